// Gated regression readout (sparse:220-231, dense:119-129) for one task, fused:
//   val[v]  = sigmoid([h_T[v] | h_0[v]] . w_gate + b_gate) * (h_T[v] . w_trans + b_trans) * mask[v]
//   out[g]  = sum of val over the nodes of graph g          (tf.unsorted_segment_sum / masked reduce_sum)
// The reference's readout MLPs have no hidden layers (chem_tensorflow.py:153-157), so each is one affine map to a scalar.
// Forward: one warp per node computes val[v] (every node row read once, float4), then one thread per graph adds its nodes in
// order (the serial order of TF's CPU segment sum, deterministic); node lists not grouped by graph take an atomicAdd stage.  Backward: one warp per node
// recomputes the two dot products, writes d h_T, accumulates the weight gradients in registers and reduces them per block.
#pragma once
#include "ggnn_common.cuh"

namespace ggnn {
namespace readout {

constexpr int MAX_D_PER_LANE = 8;   // D <= 256

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct Weights {
    const float* w_gate;    // [2D]  rows of the [2D,1] kernel: first D act on h_T, last D on h_0
    const float* b_gate;    // [1]
    const float* w_trans;   // [D]
    const float* b_trans;   // [1]
};

__device__ __forceinline__ void node_dots(const float* __restrict__ hT, const float* __restrict__ h0, const Weights& w, int D, int lane,
                                          float& gate_pre, float& trans_pre) {
    float g = 0.f, t = 0.f;
    for (int d = lane; d < D; d += 32) {
        const float a = hT[d];
        g = fmaf(a, w.w_gate[d], g);
        g = fmaf(h0[d], w.w_gate[D + d], g);
        t = fmaf(a, w.w_trans[d], t);
    }
    gate_pre = warp_sum(g) + w.b_gate[0];
    trans_pre = warp_sum(t) + w.b_trans[0];
}

// stage 1: one warp per node -> val[v]   (fully parallel; the node rows are read exactly once, 16 bytes per lane)
__global__ void __launch_bounds__(256) readout_node_kernel(const float* __restrict__ h_last, const float* __restrict__ h0, Weights w,
                                                           const float* __restrict__ mask, float* __restrict__ val, int V, int D) {
    const int v = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (v >= V) return;
    const float4* hT = reinterpret_cast<const float4*>(h_last + (size_t)v * D);
    const float4* hz = reinterpret_cast<const float4*>(h0 + (size_t)v * D);
    const float4* wa = reinterpret_cast<const float4*>(w.w_gate);
    const float4* wb = reinterpret_cast<const float4*>(w.w_gate + D);
    const float4* wt = reinterpret_cast<const float4*>(w.w_trans);
    float g = 0.f, t = 0.f;
    for (int q = lane; q < (D >> 2); q += 32) {
        const float4 a = hT[q], z = hz[q], ga = wa[q], gb = wb[q], tt = wt[q];
        g += a.x * ga.x + a.y * ga.y + a.z * ga.z + a.w * ga.w + z.x * gb.x + z.y * gb.y + z.z * gb.z + z.w * gb.w;
        t += a.x * tt.x + a.y * tt.y + a.z * tt.z + a.w * tt.w;
    }
    g = warp_sum(g) + w.b_gate[0];
    t = warp_sum(t) + w.b_trans[0];
    float r = sigmoidf_acc(g) * t;
    if (mask) r *= mask[v];
    if (lane == 0) val[v] = r;
}
// stage 2, graphs grouped: graph g owns nodes [graph_start[g], graph_start[g+1]); summed in node order (deterministic, the order of
// TF's CPU unsorted_segment_sum)
__global__ void __launch_bounds__(128) readout_sum_grouped_kernel(const float* __restrict__ val, const int* __restrict__ graph_start,
                                                                  float* __restrict__ out, int G) {
    const int g = blockIdx.x * 128 + threadIdx.x;
    if (g >= G) return;
    float acc = 0.f;
    for (int v = graph_start[g]; v < graph_start[g + 1]; ++v) acc += val[v];
    out[g] = acc;
}
// stage 2, arbitrary graph_of[v]: out must be zeroed by the caller
__global__ void __launch_bounds__(256) readout_sum_atomic_kernel(const float* __restrict__ val, const int* __restrict__ graph_of,
                                                                 float* __restrict__ out, int V) {
    const int v = blockIdx.x * 256 + threadIdx.x;
    if (v < V) atomicAdd(out + graph_of[v], val[v]);
}

// d_out[G] -> d_h_last[V,D] (written), d_w_gate[2D] / d_b_gate[1] / d_w_trans[D] / d_b_trans[1] (accumulated, atomics per block)
__global__ void __launch_bounds__(256) readout_bwd_kernel(const float* __restrict__ h_last, const float* __restrict__ h0, Weights w,
                                                          const int* __restrict__ graph_of, const float* __restrict__ mask,
                                                          const float* __restrict__ d_out, float* __restrict__ d_h_last,
                                                          float* __restrict__ d_w_gate, float* __restrict__ d_b_gate,
                                                          float* __restrict__ d_w_trans, float* __restrict__ d_b_trans, int V, int D) {
    __shared__ float red[3 * 256 + 2];   // [d_w_gate(h_T part) | d_w_gate(h_0 part) | d_w_trans] for d < 256, then the two biases
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 3 * 256 + 2; i += 256) red[i] = 0.f;
    __syncthreads();
    float gw_a[MAX_D_PER_LANE], gw_b[MAX_D_PER_LANE], tw[MAX_D_PER_LANE];
#pragma unroll
    for (int j = 0; j < MAX_D_PER_LANE; ++j) gw_a[j] = gw_b[j] = tw[j] = 0.f;
    float gb = 0.f, tb = 0.f;
    for (int v = blockIdx.x * 8 + warp; v < V; v += gridDim.x * 8) {
        const float* hT = h_last + (size_t)v * D;
        const float* hz = h0 + (size_t)v * D;
        float gp, tp;
        node_dots(hT, hz, w, D, lane, gp, tp);
        const float go = d_out[graph_of[v]] * (mask ? mask[v] : 1.0f);
        const float g = sigmoidf_acc(gp);
        const float dgp = go * tp * g * (1.0f - g);   // d gate pre-activation
        const float dt = go * g;                      // d transform output
#pragma unroll
        for (int j = 0; j < MAX_D_PER_LANE; ++j) {
            const int d = lane + 32 * j;
            if (d < D) {
                const float a = hT[d];
                d_h_last[(size_t)v * D + d] = fmaf(dgp, w.w_gate[d], dt * w.w_trans[d]);
                gw_a[j] = fmaf(dgp, a, gw_a[j]);
                gw_b[j] = fmaf(dgp, hz[d], gw_b[j]);
                tw[j] = fmaf(dt, a, tw[j]);
            }
        }
        gb += dgp; tb += dt;
    }
#pragma unroll
    for (int j = 0; j < MAX_D_PER_LANE; ++j) {
        const int d = lane + 32 * j;
        if (d < D) { atomicAdd(&red[d], gw_a[j]); atomicAdd(&red[256 + d], gw_b[j]); atomicAdd(&red[512 + d], tw[j]); }
    }
    if (lane == 0) { atomicAdd(&red[768], gb); atomicAdd(&red[769], tb); }
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += 256) {
        if (d_w_gate) { atomicAdd(d_w_gate + d, red[d]); atomicAdd(d_w_gate + D + d, red[256 + d]); }
        if (d_w_trans) atomicAdd(d_w_trans + d, red[512 + d]);
    }
    if (threadIdx.x == 0) {
        if (d_b_gate) atomicAdd(d_b_gate, red[768]);
        if (d_b_trans) atomicAdd(d_b_trans, red[769]);
    }
}

// Masked regression loss of one task (chem_tensorflow.py:161-166): diff = (computed - target) * mask,
//   loss = sum 0.5*diff^2 / (sum mask + 1e-7),  accuracy (= MAE) = sum |diff| / (sum mask + 1e-7).
// One block per task; every thread adds a strided slice in index order and the block reduces in a fixed tree: run-to-run identical.
// out[task] = loss, out[num_tasks + task] = accuracy.
__global__ void __launch_bounds__(256) masked_loss_kernel(const float* __restrict__ computed, const float* __restrict__ target_values,
                                                          const float* __restrict__ target_mask, float* __restrict__ out, int G, int num_tasks) {
    __shared__ float s_sq[256], s_abs[256], s_cnt[256];
    const int task = blockIdx.x, tid = threadIdx.x;
    const float* c = computed + (size_t)task * G;
    const float* tv = target_values + (size_t)task * G;
    const float* tm = target_mask + (size_t)task * G;
    float sq = 0.f, ab = 0.f, cnt = 0.f;
    for (int g = tid; g < G; g += 256) {
        const float m = tm[g], d = (c[g] - tv[g]) * m;
        sq += 0.5f * d * d; ab += fabsf(d); cnt += m;
    }
    s_sq[tid] = sq; s_abs[tid] = ab; s_cnt[tid] = cnt;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) { s_sq[tid] += s_sq[tid + o]; s_abs[tid] += s_abs[tid + o]; s_cnt[tid] += s_cnt[tid + o]; }
        __syncthreads();
    }
    if (tid == 0) {
        const float num = s_cnt[0] + 1e-7f;   // SMALL_NUMBER, utils.py:8
        out[task] = s_sq[0] / num;
        out[num_tasks + task] = s_abs[0] / num;
    }
}

}  // namespace readout
}  // namespace ggnn
