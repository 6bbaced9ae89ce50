// Gated regression readout (sparse:220-231, dense:119-129) for one task, fused:
//   val[v]  = sigmoid([h_T[v] | h_0[v]] . w_gate + b_gate) * (h_T[v] . w_trans + b_trans) * mask[v]
//   out[g]  = sum of val over the nodes of graph g          (tf.unsorted_segment_sum / masked reduce_sum)
// The reference's readout MLPs have no hidden layers (chem_tensorflow.py:153-157), so each is one affine map to a scalar.
// Forward: one warp per graph walks the graph's nodes in order (the serial order of TF's CPU segment sum, deterministic);
// node lists that are not grouped by graph take the one-warp-per-node + atomicAdd variant.  Backward: one warp per node
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

// graphs grouped: graph g owns nodes [graph_start[g], graph_start[g+1])
__global__ void __launch_bounds__(256) readout_fwd_grouped_kernel(const float* __restrict__ h_last, const float* __restrict__ h0, Weights w,
                                                                  const int* __restrict__ graph_start, const float* __restrict__ mask,
                                                                  float* __restrict__ out, int G, int D) {
    const int g = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (g >= G) return;
    float acc = 0.f;
    for (int v = graph_start[g]; v < graph_start[g + 1]; ++v) {
        float gp, tp;
        node_dots(h_last + (size_t)v * D, h0 + (size_t)v * D, w, D, lane, gp, tp);
        float val = sigmoidf_acc(gp) * tp;
        if (mask) val *= mask[v];
        acc += val;
    }
    if (lane == 0) out[g] = acc;
}

// arbitrary graph_of[v]: out must be zeroed by the caller
__global__ void __launch_bounds__(256) readout_fwd_atomic_kernel(const float* __restrict__ h_last, const float* __restrict__ h0, Weights w,
                                                                 const int* __restrict__ graph_of, const float* __restrict__ mask,
                                                                 float* __restrict__ out, int V, int D) {
    const int v = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (v >= V) return;
    float gp, tp;
    node_dots(h_last + (size_t)v * D, h0 + (size_t)v * D, w, D, lane, gp, tp);
    float val = sigmoidf_acc(gp) * tp;
    if (mask) val *= mask[v];
    if (lane == 0) atomicAdd(out + graph_of[v], val);
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

}  // namespace readout
}  // namespace ggnn
