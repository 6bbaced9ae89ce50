// Backward of the GGNN propagation step (what optimizer.compute_gradients builds for sparse:117-218 / dense:93-117,
// chem_tensorflow.py:184), fp32 on CUDA cores.  Per timestep, in reverse, with the activations the forward saved
// (state entering the step h, aggregated messages x, gates r/u, candidate c):
//
//   GRU :  dc = dh'*(1-u)   du = dh'*(h-c)   dh = dh'*u
//          dpc = dc*act'(c)        d[res..,x,rh] = dpc . K_c^T     dK_c += [res..,x,rh]^T . dpc     db_c += sum dpc
//          dr = drh*h   dh += drh*r   dpr = dr*r(1-r)   dpu = du*u(1-u)
//          d[res..,x,h] += [dpr|dpu] . K_g^T                       dK_g += [res..,x,h]^T . [dpr|dpu]  db_g += sum
//   RNN :  dpc = dh'*act'(h')      d[res..,x,h] = dpc . K^T        dK += [res..,x,h]^T . dpc          db += sum dpc
//   msgs:  dx' = dx / (deg+1e-7)   dB[t] += sum_v indeg[v,t] dx'[v]
//          dW_t += A_t^T . dx'   (A_t = per-type gathered source states, recomputed from the target CSR)
//          dh   += G_t . W_t^T   (G_t[s] = sum of dx'[target] over the type-t messages LEAVING s: source CSR)
//
// Kernels: elementwise cell gradients, CSR gathers, a 64x64-tile FFMA GEMM  C (+)= A . B^T  for the data gradients and a
// split-row  C += A^T . B  with fp32 atomics for the weight gradients.  Clarity first; the forward is the hot path.
#pragma once
#include "ggnn_common.cuh"

namespace ggnn {
namespace bwd {

// ---------------------------------------------------------------- C[M,N] (+)= A[M,K] . B[N,K]^T
template <bool ACC>
__global__ void __launch_bounds__(256) gemm_nt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                                      float* __restrict__ C, int ldc, int M, int N, int K) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int idx = threadIdx.x; idx < 64 * 16; idx += 256) {
            const int r = idx >> 4, kk = idx & 15;
            As[kk][r] = (m0 + r < M && k0 + kk < K) ? A[(size_t)(m0 + r) * lda + k0 + kk] : 0.f;
            Bs[kk][r] = (n0 + r < N && k0 + kk < K) ? B[(size_t)(n0 + r) * ldb + k0 + kk] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < M && n < N) {
                float* c = C + (size_t)m * ldc + n;
                *c = ACC ? (*c + acc[i][j]) : acc[i][j];
            }
        }
}

// ---------------------------------------------------------------- C[K,N] += A[M,K]^T . B[M,N]   (rows split over blockIdx.z, fp32 atomics)
__global__ void __launch_bounds__(256) gemm_tn_atomic_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                                             float* __restrict__ C, int ldc, int M, int N, int K, int rows_per_split) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int k0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int mb = blockIdx.z * rows_per_split, me = min(M, mb + rows_per_split);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int m0 = mb; m0 < me; m0 += 16) {
        for (int idx = threadIdx.x; idx < 16 * 64; idx += 256) {
            const int mm = idx >> 6, c = idx & 63;
            As[mm][c] = (m0 + mm < me && k0 + c < K) ? A[(size_t)(m0 + mm) * lda + k0 + c] : 0.f;
            Bs[mm][c] = (m0 + mm < me && n0 + c < N) ? B[(size_t)(m0 + mm) * ldb + n0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int mm = 0; mm < 16; ++mm) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[mm][ty * 4 + i]; b[i] = Bs[mm][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (k < K && n < N) atomicAdd(C + (size_t)k * ldc + n, acc[i][j]);
        }
}

// ---------------------------------------------------------------- column sums: dst[n] += sum_m src[m, n]  (optionally weighted by w[m*wstride])
__global__ void __launch_bounds__(256) colsum_atomic_kernel(const float* __restrict__ src, int ld, const float* __restrict__ w, int wstride,
                                                            float* __restrict__ dst, int M, int N, int rows_per_block) {
    const int n = blockIdx.x * 256 + threadIdx.x;
    if (n >= N) return;
    const int mb = blockIdx.y * rows_per_block, me = min(M, mb + rows_per_block);
    float s = 0.f;
    for (int m = mb; m < me; ++m) s = fmaf(w ? w[(size_t)m * wstride] : 1.0f, src[(size_t)m * ld + n], s);
    atomicAdd(dst + n, s);
}

// ---------------------------------------------------------------- elementwise cell gradients
__device__ __forceinline__ float act_grad_from_output(float y, int act) { return act == ACT_TANH ? (1.0f - y * y) : (y > 0.0f ? 1.0f : 0.0f); }

// GRU part 1: dpc = dh'(1-u) act'(c) ; dpg[:, D:2D] = dh'(h-c) u(1-u) ; dh = dh' u ; rh = r h
__global__ void gru_bwd1_kernel(const float* __restrict__ dhn, const float* __restrict__ h, const float* __restrict__ r, const float* __restrict__ u,
                                const float* __restrict__ c, float* __restrict__ dpc, float* __restrict__ dpg, float* __restrict__ dh,
                                float* __restrict__ rh, long long n, int D, int act) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / D;
        const int col = (int)(i - row * D);
        const float g = dhn[i], uu = u[i], cc = c[i], hh = h[i];
        dpc[i] = g * (1.0f - uu) * act_grad_from_output(cc, act);
        dpg[row * 2 * D + D + col] = g * (hh - cc) * uu * (1.0f - uu);
        dh[i] = g * uu;
        rh[i] = r[i] * hh;
    }
}
// GRU part 2: drh = dXc[:, rh segment] ; dh += drh r ; dpg[:, 0:D] = drh h r(1-r)
__global__ void gru_bwd2_kernel(const float* __restrict__ dXc, int ldx, int rh_off, const float* __restrict__ h, const float* __restrict__ r,
                                float* __restrict__ dpg, float* __restrict__ dh, long long n, int D) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / D;
        const int col = (int)(i - row * D);
        const float drh = dXc[row * ldx + rh_off + col], rr = r[i];
        dh[i] += drh * rr;
        dpg[row * 2 * D + col] = drh * h[i] * rr * (1.0f - rr);
    }
}
// RNN: dpc = dh' act'(h')   (yscale = keep_prob undoes the state dropout's 1/keep on the saved output; dropped entries have dh' = 0)
__global__ void rnn_bwd1_kernel(const float* __restrict__ dhn, const float* __restrict__ hnew, float* __restrict__ dpc, long long n, int act, float yscale) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dpc[i] = dhn[i] * act_grad_from_output(hnew[i] * yscale, act);
}
// state dropout backward, in place: d(pre-dropout state) = d(state) * mask / keep   (mask regenerated, ggnn_common.cuh)
__global__ void dropout_grad_kernel(float* __restrict__ dhn, unsigned long long seed, int gstep, int V, int D, float keep, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / D), col = (int)(i - (long long)row * D);
        dhn[i] = dropout_apply(dhn[i], seed, gstep, V, D, row, col, keep);
    }
}
// Split the gradient of the cell input row [res_0 .. res_{R-1} | x | h-or-rh]:
//   dres_i[v] += dX[v, i*D..]   dxp[v] = dX[v, x segment] (/ denom)   (GRU second pass / RNN: dh (+)= dX[v, last segment])
__global__ void split_input_grad_kernel(const float* __restrict__ dXa, const float* __restrict__ dXb, int ldx, int nres,
                                        float* const* __restrict__ dres, float* __restrict__ dxp, const float* __restrict__ denom,
                                        float* __restrict__ dh, int dh_from_a, int dh_from_b, int dh_accumulate, long long n, int D) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / D;
        const int col = (int)(i - row * D);
        const float* a = dXa + row * ldx;
        const float* b = dXb ? dXb + row * ldx : nullptr;
        for (int s = 0; s < nres; ++s) dres[s][i] += a[s * D + col] + (b ? b[s * D + col] : 0.0f);
        float x = a[nres * D + col] + (b ? b[nres * D + col] : 0.0f);
        if (denom) x = x / denom[row];
        dxp[i] = x;
        float hg = 0.0f;
        if (dh_from_a) hg += a[(nres + 1) * D + col];
        if (dh_from_b && b) hg += b[(nres + 1) * D + col];
        if (dh_from_a || dh_from_b) dh[i] = dh_accumulate ? dh[i] + hg : hg;
    }
}
__global__ void add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] += src[i];
}

// ---------------------------------------------------------------- gathers: out[v] = sum_{slots of row (v*T+t)} in[idx[slot]]   (one warp per node)
__global__ void __launch_bounds__(256) csr_gather_sum_kernel(const int* __restrict__ row_ptr, const int* __restrict__ idx, const float* __restrict__ in,
                                                             float* __restrict__ out, int V, int D, int T, int t) {
    const int v = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (v >= V) return;
    const int beg = row_ptr[(size_t)v * T + t], end = row_ptr[(size_t)v * T + t + 1];
    for (int c4 = lane; c4 < (D >> 2); c4 += 32) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int m = beg; m < end; ++m) {
            const float4 x = *reinterpret_cast<const float4*>(in + (size_t)idx[m] * D + (c4 << 2));
            s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
        }
        *reinterpret_cast<float4*>(out + (size_t)v * D + (c4 << 2)) = s;
    }
}
// dense adjacency [b][T][v][v]: out[g*nv+i] = sum_j A[g,t,i,j] in[g*nv+j]   (transpose: sum_j A[g,t,j,i] in[g*nv+j])
__global__ void __launch_bounds__(256) dense_gather_sum_kernel(const float* __restrict__ adj, const float* __restrict__ in, float* __restrict__ out,
                                                               int V, int D, int T, int t, int nv, int transpose) {
    const int v = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (v >= V) return;
    const int g = v / nv, i = v - g * nv;
    const float* base = adj + ((size_t)g * T + t) * nv * nv;
    for (int c4 = lane; c4 < (D >> 2); c4 += 32) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < nv; ++j) {
            const float a = transpose ? base[(size_t)j * nv + i] : base[(size_t)i * nv + j];
            if (a != 0.0f) {
                const float4 x = *reinterpret_cast<const float4*>(in + (size_t)(g * nv + j) * D + (c4 << 2));
                s.x = fmaf(a, x.x, s.x); s.y = fmaf(a, x.y, s.y); s.z = fmaf(a, x.z, s.z); s.w = fmaf(a, x.w, s.w);
            }
        }
        *reinterpret_cast<float4*>(out + (size_t)v * D + (c4 << 2)) = s;
    }
}

}  // namespace bwd
}  // namespace ggnn
