// Streaming GGNN propagation on 5th-gen tensor cores (tcgen05 / TMEM), sm_100a: the path for hidden sizes whose
// per-tile operands do not fit one SM (D > 128, BASELINE config 4) and for batches / graphs too large for the
// tile-local fused kernel (ggnn_fwd_tc.cuh).
//
// One timestep (sparse:153-216) = three launches of ONE kernel template, each a 128-row x NC-column output tile
// per CTA whose K dimension is streamed through a shared-memory ring, one UMMA K-step (16 columns) per stage:
//   EPI_AGG   agg       = [A_0 | .. | A_{T-1}] . [W_0; ..; W_{T-1}]      A_t[v] = sum of h[src] over the type-t messages into v,
//                         + indeg.B, / (deg + 1e-7)                       gathered by the worker warps straight into the ring (GATHER)
//   EPI_GATE  [r | u]   = sigmoid([res.. | agg | h] . K_g + b_g)          writes r*h (operand image) and u
//   EPI_CAND  h'        = u*h + (1-u)*act([res.. | agg | r*h] . K_c + b_c) (RNN: act([res.. | agg | h] . K + b))
// Node-state operands live in HBM/L2 as bf16 hi/lo "images" in the canonical K-major no-swizzle UMMA layout, tile-major:
//   byte(tile, kstep, part, kgroup, row, j) = ((tile*NKS + kstep)*2 + part)*4096 + kgroup*2048 + row*16 + j*2
// so one K-step of a 128-row A operand (hi + lo) is ONE contiguous 8 KB bulk copy (cp.async.bulk, 1-D TMA), and an
// epilogue thread (= one TMEM lane = one row) writes 16-byte chunks that are contiguous across the warp.  Weights are
// pre-split and pre-tiled per (N block, K-step) into contiguous 64*NC-byte stages (ggnn_tile_weights_stream_kernel).
// fp32 accuracy on bf16 tensor cores as in ggnn_fwd_tc.cuh: x = hi + lo, product = Ah.Bh + Ah.Bl + Al.Bh (3 MMAs).
//
// Roles: warp 0 = producer (one thread: bulk copies), warp 1 = MMA issuer (+ TMEM allocator), warps 2.. = workers
// (gather groups of 4 warps in the edge kernel; epilogue: TMEM lane quarter = warp % 4, column group = worker / 4).
// The TMA-fed variants use <= ~100 KB of shared memory and <= 256 TMEM columns so that two CTAs share an SM: one
// CTA's epilogue runs under the other's MMAs.  Every mbarrier wait is bounded; on timeout an error code is written.
#pragma once
#include "ggnn_fwd_tc.cuh"

namespace ggnn {
namespace ts {

using tc::smem_u32;

constexpr int TILE_M = 128;
constexpr int MAX_SEG = MAX_RES + 2;
constexpr int MAX_NS = 8;            // ring stages
constexpr int A_STAGE_B = 8192;      // one K-step of a 128-row A operand: 2 parts x 2 k-groups x 128 rows x 16 B
enum { EPI_AGG = 0, EPI_GATE = 1, EPI_CAND = 2 };

struct StreamParams {
    int V, D, DP, T;
    int NC;                // output columns per CTA (multiple of 16, <= 256); grid.y = number of N blocks
    int nstages;           // ring depth
    int nparts;            // 3: bf16x3, 1: single bf16 MMA
    int tmem_cols;         // power of two >= max(32, NC)
    int epi;               // EPI_*
    int cell, act, use_bias, use_avg;
    // ---- A operand, TMA-fed: nseg K segments, each a DP-wide tile-major image
    int nseg;
    const uint8_t* seg[MAX_SEG];
    // ---- A operand, gathered (EPI_AGG): per present edge type a DP-wide segment of per-type source sums
    const float* g_src;          // fp32 state the messages are gathered from [V][D]
    const int* row_ptr;          // [V*T+1] target-keyed CSR
    const int* csr_src;          // [M]
    const unsigned* tile_mask;   // [ntiles] bit t: some row of the tile receives a type-t message
    int csr_cap;                 // capacity (ints) of the shared copy of the tile's source list; larger tiles read it from L2
    // ---- B operand
    const uint8_t* w;            // [nblk][kt_all] stages of 64*NC bytes: [hi: 2 k-groups x NC x 16 B | lo: same]
    int kt_all;                  // K-steps per N block in `w`
    // ---- epilogue
    const float* bias;           // AGG: edge_biases [T][D] or null; GATE: gate_bias [2D]; CAND: cand_bias [D]
    const float* indeg;          // [V][T]
    const float* denom;          // [V]
    const float* h_in;           // fp32 state entering the step [V][D]   (GATE: r*h, save; CAND: blend)
    float* u_buf;                // [V][D] update gate: written by GATE, read by CAND
    float* h_out;                // CAND: new state fp32 [V][D]
    uint8_t* img_out;            // AGG: agg image; GATE: r*h image; CAND: image of the new state
    float* sv_h; float* sv_agg; float* sv_r; float* sv_c;   // this step's save-for-backward slots or null
    float drop_keep; unsigned long long drop_seed; int gstep;
    int* error_flag;
};

__device__ __forceinline__ size_t img_tile_bytes(int NKS) { return (size_t)NKS * A_STAGE_B; }

// one [row, 8 columns] chunk of an image (both parts); col0 % 8 == 0
__device__ __forceinline__ void img_store_chunk(uint8_t* img, int NKS, int tile, int row, int col0, const float (&x)[8]) {
    uint4 hi, lo;
    tc::split8(x, hi, lo);
    uint8_t* p = img + ((size_t)tile * NKS + (col0 >> 4)) * A_STAGE_B + (size_t)((col0 >> 3) & 1) * 2048 + (size_t)row * 16;
    *reinterpret_cast<uint4*>(p) = hi;
    *reinterpret_cast<uint4*>(p + 4096) = lo;
}

template <int NWORK, bool GATHER>
__global__ void __launch_bounds__((NWORK + 2) * 32, GATHER ? 1 : 2) ggnn_stream_kernel(const __grid_constant__ StreamParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[MAX_NS];    // B (and TMA-fed A) bytes landed
    __shared__ __align__(8) uint64_t bar_afull[MAX_NS];   // gathered A written (GATHER)
    __shared__ __align__(8) uint64_t bar_empty[MAX_NS];   // the MMAs that read the stage are complete
    __shared__ __align__(8) uint64_t bar_acc;             // all MMAs of the tile are complete
    __shared__ uint32_t s_tmem;
    __shared__ int s_abort;
    __shared__ int s_types[32];
    __shared__ int s_ntypes;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x, nb = blockIdx.y;
    const int D = p.D, DP = p.DP, T = p.T, NC = p.NC, NS = p.nstages;
    const int NKS = DP >> 4;
    const int row0 = tile * TILE_M;
    const int rows = min(TILE_M, p.V - row0);
    const uint32_t B_STAGE_B = 64u * (uint32_t)NC;
    const uint32_t STAGE_B = (uint32_t)A_STAGE_B + B_STAGE_B;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    int* sPtr = reinterpret_cast<int*>(smem + (size_t)NS * STAGE_B);    // [128*T + 1] tile-relative CSR row offsets (GATHER)
    int* sSrc = sPtr + ((TILE_M * T + 1 + 3) & ~3);                     // [csr_cap] global source ids

    if (tid == 0) {
        s_abort = 0;
        for (int i = 0; i < MAX_NS; ++i) { tc::mbar_init(&bar_full[i], 1); tc::mbar_init(&bar_afull[i], 4); tc::mbar_init(&bar_empty[i], 1); }
        tc::mbar_init(&bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        int n = 0;
        if (GATHER) {
            const unsigned m = p.tile_mask[tile];
            for (int t = 0; t < T; ++t) if ((m >> t) & 1u) s_types[n++] = t;
        }
        s_ntypes = n;
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = s_tmem;
    volatile int* abortp = &s_abort;
    const int nk = GATHER ? s_ntypes * NKS : p.nseg * NKS;   // K-steps of this tile

    if (warp == 0) {
        // =============================================================================== PRODUCER (one thread)
        if (lane == 0) {
            bool ok = true;
            const uint8_t* wb = p.w + (size_t)nb * p.kt_all * B_STAGE_B;
            for (int k = 0; k < nk && ok; ++k) {
                const int s = k % NS, it = k / NS;
                if (it > 0 && !tc::mbar_wait(&bar_empty[s], (uint32_t)(it - 1) & 1u, abortp)) { ok = false; break; }
                uint8_t* st = smem + (size_t)s * STAGE_B;
                if (GATHER) {
                    const int kk = s_types[k / NKS] * NKS + (k % NKS);
                    tc::mbar_arrive_expect_tx(&bar_full[s], B_STAGE_B);
                    tc::bulk_copy_g2s(st + A_STAGE_B, wb + (size_t)kk * B_STAGE_B, B_STAGE_B, &bar_full[s]);
                } else {
                    const int sg = k / NKS, ks = k - sg * NKS;
                    tc::mbar_arrive_expect_tx(&bar_full[s], (uint32_t)A_STAGE_B + B_STAGE_B);
                    tc::bulk_copy_g2s(st, p.seg[sg] + ((size_t)tile * NKS + ks) * A_STAGE_B, A_STAGE_B, &bar_full[s]);
                    tc::bulk_copy_g2s(st + A_STAGE_B, wb + (size_t)k * B_STAGE_B, B_STAGE_B, &bar_full[s]);
                }
            }
            if (!ok) atomicExch(p.error_flag, 13);
        }
    } else if (warp == 1) {
        // =============================================================================== MMA ISSUER
        bool ok = true;
        const bool x3 = p.nparts == 3;
        const uint64_t descA = tc::make_desc(0, 2048, 128);
        const uint64_t descB = tc::make_desc(0, 16u * (uint32_t)NC, 128);
        const uint32_t idesc = tc::make_idesc_bf16(NC);
        const uint32_t b_lo16 = (32u * (uint32_t)NC) >> 4;
        const uint32_t smem16 = smem_u32(smem) >> 4, stage16 = STAGE_B >> 4;
        const uint32_t tm_d = __shfl_sync(0xffffffffu, tmem, 0);
        for (int k = 0; k < nk && ok; ++k) {
            const int s = k % NS;
            const uint32_t par = (uint32_t)(k / NS) & 1u;
            if (!tc::mbar_wait(&bar_full[s], par, abortp)) ok = false;
            if (GATHER && ok && !tc::mbar_wait(&bar_afull[s], par, abortp)) ok = false;
            ok = __all_sync(0xffffffffu, ok);
            if (!ok) break;
            tc::tc_fence_after();
            if (tc::elect_one()) {
                const uint32_t a16 = smem16 + (uint32_t)s * stage16, b16 = a16 + (A_STAGE_B >> 4);
                const uint64_t ah = descA | (uint64_t)a16, al = descA | (uint64_t)(a16 + (4096u >> 4));
                const uint64_t bh = descB | (uint64_t)b16, bl = descB | (uint64_t)(b16 + b_lo16);
                tc::umma_bf16(tm_d, ah, bh, idesc, k > 0 ? 1u : 0u);
                if (x3) {
                    tc::umma_bf16(tm_d, ah, bl, idesc, 1u);
                    tc::umma_bf16(tm_d, al, bh, idesc, 1u);
                }
                tc::umma_commit(&bar_empty[s]);
                if (k == nk - 1) tc::umma_commit(&bar_acc);
            }
            __syncwarp();
        }
        if (!ok && lane == 0) atomicExch(p.error_flag, 12);
    } else {
        // =============================================================================== WORKERS
        const int wi = warp - 2;
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;                // tile row == TMEM lane
        const bool row_ok = row < rows;
        const int grow = row0 + (row_ok ? row : 0);
        bool ok = true;
        if (GATHER) {
            constexpr int NG = NWORK / 4;             // gather groups (4 warps = 128 rows each)
            const int grp = wi >> 2;
            // ---- the tile's CSR slice -> shared memory
            const int base = p.row_ptr[(size_t)row0 * T];
            const int nptr = rows * T + 1;
            const int wt = tid - 64;
            for (int i = wt; i < TILE_M * T + 1; i += NWORK * 32) sPtr[i] = p.row_ptr[(size_t)row0 * T + min(i, nptr - 1)] - base;
            const int mt = p.row_ptr[(size_t)(row0 + rows) * T] - base;
            const bool cached = mt <= p.csr_cap;
            if (cached) for (int i = wt; i < mt; i += NWORK * 32) sSrc[i] = p.csr_src[base + i];
            asm volatile("bar.sync 1, %0;" ::"n"(NWORK * 32) : "memory");
            const int* srcs = cached ? sSrc : p.csr_src + base;
            for (int k = grp; k < nk && ok; k += NG) {
                const int s = k % NS, it = k / NS;
                const int t = s_types[k / NKS], col0 = (k % NKS) * 16;
                int beg = 0, end = 0;
                if (row_ok) { beg = sPtr[row * T + t]; end = sPtr[row * T + t + 1]; }
                float a[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) a[j] = 0.0f;
                for (int m = beg; m < end; m += 2) {   // two source rows in flight
                    const bool two = m + 1 < end;
                    const float* s0 = p.g_src + (size_t)srcs[m] * D;
                    const float* s1 = p.g_src + (size_t)srcs[two ? m + 1 : m] * D;
                    float v0[16], v1[16];
                    tc::load8_guarded_cg(s0, col0, D, *reinterpret_cast<float(*)[8]>(&v0[0]));
                    tc::load8_guarded_cg(s0, col0 + 8, D, *reinterpret_cast<float(*)[8]>(&v0[8]));
                    tc::load8_guarded_cg(s1, col0, two ? D : 0, *reinterpret_cast<float(*)[8]>(&v1[0]));
                    tc::load8_guarded_cg(s1, col0 + 8, two ? D : 0, *reinterpret_cast<float(*)[8]>(&v1[8]));
#pragma unroll
                    for (int j = 0; j < 16; ++j) a[j] += v0[j];
#pragma unroll
                    for (int j = 0; j < 16; ++j) a[j] += v1[j];
                }
                uint4 h0, l0, h1, l1;
                tc::split8(*reinterpret_cast<float(*)[8]>(&a[0]), h0, l0);
                tc::split8(*reinterpret_cast<float(*)[8]>(&a[8]), h1, l1);
                if (it > 0 && !tc::mbar_wait(&bar_empty[s], (uint32_t)(it - 1) & 1u, abortp)) { ok = false; }
                ok = __all_sync(0xffffffffu, ok);
                if (!ok) break;
                uint8_t* ap = smem + (size_t)s * STAGE_B + (size_t)row * 16;
                *reinterpret_cast<uint4*>(ap) = h0;
                *reinterpret_cast<uint4*>(ap + 2048) = h1;
                *reinterpret_cast<uint4*>(ap + 4096) = l0;
                *reinterpret_cast<uint4*>(ap + 6144) = l1;
                tc::fence_async_smem();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bar_afull[s]);
            }
        }
        // ---- epilogue: accumulator -> registers -> outputs
        if (ok && nk > 0) {
            if (!tc::mbar_wait(&bar_acc, 0, abortp)) ok = false;
            ok = __all_sync(0xffffffffu, ok);
            tc::tc_fence_after();
        }
        if (ok) {
            constexpr int NCG = NWORK / 4;            // column groups
            const int cgp = wi >> 2;
            const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
            const bool have_acc = nk > 0;
            const int nchunks = NC >> 3;
            if (p.epi == EPI_AGG) {
                const float den = (p.use_avg && row_ok) ? p.denom[grow] : 1.0f;
                for (int c = cgp; c < nchunks; c += NCG) {
                    const int col = nb * NC + c * 8;
                    if (col >= DP) break;
                    float v[8];
                    if (have_acc) tc::tmem_ld8(tmem + lane_addr + c * 8, v);
                    else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = 0.0f;
                    }
                    if (p.use_bias && row_ok) {
                        for (int t = 0; t < T; ++t) {
                            const float ind = p.indeg[(size_t)grow * T + t];
                            float b[8];
                            tc::load8_guarded(p.bias + (size_t)t * D, col, D, b);
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = fmaf(ind, b[j], v[j]);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = row_ok ? (p.use_avg ? __fdiv_rn(v[j], den) : v[j]) : 0.0f;   // sparse:207-209 divides
                    if (p.sv_agg && row_ok) tc::store8_guarded(p.sv_agg + (size_t)grow * D, col, D, v);
                    img_store_chunk(p.img_out, NKS, tile, row, col, v);
                }
            } else if (p.epi == EPI_GATE) {
                for (int c = cgp; c < nchunks; c += NCG) {
                    const int colp = nb * NC + c * 8;
                    if (colp >= 2 * DP) break;
                    const bool is_r = colp < DP;
                    const int col = is_r ? colp : colp - DP;
                    float g[8], b[8];
                    tc::tmem_ld8_nowait(tmem + lane_addr + c * 8, g);
                    tc::load8_guarded(p.bias + (is_r ? 0 : D), col, D, b);
                    if (is_r) {
                        float h[8], rh[8];
                        tc::load8_guarded_cg(p.h_in + (size_t)grow * D, col, row_ok ? D : 0, h);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 8; ++j) { g[j] = tc::sigmoid_fast(g[j] + b[j]); rh[j] = g[j] * h[j]; }
                        if (p.sv_r && row_ok) {
                            tc::store8_guarded(p.sv_r + (size_t)grow * D, col, D, g);
                            tc::store8_guarded(p.sv_h + (size_t)grow * D, col, D, h);
                        }
                        img_store_chunk(p.img_out, NKS, tile, row, col, rh);
                    } else {
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 8; ++j) g[j] = tc::sigmoid_fast(g[j] + b[j]);
                        if (row_ok) tc::store8_guarded(p.u_buf + (size_t)grow * D, col, D, g);
                    }
                }
            } else {
                const bool gru = p.cell == CELL_GRU;
                for (int c = cgp; c < nchunks; c += NCG) {
                    const int col = nb * NC + c * 8;
                    if (col >= DP) break;
                    float cv[8], b[8], h[8], u[8], hn[8];
                    tc::tmem_ld8_nowait(tmem + lane_addr + c * 8, cv);
                    tc::load8_guarded(p.bias, col, D, b);
                    if (gru || p.sv_h) tc::load8_guarded_cg(p.h_in + (size_t)grow * D, col, row_ok ? D : 0, h);
                    if (gru) tc::load8_guarded_cg(p.u_buf + (size_t)grow * D, col, row_ok ? D : 0, u);
                    tc::tmem_ld_wait();
                    if (gru) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            cv[j] = tc::act_fast(cv[j] + b[j], p.act);
                            hn[j] = fmaf(u[j], h[j] - cv[j], cv[j]);   // u*h + (1-u)*c
                        }
                        if (p.sv_c && row_ok) tc::store8_guarded(p.sv_c + (size_t)grow * D, col, D, cv);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) hn[j] = tc::act_fast(cv[j] + b[j], p.act);
                        if (p.sv_h && row_ok) tc::store8_guarded(p.sv_h + (size_t)grow * D, col, D, h);
                    }
                    if (p.drop_keep < 1.0f) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) hn[j] = dropout_apply(hn[j], p.drop_seed, p.gstep, p.V, D, grow, col + j, p.drop_keep);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) hn[j] = (row_ok && col + j < D) ? hn[j] : 0.0f;
                    if (row_ok) tc::store8_guarded(p.h_out + (size_t)grow * D, col, D, hn);
                    img_store_chunk(p.img_out, NKS, tile, row, col, hn);
                }
            }
        }
        if (!ok && lane == 0) atomicExch(p.error_flag, 11);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols) : "memory");
    }
}

// fp32 [V][D] row-major -> tile-major bf16 hi/lo image ([ntiles*128][DP], zero padded)
__global__ void ggnn_image_kernel(const float* __restrict__ x, uint8_t* __restrict__ img, int V, int D, int DP, int ntiles) {
    const int NKC = DP >> 3, NKS = DP >> 4;
    const long long total = (long long)ntiles * TILE_M * NKC;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        // consecutive threads = consecutive rows of one chunk column: coalesced image writes
        const int row = (int)(idx % TILE_M);
        const int kc = (int)((idx / TILE_M) % NKC);
        const int tile = (int)(idx / ((long long)TILE_M * NKC));
        const int grow = tile * TILE_M + row;
        float v[8];
        tc::load8_guarded(x + (size_t)(grow < V ? grow : 0) * D, kc * 8, grow < V ? D : 0, v);
        img_store_chunk(img, NKS, tile, row, kc * 8, v);
    }
}

// Weight pre-tiling for the streaming kernel.  Source: fp32 row-major W[(nseg*D) rows][src_ld cols]; the padded operand has
// K = nseg*DP rows (segment s, row kk < D -> source row s*D + kk) and N = nblk*NC columns, where padded column n maps to source
// column (n / DP)*D + n % DP when n % DP < D and n / DP < ncolblk (the [r | u] gate kernel has two D-wide column blocks), else zero.
//   out: [nblk][K/16] stages of 64*NC bytes:  byte(part, kg, n, j) = part*32*NC + kg*16*NC + n*16 + j*2
__global__ void ggnn_tile_weights_stream_kernel(const float* __restrict__ W, uint8_t* __restrict__ out, int D, int DP, int nseg, int ncolblk,
                                                int src_ld, int NC, int nblk) {
    const int kt = nseg * DP / 16;
    const long long total = (long long)nblk * kt * 2 * NC;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(idx % NC);
        const int kg = (int)((idx / NC) % 2);
        const int k = (int)((idx / (2 * NC)) % kt);
        const int nb = (int)(idx / ((long long)2 * NC * kt));
        const int np = nb * NC + n;
        const int cb = np / DP, nn = np - cb * DP;
        const bool col_ok = cb < ncolblk && nn < D;
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int kp = k * 16 + kg * 8 + j;
            const int sg = kp / DP, kk = kp - sg * DP;
            x[j] = (col_ok && kk < D) ? W[(size_t)(sg * D + kk) * src_ld + cb * D + nn] : 0.0f;
        }
        uint4 hi, lo;
        tc::split8(x, hi, lo);
        uint8_t* base = out + ((size_t)nb * kt + k) * 64 * NC + (size_t)kg * 16 * NC + (size_t)n * 16;
        *reinterpret_cast<uint4*>(base) = hi;
        *reinterpret_cast<uint4*>(base + (size_t)32 * NC) = lo;
    }
}

}  // namespace ts
}  // namespace ggnn
