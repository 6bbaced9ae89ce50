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
// The fp32 master copy of every state (and the update gate u) is kept in a second, chunk-major layout
//   float(tile, chunk = col/8, row, j) = ((tile*NKC + chunk)*128 + row)*8 + j
// so that an epilogue warp (32 rows x 8 columns) reads and writes 1 KB contiguous instead of 32 scattered sectors (the row-per-thread
// pattern on a row-major [V, D] array costs one L1TEX tag cycle per row and instruction and made the epilogues longer than the GEMMs);
// the user-visible row-major [V, D] arrays are written only for node_states_per_layer entries and for the backward pass.
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
    const uint8_t* g_img;        // image of the state the messages are gathered from (a row with one type-t message is a 64-byte copy)
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
    const float* h_chk;          // fp32 state entering the step, chunk-major   (GATE: r*h, save; CAND: blend)
    float* u_buf;                // update gate, chunk-major: written by GATE, read by CAND
    float* h_chk_out;            // CAND: new state fp32, chunk-major
    float* h_out;                // CAND: new state fp32 row-major [V][D], or null (only node_states_per_layer entries need it)
    float* sv_u;                 // GATE: row-major copy of u for the backward pass, or null
    uint8_t* img_out;            // AGG: agg image; GATE: r*h image; CAND: image of the new state
    float* sv_h; float* sv_agg; float* sv_r; float* sv_c;   // this step's save-for-backward slots or null
    float drop_keep; unsigned long long drop_seed; int gstep;
    int* error_flag;
    long long* dbg;   // optional per-CTA phase stamps [grid.y][grid.x][16] (GGNN_TS_DEBUG=1), or nullptr
};

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ size_t chunk_off(int NKC, int tile, int c, int row) { return (((size_t)tile * NKC + c) * TILE_M + row) * 8; }
__device__ __forceinline__ void chunk_load(const float* base, int NKC, int tile, int c, int row, float (&v)[8]) {
    const float4* q = reinterpret_cast<const float4*>(base + chunk_off(NKC, tile, c, row));
    const float4 a = __ldcg(q), b = __ldcg(q + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void chunk_store(float* base, int NKC, int tile, int c, int row, const float (&v)[8]) {
    float4* q = reinterpret_cast<float4*>(base + chunk_off(NKC, tile, c, row));
    q[0] = make_float4(v[0], v[1], v[2], v[3]);
    q[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void unpack8(const uint4& a, float (&x)[8]) {
    const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { x[2 * j] = __uint_as_float(w[j] << 16); x[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u); }
}

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
    uint8_t* sPerm = reinterpret_cast<uint8_t*>(sSrc + p.csr_cap);       // [T][128] rows of the tile ordered: >= 2 type-t messages | exactly 1 | none
    __shared__ int s_n0[32], s_n1[32];                                   // per present type: rows with >= 2 messages, rows with exactly 1

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
            long long waited = 0;
            for (int k = 0; k < nk && ok; ++k) {
                const int s = k % NS, it = k / NS;
                if (it > 0) {
                    const long long w0 = p.dbg ? clock64() : 0;
                    if (!tc::mbar_wait(&bar_empty[s], (uint32_t)(it - 1) & 1u, abortp)) { ok = false; break; }
                    if (p.dbg) waited += clock64() - w0;
                }
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
            if (p.dbg) p.dbg[((size_t)nb * gridDim.x + tile) * 16 + 4] = waited;
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
        long long waited_b = 0, waited_a = 0;
        for (int k = 0; k < nk && ok; ++k) {
            const int s = k % NS;
            const uint32_t par = (uint32_t)(k / NS) & 1u;
            const long long w0 = p.dbg ? clock64() : 0;
            if (!tc::mbar_wait(&bar_full[s], par, abortp)) ok = false;
            const long long w1 = p.dbg ? clock64() : 0;
            if (GATHER && ok && !tc::mbar_wait(&bar_afull[s], par, abortp)) ok = false;
            if (p.dbg) { waited_b += w1 - w0; waited_a += clock64() - w1; }
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
        if (p.dbg && lane == 0) { long long* d = p.dbg + ((size_t)nb * gridDim.x + tile) * 16; d[5] = waited_b; d[6] = waited_a; }
    } else {
        // =============================================================================== WORKERS
        const int wi = warp - 2;
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;                // tile row == TMEM lane
        const bool row_ok = row < rows;
        const int grow = row0 + (row_ok ? row : 0);
        constexpr int NCG = NWORK / 4;                // epilogue column groups
        const int cgp = wi >> 2;
        const int nchunks = NC >> 3;
        const int colb = nb * NC;                     // first (padded) output column of this CTA
        bool ok = true;
        const bool stamp = p.dbg && wi == 0 && lane == 0;
        long long t0 = 0, t_gather = 0, t_acc = 0, g_load = 0, g_wait = 0, g_tail = 0, t_setup = 0;
        if (stamp) t0 = clock64();
        const int NKC = DP >> 3;
        if (GATHER) {
            constexpr int NG = NWORK / 4;             // gather groups (4 warps = 128 rows each)
            const int grp = wi >> 2;
            const int gi = (wi & 3) * 32 + lane;      // index within the group: which entry of the per-type row order this thread serves
            // ---- the tile's CSR slice -> shared memory
            const int base = p.row_ptr[(size_t)row0 * T];
            const int nptr = rows * T + 1;
            const int wt = tid - 64;
            for (int i = wt; i < TILE_M * T + 1; i += NWORK * 32) sPtr[i] = p.row_ptr[(size_t)row0 * T + min(i, nptr - 1)] - base;
            const int mt = p.row_ptr[(size_t)(row0 + rows) * T] - base;
            const bool cached = mt <= p.csr_cap;
            if (cached) for (int i = wt; i < mt; i += NWORK * 32) sSrc[i] = p.csr_src[base + i];
            asm volatile("bar.sync 1, %0;" ::"n"(NWORK * 32) : "memory");
            // ---- per present type: order the rows by message count class so that whole warps take the same path.  Only ~1/4 of the
            // (row, type) pairs of a molecule batch have a message at all, and most of those exactly one.
            for (int ti = wi; ti < s_ntypes; ti += NWORK) {
                const int t = s_types[ti];
                int cls[4], n0 = 0, n1 = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = i * 32 + lane;
                    const int cnt = r < rows ? sPtr[r * T + t + 1] - sPtr[r * T + t] : 0;
                    cls[i] = cnt >= 2 ? 0 : (cnt == 1 ? 1 : 2);
                    n0 += __popc(__ballot_sync(0xffffffffu, cls[i] == 0));
                    n1 += __popc(__ballot_sync(0xffffffffu, cls[i] == 1));
                }
                int run[3] = {0, n0, n0 + n1};
                const unsigned lt = (1u << lane) - 1u;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const unsigned m = __ballot_sync(0xffffffffu, cls[i] == c);
                        if (cls[i] == c) sPerm[ti * TILE_M + run[c] + __popc(m & lt)] = (uint8_t)(i * 32 + lane);
                        run[c] += __popc(m);
                    }
                }
                if (lane == 0) { s_n0[ti] = n0; s_n1[ti] = n1; }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NWORK * 32) : "memory");
            const int* srcs = cached ? sSrc : p.csr_src + base;
            if (stamp) t_setup = clock64();
            for (int k = grp; k < nk && ok; k += NG) {
                const long long c0 = stamp ? clock64() : 0;
                const int s = k % NS, it = k / NS;
                const int ti = k / NKS, ks = k - ti * NKS;
                const int t = s_types[ti];
                const int prow = sPerm[ti * TILE_M + gi];
                const int n0 = s_n0[ti], n01 = n0 + s_n1[ti];
                uint4 o0 = make_uint4(0u, 0u, 0u, 0u), o1 = o0, o2 = o0, o3 = o0;   // hi k-group 0 | hi k-group 1 | lo k-group 0 | lo k-group 1
                if (gi < n01) {
                    const int beg = sPtr[prow * T + t];
                    {
                        const int src = srcs[beg];
                        const uint8_t* sp = p.g_img + ((size_t)(src >> 7) * NKS + ks) * A_STAGE_B + (size_t)(src & 127) * 16;
                        o0 = __ldcg(reinterpret_cast<const uint4*>(sp));
                        o1 = __ldcg(reinterpret_cast<const uint4*>(sp + 2048));
                        o2 = __ldcg(reinterpret_cast<const uint4*>(sp + 4096));
                        o3 = __ldcg(reinterpret_cast<const uint4*>(sp + 6144));
                    }
                    if (gi < n0) {   // two or more messages: fp32 sum in message order, then re-split (two source rows in flight)
                        const int end = sPtr[prow * T + t + 1];
                        auto img_row = [&](int m) {
                            const int src = srcs[m];
                            return p.g_img + ((size_t)(src >> 7) * NKS + ks) * A_STAGE_B + (size_t)(src & 127) * 16;
                        };
                        const uint8_t* sp1 = img_row(beg + 1);
                        uint4 q0 = __ldcg(reinterpret_cast<const uint4*>(sp1)), q1 = __ldcg(reinterpret_cast<const uint4*>(sp1 + 2048));
                        uint4 q2 = __ldcg(reinterpret_cast<const uint4*>(sp1 + 4096)), q3 = __ldcg(reinterpret_cast<const uint4*>(sp1 + 6144));
                        float a0[8], a1[8];
                        unpack8(o0, a0); unpack8(o1, a1);
                        tc::unpack8_add(o2, a0, 1.0f); tc::unpack8_add(o3, a1, 1.0f);
                        for (int m = beg + 2; ; ++m) {
                            const uint4 c0 = q0, c1 = q1, c2 = q2, c3 = q3;
                            if (m < end) {
                                const uint8_t* sp = img_row(m);
                                q0 = __ldcg(reinterpret_cast<const uint4*>(sp)); q1 = __ldcg(reinterpret_cast<const uint4*>(sp + 2048));
                                q2 = __ldcg(reinterpret_cast<const uint4*>(sp + 4096)); q3 = __ldcg(reinterpret_cast<const uint4*>(sp + 6144));
                            }
                            tc::unpack8_add(c0, a0, 1.0f); tc::unpack8_add(c2, a0, 1.0f);
                            tc::unpack8_add(c1, a1, 1.0f); tc::unpack8_add(c3, a1, 1.0f);
                            if (m >= end) break;
                        }
                        tc::split8(a0, o0, o2);
                        tc::split8(a1, o1, o3);
                    }
                }
                const long long c1 = stamp ? clock64() : 0;
                if (it > 0 && !tc::mbar_wait(&bar_empty[s], (uint32_t)(it - 1) & 1u, abortp)) { ok = false; }
                ok = __all_sync(0xffffffffu, ok);
                if (!ok) break;
                const long long c2 = stamp ? clock64() : 0;
                uint8_t* ap = smem + (size_t)s * STAGE_B + (size_t)prow * 16;
                *reinterpret_cast<uint4*>(ap) = o0;
                *reinterpret_cast<uint4*>(ap + 2048) = o1;
                *reinterpret_cast<uint4*>(ap + 4096) = o2;
                *reinterpret_cast<uint4*>(ap + 6144) = o3;
                tc::fence_async_smem();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bar_afull[s]);
                if (stamp) { const long long c3 = clock64(); g_load += c1 - c0; g_wait += c2 - c1; g_tail += c3 - c2; }
            }
            if (stamp) t_gather = clock64();
        }
        // ---- epilogue: accumulator -> registers -> outputs.  The global operands of the first chunk are requested BEFORE the wait for the
        // accumulator, those of chunk c+1 before the math of chunk c (the loads are L2 hits after the prefetch above).
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const bool have_acc = nk > 0;
        const bool gru = p.cell == CELL_GRU;
        // the global operands of this thread's next chunk are requested before the math of the current one
        float hA[8], uA[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { hA[j] = 0.0f; uA[j] = 0.0f; }
        const bool cand_h = p.epi == EPI_CAND && (gru || p.sv_h);
        auto load_ops = [&](int c, float (&hb)[8], float (&ub)[8]) {   // operands of chunk c (if it exists)
            const int colp = colb + c * 8;
            if (c >= nchunks) return;
            if (p.epi == EPI_CAND) {
                if (colp >= DP) return;
                if (cand_h) chunk_load(p.h_chk, NKC, tile, colp >> 3, row, hb);
                if (gru) chunk_load(p.u_buf, NKC, tile, colp >> 3, row, ub);
            } else if (p.epi == EPI_GATE) {
                if (colp < DP) chunk_load(p.h_chk, NKC, tile, colp >> 3, row, hb);
            }
        };
        if (!GATHER) load_ops(cgp, hA, uA);
        if (ok && nk > 0) {
            if (!tc::mbar_wait(&bar_acc, 0, abortp)) ok = false;
            ok = __all_sync(0xffffffffu, ok);
            tc::tc_fence_after();
        }
        if (stamp) t_acc = clock64();
        if (ok) {
            if (p.epi == EPI_AGG) {
                const float den = (p.use_avg && row_ok) ? p.denom[grow] : 1.0f;
                for (int c = cgp; c < nchunks; c += NCG) {
                    const int col = colb + c * 8;
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
                auto gate_chunk = [&](int c, float (&hb)[8], float (&ub)[8]) -> bool {
                    const int colp = colb + c * 8;
                    if (c >= nchunks || colp >= 2 * DP) return false;
                    const bool is_r = colp < DP;
                    const int col = is_r ? colp : colp - DP;
                    float g[8], b[8], h[8];
                    tc::tmem_ld8_nowait(tmem + lane_addr + c * 8, g);
#pragma unroll
                    for (int j = 0; j < 8; ++j) h[j] = hb[j];
                    load_ops(c + NCG, hb, ub);
                    tc::load8_guarded(p.bias + (is_r ? 0 : D), col, D, b);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) g[j] = tc::sigmoid_fast(g[j] + b[j]);
                    if (is_r) {
                        float rh[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) rh[j] = g[j] * h[j];
                        if (p.sv_r && row_ok) {
                            tc::store8_guarded(p.sv_r + (size_t)grow * D, col, D, g);
                            tc::store8_guarded(p.sv_h + (size_t)grow * D, col, D, h);
                        }
                        img_store_chunk(p.img_out, NKS, tile, row, col, rh);
                    } else {
                        chunk_store(p.u_buf, NKC, tile, col >> 3, row, g);
                        if (p.sv_u && row_ok) tc::store8_guarded(p.sv_u + (size_t)grow * D, col, D, g);
                    }
                    return true;
                };
                for (int c = cgp; c < nchunks; c += NCG)
                    if (!gate_chunk(c, hA, uA)) break;
            } else {
                auto cand_chunk = [&](int c, float (&hb)[8], float (&ub)[8]) -> bool {
                    const int col = colb + c * 8;
                    if (c >= nchunks || col >= DP) return false;
                    float cv[8], b[8], h[8], u[8], hn[8];
                    tc::tmem_ld8_nowait(tmem + lane_addr + c * 8, cv);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { h[j] = hb[j]; u[j] = ub[j]; }
                    load_ops(c + NCG, hb, ub);
                    tc::load8_guarded(p.bias, col, D, b);
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
                    chunk_store(p.h_chk_out, NKC, tile, col >> 3, row, hn);
                    if (p.h_out && row_ok) tc::store8_guarded(p.h_out + (size_t)grow * D, col, D, hn);
                    img_store_chunk(p.img_out, NKS, tile, row, col, hn);
                    return true;
                };
                for (int c = cgp; c < nchunks; c += NCG)
                    if (!cand_chunk(c, hA, uA)) break;
            }
        }
        if (!ok && lane == 0) atomicExch(p.error_flag, 11);
        if (stamp) {
            long long* d = p.dbg + ((size_t)nb * gridDim.x + tile) * 16;
            d[0] = t0; d[1] = t_gather - t0; d[2] = t_acc - t0; d[3] = clock64() - t0; d[7] = nk;
            d[8] = g_load; d[9] = g_wait; d[10] = g_tail; d[11] = t_setup - t0;
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols) : "memory");
    }
}

// fp32 [V][D] row-major -> tile-major bf16 hi/lo image + chunk-major fp32 copy ([ntiles*128][DP], zero padded)
__global__ void ggnn_image_kernel(const float* __restrict__ x, uint8_t* __restrict__ img, float* __restrict__ chk, int V, int D, int DP, int ntiles) {
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
        chunk_store(chk, NKC, tile, kc, row, v);
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
