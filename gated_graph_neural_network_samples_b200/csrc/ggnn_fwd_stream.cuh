// Streaming GGNN propagation on 5th-gen tensor cores (tcgen05 / TMEM), sm_100a: the path for hidden sizes whose
// per-tile operands do not fit one SM (D > 128, BASELINE config 4) and for batches / graphs too large for the
// tile-local fused kernel (ggnn_fwd_tc.cuh).
//
// One timestep (sparse:153-216) = three launches of ONE kernel template, each a 128-row x NC-column output tile
// per CTA whose K dimension is streamed through a shared-memory ring, one UMMA K-step (16 columns) per stage:
//   EPI_AGG   agg       = [A_0 | .. | A_{T-1}] . [W_0; ..; W_{T-1}]      A_t[v] = sum of h[src] over the type-t messages into v,
//                         + indeg.B, / (deg + 1e-7)                       gathered by the worker warps straight into the ring (GATHER):
//                         a (target, type) pair with ONE message is a 64-byte asynchronous copy (cp.async, completion on the stage's
//                         mbarrier) of the source row's image chunks, a pair with none is zeros, and the few pairs with several
//                         messages are summed once per launch into "virtual rows" (prologue) and then copied like the others --
//                         no load latency sits between two K-steps of a gather warp
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
    int ksteps;            // K-steps (16 columns) per ring stage (4 unless the hidden size has fewer)
    int nparts;            // 3: bf16x3, 1: single bf16 MMA
    int npass;             // N blocks this CTA computes one after the other (TMA-fed kernels: accumulator of pass q in TMEM columns [q*NC, (q+1)*NC),
                           // so the epilogue of pass q runs under the mainloop of pass q+1); grid.y * npass = number of N blocks
    int tmem_cols;         // power of two >= max(32, npass*NC)
    int epi;               // EPI_*
    int cell, act, use_bias, use_avg;
    // ---- A operand, TMA-fed: nseg K segments, each a DP-wide tile-major image
    int nseg;
    const uint8_t* seg[MAX_SEG];
    // ---- A operand, gathered (EPI_AGG): per present edge type a DP-wide segment of per-type source sums
    const uint8_t* g_img;        // image of the state the messages are gathered from (a row with one type-t message is a 64-byte copy)
    const unsigned* tile_mask;   // [ntiles] bit t: some row of the tile receives a type-t message
    const int* pair_src;         // [ntiles*128*T] per (target, type): -1 no message | source node (exactly one message) | -(2 + vid) several
    const int* vrow_ptr;         // [NV+1] messages of the pairs with several messages (their "virtual rows"), CSR over vid ...
    const int* vsrc;             // ... source nodes in message order
    const int4* vinfo;           // [NV][2]: {count, src0, src1, src2 | src3 .. src6}: the first sources inline, one 32-byte load per virtual row
    const int* tile_vptr;        // [ntiles+1] vid range of each tile
    uint8_t* virt_img;           // image rows of the virtual rows (written in the prologue of every launch, row index = vid)
    int virt_rows;               // 1: pairs with several messages are pre-summed into virtual rows (molecule batches); 0: summed in the gather loop
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
    long long* dbg2;  // optional per-K-step timeline of CTA (0,0): [256][8] clocks (issuer: B landed, A landed, issued | gather: start, stage free, issued | producer: stage free, issued)
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

// Sum of the image rows of one (target, type) pair with several messages, one K-step (16 columns), fp32 in message order, re-split:
// (hi k-group 0, hi k-group 1, lo k-group 0, lo k-group 1).  `info` = vinfo[2*vid..]: {count, src0..src6}; longer lists continue in vsrc.
__device__ __forceinline__ void sum_pair_sources(const uint8_t* __restrict__ g_img, int NKS, int ks, const int4& i0, const int4& i1,
                                                 const int* __restrict__ vsrc_tail, uint4& h0, uint4& h1, uint4& l0, uint4& l1) {
    auto img_row = [&](int src) { return g_img + ((size_t)(src >> 7) * NKS + ks) * A_STAGE_B + (size_t)(src & 127) * 16; };
    float a0[8], a1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { a0[j] = 0.0f; a1[j] = 0.0f; }
    auto add_row = [&](const uint8_t* sp) {
        const uint4 y0 = __ldcg(reinterpret_cast<const uint4*>(sp)), y1 = __ldcg(reinterpret_cast<const uint4*>(sp + 2048));
        const uint4 y2 = __ldcg(reinterpret_cast<const uint4*>(sp + 4096)), y3 = __ldcg(reinterpret_cast<const uint4*>(sp + 6144));
        tc::unpack8_add(y0, a0, 1.0f); tc::unpack8_add(y2, a0, 1.0f);
        tc::unpack8_add(y1, a1, 1.0f); tc::unpack8_add(y3, a1, 1.0f);
    };
    const int cnt = i0.x;
    {   // the first two sources are always there: request both before the first add
        const uint8_t* s0 = img_row(i0.y);
        const uint8_t* s1 = img_row(i0.z);
        const uint4 x0 = __ldcg(reinterpret_cast<const uint4*>(s0)), x1 = __ldcg(reinterpret_cast<const uint4*>(s0 + 2048));
        const uint4 x2 = __ldcg(reinterpret_cast<const uint4*>(s0 + 4096)), x3 = __ldcg(reinterpret_cast<const uint4*>(s0 + 6144));
        const uint4 y0 = __ldcg(reinterpret_cast<const uint4*>(s1)), y1 = __ldcg(reinterpret_cast<const uint4*>(s1 + 2048));
        const uint4 y2 = __ldcg(reinterpret_cast<const uint4*>(s1 + 4096)), y3 = __ldcg(reinterpret_cast<const uint4*>(s1 + 6144));
        tc::unpack8_add(x0, a0, 1.0f); tc::unpack8_add(x2, a0, 1.0f);
        tc::unpack8_add(x1, a1, 1.0f); tc::unpack8_add(x3, a1, 1.0f);
        tc::unpack8_add(y0, a0, 1.0f); tc::unpack8_add(y2, a0, 1.0f);
        tc::unpack8_add(y1, a1, 1.0f); tc::unpack8_add(y3, a1, 1.0f);
    }
    if (cnt > 2) add_row(img_row(i0.w));
    if (cnt > 3) add_row(img_row(i1.x));
    if (cnt > 4) add_row(img_row(i1.y));
    if (cnt > 5) add_row(img_row(i1.z));
    if (cnt > 6) add_row(img_row(i1.w));
    for (int m = 7; m < cnt; ++m) add_row(img_row(vsrc_tail[m]));   // rare tail, in message order
    tc::split8(a0, h0, l0);
    tc::split8(a1, h1, l1);
}

template <int NWORK, bool GATHER>
__global__ void __launch_bounds__((NWORK + 2) * 32, 1) ggnn_stream_kernel(const __grid_constant__ StreamParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[MAX_NS];    // B (and TMA-fed A) bytes landed
    __shared__ __align__(8) uint64_t bar_afull[MAX_NS];   // gathered A written (GATHER)
    __shared__ __align__(8) uint64_t bar_empty[MAX_NS];   // the MMAs that read the stage are complete
    __shared__ __align__(8) uint64_t bar_acc[2];          // all MMAs of pass q are complete
    __shared__ __align__(8) uint64_t bar_virt[16];        // GATHER: the virtual rows of K group j are written (by the gather groups that have no ring slot)
    __shared__ uint32_t s_tmem;
    __shared__ int s_abort;
    __shared__ int s_types[32];
    __shared__ int s_ntypes;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int npass = GATHER ? 1 : p.npass;
    const int tile = blockIdx.x, nb0 = blockIdx.y * npass;   // first N block of this CTA
    const int D = p.D, DP = p.DP, T = p.T, NC = p.NC, NS = p.nstages, KS = p.ksteps;
    const int NKS = DP >> 4;
    const int GPS = (NKS + KS - 1) / KS;              // stages ("K groups") per K segment; the last one of a segment may be partial
    const int row0 = tile * TILE_M;
    const int rows = min(TILE_M, p.V - row0);
    const uint32_t B_STEP_B = 64u * (uint32_t)NC;    // one K-step of the B operand: [hi: 2 k-groups x NC x 16 B | lo: same]
    const uint32_t A_REGION_B = (uint32_t)KS * A_STAGE_B;
    const uint32_t STAGE_B = A_REGION_B + (uint32_t)KS * B_STEP_B;   // a stage = KS K-steps of A, then KS K-steps of B
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    int* sPair = reinterpret_cast<int*>(smem + (size_t)NS * STAGE_B);   // [128*T] the tile's slice of pair_src (GATHER)

    if (tid == 0) {
        s_abort = 0;
        for (int i = 0; i < MAX_NS; ++i) { tc::mbar_init(&bar_full[i], 1); tc::mbar_init(&bar_afull[i], TILE_M); tc::mbar_init(&bar_empty[i], 1); }
        tc::mbar_init(&bar_acc[0], 1); tc::mbar_init(&bar_acc[1], 1);
        {   // gather groups beyond the ring depth have no stage to fill: they pre-sum the virtual rows of K groups 1.. while the others gather
            const int idle_warps = (NWORK / 4 - min(NWORK / 4, NS)) * 4;
            for (int i = 0; i < 16; ++i) tc::mbar_init(&bar_virt[i], idle_warps > 0 ? idle_warps : 1);
        }
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
    const int nsegs = GATHER ? s_ntypes : p.nseg;    // K segments of this tile (present edge types / operand images)
    const int ng = nsegs * GPS;                        // stages to stream
    const int nk = nsegs * NKS;                        // K-steps in total

    if (warp == 0) {
        // =============================================================================== PRODUCER (one thread)
        // Issuing a bulk copy costs this thread several hundred cycles whatever its size, so a stage carries KS K-steps: the A operand of
        // those K-steps is ONE contiguous piece of the image, their B operand ONE contiguous piece of the pre-tiled weights.
        if (lane == 0) {
            bool ok = true;
            long long waited = 0;
            int s = 0, round = 0;
            for (int pass = 0; pass < npass && ok; ++pass) {
            const uint8_t* wb = p.w + (size_t)(nb0 + pass) * p.kt_all * B_STEP_B;
            int sg = 0, j = 0;
            for (int g = 0; g < ng && ok; ++g) {
                if (round > 0) {
                    const long long w0 = p.dbg ? clock64() : 0;
                    if (!tc::mbar_wait(&bar_empty[s], (uint32_t)(round - 1) & 1u, abortp)) { ok = false; break; }
                    if (p.dbg) waited += clock64() - w0;
                }
                long long* d2 = (p.dbg2 && tile == 0 && nb0 == 0 && g < 256) ? p.dbg2 + g * 8 : nullptr;
                if (d2) d2[6] = clock64();
                const int ks0 = j * KS, nks = min(KS, NKS - ks0);
                uint8_t* st = smem + (size_t)s * STAGE_B;
                const int seg_k0 = (GATHER ? s_types[sg] : sg) * NKS + ks0;   // first K-step of the stage in the weight stream
                if (GATHER) {
                    tc::mbar_arrive_expect_tx(&bar_full[s], (uint32_t)nks * B_STEP_B);
                } else {
                    tc::mbar_arrive_expect_tx(&bar_full[s], (uint32_t)nks * ((uint32_t)A_STAGE_B + B_STEP_B));
                    tc::bulk_copy_g2s(st, p.seg[sg] + ((size_t)tile * NKS + ks0) * A_STAGE_B, (uint32_t)nks * A_STAGE_B, &bar_full[s]);
                }
                tc::bulk_copy_g2s(st + A_REGION_B, wb + (size_t)seg_k0 * B_STEP_B, (uint32_t)nks * B_STEP_B, &bar_full[s]);
                if (d2) d2[7] = clock64();
                // K order: TMA-fed kernels walk segment by segment; the gather GEMM walks K group by K group over all edge types (the virtual
                // rows of K group 0 are then enough to start, the rest are pre-summed while the ring already turns)
                if (GATHER) { if (++sg == nsegs) { sg = 0; ++j; } }
                else if (++j == GPS) { j = 0; ++sg; }
                if (++s == NS) { s = 0; ++round; }
            }
            }
            if (!ok) atomicExch(p.error_flag, 13);
            if (p.dbg) p.dbg[((size_t)blockIdx.y * gridDim.x + tile) * 16 + 4] = waited;
        }
    } else if (warp == 1) {
        // =============================================================================== MMA ISSUER
        bool ok = true;
        const bool x3 = p.nparts == 3;
        const uint64_t descA = tc::make_desc(0, 2048, 128);
        const uint64_t descB = tc::make_desc(0, 16u * (uint32_t)NC, 128);
        const uint32_t idesc = tc::make_idesc_bf16(NC);
        const uint32_t b_lo16 = (32u * (uint32_t)NC) >> 4;
        const uint32_t smem16 = smem_u32(smem) >> 4, stage16 = STAGE_B >> 4, aregion16 = A_REGION_B >> 4, bstep16 = B_STEP_B >> 4;
        const uint32_t tm_d = __shfl_sync(0xffffffffu, tmem, 0);
        long long waited_b = 0, waited_a = 0;
        int s = 0;
        uint32_t par = 0;
        for (int pass = 0; pass < npass && ok; ++pass) {
        const uint32_t tm_p = tm_d + (uint32_t)(pass * NC);
        int j = 0, sgi = 0;
        for (int g = 0; g < ng && ok; ++g) {
            const long long w0 = p.dbg ? clock64() : 0;
            if (!tc::mbar_wait(&bar_full[s], par, abortp)) ok = false;
            const long long w1 = p.dbg ? clock64() : 0;
            if (GATHER && ok && !tc::mbar_wait(&bar_afull[s], par, abortp)) ok = false;
            if (p.dbg) { waited_b += w1 - w0; waited_a += clock64() - w1; }
            long long* d2 = (p.dbg2 && tile == 0 && nb0 == 0 && g < 256 && lane == 0) ? p.dbg2 + g * 8 : nullptr;
            if (d2) { d2[0] = w1; d2[1] = clock64(); }
            ok = __all_sync(0xffffffffu, ok);
            if (!ok) break;
            if (GATHER) tc::fence_async_smem();   // the gathered A stage was written through the generic proxy (cp.async / st.shared): order it before the MMA's reads
            tc::tc_fence_after();
            const int nks = min(KS, NKS - j * KS);
            if (tc::elect_one()) {
                const uint32_t a16 = smem16 + (uint32_t)s * stage16, b16 = a16 + aregion16;
                for (int i = 0; i < nks; ++i) {
                    const uint64_t ah = descA | (uint64_t)(a16 + (uint32_t)i * (A_STAGE_B >> 4)), al = ah + (4096u >> 4);
                    const uint64_t bh = descB | (uint64_t)(b16 + (uint32_t)i * bstep16), bl = bh + b_lo16;
                    tc::umma_bf16(tm_p, ah, bh, idesc, (g > 0 || i > 0) ? 1u : 0u);
                    if (x3) {
                        tc::umma_bf16(tm_p, ah, bl, idesc, 1u);
                        tc::umma_bf16(tm_p, al, bh, idesc, 1u);
                    }
                }
                tc::umma_commit(&bar_empty[s]);
                if (g == ng - 1) tc::umma_commit(&bar_acc[pass]);
            }
            __syncwarp();
            if (d2) d2[2] = clock64();
            if (GATHER) { if (++sgi == nsegs) { sgi = 0; ++j; } }
            else if (++j == GPS) j = 0;
            if (++s == NS) { s = 0; par ^= 1u; }
        }
        }
        if (!ok && lane == 0) atomicExch(p.error_flag, 12);
        if (p.dbg && lane == 0) { long long* d = p.dbg + ((size_t)blockIdx.y * gridDim.x + tile) * 16; d[5] = waited_b; d[6] = waited_a; }
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
        bool ok = true;
        const bool stamp = p.dbg && wi == 0 && lane == 0;
        long long t0 = 0, t_gather = 0, t_acc = 0, g_load = 0, g_wait = 0, g_tail = 0, t_setup = 0;
        if (stamp) t0 = clock64();
        const int NKC = DP >> 3;
        if (GATHER) {
            constexpr int NG = NWORK / 4;             // gather groups (4 warps = 128 rows each)
            const int grp = wi >> 2;
            const int gi = (wi & 3) * 32 + lane;      // the tile row this thread gathers
            const int wt = tid - 64;
            // ---- the tile's (target, type) -> source table
            for (int i = wt; i < TILE_M * T; i += NWORK * 32) sPair[i] = p.pair_src[(size_t)row0 * T + i];
            const long long t_pair = stamp ? clock64() : 0;
            // ---- virtual rows: the pairs of this tile with several messages, summed in message order (fp32), re-split, stored as image rows.
            // The K loop of the gather GEMM is K-group major, so only the virtual rows of K group 0 must exist before the first stage:
            // all workers pre-sum those; the groups beyond the ring depth (no stage to fill, see below) then pre-sum K groups 1.. and signal
            // each one on bar_virt[j] while the other groups already gather.  Without surplus groups everything is pre-summed up front.
            // (p.virt_rows == 0: pairs with several messages are summed inside the gather loop instead.)
            const int NGE = min(NG, NS);
            const int v0 = p.tile_vptr[tile], nv = p.virt_rows ? p.tile_vptr[tile + 1] - v0 : 0;
            auto presum = [&](int jg, int first, int nthreads_) {   // virtual rows of K group jg, tasks dealt to `nthreads_` threads
                const int ks0 = jg * KS, nks = min(KS, NKS - ks0);
                for (int task = first; task < nv * nks; task += nthreads_) {
                    const int vl = task / nks, vid = v0 + vl, ks = ks0 + (task - vl * nks);
                    const int4 i0 = __ldg(p.vinfo + 2 * (size_t)vid), i1 = __ldg(p.vinfo + 2 * (size_t)vid + 1);
                    uint4 h0, l0, h1, l1;
                    sum_pair_sources(p.g_img, NKS, ks, i0, i1, p.vsrc + p.vrow_ptr[i0.x > 7 ? vid : 0], h0, h1, l0, l1);
                    uint8_t* vp = p.virt_img + ((size_t)(vid >> 7) * NKS + ks) * A_STAGE_B + (size_t)(vid & 127) * 16;
                    *reinterpret_cast<uint4*>(vp) = h0;
                    *reinterpret_cast<uint4*>(vp + 2048) = h1;
                    *reinterpret_cast<uint4*>(vp + 4096) = l0;
                    *reinterpret_cast<uint4*>(vp + 6144) = l1;
                }
            };
            // worth it only if the surplus groups finish a K group's virtual rows (~7 k cycles per round of tasks) within the ~450 cycles per
            // K-step the ring needs to get there: molecule batches yes (cfg4: 42 virtual rows per tile), dense graphs no (cfg5: ~300)
            const int nidle_thr = (NG - NGE) * 128;
            const bool overlap = NGE < NG && ((nv * KS + nidle_thr - 1) / max(nidle_thr, 1)) * 16 <= nsegs * KS;
            if (p.virt_rows) {
                for (int jg = 0; jg < (overlap ? 1 : GPS); ++jg) presum(jg, wt, NWORK * 32);
                const long long t_virt = stamp ? clock64() : 0;
                __threadfence();   // the copies below read these rows back through L2 (cp.async.cg)
                if (stamp) { long long* d = p.dbg + ((size_t)blockIdx.y * gridDim.x + tile) * 16; d[12] = t_pair - t0; d[13] = t_virt - t0; d[14] = clock64() - t0; d[15] = nv; }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NWORK * 32) : "memory");
            if (stamp) t_setup = clock64();
            if (overlap && grp >= NGE && p.virt_rows) {
                const int nidle = (NG - NGE) * 128, me = (grp - NGE) * 128 + gi;
                for (int jg = 1; jg < GPS; ++jg) {
                    presum(jg, me, nidle);
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&bar_virt[jg]);
                }
            }
            // group `grp` fills stages grp, grp + NGE, ...: every row is an asynchronous 64-byte copy per K-step (or zeros).
            // No more groups than ring stages: a parity wait on a stage's barrier is only sound if the waiter cannot be two phases ahead
            // of it, and group X waiting for round r of a stage has (through its previous stage, NGE K-groups back) only seen round r-2
            // complete when NS >= NGE.  Surplus groups pre-sum virtual rows instead (above).
            int j_seen = 0;
            for (int g = grp < NGE ? grp : ng; g < ng && ok; g += NGE) {
                const int j = g / nsegs, sg = g - j * nsegs;   // K group major: stage g = (K group j, present edge type sg)
                const long long c0 = stamp ? clock64() : 0;
                const int s = g % NS, round = g / NS;
                long long* d2 = (p.dbg2 && tile == 0 && nb0 == 0 && g < 256 && gi == 0) ? p.dbg2 + g * 8 : nullptr;
                if (d2) d2[3] = clock64();
                const int ks0 = j * KS, nks = min(KS, NKS - ks0);
                const int ps = sPair[gi * T + s_types[sg]];
                if (ps < -1 && overlap && p.virt_rows && j > j_seen) {   // this K group's virtual rows come from the surplus groups
                    if (!tc::mbar_wait(&bar_virt[j], 0, abortp)) *abortp = 1;
                    j_seen = j;
                }
                const uint8_t* sp = nullptr;
                if (ps >= 0) sp = p.g_img + ((size_t)(ps >> 7) * NKS + ks0) * A_STAGE_B + (size_t)(ps & 127) * 16;
                else if (ps < -1 && p.virt_rows) { const int vid = -(ps + 2); sp = p.virt_img + ((size_t)(vid >> 7) * NKS + ks0) * A_STAGE_B + (size_t)(vid & 127) * 16; }
                const long long c1 = stamp ? clock64() : 0;
                // one warp of the group polls the stage's barrier, the other three block on a hardware barrier (polling warps cost issue slots)
                if (round > 0 && (wi & 3) == 0 && !tc::mbar_wait(&bar_empty[s], (uint32_t)(round - 1) & 1u, abortp)) *abortp = 1;
                asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory");
                if (*abortp) { ok = false; break; }
                const long long c2 = stamp ? clock64() : 0;
                if (d2) d2[4] = clock64();
                uint8_t* ap = smem + (size_t)s * STAGE_B + (size_t)gi * 16;
                const uint32_t bar = smem_u32(&bar_afull[s]);
                if (sp) {
                    uint32_t dst = smem_u32(ap);
                    for (int i = 0; i < nks; ++i, dst += A_STAGE_B, sp += A_STAGE_B) {
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(sp) : "memory");
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst + 2048u), "l"(sp + 2048) : "memory");
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst + 4096u), "l"(sp + 4096) : "memory");
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst + 6144u), "l"(sp + 6144) : "memory");
                    }
                    // one arrival on the stage's barrier when this thread's copies have landed (the barrier counts 128 threads)
                    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(bar) : "memory");
                } else if (ps < -1) {   // several messages, summed here (dense graphs): loads -> registers -> stage
                    const int vid = -(ps + 2);
                    const int4 i0 = __ldg(p.vinfo + 2 * (size_t)vid), i1 = __ldg(p.vinfo + 2 * (size_t)vid + 1);
                    const int* tail = p.vsrc + p.vrow_ptr[i0.x > 7 ? vid : 0];
                    for (int i = 0; i < nks; ++i, ap += A_STAGE_B) {
                        uint4 h0, l0, h1, l1;
                        sum_pair_sources(p.g_img, NKS, ks0 + i, i0, i1, tail, h0, h1, l0, l1);
                        *reinterpret_cast<uint4*>(ap) = h0;
                        *reinterpret_cast<uint4*>(ap + 2048) = h1;
                        *reinterpret_cast<uint4*>(ap + 4096) = l0;
                        *reinterpret_cast<uint4*>(ap + 6144) = l1;
                    }
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
                } else {
                    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                    for (int i = 0; i < nks; ++i, ap += A_STAGE_B) {
                        *reinterpret_cast<uint4*>(ap) = z;
                        *reinterpret_cast<uint4*>(ap + 2048) = z;
                        *reinterpret_cast<uint4*>(ap + 4096) = z;
                        *reinterpret_cast<uint4*>(ap + 6144) = z;
                    }
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
                }
                if (d2) d2[5] = clock64();
                if (stamp) { const long long c3 = clock64(); g_load += c1 - c0; g_wait += c2 - c1; g_tail += c3 - c2; }
            }
            if (stamp) t_gather = clock64();
        }
        // ---- epilogue: accumulator -> registers -> outputs.  The global operands of the first chunk are requested BEFORE the wait for the
        // accumulator, those of chunk c+1 before the math of chunk c (the loads are L2 hits after the prefetch above).
        const bool have_acc = nk > 0;
        const bool gru = p.cell == CELL_GRU;
        for (int pass = 0; pass < npass; ++pass) {
        const int colb = (nb0 + pass) * NC;           // first (padded) output column of this pass
        const uint32_t lane_addr = ((uint32_t)(q * 32) << 16) + (uint32_t)(pass * NC);
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
        if (nk > 0) {   // worker warp 0 polls for the accumulator, the others block on the hardware barrier behind it
            if (wi == 0 && ok && !tc::mbar_wait(&bar_acc[pass], 0, abortp)) *abortp = 1;
            asm volatile("bar.sync 1, %0;" ::"n"(NWORK * 32) : "memory");
            if (*abortp) ok = false;
            tc::tc_fence_after();
        }
        if (stamp) t_acc = clock64();
        if (ok) {
            if (p.epi == EPI_AGG) {
                // sparse:207-209 divides by (sum of in-degrees + 1e-7); one IEEE reciprocal per row, then a multiply per element (differs from
                // the quotient by at most 1 ulp; tests/test_gpu_stream.py bounds the growth over 32 timesteps)
                const float inv_den = (p.use_avg && row_ok) ? __frcp_rn(p.denom[grow]) : 1.0f;
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
                    for (int j = 0; j < 8; ++j) v[j] = row_ok ? v[j] * inv_den : 0.0f;
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
        }   // passes
        if (!ok && lane == 0) atomicExch(p.error_flag, 11);
        if (stamp) {
            long long* d = p.dbg + ((size_t)blockIdx.y * gridDim.x + tile) * 16;
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
