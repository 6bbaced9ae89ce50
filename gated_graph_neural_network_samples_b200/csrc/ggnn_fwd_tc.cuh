// Fused GGNN propagation on 5th-gen tensor cores (tcgen05 / TMEM), sm_100a.
//
// One CTA owns a tile of up to 128 node rows (UMMA M = 128, cta_group::1).  hidden_size D is padded to
// DP = roundup(D, 16) <= 128 inside the kernel only.  Per timestep (sparse:153-216):
//   G1  agg   [128 x DP ]  = sum_t A_t[128 x DP] . W_t          (A_t = per-type sum of gathered source states)
//   G2  gates [128 x 2DP]  = [agg | h] . K_g                     (+ residual pre-product, + b_g, sigmoid)
//   G3  cand  [128 x DP ]  = [agg | r*h] . K_c                   (+ residual pre-product, + b_c, act)
//   h' = u*h + (1-u)*cand
// fp32 accuracy on bf16 tensor cores: every operand is split x = hi + lo (two bf16, 16 mantissa bits) and each
// product is issued as 3 MMAs  Ah.Bh + Ah.Bl + Al.Bh  (the dropped Al.Bl term is 2^-16 relative); "fast" mode
// issues Ah.Bh only.  Accumulators live in TMEM ([h fp32 | agg/cand acc | gate acc] = 4*DP <= 512 columns);
// the fp32 master copy of the node states also lives in TMEM, so the recurrence never leaves the SM in LOCAL mode.
//
// Shared memory: three A-operand tiles (h, agg, A_t / r*h), each hi+lo in the canonical K-major no-swizzle UMMA
// layout  byte(row, k) = part*(DP*256) + (k/8)*2048 + row*16 + (k%8)*2  (half of both strides when every tile has <= 64 rows:
// compact tiles, see KGS below),  plus a ring of weight stages that a
// producer thread fills with cp.async.bulk (1-D TMA) from a pre-split, pre-tiled bf16 copy of the weights.  The stream is
// latency bound (bytes in flight / slot round trip), so the ring is extended into A-operand tiles while they hold no live
// operand: the A_t / r*h tile during the gate GEMM, the h tile during the candidate GEMM.
// Warp roles: warps 0-15 = workers (gather, epilogues; TMEM lane quarter = warp%4, column-chunk group = warp/4),
// warps 16-17 = MMA issuers (each owns a fixed subset of the weight slots; every MMA accumulates into accumulators the workers zero after
// reading, so cross-warp issue order is irrelevant), warp 18 = weight producer (ONE thread, strictly in order: parity-based
// mbarrier waits are only sound when no agent can lap another by a whole ring) + TMEM allocator.
// Every mbarrier wait is bounded; on timeout an error code is written and all roles drain.
#pragma once
#include <cuda_bf16.h>

#include "ggnn_common.cuh"

namespace ggnn {
namespace tc {

// Profiling hooks (phase stamps + per-role event log of tile 0, read by tools/tc_trace.py / tc_phase_timing.py) are compiled in only
// with -DGGNN_TC_TRACE (GGNN_TC_TRACE=1 python -m ..._build): in the shipped kernel they were 10 % of the issued warp instructions.
#ifdef GGNN_TC_TRACE
#define GGNN_TRACE_ON(p) ((p).dbg != nullptr)
#else
#define GGNN_TRACE_ON(p) false
#endif

constexpr int TILE_M = 128;
constexpr int NUM_WORKERS = 512;          // 16 worker warps: 4 per TMEM lane quarter
#ifndef GGNN_TC_ISSUERS
#define GGNN_TC_ISSUERS 2
#endif
constexpr int NUM_ISSUERS = GGNN_TC_ISSUERS;            // MMA-issuing warps; issuer i owns the weight slots s with s % NUM_ISSUERS == i (2 and 3 verified, 4 deadlocks)
constexpr int WARP_MMA = 16;              // first issuer warp
#ifndef GGNN_TC_PRODUCERS
#define GGNN_TC_PRODUCERS 2
#endif
constexpr int NUM_PRODUCERS = GGNN_TC_PRODUCERS;   // weight-producer warps (one thread each); producer j owns the slots s with s % NUM_PRODUCERS == j
constexpr int WARP_PROD = WARP_MMA + NUM_ISSUERS;   // first producer warp (also the TMEM allocator)
constexpr int NTHREADS = (WARP_PROD + NUM_PRODUCERS) * 32;
constexpr int MAX_STAGES = 10;            // ring slots proper; a slot holds TWO K-step stages (2 x 64*DP bytes, one bulk copy)
constexpr int EXT_SLOTS = 4;              // an idle 128-row A-operand tile holds 4 more slots (DP*512 / DP*128); a compact 64-row tile holds 2
constexpr int MAX_SLOTS = MAX_STAGES + 2 * EXT_SLOTS;
enum { SET_BASE = 0, SET_XA = 1, SET_XH = 2 };   // ring only | ring + opA tile (gate phase) | ring + opH tile (candidate phase)

struct TcLayer {
    // pre-split (bf16 hi/lo), pre-tiled weights: one 64*DP-byte stage per K-step (16 rows) of a [K x DP] block
    const uint8_t* w_edge;    // [T*DP/16]       W_t, t-major
    const uint8_t* w_gate;    // [(R+2)*DP/16] x 2 slots: K_g as one N = 2*DP operand (cols r | u): hi slot then lo slot per K-step
                              //                 rows: residual segments..., agg, h
    const uint8_t* w_cand;    // [(R+2)*DP/16]   K_c (RNN: the only kernel)
    const float* edge_b;    // fp32 originals (unpadded)
    const float* gate_b;
    const float* cand_b;
    int steps, nres;
    int res[MAX_RES];
};

struct TcParams {
    int V, D, DP, T, L;
    int use_bias, use_avg, cell, act;
    int gather_mode, dense_v, save;
    int nparts;   // 3: bf16x3 (fp32-accurate), 1: single bf16 MMA
    int nstages;  // weight ring depth
    int kgs;            // A-operand k-group stride in bytes: 2048 (128-row tiles) or 1024 (compact: every tile has <= 64 rows)
    int ngbuf;          // gather buffers: 2 (A_t alternates between the opA and opX tiles) or 4 (two more compact tiles: with <= 4 present edge
                        // types every gather runs ahead of the MMAs of the previous type and the issuer never waits for an operand)
    int csr_cache;      // LOCAL sparse only: the tile's CSR slice is staged in shared memory (uint16 row offsets, uint8 local sources)
    int csr_cap_msgs;   // capacity of the shared source array
    const int* tile_start;
    const unsigned* tile_mask;
    const int* row_ptr;
    const int* csr_src;
    const float* dense_adj;
    const float* indeg;
    const float* denom;
    const float* state[MAX_LAYERS + 1];
    float* state_w[MAX_LAYERS + 1];
    TcLayer layer[MAX_LAYERS];
    SaveDev save_buf;
    int step_base[MAX_LAYERS];
    float* res_pre;  // [ntiles][128][3*DP] fp32 scratch: residual pre-products of the current layer
    int g_layer, g_step;
    const float* g_in;
    float* g_out;
    float drop_keep;                // state dropout (ggnn_common.cuh dropout_apply); off when >= 1
    unsigned long long drop_seed;
    int* error_flag;
    long long* dbg;  // optional [512] profiling aid written by tile 0, or nullptr: [0,64) phase stamps (see tools/tc_phase_timing.py),
                     // then (code, clock64) event pairs of the SECOND timestep: [64,192) worker thread 0, [192,320) issuer 0, [320,448) producer 0
};

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait (~1 s at 2 GHz).  Returns false on timeout or if another role already aborted.
__device__ __forceinline__ bool mbar_try(uint32_t addr, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __noinline__ bool mbar_wait_slow(uint32_t addr, uint32_t parity, volatile int* abort_flag) {
    const long long t0 = clock64();
    for (unsigned it = 1;; ++it) {
        uint32_t ok;
        // the hardware suspends the thread (up to the hint, in ns) instead of spinning; it wakes on barrier events
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(addr), "r"(parity), "r"(20000u) : "memory");
        if (ok) return true;
        if ((it & 63u) == 0u) {   // keep the common iteration tiny: pollers share issue slots with the MMA warp
            if (*abort_flag) return false;
            if (clock64() - t0 > 2000000000LL) { *abort_flag = 1; return false; }
        }
    }
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, volatile int* abort_flag) {
    const uint32_t addr = smem_u32(bar);
    if (mbar_try(addr, parity)) return true;
    return mbar_wait_slow(addr, parity, abort_flag);
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
}
// issue only; the destination registers are valid after tmem_ld_wait()
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, float (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                   "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// smem matrix descriptor: K-major, SWIZZLE_NONE, version 1 (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptor: kind::f16, A/B = BF16, D = F32, both K-major, M = 128
__device__ __forceinline__ uint32_t make_idesc_bf16(int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {   // true in exactly one (converged) lane
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// x = hi + lo with hi, lo bf16 (round to nearest): 16 mantissa bits kept
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 hp = __floats2bfloat162_rn(x[2 * j], x[2 * j + 1]);   // one packed cvt
        h[j] = *reinterpret_cast<const uint32_t*>(&hp);
        const float r0 = x[2 * j] - __uint_as_float(h[j] << 16);
        const float r1 = x[2 * j + 1] - __uint_as_float(h[j] & 0xFFFF0000u);
        const __nv_bfloat162 lp = __floats2bfloat162_rn(r0, r1);
        l[j] = *reinterpret_cast<const uint32_t*>(&lp);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void unpack8_add(const uint4& a, float (&acc)[8], float scale) {
    const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        acc[2 * j] = fmaf(scale, __uint_as_float(w[j] << 16), acc[2 * j]);
        acc[2 * j + 1] = fmaf(scale, __uint_as_float(w[j] & 0xFFFF0000u), acc[2 * j + 1]);
    }
}
__device__ __forceinline__ float sigmoid_fast(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }
__device__ __forceinline__ float act_fast(float v, int act) {
    return act == ACT_TANH ? (1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * v))) : fmaxf(v, 0.0f);
}

// store one [row, 8-column chunk] of an A operand (both parts)
// (rows at or beyond the allocated row count of a compact tile are not stored: they would alias the next k-group)
__device__ __forceinline__ void store_operand_chunk(uint8_t* op, uint32_t kgs, uint32_t part_b, int kc, int row, const float (&x)[8]) {
    if ((uint32_t)row * 16u >= kgs) return;
    uint4 hi, lo;
    split8(x, hi, lo);
    uint8_t* p = op + (size_t)kc * kgs + (size_t)row * 16;
    *reinterpret_cast<uint4*>(p) = hi;
    *reinterpret_cast<uint4*>(p + part_b) = lo;
}

// ------------------------------------------------------------------------------------------------ the kernel
// 8-wide helpers on [row, 8-column chunk] tiles.  D % 4 == 0, so every float4 of a chunk is entirely inside or
// entirely outside the D real columns.
__device__ __forceinline__ void load8_guarded(const float* base, int col0, int D, float (&v)[8]) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (col0 + 4 <= D) a = *reinterpret_cast<const float4*>(base + col0);
    if (col0 + 8 <= D) b = *reinterpret_cast<const float4*>(base + col0 + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8_guarded_cg(const float* base, int col0, int D, float (&v)[8]) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (col0 + 4 <= D) a = __ldcg(reinterpret_cast<const float4*>(base + col0));
    if (col0 + 8 <= D) b = __ldcg(reinterpret_cast<const float4*>(base + col0 + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8_guarded(float* base, int col0, int D, const float (&v)[8]) {
    if (col0 + 4 <= D) *reinterpret_cast<float4*>(base + col0) = make_float4(v[0], v[1], v[2], v[3]);
    if (col0 + 8 <= D) *reinterpret_cast<float4*>(base + col0 + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void lds8(const float* s, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(s), b = *reinterpret_cast<const float4*>(s + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

template <bool LOCAL>
__global__ void __launch_bounds__(NTHREADS, 1) ggnn_fwd_tc_kernel(const __grid_constant__ TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_w_full[MAX_SLOTS];
    __shared__ __align__(8) uint64_t bar_w_empty[MAX_SLOTS];
    __shared__ __align__(8) uint64_t bar_xa_free;   // opA tile holds no live operand: the producer may use it as ring slots (gate phase)
    __shared__ __align__(8) uint64_t bar_xh_free;   // opH tile dead until the state update rewrites it (candidate phase)
    __shared__ __align__(8) uint64_t bar_a_ready;
    __shared__ __align__(8) uint64_t bar_mma_done;
    __shared__ __align__(8) uint64_t bar_g1_done[4];   // one per gather buffer (opA, opX, opB, opC): its MMAs are complete
    __shared__ __align__(8) uint64_t bar_g_ready[4];   // one per gather buffer: A_t written (never more than one phase pending each)
    __shared__ uint32_t s_tmem_base;
    __shared__ int s_abort;

    const int D = p.D, DP = p.DP, T = p.T;
    const int NKC = DP >> 3;   // 8-column chunks per DP
    const int NKS = DP >> 4;   // UMMA K-steps (16) per DP-wide operand
    // A operand tile: byte(row, k) = part*PART_B + (k/8)*KGS + row*16 + (k%8)*2.  KGS = 16 * allocated rows.  A compact tile (64 rows) is
    // still multiplied with M = 128: the MMA's rows 64..127 of k-group g alias rows 0..63 of k-group g+1 -- in-bounds bytes whose products land
    // in TMEM lanes 64..127, which nobody reads.  Half the operand bytes leave room for a ~3x deeper weight ring.
    const uint32_t KGS = (uint32_t)p.kgs;
    const uint32_t PART_B = (uint32_t)DP * KGS / 8u;   // bytes per part (hi or lo)
    const uint32_t OPB = 2u * PART_B;                  // bytes per A operand (hi + lo)
    const uint32_t EXT = KGS / 512u;                   // ring slots an idle operand tile can hold
    const uint32_t STAGE_B = (uint32_t)DP * 64u;       // bytes per weight stage (K = 16 x N = DP, hi + lo)
    uint8_t* opH = smem;
    uint8_t* opX = opH + OPB;
    uint8_t* opA = opX + OPB;
    const int NGB = p.ngbuf;                               // gather buffers 2 and 3 (when present) sit between the opA tile and the ring
    uint8_t* ring = opA + (size_t)OPB * (size_t)(NGB - 1); // nstages slots of 2*STAGE_B, 1024-byte aligned (DP*512 and DP*128 are multiples of 1024)
    float* sBias = reinterpret_cast<float*>(ring + (size_t)p.nstages * 2 * STAGE_B);   // [3*DP]: gate r | gate u | cand, zero padded
    uint16_t* sRowPtr = reinterpret_cast<uint16_t*>(sBias + 3 * DP);              // [128*T + 1] (csr_cache)
    uint8_t* sSrc = reinterpret_cast<uint8_t*>(sRowPtr + ((TILE_M * T + 1 + 7) & ~7)); // [csr_cap_msgs]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x;
    const int row0 = p.tile_start[tile];
    const int rows = p.tile_start[tile + 1] - row0;
    const unsigned tmask = p.tile_mask[tile];
    const size_t VD = (size_t)p.V * D;
    const int nst = p.nstages;

    if (tid == 0) {
        s_abort = 0;
        for (int i = 0; i < MAX_SLOTS; ++i) { mbar_init(&bar_w_full[i], 1); mbar_init(&bar_w_empty[i], 1); }
        mbar_init(&bar_xa_free, NUM_ISSUERS);
        mbar_init(&bar_xh_free, NUM_ISSUERS);
        mbar_init(&bar_a_ready, NUM_WORKERS / 32);   // one arrival per worker warp
        mbar_init(&bar_mma_done, NUM_ISSUERS);      // one tcgen05.commit per issuer warp
        for (int i = 0; i < 4; ++i) { mbar_init(&bar_g1_done[i], NUM_ISSUERS); mbar_init(&bar_g_ready[i], NUM_WORKERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == WARP_PROD) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem_base;
    const uint32_t TM_H = tmem, TM_ACC = tmem + (uint32_t)DP, TM_GATE = tmem + 2u * (uint32_t)DP;
    volatile int* abortp = &s_abort;

    const int l_begin = LOCAL ? 0 : p.g_layer;
    const int l_end = LOCAL ? p.L : p.g_layer + 1;

    if (warp < NUM_WORKERS / 32) {
        // =============================================================================== WORKERS
        const int q = warp & 3, cg = warp >> 2;      // TMEM lane quarter, column-chunk group
        // compact tiles (<= 64 rows): the warps of lane quarters 2 and 3 own no rows -- they keep every rendezvous but skip the column loops
        const int nkc_tile = NKC;
        const int NKC = ((uint32_t)(q * 32) * 16u < KGS) ? nkc_tile : 0;
        constexpr int NCG = NUM_WORKERS / 128;       // chunk groups (warps per lane quarter)
        const int row = q * 32 + lane;               // tile row == TMEM lane
        const bool row_ok = row < rows;
        const int grow = row_ok ? row0 + row : row0; // global node id (clamped for padding rows)
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        uint32_t ph_done = 0;                         // parity of bar_mma_done
        bool ok = true;
        uint32_t ph_g1 = 0;                           // bit b = parity of bar_g1_done[b]
        // All-worker rendezvous on hardware named barrier 1 (blocked warps issue nothing).  Waiting for the MMA warp is
        // done by worker warp 0 alone (one poller instead of sixteen); the others block on the named barrier behind it.
        // On a timeout the poller raises s_abort BEFORE the rendezvous, so every worker leaves together.
        auto workers_sync = [&]() {
            asm volatile("bar.sync 1, %0;" ::"n"(NUM_WORKERS) : "memory");
            if (*abortp) ok = false;
        };
        int* ev_ip = nullptr; int* ev_sp = nullptr;   // set below (the event log of thread 0)
        auto ev_raw = [&](int code) {
            if (GGNN_TRACE_ON(p) && ev_ip && tile == 0 && tid == 0 && *ev_sp == 1 && *ev_ip < 64) { p.dbg[64 + 2 * *ev_ip] = code; p.dbg[65 + 2 * *ev_ip] = clock64(); ++*ev_ip; }
        };
        auto wait_on = [&](uint64_t* bar, uint32_t parity) {
            ev_raw(1);
            if (warp == 0 && ok) { if (!mbar_wait(bar, parity, abortp)) *abortp = 1; }
            ev_raw(2);
            asm volatile("bar.sync 1, %0;" ::"n"(NUM_WORKERS) : "memory");
            ev_raw(3);
            if (*abortp) ok = false;
            if (ok) tc_fence_after();
        };
        auto wait_mma = [&]() { wait_on(&bar_mma_done, ph_done & 1); ++ph_done; };
        auto wait_g1 = [&](int b) { wait_on(&bar_g1_done[b], (ph_g1 >> b) & 1u); ph_g1 ^= 1u << b; };
        // every lane orders its own smem/TMEM writes, then one lane per warp signals the MMA thread
        auto publish = [&]() { tmem_st_wait(); tc_fence_before(); fence_async_smem(); __syncwarp(); if (lane == 0) mbar_arrive(&bar_a_ready); };
        auto publish_g = [&](int b) { tc_fence_before(); fence_async_smem(); __syncwarp(); if (lane == 0) mbar_arrive(&bar_g_ready[b]); };
        const float zeros8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float* res_pre = p.res_pre + (size_t)tile * TILE_M * 3 * DP + (size_t)row * 3 * DP;
        int dbg_i = 0;
        int ev_i = 0, ev_step = -1;
        ev_ip = &ev_i; ev_sp = &ev_step;
        auto ev = [&](int code) {
            if (GGNN_TRACE_ON(p) && tile == 0 && tid == 0 && ev_step == 1 && ev_i < 64) { p.dbg[64 + 2 * ev_i] = code; p.dbg[65 + 2 * ev_i] = clock64(); ++ev_i; }
        };
        auto stamp = [&]() { if (GGNN_TRACE_ON(p) && tile == 0 && tid == 0 && dbg_i < 40) p.dbg[dbg_i++] = clock64(); ev(9); };
        stamp();   // 0: start

        // ---- initial state: global fp32 -> TMEM (fp32 master) + opH (bf16 hi/lo)
        {
            const float* hin = (LOCAL ? p.state[0] : p.g_in) + (size_t)grow * D;
            for (int kc = cg; kc < NKC; kc += NCG) {
                float v[8];
                load8_guarded_cg(hin, kc * 8, row_ok ? D : 0, v);
                tmem_st8(TM_H + lane_addr + kc * 8, v);
                store_operand_chunk(opH, KGS, PART_B, kc, row, v);
            }
            for (int c = cg; c < 3 * NKC; c += NCG) tmem_st8(TM_ACC + lane_addr + c * 8, zeros8);   // acc + gate accumulators start at zero
            tmem_st_wait();
        }
        const bool csr_smem = LOCAL && p.csr_cache && p.gather_mode == GATHER_SPARSE;
        if (csr_smem) {
            const int base = p.row_ptr[(size_t)row0 * T];
            const int nptr = rows * T + 1;
            for (int i = tid; i < TILE_M * T + 1; i += NUM_WORKERS) sRowPtr[i] = (uint16_t)(p.row_ptr[(size_t)row0 * T + min(i, nptr - 1)] - base);
            const int mt = p.row_ptr[(size_t)(row0 + rows) * T] - base;
            for (int i = tid; i < mt; i += NUM_WORKERS) sSrc[i] = (uint8_t)(p.csr_src[base + i] - row0);
        }
        workers_sync();
        stamp();   // 1: state loaded

        for (int l = l_begin; l < l_end && ok; ++l) {
            const TcLayer& ly = p.layer[l];
            const int s_begin = LOCAL ? 0 : p.g_step;
            const int s_end = LOCAL ? ly.steps : p.g_step + 1;
            const bool gru = p.cell == CELL_GRU;
            // ---- this layer's cell biases -> shared (zero padded to DP)
            for (int i = tid; i < 3 * DP; i += NUM_WORKERS) {
                const int blk = i / DP, col = i - blk * DP;
                float b = 0.0f;
                if (col < D) b = (blk < 2) ? (gru ? ly.gate_b[blk * D + col] : 0.0f) : ly.cand_b[col];
                sBias[i] = b;
            }
            workers_sync();
            // ---- residual pre-products (constant over the layer's timesteps): Pg = res . K_g[res rows], Pc = res . K_c[res rows]
            if (ly.nres > 0 && s_end > s_begin) {
                for (int i = 0; i < ly.nres && ok; ++i) {
                    if (i > 0) wait_mma();
                    if (!ok) break;
                    const float* rs = p.state[ly.res[i]] + (size_t)grow * D;
                    for (int kc = cg; kc < NKC; kc += NCG) {
                        float v[8];
                        load8_guarded_cg(rs, kc * 8, row_ok ? D : 0, v);
                        store_operand_chunk(opA, KGS, PART_B, kc, row, v);
                    }
                    publish();
                }
                wait_mma();
                if (ok) {
                    const int ngate = gru ? 2 * NKC : 0;
                    for (int c = cg; c < ngate + NKC; c += NCG) {   // gate chunks [0,2NKC) then cand chunks
                        float v[8];
                        const uint32_t ta = (c < ngate ? TM_GATE + c * 8 : TM_ACC + (c - ngate) * 8) + lane_addr;
                        tmem_ld8(ta, v);
                        tmem_st8(ta, zeros8);
                        float* dst = res_pre + (c < ngate ? c * 8 : 2 * DP + (c - ngate) * 8);
                        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                        *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
                    }
                    tmem_st_wait();
                    tc_fence_before();
                    workers_sync();
                }
            }
            for (int s = s_begin; s < s_end && ok; ++s) {
                ev_step = (l == l_begin) ? s - s_begin : -1;
                ev(8);   // step start
                const size_t save_off = (size_t)(p.step_base[l] + s) * VD + (size_t)grow * D;
                // ------------------------------------------------------------ gather + G1 per present type
                // A_t tiles alternate between opA and opX so that gathering type n+1 overlaps the MMAs of type n
                int nty = 0;
                // The gather touches shared memory only (no TMEM lane-quarter restriction), so with compact tiles -- rows live in lane
                // quarters 0 and 1 only -- all 16 worker warps share it: row group = warp & 1, eight column-chunk groups instead of four.
                const bool gsplit = KGS == 1024u;
                const int g_row = gsplit ? ((warp & 1) * 32 + lane) : row;
                const bool g_row_ok = g_row < rows;
                const int g_grow = g_row_ok ? row0 + g_row : row0;
                const int g_cg = gsplit ? (warp >> 1) : cg, g_ncg = gsplit ? NUM_WORKERS / 64 : NCG;
                const int g_nkc = gsplit ? nkc_tile : NKC;
                for (int t = 0; t < T && ok; ++t) {
                    if (!((tmask >> t) & 1u)) continue;
                    const int b = nty % NGB;
                    if (nty >= NGB) { wait_g1(b); if (!ok) break; }   // the MMAs that read this buffer NGB types ago are done
                    uint8_t* gdst = b == 0 ? opA : (b == 1 ? opX : opA + (size_t)OPB * (size_t)(b - 1));
                    ++nty;
                    int beg = 0, end = 0, dg = 0, di = 0;
                    if (g_row_ok) {
                        if (p.gather_mode == GATHER_SPARSE) {
                            if (csr_smem) { beg = sRowPtr[g_row * T + t]; end = sRowPtr[g_row * T + t + 1]; }
                            else { beg = p.row_ptr[(size_t)g_grow * T + t]; end = p.row_ptr[(size_t)g_grow * T + t + 1]; }
                        } else { dg = g_grow / p.dense_v; di = g_grow - dg * p.dense_v; }
                    }
                    for (int kc = g_cg; kc < g_nkc; kc += g_ncg) {
                        float acc[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
                        if (LOCAL && csr_smem) {
                            // fast path: local source indices from shared memory, two messages in flight
                            const uint8_t* colbase = opH + (size_t)kc * KGS;
                            const uint32_t lo_off = PART_B;
                            int m = beg;
                            if (p.nparts == 3) {
                                for (; m + 1 < end; m += 2) {
                                    const uint8_t* s0 = colbase + (size_t)sSrc[m] * 16;
                                    const uint8_t* s1 = colbase + (size_t)sSrc[m + 1] * 16;
                                    const uint4 h0 = *reinterpret_cast<const uint4*>(s0), l0 = *reinterpret_cast<const uint4*>(s0 + lo_off);
                                    const uint4 h1 = *reinterpret_cast<const uint4*>(s1), l1 = *reinterpret_cast<const uint4*>(s1 + lo_off);
                                    unpack8_add(h0, acc, 1.0f); unpack8_add(h1, acc, 1.0f);
                                    unpack8_add(l0, acc, 1.0f); unpack8_add(l1, acc, 1.0f);
                                }
                                if (m < end) {
                                    const uint8_t* s0 = colbase + (size_t)sSrc[m] * 16;
                                    unpack8_add(*reinterpret_cast<const uint4*>(s0), acc, 1.0f);
                                    unpack8_add(*reinterpret_cast<const uint4*>(s0 + lo_off), acc, 1.0f);
                                }
                            } else {
                                for (; m < end; ++m) unpack8_add(*reinterpret_cast<const uint4*>(colbase + (size_t)sSrc[m] * 16), acc, 1.0f);
                            }
                        } else if (p.gather_mode == GATHER_SPARSE) {
                            for (int m = beg; m < end; ++m) {
                                if (LOCAL) {
                                    const int sl = csr_smem ? (int)sSrc[m] : (p.csr_src[m] - row0);
                                    const uint8_t* sp = opH + (size_t)kc * KGS + (size_t)sl * 16;
                                    unpack8_add(*reinterpret_cast<const uint4*>(sp), acc, 1.0f);
                                    if (p.nparts == 3) unpack8_add(*reinterpret_cast<const uint4*>(sp + PART_B), acc, 1.0f);
                                } else {
                                    // GLOBAL mode: source rows come from the previous step's fp32 state in L2; keep 4 rows in flight
                                    float hv[4][8];
                                    const int nb = min(4, end - m);
#pragma unroll
                                    for (int q = 0; q < 4; ++q)
                                        if (q < nb) load8_guarded_cg(p.g_in + (size_t)p.csr_src[m + q] * D, kc * 8, D, hv[q]);
#pragma unroll
                                    for (int q = 0; q < 4; ++q)
                                        if (q < nb) {
#pragma unroll
                                            for (int j = 0; j < 8; ++j) acc[j] += hv[q][j];
                                        }
                                    m += nb - 1;
                                }
                            }
                        } else if (g_row_ok) {
                            const int nv = p.dense_v;
                            const float* arow = p.dense_adj + (((size_t)dg * T + t) * nv + di) * nv;
                            for (int jn = 0; jn < nv; ++jn) {
                                const float a = arow[jn];
                                if (a != 0.0f) {
                                    const int src = dg * nv + jn;
                                    if (LOCAL) {
                                        const uint8_t* sp = opH + (size_t)kc * KGS + (size_t)(src - row0) * 16;
                                        unpack8_add(*reinterpret_cast<const uint4*>(sp), acc, a);
                                        if (p.nparts == 3) unpack8_add(*reinterpret_cast<const uint4*>(sp + PART_B), acc, a);
                                    } else {
                                        float hv[8];
                                        load8_guarded_cg(p.g_in + (size_t)src * D, kc * 8, D, hv);
#pragma unroll
                                        for (int j = 0; j < 8; ++j) acc[j] = fmaf(a, hv[j], acc[j]);
                                    }
                                }
                            }
                        }
                        store_operand_chunk(gdst, KGS, PART_B, kc, g_row, acc);
                    }
                    publish_g(b);
                    stamp();   // gather of one type done
                }
                if (!ok) break;
                const bool have_msgs = nty > 0;
                // all G1 MMAs complete (last use of each gather buffer) before opX is rewritten / the accumulators are read
                for (int i = max(0, nty - NGB); i < nty && ok; ++i) wait_g1(i % NGB);
                if (!ok) break;
                stamp();   // G1 done (agg accumulators ready)
                // ------------------------------------------------------------ agg epilogue: + indeg.B, / (deg + 1e-7) -> opX
                {
                    const float inv_den = (p.use_avg && row_ok) ? __fdividef(1.0f, p.denom[grow]) : 1.0f;
                    for (int kc = cg; kc < NKC; kc += NCG) {
                        float v[8];
                        if (have_msgs) { tmem_ld8(TM_ACC + lane_addr + kc * 8, v); tmem_st8(TM_ACC + lane_addr + kc * 8, zeros8); }
                        else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = 0.0f;
                        }
                        if (p.use_bias) {
                            for (int t = 0; t < T; ++t) {
                                const float ind = p.indeg[(size_t)grow * T + t];
                                float b[8];
                                load8_guarded(ly.edge_b + (size_t)t * D, kc * 8, D, b);
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[j] = fmaf(ind, b[j], v[j]);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = row_ok ? v[j] * inv_den : 0.0f;
                        if (p.save && row_ok) store8_guarded(p.save_buf.agg + save_off, kc * 8, D, v);
                        store_operand_chunk(opX, KGS, PART_B, kc, row, v);
                    }
                    publish();
                }
                stamp();   // agg epilogue done
                if (gru) {
                    // -------------------------------------------------------- r ready: r*h -> opA  (u and the candidate's agg part keep the tensor core busy meanwhile)
                    wait_mma(); if (!ok) break;
                    stamp();   // gates ready
                    for (int kc = cg; kc < NKC; kc += NCG) {
                        float g[8], h[8], b[8], rh[8];
                        tmem_ld8_nowait(TM_GATE + lane_addr + kc * 8, g);
                        tmem_ld8_nowait(TM_H + lane_addr + kc * 8, h);
                        lds8(sBias + kc * 8, b);
                        if (ly.nres > 0) {
                            float rp[8];
                            lds8(res_pre + kc * 8, rp);   // generic load (global)
#pragma unroll
                            for (int j = 0; j < 8; ++j) b[j] += rp[j];
                        }
                        tmem_ld_wait();
                        tmem_st8(TM_GATE + lane_addr + kc * 8, zeros8);
#pragma unroll
                        for (int j = 0; j < 8; ++j) { g[j] = sigmoid_fast(g[j] + b[j]); rh[j] = g[j] * h[j]; }
                        if (p.save && row_ok) {
                            store8_guarded(p.save_buf.r + save_off, kc * 8, D, g);
                            store8_guarded(p.save_buf.h_in + save_off, kc * 8, D, h);
                        }
                        store_operand_chunk(opA, KGS, PART_B, kc, row, rh);
                    }
                    publish();
                    stamp();   // r*h epilogue done
                    // u = sigmoid(gate_u + b_u) while the candidate's recurrent MMAs run; parked in the gate-u accumulator columns (the candidate
                    // GEMM only touches the agg/cand accumulator) so that the state update below is just tanh + blend
                    for (int kc = cg; kc < NKC; kc += NCG) {
                        float u[8], bu[8];
                        tmem_ld8_nowait(TM_GATE + lane_addr + DP + kc * 8, u);
                        lds8(sBias + DP + kc * 8, bu);
                        if (ly.nres > 0) {
                            float rp[8];
                            lds8(res_pre + DP + kc * 8, rp);
#pragma unroll
                            for (int j = 0; j < 8; ++j) bu[j] += rp[j];
                        }
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 8; ++j) u[j] = sigmoid_fast(u[j] + bu[j]);
                        if (p.save && row_ok) store8_guarded(p.save_buf.u + save_off, kc * 8, D, u);
                        tmem_st8(TM_GATE + lane_addr + DP + kc * 8, u);
                    }
                    tmem_st_wait();
                }
                // ------------------------------------------------------------ candidate ready: new state
                wait_mma(); if (!ok) break;
                stamp();   // candidate ready
                {
                    const bool last_of_layer = (s == ly.steps - 1);
                    float* outp = LOCAL ? (last_of_layer ? p.state_w[l + 1] : nullptr) : p.g_out;
                    for (int kc = cg; kc < NKC; kc += NCG) {
                        float c[8], h[8], u[8], bc[8], hn[8];
                        tmem_ld8_nowait(TM_ACC + lane_addr + kc * 8, c);
                        lds8(sBias + 2 * DP + kc * 8, bc);
                        if (gru) {
                            tmem_ld8_nowait(TM_GATE + lane_addr + DP + kc * 8, u);   // already sigmoid(gate_u + b_u), see above
                            tmem_ld8_nowait(TM_H + lane_addr + kc * 8, h);
                            if (ly.nres > 0) {
                                float rp[8];
                                lds8(res_pre + 2 * DP + kc * 8, rp);
#pragma unroll
                                for (int j = 0; j < 8; ++j) bc[j] += rp[j];
                            }
                            tmem_ld_wait();
                            tmem_st8(TM_ACC + lane_addr + kc * 8, zeros8);
                            tmem_st8(TM_GATE + lane_addr + DP + kc * 8, zeros8);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                c[j] = act_fast(c[j] + bc[j], p.act);
                                hn[j] = fmaf(u[j], h[j] - c[j], c[j]);   // u*h + (1-u)*c
                            }
                            if (p.save && row_ok) store8_guarded(p.save_buf.c + save_off, kc * 8, D, c);
                        } else {
                            if (p.save) tmem_ld8_nowait(TM_H + lane_addr + kc * 8, h);
                            if (ly.nres > 0) {
                                float rp[8];
                                lds8(res_pre + 2 * DP + kc * 8, rp);
#pragma unroll
                                for (int j = 0; j < 8; ++j) bc[j] += rp[j];
                            }
                            tmem_ld_wait();
                            tmem_st8(TM_ACC + lane_addr + kc * 8, zeros8);
#pragma unroll
                            for (int j = 0; j < 8; ++j) hn[j] = act_fast(c[j] + bc[j], p.act);
                            if (p.save && row_ok) store8_guarded(p.save_buf.h_in + save_off, kc * 8, D, h);
                        }
                        if (p.drop_keep < 1.0f) {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                hn[j] = dropout_apply(hn[j], p.drop_seed, p.step_base[l] + s, p.V, D, grow, kc * 8 + j, p.drop_keep);
                        }
                        tmem_st8(TM_H + lane_addr + kc * 8, hn);
                        store_operand_chunk(opH, KGS, PART_B, kc, row, hn);
                        if (outp && row_ok) store8_guarded(outp + (size_t)grow * D, kc * 8, D, hn);
                    }
                    tmem_st_wait();
                    tc_fence_before();
                    workers_sync();   // opH complete before anyone gathers from it
                    stamp();   // state update done
                }
            }  // steps
            if (LOCAL && ok && ly.steps == 0) {   // a layer without timesteps aliases the previous state (sparse:152)
                for (int kc = cg; kc < NKC; kc += NCG) {
                    float h[8];
                    tmem_ld8(TM_H + lane_addr + kc * 8, h);
                    if (row_ok) store8_guarded(p.state_w[l + 1] + (size_t)grow * D, kc * 8, D, h);
                }
            }
            if (LOCAL && l + 1 < l_end) { __threadfence(); workers_sync(); }   // layer output visible before it is read as a residual
        }  // layers
        if (!ok && lane == 0) atomicExch(p.error_flag, 1);
    } else if (warp < WARP_PROD) {   // issuers
        // =============================================================================== MMA ISSUERS
        // Each issuer warp runs the (warp-uniform) control flow; one elected lane issues the tcgen05 instructions.  K-steps are
        // dealt round-robin to the issuer warps; all of them (and the producer) walk the same sequence of weight slots.
        {
            const bool x3 = p.nparts == 3;
            const uint32_t nstg = (uint32_t)nst;
            const uint32_t iw = (uint32_t)(warp - WARP_MMA);
            const uint32_t a_lo16 = PART_B >> 4;                   // hi -> lo part of an A operand (16-byte units)
            const uint32_t a_k16 = KGS >> 3;                       // one K-step (16 columns = two k-groups) of an A operand (16-byte units)
            const uint32_t b_lo16 = ((uint32_t)DP * 32u) >> 4;    // hi -> lo part of a narrow weight slot
            const uint32_t stage16 = STAGE_B >> 4;
            const uint64_t descA = make_desc(0, KGS, 128);
            const uint64_t descB1 = make_desc(0, 16u * (uint32_t)DP, 128);   // N = DP   operand: K-chunk stride 16*DP
            const uint64_t descB2 = make_desc(0, 32u * (uint32_t)DP, 128);   // N = 2*DP operand: K-chunk stride 32*DP
            const uint32_t opH16 = smem_u32(opH) >> 4, opX16 = smem_u32(opX) >> 4, opA16 = smem_u32(opA) >> 4, ring16 = smem_u32(ring) >> 4;
            const uint32_t full0 = smem_u32(&bar_w_full[0]), empty0 = smem_u32(&bar_w_empty[0]);
            const uint32_t idesc = make_idesc_bf16(DP), idesc2 = make_idesc_bf16(2 * DP);
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
            const bool dbg_on = GGNN_TRACE_ON(p) && tile == 0 && lane == 0 && iw == 0;
            long long dbg_ready = 0;
            const long long dbg_t0 = clock64();
            int dbg_g = 0;
            int iev_i = 0, iev_step = -1;
            auto iev = [&](int code) {
                if (dbg_on && iev_step == 1 && iev_i < 64) { p.dbg[192 + 2 * iev_i] = code; p.dbg[193 + 2 * iev_i] = clock64(); ++iev_i; }
            };
            uint32_t ph_ready = 0, ph_g = 0;
            uint32_t cur[3] = {0, 0, 0};   // round-robin cursor of each slot set
            uint32_t fpar = 0;             // bit s = parity to wait for on bar_w_full[s]
            bool ok = true;
            auto next_slot = [&](int set) -> uint32_t {
                const uint32_t n = (set == SET_BASE) ? nstg : nstg + EXT;
                const uint32_t i = cur[set];
                cur[set] = (i + 1 == n) ? 0u : i + 1;
                return (set == SET_XH && i >= nstg) ? i + EXT : i;
            };
            auto slot16 = [&](uint32_t sl) -> uint32_t {
                return sl < nstg ? ring16 + sl * 2u * stage16
                     : sl < nstg + EXT ? opA16 + (sl - nstg) * 2u * stage16 : opH16 + (sl - nstg - EXT) * 2u * stage16;
            };
            auto wait_full = [&](uint32_t sl) {
                const uint32_t par = (fpar >> sl) & 1u;
                fpar ^= 1u << sl;
                const uint32_t fb = full0 + sl * 8u;
                if (!mbar_try(fb, par)) {
                    if (!mbar_wait_slow(fb, par, abortp)) ok = false;
                    ok = __all_sync(0xffffffffu, ok);
                }
            };
            auto commit_empty = [&](uint32_t sl) {
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty0 + sl * 8u) : "memory");
            };
            // acc(tm_col) += A(op) . B(next weight slots of `set`); accumulators are zeroed by the workers.  A slot = two stages:
            //   narrow (N = DP):   K-steps 2i and 2i+1 of a DP-wide block, each stage = [hi | lo] halves        (odd tail: one stage)
            //   wide   (N = 2*DP): one K-step of the [r | u] gate block, stage 0 = hi part, stage 1 = lo part
            // Each slot belongs to one issuer warp (slot index modulo the number of issuers).
            auto gemm = [&](uint32_t op16, uint32_t tm_col, bool wide, int set) {
                if (!ok) return;
                if (dbg_on && dbg_g < 20) p.dbg[40 + dbg_g++] = clock64();   // start of each of the first 20 GEMM blocks
                iev(10);
                bool first_w = true;
                const uint32_t tm_d = tmem_u + tm_col;
                const int nslots = wide ? NKS : (NKS + 1) / 2;
#pragma unroll 1
                for (int i = 0; i < nslots; ++i) {
                    const uint32_t sl = next_slot(set);
                    // STATIC slot ownership: every use of a slot is consumed by the same issuer, so that issuer observes each
                    // phase of the slot's barrier in order.  (Parity waits alias if a waiter can be two phases ahead, which
                    // round-robin ownership allows when bulk copies complete out of order or the ring is a single slot.)
                    if (sl % NUM_ISSUERS != iw) continue;
                    wait_full(sl);
                    if (!ok) return;
                    if (first_w) { iev(11); first_w = false; }
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t b16 = slot16(sl);
                        if (wide) {
                            const uint64_t ad = descA | (uint64_t)(op16 + (uint32_t)i * a_k16);
                            const uint64_t bh = descB2 | (uint64_t)b16, bl = descB2 | (uint64_t)(b16 + stage16);
                            umma_bf16(tm_d, ad, bh, idesc2, 1u);
                            if (x3) {
                                umma_bf16(tm_d, ad, bl, idesc2, 1u);
                                umma_bf16(tm_d, ad + a_lo16, bh, idesc2, 1u);
                            }
                        } else {
                            const int nk = (2 * i + 1 < NKS) ? 2 : 1;
                            for (int h = 0; h < nk; ++h) {
                                const uint64_t ad = descA | (uint64_t)(op16 + (uint32_t)(2 * i + h) * a_k16);
                                const uint64_t bd = descB1 | (uint64_t)(b16 + (uint32_t)h * stage16);
                                umma_bf16(tm_d, ad, bd, idesc, 1u);
                                if (x3) {
                                    umma_bf16(tm_d, ad, bd + b_lo16, idesc, 1u);
                                    umma_bf16(tm_d, ad + a_lo16, bd, idesc, 1u);
                                }
                            }
                        }
                        commit_empty(sl);   // slot reusable once these MMAs have read it
                    }
                    __syncwarp();
                }
                iev(12);
            };
            auto commit_to = [&](uint64_t* bar) { if (ok) { if (elect_one()) umma_commit(bar); __syncwarp(); } };
            auto wait_bar = [&](uint32_t rb, uint32_t par) {
                if (!ok) return;
                iev(13);
                if (!mbar_try(rb, par)) {
                    const long long w0 = dbg_on ? clock64() : 0;
                    if (!mbar_wait_slow(rb, par, abortp)) ok = false;
                    if (dbg_on) dbg_ready += clock64() - w0;
                }
                ok = __all_sync(0xffffffffu, ok);
                tc_fence_after();
                iev(14);
            };
            auto wait_ready = [&]() { wait_bar(smem_u32(&bar_a_ready), ph_ready & 1); ++ph_ready; };
            auto wait_g_ready = [&](int b) { wait_bar(smem_u32(&bar_g_ready[b]), (ph_g >> b) & 1u); ph_g ^= 1u << b; };
            for (int l = l_begin; l < l_end && ok; ++l) {
                const TcLayer& ly = p.layer[l];
                const int s_begin = LOCAL ? 0 : p.g_step;
                const int s_end = LOCAL ? ly.steps : p.g_step + 1;
                const int nres = ly.nres;
                const bool gru = p.cell == CELL_GRU;
                if (nres > 0 && s_end > s_begin) {
                    for (int i = 0; i < nres && ok; ++i) {
                        wait_ready();
                        if (gru) gemm(opA16, 2u * DP, true, SET_BASE);
                        gemm(opA16, (uint32_t)DP, false, SET_BASE);
                        commit_to(&bar_mma_done);
                    }
                }
                for (int s = s_begin; s < s_end && ok; ++s) {
                    iev_step = (l == l_begin) ? s - s_begin : -1;
                    int nty = 0;
                    for (int t = 0; t < T && ok; ++t) {
                        if (!((tmask >> t) & 1u)) continue;
                        const int b = nty % NGB;
                        wait_g_ready(b);
                        gemm(b == 0 ? opA16 : (b == 1 ? opX16 : opA16 + (uint32_t)(b - 1) * (OPB >> 4)), (uint32_t)DP, false, SET_BASE);
                        commit_to(&bar_g1_done[b]);
                        ++nty;
                    }
                    commit_to(&bar_xa_free);                        // every MMA that reads the opA tile (this step's G1, last step's G3) is tracked
                    wait_ready();                                   // agg operand (opX) ready
                    if (gru) {
                        gemm(opX16, 2u * DP, true, SET_XA); gemm(opH16, 2u * DP, true, SET_XA);   // [r | u] in one N = 2*DP MMA stream
                        commit_to(&bar_mma_done);                   // gates ready
                        commit_to(&bar_xh_free);                    // ... and the opH tile is dead until the state update
                        gemm(opX16, (uint32_t)DP, false, SET_XH);   // candidate, agg part: overlaps the r*h epilogue
                        wait_ready();                               // r*h operand (opA) ready
                        gemm(opA16, (uint32_t)DP, false, SET_XH);
                    } else {
                        gemm(opX16, (uint32_t)DP, false, SET_XA); gemm(opH16, (uint32_t)DP, false, SET_XA);
                    }
                    commit_to(&bar_mma_done);                       // candidate ready
                }
            }
            if (!ok && lane == 0) atomicExch(p.error_flag, 2);
            if (dbg_on) { p.dbg[60] = 0; p.dbg[61] = dbg_ready; p.dbg[62] = clock64() - dbg_t0; }
        }
    } else {
        // =============================================================================== WEIGHT PRODUCER
        // One thread per producer warp.  The stream is bound by the per-push issue cost (wait, expect_tx, bulk copy: ~430 cycles),
        // so a push moves a whole slot = two stages (2 x 64*DP bytes, contiguous in the pre-tiled weights) with one bulk copy, and the
        // slots are dealt statically to NUM_PRODUCERS threads (a slot always has the same producer and the same issuer, so every mbarrier
        // is waited on phase by phase -- parity waits alias when an agent can lap another).
        if (lane == 0) {
            const uint32_t nstg = (uint32_t)nst;
            const uint32_t pw = (uint32_t)(warp - WARP_PROD);   // every producer walks the whole push sequence and fills the slots it owns
            uint32_t cur[3] = {0, 0, 0};
            uint32_t used = 0, epar = 0;   // bit s: slot s has been filled before / parity to wait for on bar_w_empty[s]
            uint32_t ph_xa = 0, ph_xh = 0;
            bool ok = true;
            int pev_i = 0, pev_step = -1;
            auto pev = [&](int code) {
                if (GGNN_TRACE_ON(p) && tile == 0 && pw == 0 && pev_step == 1 && pev_i < 64) { p.dbg[320 + 2 * pev_i] = code; p.dbg[321 + 2 * pev_i] = clock64(); ++pev_i; }
            };
            auto next_slot = [&](int set) -> uint32_t {
                const uint32_t n = (set == SET_BASE) ? nstg : nstg + EXT;
                const uint32_t i = cur[set];
                cur[set] = (i + 1 == n) ? 0u : i + 1;
                return (set == SET_XH && i >= nstg) ? i + EXT : i;
            };
            auto slot_ptr = [&](uint32_t sl) -> uint8_t* {
                return sl < nstg ? ring + sl * 2u * STAGE_B
                     : sl < nstg + EXT ? opA + (sl - nstg) * 2u * STAGE_B : opH + (sl - nstg - EXT) * 2u * STAGE_B;
            };
            // stream `nstages` consecutive 64*DP-byte stages of a pre-tiled matrix, two per slot, into the next slots of `set`
            auto push = [&](const uint8_t* src, int nstages, int set) {
                pev(20);
                for (int i = 0; i < nstages && ok; i += 2) {
                    const uint32_t bytes = (i + 1 < nstages) ? 2u * STAGE_B : STAGE_B;
                    const uint32_t sl = next_slot(set);
                    if (sl % NUM_PRODUCERS != pw) continue;   // static ownership: each slot's barriers see ONE producer, in order
                    // a slot is free once the MMAs that read its previous contents have completed (first use: free)
                    if ((used >> sl) & 1u) {
                        if (!mbar_wait(&bar_w_empty[sl], (epar >> sl) & 1u, abortp)) { ok = false; break; }
                        epar ^= 1u << sl;
                    }
                    used |= 1u << sl;
                    mbar_arrive_expect_tx(&bar_w_full[sl], bytes);
                    bulk_copy_g2s(slot_ptr(sl), src + (size_t)i * STAGE_B, bytes, &bar_w_full[sl]);
                }
                pev(21);
            };
            for (int l = l_begin; l < l_end && ok; ++l) {
                const TcLayer& ly = p.layer[l];
                const int s_begin = LOCAL ? 0 : p.g_step;
                const int s_end = LOCAL ? ly.steps : p.g_step + 1;
                const bool gru = p.cell == CELL_GRU;
                const size_t blk = (size_t)NKS * STAGE_B;   // one DP x DP block; a gate block is two of these per K segment
                if (ly.nres > 0 && s_end > s_begin) {
                    for (int i = 0; i < ly.nres && ok; ++i) {
                        if (gru) push(ly.w_gate + (size_t)i * 2 * blk, 2 * NKS, SET_BASE);
                        push(ly.w_cand + (size_t)i * blk, NKS, SET_BASE);
                    }
                }
                const size_t kx = (size_t)ly.nres, kh = (size_t)ly.nres + 1;
                for (int s = s_begin; s < s_end && ok; ++s) {
                    pev_step = (l == l_begin) ? s - s_begin : -1;
                    for (int t = 0; t < T && ok; ++t) {
                        if (!((tmask >> t) & 1u)) continue;
                        push(ly.w_edge + (size_t)t * blk, NKS, SET_BASE);
                    }
                    // the opA tile joins the ring once no MMA reads it any more (this step's G1 / last step's candidate are complete)
                    pev(22);
                    if (ok) { ok = mbar_wait(&bar_xa_free, ph_xa & 1, abortp); ++ph_xa; }
                    pev(23);
                    if (gru) {
                        push(ly.w_gate + kx * 2 * blk, 2 * NKS, SET_XA); push(ly.w_gate + kh * 2 * blk, 2 * NKS, SET_XA);
                        // the opH tile joins the ring once the gate GEMM (its last reader) is complete
                        pev(24);
                        if (ok) { ok = mbar_wait(&bar_xh_free, ph_xh & 1, abortp); ++ph_xh; }
                        pev(25);
                        push(ly.w_cand + kx * blk, NKS, SET_XH); push(ly.w_cand + kh * blk, NKS, SET_XH);
                    } else {
                        push(ly.w_cand + kx * blk, NKS, SET_XA); push(ly.w_cand + kh * blk, NKS, SET_XA);
                    }
                }
            }
            if (!ok) atomicExch(p.error_flag, 3);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_PROD) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ weight pre-tiling
// fp32 row-major W[(nseg*D) rows][src_ld cols], columns [src_col0, src_col0 + nblk*D)  ->  per K-step s (16 padded rows) nblk consecutive ring slots of
// 64*DP bytes.  Np = nblk*DP output rows (column block b of the source lands at n = b*DP + col, zero padded):
//   nblk == 1:  slot = [hi part | lo part], part = 2 K-chunks x DP rows x 16 B            (one DP-wide weight block)
//   nblk == 2:  slot 0 = hi part, slot 1 = lo part, part = 2 K-chunks x 2*DP rows x 16 B  (the [r | u] gate block)
// element: byte(s, part, c, n, j) = s*nblk*64*DP + part*(32*Np) + c*(16*Np) + n*16 + j*2 = bf16 part of W[row(s*16+c*8+j)][col(n)]
__global__ void ggnn_tile_weights_kernel(const float* __restrict__ W, uint8_t* __restrict__ out, int D, int DP, int nseg, int nblk,
                                         int src_ld, int src_col0) {
    const int Np = nblk * DP;
    const int ksteps = nseg * DP / 16;
    const long long total = (long long)ksteps * 2 * Np;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(idx % Np);
        const int c = (int)((idx / Np) % 2);
        const int s = (int)(idx / (2 * Np));
        const int blk = n / DP, nn = n - blk * DP;
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int kp = s * 16 + c * 8 + j;
            const int seg = kp / DP, kk = kp - seg * DP;
            x[j] = (kk < D && nn < D) ? W[(size_t)(seg * D + kk) * src_ld + src_col0 + blk * D + nn] : 0.0f;
        }
        uint4 hi, lo;
        split8(x, hi, lo);
        uint8_t* base = out + (size_t)s * nblk * 64 * DP + (size_t)c * 16 * Np + (size_t)n * 16;
        *reinterpret_cast<uint4*>(base) = hi;
        *reinterpret_cast<uint4*>(base + (size_t)32 * Np) = lo;
    }
}

}  // namespace tc
}  // namespace ggnn
