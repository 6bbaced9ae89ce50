// Standalone hardware probe (not part of the product): validates the tcgen05 shared-memory / instruction
// descriptor encodings used by the engine's tensor-core path and measures the pipes that bound the design.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe tc_probe.cu && ./tc_probe
// Every mbarrier wait is bounded (a failed wait prints TIMEOUT and the kernel exits cleanly).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity, int max_iters = 1 << 20) {
    for (int i = 0; i < max_iters; ++i) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

// K-major, no swizzle canonical layout: element (row r, k) of a [R x K] bf16 operand lives at
//   (k/8)*LBO + (r/8)*SBO + (r%8)*16 + (k%8)*2   bytes, with SBO = 128 and LBO = R*16.
__host__ __device__ inline uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
    return d;                // base_offset = 0, lbo_mode = 0, layout_type = 0 (SWIZZLE_NONE)
}

// kind::f16, A/B = BF16 (1) or kind::tf32 A/B = TF32 (2); D = F32; both K-major; M, N as given.
__host__ __device__ inline uint32_t make_idesc(int M, int N, int ab_format) {
    uint32_t d = 0;
    d |= 1u << 4;                       // c_format = F32
    d |= (uint32_t)ab_format << 7;      // a_format
    d |= (uint32_t)ab_format << 10;     // b_format
    d |= (uint32_t)(N >> 3) << 17;      // n_dim
    d |= (uint32_t)(M >> 4) << 24;      // m_dim
    return d;
}

template <int KIND>  // 0: kind::f16 (bf16)   1: kind::tf32
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                     :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                     :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}

// D[128 x N] = A[128 x K] . B[K x N]; A given row-major fp32 [128][K], B row-major fp32 [K][N].
// KIND 0: operands rounded to bf16 (2-byte elements, 8 per 16-byte core row, MMA K = 16)
// KIND 1: operands as tf32 (4-byte elements, 4 per core row, MMA K = 8)
template <int KIND>
__global__ void __launch_bounds__(128) umma_probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ Dout,
                                                  int N, int K, int reps, long long* cycles, int* status) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int ESZ = KIND == 0 ? 2 : 4;
    constexpr int EPC = 16 / ESZ;        // elements per 16-byte core-matrix row
    constexpr int MMAK = 32 / ESZ;       // K per instruction
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* sA = smem;                                  // [K/EPC][128][16 B]
    uint8_t* sB = smem + (size_t)(K / EPC) * 128 * 16;   // [K/EPC][N][16 B]
    for (int idx = tid; idx < 128 * K; idx += 128) {
        int r = idx / K, k = idx % K;
        size_t off = (size_t)(k / EPC) * (128 * 16) + (size_t)r * 16 + (size_t)(k % EPC) * ESZ;
        if (KIND == 0) *reinterpret_cast<__nv_bfloat16*>(sA + off) = __float2bfloat16(A[idx]);
        else *reinterpret_cast<float*>(sA + off) = A[idx];
    }
    for (int idx = tid; idx < K * N; idx += 128) {
        int k = idx / N, n = idx % N;
        size_t off = (size_t)(k / EPC) * ((size_t)N * 16) + (size_t)n * 16 + (size_t)(k % EPC) * ESZ;
        if (KIND == 0) *reinterpret_cast<__nv_bfloat16*>(sB + off) = __float2bfloat16(B[idx]);
        else *reinterpret_cast<float*>(sB + off) = B[idx];
    }
    int ncols = 32;
    while (ncols < N) ncols <<= 1;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_smem)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the async (MMA) proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_smem;
    const uint32_t idesc = make_idesc(128, N, KIND == 0 ? 1 : 2);
    long long t0 = 0, t1 = 0;
    bool ok = true;
    if (tid == 0) {
        t0 = clock64();
        for (int rep = 0; rep < reps; ++rep) {
            for (int ks = 0; ks < K / MMAK; ++ks) {
                const uint32_t a_addr = smem_u32(sA) + (uint32_t)(ks * 2) * (128 * 16);
                const uint32_t b_addr = smem_u32(sB) + (uint32_t)(ks * 2) * (uint32_t)(N * 16);
                const uint64_t ad = make_smem_desc(a_addr, 128 * 16, 128);
                const uint64_t bd = make_smem_desc(b_addr, (uint32_t)N * 16, 128);
                umma<KIND>(tmem, ad, bd, idesc, (rep > 0 || ks > 0) ? 1u : 0u);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    ok = mbar_wait_bounded(smem_u32(&bar), 0);
    if (tid == 0) { t1 = clock64(); *cycles = t1 - t0; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok) { if (tid == 0) *status = 2; }
    else {
        // warp w reads TMEM lanes 32w..32w+31 (= D rows), 8 columns at a time
        for (int c0 = 0; c0 < N; c0 += 8) {
            uint32_t v[8];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int row = warp * 32 + lane;
            for (int j = 0; j < 8; ++j) Dout[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
        }
        if (tid == 0) *status = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(ncols) : "memory");
}

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }
static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y; }

template <int KIND>
static void run_umma(int N, int K, int reps) {
    std::vector<float> A(128 * K), B((size_t)K * N), D(128 * (size_t)N, -1.f), ref(128 * (size_t)N);
    srand(1);
    for (auto& x : A) x = (rand() / (float)RAND_MAX - 0.5f);
    for (auto& x : B) x = (rand() / (float)RAND_MAX - 0.5f);
    for (int i = 0; i < 128; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0;
            for (int k = 0; k < K; ++k) {
                float a = A[i * K + k], b = B[(size_t)k * N + j];
                if (KIND == 0) { a = bf16_round(a); b = bf16_round(b); } else { a = tf32_trunc(a); b = tf32_trunc(b); }
                s += (double)a * b;
            }
            ref[(size_t)i * N + j] = (float)(s * reps);
        }
    float *dA, *dB, *dD; long long* dc; int* ds;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMalloc(&dc, 8)); CK(cudaMalloc(&ds, 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(ds, 0, 4));
    const int ESZ = KIND == 0 ? 2 : 4;
    size_t smem = (size_t)(K * ESZ / 16) * (128 + N) * 16 + 256;
    CK(cudaFuncSetAttribute(umma_probe<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_probe<KIND><<<1, 128, smem>>>(dA, dB, dD, N, K, reps, dc, ds);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("UMMA kind=%s N=%d K=%d: launch/exec error %s\n", KIND ? "tf32" : "bf16", N, K, cudaGetErrorString(e)); exit(2); }
    long long cyc; int st;
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (size_t i = 0; i < D.size(); ++i) { maxerr = fmax(maxerr, fabs((double)D[i] - ref[i])); maxref = fmax(maxref, fabs((double)ref[i])); }
    const int nmma = reps * (K / (KIND == 0 ? 16 : 8));
    printf("UMMA kind=%-4s M=128 N=%3d K=%3d reps=%4d status=%s max_abs_err=%.3e (max|ref|=%.3e) %s | %lld cycles total, %.1f cycles/MMA\n",
           KIND ? "tf32" : "bf16", N, K, reps, st == 1 ? "done" : st == 2 ? "TIMEOUT" : "none", maxerr, maxref,
           (st == 1 && maxerr <= 2e-3 * fmax(1.0, maxref)) ? "PASS" : "FAIL", cyc, (double)cyc / nmma);
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dc); cudaFree(ds);
}


// ---------------------------------------------------------------------------------- issue-path cost of the MMA loop
// mode 0: MMAs only (descriptors precomputed)   1: + tcgen05.commit every 3 MMAs   2: + try_wait on a completed mbarrier every 3 MMAs
__global__ void __launch_bounds__(128) umma_issue_probe(int N, int nmma, int mode, long long* cycles, int* status) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar, bar2, bar3;
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (7 * 4096 * 2 + 8 * 16384) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar2)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar3)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(&bar3)) : "memory");   // bar3: phase 0 complete
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_smem)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_smem;
    if (tid == 0) {
        const uint32_t idesc = make_idesc(128, N, 1);
        const uint64_t ad0 = make_smem_desc(smem_u32(smem), 128 * 16, 128);
        const uint64_t bd0 = make_smem_desc(smem_u32(smem) + 7 * 4096 * 2, (uint32_t)N * 16, 128);
        uint64_t ad = ad0, bd = bd0;
        const long long t0 = clock64();
        long long sink = 0;
        for (int i = 0; i < nmma; i += 3) {
            if (mode & 4) {   // fresh operand addresses every group (7 K-steps of A, 8 weight stages), like the engine
                const uint32_t ks = (uint32_t)(i / 3) % 7u, stg = (uint32_t)(i / 3) % 8u;
                ad = ad0 + ((ks * 4096u) >> 4);
                bd = bd0 + ((stg * 7168u) >> 4);
            }
            if (mode & 8) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (mode & 16) sink += clock64();
            if ((mode & 3) >= 2) {
                uint32_t ok;
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                             : "=r"(ok) : "r"(smem_u32(&bar3)), "r"(0u) : "memory");
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            umma<0>(tmem, ad, bd, idesc, 1u);
            umma<0>(tmem, ad, bd + 2, idesc, 1u);
            umma<0>(tmem, ad + 2, bd, idesc, 1u);
            if ((mode & 3) >= 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar2)) : "memory");
        }
        const long long t1 = clock64();
        if (sink == 12345) cycles[1] = sink;
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
        const bool ok = mbar_wait_bounded(smem_u32(&bar), 0);
        const long long t2 = clock64();
        cycles[0] = t1 - t0; cycles[1] = t2 - t0;
        *status = ok ? 1 : 2;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256) : "memory");
}

static void run_issue_probe(int N, int nmma, int mode) {
    long long* dc; int* ds;
    CK(cudaMalloc(&dc, 16)); CK(cudaMalloc(&ds, 4)); CK(cudaMemset(ds, 0, 4));
    size_t smem = 7 * 4096 * 2 + 8 * 16384 + 256;
    CK(cudaFuncSetAttribute(umma_issue_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_issue_probe<<<1, 128, smem>>>(N, nmma, mode, dc, ds);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("issue probe error %s\n", cudaGetErrorString(e)); exit(2); }
    long long c[2]; int st;
    CK(cudaMemcpy(c, dc, 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost));
    printf("issue probe N=%3d nmma=%4d mode=%d status=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA\n", N, nmma, mode, st,
           (double)c[0] / nmma, (double)c[1] / nmma);
    cudaFree(dc); cudaFree(ds);
}


// ---------------------------------------------------------------------------------- MMA rate vs shared-memory layout
// layout: 0 = no swizzle (LBO = rows*16, SBO = 128), 6 = SWIZZLE_32B (SBO 256), 4 = SWIZZLE_64B (SBO 512), 2 = SWIZZLE_128B (SBO 1024)
// Operands hold constants, so only the fetch pattern differs.  A K-step advances the start address by the layout's
// K-step stride; weight stages rotate over 8 slots; 3 MMAs per group (hi.hi, hi.lo, lo.hi pattern: A part +28 KB / B part +half stage).
__global__ void __launch_bounds__(128) umma_layout_probe(int N, int nmma, int layout, long long* cycles, int* status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (2 * 32768 + 8 * 16384) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_smem)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_smem;
    if (tid == 0) {
        const uint32_t idesc = make_idesc(128, N, 1);
        uint32_t a_lbo, a_sbo, b_lbo, b_sbo, a_kstride, a_part = 32768, b_part = 8192;
        if (layout == 0) { a_lbo = 2048; a_sbo = 128; b_lbo = (uint32_t)N * 16; b_sbo = 128; a_kstride = 4096; }
        else if (layout == 6) { a_lbo = 16; a_sbo = 256; b_lbo = 16; b_sbo = 256; a_kstride = 4096; }
        else if (layout == 4) { a_lbo = 16; a_sbo = 512; b_lbo = 16; b_sbo = 512; a_kstride = 32; }
        else { a_lbo = 16; a_sbo = 1024; b_lbo = 16; b_sbo = 1024; a_kstride = 32; }
        const uint64_t lt = (uint64_t)layout << 61;
        const uint64_t ad0 = make_smem_desc(smem_u32(smem), a_lbo, a_sbo) | lt;
        const uint64_t bd0 = make_smem_desc(smem_u32(smem) + 2 * 32768, b_lbo, b_sbo) | lt;
        const long long t0 = clock64();
        for (int i = 0; i < nmma; i += 3) {
            const uint32_t g = (uint32_t)(i / 3);
            const uint32_t ks = (layout == 4) ? (g % 2u) : (layout == 2) ? (g % 4u) : (g % 7u);
            const uint64_t ad = ad0 + ((ks * a_kstride) >> 4);
            const uint64_t bd = bd0 + (((g % 8u) * 16384u) >> 4);
            umma<0>(tmem, ad, bd, idesc, 1u);
            umma<0>(tmem, ad, bd + (b_part >> 4), idesc, 1u);
            umma<0>(tmem, ad + (a_part >> 4), bd, idesc, 1u);
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
        const bool ok = mbar_wait_bounded(smem_u32(&bar), 0);
        const long long t2 = clock64();
        cycles[0] = t1 - t0; cycles[1] = t2 - t0;
        *status = ok ? 1 : 2;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256) : "memory");
}
static void run_layout_probe(int N, int nmma, int layout) {
    long long* dc; int* ds;
    CK(cudaMalloc(&dc, 16)); CK(cudaMalloc(&ds, 4)); CK(cudaMemset(ds, 0, 4));
    size_t smem = 2 * 32768 + 8 * 16384 + 1024;
    CK(cudaFuncSetAttribute(umma_layout_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_layout_probe<<<1, 128, smem>>>(N, nmma, layout, dc, ds);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("layout probe error %s\n", cudaGetErrorString(e)); exit(2); }
    long long c[2]; int st;
    CK(cudaMemcpy(c, dc, 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost));
    printf("layout probe N=%3d layout=%d status=%d: %.1f cyc/MMA\n", N, layout, st, (double)c[1] / nmma);
    cudaFree(dc); cudaFree(ds);
}


// ---------------------------------------------------------------------------------- warp-converged issue with elect.sync
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred;
}
// mode 0: `if (threadIdx.x == 0)` single-lane loop (what the engine did)   1: whole warp runs the loop, elect.sync lane issues
// N small on purpose (issue-bound): reports cycles per MMA.  try: 1 = also a try_wait on a completed barrier + commit per group
__global__ void __launch_bounds__(128) umma_elect_probe(int N, int nmma, int mode, int with_sync, long long* cycles, int* status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar, bar2, bar3;
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (2 * 32768 + 8 * 16384) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar2)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar3)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(&bar3)) : "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_smem)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_smem;
    const uint32_t idesc = make_idesc(128, N, 1);
    const uint64_t ad0 = make_smem_desc(smem_u32(smem), 2048, 128);
    const uint64_t bd0 = make_smem_desc(smem_u32(smem) + 2 * 32768, (uint32_t)N * 16, 128);
    long long t0 = 0, t1 = 0;
    if (warp == 1) {
        if (mode == 0) {
            if ((tid & 31) == 0) {
                t0 = clock64();
                for (int i = 0; i < nmma; i += 3) {
                    const uint32_t g = (uint32_t)(i / 3);
                    if (with_sync) { uint32_t okk; asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(okk) : "r"(smem_u32(&bar3)), "r"(0u) : "memory"); }
                    const uint64_t ad = ad0 + (((g % 7u) * 4096u) >> 4), bd = bd0 + (((g % 8u) * 16384u) >> 4);
                    umma<0>(tmem, ad, bd, idesc, 1u);
                    umma<0>(tmem, ad, bd + (8192 >> 4), idesc, 1u);
                    umma<0>(tmem, ad + (32768 >> 4), bd, idesc, 1u);
                    if (with_sync) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar2)) : "memory");
                }
                t1 = clock64();
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
            }
        } else {
            t0 = clock64();
            for (int i = 0; i < nmma; i += 3) {
                const uint32_t g = (uint32_t)(i / 3);
                if (with_sync) { uint32_t okk; asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(okk) : "r"(smem_u32(&bar3)), "r"(0u) : "memory"); }
                const uint64_t ad = ad0 + (((g % 7u) * 4096u) >> 4), bd = bd0 + (((g % 8u) * 16384u) >> 4);
                if (elect_one()) {
                    umma<0>(tmem, ad, bd, idesc, 1u);
                    umma<0>(tmem, ad, bd + (8192 >> 4), idesc, 1u);
                    umma<0>(tmem, ad + (32768 >> 4), bd, idesc, 1u);
                    if (with_sync) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar2)) : "memory");
                }
                __syncwarp();
            }
            t1 = clock64();
            if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
        }
    }
    const bool ok = mbar_wait_bounded(smem_u32(&bar), 0);
    if (tid == 32) { cycles[0] = t1 - t0; cycles[1] = clock64() - t0; *status = ok ? 1 : 2; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256) : "memory");
}
static void run_elect_probe(int N, int nmma, int mode, int with_sync) {
    long long* dc; int* ds;
    CK(cudaMalloc(&dc, 16)); CK(cudaMalloc(&ds, 4)); CK(cudaMemset(ds, 0, 4));
    size_t smem = 2 * 32768 + 8 * 16384 + 16384 + 1024;
    CK(cudaFuncSetAttribute(umma_elect_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_elect_probe<<<1, 128, smem>>>(N, nmma, mode, with_sync, dc, ds);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("elect probe error %s\n", cudaGetErrorString(e)); exit(2); }
    long long c[2]; int st;
    CK(cudaMemcpy(c, dc, 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost));
    printf("elect probe N=%3d mode=%d sync=%d status=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA\n", N, mode, with_sync, st, (double)c[0] / nmma, (double)c[1] / nmma);
    cudaFree(dc); cudaFree(ds);
}


// ---------------------------------------------------------------------------------- replicate the engine's exact operand geometry
// variant bits: 1 = TMEM accumulator at column 112 (else 0); 2 = engine smem geometry (A lo part +28672, B stages 7168 apart at
// 172032, B lo part +3584) else the aligned probe geometry; 4 = pseudo-random operand data instead of constants
__global__ void __launch_bounds__(128) umma_geom_probe(int nmma, int variant, long long* cycles, int* status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int N = 112;
    for (int i = tid; i < 229376 / 4; i += 128) {
        uint32_t v = 0x3c003c00u;
        if (variant & 4) { uint32_t x = (uint32_t)i * 2654435761u; v = 0x3c003c00u ^ ((x >> 9) & 0x00ff00ffu) ^ ((x & 1u) << 15) ^ ((x & 2u) << 30); }
        reinterpret_cast<uint32_t*>(smem)[i] = v;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_smem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_smem + ((variant & 1) ? 112u : 0u);
    const uint32_t idesc = make_idesc(128, N, 1);
    const bool eng = (variant & 2) != 0;
    const uint32_t a_lo = eng ? 28672u : 32768u, b_base = eng ? 172032u : 65536u, b_stage = eng ? 7168u : 16384u, b_lo = eng ? 3584u : 8192u;
    const uint64_t ad0 = make_smem_desc(smem_u32(smem), 2048, 128);
    const uint64_t bd0 = make_smem_desc(smem_u32(smem) + b_base, (uint32_t)N * 16, 128);
    long long t0 = 0, t1 = 0;
    if (warp == 1) {
        t0 = clock64();
        for (int i = 0; i < nmma; i += 3) {
            const uint32_t g = (uint32_t)(i / 3);
            const uint64_t ad = ad0 + (((g % 7u) * 4096u) >> 4), bd = bd0 + (((g % 8u) * b_stage) >> 4);
            if (elect_one()) {
                umma<0>(tmem, ad, bd, idesc, 1u);
                umma<0>(tmem, ad, bd + (b_lo >> 4), idesc, 1u);
                umma<0>(tmem, ad + (a_lo >> 4), bd, idesc, 1u);
            }
            __syncwarp();
        }
        t1 = clock64();
        if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    const bool ok = mbar_wait_bounded(smem_u32(&bar), 0);
    if (tid == 32 && blockIdx.x == 0) { cycles[0] = t1 - t0; cycles[1] = clock64() - t0; *status = ok ? 1 : 2; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base_smem), "r"(512) : "memory");
}
static int g_geom_grid = 1;
static void run_geom_probe(int nmma, int variant) {
    long long* dc; int* ds;
    CK(cudaMalloc(&dc, 16)); CK(cudaMalloc(&ds, 4)); CK(cudaMemset(ds, 0, 4));
    size_t smem = 229376 + 1024;
    CK(cudaFuncSetAttribute(umma_geom_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_geom_probe<<<g_geom_grid, 128, smem>>>(nmma, variant, dc, ds);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("geom probe error %s\n", cudaGetErrorString(e)); exit(2); }
    long long c[2]; int st;
    CK(cudaMemcpy(c, dc, 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost));
    printf("grid=%3d ", g_geom_grid);
    printf("geom probe variant=%d (tmem col %s, %s geometry, %s data) status=%d: %.1f cyc/MMA\n", variant, (variant & 1) ? "112" : "0",
           (variant & 2) ? "engine" : "aligned", (variant & 4) ? "random" : "const", st, (double)c[1] / nmma);
    cudaFree(dc); cudaFree(ds);
}

// ---------------------------------------------------------------------------------- legacy mma.sync / FFMA throughput
__global__ void __launch_bounds__(256) mma_sync_bf16_tput(float* out, int iters) {
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 12345.678f) out[0] = s;
}
__global__ void __launch_bounds__(256) mma_sync_tf32_tput(float* out, int iters) {
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 12345.678f) out[0] = s;
}
__global__ void __launch_bounds__(256) ffma_tput(float* out, int iters) {
    float c[16];
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 0.001f + i;
    float a = 1.0001f, b = 0.9999f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fmaf(c[i], a, b);
    }
    float s = 0; for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 12345.678f) out[0] = s;
}

template <typename F>
static float time_kernel(F launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device: %s, %d SMs, cc %d.%d\n", p.name, p.multiProcessorCount, p.major, p.minor);
    // correctness of descriptor encodings (single MMA chain)
    run_umma<0>(112, 112, 1);
    run_umma<0>(208, 112, 1);
    run_umma<0>(256, 64, 1);
    run_umma<0>(64, 32, 1);
    run_umma<1>(112, 104, 1);
    run_umma<1>(208, 56, 1);
    // issue-rate / pipe-rate (same operands re-accumulated)
    run_umma<0>(112, 112, 64);
    run_umma<0>(208, 112, 64);
    run_umma<0>(256, 128, 64);
    run_umma<1>(112, 104, 64);
    run_umma<1>(256, 64, 64);
    for (int grid : {1, 2, 39, 74, 148}) { g_geom_grid = grid; run_geom_probe(768, 3); run_geom_probe(21, 3); }
    return 0;
    run_issue_probe(112, 21, 14); run_issue_probe(112, 63, 14);
    float* d; CK(cudaMalloc(&d, 4));
    const int sms = p.multiProcessorCount, iters = 4096;
    for (int bps = 1; bps <= 2; ++bps) {
        float ms = time_kernel([&] { mma_sync_bf16_tput<<<sms * bps, 256>>>(d, iters); });
        double flops = (double)sms * bps * 8 * iters * 8 * (2.0 * 16 * 8 * 16);
        printf("mma.sync bf16 m16n8k16: %d blk/SM x 8 warps: %.2f TFLOP/s (%.0f MAC/clk/SM @1.9GHz)\n", bps, flops / ms * 1e-9, flops / 2 / (ms * 1e-3) / sms / 1.9e9);
        ms = time_kernel([&] { mma_sync_tf32_tput<<<sms * bps, 256>>>(d, iters); });
        flops = (double)sms * bps * 8 * iters * 8 * (2.0 * 16 * 8 * 8);
        printf("mma.sync tf32 m16n8k8 : %d blk/SM x 8 warps: %.2f TFLOP/s (%.0f MAC/clk/SM @1.9GHz)\n", bps, flops / ms * 1e-9, flops / 2 / (ms * 1e-3) / sms / 1.9e9);
    }
    {
        float ms = time_kernel([&] { ffma_tput<<<sms * 4, 256>>>(d, iters); });
        double flops = (double)sms * 4 * 256 * iters * 16 * 2.0;
        printf("FFMA: %.2f TFLOP/s (%.0f FMA/clk/SM @1.9GHz)\n", flops / ms * 1e-9, flops / 2 / (ms * 1e-3) / sms / 1.9e9);
    }
    return 0;
}
