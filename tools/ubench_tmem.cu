// Microbenchmarks that decide whether the search filter can move to the tensor cores (DESIGN.md
// "tcgen05 filter"): (A) tcgen05.ld (TMEM -> registers) bandwidth per SM, (B) a kind::tf32
// tcgen05.mma with the no-swizzle K-major shared-memory layout: layout check against the CPU,
// accumulate numerics probes, and issue throughput.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_tmem tools/ubench_tmem.cu
// Run:   tools/ubench_tmem ldtm | mma
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc_512(uint32_t *slot)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free_512(uint32_t addr)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

#define R8(r, o) "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7])
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t *r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : R8(r, 0), R8(r, 8), R8(r, 16), R8(r, 24)
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t *r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : R8(r, 0) : "r"(taddr) : "memory");
}

// ------------------------------------------------------------------------------------------
// (A) TMEM read bandwidth: every warp streams x32 loads from its own 32-lane quadrant.
template <int LOADS_IN_FLIGHT>
__global__ void ldtm_kernel(unsigned long long *cycles, unsigned *sink, int iters)
{
    __shared__ uint32_t base_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc_512(&base_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = base_slot;
    const uint32_t lane_field = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t r[LOADS_IN_FLIGHT][32];
#pragma unroll
        for (int j = 0; j < LOADS_IN_FLIGHT; ++j) {
            const uint32_t col = (uint32_t)(((i * LOADS_IN_FLIGHT + j + (warp >> 2)) * 32) & 511);
            tmem_ld_x32(base + lane_field + col, r[j]);
        }
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < LOADS_IN_FLIGHT; ++j)
#pragma unroll
            for (int k = 0; k < 32; k += 2) acc ^= r[j][k] ^ r[j][k + 1];
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 0x12345678u) sink[0] = acc;
    __syncthreads();
    if (warp == 0) tmem_free_512(base);
}

static void run_ldtm()
{
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    unsigned long long *cyc;
    unsigned *sink;
    CK(cudaMalloc(&cyc, sms * sizeof(unsigned long long)));
    CK(cudaMalloc(&sink, 4));
    std::vector<unsigned long long> h(sms);
    const int iters = 4096;
    for (int inflight = 1; inflight <= 2; ++inflight)
        for (int warps = 4; warps <= 16; warps *= 2) {
            for (int rep = 0; rep < 2; ++rep) {
                if (inflight == 1) ldtm_kernel<1><<<sms, warps * 32>>>(cyc, sink, iters);
                else ldtm_kernel<2><<<sms, warps * 32>>>(cyc, sink, iters);
                CK(cudaDeviceSynchronize());
            }
            CK(cudaMemcpy(h.data(), cyc, sms * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            double mean = 0;
            for (int i = 0; i < sms; ++i) mean += (double)h[i];
            mean /= sms;
            const double bytes = (double)iters * inflight * warps * 32 * 32 * 4;
            printf("ldtm x32  in_flight %d  warps %2d : %.1f cycles/iter  %.1f B/clk/SM\n", inflight, warps, mean / iters,
                   bytes / mean);
        }
}

// ------------------------------------------------------------------------------------------
// (B) kind::tf32 MMA, M=128, N=256, K=8 per instruction, operands in the no-swizzle K-major
// canonical layout: element (row, k) at (k/4)*rows*16 + (row/8)*128 + (row%8)*16 + (k%4)*4 bytes.
__device__ __forceinline__ uint64_t smem_desc(unsigned addr, unsigned lbo_bytes, unsigned sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46; // descriptor version (Blackwell)
    return d;               // layout type 0 = no swizzle, base offset 0
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
// A operand read from TMEM (lane = row, one 32-bit column per k)
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
                 : "memory");
}

constexpr int MM = 128, NN = 256;
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(MM >> 4) << 24);

__host__ __device__ inline int canon_index(int row, int k, int rows) // in floats
{
    return (k / 4) * rows * 4 + (row / 8) * 32 + (row % 8) * 4 + (k % 4);
}

// ksteps MMAs of K=8 accumulate into one 128x256 tile; a/b are (rows, 8*ksteps) row-major in global
__global__ void mma_check_kernel(const float *a, const float *b, const float *c_init, float *d, int ksteps, int use_c)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t base_slot;
    __shared__ __align__(8) uint64_t bar;
    const int K = 8 * ksteps;
    float *sa = (float *)smem, *sb = sa + MM * K;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < MM * K; i += blockDim.x) sa[canon_index(i / K, i % K, MM)] = a[i];
    for (int i = threadIdx.x; i < NN * K; i += blockDim.x) sb[canon_index(i / K, i % K, NN)] = b[i];
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    if (warp == 0) tmem_alloc_512(&base_slot);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = base_slot;
    if (use_c) { // preload the accumulator through tcgen05.st
        const int row = (warp & 3) * 32 + lane;
        for (int c = 0; c < NN; c += 8) {
            uint32_t v[8];
            for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(c_init[row * NN + c + j]);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(
                             base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c),
                         "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                         : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < ksteps; ++s) {
            const uint64_t da = smem_desc(smem_u32(sa) + s * 2 * MM * 16, MM * 16, 128);
            const uint64_t db = smem_desc(smem_u32(sb) + s * 2 * NN * 16, NN * 16, 128);
            mma_tf32(base, da, db, IDESC, (s > 0 || use_c) ? 1u : 0u);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    {
        const int row = (warp & 3) * 32 + lane;
        for (int c = 0; c < NN; c += 8) {
            uint32_t v[8];
            tmem_ld_x8(base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c, v);
            tmem_wait_ld();
            for (int j = 0; j < 8; ++j) d[row * NN + c + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free_512(base);
}

__global__ void mma_rate_kernel(unsigned long long *cycles, int n_mma, int ksteps, int nn, int ts)
{
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(nn >> 3) << 17) | ((uint32_t)(MM >> 4) << 24);
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t base_slot;
    __shared__ __align__(8) uint64_t bar;
    const int K = 8 * ksteps;
    float *sa = (float *)smem, *sb = sa + MM * K;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (MM + NN) * K; i += blockDim.x) sa[i] = 1.0f;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    if (warp == 0) tmem_alloc_512(&base_slot);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = base_slot;
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        t0 = clock64();
        // descriptors precomputed and the 3 k-steps unrolled: the loop measures the tensor core, not this thread
        uint64_t da[3], db[3];
        for (int s = 0; s < 3; ++s) {
            da[s] = smem_desc(smem_u32(sa) + s * 2 * MM * 16, MM * 16, 128);
            db[s] = smem_desc(smem_u32(sb) + s * 2 * nn * 16, nn * 16, 128);
        }
        if (ts) {
            for (int i = 0; i < n_mma / 3; ++i) {
                const uint32_t d = base + (uint32_t)(i & 1) * 256u;
                mma_tf32_ts(d, base + 448u, db[0], idesc, 0u);
                mma_tf32_ts(d, base + 456u, db[1], idesc, 1u);
                mma_tf32_ts(d, base + 464u, db[2], idesc, 1u);
            }
        } else {
            for (int i = 0; i < n_mma / 3; ++i) {
                const uint32_t d = base + (uint32_t)(i & 1) * 256u;
                mma_tf32(d, da[0], db[0], idesc, 0u);
                mma_tf32(d, da[1], db[1], idesc, 1u);
                mma_tf32(d, da[2], db[2], idesc, 1u);
            }
        }
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        t1 = clock64();
        cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free_512(base);
}

static float tf32_trunc(float x) { uint32_t b; memcpy(&b, &x, 4); b &= 0xffffe000u; memcpy(&x, &b, 4); return x; }

static void run_check(const char *name, const std::vector<float> &a, const std::vector<float> &b, const std::vector<float> *c,
                      int ksteps, bool print_rows)
{
    const int K = 8 * ksteps;
    float *da, *db, *dc = nullptr, *dd;
    CK(cudaMalloc(&da, a.size() * 4)); CK(cudaMalloc(&db, b.size() * 4)); CK(cudaMalloc(&dd, MM * NN * 4));
    CK(cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
    if (c) { CK(cudaMalloc(&dc, MM * NN * 4)); CK(cudaMemcpy(dc, c->data(), MM * NN * 4, cudaMemcpyHostToDevice)); }
    const int smem = (MM + NN) * K * 4;
    CK(cudaFuncSetAttribute(mma_check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mma_check_kernel<<<1, 128, smem>>>(da, db, dc, dd, ksteps, c ? 1 : 0);
    CK(cudaDeviceSynchronize());
    std::vector<float> d(MM * NN);
    CK(cudaMemcpy(d.data(), dd, MM * NN * 4, cudaMemcpyDeviceToHost));
    double worst = 0, worst_rel_abs = 0; long mism = 0;
    for (int r = 0; r < MM; ++r)
        for (int n = 0; n < NN; ++n) {
            double s = c ? (double)(*c)[r * NN + n] : 0.0, sabs = fabs(s);
            for (int k = 0; k < K; ++k) { double p = (double)tf32_trunc(a[r * K + k]) * (double)tf32_trunc(b[n * K + k]); s += p; sabs += fabs(p); }
            const double err = fabs((double)d[r * NN + n] - s);
            if (err > 0) ++mism;
            if (err > worst) worst = err;
            if (sabs > 0 && err / sabs > worst_rel_abs) worst_rel_abs = err / sabs;
        }
    printf("%-28s ksteps %d: inexact %ld / %d  max|err| %.3g  max err/sum|terms| %.3g = %.2f u (u=2^-24)\n", name, ksteps, mism,
           MM * NN, worst, worst_rel_abs, worst_rel_abs * 16777216.0);
    if (print_rows)
        for (int r = 0; r < 8; ++r) printf("   row %d: d[0]=%.10g d[1]=%.10g d[255]=%.10g\n", r, d[r * NN], d[r * NN + 1], d[r * NN + 255]);
    cudaFree(da); cudaFree(db); cudaFree(dd); if (dc) cudaFree(dc);
}

static void run_mma()
{
    // 1. layout check: small integers, every (row, n) distinct
    {
        const int ks = 3, K = 24;
        std::vector<float> a(MM * K), b(NN * K);
        for (int r = 0; r < MM; ++r) for (int k = 0; k < K; ++k) a[r * K + k] = (float)((r * 7 + k * 3) % 11 - 5);
        for (int n = 0; n < NN; ++n) for (int k = 0; k < K; ++k) b[n * K + k] = (float)((n * 5 + k * 13) % 17 - 8);
        run_check("layout (small ints)", a, b, nullptr, ks, false);
    }
    // 2. numerics probes, one per row: a row times a column of ones
    {
        const int ks = 1, K = 8;
        std::vector<float> a(MM * K, 0.f), b(NN * K, 1.f);
        auto set = [&](int r, std::initializer_list<float> v) { int k = 0; for (float x : v) a[r * K + k++] = x; };
        set(0, {16777216.f, 1, 1, 1, 1, 1, 1, 1});            // 2^24 + 7 ones: exact 16777223 -> RN 16777224, trunc-each 16777216
        set(1, {8388608.f, 0.5f, 0.5f, 0.5f, 0.5f, 0, 0, 0});  // 2^23 + 4 halves: exact 8388610
        set(2, {16777216.f, 1, 0, 0, 0, 0, 0, 0});             // tie: RN-even 16777216, trunc 16777216
        set(3, {16777216.f, 3, 0, 0, 0, 0, 0, 0});             // exact 16777219 -> RN 16777220, trunc 16777218
        set(4, {16777216.f, -1, 0, 0, 0, 0, 0, 0});            // exact 16777215 (representable)
        set(5, {33554432.f, 1, 1, 1, 1, 0, 0, 0});             // 2^25 + 4: exact representable only if summed before rounding
        set(6, {1.f, 5.9604645e-8f, 5.9604645e-8f, 0, 0, 0, 0, 0}); // 1 + 2*2^-24 = 1 + 2^-23 exactly
        set(7, {16777216.f, 1.5f, 1.5f, 0, 0, 0, 0, 0});       // exact 16777219
        run_check("probes (A row x ones)", a, b, nullptr, ks, true);
        std::vector<float> c(MM * NN, 16777216.f);
        std::vector<float> a2(MM * K, 0.f);
        for (int r = 0; r < MM; ++r) for (int k = 0; k < (r % 8) + 1; ++k) a2[r * K + k] = 1.f; // C=2^24 plus 1..8 ones
        run_check("probes (C=2^24 + r ones)", a2, b, &c, ks, true);
    }
    // 3. random tf32-exact operands, wide exponent range
    {
        srand(7);
        for (int ks = 1; ks <= 3; ks += 2) {
            const int K = 8 * ks;
            std::vector<float> a(MM * K), b(NN * K);
            auto rnd = []() { double m = (rand() / (double)RAND_MAX) * 2 - 1; int e = rand() % 24 - 12; return tf32_trunc((float)ldexp(m, e)); };
            for (auto &x : a) x = rnd();
            for (auto &x : b) x = rnd();
            run_check("random tf32 operands", a, b, nullptr, ks, false);
        }
    }
    // 4. fp32 operands that are NOT tf32-exact: does the unit truncate or round the low 13 bits?
    {
        const int ks = 1, K = 8;
        std::vector<float> a(MM * K, 0.f), b(NN * K, 0.f);
        for (int r = 0; r < MM; ++r) a[r * K] = 1.0f + (float)(r + 1) * 1.1920929e-7f * 512.f; // low bits set
        for (int n = 0; n < NN; ++n) b[n * K] = 1.0f;
        run_check("non-tf32 operands (trunc?)", a, b, nullptr, ks, true);
    }
    // 5. issue throughput
    {
        int sms = 0;
        CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
        unsigned long long *cyc;
        CK(cudaMalloc(&cyc, sms * 8));
        std::vector<unsigned long long> h(sms);
        const int ks = 3, smem = (MM + NN) * 8 * ks * 4;
        CK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        const int n_mma = 3000;
        for (int ts = 0; ts < 2; ++ts)
        for (int nn = 64; nn <= (ts ? 192 : 256); nn += 32) {
            for (int rep = 0; rep < 2; ++rep) { mma_rate_kernel<<<sms, 128, smem>>>(cyc, n_mma, ks, nn, ts); CK(cudaDeviceSynchronize()); }
            CK(cudaMemcpy(h.data(), cyc, sms * 8, cudaMemcpyDeviceToHost));
            double mean = 0; for (int i = 0; i < sms; ++i) mean += (double)h[i]; mean /= sms;
            printf("mma tf32 %s 128x%dx8: %d back-to-back on %d SMs: %.1f cycles each  (%.0f flop/clk/SM)\n", ts ? "A in TMEM" : "A in smem", nn, n_mma, sms, mean / n_mma,
                   2.0 * MM * nn * 8 * n_mma / mean);
        }
    }
}


// ------------------------------------------------------------------------------------------
// (C) do tcgen05.ld and tcgen05.mma share a TMEM port?  `ld_warps` warps stream x32 loads from
// columns 384.. while (optionally) one more warp issues 128x192x8 MMAs into columns 0..383.
__global__ void contention_kernel(unsigned long long *cycles, unsigned *sink, int iters, int n_mma, int ld_warps, int do_mma)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t base_slot;
    __shared__ __align__(8) uint64_t bar;
    float *sa = (float *)smem, *sb = sa + MM * 24;
    const int warp = threadIdx.x >> 5;
    const int nn = 192;
    for (int i = threadIdx.x; i < (MM + nn) * 24; i += blockDim.x) sa[i] = 1.0f;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    if (warp == 0) tmem_alloc_512(&base_slot);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = base_slot;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(nn >> 3) << 17) | ((uint32_t)(MM >> 4) << 24);
    uint32_t acc = 0;
    if (warp < ld_warps) {
        const uint32_t lane_field = (uint32_t)((warp & 3) * 32) << 16;
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            uint32_t r[32];
            tmem_ld_x32(base + lane_field + 384u + (uint32_t)(((i + (warp >> 2)) * 32) & 127), r);
            tmem_wait_ld();
#pragma unroll
            for (int k = 0; k < 32; k += 2) acc ^= r[k] ^ r[k + 1];
        }
        const long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * 32 + warp] = (unsigned long long)(t1 - t0);
    } else if (warp == ld_warps && do_mma) {
        if ((threadIdx.x & 31) == 0) {
            uint64_t da[3], db[3];
            for (int s = 0; s < 3; ++s) {
                da[s] = smem_desc(smem_u32(sa) + s * 2 * MM * 16, MM * 16, 128);
                db[s] = smem_desc(smem_u32(sb) + s * 2 * nn * 16, nn * 16, 128);
            }
            const long long t0 = clock64();
            for (int i = 0; i < n_mma / 3; ++i) {
                const uint32_t d = base + (uint32_t)(i & 1) * 192u;
                mma_tf32(d, da[0], db[0], idesc, 0u);
                mma_tf32(d, da[1], db[1], idesc, 1u);
                mma_tf32(d, da[2], db[2], idesc, 1u);
            }
            mma_commit(&bar);
            mbar_wait(&bar, 0);
            const long long t1 = clock64();
            cycles[blockIdx.x * 32 + 31] = (unsigned long long)(t1 - t0);
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free_512(base);
}

static void run_contention()
{
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    unsigned long long *cyc;
    unsigned *sink;
    CK(cudaMalloc(&cyc, sms * 32 * 8));
    CK(cudaMalloc(&sink, 4));
    std::vector<unsigned long long> h(sms * 32);
    const int smem = (MM + 192) * 24 * 4;
    CK(cudaFuncSetAttribute(contention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int ld_warps = 4; ld_warps <= 16; ld_warps *= 2)
        for (int do_mma = 0; do_mma < 2; ++do_mma) {
            const int iters = 3000, n_mma = 9000; // ~290 k cycles of MMAs, ~290 k cycles of loads
            CK(cudaMemset(cyc, 0, sms * 32 * 8));
            for (int rep = 0; rep < 2; ++rep) {
                contention_kernel<<<sms, (ld_warps + 1) * 32, smem>>>(cyc, sink, iters, n_mma, ld_warps, do_mma);
                CK(cudaDeviceSynchronize());
            }
            CK(cudaMemcpy(h.data(), cyc, sms * 32 * 8, cudaMemcpyDeviceToHost));
            double ld = 0, mma = 0;
            for (int b = 0; b < sms; ++b) { for (int w = 0; w < ld_warps; ++w) ld += (double)h[b * 32 + w]; mma += (double)h[b * 32 + 31]; }
            ld /= (double)sms * ld_warps; mma /= sms;
            printf("contention: %2d ld warps, mma %d : %.1f cycles per x32 load per warp (%.0f B/clk/SM)", ld_warps, do_mma, ld / iters,
                   (double)iters * ld_warps * 4096.0 / ld);
            if (do_mma) printf("   %.1f cycles per 128x192x8 MMA", mma / n_mma);
            printf("\n");
        }
}

int main(int argc, char **argv)
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    printf("%s, %d SMs, %d MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1000);
    if (argc > 1 && !strcmp(argv[1], "mma")) run_mma();
    else if (argc > 1 && !strcmp(argv[1], "contention")) run_contention();
    else run_ldtm();
    return 0;
}
