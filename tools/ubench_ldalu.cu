// Does a tcgen05.ld in flight slow down the CUDA-core work of its own warp / its SM sub-partition?
// Per warp and iteration: mode 0 = load 2 x 48 TMEM columns, wait after each; mode 1 = the linear-form chunk test
// (16 FMNMX + 16 FADD + 8 FMNMX3 + vote) on register data, twice; mode 2 = both, software pipelined (the load of the
// other buffer is issued, then this chunk is tested, then tcgen05.wait::ld); mode 3 = load, wait, test in series.
// If load and test overlapped perfectly inside a warp, mode 2 would cost max(mode 0, mode 1).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_ldalu tools/ubench_ldalu.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float fmax_nan(float a, float b) { float d; asm("max.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float fmax3_nan(float a, float b, float c) { float d; asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
#define R8(r, o) "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7])
__device__ __forceinline__ void ld48(uint32_t taddr, uint32_t *r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : R8(r, 0), R8(r, 8), R8(r, 16), R8(r, 24) : "r"(taddr) : "memory");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : R8(r, 32), R8(r, 40) : "r"(taddr + 32u) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ unsigned test48(const uint32_t (&r)[48], float kk)
{
    float top[2] = {-3.0e38f, -3.0e38f};
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float fa = fabsf(__uint_as_float(r[6 * q + 0])) - fmax_nan(fabsf(__uint_as_float(r[6 * q + 2])), fabsf(__uint_as_float(r[6 * q + 4])));
        const float fb = fabsf(__uint_as_float(r[6 * q + 1])) - fmax_nan(fabsf(__uint_as_float(r[6 * q + 3])), fabsf(__uint_as_float(r[6 * q + 5])));
        top[q & 1] = fmax3_nan(top[q & 1], fa, fb);
    }
    const bool some = !(fmax_nan(top[0], top[1]) + kk < 0.0f);
    unsigned acc = 0;
    if (__any_sync(0xffffffffu, some)) {
        if (some) acc = 1;
        __syncwarp();
    }
    return acc;
}

template <int MODE>
__global__ void k(unsigned *out, unsigned long long *cycles, int reps, float kk)
{
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(((warp >> 2) & 3) * 96);
    uint32_t rA[48], rB[48];
#pragma unroll
    for (int i = 0; i < 48; ++i) rA[i] = rB[i] = (i % 6) < 2 ? 0u : 0x3f800000u;
    unsigned acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < reps; ++it) {
        if (MODE == 0) {
            ld48(base, rA);
            wait_ld();
#pragma unroll
            for (int i = 0; i < 48; ++i) asm volatile("" ::"r"(rA[i]));
            ld48(base + 48u, rB);
            wait_ld();
#pragma unroll
            for (int i = 0; i < 48; ++i) asm volatile("" ::"r"(rB[i]));
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 48; ++i) {
                asm volatile("" : "+r"(rA[i]));
                asm volatile("" : "+r"(rB[i]));
            }
            acc += test48(rA, kk);
            acc += test48(rB, kk);
        } else if (MODE == 2) {
            ld48(base + 48u, rB);
#pragma unroll
            for (int i = 0; i < 48; ++i) asm volatile("" : "+r"(rA[i]));
            acc += test48(rA, kk);
            wait_ld();
            ld48(base, rA);
#pragma unroll
            for (int i = 0; i < 48; ++i) asm volatile("" : "+r"(rB[i]));
            acc += test48(rB, kk);
            wait_ld();
        } else if (MODE == 3) {
            ld48(base, rA);
            wait_ld();
            acc += test48(rA, kk);
            ld48(base + 48u, rB);
            wait_ld();
            acc += test48(rB, kk);
        } else { // MODE 4: the order of search_tc_kernel: wait, test A, load A', test B, load B'
            wait_ld();
#pragma unroll
            for (int i = 0; i < 48; ++i) asm volatile("" : "+r"(rA[i]));
            acc += test48(rA, kk);
            ld48(base, rA);
#pragma unroll
            for (int i = 0; i < 48; ++i) asm volatile("" : "+r"(rB[i]));
            acc += test48(rB, kk);
            ld48(base + 48u, rB);
        }
    }
    const long long t1 = clock64();
    wait_ld();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + rA[5] + rB[7];
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}


// ---- an MMA stream beside the epilogue (mode "mma"): does the triple {TMEM load, chunk test, MMA} still overlap?
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(unsigned addr, unsigned lbo_bytes, unsigned sbo_bytes)
{
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// EPI_MODE: 2 = one load in flight during every test, 4 = the kernel's order; epilogue warps 0..7 (two per sub-partition)
// read the two 192-column accumulators alternately like search_tc_kernel; warp 8 (do_mma) issues 3 MMAs (128 x 192 x 8)
// per step into the same accumulators, free-running.
// BLOAT = 1: every chunk test is followed by a never-taken copy of the kernel's candidate branch (16-bit mask from the
// registers + append to a shared-memory queue) and the step body is unrolled four times, like search_tc_kernel<4>:
// the instruction footprint of the loop grows from ~0.2 K to ~1.4 K instructions, the executed instructions do not change.
__device__ __forceinline__ unsigned rare_block(const uint32_t (&r)[48], float kk, int *q_count, int4 *queue, int tag)
{
    unsigned mask = 0;
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        const int c = 6 * (t >> 1) + (t & 1);
        const float f = fabsf(__uint_as_float(r[c])) - fmax_nan(fabsf(__uint_as_float(r[c + 2])), fabsf(__uint_as_float(r[c + 4])));
        if (!(f + kk < 0.0f)) mask |= 1u << t;
    }
    const int cnt = __popc(mask);
    int pos = atomicAdd(q_count, cnt);
    unsigned m = mask;
    while (m) {
        const int b = __ffs((int)m) - 1;
        m &= m - 1;
        if (pos < 256) queue[pos++] = make_int4(tag, b, 0, 0);
    }
    return (unsigned)cnt;
}
template <int BLOAT>
__device__ __forceinline__ unsigned test48b(const uint32_t (&r)[48], float kk, int *q_count, int4 *queue, int tag, int never)
{
    float top[2] = {-3.0e38f, -3.0e38f};
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float fa = fabsf(__uint_as_float(r[6 * q + 0])) - fmax_nan(fabsf(__uint_as_float(r[6 * q + 2])), fabsf(__uint_as_float(r[6 * q + 4])));
        const float fb = fabsf(__uint_as_float(r[6 * q + 1])) - fmax_nan(fabsf(__uint_as_float(r[6 * q + 3])), fabsf(__uint_as_float(r[6 * q + 5])));
        top[q & 1] = fmax3_nan(top[q & 1], fa, fb);
    }
    const bool some = !(fmax_nan(top[0], top[1]) + kk < 0.0f) && never != 0; // `never` is 0 at run time: the branch is never taken
    unsigned acc = 0;
    if (__any_sync(0xffffffffu, some)) {
        if (some) acc = BLOAT ? rare_block(r, kk, q_count, queue, tag) : 1u;
        __syncwarp();
    }
    return acc;
}

template <int EPI_MODE, int BLOAT = 0>
__global__ void triple_kernel(unsigned *out, unsigned long long *cycles, int reps, float kk, int do_mma, int never)
{
    __shared__ int4 queue_mem[8 * 256];
    __shared__ int q_counts[8 * 4];
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bar;
    float *sa = (float *)smem, *sb = sa + 128 * 24;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (128 + 192) * 24; i += blockDim.x) sa[i] = 1.0f + (float)(i % 7);
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = slot;
    unsigned acc = 0;
    if (warp < 8) {
        const uint32_t base = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(((warp >> 2) & 1) * 96);
        uint32_t rA[48], rB[48];
#pragma unroll
        for (int i = 0; i < 48; ++i) rA[i] = rB[i] = (i % 6) < 2 ? 0u : 0x3f800000u;
        int *q_count = q_counts + (warp & 7) * 4;
        int4 *queue = queue_mem + (warp & 7) * 256;
        if ((threadIdx.x & 31) == 0) *q_count = 0;
        __syncwarp();
        const long long t0 = clock64();
#pragma unroll 1
        for (int it4 = 0; it4 < reps; it4 += (BLOAT ? 4 : 1)) {
#pragma unroll
          for (int u = 0; u < (BLOAT ? 4 : 1); ++u) {
            const int it = it4 + u;
            const uint32_t a0 = base + (uint32_t)(((BLOAT ? u : it) & 1) * 256);
            if (EPI_MODE == 2) {
                ld48(a0 + 48u, rB);
#pragma unroll
                for (int i = 0; i < 48; ++i) asm volatile("" : "+r"(rA[i]));
                acc += test48b<BLOAT>(rA, kk, q_count, queue, it, never);
                wait_ld();
                ld48(a0 ^ 256u, rA);
#pragma unroll
                for (int i = 0; i < 48; ++i) asm volatile("" : "+r"(rB[i]));
                acc += test48b<BLOAT>(rB, kk, q_count, queue, it, never);
                wait_ld();
            } else {
                wait_ld();
#pragma unroll
                for (int i = 0; i < 48; ++i) asm volatile("" : "+r"(rA[i]));
                acc += test48b<BLOAT>(rA, kk, q_count, queue, it, never);
                ld48(a0, rA);
#pragma unroll
                for (int i = 0; i < 48; ++i) asm volatile("" : "+r"(rB[i]));
                acc += test48b<BLOAT>(rB, kk, q_count, queue, it, never);
                ld48(a0 + 48u, rB);
            }
          }
        }
        const long long t1 = clock64();
        wait_ld();
        out[blockIdx.x * blockDim.x + threadIdx.x] = acc + rA[5] + rB[7];
        if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * 16 + warp] = (unsigned long long)(t1 - t0);
    } else if (warp == 8 && do_mma) {
        if ((threadIdx.x & 31) == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(192 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            uint64_t da[3], db[3];
            for (int s2 = 0; s2 < 3; ++s2) {
                da[s2] = smem_desc(smem_u32(sa) + s2 * 2 * 128 * 16, 128 * 16, 128);
                db[s2] = smem_desc(smem_u32(sb) + s2 * 2 * 192 * 16, 192 * 16, 128);
            }
            const long long t0 = clock64();
            for (int i = 0; i < reps; ++i) {
                const uint32_t d = tbase + (uint32_t)(i & 1) * 256u;
                mma_tf32(d, da[0], db[0], idesc, 0u);
                mma_tf32(d, da[1], db[1], idesc, 1u);
                mma_tf32(d, da[2], db[2], idesc, 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            unsigned done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
            const long long t1 = clock64();
            cycles[blockIdx.x * 16 + 8] = (unsigned long long)(t1 - t0);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

int main()
{
    const int sms = 148, reps = 2000;
    unsigned *out;
    unsigned long long *cyc;
    cudaMalloc(&out, sms * 1024 * 4);
    cudaMalloc(&cyc, sms * 8);
    unsigned long long hc[148];
    const char *names[5] = {"load only (2 x 48 columns, wait after each)", "test only (2 chunks)", "pipelined load + test", "serial load, wait, test", "kernel order: wait, test A, ld A, test B, ld B"};
    for (int mode = 0; mode < 5; ++mode)
        for (int warps = 4; warps <= 16; warps *= 2) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<sms, warps * 32>>>(out, cyc, reps, -2.0f);
                else if (mode == 1) k<1><<<sms, warps * 32>>>(out, cyc, reps, -2.0f);
                else if (mode == 2) k<2><<<sms, warps * 32>>>(out, cyc, reps, -2.0f);
                else if (mode == 3) k<3><<<sms, warps * 32>>>(out, cyc, reps, -2.0f);
                else k<4><<<sms, warps * 32>>>(out, cyc, reps, -2.0f);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(hc, cyc, sms * 8, cudaMemcpyDeviceToHost);
            double m = 0;
            for (int i = 0; i < sms; ++i) m += hc[i];
            m /= sms;
            printf("mode %d %-46s %2d warps/CTA (%d per SMSP): %7.1f cycles per 2 chunks per warp = %6.1f per chunk per SMSP\n", mode,
                   names[mode], warps, warps / 4, m / reps, m / reps / 2 / (warps / 4));
        }
    {
        unsigned long long *cyc2;
        cudaMalloc(&cyc2, sms * 16 * 8);
        unsigned long long h2[148 * 16];
        const int smem = (128 + 192) * 24 * 4;
        cudaFuncSetAttribute(triple_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(triple_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(triple_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int mode = 2; mode <= 5; ++mode) {
          if (mode == 3) continue;
            for (int do_mma = 0; do_mma < 2; ++do_mma) {
                cudaMemset(cyc2, 0, sms * 16 * 8);
                for (int rep = 0; rep < 2; ++rep) {
                    if (mode == 2) triple_kernel<2><<<sms, 9 * 32, smem>>>(out, cyc2, reps, -2.0f, do_mma, 0);
                    else if (mode == 4) triple_kernel<4><<<sms, 9 * 32, smem>>>(out, cyc2, reps, -2.0f, do_mma, 0);
                    else triple_kernel<4, 1><<<sms, 9 * 32, smem>>>(out, cyc2, reps, -2.0f, do_mma, 0);
                    cudaDeviceSynchronize();
                }
                cudaMemcpy(h2, cyc2, sms * 16 * 8, cudaMemcpyDeviceToHost);
                double e = 0, m = 0;
                for (int b = 0; b < sms; ++b) { for (int w = 0; w < 8; ++w) e += (double)h2[b * 16 + w]; m += (double)h2[b * 16 + 8]; }
                e /= (double)sms * 8; m /= sms;
                printf("triple: 8 epilogue warps, order %s, MMA stream %s: %.1f cycles per step per epilogue warp", mode == 2 ? "one-load-in-flight" : mode == 4 ? "kernel" : "kernel + inline candidate branch, unrolled x4", do_mma ? "ON " : "off", e / reps);
                if (do_mma) printf(", %.1f cycles per 3 MMAs (128 x 192 x 8)", m / reps);
                printf("\n");
            }
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
