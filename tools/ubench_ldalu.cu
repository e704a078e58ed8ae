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
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
