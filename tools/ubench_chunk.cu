#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
// per-chunk epilogue arithmetic (square form) on register data; vary warps per SMSP and ILP shape
__device__ __forceinline__ unsigned or_and(unsigned a, unsigned b, unsigned c){ unsigned d; asm("lop3.b32 %0, %1, %2, %3, 0xA8;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
template <int MODE>
__global__ void chunk_kernel(const uint32_t *in, unsigned *out, unsigned long long *cycles, int reps)
{
    uint32_t r[48];
    for (int k = 0; k < 48; ++k) r[k] = in[(threadIdx.x & 31) * 48 + k];
    const float kf = __uint_as_float(in[0]);
    const float2 kk = make_float2(kf, kf);
    unsigned acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < reps; ++it) {
#pragma unroll
        for (int q = 0; q < 8; ++q) { r[6 * q] += 8u; } // new data every iteration (8 IADD)
        float2 du[8], dv[8];
        unsigned none[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float2 den = make_float2(__uint_as_float(r[6 * q + 0]), __uint_as_float(r[6 * q + 1]));
            const float2 au = make_float2(__uint_as_float(r[6 * q + 2]), __uint_as_float(r[6 * q + 3]));
            const float2 av = make_float2(__uint_as_float(r[6 * q + 4]), __uint_as_float(r[6 * q + 5]));
            const float2 d2 = __ffma2_rn(den, den, kk);
            du[q] = __ffma2_rn(make_float2(-au.x, -au.y), au, d2);
            dv[q] = __ffma2_rn(make_float2(-av.x, -av.y), av, d2);
            if (MODE != 2) {
                none[(2 * q) & 3] = or_and(__float_as_uint(du[q].x), __float_as_uint(dv[q].x), none[(2 * q) & 3]);
                none[(2 * q + 1) & 3] = or_and(__float_as_uint(du[q].y), __float_as_uint(dv[q].y), none[(2 * q + 1) & 3]);
            } else { // FP-only accumulate: min via fminf (ALU) replaced by nothing: just xor to keep values live
                none[q & 3] ^= __float_as_uint(du[q].x) ^ __float_as_uint(dv[q].y);
            }
        }
        const bool some = (int)(none[0] & none[1] & none[2] & none[3]) >= 0;
        if (MODE == 0) { if (__any_sync(0xffffffffu, some)) { if (some) acc += 1; __syncwarp(); } }
        else { if (some) acc += 1; }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}
int main()
{
    int sms = 148;
    uint32_t *in; unsigned *out; unsigned long long *cyc;
    cudaMalloc(&in, 32 * 48 * 4); cudaMalloc(&out, sms * 1024 * 4); cudaMalloc(&cyc, sms * 8);
    float *h = (float *)malloc(32 * 48 * 4);
    for (int i = 0; i < 32 * 48; ++i) h[i] = ((i % 6) < 2) ? 1.0f + (i % 7) : 3.0f + (float)(i % 13);   // |a| > |den|: never a candidate
    h[0] = 1e-7f;
    cudaMemcpy(in, h, 32 * 48 * 4, cudaMemcpyHostToDevice);
    unsigned long long hc[148];
    const int reps = 4000;
    for (int mode = 0; mode < 3; ++mode)
        for (int wps = 1; wps <= 4; ++wps) {
            int warps = wps * 4;
            for (int k = 0; k < 2; ++k) {
                if (mode == 0) chunk_kernel<0><<<sms, warps * 32>>>(in, out, cyc, reps);
                else if (mode == 1) chunk_kernel<1><<<sms, warps * 32>>>(in, out, cyc, reps);
                else chunk_kernel<2><<<sms, warps * 32>>>(in, out, cyc, reps);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(hc, cyc, sms * 8, cudaMemcpyDeviceToHost);
            double m = 0; for (int i = 0; i < sms; ++i) m += hc[i]; m /= sms;
            printf("mode %d (%s)  %d warps/SMSP: %.1f cycles per chunk per warp, %.1f cycles per chunk per SMSP\n", mode,
                   mode == 0 ? "vote-uniform branch" : mode == 1 ? "divergent branch" : "no LOP3 accumulate", wps, m / reps, m / reps / wps);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
