// ubench_fp32.cu - FP32 issue-rate microbenchmarks for B200 (sm_100a).
// Decides the formulation of the search filter: scalar FFMA vs packed FFMA2 (fma.rn.f32x2),
// co-issue with ALU-pipe ops (LOP3) and with broadcast LDS.128.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_fp32 ubench_fp32.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define ITERS 2048
#define NACC 8

__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

__global__ void k_ffma(float *out, float x, float y, long long *cyc)
{
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = fmaf(acc[i], x, y);
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// 3 distinct register sources per FFMA (a*b+c with a,b varying registers)
__global__ void k_ffma3(float *out, const float *in, long long *cyc)
{
    float acc[NACC], xs[NACC], ys[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i] = threadIdx.x * 1e-3f + i; xs[i] = in[i]; ys[i] = in[i + NACC]; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = fmaf(xs[(i + u) % NACC], ys[i], acc[i]);
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void k_ffma2(float *out, const float *in, long long *cyc)
{
    uint64_t acc[NACC], xs[NACC], ys[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        float2 a = make_float2(threadIdx.x * 1e-3f + i, 1.0f + i);
        float2 b = make_float2(in[i], in[i + 1]);
        float2 c = make_float2(in[i + NACC], in[i + NACC + 1]);
        acc[i] = *reinterpret_cast<uint64_t *>(&a);
        xs[i] = *reinterpret_cast<uint64_t *>(&b);
        ys[i] = *reinterpret_cast<uint64_t *>(&c);
    }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = ffma2(xs[(i + u) % NACC], ys[i], acc[i]);
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) { float2 a = *reinterpret_cast<float2 *>(&acc[i]); s += a.x + a.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// FFMA : LOP3 = 4 : 1 (scalar) - does the ALU pipe co-issue for free?
__global__ void k_ffma_lop(float *out, const float *in, long long *cyc)
{
    float acc[NACC], xs[NACC], ys[NACC];
    unsigned bits[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i] = threadIdx.x * 1e-3f + i; xs[i] = in[i]; ys[i] = in[i + NACC]; bits[i] = threadIdx.x + i; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = fmaf(xs[(i + u) % NACC], ys[i], acc[i]);
#pragma unroll
            for (int i = 0; i < NACC; i += 4) bits[i] = (bits[i] ^ __float_as_uint(acc[i])) | (bits[(i + 4) % NACC] & 0x80000000u);
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i] + (float)bits[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// packed FFMA2 + LOP3 1:1 per pair (i.e. 2 lane-FMAs per LOP3)
__global__ void k_ffma2_lop(float *out, const float *in, long long *cyc)
{
    uint64_t acc[NACC], xs[NACC], ys[NACC];
    unsigned bits[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        float2 a = make_float2(threadIdx.x * 1e-3f + i, 1.0f + i);
        float2 b = make_float2(in[i], in[i + 1]);
        float2 c = make_float2(in[i + NACC], in[i + NACC + 1]);
        acc[i] = *reinterpret_cast<uint64_t *>(&a);
        xs[i] = *reinterpret_cast<uint64_t *>(&b);
        ys[i] = *reinterpret_cast<uint64_t *>(&c);
        bits[i] = threadIdx.x + i;
    }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = ffma2(xs[(i + u) % NACC], ys[i], acc[i]);
#pragma unroll
            for (int i = 0; i < NACC; i += 2) bits[i] = (bits[i] ^ (unsigned)acc[i]) | (bits[(i + 2) % NACC] & 0x80000000u);
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) { float2 a = *reinterpret_cast<float2 *>(&acc[i]); s += a.x + a.y + (float)bits[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// FFMA with one broadcast LDS.128 per 16 FFMA (the filter's ratio at R = 1)
__global__ void k_ffma_lds(float *out, const float *in, long long *cyc)
{
    __shared__ float4 tile[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) tile[i] = make_float4(in[i % 16], in[(i + 1) % 16], 1.f, 0.5f);
    __syncthreads();
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            float4 q = tile[(it * 2 + u) & 511];
#pragma unroll
            for (int i = 0; i < NACC; i += 2) {
                acc[i] = fmaf(q.x, acc[i], q.y);
                acc[i + 1] = fmaf(q.z, acc[i + 1], q.w);
            }
#pragma unroll
            for (int i = 0; i < NACC; i += 2) {
                acc[i] = fmaf(q.y, acc[i], q.x);
                acc[i + 1] = fmaf(q.w, acc[i + 1], q.z);
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <class K, class... A>
static void run(const char *name, double lane_fma_per_thread, int blocks_per_sm, int threads, K kern, float *out, long long *cyc, A... args)
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int grid = sms * blocks_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int w = 0; w < 2; ++w) kern<<<grid, threads>>>(out, args..., cyc);
    cudaEventRecord(e0);
    const int reps = 5;
    for (int r = 0; r < reps; ++r) kern<<<grid, threads>>>(out, args..., cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    long long c = 0;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    double fma = lane_fma_per_thread * (double)grid * threads;
    double per_sm_clk = lane_fma_per_thread * (double)blocks_per_sm * threads / (double)c;
    printf("%-14s grid=%d x %d  %.3f ms  %.2f TFLOP/s  block0: %lld cyc  => %.1f lane-FMA/clk/SM (%.2f GHz eff)  err=%s\n", name, grid,
           threads, ms, 2.0 * fma / (ms * 1e-3) / 1e12, c, per_sm_clk, (double)c / (ms * 1e-3) / 1e9,
           cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    float *out, *in;
    long long *cyc;
    cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
    cudaMalloc(&in, 64 * sizeof(float));
    cudaMalloc(&cyc, 8);
    float h[64];
    for (int i = 0; i < 64; ++i) h[i] = 1.0f + 1e-4f * i;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    const double per = (double)ITERS * 4 * NACC;
    for (int bps = 1; bps <= 4; bps *= 2) {
        printf("--- %d block(s) of 256 threads per SM\n", bps);
        run("ffma_imm", per, bps, 256, k_ffma, out, cyc, 1.0001f, 0.5f);
        run("ffma_3reg", per, bps, 256, k_ffma3, out, cyc, (const float *)in);
        run("ffma2", per * 2, bps, 256, k_ffma2, out, cyc, (const float *)in);
        run("ffma+lop 4:1", per, bps, 256, k_ffma_lop, out, cyc, (const float *)in);
        run("ffma2+lop", per * 2, bps, 256, k_ffma2_lop, out, cyc, (const float *)in);
        run("ffma+lds", (double)ITERS * 2 * NACC * 2, bps, 256, k_ffma_lds, out, cyc, (const float *)in);
    }
    return 0;
}
