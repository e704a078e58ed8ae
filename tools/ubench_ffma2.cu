// ubench_ffma2.cu - which ingredient of the search loop limits FFMA2 throughput on B200?
//   A: FFMA2 scalar-broadcast x reused pair + accumulator pair, R accumulators (pure register traffic)
//   B: A + the 8 broadcast LDS.128 per pair
//   C: B + the 28 LOP3 per pair (sign transfer / candidate mask)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_ffma2 ubench_ffma2.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2_s(float a, u64 b, u64 c)
{
    u64 d;
    asm volatile("{\n.reg .b64 t;\nmov.b64 t, {%1, %1};\nfma.rn.f32x2 %0, t, %2, %3;\n}" : "=l"(d) : "f"(a), "l"(b), "l"(c));
    return d;
}
#define ITERS 4096

template <int R, int MODE>
__global__ void __launch_bounds__(256) k(float *out, const float *in, long long *cyc)
{
    __shared__ ulonglong2 tile[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        float2 a = make_float2(in[i % 16] * 1e-3f, in[(i + 1) % 16] * 1e-3f);
        tile[i].x = *reinterpret_cast<u64 *>(&a);
        tile[i].y = tile[i].x;
    }
    __syncthreads();
    float x[R][7];
    u64 acc[R][3];
    unsigned msk = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int c = 0; c < 7; ++c) x[r][c] = in[c] + 1e-3f * (threadIdx.x + r);
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[r][c] = (u64)(threadIdx.x + c + 1);
    }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        ulonglong2 l[8];
        if (MODE >= 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) l[j] = tile[((it * 8) & 1023) + j];
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) { l[j].x = acc[0][0] ^ (u64)j; l[j].y = acc[0][1] ^ (u64)j; }
        }
        const u64 *P = reinterpret_cast<const u64 *>(l);
        // 19 packed ops per ray: 3 den + 1 + 6 + 6 + 3
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r][0] = ffma2_s(x[r][c], P[c], acc[r][0]);
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r][1] = ffma2_s(x[r][6], P[3], acc[r][1]);
        if (MODE >= 2) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                unsigned lo = ((unsigned)acc[r][0] & 0x80000000u) ^ (unsigned)acc[r][1];
                unsigned hi = ((unsigned)(acc[r][0] >> 32) & 0x80000000u) ^ (unsigned)(acc[r][1] >> 32);
                acc[r][1] = ((u64)hi << 32) | lo;
            }
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) {
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r][1] = ffma2_s(x[r][c], P[4 + c], acc[r][1]);
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r][2] = ffma2_s(x[r][c], P[10 + c], acc[r][2]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r][0] = ffma2_s(x[r][c + 3], P[c], acc[r][0]);
        if (MODE >= 2) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                unsigned d0 = (unsigned)acc[r][0], d1 = (unsigned)(acc[r][0] >> 32);
                unsigned b0 = (((unsigned)acc[r][1] ^ d0) | ((unsigned)acc[r][2] ^ d0)) | (d0 ^ 0x1234u);
                unsigned b1 = (((unsigned)(acc[r][1] >> 32) ^ d1) | ((unsigned)(acc[r][2] >> 32) ^ d1)) | (d1 ^ 0x4321u);
                msk |= b0 & b1;
            }
        }
    }
    long long t1 = clock64();
    float s = (float)msk;
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) s += (float)(acc[r][c] & 0xffff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int R, int MODE>
static void run(const char *name, int bps, float *out, const float *in, long long *cyc)
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int grid = sms * bps;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<R, MODE><<<grid, 256>>>(out, in, cyc);
    cudaEventRecord(e0);
    for (int r = 0; r < 3; ++r) k<R, MODE><<<grid, 256>>>(out, in, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 3;
    double packed = (double)ITERS * 19 * R * grid * 256;
    printf("%-22s R=%d bps=%d  %.3f ms  %.2f lane-TFMA/s  = %.1f%% of 128 lanes x %d SMs x 1.965 GHz   %s\n", name, R, bps, ms,
           2 * packed / (ms * 1e-3) / 1e12, 100.0 * 2 * packed / (ms * 1e-3) / (2 * 128.0 * sms * 1.965e9), sms,
           cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    float *out, *in;
    long long *cyc;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaMalloc(&in, 64 * sizeof(float));
    cudaMalloc(&cyc, 8);
    float h[64];
    for (int i = 0; i < 64; ++i) h[i] = 1.0f + 1e-4f * i;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int bps = 1; bps <= 3; ++bps) {
        run<4, 0>("A regs only", bps, out, in, cyc);
        run<4, 1>("B +LDS", bps, out, in, cyc);
        run<4, 2>("C +LDS+LOP3", bps, out, in, cyc);
        run<8, 0>("A regs only", bps, out, in, cyc);
        run<8, 2>("C +LDS+LOP3", bps, out, in, cyc);
        run<2, 2>("C +LDS+LOP3", bps, out, in, cyc);
    }
    return 0;
}
