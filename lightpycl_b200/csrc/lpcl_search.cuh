// lpcl_search.cuh - brute-force ray x triangle nearest-hit search for sm_100a.
//
// Replaces the reference `intersect` kernel (cl:243-289, launched at iterative_tracer.py:288).
//
// Design (B200-first, not a translation of the OpenCL kernel):
//   * one thread owns R rays whose per-ray constants (D, M = O x D, K) stay in registers for the
//     whole triangle sweep; rays are read once as coalesced float4;
//   * triangles are pre-digested once per scene into 64-byte FILTER RECORDS (lpcl_math.cuh), stored
//     PAIR-INTERLEAVED (component k of triangles 2p and 2p+1 adjacent) in tiles of TILE_TRIS that
//     never straddle a mesh; a CTA streams the tiles of its chunk into shared memory with TMA bulk
//     copies (cp.async.bulk + mbarrier complete_tx), STAGES deep, one elected thread producing,
//     every warp consuming through full/empty mbarriers;
//   * the per-pair work is a division-free, branch-free filter evaluated with PACKED FP32
//     (fma.rn.f32x2 -> FFMA2: two triangles per instruction, the ray value broadcast as the scalar
//     .F32 operand).  Measured on B200 (tools/ubench_fp32.cu): a 3-register FFMA issues every 2
//     cycles (register-bank bandwidth), so the scalar formulation tops out near 55% of the FMA
//     peak; FFMA2 with the triangle pair held in the operand-reuse cache across the R rays of a
//     thread reads one even + two odd registers and runs at the full lane rate with half the issue
//     slots.  The filter is two sign-agnostic products (lpcl_math.cuh): 19.25 FMA-pipe lane-ops
//     + 1 ALU op ((a|b)&c on the sign bits) per test; measured 1.45e12 tests/s per B200 with the
//     FMA pipe 77 % busy (profiles/);
//   * only the rare candidates (a few per ray per sweep) take the exact, reference-ordered
//     evaluation that decides the result;
//   * per (ray, mesh) results (nearest t + triangle packed into one 64-bit key, crossing count)
//     are merged with atomicMin / atomicAdd so the triangle range can be split over blockIdx.y
//     (used to fill the machine when rays are few and to trim the last-wave tail when they are
//     many).  Key order (t bits, triangle index) reproduces the reference tie rule "first triangle
//     index wins" (cl:278) across chunks.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "lpcl_math.cuh"

namespace lpcl {

constexpr int TILE_TRIS = 128;                  // triangles per tile
constexpr int TILE_F4 = TILE_TRIS * 4;          // float4 per tile
constexpr int TILE_BYTES = TILE_TRIS * 64;      // 8 KB
constexpr int STAGES = 4;                       // TMA pipeline depth
#ifndef LPCL_PAIR_UNROLL
#define LPCL_PAIR_UNROLL 4                      // unroll factor of the triangle-pair loop (measured: 1 -> 4 = +3%)
#endif
#define LPCL_STR_(x) #x
#define LPCL_UNROLL(n) _Pragma(LPCL_STR_(unroll n))

struct TileInfo {
    int tri0;   // global index of the tile's first triangle
    int count;  // real triangles in the tile (<= TILE_TRIS)
    int mesh;   // mesh id (tiles never straddle meshes)
    int pad;
};

struct SceneDev {
    const float4 *v0, *v1, *v2;   // caller-owned (T,4)
    const float4 *filter;         // n_tiles * TILE_F4
    const TileInfo *tiles;        // n_tiles
    const int *mat_type;          // mesh_count
    const float *ior, *refl, *diss;
    const float *vmax;            // device scalar: max vertex norm
    const float *tc_filter;       // tc_n_tiles * TC_TILE_FLOATS (lpcl_search_tc.cuh)
    const TileInfo *tc_tiles;     // tc_n_tiles
    int n_tiles, tri_count, mesh_count, tc_n_tiles;
};

// A ray generation stored in two segments (see include/lightpycl_b200.h).
struct RaysDev {
    float4 *org_pow;
    float4 *dir_prev;
    unsigned long long *lineage;
    long long n, split, base;
    __device__ __forceinline__ long long phys(long long i) const { return i < split ? i : base + (i - split); }
};

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LPCL_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n" // suspend-time hint: sleep, do not spin
        "@p bra LPCL_DONE;\n"
        "bra LPCL_WAIT;\n"
        "LPCL_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_bulk(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ unsigned long long pack_key(float t, int tri)
{
    return ((unsigned long long)__float_as_uint(t) << 32) | (unsigned long long)(unsigned)tri;
}
__device__ __forceinline__ float key_t(unsigned long long k) { return __uint_as_float((unsigned)(k >> 32)); }
__device__ __forceinline__ int key_tri(unsigned long long k) { return (int)(unsigned)(k & 0xffffffffull); }

// ------------------------------------------------------------------------------------------ scene build
__global__ void vmax_kernel(const float4 *v0, const float4 *v1, const float4 *v2, int tri_count, unsigned *vmax2_bits)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float m = 0.0f;
    if (i < tri_count) {
        float4 a = v0[i], b = v1[i], c = v2[i];
        m = fmaxf(fmaxf(a.x * a.x + a.y * a.y + a.z * a.z, b.x * b.x + b.y * b.y + b.z * b.z),
                  c.x * c.x + c.y * c.y + c.z * c.z);
    }
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(vmax2_bits, __float_as_uint(m)); // positive floats order as uints
}

__global__ void vmax_finish_kernel(const unsigned *vmax2_bits, float *vmax)
{
    *vmax = sqrtf(__uint_as_float(*vmax2_bits)) * 1.0001f;
}

__global__ void build_filter_kernel(const float4 *v0, const float4 *v1, const float4 *v2, const TileInfo *tiles,
                                    int n_tiles, const float *vmax, float4 *filter)
{
    long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= (long long)n_tiles * TILE_TRIS) return;
    int tile = (int)(slot / TILE_TRIS), j = (int)(slot % TILE_TRIS);
    TileInfo ti = tiles[tile];
    FilterRec r;
    if (j < ti.count) {
        float4 a = v0[ti.tri0 + j], b = v1[ti.tri0 + j], c = v2[ti.tri0 + j];
        r = make_filter_rec(v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), *vmax);
    } else { // padding: never swept (loops stop at ti.count); negative margin makes it inert anyway
        r.nx = r.ny = r.nz = 0.f; r.g = -1.f;
        r.e2x = r.e2y = r.e2z = r.ax = 0.f;
        r.ay = r.az = r.e1x = r.e1y = 0.f;
        r.e1z = r.bx = r.by = r.bz = 0.f;
    }
    // pair-interleaved: component k of triangles (2p, 2p+1) sits at float index p*32 + 2k + (j & 1)
    float *dst = reinterpret_cast<float *>(filter) + (slot / 2) * 32 + (slot & 1);
    const float comp[16] = {r.nx, r.ny, r.nz, r.g, r.e2x, r.e2y, r.e2z, r.ax, r.ay, r.az, r.e1x, r.e1y, r.e1z, r.bx, r.by, r.bz};
#pragma unroll
    for (int k = 0; k < 16; ++k) dst[2 * k] = comp[k];
}

// ------------------------------------------------------------------------------------------ scratch
// Also clears up to three small 64-bit-word areas (the step counters, the look-back words of the
// step kernel's chained scan, the record counter of the deferred-flag search), so that one generation needs no separate memset launches.
__global__ void scratch_init_kernel(unsigned long long *key, int *cnt, long long total, unsigned long long init_key,
                                    unsigned long long *zero_a, long long words_a, unsigned long long *zero_b,
                                    long long words_b, unsigned long long *zero_c = nullptr, long long words_c = 0)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k = i; k < words_a; k += stride) zero_a[k] = 0ull;
    for (long long k = i; k < words_b; k += stride) zero_b[k] = 0ull;
    for (long long k = i; k < words_c; k += stride) zero_c[k] = 0ull;
    for (; i < total; i += stride) {
        key[i] = init_key;
        cnt[i] = 0;
    }
}

// mesh-major scratch -> the reference's three (n, mesh_count) row-major arrays
__global__ void scratch_unpack_kernel(const unsigned long long *key, const int *cnt, long long n, int mesh_count,
                                      float *near_t, int *hits, int *near_tri)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int m = 0; m < mesh_count; ++m) {
        unsigned long long k = key[(long long)m * n + i];
        near_t[i * mesh_count + m] = key_t(k);
        near_tri[i * mesh_count + m] = key_tri(k);
        hits[i * mesh_count + m] = cnt[(long long)m * n + i];
    }
}

// ------------------------------------------------------------------------------------------ exact path
// Rare: a few calls per ray per sweep.  Re-reads the ray and the triangle's vertices from global
// memory and evaluates cl:50-101 + the t > EPSILON test of cl:277 with reference rounding.
__device__ __noinline__ float exact_candidate(const float4 *tv0, const float4 *tv1, const float4 *tv2,
                                              const float4 *org_pow, const float4 *dir_prev, long long phys, int tri,
                                              float eps)
{
    float4 o = org_pow[phys], d = dir_prev[phys];
    float4 a = __ldg(tv0 + tri), b = __ldg(tv1 + tri), c = __ldg(tv2 + tri);
    float t;
    if (ray_triangle_exact(v3(o.x, o.y, o.z), v3(d.x, d.y, d.z), v3(a.x, a.y, a.z), v3(b.x, b.y, b.z),
                           v3(c.x, c.y, c.z), t) &&
        t > eps)
        return t;
    return -1.0f;
}

// ------------------------------------------------------------------------------------------ packed FP32 helpers
typedef unsigned long long u64;
// d = a * b + c on two lanes; `a` is a scalar broadcast to both lanes (SASS: FFMA2 Rd, Ra.F32, Rb, Rc)
__device__ __forceinline__ u64 ffma2_s(float a, u64 b, u64 c)
{
    u64 d;
    asm("{\n.reg .b64 t;\nmov.b64 t, {%1, %1};\nfma.rn.f32x2 %0, t, %2, %3;\n}" : "=l"(d) : "f"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 fmul2_s(float a, u64 b)
{
    u64 d;
    asm("{\n.reg .b64 t;\nmov.b64 t, {%1, %1};\nmul.rn.f32x2 %0, t, %2;\n}" : "=l"(d) : "f"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 fsub2(u64 a, u64 b)
{
    u64 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// (a | b) & c as ONE LOP3 (opaque to the optimiser, which would otherwise materialise a | b)
__device__ __forceinline__ unsigned or_and(unsigned a, unsigned b, unsigned c)
{
    unsigned d;
    asm("lop3.b32 %0, %1, %2, %3, 0xA8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ unsigned lo32(u64 x) { return (unsigned)x; }
__device__ __forceinline__ unsigned hi32(u64 x) { return (unsigned)(x >> 32); }
__device__ __forceinline__ u64 pack2(unsigned lo, unsigned hi) { return ((u64)hi << 32) | lo; }

// ------------------------------------------------------------------------------------------ the sweep
template <int R, int THREADS>
__global__ void __launch_bounds__(THREADS)
search_kernel(SceneDev sc, RaysDev rays, float max_ray_len, unsigned long long *__restrict__ key,
              int *__restrict__ cnt, int tiles_per_chunk, unsigned long long *cand_total)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *tile_buf = reinterpret_cast<float4 *>(smem_raw);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem_raw + STAGES * TILE_BYTES);
    uint64_t *empty_bar = full_bar + STAGES;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int c0 = blockIdx.y * tiles_per_chunk;
    const int c1 = min(c0 + tiles_per_chunk, sc.n_tiles);
    const int nt = c1 - c0;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], THREADS / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // producer prologue: fill every stage
    if (tid == 0) {
        for (int s = 0; s < STAGES && s < nt; ++s) {
            mbar_arrive_expect_tx(&full_bar[s], TILE_BYTES);
            tma_load_bulk(tile_buf + s * TILE_F4, sc.filter + (long long)(c0 + s) * TILE_F4, TILE_BYTES, &full_bar[s]);
        }
    }

    // per-ray constants
    const float eps = eps_len(max_ray_len);
    const float vmax = *sc.vmax;
    const long long ray0 = (long long)blockIdx.x * (R * THREADS) + tid;
    float dx[R], dy[R], dz[R], mx[R], my[R], mz[R], tmin[R];
    int best[R], hits[R];
    float kmax = 0.0f; // margin scale K2 shared by the thread's R rays (their maximum): one FMUL2 per pair
#pragma unroll
    for (int r = 0; r < R; ++r) {
        // a lane beyond the last ray sweeps a copy of the last ray (same cost, never flushed)
        const long long i = min(ray0 + (long long)r * THREADS, rays.n - 1);
        const long long ph = rays.phys(i);
        const float4 o = rays.org_pow[ph], d = rays.dir_prev[ph];
        dx[r] = d.x; dy[r] = d.y; dz[r] = d.z;
        mx[r] = o.y * d.z - o.z * d.y;
        my[r] = o.z * d.x - o.x * d.z;
        mz[r] = o.x * d.y - o.y * d.x;
        kmax = fmaxf(kmax, filter_ray_scale(v3(o.x, o.y, o.z), v3(d.x, d.y, d.z), vmax));
        tmin[r] = max_ray_len;
        best[r] = -1;
        hits[r] = 0;
    }
    unsigned n_cand = 0;

    auto flush = [&](int mesh) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long i = ray0 + (long long)r * THREADS;
            if (i < rays.n) {
                const long long slot = (long long)mesh * rays.n + i;
                if (hits[r]) atomicAdd(&cnt[slot], hits[r]);
                if (best[r] >= 0) atomicMin(&key[slot], pack_key(tmin[r], best[r]));
            }
            tmin[r] = max_ray_len;
            best[r] = -1;
            hits[r] = 0;
        }
    };

    int cur_mesh = -1;
    for (int it = 0; it < nt; ++it) {
        const int stage = it % STAGES;
        const int4 ti4 = __ldg(reinterpret_cast<const int4 *>(sc.tiles) + (c0 + it));
        const int tri0 = ti4.x, count = ti4.y, mesh = ti4.z;
        if (mesh != cur_mesh) {
            if (cur_mesh >= 0) flush(cur_mesh);
            cur_mesh = mesh;
        }
        mbar_wait(&full_bar[stage], (unsigned)((it / STAGES) & 1));
        const ulonglong2 *tb = reinterpret_cast<const ulonglong2 *>(tile_buf + stage * TILE_F4);
        const int npairs = (count + 1) >> 1;

        LPCL_UNROLL(LPCL_PAIR_UNROLL)
        for (int p = 0; p < npairs; ++p) {
            // 8 x LDS.128 (warp-uniform address => broadcast): 16 components x 2 triangles
            const ulonglong2 l0 = tb[p * 8 + 0], l1 = tb[p * 8 + 1], l2 = tb[p * 8 + 2], l3 = tb[p * 8 + 3];
            const ulonglong2 l4 = tb[p * 8 + 4], l5 = tb[p * 8 + 5], l6 = tb[p * 8 + 6], l7 = tb[p * 8 + 7];
            u64 den[R], un[R], vn[R];
            // component-major order: the triangle pair is the same operand in R consecutive
            // FFMA2s, so it is served by the operand-reuse cache
#pragma unroll
            for (int r = 0; r < R; ++r) den[r] = fmul2_s(dx[r], l0.x);          // Nn.x
#pragma unroll
            for (int r = 0; r < R; ++r) un[r] = fmul2_s(mx[r], l2.x);           // E2.x
#pragma unroll
            for (int r = 0; r < R; ++r) vn[r] = fmul2_s(mx[r], l5.x);           // E1n.x
#pragma unroll
            for (int r = 0; r < R; ++r) den[r] = ffma2_s(dy[r], l0.y, den[r]);  // Nn.y
#pragma unroll
            for (int r = 0; r < R; ++r) un[r] = ffma2_s(my[r], l2.y, un[r]);    // E2.y
#pragma unroll
            for (int r = 0; r < R; ++r) vn[r] = ffma2_s(my[r], l5.y, vn[r]);    // E1n.y
#pragma unroll
            for (int r = 0; r < R; ++r) den[r] = ffma2_s(dz[r], l1.x, den[r]);  // Nn.z
#pragma unroll
            for (int r = 0; r < R; ++r) un[r] = ffma2_s(mz[r], l3.x, un[r]);    // E2.z
#pragma unroll
            for (int r = 0; r < R; ++r) vn[r] = ffma2_s(mz[r], l6.x, vn[r]);    // E1n.z
#pragma unroll
            for (int r = 0; r < R; ++r) un[r] = ffma2_s(dx[r], l3.y, un[r]);    // A.x
#pragma unroll
            for (int r = 0; r < R; ++r) vn[r] = ffma2_s(dx[r], l6.y, vn[r]);    // B.x
#pragma unroll
            for (int r = 0; r < R; ++r) un[r] = ffma2_s(dy[r], l4.x, un[r]);    // A.y
#pragma unroll
            for (int r = 0; r < R; ++r) vn[r] = ffma2_s(dy[r], l7.x, vn[r]);    // B.y
#pragma unroll
            for (int r = 0; r < R; ++r) un[r] = ffma2_s(dz[r], l4.y, un[r]);    // A.z
#pragma unroll
            for (int r = 0; r < R; ++r) vn[r] = ffma2_s(dz[r], l7.y, vn[r]);    // B.z
            const u64 m = fmul2_s(kmax, l1.y);                                  // margin = K2 * G2_tri
            u64 pu[R], pv[R];
            unsigned none = 0xffffffffu; // sign bit stays set only if every test of the pair failed
#pragma unroll
            for (int r = 0; r < R; ++r) {
                pu[r] = ffma2(un[r], fsub2(den[r], un[r]), m);                  // uN (DEN - uN) + m
                pv[r] = ffma2(vn[r], fsub2(den[r], vn[r]), m);                  // vN (DEN - vN) + m
                none = or_and(lo32(pu[r]), lo32(pv[r]), none);                  // one LOP3 per test
                none = or_and(hi32(pu[r]), hi32(pv[r]), none);
            }
            if ((int)none >= 0) { // rare: some (ray, triangle) of this pair passed both product tests
#pragma unroll
                for (int r = 0; r < R; ++r) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        // (separate sign tests here keep the fused (a|b)&c LOP3 form in the hot path)
                        const int su = (int)(h ? hi32(pu[r]) : lo32(pu[r])), sv = (int)(h ? hi32(pv[r]) : lo32(pv[r]));
                        if (su >= 0 && sv >= 0) {
                            ++n_cand;
                            const long long i = ray0 + (long long)r * THREADS;
                            const int j = 2 * p + h;
                            if (i < rays.n && j < count) {
                                const int tri = tri0 + j;
                                const float t = exact_candidate(sc.v0, sc.v1, sc.v2, rays.org_pow, rays.dir_prev,
                                                                rays.phys(i), tri, eps);
                                if (t > 0.0f) {
                                    hits[r] += 1;
                                    if (t < tmin[r]) { // strict: first triangle index wins ties (cl:278)
                                        tmin[r] = t;
                                        best[r] = tri;
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }

        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        // producer: refill the stage consumed one tile ago (every warp is almost surely past it)
        if (tid == 0 && it >= 1) {
            const int j = it - 1;
            if (j + STAGES < nt) {
                const int s = j % STAGES;
                mbar_wait(&empty_bar[s], (unsigned)((j / STAGES) & 1));
                mbar_arrive_expect_tx(&full_bar[s], TILE_BYTES);
                tma_load_bulk(tile_buf + s * TILE_F4, sc.filter + (long long)(c0 + j + STAGES) * TILE_F4, TILE_BYTES,
                              &full_bar[s]);
            }
        }
    }
    if (cur_mesh >= 0) flush(cur_mesh);

    if (cand_total) {
        for (int o = 16; o; o >>= 1) n_cand += __shfl_xor_sync(0xffffffffu, n_cand, o);
        if (lane == 0 && n_cand) atomicAdd(cand_total, (unsigned long long)n_cand);
    }
}

constexpr int SEARCH_SMEM_BYTES = STAGES * TILE_BYTES + 2 * STAGES * 8;

} // namespace lpcl
