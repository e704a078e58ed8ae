// lpcl_search.cuh - brute-force ray x triangle nearest-hit search for sm_100a.
//
// Replaces the reference `intersect` kernel (cl:243-289, launched at iterative_tracer.py:288).
//
// Design (B200-first, not a translation of the OpenCL kernel):
//   * one thread owns R rays whose per-ray constants (D, M = O x D, K) stay in registers for the
//     whole triangle sweep; rays are read once as coalesced float4;
//   * triangles are pre-digested once per scene into 64-byte FILTER RECORDS (lpcl_math.cuh) grouped
//     in tiles of TILE_TRIS that never straddle a mesh; a CTA streams the tiles of its chunk into
//     shared memory with TMA bulk copies (cp.async.bulk + mbarrier complete_tx), STAGES deep, one
//     elected thread producing, every warp consuming through full/empty mbarriers;
//   * the per-pair work is a division-free, branch-free FMA filter (19 FMA-pipe + 4 ALU ops) that
//     every lane of the warp runs on broadcast LDS.128 operands; only the rare candidates (a few
//     per ray per sweep) take the exact, reference-ordered evaluation that decides the result;
//   * per (ray, mesh) results (nearest t + triangle packed into one 64-bit key, crossing count)
//     are merged with atomicMin / atomicAdd so the triangle range can be split over blockIdx.y
//     when there are too few rays to fill 148 SMs.  Key order (t bits, triangle index) reproduces
//     the reference tie rule "first triangle index wins" (cl:278) across chunks.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "lpcl_math.cuh"

namespace lpcl {

constexpr int TILE_TRIS = 128;                  // triangles per tile
constexpr int TILE_F4 = TILE_TRIS * 4;          // float4 per tile
constexpr int TILE_BYTES = TILE_TRIS * 64;      // 8 KB
constexpr int STAGES = 4;                       // TMA pipeline depth
constexpr int SEARCH_THREADS = 256;

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
    int n_tiles, tri_count, mesh_count;
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
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
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
                                    int n_tiles, float4 *filter)
{
    long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= (long long)n_tiles * TILE_TRIS) return;
    int tile = (int)(slot / TILE_TRIS), j = (int)(slot % TILE_TRIS);
    TileInfo ti = tiles[tile];
    FilterRec r;
    if (j < ti.count) {
        float4 a = v0[ti.tri0 + j], b = v1[ti.tri0 + j], c = v2[ti.tri0 + j];
        r = make_filter_rec(v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z));
    } else { // padding: never swept (loops stop at ti.count); negative margin makes it inert anyway
        r.nx = r.ny = r.nz = 0.f; r.g = -1.f;
        r.e2x = r.e2y = r.e2z = r.ax = 0.f;
        r.ay = r.az = r.e1x = r.e1y = 0.f;
        r.e1z = r.bx = r.by = r.bz = 0.f;
    }
    float4 *dst = filter + slot * 4;
    dst[0] = make_float4(r.nx, r.ny, r.nz, r.g);
    dst[1] = make_float4(r.e2x, r.e2y, r.e2z, r.ax);
    dst[2] = make_float4(r.ay, r.az, r.e1x, r.e1y);
    dst[3] = make_float4(r.e1z, r.bx, r.by, r.bz);
}

// ------------------------------------------------------------------------------------------ scratch
__global__ void scratch_init_kernel(unsigned long long *key, int *cnt, long long total, unsigned long long init_key)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
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

// ------------------------------------------------------------------------------------------ the sweep
template <int R>
__global__ void __launch_bounds__(SEARCH_THREADS)
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
            mbar_init(&empty_bar[s], SEARCH_THREADS / 32);
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
    const long long ray0 = (long long)blockIdx.x * (R * SEARCH_THREADS) + tid;
    float dx[R], dy[R], dz[R], mx[R], my[R], mz[R], kk[R], tmin[R];
    int best[R], hits[R];
    long long ph[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        long long i = ray0 + (long long)r * SEARCH_THREADS;
        if (i < rays.n) {
            ph[r] = rays.phys(i);
            float4 o = rays.org_pow[ph[r]], d = rays.dir_prev[ph[r]];
            dx[r] = d.x; dy[r] = d.y; dz[r] = d.z;
            mx[r] = o.y * d.z - o.z * d.y;
            my[r] = o.z * d.x - o.x * d.z;
            mz[r] = o.x * d.y - o.y * d.x;
            kk[r] = filter_ray_scale(v3(o.x, o.y, o.z), v3(d.x, d.y, d.z), vmax);
        } else {
            ph[r] = -1;
            dx[r] = dy[r] = dz[r] = mx[r] = my[r] = mz[r] = 0.0f;
            kk[r] = -1.0f; // negative margin: an absent ray is never a candidate
        }
        tmin[r] = max_ray_len;
        best[r] = -1;
        hits[r] = 0;
    }
    unsigned n_cand = 0;

    auto flush = [&](int mesh) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            long long i = ray0 + (long long)r * SEARCH_THREADS;
            if (i < rays.n) {
                long long slot = (long long)mesh * rays.n + i;
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
        TileInfo ti;
        ti.tri0 = ti4.x;
        ti.count = ti4.y;
        ti.mesh = ti4.z;
        if (ti.mesh != cur_mesh) {
            if (cur_mesh >= 0) flush(cur_mesh);
            cur_mesh = ti.mesh;
        }
        mbar_wait(&full_bar[stage], (unsigned)((it / STAGES) & 1));
        const float4 *tb = tile_buf + stage * TILE_F4;

#pragma unroll 2
        for (int j = 0; j < ti.count; ++j) {
            const float4 q0 = tb[j * 4 + 0]; // Nn, G
            const float4 q1 = tb[j * 4 + 1]; // E2, A.x
            const float4 q2 = tb[j * 4 + 2]; // A.yz, E1n.xy
            const float4 q3 = tb[j * 4 + 3]; // E1n.z, B
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float den = dx[r] * q0.x;
                den = fmaf(dy[r], q0.y, den);
                den = fmaf(dz[r], q0.z, den);
                const float e = kk[r] * q0.w;
                // es = eps carrying DEN's sign:  s*x + eps >= 0  <=>  (x + s*eps) has DEN's sign
                const float es = __uint_as_float((__float_as_uint(den) & 0x80000000u) ^ __float_as_uint(e));
                float un = fmaf(mx[r], q1.x, es);
                un = fmaf(my[r], q1.y, un);
                un = fmaf(mz[r], q1.z, un);
                un = fmaf(dx[r], q1.w, un);
                un = fmaf(dy[r], q2.x, un);
                un = fmaf(dz[r], q2.y, un);
                float vn = fmaf(mx[r], q2.z, es);
                vn = fmaf(my[r], q2.w, vn);
                vn = fmaf(mz[r], q3.x, vn);
                vn = fmaf(dx[r], q3.y, vn);
                vn = fmaf(dy[r], q3.z, vn);
                vn = fmaf(dz[r], q3.w, vn);
                const float wn = (fmaf(3.0f, es, den) - un) - vn;
                const unsigned db = __float_as_uint(den);
                const unsigned bad = ((__float_as_uint(un) ^ db) | (__float_as_uint(vn) ^ db)) | (__float_as_uint(wn) ^ db);
                if ((int)bad >= 0) { // all three carry DEN's sign: candidate
                    ++n_cand;
                    if (ph[r] >= 0) {
                        const int tri = ti.tri0 + j;
                        const float t = exact_candidate(sc.v0, sc.v1, sc.v2, rays.org_pow, rays.dir_prev, ph[r], tri, eps);
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
