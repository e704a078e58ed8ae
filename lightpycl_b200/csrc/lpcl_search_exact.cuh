// lpcl_search_exact.cuh - the nearest-hit search with NO filter: every (ray, triangle) pair takes the
// exact, reference-ordered evaluation (ray_triangle_exact, cl:50-101) and the t > EPSILON /
// strict-minimum rules of the reference `intersect` kernel (cl:243-289).
//
// Selected with LPCL_SEARCH_IMPL=exact.  It is the on-GPU yardstick the two filtered searches
// (lpcl_search_tc.cuh, lpcl_search.cuh) are compared with at BASELINE size, where the CPU oracle is
// far too slow: if a conservative filter ever dropped a pair the reference accepts, the per-(ray,
// mesh) nearest t / triangle / crossing count of the filtered search would differ from this kernel's
// (tests/test_gpu_fullsize.py asserts zero differences on the full C2-C5 scenes).  Same scratch
// contract as the other searches: mesh-major packed (t, triangle) keys merged with atomicMin, counts
// with atomicAdd, so the triangle range can be split over blockIdx.y.
//
// One thread owns one ray; a CTA stages the raw vertices of one 128-triangle tile (the packed-FP32
// search's tile table: tiles never straddle a mesh) in shared memory, double buffered, and every
// thread walks the tile with warp-uniform (broadcast) reads.  ~1e12 tests/s on B200 - a checker, not
// a product path.
#pragma once

#include "lpcl_search.cuh"

namespace lpcl {

constexpr int EXACT_THREADS = 256;

__global__ void __launch_bounds__(EXACT_THREADS)
search_exact_kernel(SceneDev sc, RaysDev rays, float max_ray_len, unsigned long long *__restrict__ key,
                    int *__restrict__ cnt, int tiles_per_chunk, unsigned long long *cand_total)
{
    __shared__ float4 s_v[2][3][TILE_TRIS]; // 12 KB

    const int tid = threadIdx.x;
    const int c0 = blockIdx.y * tiles_per_chunk;
    const int c1 = min(c0 + tiles_per_chunk, sc.n_tiles);
    const long long i = (long long)blockIdx.x * EXACT_THREADS + tid;
    const bool live = i < rays.n;
    const long long ph = rays.phys(live ? i : rays.n - 1);
    const float4 o4 = rays.org_pow[ph], d4 = rays.dir_prev[ph];
    const V3 o = v3(o4.x, o4.y, o4.z), d = v3(d4.x, d4.y, d4.z);
    const float eps = eps_len(max_ray_len);

    float tmin = max_ray_len; // cl:252-254: running values start at (0, max_ray_len, -1)
    int best = -1, hits = 0, cur_mesh = -1;

    auto stage = [&](int tile, int buf) {
        const int4 ti = __ldg(reinterpret_cast<const int4 *>(sc.tiles) + tile);
        for (int k = tid; k < 3 * ti.y; k += EXACT_THREADS) {
            const int which = k / ti.y, j = k - which * ti.y;
            const float4 *src = which == 0 ? sc.v0 : (which == 1 ? sc.v1 : sc.v2);
            s_v[buf][which][j] = __ldg(src + ti.x + j);
        }
    };
    auto flush = [&](int mesh) { // cl:262-269 / 286-288
        if (live) {
            const long long slot = (long long)mesh * rays.n + i;
            if (hits) atomicAdd(&cnt[slot], hits);
            if (best >= 0) atomicMin(&key[slot], pack_key(tmin, best));
        }
        tmin = max_ray_len;
        best = -1;
        hits = 0;
    };

    if (c0 < c1) stage(c0, 0);
    __syncthreads();
    for (int tile = c0; tile < c1; ++tile) {
        const int buf = (tile - c0) & 1;
        if (tile + 1 < c1) stage(tile + 1, buf ^ 1); // the other buffer was last read one iteration ago (barrier below)
        const int4 ti = __ldg(reinterpret_cast<const int4 *>(sc.tiles) + tile);
        if (ti.z != cur_mesh) {
            if (cur_mesh >= 0) flush(cur_mesh);
            cur_mesh = ti.z;
        }
        for (int j = 0; j < ti.y; ++j) {
            const float4 a = s_v[buf][0][j], b = s_v[buf][1][j], c = s_v[buf][2][j];
            float t;
            if (ray_triangle_exact(o, d, v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), t) && t > eps) { // cl:277
                hits += 1;
                if (t < tmin) { // cl:278, strict: the first triangle index wins ties
                    tmin = t;
                    best = ti.x + j;
                }
            }
        }
        __syncthreads();
    }
    if (cur_mesh >= 0) flush(cur_mesh);
    (void)cand_total; // no filter, no candidates to report
}

} // namespace lpcl
