// lpcl_small.cuh - the WHOLE iterative trace of a small scene in ONE kernel launch.
//
// The reference example example_nested_cubes_refraction.py (BASELINE config C1: 10 rays x 48
// triangles) and anything of that size is latency-bound in the general loop: three launches plus a
// host round trip per generation (~100 us) to trace a few dozen rays.  Here one CTA runs every
// generation back to back - search, media resolution, Fresnel split, order-preserving compaction,
// the reference's two termination tests (it:372-391) - and the host reads one small table at the
// end.  Same device functions as the general path (ray_triangle_exact, resolve_media, scatter), so
// the results are bit-identical; the search is the unfiltered exact evaluation of every pair
// (cl:243-289), a warp per ray with the triangles dealt to its lanes.
//
// Memory: an arena of ray slots the caller provides.  Generation k occupies a region of
// capacity cap_k (cap_0 = n_0, cap_{k+1} = 2 n_k) in the general path's packed two-segment layout
// (reflected survivors at [0, n_refl), refracted at [n_k, n_k + n_refr)), followed by its per-ray
// results (dest_meas, optional hit record) at the same ray offsets.  The kernel stops BEFORE a
// generation that no longer qualifies (too many ray x triangle pairs for one SM, arena exhausted,
// lineage branch bits used up) and reports where it stopped; the caller continues with the general
// loop from that generation (lightpycl_b200/iterative_tracer.py).
#pragma once

#include "lpcl_material.cuh"

namespace lpcl {

constexpr int SMALL_THREADS = 1024;
constexpr int SMALL_MAX_MESHES = 8;
constexpr int SMALL_MAX_ITERS = 31; // one lineage epoch

struct SmallIter { // one row per generation, mirrors lpcl_small_iter
    long long n_rays, n_reflect, n_refract, n_measured, n_terminated;
    long long offset;   // arena slot of the generation's first ray
    double live_power, measured_power;
};

struct SmallHeader { // mirrors lpcl_small_header
    long long iterations;      // generations processed
    long long stop_reason;     // 0 = the trace ended (iteration limit, power test or no live rays), 1 = handed over
    long long next_offset, next_n, next_split, next_base; // the generation to continue with when handed over
    long long measured_total, reserved;
};

// block-wide exclusive scan of three 0/1 flags over SMALL_THREADS threads; returns this thread's three ranks and the totals
__device__ __forceinline__ void small_scan3(bool f0, bool f1, bool f2, unsigned (&s_w)[3][32], unsigned &r0, unsigned &r1,
                                            unsigned &r2, unsigned &t0, unsigned &t1, unsigned &t2)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned below = (1u << lane) - 1u;
    const unsigned b0 = __ballot_sync(0xffffffffu, f0), b1 = __ballot_sync(0xffffffffu, f1), b2 = __ballot_sync(0xffffffffu, f2);
    __syncthreads(); // s_w may still be read from the previous round
    if (lane == 0) {
        s_w[0][warp] = __popc(b0);
        s_w[1][warp] = __popc(b1);
        s_w[2][warp] = __popc(b2);
    }
    __syncthreads();
    unsigned base[3] = {0, 0, 0}, tot[3] = {0, 0, 0};
#pragma unroll
    for (int f = 0; f < 3; ++f) {
        const unsigned v = s_w[f][lane]; // 32 warps: one value per lane
        unsigned inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        tot[f] = __shfl_sync(0xffffffffu, inc, 31);
        base[f] = __shfl_sync(0xffffffffu, inc - v, warp);
    }
    r0 = base[0] + __popc(b0 & below);
    r1 = base[1] + __popc(b1 & below);
    r2 = base[2] + __popc(b2 & below);
    t0 = tot[0];
    t1 = tot[1];
    t2 = tot[2];
}

__global__ void __launch_bounds__(SMALL_THREADS, 1)
trace_small_kernel(SceneDev sc, float4 *org_pow, float4 *dir_prev, unsigned long long *lineage, float4 *dest_meas, int4 *hit_out,
                   long long arena_rays, long long n0, unsigned long long *key, int *cnt, long long scratch_rays,
                   float4 *meas_out, unsigned long long *meas_lineage, long long meas_cap, float max_ray_len, float ior_env,
                   int max_iterations, double power_threshold, long long pair_limit, SmallIter *iters, SmallHeader *header)
{
    __shared__ unsigned s_w[3][32];
    __shared__ double s_pow[2][32];
    __shared__ unsigned s_term[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = sc.mesh_count;
    const float eps = eps_len(max_ray_len);
    const unsigned long long init_key = ((unsigned long long)__float_as_uint(max_ray_len) << 32) | 0xffffffffull;

    long long off = 0, cap = n0, n = n0, split = n0, base = n0, meas_base = 0;
    int it = 0;
    long long stop_reason = 0;
    for (; it < max_iterations; ++it) {
        // hand over to the general loop before a generation this kernel should not run
        const long long child_off = off + cap; // this generation's region holds `cap` ray slots
        if (n * (long long)sc.tri_count > pair_limit || child_off + 2 * n > arena_rays || n > scratch_rays ||
            meas_base + n > meas_cap || it >= SMALL_MAX_ITERS) {
            stop_reason = 1;
            break;
        }
        RaysDev par;
        par.org_pow = org_pow + off;
        par.dir_prev = dir_prev + off;
        par.lineage = lineage ? lineage + off : nullptr;
        par.n = n;
        par.split = split;
        par.base = base;

        // ---- search (cl:243-289): a warp per ray, the triangles of every tile dealt to its lanes
        for (long long i = warp; i < n; i += SMALL_THREADS / 32) {
            const long long ph = par.phys(i);
            const float4 o4 = par.org_pow[ph], d4 = par.dir_prev[ph];
            const V3 o = v3(o4.x, o4.y, o4.z), d = v3(d4.x, d4.y, d4.z);
            int cur_mesh = -1, hits = 0;
            unsigned long long best = init_key;
            auto flush = [&](int mesh) {
                __syncwarp();
                for (int s = 16; s; s >>= 1) {
                    hits += __shfl_xor_sync(0xffffffffu, hits, s);
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, s);
                    best = other < best ? other : best; // (t bits, triangle): the first triangle index wins ties (cl:278)
                }
                if (lane == 0) {
                    key[(long long)mesh * n + i] = best;
                    cnt[(long long)mesh * n + i] = hits;
                }
                hits = 0;
                best = init_key;
            };
            for (int tile = 0; tile < sc.n_tiles; ++tile) {
                const int4 ti = __ldg(reinterpret_cast<const int4 *>(sc.tiles) + tile);
                if (ti.z != cur_mesh) {
                    if (cur_mesh >= 0) flush(cur_mesh);
                    cur_mesh = ti.z;
                }
                for (int j = lane; j < ti.y; j += 32) {
                    const int tri = ti.x + j;
                    const float4 a = __ldg(sc.v0 + tri), b = __ldg(sc.v1 + tri), c = __ldg(sc.v2 + tri);
                    float t;
                    if (ray_triangle_exact(o, d, v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), t) && t > eps) {
                        hits += 1; // cl:277: counted even beyond max_ray_len
                        if (t < max_ray_len) {
                            const unsigned long long k = pack_key(t, tri);
                            best = k < best ? k : best;
                        }
                    }
                }
            }
            if (cur_mesh >= 0) flush(cur_mesh);
        }
        __syncthreads();

        // ---- resolve, scatter, compaction: one thread per ray, SMALL_THREADS rays per round
        RaysDev chi;
        chi.org_pow = org_pow + child_off;
        chi.dir_prev = dir_prev + child_off;
        chi.lineage = lineage ? lineage + child_off : nullptr;
        long long done_r = 0, done_t = 0, done_m = 0, n_term_total = 0;
        double live_power = 0.0, measured_power = 0.0;
        for (long long r0 = 0; r0 < n; r0 += SMALL_THREADS) {
            const long long i = r0 + tid;
            const bool valid = i < n;
            bool keep_r = false, keep_t = false, is_meas = false, is_term = false;
            V3 end = v3(0.f, 0.f, 0.f), rdir = end, tdir = end;
            float rpow = 0.f, tpow = 0.f, ppow = 0.f;
            int hit_mesh = -1;
            unsigned long long lin = 0ull;
            if (valid) {
                const long long ph = par.phys(i);
                const float4 o4 = par.org_pow[ph], d4 = par.dir_prev[ph];
                lin = par.lineage ? par.lineage[ph] : 0ull;
                const V3 o = v3(o4.x, o4.y, o4.z), d = v3(d4.x, d4.y, d4.z);
                auto fetch = [&](int j, float &t, int &h, int &tri) {
                    const unsigned long long kk = key[(long long)j * n + i];
                    t = key_t(kk);
                    tri = key_tri(kk);
                    h = cnt[(long long)j * n + i];
                };
                const Resolve r = resolve_media<0>(M, __float_as_int(d4.w), sc.mat_type, max_ray_len, fetch);
                end = ray_end(o, d, r.t);
                auto tri = [&](int t, V3 &a, V3 &b, V3 &c) {
                    const float4 fa = __ldg(sc.v0 + t), fb = __ldg(sc.v1 + t), fc = __ldg(sc.v2 + t);
                    a = v3(fa.x, fa.y, fa.z);
                    b = v3(fb.x, fb.y, fb.z);
                    c = v3(fc.x, fc.y, fc.z);
                };
                const Scatter s = scatter(o, end, d, o4.w, 0, r.n1, r.n2, r.hit_mesh, r.hit_tri, sc.mat_type, sc.ior, sc.refl,
                                          sc.diss, ior_env, tri);
                dest_meas[off + i] = make_float4(end.x, end.y, end.z, __int_as_float(s.measured));
                if (hit_out) hit_out[off + i] = make_int4(r.hit_mesh, r.hit_tri, r.n1, r.n2);
                if (s.power_changed) par.org_pow[ph].w = s.power; // cl:392
                keep_r = s.refl.measured == 0;
                keep_t = s.refr.measured == 0;
                is_meas = s.measured == 1;
                is_term = s.measured == -1;
                rdir = s.refl.dir;
                tdir = s.refr.dir;
                rpow = s.refl.power;
                tpow = s.refr.power;
                ppow = s.power;
                hit_mesh = r.hit_mesh;
            }
            unsigned rank_r, rank_t, rank_m, tot_r, tot_t, tot_m;
            small_scan3(keep_r, keep_t, is_meas, s_w, rank_r, rank_t, rank_m, tot_r, tot_t, tot_m);
            if (keep_r) {
                const long long c = done_r + rank_r;
                chi.org_pow[c] = make_float4(end.x, end.y, end.z, rpow);
                chi.dir_prev[c] = make_float4(rdir.x, rdir.y, rdir.z, __int_as_float(hit_mesh));
                if (chi.lineage) chi.lineage[c] = lin;
            }
            if (keep_t) { // second segment, starting at physical index n of the children region
                const long long c = n + done_t + rank_t;
                chi.org_pow[c] = make_float4(end.x, end.y, end.z, tpow);
                chi.dir_prev[c] = make_float4(tdir.x, tdir.y, tdir.z, __int_as_float(hit_mesh));
                if (chi.lineage) chi.lineage[c] = lin | (1ull << (32 + it));
            }
            if (is_meas && meas_out) {
                const long long c = meas_base + done_m + rank_m;
                meas_out[c] = make_float4(end.x, end.y, end.z, ppow);
                if (meas_lineage) meas_lineage[c] = lin;
            }
            // power sums (float64, as the general path) and the terminated count
            double lp = (keep_r ? (double)rpow : 0.0) + (keep_t ? (double)tpow : 0.0), mp = is_meas ? (double)ppow : 0.0;
            const unsigned bx = __ballot_sync(0xffffffffu, is_term);
            for (int s = 16; s; s >>= 1) {
                lp += __shfl_xor_sync(0xffffffffu, lp, s);
                mp += __shfl_xor_sync(0xffffffffu, mp, s);
            }
            if (lane == 0) {
                s_pow[0][warp] = lp;
                s_pow[1][warp] = mp;
                s_term[warp] = __popc(bx);
            }
            __syncthreads();
            for (int w = 0; w < 32; ++w) { // every thread adds the same values in the same order: uniform totals
                live_power += s_pow[0][w];
                measured_power += s_pow[1][w];
                n_term_total += s_term[w];
            }
            done_r += tot_r;
            done_t += tot_t;
            done_m += tot_m;
            __syncthreads();
        }
        if (tid == 0) {
            SmallIter rec;
            rec.n_rays = n;
            rec.n_reflect = done_r;
            rec.n_refract = done_t;
            rec.n_measured = done_m;
            rec.n_terminated = n_term_total;
            rec.offset = off;
            rec.live_power = live_power;
            rec.measured_power = measured_power;
            iters[it] = rec;
        }
        meas_base += done_m;
        // next generation
        off = child_off;
        cap = 2 * n;
        split = done_r;
        base = n;
        n = done_r + done_t;
        // it:383-391: the power test first, then "no rays left"
        if (live_power < power_threshold || n == 0) {
            ++it;
            break;
        }
    }
    if (tid == 0) {
        SmallHeader h;
        h.iterations = it;
        h.stop_reason = stop_reason;
        h.next_offset = off;
        h.next_n = n;
        h.next_split = split;
        h.next_base = base;
        h.measured_total = meas_base;
        h.reserved = 0;
        *header = h;
    }
}

} // namespace lpcl
