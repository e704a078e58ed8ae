// lpcl_material.cuh - media resolution, Fresnel split and deterministic stream compaction.
//
//   postproc_ref_kernel / scatter_ref_kernel : the reference's `intersect_postproc` (cl:105-240) and
//       `reflect_refract_rays` (cl:346-474) with the reference's buffer layout, for seam-level parity.
//   step_kernel : the production kernel.  One pass per generation: resolve media, end point,
//       Beer-Lambert, Snell/Fresnel or mirror, terminator/measure flags, then
//       ORDER-PRESERVING compaction of the surviving children (reflected survivors first, then
//       refracted, parent order kept - exactly the order NumPy pruning produces at
//       iterative_tracer.py:366-373) and of the measured parents (get_measured_rays order,
//       it:395-411).  The three exclusive prefixes come from a single-pass chained scan
//       (decoupled look-back): a block publishes its aggregate, then inspects its predecessors
//       (lower blockIdx.x, dispatched earlier, hence always resident or finished).
//       HBM-bound: ~44 B read + 48 B triangle gather + 16 B result + 32(+8) B per child.
#pragma once

#include "lpcl_search.cuh"

namespace lpcl {

#ifndef LPCL_STEP_THREADS
#define LPCL_STEP_THREADS 256
#endif
constexpr int STEP_THREADS = LPCL_STEP_THREADS;

struct StepCounters { // mirrors lpcl_step_counters
    long long n_reflect, n_refract, n_measured, n_terminated;
    double live_power, measured_power;
    long long candidates, reserved;
};

// ------------------------------------------------------------------------------------------ reference seams
__global__ void postproc_ref_kernel(SceneDev sc, const float4 *origin, const float4 *dir, float4 *dest,
                                    const int *prev_mesh, int *n1, int *n2, int *entering, int *hit_mesh,
                                    int *hit_tri, const float *near_t, const int *hits, const int *near_tri,
                                    long long n, float max_ray_len)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int M = sc.mesh_count;
    auto fetch = [&](int j, float &t, int &h, int &tri) {
        t = near_t[i * M + j];
        h = hits[i * M + j];
        tri = near_tri[i * M + j];
    };
    Resolve r = resolve_media(M, prev_mesh[i], sc.mat_type, max_ray_len, fetch);
    float4 o = origin[i], d = dir[i];
    V3 e = ray_end(v3(o.x, o.y, o.z), v3(d.x, d.y, d.z), r.t);
    dest[i] = make_float4(e.x, e.y, e.z, 0.0f);
    n1[i] = r.n1;
    n2[i] = r.n2;
    if (r.hit_mesh >= 0) entering[i] = r.entering; // cl:143: written only on a hit
    hit_mesh[i] = r.hit_mesh;
    hit_tri[i] = r.hit_tri;
}

__global__ void scatter_ref_kernel(SceneDev sc, const float4 *origin, const float4 *dest, const float4 *dir,
                                   float *power, int *measured, const int *n1, const int *n2, float4 *refl_origin,
                                   float4 *refl_dir, float *refl_power, int *refl_meas, float4 *refr_origin,
                                   float4 *refr_dir, float *refr_power, int *refr_meas, const int *hit_mesh,
                                   const int *hit_tri, long long n, float ior_env)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 o = origin[i], e = dest[i], d = dir[i];
    auto tri = [&](int t, V3 &a, V3 &b, V3 &c) {
        float4 fa = __ldg(sc.v0 + t), fb = __ldg(sc.v1 + t), fc = __ldg(sc.v2 + t);
        a = v3(fa.x, fa.y, fa.z);
        b = v3(fb.x, fb.y, fb.z);
        c = v3(fc.x, fc.y, fc.z);
    };
    Scatter s = scatter(v3(o.x, o.y, o.z), v3(e.x, e.y, e.z), v3(d.x, d.y, d.z), power[i], measured[i], n1[i], n2[i],
                        hit_mesh[i], hit_tri[i], sc.mat_type, sc.ior, sc.refl, sc.diss, ior_env, tri);
    if (s.power_changed) power[i] = s.power;
    measured[i] = s.measured;
    float4 at = make_float4(e.x, e.y, e.z, 0.0f); // children start at the end point (cl:319,324,456,461)
    refl_origin[i] = at;
    refl_dir[i] = make_float4(s.refl.dir.x, s.refl.dir.y, s.refl.dir.z, 0.0f);
    refl_power[i] = s.refl.power;
    refl_meas[i] = s.refl.measured;
    refr_origin[i] = at;
    refr_dir[i] = make_float4(s.refr.dir.x, s.refr.dir.y, s.refr.dir.z, 0.0f);
    refr_power[i] = s.refr.power;
    refr_meas[i] = s.refr.measured;
}

__global__ void pack_rays_kernel(const float4 *origin, const float4 *dir, const float *power, const int *prev_mesh,
                                 long long n, long long lineage_base, float4 *org_pow, float4 *dir_prev,
                                 unsigned long long *lineage)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 o = origin[i], d = dir[i];
    org_pow[i] = make_float4(o.x, o.y, o.z, power[i]);
    dir_prev[i] = make_float4(d.x, d.y, d.z, __int_as_float(prev_mesh ? prev_mesh[i] : -2));
    if (lineage) lineage[i] = (unsigned long long)(lineage_base + i) & 0xffffffffull;
}

// ------------------------------------------------------------------------------------------ chained scan
// Look-back words: [63:62] status (0 empty, 1 block aggregate, 2 inclusive prefix).
//   word A: [61:31] reflected survivors, [30:0] refracted survivors      word B: [61:0] measured parents
constexpr unsigned long long ST_AGG = 1ull << 62, ST_INC = 2ull << 62, ST_MASK = 3ull << 62;

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Called by warp 0.  agg_a / agg_b: this block's packed aggregates.  Returns the exclusive prefixes.
__device__ __forceinline__ void chained_scan(unsigned long long *words, int bid, int lane, unsigned long long agg_a,
                                             unsigned long long agg_b, unsigned long long &excl_a,
                                             unsigned long long &excl_b)
{
    const unsigned long long VAL = ~ST_MASK;
    unsigned long long ea = 0, eb = 0;
    if (bid > 0) {
        if (lane == 0) {
            st_relaxed_u64(&words[2 * bid], ST_AGG | agg_a);
            st_relaxed_u64(&words[2 * bid + 1], ST_AGG | agg_b);
        }
        int look = bid - 1;
        while (true) {
            const int b = look - lane;
            unsigned long long wa = ST_INC, wb = ST_INC; // blocks before 0 behave as an inclusive prefix of 0
            if (b >= 0) {
                do {
                    wa = ld_relaxed_u64(&words[2 * b]);
                } while ((wa & ST_MASK) == 0);
                do {
                    wb = ld_relaxed_u64(&words[2 * b + 1]);
                } while ((wb & ST_MASK) == 0);
            }
            // nearest predecessor holding an inclusive prefix in BOTH words
            const bool inc = ((wa & ST_MASK) == ST_INC) && ((wb & ST_MASK) == ST_INC);
            const unsigned inc_mask = __ballot_sync(0xffffffffu, inc);
            const int stop = inc_mask ? (__ffs(inc_mask) - 1) : 32;
            // lanes [0, stop) contribute aggregates, lane `stop` its inclusive prefix.  A word whose
            // status is INC while its sibling is still AGG must not be used as a prefix: in that
            // case (rare) treat the lane as "not yet inclusive" and use nothing from it but wait.
            bool usable = true;
            if (lane < stop) usable = ((wa & ST_MASK) == ST_AGG) && ((wb & ST_MASK) == ST_AGG);
            const unsigned unusable = __ballot_sync(0xffffffffu, !usable && lane < stop);
            if (unusable) continue; // re-read: the publishing block is between its two stores
            unsigned long long ca = (lane <= stop) ? (wa & VAL) : 0ull;
            unsigned long long cb = (lane <= stop) ? (wb & VAL) : 0ull;
            for (int o = 16; o; o >>= 1) {
                ca += __shfl_xor_sync(0xffffffffu, ca, o);
                cb += __shfl_xor_sync(0xffffffffu, cb, o);
            }
            ea += ca;
            eb += cb;
            if (inc_mask) break;
            look -= 32;
        }
    }
    if (lane == 0) {
        st_relaxed_u64(&words[2 * bid], ST_INC | (ea + agg_a));
        st_relaxed_u64(&words[2 * bid + 1], ST_INC | (eb + agg_b));
    }
    excl_a = ea;
    excl_b = eb;
}

// ------------------------------------------------------------------------------------------ production step
__global__ void __launch_bounds__(STEP_THREADS)
step_kernel(SceneDev sc, RaysDev par, RaysDev chi, const unsigned long long *__restrict__ key,
            const int *__restrict__ cnt, float4 *__restrict__ dest_meas, int4 *__restrict__ hit_out,
            float4 *__restrict__ meas_out, unsigned long long *__restrict__ meas_lineage, long long meas_base,
            long long meas_cap, StepCounters *ctr,
            unsigned long long *words, unsigned *ticket, float max_ray_len, float ior_env, int iteration)
{
    __shared__ unsigned s_warp[3][STEP_THREADS / 32];
    __shared__ unsigned long long s_excl[2];
    __shared__ double s_pow[2][STEP_THREADS / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Block id = blockIdx.x: the look-back only ever waits on LOWER block ids, which a 1-D grid
    // dispatches first (the assumption CUB's single-pass scan makes too).  An atomic ticket was
    // measured to cost half of the kernel (3,907 same-address atomics serialise in L2).
    (void)ticket;
    const int bid = (int)blockIdx.x;
    const long long n = par.n;
    const long long i = (long long)bid * STEP_THREADS + tid;
    const bool valid = i < n;

    bool keep_r = false, keep_t = false, is_meas = false, is_term = false;
    V3 end = v3(0.f, 0.f, 0.f);
    Scatter s;
    s.power = 0.f;
    s.refl.power = s.refr.power = 0.f;
    s.refl.dir = s.refr.dir = v3(0.f, 0.f, 0.f);
    int hit_mesh = -1;
    unsigned long long lin = 0;

    if (valid) {
        const long long ph = par.phys(i);
        const float4 o4 = par.org_pow[ph], d4 = par.dir_prev[ph];
        const V3 o = v3(o4.x, o4.y, o4.z), d = v3(d4.x, d4.y, d4.z);
        const int prev = __float_as_int(d4.w);
        const int M = sc.mesh_count;
        if (par.lineage) lin = par.lineage[ph];
        Resolve r;
        constexpr int MREG = 8;
        if (M <= MREG) {
            // few meshes (every BASELINE config): issue all 2M scratch loads at once, resolve from registers
            unsigned long long kreg[MREG];
            int creg[MREG];
#pragma unroll
            for (int j = 0; j < MREG; ++j) {
                kreg[j] = 0;
                creg[j] = 0;
                if (j < M) {
                    kreg[j] = key[(long long)j * n + i];
                    creg[j] = cnt[(long long)j * n + i];
                }
            }
            auto fetch = [&](int j, float &t, int &h, int &tri) {
                t = key_t(kreg[j]);
                tri = key_tri(kreg[j]);
                h = creg[j];
            };
            r = resolve_media<MREG>(M, prev, sc.mat_type, max_ray_len, fetch);
        } else {
            auto fetch = [&](int j, float &t, int &h, int &tri) {
                unsigned long long k = key[(long long)j * n + i];
                t = key_t(k);
                tri = key_tri(k);
                h = cnt[(long long)j * n + i];
            };
            r = resolve_media<0>(M, prev, sc.mat_type, max_ray_len, fetch);
        }
        end = ray_end(o, d, r.t);
        auto tri = [&](int t, V3 &a, V3 &b, V3 &c) {
            float4 fa = __ldg(sc.v0 + t), fb = __ldg(sc.v1 + t), fc = __ldg(sc.v2 + t);
            a = v3(fa.x, fa.y, fa.z);
            b = v3(fb.x, fb.y, fb.z);
            c = v3(fc.x, fc.y, fc.z);
        };
        s = scatter(o, end, d, o4.w, 0, r.n1, r.n2, r.hit_mesh, r.hit_tri, sc.mat_type, sc.ior, sc.refl, sc.diss,
                    ior_env, tri);
        hit_mesh = r.hit_mesh;
        dest_meas[i] = make_float4(end.x, end.y, end.z, __int_as_float(s.measured));
        if (hit_out) hit_out[i] = make_int4(r.hit_mesh, r.hit_tri, r.n1, r.n2);
        if (s.power_changed) par.org_pow[ph].w = s.power; // cl:392
        keep_r = s.refl.measured == 0;
        keep_t = s.refr.measured == 0;
        is_meas = s.measured == 1;
        is_term = s.measured == -1;
    }

    // block-level exclusive ranks of the three flags
    const unsigned br = __ballot_sync(0xffffffffu, keep_r), bt = __ballot_sync(0xffffffffu, keep_t),
                   bm = __ballot_sync(0xffffffffu, is_meas), bx = __ballot_sync(0xffffffffu, is_term);
    const unsigned below = (1u << lane) - 1u;
    unsigned rank_r = __popc(br & below), rank_t = __popc(bt & below), rank_m = __popc(bm & below);
    double lp = (keep_r ? (double)s.refl.power : 0.0) + (keep_t ? (double)s.refr.power : 0.0);
    double mp = is_meas ? (double)s.power : 0.0;
    for (int o = 16; o; o >>= 1) {
        lp += __shfl_xor_sync(0xffffffffu, lp, o);
        mp += __shfl_xor_sync(0xffffffffu, mp, o);
    }
    if (lane == 0) {
        s_warp[0][warp] = __popc(br);
        s_warp[1][warp] = __popc(bt);
        s_warp[2][warp] = __popc(bm);
        s_pow[0][warp] = lp;
        s_pow[1][warp] = mp;
        if (bx) atomicAdd((unsigned long long *)&ctr->n_terminated, (unsigned long long)__popc(bx));
    }
    __syncthreads();
    unsigned tot_r = 0, tot_t = 0, tot_m = 0;
#pragma unroll
    for (int w = 0; w < STEP_THREADS / 32; ++w) {
        const unsigned a = s_warp[0][w], b = s_warp[1][w], c = s_warp[2][w];
        if (w < warp) {
            rank_r += a;
            rank_t += b;
            rank_m += c;
        }
        tot_r += a;
        tot_t += b;
        tot_m += c;
    }
    if (warp == 0) {
        unsigned long long ea, eb;
        chained_scan(words, bid, lane, ((unsigned long long)tot_r << 31) | tot_t, tot_m, ea, eb);
        if (lane == 0) {
            s_excl[0] = ea;
            s_excl[1] = eb;
            double blp = 0.0, bmp = 0.0;
            for (int w = 0; w < STEP_THREADS / 32; ++w) {
                blp += s_pow[0][w];
                bmp += s_pow[1][w];
            }
            if (blp != 0.0) atomicAdd(&ctr->live_power, blp);
            if (bmp != 0.0) atomicAdd(&ctr->measured_power, bmp);
            if ((long long)(bid + 1) * STEP_THREADS >= n) { // last block: inclusive prefixes are the totals
                ctr->n_reflect = (long long)(ea >> 31) + tot_r;
                ctr->n_refract = (long long)(ea & 0x7fffffffull) + tot_t;
                ctr->n_measured = (long long)eb + tot_m;
            }
        }
    }
    __syncthreads();
    const long long base_r = (long long)(s_excl[0] >> 31), base_t = (long long)(s_excl[0] & 0x7fffffffull);
    const long long base_m = (long long)s_excl[1];

    if (keep_r) { // reflected survivors: first segment of the next generation
        const long long c = base_r + rank_r;
        chi.org_pow[c] = make_float4(end.x, end.y, end.z, s.refl.power);
        chi.dir_prev[c] = make_float4(s.refl.dir.x, s.refl.dir.y, s.refl.dir.z, __int_as_float(hit_mesh));
        if (chi.lineage) chi.lineage[c] = lin;
    }
    if (keep_t) { // refracted survivors: second segment, starting at physical index n
        const long long c = n + base_t + rank_t;
        chi.org_pow[c] = make_float4(end.x, end.y, end.z, s.refr.power);
        chi.dir_prev[c] = make_float4(s.refr.dir.x, s.refr.dir.y, s.refr.dir.z, __int_as_float(hit_mesh));
        if (chi.lineage) chi.lineage[c] = (iteration < 31) ? (lin | (1ull << (32 + iteration))) : lin;
    }
    if (is_meas && meas_out) {
        const long long c = meas_base + base_m + rank_m;
        if (c < meas_cap) {
            meas_out[c] = make_float4(end.x, end.y, end.z, s.power);
            if (meas_lineage) meas_lineage[c] = lin;
        }
    }
}

} // namespace lpcl
