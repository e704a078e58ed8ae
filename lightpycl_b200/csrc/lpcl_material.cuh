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
#ifndef LPCL_STEP_ITEMS
#define LPCL_STEP_ITEMS 2                        // measured on B200 (C3, 1M rays): 1 -> 80 us, 2 -> 70 us, 4 -> 74 us
#endif
constexpr int STEP_ITEMS = LPCL_STEP_ITEMS;      // rays per thread of the step kernel
constexpr int STEP_RAYS = STEP_THREADS * STEP_ITEMS;

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
// One thread handles STEP_ITEMS rays: item k of thread tid is ray  block0 + k*STEP_THREADS + tid
// (coalesced per item).  One block-wide scan and ONE look-back serve STEP_THREADS*STEP_ITEMS rays,
// and the independent items give the memory system several loads in flight per thread.
struct StepItem {
    V3 end, rdir, tdir;        // end point, reflected / refracted direction
    float rpow, tpow, ppow;    // child powers, parent power after dissipation
    int hit_mesh;
    unsigned long long lin;
    unsigned bits;             // [0] keep_r [1] keep_t [2] measured [3] terminated | rank_r<<8 | rank_t<<16 | rank_m<<24
};

__global__ void __launch_bounds__(STEP_THREADS)
step_kernel(SceneDev sc, RaysDev par, RaysDev chi, const unsigned long long *__restrict__ key,
            const int *__restrict__ cnt, float4 *__restrict__ dest_meas, int4 *__restrict__ hit_out,
            float4 *__restrict__ meas_out, unsigned long long *__restrict__ meas_lineage, long long meas_base,
            long long meas_cap, StepCounters *ctr,
            unsigned long long *words, float max_ray_len, float ior_env, int iteration)
{
    constexpr int NW = STEP_THREADS / 32;          // warps per block
    constexpr int SLOTS = NW * STEP_ITEMS;         // (item, warp) pairs in ray order, item-major
    static_assert(SLOTS <= 32, "the (item, warp) totals are scanned by one warp");
    __shared__ unsigned s_cnt[3][SLOTS];           // per (item, warp): survivors / measured
    __shared__ unsigned s_off[3][SLOTS];           // their exclusive prefix inside the block
    __shared__ unsigned long long s_excl[2];       // block-exclusive prefixes from the look-back
    __shared__ double s_pow[2][NW];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Block id = blockIdx.x: the look-back only ever waits on LOWER block ids, which a 1-D grid
    // dispatches first (the assumption CUB's single-pass scan makes too).  (An atomic ticket
    // instead was measured: no faster, and 3,907 same-address atomics per launch.)
    const int bid = (int)blockIdx.x;
    const long long n = par.n;
    const long long block0 = (long long)bid * (STEP_THREADS * STEP_ITEMS);
    const int M = sc.mesh_count;
    const unsigned below = (1u << lane) - 1u;

    StepItem it[STEP_ITEMS];
    double lp = 0.0, mp = 0.0;
    unsigned n_term = 0;

    // ray state of every item first: STEP_ITEMS independent 2 x 16-byte loads in flight
    float4 o4[STEP_ITEMS], d4[STEP_ITEMS];
    long long ph[STEP_ITEMS];
#pragma unroll
    for (int k = 0; k < STEP_ITEMS; ++k) {
        const long long i = block0 + (long long)k * STEP_THREADS + tid;
        ph[k] = par.phys(i < n ? i : n - 1);
        o4[k] = par.org_pow[ph[k]];
        d4[k] = par.dir_prev[ph[k]];
        it[k].lin = par.lineage ? par.lineage[ph[k]] : 0ull;
    }

#pragma unroll
    for (int k = 0; k < STEP_ITEMS; ++k) {
        const long long i = block0 + (long long)k * STEP_THREADS + tid;
        const bool valid = i < n;
        bool keep_r = false, keep_t = false, is_meas = false, is_term = false;
        it[k].end = it[k].rdir = it[k].tdir = v3(0.f, 0.f, 0.f);
        it[k].rpow = it[k].tpow = it[k].ppow = 0.f;
        it[k].hit_mesh = -1;
        if (valid) {
            const V3 o = v3(o4[k].x, o4[k].y, o4[k].z), d = v3(d4[k].x, d4[k].y, d4[k].z);
            const int prev = __float_as_int(d4[k].w);
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
                    unsigned long long kk = key[(long long)j * n + i];
                    t = key_t(kk);
                    tri = key_tri(kk);
                    h = cnt[(long long)j * n + i];
                };
                r = resolve_media<0>(M, prev, sc.mat_type, max_ray_len, fetch);
            }
            const V3 end = ray_end(o, d, r.t);
            auto tri = [&](int t, V3 &a, V3 &b, V3 &c) {
                float4 fa = __ldg(sc.v0 + t), fb = __ldg(sc.v1 + t), fc = __ldg(sc.v2 + t);
                a = v3(fa.x, fa.y, fa.z);
                b = v3(fb.x, fb.y, fb.z);
                c = v3(fc.x, fc.y, fc.z);
            };
            const Scatter s = scatter(o, end, d, o4[k].w, 0, r.n1, r.n2, r.hit_mesh, r.hit_tri, sc.mat_type, sc.ior,
                                      sc.refl, sc.diss, ior_env, tri);
            dest_meas[i] = make_float4(end.x, end.y, end.z, __int_as_float(s.measured));
            if (hit_out) hit_out[i] = make_int4(r.hit_mesh, r.hit_tri, r.n1, r.n2);
            if (s.power_changed) par.org_pow[ph[k]].w = s.power; // cl:392
            keep_r = s.refl.measured == 0;
            keep_t = s.refr.measured == 0;
            is_meas = s.measured == 1;
            is_term = s.measured == -1;
            it[k].end = end;
            it[k].rdir = s.refl.dir;
            it[k].tdir = s.refr.dir;
            it[k].rpow = s.refl.power;
            it[k].tpow = s.refr.power;
            it[k].ppow = s.power;
            it[k].hit_mesh = r.hit_mesh;
            lp += (keep_r ? (double)s.refl.power : 0.0) + (keep_t ? (double)s.refr.power : 0.0);
            mp += is_meas ? (double)s.power : 0.0;
        }
        // ranks inside this (item, warp) slot
        const unsigned br = __ballot_sync(0xffffffffu, keep_r), bt = __ballot_sync(0xffffffffu, keep_t),
                       bm = __ballot_sync(0xffffffffu, is_meas), bx = __ballot_sync(0xffffffffu, is_term);
        it[k].bits = (keep_r ? 1u : 0u) | (keep_t ? 2u : 0u) | (is_meas ? 4u : 0u) | (__popc(br & below) << 8) |
                     (__popc(bt & below) << 16) | (__popc(bm & below) << 24);
        if (lane == 0) {
            s_cnt[0][k * NW + warp] = __popc(br);
            s_cnt[1][k * NW + warp] = __popc(bt);
            s_cnt[2][k * NW + warp] = __popc(bm);
            n_term += __popc(bx);
        }
    }
    for (int o = 16; o; o >>= 1) {
        lp += __shfl_xor_sync(0xffffffffu, lp, o);
        mp += __shfl_xor_sync(0xffffffffu, mp, o);
    }
    if (lane == 0) {
        s_pow[0][warp] = lp;
        s_pow[1][warp] = mp;
        if (n_term) atomicAdd((unsigned long long *)&ctr->n_terminated, (unsigned long long)n_term);
    }
    __syncthreads();

    if (warp == 0) {
        // exclusive scan of the SLOTS (item, warp) totals, then one look-back for the whole block
        unsigned v[3], inc[3];
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            v[f] = lane < SLOTS ? s_cnt[f][lane] : 0u;
            inc[f] = v[f];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, inc[f], o);
                if (lane >= o) inc[f] += t;
            }
            if (lane < SLOTS) s_off[f][lane] = inc[f] - v[f];
        }
        const unsigned tot_r = __shfl_sync(0xffffffffu, inc[0], 31), tot_t = __shfl_sync(0xffffffffu, inc[1], 31),
                       tot_m = __shfl_sync(0xffffffffu, inc[2], 31);
        unsigned long long ea, eb;
        chained_scan(words, bid, lane, ((unsigned long long)tot_r << 31) | tot_t, tot_m, ea, eb);
        if (lane == 0) {
            s_excl[0] = ea;
            s_excl[1] = eb;
            double blp = 0.0, bmp = 0.0;
            for (int w = 0; w < NW; ++w) {
                blp += s_pow[0][w];
                bmp += s_pow[1][w];
            }
            if (blp != 0.0) atomicAdd(&ctr->live_power, blp);
            if (bmp != 0.0) atomicAdd(&ctr->measured_power, bmp);
            if (block0 + (long long)STEP_THREADS * STEP_ITEMS >= n) { // last block: inclusive prefixes are the totals
                ctr->n_reflect = (long long)(ea >> 31) + tot_r;
                ctr->n_refract = (long long)(ea & 0x7fffffffull) + tot_t;
                ctr->n_measured = (long long)eb + tot_m;
            }
        }
    }
    __syncthreads();
    const long long base_r = (long long)(s_excl[0] >> 31), base_t = (long long)(s_excl[0] & 0x7fffffffull);
    const long long base_m = (long long)s_excl[1];

#pragma unroll
    for (int k = 0; k < STEP_ITEMS; ++k) {
        const unsigned bits = it[k].bits;
        const int slot = k * NW + warp;
        if (bits & 1u) { // reflected survivors: first segment of the next generation
            const long long c = base_r + s_off[0][slot] + ((bits >> 8) & 0xffu);
            chi.org_pow[c] = make_float4(it[k].end.x, it[k].end.y, it[k].end.z, it[k].rpow);
            chi.dir_prev[c] = make_float4(it[k].rdir.x, it[k].rdir.y, it[k].rdir.z, __int_as_float(it[k].hit_mesh));
            if (chi.lineage) chi.lineage[c] = it[k].lin;
        }
        if (bits & 2u) { // refracted survivors: second segment, starting at physical index chi.base (= n for a whole generation)
            const long long c = chi.base + base_t + s_off[1][slot] + ((bits >> 16) & 0xffu);
            chi.org_pow[c] = make_float4(it[k].end.x, it[k].end.y, it[k].end.z, it[k].tpow);
            chi.dir_prev[c] = make_float4(it[k].tdir.x, it[k].tdir.y, it[k].tdir.z, __int_as_float(it[k].hit_mesh));
            if (chi.lineage) chi.lineage[c] = (iteration < 31) ? (it[k].lin | (1ull << (32 + iteration))) : it[k].lin;
        }
        if ((bits & 4u) && meas_out) {
            const long long c = meas_base + base_m + s_off[2][slot] + ((bits >> 24) & 0xffu);
            if (c < meas_cap) {
                meas_out[c] = make_float4(it[k].end.x, it[k].end.y, it[k].end.z, it[k].ppow);
                if (meas_lineage) meas_lineage[c] = it[k].lin;
            }
        }
    }
}

} // namespace lpcl
