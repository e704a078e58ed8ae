/* trace_oracle.c - CPU restatement of LightPyCL's per-iteration trace step.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product
 * (lightpycl_b200/) never does and has no CPU fallback.
 *
 * What it restates (all citations: /root/reference/kernel_reflect_refract_intersect.cl):
 *   orc_ray_triangle        <- intersect_triangle      cl:50-101
 *   orc_intersect           <- __kernel intersect      cl:243-289
 *   orc_postproc            <- __kernel intersect_postproc  cl:105-240
 *   orc_fresnel_split       <- reflect_refract         cl:293-343
 *   orc_reflect_refract     <- __kernel reflect_refract_rays cl:346-474
 *
 * Pinned semantics: IEEE-754 binary32, round-to-nearest,
 * NO fused multiply-add (build with -ffp-contract=off), operations in the
 * reference's source order; unsuffixed literals are double, exactly as in
 * OpenCL C on an fp64 device.  dot = (x*x' + y*y') + z*z'; cross in the
 * textbook component order; length = sqrtf(dot); normalize = v / length.
 *
 * Pinning: validated bit-for-bit against oracle/_ref/libref_strict.so (the
 * UNMODIFIED reference .cl compiled by path through cl_shim.h) by
 * tests/test_oracle_pinning.py and against tests/golden/ fixtures generated
 * from that library.  The reference ships no tests or golden vectors of its own.
 *
 * Layout: every 3-vector lives in a 16-byte slot (x, y, z, pad) like OpenCL's
 * float3 in global memory (reference it:141-149 passes (n,4) float32 arrays).
 */
#include <math.h>
#include <stddef.h>

typedef struct { float x, y, z, w; } vec4f;

static inline vec4f v_make(float x, float y, float z) { vec4f r = {x, y, z, 0.0f}; return r; }
static inline vec4f v_sub(vec4f a, vec4f b) { return v_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline vec4f v_add(vec4f a, vec4f b) { return v_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline vec4f v_scale(vec4f a, float s) { return v_make(a.x * s, a.y * s, a.z * s); }
static inline vec4f v_neg(vec4f a) { return v_make(-a.x, -a.y, -a.z); }
static inline float v_dot(vec4f a, vec4f b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline vec4f v_cross(vec4f a, vec4f b)
{
    return v_make(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float sq(float x) { return x * x; } /* pown(x, 2) */

/* ------------------------------------------------------------------------------------------
 * Ray x triangle, Moeller-Trumbore as the reference evaluates it (cl:50-101).
 * No back-face culling, no sign test on t; |DEN| < 1e-6 is an ABSOLUTE reject (cl:79).
 * Returns 1 and fills t,u,v on a hit; on a miss the outputs written so far stay written.
 */
int orc_ray_triangle(const float *O, const float *D, const float *V0, const float *V1, const float *V2,
                     float *t, float *u, float *v)
{
    const float tiny = 0.000001; /* cl:56: a float constant initialised from a double literal */
    vec4f o = v_make(O[0], O[1], O[2]), d = v_make(D[0], D[1], D[2]);
    vec4f a = v_make(V0[0], V0[1], V0[2]), b = v_make(V1[0], V1[1], V1[2]), c = v_make(V2[0], V2[1], V2[2]);

    vec4f e1 = v_sub(b, a);          /* cl:72 */
    vec4f e2 = v_sub(c, a);          /* cl:73 */
    vec4f p = v_cross(d, e2);        /* cl:75 */
    float den = v_dot(p, e1);        /* cl:76 */
    if (den > -tiny && den < tiny)   /* cl:79 */
        return 0;
    float inv = 1.0 / den;           /* cl:82: double divide rounded to float == float divide */
    vec4f tv = v_sub(o, a);          /* cl:83 */
    *u = v_dot(p, tv) * inv;         /* cl:86 */
    if (*u < 0.0 || *u > 1.0)        /* cl:87 */
        return 0;
    vec4f q = v_cross(tv, e1);       /* cl:90 */
    *v = v_dot(q, d) * inv;          /* cl:93 */
    if (*v < 0.0 || *u + *v > 1.0)   /* cl:94 (u+v is a float add) */
        return 0;
    *t = v_dot(q, e2) * inv;         /* cl:98 */
    return 1;
}

/* ------------------------------------------------------------------------------------------
 * Per ray, per mesh: number of crossings with t > EPSILON (counted even beyond max_ray_len),
 * nearest crossing distance (< max_ray_len, strict) and its GLOBAL triangle index; first index
 * wins ties (cl:277-283).  Results land in slot [ray*mesh_count + mesh]; the reference flushes
 * the running values to "current slot - 1" when the mesh id changes (cl:262-269) and to the
 * current slot after the loop (cl:286-288) - reproduced literally, so mesh ids must be
 * contiguous ascending and every mesh non-empty.
 */
static void intersect_one(int ray, const vec4f *origin, const vec4f *dir, const vec4f *v0, const vec4f *v1,
                          const vec4f *v2, const int *tri_mesh, float *near_t, int *hits, int *near_tri,
                          int mesh_count, int tri_count, float max_ray_len)
{
    const float eps_len = 0.000001 * max_ray_len; /* cl:245: double product rounded to float */
    const vec4f o = origin[ray], d = dir[ray];
    int run_hits = 0, run_tri = -1, slot = 0;
    float run_t = max_ray_len;
    float t, u, v;

    for (int i = 0; i < tri_count; ++i) {
        slot = mesh_count * ray + tri_mesh[i];
        if (i > 0 && tri_mesh[i - 1] != tri_mesh[i]) {
            hits[slot - 1] = run_hits;
            near_t[slot - 1] = run_t;
            near_tri[slot - 1] = run_tri;
            run_hits = 0;
            run_t = max_ray_len;
            run_tri = -1;
        }
        if (orc_ray_triangle(&o.x, &d.x, &v0[i].x, &v1[i].x, &v2[i].x, &t, &u, &v) && t > eps_len) {
            if (t < run_t) {
                run_t = t;
                run_tri = i;
            }
            run_hits += 1;
        }
    }
    hits[slot] = run_hits;
    near_t[slot] = run_t;
    near_tri[slot] = run_tri;
}

void orc_intersect(int n, const float *origin, const float *dir, const float *v0, const float *v1, const float *v2,
                   const int *tri_mesh, float *near_t, int *hits, int *near_tri, int mesh_count, int tri_count,
                   float max_ray_len)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (int r = 0; r < n; ++r)
        intersect_one(r, (const vec4f *)origin, (const vec4f *)dir, (const vec4f *)v0, (const vec4f *)v1,
                      (const vec4f *)v2, tri_mesh, near_t, hits, near_tri, mesh_count, tri_count, max_ray_len);
}

/* ------------------------------------------------------------------------------------------
 * Resolve the per-mesh scratch into: nearest mesh, entering/exiting, the two media the ray
 * travels between (n1 -> n2, -1 = environment) and the end point (cl:105-240).
 *
 * prev_mesh: -2 fresh from a source, -1 "outside everything", >= 0 the PARENT'S HIT-SURFACE mesh
 * (reference it:344,373 - not the medium).
 */
static void postproc_one(int ray, const vec4f *origin, const vec4f *dir, vec4f *dest, const int *prev_mesh,
                         int *n1_out, int *n2_out, int *entering_out, int *hit_mesh_out, int *hit_tri_out,
                         const int *mat_type, const float *near_t, const int *hits, const int *near_tri,
                         int mesh_count, float max_ray_len)
{
    const float eps_len = 0.000001 * max_ray_len; /* cl:112 */
    const float *rt = near_t + (size_t)mesh_count * ray;
    const int *rh = hits + (size_t)mesh_count * ray;
    const int *ri = near_tri + (size_t)mesh_count * ray;

    /* (i) nearest mesh; strict '<' so the lowest mesh id wins ties (cl:127-135) */
    float t_hit = max_ray_len;
    int m_hit = -1, tri_hit = -1;
    for (int j = 0; j < mesh_count; ++j) {
        if (rt[j] < t_hit) {
            t_hit = rt[j];
            m_hit = j;
            tri_hit = ri[j];
        }
    }

    int n1 = -1, n2 = -1;
    if (m_hit >= 0) {
        /* (ii) even number of crossings ahead => currently outside => entering (cl:142) */
        int entering = 1 - (rh[m_hit] % 2);
        entering_out[ray] = entering;

        /* (iii) media ids from the previous-surface id (cl:145-177) */
        int prev = prev_mesh[ray];
        if (entering == 1) {
            n1 = (prev == -2 || prev == -1) ? -1 : prev; /* cl:145-171 */
            n2 = m_hit;
        } else {
            n1 = m_hit;
            n2 = -1; /* fixed up by the window scan below */
        }

        /* (iv) coincident / nested surfaces (cl:187-229).  Only refractive (0) and anisotropic (4)
         * meshes take part (cl:194).  Window A = [t_hit, t_hit+EPS]: the LAST mesh j (>= / <=
         * comparisons let later ids win ties) whose nearest crossing lies in the window and is not
         * closer than the running window maximum.  Window B = (t_hit+EPS, max]: the last EXITING
         * mesh with the smallest crossing. */
        const float t_first = t_hit;
        float t_winA = t_hit, t_winB = max_ray_len;
        int winA = -1, winA_entering = 0, winB = -1;
        for (int j = 0; j < mesh_count; ++j) {
            if (mat_type[j] != 0 && mat_type[j] != 4)
                continue;
            float tj = rt[j];
            int ent_j = 1 - (rh[j] % 2);
            if (tj <= t_first + eps_len && tj >= t_winA) { /* cl:202 */
                t_winA = tj;
                winA = j;
                winA_entering = ent_j;
            }
            if (tj > t_first + eps_len && ent_j == 0 && tj <= t_winB) { /* cl:209 */
                t_winB = tj;
                winB = j;
            }
        }
        if (winA_entering == 1) { /* cl:216-219: step straight into the adjacent mesh */
            t_hit = t_winA;
            n2 = winA;
        } else {                  /* cl:220-229 */
            if (winA >= 0)
                t_hit = t_winA;
            if (winB >= 0)
                n2 = winB;
        }
    }

    n1_out[ray] = n1;                                         /* cl:235 */
    n2_out[ray] = n2;                                         /* cl:236 */
    dest[ray] = v_add(origin[ray], v_scale(dir[ray], t_hit)); /* cl:237 */
    hit_mesh_out[ray] = m_hit;                                /* cl:238: stays the step-(i) mesh */
    hit_tri_out[ray] = tri_hit;                               /* cl:239 */
}

void orc_postproc(int n, const float *origin, const float *dir, float *dest, const int *prev_mesh, int *n1,
                  int *n2, int *entering, int *hit_mesh, int *hit_tri, const int *mat_type, const float *near_t,
                  const int *hits, const int *near_tri, int mesh_count, float max_ray_len)
{
#pragma omp parallel for schedule(static)
    for (int r = 0; r < n; ++r)
        postproc_one(r, (const vec4f *)origin, (const vec4f *)dir, (vec4f *)dest, prev_mesh, n1, n2, entering,
                     hit_mesh, hit_tri, mat_type, near_t, hits, near_tri, mesh_count, max_ray_len);
}

/* ------------------------------------------------------------------------------------------
 * Snell + unpolarised Fresnel split at one surface point (cl:293-343).
 * Children: index 0 = reflected, 1 = refracted.  Returns 1 (nothing written) when the TIR
 * discriminant is NaN, exactly like the reference falls through both `if`s (cl:342).
 */
typedef struct { vec4f origin, dir; float power; int measured; } child_t;

static int fresnel_split(vec4f at, vec4f d, float power, vec4f normal_in, float ior1, float ior2, child_t out[2])
{
    vec4f nrm = normal_in;
    float r = ior1 / ior2;                        /* cl:299 */
    float cos1 = -v_dot(nrm, d);                  /* cl:303 */
    if (cos1 < 0.0) {                             /* cl:304-307: make the normal face the ray */
        nrm = v_neg(normal_in);
        cos1 = -v_dot(nrm, d);
    }
    /* cl:309: evaluated in double (literals 1.0), pown() results are floats; rounded to float */
    float tir = 1.0 - sq(r) * (1.0 - sq(cos1));
    if (tir >= 0.0) {                             /* cl:310-328 */
        float cos2 = sqrtf(tir);
        float rs = sq(fabsf((ior1 * cos1 - ior2 * cos2) / (ior1 * cos1 + ior2 * cos2))); /* cl:313 */
        float rp = sq(fabsf((ior1 * cos2 - ior2 * cos1) / (ior1 * cos2 + ior2 * cos1))); /* cl:314 */
        float p_refl = power * (rs + rp) / 2.0;   /* cl:316 */
        out[0].dir = v_add(d, v_scale(nrm, (float)(2.0 * cos1))); /* cl:318 */
        out[0].origin = at;
        out[0].power = p_refl;
        out[0].measured = 0;
        out[1].dir = v_add(v_scale(d, r), v_scale(nrm, r * cos1 - cos2)); /* cl:323 */
        out[1].origin = at;
        out[1].power = power - p_refl;            /* cl:325 */
        out[1].measured = 0;
        return 0;
    }
    if (tir < 0.0) {                              /* cl:329-341: total internal reflection */
        out[0].dir = v_add(d, v_scale(nrm, (float)(2.0 * cos1)));
        out[0].origin = at;
        out[0].power = power;
        out[0].measured = 0;
        out[1].dir = v_make(0.0f, 0.0f, 0.0f);
        out[1].origin = at;
        out[1].power = 0.0;
        out[1].measured = -1;
        return 0;
    }
    return 1;
}

/* exported single-point probe for known-answer tests */
int orc_fresnel_split(const float *at, const float *d, float power, const float *normal, float ior1, float ior2,
                      float *refl_dir, float *refl_power, int *refl_meas, float *refr_dir, float *refr_power,
                      int *refr_meas)
{
    child_t c[2] = {{{0, 0, 0, 0}, {0, 0, 0, 0}, 0.0f, 0}, {{0, 0, 0, 0}, {0, 0, 0, 0}, 0.0f, 0}};
    int rc = fresnel_split(v_make(at[0], at[1], at[2]), v_make(d[0], d[1], d[2]), power,
                           v_make(normal[0], normal[1], normal[2]), ior1, ior2, c);
    refl_dir[0] = c[0].dir.x; refl_dir[1] = c[0].dir.y; refl_dir[2] = c[0].dir.z;
    refr_dir[0] = c[1].dir.x; refr_dir[1] = c[1].dir.y; refr_dir[2] = c[1].dir.z;
    *refl_power = c[0].power; *refl_meas = c[0].measured;
    *refr_power = c[1].power; *refr_meas = c[1].measured;
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * Material handling for one parent ray (cl:346-474): Beer-Lambert loss in the medium it
 * travelled through, Fresnel split or mirror reflection, terminator / measure flagging.
 * power[] and measured[] of the PARENT are updated in place (cl:392, 467, 470).
 */
static void material_one(int ray, const vec4f *origin, const vec4f *dest, const vec4f *dir, float *power,
                         int *measured, const int *n1_id, const int *n2_id, vec4f *refl_origin, vec4f *refl_dir,
                         float *refl_power, int *refl_meas, vec4f *refr_origin, vec4f *refr_dir, float *refr_power,
                         int *refr_meas, const int *hit_mesh, const int *hit_tri, const vec4f *v0, const vec4f *v1,
                         const vec4f *v2, const int *mat_type, const float *ior, const float *refl,
                         const float *diss, float ior_env)
{
    const float tiny = 0.000001; /* cl:359 */
    const int hm = hit_mesh[ray];
    int material = 2;            /* nothing hit => behave like a terminator (cl:364) */
    float mirror_r = 0.0;
    if (hm >= 0) {
        material = mat_type[hm];
        mirror_r = refl[hm];
    }

    const int m1 = n1_id[ray], m2 = n2_id[ray];
    float ior1 = ior_env, ior2 = ior_env;
    float p = power[ray];
    const vec4f at = dest[ray];

    if (m1 >= 0) {
        ior1 = ior[m1];
        if (mat_type[m1] == 0 && diss[m1] > tiny) { /* cl:389-393 */
            vec4f seg = v_sub(at, origin[ray]);
            float len = sqrtf(v_dot(seg, seg));
            /* exp() of a float: pinned as the CORRECTLY ROUNDED float exponential, evaluated as the
             * double exp rounded to float (OpenCL only bounds exp to 3 ulp; glibc's expf is faithful,
             * not correctly rounded, and differs between its FMA / non-FMA ifunc variants) */
            p = p * (float)exp((double)(-diss[m1] * len));
            power[ray] = p;
        }
    }
    if (m2 >= 0)
        ior2 = ior[m2];

    if (measured[ray] == 0 && hm >= 0 && (material == 0 || material == 1)) { /* cl:408 */
        const int ti = hit_tri[ray];
        vec4f nraw = v_cross(v_sub(v1[ti], v0[ti]), v_sub(v2[ti], v1[ti])); /* cl:413: (v1-v0) x (v2-v1) */
        float nlen = sqrtf(v_dot(nraw, nraw));
        vec4f nrm = v_make(nraw.x / nlen, nraw.y / nlen, nraw.z / nlen);

        child_t c[2] = {{{0, 0, 0, 0}, {0, 0, 0, 0}, 0.0f, 0}, {{0, 0, 0, 0}, {0, 0, 0, 0}, 0.0f, 0}};
        fresnel_split(at, dir[ray], p, nrm, ior1, ior2, c); /* NaN case leaves zeros (reference: garbage) */

        refl_origin[ray] = c[0].origin;
        refl_dir[ray] = c[0].dir;
        refl_meas[ray] = c[0].measured;
        refr_origin[ray] = c[1].origin;
        if (material == 0) {                     /* cl:429-439 */
            refl_power[ray] = c[0].power;
            refr_dir[ray] = c[1].dir;
            refr_power[ray] = c[1].power;
            refr_meas[ray] = c[1].measured;
        } else {                                 /* mirror, cl:440-452 */
            refl_power[ray] = p * mirror_r;
            refr_dir[ray] = v_make(0.0f, 0.0f, 0.0f);
            refr_power[ray] = 0.0;
            refr_meas[ray] = -1;
        }
    } else {                                     /* cl:455-472 */
        refl_origin[ray] = at;
        refl_dir[ray] = v_make(0.0f, 0.0f, 0.0f);
        refl_power[ray] = 0.0;
        refl_meas[ray] = -1;
        refr_origin[ray] = at;
        refr_dir[ray] = v_make(0.0f, 0.0f, 0.0f);
        refr_power[ray] = 0.0;
        refr_meas[ray] = -1;
        if (material == 2 || hm < 0)
            measured[ray] = -1;
        if (material == 3 && hm >= 0)
            measured[ray] = 1;
    }
}

void orc_reflect_refract(int n, const float *origin, const float *dest, const float *dir, float *power,
                         int *measured, const int *n1_id, const int *n2_id, float *refl_origin, float *refl_dir,
                         float *refl_power, int *refl_meas, float *refr_origin, float *refr_dir, float *refr_power,
                         int *refr_meas, const int *hit_mesh, const int *hit_tri, const float *v0, const float *v1,
                         const float *v2, const int *mat_type, const float *ior, const float *refl,
                         const float *diss, float ior_env)
{
#pragma omp parallel for schedule(static)
    for (int r = 0; r < n; ++r)
        material_one(r, (const vec4f *)origin, (const vec4f *)dest, (const vec4f *)dir, power, measured, n1_id,
                     n2_id, (vec4f *)refl_origin, (vec4f *)refl_dir, refl_power, refl_meas, (vec4f *)refr_origin,
                     (vec4f *)refr_dir, refr_power, refr_meas, hit_mesh, hit_tri, (const vec4f *)v0,
                     (const vec4f *)v1, (const vec4f *)v2, mat_type, ior, refl, diss, ior_env);
}

int orc_abi_version(void) { return 1; }

/* thread control for the timing builds (torchrun exports OMP_NUM_THREADS=1) */
#ifdef _OPENMP
#include <omp.h>
void orc_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int orc_max_threads(void) { return omp_get_max_threads(); }
#else
void orc_set_threads(int n) { (void)n; }
int orc_max_threads(void) { return 1; }
#endif
