// lpcl_math.cuh - arithmetic of the LightPyCL trace step with pinned rounding.
//
// Everything that decides a RESULT (hit / no hit, t, dest, media ids, child directions and
// powers) is evaluated with the reference kernel's operation order in IEEE-754 binary32,
// round-to-nearest, with NO fused multiply-add: the __f*_rn / __d*_rn intrinsics are never
// contracted by nvcc.  Sub-expressions that OpenCL C evaluates in double (unsuffixed literals,
// SURVEY.md Q9) are evaluated in double here too.  Citations: /root/reference/
// kernel_reflect_refract_intersect.cl ("cl:").
//
// The header also compiles for the host (g++ -ffp-contract=off) so that the conservative search
// filter can be property-tested against the exact test on the CPU (tests/test_filter_host.py).
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define LPCL_HD __host__ __device__ __forceinline__
#define LPCL_MUL(a, b) __fmul_rn((a), (b))
#define LPCL_ADD(a, b) __fadd_rn((a), (b))
#define LPCL_SUB(a, b) __fsub_rn((a), (b))
#define LPCL_DIV(a, b) __fdiv_rn((a), (b))
#define LPCL_SQRT(a) __fsqrt_rn((a))
#define LPCL_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define LPCL_DMUL(a, b) __dmul_rn((a), (b))
#define LPCL_DSUB(a, b) __dsub_rn((a), (b))
#else
#if defined(__CUDACC__)
#define LPCL_HD __host__ __device__ inline
#else
#define LPCL_HD static inline
#endif
#define LPCL_MUL(a, b) ((float)(a) * (float)(b))
#define LPCL_ADD(a, b) ((float)(a) + (float)(b))
#define LPCL_SUB(a, b) ((float)(a) - (float)(b))
#define LPCL_DIV(a, b) ((float)(a) / (float)(b))
#define LPCL_SQRT(a) sqrtf((a))
#define LPCL_FMA(a, b, c) fmaf((a), (b), (c))
#define LPCL_DMUL(a, b) ((double)(a) * (double)(b))
#define LPCL_DSUB(a, b) ((double)(a) - (double)(b))
#endif

namespace lpcl {

struct V3 {
    float x, y, z;
};

LPCL_HD V3 v3(float x, float y, float z)
{
    V3 r;
    r.x = x;
    r.y = y;
    r.z = z;
    return r;
}
LPCL_HD V3 sub(V3 a, V3 b) { return v3(LPCL_SUB(a.x, b.x), LPCL_SUB(a.y, b.y), LPCL_SUB(a.z, b.z)); }
LPCL_HD V3 add(V3 a, V3 b) { return v3(LPCL_ADD(a.x, b.x), LPCL_ADD(a.y, b.y), LPCL_ADD(a.z, b.z)); }
LPCL_HD V3 scale(V3 a, float s) { return v3(LPCL_MUL(a.x, s), LPCL_MUL(a.y, s), LPCL_MUL(a.z, s)); }
LPCL_HD V3 neg(V3 a) { return v3(-a.x, -a.y, -a.z); }
// dot = (x*x' + y*y') + z*z', three products and two sums, each rounded
LPCL_HD float dot(V3 a, V3 b)
{
    return LPCL_ADD(LPCL_ADD(LPCL_MUL(a.x, b.x), LPCL_MUL(a.y, b.y)), LPCL_MUL(a.z, b.z));
}
LPCL_HD V3 cross(V3 a, V3 b)
{
    return v3(LPCL_SUB(LPCL_MUL(a.y, b.z), LPCL_MUL(a.z, b.y)), LPCL_SUB(LPCL_MUL(a.z, b.x), LPCL_MUL(a.x, b.z)),
              LPCL_SUB(LPCL_MUL(a.x, b.y), LPCL_MUL(a.y, b.x)));
}
LPCL_HD float sq(float x) { return LPCL_MUL(x, x); } // pown(x, 2)

// EPSILON = (float)(0.000001 * (double)max_ray_len)   (cl:112, 245, 357)
LPCL_HD float eps_len(float max_ray_len) { return (float)LPCL_DMUL(0.000001, (double)max_ray_len); }

// ------------------------------------------------------------------------------------------
// Exact ray x triangle test, cl:50-101.  Returns true on a hit and sets t.
LPCL_HD bool ray_triangle_exact(V3 o, V3 d, V3 a, V3 b, V3 c, float &t)
{
    const float tiny = 0.000001f; // (float)1e-6, cl:56
    V3 e1 = sub(b, a);
    V3 e2 = sub(c, a);
    V3 p = cross(d, e2);
    float den = dot(p, e1);
    if (den > -tiny && den < tiny) return false; // cl:79
    float inv = LPCL_DIV(1.0f, den);             // cl:82 (double divide rounded to float == float divide)
    V3 tv = sub(o, a);
    float u = LPCL_MUL(dot(p, tv), inv);
    if (u < 0.0f || u > 1.0f) return false;      // cl:87
    V3 q = cross(tv, e1);
    float v = LPCL_MUL(dot(q, d), inv);
    if (v < 0.0f || LPCL_ADD(u, v) > 1.0f) return false; // cl:94
    t = LPCL_MUL(dot(q, e2), inv);               // cl:98
    return true;
}

// ------------------------------------------------------------------------------------------
// Conservative search filter.
//
// With M = O x D (per ray), and per triangle (E1 = fl(V1-V0), E2 = fl(V2-V0) as the reference
// rounds them):
//   DEN = (D x E2).E1        = D . Nn            Nn  = E2 x E1
//   uN  = (D x E2).(O - V0)  = E2 . M + D . A    A   = V0 x E2
//   vN  = ((O-V0) x E1).D    = E1n . M + D . B   E1n = -E1,  B = E1 x V0
// (scalar triple-product identities; u = uN/DEN, v = vN/DEN in the reference).  The reference
// accepts only if 0 <= u <= 1 and 0 <= v <= 1 (v <= 1 follows from u + v <= 1), i.e. only if
// uN (DEN - uN) >= 0 and vN (DEN - vN) >= 0 - two sign-agnostic products, no division, no
// branch on the sign of DEN.  A triangle is a CANDIDATE when
//   uN (DEN - uN) + m >= 0   and   vN (DEN - vN) + m >= 0,          m = K2_ray * G2_tri.
// The parallelogram spanned by the triangle is tested instead of the triangle itself (twice the
// candidates, still a few per ray per sweep) - that trades three sign tests for two and removes
// every sign transfer from the inner loop.
//
// Margin.  Let eps bound |X_filter - X_reference| for X in {DEN, uN, vN}: the FMA evaluation
// here carries <= 9 u |D| Emax (|O| + |V0|) and the reference's own float evaluation <= 7 u ...,
// u = 2^-24, Emax = max(|E1|,|E2|); with L = |O| + 2 Vmax (Vmax = max vertex norm) and a 1.5x
// allowance take eps = 24 u |D| Emax L.  If the reference accepts (a = uN_ref, d = DEN_ref of
// equal sign, |a| <= |d|(1+3u)) then a_f (d_f - a_f) >= -(4u d^2 + 3 eps |d| + 2 eps^2) with
// |d| <= |D| E12, E12 = |E1||E2|, i.e. >= -u |D|^2 Emax L (76 E12 + 1152 u Emax L), which is
// covered with a factor 2 to spare by the separable product of
//   K2_ray = 2 u |D|^2 L max(1, L / Lcap),   G2_tri = Emax (76 E12 + 1152 u Emax Lcap),  Lcap = 4 Vmax.
// Every candidate is re-evaluated with ray_triangle_exact, so the filter can only cost time,
// never change a result; it must merely never reject a triangle the reference accepts
// (property-tested on the CPU by tests/test_filter_host.py, and by every parity test).  DEN's
// 1e-6 cut-off, u + v <= 1 and t > EPSILON are left to the exact evaluation.
struct FilterRec { // 64 bytes
    float nx, ny, nz, g;       // Nn, G2_tri
    float e2x, e2y, e2z, ax;   // E2, A.x
    float ay, az, e1x, e1y;    // A.yz, E1n.xy
    float e1z, bx, by, bz;     // E1n.z, B
};

LPCL_HD FilterRec make_filter_rec(V3 v0, V3 v1, V3 v2, float vmax)
{
    V3 e1 = sub(v1, v0), e2 = sub(v2, v0); // float-rounded edges, exactly the reference's E1/E2
    double e1x = e1.x, e1y = e1.y, e1z = e1.z, e2x = e2.x, e2y = e2.y, e2z = e2.z;
    double ox = v0.x, oy = v0.y, oz = v0.z;
    FilterRec r;
    r.nx = (float)(e2y * e1z - e2z * e1y);
    r.ny = (float)(e2z * e1x - e2x * e1z);
    r.nz = (float)(e2x * e1y - e2y * e1x);
    r.e2x = e2.x;
    r.e2y = e2.y;
    r.e2z = e2.z;
    r.ax = (float)(oy * e2z - oz * e2y);
    r.ay = (float)(oz * e2x - ox * e2z);
    r.az = (float)(ox * e2y - oy * e2x);
    r.e1x = -e1.x;
    r.e1y = -e1.y;
    r.e1z = -e1.z;
    r.bx = (float)(e1y * oz - e1z * oy);
    r.by = (float)(e1z * ox - e1x * oz);
    r.bz = (float)(e1x * oy - e1y * ox);
    double l1 = sqrt(e1x * e1x + e1y * e1y + e1z * e1z), l2 = sqrt(e2x * e2x + e2y * e2y + e2z * e2z);
    double emax = l1 > l2 ? l1 : l2;
    double g = emax * (76.0 * l1 * l2 + 6.9e-5 * emax * 4.0 * (double)vmax) * 1.001; // 1152 * 2^-24 = 6.87e-5
    r.g = (float)(g > 1e-30 ? g : 1e-30);
    return r;
}

// K2_ray = 2 u |D|^2 L max(1, L / (4 Vmax)), L = |O| + 2 Vmax, rounded up
LPCL_HD float filter_ray_scale(V3 o, V3 d, float vmax)
{
    float d2 = d.x * d.x + d.y * d.y + d.z * d.z;
    float lo = sqrtf(o.x * o.x + o.y * o.y + o.z * o.z);
    float L = lo + 2.0f * vmax;
    float f = L / (4.0f * vmax);
    return 1.2e-7f * d2 * L * (f > 1.0f ? f : 1.0f); // 2 * 2^-24 = 1.19e-7
}
// Ray part of the tensor-core filter's margin (derivation below, "the same filter evaluated by the
// tensor cores"): k = 268 u |D| L with L = |O| + 2 V, V = a bound of the vertex norms of ONE TILE of
// triangles (the derivation only needs |V0| + Emax <= 2 V for the triangle at hand).  The ray keeps |O|
// and 268 u |D|, the tile brings 2 V: one multiply-add when V changes.  A scene with a large far shell
// (a measure dome) does not inflate the margin of its small inner triangles.
struct RayScale {
    float lo;  // |O|
    float du;  // 1.6e-5 |D|    (268 * 2^-24 = 1.597e-5)
};
LPCL_HD RayScale filter_ray_parts(V3 o, V3 d)
{
    RayScale r;
    r.lo = sqrtf(o.x * o.x + o.y * o.y + o.z * o.z);
    r.du = 1.6e-5f * sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
    return r;
}
LPCL_HD float filter_ray_scale_tile(RayScale r, float two_v)
{
    return r.du * (r.lo + two_v);
}

// The filter for one (ray, triangle), operation by operation as the packed device loop evaluates
// each lane (fma.rn.f32x2 == two fmaf).  Used by the host property test and by nothing on the GPU.
LPCL_HD bool filter_candidate(const FilterRec &q, float dx, float dy, float dz, float mx, float my, float mz, float k2)
{
    float den = dx * q.nx;
    den = LPCL_FMA(dy, q.ny, den);
    den = LPCL_FMA(dz, q.nz, den);
    float un = mx * q.e2x;
    un = LPCL_FMA(my, q.e2y, un);
    un = LPCL_FMA(mz, q.e2z, un);
    un = LPCL_FMA(dx, q.ax, un);
    un = LPCL_FMA(dy, q.ay, un);
    un = LPCL_FMA(dz, q.az, un);
    float vn = mx * q.e1x;
    vn = LPCL_FMA(my, q.e1y, vn);
    vn = LPCL_FMA(mz, q.e1z, vn);
    vn = LPCL_FMA(dx, q.bx, vn);
    vn = LPCL_FMA(dy, q.by, vn);
    vn = LPCL_FMA(dz, q.bz, vn);
    const float m = k2 * q.g;
    const float pu = LPCL_FMA(un, LPCL_SUB(den, un), m);
    const float pv = LPCL_FMA(vn, LPCL_SUB(den, vn), m);
    return !(signbit(pu) || signbit(pv));
}

// ------------------------------------------------------------------------------------------
// The same filter evaluated by the tensor cores (lpcl_search_tc.cuh).
//
// DEN, uN and vN are linear in the ray's six numbers (M, D), so a tile of rays x a tile of
// triangles is a GEMM with K = 6: D[ray][3 t + {0,1,2}] = (M, D)_ray . row_{t,{0,1,2}}.  The
// tensor cores multiply TF32 (11-bit significands), so every fp32 operand x is split into
// x = hi + lo (+ <= 2^-22 |x|), hi = tf32(x), lo = tf32(x - hi), and the four partial products
// hi.hi + hi.lo + lo.hi + lo.lo are laid along K (4 x 6 = 24 = three kind::tf32 MMAs of K = 8).
//
// The three rows per triangle are DEN, aU = 2 uN - DEN and aV = 2 vN - DEN (still linear in the
// ray).  The reference accepts only if 0 <= uN/DEN <= 1 and 0 <= vN/DEN <= 1, i.e. only if
// |aU| <= |DEN| and |aV| <= |DEN| whatever the sign of DEN (the parallelogram spanned by the triangle), so
//   candidate  <=>  |DEN| - max(|aU|, |aV|) + k >= 0          (the LINEAR form)
// - per triangle one FMNMX and one FADD (absolute values are SASS operand modifiers), and the chunk-wide
// "any candidate?" is a running maximum, one three-input FMNMX3 per PAIR of triangles: 2.5 issue slots
// per test instead of the 4 of the square form DEN^2 + m - aU^2 >= 0 (three packed multiply-adds and two
// LOP3 per pair, a packed instruction holding the issue port for two cycles), which this kernel used
// before: the CUDA-core epilogue, not the tensor pipe, bounds the kernel.  Every maximum propagates NaN
// (max.NaN), and a NaN passes the final comparison: overflowing inputs become candidates, never misses.
// Each triangle's three rows are pre-scaled by s = 1 / Emax_tri, which turns the margin
// k = 268 u |D| L Emax_tri into the per-RAY constant k_ray = 268 u |D| L.
//
// Margin.  Measured on B200 (tools/ubench_tmem.cu, profiles/r1_ubench_tmem.txt): a kind::tf32
// tcgen05.mma sums its K = 8 products and the accumulator with extra bits and truncates once
// (|err| <= 2.7 u sum|terms| observed per instruction, 6.2 u over three); budget 16 u.  Operand
// splitting costs <= 4 u per operand, M = O x D in fp32 <= 2 u, the scaled coefficient's
// rounding <= 1 u: eps_filter <= (16 + 6 + 5) u |D| Emax (|O| + |V0|), plus the reference's own
// 7 u, times the same 1.5 allowance: eps = 56 u |D| Emax L bounds |DEN_f - DEN_ref|; the aU / aV
// rows carry three times the terms, 3 eps.  If the reference accepts, uN_ref and DEN_ref have equal
// signs and |uN_ref| <= |DEN_ref| (1 + 3u), i.e. |aU_ref| <= |d| (1 + 6u) with d = DEN_ref; then
//   |aU_f| <= |aU_ref| + 3 eps <= (|DEN_f| + eps)(1 + 6u) + 3 eps,
//   |aU_f| - |DEN_f| <= 6u |DEN_f| + (4 + 6u) eps <= (6 + 224 + 1) u |D| Emax L = 231 u |D| Emax L
// (|DEN_f| <= |D| E12 + eps and E12 <= Emax . 2V <= Emax L were used), and the same for aV.  The epilogue's own
// subtraction and the addition of k round by <= u relative each.  k = 268 u |D| Emax L covers that with 16 % to
// spare.  In the units of the square form (aU^2 - DEN^2 <= m means |aU| - |DEN| <= m / (|aU| + |DEN|) ~ 463 .. 926 u
// |D| Emax L / cos(incidence)) the linear margin is 2 - 4 x narrower: fewer false candidates as well.
LPCL_HD float tf32_rna(float x) // round to nearest (ties away) to 10 explicit mantissa bits, as cvt.rna.tf32.f32
{
    union { float f; uint32_t u; } v;
    v.f = x;
    v.u = (v.u + 0x1000u) & 0xffffe000u;
    return v.f;
}
LPCL_HD void tf32_split(float x, float &hi, float &lo)
{
    hi = tf32_rna(x);
    lo = tf32_rna(x - hi); // x - hi is exact in binary32
}

struct TcRows {
    float c[3][6]; // rows DEN', aU', aV' over (Mx, My, Mz, Dx, Dy, Dz)
};

LPCL_HD TcRows make_tc_rows(V3 v0, V3 v1, V3 v2, float vmax)
{
    V3 e1 = sub(v1, v0), e2 = sub(v2, v0);
    double e1x = e1.x, e1y = e1.y, e1z = e1.z, e2x = e2.x, e2y = e2.y, e2z = e2.z;
    double ox = v0.x, oy = v0.y, oz = v0.z;
    double l1 = sqrt(e1x * e1x + e1y * e1y + e1z * e1z), l2 = sqrt(e2x * e2x + e2y * e2y + e2z * e2z);
    double emax = l1 > l2 ? l1 : l2;
    double s = 1.0 / (emax > 1e-30 ? emax : 1e-30); // rows in units of Emax: the margin becomes a per-ray constant
    const double nn[3] = {e2y * e1z - e2z * e1y, e2z * e1x - e2x * e1z, e2x * e1y - e2y * e1x}; // Nn = E2 x E1
    const double a[3] = {oy * e2z - oz * e2y, oz * e2x - ox * e2z, ox * e2y - oy * e2x};         // A = V0 x E2
    const double b[3] = {e1y * oz - e1z * oy, e1z * ox - e1x * oz, e1x * oy - e1y * ox};         // B = E1 x V0
    const double e2d[3] = {e2x, e2y, e2z}, e1d[3] = {e1x, e1y, e1z};
    TcRows r;
    for (int k = 0; k < 3; ++k) {
        r.c[0][k] = 0.0f;                                   // DEN = D . Nn
        r.c[0][3 + k] = (float)(s * nn[k]);
        r.c[1][k] = (float)(s * 2.0 * e2d[k]);              // aU = 2 (E2 . M + D . A) - D . Nn
        r.c[1][3 + k] = (float)(s * (2.0 * a[k] - nn[k]));
        r.c[2][k] = (float)(-s * 2.0 * e1d[k]);             // aV = 2 (-E1 . M + D . B) - D . Nn
        r.c[2][3 + k] = (float)(s * (2.0 * b[k] - nn[k]));
    }
    return r;
}

// Host model of the tensor-core filter for the property test: the three GEMM outputs are formed
// exactly from the split operands (double), then pushed by +-16 u sum|terms| in the direction
// `push` selects (bit 0: DEN', bit 1: aU', bit 2: aV'); the caller requires a reference hit to be a
// candidate for every push.  The epilogue's maximum, subtraction and addition are evaluated as the GPU does.
// Not used on the GPU.
LPCL_HD float filter_value_tc(const TcRows &q, float dx, float dy, float dz, float mx, float my, float mz, int push)
{
    const float a[6] = {mx, my, mz, dx, dy, dz};
    float out[3];
    for (int r = 0; r < 3; ++r) {
        double s = 0.0, sabs = 0.0;
        for (int k = 0; k < 6; ++k) {
            float ah, al, bh, bl;
            tf32_split(a[k], ah, al);
            tf32_split(q.c[r][k], bh, bl);
            double p = ((double)ah + (double)al) * ((double)bh + (double)bl);
            s += p;
            sabs += p < 0 ? -p : p;
        }
        const double e = 16.0 * 5.9604644775390625e-8 * sabs;
        out[r] = (float)(s + (((push >> r) & 1) ? e : -e));
    }
    return fabsf(out[0]) - fmaxf(fabsf(out[1]), fabsf(out[2]));
}
LPCL_HD bool filter_candidate_tc(const TcRows &q, float dx, float dy, float dz, float mx, float my, float mz, float kray, int push)
{
    return !(filter_value_tc(q, dx, dy, dz, mx, my, mz, push) + kray < 0.0f);
}

// ------------------------------------------------------------------------------------------
// Per-ray resolution of the per-mesh search results, cl:105-240.
struct Resolve {
    float t;                 // distance to the end point
    int hit_mesh, hit_tri;   // nearest mesh / its triangle (step (i)); -1 = none
    int n1, n2, entering;
};

// `fetch(j, t_j, hits_j, tri_j)` returns mesh j's nearest distance, crossing count and triangle.
// UNROLL > 0: mesh_count <= UNROLL is guaranteed by the caller and both loops are fully unrolled
// (the per-mesh values then live in registers); UNROLL == 0: plain run-time loops.
template <int UNROLL = 0, class Fetch>
LPCL_HD Resolve resolve_media(int mesh_count, int prev_mesh, const int *mat_type, float max_ray_len, Fetch fetch)
{
    const float eps = eps_len(max_ray_len);
    Resolve r;
    r.t = max_ray_len;
    r.hit_mesh = -1;
    r.hit_tri = -1;
    r.n1 = -1;
    r.n2 = -1;
    r.entering = 0;
    int hits_hit = 0;
    const int bound = UNROLL > 0 ? UNROLL : mesh_count;
#pragma unroll
    for (int j = 0; j < bound; ++j) { // cl:127-135, strict '<': lowest mesh id wins ties
        if (UNROLL > 0 && j >= mesh_count) break;
        float tj;
        int hj, ij;
        fetch(j, tj, hj, ij);
        if (tj < r.t) {
            r.t = tj;
            r.hit_mesh = j;
            r.hit_tri = ij;
            hits_hit = hj;
        }
    }
    if (r.hit_mesh < 0) return r;

    r.entering = 1 - (hits_hit % 2); // cl:142
    if (r.entering == 1) {           // cl:145-177
        r.n1 = (prev_mesh == -2 || prev_mesh == -1) ? -1 : prev_mesh;
        r.n2 = r.hit_mesh;
    } else {
        r.n1 = r.hit_mesh;
        r.n2 = -1;
    }
    // cl:187-229: coincident / nested refractive surfaces
    const float t_first = r.t;
    const float t_edge = LPCL_ADD(t_first, eps);
    float t_a = r.t, t_b = max_ray_len;
    int win_a = -1, win_a_entering = 0, win_b = -1;
#pragma unroll
    for (int j = 0; j < bound; ++j) {
        if (UNROLL > 0 && j >= mesh_count) break;
        int mt = mat_type[j];
        if (mt != 0 && mt != 4) continue; // cl:194
        float tj;
        int hj, ij;
        fetch(j, tj, hj, ij);
        int ent = 1 - (hj % 2);
        if (tj <= t_edge && tj >= t_a) { // cl:202
            t_a = tj;
            win_a = j;
            win_a_entering = ent;
        }
        if (tj > t_edge && ent == 0 && tj <= t_b) { // cl:209
            t_b = tj;
            win_b = j;
        }
    }
    if (win_a_entering == 1) { // cl:216-219
        r.t = t_a;
        r.n2 = win_a;
    } else {                   // cl:220-229
        if (win_a >= 0) r.t = t_a;
        if (win_b >= 0) r.n2 = win_b;
    }
    return r;
}

LPCL_HD V3 ray_end(V3 o, V3 d, float t) { return add(o, scale(d, t)); } // cl:237

// ------------------------------------------------------------------------------------------
// Snell + unpolarised Fresnel at one point, cl:293-343.  child 0 = reflected, 1 = refracted.
struct Child {
    V3 dir;
    float power;
    int measured; // 0 alive, -1 dead
};

// returns false when the TIR discriminant is NaN (reference returns 1 and writes nothing, cl:342)
LPCL_HD bool fresnel_split(V3 d, float power, V3 normal_in, float ior1, float ior2, Child &refl, Child &refr)
{
    V3 nrm = normal_in;
    float r = LPCL_DIV(ior1, ior2);
    float cos1 = -dot(nrm, d);
    if (cos1 < 0.0f) { // cl:304-307
        nrm = neg(normal_in);
        cos1 = -dot(nrm, d);
    }
    // cl:309, double: 1.0 - pown(r,2) * (1.0 - pown(cosT1,2)), rounded to float
    float tir = (float)LPCL_DSUB(1.0, LPCL_DMUL((double)sq(r), LPCL_DSUB(1.0, (double)sq(cos1))));
    float two_cos1 = LPCL_MUL(2.0f, cos1); // (float)(2.0*cosT1), exact
    if (tir >= 0.0f) {
        float cos2 = LPCL_SQRT(tir);
        float a1 = LPCL_MUL(ior1, cos1), b2 = LPCL_MUL(ior2, cos2);
        float rs = sq(fabsf(LPCL_DIV(LPCL_SUB(a1, b2), LPCL_ADD(a1, b2)))); // cl:313
        float a2 = LPCL_MUL(ior1, cos2), b1 = LPCL_MUL(ior2, cos1);
        float rp = sq(fabsf(LPCL_DIV(LPCL_SUB(a2, b1), LPCL_ADD(a2, b1)))); // cl:314
        float p_refl = LPCL_MUL(LPCL_MUL(power, LPCL_ADD(rs, rp)), 0.5f);   // cl:316 (/2.0 is exact)
        refl.dir = add(d, scale(nrm, two_cos1));                           // cl:318
        refl.power = p_refl;
        refl.measured = 0;
        refr.dir = add(scale(d, r), scale(nrm, LPCL_SUB(LPCL_MUL(r, cos1), cos2))); // cl:323
        refr.power = LPCL_SUB(power, p_refl);
        refr.measured = 0;
        return true;
    }
    if (tir < 0.0f) { // cl:329-341
        refl.dir = add(d, scale(nrm, two_cos1));
        refl.power = power;
        refl.measured = 0;
        refr.dir = v3(0.0f, 0.0f, 0.0f);
        refr.power = 0.0f;
        refr.measured = -1;
        return true;
    }
    return false;
}

// Material handling of one parent, cl:346-474.
struct Scatter {
    float power;       // parent power after dissipation
    bool power_changed;
    int measured;      // parent flag after the step
    Child refl, refr;  // children; origin of both = dest
};

LPCL_HD float exp_neg(float x)
{
    // cl:391 exp() of a float.  OpenCL bounds it to 3 ulp only; the oracle pins it as the CORRECTLY
    // ROUNDED float exponential = double exp rounded to float (oracle/cl_shim.h).  CUDA's double
    // exp (<= 1 ulp in double) gives the same float outside ~2^-28 of the inputs.
    return (float)exp((double)x);
}

template <class TriFetch>
LPCL_HD Scatter scatter(V3 origin, V3 dest, V3 dir, float power, int measured, int n1, int n2, int hit_mesh,
                        int hit_tri, const int *mat_type, const float *ior, const float *refl_coef,
                        const float *diss, float ior_env, TriFetch tri)
{
    const float tiny = 0.000001f; // cl:359
    Scatter s;
    int material = 2; // nothing hit => terminate (cl:364)
    float mirror_r = 0.0f;
    if (hit_mesh >= 0) {
        material = mat_type[hit_mesh];
        mirror_r = refl_coef[hit_mesh];
    }
    float ior1 = ior_env, ior2 = ior_env;
    float p = power;
    s.power_changed = false;
    if (n1 >= 0) {
        ior1 = ior[n1];
        float dz = diss[n1];
        if (mat_type[n1] == 0 && dz > tiny) { // cl:389-393
            V3 seg = sub(dest, origin);
            float len = LPCL_SQRT(dot(seg, seg));
            p = LPCL_MUL(p, exp_neg(LPCL_MUL(-dz, len)));
            s.power_changed = true;
        }
    }
    if (n2 >= 0) ior2 = ior[n2];
    s.power = p;
    s.measured = measured;

    if (measured == 0 && hit_mesh >= 0 && (material == 0 || material == 1)) { // cl:408
        V3 a, b, c;
        tri(hit_tri, a, b, c);
        V3 nraw = cross(sub(b, a), sub(c, b)); // cl:413: (v1-v0) x (v2-v1)
        float nlen = LPCL_SQRT(dot(nraw, nraw));
        V3 nrm = v3(LPCL_DIV(nraw.x, nlen), LPCL_DIV(nraw.y, nlen), LPCL_DIV(nraw.z, nlen));
        s.refl.dir = v3(0.0f, 0.0f, 0.0f);
        s.refl.power = 0.0f;
        s.refl.measured = 0;
        s.refr = s.refl;
        fresnel_split(dir, p, nrm, ior1, ior2, s.refl, s.refr);
        if (material == 1) { // mirror, cl:440-452
            s.refl.power = LPCL_MUL(p, mirror_r);
            s.refr.dir = v3(0.0f, 0.0f, 0.0f);
            s.refr.power = 0.0f;
            s.refr.measured = -1;
        }
    } else { // cl:455-472
        s.refl.dir = v3(0.0f, 0.0f, 0.0f);
        s.refl.power = 0.0f;
        s.refl.measured = -1;
        s.refr = s.refl;
        if (material == 2 || hit_mesh < 0) s.measured = -1;
        if (material == 3 && hit_mesh >= 0) s.measured = 1;
    }
    return s;
}

} // namespace lpcl
