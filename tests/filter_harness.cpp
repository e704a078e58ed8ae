// Host property test of the conservative search filter: for adversarial random (ray, triangle)
// pairs, whenever the reference-ordered exact test (lpcl::ray_triangle_exact, cl:50-101) reports a
// hit, the FMA filter (lpcl::filter_candidate, lane-for-lane what the packed CUDA loop computes)
// must flag the pair as a candidate.  Build: g++ -O2 -ffp-contract=off -shared -fPIC.
#include <stdint.h>
#include <math.h>
#include "../lightpycl_b200/csrc/lpcl_math.cuh"

using namespace lpcl;

static uint64_t s_state;
static inline double rnd() { // xorshift64*
    s_state ^= s_state >> 12; s_state ^= s_state << 25; s_state ^= s_state >> 27;
    return (double)((s_state * 2685821237468027973ULL) >> 11) / 9007199254740992.0;
}
static inline double srnd() { return 2.0 * rnd() - 1.0; }

static long property(int tc, long n, unsigned seed, double scene_radius, double tri_size, double origin_radius,
                     int unit_dirs, long *n_exact, long *n_cand, double *worst)
{
    s_state = 0x9E3779B97F4A7C15ULL ^ ((uint64_t)seed << 17);
    long violations = 0;
    *n_exact = 0; *n_cand = 0; *worst = 0.0;
    const float vmax = (float)((scene_radius + 2.0 * tri_size) * 1.7320508 * 1.0001);
    for (long it = 0; it < n; ++it) {
        // triangle somewhere in the scene cube, edges of size ~tri_size (log-uniform over 3 decades)
        double sz = tri_size * pow(10.0, -3.0 * rnd());
        V3 a = v3((float)(scene_radius * srnd()), (float)(scene_radius * srnd()), (float)(scene_radius * srnd()));
        V3 b = v3((float)(a.x + sz * srnd()), (float)(a.y + sz * srnd()), (float)(a.z + sz * srnd()));
        V3 c = v3((float)(a.x + sz * srnd()), (float)(a.y + sz * srnd()), (float)(a.z + sz * srnd()));
        if (rnd() < 0.05) c = b;                       // degenerate triangles
        // ray origin; aim at a point on / just off the triangle (edges and corners over-weighted)
        V3 o = v3((float)(origin_radius * srnd()), (float)(origin_radius * srnd()), (float)(origin_radius * srnd()));
        double u = rnd(), v = rnd();
        int mode = (int)(rnd() * 6);
        if (mode == 0) u = 0; else if (mode == 1) v = 0; else if (mode == 2) { v = 1 - u; } else if (mode == 3) { u = 0; v = 0; }
        if (u + v > 1 && mode > 3) { u = 1 - u; v = 1 - v; }
        double jit = (rnd() < 0.5) ? 0.0 : 1e-6 * srnd();
        u += jit; v += jit * srnd();
        double px = a.x + u * ((double)b.x - a.x) + v * ((double)c.x - a.x);
        double py = a.y + u * ((double)b.y - a.y) + v * ((double)c.y - a.y);
        double pz = a.z + u * ((double)b.z - a.z) + v * ((double)c.z - a.z);
        double dxd = px - o.x, dyd = py - o.y, dzd = pz - o.z;
        double len = sqrt(dxd * dxd + dyd * dyd + dzd * dzd);
        if (len == 0) continue;
        double sc = unit_dirs ? 1.0 / len : pow(10.0, 2.0 * srnd()) / len;
        if (rnd() < 0.3) sc = -sc;                     // hits behind the origin count for the parity too
        V3 d = v3((float)(dxd * sc), (float)(dyd * sc), (float)(dzd * sc));
        float t;
        bool hit = ray_triangle_exact(o, d, a, b, c, t);
        float mx = o.y * d.z - o.z * d.y, my = o.z * d.x - o.x * d.z, mz = o.x * d.y - o.y * d.x;
        float k2 = filter_ray_scale(o, d, vmax);
        bool cand;
        if (tc) { // tensor-core model: a candidate under every admissible rounding of the three GEMM outputs
            // the tile-local vertex-norm bound: anything from this triangle's own largest vertex norm upwards
            float na = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z), nb = sqrtf(b.x * b.x + b.y * b.y + b.z * b.z), nc = sqrtf(c.x * c.x + c.y * c.y + c.z * c.z);
            float vloc = fmaxf(fmaxf(na, nb), fmaxf(nc, 1e-30f)) * 1.0001f * (float)(1.0 + (rnd() < 0.5 ? 0.0 : 3.0 * rnd()));
            k2 = filter_ray_scale_tile(filter_ray_parts(o, d), 2.0f * vloc);
            TcRows q = make_tc_rows(a, b, c, vloc);
            cand = true;
            for (int push = 0; push < 8; ++push) {
                cand = cand && filter_candidate_tc(q, d.x, d.y, d.z, mx, my, mz, k2, push);
                // the share of the margin a reference hit actually needed (1 = the filter would have dropped it)
                if (hit) *worst = fmax(*worst, -(double)filter_value_tc(q, d.x, d.y, d.z, mx, my, mz, push) / (double)k2);
            }
        } else {
            FilterRec q = make_filter_rec(a, b, c, vmax);
            cand = filter_candidate(q, d.x, d.y, d.z, mx, my, mz, k2);
        }
        *n_exact += hit;
        *n_cand += cand;
        if (hit && !cand) ++violations;
    }
    return violations;
}

extern "C" long filter_property(long n, unsigned seed, double scene_radius, double tri_size, double origin_radius,
                                int unit_dirs, long *n_exact, long *n_cand, double *worst)
{
    return property(0, n, seed, scene_radius, tri_size, origin_radius, unit_dirs, n_exact, n_cand, worst);
}

extern "C" long filter_property_tc(long n, unsigned seed, double scene_radius, double tri_size, double origin_radius,
                                   int unit_dirs, long *n_exact, long *n_cand, double *worst)
{
    return property(1, n, seed, scene_radius, tri_size, origin_radius, unit_dirs, n_exact, n_cand, worst);
}
