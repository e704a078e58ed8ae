/* The reference's measured-ray projection kernels - `rotate_vec` (.cl:476-484), `stereograph_project`
 * (.cl:488-506) and `angular_project` (.cl:509-538) - compiled FROM THE REFERENCE SOURCE as a CPU
 * checker.  TEST INFRASTRUCTURE (oracle/_ref/libref_proj.so), never part of the product path.
 *
 * These three kernels build float3 values with OpenCL's vector-literal syntax `(float3)(u,v,w)`,
 * which C++ parses as a cast of the comma expression (= w).  oracle/Makefile therefore pipes the
 * reference file through `sed 's/(float3)(/float3(/g'` (vector literal -> 3-argument constructor,
 * nothing else changes) straight into the compiler: this translation unit includes the text from
 * /dev/stdin (macro LPCL_REF_TEXT).  No copy of the reference source is stored anywhere; the only
 * output is the git-ignored .so.  The hot-path oracle (libref_strict.so) keeps compiling the
 * UNMODIFIED file (ref_shim.cpp); this library exports only the projection launchers.
 *
 * Semantics pinned by cl_shim.h (strict IEEE, no contraction, unsuffixed literals double);
 * sqrt/atan2/acos/cos/sin of a float are the C++ float overloads (glibc sqrtf, atan2f, acosf, cosf,
 * sinf) - OpenCL bounds them to 4-5 ulp only, so parity with them is a tolerance, not bit-exact.
 */
#include "cl_shim.h"

thread_local int lpcl_shim_gid = 0;

#ifndef LPCL_REF_TEXT
#error "build through oracle/Makefile: sed ... reference.cl | g++ -DLPCL_REF_TEXT='\"/dev/stdin\"' proj_shim.cpp"
#endif
static inline float lpcl_shim_exp(float x) { return (float)::exp((double)x); }
#define exp(x) lpcl_shim_exp(x)
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Wunused-parameter"
#pragma GCC diagnostic ignored "-Wunused-variable"
#pragma GCC diagnostic ignored "-Wunused-value"
#include LPCL_REF_TEXT
#pragma GCC diagnostic pop
#undef exp

extern "C" {

/* reference .cl:488 `stereograph_project`, launched at iterative_tracer.py:515 with global size pwr0.shape */
void ref_launch_stereograph_project(int n, const float3 *vecs, const float *pwrs, const float3 *rot_mtx, float3 *pivot,
                                    float *x, float *y, float *pwrs_cor)
{
#pragma omp parallel for schedule(static)
    for (int gid = 0; gid < n; ++gid) {
        lpcl_shim_gid = gid;
        stereograph_project(vecs, pwrs, rot_mtx, pivot, x, y, pwrs_cor);
    }
}

/* reference .cl:509 `angular_project`, launched at iterative_tracer.py:546 and 602 */
void ref_launch_angular_project(int n, const float3 *vecs, const float *pwrs, const float3 *rot_mtx, float3 *pivot, float *x,
                                float *y, float *pwrs_cor)
{
#pragma omp parallel for schedule(static)
    for (int gid = 0; gid < n; ++gid) {
        lpcl_shim_gid = gid;
        angular_project(vecs, pwrs, rot_mtx, pivot, x, y, pwrs_cor);
    }
}

/* reference .cl:476 `rotate_vec` */
void ref_launch_rotate_vec(int n, const float3 *vecs, const float3 *rot_mtx, float3 *pivot, float3 *vecs_rot)
{
#pragma omp parallel for schedule(static)
    for (int gid = 0; gid < n; ++gid) {
        lpcl_shim_gid = gid;
        rotate_vec(vecs, rot_mtx, pivot, vecs_rot);
    }
}

} /* extern "C" */
