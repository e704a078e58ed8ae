/* Compiles the UNMODIFIED reference kernel file by path (macro LPCL_REF_CL,
 * set by oracle/Makefile to /root/reference/kernel_reflect_refract_intersect.cl)
 * into oracle/_ref/ and adds per-kernel "NDRange" launchers.  No reference
 * source is copied into this repository; the only output is a git-ignored .so.
 * TEST INFRASTRUCTURE (checker + CPU baseline), never part of the product path.
 *
 * Launch semantics mirror iterative_tracer.py:288/303/318 of the reference:
 * global size (n,), one work-item per ray, no local size.
 */
#include "cl_shim.h"

thread_local int lpcl_shim_gid = 0;

#ifndef LPCL_REF_CL
#error "build with -DLPCL_REF_CL=\"/abs/path/kernel_reflect_refract_intersect.cl\""
#endif
/* exp(float) inside the kernels (cl:391) is pinned as the correctly rounded float exponential:
 * double exp rounded to float (see cl_shim.h).  The macro only affects the included .cl text. */
static inline float lpcl_shim_exp(float x) { return (float)::exp((double)x); }
#define exp(x) lpcl_shim_exp(x)
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Wunused-parameter"
#pragma GCC diagnostic ignored "-Wunused-variable"
#pragma GCC diagnostic ignored "-Wunused-value"
#include LPCL_REF_CL
#pragma GCC diagnostic pop
#undef exp

extern "C" {

/* reference .cl:243 `intersect`, launched at iterative_tracer.py:288 */
void ref_launch_intersect(int n, const float3 *rays_origin, const float3 *rays_dir, float3 *rays_dest,
                          int *ray_entering, int *ray_isect_mesh_id, int *ray_isect_mesh_idx,
                          const float3 *mesh_v0, const float3 *mesh_v1, const float3 *mesh_v2, const int *mesh_id,
                          float *isect_min_ray_len, int *isects_count, int *ray_isect_mesh_idx_tmp,
                          int mesh_count, int tri_count, int ray_count, float max_ray_len)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (int rid = 0; rid < n; ++rid) {
        lpcl_shim_gid = rid;
        intersect(rays_origin, rays_dir, rays_dest, ray_entering, ray_isect_mesh_id, ray_isect_mesh_idx,
                  mesh_v0, mesh_v1, mesh_v2, mesh_id, isect_min_ray_len, isects_count,
                  ray_isect_mesh_idx_tmp, mesh_count, tri_count, ray_count, max_ray_len);
    }
}

/* reference .cl:105 `intersect_postproc`, launched at iterative_tracer.py:303 */
void ref_launch_intersect_postproc(int n, const float3 *rays_origin, const float3 *rays_dir, float3 *rays_dest,
                                   int *rays_prev_isect_mesh_id, int *rays_n1_mesh_id, int *rays_n2_mesh_id,
                                   int *ray_entering, int *ray_isect_mesh_id, int *ray_isect_mesh_idx,
                                   const float3 *mesh_v0, const float3 *mesh_v1, const float3 *mesh_v2,
                                   const int *mesh_id, const int *mesh_mat_type, float *isect_min_ray_len,
                                   int *isects_count, int *ray_isect_mesh_idx_tmp, int mesh_count, int ray_count,
                                   float max_ray_len)
{
#pragma omp parallel for schedule(static)
    for (int rid = 0; rid < n; ++rid) {
        lpcl_shim_gid = rid;
        intersect_postproc(rays_origin, rays_dir, rays_dest, rays_prev_isect_mesh_id, rays_n1_mesh_id,
                           rays_n2_mesh_id, ray_entering, ray_isect_mesh_id, ray_isect_mesh_idx, mesh_v0,
                           mesh_v1, mesh_v2, mesh_id, mesh_mat_type, isect_min_ray_len, isects_count,
                           ray_isect_mesh_idx_tmp, mesh_count, ray_count, max_ray_len);
    }
}

/* reference .cl:346 `reflect_refract_rays`, launched at iterative_tracer.py:318 */
void ref_launch_reflect_refract_rays(int n, const float3 *in_rays_origin, const float3 *in_rays_dest,
                                     const float3 *in_rays_dir, float *in_rays_power, int *in_rays_measured,
                                     const int *in_ray_entering, int *rays_n1_mesh_id, int *rays_n2_mesh_id,
                                     float3 *rays_reflect_origin, float3 *rays_reflect_dir,
                                     float *rays_reflect_power, int *rays_reflect_measured,
                                     float3 *rays_refract_origin, float3 *rays_refract_dir,
                                     float *rays_refract_power, int *rays_refract_measured,
                                     int *ray_isect_mesh_id, int *ray_isect_mesh_idx, const float3 *mesh_v0,
                                     const float3 *mesh_v1, const float3 *mesh_v2, const int *mesh_id,
                                     const int *mesh_mat_type, const float *mesh_ior, const float *mesh_refl,
                                     const float *mesh_diss, float IOR_env, int mesh_count, int ray_count,
                                     float max_ray_len)
{
#pragma omp parallel for schedule(static)
    for (int rid = 0; rid < n; ++rid) {
        lpcl_shim_gid = rid;
        reflect_refract_rays(in_rays_origin, in_rays_dest, in_rays_dir, in_rays_power, in_rays_measured,
                             in_ray_entering, rays_n1_mesh_id, rays_n2_mesh_id, rays_reflect_origin,
                             rays_reflect_dir, rays_reflect_power, rays_reflect_measured, rays_refract_origin,
                             rays_refract_dir, rays_refract_power, rays_refract_measured, ray_isect_mesh_id,
                             ray_isect_mesh_idx, mesh_v0, mesh_v1, mesh_v2, mesh_id, mesh_mat_type, mesh_ior,
                             mesh_refl, mesh_diss, IOR_env, mesh_count, ray_count, max_ray_len);
    }
}

/* single ray x single triangle probe of reference .cl:50 `intersect_triangle` (KAT helper) */
int ref_intersect_triangle(const float *O, const float *D, const float *V0, const float *V1, const float *V2,
                           float *t, float *u, float *v)
{
    return intersect_triangle(float3(O[0], O[1], O[2]), float3(D[0], D[1], D[2]), float3(V0[0], V0[1], V0[2]),
                              float3(V1[0], V1[1], V1[2]), float3(V2[0], V2[1], V2[2]), t, u, v);
}

/* thread control for the timing build (torchrun exports OMP_NUM_THREADS=1) */
#ifdef _OPENMP
#include <omp.h>
void ref_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int ref_max_threads(void) { return omp_get_max_threads(); }
#else
void ref_set_threads(int n) { (void)n; }
int ref_max_threads(void) { return 1; }
#endif

int ref_build_flags(void)
{
#ifdef LPCL_REF_FAST
    return 1; /* -O3 + host ISA, contraction allowed: timing build only */
#else
    return 0; /* -O2 -ffp-contract=off: the parity build */
#endif
}

} /* extern "C" */
