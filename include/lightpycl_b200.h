/* lightpycl_b200.h - C ABI of the B200-native LightPyCL trace step.
 *
 * One shared library (lightpycl_b200/_lib/liblightpycl_b200.so, sm_100a) replaces the three
 * OpenCL kernels of the reference's hot path and the device side of its host loop:
 *
 *   reference interface (what a binding would have called)                 entry point here
 *   --------------------------------------------------------------------   -----------------------------
 *   cl.Program(ctx, src).build(); to_device(m_v0,m_v1,m_v2,m_id, materials) lpcl_scene_create / _destroy
 *     iterative_tracer.py:66-71, 161-169
 *   prg.intersect(queue,(n,),None, 13 buffers, mesh_count, tri_count,       lpcl_intersect
 *     ray_count, max_ray_len)          iterative_tracer.py:288-294, .cl:243
 *   prg.intersect_postproc(... 17 buffers ...)                              lpcl_intersect_postproc
 *                                      iterative_tracer.py:303-309, .cl:105
 *   prg.reflect_refract_rays(... 26 buffers ...)                            lpcl_reflect_refract_rays
 *                                      iterative_tracer.py:318-325, .cl:346
 *   the three launches + 12 result downloads + NumPy pruning of one         lpcl_trace_step
 *     iteration                        iterative_tracer.py:261-375
 *
 * Conventions
 *   - Every pointer named d_* is a DEVICE pointer owned by the caller (e.g. tensor.data_ptr());
 *     h_* is a HOST pointer.  The library allocates device memory only inside lpcl_scene_create
 *     (filter tiles + material tables), released by lpcl_scene_destroy.  No hidden globals.
 *   - 3-vectors occupy 16 bytes (x, y, z, pad) like OpenCL float3 / the reference's (n,4) float32
 *     arrays; ints are 32-bit.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  All entry points are
 *     asynchronous on that stream and never synchronise; the caller syncs before reading results.
 *   - Return value: 0 = OK; negative = argument error (LPCL_E_*); positive = cudaError_t.
 *     lpcl_last_error() returns a thread-local message for the last failure.  Nothing throws or
 *     exits across this boundary.
 *   - A scene handle is immutable after creation; concurrent calls on one handle are allowed as
 *     long as their caller-owned buffers are distinct.  One handle per GPU.
 *   - Arithmetic contract: IEEE-754 binary32, round-to-nearest, no FMA contraction, operations in
 *     the reference kernel's source order (double where OpenCL C promotes literals), so that all
 *     outputs are bit-identical to the strict-IEEE CPU build of the reference kernels.  The
 *     nearest-hit SEARCH uses a conservative FMA filter and re-evaluates every candidate with the
 *     reference arithmetic; it never changes results.
 */
#ifndef LIGHTPYCL_B200_H
#define LIGHTPYCL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LPCL_ABI_VERSION 1

#define LPCL_E_NULL (-1)      /* required pointer is NULL */
#define LPCL_E_RANGE (-2)     /* negative / zero count, capacity too small */
#define LPCL_E_MESH_IDS (-3)  /* triangle mesh ids not contiguous ascending from 0, or an empty mesh
                                 (the reference kernel silently misbehaves here: .cl:262-265) */
#define LPCL_E_ALIGN (-4)     /* a float4 buffer is not 16-byte aligned */

typedef struct lpcl_scene lpcl_scene;

int lpcl_abi_version(void);
const char *lpcl_last_error(void);

/* ---------------------------------------------------------------- scene
 * d_v0,d_v1,d_v2 : (tri_count,4) float32, caller-owned, must outlive the handle
 *                  (reference m_v0_buf..m_v2_buf, iterative_tracer.py:161-163)
 * h_tri_mesh     : (tri_count,) int32 mesh id per triangle (reference m_id, it:138-149)
 * h_mat_type,h_ior,h_refl,h_diss : (mesh_count,) per-mesh material (it:122-135; codes .cl:364)
 * Builds the per-triangle filter records (64 B each, tiles padded at mesh boundaries) on `stream`.
 */
int lpcl_scene_create(lpcl_scene **out, const float *d_v0, const float *d_v1, const float *d_v2,
                      const int32_t *h_tri_mesh, int32_t tri_count, const int32_t *h_mat_type,
                      const float *h_ior, const float *h_refl, const float *h_diss, int32_t mesh_count,
                      void *stream);
/* Same, but the device memory of the handle (filter tiles, tile table, materials) is a caller-owned
 * buffer of at least lpcl_scene_bytes(tri_count, mesh_count) bytes, 256-byte aligned, that must
 * outlive the handle: nothing is cudaMalloc'ed or cudaFree'd (the allocator stays the caller's). */
int64_t lpcl_scene_bytes(int32_t tri_count, int32_t mesh_count);
int lpcl_scene_create_in(lpcl_scene **out, const float *d_v0, const float *d_v1, const float *d_v2,
                         const int32_t *h_tri_mesh, int32_t tri_count, const int32_t *h_mat_type,
                         const float *h_ior, const float *h_refl, const float *h_diss, int32_t mesh_count,
                         void *d_scene_mem, int64_t scene_mem_bytes, void *stream);
void lpcl_scene_destroy(lpcl_scene *scene);
int32_t lpcl_scene_tri_count(const lpcl_scene *scene);
int32_t lpcl_scene_mesh_count(const lpcl_scene *scene);
int64_t lpcl_scene_device_bytes(const lpcl_scene *scene); /* memory owned by the handle */

/* Bytes of caller-owned device scratch the search needs for n rays: per (mesh, ray) one packed
 * 64-bit (nearest t, triangle) key and one 32-bit crossing count, mesh-major; plus the record buffer through which
 * the default (deferred-flag) search hands flagged (ray, 16-triangle chunk) pairs to its second pass: 8 bytes per
 * record, LPCL_REC_PER_RAY records per ray (environment, read once per process; default 32, i.e. 256 bytes per ray;
 * records that do not fit are evaluated inside the sweep kernel - slower, same results). */
int64_t lpcl_scratch_bytes(const lpcl_scene *scene, int64_t n_rays);

/* ---------------------------------------------------------------- reference-seam kernels
 * Same buffer meaning as the reference kernels so each can be checked in isolation. */

/* reference `intersect` (.cl:243): per ray and mesh, the number of crossings with t > EPSILON,
 * the nearest crossing distance (< max_ray_len) and its global triangle index (-1 = none).
 * d_near_t, d_hits, d_near_tri : (n, mesh_count) row-major, like isect_min_ray_len / isects_count /
 * ray_isect_mesh_idx_tmp.  d_scratch: lpcl_scratch_bytes(scene, n) bytes. */
int lpcl_intersect(const lpcl_scene *scene, const float *d_origin, const float *d_dir, int64_t n,
                   float max_ray_len, float *d_near_t, int32_t *d_hits, int32_t *d_near_tri, void *d_scratch,
                   void *stream);

/* reference `intersect_postproc` (.cl:105). */
int lpcl_intersect_postproc(const lpcl_scene *scene, const float *d_origin, const float *d_dir, float *d_dest,
                            const int32_t *d_prev_mesh, int32_t *d_n1, int32_t *d_n2, int32_t *d_entering,
                            int32_t *d_hit_mesh, int32_t *d_hit_tri, const float *d_near_t, const int32_t *d_hits,
                            const int32_t *d_near_tri, int64_t n, float max_ray_len, void *stream);

/* reference `reflect_refract_rays` (.cl:346); d_power and d_measured are updated in place
 * (.cl:392,467,470); children are written uncompacted at the parent's index. */
int lpcl_reflect_refract_rays(const lpcl_scene *scene, const float *d_origin, const float *d_dest,
                              const float *d_dir, float *d_power, int32_t *d_measured, const int32_t *d_n1,
                              const int32_t *d_n2, float *d_refl_origin, float *d_refl_dir, float *d_refl_power,
                              int32_t *d_refl_measured, float *d_refr_origin, float *d_refr_dir,
                              float *d_refr_power, int32_t *d_refr_measured, const int32_t *d_hit_mesh,
                              const int32_t *d_hit_tri, int64_t n, float ior_env, void *stream);

/* ---------------------------------------------------------------- production step
 * One generation of the trace loop, device resident (replaces iterative_tracer.py:261-375).
 *
 * Ray generation buffers (packed, 32 B per ray):
 *   org_pow  (cap,4) f32 : origin.xyz, w = power
 *   dir_prev (cap,4) f32 : direction.xyz, w = int32 bits of the previous hit-surface mesh id
 *                          (-2 for rays fresh from a source; reference it:118,344,373)
 *   lineage  (cap,) u64  : optional (may be NULL): (path << 32) | index, where bit k of `path` is set
 *                          when the ray took the REFRACTED branch at iteration k of the current
 *                          lineage epoch (k < 31, so the id stays positive as int64).  Sorting a
 *                          generation by lineage gives the reference order.  `index` is the source
 *                          ray index in the first epoch; before the 31 branch bits run out the
 *                          caller starts a new epoch by renumbering the live rays 0..n-1 in
 *                          generation order (lightpycl_b200/dist.py renumber_lineage) and passes
 *                          iteration numbers relative to the epoch start.
 * A generation of n rays is stored in two segments: logical ray i lives at physical index i for
 * i < seg_split, and at seg_base + (i - seg_split) otherwise.  (Generation 0: seg_split = n.)
 *
 * Outputs for the n parents (logical order == the reference's ray order):
 *   d_dest_meas (n,4) f32 : dest.xyz, w = int32 bits of `measured` (-1 terminated / 0 split / +1 measured)
 *   d_hit       (n,4) i32 : optional (may be NULL): hit mesh, hit triangle, n1, n2 (.cl:235-239)
 *   parent power is rewritten in place in org_pow.w when the ray crossed a dissipative medium.
 * Children (surviving ones only, reference it:366-373) go to the next generation's buffers:
 *   reflected survivors, parent order kept, at [0, n_reflect); refracted survivors at
 *   [children->seg_base, children->seg_base + n_refract), seg_base >= n.  With seg_base = n the next
 *   generation is (seg_split = n_reflect, seg_base = n); capacity needed: seg_base + n rays.
 *   A generation may also be processed in RAY TILES (the reference's <= 50,000-ray partitions,
 *   it:218-273): call the step on consecutive logical ranges of the parents (pointers advanced, see
 *   lightpycl_b200/iterative_tracer.py _Generation.c_struct) with the children pointers advanced by the
 *   reflected survivors so far and seg_base = N + refracted so far - reflected so far; the tiles'
 *   outputs then form the same two segments an untiled step writes, bit for bit.
 * Measured parents are appended, in ray order, as (dest.xyz, power) records to d_measured_out
 *   starting at record `measured_base` (reference get_measured_rays, it:395-411); when both
 *   parents->d_lineage and d_measured_lineage are given, the parent's lineage is stored alongside
 *   (used to merge the per-GPU buffers of a ray-sharded run back into the reference order).
 * d_counters receives the totals (read it after synchronising the stream).
 * d_workspace: lpcl_step_workspace_bytes(scene, n) bytes of device scratch (initialised by the call).
 */
typedef struct lpcl_step_counters {
    int64_t n_reflect;      /* surviving reflected children */
    int64_t n_refract;      /* surviving refracted children */
    int64_t n_measured;     /* parents that hit a measure surface this step */
    int64_t n_terminated;   /* parents terminated (terminator surface or no hit) */
    double live_power;      /* sum of surviving children's power (termination test it:372-387) */
    double measured_power;  /* sum of measured parents' power */
    int64_t candidates;     /* search statistics: filter candidates re-evaluated exactly */
    int64_t reserved;
} lpcl_step_counters;

typedef struct lpcl_generation {
    float *d_org_pow;      /* (cap,4) f32 */
    float *d_dir_prev;     /* (cap,4) f32 */
    uint64_t *d_lineage;   /* (cap,) or NULL */
    int64_t n;             /* logical ray count */
    int64_t seg_split;     /* rays in the first segment */
    int64_t seg_base;      /* physical start of the second segment */
    int64_t capacity;      /* rays the buffers can hold */
} lpcl_generation;

/* Tuning hint, not part of the reference interface (it has one kernel, cl:243): which epilogue variant of the
 * tensor-core search the searches of this scene use from now on.  0 = library default: the deferred-flag sweep (one
 * sign bit per 16-triangle chunk test, no branch in the sweep) followed by a second pass that evaluates every flagged
 * (ray, chunk) exactly (search_tc_kernel<RT, true> + resolve_records_kernel); 8 = eight fat epilogue warps that vote
 * after every chunk test and evaluate candidates from per-warp queues (search_tc_kernel<RT, false>); 16 = sixteen thin
 * warps in two groups (search_tc16_kernel).  The results are bit-identical in every case (the exact evaluation alone
 * decides); the environment variable LPCL_TC_WARPS, when set, overrides the hint. */
int lpcl_scene_set_tc_warps(lpcl_scene *scene, int32_t warps);
/* Kernel launches one lpcl_trace_search / lpcl_intersect call issues under the scene's current hint and the environment
 * (2 = scratch initialisation + search; 3 = + the second pass of the deferred-flag search).  For launch accounting. */
int32_t lpcl_search_launch_count(const lpcl_scene *scene);

int64_t lpcl_step_workspace_bytes(const lpcl_scene *scene, int64_t n_rays);

int lpcl_trace_step(const lpcl_scene *scene, const lpcl_generation *parents, const lpcl_generation *children,
                    float *d_dest_meas, int32_t *d_hit, float *d_measured_out, uint64_t *d_measured_lineage,
                    int64_t measured_base, int64_t measured_capacity, lpcl_step_counters *d_counters,
                    void *d_workspace, float max_ray_len, float ior_env, int32_t iteration, void *stream);

/* The two halves of lpcl_trace_step, callable separately so the caller can put events between
 * them: lpcl_trace_search initialises d_workspace, ZEROES *d_counters (when != NULL), runs the
 * nearest-hit search into the workspace and adds its candidate count to d_counters->candidates;
 * lpcl_trace_resolve runs the fused media-resolution / Fresnel / compaction kernel on the workspace
 * lpcl_trace_search just filled for the same parents and accumulates into d_counters.
 * lpcl_trace_step == search + resolve: three kernel launches (four with the default search, whose second pass is its
 * own launch; lpcl_search_launch_count), no memsets. */
int lpcl_trace_search(const lpcl_scene *scene, const lpcl_generation *parents, void *d_workspace,
                      float max_ray_len, lpcl_step_counters *d_counters, void *stream);
int lpcl_trace_resolve(const lpcl_scene *scene, const lpcl_generation *parents, const lpcl_generation *children,
                       float *d_dest_meas, int32_t *d_hit, float *d_measured_out, uint64_t *d_measured_lineage,
                       int64_t measured_base, int64_t measured_capacity, lpcl_step_counters *d_counters,
                       void *d_workspace, float max_ray_len, float ior_env, int32_t iteration, void *stream);

/* ---------------------------------------------------------------- small scenes: the whole trace in one launch
 * Replaces the complete host loop iterative_tracer.py:241-391 (every generation, both termination
 * tests it:383-391) for scenes of the size of the reference's example_nested_cubes_refraction.py
 * (10 rays x 48 triangles): one CTA runs the generations back to back, the caller reads the table
 * once.  Bit-identical to the general path (same device functions; unfiltered exact search).
 *
 * Arena: d_org_pow / d_dir_prev (/ d_lineage) hold `arena_rays` ray slots in the packed layout above;
 * the caller packs generation 0 (n0 rays) at slot 0 (lpcl_pack_rays).  Generation k occupies the
 * region [offset_k, offset_k + cap_k), cap_0 = n0, cap_{k+1} = 2 n_k, in the two-segment layout
 * (seg_split = n_reflect of its parents, seg_base = n of its parents); its per-ray results go to
 * d_dest_meas / d_hit at slots offset_k + logical ray index.  d_scratch: lpcl_small_scratch_bytes(scene,
 * scratch_rays) bytes.  Measured parents are appended to d_measured_out from record 0.
 * The kernel processes at most min(max_iterations, 31) generations and stops BEFORE one that has more
 * than pair_limit ray x triangle pairs, more than scratch_rays rays, or does not fit the arena / the
 * measured buffer: d_header then says stop_reason = 1 and describes the generation to continue with
 * (lpcl_trace_search / lpcl_trace_resolve on the same arena buffers).  power_threshold =
 * (1 - trace_until_dissipated) * input_power (it:383). */
typedef struct lpcl_small_iter {
    int64_t n_rays, n_reflect, n_refract, n_measured, n_terminated;
    int64_t offset;          /* arena slot of the generation's region */
    double live_power, measured_power;
} lpcl_small_iter;

typedef struct lpcl_small_header {
    int64_t iterations;      /* generations processed (rows of d_iters that are valid) */
    int64_t stop_reason;     /* 0: the trace ended; 1: handed over before generation `iterations` */
    int64_t next_offset, next_n, next_split, next_base; /* that generation: region, ray count, seg_split, seg_base */
    int64_t measured_total, reserved;
} lpcl_small_header;

int64_t lpcl_small_scratch_bytes(const lpcl_scene *scene, int64_t scratch_rays);
int lpcl_trace_small(const lpcl_scene *scene, float *d_org_pow, float *d_dir_prev, uint64_t *d_lineage, float *d_dest_meas,
                     int32_t *d_hit, int64_t arena_rays, int64_t n0, void *d_scratch, int64_t scratch_rays,
                     float *d_measured_out, uint64_t *d_measured_lineage, int64_t measured_capacity, float max_ray_len,
                     float ior_env, int32_t max_iterations, double power_threshold, int64_t pair_limit,
                     lpcl_small_iter *d_iters, lpcl_small_header *d_header, void *stream);

/* Pack / unpack helpers between the reference layout and the packed generation layout. */
int lpcl_pack_rays(const float *d_origin, const float *d_dir, const float *d_power, const int32_t *d_prev_mesh,
                   int64_t n, int64_t lineage_base, float *d_org_pow, float *d_dir_prev, uint64_t *d_lineage,
                   void *stream);

/* ---------------------------------------------------------------- measured-ray post-processing
 * (the consumer of the measured buffer; O(K), once per run.)
 *
 * lpcl_project: the reference kernels `stereograph_project` (.cl:488, launched it:515; kind = 1) and
 * `angular_project` (.cl:509, launched it:546,602; kind = 2), or the plain (x, y, power) that
 * get_binned_data bins (it:413-418; kind = 0).  d_records: (k,4) f32 (x, y, z, power) as written
 * by lpcl_trace_resolve; h_rot9: row-major 3x3 rotation (reference R_dev rows 0..2), NULL = identity;
 * outputs d_x, d_y, d_pwr_cor: (k,) f32 like the kernels' x / y / pwrs_cor buffers.
 *
 * lpcl_bin2d: np.histogram2d(x, y, bins=points, range=limits, weights=float64(w)/denom) as called at
 * it:416,529,560.  d_xedges, d_yedges: (points+1,) f64 bin edges (np.linspace(lo, hi, points+1));
 * d_hist: (points, points) f64, zeroed by the call, H[x_bin][y_bin]; right-most edge inclusive.
 */
int lpcl_project(const float *d_records, int64_t k, const float *h_rot9, int32_t kind, float *d_x, float *d_y,
                 float *d_pwr_cor, void *stream);
int lpcl_bin2d(const float *d_x, const float *d_y, const float *d_w, int64_t k, const double *d_xedges,
               const double *d_yedges, int32_t points, double denom, double *d_hist, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LIGHTPYCL_B200_H */
