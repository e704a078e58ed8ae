// lpcl_cabi.cu - extern "C" boundary of liblightpycl_b200.so (see include/lightpycl_b200.h).
// Host-side argument validation, scene construction (tile table) and kernel launches.
// No torch types, no exceptions across the boundary, nothing synchronises the stream.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/lightpycl_b200.h"
#include "lpcl_material.cuh"
#include "lpcl_search_exact.cuh"
#include "lpcl_search_tc.cuh"
#include "lpcl_small.cuh"

using namespace lpcl;

static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess) return fail((int)e__, "%s: %s", #expr, cudaGetErrorString(e__));    \
    } while (0)

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

typedef void (*SearchFn)(SceneDev, RaysDev, float, unsigned long long *, int *, int, unsigned long long *);
struct SearchVariant {
    int R, threads;
    SearchFn fn;
};
// rays per thread x threads per CTA; ordered by rays per CTA
static const SearchVariant kVariants[] = {
    {1, 128, search_kernel<1, 128>}, {1, 256, search_kernel<1, 256>}, {2, 256, search_kernel<2, 256>},
    {4, 128, search_kernel<4, 128>}, {4, 256, search_kernel<4, 256>}, {8, 128, search_kernel<8, 128>},
};
static const int kNumVariants = (int)(sizeof(kVariants) / sizeof(kVariants[0]));
static const int kDefaultVariant = 4;

// tensor-core search: ray tiles (128 rays each) per CTA
struct TcVariant {
    int rt;
    SearchFn fn;
};
typedef void (*SearchRecFn)(SceneDev, RaysDev, float, unsigned long long *, int *, int, unsigned long long *, RecordBuf);
struct TcRecVariant {
    int rt;
    SearchRecFn fn;
};
static const TcRecVariant kTcVariants[] = {
    {1, search_tc_kernel<1>}, {2, search_tc_kernel<2>}, {4, search_tc_kernel<4>}, {8, search_tc_kernel<8>}};
static const int kNumTcVariants = (int)(sizeof(kTcVariants) / sizeof(kTcVariants[0]));
static const TcRecVariant kTcfVariants[] = { // deferred-flag epilogue
    {1, search_tc_kernel<1, true>}, {2, search_tc_kernel<2, true>}, {4, search_tc_kernel<4, true>}, {8, search_tc_kernel<8, true>}};
static const TcVariant kTc16Variants[] = { // LPCL_TC_WARPS=16: two groups of eight thin epilogue warps (RT even; RT = 1 uses the default kernel)
    {1, nullptr}, {2, search_tc16_kernel<2>}, {4, search_tc16_kernel<4>}, {8, search_tc16_kernel<8>}};

struct lpcl_scene {
    SceneDev dev;
    int occupancy[8] = {1, 1, 1, 1, 1, 1, 1, 1};
    void *owned = nullptr;      // one allocation: filter tiles | tile table | materials | vmax
    int64_t owned_bytes = 0;
    int sm_count = 148;
    int device = 0;
    int tc_warps = 0;           // lpcl_scene_set_tc_warps: 0 = default, 8 / 16 = epilogue variant of the tensor-core search
};

static_assert(sizeof(StepCounters) == sizeof(lpcl_step_counters), "counter structs must match");
static_assert(sizeof(SmallIter) == sizeof(lpcl_small_iter) && sizeof(SmallHeader) == sizeof(lpcl_small_header),
              "small-trace structs must match");

extern "C" int lpcl_abi_version(void) { return LPCL_ABI_VERSION; }
extern "C" const char *lpcl_last_error(void) { return g_err; }

static int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// upper bound of the device bytes a scene needs: tiles never straddle meshes, so at most
// ceil(T / TILE_TRIS) + mesh_count tiles
static int64_t scene_bytes_bound(int64_t tri_count, int64_t mesh_count)
{
    const int64_t max_tiles = (tri_count + TILE_TRIS - 1) / TILE_TRIS + mesh_count;
    const int64_t max_tc_tiles = (tri_count + TC_TILE_TRIS - 1) / TC_TILE_TRIS + mesh_count;
    return align_up(max_tiles * TILE_BYTES, 256) + align_up(max_tiles * (int64_t)sizeof(TileInfo), 256) +
           align_up(max_tc_tiles * TC_TILE_BYTES, 256) + align_up(max_tc_tiles * (int64_t)sizeof(TileInfo), 256) +
           4 * align_up(mesh_count * 4, 256) + 256;
}

extern "C" int64_t lpcl_scene_bytes(int32_t tri_count, int32_t mesh_count)
{
    if (tri_count <= 0 || mesh_count <= 0) return 0;
    return scene_bytes_bound(tri_count, mesh_count);
}

static int scene_create_impl(lpcl_scene **out, const float *d_v0, const float *d_v1, const float *d_v2,
                             const int32_t *h_tri_mesh, int32_t tri_count, const int32_t *h_mat_type,
                             const float *h_ior, const float *h_refl, const float *h_diss, int32_t mesh_count,
                             void *d_mem, int64_t mem_bytes, bool borrow, void *stream_)
{
    if (!out || !d_v0 || !d_v1 || !d_v2 || !h_tri_mesh || !h_mat_type || !h_ior || !h_refl || !h_diss)
        return fail(LPCL_E_NULL, "lpcl_scene_create: NULL argument");
    if (tri_count <= 0 || mesh_count <= 0) return fail(LPCL_E_RANGE, "lpcl_scene_create: tri_count/mesh_count must be > 0");
    if (!aligned16(d_v0) || !aligned16(d_v1) || !aligned16(d_v2))
        return fail(LPCL_E_ALIGN, "lpcl_scene_create: vertex buffers must be 16-byte aligned");
    cudaStream_t stream = (cudaStream_t)stream_;

    // tile table: contiguous ascending mesh ids, every mesh non-empty (reference relies on it, cl:262-265)
    std::vector<TileInfo> tiles, tc_tiles;
    int expect = 0;
    if (h_tri_mesh[0] != 0) return fail(LPCL_E_MESH_IDS, "mesh ids must start at 0 (got %d)", h_tri_mesh[0]);
    int start = 0;
    for (int i = 1; i <= tri_count; ++i) {
        if (i == tri_count || h_tri_mesh[i] != h_tri_mesh[i - 1]) {
            if (i < tri_count && h_tri_mesh[i] != h_tri_mesh[i - 1] + 1)
                return fail(LPCL_E_MESH_IDS, "mesh ids must ascend by 1 (triangle %d: %d after %d)", i, h_tri_mesh[i],
                            h_tri_mesh[i - 1]);
            for (int t0 = start; t0 < i; t0 += TILE_TRIS) {
                TileInfo ti;
                ti.tri0 = t0;
                ti.count = (i - t0 < TILE_TRIS) ? (i - t0) : TILE_TRIS;
                ti.mesh = expect;
                ti.pad = 0;
                tiles.push_back(ti);
            }
            for (int t0 = start; t0 < i; t0 += TC_TILE_TRIS) {
                TileInfo ti;
                ti.tri0 = t0;
                ti.count = (i - t0 < TC_TILE_TRIS) ? (i - t0) : TC_TILE_TRIS;
                ti.mesh = expect;
                ti.pad = 0;
                tc_tiles.push_back(ti);
            }
            start = i;
            ++expect;
        }
    }
    if (expect != mesh_count)
        return fail(LPCL_E_MESH_IDS, "%d meshes declared but triangles reference %d (empty meshes are not allowed)",
                    mesh_count, expect);

    lpcl_scene *sc = new lpcl_scene();
    cudaGetDevice(&sc->device);
    cudaDeviceGetAttribute(&sc->sm_count, cudaDevAttrMultiProcessorCount, sc->device);
    const int n_tiles = (int)tiles.size();
    const int64_t off_filter = 0;
    const int64_t off_tiles = align_up(off_filter + (int64_t)n_tiles * TILE_BYTES, 256);
    const int64_t off_mat = align_up(off_tiles + (int64_t)n_tiles * sizeof(TileInfo), 256);
    const int64_t off_ior = off_mat + align_up((int64_t)mesh_count * 4, 256);
    const int64_t off_refl = off_ior + align_up((int64_t)mesh_count * 4, 256);
    const int64_t off_diss = off_refl + align_up((int64_t)mesh_count * 4, 256);
    const int64_t off_vmax = off_diss + align_up((int64_t)mesh_count * 4, 256);
    const int tc_n_tiles = (int)tc_tiles.size();
    const int64_t off_tc_filter = off_vmax + 256;
    const int64_t off_tc_tiles = align_up(off_tc_filter + (int64_t)tc_n_tiles * TC_TILE_BYTES, 256);
    const int64_t total = align_up(off_tc_tiles + (int64_t)tc_n_tiles * sizeof(TileInfo), 256);
    char *base = nullptr;
    if (borrow) { // caller-owned memory (e.g. a torch tensor): no cudaMalloc / cudaFree on this path
        if (!d_mem || mem_bytes < total || (reinterpret_cast<uintptr_t>(d_mem) & 255u)) {
            delete sc;
            return fail(d_mem ? LPCL_E_RANGE : LPCL_E_NULL, "lpcl_scene_create_in: need %lld bytes of 256-byte aligned device memory",
                        (long long)total);
        }
        base = (char *)d_mem;
        sc->owned = nullptr;
    } else {
        cudaError_t e = cudaMalloc(&sc->owned, (size_t)total);
        if (e != cudaSuccess) {
            delete sc;
            return fail((int)e, "cudaMalloc(%lld): %s", (long long)total, cudaGetErrorString(e));
        }
        base = (char *)sc->owned;
    }
    sc->owned_bytes = total;
    SceneDev &d = sc->dev;
    d.v0 = (const float4 *)d_v0;
    d.v1 = (const float4 *)d_v1;
    d.v2 = (const float4 *)d_v2;
    d.filter = (const float4 *)(base + off_filter);
    d.tiles = (const TileInfo *)(base + off_tiles);
    d.mat_type = (const int *)(base + off_mat);
    d.ior = (const float *)(base + off_ior);
    d.refl = (const float *)(base + off_refl);
    d.diss = (const float *)(base + off_diss);
    d.vmax = (const float *)(base + off_vmax);
    d.tc_filter = (const float *)(base + off_tc_filter);
    d.tc_tiles = (const TileInfo *)(base + off_tc_tiles);
    d.tc_n_tiles = tc_n_tiles;
    d.n_tiles = n_tiles;
    d.tri_count = tri_count;
    d.mesh_count = mesh_count;

#define SCENE_TRY(expr)                                                                  \
    do {                                                                                 \
        cudaError_t e__ = (expr);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            if (sc->owned) cudaFree(sc->owned);                                          \
            delete sc;                                                                   \
            return fail((int)e__, "%s: %s", #expr, cudaGetErrorString(e__));             \
        }                                                                                \
    } while (0)
    // small host->device copies; the sources are pageable so these complete before returning
    SCENE_TRY(cudaMemcpyAsync(base + off_tiles, tiles.data(), n_tiles * sizeof(TileInfo), cudaMemcpyHostToDevice, stream));
    SCENE_TRY(cudaMemcpyAsync(base + off_tc_tiles, tc_tiles.data(), tc_n_tiles * sizeof(TileInfo), cudaMemcpyHostToDevice, stream));
    SCENE_TRY(cudaMemcpyAsync(base + off_mat, h_mat_type, mesh_count * 4, cudaMemcpyHostToDevice, stream));
    SCENE_TRY(cudaMemcpyAsync(base + off_ior, h_ior, mesh_count * 4, cudaMemcpyHostToDevice, stream));
    SCENE_TRY(cudaMemcpyAsync(base + off_refl, h_refl, mesh_count * 4, cudaMemcpyHostToDevice, stream));
    SCENE_TRY(cudaMemcpyAsync(base + off_diss, h_diss, mesh_count * 4, cudaMemcpyHostToDevice, stream));
    SCENE_TRY(cudaMemsetAsync(base + off_vmax, 0, 256, stream));
    unsigned *vmax_bits = (unsigned *)(base + off_vmax + 128);
    vmax_kernel<<<(tri_count + 255) / 256, 256, 0, stream>>>(d.v0, d.v1, d.v2, tri_count, vmax_bits);
    vmax_finish_kernel<<<1, 1, 0, stream>>>(vmax_bits, (float *)(base + off_vmax));
    const long long slots = (long long)n_tiles * TILE_TRIS;
    build_filter_kernel<<<(unsigned)((slots + 255) / 256), 256, 0, stream>>>(d.v0, d.v1, d.v2, d.tiles, n_tiles, d.vmax,
                                                                            (float4 *)(base + off_filter));
    const long long tc_slots = (long long)tc_n_tiles * TC_TILE_TRIS;
    TileInfo *tc_tiles_dev = (TileInfo *)(base + off_tc_tiles);
    tc_tile_vmax_kernel<<<(unsigned)((tc_slots + 127) / 128), 128, 0, stream>>>(d.v0, d.v1, d.v2, tc_tiles_dev, tc_n_tiles);
    tc_tile_vmax_finish_kernel<<<(unsigned)((tc_n_tiles + 127) / 128), 128, 0, stream>>>(tc_tiles_dev, tc_n_tiles);
    build_tc_kernel<<<(unsigned)((tc_slots + 127) / 128), 128, 0, stream>>>(d.v0, d.v1, d.v2, d.tc_tiles, tc_n_tiles,
                                                                           (float *)(base + off_tc_filter));
    SCENE_TRY(cudaGetLastError());
    for (int v = 0; v < kNumTcVariants; ++v) {
        SCENE_TRY(cudaFuncSetAttribute(kTcVariants[v].fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       tc_smem_bytes(kTcVariants[v].rt)));
        SCENE_TRY(cudaFuncSetAttribute(kTcfVariants[v].fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       tcf_smem_bytes(kTcfVariants[v].rt)));
        if (kTc16Variants[v].fn)
            SCENE_TRY(cudaFuncSetAttribute(kTc16Variants[v].fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           tc16_smem_bytes(kTc16Variants[v].rt)));
    }
    for (int v = 0; v < kNumVariants; ++v) {
        SCENE_TRY(cudaFuncSetAttribute(kVariants[v].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SEARCH_SMEM_BYTES));
        int occ = 0;
        SCENE_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kVariants[v].fn, kVariants[v].threads, SEARCH_SMEM_BYTES));
        sc->occupancy[v] = occ > 0 ? occ : 1;
    }
#undef SCENE_TRY
    *out = sc;
    return 0;
}

extern "C" int lpcl_scene_create(lpcl_scene **out, const float *d_v0, const float *d_v1, const float *d_v2,
                                 const int32_t *h_tri_mesh, int32_t tri_count, const int32_t *h_mat_type,
                                 const float *h_ior, const float *h_refl, const float *h_diss, int32_t mesh_count,
                                 void *stream_)
{
    return scene_create_impl(out, d_v0, d_v1, d_v2, h_tri_mesh, tri_count, h_mat_type, h_ior, h_refl, h_diss, mesh_count,
                             nullptr, 0, false, stream_);
}

extern "C" int lpcl_scene_create_in(lpcl_scene **out, const float *d_v0, const float *d_v1, const float *d_v2,
                                    const int32_t *h_tri_mesh, int32_t tri_count, const int32_t *h_mat_type,
                                    const float *h_ior, const float *h_refl, const float *h_diss, int32_t mesh_count,
                                    void *d_scene_mem, int64_t scene_mem_bytes, void *stream_)
{
    return scene_create_impl(out, d_v0, d_v1, d_v2, h_tri_mesh, tri_count, h_mat_type, h_ior, h_refl, h_diss, mesh_count,
                             d_scene_mem, scene_mem_bytes, true, stream_);
}

extern "C" int lpcl_scene_set_tc_warps(lpcl_scene *scene, int32_t warps)
{
    if (!scene) return fail(LPCL_E_NULL, "lpcl_scene_set_tc_warps: NULL scene");
    if (warps != 0 && warps != 8 && warps != 16) return fail(LPCL_E_RANGE, "lpcl_scene_set_tc_warps: warps must be 0, 8 or 16 (got %d)", warps);
    scene->tc_warps = warps;
    return 0;
}

static int env_int(const char *name, int dflt);

extern "C" int32_t lpcl_search_launch_count(const lpcl_scene *scene)
{
    if (!scene) return 0;
    const char *impl = getenv("LPCL_SEARCH_IMPL");
    if (impl && (strcmp(impl, "exact") == 0 || strcmp(impl, "simt") == 0)) return 2;
    int tc_warps = env_int("LPCL_TC_WARPS", -1);
    if (tc_warps < 0) tc_warps = scene->tc_warps;
    return tc_warps == 8 || tc_warps == 16 ? 2 : 3;
}

extern "C" void lpcl_scene_destroy(lpcl_scene *scene)
{
    if (!scene) return;
    if (scene->owned) cudaFree(scene->owned);
    delete scene;
}

extern "C" int32_t lpcl_scene_tri_count(const lpcl_scene *s) { return s ? s->dev.tri_count : 0; }
extern "C" int32_t lpcl_scene_mesh_count(const lpcl_scene *s) { return s ? s->dev.mesh_count : 0; }
extern "C" int64_t lpcl_scene_device_bytes(const lpcl_scene *s) { return s ? s->owned_bytes : 0; }

// scratch layout: key[mesh_count * n] (u64) | cnt[mesh_count * n] (i32), 256-byte aligned parts
static int64_t scratch_key_bytes(const lpcl_scene *s, int64_t n) { return align_up((int64_t)s->dev.mesh_count * n * 8, 256); }
static int64_t scratch_cnt_bytes(const lpcl_scene *s, int64_t n) { return align_up((int64_t)s->dev.mesh_count * n * 4, 256); }

// ... | record header (256 bytes: the append counter) | records: (ray, chunk) pairs the deferred-flag search hands to
// resolve_records_kernel (lpcl_search_tc.cuh).  Records that do not fit are evaluated inside the search kernel, so the
// budget per ray is a matter of speed only: LPCL_REC_PER_RAY (default 32 records = 256 bytes per ray; C4 needs 29)
static int64_t scratch_rec_capacity(int64_t n)
{
    static const int per_ray = []() { const char *v = getenv("LPCL_REC_PER_RAY"); int k = v ? atoi(v) : 32; return k < 0 ? 0 : k; }();
    int64_t cap = n * per_ray;
    if (per_ray > 0 && cap < 4096) cap = 4096;
    if (cap > (1LL << 28)) cap = 1LL << 28;
    return cap;
}
static int64_t scratch_rec_bytes(int64_t n) { return 256 + align_up(scratch_rec_capacity(n) * 8, 256); }

extern "C" int64_t lpcl_scratch_bytes(const lpcl_scene *s, int64_t n)
{
    if (!s || n < 0) return 0;
    return scratch_key_bytes(s, n) + scratch_cnt_bytes(s, n) + scratch_rec_bytes(n) + 256;
}

static int64_t step_blocks(int64_t n) { return (n + STEP_RAYS - 1) / STEP_RAYS; }

extern "C" int64_t lpcl_step_workspace_bytes(const lpcl_scene *s, int64_t n)
{
    if (!s || n < 0) return 0;
    return lpcl_scratch_bytes(s, n) + align_up(step_blocks(n) * 16 + 16, 256) + 256; // scratch | scan_bytes(n) | slack
}

// LPCL_SEARCH_R / LPCL_SEARCH_THREADS / LPCL_SEARCH_CHUNKS override the launch heuristics without rebuilding.
static int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

// zero_a / zero_b: optional 64-bit-word areas cleared by the scratch initialisation launch
static int launch_search(const lpcl_scene *s, const RaysDev &rays, float max_ray_len, void *d_scratch,
                         unsigned long long *cand_total, cudaStream_t stream, unsigned long long *zero_a = nullptr,
                         long long words_a = 0, unsigned long long *zero_b = nullptr, long long words_b = 0)
{
    const int64_t n = rays.n;
    unsigned long long *key = (unsigned long long *)d_scratch;
    int *cnt = (int *)((char *)d_scratch + scratch_key_bytes(s, n));
    const long long total = (long long)s->dev.mesh_count * n;
    RecordBuf records;
    records.count = (unsigned long long *)((char *)d_scratch + scratch_key_bytes(s, n) + scratch_cnt_bytes(s, n));
    records.rec = (int2 *)((char *)records.count + 256);
    records.capacity = scratch_rec_capacity(n);
    const unsigned long long init_key =
        ((unsigned long long)(*reinterpret_cast<const unsigned *>(&max_ray_len)) << 32) | 0xffffffffull;
    {
        long long blocks = (total + 255) / 256;
        if (blocks > (long long)s->sm_count * 32) blocks = (long long)s->sm_count * 32;
        scratch_init_kernel<<<(unsigned)blocks, 256, 0, stream>>>(key, cnt, total, init_key, zero_a, words_a, zero_b, words_b,
                                                                  records.count, 256 / 8);
    }
    // Default: the filter on the tensor cores (lpcl_search_tc.cuh).  LPCL_SEARCH_IMPL=simt selects the
    // packed-FP32 filter (lpcl_search.cuh) - kept as the second implementation the tests cross-check.
    const char *impl = getenv("LPCL_SEARCH_IMPL");
    if (impl && strcmp(impl, "exact") == 0) {
        // no filter at all: every pair takes the exact evaluation (lpcl_search_exact.cuh) - the on-GPU
        // yardstick of the full-size parity tests, ~3x slower than the filtered searches
        const int64_t ray_blocks = (n + EXACT_THREADS - 1) / EXACT_THREADS;
        int64_t chunks = env_int("LPCL_SEARCH_CHUNKS", 0);
        if (chunks <= 0) chunks = (32 * (int64_t)s->sm_count + ray_blocks - 1) / ray_blocks;
        if (chunks < 1) chunks = 1;
        if (chunks > s->dev.n_tiles) chunks = s->dev.n_tiles;
        if (chunks > 65535) chunks = 65535;
        int tiles_per_chunk = (int)((s->dev.n_tiles + chunks - 1) / chunks);
        chunks = (s->dev.n_tiles + tiles_per_chunk - 1) / tiles_per_chunk;
        dim3 grid((unsigned)ray_blocks, (unsigned)chunks);
        search_exact_kernel<<<grid, EXACT_THREADS, 0, stream>>>(s->dev, rays, max_ray_len, key, cnt, tiles_per_chunk, cand_total);
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    if (!impl || strcmp(impl, "simt") != 0) {
        const int want_rt = env_int("LPCL_SEARCH_RT", 0);
        int v = -1;
        for (int k = 0; k < kNumTcVariants; ++k)
            if (kTcVariants[k].rt == want_rt) v = k;
        if (v < 0) { // 4 ray tiles per CTA (measured best on C3), fewer while the ray blocks would be mostly padding
            v = 2;
            while (v > 0 && n < 8LL * kTcVariants[v].rt * TC_M) --v;
        }
        const int64_t per_cta = (int64_t)kTcVariants[v].rt * TC_M;
        const int64_t ray_blocks = (n + per_cta - 1) / per_cta;
        int64_t chunks = env_int("LPCL_SEARCH_CHUNKS", 0);
        if (chunks <= 0) { // one CTA per SM
            // fill: at least 8 waves of CTAs when rays are few; tail: with many ray blocks, cut the sweep so
            // that the partially filled last wave is short (32 waves) as long as a CTA keeps >= 200 tiles of
            // work to amortise its set-up (measured on C3: 1 chunk 2.69e12, 3-4 chunks 2.84e12 tests/s)
            const int64_t fill = (8 * (int64_t)s->sm_count + ray_blocks - 1) / ray_blocks;
            int64_t tail = (32 * (int64_t)s->sm_count + ray_blocks - 1) / ray_blocks;
            const int64_t cap = s->dev.tc_n_tiles / 200 > 1 ? s->dev.tc_n_tiles / 200 : 1;
            if (tail > cap) tail = cap;
            chunks = fill > tail ? fill : tail;
        }
        if (chunks < 1) chunks = 1;
        if (chunks > s->dev.tc_n_tiles) chunks = s->dev.tc_n_tiles;
        if (chunks > 65535) chunks = 65535;
        int tiles_per_chunk = (int)((s->dev.tc_n_tiles + chunks - 1) / chunks);
        chunks = (s->dev.tc_n_tiles + tiles_per_chunk - 1) / tiles_per_chunk;
        dim3 grid((unsigned)ray_blocks, (unsigned)chunks);
        // epilogue variant: 0 = deferred flags + second pass (default), 8 / 16 = a vote per chunk test, fat / thin warps
        int tc_warps = env_int("LPCL_TC_WARPS", -1); // the environment (tests, tuning) beats the scene's hint
        if (tc_warps < 0) tc_warps = s->tc_warps;
        if (tc_warps == 16 && kTc16Variants[v].fn)
            kTc16Variants[v].fn<<<grid, TC16_THREADS, tc16_smem_bytes(kTc16Variants[v].rt), stream>>>(s->dev, rays, max_ray_len, key, cnt,
                                                                                                      tiles_per_chunk, cand_total);
        else if (tc_warps == 8 || tc_warps == 16)
            kTcVariants[v].fn<<<grid, TC_THREADS, tc_smem_bytes(kTcVariants[v].rt), stream>>>(s->dev, rays, max_ray_len, key, cnt,
                                                                                              tiles_per_chunk, cand_total, records);
        else {
            kTcfVariants[v].fn<<<grid, TC_THREADS, tcf_smem_bytes(kTcfVariants[v].rt), stream>>>(s->dev, rays, max_ray_len, key, cnt,
                                                                                                tiles_per_chunk, cand_total, records);
            CUDA_TRY(cudaGetLastError());
            // second pass: the exact evaluation of every flagged (ray, chunk) record, at full occupancy
            resolve_records_kernel<<<(unsigned)(s->sm_count * 8), 256, 0, stream>>>(s->dev, rays, max_ray_len, key, cnt, records);
        }
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    // variant: forced by the environment (tests / tuning), else the default, stepping down to fewer
    // rays per CTA while the last ray block would be mostly padding
    const int want_r = env_int("LPCL_SEARCH_R", 0), want_t = env_int("LPCL_SEARCH_THREADS", 0);
    int v = -1;
    for (int k = 0; k < kNumVariants; ++k)
        if (kVariants[k].R == want_r && (want_t == 0 || kVariants[k].threads == want_t)) { v = k; break; }
    if (v < 0) {
        v = kDefaultVariant;
        while (v > 0 && n < 16LL * kVariants[v].R * kVariants[v].threads) --v;
    }
    const int64_t per_cta = (int64_t)kVariants[v].R * kVariants[v].threads;
    const int64_t ray_blocks = (n + per_cta - 1) / per_cta;
    // split the triangle range over blockIdx.y until the grid is >= 8 full waves: fills the SMs when
    // rays are few and keeps the partially filled last wave below ~1/8 of the run when they are many
    int64_t chunks = env_int("LPCL_SEARCH_CHUNKS", 0);
    if (chunks <= 0) {
        const int64_t wave = (int64_t)s->sm_count * s->occupancy[v];
        chunks = (8 * wave + ray_blocks - 1) / ray_blocks;
    }
    if (chunks < 1) chunks = 1;
    if (chunks > s->dev.n_tiles) chunks = s->dev.n_tiles;
    if (chunks > 65535) chunks = 65535;
    int tiles_per_chunk = (int)((s->dev.n_tiles + chunks - 1) / chunks);
    chunks = (s->dev.n_tiles + tiles_per_chunk - 1) / tiles_per_chunk;
    dim3 grid((unsigned)ray_blocks, (unsigned)chunks);
    kVariants[v].fn<<<grid, kVariants[v].threads, SEARCH_SMEM_BYTES, stream>>>(s->dev, rays, max_ray_len, key, cnt,
                                                                              tiles_per_chunk, cand_total);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int check_gen(const lpcl_generation *g, const char *who, bool need_rays)
{
    if (!g) return fail(LPCL_E_NULL, "%s: NULL generation", who);
    if (!g->d_org_pow || !g->d_dir_prev) return fail(LPCL_E_NULL, "%s: NULL ray buffer", who);
    if (!aligned16(g->d_org_pow) || !aligned16(g->d_dir_prev)) return fail(LPCL_E_ALIGN, "%s: ray buffers must be 16-byte aligned", who);
    if (need_rays) {
        if (g->n <= 0 || g->n >= (1LL << 31)) return fail(LPCL_E_RANGE, "%s: n=%lld out of range", who, (long long)g->n);
        if (g->seg_split < 0 || g->seg_split > g->n) return fail(LPCL_E_RANGE, "%s: seg_split out of range", who);
        const int64_t second = g->n - g->seg_split;
        if (g->seg_split > g->capacity || (second > 0 && (g->seg_base < g->seg_split || g->seg_base + second > g->capacity)))
            return fail(LPCL_E_RANGE, "%s: segments exceed capacity", who);
    }
    return 0;
}

static RaysDev to_dev(const lpcl_generation *g)
{
    RaysDev r;
    r.org_pow = (float4 *)g->d_org_pow;
    r.dir_prev = (float4 *)g->d_dir_prev;
    r.lineage = (unsigned long long *)g->d_lineage;
    r.n = g->n;
    r.split = g->seg_split;
    r.base = g->seg_base;
    return r;
}

extern "C" int lpcl_intersect(const lpcl_scene *s, const float *d_origin, const float *d_dir, int64_t n,
                              float max_ray_len, float *d_near_t, int32_t *d_hits, int32_t *d_near_tri, void *d_scratch,
                              void *stream_)
{
    if (!s || !d_origin || !d_dir || !d_near_t || !d_hits || !d_near_tri || !d_scratch)
        return fail(LPCL_E_NULL, "lpcl_intersect: NULL argument");
    if (n <= 0 || n >= (1LL << 31)) return fail(LPCL_E_RANGE, "lpcl_intersect: n=%lld out of range", (long long)n);
    if (!aligned16(d_origin) || !aligned16(d_dir)) return fail(LPCL_E_ALIGN, "lpcl_intersect: ray buffers must be 16-byte aligned");
    cudaStream_t stream = (cudaStream_t)stream_;
    RaysDev rays;
    rays.org_pow = (float4 *)d_origin; // .w lanes are ignored by the search
    rays.dir_prev = (float4 *)d_dir;
    rays.lineage = nullptr;
    rays.n = n;
    rays.split = n;
    rays.base = n;
    int rc = launch_search(s, rays, max_ray_len, d_scratch, nullptr, stream);
    if (rc) return rc;
    const unsigned long long *key = (const unsigned long long *)d_scratch;
    const int *cnt = (const int *)((char *)d_scratch + scratch_key_bytes(s, n));
    scratch_unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(key, cnt, n, s->dev.mesh_count, d_near_t,
                                                                          d_hits, d_near_tri);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int lpcl_intersect_postproc(const lpcl_scene *s, const float *d_origin, const float *d_dir, float *d_dest,
                                       const int32_t *d_prev_mesh, int32_t *d_n1, int32_t *d_n2, int32_t *d_entering,
                                       int32_t *d_hit_mesh, int32_t *d_hit_tri, const float *d_near_t,
                                       const int32_t *d_hits, const int32_t *d_near_tri, int64_t n, float max_ray_len,
                                       void *stream_)
{
    if (!s || !d_origin || !d_dir || !d_dest || !d_prev_mesh || !d_n1 || !d_n2 || !d_entering || !d_hit_mesh ||
        !d_hit_tri || !d_near_t || !d_hits || !d_near_tri)
        return fail(LPCL_E_NULL, "lpcl_intersect_postproc: NULL argument");
    if (n <= 0) return fail(LPCL_E_RANGE, "lpcl_intersect_postproc: n must be > 0");
    postproc_ref_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(
        s->dev, (const float4 *)d_origin, (const float4 *)d_dir, (float4 *)d_dest, d_prev_mesh, d_n1, d_n2, d_entering,
        d_hit_mesh, d_hit_tri, d_near_t, d_hits, d_near_tri, n, max_ray_len);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int lpcl_reflect_refract_rays(const lpcl_scene *s, const float *d_origin, const float *d_dest,
                                         const float *d_dir, float *d_power, int32_t *d_measured, const int32_t *d_n1,
                                         const int32_t *d_n2, float *d_refl_origin, float *d_refl_dir,
                                         float *d_refl_power, int32_t *d_refl_measured, float *d_refr_origin,
                                         float *d_refr_dir, float *d_refr_power, int32_t *d_refr_measured,
                                         const int32_t *d_hit_mesh, const int32_t *d_hit_tri, int64_t n, float ior_env,
                                         void *stream_)
{
    if (!s || !d_origin || !d_dest || !d_dir || !d_power || !d_measured || !d_n1 || !d_n2 || !d_refl_origin ||
        !d_refl_dir || !d_refl_power || !d_refl_measured || !d_refr_origin || !d_refr_dir || !d_refr_power ||
        !d_refr_measured || !d_hit_mesh || !d_hit_tri)
        return fail(LPCL_E_NULL, "lpcl_reflect_refract_rays: NULL argument");
    if (n <= 0) return fail(LPCL_E_RANGE, "lpcl_reflect_refract_rays: n must be > 0");
    scatter_ref_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(
        s->dev, (const float4 *)d_origin, (const float4 *)d_dest, (const float4 *)d_dir, d_power, d_measured, d_n1,
        d_n2, (float4 *)d_refl_origin, (float4 *)d_refl_dir, d_refl_power, d_refl_measured, (float4 *)d_refr_origin,
        (float4 *)d_refr_dir, d_refr_power, d_refr_measured, d_hit_mesh, d_hit_tri, n, ior_env);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int lpcl_pack_rays(const float *d_origin, const float *d_dir, const float *d_power,
                              const int32_t *d_prev_mesh, int64_t n, int64_t lineage_base, float *d_org_pow,
                              float *d_dir_prev, uint64_t *d_lineage, void *stream_)
{
    if (!d_origin || !d_dir || !d_power || !d_org_pow || !d_dir_prev) return fail(LPCL_E_NULL, "lpcl_pack_rays: NULL argument");
    if (n <= 0) return fail(LPCL_E_RANGE, "lpcl_pack_rays: n must be > 0");
    pack_rays_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(
        (const float4 *)d_origin, (const float4 *)d_dir, d_power, d_prev_mesh, n, lineage_base, (float4 *)d_org_pow,
        (float4 *)d_dir_prev, (unsigned long long *)d_lineage);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int64_t scan_bytes(int64_t n) { return align_up(step_blocks(n) * 16 + 16, 256); }

extern "C" int lpcl_trace_search(const lpcl_scene *s, const lpcl_generation *parents, void *d_workspace,
                                 float max_ray_len, lpcl_step_counters *d_counters, void *stream_)
{
    if (!s || !d_workspace) return fail(LPCL_E_NULL, "lpcl_trace_search: NULL argument");
    int rc = check_gen(parents, "lpcl_trace_search(parents)", true);
    if (rc) return rc;
    // the scratch initialisation also clears the counters and the look-back words lpcl_trace_resolve
    // scans with: one generation = 3 launches (init, search, step), no memsets
    char *scan = (char *)d_workspace + lpcl_scratch_bytes(s, parents->n);
    return launch_search(s, to_dev(parents), max_ray_len, d_workspace,
                         d_counters ? (unsigned long long *)&d_counters->candidates : nullptr, (cudaStream_t)stream_,
                         (unsigned long long *)d_counters, d_counters ? (long long)(sizeof(lpcl_step_counters) / 8) : 0,
                         (unsigned long long *)scan, scan_bytes(parents->n) / 8);
}

extern "C" int lpcl_trace_resolve(const lpcl_scene *s, const lpcl_generation *parents, const lpcl_generation *children,
                                  float *d_dest_meas, int32_t *d_hit, float *d_measured_out,
                                  uint64_t *d_measured_lineage, int64_t measured_base, int64_t measured_capacity,
                                  lpcl_step_counters *d_counters, void *d_workspace, float max_ray_len, float ior_env,
                                  int32_t iteration, void *stream_)
{
    if (!s || !d_dest_meas || !d_counters || !d_workspace) return fail(LPCL_E_NULL, "lpcl_trace_resolve: NULL argument");
    int rc = check_gen(parents, "lpcl_trace_resolve(parents)", true);
    if (rc) return rc;
    rc = check_gen(children, "lpcl_trace_resolve(children)", false);
    if (rc) return rc;
    const int64_t n = parents->n;
    // children: reflected survivors fill [0, n_reflect), refracted ones [seg_base, seg_base + n_refract); both counts <= n
    if (children->seg_base < n || children->capacity < children->seg_base + n)
        return fail(LPCL_E_RANGE, "lpcl_trace_resolve: children need seg_base >= n and capacity >= seg_base + n (n = %lld, "
                    "seg_base = %lld, capacity = %lld)", (long long)n, (long long)children->seg_base, (long long)children->capacity);
    if (d_measured_out && measured_capacity < measured_base + n)
        return fail(LPCL_E_RANGE, "lpcl_trace_resolve: measured buffer too small (%lld < %lld)",
                    (long long)measured_capacity, (long long)(measured_base + n));
    if (!aligned16(d_dest_meas) || (d_measured_out && !aligned16(d_measured_out)) || (d_hit && !aligned16(d_hit)))
        return fail(LPCL_E_ALIGN, "lpcl_trace_resolve: output buffers must be 16-byte aligned");
    cudaStream_t stream = (cudaStream_t)stream_;
    RaysDev par = to_dev(parents), chi = to_dev(children);
    char *ws = (char *)d_workspace;
    const unsigned long long *key = (const unsigned long long *)ws;
    const int *cnt = (const int *)(ws + scratch_key_bytes(s, n));
    char *scan = ws + lpcl_scratch_bytes(s, n); // cleared by lpcl_trace_search
    unsigned long long *words = (unsigned long long *)(scan + 16); // two look-back status words per block
    step_kernel<<<(unsigned)step_blocks(n), STEP_THREADS, 0, stream>>>(
        s->dev, par, chi, key, cnt, (float4 *)d_dest_meas, (int4 *)d_hit, (float4 *)d_measured_out,
        (unsigned long long *)d_measured_lineage, measured_base, measured_capacity, (StepCounters *)d_counters, words,
        max_ray_len, ior_env, iteration);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int lpcl_trace_step(const lpcl_scene *s, const lpcl_generation *parents, const lpcl_generation *children,
                               float *d_dest_meas, int32_t *d_hit, float *d_measured_out, uint64_t *d_measured_lineage,
                               int64_t measured_base, int64_t measured_capacity, lpcl_step_counters *d_counters,
                               void *d_workspace, float max_ray_len, float ior_env, int32_t iteration, void *stream_)
{
    if (!d_counters) return fail(LPCL_E_NULL, "lpcl_trace_step: NULL counters");
    int rc = lpcl_trace_search(s, parents, d_workspace, max_ray_len, d_counters, stream_);
    if (rc) return rc;
    return lpcl_trace_resolve(s, parents, children, d_dest_meas, d_hit, d_measured_out, d_measured_lineage,
                              measured_base, measured_capacity, d_counters, d_workspace, max_ray_len, ior_env, iteration,
                              stream_);
}

// ------------------------------------------------------------------------------------------ small scenes
extern "C" int64_t lpcl_small_scratch_bytes(const lpcl_scene *s, int64_t scratch_rays)
{
    if (!s || scratch_rays < 0) return 0;
    return align_up((int64_t)s->dev.mesh_count * scratch_rays * 8, 256) + align_up((int64_t)s->dev.mesh_count * scratch_rays * 4, 256);
}

extern "C" int lpcl_trace_small(const lpcl_scene *s, float *d_org_pow, float *d_dir_prev, uint64_t *d_lineage, float *d_dest_meas,
                                int32_t *d_hit, int64_t arena_rays, int64_t n0, void *d_scratch, int64_t scratch_rays,
                                float *d_measured_out, uint64_t *d_measured_lineage, int64_t measured_capacity, float max_ray_len,
                                float ior_env, int32_t max_iterations, double power_threshold, int64_t pair_limit,
                                lpcl_small_iter *d_iters, lpcl_small_header *d_header, void *stream_)
{
    if (!s || !d_org_pow || !d_dir_prev || !d_dest_meas || !d_scratch || !d_iters || !d_header)
        return fail(LPCL_E_NULL, "lpcl_trace_small: NULL argument");
    if (n0 <= 0 || arena_rays < n0 || scratch_rays <= 0 || max_iterations <= 0)
        return fail(LPCL_E_RANGE, "lpcl_trace_small: need n0 > 0, arena_rays >= n0, scratch_rays > 0, max_iterations > 0");
    if (s->dev.mesh_count > SMALL_MAX_MESHES * 1024) return fail(LPCL_E_RANGE, "lpcl_trace_small: too many meshes");
    if (!aligned16(d_org_pow) || !aligned16(d_dir_prev) || !aligned16(d_dest_meas) || (d_hit && !aligned16(d_hit)) ||
        (d_measured_out && !aligned16(d_measured_out)))
        return fail(LPCL_E_ALIGN, "lpcl_trace_small: buffers must be 16-byte aligned");
    unsigned long long *key = (unsigned long long *)d_scratch;
    int *cnt = (int *)((char *)d_scratch + align_up((int64_t)s->dev.mesh_count * scratch_rays * 8, 256));
    trace_small_kernel<<<1, SMALL_THREADS, 0, (cudaStream_t)stream_>>>(
        s->dev, (float4 *)d_org_pow, (float4 *)d_dir_prev, (unsigned long long *)d_lineage, (float4 *)d_dest_meas, (int4 *)d_hit,
        arena_rays, n0, key, cnt, scratch_rays, (float4 *)d_measured_out, (unsigned long long *)d_measured_lineage,
        d_measured_out ? measured_capacity : (1LL << 62), max_ray_len, ior_env, max_iterations, power_threshold, pair_limit,
        (SmallIter *)d_iters, (SmallHeader *)d_header);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

#ifdef LPCL_TC_TRACE
extern "C" int lpcl_debug_read_trace(void *h_dst, int64_t bytes) // development build only (see lpcl_search_tc.cuh)
{
    cudaDeviceSynchronize();
    CUDA_TRY(cudaMemcpyFromSymbol(h_dst, g_tc_trace, bytes < (int64_t)sizeof(g_tc_trace) ? (size_t)bytes : sizeof(g_tc_trace)));
    return 0;
}
extern "C" int lpcl_debug_read_clk(long long *h_dst4)
{
    cudaDeviceSynchronize();
    CUDA_TRY(cudaMemcpyFromSymbol(h_dst4, g_tc_clk, sizeof(g_tc_clk)));
    return 0;
}
#endif

// ------------------------------------------------------------------------------------------ post-processing
#include "lpcl_project.cuh"

extern "C" int lpcl_project(const float *d_records, int64_t k, const float *h_rot9, int32_t kind, float *d_x, float *d_y,
                            float *d_pwr_cor, void *stream_)
{
    if (!d_records || !d_x || !d_y || !d_pwr_cor) return fail(LPCL_E_NULL, "lpcl_project: NULL argument");
    if (k <= 0) return fail(LPCL_E_RANGE, "lpcl_project: k must be > 0");
    if (kind < 0 || kind > 2) return fail(LPCL_E_RANGE, "lpcl_project: kind must be 0 (xy), 1 (stereographic) or 2 (angular)");
    if (!aligned16(d_records)) return fail(LPCL_E_ALIGN, "lpcl_project: records must be 16-byte aligned");
    Rot3 rot;
    for (int i = 0; i < 9; ++i) rot.r[i] = h_rot9 ? h_rot9[i] : ((i % 4 == 0) ? 1.0f : 0.0f);
    project_kernel<<<(unsigned)((k + 255) / 256), 256, 0, (cudaStream_t)stream_>>>((const float4 *)d_records, k, rot, kind,
                                                                                 d_x, d_y, d_pwr_cor);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int lpcl_bin2d(const float *d_x, const float *d_y, const float *d_w, int64_t k, const double *d_xedges,
                          const double *d_yedges, int32_t points, double denom, double *d_hist, void *stream_)
{
    if (!d_x || !d_y || !d_w || !d_xedges || !d_yedges || !d_hist) return fail(LPCL_E_NULL, "lpcl_bin2d: NULL argument");
    if (k < 0 || points <= 0) return fail(LPCL_E_RANGE, "lpcl_bin2d: k >= 0 and points > 0 required");
    cudaStream_t stream = (cudaStream_t)stream_;
    CUDA_TRY(cudaMemsetAsync(d_hist, 0, sizeof(double) * (size_t)points * points, stream));
    if (k > 0) {
        bin2d_kernel<<<(unsigned)((k + 255) / 256), 256, 0, stream>>>(d_x, d_y, d_w, k, d_xedges, d_yedges, points, denom,
                                                                     d_hist);
        CUDA_TRY(cudaGetLastError());
    }
    return 0;
}
