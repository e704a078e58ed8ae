// lpcl_search_tc.cuh - the nearest-hit search with its filter on the 5th-generation tensor cores.
//
// Same contract as search_kernel (lpcl_search.cuh; reference `intersect`, cl:243-289): per
// (ray, mesh) the nearest accepted t + triangle and the crossing count, merged with atomicMin /
// atomicAdd into the mesh-major scratch.  What changes is who evaluates the conservative filter
// (lpcl_math.cuh, "the same filter evaluated by the tensor cores"):
//
//   D[128 rays][192] = A[128 rays][K = 24] . B[64 triangles x {DEN', aU', aV'}][K = 24]^T
//
// is three kind::tf32 tcgen05.mma (M = 128, N = 192, K = 8) per (ray tile, triangle tile); the 24
// K-slots carry the four partial products of the hi/lo TF32 split of the six ray numbers (M, D)
// and of the triangle's scaled coefficients.  The accumulator lives in TMEM (two 192-column
// buffers, so the MMAs of the next ray tile overlap the epilogue of this one); the epilogue warps
// read it back with tcgen05.ld (lane = ray, 6 consecutive columns = one pair of triangles) and apply
// the sign-agnostic test in its linear form, |DEN'| - max(|aU'|, |aV'|) + k >= 0 (FMNMX + FADD per
// triangle, one FMNMX3 per pair to fold a 16-triangle chunk; the 15 FMAs per test of the SIMT kernel
// are gone).  Only the exact, reference-ordered evaluation decides results; who runs it differs:
//   search_tc_kernel<RT, true> (default)  the sweep keeps ONE sign bit per chunk test in a per-lane flag
//              word - no vote, no branch in the inner loop - and, once per tile, appends flagged (ray, chunk)
//              records to a queue; the records go to global memory and resolve_records_kernel evaluates
//              every triangle of every flagged chunk at full occupancy (a second launch);
//   search_tc_kernel<RT, false>, search_tc16_kernel (LPCL_TC_WARPS=8 / 16, round 2b)  vote after every chunk
//              test, recompute the per-triangle mask in the rare branch, queue (ray, triangle) candidates
//              per warp and drain them inside the kernel.
//
// CTA = 10 warps, one CTA per SM:
//   warps 0-7  epilogue: warp w reads TMEM lanes 32 (w % 4) .. +31 (its rays) and the two
//              16-triangle chunks of half w / 4 of EVERY (ray tile, triangle tile) step through two
//              48-register buffers (the tcgen05.ld of the next step's chunks are in flight while this
//              step's are folded).  The epilogue warps also stage the A operand (RT ray tiles) once per CTA;
//   warp 8     TMA producer: one bulk copy (18 KB) per triangle tile, TC_STAGES deep;
//   warp 9     MMA issuer (one elected lane): waits tile + free accumulator, issues 3 MMAs, commits.
// Shared-memory operands use the no-swizzle K-major canonical layout (8-row x 16-byte core
// matrices): element (row, k) at (k / 4) * rows * 16 + (row / 8) * 128 + (row % 8) * 16 + (k % 4) * 4.
// The triangle tiles are stored in exactly that layout in global memory, so a tile is one 1-D TMA
// bulk copy.  Measured pieces (tools/ubench_tmem.cu, profiles/r1_tc_notes.txt): a 128 x N x 8 tf32
// MMA takes N / 2 cycles (96 here; 4092 flop/clk/SM); tcgen05.ld moves ~43 B/clk per warp whatever
// its shape (310-340 B/clk/SM with 8 warps) and does not contend with the MMAs.
#pragma once

#include "lpcl_search.cuh"

namespace lpcl {

#ifdef LPCL_TC_TRACE
// development build only: clock64 stamps of CTA (0,0)'s epilogue warps 0 and 4 (one SM sub-partition) and of the MMA
// warp at the segment boundaries of steps [TC_TRACE_FIRST, TC_TRACE_FIRST + TC_TRACE_STEPS); read back with
// lpcl_debug_read_trace (tools/kbench.py --trace)
constexpr int TC_TRACE_FIRST = 400, TC_TRACE_STEPS = 48, TC_TRACE_SLOTS = 8;
__device__ long long g_tc_trace[3][TC_TRACE_STEPS][TC_TRACE_SLOTS];
__device__ long long g_tc_clk[32][4]; // CTA (64 k, 0), k < 32: CTA (0,0): clock64 at start / end, globaltimer (ns) at start / end -> the SM clock under load
#define TC_STAMP(who, step, slot)                                                                                     \
    do {                                                                                                              \
        if (trace_on && ((who) == 2 || (warp & 3) == 0) && (step) >= TC_TRACE_FIRST && (step) < TC_TRACE_FIRST + TC_TRACE_STEPS && (threadIdx.x & 31) == 0) \
            g_tc_trace[who][(step) - TC_TRACE_FIRST][slot] = clock64();                                             \
    } while (0)
#else
#define TC_STAMP(who, step, slot) do {} while (0)
#endif

constexpr int TC_TILE_TRIS = 64;                       // triangles per B tile
constexpr int TC_N = 3 * TC_TILE_TRIS;                 // 192 accumulator columns
constexpr int TC_K = 24;                               // 4 partial products x 6
constexpr int TC_M = 128;                              // rays per A tile (= TMEM lanes)
constexpr int TC_TILE_FLOATS = TC_N * TC_K;            // 4608
constexpr int TC_TILE_BYTES = TC_TILE_FLOATS * 4;      // 18432
constexpr int TC_A_FLOATS = TC_M * TC_K;               // 3072
constexpr int TC_A_BYTES = TC_A_FLOATS * 4;            // 12288
constexpr int TC_STAGES = 4;
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_PRODUCER_WARP = 8;
constexpr int TC_MMA_WARP = 9;
constexpr int TC_THREADS = 320;
constexpr int TC_CHUNK_TRIS = 16;                      // triangles per epilogue chunk (48 columns)
constexpr int TC_ACC_STRIDE = 256;                     // TMEM columns between the two accumulators
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

__host__ __device__ inline int tc_canon(int row, int k, int rows) // float index inside an operand tile
{
    return (k >> 2) * rows * 4 + (row >> 3) * 32 + (row & 7) * 4 + (k & 3);
}
// ------------------------------------------------------------------------------------------ scene build
// TileInfo.pad of a tensor-core tile = float bits of Vc, a bound of the vertex norms of ITS triangles:
// the filter margin of the tile is built on Vc instead of the scene-wide Vmax (lpcl_math.cuh,
// filter_ray_scale_tile).  Pass 1: max |v|^2 per tile (non-negative floats order as unsigned ints).
__global__ void tc_tile_vmax_kernel(const float4 *v0, const float4 *v1, const float4 *v2, TileInfo *tiles, int n_tiles)
{
    long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= (long long)n_tiles * TC_TILE_TRIS) return;
    const int tile = (int)(slot / TC_TILE_TRIS), j = (int)(slot % TC_TILE_TRIS);
    if (j >= tiles[tile].count) return;
    const int t = tiles[tile].tri0 + j;
    const float4 a = v0[t], b = v1[t], c = v2[t];
    const float m = fmaxf(fmaxf(a.x * a.x + a.y * a.y + a.z * a.z, b.x * b.x + b.y * b.y + b.z * b.z),
                          c.x * c.x + c.y * c.y + c.z * c.z);
    if (m > 0.0f) atomicMax(reinterpret_cast<unsigned *>(&tiles[tile].pad), __float_as_uint(m));
}
// Pass 2: Vc = sqrt(max |v|^2) rounded up to a power of two (never zero): neighbouring tiles then
// share their Vc, and the sweep rebuilds its per-ray margin constants only when Vc changes
__global__ void tc_tile_vmax_finish_kernel(TileInfo *tiles, int n_tiles)
{
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= n_tiles) return;
    const float v = fmaxf(sqrtf(__int_as_float(tiles[tile].pad)) * 1.0001f, 1e-30f);
    int e;
    const float m = frexpf(v, &e); // v = m 2^e, 0.5 <= m < 1
    tiles[tile].pad = __float_as_int(ldexpf(1.0f, m > 0.5f ? e : e - 1));
}

// one thread per (tile, triangle slot): three rows of 24 floats scattered into the canonical layout
__global__ void build_tc_kernel(const float4 *v0, const float4 *v1, const float4 *v2, const TileInfo *tiles, int n_tiles,
                                float *out)
{
    long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= (long long)n_tiles * TC_TILE_TRIS) return;
    const int tile = (int)(slot / TC_TILE_TRIS), j = (int)(slot % TC_TILE_TRIS);
    const TileInfo ti = tiles[tile];
    float *dst = out + (long long)tile * TC_TILE_FLOATS;
    TcRows r;
    if (j < ti.count) {
        float4 a = v0[ti.tri0 + j], b = v1[ti.tri0 + j], c = v2[ti.tri0 + j];
        r = make_tc_rows(v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), __int_as_float(ti.pad));
    } else {
        // padding slot of a mesh's last tile: rows that fail the filter, DEN' = 0 and |aU'|, |aV'| = 2^40 |a . D| for two
        // fixed directions a (a zero row would PASS - 0 + k >= 0 - and every ray would carry a record for every padded
        // chunk into the second pass).  A ray along (3, -1, -2) still passes; the exact pass looks at valid slots only
        for (int q = 0; q < 3; ++q)
            for (int k = 0; k < 6; ++k) r.c[q][k] = 0.0f;
        const float big = 1099511627776.0f; // 2^40: above any margin k = 1.6e-5 |D| (|O| + 2 Vc), far below overflow
        r.c[1][3] = big, r.c[1][4] = big, r.c[1][5] = big;         // aU' = 2^40 (Dx + Dy + Dz)
        r.c[2][3] = big, r.c[2][4] = -big, r.c[2][5] = 2.0f * big;  // aV' = 2^40 (Dx - Dy + 2 Dz)
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const int row = 6 * (j >> 1) + 2 * q + (j & 1); // pair-interleaved: DEN'(2p,2p+1) uN'(..) vN'(..)
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            float hi, lo;
            tf32_split(r.c[q][k], hi, lo);
            dst[tc_canon(row, k, TC_N)] = hi;       // x A_hi
            dst[tc_canon(row, 6 + k, TC_N)] = lo;   // x A_hi
            dst[tc_canon(row, 12 + k, TC_N)] = hi;  // x A_lo
            dst[tc_canon(row, 18 + k, TC_N)] = lo;  // x A_lo
        }
    }
}

// ------------------------------------------------------------------------------------------ tcgen05 PTX
__device__ __forceinline__ void tc_alloc_512(uint32_t *slot)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_free_512(uint32_t addr)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one lane of the (converged) warp
__device__ __forceinline__ bool tc_elect_one()
{
    unsigned pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// shared-memory matrix descriptor: no swizzle, K-major; LBO = bytes between the two 16-byte K halves
// of one MMA, SBO = bytes between 8-row groups
__device__ __forceinline__ uint64_t tc_smem_desc(unsigned addr, unsigned lbo_bytes, unsigned sbo_bytes)
{
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate,
                                            uint32_t idesc = TC_IDESC)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                 "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
// arrives on the mbarrier when every MMA issued so far by this thread has completed
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#define LPCL_R8(r, o) "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7])
// 48 consecutive accumulator columns of this thread's TMEM lane (SASS: LDTM.x32 + LDTM.x16)
__device__ __forceinline__ void tc_ld_48(uint32_t taddr, uint32_t *r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : LPCL_R8(r, 0), LPCL_R8(r, 8), LPCL_R8(r, 16), LPCL_R8(r, 24)
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : LPCL_R8(r, 32), LPCL_R8(r, 40)
                 : "r"(taddr + 32u)
                 : "memory");
}

// maxima that propagate NaN (an overflowed filter value must become a candidate, never a silent miss)
__device__ __forceinline__ float fmax_nan(float a, float b)
{
    float d;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}
__device__ __forceinline__ float fmax3_nan(float a, float b, float c) // FMNMX3 (sm_100+)
{
    float d;
    asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// ------------------------------------------------------------------------------------------ exact path
// Candidates are not evaluated where they are found: a lane that finds some appends (ray, triangle,
// mesh) records to its warp's queue in shared memory and the sweep goes on.  When the queue is
// half full (and at the end of the sweep) the whole warp drains it, one candidate per lane - the
// exact evaluation runs with 32 active lanes instead of one or two, and its long memory latencies
// no longer sit inside the MMA <-> epilogue hand-shake.  Accepted hits go straight to the scratch
// with atomicAdd / atomicMin; the packed key (t bits, triangle) reproduces "first triangle index
// wins ties" (cl:278) whatever the evaluation order.
constexpr int TC_QUEUE_CAP = 256;   // records per warp
constexpr int TC_QUEUE_DRAIN = 128; // drain threshold, checked once per triangle tile

struct ExactLimits {
    float eps;          // EPSILON = (float)(1e-6 * (double)max_ray_len): hits at t <= eps are ignored (cl:277)
    float max_ray_len;  // hits at t >= max_ray_len are COUNTED but never become the nearest one (cl:252,278)
};

__device__ __forceinline__ void tc_exact_one(const SceneDev &sc, const RaysDev &rays, long long i, int tri, int mesh, ExactLimits lim,
                                             unsigned long long *key, int *cnt)
{
#ifdef LPCL_TC_NOEXACT // timing experiment only: results are wrong
    return;
#endif
    const long long ph = rays.phys(i);
    const float4 o = rays.org_pow[ph], d = rays.dir_prev[ph];
    const float4 a = __ldg(sc.v0 + tri), b = __ldg(sc.v1 + tri), c = __ldg(sc.v2 + tri);
    float t;
    if (ray_triangle_exact(v3(o.x, o.y, o.z), v3(d.x, d.y, d.z), v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), t) &&
        t > lim.eps) {
        const long long slot = (long long)mesh * rays.n + i;
        atomicAdd(&cnt[slot], 1);
        // strict '<' against the running minimum that starts at max_ray_len (cl:278): t == max_ray_len must not
        // replace the initial key even though its packed (t, triangle) value is smaller than (max_ray_len, -1)
        if (t < lim.max_ray_len) atomicMin(&key[slot], pack_key(t, tri));
    }
}

// warp-collective
__device__ __noinline__ void tc_drain(int4 *queue, int *q_count, const SceneDev &sc, const RaysDev &rays, ExactLimits eps,
                                      unsigned long long *key, int *cnt)
{
    __syncwarp();
    const int lane = threadIdx.x & 31;
    const int total = min(*q_count, TC_QUEUE_CAP);
    for (int k = lane; k < total; k += 32) {
        const int4 e = queue[k];
        if (e.y >= 0) {
            queue[k].y = -1; // holes left by an overflowing append stay negative
            tc_exact_one(sc, rays, (long long)(unsigned)e.x, e.y, e.z, eps, key, cnt);
        }
    }
    __syncwarp();
    if (lane == 0) *q_count = 0;
    __syncwarp();
}

// Deferred-flag epilogue (search_tc_kernel<RT, true>): a record is one flagged (ray, 16-triangle chunk) - 8 bytes:
// {ray, 4 * tile + chunk}.  The exact evaluation alone decides (the filter only prunes), so results do not depend on
// which pairs reach it: all (up to 16) triangles of a flagged chunk take it.  Where they take it matters a great deal:
// inside the sweep the exact evaluation is a chain of L2 latencies run by eight warps per SM, and a warp that is busy
// with it cannot release its accumulator, which stops the whole SM within a step or two (measured: 4.5 ms of a 23 ms
// launch on C3, more than the sweep itself on C4).  So the sweep only COLLECTS: a warp's records gather in shared memory
// and are copied to a record buffer in global memory (one atomicAdd per copy) after the sweep, or when the queue is about
// to overflow; resolve_records_kernel then evaluates all of them at full occupancy (0.1 - 0.3 ms for the same records).
// Records that do not fit the buffer are evaluated in place (tc_drain_chunks), so its size is a matter of speed only.
struct RecordBuf {
    int2 *rec;                       // capacity records
    unsigned long long *count;       // records appended so far (may exceed capacity: the excess was evaluated in place)
    long long capacity;
};

__device__ __forceinline__ void tc_exact_chunk_lane(const SceneDev &sc, const RaysDev &rays, int2 e, int t, ExactLimits lim,
                                                    unsigned long long *key, int *cnt)
{
    const int4 ti = __ldg(reinterpret_cast<const int4 *>(sc.tc_tiles) + (e.y >> 2)); // tri0, count, mesh, Vc
    const int first = (e.y & 3) * TC_CHUNK_TRIS + t;
    if (first < ti.y) tc_exact_one(sc, rays, (long long)(unsigned)e.x, ti.x + first, ti.z, lim, key, cnt);
}

// warp-collective: records [from, *q_count) of the warp's queue, sixteen lanes per record
__device__ __noinline__ void tc_drain_chunks(const int2 *queue, int from, int total, const SceneDev &sc, const RaysDev &rays,
                                             ExactLimits lim, unsigned long long *key, int *cnt)
{
    const int lane = threadIdx.x & 31;
    for (int k = from + (lane >> 4); k < total; k += 2) tc_exact_chunk_lane(sc, rays, queue[k], lane & 15, lim, key, cnt);
    __syncwarp();
}

// warp-collective: the warp's queue -> the global record buffer (what does not fit is evaluated here)
__device__ __noinline__ void tc_flush_records(int2 *queue, int *q_count, RecordBuf out, const SceneDev &sc, const RaysDev &rays,
                                              ExactLimits lim, unsigned long long *key, int *cnt)
{
    __syncwarp();
    const int lane = threadIdx.x & 31;
    const int total = *(volatile int *)q_count;
    if (total > 0) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(out.count, (unsigned long long)total);
        base = __shfl_sync(0xffffffffu, base, 0);
        const long long room = out.capacity - (long long)base;
        const int fits = room >= total ? total : (room > 0 ? (int)room : 0);
        for (int k = lane; k < fits; k += 32) out.rec[base + k] = queue[k];
        if (fits < total) tc_drain_chunks(queue, fits, total, sc, rays, lim, key, cnt);
    }
    __syncwarp();
    if (lane == 0) *q_count = 0;
    __syncwarp();
}

// second pass of the deferred-flag search: sixteen threads per record, grid-stride
__global__ void __launch_bounds__(256, 8) resolve_records_kernel(SceneDev sc, RaysDev rays, float max_ray_len, unsigned long long *key,
                                                              int *cnt, RecordBuf in)
{
    const ExactLimits lim = {eps_len(max_ray_len), max_ray_len};
    const unsigned long long appended = *in.count;
    const long long total = appended < (unsigned long long)in.capacity ? (long long)appended : in.capacity;
    const int t = threadIdx.x & 15;
    const long long stride = ((long long)gridDim.x * blockDim.x) >> 4;
    for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4; r < total; r += stride)
        tc_exact_chunk_lane(sc, rays, in.rec[r], t, lim, key, cnt);
}

// queue full (pathological candidate density): evaluate in place
__device__ __noinline__ void tc_overflow(unsigned mask, int tri_base, long long i, int mesh, const SceneDev &sc,
                                         const RaysDev &rays, ExactLimits eps, unsigned long long *key, int *cnt)
{
    while (mask) {
        const int b = __ffs((int)mask) - 1;
        mask &= mask - 1;
        tc_exact_one(sc, rays, i, tri_base + b, mesh, eps, key, cnt);
    }
}

// one lane, inside the divergent rare branch: mask bit b = triangle tri_base + b passed the filter
__device__ __forceinline__ void tc_enqueue(unsigned mask, int tri_base, long long i, int mesh, int4 *queue, int *q_count,
                                           const SceneDev &sc, const RaysDev &rays, ExactLimits eps, unsigned long long *key, int *cnt)
{
    const int c = __popc(mask);
    const int pos = atomicAdd(q_count, c);
    if (pos + c <= TC_QUEUE_CAP) {
        int k = pos;
        while (mask) {
            const int b = __ffs((int)mask) - 1;
            mask &= mask - 1;
            queue[k++] = make_int4((int)i, tri_base + b, mesh, 0);
        }
    } else {
        tc_overflow(mask, tri_base, i, mesh, sc, rays, eps, key, cnt);
    }
}

// ------------------------------------------------------------------------------------------ the sweep
__host__ __device__ constexpr int tc_smem_bytes(int rt)
{
    // padded past half an SM so that two CTAs (two 512-column TMEM allocations) never share an SM
    return (TC_STAGES * TC_TILE_BYTES + rt * TC_A_BYTES + TC_EPI_WARPS * (TC_QUEUE_CAP * 16 + 16) + 256) > 120 * 1024
               ? (TC_STAGES * TC_TILE_BYTES + rt * TC_A_BYTES + TC_EPI_WARPS * (TC_QUEUE_CAP * 16 + 16) + 256)
               : 120 * 1024;
}

// deferred-flag epilogue: 8-byte records, a queue that holds a whole sweep in the common case
__host__ __device__ constexpr int tcf_queue_cap(int rt) { return rt <= 4 ? 1024 : 512; }
__host__ __device__ constexpr int tcf_smem_bytes(int rt)
{
    return TC_STAGES * TC_TILE_BYTES + rt * TC_A_BYTES + TC_EPI_WARPS * (tcf_queue_cap(rt) * 8 + 16) + 256;
}

template <int RT, bool DEFER = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
search_tc_kernel(SceneDev sc, RaysDev rays, float max_ray_len, unsigned long long *__restrict__ key, int *__restrict__ cnt,
                 int tiles_per_chunk, unsigned long long *cand_total, RecordBuf records)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *b_buf = reinterpret_cast<float *>(smem_raw);
    float *a_buf = reinterpret_cast<float *>(smem_raw + TC_STAGES * TC_TILE_BYTES);
    int4 *queues = reinterpret_cast<int4 *>(smem_raw + TC_STAGES * TC_TILE_BYTES + RT * TC_A_BYTES);
    constexpr int QUEUE_BYTES = DEFER ? tcf_queue_cap(RT) * 8 : TC_QUEUE_CAP * 16; // per epilogue warp
    int *q_counts = reinterpret_cast<int *>(reinterpret_cast<unsigned char *>(queues) + TC_EPI_WARPS * QUEUE_BYTES); // one per 16 bytes
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(q_counts + TC_EPI_WARPS * 4);
    uint64_t *empty_bar = full_bar + TC_STAGES;
    uint64_t *tfull_bar = empty_bar + TC_STAGES; // [2] accumulator ready   (MMA commit -> epilogue)
    uint64_t *tempty_bar = tfull_bar + 2;        // [2] accumulator drained (epilogue -> MMA issuer)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty_bar + 2);

    const int tid = threadIdx.x, lane = tid & 31;
    // the warp index through a shuffle: the compiler then knows it is warp-uniform and keeps everything derived
    // from it (TMEM addresses, queue pointers, role branches) in uniform registers (LDTM with a uniform base and
    // immediate offsets instead of two R2UR per load: +20 % together with the elected arrive and the vote below)
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int c0 = blockIdx.y * tiles_per_chunk;
    const int c1 = min(c0 + tiles_per_chunk, sc.tc_n_tiles);
    const int nt = max(c1 - c0, 0);
#ifdef LPCL_TC_TRACE
    const bool trace_on = blockIdx.x == 0 && blockIdx.y == 0;
#endif

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull_bar[b], 1);
            mbar_init(&tempty_bar[b], TC_EPI_WARPS);
        }
        mbar_fence_init();
    }
    if (warp == TC_MMA_WARP) tc_alloc_512(tmem_slot);

    // epilogue threads: the margin constant of each of their rays, the A operand (each ray row is
    // written by one of its two threads) and an empty candidate queue
    const int row = (warp & 3) * 32 + lane; // TMEM lane == ray row of every ray tile
    const int half = (warp >> 2) & 1;
    const long long ray0 = (long long)blockIdx.x * (RT * TC_M) + row;
    const ExactLimits eps = {eps_len(max_ray_len), max_ray_len};
    RayScale rs[RT]; // |O| and 268 u |D| of this thread's ray in each ray tile: the margin constant k is rebuilt when the tile bound Vc changes
    if (warp < TC_EPI_WARPS) {
#pragma unroll
        for (int j = 0; j < RT; ++j) {
            // a row beyond the last ray sweeps a copy of the last ray (its candidates are dropped)
            const long long i = min(ray0 + (long long)j * TC_M, rays.n - 1);
            const long long ph = rays.phys(i);
            const float4 o = rays.org_pow[ph], d = rays.dir_prev[ph];
            rs[j] = filter_ray_parts(v3(o.x, o.y, o.z), v3(d.x, d.y, d.z));
            if ((j & 1) == half) {
                const float a6[6] = {o.y * d.z - o.z * d.y, o.z * d.x - o.x * d.z, o.x * d.y - o.y * d.x, d.x, d.y, d.z};
                float *dst = a_buf + j * TC_A_FLOATS;
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    float hi, lo;
                    tf32_split(a6[k], hi, lo);
                    dst[tc_canon(row, k, TC_M)] = hi;       // x B_hi
                    dst[tc_canon(row, 6 + k, TC_M)] = hi;   // x B_lo
                    dst[tc_canon(row, 12 + k, TC_M)] = lo;  // x B_hi
                    dst[tc_canon(row, 18 + k, TC_M)] = lo;  // x B_lo
                }
            }
        }
        if (!DEFER)
            for (int k = lane; k < TC_QUEUE_CAP; k += 32) queues[warp * TC_QUEUE_CAP + k] = make_int4(0, -1, 0, 0);
        if (lane == 0) q_counts[warp * 4] = 0;
    }
    fence_async_smem(); // generic-proxy writes of A -> visible to the tensor core's async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // The producer and MMA warps run their loops warp-converged and elect the issuing lane with
    // elect.sync: addresses and descriptors then live in uniform registers and each tcgen05.mma is
    // a single UTCHMMA (a plain `if (lane == 0)` makes nvcc wrap every one in an election loop fed
    // by R2UR moves - measured 330 + 300 cycles of issue-thread time per step instead of ~100).
    if (warp == TC_PRODUCER_WARP) {
        for (int it = 0; it < nt; ++it) {
            const int s = it % TC_STAGES;
            if (it >= TC_STAGES) mbar_wait(&empty_bar[s], (unsigned)((it / TC_STAGES - 1) & 1));
            if (tc_elect_one()) {
                mbar_arrive_expect_tx(&full_bar[s], TC_TILE_BYTES);
                tma_load_bulk(b_buf + s * TC_TILE_FLOATS, sc.tc_filter + (long long)(c0 + it) * TC_TILE_FLOATS, TC_TILE_BYTES,
                              &full_bar[s]);
            }
            __syncwarp();
        }
    } else if (warp == TC_MMA_WARP) {
        // descriptors differ only in their start-address field (bits 0-13, 16-byte units)
        const uint64_t a_desc0 = tc_smem_desc(smem_u32(a_buf), TC_M * 16, 128);
        const uint64_t b_desc0 = tc_smem_desc(smem_u32(b_buf), TC_N * 16, 128);
        int g = 0;
        for (int it = 0; it < nt; ++it) {
            const int s = it % TC_STAGES;
            mbar_wait(&full_bar[s], (unsigned)((it / TC_STAGES) & 1));
            tc_fence_after();
            const uint64_t db = b_desc0 + (uint64_t)((s * TC_TILE_BYTES) >> 4);
#pragma unroll
            for (int j = 0; j < RT; ++j, ++g) {
                const int buf = g & 1, use = g >> 1;
                TC_STAMP(2, g, 0);
                if (use > 0) mbar_wait(&tempty_bar[buf], (unsigned)((use - 1) & 1));
                TC_STAMP(2, g, 1);
                tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)(buf * TC_ACC_STRIDE);
                const uint64_t da = a_desc0 + (uint64_t)((j * TC_A_BYTES) >> 4);
                if (tc_elect_one()) {
                    tc_mma_tf32(acc, da, db, 0u);
                    tc_mma_tf32(acc, da + (uint64_t)((2 * TC_M * 16) >> 4), db + (uint64_t)((2 * TC_N * 16) >> 4), 1u);
                    tc_mma_tf32(acc, da + (uint64_t)((4 * TC_M * 16) >> 4), db + (uint64_t)((4 * TC_N * 16) >> 4), 1u);
                    tc_commit(&tfull_bar[buf]);
                    if (j == RT - 1) tc_commit(&empty_bar[s]); // the tile's shared-memory slot is free once its MMAs completed
                }
                __syncwarp();
                TC_STAMP(2, g, 2);
            }
        }
        if (DEFER && nt > 0) { // the epilogue's wait for the step after the last one
            if (tc_elect_one()) tc_commit(&tfull_bar[g & 1]);
            __syncwarp();
        }
    } else if (DEFER) {
        // Deferred-flag epilogue.  The sweep's inner loop has no branch and no vote: a chunk test ends in ONE sign bit
        // (set = no triangle of the chunk passed), funnel-shifted into a per-lane flag word; the word is examined once
        // per 4 steps (8 chunk tests), and only then - rarely - flagged (ray, chunk) records go to the warp's queue.
        // Measured with knock-out builds (profiles/r2c_knockouts.txt): the per-chunk vote + branch of the other
        // epilogue costs ~70 of its 545 cycles per step although it is almost never taken (the branch waits for the end
        // of the test's dependency chain and nothing of the next chunk can issue before it), its inlined candidate
        // code another ~30 (instruction footprint).
        unsigned n_cand = 0;
#ifdef LPCL_TC_TRACE
        if (blockIdx.y == 0 && (blockIdx.x & 63) == 0 && blockIdx.x < 2048 && warp == 0 && lane == 0) {
            g_tc_clk[blockIdx.x >> 6][0] = clock64();
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_tc_clk[blockIdx.x >> 6][2]));
        }
#endif
        constexpr int CAP = tcf_queue_cap(RT);
        int2 *queue = reinterpret_cast<int2 *>(queues) + warp * CAP;
        int *q_count = q_counts + warp * 4;
        const int chunk0 = half * 2;
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(chunk0 * 48);
        uint32_t rA[48], rB[48];
        auto fold = [&](const uint32_t(&r)[48], float kk) -> uint32_t {
            float top[2] = {-3.0e38f, -3.0e38f};
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float fa = fabsf(__uint_as_float(r[6 * q + 0])) - fmax_nan(fabsf(__uint_as_float(r[6 * q + 2])), fabsf(__uint_as_float(r[6 * q + 4])));
                const float fb = fabsf(__uint_as_float(r[6 * q + 1])) - fmax_nan(fabsf(__uint_as_float(r[6 * q + 3])), fabsf(__uint_as_float(r[6 * q + 5])));
                top[q & 1] = fmax3_nan(top[q & 1], fa, fb);
            }
            // sign bit clear = a triangle of the chunk passed, exactly the other kernels' !(x + k < 0): a NaN is the
            // canonical one (sign clear) and passes; -0 cannot occur (|DEN| - m is +0 when they are equal, k >= +0)
            return __float_as_uint(fmax_nan(top[0], top[1]) + kk);
        };
        if (nt > 0) {
            mbar_wait(&tfull_bar[0], 0u);
            tc_fence_after();
            tc_ld_48(lane_addr, rA);
            tc_ld_48(lane_addr + 48u, rB);
        }
        int4 ti_next = nt > 0 ? __ldg(reinterpret_cast<const int4 *>(sc.tc_tiles) + c0) : make_int4(0, 0, 0, 0);
        float k2s[RT];
        int vc_bits = 0;
        unsigned flags = 0xffffffffu;
        constexpr int CHECK = RT < 4 ? RT : 4;              // steps per examination of the flag word
        constexpr unsigned CHECK_MASK = (1u << (2 * CHECK)) - 1u;
        static_assert(32 * 2 * CHECK <= CAP, "one examination must fit the queue");
        for (int it = 0; it < nt; ++it) {
            const int4 ti4 = ti_next;
            if (it + 1 < nt) ti_next = __ldg(reinterpret_cast<const int4 *>(sc.tc_tiles) + (c0 + it + 1));
            const int count = ti4.y;
            if (ti4.w != vc_bits) {
                vc_bits = ti4.w;
                const float two_vc = 2.0f * __int_as_float(vc_bits);
#pragma unroll
                for (int j = 0; j < RT; ++j) k2s[j] = filter_ray_scale_tile(rs[j], two_vc);
            }
#pragma unroll
            for (int j = 0; j < RT; ++j) {
                const int g = it * RT + j, buf = g & 1;
                const float k2 = k2s[j];
                const uint32_t next_addr = lane_addr + (uint32_t)((buf ^ 1) * TC_ACC_STRIDE);
                TC_STAMP(half, g, 0);
                tc_wait_ld(); // both chunks of step g are in registers: hand the accumulator back
                TC_STAMP(half, g, 1);
                tc_fence_before();
                __syncwarp();
                if (tc_elect_one()) mbar_arrive(&tempty_bar[buf]);
                TC_STAMP(half, g, 2);
                // the accumulator of step g + 1 (the MMA warp commits once more after its last step, so that this wait
                // needs no "is there a next step" branch; after the last step the loads below fetch stale columns that
                // nothing reads): with the wait up here nothing separates the two folds from the two loads, and ptxas
                // interleaves them - the load of the next step's first chunk is in flight during the second fold
                mbar_wait(&tfull_bar[buf ^ 1], (unsigned)(((g + 1) >> 1) & 1));
                TC_STAMP(half, g, 3);
                tc_fence_after();
                flags = __funnelshift_l(fold(rA, k2), flags, 1);
                TC_STAMP(half, g, 4);
                tc_ld_48(next_addr, rA);
                TC_STAMP(half, g, 5);
                flags = __funnelshift_l(fold(rB, k2), flags, 1);
                TC_STAMP(half, g, 6);
                tc_ld_48(next_addr + 48u, rB);
                TC_STAMP(half, g, 7);
                if ((j % CHECK) == CHECK - 1) {
                    // bit b of the low 2 CHECK bits: chunk test number 2 CHECK - 1 - b of the last CHECK steps
                    const bool some = (~flags & CHECK_MASK) != 0u;
                    if (__any_sync(0xffffffffu, some)) {
                        if (*(volatile int *)q_count > CAP - 32 * 2 * CHECK) // no room for every lane's records: flush in mid-sweep
                            tc_flush_records(queue, q_count, records, sc, rays, eps, key, cnt);
                        unsigned m = ~flags & CHECK_MASK;
                        // chunks that are all padding (a mesh's last tile) and rows beyond the last ray are dropped
                        const int validA = count - chunk0 * TC_CHUNK_TRIS, validB = validA - TC_CHUNK_TRIS;
                        if (validA <= 0) m &= 0x55555555u; // even test numbers = first chunk = odd bits
                        if (validB <= 0) m &= 0xaaaaaaaau;
                        const int j0 = j - (CHECK - 1);
                        if (m) {
                            unsigned keep = 0;
#pragma unroll
                            for (int q = 0; q < CHECK; ++q)
                                if (ray0 + (long long)(j0 + q) * TC_M < rays.n) keep |= 3u << (2 * (CHECK - 1 - q));
                            m &= keep;
                        }
                        if (m) {
                            int pos = atomicAdd(q_count, __popc(m));
                            n_cand += __popc(m);
                            while (m) {
                                const int b = 31 - __clz((int)m);
                                m &= ~(1u << b);
                                const int tst = 2 * CHECK - 1 - b; // chunk test number: step j0 + tst / 2, chunk chunk0 + tst % 2
                                queue[pos++] = make_int2((int)(ray0 + (long long)(j0 + (tst >> 1)) * TC_M), (c0 + it) * 4 + chunk0 + (tst & 1));
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
        tc_wait_ld();
#ifdef LPCL_TC_TRACE
        if (blockIdx.y == 0 && (blockIdx.x & 63) == 0 && blockIdx.x < 2048 && warp == 0 && lane == 0) {
            g_tc_clk[blockIdx.x >> 6][1] = clock64();
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_tc_clk[blockIdx.x >> 6][3]));
        }
#endif
        tc_flush_records(queue, q_count, records, sc, rays, eps, key, cnt);
        if (cand_total) {
            for (int o = 16; o; o >>= 1) n_cand += __shfl_xor_sync(0xffffffffu, n_cand, o);
            if (lane == 0 && n_cand) atomicAdd(cand_total, (unsigned long long)n_cand);
        }
    } else {
        unsigned n_cand = 0;
        int4 *queue = queues + warp * TC_QUEUE_CAP;
        int *q_count = q_counts + warp * 4;
        const int chunk0 = half * 2; // this warp tests chunks chunk0 and chunk0 + 1 (16 triangles = 48 columns each)
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(chunk0 * 48);
        const int steps = nt * RT;
        uint32_t rA[48], rB[48];

        // the test on 16 triangles x this lane's ray in the linear form (lpcl_math.cuh): per triangle
        //   f = |DEN'| - max(|aU'|, |aV'|)        (FMNMX + FADD, absolute values are operand modifiers)
        // and the chunk's largest f through a three-input FMNMX3 per PAIR of triangles (two independent chains);
        // f + k < 0 = not a candidate.  Every maximum propagates NaN and a NaN passes.  Candidates go to the warp's queue
        auto test_chunk = [&](const uint32_t(&r)[48], float kk, int first, int tri0, int count, int mesh, long long i) {
            float top[2] = {-3.0e38f, -3.0e38f};
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float fa = fabsf(__uint_as_float(r[6 * q + 0])) - fmax_nan(fabsf(__uint_as_float(r[6 * q + 2])), fabsf(__uint_as_float(r[6 * q + 4])));
                const float fb = fabsf(__uint_as_float(r[6 * q + 1])) - fmax_nan(fabsf(__uint_as_float(r[6 * q + 3])), fabsf(__uint_as_float(r[6 * q + 5])));
                top[q & 1] = fmax3_nan(top[q & 1], fa, fb);
            }
            const bool some = !(fmax_nan(top[0], top[1]) + kk < 0.0f); // rare: a triangle of the chunk passed
            if (__any_sync(0xffffffffu, some)) { // warp-uniform branch
                if (some) {
                    unsigned mask = 0; // bit t = triangle `first + t` passed
#pragma unroll
                    for (int t = 0; t < TC_CHUNK_TRIS; ++t) {
                        const int c = 6 * (t >> 1) + (t & 1);
                        const float f = fabsf(__uint_as_float(r[c])) - fmax_nan(fabsf(__uint_as_float(r[c + 2])), fabsf(__uint_as_float(r[c + 4])));
                        if (!(f + kk < 0.0f)) mask |= 1u << t;
                    }
                    const int valid = count - first; // padded triangle slots of a mesh's last tile are not candidates
                    if (valid < TC_CHUNK_TRIS) mask &= valid > 0 ? (1u << valid) - 1u : 0u;
                    n_cand += __popc(mask);
                    if (mask && i < rays.n) tc_enqueue(mask, tri0 + first, i, mesh, queue, q_count, sc, rays, eps, key, cnt);
                }
                __syncwarp(); // reconverge: tcgen05.ld and the barrier hand-shake are warp-collective
            }
        };

        if (steps > 0) { // prologue: step 0 into both buffers
            mbar_wait(&tfull_bar[0], 0u);
            tc_fence_after();
            tc_ld_48(lane_addr, rA);
            tc_ld_48(lane_addr + 48u, rB);
        }
        int4 ti_next = nt > 0 ? __ldg(reinterpret_cast<const int4 *>(sc.tc_tiles) + c0) : make_int4(0, 0, 0, 0);
        float k2s[RT];
        int vc_bits = 0; // Vc is a positive float: the first tile always differs
        for (int it = 0; it < nt; ++it) {
            const int4 ti4 = ti_next; // fetched one tile ahead: the L2 latency is off the critical path
            if (it + 1 < nt) ti_next = __ldg(reinterpret_cast<const int4 *>(sc.tc_tiles) + (c0 + it + 1));
            const int tri0 = ti4.x, count = ti4.y, mesh = ti4.z;
            if (ti4.w != vc_bits) { // K2 of (this thread's ray in each ray tile, tiles with this Vc): rebuilt when Vc changes
                vc_bits = ti4.w;
                const float two_vc = 2.0f * __int_as_float(vc_bits);
#pragma unroll
                for (int j = 0; j < RT; ++j) k2s[j] = filter_ray_scale_tile(rs[j], two_vc);
            }
#pragma unroll
            for (int j = 0; j < RT; ++j) {
                const int g = it * RT + j, buf = g & 1;
                const float k2 = k2s[j];
                const bool more = g + 1 < steps;
                const uint32_t next_addr = lane_addr + (uint32_t)((buf ^ 1) * TC_ACC_STRIDE);
                const long long i = ray0 + (long long)j * TC_M;
                TC_STAMP(half, g, 0);
                tc_wait_ld(); // both chunks of step g are in registers: hand the accumulator back
                TC_STAMP(half, g, 1);
                tc_fence_before();
                __syncwarp();
                if (tc_elect_one()) mbar_arrive(&tempty_bar[buf]);
                if (j == 0 && *(volatile int *)q_count >= TC_QUEUE_DRAIN) // once per triangle tile; an overflowing append evaluates in place
                    tc_drain(queue, q_count, sc, rays, eps, key, cnt);
                TC_STAMP(half, g, 2);
                test_chunk(rA, k2, chunk0 * TC_CHUNK_TRIS, tri0, count, mesh, i);
                TC_STAMP(half, g, 3);
                if (more) { // the first chunk of step g + 1 loads while the second chunk of step g is tested
                    mbar_wait(&tfull_bar[buf ^ 1], (unsigned)(((g + 1) >> 1) & 1));
                    TC_STAMP(half, g, 4);
                    tc_fence_after();
                    tc_ld_48(next_addr, rA);
                }
                TC_STAMP(half, g, 5);
                test_chunk(rB, k2, (chunk0 + 1) * TC_CHUNK_TRIS, tri0, count, mesh, i);
                TC_STAMP(half, g, 6);
                if (more) tc_ld_48(next_addr + 48u, rB);
                TC_STAMP(half, g, 7);
            }
        }
        tc_drain(queue, q_count, sc, rays, eps, key, cnt);
        if (cand_total) {
            for (int o = 16; o; o >>= 1) n_cand += __shfl_xor_sync(0xffffffffu, n_cand, o);
            if (lane == 0 && n_cand) atomicAdd(cand_total, (unsigned long long)n_cand);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == TC_MMA_WARP) tc_free_512(tmem_base);
}


// ------------------------------------------------------------------------------------------ thin-warp variant
// LPCL_TC_WARPS=16: sixteen epilogue warps (four per SM sub-partition, 96-register cap) in TWO GROUPS of eight.  Group h
// owns accumulator h, i.e. every other (ray tile, triangle tile) step; inside a group warp w reads TMEM lanes 32 (w % 4)..
// and tests the two 16-triangle chunks of half (w / 4) % 2 one after the other through a single 48-register buffer
// (load, wait, test, load, wait, release, test - nothing is software pipelined inside a warp; the other group, half a
// period away, fills the gaps).  Same filter arithmetic, queues and exact path as search_tc_kernel, bit-identical
// results.  Measured (profiles/r2b_kbench_variants.log): C3 4.14e12 vs 4.19e12 tests/s for the default kernel, C2 3.34e12
// vs 2.96e12 (+13 %: a sixth of C2's chunk tests take the candidate branch, which stalls a fat warp's whole pipeline),
// C4 2.96e12 vs 2.91e12.  Selectable, not the default (the headline workload is C3).
constexpr int TC16_EPI_WARPS = 16;
constexpr int TC16_PRODUCER_WARP = 16;
constexpr int TC16_MMA_WARP = 17;
constexpr int TC16_THREADS = 576;
constexpr int TC16_QUEUE_CAP = 128;
constexpr int TC16_QUEUE_DRAIN = 64;

__host__ __device__ constexpr int tc16_smem_bytes(int rt)
{
    return (TC_STAGES * TC_TILE_BYTES + rt * TC_A_BYTES + TC16_EPI_WARPS * (TC16_QUEUE_CAP * 16 + 16) + 256) > 120 * 1024
               ? (TC_STAGES * TC_TILE_BYTES + rt * TC_A_BYTES + TC16_EPI_WARPS * (TC16_QUEUE_CAP * 16 + 16) + 256)
               : 120 * 1024;
}

__device__ __noinline__ void tc16_drain(int4 *queue, int *q_count, const SceneDev &sc, const RaysDev &rays, ExactLimits eps,
                                        unsigned long long *key, int *cnt)
{
    __syncwarp();
    const int lane = threadIdx.x & 31;
    const int total = min(*q_count, TC16_QUEUE_CAP);
    for (int k = lane; k < total; k += 32) {
        const int4 e = queue[k];
        if (e.y >= 0) {
            queue[k].y = -1;
            tc_exact_one(sc, rays, (long long)(unsigned)e.x, e.y, e.z, eps, key, cnt);
        }
    }
    __syncwarp();
    if (lane == 0) *q_count = 0;
    __syncwarp();
}

__device__ __forceinline__ void tc16_enqueue(unsigned mask, int tri_base, long long i, int mesh, int4 *queue, int *q_count,
                                             const SceneDev &sc, const RaysDev &rays, ExactLimits eps, unsigned long long *key, int *cnt)
{
    const int c = __popc(mask);
    const int pos = atomicAdd(q_count, c);
    if (pos + c <= TC16_QUEUE_CAP) {
        int k = pos;
        while (mask) {
            const int b = __ffs((int)mask) - 1;
            mask &= mask - 1;
            queue[k++] = make_int4((int)i, tri_base + b, mesh, 0);
        }
    } else {
        tc_overflow(mask, tri_base, i, mesh, sc, rays, eps, key, cnt);
    }
}

template <int RT>
__global__ void __launch_bounds__(TC16_THREADS, 1)
search_tc16_kernel(SceneDev sc, RaysDev rays, float max_ray_len, unsigned long long *__restrict__ key, int *__restrict__ cnt,
                   int tiles_per_chunk, unsigned long long *cand_total)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *b_buf = reinterpret_cast<float *>(smem_raw);
    float *a_buf = reinterpret_cast<float *>(smem_raw + TC_STAGES * TC_TILE_BYTES);
    int4 *queues = reinterpret_cast<int4 *>(smem_raw + TC_STAGES * TC_TILE_BYTES + RT * TC_A_BYTES);
    int *q_counts = reinterpret_cast<int *>(queues + TC16_EPI_WARPS * TC16_QUEUE_CAP); // one per 16 bytes
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(q_counts + TC16_EPI_WARPS * 4);
    uint64_t *empty_bar = full_bar + TC_STAGES;
    uint64_t *tfull_bar = empty_bar + TC_STAGES;
    uint64_t *tempty_bar = tfull_bar + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty_bar + 2);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int c0 = blockIdx.y * tiles_per_chunk;
    const int c1 = min(c0 + tiles_per_chunk, sc.tc_n_tiles);
    const int nt = max(c1 - c0, 0);

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull_bar[b], 1);
            mbar_init(&tempty_bar[b], TC16_EPI_WARPS / 2); // only the accumulator's own group of warps releases it
        }
        mbar_fence_init();
    }
    if (warp == TC16_MMA_WARP) tc_alloc_512(tmem_slot);

    const int row = (warp & 3) * 32 + lane; // TMEM lane == ray row of every ray tile
    const int chunk = (warp >> 2) & 3;      // (A staging only) which of the four threads of a row stages it
    const int group = (warp >> 3) & 1;      // the group owns accumulator `group`, i.e. the steps g = group (mod 2)
    const int sub = (warp >> 2) & 1;        // this warp tests chunks 2 sub and 2 sub + 1 of its group's steps
    const long long ray0 = (long long)blockIdx.x * (RT * TC_M) + row;
    const ExactLimits eps = {eps_len(max_ray_len), max_ray_len};
    if (warp < TC16_EPI_WARPS) {
#pragma unroll
        for (int j = 0; j < RT; ++j) {
            if ((j & 3) == chunk) { // each ray row of each ray tile is staged by one of its four threads
                const long long i = min(ray0 + (long long)j * TC_M, rays.n - 1);
                const long long ph = rays.phys(i);
                const float4 o = rays.org_pow[ph], d = rays.dir_prev[ph];
                const float a6[6] = {o.y * d.z - o.z * d.y, o.z * d.x - o.x * d.z, o.x * d.y - o.y * d.x, d.x, d.y, d.z};
                float *dst = a_buf + j * TC_A_FLOATS;
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    float hi, lo;
                    tf32_split(a6[k], hi, lo);
                    dst[tc_canon(row, k, TC_M)] = hi;
                    dst[tc_canon(row, 6 + k, TC_M)] = hi;
                    dst[tc_canon(row, 12 + k, TC_M)] = lo;
                    dst[tc_canon(row, 18 + k, TC_M)] = lo;
                }
            }
        }
        for (int k = lane; k < TC16_QUEUE_CAP; k += 32) queues[warp * TC16_QUEUE_CAP + k] = make_int4(0, -1, 0, 0);
        if (lane == 0) q_counts[warp * 4] = 0;
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == TC16_PRODUCER_WARP) {
        for (int it = 0; it < nt; ++it) {
            const int s = it % TC_STAGES;
            if (it >= TC_STAGES) mbar_wait(&empty_bar[s], (unsigned)((it / TC_STAGES - 1) & 1));
            if (tc_elect_one()) {
                mbar_arrive_expect_tx(&full_bar[s], TC_TILE_BYTES);
                tma_load_bulk(b_buf + s * TC_TILE_FLOATS, sc.tc_filter + (long long)(c0 + it) * TC_TILE_FLOATS, TC_TILE_BYTES,
                              &full_bar[s]);
            }
            __syncwarp();
        }
    } else if (warp == TC16_MMA_WARP) {
        const uint64_t a_desc0 = tc_smem_desc(smem_u32(a_buf), TC_M * 16, 128);
        const uint64_t b_desc0 = tc_smem_desc(smem_u32(b_buf), TC_N * 16, 128);
        int g = 0;
        for (int it = 0; it < nt; ++it) {
            const int s = it % TC_STAGES;
            mbar_wait(&full_bar[s], (unsigned)((it / TC_STAGES) & 1));
            tc_fence_after();
            const uint64_t db = b_desc0 + (uint64_t)((s * TC_TILE_BYTES) >> 4);
#pragma unroll
            for (int j = 0; j < RT; ++j, ++g) {
                const int buf = g & 1, use = g >> 1;
                if (use > 0) mbar_wait(&tempty_bar[buf], (unsigned)((use - 1) & 1));
                tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)(buf * TC_ACC_STRIDE);
                const uint64_t da = a_desc0 + (uint64_t)((j * TC_A_BYTES) >> 4);
                if (tc_elect_one()) {
                    tc_mma_tf32(acc, da, db, 0u);
                    tc_mma_tf32(acc, da + (uint64_t)((2 * TC_M * 16) >> 4), db + (uint64_t)((2 * TC_N * 16) >> 4), 1u);
                    tc_mma_tf32(acc, da + (uint64_t)((4 * TC_M * 16) >> 4), db + (uint64_t)((4 * TC_N * 16) >> 4), 1u);
                    tc_commit(&tfull_bar[buf]);
                    if (j == RT - 1) tc_commit(&empty_bar[s]);
                }
                __syncwarp();
            }
        }
    } else {
        unsigned n_cand = 0;
        int4 *queue = queues + warp * TC16_QUEUE_CAP;
        int *q_count = q_counts + warp * 4;
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(group * TC_ACC_STRIDE) + (uint32_t)(sub * 96);
        uint32_t r[48];
        int4 ti_next = nt > 0 ? __ldg(reinterpret_cast<const int4 *>(sc.tc_tiles) + c0) : make_int4(0, 0, 0, 0);
        float k2s[RT];
        int vc_bits = 0;
        auto test = [&](float kk, int first, int tri0, int count, int mesh, long long i) {
            float top[2] = {-3.0e38f, -3.0e38f};
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float fa = fabsf(__uint_as_float(r[6 * q + 0])) - fmax_nan(fabsf(__uint_as_float(r[6 * q + 2])), fabsf(__uint_as_float(r[6 * q + 4])));
                const float fb = fabsf(__uint_as_float(r[6 * q + 1])) - fmax_nan(fabsf(__uint_as_float(r[6 * q + 3])), fabsf(__uint_as_float(r[6 * q + 5])));
                top[q & 1] = fmax3_nan(top[q & 1], fa, fb);
            }
            const bool some = !(fmax_nan(top[0], top[1]) + kk < 0.0f);
            if (__any_sync(0xffffffffu, some)) {
                if (some) {
                    unsigned mask = 0;
#pragma unroll
                    for (int t = 0; t < TC_CHUNK_TRIS; ++t) {
                        const int c = 6 * (t >> 1) + (t & 1);
                        const float f = fabsf(__uint_as_float(r[c])) - fmax_nan(fabsf(__uint_as_float(r[c + 2])), fabsf(__uint_as_float(r[c + 4])));
                        if (!(f + kk < 0.0f)) mask |= 1u << t;
                    }
                    const int valid = count - first;
                    if (valid < TC_CHUNK_TRIS) mask &= valid > 0 ? (1u << valid) - 1u : 0u;
                    n_cand += __popc(mask);
                    if (mask && i < rays.n) tc16_enqueue(mask, tri0 + first, i, mesh, queue, q_count, sc, rays, eps, key, cnt);
                }
                __syncwarp();
            }
        };
        static_assert(RT % 2 == 0, "the two groups alternate ray tiles");
        for (int it = 0; it < nt; ++it) {
            const int4 ti4 = ti_next;
            if (it + 1 < nt) ti_next = __ldg(reinterpret_cast<const int4 *>(sc.tc_tiles) + (c0 + it + 1));
            const int tri0 = ti4.x, count = ti4.y, mesh = ti4.z;
            if (ti4.w != vc_bits) {
                vc_bits = ti4.w;
                const float two_vc = 2.0f * __int_as_float(vc_bits);
#pragma unroll
                for (int j = 0; j < RT; ++j) {
                    const long long i = min(ray0 + (long long)j * TC_M, rays.n - 1);
                    const long long ph = rays.phys(i);
                    const float4 o = rays.org_pow[ph], d = rays.dir_prev[ph];
                    k2s[j] = filter_ray_scale_tile(filter_ray_parts(v3(o.x, o.y, o.z), v3(d.x, d.y, d.z)), two_vc);
                }
            }
            if (*(volatile int *)q_count >= TC16_QUEUE_DRAIN) tc16_drain(queue, q_count, sc, rays, eps, key, cnt);
#pragma unroll
            for (int jj = 0; jj < RT / 2; ++jj) {
                // own step g = it RT + 2 jj + group: ray tile j = 2 jj + group (k2s is indexed at compile time for both groups)
                const int g = it * RT + 2 * jj + group;
                const float kk = group ? k2s[2 * jj + 1] : k2s[2 * jj];
                const long long i = ray0 + (long long)(2 * jj + group) * TC_M;
                mbar_wait(&tfull_bar[group], (unsigned)((g >> 1) & 1));
                tc_fence_after();
                tc_ld_48(lane_addr, r);
                tc_wait_ld();
                test(kk, (2 * sub) * TC_CHUNK_TRIS, tri0, count, mesh, i);
                tc_ld_48(lane_addr + 48u, r);
                tc_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (tc_elect_one()) mbar_arrive(&tempty_bar[group]);
                test(kk, (2 * sub + 1) * TC_CHUNK_TRIS, tri0, count, mesh, i);
            }
        }
        tc16_drain(queue, q_count, sc, rays, eps, key, cnt);
        if (cand_total) {
            for (int o = 16; o; o >>= 1) n_cand += __shfl_xor_sync(0xffffffffu, n_cand, o);
            if (lane == 0 && n_cand) atomicAdd(cand_total, (unsigned long long)n_cand);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == TC16_MMA_WARP) tc_free_512(tmem_base);
}

} // namespace lpcl
