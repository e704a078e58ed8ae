// lpcl_project.cuh - projection of measured end points and weighted 2-D binning on the GPU.
//
// "Next" row (f)1 of SURVEY.md section 8: the immediate consumer of the measured-ray buffer.
//   project_kernel : the reference's `stereograph_project` (cl:488-506) and `angular_project`
//                    (cl:509-538) kernels (kind 1 / 2) or the plain x,y of get_binned_data (kind 0,
//                    it:413-418), one thread per measured record (x, y, z, power).
//   bin2d_kernel   : np.histogram2d(x, y, bins=points, range=limits, weights=w) as the reference
//                    calls it (it:416,529,560): float64 weights accumulated with atomicAdd(double),
//                    bin membership decided against the SAME float64 edge arrays NumPy builds
//                    (np.linspace), right-most edge inclusive.
// Float32 arithmetic in the reference's operation order (strict __f*_rn), double where OpenCL C
// promotes (the literals 1.0 / 4.0 of cl:501,529).  acos/atan2/sin/cos are CUDA's float
// implementations: OpenCL leaves their last-bit accuracy to the driver, so parity with any given
// libm is a tolerance (1e-6 relative on x, y), not bit-exact.
#pragma once

#include <cuda_runtime.h>

#include "lpcl_math.cuh"

namespace lpcl {

struct Rot3 {
    float r[9]; // row-major rot_mtx[0..2] (the reference uploads a 4x4 with zero last row/column)
};

__global__ void project_kernel(const float4 *rec, long long k, Rot3 rot, int kind, float *x, float *y, float *pw)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    const float4 q = rec[i];
    if (kind == 0) { // get_binned_data: raw x, y, power
        x[i] = q.x;
        y[i] = q.y;
        pw[i] = q.w;
        return;
    }
    const V3 vec = v3(q.x, q.y, q.z); // pivot is the origin (it:511,542)
    const float u = dot(vec, v3(rot.r[0], rot.r[1], rot.r[2]));
    const float v = dot(vec, v3(rot.r[3], rot.r[4], rot.r[5]));
    const float w = dot(vec, v3(rot.r[6], rot.r[7], rot.r[8]));
    const float l = LPCL_SQRT(LPCL_ADD(LPCL_ADD(sq(u), sq(v)), sq(w)));
    if (kind == 1) { // cl:499-505
        const float lw = LPCL_ADD(l, w);
        const float xt = LPCL_DIV(u, lw), yt = LPCL_DIV(v, lw);
        // A = 4.0 / pown(1.0 + pown(xt,2) + pown(yt,2), 2): the sum and its square are double
        const double s = __dadd_rn(__dadd_rn(1.0, (double)sq(xt)), (double)sq(yt));
        const float a = (float)__ddiv_rn(4.0, __dmul_rn(s, s));
        x[i] = xt;
        y[i] = yt;
        pw[i] = LPCL_DIV(q.w, a);
    } else { // cl:519-537
        const float cos_t = LPCL_DIV(w, l); // dot((0,0,1), (u,v,w)/l)
        const float phi = atan2f(v, u);
        const float r = acosf(cos_t);
        const float xt = LPCL_MUL(r, cosf(phi)), yt = LPCL_MUL(r, sinf(phi));
        float a = 1.0f;
        if (r > 0.000001f) a = LPCL_DIV(sinf(r), r);
        x[i] = xt;
        y[i] = yt;
        pw[i] = LPCL_DIV(q.w, a);
    }
}

// np.searchsorted(edges, v, side='right') - 1 with the last edge inclusive; -1 = outside
__device__ __forceinline__ int bin_of(double v, const double *edges, int nb)
{
    const double lo = edges[0], hi = edges[nb];
    if (!(v >= lo) || !(v <= hi)) return -1;
    if (v == hi) return nb - 1;
    int b = (int)((v - lo) / (hi - lo) * nb);
    b = b < 0 ? 0 : (b > nb - 1 ? nb - 1 : b);
    while (b > 0 && v < edges[b]) --b;          // the guess can be off by one at an edge
    while (b < nb - 1 && v >= edges[b + 1]) ++b;
    return b;
}

__global__ void bin2d_kernel(const float *x, const float *y, const float *w, long long k, const double *xe,
                             const double *ye, int nb, double denom, double *hist)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    const int bx = bin_of((double)x[i], xe, nb), by = bin_of((double)y[i], ye, nb);
    if (bx < 0 || by < 0) return;
    atomicAdd(&hist[(long long)bx * nb + by], (double)w[i] / denom); // H[x_bin, y_bin]; w / (dx*dy) as it:553-558
}

} // namespace lpcl
