/* OpenCL-C -> C++ shim used ONLY to compile the reference kernel file
 * (/root/reference/kernel_reflect_refract_intersect.cl) BY PATH into
 * oracle/_ref/ as a CPU checker.  TEST INFRASTRUCTURE - never linked into,
 * imported by, or shipped with the product (lightpycl_b200/).
 *
 * The shim pins the semantics the OpenCL spec leaves open:
 *   - strict IEEE-754 binary32, no FMA contraction (build with
 *     -ffp-contract=off), operations in source order;
 *   - unsuffixed literals are double (C++ rule == OpenCL C with cl_khr_fp64);
 *   - dot(a,b)   = (a.x*b.x + a.y*b.y) + a.z*b.z
 *   - cross(a,b) = (a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x)
 *   - length(a)  = sqrtf(dot(a,a));  normalize(a) = a / length(a)
 *   - pown(x,2)  = x*x (float)
 *   - exp(float) = correctly rounded: (float)exp((double)x)  (ref_shim.cpp maps the kernels'
 *     exp() to it; OpenCL only promises <= 3 ulp, glibc's expf is faithful but not correctly
 *     rounded and has FMA / non-FMA variants, so it cannot serve as a portable definition)
 *   - scalar (float or double) * float3 converts the scalar to float first.
 * float3 occupies 16 bytes in __global memory exactly like OpenCL's float3.
 *
 * `(float3)(0,0,0)` parses in C++ as a cast of the comma expression `0`,
 * i.e. the zero vector - correct for every use on the hot path
 * (reference .cl lines 336, 447, 457, 462).  It is NOT correct for
 * `(float3)(u,v,w)` in the projection kernels (.cl 483, 519-520), so those
 * kernels are never called through this shim.
 */
#ifndef LPCL_ORACLE_CL_SHIM_H
#define LPCL_ORACLE_CL_SHIM_H

#include <math.h>

struct float3 {
    float x, y, z, w;
    float3() : x(0.f), y(0.f), z(0.f), w(0.f) {}
    float3(float a, float b, float c) : x(a), y(b), z(c), w(0.f) {}
    /* cast-constructor used by `(float3)(0,0,0)` (comma expression -> int) */
    float3(int s) : x((float)s), y((float)s), z((float)s), w(0.f) {}
};

static inline float3 operator+(const float3 &a, const float3 &b) { return float3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline float3 operator-(const float3 &a, const float3 &b) { return float3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline float3 operator-(const float3 &a) { return float3(-a.x, -a.y, -a.z); }
static inline float3 operator*(const float3 &a, float s) { return float3(a.x * s, a.y * s, a.z * s); }
static inline float3 operator*(float s, const float3 &a) { return float3(s * a.x, s * a.y, s * a.z); }
static inline float3 operator*(const float3 &a, double s) { return a * (float)s; }
static inline float3 operator*(double s, const float3 &a) { return (float)s * a; }
static inline float3 operator/(const float3 &a, float s) { return float3(a.x / s, a.y / s, a.z / s); }

static inline float dot(const float3 &a, const float3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float3 cross(const float3 &a, const float3 &b) {
    return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float length(const float3 &a) { return sqrtf(dot(a, a)); }
static inline float3 normalize(const float3 &a) { return a / length(a); }
static inline float pown(float x, int n) {
    float r = 1.0f;
    for (int i = 0; i < n; ++i) r = r * x; /* n == 2 everywhere: 1*x*x == x*x exactly */
    return r;
}

/* OpenCL C with cl_khr_fp64 also has double pown(double, int): `pown((1.0 + pown(xt,2) + pown(yt,2)), 2)` in
 * stereograph_project (.cl:501) squares a DOUBLE (the literal 1.0 promotes the sum).  Every pown on
 * the hot path has a float argument and keeps resolving to the float overload above. */
static inline double pown(double x, int n) {
    double r = 1.0;
    for (int i = 0; i < n; ++i) r = r * x;
    return r;
}

extern thread_local int lpcl_shim_gid;
static inline int get_global_id(int) { return lpcl_shim_gid; }

#define __kernel extern "C"
#define __global

#endif
