// dmath.cuh -- float3 helpers for the kernels.  The library is built with --fmad=false, and the
// predicates that decide pair/contact membership evaluate in the same operation order as ODE
// (left to right), so their rounding is reproducible ("box-box bit-exactness").
#pragma once

#include <cuda_runtime.h>
#include <math.h>

namespace ob {

struct V3 {
    float x, y, z;
};

__host__ __device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__host__ __device__ __forceinline__ V3 v3(const float4 &a) { return V3{a.x, a.y, a.z}; }
__host__ __device__ __forceinline__ V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__host__ __device__ __forceinline__ V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__host__ __device__ __forceinline__ V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
__host__ __device__ __forceinline__ V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
__host__ __device__ __forceinline__ V3 operator*(float s, V3 a) { return V3{a.x * s, a.y * s, a.z * s}; }
// ODE dCalcVectorDot3: a0*b0 + a1*b1 + a2*b2, left to right
__host__ __device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// ODE dCalcVectorCross3(a, b, c): a = b x c
__host__ __device__ __forceinline__ V3 cross(V3 b, V3 c) {
    return V3{b.y * c.z - b.z * c.y, b.z * c.x - b.x * c.z, b.x * c.y - b.y * c.x};
}
__host__ __device__ __forceinline__ float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// 3x3 rotation, rows r0 r1 r2 (ODE dMatrix3 without the pad column)
struct M3 {
    V3 r0, r1, r2;
};
__host__ __device__ __forceinline__ V3 col(const M3 &R, int j) {
    return V3{comp(R.r0, j), comp(R.r1, j), comp(R.r2, j)};
}
// dMultiply0_331: R * v
__host__ __device__ __forceinline__ V3 mul(const M3 &R, V3 v) { return V3{dot(R.r0, v), dot(R.r1, v), dot(R.r2, v)}; }
// dMultiply1_331: R^T * v, each component as a column dot (a[0]*b[0] + a[4]*b[1] + a[8]*b[2])
__host__ __device__ __forceinline__ V3 mulT(const M3 &R, V3 v) {
    return V3{R.r0.x * v.x + R.r1.x * v.y + R.r2.x * v.z, R.r0.y * v.x + R.r1.y * v.y + R.r2.y * v.z,
              R.r0.z * v.x + R.r1.z * v.y + R.r2.z * v.z};
}

// dSafeNormalize3 (scale by the largest component first)
__host__ __device__ __forceinline__ V3 safe_normalize3(V3 a) {
    float aa0 = fabsf(a.x), aa1 = fabsf(a.y), aa2 = fabsf(a.z), s;
    if (aa1 > aa0) {
        s = (aa2 > aa1) ? aa2 : aa1;
    } else if (aa2 > aa0) {
        s = aa2;
    } else {
        if (aa0 <= 0) return V3{1, 0, 0};
        s = aa0;
    }
    a.x /= s; a.y /= s; a.z /= s;
    float l = 1.0f / sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
    return V3{a.x * l, a.y * l, a.z * l};
}

// dQtoR
__host__ __device__ __forceinline__ M3 q_to_r(float4 q) { // q = (w,x,y,z) in (x,y,z,w) slots 0..3
    const float q0 = q.x, q1 = q.y, q2 = q.z, q3 = q.w;
    float qq1 = 2 * q1 * q1, qq2 = 2 * q2 * q2, qq3 = 2 * q3 * q3;
    M3 R;
    R.r0 = V3{1 - qq2 - qq3, 2 * (q1 * q2 - q0 * q3), 2 * (q1 * q3 + q0 * q2)};
    R.r1 = V3{2 * (q1 * q2 + q0 * q3), 1 - qq1 - qq3, 2 * (q2 * q3 - q0 * q1)};
    R.r2 = V3{2 * (q1 * q3 - q0 * q2), 2 * (q2 * q3 + q0 * q1), 1 - qq1 - qq2};
    return R;
}

// dPlaneSpace, split in two: the only expensive part is k = 1/sqrt(a) (IEEE sqrt + divide); the
// solver computes it once per contact in the row builder and reuses it every iteration.
__host__ __device__ __forceinline__ float plane_space_k(V3 n) {
    if (fabsf(n.z) > 0.70710678118654752440f) return 1.0f / sqrtf(n.y * n.y + n.z * n.z);
    return 1.0f / sqrtf(n.x * n.x + n.y * n.y);
}
__host__ __device__ __forceinline__ void plane_space_with_k(V3 n, float k, V3 &p, V3 &q) {
    if (fabsf(n.z) > 0.70710678118654752440f) {
        float a = n.y * n.y + n.z * n.z;
        p = V3{0, -n.z * k, n.y * k};
        q = V3{a * k, -n.x * p.z, n.x * p.y};
    } else {
        float a = n.x * n.x + n.y * n.y;
        p = V3{-n.y * k, n.x * k, 0};
        q = V3{-n.z * p.y, n.z * p.x, a * k};
    }
}
__host__ __device__ __forceinline__ void plane_space(V3 n, V3 &p, V3 &q) { plane_space_with_k(n, plane_space_k(n), p, q); }

__device__ __forceinline__ M3 load_m3(const float4 *__restrict__ R, int i) {
    float4 a = R[3 * i], b = R[3 * i + 1], c = R[3 * i + 2];
    return M3{v3(a), v3(b), v3(c)};
}
__device__ __forceinline__ void store_m3(float4 *R, int i, const M3 &m) {
    R[3 * i] = make_float4(m.r0.x, m.r0.y, m.r0.z, 0.f);
    R[3 * i + 1] = make_float4(m.r1.x, m.r1.y, m.r1.z, 0.f);
    R[3 * i + 2] = make_float4(m.r2.x, m.r2.y, m.r2.z, 0.f);
}

// monotone float <-> uint encoding for atomicMin/atomicMax on floats
__host__ __device__ __forceinline__ unsigned f2ord(float f) {
#ifdef __CUDA_ARCH__
    unsigned u = __float_as_uint(f);
#else
    union { float f; unsigned u; } c; c.f = f; unsigned u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(unsigned o) {
    unsigned u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; unsigned u; } c; c.u = u; return c.f;
#endif
}

} // namespace ob
