// narrowphase.cu -- K4a..K4e: contact generation per pair class (what dCollide computes for the
// reference's NearCallback, /root/reference/src/main.c:678).  One kernel per class over that
// class's slice of the pair list, so warps never mix the cheap sphere tests with the 15-axis
// box-box SAT.  Contacts go to k-major slots ([k * stride + pair]) and never leave the device in
// device-resident mode.  Normal convention: from geom 2 into geom 1 (ODE).
//
// Operation order inside every predicate follows ODE's (and the test oracle's) so that contact
// membership rounds identically; the library is compiled with --fmad=false.
#include "dev.cuh"

namespace ob {

struct GeomPose {
    V3 p;
    M3 R;
    float4 d;
};

__device__ __forceinline__ GeomPose load_geom(const GeomArrays &g, int i) {
    GeomPose o;
    o.p = v3(g.pos[i]);
    o.R = load_m3(g.R, i);
    o.d = g.dims[i];
    return o;
}

__device__ __forceinline__ void put_contact(const ContactSlots &cs, int p, int k, V3 pos, float depth, V3 n, int side) {
    cs.pd[(size_t)k * cs.stride + p] = make_float4(pos.x, pos.y, pos.z, depth);
    cs.ns[(size_t)k * cs.stride + p] = make_float4(n.x, n.y, n.z, __int_as_float(side));
}

// ---------------------------------------------------------------- sphere-sphere (dCollideSpheres)
__global__ void __launch_bounds__(256) k_np_sphere_sphere(const BroadCounters *__restrict__ bc, const int2 *__restrict__ pairs,
                                                           GeomArrays g, ContactSlots cs) {
    const int s = bc->class_start[PC_SPHERE_SPHERE], e = bc->class_start[PC_SPHERE_SPHERE + 1];
    for (int p = s + blockIdx.x * blockDim.x + threadIdx.x; p < e; p += gridDim.x * blockDim.x) {
        const int2 pr = pairs[p];
        const float4 p1 = g.pos[pr.x], p2 = g.pos[pr.y];
        const float r1 = g.dims[pr.x].x, r2 = g.dims[pr.y].x;
        const float dx = p1.x - p2.x, dy = p1.y - p2.y, dz = p1.z - p2.z;
        const float d = sqrtf(dx * dx + dy * dy + dz * dz);
        int nc = 0;
        if (!(d > (r1 + r2))) {
            if (d <= 0) {
                put_contact(cs, p, 0, v3(p1), r1 + r2, v3(1.f, 0.f, 0.f), -1);
            } else {
                const float d1 = 1.0f / d;
                const V3 n = v3(dx * d1, dy * d1, dz * d1);
                const float k = 0.5f * (r2 - r1 - d);
                put_contact(cs, p, 0, v3(p1.x + n.x * k, p1.y + n.y * k, p1.z + n.z * k), r1 + r2 - d, n, -1);
            }
            nc = 1;
        }
        cs.nc[p] = nc;
    }
}

// ---------------------------------------------------------------- sphere-box (dCollideSphereBox)
__global__ void __launch_bounds__(256) k_np_sphere_box(const BroadCounters *__restrict__ bc, const int2 *__restrict__ pairs,
                                                        GeomArrays g, ContactSlots cs) {
    const int s = bc->class_start[PC_SPHERE_BOX], e = bc->class_start[PC_SPHERE_BOX + 1];
    for (int pi = s + blockIdx.x * blockDim.x + threadIdx.x; pi < e; pi += gridDim.x * blockDim.x) {
        const int2 pr = pairs[pi];
        const V3 ps = v3(g.pos[pr.x]);
        const float radius = g.dims[pr.x].x;
        const GeomPose bx = load_geom(g, pr.y);
        const V3 p = ps - bx.p;
        float l[3] = {bx.d.x * 0.5f, bx.d.y * 0.5f, bx.d.z * 0.5f}, t[3];
        bool onborder = false;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            t[k] = dot(p, col(bx.R, k));
            if (t[k] < -l[k]) { t[k] = -l[k]; onborder = true; }
            if (t[k] > l[k]) { t[k] = l[k]; onborder = true; }
        }
        int nc = 0;
        if (!onborder) {
            float min_distance = l[0] - fabsf(t[0]);
            int mini = 0;
#pragma unroll
            for (int i = 1; i < 3; i++) {
                float face_distance = l[i] - fabsf(t[i]);
                if (face_distance < min_distance) { min_distance = face_distance; mini = i; }
            }
            float tm = mini == 0 ? t[0] : (mini == 1 ? t[1] : t[2]);
            V3 tmp = v3(0.f, 0.f, 0.f);
            const float sg = (tm > 0) ? 1.0f : -1.0f;
            if (mini == 0) tmp.x = sg; else if (mini == 1) tmp.y = sg; else tmp.z = sg;
            put_contact(cs, pi, 0, ps, min_distance + radius, mul(bx.R, tmp), -1);
            nc = 1;
        } else {
            const V3 q = mul(bx.R, v3(t[0], t[1], t[2]));
            const V3 r = p - q;
            const float depth = radius - sqrtf(dot(r, r));
            if (!(depth < 0)) {
                put_contact(cs, pi, 0, q + bx.p, depth, safe_normalize3(r), -1);
                nc = 1;
            }
        }
        cs.nc[pi] = nc;
    }
}

// ---------------------------------------------------------------- sphere-plane (dCollideSpherePlane)
__global__ void __launch_bounds__(256) k_np_sphere_plane(const BroadCounters *__restrict__ bc, const int2 *__restrict__ pairs,
                                                          GeomArrays g, ContactSlots cs) {
    const int s = bc->class_start[PC_SPHERE_PLANE], e = bc->class_start[PC_SPHERE_PLANE + 1];
    for (int pi = s + blockIdx.x * blockDim.x + threadIdx.x; pi < e; pi += gridDim.x * blockDim.x) {
        const int2 pr = pairs[pi];
        const V3 ps = v3(g.pos[pr.x]);
        const float radius = g.dims[pr.x].x;
        const float4 pl = g.dims[pr.y];
        const V3 n = v3(pl);
        const float k = dot(ps, n);
        const float depth = pl.w - k + radius;
        int nc = 0;
        if (!(depth < 0)) {
            put_contact(cs, pi, 0, v3(ps.x - n.x * radius, ps.y - n.y * radius, ps.z - n.z * radius), depth, n, -1);
            nc = 1;
        }
        cs.nc[pi] = nc;
    }
}

// ---------------------------------------------------------------- box-plane (dCollideBoxPlane)
__global__ void __launch_bounds__(256) k_np_box_plane(const BroadCounters *__restrict__ bc, const int2 *__restrict__ pairs,
                                                       GeomArrays g, ContactSlots cs, int maxc_in) {
    const int s = bc->class_start[PC_BOX_PLANE], e = bc->class_start[PC_BOX_PLANE + 1];
    for (int pi = s + blockIdx.x * blockDim.x + threadIdx.x; pi < e; pi += gridDim.x * blockDim.x) {
        const int2 pr = pairs[pi];
        const GeomPose bx = load_geom(g, pr.x);
        const float4 pl = g.dims[pr.y];
        const V3 n = v3(pl);
        const float side[3] = {bx.d.x, bx.d.y, bx.d.z};
        float A[3], B[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float Q = dot(n, col(bx.R, k));
            A[k] = side[k] * Q;
            B[k] = fabsf(A[k]);
        }
        const float depth = pl.w + 0.5f * (B[0] + B[1] + B[2]) - dot(n, bx.p);
        int ret = 0;
        if (!(depth < 0)) {
            int maxc = maxc_in > 4 ? 4 : maxc_in;
            if (maxc < 1) maxc = 1;
            V3 p = bx.p;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const V3 c = col(bx.R, i);
                const V3 term = v3(0.5f * side[i] * c.x, 0.5f * side[i] * c.y, 0.5f * side[i] * c.z);
                if (A[i] > 0) p = p - term; else p = p + term;
            }
            put_contact(cs, pi, 0, p, depth, n, -1);
            ret = 1;
            V3 cpos[3];
            float cdep[3];
            if (maxc > 1) {
                int first, second;
                if (B[0] < B[1]) {
                    if (B[2] < B[0]) { first = 2; second = (B[0] < B[1]) ? 0 : 1; }
                    else { first = 0; second = (B[1] < B[2]) ? 1 : 2; }
                } else {
                    if (B[2] < B[1]) { first = 2; second = (B[0] < B[1]) ? 0 : 1; }
                    else { first = 1; second = (B[0] < B[2]) ? 0 : 2; }
                }
                for (int sI = 0; sI < 2; sI++) {
                    if (sI == 1 && maxc == 2) break;
                    const int j = sI == 0 ? first : second;
                    const float Bj = j == 0 ? B[0] : (j == 1 ? B[1] : B[2]);
                    const float Aj = j == 0 ? A[0] : (j == 1 ? A[1] : A[2]);
                    const float sj = j == 0 ? side[0] : (j == 1 ? side[1] : side[2]);
                    if (depth - Bj < 0) break;
                    const V3 c = col(bx.R, j);
                    const V3 term = v3(sj * c.x, sj * c.y, sj * c.z);
                    cpos[ret] = (Aj > 0) ? (p + term) : (p - term);
                    cdep[ret] = depth - Bj;
                    put_contact(cs, pi, ret, cpos[ret], cdep[ret], n, -1);
                    ret++;
                }
            }
            if (maxc == 4 && ret == 3) {
                const float d4 = cdep[1] + cdep[2] - depth;
                if (d4 > 0) {
                    const V3 p4 = v3(cpos[1].x + cpos[2].x - p.x, cpos[1].y + cpos[2].y - p.y, cpos[1].z + cpos[2].z - p.z);
                    put_contact(cs, pi, 3, p4, d4, n, -1);
                    ret++;
                }
            }
        }
        cs.nc[pi] = ret;
    }
}

// ---------------------------------------------------------------- box-box (dBoxBox)

__device__ __forceinline__ float d14(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[4] + a[2] * b[8]; }
__device__ __forceinline__ float d41(const float *a, const float *b) { return a[0] * b[0] + a[4] * b[1] + a[8] * b[2]; }
__device__ __forceinline__ float d44(const float *a, const float *b) { return a[0] * b[0] + a[4] * b[4] + a[8] * b[8]; }
__device__ __forceinline__ float d11(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// clip the incident quad against the reference rectangle (intersectRectQuad): <= 8 points
__device__ int clip_rect_quad(const float h[2], const float p[8], float ret[16]) {
    int nq = 4, nr = 0;
    float buffer[16];
    const float *q = p;
    float *r = ret;
    for (int dir = 0; dir <= 1; dir++) {
        for (int sign = -1; sign <= 1; sign += 2) {
            const float *pq = q;
            float *pr = r;
            nr = 0;
            for (int i = nq; i > 0; i--) {
                if (sign * pq[dir] < h[dir]) {
                    pr[0] = pq[0]; pr[1] = pq[1];
                    pr += 2; nr++;
                    if (nr & 8) { q = r; goto done; }
                }
                const float *nextq = (i > 1) ? pq + 2 : q;
                if ((sign * pq[dir] < h[dir]) ^ (sign * nextq[dir] < h[dir])) {
                    pr[1 - dir] = pq[1 - dir] + (nextq[1 - dir] - pq[1 - dir]) / (nextq[dir] - pq[dir]) * (sign * h[dir] - pq[dir]);
                    pr[dir] = sign * h[dir];
                    pr += 2; nr++;
                    if (nr & 8) { q = r; goto done; }
                }
                pq += 2;
            }
            q = r;
            r = (q == ret) ? buffer : ret;
            nq = nr;
        }
    }
done:
    if (q != ret)
        for (int i = 0; i < nr * 2; i++) ret[i] = q[i];
    return nr;
}

// cullPoints: choose m of n clipped points spread in angle around the centroid, i0 first
__device__ void cull_points(int n, const float p[], int m, int i0, int iret[]) {
    float a, cx, cy, q;
    if (n == 1) { cx = p[0]; cy = p[1]; }
    else if (n == 2) { cx = 0.5f * (p[0] + p[2]); cy = 0.5f * (p[1] + p[3]); }
    else {
        a = 0; cx = 0; cy = 0;
        for (int i = 0; i < n - 1; i++) {
            q = p[i * 2] * p[i * 2 + 3] - p[i * 2 + 2] * p[i * 2 + 1];
            a += q;
            cx += q * (p[i * 2] + p[i * 2 + 2]);
            cy += q * (p[i * 2 + 1] + p[i * 2 + 3]);
        }
        q = p[n * 2 - 2] * p[1] - p[0] * p[n * 2 - 1];
        a = 1.0f / (3.0f * (a + q));
        cx = a * (cx + q * (p[n * 2 - 2] + p[0]));
        cy = a * (cy + q * (p[n * 2 - 1] + p[1]));
    }
    float A[8];
    int avail[8];
    for (int i = 0; i < n; i++) { A[i] = atan2f(p[i * 2 + 1] - cy, p[i * 2] - cx); avail[i] = 1; }
    avail[i0] = 0;
    iret[0] = i0;
    const float pi = 3.14159265358979323846f;
    for (int j = 1; j < m; j++) {
        a = (float)j * (2 * pi / m) + A[i0];
        if (a > pi) a -= 2 * pi;
        float maxdiff = 1e9f, diff;
        int best = i0;
        for (int i = 0; i < n; i++) {
            if (avail[i]) {
                diff = fabsf(A[i] - a);
                if (diff > pi) diff = 2 * pi - diff;
                if (diff < maxdiff) { maxdiff = diff; best = i; }
            }
        }
        avail[best] = 0;
        iret[j] = best;
    }
}

struct BoxBoxOut {
    float pos[8][3];
    float dep[8];
    float normal[3];
    int n;
};

// sat_only: stop after the 15-axis test and return 1 when no separating axis exists (the first pass of the two-pass
// kernel below); the arithmetic up to that point is the full function's, so both passes take the same decision.
__device__ int box_box(const float p1[3], const float R1[12], const float side1[3], const float p2[3],
                       const float R2[12], const float side2[3], int maxc_in, BoxBoxOut &out, bool sat_only = false) {
    const float fudge_factor = 1.05f;
    float p[3], pp[3], normalC[3] = {0, 0, 0};
    const float *normalR = nullptr;
    float A[3], B[3], Rm[3][3], Q[3][3], s, s2, l, e1;
    int invert_normal = 0, code = 0;
    float *normal = out.normal;

    for (int k = 0; k < 3; k++) p[k] = p2[k] - p1[k];
    pp[0] = d41(R1 + 0, p); pp[1] = d41(R1 + 1, p); pp[2] = d41(R1 + 2, p);
    for (int k = 0; k < 3; k++) { A[k] = side1[k] * 0.5f; B[k] = side2[k] * 0.5f; }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) { Rm[i][j] = d44(R1 + i, R2 + j); Q[i][j] = fabsf(Rm[i][j]); }
    s = -INFINITY;

#define OB_TST1(expr1, expr2, norm, cc)              \
    e1 = (expr1);                                    \
    s2 = fabsf(e1) - (expr2);                        \
    if (s2 > 0) return 0;                            \
    if (s2 > s) { s = s2; normalR = (norm); invert_normal = (e1 < 0); code = (cc); }

    OB_TST1(pp[0], (A[0] + B[0] * Q[0][0] + B[1] * Q[0][1] + B[2] * Q[0][2]), R1 + 0, 1)
    OB_TST1(pp[1], (A[1] + B[0] * Q[1][0] + B[1] * Q[1][1] + B[2] * Q[1][2]), R1 + 1, 2)
    OB_TST1(pp[2], (A[2] + B[0] * Q[2][0] + B[1] * Q[2][1] + B[2] * Q[2][2]), R1 + 2, 3)
    OB_TST1(d41(R2 + 0, p), (A[0] * Q[0][0] + A[1] * Q[1][0] + A[2] * Q[2][0] + B[0]), R2 + 0, 4)
    OB_TST1(d41(R2 + 1, p), (A[0] * Q[0][1] + A[1] * Q[1][1] + A[2] * Q[2][1] + B[1]), R2 + 1, 5)
    OB_TST1(d41(R2 + 2, p), (A[0] * Q[0][2] + A[1] * Q[1][2] + A[2] * Q[2][2] + B[2]), R2 + 2, 6)
#undef OB_TST1

#define OB_TST2(expr1, expr2, n1, n2, n3, cc)                                \
    e1 = (expr1);                                                           \
    s2 = fabsf(e1) - (expr2);                                               \
    if (s2 > 0) return 0;                                                   \
    l = sqrtf((n1) * (n1) + (n2) * (n2) + (n3) * (n3));                     \
    if (l > 0) {                                                            \
        s2 /= l;                                                            \
        if (s2 * fudge_factor > s) {                                        \
            s = s2; normalR = nullptr;                                      \
            normalC[0] = (n1) / l; normalC[1] = (n2) / l; normalC[2] = (n3) / l; \
            invert_normal = (e1 < 0); code = (cc);                          \
        }                                                                   \
    }

    OB_TST2(pp[2] * Rm[1][0] - pp[1] * Rm[2][0], (A[1] * Q[2][0] + A[2] * Q[1][0] + B[1] * Q[0][2] + B[2] * Q[0][1]), 0, -Rm[2][0], Rm[1][0], 7)
    OB_TST2(pp[2] * Rm[1][1] - pp[1] * Rm[2][1], (A[1] * Q[2][1] + A[2] * Q[1][1] + B[0] * Q[0][2] + B[2] * Q[0][0]), 0, -Rm[2][1], Rm[1][1], 8)
    OB_TST2(pp[2] * Rm[1][2] - pp[1] * Rm[2][2], (A[1] * Q[2][2] + A[2] * Q[1][2] + B[0] * Q[0][1] + B[1] * Q[0][0]), 0, -Rm[2][2], Rm[1][2], 9)
    OB_TST2(pp[0] * Rm[2][0] - pp[2] * Rm[0][0], (A[0] * Q[2][0] + A[2] * Q[0][0] + B[1] * Q[1][2] + B[2] * Q[1][1]), Rm[2][0], 0, -Rm[0][0], 10)
    OB_TST2(pp[0] * Rm[2][1] - pp[2] * Rm[0][1], (A[0] * Q[2][1] + A[2] * Q[0][1] + B[0] * Q[1][2] + B[2] * Q[1][0]), Rm[2][1], 0, -Rm[0][1], 11)
    OB_TST2(pp[0] * Rm[2][2] - pp[2] * Rm[0][2], (A[0] * Q[2][2] + A[2] * Q[0][2] + B[0] * Q[1][1] + B[1] * Q[1][0]), Rm[2][2], 0, -Rm[0][2], 12)
    OB_TST2(pp[1] * Rm[0][0] - pp[0] * Rm[1][0], (A[0] * Q[1][0] + A[1] * Q[0][0] + B[1] * Q[2][2] + B[2] * Q[2][1]), -Rm[1][0], Rm[0][0], 0, 13)
    OB_TST2(pp[1] * Rm[0][1] - pp[0] * Rm[1][1], (A[0] * Q[1][1] + A[1] * Q[0][1] + B[0] * Q[2][2] + B[2] * Q[2][0]), -Rm[1][1], Rm[0][1], 0, 14)
    OB_TST2(pp[1] * Rm[0][2] - pp[0] * Rm[1][2], (A[0] * Q[1][2] + A[1] * Q[0][2] + B[0] * Q[2][1] + B[1] * Q[2][0]), -Rm[1][2], Rm[0][2], 0, 15)
#undef OB_TST2

    if (!code) return 0;
    if (sat_only) return 1;

    if (normalR) { normal[0] = normalR[0]; normal[1] = normalR[4]; normal[2] = normalR[8]; }
    else { normal[0] = d11(R1, normalC); normal[1] = d11(R1 + 4, normalC); normal[2] = d11(R1 + 8, normalC); }
    if (invert_normal) { normal[0] = -normal[0]; normal[1] = -normal[1]; normal[2] = -normal[2]; }
    const float depth = -s;

    if (code > 6) {
        float pa[3], pb[3], sign;
        for (int i = 0; i < 3; i++) pa[i] = p1[i];
        for (int j = 0; j < 3; j++) {
            sign = (d14(normal, R1 + j) > 0) ? 1.0f : -1.0f;
            for (int i = 0; i < 3; i++) pa[i] += sign * A[j] * R1[i * 4 + j];
        }
        for (int i = 0; i < 3; i++) pb[i] = p2[i];
        for (int j = 0; j < 3; j++) {
            sign = (d14(normal, R2 + j) > 0) ? -1.0f : 1.0f;
            for (int i = 0; i < 3; i++) pb[i] += sign * B[j] * R2[i * 4 + j];
        }
        float ua[3], ub[3];
        for (int i = 0; i < 3; i++) ua[i] = R1[((code)-7) / 3 + i * 4];
        for (int i = 0; i < 3; i++) ub[i] = R2[((code)-7) % 3 + i * 4];
        // dLineClosestApproach
        float pd[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
        float uaub = d11(ua, ub);
        float q1 = d11(ua, pd);
        float q2 = -d11(ub, pd);
        float d = 1 - uaub * uaub, alpha, beta;
        if (d <= 0.0001f) { alpha = 0; beta = 0; }
        else {
            d = 1.0f / d;
            alpha = (q1 + uaub * q2) * d;
            beta = (uaub * q1 + q2) * d;
        }
        for (int i = 0; i < 3; i++) pa[i] += ua[i] * alpha;
        for (int i = 0; i < 3; i++) pb[i] += ub[i] * beta;
        for (int i = 0; i < 3; i++) out.pos[0][i] = 0.5f * (pa[i] + pb[i]);
        out.dep[0] = depth;
        return 1;
    }

    const float *Ra, *Rb, *pa, *pb, *Sa, *Sb;
    if (code <= 3) { Ra = R1; Rb = R2; pa = p1; pb = p2; Sa = A; Sb = B; }
    else { Ra = R2; Rb = R1; pa = p2; pb = p1; Sa = B; Sb = A; }

    float normal2[3], nr[3], anr[3];
    if (code <= 3) { normal2[0] = normal[0]; normal2[1] = normal[1]; normal2[2] = normal[2]; }
    else { normal2[0] = -normal[0]; normal2[1] = -normal[1]; normal2[2] = -normal[2]; }
    nr[0] = d41(Rb + 0, normal2); nr[1] = d41(Rb + 1, normal2); nr[2] = d41(Rb + 2, normal2);
    anr[0] = fabsf(nr[0]); anr[1] = fabsf(nr[1]); anr[2] = fabsf(nr[2]);

    int lanr, a1, a2;
    if (anr[1] > anr[0]) {
        if (anr[1] > anr[2]) { a1 = 0; lanr = 1; a2 = 2; }
        else { a1 = 0; a2 = 1; lanr = 2; }
    } else {
        if (anr[0] > anr[2]) { lanr = 0; a1 = 1; a2 = 2; }
        else { a1 = 0; a2 = 1; lanr = 2; }
    }

    float center[3];
    if (nr[lanr] < 0) { for (int i = 0; i < 3; i++) center[i] = pb[i] - pa[i] + Sb[lanr] * Rb[i * 4 + lanr]; }
    else { for (int i = 0; i < 3; i++) center[i] = pb[i] - pa[i] - Sb[lanr] * Rb[i * 4 + lanr]; }

    int codeN, code1, code2;
    codeN = (code <= 3) ? code - 1 : code - 4;
    if (codeN == 0) { code1 = 1; code2 = 2; }
    else if (codeN == 1) { code1 = 0; code2 = 2; }
    else { code1 = 0; code2 = 1; }

    float quad[8], c1, c2, m11, m12, m21, m22;
    c1 = d14(center, Ra + code1);
    c2 = d14(center, Ra + code2);
    m11 = d44(Ra + code1, Rb + a1);
    m12 = d44(Ra + code1, Rb + a2);
    m21 = d44(Ra + code2, Rb + a1);
    m22 = d44(Ra + code2, Rb + a2);
    {
        float k1 = m11 * Sb[a1], k2 = m21 * Sb[a1], k3 = m12 * Sb[a2], k4 = m22 * Sb[a2];
        quad[0] = c1 - k1 - k3; quad[1] = c2 - k2 - k4;
        quad[2] = c1 - k1 + k3; quad[3] = c2 - k2 + k4;
        quad[4] = c1 + k1 + k3; quad[5] = c2 + k2 + k4;
        quad[6] = c1 + k1 - k3; quad[7] = c2 + k2 - k4;
    }
    float rect[2] = {Sa[code1], Sa[code2]};
    float ret[16];
    int n = clip_rect_quad(rect, quad, ret);
    if (n < 1) return 0;

    float point[3 * 8], dep[8];
    float det1 = 1.0f / (m11 * m22 - m12 * m21);
    m11 *= det1; m12 *= det1; m21 *= det1; m22 *= det1;
    int cnum = 0;
    for (int j = 0; j < n; j++) {
        float k1 = m22 * (ret[j * 2] - c1) - m12 * (ret[j * 2 + 1] - c2);
        float k2 = -m21 * (ret[j * 2] - c1) + m11 * (ret[j * 2 + 1] - c2);
        for (int i = 0; i < 3; i++) point[cnum * 3 + i] = center[i] + k1 * Rb[i * 4 + a1] + k2 * Rb[i * 4 + a2];
        dep[cnum] = Sa[codeN] - d11(normal2, point + cnum * 3);
        if (dep[cnum] >= 0) {
            ret[cnum * 2] = ret[j * 2];
            ret[cnum * 2 + 1] = ret[j * 2 + 1];
            cnum++;
        }
    }
    if (cnum < 1) return 0;

    int maxc = maxc_in;
    if (maxc > cnum) maxc = cnum;
    if (maxc < 1) maxc = 1;

    if (cnum <= maxc) {
        if (code < 4) {
            for (int j = 0; j < cnum; j++) {
                for (int i = 0; i < 3; i++) out.pos[j][i] = point[j * 3 + i] + pa[i];
                out.dep[j] = dep[j];
            }
        } else {
            for (int j = 0; j < cnum; j++) {
                for (int i = 0; i < 3; i++) out.pos[j][i] = point[j * 3 + i] + pa[i] - normal[i] * dep[j];
                out.dep[j] = dep[j];
            }
        }
    } else {
        int i1 = 0;
        float maxdepth = dep[0];
        for (int i = 1; i < cnum; i++) if (dep[i] > maxdepth) { maxdepth = dep[i]; i1 = i; }
        int iret[8];
        cull_points(cnum, ret, maxc, i1, iret);
        for (int j = 0; j < maxc; j++) {
            for (int i = 0; i < 3; i++) out.pos[j][i] = point[iret[j] * 3 + i] + pa[i];
            out.dep[j] = dep[iret[j]];
        }
        cnum = maxc;
    }
    return cnum;
}

__device__ __forceinline__ void m3_to_arr(const M3 &R, float a[12]) {
    a[0] = R.r0.x; a[1] = R.r0.y; a[2] = R.r0.z; a[3] = 0;
    a[4] = R.r1.x; a[5] = R.r1.y; a[6] = R.r1.z; a[7] = 0;
    a[8] = R.r2.x; a[9] = R.r2.y; a[10] = R.r2.z; a[11] = 0;
}

// Two passes.  ncu had the one-pass kernel at 5.8 of 32 active lanes per instruction on the 1 M-body pile: two thirds of
// the AABB-overlapping box pairs leave dBoxBox at a separating axis after a few dozen instructions while their warp
// mates go on through rectangle clipping and cullPoints.  Pass 1 runs only the 15-axis test for every pair, writes
// nc = 0 for the separated ones and compacts the others (warp ballot + one atomic per warp; the list's order does not
// matter, every pair writes its own contact slots); pass 2 runs the whole of dBoxBox on the compacted list, so its
// warps are full of pairs that all clip.  Same function, same arithmetic, same contacts.
__device__ __forceinline__ void load_box_pair(const GeomArrays &g, int2 pr, float p1[3], float R1[12], float s1[3], float p2[3],
                                              float R2[12], float s2[3]) {
    const GeomPose b1 = load_geom(g, pr.x), b2 = load_geom(g, pr.y);
    m3_to_arr(b1.R, R1);
    m3_to_arr(b2.R, R2);
    p1[0] = b1.p.x; p1[1] = b1.p.y; p1[2] = b1.p.z; p2[0] = b2.p.x; p2[1] = b2.p.y; p2[2] = b2.p.z;
    s1[0] = b1.d.x; s1[1] = b1.d.y; s1[2] = b1.d.z; s2[0] = b2.d.x; s2[1] = b2.d.y; s2[2] = b2.d.z;
}

__global__ void __launch_bounds__(128) k_np_box_box_sat(const BroadCounters *__restrict__ bc, const int2 *__restrict__ pairs,
                                                         GeomArrays g, ContactSlots cs, int *__restrict__ list, int *__restrict__ n_list) {
    const int s = bc->class_start[PC_BOX_BOX], e = bc->class_start[PC_BOX_BOX + 1];
    const int lane = threadIdx.x & 31;
    const int span = ((e - s) + 31) & ~31; // whole warps enter the loop together (ballot)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < span; i += gridDim.x * blockDim.x) {
        const int pi = s + i;
        bool hit = false;
        if (pi < e) {
            float p1[3], p2[3], s1[3], s2[3], R1[12], R2[12];
            load_box_pair(g, pairs[pi], p1, R1, s1, p2, R2, s2);
            BoxBoxOut out;
            hit = box_box(p1, R1, s1, p2, R2, s2, 8, out, true) != 0;
            if (!hit) cs.nc[pi] = 0;
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(n_list, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (hit) list[base + __popc(m & ((1u << lane) - 1u))] = pi;
        }
    }
}

__global__ void __launch_bounds__(128) k_np_box_box(const int2 *__restrict__ pairs, GeomArrays g, ContactSlots cs, int maxc,
                                                     const int *__restrict__ list, const int *__restrict__ n_list) {
    const int n = *n_list;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int pi = list[i];
        float p1[3], p2[3], s1[3], s2[3], R1[12], R2[12];
        load_box_pair(g, pairs[pi], p1, R1, s1, p2, R2, s2);
        BoxBoxOut out;
        const int nc = box_box(p1, R1, s1, p2, R2, s2, maxc > 8 ? 8 : maxc, out);
        // dCollideBoxBox: contact normal = -dBoxBox normal
        const V3 nn = v3(-out.normal[0], -out.normal[1], -out.normal[2]);
        for (int k = 0; k < nc; k++) put_contact(cs, pi, k, v3(out.pos[k][0], out.pos[k][1], out.pos[k][2]), out.dep[k], nn, -1);
        cs.nc[pi] = nc;
    }
}

// ---------------------------------------------------------------- pairs without a collider
__global__ void __launch_bounds__(256) k_np_none(const BroadCounters *__restrict__ bc, ContactSlots cs) {
    const int s = bc->class_start[PC_NONE], e = bc->class_start[PC_NONE + 1];
    for (int p = s + blockIdx.x * blockDim.x + threadIdx.x; p < e; p += gridDim.x * blockDim.x) cs.nc[p] = 0;
}

// ---------------------------------------------------------------- sphere-trimesh
//
// One warp per (sphere, trimesh) pair; the mesh (vertices + indices) is staged once per CTA in
// shared memory with a TMA bulk copy (cp.async.bulk global->shared completing on an mbarrier)
// when it fits, which the reference's teapot.obj does (8884 triangles, 165 KB).  The warp visits only the
// cells of the mesh's triangle grid (engine.cu, eng_add_mesh) that the collider's box touches: one lane per
// cell fetches the cell's triangle range, a warp scan flattens the ranges, and the lanes then take the listed
// triangles 32 at a time (round 1 walked all 8884 triangles of the teapot for every pair).  A triangle listed in
// several visited cells is taken in exactly one of them (the lowest cell, per axis, of the overlap between its own
// cell range and the query's).  Hits go to a per-warp candidate list; the <= 8 contacts are then chosen by the
// order-independent rule documented in DESIGN.md "sphere-trimesh" (depth desc, triangle index asc,
// duplicates within 1e-3 r dropped) -- so the contacts equal the brute-force walk's bit for bit.

constexpr int TM_WARPS = 8;
constexpr int TM_CAND = 64;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// The warp's candidate list keeps the TM_CAND candidates that come FIRST in the selection order (depth descending,
// triangle index ascending).  While there is room, hits are appended (ballot + prefix); once the list is full every further
// hit replaces the current last-in-order entry if it precedes it.  The greedy selection walks the candidates in that
// order, so when it finds its <= 8 contacts inside the kept ones the result equals the unbounded list's (the oracle's);
// only when it runs out of kept candidates first could a dropped one have mattered -- that case alone is flagged.
template <typename Cand>
__device__ __forceinline__ void cand_push(Cand *mine, int &ncand, bool &dropped, bool hit, const Cand &cd, int lane) {
    const unsigned hm = __ballot_sync(0xffffffffu, hit);
    if (!hm) return;
    const int nh = __popc(hm);
    if (ncand + nh <= TM_CAND) {
        if (hit) mine[ncand + __popc(hm & ((1u << lane) - 1u))] = cd;
        ncand += nh;
        __syncwarp();
        return;
    }
    for (unsigned rest = hm; rest; rest &= rest - 1) { // rare: one hit at a time
        const int src = __ffs(rest) - 1;
        const float d = __shfl_sync(0xffffffffu, cd.depth, src);
        const int t = __shfl_sync(0xffffffffu, cd.tri, src);
        if (ncand < TM_CAND) {
            if (lane == src) mine[ncand] = cd;
            ncand++;
            __syncwarp();
            continue;
        }
        // the kept entry that comes last in the order: smallest depth, largest triangle index among equals
        float wd = INFINITY;
        int wt = -1, wi = -1;
        for (int i = lane; i < TM_CAND; i += 32) {
            const float di = mine[i].depth;
            const int ti = mine[i].tri;
            if (di < wd || (di == wd && ti > wt)) { wd = di; wt = ti; wi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, wd, o);
            const int ot = __shfl_xor_sync(0xffffffffu, wt, o);
            const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
            if (od < wd || (od == wd && ot > wt)) { wd = od; wt = ot; wi = oi; }
        }
        dropped = true;
        if (d > wd || (d == wd && t < wt)) {
            if (lane == src) mine[wi] = cd;
        }
        __syncwarp();
    }
}

__device__ __forceinline__ V3 closest_pt_triangle(V3 p, V3 a, V3 b, V3 c) {
    const V3 ab = b - a, ac = c - a, ap = p - a;
    const float d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0 && d2 <= 0) return a;
    const V3 bp = p - b;
    const float d3 = dot(ab, bp), d4 = dot(ac, bp);
    if (d3 >= 0 && d4 <= d3) return b;
    const float vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) {
        const float v = d1 / (d1 - d3);
        return v3(a.x + v * ab.x, a.y + v * ab.y, a.z + v * ab.z);
    }
    const V3 cp = p - c;
    const float d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0 && d5 <= d6) return c;
    const float vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) {
        const float w = d2 / (d2 - d6);
        return v3(a.x + w * ac.x, a.y + w * ac.y, a.z + w * ac.z);
    }
    const float va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
        const float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        return v3(b.x + w * (c.x - b.x), b.y + w * (c.y - b.y), b.z + w * (c.z - b.z));
    }
    const float denom = 1.0f / (va + vb + vc);
    const float v = vb * denom, w = vc * denom;
    return v3(a.x + ab.x * v + ac.x * w, a.y + ab.y * v + ac.y * w, a.z + ab.z * v + ac.z * w);
}

struct TriCand {
    float depth;
    int tri;
    float qx, qy, qz, nx, ny, nz;
};

__device__ __forceinline__ int mesh_cell_of(const MeshInfo &mesh, float v, int k) {
    const int c = (int)floorf((v - mesh.lo[k]) * mesh.ginv[k]); // the host's arithmetic (eng_add_mesh), float for float
    return min(max(c, 0), mesh.gd[k] - 1);
}

// Warp-cooperative walk over the triangles listed in the grid cells [q0, q1]: calls body(t, a, b, c) with, per lane,
// one triangle (t >= 0) or nothing (t = -1), 32 triangles per call, every triangle whose box overlaps the cells' union
// exactly once.  All lanes call body together (it may use warp votes).
template <typename Body>
__device__ __forceinline__ void for_each_grid_triangle(const MeshInfo &mesh, const float *verts, const int *tris,
                                                       const int q0[3], const int q1[3], int lane, Body &&body) {
    const int ncx = q1[0] - q0[0] + 1, ncy = q1[1] - q0[1] + 1, ncz = q1[2] - q0[2] + 1;
    const int ncell = ncx * ncy * ncz;
    for (int cb = 0; cb < ncell; cb += 32) {
        const int ci = cb + lane;
        const bool valid = ci < ncell;
        const int cx = q0[0] + ci % ncx, cy = q0[1] + (ci / ncx) % ncy, cz = q0[2] + ci / (ncx * ncy);
        int s0 = 0, cnt = 0;
        if (valid) {
            const int cid = (cz * mesh.gd[1] + cy) * mesh.gd[0] + cx;
            s0 = mesh.cell_start[cid];
            cnt = mesh.cell_start[cid + 1] - s0;
        }
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        const int total = __shfl_sync(0xffffffffu, inc, 31);
        const int excl = inc - cnt;
        for (int k0 = 0; k0 < total; k0 += 32) {
            const int k = k0 + lane;
            // owner = the last lane whose exclusive prefix is <= k
            int j = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const int cand = j + step;
                const int e = __shfl_sync(0xffffffffu, excl, cand & 31);
                if (cand < 32 && e <= k) j = cand;
            }
            const int os0 = __shfl_sync(0xffffffffu, s0, j), oex = __shfl_sync(0xffffffffu, excl, j);
            const int ox = __shfl_sync(0xffffffffu, cx, j), oy = __shfl_sync(0xffffffffu, cy, j), oz = __shfl_sync(0xffffffffu, cz, j);
            int t = -1;
            V3 a = v3(0.f, 0.f, 0.f), b = a, c = a;
            if (k < total) {
                t = mesh.cell_tris[os0 + (k - oex)];
                const int i0 = tris[3 * t], i1 = tris[3 * t + 1], i2 = tris[3 * t + 2];
                a = v3(verts[3 * i0], verts[3 * i0 + 1], verts[3 * i0 + 2]);
                b = v3(verts[3 * i1], verts[3 * i1 + 1], verts[3 * i1 + 2]);
                c = v3(verts[3 * i2], verts[3 * i2 + 1], verts[3 * i2 + 2]);
                // taken only in the lowest visited cell of the triangle's own cell range
                const int tx = max(mesh_cell_of(mesh, fminf(a.x, fminf(b.x, c.x)), 0), q0[0]);
                const int ty = max(mesh_cell_of(mesh, fminf(a.y, fminf(b.y, c.y)), 1), q0[1]);
                const int tz = max(mesh_cell_of(mesh, fminf(a.z, fminf(b.z, c.z)), 2), q0[2]);
                if (tx != ox || ty != oy || tz != oz) t = -1;
            }
            body(t, a, b, c);
        }
    }
}

__global__ void __launch_bounds__(TM_WARPS * 32) k_np_sphere_trimesh(const BroadCounters *__restrict__ bc,
                                                                      const int2 *__restrict__ pairs, GeomArrays g,
                                                                      MeshInfo mesh, int mesh_id, ContactSlots cs,
                                                                      int maxc, int stage_bytes_v, int stage_bytes_t,
                                                                      StepStats *__restrict__ stats) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ TriCand cand[TM_WARPS][TM_CAND];
    const float *verts = mesh.verts;
    const int *tris = mesh.tris;
    if (stage_bytes_v > 0) {
        // TMA bulk copy of the whole mesh into shared memory, completion tracked by an mbarrier
        float *sv = reinterpret_cast<float *>(smem_raw);
        int *stri = reinterpret_cast<int *>(smem_raw + stage_bytes_v);
        if (threadIdx.x == 0) {
            const uint32_t bar = smem_u32(&mbar);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                         "r"((uint32_t)(stage_bytes_v + stage_bytes_t))
                         : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sv)),
                "l"(mesh.verts), "r"((uint32_t)stage_bytes_v), "r"(bar)
                : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(stri)),
                "l"(mesh.tris), "r"((uint32_t)stage_bytes_t), "r"(bar)
                : "memory");
        }
        __syncthreads();
        {
            const uint32_t bar = smem_u32(&mbar);
            uint32_t done = 0;
            while (!done) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                    : "=r"(done)
                    : "r"(bar), "r"(0u)
                    : "memory");
            }
        }
        verts = sv;
        tris = stri;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int s = bc->class_start[PC_SPHERE_TRIMESH], e = bc->class_start[PC_SPHERE_TRIMESH + 1];
    const unsigned lt = (1u << lane) - 1u;
    TriCand *mine = cand[wid];
    if (maxc > 8) maxc = 8;
    for (int pi = s + blockIdx.x * TM_WARPS + wid; pi < e; pi += gridDim.x * TM_WARPS) {
        const int2 pr = pairs[pi];
        if (g.mesh[pr.y] != mesh_id) continue;
        const V3 ps = v3(g.pos[pr.x]);
        const V3 mp = v3(g.pos[pr.y]);
        const M3 mR = load_m3(g.R, pr.y);
        const V3 c = mulT(mR, ps - mp);
        const bool is_box = g.type[pr.x] == G_BOX;   // warp-uniform: one pair per warp
        float r = g.dims[pr.x].x;                     // sphere radius, or (box) the smallest half side
        int ncand = 0;
        bool dropped = false; // the candidate list was full and something was left out (cand_push)
        if (is_box) {
            // box vs trimesh: vertex/face rule of DESIGN.md "box-trimesh" (restated by the oracle).  Lane = triangle;
            // the 8 box vertices and the 3 triangle vertices are tried as sub-items in lockstep.
            const float4 dm = g.dims[pr.x];
            const float h[3] = {0.5f * dm.x, 0.5f * dm.y, 0.5f * dm.z};
            const M3 bR = load_m3(g.R, pr.x);
            const V3 A[3] = {mulT(mR, v3(bR.r0.x, bR.r1.x, bR.r2.x)), mulT(mR, v3(bR.r0.y, bR.r1.y, bR.r2.y)),
                             mulT(mR, v3(bR.r0.z, bR.r1.z, bR.r2.z))};
            const V3 ext = v3(fabsf(A[0].x) * h[0] + fabsf(A[1].x) * h[1] + fabsf(A[2].x) * h[2],
                              fabsf(A[0].y) * h[0] + fabsf(A[1].y) * h[1] + fabsf(A[2].y) * h[2],
                              fabsf(A[0].z) * h[0] + fabsf(A[1].z) * h[1] + fabsf(A[2].z) * h[2]);
            const float hmin = fminf(h[0], fminf(h[1], h[2]));
            const float maxdepth = 2.0f * hmin;
            r = hmin;
            int q0[3], q1[3];
            {
                const float cl[3] = {c.x, c.y, c.z}, ex[3] = {ext.x, ext.y, ext.z};
                for (int k = 0; k < 3; k++) {
                    const float sl = 1e-5f * (fabsf(cl[k]) + ex[k]) + 1e-6f; // a wider window only costs tests
                    q0[k] = mesh_cell_of(mesh, cl[k] - ex[k] - sl, k);
                    q1[k] = mesh_cell_of(mesh, cl[k] + ex[k] + sl, k);
                }
            }
            for_each_grid_triangle(mesh, verts, tris, q0, q1, lane, [&](int t, V3 a, V3 b, V3 cc) {
                bool live = false;
                V3 nr, n;
                nr = n = v3(0.f, 0.f, 0.f);
                if (t >= 0) {
                    bool skip = false;
                    skip |= (c.x - ext.x > fmaxf(a.x, fmaxf(b.x, cc.x))) || (c.x + ext.x < fminf(a.x, fminf(b.x, cc.x)));
                    skip |= (c.y - ext.y > fmaxf(a.y, fmaxf(b.y, cc.y))) || (c.y + ext.y < fminf(a.y, fminf(b.y, cc.y)));
                    skip |= (c.z - ext.z > fmaxf(a.z, fmaxf(b.z, cc.z))) || (c.z + ext.z < fminf(a.z, fminf(b.z, cc.z)));
                    if (!skip) {
                        nr = cross(b - a, cc - a);
                        const float l2 = dot(nr, nr);
                        if (l2 > 0) {
                            const float inv = 1.0f / sqrtf(l2);
                            n = v3(nr.x * inv, nr.y * inv, nr.z * inv);
                            live = true;
                        }
                    }
                }
                if (!__any_sync(0xffffffffu, live)) return;
                for (int sub = 0; sub < 11; sub++) {
                    bool hit = false;
                    TriCand cd;
                    if (live) {
                        if (sub < 8) {
                            const float s0 = (sub & 1) ? h[0] : -h[0], s1 = (sub & 2) ? h[1] : -h[1], s2 = (sub & 4) ? h[2] : -h[2];
                            const V3 pv = v3(((c.x + s0 * A[0].x) + s1 * A[1].x) + s2 * A[2].x,
                                             ((c.y + s0 * A[0].y) + s1 * A[1].y) + s2 * A[2].y,
                                             ((c.z + s0 * A[0].z) + s1 * A[1].z) + s2 * A[2].z);
                            const float dd = dot(n, pv - a);
                            if (dd < 0 && dd >= -maxdepth) {
                                const V3 pq = v3(pv.x - dd * n.x, pv.y - dd * n.y, pv.z - dd * n.z);
                                bool inside = true;
                                inside = inside && (dot(cross(b - a, pq - a), nr) >= 0);
                                inside = inside && (dot(cross(cc - b, pq - b), nr) >= 0);
                                inside = inside && (dot(cross(a - cc, pq - cc), nr) >= 0);
                                if (inside) {
                                    hit = true;
                                    cd.depth = -dd; cd.tri = t * 16 + sub;
                                    cd.qx = pv.x; cd.qy = pv.y; cd.qz = pv.z;
                                    cd.nx = n.x; cd.ny = n.y; cd.nz = n.z;
                                }
                            }
                        } else {
                            const V3 tv = sub == 8 ? a : (sub == 9 ? b : cc);
                            const V3 e = tv - c;
                            const float loc[3] = {dot(A[0], e), dot(A[1], e), dot(A[2], e)};
                            bool inside = true;
                            int ks = 0;
                            float best = 0.f;
#pragma unroll
                            for (int k = 0; k < 3; k++) {
                                if (!(fabsf(loc[k]) <= h[k])) inside = false;
                                const float pen = h[k] - fabsf(loc[k]);
                                if (k == 0 || pen < best) { best = pen; ks = k; }
                            }
                            if (inside) {
                                const float sg = loc[ks] < 0 ? 1.0f : -1.0f;
                                const V3 ax = ks == 0 ? A[0] : (ks == 1 ? A[1] : A[2]);
                                hit = true;
                                cd.depth = best; cd.tri = t * 16 + sub;
                                cd.qx = tv.x; cd.qy = tv.y; cd.qz = tv.z;
                                cd.nx = sg * ax.x; cd.ny = sg * ax.y; cd.nz = sg * ax.z;
                            }
                        }
                    }
                    cand_push(mine, ncand, dropped, hit, cd, lane);
                }
            });
        } else {
        int q0[3], q1[3];
        {
            const float cl[3] = {c.x, c.y, c.z};
            for (int k = 0; k < 3; k++) {
                const float sl = 1e-5f * (fabsf(cl[k]) + r) + 1e-6f;
                q0[k] = mesh_cell_of(mesh, cl[k] - r - sl, k);
                q1[k] = mesh_cell_of(mesh, cl[k] + r + sl, k);
            }
        }
        for_each_grid_triangle(mesh, verts, tris, q0, q1, lane, [&](int t, V3 a, V3 b, V3 cc) {
            bool hit = false;
            TriCand cd;
            if (t >= 0) {
                bool skip = false;
                skip |= (c.x - r > fmaxf(a.x, fmaxf(b.x, cc.x))) || (c.x + r < fminf(a.x, fminf(b.x, cc.x)));
                skip |= (c.y - r > fmaxf(a.y, fmaxf(b.y, cc.y))) || (c.y + r < fminf(a.y, fminf(b.y, cc.y)));
                skip |= (c.z - r > fmaxf(a.z, fmaxf(b.z, cc.z))) || (c.z + r < fminf(a.z, fminf(b.z, cc.z)));
                if (!skip) {
                    const V3 q = closest_pt_triangle(c, a, b, cc);
                    const V3 dv = c - q;
                    const float d2 = dot(dv, dv);
                    if (!(d2 > r * r)) {
                        const float dist = sqrtf(d2);
                        if (!(dist > r)) {
                            V3 n;
                            bool ok = true;
                            if (dist > 0) {
                                const float inv = 1.0f / dist;
                                n = v3(dv.x * inv, dv.y * inv, dv.z * inv);
                            } else {
                                n = cross(b - a, cc - a);
                                const float l2 = dot(n, n);
                                if (!(l2 > 0)) ok = false;
                                else {
                                    const float inv = 1.0f / sqrtf(l2);
                                    n = v3(n.x * inv, n.y * inv, n.z * inv);
                                }
                            }
                            if (ok) {
                                hit = true;
                                cd.depth = r - dist; cd.tri = t * 16;
                                cd.qx = q.x; cd.qy = q.y; cd.qz = q.z;
                                cd.nx = n.x; cd.ny = n.y; cd.nz = n.z;
                            }
                        }
                    }
                }
            }
            cand_push(mine, ncand, dropped, hit, cd, lane);
        });
        }
        __syncwarp();
        // greedy selection in (depth desc, tri asc) order with duplicate suppression
        int nout = 0;
        const float tol2 = (1e-3f * r) * (1e-3f * r);
        while (nout < maxc) {
            float bd = -1.f;
            int bt = 0x7fffffff, bi = -1;
            for (int i = lane; i < ncand; i += 32) {
                const float d = mine[i].depth;
                const int t = mine[i].tri;
                if (d >= 0.f && (d > bd || (d == bd && t < bt))) { bd = d; bt = t; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, bd, o);
                const int ot = __shfl_xor_sync(0xffffffffu, bt, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (od > bd || (od == bd && ot < bt)) { bd = od; bt = ot; bi = oi; }
            }
            if (bi < 0) {
                // ran out of kept candidates below the contact limit although some were dropped: one of those might have
                // been a contact -- flagged, never silent
                if (dropped && lane == 0) atomicOr(&stats->flags, SF_CAND_OVERFLOW);
                break;
            }
            const TriCand w = mine[bi];
            __syncwarp();
            if (lane == 0) {
                const V3 pw = mul(mR, v3(w.qx, w.qy, w.qz));
                const V3 nw = mul(mR, v3(w.nx, w.ny, w.nz));
                put_contact(cs, pi, nout, pw + mp, w.depth, nw, w.tri >> 4);
            }
            for (int i = lane; i < ncand; i += 32) {
                if (mine[i].depth >= 0.f) {
                    const float ex = mine[i].qx - w.qx, ey = mine[i].qy - w.qy, ez = mine[i].qz - w.qz;
                    if (ex * ex + ey * ey + ez * ez <= tol2) mine[i].depth = -1.f;
                }
            }
            __syncwarp();
            nout++;
        }
        if (lane == 0) cs.nc[pi] = nout;
        __syncwarp();
    }
}

void narrowphase_run(const BroadPhase &bp, GeomArrays g, MeshTable meshes, const std::vector<TriMesh> &host_meshes,
                     ContactSlots cs, int max_contacts, StepStats *d_stats, int num_sms, cudaStream_t st) {
    if (g.n == 0) return;
    // persistent grids: a multiple of the SM count, grid-stride over the class slice
    const unsigned grid = (unsigned)(num_sms * 8);
    // The class kernels work on disjoint slices of the pair list.  While the tick is being captured into a CUDA
    // graph they are forked onto side streams, so the graph runs them as parallel branches next to the long
    // box-box kernel; plain launches stay on the one stream (the fork/join events would cost more than they buy).
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    OB_CUDA(cudaStreamIsCapturing(st, &cap));
    struct Fork { cudaStream_t side[5]; cudaEvent_t fork_ev, join_ev[5]; };
    static Fork forks[16] = {}; // per device (a process may hold worlds on several GPUs)
    int dev = 0;
    OB_CUDA(cudaGetDevice(&dev));
    const bool forked = cap == cudaStreamCaptureStatusActive && dev < 16;
    Fork &fk = forks[dev < 16 ? dev : 0];
    cudaStream_t *side = fk.side;
    cudaEvent_t &fork_ev = fk.fork_ev;
    cudaEvent_t *join_ev = fk.join_ev;
    if (forked && !fork_ev) {
        OB_CUDA(cudaEventCreateWithFlags(&fork_ev, cudaEventDisableTiming));
        for (int i = 0; i < 5; i++) {
            OB_CUDA(cudaStreamCreateWithFlags(&side[i], cudaStreamNonBlocking));
            OB_CUDA(cudaEventCreateWithFlags(&join_ev[i], cudaEventDisableTiming));
        }
    }
    cudaStream_t s1 = st, s2 = st, s3 = st, s4 = st, s5 = st;
    if (forked) {
        OB_CUDA(cudaEventRecord(fork_ev, st));
        for (int i = 0; i < 5; i++) OB_CUDA(cudaStreamWaitEvent(side[i], fork_ev, 0));
        s1 = side[0]; s2 = side[1]; s3 = side[2]; s4 = side[3]; s5 = side[4];
    }
    // two-pass box-box: the compacted list lives in the sweep's (now idle) per-thread hit counters
    k_np_box_box_sat<<<grid, 128, 0, s3>>>(bp.counters, bp.pairs, g, cs, bp.bb_list, &bp.counters->n_bb);
    OB_CHECK_KERNEL("k_np_box_box_sat", s3);
    k_np_box_box<<<grid, 128, 0, s3>>>(bp.pairs, g, cs, max_contacts, bp.bb_list, &bp.counters->n_bb);
    OB_CHECK_KERNEL("k_np_box_box", s3);
    k_np_sphere_sphere<<<grid, 256, 0, s1>>>(bp.counters, bp.pairs, g, cs);
    OB_CHECK_KERNEL("k_np_sphere_sphere", s1);
    k_np_sphere_box<<<grid, 256, 0, s2>>>(bp.counters, bp.pairs, g, cs);
    OB_CHECK_KERNEL("k_np_sphere_box", s2);
    k_np_sphere_plane<<<grid, 256, 0, s3>>>(bp.counters, bp.pairs, g, cs);
    OB_CHECK_KERNEL("k_np_sphere_plane", s3);
    k_np_box_plane<<<grid, 256, 0, s4>>>(bp.counters, bp.pairs, g, cs, max_contacts);
    OB_CHECK_KERNEL("k_np_box_plane", s4);
    k_np_none<<<grid, 256, 0, s5>>>(bp.counters, cs);
    OB_CHECK_KERNEL("k_np_none", s5);
    if (forked)
        for (int i = 0; i < 5; i++) {
            OB_CUDA(cudaEventRecord(join_ev[i], side[i]));
            OB_CUDA(cudaStreamWaitEvent(st, join_ev[i], 0));
        }
    for (int m = 0; m < meshes.n; m++) {
        const TriMesh &hm = host_meshes[m];
        int bv = ((hm.nv * 3 * (int)sizeof(float) + 15) / 16) * 16;
        int bt = ((hm.nt * 3 * (int)sizeof(int) + 15) / 16) * 16;
        int dyn = bv + bt;
        static bool attr_set[64] = {false}; // per device: function attributes live in the device's context
        int dev = 0;
        OB_CUDA(cudaGetDevice(&dev));
        const int max_dyn = 200 * 1024;
        if (dyn > max_dyn) { bv = 0; bt = 0; dyn = 0; } // mesh too large to stage: read through L2
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
            OB_CUDA(cudaFuncSetAttribute(k_np_sphere_trimesh, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn));
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
        k_np_sphere_trimesh<<<(unsigned)num_sms, TM_WARPS * 32, dyn, st>>>(bp.counters, bp.pairs, g, meshes.m[m], m, cs,
                                                                          max_contacts, bv, bt, d_stats);
        OB_CHECK_KERNEL("k_np_sphere_trimesh", st);
    }
}

} // namespace ob
