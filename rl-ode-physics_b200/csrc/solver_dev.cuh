// solver_dev.cuh -- device helpers shared by the solver kernels (solver.cu: global graph-coloured solver and the
// generic island solver; solver_env.cu: the lane-pair island solver of batched worlds): per-body step preparation,
// colouring priorities, one row of dxJointContact::getInfo2 + QuickStep's rhs / Ad, the integrator + snapshot pack.
#pragma once

#include "engine_impl.h"

namespace ob {

constexpr int MODE_MU2 = 0x001, MODE_BOUNCE = 0x004, MODE_SOFT_ERP = 0x008, MODE_SOFT_CFM = 0x010,
              MODE_MOTION1 = 0x020, MODE_MOTION2 = 0x040, MODE_MOTIONN = 0x080, MODE_SLIP1 = 0x100,
              MODE_SLIP2 = 0x200, MODE_APPROX1_1 = 0x1000, MODE_APPROX1_2 = 0x2000;

constexpr int REC_REV = 1 << 8, REC_DYN1 = 1 << 9, REC_DYN2 = 1 << 10;
constexpr int OVERFLOW_COLOUR = 64;

// ------------------------------------------------------------------ per-body step preparation

// world-frame inverse inertia, gyroscopic torque, gravity, v/h + M^-1 f; clears the accumulators
__device__ __forceinline__ void body_prep(int i, const BodyArrays &B, const StepConfig &cfg) {
    const float4 p = B.pos[i];
    const float invM = p.w;
    const M3 R = load_m3(B.R, i);
    const M3 iIb = load_m3(B.invI, i);
    const int flags = B.flags[i];
    // dMultiply2_333(tmp, invI, R): tmp = invI * R^T ; dMultiply0_333(out, R, tmp)
    M3 tmp;
    tmp.r0 = v3(dot(iIb.r0, R.r0), dot(iIb.r0, R.r1), dot(iIb.r0, R.r2));
    tmp.r1 = v3(dot(iIb.r1, R.r0), dot(iIb.r1, R.r1), dot(iIb.r1, R.r2));
    tmp.r2 = v3(dot(iIb.r2, R.r0), dot(iIb.r2, R.r1), dot(iIb.r2, R.r2));
    M3 iIw;
    iIw.r0 = v3(dot(R.r0, col(tmp, 0)), dot(R.r0, col(tmp, 1)), dot(R.r0, col(tmp, 2)));
    iIw.r1 = v3(dot(R.r1, col(tmp, 0)), dot(R.r1, col(tmp, 1)), dot(R.r1, col(tmp, 2)));
    iIw.r2 = v3(dot(R.r2, col(tmp, 0)), dot(R.r2, col(tmp, 1)), dot(R.r2, col(tmp, 2)));
    const float4 lv4 = B.lvel[i], av4 = B.avel[i];
    const V3 lv = v3(lv4), av = v3(av4);
    V3 f = v3(B.facc[i]), t = v3(B.tacc[i]);
    if ((flags & BF_GYRO) && !(flags & BF_KINEMATIC)) {
        const M3 Ib = load_m3(B.I, i);
        M3 t2;
        t2.r0 = v3(dot(Ib.r0, R.r0), dot(Ib.r0, R.r1), dot(Ib.r0, R.r2));
        t2.r1 = v3(dot(Ib.r1, R.r0), dot(Ib.r1, R.r1), dot(Ib.r1, R.r2));
        t2.r2 = v3(dot(Ib.r2, R.r0), dot(Ib.r2, R.r1), dot(Ib.r2, R.r2));
        M3 Iw;
        Iw.r0 = v3(dot(R.r0, col(t2, 0)), dot(R.r0, col(t2, 1)), dot(R.r0, col(t2, 2)));
        Iw.r1 = v3(dot(R.r1, col(t2, 0)), dot(R.r1, col(t2, 1)), dot(R.r1, col(t2, 2)));
        Iw.r2 = v3(dot(R.r2, col(t2, 0)), dot(R.r2, col(t2, 1)), dot(R.r2, col(t2, 2)));
        const V3 L = mul(Iw, av);
        const V3 gt = cross(av, L);
        t = t - gt;
    }
    if (!(flags & BF_NOGRAVITY)) {
        const float mass = lv4.w;
        f.x += mass * cfg.gx; f.y += mass * cfg.gy; f.z += mass * cfg.gz;
    }
    B.facc[i] = make_float4(f.x, f.y, f.z, 0.f);
    B.tacc[i] = make_float4(t.x, t.y, t.z, 0.f);
    const float h1 = 1.0f / cfg.h;
    const V3 it = mul(iIw, t);
    B.tmp[2 * i] = make_float4(f.x * invM + lv.x * h1, f.y * invM + lv.y * h1, f.z * invM + lv.z * h1, 0.f);
    B.tmp[2 * i + 1] = make_float4(it.x + av.x * h1, it.y + av.y * h1, it.z + av.z * h1, 0.f);
    B.inv[3 * i] = make_float4(iIw.r0.x, iIw.r0.y, iIw.r0.z, invM);
    B.inv[3 * i + 1] = make_float4(iIw.r1.x, iIw.r1.y, iIw.r1.z, 0.f);
    B.inv[3 * i + 2] = make_float4(iIw.r2.x, iIw.r2.y, iIw.r2.z, 0.f);
    B.fc[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
    B.fc[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    B.colmask[i] = 0ull;
    B.prio[i] = ~0ull;
}


// ------------------------------------------------------------------ edge colouring

__device__ __forceinline__ unsigned long long manifold_prio(int tie, int lb1, int lb2) {
    // hash of the env-local body pair; the pair index only breaks ties (its relative order inside a
    // world does not depend on the other worlds of a batch)
    unsigned x = ((unsigned)lb1 * 0x9E3779B1u) ^ (((unsigned)lb2 + 0x7F4A7C15u) * 0x85EBCA6Bu);
    x ^= x >> 15; x *= 0x85EBCA77u; x ^= x >> 13; x *= 0xC2B2AE3Du; x ^= x >> 16;
    // 24 hash bits + 32 tie bits: the top byte is left free for k_colour's round stamp
    return ((unsigned long long)(x >> 8) << 32) | (unsigned)tie;
}

// colour choice of a winner: lowest free colour, or -- with spread K > 0 -- the first free colour at or
// after a hashed start within [0, K) (cyclic), falling back to the lowest free colour >= K.  The
// spread rule equalises the colour classes (greedy lowest-first makes the first colours large and the
// last ones tiny, which leaves most lanes of the island solver idle in the late colours).
__device__ __forceinline__ int pick_colour(unsigned long long mask, unsigned long long pr, int K) {
    const unsigned long long fre = ~mask;
    if (fre == 0ull) return OVERFLOW_COLOUR;
    if (K > 0) {
        const unsigned long long low = fre & ((1ull << K) - 1ull);
        if (low) {
            const int start = (int)((pr >> 32) % (unsigned)K);
            const unsigned long long at = low >> start;
            return at ? start + __ffsll((long long)at) - 1 : __ffsll((long long)low) - 1;
        }
    }
    return __ffsll((long long)fre) - 1;
}


// ------------------------------------------------------------------ row build

struct ContactSource {
    const float4 *pd, *ns;
    const Surface *surf; // per contact (compat) or nullptr
    int kstride;         // index = cbase + k * kstride
};

__device__ __forceinline__ int surface_rows(const Surface &s) {
    // dxJointContact::getInfo1
    int m = 1;
    const float mu = s.mu < 0 ? 0 : s.mu;
    if (s.mode & MODE_MU2) {
        const float mu2 = s.mu2 < 0 ? 0 : s.mu2;
        if (mu > 0) m++;
        if (mu2 > 0) m++;
    } else if (mu > 0) m += 2;
    return m;
}

struct BodyKin {
    V3 x, lv, av, tv, tw;
    M3 iI;
    float invM;
};

__device__ __forceinline__ BodyKin load_kin(const BodyArrays &B, int b) {
    BodyKin k;
    const float4 p = B.pos[b];
    k.x = v3(p);
    k.lv = v3(B.lvel[b]);
    k.av = v3(B.avel[b]);
    k.tv = v3(B.tmp[2 * b]);
    k.tw = v3(B.tmp[2 * b + 1]);
    const float4 i0 = B.inv[3 * b], i1 = B.inv[3 * b + 1], i2 = B.inv[3 * b + 2];
    k.iI = M3{v3(i0), v3(i1), v3(i2)};
    k.invM = i0.w;
    return k;
}

// one constraint row of dxJointContact::getInfo2 + QuickStep's rhs / Ad
__device__ __forceinline__ void build_row(V3 dir, V3 c1, V3 c2, const BodyKin &k1, const BodyKin &k2, bool two,
                                          float cval, float cfm, const StepConfig &cfg, float &rhs_s, float &Ad,
                                          float &Adcfm) {
    const float h1 = 1.0f / cfg.h;
    const V3 J1a = cross(c1, dir);
    V3 J2l = v3(0.f, 0.f, 0.f), J2a = v3(0.f, 0.f, 0.f);
    // rhs = c/h - J (v/h + invM fe)
    float sum = 0.f;
    sum += dir.x * k1.tv.x; sum += dir.y * k1.tv.y; sum += dir.z * k1.tv.z;
    sum += J1a.x * k1.tw.x; sum += J1a.y * k1.tw.y; sum += J1a.z * k1.tw.z;
    if (two) {
        J2l = -dir;
        J2a = -cross(c2, dir);
        sum += J2l.x * k2.tv.x; sum += J2l.y * k2.tv.y; sum += J2l.z * k2.tv.z;
        sum += J2a.x * k2.tw.x; sum += J2a.y * k2.tw.y; sum += J2a.z * k2.tw.z;
    }
    const float rhs = cval * h1 - sum;
    const float cfm_h = cfm * h1;
    // Ad = w / (J invM J^T + cfm)
    const V3 iM1l = v3(k1.invM * dir.x, k1.invM * dir.y, k1.invM * dir.z);
    const V3 iM1a = mul(k1.iI, J1a);
    float d = 0.f;
    d += iM1l.x * dir.x; d += iM1l.y * dir.y; d += iM1l.z * dir.z;
    d += iM1a.x * J1a.x; d += iM1a.y * J1a.y; d += iM1a.z * J1a.z;
    if (two) {
        const V3 iM2l = v3(k2.invM * J2l.x, k2.invM * J2l.y, k2.invM * J2l.z);
        const V3 iM2a = mul(k2.iI, J2a);
        d += iM2l.x * J2l.x; d += iM2l.y * J2l.y; d += iM2l.z * J2l.z;
        d += iM2a.x * J2a.x; d += iM2a.y * J2a.y; d += iM2a.z * J2a.z;
    }
    Ad = cfg.sor_w / (d + cfm_h);
    rhs_s = rhs * Ad;
    Adcfm = Ad * cfm_h;
}


// Snapshot record of one body.  Format 0: column-major 4x4 = transpose of R, translation in 12..14
// (GetTransformMat, src/main.c:602-622).  Format 1: the same without its four constant floats (columns of 3, then
// the translation: 48 B).  Format 2: position + quaternion (32 B; dSnapshotExpandB200 rebuilds format 0 on the host
// with the same dQtoR arithmetic).  Records are `snap_stride` floats apart.
__device__ __forceinline__ int snap_stride(int fmt) { return fmt == 0 ? 16 : (fmt == 1 ? 12 : 8); }
__device__ __forceinline__ void snapshot_store(const BodyArrays &B, int i, float4 p, float4 q, const M3 &R) {
    // streaming stores: no kernel reads the snapshot back, it should not displace the solver's rows in L2
    float4 *sn = reinterpret_cast<float4 *>(B.snap + (size_t)snap_stride(B.snap_fmt) * (size_t)i);
    if (B.snap_fmt == 0) {
        __stcs(&sn[0], make_float4(R.r0.x, R.r1.x, R.r2.x, 0.f));
        __stcs(&sn[1], make_float4(R.r0.y, R.r1.y, R.r2.y, 0.f));
        __stcs(&sn[2], make_float4(R.r0.z, R.r1.z, R.r2.z, 0.f));
        __stcs(&sn[3], make_float4(p.x, p.y, p.z, 1.f));
    } else if (B.snap_fmt == 1) {
        __stcs(&sn[0], make_float4(R.r0.x, R.r1.x, R.r2.x, R.r0.y));
        __stcs(&sn[1], make_float4(R.r1.y, R.r2.y, R.r0.z, R.r1.z));
        __stcs(&sn[2], make_float4(R.r2.z, p.x, p.y, p.z));
    } else {
        __stcs(&sn[0], make_float4(p.x, p.y, p.z, 1.f));
        __stcs(&sn[1], q);
    }
}

// velocity update, dxStepBody (semi-implicit Euler + quaternion renormalisation + dQtoR) and the
// reference's GetTransformMat pack, for one body
__device__ __forceinline__ void integrate_body(int i, const BodyArrays &B, float h, float4 fl, float4 fa) {
    float4 p = B.pos[i];
    float4 lv4 = B.lvel[i], av4 = B.avel[i];
    const float4 f = B.facc[i], t = B.tacc[i];
    const float invM = p.w;
    lv4.x += h * fl.x; lv4.y += h * fl.y; lv4.z += h * fl.z;
    av4.x += h * fa.x; av4.y += h * fa.y; av4.z += h * fa.z;
    lv4.x += h * invM * f.x; lv4.y += h * invM * f.y; lv4.z += h * invM * f.z;
    const float4 i0 = B.inv[3 * i], i1 = B.inv[3 * i + 1], i2 = B.inv[3 * i + 2];
    const V3 th = v3(t.x * h, t.y * h, t.z * h);
    av4.x += dot(v3(i0), th); av4.y += dot(v3(i1), th); av4.z += dot(v3(i2), th);
    p.x += h * lv4.x; p.y += h * lv4.y; p.z += h * lv4.z;
    float4 q = B.quat[i]; // (w,x,y,z) in x,y,z,w slots
    const float q0 = q.x, q1 = q.y, q2 = q.z, q3 = q.w;
    const float dq0 = 0.5f * (-av4.x * q1 - av4.y * q2 - av4.z * q3);
    const float dq1 = 0.5f * (av4.x * q0 + av4.y * q3 - av4.z * q2);
    const float dq2 = 0.5f * (-av4.x * q3 + av4.y * q0 + av4.z * q1);
    const float dq3 = 0.5f * (av4.x * q2 - av4.y * q1 + av4.z * q0);
    q.x = q0 + h * dq0; q.y = q1 + h * dq1; q.z = q2 + h * dq2; q.w = q3 + h * dq3;
    float l = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    if (l > 0) {
        l = 1.0f / sqrtf(l);
        q.x *= l; q.y *= l; q.z *= l; q.w *= l;
    } else {
        q = make_float4(1.f, 0.f, 0.f, 0.f);
    }
    const M3 R = q_to_r(q);
    B.pos[i] = p;
    B.lvel[i] = lv4;
    B.avel[i] = av4;
    B.quat[i] = q;
    store_m3(B.R, i, R);
    B.facc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    B.tacc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    snapshot_store(B, i, p, q, R);
}


} // namespace ob
