// solver.cu -- K5..K8: contact manifolds -> graph colouring -> row build -> SOR/PGS iterations with
// the integrator and snapshot pack fused into the solver tail.
//
// Replaces what libode does inside dWorldStep/dWorldQuickStep (/root/reference/src/main.c:213) for
// the contact joints the reference's NearCallback creates (src/main.c:683-691), plus the app-side
// GetTransformMat pack (src/main.c:602-622, used at :236).
//
// Parallelisation: one thread owns one manifold (all contacts between one geom pair) and solves
// its rows in order normal, tangent 1, tangent 2 per contact.  Manifolds are edge-coloured on the
// body graph so that manifolds of one colour share no dynamic body; colours run one after the
// other separated by grid-wide barriers inside one persistent cooperative kernel.  This is exactly
// a sequential Gauss-Seidel sweep in (colour, manifold, contact, row) order -- the order the test
// oracle is given -- so results match a CPU run of that order bit for bit.
//
// Rows are not stored as 12-float Jacobians: per contact the kernel streams the normal, the two
// lever arms and 9 row scalars (96 B with lambda) and rebuilds J and M^-1 J^T from the bodies'
// world inverse inertia each iteration; arithmetic is free on a kernel that is HBM-bound.
#include <cooperative_groups.h>

#include "solver_dev.cuh"

namespace cg = cooperative_groups;

namespace ob {

// start-of-step resets in one launch: statistics and, for batched worlds, the per-env counters (kernels, not
// memsets/copies: copy-engine work on the compute stream queues behind the application's own transfers)
__global__ void __launch_bounds__(256) k_stats_reset(StepStats *__restrict__ stats, int *__restrict__ env_cnt, int *__restrict__ env_fill,
                                                      int n_env_slots) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        stats->n_rows = 0; stats->n_rows1 = 0; stats->n_rows2 = 0; stats->n_contacts = 0;
        stats->n_manifolds = 0; stats->n_colours = 0; stats->n_overflow = 0; stats->colour_rounds = 0;
        stats->exact_status = -1; stats->n_islands = 0; stats->max_island_rows = 0; stats->pivot_rounds = 0;
        stats->env_trips = 0; stats->env_lanes = 0;
    }
    for (int k = i; k < n_env_slots; k += gridDim.x * blockDim.x) { env_cnt[k] = 0; env_fill[k] = 0; }
}

__global__ void __launch_bounds__(256) k_body_prep(BodyArrays B, StepConfig cfg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B.n) body_prep(i, B, cfg);
}

// ------------------------------------------------------------------ manifolds from device contacts

__global__ void __launch_bounds__(256) k_manifold_flags(const BroadCounters *__restrict__ bc, const int2 *__restrict__ pairs,
                                                         const int *__restrict__ g_body, const int *__restrict__ nc,
                                                         int *__restrict__ flag, int per_contact) {
    const int n = bc->n_pairs;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const int2 pr = pairs[p];
        const int c = nc[p];
        // solver units of this pair: one manifold, or one unit per contact
        flag[p] = (c > 0 && (g_body[pr.x] >= 0 || g_body[pr.y] >= 0)) ? (per_contact ? c : 1) : 0;
    }
}

__global__ void __launch_bounds__(256) k_manifold_write(const BroadCounters *__restrict__ bc, const int2 *__restrict__ pairs,
                                                         const int *__restrict__ g_body, const int *__restrict__ nc,
                                                         const int *__restrict__ scanned, const float4 *__restrict__ b_pos,
                                                         ManifoldArrays M, StepStats *__restrict__ stats, int per_contact,
                                                         int kstride) {
    const int n = bc->n_pairs;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const int2 pr = pairs[p];
        int b1 = g_body[pr.x], b2 = g_body[pr.y];
        const int c = nc[p];
        if (!(c > 0 && (b1 >= 0 || b2 >= 0))) continue;
        const int m = scanned[p];
        const int units = per_contact ? c : 1;
        if (m + units > M.cap) { atomicOr(&stats->flags, SF_MANIFOLD_OVERFLOW); continue; }
        int w = per_contact ? 1 : c;
        if (b1 < 0) { b1 = b2; b2 = -1; w |= REC_REV; } // dJointAttach: NULL body1 swaps, REVERSE
        if (b_pos[b1].w > 0.f) w |= REC_DYN1;
        if (b2 >= 0 && b_pos[b2].w > 0.f) w |= REC_DYN2;
        // .z = index of the unit's first contact in the slot arrays (contact k of pair p: p + k * stride)
        for (int u = 0; u < units; u++) M.rec[m + u] = make_int4(b1, b2, p + u * kstride, w);
    }
}

__global__ void k_manifold_count(const int *__restrict__ total, ManifoldArrays M, StepStats *__restrict__ stats) {
    int n = *total;
    if (n > M.cap) n = M.cap;
    *M.count = n;
    stats->n_manifolds = n;
    M.meta[0] = 0; M.meta[1] = 0; M.meta[2] = 0; M.meta[3] = 0; M.meta[4] = 0; M.meta[7] = 0;
}

// flag DYN bits for host-provided manifold records
__global__ void __launch_bounds__(256) k_manifold_dynbits(ManifoldArrays M, const float4 *__restrict__ b_pos) {
    const int n = *M.count;
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < n; m += gridDim.x * blockDim.x) {
        int4 r = M.rec[m];
        r.w &= ~(REC_DYN1 | REC_DYN2);
        if (b_pos[r.x].w > 0.f) r.w |= REC_DYN1;
        if (r.y >= 0 && b_pos[r.y].w > 0.f) r.w |= REC_DYN2;
        M.rec[m] = r;
    }
}

// grid-wide barrier of the persistent solver (all CTAs are co-resident: cooperative launch).  One
// release-add per CTA on a monotone counter, thread 0 spins with acquire loads.
__device__ __forceinline__ void grid_barrier(unsigned *ctr, unsigned &target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(ctr, 1u);
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

// ------------------------------------------------------------------ edge colouring

// Deterministic parallel greedy colouring: in every round an uncoloured manifold wins if it has
// the smallest hashed priority among the uncoloured manifolds at both of its dynamic bodies; a
// winner takes the lowest colour free at both bodies.  Only kinematic/static ends never conflict.
__global__ void __launch_bounds__(256) k_colour(ManifoldArrays M, BodyArrays B, int spread) {
    // (cg's grid sync, not the solver's own barrier: at this kernel's ~1200 CTAs one spinning thread per CTA on a
    // single counter is slower -- measured on C3: prepare 1.02 -> 1.25 ms)
    cg::grid_group grid = cg::this_grid();
    const int n = *M.count;
    const int gt = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;
    for (int m = gt; m < n; m += gs) M.colour[m] = -1;
    grid.sync();
    int round = 0;
    for (;; round++) {
        // three rotating counters: the one zeroed here was last read two grid barriers ago
        const int slot[3] = {2, 3, 7};
        int *rem_cur = &M.meta[slot[round % 3]], *rem_next = &M.meta[slot[(round + 1) % 3]];
        if (gt == 0) *rem_next = 0;
        // Later rounds stamp their priorities with a smaller top byte, so atomicMin prefers them over
        // whatever earlier rounds left behind and no reset pass (and no third grid barrier) is needed.
        const bool stamped = round < 254;
        const unsigned long long stamp = stamped ? ((unsigned long long)(254 - round) << 56) : 0ull;
        for (int m = gt; m < n; m += gs) {
            if (M.colour[m] >= 0) continue;
            const int4 r = M.rec[m];
            const unsigned long long pr = stamp | manifold_prio(r.z, B.local[r.x], r.y >= 0 ? B.local[r.y] : -1);
            if (r.w & REC_DYN1) atomicMin(&B.prio[r.x], pr);
            if (r.w & REC_DYN2) atomicMin(&B.prio[r.y], pr);
        }
        grid.sync();
        int local = 0;
        for (int m = gt; m < n; m += gs) {
            if (M.colour[m] >= 0) continue;
            const int4 r = M.rec[m];
            const unsigned long long base = manifold_prio(r.z, B.local[r.x], r.y >= 0 ? B.local[r.y] : -1);
            const unsigned long long pr = stamp | base;
            const bool d1 = r.w & REC_DYN1, d2 = r.w & REC_DYN2;
            const bool ok = (!d1 || B.prio[r.x] == pr) && (!d2 || B.prio[r.y] == pr);
            if (ok) {
                unsigned long long mask = 0ull;
                if (d1) mask |= B.colmask[r.x];
                if (d2) mask |= B.colmask[r.y];
                const int c = pick_colour(mask, base, spread);
                if (c == OVERFLOW_COLOUR) {
                    atomicAdd(&M.meta[1], 1);
                } else {
                    const unsigned long long bit = 1ull << c;
                    if (d1) B.colmask[r.x] |= bit;
                    if (d2) B.colmask[r.y] |= bit;
                    atomicMax(&M.meta[0], c + 1);
                }
                M.colour[m] = c;
            } else {
                local++;
            }
        }
        // warp-aggregate the remaining count
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if ((threadIdx.x & 31) == 0 && local) atomicAdd(rem_cur, local);
        grid.sync();
        const int rem = *rem_cur;
        if (rem == 0) break;
        if (!stamped) { // out of stamps (> 254 rounds): fall back to resetting the priorities
            for (int m = gt; m < n; m += gs) {
                if (M.colour[m] >= 0) continue;
                const int4 r = M.rec[m];
                if (r.w & REC_DYN1) B.prio[r.x] = ~0ull;
                if (r.w & REC_DYN2) B.prio[r.y] = ~0ull;
            }
            grid.sync();
        }
    }
    if (gt == 0) M.meta[4] = round + 1;
}

__global__ void __launch_bounds__(256) k_colour_keys(ManifoldArrays M) {
    const int n = *M.count;
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < n; m += gridDim.x * blockDim.x) {
        const int nc = M.rec[m].w & 0xff;
        M.skey[m] = ((uint32_t)M.colour[m] << 3) | (uint32_t)(8 - nc);
        M.sidx[m] = m;
    }
}

__global__ void __launch_bounds__(256) k_colour_bounds(ManifoldArrays M) {
    const int n = *M.count;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const int c = (int)(M.skey[s] >> 3);
        if (s == 0 || (int)(M.skey[s - 1] >> 3) != c) M.colour_start[c] = s;
    }
}

__global__ void k_colour_fixup(ManifoldArrays M, StepStats *__restrict__ stats) {
    const int n = *M.count;
    M.colour_start[OVERFLOW_COLOUR + 1] = n;
    for (int c = OVERFLOW_COLOUR; c >= 0; c--)
        if (M.colour_start[c] < 0) M.colour_start[c] = M.colour_start[c + 1];
    stats->n_colours = M.meta[0];
    stats->n_overflow = M.meta[1];
    stats->colour_rounds = M.meta[4];
}

__global__ void k_colour_start_init(ManifoldArrays M) {
    if (threadIdx.x < OVERFLOW_COLOUR + 2) M.colour_start[threadIdx.x] = -1;
}

// ------------------------------------------------------------------ row build

// all rows of one manifold -> solver slot s (getInfo2 + QuickStep rhs/Ad per row)
__device__ __forceinline__ void build_manifold_rows(int s, const int4 rec, const BodyArrays &B, const ContactSource &src,
                                                    const Surface &usurf, const SolverArrays &S, const StepConfig &cfg,
                                                    int &rows1, int &rows2, int &ncont) {
    const int nc = rec.w & 0xff;
    const bool rev = rec.w & REC_REV, two = rec.y >= 0;
    S.mrec[s] = make_int4(rec.x, rec.y, nc, rec.z);
    const BodyKin k1 = load_kin(B, rec.x);
    BodyKin k2;
    if (two) k2 = load_kin(B, rec.y);
    else {
        k2.x = v3(0.f, 0.f, 0.f); k2.lv = k2.av = k2.tv = k2.tw = k2.x;
        k2.iI = M3{k2.x, k2.x, k2.x}; k2.invM = 0.f;
    }
    for (int k = 0; k < nc; k++) {
        const size_t ci = (size_t)rec.z + (size_t)k * src.kstride;
        const float4 pd = src.pd[ci], ns = src.ns[ci];
        const Surface sf = src.surf ? src.surf[ci] : usurf;
        const int the_m = surface_rows(sf);
        V3 normal = v3(ns);
        if (rev) normal = -normal;
        const V3 pos = v3(pd);
        const V3 c1 = pos - k1.x;
        const V3 c2 = two ? (pos - k2.x) : v3(0.f, 0.f, 0.f);
        // normal row: pushout, max_vel cap, bounce
        float erp = cfg.erp;
        if (sf.mode & MODE_SOFT_ERP) erp = sf.soft_erp;
        const float kk = (1.0f / cfg.h) * erp;
        float depth = pd.w - cfg.min_depth;
        if (depth < 0) depth = 0;
        float cfmN = cfg.cfm;
        if (sf.mode & MODE_SOFT_CFM) cfmN = sf.soft_cfm;
        float motionN = 0.f;
        if (sf.mode & MODE_MOTIONN) motionN = sf.motionN;
        float cN = kk * depth + motionN;
        if (cN > cfg.max_vel) cN = cfg.max_vel;
        if (sf.mode & MODE_BOUNCE) {
            const V3 J1a = cross(c1, normal);
            float outgoing = dot(normal, k1.lv) + dot(J1a, k1.av);
            if (two) {
                const V3 J2l = -normal;
                const V3 J2a = -cross(c2, normal);
                outgoing += dot(J2l, k2.lv) + dot(J2a, k2.av);
            }
            outgoing -= motionN;
            if (sf.bounce_vel >= 0 && (-outgoing) > sf.bounce_vel) {
                const float newc = -sf.bounce * outgoing + motionN;
                if (newc > cN) cN = newc;
            }
        }
        float rhsN, AdN, AdcfmN;
        build_row(normal, c1, c2, k1, k2, two, cN, cfmN, cfg, rhsN, AdN, AdcfmN);
        float4 q3 = make_float4(0.f, 0.f, 0.f, 0.f), q4 = q3;
        int lflags = the_m;
        if (the_m >= 2) {
            V3 t1, t2;
            const float psk = plane_space_k(normal);
            plane_space_with_k(normal, psk, t1, t2);
            const float mu = sf.mu < 0 ? 0 : sf.mu;
            float c1v = (sf.mode & MODE_MOTION1) ? sf.motion1 : 0.f;
            float cfm1 = (sf.mode & MODE_SLIP1) ? sf.slip1 : cfg.cfm;
            build_row(t1, c1, c2, k1, k2, two, c1v, cfm1, cfg, q3.x, q3.y, q3.z);
            q3.w = psk; // 1/sqrt of dPlaneSpace, reused by every iteration
            q4.w = mu;
            if (sf.mode & MODE_APPROX1_1) lflags |= 0x10;
            if (the_m >= 3) {
                float c2v = (sf.mode & MODE_MOTION2) ? sf.motion2 : 0.f;
                float cfm2 = (sf.mode & MODE_SLIP2) ? sf.slip2 : cfg.cfm;
                build_row(t2, c1, c2, k1, k2, two, c2v, cfm2, cfg, q4.x, q4.y, q4.z);
                if (sf.mode & MODE_MU2) { // second friction limit differs: rare, kept out of the streamed records
                    lflags |= 0x40;
                    S.q5[(size_t)k * S.cap + s] = make_float4(sf.mu2 < 0 ? 0 : sf.mu2, 0.f, 0.f, 0.f);
                }
                if (sf.mode & MODE_APPROX1_2) lflags |= 0x20;
            }
        }
        const size_t si = (size_t)k * S.cap + s;
        S.q0[si] = make_float4(normal.x, normal.y, normal.z, rhsN);
        S.q1[si] = make_float4(c1.x, c1.y, c1.z, AdN);
        S.q2[si] = make_float4(c2.x, c2.y, c2.z, AdcfmN);
        S.q3[si] = q3;
        S.q4[si] = q4;
        S.lam[si] = make_float4(0.f, 0.f, 0.f, __int_as_float(lflags));
        if (two) rows2 += the_m; else rows1 += the_m;
        ncont++;
    }
}

__global__ void __launch_bounds__(128) k_rows(ManifoldArrays M, BodyArrays B, ContactSource src, Surface usurf,
                                               SolverArrays S, StepConfig cfg, StepStats *__restrict__ stats) {
    const int n = *M.count;
    int rows1 = 0, rows2 = 0, ncont = 0;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x)
        build_manifold_rows(s, M.rec[M.sidx[s]], B, src, usurf, S, cfg, rows1, rows2, ncont);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        rows1 += __shfl_xor_sync(0xffffffffu, rows1, o);
        rows2 += __shfl_xor_sync(0xffffffffu, rows2, o);
        ncont += __shfl_xor_sync(0xffffffffu, ncont, o);
    }
    if ((threadIdx.x & 31) == 0 && (rows1 | rows2 | ncont)) {
        atomicAdd(&stats->n_rows1, rows1);
        atomicAdd(&stats->n_rows2, rows2);
        atomicAdd(&stats->n_rows, rows1 + rows2);
        atomicAdd(&stats->n_contacts, ncont);
    }
}

// ------------------------------------------------------------------ SOR/PGS + fused tail

struct FC {
    V3 l, a;
};

// one row of ODE's SOR_LCP inner loop (quickstep.cpp), J rebuilt from (dir, c1, c2)
__device__ __forceinline__ void solve_row(V3 dir, V3 c1, V3 c2, bool two, float invM1, const M3 &iI1, float invM2,
                                          const M3 &iI2, float rhs_s, float Ad, float Adcfm, float lo, float hi,
                                          float &lambda, FC &f1, FC &f2) {
    const V3 J1a = cross(c1, dir);
    const float old_lambda = lambda;
    float delta = rhs_s - old_lambda * Adcfm;
    // J is pre-scaled by Ad in ODE; scale component-wise so the products round identically
    delta -= f1.l.x * (dir.x * Ad) + f1.l.y * (dir.y * Ad) + f1.l.z * (dir.z * Ad) + f1.a.x * (J1a.x * Ad) +
             f1.a.y * (J1a.y * Ad) + f1.a.z * (J1a.z * Ad);
    V3 J2l = v3(0.f, 0.f, 0.f), J2a = J2l;
    if (two) {
        J2l = -dir;
        J2a = -cross(c2, dir);
        delta -= f2.l.x * (J2l.x * Ad) + f2.l.y * (J2l.y * Ad) + f2.l.z * (J2l.z * Ad) + f2.a.x * (J2a.x * Ad) +
                 f2.a.y * (J2a.y * Ad) + f2.a.z * (J2a.z * Ad);
    }
    const float new_lambda = old_lambda + delta;
    if (new_lambda < lo) { delta = lo - old_lambda; lambda = lo; }
    else if (new_lambda > hi) { delta = hi - old_lambda; lambda = hi; }
    else lambda = new_lambda;
    const V3 iM1a = mul(iI1, J1a);
    f1.l.x += delta * (invM1 * dir.x); f1.l.y += delta * (invM1 * dir.y); f1.l.z += delta * (invM1 * dir.z);
    f1.a.x += delta * iM1a.x; f1.a.y += delta * iM1a.y; f1.a.z += delta * iM1a.z;
    if (two) {
        const V3 iM2a = mul(iI2, J2a);
        f2.l.x += delta * (invM2 * J2l.x); f2.l.y += delta * (invM2 * J2l.y); f2.l.z += delta * (invM2 * J2l.z);
        f2.a.x += delta * iM2a.x; f2.a.y += delta * iM2a.y; f2.a.z += delta * iM2a.z;
    }
}

// the six float4 records of one contact of one sorted manifold (DESIGN.md "solver rows")
struct RowRec {
    float4 q0, q1, q2, q3, q4, lam;
};
template <bool NC = true>
__device__ __forceinline__ RowRec load_rows(const SolverArrays &S, size_t si) {
    RowRec r;
    if (!NC) { // the arrays may live in shared memory (small-world solver): generic loads
        r.q0 = S.q0[si]; r.q1 = S.q1[si]; r.q2 = S.q2[si]; r.q3 = S.q3[si]; r.q4 = S.q4[si];
        r.lam = S.lam[si];
        return r;
    }
    // rows are written by k_rows before this kernel starts: read-only path; lambda is thread-private
    r.q0 = __ldg(&S.q0[si]); r.q1 = __ldg(&S.q1[si]); r.q2 = __ldg(&S.q2[si]);
    r.q3 = __ldg(&S.q3[si]); r.q4 = __ldg(&S.q4[si]);
    r.lam = S.lam[si];
    return r;
}

// All rows of one manifold, in order (normal, tangent 1, tangent 2) per contact.  The loads of the
// manifold's first contact are issued together with the body gathers, and contact k+1 is fetched
// while contact k is being solved, so the dependent chain per manifold is: record -> {bodies, rows}
// -> arithmetic, not one DRAM round trip per contact.  The body accumulators fc are exchanged
// between SMs from one colour to the next, so they bypass L1 (ld.cg / st.cg).
template <bool L2ONLY>
__device__ __forceinline__ float4 ld_fc(const float4 *p) { return L2ONLY ? __ldcg(p) : *p; }
template <bool L2ONLY>
__device__ __forceinline__ void st_fc(float4 *p, float4 v) {
    if (L2ONLY) __stcg(p, v); else *p = v;
}

// fcp / invp: where the accumulators and world inverse inertias of the unit's bodies live -- the global
// arrays (indexed by body) or, on the island path, the env's copy in shared memory (indexed by local body)
template <bool L2ONLY, bool SINGLE = false, bool NC = true>
__device__ __forceinline__ void solve_manifold_core(int s, const int4 rec, RowRec cur, const SolverArrays &S, float4 *fcp,
                                                    const float4 *invp, int fs = 2, int fo = 1, float *maxd = nullptr) {
    const int b1 = rec.x, b2 = rec.y, nc = SINGLE ? 1 : rec.z; // SINGLE: per-contact units, no contact loop
    const bool two = b2 >= 0;
    FC f1, f2;
    {
        const float4 a = ld_fc<L2ONLY>(&fcp[fs * b1]), b = ld_fc<L2ONLY>(&fcp[fs * b1 + fo]);
        f1.l = v3(a); f1.a = v3(b);
    }
    const float4 i10 = invp[3 * b1], i11 = invp[3 * b1 + 1], i12 = invp[3 * b1 + 2];
    const M3 iI1 = M3{v3(i10), v3(i11), v3(i12)};
    const float invM1 = i10.w;
    M3 iI2 = M3{v3(0.f, 0.f, 0.f), v3(0.f, 0.f, 0.f), v3(0.f, 0.f, 0.f)};
    float invM2 = 0.f;
    f2.l = v3(0.f, 0.f, 0.f); f2.a = f2.l;
    if (two) {
        const float4 a = ld_fc<L2ONLY>(&fcp[fs * b2]), b = ld_fc<L2ONLY>(&fcp[fs * b2 + fo]);
        f2.l = v3(a); f2.a = v3(b);
        const float4 i20 = invp[3 * b2], i21 = invp[3 * b2 + 1], i22 = invp[3 * b2 + 2];
        iI2 = M3{v3(i20), v3(i21), v3(i22)};
        invM2 = i20.w;
    }
    for (int k = 0; k < nc; k++) {
        const size_t si = (size_t)k * S.cap + s;
        RowRec nxt = cur;
        if (!SINGLE && k + 1 < nc) nxt = load_rows<NC>(S, si + S.cap);
        float4 lam = cur.lam;
        const int lflags = __float_as_int(lam.w);
        const int the_m = lflags & 0xf;
        const V3 n = v3(cur.q0), c1 = v3(cur.q1), c2 = v3(cur.q2);
        solve_row(n, c1, c2, two, invM1, iI1, invM2, iI2, cur.q0.w, cur.q1.w, cur.q2.w, 0.f, INFINITY, lam.x, f1, f2);
        if (the_m >= 2) {
            const float4 q3 = cur.q3;
            const float mu = cur.q4.w;
            V3 t1, t2;
            plane_space_with_k(n, q3.w, t1, t2);
            float hi = mu, lo = -mu;
            if (lflags & 0x10) { hi = fabsf(mu * lam.x); lo = -hi; }
            solve_row(t1, c1, c2, two, invM1, iI1, invM2, iI2, q3.x, q3.y, q3.z, lo, hi, lam.y, f1, f2);
            if (the_m >= 3) {
                const float4 q4 = cur.q4;
                const float mu2 = (lflags & 0x40) ? S.q5[si].x : mu;
                hi = mu2; lo = -mu2;
                if (lflags & 0x20) { hi = fabsf(mu2 * lam.x); lo = -hi; }
                solve_row(t2, c1, c2, two, invM1, iI1, invM2, iI2, q4.x, q4.y, q4.z, lo, hi, lam.z, f1, f2);
            }
        }
        S.lam[si] = lam;
        if (maxd) // residual-terminated mode: largest |delta lambda| this thread produced in the sweep
            *maxd = fmaxf(*maxd, fmaxf(fabsf(lam.x - cur.lam.x), fmaxf(fabsf(lam.y - cur.lam.y), fabsf(lam.z - cur.lam.z))));
        if (!SINGLE) cur = nxt;
    }
    st_fc<L2ONLY>(&fcp[fs * b1], make_float4(f1.l.x, f1.l.y, f1.l.z, 0.f));
    st_fc<L2ONLY>(&fcp[fs * b1 + fo], make_float4(f1.a.x, f1.a.y, f1.a.z, 0.f));
    if (two) {
        st_fc<L2ONLY>(&fcp[fs * b2], make_float4(f2.l.x, f2.l.y, f2.l.z, 0.f));
        st_fc<L2ONLY>(&fcp[fs * b2 + fo], make_float4(f2.a.x, f2.a.y, f2.a.z, 0.f));
    }
}

template <bool L2ONLY, bool SINGLE = false, bool NC = true>
__device__ __forceinline__ void solve_manifold(int s, const SolverArrays &S, float4 *fcp, const float4 *invp, int fs = 2,
                                               int fo = 1, float *maxd = nullptr) {
    const int4 rec = NC ? __ldg(&S.mrec[s]) : S.mrec[s];
    solve_manifold_core<L2ONLY, SINGLE, NC>(s, rec, load_rows<NC>(S, (size_t)s), S, fcp, invp, fs, fo, maxd);
}

#ifdef OB_ENV_PROFILE
__device__ unsigned long long g_phase_t[256];
__device__ int g_phase_n[256];
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#endif

// Persistent solver of one big world: colour phases separated by grid barriers (1.35 us each at 296 CTAs,
// measured).  Per-phase timing on the 1 M-body pile (profiles/README.md): the large early colours run at
// ~4.9 TB/s of rows + body data, the small late colours sit on a ~6 us latency floor.  (Prefetching the
// next phase's rows with cp.async before the barrier, L2 evict-first hints on the rows and a persisting-L2
// window on the body data, and fetching each thread's first record + rows of the next phase into registers
// before the barrier, were all measured and did not help; they are not in the code.)
// TOL: residual-terminated sweeps (the dWorldStep parity mode): every thread tracks the
// largest |delta lambda| of its sweep, the grid max goes through one of three rotating slots (meta[8..10]:
// a slot is re-zeroed only after every CTA has passed the barrier that follows its last read), and all CTAs
// take the same exit decision after the sweep's last barrier.
#ifndef OB_SOLVE_CTAS
#define OB_SOLVE_CTAS 2 // CTAs of 256 threads per SM the register budget is sized for
#endif
template <bool TOL>
__global__ void __launch_bounds__(256, OB_SOLVE_CTAS) k_solve(ManifoldArrays M, SolverArrays S, BodyArrays B, StepConfig cfg,
                                                  StepStats *__restrict__ stats, const int *__restrict__ done_flag) {
    if (done_flag && *done_flag) return; // the small-world solver already did the whole solve + tail
    const int n = *M.count;
    const int gt = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;
    unsigned *bar = reinterpret_cast<unsigned *>(&M.meta[6]);
    unsigned target = 0;
    __shared__ int cs[OVERFLOW_COLOUR + 2]; // colour bucket starts: read once, not once per phase
    if (threadIdx.x < OVERFLOW_COLOUR + 2) cs[threadIdx.x] = M.colour_start[threadIdx.x];
    __syncthreads();
    if (n > 0) {
        const int ncol = M.meta[0];
        const int ovf0 = cs[OVERFLOW_COLOUR], ovf1 = cs[OVERFLOW_COLOUR + 1];
#ifdef OB_ENV_PROFILE
        if (gt == 0) g_phase_t[0] = gtimer();
#endif
        unsigned *resid = reinterpret_cast<unsigned *>(&M.meta[8]);
        int it = 0;
        for (; it < cfg.iters; it++) {
            float maxd = 0.f;
            for (int c = 0; c < ncol; c++) {
                const int s0 = cs[c], s1 = cs[c + 1];
                for (int s = s0 + gt; s < s1; s += gs) solve_manifold<true>(s, S, B.fc, B.inv, 2, 1, TOL ? &maxd : nullptr);
                if (TOL && c == ncol - 1 && ovf1 == ovf0) {
                    maxd = fmaxf(maxd, __shfl_xor_sync(0xffffffffu, maxd, 16)); maxd = fmaxf(maxd, __shfl_xor_sync(0xffffffffu, maxd, 8));
                    maxd = fmaxf(maxd, __shfl_xor_sync(0xffffffffu, maxd, 4)); maxd = fmaxf(maxd, __shfl_xor_sync(0xffffffffu, maxd, 2));
                    maxd = fmaxf(maxd, __shfl_xor_sync(0xffffffffu, maxd, 1));
                    if ((threadIdx.x & 31) == 0 && maxd > 0.f) atomicMax(&resid[it % 3], __float_as_uint(maxd));
                }
                grid_barrier(bar, target);
#ifdef OB_ENV_PROFILE
                if (gt == 0 && it * ncol + c < 255) { g_phase_t[it * ncol + c + 1] = gtimer(); g_phase_n[it * ncol + c + 1] = s1 - s0; }
#endif
            }
            if (ovf1 > ovf0) {
                // manifolds that found no free colour (> 64 neighbours): one thread, in order
                if (gt == 0)
                    for (int s = ovf0; s < ovf1; s++) solve_manifold<true>(s, S, B.fc, B.inv, 2, 1, TOL ? &maxd : nullptr);
                if (TOL) {
                    maxd = fmaxf(maxd, __shfl_xor_sync(0xffffffffu, maxd, 16)); maxd = fmaxf(maxd, __shfl_xor_sync(0xffffffffu, maxd, 8));
                    maxd = fmaxf(maxd, __shfl_xor_sync(0xffffffffu, maxd, 4)); maxd = fmaxf(maxd, __shfl_xor_sync(0xffffffffu, maxd, 2));
                    maxd = fmaxf(maxd, __shfl_xor_sync(0xffffffffu, maxd, 1));
                    if ((threadIdx.x & 31) == 0 && maxd > 0.f) atomicMax(&resid[it % 3], __float_as_uint(maxd));
                }
                grid_barrier(bar, target);
            }
            if (TOL) {
                unsigned r;
                asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(r) : "l"(&resid[it % 3]) : "memory");
                if (gt == 0) resid[(it + 2) % 3] = 0u; // last read before this sweep's barrier, next written two sweeps on
                if (__uint_as_float(r) < cfg.tol) { it++; break; }
            }
        }
        if (gt == 0) stats->solver_iters = it;
    }
    for (int i = gt; i < B.n; i += gs) integrate_body(i, B, cfg.h, __ldcg(&B.fc[2 * i]), __ldcg(&B.fc[2 * i + 1]));
}

// Small single worlds (the reference's own scene: 68 bodies, ~80 manifolds): the colour phases of k_solve are
// ~2 us each -- a grid barrier plus a chain of L2 round trips -- for a handful of manifolds.  One CTA with the
// rows, the manifold records, the accumulators and the world inverse inertias in shared memory runs the same
// phases (same colours, same order inside a manifold: bit-identical results) separated by __syncthreads only.
// Falls through (flag stays 0) when the world does not fit; k_solve then does the work.
constexpr int TINY_MANIFOLDS = 160, TINY_BODIES = 256, TINY_THREADS = 256;
__global__ void __launch_bounds__(TINY_THREADS) k_tiny_solve(ManifoldArrays M, SolverArrays S, BodyArrays B, StepConfig cfg,
                                                              StepStats *__restrict__ stats, int *__restrict__ done_flag) {
    extern __shared__ __align__(16) unsigned char tiny_smem[];
    const int n = *M.count, nb = B.n, tid = threadIdx.x;
    if (n > TINY_MANIFOLDS || nb > TINY_BODIES) {
        if (tid == 0) *done_flag = 0;
        return;
    }
    __shared__ int cs[OVERFLOW_COLOUR + 2];
    SolverArrays T;
    T.cap = n;
    float4 *p = reinterpret_cast<float4 *>(tiny_smem);
    T.q0 = p; p += 8 * n; T.q1 = p; p += 8 * n; T.q2 = p; p += 8 * n; T.q3 = p; p += 8 * n;
    T.q4 = p; p += 8 * n; T.q5 = p; p += 8 * n; T.lam = p; p += 8 * n;
    T.mrec = reinterpret_cast<int4 *>(p); p += n;
    float4 *fc = p; p += 2 * nb;
    float4 *inv = p;
    if (tid < OVERFLOW_COLOUR + 2) cs[tid] = M.colour_start[tid];
    for (int s = tid; s < n; s += TINY_THREADS) {
        const int4 rec = S.mrec[s];
        T.mrec[s] = rec;
        for (int k = 0; k < rec.z; k++) {
            const size_t g = (size_t)k * S.cap + s;
            const int t = k * n + s;
            T.q0[t] = S.q0[g]; T.q1[t] = S.q1[g]; T.q2[t] = S.q2[g]; T.q3[t] = S.q3[g]; T.q4[t] = S.q4[g];
            const float4 lam = S.lam[g];
            T.lam[t] = lam;
            if (__float_as_int(lam.w) & 0x40) T.q5[t] = S.q5[g];
        }
    }
    for (int i = tid; i < 2 * nb; i += TINY_THREADS) fc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < 3 * nb; i += TINY_THREADS) inv[i] = B.inv[i];
    __syncthreads();
    if (n > 0) {
        const int ncol = M.meta[0];
        const int ovf0 = cs[OVERFLOW_COLOUR], ovf1 = cs[OVERFLOW_COLOUR + 1];
        for (int it = 0; it < cfg.iters; it++) {
            for (int c = 0; c < ncol; c++) {
                for (int s = cs[c] + tid; s < cs[c + 1]; s += TINY_THREADS) solve_manifold<false, false, false>(s, T, fc, inv);
                __syncthreads();
            }
            if (ovf1 > ovf0) {
                if (tid == 0)
                    for (int s = ovf0; s < ovf1; s++) solve_manifold<false, false, false>(s, T, fc, inv);
                __syncthreads();
            }
        }
    }
    for (int i = tid; i < nb; i += TINY_THREADS) {
        B.fc[2 * i] = fc[2 * i]; B.fc[2 * i + 1] = fc[2 * i + 1]; // dWorldPackImpulsesDeviceB200 reads them
        integrate_body(i, B, cfg.h, fc[2 * i], fc[2 * i + 1]);
    }
    if (tid == 0) { *done_flag = 1; stats->solver_iters = cfg.iters; }
}

// micro-benchmark hook: cost of one grid barrier at the solver's launch shape
__global__ void __launch_bounds__(256, 2) k_barrier_bench(unsigned *bar, int iters) {
    unsigned target = 0;
    for (int i = 0; i < iters; i++) grid_barrier(bar, target);
}

float solver_barrier_bench(Engine *e, int iters) {
    unsigned *bar = nullptr;
    OB_CUDA(ob_malloc(&bar, sizeof(unsigned)));
    OB_CUDA(cudaMemset(bar, 0, sizeof(unsigned)));
    int per_sm = 0;
    OB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)k_barrier_bench, 256, 0));
    if (per_sm > 2) per_sm = 2;
    int grid = per_sm * e->num_sms;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    void *args[] = {(void *)&bar, (void *)&iters};
    OB_CUDA(cudaLaunchCooperativeKernel((const void *)k_barrier_bench, dim3((unsigned)grid), dim3(256), args, 0, e->st));
    OB_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned), e->st));
    OB_CUDA(cudaEventRecord(a, e->st));
    OB_CUDA(cudaLaunchCooperativeKernel((const void *)k_barrier_bench, dim3((unsigned)grid), dim3(256), args, 0, e->st));
    OB_CUDA(cudaEventRecord(b, e->st));
    OB_CUDA(cudaEventSynchronize(b));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    ob_free(bar);
    return ms * 1000.f / iters; // microseconds per barrier
}

// ------------------------------------------------------------------ island solver for batched worlds
//
// Scenes made of many small independent worlds ("envs", BASELINE config 4) do not need grid-wide
// barriers: an env is an island.  Manifolds are bucketed per env, and one group of G lanes of a warp
// owns one env for the whole solve: it colours the env's manifolds (same rule and priorities as
// k_colour, so the colours -- and therefore the Gauss-Seidel order and every result bit -- are the
// same as on the global path), orders them by colour, builds their rows and runs all iterations
// with __syncwarp() between colours.  Rows and accumulators of an env stay in L1/L2 for its 20
// iterations instead of being streamed from HBM 20 times.

__global__ void __launch_bounds__(256) k_env_count(const BroadCounters *__restrict__ bc, const int2 *__restrict__ pairs,
                                                    const int *__restrict__ g_body, const int *__restrict__ nc,
                                                    const int *__restrict__ b_env, int *__restrict__ env_cnt,
                                                    int per_contact) {
    const int n = bc->n_pairs;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const int2 pr = pairs[p];
        const int b1 = g_body[pr.x], b2 = g_body[pr.y];
        const int c = nc[p];
        if (c > 0 && (b1 >= 0 || b2 >= 0)) atomicAdd(&env_cnt[b_env[b1 >= 0 ? b1 : b2]], per_contact ? c : 1);
    }
}

__global__ void __launch_bounds__(256) k_env_bucket(const BroadCounters *__restrict__ bc, const int2 *__restrict__ pairs,
                                                     const int *__restrict__ g_body, const int *__restrict__ nc,
                                                     const float4 *__restrict__ b_pos, const int *__restrict__ b_env,
                                                     EnvArrays E, StepStats *__restrict__ stats, int per_contact, int kstride,
                                                     int *__restrict__ m_count) {
    const int n = bc->n_pairs;
    if (E.start[E.n_envs] > E.cap) { // more solver units than the arrays hold: flagged, never silent; the solve is skipped
        if (blockIdx.x == 0 && threadIdx.x == 0) { atomicOr(&stats->flags, SF_MANIFOLD_OVERFLOW); stats->n_manifolds = 0; *m_count = 0; }
        return;
    }
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const int2 pr = pairs[p];
        int b1 = g_body[pr.x], b2 = g_body[pr.y];
        const int c = nc[p];
        if (!(c > 0 && (b1 >= 0 || b2 >= 0))) continue;
        const int units = per_contact ? c : 1;
        int w = per_contact ? 1 : c;
        if (b1 < 0) { b1 = b2; b2 = -1; w |= REC_REV; }
        if (b_pos[b1].w > 0.f) w |= REC_DYN1;
        if (b2 >= 0 && b_pos[b2].w > 0.f) w |= REC_DYN2;
        const int e = b_env[b1];
        const int slot = E.start[e] + atomicAdd(&E.fill[e], units);
        for (int u = 0; u < units; u++) E.rec[slot + u] = make_int4(b1, b2, p + u * kstride, w);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { stats->n_manifolds = E.start[E.n_envs]; *m_count = E.start[E.n_envs]; }
}

// Work queue order of the island solver: envs by decreasing unit count (64 size classes), so the warps' last
// envs are the small ones and the kernel's tail is short.  The order inside a class is arbitrary; it does not
// affect any result (envs are independent).
__global__ void __launch_bounds__(1024) k_env_order(const int *__restrict__ cnt, int n_envs, int *__restrict__ order) {
    __shared__ int hist[64], start[64];
    if (threadIdx.x < 64) hist[threadIdx.x] = 0;
    __syncthreads();
    for (int e = threadIdx.x; e < n_envs; e += blockDim.x) atomicAdd(&hist[63 - min(cnt[e] >> 3, 63)], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int b = 0; b < 64; b++) { start[b] = acc; acc += hist[b]; }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < n_envs; e += blockDim.x) order[atomicAdd(&start[63 - min(cnt[e] >> 3, 63)], 1)] = e;
}

#ifdef OB_ENV_PROFILE
__device__ unsigned long long g_env_prof[8];
#define PROF_T(var) const long long var = clock64()
#define PROF_ADD(i, a, b) if (g == 0) atomicAdd(&g_env_prof[i], (unsigned long long)((b) - (a)))
#else
#define PROF_T(var)
#define PROF_ADD(i, a, b)
#endif

#ifndef OB_ENV_CTAS
#define OB_ENV_CTAS 5 // CTAs per SM the register budget is sized for (measured on C4: 5 >= 4 > 6, 8)
#endif

// The W = 32/G envs that share a warp run in LOCKSTEP: every loop bound is made warp-uniform (the
// maximum over the warp's groups) and lanes without work are predicated off, so the groups never
// diverge into serialised code paths.  The solver is issue-bound (ncu: 64 % issue-active, DRAM 10 %),
// so what matters is how many lanes of each issued instruction do useful work.
template <int G, bool SINGLE>
__global__ void __launch_bounds__(128, OB_ENV_CTAS) k_env_solve(EnvArrays E, BodyArrays B, ContactSource src, Surface usurf,
                                                    SolverArrays S, StepConfig cfg, int spread, int stage, int fused,
                                                    StepStats *__restrict__ stats) {
    extern __shared__ __align__(16) unsigned char env_smem[];
    constexpr int GROUPS = 128 / G;   // envs per CTA
    constexpr unsigned FULL = 0xffffffffu;
    const int grp = threadIdx.x / G, g = threadIdx.x % G;
    const int lane = threadIdx.x & 31;
    const int mb = (E.max_bodies + 31) & ~31;
    // per group: a body region -- colour masks + priorities while colouring (16 B per body), then, with
    // `stage`, the env's accumulators fc and world inverse inertias (80 B per body) for the iterations --
    // followed by the colour bucket starts and cursors
    const size_t region = (size_t)mb * (stage ? 80 : 16);
    unsigned long long *masks = reinterpret_cast<unsigned long long *>(env_smem + (size_t)grp * region);
    unsigned long long *prio = masks + mb;
    float4 *sm_fc = reinterpret_cast<float4 *>(env_smem + (size_t)grp * region);
    float4 *sm_inv = sm_fc + 2 * (size_t)mb;
    int *cstart = reinterpret_cast<int *>(env_smem + (size_t)GROUPS * region) + grp * 136;
    int *cursor = cstart + 68;
    int rows1 = 0, rows2 = 0, ncont = 0, max_col = 0, max_rounds = 0;
    // Persistent warps: a warp takes the next block of 32/G envs from a counter when it is done, instead of a
    // CTA of four warps waiting for its slowest env before the next CTA can start (envs differ by +-15 % in
    // contacts and by a few colours).
    constexpr int W = 32 / G; // envs per warp
    int *next_item = &E.fill[E.n_envs];
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(next_item, 1);
        item = __shfl_sync(FULL, item, 0);
        if (item * W >= E.n_envs) break;
        const int qi = item * W + lane / G;
        const bool have = qi < E.n_envs;
        const int env = have ? E.order[qi] : E.n_envs;
        const bool fits = E.start[E.n_envs] <= E.cap; // unit overflow (flagged by k_env_bucket): free flight
        const int ms = (have && fits) ? E.start[env] : 0, me = (have && fits) ? E.start[env + 1] : 0;
        const int trips = (__reduce_max_sync(FULL, me - ms) + G - 1) / G; // warp-uniform
        const int fb = have ? E.first_body[env] : 0, nbod = have ? E.n_body[env] : 0;
        if (fused) { // per-body step preparation of this env (k_body_prep's work), fused in
            for (int i = g; i < nbod; i += G) body_prep(fb + i, B, cfg);
            __syncwarp();
        }
        if (trips == 0) { // no contacts in any env of this warp: free flight
            if (fused)
                for (int i = g; i < nbod; i += G) {
                    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                    integrate_body(fb + i, B, cfg.h, z, z);
                    if (fused == 2) { B.fc[2 * (size_t)(fb + i)] = z; B.fc[2 * (size_t)(fb + i) + 1] = z; }
                }
            continue;
        }
        PROF_T(t0);
        // ---- colouring (rule of k_colour, at group scope)
        for (int i = g; i < mb; i += G) { masks[i] = 0ull; prio[i] = ~0ull; }
        for (int i = g; i < 68; i += G) { cstart[i] = 0; cursor[i] = 0; }
        // With staging, 64 of the region's 80 bytes per body are idle until the iterations: they cache each unit's
        // priority, env-local body ids, dynamic bits and colour, so the rounds below touch shared memory only
        // (measured: the rounds were 20 % of the kernel, one L2 round trip after another).
        const int nun = me - ms;
        const bool fastc = stage && __all_sync(FULL, nun <= 4 * mb);
        uint4 *uc = reinterpret_cast<uint4 *>(env_smem + (size_t)grp * region + 16 * (size_t)mb);
        constexpr unsigned UC_D1 = 1u << 24, UC_D2 = 1u << 25, UC_REV = 1u << 26; // bits 27..30: contacts of the unit
        if (fastc) {
            // all loads of a stage are issued before the first use (two dependent L2 round trips in total, not per trip)
            constexpr int TB = 6;
            for (int j0 = 0; j0 < trips; j0 += TB) {
                int4 r[TB];
                int l1[TB], l2[TB];
#pragma unroll
                for (int k = 0; k < TB; k++) {
                    const int m = ms + g + (j0 + k) * G;
                    r[k] = (j0 + k < trips && m < me) ? E.rec[m] : make_int4(-1, -1, 0, 0);
                }
#pragma unroll
                for (int k = 0; k < TB; k++) {
                    l1[k] = r[k].x >= 0 ? B.local[r[k].x] : -1;
                    l2[k] = r[k].y >= 0 ? B.local[r[k].y] : -1;
                }
#pragma unroll
                for (int k = 0; k < TB; k++) {
                    if (r[k].x < 0) continue;
                    const unsigned long long pr = manifold_prio(r[k].z, l1[k], l2[k]);
                    uc[g + (j0 + k) * G] = make_uint4((unsigned)pr, (unsigned)(pr >> 32),
                                                       ((unsigned)l1[k] & 0xfffu) | (((unsigned)l2[k] & 0xfffu) << 12) |
                                                           ((r[k].w & REC_DYN1) ? UC_D1 : 0u) | ((r[k].w & REC_DYN2) ? UC_D2 : 0u) |
                                                           ((r[k].w & REC_REV) ? UC_REV : 0u) | (((unsigned)r[k].w & 0xfu) << 27),
                                                       255u);
                }
            }
        } else {
            for (int j = 0; j < trips; j++) {
                const int m = ms + g + j * G;
                if (m < me) E.col[m] = 255;
            }
        }
        __syncwarp();
        int rounds = 0;
        if (fastc) {
            for (;; rounds++) {
                // later rounds stamp their priorities with a smaller top byte (as k_colour does): atomicMin then
                // prefers them over what earlier rounds left behind and the losers need no reset pass
                const bool stamped = rounds < 254;
                const unsigned long long stamp = stamped ? ((unsigned long long)(254 - rounds) << 56) : 0ull;
                for (int j = 0; j < trips; j++) {
                    const int u = g + j * G;
                    if (u < nun) {
                        const uint4 q = uc[u];
                        if (q.w == 255u) {
                            const unsigned long long pr = stamp | (unsigned long long)q.x | ((unsigned long long)q.y << 32);
                            if (q.z & UC_D1) atomicMin(&prio[q.z & 0xfffu], pr);
                            if (q.z & UC_D2) atomicMin(&prio[(q.z >> 12) & 0xfffu], pr);
                        }
                    }
                }
                __syncwarp();
                bool left = false;
                for (int j = 0; j < trips; j++) {
                    const int u = g + j * G;
                    if (u < nun) {
                        const uint4 q = uc[u];
                        if (q.w == 255u) {
                            const unsigned long long base = (unsigned long long)q.x | ((unsigned long long)q.y << 32);
                            const unsigned long long pr = stamp | base;
                            const int l1 = (int)(q.z & 0xfffu), l2 = (int)((q.z >> 12) & 0xfffu);
                            const bool d1 = q.z & UC_D1, d2 = q.z & UC_D2;
                            if ((!d1 || prio[l1] == pr) && (!d2 || prio[l2] == pr)) {
                                unsigned long long mask = 0ull;
                                if (d1) mask |= masks[l1];
                                if (d2) mask |= masks[l2];
                                const int c = pick_colour(mask, base, spread);
                                if (c != OVERFLOW_COLOUR) {
                                    const unsigned long long bit = 1ull << c;
                                    if (d1) masks[l1] |= bit;
                                    if (d2) masks[l2] |= bit;
                                }
                                uc[u].w = (unsigned)c;
                                atomicAdd(&cstart[c + 1], 1);
                            } else left = true;
                        }
                    }
                }
                __syncwarp();
                if (!__any_sync(FULL, left)) break;
                if (!stamped) { // out of stamps: reset the losers' bodies
                    for (int j = 0; j < trips; j++) {
                        const int u = g + j * G;
                        if (u < nun) {
                            const uint4 q = uc[u];
                            if (q.w == 255u) {
                                if (q.z & UC_D1) prio[q.z & 0xfffu] = ~0ull;
                                if (q.z & UC_D2) prio[(q.z >> 12) & 0xfffu] = ~0ull;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
        } else
        for (;; rounds++) {
            for (int j = 0; j < trips; j++) {
                const int m = ms + g + j * G;
                if (m < me && E.col[m] == 255) {
                    const int4 r = E.rec[m];
                    const int l1 = B.local[r.x], l2 = r.y >= 0 ? B.local[r.y] : -1;
                    const unsigned long long pr = manifold_prio(r.z, l1, l2);
                    if (r.w & REC_DYN1) atomicMin(&prio[l1], pr);
                    if (r.w & REC_DYN2) atomicMin(&prio[l2], pr);
                }
            }
            __syncwarp();
            bool left = false;
            for (int j = 0; j < trips; j++) {
                const int m = ms + g + j * G;
                if (m < me && E.col[m] == 255) {
                    const int4 r = E.rec[m];
                    const int l1 = B.local[r.x], l2 = r.y >= 0 ? B.local[r.y] : -1;
                    const unsigned long long pr = manifold_prio(r.z, l1, l2);
                    const bool d1 = r.w & REC_DYN1, d2 = r.w & REC_DYN2;
                    if ((!d1 || prio[l1] == pr) && (!d2 || prio[l2] == pr)) {
                        unsigned long long mask = 0ull;
                        if (d1) mask |= masks[l1];
                        if (d2) mask |= masks[l2];
                        const int c = pick_colour(mask, pr, spread);
                        if (c != OVERFLOW_COLOUR) {
                            const unsigned long long bit = 1ull << c;
                            if (d1) masks[l1] |= bit;
                            if (d2) masks[l2] |= bit;
                        }
                        E.col[m] = (unsigned char)c;
                        atomicAdd(&cstart[c + 1], 1);
                    } else left = true;
                }
            }
            __syncwarp();
            if (!__any_sync(FULL, left)) break;
            for (int j = 0; j < trips; j++) {
                const int m = ms + g + j * G;
                if (m < me && E.col[m] == 255) {
                    const int4 r = E.rec[m];
                    if (r.w & REC_DYN1) prio[B.local[r.x]] = ~0ull;
                    if (r.w & REC_DYN2) prio[B.local[r.y]] = ~0ull;
                }
            }
            __syncwarp();
        }
        PROF_T(t1);
        // ---- order the env's manifolds by colour (counting sort at group scope)
        if (g == 0) {
            int acc = 0;
            for (int c = 0; c <= OVERFLOW_COLOUR + 1; c++) { acc += cstart[c]; cstart[c] = acc; }
        }
        __syncwarp();
        int ncol = 0;
        for (int j = 0; j < trips; j++) {
            const int m = ms + g + j * G;
            if (m < me) {
                const int c = fastc ? (int)uc[m - ms].w : (int)E.col[m];
                const int r = atomicAdd(&cursor[c], 1);
                // sorted position -> unit: in the cache's (now dead) high priority word, else in global memory
                if (fastc) uc[cstart[c] + r].y = (unsigned)(m - ms);
                else E.perm[ms + cstart[c] + r] = m;
                if (c < OVERFLOW_COLOUR && c + 1 > ncol) ncol = c + 1;
            }
        }
        __syncwarp();
        ncol = __reduce_max_sync(FULL, ncol); // colours of the busiest env of the warp
        PROF_T(t2);
        // ---- rows
        for (int j = 0; j < trips; j++) {
            const int s = ms + g + j * G;
            if (s < me) {
                int4 r;
                int l1 = 0, l2 = -1;
                if (fastc) {
                    // the whole record comes out of the cache: unit id, contact slot (= the priority's tie word),
                    // local body ids (bodies of an env are one index range), flags -- no global reads before the rows
                    const uint4 q = uc[uc[s - ms].y];
                    l1 = (int)(q.z & 0xfffu);
                    l2 = (int)((q.z >> 12) & 0xfffu);
                    if (l2 == 0xfff) l2 = -1;
                    const int w = (int)((q.z >> 27) & 0xfu) | ((q.z & UC_REV) ? REC_REV : 0) | ((q.z & UC_D1) ? REC_DYN1 : 0) |
                                  ((q.z & UC_D2) ? REC_DYN2 : 0);
                    r = make_int4(fb + l1, l2 >= 0 ? fb + l2 : -1, (int)q.x, w);
                } else {
                    r = E.rec[E.perm[s]];
                    if (stage) { l1 = B.local[r.x]; l2 = r.y >= 0 ? B.local[r.y] : -1; }
                }
                build_manifold_rows(s, r, B, src, usurf, S, cfg, rows1, rows2, ncont);
                if (stage) S.mrec[s] = make_int4(l1, l2, r.w & 0xff, r.z); // the iterations index the shared-memory copies by local body
            }
        }
        __syncwarp();
        float4 *fcp = B.fc;
        int fs = 2, fo = 1; // fc layout: linear part at fs * b, angular part fo further on
        const float4 *invp = B.inv;
        if (stage) { // colouring scratch is dead: reuse the region for fc (zero) and inv (copied once)
            for (int i = g; i < 2 * mb; i += G) sm_fc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = g; i < 3 * nbod; i += G) sm_inv[i] = B.inv[3 * (size_t)fb + i];
            fcp = sm_fc;
            // split layout in shared memory: with the linear and angular parts interleaved (32 B per body) a
            // quarter-warp's 16-byte accesses can only reach every other bank group -- a built-in 2-way conflict
            fs = 1; fo = mb;
            invp = sm_inv;
            __syncwarp();
        }
        PROF_T(t3);
        // ---- SOR/PGS iterations, colours separated by warp-level barriers only
        const int novf = cstart[OVERFLOW_COLOUR + 1] - cstart[OVERFLOW_COLOUR];
        const bool any_ovf = __any_sync(FULL, novf > 0);
        for (int it = 0; it < cfg.iters; it++) {
            for (int c = 0; c < ncol; c++) {
                const int s0 = ms + cstart[c], s1 = ms + cstart[c + 1];
                const int t = (__reduce_max_sync(FULL, s1 - s0) + G - 1) / G;
                for (int j = 0; j < t; j++) {
                    const int s = s0 + g + j * G;
                    if (s < s1) solve_manifold<false, SINGLE>(s, S, fcp, invp, fs, fo);
                }
                __syncwarp();
            }
            if (any_ovf) {
                // manifolds that found no free colour (> 64 neighbours): one lane per env, in order
                if (g == 0)
                    for (int s = ms + cstart[OVERFLOW_COLOUR]; s < ms + cstart[OVERFLOW_COLOUR + 1]; s++)
                        solve_manifold<false, SINGLE>(s, S, fcp, invp, fs, fo);
                __syncwarp();
            }
        }
        if (fused) { // solver tail: velocity update, dxStepBody, snapshot pack for this env's bodies
            for (int i = g; i < nbod; i += G) {
                const float4 fl = stage ? sm_fc[i] : B.fc[2 * (size_t)(fb + i)];
                const float4 fa = stage ? sm_fc[mb + i] : B.fc[2 * (size_t)(fb + i) + 1];
                integrate_body(fb + i, B, cfg.h, fl, fa);
                if (fused == 2) { // the caller wants the step's impulses (dWorldPackImpulsesDeviceB200)
                    B.fc[2 * (size_t)(fb + i)] = fl;
                    B.fc[2 * (size_t)(fb + i) + 1] = fa;
                }
            }
            __syncwarp();
        } else if (stage) { // hand the accumulators to k_integrate
            for (int i = g; i < nbod; i += G) {
                B.fc[2 * (size_t)(fb + i)] = sm_fc[i];
                B.fc[2 * (size_t)(fb + i) + 1] = sm_fc[mb + i];
            }
            __syncwarp();
        }
        PROF_T(t4);
        PROF_ADD(0, t0, t1); PROF_ADD(1, t1, t2); PROF_ADD(2, t2, t3); PROF_ADD(3, t3, t4); PROF_ADD(4, 0, 1);
        max_col = max(max_col, ncol);
        max_rounds = max(max_rounds, rounds + 1);
        if (g == 0 && novf > 0) atomicAdd(&stats->n_overflow, novf);
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        rows1 += __shfl_xor_sync(FULL, rows1, o);
        rows2 += __shfl_xor_sync(FULL, rows2, o);
        ncont += __shfl_xor_sync(FULL, ncont, o);
        max_col = max(max_col, __shfl_xor_sync(FULL, max_col, o));
        max_rounds = max(max_rounds, __shfl_xor_sync(FULL, max_rounds, o));
    }
    if (lane == 0) {
        if (rows1 | rows2 | ncont) {
            atomicAdd(&stats->n_rows1, rows1);
            atomicAdd(&stats->n_rows2, rows2);
            atomicAdd(&stats->n_rows, rows1 + rows2);
            atomicAdd(&stats->n_contacts, ncont);
        }
        atomicMax(&stats->n_colours, max_col);
        atomicMax(&stats->colour_rounds, max_rounds);
        if (blockIdx.x == 0 && threadIdx.x == 0) stats->solver_iters = cfg.iters;
    }
}

__global__ void __launch_bounds__(256) k_integrate(BodyArrays B, StepConfig cfg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B.n) integrate_body(i, B, cfg.h, __ldcg(&B.fc[2 * i]), __ldcg(&B.fc[2 * i + 1]));
}

void exact_solve_launch(Engine *e, const ManifoldArrays &M, const SolverArrays &S, const BodyArrays &B, const StepConfig &cfg,
                        int *done_flag, cudaStream_t st);
bool exact_solve_fits(int n_bodies);
void env_solve2_launch(Engine *e, const EnvArrays &E, const BodyArrays &B, const ContactSource &src, const Surface &usurf,
                       const SolverArrays &S, const StepConfig &cfg, int fused, int rows_smem, cudaStream_t st);

void solver_profile_dump() {
#ifdef OB_ENV_PROFILE
    unsigned long long h[8];
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(h, g_env_prof, sizeof(h));
    {
        unsigned long long t[256];
        int nn[256];
        cudaMemcpyFromSymbol(t, g_phase_t, sizeof(t));
        cudaMemcpyFromSymbol(nn, g_phase_n, sizeof(nn));
        if (t[1]) {
            fprintf(stderr, "k_solve phases (manifolds: microseconds):");
            for (int i = 1; i < 40 && t[i]; i++) fprintf(stderr, " %d:%.1f", nn[i], (double)(t[i] - t[i - 1]) / 1000.0);
            fprintf(stderr, "\n");
        }
    }
    if (h[4])
        fprintf(stderr, "env-solve profile (avg cycles per env-solve): colour %.0f  sort %.0f  rows %.0f  iterations %.0f  (n=%llu)\n",
                (double)h[0] / h[4], (double)h[1] / h[4], (double)h[2] / h[4], (double)h[3] / h[4], h[4]);
#endif
}

// ------------------------------------------------------------------ host orchestration

static int coop_grid(Engine *e, const void *kernel, int threads, long work_items) {
    int per_sm = 0;
    OB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0));
    if (per_sm < 1) per_sm = 1;
    long max_blocks = (long)per_sm * e->num_sms;
    long want = (work_items + threads - 1) / threads;
    if (want < 1) want = 1;
    return (int)(want < max_blocks ? want : max_blocks);
}

void solver_step(Engine *e, float h, bool host_contacts, const Surface *uniform_surface) {
    cudaStream_t st = e->st;
    BodyArrays B = e->B;
    const int nb = B.n;
    if (nb == 0) return;
    StepConfig cfg;
    cfg.h = h; cfg.erp = e->params.erp; cfg.cfm = e->params.cfm; cfg.sor_w = e->params.sor_w;
    cfg.max_vel = e->params.max_vel; cfg.min_depth = e->params.min_depth;
    cfg.gx = e->params.gravity[0]; cfg.gy = e->params.gravity[1]; cfg.gz = e->params.gravity[2];
    cfg.iters = e->params.iters;
    cfg.tol = e->params.tol;

    // island path with contiguous envs: per-body preparation and the integrate/pack tail run inside
    // k_env_solve, env by env; otherwise they are separate passes over all bodies
    const bool island = !host_contacts && e->have_device_contacts && e->n_envs > 1 && e->E.max_bodies <= 1024 && e->solver_mode != 1 &&
                        !(cfg.tol > 0.f); // residual termination is a grid-wide decision: global solver
    const int fused = (island && e->E.contiguous && e->env_fuse) ? (e->keep_fc ? 2 : 1) : 0;
    {
        const int slots = island ? e->n_envs + 1 : 0;
        k_stats_reset<<<(unsigned)std::max(1, (slots + 255) / 256), 256, 0, st>>>(e->d_stats, e->E.cnt, e->E.fill, slots);
        OB_CHECK_KERNEL("k_stats_reset", st);
    }
    e->last_h = h;
    e->fc_valid = fused != 1; // the fused island path keeps the accumulators in shared memory only
    if (!fused) {
        k_body_prep<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(B, cfg);
        OB_CHECK_KERNEL("k_body_prep", st);
    }

    ManifoldArrays M = e->M;
    const unsigned pgrid = (unsigned)(e->num_sms * 8);
    long max_manifolds;
    ContactSource src;
    Surface usurf{};
    // solver unit: a whole manifold (all contacts of a pair, fewest colours) or a single contact (more
    // colours, but every phase costs one contact instead of the longest manifold of the phase)
    const int per_contact = e->contact_units >= 0 ? e->contact_units : (e->n_envs > 1 ? 1 : 0);
    // batched independent worlds with device-resident contacts take the island path
    const bool env_path = island;
    if (env_path) {
        if (uniform_surface) usurf = *uniform_surface;
        src.pd = e->cs.pd; src.ns = e->cs.ns; src.surf = nullptr; src.kstride = e->cs.stride;
        EnvArrays E = e->E;
        const int ne = E.n_envs;
        k_env_count<<<pgrid, 256, 0, st>>>(e->bp.counters, e->bp.pairs, e->G.body, e->cs.nc, B.env, E.cnt, per_contact);
        OB_CHECK_KERNEL("k_env_count", st);
        scan_exclusive(E.cnt, E.start, (long)ne + 1, nullptr, nullptr, e->scan, st);
        k_env_order<<<1, 1024, 0, st>>>(E.cnt, ne, E.order);
        OB_CHECK_KERNEL("k_env_order", st);
        k_env_bucket<<<pgrid, 256, 0, st>>>(e->bp.counters, e->bp.pairs, e->G.body, e->cs.nc, B.pos, B.env, E, e->d_stats,
                                            per_contact, e->cs.stride, M.count);
        OB_CHECK_KERNEL("k_env_bucket", st);
        if (e->timing) {
            OB_CUDA(cudaEventRecord(e->ev[2], st));
            OB_CUDA(cudaEventRecord(e->ev[3], st));
        }
        // lanes per env (8, 16 or 32): the 32/G envs of a warp run in lockstep
        int G = e->env_group;
        if (G != 8 && G != 16 && G != 32) G = 32; // measured on C4: 32 >= 16 > 8
        const int groups = 128 / G;
        const int mb = (E.max_bodies + 31) & ~31;
        // stage the env's body data in shared memory when its bodies are one index range and it fits
        const int stage = (E.contiguous && mb <= 160 && e->env_stage != 0) ? 1 : 0;
        const size_t smem = (size_t)groups * ((size_t)mb * (stage ? 80 : 16) + 136 * sizeof(int));
        unsigned grid = (unsigned)((ne + groups - 1) / groups);
        SolverArrays S = e->S;
        // per-contact units on contiguous envs: the lane-pair island solver (solver_env.cu), one body's half of a
        // contact per lane; everything else (manifold units, scattered envs, lane groups of 8/16) takes the generic one
        if (per_contact && stage && G == 32 && e->env_pair != 0) {
            env_solve2_launch(e, E, B, src, usurf, S, cfg, fused, e->env_pair_rows, st);
            OB_CHECK_KERNEL("k_env_solve2", st);
            if (!fused) {
                k_integrate<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(B, cfg);
                OB_CHECK_KERNEL("k_integrate", st);
            }
            if (e->timing) OB_CUDA(cudaEventRecord(e->ev[4], st));
            return;
        }
        // one resident wave of CTAs: their warps pull env blocks from the counter E.fill[ne] (zeroed above)
#define OB_LAUNCH_ENV(GG, SS)                                                                                        \
    do {                                                                                                             \
        if (smem > 48 * 1024)                                                                                        \
            OB_CUDA(cudaFuncSetAttribute(k_env_solve<GG, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        int per_sm = 0;                                                                                              \
        OB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_env_solve<GG, SS>, 128, smem));             \
        if (per_sm > 0) grid = std::min(grid, (unsigned)(per_sm * e->num_sms));                                      \
        k_env_solve<GG, SS><<<grid, 128, smem, st>>>(E, B, src, usurf, S, cfg, e->colour_spread, stage, fused, e->d_stats);        \
    } while (0)
        if (per_contact) {
            if (G == 8) OB_LAUNCH_ENV(8, true); else if (G == 16) OB_LAUNCH_ENV(16, true); else OB_LAUNCH_ENV(32, true);
        } else {
            if (G == 8) OB_LAUNCH_ENV(8, false); else if (G == 16) OB_LAUNCH_ENV(16, false); else OB_LAUNCH_ENV(32, false);
        }
#undef OB_LAUNCH_ENV
        OB_CHECK_KERNEL("k_env_solve", st);
        if (!fused) {
            k_integrate<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(B, cfg);
            OB_CHECK_KERNEL("k_integrate", st);
        }
        if (e->timing) OB_CUDA(cudaEventRecord(e->ev[4], st));
        return;
    }
    if (host_contacts) {
        // records + contacts were uploaded by eng_step_host_contacts; *M.count set there
        src.pd = e->hc_pd; src.ns = e->hc_ns; src.surf = e->hc_surf; src.kstride = 1;
        k_manifold_dynbits<<<pgrid, 256, 0, st>>>(M, B.pos);
        OB_CHECK_KERNEL("k_manifold_dynbits", st);
        max_manifolds = (long)e->st_mrec.size();
    } else {
        if (uniform_surface) usurf = *uniform_surface;
        src.pd = e->cs.pd; src.ns = e->cs.ns; src.surf = nullptr; src.kstride = e->cs.stride;
        if (e->have_device_contacts) {
            k_manifold_flags<<<pgrid, 256, 0, st>>>(e->bp.counters, e->bp.pairs, e->G.body, e->cs.nc, M.flag, per_contact);
            OB_CHECK_KERNEL("k_manifold_flags", st);
            scan_exclusive(M.flag, M.flag, e->bp.cap_pairs, &e->bp.counters->n_pairs, &M.meta[5], e->scan, st);
            k_manifold_write<<<pgrid, 256, 0, st>>>(e->bp.counters, e->bp.pairs, e->G.body, e->cs.nc, M.flag, B.pos, M,
                                                    e->d_stats, per_contact, e->cs.stride);
            OB_CHECK_KERNEL("k_manifold_write", st);
            k_manifold_count<<<1, 1, 0, st>>>(&M.meta[5], M, e->d_stats);
            OB_CHECK_KERNEL("k_manifold_count", st);
            max_manifolds = M.cap;
        } else {
            OB_CUDA(cudaMemsetAsync(M.count, 0, sizeof(int), st));
            OB_CUDA(cudaMemsetAsync(M.meta, 0, 8 * sizeof(int), st));
            max_manifolds = 0;
        }
    }
    if (e->timing) OB_CUDA(cudaEventRecord(e->ev[2], st));

    if (max_manifolds > 0) {
        {
            int grid = coop_grid(e, (const void *)k_colour, 256, max_manifolds);
            int spread = e->colour_spread;
            void *args[] = {(void *)&M, (void *)&B, (void *)&spread};
            OB_CUDA(cudaLaunchCooperativeKernel((const void *)k_colour, dim3((unsigned)grid), dim3(256), args, 0, st));
            OB_CHECK_KERNEL("k_colour", st);
        }
        k_colour_keys<<<pgrid, 256, 0, st>>>(M);
        OB_CHECK_KERNEL("k_colour_keys", st);
        sort_pairs(M.skey, M.sidx, max_manifolds, M.count, 10, e->sort, st);
        k_colour_start_init<<<1, 128, 0, st>>>(M);
        OB_CHECK_KERNEL("k_colour_start_init", st);
        k_colour_bounds<<<pgrid, 256, 0, st>>>(M);
        OB_CHECK_KERNEL("k_colour_bounds", st);
        k_colour_fixup<<<1, 1, 0, st>>>(M, e->d_stats);
        OB_CHECK_KERNEL("k_colour_fixup", st);
        k_rows<<<pgrid, 128, 0, st>>>(M, B, src, usurf, e->S, cfg, e->d_stats);
        OB_CHECK_KERNEL("k_rows", st);
    }
    if (e->timing) OB_CUDA(cudaEventRecord(e->ev[3], st));
    {
        SolverArrays S = e->S;
        OB_CUDA(cudaMemsetAsync(&M.meta[6], 0, sizeof(int), st)); // grid barrier counter
        long work = max_manifolds > nb ? max_manifolds : nb;
        const void *fn = cfg.tol > 0.f ? (const void *)k_solve<true> : (const void *)k_solve<false>;
        if (cfg.tol > 0.f) OB_CUDA(cudaMemsetAsync(&M.meta[8], 0, 3 * sizeof(int), st)); // residual slots
        StepStats *d_stats = e->d_stats;
        const int *done_flag = nullptr;
        // dWorldStep on a world the size of the reference's: solve the step's LCP exactly (solver_exact.cu); when that
        // succeeds it also integrates and raises the flag that makes the sweep kernels below return at once
        const bool want_exact = e->params.exact && !(cfg.tol > 0.f) && max_manifolds > 0 && exact_solve_fits(nb);
        if (want_exact) {
            exact_solve_launch(e, M, S, B, cfg, &M.meta[12], st);
            done_flag = &M.meta[12];
        } else if (nb <= TINY_BODIES && !(cfg.tol > 0.f) && e->tiny_solver) {
            const size_t smem = (size_t)TINY_MANIFOLDS * (7 * 8 + 1) * 16 + (size_t)nb * 5 * 16;
            // per device (function attributes live in the context): once per engine, not once per process
            if (!e->tiny_attr_set) {
                OB_CUDA(cudaFuncSetAttribute(k_tiny_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                e->tiny_attr_set = true;
            }
            k_tiny_solve<<<1, TINY_THREADS, smem, st>>>(M, S, B, cfg, d_stats, &M.meta[11]);
            OB_CHECK_KERNEL("k_tiny_solve", st);
            done_flag = &M.meta[11];
        }
        int grid = coop_grid(e, fn, 256, work);
        void *args[] = {(void *)&M, (void *)&S, (void *)&B, (void *)&cfg, (void *)&d_stats, (void *)&done_flag};
        OB_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(256), args, 0, st));
        OB_CHECK_KERNEL("k_solve", st);
    }
    if (e->timing) OB_CUDA(cudaEventRecord(e->ev[4], st));
}

} // namespace ob
