// solver_env.cu -- the lane-pair island solver of batched worlds (BASELINE config 4: thousands of independent
// 128-body worlds), K5..K8 for one env in one warp.
//
// Replaces what libode does inside dWorldQuickStep (/root/reference/src/main.c:213 calls dWorldStep; north_star:
// QuickStep, 20 sweeps) for the contact joints of the reference's NearCallback (src/main.c:683-691) plus the
// app-side GetTransformMat pack (src/main.c:602-622).
//
// Why a second island kernel.  ncu of the generic island solver (solver.cu, k_env_solve<32,true>) on C4: 64 %
// issue-active but only 17.9 of 32 lanes per instruction -- a warp owns an env, a lane owns a contact, and a
// settled 128-body world has colour classes of 73, 36, 17, 12, 7, 3, 3, 3, 1 contacts, 58 % of them against the
// static plane (one body) -- so most trips run a 450-instruction two-body row update on a handful of lanes.
// Here a LANE owns ONE BODY'S HALF of a contact: a two-body contact takes an adjacent lane pair, a one-body
// contact a single lane, all packed into the same trips (lanes [0, 2T) = the colour's T two-body contacts, lanes
// [2T, 2T+O) its O one-body contacts).  Each lane computes its body's part of J.fc (same products, same order as
// ODE's SOR_LCP), the pair exchanges the two partial sums with one shuffle, both lanes form the same delta and
// clamp, and each updates its own body's accumulator.  Every floating-point operation is the one the CPU oracle
// and the global solver execute, in the same order: results are bit-identical (tests: island == global == oracle).
// Per trip a lane issues ~270 instructions instead of ~450 and a settled C4 world needs 13 half-cost trips per
// sweep instead of 12 full-cost ones.
//
// Row records (96 B per contact, six float4 planes indexed by the env-sorted slot): A = (n, flags), R1 = (r1, -),
// R2 = (r2, mu2), C = (rhsN, rhsT1, rhsT2, 1/sqrt of dPlaneSpace) * Ad, D = (AdN, AdT1, AdT2, mu), L = lambda.
// Ad*cfm is recomputed (one multiply; the island path has one surface for the whole world, so the three CFMs are
// kernel constants).  With ROWS_SMEM the planes of the first `row_cap` contacts of the env live in shared memory
// for the 20 sweeps (the rest, if any, in global memory); otherwise all rows stay in global memory (L2-resident).
#include "solver_dev.cuh"

namespace ob {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int CS2 = 2 * (OVERFLOW_COLOUR + 1) + 2; // bucket starts: (colour, two-body | one-body), overflow, end

// unit cache while colouring (16 B per unit): x,y = priority (y later: sorted position -> unit), z = packed ids
// and flags, w = colour (255 = none yet)
constexpr unsigned UC_D1 = 1u << 24, UC_D2 = 1u << 25, UC_REV = 1u << 26;

// row flags (A.w)
constexpr int RF_TWO = 1 << 28, RF_APPROX1 = 1 << 26, RF_APPROX2 = 1 << 27, RF_MU2 = 1 << 29; // bits 0..11 l1, 12..23 l2, 24..25 rows

struct EnvRowPlanes {
    float4 *A, *R1, *R2, *C, *D, *L; // global planes, indexed by sorted slot (ms + position)
};

template <bool ROWS_SMEM>
struct RowAccess {
    EnvRowPlanes g;
    float4 *sm; // shared planes of this warp: [plane][row_cap]
    int row_cap, ms;
    __device__ __forceinline__ float4 *plane(int p, int pos) const {
        if (ROWS_SMEM && pos < row_cap) return sm + p * row_cap + pos;
        float4 *base = p == 0 ? g.A : p == 1 ? g.R1 : p == 2 ? g.R2 : p == 3 ? g.C : p == 4 ? g.D : g.L;
        return base + ms + pos;
    }
};

struct HalfBody {
    V3 fl, fa;
    M3 iI;
    float invM;
};

// One row, one body's half.  A lane that holds body 2 works with the negated Jacobian J2 = -(d, r2 x d).  Negation
// commutes exactly with IEEE multiplication and with sums of negated terms, so instead of negating six components per row
// the lane carries the sign in the two scalars: sAd = -Ad in the dot (each product fc * (J * Ad) comes out as ODE's) and
// sdelta = -delta in the update (fc += delta * (M^-1 J^T) likewise).  Bit-identical to the explicit form, ~10 instructions
// per row cheaper.  Returns this body's part of J.fc, scaled by Ad component-wise like ODE's pre-scaled J.
__device__ __forceinline__ float half_dot(V3 dir, V3 r, float sAd, const HalfBody &hb, V3 &Ja) {
    Ja = cross(r, dir);
    return hb.fl.x * (dir.x * sAd) + hb.fl.y * (dir.y * sAd) + hb.fl.z * (dir.z * sAd) + hb.fa.x * (Ja.x * sAd) + hb.fa.y * (Ja.y * sAd) +
           hb.fa.z * (Ja.z * sAd);
}
__device__ __forceinline__ void half_apply(float sdelta, V3 dir, V3 Ja, HalfBody &hb) {
    const V3 iMa = mul(hb.iI, Ja);
    hb.fl.x += sdelta * (hb.invM * dir.x); hb.fl.y += sdelta * (hb.invM * dir.y); hb.fl.z += sdelta * (hb.invM * dir.z);
    hb.fa.x += sdelta * iMa.x; hb.fa.y += sdelta * iMa.y; hb.fa.z += sdelta * iMa.z;
}
// ODE's SOR_LCP row update given the two partial sums (body 1's, body 2's)
__device__ __forceinline__ float row_delta(float rhs_s, float Adcfm, float s1, float s2, bool two, float lo, float hi, float &lambda) {
    const float old_lambda = lambda;
    float delta = rhs_s - old_lambda * Adcfm;
    delta -= s1;
    if (two) delta -= s2;
    const float new_lambda = old_lambda + delta;
    if (new_lambda < lo) { delta = lo - old_lambda; lambda = lo; }
    else if (new_lambda > hi) { delta = hi - old_lambda; lambda = hi; }
    else lambda = new_lambda;
    return delta;
}

} // namespace

#ifndef OB_ENV2_THREADS
#define OB_ENV2_THREADS 32 // one warp (= one env at a time) per CTA: measured 1.09 ms vs 1.12 (64) and 1.14 (128) on C4
#endif
#ifndef OB_ENV2_WARPS_SM
#define OB_ENV2_WARPS_SM 16 // resident warps per SM the register budget is sized for: 126 registers, no spills (20 warps = 96 registers: +6 % time)
#endif
#ifndef OB_ENV2_WARPS_SM_ROWS
#define OB_ENV2_WARPS_SM_ROWS 8 // ... of the shared-memory-rows variant (shared memory allows no more)
#endif

template <bool ROWS_SMEM>
__global__ void __launch_bounds__(OB_ENV2_THREADS, (ROWS_SMEM ? OB_ENV2_WARPS_SM_ROWS : OB_ENV2_WARPS_SM) * 32 / OB_ENV2_THREADS) k_env_solve2(EnvArrays E, BodyArrays B, ContactSource src, Surface usurf, SolverArrays S,
                                                              StepConfig cfg, int spread, int fused, int row_cap,
                                                              StepStats *__restrict__ stats) {
    extern __shared__ __align__(16) unsigned char env_smem[];
    constexpr int WARPS = OB_ENV2_THREADS / 32;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mb = (E.max_bodies + 31) & ~31;
    // per warp: body region (80 B per body: colouring scratch, then accumulators + world inverse inertias), bucket
    // starts + cursors, and with ROWS_SMEM the row planes
    const size_t region = (size_t)mb * 80;
    const size_t per_warp = region + 2 * CS2 * sizeof(int) + (ROWS_SMEM ? (size_t)row_cap * 96 : 0);
    unsigned char *wbase = env_smem + (size_t)wid * per_warp;
    unsigned long long *masks = reinterpret_cast<unsigned long long *>(wbase);
    unsigned long long *prio = masks + mb;
    float4 *sm_fc = reinterpret_cast<float4 *>(wbase);
    float4 *sm_inv = sm_fc + 2 * (size_t)mb;
    uint4 *uc_sm = reinterpret_cast<uint4 *>(wbase + 16 * (size_t)mb);
    int *cs2 = reinterpret_cast<int *>(wbase + region);
    int *cursor = cs2 + CS2;
    RowAccess<ROWS_SMEM> rows;
    rows.g = EnvRowPlanes{S.q0, S.q1, S.q2, S.q3, S.q4, S.lam};
    rows.sm = reinterpret_cast<float4 *>(wbase + region + 2 * CS2 * sizeof(int));
    rows.row_cap = row_cap;
    // the three CFMs of the world's one surface (build_row: Adcfm = Ad * (cfm / h))
    const float h1 = 1.0f / cfg.h;
    const float cfmhN = ((usurf.mode & MODE_SOFT_CFM) ? usurf.soft_cfm : cfg.cfm) * h1;
    const float cfmh1 = ((usurf.mode & MODE_SLIP1) ? usurf.slip1 : cfg.cfm) * h1;
    const float cfmh2 = ((usurf.mode & MODE_SLIP2) ? usurf.slip2 : cfg.cfm) * h1;
    const int the_m = surface_rows(usurf); // rows per contact: the same for every contact of the world

    int rows1 = 0, rows2 = 0, ncont = 0, max_col = 0, max_rounds = 0, ntrips = 0, nlanes = 0;
    int *next_item = &E.fill[E.n_envs];
    for (;;) {
        // persistent warps: the next env comes from a counter, largest envs first (k_env_order)
        int item = 0;
        if (lane == 0) item = atomicAdd(next_item, 1);
        item = __shfl_sync(FULL, item, 0);
        if (item >= E.n_envs) break;
        const int env = E.order[item];
        const bool fits = E.start[E.n_envs] <= E.cap; // unit overflow (flagged by k_env_bucket): free flight
        const int ms = fits ? E.start[env] : 0, me = fits ? E.start[env + 1] : 0;
        const int nun = me - ms;
        const int fb = E.first_body[env], nbod = E.n_body[env];
        rows.ms = ms;
        if (fused)
            for (int i = lane; i < nbod; i += 32) body_prep(fb + i, B, cfg);
        __syncwarp();
        if (nun == 0) { // free flight
            if (fused)
                for (int i = lane; i < nbod; i += 32) {
                    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                    integrate_body(fb + i, B, cfg.h, z, z);
                    if (fused == 2) { B.fc[2 * (size_t)(fb + i)] = z; B.fc[2 * (size_t)(fb + i) + 1] = z; }
                }
            continue;
        }
        // ---- colouring: the rule of k_colour (solver.cu) at warp scope, on a cache of the env's units
        for (int i = lane; i < mb; i += 32) { masks[i] = 0ull; prio[i] = ~0ull; }
        for (int i = lane; i < 2 * CS2; i += 32) cs2[i] = 0;
        // the cache lives in the idle 64 B per body of the region, or (envs with more than 4 units per body slot) in the
        // k = 1 plane of q0, which per-contact units never use
        uint4 *uc = (nun <= 4 * mb) ? uc_sm : reinterpret_cast<uint4 *>(S.q0 + S.cap) + ms;
        const int trips = (nun + 31) >> 5;
        constexpr int TB = 6;
        for (int j0 = 0; j0 < trips; j0 += TB) {
            int4 r[TB];
            int l1[TB], l2[TB];
#pragma unroll
            for (int k = 0; k < TB; k++) {
                const int m = ms + lane + (j0 + k) * 32;
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
                uc[lane + (j0 + k) * 32] = make_uint4((unsigned)pr, (unsigned)(pr >> 32),
                                                      ((unsigned)l1[k] & 0xfffu) | (((unsigned)l2[k] & 0xfffu) << 12) |
                                                          ((r[k].w & REC_DYN1) ? UC_D1 : 0u) | ((r[k].w & REC_DYN2) ? UC_D2 : 0u) |
                                                          ((r[k].w & REC_REV) ? UC_REV : 0u),
                                                      255u);
            }
        }
        __syncwarp();
        int rounds = 0;
        for (;; rounds++) {
            const bool stamped = rounds < 254;
            const unsigned long long stamp = stamped ? ((unsigned long long)(254 - rounds) << 56) : 0ull;
            for (int u = lane; u < nun; u += 32) {
                const uint4 q = uc[u];
                if (q.w == 255u) {
                    const unsigned long long pr = stamp | (unsigned long long)q.x | ((unsigned long long)q.y << 32);
                    if (q.z & UC_D1) atomicMin(&prio[q.z & 0xfffu], pr);
                    if (q.z & UC_D2) atomicMin(&prio[(q.z >> 12) & 0xfffu], pr);
                }
            }
            __syncwarp();
            bool left = false;
            for (int u = lane; u < nun; u += 32) {
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
                        // bucket = (colour, one-body?) -- the overflow class keeps one bucket
                        const int bkt = c == OVERFLOW_COLOUR ? 2 * OVERFLOW_COLOUR : 2 * c + (l2 == 0xfff ? 1 : 0);
                        atomicAdd(&cs2[bkt + 1], 1);
                    } else left = true;
                }
            }
            __syncwarp();
            if (!__any_sync(FULL, left)) break;
            if (!stamped) {
                for (int u = lane; u < nun; u += 32) {
                    const uint4 q = uc[u];
                    if (q.w == 255u) {
                        if (q.z & UC_D1) prio[q.z & 0xfffu] = ~0ull;
                        if (q.z & UC_D2) prio[(q.z >> 12) & 0xfffu] = ~0ull;
                    }
                }
                __syncwarp();
            }
        }
        // ---- counting sort by (colour, two-body first)
        if (lane == 0) {
            int acc = 0;
            for (int c = 0; c < CS2; c++) { acc += cs2[c]; cs2[c] = acc; }
        }
        __syncwarp();
        int ncol = 0;
        for (int u = lane; u < nun; u += 32) {
            const uint4 q = uc[u];
            const int c = (int)q.w;
            const int bkt = c == OVERFLOW_COLOUR ? 2 * OVERFLOW_COLOUR : 2 * c + (((q.z >> 12) & 0xfffu) == 0xfff ? 1 : 0);
            const int r = atomicAdd(&cursor[bkt], 1);
            uc[cs2[bkt] + r].y = (unsigned)u; // sorted position -> unit (the priority's high word is dead now)
            if (c < OVERFLOW_COLOUR && c + 1 > ncol) ncol = c + 1;
        }
        __syncwarp();
        ncol = __reduce_max_sync(FULL, ncol);
        // ---- rows (dxJointContact::getInfo2 + QuickStep's rhs and Ad), one lane per contact
        for (int pos = lane; pos < nun; pos += 32) {
            const uint4 q = uc[uc[pos].y];
            const int l1 = (int)(q.z & 0xfffu);
            int l2 = (int)((q.z >> 12) & 0xfffu);
            if (l2 == 0xfff) l2 = -1;
            const bool two = l2 >= 0, rev = q.z & UC_REV;
            const int cslot = (int)q.x; // the priority's tie word = contact slot
            const BodyKin k1 = load_kin(B, fb + l1);
            BodyKin k2;
            if (two) k2 = load_kin(B, fb + l2);
            else {
                k2.x = v3(0.f, 0.f, 0.f); k2.lv = k2.av = k2.tv = k2.tw = k2.x;
                k2.iI = M3{k2.x, k2.x, k2.x}; k2.invM = 0.f;
            }
            const float4 pd = src.pd[cslot], ns = src.ns[cslot];
            V3 normal = v3(ns);
            if (rev) normal = -normal;
            const V3 cp = v3(pd);
            const V3 c1 = cp - k1.x;
            const V3 c2 = two ? (cp - k2.x) : v3(0.f, 0.f, 0.f);
            float erp = cfg.erp;
            if (usurf.mode & MODE_SOFT_ERP) erp = usurf.soft_erp;
            const float kk = (1.0f / cfg.h) * erp;
            float depth = pd.w - cfg.min_depth;
            if (depth < 0) depth = 0;
            float cfmN = cfg.cfm;
            if (usurf.mode & MODE_SOFT_CFM) cfmN = usurf.soft_cfm;
            float motionN = 0.f;
            if (usurf.mode & MODE_MOTIONN) motionN = usurf.motionN;
            float cN = kk * depth + motionN;
            if (cN > cfg.max_vel) cN = cfg.max_vel;
            if (usurf.mode & MODE_BOUNCE) {
                const V3 J1a = cross(c1, normal);
                float outgoing = dot(normal, k1.lv) + dot(J1a, k1.av);
                if (two) {
                    const V3 J2l = -normal;
                    const V3 J2a = -cross(c2, normal);
                    outgoing += dot(J2l, k2.lv) + dot(J2a, k2.av);
                }
                outgoing -= motionN;
                if (usurf.bounce_vel >= 0 && (-outgoing) > usurf.bounce_vel) {
                    const float newc = -usurf.bounce * outgoing + motionN;
                    if (newc > cN) cN = newc;
                }
            }
            float4 C = make_float4(0.f, 0.f, 0.f, 0.f), D = make_float4(0.f, 0.f, 0.f, 0.f);
            float unused;
            build_row(normal, c1, c2, k1, k2, two, cN, cfmN, cfg, C.x, D.x, unused);
            int flags = (l1 & 0xfff) | ((l2 & 0xfff) << 12) | (the_m << 24) | (two ? RF_TWO : 0);
            float mu2 = 0.f;
            if (the_m >= 2) {
                V3 t1, t2;
                const float psk = plane_space_k(normal);
                plane_space_with_k(normal, psk, t1, t2);
                C.w = psk;
                D.w = usurf.mu < 0 ? 0 : usurf.mu;
                const float c1v = (usurf.mode & MODE_MOTION1) ? usurf.motion1 : 0.f;
                const float cfm1 = (usurf.mode & MODE_SLIP1) ? usurf.slip1 : cfg.cfm;
                build_row(t1, c1, c2, k1, k2, two, c1v, cfm1, cfg, C.y, D.y, unused);
                if (usurf.mode & MODE_APPROX1_1) flags |= RF_APPROX1;
                if (the_m >= 3) {
                    const float c2v = (usurf.mode & MODE_MOTION2) ? usurf.motion2 : 0.f;
                    const float cfm2 = (usurf.mode & MODE_SLIP2) ? usurf.slip2 : cfg.cfm;
                    build_row(t2, c1, c2, k1, k2, two, c2v, cfm2, cfg, C.z, D.z, unused);
                    if (usurf.mode & MODE_MU2) { flags |= RF_MU2; mu2 = usurf.mu2 < 0 ? 0 : usurf.mu2; }
                    if (usurf.mode & MODE_APPROX1_2) flags |= RF_APPROX2;
                }
            }
            *rows.plane(0, pos) = make_float4(normal.x, normal.y, normal.z, __int_as_float(flags));
            *rows.plane(1, pos) = make_float4(c1.x, c1.y, c1.z, 0.f);
            *rows.plane(2, pos) = make_float4(c2.x, c2.y, c2.z, mu2);
            *rows.plane(3, pos) = C;
            *rows.plane(4, pos) = D;
            *rows.plane(5, pos) = make_float4(0.f, 0.f, 0.f, 0.f);
            S.mrec[ms + pos] = make_int4(l1, l2, 1, cslot); // the solver order export reads .w (tests)
            if (two) rows2 += the_m; else rows1 += the_m;
            ncont++;
        }
        __syncwarp();
        // ---- the colouring scratch is dead: accumulators (zero) and world inverse inertias take the region
        for (int i = lane; i < 2 * mb; i += 32) sm_fc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = lane; i < 3 * nbod; i += 32) sm_inv[i] = B.inv[3 * (size_t)fb + i];
        __syncwarp();
        // ---- SOR/PGS sweeps.  Colour c: lanes [0, 2T) hold the halves of its T two-body contacts, lanes [2T, 2T + O)
        // its O one-body contacts; colours are separated by __syncwarp only (an env is an island).
        const int ovf0 = cs2[2 * OVERFLOW_COLOUR], ovf1 = cs2[2 * OVERFLOW_COLOUR + 1];
        if (lane == 0) // diagnostics: 32-lane trips and occupied lanes of one sweep (dStepStatsB200.env_trips / env_lanes)
            for (int c = 0; c < ncol; c++) {
                const int nl = 2 * (cs2[2 * c + 1] - cs2[2 * c]) + (cs2[2 * c + 2] - cs2[2 * c + 1]);
                ntrips += (nl + 31) >> 5;
                nlanes += nl;
            }
        // one 32-lane trip: every lane updates one body's half of one contact (rows: normal, tangent 1, tangent 2)
        auto solve_trip = [&](int pos, bool active, bool neg, const float4 A, const float4 Rh, const float4 C, const float4 D) {
            float4 L = *rows.plane(5, pos);
            const int flags = __float_as_int(A.w);
            const bool two = flags & RF_TWO;
            const int b = neg ? ((flags >> 12) & 0xfff) : (flags & 0xfff);
            HalfBody hb;
            {
                const float4 a = sm_fc[b], w = sm_fc[mb + b];
                hb.fl = v3(a); hb.fa = v3(w);
                const float4 i0 = sm_inv[3 * b], i1 = sm_inv[3 * b + 1], i2 = sm_inv[3 * b + 2];
                hb.iI = M3{v3(i0), v3(i1), v3(i2)};
                hb.invM = i0.w;
            }
            const V3 n = v3(A), r = v3(Rh);
            const float sgn = neg ? -1.0f : 1.0f;
            V3 Ja;
            // normal row
            float s = half_dot(n, r, sgn * D.x, hb, Ja);
            float so = __shfl_xor_sync(FULL, s, 1);
            float delta = row_delta(C.x, D.x * cfmhN, neg ? so : s, neg ? s : so, two, 0.f, INFINITY, L.x);
            half_apply(sgn * delta, n, Ja, hb);
            if (the_m >= 2) {
                V3 t1, t2;
                plane_space_with_k(n, C.w, t1, t2);
                const float mu = D.w;
                float hi = mu, lo = -mu;
                if (flags & RF_APPROX1) { hi = fabsf(mu * L.x); lo = -hi; }
                s = half_dot(t1, r, sgn * D.y, hb, Ja);
                so = __shfl_xor_sync(FULL, s, 1);
                delta = row_delta(C.y, D.y * cfmh1, neg ? so : s, neg ? s : so, two, lo, hi, L.y);
                half_apply(sgn * delta, t1, Ja, hb);
                if (the_m >= 3) {
                    const float mu2 = (flags & RF_MU2) ? (*rows.plane(2, pos)).w : mu;
                    hi = mu2; lo = -mu2;
                    if (flags & RF_APPROX2) { hi = fabsf(mu2 * L.x); lo = -hi; }
                    s = half_dot(t2, r, sgn * D.z, hb, Ja);
                    so = __shfl_xor_sync(FULL, s, 1);
                    delta = row_delta(C.z, D.z * cfmh2, neg ? so : s, neg ? s : so, two, lo, hi, L.z);
                    half_apply(sgn * delta, t2, Ja, hb);
                }
            }
            if (active) {
                sm_fc[b] = make_float4(hb.fl.x, hb.fl.y, hb.fl.z, 0.f);
                sm_fc[mb + b] = make_float4(hb.fa.x, hb.fa.y, hb.fa.z, 0.f);
                if (!neg) *rows.plane(5, pos) = L;
            }
        };
        // lane -> (row slot, body half) of trip (c, gl0)
        auto trip_lane = [&](int c, int gl0, int &pos, bool &active, bool &neg) {
            const int p0 = cs2[2 * c], T = cs2[2 * c + 1] - p0, O = cs2[2 * c + 2] - cs2[2 * c + 1];
            const int gl = gl0 + lane;
            active = gl < 2 * T + O;
            const bool pairlane = gl < 2 * T;
            pos = active ? (pairlane ? p0 + (gl >> 1) : p0 + gl - T) : p0;
            neg = pairlane && (gl & 1);
        };
        // (Fetching the next trip's immutable row planes a whole trip ahead -- a trip table in shared memory, 16 more
        // registers -- was measured: C4 solve 1.09 -> 1.17 ms.  The loads of a trip already overlap the previous trip of
        // the other resident warps; the extra live registers and the table look-ups cost more than the stall they hide.
        // Staging the next trip's four immutable planes in shared memory with cp.async (no registers in flight; lambda one
        // trip ahead in registers) was measured as well: 1.12 -> 1.20 ms.  ncu: long-scoreboard stalls fall from 1.6 to 1.0
        // per issue but short-scoreboard ones rise by as much; the samples sit on the row's dependent chain, not on loads.)
        // (Storing r x d and I^-1 (r x d) per row half at row build instead of recomputing them every sweep -- 72 fewer
        // instructions per trip, bit-identical -- was measured too: 224 B instead of 96 B per contact from L2 every sweep,
        // solve 1.12 -> 1.35 ms.  Arithmetic is cheaper than L2 here.)
        for (int it = 0; it < cfg.iters; it++) {
            for (int c = 0; c < ncol; c++) {
                const int nl = 2 * (cs2[2 * c + 1] - cs2[2 * c]) + (cs2[2 * c + 2] - cs2[2 * c + 1]);
                for (int gl0 = 0; gl0 < nl; gl0 += 32) {
                    int pos;
                    bool active, neg;
                    trip_lane(c, gl0, pos, active, neg);
                    solve_trip(pos, active, neg, *rows.plane(0, pos), *rows.plane(neg ? 2 : 1, pos), *rows.plane(3, pos), *rows.plane(4, pos));
                }
                __syncwarp();
            }
            if (ovf1 > ovf0) {
                // contacts that found no free colour (a body with more than 64 contacts): one lane, both halves, in order
                if (lane == 0) {
                    for (int pos = ovf0; pos < ovf1; pos++) {
                        const float4 A = *rows.plane(0, pos), R1 = *rows.plane(1, pos), R2 = *rows.plane(2, pos);
                        const float4 C = *rows.plane(3, pos), D = *rows.plane(4, pos);
                        float4 L = *rows.plane(5, pos);
                        const int flags = __float_as_int(A.w);
                        const bool two = flags & RF_TWO;
                        const int b1 = flags & 0xfff, b2 = two ? ((flags >> 12) & 0xfff) : b1;
                        HalfBody h1b, h2b;
                        {
                            const float4 a = sm_fc[b1], w = sm_fc[mb + b1];
                            h1b.fl = v3(a); h1b.fa = v3(w);
                            const float4 i0 = sm_inv[3 * b1], i1 = sm_inv[3 * b1 + 1], i2 = sm_inv[3 * b1 + 2];
                            h1b.iI = M3{v3(i0), v3(i1), v3(i2)}; h1b.invM = i0.w;
                        }
                        {
                            const float4 a = sm_fc[b2], w = sm_fc[mb + b2];
                            h2b.fl = v3(a); h2b.fa = v3(w);
                            const float4 i0 = sm_inv[3 * b2], i1 = sm_inv[3 * b2 + 1], i2 = sm_inv[3 * b2 + 2];
                            h2b.iI = M3{v3(i0), v3(i1), v3(i2)}; h2b.invM = i0.w;
                        }
                        const V3 n = v3(A), r1 = v3(R1), r2 = v3(R2);
                        V3 t1 = n, t2 = n;
                        if (the_m >= 2) plane_space_with_k(n, C.w, t1, t2);
                        const float mu = D.w, mu2 = (flags & RF_MU2) ? R2.w : mu;
                        for (int row = 0; row < the_m; row++) {
                            const V3 d = row == 0 ? n : (row == 1 ? t1 : t2);
                            const float Ad = row == 0 ? D.x : (row == 1 ? D.y : D.z);
                            const float rhs = row == 0 ? C.x : (row == 1 ? C.y : C.z);
                            const float cfmh = row == 0 ? cfmhN : (row == 1 ? cfmh1 : cfmh2);
                            float lo = 0.f, hi = INFINITY;
                            if (row == 1) { hi = mu; lo = -mu; if (flags & RF_APPROX1) { hi = fabsf(mu * L.x); lo = -hi; } }
                            if (row == 2) { hi = mu2; lo = -mu2; if (flags & RF_APPROX2) { hi = fabsf(mu2 * L.x); lo = -hi; } }
                            V3 J1a, J2a;
                            const float s1 = half_dot(d, r1, Ad, h1b, J1a);
                            const float s2 = two ? half_dot(d, r2, -Ad, h2b, J2a) : 0.f;
                            float &lam = row == 0 ? L.x : (row == 1 ? L.y : L.z);
                            const float delta = row_delta(rhs, Ad * cfmh, s1, s2, two, lo, hi, lam);
                            half_apply(delta, d, J1a, h1b);
                            if (two) half_apply(-delta, d, J2a, h2b);
                        }
                        sm_fc[b1] = make_float4(h1b.fl.x, h1b.fl.y, h1b.fl.z, 0.f);
                        sm_fc[mb + b1] = make_float4(h1b.fa.x, h1b.fa.y, h1b.fa.z, 0.f);
                        if (two) {
                            sm_fc[b2] = make_float4(h2b.fl.x, h2b.fl.y, h2b.fl.z, 0.f);
                            sm_fc[mb + b2] = make_float4(h2b.fa.x, h2b.fa.y, h2b.fa.z, 0.f);
                        }
                        *rows.plane(5, pos) = L;
                    }
                }
                __syncwarp();
            }
        }
        // ---- solver tail: velocity update, dxStepBody, snapshot pack (or hand the accumulators to k_integrate)
        if (fused) {
            for (int i = lane; i < nbod; i += 32) {
                const float4 fl = sm_fc[i], fa = sm_fc[mb + i];
                integrate_body(fb + i, B, cfg.h, fl, fa);
                if (fused == 2) {
                    B.fc[2 * (size_t)(fb + i)] = fl;
                    B.fc[2 * (size_t)(fb + i) + 1] = fa;
                }
            }
        } else {
            for (int i = lane; i < nbod; i += 32) {
                B.fc[2 * (size_t)(fb + i)] = sm_fc[i];
                B.fc[2 * (size_t)(fb + i) + 1] = sm_fc[mb + i];
            }
        }
        __syncwarp();
        max_col = max(max_col, ncol);
        max_rounds = max(max_rounds, rounds + 1);
        if (lane == 0 && ovf1 > ovf0) atomicAdd(&stats->n_overflow, ovf1 - ovf0);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        rows1 += __shfl_xor_sync(FULL, rows1, o);
        rows2 += __shfl_xor_sync(FULL, rows2, o);
        ncont += __shfl_xor_sync(FULL, ncont, o);
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
        if (ntrips) { atomicAdd(&stats->env_trips, ntrips); atomicAdd(&stats->env_lanes, nlanes); }
        if (blockIdx.x == 0 && threadIdx.x == 0) stats->solver_iters = cfg.iters;
    }
    (void)WARPS;
}

// Launch the lane-pair island solver.  Preconditions (checked by the caller, solver_step): per-contact units, every
// env's bodies one contiguous index range of at most 160, one surface for the whole world.
// rows_smem: 0 = rows in global memory (L2-resident), N > 0 = the first N contacts of an env in shared memory.
void env_solve2_launch(Engine *e, const EnvArrays &E, const BodyArrays &B, const ContactSource &src, const Surface &usurf,
                       const SolverArrays &S, const StepConfig &cfg, int fused, int rows_smem, cudaStream_t st) {
    const int mb = (E.max_bodies + 31) & ~31;
    constexpr int WARPS = OB_ENV2_THREADS / 32;
    const size_t per_warp = (size_t)mb * 80 + 2 * CS2 * sizeof(int) + (size_t)rows_smem * 96;
    const size_t smem = WARPS * per_warp;
    unsigned grid = (unsigned)((E.n_envs + WARPS - 1) / WARPS);
    int per_sm = 0;
    if (rows_smem > 0) {
        OB_CUDA(cudaFuncSetAttribute(k_env_solve2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        OB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_env_solve2<true>, OB_ENV2_THREADS, smem));
        if (per_sm > 0) grid = std::min(grid, (unsigned)(per_sm * e->num_sms));
        k_env_solve2<true><<<grid, OB_ENV2_THREADS, smem, st>>>(E, B, src, usurf, S, cfg, e->colour_spread, fused, rows_smem, e->d_stats);
    } else {
        OB_CUDA(cudaFuncSetAttribute(k_env_solve2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        OB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_env_solve2<false>, OB_ENV2_THREADS, smem));
        if (per_sm > 0) grid = std::min(grid, (unsigned)(per_sm * e->num_sms));
        k_env_solve2<false><<<grid, OB_ENV2_THREADS, smem, st>>>(E, B, src, usurf, S, cfg, e->colour_spread, fused, 0, e->d_stats);
    }
}

} // namespace ob
