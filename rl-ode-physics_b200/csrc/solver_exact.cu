// solver_exact.cu -- dWorldStep as the reference means it (/root/reference/src/main.c:213): the step's LCP solved
// EXACTLY, island by island, for worlds the size of the reference's own scene.
//
// libode's dWorldStep builds A = J M^-1 J^T + cfm/h per island and hands  A lambda = rhs + w,  lo <= lambda <= hi  to its
// Dantzig solver (dSolveLCP); A is symmetric positive definite for cfm > 0, so the bounded LCP has exactly one solution
// and any exact method returns it.  Here: islands by label propagation over the contact graph (static and kinematic
// ends do not connect, as in dxProcessIslands their rows do not couple), one CTA per island, A assembled in double from
// the same float J and M^-1 J^T rows the sweeps use, and block principal pivoting (Judice & Pires) with a dense Cholesky
// of the free block -- the method the test oracle's exact mode restates on the host (tests compare the two).
// Limits: worlds of at most EX_MAX_BODIES bodies and EX_MAX_UNITS solver units, islands of at most EX_MAX_ROWS rows,
// no dContactApprox1 rows (their bounds depend on another row's lambda: not an LCP with fixed bounds).  Anything
// beyond falls back to the SOR/PGS sweeps inside the same step and says so in dStepStatsB200.exact_status.
#include "solver_dev.cuh"

namespace ob {

constexpr int EX_MAX_BODIES = 1024, EX_MAX_UNITS = 4096, EX_MAX_ROWS = 384;
constexpr int EX_THREADS = 512;

struct ExactArrays {
    int *label;      // [EX_MAX_BODIES] island label of a body (smallest body index of the island), -1: not dynamic
    int *desc;       // row descriptors, island-major: slot | contact << 16 | row << 20
    int *isl_row0;   // [EX_MAX_BODIES + 1] first row of island i
    size_t *isl_mat; // [EX_MAX_BODIES + 1] offset of island i's matrix
    int *isl_label;  // [EX_MAX_BODIES] label of island i
    int *meta;       // [0] number of islands, [1] status (0 ok, 1 too large, 2 unsupported rows, 3 no convergence),
                     // [2] total rows, [3] largest island, [4] pivoting rounds (max over islands)
    double *A, *C;   // island matrices and Cholesky work copies
    size_t cap_mat;
    int cap_rows;
};

// ---- islands + row enumeration (one CTA)
__global__ void __launch_bounds__(1024) k_exact_islands(ManifoldArrays M, SolverArrays S, BodyArrays B, ExactArrays X) {
    __shared__ int label[EX_MAX_BODIES];
    __shared__ int cnt[EX_MAX_BODIES];
    __shared__ short uisl[EX_MAX_UNITS];
    __shared__ short urows[EX_MAX_UNITS];
    __shared__ short urow0[EX_MAX_UNITS];
    __shared__ int changed, n_isl, status;
    const int tid = threadIdx.x, T = blockDim.x;
    const int nb = B.n, n = *M.count;
    if (tid == 0) { status = (nb > EX_MAX_BODIES || n > EX_MAX_UNITS) ? 1 : 0; n_isl = 0; }
    __syncthreads();
    if (status) {
        if (tid == 0) { X.meta[0] = 0; X.meta[1] = 1; X.meta[2] = 0; X.meta[3] = 0; X.meta[4] = 0; }
        return;
    }
    for (int b = tid; b < nb; b += T) { label[b] = B.pos[b].w > 0.f ? b : -1; cnt[b] = 0; }
    __syncthreads();
    // label propagation with pointer jumping: every dynamic body ends with the smallest body index of its island
    for (;;) {
        if (tid == 0) changed = 0;
        __syncthreads();
        for (int s = tid; s < n; s += T) {
            const int4 r = S.mrec[s];
            if (r.y < 0) continue;
            const int l1 = label[r.x], l2 = label[r.y];
            if (l1 < 0 || l2 < 0 || l1 == l2) continue;
            const int m = min(l1, l2);
            if (l1 != m) { atomicMin(&label[r.x], m); changed = 1; }
            if (l2 != m) { atomicMin(&label[r.y], m); changed = 1; }
        }
        __syncthreads();
        for (int b = tid; b < nb; b += T) {
            int l = label[b];
            if (l >= 0) {
                int ll = label[l];
                while (ll != l) { l = ll; ll = label[l]; }
                if (l != label[b]) { label[b] = l; changed = 1; }
            }
        }
        __syncthreads();
        if (!changed) break;
        __syncthreads();
    }
    // rows per unit, unit -> label
    for (int s = tid; s < n; s += T) {
        const int4 r = S.mrec[s];
        int lab = label[r.x];
        if (lab < 0 && r.y >= 0) lab = label[r.y];
        int rows = 0;
        if (lab >= 0)
            for (int k = 0; k < r.z; k++) {
                const int lf = __float_as_int(S.lam[(size_t)k * S.cap + s].w);
                if (lf & 0x30) status = 2; // dContactApprox1: bounds that follow another lambda
                rows += lf & 0xf;
            }
        uisl[s] = (short)lab; // label for now
        urows[s] = (short)rows;
        if (lab >= 0 && rows) atomicAdd(&cnt[lab], rows);
    }
    __syncthreads();
    // compact the labels into island ids (ascending label), row and matrix offsets
    if (tid == 0) {
        int ni = 0, row = 0, big = 0;
        size_t mat = 0;
        for (int b = 0; b < nb; b++) {
            if (cnt[b] > 0) {
                X.isl_row0[ni] = row; X.isl_mat[ni] = mat; X.isl_label[ni] = b;
                if (cnt[b] > EX_MAX_ROWS) status = status ? status : 1;
                row += cnt[b];
                mat += (size_t)cnt[b] * (size_t)cnt[b];
                big = max(big, cnt[b]);
                cnt[b] = ni; // label -> island id
                ni++;
            } else cnt[b] = -1;
        }
        X.isl_row0[ni] = row; X.isl_mat[ni] = mat;
        if (row > X.cap_rows || mat > X.cap_mat) status = status ? status : 1;
        n_isl = ni;
        X.meta[0] = ni; X.meta[2] = row; X.meta[3] = big; X.meta[4] = 0;
    }
    __syncthreads();
    if (tid == 0) X.meta[1] = status;
    for (int b = tid; b < nb; b += T) X.label[b] = label[b];
    if (status) return;
    for (int s = tid; s < n; s += T) uisl[s] = (short)((uisl[s] >= 0 && urows[s]) ? cnt[uisl[s]] : -1);
    __syncthreads();
    // position of each unit's rows inside its island, in slot order (deterministic): one warp scans all units per island
    const int lane = tid & 31, wid = tid >> 5, nw = T >> 5;
    for (int i = wid; i < n_isl; i += nw) {
        int running = 0;
        for (int base = 0; base < n; base += 32) {
            const int s = base + lane;
            const bool mine = s < n && uisl[s] == i;
            const int r = mine ? urows[s] : 0;
            int inc = r;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            if (mine) urow0[s] = (short)(running + inc - r);
            running += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
    __syncthreads();
    for (int s = tid; s < n; s += T) {
        const int i = uisl[s];
        if (i < 0) continue;
        int at = X.isl_row0[i] + urow0[s];
        const int nc = S.mrec[s].z;
        for (int k = 0; k < nc; k++) {
            const int rows = __float_as_int(S.lam[(size_t)k * S.cap + s].w) & 0xf;
            for (int t = 0; t < rows; t++) X.desc[at++] = s | (k << 16) | (t << 20);
        }
    }
}

// ---- one island per CTA: rows, A, block principal pivoting, accumulators
__global__ void __launch_bounds__(EX_THREADS) k_exact_solve(SolverArrays S, BodyArrays B, ExactArrays X, StepConfig cfg) {
    if (X.meta[1] != 0) return;
    const int isl = blockIdx.x;
    if (isl >= X.meta[0]) return;
    extern __shared__ __align__(16) unsigned char ex_smem[];
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = T >> 5;
    const int row0 = X.isl_row0[isl], m = X.isl_row0[isl + 1] - row0;
    double *A = X.A + X.isl_mat[isl], *C = X.C + X.isl_mat[isl];
    // shared: iMJ (12 doubles per row), J (12 floats per row), b, x, w (double), lo, hi (float), bodies (2 ints), state, free list.
    // M^-1 J^T is formed in DOUBLE from the float J and the float inverse masses / world inverse inertias: with the float
    // products of the sweeps, A carries 1e-7 relative errors, which the nearly singular contact block of a resting box
    // (four coplanar contacts, only cfm/h on the diagonal) amplifies to 4e-4 m/s -- found by fuzzing against the oracle.
    double *iMJ = reinterpret_cast<double *>(ex_smem);
    float *J = reinterpret_cast<float *>(iMJ + 12 * EX_MAX_ROWS);
    double *bvec = reinterpret_cast<double *>(J + 12 * EX_MAX_ROWS);
    double *x = bvec + EX_MAX_ROWS, *wv = x + EX_MAX_ROWS, *rr = wv + EX_MAX_ROWS;
    float *lo = reinterpret_cast<float *>(rr + EX_MAX_ROWS), *hi = lo + EX_MAX_ROWS, *cfmh = hi + EX_MAX_ROWS;
    int *jb = reinterpret_cast<int *>(cfmh + EX_MAX_ROWS); // 2 per row
    short *F = reinterpret_cast<short *>(jb + 2 * EX_MAX_ROWS);
    char *st = reinterpret_cast<char *>(F + EX_MAX_ROWS);
    __shared__ int s_nf, s_ninf, s_last, s_best, s_p, s_all, s_done, s_rounds;

    // rows: J and M^-1 J^T exactly as the sweeps rebuild them (solver.cu solve_row), scalars un-scaled from Ad
    for (int r = tid; r < m; r += T) {
        const int d = X.desc[row0 + r];
        const int s = d & 0xffff, k = (d >> 16) & 0xf, t = (d >> 20) & 0x3;
        const size_t si = (size_t)k * S.cap + s;
        const int4 rec = S.mrec[s];
        const float4 q0 = S.q0[si], q1 = S.q1[si], q2 = S.q2[si], q3 = S.q3[si], q4 = S.q4[si];
        const int lflags = __float_as_int(S.lam[si].w);
        const V3 n = v3(q0), c1 = v3(q1), c2 = v3(q2);
        V3 dir = n;
        float rhs_s = q0.w, Ad = q1.w, Adcfm = q2.w, l = 0.f, h = INFINITY;
        if (t > 0) {
            V3 t1, t2;
            plane_space_with_k(n, q3.w, t1, t2);
            const float mu = q4.w;
            if (t == 1) { dir = t1; rhs_s = q3.x; Ad = q3.y; Adcfm = q3.z; l = -mu; h = mu; }
            else {
                const float mu2 = (lflags & 0x40) ? S.q5[si].x : mu;
                dir = t2; rhs_s = q4.x; Ad = q4.y; Adcfm = q4.z; l = -mu2; h = mu2;
            }
        }
        const bool two = rec.y >= 0;
        const V3 J1a = cross(c1, dir);
        float *Jr = J + 12 * r;
        double *im = iMJ + 12 * r;
        Jr[0] = dir.x; Jr[1] = dir.y; Jr[2] = dir.z; Jr[3] = J1a.x; Jr[4] = J1a.y; Jr[5] = J1a.z;
        for (int q = 6; q < 12; q++) { Jr[q] = 0.f; im[q] = 0.0; }
        V3 J2a = v3(0.f, 0.f, 0.f);
        if (two) {
            const V3 J2l = -dir;
            J2a = -cross(c2, dir);
            Jr[6] = J2l.x; Jr[7] = J2l.y; Jr[8] = J2l.z; Jr[9] = J2a.x; Jr[10] = J2a.y; Jr[11] = J2a.z;
        }
        for (int s1 = 0; s1 < (two ? 2 : 1); s1++) {
            const int b = s1 ? rec.y : rec.x;
            const float4 i0 = B.inv[3 * b], i1 = B.inv[3 * b + 1], i2 = B.inv[3 * b + 2];
            const float *Jh = Jr + 6 * s1;
            double *ih = im + 6 * s1;
            for (int q = 0; q < 3; q++) ih[q] = (double)i0.w * (double)Jh[q];
            ih[3] = (double)i0.x * Jh[3] + (double)i0.y * Jh[4] + (double)i0.z * Jh[5];
            ih[4] = (double)i1.x * Jh[3] + (double)i1.y * Jh[4] + (double)i1.z * Jh[5];
            ih[5] = (double)i2.x * Jh[3] + (double)i2.y * Jh[4] + (double)i2.z * Jh[5];
        }
        jb[2 * r] = rec.x; jb[2 * r + 1] = rec.y;
        bvec[r] = (double)rhs_s / (double)Ad;   // the sweeps store rhs * Ad (ODE's pre-scaled rows)
        cfmh[r] = (float)((double)Adcfm / (double)Ad);
        lo[r] = l; hi[r] = h;
        st[r] = 0; x[r] = 0.0;
    }
    if (tid == 0) { s_best = m + 1; s_p = 10; s_done = 0; s_rounds = 0; }
    __syncthreads();
    // A = J M^-1 J^T + cfm/h, accumulated in double from the float rows (oracle: order_mode 3)
    for (int e = tid; e < m * m; e += T) {
        const int i = e / m, j = e - i * m;
        const double *im = iMJ + 12 * i;
        const float *Jj = J + 12 * j;
        double a = 0.0;
        for (int s1 = 0; s1 < 2; s1++) {
            const int bi = jb[2 * i + s1];
            if (bi < 0) continue;
            for (int s2 = 0; s2 < 2; s2++)
                if (jb[2 * j + s2] == bi)
                    for (int k = 0; k < 6; k++) a += im[6 * s1 + k] * (double)Jj[6 * s2 + k];
        }
        if (i == j) a += (double)cfmh[i];
        A[e] = a;
    }
    __syncthreads();

    const double eps = 1e-11;
    const int max_rounds = 60 + 4 * m;
    for (int round = 0; round < max_rounds; round++) {
        // free set, bound values
        if (tid == 0) {
            int nf = 0;
            for (int i = 0; i < m; i++) {
                if (st[i] == 0) F[nf++] = (short)i;
                else x[i] = st[i] == 1 ? (double)lo[i] : (double)hi[i];
            }
            s_nf = nf;
        }
        __syncthreads();
        const int nf = s_nf;
        // r_F = b_F - A_F,bound x_bound ; C = A_FF (lower triangle)
        for (int a = wid; a < nf; a += nw) {
            const int i = F[a];
            double sum = 0.0;
            for (int j = lane; j < m; j += 32)
                if (st[j]) {
                    const double xj = x[j];
                    if (xj != 0.0) sum += A[(size_t)i * m + j] * xj;
                }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (lane == 0) rr[a] = bvec[i] - sum;
            for (int c = lane; c <= a; c += 32) C[(size_t)a * nf + c] = A[(size_t)i * m + F[c]];
        }
        __syncthreads();
        // Cholesky C = L L^T (right-looking, in place, lower)
        for (int k = 0; k < nf; k++) {
            if (tid == 0) {
                double s = C[(size_t)k * nf + k];
                if (s <= 0) s = 1e-300;
                C[(size_t)k * nf + k] = sqrt(s);
            }
            __syncthreads();
            const double dkk = C[(size_t)k * nf + k];
            for (int i = k + 1 + tid; i < nf; i += T) C[(size_t)i * nf + k] /= dkk;
            __syncthreads();
            for (int i = k + 1 + wid; i < nf; i += nw) {
                const double lik = C[(size_t)i * nf + k];
                for (int j = k + 1 + lane; j <= i; j += 32) C[(size_t)i * nf + j] -= lik * C[(size_t)j * nf + k];
            }
            __syncthreads();
        }
        // forward and back substitution (column oriented: one barrier per column)
        for (int a = 0; a < nf; a++) {
            if (tid == 0) rr[a] /= C[(size_t)a * nf + a];
            __syncthreads();
            const double ra = rr[a];
            for (int i = a + 1 + tid; i < nf; i += T) rr[i] -= C[(size_t)i * nf + a] * ra;
            __syncthreads();
        }
        for (int a = nf - 1; a >= 0; a--) {
            if (tid == 0) rr[a] /= C[(size_t)a * nf + a];
            __syncthreads();
            const double ra = rr[a];
            for (int i = tid; i < a; i += T) rr[i] -= C[(size_t)a * nf + i] * ra;
            __syncthreads();
        }
        for (int a = tid; a < nf; a += T) x[F[a]] = rr[a];
        if (tid == 0) { s_ninf = 0; s_last = -1; }
        __syncthreads();
        // w = A x - b, infeasibilities
        for (int i = wid; i < m; i += nw) {
            double sum = 0.0;
            for (int j = lane; j < m; j += 32) sum += A[(size_t)i * m + j] * x[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (lane == 0) {
                const double w = sum - bvec[i];
                wv[i] = w;
                bool bad;
                if (st[i] == 0) bad = (x[i] < (double)lo[i] - eps) || (x[i] > (double)hi[i] + eps);
                else if (st[i] == 1) bad = w < -eps;
                else bad = w > eps;
                if (bad) { atomicAdd(&s_ninf, 1); atomicMax(&s_last, i); }
            }
        }
        __syncthreads();
        if (tid == 0) {
            s_rounds = round + 1;
            if (s_ninf == 0) s_done = 1;
            else {
                // Judice-Pires: flip every infeasible row while that keeps reducing their number, p more tries after a
                // failure, then only the infeasible row of largest index (guarantees termination)
                if (s_ninf < s_best) { s_best = s_ninf; s_p = 10; s_all = 1; }
                else if (s_p > 0) { s_p--; s_all = 1; }
                else s_all = 0;
            }
        }
        __syncthreads();
        if (s_done) break;
        for (int i = tid; i < m; i += T) {
            if (!s_all && i != s_last) continue;
            if (st[i] == 0) {
                if (x[i] < (double)lo[i] - eps) st[i] = 1;
                else if (x[i] > (double)hi[i] + eps) st[i] = 2;
            } else if ((st[i] == 1 && wv[i] < -eps) || (st[i] == 2 && wv[i] > eps)) st[i] = 0;
        }
        __syncthreads();
    }
    if (tid == 0) {
        if (!s_done) atomicMax(&X.meta[1], 3);
        atomicMax(&X.meta[4], s_rounds);
    }
    // lambda out (row order of the sweeps' arrays) and the accumulators fc = M^-1 J^T lambda of the island's bodies,
    // each body summed by one thread over the island's rows in order
    for (int r = tid; r < m; r += T) {
        const int d = X.desc[row0 + r];
        const int s = d & 0xffff, k = (d >> 16) & 0xf, t = (d >> 20) & 0x3;
        float *lam = reinterpret_cast<float *>(&S.lam[(size_t)k * S.cap + s]);
        lam[t] = (float)x[r];
    }
    const int lab = X.isl_label[isl];
    for (int b = tid; b < B.n; b += T) {
        if (X.label[b] != lab) continue;
        double acc[6] = {0, 0, 0, 0, 0, 0};
        for (int r = 0; r < m; r++) {
            if (jb[2 * r] == b)
                for (int k = 0; k < 6; k++) acc[k] += x[r] * iMJ[12 * r + k];
            if (jb[2 * r + 1] == b)
                for (int k = 0; k < 6; k++) acc[k] += x[r] * iMJ[12 * r + 6 + k];
        }
        B.fc[2 * b] = make_float4((float)acc[0], (float)acc[1], (float)acc[2], 0.f);
        B.fc[2 * b + 1] = make_float4((float)acc[3], (float)acc[4], (float)acc[5], 0.f);
    }
}

// ---- exact solve succeeded: velocity update + dxStepBody + snapshot pack, and tell the sweep kernels to stand down;
// otherwise clear what the islands wrote so the sweeps start from fc = 0
__global__ void __launch_bounds__(256) k_exact_finish(BodyArrays B, ExactArrays X, StepConfig cfg, int *__restrict__ done_flag,
                                                       StepStats *__restrict__ stats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int status = X.meta[1];
    if (i == 0) {
        *done_flag = status == 0 ? 1 : 0;
        stats->exact_status = status;
        stats->n_islands = X.meta[0];
        stats->max_island_rows = X.meta[3];
        stats->pivot_rounds = X.meta[4];
        if (status == 0) stats->solver_iters = 0;
    }
    if (i >= B.n) return;
    if (status == 0) integrate_body(i, B, cfg.h, B.fc[2 * i], B.fc[2 * i + 1]);
    else { B.fc[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f); B.fc[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f); }
}

static size_t exact_smem_bytes() {
    return (size_t)EX_MAX_ROWS * (12 * sizeof(double) + 12 * sizeof(float) + 4 * sizeof(double) + 3 * sizeof(float) + 2 * sizeof(int) + sizeof(short) + 1) + 64;
}

// Try the exact solve of the rows k_rows just built.  *done_flag = 1 when it succeeded (integration included).
void exact_solve_launch(Engine *e, const ManifoldArrays &M, const SolverArrays &S, const BodyArrays &B, const StepConfig &cfg,
                        int *done_flag, cudaStream_t st) {
    if (!e->ex_label) {
        const size_t cap_rows = (size_t)EX_MAX_UNITS * 8 * 3, cap_mat = (size_t)EX_MAX_ROWS * cap_rows / 8;
        OB_CUDA(ob_malloc(&e->ex_label, sizeof(int) * EX_MAX_BODIES));
        OB_CUDA(ob_malloc(&e->ex_desc, sizeof(int) * cap_rows));
        OB_CUDA(ob_malloc(&e->ex_isl_row0, sizeof(int) * (EX_MAX_BODIES + 1)));
        OB_CUDA(ob_malloc(&e->ex_isl_mat, sizeof(size_t) * (EX_MAX_BODIES + 1)));
        OB_CUDA(ob_malloc(&e->ex_isl_label, sizeof(int) * EX_MAX_BODIES));
        OB_CUDA(ob_malloc(&e->ex_meta, sizeof(int) * 8));
        OB_CUDA(ob_malloc(&e->ex_A, sizeof(double) * cap_mat));
        OB_CUDA(ob_malloc(&e->ex_C, sizeof(double) * cap_mat));
        e->ex_cap_mat = cap_mat;
        e->ex_cap_rows = (int)cap_rows;
        OB_CUDA(cudaFuncSetAttribute(k_exact_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)exact_smem_bytes()));
    }
    ExactArrays X;
    X.label = e->ex_label; X.desc = e->ex_desc; X.isl_row0 = e->ex_isl_row0; X.isl_mat = e->ex_isl_mat;
    X.isl_label = e->ex_isl_label; X.meta = e->ex_meta; X.A = e->ex_A; X.C = e->ex_C;
    X.cap_mat = e->ex_cap_mat; X.cap_rows = e->ex_cap_rows;
    k_exact_islands<<<1, 1024, 0, st>>>(M, S, B, X);
    OB_CHECK_KERNEL("k_exact_islands", st);
    const int grid = std::min(B.n, EX_MAX_BODIES);
    k_exact_solve<<<(unsigned)std::max(grid, 1), EX_THREADS, exact_smem_bytes(), st>>>(S, B, X, cfg);
    OB_CHECK_KERNEL("k_exact_solve", st);
    k_exact_finish<<<(unsigned)((B.n + 255) / 256), 256, 0, st>>>(B, X, cfg, done_flag, e->d_stats);
    OB_CHECK_KERNEL("k_exact_finish", st);
}

bool exact_solve_fits(int n_bodies) { return n_bodies <= EX_MAX_BODIES; } // units and island sizes are checked on the device

} // namespace ob
