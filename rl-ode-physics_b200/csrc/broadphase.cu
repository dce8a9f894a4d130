// broadphase.cu -- K1..K3 of SURVEY.md section 2: AABBs, uniform-grid keys, sorted sweep that
// emits every pair passing ODE's collideAABBs filter (what dxHashSpace::collide hands to the
// reference's NearCallback, /root/reference/src/main.c:212 + :674).
//
// Geoms whose AABB is finite and no larger than the largest dynamic geom (extent M) are "small": they are
// binned by AABB centre into a uniform grid of cell M/2.  A small geom j can only overlap geom i if its
// centre lies in i's AABB grown by M/2, so i visits the cells of that window only -- (e_i/M + 1.5)^3 cell
// volumes of M^3 instead of the 27 of a cell-M grid with a 3x3x3 neighbourhood; on the settled 1 M-body pile
// that is 3.6x fewer candidate tests.  Everything else (planes, trimeshes, the
// reference's 100x1x100 floor box) is "big" and is tested against every geom, like ODE's big-box
// list.  Pairs are produced in two deterministic passes (count, scan, fill) grouped by collider
// class; no atomics decide a position, so the pair list is bit-reproducible.
#include "dev.cuh"

namespace ob {

constexpr float GRID_CELL_FRACTION = 0.5f;

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// K1: refresh geom poses from their bodies, compute ODE's AABBs, reduce the grid inputs.
// acc[0..2] = min centre, acc[3..5] = max centre, acc[6] = max extent (order-encoded floats)
__global__ void __launch_bounds__(256) k_geom_update(GeomArrays g, const float4 *__restrict__ b_pos,
                                                      const float4 *__restrict__ b_R, MeshTable meshes,
                                                      float big_extent, unsigned *__restrict__ acc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float cmin[3] = {INFINITY, INFINITY, INFINITY}, cmax[3] = {-INFINITY, -INFINITY, -INFINITY}, ext = 0.f;
    if (i < g.n) {
        const int type = g.type[i], body = g.body[i];
        float4 p4;
        M3 R;
        if (body >= 0) {
            p4 = b_pos[body];
            R = load_m3(b_R, body);
            g.pos[i] = make_float4(p4.x, p4.y, p4.z, 0.f);
            store_m3(g.R, i, R);
        } else {
            p4 = g.pos[i];
            R = load_m3(g.R, i);
        }
        const float4 d = g.dims[i];
        float lo[3], hi[3];
        const float pos[3] = {p4.x, p4.y, p4.z};
        if (type == G_SPHERE) {
            for (int k = 0; k < 3; k++) { lo[k] = pos[k] - d.x; hi[k] = pos[k] + d.x; }
        } else if (type == G_BOX) {
            const V3 rows[3] = {R.r0, R.r1, R.r2};
            for (int k = 0; k < 3; k++) {
                float range = 0.5f * (fabsf(rows[k].x * d.x) + fabsf(rows[k].y * d.y) + fabsf(rows[k].z * d.z));
                lo[k] = pos[k] - range;
                hi[k] = pos[k] + range;
            }
        } else if (type == G_PLANE) {
            for (int k = 0; k < 3; k++) { lo[k] = -INFINITY; hi[k] = INFINITY; }
            if (d.y == 0.0f && d.z == 0.0f) {
                lo[0] = (d.x > 0) ? -INFINITY : -d.w;
                hi[0] = (d.x > 0) ? d.w : INFINITY;
            } else if (d.x == 0.0f && d.z == 0.0f) {
                lo[1] = (d.y > 0) ? -INFINITY : -d.w;
                hi[1] = (d.y > 0) ? d.w : INFINITY;
            } else if (d.x == 0.0f && d.y == 0.0f) {
                lo[2] = (d.z > 0) ? -INFINITY : -d.w;
                hi[2] = (d.z > 0) ? d.w : INFINITY;
            }
        } else if (type == G_TRIMESH) {
            const MeshInfo mi = meshes.m[g.mesh[i]];
            const float c[3] = {0.5f * (mi.lo[0] + mi.hi[0]), 0.5f * (mi.lo[1] + mi.hi[1]), 0.5f * (mi.lo[2] + mi.hi[2])};
            const float e[3] = {0.5f * (mi.hi[0] - mi.lo[0]), 0.5f * (mi.hi[1] - mi.lo[1]), 0.5f * (mi.hi[2] - mi.lo[2])};
            const V3 rows[3] = {R.r0, R.r1, R.r2};
            for (int k = 0; k < 3; k++) {
                float wc = pos[k] + (rows[k].x * c[0] + rows[k].y * c[1] + rows[k].z * c[2]);
                float range = fabsf(rows[k].x * e[0]) + fabsf(rows[k].y * e[1]) + fabsf(rows[k].z * e[2]);
                lo[k] = wc - range;
                hi[k] = wc + range;
            }
        } else {
            for (int k = 0; k < 3; k++) { lo[k] = INFINITY; hi[k] = -INFINITY; }
        }
        g.amin[i] = make_float4(lo[0], lo[1], lo[2], 0.f);
        g.amax[i] = make_float4(hi[0], hi[1], hi[2], 0.f);
        if (g.alive[i] && body >= 0 && (type == G_SPHERE || type == G_BOX)) {
            float e = fmaxf(hi[0] - lo[0], fmaxf(hi[1] - lo[1], hi[2] - lo[2]));
            if (e <= big_extent && isfinite(e)) {
                ext = e;
                for (int k = 0; k < 3; k++) {
                    float c = 0.5f * (lo[k] + hi[k]);
                    if (isfinite(c)) { cmin[k] = c; cmax[k] = c; }
                }
            }
        }
    }
    // warp shuffle reduction, then one set of atomics per CTA (not per warp: 32k warps hammering
    // seven addresses serialise in the L2 atomic unit)
    __shared__ float red[8][7];
    for (int k = 0; k < 3; k++) { cmin[k] = warp_min(cmin[k]); cmax[k] = warp_max(cmax[k]); }
    ext = warp_max(ext);
    const int wid = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        for (int k = 0; k < 3; k++) { red[wid][k] = cmin[k]; red[wid][3 + k] = cmax[k]; }
        red[wid][6] = ext;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        const int k = threadIdx.x;
        float v = red[0][k];
        for (int w = 1; w < 8; w++) v = (k < 3) ? fminf(v, red[w][k]) : fmaxf(v, red[w][k]);
        if (k < 3) { if (v < INFINITY) atomicMin(&acc[k], f2ord(v)); }
        else if (k < 6) { if (v > -INFINITY) atomicMax(&acc[k], f2ord(v)); }
        else if (v > 0.f) atomicMax(&acc[6], f2ord(v));
    }
}

// The reduction slots are re-armed on the device by whichever one-thread kernel follows k_geom_update (no
// host-to-device copy on the tick path: copy-engine work of the compute stream queues behind the application's
// own snapshot / force transfers -- measured: collide 0.33 -> 0.52 ms while a 64 MB D2H copy was in flight).
__device__ __forceinline__ void acc_rearm(unsigned *acc) {
    acc[0] = acc[1] = acc[2] = 0xffffffffu;
    acc[3] = acc[4] = acc[5] = acc[6] = acc[7] = 0u;
}
__global__ void k_acc_init(unsigned *acc) { acc_rearm(acc); }

// one thread: turn the reductions into grid parameters, shrink the grid until it fits the table
__global__ void k_grid_params(unsigned *__restrict__ acc, GridParams *__restrict__ gp, int n_envs,
                              int cap_cells, int n_geoms, BroadCounters *__restrict__ bc) {
    float ext = ord2f(acc[6]);
    float lo[3], hi[3];
    bool any = acc[6] != f2ord(0.f) && ext > 0.f;
    for (int k = 0; k < 3; k++) { lo[k] = ord2f(acc[k]); hi[k] = ord2f(acc[3 + k]); }
    if (!any || !(lo[0] <= hi[0])) {
        ext = 1.f;
        for (int k = 0; k < 3; k++) { lo[k] = 0.f; hi[k] = 0.f; }
    }
    float cell = ext * (1.01f * GRID_CELL_FRACTION);
    int d[3];
    for (int iter = 0; iter < 64; iter++) {
        double tot = (double)n_envs;
        for (int k = 0; k < 3; k++) {
            float span = (hi[k] - lo[k]) / cell;
            d[k] = (span < 1.0e6f) ? (int)floorf(span) + 1 : 1000000;
            tot *= (double)d[k];
        }
        if (tot <= (double)cap_cells) break;
        cell *= 1.26f;
    }
    gp->ox = lo[0]; gp->oy = lo[1]; gp->oz = lo[2];
    gp->cell = cell; gp->inv_cell = 1.0f / cell;
    gp->dx = d[0]; gp->dy = d[1]; gp->dz = d[2];
    gp->per_env = d[0] * d[1] * d[2];
    gp->n_envs = n_envs;
    gp->small_extent = any ? ext : 0.f;
    bc->first_big = n_geoms;
    bc->first_dead = n_geoms;
    bc->n_pairs = 0;
    bc->n_bb = 0;
    acc_rearm(acc);
}

// cell = this fraction of the largest small extent M (measured on the settled pile, C3: 1 -> 0.5 cuts the
// candidate tests 3.6x for 2.3x more cell-table look-ups; 0.25 would need 5x more look-ups than tests it saves)
// K2: grid key per geom (BIG / DEAD sentinels sort to the tail)
__global__ void __launch_bounds__(256) k_cell_keys(GeomArrays g, const GridParams *__restrict__ gpp, int cap_cells,
                                                    uint32_t *__restrict__ keys, int *__restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    const GridParams gp = *gpp;
    const float4 lo = g.amin[i], hi = g.amax[i];
    uint32_t key;
    const int env = g.env[i];
    if (!g.alive[i]) {
        key = (uint32_t)cap_cells + 1u;
    } else {
        float e = fmaxf(hi.x - lo.x, fmaxf(hi.y - lo.y, hi.z - lo.z));
        bool small_geom = isfinite(e) && e <= gp.small_extent && (env >= 0 || gp.n_envs == 1) && env < gp.n_envs;
        if (!small_geom) {
            key = (uint32_t)cap_cells;
        } else {
            float cx = 0.5f * (lo.x + hi.x), cy = 0.5f * (lo.y + hi.y), cz = 0.5f * (lo.z + hi.z);
            int ix = (int)floorf((cx - gp.ox) * gp.inv_cell);
            int iy = (int)floorf((cy - gp.oy) * gp.inv_cell);
            int iz = (int)floorf((cz - gp.oz) * gp.inv_cell);
            ix = min(max(ix, 0), gp.dx - 1);
            iy = min(max(iy, 0), gp.dy - 1);
            iz = min(max(iz, 0), gp.dz - 1);
            int e0 = env < 0 ? 0 : env;
            key = (uint32_t)(((e0 * gp.dz + iz) * gp.dy + iy) * gp.dx + ix);
        }
    }
    keys[i] = key;
    idx[i] = i;
}

// K3a: gather the sweep records into sorted order, mark cell ranges and the big/dead boundaries
__global__ void __launch_bounds__(256) k_sorted_records(GeomArrays g, const uint32_t *__restrict__ keys,
                                                         const int *__restrict__ idx, int cap_cells,
                                                         float4 *__restrict__ s_min, float4 *__restrict__ s_max,
                                                         uint4 *__restrict__ s_flt, int *__restrict__ cell_start,
                                                         int *__restrict__ cell_end, BroadCounters *__restrict__ bc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    const int gi = idx[i];
    const uint32_t key = keys[i];
    float4 lo = g.amin[gi], hi = g.amax[gi];
    lo.w = __int_as_float(gi);
    hi.w = __int_as_float(g.body[gi]);
    s_min[i] = lo;
    s_max[i] = hi;
    s_flt[i] = make_uint4(g.cat[gi], g.col[gi], (unsigned)g.env[gi], (unsigned)g.type[gi]);
    const uint32_t prev = (i > 0) ? keys[i - 1] : 0xffffffffu;
    const uint32_t next = (i + 1 < g.n) ? keys[i + 1] : 0xffffffffu;
    if (key < (uint32_t)cap_cells) {
        if (i == 0 || prev != key) cell_start[key] = i;
        if (next != key) cell_end[key] = i + 1;
    } else if (key == (uint32_t)cap_cells) {
        if (i == 0 || prev < (uint32_t)cap_cells) bc->first_big = i;
    } else {
        if (i == 0 || prev <= (uint32_t)cap_cells) {
            bc->first_dead = i;
            if (i == 0 || prev < (uint32_t)cap_cells) bc->first_big = i; // no big geoms at all
        }
    }
}

__global__ void __launch_bounds__(256) k_cell_clear(int n, const uint32_t *__restrict__ keys, int cap_cells,
                                                     int *__restrict__ cell_end) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t key = keys[i];
    if (key < (uint32_t)cap_cells) cell_end[key] = 0;
}

__device__ __forceinline__ int pair_class(int ta, int tb) { // ta <= tb
    if (ta == G_SPHERE) {
        if (tb == G_SPHERE) return PC_SPHERE_SPHERE;
        if (tb == G_BOX) return PC_SPHERE_BOX;
        if (tb == G_PLANE) return PC_SPHERE_PLANE;
        if (tb == G_TRIMESH) return PC_SPHERE_TRIMESH;
    } else if (ta == G_BOX) {
        if (tb == G_BOX) return PC_BOX_BOX;
        if (tb == G_PLANE) return PC_BOX_PLANE;
        if (tb == G_TRIMESH) return PC_SPHERE_TRIMESH; // the trimesh kernel takes spheres and boxes
    }
    return PC_NONE;
}

// ODE collideAABBs: same-body, category/collide bits, non-strict AABB overlap; plus the env rule.
// Returns the pair class or -1.
__device__ __forceinline__ int test_pair(const float4 &lo1, const float4 &hi1, const uint4 &f1, const float4 &lo2,
                                         const float4 &hi2, const uint4 &f2) {
    if (lo1.x > hi2.x || hi1.x < lo2.x || lo1.y > hi2.y || hi1.y < lo2.y || lo1.z > hi2.z || hi1.z < lo2.z) return -1;
    const int b1 = __float_as_int(hi1.w), b2 = __float_as_int(hi2.w);
    if (b1 == b2 && b1 >= 0) return -1;
    const int e1 = (int)f1.z, e2 = (int)f2.z;
    if (e1 >= 0 && e2 >= 0 && e1 != e2) return -1;
    if (!((f1.x & f2.y) || (f2.x & f1.y))) return -1;
    int ta = (int)f1.w, tb = (int)f2.w;
    if (ta > tb) { int t = ta; ta = tb; tb = t; }
    return pair_class(ta, tb);
}

// The count pass also parks each thread's first SWEEP_TCAP hits (k-major, coalesced) so that the fill
// pass is a pure compaction for almost every thread instead of a second traversal of the grid.
// (16: in a settled pile a geom emits 4-7 pairs on average and a third of them more than 8; one such lane makes
// its whole warp wait for a second traversal -- C3: fill pass 253 us at 8)

struct PairSink {
    int cnt[PC_COUNT];
    int total;
};

// Pair positions without a scan over 7 counters per geom: the count pass also reduces each class over its
// 128-thread block (7 sums per block); only those sums are scanned; the fill pass re-derives each thread's
// offsets as (scanned base of its block and class) + (prefix of the per-thread counts inside the block).
constexpr int SWEEP_THREADS = 128;

__device__ __forceinline__ void block_class_sums(const int (&v)[PC_COUNT], int *__restrict__ blk, int nblk) {
    __shared__ int wsum[PC_COUNT][SWEEP_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < PC_COUNT; c++) {
        const int t = __reduce_add_sync(0xffffffffu, v[c]);
        if (lane == 0) wsum[c][wid] = t;
    }
    __syncthreads();
    if (threadIdx.x < PC_COUNT) {
        int t = 0;
        for (int w = 0; w < SWEEP_THREADS / 32; w++) t += wsum[threadIdx.x][w];
        blk[threadIdx.x * nblk + blockIdx.x] = t;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) blk[PC_COUNT * nblk] = 0; // the scan's total slot
}

__device__ __forceinline__ void block_offsets(const int *__restrict__ cnt, const int *__restrict__ blkoff, int n, int i, int nblk,
                                              int (&off)[PC_COUNT]) {
    __shared__ int wsum[PC_COUNT][SWEEP_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int v[PC_COUNT], inc[PC_COUNT];
#pragma unroll
    for (int c = 0; c < PC_COUNT; c++) {
        v[c] = i < n ? cnt[c * n + i] : 0;
        int x = v[c];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += t;
        }
        inc[c] = x;
        if (lane == 31) wsum[c][wid] = x;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < PC_COUNT; c++) {
        int base = blkoff[c * nblk + blockIdx.x];
        for (int w = 0; w < wid; w++) base += wsum[c][w];
        off[c] = base + inc[c] - v[c];
    }
}

template <bool FILL>
__device__ __forceinline__ void emit(int cls, int i, int n, const float4 &lo1, const uint4 &f1, const float4 &lo2,
                                     const uint4 &f2, PairSink &sink, const int *__restrict__ off,
                                     int2 *__restrict__ pairs, int cap_pairs, int2 *__restrict__ tmp) {
    int ga = __float_as_int(lo1.w), gb = __float_as_int(lo2.w);
    const int ta = (int)f1.w, tb = (int)f2.w;
    // canonical callback order: lower class first, then lower geom id
    if (ta > tb || (ta == tb && ga > gb)) { int t = ga; ga = gb; gb = t; }
    if (FILL) {
        const int pos = off[cls] + sink.cnt[cls]; // off: this thread's offsets (block_offsets)
        if (pos < cap_pairs) pairs[pos] = make_int2(ga, gb);
    } else if (sink.total < SWEEP_TCAP) {
        tmp[(size_t)sink.total * n + i] = make_int2(ga | (cls << 28), gb);
    }
    sink.cnt[cls]++;
    sink.total++;
}

// K3b: the sweep. Thread i owns sorted record i and visits only records after it in sort order: the rest
// of its own cell row and the later rows of its window (see the file header), each a contiguous run of the
// sorted records; then the big list.  Consecutive lanes hold neighbouring geoms, so a warp's runs overlap
// and its loads hit the same L1 lines.
template <bool FILL>
__global__ void __launch_bounds__(SWEEP_THREADS) k_sweep(int n, const uint32_t *__restrict__ keys,
                                                const float4 *__restrict__ s_min, const float4 *__restrict__ s_max,
                                                const uint4 *__restrict__ s_flt, const int *__restrict__ cell_start,
                                                const int *__restrict__ cell_end, const GridParams *__restrict__ gpp,
                                                const BroadCounters *__restrict__ bc, int *__restrict__ cnt,
                                                int *__restrict__ blk, const int *__restrict__ blkoff,
                                                int2 *__restrict__ pairs, int cap_pairs, int2 *__restrict__ tmp,
                                                int *__restrict__ tot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nblk = gridDim.x;
    const int first_big = bc->first_big, first_dead = bc->first_dead;
    PairSink sink;
#pragma unroll
    for (int c = 0; c < PC_COUNT; c++) sink.cnt[c] = 0;
    sink.total = 0;
    int off[PC_COUNT];
    bool traverse = i < n;
    if (FILL) {
        block_offsets(cnt, blkoff, n, i, nblk, off); // every thread of the block takes part
        if (i < n) {
            const int t = tot[i];
            if (t <= SWEEP_TCAP) { // every hit of this thread was parked by the count pass: compact, in order
                for (int k = 0; k < t; k++) {
                    const int2 e = tmp[(size_t)k * n + i];
                    const int cls = e.x >> 28;
                    const int pos = off[cls] + sink.cnt[cls]++;
                    if (pos < cap_pairs) pairs[pos] = make_int2(e.x & 0x0fffffff, e.y);
                }
                traverse = false;
            }
        }
    }
    if (traverse && i < first_dead) {
        const float4 lo1 = s_min[i], hi1 = s_max[i];
        const uint4 f1 = s_flt[i];
        if (i < first_big) {
            const GridParams gp = *gpp;
            const int key = (int)keys[i];
            const int x = key % gp.dx;
            const int y = (key / gp.dx) % gp.dy;
            const int z = (key / (gp.dx * gp.dy)) % gp.dz;
            const int envbase = (key / gp.per_env) * gp.per_env;
            // window of cells that can hold the CENTRE of a small geom overlapping mine: my AABB grown by M/2 (the
            // slack covers the rounding of the centres and of the cell arithmetic; a wider window only costs tests,
            // the exact filter is test_pair)
            const float hw = 0.5f * gp.small_extent * 1.001f;
            const float sx = hw + 2e-6f * (fabsf(lo1.x) + fabsf(hi1.x)), sy = hw + 2e-6f * (fabsf(lo1.y) + fabsf(hi1.y)),
                        sz = hw + 2e-6f * (fabsf(lo1.z) + fabsf(hi1.z));
            const int x0 = min(max((int)floorf((lo1.x - sx - gp.ox) * gp.inv_cell), 0), x);
            const int x1 = max(min((int)floorf((hi1.x + sx - gp.ox) * gp.inv_cell), gp.dx - 1), x);
            const int y0 = min(max((int)floorf((lo1.y - sy - gp.oy) * gp.inv_cell), 0), y);
            const int y1 = max(min((int)floorf((hi1.y + sy - gp.oy) * gp.inv_cell), gp.dy - 1), y);
            const int z1 = max(min((int)floorf((hi1.z + sz - gp.oz) * gp.inv_cell), gp.dz - 1), z);
            // records after me in sort order (key = env, z, y, x): the rest of my own row, then every row (zz, yy) of
            // the window that sorts after (z, y); each row is one contiguous run of the sorted records
            for (int zz = z; zz <= z1; zz++) {
                for (int yy = (zz == z) ? y : y0; yy <= y1; yy++) {
                    const int rowbase = envbase + (zz * gp.dy + yy) * gp.dx;
                    int s = 0, e = 0; // empty run unless a cell of the row is occupied
                    for (int xx = x0; xx <= x1; xx++) {
                        const int ce = cell_end[rowbase + xx];
                        if (ce) {
                            if (e == 0) s = cell_start[rowbase + xx];
                            e = ce;
                        }
                    }
                    if (zz == z && yy == y) s = max(s, i + 1);
                    for (int j = s; j < e; j++) {
                        const float4 lo2 = s_min[j], hi2 = s_max[j];
                        if (lo1.x > hi2.x || hi1.x < lo2.x || lo1.y > hi2.y || hi1.y < lo2.y || lo1.z > hi2.z || hi1.z < lo2.z)
                            continue; // the filter record is only fetched for overlapping boxes
                        const uint4 f2 = s_flt[j];
                        int cls = test_pair(lo1, hi1, f1, lo2, hi2, f2);
                        if (cls >= 0) emit<FILL>(cls, i, n, lo1, f1, lo2, f2, sink, off, pairs, cap_pairs, tmp);
                    }
                }
            }
            for (int j = first_big; j < first_dead; j++) {
                const float4 lo2 = s_min[j], hi2 = s_max[j];
                const uint4 f2 = s_flt[j];
                int cls = test_pair(lo1, hi1, f1, lo2, hi2, f2);
                if (cls >= 0) emit<FILL>(cls, i, n, lo1, f1, lo2, f2, sink, off, pairs, cap_pairs, tmp);
            }
        } else {
            for (int j = i + 1; j < first_dead; j++) {
                const float4 lo2 = s_min[j], hi2 = s_max[j];
                const uint4 f2 = s_flt[j];
                int cls = test_pair(lo1, hi1, f1, lo2, hi2, f2);
                if (cls >= 0) emit<FILL>(cls, i, n, lo1, f1, lo2, f2, sink, off, pairs, cap_pairs, tmp);
            }
        }
    }
    if (!FILL) {
        if (i < n) {
#pragma unroll
            for (int c = 0; c < PC_COUNT; c++) cnt[c * n + i] = sink.cnt[c];
            tot[i] = sink.total;
        }
        block_class_sums(sink.cnt, blk, nblk);
    }
}


// ---- all pairs per env -------------------------------------------------------------------------
// Batched worlds have a few hundred geoms each, so "which geoms share a world" already is the
// broadphase partition: no keys, no sort, no cell table.  Thread i owns geom i and tests the next
// floor((c-1)/2) geoms of its env's index range cyclically (plus the opposite one for even c, from the
// lower half only), which visits each unordered pair exactly once with equal trip counts across a warp;
// then the geoms shared by all envs.  A bounding-sphere record (16 B) screens each candidate before
// the two AABB loads.  Same collideAABBs filter, same count -> scan -> fill emission as the grid path.
__global__ void __launch_bounds__(256) k_env_bounds(GeomArrays g, float4 *__restrict__ cr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    const float4 lo = g.amin[i], hi = g.amax[i];
    const float cx = 0.5f * (lo.x + hi.x), cy = 0.5f * (lo.y + hi.y), cz = 0.5f * (lo.z + hi.z);
    const float hx = 0.5f * (hi.x - lo.x), hy = 0.5f * (hi.y - lo.y), hz = 0.5f * (hi.z - lo.z);
    // inflated so that rounding can never reject a pair whose AABBs overlap
    const float r = sqrtf(hx * hx + hy * hy + hz * hz) * 1.0001f + 1e-6f * (fabsf(cx) + fabsf(cy) + fabsf(cz)) + 1e-6f;
    cr[i] = make_float4(cx, cy, cz, r);
}

struct GeomRec {
    float4 lo, hi; // lo.w = geom index, hi.w = body
    uint4 f;       // cat, col, env, type
};

__device__ __forceinline__ GeomRec load_rec(const GeomArrays &g, int j, bool single) {
    GeomRec r;
    r.lo = g.amin[j]; r.hi = g.amax[j];
    r.lo.w = __int_as_float(j);
    r.hi.w = __int_as_float(g.body[j]);
    r.f = make_uint4(g.cat[j], g.col[j], single ? 0u : (unsigned)g.env[j], (unsigned)g.type[j]);
    return r;
}

template <bool FILL>
__global__ void __launch_bounds__(SWEEP_THREADS) k_env_sweep(GeomArrays g, const float4 *__restrict__ cr,
                                                    const int *__restrict__ efirst, const int *__restrict__ ecount,
                                                    const int *__restrict__ shared, int n_shared, int single,
                                                    int *__restrict__ cnt, int *__restrict__ blk, const int *__restrict__ blkoff,
                                                    int2 *__restrict__ pairs, int cap_pairs, int2 *__restrict__ tmp,
                                                    int *__restrict__ tot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = g.n;
    const int nblk = gridDim.x;
    PairSink sink;
#pragma unroll
    for (int c = 0; c < PC_COUNT; c++) sink.cnt[c] = 0;
    sink.total = 0;
    int off[PC_COUNT];
    bool traverse = i < n;
    if (FILL) {
        block_offsets(cnt, blkoff, n, i, nblk, off);
        if (i < n) {
            const int t = tot[i];
            if (t <= SWEEP_TCAP) {
                for (int k = 0; k < t; k++) {
                    const int2 e = tmp[(size_t)k * n + i];
                    const int cls = e.x >> 28;
                    const int pos = off[cls] + sink.cnt[cls]++;
                    if (pos < cap_pairs) pairs[pos] = make_int2(e.x & 0x0fffffff, e.y);
                }
                traverse = false;
            }
        }
    }
    if (traverse && g.alive[i]) {
        const GeomRec me = load_rec(g, i, single != 0);
        const float4 c1 = cr[i];
        const int env = (int)me.f.z;
        auto visit2 = [&](int j, const float4 c2) {
            const float dx = c1.x - c2.x, dy = c1.y - c2.y, dz = c1.z - c2.z, rr = c1.w + c2.w;
            if (dx * dx + dy * dy + dz * dz > rr * rr) return; // NaN/inf (planes) fall through to the AABB test
            const float4 lo2 = g.amin[j], hi2 = g.amax[j];
            if (me.lo.x > hi2.x || me.hi.x < lo2.x || me.lo.y > hi2.y || me.hi.y < lo2.y || me.lo.z > hi2.z ||
                me.hi.z < lo2.z)
                return;
            if (!g.alive[j]) return;
            const GeomRec o = load_rec(g, j, single != 0);
            const int cls = test_pair(me.lo, me.hi, me.f, o.lo, o.hi, o.f);
            if (cls >= 0) emit<FILL>(cls, i, n, me.lo, me.f, o.lo, o.f, sink, off, pairs, cap_pairs, tmp);
        };
        auto visit = [&](int j) { visit2(j, cr[j]); };
        if (env >= 0) {
            const int first = efirst[env], c = ecount[env], li = i - first;
            const int half = (c - 1) >> 1;
            // four bounding-sphere records in flight per step (the loads are independent; the tests keep their order)
            for (int d = 1; d <= half; d += 4) {
                int jj[4];
                float4 cc[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    int lj = li + d + k;
                    if (lj >= c) lj -= c;
                    jj[k] = first + lj;
                    cc[k] = (d + k <= half) ? cr[jj[k]] : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (d + k <= half) visit2(jj[k], cc[k]);
            }
            if (!(c & 1) && li < (c >> 1)) visit(first + li + (c >> 1));
            for (int s = 0; s < n_shared; s++) visit(shared[s]);
        } else {
            for (int s = 0; s < n_shared; s++) {
                const int j = shared[s];
                if (j > i) visit(j);
            }
        }
    }
    if (!FILL) {
        if (i < n) {
#pragma unroll
            for (int c = 0; c < PC_COUNT; c++) cnt[c * n + i] = sink.cnt[c];
            tot[i] = sink.total;
        }
        block_class_sums(sink.cnt, blk, nblk);
    }
}

// ---- sort and sweep per env ----------------------------------------------------------------------
// Envs of at most 128 geoms (BASELINE config 4: 128 bodies per world; the reference's own 72-geom scene): one CTA per
// env.  The CTA sorts the env's boxes by their lower x bound in shared memory (rank sort: 128 comparisons per thread) and thread t
// sweeps from sorted position t forward until a box starts beyond its own upper x bound -- ~7 candidates per geom in a
// settled 128-body world instead of the 64 of the all-pairs sweep above (ncu: that sweep was compute-bound, 68 % SM
// throughput at 30 lanes per instruction).  Geoms shared by all envs (planes) are tested by every geom's thread; the pairs
// among the shared geoms themselves come from one extra CTA.  Same collideAABBs filter, same count -> scan -> fill
// emission (per-thread counters are indexed by GEOM, so the fill pass finds them whatever the sorted order was), same
// pair SET as the other broadphases (tested against them and against the oracle's hash space).
template <bool FILL>
__global__ void __launch_bounds__(SWEEP_THREADS) k_env_sap(GeomArrays g, const int *__restrict__ efirst, const int *__restrict__ ecount,
                                                            const int *__restrict__ shared, int n_shared, int single, int n_envs,
                                                            int *__restrict__ cnt, int *__restrict__ blk, const int *__restrict__ blkoff,
                                                            int2 *__restrict__ pairs, int cap_pairs, int2 *__restrict__ tmp,
                                                            int *__restrict__ tot) {
    __shared__ __align__(16) float key[SWEEP_THREADS];
    __shared__ float skey[SWEEP_THREADS]; // keys in sorted order
    __shared__ int sidx[SWEEP_THREADS];
    __shared__ float4 slo[SWEEP_THREADS], shi[SWEEP_THREADS];
    __shared__ uint4 sflt[SWEEP_THREADS];
    __shared__ int spos[SWEEP_THREADS];
    const int t = threadIdx.x, n = g.n, nblk = gridDim.x;
    const bool tail = (int)blockIdx.x >= n_envs; // the CTA of the shared geoms
    const int first = tail ? 0 : efirst[blockIdx.x];
    const int c = tail ? n_shared : ecount[blockIdx.x];
    const int gi = t < c ? (tail ? shared[t] : first + t) : -1;
    PairSink sink;
#pragma unroll
    for (int k = 0; k < PC_COUNT; k++) sink.cnt[k] = 0;
    sink.total = 0;
    int off[PC_COUNT];
    bool traverse = gi >= 0;
    if (FILL) {
        block_offsets(cnt, blkoff, n, gi >= 0 ? gi : n, nblk, off);
        if (gi >= 0) {
            const int tt = tot[gi];
            if (tt <= SWEEP_TCAP) { // every hit of this geom was parked by the count pass: compact, in order
                for (int k = 0; k < tt; k++) {
                    const int2 e = tmp[(size_t)k * n + gi];
                    const int cls = e.x >> 28;
                    const int pos = off[cls] + sink.cnt[cls]++;
                    if (pos < cap_pairs) pairs[pos] = make_int2(e.x & 0x0fffffff, e.y);
                }
                traverse = false;
            }
        }
        if (!__syncthreads_or(traverse)) return; // nobody of this env has to walk again
    }
    // members: boxes, filter words; dead or absent slots get an empty box that sorts last
    {
        float4 lo = make_float4(INFINITY, INFINITY, INFINITY, __int_as_float(-1)), hi = make_float4(-INFINITY, -INFINITY, -INFINITY, __int_as_float(-1));
        uint4 f = make_uint4(0u, 0u, 0u, 0u);
        if (gi >= 0 && g.alive[gi]) {
            const GeomRec r = load_rec(g, gi, single != 0);
            lo = r.lo; hi = r.hi; f = r.f;
        }
        slo[t] = lo; shi[t] = hi; sflt[t] = f;
        key[t] = lo.x;
    }
    __syncthreads();
    // rank sort of (key, member) ascending: every thread counts the members that sort before its own (128 broadcast reads,
    // one barrier -- ncu had a 28-step bitonic network at twice the instructions)
    {
        const float kt = key[t];
        int rank = 0;
        const float4 *k4 = reinterpret_cast<const float4 *>(key);
#pragma unroll 8
        for (int q = 0; q < SWEEP_THREADS / 4; q++) {
            const float4 kk = k4[q];
            rank += (kk.x < kt || (kk.x == kt && 4 * q + 0 < t)) ? 1 : 0;
            rank += (kk.y < kt || (kk.y == kt && 4 * q + 1 < t)) ? 1 : 0;
            rank += (kk.z < kt || (kk.z == kt && 4 * q + 2 < t)) ? 1 : 0;
            rank += (kk.w < kt || (kk.w == kt && 4 * q + 3 < t)) ? 1 : 0;
        }
        __syncthreads(); // everybody has read the unsorted keys
        spos[t] = rank;
        sidx[rank] = t;
        skey[rank] = kt;
    }
    __syncthreads();
    if (traverse) {
        // thread t keeps member t (the per-geom counters are indexed by its geom) and sweeps from that member's sorted position
        const float4 lo1 = slo[t], hi1 = shi[t];
        const uint4 f1 = sflt[t];
        if (__float_as_int(lo1.w) >= 0) {
            for (int q = spos[t] + 1; q < SWEEP_THREADS; q++) {
                if (skey[q] > hi1.x) break; // every later box starts beyond mine (empty slots: +inf)
                const int m = sidx[q];
                const float4 lo2 = slo[m], hi2 = shi[m];
                if (lo1.y > hi2.y || hi1.y < lo2.y || lo1.z > hi2.z || hi1.z < lo2.z) continue;
                const uint4 f2 = sflt[m];
                const int cls = test_pair(lo1, hi1, f1, lo2, hi2, f2);
                if (cls >= 0) emit<FILL>(cls, gi, n, lo1, f1, lo2, f2, sink, off, pairs, cap_pairs, tmp);
            }
            if (!tail)
                for (int s = 0; s < n_shared; s++) {
                    const int j = shared[s];
                    if (!g.alive[j]) continue;
                    const GeomRec o = load_rec(g, j, single != 0);
                    const int cls = test_pair(lo1, hi1, f1, o.lo, o.hi, o.f);
                    if (cls >= 0) emit<FILL>(cls, gi, n, lo1, f1, o.lo, o.f, sink, off, pairs, cap_pairs, tmp);
                }
        }
    }
    if (!FILL) {
        if (gi >= 0) {
#pragma unroll
            for (int k = 0; k < PC_COUNT; k++) cnt[k * n + gi] = sink.cnt[k];
            tot[gi] = sink.total;
        }
        block_class_sums(sink.cnt, blk, nblk);
    }
}

__global__ void k_env_counters(BroadCounters *__restrict__ bc, GridParams *__restrict__ gp, int n_alive, int n_shared,
                               int n_envs, unsigned *__restrict__ acc) {
    acc_rearm(acc);
    bc->first_dead = n_alive;
    bc->first_big = n_alive - n_shared;
    bc->n_pairs = 0;
    bc->n_bb = 0;
    gp->cell = 0.f; gp->dx = gp->dy = gp->dz = 0; gp->n_envs = n_envs;
}

__global__ void k_pairs_finish(int nblk, const int *__restrict__ off, int cap_pairs, BroadCounters *__restrict__ bc,
                               StepStats *__restrict__ stats, const GridParams *__restrict__ gp) {
    int total = off[PC_COUNT * nblk]; // off: exclusive scan of the per-block class sums
    int flags = 0;
    if (total > cap_pairs) { total = cap_pairs; flags |= SF_PAIR_OVERFLOW; }
    bc->n_pairs = total;
    for (int c = 0; c < PC_COUNT; c++) {
        int s = off[c * nblk];
        if (s > cap_pairs) s = cap_pairs;
        bc->class_start[c] = s;
    }
    bc->class_start[PC_COUNT] = total;
    stats->n_geoms = bc->first_dead;
    stats->n_big = bc->first_dead - bc->first_big;
    stats->n_pairs = total;
    stats->flags = flags;
    for (int c = 0; c < PC_COUNT; c++) stats->class_count[c] = bc->class_start[c + 1] - bc->class_start[c];
    stats->cell_size = gp->cell;
    stats->grid_dims[0] = gp->dx; stats->grid_dims[1] = gp->dy; stats->grid_dims[2] = gp->dz;
}

// dCollide(o1, o2) outside a space traversal: refresh the poses, then hand the narrowphase a one-pair list
__global__ void k_single_pair(GeomArrays g, int g1, int g2, int2 *__restrict__ pairs, BroadCounters *__restrict__ bc,
                              StepStats *__restrict__ stats, unsigned *__restrict__ acc) {
    acc_rearm(acc);
    int ta = g.type[g1], tb = g.type[g2];
    int ga = g1, gb = g2;
    if (ta > tb || (ta == tb && ga > gb)) { int t = ga; ga = gb; gb = t; t = ta; ta = tb; tb = t; }
    const int cls = (g.alive[g1] && g.alive[g2]) ? pair_class(ta, tb) : PC_NONE;
    pairs[0] = make_int2(ga, gb);
    bc->n_pairs = 1;
    bc->n_bb = 0;
    bc->first_big = bc->first_dead = g.n;
    for (int c = 0; c <= PC_COUNT; c++) bc->class_start[c] = c <= cls ? 0 : 1;
    stats->n_pairs = 1;
    stats->flags = 0;
    for (int c = 0; c < PC_COUNT; c++) stats->class_count[c] = c == cls ? 1 : 0;
}

void broadphase_acc_init(BroadPhase &bp, cudaStream_t st) {
    k_acc_init<<<1, 1, 0, st>>>(bp.acc);
    OB_CHECK_KERNEL("k_acc_init", st);
}

void broadphase_single_pair(BroadPhase &bp, GeomArrays g, const float4 *b_pos, const float4 *b_R, MeshTable meshes, float big_extent,
                            int g1, int g2, StepStats *d_stats, cudaStream_t st) {
    const unsigned nb = (unsigned)((g.n + 255) / 256);
    k_geom_update<<<nb, 256, 0, st>>>(g, b_pos, b_R, meshes, big_extent, bp.acc);
    OB_CHECK_KERNEL("k_geom_update", st);
    k_single_pair<<<1, 1, 0, st>>>(g, g1, g2, bp.pairs, bp.counters, d_stats, bp.acc);
    OB_CHECK_KERNEL("k_single_pair", st);
}

void broadphase_run(BroadPhase &bp, GeomArrays g, const float4 *b_pos, const float4 *b_R, MeshTable meshes,
                    int n_envs, float big_extent, const EnvBroad &eb, StepStats *d_stats, cudaStream_t st) {
    const int n = g.n;
    if (n == 0) {
        OB_CUDA(cudaMemsetAsync(bp.counters, 0, sizeof(BroadCounters), st));
        OB_CUDA(cudaMemsetAsync(d_stats, 0, sizeof(StepStats), st));
        return;
    }
    const unsigned nb = (unsigned)((n + 255) / 256);
    k_geom_update<<<nb, 256, 0, st>>>(g, b_pos, b_R, meshes, big_extent, bp.acc);
    OB_CHECK_KERNEL("k_geom_update", st);
    if (eb.enabled) {
        const unsigned nb2 = (unsigned)((n + SWEEP_THREADS - 1) / SWEEP_THREADS);
        if (eb.max_count <= SWEEP_THREADS && eb.n_shared <= SWEEP_THREADS && eb.sap &&
            (long)PC_COUNT * (eb.n_envs + 2) + 1 <= bp.cap_blk) {
            // sort and sweep, one CTA per env (+ one for the shared geoms)
            const unsigned nbe = (unsigned)eb.n_envs + 1u;
            k_env_counters<<<1, 1, 0, st>>>(bp.counters, bp.gp, eb.n_alive, eb.n_shared, n_envs, bp.acc);
            OB_CHECK_KERNEL("k_env_counters", st);
            k_env_sap<false><<<nbe, SWEEP_THREADS, 0, st>>>(g, eb.first, eb.count, eb.shared, eb.n_shared, eb.single, eb.n_envs, bp.cnt,
                                                           bp.blk, nullptr, nullptr, 0, bp.sweep_tmp, bp.sweep_tot);
            OB_CHECK_KERNEL("k_env_sap", st);
            scan_exclusive(bp.blk, bp.blk, (long)PC_COUNT * nbe + 1, nullptr, nullptr, bp.scan, st);
            k_env_sap<true><<<nbe, SWEEP_THREADS, 0, st>>>(g, eb.first, eb.count, eb.shared, eb.n_shared, eb.single, eb.n_envs, bp.cnt,
                                                          nullptr, bp.blk, bp.pairs, bp.cap_pairs, bp.sweep_tmp, bp.sweep_tot);
            OB_CHECK_KERNEL("k_env_sap", st);
            k_pairs_finish<<<1, 1, 0, st>>>((int)nbe, bp.blk, bp.cap_pairs, bp.counters, d_stats, bp.gp);
            OB_CHECK_KERNEL("k_pairs_finish", st);
            return;
        }
        float4 *cr = bp.s_min;
        k_env_bounds<<<nb, 256, 0, st>>>(g, cr);
        OB_CHECK_KERNEL("k_env_bounds", st);
        k_env_counters<<<1, 1, 0, st>>>(bp.counters, bp.gp, eb.n_alive, eb.n_shared, n_envs, bp.acc);
        OB_CHECK_KERNEL("k_env_counters", st);
        k_env_sweep<false><<<nb2, SWEEP_THREADS, 0, st>>>(g, cr, eb.first, eb.count, eb.shared, eb.n_shared, eb.single, bp.cnt,
                                                          bp.blk, nullptr, nullptr, 0, bp.sweep_tmp, bp.sweep_tot);
        OB_CHECK_KERNEL("k_env_sweep", st);
        scan_exclusive(bp.blk, bp.blk, (long)PC_COUNT * nb2 + 1, nullptr, nullptr, bp.scan, st);
        k_env_sweep<true><<<nb2, SWEEP_THREADS, 0, st>>>(g, cr, eb.first, eb.count, eb.shared, eb.n_shared, eb.single, bp.cnt,
                                                         nullptr, bp.blk, bp.pairs, bp.cap_pairs, bp.sweep_tmp, bp.sweep_tot);
        OB_CHECK_KERNEL("k_env_sweep", st);
        k_pairs_finish<<<1, 1, 0, st>>>((int)nb2, bp.blk, bp.cap_pairs, bp.counters, d_stats, bp.gp);
        OB_CHECK_KERNEL("k_pairs_finish", st);
        return;
    }
    k_grid_params<<<1, 1, 0, st>>>(bp.acc, bp.gp, n_envs, bp.cap_cells, n, bp.counters);
    OB_CHECK_KERNEL("k_grid_params", st);
    k_cell_keys<<<nb, 256, 0, st>>>(g, bp.gp, bp.cap_cells, bp.keys, bp.idx);
    OB_CHECK_KERNEL("k_cell_keys", st);
    sort_pairs(bp.keys, bp.idx, n, nullptr, bp.key_bits, bp.sort, st);
    k_sorted_records<<<nb, 256, 0, st>>>(g, bp.keys, bp.idx, bp.cap_cells, bp.s_min, bp.s_max, bp.s_flt, bp.cell_start,
                                         bp.cell_end, bp.counters);
    OB_CHECK_KERNEL("k_sorted_records", st);
    const unsigned nb2 = (unsigned)((n + SWEEP_THREADS - 1) / SWEEP_THREADS);
    k_sweep<false><<<nb2, SWEEP_THREADS, 0, st>>>(n, bp.keys, bp.s_min, bp.s_max, bp.s_flt, bp.cell_start, bp.cell_end, bp.gp,
                                                  bp.counters, bp.cnt, bp.blk, nullptr, nullptr, 0, bp.sweep_tmp, bp.sweep_tot);
    OB_CHECK_KERNEL("k_sweep", st);
    scan_exclusive(bp.blk, bp.blk, (long)PC_COUNT * nb2 + 1, nullptr, nullptr, bp.scan, st);
    k_sweep<true><<<nb2, SWEEP_THREADS, 0, st>>>(n, bp.keys, bp.s_min, bp.s_max, bp.s_flt, bp.cell_start, bp.cell_end, bp.gp,
                                                 bp.counters, bp.cnt, nullptr, bp.blk, bp.pairs, bp.cap_pairs, bp.sweep_tmp,
                                                 bp.sweep_tot);
    OB_CHECK_KERNEL("k_sweep", st);
    k_pairs_finish<<<1, 1, 0, st>>>((int)nb2, bp.blk, bp.cap_pairs, bp.counters, d_stats, bp.gp);
    OB_CHECK_KERNEL("k_pairs_finish", st);
    k_cell_clear<<<nb, 256, 0, st>>>(n, bp.keys, bp.cap_cells, bp.cell_end);
    OB_CHECK_KERNEL("k_cell_clear", st);
}

} // namespace ob
