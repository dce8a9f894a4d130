// dev.cuh -- device-side array bundles shared by the kernels of libode_b200 (see DESIGN.md for
// the HBM layout).  All per-element arrays are struct-of-arrays of float4 / int so warps read and
// write full 16-byte lanes; rotations are kept as 3 consecutive float4 rows per element (48 B, one
// and a half sectors) because they are gathered by index, not streamed.
#pragma once

#include "dmath.cuh"
#include "engine.h"
#include "prims.cuh"

namespace ob {

struct BodyArrays {
    int n;
    float4 *pos;   // xyz, invMass
    float4 *quat;  // w x y z
    float4 *R;     // 3 per body
    float4 *lvel;  // xyz, mass
    float4 *avel;  // xyz, -
    float4 *I;     // 3 per body, body frame
    float4 *invI;  // 3 per body, body frame
    float4 *facc, *tacc;
    int *flags;
    int *local;    // body index relative to the first body of its env (env-invariant colouring priority)
    int *env;      // env (independent world) of the body
    // per-step solver views
    float4 *inv;   // 3 per body: rows of the world-frame inverse inertia; row 0 .w = invMass
    float4 *tmp;   // 2 per body: v/h + invM*f, w/h + invI*t
    float4 *fc;    // 2 per body: constraint acceleration accumulators (lin, ang)
    float *snap;   // snapshot records, 16 / 12 / 8 floats each by snap_fmt: GetTransformMat layout or its compact forms
    int snap_fmt;  // 0: 16 floats; 1: the 12 non-constant ones; 2: pos + quaternion (8 floats)
    // colouring scratch
    unsigned long long *colmask;
    unsigned long long *prio;
};

struct GeomArrays {
    int n;
    int *type;
    float4 *dims;
    int *body;
    float4 *pos;
    float4 *R;     // 3 per geom
    uint32_t *cat, *col;
    int *env;
    int *mesh;
    int *alive;
    float4 *amin, *amax;
};

struct MeshInfo {
    const float *verts; // 3*nv, 16-byte padded
    const int *tris;    // 3*nt
    int nv, nt;
    float lo[3], hi[3];
    // uniform grid over the mesh-local bounds, built once at dGeomTriMeshDataBuild*: triangle t is listed in every
    // cell its box touches; cell (x, y, z) holds cell_tris[cell_start[c] .. cell_start[c + 1]), c = (z * gd[1] + y) * gd[0] + x
    const int *cell_start;
    const int *cell_tris;
    int gd[3];
    float gcell[3], ginv[3];
};
constexpr int MAX_MESHES = 8;
struct MeshTable {
    MeshInfo m[MAX_MESHES];
    int n;
};

struct GridParams {
    float ox, oy, oz, cell, inv_cell, small_extent;
    int dx, dy, dz, per_env, n_envs;
};

struct BroadCounters {
    int first_big, first_dead, n_pairs;
    int n_bb; // box-box pairs that passed the separating-axis pass (two-pass narrowphase)
    int class_start[PC_COUNT + 1];
};

constexpr int SWEEP_TCAP = 16; // hits a sweep thread can park in the count pass (broadphase.cu)

struct BroadPhase {
    int cap_geoms = 0, cap_cells = 0, key_bits = 0, cap_pairs = 0;
    long cap_blk = 0; // ints in blk
    unsigned *acc = nullptr;
    GridParams *gp = nullptr;
    BroadCounters *counters = nullptr;
    uint32_t *keys = nullptr;
    int *idx = nullptr;
    float4 *s_min = nullptr, *s_max = nullptr;
    uint4 *s_flt = nullptr;
    int *cell_start = nullptr, *cell_end = nullptr;
    int *cnt = nullptr; // PC_COUNT * n: pairs per sweep thread and class
    int *blk = nullptr; // PC_COUNT * (sweep blocks) + 1: per-block class sums, scanned in place
    int2 *sweep_tmp = nullptr; // SWEEP_TCAP layers of n parked hits
    int *sweep_tot = nullptr;  // hits per sweep thread
    int2 *pairs = nullptr;
    int *bb_list = nullptr; // box-box pairs without a separating axis (compacted by k_np_box_box_sat)
    SortWorkspace sort;
    ScanWorkspace scan;
};

// all-pairs-per-env broadphase (batched worlds, or one small world): no sort, no grid.  Geoms of env e are
// the index range [first[e], first[e] + count[e]); geoms shared by every env are listed in `shared`.
struct EnvBroad {
    int enabled = 0;
    int single = 0;      // one world: every geom belongs to the one range, env ids are ignored
    int n_shared = 0, n_alive = 0;
    int sap = 1; // sort-and-sweep CTA per env when every env has <= 128 geoms (0: always the all-pairs sweep)
    int n_envs = 0, max_count = 0; // number of env ranges and the largest one (<= 128: sort-and-sweep CTA per env)
    int *first = nullptr, *count = nullptr, *shared = nullptr;
};

void broadphase_run(BroadPhase &bp, GeomArrays g, const float4 *b_pos, const float4 *b_R, MeshTable meshes,
                    int n_envs, float big_extent, const EnvBroad &eb, StepStats *d_stats, cudaStream_t st);

void broadphase_acc_init(BroadPhase &bp, cudaStream_t st);
void broadphase_single_pair(BroadPhase &bp, GeomArrays g, const float4 *b_pos, const float4 *b_R, MeshTable meshes, float big_extent,
                            int g1, int g2, StepStats *d_stats, cudaStream_t st);

// contact slot storage: contact k of pair p lives at [k * stride + p]
struct ContactSlots {
    float4 *pd;   // pos.xyz, depth
    float4 *ns;   // normal.xyz, bitcast(int side2 / triangle index)
    int *nc;      // contacts per pair
    int stride;   // = pair capacity
};

void narrowphase_run(const BroadPhase &bp, GeomArrays g, MeshTable meshes, const std::vector<TriMesh> &host_meshes,
                     ContactSlots cs, int max_contacts, StepStats *d_stats, int num_sms, cudaStream_t st);

// manifold = all contacts of one geom pair (device mode) or one run of contact joints with the
// same body pair (compat mode); one solver thread owns a manifold.
struct ManifoldArrays {
    int cap;
    int4 *rec;        // b1, b2, cbase, nc | reverse << 8
    int *colour;      // per manifold
    uint32_t *skey;   // sort key (colour, nc)
    int *sidx;        // sorted -> manifold
    int *flag;        // compaction flags / scan
    int *count;       // device: number of manifolds
    int *colour_start; // [66]
    int *meta;        // [0] n_colours, [1] n_overflow, [2],[3] remaining (alternating), [4] rounds, [5] scan total
};

// batched independent worlds: manifolds bucketed per env for the island solver (one lane group per env)
struct EnvArrays {
    int n_envs;
    int cap;         // capacity of rec / perm / col (solver units); more units than this: flagged, the step runs without contacts
    int max_bodies;  // largest number of bodies in one env (host-known)
    int contiguous;  // 1: every env's bodies are one contiguous index range (enables shared-memory staging)
    int *first_body; // first body of each env
    int *n_body;     // bodies of each env
    int *cnt;        // manifolds per env (n_envs + 1 for the scan)
    int *start;      // exclusive scan of cnt, [n_envs] = total
    int *fill;       // bucket cursors
    int *order;      // envs by decreasing unit count (the island solver's work queue)
    int4 *rec;       // manifold records bucketed by env
    int *perm;       // per env: bucket slots ordered by colour
    unsigned char *col; // colour per bucket slot
};

// solver rows, k-major: contact k of sorted manifold s at [k * cap + s]
struct SolverArrays {
    int cap;
    float4 *q0; // n.xyz, rhsN
    float4 *q1; // r1.xyz, AdN
    float4 *q2; // r2.xyz, AdcfmN
    float4 *q3; // rhsT1, AdT1, AdcfmT1, k = 1/sqrt(.) of dPlaneSpace(n)
    float4 *q4; // rhsT2, AdT2, AdcfmT2, mu
    float4 *q5; // (mu2, -, -, -): only written/read for dContactMu2 contacts
    float4 *lam; // lambdaN, lambdaT1, lambdaT2, bitcast(rows | findex flags)
    int4 *mrec;  // per sorted manifold: b1, b2, nc, manifold id
};

struct StepConfig {
    float h, erp, cfm, sor_w, max_vel, min_depth;
    float gx, gy, gz;
    int iters;
    float tol; // > 0: stop the sweeps when the largest |delta lambda| of a sweep falls below it (global solver)
};

void solver_step(Engine *e, float h, bool host_contacts, const Surface *uniform_surface);
void solver_profile_dump();
float solver_barrier_bench(Engine *e, int iters);

} // namespace ob
