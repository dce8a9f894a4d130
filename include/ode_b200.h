/* ode_b200.h -- libode_b200 extensions to the ODE C API (C ABI: plain pointers and sizes only).
 *
 * Why they exist: the reference drives collision through a per-pair host callback
 * (dSpaceCollide -> NearCallback -> dJointCreateContact, /root/reference/src/main.c:212,674-693).
 * That shape is kept bit-for-bit in <ode/ode.h> ("compat mode"), but at 10^6 bodies it would mean
 * ~10^7 host calls per tick.  The reference's callback is a fixed policy -- uniform surface, up to
 * 8 contacts, attach the geoms' bodies -- so the device-resident mode below declares that policy
 * once and pairs, contacts and rows never leave the GPU.
 */
#ifndef ODE_B200_EXT_H
#define ODE_B200_EXT_H

#include <stddef.h>

#include "ode/ode.h"

#ifdef __cplusplus
extern "C" {
#endif

/* CUDA device used by worlds created afterwards (default: env ODE_B200_DEVICE, LOCAL_RANK, or 0) */
void dSetDeviceB200(int device);
int dGetDeviceB200(void);
int dWorldGetDeviceB200(dWorldID);
/* page-locked host memory for the buffers of dWorldSetForcesB200 / dWorldGetSnapshotB200 & co, so that a plain-C host needs no
 * CUDA headers to get asynchronous copies at full PCIe rate.  write_combined != 0: for buffers the host only WRITES (force
 * uploads) -- the GPU reads them without snooping the CPU caches; reading such memory from the CPU is slow. */
void *dAllocPinnedB200(size_t bytes, int write_combined);
void dFreePinnedB200(void *); /* the CUDA device this world lives on (fixed at dWorldCreate) */

/* device-resident replacement of `dSpaceCollide(space, 0, NearCallback)` (src/main.c:212): runs
 * broadphase + narrowphase and keeps the contacts on the GPU for the next dWorldQuickStep/dWorldStep,
 * which applies the world's uniform surface to every contact.  Asynchronous. */
void dSpaceCollideDeviceB200(dSpaceID, int max_contacts);
/* the uniform surface policy; default = the reference's NearCallback (src/main.c:684-687):
 * mode dContactBounce, bounce 0.2, bounce_vel 0.1, mu dInfinity */
void dWorldSetSurfaceB200(dWorldID, const dSurfaceParameters *);
void dWorldGetSurfaceB200(dWorldID, dSurfaceParameters *);
/* max contacts per pair precomputed by dSpaceCollide in compat mode (default 8, src/main.c:675) */
void dWorldSetMaxContactsB200(dWorldID, int max_contacts);

/* batched independent worlds ("envs") inside one dWorldID: geoms of different envs never collide;
 * env -1 on a static geom means "present in every env". */
void dWorldSetNumEnvsB200(dWorldID, int n_envs);
/* slot re-use (default off): with on != 0, dBodyDestroy / dGeomDestroy remember the freed index and the next dBodyCreate /
 * dCreateSphere / dCreateBox takes the lowest remembered one instead of growing the world's arrays -- for applications that
 * spawn and destroy for ever (the reference's server: src/main.c:695-733 re-uses its own 512 slots the same way; the slab
 * driver's migration).  Off, indices are handed out in creation order, which the parity tests rely on. */
void dWorldSetSlotReuseB200(dWorldID, int on);
void dBodySetEnvB200(dBodyID, int env);
void dGeomSetEnvB200(dGeomID, int env);

/* bulk creation (one call instead of ~10 per body). Any pointer but pos may be NULL for defaults
 * (q identity, zero velocity, mass 1 / identity inertia, flags 0, env 0). inertia9 is row-major.
 * flags: 1 kinematic, 2 no gravity, 4 gyroscopic. Returns the index of the first new body. */
int dWorldAddBodiesB200(dWorldID, int n, const float *pos3, const float *quat4, const float *lvel3,
                        const float *avel3, const float *mass, const float *inertia9,
                        const int *flags, const int *env);
/* type: dSphereClass/dBoxClass/dPlaneClass/dTriMeshClass; dims4: r | lx,ly,lz | a,b,c,d | mesh id;
 * body: body index within the world or -1 (then pos3/R12 give the static pose; NULL = identity).
 * Returns the index of the first new geom. */
int dSpaceAddGeomsB200(dSpaceID, dWorldID, int n, const int *type, const float *dims4,
                       const int *body, const float *pos3, const float *R12,
                       const unsigned *category, const unsigned *collide, const int *env);
int dWorldAddTriMeshB200(dWorldID, const float *verts3, int n_verts, const int *tris3, int n_tris);
dBodyID dWorldGetBodyB200(dWorldID, int index);
dGeomID dSpaceGetGeomB200(dSpaceID, int index);
int dWorldGetNumBodiesB200(dWorldID);
int dBodyGetIndexB200(dBodyID);
int dGeomGetIndexB200(dGeomID);

/* bulk state access (blocking). Any output pointer may be NULL. */
void dWorldGetStateB200(dWorldID, float *pos3, float *quat4, float *lvel3, float *avel3, float *R12);
/* per-step external force + torque, 6 floats per body (what dBodyAddForce/dBodyAddTorque set) */
void dWorldSetForcesB200(dWorldID, const float *force_torque6, int n);

/* fused snapshot: 16 floats per body in the layout of the reference's GetTransformMat
 * (src/main.c:602-622), written by the solver-tail kernel. Copies bodies [first, first+count). */
void dWorldGetSnapshotB200(dWorldID, float *dst16, int first, int count, int blocking);
const float *dWorldGetSnapshotDeviceB200(dWorldID);
/* Snapshot record format.  0 (default): the 16 floats of GetTransformMat.  1: the same without its four constant
 * floats -- columns 0..2 of the 4x4 (3 floats each), then the translation: 12 floats.  2: position (x y z 1) and the
 * quaternion (w x y z): 8 floats.  dWorldGetSnapshotB200 then delivers count * {16, 12, 8} floats: the formats cut
 * the per-tick device-to-host traffic to 3/4 and 1/2 (at 8 GPUs the host's PCIe complex, not the GPUs, bounds the
 * end-to-end tick rate).  dSnapshotExpandB200 turns `count` compact records into the 16-float layout on the host,
 * with `threads` worker threads, bit-identical to format 0 (format 2 re-evaluates dQtoR in the solver tail's own
 * arithmetic).  Bodies spawned or moved since the last step are included with their current pose. */
void dWorldSetSnapshotFormatB200(dWorldID, int format);
int dWorldGetSnapshotFormatB200(dWorldID);
void dSnapshotExpandB200(const float *compact, int format, int count, float *dst16, int threads);
void dWorldWaitB200(dWorldID);
/* halo exchange of a slab-decomposed world: gather the states of the listed
 * bodies into a DEVICE buffer / scatter received states into the listed (ghost) bodies.  16 floats per
 * body: pos3 pad, quat4, lvel3 pad, avel3 pad.  d_idx and the buffers are device pointers (so NCCL can
 * send them GPU to GPU); asynchronous on the world's stream. */
void dWorldPackStatesDeviceB200(dWorldID, const int *d_idx, int n, float *d_out16);
void dWorldUnpackStatesDeviceB200(dWorldID, const int *d_idx, int n, const float *d_in16);
/* impulse half of the halo exchange: the velocity change the last step's contacts
 * gave each listed body -- 8 floats per body: h*fc linear xyz, pad, angular xyz, pad -- gathered into a DEVICE
 * buffer; the receiving world adds such a buffer to the listed bodies' velocities (kinematic bodies are
 * skipped).  A slab that solves a contact against a dynamic ghost sends the ghost's impulse to its owner.
 * Batched worlds keep their accumulators on chip unless dWorldSetKeepImpulsesB200(world, 1). */
void dWorldPackImpulsesDeviceB200(dWorldID, const int *d_idx, int n, float *d_out8);
void dWorldAddImpulsesDeviceB200(dWorldID, const int *d_idx, int n, const float *d_in8);
void dWorldSetKeepImpulsesB200(dWorldID, int on);
/* dynamic halo (bodies move, so the set near a slab face is rebuilt every tick, on the device):
 *  - Select: indices of the bodies with mask[b] != 0 (mask may be NULL) and lo <= pos[axis] < hi, ascending, into
 *    d_idx_out[0..cap); unused entries are -1; *d_count = number selected (may exceed cap: then truncated).
 *  - PackBodies: 48 floats per list entry -- pos3 invM | quat4 | lvel3 mass | avel3 geom-type | geom dims4 |
 *    I (3x4) | invI (3x4) | body flags, pad3 -- i.e. everything the other side needs to simulate the body;
 *    d_body_geom maps a body to its geom; entries with index -1 are marked empty.
 *  - UnpackBodies: slot i of the message goes into body d_ghost_body[i] / geom d_ghost_geom[i] (a pool of
 *    cap ghost slots created once); empty entries switch the slot off (geom not alive, body inert).
 * The impulse calls above skip -1 entries, so the same list drives the return path. */
void dWorldSelectBodiesDeviceB200(dWorldID, int axis, float lo, float hi, const int *d_mask, int *d_idx_out, int cap,
                                  int *d_count);
void dWorldPackBodiesDeviceB200(dWorldID, const int *d_idx, int cap, const int *d_body_geom, float *d_out48);
void dWorldUnpackBodiesDeviceB200(dWorldID, const int *d_ghost_body, const int *d_ghost_geom, int cap, const float *d_in48);
/* the CUDA stream (cudaStream_t) every kernel and copy of this world is queued on */
void *dWorldGetStreamB200(dWorldID);
/* debugging aid: with ODE_B200_DEBUG_GUARD=1 in the environment every device allocation of the library carries a
 * 256-byte guard band on both sides; this call synchronises, reads the bands back and returns how many allocations a
 * kernel wrote outside of (naming them on stderr when verbose != 0).  Returns -1 when the guards are off. */
int dCheckGuardsB200(int verbose);
/* self-test of that mechanism: allocates a guarded scratch buffer of 1000 bytes, writes `overrun` bytes past its end
 * (0 = stays inside), runs the check, frees the buffer; returns what dCheckGuardsB200 returned in between. */
int dGuardSelfTestB200(int overrun);

/* Slab decomposition of ONE large world over several GPUs (BASELINE config 5), driven from C:
 * one process per GPU, each owning the bodies whose centre lies in its x-interval [face_left, face_right); a contact
 * across a face belongs to the LOWER slab, which mirrors its upper neighbour's boundary bodies (x < face + margin)
 * as dynamic ghosts in a pool of `pool` body/geom slots created by the application.  dSlabTickB200 = state halo down
 * -> dSpaceCollideDeviceB200 + dWorldQuickStep -> impulse halo up, NCCL send/recv inside the library, ordered by CUDA
 * events only (no host synchronisation).  dSlabMigrateB200 (every few ticks) moves the ownership of bodies that
 * crossed a face by more than `hyst`.  The communicator is built from an NCCL unique id the application distributes
 * (rank 0: dSlabGetUniqueIdB200, then MPI / sockets / a file).  Several slabs in one process (one GPU, tests) are
 * connected with dSlabConnectLocalB200 and ticked together with dSlabTickLocalB200: same phases, device copies. */
typedef struct dxSlabB200 *dSlabID;
typedef struct dSlabLayoutB200 {
    float face_left, face_right, margin, hyst;
    int n_own;            /* bodies [0, n_own) are owned at set-up */
    int n_static;         /* static geoms precede the body geoms: geom of body b = n_static + b */
    int pool, pool_first_body, pool_first_geom; /* ghost slots (ignored on the last rank) */
    int mig_cap;          /* bodies that may change owner per face per migration */
} dSlabLayoutB200;
typedef struct dSlabInfoB200 {
    int n_owned, halo_selected, halo_overflow, mig_overflow;
    long migrated_in, migrated_out, ticks, halo_bytes_per_tick;
} dSlabInfoB200;
int dSlabGetUniqueIdB200(char id128[128]); /* 0: NCCL not available */
dSlabID dSlabCreateB200(dWorldID, dSpaceID, int rank, int n_ranks, const char *nccl_id128 /* NULL: local transport */,
                        const dSlabLayoutB200 *);
void dSlabDestroyB200(dSlabID);
void dSlabTickB200(dSlabID, dReal h, int max_contacts);
void dSlabMigrateB200(dSlabID);
void dSlabConnectLocalB200(dSlabID lower, dSlabID upper);
void dSlabTickLocalB200(dSlabID *slabs, int n, dReal h, int max_contacts);
void dSlabMigrateLocalB200(dSlabID *slabs, int n);
void dSlabGetInfoB200(dSlabID, dSlabInfoB200 *); /* blocking */

/* CUDA-event timer on the world's stream: everything queued between start and stop */
void dWorldTimerStartB200(dWorldID);
void dWorldTimerStopB200(dWorldID);
float dWorldTimerElapsedB200(dWorldID); /* ms, blocking */
/* several worlds of one device ticking side by side (each world has its own stream, so their ticks overlap on the GPU):
 * ms from start_world's start event to stop_world's stop event, blocking */
float dWorldTimerElapsedBetweenB200(dWorldID start_world, dWorldID stop_world);
/* number of CUDA kernels this library has launched in this process so far */
long dGetKernelLaunchCountB200(void);

/* Wire image of the reference's MsgUpdateBodies (inc/msgs.h:30-33; what src/main.c:221-242 assembles with a
 * 512-iteration host loop + memcpy): bind the application's slot table once -- per slot the dBodyID (or
 * NULL), the dGeomID (static slots), BodyType, size[3] and RGBA colour, i.e. the fields of Body/BodyState
 * (inc/body.h:20-31) -- then, after any step, one kernel writes {int msg; BodyState bodies[n_slots]} (84 bytes
 * per slot) and one copy brings it to `dst`.  Returns the image size in bytes (4 + 84 * n_slots). */
void dWorldBindSnapshotSlotsB200(dWorldID, int n_slots, const dBodyID *bodies, const dGeomID *geoms, const int *types,
                                 const float *size3, const unsigned *rgba);
size_t dWorldPackMsgUpdateBodiesB200(dWorldID, void *dst, int msg_type, int blocking);

/* capacities (pairs, manifolds); 0 = automatic (8 and 6 per geom). Overflow sets a stats flag. */
void dWorldSetCapacityB200(dWorldID, long max_pairs, long max_manifolds);
/* solver selection: mode 0 = automatic (island solver for batched worlds, grid-barrier solver
 * otherwise), 1 = always the grid-barrier solver. env_group = lanes per env of the island solver
 * (8, 16, 32; 0 = automatic). Both solvers produce bit-identical results. */
void dWorldSetSolverModeB200(dWorldID, int mode, int env_group);
/* Gauss-Seidel unit of the device-resident solver: 0 = one manifold (all contacts of a geom pair),
 * 1 = one contact, -1 = automatic (per contact for batched worlds, per manifold otherwise).  Either is a
 * valid row order; results differ in the last bits between the two settings. */
void dWorldSetContactUnitsB200(dWorldID, int per_contact);
/* dynamic geoms whose AABB extent exceeds this are treated like static "big" geoms (default inf) */
void dWorldSetBigExtentB200(dWorldID, float extent);
/* Wavefront OBJ (v / f records, 1-based or negative indices, polygons fan-triangulated) into a trimesh data
 * object, in place of dGeomTriMeshDataBuildSingle (the reference ships res/teapot.obj and res/grassPlane.obj;
 * BASELINE config 2 uses the teapot as a static collision mesh).  Returns the triangle count or -1. */
int dGeomTriMeshDataBuildFromOBJB200(dTriMeshDataID, const char *path);
/* read a trimesh data object back: copies up to cap_verts vertices (3 floats each) and cap_tris triangles (3 ints
 * each); returns the triangle count and stores the vertex count in *n_verts (either buffer may be NULL) */
int dGeomTriMeshDataGetB200(dTriMeshDataID, float *verts3, int cap_verts, int *tris3, int cap_tris, int *n_verts);
/* What dWorldStep runs (the reference calls dWorldStep, src/main.c:213; libode's dWorldStep solves the step's LCP
 * exactly, island by island, with a Dantzig solver).
 *   max_iters == 0 (default): the exact solution too -- islands on the device, A = J M^-1 J^T + cfm/h per island,
 *     block principal pivoting with a dense double-precision Cholesky, one CTA per island -- for worlds of at most
 *     1024 bodies / 4096 contact manifolds whose islands have at most 384 rows (the reference's scene: 68 bodies);
 *     larger worlds, larger islands and dContactApprox1 contacts fall back to dWorldQuickStep's sweeps inside the same
 *     call, and dStepStatsB200.exact_status says so.
 *   max_iters  > 0: up to max_iters SOR/PGS sweeps, stopping once the largest |delta lambda| of a sweep is below tol
 *     (tol 0: always max_iters) -- converges to the same solution, slowly.
 *   max_iters  < 0: dWorldStep == dWorldQuickStep.
 * dWorldQuickStep is not affected. */
void dWorldSetStepSolverB200(dWorldID, int max_iters, float tol);
/* broadphase layout: -1 automatic (default), 0 uniform grid (sort + cell sweep), 1 all pairs per env (batched
 * worlds whose geoms were added env by env, or one world of <= 4096 geoms; falls back to the grid
 * otherwise).  Both emit the same pair SET; the order inside the list differs. */
void dWorldSetBroadphaseB200(dWorldID, int mode);

typedef struct dStepStatsB200 {
    int n_geoms, n_big, n_pairs, n_contacts, n_manifolds, n_colours, n_overflow, flags;
    int class_count[7]; /* sphere-sphere, sphere-box, box-box, sphere-plane, box-plane, sphere-trimesh, none */
    int n_rows, n_rows1, n_rows2;
    int colour_rounds;
    float cell_size;
    int grid_dims[3];
    int solver_iters; /* sweeps the last solve ran (fewer than the limit when residual-terminated; 0: exact solve) */
    /* dWorldStep's exact solve: -1 not attempted (dWorldQuickStep, or sweeps requested), 0 the step's LCP was solved
     * exactly, 1 world or island too large, 2 dContactApprox1 rows, 3 no convergence -- 1..3 fell back to the sweeps */
    int exact_status, n_islands, max_island_rows, pivot_rounds;
    /* island solver of batched worlds: 32-lane trips and occupied lanes of one sweep, summed over the worlds */
    int env_trips, env_lanes;
} dStepStatsB200;
void dWorldGetStatsB200(dWorldID, dStepStatsB200 *); /* blocking */
void dWorldEnableTimingB200(dWorldID, int on);
/* CUDA-event times of the last tick, ms: collide, prepare (manifolds+colouring+rows), solve (PGS
 * iterations + integrate + pack), whole tick */
void dWorldGetTimingsB200(dWorldID, float out_ms[4]);
/* the same tick split further: broadphase, narrowphase, prepare, solve, whole tick */
void dWorldGetStageTimingsB200(dWorldID, float out_ms[5]);

/* parity-test hooks: the broadphase pair list and the contacts of the last collide (blocking).
 * pairs2: (g1,g2) per pair in device order; returns the number of pairs (may exceed cap). */
int dSpaceGetPairsB200(dSpaceID, int *pairs2, int cap);
/* per pair: count; contacts compacted pair-major: pos3+depth (4), normal3 + side (4, side as int bits) */
int dSpaceGetContactsB200(dSpaceID, int *count_per_pair, int cap_pairs, float *pos_depth4,
                          float *normal_side4, int cap_contacts);
/* the solver's Gauss-Seidel order of the last device-resident step: per contact (g1, g2, k) */
int dWorldGetSolverOrderB200(dWorldID, int *g1, int *g2, int *k, int cap);

#ifdef __cplusplus
}
#endif

#endif
