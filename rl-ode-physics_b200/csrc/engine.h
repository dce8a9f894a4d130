// engine.h -- device-resident rigid-body world behind libode_b200.so (sm_100a only).
//
// Replaces, as hand-written CUDA kernels, the libode internals the reference reaches from
// /root/reference/src/main.c:212-214 (dSpaceCollide -> NearCallback/dCollide -> dWorldStep):
// broadphase, narrowphase, contact rows, SOR/PGS solve, integration and the snapshot pack of
// src/main.c:218-243.  Layout and kernel inventory: DESIGN.md.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <initializer_list>
#include <vector>

namespace ob {

// ---------------------------------------------------------------------------------------------
// error handling: no CPU fallback anywhere -- a CUDA failure aborts with a message (ODE's dError
// behaviour, SURVEY.md section 8b "Errors").
#define OB_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            fprintf(stderr, "libode_b200: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e__), \
                    __FILE__, __LINE__, cudaGetErrorString(e__));                              \
            abort();                                                                           \
        }                                                                                      \
    } while (0)

// ODE_B200_DEBUG_SYNC=1: synchronise after every kernel and name the one that faulted
inline bool ob_debug_sync() {
    static int v = -1;
    if (v < 0) {
        const char *s = getenv("ODE_B200_DEBUG_SYNC");
        v = (s && s[0] == '1') ? 1 : 0;
    }
    return v == 1;
}
// every kernel launch of the library passes through OB_CHECK_KERNEL (or ob_count_launch for the
// cooperative ones): the counter is what bench.py reports as gpu_launches
extern long g_ob_launches;
inline void ob_count_launch() { ++g_ob_launches; }
#define OB_CHECK_KERNEL(name, st)                                                                  \
    do {                                                                                           \
        ob_count_launch();                                                                         \
        cudaError_t e__ = cudaGetLastError();                                                      \
        if (e__ == cudaSuccess && ob_debug_sync()) e__ = cudaStreamSynchronize(st);                \
        if (e__ != cudaSuccess) {                                                                  \
            fprintf(stderr, "libode_b200: kernel %s failed: %s (%s:%d)\n", name,                  \
                    cudaGetErrorString(e__), __FILE__, __LINE__);                                  \
            abort();                                                                               \
        }                                                                                          \
    } while (0)

// Device memory of the library goes through ob_malloc / ob_free.  With ODE_B200_DEBUG_GUARD=1 every allocation gets a
// 256-byte guard band on both sides, filled with a pattern; dCheckGuardsB200() (ode_b200.h) reads the bands back and
// reports every allocation a kernel wrote outside of (the pool's compute-sanitizer is closed, so this is the library's
// own overrun check; tests/test_guards_gpu.py runs the scenes of the parity tests under it).
cudaError_t guard_malloc(void **p, size_t bytes, const char *file, int line);
cudaError_t guard_free(void *p);
int guard_check(int verbose); // number of allocations with a damaged band; -1 = guards are off
template <typename T>
inline cudaError_t ob_malloc_at(T **p, size_t bytes, const char *file, int line) {
    return guard_malloc(reinterpret_cast<void **>(p), bytes, file, line);
}
#define ob_malloc(p, bytes) ob::ob_malloc_at(p, bytes, __FILE__, __LINE__)
#define ob_free(p) ob::guard_free(p)

enum GeomType { G_SPHERE = 0, G_BOX = 1, G_PLANE = 4, G_TRIMESH = 8 };
enum BodyFlags { BF_KINEMATIC = 1, BF_NOGRAVITY = 2, BF_GYRO = 4 };

// pair classes, in the order the pair list is grouped (narrowphase launches one kernel per class)
enum PairClass {
    PC_SPHERE_SPHERE = 0,
    PC_SPHERE_BOX = 1,
    PC_BOX_BOX = 2,
    PC_SPHERE_PLANE = 3,
    PC_BOX_PLANE = 4,
    PC_SPHERE_TRIMESH = 5,
    PC_NONE = 6, // AABBs overlap but no collider exists (plane-plane, box-trimesh, ...)
    PC_COUNT = 7
};

// surface policy of a contact (dSurfaceParameters subset); mode bits are ODE's dContact* values
struct Surface {
    int mode;
    float mu, mu2, bounce, bounce_vel, soft_erp, soft_cfm;
    float motion1, motion2, motionN, slip1, slip2;
};

// host-side contact joint record (compat mode: dJointCreateContact + dJointAttach)
struct HostContact {
    float pos[3], depth;
    float normal[3];
    int b1, b2; // engine body indices, -1 = NULL
    Surface surf;
};

struct TriMesh {
    int nv = 0, nt = 0;
    float *d_verts = nullptr; // 3*nv floats, padded to 16 B
    int *d_tris = nullptr;    // 3*nt ints
    float lo[3], hi[3];       // local bounds
    std::vector<float> h_verts;
    std::vector<int> h_tris;
    int *d_cell_start = nullptr, *d_cell_tris = nullptr; // triangle grid (device)
};

// per-step counters read back once per step (pinned)
struct StepStats {
    int n_geoms, n_big, n_pairs, n_contacts, n_manifolds, n_colours, n_overflow, flags;
    int class_count[PC_COUNT];
    int n_rows, n_rows1, n_rows2; // total rows, rows of one-body / two-body manifolds
    int colour_rounds;
    float cell_size;
    int grid_dims[3];
    int solver_iters; // sweeps the last solve ran (< iterations when residual-terminated; 0: exact solve)
    // exact dWorldStep (solver_exact.cu): -1 not attempted, 0 solved exactly, 1 world/island too large, 2 rows it does not
    // take (dContactApprox1), 3 pivoting did not converge -- 1..3: the step fell back to the sweeps
    int exact_status, n_islands, max_island_rows, pivot_rounds;
    // lane-pair island solver: 32-lane trips and occupied lanes of ONE sweep, summed over the envs (lane fill = lanes / (32 trips))
    int env_trips, env_lanes;
};
enum StatFlags { SF_PAIR_OVERFLOW = 1, SF_MANIFOLD_OVERFLOW = 2, SF_CAND_OVERFLOW = 4 };

struct WorldParams {
    float gravity[3] = {0, 0, 0};
    float erp = 0.2f, cfm = 1e-5f, sor_w = 1.3f;
    int iters = 20;
    float tol = 0.f; // > 0: residual-terminated sweeps (dWorldStep parity mode)
    int exact = 0;   // this step: solve the LCP exactly if the world is small enough (dWorldStep)
    float max_vel = INFINITY, min_depth = 0.0f;
};

struct Engine; // opaque to the shim

// lifecycle
Engine *eng_create(int device);
void eng_destroy(Engine *);
WorldParams &eng_params(Engine *);
int eng_device(Engine *);
cudaStream_t eng_stream(Engine *);

// host mirrors (the shim reads/writes these, then marks dirty)
// Growable host array whose elements NEVER move: dBodyGetPosition / Rotation / Quaternion / LinearVel / AngularVel and
// dGeomGetPosition / Rotation hand out pointers into these mirrors, and libode guarantees such a pointer for the body's
// lifetime (a later dBodyCreate must not invalidate it).  The array reserves address space once (mmap, PROT_NONE, no memory
// committed) and commits pages as it grows, so data() is contiguous -- the bulk uploads / read-backs copy it in one piece.
template <typename T>
class StableVec {
  public:
    StableVec() = default;
    StableVec(const StableVec &) = delete;
    StableVec &operator=(const StableVec &) = delete;
    ~StableVec();
    size_t size() const { return n_; }
    bool empty() const { return n_ == 0; }
    T *data() { return p_; }
    const T *data() const { return p_; }
    T &operator[](size_t i) { return p_[i]; }
    const T &operator[](size_t i) const { return p_[i]; }
    T *begin() { return p_; }
    T *end() { return p_ + n_; }
    void resize(size_t n, T v = T());
    // append only (pos must be end()): the two forms the engine uses
    void insert(T *pos, std::initializer_list<T> il) { append(pos, il.begin(), il.size()); }
    void insert(T *pos, const T *first, const T *last) { append(pos, first, (size_t)(last - first)); }

  private:
    void append(T *pos, const T *src, size_t k);
    void grow(size_t n); // make room for n elements
    T *p_ = nullptr;
    size_t n_ = 0, committed_ = 0, reserved_ = 0; // elements in use; BYTES committed; BYTES of address space reserved
};
void *stable_reserve(size_t *bytes);                     // engine.cu; may reserve less than asked (address-space limits)
void stable_commit(void *base, size_t old_bytes, size_t new_bytes);
void stable_release(void *base, size_t bytes);
// address space per array: 3 GiB = the rotation matrices (48 B) of 67 M bodies; 11 GiB per world with all seven mirrors, so
// a process may hold thousands of worlds (the 47-bit address space has room for ~11 000)
constexpr size_t STABLE_RESERVE = (size_t)3 << 30;
template <typename T>
StableVec<T>::~StableVec() {
    if (p_) stable_release(p_, reserved_);
}
template <typename T>
void StableVec<T>::grow(size_t n) {
    const size_t need = n * sizeof(T);
    if (need <= committed_) return;
    if (!p_) {
        reserved_ = STABLE_RESERVE;
        p_ = static_cast<T *>(stable_reserve(&reserved_));
    }
    if (need > reserved_) {
        fprintf(stderr, "libode_b200: a host mirror outgrew its %zu-byte address reservation\n", reserved_);
        abort();
    }
    size_t want = committed_ ? committed_ * 2 : ((size_t)1 << 16);
    while (want < need) want *= 2;
    if (want > reserved_) want = reserved_;
    stable_commit(p_, committed_, want);
    committed_ = want;
}
template <typename T>
void StableVec<T>::resize(size_t n, T v) {
    grow(n);
    for (size_t i = n_; i < n; i++) p_[i] = v;
    n_ = n;
}
template <typename T>
void StableVec<T>::append(T *pos, const T *src, size_t k) {
    if (pos != end()) {
        fprintf(stderr, "libode_b200: StableVec supports appending only\n");
        abort();
    }
    grow(n_ + k);
    for (size_t i = 0; i < k; i++) p_[n_ + i] = src[i];
    n_ += k;
}

struct HostBodies {
    StableVec<float> pos;   // 4 per body: x y z invMass      (the five arrays the dBodyGet* pointers point into)
    StableVec<float> quat;  // 4: w x y z
    StableVec<float> R;     // 12 row-major 3x4
    StableVec<float> lvel;  // 4: xyz, mass
    StableVec<float> avel;  // 4: xyz, pad
    std::vector<float> I;     // 12 body-frame inertia (3x4)
    std::vector<float> invI;  // 12 body-frame inverse inertia (3x4)
    std::vector<float> facc;  // 4
    std::vector<float> tacc;  // 4
    std::vector<int> flags;   // BodyFlags
    std::vector<int> env;
    int n = 0;
};
struct HostGeoms {
    std::vector<int> type;
    std::vector<float> dims;  // 4
    std::vector<int> body;    // -1 static
    StableVec<float> pos;   // 4 (static pose; refreshed from body for getters)   (dGeomGetPosition / Rotation)
    StableVec<float> R;     // 12
    std::vector<uint32_t> cat, col;
    std::vector<int> env;     // -1 = every env
    std::vector<int> alive;   // 0 = destroyed (never collides)
    int n = 0;
};
HostBodies &eng_bodies(Engine *);
HostGeoms &eng_geoms(Engine *);
int eng_add_body(Engine *);  // default dBodyCreate state; returns index
int eng_add_geom(Engine *);  // returns index
void eng_reset_body(Engine *, int i); // slot re-use: creation defaults, whole record re-sent
void eng_reset_geom(Engine *, int i);
int eng_add_mesh(Engine *, const float *verts, int nv, const int *tris, int nt);
void eng_mark_bodies_dirty(Engine *);  // host mirror changed -> upload before next device op
void eng_mark_geoms_dirty(Engine *);
void eng_mark_forces_dirty(Engine *);
void eng_set_num_envs(Engine *, int n);
void eng_set_capacity(Engine *, long max_pairs, long max_manifolds);
// incremental edits: only the named fields of the named body / the named geom are sent to the device at the
// next collide or step (one packed copy + one scatter kernel), the host mirrors need not be fresh
enum BodyField { FLD_POS = 1, FLD_ROT = 2, FLD_LVEL = 4, FLD_AVEL = 8, FLD_MASS = 16, FLD_FORCE = 32, FLD_ALL = 63 };
void eng_mark_body_fields(Engine *, int body, int fields);
void eng_mark_geom(Engine *, int geom);
void eng_set_big_extent(Engine *, float extent);
void eng_set_broadphase(Engine *, int mode);
void eng_set_solver_mode(Engine *, int mode, int env_group);
void eng_set_colour_spread(Engine *, int k);
void eng_set_contact_units(Engine *, int per_contact);

// device ops (all asynchronous on the engine stream unless stated)
void eng_sync_to_device(Engine *);              // upload dirty mirrors
void eng_sync_to_host(Engine *);                // download body state into the mirrors (blocking)
void eng_collide(Engine *, int max_contacts);   // broadphase + narrowphase; contacts stay on device
// compat mode: fetch pair list + contacts to the host (blocking). Arrays are engine-owned.
struct HostPairs {
    int n_pairs = 0;
    const int *g1 = nullptr, *g2 = nullptr;     // geom indices, canonical order
    const int *first = nullptr, *count = nullptr; // contact range per pair
    const float *pos_depth = nullptr;           // 4 per contact
    const float *normal_side = nullptr;         // 4 per contact (w = bitcast int side2 / triangle)
};
HostPairs eng_fetch_pairs(Engine *);
HostPairs eng_collide_pair(Engine *, int g1, int g2, int max_contacts);
// step with device-resident contacts of the last eng_collide and one surface for every contact
void eng_step_device_contacts(Engine *, float h, const Surface &surf);
// step with host-provided contact joints (compat mode)
void eng_step_host_contacts(Engine *, float h, const HostContact *contacts, int n);
// snapshot: 16 floats per body in the reference's GetTransformMat layout (src/main.c:602-622)
const float *eng_snapshot_device(Engine *);
void eng_set_snapshot_format(Engine *, int fmt); // 0: 16 floats, 1: 12 floats, 2: pos + quaternion
int eng_snapshot_format(Engine *);
void eng_snapshot_to_host(Engine *, float *dst, int first, int count, bool blocking);
// upload per-body external force/torque (6 floats per body) for the next step
void eng_set_forces(Engine *, const float *f6, int n);
// halo exchange: gather / scatter body states (16 floats each) through device buffers (async)
void eng_pack_states_device(Engine *, const int *d_idx, int n, float *d_out);
void eng_unpack_states_device(Engine *, const int *d_idx, int n, const float *d_in);
void eng_pack_impulses_device(Engine *, const int *d_idx, int n, float *d_out);
void eng_add_impulses_device(Engine *, const int *d_idx, int n, const float *d_in);
void eng_set_keep_impulses(Engine *, int on);
void eng_select_bodies_device(Engine *, int axis, float lo, float hi, const int *d_mask, int *d_idx_out, int cap, int *d_count);
void eng_pack_bodies_device(Engine *, const int *d_idx, int cap, const int *d_body_geom, float *d_out);
void eng_unpack_bodies_device(Engine *, const int *d_ghost_body, const int *d_ghost_geom, int cap, const float *d_in);
// wait for everything queued on the engine stream
void eng_wait(Engine *);
StepStats eng_stats(Engine *);                  // blocking: stats of the last collide/step
// solver order export for parity tests: per device contact (pair-major) the solve rank
int eng_export_solver_order(Engine *, int *pair_g1, int *pair_g2, int *pair_k, int cap);
// device timer on the engine stream (CUDA events): start / stop / elapsed ms (blocking)
void eng_timer_start(Engine *);
void eng_timer_stop(Engine *);
float eng_timer_elapsed_ms(Engine *);
float eng_timer_elapsed_between_ms(Engine *start, Engine *stop); // start's start event -> stop's stop event (same device)
long eng_launch_count();
// wire image of the reference's MsgUpdateBodies: bind the slot table once, pack after any step
void eng_bind_msg_slots(Engine *, int n_slots, const int *body, const int *geom, const int *type, const float *size3,
                        const unsigned *rgba);
size_t eng_pack_msg(Engine *, void *dst, int msg_type, bool blocking);
float eng_barrier_bench(Engine *, int iters); // microseconds per grid barrier (diagnostic)
// event-timed sections of the last step, milliseconds (collide, prep+colour+rows, solve+tail)
void eng_last_timings(Engine *, float out[4]);
void eng_enable_timing(Engine *, int on);
void eng_stage_timings(Engine *, float out[5]); // broadphase, narrowphase, prepare, solve, whole tick

} // namespace ob
