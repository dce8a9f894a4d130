// slab.cu -- the slab-decomposition driver of one large world (BASELINE config 5) in C/C++,
// behind include/ode_b200.h: the host application stays C (north_star), the halo exchange is NCCL send/recv over
// NVLink issued by the library itself, and a tick is ordered by CUDA events only -- no host synchronisation between
// the pack kernels, the transfers, the tick and the impulse return.
//
// Per tick (lower slab owns a contact across the face it shares with the next slab):
//   1. rank r selects, on the device, its bodies with x < face_left + margin and packs them whole (48 floats each)
//      for rank r-1;                                                              [engine stream]
//   2. state halo: ncclSend to r-1 / ncclRecv from r+1 in one group              [comm stream, after event 1]
//   3. unpack into the ghost pool, dSpaceCollideDeviceB200 + dWorldQuickStep, pack the ghosts' contact impulses
//                                                                                 [engine stream, after event 2]
//   4. impulse halo: ncclSend to r+1 / ncclRecv from r-1                          [comm stream, after event 3]
//   5. add the received impulses to the owned boundary bodies                    [engine stream, after event 4]
// Messages have the ghost pool's fixed capacity (the selected count never leaves the device); at 2 M bodies per
// GPU that is 18.9 MB down + 3.1 MB up per face per tick, ~30 us of a ~10 ms tick at NVLink rates.  Ownership
// migration (every few ticks, host-assisted: records come to the host, bodies are re-created through the ODE handle
// API on the new owner) is dSlabMigrateB200.  With several slabs in ONE process (tests on one GPU) the transport is a
// device-to-device copy ordered by the same events (dSlabConnectLocalB200 / dSlabTickLocalB200).
//
// NCCL is looked up at run time (dlopen libnccl.so.2), so libode_b200.so does not depend on it unless a slab
// communicator is created.
#include <dlfcn.h>
#include <string.h>

#include <vector>

#include "engine.h"
#include "ode/ode.h"
#include "ode_b200.h"

using namespace ob;

namespace {

constexpr int REC = 48; // floats per whole-body record (dWorldPackBodiesDeviceB200)
constexpr unsigned long CAT_MAP = 1, CAT_OBJ = 2, CAT_GHOST = 4;

// ---- the few NCCL entry points the driver needs
typedef struct ncclComm *ncclComm_t;
struct NcclId { char internal[128]; };
struct NcclApi {
    int (*GetUniqueId)(NcclId *);
    int (*CommInitRank)(ncclComm_t *, int, NcclId, int);
    int (*CommDestroy)(ncclComm_t);
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*GroupStart)();
    int (*GroupEnd)();
    const char *(*GetErrorString)(int);
    void *lib = nullptr;
};
NcclApi g_nccl;
constexpr int NCCL_FLOAT32 = 7;

bool nccl_load() {
    if (g_nccl.lib) return true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return false;
#define LOAD(field, name) *(void **)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) return false
    LOAD(GetUniqueId, "ncclGetUniqueId"); LOAD(CommInitRank, "ncclCommInitRank"); LOAD(CommDestroy, "ncclCommDestroy");
    LOAD(Send, "ncclSend"); LOAD(Recv, "ncclRecv"); LOAD(GroupStart, "ncclGroupStart"); LOAD(GroupEnd, "ncclGroupEnd");
    LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
    g_nccl.lib = h;
    return true;
}
#define OB_NCCL(call)                                                                                         \
    do {                                                                                                      \
        int r__ = (call);                                                                                     \
        if (r__ != 0) {                                                                                       \
            fprintf(stderr, "libode_b200: NCCL error %s at %s:%d\n", g_nccl.GetErrorString(r__), __FILE__, __LINE__); \
            abort();                                                                                          \
        }                                                                                                     \
    } while (0)

template <typename T>
T *dalloc(size_t n) {
    T *p = nullptr;
    OB_CUDA(ob_malloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    return p;
}

} // namespace

struct dxSlabB200 {
    dWorldID world;
    dSpaceID space;
    int rank, n_ranks, device;
    dSlabLayoutB200 L;
    cudaStream_t st, comm_st;
    cudaEvent_t ev_pack, ev_comm;
    ncclComm_t comm = nullptr;
    dxSlabB200 *lower = nullptr, *upper = nullptr; // same-process neighbours (local transport)
    bool has_left, has_right;
    int cap_bodies, n_bodies, n_owned;
    std::vector<int> h_mask, h_geom; // host mirrors of own_mask / body_geom (only the host changes them)
    int *own_mask, *body_geom, *count;
    // face shared with rank-1: our boundary bodies go down, their impulses come back
    int *send_state_idx = nullptr;
    float *send_state_buf = nullptr, *recv_imp_buf = nullptr;
    // face shared with rank+1: ghost pool
    int *ghost_body = nullptr, *ghost_geom = nullptr;
    float *recv_state_buf = nullptr, *send_imp_buf = nullptr;
    // migration, both faces: [0] = left, [1] = right
    int *send_mig_idx[2] = {nullptr, nullptr};
    float *send_mig_buf[2] = {nullptr, nullptr}, *recv_mig_buf[2] = {nullptr, nullptr};
    long migrated_in = 0, migrated_out = 0, ticks = 0;
};

extern "C" int dSlabGetUniqueIdB200(char id[128]) {
    if (!nccl_load()) return 0;
    NcclId u;
    OB_NCCL(g_nccl.GetUniqueId(&u));
    memcpy(id, u.internal, 128);
    return 1;
}

extern "C" dSlabID dSlabCreateB200(dWorldID world, dSpaceID space, int rank, int n_ranks, const char *nccl_id,
                                   const dSlabLayoutB200 *layout) {
    dxSlabB200 *s = new dxSlabB200();
    s->world = world; s->space = space; s->rank = rank; s->n_ranks = n_ranks; s->L = *layout;
    s->has_left = rank > 0; s->has_right = rank < n_ranks - 1;
    dWorldSetSlotReuseB200(world, 1); // arriving migrants take the body / geom slots of the ones that left
    s->st = (cudaStream_t)dWorldGetStreamB200(world);
    s->device = dWorldGetDeviceB200(world); // not the calling thread's current device
    OB_CUDA(cudaSetDevice(s->device));
    OB_CUDA(cudaStreamCreateWithFlags(&s->comm_st, cudaStreamNonBlocking));
    OB_CUDA(cudaEventCreateWithFlags(&s->ev_pack, cudaEventDisableTiming));
    OB_CUDA(cudaEventCreateWithFlags(&s->ev_comm, cudaEventDisableTiming));
    const int P = layout->pool, M = layout->mig_cap;
    s->n_bodies = dWorldGetNumBodiesB200(world);
    s->n_owned = layout->n_own;
    s->cap_bodies = s->n_bodies + 16 * M;
    s->h_mask.assign((size_t)s->cap_bodies, 0);
    s->h_geom.assign((size_t)s->cap_bodies, -1);
    for (int b = 0; b < layout->n_own; b++) s->h_mask[b] = 1;
    for (int b = 0; b < s->n_bodies; b++) s->h_geom[b] = layout->n_static + b; // scene order: static geoms, then one geom per body
    s->own_mask = dalloc<int>(s->cap_bodies); s->body_geom = dalloc<int>(s->cap_bodies); s->count = dalloc<int>(16);
    OB_CUDA(cudaMemcpy(s->own_mask, s->h_mask.data(), sizeof(int) * s->cap_bodies, cudaMemcpyHostToDevice));
    OB_CUDA(cudaMemcpy(s->body_geom, s->h_geom.data(), sizeof(int) * s->cap_bodies, cudaMemcpyHostToDevice));
    OB_CUDA(cudaMemset(s->count, 0, sizeof(int) * 16));
    if (s->has_left) {
        s->send_state_idx = dalloc<int>(P); s->send_state_buf = dalloc<float>((size_t)P * REC); s->recv_imp_buf = dalloc<float>((size_t)P * 8);
        OB_CUDA(cudaMemset(s->send_state_idx, 0xff, sizeof(int) * P));
        OB_CUDA(cudaMemset(s->recv_imp_buf, 0, sizeof(float) * 8 * P));
    }
    if (s->has_right) {
        std::vector<int> gb((size_t)P), gg((size_t)P);
        for (int i = 0; i < P; i++) { gb[i] = layout->pool_first_body + i; gg[i] = layout->pool_first_geom + i; }
        s->ghost_body = dalloc<int>(P); s->ghost_geom = dalloc<int>(P);
        OB_CUDA(cudaMemcpy(s->ghost_body, gb.data(), sizeof(int) * P, cudaMemcpyHostToDevice));
        OB_CUDA(cudaMemcpy(s->ghost_geom, gg.data(), sizeof(int) * P, cudaMemcpyHostToDevice));
        s->recv_state_buf = dalloc<float>((size_t)P * REC); s->send_imp_buf = dalloc<float>((size_t)P * 8);
        OB_CUDA(cudaMemset(s->recv_state_buf, 0, sizeof(float) * REC * P));
    }
    for (int f = 0; f < 2; f++) {
        if (!(f == 0 ? s->has_left : s->has_right)) continue;
        s->send_mig_idx[f] = dalloc<int>(M); s->send_mig_buf[f] = dalloc<float>((size_t)M * REC); s->recv_mig_buf[f] = dalloc<float>((size_t)M * REC);
        OB_CUDA(cudaMemset(s->send_mig_idx[f], 0xff, sizeof(int) * M));
    }
    dWorldSetKeepImpulsesB200(world, 1); // the impulse halo reads the step's accumulators
    if (nccl_id && n_ranks > 1) {
        if (!nccl_load()) { fprintf(stderr, "libode_b200: dSlabCreateB200: libnccl.so.2 not found\n"); abort(); }
        NcclId u;
        memcpy(u.internal, nccl_id, 128);
        OB_NCCL(g_nccl.CommInitRank(&s->comm, n_ranks, u, rank));
    }
    return s;
}

extern "C" void dSlabDestroyB200(dSlabID s) {
    if (!s) return;
    cudaStreamSynchronize(s->st); cudaStreamSynchronize(s->comm_st);
    if (s->comm) g_nccl.CommDestroy(s->comm);
    ob_free(s->own_mask); ob_free(s->body_geom); ob_free(s->count);
    ob_free(s->send_state_idx); ob_free(s->send_state_buf); ob_free(s->recv_imp_buf);
    ob_free(s->ghost_body); ob_free(s->ghost_geom); ob_free(s->recv_state_buf); ob_free(s->send_imp_buf);
    for (int f = 0; f < 2; f++) { ob_free(s->send_mig_idx[f]); ob_free(s->send_mig_buf[f]); ob_free(s->recv_mig_buf[f]); }
    cudaEventDestroy(s->ev_pack); cudaEventDestroy(s->ev_comm); cudaStreamDestroy(s->comm_st);
    delete s;
}

extern "C" void dSlabConnectLocalB200(dSlabID lower, dSlabID upper) { lower->upper = upper; upper->lower = lower; }

// ---- phases of a tick (each only enqueues work)
static void phase_pack_state(dxSlabB200 *s) {
    // local transport: the lower neighbour's copy out of send_state_buf (last tick) must be over before it is rewritten
    if (!s->comm && s->lower) OB_CUDA(cudaStreamWaitEvent(s->st, s->lower->ev_comm, 0));
    if (s->has_left) {
        dWorldSelectBodiesDeviceB200(s->world, 0, -INFINITY, s->L.face_left + s->L.margin, s->own_mask, s->send_state_idx, s->L.pool, s->count);
        dWorldPackBodiesDeviceB200(s->world, s->send_state_idx, s->L.pool, s->body_geom, s->send_state_buf);
    }
    OB_CUDA(cudaEventRecord(s->ev_pack, s->st));
}
// kind 0: states down (send to rank-1, receive from rank+1); kind 1: impulses up (send to rank+1, receive from rank-1)
static void phase_exchange(dxSlabB200 *s, int kind) {
    const size_t P = (size_t)s->L.pool;
    OB_CUDA(cudaStreamWaitEvent(s->comm_st, s->ev_pack, 0));
    if (s->comm) {
        OB_NCCL(g_nccl.GroupStart());
        if (kind == 0) {
            if (s->has_left) OB_NCCL(g_nccl.Send(s->send_state_buf, P * REC, NCCL_FLOAT32, s->rank - 1, s->comm, s->comm_st));
            if (s->has_right) OB_NCCL(g_nccl.Recv(s->recv_state_buf, P * REC, NCCL_FLOAT32, s->rank + 1, s->comm, s->comm_st));
        } else {
            if (s->has_right) OB_NCCL(g_nccl.Send(s->send_imp_buf, P * 8, NCCL_FLOAT32, s->rank + 1, s->comm, s->comm_st));
            if (s->has_left) OB_NCCL(g_nccl.Recv(s->recv_imp_buf, P * 8, NCCL_FLOAT32, s->rank - 1, s->comm, s->comm_st));
        }
        OB_NCCL(g_nccl.GroupEnd());
    } else {
        // local transport: the receiver pulls from its same-process neighbour, after that neighbour's pack
        if (kind == 0 && s->upper) {
            OB_CUDA(cudaStreamWaitEvent(s->comm_st, s->upper->ev_pack, 0));
            OB_CUDA(cudaMemcpyAsync(s->recv_state_buf, s->upper->send_state_buf, P * REC * sizeof(float), cudaMemcpyDeviceToDevice, s->comm_st));
        }
        if (kind == 1 && s->lower) {
            OB_CUDA(cudaStreamWaitEvent(s->comm_st, s->lower->ev_pack, 0));
            OB_CUDA(cudaMemcpyAsync(s->recv_imp_buf, s->lower->send_imp_buf, P * 8 * sizeof(float), cudaMemcpyDeviceToDevice, s->comm_st));
        }
    }
    OB_CUDA(cudaEventRecord(s->ev_comm, s->comm_st));
    OB_CUDA(cudaStreamWaitEvent(s->st, s->ev_comm, 0));
}
static void phase_run(dxSlabB200 *s, float h, int max_contacts) {
    if (!s->comm && s->upper) OB_CUDA(cudaStreamWaitEvent(s->st, s->upper->ev_comm, 0)); // ... and of send_imp_buf
    if (s->has_right) dWorldUnpackBodiesDeviceB200(s->world, s->ghost_body, s->ghost_geom, s->L.pool, s->recv_state_buf);
    dSpaceCollideDeviceB200(s->space, max_contacts);
    dWorldQuickStep(s->world, h);
    if (s->has_right) dWorldPackImpulsesDeviceB200(s->world, s->ghost_body, s->L.pool, s->send_imp_buf);
    OB_CUDA(cudaEventRecord(s->ev_pack, s->st));
}
static void phase_finish(dxSlabB200 *s) {
    if (s->has_left) dWorldAddImpulsesDeviceB200(s->world, s->send_state_idx, s->L.pool, s->recv_imp_buf);
    s->ticks++;
}

extern "C" void dSlabTickB200(dSlabID s, dReal h, int max_contacts) {
    phase_pack_state(s);
    phase_exchange(s, 0);
    phase_run(s, h, max_contacts);
    phase_exchange(s, 1);
    phase_finish(s);
}
extern "C" void dSlabTickLocalB200(dSlabID *slabs, int n, dReal h, int max_contacts) {
    for (int i = 0; i < n; i++) { OB_CUDA(cudaSetDevice(slabs[i]->device)); phase_pack_state(slabs[i]); }
    for (int i = 0; i < n; i++) phase_exchange(slabs[i], 0);
    for (int i = 0; i < n; i++) phase_run(slabs[i], h, max_contacts);
    for (int i = 0; i < n; i++) phase_exchange(slabs[i], 1);
    for (int i = 0; i < n; i++) phase_finish(slabs[i]);
}

// ---- ownership migration: bodies whose centre crossed a face by more than the hysteresis change owner
static void mig_pack(dxSlabB200 *s) {
    for (int f = 0; f < 2; f++) {
        if (!s->send_mig_idx[f]) continue;
        const float lo = f == 0 ? -INFINITY : s->L.face_right + s->L.hyst, hi = f == 0 ? s->L.face_left - s->L.hyst : INFINITY;
        dWorldSelectBodiesDeviceB200(s->world, 0, lo, hi, s->own_mask, s->send_mig_idx[f], s->L.mig_cap, s->count + 4 * (f + 1));
        dWorldPackBodiesDeviceB200(s->world, s->send_mig_idx[f], s->L.mig_cap, s->body_geom, s->send_mig_buf[f]);
    }
    OB_CUDA(cudaEventRecord(s->ev_pack, s->st));
}
static void mig_exchange(dxSlabB200 *s) {
    const size_t n = (size_t)s->L.mig_cap * REC;
    OB_CUDA(cudaStreamWaitEvent(s->comm_st, s->ev_pack, 0));
    if (s->comm) {
        OB_NCCL(g_nccl.GroupStart());
        if (s->has_left) {
            OB_NCCL(g_nccl.Send(s->send_mig_buf[0], n, NCCL_FLOAT32, s->rank - 1, s->comm, s->comm_st));
            OB_NCCL(g_nccl.Recv(s->recv_mig_buf[0], n, NCCL_FLOAT32, s->rank - 1, s->comm, s->comm_st));
        }
        if (s->has_right) {
            OB_NCCL(g_nccl.Send(s->send_mig_buf[1], n, NCCL_FLOAT32, s->rank + 1, s->comm, s->comm_st));
            OB_NCCL(g_nccl.Recv(s->recv_mig_buf[1], n, NCCL_FLOAT32, s->rank + 1, s->comm, s->comm_st));
        }
        OB_NCCL(g_nccl.GroupEnd());
    } else {
        if (s->lower) {
            OB_CUDA(cudaStreamWaitEvent(s->comm_st, s->lower->ev_pack, 0));
            OB_CUDA(cudaMemcpyAsync(s->recv_mig_buf[0], s->lower->send_mig_buf[1], n * sizeof(float), cudaMemcpyDeviceToDevice, s->comm_st));
        }
        if (s->upper) {
            OB_CUDA(cudaStreamWaitEvent(s->comm_st, s->upper->ev_pack, 0));
            OB_CUDA(cudaMemcpyAsync(s->recv_mig_buf[1], s->upper->send_mig_buf[0], n * sizeof(float), cudaMemcpyDeviceToDevice, s->comm_st));
        }
    }
    OB_CUDA(cudaEventRecord(s->ev_comm, s->comm_st));
}
static void mig_apply(dxSlabB200 *s) {
    OB_CUDA(cudaEventSynchronize(s->ev_comm)); // migration is the one host-assisted step (every few ticks)
    const int M = s->L.mig_cap;
    std::vector<int> out_idx((size_t)M);
    std::vector<float> rec((size_t)M * REC);
    for (int f = 0; f < 2; f++) {
        if (!s->send_mig_idx[f]) continue;
        OB_CUDA(cudaMemcpy(out_idx.data(), s->send_mig_idx[f], sizeof(int) * M, cudaMemcpyDeviceToHost));
        OB_CUDA(cudaMemcpy(rec.data(), s->recv_mig_buf[f], sizeof(float) * REC * M, cudaMemcpyDeviceToHost));
        // leaving: destroy what was sent
        for (int i = 0; i < M; i++) {
            const int b = out_idx[i];
            if (b < 0) continue;
            dGeomDestroy(dSpaceGetGeomB200(s->space, s->h_geom[b]));
            dBodyDestroy(dWorldGetBodyB200(s->world, b));
            s->h_mask[b] = 0;
            OB_CUDA(cudaMemcpyAsync(s->own_mask + b, &s->h_mask[b], sizeof(int), cudaMemcpyHostToDevice, s->st));
            s->n_owned--; s->migrated_out++;
        }
        // arriving: re-create through the handle API (queued as patches, sent with the next collide)
        for (int i = 0; i < M; i++) {
            const float *r = rec.data() + (size_t)i * REC;
            int type;
            memcpy(&type, r + 15, 4);
            if (type < 0) continue;
            dBodyID b = dBodyCreate(s->world);
            dBodySetPosition(b, r[0], r[1], r[2]);
            dBodySetQuaternion(b, r + 4);
            dBodySetLinearVel(b, r[8], r[9], r[10]);
            dBodySetAngularVel(b, r[12], r[13], r[14]);
            dMass m;
            memset(&m, 0, sizeof(m));
            m.mass = r[11];
            for (int k = 0; k < 12; k++) m.I[k] = r[20 + k];
            dBodySetMass(b, &m);
            int flags;
            memcpy(&flags, r + 44, 4);
            dBodySetGyroscopicMode(b, (flags & BF_GYRO) ? 1 : 0);
            dGeomID g = type == G_SPHERE ? dCreateSphere(s->space, r[16]) : dCreateBox(s->space, r[16], r[17], r[18]);
            dGeomSetBody(g, b);
            dGeomSetCategoryBits(g, CAT_OBJ);
            dGeomSetCollideBits(g, CAT_OBJ | CAT_MAP | CAT_GHOST);
            const int bi = dBodyGetIndexB200(b), gi = dGeomGetIndexB200(g);
            if (bi >= s->cap_bodies) {
                fprintf(stderr, "libode_b200: dSlabMigrateB200: body capacity (%d) exhausted by migration\n", s->cap_bodies);
                abort();
            }
            s->h_mask[bi] = 1; s->h_geom[bi] = gi;
            OB_CUDA(cudaMemcpyAsync(s->own_mask + bi, &s->h_mask[bi], sizeof(int), cudaMemcpyHostToDevice, s->st));
            OB_CUDA(cudaMemcpyAsync(s->body_geom + bi, &s->h_geom[bi], sizeof(int), cudaMemcpyHostToDevice, s->st));
            s->n_bodies++; s->n_owned++; s->migrated_in++;
        }
    }
    OB_CUDA(cudaStreamSynchronize(s->st)); // the mirrors' entries were sources of the small copies above
}
extern "C" void dSlabMigrateB200(dSlabID s) { mig_pack(s); mig_exchange(s); mig_apply(s); }
extern "C" void dSlabMigrateLocalB200(dSlabID *slabs, int n) {
    for (int i = 0; i < n; i++) mig_pack(slabs[i]);
    for (int i = 0; i < n; i++) mig_exchange(slabs[i]);
    for (int i = 0; i < n; i++) mig_apply(slabs[i]);
}

extern "C" void dSlabGetInfoB200(dSlabID s, dSlabInfoB200 *o) {
    o->n_owned = s->n_owned; o->migrated_in = s->migrated_in; o->migrated_out = s->migrated_out; o->ticks = s->ticks;
    o->halo_bytes_per_tick = (long)((s->has_left ? (size_t)s->L.pool * REC * 4 : 0) + (s->has_right ? (size_t)s->L.pool * 8 * 4 : 0));
    int c[16];
    OB_CUDA(cudaMemcpy(c, s->count, sizeof(c), cudaMemcpyDeviceToHost)); // blocking: diagnostics only
    o->halo_selected = c[0];
    o->halo_overflow = c[0] > s->L.pool ? c[0] - s->L.pool : 0;
    o->mig_overflow = std::max(0, c[4] - s->L.mig_cap) + std::max(0, c[8] - s->L.mig_cap);
}
