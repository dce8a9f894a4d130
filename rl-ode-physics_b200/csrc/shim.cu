// shim.cu -- the C ABI of libode_b200.so: ODE's handle-based API (include/ode/ode.h) and the
// device-resident extensions (include/ode_b200.h) on top of the engine.
//
// Handle model: IDs are pointers to small host structs that hold an
// index into the world's arrays.  Bodies and geoms have host mirrors so that dBodyGetPosition &co
// return stable `const dReal*`; the mirrors are refreshed lazily after a step.  Everything that
// computes physics runs on the GPU; this file only does bookkeeping and ODE's host-side
// conversions (dRtoQ, dMass helpers).
#include <string.h>

#include <algorithm>
#include <deque>
#include <thread>
#include <vector>

#include "dmath.cuh"
#include "engine.h"
#include "ode_b200.h"

using namespace ob;

struct dxJointGroup;

struct dxGeom;
struct dxBody {
    dxWorld *w;
    int idx;
    void *data;
    bool alive;
    std::vector<dxGeom *> geoms; // geoms attached with dGeomSetBody (so destroying a body is not a scan of the space)
};

struct dxGeom {
    dxSpace *space;
    int idx; // engine geom index
    dxBody *body;
    void *data;
    bool alive;
    float aabb_tmp[6];
};

struct dxJoint {
    dxWorld *w;
    HostContact hc;
    dxBody *b1, *b2;
};

struct dxJointGroup {
    std::deque<dxJoint> joints;
};

struct dxTriMeshData {
    std::vector<float> verts;
    std::vector<int> tris;
    dxWorld *bound_world = nullptr;
    int mesh_id = -1;
};

struct dxSpace {
    dxWorld *w;
    std::vector<dxGeom *> geoms; // by creation order (alive or not)
};

struct dxWorld {
    Engine *eng;
    std::deque<dxBody> bodies;
    std::deque<dxGeom> geoms;          // all geoms of all spaces bound to this world, by engine index
    std::vector<dxSpace *> spaces;
    std::vector<dxJointGroup *> groups; // groups that hold joints of this world
    dSurfaceParameters surface;
    int max_contacts = 8;
    int step_iters = 0;   // dWorldStep parity mode (dWorldSetStepSolverB200)
    float step_tol = 0.f;
    bool device_contacts_pending = false;
    // dWorldSetSlotReuseB200: indices of destroyed bodies / sphere and box geoms, handed out again (lowest first) by the next
    // dBodyCreate / dCreateSphere / dCreateBox instead of growing the arrays
    bool reuse_slots = false;
    std::vector<int> free_bodies, free_geoms;
    // callback context
    bool in_callback = false;
    int cb_g1 = -1, cb_g2 = -1, cb_first = 0, cb_count = 0;
    HostPairs cb_pairs;
    float ret_vec[4];
};

static std::vector<dxWorld *> g_worlds;
static int g_device = -1;

static int default_device() {
    if (g_device >= 0) return g_device;
    const char *s = getenv("ODE_B200_DEVICE");
    if (!s) s = getenv("LOCAL_RANK");
    return s ? atoi(s) : 0;
}

static void fatal(const char *msg) {
    fprintf(stderr, "libode_b200: %s\n", msg);
    abort();
}

static dxWorld *current_world() {
    for (size_t i = g_worlds.size(); i-- > 0;)
        if (g_worlds[i]) return g_worlds[i];
    return nullptr;
}

static dxWorld *space_world(dxSpace *s) {
    if (!s->w) {
        s->w = current_world();
        if (!s->w) fatal("a space needs a world: create a dWorldID before adding geoms");
        s->w->spaces.push_back(s);
    }
    return s->w;
}

// host mirrors must be current before the host edits or reads them
static void fresh(dxWorld *w) { eng_sync_to_host(w->eng); }

// ------------------------------------------------------------------ host-side ODE math helpers

extern "C" void dQtoR(const dQuaternion q, dMatrix3 R) {
    const float qq1 = 2 * q[1] * q[1], qq2 = 2 * q[2] * q[2], qq3 = 2 * q[3] * q[3];
    R[0] = 1 - qq2 - qq3; R[1] = 2 * (q[1] * q[2] - q[0] * q[3]); R[2] = 2 * (q[1] * q[3] + q[0] * q[2]); R[3] = 0;
    R[4] = 2 * (q[1] * q[2] + q[0] * q[3]); R[5] = 1 - qq1 - qq3; R[6] = 2 * (q[2] * q[3] - q[0] * q[1]); R[7] = 0;
    R[8] = 2 * (q[1] * q[3] - q[0] * q[2]); R[9] = 2 * (q[2] * q[3] + q[0] * q[1]); R[10] = 1 - qq1 - qq2; R[11] = 0;
}

extern "C" void dRtoQ(const dMatrix3 R, dQuaternion q) {
    float tr = R[0] + R[5] + R[10], s;
    if (tr >= 0) {
        s = sqrtf(tr + 1);
        q[0] = 0.5f * s;
        s = 0.5f / s;
        q[1] = (R[9] - R[6]) * s; q[2] = (R[2] - R[8]) * s; q[3] = (R[4] - R[1]) * s;
    } else {
        int c;
        if (R[5] > R[0]) c = (R[10] > R[5]) ? 2 : 1;
        else c = (R[10] > R[0]) ? 2 : 0;
        if (c == 0) {
            s = sqrtf((R[0] - (R[5] + R[10])) + 1);
            q[1] = 0.5f * s; s = 0.5f / s;
            q[2] = (R[1] + R[4]) * s; q[3] = (R[8] + R[2]) * s; q[0] = (R[9] - R[6]) * s;
        } else if (c == 1) {
            s = sqrtf((R[5] - (R[10] + R[0])) + 1);
            q[2] = 0.5f * s; s = 0.5f / s;
            q[3] = (R[6] + R[9]) * s; q[1] = (R[1] + R[4]) * s; q[0] = (R[2] - R[8]) * s;
        } else {
            s = sqrtf((R[10] - (R[0] + R[5])) + 1);
            q[3] = 0.5f * s; s = 0.5f / s;
            q[1] = (R[8] + R[2]) * s; q[2] = (R[6] + R[9]) * s; q[0] = (R[4] - R[1]) * s;
        }
    }
}

static void normalize4(float *a) {
    float l = a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3];
    if (l > 0) {
        l = 1.0f / sqrtf(l);
        a[0] *= l; a[1] *= l; a[2] *= l; a[3] *= l;
    } else {
        a[0] = 1; a[1] = 0; a[2] = 0; a[3] = 0;
    }
}

extern "C" void dRSetIdentity(dMatrix3 R) {
    memset(R, 0, sizeof(dMatrix3));
    R[0] = R[5] = R[10] = 1;
}
extern "C" void dQSetIdentity(dQuaternion q) { q[0] = 1; q[1] = q[2] = q[3] = 0; }
extern "C" void dQFromAxisAndAngle(dQuaternion q, dReal ax, dReal ay, dReal az, dReal angle) {
    float l = ax * ax + ay * ay + az * az;
    if (l > 0) {
        angle *= 0.5f;
        q[0] = cosf(angle);
        l = sinf(angle) * (1.0f / sqrtf(l));
        q[1] = ax * l; q[2] = ay * l; q[3] = az * l;
    } else dQSetIdentity(q);
}
extern "C" void dRFromAxisAndAngle(dMatrix3 R, dReal ax, dReal ay, dReal az, dReal angle) {
    dQuaternion q;
    dQFromAxisAndAngle(q, ax, ay, az, angle);
    dQtoR(q, R);
}
extern "C" void dRFromEulerAngles(dMatrix3 R, dReal phi, dReal theta, dReal psi) {
    const float sphi = sinf(phi), cphi = cosf(phi), stheta = sinf(theta), ctheta = cosf(theta), spsi = sinf(psi),
                cpsi = cosf(psi);
    R[0] = cpsi * ctheta; R[1] = spsi * ctheta; R[2] = -stheta; R[3] = 0;
    R[4] = cpsi * stheta * sphi - spsi * cphi; R[5] = spsi * stheta * sphi + cpsi * cphi; R[6] = ctheta * sphi; R[7] = 0;
    R[8] = cpsi * stheta * cphi + spsi * sphi; R[9] = spsi * stheta * cphi - cpsi * sphi; R[10] = ctheta * cphi; R[11] = 0;
}
extern "C" void dPlaneSpace(const dVector3 n, dVector3 p, dVector3 q) {
    V3 pp, qq;
    plane_space(v3(n[0], n[1], n[2]), pp, qq);
    p[0] = pp.x; p[1] = pp.y; p[2] = pp.z; p[3] = 0;
    q[0] = qq.x; q[1] = qq.y; q[2] = qq.z; q[3] = 0;
}

// ------------------------------------------------------------------ mass

extern "C" void dMassSetZero(dMass *m) { memset(m, 0, sizeof(*m)); }
extern "C" void dMassSetParameters(dMass *m, dReal themass, dReal cgx, dReal cgy, dReal cgz, dReal I11, dReal I22,
                                   dReal I33, dReal I12, dReal I13, dReal I23) {
    dMassSetZero(m);
    m->mass = themass;
    m->c[0] = cgx; m->c[1] = cgy; m->c[2] = cgz;
    m->I[0] = I11; m->I[5] = I22; m->I[10] = I33;
    m->I[1] = I12; m->I[2] = I13; m->I[6] = I23;
    m->I[4] = I12; m->I[8] = I13; m->I[9] = I23;
}
extern "C" void dMassSetSphereTotal(dMass *m, dReal total_mass, dReal radius) {
    dMassSetZero(m);
    m->mass = total_mass;
    const float II = 0.4f * total_mass * radius * radius;
    m->I[0] = II; m->I[5] = II; m->I[10] = II;
}
extern "C" void dMassSetSphere(dMass *m, dReal density, dReal radius) {
    dMassSetSphereTotal(m, (4.0f / 3.0f) * (float)M_PI * radius * radius * radius * density, radius);
}
extern "C" void dMassSetBoxTotal(dMass *m, dReal total_mass, dReal lx, dReal ly, dReal lz) {
    dMassSetZero(m);
    m->mass = total_mass;
    m->I[0] = total_mass / 12.0f * (ly * ly + lz * lz);
    m->I[5] = total_mass / 12.0f * (lx * lx + lz * lz);
    m->I[10] = total_mass / 12.0f * (lx * lx + ly * ly);
}
extern "C" void dMassSetBox(dMass *m, dReal density, dReal lx, dReal ly, dReal lz) {
    dMassSetBoxTotal(m, lx * ly * lz * density, lx, ly, lz);
}
extern "C" void dMassAdjust(dMass *m, dReal newmass) {
    const float scale = newmass / m->mass;
    m->mass = newmass;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) m->I[i * 4 + j] *= scale;
}

static void invert3(const float *I, float *inv) {
    const float a = I[0], b = I[1], c = I[2], d = I[4], e = I[5], f = I[6], g = I[8], h = I[9], i = I[10];
    const float A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
    const float det = a * A + b * B + c * C;
    const float id = 1.0f / det;
    memset(inv, 0, 12 * sizeof(float));
    inv[0] = A * id; inv[1] = -(b * i - c * h) * id; inv[2] = (b * f - c * e) * id;
    inv[4] = B * id; inv[5] = (a * i - c * g) * id; inv[6] = -(a * f - c * d) * id;
    inv[8] = C * id; inv[9] = -(a * h - b * g) * id; inv[10] = (a * e - b * d) * id;
}

// ------------------------------------------------------------------ library / world

extern "C" void dInitODE(void) {}
extern "C" int dInitODE2(unsigned int) { return 1; }
extern "C" void dCloseODE(void) {}
extern "C" void dSetDeviceB200(int device) { g_device = device; }
extern "C" int dGetDeviceB200(void) { return default_device(); }
extern "C" int dWorldGetDeviceB200(dWorldID w) { return eng_device(w->eng); }
extern "C" void *dAllocPinnedB200(size_t bytes, int write_combined) {
    void *p = nullptr;
    OB_CUDA(cudaSetDevice(default_device()));
    OB_CUDA(cudaHostAlloc(&p, bytes ? bytes : 1, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
    return p;
}
extern "C" void dFreePinnedB200(void *p) {
    if (p) OB_CUDA(cudaFreeHost(p));
}

extern "C" dWorldID dWorldCreate(void) {
    dxWorld *w = new dxWorld();
    w->eng = eng_create(default_device());
    memset(&w->surface, 0, sizeof(w->surface));
    // the reference's NearCallback policy, src/main.c:684-687
    w->surface.mode = dContactBounce;
    w->surface.bounce = 0.2f;
    w->surface.bounce_vel = 0.1f;
    w->surface.mu = dInfinity;
    g_worlds.push_back(w);
    return w;
}

extern "C" void dWorldDestroy(dWorldID w) {
    if (!w) return;
    for (auto &p : g_worlds) if (p == w) p = nullptr;
    for (dxSpace *s : w->spaces)
        if (s) s->w = nullptr;
    eng_destroy(w->eng);
    delete w;
}

extern "C" void dWorldSetGravity(dWorldID w, dReal x, dReal y, dReal z) {
    WorldParams &p = eng_params(w->eng);
    p.gravity[0] = x; p.gravity[1] = y; p.gravity[2] = z;
}
extern "C" void dWorldGetGravity(dWorldID w, dVector3 g) {
    WorldParams &p = eng_params(w->eng);
    g[0] = p.gravity[0]; g[1] = p.gravity[1]; g[2] = p.gravity[2]; g[3] = 0;
}
extern "C" void dWorldSetERP(dWorldID w, dReal erp) { eng_params(w->eng).erp = erp; }
extern "C" dReal dWorldGetERP(dWorldID w) { return eng_params(w->eng).erp; }
extern "C" void dWorldSetCFM(dWorldID w, dReal cfm) { eng_params(w->eng).cfm = cfm; }
extern "C" dReal dWorldGetCFM(dWorldID w) { return eng_params(w->eng).cfm; }
extern "C" void dWorldSetQuickStepNumIterations(dWorldID w, int num) { eng_params(w->eng).iters = num; }
extern "C" int dWorldGetQuickStepNumIterations(dWorldID w) { return eng_params(w->eng).iters; }
extern "C" void dWorldSetQuickStepW(dWorldID w, dReal v) { eng_params(w->eng).sor_w = v; }
extern "C" dReal dWorldGetQuickStepW(dWorldID w) { return eng_params(w->eng).sor_w; }
extern "C" void dWorldSetContactMaxCorrectingVel(dWorldID w, dReal vel) { eng_params(w->eng).max_vel = vel; }
extern "C" dReal dWorldGetContactMaxCorrectingVel(dWorldID w) { return eng_params(w->eng).max_vel; }
extern "C" void dWorldSetContactSurfaceLayer(dWorldID w, dReal depth) { eng_params(w->eng).min_depth = depth; }
extern "C" dReal dWorldGetContactSurfaceLayer(dWorldID w) { return eng_params(w->eng).min_depth; }

static Surface to_surface(const dSurfaceParameters &s) {
    Surface o;
    o.mode = s.mode; o.mu = s.mu; o.mu2 = s.mu2; o.bounce = s.bounce; o.bounce_vel = s.bounce_vel;
    o.soft_erp = s.soft_erp; o.soft_cfm = s.soft_cfm; o.motion1 = s.motion1; o.motion2 = s.motion2;
    o.motionN = s.motionN; o.slip1 = s.slip1; o.slip2 = s.slip2;
    if (o.mode & dContactFDir1) fatal("dContactFDir1 is not supported (friction directions come from dPlaneSpace)");
    return o;
}

extern "C" int dWorldQuickStep(dWorldID w, dReal h) {
    if (!w || !(h > 0)) return 0;
    // contact joints created through dJointCreateContact (compat mode) take precedence
    std::vector<HostContact> hc;
    for (dxJointGroup *g : w->groups)
        for (dxJoint &j : g->joints)
            if (j.w == w) {
                j.hc.b1 = (j.b1 && j.b1->alive) ? j.b1->idx : -1;
                j.hc.b2 = (j.b2 && j.b2->alive) ? j.b2->idx : -1;
                hc.push_back(j.hc);
            }
    if (!hc.empty() || !w->device_contacts_pending) {
        eng_step_host_contacts(w->eng, h, hc.data(), (int)hc.size());
    } else {
        eng_step_device_contacts(w->eng, h, to_surface(w->surface));
    }
    w->device_contacts_pending = false;
    return 1;
}
// dWorldStep is libode's exact (Dantzig) stepper (what the reference calls, src/main.c:213).  Default: the step's LCP is
// solved exactly on the device (solver_exact.cu) when the world is small enough, else QuickStep's sweeps;
// dWorldSetStepSolverB200 selects residual-terminated sweeps (max_iters > 0) or plain QuickStep (max_iters < 0) instead.
extern "C" int dWorldStep(dWorldID w, dReal h) {
    if (!w) return 0;
    WorldParams &p = eng_params(w->eng);
    if (w->step_iters < 0) return dWorldQuickStep(w, h);
    if (w->step_iters == 0) {
        p.exact = 1;
        const int r = dWorldQuickStep(w, h);
        p.exact = 0;
        return r;
    }
    const int iters = p.iters;
    p.iters = w->step_iters; p.tol = w->step_tol;
    const int r = dWorldQuickStep(w, h);
    p.iters = iters; p.tol = 0.f;
    return r;
}
extern "C" void dWorldSetStepSolverB200(dWorldID w, int max_iters, float tol) { w->step_iters = max_iters; w->step_tol = tol < 0 ? 0 : tol; }

extern "C" void dWorldSetSurfaceB200(dWorldID w, const dSurfaceParameters *s) { w->surface = *s; }
extern "C" void dWorldGetSurfaceB200(dWorldID w, dSurfaceParameters *s) { *s = w->surface; }
extern "C" void dWorldSetMaxContactsB200(dWorldID w, int n) { w->max_contacts = n < 1 ? 1 : (n > 8 ? 8 : n); }
extern "C" void dWorldSetNumEnvsB200(dWorldID w, int n) { eng_set_num_envs(w->eng, n); }
extern "C" void dWorldSetSlotReuseB200(dWorldID w, int on) { w->reuse_slots = on != 0; }
extern "C" void dWorldSetCapacityB200(dWorldID w, long mp, long mm) { eng_set_capacity(w->eng, mp, mm); }
extern "C" void dWorldSetBigExtentB200(dWorldID w, float e) { eng_set_big_extent(w->eng, e); }
extern "C" void dWorldSetBroadphaseB200(dWorldID w, int mode) { eng_set_broadphase(w->eng, mode); }
extern "C" void dWorldSetSolverModeB200(dWorldID w, int mode, int env_group) { eng_set_solver_mode(w->eng, mode, env_group); }
extern "C" void dWorldSetContactUnitsB200(dWorldID w, int per_contact) { eng_set_contact_units(w->eng, per_contact); }
extern "C" void dWorldWaitB200(dWorldID w) { eng_wait(w->eng); }
extern "C" void *dWorldGetStreamB200(dWorldID w) { return (void *)eng_stream(w->eng); }
extern "C" int dCheckGuardsB200(int verbose) { return ob::guard_check(verbose); }
extern "C" int dGuardSelfTestB200(int overrun) {
    if (ob::guard_check(0) < 0) return -1; // guards off: there is no band to write into
    unsigned char *p = nullptr;
    OB_CUDA(ob_malloc(&p, 1000));
    OB_CUDA(cudaMemset(p, 0, (size_t)(1000 + (overrun > 0 ? (overrun < 256 ? overrun : 256) : 0)))); // never leaves the band
    const int bad = ob::guard_check(0);
    ob_free(p);
    return bad;
}
extern "C" void dWorldPackStatesDeviceB200(dWorldID w, const int *d_idx, int n, float *d_out) { eng_pack_states_device(w->eng, d_idx, n, d_out); }
extern "C" void dWorldPackImpulsesDeviceB200(dWorldID w, const int *d_idx, int n, float *d_out) { eng_pack_impulses_device(w->eng, d_idx, n, d_out); }
extern "C" void dWorldAddImpulsesDeviceB200(dWorldID w, const int *d_idx, int n, const float *d_in) { eng_add_impulses_device(w->eng, d_idx, n, d_in); }
extern "C" void dWorldSetKeepImpulsesB200(dWorldID w, int on) { eng_set_keep_impulses(w->eng, on); }
extern "C" void dWorldSelectBodiesDeviceB200(dWorldID w, int axis, float lo, float hi, const int *d_mask, int *d_idx_out, int cap, int *d_count) {
    eng_select_bodies_device(w->eng, axis, lo, hi, d_mask, d_idx_out, cap, d_count);
}
extern "C" void dWorldPackBodiesDeviceB200(dWorldID w, const int *d_idx, int cap, const int *d_body_geom, float *d_out48) {
    eng_pack_bodies_device(w->eng, d_idx, cap, d_body_geom, d_out48);
}
extern "C" void dWorldUnpackBodiesDeviceB200(dWorldID w, const int *d_ghost_body, const int *d_ghost_geom, int cap, const float *d_in48) {
    eng_unpack_bodies_device(w->eng, d_ghost_body, d_ghost_geom, cap, d_in48);
}
extern "C" void dWorldUnpackStatesDeviceB200(dWorldID w, const int *d_idx, int n, const float *d_in) { eng_unpack_states_device(w->eng, d_idx, n, d_in); }
extern "C" void dWorldTimerStartB200(dWorldID w) { eng_timer_start(w->eng); }
extern "C" void dWorldTimerStopB200(dWorldID w) { eng_timer_stop(w->eng); }
extern "C" float dWorldTimerElapsedB200(dWorldID w) { return eng_timer_elapsed_ms(w->eng); }
extern "C" float dWorldTimerElapsedBetweenB200(dWorldID a, dWorldID b) { return eng_timer_elapsed_between_ms(a->eng, b->eng); }
extern "C" long dGetKernelLaunchCountB200(void) { return eng_launch_count(); }
extern "C" void dWorldBindSnapshotSlotsB200(dWorldID w, int n_slots, const dBodyID *bodies, const dGeomID *geoms, const int *types,
                                            const float *size3, const unsigned *rgba) {
    std::vector<int> bi((size_t)n_slots, -1), gi((size_t)n_slots, -1);
    for (int i = 0; i < n_slots; i++) {
        if (bodies && bodies[i] && bodies[i]->alive) bi[i] = bodies[i]->idx;
        if (geoms && geoms[i] && geoms[i]->alive) gi[i] = geoms[i]->idx;
    }
    eng_bind_msg_slots(w->eng, n_slots, bi.data(), gi.data(), types, size3, rgba);
}
extern "C" size_t dWorldPackMsgUpdateBodiesB200(dWorldID w, void *dst, int msg_type, int blocking) {
    return eng_pack_msg(w->eng, dst, msg_type, blocking != 0);
}
extern "C" float dTestGridBarrierB200(dWorldID w, int iters) { return eng_barrier_bench(w->eng, iters); }
// test hook: take the whole-array upload path (refresh the mirrors, then re-send everything) instead of
// the queued field patches, so tests can check that the two are equivalent
extern "C" void dTestForceFullSyncB200(dWorldID w) {
    fresh(w);
    eng_mark_bodies_dirty(w->eng);
    eng_mark_geoms_dirty(w->eng);
}
extern "C" void dWorldEnableTimingB200(dWorldID w, int on) { eng_enable_timing(w->eng, on); }
extern "C" void dWorldGetTimingsB200(dWorldID w, float out[4]) { eng_last_timings(w->eng, out); }
extern "C" void dWorldGetStageTimingsB200(dWorldID w, float out[5]) { eng_stage_timings(w->eng, out); }
extern "C" int dWorldGetNumBodiesB200(dWorldID w) { return eng_bodies(w->eng).n; }

extern "C" void dWorldGetStatsB200(dWorldID w, dStepStatsB200 *out) {
    static_assert(sizeof(dStepStatsB200) == sizeof(StepStats), "stats layout");
    StepStats s = eng_stats(w->eng);
    memcpy(out, &s, sizeof(s));
}

// ------------------------------------------------------------------ bodies

static int pop_lowest(std::vector<int> &v) {
    auto it = std::min_element(v.begin(), v.end());
    const int i = *it;
    *it = v.back();
    v.pop_back();
    return i;
}
static dxBody *new_body(dxWorld *w) {
    if (w->reuse_slots && !w->free_bodies.empty()) {
        const int idx = pop_lowest(w->free_bodies);
        eng_reset_body(w->eng, idx);
        w->bodies[idx] = dxBody{w, idx, nullptr, true, {}};
        return &w->bodies[idx];
    }
    const int idx = eng_add_body(w->eng);
    w->bodies.push_back(dxBody{w, idx, nullptr, true, {}});
    return &w->bodies.back();
}

extern "C" dBodyID dBodyCreate(dWorldID w) {
    dxBody *b = new_body(w);
    eng_bodies(w->eng).flags[b->idx] = BF_GYRO; // ODE >= 0.13: gyroscopic mode on by default
    return b;
}

extern "C" void dBodyDestroy(dBodyID b) {
    if (!b || !b->alive) return;
    dxWorld *w = b->w;
    // ODE detaches the body's geoms; the slot stays as an inert kinematic body
    for (dxGeom *gp : b->geoms) {
        dxGeom &g = *gp;
        if (g.alive && g.body == b) {
            // the detached geom stays where the body was
            const float *bp = dBodyGetPosition(b), *bR = dBodyGetRotation(b);
            HostGeoms &hg = eng_geoms(w->eng);
            memcpy(&hg.pos[4 * g.idx], bp, 3 * sizeof(float));
            memcpy(&hg.R[12 * g.idx], bR, 12 * sizeof(float));
            g.body = nullptr;
            hg.body[g.idx] = -1;
            eng_mark_geom(w->eng, g.idx);
        }
    }
    b->geoms.clear();
    HostBodies &hb = eng_bodies(w->eng);
    hb.flags[b->idx] = BF_KINEMATIC | BF_NOGRAVITY;
    hb.pos[4 * b->idx + 3] = 0.f;
    for (int k = 0; k < 3; k++) { hb.lvel[4 * b->idx + k] = 0; hb.avel[4 * b->idx + k] = 0; }
    for (int k = 0; k < 12; k++) hb.invI[12 * b->idx + k] = 0;
    eng_mark_body_fields(w->eng, b->idx, FLD_MASS | FLD_LVEL | FLD_AVEL);
    b->alive = false;
    if (w->reuse_slots && hb.env[b->idx] == 0) w->free_bodies.push_back(b->idx);
}

#define HB(b) eng_bodies((b)->w->eng)

extern "C" void dBodySetPosition(dBodyID b, dReal x, dReal y, dReal z) {
    float *p = &HB(b).pos[4 * b->idx];
    p[0] = x; p[1] = y; p[2] = z;
    eng_mark_body_fields(b->w->eng, b->idx, FLD_POS);
}
extern "C" void dBodySetRotation(dBodyID b, const dMatrix3 R) {
    float *r = &HB(b).R[12 * b->idx], *q = &HB(b).quat[4 * b->idx];
    memcpy(r, R, 12 * sizeof(float));
    r[3] = r[7] = r[11] = 0;
    dRtoQ(r, q);
    normalize4(q);
    eng_mark_body_fields(b->w->eng, b->idx, FLD_ROT);
}
extern "C" void dBodySetQuaternion(dBodyID b, const dQuaternion qq) {
    float *r = &HB(b).R[12 * b->idx], *q = &HB(b).quat[4 * b->idx];
    memcpy(q, qq, 4 * sizeof(float));
    normalize4(q);
    dQtoR(q, r);
    eng_mark_body_fields(b->w->eng, b->idx, FLD_ROT);
}
extern "C" void dBodySetLinearVel(dBodyID b, dReal x, dReal y, dReal z) {
    float *p = &HB(b).lvel[4 * b->idx];
    p[0] = x; p[1] = y; p[2] = z;
    eng_mark_body_fields(b->w->eng, b->idx, FLD_LVEL);
}
extern "C" void dBodySetAngularVel(dBodyID b, dReal x, dReal y, dReal z) {
    float *p = &HB(b).avel[4 * b->idx];
    p[0] = x; p[1] = y; p[2] = z;
    eng_mark_body_fields(b->w->eng, b->idx, FLD_AVEL);
}
extern "C" const dReal *dBodyGetPosition(dBodyID b) { fresh(b->w); return &HB(b).pos[4 * b->idx]; }
extern "C" const dReal *dBodyGetRotation(dBodyID b) { fresh(b->w); return &HB(b).R[12 * b->idx]; }
extern "C" const dReal *dBodyGetQuaternion(dBodyID b) { fresh(b->w); return &HB(b).quat[4 * b->idx]; }
extern "C" const dReal *dBodyGetLinearVel(dBodyID b) { fresh(b->w); return &HB(b).lvel[4 * b->idx]; }
extern "C" const dReal *dBodyGetAngularVel(dBodyID b) { fresh(b->w); return &HB(b).avel[4 * b->idx]; }
extern "C" const dReal *dBodyGetForce(dBodyID b) { return &HB(b).facc[4 * b->idx]; }
extern "C" const dReal *dBodyGetTorque(dBodyID b) { return &HB(b).tacc[4 * b->idx]; }

extern "C" void dBodySetMass(dBodyID b, const dMass *m) {
    if (!(m->mass > 0)) fatal("dBodySetMass: mass must be > 0");
    if (m->c[0] != 0 || m->c[1] != 0 || m->c[2] != 0) fatal("dBodySetMass: centre of mass must be at the origin");
    HostBodies &hb = HB(b);
    hb.lvel[4 * b->idx + 3] = m->mass;
    memcpy(&hb.I[12 * b->idx], m->I, 12 * sizeof(float));
    if (!(hb.flags[b->idx] & BF_KINEMATIC)) {
        hb.pos[4 * b->idx + 3] = 1.0f / m->mass;
        invert3(m->I, &hb.invI[12 * b->idx]);
    }
    eng_mark_body_fields(b->w->eng, b->idx, FLD_MASS);
}
extern "C" void dBodyGetMass(dBodyID b, dMass *m) {
    HostBodies &hb = HB(b);
    dMassSetZero(m);
    m->mass = hb.lvel[4 * b->idx + 3];
    memcpy(m->I, &hb.I[12 * b->idx], 12 * sizeof(float));
}
extern "C" void dBodySetKinematic(dBodyID b) {
    HostBodies &hb = HB(b);
    hb.flags[b->idx] |= BF_KINEMATIC;
    hb.pos[4 * b->idx + 3] = 0.f; // invMass = 0, invI = 0
    for (int k = 0; k < 12; k++) hb.invI[12 * b->idx + k] = 0;
    eng_mark_body_fields(b->w->eng, b->idx, FLD_MASS);
}
extern "C" void dBodySetDynamic(dBodyID b) {
    HostBodies &hb = HB(b);
    hb.flags[b->idx] &= ~BF_KINEMATIC;
    hb.pos[4 * b->idx + 3] = 1.0f / hb.lvel[4 * b->idx + 3];
    invert3(&hb.I[12 * b->idx], &hb.invI[12 * b->idx]);
    eng_mark_body_fields(b->w->eng, b->idx, FLD_MASS);
}
extern "C" int dBodyIsKinematic(dBodyID b) { return (HB(b).flags[b->idx] & BF_KINEMATIC) != 0; }
static void set_flag(dBodyID b, int flag, bool on) {
    int &f = HB(b).flags[b->idx];
    f = on ? (f | flag) : (f & ~flag);
    eng_mark_body_fields(b->w->eng, b->idx, FLD_MASS);
}
extern "C" void dBodySetGravityMode(dBodyID b, int mode) { set_flag(b, BF_NOGRAVITY, mode == 0); }
extern "C" int dBodyGetGravityMode(dBodyID b) { return (HB(b).flags[b->idx] & BF_NOGRAVITY) == 0; }
extern "C" void dBodySetGyroscopicMode(dBodyID b, int on) { set_flag(b, BF_GYRO, on != 0); }
extern "C" int dBodyGetGyroscopicMode(dBodyID b) { return (HB(b).flags[b->idx] & BF_GYRO) != 0; }
extern "C" void dBodyAddForce(dBodyID b, dReal fx, dReal fy, dReal fz) {
    float *f = &HB(b).facc[4 * b->idx];
    f[0] += fx; f[1] += fy; f[2] += fz;
    eng_mark_body_fields(b->w->eng, b->idx, FLD_FORCE);
}
extern "C" void dBodyAddTorque(dBodyID b, dReal fx, dReal fy, dReal fz) {
    float *f = &HB(b).tacc[4 * b->idx];
    f[0] += fx; f[1] += fy; f[2] += fz;
    eng_mark_body_fields(b->w->eng, b->idx, FLD_FORCE);
}
extern "C" void dBodySetData(dBodyID b, void *d) { b->data = d; }
extern "C" void *dBodyGetData(dBodyID b) { return b->data; }
extern "C" dWorldID dBodyGetWorld(dBodyID b) { return b->w; }
extern "C" void dBodySetEnvB200(dBodyID b, int env) {
    if (HB(b).env[b->idx] == env) return;
    fresh(b->w); // env membership shapes the per-env tables: full re-upload
    HB(b).env[b->idx] = env;
    eng_mark_bodies_dirty(b->w->eng);
}
extern "C" int dBodyGetIndexB200(dBodyID b) { return b->idx; }
extern "C" dBodyID dWorldGetBodyB200(dWorldID w, int i) { return (i >= 0 && i < (int)w->bodies.size()) ? &w->bodies[i] : nullptr; }

// ------------------------------------------------------------------ spaces and geoms

extern "C" dSpaceID dHashSpaceCreate(dSpaceID) {
    dxSpace *s = new dxSpace();
    s->w = current_world();
    if (s->w) s->w->spaces.push_back(s);
    return s;
}
extern "C" dSpaceID dSimpleSpaceCreate(dSpaceID p) { return dHashSpaceCreate(p); }
extern "C" void dSpaceDestroy(dSpaceID s) {
    if (!s) return;
    if (s->w) {
        for (dxGeom *g : s->geoms)
            if (g->alive) dGeomDestroy(g);
        for (auto &p : s->w->spaces) if (p == s) p = nullptr;
    }
    delete s;
}
extern "C" int dSpaceGetNumGeoms(dSpaceID s) {
    int n = 0;
    for (dxGeom *g : s->geoms) n += g->alive ? 1 : 0;
    return n;
}
extern "C" dGeomID dSpaceGetGeom(dSpaceID s, int i) {
    for (dxGeom *g : s->geoms)
        if (g->alive && i-- == 0) return g;
    return nullptr;
}
extern "C" dGeomID dSpaceGetGeomB200(dSpaceID s, int i) { return (i >= 0 && i < (int)s->geoms.size()) ? s->geoms[i] : nullptr; }

#define HG(g) eng_geoms((g)->space->w->eng)

static dxGeom *new_geom(dxSpace *s, int type, float d0, float d1, float d2, float d3) {
    if (!s) fatal("geoms must be created in a space (dSpaceID 0 is not supported)");
    dxWorld *w = space_world(s);
    if (w->reuse_slots && !w->free_geoms.empty() && (type == G_SPHERE || type == G_BOX)) {
        const int idx = pop_lowest(w->free_geoms);
        eng_reset_geom(w->eng, idx);
        HostGeoms &hg = eng_geoms(w->eng);
        hg.type[idx] = type;
        hg.dims[4 * idx] = d0; hg.dims[4 * idx + 1] = d1; hg.dims[4 * idx + 2] = d2; hg.dims[4 * idx + 3] = d3;
        dxGeom *g = &w->geoms[idx];
        if (g->space != s) {
            std::vector<dxGeom *> &v = g->space->geoms;
            v.erase(std::remove(v.begin(), v.end(), g), v.end());
            s->geoms.push_back(g);
        }
        *g = dxGeom{s, idx, nullptr, nullptr, true, {0, 0, 0, 0, 0, 0}};
        return g;
    }
    const int idx = eng_add_geom(w->eng);
    HostGeoms &hg = eng_geoms(w->eng);
    hg.type[idx] = type;
    hg.dims[4 * idx] = d0; hg.dims[4 * idx + 1] = d1; hg.dims[4 * idx + 2] = d2; hg.dims[4 * idx + 3] = d3;
    hg.env[idx] = -1;
    w->geoms.push_back(dxGeom{s, idx, nullptr, nullptr, true, {0, 0, 0, 0, 0, 0}});
    dxGeom *g = &w->geoms.back();
    s->geoms.push_back(g);
    return g;
}

extern "C" dGeomID dCreateSphere(dSpaceID s, dReal r) { return new_geom(s, G_SPHERE, r, 0, 0, 0); }
extern "C" dGeomID dCreateBox(dSpaceID s, dReal lx, dReal ly, dReal lz) { return new_geom(s, G_BOX, lx, ly, lz, 0); }
static void plane_normalize(float *p) {
    // make_sure_plane_normal_has_unit_length
    float l = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
    if (l > 0) {
        l = 1.0f / sqrtf(l);
        p[0] *= l; p[1] *= l; p[2] *= l; p[3] *= l;
    } else {
        p[0] = 1; p[1] = 0; p[2] = 0; p[3] = 0;
    }
}
extern "C" dGeomID dCreatePlane(dSpaceID s, dReal a, dReal b, dReal c, dReal d) {
    float p[4] = {a, b, c, d};
    plane_normalize(p);
    return new_geom(s, G_PLANE, p[0], p[1], p[2], p[3]);
}
extern "C" void dGeomDestroy(dGeomID g) {
    if (!g || !g->alive) return;
    if (g->body) {
        std::vector<dxGeom *> &v = g->body->geoms;
        v.erase(std::remove(v.begin(), v.end(), g), v.end());
    }
    HG(g).alive[g->idx] = 0;
    eng_mark_geom(g->space->w->eng, g->idx);
    g->alive = false;
    dxWorld *w = g->space->w;
    const int t = HG(g).type[g->idx];
    // (slots of env 0 only: a re-used slot keeps its env, and env 0 is where the handle API creates things)
    if (w->reuse_slots && (t == G_SPHERE || t == G_BOX) && HG(g).env[g->idx] == 0) w->free_geoms.push_back(g->idx);
}
extern "C" void dGeomSetBody(dGeomID g, dBodyID b) {
    if (b && b->w != g->space->w) fatal("dGeomSetBody: the body belongs to a different world than the geom's space");
    if (g->body && g->body != b) {
        if (!b) { // detached: the geom keeps the pose it had on the body (ODE: dGeomSetBody(g, 0) leaves the geom in place)
            const float *bp = dBodyGetPosition(g->body), *bR = dBodyGetRotation(g->body);
            memcpy(&HG(g).pos[4 * g->idx], bp, 3 * sizeof(float));
            memcpy(&HG(g).R[12 * g->idx], bR, 12 * sizeof(float));
        }
        std::vector<dxGeom *> &v = g->body->geoms;
        v.erase(std::remove(v.begin(), v.end(), g), v.end());
    }
    if (b && g->body != b) b->geoms.push_back(g);
    g->body = b;
    HostGeoms &hg = HG(g);
    hg.body[g->idx] = b ? b->idx : -1;
    if (b && hg.env[g->idx] < 0) hg.env[g->idx] = eng_bodies(b->w->eng).env[b->idx];
    eng_mark_geom(g->space->w->eng, g->idx);
}
extern "C" dBodyID dGeomGetBody(dGeomID g) { return g->body; }
extern "C" void dGeomSetPosition(dGeomID g, dReal x, dReal y, dReal z) {
    if (g->body) { dBodySetPosition(g->body, x, y, z); return; }
    float *p = &HG(g).pos[4 * g->idx];
    p[0] = x; p[1] = y; p[2] = z;
    eng_mark_geom(g->space->w->eng, g->idx);
}
extern "C" void dGeomSetRotation(dGeomID g, const dMatrix3 R) {
    if (g->body) { dBodySetRotation(g->body, R); return; }
    float *r = &HG(g).R[12 * g->idx];
    memcpy(r, R, 12 * sizeof(float));
    r[3] = r[7] = r[11] = 0;
    eng_mark_geom(g->space->w->eng, g->idx);
}
extern "C" void dGeomSetQuaternion(dGeomID g, const dQuaternion q) {
    if (g->body) { dBodySetQuaternion(g->body, q); return; }
    dQuaternion qq = {q[0], q[1], q[2], q[3]};
    normalize4(qq);
    dQtoR(qq, &HG(g).R[12 * g->idx]);
    eng_mark_geom(g->space->w->eng, g->idx);
}
extern "C" const dReal *dGeomGetPosition(dGeomID g) {
    if (g->body) return dBodyGetPosition(g->body);
    return &HG(g).pos[4 * g->idx];
}
extern "C" const dReal *dGeomGetRotation(dGeomID g) {
    if (g->body) return dBodyGetRotation(g->body);
    return &HG(g).R[12 * g->idx];
}
extern "C" void dGeomGetQuaternion(dGeomID g, dQuaternion q) {
    if (g->body) { memcpy(q, dBodyGetQuaternion(g->body), 4 * sizeof(float)); return; }
    dRtoQ(&HG(g).R[12 * g->idx], q);
}
extern "C" int dGeomGetClass(dGeomID g) { return HG(g).type[g->idx]; }
extern "C" void dGeomSetCategoryBits(dGeomID g, unsigned long bits) {
    HG(g).cat[g->idx] = (uint32_t)bits;
    eng_mark_geom(g->space->w->eng, g->idx);
}
extern "C" void dGeomSetCollideBits(dGeomID g, unsigned long bits) {
    HG(g).col[g->idx] = (uint32_t)bits;
    eng_mark_geom(g->space->w->eng, g->idx);
}
extern "C" unsigned long dGeomGetCategoryBits(dGeomID g) { return HG(g).cat[g->idx]; }
extern "C" unsigned long dGeomGetCollideBits(dGeomID g) { return HG(g).col[g->idx]; }
extern "C" void dGeomSetData(dGeomID g, void *d) { g->data = d; }
extern "C" void *dGeomGetData(dGeomID g) { return g->data; }
extern "C" dSpaceID dGeomGetSpace(dGeomID g) { return g->space; }
extern "C" dReal dGeomSphereGetRadius(dGeomID g) { return HG(g).dims[4 * g->idx]; }
extern "C" void dGeomSphereSetRadius(dGeomID g, dReal r) {
    HG(g).dims[4 * g->idx] = r;
    eng_mark_geom(g->space->w->eng, g->idx);
}
extern "C" void dGeomBoxGetLengths(dGeomID g, dVector3 out) {
    for (int k = 0; k < 3; k++) out[k] = HG(g).dims[4 * g->idx + k];
    out[3] = 0;
}
extern "C" void dGeomBoxSetLengths(dGeomID g, dReal lx, dReal ly, dReal lz) {
    float *d = &HG(g).dims[4 * g->idx];
    d[0] = lx; d[1] = ly; d[2] = lz;
    eng_mark_geom(g->space->w->eng, g->idx);
}
extern "C" void dGeomPlaneGetParams(dGeomID g, dVector4 out) {
    for (int k = 0; k < 4; k++) out[k] = HG(g).dims[4 * g->idx + k];
}
extern "C" void dGeomPlaneSetParams(dGeomID g, dReal a, dReal b, dReal c, dReal d) {
    float p[4] = {a, b, c, d};
    plane_normalize(p);
    memcpy(&HG(g).dims[4 * g->idx], p, sizeof(p));
    eng_mark_geom(g->space->w->eng, g->idx);
}
extern "C" void dGeomSetEnvB200(dGeomID g, int env) {
    HG(g).env[g->idx] = env;
    eng_mark_geom(g->space->w->eng, g->idx);
}
extern "C" int dGeomGetIndexB200(dGeomID g) { return g->idx; }

extern "C" void dGeomGetAABB(dGeomID g, dReal aabb[6]) {
    // host restatement of the AABB rules for a single geom (debug accessor; not on the hot path)
    const HostGeoms &hg = HG(g);
    const float *pos = dGeomGetPosition(g), *R = dGeomGetRotation(g), *d = &hg.dims[4 * g->idx];
    const int type = hg.type[g->idx];
    for (int k = 0; k < 3; k++) { aabb[2 * k] = -INFINITY; aabb[2 * k + 1] = INFINITY; }
    if (type == G_SPHERE) {
        for (int k = 0; k < 3; k++) { aabb[2 * k] = pos[k] - d[0]; aabb[2 * k + 1] = pos[k] + d[0]; }
    } else if (type == G_BOX) {
        for (int k = 0; k < 3; k++) {
            float range = 0.5f * (fabsf(R[k * 4] * d[0]) + fabsf(R[k * 4 + 1] * d[1]) + fabsf(R[k * 4 + 2] * d[2]));
            aabb[2 * k] = pos[k] - range;
            aabb[2 * k + 1] = pos[k] + range;
        }
    } else if (type == G_PLANE) {
        if (d[1] == 0.0f && d[2] == 0.0f) { aabb[0] = (d[0] > 0) ? -INFINITY : -d[3]; aabb[1] = (d[0] > 0) ? d[3] : INFINITY; }
        else if (d[0] == 0.0f && d[2] == 0.0f) { aabb[2] = (d[1] > 0) ? -INFINITY : -d[3]; aabb[3] = (d[1] > 0) ? d[3] : INFINITY; }
        else if (d[0] == 0.0f && d[1] == 0.0f) { aabb[4] = (d[2] > 0) ? -INFINITY : -d[3]; aabb[5] = (d[2] > 0) ? d[3] : INFINITY; }
    }
}

// trimesh
extern "C" dTriMeshDataID dGeomTriMeshDataCreate(void) { return new dxTriMeshData(); }
extern "C" void dGeomTriMeshDataDestroy(dTriMeshDataID d) { delete d; }
extern "C" void dGeomTriMeshDataBuildSingle(dTriMeshDataID d, const void *Vertices, int VertexStride, int VertexCount,
                                            const void *Indices, int IndexCount, int TriStride) {
    d->verts.resize((size_t)VertexCount * 3);
    for (int i = 0; i < VertexCount; i++) {
        const float *v = (const float *)((const char *)Vertices + (size_t)i * VertexStride);
        d->verts[3 * i] = v[0]; d->verts[3 * i + 1] = v[1]; d->verts[3 * i + 2] = v[2];
    }
    const int nt = IndexCount / 3;
    d->tris.resize((size_t)nt * 3);
    for (int t = 0; t < nt; t++) {
        const dTriIndex *ix = (const dTriIndex *)((const char *)Indices + (size_t)t * TriStride);
        d->tris[3 * t] = (int)ix[0]; d->tris[3 * t + 1] = (int)ix[1]; d->tris[3 * t + 2] = (int)ix[2];
    }
    d->bound_world = nullptr;
    d->mesh_id = -1;
}
// Wavefront OBJ -> trimesh data (res/teapot.obj, res/grassPlane.obj).  Reads `v x y z`
// and `f` records (`a`, `a/b`, `a//c`, `a/b/c`; 1-based, negative = relative to the vertices read so far);
// polygons are fan-triangulated; everything else is skipped.  Returns the triangle count, -1 if the file cannot
// be read or holds no triangle.
extern "C" int dGeomTriMeshDataBuildFromOBJB200(dTriMeshDataID d, const char *path) {
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    d->verts.clear(); d->tris.clear();
    char line[4096];
    while (fgets(line, sizeof(line), f)) {
        const char *p = line;
        while (*p == ' ' || *p == '\t') p++;
        if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t')) {
            float x, y, z;
            if (sscanf(p + 1, "%f %f %f", &x, &y, &z) == 3) { d->verts.push_back(x); d->verts.push_back(y); d->verts.push_back(z); }
        } else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t')) {
            const int nv = (int)d->verts.size() / 3;
            int idx[64], n = 0;
            p += 1;
            while (n < 64) {
                while (*p == ' ' || *p == '\t') p++;
                if (*p == 0 || *p == '\n' || *p == '\r' || *p == '#') break;
                char *end;
                long v = strtol(p, &end, 10);
                if (end == p) break;
                if (v < 0) v = nv + v + 1;
                idx[n++] = (int)v - 1;
                p = end;
                while (*p && *p != ' ' && *p != '\t' && *p != '\n' && *p != '\r') p++; // skip /vt/vn
            }
            bool ok = n >= 3;
            for (int k = 0; k < n; k++) ok = ok && idx[k] >= 0 && idx[k] < nv;
            if (ok)
                for (int k = 1; k + 1 < n; k++) { d->tris.push_back(idx[0]); d->tris.push_back(idx[k]); d->tris.push_back(idx[k + 1]); }
        }
    }
    fclose(f);
    d->bound_world = nullptr;
    d->mesh_id = -1;
    const int nt = (int)d->tris.size() / 3;
    return nt > 0 ? nt : -1;
}
extern "C" int dGeomTriMeshDataGetB200(dTriMeshDataID d, float *verts3, int cap_verts, int *tris3, int cap_tris, int *n_verts) {
    const int nv = (int)d->verts.size() / 3, nt = (int)d->tris.size() / 3;
    if (n_verts) *n_verts = nv;
    if (verts3 && cap_verts > 0) memcpy(verts3, d->verts.data(), sizeof(float) * 3 * (size_t)std::min(nv, cap_verts));
    if (tris3 && cap_tris > 0) memcpy(tris3, d->tris.data(), sizeof(int) * 3 * (size_t)std::min(nt, cap_tris));
    return nt;
}
extern "C" dGeomID dCreateTriMesh(dSpaceID s, dTriMeshDataID d, dTriCallback *, dTriArrayCallback *, dTriRayCallback *) {
    dxWorld *w = space_world(s);
    if (d->bound_world != w || d->mesh_id < 0) {
        d->mesh_id = eng_add_mesh(w->eng, d->verts.data(), (int)d->verts.size() / 3, d->tris.data(), (int)d->tris.size() / 3);
        d->bound_world = w;
    }
    return new_geom(s, G_TRIMESH, (float)d->mesh_id, 0, 0, 0);
}
extern "C" int dWorldAddTriMeshB200(dWorldID w, const float *verts, int nv, const int *tris, int nt) {
    return eng_add_mesh(w->eng, verts, nv, tris, nt);
}

// ------------------------------------------------------------------ bulk creation / state

extern "C" int dWorldAddBodiesB200(dWorldID w, int n, const float *pos3, const float *quat4, const float *lvel3,
                                   const float *avel3, const float *mass, const float *inertia9, const int *flags,
                                   const int *env) {
    HostBodies &hb = eng_bodies(w->eng);
    const int first = hb.n;
    for (int i = 0; i < n; i++) {
        dxBody *b = new_body(w);
        const int k = b->idx;
        for (int c = 0; c < 3; c++) hb.pos[4 * k + c] = pos3[3 * i + c];
        if (quat4) {
            float *q = &hb.quat[4 * k];
            memcpy(q, quat4 + 4 * i, 16);
            normalize4(q);
            dQtoR(q, &hb.R[12 * k]);
        }
        if (lvel3) for (int c = 0; c < 3; c++) hb.lvel[4 * k + c] = lvel3[3 * i + c];
        if (avel3) for (int c = 0; c < 3; c++) hb.avel[4 * k + c] = avel3[3 * i + c];
        const int fl = flags ? flags[i] : 0;
        hb.flags[k] = fl;
        hb.env[k] = env ? env[i] : 0;
        const float m = mass ? mass[i] : 1.0f;
        hb.lvel[4 * k + 3] = m;
        if (inertia9)
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) hb.I[12 * k + 4 * r + c] = inertia9[9 * i + 3 * r + c];
        if (fl & BF_KINEMATIC) {
            hb.pos[4 * k + 3] = 0.f;
            for (int c = 0; c < 12; c++) hb.invI[12 * k + c] = 0.f;
        } else {
            hb.pos[4 * k + 3] = 1.0f / m;
            if (inertia9) invert3(&hb.I[12 * k], &hb.invI[12 * k]);
        }
    }
    return first; // eng_add_body queued them (or asked for the first full upload)
}

extern "C" int dSpaceAddGeomsB200(dSpaceID s, dWorldID w, int n, const int *type, const float *dims4, const int *body,
                                  const float *pos3, const float *R12, const unsigned *cat, const unsigned *col,
                                  const int *env) {
    if (!s->w) { s->w = w; w->spaces.push_back(s); }
    if (s->w != w) fatal("dSpaceAddGeomsB200: the space is bound to another world");
    HostGeoms &hg = eng_geoms(w->eng);
    const int first = hg.n;
    for (int i = 0; i < n; i++) {
        float d[4] = {dims4[4 * i], dims4[4 * i + 1], dims4[4 * i + 2], dims4[4 * i + 3]};
        if (type[i] == G_PLANE) plane_normalize(d);
        dxGeom *g = new_geom(s, type[i], d[0], d[1], d[2], d[3]);
        const int k = g->idx;
        const int b = body ? body[i] : -1;
        if (b >= 0) {
            if (b >= (int)w->bodies.size()) fatal("dSpaceAddGeomsB200: body index out of range");
            g->body = &w->bodies[b];
            g->body->geoms.push_back(g);
            hg.body[k] = b;
        } else {
            if (pos3) for (int c = 0; c < 3; c++) hg.pos[4 * k + c] = pos3[3 * i + c];
            if (R12) {
                memcpy(&hg.R[12 * k], R12 + 12 * i, 48);
                hg.R[12 * k + 3] = hg.R[12 * k + 7] = hg.R[12 * k + 11] = 0;
            }
        }
        if (cat) hg.cat[k] = cat[i];
        if (col) hg.col[k] = col[i];
        hg.env[k] = env ? env[i] : (b >= 0 ? eng_bodies(w->eng).env[b] : -1);
    }
    return first;
}

extern "C" void dWorldGetStateB200(dWorldID w, float *pos3, float *quat4, float *lvel3, float *avel3, float *R12) {
    fresh(w);
    const HostBodies &hb = eng_bodies(w->eng);
    for (int i = 0; i < hb.n; i++) {
        if (pos3) for (int c = 0; c < 3; c++) pos3[3 * i + c] = hb.pos[4 * i + c];
        if (quat4) memcpy(quat4 + 4 * i, &hb.quat[4 * i], 16);
        if (lvel3) for (int c = 0; c < 3; c++) lvel3[3 * i + c] = hb.lvel[4 * i + c];
        if (avel3) for (int c = 0; c < 3; c++) avel3[3 * i + c] = hb.avel[4 * i + c];
        if (R12) memcpy(R12 + 12 * i, &hb.R[12 * i], 48);
    }
}

extern "C" void dWorldSetForcesB200(dWorldID w, const float *f6, int n) { eng_set_forces(w->eng, f6, n); }
extern "C" void dWorldGetSnapshotB200(dWorldID w, float *dst, int first, int count, int blocking) {
    eng_snapshot_to_host(w->eng, dst, first, count, blocking != 0);
}
extern "C" const float *dWorldGetSnapshotDeviceB200(dWorldID w) { return eng_snapshot_device(w->eng); }
extern "C" void dWorldSetSnapshotFormatB200(dWorldID w, int fmt) { eng_set_snapshot_format(w->eng, fmt); }
extern "C" int dWorldGetSnapshotFormatB200(dWorldID w) { return eng_snapshot_format(w->eng); }

// Host-side expansion of compact snapshot records into the reference's GetTransformMat layout (src/main.c:602-622).
// Format 2 rebuilds R with dQtoR in the arithmetic the solver tail uses (dmath.cuh q_to_r, no FMA contraction in
// either build), so the result equals the format-0 snapshot bit for bit (tests/test_api_gpu.py).
static void expand_range(const float *src, int fmt, long i0, long i1, float *dst) {
    if (fmt == 1) {
        for (long i = i0; i < i1; i++) {
            const float *s = src + 12 * i;
            float *d = dst + 16 * i;
            d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = 0.f;
            d[4] = s[3]; d[5] = s[4]; d[6] = s[5]; d[7] = 0.f;
            d[8] = s[6]; d[9] = s[7]; d[10] = s[8]; d[11] = 0.f;
            d[12] = s[9]; d[13] = s[10]; d[14] = s[11]; d[15] = 1.f;
        }
    } else if (fmt == 2) {
        for (long i = i0; i < i1; i++) {
            const float *s = src + 8 * i;
            float *d = dst + 16 * i;
            const ob::M3 R = ob::q_to_r(make_float4(s[4], s[5], s[6], s[7]));
            d[0] = R.r0.x; d[1] = R.r1.x; d[2] = R.r2.x; d[3] = 0.f;
            d[4] = R.r0.y; d[5] = R.r1.y; d[6] = R.r2.y; d[7] = 0.f;
            d[8] = R.r0.z; d[9] = R.r1.z; d[10] = R.r2.z; d[11] = 0.f;
            d[12] = s[0]; d[13] = s[1]; d[14] = s[2]; d[15] = 1.f;
        }
    } else {
        memcpy(dst + 16 * i0, src + 16 * i0, (size_t)(i1 - i0) * 64);
    }
}
extern "C" void dSnapshotExpandB200(const float *src, int fmt, int count, float *dst16, int threads) {
    if (count <= 0) return;
    if (threads <= 1 || count < 4096) { expand_range(src, fmt, 0, count, dst16); return; }
    if (threads > 64) threads = 64;
    std::vector<std::thread> pool;
    const long chunk = ((long)count + threads - 1) / threads;
    for (int t = 1; t < threads; t++) {
        const long a = t * chunk, b = std::min<long>((long)count, a + chunk);
        if (a < b) pool.emplace_back(expand_range, src, fmt, a, b, dst16);
    }
    expand_range(src, fmt, 0, std::min<long>(count, chunk), dst16);
    for (auto &th : pool) th.join();
}

// ------------------------------------------------------------------ collision

extern "C" void dSpaceCollideDeviceB200(dSpaceID s, int max_contacts) {
    dxWorld *w = space_world(s);
    eng_collide(w->eng, max_contacts);
    w->device_contacts_pending = true;
}

extern "C" void dSpaceCollide(dSpaceID s, void *data, dNearCallback *callback) {
    dxWorld *w = space_world(s);
    eng_collide(w->eng, w->max_contacts);
    w->device_contacts_pending = false;
    w->cb_pairs = eng_fetch_pairs(w->eng);
    const HostPairs &hp = w->cb_pairs;
    w->in_callback = true;
    for (int p = 0; p < hp.n_pairs; p++) {
        w->cb_g1 = hp.g1[p]; w->cb_g2 = hp.g2[p];
        w->cb_first = hp.first[p]; w->cb_count = hp.count[p];
        callback(data, &w->geoms[hp.g1[p]], &w->geoms[hp.g2[p]]);
    }
    w->in_callback = false;
    w->cb_g1 = w->cb_g2 = -1;
}

static int copy_contacts(dxWorld *w, int first, int count, bool swapped, dGeomID o1, dGeomID o2, int maxc,
                         dContactGeom *contact, int skip) {
    const HostPairs &hp = w->cb_pairs;
    const int n = count < maxc ? count : maxc;
    for (int k = 0; k < n; k++) {
        dContactGeom *c = (dContactGeom *)((char *)contact + (size_t)k * skip);
        const float *pd = hp.pos_depth + 4 * (size_t)(first + k), *ns = hp.normal_side + 4 * (size_t)(first + k);
        c->pos[0] = pd[0]; c->pos[1] = pd[1]; c->pos[2] = pd[2]; c->pos[3] = 0;
        const float sg = swapped ? -1.f : 1.f;
        c->normal[0] = sg * ns[0]; c->normal[1] = sg * ns[1]; c->normal[2] = sg * ns[2]; c->normal[3] = 0;
        c->depth = pd[3];
        c->g1 = o1; c->g2 = o2;
        int side;
        memcpy(&side, &ns[3], sizeof(int));
        c->side1 = swapped ? side : -1;
        c->side2 = swapped ? -1 : side;
    }
    return n;
}

extern "C" int dCollide(dGeomID o1, dGeomID o2, int flags, dContactGeom *contact, int skip) {
    const int maxc = flags & 0xffff;
    if (!o1 || !o2 || o1 == o2 || maxc < 1) return 0;
    dxWorld *w = o1->space->w;
    if (o2->space->w != w) return 0;
    if (w->in_callback && ((o1->idx == w->cb_g1 && o2->idx == w->cb_g2) || (o1->idx == w->cb_g2 && o2->idx == w->cb_g1))) {
        const bool swapped = o1->idx != w->cb_g1;
        if (w->cb_count <= maxc || w->cb_count <= 1)
            return copy_contacts(w, w->cb_first, w->cb_count, swapped, o1, o2, maxc, contact, skip);
        // fewer contacts requested than were precomputed: ODE culls differently, so the exact
        // answer needs the pair re-collided with that limit (set dWorldSetMaxContactsB200 to avoid)
        fatal("dCollide: max contacts smaller than the world's precomputed limit; call dWorldSetMaxContactsB200 first");
    }
    if (w->in_callback) fatal("dCollide inside the callback is served for the callback's own pair only");
    // anywhere else: the pair goes through the narrowphase kernels on its own (a GPU round trip per call)
    HostPairs hp = eng_collide_pair(w->eng, o1->idx, o2->idx, maxc);
    if (hp.n_pairs != 1 || hp.count[0] == 0) return 0;
    const HostPairs saved = w->cb_pairs;
    w->cb_pairs = hp;
    const int n = copy_contacts(w, hp.first[0], hp.count[0], hp.g1[0] != o1->idx, o1, o2, maxc, contact, skip);
    w->cb_pairs = saved;
    return n;
}

extern "C" int dSpaceGetPairsB200(dSpaceID s, int *pairs2, int cap) {
    dxWorld *w = space_world(s);
    HostPairs hp = eng_fetch_pairs(w->eng);
    w->cb_pairs = hp;
    for (int p = 0; p < hp.n_pairs && p < cap; p++) { pairs2[2 * p] = hp.g1[p]; pairs2[2 * p + 1] = hp.g2[p]; }
    return hp.n_pairs;
}

extern "C" int dSpaceGetContactsB200(dSpaceID s, int *count_per_pair, int cap_pairs, float *pos_depth4,
                                     float *normal_side4, int cap_contacts) {
    dxWorld *w = space_world(s);
    HostPairs hp = eng_fetch_pairs(w->eng);
    w->cb_pairs = hp;
    int total = 0;
    for (int p = 0; p < hp.n_pairs; p++) {
        if (p < cap_pairs) count_per_pair[p] = hp.count[p];
        for (int k = 0; k < hp.count[p]; k++) {
            if (total < cap_contacts) {
                memcpy(pos_depth4 + 4 * (size_t)total, hp.pos_depth + 4 * (size_t)(hp.first[p] + k), 16);
                memcpy(normal_side4 + 4 * (size_t)total, hp.normal_side + 4 * (size_t)(hp.first[p] + k), 16);
            }
            total++;
        }
    }
    return total;
}

extern "C" int dWorldGetSolverOrderB200(dWorldID w, int *g1, int *g2, int *k, int cap) {
    return eng_export_solver_order(w->eng, g1, g2, k, cap);
}

// ------------------------------------------------------------------ contact joints

extern "C" dJointGroupID dJointGroupCreate(int) { return new dxJointGroup(); }
extern "C" void dJointGroupEmpty(dJointGroupID g) {
    if (g) g->joints.clear();
}
extern "C" void dJointGroupDestroy(dJointGroupID g) {
    if (!g) return;
    for (dxWorld *w : g_worlds)
        if (w)
            for (auto &p : w->groups) if (p == g) p = nullptr;
    for (dxWorld *w : g_worlds)
        if (w) {
            std::vector<dxJointGroup *> keep;
            for (auto p : w->groups) if (p) keep.push_back(p);
            w->groups.swap(keep);
        }
    delete g;
}

static dxJointGroup g_default_group;

extern "C" dJointID dJointCreateContact(dWorldID w, dJointGroupID g, const dContact *c) {
    if (!g) g = &g_default_group;
    bool known = false;
    for (dxJointGroup *p : w->groups) if (p == g) known = true;
    if (!known) w->groups.push_back(g);
    dxJoint j;
    j.w = w; j.b1 = j.b2 = nullptr;
    j.hc.pos[0] = c->geom.pos[0]; j.hc.pos[1] = c->geom.pos[1]; j.hc.pos[2] = c->geom.pos[2];
    j.hc.depth = c->geom.depth;
    j.hc.normal[0] = c->geom.normal[0]; j.hc.normal[1] = c->geom.normal[1]; j.hc.normal[2] = c->geom.normal[2];
    j.hc.b1 = j.hc.b2 = -1;
    // only the fields the mode selects are read: the reference leaves the rest uninitialised (src/main.c:676)
    const dSurfaceParameters &s = c->surface;
    Surface &o = j.hc.surf;
    memset(&o, 0, sizeof(o));
    o.mode = s.mode;
    o.mu = s.mu;
    if (s.mode & dContactMu2) o.mu2 = s.mu2;
    if (s.mode & dContactBounce) { o.bounce = s.bounce; o.bounce_vel = s.bounce_vel; }
    if (s.mode & dContactSoftERP) o.soft_erp = s.soft_erp;
    if (s.mode & dContactSoftCFM) o.soft_cfm = s.soft_cfm;
    if (s.mode & dContactMotion1) o.motion1 = s.motion1;
    if (s.mode & dContactMotion2) o.motion2 = s.motion2;
    if (s.mode & dContactMotionN) o.motionN = s.motionN;
    if (s.mode & dContactSlip1) o.slip1 = s.slip1;
    if (s.mode & dContactSlip2) o.slip2 = s.slip2;
    if (s.mode & dContactFDir1) fatal("dContactFDir1 is not supported (friction directions come from dPlaneSpace)");
    g->joints.push_back(j);
    return &g->joints.back();
}
extern "C" void dJointAttach(dJointID j, dBodyID b1, dBodyID b2) {
    j->b1 = b1;
    j->b2 = b2;
}
extern "C" dBodyID dJointGetBody(dJointID j, int index) { return index == 0 ? j->b1 : j->b2; }

// ------------------------------------------------------------------ primitive test hooks
// (host arrays in, host arrays out; used by tests/test_prims_gpu.py to check the hand-written scan
// and radix sort against numpy at awkward sizes)
#include "prims.cuh"

extern "C" int dTestScanB200(const int *in, int *out, long n, int *total) {
    OB_CUDA(cudaSetDevice(default_device()));
    int *d = nullptr, *dt = nullptr;
    OB_CUDA(ob_malloc(&d, (size_t)(n > 0 ? n : 1) * sizeof(int)));
    OB_CUDA(ob_malloc(&dt, sizeof(int)));
    OB_CUDA(cudaMemcpy(d, in, (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
    ScanWorkspace ws;
    scan_exclusive(d, d, n, nullptr, dt, ws, 0);
    OB_CUDA(cudaDeviceSynchronize());
    OB_CUDA(cudaMemcpy(out, d, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
    OB_CUDA(cudaMemcpy(total, dt, sizeof(int), cudaMemcpyDeviceToHost));
    scan_workspace_free(ws);
    ob_free(d);
    ob_free(dt);
    return 0;
}

extern "C" int dTestSortB200(unsigned *keys, int *vals, long n, int bits) {
    OB_CUDA(cudaSetDevice(default_device()));
    uint32_t *dk = nullptr;
    int *dv = nullptr;
    OB_CUDA(ob_malloc(&dk, (size_t)(n > 0 ? n : 1) * sizeof(uint32_t)));
    OB_CUDA(ob_malloc(&dv, (size_t)(n > 0 ? n : 1) * sizeof(int)));
    OB_CUDA(cudaMemcpy(dk, keys, (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice));
    OB_CUDA(cudaMemcpy(dv, vals, (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
    SortWorkspace ws;
    sort_pairs(dk, dv, n, nullptr, bits, ws, 0);
    OB_CUDA(cudaDeviceSynchronize());
    OB_CUDA(cudaMemcpy(keys, dk, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    OB_CUDA(cudaMemcpy(vals, dv, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
    sort_workspace_free(ws);
    ob_free(dk);
    ob_free(dv);
    return 0;
}
