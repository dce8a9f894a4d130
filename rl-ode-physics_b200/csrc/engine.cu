// engine.cu -- world lifetime, HBM allocation, host<->device synchronisation and the per-tick
// orchestration (collide -> step) of libode_b200.  The tick mirrors the reference's
// /root/reference/src/main.c:212-214: dSpaceCollide(+NearCallback), dWorldStep, dJointGroupEmpty.
#include <string.h>
#include <sys/mman.h>

#include <algorithm>
#include <cmath>

#include "solver_dev.cuh"

namespace ob {

static void step_begin(Engine *e);
static void step_end(Engine *e);
static void apply_pending_forces(Engine *e);

template <typename T>
static void dev_realloc(T *&p, size_t old_n, size_t new_n, cudaStream_t st, bool keep = true) {
    T *q = nullptr;
    OB_CUDA(ob_malloc(&q, std::max<size_t>(new_n, 1) * sizeof(T)));
    if (p && keep && old_n) OB_CUDA(cudaMemcpyAsync(q, p, std::min(old_n, new_n) * sizeof(T), cudaMemcpyDeviceToDevice, st));
    if (p) {
        OB_CUDA(cudaStreamSynchronize(st));
        OB_CUDA(ob_free(p));
    }
    p = q;
}
template <typename T>
static void dev_free(T *&p) {
    if (p) ob_free(p);
    p = nullptr;
}

// address-space reservations of the StableVec host mirrors (engine.h)
void *stable_reserve(size_t *bytes) {
    // under an address-space limit (ulimit -v) the full reservation may be refused: take what can be had, down to 64 MiB
    for (size_t want = *bytes; want >= ((size_t)64 << 20); want >>= 1) {
        void *p = mmap(nullptr, want, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p != MAP_FAILED) {
            *bytes = want;
            return p;
        }
    }
    perror("libode_b200: mmap (host mirror reservation)");
    abort();
}
void stable_commit(void *base, size_t old_bytes, size_t new_bytes) {
    if (mprotect(static_cast<char *>(base) + old_bytes, new_bytes - old_bytes, PROT_READ | PROT_WRITE) != 0) {
        perror("libode_b200: mprotect (host mirror growth)");
        abort();
    }
}
void stable_release(void *base, size_t bytes) { munmap(base, bytes); }

Engine *eng_create(int device) {
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0) {
        fprintf(stderr, "libode_b200: no CUDA device available (%s); this library has no CPU path\n",
                cudaGetErrorString(err));
        abort();
    }
    if (device < 0 || device >= count) device = 0;
    OB_CUDA(cudaSetDevice(device));
    Engine *e = new Engine();
    e->device = device;
    cudaDeviceProp prop;
    OB_CUDA(cudaGetDeviceProperties(&prop, device));
    e->num_sms = prop.multiProcessorCount;
    if (!prop.cooperativeLaunch) {
        fprintf(stderr, "libode_b200: device lacks cooperative launch\n");
        abort();
    }
    OB_CUDA(cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking));
    OB_CUDA(cudaStreamCreateWithFlags(&e->copy_st, cudaStreamNonBlocking));
    OB_CUDA(cudaStreamCreateWithFlags(&e->h2d_st, cudaStreamNonBlocking));
    OB_CUDA(cudaEventCreateWithFlags(&e->ev_step_done, cudaEventDisableTiming));
    OB_CUDA(cudaEventCreateWithFlags(&e->ev_snap_copied[0], cudaEventDisableTiming));
    OB_CUDA(cudaEventCreateWithFlags(&e->ev_snap_copied[1], cudaEventDisableTiming));
    OB_CUDA(cudaEventCreateWithFlags(&e->ev_f6, cudaEventDisableTiming));
    OB_CUDA(cudaEventCreateWithFlags(&e->ev_f6_consumed[0], cudaEventDisableTiming));
    OB_CUDA(cudaEventCreateWithFlags(&e->ev_f6_consumed[1], cudaEventDisableTiming));
    OB_CUDA(ob_malloc(&e->d_stats, sizeof(StepStats)));
    OB_CUDA(cudaMemset(e->d_stats, 0, sizeof(StepStats)));
    OB_CUDA(cudaMallocHost(&e->h_stats, sizeof(StepStats)));
    memset(e->h_stats, 0, sizeof(StepStats));
    for (int i = 0; i < 5; i++) OB_CUDA(cudaEventCreate(&e->ev[i]));
    OB_CUDA(cudaEventCreate(&e->ev_bp));
    OB_CUDA(ob_malloc(&e->M.count, sizeof(int)));
    OB_CUDA(ob_malloc(&e->M.colour_start, 72 * sizeof(int)));
    OB_CUDA(ob_malloc(&e->M.meta, 16 * sizeof(int)));
    OB_CUDA(cudaMemset(e->M.count, 0, sizeof(int)));
    OB_CUDA(cudaMemset(e->M.meta, 0, 16 * sizeof(int)));
    OB_CUDA(ob_malloc(&e->bp.acc, 8 * sizeof(unsigned)));
    broadphase_acc_init(e->bp, e->st); // re-armed on the device after every use from here on
    OB_CUDA(ob_malloc(&e->bp.gp, sizeof(GridParams)));
    OB_CUDA(ob_malloc(&e->bp.counters, sizeof(BroadCounters)));
    OB_CUDA(cudaMemset(e->bp.counters, 0, sizeof(BroadCounters)));
    e->meshes.n = 0;
    if (const char *g = getenv("ODE_B200_ENV_GROUP")) e->env_group = atoi(g);
    if (const char *g = getenv("ODE_B200_COLOUR_SPREAD")) eng_set_colour_spread(e, atoi(g));
    if (const char *g = getenv("ODE_B200_CONTACT_UNITS")) e->contact_units = atoi(g);
    if (const char *g = getenv("ODE_B200_ENV_STAGE")) e->env_stage = atoi(g);
    if (const char *g = getenv("ODE_B200_ENV_SAP")) e->EB.sap = atoi(g);
    if (const char *g = getenv("ODE_B200_ENV_PAIR")) e->env_pair = atoi(g);
    if (const char *g = getenv("ODE_B200_ENV_PAIR_ROWS")) e->env_pair_rows = atoi(g);
    if (const char *g = getenv("ODE_B200_L2_PERSIST")) e->l2_persist = atoi(g);
    if (const char *g = getenv("ODE_B200_ENV_FUSE")) e->env_fuse = atoi(g);
    if (const char *g = getenv("ODE_B200_TINY_SOLVER")) e->tiny_solver = atoi(g);
    if (const char *g = getenv("ODE_B200_GRAPHS")) e->graphs = atoi(g);
    if (const char *g = getenv("ODE_B200_BROADPHASE")) e->broad_mode = !strcmp(g, "grid") ? 0 : !strcmp(g, "env") ? 1 : -1;
    return e;
}

void eng_destroy(Engine *e) {
    if (!e) return;
    cudaSetDevice(e->device);
    solver_profile_dump();
    cudaStreamSynchronize(e->st);
    cudaStreamSynchronize(e->copy_st);
    cudaStreamSynchronize(e->h2d_st);
    BodyArrays &B = e->B;
    dev_free(B.pos); dev_free(B.quat); dev_free(B.R); dev_free(B.lvel); dev_free(B.avel); dev_free(B.I);
    dev_free(B.invI); dev_free(B.facc); dev_free(B.tacc); dev_free(B.flags); dev_free(B.local); dev_free(B.env); dev_free(e->body_hot); dev_free(B.tmp);
    dev_free(e->snap_buf[0]); dev_free(e->snap_buf[1]); dev_free(B.colmask); dev_free(B.prio);
    GeomArrays &G = e->G;
    dev_free(G.type); dev_free(G.dims); dev_free(G.body); dev_free(G.pos); dev_free(G.R); dev_free(G.cat);
    dev_free(G.col); dev_free(G.env); dev_free(G.mesh); dev_free(G.alive); dev_free(G.amin); dev_free(G.amax);
    BroadPhase &bp = e->bp;
    dev_free(e->EB.first); dev_free(e->EB.count); dev_free(e->EB.shared);
    if (e->g_collide.exec) cudaGraphExecDestroy(e->g_collide.exec);
    for (int i = 0; i < 2; i++) if (e->g_step[i].exec) cudaGraphExecDestroy(e->g_step[i].exec);
    dev_free(e->sel_flag);
    if (e->h_patch) cudaFreeHost(e->h_patch);
    if (e->d_patch) ob_free(e->d_patch);
    if (e->ev_patch) cudaEventDestroy(e->ev_patch);
    dev_free(bp.blk);
    dev_free(bp.acc); dev_free(bp.gp); dev_free(bp.counters); dev_free(bp.keys); dev_free(bp.idx);
    dev_free(bp.s_min); dev_free(bp.s_max); dev_free(bp.s_flt); dev_free(bp.cell_start); dev_free(bp.cell_end);
    dev_free(bp.cnt); dev_free(bp.pairs); dev_free(bp.bb_list); dev_free(bp.sweep_tmp); dev_free(bp.sweep_tot);
    sort_workspace_free(bp.sort); scan_workspace_free(bp.scan);
    dev_free(e->cs.pd); dev_free(e->cs.ns); dev_free(e->cs.nc);
    ManifoldArrays &M = e->M;
    dev_free(M.rec); dev_free(M.colour); dev_free(M.skey); dev_free(M.sidx); dev_free(M.flag); dev_free(M.count);
    dev_free(M.colour_start); dev_free(M.meta);
    dev_free(e->ex_label); dev_free(e->ex_desc); dev_free(e->ex_isl_row0); dev_free(e->ex_isl_mat); dev_free(e->ex_isl_label);
    dev_free(e->ex_meta); dev_free(e->ex_A); dev_free(e->ex_C);
    SolverArrays &S = e->S;
    dev_free(S.q0); dev_free(S.q1); dev_free(S.q2); dev_free(S.q3); dev_free(S.q4); dev_free(S.q5); dev_free(S.lam); dev_free(S.mrec);
    sort_workspace_free(e->sort); scan_workspace_free(e->scan);
    dev_free(e->E.first_body); dev_free(e->E.n_body); dev_free(e->E.cnt); dev_free(e->E.start); dev_free(e->E.fill); dev_free(e->E.order); dev_free(e->E.rec); dev_free(e->E.perm); dev_free(e->E.col);
    dev_free(e->hc_pd); dev_free(e->hc_ns); dev_free(e->hc_surf); dev_free(e->hc_mrec);
    dev_free(e->dl_first); dev_free(e->dl_pd); dev_free(e->dl_ns);
    for (auto &m : e->hmeshes) { dev_free(m.d_verts); dev_free(m.d_tris); dev_free(m.d_cell_start); dev_free(m.d_cell_tris); }
    dev_free(e->d_stats);
    dev_free(e->msg_body); dev_free(e->msg_geom); dev_free(e->msg_type); dev_free(e->msg_size); dev_free(e->msg_col); dev_free(e->msg_out);
    dev_free(e->d_f6[0]); dev_free(e->d_f6[1]);
    if (e->tev[0]) { cudaEventDestroy(e->tev[0]); cudaEventDestroy(e->tev[1]); }
    if (e->h_stats) cudaFreeHost(e->h_stats);
    if (e->x_host) cudaFreeHost(e->x_host);
    if (e->hcs_host) cudaFreeHost(e->hcs_host);
    dev_free(e->hcs_dev);
    for (int i = 0; i < 5; i++) if (e->ev[i]) cudaEventDestroy(e->ev[i]);
    if (e->ev_bp) cudaEventDestroy(e->ev_bp);
    cudaStreamDestroy(e->st);
    cudaStreamDestroy(e->copy_st);
    cudaStreamDestroy(e->h2d_st);
    cudaEventDestroy(e->ev_step_done); cudaEventDestroy(e->ev_snap_copied[0]); cudaEventDestroy(e->ev_snap_copied[1]);
    cudaEventDestroy(e->ev_f6); cudaEventDestroy(e->ev_f6_consumed[0]); cudaEventDestroy(e->ev_f6_consumed[1]);
    delete e;
}

WorldParams &eng_params(Engine *e) { return e->params; }
int eng_device(Engine *e) { return e->device; }
cudaStream_t eng_stream(Engine *e) { return e->st; }
HostBodies &eng_bodies(Engine *e) { return e->hb; }
HostGeoms &eng_geoms(Engine *e) { return e->hg; }

int eng_add_body(Engine *e) {
    HostBodies &b = e->hb;
    const int i = b.n++;
    static const float ident[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    b.pos.insert(b.pos.end(), {0.f, 0.f, 0.f, 1.f});
    b.quat.insert(b.quat.end(), {1.f, 0.f, 0.f, 0.f});
    b.R.insert(b.R.end(), ident, ident + 12);
    b.lvel.insert(b.lvel.end(), {0.f, 0.f, 0.f, 1.f});
    b.avel.insert(b.avel.end(), {0.f, 0.f, 0.f, 0.f});
    b.I.insert(b.I.end(), ident, ident + 12);
    b.invI.insert(b.invI.end(), ident, ident + 12);
    b.facc.insert(b.facc.end(), {0.f, 0.f, 0.f, 0.f});
    b.tacc.insert(b.tacc.end(), {0.f, 0.f, 0.f, 0.f});
    b.flags.push_back(0);
    b.env.push_back(0);
    if (e->n_b_dev == 0) e->bodies_dirty = true; // nothing on the device yet: the first sync uploads everything
    else eng_mark_body_fields(e, i, FLD_ALL);
    return i;
}

int eng_add_geom(Engine *e) {
    HostGeoms &g = e->hg;
    const int i = g.n++;
    static const float ident[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    g.type.push_back(G_SPHERE);
    g.dims.insert(g.dims.end(), {0.f, 0.f, 0.f, 0.f});
    g.body.push_back(-1);
    g.pos.insert(g.pos.end(), {0.f, 0.f, 0.f, 0.f});
    g.R.insert(g.R.end(), ident, ident + 12);
    g.cat.push_back(0xffffffffu);
    g.col.push_back(0xffffffffu);
    g.env.push_back(-1);
    g.alive.push_back(1);
    if (e->n_g_dev == 0) e->geoms_dirty = true;
    else eng_mark_geom(e, i);
    return i;
}

// Slot re-use (dWorldSetSlotReuseB200): a destroyed body / geom slot is given back its creation defaults and sent to the
// device as one whole-record patch, exactly like a fresh append.  The env of the slot is kept.
void eng_reset_body(Engine *e, int i) {
    HostBodies &b = e->hb;
    static const float ident[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    const float p4[4] = {0.f, 0.f, 0.f, 1.f}, q4[4] = {1.f, 0.f, 0.f, 0.f}, z4[4] = {0.f, 0.f, 0.f, 0.f};
    memcpy(&b.pos[4 * (size_t)i], p4, sizeof(p4)); memcpy(&b.quat[4 * (size_t)i], q4, sizeof(q4));
    memcpy(&b.lvel[4 * (size_t)i], p4, sizeof(p4)); memcpy(&b.avel[4 * (size_t)i], z4, sizeof(z4));
    memcpy(&b.facc[4 * (size_t)i], z4, sizeof(z4)); memcpy(&b.tacc[4 * (size_t)i], z4, sizeof(z4));
    memcpy(&b.R[12 * (size_t)i], ident, sizeof(ident)); memcpy(&b.I[12 * (size_t)i], ident, sizeof(ident));
    memcpy(&b.invI[12 * (size_t)i], ident, sizeof(ident));
    b.flags[i] = 0;
    if (e->n_b_dev == 0) e->bodies_dirty = true;
    else eng_mark_body_fields(e, i, FLD_ALL);
}
void eng_reset_geom(Engine *e, int i) {
    HostGeoms &g = e->hg;
    static const float ident[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    const float z4[4] = {0.f, 0.f, 0.f, 0.f};
    g.type[i] = G_SPHERE;
    memcpy(&g.dims[4 * (size_t)i], z4, sizeof(z4)); memcpy(&g.pos[4 * (size_t)i], z4, sizeof(z4));
    memcpy(&g.R[12 * (size_t)i], ident, sizeof(ident));
    g.body[i] = -1; g.cat[i] = 0xffffffffu; g.col[i] = 0xffffffffu; g.alive[i] = 1;
    if (e->n_g_dev == 0) e->geoms_dirty = true;
    else eng_mark_geom(e, i);
}

int eng_add_mesh(Engine *e, const float *verts, int nv, const int *tris, int nt) {
    if ((int)e->hmeshes.size() >= MAX_MESHES) {
        fprintf(stderr, "libode_b200: at most %d trimesh data objects per world\n", MAX_MESHES);
        abort();
    }
    OB_CUDA(cudaSetDevice(e->device));
    TriMesh m;
    m.nv = nv; m.nt = nt;
    m.h_verts.assign(verts, verts + 3 * (size_t)nv);
    m.h_tris.assign(tris, tris + 3 * (size_t)nt);
    for (int k = 0; k < 3; k++) { m.lo[k] = INFINITY; m.hi[k] = -INFINITY; }
    for (int i = 0; i < nv; i++)
        for (int k = 0; k < 3; k++) {
            m.lo[k] = fminf(m.lo[k], verts[3 * i + k]);
            m.hi[k] = fmaxf(m.hi[k], verts[3 * i + k]);
        }
    const size_t bv = ((size_t)nv * 3 * sizeof(float) + 15) / 16 * 16, bt = ((size_t)nt * 3 * sizeof(int) + 15) / 16 * 16;
    OB_CUDA(ob_malloc(&m.d_verts, bv));
    OB_CUDA(ob_malloc(&m.d_tris, bt));
    OB_CUDA(cudaMemset(m.d_verts, 0, bv));
    OB_CUDA(cudaMemset(m.d_tris, 0, bt));
    OB_CUDA(cudaMemcpy(m.d_verts, verts, (size_t)nv * 3 * sizeof(float), cudaMemcpyHostToDevice));
    OB_CUDA(cudaMemcpy(m.d_tris, tris, (size_t)nt * 3 * sizeof(int), cudaMemcpyHostToDevice));
    const int id = (int)e->hmeshes.size();
    MeshInfo &mi = e->meshes.m[id];
    mi.verts = m.d_verts; mi.tris = m.d_tris; mi.nv = nv; mi.nt = nt;
    for (int k = 0; k < 3; k++) { mi.lo[k] = m.lo[k]; mi.hi[k] = m.hi[k]; }
    {
        // Triangle grid: ~2 cells per triangle, cell edges proportional to the bounds (a flat mesh gets a flat grid).
        // A collider visits only the cells its box touches instead of every triangle of the mesh.
        double ext[3], vol = 1.0;
        int flat = 0;
        for (int k = 0; k < 3; k++) { ext[k] = std::max((double)m.hi[k] - (double)m.lo[k], 0.0); if (ext[k] <= 0) flat++; }
        double emax = std::max(ext[0], std::max(ext[1], ext[2]));
        if (emax <= 0) emax = 1.0;
        for (int k = 0; k < 3; k++) vol *= std::max(ext[k], 1e-3 * emax);
        const double target = std::max(8.0, 2.0 * nt);
        const double cell = std::cbrt(vol / target);
        long total = 1;
        for (int k = 0; k < 3; k++) {
            int d = (int)std::floor(ext[k] / cell) + 1;
            d = std::min(std::max(d, 1), 256);
            mi.gd[k] = d;
            mi.gcell[k] = (float)(ext[k] > 0 ? ext[k] / d : 1.0);
            mi.ginv[k] = 1.0f / mi.gcell[k];
            total *= d;
        }
        auto cell_of = [&](float v, int k) {
            int c = (int)floorf((v - m.lo[k]) * mi.ginv[k]);
            return std::min(std::max(c, 0), mi.gd[k] - 1);
        };
        std::vector<int> start((size_t)total + 1, 0);
        std::vector<int> range((size_t)nt * 6);
        for (int t = 0; t < nt; t++) {
            for (int k = 0; k < 3; k++) {
                float lo = INFINITY, hi = -INFINITY;
                for (int c = 0; c < 3; c++) {
                    const float v = verts[3 * (size_t)tris[3 * (size_t)t + c] + k];
                    lo = fminf(lo, v); hi = fmaxf(hi, v);
                }
                range[6 * (size_t)t + 2 * k] = cell_of(lo, k);
                range[6 * (size_t)t + 2 * k + 1] = cell_of(hi, k);
            }
            const int *r = &range[6 * (size_t)t];
            for (int z = r[4]; z <= r[5]; z++)
                for (int y = r[2]; y <= r[3]; y++)
                    for (int x = r[0]; x <= r[1]; x++) start[((size_t)z * mi.gd[1] + y) * mi.gd[0] + x + 1]++;
        }
        for (long c = 0; c < total; c++) start[c + 1] += start[c];
        std::vector<int> list((size_t)std::max(start[total], 1));
        std::vector<int> cur(start.begin(), start.end() - 1);
        for (int t = 0; t < nt; t++) { // ascending triangle index inside every cell
            const int *r = &range[6 * (size_t)t];
            for (int z = r[4]; z <= r[5]; z++)
                for (int y = r[2]; y <= r[3]; y++)
                    for (int x = r[0]; x <= r[1]; x++) list[cur[((size_t)z * mi.gd[1] + y) * mi.gd[0] + x]++] = t;
        }
        OB_CUDA(ob_malloc(&m.d_cell_start, start.size() * sizeof(int)));
        OB_CUDA(ob_malloc(&m.d_cell_tris, list.size() * sizeof(int)));
        OB_CUDA(cudaMemcpy(m.d_cell_start, start.data(), start.size() * sizeof(int), cudaMemcpyHostToDevice));
        OB_CUDA(cudaMemcpy(m.d_cell_tris, list.data(), list.size() * sizeof(int), cudaMemcpyHostToDevice));
        mi.cell_start = m.d_cell_start; mi.cell_tris = m.d_cell_tris;
    }
    e->meshes.n = id + 1;
    e->hmeshes.push_back(std::move(m));
    return id;
}

void eng_mark_bodies_dirty(Engine *e) { e->bodies_dirty = true; }
void eng_mark_geoms_dirty(Engine *e) { e->geoms_dirty = true; }
void eng_mark_forces_dirty(Engine *e) { e->forces_dirty = true; }
void eng_set_num_envs(Engine *e, int n) {
    if ((n < 1 ? 1 : n) == e->n_envs) return;
    eng_sync_to_host(e);
    e->n_envs = n < 1 ? 1 : n;
    e->geoms_dirty = true; // the per-env body and geom ranges depend on the env count
    e->bodies_dirty = true;
    // measured on C4 (profiles/README.md): spreading the colours makes MORE phases whose cost is set by
    // the longest manifold in the phase, so lowest-free colouring stays the default for batched worlds too
}
void eng_set_capacity(Engine *e, long max_pairs, long max_manifolds) { e->want_pairs = max_pairs; e->want_manifolds = max_manifolds; }
void eng_set_big_extent(Engine *e, float extent) { e->big_extent = extent; }
void eng_set_broadphase(Engine *e, int mode) { e->broad_mode = mode; e->geoms_dirty = true; }
void eng_set_solver_mode(Engine *e, int mode, int env_group) { e->solver_mode = mode; e->env_group = env_group; }
void eng_set_contact_units(Engine *e, int per_contact) { e->contact_units = per_contact; }
void eng_set_colour_spread(Engine *e, int k) { e->colour_spread = k < 0 ? 0 : (k > 32 ? 32 : k); e->colour_spread_auto = false; }
void eng_enable_timing(Engine *e, int on) { e->timing = on != 0; }

// grow body / geom arrays to hold the host mirrors
void engine_ensure_capacity(Engine *e) {
    cudaStream_t st = e->st;
    if (e->hb.n > e->cap_b) {
        const size_t o = (size_t)e->cap_b, n = (size_t)e->hb.n + (size_t)e->hb.n / 4 + 64;
        BodyArrays &B = e->B;
        dev_realloc(B.pos, o, n, st); dev_realloc(B.quat, o, n, st); dev_realloc(B.R, 3 * o, 3 * n, st);
        dev_realloc(B.lvel, o, n, st); dev_realloc(B.avel, o, n, st); dev_realloc(B.I, 3 * o, 3 * n, st);
        dev_realloc(B.invI, 3 * o, 3 * n, st); dev_realloc(B.facc, o, n, st); dev_realloc(B.tacc, o, n, st);
        dev_realloc(B.flags, o, n, st); dev_realloc(B.local, o, n, st); dev_realloc(B.env, o, n, st);
        dev_realloc(B.tmp, 2 * o, 2 * n, st, false);
        // the solver's hot per-body data -- accumulators fc (32 B) + world inverse inertia (48 B) -- live in
        // ONE allocation so that a single L2 access-policy window can keep them resident while the row
        // records stream through (ncu on the 1 M-body pile: they were 2/3 of the solver's DRAM traffic)
        dev_realloc(e->body_hot, 0, 5 * n, st, false);
        B.fc = e->body_hot;
        B.inv = e->body_hot + 2 * n;
        {
            cudaDeviceProp prop;
            OB_CUDA(cudaGetDeviceProperties(&prop, e->device));
            size_t bytes = 5 * n * sizeof(float4);
            size_t persist = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, bytes);
            cudaStreamAttrValue attr;
            memset(&attr, 0, sizeof(attr));
            if (persist > 0 && prop.accessPolicyMaxWindowSize > 0 && e->l2_persist) {
                OB_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist));
                attr.accessPolicyWindow.base_ptr = e->body_hot;
                attr.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, (size_t)prop.accessPolicyMaxWindowSize);
                attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)persist / (double)attr.accessPolicyWindow.num_bytes);
                attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            }
            OB_CUDA(cudaStreamSetAttribute(e->st, cudaStreamAttributeAccessPolicyWindow, &attr));
        }
        OB_CUDA(cudaStreamSynchronize(e->copy_st));
        dev_realloc(e->snap_buf[0], 16 * o, 16 * n, st); dev_realloc(e->snap_buf[1], 16 * o, 16 * n, st);
        B.snap = e->snap_buf[e->snap_cur];
        B.snap_fmt = e->snap_fmt;
        dev_realloc(B.colmask, o, n, st, false);
        dev_realloc(B.prio, o, n, st, false);
        e->cap_b = (int)n;
    }
    if (e->hg.n > e->cap_g) {
        const size_t o = (size_t)e->cap_g, n = (size_t)e->hg.n + (size_t)e->hg.n / 4 + 64;
        GeomArrays &G = e->G;
        dev_realloc(G.type, o, n, st); dev_realloc(G.dims, o, n, st); dev_realloc(G.body, o, n, st);
        dev_realloc(G.pos, o, n, st); dev_realloc(G.R, 3 * o, 3 * n, st); dev_realloc(G.cat, o, n, st);
        dev_realloc(G.col, o, n, st); dev_realloc(G.env, o, n, st); dev_realloc(G.mesh, o, n, st);
        dev_realloc(G.alive, o, n, st); dev_realloc(G.amin, o, n, st, false); dev_realloc(G.amax, o, n, st, false);
        BroadPhase &bp = e->bp;
        dev_realloc(bp.keys, 0, n, st, false); dev_realloc(bp.idx, 0, n, st, false);
        dev_realloc(bp.s_min, 0, n, st, false); dev_realloc(bp.s_max, 0, n, st, false);
        dev_realloc(bp.s_flt, 0, n, st, false);
        dev_realloc(bp.cnt, 0, (size_t)PC_COUNT * n + 1, st, false);
        bp.cap_blk = (long)PC_COUNT * (long)(n / 128 + 2 + (size_t)std::max(e->n_envs, 1) + 1) + 1;
        dev_realloc(bp.blk, 0, (size_t)bp.cap_blk, st, false);
        dev_realloc(bp.sweep_tmp, 0, (size_t)SWEEP_TCAP * n, st, false);
        dev_realloc(bp.sweep_tot, 0, n, st, false);
        bp.cap_geoms = (int)n;
        e->cap_g = (int)n;
    }
    e->B.n = e->hb.n;
    e->G.n = e->hg.n;
}

// pair / manifold / solver capacities. Defaults: 8 pairs and 6 manifolds per geom (+ slack), which
// covers dense piles (measured ~3-6 pairs per body); overflow is flagged in StepStats, never silent.
void engine_ensure_pair_capacity(Engine *e) {
    cudaStream_t st = e->st;
    BroadPhase &bp = e->bp;
    if (!bp.cell_start) {
        bp.cap_cells = (1 << 24) - 2;
        bp.key_bits = 24;
        OB_CUDA(ob_malloc(&bp.cell_start, ((size_t)bp.cap_cells + 2) * sizeof(int)));
        OB_CUDA(ob_malloc(&bp.cell_end, ((size_t)bp.cap_cells + 2) * sizeof(int)));
        OB_CUDA(cudaMemsetAsync(bp.cell_start, 0, ((size_t)bp.cap_cells + 2) * sizeof(int), st));
        OB_CUDA(cudaMemsetAsync(bp.cell_end, 0, ((size_t)bp.cap_cells + 2) * sizeof(int), st));
    }
    // sized from the geom CAPACITY (which grows geometrically), so that a spawn does not re-allocate these
    const long ng = std::max(e->cap_g, e->hg.n);
    long wantp = e->want_pairs > 0 ? e->want_pairs : ng * 8 + 1024;
    if (wantp > bp.cap_pairs) {
        const size_t n = (size_t)wantp;
        dev_realloc(bp.pairs, 0, n, st, false);
        dev_realloc(bp.bb_list, 0, n, st, false);
        dev_realloc(e->cs.pd, 0, n * 8, st, false);
        dev_realloc(e->cs.ns, 0, n * 8, st, false);
        dev_realloc(e->cs.nc, 0, n, st, false);
        dev_realloc(e->M.flag, 0, n + 1, st, false);
        bp.cap_pairs = (int)n;
        e->cs.stride = (int)n;
        e->have_device_contacts = false;
    }
    // solver units: manifolds (<= pairs) or, with per-contact units, contacts
    const bool per_contact = e->contact_units >= 0 ? e->contact_units == 1 : e->n_envs > 1;
    long wantm = e->want_manifolds > 0 ? e->want_manifolds : ng * (per_contact ? 8 : 6) + 1024;
    if (!per_contact && wantm > bp.cap_pairs) wantm = bp.cap_pairs;
    if ((long)e->st_mrec.size() > wantm) wantm = (long)e->st_mrec.size();
    if (wantm > e->M.cap) {
        const size_t n = (size_t)wantm;
        ManifoldArrays &M = e->M;
        dev_realloc(M.rec, 0, n, st, false); dev_realloc(M.colour, 0, n, st, false);
        dev_realloc(M.skey, 0, n, st, false); dev_realloc(M.sidx, 0, n, st, false);
        M.cap = (int)n;
        SolverArrays &S = e->S;
        dev_realloc(S.q0, 0, n * 8, st, false); dev_realloc(S.q1, 0, n * 8, st, false);
        dev_realloc(S.q2, 0, n * 8, st, false); dev_realloc(S.q3, 0, n * 8, st, false);
        dev_realloc(S.q4, 0, n * 8, st, false); dev_realloc(S.q5, 0, n * 8, st, false);
        dev_realloc(S.lam, 0, n * 8, st, false);
        dev_realloc(S.mrec, 0, n, st, false);
        S.cap = (int)n;
    }
    if (e->n_envs > 1) {
        if (e->M.cap > e->cap_env_rec) {
            const size_t n = (size_t)e->M.cap;
            dev_realloc(e->E.rec, 0, n, st, false); dev_realloc(e->E.perm, 0, n, st, false);
            dev_realloc(e->E.col, 0, n, st, false);
            e->cap_env_rec = (int)n;
        }
        if (e->n_envs + 1 > e->cap_envs) {
            const size_t n = (size_t)e->n_envs + 1;
            dev_realloc(e->E.cnt, 0, n, st, false); dev_realloc(e->E.start, 0, n, st, false);
            dev_realloc(e->E.fill, 0, n, st, false);
            dev_realloc(e->E.order, 0, n, st, false);
            e->cap_envs = (int)n;
        }
    }
    e->E.n_envs = e->n_envs;
    e->E.cap = e->cap_env_rec;
}

template <typename T>
static void upload(T *dst, const void *src, size_t count, cudaStream_t st) {
    if (count) OB_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyHostToDevice, st));
}

// ---- incremental ingestion --------------------------------------------------------------------
// A spawn (reference: MSGTYPE_S_NEW_BODY -> AddBody, src/main.c:178-182, 695-733) or a per-tick setter
// (dBodySetPosition of a player body) touches a handful of entries of a world that may hold millions.
// Such edits are queued as (index, field mask); the next collide/step packs them into one pinned
// staging buffer, sends it with one copy and applies it with one scatter kernel -- no device-to-host
// refresh, no re-upload of the arrays, no synchronisation per body.
struct BodyPatch {
    int idx, mask, flags, env;
    int local, pad0, pad1, pad2;
    float4 pos, quat, lvel, avel, facc, tacc;
    float4 R[3], I[3], invI[3];
};
struct GeomPatch {
    int idx, type, body, env;
    unsigned cat, col;
    int alive, mesh;
    float4 dims, pos;
    float4 R[3];
};

__global__ void __launch_bounds__(128) k_patch_bodies(BodyArrays B, const BodyPatch *__restrict__ p, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const BodyPatch q = p[t];
    const int i = q.idx, m = q.mask;
    const bool fresh_body = m & 64;
    float4 pos = fresh_body ? q.pos : B.pos[i], lv = fresh_body ? q.lvel : B.lvel[i];
    if (m & FLD_POS) { pos.x = q.pos.x; pos.y = q.pos.y; pos.z = q.pos.z; }
    if (m & FLD_LVEL) { lv.x = q.lvel.x; lv.y = q.lvel.y; lv.z = q.lvel.z; }
    if (m & FLD_MASS) {
        pos.w = q.pos.w; lv.w = q.lvel.w; // inverse mass rides in pos.w, mass in lvel.w
        B.flags[i] = q.flags;
        for (int k = 0; k < 3; k++) { B.I[3 * i + k] = q.I[k]; B.invI[3 * i + k] = q.invI[k]; }
    }
    if (m & (FLD_POS | FLD_MASS)) B.pos[i] = pos;
    if (m & (FLD_LVEL | FLD_MASS)) B.lvel[i] = lv;
    if (m & FLD_ROT) {
        B.quat[i] = q.quat;
        for (int k = 0; k < 3; k++) B.R[3 * i + k] = q.R[k];
    }
    if (m & FLD_AVEL) B.avel[i] = q.avel;
    if (m & FLD_FORCE) { B.facc[i] = q.facc; B.tacc[i] = q.tacc; }
    if (fresh_body) { B.env[i] = q.env; B.local[i] = q.local; }
}

__global__ void __launch_bounds__(128) k_patch_geoms(GeomArrays G, const GeomPatch *__restrict__ p, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const GeomPatch q = p[t];
    const int i = q.idx;
    G.type[i] = q.type; G.body[i] = q.body; G.env[i] = q.env; G.cat[i] = q.cat; G.col[i] = q.col;
    G.alive[i] = q.alive; G.mesh[i] = q.mesh; G.dims[i] = q.dims;
    if (q.body < 0) { // body-attached geoms take their pose from the body at every collide
        G.pos[i] = q.pos;
        for (int k = 0; k < 3; k++) G.R[3 * i + k] = q.R[k];
    }
}

static void *patch_stage(Engine *e, size_t bytes) {
    if (e->patch_inflight) { // the previous batch may still be reading the staging buffer
        OB_CUDA(cudaEventSynchronize(e->ev_patch));
        e->patch_inflight = false;
    }
    if (bytes > e->cap_patch) {
        if (e->h_patch) OB_CUDA(cudaFreeHost(e->h_patch));
        if (e->d_patch) OB_CUDA(ob_free(e->d_patch));
        e->cap_patch = bytes * 2 + 4096;
        OB_CUDA(cudaMallocHost(&e->h_patch, e->cap_patch));
        OB_CUDA(ob_malloc(&e->d_patch, e->cap_patch));
    }
    if (!e->ev_patch) OB_CUDA(cudaEventCreateWithFlags(&e->ev_patch, cudaEventDisableTiming));
    return e->h_patch;
}

template <typename V>
static inline float4 ld4(const V &v, size_t i) { return make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]); }

// per-env body ranges: index of a body relative to the first body of its env makes the colouring priorities,
// and with them the Gauss-Seidel order of a world, independent of which other worlds share the batch;
// contiguous envs let the island solver stage body data in shared memory.  Bodies are only ever appended.
static int env_note_body(Engine *e, int i, int en) {
    if (en < 0 || en >= (int)e->env_first.size()) { e->env_contig = false; e->env_max_local = std::max(e->env_max_local, i); return i; }
    if (e->env_cnt[en] == 0) e->env_first[en] = i;
    else if (e->env_first[en] + e->env_cnt[en] != i) e->env_contig = false;
    e->env_cnt[en]++;
    const int local = i - e->env_first[en];
    e->env_max_local = std::max(e->env_max_local, local);
    return local;
}

static void env_upload_tables(Engine *e, cudaStream_t st) {
    const size_t ne = e->env_first.size();
    if ((int)ne > e->cap_env_bodies) {
        dev_realloc(e->E.first_body, 0, ne + 1, st, false);
        dev_realloc(e->E.n_body, 0, ne + 1, st, false);
        e->cap_env_bodies = (int)ne;
    }
    e->E.max_bodies = e->env_max_local + 1;
    e->E.contiguous = e->env_contig ? 1 : 0;
    upload(e->E.first_body, e->env_first.data(), ne, st);
    upload(e->E.n_body, e->env_cnt.data(), ne, st);
}

// per-env geom ranges of the all-pairs-per-env broadphase (same append-only bookkeeping)
static void envg_note_geom(Engine *e, int i) {
    const HostGeoms &g = e->hg;
    const int ne = (int)e->envg_first.size();
    if (ne == 1) { e->envg_cnt[0] = i + 1; e->envg_max = i + 1; return; }
    const int en = g.env[i];
    if (en < 0) { e->envg_shared.push_back(i); return; }
    if (en >= ne) { e->envg_ok = false; return; }
    if (e->envg_cnt[en] == 0) e->envg_first[en] = i;
    else if (e->envg_first[en] + e->envg_cnt[en] != i) e->envg_ok = false; // not one contiguous range
    e->envg_cnt[en]++;
    e->envg_max = std::max(e->envg_max, e->envg_cnt[en]);
}

static void envg_upload_tables(Engine *e, cudaStream_t st) {
    EnvBroad &eb = e->EB;
    const int ne = (int)e->envg_first.size();
    bool ok = e->broad_mode != 0 && e->hg.n > 0 && e->envg_ok;
    if (ok && ne == 1) ok = e->hg.n <= (e->broad_mode == 1 ? 4096 : 1024);
    else if (ok) ok = e->envg_max <= 2048 && e->envg_shared.size() <= 256;
    eb.enabled = 0;
    if (!ok) return;
    if (ne > e->cap_eb_envs) {
        dev_realloc(eb.first, 0, (size_t)ne, st, false); dev_realloc(eb.count, 0, (size_t)ne, st, false);
        e->cap_eb_envs = ne;
    }
    if ((int)e->envg_shared.size() > e->cap_eb_shared) {
        const size_t cap = e->envg_shared.size() + 64;
        dev_realloc(eb.shared, 0, cap, st, false);
        e->cap_eb_shared = (int)cap;
    }
    upload(eb.first, e->envg_first.data(), (size_t)ne, st); upload(eb.count, e->envg_cnt.data(), (size_t)ne, st);
    upload(eb.shared, e->envg_shared.data(), e->envg_shared.size(), st);
    eb.enabled = 1;
    eb.single = ne == 1 ? 1 : 0;
    eb.n_shared = (int)e->envg_shared.size();
    eb.n_alive = e->envg_alive;
    eb.n_envs = ne;
    eb.max_count = e->envg_max;
}

static void clear_dirty_bodies(Engine *e) {
    for (int i : e->dirty_b) e->mask_b[i] = 0;
    e->dirty_b.clear();
}
static void clear_dirty_geoms(Engine *e) {
    for (int i : e->dirty_g) e->mask_g[i] = 0;
    e->dirty_g.clear();
}

void eng_mark_body_fields(Engine *e, int i, int fields) {
    if (fields & FLD_FORCE) { // accumulators are consumed by the next step: remember whose mirror to clear then
        if ((int)e->inforce_b.size() < e->hb.n) e->inforce_b.resize((size_t)e->hb.n, 0);
        if (!e->inforce_b[i]) { e->inforce_b[i] = 1; e->force_b.push_back(i); }
    }
    if (e->bodies_dirty) return; // a full upload is pending anyway
    if ((int)e->mask_b.size() < e->hb.n) e->mask_b.resize((size_t)e->hb.n, 0);
    if (!e->mask_b[i]) e->dirty_b.push_back(i);
    e->mask_b[i] |= (unsigned char)fields;
}
void eng_mark_geom(Engine *e, int i) {
    if (e->geoms_dirty) return;
    if ((int)e->mask_g.size() < e->hg.n) e->mask_g.resize((size_t)e->hg.n, 0);
    if (!e->mask_g[i]) e->dirty_g.push_back(i);
    e->mask_g[i] = 1;
}

void eng_sync_to_device(Engine *e) {
    OB_CUDA(cudaSetDevice(e->device));
    // A full upload sends the host mirrors as they are, so it needs them coherent.  Setters queue field
    // patches instead; only the bulk calls (and the very first sync) ask for a full upload, and they
    // refresh the mirrors first (shim: fresh()).
    if (e->bodies_dirty && e->host_stale && e->n_b_dev > 0) {
        fprintf(stderr, "libode_b200: internal error: full upload requested over stale host mirrors\n");
        abort();
    }
    engine_ensure_capacity(e);
    cudaStream_t st = e->st;
    if (e->bodies_dirty) {
        const HostBodies &b = e->hb;
        const size_t n = (size_t)b.n;
        upload(e->B.pos, b.pos.data(), n, st); upload(e->B.quat, b.quat.data(), n, st);
        upload(e->B.R, b.R.data(), 3 * n, st); upload(e->B.lvel, b.lvel.data(), n, st);
        upload(e->B.avel, b.avel.data(), n, st); upload(e->B.I, b.I.data(), 3 * n, st);
        upload(e->B.invI, b.invI.data(), 3 * n, st); upload(e->B.facc, b.facc.data(), n, st);
        upload(e->B.tacc, b.tacc.data(), n, st); upload(e->B.flags, b.flags.data(), n, st);
        {
            const size_t ne = (size_t)std::max(e->n_envs, 1);
            e->env_first.assign(ne, 0); e->env_cnt.assign(ne, 0);
            e->env_max_local = 0; e->env_contig = true;
            std::vector<int> local(n);
            for (size_t i = 0; i < n; i++) local[i] = env_note_body(e, (int)i, b.env[i]);
            env_upload_tables(e, st);
            upload(e->B.local, local.data(), n, st);
            upload(e->B.env, b.env.data(), n, st);
            OB_CUDA(cudaStreamSynchronize(st)); // `local` is a temporary
        }
        e->bodies_dirty = false;
        e->forces_dirty = false;
        e->n_b_dev = b.n;
        e->snap_stale = true;
        clear_dirty_bodies(e);
    } else if (!e->dirty_b.empty()) {
        const HostBodies &b = e->hb;
        const size_t nd = e->dirty_b.size();
        BodyPatch *hp = static_cast<BodyPatch *>(patch_stage(e, nd * sizeof(BodyPatch)));
        std::sort(e->dirty_b.begin(), e->dirty_b.end()); // new bodies join their env ranges in index order
        bool grew = false;
        for (size_t k = 0; k < nd; k++) {
            const int i = e->dirty_b[k];
            BodyPatch &q = hp[k];
            q.idx = i; q.mask = e->mask_b[i]; q.flags = b.flags[i]; q.env = b.env[i];
            q.local = 0; q.pad0 = q.pad1 = q.pad2 = 0;
            if (i >= e->n_b_dev) { // appended since the last sync
                q.mask = FLD_ALL | 64;
                q.local = env_note_body(e, i, b.env[i]);
                grew = true;
            }
            q.pos = ld4(b.pos, i); q.quat = ld4(b.quat, i); q.lvel = ld4(b.lvel, i); q.avel = ld4(b.avel, i);
            q.facc = ld4(b.facc, i); q.tacc = ld4(b.tacc, i);
            for (int r = 0; r < 3; r++) {
                q.R[r] = ld4(b.R, 3 * (size_t)i + r); q.I[r] = ld4(b.I, 3 * (size_t)i + r);
                q.invI[r] = ld4(b.invI, 3 * (size_t)i + r);
            }
        }
        OB_CUDA(cudaMemcpyAsync(e->d_patch, hp, nd * sizeof(BodyPatch), cudaMemcpyHostToDevice, st));
        k_patch_bodies<<<(unsigned)((nd + 127) / 128), 128, 0, st>>>(e->B, static_cast<const BodyPatch *>(e->d_patch), (int)nd);
        OB_CHECK_KERNEL("k_patch_bodies", st);
        OB_CUDA(cudaEventRecord(e->ev_patch, st));
        e->patch_inflight = true;
        if (grew) env_upload_tables(e, st);
        e->n_b_dev = b.n;
        e->snap_stale = true; // spawned or moved bodies: their snapshot records are rebuilt on demand
        clear_dirty_bodies(e);
    }
    if (e->forces_dirty) { // bulk force upload (whole arrays)
        const HostBodies &b = e->hb;
        upload(e->B.facc, b.facc.data(), (size_t)b.n, st);
        upload(e->B.tacc, b.tacc.data(), (size_t)b.n, st);
        e->forces_dirty = false;
    }
    if (e->geoms_dirty) {
        const HostGeoms &g = e->hg;
        const size_t n = (size_t)g.n;
        std::vector<int> mesh(n, 0);
        for (size_t i = 0; i < n; i++) if (g.type[i] == G_TRIMESH) mesh[i] = (int)g.dims[4 * i];
        upload(e->G.type, g.type.data(), n, st); upload(e->G.dims, g.dims.data(), n, st);
        upload(e->G.body, g.body.data(), n, st); upload(e->G.pos, g.pos.data(), n, st);
        upload(e->G.R, g.R.data(), 3 * n, st); upload(e->G.cat, g.cat.data(), n, st);
        upload(e->G.col, g.col.data(), n, st); upload(e->G.env, g.env.data(), n, st);
        upload(e->G.alive, g.alive.data(), n, st);
        upload(e->G.mesh, mesh.data(), n, st);
        {
            const size_t ne = (size_t)std::max(e->n_envs, 1);
            e->envg_first.assign(ne, 0); e->envg_cnt.assign(ne, 0); e->envg_shared.clear();
            e->envg_max = 0; e->envg_ok = true; e->envg_alive = 0;
            e->alive_dev.assign(n, 0);
            for (size_t i = 0; i < n; i++) {
                envg_note_geom(e, (int)i);
                e->alive_dev[i] = g.alive[i] ? 1 : 0;
                e->envg_alive += e->alive_dev[i];
            }
            envg_upload_tables(e, st);
        }
        OB_CUDA(cudaStreamSynchronize(st)); // `mesh` is a temporary
        e->geoms_dirty = false;
        e->n_g_dev = g.n;
        clear_dirty_geoms(e);
    } else if (!e->dirty_g.empty()) {
        const HostGeoms &g = e->hg;
        const size_t nd = e->dirty_g.size();
        // the body batch may still be in flight in the shared staging buffer: geoms get the second half
        if (e->patch_inflight) { OB_CUDA(cudaEventSynchronize(e->ev_patch)); e->patch_inflight = false; }
        GeomPatch *hp = static_cast<GeomPatch *>(patch_stage(e, nd * sizeof(GeomPatch)));
        std::sort(e->dirty_g.begin(), e->dirty_g.end());
        bool tables = false;
        for (size_t k = 0; k < nd; k++) {
            const int i = e->dirty_g[k];
            GeomPatch &q = hp[k];
            q.idx = i; q.type = g.type[i]; q.body = g.body[i]; q.env = g.env[i]; q.cat = g.cat[i]; q.col = g.col[i];
            q.alive = g.alive[i]; q.mesh = g.type[i] == G_TRIMESH ? (int)g.dims[4 * (size_t)i] : 0;
            q.dims = ld4(g.dims, i); q.pos = ld4(g.pos, i);
            for (int r = 0; r < 3; r++) q.R[r] = ld4(g.R, 3 * (size_t)i + r);
            if ((int)e->alive_dev.size() < g.n) e->alive_dev.resize((size_t)g.n, 0);
            const unsigned char now = g.alive[i] ? 1 : 0;
            if (i >= e->n_g_dev) { envg_note_geom(e, i); e->envg_alive += now; tables = true; }
            else if (now != e->alive_dev[i]) e->envg_alive += (int)now - (int)e->alive_dev[i];
            e->alive_dev[i] = now;
        }
        OB_CUDA(cudaMemcpyAsync(e->d_patch, hp, nd * sizeof(GeomPatch), cudaMemcpyHostToDevice, st));
        k_patch_geoms<<<(unsigned)((nd + 127) / 128), 128, 0, st>>>(e->G, static_cast<const GeomPatch *>(e->d_patch), (int)nd);
        OB_CHECK_KERNEL("k_patch_geoms", st);
        OB_CUDA(cudaEventRecord(e->ev_patch, st));
        e->patch_inflight = true;
        // (the per-env ranges are append-only; the alive count follows the patches)
        if (tables) envg_upload_tables(e, st);
        else { e->EB.n_alive = e->envg_alive; e->EB.max_count = e->envg_max; }
        e->n_g_dev = g.n;
        clear_dirty_geoms(e);
    }
}

void eng_sync_to_host(Engine *e) {
    if (!e->host_stale) return;
    OB_CUDA(cudaSetDevice(e->device));
    // queued field edits live only in the mirrors: send them before the mirrors are overwritten
    if (!e->dirty_b.empty() || !e->dirty_g.empty()) eng_sync_to_device(e);
    HostBodies &b = e->hb;
    const size_t n = (size_t)std::min(b.n, e->B.n);
    cudaStream_t st = e->st;
    if (n) {
        OB_CUDA(cudaMemcpyAsync(b.pos.data(), e->B.pos, n * 16, cudaMemcpyDeviceToHost, st));
        OB_CUDA(cudaMemcpyAsync(b.quat.data(), e->B.quat, n * 16, cudaMemcpyDeviceToHost, st));
        OB_CUDA(cudaMemcpyAsync(b.R.data(), e->B.R, n * 48, cudaMemcpyDeviceToHost, st));
        OB_CUDA(cudaMemcpyAsync(b.lvel.data(), e->B.lvel, n * 16, cudaMemcpyDeviceToHost, st));
        OB_CUDA(cudaMemcpyAsync(b.avel.data(), e->B.avel, n * 16, cudaMemcpyDeviceToHost, st));
    }
    OB_CUDA(cudaStreamSynchronize(st));
    e->host_stale = false;
}

// ---- CUDA graphs for the tick of batched worlds ----------------------------------------------------
// A C4 tick is ~20 short launches before the one long solver kernel.  Each launch makes the GPU fetch its
// commands from host memory; while the application's own snapshot copy saturates the PCIe link those fetches
// queue behind it (measured: collide 0.33 -> 0.52 ms, prepare 0.08 -> 0.20 ms during a 64 MB D2H copy).  After
// three identical ticks the launch sequence is captured once and replayed as a graph; the key covers every
// pointer, size, option and parameter the captured launches depend on, so any change falls back to plain
// launches and re-captures.
static inline unsigned long long gk_mix(unsigned long long h, unsigned long long v) {
    return h ^ (v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
}
static inline unsigned long long gk_ptr(unsigned long long h, const void *p) { return gk_mix(h, (unsigned long long)(uintptr_t)p); }
static inline unsigned long long gk_bytes(unsigned long long h, const void *p, size_t n) {
    const unsigned char *b = static_cast<const unsigned char *>(p);
    for (size_t i = 0; i < n; i++) h = gk_mix(h, b[i]);
    return h;
}
static unsigned long long graph_key(Engine *e, bool step) {
    unsigned long long h = 0x243F6A8885A308D3ull;
    if (step) h = gk_ptr(h, e->B.snap); // alternates between the two snapshot buffers: one step graph per buffer
    const void *ptrs[] = {e->B.pos, e->B.quat, e->B.fc, e->B.local, e->G.pos, e->G.amin, e->bp.pairs, e->bp.bb_list, e->bp.cnt, e->bp.blk,
                          e->bp.sweep_tmp, e->bp.keys, e->bp.scan.sums[0], e->bp.scan.sums[1], e->scan.sums[0], e->scan.sums[1],
                          e->cs.pd, e->cs.nc, e->M.rec, e->M.flag, e->S.q0, e->S.mrec, e->E.rec, e->E.cnt, e->E.start, e->E.fill,
                          e->E.first_body, e->EB.first, e->EB.count, e->EB.shared, e->d_stats};
    for (const void *p : ptrs) h = gk_ptr(h, p);
    const void *sorts[] = {e->bp.sort.hist, e->bp.sort.keys_tmp, e->bp.sort.vals_tmp, e->bp.sort.scan.sums[0], e->bp.sort.scan.sums[1],
                           e->sort.hist, e->sort.keys_tmp, e->sort.vals_tmp, e->sort.scan.sums[0], e->sort.scan.sums[1],
                           e->bp.cell_start, e->bp.s_min, e->M.skey, e->M.sidx, e->M.colour};
    for (const void *p : sorts) h = gk_ptr(h, p);
    for (int m = 0; m < e->meshes.n; m++) {
        h = gk_ptr(h, e->meshes.m[m].verts);
        h = gk_ptr(h, e->meshes.m[m].tris);
        h = gk_mix(h, ((unsigned long long)e->meshes.m[m].nt << 32) | (unsigned)e->meshes.m[m].nv);
    }
    const int ints[] = {e->B.n, e->G.n, e->cap_b, e->cap_g, e->bp.cap_pairs, e->cs.stride, e->M.cap, e->S.cap, e->n_envs, e->max_contacts,
                        e->env_group, e->contact_units, e->solver_mode, e->env_stage, e->env_pair, e->env_pair_rows, e->env_fuse, e->colour_spread, e->broad_mode,
                        e->tiny_solver, (int)e->keep_fc, e->EB.enabled, e->EB.single, e->EB.n_shared, e->EB.n_alive, e->EB.max_count, e->EB.n_envs, e->EB.sap, e->E.contiguous,
                        e->E.max_bodies, e->meshes.n, (int)e->have_device_contacts, e->snap_fmt};
    h = gk_bytes(h, ints, sizeof(ints));
    h = gk_bytes(h, &e->params, sizeof(e->params));
    h = gk_bytes(h, &e->big_extent, sizeof(float));
    return h;
}
static bool graphs_usable(Engine *e) {
    // every device-resident tick: batched worlds (island solver) and single worlds (cooperative colouring + solver)
    return e->graphs && !e->timing && !ob_debug_sync() && !(e->params.tol > 0.f);
}
template <typename F>
static void run_graphed(Engine *e, Engine::TickGraph &g, unsigned long long key, F &&enqueue) {
    if (g.exec && g.key == key) {
        OB_CUDA(cudaGraphLaunch(g.exec, e->st));
        g_ob_launches += g.kernels;
        e->fc_valid = g.fc_valid;
        return;
    }
    if (g.warm_key == key) g.warm++;
    else { g.warm_key = key; g.warm = 0; }
    if (g.warm < 3) { // workspaces are sized and function attributes set by the plain launches of these ticks
        enqueue();
        return;
    }
    if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    const long before = g_ob_launches;
    cudaGraph_t graph = nullptr;
    OB_CUDA(cudaStreamBeginCapture(e->st, cudaStreamCaptureModeThreadLocal));
    enqueue();
    OB_CUDA(cudaStreamEndCapture(e->st, &graph));
    g.kernels = (int)(g_ob_launches - before);
    g.fc_valid = e->fc_valid;
    OB_CUDA(cudaGraphInstantiate(&g.exec, graph, 0));
    OB_CUDA(cudaGraphDestroy(graph));
    g.key = key;
    if (e->graphs >= 2) fprintf(stderr, "libode_b200: captured a tick graph of %d launches\n", g.kernels);
    OB_CUDA(cudaGraphLaunch(g.exec, e->st));
}

void eng_collide(Engine *e, int max_contacts) {
    eng_sync_to_device(e);
    engine_ensure_pair_capacity(e);
    if (max_contacts < 1) max_contacts = 1;
    if (max_contacts > 8) max_contacts = 8;
    e->max_contacts = max_contacts;
    if (e->timing) OB_CUDA(cudaEventRecord(e->ev[0], e->st));
    auto enqueue = [&]() {
        broadphase_run(e->bp, e->G, e->B.pos, e->B.R, e->meshes, e->n_envs, e->big_extent, e->EB, e->d_stats, e->st);
        if (e->timing) OB_CUDA(cudaEventRecord(e->ev_bp, e->st)); // (timing on: plain launches, never captured)
        narrowphase_run(e->bp, e->G, e->meshes, e->hmeshes, e->cs, max_contacts, e->d_stats, e->num_sms, e->st);
    };
    if (graphs_usable(e)) run_graphed(e, e->g_collide, graph_key(e, false), enqueue);
    else enqueue();
    if (e->timing) OB_CUDA(cudaEventRecord(e->ev[1], e->st));
    e->have_device_contacts = true;
}

// pair-major compaction of the contact slots for the host (compat mode)
__global__ void __launch_bounds__(256) k_compact_contacts(const BroadCounters *__restrict__ bc, ContactSlots cs,
                                                           const int *__restrict__ first, float4 *__restrict__ pd,
                                                           float4 *__restrict__ ns) {
    const int n = bc->n_pairs;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const int c = cs.nc[p], f = first[p];
        for (int k = 0; k < c; k++) {
            pd[f + k] = cs.pd[(size_t)k * cs.stride + p];
            ns[f + k] = cs.ns[(size_t)k * cs.stride + p];
        }
    }
}

// compat mode, fast path: pairs, contact ranges and pair-major contacts written into the mapped host buffer
struct ExportView {
    int *hdr, *g1, *g2, *first, *count; // hdr: n_pairs, n_contacts, spare, spare
    float4 *pd, *ns;
};
static ExportView export_view(unsigned char *base, int cap_pairs, int cap_contacts) {
    ExportView v;
    const size_t cp = ((size_t)cap_pairs + 4 + 3) & ~(size_t)3; // ints per array, 16-byte multiples
    v.hdr = reinterpret_cast<int *>(base);
    v.g1 = v.hdr + 4; v.g2 = v.g1 + cp; v.first = v.g2 + cp; v.count = v.first + cp;
    v.pd = reinterpret_cast<float4 *>(v.count + cp);
    v.ns = v.pd + cap_contacts;
    return v;
}
static size_t export_bytes(int cap_pairs, int cap_contacts) {
    const size_t cp = ((size_t)cap_pairs + 4 + 3) & ~(size_t)3;
    return 16 + 4 * cp * sizeof(int) + 2 * (size_t)cap_contacts * sizeof(float4);
}
__global__ void __launch_bounds__(256) k_export_pairs(const BroadCounters *__restrict__ bc, const int2 *__restrict__ pairs,
                                                      ContactSlots cs, const int *__restrict__ first, const int *__restrict__ total,
                                                      ExportView x, int cap_pairs, int cap_contacts) {
    const int np = bc->n_pairs, n = min(np, cap_pairs);
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const int2 pr = pairs[p];
        const int c = cs.nc[p], f = first[p];
        x.g1[p] = pr.x; x.g2[p] = pr.y; x.first[p] = f; x.count[p] = c;
        for (int k = 0; k < c; k++)
            if (f + k < cap_contacts) {
                x.pd[f + k] = cs.pd[(size_t)k * cs.stride + p];
                x.ns[f + k] = cs.ns[(size_t)k * cs.stride + p];
            }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        x.hdr[0] = np; x.hdr[1] = *total; x.hdr[2] = 0; x.hdr[3] = 0;
        x.first[n] = *total;
    }
}
constexpr int EXPORT_MAX_PAIRS = 1 << 16, EXPORT_MAX_CONTACTS = 1 << 17; // beyond: the copying path (bandwidth matters there)

static void export_ensure(Engine *e, int pairs, int contacts) {
    if (pairs <= e->x_cap_pairs && contacts <= e->x_cap_contacts) return;
    if (e->x_host) { OB_CUDA(cudaStreamSynchronize(e->st)); OB_CUDA(cudaFreeHost(e->x_host)); }
    e->x_cap_pairs = std::max(pairs, e->x_cap_pairs);
    e->x_cap_contacts = std::max(contacts, e->x_cap_contacts);
    OB_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&e->x_host), export_bytes(e->x_cap_pairs, e->x_cap_contacts), cudaHostAllocMapped));
    OB_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void **>(&e->x_dev), e->x_host, 0));
}

// dCollide outside a space traversal: one pair through the same narrowphase kernels
HostPairs eng_collide_pair(Engine *e, int g1, int g2, int max_contacts) {
    if (e->have_device_contacts) {
        fprintf(stderr, "libode_b200: dCollide between dSpaceCollideDeviceB200 and the step would discard the world's "
                        "contacts; call it before the collide or after the step\n");
        abort();
    }
    eng_sync_to_device(e);
    engine_ensure_pair_capacity(e);
    if (max_contacts < 1) max_contacts = 1;
    if (max_contacts > 8) max_contacts = 8;
    broadphase_single_pair(e->bp, e->G, e->B.pos, e->B.R, e->meshes, e->big_extent, g1, g2, e->d_stats, e->st);
    narrowphase_run(e->bp, e->G, e->meshes, e->hmeshes, e->cs, max_contacts, e->d_stats, e->num_sms, e->st);
    e->have_device_contacts = true;
    HostPairs hp = eng_fetch_pairs(e);
    e->have_device_contacts = false;
    return hp;
}

HostPairs eng_fetch_pairs(Engine *e) {
    HostPairs hp;
    OB_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = e->st;
    if (!e->have_device_contacts || e->G.n == 0) return hp;
    // fast path: sized by what earlier traversals needed (+ headroom); a tick that outgrows the buffer takes the copying
    // path below once.  Worlds beyond EXPORT_MAX_* always copy.
    if (e->x_want_pairs > EXPORT_MAX_PAIRS || e->x_want_contacts > EXPORT_MAX_CONTACTS) {
        if (e->x_host) { OB_CUDA(cudaStreamSynchronize(st)); OB_CUDA(cudaFreeHost(e->x_host)); e->x_host = nullptr; }
        e->x_cap_pairs = e->x_cap_contacts = 0;
    } else {
        export_ensure(e, e->x_want_pairs, e->x_want_contacts);
        const int capx = e->x_cap_pairs, capc = e->x_cap_contacts;
        if (capx + 2 > e->cap_dl) {
            dev_realloc(e->dl_first, 0, (size_t)capx + 1026, st, false);
            e->cap_dl = capx + 1024;
        }
        int *d_total = e->dl_first + capx + 1;
        scan_exclusive(e->cs.nc, e->dl_first, capx, &e->bp.counters->n_pairs, d_total, e->scan, st);
        const ExportView xd = export_view(e->x_dev, capx, capc);
        k_export_pairs<<<(unsigned)std::min((capx + 255) / 256, e->num_sms * 4), 256, 0, st>>>(e->bp.counters, e->bp.pairs, e->cs, e->dl_first,
                                                                                              d_total, xd, capx, capc);
        OB_CHECK_KERNEL("k_export_pairs", st);
        OB_CUDA(cudaStreamSynchronize(st));
        const ExportView xh = export_view(e->x_host, capx, capc);
        const int np = xh.hdr[0], total = xh.hdr[1];
        if (np <= capx && total <= capc) {
            hp.n_pairs = np;
            hp.g1 = xh.g1; hp.g2 = xh.g2; hp.first = xh.first; hp.count = xh.count;
            hp.pos_depth = reinterpret_cast<const float *>(xh.pd); hp.normal_side = reinterpret_cast<const float *>(xh.ns);
            if (np + np / 4 > capx) e->x_want_pairs = 2 * np + 64;             // keep a quarter of headroom
            if (total + total / 4 > capc) e->x_want_contacts = 2 * total + 64;
            return hp;
        }
        // outgrown (the scan above covered only capx pairs, so `total` is only valid when the pairs fitted)
        if (np > capx) e->x_want_pairs = np + np / 2 + 64;
        e->x_want_contacts = np > capx ? std::max(2 * capc, e->x_want_contacts) : total + total / 2 + 64;
    }
    BroadCounters bc;
    OB_CUDA(cudaMemcpyAsync(&bc, e->bp.counters, sizeof(bc), cudaMemcpyDeviceToHost, st));
    OB_CUDA(cudaStreamSynchronize(st));
    const int np = bc.n_pairs;
    hp.n_pairs = np;
    if (np == 0) return hp;
    if (np + 1 > e->cap_dl) {
        dev_realloc(e->dl_first, 0, (size_t)np + 1025, st, false);
        e->cap_dl = np + 1024;
    }
    int *d_total = e->dl_first + np; // scan total lands after the last prefix
    scan_exclusive(e->cs.nc, e->dl_first, np, nullptr, d_total, e->scan, st);
    e->h_first.resize((size_t)np + 1);
    e->h_count.resize((size_t)np);
    e->h_g1.resize((size_t)np);
    e->h_g2.resize((size_t)np);
    std::vector<int2> tmp_pairs((size_t)np);
    OB_CUDA(cudaMemcpyAsync(e->h_first.data(), e->dl_first, ((size_t)np + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
    OB_CUDA(cudaMemcpyAsync(e->h_count.data(), e->cs.nc, (size_t)np * sizeof(int), cudaMemcpyDeviceToHost, st));
    OB_CUDA(cudaMemcpyAsync(tmp_pairs.data(), e->bp.pairs, (size_t)np * sizeof(int2), cudaMemcpyDeviceToHost, st));
    OB_CUDA(cudaStreamSynchronize(st));
    const int total = e->h_first[(size_t)np];
    for (int i = 0; i < np; i++) { e->h_g1[i] = tmp_pairs[i].x; e->h_g2[i] = tmp_pairs[i].y; }
    e->h_pd.resize((size_t)std::max(total, 1) * 4);
    e->h_ns.resize((size_t)std::max(total, 1) * 4);
    if (total > 0) {
        if (total > e->cap_dlc) {
            e->cap_dlc = total + total / 2 + 256;
            dev_realloc(e->dl_pd, 0, (size_t)e->cap_dlc, st, false);
            dev_realloc(e->dl_ns, 0, (size_t)e->cap_dlc, st, false);
        }
        k_compact_contacts<<<(unsigned)std::min((np + 255) / 256, e->num_sms * 8), 256, 0, st>>>(e->bp.counters, e->cs,
                                                                                               e->dl_first, e->dl_pd, e->dl_ns);
        OB_CHECK_KERNEL("k_compact_contacts", st);
        OB_CUDA(cudaMemcpyAsync(e->h_pd.data(), e->dl_pd, (size_t)total * 16, cudaMemcpyDeviceToHost, st));
        OB_CUDA(cudaMemcpyAsync(e->h_ns.data(), e->dl_ns, (size_t)total * 16, cudaMemcpyDeviceToHost, st));
        OB_CUDA(cudaStreamSynchronize(st));
    }
    hp.g1 = e->h_g1.data(); hp.g2 = e->h_g2.data();
    hp.first = e->h_first.data(); hp.count = e->h_count.data();
    hp.pos_depth = e->h_pd.data(); hp.normal_side = e->h_ns.data();
    return hp;
}

// the step consumes the force accumulators (the device clears its copy in the integrate tail)
static void consume_force_mirrors(Engine *e) {
    HostBodies &b = e->hb;
    for (int i : e->force_b) {
        for (int k = 0; k < 4; k++) { b.facc[4 * (size_t)i + k] = 0.f; b.tacc[4 * (size_t)i + k] = 0.f; }
        e->inforce_b[i] = 0;
    }
    e->force_b.clear();
}

void eng_step_device_contacts(Engine *e, float h, const Surface &surf) {
    eng_sync_to_device(e);
    consume_force_mirrors(e);
    engine_ensure_pair_capacity(e);
    if (e->timing && !e->have_device_contacts) {
        OB_CUDA(cudaEventRecord(e->ev[0], e->st));
        OB_CUDA(cudaEventRecord(e->ev[1], e->st));
    }
    apply_pending_forces(e);
    step_begin(e);
    if (graphs_usable(e) && e->have_device_contacts) {
        unsigned long long key = graph_key(e, true);
        key = gk_bytes(key, &h, sizeof(h));
        key = gk_bytes(key, &surf, sizeof(surf));
        run_graphed(e, e->g_step[e->snap_cur], key, [&]() { solver_step(e, h, false, &surf); });
        e->last_h = h;
    } else {
        solver_step(e, h, false, &surf);
    }
    step_end(e);
    e->have_device_contacts = false;
    e->host_stale = true;
    e->ev_valid = e->timing;
}

__global__ void __launch_bounds__(256) k_file_host_contacts(const float4 *__restrict__ pd, const float4 *__restrict__ ns,
                                                            const Surface *__restrict__ surf, const int4 *__restrict__ mrec, int nc, int nm,
                                                            float4 *__restrict__ o_pd, float4 *__restrict__ o_ns, Surface *__restrict__ o_surf,
                                                            int4 *__restrict__ o_mrec, int *__restrict__ m_count, int *__restrict__ meta) {
    const int gt = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;
    for (int i = gt; i < nc; i += gs) { o_pd[i] = pd[i]; o_ns[i] = ns[i]; o_surf[i] = surf[i]; }
    for (int i = gt; i < nm; i += gs) o_mrec[i] = mrec[i];
    if (gt == 0) *m_count = nm;
    if (gt < 8) meta[gt] = 0;
}

void eng_step_host_contacts(Engine *e, float h, const HostContact *contacts, int n) {
    eng_sync_to_device(e);
    consume_force_mirrors(e);
    // group consecutive joints that attach the same ordered body pair into manifolds (<= 8 each)
    e->st_pd.clear(); e->st_ns.clear(); e->st_surf.clear(); e->st_mrec.clear();
    int cur_b1 = -2, cur_b2 = -2, cur_rev = 0;
    for (int i = 0; i < n; i++) {
        const HostContact &c = contacts[i];
        int b1 = c.b1, b2 = c.b2, rev = 0;
        if (b1 < 0 && b2 < 0) continue; // both NULL: inert joint
        if (b1 < 0) { b1 = b2; b2 = -1; rev = 1; }
        const bool same = !e->st_mrec.empty() && b1 == cur_b1 && b2 == cur_b2 && rev == cur_rev &&
                          (e->st_mrec.back().w & 0xff) < 8;
        if (!same) {
            e->st_mrec.push_back(make_int4(b1, b2, (int)e->st_pd.size(), rev ? (1 << 8) : 0));
            cur_b1 = b1; cur_b2 = b2; cur_rev = rev;
        }
        e->st_mrec.back().w += 1;
        e->st_pd.push_back(make_float4(c.pos[0], c.pos[1], c.pos[2], c.depth));
        e->st_ns.push_back(make_float4(c.normal[0], c.normal[1], c.normal[2], 0.f));
        e->st_surf.push_back(c.surf);
    }
    engine_ensure_pair_capacity(e);
    cudaStream_t st = e->st;
    const int nc = (int)e->st_pd.size(), nm = (int)e->st_mrec.size();
    if (nc > e->cap_hc) {
        const size_t cap = (size_t)nc + (size_t)nc / 2 + 256;
        dev_realloc(e->hc_pd, 0, cap, st, false);
        dev_realloc(e->hc_ns, 0, cap, st, false);
        dev_realloc(e->hc_surf, 0, cap, st, false);
        e->cap_hc = (int)cap;
    }
    {
        // one pinned blob [pd | ns | surfaces | manifold records], one copy, one kernel that files it (instead of five
        // copies and a memset from pageable vectors)
        const size_t o_ns = (size_t)nc * 16, o_surf = 2 * o_ns, o_mrec = o_surf + (((size_t)nc * sizeof(Surface) + 15) & ~(size_t)15);
        const size_t bytes = o_mrec + (size_t)nm * 16;
        if (bytes > e->hcs_cap) {
            if (e->hcs_host) { OB_CUDA(cudaFreeHost(e->hcs_host)); dev_free(e->hcs_dev); }
            e->hcs_cap = bytes + bytes / 2 + 4096;
            OB_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&e->hcs_host), e->hcs_cap, cudaHostAllocDefault));
            dev_realloc(e->hcs_dev, 0, e->hcs_cap, st, false);
        }
        if (nc) {
            memcpy(e->hcs_host, e->st_pd.data(), o_ns);
            memcpy(e->hcs_host + o_ns, e->st_ns.data(), o_ns);
            memcpy(e->hcs_host + o_surf, e->st_surf.data(), (size_t)nc * sizeof(Surface));
            memcpy(e->hcs_host + o_mrec, e->st_mrec.data(), (size_t)nm * 16);
            OB_CUDA(cudaMemcpyAsync(e->hcs_dev, e->hcs_host, bytes, cudaMemcpyHostToDevice, st));
        }
        k_file_host_contacts<<<(unsigned)std::max(1, std::min((nc + 255) / 256, e->num_sms * 2)), 256, 0, st>>>(
            reinterpret_cast<const float4 *>(e->hcs_dev), reinterpret_cast<const float4 *>(e->hcs_dev + o_ns),
            reinterpret_cast<const Surface *>(e->hcs_dev + o_surf), reinterpret_cast<const int4 *>(e->hcs_dev + o_mrec), nc, nm, e->hc_pd,
            e->hc_ns, e->hc_surf, e->M.rec, e->M.count, e->M.meta);
        OB_CHECK_KERNEL("k_file_host_contacts", st);
    }
    if (e->timing) {
        OB_CUDA(cudaEventRecord(e->ev[0], st));
        OB_CUDA(cudaEventRecord(e->ev[1], st));
    }
    apply_pending_forces(e);
    step_begin(e);
    solver_step(e, h, true, nullptr);
    step_end(e);
    OB_CUDA(cudaStreamSynchronize(st)); // staging vectors and &nm must outlive the copies
    e->have_device_contacts = false;
    e->host_stale = true;
    e->ev_valid = e->timing;
}

static void snapshot_refresh_if_stale(Engine *e);
const float *eng_snapshot_device(Engine *e) {
    eng_sync_to_device(e);
    snapshot_refresh_if_stale(e);
    return e->snap_buf[e->snap_cur];
}

// called by both step entry points around solver_step: flip the snapshot buffer, make the kernel that
// will overwrite it wait for a still-draining copy, apply pending force uploads
static void step_begin(Engine *e) {
    const int next = e->snap_cur ^ 1;
    if (e->snap_copy_pending[next]) {
        OB_CUDA(cudaStreamWaitEvent(e->st, e->ev_snap_copied[next], 0));
        e->snap_copy_pending[next] = false;
    }
    e->snap_cur = next;
    e->B.snap = e->snap_buf[next];
    e->B.snap_fmt = e->snap_fmt;
    e->snap_stale = false; // the step's tail writes every body's record
}
static void step_end(Engine *e) { OB_CUDA(cudaEventRecord(e->ev_step_done, e->st)); }

// Snapshot records straight from the body state: for bodies spawned or moved by the host since the last step
// (the reference's broadcast loop reads dBodyGetPosition/Rotation of a freshly added body without a step in
// between, src/main.c:178-182 then :221-242) and after a change of the snapshot format.
__global__ void __launch_bounds__(256) k_snapshot_refresh(BodyArrays B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.n) return;
    snapshot_store(B, i, B.pos[i], B.quat[i], load_m3(B.R, i));
}
static void snapshot_refresh_if_stale(Engine *e) {
    if (!e->snap_stale || e->B.n == 0) return;
    e->B.snap = e->snap_buf[e->snap_cur];
    e->B.snap_fmt = e->snap_fmt;
    if (e->snap_copy_pending[e->snap_cur]) { // a copy of this buffer may still be draining
        OB_CUDA(cudaStreamWaitEvent(e->st, e->ev_snap_copied[e->snap_cur], 0));
        e->snap_copy_pending[e->snap_cur] = false;
    }
    k_snapshot_refresh<<<(unsigned)((e->B.n + 255) / 256), 256, 0, e->st>>>(e->B);
    OB_CHECK_KERNEL("k_snapshot_refresh", e->st);
    e->snap_stale = false;
}
void eng_set_snapshot_format(Engine *e, int fmt) {
    if (fmt < 0 || fmt > 2) { fprintf(stderr, "libode_b200: snapshot format %d does not exist (0, 1, 2)\n", fmt); abort(); }
    if (fmt == e->snap_fmt) return;
    e->snap_fmt = fmt;
    e->snap_stale = true;
}
int eng_snapshot_format(Engine *e) { return e->snap_fmt; }

void eng_snapshot_to_host(Engine *e, float *dst, int first, int count, bool blocking) {
    OB_CUDA(cudaSetDevice(e->device));
    if (count <= 0) return;
    eng_sync_to_device(e);
    snapshot_refresh_if_stale(e);
    const size_t stride = e->snap_fmt == 0 ? 16 : (e->snap_fmt == 1 ? 12 : 8); // floats per record
    // the copy runs on the copy stream, after the step that produced the snapshot, and overlaps later ticks
    const int cur = e->snap_cur;
    OB_CUDA(cudaEventRecord(e->ev_step_done, e->st));
    OB_CUDA(cudaStreamWaitEvent(e->copy_st, e->ev_step_done, 0));
    OB_CUDA(cudaMemcpyAsync(dst, e->snap_buf[cur] + stride * (size_t)first, (size_t)count * stride * sizeof(float),
                            cudaMemcpyDeviceToHost, e->copy_st));
    OB_CUDA(cudaEventRecord(e->ev_snap_copied[cur], e->copy_st));
    e->snap_copy_pending[cur] = true;
    if (blocking) OB_CUDA(cudaStreamSynchronize(e->copy_st));
}

// 6 floats per body (force, torque) -> the float4 accumulators the step consumes
__global__ void __launch_bounds__(256) k_scatter_forces(int n, const float *__restrict__ f6, float4 *__restrict__ facc,
                                                         float4 *__restrict__ tacc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *f = f6 + 6 * (size_t)i;
    facc[i] = make_float4(f[0], f[1], f[2], 0.f);
    tacc[i] = make_float4(f[3], f[4], f[5], 0.f);
}

void eng_set_forces(Engine *e, const float *f6, int n) {
    // One H2D copy straight from the caller's (ideally pinned) buffer on the upload stream; the scatter
    // into the float4 accumulators is deferred to the start of the next step, so the copy overlaps the
    // collide phase.  Replaces the accumulators of bodies [0, n), like dBodySetForce/Torque.
    eng_sync_to_device(e);
    if (n > e->B.n) n = e->B.n;
    if (n <= 0) return;
    if (n > e->cap_f6) {
        OB_CUDA(cudaStreamSynchronize(e->st));
        OB_CUDA(cudaStreamSynchronize(e->h2d_st));
        e->cap_f6 = n + n / 4 + 64;
        for (int i = 0; i < 2; i++) {
            if (e->d_f6[i]) OB_CUDA(ob_free(e->d_f6[i]));
            OB_CUDA(ob_malloc(&e->d_f6[i], (size_t)e->cap_f6 * 6 * sizeof(float)));
            e->f6_inflight[i] = false;
        }
    }
    // the scatter that last read this staging buffer (two uploads ago) must be done
    const int cur = e->f6_cur ^ 1;
    if (e->f6_inflight[cur]) OB_CUDA(cudaStreamWaitEvent(e->h2d_st, e->ev_f6_consumed[cur], 0));
    OB_CUDA(cudaMemcpyAsync(e->d_f6[cur], f6, (size_t)n * 6 * sizeof(float), cudaMemcpyHostToDevice, e->h2d_st));
    OB_CUDA(cudaEventRecord(e->ev_f6, e->h2d_st));
    e->f6_cur = cur;
    e->pending_f6 = n;
}

// before a step consumes the accumulators: wait for the upload and scatter it
static void apply_pending_forces(Engine *e) {
    if (e->pending_f6 <= 0) return;
    const int n = e->pending_f6;
    OB_CUDA(cudaStreamWaitEvent(e->st, e->ev_f6, 0));
    k_scatter_forces<<<(unsigned)((n + 255) / 256), 256, 0, e->st>>>(n, e->d_f6[e->f6_cur], e->B.facc, e->B.tacc);
    OB_CHECK_KERNEL("k_scatter_forces", e->st);
    OB_CUDA(cudaEventRecord(e->ev_f6_consumed[e->f6_cur], e->st));
    e->f6_inflight[e->f6_cur] = true;
    e->pending_f6 = 0;
}

// ---- halo exchange support (slab-decomposed worlds): gather / scatter of body states through DEVICE
// buffers, so NCCL can move them GPU to GPU without touching the host.  16 floats per body:
// pos(3) pad, quat(4), lvel(3) pad, avel(3) pad.
__global__ void __launch_bounds__(256) k_pack_states(int n, const int *__restrict__ idx, BodyArrays B, float4 *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = idx[i];
    const float4 p = B.pos[b], q = B.quat[b], lv = B.lvel[b], av = B.avel[b];
    out[4 * (size_t)i] = make_float4(p.x, p.y, p.z, 0.f);
    out[4 * (size_t)i + 1] = q;
    out[4 * (size_t)i + 2] = make_float4(lv.x, lv.y, lv.z, 0.f);
    out[4 * (size_t)i + 3] = make_float4(av.x, av.y, av.z, 0.f);
}

__global__ void __launch_bounds__(256) k_unpack_states(int n, const int *__restrict__ idx, BodyArrays B, const float4 *__restrict__ in) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = idx[i];
    const float4 p = in[4 * (size_t)i], q = in[4 * (size_t)i + 1], lv = in[4 * (size_t)i + 2], av = in[4 * (size_t)i + 3];
    const float invM = B.pos[b].w, mass = B.lvel[b].w;
    B.pos[b] = make_float4(p.x, p.y, p.z, invM);
    B.quat[b] = q;
    store_m3(B.R, b, q_to_r(q));
    B.lvel[b] = make_float4(lv.x, lv.y, lv.z, mass);
    B.avel[b] = make_float4(av.x, av.y, av.z, 0.f);
}

void eng_pack_states_device(Engine *e, const int *d_idx, int n, float *d_out) {
    eng_sync_to_device(e);
    if (n <= 0) return;
    k_pack_states<<<(unsigned)((n + 255) / 256), 256, 0, e->st>>>(n, d_idx, e->B, reinterpret_cast<float4 *>(d_out));
    OB_CHECK_KERNEL("k_pack_states", e->st);
}
// Constraint impulses of the last step, per listed body: 8 floats = h * fc (linear xyz, pad, angular xyz, pad),
// i.e. exactly the velocity change the step's contacts gave the body.  A slab that solves a cross-face contact
// against a dynamic ghost sends these to the ghost's owner, which adds them.
__global__ void __launch_bounds__(256) k_pack_impulses(int n, const int *__restrict__ idx, const float4 *__restrict__ fc, float h,
                                                        float4 *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = idx[i];
    if (b < 0) { // empty slot of a counted list
        out[2 * (size_t)i] = make_float4(0.f, 0.f, 0.f, 0.f);
        out[2 * (size_t)i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const float4 l = fc[2 * (size_t)b], a = fc[2 * (size_t)b + 1];
    out[2 * (size_t)i] = make_float4(h * l.x, h * l.y, h * l.z, 0.f);
    out[2 * (size_t)i + 1] = make_float4(h * a.x, h * a.y, h * a.z, 0.f);
}
__global__ void __launch_bounds__(256) k_add_impulses(int n, const int *__restrict__ idx, BodyArrays B, const float4 *__restrict__ in) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = idx[i];
    if (b < 0 || !(B.pos[b].w > 0.f)) return; // empty slots, kinematic / destroyed bodies take no impulse
    const float4 l = in[2 * (size_t)i], a = in[2 * (size_t)i + 1];
    float4 lv = B.lvel[b], av = B.avel[b];
    lv.x += l.x; lv.y += l.y; lv.z += l.z;
    av.x += a.x; av.y += a.y; av.z += a.z;
    B.lvel[b] = lv;
    B.avel[b] = av;
}
void eng_pack_impulses_device(Engine *e, const int *d_idx, int n, float *d_out) {
    if (!e->fc_valid) {
        fprintf(stderr, "libode_b200: dWorldPackImpulsesDeviceB200: no accumulators from the last step (call "
                        "dWorldSetKeepImpulsesB200(world, 1) before stepping batched worlds)\n");
        abort();
    }
    if (n <= 0) return;
    k_pack_impulses<<<(unsigned)((n + 255) / 256), 256, 0, e->st>>>(n, d_idx, e->B.fc, e->last_h, reinterpret_cast<float4 *>(d_out));
    OB_CHECK_KERNEL("k_pack_impulses", e->st);
}
void eng_add_impulses_device(Engine *e, const int *d_idx, int n, const float *d_in) {
    eng_sync_to_device(e);
    if (n <= 0) return;
    k_add_impulses<<<(unsigned)((n + 255) / 256), 256, 0, e->st>>>(n, d_idx, e->B, reinterpret_cast<const float4 *>(d_in));
    OB_CHECK_KERNEL("k_add_impulses", e->st);
    e->host_stale = true;
}
void eng_set_keep_impulses(Engine *e, int on) { e->keep_fc = on != 0; }

// ---- dynamic halo: which bodies sit near a slab face changes as the pile moves, so the list is rebuilt on the
// device every tick (flag -> scan -> compact: ascending, deterministic), and the message carries whole bodies
// (state + mass properties + the shape of the body's geom) into a pool of ghost slots on the other side.
__global__ void __launch_bounds__(256) k_select_flag(BodyArrays B, int axis, float lo, float hi, const int *__restrict__ mask,
                                                      int *__restrict__ flag) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B.n) return;
    const float4 p = B.pos[b];
    const float x = axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
    flag[b] = ((!mask || mask[b]) && x >= lo && x < hi) ? 1 : 0;
}
__global__ void __launch_bounds__(256) k_select_write(int n, const int *__restrict__ off, const int *__restrict__ total, int cap,
                                                       const float4 *__restrict__ pos, int axis, float lo, float hi,
                                                       const int *__restrict__ mask, int *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float4 p = pos[i];
        const float x = axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
        if ((!mask || mask[i]) && x >= lo && x < hi && off[i] < cap) out[off[i]] = i;
    }
    if (i < cap && i >= *total) out[i] = -1;
}
void eng_select_bodies_device(Engine *e, int axis, float lo, float hi, const int *d_mask, int *d_idx_out, int cap, int *d_count) {
    eng_sync_to_device(e);
    const int n = e->B.n;
    if (n + 1 > e->cap_sel) {
        dev_realloc(e->sel_flag, 0, (size_t)n + (size_t)n / 4 + 65, e->st, false);
        e->cap_sel = n + n / 4 + 64;
    }
    const int m = std::max(n, cap);
    if (m <= 0) return;
    if (n > 0) {
        k_select_flag<<<(unsigned)((n + 255) / 256), 256, 0, e->st>>>(e->B, axis, lo, hi, d_mask, e->sel_flag);
        OB_CHECK_KERNEL("k_select_flag", e->st);
        scan_exclusive(e->sel_flag, e->sel_flag, n, nullptr, d_count, e->scan, e->st);
    } else {
        OB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), e->st));
    }
    k_select_write<<<(unsigned)((m + 255) / 256), 256, 0, e->st>>>(n, e->sel_flag, d_count, cap, e->B.pos, axis, lo, hi, d_mask, d_idx_out);
    OB_CHECK_KERNEL("k_select_write", e->st);
}

// 12 float4 per body: pos+invM | quat | lvel+mass | avel+geom type (int bits, -1: empty slot) | geom dims |
// I rows (3) | invI rows (3) | body flags (int bits)
__global__ void __launch_bounds__(128) k_pack_bodies(int cap, const int *__restrict__ idx, const int *__restrict__ body_geom,
                                                      BodyArrays B, GeomArrays G, float4 *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    float4 *o = out + 12 * (size_t)i;
    const int b = idx[i];
    const int g = b >= 0 ? body_geom[b] : -1;
    if (b < 0 || g < 0) {
        o[3] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
        return;
    }
    o[0] = B.pos[b]; o[1] = B.quat[b]; o[2] = B.lvel[b];
    const float4 av = B.avel[b];
    o[3] = make_float4(av.x, av.y, av.z, __int_as_float(G.type[g]));
    o[4] = G.dims[g];
    for (int k = 0; k < 3; k++) { o[5 + k] = B.I[3 * (size_t)b + k]; o[8 + k] = B.invI[3 * (size_t)b + k]; }
    o[11] = make_float4(__int_as_float(B.flags[b]), 0.f, 0.f, 0.f);
}
__global__ void __launch_bounds__(128) k_unpack_bodies(int cap, const int *__restrict__ ghost_body, const int *__restrict__ ghost_geom,
                                                        BodyArrays B, GeomArrays G, const float4 *__restrict__ in) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    const float4 *r = in + 12 * (size_t)i;
    const int b = ghost_body[i], g = ghost_geom[i];
    const float4 r3 = r[3];
    const int type = __float_as_int(r3.w);
    if (type < 0) { // unused slot: no geom, inert body
        G.alive[g] = 0;
        float4 p = B.pos[b];
        p.w = 0.f;
        B.pos[b] = p;
        B.flags[b] = BF_KINEMATIC | BF_NOGRAVITY;
        const float4 lv = B.lvel[b];
        B.lvel[b] = make_float4(0.f, 0.f, 0.f, lv.w);
        B.avel[b] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    B.pos[b] = r[0];
    B.quat[b] = r[1];
    store_m3(B.R, b, q_to_r(r[1]));
    B.lvel[b] = r[2];
    B.avel[b] = make_float4(r3.x, r3.y, r3.z, 0.f);
    for (int k = 0; k < 3; k++) { B.I[3 * (size_t)b + k] = r[5 + k]; B.invI[3 * (size_t)b + k] = r[8 + k]; }
    B.flags[b] = __float_as_int(r[11].x);
    G.type[g] = type;
    G.dims[g] = r[4];
    G.alive[g] = 1;
}
void eng_pack_bodies_device(Engine *e, const int *d_idx, int cap, const int *d_body_geom, float *d_out) {
    eng_sync_to_device(e);
    if (cap <= 0) return;
    k_pack_bodies<<<(unsigned)((cap + 127) / 128), 128, 0, e->st>>>(cap, d_idx, d_body_geom, e->B, e->G, reinterpret_cast<float4 *>(d_out));
    OB_CHECK_KERNEL("k_pack_bodies", e->st);
}
void eng_unpack_bodies_device(Engine *e, const int *d_ghost_body, const int *d_ghost_geom, int cap, const float *d_in) {
    eng_sync_to_device(e);
    if (cap <= 0) return;
    k_unpack_bodies<<<(unsigned)((cap + 127) / 128), 128, 0, e->st>>>(cap, d_ghost_body, d_ghost_geom, e->B, e->G,
                                                                     reinterpret_cast<const float4 *>(d_in));
    OB_CHECK_KERNEL("k_unpack_bodies", e->st);
    e->host_stale = true;
}
void eng_unpack_states_device(Engine *e, const int *d_idx, int n, const float *d_in) {
    eng_sync_to_device(e);
    if (n <= 0) return;
    k_unpack_states<<<(unsigned)((n + 255) / 256), 256, 0, e->st>>>(n, d_idx, e->B, reinterpret_cast<const float4 *>(d_in));
    OB_CHECK_KERNEL("k_unpack_states", e->st);
    e->host_stale = true;
}

// ---- wire image of the reference's MsgUpdateBodies (inc/msgs.h:30-33): int msg type, then n_slots BodyState
// records of 84 bytes {int type; float transform[16]; float size[3]; uchar col[4]} (inc/body.h:26-31).
// Slot -> body (transform from the fused snapshot) or static geom (GetTransformMat of its pose).
__global__ void __launch_bounds__(256) k_pack_msg(int n_slots, const int *__restrict__ slot_body, const int *__restrict__ slot_geom,
                                                   const int *__restrict__ slot_type, const float *__restrict__ slot_size,
                                                   const unsigned *__restrict__ slot_col, const float4 *__restrict__ b_pos,
                                                   const float4 *__restrict__ b_R, const float4 *__restrict__ g_pos, const float4 *__restrict__ g_R, int msg_type,
                                                   unsigned *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) out[0] = (unsigned)msg_type;
    if (i >= n_slots) return;
    unsigned *rec = out + 1 + (size_t)i * 21; // 84 bytes = 21 words
    const int type = slot_type[i];
    rec[0] = (unsigned)type;
    float t[16];
    for (int k = 0; k < 16; k++) t[k] = 0.f;
    if (type != 0) {
        const int b = slot_body[i], g = slot_geom[i];
        if (b >= 0 || g >= 0) { // GetTransformMat(pos, rot): columns of the result = rows of ODE's R
            // (from the body state itself, not the snapshot buffer: valid for bodies spawned since the last step and
            // whatever the snapshot format)
            const float4 p = b >= 0 ? b_pos[b] : g_pos[g];
            const float4 r0 = b >= 0 ? b_R[3 * b] : g_R[3 * g], r1 = b >= 0 ? b_R[3 * b + 1] : g_R[3 * g + 1],
                         r2 = b >= 0 ? b_R[3 * b + 2] : g_R[3 * g + 2];
            t[0] = r0.x; t[1] = r1.x; t[2] = r2.x;
            t[4] = r0.y; t[5] = r1.y; t[6] = r2.y;
            t[8] = r0.z; t[9] = r1.z; t[10] = r2.z;
            t[12] = p.x; t[13] = p.y; t[14] = p.z; t[15] = 1.f;
        }
    }
    for (int k = 0; k < 16; k++) rec[1 + k] = __float_as_uint(t[k]);
    for (int k = 0; k < 3; k++) rec[17 + k] = __float_as_uint(slot_size[3 * i + k]);
    rec[20] = slot_col[i];
}

void eng_bind_msg_slots(Engine *e, int n_slots, const int *body, const int *geom, const int *type, const float *size3,
                        const unsigned *rgba) {
    OB_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = e->st;
    OB_CUDA(cudaStreamSynchronize(st));
    dev_free(e->msg_body); dev_free(e->msg_geom); dev_free(e->msg_type); dev_free(e->msg_size); dev_free(e->msg_col);
    dev_free(e->msg_out);
    e->msg_slots = n_slots;
    if (n_slots <= 0) return;
    const size_t n = (size_t)n_slots;
    OB_CUDA(ob_malloc(&e->msg_body, n * 4)); OB_CUDA(ob_malloc(&e->msg_geom, n * 4)); OB_CUDA(ob_malloc(&e->msg_type, n * 4));
    OB_CUDA(ob_malloc(&e->msg_size, n * 12)); OB_CUDA(ob_malloc(&e->msg_col, n * 4));
    OB_CUDA(ob_malloc(&e->msg_out, 4 + n * 84));
    OB_CUDA(cudaMemcpy(e->msg_body, body, n * 4, cudaMemcpyHostToDevice));
    OB_CUDA(cudaMemcpy(e->msg_geom, geom, n * 4, cudaMemcpyHostToDevice));
    OB_CUDA(cudaMemcpy(e->msg_type, type, n * 4, cudaMemcpyHostToDevice));
    OB_CUDA(cudaMemcpy(e->msg_size, size3, n * 12, cudaMemcpyHostToDevice));
    OB_CUDA(cudaMemcpy(e->msg_col, rgba, n * 4, cudaMemcpyHostToDevice));
}

size_t eng_pack_msg(Engine *e, void *dst, int msg_type, bool blocking) {
    if (e->msg_slots <= 0) return 0;
    eng_sync_to_device(e);
    const size_t bytes = 4 + (size_t)e->msg_slots * 84;
    k_pack_msg<<<(unsigned)((e->msg_slots + 255) / 256), 256, 0, e->st>>>(e->msg_slots, e->msg_body, e->msg_geom, e->msg_type,
                                                                       e->msg_size, e->msg_col, e->B.pos, e->B.R, e->G.pos,
                                                                       e->G.R, msg_type, e->msg_out);
    OB_CHECK_KERNEL("k_pack_msg", e->st);
    OB_CUDA(cudaMemcpyAsync(dst, e->msg_out, bytes, cudaMemcpyDeviceToHost, e->st));
    if (blocking) OB_CUDA(cudaStreamSynchronize(e->st));
    return bytes;
}

float eng_barrier_bench(Engine *e, int iters) { return solver_barrier_bench(e, iters); }

long g_ob_launches = 0;
long eng_launch_count() { return g_ob_launches; }

void eng_timer_start(Engine *e) {
    OB_CUDA(cudaSetDevice(e->device));
    if (!e->tev[0]) { OB_CUDA(cudaEventCreate(&e->tev[0])); OB_CUDA(cudaEventCreate(&e->tev[1])); }
    OB_CUDA(cudaEventRecord(e->tev[0], e->st));
}
void eng_timer_stop(Engine *e) {
    if (e->tev[1]) OB_CUDA(cudaEventRecord(e->tev[1], e->st));
}
float eng_timer_elapsed_ms(Engine *e) {
    float ms = 0.f;
    if (!e->tev[0]) return -1.f;
    OB_CUDA(cudaEventSynchronize(e->tev[1]));
    OB_CUDA(cudaEventElapsedTime(&ms, e->tev[0], e->tev[1]));
    return ms;
}

float eng_timer_elapsed_between_ms(Engine *start, Engine *stop) {
    float ms = 0.f;
    if (!start->tev[0] || !stop->tev[1]) return -1.f; // a timer that was never started
    OB_CUDA(cudaEventSynchronize(stop->tev[1]));
    OB_CUDA(cudaEventElapsedTime(&ms, start->tev[0], stop->tev[1]));
    return ms;
}

void eng_wait(Engine *e) {
    OB_CUDA(cudaSetDevice(e->device));
    OB_CUDA(cudaStreamSynchronize(e->st));
    OB_CUDA(cudaStreamSynchronize(e->copy_st));
    OB_CUDA(cudaStreamSynchronize(e->h2d_st));
}

StepStats eng_stats(Engine *e) {
    OB_CUDA(cudaSetDevice(e->device));
    OB_CUDA(cudaMemcpyAsync(e->h_stats, e->d_stats, sizeof(StepStats), cudaMemcpyDeviceToHost, e->st));
    OB_CUDA(cudaStreamSynchronize(e->st));
    return *e->h_stats;
}

void eng_last_timings(Engine *e, float out[4]) {
    out[0] = out[1] = out[2] = out[3] = 0.f;
    if (!e->ev_valid) return;
    OB_CUDA(cudaSetDevice(e->device));
    OB_CUDA(cudaEventSynchronize(e->ev[4]));
    cudaEventElapsedTime(&out[0], e->ev[0], e->ev[1]); // collide
    cudaEventElapsedTime(&out[1], e->ev[1], e->ev[3]); // prep + manifolds + colouring + rows
    cudaEventElapsedTime(&out[2], e->ev[3], e->ev[4]); // solve + integrate + pack
    cudaEventElapsedTime(&out[3], e->ev[0], e->ev[4]); // whole tick
}

// broadphase, narrowphase, prepare, solve, whole tick (ms) of the last tick run with timing on
void eng_stage_timings(Engine *e, float out[5]) {
    for (int i = 0; i < 5; i++) out[i] = 0.f;
    if (!e->ev_valid) return;
    OB_CUDA(cudaSetDevice(e->device));
    OB_CUDA(cudaEventSynchronize(e->ev[4]));
    if (cudaEventElapsedTime(&out[0], e->ev[0], e->ev_bp) != cudaSuccess) { out[0] = 0.f; (void)cudaGetLastError(); }
    if (cudaEventElapsedTime(&out[1], e->ev_bp, e->ev[1]) != cudaSuccess) { out[1] = 0.f; (void)cudaGetLastError(); }
    cudaEventElapsedTime(&out[2], e->ev[1], e->ev[3]);
    cudaEventElapsedTime(&out[3], e->ev[3], e->ev[4]);
    cudaEventElapsedTime(&out[4], e->ev[0], e->ev[4]);
}

int eng_export_solver_order(Engine *e, int *pair_g1, int *pair_g2, int *pair_k, int cap) {
    OB_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = e->st;
    OB_CUDA(cudaStreamSynchronize(st));
    int nm = 0;
    OB_CUDA(cudaMemcpy(&nm, e->M.count, sizeof(int), cudaMemcpyDeviceToHost));
    if (nm <= 0) return 0;
    std::vector<int4> mrec((size_t)nm);
    OB_CUDA(cudaMemcpy(mrec.data(), e->S.mrec, (size_t)nm * sizeof(int4), cudaMemcpyDeviceToHost));
    BroadCounters bc;
    OB_CUDA(cudaMemcpy(&bc, e->bp.counters, sizeof(bc), cudaMemcpyDeviceToHost));
    std::vector<int2> pairs((size_t)std::max(bc.n_pairs, 1));
    if (bc.n_pairs) OB_CUDA(cudaMemcpy(pairs.data(), e->bp.pairs, (size_t)bc.n_pairs * sizeof(int2), cudaMemcpyDeviceToHost));
    int out = 0;
    for (int s = 0; s < nm; s++) {
        // .w = slot index of the unit's first contact = pair + k0 * stride
        const int stride = e->cs.stride > 0 ? e->cs.stride : 1;
        const int p = mrec[s].w % stride, k0 = mrec[s].w / stride, nc = mrec[s].z;
        for (int k = 0; k < nc; k++) {
            if (out < cap) {
                pair_g1[out] = (p < bc.n_pairs) ? pairs[p].x : -1;
                pair_g2[out] = (p < bc.n_pairs) ? pairs[p].y : -1;
                pair_k[out] = k0 + k;
            }
            out++;
        }
    }
    return out;
}

} // namespace ob
