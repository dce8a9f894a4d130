// engine_impl.h -- the Engine object: host mirrors + device arrays of one world.
#pragma once

#include "dev.cuh"

namespace ob {

struct Engine {
    int device = 0;
    int num_sms = 148;
    cudaStream_t st = nullptr;
    cudaStream_t copy_st = nullptr;
    WorldParams params;

    HostBodies hb;
    HostGeoms hg;
    bool bodies_dirty = false, geoms_dirty = false, forces_dirty = false;
    bool host_stale = false; // device state is newer than the host mirrors
    bool keep_fc = false, fc_valid = false; // the last step left the bodies' accumulators fc in global memory
    float last_h = 0.f;
    int *sel_flag = nullptr; // scratch of eng_select_bodies_device
    int cap_sel = 0;
    // incremental ingestion (spawns, per-tick setters): dirty lists + per-entry field masks
    std::vector<int> dirty_b, dirty_g, force_b;
    std::vector<unsigned char> mask_b, mask_g, inforce_b, alive_dev;
    int n_b_dev = 0, n_g_dev = 0;   // bodies / geoms the device already holds
    std::vector<int> env_first, env_cnt, envg_first, envg_cnt, envg_shared; // persistent per-env ranges
    int env_max_local = 0, envg_max = 0, envg_alive = 0;
    bool env_contig = true, envg_ok = true;
    void *h_patch = nullptr, *d_patch = nullptr;
    size_t cap_patch = 0;
    cudaEvent_t ev_patch = nullptr;
    bool patch_inflight = false;
    int n_envs = 1;
    float big_extent = INFINITY;

    int cap_b = 0, cap_g = 0;
    BodyArrays B{};
    GeomArrays G{};
    MeshTable meshes{};
    std::vector<TriMesh> hmeshes;

    long want_pairs = 0, want_manifolds = 0; // user capacities (0 = auto)
    BroadPhase bp;
    ContactSlots cs{};
    int max_contacts = 8;
    bool have_device_contacts = false;

    ManifoldArrays M{};
    EnvArrays E{};
    int cap_envs = 0, cap_env_rec = 0, cap_env_bodies = 0;
    int env_group = 0; // lanes per env of the island solver (0 = automatic)
    // wire image of MsgUpdateBodies (slot table bound by the host)
    int msg_slots = 0;
    int *msg_body = nullptr, *msg_geom = nullptr, *msg_type = nullptr;
    float *msg_size = nullptr;
    unsigned *msg_col = nullptr, *msg_out = nullptr;
    float4 *body_hot = nullptr; // fc (2 per body) followed by inv (3 per body)
    int l2_persist = 0;        // measured: the persisting carve-out costs more than it gives (profiles/README.md)        // keep body_hot resident in L2 through an access-policy window
    // CUDA graphs of the device-resident tick of batched worlds (collide; step, one per snapshot buffer)
    struct TickGraph {
        cudaGraphExec_t exec = nullptr;
        unsigned long long key = 0, warm_key = 0; // configuration captured / seen on the last ticks
        int warm = 0, kernels = 0;
        bool fc_valid = false;
    };
    TickGraph g_collide, g_step[2];
    int graphs = 1;
    bool tiny_attr_set = false; // k_tiny_solve's dynamic shared memory limit raised on this engine's device
    int tiny_solver = 1; // small single worlds: one-CTA shared-memory solver (k_tiny_solve) before k_solve
    int env_fuse = 1;  // island solver: run body preparation and the integrate/pack tail inside k_env_solve
    int env_stage = 1; // island solver: stage body data in shared memory when possible
    int env_pair = 1;      // island solver: lane-pair kernel (solver_env.cu) for per-contact units on contiguous envs
    int env_pair_rows = 0; // ... with the first N row records of an env in shared memory (0: rows stay in global memory / L2)
    int solver_mode = 0; // 0 automatic, 1 force the global (grid-barrier) solver
    int contact_units = -1; // -1 automatic (per contact for batched worlds), 0 manifold units, 1 contact units
    int broad_mode = -1;   // -1 auto, 0 uniform grid, 1 all pairs per env
    EnvBroad EB;
    int cap_eb_envs = 0, cap_eb_shared = 0;
    int colour_spread = 0; // 0: lowest free colour; K > 0: hashed start within the first K colours
    bool colour_spread_auto = true;
    SolverArrays S{};
    ScanWorkspace scan;
    SortWorkspace sort;

    // compat-mode contact upload buffers (device) + pinned staging
    int cap_hc = 0;
    float4 *hc_pd = nullptr, *hc_ns = nullptr;
    Surface *hc_surf = nullptr;
    int4 *hc_mrec = nullptr;
    std::vector<float4> st_pd, st_ns;
    std::vector<Surface> st_surf;
    std::vector<int4> st_mrec;
    unsigned char *hcs_host = nullptr, *hcs_dev = nullptr; // one pinned staging blob + its device twin: the step's joints in ONE copy
    size_t hcs_cap = 0;

    // compat-mode pair download
    int cap_dl = 0, cap_dlc = 0;
    int *dl_first = nullptr;     // device: exclusive scan of nc
    float4 *dl_pd = nullptr, *dl_ns = nullptr; // device: pair-major compacted contacts
    std::vector<int> h_g1, h_g2, h_first, h_count;
    std::vector<float> h_pd, h_ns;
    // ... fast path for small and medium worlds: ONE kernel writes pairs, contact ranges and contacts straight into a
    // mapped pinned host buffer (zero-copy), one synchronisation -- instead of three round trips with six small copies
    unsigned char *x_host = nullptr, *x_dev = nullptr; // the same buffer: host address / device address
    int x_cap_pairs = 0, x_cap_contacts = 0;
    int x_want_pairs = 1024, x_want_contacts = 2048; // applied at the next traversal (the current one's pointers stay valid)

    // exact dWorldStep of small worlds (solver_exact.cu): island tables, row descriptors, island matrices
    int *ex_label = nullptr, *ex_desc = nullptr, *ex_isl_row0 = nullptr, *ex_isl_label = nullptr, *ex_meta = nullptr;
    size_t *ex_isl_mat = nullptr;
    double *ex_A = nullptr, *ex_C = nullptr;
    size_t ex_cap_mat = 0;
    int ex_cap_rows = 0;

    StepStats *d_stats = nullptr;
    StepStats *h_stats = nullptr; // pinned
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_bp = nullptr; // end of the broadphase inside the collide stage
    bool timing = false;
    float last_ms[4] = {0, 0, 0, 0};
    bool ev_valid = false;

    int coop_blocks_colour = 0, coop_blocks_solve = 0;
    cudaEvent_t tev[2] = {nullptr, nullptr};
    // host <-> device traffic of the tick runs on its own streams so it overlaps the next tick's kernels:
    // snapshots are double-buffered (the tail of tick t+1 writes the other buffer while tick t's copy drains)
    cudaStream_t h2d_st = nullptr;
    float *snap_buf[2] = {nullptr, nullptr};
    int snap_cur = 0;                 // buffer the last step wrote (B.snap points at it)
    int snap_fmt = 0;                 // record format (dWorldSetSnapshotFormatB200)
    bool snap_stale = false;          // bodies were spawned / moved by the host since the records were written
    cudaEvent_t ev_step_done = nullptr;
    cudaEvent_t ev_snap_copied[2] = {nullptr, nullptr};
    bool snap_copy_pending[2] = {false, false};
    float *d_f6[2] = {nullptr, nullptr};   // double-buffered staging of the 6-float force records
    int f6_cur = 0;
    int cap_f6 = 0, pending_f6 = 0;   // pending_f6 > 0: forces uploaded on h2d_st, scatter before the next step
    cudaEvent_t ev_f6 = nullptr;
    cudaEvent_t ev_f6_consumed[2] = {nullptr, nullptr};
    bool f6_inflight[2] = {false, false};
};

void engine_ensure_capacity(Engine *e);
void engine_ensure_pair_capacity(Engine *e);

} // namespace ob
