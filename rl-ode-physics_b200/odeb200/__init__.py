"""ctypes binding of libode_b200.so (the C ABI declared in include/ode/ode.h and include/ode_b200.h).

This is host-side glue for tests and the benchmark: every physics call goes through the C ABI of the
shared library, whose kernels are hand-written CUDA for sm_100a.  There is no CPU path: loading
fails loudly when the library is missing, and creating a world aborts when no GPU is present.
"""
import ctypes as C
import os

import numpy as np

from . import scenes  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "..", "libode_b200.so")
_LIB = None

dContactBounce = 0x004
INF = float("inf")


class SurfaceParameters(C.Structure):
    _fields_ = [("mode", C.c_int), ("mu", C.c_float), ("mu2", C.c_float), ("rho", C.c_float), ("rho2", C.c_float),
                ("rhoN", C.c_float), ("bounce", C.c_float), ("bounce_vel", C.c_float), ("soft_erp", C.c_float),
                ("soft_cfm", C.c_float), ("motion1", C.c_float), ("motion2", C.c_float), ("motionN", C.c_float),
                ("slip1", C.c_float), ("slip2", C.c_float)]


class ContactGeom(C.Structure):
    _fields_ = [("pos", C.c_float * 4), ("normal", C.c_float * 4), ("depth", C.c_float), ("g1", C.c_void_p),
                ("g2", C.c_void_p), ("side1", C.c_int), ("side2", C.c_int)]


class Contact(C.Structure):
    _fields_ = [("surface", SurfaceParameters), ("geom", ContactGeom), ("fdir1", C.c_float * 4)]


class Mass(C.Structure):
    _fields_ = [("mass", C.c_float), ("c", C.c_float * 4), ("I", C.c_float * 12)]


class StepStats(C.Structure):
    _fields_ = [("n_geoms", C.c_int), ("n_big", C.c_int), ("n_pairs", C.c_int), ("n_contacts", C.c_int),
                ("n_manifolds", C.c_int), ("n_colours", C.c_int), ("n_overflow", C.c_int), ("flags", C.c_int),
                ("class_count", C.c_int * 7), ("n_rows", C.c_int), ("n_rows1", C.c_int), ("n_rows2", C.c_int),
                ("colour_rounds", C.c_int), ("cell_size", C.c_float), ("grid_dims", C.c_int * 3),
                ("solver_iters", C.c_int), ("exact_status", C.c_int), ("n_islands", C.c_int), ("max_island_rows", C.c_int),
                ("pivot_rounds", C.c_int), ("env_trips", C.c_int), ("env_lanes", C.c_int)]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if hasattr(v, "__len__") else v
        return d


class SlabLayout(C.Structure):
    _fields_ = [("face_left", C.c_float), ("face_right", C.c_float), ("margin", C.c_float), ("hyst", C.c_float),
                ("n_own", C.c_int), ("n_static", C.c_int), ("pool", C.c_int), ("pool_first_body", C.c_int),
                ("pool_first_geom", C.c_int), ("mig_cap", C.c_int)]


class SlabInfo(C.Structure):
    _fields_ = [("n_owned", C.c_int), ("halo_selected", C.c_int), ("halo_overflow", C.c_int), ("mig_overflow", C.c_int),
                ("migrated_in", C.c_long), ("migrated_out", C.c_long), ("ticks", C.c_long), ("halo_bytes_per_tick", C.c_long)]


NearCallback = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p)

# (name, restype, argtypes) for every exported entry point the binding uses
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_up = C.POINTER(C.c_uint)
_vp = C.c_void_p
_f = C.c_float
_i = C.c_int
_SIGS = [
    ("dInitODE", None, []), ("dCloseODE", None, []),
    ("dWorldCreate", _vp, []), ("dWorldDestroy", None, [_vp]),
    ("dWorldSetGravity", None, [_vp, _f, _f, _f]),
    ("dWorldSetERP", None, [_vp, _f]), ("dWorldSetCFM", None, [_vp, _f]),
    ("dWorldSetQuickStepNumIterations", None, [_vp, _i]), ("dWorldSetQuickStepW", None, [_vp, _f]),
    ("dWorldGetQuickStepNumIterations", _i, [_vp]),
    ("dWorldStep", _i, [_vp, _f]), ("dWorldQuickStep", _i, [_vp, _f]),
    ("dBodyCreate", _vp, [_vp]), ("dBodyDestroy", None, [_vp]),
    ("dBodySetPosition", None, [_vp, _f, _f, _f]), ("dBodySetRotation", None, [_vp, _fp]),
    ("dBodySetQuaternion", None, [_vp, _fp]), ("dBodySetLinearVel", None, [_vp, _f, _f, _f]),
    ("dBodySetAngularVel", None, [_vp, _f, _f, _f]),
    ("dBodyGetPosition", _fp, [_vp]), ("dBodyGetRotation", _fp, [_vp]), ("dBodyGetQuaternion", _fp, [_vp]),
    ("dBodyGetLinearVel", _fp, [_vp]), ("dBodyGetAngularVel", _fp, [_vp]),
    ("dBodySetMass", None, [_vp, C.POINTER(Mass)]), ("dBodySetKinematic", None, [_vp]),
    ("dBodySetGyroscopicMode", None, [_vp, _i]), ("dBodyAddForce", None, [_vp, _f, _f, _f]),
    ("dMassSetBox", None, [C.POINTER(Mass), _f, _f, _f, _f]), ("dMassSetSphere", None, [C.POINTER(Mass), _f, _f]),
    ("dMassSetBoxTotal", None, [C.POINTER(Mass), _f, _f, _f, _f]),
    ("dMassSetSphereTotal", None, [C.POINTER(Mass), _f, _f]),
    ("dHashSpaceCreate", _vp, [_vp]), ("dSpaceDestroy", None, [_vp]),
    ("dSpaceCollide", None, [_vp, _vp, NearCallback]),
    ("dCollide", _i, [_vp, _vp, _i, C.POINTER(ContactGeom), _i]),
    ("dCreateSphere", _vp, [_vp, _f]), ("dCreateBox", _vp, [_vp, _f, _f, _f]),
    ("dCreatePlane", _vp, [_vp, _f, _f, _f, _f]), ("dGeomDestroy", None, [_vp]),
    ("dGeomSetBody", None, [_vp, _vp]), ("dGeomGetBody", _vp, [_vp]),
    ("dGeomSetPosition", None, [_vp, _f, _f, _f]), ("dGeomSetRotation", None, [_vp, _fp]),
    ("dGeomGetPosition", _fp, [_vp]), ("dGeomGetRotation", _fp, [_vp]),
    ("dGeomSetCategoryBits", None, [_vp, C.c_ulong]), ("dGeomSetCollideBits", None, [_vp, C.c_ulong]),
    ("dGeomTriMeshDataCreate", _vp, []), ("dGeomTriMeshDataDestroy", None, [_vp]),
    ("dGeomTriMeshDataBuildSingle", None, [_vp, _vp, _i, _i, _vp, _i, _i]),
    ("dCreateTriMesh", _vp, [_vp, _vp, _vp, _vp, _vp]),
    ("dGeomTriMeshDataBuildFromOBJB200", _i, [_vp, C.c_char_p]),
    ("dJointGroupCreate", _vp, [_i]), ("dJointGroupEmpty", None, [_vp]), ("dJointGroupDestroy", None, [_vp]),
    ("dJointCreateContact", _vp, [_vp, _vp, C.POINTER(Contact)]), ("dJointAttach", None, [_vp, _vp, _vp]),
    # extensions
    ("dSetDeviceB200", None, [_i]),
    ("dSpaceCollideDeviceB200", None, [_vp, _i]),
    ("dWorldSetSurfaceB200", None, [_vp, C.POINTER(SurfaceParameters)]),
    ("dWorldGetSurfaceB200", None, [_vp, C.POINTER(SurfaceParameters)]),
    ("dWorldSetMaxContactsB200", None, [_vp, _i]),
    ("dWorldSetNumEnvsB200", None, [_vp, _i]),
    ("dWorldSetSlotReuseB200", None, [_vp, _i]),
    ("dWorldAddBodiesB200", _i, [_vp, _i, _fp, _fp, _fp, _fp, _fp, _fp, _ip, _ip]),
    ("dSpaceAddGeomsB200", _i, [_vp, _vp, _i, _ip, _fp, _ip, _fp, _fp, _up, _up, _ip]),
    ("dWorldAddTriMeshB200", _i, [_vp, _fp, _i, _ip, _i]),
    ("dTestForceFullSyncB200", None, [_vp]),
    ("dWorldGetBodyB200", _vp, [_vp, _i]), ("dSpaceGetGeomB200", _vp, [_vp, _i]),
    ("dWorldGetNumBodiesB200", _i, [_vp]),
    ("dBodyGetIndexB200", _i, [_vp]), ("dGeomGetIndexB200", _i, [_vp]),
    ("dWorldGetStateB200", None, [_vp, _fp, _fp, _fp, _fp, _fp]),
    ("dWorldSetForcesB200", None, [_vp, _fp, _i]),
    ("dWorldGetSnapshotB200", None, [_vp, _vp, _i, _i, _i]),
    ("dWorldGetSnapshotDeviceB200", _vp, [_vp]),
    ("dWorldSetSnapshotFormatB200", None, [_vp, _i]), ("dWorldGetSnapshotFormatB200", _i, [_vp]),
    ("dSnapshotExpandB200", None, [_vp, _i, _i, _vp, _i]),
    ("dWorldGetStageTimingsB200", None, [_vp, _fp]),
    ("dGeomTriMeshDataGetB200", _i, [_vp, _fp, _i, _ip, _i, _ip]),
    ("dWorldWaitB200", None, [_vp]),
    ("dWorldPackStatesDeviceB200", None, [_vp, _vp, _i, _vp]), ("dWorldUnpackStatesDeviceB200", None, [_vp, _vp, _i, _vp]),
    ("dWorldPackImpulsesDeviceB200", None, [_vp, _vp, _i, _vp]), ("dWorldAddImpulsesDeviceB200", None, [_vp, _vp, _i, _vp]),
    ("dWorldSetKeepImpulsesB200", None, [_vp, _i]),
    ("dWorldSelectBodiesDeviceB200", None, [_vp, _i, _f, _f, _vp, _vp, _i, _vp]),
    ("dWorldPackBodiesDeviceB200", None, [_vp, _vp, _i, _vp, _vp]),
    ("dWorldUnpackBodiesDeviceB200", None, [_vp, _vp, _vp, _i, _vp]),
    ("dWorldGetStreamB200", _vp, [_vp]),
    ("dCheckGuardsB200", _i, [_i]),
    ("dWorldGetDeviceB200", _i, [_vp]),
    ("dAllocPinnedB200", _vp, [C.c_size_t, _i]), ("dFreePinnedB200", None, [_vp]),
    ("dGuardSelfTestB200", _i, [_i]),
    ("dSlabGetUniqueIdB200", _i, [C.c_char_p]),
    ("dSlabCreateB200", _vp, [_vp, _vp, _i, _i, C.c_char_p, C.POINTER(SlabLayout)]),
    ("dSlabDestroyB200", None, [_vp]), ("dSlabTickB200", None, [_vp, _f, _i]), ("dSlabMigrateB200", None, [_vp]),
    ("dSlabConnectLocalB200", None, [_vp, _vp]), ("dSlabTickLocalB200", None, [C.POINTER(_vp), _i, _f, _i]),
    ("dSlabMigrateLocalB200", None, [C.POINTER(_vp), _i]), ("dSlabGetInfoB200", None, [_vp, C.POINTER(SlabInfo)]),
    ("dWorldTimerStartB200", None, [_vp]), ("dWorldTimerStopB200", None, [_vp]),
    ("dWorldTimerElapsedB200", C.c_float, [_vp]), ("dGetKernelLaunchCountB200", C.c_long, []),
    ("dWorldTimerElapsedBetweenB200", C.c_float, [_vp, _vp]),
    ("dWorldSetCapacityB200", None, [_vp, C.c_long, C.c_long]),
    ("dWorldSetBigExtentB200", None, [_vp, _f]),
    ("dWorldSetBroadphaseB200", None, [_vp, _i]),
    ("dWorldSetStepSolverB200", None, [_vp, _i, _f]),
    ("dWorldSetSolverModeB200", None, [_vp, _i, _i]),
    ("dWorldSetContactUnitsB200", None, [_vp, _i]),
    ("dWorldGetStatsB200", None, [_vp, C.POINTER(StepStats)]),
    ("dWorldEnableTimingB200", None, [_vp, _i]),
    ("dWorldGetTimingsB200", None, [_vp, _fp]),
    ("dSpaceGetPairsB200", _i, [_vp, _ip, _i]),
    ("dSpaceGetContactsB200", _i, [_vp, _ip, _i, _fp, _fp, _i]),
    ("dWorldGetSolverOrderB200", _i, [_vp, _ip, _ip, _ip, _i]),
]


def lib():
    """Load libode_b200.so (built in-tree by __graft_entry__.build()). Fails loudly if absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.abspath(os.environ.get("ODE_B200_LIB", LIB_PATH))  # override only for instrumented builds
    if not os.path.exists(path):
        raise RuntimeError("libode_b200.so is not built (%s): run `python __graft_entry__.py` or make -C "
                           "rl-ode-physics_b200/csrc; there is no CPU fallback" % path)
    L = C.CDLL(path)
    for name, res, args in _SIGS:
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = L
    return L


def _a(x, dtype):
    return None if x is None else np.ascontiguousarray(x, dtype=dtype)


def _p(x, ptype):
    return None if x is None else x.ctypes.data_as(ptype)


def reference_surface():
    """The reference NearCallback's surface (src/main.c:684-687)."""
    s = SurfaceParameters()
    s.mode = dContactBounce
    s.bounce = 0.2
    s.bounce_vel = 0.1
    s.mu = INF
    return s


class World:
    """One dWorldID + one dSpaceID, driven through the device-resident path."""

    def __init__(self, gravity=(0.0, -9.8, 0.0), erp=None, cfm=None, iters=None, sor_w=None, device=None):
        self.L = lib()
        if device is not None:
            self.L.dSetDeviceB200(int(device))
        self.L.dInitODE()
        self.w = C.c_void_p(self.L.dWorldCreate())
        self.space = C.c_void_p(self.L.dHashSpaceCreate(None))
        self.L.dWorldSetGravity(self.w, *[float(g) for g in gravity])
        if erp is not None:
            self.L.dWorldSetERP(self.w, erp)
        if cfm is not None:
            self.L.dWorldSetCFM(self.w, cfm)
        if iters is not None:
            self.L.dWorldSetQuickStepNumIterations(self.w, iters)
        if sor_w is not None:
            self.L.dWorldSetQuickStepW(self.w, sor_w)
        self.n_bodies = 0
        self.n_geoms = 0

    def close(self):
        if getattr(self, "w", None):
            self.L.dSpaceDestroy(self.space)
            self.L.dWorldDestroy(self.w)
            self.w = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- construction
    def load_scene(self, sc):
        for v, t in sc.get("meshes", []):
            v = _a(v, np.float32)
            t = _a(t, np.int32)
            self.L.dWorldAddTriMeshB200(self.w, _p(v, _fp), len(v), _p(t, _ip), len(t))
        b = sc["bodies"]
        n = len(b["pos"])
        n_envs = int(max(1, (b["env"].max() + 1) if n else 1))
        self.L.dWorldSetNumEnvsB200(self.w, n_envs)
        arrs = [_a(b["pos"], np.float32), _a(b["quat"], np.float32), _a(b["lvel"], np.float32), _a(b["avel"], np.float32),
                _a(b["mass"], np.float32), _a(b.get("inertia"), np.float32)]
        fl, env = _a(b["flags"], np.int32), _a(b["env"], np.int32)
        if n:
            self.L.dWorldAddBodiesB200(self.w, n, *[_p(x, _fp) for x in arrs], _p(fl, _ip), _p(env, _ip))
        g = sc["geoms"]
        ng = len(g["type"])
        ga = [_a(g["type"], np.int32), _a(g["dims"], np.float32), _a(g["body"], np.int32), _a(g["pos"], np.float32),
              _a(g["R"], np.float32), _a(g["cat"], np.uint32), _a(g["col"], np.uint32), _a(g["env"], np.int32)]
        if ng:
            self.L.dSpaceAddGeomsB200(self.space, self.w, ng, _p(ga[0], _ip), _p(ga[1], _fp), _p(ga[2], _ip), _p(ga[3], _fp),
                                      _p(ga[4], _fp), _p(ga[5], _up), _p(ga[6], _up), _p(ga[7], _ip))
        self.n_bodies += n
        self.n_geoms += ng

    def spawn(self, pos, kind, dims, R=None, cat=2, col=3):
        """The reference's AddBody (src/main.c:695-733) through the handle API: one body + one geom."""
        L = self.L
        b = C.c_void_p(L.dBodyCreate(self.w))
        L.dBodySetPosition(b, float(pos[0]), float(pos[1]), float(pos[2]))
        if R is not None:
            r = np.ascontiguousarray(R, np.float32)
            L.dBodySetRotation(b, _p(r, _fp))
        if kind == "sphere":
            g = C.c_void_p(L.dCreateSphere(self.space, float(dims[0])))
        else:
            g = C.c_void_p(L.dCreateBox(self.space, float(dims[0]), float(dims[1]), float(dims[2])))
        L.dGeomSetBody(g, b)
        L.dGeomSetCategoryBits(g, cat)
        L.dGeomSetCollideBits(g, col)
        self.n_bodies += 1
        self.n_geoms += 1
        return b, g

    def body_handle(self, i):
        return C.c_void_p(self.L.dWorldGetBodyB200(self.w, int(i)))

    def force_full_sync(self):
        self.L.dTestForceFullSyncB200(self.w)

    def set_surface(self, s):
        self.L.dWorldSetSurfaceB200(self.w, C.byref(s))

    def set_solver_mode(self, mode=0, env_group=0):
        self.L.dWorldSetSolverModeB200(self.w, int(mode), int(env_group))

    def set_broadphase(self, mode):
        self.L.dWorldSetBroadphaseB200(self.w, int(mode))

    def set_contact_units(self, per_contact):
        self.L.dWorldSetContactUnitsB200(self.w, int(per_contact))

    def set_capacity(self, max_pairs, max_manifolds):
        self.L.dWorldSetCapacityB200(self.w, int(max_pairs), int(max_manifolds))

    # -- ticking
    def collide(self, max_contacts=8):
        self.L.dSpaceCollideDeviceB200(self.space, max_contacts)

    def step(self, h):
        return self.L.dWorldQuickStep(self.w, float(h))

    def set_step_solver(self, max_iters, tol=0.0):
        """dWorldStep parity mode: dWorldStep runs up to max_iters sweeps, stopping below tol."""
        self.L.dWorldSetStepSolverB200(self.w, int(max_iters), float(tol))

    def world_step(self, h):
        """dWorldStep (what the reference calls, src/main.c:213)."""
        return self.L.dWorldStep(self.w, float(h))

    def tick(self, h, max_contacts=8):
        """One reference tick (src/main.c:212-214), device-resident."""
        self.collide(max_contacts)
        return self.step(h)

    def wait(self):
        self.L.dWorldWaitB200(self.w)

    # -- readback
    def state(self):
        n = self.L.dWorldGetNumBodiesB200(self.w)
        pos = np.zeros((n, 3), np.float32); q = np.zeros((n, 4), np.float32)
        lv = np.zeros((n, 3), np.float32); av = np.zeros((n, 3), np.float32); R = np.zeros((n, 12), np.float32)
        self.L.dWorldGetStateB200(self.w, _p(pos, _fp), _p(q, _fp), _p(lv, _fp), _p(av, _fp), _p(R, _fp))
        return {"pos": pos, "quat": q, "lvel": lv, "avel": av, "R": R}

    SNAP_FLOATS = {0: 16, 1: 12, 2: 8}

    def set_snapshot_format(self, fmt):
        """0: GetTransformMat's 16 floats; 1: its 12 non-constant floats; 2: position + quaternion (8 floats)."""
        self.L.dWorldSetSnapshotFormatB200(self.w, int(fmt))

    def snapshot(self, first=0, count=None):
        """Snapshot records of bodies [first, first + count) in the world's current snapshot format."""
        n = self.L.dWorldGetNumBodiesB200(self.w)
        count = n - first if count is None else count
        out = np.zeros((count, self.SNAP_FLOATS[self.L.dWorldGetSnapshotFormatB200(self.w)]), np.float32)
        self.L.dWorldGetSnapshotB200(self.w, out.ctypes.data_as(C.c_void_p), first, count, 1)
        return out

    def expand_snapshot(self, compact, fmt, threads=4):
        """dSnapshotExpandB200: compact records -> the reference's 16-float transforms (host side)."""
        compact = np.ascontiguousarray(compact, np.float32)
        out = np.zeros((len(compact), 16), np.float32)
        self.L.dSnapshotExpandB200(compact.ctypes.data_as(C.c_void_p), int(fmt), len(compact), out.ctypes.data_as(C.c_void_p), threads)
        return out

    def stats(self):
        s = StepStats()
        self.L.dWorldGetStatsB200(self.w, C.byref(s))
        return s.as_dict()

    def enable_timing(self, on=True):
        self.L.dWorldEnableTimingB200(self.w, 1 if on else 0)

    def stage_timings(self):
        t = np.zeros(5, np.float32)
        self.L.dWorldGetStageTimingsB200(self.w, _p(t, _fp))
        return {"broadphase_ms": float(t[0]), "narrowphase_ms": float(t[1]), "prepare_ms": float(t[2]), "solve_ms": float(t[3]),
                "tick_ms": float(t[4])}

    def timings(self):
        t = np.zeros(4, np.float32)
        self.L.dWorldGetTimingsB200(self.w, _p(t, _fp))
        return {"collide_ms": float(t[0]), "prepare_ms": float(t[1]), "solve_ms": float(t[2]), "tick_ms": float(t[3])}

    def pairs(self):
        """Broadphase pair list of the last collide as a sorted (min id, max id) array."""
        n = self.L.dSpaceGetPairsB200(self.space, None, 0)
        out = np.zeros((max(n, 1), 2), np.int32)
        self.L.dSpaceGetPairsB200(self.space, _p(out, _ip), n)
        return out[:n]

    def contacts(self):
        """(pairs[n,2] in device order, counts[n], pos_depth[c,4], normal[c,3], side[c])."""
        pr = self.pairs()
        n = len(pr)
        counts = np.zeros(max(n, 1), np.int32)
        total = self.L.dSpaceGetContactsB200(self.space, _p(counts, _ip), n, None, None, 0)
        pd = np.zeros((max(total, 1), 4), np.float32)
        ns = np.zeros((max(total, 1), 4), np.float32)
        self.L.dSpaceGetContactsB200(self.space, _p(counts, _ip), n, _p(pd, _fp), _p(ns, _fp), total)
        side = ns[:total, 3].copy().view(np.int32)
        return pr, counts[:n], pd[:total], ns[:total, :3].copy(), side

    def solver_order(self):
        n = self.L.dWorldGetSolverOrderB200(self.w, None, None, None, 0)
        g1 = np.zeros(max(n, 1), np.int32); g2 = np.zeros(max(n, 1), np.int32); k = np.zeros(max(n, 1), np.int32)
        self.L.dWorldGetSolverOrderB200(self.w, _p(g1, _ip), _p(g2, _ip), _p(k, _ip), n)
        return g1[:n], g2[:n], k[:n]
