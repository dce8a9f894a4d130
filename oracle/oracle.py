"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module (see oracle/ode_oracle.h: the oracle is the checker, never the product; "parity
unpinned").
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SPHERE, BOX, PLANE, TRIMESH = 0, 1, 4, 8
BODY_KINEMATIC, BODY_NOGRAVITY, BODY_GYRO = 1, 2, 4


class ContactGeom(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("normal", C.c_float * 3), ("depth", C.c_float),
                ("g1", C.c_int), ("g2", C.c_int), ("side1", C.c_int), ("side2", C.c_int)]


class Surface(C.Structure):
    _fields_ = [("mode", C.c_int), ("mu", C.c_float), ("mu2", C.c_float), ("bounce", C.c_float),
                ("bounce_vel", C.c_float), ("soft_erp", C.c_float), ("soft_cfm", C.c_float),
                ("motion1", C.c_float), ("motion2", C.c_float), ("motionN", C.c_float),
                ("slip1", C.c_float), ("slip2", C.c_float), ("fdir1", C.c_float * 3)]


def reference_surface():
    """NearCallback's surface, /root/reference/src/main.c:684-687."""
    s = Surface()
    s.mode = 0x004  # dContactBounce
    s.bounce = 0.2
    s.bounce_vel = 0.1
    s.mu = float("inf")
    return s


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "liboracle.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    fp = C.POINTER(C.c_float)
    ip = C.POINTER(C.c_int)
    L.orc_create.restype = C.c_void_p
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_set_gravity.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
    L.orc_set_params.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_float]
    L.orc_set_contact_params.argtypes = [C.c_void_p, C.c_float, C.c_float]
    L.orc_add_body.argtypes = [C.c_void_p, fp, fp, fp, fp, fp, C.c_float, fp, C.c_int, C.c_int]
    L.orc_add_body.restype = C.c_int
    L.orc_num_bodies.argtypes = [C.c_void_p]
    L.orc_get_body.argtypes = [C.c_void_p, C.c_int, fp, fp, fp, fp, fp]
    L.orc_set_body_state.argtypes = [C.c_void_p, C.c_int, fp, fp, fp, fp, fp]
    L.orc_add_force.argtypes = [C.c_void_p, C.c_int, fp, fp]
    L.orc_add_mesh.argtypes = [C.c_void_p, fp, C.c_int, ip, C.c_int]
    L.orc_add_geom.argtypes = [C.c_void_p, C.c_int, fp, C.c_int, fp, fp, C.c_uint, C.c_uint, C.c_int]
    L.orc_num_geoms.argtypes = [C.c_void_p]
    L.orc_get_aabb.argtypes = [C.c_void_p, C.c_int, fp]
    L.orc_broadphase.argtypes = [C.c_void_p, C.c_int, ip, C.c_long]
    L.orc_broadphase.restype = C.c_long
    L.orc_collide.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(ContactGeom)]
    L.orc_add_contact.argtypes = [C.c_void_p, C.POINTER(ContactGeom), C.POINTER(Surface), C.c_int, C.c_int]
    L.orc_clear_contacts.argtypes = [C.c_void_p]
    L.orc_num_contacts.argtypes = [C.c_void_p]
    L.orc_collide_all.argtypes = [C.c_void_p, C.c_int, C.POINTER(Surface)]
    L.orc_collide_all.restype = C.c_long
    L.orc_quickstep.argtypes = [C.c_void_p, C.c_float, C.c_int, ip]
    L.orc_num_rows.argtypes = [C.c_void_p]
    L.orc_last_lambda.argtypes = [C.c_void_p, fp, C.c_int]
    L.orc_set_keep_rows.argtypes = [C.c_void_p, C.c_int]
    L.orc_last_rows.argtypes = [C.c_void_p, fp, fp, fp, fp, fp, fp, ip, C.c_int]
    L.orc_pack_body_transform.argtypes = [C.c_void_p, C.c_int, fp]
    L.orc_pack_geom_transform.argtypes = [C.c_void_p, C.c_int, fp]
    L.orc_q_to_r.argtypes = [fp, fp]
    L.orc_r_to_q.argtypes = [fp, fp]
    L.orc_plane_space.argtypes = [fp, fp, fp]
    L.orc_rand_next.argtypes = [C.POINTER(C.c_uint)]
    L.orc_rand_next.restype = C.c_uint
    _LIB = L
    return L


def _f(a):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a.ctypes.data_as(C.POINTER(C.c_float)), a


def _fp(a):
    if a is None:
        return None
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


class OracleWorld:
    """Flat-array CPU world mirroring the subset of ODE the reference drives."""

    def __init__(self, gravity=(0.0, -9.8, 0.0), erp=0.2, cfm=1e-5, iters=20, sor_w=1.3):
        self.L = lib()
        self.w = C.c_void_p(self.L.orc_create())
        self.L.orc_set_gravity(self.w, *[float(x) for x in gravity])
        self.L.orc_set_params(self.w, erp, cfm, iters, sor_w)

    def close(self):
        if self.w:
            self.L.orc_destroy(self.w)
            self.w = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- construction
    def add_body(self, pos, q=None, R=None, lvel=None, avel=None, mass=1.0, inertia=None, flags=0, env=0):
        a = [_fp(pos), _fp(q), _fp(R), _fp(lvel), _fp(avel), _fp(inertia)]
        return self.L.orc_add_body(self.w, _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]), _ptr(a[4]),
                                   float(mass), _ptr(a[5]), int(flags), int(env))

    def add_mesh(self, verts, tris):
        v = np.ascontiguousarray(verts, dtype=np.float32)
        t = np.ascontiguousarray(tris, dtype=np.int32)
        return self.L.orc_add_mesh(self.w, _ptr(v), len(v), t.ctypes.data_as(C.POINTER(C.c_int)), len(t))

    def add_geom(self, gtype, dims, body=-1, pos=None, R=None, cat=0xFFFFFFFF, col=0xFFFFFFFF, env=-1):
        d = np.zeros(4, dtype=np.float32)
        dims = np.atleast_1d(np.asarray(dims, dtype=np.float32))
        d[:len(dims)] = dims
        p, r = _fp(pos), _fp(R)
        return self.L.orc_add_geom(self.w, int(gtype), _ptr(d), int(body), _ptr(p), _ptr(r),
                                   int(cat) & 0xFFFFFFFF, int(col) & 0xFFFFFFFF, int(env))

    def load_scene(self, sc):
        """sc: scene dict produced by odeb200.scenes (arrays of bodies / geoms / meshes)."""
        for v, t in sc.get("meshes", []):
            self.add_mesh(v, t)
        b = sc["bodies"]
        for i in range(len(b["pos"])):
            inertia = None if b.get("inertia") is None else b["inertia"][i]
            self.add_body(b["pos"][i], q=b["quat"][i], lvel=b["lvel"][i], avel=b["avel"][i],
                          mass=float(b["mass"][i]), inertia=inertia, flags=int(b["flags"][i]), env=int(b["env"][i]))
        g = sc["geoms"]
        for i in range(len(g["type"])):
            self.add_geom(int(g["type"][i]), g["dims"][i], body=int(g["body"][i]), pos=g["pos"][i], R=g["R"][i],
                          cat=int(g["cat"][i]), col=int(g["col"][i]), env=int(g["env"][i]))

    # -- queries
    @property
    def num_bodies(self):
        return self.L.orc_num_bodies(self.w)

    @property
    def num_geoms(self):
        return self.L.orc_num_geoms(self.w)

    def body(self, i):
        pos = np.zeros(3, np.float32); q = np.zeros(4, np.float32); R = np.zeros(12, np.float32)
        lv = np.zeros(3, np.float32); av = np.zeros(3, np.float32)
        self.L.orc_get_body(self.w, i, _ptr(pos), _ptr(q), _ptr(R), _ptr(lv), _ptr(av))
        return pos, q, R, lv, av

    def state(self):
        n = self.num_bodies
        pos = np.zeros((n, 3), np.float32); q = np.zeros((n, 4), np.float32); R = np.zeros((n, 12), np.float32)
        lv = np.zeros((n, 3), np.float32); av = np.zeros((n, 3), np.float32)
        for i in range(n):
            self.L.orc_get_body(self.w, i, _ptr(pos[i]), _ptr(q[i]), _ptr(R[i]), _ptr(lv[i]), _ptr(av[i]))
        return {"pos": pos, "quat": q, "R": R, "lvel": lv, "avel": av}

    def set_body_state(self, i, pos=None, q=None, R=None, lvel=None, avel=None):
        a = [_fp(pos), _fp(q), _fp(R), _fp(lvel), _fp(avel)]
        self.L.orc_set_body_state(self.w, i, *[_ptr(x) for x in a])

    def aabb(self, g):
        a = np.zeros(6, np.float32)
        self.L.orc_get_aabb(self.w, g, _ptr(a))
        return a

    def broadphase(self, method=0):
        n = self.L.orc_broadphase(self.w, method, None, 0)
        out = np.zeros((max(n, 1), 2), np.int32)
        n2 = self.L.orc_broadphase(self.w, method, out.ctypes.data_as(C.POINTER(C.c_int)), n)
        assert n2 == n
        return out[:n]

    def collide(self, g1, g2, maxc=8):
        buf = (ContactGeom * 8)()
        n = self.L.orc_collide(self.w, g1, g2, maxc, buf)
        return [buf[i] for i in range(n)]

    def add_contact(self, cg, surf, b1, b2):
        return self.L.orc_add_contact(self.w, C.byref(cg), C.byref(surf), b1, b2)

    def clear_contacts(self):
        self.L.orc_clear_contacts(self.w)

    @property
    def num_contacts(self):
        return self.L.orc_num_contacts(self.w)

    def collide_all(self, maxc=8, surf=None):
        surf = surf or reference_surface()
        return self.L.orc_collide_all(self.w, maxc, C.byref(surf))

    def quickstep(self, h, order_mode=0, perm=None):
        p = None
        if perm is not None:
            perm = np.ascontiguousarray(perm, dtype=np.int32)
            p = perm.ctypes.data_as(C.POINTER(C.c_int))
        return self.L.orc_quickstep(self.w, float(h), order_mode, p)

    def tick(self, h, maxc=8, surf=None, order_mode=0):
        """One reference tick: dSpaceCollide + NearCallback, step, dJointGroupEmpty (src/main.c:212-214)."""
        nc = self.collide_all(maxc, surf)
        self.quickstep(h, order_mode)
        self.clear_contacts()
        return nc

    @property
    def num_rows(self):
        return self.L.orc_num_rows(self.w)

    def last_lambda(self):
        n = self.num_rows
        out = np.zeros(max(n, 1), np.float32)
        self.L.orc_last_lambda(self.w, _ptr(out), n)
        return out[:n]

    def keep_rows(self, on=True):
        self.L.orc_set_keep_rows(self.w, 1 if on else 0)

    def last_rows(self):
        """Rows of the last step (before SOR_LCP's Ad scaling): dict of J[m,12], c, cfm (= cfm/h), lo, hi, rhs,
        jb[m,2] (body pair, -1 = none) and lam (the lambda the step ended with)."""
        m = self.L.orc_last_rows(self.w, None, None, None, None, None, None, None, 0)
        J = np.zeros((max(m, 1), 12), np.float32)
        v = [np.zeros(max(m, 1), np.float32) for _ in range(5)]
        jb = np.zeros((max(m, 1), 2), np.int32)
        self.L.orc_last_rows(self.w, _ptr(J), *[_ptr(x) for x in v], jb.ctypes.data_as(C.POINTER(C.c_int)), m)
        return {"J": J[:m], "c": v[0][:m], "cfm": v[1][:m], "lo": v[2][:m], "hi": v[3][:m], "rhs": v[4][:m], "jb": jb[:m],
                "lam": self.last_lambda()[:m]}

    def body_transform(self, b):
        out = np.zeros(16, np.float32)
        self.L.orc_pack_body_transform(self.w, b, _ptr(out))
        return out

    def geom_transform(self, g):
        out = np.zeros(16, np.float32)
        self.L.orc_pack_geom_transform(self.w, g, _ptr(out))
        return out
