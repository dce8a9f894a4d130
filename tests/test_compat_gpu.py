"""GPU: the exact drop-in ("compat") path -- dSpaceCollide invoking a host NearCallback that calls
dCollide / dJointCreateContact / dJointAttach, as /root/reference/src/main.c:674-693 does -- gives the
same physics as the device-resident path, bit for bit; and the headless C host harness
(rl-ode-physics_b200/host/physics_server.c, the reference's StartServer loop against our ode/ode.h)
produces the MsgUpdateBodies image the Python-built scene predicts."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

import odeb200
import util
from odeb200 import scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class CompatWorld:
    """The reference's AddBody/AddBodyMap/NearCallback/tick written against the ODE handle API."""

    def __init__(self, sc):
        L = self.L = odeb200.lib()
        L.dInitODE()
        self.world = C.c_void_p(L.dWorldCreate())
        L.dWorldSetGravity(self.world, *[float(g) for g in sc["gravity"]])
        self.space = C.c_void_p(L.dHashSpaceCreate(None))
        self.group = C.c_void_p(L.dJointGroupCreate(0))
        self.bodies, self.geoms = [], []
        b, g = sc["bodies"], sc["geoms"]
        fp = C.POINTER(C.c_float)
        for i in range(len(b["pos"])):
            h = C.c_void_p(L.dBodyCreate(self.world))
            L.dBodySetPosition(h, *[float(x) for x in b["pos"][i]])
            q = np.ascontiguousarray(b["quat"][i], np.float32)
            L.dBodySetQuaternion(h, q.ctypes.data_as(fp))
            L.dBodySetLinearVel(h, *[float(x) for x in b["lvel"][i]])
            L.dBodySetAngularVel(h, *[float(x) for x in b["avel"][i]])
            if int(b["flags"][i]) & scenes.BODY_KINEMATIC:
                L.dBodySetKinematic(h)
            L.dBodySetGyroscopicMode(h, 1 if int(b["flags"][i]) & scenes.BODY_GYRO else 0)
            self.bodies.append(h)
        for i in range(len(g["type"])):
            t, d = int(g["type"][i]), [float(x) for x in g["dims"][i]]
            if t == scenes.SPHERE:
                h = L.dCreateSphere(self.space, d[0])
            elif t == scenes.BOX:
                h = L.dCreateBox(self.space, d[0], d[1], d[2])
            else:
                h = L.dCreatePlane(self.space, d[0], d[1], d[2], d[3])
            h = C.c_void_p(h)
            if int(g["body"][i]) >= 0:
                L.dGeomSetBody(h, self.bodies[int(g["body"][i])])
            elif t != scenes.PLANE:
                L.dGeomSetPosition(h, *[float(x) for x in g["pos"][i]])
                R = np.ascontiguousarray(g["R"][i], np.float32)
                L.dGeomSetRotation(h, R.ctypes.data_as(fp))
            L.dGeomSetCategoryBits(h, int(g["cat"][i]))
            L.dGeomSetCollideBits(h, int(g["col"][i]))
            self.geoms.append(h)
        self.n_callbacks = 0
        self.n_joints = 0
        # the comparisons in this file are about the callback path (same contacts, same rows as the device-resident
        # path, which steps with dWorldQuickStep): make dWorldStep run QuickStep's sweeps here; its default -- the exact
        # solve of the step's LCP -- is tested against the oracle in test_parity_gpu.py
        L.dWorldSetStepSolverB200(self.world, -1, 0.0)

        def near(data, o1, o2):
            # src/main.c:674-693
            self.n_callbacks += 1
            contacts = (odeb200.Contact * 8)()
            geom0 = C.cast(C.addressof(contacts) + odeb200.Contact.geom.offset, C.POINTER(odeb200.ContactGeom))
            nc = L.dCollide(o1, o2, 8, geom0, C.sizeof(odeb200.Contact))
            for k in range(nc):
                contacts[k].surface.mode = odeb200.dContactBounce
                contacts[k].surface.bounce = 0.2
                contacts[k].surface.bounce_vel = 0.1
                contacts[k].surface.mu = float("inf")
                j = L.dJointCreateContact(self.world, self.group, C.byref(contacts[k]))
                L.dJointAttach(j, L.dGeomGetBody(o1), L.dGeomGetBody(o2))
                self.n_joints += 1

        self._cb = odeb200.NearCallback(near)

    def tick(self, h):
        L = self.L
        L.dSpaceCollide(self.space, None, self._cb)       # src/main.c:212
        L.dWorldStep(self.world, float(h))                # :213
        L.dJointGroupEmpty(self.group)                    # :214

    def state(self):
        L = self.L
        pos = np.array([[L.dBodyGetPosition(b)[k] for k in range(3)] for b in self.bodies], np.float32)
        R = np.array([[L.dBodyGetRotation(b)[k] for k in range(12)] for b in self.bodies], np.float32)
        q = np.array([[L.dBodyGetQuaternion(b)[k] for k in range(4)] for b in self.bodies], np.float32)
        lv = np.array([[L.dBodyGetLinearVel(b)[k] for k in range(3)] for b in self.bodies], np.float32)
        av = np.array([[L.dBodyGetAngularVel(b)[k] for k in range(3)] for b in self.bodies], np.float32)
        return {"pos": pos, "R": R, "quat": q, "lvel": lv, "avel": av}

    def close(self):
        L = self.L
        for g in self.geoms:
            L.dGeomDestroy(g)
        for b in self.bodies:
            L.dBodyDestroy(b)
        L.dJointGroupDestroy(self.group)
        L.dSpaceDestroy(self.space)
        L.dWorldDestroy(self.world)
        L.dCloseODE()


@pytest.mark.parametrize("name", ["c1_low", "c1p_low"])
def test_callback_path_equals_device_resident_path(name):
    sc = scenes.server_scene(seed=1, y_range=(1.0, 6.0), floor_plane=(name == "c1p_low"))
    cw = CompatWorld(sc)
    ew = util.engine_world(sc)
    for step in range(12):
        cw.tick(sc["h"])
        ew.tick(sc["h"])
        a, b = cw.state(), ew.state()
        for k in ("pos", "quat", "lvel", "avel", "R"):
            assert np.array_equal(a[k], b[k]), (step, k)
    assert cw.n_callbacks >= 12 * 40 and cw.n_joints > 12 * 30
    # and against the oracle
    ow = util.oracle_world(sc)
    ow._types = [int(t) for t in sc["geoms"]["type"]]
    ow._bodies = [int(x) for x in sc["geoms"]["body"]]
    ew2 = util.engine_world(sc)
    ew2.tick(sc["h"])
    util.oracle_tick_in_engine_order(ow, ew2, sc["h"])
    cw2 = CompatWorld(sc)
    cw2.tick(sc["h"])
    s, o = cw2.state(), ow.state()
    for k in ("pos", "quat", "lvel", "avel"):
        assert util.rel_err(s[k], o[k]).max() <= 1e-4
    for w in (cw, cw2):
        w.close()
    ew.close(); ew2.close()


def _run_server(mode, ticks, out, stepper="quick"):
    exe = os.path.join(ROOT, "rl-ode-physics_b200", "host", "physics_server")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(exe)])
    dt = repr(float(np.float32(1.0 / 60.0)))
    subprocess.check_call([exe, "1", "64", "4", str(ticks), dt, mode, out, stepper])
    raw = np.fromfile(out, dtype=np.uint8)
    assert raw.size == 43012                      # sizeof(MsgUpdateBodies) with float dReal
    msg = int(np.frombuffer(raw[:4].tobytes(), np.int32)[0])
    states = raw[4:].reshape(512, 84)
    types = np.frombuffer(states[:, 0:4].tobytes(), np.int32)
    transforms = np.frombuffer(states[:, 4:68].tobytes(), np.float32).reshape(512, 16)
    sizes = np.frombuffer(states[:, 68:80].tobytes(), np.float32).reshape(512, 3)
    return msg, types, transforms, sizes


def test_headless_reference_server_loop_drop_in():
    ticks = 90
    with tempfile.TemporaryDirectory() as d:
        msg_c, types_c, tr_c, sz_c = _run_server("compat", ticks, os.path.join(d, "c.bin"))
        msg_d, types_d, tr_d, sz_d = _run_server("device", ticks, os.path.join(d, "d.bin"))
        raw_c = np.fromfile(os.path.join(d, "c.bin"), dtype=np.uint8)
        _run_server("wire", ticks, os.path.join(d, "w.bin"))
        raw_w = np.fromfile(os.path.join(d, "w.bin"), dtype=np.uint8)
        # the reference's own stepper: dWorldStep with its default, the exact solve (both contact paths again)
        _, _, tr_xc, _ = _run_server("compat", ticks, os.path.join(d, "xc.bin"), "step")
        _, _, tr_xd, _ = _run_server("device", ticks, os.path.join(d, "xd.bin"), "step")
    assert np.array_equal(tr_xc, tr_xd)
    assert not np.array_equal(tr_xc, tr_c) and np.abs(tr_xc[4:72, 12:15] - tr_c[4:72, 12:15]).max() < 2.0
    assert msg_c == 3 and msg_d == 3              # MSGTYPE_C_UPDATE_BODIES
    assert np.array_equal(types_c, types_d) and np.array_equal(tr_c, tr_d)
    # the GPU-packed wire image (dWorldPackMsgUpdateBodiesB200) equals the host-assembled one byte for byte,
    # except the unused padding-free NULL slots, which both leave zeroed
    assert np.array_equal(raw_c, raw_w)
    # slots: 4 map boxes, 64 spawned bodies, 4 kinematic spheres, the rest BODYTYPE_NULL
    assert (types_c[:4] == 2).all() and (types_c[72:] == 0).all() and set(types_c[4:68].tolist()) <= {1, 2}
    sc = scenes.server_scene(seed=1)
    assert np.array_equal(sz_c[4:68][types_c[4:68] == 2], sc["geoms"]["dims"][4:68][sc["geoms"]["type"][4:68] == scenes.BOX][:, :3])
    ew = util.engine_world(sc)
    for _ in range(ticks):
        ew.tick(sc["h"])
    snap = ew.snapshot()
    assert np.array_equal(tr_c[4:72], snap)       # same scene through the bulk API: identical transforms
    # static map slots: the broadcast loop re-packs dGeomGetRotation with GetTransformMat, i.e. the
    # TRANSPOSE of the GetTransformMatV rows the geom was created with
    rm = scenes.transform_mat_v((4, 3, 0), (0, 0, -0.5))[:12].reshape(3, 4)
    assert np.array_equal(tr_c[1].reshape(4, 4)[:3, :3], rm[:, :3].T.T.T)
    assert np.array_equal(tr_c[1][12:], np.array([4, 3, 0, 1], np.float32))
    ew.close()


@pytest.mark.parametrize("seed,n", [(701, 5), (702, 60), (703, 400), (704, 1500), (705, 2500)])
def test_fuzz_callback_path_equals_device_path_while_the_world_grows(seed, n):
    """Random soups of 5 .. 2500 bodies through the callback path and through the device-resident path: identical bits for
    six ticks.  The larger ones outgrow the mapped read-back buffer of the callback path (sized for 1024 pairs at first),
    so the copying fallback of the outgrown tick and the enlarged buffer of the following ticks are exercised too."""
    sc = scenes.random_soup(n, seed=seed, extent=1.5 + 0.25 * n ** (1.0 / 3.0), with_static_box=(seed % 2 == 0))
    cw = CompatWorld(sc)
    ew = util.engine_world(sc)
    for step in range(6):
        cw.tick(sc["h"])
        ew.tick(sc["h"])
        a, b = cw.state(), ew.state()
        for k in ("pos", "quat", "lvel", "avel", "R"):
            assert np.array_equal(a[k], b[k]), (step, k)
    assert n < 60 or cw.n_joints > n // 4
    cw.close()
    ew.close()
