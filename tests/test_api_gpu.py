"""GPU: behaviour of the ODE handle API beyond the main tick: spawning bodies while the simulation runs
(the reference's MSGTYPE_S_NEW_BODY -> AddBody, src/main.c:178-182), dMass*/dBodySetMass, world
parameters, destroying objects (src/main.c:258-267), empty and static-only worlds."""
import ctypes as C

import numpy as np
import pytest

import odeb200
import oracle as O
import util
from odeb200 import scenes

pytestmark = pytest.mark.gpu
fp = C.POINTER(C.c_float)
H = 1.0 / 60.0


def _near_callback(L, world, group, counter):
    def near(data, o1, o2):
        contacts = (odeb200.Contact * 8)()
        geom0 = C.cast(C.addressof(contacts) + odeb200.Contact.geom.offset, C.POINTER(odeb200.ContactGeom))
        nc = L.dCollide(o1, o2, 8, geom0, C.sizeof(odeb200.Contact))
        counter[0] += 1
        for k in range(nc):
            contacts[k].surface.mode = odeb200.dContactBounce
            contacts[k].surface.bounce = 0.2
            contacts[k].surface.bounce_vel = 0.1
            contacts[k].surface.mu = float("inf")
            j = L.dJointCreateContact(world, group, C.byref(contacts[k]))
            L.dJointAttach(j, L.dGeomGetBody(o1), L.dGeomGetBody(o2))
    return odeb200.NearCallback(near)


class Server:
    """the reference's world setup + tick through the handle API"""

    def __init__(self):
        L = self.L = odeb200.lib()
        L.dInitODE()
        self.world = C.c_void_p(L.dWorldCreate())
        L.dWorldSetGravity(self.world, 0.0, -9.8, 0.0)
        self.space = C.c_void_p(L.dHashSpaceCreate(None))
        self.group = C.c_void_p(L.dJointGroupCreate(0))
        # these tests compare with the oracle's 20-sweep QuickStep: make dWorldStep run the sweeps (its default, the exact
        # solve, has its own tests in test_parity_gpu.py)
        L.dWorldSetStepSolverB200(self.world, -1, 0.0)
        self.calls = [0]
        self.cb = _near_callback(L, self.world, self.group, self.calls)
        self.bodies, self.geoms = [], []

    def add_static_box(self, pos, size):
        g = C.c_void_p(self.L.dCreateBox(self.space, *[float(x) for x in size]))
        self.L.dGeomSetPosition(g, *[float(x) for x in pos])
        self.geoms.append(g)
        return g

    def add_body(self, pos, kind, dims, mass=None):
        L = self.L
        b = C.c_void_p(L.dBodyCreate(self.world))
        L.dBodySetPosition(b, *[float(x) for x in pos])
        if kind == "sphere":
            g = C.c_void_p(L.dCreateSphere(self.space, float(dims[0])))
        else:
            g = C.c_void_p(L.dCreateBox(self.space, *[float(x) for x in dims]))
        if mass is not None:
            m = odeb200.Mass()
            if kind == "sphere":
                L.dMassSetSphere(C.byref(m), float(mass), float(dims[0]))
            else:
                L.dMassSetBox(C.byref(m), float(mass), *[float(x) for x in dims])
            L.dBodySetMass(b, C.byref(m))
        L.dGeomSetBody(g, b)
        self.bodies.append(b); self.geoms.append(g)
        return b, g

    def tick(self, h=H):
        self.L.dSpaceCollide(self.space, None, self.cb)
        self.L.dWorldStep(self.world, float(h))
        self.L.dJointGroupEmpty(self.group)

    def pos(self, b):
        p = self.L.dBodyGetPosition(b)
        return np.array([p[0], p[1], p[2]], np.float32)

    def vel(self, b):
        p = self.L.dBodyGetLinearVel(b)
        return np.array([p[0], p[1], p[2]], np.float32)

    def close(self):
        L = self.L
        for g in self.geoms:
            L.dGeomDestroy(g)
        for b in self.bodies:
            L.dBodyDestroy(b)
        L.dJointGroupDestroy(self.group)
        L.dSpaceDestroy(self.space)
        L.dWorldDestroy(self.world)


def test_bodies_spawned_while_running_match_the_oracle():
    s = Server()
    ow = O.OracleWorld()
    s.add_static_box((0, 0, 0), (100, 1, 100))
    ow.add_geom(O.BOX, [100, 1, 100], pos=[0, 0, 0])
    rs = np.random.RandomState(5)
    for step in range(90):
        if step % 10 == 0:                       # a client pressed `M` (src/main.c:502-522)
            pos = [rs.uniform(-1, 1), rs.uniform(1.5, 3.0), rs.uniform(-1, 1)]
            if rs.randint(2):
                dims = list(rs.uniform(0.2, 1.0, 3)); kind, t = "box", O.BOX
            else:
                dims = [rs.uniform(0.1, 0.4)]; kind, t = "sphere", O.SPHERE
            s.add_body(pos, kind, dims)
            b = ow.add_body(np.float32(pos), flags=O.BODY_GYRO)
            ow.add_geom(t, np.float32(dims), body=b)
        s.tick()
        ow.tick(H, order_mode=0)
    assert s.calls[0] > 90
    # both integrate the same scene; row orders differ (ODE's shuffle vs colours), so compare loosely
    for i, b in enumerate(s.bodies):
        po = ow.body(i)[0]
        assert np.abs(s.pos(b) - po).max() < 0.08, i
        assert s.pos(b)[1] > 0.5
    s.close()


def test_dmass_and_world_parameters_take_effect():
    s = Server()
    L = s.L
    L.dWorldSetERP(s.world, 0.5); L.dWorldSetCFM(s.world, 1e-4)
    L.dWorldSetQuickStepNumIterations(s.world, 7); L.dWorldSetQuickStepW(s.world, 1.1)
    assert L.dWorldGetQuickStepNumIterations(s.world) == 7
    ow = O.OracleWorld(erp=0.5, cfm=1e-4, iters=7, sor_w=1.1)
    s.add_static_box((0, 0, 0), (20, 1, 20)); ow.add_geom(O.BOX, [20, 1, 20], pos=[0, 0, 0])
    b1, g1 = s.add_body((0, 0.99, 0), "box", (1.0, 1.0, 0.5), mass=3.0)            # density 3
    b2, g2 = s.add_body((0.1, 1.88, 0.0), "sphere", (0.4,), mass=2.0)
    m1 = 3.0 * 1.0 * 1.0 * 0.5
    I1 = np.diag([m1 / 12 * (1 + 0.25), m1 / 12 * (1 + 0.25), m1 / 12 * 2.0]).astype(np.float32)
    m2 = np.float32(4.0 / 3.0) * np.float32(np.pi) * np.float32(0.4) ** 3 * np.float32(2.0)
    I2 = (np.eye(3) * 0.4 * m2 * 0.16).astype(np.float32)
    o1 = ow.add_body([0, 0.99, 0], mass=m1, inertia=I1.reshape(9), flags=O.BODY_GYRO); ow.add_geom(O.BOX, [1, 1, 0.5], body=o1)
    o2 = ow.add_body(np.float32([0.1, 1.88, 0.0]), mass=float(m2), inertia=I2.reshape(9), flags=O.BODY_GYRO); ow.add_geom(O.SPHERE, [0.4], body=o2)
    # one tick solved in the same (engine) order: build the engine's order from a device-resident twin
    sc_types = [O.BOX, O.BOX, O.SPHERE]
    s.tick()
    ow._types, ow._bodies = sc_types, [-1, 0, 1]
    # oracle in joint order == engine order here (2 pairs, disjoint colour per body chain) up to tolerance
    ow.tick(H, order_mode=1)
    for (b, o) in ((b1, o1), (b2, o2)):
        assert np.abs(s.pos(b) - ow.body(o)[0]).max() < 2e-4
        assert np.abs(s.vel(b) - ow.body(o)[3]).max() < 2e-2
    # heavier sphere sinks the same way in both; parameters really were used: ERP 0.5 pushes out faster than 0.2
    s.close()


def test_destroy_and_empty_worlds():
    s = Server()
    s.tick()                                            # empty space: nothing to do, must not fail
    floor = s.add_static_box((0, 0, 0), (10, 1, 10))
    s.tick()                                            # static only
    b, g = s.add_body((0, 0.9, 0), "sphere", (0.5,))
    b2, g2 = s.add_body((2, 0.9, 0), "sphere", (0.5,))
    for _ in range(5):
        s.tick()
    assert s.pos(b)[1] > 0.9 and s.pos(b2)[1] > 0.9     # resting on / pushed out of the floor
    # destroy the floor geom: the spheres now fall freely
    s.L.dGeomDestroy(floor)
    y0 = s.pos(b)[1]
    for _ in range(30):
        s.tick()
    assert s.pos(b)[1] < y0 - 0.8
    # destroy one body + its geom: the other keeps simulating, the handle API stays usable
    s.L.dGeomDestroy(g2); s.L.dBodyDestroy(b2)
    s.bodies.remove(b2); s.geoms.remove(g2); s.geoms.remove(floor)
    y1 = s.pos(b)[1]
    s.tick()
    assert s.pos(b)[1] < y1
    s.close()


def test_kinematic_body_via_api_and_velocity_setters():
    s = Server()
    L = s.L
    s.add_static_box((0, 0, 0), (10, 1, 10))
    k, gk = s.add_body((0, 1.0, 0), "sphere", (0.5,))
    L.dBodySetKinematic(k)
    L.dBodySetLinearVel(k, 1.0, 0.0, 0.0)
    d, gd = s.add_body((0.9, 1.0, 0), "sphere", (0.5,))
    for _ in range(10):
        s.tick()
    assert s.pos(k)[1] == 1.0 and abs(s.pos(k)[0] - 10 * H) < 1e-5
    assert s.vel(d)[0] > 0.5
    s.close()


def _identity_R12():
    return np.float32([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0])


@pytest.mark.parametrize("scene", ["soup", "batch"])
def test_incremental_ingestion_equals_full_upload(scene):
    """Spawns and setters on a running world are sent as queued field patches (one packed copy + one scatter
    kernel, no mirror refresh).  The result must be bit-identical to refreshing the mirrors and re-uploading
    every array after each edit (the path every world takes at its first sync)."""
    outs = []
    for full in (False, True):
        if scene == "soup":
            sc = scenes.random_soup(200, seed=7)
        else:
            sc = scenes.batched_worlds_scene(4, seed=4, spacing=0.7)
        ew = util.engine_world(sc)
        L = ew.L
        rs = np.random.RandomState(3)
        player = None
        for step in range(30):
            if step == 5 and scene == "soup":       # a kinematic "player" sphere appears (src/main.c:150-160)
                player, _ = ew.spawn([0.0, 1.0, 0.0], "sphere", [0.5])
                L.dBodySetKinematic(player)
            if step >= 5 and step % 6 == 5 and scene == "soup":   # a client pressed M (src/main.c:502-522)
                pos = [rs.uniform(-2, 2), rs.uniform(2, 4), rs.uniform(-2, 2)]
                if rs.randint(2):
                    ew.spawn(pos, "box", list(rs.uniform(0.2, 1.0, 3)), R=_identity_R12())
                else:
                    ew.spawn(pos, "sphere", [rs.uniform(0.1, 0.4)])
            if player is not None:                   # MSGTYPE_S_PLAYER_UPDATE: the player body is teleported
                L.dBodySetPosition(player, float(0.02 * step), 1.0, float(-0.01 * step))
            if step % 4 == 1:                        # velocity / force setters on bodies that are mid-flight
                b = ew.body_handle(3)
                L.dBodyAddForce(b, 0.0, 30.0, 0.0)
                L.dBodyAddForce(b, 5.0, 0.0, 0.0)
                L.dBodySetAngularVel(ew.body_handle(7), 0.0, 2.0, 0.0)
                L.dBodySetLinearVel(ew.body_handle(11), 1.0, 0.5, 0.0)
            if full:
                ew.force_full_sync()
            ew.tick(sc["h"])
        outs.append(ew.state())
        assert ew.stats()["n_contacts"] > 10
        ew.close()
    a, b = outs
    assert len(a["pos"]) == len(b["pos"]) and (scene != "soup" or len(a["pos"]) > 200)
    for k in ("pos", "quat", "lvel", "avel"):
        assert np.array_equal(a[k], b[k]), k


def test_destroyed_body_leaves_its_geom_in_place_incrementally():
    """dBodyDestroy on a running world: the geom stays where the body was (ODE detaches it), other bodies keep
    colliding with it; done through queued patches, no full refresh."""
    sc = scenes.random_soup(50, seed=2)
    ew = util.engine_world(sc)
    for _ in range(5):
        ew.tick(sc["h"])
    st = ew.state()
    b = ew.body_handle(4)
    ew.L.dBodyDestroy(b)
    for _ in range(3):
        ew.tick(sc["h"])
    st2 = ew.state()
    # the destroyed body's slot is inert now; everything else kept moving
    assert np.array_equal(st2["pos"][4], st["pos"][4])
    assert not np.array_equal(st2["pos"][5], st["pos"][5])
    ew.close()


def test_geom_detached_from_its_body_stays_where_the_body_was():
    """dGeomSetBody(g, 0): ODE leaves the geom at the pose it had on the body; it becomes a static obstacle there."""
    s = Server()
    s.add_static_box((0, -0.5, 0), (40, 1, 40))
    b, g = s.add_body((2.0, 0.3, 1.0), "box", (1.0, 0.6, 1.0))
    for _ in range(60):
        s.tick()
    rest = s.pos(b)
    assert abs(rest[1] - 0.3) < 0.02
    s.L.dGeomSetBody(g, None)
    gp = s.L.dGeomGetPosition(g)
    assert np.array_equal(np.float32([gp[0], gp[1], gp[2]]), rest)          # not the origin, not the spawn pose
    ball, _ = s.add_body((2.0, 3.0, 1.0), "sphere", (0.25,))
    for _ in range(120):
        s.tick()
    assert abs(s.pos(ball)[1] - (rest[1] + 0.3 + 0.25)) < 0.03                # it landed on the detached box
    s.close()


def test_obj_loader_builds_the_same_trimesh_as_buildsingle(tmp_path):
    """SURVEY section 8 f4 (loader half): an OBJ with quads, `a/b/c` references, negative indices and noise lines
    must give the same collision mesh -- hence bit-identical contacts -- as dGeomTriMeshDataBuildSingle fed
    with the triangulation done here."""
    verts = np.float32([[-2, 0, -2], [2, 0, -2], [2, 0, 2], [-2, 0, 2], [0, 1.0, 0]])
    lines = ["# test mesh", "o pyramid", "mtllib none.mtl"]
    lines += ["v %g %g %g" % tuple(v) for v in verts]
    lines += ["vn 0 1 0", "vt 0 0", "usemtl m", "s off",
              "f 1/1/1 2/1/1 3/1/1 4/1/1",        # quad -> fan: (0,1,2), (0,2,3)
              "f 1//1 5//1 2//1",
              "f -4 -1 -3",                         # relative: (1, 4, 2)
              "f 3 5 4", "f 4 5 1"]
    path = tmp_path / "pyramid.obj"
    path.write_text("\n".join(lines) + "\n")
    tris = np.int32([[0, 1, 2], [0, 2, 3], [0, 4, 1], [1, 4, 2], [2, 4, 3], [3, 4, 0]])
    L = odeb200.lib()
    results = []
    for use_obj in (True, False):
        s = Server()
        d = C.c_void_p(L.dGeomTriMeshDataCreate())
        if use_obj:
            assert L.dGeomTriMeshDataBuildFromOBJB200(d, str(path).encode()) == len(tris)
        else:
            v = np.ascontiguousarray(verts); t = np.ascontiguousarray(tris)
            L.dGeomTriMeshDataBuildSingle(d, v.ctypes.data, 12, len(v), t.ctypes.data, 3 * len(t), 12)
        g = C.c_void_p(L.dCreateTriMesh(s.space, d, None, None, None))
        s.geoms.append(g)
        balls = [s.add_body(p, "sphere", (0.3,))[0] for p in ([0.5, 1.2, 0.3], [-1.0, 0.9, 0.2], [1.2, 0.5, -0.8], [0.0, 1.6, 0.0])]
        for _ in range(40):
            s.tick()
        results.append(np.stack([s.pos(b) for b in balls]))
        s.close()
        L.dGeomTriMeshDataDestroy(d)
    assert np.array_equal(results[0], results[1])
    assert results[0][:, 1].min() > 0.2              # nobody fell through the mesh
    assert L.dGeomTriMeshDataBuildFromOBJB200(C.c_void_p(L.dGeomTriMeshDataCreate()), b"/nonexistent.obj") == -1


def test_dcollide_outside_the_space_traversal_matches_the_oracle():
    """dCollide(o1, o2) called directly (not from dSpaceCollide's callback) runs the one pair through the same
    narrowphase kernels: contacts bit-equal to the oracle's dCollide, in either argument order."""
    s = Server()
    L = s.L
    ow = O.OracleWorld()
    floor = s.add_static_box((0, 0, 0), (10, 1, 10)); ow.add_geom(O.BOX, [10, 1, 10], pos=[0, 0, 0])
    b1, g1 = s.add_body((0.1, 0.9, 0.0), "box", (1.0, 1.0, 1.0))
    o1 = ow.add_body(np.float32([0.1, 0.9, 0.0]), flags=O.BODY_GYRO); ow.add_geom(O.BOX, [1, 1, 1], body=o1)
    b2, g2 = s.add_body((0.6, 1.5, 0.2), "sphere", (0.45,))
    o2 = ow.add_body(np.float32([0.6, 1.5, 0.2]), flags=O.BODY_GYRO); ow.add_geom(O.SPHERE, [0.45], body=o2)
    contacts = (odeb200.Contact * 8)()
    geom0 = C.cast(C.addressof(contacts) + odeb200.Contact.geom.offset, C.POINTER(odeb200.ContactGeom))
    for (ga, gb, ia, ib) in ((floor, g1, 0, 1), (g1, g2, 1, 2), (g2, g1, 2, 1), (floor, g2, 0, 2)):
        n = L.dCollide(ga, gb, 8, geom0, C.sizeof(odeb200.Contact))
        ref = ow.collide(ia, ib, 8)
        assert n == len(ref), (ia, ib, n, len(ref))
        for k in range(n):
            cg = contacts[k].geom
            assert np.array_equal(np.float32(list(cg.pos)[:3]), np.float32(list(ref[k].pos)[:3])), (ia, ib, k)
            assert np.array_equal(np.float32(list(cg.normal)[:3]), np.float32(list(ref[k].normal)[:3])), (ia, ib, k)
            assert np.float32(cg.depth) == np.float32(ref[k].depth)
    # the world still ticks normally afterwards
    for _ in range(3):
        s.tick()
    assert s.pos(b1)[1] > 0.9
    s.close()


@pytest.mark.parametrize("scene", ["soup", "batch"])
def test_compact_snapshot_formats_expand_to_the_reference_layout_bit_for_bit(scene):
    """dWorldSetSnapshotFormatB200: format 1 (the 12 non-constant floats) and format 2 (position + quaternion),
    expanded on the host by dSnapshotExpandB200, must equal the 16-float GetTransformMat records (src/main.c:602-622)
    byte for byte -- on the global solver's tail (soup) and on the island solver's fused tail (batch)."""
    sc = scenes.random_soup(300, seed=5) if scene == "soup" else scenes.batched_worlds_scene(40, seed=4, spacing=0.7)
    snaps = {}
    for fmt in (0, 1, 2):
        ew = util.engine_world(sc)
        ew.set_snapshot_format(fmt)
        for _ in range(12):
            ew.tick(sc["h"])
        raw = ew.snapshot()
        assert raw.shape[1] == {0: 16, 1: 12, 2: 8}[fmt]
        snaps[fmt] = ew.expand_snapshot(raw, fmt, threads=3)
        if fmt == 2:                      # the window copy honours the record size
            assert np.array_equal(ew.snapshot(first=7, count=9), raw[7:16])
            st = ew.state()
            assert np.array_equal(raw[:, :3], st["pos"]) and np.array_equal(raw[:, 4:], st["quat"])
        ew.close()
    assert np.array_equal(snaps[0].view(np.uint32), snaps[1].view(np.uint32))
    assert np.array_equal(snaps[0].view(np.uint32), snaps[2].view(np.uint32))
    assert (snaps[0][:, 15] == 1).all() and np.abs(snaps[0][:, :3]).max() <= 1.0 + 1e-6
    # switching the format between ticks re-derives the records from the state, without a step
    ew = util.engine_world(sc)
    for _ in range(3):
        ew.tick(sc["h"])
    full = ew.snapshot()
    ew.set_snapshot_format(2)
    assert np.array_equal(ew.expand_snapshot(ew.snapshot(), 2), full)
    ew.close()


def test_snapshot_of_a_body_spawned_since_the_last_step():
    """The reference's broadcast loop (src/main.c:221-242) may run right after MSGTYPE_S_NEW_BODY -> AddBody
    (src/main.c:178-182) with no physics step in between: the snapshot and the MsgUpdateBodies wire image must
    carry the spawn pose, not whatever the snapshot buffer held."""
    sc = scenes.server_scene(seed=1, y_range=(1.0, 6.0))
    ew = util.engine_world(sc)
    for _ in range(5):
        ew.tick(sc["h"])
    n0 = ew.L.dWorldGetNumBodiesB200(ew.w)
    R = np.float32([0, 0, 1, 0, 0, 1, 0, 0, -1, 0, 0, 0])          # a quarter turn about y
    ew.spawn((1.5, 7.0, -2.0), "box", (0.4, 0.5, 0.6), R=R)
    snap = ew.snapshot()
    assert len(snap) == n0 + 1
    want = np.float32([0, 0, -1, 0, 0, 1, 0, 0, 1, 0, 0, 0, 1.5, 7.0, -2.0, 1.0])   # GetTransformMat: transpose of R, then pos
    assert np.array_equal(snap[n0], want)
    # the other bodies keep the records of the last step
    st = ew.state()
    assert np.array_equal(snap[:n0, 12:15], st["pos"][:n0])
    # moving an existing body without stepping is reflected too
    b3 = ew.body_handle(3)
    ew.L.dBodySetPosition(b3, 9.0, 8.0, 7.0)
    assert np.array_equal(ew.snapshot()[3, 12:15], np.float32([9, 8, 7]))
    ew.close()


def test_boxes_settle_on_the_reference_terrain_mesh_loaded_from_obj(tmp_path):
    """res/grassPlane.obj (266 triangles; committed as tests/golden/grassplane_mesh.npz) goes
    through dGeomTriMeshDataBuildFromOBJB200 in Blender's OBJ dialect and serves as the static terrain: spheres and
    boxes dropped on it come to rest on it, and the contacts equal those of the same mesh passed as arrays."""
    import os
    m = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grassplane_mesh.npz"))
    verts, tris = m["verts"], m["tris"]
    lines = ["# Blender", "o Plane"] + ["v %.6f %.6f %.6f" % tuple(v) for v in verts] + ["vn 0 1 0", "vt 0 0", "s off", "l 100 113"]
    lines += ["f %d/1/1 %d/1/1 %d/1/1" % tuple(t + 1) for t in tris]
    path = tmp_path / "grassPlane.obj"
    path.write_text("\n".join(lines) + "\n")
    L = odeb200.lib()
    rs = np.random.RandomState(2)
    spots = rs.uniform(-25, 25, size=(24, 2))
    finals = []
    for use_obj in (True, False):
        s = Server()
        d = C.c_void_p(L.dGeomTriMeshDataCreate())
        if use_obj:
            assert L.dGeomTriMeshDataBuildFromOBJB200(d, str(path).encode()) == 266
            # %.6f text round trip: parse what the loader parsed so that both runs use the same floats
            nv = C.c_int(0)
            L.dGeomTriMeshDataGetB200(d, None, 0, None, 0, C.byref(nv))
            v = np.zeros((nv.value, 3), np.float32)
            L.dGeomTriMeshDataGetB200(d, v.ctypes.data_as(fp), nv.value, None, 0, C.byref(nv))
            verts = v
        else:
            v = np.ascontiguousarray(verts); t = np.ascontiguousarray(tris)
            L.dGeomTriMeshDataBuildSingle(d, v.ctypes.data, 12, len(v), t.ctypes.data, 3 * len(t), 12)
        g = C.c_void_p(L.dCreateTriMesh(s.space, d, None, None, None))
        s.geoms.append(g)
        bodies = []
        for i, (x, z) in enumerate(spots):
            if i % 2:
                bodies.append(s.add_body((x, 12.0, z), "sphere", (0.8,))[0])
            else:
                bodies.append(s.add_body((x, 12.0, z), "box", (1.2, 0.9, 1.5))[0])
        for _ in range(360):
            s.tick()
        finals.append(np.stack([np.concatenate([s.pos(b), s.vel(b)]) for b in bodies]))
        s.close()
        L.dGeomTriMeshDataDestroy(d)
    assert np.array_equal(finals[0], finals[1])
    y = finals[0][:, 1]
    assert y.min() > verts[:, 1].min() - 0.5 and y.max() < verts[:, 1].max() + 2.0      # on the terrain, not through it
    assert np.abs(finals[0][::2, 3:]).max() < 1.0                                         # the boxes came to rest (spheres may still roll)


def test_getter_pointers_stay_valid_for_the_bodys_lifetime():
    """libode guarantees the const dReal* of dBodyGetPosition & co for the body's lifetime.  The host
    mirrors are address-stable arrays: a pointer taken before 70 000 more bodies are created (the mirrors grow by three
    orders of magnitude) is still THE pointer afterwards, and reads the body's current state after a step."""
    s = Server()
    L = s.L
    s.add_static_box((0, 0, 0), (100, 1, 100))
    b0, g0 = s.add_body((0.25, 3.0, -0.5), "box", (0.4, 0.4, 0.4))
    getters = [L.dBodyGetPosition, L.dBodyGetRotation, L.dBodyGetQuaternion, L.dBodyGetLinearVel, L.dBodyGetAngularVel]
    before = [C.cast(f(b0), C.c_void_p).value for f in getters]
    gp_before = C.cast(L.dGeomGetPosition(s.geoms[0]), C.c_void_p).value   # the static floor's own pose
    p = L.dBodyGetPosition(b0)
    assert (p[0], p[1], p[2]) == (0.25, 3.0, -0.5)
    rs = np.random.RandomState(5)
    for i in range(70000):
        bi = C.c_void_p(L.dBodyCreate(s.world))
        L.dBodySetPosition(bi, float(100 + 2 * (i % 300)), 50.0 + 2 * (i // 300), float(rs.rand()))
        s.bodies.append(bi)
    after = [C.cast(f(b0), C.c_void_p).value for f in getters]
    assert before == after
    assert C.cast(L.dGeomGetPosition(s.geoms[0]), C.c_void_p).value == gp_before
    assert (p[0], p[1], p[2]) == (0.25, 3.0, -0.5)          # the old pointer still reads the body
    s.tick()
    q = L.dBodyGetPosition(b0)                               # refreshes the mirrors after the step
    assert C.cast(q, C.c_void_p).value == before[0]
    assert p[1] < 3.0 and p[1] == q[1]                       # ... and the OLD pointer sees the new state
    s.close()


def test_slot_reuse_gives_the_same_physics_without_growing_the_world():
    """dWorldSetSlotReuseB200: a server that spawns and destroys for ever.  With re-use on, the world's arrays stop
    growing (indices of destroyed bodies / geoms are handed out again, lowest first), and the live bodies move exactly
    as in a world that appends: compared handle by handle over 40 ticks with a spawn + a destroy every other tick."""
    rs = np.random.RandomState(11)
    plan = [(rs.uniform(-1.5, 1.5), 2.0 + rs.uniform(0, 2), rs.uniform(-1.5, 1.5), rs.rand() < 0.5, rs.uniform(0.2, 0.5, 3))
            for _ in range(40)]
    runs = []
    for reuse in (0, 1):
        s = Server()
        L = s.L
        L.dWorldSetSlotReuseB200(s.world, reuse)
        s.add_static_box((0, 0, 0), (100, 1, 100))
        live, trace = [], []
        for t in range(40):
            x, y, z, sphere, d = plan[t]
            if t % 2 == 0:
                live.append(s.add_body((x, y, z), "sphere" if sphere else "box", (d[0],) if sphere else tuple(d)))
            elif len(live) > 3:
                b, g = live.pop(1)
                L.dGeomDestroy(g)
                L.dBodyDestroy(b)
            s.tick()
            trace.append(np.array([s.pos(b) for b, _ in live]))
        runs.append((trace, L.dWorldGetNumBodiesB200(s.world), max(L.dBodyGetIndexB200(b) for b, _ in live)))
        s.close()
    (ta, na, _), (tb, nb, top) = runs
    for a, b in zip(ta, tb):
        assert np.array_equal(a, b)
    assert na == 20 and nb < na and top < nb     # appended 20 slots vs re-used ones


def test_many_worlds_alive_at_once():
    """The address-stable host mirrors reserve address space per world (up to 3 GiB per mirror in use): 300 worlds alive
    at once, each with a body and a geom, must work and keep their getter pointers apart."""
    L = odeb200.lib()
    L.dInitODE()
    worlds = []
    for i in range(300):
        w = C.c_void_p(L.dWorldCreate())
        sp = C.c_void_p(L.dHashSpaceCreate(None))
        b = C.c_void_p(L.dBodyCreate(w))
        L.dBodySetPosition(b, float(i), 1.0, 0.0)
        g = C.c_void_p(L.dCreateSphere(sp, 0.25))
        L.dGeomSetBody(g, b)
        worlds.append((w, sp, b, g))
    ptrs = set()
    for i, (w, sp, b, g) in enumerate(worlds):
        p = L.dBodyGetPosition(b)
        assert (p[0], p[1]) == (float(i), 1.0)
        ptrs.add(C.cast(p, C.c_void_p).value)
    assert len(ptrs) == 300
    for w, sp, b, g in worlds:
        L.dGeomDestroy(g); L.dBodyDestroy(b); L.dSpaceDestroy(sp); L.dWorldDestroy(w)


def test_world_lifetimes_leak_no_device_memory():
    """Forty worlds created, loaded, ticked, read back and destroyed one after the other (the handle path with its mapped
    read-back and pinned staging buffers, and the bulk path): the device's free memory after cycle 40 is what it was
    after cycle 5."""
    import torch
    torch.zeros(1, device="cuda")

    def free():
        torch.cuda.synchronize()
        return torch.cuda.mem_get_info()[0]
    sc = scenes.pile_scene(8, 8, 4, seed=5, spacing=0.6)
    base = None
    for i in range(40):
        w = odeb200.World(gravity=sc["gravity"])
        w.load_scene(sc)
        for _ in range(3):
            w.tick(sc["h"])
        w.state()
        w.close()
        s = Server()
        s.add_static_box((0, 0, 0), (100, 1, 100))
        for k in range(6):
            s.add_body((0.3 * k, 1.0 + 0.2 * k, 0.0), "box", (0.4, 0.4, 0.4))
        for _ in range(3):
            s.tick()
        s.close()
        if i == 4:
            base = free()
    assert free() >= base - (1 << 20), (base, free())
