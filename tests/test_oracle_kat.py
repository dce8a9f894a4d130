"""CPU: known-answer tests that pin the oracle (oracle/ode_oracle.c) to hand-derivable results of
the ODE algorithms it restates.  The reference has no tests or golden
vectors for this path and libode is not available, so these are the oracle's only pins ("parity
unpinned" otherwise)."""
import math

import numpy as np
import pytest

import oracle as O

H = 1.0 / 60.0


def test_free_fall_is_semi_implicit_euler(oracle_lib):
    # v' = v + h g ; y' = y + h v'  (new velocity used for the position)
    w = O.OracleWorld()
    b = w.add_body([0, 10, 0])
    w.add_geom(O.SPHERE, [0.5], body=b)
    w.quickstep(H)
    pos, q, R, lv, av = w.body(b)
    v = np.float32(0) + np.float32(H) * np.float32(1.0) * np.float32(-9.8)
    assert lv[1] == pytest.approx(float(v), rel=1e-6)
    assert pos[1] == pytest.approx(10.0 + H * float(v), rel=1e-6)
    assert np.allclose(q, [1, 0, 0, 0])


def test_sphere_resting_on_plane_erp_pushout(oracle_lib):
    # one contact, 3 decoupled rows; after the solve v_n ~= erp * depth / h
    w = O.OracleWorld()
    w.add_geom(O.PLANE, [0, 1, 0, 0])
    b = w.add_body([0, 0.49, 0])
    w.add_geom(O.SPHERE, [0.5], body=b)
    assert w.tick(H, order_mode=1) == 1
    _, _, _, lv, _ = w.body(b)
    assert lv[1] == pytest.approx(0.2 * 0.01 / H, rel=5e-3)
    assert abs(lv[0]) < 1e-6 and abs(lv[2]) < 1e-6


def test_normal_row_converges_geometrically(oracle_lib):
    # SOR with w = 1.3 on one decoupled row: error shrinks by |1 - w| = 0.3 per iteration
    errs = []
    for iters in (1, 2, 3):
        w = O.OracleWorld(iters=iters, cfm=0.0)
        w.add_geom(O.PLANE, [0, 1, 0, 0])
        b = w.add_body([0, 0.49, 0])
        w.add_geom(O.SPHERE, [0.5], body=b)
        w.collide_all(8)
        w.quickstep(H, order_mode=1)
        lam = w.last_lambda()[0]
        exact = (0.2 * 0.01 / H / H + 9.8)  # rhs / (1/m)
        errs.append(abs(lam - exact))
    assert errs[1] / errs[0] == pytest.approx(0.3, rel=2e-2)
    assert errs[2] / errs[1] == pytest.approx(0.3, rel=5e-2)


def test_sphere_sphere_contact(oracle_lib):
    w = O.OracleWorld()
    b1 = w.add_body([0, 0, 0]); g1 = w.add_geom(O.SPHERE, [1.0], body=b1)
    b2 = w.add_body([1.5, 0, 0]); g2 = w.add_geom(O.SPHERE, [0.75], body=b2)
    c = w.collide(g1, g2)
    assert len(c) == 1
    assert list(c[0].normal) == pytest.approx([-1, 0, 0])     # from g2 into g1
    assert c[0].depth == pytest.approx(0.25)
    # pos = p1 + n * 0.5 * (r2 - r1 - d)
    assert c[0].pos[0] == pytest.approx(0.0 + (-1) * 0.5 * (0.75 - 1.0 - 1.5))
    assert w.collide(g1, g2)[0].g1 == g1
    # coincident centres: fixed normal, depth r1 + r2
    w.set_body_state(b2, pos=[0, 0, 0])
    c = w.collide(g1, g2)
    assert list(c[0].normal) == [1, 0, 0] and c[0].depth == pytest.approx(1.75)
    # separated
    w.set_body_state(b2, pos=[2.0, 0, 0])
    assert len(w.collide(g1, g2)) == 0


def test_dcollide_swaps_for_reversed_class_order(oracle_lib):
    w = O.OracleWorld()
    bs = w.add_body([0, 0.9, 0]); gs = w.add_geom(O.SPHERE, [0.5], body=bs)
    gb = w.add_geom(O.BOX, [2, 1, 2], pos=[0, 0, 0])
    a = w.collide(gs, gb)[0]
    b = w.collide(gb, gs)[0]
    assert list(a.normal) == pytest.approx([0, 1, 0])        # box pushes the sphere up
    assert list(b.normal) == pytest.approx([0, -1, 0])
    assert (a.g1, a.g2) == (gs, gb) and (b.g1, b.g2) == (gb, gs)
    assert a.depth == pytest.approx(0.1) and b.depth == pytest.approx(0.1)


def test_sphere_box_face_edge_corner_inside(oracle_lib):
    w = O.OracleWorld()
    bs = w.add_body([0, 0, 0]); gs = w.add_geom(O.SPHERE, [0.5], body=bs)
    gb = w.add_geom(O.BOX, [2, 2, 2], pos=[0, 0, 0])
    # face
    w.set_body_state(bs, pos=[0.2, 1.4, -0.3])
    c = w.collide(gs, gb)[0]
    assert list(c.normal) == pytest.approx([0, 1, 0]) and c.depth == pytest.approx(0.1)
    assert list(c.pos) == pytest.approx([0.2, 1.0, -0.3])
    # edge
    w.set_body_state(bs, pos=[1.3, 1.3, 0.0])
    c = w.collide(gs, gb)[0]
    s = math.sqrt(0.5)
    assert list(c.normal) == pytest.approx([s, s, 0], abs=1e-6)
    assert c.depth == pytest.approx(0.5 - math.sqrt(0.18), abs=1e-6)
    # corner
    w.set_body_state(bs, pos=[1.2, 1.2, 1.2])
    c = w.collide(gs, gb)[0]
    assert list(c.normal) == pytest.approx([1 / math.sqrt(3)] * 3, abs=1e-6)
    assert list(c.pos) == pytest.approx([1, 1, 1])
    # centre inside the box: pushed out through the nearest face, depth = distance + r
    w.set_body_state(bs, pos=[0.1, 0.8, -0.2])
    c = w.collide(gs, gb)[0]
    assert list(c.normal) == pytest.approx([0, 1, 0]) and c.depth == pytest.approx(0.2 + 0.5)
    assert list(c.pos) == pytest.approx([0.1, 0.8, -0.2])
    # miss
    w.set_body_state(bs, pos=[1.6, 1.6, 0])
    assert len(w.collide(gs, gb)) == 0


def test_box_plane_contact_counts(oracle_lib):
    w = O.OracleWorld()
    gp = w.add_geom(O.PLANE, [0, 1, 0, 0])
    b = w.add_body([0, 0.45, 0]); gb = w.add_geom(O.BOX, [1, 1, 1], body=b)
    c = w.collide(gb, gp)
    assert len(c) == 4                                       # face flat in the plane: 4 corners
    assert all(cc.depth == pytest.approx(0.05) for cc in c)
    assert sorted((round(cc.pos[0], 3), round(cc.pos[2], 3)) for cc in c) == [(-0.5, -0.5), (-0.5, 0.5), (0.5, -0.5), (0.5, 0.5)]
    # tilted about z by 45 degrees: an edge touches -> 2 contacts
    q = [math.cos(math.pi / 8), 0, 0, math.sin(math.pi / 8)]
    w.set_body_state(b, pos=[0, math.sqrt(0.5) - 0.01, 0], q=q)
    assert len(w.collide(gb, gp)) == 2
    # generic orientation, shallow: a single corner
    ax = np.array([1.0, 0.3, 0.7]); ax /= np.linalg.norm(ax)
    ang = 0.9
    q = [math.cos(ang / 2)] + list(math.sin(ang / 2) * ax)
    w.set_body_state(b, pos=[0, 0.0, 0], q=q)
    R = w.body(b)[2].reshape(3, 4)[:, :3]
    reach = 0.5 * np.abs(R[1]).sum()
    w.set_body_state(b, pos=[0, reach - 0.005, 0])
    c = w.collide(gb, gp)
    assert len(c) == 1 and c[0].depth == pytest.approx(0.005, abs=1e-5)
    assert len(w.collide(gb, gp, maxc=3)) <= 3


def test_box_box_face_face_and_edge_edge(oracle_lib):
    w = O.OracleWorld()
    b1 = w.add_body([0, 0, 0]); g1 = w.add_geom(O.BOX, [2, 1, 2], body=b1)
    b2 = w.add_body([0.2, 0.95, 0.1]); g2 = w.add_geom(O.BOX, [1, 1, 1], body=b2)
    c = w.collide(g1, g2)
    assert len(c) == 4
    assert all(list(cc.normal) == pytest.approx([0, -1, 0]) for cc in c)   # from g2 (top) into g1
    assert all(cc.depth == pytest.approx(0.05, abs=1e-6) for cc in c)
    xs = sorted((round(cc.pos[0], 3), round(cc.pos[2], 3)) for cc in c)
    assert xs == [(-0.3, -0.4), (-0.3, 0.6), (0.7, -0.4), (0.7, 0.6)]
    # edge-edge: two unit cubes, the upper rotated 45 deg about x then 45 about z... use crossed edges
    w2 = O.OracleWorld()
    qa = [math.cos(math.pi / 8), math.sin(math.pi / 8), 0, 0]      # 45 deg about x
    qb = [math.cos(math.pi / 8), 0, 0, math.sin(math.pi / 8)]      # 45 deg about z
    a = w2.add_body([0, 0, 0], q=qa); ga = w2.add_geom(O.BOX, [4, 1, 1], body=a)
    d = 2 * math.sqrt(0.5) - 0.02
    b = w2.add_body([0, d, 0], q=[math.cos(math.pi / 4) * qb[0], 0, 0, 0]); w2.set_body_state(b, q=None)
    # upper box: long axis along z, rotated 45 deg about its long axis so an edge points down
    qz = [math.cos(math.pi / 8), 0, 0, math.sin(math.pi / 8)]
    w2.set_body_state(b, pos=[0, d, 0], q=qz)
    gb = w2.add_geom(O.BOX, [1, 1, 4], body=b)
    c = w2.collide(ga, gb)
    assert len(c) == 1
    assert c[0].depth == pytest.approx(0.02, abs=1e-5)
    assert list(c[0].normal) == pytest.approx([0, -1, 0], abs=1e-5)
    assert list(c[0].pos) == pytest.approx([0, d / 2, 0], abs=1e-5)


def test_plane_space_both_branches(oracle_lib):
    L = O.lib()
    for n in ([0, 0, 1], [0, 1, 0], [1, 0, 0], [0.6, 0.0, 0.8], [0.48, 0.6, 0.64]):
        n = np.asarray(n, np.float32); n /= np.linalg.norm(n)
        p = np.zeros(3, np.float32); q = np.zeros(3, np.float32)
        L.orc_plane_space(O._ptr(n), O._ptr(p), O._ptr(q))
        assert abs(np.dot(n, p)) < 1e-6 and abs(np.dot(n, q)) < 1e-6 and abs(np.dot(p, q)) < 1e-6
        assert np.linalg.norm(p) == pytest.approx(1, abs=1e-6) and np.linalg.norm(q) == pytest.approx(1, abs=1e-6)
        assert np.allclose(np.cross(n, p), q, atol=1e-6)
    # |n.z| > sqrt(1/2) branch: p has no x component
    n = np.array([0, 0, 1], np.float32); p = np.zeros(3, np.float32); q = np.zeros(3, np.float32)
    L.orc_plane_space(O._ptr(n), O._ptr(p), O._ptr(q))
    assert list(p) == [0, -1, 0] and list(q) == [1, 0, 0]


def test_quaternion_roundtrip_and_renorm(oracle_lib):
    L = O.lib()
    rs = np.random.RandomState(0)
    for _ in range(50):
        q = rs.normal(size=4).astype(np.float32); q /= np.linalg.norm(q)
        if q[0] < 0:
            q = -q
        R = np.zeros(12, np.float32); q2 = np.zeros(4, np.float32)
        L.orc_q_to_r(O._ptr(q), O._ptr(R)); L.orc_r_to_q(O._ptr(R), O._ptr(q2))
        if q2[0] < 0:
            q2 = -q2
        assert np.allclose(q, q2, atol=2e-6)
        M = R.reshape(3, 4)[:, :3]
        assert np.allclose(M @ M.T, np.eye(3), atol=1e-5)
    # integration renormalises: spin a body for 100 steps
    w = O.OracleWorld(gravity=(0, 0, 0))
    b = w.add_body([0, 0, 0], avel=[3, -2, 5])
    for _ in range(100):
        w.quickstep(H)
    assert np.linalg.norm(w.body(b)[1]) == pytest.approx(1, abs=1e-6)
    assert list(w.body(b)[4]) == pytest.approx([3, -2, 5])            # I = identity: no gyroscopic drift


def test_kinematic_body_ignores_forces_and_pushes(oracle_lib):
    w = O.OracleWorld()
    k = w.add_body([0, 0, 0], lvel=[1, 0, 0], flags=O.BODY_KINEMATIC); w.add_geom(O.SPHERE, [0.5], body=k)
    d = w.add_body([0.9, 0, 0]); w.add_geom(O.SPHERE, [0.5], body=d)
    for _ in range(10):
        w.tick(H)
    assert list(w.body(k)[3]) == [1, 0, 0]
    assert w.body(k)[0][0] == pytest.approx(10 * H, rel=1e-5) and w.body(k)[0][1] == 0
    assert w.body(d)[3][0] > 0.5                                       # shoved along by the kinematic sphere


def test_hash_space_equals_brute_force(oracle_lib):
    from odeb200 import scenes
    for seed in (1, 2, 3):
        sc = scenes.random_soup(300, seed=seed)
        w = O.OracleWorld(); w.load_scene(sc)
        assert np.array_equal(w.broadphase(0), w.broadphase(1))
    sc = scenes.server_scene(y_range=(1.0, 6.0))
    w = O.OracleWorld(); w.load_scene(sc)
    p = w.broadphase(0)
    assert np.array_equal(p, w.broadphase(1))
    # static-static pairs are emitted (both bodies NULL != same body): floor vs the three walls
    stat = {(a, b) for a, b in p.tolist() if a < 4 and b < 4}
    assert stat == {(0, 1), (0, 2), (0, 3), (1, 2), (1, 3)}


def test_category_collide_bits_and_same_body(oracle_lib):
    w = O.OracleWorld()
    b = w.add_body([0, 0, 0])
    g1 = w.add_geom(O.SPHERE, [0.5], body=b, cat=2, col=3)
    g2 = w.add_geom(O.SPHERE, [0.5], body=b, cat=2, col=3)           # same body: never a pair
    g3 = w.add_geom(O.SPHERE, [0.5], pos=[0.2, 0, 0], cat=4, col=8)  # bits do not meet
    g4 = w.add_geom(O.SPHERE, [0.5], pos=[0.2, 0, 0], cat=1, col=0)  # g1.col & g4.cat
    pairs = {tuple(p) for p in w.broadphase(1).tolist()}
    assert (g1, g2) not in pairs and (g1, g3) not in pairs
    assert (g1, g4) in pairs and (g2, g4) in pairs
    # touching AABBs count as overlapping (non-strict test)
    w2 = O.OracleWorld()
    a = w2.add_geom(O.BOX, [1, 1, 1], pos=[0, 0, 0]); c = w2.add_geom(O.BOX, [1, 1, 1], pos=[1, 0, 0])
    assert w2.broadphase(1).tolist() == [[a, c]]


def test_axis_aligned_plane_has_half_space_aabb(oracle_lib):
    w = O.OracleWorld()
    gp = w.add_geom(O.PLANE, [0, 2, 0, 1.0])                          # normalised to (0,1,0,0.5)
    a = w.aabb(gp)
    assert a[3] == pytest.approx(0.5) and np.isinf(a[2]) and np.isinf(a[0]) and np.isinf(a[5])
    hi = w.add_body([0, 5, 0]); w.add_geom(O.SPHERE, [0.5], body=hi)   # entirely above: no pair
    lo = w.add_body([0, 0.9, 0]); g_lo = w.add_geom(O.SPHERE, [0.5], body=lo)
    assert w.broadphase(0).tolist() == [[gp, g_lo]]


def test_snapshot_layout_is_transposed_rotation(oracle_lib):
    w = O.OracleWorld()
    q = np.array([0.9, 0.1, -0.3, 0.2], np.float32); q /= np.linalg.norm(q)
    b = w.add_body([1, 2, 3], q=q)
    t = w.body_transform(b)
    R = w.body(b)[2]
    assert list(t[0:4]) == [R[0], R[4], R[8], 0] and list(t[4:8]) == [R[1], R[5], R[9], 0]
    assert list(t[8:12]) == [R[2], R[6], R[10], 0] and list(t[12:16]) == [1, 2, 3, 1]


def test_reference_prng_stream(oracle_lib):
    import ctypes as C
    from odeb200.scenes import RefRand
    L = O.lib()
    st = C.c_uint(12345)
    seq = [L.orc_rand_next(C.byref(st)) for _ in range(8)]
    r = RefRand(12345)
    assert seq == [int(x) for x in r.next_block(8)]
    # first principles for the first draw
    s = (12345 + 0xE120FC15) & 0xFFFFFFFF
    t = s * 0x4A39B70D
    m1 = ((t >> 32) ^ t) & 0xFFFFFFFF
    t = m1 * 0x12FAD5C9
    assert seq[0] == ((t >> 32) ^ t) & 0xFFFFFFFF


def test_trimesh_sphere_rule(oracle_lib):
    # two triangles forming a unit square in the plane y = 0
    verts = np.array([[0, 0, 0], [1, 0, 0], [1, 0, 1], [0, 0, 1]], np.float32)
    tris = np.array([[0, 2, 1], [0, 3, 2]], np.int32)
    w = O.OracleWorld()
    w.add_mesh(verts, tris)
    gm = w.add_geom(O.TRIMESH, [0])
    b = w.add_body([0.9, 0.3, 0.1]); gs = w.add_geom(O.SPHERE, [0.5], body=b)
    c = w.collide(gs, gm)
    assert len(c) == 1 and c[0].side2 == 0
    assert list(c[0].normal) == pytest.approx([0, 1, 0]) and c[0].depth == pytest.approx(0.2)
    assert list(c[0].pos) == pytest.approx([0.9, 0.0, 0.1])
    # nearer the diagonal the neighbouring triangle is within reach too: a second, shallower contact
    w.set_body_state(b, pos=[0.7, 0.3, 0.2])
    c = w.collide(gs, gm)
    assert [cc.side2 for cc in c] == [0, 1] and c[0].depth > c[1].depth
    # on the shared diagonal both triangles give the same closest point: duplicates are dropped
    w.set_body_state(b, pos=[0.5, 0.3, 0.5])
    c = w.collide(gs, gm)
    assert len(c) == 1 and c[0].side2 == 0
    # beside the square: closest point on the border edge, slanted normal
    w.set_body_state(b, pos=[1.3, 0.2, 0.5])
    c = w.collide(gs, gm)
    assert len(c) == 1
    n = np.array([0.3, 0.2, 0.0]); n /= np.linalg.norm(n)
    assert list(c[0].normal) == pytest.approx(list(n), abs=1e-6)
    assert c[0].depth == pytest.approx(0.5 - math.hypot(0.3, 0.2), abs=1e-6)
    w.set_body_state(b, pos=[1.6, 0.2, 0.5])
    assert len(w.collide(gs, gm)) == 0


def test_one_step_tolerance_is_meaningful_under_fma_rounding(oracle_lib):
    """Why the engine keeps --fmad=false: contracting the solver's dot products into FMAs (what nvcc would do
    by default) moves one-step velocities by up to ~1e-5 relative -- inside north_star's 1e-4 bar after ONE
    step, but the drift compounds over steps.  The GPU path avoids the question by rounding exactly like the
    oracle, so its parity tests use zero tolerance on state as well."""
    import ctypes as C
    from odeb200 import scenes
    L = O.lib()
    L.orc_set_perturb_fma.argtypes = [C.c_void_p, C.c_int]
    sc = scenes.random_soup(400, seed=21, extent=4.5)
    a = O.OracleWorld(); a.load_scene(sc)
    b = O.OracleWorld(); b.load_scene(sc)
    L.orc_set_perturb_fma(b.w, 1)
    a.tick(sc["h"], order_mode=1); b.tick(sc["h"], order_mode=1)
    sa, sb = a.state(), b.state()
    assert a.num_rows > 300
    for k, floor in (("pos", 1.0), ("quat", 1.0), ("lvel", 0.1), ("avel", 0.1)):
        d = np.abs(sa[k].astype(np.float64) - sb[k]) / np.maximum(np.abs(sa[k]), floor)
        assert d.max() < 1e-4, k
    assert np.abs(sa["lvel"] - sb["lvel"]).max() > 0          # ... but it is not bit-identical


def test_exact_lcp_mode_is_the_limit_of_the_sweeps(oracle_lib):
    """order_mode 3 (the dWorldStep restatement: exact LCP by principal pivoting, double precision) must be
    what SOR/PGS converges to -- A = J invM J^T + cfm/h is SPD, so the bounded LCP has one solution -- and
    20 sweeps must be measurably further from it than 2000."""
    from odeb200 import scenes
    sc = scenes.server_scene(seed=1, y_range=(1.0, 6.0))
    dist = {}
    ref = None
    for iters, mode in ((0, 3), (20, 1), (200, 1), (5000, 1)):
        ow = O.OracleWorld(iters=max(iters, 1))
        ow.load_scene(sc)
        nc = ow.collide_all(8, O.reference_surface())
        assert nc > 20
        ow.quickstep(sc["h"], order_mode=mode)
        v = np.concatenate([ow.state()["lvel"], ow.state()["avel"]], axis=1).astype(np.float64)
        if mode == 3:
            ref = v
        else:
            dist[iters] = np.abs(v - ref).max()
    assert dist[5000] < 2e-3, dist
    assert dist[20] > 5 * dist[5000], dist
    assert dist[200] < dist[20], dist


def _quad_mesh(size=4.0):
    v = np.float32([[-size, 0, -size], [size, 0, -size], [size, 0, size], [-size, 0, size]])
    t = np.int32([[0, 2, 1], [0, 3, 2]])          # normals +y
    return v, t


def test_box_trimesh_vertex_face_rule(oracle_lib):
    """box-trimesh (engine-defined rule): a box sunk 0.05 into a flat two-triangle floor gets its four
    lower vertices as contacts, normal +y (from the mesh into the box), depth 0.05; tilted so that one vertex is
    lowest, that vertex comes first (deepest)."""
    w = O.OracleWorld()
    v, t = _quad_mesh()
    m = w.add_mesh(v, t)
    gm = w.add_geom(O.TRIMESH, [m])
    b = w.add_body([0.3, 0.45, -0.2])
    gb = w.add_geom(O.BOX, [1.0, 1.0, 1.0], body=b)
    cs = w.collide(gb, gm, 8)
    assert len(cs) == 4
    for c in cs:
        assert tuple(c.normal)[:3] == (0.0, 1.0, 0.0)
        assert c.depth == pytest.approx(0.05, abs=1e-6)
        assert c.pos[1] == pytest.approx(-0.05, abs=1e-6) and abs(abs(c.pos[0] - 0.3) - 0.5) < 1e-6
        assert c.side2 in (0, 1) and c.side1 == -1
    # argument order swapped: normals negated, sides exchanged (dCollide's contract)
    cs2 = w.collide(gm, gb, 8)
    assert len(cs2) == 4 and tuple(cs2[0].normal)[:3] == (0.0, -1.0, 0.0) and cs2[0].side1 in (0, 1)
    # a box high above the floor: nothing
    w2 = O.OracleWorld()
    m2 = w2.add_mesh(v, t)
    gm2 = w2.add_geom(O.TRIMESH, [m2])
    b2 = w2.add_body([0.0, 0.51, 0.0])
    gb2 = w2.add_geom(O.BOX, [1.0, 1.0, 1.0], body=b2)
    assert len(w2.collide(gb2, gm2, 8)) == 0


def test_box_trimesh_mesh_vertex_inside_box(oracle_lib):
    """a spike of the mesh poking into the bottom face of a box: contact at the spike's tip, pushing the box up
    by the tip's distance to the bottom face"""
    w = O.OracleWorld()
    v = np.float32([[-1, 0, -1], [1, 0, -1], [0, 0, 1], [0, 1.0, 0]])
    t = np.int32([[0, 3, 1], [1, 3, 2], [2, 3, 0]])
    m = w.add_mesh(v, t)
    gm = w.add_geom(O.TRIMESH, [m])
    b = w.add_body([0.0, 1.4, 0.0])
    gb = w.add_geom(O.BOX, [3.0, 1.0, 3.0], body=b)          # bottom face at y = 0.9: the tip is 0.1 inside
    cs = w.collide(gb, gm, 8)
    assert len(cs) == 1
    c = cs[0]
    assert tuple(c.pos)[:3] == (0.0, 1.0, 0.0) and tuple(c.normal)[:3] == (0.0, 1.0, 0.0)
    assert c.depth == pytest.approx(0.1, abs=1e-6)


def test_exact_lcp_single_contact_closed_form(oracle_lib):
    """order_mode 3 on a hand-solvable step: a unit sphere of mass 1 sunk 0.05 into the plane y = 0, at rest.
    One contact, decoupled rows: (1/m + cfm/h) lambda = c/h - (v/h + g.n) with c = min(erp*depth/h, max_vel), so the
    new normal velocity is h*(lambda/m + g.n) = (c + 9.8 h)/(1 + cfm/h) - 9.8 h, and the tangential rows stay 0."""
    w = O.OracleWorld()
    w.add_geom(O.PLANE, [0, 1, 0, 0])
    b = w.add_body([0, 0.45, 0])
    w.add_geom(O.SPHERE, [0.5], body=b)
    assert w.collide_all(8, O.reference_surface()) == 1
    w.quickstep(H, order_mode=3)
    pos, q, R, lv, av = w.body(b)
    c = 0.2 * 0.05 / H
    expect = (c + 9.8 * H) / (1.0 + 1e-5 / H) - 9.8 * H
    assert lv[1] == pytest.approx(expect, rel=2e-6)
    assert abs(lv[0]) < 1e-7 and abs(lv[2]) < 1e-7 and np.abs(av).max() < 1e-7
    # 20 SOR sweeps at w = 1.3 on the same step have not converged to it yet -- the modes really differ
    w2 = O.OracleWorld()
    w2.add_geom(O.PLANE, [0, 1, 0, 0])
    b2 = w2.add_body([0, 0.45, 0])
    w2.add_geom(O.SPHERE, [0.5], body=b2)
    w2.collide_all(8, O.reference_surface())
    w2.quickstep(H, order_mode=1)
    assert abs(w2.body(b2)[3][1] - expect) < 1e-3


def test_exact_lcp_head_on_bounce_closed_form(oracle_lib):
    """two unit spheres (mass 1) head-on at +-1 m/s, 0.02 deep, no gravity, reference surface (bounce 0.2): the row
    target is c = max(erp*depth/h, bounce * closing speed) = 0.4, A = 1/m1 + 1/m2 + cfm/h, so the separation speed
    after the exact step is v_rel + 2 (c - v_rel) / A with v_rel = -2; momentum stays zero."""
    w = O.OracleWorld(gravity=(0.0, 0.0, 0.0))
    b1 = w.add_body([-0.49, 0, 0], lvel=[1.0, 0, 0])
    w.add_geom(O.SPHERE, [0.5], body=b1)
    b2 = w.add_body([0.49, 0, 0], lvel=[-1.0, 0, 0])
    w.add_geom(O.SPHERE, [0.5], body=b2)
    assert w.collide_all(8, O.reference_surface()) == 1
    w.quickstep(H, order_mode=3)
    v1, v2 = w.body(b1)[3], w.body(b2)[3]
    A = 2.0 + 1e-5 / H
    sep = -2.0 + 2.0 * (0.4 + 2.0) / A
    assert (v2[0] - v1[0]) == pytest.approx(sep, rel=1e-5)
    assert v1[0] + v2[0] == pytest.approx(0.0, abs=1e-6)
    assert np.abs(np.concatenate([v1[1:], v2[1:]])).max() < 1e-7
