"""GPU: parity of the CUDA path (through the C ABI of libode_b200.so) with the CPU oracle on the same
seeded inputs, and with the committed golden fixtures.

Bars (BASELINE.json north_star): broadphase pair lists and per-pair contact counts bit-exact as
sorted sets; post-step position / orientation / velocity within relative 1e-4 after one step at
equal iteration count (the oracle is run in the engine's Gauss-Seidel order); bounded
constraint residual over 600 steps."""
import os

import numpy as np
import pytest

import oracle as O
import util
from odeb200 import scenes

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
STATE_RTOL = 1e-4  # north_star: relative 1e-4 in float32 after one step at equal iteration count


def _scene(name):
    return {
        "c1_low": lambda: scenes.server_scene(seed=1, y_range=(1.0, 6.0)),
        "c1p_low": lambda: scenes.server_scene(seed=1, y_range=(1.0, 6.0), floor_plane=True),
        "soup200": lambda: scenes.random_soup(200, seed=7),
        "soup_axis": lambda: scenes.random_soup(150, seed=9, rotated=False),
        "teapot256": lambda: scenes.trimesh_contact_scene(256, seed=11),
        "batch8": lambda: scenes.batched_worlds_scene(8, seed=4, spacing=0.7),
        "teapot_boxes": lambda: scenes.trimesh_contact_scene(192, seed=13, box_fraction=0.6),
        "soup1500": lambda: scenes.random_soup(1500, seed=21, extent=7.0),
        "pile_dense": lambda: scenes.pile_scene(12, 12, 6, seed=5, spacing=0.6),
    }[name]()


GOLDEN = ["c1_low", "c1p_low", "soup200", "soup_axis", "teapot256", "batch8", "teapot_boxes"]
ALL = GOLDEN + ["soup1500", "pile_dense"]


def _engine_contacts_by_pair(ew):
    pr, cnt, pd, nrm, side = ew.contacts()
    out, f = {}, 0
    for (a, b), c in zip(pr.tolist(), cnt.tolist()):
        out[(a, b)] = (pd[f:f + c], nrm[f:f + c], side[f:f + c])
        f += c
    return out


@pytest.mark.parametrize("name", GOLDEN)
def test_pairs_and_contacts_match_golden_bit_exact(name):
    gold = np.load(os.path.join(HERE, "golden", name + ".npz"))
    ew = util.engine_world(_scene(name))
    ew.collide(8)
    assert np.array_equal(util.sorted_pair_set(ew.pairs()), gold["pairs"].astype(np.int64))
    got = _engine_contacts_by_pair(ew)
    f = 0
    for (g1, g2), c in zip(gold["canon"].tolist(), gold["count"].tolist()):
        pd, nrm, side = got[(g1, g2)]
        assert len(pd) == c, (g1, g2)
        assert np.array_equal(pd, gold["pos_depth"][f:f + c]), (g1, g2)
        assert np.array_equal(nrm, gold["normal"][f:f + c]), (g1, g2)
        assert np.array_equal(side, gold["side"][f:f + c]), (g1, g2)
        f += c
    assert ew.stats()["flags"] == 0
    ew.close()


@pytest.mark.parametrize("name", ALL)
def test_tick_parity_with_oracle(name):
    sc = _scene(name)
    ow, ew = util.load_both(sc)
    for step in range(4):
        ew.collide(8)
        ep = util.sorted_pair_set(ew.pairs())
        op = util.sorted_pair_set(ow.broadphase(0))
        assert np.array_equal(ep, op), "step %d: broadphase pair sets differ" % step
        got = _engine_contacts_by_pair(ew)
        ref = util.oracle_contacts(ow)
        assert set(got) == set(ref)
        for key, cs in ref.items():
            pd, nrm, side = got[key]
            assert len(pd) == len(cs), (step, key)                      # contact counts: exact
            for k, c in enumerate(cs):
                assert np.array_equal(pd[k], np.array(list(c.pos) + [c.depth], np.float32)), (step, key, k)
                assert np.array_equal(nrm[k], np.array(list(c.normal), np.float32)), (step, key, k)
        ew.step(sc["h"])
        util.oracle_tick_in_engine_order(ow, ew, sc["h"])
        es, os_ = ew.state(), ow.state()
        for k in ("pos", "quat", "lvel", "avel"):
            assert util.rel_err(es[k], os_[k]).max() <= STATE_RTOL, (step, k)
        st = ew.stats()
        assert st["flags"] == 0 and st["n_overflow"] == 0
    ew.close()


def test_snapshot_is_the_reference_transform_layout():
    sc = _scene("soup200")
    ow, ew = util.load_both(sc)
    ew.tick(sc["h"])
    util.oracle_tick_in_engine_order(ow, ew, sc["h"])
    snap = ew.snapshot()
    ref = np.stack([ow.body_transform(i) for i in range(ow.num_bodies)])
    assert np.allclose(snap, ref, rtol=STATE_RTOL, atol=1e-6)
    st = ew.state()
    R = st["R"].reshape(-1, 3, 4)
    assert np.array_equal(snap[:, 0:3], R[:, :, 0]) and np.array_equal(snap[:, 4:7], R[:, :, 1])   # transpose of R
    assert np.array_equal(snap[:, 12:15], st["pos"]) and (snap[:, 15] == 1).all() and (snap[:, 3] == 0).all()
    # partial window copy
    assert np.array_equal(ew.snapshot(first=10, count=5), snap[10:15])
    ew.close()


def test_engine_is_deterministic():
    sc = _scene("soup1500")
    outs = []
    for _ in range(2):
        ew = util.engine_world(sc)
        for _ in range(10):
            ew.tick(sc["h"])
        outs.append(ew.state())
        ew.close()
    for k in outs[0]:
        assert np.array_equal(outs[0][k], outs[1][k]), k


def test_batched_worlds_are_independent_bit_exact():
    """World w inside a batch evolves exactly as the same world stepped alone: the property that makes
    sharding worlds over GPUs collective-free and bit-reproducible."""
    batch = scenes.batched_worlds_scene(6, seed=4, spacing=0.7)
    ew = util.engine_world(batch)
    for _ in range(25):
        ew.tick(batch["h"])
    sb = ew.state()
    st = ew.stats()
    assert st["n_contacts"] > 100
    ew.close()
    for w in (0, 3, 5):
        alone = scenes.batched_worlds_scene(1, seed=4, spacing=0.7, first_world=w)
        e1 = util.engine_world(alone)
        e1.set_contact_units(1)          # same Gauss-Seidel unit as the batch default
        for _ in range(25):
            e1.tick(alone["h"])
        s1 = e1.state()
        e1.close()
        for k in ("pos", "quat", "lvel", "avel"):
            assert np.array_equal(sb[k][w * 128:(w + 1) * 128], s1[k]), (w, k)


def test_world_batches_ticked_side_by_side_equal_one_batch():
    """bench.py holds the worlds of a GPU in several dWorld objects whose ticks overlap on the GPU (one stream each).  The
    worlds do not care how they are grouped: 24 worlds as 3 dWorlds of 8, ticked interleaved without waiting, end in the
    very bits of one dWorld of 24."""
    one = scenes.batched_worlds_scene(24, seed=4, spacing=0.7)
    e1 = util.engine_world(one)
    parts = []
    for k in range(3):
        sc = scenes.batched_worlds_scene(8, seed=4, spacing=0.7, first_world=8 * k)
        parts.append(util.engine_world(sc))
    for _ in range(30):
        e1.tick(one["h"])
        for p in parts:
            p.tick(one["h"])      # queued on three streams; nobody waits
    s1 = e1.state()
    for k, p in enumerate(parts):
        sp = p.state()
        for f in ("pos", "quat", "lvel", "avel"):
            assert np.array_equal(s1[f][k * 1024:(k + 1) * 1024], sp[f]), (k, f)
        p.close()
    assert e1.stats()["n_contacts"] > 500
    e1.close()


@pytest.mark.parametrize("name", ["batch8", "c1_low", "soup200"])
def test_env_broadphase_equals_grid_broadphase(name):
    """The all-pairs-per-env sweep and the uniform-grid sweep must hand the narrowphase the same pair SET
    (dxHashSpace::collide's, per the golden test) and therefore the same contacts; only list order differs."""
    sc = _scene(name)
    got = []
    for mode in (0, 1):
        ew = util.engine_world(sc)
        ew.set_broadphase(mode)
        ew.collide()      # same input state for both (a tick would let the list order reach the last bits)
        by_pair = _engine_contacts_by_pair(ew)
        st = ew.stats()
        got.append((by_pair, st))
        ew.close()
    (a, sa), (b, sb) = got
    assert sa["n_pairs"] == sb["n_pairs"] and sa["n_contacts"] == sb["n_contacts"]
    assert sorted(a) == sorted(b)
    for k in a:
        for x, y in zip(a[k], b[k]):
            assert np.array_equal(x, y), k
    if name != "batch8":
        assert sa["grid_dims"][0] > 0 and sb["grid_dims"][0] == 0  # the second run really took the env path


def _max_penetration(ew, sc):
    """deepest contact among pairs that involve a dynamic body (the static map boxes overlap each
    other, and a kinematic player sphere sits inside the slanted wall by construction)"""
    gb = sc["geoms"]["body"]
    dyn_body = (sc["bodies"]["flags"] & scenes.BODY_KINEMATIC) == 0
    geom_dyn = (gb >= 0) & dyn_body[np.maximum(gb, 0)]
    pr, cnt, pd, nrm, side = ew.contacts()
    active = geom_dyn[pr[:, 0]] | geom_dyn[pr[:, 1]]
    per_contact = np.repeat(active, cnt)
    d = pd[per_contact, 3]
    return float(d.max()) if len(d) else 0.0


def test_c1_at_rest_constraint_residual():
    """After settling, the solved velocities satisfy the contact rows: normal relative velocity at
    every active contact stays within the ERP push-out bound and bodies stop moving."""
    sc = scenes.server_scene(seed=1, y_range=(1.0, 4.0))
    ew = util.engine_world(sc)
    for _ in range(600):
        ew.tick(sc["h"])
    ew.collide(8)
    pen = _max_penetration(ew, sc)
    s = ew.state()
    assert pen < 0.02
    assert np.abs(s["lvel"][:64]).max() < 0.25 and np.abs(s["avel"][:64]).max() < 2.5
    ew.close()


def test_full_size_pile_properties():
    """BASELINE config 3 at full size (1,048,576 bodies): size-independent properties -- no capacity
    overflow, finite state, unit quaternions, nothing below the floor beyond the contact slop, and
    the pair list is duplicate-free and canonical."""
    sc = scenes.pile_scene(256, 256, 16, seed=3)
    ew = util.engine_world(sc)
    for _ in range(120):
        ew.tick(sc["h"])
    ew.collide(8)
    st = ew.stats()
    assert st["flags"] == 0 and st["n_overflow"] == 0 and st["n_pairs"] > 100000
    pr = ew.pairs()
    key = pr[:, 0].astype(np.int64) * (1 << 32) + pr[:, 1]
    assert len(np.unique(key)) == len(key)
    types = sc["geoms"]["type"]
    assert (types[pr[:, 0]] <= types[pr[:, 1]]).all()
    ew.step(sc["h"])
    s = ew.state()
    assert np.isfinite(s["pos"]).all() and np.isfinite(s["lvel"]).all()
    assert np.allclose(np.linalg.norm(s["quat"], axis=1), 1.0, atol=1e-5)
    assert s["pos"][:, 1].min() > -0.5
    assert np.abs(s["pos"][:, 0]).max() < 0.5 * 256 * 1.8 + 2.0
    ew.close()


def test_full_size_batched_worlds_copies_agree_bit_for_bit():
    """BASELINE config 4 at full size (8192 worlds x 128 bodies = 1,048,576 bodies): the batch is 128 tiles of
    the same 64 worlds, so after 60 ticks every tile must equal tile 0 bit for bit -- each world's result
    depends on nothing but that world, whichever CTA, warp and L2 slice happened to process it."""
    base = scenes.batched_worlds_scene(64, seed=4)
    tiles, per_world, nw = 128, 128, 64
    nb = nw * per_world
    b, g = base["bodies"], base["geoms"]
    bodies = {k: np.concatenate([v] * tiles) for k, v in b.items()}
    bodies["env"] = (np.arange(tiles * nb) // per_world).astype(np.int32)
    dyn = {k: v[1:] for k, v in g.items()}                      # geom 0 is the shared plane
    geoms = {k: np.concatenate([g[k][:1]] + [dyn[k]] * tiles) for k in g}
    geoms["body"] = np.concatenate([[-1], np.arange(tiles * nb)]).astype(np.int32)
    geoms["env"] = np.concatenate([[-1], np.arange(tiles * nb) // per_world]).astype(np.int32)
    sc = scenes.from_arrays("C4-tiled", bodies, geoms, h=base["h"])
    ew = util.engine_world(sc)
    for _ in range(60):
        ew.tick(sc["h"])
    st = ew.stats()
    assert st["flags"] == 0 and st["n_overflow"] == 0 and st["n_contacts"] > 500000
    s = ew.state()
    ew.close()
    for k in ("pos", "quat", "lvel", "avel"):
        a = s[k].reshape(tiles, nb, -1)
        assert np.isfinite(a).all()
        assert (a == a[0:1]).all(), k
    assert s["pos"][:, 1].min() > -0.05


def test_mass_and_inertia_parity():
    """dMassSetBox-style inertia (non-identity, gyroscopic term active) against the oracle."""
    rs = np.random.RandomState(3)
    sc = scenes.random_soup(120, seed=33)
    b = sc["bodies"]
    g = sc["geoms"]
    for gi in range(len(g["type"])):
        bi = g["body"][gi]
        if bi < 0:
            continue
        dens = 2.0 + rs.rand()
        if g["type"][gi] == scenes.BOX:
            lx, ly, lz = g["dims"][gi][:3]
            m = dens * lx * ly * lz
            I = np.diag([m / 12 * (ly * ly + lz * lz), m / 12 * (lx * lx + lz * lz), m / 12 * (lx * lx + ly * ly)])
        else:
            r = g["dims"][gi][0]
            m = dens * 4.0 / 3.0 * np.pi * r ** 3
            I = np.eye(3) * 0.4 * m * r * r
        b["mass"][bi] = m
        b["inertia"][bi] = I.reshape(9)
        b["flags"][bi] = scenes.BODY_GYRO
    ow, ew = util.load_both(sc)
    for step in range(3):
        ew.tick(sc["h"])
        util.oracle_tick_in_engine_order(ow, ew, sc["h"])
        es, os_ = ew.state(), ow.state()
        for k in ("pos", "quat", "lvel", "avel"):
            assert util.rel_err(es[k], os_[k]).max() <= STATE_RTOL, (step, k)
    ew.close()


def test_island_path_with_kinematic_heavy_and_gyroscopic_bodies():
    """Batched worlds on the lane-pair island solver with everything a body can be: kinematic bodies ploughing through
    the heap (their rows are one-body rows: a kinematic end never receives an impulse), masses over two decades, box
    inertias with the gyroscopic term, bodies that ignore gravity.  Oracle of worlds 0, 3 and 7, stepped in the
    engine's order, for 24 ticks (the heaps reach the plane): state within the tolerance (in fact bit-equal)."""
    rs = np.random.RandomState(17)
    nw, per = 8, 128
    sc = scenes.batched_worlds_scene(nw, seed=40, spacing=0.7)
    b, g = sc["bodies"], sc["geoms"]
    body_geom = {int(bi): gi for gi, bi in enumerate(g["body"]) if bi >= 0}
    for bi in range(nw * per):
        gi = body_geom[bi]
        k = bi % per
        if k % 16 == 5:  # kinematic, moving sideways through the lattice
            b["flags"][bi] = scenes.BODY_KINEMATIC
            b["lvel"][bi] = (1.5 * (rs.rand() - 0.5), 0.0, 1.5 * (rs.rand() - 0.5))
            b["avel"][bi] = (0.0, 2.0 * (rs.rand() - 0.5), 0.0)
            continue
        m = float(10.0 ** rs.uniform(-1.0, 1.0))
        if g["type"][gi] == scenes.BOX:
            lx, ly, lz = g["dims"][gi][:3]
            I = np.diag([m / 12 * (ly * ly + lz * lz), m / 12 * (lx * lx + lz * lz), m / 12 * (lx * lx + ly * ly)])
        else:
            r = g["dims"][gi][0]
            I = np.eye(3) * 0.4 * m * r * r
        b["mass"][bi] = m
        b["inertia"][bi] = I.reshape(9)
        b["flags"][bi] = scenes.BODY_GYRO | (scenes.BODY_NOGRAVITY if k % 16 == 9 else 0)
        b["avel"][bi] = 3.0 * (rs.rand(3) - 0.5)
    ew = util.engine_world(sc)
    subs = [(w,) + _world_of_batch(sc, w, per) for w in (0, 3, 7)]
    for step in range(24):
        ew.tick(sc["h"])
        st = ew.stats()
        assert st["flags"] == 0 and st["env_trips"] > 0  # the lane-pair island solver ran
        order = ew.solver_order()
        es = ew.state()
        for w, sub, ow, gmap in subs:
            util.oracle_tick_in_engine_order(ow, ew, sc["h"], geom_map=gmap, order=order)
            os_ = ow.state()
            sl = slice(w * per, (w + 1) * per)
            for k in ("pos", "quat", "lvel", "avel"):
                assert util.rel_err(es[k][sl], os_[k]).max() <= STATE_RTOL, (step, w, k)
    assert st["n_rows1"] > 0 and st["n_rows2"] > 0
    ew.close()


def _world_of_batch(sc, w, per):
    """(sub-scene, oracle world, engine geom id -> oracle geom id) of world w of a batched scene."""
    b, g = sc["bodies"], sc["geoms"]
    keep_g = np.where((g["env"] == w) | (g["env"] < 0))[0]
    bodies = {k: v[w * per:(w + 1) * per].copy() for k, v in b.items()}
    bodies["env"][:] = 0
    geoms = {k: v[keep_g].copy() for k, v in g.items()}
    geoms["body"] = np.where(geoms["body"] >= 0, geoms["body"] - w * per, -1).astype(np.int32)
    geoms["env"] = np.where(geoms["env"] >= 0, 0, -1).astype(np.int32)
    sub = scenes.from_arrays("world%d" % w, bodies, geoms, h=sc["h"])
    ow = util.oracle_world(sub)
    ow._types = [int(t) for t in sub["geoms"]["type"]]
    ow._bodies = [int(x) for x in sub["geoms"]["body"]]
    gmap = -np.ones(len(g["type"]), np.int64)
    gmap[keep_g] = np.arange(len(keep_g))
    return sub, ow, gmap


@pytest.mark.parametrize("group", [8, 16, 32])
def test_island_solver_equals_grid_barrier_solver(group):
    """Batched worlds take the island solver (one lane group per env, no grid barriers); it must give
    the same bits as the global graph-coloured solver."""
    sc = scenes.batched_worlds_scene(24, seed=4, spacing=0.7)
    outs = []
    for mode, g in ((1, 0), (0, group)):
        ew = util.engine_world(sc)
        ew.set_solver_mode(mode, g)
        for _ in range(30):
            ew.tick(sc["h"])
        outs.append((ew.state(), ew.stats()))
        ew.close()
    (a, sa), (b, sb) = outs
    for k in ("pos", "quat", "lvel", "avel"):
        assert np.array_equal(a[k], b[k]), k
    for k in ("n_contacts", "n_manifolds", "n_rows1", "n_rows2", "n_colours"):
        assert sa[k] == sb[k], k
    assert sa["n_contacts"] > 500


def _surface(mode=0, **kw):
    import oracle as O
    import odeb200
    so, se = O.Surface(), odeb200.SurfaceParameters()
    so.mode = se.mode = mode
    for k, v in kw.items():
        setattr(so, k, v)
        setattr(se, k, v)
    return so, se


def _surface_case(case):
    B, MU2, SOFT_ERP, SOFT_CFM = 0x004, 0x001, 0x008, 0x010
    M1, M2, MN, S1, S2, A1 = 0x020, 0x040, 0x080, 0x100, 0x200, 0x3000
    return {
        "mu0": lambda: _surface(B, mu=0.0, bounce=0.1, bounce_vel=0.05),
        "mu_finite": lambda: _surface(B, mu=0.7, bounce=0.3, bounce_vel=0.2),
        "mu2": lambda: _surface(MU2, mu=0.9, mu2=0.2),
        "approx1": lambda: _surface(A1 | B, mu=0.5, bounce=0.2, bounce_vel=0.1),
        "soft": lambda: _surface(SOFT_ERP | SOFT_CFM | B, mu=float("inf"), soft_erp=0.5, soft_cfm=1e-3, bounce=0.2, bounce_vel=0.1),
        "slip_motion": lambda: _surface(S1 | S2 | M1 | M2 | MN, mu=2.0, slip1=0.01, slip2=0.02, motion1=0.1, motion2=-0.2, motionN=0.05),
    }[case]()


SURFACE_CASES = ["mu0", "mu_finite", "mu2", "approx1", "soft", "slip_motion"]


@pytest.mark.parametrize("case", SURFACE_CASES)
def test_surface_modes_parity(case):
    """dSurfaceParameters variants beyond the reference's (bounce, mu=inf): frictionless (1 row),
    box friction, mu2, friction pyramid approximation (findex), soft ERP/CFM, slip and motion."""
    so, se = _surface_case(case)
    sc = scenes.random_soup(160, seed=13)
    ow, ew = util.load_both(sc)
    ew.set_surface(se)
    for step in range(3):
        ew.tick(sc["h"])
        util.oracle_tick_in_engine_order(ow, ew, sc["h"], surf=so, rows_per_contact=1 if case == "mu0" else 3)
        es, os_ = ew.state(), ow.state()
        for k in ("pos", "quat", "lvel", "avel"):
            assert util.rel_err(es[k], os_[k]).max() <= STATE_RTOL, (case, step, k)
    ew.close()


@pytest.mark.parametrize("case", SURFACE_CASES)
def test_surface_modes_parity_on_the_island_path(case):
    """the same surfaces on batched worlds: the lane-pair island solver (solver_env.cu) bakes the world's one surface
    into its row records (rows per contact, the three CFMs, friction limits) -- against the oracle, world by world"""
    so, se = _surface_case(case)
    sc = scenes.batched_worlds_scene(6, seed=4, spacing=0.7)
    ow, ew = util.load_both(sc)
    ew.set_surface(se)
    for step in range(12):
        ew.tick(sc["h"])
        util.oracle_tick_in_engine_order(ow, ew, sc["h"], surf=so, rows_per_contact=1 if case == "mu0" else 3)
        es, os_ = ew.state(), ow.state()
        for k in ("pos", "quat", "lvel", "avel"):
            assert util.rel_err(es[k], os_[k]).max() <= STATE_RTOL, (case, step, k)
    st = ew.stats()
    assert st["n_contacts"] > 30 and st["flags"] == 0
    ew.close()


def test_full_size_trimesh_scene():
    """BASELINE config 2 at full size: teapot.obj (8884 triangles) vs 10,000 spheres.  Contact counts of a
    sample of sphere-mesh pairs are checked against the oracle; the rest through properties."""
    sc = scenes.trimesh_scene(100, seed=2)
    ow, ew = util.load_both(sc)
    for _ in range(150):
        ew.tick(sc["h"])
    ew.collide(8)
    st = ew.stats()
    assert st["flags"] == 0 and st["class_count"][5] > 1000            # sphere-trimesh pairs
    pr, cnt, pd, nrm, side = ew.contacts()
    s = ew.state()
    for i in range(len(s["pos"])):
        ow.set_body_state(i, pos=s["pos"][i], q=s["quat"][i], lvel=s["lvel"][i], avel=s["avel"][i])
    types = sc["geoms"]["type"]
    first = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    mesh_pairs = [i for i in range(len(pr)) if types[pr[i, 1]] == scenes.TRIMESH and cnt[i] > 0]
    assert len(mesh_pairs) > 200
    ntri = 8884
    for i in mesh_pairs[::max(1, len(mesh_pairs) // 150)]:
        ref = ow.collide(int(pr[i, 0]), int(pr[i, 1]), 8)
        assert len(ref) == cnt[i]
        for k, c in enumerate(ref):
            assert np.array_equal(pd[first[i] + k], np.array(list(c.pos) + [c.depth], np.float32))
            assert side[first[i] + k] == c.side2 and 0 <= c.side2 < ntri
    assert np.isfinite(s["pos"]).all() and s["pos"][:, 1].min() > -2.0     # nothing fell through the mesh + plane
    ew.close()


def test_forces_and_async_snapshot_pipeline():
    """dWorldSetForcesB200 (uploaded on its own stream, applied at the next step) and the non-blocking,
    double-buffered snapshot copy give the same numbers as the oracle with dBodyAddForce semantics."""
    import ctypes as C
    import oracle as O
    sc = scenes.random_soup(100, seed=17, with_plane=False, with_static_box=False, extent=20.0)  # sparse: no contacts
    ow, ew = util.load_both(sc)
    n = len(sc["bodies"]["pos"])
    rs = np.random.RandomState(0)
    snaps = [np.zeros((n, 16), np.float32) for _ in range(2)]
    for step in range(6):
        f6 = rs.uniform(-5, 5, size=(n, 6)).astype(np.float32)
        ew.L.dWorldSetForcesB200(ew.w, f6.ctypes.data_as(C.POINTER(C.c_float)), n)
        for i in range(n):
            ow.L.orc_add_force(ow.w, i, O._ptr(f6[i, :3].copy()), O._ptr(f6[i, 3:].copy()))
        ew.tick(sc["h"])
        ew.L.dWorldGetSnapshotB200(ew.w, snaps[step & 1].ctypes.data_as(C.c_void_p), 0, n, 0)   # non-blocking
        ow.tick(sc["h"], order_mode=1)
        if step >= 1:
            ew.wait()
            ref = np.stack([ow.body_transform(i) for i in range(n)])
            assert np.array_equal(snaps[step & 1], ref), step
    es, os_ = ew.state(), ow.state()
    for k in ("pos", "quat", "lvel", "avel"):
        assert np.array_equal(es[k], os_[k]), k
    ew.close()


def _crowded_scene(n_small=90, batched=False):
    """One big dynamic slab carrying n_small spheres: the slab has more than 64 contact neighbours, so some of
    its manifolds cannot get one of the 64 colours and go through the solvers' serial overflow class."""
    sc = scenes._empty_scene("crowded")
    scenes._add_geom(sc, scenes.PLANE, (0, 1, 0, 0.0), cat=scenes.CMASK_ALL & ~scenes.CMASK_MAP)
    n_env = 2 if batched else 1
    for e in range(n_env):
        slab = scenes._add_body(sc, (0.0, 0.26, 0.0), mass=50.0, inertia=(400, 0, 0, 0, 800, 0, 0, 0, 400), env=e)
        scenes._add_geom(sc, scenes.BOX, (10.0, 0.5, 10.0), body=slab, cat=scenes.CMASK_OBJ, col=scenes.CMASK_OBJ | scenes.CMASK_MAP, env=e)
        k = 0
        for ix in range(10):
            for iz in range(10):
                if k >= n_small:
                    break
                b = scenes._add_body(sc, (-4.5 + ix * 1.0, 0.51 + 0.29, -4.5 + iz * 1.0), env=e)
                scenes._add_geom(sc, scenes.SPHERE, (0.3,), body=b, cat=scenes.CMASK_OBJ, col=scenes.CMASK_OBJ | scenes.CMASK_MAP, env=e)
                k += 1
    return scenes.finalize(sc)


@pytest.mark.parametrize("batched", [False, True, "contacts"])
def test_more_than_64_neighbours_overflow_class(batched):
    sc = _crowded_scene(90, bool(batched))
    ow, ew = util.load_both(sc)
    if batched is True:
        ew.set_contact_units(0)          # manifold units on the island path too: 90 + 4 manifolds on the slab
    # "contacts": the batched default -- per-contact units, the lane-pair island solver's serial overflow class
    for step in range(3):
        ew.tick(sc["h"])
        st = ew.stats()
        assert st["n_overflow"] > 0 and st["n_colours"] == 64 and st["flags"] == 0
        util.oracle_tick_in_engine_order(ow, ew, sc["h"])
        es, os_ = ew.state(), ow.state()
        for k in ("pos", "quat", "lvel", "avel"):
            assert util.rel_err(es[k], os_[k]).max() <= STATE_RTOL, (step, k)
    ew.close()


def test_capacity_overflow_is_flagged_not_silent():
    sc = scenes.random_soup(300, seed=3, extent=3.0)
    ew = util.engine_world(sc)
    ew.set_capacity(64, 32)              # far too small on purpose
    ew.tick(sc["h"])
    st = ew.stats()
    assert st["flags"] & 1 and st["n_pairs"] == 64          # SF_PAIR_OVERFLOW, list truncated at capacity
    s = ew.state()
    assert np.isfinite(s["pos"]).all()
    ew.close()


def test_unit_capacity_overflow_on_the_island_path_is_flagged_not_silent():
    """batched worlds whose contacts outnumber the solver-unit arrays: the step says so (flag 2) and moves the bodies
    without contact forces for that tick instead of writing past the arrays"""
    sc = scenes.batched_worlds_scene(12, seed=4, spacing=0.7)
    ew = util.engine_world(sc)
    ew.set_capacity(0, 200)              # ~1200 contacts form in these 12 dense worlds
    for _ in range(4):
        ew.tick(sc["h"])
    st = ew.stats()
    assert st["flags"] & 2 and st["n_contacts"] == 0, st
    s = ew.state()
    assert np.isfinite(s["pos"]).all() and np.isfinite(s["lvel"]).all()
    ew.close()


def test_dworldstep_parity_mode_converges_to_the_exact_lcp():
    """the reference calls dWorldStep (src/main.c:213), libode's exact Dantzig stepper.
    dWorldSetStepSolverB200(world, n > 0, tol) makes the engine's dWorldStep run residual-terminated sweeps; their
    distance from the oracle's exact LCP restatement (order_mode 3) must shrink with the sweep budget, the tolerance
    must end the sweeps early, and the DEFAULT dWorldStep -- the exact solve on the device (solver_exact.cu) -- must
    land on the oracle's answer.  Distances are velocities after one C1 tick (m/s, rad/s)."""
    sc = _scene("c1_low")
    ow = util.oracle_world(sc)
    ow.collide_all(8, O.reference_surface())
    ow.quickstep(sc["h"], order_mode=3)
    so = ow.state()
    ref = np.concatenate([so["lvel"], so["avel"]], axis=1).astype(np.float64)
    dist, used, status = {}, {}, {}
    for iters, tol in ((-1, 0.0), (100, 0.0), (2000, 0.0), (20000, 1e-4), (0, 0.0)):
        ew = util.engine_world(sc)
        ew.set_step_solver(iters, tol)
        ew.collide()
        ew.world_step(sc["h"])
        se = ew.state()
        v = np.concatenate([se["lvel"], se["avel"]], axis=1).astype(np.float64)
        dist[iters] = float(np.abs(v - ref).max())
        st = ew.stats()
        used[iters], status[iters] = st["solver_iters"], st["exact_status"]
        if iters == 0:
            assert st["n_islands"] >= 3 and 3 <= st["max_island_rows"] <= 384 and st["pivot_rounds"] >= 1
        ew.close()
    print("dWorldStep: max |v - v_exact| by mode (-1 QuickStep, n sweeps, 0 exact):", dist, "sweeps used:", used)
    assert used[-1] == 20 and used[100] == 100 and used[2000] == 2000 and status[-1] == -1 and status[100] == -1
    assert 20 < used[20000] < 20000                 # the tolerance ended the sweeps
    assert dist[-1] > 0.01                          # 20 QuickStep sweeps are visibly not the LCP solution
    assert dist[100] < 0.2 * dist[-1]
    assert dist[2000] < 2e-3 and dist[20000] < 2e-3
    assert status[0] == 0 and used[0] == 0 and dist[0] <= 1e-4     # default dWorldStep: the exact answer


@pytest.mark.parametrize("h", [1.0 / 60.0, 1.0 / 120.0])
def test_default_dworldstep_follows_the_exact_oracle_tick_by_tick(h):
    """the reference's loop as it is written -- dSpaceCollide + dWorldStep(world, 1/120) (src/main.c:208-213) -- on its own
    scene: every tick starts from the engine's state, the oracle solves that tick's LCP exactly (order_mode 3), and the
    engine's dWorldStep must agree within 1e-4 on positions and 2e-4 m/s on velocities.  Ticks whose islands exceed
    the exact solver's limits must say so (exact_status 1) and equal dWorldQuickStep."""
    sc = scenes.server_scene(seed=1, h=h, y_range=(1.0, 9.0))
    ow, ew = util.load_both(sc)
    exact = fallback = 0
    worst = 0.0
    for step in range(150):
        pre = ew.state()
        for i in range(len(pre["pos"])):
            ow.set_body_state(i, pos=pre["pos"][i], q=pre["quat"][i], lvel=pre["lvel"][i], avel=pre["avel"][i])
        ew.collide(8)
        ew.world_step(h)
        st = ew.stats()
        assert st["exact_status"] in (0, 1), st
        ow.clear_contacts()
        ow.collide_all(8, O.reference_surface())
        if st["exact_status"] == 0:
            exact += 1
            ow.quickstep(h, order_mode=3)
            es, os_ = ew.state(), ow.state()
            for k in ("lvel", "avel"):
                d = float(np.abs(es[k].astype(np.float64) - os_[k]).max())
                worst = max(worst, d)
                assert d <= 2e-4, (step, k, d)
            assert util.rel_err(es["pos"], os_["pos"]).max() <= 1e-4 and util.rel_err(es["quat"], os_["quat"]).max() <= 1e-4
        else:
            fallback += 1
        ow.clear_contacts()
    print("default dWorldStep, h=1/%d: %d exact ticks (worst |dv| %.2e), %d fell back to the sweeps" % (round(1 / h), exact, worst, fallback))
    assert exact >= 140
    ew.close()


def test_dworldstep_falls_back_to_the_sweeps_when_an_island_is_too_large():
    """a dense pile is one island of thousands of rows: dWorldStep reports exact_status 1 and does what
    dWorldQuickStep does, bit for bit; dContactApprox1 rows (bounds that follow another lambda) report 2."""
    sc = _scene("pile_dense")
    outs = []
    for mode in ("step", "quick"):
        ew = util.engine_world(sc)
        for _ in range(5):
            ew.collide()
            if mode == "step":
                ew.world_step(sc["h"])
            else:
                ew.step(sc["h"])
        st = ew.stats()
        assert st["exact_status"] == (1 if mode == "step" else -1), st
        assert st["solver_iters"] == 20
        outs.append(ew.state())
        ew.close()
    for k in ("pos", "quat", "lvel", "avel"):
        assert np.array_equal(outs[0][k], outs[1][k]), k
    so, se = _surface_case("approx1")
    sc = _scene("c1_low")
    ew = util.engine_world(sc)
    ew.set_surface(se)
    ew.collide()
    ew.world_step(sc["h"])
    assert ew.stats()["exact_status"] == 2 and ew.stats()["solver_iters"] == 20
    ew.close()


def test_boxes_rest_on_a_trimesh_floor():
    """box-trimesh end to end: boxes dropped on a two-triangle floor mesh come to rest on it (as they would on
    dCreatePlane), none tunnels through, and the class counter shows the trimesh kernel took the pairs."""
    sc = scenes._empty_scene("boxes-on-mesh")
    v = np.float32([[-8, 0, -8], [8, 0, -8], [8, 0, 8], [-8, 0, 8]])
    t = np.int32([[0, 2, 1], [0, 3, 2]])
    scenes._add_geom(sc, scenes.TRIMESH, (0.0, 0, 0, 0))
    rs = np.random.RandomState(4)
    sizes = []
    for i in range(12):
        s = rs.uniform(0.4, 1.0, 3)
        sizes.append(s)
        b = scenes._add_body(sc, (float(-5 + 2.5 * (i % 4)), 1.5 + 0.2 * i, float(-3 + 3.0 * (i // 4))), flags=0)
        scenes._add_geom(sc, scenes.BOX, (s[0], s[1], s[2], 0), body=b)
    sc = scenes.finalize(sc)
    sc["meshes"] = [(v, t)]
    ow, ew = util.load_both(sc)
    for step in range(240):
        ew.tick(sc["h"])
    st = ew.stats()
    assert st["flags"] == 0 and st["class_count"][5] == 12          # every box is in touch with the mesh
    s = ew.state()
    half_y = np.array([x[1] for x in sizes]) * 0.5
    assert np.abs(s["lvel"]).max() < 0.05
    assert (s["pos"][:, 1] > half_y - 0.03).all() and (s["pos"][:, 1] < half_y + 0.03).all()
    ew.close()


@pytest.mark.parametrize("seed", [101, 102, 103, 104, 105, 106])
def test_random_soups_pairs_and_contacts(seed):
    """more seeds of the dense random soup (rotated boxes + spheres + plane + static box), both broadphase layouts:
    pair sets equal to the oracle's hash space, contacts equal bit for bit, one tick within the tolerance"""
    sc = scenes.random_soup(260 + 7 * (seed % 5), seed=seed, extent=3.0 + 0.5 * (seed % 3), rotated=(seed % 2 == 0))
    ow, ew0 = util.load_both(sc)
    ew0.close()
    op = util.sorted_pair_set(ow.broadphase(0))
    ref = util.oracle_contacts(ow)
    for mode in (0, 1):
        ew = util.engine_world(sc)
        ew.set_broadphase(mode)
        ew.collide(8)
        assert np.array_equal(util.sorted_pair_set(ew.pairs()), op), mode
        got = _engine_contacts_by_pair(ew)
        assert set(got) == set(ref)
        for key, cs in ref.items():
            pd, nrm, side = got[key]
            assert len(pd) == len(cs), (mode, key)
            for k, c in enumerate(cs):
                assert np.array_equal(pd[k], np.array(list(c.pos) + [c.depth], np.float32)), (mode, key, k)
                assert np.array_equal(nrm[k], np.array(list(c.normal), np.float32)), (mode, key, k)
        if mode == 0:
            ew.step(sc["h"])
            util.oracle_tick_in_engine_order(ow, ew, sc["h"])
            es, os_ = ew.state(), ow.state()
            for k in ("pos", "quat", "lvel", "avel"):
                assert util.rel_err(es[k], os_[k]).max() <= STATE_RTOL, k
        ew.close()


@pytest.mark.parametrize("seed0", [200, 220, 240])
def test_fuzz_scenes_and_surfaces_against_the_oracle(seed0):
    """Twenty random worlds per parameter -- soups with and without plane / static box / rotations, dense piles with walls,
    the server scene with a box or a plane floor, each under the reference surface or one of six others -- three ticks each,
    state against the oracle stepped in the engine's order.  (A sweep of 60 such worlds, 180 ticks, was bit-equal.)"""
    for seed in range(seed0, seed0 + 20):
        rs = np.random.RandomState(seed)
        kind = seed % 3
        if kind == 0:
            sc = scenes.random_soup(int(rs.randint(20, 400)), seed=seed, extent=float(rs.uniform(2.0, 6.0)), rotated=bool(rs.rand() < 0.7),
                                    with_plane=bool(rs.rand() < 0.8), with_static_box=bool(rs.rand() < 0.5))
        elif kind == 1:
            sc = scenes.pile_scene(int(rs.randint(3, 9)), int(rs.randint(3, 9)), int(rs.randint(2, 6)), seed=seed,
                                   spacing=float(rs.uniform(0.5, 1.0)))
        else:
            sc = scenes.server_scene(seed=seed, y_range=(1.0, float(rs.uniform(3.0, 8.0))), floor_plane=bool(rs.rand() < 0.5))
        case = SURFACE_CASES[seed % len(SURFACE_CASES)] if rs.rand() < 0.5 else None
        ow, ew = util.load_both(sc)
        so, rows = None, 3
        if case:
            so, se = _surface_case(case)
            ew.set_surface(se)
            rows = 1 if case == "mu0" else 3
        for step in range(3):
            ew.tick(sc["h"])
            util.oracle_tick_in_engine_order(ow, ew, sc["h"], surf=so, rows_per_contact=rows)
            es, os_ = ew.state(), ow.state()
            for k in ("pos", "quat", "lvel", "avel"):
                assert util.rel_err(es[k], os_[k]).max() <= STATE_RTOL, (seed, kind, case, step, k)
        assert ew.stats()["flags"] == 0
        ew.close()


def _check_island_worlds(sc, nw, per, worlds, ticks, case=None, min_contacts_per_env=0):
    ew = util.engine_world(sc)
    so, rows = None, 3
    if case:
        so, se = _surface_case(case)
        ew.set_surface(se)
        rows = 1 if case == "mu0" else 3
    subs = [(w,) + _world_of_batch(sc, w, per) for w in worlds]
    st = None
    for step in range(ticks):
        ew.tick(sc["h"])
        st = ew.stats()
        assert st["flags"] == 0
        order = ew.solver_order()
        es = ew.state()
        for w, sub, ow, gmap in subs:
            util.oracle_tick_in_engine_order(ow, ew, sc["h"], surf=so, rows_per_contact=rows, geom_map=gmap, order=order)
            os_ = ow.state()
            sl = slice(w * per, (w + 1) * per)
            for k in ("pos", "quat", "lvel", "avel"):
                assert util.rel_err(es[k][sl], os_[k]).max() <= STATE_RTOL, (step, w, k)
    assert st["n_contacts"] >= min_contacts_per_env * nw
    ew.close()
    return st


@pytest.mark.parametrize("seed0", [300, 320])
def test_fuzz_island_path_against_the_oracle(seed0):
    """Twenty batches per parameter: 2-6 worlds of random lattice shape (1 to 160 bodies), spacing 0.4-1.2, the reference
    surface or one of six others; first and last world against the oracle for four ticks on the island path."""
    for seed in range(seed0, seed0 + 20):
        rs = np.random.RandomState(seed)
        nx, ny, nz = int(rs.randint(1, 9)), int(rs.randint(1, 6)), int(rs.randint(1, 5))
        while nx * ny * nz > 160:
            nx -= 1
        nw = int(rs.randint(2, 7))
        spacing = float(rs.uniform(0.4, 1.2))
        sc = scenes.batched_worlds_scene(nw, seed=seed, nx=nx, ny=ny, nz=nz, spacing=spacing)
        case = SURFACE_CASES[seed % len(SURFACE_CASES)] if rs.rand() < 0.5 else None
        _check_island_worlds(sc, nw, nx * ny * nz, sorted({0, nw - 1}), 4, case)


def test_island_path_with_more_units_than_the_shared_memory_cache_holds():
    """48 heavily overlapping bodies per world: far more than four contacts per body slot, so the colouring's unit cache
    lives in global memory instead of the warp's shared-memory region (solver_env.cu); still the oracle's bits."""
    sc = scenes.batched_worlds_scene(3, seed=77, nx=4, ny=4, nz=3, spacing=0.35)
    st = _check_island_worlds(sc, 3, 48, [0, 2], 3, min_contacts_per_env=4 * 64 + 1)
    assert st["env_trips"] > 0


def test_fuzz_exact_dworldstep_against_the_oracles_exact_lcp():
    """The default dWorldStep on 24 random small worlds (soups with rotated boxes, dense piles, server scenes; h = 1/60,
    1/120, 1/240), 25 ticks each, every tick started from the engine's state: velocities within 5e-5 m/s of the oracle's
    exact LCP solution.  (This fuzz found that both sides formed M^-1 J^T in float: a resting box's nearly singular
    contact block amplified the 1e-7 errors of A to 4e-4 .. 4e-3 m/s.  Both now form it in double; worst over 1600
    ticks: 7.6e-6.)"""
    worst, exact = 0.0, 0
    for seed in range(400, 424):
        rs = np.random.RandomState(seed)
        kind = seed % 3
        h = [1 / 60.0, 1 / 120.0, 1 / 240.0][seed % 3]
        if kind == 0:
            sc = scenes.random_soup(int(rs.randint(10, 90)), seed=seed, extent=float(rs.uniform(2.5, 6.0)), rotated=bool(rs.rand() < 0.7))
        elif kind == 1:
            sc = scenes.pile_scene(int(rs.randint(2, 5)), int(rs.randint(2, 5)), int(rs.randint(1, 4)), seed=seed,
                                   spacing=float(rs.uniform(0.7, 1.1)))
        else:
            sc = scenes.server_scene(seed=seed, h=h, y_range=(1.0, float(rs.uniform(2.0, 6.0))), n_dropped=int(rs.randint(8, 64)))
        ow, ew = util.load_both(sc)
        for step in range(25):
            pre = ew.state()
            for i in range(len(pre["pos"])):
                ow.set_body_state(i, pos=pre["pos"][i], q=pre["quat"][i], lvel=pre["lvel"][i], avel=pre["avel"][i])
            ew.collide(8)
            ew.world_step(h)
            st = ew.stats()
            assert st["exact_status"] in (0, 1), (seed, step, st["exact_status"])
            ow.clear_contacts()
            ow.collide_all(8, O.reference_surface())
            if st["exact_status"] == 0:
                exact += 1
                ow.quickstep(h, order_mode=3)
                es, os_ = ew.state(), ow.state()
                d = max(float(np.abs(es[k].astype(np.float64) - os_[k]).max()) for k in ("lvel", "avel"))
                worst = max(worst, d)
                assert d <= 5e-5, (seed, kind, step, d, st["n_islands"], st["max_island_rows"])
            ow.clear_contacts()
        ew.close()
    print("exact dWorldStep fuzz: %d exact ticks, worst |dv| %.2e" % (exact, worst))
    assert exact >= 500


def _fuzz_heightfield(rs, n, scale, amp):
    xs = np.linspace(-0.5, 0.5, n) * scale
    X, Z = np.meshgrid(xs, xs, indexing="ij")
    Y = amp * rs.uniform(-1, 1, size=(n, n))
    verts = np.stack([X, Y, Z], axis=-1).reshape(-1, 3).astype(np.float32)
    tris = []
    for i in range(n - 1):
        for j in range(n - 1):
            a, b, c, d = i * n + j, (i + 1) * n + j, i * n + j + 1, (i + 1) * n + j + 1
            tris += [(a, c, b), (b, c, d)]
    return verts, np.asarray(tris, np.int32)


def _fuzz_triangle_soup(rs, nt, scale):
    c = rs.uniform(-0.5, 0.5, size=(nt, 1, 3)) * scale
    e = rs.normal(size=(nt, 3, 3)) * rs.choice([0.02, 0.1, 0.5], size=(nt, 1, 1)) * scale
    verts = (c + e).reshape(-1, 3).astype(np.float32)
    tris = np.arange(3 * nt, dtype=np.int32).reshape(nt, 3)
    if nt > 3:   # a degenerate triangle and a duplicated one
        verts[3:6] = verts[3]
        tris = np.concatenate([tris, tris[:1]])
    return verts, tris


@pytest.mark.parametrize("seed0", [500, 520, 540])
def test_fuzz_random_meshes_against_the_oracles_brute_force(seed0):
    """Twenty random meshes per parameter -- height fields of 2..40 vertices a side (vertices ON the triangle grid's cell
    borders, flat to bumpy) and triangle soups with degenerate and duplicated triangles, 1 to 400 triangles, at scales 1,
    10, 100 -- with spheres and rotated boxes from a fiftieth to half the mesh's size resting on them: pairs and every
    contact bit-exact against the oracle, which walks all triangles; no flags.  (The fuzz found spheres touching more
    than 64 triangles losing contacts to the candidate list's capacity: the list now keeps the 64 first in order.)"""
    total = 0
    for seed in range(seed0, seed0 + 20):
        rs = np.random.RandomState(seed)
        scale = float(rs.choice([1.0, 10.0, 100.0]))
        if seed % 2 == 0:
            mesh = _fuzz_heightfield(rs, int(rs.randint(2, 40)), scale, 0.05 * scale * rs.rand())
        else:
            mesh = _fuzz_triangle_soup(rs, int(rs.randint(1, 400)), scale)
        nt = len(mesh[1])
        n = int(min(nt, rs.randint(1, 200)))
        sc = scenes.trimesh_contact_scene(n, seed=seed, sphere_scale=float(scale * rs.uniform(0.02, 0.5)), mesh=mesh,
                                          box_fraction=float(rs.choice([0.0, 0.5, 1.0])))
        ow, ew = util.load_both(sc)
        for step in range(2):
            ew.collide(8)
            assert np.array_equal(util.sorted_pair_set(ew.pairs()), util.sorted_pair_set(ow.broadphase(0))), seed
            got = _engine_contacts_by_pair(ew)
            ref = util.oracle_contacts(ow)
            assert set(got) == set(ref), seed
            for key, cs in ref.items():
                pd, nrm, side = got[key]
                assert len(pd) == len(cs), (seed, step, key, len(pd), len(cs))
                for k, c in enumerate(cs):
                    assert np.array_equal(pd[k], np.array(list(c.pos) + [c.depth], np.float32)), (seed, step, key, k)
                    assert np.array_equal(nrm[k], np.array(list(c.normal), np.float32)), (seed, step, key, k)
                    total += 1
            assert ew.stats()["flags"] == 0, seed
            ew.step(sc["h"])
            util.oracle_tick_in_engine_order(ow, ew, sc["h"])
        ew.close()
    assert total > 5000
