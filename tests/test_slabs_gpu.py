"""GPU (one device): two slabs of a pile emulated as two worlds that exchange boundary-body states through
device buffers each tick -- the same code path bench.py runs over NCCL with one slab per GPU."""
import numpy as np
import pytest

import util
from odeb200 import scenes, slabs

pytestmark = pytest.mark.gpu


def _build(n_slabs, **kw):
    import odeb200
    out = []
    for r in range(n_slabs):
        sc, halo = slabs.slab_scene(r, n_slabs, **kw)
        w = odeb200.World(gravity=sc["gravity"])
        w.load_scene(sc)
        out.append((sc, slabs.SlabWorld(w, halo, "cuda:0")))
    return out


def test_two_slab_pile_with_halo_exchange():
    kw = dict(nx_per_slab=8, nz=32, ny=6, seed=5, spacing=0.8, margin_cols=3, coupling="ghost")
    built = _build(2, **kw)
    sl = [s for _, s in built]
    h = built[0][0]["h"]
    for step in range(150):
        for s in sl:
            s.pack("state")
        slabs.exchange_local(sl, "state")
        for s in sl:
            s.unpack("state")
        if step == 149:
            break
        for s in sl:
            s.w.tick(h)
    st = [s.w.state() for s in sl]
    n0, n1 = built[0][0]["n_owned"], built[1][0]["n_owned"]
    # after an exchange the ghosts carry exactly the owners' states
    send0, recv0 = sl[0].sides["right"]["send_state_idx"].cpu().numpy(), sl[0].sides["right"]["recv_state_idx"].cpu().numpy()
    send1, recv1 = sl[1].sides["left"]["send_state_idx"].cpu().numpy(), sl[1].sides["left"]["recv_state_idx"].cpu().numpy()
    for k in ("pos", "quat", "lvel", "avel"):
        assert np.array_equal(st[0][k][send0], st[1][k][recv1]), k
        assert np.array_equal(st[1][k][send1], st[0][k][recv0]), k
    # the pile settled: finite, above the ground, slow, and bodies pressed against the interface did not
    # tunnel through their ghost neighbours (no deep penetration between owned bodies and ghosts)
    for r, s in enumerate(sl):
        own = slice(0, (n0, n1)[r])
        assert np.isfinite(st[r]["pos"]).all()
        assert st[r]["pos"][own, 1].min() > 0.0
        assert np.abs(st[r]["lvel"][own]).max() < 9.0     # still settling after 150 ticks, but nothing was launched
        s.w.collide(8)
        pr, cnt, pd, nrm, side = s.w.contacts()
        assert s.w.stats()["flags"] == 0
        gb = built[r][0]["geoms"]["body"]
        is_ghost = gb >= (n0, n1)[r]
        cross = np.repeat(is_ghost[pr[:, 0]] | is_ghost[pr[:, 1]], cnt)
        assert cross.sum() >= 4                      # the interface is active ...
        assert pd[cross, 3].max() < 0.15             # ... and contacts across it stay shallow
    # same pile in one world (no decomposition): bulk statistics agree
    b = scenes._concat([built[0][0]["bodies"], built[1][0]["bodies"]])
    keep0 = np.arange(n0)
    keep1 = len(built[0][0]["bodies"]["pos"]) + np.arange(n1)
    keep = np.concatenate([keep0, keep1])
    bodies = {k: v[keep] for k, v in b.items()}
    g0, g1 = built[0][0]["geoms"], built[1][0]["geoms"]
    geoms = scenes._concat([{k: v[:5] for k, v in g0.items()}, {k: v[5:5 + n0] for k, v in g0.items()},
                            {k: v[5:5 + n1] for k, v in g1.items()}])
    geoms["body"] = np.concatenate([np.full(5, -1, np.int32), np.arange(n0 + n1, dtype=np.int32)])
    one = util.engine_world(scenes.from_arrays("merged", bodies, geoms, h=h))
    for _ in range(149):
        one.tick(h)
    so = one.state()
    y_dec = np.concatenate([st[0]["pos"][:n0, 1], st[1]["pos"][:n1, 1]])
    assert abs(y_dec.mean() - so["pos"][:, 1].mean()) < 0.05 * so["pos"][:, 1].mean()
    assert abs(y_dec.max() - so["pos"][:, 1].max()) < 0.35
    one.close()
    for s in sl:
        s.w.close()


def _two_sphere_slabs(coupling, va, vb):
    """Two unit-mass spheres on a collision course along x in free space, one per slab (face at x = 0)."""
    import odeb200
    CAT_OBJ, CAT_GHOST = slabs.CAT_OBJ, slabs.CAT_GHOST
    pa, pb = (-0.45, 0.0, 0.0), (0.45, 0.0, 0.0)
    worlds, halos = [], []
    for r in range(2):
        sc = scenes._empty_scene("pair%d" % r, gravity=(0.0, 0.0, 0.0))
        own_p, own_v = (pa, va) if r == 0 else (pb, vb)
        oth_p, oth_v = (pb, vb) if r == 0 else (pa, va)
        b0 = scenes._add_body(sc, own_p, lvel=own_v, flags=0)
        scenes._add_geom(sc, scenes.SPHERE, (0.5, 0, 0, 0), body=b0, cat=CAT_OBJ, col=CAT_OBJ | CAT_GHOST)
        hs = {"send_state": None, "recv_state": None, "send_imp": None, "recv_imp": None}
        mirror = coupling == "ghost" or r == 0
        if mirror:
            b1 = scenes._add_body(sc, oth_p, lvel=oth_v, flags=scenes.BODY_KINEMATIC if coupling == "ghost" else 0)
            scenes._add_geom(sc, scenes.SPHERE, (0.5, 0, 0, 0), body=b1, cat=CAT_GHOST, col=0)
            hs["recv_state"] = np.int32([b1])
            if coupling == "impulse":
                hs["send_imp"] = np.int32([b1])
        if coupling == "ghost" or r == 1:
            hs["send_state"] = np.int32([b0])
        if coupling == "impulse" and r == 1:
            hs["recv_imp"] = np.int32([b0])
        sc = scenes.finalize(sc)
        w = odeb200.World(gravity=sc["gravity"])
        w.load_scene(sc)
        halo = {"left": hs if r == 1 else None, "right": hs if r == 0 else None}
        worlds.append(slabs.SlabWorld(w, halo, "cuda:0"))
    return worlds


def _run_pair(coupling, va, vb, ticks=12):
    sl = _two_sphere_slabs(coupling, va, vb)
    exch = lambda kind: slabs.exchange_local(sl, kind)   # noqa: E731
    for _ in range(ticks):
        # the same order bench.py runs per rank: states, tick, impulses -- here phase by phase over both slabs
        for s in sl:
            s.pack("state")
        exch("state")
        for s in sl:
            s.unpack("state")
        for s in sl:
            s.w.tick(1.0 / 60.0)
        if sl[0].has_imp:
            for s in sl:
                s.pack("imp")
            exch("imp")
            for s in sl:
                s.unpack("imp")
    v = [s.w.state()["lvel"][0].astype(np.float64) for s in sl]
    for s in sl:
        s.w.close()
    return v


def test_impulse_coupling_conserves_momentum_across_the_face():
    """SURVEY.md section 8e step 3.  A moving sphere (slab 0) hits a resting one (slab 1).  With the impulse halo
    the lower slab solves the two-body contact and the owner of the struck sphere receives the opposite impulse:
    total momentum is conserved and the result matches the same two spheres in ONE world.  With kinematic ghosts
    each side sees an immovable obstacle and over-corrects."""
    va, vb = (2.0, 0.0, 0.0), (0.0, 0.0, 0.0)
    v_imp = _run_pair("impulse", va, vb)
    p_imp = v_imp[0][0] + v_imp[1][0]
    assert abs(p_imp - 2.0) < 1e-4, v_imp
    assert v_imp[1][0] > 0.5 and v_imp[0][0] < 1.5            # the struck sphere really moves off
    # one world, no decomposition
    sc = scenes._empty_scene("pair", gravity=(0.0, 0.0, 0.0))
    for p, v in (((-0.45, 0, 0), va), ((0.45, 0, 0), vb)):
        b = scenes._add_body(sc, p, lvel=v, flags=0)
        scenes._add_geom(sc, scenes.SPHERE, (0.5, 0, 0, 0), body=b)
    one = util.engine_world(scenes.finalize(sc))
    for _ in range(12):
        one.tick(1.0 / 60.0)
    vo = one.state()["lvel"].astype(np.float64)
    one.close()
    assert abs(vo[0, 0] + vo[1, 0] - 2.0) < 1e-4
    print("impulse halo:", v_imp, "one world:", vo[:, 0])
    assert abs(v_imp[0][0] - vo[0, 0]) < 0.25 and abs(v_imp[1][0] - vo[1, 0]) < 0.25, (v_imp, vo)
    v_gh = _run_pair("ghost", va, vb)
    # each side resolves the whole relative velocity against an immovable obstacle: the pair flies apart much
    # faster than physics allows (the impulse halo does not)
    sep_one = vo[1, 0] - vo[0, 0]
    assert (v_gh[1][0] - v_gh[0][0]) > 2.0 * sep_one and abs((v_imp[1][0] - v_imp[0][0]) - sep_one) < 0.3, (v_gh, v_imp, vo)


def test_two_slab_pile_with_impulse_halo():
    """the default coupling on a pile: states down, impulses up; the pile settles like the undecomposed one"""
    kw = dict(nx_per_slab=8, nz=32, ny=6, seed=5, spacing=0.8, margin_cols=3, coupling="impulse")
    built = _build(2, **kw)
    sl = [s for _, s in built]
    h = built[0][0]["h"]
    exch = lambda kind: slabs.exchange_local(sl, kind)   # noqa: E731
    for step in range(150):
        for kind in ("state",):
            for s in sl:
                s.pack(kind)
            exch(kind)
            for s in sl:
                s.unpack(kind)
        for s in sl:
            s.w.tick(h)
        for s in sl:
            s.pack("imp")
        exch("imp")
        for s in sl:
            s.unpack("imp")
    st = [s.w.state() for s in sl]
    n0, n1 = built[0][0]["n_owned"], built[1][0]["n_owned"]
    assert "left" not in sl[0].sides and "right" not in sl[1].sides
    assert len(st[1]["pos"]) == n1                              # the upper slab holds no ghosts
    for r, n in ((0, n0), (1, n1)):
        assert np.isfinite(st[r]["pos"]).all()
        assert st[r]["pos"][:n, 1].min() > 0.0
        assert np.abs(st[r]["lvel"][:n]).max() < 9.0
        assert sl[r].w.stats()["flags"] == 0
    # cross-face contacts exist on the lower slab and stay shallow
    sl[0].w.collide(8)
    pr, cnt, pd, nrm, side = sl[0].w.contacts()
    gb = built[0][0]["geoms"]["body"]
    cross = np.repeat((gb[pr[:, 0]] >= n0) | (gb[pr[:, 1]] >= n0), cnt)
    assert cross.sum() >= 4 and pd[cross, 3].max() < 0.15
    y_dec = np.concatenate([st[0]["pos"][:n0, 1], st[1]["pos"][:n1, 1]])
    assert 0.2 < y_dec.mean() < 3.0
    for s in sl:
        s.w.close()


def _build_dynamic(n_slabs, **kw):
    import odeb200
    out = []
    for r in range(n_slabs):
        sc, info = slabs.dynamic_slab_scene(r, n_slabs, **kw)
        w = odeb200.World(gravity=sc["gravity"])
        w.load_scene(sc)
        out.append((sc, slabs.DynamicSlabWorld(w, info, "cuda:0")))
    return out


def _phase(sl, kind):
    for s in sl:
        s.pack(kind)
    slabs.exchange_local(sl, kind)
    for s in sl:
        s.unpack(kind)


def test_dynamic_halo_selects_by_position_and_migrates_bodies():
    """Dynamic halo: the boundary set is re-selected on the device from the bodies' current x each tick and sent
    as whole bodies into the lower slab's ghost pool; bodies whose centre crossed a face change owner (destroyed on
    one side, spawned through the handle API on the other).  A fast sphere is shot from slab 0 into slab 1."""
    kw = dict(nx_per_slab=6, nz=8, ny=3, seed=5, spacing=1.0, margin_cols=2, mig_cap=256)
    built = _build_dynamic(2, **kw)
    sl = [s for _, s in built]
    h = built[0][0]["h"]
    n0, n1 = sl[0].n_owned, sl[1].n_owned
    # the projectile: body 0 of slab 0, lifted above the pile and thrown along +x
    L = sl[0].w.L
    shot = sl[0].w.body_handle(0)
    L.dBodySetPosition(shot, -1.0, 6.0, 0.0)
    L.dBodySetLinearVel(shot, 6.0, 0.0, 0.0)
    owner_history = []
    for step in range(60):
        if step % 4 == 0 and step > 0:
            _phase(sl, "mig")
        _phase(sl, "state")
        if step == 1:
            # ghosts on slab 0 are exactly slab 1's current boundary bodies, whole (state, mass, shape)
            st0, st1 = sl[0].w.state(), sl[1].w.state()
            idx = sl[1].sides["left"]["send_state_idx"].cpu().numpy()
            k = int((idx >= 0).sum())
            assert 0 < k < len(idx) and (idx[k:] == -1).all()
            sel = np.nonzero(st1["pos"][:n1, 0] < sl[1].info["face_left"] + sl[1].info["margin"])[0]
            assert np.array_equal(idx[:k], sel)                       # selection = positions, ascending
            gb = sl[0].sides["right"]["ghost_body"].cpu().numpy()[:k]
            for key in ("pos", "quat", "lvel", "avel"):
                assert np.array_equal(st0[key][gb], st1[key][idx[:k]]), key
        for s in sl:
            s.w.tick(h)
        _phase(sl, "imp")
        owner_history.append((sl[0].n_owned, sl[1].n_owned))
    assert owner_history[0] == (n0, n1)
    assert sl[0].migrated_out >= 1 and sl[1].migrated_in == sl[0].migrated_out   # the shot (at least) changed owner
    assert sl[0].n_owned + sl[1].n_owned == n0 + n1                               # nobody lost, nobody duplicated
    st0, st1 = sl[0].w.state(), sl[1].w.state()
    m0 = sl[0].own_mask.cpu().numpy()[:len(st0["pos"])].astype(bool)
    m1 = sl[1].own_mask.cpu().numpy()[:len(st1["pos"])].astype(bool)
    assert m0.sum() == sl[0].n_owned and m1.sum() == sl[1].n_owned and not m0[0]  # body 0 of slab 0 is gone ...
    xs1 = st1["pos"][m1]
    assert (xs1[:, 0] > sl[1].info["face_left"] + 2.0).any()                       # ... and flies on inside slab 1
    for st, m in ((st0, m0), (st1, m1)):
        assert np.isfinite(st["pos"][m]).all() and st["pos"][m, 1].min() > 0.0
    for s in sl:
        assert s.w.stats()["flags"] == 0
        s.w.close()


def test_c_slab_driver_equals_the_python_orchestration_bit_for_bit():
    """The product path of config 5 is the C driver inside libode_b200.so (dSlabCreateB200 / dSlabTickLocalB200 /
    dSlabMigrateLocalB200: buffers, transfers and event ordering in the library, no host synchronisation per tick).  Three
    slabs on one GPU, a projectile crossing two faces: after 48 ticks with a migration every 4 the owned bodies' states
    must equal, bit for bit, those of the Python-orchestrated run above (same kernels, same order of operations)."""
    import odeb200
    kw = dict(nx_per_slab=6, nz=8, ny=3, seed=5, spacing=1.0, margin_cols=2, mig_cap=256)

    def shoot(worlds):
        L = worlds[0].L
        shot = worlds[0].body_handle(0)
        L.dBodySetPosition(shot, -4.0, 6.0, 0.0)
        L.dBodySetLinearVel(shot, 12.0, 0.0, 0.0)

    # reference: Python orchestration
    built = _build_dynamic(3, **kw)
    sl = [s for _, s in built]
    h = built[0][0]["h"]
    shoot([s.w for s in sl])
    for step in range(48):
        if step % 4 == 0 and step > 0:
            _phase(sl, "mig")
        _phase(sl, "state")
        for s in sl:
            s.w.tick(h)
        _phase(sl, "imp")
    ref = [(s.w.state(), s.own_mask.cpu().numpy(), s.n_owned, s.migrated_in, s.migrated_out) for s in sl]
    for s in sl:
        s.w.close()
    assert ref[0][4] >= 1 and ref[2][3] + ref[1][3] >= 1
    # product: the C driver
    cs = []
    for r in range(3):
        sc, info = slabs.dynamic_slab_scene(r, 3, **kw)
        w = odeb200.World(gravity=sc["gravity"])
        w.load_scene(sc)
        cs.append(slabs.CSlab(w, info))
    for a, b in zip(cs[:-1], cs[1:]):
        slabs.CSlab.connect(a, b)
    shoot([c.w for c in cs])
    for step in range(48):
        if step % 4 == 0 and step > 0:
            slabs.CSlab.migrate_local(cs)
        slabs.CSlab.tick_local(cs, h)
    for r, c in enumerate(cs):
        st = c.w.state()
        info = c.get_info()
        rst, rmask, rown, rin, rout = ref[r]
        assert info["n_owned"] == rown and info["migrated_in"] == rin and info["migrated_out"] == rout, (r, info)
        assert info["halo_overflow"] == 0 and info["mig_overflow"] == 0 and info["ticks"] == 48
        n = len(st["pos"])
        assert n == len(rst["pos"])
        own = rmask[:n].astype(bool)
        for k in ("pos", "quat", "lvel", "avel"):
            assert np.array_equal(st[k][own], rst[k][own]), (r, k)
        assert c.w.stats()["flags"] == 0
        c.close()
        c.w.close()


def test_union_of_slab_pair_sets_equals_the_single_world_pair_set():
    """The broadphase depends on the state only, so the decomposition can be checked EXACTLY: after the state halo, the
    pairs slab 0 finds (own-own, own-ghost, own-static) united with the pairs slab 1 finds, mapped to global body
    numbers, must be precisely the pair set of one undecomposed world holding the same bodies in the same states --
    nothing lost at the cut, nothing found twice -- and the contacts of every cross-face pair must be bit-identical."""
    import odeb200
    kw = dict(nx_per_slab=6, nz=8, ny=3, seed=5, spacing=0.6, margin_cols=3, mig_cap=256)   # dense: neighbours overlap
    built = _build_dynamic(2, **kw)
    sl = [s for _, s in built]
    scs = [sc for sc, _ in built]
    h = scs[0]["h"]
    for step in range(25):                       # let the pile collapse across the face (no migration: ids stay put)
        _phase(sl, "state")
        for s in sl:
            s.w.tick(h)
        _phase(sl, "imp")
    _phase(sl, "state")                          # ghosts = the upper slab's current boundary bodies
    n_own = [s.info["n_own"] for s in sl]
    n_static = sl[0].info["n_static"]
    states = [s.w.state() for s in sl]
    # ghost slot k of slab 0 mirrors body send_idx[k] of slab 1
    send_idx = sl[1].sides["left"]["send_state_idx"].cpu().numpy()
    ghost_body = sl[0].sides["right"]["ghost_body"].cpu().numpy()
    n_sel = int((send_idx >= 0).sum())
    assert n_sel > 10

    def gid(r, b):                               # global body number in the merged world
        return b if r == 0 else n_own[0] + b

    ghost_to_global = {int(ghost_body[k]): gid(1, int(send_idx[k])) for k in range(n_sel)}
    pair_sets, cross_contacts = [], {}
    for r, s in enumerate(sl):
        s.w.collide(8)
        pr, cnt, pd, nrm, side = s.w.contacts()
        first = np.concatenate([[0], np.cumsum(cnt)[:-1]])
        gb = scs[r]["geoms"]["body"]
        out = set()
        for i, (a, b) in enumerate(pr.tolist()):
            ends = []
            for g in (a, b):
                body = int(gb[g])
                if body < 0:
                    ends.append(("static", g))                       # the five planes: same geom numbers everywhere
                elif body < n_own[r]:
                    ends.append(("body", gid(r, body)))
                else:
                    assert r == 0 and body in ghost_to_global, "a pair with an unused ghost slot"
                    ends.append(("body", ghost_to_global[body]))
            if ends[0][0] == "static" and ends[1][0] == "static":
                continue                                              # plane-plane pairs exist in every world
            key = tuple(sorted(ends))
            assert key not in out
            out.add(key)
            if r == 0 and any(int(gb[g]) >= n_own[0] for g in (a, b)):
                cross_contacts[key] = (pd[first[i]:first[i] + cnt[i]].copy(), nrm[first[i]:first[i] + cnt[i]].copy())
        pair_sets.append(out)
    assert not (pair_sets[0] & pair_sets[1]), "a pair was found by both slabs"
    assert len(cross_contacts) >= 3, "the test needs contacts across the face"
    # the undecomposed world: the planes + both slabs' own bodies in their current states
    bodies = {k: np.concatenate([scs[r]["bodies"][k][:n_own[r]] for r in range(2)]) for k in scs[0]["bodies"]}
    for k in ("pos", "quat", "lvel", "avel"):
        bodies[k] = np.concatenate([states[r][k][:n_own[r]] for r in range(2)])
    geoms = {k: np.concatenate([scs[0]["geoms"][k][:n_static]] + [scs[r]["geoms"][k][n_static:n_static + n_own[r]] for r in range(2)])
             for k in scs[0]["geoms"]}
    geoms["body"] = np.concatenate([np.full(n_static, -1), np.arange(n_own[0] + n_own[1])]).astype(np.int32)
    geoms["cat"][n_static:] = slabs.CAT_OBJ
    geoms["col"][n_static:] = slabs.CAT_OBJ | slabs.CAT_MAP
    one = odeb200.World(gravity=scs[0]["gravity"])
    one.load_scene(scenes.from_arrays("one", bodies, geoms, h=h))
    # bit-exact states: the bulk loader re-normalises quaternions (dBodySetQuaternion semantics), the state scatter does not
    import torch
    nall = n_own[0] + n_own[1]
    rec = np.zeros((nall, 16), np.float32)
    rec[:, 0:3] = bodies["pos"]; rec[:, 4:8] = bodies["quat"]; rec[:, 8:11] = bodies["lvel"]; rec[:, 12:15] = bodies["avel"]
    d_rec = torch.as_tensor(rec, device="cuda:0")
    d_idx = torch.arange(nall, dtype=torch.int32, device="cuda:0")
    one.collide(8)                               # (first sync of the world to the device)
    one.L.dWorldUnpackStatesDeviceB200(one.w, d_idx.data_ptr(), nall, d_rec.data_ptr())
    one.collide(8)
    st1 = one.state()
    for k in ("pos", "quat"):
        assert np.array_equal(st1[k], bodies[k])
    pr, cnt, pd, nrm, side = one.contacts()
    first = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    ref, ref_contacts = set(), {}
    for i, (a, b) in enumerate(pr.tolist()):
        ends = [("static", g) if g < n_static else ("body", g - n_static) for g in (a, b)]
        if ends[0][0] == "static" and ends[1][0] == "static":
            continue
        key = tuple(sorted(ends))
        ref.add(key)
        ref_contacts[key] = (pd[first[i]:first[i] + cnt[i]], nrm[first[i]:first[i] + cnt[i]])
    union = pair_sets[0] | pair_sets[1]
    assert union == ref, (len(union), len(ref), sorted(union - ref)[:5], sorted(ref - union)[:5])
    n_with_contacts = 0
    for key, (cpd, cn) in cross_contacts.items():
        rpd, rn = ref_contacts[key]
        assert len(cpd) == len(rpd), key
        if len(cpd):
            n_with_contacts += 1
            # the slab lists the ghost after its own bodies, the single world by global number: the same two geoms may be
            # handed to the collider in the other order, which flips the normal's sign but nothing else
            same = np.array_equal(cpd, rpd) and np.array_equal(cn, rn)
            flipped = np.allclose(np.sort(cpd[:, 3]), np.sort(rpd[:, 3]), atol=1e-5) and np.allclose(np.abs(cn), np.abs(rn), atol=1e-5)
            assert same or flipped, key
    assert n_with_contacts >= 1
    one.close()
    for s in sl:
        s.w.close()


def test_c_host_drives_the_slab_decomposition():
    """rl-ode-physics_b200/host/slab_server.c: a plain-C host builds two slabs of a lattice pile through the ODE handle
    API and runs them through the C slab driver (local transport: one GPU): every body stays owned by exactly one slab,
    nothing overflows, bodies cross the face in both directions of the bookkeeping (migrated out == migrated in), the
    halo carries bodies, and the pile comes to rest on the plane."""
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "rl-ode-physics_b200", "host", "slab_server")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(exe)])
    cols, nz, ny, ticks = 12, 16, 6, 320
    r = subprocess.run([exe, "local", "2", str(cols), str(nz), str(ny), str(ticks)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rows = []
    for line in r.stdout.splitlines():
        m = re.match(r"slab (\d+): (.*)", line)
        if m:
            toks = m.group(2).split()
            d, i = {}, 0
            while i < len(toks):
                k = toks[i]
                n = 3 if k == "momentum" else 1
                d[k] = [float(x) for x in toks[i + 1:i + 1 + n]]
                i += 1 + n
            rows.append(d)
    assert len(rows) == 2, r.stdout
    assert sum(d["owned"][0] for d in rows) == 2 * cols * nz * ny
    assert sum(d["migrated_in"][0] for d in rows) == sum(d["migrated_out"][0] for d in rows) > 0   # the kicked columns
    for d in rows:
        assert d["halo_overflow"][0] == 0 and d["mig_overflow"][0] == 0
        assert d["ymin"][0] > -0.05 and d["vmax"][0] < 6.0, d
    assert rows[1]["halo_selected"][0] > 0 and rows[1]["halo_bytes_per_tick"][0] > 0
