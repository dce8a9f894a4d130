"""GPU: oracle parity ON THE BASELINE CONFIGS THEMSELVES (BASELINE.json configs 1, 3, 4 at the sizes and
spacings bench.py runs), not on reduced stand-ins:

* C4 -- the bench scene (8192 worlds x 128 bodies, lattice spacing 1.8) after the bench's 100 settle ticks:
  worlds 0, 4095 and 8191 are pulled out, their state is injected into an oracle world, and pair set,
  contacts (bit-exact) and one tick (relative 1e-4, in the engine's Gauss-Seidel order) are compared;
* C3 -- the 1,048,576-body pile after the bench's 300 settle ticks: a spatial crop of > 2000 bodies is injected
  into the oracle; the engine's pair set restricted to the crop and every contact of those pairs must equal the
  oracle's bit for bit.  (One-tick STATE parity needs a closed system -- a crop's boundary bodies touch bodies
  outside it -- so it is checked on a closed 36,864-body pile of the same construction.)
* C1 -- exactly SURVEY.md section 8(d): the reference's server scene with the reference's spawn heights
  y in [20, 50], at h = 1/60 and at the reference's own h = 1/120 (src/main.c:208), 600 ticks, the oracle stepped
  alongside in the engine's row order; every tick must agree within 1e-4, and BOTH sides must keep the
  constraint residual max |J v+ - c + cfm lambda| (complementarity-aware), the penetration and the energy bounded.
"""
import numpy as np
import pytest

import oracle as O
import util
from odeb200 import scenes
from test_oracle_pins import lcp_residual

pytestmark = pytest.mark.gpu
STATE_RTOL = 1e-4


def _contacts_by_pair(ew):
    pr, cnt, pd, nrm, side = ew.contacts()
    first = np.concatenate([[0], np.cumsum(cnt)[:-1]]) if len(cnt) else np.zeros(0, np.int64)
    return pr, cnt, first, pd, nrm, side


def _sub_scene(sc, static_geoms, body_idx, state):
    """Scene made of the given static geoms + the geoms of the given bodies, with the engine's current state."""
    b, g = sc["bodies"], sc["geoms"]
    body_idx = np.asarray(body_idx)
    bodies = {k: v[body_idx].copy() for k, v in b.items()}
    bodies["pos"] = state["pos"][body_idx].copy(); bodies["quat"] = state["quat"][body_idx].copy()
    bodies["lvel"] = state["lvel"][body_idx].copy(); bodies["avel"] = state["avel"][body_idx].copy()
    bodies["env"] = np.zeros(len(body_idx), np.int32)
    new_body = -np.ones(len(b["pos"]), np.int64)
    new_body[body_idx] = np.arange(len(body_idx))
    dyn = np.nonzero((g["body"] >= 0) & (new_body[np.maximum(g["body"], 0)] >= 0))[0]
    gidx = np.concatenate([np.asarray(static_geoms, np.int64), dyn])
    geoms = {k: v[gidx].copy() for k, v in g.items()}
    geoms["body"] = np.where(geoms["body"] >= 0, new_body[np.maximum(geoms["body"], 0)], -1).astype(np.int32)
    geoms["env"] = np.where(geoms["env"] >= 0, 0, -1).astype(np.int32)
    geom_map = -np.ones(len(g["type"]), np.int64)
    geom_map[gidx] = np.arange(len(gidx))
    sub = scenes.from_arrays("sub", bodies, geoms, meshes=sc.get("meshes", []), gravity=sc["gravity"], h=sc["h"])
    return sub, gidx, geom_map


def _compare_pairs_and_contacts(ow, sub, gidx, geom_map, pr, cnt, first, pd, nrm):
    """engine pairs with both geoms in the sub-world == oracle hash-space pairs; contacts bit-exact."""
    inside = (geom_map[pr[:, 0]] >= 0) & (geom_map[pr[:, 1]] >= 0)
    sel = np.nonzero(inside)[0]
    eng = util.sorted_pair_set(np.stack([geom_map[pr[sel, 0]], geom_map[pr[sel, 1]]], axis=1))
    orc = util.sorted_pair_set(ow.broadphase(0))
    assert np.array_equal(eng, orc), "pair sets differ: engine %d, oracle %d" % (len(eng), len(orc))
    types = sub["geoms"]["type"]
    n_contacts = 0
    for i in sel:
        g1, g2 = int(geom_map[pr[i, 0]]), int(geom_map[pr[i, 1]])
        assert types[g1] <= types[g2]
        ref = ow.collide(g1, g2, 8)
        assert len(ref) == cnt[i], (g1, g2, len(ref), cnt[i])                 # contact counts: exact
        for k, c in enumerate(ref):
            assert np.array_equal(pd[first[i] + k], np.array(list(c.pos) + [c.depth], np.float32)), (g1, g2, k)
            assert np.array_equal(nrm[first[i] + k], np.array(list(c.normal), np.float32)), (g1, g2, k)
        n_contacts += len(ref)
    return len(sel), n_contacts


def _oracle_of(sub):
    ow = O.OracleWorld(gravity=sub["gravity"])
    ow.load_scene(sub)
    b = sub["bodies"]
    for i in range(len(b["pos"])):     # the engine's exact state (add_body re-normalises the quaternion like dBodySetQuaternion)
        ow.set_body_state(i, pos=b["pos"][i], q=b["quat"][i], lvel=b["lvel"][i], avel=b["avel"][i])
    ow._types = [int(t) for t in sub["geoms"]["type"]]
    ow._bodies = [int(b) for b in sub["geoms"]["body"]]
    return ow


def test_c4_bench_scene_worlds_against_oracle():
    nw, per = 8192, 128
    sc = scenes.batched_worlds_scene(nw, seed=4)          # the bench scene: spacing 1.8
    ew = util.engine_world(sc)
    h = sc["h"]
    for _ in range(100):                                   # bench.py's settle ticks
        ew.tick(h)
    ew.collide(8)
    pre = ew.state()
    pr, cnt, first, pd, nrm, side = _contacts_by_pair(ew)
    ew.step(h)
    post = ew.state()
    order = ew.solver_order()
    st = ew.stats()
    assert st["flags"] == 0 and st["n_overflow"] == 0 and st["n_contacts"] > 1000000
    ew.close()
    for w in (0, 4095, 8191):
        bidx = np.arange(w * per, (w + 1) * per)
        sub, gidx, gmap = _sub_scene(sc, [0], bidx, pre)
        ow = _oracle_of(sub)
        npairs, ncont = _compare_pairs_and_contacts(ow, sub, gidx, gmap, pr, cnt, first, pd, nrm)
        assert npairs > 100 and ncont > 100, (w, npairs, ncont)
        util.oracle_tick_in_engine_order(ow, None, h, geom_map=gmap, order=order)
        os_ = ow.state()
        for k in ("pos", "quat", "lvel", "avel"):
            err = util.rel_err(post[k][bidx], os_[k]).max()
            assert err <= STATE_RTOL, (w, k, err)
        ow.close()


def test_c3_full_size_pile_crop_against_oracle():
    sc = scenes.pile_scene(256, 256, 16, seed=3)           # the bench scene, 1,048,576 bodies
    ew = util.engine_world(sc)
    h = sc["h"]
    for _ in range(300):                                   # bench.py's settle ticks
        ew.tick(h)
    ew.collide(8)
    pre = ew.state()
    pr, cnt, first, pd, nrm, side = _contacts_by_pair(ew)
    st = ew.stats()
    assert st["flags"] == 0 and st["n_pairs"] > 1000000
    ew.close()
    # crop 1: the middle of the pile; crop 2: a corner, where the wall planes carry contacts
    lim = 0.5 * 256 * 1.8
    for name, mask in (("middle", (np.abs(pre["pos"][:, 0]) < 11.0) & (np.abs(pre["pos"][:, 2]) < 11.0)),
                       ("corner", (pre["pos"][:, 0] > lim - 16.0) & (pre["pos"][:, 2] > lim - 16.0))):
        bidx = np.nonzero(mask)[0]
        assert len(bidx) >= 1200, (name, len(bidx))
        sub, gidx, gmap = _sub_scene(sc, [0, 1, 2, 3, 4], bidx, pre)
        ow = _oracle_of(sub)
        npairs, ncont = _compare_pairs_and_contacts(ow, sub, gidx, gmap, pr, cnt, first, pd, nrm)
        print("C3 crop %s: %d bodies, %d pairs, %d contacts bit-exact" % (name, len(bidx), npairs, ncont))
        assert npairs > 2 * len(bidx) and ncont > len(bidx)
        ow.close()
    assert ((np.abs(pre["pos"][:, 0]) < 11.0) & (np.abs(pre["pos"][:, 2]) < 11.0)).sum() >= 2000


def test_closed_pile_one_tick_state_against_oracle():
    sc = scenes.pile_scene(48, 48, 16, seed=3)             # same construction as C3, 36,864 bodies, closed by its walls
    ow, ew = util.load_both(sc)
    h = sc["h"]
    for _ in range(300):
        ew.tick(h)
    ew.collide(8)
    pre = ew.state()
    for i in range(len(pre["pos"])):
        ow.set_body_state(i, pos=pre["pos"][i], q=pre["quat"][i], lvel=pre["lvel"][i], avel=pre["avel"][i])
    assert np.array_equal(util.sorted_pair_set(ew.pairs()), util.sorted_pair_set(ow.broadphase(0)))
    ew.step(h)
    nc = util.oracle_tick_in_engine_order(ow, ew, h)
    st = ew.stats()
    assert nc == st["n_contacts"] and nc > 40000 and st["flags"] == 0
    es, os_ = ew.state(), ow.state()
    for k in ("pos", "quat", "lvel", "avel"):
        err = util.rel_err(es[k], os_[k]).max()
        assert err <= STATE_RTOL, (k, err)
    ew.close()


def _energy(s, nd):
    return float(0.5 * (s["lvel"][:nd].astype(np.float64) ** 2).sum() + 0.5 * (s["avel"][:nd].astype(np.float64) ** 2).sum() +
                 9.8 * s["pos"][:nd, 1].astype(np.float64).sum())      # m = 1, I = identity (dBodyCreate defaults)


@pytest.mark.parametrize("h", [1.0 / 60.0, 1.0 / 120.0])
def test_c1_600_ticks_alongside_the_oracle(h):
    """static map verbatim, 64 bodies spawned at y in [20, 50] (src/main.c:504-521), 4 kinematic
    spheres; h = 1/60 (BASELINE config 1) and 1/120 (what the reference steps at, src/main.c:208)."""
    sc = scenes.server_scene(seed=1, h=h)
    ow, ew = util.load_both(sc)
    ow.keep_rows()
    nd = 64
    dyn_geom = np.array([(b >= 0 and b < nd) for b in sc["geoms"]["body"]])
    E0 = _energy(ew.state(), nd)
    Ee = Eo = E0
    inc_e = inc_o = 0.0
    res_e = res_o = pen = 0.0
    late_e = 0.0
    biteq = 0
    vmax = np.sqrt(2 * 9.8 * 50.0) + 0.5                       # free-fall speed from the highest spawn
    for step in range(600):
        ew.collide(8)
        pr, cnt, first, pd, nrm, side = _contacts_by_pair(ew)
        if len(pr):
            act = np.repeat(dyn_geom[pr[:, 0]] | dyn_geom[pr[:, 1]], cnt)
            if act.any():
                pen = max(pen, float(pd[:len(act)][act, 3].max()))
        ew.step(h)
        util.oracle_tick_in_engine_order(ow, ew, h)            # asserts the oracle finds the same pairs and contacts
        es, os_ = ew.state(), ow.state()
        for k in ("pos", "quat", "lvel", "avel"):
            err = util.rel_err(es[k], os_[k]).max()
            assert err <= STATE_RTOL, (step, k, err)
        same = all(np.array_equal(es[k], os_[k]) for k in ("pos", "quat", "lvel", "avel"))
        biteq += same
        rows = ow.last_rows() if ow.num_rows else None
        if rows is not None and len(rows["c"]):
            ve = np.concatenate([es["lvel"], es["avel"]], axis=1).astype(np.float64)
            vo = np.concatenate([os_["lvel"], os_["avel"]], axis=1).astype(np.float64)
            r_e = float(lcp_residual(rows, ve, h)[0].max())
            r_o = float(lcp_residual(rows, vo, h)[0].max())
            res_e, res_o = max(res_e, r_e), max(res_o, r_o)
            if step >= 450:
                late_e = max(late_e, r_e)
        e1, e2 = _energy(es, nd), _energy(os_, nd)
        inc_e, inc_o = max(inc_e, e1 - Ee), max(inc_o, e2 - Eo)
        Ee, Eo = e1, e2
        if not same:                                            # keep stepping from identical states
            for i in range(len(es["pos"])):
                ow.set_body_state(i, pos=es["pos"][i], q=es["quat"][i], lvel=es["lvel"][i], avel=es["avel"][i])
    print("C1 h=1/%d: bit-equal ticks %d/600, max residual engine %.4f oracle %.4f (late %.4f), max penetration %.4f, "
          "energy %.1f -> %.1f, max one-tick energy increase %.2e" % (round(1 / h), biteq, res_e, res_o, late_e, pen, E0, Ee, inc_e))
    # constraint residual of 20 sweeps: bounded by the impact speeds, and the engine is no worse than the oracle
    assert res_o < 0.5 * vmax and res_e <= 1.001 * res_o + 1e-4
    assert late_e < 1.5                                         # once the pile has formed
    # penetration: never more than one tick of travel at the top speed, and small at the end
    assert pen <= vmax * h + 0.02
    # energy: bounce 0.2 and the sweeps dissipate; ERP push-out may add a little per tick
    assert inc_e <= 1e-3 * E0 and inc_o <= 1e-3 * E0 and Ee < 0.1 * E0 and abs(Ee - Eo) <= 1e-4 * E0
    s = ew.state()
    assert s["pos"][:nd, 1].min() > 0.5 and np.array_equal(s["pos"][nd:, 1], np.full(4, 2.0, np.float32))
    assert np.allclose(np.linalg.norm(s["quat"], axis=1), 1.0, atol=1e-5)
    ew.close()
