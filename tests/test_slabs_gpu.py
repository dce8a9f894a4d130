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
    kw = dict(nx_per_slab=8, nz=32, ny=6, seed=5, spacing=0.8, margin_cols=3)
    built = _build(2, **kw)
    sl = [s for _, s in built]
    h = built[0][0]["h"]
    for step in range(150):
        for s in sl:
            s.pack()
        slabs.exchange_local(sl)
        for s in sl:
            s.unpack()
        if step == 149:
            break
        for s in sl:
            s.w.tick(h)
    st = [s.w.state() for s in sl]
    n0, n1 = built[0][0]["n_owned"], built[1][0]["n_owned"]
    # after an exchange the ghosts carry exactly the owners' states
    send0, recv0 = sl[0].sides["right"]["send_idx"].cpu().numpy(), sl[0].sides["right"]["recv_idx"].cpu().numpy()
    send1, recv1 = sl[1].sides["left"]["send_idx"].cpu().numpy(), sl[1].sides["left"]["recv_idx"].cpu().numpy()
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
