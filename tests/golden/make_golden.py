"""Generates tests/golden/*.npz with the CPU oracle (oracle/ode_oracle.c): for every named scene the
sorted broadphase pair set, the per-pair contact lists of step 0, and the body state after one
reference tick solved in joint order (order_mode 1) and in ODE's randomised order (order_mode 0).

The reference has no golden vectors of its own (no tests, libode un-vendored), so these fixtures pin
the ORACLE (regression) and give the GPU tests committed pair/contact sets to match bit for bit.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
for p in (os.path.join(ROOT, "rl-ode-physics_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import oracle as O  # noqa: E402
from odeb200 import scenes  # noqa: E402


def golden_scenes():
    return {
        "c1_low": scenes.server_scene(seed=1, y_range=(1.0, 6.0)),
        "c1p_low": scenes.server_scene(seed=1, y_range=(1.0, 6.0), floor_plane=True),
        "soup200": scenes.random_soup(200, seed=7),
        "soup_axis": scenes.random_soup(150, seed=9, rotated=False),
        "teapot256": scenes.trimesh_contact_scene(256, seed=11),
        "batch8": scenes.batched_worlds_scene(8, seed=4, spacing=0.7),
        "teapot_boxes": scenes.trimesh_contact_scene(192, seed=13, box_fraction=0.6),
    }


def oracle_record(sc):
    types = [int(t) for t in sc["geoms"]["type"]]
    w = O.OracleWorld(gravity=sc["gravity"])
    w.load_scene(sc)
    pairs = w.broadphase(0)
    cnt, pd, nrm, side = [], [], [], []
    canon = []
    for a, b in pairs.tolist():
        g1, g2 = (a, b) if types[a] <= types[b] else (b, a)
        canon.append((g1, g2))
        cs = w.collide(g1, g2, 8)
        cnt.append(len(cs))
        for c in cs:
            pd.append(list(c.pos) + [c.depth]); nrm.append(list(c.normal)); side.append(c.side2)
    out = {"pairs": pairs.astype(np.int32), "canon": np.asarray(canon, np.int32).reshape(-1, 2),
           "count": np.asarray(cnt, np.int32), "pos_depth": np.asarray(pd, np.float32).reshape(-1, 4),
           "normal": np.asarray(nrm, np.float32).reshape(-1, 3), "side": np.asarray(side, np.int32)}
    for mode in (1, 0):
        w2 = O.OracleWorld(gravity=sc["gravity"])
        w2.load_scene(sc)
        w2.tick(sc["h"], order_mode=mode)
        st = w2.state()
        for k in ("pos", "quat", "lvel", "avel"):
            out["state%d_%s" % (mode, k)] = st[k]
    return out


if __name__ == "__main__":
    O.build()
    only = sys.argv[1:]
    for name, sc in golden_scenes().items():
        if only and name not in only:
            continue
        rec = oracle_record(sc)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
        print(name, "pairs", len(rec["pairs"]), "contacts", int(rec["count"].sum()))
