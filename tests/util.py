"""Shared helpers of the parity tests: build the same scene in the engine (through the C ABI) and
in the CPU oracle, and compare pair sets, contacts and post-step state."""
import numpy as np

import oracle as O
import odeb200
from odeb200 import scenes


def oracle_world(sc, iters=20):
    w = O.OracleWorld(gravity=sc["gravity"], iters=iters)
    w.load_scene(sc)
    return w


def engine_world(sc, iters=20):
    w = odeb200.World(gravity=sc["gravity"], iters=iters)
    w.load_scene(sc)
    return w


def sorted_pair_set(pairs):
    p = np.sort(np.asarray(pairs, np.int64).reshape(-1, 2), axis=1)
    if len(p) == 0:
        return p
    order = np.lexsort((p[:, 1], p[:, 0]))
    return p[order]


def oracle_contacts(ow, maxc=8):
    """dict (g1,g2) canonical -> list of ContactGeom, using the engine's callback order rule."""
    out = {}
    gt = ow_geom_types(ow)
    for a, b in ow.broadphase(0):
        g1, g2 = int(a), int(b)
        if gt[g1] > gt[g2]:
            g1, g2 = g2, g1
        out[(g1, g2)] = ow.collide(g1, g2, maxc)
    return out


def ow_geom_types(ow):
    return ow._types


def load_both(sc, iters=20):
    ow = oracle_world(sc, iters)
    ow._types = [int(t) for t in sc["geoms"]["type"]]
    ow._bodies = [int(b) for b in sc["geoms"]["body"]]
    ew = engine_world(sc, iters)
    return ow, ew


def oracle_tick_in_engine_order(ow, ew, h, maxc=8, surf=None, rows_per_contact=3, geom_map=None, order=None):
    """Run the oracle's collide + QuickStep with the row order the engine used for its last step.
    Returns the number of oracle contacts.

    geom_map: when the oracle world is a PART of the engine's world (one env of a batch), an int array
    engine geom id -> oracle geom id (-1: not in the oracle world); the engine's solver order is then
    restricted to the units whose geoms are all in the oracle world.  order: a solver order fetched earlier."""
    surf = surf or O.reference_surface()
    ow.clear_contacts()
    nc = ow.collide_all(maxc, surf)
    # oracle joint order = sorted pairs, contacts k ascending; rebuild the same enumeration here
    gt = ow._types
    gb = ow._bodies
    joint_of = {}
    row0 = {}
    j = 0
    rows = 0
    for a, b in ow.broadphase(0):
        g1, g2 = int(a), int(b)
        if gt[g1] > gt[g2]:
            g1, g2 = g2, g1
        n = len(ow.collide(g1, g2, maxc))
        active = gb[g1] >= 0 or gb[g2] >= 0
        for k in range(n):
            joint_of[(g1, g2, k)] = j
            if active:
                row0[j] = rows
                rows += rows_per_contact
            j += 1
    assert j == nc
    eg1, eg2, ek = order if order is not None else ew.solver_order()
    if geom_map is not None:
        gm = np.asarray(geom_map)
        m1, m2 = gm[eg1], gm[eg2]
        keep = (m1 >= 0) & (m2 >= 0)
        eg1, eg2, ek = m1[keep], m2[keep], ek[keep]
    perm = []
    for a, b, k in zip(eg1, eg2, ek):
        jj = joint_of[(int(a), int(b), int(k))]
        r = row0[jj]
        perm += [r + i for i in range(rows_per_contact)]
    assert len(perm) == rows, (len(perm), rows)
    assert sorted(perm) == list(range(rows))
    ow.quickstep(h, order_mode=2, perm=np.asarray(perm, np.int32))
    ow.clear_contacts()
    return nc


def rel_err(a, b, floor=1e-3):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)
