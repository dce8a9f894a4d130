"""CPU: INDEPENDENT pins of the oracle (oracle/ode_oracle.c).

libode is absent from the reference tree and from this image, so the oracle cannot be pinned against
the reference's own build ("parity unpinned").  The known-answer tests in
test_oracle_kat.py pin single cases by hand; this file pins the restatement against SECOND,
differently written implementations in float64 numpy on thousands of random inputs:

* box-box: a 15-axis float64 SAT overlap predicate  <=>  dBoxBox returns contacts; the returned normal is
  a near-minimal overlap axis (the 1.05 edge fudge), depths never exceed the overlap along it;
* sphere-box / sphere-plane / box-plane / sphere-sphere: depth and position against brute-force geometry
  (closest point on a box by clamping in float64, all 8 corners of a box against a plane);
* sphere-trimesh (engine's own rule): deepest contact = r - min distance to any triangle, with the
  point-triangle distance computed by a different method (plane projection + three segment tests);
* rows (dxJointContact::getInfo2 + QuickStep rhs): J, c and rhs re-derived in float64 from the contact
  geometry and body state;
* SOR_LCP: the 20-sweep lambda against a float64 projected-SOR on the DENSE matrix form
  A = J M^-1 J^T + cfm/h (the oracle, like ODE, never forms A: it carries the accumulator fc = M^-1 J^T lambda);
* the velocity update v+ = v + h M^-1 (f + J^T lambda) and the LCP residual J v+ - c + cfm lambda of the sweeps;
* order_mode 3 (dWorldStep's answer): KKT conditions of the bounded LCP on the dense matrix.
"""
import numpy as np
import pytest

import oracle as O
from odeb200 import scenes

H = 1.0 / 60.0


# ------------------------------------------------------------------ geometry helpers (float64, independent)

def _rand_rot(rs):
    q = rs.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    return q, R


def _body_R64(w, b):
    pos, q, R, lv, av = w.body(b)
    return pos.astype(np.float64), R.astype(np.float64).reshape(3, 4)[:, :3]


def _sat(pa, Ra, ha, pb, Rb, hb):
    """15-axis separating-axis test of two oriented boxes.  Returns (min signed overlap over all axes,
    overlap per axis list, axes).  overlap > 0 on every axis <=> the boxes intersect."""
    d = pb - pa
    axes = [Ra[:, i] for i in range(3)] + [Rb[:, i] for i in range(3)]
    for i in range(3):
        for j in range(3):
            c = np.cross(Ra[:, i], Rb[:, j])
            n = np.linalg.norm(c)
            if n > 1e-9:
                axes.append(c / n)
    ov = []
    for a in axes:
        ra = sum(ha[i] * abs(np.dot(a, Ra[:, i])) for i in range(3))
        rb = sum(hb[i] * abs(np.dot(a, Rb[:, i])) for i in range(3))
        ov.append(ra + rb - abs(np.dot(a, d)))
    return min(ov), ov, axes


def _overlap_along(n, pa, Ra, ha, pb, Rb, hb):
    ra = sum(ha[i] * abs(np.dot(n, Ra[:, i])) for i in range(3))
    rb = sum(hb[i] * abs(np.dot(n, Rb[:, i])) for i in range(3))
    return ra + rb - abs(np.dot(n, pb - pa))


def test_box_box_contacts_iff_float64_sat_overlap(oracle_lib):
    rs = np.random.RandomState(1234)
    n_hit = n_miss = n_skipped = 0
    for trial in range(4000):
        w = O.OracleWorld()
        sides, poses = [], []
        for k in range(2):
            q, R = _rand_rot(rs)
            s = rs.uniform(0.2, 1.0, 3)
            p = rs.uniform(-0.55, 0.55, 3) if k else np.zeros(3)
            b = w.add_body(p, q=q)
            w.add_geom(O.BOX, s, body=b)
            sides.append(s)
        (pa, Ra), (pb, Rb) = _body_R64(w, 0), _body_R64(w, 1)
        ha, hb = np.float32(sides[0]).astype(np.float64) * 0.5, np.float32(sides[1]).astype(np.float64) * 0.5
        sep, ov, axes = _sat(pa, Ra, ha, pb, Rb, hb)
        cs = w.collide(0, 1, 8)
        w.close()
        if abs(sep) < 2e-5:       # too close to tangency for a float32 predicate to be decidable
            n_skipped += 1
            continue
        assert (len(cs) > 0) == (sep > 0), (trial, sep, len(cs))
        if not cs:
            n_miss += 1
            continue
        n_hit += 1
        n = np.array(cs[0].normal, np.float64)
        assert abs(np.linalg.norm(n) - 1.0) < 1e-5
        for c in cs:
            assert np.allclose(np.array(c.normal), n)                 # one normal per manifold
        ovn = _overlap_along(n, pa, Ra, ha, pb, Rb, hb)
        # dBoxBox picks the axis of least overlap, preferring faces: an edge axis must beat the best face by 1.05
        assert ovn <= 1.05 * sep + 1e-5, (trial, ovn, sep)
        # normal points from box 2 into box 1
        assert np.dot(n, pa - pb) > -1e-6
        depths = np.array([c.depth for c in cs], np.float64)
        assert depths.max() <= ovn + 1e-4 and depths.min() >= -1e-6
        assert 1 <= len(cs) <= 8
        # every contact point lies in the (slightly inflated) intersection slab of the two boxes along n
        for c in cs:
            p = np.array(c.pos, np.float64)
            for (pc, Rc, hc) in ((pa, Ra, ha), (pb, Rb, hb)):
                loc = Rc.T @ (p - pc)
                assert (np.abs(loc) <= hc + ovn + 1e-4).all(), (trial, loc, hc)
    assert n_hit > 500 and n_miss > 500 and n_skipped < 50


def test_sphere_box_against_float64_closest_point(oracle_lib):
    rs = np.random.RandomState(5)
    hits = 0
    for trial in range(3000):
        w = O.OracleWorld()
        q, R = _rand_rot(rs)
        s = rs.uniform(0.2, 1.0, 3)
        bb = w.add_body(np.zeros(3), q=q); w.add_geom(O.BOX, s, body=bb)
        r = rs.uniform(0.1, 0.4)
        p = rs.uniform(-0.9, 0.9, 3)
        bs = w.add_body(p); w.add_geom(O.SPHERE, [r], body=bs)
        pb, Rb = _body_R64(w, bb)
        ps, _ = _body_R64(w, bs)
        hb = np.float32(s).astype(np.float64) * 0.5
        r64 = float(np.float32(r))
        loc = Rb.T @ (ps - pb)
        cl = np.clip(loc, -hb, hb)
        dist = np.linalg.norm(loc - cl)
        inside = (np.abs(loc) < hb).all()
        cs = w.collide(1, 0, 8)      # (sphere, box): ODE's stored collider order
        w.close()
        if inside:
            depth = r64 + (hb - np.abs(loc)).min()
            assert len(cs) == 1 and cs[0].depth == pytest.approx(depth, abs=2e-5)
            hits += 1
            continue
        if abs(dist - r64) < 2e-5:
            continue
        assert (len(cs) == 1) == (dist < r64), (trial, dist, r64)
        if cs:
            hits += 1
            assert cs[0].depth == pytest.approx(r64 - dist, abs=2e-5)
            assert np.allclose(np.array(cs[0].pos), pb + Rb @ cl, atol=2e-5)       # contact at the closest point
            nrm = (loc - cl) / dist
            assert np.allclose(np.array(cs[0].normal), Rb @ nrm, atol=2e-4)        # from the box into the sphere
    assert hits > 300


def test_box_plane_against_all_eight_corners(oracle_lib):
    rs = np.random.RandomState(6)
    seen = set()
    for trial in range(3000):
        w = O.OracleWorld()
        nrm = rs.normal(size=3); nrm /= np.linalg.norm(nrm)
        nrm32 = np.float32(nrm)
        d = float(np.float32(rs.uniform(-0.3, 0.3)))
        w.add_geom(O.PLANE, [nrm32[0], nrm32[1], nrm32[2], d])
        q, R = _rand_rot(rs)
        s = rs.uniform(0.2, 1.0, 3)
        b = w.add_body(rs.uniform(-0.5, 0.5, 3), q=q); w.add_geom(O.BOX, s, body=b)
        pb, Rb = _body_R64(w, b)
        hb = np.float32(s).astype(np.float64) * 0.5
        n64 = nrm32.astype(np.float64)
        corners = np.array([pb + Rb @ (hb * np.array([sx, sy, sz])) for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)])
        dep = d - corners @ n64            # penetration of each corner (n is normalised to ~1e-7)
        cs = w.collide(1, 0, 8)            # (box, plane)
        w.close()
        if abs(dep.max()) < 2e-5:
            continue
        assert (len(cs) > 0) == (dep.max() > 0), (trial, dep.max())
        if not cs:
            continue
        seen.add(len(cs))
        assert len(cs) <= 4
        assert cs[0].depth == pytest.approx(dep.max(), abs=3e-5)          # first contact = deepest corner
        for c in cs:
            p = np.array(c.pos, np.float64)
            k = np.argmin(np.linalg.norm(corners - p, axis=1))
            assert np.linalg.norm(corners[k] - p) < 5e-5                  # every contact sits on a corner
            assert c.depth == pytest.approx(dep[k], abs=5e-5) and c.depth >= 0
            assert np.allclose(np.array(c.normal), n64, atol=1e-6)
    assert seen >= {1, 2, 3, 4}


def test_sphere_sphere_and_sphere_plane_against_float64(oracle_lib):
    rs = np.random.RandomState(7)
    for trial in range(2000):
        w = O.OracleWorld()
        r1, r2 = rs.uniform(0.1, 0.4, 2)
        p2 = rs.uniform(-0.5, 0.5, 3)
        b1 = w.add_body(np.zeros(3)); w.add_geom(O.SPHERE, [r1], body=b1)
        b2 = w.add_body(p2); w.add_geom(O.SPHERE, [r2], body=b2)
        nrm = np.float32(rs.normal(size=3)); nrm /= np.float32(np.linalg.norm(nrm.astype(np.float64)))
        d = float(np.float32(rs.uniform(-0.3, 0.3)))
        w.add_geom(O.PLANE, [nrm[0], nrm[1], nrm[2], d])
        x2, _ = _body_R64(w, b2)
        R1, R2 = float(np.float32(r1)), float(np.float32(r2))
        dist = np.linalg.norm(x2)
        cs = w.collide(0, 1, 8)
        if abs(dist - (R1 + R2)) > 2e-5:
            assert (len(cs) == 1) == (dist < R1 + R2)
            if cs:
                assert cs[0].depth == pytest.approx(R1 + R2 - dist, abs=2e-6)
                assert np.allclose(np.array(cs[0].normal), -x2 / dist, atol=1e-5)   # from sphere 2 into sphere 1
        cp = w.collide(1, 2, 8)           # (sphere, plane)
        dep = d - float(np.dot(x2, nrm.astype(np.float64))) + R2
        if abs(dep) > 2e-5:
            assert (len(cp) == 1) == (dep > 0)
            if cp:
                assert cp[0].depth == pytest.approx(dep, abs=2e-6)
                assert np.allclose(np.array(cp[0].pos), x2 - R2 * nrm.astype(np.float64), atol=1e-5)
        w.close()


def _pt_tri_dist(p, a, b, c):
    """distance point-triangle by plane projection + the three edge segments (NOT Ericson's region walk,
    which is what the oracle and the kernel use)"""
    n = np.cross(b - a, c - a)
    nn = np.linalg.norm(n)
    best = np.inf
    if nn > 0:
        n = n / nn
        t = np.dot(p - a, n)
        proj = p - t * n
        inside = True
        for u, v in ((a, b), (b, c), (c, a)):
            if np.dot(np.cross(v - u, proj - u), n) < 0:
                inside = False
        if inside:
            best = abs(t)
    for u, v in ((a, b), (b, c), (c, a)):
        e = v - u
        s = np.clip(np.dot(p - u, e) / max(np.dot(e, e), 1e-300), 0.0, 1.0)
        best = min(best, np.linalg.norm(p - (u + s * e)))
    return best


def test_sphere_trimesh_deepest_contact_against_bruteforce_distance(oracle_lib):
    sc = scenes.trimesh_contact_scene(96, seed=3)
    verts, tris = sc["meshes"][0]
    V = verts.astype(np.float64)
    w = O.OracleWorld()
    w.load_scene(sc)
    A, B, Cc = V[tris[:, 0]], V[tris[:, 1]], V[tris[:, 2]]
    lo = np.minimum(np.minimum(A, B), Cc); hi = np.maximum(np.maximum(A, B), Cc)
    checked = 0
    for g in range(1, 97):
        pos, _ = _body_R64(w, g - 1)
        r = float(sc["geoms"]["dims"][g, 0])
        near = np.nonzero(((lo <= pos + r).all(axis=1)) & ((hi >= pos - r).all(axis=1)))[0]
        dmin = min([_pt_tri_dist(pos, A[t], B[t], Cc[t]) for t in near] + [np.inf])
        cs = w.collide(g, 0, 8)           # (sphere, trimesh)
        if abs(dmin - r) < 1e-4:
            continue
        assert (len(cs) > 0) == (dmin < r), (g, dmin, r)
        if cs:
            checked += 1
            assert max(c.depth for c in cs) == pytest.approx(r - dmin, abs=2e-4)
            for c in cs:
                t = c.side2
                d = _pt_tri_dist(pos, A[t], B[t], Cc[t])
                assert c.depth == pytest.approx(r - d, abs=2e-4)         # each contact's depth belongs to its triangle
                assert np.linalg.norm(np.array(c.pos) - pos) == pytest.approx(d, abs=2e-4)
    w.close()
    assert checked > 60


# ------------------------------------------------------------------ rows + solver (float64, dense matrix form)

def _dense_problem(w, rows, state, gravity=(0.0, -9.8, 0.0)):
    """A = J M^-1 J^T + diag(cfm/h) and M^-1 (block diagonal) in float64 from the PRE-step state."""
    nb = len(state["pos"])
    m = len(rows["c"])
    Minv = np.zeros((6 * nb, 6 * nb))
    for i in range(nb):
        mass, invI_w = state["_mass"][i]
        Minv[6 * i:6 * i + 3, 6 * i:6 * i + 3] = np.eye(3) * (0.0 if mass == 0 else 1.0 / mass)
        Minv[6 * i + 3:6 * i + 6, 6 * i + 3:6 * i + 6] = invI_w
    Jd = np.zeros((m, 6 * nb))
    for r in range(m):
        b1, b2 = rows["jb"][r]
        Jd[r, 6 * b1:6 * b1 + 6] += rows["J"][r, :6]
        if b2 >= 0:
            Jd[r, 6 * b2:6 * b2 + 6] += rows["J"][r, 6:]
    A = Jd @ Minv @ Jd.T + np.diag(rows["cfm"].astype(np.float64))
    return A, Jd, Minv


def _soup_world(seed, n=60, iters=20):
    sc = scenes.random_soup(n, seed=seed, extent=2.2)
    b = sc["bodies"]
    rs = np.random.RandomState(seed)
    masses = []
    for i in range(n):                     # non-trivial masses and inertias so M^-1 matters
        mval = float(np.float32(rs.uniform(0.5, 3.0)))
        I = np.diag(rs.uniform(0.05, 0.4, 3)).astype(np.float32)
        b["mass"][i] = mval
        b["inertia"][i] = I.reshape(9)
    w = O.OracleWorld(gravity=sc["gravity"], iters=iters)
    w.load_scene(sc)
    st = w.state()
    st["_mass"] = []
    for i in range(n):
        R = st["R"][i].astype(np.float64).reshape(3, 4)[:, :3]
        Ib = b["inertia"][i].astype(np.float64).reshape(3, 3)
        st["_mass"].append((float(b["mass"][i]), R @ np.linalg.inv(Ib) @ R.T))
    return sc, w, st


@pytest.mark.parametrize("seed", [3, 4, 5])
def test_rows_and_rhs_against_float64_rederivation(oracle_lib, seed):
    sc, w, st = _soup_world(seed)
    h = sc["h"]
    w.keep_rows()
    # remember the contacts in joint order
    contacts = []
    types = sc["geoms"]["type"]; gbody = sc["geoms"]["body"]
    for a, b in w.broadphase(0):
        g1, g2 = (int(a), int(b)) if types[a] <= types[b] else (int(b), int(a))
        for c in w.collide(g1, g2, 8):
            contacts.append((c, gbody[g1], gbody[g2]))
    assert w.collide_all(8) == len(contacts)
    w.quickstep(h, order_mode=1)
    rows = w.last_rows()
    act = [(c, b1, b2) for (c, b1, b2) in contacts if b1 >= 0 or b2 >= 0]
    assert len(rows["c"]) == 3 * len(act) and len(act) > 40
    g = np.array(sc["gravity"])
    for k, (c, b1, b2) in enumerate(act):
        n = np.array(c.normal, np.float64); p = np.array(c.pos, np.float64)
        if b1 < 0:                                  # dJointAttach: NULL first body -> swap + reverse
            b1, b2, n = b2, -1, -n
        assert tuple(rows["jb"][3 * k]) == (b1, b2)
        x1 = st["pos"][b1].astype(np.float64)
        J = rows["J"][3 * k].astype(np.float64)
        assert np.allclose(J[:3], n, atol=1e-6) and np.allclose(J[3:6], np.cross(p - x1, n), atol=2e-6)
        vrel = J[:3] @ st["lvel"][b1] + J[3:6] @ st["avel"][b1]
        if b2 >= 0:
            x2 = st["pos"][b2].astype(np.float64)
            assert np.allclose(J[6:9], -n, atol=1e-6) and np.allclose(J[9:], -np.cross(p - x2, n), atol=2e-6)
            vrel += J[6:9] @ st["lvel"][b2] + J[9:] @ st["avel"][b2]
        cexp = 0.2 / h * max(c.depth, 0.0)          # ERP push-out
        if -vrel > 0.1:                              # bounce (reference surface: 0.2, threshold 0.1)
            cexp = max(cexp, -0.2 * vrel)
        assert rows["c"][3 * k] == pytest.approx(cexp, rel=2e-5, abs=2e-5)
        assert rows["lo"][3 * k] == 0 and np.isposinf(rows["hi"][3 * k])
        for t in (1, 2):                             # tangents: orthonormal frame with n, unbounded (mu = inf), c = 0
            Jt = rows["J"][3 * k + t].astype(np.float64)
            assert abs(Jt[:3] @ n) < 1e-6 and abs(np.linalg.norm(Jt[:3]) - 1) < 1e-6
            assert np.allclose(Jt[3:6], np.cross(p - x1, Jt[:3]), atol=2e-6)
            assert rows["c"][3 * k + t] == 0 and np.isneginf(rows["lo"][3 * k + t]) and np.isposinf(rows["hi"][3 * k + t])
        assert abs(rows["J"][3 * k + 1][:3].astype(np.float64) @ rows["J"][3 * k + 2][:3]) < 1e-6
    # rhs = c/h - J (v/h + M^-1 f_ext),  cfm = 1e-5 / h
    A, Jd, Minv = _dense_problem(w, rows, st)
    v = np.concatenate([np.concatenate([st["lvel"][i], st["avel"][i]]) for i in range(len(st["pos"]))]).astype(np.float64)
    fext = np.concatenate([np.concatenate([st["_mass"][i][0] * g, np.zeros(3)]) for i in range(len(st["pos"]))])
    rhs = rows["c"].astype(np.float64) / h - Jd @ (v / h + Minv @ fext)
    assert np.allclose(rows["rhs"], rhs, rtol=2e-4, atol=2e-3 * np.abs(rhs).max() * 1e-2 + 1e-3)
    assert np.allclose(rows["cfm"], 1e-5 / h, rtol=1e-6)
    w.close()


@pytest.mark.parametrize("seed", [3, 4, 5])
def test_sor_lcp_against_float64_dense_projected_sor(oracle_lib, seed):
    """20 sweeps of SOR_LCP (w = 1.3, order = row index) == 20 sweeps of projected SOR on the dense A,
    written from the textbook formula  x_i <- clamp(x_i + w (b_i - A_i x) / A_ii)."""
    sc, w, st = _soup_world(seed)
    h = sc["h"]
    w.keep_rows()
    w.collide_all(8)
    w.quickstep(h, order_mode=1)
    rows = w.last_rows()
    A, Jd, Minv = _dense_problem(w, rows, st)
    b = rows["rhs"].astype(np.float64)
    lo, hi = rows["lo"].astype(np.float64), rows["hi"].astype(np.float64)
    m = len(b)
    x = np.zeros(m)
    # ODE's update is x_i += Ad_i (b_i - A_i x) with Ad = w / (J M^-1 J^T_ii + cfm) -- the same thing
    for it in range(20):
        for i in range(m):
            x[i] = min(max(x[i] + 1.3 * (b[i] - A[i] @ x) / A[i, i], lo[i]), hi[i])
    lam = rows["lam"].astype(np.float64)
    scale = np.abs(x).max()
    assert scale > 1.0
    assert np.abs(lam - x).max() <= 2e-3 * scale, np.abs(lam - x).max() / scale
    # velocity update: v+ = v + h M^-1 (f_ext + J^T lambda) with the ORACLE's lambda
    nb = len(st["pos"])
    v = np.concatenate([np.concatenate([st["lvel"][i], st["avel"][i]]) for i in range(nb)]).astype(np.float64)
    g = np.array(sc["gravity"])
    fext = np.concatenate([np.concatenate([st["_mass"][i][0] * g, np.zeros(3)]) for i in range(nb)])
    vplus = v + h * (Minv @ (fext + Jd.T @ lam))
    s2 = w.state()
    got = np.concatenate([np.concatenate([s2["lvel"][i], s2["avel"][i]]) for i in range(nb)]).astype(np.float64)
    assert np.abs(got - vplus).max() <= 1e-4 * max(1.0, np.abs(vplus).max())
    # and dxStepBody: x+ = x + h v+
    assert np.allclose(s2["pos"], st["pos"] + h * s2["lvel"], atol=1e-5)
    w.close()


def lcp_residual(rows, vplus, h):
    """Velocity-level residual of every row after a step: w = J v+ - c + cfm lambda (cfm = the row's CFM, so the
    stored cfm/h times h).  Returns (violation per row, active mask): a free row (lo < lambda < hi) must have
    w = 0; a row at lo may have w >= 0; at hi w <= 0."""
    m = len(rows["c"])
    wv = np.zeros(m)
    for r in range(m):
        b1, b2 = rows["jb"][r]
        J = rows["J"][r].astype(np.float64)
        s = J[:6] @ vplus[b1]
        if b2 >= 0:
            s += J[6:] @ vplus[b2]
        wv[r] = s - float(rows["c"][r]) + float(rows["cfm"][r]) * h * float(rows["lam"][r])
    lam = rows["lam"].astype(np.float64)
    at_lo = lam <= rows["lo"]
    at_hi = lam >= rows["hi"]
    viol = np.where(at_lo, np.maximum(0.0, -wv), np.where(at_hi, np.maximum(0.0, wv), np.abs(wv)))
    return viol, ~(at_lo | at_hi)


def test_exact_mode_satisfies_the_kkt_conditions_of_the_dense_lcp(oracle_lib):
    sc, w, st = _soup_world(9, n=40)
    h = sc["h"]
    w.keep_rows()
    w.collide_all(8)
    w.quickstep(h, order_mode=3)
    rows = w.last_rows()
    A, Jd, Minv = _dense_problem(w, rows, st)
    lam = rows["lam"].astype(np.float64)
    wv = A @ lam - rows["rhs"].astype(np.float64)
    free = (lam > rows["lo"]) & (lam < rows["hi"])
    # the float64 solution is written back as float32 lambda: residual <= |A| |lambda| 2^-24 row by row
    bound = 16.0 * (np.abs(A) @ np.abs(lam)) * 2.0 ** -24 + 1e-9
    assert (np.abs(wv[free]) <= bound[free]).all(), (np.abs(wv[free]) / bound[free]).max()
    assert (wv[lam <= rows["lo"]] >= -bound[lam <= rows["lo"]]).all()
    assert (lam >= rows["lo"]).all() and free.sum() > 50
    # velocity-level residual J v+ - c + cfm lambda of the exact step, and of 20 sweeps on the same problem:
    # the sweeps are visibly not the solution -- lcp_residual measures convergence (used by the 600-step tests)
    s2 = w.state()
    vplus = np.concatenate([s2["lvel"], s2["avel"]], axis=1).astype(np.float64)
    viol, active = lcp_residual(rows, vplus, h)
    assert viol.max() < 2e-2 and active.sum() > 50          # m/s, with lambda up to 4.5e5 in this overlapping soup
    w.close()
    sc, w, st = _soup_world(9, n=40)
    w.keep_rows()
    w.collide_all(8)
    w.quickstep(h, order_mode=1)
    s2 = w.state()
    viol20, _ = lcp_residual(w.last_rows(), np.concatenate([s2["lvel"], s2["avel"]], axis=1).astype(np.float64), h)
    assert viol20.max() > 100 * viol.max()
    w.close()


@pytest.mark.parametrize("h", [1.0 / 60.0, 1.0 / 120.0])
def test_c1_600_ticks_oracle_residual_penetration_energy(oracle_lib, h):
    """SURVEY 4(4) on the oracle alone (the GPU runs the same check alongside it in test_baseline_configs_gpu.py):
    reference server scene, spawn heights y in [20, 50], h = 1/60 and the reference's 1/120 (src/main.c:208)."""
    sc = scenes.server_scene(seed=1, h=h)
    w = O.OracleWorld(gravity=sc["gravity"])
    w.load_scene(sc)
    w.keep_rows()
    nd = 64

    def energy(s):
        return float(0.5 * (s["lvel"][:nd].astype(np.float64) ** 2).sum() + 0.5 * (s["avel"][:nd].astype(np.float64) ** 2).sum() +
                     9.8 * s["pos"][:nd, 1].astype(np.float64).sum())
    E0 = Eprev = energy(w.state())
    max_inc = max_res = late_res = 0.0
    vmax = np.sqrt(2 * 9.8 * 50.0) + 0.5
    for step in range(600):
        w.collide_all(8)
        w.quickstep(h, order_mode=1)
        rows = w.last_rows()
        w.clear_contacts()
        s = w.state()
        E = energy(s)
        max_inc, Eprev = max(max_inc, E - Eprev), E
        if len(rows["c"]):
            viol, _ = lcp_residual(rows, np.concatenate([s["lvel"], s["avel"]], axis=1).astype(np.float64), h)
            max_res = max(max_res, float(viol.max()))
            if step >= 450:
                late_res = max(late_res, float(viol.max()))
    assert max_res < 0.5 * vmax and late_res < 1.5
    assert max_inc <= 1e-3 * E0 and Eprev < 0.1 * E0
    assert s["pos"][:nd, 1].min() > 0.5
    w.close()
