"""Synthetic scenes for the configs BASELINE.json names.

Shapes are drawn exactly as the reference's spawn code does (/root/reference/src/main.c:502-522)
from the reference's own PRNG (/root/reference/src/rand.c:7-34), with `randState` fixed to a seed
instead of time(NULL).  A scene is a plain dict of numpy arrays that both the engine binding
(odeb200.World.load_scene) and the test oracle accept.
"""
import math
import os

import numpy as np

SPHERE, BOX, PLANE, TRIMESH = 0, 1, 4, 8
BODY_KINEMATIC, BODY_NOGRAVITY, BODY_GYRO = 1, 2, 4
CMASK_MAP, CMASK_OBJ, CMASK_ALL = 1, 2, 0xFFFFFFFF


class RefRand:
    """Rand_Next / Rand_Int / Rand_Double of src/rand.c:7-34.  The generator is a Weyl sequence
    fed through two multiply-xorshift rounds, so draw k only depends on state0 + (k+1)*0xE120FC15
    and whole blocks can be drawn vectorised."""

    def __init__(self, seed):
        self.state = np.uint64(seed & 0xFFFFFFFF)

    def next_block(self, n):
        k = np.arange(1, n + 1, dtype=np.uint64)
        st = (self.state + k * np.uint64(0xE120FC15)) & np.uint64(0xFFFFFFFF)
        self.state = st[-1] if n else self.state
        t = st * np.uint64(0x4A39B70D)
        m1 = ((t >> np.uint64(32)) ^ t) & np.uint64(0xFFFFFFFF)
        t = m1 * np.uint64(0x12FAD5C9)
        return (((t >> np.uint64(32)) ^ t) & np.uint64(0xFFFFFFFF)).astype(np.uint64)

    def next(self):
        return int(self.next_block(1)[0])

    def rand_int(self, lo, hi):
        return int(self.next() % (hi - lo)) + lo

    def rand_double(self, lo, hi):
        return lo + self.next() / float(0xFFFFFFFF) * (hi - lo)

    @staticmethod
    def to_double(raw, lo, hi):
        return lo + raw.astype(np.float64) / float(0xFFFFFFFF) * (hi - lo)


def _empty_scene(name, gravity=(0.0, -9.8, 0.0), h=1.0 / 60.0):
    return {"name": name, "gravity": tuple(gravity), "h": h, "meshes": [],
            "_b": {k: [] for k in ("pos", "quat", "lvel", "avel", "mass", "inertia", "flags", "env")},
            "_g": {k: [] for k in ("type", "dims", "body", "pos", "R", "cat", "col", "env")}}


IDENT_R = (1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0)


def _add_body(sc, pos, quat=(1, 0, 0, 0), lvel=(0, 0, 0), avel=(0, 0, 0), mass=1.0,
              inertia=(1, 0, 0, 0, 1, 0, 0, 0, 1), flags=0, env=0):
    b = sc["_b"]
    b["pos"].append(pos); b["quat"].append(quat); b["lvel"].append(lvel); b["avel"].append(avel)
    b["mass"].append(mass); b["inertia"].append(inertia); b["flags"].append(flags); b["env"].append(env)
    return len(b["pos"]) - 1


def _add_geom(sc, gtype, dims, body=-1, pos=(0, 0, 0), R=IDENT_R, cat=CMASK_ALL, col=CMASK_ALL, env=-1):
    g = sc["_g"]
    d = list(dims) + [0.0] * (4 - len(dims))
    g["type"].append(gtype); g["dims"].append(d); g["body"].append(body); g["pos"].append(pos)
    g["R"].append(R); g["cat"].append(cat); g["col"].append(col); g["env"].append(env)
    return len(g["type"]) - 1


def finalize(sc):
    b, g = sc.pop("_b"), sc.pop("_g")
    nb, ng = len(b["pos"]), len(g["type"])
    sc["bodies"] = {
        "pos": np.asarray(b["pos"], np.float32).reshape(nb, 3),
        "quat": np.asarray(b["quat"], np.float32).reshape(nb, 4),
        "lvel": np.asarray(b["lvel"], np.float32).reshape(nb, 3),
        "avel": np.asarray(b["avel"], np.float32).reshape(nb, 3),
        "mass": np.asarray(b["mass"], np.float32).reshape(nb),
        "inertia": np.asarray(b["inertia"], np.float32).reshape(nb, 9),
        "flags": np.asarray(b["flags"], np.int32).reshape(nb),
        "env": np.asarray(b["env"], np.int32).reshape(nb),
    }
    sc["geoms"] = {
        "type": np.asarray(g["type"], np.int32).reshape(ng),
        "dims": np.asarray(g["dims"], np.float32).reshape(ng, 4),
        "body": np.asarray(g["body"], np.int32).reshape(ng),
        "pos": np.asarray(g["pos"], np.float32).reshape(ng, 3),
        "R": np.asarray(g["R"], np.float32).reshape(ng, 12),
        "cat": np.asarray(g["cat"], np.uint32).reshape(ng),
        "col": np.asarray(g["col"], np.uint32).reshape(ng),
        "env": np.asarray(g["env"], np.int32).reshape(ng),
    }
    return sc


def from_arrays(name, bodies, geoms, meshes=(), gravity=(0.0, -9.8, 0.0), h=1.0 / 60.0):
    return {"name": name, "gravity": tuple(gravity), "h": h, "meshes": list(meshes), "bodies": bodies, "geoms": geoms}


def transform_mat_v(pos, rot):
    """GetTransformMatV, src/main.c:624-651 (including its :639 `sx`-for-`sz` quirk), in float32."""
    f = np.float32
    cx, sx = f(math.cos(rot[0])), f(math.sin(rot[0]))
    cy, sy = f(math.cos(rot[1])), f(math.sin(rot[1]))
    cz, sz = f(math.cos(rot[2])), f(math.sin(rot[2]))
    res = np.zeros(16, np.float32)
    res[0] = cy * cz; res[1] = cz * sx * sy - cx * sz; res[2] = cx * cz * sy + sx * sz
    res[4] = cy * sz; res[5] = cx * cz + sx * sy * sz; res[6] = -cz * sx + cx * sy * sx
    res[8] = -sy; res[9] = cy * sx; res[10] = cx * cy
    res[12:15] = pos; res[15] = 1.0
    return res


def _static_map(sc, floor_plane=False):
    """The four AddBodyMap calls of src/main.c:115-121.  AddBodyMap sets the category bits twice
    (:751-752) so static geoms end with category ALL & ~MAP and the default collide bits."""
    maps = [((0, 0, 0), (0, 0, 0), (100, 1, 100)), ((4, 3, 0), (0, 0, -0.5), (0.5, 8, 12)),
            ((0, 3, 6), (0, 0, 0), (12, 8, 0.5)), ((0, 3, -6), (0, 0, 0), (12, 8, 0.5))]
    for i, (pos, rot, size) in enumerate(maps):
        if i == 0 and floor_plane:
            _add_geom(sc, PLANE, (0, 1, 0, 0.5), cat=CMASK_ALL & ~CMASK_MAP, col=CMASK_ALL)
            continue
        t = transform_mat_v(pos, rot)
        _add_geom(sc, BOX, size, pos=pos, R=tuple(t[:12]), cat=CMASK_ALL & ~CMASK_MAP, col=CMASK_ALL)


def _spawn(rng):
    """One `M`-key spawn, src/main.c:504-521: returns (pos, type, dims)."""
    pos = (rng.rand_double(-4.0, 4.0), rng.rand_double(20.0, 50.0), rng.rand_double(-4.0, 4.0))
    if rng.rand_int(0, 2) == 0:
        dims = (rng.rand_double(0.2, 1.0), rng.rand_double(0.2, 1.0), rng.rand_double(0.2, 1.0))
        gtype = BOX
    else:
        dims = (rng.rand_double(0.1, 0.4),)
        gtype = SPHERE
    for _ in range(3):  # Rand_Color draws three Rand_Int
        rng.rand_int(30, 190)
    return pos, gtype, dims


def server_scene(seed=1, n_dropped=64, n_players=4, h=1.0 / 60.0, floor_plane=False, y_range=None):
    """C1: reference server scene -- static map + dropped boxes/spheres + kinematic player spheres."""
    sc = _empty_scene("C1p" if floor_plane else "C1", h=h)
    _static_map(sc, floor_plane)
    rng = RefRand(seed)
    for _ in range(n_dropped):
        pos, gtype, dims = _spawn(rng)
        if y_range is not None:
            pos = (pos[0], y_range[0] + (pos[1] - 20.0) / 30.0 * (y_range[1] - y_range[0]), pos[2])
        b = _add_body(sc, pos, flags=BODY_GYRO)  # dBodyCreate: gyroscopic mode on (ODE >= 0.13)
        _add_geom(sc, gtype, dims, body=b, cat=CMASK_OBJ, col=CMASK_OBJ | CMASK_MAP, env=0)
    for i in range(n_players):
        b = _add_body(sc, (0.0 + 1.5 * i, 2.0, -3.0), flags=BODY_KINEMATIC | BODY_GYRO)
        _add_geom(sc, SPHERE, (0.5,), body=b, cat=CMASK_OBJ, col=CMASK_OBJ | CMASK_MAP, env=0)
    return finalize(sc)


def _lattice_bodies(rng, nx, ny, nz, spacing, origin, jitter, env, box_prob_half=True, sphere_scale=1.0,
                    spheres_only=False):
    """Vectorised jittered lattice of reference-distribution shapes. Returns body/geom arrays."""
    n = nx * ny * nz
    raw = rng.next_block(n * 7).reshape(n, 7)
    ix = np.arange(n) % nx
    iz = (np.arange(n) // nx) % nz
    iy = np.arange(n) // (nx * nz)
    jit = np.stack([RefRand.to_double(raw[:, k], -jitter, jitter) for k in range(3)], axis=1)
    pos = np.stack([origin[0] + ix * spacing, origin[1] + iy * spacing, origin[2] + iz * spacing], axis=1) + jit
    is_box = (raw[:, 3] % np.uint64(2)) == 0
    if spheres_only:
        is_box[:] = False
    dims = np.zeros((n, 4), np.float64)
    for k in range(3):
        dims[:, k] = RefRand.to_double(raw[:, 4 + k], 0.2, 1.0)
    rad = RefRand.to_double(raw[:, 4], 0.1, 0.4) * sphere_scale
    dims[~is_box, 0] = rad[~is_box]
    dims[~is_box, 1:] = 0.0
    gtype = np.where(is_box, BOX, SPHERE).astype(np.int32)
    bodies = {
        "pos": pos.astype(np.float32),
        "quat": np.tile(np.array([1, 0, 0, 0], np.float32), (n, 1)),
        "lvel": np.zeros((n, 3), np.float32), "avel": np.zeros((n, 3), np.float32),
        "mass": np.ones(n, np.float32),
        "inertia": np.tile(np.array([1, 0, 0, 0, 1, 0, 0, 0, 1], np.float32), (n, 1)),
        "flags": np.zeros(n, np.int32), "env": np.full(n, env, np.int32),
    }
    geoms = {
        "type": gtype, "dims": dims.astype(np.float32), "body": np.arange(n, dtype=np.int32),
        "pos": np.zeros((n, 3), np.float32), "R": np.tile(np.array(IDENT_R, np.float32), (n, 1)),
        "cat": np.full(n, CMASK_OBJ, np.uint32), "col": np.full(n, CMASK_OBJ | CMASK_MAP, np.uint32),
        "env": np.full(n, env, np.int32),
    }
    return bodies, geoms


def _static_geoms(entries):
    """entries: list of (type, dims4, pos3, R12, env)."""
    n = len(entries)
    return {
        "type": np.array([e[0] for e in entries], np.int32),
        "dims": np.array([list(e[1]) + [0.0] * (4 - len(e[1])) for e in entries], np.float32).reshape(n, 4),
        "body": np.full(n, -1, np.int32),
        "pos": np.array([e[2] for e in entries], np.float32).reshape(n, 3),
        "R": np.array([e[3] for e in entries], np.float32).reshape(n, 12),
        "cat": np.full(n, CMASK_ALL & ~CMASK_MAP, np.uint32), "col": np.full(n, CMASK_ALL, np.uint32),
        "env": np.array([e[4] for e in entries], np.int32),
    }


def _concat(parts):
    return {k: np.concatenate([p[k] for p in parts], axis=0) for k in parts[0]}


def pile_scene(nx=256, nz=256, ny=16, seed=3, spacing=1.8, h=1.0 / 60.0, walls=True, name="C3"):
    """C3: nx*nz*ny bodies on a jittered lattice above a plane y=0 with four plane walls."""
    rng = RefRand(seed)
    ox, oz = -0.5 * (nx - 1) * spacing, -0.5 * (nz - 1) * spacing
    bodies, geoms = _lattice_bodies(rng, nx, ny, nz, spacing, (ox, 1.0, oz), 0.03, 0)
    half_x, half_z = 0.5 * nx * spacing + 1.0, 0.5 * nz * spacing + 1.0
    st = [(PLANE, (0, 1, 0, 0.0), (0, 0, 0), IDENT_R, -1)]
    if walls:
        st += [(PLANE, (1, 0, 0, -half_x), (0, 0, 0), IDENT_R, -1), (PLANE, (-1, 0, 0, -half_x), (0, 0, 0), IDENT_R, -1),
               (PLANE, (0, 0, 1, -half_z), (0, 0, 0), IDENT_R, -1), (PLANE, (0, 0, -1, -half_z), (0, 0, 0), IDENT_R, -1)]
    sg = _static_geoms(st)
    g = _concat([sg, geoms])
    # static geoms come first; dynamic geoms keep body index = lattice index
    return from_arrays(name, bodies, g, h=h)


def batched_worlds_scene(n_worlds=8192, seed=4, nx=8, ny=4, nz=4, spacing=1.8, h=1.0 / 60.0, first_world=0):
    """C4: n_worlds independent worlds, each a plane + nx*ny*nz bodies; world w uses seed 4+w.
    `first_world` offsets the world numbering so a shard reproduces worlds [first, first+n)."""
    per = nx * ny * nz
    bparts, gparts = [], []
    ox, oz = -0.5 * (nx - 1) * spacing, -0.5 * (nz - 1) * spacing
    for w in range(n_worlds):
        rng = RefRand(seed + first_world + w)
        b, g = _lattice_bodies(rng, nx, ny, nz, spacing, (ox, 1.0, oz), 0.03, w)
        g["body"] = g["body"] + w * per
        bparts.append(b); gparts.append(g)
    sg = _static_geoms([(PLANE, (0, 1, 0, 0.0), (0, 0, 0), IDENT_R, -1)])
    return from_arrays("C4", _concat(bparts), _concat([sg] + gparts), h=h)


def load_obj(path):
    """Minimal OBJ reader: `v x y z` and triangular `f a/b/c ...` (1-based)."""
    v, f = [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith("v "):
                v.append([float(x) for x in line.split()[1:4]])
            elif line.startswith("f "):
                idx = [int(tok.split("/")[0]) - 1 for tok in line.split()[1:]]
                for k in range(1, len(idx) - 1):
                    f.append([idx[0], idx[k], idx[k + 1]])
    return np.asarray(v, np.float32), np.asarray(f, np.int32)


def teapot_mesh():
    """The teapot.obj asset (reference res/teapot.obj, 4884 vertices / 8884 triangles) as committed
    under tests/golden/teapot_mesh.npz by tests/golden/make_teapot_npz.py."""
    here = os.path.dirname(os.path.abspath(__file__))
    path = os.path.join(here, "..", "..", "tests", "golden", "teapot_mesh.npz")
    d = np.load(path)
    return d["verts"].astype(np.float32), d["tris"].astype(np.int32)


def trimesh_scene(n_side=100, seed=2, sphere_scale=5.0, h=1.0 / 60.0, mesh=None, drop=5.0):
    """C2: one static trimesh + n_side^2 spheres (radius U[0.1,0.4]*sphere_scale) dropped from a
    jittered grid `drop` units above the mesh top."""
    verts, tris = mesh if mesh is not None else teapot_mesh()
    lo, hi = verts.min(axis=0), verts.max(axis=0)
    rng = RefRand(seed)
    n = n_side * n_side
    raw = rng.next_block(n * 3).reshape(n, 3)
    ix, iz = np.arange(n) % n_side, np.arange(n) // n_side
    sx = (hi[0] - lo[0]) / n_side
    sz = (hi[2] - lo[2]) / n_side
    rad = RefRand.to_double(raw[:, 0], 0.1, 0.4) * sphere_scale
    px = lo[0] + (ix + 0.5) * sx + RefRand.to_double(raw[:, 1], -0.2, 0.2) * sx
    pz = lo[2] + (iz + 0.5) * sz + RefRand.to_double(raw[:, 2], -0.2, 0.2) * sz
    # stagger heights so the spheres do not start in contact with each other
    py = hi[1] + drop + (np.arange(n) % 7) * (2.0 * 0.4 * sphere_scale + 0.1)
    bodies = {
        "pos": np.stack([px, py, pz], axis=1).astype(np.float32),
        "quat": np.tile(np.array([1, 0, 0, 0], np.float32), (n, 1)),
        "lvel": np.zeros((n, 3), np.float32), "avel": np.zeros((n, 3), np.float32),
        "mass": np.ones(n, np.float32),
        "inertia": np.tile(np.array([1, 0, 0, 0, 1, 0, 0, 0, 1], np.float32), (n, 1)),
        "flags": np.zeros(n, np.int32), "env": np.zeros(n, np.int32),
    }
    dims = np.zeros((n, 4), np.float32)
    dims[:, 0] = rad
    geoms = {
        "type": np.full(n, SPHERE, np.int32), "dims": dims, "body": np.arange(n, dtype=np.int32),
        "pos": np.zeros((n, 3), np.float32), "R": np.tile(np.array(IDENT_R, np.float32), (n, 1)),
        "cat": np.full(n, CMASK_OBJ, np.uint32), "col": np.full(n, CMASK_OBJ | CMASK_MAP, np.uint32),
        "env": np.zeros(n, np.int32),
    }
    sg = _static_geoms([(TRIMESH, (0.0,), (0, 0, 0), IDENT_R, -1), (PLANE, (0, 1, 0, float(lo[1]) - 1.0), (0, 0, 0), IDENT_R, -1)])
    return from_arrays("C2", bodies, _concat([sg, geoms]), meshes=[(verts, tris)], h=h)


def trimesh_contact_scene(n=256, seed=11, sphere_scale=5.0, mesh=None, h=1.0 / 60.0, box_fraction=0.0):
    """Spheres placed in touch with a static trimesh (centre = triangle centroid + 0.8 r along the
    face normal): exercises sphere-vs-trimesh contacts from step 0.  With box_fraction > 0 that share of
    the bodies are randomly rotated boxes (sides 0.4..1.6 r) instead: box-vs-trimesh contacts."""
    verts, tris = mesh if mesh is not None else teapot_mesh()
    rs = np.random.RandomState(seed)
    pick = rs.choice(len(tris), size=n, replace=False)
    a, b, c = verts[tris[pick, 0]], verts[tris[pick, 1]], verts[tris[pick, 2]]
    cen = (a + b + c) / 3.0
    nrm = np.cross(b - a, c - a)
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-20)
    rad = rs.uniform(0.1, 0.4, size=n) * sphere_scale
    pos = cen + nrm * (0.8 * rad)[:, None]
    bodies = {
        "pos": pos.astype(np.float32),
        "quat": np.tile(np.array([1, 0, 0, 0], np.float32), (n, 1)),
        "lvel": np.zeros((n, 3), np.float32), "avel": np.zeros((n, 3), np.float32),
        "mass": np.ones(n, np.float32),
        "inertia": np.tile(np.array([1, 0, 0, 0, 1, 0, 0, 0, 1], np.float32), (n, 1)),
        "flags": np.zeros(n, np.int32), "env": np.zeros(n, np.int32),
    }
    dims = np.zeros((n, 4), np.float32)
    dims[:, 0] = rad
    gtype = np.full(n, SPHERE, np.int32)
    if box_fraction > 0:
        is_box = rs.uniform(size=n) < box_fraction
        sides = rs.uniform(0.4, 1.6, size=(n, 3)) * rad[:, None]
        q = rs.normal(size=(n, 4))
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        gtype[is_box] = BOX
        dims[is_box, :3] = sides[is_box]
        bodies["quat"][is_box] = q[is_box].astype(np.float32)
    geoms = {
        "type": gtype, "dims": dims, "body": np.arange(n, dtype=np.int32),
        "pos": np.zeros((n, 3), np.float32), "R": np.tile(np.array(IDENT_R, np.float32), (n, 1)),
        "cat": np.full(n, CMASK_OBJ, np.uint32), "col": np.full(n, CMASK_OBJ | CMASK_MAP, np.uint32),
        "env": np.zeros(n, np.int32),
    }
    sg = _static_geoms([(TRIMESH, (0.0,), (0, 0, 0), IDENT_R, -1)])
    return from_arrays("trimesh-contact", bodies, _concat([sg, geoms]), meshes=[(verts, tris)], h=h)


def random_soup(n=200, seed=7, extent=4.0, with_plane=True, with_static_box=True, rotated=True):
    """Dense random soup of overlapping boxes and spheres with random orientations: a stress
    case for broadphase set parity and narrowphase contact-count parity."""
    rs = np.random.RandomState(seed)
    sc = _empty_scene("soup")
    if with_plane:
        _add_geom(sc, PLANE, (0, 1, 0, -extent), cat=CMASK_ALL & ~CMASK_MAP)
    if with_static_box:
        _add_geom(sc, BOX, (2 * extent, 0.5, 2 * extent), pos=(0, -extent + 0.5, 0), cat=CMASK_ALL & ~CMASK_MAP)
    for _ in range(n):
        pos = rs.uniform(-extent, extent, 3)
        if rotated:
            q = rs.normal(size=4)
            q /= np.linalg.norm(q)
        else:
            q = np.array([1.0, 0, 0, 0])
        lv = rs.uniform(-1, 1, 3)
        av = rs.uniform(-1, 1, 3)
        b = _add_body(sc, tuple(pos), quat=tuple(q), lvel=tuple(lv), avel=tuple(av))
        if rs.randint(0, 2) == 0:
            _add_geom(sc, BOX, tuple(rs.uniform(0.2, 1.0, 3)), body=b, cat=CMASK_OBJ, col=CMASK_OBJ | CMASK_MAP, env=0)
        else:
            _add_geom(sc, SPHERE, (rs.uniform(0.1, 0.4),), body=b, cat=CMASK_OBJ, col=CMASK_OBJ | CMASK_MAP, env=0)
    return finalize(sc)
