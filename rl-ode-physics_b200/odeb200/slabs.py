"""Slab decomposition of one large world across GPUs (BASELINE config 5).

The lattice pile is cut along x into one slab per rank.  A rank owns the bodies of its slab; lattice
columns of a neighbour nearest the shared face are mirrored as GHOST bodies whose states are refreshed from
their owners every tick (one message per face, device to device: NCCL send/recv over NVLink, or a plain
device copy when several slabs are emulated on one GPU).  Two couplings:

* ``"impulse"`` (default): a contact across a face is owned by the LOWER
  slab.  Rank r mirrors only the boundary columns of rank r+1, as DYNAMIC ghosts with the owners' masses, so
  its solver treats the contact as the two-body constraint it is.  After the step the velocity change the
  contacts gave each ghost (h * fc, `dWorldPackImpulsesDeviceB200`) travels back to the owner, which adds it
  to its body (`dWorldAddImpulsesDeviceB200`): equal and opposite impulses, one tick late on the upper side.
  Per tick: states go down (r+1 -> r), impulses go up (r -> r+1).
* ``"ghost"``: both neighbours mirror each other's boundary columns as KINEMATIC ghosts; a contact across the
  face pushes only the owned body on each side (each side sees an immovable obstacle).  One exchange per tick.

Ghost geoms collide with owned bodies only (category bits), never with each other or with static geoms.
`slab_scene` / `SlabWorld` fix the boundary membership by the initial lattice column; `dynamic_slab_scene` /
`DynamicSlabWorld` (further down) re-select it from the current positions every tick and migrate bodies between
slabs.
"""
import numpy as np

from . import scenes

CAT_MAP, CAT_OBJ, CAT_GHOST = 1, 2, 4


def _slab_lattice(slab, nx, nz, ny, seed, spacing, x0, z0):
    """bodies + geoms of slab `slab` (its own PRNG stream, so neighbours can regenerate it exactly)."""
    rng = scenes.RefRand(seed + 7919 * slab)
    b, g = scenes._lattice_bodies(rng, nx, ny, nz, spacing, (x0 + slab * nx * spacing, 1.0, z0), 0.03, 0)
    ix = np.arange(nx * ny * nz) % nx
    return b, g, ix


def _halo_side():
    return {"send_state": None, "recv_state": None, "send_imp": None, "recv_imp": None}


def slab_scene(rank, n_slabs, nx_per_slab=128, nz=1024, ny=16, seed=5, spacing=1.8, margin_cols=4, h=1.0 / 60.0,
               coupling="impulse"):
    """Scene of one rank + its halo lists.

    Returns (scene, halo) with halo[side] (side = "left" | "right") = None at the ends of the row, else a dict of
    body-index arrays: `send_state` (own boundary bodies whose states go to that neighbour, ascending),
    `recv_state` (ghost bodies refreshed from that neighbour, in its send order), `send_imp` (ghosts whose
    impulses go back to that neighbour) and `recv_imp` (own bodies that take impulses from it)."""
    assert coupling in ("impulse", "ghost")
    nx_total = nx_per_slab * n_slabs
    x0 = -0.5 * (nx_total - 1) * spacing
    z0 = -0.5 * (nz - 1) * spacing
    own_b, own_g, ix = _slab_lattice(rank, nx_per_slab, nz, ny, seed, spacing, x0, z0)
    n_own = len(own_b["pos"])
    own_g["cat"][:] = CAT_OBJ
    own_g["col"][:] = CAT_OBJ | CAT_MAP | CAT_GHOST
    bparts, gparts = [own_b], [own_g]
    halo = {"left": None, "right": None}
    n_bodies = n_own
    for side, nbr in (("left", rank - 1), ("right", rank + 1)):
        if nbr < 0 or nbr >= n_slabs:
            continue
        hs = _halo_side()
        # our columns facing the neighbour
        sel_o = (np.nonzero(ix < margin_cols)[0] if side == "left" else np.nonzero(ix >= nx_per_slab - margin_cols)[0]).astype(np.int32)
        mirror = coupling == "ghost" or side == "right"   # impulse coupling: the lower slab owns the face
        if mirror:
            nb, ng, nix = _slab_lattice(nbr, nx_per_slab, nz, ny, seed, spacing, x0, z0)
            sel_n = np.nonzero(nix >= nx_per_slab - margin_cols)[0] if side == "left" else np.nonzero(nix < margin_cols)[0]
            gb = {k: v[sel_n].copy() for k, v in nb.items()}
            gg = {k: v[sel_n].copy() for k, v in ng.items()}
            if coupling == "ghost":
                gb["flags"][:] = scenes.BODY_KINEMATIC
            gg["body"] = np.arange(n_bodies, n_bodies + len(sel_n), dtype=np.int32)
            gg["cat"][:] = CAT_GHOST
            gg["col"][:] = 0
            hs["recv_state"] = gg["body"].copy()
            if coupling == "impulse":
                hs["send_imp"] = gg["body"].copy()
            n_bodies += len(sel_n)
            bparts.append(gb)
            gparts.append(gg)
        if coupling == "ghost" or side == "left":
            hs["send_state"] = sel_o
        if coupling == "impulse" and side == "left":
            hs["recv_imp"] = sel_o.copy()
        halo[side] = hs
    half_x = 0.5 * nx_total * spacing + 1.0
    half_z = 0.5 * nz * spacing + 1.0
    st = scenes._static_geoms([(scenes.PLANE, (0, 1, 0, 0.0), (0, 0, 0), scenes.IDENT_R, -1),
                               (scenes.PLANE, (1, 0, 0, -half_x), (0, 0, 0), scenes.IDENT_R, -1),
                               (scenes.PLANE, (-1, 0, 0, -half_x), (0, 0, 0), scenes.IDENT_R, -1),
                               (scenes.PLANE, (0, 0, 1, -half_z), (0, 0, 0), scenes.IDENT_R, -1),
                               (scenes.PLANE, (0, 0, -1, -half_z), (0, 0, 0), scenes.IDENT_R, -1)])
    st["cat"][:] = CAT_MAP
    st["col"][:] = CAT_OBJ
    bodies = scenes._concat(bparts)
    geoms = scenes._concat([st] + gparts)
    sc = scenes.from_arrays("C5-slab%d" % rank, bodies, geoms, h=h)
    sc["n_owned"] = n_own
    sc["coupling"] = coupling
    return sc, halo


# message kinds: (index list to gather, index list to scatter, floats per body)
_KINDS = {"state": ("send_state", "recv_state", 16), "imp": ("send_imp", "recv_imp", 8)}


class SlabWorld:
    """One rank's slab: an odeb200.World plus device-side halo buffers.  The transfer itself is supplied by the
    caller (`exchange_nccl` through torch.distributed, or `exchange_local` between emulated slabs)."""

    def __init__(self, world, halo, device):
        import torch
        self.w = world
        self.torch = torch
        self.dev = device
        self.sides = {}
        for side in ("left", "right"):
            if halo[side] is None:
                continue
            s = {}
            for kind, (ks, kr, width) in _KINDS.items():
                for key, role in ((ks, "send"), (kr, "recv")):
                    idx = halo[side][key]
                    if idx is None:
                        continue
                    assert len(idx) > 0
                    s["%s_%s_idx" % (role, kind)] = torch.as_tensor(idx, dtype=torch.int32, device=device)
                    s["%s_%s_buf" % (role, kind)] = torch.empty((len(idx), width), dtype=torch.float32, device=device)
            self.sides[side] = s
        self.has_imp = any("send_imp_buf" in s or "recv_imp_buf" in s for s in self.sides.values())

    def pack(self, kind="state"):
        """gather into the send buffers (engine stream), then wait for them"""
        L = self.w.L
        fn = L.dWorldPackStatesDeviceB200 if kind == "state" else L.dWorldPackImpulsesDeviceB200
        for s in self.sides.values():
            if "send_%s_buf" % kind in s:
                idx, buf = s["send_%s_idx" % kind], s["send_%s_buf" % kind]
                fn(self.w.w, idx.data_ptr(), idx.numel(), buf.data_ptr())
        self.w.wait()

    def unpack(self, kind="state"):
        """scatter received states into the ghosts / add received impulses to the owned boundary bodies
        (the caller has synchronised the transfer)"""
        L = self.w.L
        fn = L.dWorldUnpackStatesDeviceB200 if kind == "state" else L.dWorldAddImpulsesDeviceB200
        for s in self.sides.values():
            if "recv_%s_buf" % kind in s:
                idx, buf = s["recv_%s_idx" % kind], s["recv_%s_buf" % kind]
                fn(self.w.w, idx.data_ptr(), idx.numel(), buf.data_ptr())

    def halo_bytes(self):
        return sum(v.numel() * 4 for s in self.sides.values() for k, v in s.items() if k.startswith("send_") and k.endswith("_buf"))


def exchange_nccl(slab, rank, n_slabs, kind="state"):
    """halo exchange with the +-1 neighbours: at most one send and one recv per face, batched (NCCL group)."""
    import torch.distributed as dist
    ops = []
    for side, nbr in (("left", rank - 1), ("right", rank + 1)):
        s = slab.sides.get(side)
        if s is None:
            continue
        if "send_%s_buf" % kind in s:
            ops.append(dist.P2POp(dist.isend, s["send_%s_buf" % kind], nbr))
        if "recv_%s_buf" % kind in s:
            ops.append(dist.P2POp(dist.irecv, s["recv_%s_buf" % kind], nbr))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    slab.torch.cuda.synchronize()


def exchange_local(slabs, kind="state"):
    """several slabs emulated in one process on one GPU: the 'link' is a device-to-device copy"""
    sk, rk = "send_%s_buf" % kind, "recv_%s_buf" % kind
    for r, s in enumerate(slabs):
        if "right" not in s.sides:
            continue
        a, b = s.sides["right"], slabs[r + 1].sides["left"]
        if sk in a:
            b[rk].copy_(a[sk])
        if sk in b:
            a[rk].copy_(b[sk])
    slabs[0].torch.cuda.synchronize()


def tick(slab, exchange, h, max_contacts=8):
    """one tick of a slab: state halo, the reference tick (collide -> step), then the impulse halo.
    `exchange(kind)` moves the packed buffers of that kind."""
    slab.pack("state")
    exchange("state")
    slab.unpack("state")
    slab.w.tick(h, max_contacts)
    if slab.has_imp:
        slab.pack("imp")
        exchange("imp")
        slab.unpack("imp")


# ------------------------------------------------------------------ dynamic halo + migration
#
# The lists above are fixed at set-up, which is only right while bodies stay in their lattice columns.  The
# dynamic variant rebuilds everything from the bodies' CURRENT positions (impulse coupling only):
#   * every tick rank r selects, on the device, its own bodies with x < face_left + margin and sends them whole
#     (state + mass properties + geom shape, 48 floats) to rank r-1, which unpacks them into a POOL of ghost
#     slots (unused slots are switched off); impulses return along the same list;
#   * every `migrate_every` ticks bodies whose centre crossed a face by more than `hyst` change owner: the old
#     owner selects and packs them on the device, the records travel like a halo message, the new owner spawns
#     them through the handle API (queued patches: no re-upload of the world) and the old owner destroys them.

REC = 48  # floats per body record (dWorldPackBodiesDeviceB200)


def dynamic_slab_scene(rank, n_slabs, nx_per_slab=128, nz=1024, ny=16, seed=5, spacing=1.8, margin_cols=4, h=1.0 / 60.0,
                       pool_factor=1.5, mig_cap=4096):
    """One rank's scene for the dynamic halo: its own lattice slab + (unless it is the last rank) a pool of ghost
    slots for the right neighbour's boundary bodies.  Returns (scene, info)."""
    nx_total = nx_per_slab * n_slabs
    x0 = -0.5 * (nx_total - 1) * spacing
    z0 = -0.5 * (nz - 1) * spacing
    own_b, own_g, _ = _slab_lattice(rank, nx_per_slab, nz, ny, seed, spacing, x0, z0)
    n_own = len(own_b["pos"])
    own_g["cat"][:] = CAT_OBJ
    own_g["col"][:] = CAT_OBJ | CAT_MAP | CAT_GHOST
    pool = int(pool_factor * margin_cols * nz * ny) + 64
    bparts, gparts = [own_b], [own_g]
    has_pool = rank < n_slabs - 1
    if has_pool:
        pb = {
            "pos": np.stack([np.zeros(pool), -1000.0 - np.arange(pool), np.zeros(pool)], axis=1).astype(np.float32),
            "quat": np.tile(np.array([1, 0, 0, 0], np.float32), (pool, 1)),
            "lvel": np.zeros((pool, 3), np.float32), "avel": np.zeros((pool, 3), np.float32),
            "mass": np.ones(pool, np.float32),
            "inertia": np.tile(np.array([1, 0, 0, 0, 1, 0, 0, 0, 1], np.float32), (pool, 1)),
            "flags": np.full(pool, scenes.BODY_KINEMATIC, np.int32), "env": np.zeros(pool, np.int32),
        }
        pg = {
            "type": np.full(pool, scenes.SPHERE, np.int32), "dims": np.tile(np.float32([0.1, 0, 0, 0]), (pool, 1)),
            "body": np.arange(n_own, n_own + pool, dtype=np.int32), "pos": np.zeros((pool, 3), np.float32),
            "R": np.tile(np.array(scenes.IDENT_R, np.float32), (pool, 1)),
            "cat": np.full(pool, CAT_GHOST, np.uint32), "col": np.zeros(pool, np.uint32), "env": np.zeros(pool, np.int32),
        }
        bparts.append(pb)
        gparts.append(pg)
    half_x = 0.5 * nx_total * spacing + 1.0
    half_z = 0.5 * nz * spacing + 1.0
    st = scenes._static_geoms([(scenes.PLANE, (0, 1, 0, 0.0), (0, 0, 0), scenes.IDENT_R, -1),
                               (scenes.PLANE, (1, 0, 0, -half_x), (0, 0, 0), scenes.IDENT_R, -1),
                               (scenes.PLANE, (-1, 0, 0, -half_x), (0, 0, 0), scenes.IDENT_R, -1),
                               (scenes.PLANE, (0, 0, 1, -half_z), (0, 0, 0), scenes.IDENT_R, -1),
                               (scenes.PLANE, (0, 0, -1, -half_z), (0, 0, 0), scenes.IDENT_R, -1)])
    st["cat"][:] = CAT_MAP
    st["col"][:] = CAT_OBJ
    sc = scenes.from_arrays("C5-dyn-slab%d" % rank, scenes._concat(bparts), scenes._concat([st] + gparts), h=h)
    sc["n_owned"] = n_own
    face_left = x0 + (rank * nx_per_slab - 0.5) * spacing
    info = {
        "rank": rank, "n_slabs": n_slabs, "n_own": n_own, "n_static": 5, "pool": pool, "has_pool": has_pool,
        "pool_first_body": n_own, "pool_first_geom": 5 + n_own,
        "face_left": face_left, "face_right": face_left + nx_per_slab * spacing,
        "margin": margin_cols * spacing, "hyst": 0.25 * margin_cols * spacing, "mig_cap": mig_cap,
    }
    return sc, info


class DynamicSlabWorld:
    """Slab with a device-selected halo and host-driven migration.  Buffer names follow SlabWorld, so
    exchange_nccl / exchange_local move kinds "state", "imp" and "mig" unchanged."""

    def __init__(self, world, info, device):
        import torch
        self.w, self.torch, self.dev, self.info = world, torch, device, info
        self.rank, self.n_slabs = info["rank"], info["n_slabs"]
        n_bodies = world.n_bodies
        cap = n_bodies + 16 * info["mig_cap"]
        self.own_mask = torch.zeros(cap, dtype=torch.int32, device=device)
        self.own_mask[:info["n_own"]] = 1
        self.body_geom = torch.full((cap,), -1, dtype=torch.int32, device=device)
        self.body_geom[:n_bodies] = torch.arange(info["n_static"], info["n_static"] + n_bodies, dtype=torch.int32, device=device)
        self.n_owned = info["n_own"]
        self.count = torch.zeros(4, dtype=torch.int32, device=device)
        world.L.dWorldSetSlotReuseB200(world.w, 1)   # migrants take the slots of bodies that left (as the C driver does)
        P, M = info["pool"], info["mig_cap"]
        f32 = dict(dtype=torch.float32, device=device)
        i32 = dict(dtype=torch.int32, device=device)
        self.sides = {}
        if self.rank > 0:
            self.sides["left"] = {
                "send_state_idx": torch.full((P,), -1, **i32), "send_state_buf": torch.zeros((P, REC), **f32),
                "recv_imp_buf": torch.zeros((P, 8), **f32),
                "send_mig_idx": torch.full((M,), -1, **i32), "send_mig_buf": torch.zeros((M, REC), **f32),
                "recv_mig_buf": torch.zeros((M, REC), **f32),
            }
        if self.rank < self.n_slabs - 1:
            self.sides["right"] = {
                "ghost_body": torch.arange(info["pool_first_body"], info["pool_first_body"] + P, **i32),
                "ghost_geom": torch.arange(info["pool_first_geom"], info["pool_first_geom"] + P, **i32),
                "recv_state_buf": torch.zeros((P, REC), **f32), "send_imp_buf": torch.zeros((P, 8), **f32),
                "send_mig_idx": torch.full((M,), -1, **i32), "send_mig_buf": torch.zeros((M, REC), **f32),
                "recv_mig_buf": torch.zeros((M, REC), **f32),
            }
        self.has_imp = True
        self.migrated_in = self.migrated_out = 0

    # -- per tick
    def pack(self, kind="state"):
        L, w = self.w.L, self.w.w
        inf = float("inf")
        if kind == "state" and "left" in self.sides:
            s = self.sides["left"]
            L.dWorldSelectBodiesDeviceB200(w, 0, -inf, self.info["face_left"] + self.info["margin"], self.own_mask.data_ptr(),
                                           s["send_state_idx"].data_ptr(), s["send_state_idx"].numel(), self.count.data_ptr())
            L.dWorldPackBodiesDeviceB200(w, s["send_state_idx"].data_ptr(), s["send_state_idx"].numel(), self.body_geom.data_ptr(),
                                         s["send_state_buf"].data_ptr())
        if kind == "imp" and "right" in self.sides:
            s = self.sides["right"]
            L.dWorldPackImpulsesDeviceB200(w, s["ghost_body"].data_ptr(), s["ghost_body"].numel(), s["send_imp_buf"].data_ptr())
        if kind == "mig":
            for side, s in self.sides.items():
                lo, hi = (-inf, self.info["face_left"] - self.info["hyst"]) if side == "left" else (self.info["face_right"] + self.info["hyst"], inf)
                L.dWorldSelectBodiesDeviceB200(w, 0, lo, hi, self.own_mask.data_ptr(), s["send_mig_idx"].data_ptr(),
                                               s["send_mig_idx"].numel(), self.count.data_ptr() + (4 if side == "left" else 8))
                L.dWorldPackBodiesDeviceB200(w, s["send_mig_idx"].data_ptr(), s["send_mig_idx"].numel(), self.body_geom.data_ptr(),
                                             s["send_mig_buf"].data_ptr())
        self.w.wait()
        # overflow is reported, never silent: the selection kernels count every body in range, also those beyond capacity
        if kind in ("state", "mig"):
            cnt = self.count.cpu().numpy()
            if kind == "state" and "left" in self.sides and cnt[0] > self.sides["left"]["send_state_idx"].numel():
                raise RuntimeError("DynamicSlabWorld: %d bodies within the face margin, the ghost pool holds %d"
                                   % (cnt[0], self.sides["left"]["send_state_idx"].numel()))
            if kind == "mig":
                for side, k in (("left", 1), ("right", 2)):
                    if side in self.sides and cnt[k] > self.sides[side]["send_mig_idx"].numel():
                        import warnings
                        warnings.warn("DynamicSlabWorld: %d bodies crossed the %s face, the migration buffer holds %d; the rest "
                                      "change owner at the next migration" % (cnt[k], side, self.sides[side]["send_mig_idx"].numel()))

    def unpack(self, kind="state"):
        L, w = self.w.L, self.w.w
        if kind == "state" and "right" in self.sides:
            s = self.sides["right"]
            L.dWorldUnpackBodiesDeviceB200(w, s["ghost_body"].data_ptr(), s["ghost_geom"].data_ptr(), s["ghost_body"].numel(),
                                           s["recv_state_buf"].data_ptr())
        if kind == "imp" and "left" in self.sides:
            s = self.sides["left"]
            L.dWorldAddImpulsesDeviceB200(w, s["send_state_idx"].data_ptr(), s["send_state_idx"].numel(), s["recv_imp_buf"].data_ptr())
        if kind == "mig":
            self._apply_migration()

    # -- migration (host side; the exchange of kind "mig" has completed)
    def _apply_migration(self):
        import ctypes as C
        from . import Mass
        L, w, space = self.w.L, self.w.w, self.w.space
        for side, s in self.sides.items():
            # leaving: destroy what we sent
            out_idx = s["send_mig_idx"].cpu().numpy()
            out_idx = out_idx[out_idx >= 0]
            if len(out_idx):
                geoms = self.body_geom[self.torch.as_tensor(out_idx, dtype=self.torch.long, device=self.dev)].cpu().numpy()
                for b, g in zip(out_idx.tolist(), geoms.tolist()):
                    L.dGeomDestroy(C.c_void_p(L.dSpaceGetGeomB200(space, int(g))))
                    L.dBodyDestroy(C.c_void_p(L.dWorldGetBodyB200(w, int(b))))
                self.own_mask[self.torch.as_tensor(out_idx, dtype=self.torch.long, device=self.dev)] = 0
                self.n_owned -= len(out_idx)
                self.migrated_out += len(out_idx)
            # arriving: spawn through the handle API (queued as patches, sent with the next collide)
            rec = s["recv_mig_buf"].cpu().numpy()
            types = rec[:, 15].view(np.int32)
            new_b, new_g = [], []
            for r in rec[types >= 0]:
                t = int(r[15:16].view(np.int32)[0])
                b = C.c_void_p(L.dBodyCreate(w))
                L.dBodySetPosition(b, float(r[0]), float(r[1]), float(r[2]))
                q = np.ascontiguousarray(r[4:8], np.float32)
                L.dBodySetQuaternion(b, q.ctypes.data_as(C.POINTER(C.c_float)))
                L.dBodySetLinearVel(b, float(r[8]), float(r[9]), float(r[10]))
                L.dBodySetAngularVel(b, float(r[12]), float(r[13]), float(r[14]))
                m = Mass()
                m.mass = float(r[11])
                for k in range(12):
                    m.I[k] = float(r[20 + k])
                L.dBodySetMass(b, C.byref(m))
                flags = int(r[44:45].view(np.int32)[0])
                L.dBodySetGyroscopicMode(b, 1 if flags & scenes.BODY_GYRO else 0)
                if t == scenes.SPHERE:
                    g = C.c_void_p(L.dCreateSphere(space, float(r[16])))
                else:
                    g = C.c_void_p(L.dCreateBox(space, float(r[16]), float(r[17]), float(r[18])))
                L.dGeomSetBody(g, b)
                L.dGeomSetCategoryBits(g, CAT_OBJ)
                L.dGeomSetCollideBits(g, CAT_OBJ | CAT_MAP | CAT_GHOST)
                new_b.append(L.dBodyGetIndexB200(b))
                new_g.append(L.dGeomGetIndexB200(g))
            if new_b:
                if max(new_b) >= self.own_mask.numel():
                    raise RuntimeError("DynamicSlabWorld: body capacity exhausted by migration")
                ib = self.torch.as_tensor(new_b, dtype=self.torch.long, device=self.dev)
                self.own_mask[ib] = 1
                self.body_geom[ib] = self.torch.as_tensor(new_g, dtype=self.torch.int32, device=self.dev)
                self.w.n_bodies += len(new_b)
                self.w.n_geoms += len(new_b)
                self.n_owned += len(new_b)
                self.migrated_in += len(new_b)

    def halo_bytes(self):
        n = 0
        for s in self.sides.values():
            for k in ("send_state_buf", "send_imp_buf"):
                if k in s:
                    n += s[k].numel() * 4
        return n


def tick_dynamic(slab, exchange, h, step, migrate_every=16, max_contacts=8):
    """one tick of a dynamic slab; every `migrate_every` ticks ownership follows the bodies first"""
    if migrate_every > 0 and step % migrate_every == 0 and step > 0:
        slab.pack("mig")
        exchange("mig")
        slab.unpack("mig")
    tick(slab, exchange, h, max_contacts)


# ------------------------------------------------------------------ the C driver (csrc/slab.cu)
#
# Everything above orchestrates the halo from Python (torch tensors as buffers, torch.distributed as transport) and
# is kept as the reference the C driver is tested against.  The product path is the C driver: dSlabCreateB200 /
# dSlabTickB200 / dSlabMigrateB200 behind include/ode_b200.h -- buffers, NCCL calls and the event ordering live inside
# libode_b200.so, so a C application needs nothing but an NCCL unique id from its launcher.

class CSlab:
    """One rank's slab driven by the library (dSlab*B200).  `nccl_id`: 128 bytes from dSlabGetUniqueIdB200 of rank 0
    (None: same-process slabs, connect them with CSlab.connect and tick them with CSlab.tick_local)."""

    def __init__(self, world, info, nccl_id=None):
        from . import SlabLayout
        import ctypes as C
        self.w, self.info = world, info
        lay = SlabLayout(face_left=info["face_left"], face_right=info["face_right"], margin=info["margin"], hyst=info["hyst"],
                         n_own=info["n_own"], n_static=info["n_static"], pool=info["pool"],
                         pool_first_body=info["pool_first_body"], pool_first_geom=info["pool_first_geom"], mig_cap=info["mig_cap"])
        world.L.dSpaceCollideDeviceB200  # (the library must be loaded)
        world.wait()
        self.h = C.c_void_p(world.L.dSlabCreateB200(world.w, world.space, info["rank"], info["n_slabs"], nccl_id, C.byref(lay)))

    def tick(self, h, max_contacts=8):
        self.w.L.dSlabTickB200(self.h, float(h), int(max_contacts))

    def migrate(self):
        self.w.L.dSlabMigrateB200(self.h)

    def get_info(self):
        from . import SlabInfo
        import ctypes as C
        i = SlabInfo()
        self.w.L.dSlabGetInfoB200(self.h, C.byref(i))
        return {k: getattr(i, k) for k, _ in SlabInfo._fields_}

    def halo_bytes(self):
        return self.get_info()["halo_bytes_per_tick"]

    @property
    def migrated_out(self):
        return self.get_info()["migrated_out"]

    def close(self):
        if self.h:
            self.w.L.dSlabDestroyB200(self.h)
            self.h = None

    @staticmethod
    def connect(lower, upper):
        lower.w.L.dSlabConnectLocalB200(lower.h, upper.h)

    @staticmethod
    def _array(slabs):
        import ctypes as C
        return (C.c_void_p * len(slabs))(*[s.h for s in slabs])

    @staticmethod
    def tick_local(slabs, h, max_contacts=8):
        slabs[0].w.L.dSlabTickLocalB200(CSlab._array(slabs), len(slabs), float(h), int(max_contacts))

    @staticmethod
    def migrate_local(slabs):
        slabs[0].w.L.dSlabMigrateLocalB200(CSlab._array(slabs), len(slabs))


def nccl_unique_id(lib, rank, world):
    """128-byte NCCL unique id created by rank 0's library and broadcast to the others (here through torch.distributed,
    which the benchmark already uses for its barrier; a C application would use its own launcher's channel)."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    buf = C.create_string_buffer(128)
    if rank == 0 and not lib.dSlabGetUniqueIdB200(buf):
        raise RuntimeError("NCCL is not available to libode_b200.so (libnccl.so.2 not found)")
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone().cuda()
    if world > 1:
        dist.broadcast(t, 0)
    return bytes(t.cpu().numpy().tobytes())
