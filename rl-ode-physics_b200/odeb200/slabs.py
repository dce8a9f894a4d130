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
Boundary membership is fixed by the initial lattice column: bodies do not migrate between slabs (DESIGN.md
section 6).
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
