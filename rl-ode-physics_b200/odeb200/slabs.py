"""Slab decomposition of one large world across GPUs (BASELINE config 5).

The lattice pile is cut along x into one slab per rank.  A rank owns the bodies of its slab; the
`margin_cols` lattice columns of each neighbour nearest the shared face are mirrored as kinematic GHOST
bodies.  Every tick, before collide, the owners' current states of those boundary bodies are exchanged
(one message per face, device to device: NCCL send/recv over NVLink, or a plain device copy when several
slabs are emulated on one GPU) and scattered into the ghosts.  Contacts between an owned body and a
ghost push only the owned body -- the neighbour does the mirror-image computation for its own body -- so
the only collective on the path is this halo exchange of boundary-body states.

Round-1 limits: boundary membership is fixed by the initial lattice column (no
migration of bodies between slabs), and contact impulses are not exchanged (one-sided coupling).
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


def slab_scene(rank, n_slabs, nx_per_slab=128, nz=1024, ny=16, seed=5, spacing=1.8, margin_cols=4, h=1.0 / 60.0):
    """Scene of one rank + its halo lists.

    Returns (scene, halo) with halo = {"left": (send_idx, recv_idx) | None, "right": ...}: send_idx are the
    rank's own boundary bodies (ascending), recv_idx the ghost bodies mirroring the neighbour's boundary
    bodies, in the neighbour's send order."""
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
        nb, ng, nix = _slab_lattice(nbr, nx_per_slab, nz, ny, seed, spacing, x0, z0)
        # neighbour's columns facing us / our columns facing the neighbour
        sel_n = np.nonzero(nix >= nx_per_slab - margin_cols)[0] if side == "left" else np.nonzero(nix < margin_cols)[0]
        sel_o = np.nonzero(ix < margin_cols)[0] if side == "left" else np.nonzero(ix >= nx_per_slab - margin_cols)[0]
        gb = {k: v[sel_n].copy() for k, v in nb.items()}
        gg = {k: v[sel_n].copy() for k, v in ng.items()}
        gb["flags"][:] = scenes.BODY_KINEMATIC
        gg["body"] = np.arange(n_bodies, n_bodies + len(sel_n), dtype=np.int32)
        gg["cat"][:] = CAT_GHOST
        gg["col"][:] = 0
        halo[side] = (sel_o.astype(np.int32), gg["body"].copy())
        n_bodies += len(sel_n)
        bparts.append(gb)
        gparts.append(gg)
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
    return sc, halo


class SlabWorld:
    """One rank's slab: an odeb200.World plus device-side halo buffers.  `exchange(sendbufs) -> recvbufs`
    is supplied by the caller (NCCL through torch.distributed, or a local copy between emulated slabs)."""

    def __init__(self, world, halo, device):
        import torch
        self.w = world
        self.torch = torch
        self.dev = device
        self.sides = {}
        for side in ("left", "right"):
            if halo[side] is None:
                continue
            send_idx, recv_idx = halo[side]
            assert len(send_idx) > 0 and len(recv_idx) > 0
            self.sides[side] = {
                "send_idx": torch.as_tensor(send_idx, dtype=torch.int32, device=device),
                "recv_idx": torch.as_tensor(recv_idx, dtype=torch.int32, device=device),
                "send_buf": torch.empty((len(send_idx), 16), dtype=torch.float32, device=device),
                "recv_buf": torch.empty((len(recv_idx), 16), dtype=torch.float32, device=device),
            }

    def pack(self):
        """gather the boundary-body states into the send buffers (engine stream), then wait for them"""
        L = self.w.L
        for s in self.sides.values():
            L.dWorldPackStatesDeviceB200(self.w.w, s["send_idx"].data_ptr(), s["send_idx"].numel(), s["send_buf"].data_ptr())
        self.w.wait()

    def unpack(self):
        """scatter received states into the ghost bodies (caller has synchronised the transfer)"""
        L = self.w.L
        for s in self.sides.values():
            L.dWorldUnpackStatesDeviceB200(self.w.w, s["recv_idx"].data_ptr(), s["recv_idx"].numel(), s["recv_buf"].data_ptr())

    def halo_bytes(self):
        return sum(s["send_buf"].numel() * 4 for s in self.sides.values())


def exchange_nccl(slab, rank, n_slabs):
    """halo exchange with the +-1 neighbours: one send and one recv per face, batched (NCCL group)."""
    import torch.distributed as dist
    ops = []
    for side, nbr in (("left", rank - 1), ("right", rank + 1)):
        if side in slab.sides:
            s = slab.sides[side]
            ops.append(dist.P2POp(dist.isend, s["send_buf"], nbr))
            ops.append(dist.P2POp(dist.irecv, s["recv_buf"], nbr))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    slab.torch.cuda.synchronize()


def exchange_local(slabs):
    """several slabs emulated in one process on one GPU: the 'link' is a device-to-device copy"""
    for r, s in enumerate(slabs):
        if "right" in s.sides:
            slabs[r + 1].sides["left"]["recv_buf"].copy_(s.sides["right"]["send_buf"])
            s.sides["right"]["recv_buf"].copy_(slabs[r + 1].sides["left"]["send_buf"])
    slabs[0].torch.cuda.synchronize()


def tick(slab, exchange, h, max_contacts=8):
    """one tick of a slab: halo exchange, then the reference tick (collide -> step)"""
    slab.pack()
    exchange()
    slab.unpack()
    slab.w.tick(h, max_contacts)
