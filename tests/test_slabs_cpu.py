"""CPU: host logic of the slab decomposition (BASELINE config 5): ownership, ghost mirrors and halo lists
of neighbouring ranks line up."""
import numpy as np

from odeb200 import scenes, slabs


def test_neighbouring_slabs_agree_on_the_halo():
    n_slabs, nx, nz, ny, mc = 3, 6, 5, 3, 2
    built = [slabs.slab_scene(r, n_slabs, nx_per_slab=nx, nz=nz, ny=ny, seed=5, margin_cols=mc, coupling="ghost")
             for r in range(n_slabs)]
    for r, (sc, halo) in enumerate(built):
        n_own = sc["n_owned"]
        assert n_own == nx * nz * ny
        b, g = sc["bodies"], sc["geoms"]
        assert (b["flags"][:n_own] == 0).all() and (b["flags"][n_own:] == scenes.BODY_KINEMATIC).all()
        assert (halo["left"] is None) == (r == 0) and (halo["right"] is None) == (r == n_slabs - 1)
        # geoms: 5 planes, then one geom per body in body order
        assert np.array_equal(g["body"][5:], np.arange(len(b["pos"])))
        assert (g["cat"][5:5 + n_own] == slabs.CAT_OBJ).all() and (g["cat"][5 + n_own:] == slabs.CAT_GHOST).all()
        if halo["right"] is not None:
            send = halo["right"]["send_state"]
            recv = built[r + 1][1]["left"]["recv_state"]
            nb, ng = built[r + 1][0]["bodies"], built[r + 1][0]["geoms"]
            assert len(send) == len(recv) == mc * nz * ny
            # the neighbour's ghosts mirror exactly the bodies we send, in the same order
            assert np.array_equal(b["pos"][send], nb["pos"][recv])
            assert np.array_equal(g["dims"][5 + send], ng["dims"][5 + recv])
            assert np.array_equal(g["type"][5 + send], ng["type"][5 + recv])
            # and they are the columns nearest the shared face
            face = 0.5 * (b["pos"][:n_own, 0].max() + nb["pos"][:built[r + 1][0]["n_owned"], 0].min())
            assert (np.abs(b["pos"][send, 0] - face) < (mc + 0.5) * 1.8).all()
        if halo["left"] is not None:
            send = halo["left"]["send_state"]
            recv = built[r - 1][1]["right"]["recv_state"]
            assert np.array_equal(b["pos"][send], built[r - 1][0]["bodies"]["pos"][recv])
    # slabs tile x without gaps: owned x ranges are disjoint and ordered
    xs = [(sc["bodies"]["pos"][:sc["n_owned"], 0].min(), sc["bodies"]["pos"][:sc["n_owned"], 0].max()) for sc, _ in built]
    assert xs[0][1] < xs[1][0] and xs[1][1] < xs[2][0]


def test_impulse_coupling_lower_slab_owns_the_face():
    """a cross-slab contact is owned by the lower slab.  Rank r mirrors only rank r+1's
    boundary columns, as dynamic bodies with the owners' masses; states travel down, impulses travel up, and
    the index lists of the two ends of every message line up."""
    n_slabs, nx, nz, ny, mc = 3, 6, 5, 3, 2
    built = [slabs.slab_scene(r, n_slabs, nx_per_slab=nx, nz=nz, ny=ny, seed=5, margin_cols=mc) for r in range(n_slabs)]
    for r, (sc, halo) in enumerate(built):
        n_own = sc["n_owned"]
        b, g = sc["bodies"], sc["geoms"]
        n_ghost = len(b["pos"]) - n_own
        assert n_ghost == (mc * nz * ny if r < n_slabs - 1 else 0)
        assert (b["flags"] == 0).all()                              # ghosts are dynamic here
        assert (g["cat"][5 + n_own:] == slabs.CAT_GHOST).all() and (g["col"][5 + n_own:] == 0).all()
        if r > 0:
            left, up = halo["left"], built[r - 1][1]["right"]
            assert left["recv_state"] is None and left["send_imp"] is None
            assert np.array_equal(left["send_state"], left["recv_imp"])
            assert np.array_equal(up["recv_state"], up["send_imp"])
            assert up["send_state"] is None and up["recv_imp"] is None
            lower = built[r - 1][0]["bodies"]
            # the lower rank's ghosts mirror exactly the bodies we send: same pose, same mass, same shape
            assert np.array_equal(b["pos"][left["send_state"]], lower["pos"][up["recv_state"]])
            assert np.array_equal(b["mass"][left["send_state"]], lower["mass"][up["recv_state"]])
            assert np.array_equal(g["dims"][5 + left["send_state"]], built[r - 1][0]["geoms"]["dims"][5 + up["recv_state"]])
    assert built[0][1]["left"] is None and built[-1][1]["right"] is None


def test_ghosts_only_collide_with_owned_bodies():
    sc, halo = slabs.slab_scene(0, 2, nx_per_slab=4, nz=4, ny=2, margin_cols=2)
    g = sc["geoms"]

    def passes(i, j):
        return bool((g["cat"][i] & g["col"][j]) or (g["cat"][j] & g["col"][i]))
    n_own = sc["n_owned"]
    plane, owned, ghost = 0, 5, 5 + n_own
    assert passes(owned, owned + 1) and passes(owned, ghost) and passes(owned, plane)
    assert not passes(ghost, ghost + 1) and not passes(ghost, plane) and not passes(plane, plane + 1)


def test_dynamic_slab_scene_layout():
    """dynamic halo: every rank but the last carries a pool of switched-off ghost slots behind its own bodies;
    pools have one size (a message fills the receiver's pool slot for slot); faces tile the x axis; geom index =
    5 static planes + body index (the body->geom map DynamicSlabWorld starts from)."""
    n_slabs = 3
    built = [slabs.dynamic_slab_scene(r, n_slabs, nx_per_slab=6, nz=5, ny=3, margin_cols=2, spacing=1.8) for r in range(n_slabs)]
    pools = {info["pool"] for _, info in built}
    assert len(pools) == 1
    for r, (sc, info) in enumerate(built):
        b, g = sc["bodies"], sc["geoms"]
        n_own, pool = info["n_own"], info["pool"]
        assert info["has_pool"] == (r < n_slabs - 1)
        assert len(b["pos"]) == n_own + (pool if info["has_pool"] else 0)
        assert np.array_equal(g["body"][5:], np.arange(len(b["pos"])))
        assert (b["flags"][:n_own] == 0).all() and (b["flags"][n_own:] == scenes.BODY_KINEMATIC).all()
        assert (g["cat"][5 + n_own:] == slabs.CAT_GHOST).all() and (g["col"][5 + n_own:] == 0).all()
        assert (b["pos"][n_own:, 1] < -100).all()                       # parked far below the ground plane
        xs = b["pos"][:n_own, 0]
        assert info["face_left"] < xs.min() and xs.max() < info["face_right"]
        if r > 0:
            assert abs(info["face_left"] - built[r - 1][1]["face_right"]) < 1e-9
        assert 0 < info["hyst"] < info["margin"]


class _HostSlab:
    """stand-in for SlabWorld / DynamicSlabWorld on the CPU: the same `sides` buffer naming, no engine behind it"""

    def __init__(self, rank, n_slabs, sizes):
        import torch
        self.torch = torch
        self.sides = {}
        mk = lambda n, w, v: torch.full((n, w), float(v))   # noqa: E731
        if rank > 0:        # impulse coupling: states go down (to rank-1), impulses come up from it; migrants both ways
            self.sides["left"] = {"send_state_buf": mk(sizes["state"], 4, 100 + rank), "recv_imp_buf": mk(sizes["imp"], 2, -1),
                                  "send_mig_buf": mk(sizes["mig"], 3, 300 + rank), "recv_mig_buf": mk(sizes["mig"], 3, -1)}
        if rank < n_slabs - 1:
            self.sides["right"] = {"recv_state_buf": mk(sizes["state"], 4, -1), "send_imp_buf": mk(sizes["imp"], 2, 200 + rank),
                                   "send_mig_buf": mk(sizes["mig"], 3, 400 + rank), "recv_mig_buf": mk(sizes["mig"], 3, -1)}


def _halo_worker(rank, world_size, port, q):
    import os
    import torch.distributed as dist
    from odeb200 import sharding
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world_size), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    sharding.init_process_group("gloo")
    slab = _HostSlab(rank, world_size, {"state": 5, "imp": 5, "mig": 3})
    slab.torch.cuda.synchronize = lambda: None            # exchange_nccl ends with a device sync; nothing to sync here
    for kind in ("state", "imp", "mig"):
        slabs.exchange_nccl(slab, rank, world_size, kind)
    out = {side: {k: float(v[0, 0]) for k, v in s.items() if k.startswith("recv_")} for side, s in slab.sides.items()}
    q.put((rank, out))
    dist.destroy_process_group()


def test_halo_messages_route_between_three_ranks_gloo():
    """the exchange the C5 bench runs over NCCL, on gloo with CPU tensors and world_size 3: states travel to the lower
    neighbour, impulses back up, migrants both ways, and the middle rank talks to both sides in one batch"""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ws = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_halo_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(ws))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(ws):
        if r < ws - 1:      # from the right neighbour: its states and its left-going migrants
            assert res[r]["right"]["recv_state_buf"] == 100 + (r + 1)
            assert res[r]["right"]["recv_mig_buf"] == 300 + (r + 1)
        if r > 0:           # from the left neighbour: impulses for our boundary bodies and its right-going migrants
            assert res[r]["left"]["recv_imp_buf"] == 200 + (r - 1)
            assert res[r]["left"]["recv_mig_buf"] == 400 + (r - 1)
