#!/usr/bin/env python
"""bench.py -- body-steps/s of the reference tick (collide -> QuickStep(20) -> snapshot pack) on B200.

    python bench.py --gpus N --steps K --warmup W [--workload C4|C3|C2|C1] [--impl reference]

A "step" is one full tick of the hot path (dSpaceCollideDeviceB200 + dWorldQuickStep with the fused
integrate + snapshot pack) over the whole workload.  Default workload = BASELINE.json config 4, the one
its metric is quoted on at 1/2/4/8 B200: 8192 independent 128-body worlds PER GPU (weak scaling, no
data-path collective).  Prints ONE JSON line (see the contract in the task statement / DESIGN.md).

`--impl reference` times the CPU oracle port (libode is not installable: it is an un-vendored,
un-versioned dependency of the reference) on all host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "rl-ode-physics_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "body_steps_per_sec"
UNIT = "body-steps/s"
SETTLE = {"C4": 100, "C3": 300, "C2": 120, "C1": 200, "C5": 300}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="C4", choices=["C4", "C3", "C2", "C1", "C5"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--worlds-per-gpu", type=int, default=8192)
    ap.add_argument("--batches", type=int, default=4,
                    help="C4: the worlds of one GPU are held in this many dWorld objects (each has its own CUDA stream, so the "
                         "collide kernels of one batch overlap the solver of another); 1 = one dWorld for all of them")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary blocks (C3 roofline at N=1; strong scaling and C5 at N>1)")
    ap.add_argument("--snapshot-format", type=int, default=2, choices=[0, 1, 2],
                    help="record format of the end-to-end leg's per-tick snapshot copy: 0 = 16 floats (GetTransformMat), "
                         "1 = its 12 non-constant floats, 2 = position + quaternion (8 floats)")
    ap.add_argument("--settle", type=int, default=-1)
    ap.add_argument("--slab-cols", type=int, default=128, help="C5: lattice columns (x) per GPU; z = 1024, y = 16")
    ap.add_argument("--halo", default="dynamic", choices=["dynamic", "static"],
                    help="C5: dynamic = boundary set re-selected on the device each tick + migration between slabs; "
                         "static = boundary set fixed by the initial lattice column")
    ap.add_argument("--migrate-every", type=int, default=16, help="C5 dynamic halo: ticks between ownership updates")
    ap.add_argument("--coupling", default="impulse", choices=["impulse", "ghost"],
                    help="C5: impulse = lower slab owns cross-face contacts, impulses sent back to the owner; "
                         "ghost = kinematic ghosts on both sides")
    return ap.parse_args()


def build_scene(workload, rank, worlds_per_gpu):
    from odeb200 import scenes
    if workload == "C4":
        first = rank * worlds_per_gpu
        sc = scenes.batched_worlds_scene(worlds_per_gpu, seed=4, first_world=first)
        desc = C4_DESC % worlds_per_gpu
    elif workload == "C3":
        sc = scenes.pile_scene(256, 256, 16, seed=3)
        desc = "C3: 1,048,576-body random box/sphere pile on a plane + 4 wall planes, dt=1/60, QuickStep 20 iters"
    elif workload == "C2":
        sc = scenes.trimesh_scene(100, seed=2)
        desc = "C2: teapot.obj trimesh (8884 tris) vs 10,000 spheres, dt=1/60, QuickStep 20 iters"
    else:
        sc = scenes.server_scene(seed=1)
        desc = "C1: reference server scene, 4 static boxes + 64 dropped boxes/spheres + 4 kinematic spheres, dt=1/60"
    return sc, desc


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """samples before this point (warm-up) are dropped"""
        self.first = len(self.lines)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        first = getattr(self, "first", 0)
        lines, window = self.lines[first:], getattr(self, "window", "timed region")
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples: timed region shorter than the sampling period"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "samples": len(sm), "window": window}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(st, n_geoms, iters=20):
    """SURVEY.md section 8(d) per-stage algorithmic bytes from the run's measured counts."""
    N, P, C = n_geoms, st["n_pairs"], st["n_contacts"]
    R1, R2 = st["n_rows1"], st["n_rows2"]
    return {
        "broadphase": 184 * N + 8 * P,
        "narrowphase": 104 * P + 48 * C,
        "row_build": 240 * C + 128 * (R1 + R2),
        "solver": iters * (228 * R2 + 132 * R1),
        "integrate_pack": 312 * N,
    }


def compulsory_bytes(st, n_bodies, workload):
    """Once-through HBM bytes of the dominant kernel (what an ideal kernel that keeps every row on chip would still have
    to move).  Island solver (C4, one fused kernel per solve): per body the state in and out + snapshot (SURVEY 8d:
    312 B), per contact the narrowphase record (pos/depth + normal: 32 B) and the unit record (16 B).  Global solver
    (C3): the same plus the row records, which it cannot keep on chip: 96 B written by k_rows is not its traffic, but
    each of the 20 sweeps reads 96 B and writes 16 B per contact."""
    C = st["n_contacts"]
    base = 312 * n_bodies + 48 * C
    if workload == "C4":
        return base
    return base + 20 * 112 * C


def kernel_counters(workload):
    """ncu counters of the dominant kernel from the committed capture (profiles/kernel_counters.json), stamped with the
    commit and the contact count they were measured at; never re-measured inside a bench run (a number taken under a
    profiler is not a bench value, and ncu is not run by bench.py)."""
    path = os.path.join(ROOT, "profiles", "kernel_counters.json")
    try:
        return json.load(open(path)).get(workload)
    except Exception:
        return None


def cpu_port_all_cores(steps, warmup, settle):
    """The CPU oracle port (restatement of libode's QuickStep path, not libode) on every host thread: one block of 16
    C4 worlds per thread, each thread stepping its own OracleWorld (ctypes releases the GIL inside the C call)."""
    import oracle as O
    from odeb200 import scenes
    O.build()
    cores = os.cpu_count() or 1
    wpt = 16  # worlds per thread: 2048 bodies
    worlds = []
    for t in range(cores):
        sc = scenes.batched_worlds_scene(wpt, seed=4, first_world=t * wpt)
        w = O.OracleWorld(gravity=sc["gravity"])
        w.load_scene(sc)
        worlds.append((w, sc["h"]))

    def run(n):
        def work(w, h):
            for _ in range(n):
                w.tick(h)
        ths = [threading.Thread(target=work, args=wh) for wh in worlds]
        for th in ths:
            th.start()
        for th in ths:
            th.join()

    run(settle)
    run(max(warmup, 0))
    t0 = time.perf_counter()
    run(steps)
    dt = time.perf_counter() - t0
    for w, _ in worlds:
        w.close()
    bodies = cores * wpt * 128
    sample = ("%d host threads x %d C4 worlds (%d bodies = a sample of the %d-world workload; worlds are independent, so "
              "per-body throughput carries over), %d ticks after %d settle ticks; CPU restatement of libode's QuickStep path "
              "(oracle port), not libode" % (cores, wpt, bodies, 8192, steps, settle))
    return bodies * steps / dt, dt / steps * 1e3, cores, sample


C4_DESC = "C4: %d independent 128-body worlds (plane + 8x4x4 lattice) per GPU, dt=1/60, QuickStep 20 iters"


def run_reference(args, rank):
    """--impl reference: the CPU port on all host threads, on a bounded sample of the GPU arm's workload."""
    if rank != 0:
        return
    settle = SETTLE["C4"] if args.settle < 0 else args.settle
    value, ms, cores, sample = cpu_port_all_cores(args.steps, args.warmup, settle)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": C4_DESC % args.worlds_per_gpu, "sample": sample,
                   "note": "libode is an un-vendored, un-versioned dependency of the reference and is not installable here: the "
                           "reference arm is the oracle port; each step advances the sample, not the whole workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_line(line)


def bind_to_gpu_numa_node(local_rank):
    """Run this rank (and first-touch its pinned staging buffers) on the CPUs nearest its GPU: at 8 ranks the
    per-tick snapshot / force copies otherwise cross sockets.  Returns the previous affinity (restored for the
    cpu_baseline leg, which uses every host thread)."""
    try:
        import pynvml
        prev = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n + 63) // 64)
        cpus = {i for i in range(n) if (int(words[i // 64]) >> (i % 64)) & 1} & prev
        if cpus:
            os.sched_setaffinity(0, cpus)
        return prev
    except Exception:
        return None


class WorldGroup:
    """The dWorld objects one rank ticks side by side.  Every dWorld has its own CUDA stream and its own CUDA-graph
    replay, so the ticks of different batches overlap on the GPU: the short, latency-bound collide kernels of one batch
    run under the long island-solver kernel of another, and one solver's tail is filled by the next one's start
    (measured on C4, 8192 worlds: 1 batch 1.47 ms per tick, 4 batches 1.39 ms).  Every world is ticked every step."""

    def __init__(self, worlds, n_bodies):
        self.ws, self.nb = list(worlds), list(n_bodies)

    def tick(self, h):
        for w in self.ws:
            w.tick(h)

    def wait(self):
        for w in self.ws:
            w.wait()

    def close(self):
        for w in self.ws:
            w.close()

    def stats(self):
        sts = [w.stats() for w in self.ws]
        out = {}
        for k in sts[0]:
            vals = [st[k] for st in sts]
            if k == "flags":
                v = 0
                for x in vals:
                    v |= int(x)
                out[k] = v
            elif k in ("n_colours", "colour_rounds", "solver_iters", "max_island_rows", "exact_status", "pivot_rounds", "cell_size"):
                out[k] = max(vals)
            elif isinstance(vals[0], list):
                out[k] = [max(c) for c in zip(*vals)] if k == "grid_dims" else [sum(c) for c in zip(*vals)]
            else:
                out[k] = sum(vals)
        return out

    def stage_timings_serial(self, h, reps):
        """per-stage CUDA-event times with one batch on the GPU at a time, summed over the batches (the per-launch
        duration of a kernel, not its share of an overlapped tick)"""
        for w in self.ws:
            w.enable_timing(True)
        acc = []
        for _ in range(reps):
            tot = {}
            for w in self.ws:
                self.wait()
                w.tick(h)
                w.wait()
                for k, v in w.stage_timings().items():
                    tot[k] = tot.get(k, 0.0) + float(v)
            acc.append(tot)
        for w in self.ws:
            w.enable_timing(False)
        return {k: float(np.mean([t[k] for t in acc])) for k in acc[0]}


def c4_world_group(odeb200, n_worlds, first_world, n_batches, device):
    """worlds [first_world, first_world + n_worlds) of BASELINE config 4 in n_batches dWorld objects"""
    from odeb200 import scenes
    ws, nb, ng, first = [], [], [], first_world
    for k in range(n_batches):
        cnt = n_worlds // n_batches + (1 if k < n_worlds % n_batches else 0)
        sc = scenes.batched_worlds_scene(cnt, seed=4, first_world=first)
        first += cnt
        w = odeb200.World(gravity=sc["gravity"], device=device)
        w.load_scene(sc)
        ws.append(w); nb.append(len(sc["bodies"]["pos"])); ng.append(len(sc["geoms"]["type"]))
    grp = WorldGroup(ws, nb)
    grp.ng = ng
    return grp


def timed_device_ticks(L, grp, do_tick, steps, sharding, torch):
    """exactly `steps` ticks between a barrier + synchronize on both sides; CUDA events on the engines' streams: from
    the first world's start event (recorded while every stream is idle) to the LAST stop event of any world"""
    grp.wait()
    sharding.barrier()
    torch.cuda.synchronize()
    for w in grp.ws:
        L.dWorldTimerStartB200(w.w)
    for _ in range(steps):
        do_tick()
    for w in grp.ws:
        L.dWorldTimerStopB200(w.w)
    grp.wait()
    torch.cuda.synchronize()
    sharding.barrier()
    return max(float(L.dWorldTimerElapsedBetweenB200(grp.ws[0].w, w.w)) for w in grp.ws)


def timed_e2e(L, grp, do_tick_of, steps, fmt, expand, sharding, torch, odeb200, wc_forces=True):
    """The same ticks through the C ABI with HOST buffers: per tick and per world the H2D copy of the per-body
    force/torque input (24 B/body, pinned) and the D2H copy of the step's snapshot (64 / 48 / 32 B per body by format),
    both inside the timed region; `expand` additionally rebuilds the reference's 16-float transforms on the host
    (dSnapshotExpandB200).  do_tick_of(k) queues one tick of world k."""
    C = odeb200.C
    floats = {0: 16, 1: 12, 2: 8}[fmt]
    K = len(grp.ws)
    for w in grp.ws:
        L.dWorldSetSnapshotFormatB200(w.w, fmt)
    # force / torque input: page-locked, write-combined (the host only writes it) -- dAllocPinnedB200; or torch's plain pinned
    f6_raw = []
    if wc_forces:
        f6 = []
        for n in grp.nb:
            ptr = L.dAllocPinnedB200(n * 24, 1)
            f6_raw.append(ptr)
            C.memset(ptr, 0, n * 24)
            f6.append(ptr)
        fp_list = [C.cast(ptr, C.POINTER(C.c_float)) for ptr in f6]
    else:
        f6 = [torch.zeros((n, 6), dtype=torch.float32).pin_memory() for n in grp.nb]
        fp_list = None
    snap = [[torch.empty((n, floats), dtype=torch.float32).pin_memory() for _ in range(2)] for n in grp.nb]
    full = [torch.empty((n, 16), dtype=torch.float32) if expand else None for n in grp.nb]
    fp = fp_list if fp_list is not None else [C.cast(t.data_ptr(), C.POINTER(C.c_float)) for t in f6]
    threads = max(1, (os.cpu_count() or 1) // max(1, sharding.dist_env()[2]))

    def step(i):
        for k, w in enumerate(grp.ws):
            L.dWorldSetForcesB200(w.w, fp[k], grp.nb[k])
            do_tick_of(k)
            L.dWorldGetSnapshotB200(w.w, snap[k][i & 1].data_ptr(), 0, grp.nb[k], 0)
            if expand and i > 0:      # expand tick i-1's records (already on the host) while tick i runs on the GPU
                L.dSnapshotExpandB200(snap[k][(i - 1) & 1].data_ptr(), fmt, grp.nb[k], full[k].data_ptr(), threads)

    for i in range(3):
        step(i)
    grp.wait()
    sharding.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    grp.wait()
    if expand:
        for k in range(K):
            L.dSnapshotExpandB200(snap[k][(steps - 1) & 1].data_ptr(), fmt, grp.nb[k], full[k].data_ptr(), threads)
    torch.cuda.synchronize()
    dt_ms = (time.perf_counter() - t0) * 1e3
    sharding.barrier()
    for k in range(K):
        ok = float(snap[k][0][0, 3 if fmt == 2 else floats - 1]) == 1.0 if fmt != 1 else True
        assert ok and (not expand or float(full[k][grp.nb[k] - 1, 15]) == 1.0)
    for w in grp.ws:
        L.dWorldSetSnapshotFormatB200(w.w, 0)
    for ptr in f6_raw:
        L.dFreePinnedB200(ptr)
    return dt_ms


def secondary_c3(L, odeb200, local_rank, peak):
    """The HBM-bound config (BASELINE config 3, the one north_star's ">= 50 % of HBM" is about) on this rank's GPU:
    20 timed ticks + per-stage CUDA-event times -> solver and broadphase fractions of the measured HBM peak."""
    sc, desc = build_scene("C3", 0, 0)
    nb, ng = len(sc["bodies"]["pos"]), len(sc["geoms"]["type"])
    ew = odeb200.World(gravity=sc["gravity"], device=local_rank)
    ew.load_scene(sc)
    h = sc["h"]
    for _ in range(SETTLE["C3"] + 5):
        ew.tick(h)
    ew.wait()
    L.dWorldTimerStartB200(ew.w)
    for _ in range(20):
        ew.tick(h)
    L.dWorldTimerStopB200(ew.w)
    ms = float(L.dWorldTimerElapsedB200(ew.w)) / 20
    st = ew.stats()
    ew.enable_timing(True)
    tm = []
    for _ in range(6):
        ew.tick(h)
        ew.wait()
        tm.append(ew.stage_timings())
    ew.enable_timing(False)
    ew.close()
    tmean = {k: float(np.mean([t[k] for t in tm])) for k in tm[0]}
    ab = algorithmic_bytes(st, ng)
    solver_alg = ab["solver"] + ab["integrate_pack"]
    kc = kernel_counters("C3") or {}
    out = {
        "workload": desc, "ms_per_step": ms, "value": nb / (ms * 1e-3), "unit": UNIT, "stage_ms": tmean,
        "counts": {k: st[k] for k in ("n_pairs", "n_contacts", "n_manifolds", "n_rows1", "n_rows2", "n_colours")},
        "solver": {"kernel": "k_solve (20 PGS sweeps x colour phases separated by grid barriers, fused integrate + snapshot pack)",
                   "bound": "hbm", "algorithmic_bytes_per_launch": solver_alg, "compulsory_bytes_per_launch": compulsory_bytes(st, nb, "C3"),
                   "kernel_ms": tmean["solve_ms"],
                   "achieved": solver_alg / (tmean["solve_ms"] * 1e-3) / 1e9, "unit": "GB/s",
                   "frac": solver_alg / (tmean["solve_ms"] * 1e-3) / 1e9 / peak,
                   "compulsory_frac": compulsory_bytes(st, nb, "C3") / (tmean["solve_ms"] * 1e-3) / 1e9 / peak,
                   "note": "achieved = SURVEY 8(d) algorithmic bytes (a row-streaming QuickStep: 228 / 132 B per row-sweep) / event-timed "
                           "kernel time; compulsory = this engine's own layout (96 B read + 16 B written per contact-sweep, 312 B per "
                           "body, 48 B per contact once); ncu DRAM bytes of the committed capture in `ncu`",
                   "ncu": kc.get("k_solve")},
        "broadphase": {"kernels": "k_geom_update, k_cell_keys, radix sort, k_sorted_records, k_sweep<count>, scan, k_sweep<fill>",
                       "bound": "issue / latency (ncu: lanes per instruction, L1-served candidate reads)",
                       "algorithmic_bytes": ab["broadphase"], "stage_ms": tmean["broadphase_ms"],
                       "achieved": ab["broadphase"] / (tmean["broadphase_ms"] * 1e-3) / 1e9, "unit": "GB/s",
                       "frac": ab["broadphase"] / (tmean["broadphase_ms"] * 1e-3) / 1e9 / peak,
                       "ncu": kc.get("k_sweep")},
        "narrowphase": {"algorithmic_bytes": ab["narrowphase"], "stage_ms": tmean["narrowphase_ms"],
                        "frac": ab["narrowphase"] / (tmean["narrowphase_ms"] * 1e-3) / 1e9 / peak},
    }
    return out


def secondary_c5(L, odeb200, rank, local_rank, world, dev, sharding, torch, cols=64, settle=200, steps=20):
    """BASELINE config 5 on the N GPUs of this run: one slab of `cols` x 1024 x 16 lattice columns per GPU (1 M bodies per
    GPU), the C slab driver (dSlabTickB200: NCCL send/recv inside the library, event-ordered), dynamic halo + impulse
    coupling + migration every 16 ticks.  The same ticks without the halo (ghosts frozen) give the exchange's share."""
    from odeb200 import slabs
    sc, info = slabs.dynamic_slab_scene(rank, world, nx_per_slab=cols, nz=1024, ny=16, seed=5, margin_cols=4)
    ew = odeb200.World(gravity=sc["gravity"], device=local_rank)
    ew.load_scene(sc)
    h = sc["h"]
    slab = slabs.CSlab(ew, info, slabs.nccl_unique_id(L, rank, world))
    n = [0]

    def tick():
        if n[0] % 16 == 0 and n[0] > 0:
            slab.migrate()
        slab.tick(h)
        n[0] += 1
    for _ in range(settle):
        tick()
    n[0] = 1                                   # no migration inside the timed region's first tick
    grp = WorldGroup([ew], [0])
    ms = sharding.all_reduce_max(timed_device_ticks(L, grp, tick, steps, sharding, torch), dev) / steps
    ms_local = sharding.all_reduce_max(timed_device_ticks(L, grp, lambda: ew.tick(h), steps, sharding, torch), dev) / steps
    inf = slab.get_info()
    owned = sharding.all_reduce_sum(inf["n_owned"], dev)
    sel = sharding.all_reduce_max(inf["halo_selected"], dev)
    ovf = sharding.all_reduce_max(inf["halo_overflow"] + inf["mig_overflow"], dev)
    mig = sharding.all_reduce_sum(inf["migrated_out"], dev)
    sent = sharding.all_reduce_max(inf["halo_bytes_per_tick"], dev)
    out = {"workload": "C5: slab-decomposed single world, %d x 1024 x 16 lattice columns per GPU, C slab driver (NCCL send/recv "
                       "inside libode_b200.so, event-ordered), dynamic halo + impulse coupling + migration every 16 ticks" % cols,
           "bodies_total": int(owned), "ms_per_step": ms, "value": owned / (ms * 1e-3), "unit": UNIT,
           "ms_per_step_without_halo": ms_local, "halo_share": max(0.0, 1.0 - ms_local / ms),
           "halo_bytes_sent_per_tick_per_gpu_max": int(sent), "halo_bodies_selected_per_face_max": int(sel),
           "halo_or_migration_overflow_max": int(ovf), "bodies_migrated_total": int(mig)}
    slab.close()
    ew.close()
    return out


_REAL_STDOUT = None


def emit_line(line):
    """the ONE line of the contract, on the process's real stdout"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # stdout carries ONE JSON line and nothing else.  Libraries print there too (NCCL's "NCCL version ..." banner at
    # NCCL_DEBUG >= VERSION, from torch's communicator and from the one the slab driver creates inside libode_b200.so), so
    # file descriptor 1 points at stderr for the whole run and the line goes to a duplicate of the real stdout.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    from odeb200 import sharding
    rank, local_rank, world = sharding.dist_env()
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import odeb200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libode_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    prev_affinity = bind_to_gpu_numa_node(local_rank) if (world > 1 and not os.environ.get("ODE_B200_NO_BIND")) else None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        sharding.init_process_group("nccl")
    dev = "cuda:%d" % local_rank
    L = odeb200.lib()

    slab = None
    if args.workload == "C5":
        # one slab of the 1024(z) x 16(y) lattice per GPU, halo exchange of boundary bodies over NCCL
        from odeb200 import slabs
        if args.halo == "dynamic":
            sc, halo = slabs.dynamic_slab_scene(rank, world, nx_per_slab=args.slab_cols, nz=1024, ny=16, seed=5, margin_cols=4)
            args.coupling = "impulse"
        else:
            sc, halo = slabs.slab_scene(rank, world, nx_per_slab=args.slab_cols, nz=1024, ny=16, seed=5, margin_cols=4,
                                         coupling=args.coupling)
        desc = ("C5: slab-decomposed single world, %d x 1024 x 16 lattice columns per GPU (%d bodies/GPU), NCCL halo exchange "
                "each tick (%s; %s), dt=1/60, QuickStep 20 iters"
                % (args.slab_cols, args.slab_cols * 1024 * 16,
                   ("boundary set selected on the device from current positions, whole-body records into a ghost pool, "
                    "ownership migrates every %d ticks" % args.migrate_every) if args.halo == "dynamic"
                   else "boundary set fixed by the initial lattice column",
                   "boundary-body states to the lower slab, contact impulses back to the owner" if args.coupling == "impulse"
                   else "boundary-body states both ways, kinematic ghosts"))
        n_bodies = sc["n_owned"]
    # (overlap pays down to ~1024 worlds per batch: measured, 4096 worlds per GPU run 1.38e9 in four batches and 1.25e9 in one
    # at 2 GPUs; 1024 worlds per GPU run 2.54e9 in four batches of 256 and 2.66e9 in one at 8 GPUs)
    def batches_for(n_worlds):
        return max(1, min(args.batches, n_worlds // 1024))
    n_batches = batches_for(args.worlds_per_gpu) if args.workload == "C4" else 1
    if n_batches > 1:
        # the same worlds as one big batch (seed 4 + global world index), held in n_batches dWorld objects
        grp = c4_world_group(odeb200, args.worlds_per_gpu, rank * args.worlds_per_gpu, n_batches, local_rank)
        ew = grp.ws[0]
        n_bodies, n_geoms, h = sum(grp.nb), sum(grp.ng), 1.0 / 60.0
        desc = C4_DESC % args.worlds_per_gpu
    else:
        if args.workload != "C5":
            sc, desc = build_scene(args.workload, rank, args.worlds_per_gpu)
            n_bodies = len(sc["bodies"]["pos"])
        n_geoms = len(sc["geoms"]["type"])
        h = sc["h"]
        ew = odeb200.World(gravity=sc["gravity"], device=local_rank)
        ew.load_scene(sc)
        grp = WorldGroup([ew], [n_bodies])
        del sc
    if args.workload == "C5":
        tick_no = [0]
        if args.halo == "dynamic":
            # the C driver inside libode_b200.so: buffers, NCCL send/recv and the event ordering live in the library
            slab = slabs.CSlab(ew, halo, slabs.nccl_unique_id(L, rank, world) if world > 1 else None)

            def do_tick():
                if args.migrate_every > 0 and tick_no[0] % args.migrate_every == 0 and tick_no[0] > 0:
                    slab.migrate()
                slab.tick(h)
                tick_no[0] += 1
        else:
            slab = slabs.SlabWorld(ew, halo, dev)
            exch = (lambda kind: slabs.exchange_nccl(slab, rank, world, kind)) if world > 1 else (lambda kind: None)

            def do_tick():
                slabs.tick(slab, exch, h)
                tick_no[0] += 1
    else:
        def do_tick():
            grp.tick(h)
    if args.workload == "C5":
        do_tick_of = lambda k: do_tick()          # noqa: E731  (one world per rank)
    else:
        do_tick_of = lambda k: grp.ws[k].tick(h)  # noqa: E731
    settle = SETTLE[args.workload] if args.settle < 0 else args.settle
    for _ in range(settle):          # scene preparation (bodies dropped onto the ground), untimed
        do_tick()
    grp.wait()

    # ---------------- device-resident throughput: W warm-up ticks, then exactly K timed ticks
    sampler = ClockSampler(local_rank)
    sampler.start()                  # nvidia-smi needs ~1 s to start; only samples taken under load are kept
    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        do_tick()
    grp.wait()
    sampler.mark()
    launches0 = L.dGetKernelLaunchCountB200()
    elapsed_ms = timed_device_ticks(L, grp, do_tick, args.steps, sharding, torch)
    launches = L.dGetKernelLaunchCountB200() - launches0
    t_timed_end = time.perf_counter()
    st = grp.stats()
    if st["flags"] != 0:
        # a capacity overflow (pairs, solver units, trimesh candidates) means contacts were dropped: the engine flags it and
        # carries on, but a throughput measured that way would be a number with work skipped
        raise SystemExit("bench.py: the engine flagged a capacity overflow (dStepStatsB200.flags = %d) inside the timed region" % st["flags"])
    # per-kernel duration of the dominant kernel: CUDA events on the engine's stream around the solver launch, on
    # eight more live ticks right behind the timed ones (stage events are off inside the timed region: with them
    # the engine does not replay the tick as a CUDA graph)
    if args.workload == "C5":
        ew.enable_timing(True)
        stage = []
        for _ in range(8):
            do_tick()
            ew.wait()
            stage.append(ew.stage_timings())
        ew.enable_timing(False)
        tm = {k: float(np.mean([t[k] for t in stage])) for k in stage[0]}
    else:
        tm = grp.stage_timings_serial(h, 8)
    t_max = sharding.all_reduce_max(elapsed_ms, dev)
    total_bodies = sharding.all_reduce_sum(n_bodies, dev)
    value = total_bodies * args.steps / (t_max * 1e-3)

    # ---------------- end to end through the C ABI with HOST buffers (pinned)
    e2e = None
    if not args.no_e2e:
        fmt = args.snapshot_format
        rec = {0: 64, 1: 48, 2: 32}
        dt = sharding.all_reduce_max(timed_e2e(L, grp, do_tick_of, args.steps, fmt, False, sharding, torch, odeb200), dev)
        e2e = {"value": total_bodies * args.steps / (dt * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(n_bodies * 24 * world),
               "d2h_bytes_per_step": int(n_bodies * rec[fmt] * world), "ms_per_step": dt / args.steps,
               "snapshot_format": {0: "16 floats per body: the reference's GetTransformMat layout (src/main.c:602-622)",
                                   1: "12 floats per body: GetTransformMat without its four constant floats",
                                   2: "8 floats per body: position + quaternion (dBodyGetPosition / dBodyGetQuaternion)"}[fmt]}
        if fmt != 0 and not args.no_secondary:
            # the same loop with the full 64 B records, and with the compact records expanded to them on the host
            d0 = sharding.all_reduce_max(timed_e2e(L, grp, do_tick_of, args.steps, 0, False, sharding, torch, odeb200), dev)
            dx = sharding.all_reduce_max(timed_e2e(L, grp, do_tick_of, args.steps, fmt, True, sharding, torch, odeb200), dev)
            dp = sharding.all_reduce_max(timed_e2e(L, grp, do_tick_of, args.steps, fmt, False, sharding, torch, odeb200, wc_forces=False), dev)
            e2e["variants_note"] = ("the variant legs run after the default leg, i.e. later in the simulation (the heaps carry more "
                                    "contacts by then): compare them with each other, not with the default leg")
            e2e["variants"] = {
                "forces_in_plain_pinned_memory": {"value": total_bodies * args.steps / (dp * 1e-3), "ms_per_step": dp / args.steps,
                                                  "note": "the default leg uploads the forces from write-combined pinned memory (dAllocPinnedB200)"},
                "format0_64B_per_body": {"value": total_bodies * args.steps / (d0 * 1e-3), "ms_per_step": d0 / args.steps,
                                         "d2h_bytes_per_step": int(n_bodies * 64 * world)},
                "compact_then_expanded_on_host": {"value": total_bodies * args.steps / (dx * 1e-3), "ms_per_step": dx / args.steps,
                                                  "note": "dSnapshotExpandB200 rebuilds the 16-float transforms of tick t-1 on "
                                                          "the host's threads while tick t runs"}}

    # ---------------- clocks: the sampler has been running since the start of the device-timed region.  An nvidia-smi query
    # takes ~0.1 s, so short regions see one sample or none: keep the same ticks running (untimed, after every measured leg,
    # so that no leg sees a later phase of the simulation) until ~0.8 s of load have been sampled, and say so
    loaded = time.perf_counter() - t_timed_end + elapsed_ms * 1e-3
    # (a tick COUNT agreed by all ranks, not a deadline: C5's ticks exchange halos, so every rank must run the same number)
    n_extra = int(max(0.0, 0.8 - loaded) / max(elapsed_ms * 1e-3 / args.steps, 1e-5))
    n_extra = int(sharding.all_reduce_max(float(n_extra), dev))
    for i in range(n_extra):
        do_tick()
        if i % 10 == 9:
            grp.wait()
    grp.wait()
    sampler.window = ("device-timed region (%.0f ms), the stage-timing and end-to-end legs, and the same ticks kept running after them: "
                      "%.1f s of load" % (elapsed_ms, max(loaded, 0.8)))
    clocks = sampler.stop()

    # ---------------- BASELINE config 4 as written: 8192 worlds in TOTAL, sharded over the GPUs (strong scaling)
    strong = None
    if world > 1 and args.workload == "C4" and not args.no_secondary:
        first, cnt = sharding.shard_range(args.worlds_per_gpu, rank, world)
        nb2 = cnt * 128
        grp2 = c4_world_group(odeb200, cnt, first, batches_for(cnt), local_rank)
        tick2 = lambda: grp2.tick(h)  # noqa: E731
        for _ in range(settle + warmup):
            tick2()
        ms2 = sharding.all_reduce_max(timed_device_ticks(L, grp2, tick2, args.steps, sharding, torch), dev)
        tot2 = sharding.all_reduce_sum(nb2, dev)
        strong = {"scaling": "strong", "workload": "BASELINE config 4 as written: %d worlds in total, %d per GPU (in %d dWorld batches)"
                  % (args.worlds_per_gpu, cnt, len(grp2.ws)),
                  "value": tot2 * args.steps / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2 / args.steps}
        if not args.no_e2e:
            dt2 = sharding.all_reduce_max(timed_e2e(L, grp2, lambda k: grp2.ws[k].tick(h), args.steps, args.snapshot_format, False, sharding, torch, odeb200), dev)
            strong["e2e"] = {"value": tot2 * args.steps / (dt2 * 1e-3), "ms_per_step": dt2 / args.steps}
        grp2.close()

    # ---------------- BASELINE config 5 under the driver's eyes: the slab-decomposed single world on the same N GPUs
    c5 = None
    if world > 1 and args.workload == "C4" and not args.no_secondary:
        grp.close()
        grp = None
        c5 = secondary_c5(L, odeb200, rank, local_rank, world, dev, sharding, torch)

    if rank == 0:
        peak, peak_src = measured_peaks()
        ab = algorithmic_bytes(st, n_geoms)
        t_solve = tm["solve_ms"] * 1e-3
        island = args.workload == "C4"
        comp = compulsory_bytes(st, n_bodies, "C4" if island else "C3")
        solver_alg = ab["solver"] + ab["integrate_pack"]
        kc = (kernel_counters(args.workload) or {}).get("k_env_solve2" if island else "k_solve")
        if island:
            # The island solver keeps an env's rows in L2 / shared memory for its 20 sweeps, so SURVEY 8(d)'s row-streaming byte
            # model does not describe it (it would read 1.8x the HBM peak): the kernel is bound by instruction issue.  frac =
            # once-through compulsory bytes / kernel time / peak says how far from an HBM-bound kernel it is.
            roofline = {"bound": "issue", "kernel": "k_env_solve2 (lane-pair island solver: body preparation, colouring, rows, 20 PGS "
                        "sweeps, integrate + snapshot pack of one world per warp, one launch per solve)",
                        "achieved": comp / t_solve / 1e9, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                        "frac": comp / t_solve / 1e9 / peak, "compulsory_bytes_per_launch": comp,
                        "row_streaming_model_bytes_per_launch": solver_alg, "row_streaming_model_GBps": solver_alg / t_solve / 1e9}
        else:
            roofline = {"bound": "hbm", "kernel": "k_solve (20 PGS sweeps x colour phases, fused integrate + snapshot pack)",
                        "achieved": solver_alg / t_solve / 1e9, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                        "frac": solver_alg / t_solve / 1e9 / peak, "algorithmic_bytes_per_launch": solver_alg,
                        "compulsory_bytes_per_launch": comp, "compulsory_frac": comp / t_solve / 1e9 / peak}
        if n_batches > 1:
            roofline["launches_per_step"] = n_batches
            kb = (kernel_counters(args.workload) or {}).get("k_env_solve2_batch2048")
            if kb:
                roofline["ncu_one_batch_launch"] = kb   # the committed capture of ONE 2048-world launch (ncu below: 8192 worlds in one)
            roofline["kernel_ms_note"] = ("kernel_ms and stage_ms are sums over the %d batches, each timed with the GPU to itself "
                                          "(the per-launch durations); inside the timed region the batches overlap" % n_batches)
        roofline.update({"frac_of_8TBps_spec": roofline["achieved"] / 8000.0, "kernel_ms": t_solve * 1e3, "stage_ms": tm,
                         "whole_tick_algorithmic_GBps": sum(ab.values()) / (tm["tick_ms"] * 1e-3) / 1e9})
        roofline["traffic"] = None
        if kc:
            # ncu --set full capture of the same kernel (per launch), committed under profiles/ and stamped with its commit:
            # not re-measured by this run
            roofline["traffic"] = kc.get("dram_bytes_per_launch")
            roofline["ncu"] = kc
            if kc.get("dram_bytes_per_launch"):
                roofline["dram_GBps"] = kc["dram_bytes_per_launch"] / t_solve / 1e9
                roofline["dram_frac"] = roofline["dram_GBps"] / peak
        cpu = None
        secondary = None
        if world == 1 and not args.no_cpu_baseline:
            if prev_affinity:
                os.sched_setaffinity(0, prev_affinity)
            v, ms, cores, sample = cpu_port_all_cores(20, 3, SETTLE["C4"])
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        if world == 1 and args.workload == "C4" and not args.no_secondary:
            grp.close()
            grp = None
            secondary = {"C3": secondary_c3(L, odeb200, local_rank, peak)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": t_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": desc, "bodies_per_gpu": n_bodies, "settle_steps": settle,
                       "world_batches": ("the %d worlds of a GPU are held in %d dWorld objects of %d worlds (one CUDA stream each, their "
                                         "ticks overlap on the GPU); every world is ticked every step" %
                                         (args.worlds_per_gpu, n_batches, args.worlds_per_gpu // n_batches)) if n_batches > 1 else 1,
                       "halo_bytes_per_tick_per_gpu": (slab.halo_bytes() if slab else 0),
                       "migrated_out_rank0": (getattr(slab, "migrated_out", 0) if slab else 0),
                       "l2": "inputs larger than L2: ~%.0f MB of body, contact and row arrays are touched per tick (126 MB L2)"
                             % ((sum(ab.values()) / 20 + 200 * n_bodies) / 1e6),
                       "counts": {k: st[k] for k in ("n_pairs", "n_contacts", "n_manifolds", "n_rows1", "n_rows2", "n_colours", "env_trips", "env_lanes", "flags")}},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        if secondary:
            line["secondary"] = secondary
        if strong:
            line["strong_scaling"] = strong
        if c5:
            line.setdefault("secondary", {})["C5"] = c5
        emit_line(line)
    if grp is not None:
        grp.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
