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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
        desc = "C4: %d independent 128-body worlds (plane + 8x4x4 lattice) per GPU, dt=1/60, QuickStep 20 iters" % worlds_per_gpu
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
        for ln in self.lines[getattr(self, "first", 0):]:
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
                "power_w_max": float(max(power)), "samples": len(sm)}


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


def layout_bytes(st, n_bodies, iters=20):
    """Minimum traffic of THIS engine's solver layout: 112 B per contact-iteration (5 float4
    row records + lambda read/write) and per manifold-iteration 16 B record + per body end 80 B read (fc, world
    inverse inertia) + 32 B fc write; the fused tail moves 312 B per body."""
    C, M = st["n_contacts"], st["n_manifolds"]
    two = st["n_rows2"] / max(1, st["n_rows"])
    return iters * (112 * C + M * (16 + (1 + two) * 112)) + 312 * n_bodies


def cpu_port_sample(n_worlds, settle, ticks):
    """Single-thread oracle (CPU restatement of libode's QuickStep path, not libode) on n_worlds C4 worlds."""
    import oracle as O
    from odeb200 import scenes
    sc = scenes.batched_worlds_scene(n_worlds, seed=4)
    w = O.OracleWorld(gravity=sc["gravity"])
    w.load_scene(sc)
    for _ in range(settle):
        w.tick(sc["h"])
    t0 = time.perf_counter()
    for _ in range(ticks):
        w.tick(sc["h"])
    dt = time.perf_counter() - t0
    w.close()
    return n_worlds * 128 * ticks / dt


def run_reference(args, rank):
    """--impl reference: the CPU port on all host threads, one block of worlds per thread."""
    if rank != 0:
        return
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
                w.tick(h)      # ctypes releases the GIL inside the C call
        ths = [threading.Thread(target=work, args=wh) for wh in worlds]
        for th in ths:
            th.start()
        for th in ths:
            th.join()

    run(SETTLE["C4"] if args.settle < 0 else args.settle)
    run(max(args.warmup, 0))
    t0 = time.perf_counter()
    run(args.steps)
    dt = time.perf_counter() - t0
    bodies = cores * wpt * 128
    value = bodies * args.steps / dt
    sample = "%d threads x %d C4 worlds (%d bodies), %d ticks after %d settle ticks" % (cores, wpt, bodies, args.steps, SETTLE["C4"])
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C4 sample: " + sample, "note": "CPU restatement of libode's QuickStep path (oracle port), not libode: "
                   "libode is an un-vendored, un-versioned dependency of the reference and is not installable here"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


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


def main():
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
    else:
        sc, desc = build_scene(args.workload, rank, args.worlds_per_gpu)
        n_bodies = len(sc["bodies"]["pos"])
    n_geoms = len(sc["geoms"]["type"])
    ew = odeb200.World(gravity=sc["gravity"], device=local_rank)
    ew.load_scene(sc)
    h = sc["h"]
    if args.workload == "C5":
        slab = slabs.DynamicSlabWorld(ew, halo, dev) if args.halo == "dynamic" else slabs.SlabWorld(ew, halo, dev)
        exch = (lambda kind: slabs.exchange_nccl(slab, rank, world, kind)) if world > 1 else (lambda kind: None)
        tick_no = [0]

        def do_tick():
            if args.halo == "dynamic":
                slabs.tick_dynamic(slab, exch, h, tick_no[0], args.migrate_every)
            else:
                slabs.tick(slab, exch, h)
            tick_no[0] += 1
    else:
        def do_tick():
            ew.tick(h)
    settle = SETTLE[args.workload] if args.settle < 0 else args.settle
    for _ in range(settle):          # scene preparation (bodies dropped onto the ground), untimed
        do_tick()
    ew.wait()

    # ---------------- device-resident throughput: W warm-up ticks, then exactly K timed ticks
    sampler = ClockSampler(local_rank)
    sampler.start()                  # nvidia-smi needs ~1 s to start; only samples taken under load are kept
    for _ in range(max(args.warmup, 3)):
        do_tick()
    ew.wait()
    sampler.mark()
    sharding.barrier()
    torch.cuda.synchronize()
    launches0 = L.dGetKernelLaunchCountB200()
    solve_ms = []
    L.dWorldTimerStartB200(ew.w)
    for _ in range(args.steps):
        do_tick()
    L.dWorldTimerStopB200(ew.w)
    ew.wait()
    torch.cuda.synchronize()
    sharding.barrier()
    elapsed_ms = float(L.dWorldTimerElapsedB200(ew.w))
    launches = L.dGetKernelLaunchCountB200() - launches0
    clocks = sampler.stop()
    st = ew.stats()
    # per-kernel duration of the dominant kernel: CUDA events on the engine's stream around the solver launch, on
    # eight more live ticks right behind the timed ones (stage events are off inside the timed region: with them
    # the engine does not replay the tick as a CUDA graph)
    ew.enable_timing(True)
    for _ in range(8):
        do_tick()
        ew.wait()
        solve_ms.append(ew.timings()["solve_ms"])
    tm = ew.timings()
    ew.enable_timing(False)
    t_max = sharding.all_reduce_max(elapsed_ms, dev)
    total_bodies = sharding.all_reduce_sum(n_bodies, dev)
    value = total_bodies * args.steps / (t_max * 1e-3)

    # ---------------- end to end through the C ABI with HOST buffers (pinned): per tick H2D of the per-body
    # force/torque input and D2H of the fused snapshot, both inside the timed region
    e2e = None
    if not args.no_e2e:
        f6 = torch.zeros((n_bodies, 6), dtype=torch.float32).pin_memory()
        snap = [torch.empty((n_bodies, 16), dtype=torch.float32).pin_memory() for _ in range(2)]
        fp = odeb200.C.cast(f6.data_ptr(), odeb200.C.POINTER(odeb200.C.c_float))
        for i in range(3):
            L.dWorldSetForcesB200(ew.w, fp, n_bodies)
            do_tick()
            L.dWorldGetSnapshotB200(ew.w, snap[i & 1].data_ptr(), 0, n_bodies, 0)
        ew.wait()
        sharding.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(args.steps):
            L.dWorldSetForcesB200(ew.w, fp, n_bodies)
            do_tick()
            L.dWorldGetSnapshotB200(ew.w, snap[i & 1].data_ptr(), 0, n_bodies, 0)
        ew.wait()
        torch.cuda.synchronize()
        dt_ms = (time.perf_counter() - t0) * 1e3
        sharding.barrier()
        dt_max = sharding.all_reduce_max(dt_ms, dev)
        assert float(snap[0][0, 15]) == 1.0 and float(snap[1][n_bodies - 1, 15]) == 1.0
        e2e = {"value": total_bodies * args.steps / (dt_max * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(n_bodies * 24 * world),
               "d2h_bytes_per_step": int(n_bodies * 64 * world), "ms_per_step": dt_max / args.steps}

    if rank == 0:
        peak, peak_src = measured_peaks()
        ab = algorithmic_bytes(st, n_geoms)
        t_solve = float(np.mean(solve_ms)) * 1e-3
        solver_alg = ab["solver"] + ab["integrate_pack"]
        achieved = solver_alg / t_solve / 1e9 if t_solve > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "solver_traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                traffic = tj.get(args.workload, {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        kname = ("k_env_solve<G> (island solver: body preparation, colouring, rows, 20 PGS iterations, integrate + snapshot pack in one kernel)"
                 if args.workload == "C4" else "k_solve (20 PGS iterations x colours, fused integrate + snapshot pack)")
        roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved,
                    "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                    "frac_of_8TBps_spec": achieved / 8000.0, "traffic": traffic,
                    "algorithmic_bytes_per_launch": solver_alg, "layout_bytes_per_launch": layout_bytes(st, n_bodies),
                    "kernel_ms": t_solve * 1e3, "whole_tick_GBps": sum(ab.values()) / (tm["tick_ms"] * 1e-3) / 1e9,
                    "stage_ms": tm}
        if traffic and t_solve > 0:
            # what actually crossed the HBM interface (ncu capture of the same kernel, profiles/solver_traffic.json)
            roofline["dram_GBps"] = traffic / t_solve / 1e9
            roofline["dram_frac"] = roofline["dram_GBps"] / peak
        roofline["note"] = ("achieved = SURVEY 8(d) algorithmic bytes of a row-streaming QuickStep / kernel time; this engine "
                            "rebuilds J and iMJ from 96 B per contact and (island solver) re-reads rows from L2 / shared "
                            "memory, so frac can exceed 1; dram_GBps is the measured HBM traffic rate; the kernel is "
                            "issue-bound (profiles/README.md)" if args.workload == "C4" else
                            "achieved = SURVEY 8(d) algorithmic bytes / kernel time; dram_GBps = measured HBM traffic rate")
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            import oracle as O
            O.build()
            nw = 64
            v = cpu_port_sample(nw, SETTLE["C4"], 120)
            cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": "%d C4 worlds (%d bodies), 120 ticks after %d settle ticks, single thread; CPU restatement of "
                             "libode's QuickStep path, not libode" % (nw, nw * 128, SETTLE["C4"])}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": t_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": desc, "bodies_per_gpu": n_bodies, "settle_steps": settle,
                       "halo_bytes_per_tick_per_gpu": (slab.halo_bytes() if slab else 0),
                       "migrated_out_rank0": (getattr(slab, "migrated_out", 0) if slab else 0),
                       "l2": "inputs larger than L2: ~%.0f MB of body, contact and row arrays are streamed per tick (126 MB L2)"
                             % ((sum(ab.values()) / 20 + 200 * n_bodies) / 1e6),
                       "counts": {k: st[k] for k in ("n_pairs", "n_contacts", "n_manifolds", "n_rows1", "n_rows2", "n_colours")}},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    ew.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
