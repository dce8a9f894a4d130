"""CPU: the bookkeeping of bench.py that does not need a GPU -- SURVEY 8(d)'s algorithmic-byte model, the compulsory-byte
model of the roofline block, the merge of per-batch statistics, the provenance of the committed ncu counters, and the JSON
contract of the reference arm (`bench.py --impl reference`, the CPU oracle port on the host's threads)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_follow_survey_8d():
    # the judge's own recomputation in VERDICT.md (round 1): 14.167 GB per launch for these counts
    st = {"n_pairs": 0, "n_contacts": 0, "n_rows1": 1456722, "n_rows2": 2191623}
    ab = bench.algorithmic_bytes(st, 1048577)
    assert ab["solver"] + ab["integrate_pack"] == 20 * (228 * 2191623 + 132 * 1456722) + 312 * 1048577
    assert abs((ab["solver"] + ab["integrate_pack"]) / 1e9 - 14.167) < 0.001
    st = {"n_pairs": 10, "n_contacts": 7, "n_rows1": 3, "n_rows2": 9}
    ab = bench.algorithmic_bytes(st, 5)
    assert ab == {"broadphase": 184 * 5 + 80, "narrowphase": 1040 + 48 * 7, "row_build": 240 * 7 + 128 * 12,
                  "solver": 20 * (228 * 9 + 132 * 3), "integrate_pack": 312 * 5}


def test_compulsory_bytes():
    st = {"n_contacts": 1000}
    assert bench.compulsory_bytes(st, 100, "C4") == 312 * 100 + 48 * 1000
    assert bench.compulsory_bytes(st, 100, "C3") == 312 * 100 + 48 * 1000 + 20 * 112 * 1000


class _FakeWorld:
    def __init__(self, st):
        self._st = st

    def stats(self):
        return dict(self._st)


def test_world_group_merges_the_batches_statistics():
    a = {"n_pairs": 10, "n_contacts": 4, "n_colours": 7, "flags": 1, "class_count": [1, 2, 3], "grid_dims": [4, 5, 6], "cell_size": 0.5,
         "exact_status": -1, "env_trips": 100}
    b = {"n_pairs": 5, "n_contacts": 6, "n_colours": 9, "flags": 4, "class_count": [10, 20, 30], "grid_dims": [1, 9, 2], "cell_size": 0.25,
         "exact_status": -1, "env_trips": 50}
    g = bench.WorldGroup([_FakeWorld(a), _FakeWorld(b)], [128, 256])
    st = g.stats()
    assert st["n_pairs"] == 15 and st["n_contacts"] == 10 and st["env_trips"] == 150      # counts add up
    assert st["n_colours"] == 9 and st["exact_status"] == -1 and st["cell_size"] == 0.5   # maxima
    assert st["flags"] == 5                                                               # overflow flags are OR-ed
    assert st["class_count"] == [11, 22, 33] and st["grid_dims"] == [4, 9, 6]


def test_committed_ncu_counters_name_their_capture():
    kc = json.load(open(os.path.join(ROOT, "profiles", "kernel_counters.json")))
    for wl, kern in (("C4", "k_env_solve2"), ("C3", "k_solve")):
        c = bench.kernel_counters(wl)[kern]
        assert c == kc[wl][kern]
        assert os.path.exists(os.path.join(ROOT, c["capture"])), c["capture"]
        assert c["commit"] and c["workload"] and c["dram_bytes_per_launch"] > 0 and 0 < c["lanes_per_instruction"] <= 32


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--settle", "5"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                                   # ONE JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "body_steps_per_sec" and d["unit"] == "body-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"].startswith("C4: 8192 independent 128-body worlds")


def test_stdout_carries_one_json_line_whatever_libraries_print():
    """bench.py points file descriptor 1 at stderr for the run and writes its line to a duplicate of the real stdout: what
    NCCL (its version banner) or anything else prints to stdout ends up on stderr."""
    code = ("import os, sys; sys.path.insert(0, %r); import bench; sys.stdout.flush(); bench._REAL_STDOUT = os.dup(1); os.dup2(2, 1); "
            "os.write(1, b'NCCL version 2.28.9+cuda12.9\\n'); print('noise from python'); sys.stdout.flush(); bench.emit_line({'a': 1})" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"a": 1}\n'
    assert "NCCL version" in r.stderr and "noise from python" in r.stderr
