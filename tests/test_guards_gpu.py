"""GPU: the library's own overrun check.  compute-sanitizer is not available on the B200 pool, so libode_b200 can put
256-byte guard bands around every device allocation (ODE_B200_DEBUG_GUARD=1, csrc/prims.cu) and read them back
(dCheckGuardsB200).  This test re-runs a cross-section of the GPU tests -- every kernel family: grid / per-env / sort-
and-sweep broadphase, all colliders incl. the triangle grid, colouring, the four solvers, the exact dWorldStep, spawn
patches, snapshot formats, halo / slab kernels -- in a child pytest with the guards on; tests/conftest.py checks the
bands after each of them."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

SELECTION = ("pairs_and_contacts_match_golden or tick_parity_with_oracle or island_solver_equals or surface_modes_parity_on_the_island "
             "or more_than_64 or capacity_overflow or default_dworldstep or falls_back_to_the_sweeps or boxes_rest_on_a_trimesh "
             "or spawned or incremental_ingestion or compact_snapshot or dcollide_outside or c_slab_driver or union_of_slab "
             "or env_broadphase_equals or forces_and_async")


def test_guard_bands_stay_intact_across_the_gpu_tests():
    env = dict(os.environ, ODE_B200_DEBUG_GUARD="1")
    files = [os.path.join(HERE, f) for f in ("test_parity_gpu.py", "test_api_gpu.py", "test_slabs_gpu.py")]
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider", "-k", SELECTION] + files,
                       env=env, capture_output=True, text=True, timeout=1500)
    tail = (r.stdout[-3000:] + "\n" + r.stderr[-3000:])
    assert r.returncode == 0, tail
    import re
    m = re.search(r"(\d+) passed", r.stdout)
    assert m and int(m.group(1)) >= 40 and "failed" not in r.stdout, tail  # 47 tests at the time of writing


def test_a_deliberate_overrun_is_caught():
    """The check itself (dGuardSelfTestB200): one byte past the end of a 1000-byte allocation is seen, a write that
    stays inside is not; with the guards off the calls say so."""
    code = r'''
import os, sys
sys.path.insert(0, os.path.join(%r, "..", "rl-ode-physics_b200"))
import odeb200
from odeb200 import scenes
L = odeb200.lib()
sc = scenes.pile_scene(8, 8, 4, seed=5, spacing=0.6)
w = odeb200.World(gravity=sc["gravity"])
w.load_scene(sc)
w.tick(1.0 / 60.0)
print("RESULT", L.dCheckGuardsB200(0), L.dGuardSelfTestB200(0), L.dGuardSelfTestB200(1), L.dGuardSelfTestB200(200), L.dCheckGuardsB200(0))
''' % HERE
    for flag, want in (("1", "RESULT 0 0 1 1 0"), ("0", "RESULT -1 -1 -1 -1 -1")):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, ODE_B200_DEBUG_GUARD=flag), capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0 and want in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
