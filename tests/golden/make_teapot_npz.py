"""Converts the reference asset res/teapot.obj (data, not code; 4884 vertices / 8884 triangles, the
mesh BASELINE.json config 2 names) into tests/golden/teapot_mesh.npz so that the GPU box, which has no
/root/reference, can run the trimesh config.  Run in the build container:
    python tests/golden/make_teapot_npz.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "rl-ode-physics_b200"))
from odeb200.scenes import load_obj  # noqa: E402

if __name__ == "__main__":
    v, t = load_obj("/root/reference/res/teapot.obj")
    assert v.shape == (4884, 3) and t.shape == (8884, 3), (v.shape, t.shape)
    np.savez_compressed(os.path.join(HERE, "teapot_mesh.npz"), verts=v.astype(np.float32), tris=t.astype(np.int32))
    print("teapot:", v.shape, t.shape, "bbox", v.min(0), v.max(0))
