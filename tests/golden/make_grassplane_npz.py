"""Converts the reference asset res/grassPlane.obj (data, not code; 159 vertices / 266 triangles, the terrain mesh
SURVEY.md section 8 f4 names) into tests/golden/grassplane_mesh.npz so that the GPU box, which has no
/root/reference, can use it.  Run in the build container:
    python tests/golden/make_grassplane_npz.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "rl-ode-physics_b200"))
from odeb200.scenes import load_obj  # noqa: E402

if __name__ == "__main__":
    v, t = load_obj("/root/reference/res/grassPlane.obj")
    assert v.shape == (159, 3) and t.shape == (266, 3), (v.shape, t.shape)
    np.savez_compressed(os.path.join(HERE, "grassplane_mesh.npz"), verts=v.astype(np.float32), tris=t.astype(np.int32))
    print("grassPlane:", v.shape, t.shape, "bbox", v.min(0), v.max(0))
