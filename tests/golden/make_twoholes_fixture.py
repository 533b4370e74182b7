"""Generates tests/golden/twoholes_large.npz from the reference's bundled mesh
(/root/reference/deps/MFEM/MeshData/twoholes_large.stl, BASELINE config 1).

Run in the build container only (the GPU box has no /root/reference).  The STL is read the way
meshio 4.2 reads ASCII STL for `Mesh(filename)` (src/MFEM/MUtils.jl:32-50): one triangle per facet,
bit-identical vertices merged keeping first-appearance order, z dropped.
"""
import os
import sys

import numpy as np

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/deps/MFEM/MeshData/twoholes_large.stl"
pts, index, tris, cur = [], {}, [], []
for line in open(src):
    t = line.split()
    if len(t) == 4 and t[0] == "vertex":
        key = (float(t[1]), float(t[2]), float(t[3]))
        if key not in index:
            index[key] = len(pts)
            pts.append(key)
        cur.append(index[key])
        if len(cur) == 3:
            tris.append(cur)
            cur = []
pts = np.array(pts)
assert np.all(pts[:, 2] == 0)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "twoholes_large.npz")
np.savez_compressed(out, nodes=pts[:, :2], elems=np.array(tris, dtype=np.int64))
print(out, pts.shape, len(tris))
