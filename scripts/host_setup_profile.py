#!/usr/bin/env python
"""Wall time of the host-side, mesh-static setup of one case on a HOST-ONLY handle (no GPU needed): mesh generator, adfem_mesh_create, the
symbolic CSR pattern and the two tile plans, with MD5 digests of every product so that two builds of the host code can be compared.
ADFEM_DEBUG_PLAN=1 adds the library's own phase times (stderr).

  python scripts/host_setup_profile.py CASE N      CASE: 2 (Mesh(N,N) P1) | 2g / 3 (jittered + renumbered P1, scalar / elasticity plans) |
                                                          4 / 4o (P2, renumbered / generator order) | 5 (Mesh3(N,N,N)) | 5g (jittered tets)
"""
import hashlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import adfem_jl_b200 as A
from adfem_jl_b200 import meshgen
case=sys.argv[1]; n=int(sys.argv[2])
t=time.perf_counter()
def lap(s):
    global t
    nw=time.perf_counter(); print("%-40s %.2fs"%(s,nw-t), flush=True); t=nw
def h(a): return hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()[:12]
if case=='4':
    c,e=meshgen.jitter_unstructured(n,n,1.0/n,seed=2,permute=True); lap("gen")
    m=A.Mesh(c,e,degree=2,host_only=True); nc=1
elif case=='4o':
    c,e=meshgen.jitter_unstructured(n,n,1.0/n,seed=2,permute=False); lap("gen")
    m=A.Mesh(c,e,degree=2,host_only=True); nc=1
elif case=='2':
    c,e=meshgen.tri_grid(n,n,1.0/n); lap("gen")
    m=A.Mesh(c,e,host_only=True); nc=1
elif case=='2g':
    c,e=meshgen.jitter_unstructured(n,n,1.0/n,seed=2,permute=True); lap("gen")
    m=A.Mesh(c,e,host_only=True); nc=1
elif case=='3':
    c,e=meshgen.jitter_unstructured(n,n,1.0/n,seed=2,permute=True); lap("gen")
    m=A.Mesh(c,e,host_only=True); nc=2
elif case=='5':
    m=A.Mesh3(n,n,n,1.0/n,host_only=True); nc=3
elif case=='5g':
    c,e=meshgen.tet_grid(n,n,n,1.0/n); rng=np.random.default_rng(1); c=c+rng.uniform(-0.1/n,0.1/n,c.shape)
    m=A.Mesh3(c,e,host_only=True); nc=1
lap("Mesh create")
rp,ci=m.csr_pattern(1); lap("csr_pattern"); print("  pattern", h(rp), h(ci))
for which,name in ((0,'fwd'),(1,'adj')):
    p=m.plan_array(which,nc,0,np.int64); lap(name+" plan"); b=m.plan_array(which,nc,1,np.uint8)
    print("  ",name,len(p),len(b),h(p),h(b)); t=time.perf_counter()
