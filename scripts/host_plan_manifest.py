#!/usr/bin/env python
"""Digest manifest of the host symbolic phase (pattern, slot map, forward / adjoint tile plans) over a set of small meshes of every element
family, numbering and plan kind.  Two builds of csrc/plan.cpp / host_mesh.cpp are byte-compatible when their manifests are equal — how the
round-2 rewrite of the symbolic phase was checked (together with the adjacency checksum printed under ADFEM_DEBUG_PLAN=1 and
scripts/sass_diff.py for the device code).   python scripts/host_plan_manifest.py > manifest.txt"""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import adfem_jl_b200 as A
from adfem_jl_b200 import meshgen, _lib
def h(a): return hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()[:12]
def run(tag, m, ncs, opts=()):
    for k,v in opts: m.set_option(k,v)
    rp,ci=m.csr_pattern(1)
    L=_lib.lib()
    d=m.elem_ndof
    sl=np.zeros(m.nelem*d*d,dtype=np.uint32)
    _lib.check(L.adfem_slot_to_nnz(m.handle, sl.ctypes.data_as(C.c_void_p)))
    out=[tag,'pat',h(rp),h(ci),h(sl)]
    for nc in ncs:
        for which in (0,1):
            try:
                p=m.plan_array(which,nc,0,np.int64); b=m.plan_array(which,nc,1,np.uint8)
                out+= ['%d%s'%(nc,'fa'[which]),h(p),h(b)]
            except Exception as ex:
                out+= ['%d%s'%(nc,'fa'[which]),'ERR:'+str(ex)[:40]]
    print(' '.join(out),flush=True)
rng=np.random.default_rng(5)
c,e=meshgen.jitter_unstructured(400,400,1/400,seed=2,permute=True)
run('tri_p2_perm',A.Mesh(c,e,degree=2,host_only=True),(1,2))
run('tri_p1_perm',A.Mesh(c,e,host_only=True),(1,2))
run('tri_p1_perm_rows64',A.Mesh(c,e,host_only=True),(1,),(("rows_per_tile",64),("elems_per_tile",100)))
c,e=meshgen.jitter_unstructured(300,200,1/300,seed=3,permute=False)
run('tri_p2_gen',A.Mesh(c,e,degree=2,host_only=True),(1,))
run('tri_grid_p1',A.Mesh(600,500,1/600,host_only=True),(1,2))
run('tri_grid_p2',A.Mesh(200,150,1/200,degree=2,host_only=True),(1,))
run('tet_grid_p1',A.Mesh3(24,24,20,1/24,host_only=True),(1,3))
c,e=meshgen.tet_grid(14,14,12,1/14); c=c+rng.uniform(-0.1/14,0.1/14,c.shape); perm=rng.permutation(len(c)); inv=np.empty_like(perm); inv[perm]=np.arange(len(c))
c2=c[perm]; e2=inv[e].astype(e.dtype); e2=e2[rng.permutation(len(e2))]
run('tet_p1_perm',A.Mesh3(c2,e2,host_only=True),(1,3))
run('tet_p2_perm',A.Mesh3(c2,e2,degree=2,host_only=True),(1,3))
run('tiny',A.Mesh(3,3,0.5,host_only=True),(1,2))
run('tiny_p2',A.Mesh(1,1,0.5,degree=2,host_only=True),(1,))
