"""CPU test: libadfem_cuda.so loads without a GPU and exports every function include/adfem_cuda.h declares."""
import os
import re

import adfem_jl_b200 as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "adfem_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    L = A._lib.lib()
    names = declared_functions()
    assert len(names) > 60 and "adfem_assemble_csr" in names and "FemLaplaceScalar_forward_Julia" in names
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_reference_side_bindings_name_real_symbols():
    """julia/AdFemCUDA.jl (ccall) and integration/tf_ops/FemLaplaceScalar.cpp (the TF op shell) cannot be executed here (no Julia, no
    TensorFlow): check at least that every library symbol they bind is exported, and with the argument count the header declares."""
    L = A._lib.lib()
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "adfem_cuda.h")).read(), flags=re.S)
    nargs = {}
    for name, args in re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(([^;{}]*)\)\s*;", hdr):
        nargs[name] = 0 if args.strip() in ("", "void") else args.count(",") + 1
    jl = open(os.path.join(ROOT, "julia", "AdFemCUDA.jl")).read()
    calls = re.findall(r"ccall\(\(:(\w+), LIB\[\]\),\s*\w+,\s*\(([^)]*)\)", jl)
    assert len(calls) >= 15
    for name, argtypes in calls:
        assert hasattr(L, name), name
        n = len([a for a in argtypes.split(",") if a.strip()])
        assert n == nargs[name], (name, n, nargs[name])
    cpp = open(os.path.join(ROOT, "integration", "tf_ops", "FemLaplaceScalar.cpp")).read()
    for name in ("FemLaplaceScalar_forward", "FemLaplaceScalar_backward", "mfem_get_ngauss", "mfem_get_elem_ndof"):
        assert name + "(" in cpp and hasattr(L, name)
    assert 'REGISTER_OP("FemLaplaceScalar")' in cpp and 'REGISTER_OP("FemLaplaceScalarGrad")' in cpp      # the names ADCME's load_op_and_grad looks up


def test_error_reporting_without_gpu():
    L = A._lib.lib()
    if L.adfem_device_count() == 0:
        import ctypes as C
        import numpy as np
        h = C.c_void_p()
        c = np.zeros((3, 2)); c[1, 0] = 1; c[2, 1] = 1
        e = np.array([[0, 1, 2]], dtype=np.int32)
        rc = L.adfem_mesh_create(C.byref(h), 2, c.ctypes.data_as(A._lib.c_dp), 2, 3, e.ctypes.data_as(A._lib.c_ip), 1, -1, 1, -1, 0)
        assert rc != 0 and "no CPU fallback" in A._lib.last_error()
        assert L.adfem_quad_stiffness1(None, 0, 2, 2, C.c_double(1.0), None, None, None, None) != 0
    # bad arguments are reported, not crashed on
    import numpy as np
    import pytest
    with pytest.raises(A.AdfemError):
        A.Mesh(np.zeros((3, 2)), np.array([[0, 1, 5]]), host_only=True)        # vertex index out of range
    with pytest.raises(ValueError):
        A.Mesh(np.zeros((3, 2)), np.array([[0, 1, 2]]), degree=3, host_only=True)


def test_gauss_op_lengths_on_host_only_handles():
    """adfem_gauss_op_len is host arithmetic: input / output lengths of the five Gauss-point operators (include/adfem_cuda.h)."""
    import ctypes as C
    from adfem_jl_b200 import meshgen
    L = A._lib.lib()
    c, e = meshgen.jitter_unstructured(5, 4, 0.1, seed=0)
    for mesh, dim, ns in ((A.Mesh(c, e, degree=2, host_only=True), 2, 3), (A.Mesh3(2, 2, 2, 0.5, host_only=True), 3, 6)):
        G, n, nv = mesh.ngauss, mesh.ndof, mesh.nnode
        want = {0: (nv, G), 1: (n, G), 2: (n, G * dim), 3: (dim * n, G * ns), 4: (G * ns, dim * n)}
        for kind, (nin, nout) in want.items():
            assert L.adfem_gauss_op_len(mesh.handle, C.c_int(kind), C.c_int(0)) == nin
            assert L.adfem_gauss_op_len(mesh.handle, C.c_int(kind), C.c_int(1)) == nout
        assert L.adfem_gauss_op_len(mesh.handle, C.c_int(7), C.c_int(0)) == -1
        # compute entry points refuse host-only handles (no CPU path)
        assert L.adfem_gauss_op(mesh.handle, C.c_int(1), None, None, None) != 0 and "no CPU fallback" in A._lib.last_error()
        assert L.adfem_laplace_term(mesh.handle, None, None, None, None) != 0
        assert L.adfem_assemble_csr_plane(mesh.handle, C.c_int(1), None, None, None, None) != 0
