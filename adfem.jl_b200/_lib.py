"""ctypes binding of libadfem_cuda.so (include/adfem_cuda.h).

The library is built in-tree (adfem.jl_b200/lib/libadfem_cuda.so) by `build()`; loading fails loudly
when it is missing, and every compute entry point fails when no CUDA device is usable — there is no
CPU fallback anywhere in this package.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "lib", "libadfem_cuda.so")
_lib = None

c_dp = C.POINTER(C.c_double)
c_lp = C.POINTER(C.c_longlong)
c_ip = C.POINTER(C.c_int)

OP_LAPLACE, OP_MASS, OP_STIFFNESS = 0, 1, 2
HOST_ONLY = 1
(INFO_DIM, INFO_NV, INFO_NE, INFO_NDOF, INFO_NGAUSS, INFO_ELEM_NDOF, INFO_NEDGES, INFO_GAUSS_PER_ELEM, INFO_NNZ_SCALAR,
 INFO_TILES_FWD, INFO_TILES_ADJ, INFO_PLAN_BYTES, INFO_STRUCTURED) = range(13)


class AdfemError(RuntimeError):
    pass


def build(verbose=False):
    """Compile every CUDA source for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if out.returncode != 0:
        raise AdfemError("building libadfem_cuda.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)
    return SO_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise AdfemError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(there is no CPU fallback)")
        L = C.CDLL(SO_PATH)
        L.adfem_last_error.restype = C.c_char_p
        L.adfem_mesh_info.restype = C.c_longlong
        L.adfem_csr_nnz.restype = C.c_longlong
        L.adfem_coo_nslots.restype = C.c_longlong
        L.adfem_plan_array.restype = C.c_longlong
        L.adfem_gauss_op_len.restype = C.c_longlong
        L.adfem_impose_dirichlet_count.restype = C.c_longlong
        L.adfem_dist_info.restype = C.c_longlong
        L.adfem_dist_destroy.restype = None
        L.init_nnfem_mesh.restype = c_lp
        L.init_nnfem_mesh3.restype = c_lp
        L.adfem_mesh_info.argtypes = [C.c_void_p, C.c_int]
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise AdfemError(lib().adfem_last_error().decode())


def last_error():
    return lib().adfem_last_error().decode()
