# AdFemCUDA.jl — the Julia side of the drop-in boundary: `ccall` bindings of libadfem_cuda.so for AdFem.jl.
#
# No Julia toolchain exists in the image this repository is built in (`which julia` -> none), so this file has not been executed here;
# every symbol it binds is exported by the library (tests/test_abi_symbols.py) and exercised end to end by the Python ctypes twin
# (adfem.jl_b200/_lib.py, tests/test_gpu_parity.py::test_legacy_symbols_2d / _3d) with the same argument conventions.
#
# Three layers, matching INTEGRATION.md:
#   1. `use_cuda_backend!()`   re-points AdFem's LIBMFEM / LIBMFEM3 handles (src/AdFem.jl:31-33): every existing eager `ccall`
#                              (src/MFEM/MFEM.jl:95-106, src/MFEM/MCore.jl:112-120, src/MFEM3/MFEM.jl:58-75) then runs on the GPU unchanged.
#   2. `CudaMesh` + `compute_fem_*_matrix1(coef, ::CudaMesh)`   CSR mode: the mesh-static symbolic phase once, values only per call,
#                              host arrays in / `SparseMatrixCSC` out (same public signatures as src/MFEM/MCore.jl:100-120, 203-222, 31-66).
#   3. `DistMesh`              multi-GPU interface exchange (adfem_dist_*) for one process per GPU (MPI.jl / Distributed + NCCL.jl).
module AdFemCUDA

using SparseArrays
using Libdl

export use_cuda_backend!, CudaMesh, compute_fem_laplace_matrix1, compute_fem_mass_matrix1, compute_fem_stiffness_matrix,
       compute_fem_source_term1, laplace_matrix_adjoint, DistMesh, reduce_interface!, replicate_interface!

const LIB = Ref{String}(get(ENV, "ADFEM_CUDA_LIB", joinpath(@__DIR__, "..", "adfem.jl_b200", "lib", "libadfem_cuda.so")))

last_error() = unsafe_string(ccall((:adfem_last_error, LIB[]), Cstring, ()))
check(rc::Integer) = rc == 0 || error("libadfem_cuda: " * last_error())

"""
    use_cuda_backend!(AdFem)

Point `AdFem.LIBMFEM` and `AdFem.LIBMFEM3` at libadfem_cuda.so.  The library exports the reference's symbols with the reference's
signatures (include/adfem_cuda.h group 1), so `Mesh(...)`, `compute_fem_laplace_matrix1(kappa::Array, mesh)`, `compute_fem_source_term1`,
`compute_fem_stiffness_matrix`, the `pcl_*` helpers and the 3-D twins work without touching their Julia code.  Fails when no CUDA
device is usable (there is no CPU fallback).
"""
function use_cuda_backend!(adfem::Module)
    ccall((:adfem_device_count, LIB[]), Cint, ()) > 0 || error("libadfem_cuda: no usable CUDA device")
    Core.eval(adfem, :(LIBMFEM = $(LIB[])))
    Core.eval(adfem, :(LIBMFEM3 = $(LIB[])))
    Core.eval(adfem, :(libmfem = $(LIB[])))       # ADCME's load_op_and_grad dlopens this path for the TF op shells (integration/tf_ops)
    nothing
end

# ------------------------------------------------------------------------------------------------------------------------------
# 2. handle API, host buffers (adfem_assemble_csr_host streams chunks over three CUDA streams on structured meshes)
# ------------------------------------------------------------------------------------------------------------------------------
const OP_LAPLACE, OP_MASS, OP_STIFFNESS = Cint(0), Cint(1), Cint(2)

mutable struct CudaMesh
    handle::Ptr{Cvoid}
    dim::Int
    ndof::Int
    ngauss::Int
    rowptr::Vector{Int64}        # scalar pattern, 0-based
    colind::Vector{Int32}
    function CudaMesh(nodes::Array{Float64,2}, elems::Array{Int64,2}; degree::Int = 1, order::Int = -1, lorder::Int = -1)
        dim = size(nodes, 2)
        c = Array(nodes')                                  # dim x nv, column-major == interleaved xyz per vertex
        e = Int32.(Array(elems') .- 1)                     # 0-based, (dim+1) x ne
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:adfem_mesh_create, LIB[]), Cint,
                    (Ref{Ptr{Cvoid}}, Cint, Ptr{Cdouble}, Cint, Cint, Ptr{Cint}, Cint, Cint, Cint, Cint, Cint),
                    h, dim, c, dim, size(nodes, 1), e, size(elems, 1), order, degree, lorder, 0))
        info(w) = Int(ccall((:adfem_mesh_info, LIB[]), Clonglong, (Ptr{Cvoid}, Cint), h[], w))
        ndof, ngauss = info(3), info(4)                    # ADFEM_INFO_NDOF, ADFEM_INFO_NGAUSS
        nnz = ccall((:adfem_csr_nnz, LIB[]), Clonglong, (Ptr{Cvoid}, Cint), h[], 1)
        nnz >= 0 || error("libadfem_cuda: " * last_error())
        rowptr = zeros(Int64, ndof + 1); colind = zeros(Int32, nnz)
        check(ccall((:adfem_csr_pattern, LIB[]), Cint, (Ptr{Cvoid}, Cint, Ptr{Clonglong}, Ptr{Cint}), h[], 1, rowptr, colind))
        m = new(h[], dim, ndof, ngauss, rowptr, colind)
        finalizer(x -> ccall((:adfem_mesh_destroy, LIB[]), Cvoid, (Ptr{Cvoid},), x.handle), m)
        m
    end
end
CudaMesh(mesh; kw...) = CudaMesh(mesh.nodes, mesh.elems; degree = (string(mesh.elem_type) == "P2" ? 2 : 1), order = mesh.order, lorder = mesh.lorder, kw...)

# CSR of K (0-based) -> SparseMatrixCSC: the CSR arrays of K are the CSC arrays of K'
function csr_to_csc(n::Int, rowptr::Vector{Int64}, colind::Vector{Int32}, vals::Vector{Float64})
    copy(SparseMatrixCSC(n, n, rowptr .+ 1, Int64.(colind) .+ 1, vals)')
end

function assemble_scalar(op::Cint, coef::Array{Float64,1}, cm::CudaMesh)
    @assert length(coef) == cm.ngauss
    vals = zeros(length(cm.colind))
    check(ccall((:adfem_assemble_csr_host, LIB[]), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}), cm.handle, op, coef, vals))
    csr_to_csc(cm.ndof, cm.rowptr, cm.colind, vals)
end

"compute_fem_laplace_matrix1(kappa, mesh) of src/MFEM/MCore.jl:112-120, CSR mode (no COO triplets, no `sparse(i, j, v)` re-summation)"
compute_fem_laplace_matrix1(kappa::Array{Float64,1}, cm::CudaMesh) = assemble_scalar(OP_LAPLACE, kappa, cm)
"compute_fem_mass_matrix1(rho, mesh) of src/MFEM/MCore.jl:203-222"
compute_fem_mass_matrix1(rho::Array{Float64,1}, cm::CudaMesh) = assemble_scalar(OP_MASS, rho, cm)

"""
    compute_fem_stiffness_matrix(hmat, mesh)   (src/MFEM/MCore.jl:31-66)

`hmat`: ngauss x 3 x 3 (2-D) or ngauss x 6 x 6 (3-D extension), row-major per Gauss point like the reference op's input.  Component-
blocked dofs (u then v [then w]); entry (r + a n, c + b n) of the block CSR lives at nc (a nnz + rowptr[r]) + b len(r) + j.
"""
function compute_fem_stiffness_matrix(hmat::Array{Float64,3}, cm::CudaMesh)
    nc = cm.dim
    h = permutedims(hmat, [3, 2, 1])                       # Julia column-major -> row-major per point
    nnz = length(cm.colind)
    vals = zeros(nc * nc * nnz)
    check(ccall((:adfem_assemble_csr_host, LIB[]), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}), cm.handle, OP_STIFFNESS, h, vals))
    n = cm.ndof
    I = Int64[]; J = Int64[]; V = Float64[]
    sizehint!(I, length(vals)); sizehint!(J, length(vals)); sizehint!(V, length(vals))
    for a in 0:nc-1, r in 0:n-1
        rs, len = cm.rowptr[r+1], cm.rowptr[r+2] - cm.rowptr[r+1]
        for b in 0:nc-1, j in 0:len-1
            push!(I, r + a * n + 1); push!(J, cm.colind[rs+j+1] + b * n + 1); push!(V, vals[nc*(a*nnz+rs)+b*len+j+1])
        end
    end
    sparse(I, J, V, nc * n, nc * n)
end

"gradient of a scalar loss with respect to kappa given dL/dK on the CSR pattern (the adjoint BFGS! needs; FemLaplaceScalar.h:28-54)"
function laplace_matrix_adjoint(dvals::Array{Float64,1}, cm::CudaMesh; op::Cint = OP_LAPLACE)
    @assert length(dvals) == length(cm.colind)
    g = zeros(cm.ngauss)
    check(ccall((:adfem_assemble_csr_adjoint_host, LIB[]), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}), cm.handle, op, dvals, g))
    g
end

"compute_fem_source_term1(f, mesh) of src/MFEM/MCore.jl:300-313 through the reference's own eager symbol (global 2-D mesh)"
function compute_fem_source_term1(f::Array{Float64,1}, ndof::Int)
    rhs = zeros(ndof)
    ccall((:FemSourceScalar_forward_Julia, LIB[]), Cvoid, (Ptr{Cdouble}, Ptr{Cdouble}), rhs, f)
    rhs
end

# ------------------------------------------------------------------------------------------------------------------------------
# 3. multi-GPU: one Julia process per GPU; `comm` is an ncclComm_t (NCCL.jl: `comm.handle`) or C_NULL to let the library create one
# ------------------------------------------------------------------------------------------------------------------------------
mutable struct DistMesh
    handle::Ptr{Cvoid}
    mesh::CudaMesh
end

"128-byte ncclUniqueId: call on rank 0, broadcast with MPI.jl / Distributed, pass to `nccl_comm` on every rank"
function nccl_unique_id()
    id = zeros(UInt8, 128)
    check(ccall((:adfem_dist_nccl_unique_id, LIB[]), Cint, (Ptr{UInt8},), id))
    id
end
function nccl_comm(id::Vector{UInt8}, rank::Int, world::Int)
    c = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:adfem_dist_comm_create, LIB[]), Cint, (Ref{Ptr{Cvoid}}, Ptr{UInt8}, Cint, Cint), c, id, rank, world))
    c[]
end

"""
    DistMesh(cm, comm, rank, world, send_counts, send_pos, recv_counts, recv_pos; max_ncomp = cm.dim)

Mesh-static exchange lists (0-based positions in the scalar CSR of `cm`; `recv_pos` -1 = ghost column), see include/adfem_cuda.h
group 3 and adfem.jl_b200/dist.py (`Partition`) for how they are derived from the global vertex ids.
"""
function DistMesh(cm::CudaMesh, comm::Ptr{Cvoid}, rank::Int, world::Int, send_counts::Vector{Int64}, send_pos::Vector{Int64},
                  recv_counts::Vector{Int64}, recv_pos::Vector{Int64}; max_ncomp::Int = cm.dim)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:adfem_dist_create, LIB[]), Cint,
                (Ref{Ptr{Cvoid}}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Clonglong}, Ptr{Clonglong}, Ptr{Clonglong}, Ptr{Clonglong}),
                h, cm.handle, comm, rank, world, max_ncomp, send_counts, send_pos, recv_counts, recv_pos))
    d = new_dist(h[], cm)
    d
end
new_dist(h, cm) = (d = DistMesh(h, cm); finalizer(x -> ccall((:adfem_dist_destroy, LIB[]), Cvoid, (Ptr{Cvoid},), x.handle), d); d)

"forward: sum the partial interface rows at their owners, in place, on `stream` (device pointers: CUDA.jl `pointer(::CuArray)`)"
reduce_interface!(d::DistMesh, ncomp::Int, vals_dev::Ptr{Cvoid}, ghost_dev::Ptr{Cvoid}, stream::Ptr{Cvoid}) =
    check(ccall((:adfem_dist_reduce, LIB[]), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), d.handle, ncomp, vals_dev, ghost_dev, stream))
"adjoint: owners send dL/dK of the interface entries back to every contributor; the element-level adjoint then runs rank-local"
replicate_interface!(d::DistMesh, ncomp::Int, dvals_dev::Ptr{Cvoid}, dghost_dev::Ptr{Cvoid}, stream::Ptr{Cvoid}) =
    check(ccall((:adfem_dist_replicate, LIB[]), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), d.handle, ncomp, dvals_dev, dghost_dev, stream))

end # module
