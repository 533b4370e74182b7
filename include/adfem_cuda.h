/* libadfem_cuda.so — C ABI of the B200-native differentiable FEM assembly path.
 *
 * Two groups of entry points:
 *
 *  (1) LEGACY symbols: byte-for-byte the signatures the reference's Julia side binds with `ccall`
 *      (and the TF op shells call) for this path.  Host pointers in, host pointers out, one global
 *      2-D mesh and one global 3-D mesh per process (reference quirk Q10), `void` returns.  They are
 *      thin wrappers that copy to/from the device around the kernels of group (2).
 *
 *  (2) HANDLE API (`adfem_*`): mesh handles, the mesh-static symbolic phase, and device-pointer
 *      assembly / adjoint calls on a caller-supplied CUDA stream.  Every function returns 0 on
 *      success and a non-zero code otherwise; `adfem_last_error()` gives the message.
 *
 * There is no CPU fallback: every compute entry point fails (non-zero / error message) when no CUDA
 * device is usable.  Index conventions are the reference's own and are stated per function.
 *
 * Reference paths below are relative to the kailaix/AdFem.jl checkout.
 */
#ifndef ADFEM_CUDA_H
#define ADFEM_CUDA_H

#ifdef __cplusplus
extern "C" {
#endif

/* ============================ (1) legacy symbols ============================================== */

/* deps/MFEM/API.cpp:4-14.  vertices: 3*nv xyz-interleaved (z ignored); element_indices: 3*ne, 0-based.
 * Returns malloc'ed edges[2*nedges] (edges[i] = min vertex + 1, edges[nedges+i] = max vertex + 1);
 * the CALLER frees it (Julia: unsafe_wrap(..., own=true), src/MFEM/MFEM.jl:99).  degree -1 (BDM1) is
 * out of scope and returns NULL. */
long long* init_nnfem_mesh(double* vertices, int num_vertices, int* element_indices, int num_elements,
                           int order, int lorder, int degree, long long* nedges);
int  mfem_get_ngauss(void);                                  /* API.cpp:17-19 */
void mfem_get_gauss(double* x, double* y);                   /* API.cpp:21-24 */
void mfem_get_gauss_weights(double* w);                      /* API.cpp:26-34, element-major */
void mfem_get_area(double* a);                               /* API.cpp:36-39 (Heron) */
int  mfem_get_elem_ndof(void);                               /* API.cpp:41-43 */
int  mfem_get_ndof(void);                                    /* API.cpp:45-47 */
void mfem_get_connectivity(long long* conn);                 /* API.cpp:49-56, ne*d row-major, 1-based */
void mfem_get_element_to_vertices(long long* elems);         /* API.cpp:58-65, column-major, 1-based */
int  get_LineIntegralN(void);                                /* deps/MFEM/Common.cpp:350-352 */
void get_LineIntegralPnW(double* p, double* w);              /* deps/MFEM/Common.cpp:354-362 */

/* deps/MFEM3/API.cpp:4-67.  vertices: 3*nv; element_indices: 4*ne, 0-based. */
long long* init_nnfem_mesh3(double* vertices, int num_vertices, int* element_indices, int num_elements,
                            int order, int degree, long long* nedges);
int  mfem_get_ngauss3(void);
void mfem_get_gauss3(double* x, double* y, double* z);
void mfem_get_gauss_weights3(double* w);
void mfem_get_volume3(double* v);
int  mfem_get_elem_ndof3(void);
int  mfem_get_ndof3(void);
void mfem_get_connectivity3(long long* conn);
void mfem_get_element_to_vertices3(long long* elems);

/* Physics-constrained-learning Jacobians (src/pcl.jl), host pointers, dense column-major, CALLER-ZEROED like Julia's zeros(...):
 *  pcl_FemLaplaceScalar_Jacobian: H[G x G*d*d], H[t + slot*G] = d vv[slot] / d kappa[t]     (deps/MFEM/FemLaplace1/FemLaplaceScalar.h:65-92)
 *  pcl_ImposeDirichlet: J[sN x outdof], J[k + s*sN] = 1 for the s-th kept slot k; indices 1-based column-major, bd 1-based
 *                                                                                          (deps/MFEM/ImposeDirichlet/ImposeDirichlet.h:98-112) */
void pcl_FemLaplaceScalar_Jacobian(double* H);
void pcl_ImposeDirichlet(double* J, const long long* indices, const long long* bd, int bdN, int sN);

/* Eager op entry points (host pointers).  indices: 2*N interleaved (row, col), 0-based int64. */
void FemLaplaceScalar_forward_Julia(long long* indices, double* vv, const double* kappa);      /* deps/MFEM/FemLaplace1/FemLaplaceScalar.h:59-61 */
void FemSourceScalar_forward_Julia(double* rhs, const double* f);                              /* deps/MFEM/FemSource1/FemSourceScalar.h:35-37 (adds into rhs) */
void ComputeFemStiffnessMatrixMfem_forward_Julia(long long* indices, double* vv, const double* hmat); /* deps/MFEM/ComputeFemStiffnessMatrixMfem/ComputeFemStiffnessMatrixMfem.h:84-88 */
void FemLaplaceScalarT_forward_Julia(long long* indices, double* vv, const double* kappa);     /* deps/MFEM3/FemLaplace1/FemLaplaceScalarT.h:61-63 */
void FemSourceScalarT_forward_Julia(double* rhs, const double* f);                             /* deps/MFEM3/FemSource/FemSourceScalarT.h:35-37 */

/* The bodies the TF op shells call (namespace MFEM in the reference), exported with C linkage so a
 * rebuilt op shell can link them.  Host pointers; same argument order as the reference functions. */
void FemLaplaceScalar_forward(long long* indices, double* vv, const double* kappa);            /* FemLaplaceScalar.h:3-26 */
void FemLaplaceScalar_backward(double* grad_kappa, const double* grad_vv, const long long* indices,
                               const double* vv, const double* kappa);                        /* FemLaplaceScalar.h:28-54 */
void ComputeFemMassMatrix1_forward(long long* indices, double* vv, const double* rho);         /* deps/MFEM/ComputeFemMassMatrix1/ComputeFemMassMatrixMfem.h:4-29 */
void ComputeFemMassMatrix1_backward(double* grad_rho, const double* grad_vv, const double* vv,
                                    const double* rho);                                       /* ComputeFemMassMatrixMfem.h:31-56 (overwrites grad_rho) */
void ComputeFemStiffnessMatrixMfem_forward(long long* indices, double* vv, const double* hmat);/* ComputeFemStiffnessMatrixMfem.h:4-48 */
void ComputeFemStiffnessMatrixMfem_backward(double* grad_hmat, const double* grad_vv);         /* ComputeFemStiffnessMatrixMfem.h:50-81 */
void FemSourceScalar_forward(double* rhs, const double* f);                                    /* FemSourceScalar.h:4-15 (adds into rhs) */
void FemSourceScalar_backward(double* grad_f, const double* grad_rhs, const double* rhs,
                              const double* f);                                               /* FemSourceScalar.h:17-32 */
void FemLaplaceScalarT_forward(long long* indices, double* vv, const double* kappa);           /* FemLaplaceScalarT.h:3-27 */
void FemLaplaceScalarT_backward(double* grad_kappa, const double* grad_vv, const long long* indices,
                                const double* vv, const double* kappa);                       /* FemLaplaceScalarT.h:29-56 */
void ComputeFemMassMatrixMfemT_forward(long long* indices, double* vv, const double* rho);     /* deps/MFEM3/ComputeFemMassMatrixMfem3/ComputeFemMassMatrixMfemT.h:4-27, N = ne*d*d */
void ComputeFemMassMatrixMfemT_backward(double* grad_rho, const double* grad_vv);              /* extension: the reference's Grad body is empty (…T.cpp:136-140) */
void FemSourceScalarT_forward(double* rhs, const double* f);                                   /* FemSourceScalarT.h:4-15 */
void FemSourceScalarT_backward(double* grad_f, const double* grad_rhs, const double* rhs,
                               const double* f);                                              /* FemSourceScalarT.h:17-32 */

/* Gauss-point operators and matrix-free terms on the same element tables.  Host pointers, the global 2-D mesh
 * (`…T` / `…3`: the global 3-D mesh); same argument order as the reference bodies.  Where the reference ACCUMULATES into a
 * caller-zeroed array (every `+=` below) these wrappers add into the host array as well. */
void FemToGaussPointsMfem_Julia(double* out, const double* u);                                    /* deps/MFEM/FemToGaussPoints/FemToGaussPointsMfem.h:38-49: out[G] = */
void FemToGaussPointsMfem_forward(double* out, const double* u);                                  /* …h:6-18 */
void FemToGaussPointsMfem_backward(double* grad_u, const double* grad_out, const double* out, const double* u);   /* …h:20-35: grad_u[node] += */
void DofToGaussPointsMfem_forward_Julia(double* out, const double* u);                            /* deps/MFEM/DofToGaussPoints/DofToGaussPointsMfem.h:37-39 */
void DofToGaussPointsMfem_forward(double* out, const double* u);                                  /* …h:6-18: out[G] = */
void DofToGaussPointsMfem_backward(double* grad_u, const double* grad_out, const double* out, const double* u);   /* …h:20-34: grad_u[ndof] += */
void FemGradMfem_forward(double* out, const double* u);                                           /* deps/MFEM/FemGrad/FemGradMfem.h:8-23: out[2G] = (du/dx, du/dy) interleaved */
void FemGradMfem_backward(double* grad_u, const double* grad_out, const double* out, const double* u);            /* …h:25-42: grad_u[ndof] += */
void EvalStrainOnGaussPts_forward_Julia(double* epsilon, const double* u);                        /* deps/MFEM/EvalStrainOnGaussPtsMfem/EvalStrainOnGaussPts.h:32-34 */
void EvalStrainOnGaussPts_forward(double* epsilon, const double* u);                              /* …h:4-16: epsilon[3G] += (exx, eyy, gxy); u[2 ndof] component-blocked */
void EvalStrainOnGaussPts_backward(double* grad_u, const double* grad_epsilon);                   /* …h:18-29: grad_u[2 ndof] += */
void ComputeStrainEnergyTermMfem_forward_Julia(double* out, const double* sigma);                 /* deps/MFEM/ComputeStrainEnergyTermMfem/ComputeStrainEnergyTermMfem.h:38-40 */
void ComputeStrainEnergyTermMfem_forward(double* out, const double* sigma);                       /* …h:4-19: out[2 ndof] += int sigma : eps(v); sigma[3G] = (s11, s22, s12) */
void ComputeStrainEnergyTermMfem_backward(double* grad_sigma, const double* grad_out);            /* …h:21-35: grad_sigma[3G] += */
void ComputeLaplaceTermMfem_forward_Julia(double* out, const double* nu, const double* u);        /* deps/MFEM/ComputeLaplaceTermMfem/ComputeLaplaceTermMfem.h:42-44 */
void ComputeLaplaceTermMfem_forward(double* out, const double* nu, const double* u);              /* …h:4-17: out[ndof] += int nu grad u . grad v */
void ComputeLaplaceTermMfem_backward(double* grad_nu, double* grad_u, const double* grad_out, const double* out,
                                     const double* nu, const double* u);                          /* …h:19-39: both += */
void ComputeLaplaceTermMfem3_forward_Julia(double* out, const double* nu, const double* u);       /* deps/MFEM3/ComputeLaplaceTermMfem/ComputeLaplaceTermMfemT.h:48-50 */
void ComputeLaplaceTermMfemT_forward(double* out, const double* nu, const double* u);                 /* …T.h:4-22 */
void ComputeLaplaceTermMfemT_backward(double* grad_nu, double* grad_u, const double* grad_out, const double* out,
                                      const double* nu, const double* u);                         /* …T.h:24-46 */
/* deps/MFEM/PlaneStrainAndStress/PlaneStrainAndStress.h:5-78 (the op's mode 0 / mode 1): out[9N] row-major 3x3 per point; the
 * backward OVERWRITES grad_nu[N], grad_E[N] (note the reference's argument order: grad_nu first). */
void PlaneStrainMatrix_forward(double* out, const double* E, const double* nu, int N);
void PlaneStrainMatrix_backward(double* grad_nu, double* grad_E, const double* grad_out, const double* E, const double* nu, int N);
void PlaneStressMatrix_forward(double* out, const double* E, const double* nu, int N);
void PlaneStressMatrix_backward(double* grad_nu, double* grad_E, const double* grad_out, const double* E, const double* nu, int N);

/* ============================ (2) handle API ================================================= */
/* Threading / streams: a mesh handle owns scratch that its calls share (the Gauss-summed tangents of the P1 elasticity kernels, the
 * staging buffers of the *_host calls, the side stream of the optional z-chunk pipeline).  Calls on ONE handle must therefore be issued
 * on one stream at a time, or be ordered by events; different handles are independent.  Every call is asynchronous on the stream it is
 * given (the first call of an operator also builds its mesh-static plan, which synchronises once). */
typedef struct adfem_mesh adfem_mesh;

const char* adfem_last_error(void);
int adfem_device_count(void);          /* 0 when no usable CUDA device (never a fallback) */

enum { ADFEM_HOST_ONLY = 1 };          /* flags: build host tables/plans only (inspection, CPU unit tests) */
enum { ADFEM_OP_LAPLACE = 0, ADFEM_OP_MASS = 1, ADFEM_OP_STIFFNESS = 2 };
enum { ADFEM_INFO_DIM = 0, ADFEM_INFO_NV, ADFEM_INFO_NE, ADFEM_INFO_NDOF, ADFEM_INFO_NGAUSS, ADFEM_INFO_ELEM_NDOF,
       ADFEM_INFO_NEDGES, ADFEM_INFO_GAUSS_PER_ELEM, ADFEM_INFO_NNZ_SCALAR, ADFEM_INFO_TILES_FWD, ADFEM_INFO_TILES_ADJ,
       ADFEM_INFO_PLAN_BYTES,
       ADFEM_INFO_STRUCTURED /* 1: the mesh is the structured triangulation Mesh(m,n,h) v1 on rectilinear nodes and the scalar CSR
                                operators use the index-free kernels of csrc/tri_grid.cuh (option "structured" = 0 disables);
                                2: the mesh is the structured tetrahedral grid Mesh3(n,n,l,h) on rectilinear nodes (csrc/tet_node.cuh / tet_grid.cuh,
                                used by the elasticity kernels under option "structured_elasticity");
                                3: structured connectivity of Mesh(m,n,h) on NON-rectilinear (mapped / jittered) node positions: the scalar CSR
                                operators, the source term, P1 elasticity and the scatter-type Gauss-point operators / Laplace term use the
                                index-free kernels with node positions read from the coordinate array (MAPPED instantiations) */ };

/* Replaces init_nnfem_mesh / init_nnfem_mesh3 (deps/MFEM/API.cpp:4, deps/MFEM3/API.cpp:4) without the
 * process-global singleton.  dim = 2|3; vertices: nv rows of `vertex_stride` doubles (first `dim` used);
 * elems: ne*(dim+1), 0-based; degree 1|2; order/lorder as in the reference (-1 = reference defaults,
 * src/MFEM/MFEM.jl:71-85, src/MFEM3/MFEM.jl:49-55). */
int  adfem_mesh_create(adfem_mesh** out, int dim, const double* vertices, int vertex_stride, int nv,
                       const int* elems, int ne, int order, int degree, int lorder, int flags);
void adfem_mesh_destroy(adfem_mesh* m);
long long adfem_mesh_info(const adfem_mesh* m, int what);
int adfem_mesh_edges(const adfem_mesh* m, long long* edges);                  /* 2*nedges, layout of init_nnfem_mesh */
int adfem_mesh_connectivity(const adfem_mesh* m, long long* conn);            /* as mfem_get_connectivity */
int adfem_mesh_element_to_vertices(const adfem_mesh* m, long long* elems);    /* as mfem_get_element_to_vertices */
int adfem_mesh_gauss(const adfem_mesh* m, double* xyz);                       /* dim blocks of ngauss (column-major) */
int adfem_mesh_gauss_weights(const adfem_mesh* m, double* w);
int adfem_mesh_measure(const adfem_mesh* m, double* a);                       /* Heron area (2-D) / volume (3-D) */
int adfem_set_option(adfem_mesh* m, const char* key, long long value);        /* "rows_per_tile", "elems_per_tile", "adjoint_tiled", "host_threads", "smem_budget", "tile_threads", "pipeline", "coef_prefetch", "coef_presum", "structured", "structured_pattern", "structured_elasticity", "structured_tet_scalar", "row_gather", "grid_rows", "grid_limit", "area_formula_csr", "area_formula_coo" */

/* Mesh-static symbolic phase (what Julia's sparse() / TF's sparse ops redo on every call downstream of the
 * reference's COO, src/MFEM/MCore.jl:118-119).  ncomp = 1: scalar operators, n = ndof.  ncomp = dim: the
 * elasticity operator, n = dim*ndof with component-blocked dofs (ComputeFemStiffnessMatrixMfem.h:31-34). */
int adfem_symbolic(adfem_mesh* m);
long long adfem_csr_nnz(adfem_mesh* m, int ncomp);
int adfem_csr_pattern(adfem_mesh* m, int ncomp, long long* rowptr /* n+1 */, int* colind /* nnz */);   /* host out, 0-based */
int adfem_slot_to_nnz(adfem_mesh* m, unsigned int* slot_nnz /* ne*d*d, host out */);
/* DEVICE pointers of the scalar pattern (rowptr int64[ndof+1], colind int32[nnz], 0-based), owned by the handle: together with the values
 * written by adfem_assemble_csr this is a device-resident CSR matrix that a GPU solver (cuDSS / cuSPARSE / AMGX) can take without a host
 * round trip — the hand-off SURVEY 8(f) ranks first after the assembly path itself. */
int adfem_csr_pattern_device(adfem_mesh* m, const long long** rowptr, const int** colind);


/* Inspection of the mesh-static tile plans (ADFEM_HOST_ONLY handles only; used by the CPU unit tests that
 * replay a plan against the oracle).  which_plan 0 = forward row tiles, 1 = adjoint element tiles; array_id 0 =
 * blob_ptr (int64[ntiles+1] byte offsets), 1 = the concatenated per-tile blobs (bytes; layout documented in
 * adfem.jl_b200/csrc/plan.h — it is exactly what one TMA bulk copy brings into shared memory per CTA).
 * Returns the element count (and copies when out != NULL), -1 on error. */
long long adfem_plan_array(adfem_mesh* m, int which_plan, int ncomp, int array_id, void* out);

/* CSR-mode assembly: coef and vals are DEVICE pointers, stream is a cudaStream_t.
 * op LAPLACE/MASS: coef[ngauss]; STIFFNESS: coef[ngauss*ns*ns], ns = 3 (2-D) | 6 (3-D Voigt xx,yy,zz,yz,xz,xy).
 * vals[nnz] in the order of adfem_csr_pattern: the canonical CSR of the reference op's COO output. */
int adfem_assemble_csr(adfem_mesh* m, int op, const double* coef, double* vals, void* stream);
/* reverse mode: dvals[nnz] = d loss / d vals  ->  grad_coef (same shape as coef) */
int adfem_assemble_csr_adjoint(adfem_mesh* m, int op, const double* dvals, double* grad_coef, void* stream);

/* COO-compatible mode: exactly the reference ops' outputs (N = ngauss*D*D slots; 3-D mass: ne*d*d). */
long long adfem_coo_nslots(const adfem_mesh* m, int op);
int adfem_coo_indices(adfem_mesh* m, int op, long long* indices /* device, 2*N */, void* stream);
int adfem_assemble_coo(adfem_mesh* m, int op, const double* coef, double* vv, void* stream);
int adfem_assemble_coo_adjoint(adfem_mesh* m, int op, const double* grad_vv, double* grad_coef, void* stream);

/* Source term (FemSourceScalar*): rhs[ndof] is OVERWRITTEN (no pre-zeroing needed). */
int adfem_source(adfem_mesh* m, const double* f, double* rhs, void* stream);
int adfem_source_adjoint(adfem_mesh* m, const double* grad_rhs, double* grad_f, void* stream);

/* Gauss-point operators: dof <-> Gauss-point transfers on the element tables.  DEVICE pointers; outputs are
 * OVERWRITTEN (no pre-zeroing).  Layouts are the reference's: Gauss-point arrays element-major with NQ interleaved values per point
 * ((e*g + k)*NQ + i), dof vectors component-blocked (dof + c*ndof).
 *   kind                      forward: in -> out                                    reference op
 *   ADFEM_GP_FEM_TO_GAUSS     u[nv]        -> [G]       P1 shapes, vertex values     FemToGaussPointsMfem
 *   ADFEM_GP_DOF_TO_GAUSS     u[ndof]      -> [G]       all shape functions          DofToGaussPointsMfem
 *   ADFEM_GP_GRAD             u[ndof]      -> [G*dim]   physical gradient            FemGradMfem
 *   ADFEM_GP_STRAIN           u[dim*ndof]  -> [G*ns]    (exx, eyy, gxy); 3-D Voigt   EvalStrainOnGaussPts (3-D: extension)
 *   ADFEM_GP_STRAIN_ENERGY    sigma[G*ns]  -> [dim*ndof] int sigma : eps(v)          ComputeStrainEnergyTermMfem
 * adfem_gauss_op_len(m, kind, 0|1) = input | output length.  _adjoint maps grad_out (output-shaped) to grad_in (input-shaped). */
enum { ADFEM_GP_FEM_TO_GAUSS = 0, ADFEM_GP_DOF_TO_GAUSS = 1, ADFEM_GP_GRAD = 2, ADFEM_GP_STRAIN = 3, ADFEM_GP_STRAIN_ENERGY = 4 };
long long adfem_gauss_op_len(const adfem_mesh* m, int kind, int output);
int adfem_gauss_op(adfem_mesh* m, int kind, const double* in, double* out, void* stream);
int adfem_gauss_op_adjoint(adfem_mesh* m, int kind, const double* grad_out, double* grad_in, void* stream);
/* ComputeLaplaceTermMfem / ...T: out[ndof] = int nu grad u . grad v (nu[G], u[ndof]); the adjoint gives grad_nu[G] and grad_u[ndof]
 * (either may be NULL to skip it). */
int adfem_laplace_term(adfem_mesh* m, const double* nu, const double* u, double* out, void* stream);
int adfem_laplace_term_adjoint(adfem_mesh* m, const double* nu, const double* u, const double* grad_out, double* grad_nu, double* grad_u,
                               void* stream);
/* PlaneStrainAndStress op: mode 0 = PlaneStrainMatrix, 1 = PlaneStressMatrix (the reference's names and formulas,
 * deps/MFEM/PlaneStrainAndStress/PlaneStrainAndStress.h); E[n], nu[n] -> H[9n]; the gradient overwrites grad_E[n], grad_nu[n]. */
int adfem_plane_matrix(int mode, long long n, const double* E, const double* nu, double* H, void* stream);
int adfem_plane_matrix_grad(int mode, long long n, const double* E, const double* nu, const double* grad_H, double* grad_E, double* grad_nu,
                            void* stream);

/* Fused pre-step + assembly (P1 triangles): CSR values of the elasticity operator with H = plane matrix(E, nu) per Gauss point, and the
 * gradient with respect to E and nu, WITHOUT materialising H (16 B instead of 72 B of coefficients per Gauss point).  Same values as
 * adfem_plane_matrix followed by adfem_assemble_csr(ADFEM_OP_STIFFNESS); E[G], nu[G], vals / dvals [4*nnz]. */
int adfem_assemble_csr_plane(adfem_mesh* m, int mode, const double* E, const double* nu, double* vals, void* stream);
int adfem_assemble_csr_plane_adjoint(adfem_mesh* m, int mode, const double* E, const double* nu, const double* dvals, double* grad_E, double* grad_nu,
                                     void* stream);

/* Algebraic Dirichlet conditions on a COO matrix — the ImposeDirichlet op (deps/MFEM/ImposeDirichlet/ImposeDirichlet.h:27-93,
 * op signature ImposeDirichlet.cpp:14-57).  All pointers are DEVICE pointers.  indices: sN x 2 (row, col) 0-based; bd: bdN
 * boundary dofs, 1-BASED like the reference (ImposeDirichlet.h:32); duplicates in bd: the last bdval wins.  Output: the kept
 * slots in input order followed by one (b, b, 1.0) per boundary dof in ascending order; orhs[N].  The output length is data
 * dependent: call _count first (as the op shell runs forward before allocating, ImposeDirichlet.cpp:104-110). */
long long adfem_impose_dirichlet_count(const long long* indices, long long sN, const long long* bd, long long bdN, long long N, void* stream);
int adfem_impose_dirichlet(const long long* indices, const double* vv, long long sN, const long long* bd, const double* bdval,
                           long long bdN, const double* rhs, long long N, long long* oindices, double* ov, double* orhs, void* stream);
/* ImposeDirichletGrad (ImposeDirichlet.h:63-93): grad_vv[sN], grad_rhs[N], grad_bdval[bdN] are overwritten. */
int adfem_impose_dirichlet_grad(const double* grad_ov, const double* grad_orhs, const long long* indices, const double* vv,
                                long long sN, const long long* bd, const double* bdval, long long bdN, long long N,
                                double* grad_vv, double* grad_rhs, double* grad_bdval, void* stream);

/* DirichletBd — the op "dirichlet_bd" (deps/DirichletBd/DirichletBd.h:8-60, DirichletBd.cpp:19-33) behind
 * fem_impose_Dirichlet_boundary_condition_experimental / fem_impose_coupled_Dirichlet_boundary_condition (src/InvCore.jl:6-23): COO
 * triplets (ii, jj, vv)[N] of a two-component operator on the m x n grid, boundary NODES bd[bdn] (int32 like the op input, same index
 * base as ii / jj); the boundary dof set is bd and bd + (m+1)(n+1).  Output 1: the triplets with both indices free, in input order,
 * then one (b, b, 1.0) per boundary dof in ascending order; output 2: the triplets with a free row and a boundary column, the column
 * replaced by the 1-based position of that dof in [bd, bd + (m+1)(n+1)] (last duplicate wins).  All pointers are DEVICE pointers;
 * lengths are data dependent: call _count first.  _grad (backward(), DirichletBd.h:96-112) overwrites grad_vv[N]. */
int adfem_dirichlet_bd_count(const long long* ii, const long long* jj, long long N, const int* bd, int bdn, int m, int n, long long* n1, long long* n2,
                             void* stream);
int adfem_dirichlet_bd(const long long* ii, const long long* jj, const double* vv, long long N, const int* bd, int bdn, int m, int n, long long* ii1,
                       long long* jj1, double* vv1, long long* ii2, long long* jj2, double* vv2, void* stream);
int adfem_dirichlet_bd_grad(const long long* ii, const long long* jj, long long N, const int* bd, int bdn, int m, int n, const double* grad_vv1,
                            const double* grad_vv2, double* grad_vv, void* stream);

/* Device versions of the two PCL Jacobians: H / J are DEVICE pointers, caller-zeroed; indices 0-based interleaved like adfem_impose_dirichlet. */
int adfem_pcl_laplace_jacobian(adfem_mesh* m, double* H /* G x G*d*d column-major */, void* stream);
int adfem_pcl_impose_dirichlet(const long long* indices, long long sN, const long long* bd, long long bdN, long long N, double* J /* sN x S column-major */,
                               void* stream);

/* Structured-grid Q1 operators on an m x n grid of h x h cells (device pointers; ii/jj are 1-BASED int64 like the ops and may
 * both be NULL to skip the mesh-static indices).
 *  adfem_quad_stiffness1: UnivariateFemStiffness (deps/FemStiffness1/UnivariateFemStiffness.h:7-196) = compute_fem_stiffness_matrix1
 *    (src/InvCore.jl:67-76); hmat [4mn,2,2] (rank3=1) or [2,2] (rank3=0); 64mn slots.
 *  adfem_quad_elasticity: per_gauss=0 FemStiffness (deps/FemStiffness/FemStiffness.h:7-70, hmat [3,3], 64mn slots),
 *    per_gauss=1 SpatialFemStiffness (deps/SpatialFemStiffness/SpatialFemStiffness.h:7-81, hmat [4mn,3,3], 256mn slots).
 *  adfem_svt: SpatialVaryingTangentElastic (deps/SpatialVaryingTangentElastic/SpatialVaryingTangentElastic.h:1-70), type 1|2|3. */
int adfem_quad_stiffness1(const double* hmat, int rank3, int m, int n, double h, long long* ii, long long* jj, double* vv, void* stream);
int adfem_quad_stiffness1_grad(const double* grad_vv, int rank3, int m, int n, double h, double* grad_hmat, void* stream);
int adfem_quad_elasticity(const double* hmat, int per_gauss, int m, int n, double h, long long* ii, long long* jj, double* vv, void* stream);
int adfem_quad_elasticity_grad(const double* grad_vv, int per_gauss, int m, int n, double h, double* grad_hmat, void* stream);
int adfem_svt(const double* mu, long long m, long long n, int type, double* hmat, void* stream);
int adfem_svt_grad(const double* grad_hmat, long long m, long long n, int type, double* grad_mu, void* stream);

/* Structured-grid Q1 scalar siblings, device pointers.  Cell (i,j) = id j*m+i, Gauss point k = 2q+p at (pts[p], pts[q]),
 * coefficients [4mn] indexed 4*cell+k, slot (4*cell+k)*16 + 4a + b; ii/jj are 0-BASED int64 like these ops emit them (the Julia wrappers add
 * 1, src/InvCore.jl:368,448) and may both be NULL; m, n are int64 as in the ops' signatures.
 *  adfem_quad_scalar op 0: FemLaplace (deps/FemLaplace/FemLaplace.h:11-49) = compute_fem_laplace_matrix1(K, m, n, h);
 *                    op 1: FemMass (deps/FemMass/FemMass.h:10-43) = compute_fem_mass_matrix1(rho, m, n, h); 64mn slots.
 *  adfem_quad_source: FemSource (deps/FemSource/FemSource.h:8-26) = compute_fem_source_term1(f, m, n, h); rhs[(m+1)(n+1)] is OVERWRITTEN.
 *  The _grad calls overwrite grad_coef / grad_f [4mn] (FemLaplace.h:51-82, FemMass.h:45-79, FemSource.h:30-47). */
int adfem_quad_scalar(int op, const double* coef, long long m, long long n, double h, long long* ii, long long* jj, double* vv, void* stream);
int adfem_quad_scalar_grad(int op, const double* grad_vv, long long m, long long n, double h, double* grad_coef, void* stream);
int adfem_quad_source(const double* f, long long m, long long n, double h, double* rhs, void* stream);
int adfem_quad_source_grad(const double* grad_rhs, long long m, long long n, double h, double* grad_f, void* stream);

/* SpatialVaryingTangentElastic fused into UnivariateFemStiffness: the same 64mn slots (and 1-based ii / jj) as
 * adfem_svt followed by adfem_quad_stiffness1(rank3 = 1), without materialising the 4mn x 2 x 2 tensor; mu[4mn * type], grad_mu likewise. */
int adfem_quad_stiffness1_svt(const double* mu, int type, int m, int n, double h, long long* ii, long long* jj, double* vv, void* stream);
int adfem_quad_stiffness1_svt_grad(const double* grad_vv, int type, int m, int n, double h, double* grad_mu, void* stream);

/* Host-buffer convenience calls (synchronous; H2D + kernel + D2H).  Used for end-to-end timing. */
int adfem_assemble_csr_host(adfem_mesh* m, int op, const double* coef_host, double* vals_host);
int adfem_assemble_csr_adjoint_host(adfem_mesh* m, int op, const double* dvals_host, double* grad_coef_host);

/* ============================ (3) multi-GPU interface exchange ================================
 * SURVEY 8(e).  The reference has no parallel path (no file:line to mirror); this group is what a Julia / C++ / Python host binds to run
 * the assembly path on element blocks, one process per GPU.  The host computes the mesh-static lists once (which CSR entries of rows it
 * does not own go to which owner; for every entry it will receive, the position of the same (row, col) in its own pattern or -1 when it
 * does not hold the column) and hands them over; every exchange is then pack kernel -> one ncclGroup of ncclSend/ncclRecv -> unpack
 * kernel on the caller's stream.  No atomics: an entry receiving from several ranks sums them in ascending source-rank order.
 * NCCL is bound at run time (dlopen libnccl.so.2). */
typedef struct adfem_dist adfem_dist;
int adfem_dist_nccl_unique_id(void* id128 /* out: 128-byte ncclUniqueId, call on one rank and broadcast it */);
int adfem_dist_comm_create(void** nccl_comm /* out: ncclComm_t */, const void* id128, int rank, int world);    /* ncclCommInitRank, current device */
int adfem_dist_comm_destroy(void* nccl_comm);
/* send_pos: positions (in the scalar CSR of `m`) of the entries this rank sends, concatenated by destination rank (send_counts[world]);
 * recv_pos: for the entries it receives, concatenated by source rank (recv_counts[world]), its own position of that entry or -1 (ghost
 * column).  max_ncomp: largest block size the handle will be used with (1 scalar, dim elasticity).  nccl_comm may be an existing
 * ncclComm_t of the host (NCCL.jl, torch) or one made by adfem_dist_comm_create; NULL is accepted for world == 1. */
int adfem_dist_create(adfem_dist** out, adfem_mesh* m, void* nccl_comm, int rank, int world, int max_ncomp, const long long* send_counts,
                      const long long* send_pos, const long long* recv_counts, const long long* recv_pos);
void adfem_dist_destroy(adfem_dist* d);
long long adfem_dist_info(const adfem_dist* d, int what /* 0 entries sent, 1 received, 2 ghost entries, 3 distinct destinations, 4 bytes per scalar exchange */);
/* forward: vals (layout of adfem_assemble_csr, ncomp^2 * nnz) — partial interface rows are summed into their owners in place; ghost_vals
 * [nghost * ncomp^2] (may be NULL) receives the blocks whose column the owner does not hold.  Rows a rank does not own keep their partial sums. */
int adfem_dist_reduce(adfem_dist* d, int ncomp, double* vals, double* ghost_vals, void* stream);
/* adjoint: owners send d loss / d K of the interface entries back (dghost: gradient of the ghost block or NULL = 0), so dvals becomes valid
 * on every entry this rank's elements contribute to; adfem_assemble_csr_adjoint then runs rank-local. */
int adfem_dist_replicate(adfem_dist* d, int ncomp, double* dvals, const double* dghost, void* stream);

#ifdef __cplusplus
}
#endif
#endif
