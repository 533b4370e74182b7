// TensorFlow custom-op shell for ADCME's `load_op_and_grad(libmfem, "fem_laplace_scalar")` on top of libadfem_cuda.so.
//
// Replaces deps/MFEM/FemLaplace1/FemLaplaceScalar.cpp of the reference: the op names, input / output names, dtypes and the Grad-op
// input convention ([grad of outputs..., outputs..., inputs...]) are exactly the registered ones (:14-33 there), so
// src/MFEM/MCore.jl:100-108 and BFGS! see no difference; the bodies call the C symbols of include/adfem_cuda.h instead of
// MFEM::FemLaplaceScalar_forward / _backward (FemLaplaceScalar.h:3-54).  The mesh is the library's global 2-D mesh set by
// init_nnfem_mesh (the reference's `mmesh` singleton).
//
// Build (needs TensorFlow headers, which this repository's image does not have — the file is the binding a maintainer adds):
//   g++ -std=c++14 -shared -fPIC FemLaplaceScalar.cpp -o libnnfem_mfem.so $(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_compile_flags()+tf.sysconfig.get_link_flags()))') \
//       -I../../include -L../../adfem.jl_b200/lib -ladfem_cuda -Wl,-rpath,'$ORIGIN'
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"

#include "adfem_cuda.h"

namespace tf = tensorflow;

REGISTER_OP("FemLaplaceScalar")
    .Input("kappa : double")
    .Output("indices : int64")
    .Output("vv : double")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      tf::shape_inference::ShapeHandle unused;
      TF_RETURN_IF_ERROR(c->WithRank(c->input(0), 1, &unused));
      c->set_output(0, c->Matrix(c->UnknownDim(), 2));
      c->set_output(1, c->Vector(c->UnknownDim()));
      return tf::Status::OK();
    });

REGISTER_OP("FemLaplaceScalarGrad")
    .Input("grad_vv : double")
    .Input("indices : int64")
    .Input("vv : double")
    .Input("kappa : double")
    .Output("grad_kappa : double");

namespace {

// number of COO slots the op emits: one d x d block per Gauss point (FemLaplaceScalar.h:3-26)
inline tf::int64 coo_slots() {
  const tf::int64 d = mfem_get_elem_ndof();
  return static_cast<tf::int64>(mfem_get_ngauss()) * d * d;
}

class FemLaplaceScalarKernel : public tf::OpKernel {
 public:
  using tf::OpKernel::OpKernel;
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& kappa = ctx->input(0);
    OP_REQUIRES(ctx, kappa.dims() == 1 && kappa.dim_size(0) == mfem_get_ngauss(),
                tf::errors::InvalidArgument("kappa must hold one value per Gauss point of the current mesh (", mfem_get_ngauss(), ")"));
    const tf::int64 N = coo_slots();
    tf::Tensor *indices = nullptr, *vv = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({N, 2}), &indices));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({N}), &vv));
    // host tensors in, host tensors out; the library copies to the GPU, runs k_coo_scalar_fwd + k_coo_indices and copies back
    FemLaplaceScalar_forward(reinterpret_cast<long long*>(indices->flat<tf::int64>().data()), vv->flat<double>().data(),
                             kappa.flat<double>().data());
  }
};

class FemLaplaceScalarGradKernel : public tf::OpKernel {
 public:
  using tf::OpKernel::OpKernel;
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor &grad_vv = ctx->input(0), &indices = ctx->input(1), &vv = ctx->input(2), &kappa = ctx->input(3);
    OP_REQUIRES(ctx, grad_vv.NumElements() == coo_slots(), tf::errors::InvalidArgument("grad_vv does not match the current mesh"));
    tf::Tensor* grad_kappa = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, kappa.shape(), &grad_kappa));
    FemLaplaceScalar_backward(grad_kappa->flat<double>().data(), grad_vv.flat<double>().data(),
                              reinterpret_cast<const long long*>(indices.flat<tf::int64>().data()), vv.flat<double>().data(),
                              kappa.flat<double>().data());
  }
};

}  // namespace

REGISTER_KERNEL_BUILDER(Name("FemLaplaceScalar").Device(tf::DEVICE_CPU), FemLaplaceScalarKernel);
REGISTER_KERNEL_BUILDER(Name("FemLaplaceScalarGrad").Device(tf::DEVICE_CPU), FemLaplaceScalarGradKernel);
// The same two classes serve the other scalar ops by swapping the called pair: ComputeFemMassMatrix1_forward/_backward
// ("ComputeFemMassMatrixMfem"), FemSourceScalar_forward/_backward ("FemSourceScalar"), ComputeFemStiffnessMatrixMfem_forward/_backward
// ("ComputeFemStiffnessMatrixMfem", N = ngauss * (2 d)^2) and the 3-D twins (…T).  For device-resident tensors register the kernels for
// DEVICE_GPU and call adfem_assemble_coo / adfem_assemble_coo_adjoint (include/adfem_cuda.h group 2) on
// ctx->eigen_gpu_device().stream(): nothing crosses PCIe.
