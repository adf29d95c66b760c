// TensorFlow custom-op shim over the C ABI of libfast_rnnt_b200.so.
//
// This is the file a maintainer of Samsung/tf-fast-rnnt builds in place of
// tf_fast_rnnt/python/csrc/tf_fast_rnnt_op.cc (and links against
// libfast_rnnt_b200.so instead of tf_mutual_information_core).  It keeps the two
// ops of the reference — "FastRNNTLoss" and "Cummin", same names, dtypes and
// arity (tf_fast_rnnt_op.cc:27-38) — so tf_fast_rnnt.mutual_information_recursion
// and tf_fast_rnnt.cummin keep working unchanged, and adds fused ops for the rest
// of the hot path.  Differences from the reference shim, on purpose:
//   * no cudaStreamSynchronize, no host->device copy (tf_fast_rnnt_op.cc:105-113);
//   * px_grad has the shape of px (op.cc:84 always allocated [B,S,T+1]);
//   * shape functions are registered; non-zero C-ABI status -> InvalidArgument /
//     Internal instead of the reference's "status == 1" check.
//
// NOT compiled in this repository's container (TensorFlow headers are absent);
// build line for a machine with TensorFlow:
//   g++ -std=c++17 -shared -fPIC tf_fast_rnnt_b200_ops.cc -o _tf_fast_rnnt.so \
//       $(python -c "import tensorflow as tf; print(' '.join(tf.sysconfig.get_compile_flags()))") \
//       $(python -c "import tensorflow as tf; print(' '.join(tf.sysconfig.get_link_flags()))") \
//       -I../../include -L../lib -lfast_rnnt_b200 -Wl,-rpath,'$ORIGIN/../lib' -DGOOGLE_CUDA=1
#define EIGEN_USE_GPU
#include <cuda_runtime.h>

#include <type_traits>

#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"

#include "fast_rnnt_b200.h"

namespace tf = tensorflow;
using tf::shape_inference::InferenceContext;
using tf::shape_inference::ShapeHandle;

namespace {

inline void *StreamOf(tf::OpKernelContext *ctx) {
  return static_cast<void *>(ctx->eigen_device<Eigen::GpuDevice>().stream());
}

inline tf::Status FromFrn(int rc, const char *what) {
  if (rc == FRN_OK) return tf::OkStatus();
  if (rc == FRN_EINVAL) return tf::errors::InvalidArgument(what, ": ", frn_status_string(rc));
  return tf::errors::Internal(what, ": ", frn_status_string(rc), " (cudaError ", frn_last_cuda_error(), ")");
}

// workspace as an allocate_temp uint8 tensor (TF's allocator returns >= 256-byte aligned GPU memory)
inline tf::Status Workspace(tf::OpKernelContext *ctx, size_t bytes, tf::Tensor *ws) {
  return ctx->allocate_temp(tf::DT_UINT8, tf::TensorShape({static_cast<tf::int64>(bytes < 256 ? 256 : bytes)}), ws);
}

// Shapes of the pruned-loss family: logits [B,T,R,C], symbols [B,S], ranges [B,T,R], boundary [B,4]
// (the reference validates nothing, its Python asserts are commented out: rnnt_loss.py:158-171).
inline tf::Status CheckPrunedShapes(const tf::Tensor &lg, const tf::Tensor &sym, const tf::Tensor *rg,
                                    const tf::Tensor &bd) {
  if (lg.dims() != 4) return tf::errors::InvalidArgument("logits must be [B,T,s_range,C]");
  if (sym.dims() != 2 || sym.dim_size(0) != lg.dim_size(0)) return tf::errors::InvalidArgument("symbols must be [B,S]");
  if (bd.dims() != 2 || bd.dim_size(0) != lg.dim_size(0) || bd.dim_size(1) != 4)
    return tf::errors::InvalidArgument("boundary must be [B,4]");
  if (rg != nullptr) {
    if (rg->dims() != 3 || rg->dim_size(0) != lg.dim_size(0) || rg->dim_size(1) != lg.dim_size(1) ||
        rg->dim_size(2) != lg.dim_size(2))
      return tf::errors::InvalidArgument("ranges must be [B,T,s_range], matching logits");
    if (lg.dim_size(2) > sym.dim_size(1) + 1) return tf::errors::InvalidArgument("s_range must be <= S + 1");
  }
  return tf::OkStatus();
}

}  // namespace

// ---------------------------------------------------------------------------
// FastRNNTLoss  (reference: tf_fast_rnnt_op.cc:27-34, 48-133)
// ---------------------------------------------------------------------------
REGISTER_OP("FastRNNTLoss")
    .Input("px: float32")
    .Input("py: float32")
    .Input("boundary: int32")
    .Input("calc_gradients: bool")
    .Output("ans: float32")
    .Output("px_grad: float32")
    .Output("py_grad: float32")
    .SetShapeFn([](InferenceContext *c) {
      ShapeHandle px, py;
      TF_RETURN_IF_ERROR(c->WithRank(c->input(0), 3, &px));
      TF_RETURN_IF_ERROR(c->WithRank(c->input(1), 3, &py));
      c->set_output(0, c->Vector(c->Dim(px, 0)));
      c->set_output(1, px);
      c->set_output(2, py);
      return tf::OkStatus();
    });

class FastRNNTLossOp : public tf::OpKernel {
 public:
  explicit FastRNNTLossOp(tf::OpKernelConstruction *ctx) : tf::OpKernel(ctx) {}
  void Compute(tf::OpKernelContext *ctx) override {
    const tf::Tensor &px = ctx->input(0), &py = ctx->input(1), &boundary = ctx->input(2);
    const bool calc = ctx->input(3).scalar<bool>()();
    OP_REQUIRES(ctx, px.dims() == 3 && py.dims() == 3, tf::errors::InvalidArgument("px, py must be 3-D"));
    const int B = px.dim_size(0), S = px.dim_size(1), T1 = px.dim_size(2), T = py.dim_size(2);
    OP_REQUIRES(ctx, py.dim_size(0) == B && py.dim_size(1) == S + 1 && (T1 == T || T1 == T + 1),
                tf::errors::InvalidArgument("px must be [B,S,T] or [B,S,T+1], py [B,S+1,T]"));
    OP_REQUIRES(ctx, boundary.dims() == 2 && boundary.dim_size(0) == B && boundary.dim_size(1) == 4,
                tf::errors::InvalidArgument("boundary must be [B,4]"));
    tf::Tensor *ans = nullptr, *gx = nullptr, *gy = nullptr, ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({B}), &ans));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, px.shape(), &gx));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, py.shape(), &gy));
    const size_t bytes = frn_mi_workspace_bytes(B, S, T, T1);
    OP_REQUIRES_OK(ctx, Workspace(ctx, bytes, &ws));
    if (!calc) {  // reference leaves the grads uninitialised (defect D6): zero them instead
      cudaMemsetAsync(gx->flat<float>().data(), 0, gx->NumElements() * sizeof(float),
                      static_cast<cudaStream_t>(StreamOf(ctx)));
      cudaMemsetAsync(gy->flat<float>().data(), 0, gy->NumElements() * sizeof(float),
                      static_cast<cudaStream_t>(StreamOf(ctx)));
    }
    OP_REQUIRES_OK(ctx, FromFrn(frn_mi_fwd_bwd(px.flat<float>().data(), py.flat<float>().data(),
                                               boundary.flat<tf::int32>().data(), B, S, T, T1, calc ? 1 : 0,
                                               ans->flat<float>().data(), gx->flat<float>().data(),
                                               gy->flat<float>().data(), ws.flat<tf::uint8>().data(), bytes,
                                               StreamOf(ctx)),
                                "FastRNNTLoss"));
  }
};
REGISTER_KERNEL_BUILDER(Name("FastRNNTLoss").Device(tf::DEVICE_GPU).HostMemory("calc_gradients"), FastRNNTLossOp);

// ---------------------------------------------------------------------------
// Cummin  (reference: tf_fast_rnnt_op.cc:36-38, 135-165)
// ---------------------------------------------------------------------------
REGISTER_OP("Cummin").Input("in: int32").Output("out: int32").SetShapeFn([](InferenceContext *c) {
  c->set_output(0, c->input(0));
  return tf::OkStatus();
});

class CumminOp : public tf::OpKernel {
 public:
  explicit CumminOp(tf::OpKernelConstruction *ctx) : tf::OpKernel(ctx) {}
  void Compute(tf::OpKernelContext *ctx) override {
    const tf::Tensor &in = ctx->input(0);
    OP_REQUIRES(ctx, in.dims() == 2, tf::errors::InvalidArgument("Cummin expects a matrix"));
    tf::Tensor *out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, in.shape(), &out));
    OP_REQUIRES_OK(ctx, FromFrn(frn_cummin(in.flat<tf::int32>().data(), out->flat<tf::int32>().data(),
                                           in.dim_size(0), in.dim_size(1), StreamOf(ctx)),
                                "Cummin"));
  }
};
REGISTER_KERNEL_BUILDER(Name("Cummin").Device(tf::DEVICE_GPU), CumminOp);

// ---------------------------------------------------------------------------
// FastRnntSimpleLoss: rnnt_loss_simple / rnnt_loss_smoothed fused
// (rnnt_loss.py:225-338, 1369-1494), reduction "none"; scores = -loss.
// ---------------------------------------------------------------------------
REGISTER_OP("FastRnntSimpleLoss")
    .Input("lm: T")
    .Input("am: T")
    .Input("symbols: int32")
    .Input("boundary: int32")
    .Attr("T: {float32, bfloat16, half} = DT_FLOAT")
    .Attr("termination_symbol: int")
    .Attr("rnnt_type: int = 0")
    .Attr("smoothed: bool = false")
    .Attr("lm_only_scale: float = 0.0")
    .Attr("am_only_scale: float = 0.0")
    .Attr("delay_penalty: float = 0.0")
    .Attr("calc_gradients: bool = true")
    .Output("scores: float32")
    .Output("px_grad: float32")
    .Output("py_grad: float32")
    .SetShapeFn([](InferenceContext *c) {
      c->set_output(0, c->Vector(c->Dim(c->input(0), 0)));
      c->set_output(1, c->UnknownShapeOfRank(3));
      c->set_output(2, c->UnknownShapeOfRank(3));
      return tf::OkStatus();
    });

// lm / am of type T are consumed as they are (frn_simple_loss_lp: bf16 / fp16 need no tf.cast and no float32 copy;
// the reference's op is float32-only, op.cc:28-34).  Scores and occupation counts are float32.
template <typename DT>
class FastRnntSimpleLossOp : public tf::OpKernel {
 public:
  explicit FastRnntSimpleLossOp(tf::OpKernelConstruction *ctx) : tf::OpKernel(ctx) {
    OP_REQUIRES_OK(ctx, ctx->GetAttr("termination_symbol", &term_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("rnnt_type", &type_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("smoothed", &smoothed_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("lm_only_scale", &lms_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("am_only_scale", &ams_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("delay_penalty", &dp_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("calc_gradients", &calc_));
  }
  void Compute(tf::OpKernelContext *ctx) override {
    const tf::Tensor &lm = ctx->input(0), &am = ctx->input(1), &sym = ctx->input(2), &bd = ctx->input(3);
    OP_REQUIRES(ctx, lm.dims() == 3 && am.dims() == 3 && sym.dims() == 2, tf::errors::InvalidArgument("bad ranks"));
    const int B = am.dim_size(0), T = am.dim_size(1), C = am.dim_size(2), S = lm.dim_size(1) - 1;
    const int T1 = type_ == FRN_REGULAR ? T + 1 : T;
    tf::Tensor *scores = nullptr, *gx = nullptr, *gy = nullptr, ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({B}), &scores));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({B, S, T1}), &gx));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, tf::TensorShape({B, S + 1, T}), &gy));
    const size_t bytes = frn_simple_loss_workspace_bytes(B, S, T, C);
    OP_REQUIRES_OK(ctx, Workspace(ctx, bytes, &ws));
    const int dtype = std::is_same<DT, float>::value ? FRN_F32 : (std::is_same<DT, tf::bfloat16>::value ? FRN_BF16 : FRN_F16);
    OP_REQUIRES_OK(ctx, FromFrn(frn_simple_loss_lp(lm.flat<DT>().data(), am.flat<DT>().data(), dtype,
                                                   sym.flat<tf::int32>().data(), bd.flat<tf::int32>().data(), B, S, T,
                                                   C, term_, type_, smoothed_ ? 1 : 0, lms_, ams_, nullptr, dp_,
                                                   calc_ ? 1 : 0, scores->flat<float>().data(),
                                                   gx->flat<float>().data(), gy->flat<float>().data(),
                                                   ws.flat<tf::uint8>().data(), bytes, StreamOf(ctx)),
                                "FastRnntSimpleLoss"));
  }

 private:
  int term_, type_;
  bool smoothed_, calc_;
  float lms_, ams_, dp_;
};
REGISTER_KERNEL_BUILDER(Name("FastRnntSimpleLoss").Device(tf::DEVICE_GPU).TypeConstraint<float>("T"),
                        FastRnntSimpleLossOp<float>);
REGISTER_KERNEL_BUILDER(Name("FastRnntSimpleLoss").Device(tf::DEVICE_GPU).TypeConstraint<tf::bfloat16>("T"),
                        FastRnntSimpleLossOp<tf::bfloat16>);
REGISTER_KERNEL_BUILDER(Name("FastRnntSimpleLoss").Device(tf::DEVICE_GPU).TypeConstraint<Eigen::half>("T"),
                        FastRnntSimpleLossOp<Eigen::half>);

// gradient of FastRnntSimpleLoss w.r.t. lm and am (A9: what TF autodiff derives through
// rnnt_loss.py:175-221 / 1266-1365 once _RNNTLossGrad, __init__.py:154-162, has supplied the
// occupation counts); registered from Python with RegisterGradient("FastRnntSimpleLoss").
REGISTER_OP("FastRnntSimpleLossGrad")
    .Input("lm: float32")
    .Input("am: float32")
    .Input("symbols: int32")
    .Input("boundary: int32")
    .Input("px_grad: float32")
    .Input("py_grad: float32")
    .Input("scores_grad: float32")
    .Attr("termination_symbol: int")
    .Attr("rnnt_type: int = 0")
    .Attr("smoothed: bool = false")
    .Attr("lm_only_scale: float = 0.0")
    .Attr("am_only_scale: float = 0.0")
    .Output("lm_grad: float32")
    .Output("am_grad: float32")
    .SetShapeFn([](InferenceContext *c) {
      c->set_output(0, c->input(0));
      c->set_output(1, c->input(1));
      return tf::OkStatus();
    });

class FastRnntSimpleLossGradOp : public tf::OpKernel {
 public:
  explicit FastRnntSimpleLossGradOp(tf::OpKernelConstruction *ctx) : tf::OpKernel(ctx) {
    OP_REQUIRES_OK(ctx, ctx->GetAttr("termination_symbol", &term_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("rnnt_type", &type_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("smoothed", &smoothed_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("lm_only_scale", &lms_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("am_only_scale", &ams_));
  }
  void Compute(tf::OpKernelContext *ctx) override {
    const tf::Tensor &lm = ctx->input(0), &am = ctx->input(1), &sym = ctx->input(2), &bd = ctx->input(3);
    const tf::Tensor &gx = ctx->input(4), &gy = ctx->input(5), &sg = ctx->input(6);
    const int B = am.dim_size(0), T = am.dim_size(1), C = am.dim_size(2), S = lm.dim_size(1) - 1;
    tf::Tensor *lm_g = nullptr, *am_g = nullptr, ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, lm.shape(), &lm_g));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, am.shape(), &am_g));
    const size_t bytes = frn_simple_loss_bwd_workspace_bytes(B, S, T, C);
    OP_REQUIRES_OK(ctx, Workspace(ctx, bytes, &ws));
    const int rc =
        smoothed_ ? frn_smoothed_loss_bwd(lm.flat<float>().data(), am.flat<float>().data(),
                                          sym.flat<tf::int32>().data(), bd.flat<tf::int32>().data(),
                                          gx.flat<float>().data(), gy.flat<float>().data(), sg.flat<float>().data(), B,
                                          S, T, C, term_, type_, lms_, ams_, am_g->flat<float>().data(),
                                          lm_g->flat<float>().data(), ws.flat<tf::uint8>().data(), bytes, StreamOf(ctx))
                  : frn_simple_loss_bwd(lm.flat<float>().data(), am.flat<float>().data(),
                                        sym.flat<tf::int32>().data(), bd.flat<tf::int32>().data(),
                                        gx.flat<float>().data(), gy.flat<float>().data(), sg.flat<float>().data(), B, S,
                                        T, C, term_, type_, am_g->flat<float>().data(), lm_g->flat<float>().data(),
                                        ws.flat<tf::uint8>().data(), bytes, StreamOf(ctx));
    OP_REQUIRES_OK(ctx, FromFrn(rc, "FastRnntSimpleLossGrad"));
  }

 private:
  int term_, type_;
  bool smoothed_;
  float lms_, ams_;
};
REGISTER_KERNEL_BUILDER(Name("FastRnntSimpleLossGrad").Device(tf::DEVICE_GPU), FastRnntSimpleLossGradOp);

// ---------------------------------------------------------------------------
// FastRnntPruneRanges (rnnt_loss.py:647-761) / FastRnntDoPruning (:763-812)
// ---------------------------------------------------------------------------
REGISTER_OP("FastRnntPruneRanges")
    .Input("px_grad: float32")
    .Input("py_grad: float32")
    .Input("boundary: int32")
    .Attr("s_range: int")
    .Output("ranges: int32")
    .SetShapeFn([](InferenceContext *c) {
      c->set_output(0, c->UnknownShapeOfRank(3));
      return tf::OkStatus();
    });

class FastRnntPruneRangesOp : public tf::OpKernel {
 public:
  explicit FastRnntPruneRangesOp(tf::OpKernelConstruction *ctx) : tf::OpKernel(ctx) {
    OP_REQUIRES_OK(ctx, ctx->GetAttr("s_range", &s_range_));
  }
  void Compute(tf::OpKernelContext *ctx) override {
    const tf::Tensor &gx = ctx->input(0), &gy = ctx->input(1), &bd = ctx->input(2);
    const int B = gx.dim_size(0), S = gx.dim_size(1), T1 = gx.dim_size(2), T = gy.dim_size(2);
    const int R = frn_prune_ranges_width(S, s_range_);
    tf::Tensor *ranges = nullptr, ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({B, T, R}), &ranges));
    const size_t bytes = frn_prune_ranges_workspace_bytes(B, T);
    OP_REQUIRES_OK(ctx, Workspace(ctx, bytes, &ws));
    OP_REQUIRES_OK(ctx, FromFrn(frn_prune_ranges(gx.flat<float>().data(), gy.flat<float>().data(),
                                                 bd.flat<tf::int32>().data(), B, S, T, T1, s_range_,
                                                 ranges->flat<tf::int32>().data(), ws.flat<tf::uint8>().data(),
                                                 bytes < 256 ? 256 : bytes, StreamOf(ctx)),
                                "FastRnntPruneRanges"));
  }

 private:
  int s_range_;
};
REGISTER_KERNEL_BUILDER(Name("FastRnntPruneRanges").Device(tf::DEVICE_GPU), FastRnntPruneRangesOp);

REGISTER_OP("FastRnntDoPruning")
    .Input("am: float32")
    .Input("lm: float32")
    .Input("ranges: int32")
    .Output("am_pruned: float32")
    .Output("lm_pruned: float32")
    .SetShapeFn([](InferenceContext *c) {
      c->set_output(0, c->UnknownShapeOfRank(4));
      c->set_output(1, c->UnknownShapeOfRank(4));
      return tf::OkStatus();
    });

class FastRnntDoPruningOp : public tf::OpKernel {
 public:
  explicit FastRnntDoPruningOp(tf::OpKernelConstruction *ctx) : tf::OpKernel(ctx) {}
  void Compute(tf::OpKernelContext *ctx) override {
    const tf::Tensor &am = ctx->input(0), &lm = ctx->input(1), &rg = ctx->input(2);
    const int B = am.dim_size(0), T = am.dim_size(1), C = am.dim_size(2), S = lm.dim_size(1) - 1, R = rg.dim_size(2);
    tf::Tensor *amp = nullptr, *lmp = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({B, T, R, C}), &amp));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({B, T, R, C}), &lmp));
    OP_REQUIRES_OK(ctx, FromFrn(frn_do_pruning(am.flat<float>().data(), lm.flat<float>().data(),
                                               rg.flat<tf::int32>().data(), B, S, T, R, C, amp->flat<float>().data(),
                                               lmp->flat<float>().data(), StreamOf(ctx)),
                                "FastRnntDoPruning"));
  }
};
REGISTER_KERNEL_BUILDER(Name("FastRnntDoPruning").Device(tf::DEVICE_GPU), FastRnntDoPruningOp);

// do_rnnt_pruning plus the additive joiner of the reference's tests (logits = am_pruned + lm_pruned,
// simple_rnnt_loss_test.py:120-125) in one pass; all three tensors are outputs.
REGISTER_OP("FastRnntDoPruningAddJoiner")
    .Input("am: float32")
    .Input("lm: float32")
    .Input("ranges: int32")
    .Output("am_pruned: float32")
    .Output("lm_pruned: float32")
    .Output("logits: float32")
    .SetShapeFn([](InferenceContext *c) {
      for (int i = 0; i < 3; ++i) c->set_output(i, c->UnknownShapeOfRank(4));
      return tf::OkStatus();
    });

class FastRnntDoPruningAddJoinerOp : public tf::OpKernel {
 public:
  explicit FastRnntDoPruningAddJoinerOp(tf::OpKernelConstruction *ctx) : tf::OpKernel(ctx) {}
  void Compute(tf::OpKernelContext *ctx) override {
    const tf::Tensor &am = ctx->input(0), &lm = ctx->input(1), &rg = ctx->input(2);
    const int B = am.dim_size(0), T = am.dim_size(1), C = am.dim_size(2), S = lm.dim_size(1) - 1, R = rg.dim_size(2);
    tf::Tensor *amp = nullptr, *lmp = nullptr, *lg = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({B, T, R, C}), &amp));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({B, T, R, C}), &lmp));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, tf::TensorShape({B, T, R, C}), &lg));
    OP_REQUIRES_OK(ctx, FromFrn(frn_do_pruning_add_joiner(am.flat<float>().data(), lm.flat<float>().data(),
                                                          rg.flat<tf::int32>().data(), B, S, T, R, C,
                                                          amp->flat<float>().data(), lmp->flat<float>().data(),
                                                          lg->flat<float>().data(), StreamOf(ctx)),
                                "FastRnntDoPruningAddJoiner"));
  }
};
REGISTER_KERNEL_BUILDER(Name("FastRnntDoPruningAddJoiner").Device(tf::DEVICE_GPU), FastRnntDoPruningAddJoinerOp);

// gradient of FastRnntDoPruning (registered from Python with RegisterGradient)
REGISTER_OP("FastRnntDoPruningGrad")
    .Input("am_pruned_grad: float32")
    .Input("lm_pruned_grad: float32")
    .Input("ranges: int32")
    .Attr("S: int")
    .Output("am_grad: float32")
    .Output("lm_grad: float32")
    .SetShapeFn([](InferenceContext *c) {
      c->set_output(0, c->UnknownShapeOfRank(3));
      c->set_output(1, c->UnknownShapeOfRank(3));
      return tf::OkStatus();
    });

class FastRnntDoPruningGradOp : public tf::OpKernel {
 public:
  explicit FastRnntDoPruningGradOp(tf::OpKernelConstruction *ctx) : tf::OpKernel(ctx) {
    OP_REQUIRES_OK(ctx, ctx->GetAttr("S", &S_));
  }
  void Compute(tf::OpKernelContext *ctx) override {
    const tf::Tensor &ga = ctx->input(0), &gl = ctx->input(1), &rg = ctx->input(2);
    const int B = ga.dim_size(0), T = ga.dim_size(1), R = ga.dim_size(2), C = ga.dim_size(3);
    tf::Tensor *am_g = nullptr, *lm_g = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({B, T, C}), &am_g));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({B, S_ + 1, C}), &lm_g));
    OP_REQUIRES_OK(ctx, FromFrn(frn_do_pruning_bwd(ga.flat<float>().data(), gl.flat<float>().data(),
                                                   rg.flat<tf::int32>().data(), B, S_, T, R, C,
                                                   am_g->flat<float>().data(), lm_g->flat<float>().data(),
                                                   StreamOf(ctx)),
                                "FastRnntDoPruningGrad"));
  }

 private:
  int S_;
};
REGISTER_KERNEL_BUILDER(Name("FastRnntDoPruningGrad").Device(tf::DEVICE_GPU), FastRnntDoPruningGradOp);

// ---------------------------------------------------------------------------
// FastRnntPrunedLoss: rnnt_loss_pruned fused with its logits gradient
// (rnnt_loss.py:1022-1130 + the autodiff chain through :942-1018).
// scores_grad is the upstream gradient w.r.t. scores (ones for d scores/d logits).
// ---------------------------------------------------------------------------
REGISTER_OP("FastRnntPrunedLoss")
    .Input("logits: T")
    .Input("symbols: int32")
    .Input("ranges: int32")
    .Input("boundary: int32")
    .Input("scores_grad: float32")
    .Attr("T: {float32, bfloat16} = DT_FLOAT")
    .Attr("termination_symbol: int")
    .Attr("rnnt_type: int = 0")
    .Attr("delay_penalty: float = 0.0")
    .Attr("with_logits_grad: bool = true")
    .Output("scores: float32")
    .Output("logits_grad: T")
    .SetShapeFn([](InferenceContext *c) {
      bool with_grad = true;
      TF_RETURN_IF_ERROR(c->GetAttr("with_logits_grad", &with_grad));
      c->set_output(0, c->Vector(c->Dim(c->input(0), 0)));
      c->set_output(1, with_grad ? c->input(0) : c->Vector(0));
      return tf::OkStatus();
    });

template <typename DT>
class FastRnntPrunedLossOp : public tf::OpKernel {
 public:
  explicit FastRnntPrunedLossOp(tf::OpKernelConstruction *ctx) : tf::OpKernel(ctx) {
    OP_REQUIRES_OK(ctx, ctx->GetAttr("termination_symbol", &term_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("rnnt_type", &type_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("delay_penalty", &dp_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("with_logits_grad", &with_grad_));
  }
  void Compute(tf::OpKernelContext *ctx) override {
    const tf::Tensor &lg = ctx->input(0), &sym = ctx->input(1), &rg = ctx->input(2), &bd = ctx->input(3),
                     &sg = ctx->input(4);
    OP_REQUIRES_OK(ctx, CheckPrunedShapes(lg, sym, &rg, bd));
    const int B = lg.dim_size(0), T = lg.dim_size(1), R = lg.dim_size(2), C = lg.dim_size(3), S = sym.dim_size(1);
    OP_REQUIRES(ctx, sg.dims() == 1 && sg.dim_size(0) == B, tf::errors::InvalidArgument("scores_grad must be [B]"));
    tf::Tensor *scores = nullptr, *grad = nullptr, ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({B}), &scores));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, with_grad_ ? lg.shape() : tf::TensorShape({0}), &grad));
    const size_t bytes = frn_pruned_loss_min_workspace_bytes(B, S, T, R, dp_);     // sized by the path that runs
    OP_REQUIRES_OK(ctx, Workspace(ctx, bytes, &ws));
    const int dtype = std::is_same<DT, float>::value ? FRN_F32 : FRN_BF16;
    OP_REQUIRES_OK(ctx, FromFrn(frn_pruned_loss(lg.flat<DT>().data(), dtype, sym.flat<tf::int32>().data(),
                                                rg.flat<tf::int32>().data(), bd.flat<tf::int32>().data(), B, S, T, R,
                                                C, term_, type_, dp_, sg.flat<float>().data(),
                                                scores->flat<float>().data(),
                                                with_grad_ ? static_cast<void *>(grad->flat<DT>().data()) : nullptr,
                                                ws.flat<tf::uint8>().data(), bytes, StreamOf(ctx)),
                                "FastRnntPrunedLoss"));
  }

 private:
  int term_, type_;
  float dp_;
  bool with_grad_;
};
REGISTER_KERNEL_BUILDER(Name("FastRnntPrunedLoss").Device(tf::DEVICE_GPU).TypeConstraint<float>("T"),
                        FastRnntPrunedLossOp<float>);
REGISTER_KERNEL_BUILDER(Name("FastRnntPrunedLoss").Device(tf::DEVICE_GPU).TypeConstraint<tf::bfloat16>("T"),
                        FastRnntPrunedLossOp<tf::bfloat16>);

// ---------------------------------------------------------------------------
// FastRnntJointLoss: rnnt_loss on the full joiner output [B,T,S+1,C]
// (rnnt_loss.py:340-551: get_rnnt_logprobs_joint + recursion), with its logits gradient.
// ---------------------------------------------------------------------------
REGISTER_OP("FastRnntJointLoss")
    .Input("logits: T")
    .Input("symbols: int32")
    .Input("boundary: int32")
    .Input("scores_grad: float32")
    .Attr("T: {float32, bfloat16} = DT_FLOAT")
    .Attr("termination_symbol: int")
    .Attr("rnnt_type: int = 0")
    .Attr("delay_penalty: float = 0.0")
    .Attr("with_logits_grad: bool = true")
    .Output("scores: float32")
    .Output("logits_grad: T")
    .SetShapeFn([](InferenceContext *c) {
      c->set_output(0, c->Vector(c->Dim(c->input(0), 0)));
      c->set_output(1, c->input(0));
      return tf::OkStatus();
    });

template <typename DT>
class FastRnntJointLossOp : public tf::OpKernel {
 public:
  explicit FastRnntJointLossOp(tf::OpKernelConstruction *ctx) : tf::OpKernel(ctx) {
    OP_REQUIRES_OK(ctx, ctx->GetAttr("termination_symbol", &term_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("rnnt_type", &type_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("delay_penalty", &dp_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("with_logits_grad", &with_grad_));
  }
  void Compute(tf::OpKernelContext *ctx) override {
    const tf::Tensor &lg = ctx->input(0), &sym = ctx->input(1), &bd = ctx->input(2), &sg = ctx->input(3);
    OP_REQUIRES(ctx, lg.dims() == 4, tf::errors::InvalidArgument("logits must be [B,T,S+1,C]"));
    const int B = lg.dim_size(0), T = lg.dim_size(1), C = lg.dim_size(3), S = sym.dim_size(1);
    OP_REQUIRES(ctx, lg.dim_size(2) == S + 1, tf::errors::InvalidArgument("logits.shape[2] must be S+1"));
    tf::Tensor *scores = nullptr, *grad = nullptr, ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({B}), &scores));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, with_grad_ ? lg.shape() : tf::TensorShape({0}), &grad));
    const size_t bytes = frn_joint_loss_workspace_bytes(B, S, T);
    OP_REQUIRES_OK(ctx, Workspace(ctx, bytes, &ws));
    const int dtype = std::is_same<DT, float>::value ? FRN_F32 : FRN_BF16;
    OP_REQUIRES_OK(ctx, FromFrn(frn_joint_loss(lg.flat<DT>().data(), dtype, sym.flat<tf::int32>().data(),
                                               bd.flat<tf::int32>().data(), B, S, T, C, term_, type_, dp_,
                                               sg.flat<float>().data(), scores->flat<float>().data(),
                                               with_grad_ ? static_cast<void *>(grad->flat<DT>().data()) : nullptr,
                                               ws.flat<tf::uint8>().data(), bytes, StreamOf(ctx)),
                                "FastRnntJointLoss"));
  }

 private:
  int term_, type_;
  float dp_;
  bool with_grad_;
};
REGISTER_KERNEL_BUILDER(Name("FastRnntJointLoss").Device(tf::DEVICE_GPU).TypeConstraint<float>("T"),
                        FastRnntJointLossOp<float>);
REGISTER_KERNEL_BUILDER(Name("FastRnntJointLoss").Device(tf::DEVICE_GPU).TypeConstraint<tf::bfloat16>("T"),
                        FastRnntJointLossOp<tf::bfloat16>);

// ---------------------------------------------------------------------------
// FastRnntSimpleLogprobs: get_rnnt_logprobs / get_rnnt_logprobs_smoothed (rnnt_loss.py:63-223,
// 1132-1367) -> dense px [B,S,T1], py [B,S+1,T].  FastRnntPrunedLogprobs: get_rnnt_logprobs_pruned
// (:853-1020) -> the same from pruned joiner logits.  Forward only (the public log-prob
// functions; the losses above carry their own gradients).
// ---------------------------------------------------------------------------
REGISTER_OP("FastRnntSimpleLogprobs")
    .Input("lm: float32")
    .Input("am: float32")
    .Input("symbols: int32")
    .Input("boundary: int32")
    .Attr("termination_symbol: int")
    .Attr("rnnt_type: int = 0")
    .Attr("smoothed: bool = false")
    .Attr("lm_only_scale: float = 0.0")
    .Attr("am_only_scale: float = 0.0")
    .Output("px: float32")
    .Output("py: float32")
    .SetShapeFn([](InferenceContext *c) {
      c->set_output(0, c->UnknownShapeOfRank(3));
      c->set_output(1, c->UnknownShapeOfRank(3));
      return tf::OkStatus();
    });

class FastRnntSimpleLogprobsOp : public tf::OpKernel {
 public:
  explicit FastRnntSimpleLogprobsOp(tf::OpKernelConstruction *ctx) : tf::OpKernel(ctx) {
    OP_REQUIRES_OK(ctx, ctx->GetAttr("termination_symbol", &term_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("rnnt_type", &type_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("smoothed", &smoothed_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("lm_only_scale", &lms_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("am_only_scale", &ams_));
  }
  void Compute(tf::OpKernelContext *ctx) override {
    const tf::Tensor &lm = ctx->input(0), &am = ctx->input(1), &sym = ctx->input(2), &bd = ctx->input(3);
    OP_REQUIRES(ctx, lm.dims() == 3 && am.dims() == 3 && sym.dims() == 2, tf::errors::InvalidArgument("bad ranks"));
    const int B = am.dim_size(0), T = am.dim_size(1), C = am.dim_size(2), S = lm.dim_size(1) - 1;
    const int T1 = type_ == FRN_REGULAR ? T + 1 : T;
    tf::Tensor *px = nullptr, *py = nullptr, ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({B, S, T1}), &px));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({B, S + 1, T}), &py));
    const size_t bytes = frn_simple_logprobs_workspace_bytes(B, S, T, C);
    OP_REQUIRES_OK(ctx, Workspace(ctx, bytes, &ws));
    OP_REQUIRES_OK(ctx, FromFrn(frn_simple_logprobs(lm.flat<float>().data(), am.flat<float>().data(),
                                                    sym.flat<tf::int32>().data(), bd.flat<tf::int32>().data(), B, S,
                                                    T, C, term_, type_, smoothed_ ? 1 : 0, lms_, ams_,
                                                    px->flat<float>().data(), py->flat<float>().data(),
                                                    ws.flat<tf::uint8>().data(), bytes, StreamOf(ctx)),
                                "FastRnntSimpleLogprobs"));
  }

 private:
  int term_, type_;
  bool smoothed_;
  float lms_, ams_;
};
REGISTER_KERNEL_BUILDER(Name("FastRnntSimpleLogprobs").Device(tf::DEVICE_GPU), FastRnntSimpleLogprobsOp);

REGISTER_OP("FastRnntPrunedLogprobs")
    .Input("logits: T")
    .Input("symbols: int32")
    .Input("ranges: int32")
    .Input("boundary: int32")
    .Attr("T: {float32, bfloat16} = DT_FLOAT")
    .Attr("termination_symbol: int")
    .Attr("rnnt_type: int = 0")
    .Output("px: float32")
    .Output("py: float32")
    .SetShapeFn([](InferenceContext *c) {
      c->set_output(0, c->UnknownShapeOfRank(3));
      c->set_output(1, c->UnknownShapeOfRank(3));
      return tf::OkStatus();
    });

template <typename DT>
class FastRnntPrunedLogprobsOp : public tf::OpKernel {
 public:
  explicit FastRnntPrunedLogprobsOp(tf::OpKernelConstruction *ctx) : tf::OpKernel(ctx) {
    OP_REQUIRES_OK(ctx, ctx->GetAttr("termination_symbol", &term_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("rnnt_type", &type_));
  }
  void Compute(tf::OpKernelContext *ctx) override {
    const tf::Tensor &lg = ctx->input(0), &sym = ctx->input(1), &rg = ctx->input(2), &bd = ctx->input(3);
    OP_REQUIRES_OK(ctx, CheckPrunedShapes(lg, sym, &rg, bd));
    const int B = lg.dim_size(0), T = lg.dim_size(1), R = lg.dim_size(2), C = lg.dim_size(3), S = sym.dim_size(1);
    const int T1 = type_ == FRN_REGULAR ? T + 1 : T;
    tf::Tensor *px = nullptr, *py = nullptr, ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({B, S, T1}), &px));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({B, S + 1, T}), &py));
    const size_t bytes = frn_pruned_logprobs_workspace_bytes(B, S, T, R);
    OP_REQUIRES_OK(ctx, Workspace(ctx, bytes, &ws));
    const int dtype = std::is_same<DT, float>::value ? FRN_F32 : FRN_BF16;
    OP_REQUIRES_OK(ctx, FromFrn(frn_pruned_logprobs(lg.flat<DT>().data(), dtype, sym.flat<tf::int32>().data(),
                                                    rg.flat<tf::int32>().data(), bd.flat<tf::int32>().data(), B, S, T,
                                                    R, C, term_, type_, px->flat<float>().data(),
                                                    py->flat<float>().data(), ws.flat<tf::uint8>().data(), bytes,
                                                    StreamOf(ctx)),
                                "FastRnntPrunedLogprobs"));
  }

 private:
  int term_, type_;
};
REGISTER_KERNEL_BUILDER(Name("FastRnntPrunedLogprobs").Device(tf::DEVICE_GPU).TypeConstraint<float>("T"),
                        FastRnntPrunedLogprobsOp<float>);
REGISTER_KERNEL_BUILDER(Name("FastRnntPrunedLogprobs").Device(tf::DEVICE_GPU).TypeConstraint<tf::bfloat16>("T"),
                        FastRnntPrunedLogprobsOp<tf::bfloat16>);

// ---------------------------------------------------------------------------
// FastRnntPrunedLogprobsGrad: backward of FastRnntPrunedLogprobs (what TF autodiff derives through
// rnnt_loss.py:942-1018) - the reference composes get_rnnt_logprobs_pruned with
// mutual_information_recursion under autodiff (rnnt_loss.py:1088-1119), so the building block needs its
// own gradient.  Registered from Python with RegisterGradient("FastRnntPrunedLogprobs").
// (FastRnntSimpleLogprobs uses FastRnntSimpleLossGrad: the same contractions, fed with the cotangents
// of px / py instead of occupation counts.)
// ---------------------------------------------------------------------------
REGISTER_OP("FastRnntPrunedLogprobsGrad")
    .Input("logits: T")
    .Input("symbols: int32")
    .Input("ranges: int32")
    .Input("boundary: int32")
    .Input("px_grad: float32")
    .Input("py_grad: float32")
    .Attr("T: {float32, bfloat16} = DT_FLOAT")
    .Attr("termination_symbol: int")
    .Attr("rnnt_type: int = 0")
    .Output("logits_grad: T")
    .SetShapeFn([](InferenceContext *c) {
      c->set_output(0, c->input(0));
      return tf::OkStatus();
    });

template <typename DT>
class FastRnntPrunedLogprobsGradOp : public tf::OpKernel {
 public:
  explicit FastRnntPrunedLogprobsGradOp(tf::OpKernelConstruction *ctx) : tf::OpKernel(ctx) {
    OP_REQUIRES_OK(ctx, ctx->GetAttr("termination_symbol", &term_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("rnnt_type", &type_));
  }
  void Compute(tf::OpKernelContext *ctx) override {
    const tf::Tensor &lg = ctx->input(0), &sym = ctx->input(1), &rg = ctx->input(2), &bd = ctx->input(3),
                     &gx = ctx->input(4), &gy = ctx->input(5);
    OP_REQUIRES_OK(ctx, CheckPrunedShapes(lg, sym, &rg, bd));
    const int B = lg.dim_size(0), T = lg.dim_size(1), R = lg.dim_size(2), C = lg.dim_size(3), S = sym.dim_size(1);
    const int T1 = type_ == FRN_REGULAR ? T + 1 : T;
    OP_REQUIRES(ctx, gx.dims() == 3 && gx.dim_size(0) == B && gx.dim_size(1) == S && gx.dim_size(2) == T1,
                tf::errors::InvalidArgument("px_grad must be [B,S,T1]"));
    OP_REQUIRES(ctx, gy.dims() == 3 && gy.dim_size(0) == B && gy.dim_size(1) == S + 1 && gy.dim_size(2) == T,
                tf::errors::InvalidArgument("py_grad must be [B,S+1,T]"));
    tf::Tensor *grad = nullptr, ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, lg.shape(), &grad));
    const size_t bytes = frn_pruned_logprobs_workspace_bytes(B, S, T, R);
    OP_REQUIRES_OK(ctx, Workspace(ctx, bytes, &ws));
    const int dtype = std::is_same<DT, float>::value ? FRN_F32 : FRN_BF16;
    OP_REQUIRES_OK(ctx, FromFrn(frn_pruned_logprobs_bwd(lg.flat<DT>().data(), dtype, sym.flat<tf::int32>().data(),
                                                        rg.flat<tf::int32>().data(), bd.flat<tf::int32>().data(),
                                                        gx.flat<float>().data(), gy.flat<float>().data(), B, S, T, R, C,
                                                        term_, type_, grad->flat<DT>().data(),
                                                        ws.flat<tf::uint8>().data(), bytes, StreamOf(ctx)),
                                "FastRnntPrunedLogprobsGrad"));
  }

 private:
  int term_, type_;
};
REGISTER_KERNEL_BUILDER(Name("FastRnntPrunedLogprobsGrad").Device(tf::DEVICE_GPU).TypeConstraint<float>("T"),
                        FastRnntPrunedLogprobsGradOp<float>);
REGISTER_KERNEL_BUILDER(Name("FastRnntPrunedLogprobsGrad").Device(tf::DEVICE_GPU).TypeConstraint<tf::bfloat16>("T"),
                        FastRnntPrunedLogprobsGradOp<tf::bfloat16>);
