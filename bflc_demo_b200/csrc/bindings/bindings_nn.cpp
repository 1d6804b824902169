// Bindings for the non-GEMM NN kernels (csrc/kernels/nn_kernels.cu).
#include <ATen/cuda/CUDAContext.h>
#include <torch/extension.h>

#include <optional>

#include "bflc_kernels.h"

namespace py = pybind11;
using OptT = std::optional<at::Tensor>;

namespace {
void check(cudaError_t e, const char* what) {
  TORCH_CHECK(e == cudaSuccess, "bflc::", what, " failed: ", cudaGetErrorString(e));
}
cudaStream_t st() { return at::cuda::getCurrentCUDAStream().stream(); }
template <typename T>
T* optp(const OptT& t) { return t.has_value() ? reinterpret_cast<T*>(t->data_ptr()) : nullptr; }
}  // namespace

void bind_nn(py::module_& m) {
  m.def("im2col", [](at::Tensor x, at::Tensor col, int N, int C, int H, int W, int KH, int KW,
                     int stride, int pad, int OH, int OW) {
    check(bflc::im2col_bf16(x.data_ptr(), col.data_ptr(), N, C, H, W, KH, KW, stride, pad, OH, OW,
                            col.stride(0), st()), "im2col");
  });
  m.def("col2im", [](at::Tensor col, at::Tensor dx, int N, int C, int H, int W, int KH, int KW,
                     int stride, int pad, int OH, int OW) {
    check(bflc::col2im_bf16(col.data_ptr(), dx.data_ptr(), N, C, H, W, KH, KW, stride, pad, OH, OW,
                            col.stride(0), st()), "col2im");
  });
  m.def("upsample_zero", [](at::Tensor dy, at::Tensor up, int N, int H, int W, int OH, int OW, int Cc,
                            int stride) {
    check(bflc::upsample_zero_bf16(dy.data_ptr(), up.data_ptr(), N, H, W, OH, OW, Cc, stride, st()),
          "upsample_zero");
  });
  m.def("maxpool_fwd", [](at::Tensor x, at::Tensor y, at::Tensor idx, int N, int C, int H, int W,
                          int k, int stride, int pad, int OH, int OW) {
    check(bflc::maxpool2d_fwd(x.data_ptr(), y.data_ptr(), idx.data_ptr<int32_t>(), N, C, H, W, k,
                              stride, pad, OH, OW, st()), "maxpool_fwd");
  });
  m.def("maxpool_bwd", [](at::Tensor dy, at::Tensor idx, at::Tensor dx_f32, int64_t per_out,
                          int64_t per_in) {
    check(bflc::maxpool2d_bwd(dy.data_ptr(), idx.data_ptr<int32_t>(), dx_f32.data_ptr<float>(),
                              dy.numel(), per_out, per_in, st()), "maxpool_bwd");
  });
  m.def("avgpool_fwd", [](at::Tensor x, at::Tensor y, int N, int HW, int C) {
    check(bflc::avgpool_global_fwd(x.data_ptr(), y.data_ptr(), N, HW, C, st()), "avgpool_fwd");
  });
  m.def("avgpool_bwd", [](at::Tensor dy, at::Tensor dx, int N, int HW, int C) {
    check(bflc::avgpool_global_bwd(dy.data_ptr(), dx.data_ptr(), N, HW, C, st()), "avgpool_bwd");
  });
  m.def("batchnorm_fwd", [](at::Tensor x, at::Tensor y, at::Tensor gamma, at::Tensor beta,
                            at::Tensor mean, at::Tensor rstd, const OptT& run_mean,
                            const OptT& run_var, int64_t rows, int C, double eps, double momentum,
                            bool training, bool relu, const OptT& residual) {
    check(bflc::batchnorm_fwd(x.data_ptr(), y.data_ptr(), gamma.data_ptr<float>(),
                              beta.data_ptr<float>(), mean.data_ptr<float>(), rstd.data_ptr<float>(),
                              optp<float>(run_mean), optp<float>(run_var), rows, C, (float)eps,
                              (float)momentum, training, relu,
                              residual.has_value() ? residual->data_ptr() : nullptr, st()),
          "batchnorm_fwd");
  });
  m.def("batchnorm_bwd", [](at::Tensor dy, at::Tensor x, at::Tensor y, at::Tensor gamma,
                            at::Tensor mean, at::Tensor rstd, at::Tensor dx, at::Tensor dgamma,
                            at::Tensor dbeta, const OptT& dres, int64_t rows, int C, bool relu) {
    check(bflc::batchnorm_bwd(dy.data_ptr(), x.data_ptr(), y.data_ptr(), gamma.data_ptr<float>(),
                              mean.data_ptr<float>(), rstd.data_ptr<float>(), dx.data_ptr(),
                              dgamma.data_ptr<float>(), dbeta.data_ptr<float>(),
                              dres.has_value() ? dres->data_ptr() : nullptr, rows, C, relu, st()),
          "batchnorm_bwd");
  });
  m.def("layernorm_fwd", [](at::Tensor x, at::Tensor y, at::Tensor gamma, at::Tensor beta,
                            at::Tensor mean, at::Tensor rstd, int64_t rows, int C, double eps) {
    check(bflc::layernorm_fwd(x.data_ptr(), nullptr, y.data_ptr(), gamma.data_ptr<float>(),
                              beta.data_ptr<float>(), mean.data_ptr<float>(), rstd.data_ptr<float>(),
                              rows, C, (float)eps, st()), "layernorm_fwd");
  });
  m.def("layernorm_bwd", [](at::Tensor dy, at::Tensor x, at::Tensor gamma, at::Tensor mean,
                            at::Tensor rstd, at::Tensor dx, at::Tensor dgamma, at::Tensor dbeta,
                            int64_t rows, int C) {
    check(bflc::layernorm_bwd(dy.data_ptr(), x.data_ptr(), gamma.data_ptr<float>(),
                              mean.data_ptr<float>(), rstd.data_ptr<float>(), dx.data_ptr(),
                              dgamma.data_ptr<float>(), dbeta.data_ptr<float>(), rows, C, st()),
          "layernorm_bwd");
  });
  m.def("softmax_fwd", [](at::Tensor x, at::Tensor y, int64_t rows, int cols, double scale) {
    check(bflc::softmax_rows_fwd(x.data_ptr(), y.data_ptr(), rows, cols, (float)scale, st()),
          "softmax_fwd");
  });
  m.def("softmax_bwd", [](at::Tensor dy, at::Tensor y, at::Tensor dx, int64_t rows, int cols,
                          double scale) {
    check(bflc::softmax_rows_bwd(dy.data_ptr(), y.data_ptr(), dx.data_ptr(), rows, cols,
                                 (float)scale, st()), "softmax_bwd");
  });
  m.def("embedding_fwd", [](at::Tensor ids, at::Tensor table, const OptT& pos, at::Tensor out,
                            int64_t rows, int seq, int C) {
    check(bflc::embedding_fwd(ids.data_ptr<int32_t>(), table.data_ptr(),
                              pos.has_value() ? pos->data_ptr() : nullptr, out.data_ptr(), rows, seq,
                              C, st()), "embedding_fwd");
  });
  m.def("embedding_bwd", [](at::Tensor ids, at::Tensor dy, at::Tensor dtable, const OptT& dpos,
                            int64_t rows, int seq, int C) {
    check(bflc::embedding_bwd(ids.data_ptr<int32_t>(), dy.data_ptr(), dtable.data_ptr<float>(),
                              optp<float>(dpos), rows, seq, C, st()), "embedding_bwd");
  });
  m.def("act_bwd_colsum", [](at::Tensor dy, const OptT& aux, const OptT& dz, const OptT& colsum,
                             int64_t rows, int C, int mode) {
    check(bflc::act_bwd_colsum(dy.data_ptr(), aux.has_value() ? aux->data_ptr() : nullptr,
                               dz.has_value() ? dz->data_ptr() : nullptr, optp<float>(colsum), rows,
                               C, mode, st()), "act_bwd_colsum");
  });
  // ---- block-scaled fp8 (MXFP8) ----
  m.def("mx8_sf_bytes", [](int64_t rows, int64_t K) { return bflc::mx8_sf_bytes((int)rows, (int)K); });
  m.def("quantize_mx8", [](at::Tensor x, at::Tensor q, at::Tensor sf, int64_t R, int64_t K,
                           double in_scale) {
    bflc::DType dt = x.scalar_type() == at::kFloat      ? bflc::DType::F32
                     : x.scalar_type() == at::kBFloat16 ? bflc::DType::BF16
                                                        : bflc::DType::U8;
    TORCH_CHECK(x.scalar_type() == at::kFloat || x.scalar_type() == at::kBFloat16 ||
                    x.scalar_type() == at::kByte, "quantize_mx8: f32 / bf16 / u8 input");
    TORCH_CHECK(x.stride(-1) == 1 && q.stride(-1) == 1, "quantize_mx8: contiguous rows");
    TORCH_CHECK(sf.numel() >= bflc::mx8_sf_bytes((int)R, (int)K), "quantize_mx8: sf buffer too small");
    check(bflc::quantize_mx8(x.data_ptr(), dt, x.stride(0), (int)R, (int)K, (float)in_scale,
                             q.data_ptr(), q.stride(0), sf.data_ptr(), st()), "quantize_mx8");
  });
  m.def("gemm_mx8", [](at::Tensor a, at::Tensor sfa, at::Tensor b, at::Tensor sfb, at::Tensor d,
                       int64_t M, int64_t N, int64_t K, double alpha, const OptT& bias, int64_t act) {
    bflc::Mx8Problem p;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.a = a.data_ptr(); p.lda = a.stride(0); p.sfa = static_cast<const uint8_t*>(sfa.data_ptr());
    p.b = b.data_ptr(); p.ldb = b.stride(0); p.sfb = static_cast<const uint8_t*>(sfb.data_ptr());
    TORCH_CHECK(sfa.numel() >= bflc::mx8_sf_bytes(p.M, p.K) && sfb.numel() >= bflc::mx8_sf_bytes(p.N, p.K),
                "gemm_mx8: scale-factor buffers too small");
    p.d = d.data_ptr();
    p.d_dtype = d.scalar_type() == at::kFloat ? bflc::DType::F32 : bflc::DType::BF16;
    p.ldd = d.stride(0);
    p.alpha = (float)alpha;
    p.bias = optp<const float>(bias);
    p.act = static_cast<bflc::Act>(act);
    check(bflc::gemm_mx8_sm100(p, st()), "gemm_mx8_sm100");
  });
  // fused attention (seq 128, head dim 64): q, k, v, o, gradients are [B*S, H*64] bf16
  // Implicit-GEMM convolution (ConvView in bflc_kernels.h).  `x` is the NHWC activation the
  // shifted boxes are read from, `other` the dense operand (weights, or dy for mode 2).
  m.def("conv_gemm", [](int mode, int flip, at::Tensor x, at::Tensor other, at::Tensor d, int N, int H,
                        int W, int Cc, int OH, int OW, int KH, int KW, int stride, int pad, int n_out,
                        const OptT& bias, int act, const OptT& aux_out, const OptT& aux_in, int act_bwd,
                        const OptT& colsum, int split_k, bool accumulate) {
    bflc::GemmProblem p;
    auto& cv = p.conv;
    cv.mode = mode; cv.flip = flip; cv.x = x.data_ptr();
    cv.N = N; cv.H = H; cv.W = W; cv.C = Cc; cv.OH = OH; cv.OW = OW;
    cv.KH = KH; cv.KW = KW; cv.stride = stride; cv.pad = pad;
    const int taps = KH * KW;
    const int64_t pixels = (int64_t)N * OH * OW;
    if (mode == 1) {
      p.M = (int)pixels; p.N = n_out; p.K = taps * Cc;
      p.a = {x.data_ptr(), Cc, 0, false};
      p.b = {other.data_ptr(), other.stride(0), 0, flip != 0};
    } else {
      p.M = n_out; p.N = taps * Cc; p.K = (int)pixels;
      p.a = {other.data_ptr(), other.stride(0), 0, true};   // dy [pixels][Cout]
      p.b = {x.data_ptr(), Cc, 0, true};
    }
    auto& e = p.epi;
    e.d = d.data_ptr();
    e.d_dtype = d.scalar_type() == at::kFloat ? bflc::DType::F32 : bflc::DType::BF16;
    e.ldd = d.stride(0);
    e.bias = optp<const float>(bias);
    e.act = static_cast<bflc::Act>(act);
    e.aux_out = aux_out.has_value() ? aux_out->data_ptr() : nullptr;
    e.aux_in = aux_in.has_value() ? aux_in->data_ptr() : nullptr;
    e.act_bwd = act_bwd;
    e.colsum = optp<float>(colsum);
    e.split_k = split_k;
    e.accumulate = accumulate ? 1 : 0;
    check(bflc::gemm_sm100(p, st()), "conv_gemm (implicit-GEMM convolution)");
  });
  m.def("attention_fwd", [](at::Tensor q, at::Tensor k, at::Tensor v, at::Tensor o, at::Tensor lse, int B,
                            int S, int H, double scale) {
    const int D = (int)q.size(1) / H;
    TORCH_CHECK(q.stride(1) == 1 && k.stride(1) == 1 && v.stride(1) == 1 && o.stride(1) == 1 &&
                q.stride(0) == k.stride(0) && q.stride(0) == v.stride(0) && q.stride(0) == o.stride(0),
                "attention: q, k, v, o must share one row pitch");
    check(bflc::attention_fwd_sm100(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr<float>(),
                                    B, S, H, D, q.stride(0), (float)scale, st()),
          "attention_fwd_sm100");
  });
  m.def("attention_bwd", [](at::Tensor q, at::Tensor k, at::Tensor v, at::Tensor o, at::Tensor dout, at::Tensor lse,
                            at::Tensor dq, at::Tensor dk, at::Tensor dv, int B, int S, int H, double scale) {
    const int D = (int)q.size(1) / H;
    const int64_t ld = q.stride(0);
    for (const at::Tensor* t : {&k, &v, &o, &dout, &dq, &dk, &dv})
      TORCH_CHECK(t->stride(1) == 1 && t->stride(0) == ld, "attention: all operands must share one row pitch");
    check(bflc::attention_bwd_sm100(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), dout.data_ptr(),
                                    lse.data_ptr<float>(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), B, S, H, D,
                                    ld, (float)scale, st()),
          "attention_bwd_sm100");
  });
  m.def("transpose_0213", [](at::Tensor x, at::Tensor y, int d0, int d1, int d2, int d3) {
    check(bflc::transpose_0213_bf16(x.data_ptr(), y.data_ptr(), d0, d1, d2, d3, st()),
          "transpose_0213");
  });
}
