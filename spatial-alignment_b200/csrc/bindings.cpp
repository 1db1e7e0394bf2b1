// torch custom-op layer over the C ABI of include/gpsa_b200.h:  TORCH_LIBRARY(gpsa_b200, ...).
//
// Compiled by g++ on its own (torch headers never enter the CUDA translation units: a .cu that includes
// <torch/extension.h> takes minutes to build) into libgpsa_b200_torch.so, which links against
// libgpsa_b200.so.  Every extern "C" launcher becomes one op `torch.ops.gpsa_b200.<name without the gpsa_ prefix>`:
//   * pointer parameters are Tensors (`Tensor?` for const, `Tensor(a!)?` for outputs / in-place operands; None = NULL),
//     checked here for device, dtype and contiguity -- violations raise RuntimeError (TORCH_CHECK);
//   * sizes are ints, scalars floats; the trailing cudaStream_t is NOT an argument: kernels are enqueued on
//     c10::cuda::getCurrentCUDAStream() of the tensors' device, under a CUDAGuard for that device;
//   * a non-zero return code raises.
// The ops allocate nothing and return nothing (outputs are caller-allocated tensors); forward / backward pairs are
// tied together by the torch.autograd.Function classes in gpsa/_ops.py -- explicit backward, no autograd tape below.
// The struct-based layer entry points (warp / data layer, Adam) take their fields as a flat argument list in the order
// of the C structs.
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "gpsa_b200.h"

namespace {
using at::Tensor;
using OT = c10::optional<Tensor>;

const char* err_text(int rc) {
  return rc == 1 ? "bad argument" : rc == 2 ? "CUDA launch error" : rc == 3 ? "unsupported size or kernel kind" : "unknown error";
}

template <typename T> struct ScalarOf;
template <> struct ScalarOf<float> { static constexpr at::ScalarType v = at::kFloat; static constexpr const char* n = "float32"; };
template <> struct ScalarOf<double> { static constexpr at::ScalarType v = at::kDouble; static constexpr const char* n = "float64"; };
template <> struct ScalarOf<int> { static constexpr at::ScalarType v = at::kInt; static constexpr const char* n = "int32"; };
template <> struct ScalarOf<long long> { static constexpr at::ScalarType v = at::kLong; static constexpr const char* n = "int64"; };
template <> struct ScalarOf<void> { static constexpr at::ScalarType v = at::kByte; static constexpr const char* n = "uint8"; };

template <typename T>
T* tensor_ptr(const OT& t, const char* op, int pos) {
  if (!t.has_value() || !t->defined()) return nullptr;
  TORCH_CHECK(t->is_cuda(), "gpsa_b200::", op, ": argument ", pos, " must be a CUDA tensor (gpsa_b200 has no CPU path)");
  TORCH_CHECK(t->scalar_type() == ScalarOf<T>::v, "gpsa_b200::", op, ": argument ", pos, " must be ", ScalarOf<T>::n,
              ", got ", t->scalar_type());
  TORCH_CHECK(t->is_contiguous(), "gpsa_b200::", op, ": argument ", pos, " must be contiguous");
  return static_cast<T*>(t->data_ptr());
}

// ---- C parameter type -> op parameter type, schema fragment, conversion --------------------------------------------
template <typename P> struct Arg;
template <typename T> struct Arg<const T*> {
  using op = const OT&;
  static std::string schema(int i) { return "Tensor? a" + std::to_string(i); }
  static const T* get(const OT& t, const char* name, int pos) { return tensor_ptr<T>(t, name, pos); }
  static void device(const OT& t, c10::optional<c10::Device>& d) { if (!d && t.has_value() && t->defined()) d = t->device(); }
};
template <typename T> struct Arg<T*> {
  using op = const OT&;
  static std::string schema(int i) { return "Tensor(m" + std::to_string(i) + "!)? a" + std::to_string(i); }
  static T* get(const OT& t, const char* name, int pos) { return tensor_ptr<T>(t, name, pos); }
  static void device(const OT& t, c10::optional<c10::Device>& d) { if (!d && t.has_value() && t->defined()) d = t->device(); }
};
template <typename I> struct IntArg {
  using op = int64_t;
  static std::string schema(int i) { return "int a" + std::to_string(i); }
  static I get(int64_t v, const char*, int) { return static_cast<I>(v); }
  static void device(int64_t, c10::optional<c10::Device>&) {}
};
template <> struct Arg<int> : IntArg<int> {};
template <> struct Arg<long> : IntArg<long> {};
template <> struct Arg<unsigned long> : IntArg<unsigned long> {};
template <> struct Arg<float> {
  using op = double;
  static std::string schema(int i) { return "float a" + std::to_string(i); }
  static float get(double v, const char*, int) { return static_cast<float>(v); }
  static void device(double, c10::optional<c10::Device>&) {}
};

template <typename F> struct FnTraits;
template <typename... Ps> struct FnTraits<int (*)(Ps...)> {
  using args = std::tuple<Ps...>;
  static constexpr size_t n = sizeof...(Ps);
};

// One op per extern "C" launcher whose LAST parameter is the stream.
template <auto Fn, const char* Name, size_t... I>
struct Wrapped {
  using Tr = FnTraits<decltype(Fn)>;
  template <size_t K> using P = std::tuple_element_t<K, typename Tr::args>;
  static void call(typename Arg<P<I>>::op... a) {
    auto cargs = std::make_tuple(Arg<P<I>>::get(a, Name, (int)I)...);  // device / dtype / contiguity checks happen here
    c10::optional<c10::Device> dev;
    (Arg<P<I>>::device(a, dev), ...);
    TORCH_CHECK(dev.has_value(), "gpsa_b200::", Name, ": no tensor argument");
    c10::cuda::CUDAGuard guard(*dev);
    cudaStream_t st = c10::cuda::getCurrentCUDAStream(dev->index()).stream();
    const int rc = std::apply([&](auto... c) { return Fn(c..., st); }, cargs);
    TORCH_CHECK(rc == 0, "gpsa_b200::", Name, " failed: ", err_text(rc));
  }
  static std::string schema() {
    std::string s = std::string(Name) + "(";
    const std::vector<std::string> parts = {Arg<P<I>>::schema((int)I)...};
    for (size_t k = 0; k < parts.size(); ++k) s += (k ? ", " : "") + parts[k];
    return s + ") -> ()";
  }
};
template <auto Fn, const char* Name, size_t... I>
Wrapped<Fn, Name, I...> make_wrapped(std::index_sequence<I...>) { return {}; }

#define GPSA_OP(m, fn)                                                                                   \
  do {                                                                                                   \
    static constexpr char name_[] = #fn;                                                                 \
    using W = decltype(make_wrapped<&gpsa_##fn, name_>(                                                  \
        std::make_index_sequence<FnTraits<decltype(&gpsa_##fn)>::n - 1>{}));                             \
    m.def(W::schema().c_str(), &W::call);                                                                \
  } while (0)

// ---- struct-based entry points: flat argument lists ------------------------------------------------------------------
cudaStream_t stream_of(const Tensor& t) { return c10::cuda::getCurrentCUDAStream(t.get_device()).stream(); }
#define FP(x, i) tensor_ptr<float>(x, OPNAME, i)
#define DP(x, i) tensor_ptr<double>(x, OPNAME, i)
#define IP(x, i) tensor_ptr<int>(x, OPNAME, i)

void warp_view_fwd(int64_t kind, int64_t D, int64_t M, int64_t V, int64_t v, int64_t S, int64_t n, const Tensor& Z,
                   const OT& dlt, const OT& log_ls, const OT& log_var, const OT& Omega_G, const OT& hld_Omega, const OT& X,
                   const OT& eps, const OT& Lk, const OT& Kinv, const OT& Kinv64, const OT& hld_K, const OT& info,
                   const OT& A, const OT& B, const OT& T, const OT& Ke, const OT& var, const OT& Gmean, const OT& Gs,
                   int64_t gs_stride, const OT& kl_acc, const OT& ws64, const OT& Kuu_ext, const OT& Kuf_ext) {
  static constexpr const char* OPNAME = "warp_view_fwd";
  c10::cuda::CUDAGuard guard(Z.device());
  gpsa_warp_fwd_args a = {};
  a.kind = (int)kind; a.D = (int)D; a.M = (int)M; a.V = (int)V; a.v = (int)v; a.S = (int)S; a.n = (long)n;
  a.Z = FP(Z, 7); a.dlt = FP(dlt, 8); a.log_ls = FP(log_ls, 9); a.log_var = FP(log_var, 10); a.Omega_G = FP(Omega_G, 11);
  a.hld_Omega = DP(hld_Omega, 12); a.X = FP(X, 13); a.eps = FP(eps, 14); a.Lk = FP(Lk, 15); a.Kinv = FP(Kinv, 16);
  a.Kinv64 = DP(Kinv64, 17); a.hld_K = DP(hld_K, 18); a.info = IP(info, 19); a.A = DP(A, 20); a.B = DP(B, 21);
  a.T = DP(T, 22); a.Ke = DP(Ke, 23); a.var = FP(var, 24); a.Gmean = FP(Gmean, 25); a.Gs = FP(Gs, 26);
  a.gs_stride = (long)gs_stride; a.kl_acc = DP(kl_acc, 28); a.ws64 = DP(ws64, 29); a.Kuu_ext = FP(Kuu_ext, 30);
  a.Kuf_ext = FP(Kuf_ext, 31);
  const int rc = gpsa_warp_view_fwd(&a, stream_of(Z));
  TORCH_CHECK(rc == 0, "gpsa_b200::warp_view_fwd failed: ", err_text(rc));
}

void warp_view_bwd(int64_t kind, int64_t D, int64_t M, int64_t V, int64_t v, int64_t S, int64_t n, const Tensor& Z,
                   const OT& dlt, const OT& log_ls, const OT& log_var, const OT& Omega_G, const OT& X, const OT& eps,
                   const OT& Kinv64, const OT& A, const OT& B, const OT& T, const OT& Ke, const OT& Gs_bar,
                   int64_t gs_stride, const OT& Gm_bar, const OT& kl_bar, const OT& acc_Z, const OT& acc_dlt,
                   const OT& acc_hyp, const OT& Obar_G, const OT& mubar, const OT& varbar, const OT& q1bar, const OT& Abar,
                   const OT& C, const OT& AS, const OT& ws64, const OT& Kuu_bar, const OT& Kuf_bar) {
  static constexpr const char* OPNAME = "warp_view_bwd";
  c10::cuda::CUDAGuard guard(Z.device());
  gpsa_warp_bwd_args a = {};
  a.kind = (int)kind; a.D = (int)D; a.M = (int)M; a.V = (int)V; a.v = (int)v; a.S = (int)S; a.n = (long)n;
  a.Z = FP(Z, 7); a.dlt = FP(dlt, 8); a.log_ls = FP(log_ls, 9); a.log_var = FP(log_var, 10); a.Omega_G = FP(Omega_G, 11);
  a.X = FP(X, 12); a.eps = FP(eps, 13); a.Kinv64 = DP(Kinv64, 14); a.A = DP(A, 15); a.B = DP(B, 16); a.T = DP(T, 17);
  a.Ke = DP(Ke, 18); a.Gs_bar = FP(Gs_bar, 19); a.gs_stride = (long)gs_stride; a.Gm_bar = FP(Gm_bar, 21);
  a.kl_bar = FP(kl_bar, 22); a.acc_Z = DP(acc_Z, 23); a.acc_dlt = DP(acc_dlt, 24); a.acc_hyp = DP(acc_hyp, 25);
  a.Obar_G = FP(Obar_G, 26); a.mubar = FP(mubar, 27); a.varbar = FP(varbar, 28); a.q1bar = FP(q1bar, 29);
  a.Abar = DP(Abar, 30); a.C = DP(C, 31); a.AS = DP(AS, 32); a.ws64 = DP(ws64, 33); a.Kuu_bar = FP(Kuu_bar, 34);
  a.Kuf_bar = FP(Kuf_bar, 35);
  const int rc = gpsa_warp_view_bwd(&a, stream_of(Z));
  TORCH_CHECK(rc == 0, "gpsa_b200::warp_view_bwd failed: ", err_text(rc));
}

void data_layer_fwd(int64_t kind, int64_t D, int64_t M, int64_t L, int64_t R, const Tensor& Gt, const OT& log_ls,
                    const OT& log_var, const OT& dlt, const OT& Omega, const OT& hld_Omega, const OT& G, const OT& Lk,
                    const OT& Kinv, const OT& Kinv64, const OT& hld_K, const OT& info, const OT& A, const OT& B,
                    const OT& kq, const OT& W, const OT& KD, const OT& mean, const OT& q2, const OT& kl_acc, const OT& ws64,
                    int64_t engine, const OT& tc_ws, const OT& Kuu_ext, int64_t prior_ready) {
  static constexpr const char* OPNAME = "data_layer_fwd";
  c10::cuda::CUDAGuard guard(Gt.device());
  gpsa_data_fwd_args a = {};
  a.kind = (int)kind; a.D = (int)D; a.M = (int)M; a.L = (int)L; a.R = (long)R;
  a.Gt = FP(Gt, 5); a.log_ls = FP(log_ls, 6); a.log_var = FP(log_var, 7); a.dlt = FP(dlt, 8); a.Omega = FP(Omega, 9);
  a.hld_Omega = DP(hld_Omega, 10); a.G = FP(G, 11); a.Lk = FP(Lk, 12); a.Kinv = FP(Kinv, 13); a.Kinv64 = DP(Kinv64, 14);
  a.hld_K = DP(hld_K, 15); a.info = IP(info, 16); a.A = FP(A, 17); a.B = FP(B, 18); a.kq = FP(kq, 19); a.W = FP(W, 20);
  a.KD = DP(KD, 21); a.mean = FP(mean, 22); a.q2 = FP(q2, 23); a.kl_acc = DP(kl_acc, 24); a.ws64 = DP(ws64, 25);
  a.engine = (int)engine; a.tc_ws = tensor_ptr<void>(tc_ws, OPNAME, 27);
  a.tc_ws_bytes = (tc_ws.has_value() && tc_ws->defined()) ? (size_t)tc_ws->numel() : 0; a.Kuu_ext = FP(Kuu_ext, 28);
  a.prior_ready = (int)prior_ready;
  const int rc = gpsa_data_layer_fwd(&a, stream_of(Gt));
  TORCH_CHECK(rc == 0, "gpsa_b200::data_layer_fwd failed: ", err_text(rc));
}

void data_layer_bwd(int64_t kind, int64_t D, int64_t M, int64_t L, int64_t R, const Tensor& Gt, const OT& log_ls,
                    const OT& log_var, const OT& dlt, const OT& Omega, const OT& G, const OT& Kinv, const OT& Kinv64,
                    const OT& A, const OT& B, const OT& W, const OT& KD, const OT& mean_bar, const OT& q2_bar,
                    const OT& kq_bar, const OT& kl_bar, const OT& G_bar, const OT& acc_Gt, const OT& acc_hyp,
                    const OT& dlt_bar, const OT& Obar, const OT& q1bar, const OT& Abar, const OT& C, const OT& H,
                    const OT& ws64, int64_t engine, const OT& tc_ws, const OT& Kuu_bar) {
  static constexpr const char* OPNAME = "data_layer_bwd";
  c10::cuda::CUDAGuard guard(Gt.device());
  gpsa_data_bwd_args a = {};
  a.kind = (int)kind; a.D = (int)D; a.M = (int)M; a.L = (int)L; a.R = (long)R;
  a.Gt = FP(Gt, 5); a.log_ls = FP(log_ls, 6); a.log_var = FP(log_var, 7); a.dlt = FP(dlt, 8); a.Omega = FP(Omega, 9);
  a.G = FP(G, 10); a.Kinv = FP(Kinv, 11); a.Kinv64 = DP(Kinv64, 12); a.A = FP(A, 13); a.B = FP(B, 14); a.W = FP(W, 15);
  a.KD = DP(KD, 16); a.mean_bar = FP(mean_bar, 17); a.q2_bar = FP(q2_bar, 18); a.kq_bar = FP(kq_bar, 19);
  a.kl_bar = FP(kl_bar, 20); a.G_bar = FP(G_bar, 21); a.acc_Gt = DP(acc_Gt, 22); a.acc_hyp = DP(acc_hyp, 23);
  a.dlt_bar = FP(dlt_bar, 24); a.Obar = FP(Obar, 25); a.q1bar = FP(q1bar, 26); a.Abar = FP(Abar, 27); a.C = FP(C, 28);
  a.H = FP(H, 29); a.ws64 = DP(ws64, 30); a.engine = (int)engine; a.tc_ws = tensor_ptr<void>(tc_ws, OPNAME, 32);
  a.tc_ws_bytes = (tc_ws.has_value() && tc_ws->defined()) ? (size_t)tc_ws->numel() : 0; a.Kuu_bar = FP(Kuu_bar, 33);
  const int rc = gpsa_data_layer_bwd(&a, stream_of(Gt));
  TORCH_CHECK(rc == 0, "gpsa_b200::data_layer_bwd failed: ", err_text(rc));
}

// Adam over up to GPSA_ADAM_MAX_TENSORS tensors; grads[k] may be an undefined / None entry
void adam_step(at::TensorList params, const c10::List<OT>& grads, at::TensorList exp_avg, at::TensorList exp_avg_sq,
               double lr, double beta1, double beta2, double eps, const Tensor& step) {
  static constexpr const char* OPNAME = "adam_step";
  const size_t n = params.size();
  TORCH_CHECK(n <= GPSA_ADAM_MAX_TENSORS && grads.size() == n && exp_avg.size() == n && exp_avg_sq.size() == n,
              "gpsa_b200::adam_step: at most ", GPSA_ADAM_MAX_TENSORS, " tensors per call, lists of equal length");
  if (n == 0) return;
  c10::cuda::CUDAGuard guard(params[0].device());
  gpsa_adam_args a = {};
  a.count = (int)n;
  for (size_t k = 0; k < n; ++k) {
    a.p[k] = FP(OT(params[k]), (int)k);
    a.m[k] = FP(OT(exp_avg[k]), (int)k);
    a.v[k] = FP(OT(exp_avg_sq[k]), (int)k);
    const OT g = grads.get(k);
    a.g[k] = FP(g, (int)k);
    TORCH_CHECK(!a.g[k] || g->numel() == params[k].numel(), "gpsa_b200::adam_step: gradient ", k, " has the wrong size");
    a.n[k] = (long)params[k].numel();
  }
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
  a.step = FP(OT(step), -1);
  TORCH_CHECK(step.numel() >= (int64_t)n, "gpsa_b200::adam_step: step needs one float per tensor");
  const int rc = gpsa_adam_step(&a, stream_of(params[0]));
  TORCH_CHECK(rc == 0, "gpsa_b200::adam_step failed: ", err_text(rc));
}

}  // namespace

TORCH_LIBRARY(gpsa_b200, m) {
  // covariance functions, factorisations, variational covariances
  GPSA_OP(m, kernel_matrix_fwd);
  GPSA_OP(m, kernel_matrix_bwd);
  GPSA_OP(m, potrf_batched_f32);
  GPSA_OP(m, potrf_batched_f64);
  GPSA_OP(m, potrf_batched_f32_ld64);
  GPSA_OP(m, trtri_batched_f32);
  GPSA_OP(m, trtri_batched_f64);
  GPSA_OP(m, prior_prepare);
  GPSA_OP(m, prior_prepare_ext);
  GPSA_OP(m, omega_prepare);
  GPSA_OP(m, omega_grad);
  GPSA_OP(m, omega_grad_tc);
  GPSA_OP(m, omega_grad_f32);
  GPSA_OP(m, gemm_f32);
  GPSA_OP(m, gemm_tc);
  GPSA_OP(m, tc_gemm_test);
  // quadratic-form engines
  GPSA_OP(m, feat_pack);
  GPSA_OP(m, feat_unpack);
  GPSA_OP(m, quadform_fwd_f32);
  GPSA_OP(m, quadform_bwd_omega_f32);
  GPSA_OP(m, quadform_bwd_alpha_f32);
  GPSA_OP(m, quadform_fwd_feat_tc);
  GPSA_OP(m, quadform_bwd_alpha_tc);
  GPSA_OP(m, quadform_bwd_omega_tc);
  // sampling stage, likelihood, LMC
  GPSA_OP(m, sample_fwd);
  GPSA_OP(m, sample_bwd);
  GPSA_OP(m, philox_normal);
  GPSA_OP(m, sample_ll_fused);
  GPSA_OP(m, scale_if_not_one);
  GPSA_OP(m, gaussian_ll_fwd);
  GPSA_OP(m, gaussian_ll_bwd);
  GPSA_OP(m, lmc_fwd);
  GPSA_OP(m, lmc_bwd);
  GPSA_OP(m, lmc_ll_fused);
  GPSA_OP(m, kmeans_lloyd);
  // layers and optimiser (flat argument lists in the order of the C structs)
  m.def("warp_view_fwd(int kind, int D, int M, int V, int v, int S, int n, Tensor Z, Tensor? dlt, Tensor? log_ls, "
        "Tensor? log_var, Tensor? Omega_G, Tensor? hld_Omega, Tensor? X, Tensor? eps, Tensor(o0!)? Lk, Tensor(o1!)? Kinv, "
        "Tensor(o2!)? Kinv64, Tensor(o3!)? hld_K, Tensor(o4!)? info, Tensor(o5!)? A, Tensor(o6!)? B, Tensor(o7!)? T, "
        "Tensor(o8!)? Ke, Tensor(o9!)? var, Tensor(o10!)? Gmean, Tensor(o11!)? Gs, int gs_stride, Tensor(o12!)? kl_acc, "
        "Tensor(o13!)? ws64, Tensor? Kuu_ext, Tensor? Kuf_ext) -> ()", &warp_view_fwd);
  m.def("warp_view_bwd(int kind, int D, int M, int V, int v, int S, int n, Tensor Z, Tensor? dlt, Tensor? log_ls, "
        "Tensor? log_var, Tensor? Omega_G, Tensor? X, Tensor? eps, Tensor? Kinv64, Tensor? A, Tensor? B, Tensor? T, "
        "Tensor? Ke, Tensor? Gs_bar, int gs_stride, Tensor? Gm_bar, Tensor? kl_bar, Tensor(o0!)? acc_Z, "
        "Tensor(o1!)? acc_dlt, Tensor(o2!)? acc_hyp, Tensor(o3!)? Obar_G, Tensor(o4!)? mubar, Tensor(o5!)? varbar, "
        "Tensor(o6!)? q1bar, Tensor(o7!)? Abar, Tensor(o8!)? C, Tensor(o9!)? AS, Tensor(o10!)? ws64, Tensor(o11!)? Kuu_bar, "
        "Tensor(o12!)? Kuf_bar) -> ()", &warp_view_bwd);
  m.def("data_layer_fwd(int kind, int D, int M, int L, int R, Tensor Gt, Tensor? log_ls, Tensor? log_var, Tensor? dlt, "
        "Tensor? Omega, Tensor? hld_Omega, Tensor? G, Tensor(o0!)? Lk, Tensor(o1!)? Kinv, Tensor(o2!)? Kinv64, "
        "Tensor(o3!)? hld_K, Tensor(o4!)? info, Tensor(o5!)? A, Tensor(o6!)? B, Tensor(o7!)? kq, Tensor(o8!)? W, "
        "Tensor(o9!)? KD, Tensor(o10!)? mean, Tensor(o11!)? q2, Tensor(o12!)? kl_acc, Tensor(o13!)? ws64, int engine, "
        "Tensor(o14!)? tc_ws, Tensor? Kuu_ext, int prior_ready) -> ()", &data_layer_fwd);
  m.def("data_layer_bwd(int kind, int D, int M, int L, int R, Tensor Gt, Tensor? log_ls, Tensor? log_var, Tensor? dlt, "
        "Tensor? Omega, Tensor? G, Tensor? Kinv, Tensor? Kinv64, Tensor? A, Tensor? B, Tensor? W, Tensor? KD, "
        "Tensor? mean_bar, Tensor? q2_bar, Tensor? kq_bar, Tensor? kl_bar, Tensor(o0!)? G_bar, Tensor(o1!)? acc_Gt, "
        "Tensor(o2!)? acc_hyp, Tensor(o3!)? dlt_bar, Tensor(o4!)? Obar, Tensor(o5!)? q1bar, Tensor(o6!)? Abar, "
        "Tensor(o7!)? C, Tensor(o8!)? H, Tensor(o9!)? ws64, int engine, Tensor(o10!)? tc_ws, Tensor(o11!)? Kuu_bar) -> ()",
        &data_layer_bwd);
  m.def("adam_step(Tensor(a!)[] params, Tensor?[] grads, Tensor(b!)[] exp_avg, Tensor(c!)[] exp_avg_sq, float lr, "
        "float beta1, float beta2, float eps, Tensor(d!) step) -> ()", &adam_step);
}
