"""ctypes binding of the C ABI declared in include/gpsa_b200.h.

The shared library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a).  There is no
CPU or PyTorch fallback: if the library is missing or a tensor is not a contiguous CUDA float32
tensor the call raises.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgpsa_b200.so")

KIND_RBF, KIND_MATERN12, KIND_MATERN32, KIND_EXTERNAL = 0, 1, 2, 3
OFF = 1e-5

_lib = None

_ERR = {1: "bad argument", 2: "CUDA launch error", 3: "unsupported size or kernel kind"}


class GPSALibraryError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GPSALibraryError(
                f"{LIB_PATH} not found: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()'). "
                "gpsa_b200 has no CPU fallback."
            )
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        _declare(_lib)
    return _lib


TORCH_LIB_PATH = os.path.join(_HERE, "libgpsa_b200_torch.so")
_ops_ns = None


def ops():
    """torch.ops.gpsa_b200: the C ABI registered as torch custom ops (csrc/bindings.cpp, TORCH_LIBRARY(gpsa_b200)).
    This is what gpsa/_ops.py calls; the ctypes table below stays for direct C-ABI use (tests, tools)."""
    global _ops_ns
    if _ops_ns is None:
        lib()  # the kernels' library first: the op library links against it
        if not os.path.exists(TORCH_LIB_PATH):
            raise GPSALibraryError(
                f"{TORCH_LIB_PATH} not found: build it first (python -c 'import __graft_entry__ as g; g.build()'). "
                "gpsa_b200 has no fallback path."
            )
        torch.ops.load_library(TORCH_LIB_PATH)
        _ops_ns = torch.ops.gpsa_b200
    return _ops_ns


P = C.c_void_p
I = C.c_int
LNG = C.c_long
F = C.c_float


class WarpFwdArgs(C.Structure):
    _fields_ = [
        ("kind", I), ("D", I), ("M", I), ("V", I), ("v", I), ("S", I), ("n", LNG),
        ("Z", P), ("dlt", P), ("log_ls", P), ("log_var", P), ("Omega_G", P), ("hld_Omega", P), ("X", P), ("eps", P),
        ("Lk", P), ("Kinv", P), ("Kinv64", P), ("hld_K", P), ("info", P),
        ("A", P), ("B", P), ("T", P), ("Ke", P), ("var", P), ("Gmean", P), ("Gs", P), ("gs_stride", LNG),
        ("kl_acc", P), ("ws64", P), ("Kuu_ext", P), ("Kuf_ext", P),
    ]


class WarpBwdArgs(C.Structure):
    _fields_ = [
        ("kind", I), ("D", I), ("M", I), ("V", I), ("v", I), ("S", I), ("n", LNG),
        ("Z", P), ("dlt", P), ("log_ls", P), ("log_var", P), ("Omega_G", P), ("X", P), ("eps", P),
        ("Kinv64", P), ("A", P), ("B", P), ("T", P), ("Ke", P),
        ("Gs_bar", P), ("gs_stride", LNG), ("Gm_bar", P), ("kl_bar", P),
        ("acc_Z", P), ("acc_dlt", P), ("acc_hyp", P), ("Obar_G", P),
        ("mubar", P), ("varbar", P), ("q1bar", P), ("Abar", P), ("C", P), ("AS", P),
        ("ws64", P), ("Kuu_bar", P), ("Kuf_bar", P),
    ]


class DataFwdArgs(C.Structure):
    _fields_ = [
        ("kind", I), ("D", I), ("M", I), ("L", I), ("R", LNG),
        ("Gt", P), ("log_ls", P), ("log_var", P), ("dlt", P), ("Omega", P), ("hld_Omega", P), ("G", P),
        ("Lk", P), ("Kinv", P), ("Kinv64", P), ("hld_K", P), ("info", P),
        ("A", P), ("B", P), ("kq", P), ("W", P), ("KD", P), ("mean", P), ("q2", P),
        ("kl_acc", P), ("ws64", P), ("engine", I), ("tc_ws", P), ("tc_ws_bytes", C.c_size_t),
        ("Kuu_ext", P), ("prior_ready", I),
    ]


class DataBwdArgs(C.Structure):
    _fields_ = [
        ("kind", I), ("D", I), ("M", I), ("L", I), ("R", LNG),
        ("Gt", P), ("log_ls", P), ("log_var", P), ("dlt", P), ("Omega", P), ("G", P),
        ("Kinv", P), ("Kinv64", P), ("A", P), ("B", P), ("W", P), ("KD", P),
        ("mean_bar", P), ("q2_bar", P), ("kq_bar", P), ("kl_bar", P),
        ("G_bar", P), ("acc_Gt", P), ("acc_hyp", P), ("dlt_bar", P), ("Obar", P),
        ("q1bar", P), ("Abar", P), ("C", P), ("H", P), ("ws64", P),
        ("engine", I), ("tc_ws", P), ("tc_ws_bytes", C.c_size_t), ("Kuu_bar", P),
    ]


ADAM_MAX_TENSORS = 24


class AdamArgs(C.Structure):
    _fields_ = [
        ("count", I), ("p", P * ADAM_MAX_TENSORS), ("g", P * ADAM_MAX_TENSORS), ("m", P * ADAM_MAX_TENSORS),
        ("v", P * ADAM_MAX_TENSORS), ("n", LNG * ADAM_MAX_TENSORS), ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double),
        ("eps", C.c_double),
        ("step", P),
    ]


# name -> argtypes (restype is int unless listed in _RESTYPE); kept in one table so the CPU test can
# check that every symbol of include/gpsa_b200.h is exported.
SIGNATURES = {
    "gpsa_version": [],
    "gpsa_launch_count": [],
    "gpsa_debug_disable_vec4": [I],
    "gpsa_prof_enable": [I],
    "gpsa_prof_read": [P, P],
    "gpsa_kernel_matrix_fwd": [I, I, I, LNG, P, P, P, P, P, P],
    "gpsa_kernel_matrix_bwd": [I, I, I, LNG, P, P, P, P, P, P, P, P, P, P],
    "gpsa_potrf_batched_f32": [I, I, P, P, P, P],
    "gpsa_potrf_batched_f64": [I, I, P, P, P, P],
    "gpsa_potrf_batched_f32_ld64": [I, I, P, P, P, P, P],
    "gpsa_trtri_batched_f32": [I, I, P, P, P],
    "gpsa_trtri_batched_f64": [I, I, P, P, P],
    "gpsa_gemm_f32": [I, I, LNG, F, P, LNG, LNG, LNG, P, LNG, LNG, LNG, F, P, LNG, LNG, I, P],
    "gpsa_prior_prepare": [I, I, I, P, P, P, P, P, P, P, P, P, P],
    "gpsa_prior_prepare_ext": [I, P, P, P, P, P, P, P, P],
    "gpsa_omega_prepare": [I, I, P, P, P, P, P, P, P],
    "gpsa_omega_grad": [I, I, P, P, P, P, P, P, P, P],
    "gpsa_omega_grad_tc": [I, I, P, P, P, P, P, P, P, P, C.c_size_t, P],
    "gpsa_omega_grad_f32": [I, I, P, P, P, P, P, P, P, P, C.c_size_t, P],
    "gpsa_gemm_tc_ws_bytes": [LNG, LNG, I, I],
    "gpsa_gemm_tc": [LNG, LNG, I, I, P, LNG, LNG, I, P, LNG, LNG, I, P, LNG, LNG, F, I, I, P, C.c_size_t, P],
    "gpsa_feat_count": [I],
    "gpsa_feat_pack": [I, I, P, P, P],
    "gpsa_feat_unpack": [I, I, P, P, F, P, P, P],
    "gpsa_quadform_fwd_f32": [I, LNG, I, P, P, P, P],
    "gpsa_quadform_bwd_omega_f32": [I, LNG, I, P, P, P, P],
    "gpsa_quadform_bwd_alpha_f32": [I, LNG, I, P, P, P, P, P],
    "gpsa_tc_supported": [I],
    "gpsa_quadform_tc_ws_bytes": [I, LNG, I],
    "gpsa_quadform_fwd_feat_tc": [I, LNG, I, P, P, P, P, C.c_size_t, P],
    "gpsa_quadform_bwd_alpha_tc": [I, LNG, I, P, P, P, P, P, C.c_size_t, P],
    "gpsa_quadform_bwd_omega_tc": [I, LNG, I, P, P, P, P, C.c_size_t, P],
    "gpsa_tc_gemm_test": [I, I, I, P, P, P, I, P, C.c_size_t, P],
    "gpsa_warp_view_fwd": [C.POINTER(WarpFwdArgs), P],
    "gpsa_warp_view_bwd": [C.POINTER(WarpBwdArgs), P],
    "gpsa_data_layer_fwd": [C.POINTER(DataFwdArgs), P],
    "gpsa_data_layer_bwd": [C.POINTER(DataBwdArgs), P],
    "gpsa_sample_fwd": [LNG, I, P, P, P, P, P],
    "gpsa_sample_bwd": [LNG, I, P, P, P, P, P, P],
    "gpsa_philox_normal": [LNG, I, I, P, I, I, P, P],
    "gpsa_sample_ll_fused": [LNG, I, I, P, P, P, P, P, I, I, P, P, P, P, P, P],
    "gpsa_scale_if_not_one": [LNG, P, P, P],
    "gpsa_lmc_fwd": [LNG, I, I, P, P, P, P],
    "gpsa_lmc_bwd": [LNG, I, I, P, P, P, P, P, P],
    "gpsa_lmc_max_latent": [],
    "gpsa_lmc_ll_fused": [LNG, I, I, I, P, P, P, P, P, P, P, P, P],
    "gpsa_adam_step": [C.POINTER(AdamArgs), P],
    "gpsa_kmeans_lloyd": [LNG, I, I, P, P, I, P, P, P, P],
    "gpsa_gaussian_ll_fwd": [LNG, I, I, P, P, P, P, P],
    "gpsa_gaussian_ll_bwd": [LNG, I, I, P, P, P, P, P, P, P],
}
_RESTYPE = {"gpsa_feat_count": LNG, "gpsa_launch_count": LNG, "gpsa_prof_enable": None, "gpsa_debug_disable_vec4": None,
            "gpsa_quadform_tc_ws_bytes": C.c_size_t, "gpsa_gemm_tc_ws_bytes": C.c_size_t}


def _declare(l):
    for name, argtypes in SIGNATURES.items():
        fn = getattr(l, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPE.get(name, I)


def check(rc, what):
    if rc != 0:
        raise GPSALibraryError(f"{what} failed: {_ERR.get(rc, rc)}")


def stream(device=None):
    """cudaStream_t of torch's current stream on `device` (default: the current device).  The autograd Functions in
    _ops.py run under torch.cuda.device(<their tensors' device>), so the two agree."""
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def on_device_of(argpos=0):
    """Decorator for autograd.Function.forward/backward: run the body with the CUDA device of the first tensor
    argument current, so that allocations, the stream and the library's per-device state all refer to the device
    the data lives on (a model on cuda:1 in a process whose current device is cuda:0)."""
    import functools

    def deco(fn):
        @functools.wraps(fn)
        def wrapped(*args):
            dev = None
            for a in args[argpos:]:
                if torch.is_tensor(a) and a.is_cuda:
                    dev = a.device
                    break
            if dev is None:
                for t in getattr(args[0], "saved_tensors", ()) if argpos else ():
                    if torch.is_tensor(t) and t.is_cuda:
                        dev = t.device
                        break
            if dev is None:
                return fn(*args)
            with torch.cuda.device(dev):
                return fn(*args)
        return wrapped
    return deco


def ptr(t, dtype=torch.float32):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise GPSALibraryError("gpsa_b200 is CUDA-only: got a CPU tensor (no CPU fallback exists)")
    if t.dtype != dtype:
        raise GPSALibraryError(f"expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise GPSALibraryError("expected a contiguous tensor")
    return t.data_ptr()


def feat_count(M):
    return int(lib().gpsa_feat_count(int(M)))


def tc_workspace(M, R, L, like):
    """Scratch for the tcgen05 engine (bf16 hi/lo copies of the operands), as a uint8 CUDA tensor."""
    n = int(lib().gpsa_quadform_tc_ws_bytes(int(M), int(R), int(L)))
    return torch.empty(n, dtype=torch.uint8, device=like.device)
