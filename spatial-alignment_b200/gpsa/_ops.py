"""torch.autograd.Function wrappers with EXPLICIT forward and backward around the C ABI
(include/gpsa_b200.h).  Nothing below this file uses autograd; every gradient is the analytic
backward implemented in CUDA."""
import ctypes as C

import torch

from . import _lib
from ._lib import DataBwdArgs, DataFwdArgs, WarpBwdArgs, WarpFwdArgs, check, lib, on_device_of, ptr, stream

f32, f64, i32 = torch.float32, torch.float64, torch.int32

KINDS = {"rbf": _lib.KIND_RBF, "matern12": _lib.KIND_MATERN12, "matern32": _lib.KIND_MATERN32}

# quadratic-form engine: 0 = fp32 SIMT, 1 = tcgen05 split-bf16 with the ||a^T L||^2 forward, 2 = tcgen05 split-bf16
# with the implicit-feature forward (all three products on the same generator GEMM core); "auto" = engine 2 whenever
# the shape can fill 128-row MMA tiles (the SIMT engine stays for the launch-bound toy configurations)
ENGINE = {"value": "auto"}
TC_ENGINES = (1, 2)


def pick_engine(M, R, L):
    e = ENGINE["value"]
    if e == "auto":
        return 2 if (lib().gpsa_tc_supported(int(M)) and R >= 2048 and L >= 16 and M >= 32) else 0
    if e in TC_ENGINES and not lib().gpsa_tc_supported(int(M)):
        raise _lib.GPSALibraryError(f"the tcgen05 quadratic-form engine does not cover M={M} yet (16 <= M <= 512)")
    return int(e)


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


def _new(like, *shape, dtype=f32):
    return torch.empty(shape, dtype=dtype, device=like.device)


def _zeros(like, *shape, dtype=f32):
    return torch.zeros(shape, dtype=dtype, device=like.device)


_SIDE = {}


def _side_streams(device, n):
    """A small pool of side streams per device: the views of the warp layer are independent and each is a chain of
    small latency-bound kernels (M x M factorisation, M x n products), so they run concurrently and are joined back
    into the caller's stream."""
    key = (device.index if device.index is not None else torch.cuda.current_device())
    pool = _SIDE.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=device))
    return pool[:n]


def check_info(info, what):
    """Raise like torch.cholesky does when a matrix was not positive definite (this reads a device
    flag, i.e. it synchronises; callers gate it behind GPSA_B200_CHECK=1 / debug mode)."""
    bad = torch.nonzero(info)
    if bad.numel():
        raise RuntimeError(f"{what}: matrix {int(bad[0, 0])} is not positive-definite")


# --------------------------------------------------------------------------------------------------
class KernelMatrix(torch.autograd.Function):
    """K[m, r] = k(x1[m], x2[r]) for x1 [M,D], x2 [R,D]; differentiable in all four tensors."""

    @staticmethod
    @on_device_of(1)
    def forward(ctx, kind, x1, x2, log_ls, log_var):
        x1, x2 = _c(x1.detach()), _c(x2.detach())
        log_ls, log_var = _c(log_ls.detach().reshape(1)), _c(log_var.detach().reshape(1))
        M, D = x1.shape
        R = x2.shape[0]
        if x2.shape[1] != D or not 1 <= D <= 3:
            raise ValueError("kernel inputs must share a trailing dimension of 1, 2 or 3")
        K = _new(x1, M, R)
        check(lib().gpsa_kernel_matrix_fwd(kind, D, M, R, ptr(x1), ptr(x2), ptr(log_ls), ptr(log_var), ptr(K), stream()),
              "kernel_matrix_fwd")
        ctx.kind = kind
        ctx.save_for_backward(x1, x2, log_ls, log_var)
        return K

    @staticmethod
    @on_device_of(1)
    def backward(ctx, Kbar):
        x1, x2, log_ls, log_var = ctx.saved_tensors
        M, D = x1.shape
        R = x2.shape[0]
        Kbar = _c(Kbar)
        acc_x1 = _zeros(x1, M, D, dtype=f64)
        acc_h = _zeros(x1, 2, dtype=f64)
        x2bar = _new(x1, R, D)
        check(lib().gpsa_kernel_matrix_bwd(ctx.kind, D, M, R, ptr(x1), ptr(x2), ptr(log_ls), ptr(log_var), ptr(Kbar),
                                           ptr(acc_x1, f64), ptr(x2bar), None, ptr(acc_h, f64), stream()),
              "kernel_matrix_bwd")
        h = acc_h.to(f32)
        return None, acc_x1.to(f32), x2bar, h[0:1], h[1:2]


def kernel_matrix(kind_name, x1, x2, log_ls, log_var):
    """Batch-broadcasting front end used by gpsa.rbf_kernel / gpsa.matern12_kernel:
    x1 [n1,D], x2 [..., n2, D] -> [..., n1, n2]  (reference gpsa/util/util.py:18-23 semantics)."""
    kind = KINDS[kind_name]
    if x1.dim() != 2:
        if x1.shape[:-2] == x2.shape[:-2] and x1.dim() == x2.dim():
            flat1, flat2 = x1.reshape(-1, *x1.shape[-2:]), x2.reshape(-1, *x2.shape[-2:])
            out = torch.stack([kernel_matrix(kind_name, a, b, log_ls, log_var) for a, b in zip(flat1, flat2)])
            return out.reshape(*x1.shape[:-2], x1.shape[-2], x2.shape[-2])
        raise NotImplementedError("x1 must be [n1, D] or share x2's batch dimensions")
    batch = x2.shape[:-2]
    n2, D = x2.shape[-2:]
    if log_ls.numel() != 1 or log_var.numel() != 1:
        raise ValueError("gpsa_b200's fused covariance functions take ONE lengthscale and ONE output variance "
                         f"(got {log_ls.numel()} and {log_var.numel()} elements); per-dimension lengthscales are not supported")
    K = KernelMatrix.apply(kind, x1, x2.reshape(-1, D), log_ls.reshape(1), log_var.reshape(1))
    if len(batch) == 0:
        return K
    n1 = x1.shape[0]
    return K.reshape(n1, -1, n2).movedim(0, 1).reshape(*batch, n1, n2)


# --------------------------------------------------------------------------------------------------
def omega_prepare(Osq):
    """Omega = Osq Osq^T + 1e-5 I (fp32 copy), its Cholesky factor (fp32 copy and the fp64 original),
    fp64 half log-dets, info flags."""
    with torch.cuda.device(Osq.device):
        return _omega_prepare(Osq)


def _omega_prepare(Osq):
    B, M, _ = Osq.shape
    Omega, Ltril = _new(Osq, B, M, M), _new(Osq, B, M, M)
    L64 = _new(Osq, B, M, M, dtype=f64)
    hld = _new(Osq, B, dtype=f64)
    info = _new(Osq, B, dtype=i32)
    check(lib().gpsa_omega_prepare(M, B, ptr(Osq), ptr(Omega), ptr(Ltril), ptr(L64, f64), ptr(hld, f64),
                                   ptr(info, i32), stream()), "omega_prepare")
    return Omega, Ltril, L64, hld, info


class OmegaFromSqt(torch.autograd.Function):
    """Omega = Osq Osq^T + 1e-5 I as a differentiable op (reference gpsa/models/vgpsa.py:206-210 is plain autograd)."""

    @staticmethod
    @on_device_of(1)
    def forward(ctx, Osq):
        Osq = _c(Osq.detach())
        ctx.save_for_backward(Osq)
        return _omega_prepare(Osq)[0]

    @staticmethod
    @on_device_of(1)
    def backward(ctx, Obar):
        (Osq,) = ctx.saved_tensors
        sym = _c(0.5 * (Obar + Obar.transpose(-1, -2)))  # omega_grad expects a symmetric Omega-bar
        return omega_grad(Osq, None, sym, None)


def omega_grad(Osq, L64, Obar, coef, tc=False):
    """Osq_bar = 2 (Obar + coef Omega^-1) Osq; tc=True runs the 2 Obar Osq product on the tcgen05 engine."""
    B, M, _ = Osq.shape
    Linv = _new(Osq, B, M, M, dtype=f64) if coef is not None else None
    Y = _new(Osq, B, M, M, dtype=f64) if coef is not None else None
    out = _new(Osq, B, M, M)
    ws = None
    if tc and M >= 32:
        ws = torch.empty(int(lib().gpsa_gemm_tc_ws_bytes(M, M, M, B)), dtype=torch.uint8, device=Osq.device)
    check(lib().gpsa_omega_grad_tc(M, B, ptr(Osq), ptr(L64, f64), ptr(Obar), ptr(coef), ptr(Linv, f64), ptr(Y, f64),
                                   ptr(out), ptr(ws, torch.uint8), ws.numel() if ws is not None else 0, stream()),
          "omega_grad")
    return out


# --------------------------------------------------------------------------------------------------
class WarpLayer(torch.autograd.Function):
    """All non-fixed views of the warp GP (reference gpsa/models/vgpsa.py:255-351, KL :498-516).

    forward(meta, Xtilde, delta_G, Omega_sqt_G, log_ls, log_var, *[X_v, eps_v for each free view])
      -> (KL_G, Kuu_chol_list, Omega_tril_G, info, *[Gmean_v, Gsamples_v ...])
    meta = dict(kind=int, V=int, S=int, free=[view indices], with_kl=bool)
    """

    @staticmethod
    @on_device_of(1)
    def forward(ctx, meta, Xtilde, delta_G, Osq_G, log_ls, log_var, *xe):
        Xtilde, delta_G, Osq_G = _c(Xtilde.detach()), _c(delta_G.detach()), _c(Osq_G.detach())
        log_ls, log_var = _c(log_ls.detach()), _c(log_var.detach())
        V, M, D = Xtilde.shape
        S, free, kind = meta["S"], meta["free"], meta["kind"]
        Omega_G, Ltril_G, L64_G, hld_G, info_G = omega_prepare(Osq_G)
        kl = _zeros(Xtilde, 1, dtype=f64)
        Lk_all = torch.full((V, M, M), float("nan"), dtype=f32, device=Xtilde.device)  # NaN rows for fixed views (:237-242)
        info = _zeros(Xtilde, V, dtype=i32)
        hldK = _zeros(Xtilde, V, dtype=f64)
        saved, outs = [], []
        cur = torch.cuda.current_stream()
        # meta["overlap"]: work that does not depend on the warp layer (the caller passes the preparation of the
        # gene-batched Omega_F: large kernels).  It is enqueued on the caller's stream between fork and join, so the
        # views' latency-bound chains of small kernels run underneath it.
        overlap = meta.get("overlap")
        side = _side_streams(Xtilde.device, min(len(free), 4)) if (len(free) > 1 or (overlap and free)) else []
        keep = []
        for k, v in enumerate(free):
            X, eps = _c(xe[2 * k].detach()), _c(xe[2 * k + 1].detach())
            n = X.shape[0]
            ws64 = _new(Xtilde, 2 * M * M, dtype=f64)
            keep.append(ws64)
            Kinv = _new(X, M, M, dtype=f64)
            A, B, T = _new(X, M, n, dtype=f64), _new(X, M, n, dtype=f64), _new(X, D, M, n, dtype=f64)
            Ke, var = _new(X, D, M, dtype=f64), _new(X, n, D)
            Gmean, Gs = _new(X, n, D), _new(X, S, n, D)
            if n > 0:
                a = WarpFwdArgs(kind=kind, D=D, M=M, V=V, v=v, S=S, n=n,
                                Z=ptr(Xtilde) + 4 * v * M * D, dlt=ptr(delta_G) + 4 * v * M * D,
                                log_ls=ptr(log_ls) + 4 * v, log_var=ptr(log_var) + 4 * v,
                                Omega_G=ptr(Omega_G), hld_Omega=ptr(hld_G, f64), X=ptr(X), eps=ptr(eps),
                                Lk=ptr(Lk_all) + 4 * v * M * M, Kinv=None, Kinv64=ptr(Kinv, f64),
                                hld_K=ptr(hldK, f64) + 8 * v, info=ptr(info, i32) + 4 * v, A=ptr(A, f64),
                                B=ptr(B, f64), T=ptr(T, f64), Ke=ptr(Ke, f64), var=ptr(var),
                                Gmean=ptr(Gmean), Gs=ptr(Gs), gs_stride=n * D,
                                kl_acc=ptr(kl, f64) if meta["with_kl"] else None, ws64=ptr(ws64, f64))
                st = stream()
                if side:
                    # fork HERE: everything this view reads (copies, zero fills) has been enqueued on `cur` by now
                    side[k % len(side)].wait_stream(cur)
                    st = C.c_void_p(side[k % len(side)].cuda_stream)
                check(lib().gpsa_warp_view_fwd(C.byref(a), st), "warp_view_fwd")
            # `var` is written on the side stream and used nowhere else: it has to outlive the join below, or its block
            # returns to the current stream's allocator pool while this view's chain is still running
            keep.append(var)
            saved += [X, eps, Kinv, A, B, T, Ke]
            outs += [Gmean, Gs]
        if overlap:
            overlap()
        for s_ in side:
            cur.wait_stream(s_)
        del keep
        ctx.meta = meta
        ctx.save_for_backward(Xtilde, delta_G, Osq_G, log_ls, log_var, Omega_G, L64_G, *saved)
        info_all = torch.cat([info_G, info])
        ctx.mark_non_differentiable(Lk_all, Ltril_G, info_all)
        return (kl.to(f32).reshape(()), Lk_all, Ltril_G, info_all, *outs)

    @staticmethod
    @on_device_of(1)
    def backward(ctx, kl_bar, _1, _2, _3, *gouts):
        meta = ctx.meta
        Xtilde, delta_G, Osq_G, log_ls, log_var, Omega_G, L64_G, *saved = ctx.saved_tensors
        V, M, D = Xtilde.shape
        S, free, kind = meta["S"], meta["free"], meta["kind"]
        dev = Xtilde
        use_kl = meta["with_kl"] and kl_bar is not None
        klb = _c(kl_bar.to(f32).reshape(1)) if use_kl else None
        acc_Z, acc_dlt = _zeros(dev, V, M, D, dtype=f64), _zeros(dev, V, M, D, dtype=f64)
        acc_hyp = _zeros(dev, V, 2, dtype=f64)
        xgrads = []
        cur = torch.cuda.current_stream()
        live = [k for k in range(len(free)) if saved[7 * k].shape[0] > 0]
        side = _side_streams(dev.device, min(len(live), 4)) if len(live) > 1 else []
        # views add into overlapping Omega-bar slices (v*D+j and j*V+v): one buffer per concurrent view, summed at the join
        Obars, keep = [], []
        for k, v in enumerate(free):
            X, eps, Kinv, A, B, T, Ke = saved[7 * k: 7 * k + 7]
            n = X.shape[0]
            gm, gs = gouts[2 * k], gouts[2 * k + 1]
            xgrads += [None, None]
            if n == 0:
                continue
            lane = len(Obars) % len(side) if side else 0
            if side or not Obars:
                Obars.append(_zeros(dev, V * D, M, M))
            Obar = Obars[-1]
            ws64 = _new(dev, 3 * M * M, dtype=f64)
            gm = _c(gm) if gm is not None else None
            gs = _c(gs) if gs is not None else None
            mubar, varbar, q1bar = _new(dev, n, D), _new(dev, n, D), _new(dev, n)
            Abar, Cm, AS = (_new(dev, M, n, dtype=f64), _new(dev, M, n, dtype=f64),
                            _new(dev, D, M, n, dtype=f64))
            a = WarpBwdArgs(kind=kind, D=D, M=M, V=V, v=v, S=S, n=n,
                            Z=ptr(Xtilde) + 4 * v * M * D, dlt=ptr(delta_G) + 4 * v * M * D,
                            log_ls=ptr(log_ls) + 4 * v, log_var=ptr(log_var) + 4 * v, Omega_G=ptr(Omega_G),
                            X=ptr(X), eps=ptr(eps), Kinv64=ptr(Kinv, f64), A=ptr(A, f64), B=ptr(B, f64),
                            T=ptr(T, f64), Ke=ptr(Ke, f64),
                            Gs_bar=ptr(gs), gs_stride=n * D, Gm_bar=ptr(gm), kl_bar=ptr(klb),
                            acc_Z=ptr(acc_Z, f64) + 8 * v * M * D, acc_dlt=ptr(acc_dlt, f64) + 8 * v * M * D,
                            acc_hyp=ptr(acc_hyp, f64) + 16 * v, Obar_G=ptr(Obar),
                            mubar=ptr(mubar), varbar=ptr(varbar), q1bar=ptr(q1bar), Abar=ptr(Abar, f64),
                            C=ptr(Cm, f64), AS=ptr(AS, f64), ws64=ptr(ws64, f64))
            st = stream()
            if side:
                # fork HERE, after this view's zero fills / contiguous copies were enqueued on `cur`
                side[lane].wait_stream(cur)
                st = C.c_void_p(side[lane].cuda_stream)
            check(lib().gpsa_warp_view_bwd(C.byref(a), st), "warp_view_bwd")
            keep += [gm, gs, mubar, varbar, q1bar, Abar, Cm, AS, ws64]
        for s_ in side:
            cur.wait_stream(s_)
        del keep
        if not Obars:
            Obar = _zeros(dev, V * D, M, M)
        elif len(Obars) == 1:
            Obar = Obars[0]
        else:
            Obar = torch.stack(Obars).sum(0)
        coef = None
        if use_kl:
            # d(-half_logdet Omega_{j*V+v})/dOmega = -1/2 Omega^-1 on the free views' KL slices only;
            # meta["kl_mask"] holds -0.5 there and 0 elsewhere (device tensor, built once by the model)
            coef = _c(meta["kl_mask"] * klb)
        Osq_bar = omega_grad(Osq_G, L64_G, Obar, coef)
        hyp = acc_hyp.to(f32)
        return (None, acc_Z.to(f32), acc_dlt.to(f32), Osq_bar, _c(hyp[:, 0]), _c(hyp[:, 1]), *xgrads)


# --------------------------------------------------------------------------------------------------
class DataLayer(torch.autograd.Function):
    """One modality of the data GP (reference gpsa/models/vgpsa.py:390-426, KL :520-530).

    forward(meta, Gtilde, log_ls, log_var, delta_F, Omega_sqt_F, G [S,N,D], eps [S,N,L])
      -> (F_latent [S,N,L], KL_F, Kuu_chol_F, Omega_tril_F, info)
    """

    @staticmethod
    @on_device_of(1)
    def forward(ctx, meta, Gtilde, log_ls, log_var, delta_F, Osq_F, G, eps):
        Gtilde, delta_F, Osq_F = _c(Gtilde.detach()), _c(delta_F.detach()), _c(Osq_F.detach())
        log_ls, log_var = _c(log_ls.detach().reshape(1)), _c(log_var.detach().reshape(1))
        G, eps = _c(G.detach()), _c(eps.detach())
        M, D = Gtilde.shape
        L = delta_F.shape[1]
        S, N = G.shape[0], G.shape[1]
        R = S * N
        kind = meta["kind"]
        pre = meta.get("omega")
        Omega, Ltril, L64, hld, info_O = pre if pre is not None else omega_prepare(Osq_F)
        Lk, Kinv, Kinv64 = _new(G, M, M), _new(G, M, M), _new(G, M, M, dtype=f64)
        hldK = _zeros(G, 1, dtype=f64)
        info = _zeros(G, 1, dtype=i32)
        A, B, kq = _new(G, M, R), _new(G, M, R), _new(G, R)
        engine = pick_engine(M, R, L)
        W = _new(G, _lib.feat_count(M), L) if engine == 0 else _new(G, 1)
        tc_ws = _lib.tc_workspace(M, R, L, G) if engine in TC_ENGINES else None
        KD = _new(G, M, L, dtype=f64)
        Fo, var = _new(G, S, N, L), _new(G, R, L)
        kl = _zeros(G, 1, dtype=f64)
        ws64 = _new(G, 2 * M * M, dtype=f64)
        a = DataFwdArgs(kind=kind, D=D, M=M, L=L, R=R, Gt=ptr(Gtilde), log_ls=ptr(log_ls), log_var=ptr(log_var),
                        dlt=ptr(delta_F), Omega=ptr(Omega), hld_Omega=ptr(hld, f64), G=ptr(G), eps=ptr(eps),
                        Lk=ptr(Lk), Kinv=ptr(Kinv), Kinv64=ptr(Kinv64, f64), hld_K=ptr(hldK, f64),
                        info=ptr(info, i32), A=ptr(A), B=ptr(B), kq=ptr(kq), W=ptr(W), KD=ptr(KD, f64), F=ptr(Fo),
                        var=ptr(var),
                        kl_acc=ptr(kl, f64) if meta["with_kl"] else None, ws64=ptr(ws64, f64),
                        engine=engine, Ltril=ptr(Ltril), tc_ws=ptr(tc_ws, torch.uint8),
                        tc_ws_bytes=tc_ws.numel() if tc_ws is not None else 0)
        check(lib().gpsa_data_layer_fwd(C.byref(a), stream()), "data_layer_fwd")
        ctx.meta = meta
        ctx.engine = engine
        ctx.save_for_backward(Gtilde, log_ls, log_var, delta_F, Osq_F, G, eps, Omega, L64, Kinv, Kinv64, A, B, W, KD,
                              var)
        info_all = torch.cat([info_O, info])
        ctx.mark_non_differentiable(Lk, Ltril, info_all)
        return Fo, kl.to(f32).reshape(()), Lk, Ltril, info_all

    @staticmethod
    @on_device_of(1)
    def backward(ctx, F_bar, kl_bar, _1, _2, _3):
        meta = ctx.meta
        (Gtilde, log_ls, log_var, delta_F, Osq_F, G, eps, Omega, L64, Kinv, Kinv64, A, B, W, KD,
         var) = ctx.saved_tensors
        M, D = Gtilde.shape
        L = delta_F.shape[1]
        S, N = G.shape[0], G.shape[1]
        R = S * N
        dev = G
        use_kl = meta["with_kl"] and kl_bar is not None
        klb = _c(kl_bar.to(f32).reshape(1)) if use_kl else None
        F_bar = _c(F_bar) if F_bar is not None else _zeros(dev, S, N, L)
        G_bar = _new(dev, S, N, D)
        acc_Gt, acc_hyp = _zeros(dev, M, D, dtype=f64), _zeros(dev, 2, dtype=f64)
        dlt_bar, Obar = _new(dev, M, L), _new(dev, L, M, M)
        Gm, q1bar = _new(dev, R, L), _new(dev, R)
        Abar, Cm = _new(dev, M, R), _new(dev, M, R)
        engine = ctx.engine
        H = _new(dev, _lib.feat_count(M), L)
        tc_ws = _lib.tc_workspace(M, R, L, dev) if engine in TC_ENGINES else None
        ws64 = _new(dev, 3 * M * M, dtype=f64)
        a = DataBwdArgs(kind=meta["kind"], D=D, M=M, L=L, R=R, Gt=ptr(Gtilde), log_ls=ptr(log_ls),
                        log_var=ptr(log_var), dlt=ptr(delta_F), Omega=ptr(Omega), G=ptr(G), eps=ptr(eps),
                        Kinv=ptr(Kinv), Kinv64=ptr(Kinv64, f64), A=ptr(A), B=ptr(B), W=ptr(W), KD=ptr(KD, f64),
                        var=ptr(var),
                        F_bar=ptr(F_bar), kl_bar=ptr(klb), G_bar=ptr(G_bar), acc_Gt=ptr(acc_Gt, f64),
                        acc_hyp=ptr(acc_hyp, f64), dlt_bar=ptr(dlt_bar), Obar=ptr(Obar), Gm=ptr(Gm),
                        q1bar=ptr(q1bar), Abar=ptr(Abar), C=ptr(Cm), H=ptr(H), ws64=ptr(ws64, f64),
                        engine=engine, tc_ws=ptr(tc_ws, torch.uint8),
                        tc_ws_bytes=tc_ws.numel() if tc_ws is not None else 0)
        check(lib().gpsa_data_layer_bwd(C.byref(a), stream()), "data_layer_bwd")
        coef = _c((-0.5 * klb).expand(L)) if use_kl else None
        del tc_ws
        Osq_bar = omega_grad(Osq_F, L64, Obar, coef, tc=(engine in TC_ENGINES))
        hyp = acc_hyp.to(f32)
        return None, acc_Gt.to(f32), hyp[0:1], hyp[1:2], dlt_bar, Osq_bar, G_bar, None


# --------------------------------------------------------------------------------------------------
class GaussianLL(torch.autograd.Function):
    """sum log N(Y; F, sigma) / S with sigma = exp(log_noise) + 1e-5 as the Normal scale
    (reference gpsa/models/vgpsa.py:217, :532-538).  F [S,N,P], Y [N,P], log_noise: 1 element."""

    @staticmethod
    @on_device_of(1)
    def forward(ctx, F, Y, log_noise):
        F, Y, log_noise = _c(F.detach()), _c(Y.detach()), _c(log_noise.detach().reshape(1))
        S, N, P = F.shape
        if Y.shape != (N, P):
            raise ValueError(f"outputs have shape {tuple(Y.shape)}, F_samples imply {(N, P)}")
        acc = _zeros(F, 1, dtype=f64)
        check(lib().gpsa_gaussian_ll_fwd(N, P, S, ptr(F), ptr(Y), ptr(log_noise), ptr(acc, f64), stream()), "ll_fwd")
        ctx.save_for_backward(F, Y, log_noise)
        return acc.to(f32).reshape(())

    @staticmethod
    @on_device_of(1)
    def backward(ctx, ll_bar):
        F, Y, log_noise = ctx.saved_tensors
        S, N, P = F.shape
        llb = _c(ll_bar.to(f32).reshape(1))
        F_bar = _new(F, S, N, P)
        acc = _zeros(F, 1, dtype=f64)
        check(lib().gpsa_gaussian_ll_bwd(N, P, S, ptr(F), ptr(Y), ptr(log_noise), ptr(llb), ptr(F_bar),
                                         ptr(acc, f64), stream()), "ll_bwd")
        return F_bar, None, acc.to(f32)
