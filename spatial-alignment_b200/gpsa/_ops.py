"""torch.autograd.Function wrappers with EXPLICIT forward and backward around the library's torch custom ops
(torch.ops.gpsa_b200.*, csrc/bindings.cpp: one op per extern "C" launcher of include/gpsa_b200.h).  Nothing below this
file uses autograd; every gradient is the analytic backward implemented in CUDA.  The ops check device / dtype /
contiguity and launch on torch's current stream of the tensors' device."""
import contextlib

import torch

from . import _lib
from ._lib import lib, on_device_of, ops

f32, f64, i32 = torch.float32, torch.float64, torch.int32

KINDS = {"rbf": _lib.KIND_RBF, "matern12": _lib.KIND_MATERN12, "matern32": _lib.KIND_MATERN32}

# quadratic-form engine: 0 = fp32 SIMT, 1 = tcgen05 (bf16 hi/lo split, three passes, all three products as implicit-
# feature GEMMs); "auto" = tcgen05 whenever the shape can fill 128-row MMA tiles (the SIMT engine stays for the
# launch-bound toy configurations)
ENGINE = {"value": "auto"}
TC_ENGINES = (1,)


def pick_engine(M, R, L):
    e = ENGINE["value"]
    if e == "auto":
        return 1 if (lib().gpsa_tc_supported(int(M)) and R >= 2048 and L >= 16 and M >= 32) else 0
    if e in TC_ENGINES and not lib().gpsa_tc_supported(int(M)):
        raise _lib.GPSALibraryError(f"the tcgen05 quadratic-form engine does not cover M={M} (16 <= M <= 512)")
    return int(e)


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


def _new(like, *shape, dtype=f32):
    return torch.empty(shape, dtype=dtype, device=like.device)


def _zeros(like, *shape, dtype=f32):
    return torch.zeros(shape, dtype=dtype, device=like.device)


_SIDE = {}
_OMEGA_STREAM = {}


def omega_stream(device, which="omega"):
    """The stream the gene-batched variational covariances (OmegaChain) are prepared and differentiated on;
    which="prior": the stream the data layer's K_uu is factorised on ahead of the layer (prior_prepare)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), which)
    if key not in _OMEGA_STREAM:
        _OMEGA_STREAM[key] = torch.cuda.Stream(device=device)
        # Omega_sqt_F's gradient is produced on this stream and accumulated on the parameter's: intended (autograd
        # synchronises the two), so the advisory about it is switched off
        quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if quiet is not None:
            quiet(False)
    return _OMEGA_STREAM[key]


def prior_prepare(kind, Z, log_ls, log_var):
    """fp64 factorisation of k(Z,Z) + 1e-5 I (reference gpsa/models/vgpsa.py:390-394) on the current stream:
    (Lk fp32, K^-1 fp32, K^-1 fp64, half log-det [1] fp64, info [1]) -- what DataLayerPre takes as meta["prior"].  K_uu
    depends on parameters only, so the model factorises it while the warp layer runs."""
    Z = _c(Z.detach())
    M, D = Z.shape
    with torch.cuda.device(Z.device):
        Lk, Kinv, Kinv64 = _new(Z, M, M), _new(Z, M, M), _new(Z, M, M, dtype=f64)
        hldK, info = _zeros(Z, 1, dtype=f64), _zeros(Z, 1, dtype=i32)
        ws64 = _new(Z, 2 * M * M, dtype=f64)
        ops().prior_prepare(kind, D, M, Z, _c(log_ls.detach().reshape(1)), _c(log_var.detach().reshape(1)), Lk, Kinv, Kinv64,
                            hldK, info, ws64)
    return Lk, Kinv, Kinv64, hldK, info


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
        ops().kernel_matrix_fwd(kind, D, M, R, x1, x2, log_ls, log_var, K)
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
        ops().kernel_matrix_bwd(ctx.kind, D, M, R, x1, x2, log_ls, log_var, Kbar, acc_x1, x2bar, None, acc_h)
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


# Batches of at least this many M x M variational covariances (M >= 32) are factorised in fp32 (two CTAs per SM, half
# the bytes); smaller batches -- the warp layer's V*D matrices, toy gene counts -- stay in fp64.  Omega itself is
# accumulated in fp64 either way.
OMEGA_F32_MIN_BATCH = 16


def omega_uses_f32(B, M):
    return B >= OMEGA_F32_MIN_BATCH and M >= 32


def _omega_prepare(Osq):
    B, M, _ = Osq.shape
    Omega, Ltril = _new(Osq, B, M, M), _new(Osq, B, M, M)
    L64 = None if omega_uses_f32(B, M) else _new(Osq, B, M, M, dtype=f64)
    hld = _new(Osq, B, dtype=f64)
    info = _new(Osq, B, dtype=i32)
    ops().omega_prepare(M, B, Osq, Omega, Ltril, L64, hld, info)
    # fp32 branch: the fp32 factor itself is what the backward inverts
    return Omega, Ltril, (L64 if L64 is not None else Ltril), hld, info


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


class OmegaChain(torch.autograd.Function):
    """The variational covariances of one modality as their own autograd node (reference gpsa/models/vgpsa.py:206-210,
    :410-412):  Osq [L,M,M] -> Omega = Osq Osq^T + 1e-5 I and half log-dets hld [L] (float64), both differentiable, plus
    the factor (Ltril fp32, L64 fp64 or None when the batch is factorised in fp32) and the info flags.

    Being a node of its own, it runs on whatever stream is current when it is applied -- the model applies it on a side
    stream -- and autograd runs its backward on that same stream: the gene-batched factorisation overlaps the warp
    layer's forward, and Osq_bar = 2 (Obar + hld_bar/2 Omega^-1) Osq (trtri + three batched GEMMs) overlaps the warp
    layer's backward.  meta["tc"] (set by the consumer before the backward) selects the tcgen05 engine for 2 Obar Osq."""

    @staticmethod
    @on_device_of(2)
    def forward(ctx, meta, Osq):
        Osq = _c(Osq.detach())
        Omega, Ltril, Lfac, hld, info = _omega_prepare(Osq)
        ctx.meta = meta
        # no zero tensors for the gradients of outputs nothing differentiates (autograd would fill an [L,M,M] buffer for
        # the factor on every backward); the backward handles None
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(Osq, Lfac)
        L64 = Lfac if Lfac.dtype == f64 else None
        ctx.mark_non_differentiable(Ltril, info, *([L64] if L64 is not None else []))
        return Omega, hld, Ltril, L64, info

    @staticmethod
    @on_device_of(2)
    def backward(ctx, Obar, hld_bar, _1, _2, _3):
        Osq, Lfac = ctx.saved_tensors
        cur = torch.cuda.current_stream()
        if Obar is None:
            Obar = _zeros(Osq, *Osq.shape)
        else:
            Obar = _c(Obar)
            Obar.record_stream(cur)  # produced on the consumer's stream, read here on this node's
        coef = None
        if hld_bar is not None:
            hld_bar.record_stream(cur)
            coef = _c((0.5 * hld_bar).to(f32))  # d hld / d Omega = 1/2 Omega^-1
        return None, omega_grad(Osq, Lfac, Obar, coef, tc=bool(ctx.meta.get("tc", False)))


def omega_grad(Osq, Lfac, Obar, coef, tc=False):
    """Osq_bar = 2 (Obar + coef Omega^-1) Osq; tc=True runs the 2 Obar Osq product on the tcgen05 engine.
    Lfac: the factor omega_prepare returned for the backward (fp64, or fp32 for large batches); None if coef is None."""
    B, M, _ = Osq.shape
    out = _new(Osq, B, M, M)
    ws = None
    if tc and M >= 32:
        ws = torch.empty(int(lib().gpsa_gemm_tc_ws_bytes(M, M, M, B)), dtype=torch.uint8, device=Osq.device)
    nws = ws.numel() if ws is not None else 0
    if Lfac is not None and Lfac.dtype == f32:
        Linv = _new(Osq, B, M, M) if coef is not None else None
        Y = _new(Osq, B, M, M) if coef is not None else None
        ops().omega_grad_f32(M, B, Osq, Lfac, Obar, coef, Linv, Y, out, ws, nws)
        return out
    Linv = _new(Osq, B, M, M, dtype=f64) if coef is not None else None
    Y = _new(Osq, B, M, M, dtype=f64) if coef is not None else None
    ops().omega_grad_tc(M, B, Osq, Lfac, Obar, coef, Linv, Y, out, ws, nws)
    return out


# --------------------------------------------------------------------------------------------------
class WarpLayer(torch.autograd.Function):
    """All non-fixed views of the warp GP (reference gpsa/models/vgpsa.py:255-351, KL :498-516).

    forward(meta, Xtilde, delta_G, Omega_sqt_G, log_ls, log_var, *[X_v, eps_v for each free view])
      -> (KL_G, Kuu_chol_list, Omega_tril_G, info, *[Gmean_v, Gsamples_v ...])
    meta = dict(kind=int, V=int, S=int, free=[view indices], with_kl=bool)
    kind = KIND_EXTERNAL (user-supplied covariance callable): the per-view inputs are [X_v, eps_v, Kuu_v [M,M],
    Kuf_v [M,n_v]] -- k(Z,Z) and k(Z,X) evaluated by the caller with torch -- and the backward returns dLoss/dK for both.
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
        side = _side_streams(Xtilde.device, min(len(free), 4)) if len(free) > 1 else []
        keep = []
        ext = kind == _lib.KIND_EXTERNAL
        nper = 4 if ext else 2
        for k, v in enumerate(free):
            X, eps = _c(xe[nper * k].detach()), _c(xe[nper * k + 1].detach())
            Kuu_e = _c(xe[nper * k + 2].detach().to(f32)) if ext else None
            Kuf_e = _c(xe[nper * k + 3].detach().to(f32)) if ext else None
            n = X.shape[0]
            ws64 = _new(Xtilde, 2 * M * M, dtype=f64)
            keep.append(ws64)
            Kinv = _new(X, M, M, dtype=f64)
            A, B, T = _new(X, M, n, dtype=f64), _new(X, M, n, dtype=f64), _new(X, D, M, n, dtype=f64)
            Ke, var = _new(X, D, M, dtype=f64), _new(X, n, D)
            Gmean, Gs = _new(X, n, D), _new(X, S, n, D)
            if n > 0:
                keep += [Kuu_e, Kuf_e]
                ctx_stream = contextlib.nullcontext()
                if side:
                    # fork HERE: everything this view reads (copies, zero fills) has been enqueued on `cur` by now
                    side[k % len(side)].wait_stream(cur)
                    ctx_stream = torch.cuda.stream(side[k % len(side)])
                with ctx_stream:  # the op launches on the current stream; it allocates nothing
                    ops().warp_view_fwd(kind, D, M, V, v, S, n, Xtilde[v], delta_G[v], log_ls[v:v + 1], log_var[v:v + 1],
                                        Omega_G, hld_G, X, eps, Lk_all[v], None, Kinv, hldK[v:v + 1], info[v:v + 1], A, B, T,
                                        Ke, var, Gmean, Gs, n * D, kl if meta["with_kl"] else None, ws64, Kuu_e, Kuf_e)
            # `var` is written on the side stream and used nowhere else: it has to outlive the join below, or its block
            # returns to the current stream's allocator pool while this view's chain is still running
            keep.append(var)
            saved += [X, eps, Kinv, A, B, T, Ke]
            outs += [Gmean, Gs]
        for s_ in side:
            cur.wait_stream(s_)
        del keep
        ctx.meta = meta
        ctx.set_materialize_grads(False)  # unused outputs (G_means, the cached factors) arrive as None, not as zero tensors
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
        ext = kind == _lib.KIND_EXTERNAL
        for k, v in enumerate(free):
            X, eps, Kinv, A, B, T, Ke = saved[7 * k: 7 * k + 7]
            n = X.shape[0]
            gm, gs = gouts[2 * k], gouts[2 * k + 1]
            Kuu_b = _new(dev, M, M) if ext else None
            Kuf_b = _new(dev, M, n) if ext else None
            xgrads += [None, None, Kuu_b, Kuf_b] if ext else [None, None]
            if n == 0:
                if ext:
                    Kuu_b.zero_()
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
            ctx_stream = contextlib.nullcontext()
            if side:
                # fork HERE, after this view's zero fills / contiguous copies were enqueued on `cur`
                side[lane].wait_stream(cur)
                ctx_stream = torch.cuda.stream(side[lane])
            with ctx_stream:
                ops().warp_view_bwd(kind, D, M, V, v, S, n, Xtilde[v], delta_G[v], log_ls[v:v + 1], log_var[v:v + 1], Omega_G,
                                    X, eps, Kinv, A, B, T, Ke, gs, n * D, gm, klb, acc_Z[v], acc_dlt[v], acc_hyp[v], Obar,
                                    mubar, varbar, q1bar, Abar, Cm, AS, ws64, Kuu_b, Kuf_b)
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
class DataLayerPre(torch.autograd.Function):
    """One modality of the data GP up to its predictive moments (reference gpsa/models/vgpsa.py:390-421, KL :520-530).

    forward(meta, Gtilde, log_ls, log_var, delta_F, Omega_sqt_F, G [S,N,D][, Kuu [M,M], Kuf [M,S*N][, Omega, hld]])
      -> (mean [S,N,L], q2 [S,N,L], kq [S,N], KL_F, Kuu_chol_F, Omega_tril_F, info)
    with q2[s,n,p] = a^T Omega_p a (the hot contraction) and kq = sigma^2 - a^T K a; the marginal variance is
    kq + q2 + 2e-5 and the sample F = mean + sqrt(var) eps is the sampling stage (SampleF / SampleNLL below).
    Kuu, Kuf: only with kind = KIND_EXTERNAL (user-supplied covariance callable): k(Gt,Gt) and k(Gt,G) are evaluated by
    the caller with torch and the backward returns dLoss/dK for both (and nothing for Gtilde / G); None otherwise.
    Omega [L,M,M], hld [L]: the differentiable outputs of OmegaChain, whose other outputs come in
    meta["omega_chain"] = (Ltril, info, chain_meta).  The variational covariances are then a node of their own: the
    backward returns dLoss/dOmega and dLoss/dhld for them and nothing for Omega_sqt_F.  Without them the layer prepares
    Omega itself (or takes meta["omega"]) and differentiates it inline.
    meta = dict(kind=int, with_kl=bool[, omega=omega_prepare(Omega_sqt_F) | omega_chain=(...)][, prior=prior_prepare(...)])
    """

    @staticmethod
    @on_device_of(1)
    def forward(ctx, meta, Gtilde, log_ls, log_var, delta_F, Osq_F, G, Kuu=None, Kuf=None, Omega_in=None, hld_in=None):
        Gtilde, delta_F, Osq_F = _c(Gtilde.detach()), _c(delta_F.detach()), _c(Osq_F.detach())
        log_ls, log_var = _c(log_ls.detach().reshape(1)), _c(log_var.detach().reshape(1))
        G = _c(G.detach())
        M, D = Gtilde.shape
        L = delta_F.shape[1]
        S, N = G.shape[0], G.shape[1]
        R = S * N
        kind = meta["kind"]
        ext = kind == _lib.KIND_EXTERNAL
        if ext and (Kuu is None or Kuf is None or tuple(Kuu.shape) != (M, M) or tuple(Kuf.shape) != (M, R)):
            raise ValueError("external covariance: expected Kuu [M,M] and Kuf [M,S*N]")
        chained = Omega_in is not None
        if chained:
            Ltril, info_O, chain_meta = meta["omega_chain"]
            L64 = None
            Omega, hld = _c(Omega_in.detach()), _c(hld_in.detach())
        else:
            pre = meta.get("omega")
            Omega, Ltril, L64, hld, info_O = pre if pre is not None else _omega_prepare(Osq_F)
        prior = meta.get("prior")
        if prior is not None:
            Lk, Kinv, Kinv64, hldK, info = prior
        else:
            Lk, Kinv, Kinv64 = _new(G, M, M), _new(G, M, M), _new(G, M, M, dtype=f64)
            hldK = _zeros(G, 1, dtype=f64)
            info = _zeros(G, 1, dtype=i32)
        A, kq = _new(G, M, R), _new(G, S, N)
        B = _c(Kuf.detach().to(f32)) if ext else _new(G, M, R)  # external: B comes in filled
        Kuu_e = _c(Kuu.detach().to(f32)) if ext else None
        engine = pick_engine(M, R, L)
        if chained:
            chain_meta["tc"] = engine in TC_ENGINES
        W = _new(G, _lib.feat_count(M), L) if engine == 0 else _new(G, 1)
        tc_ws = _lib.tc_workspace(M, R, L, G) if engine in TC_ENGINES else None
        KD = _new(G, M, L, dtype=f64)
        mean, q2 = _new(G, S, N, L), _new(G, S, N, L)
        kl = _zeros(G, 1, dtype=f64)
        ws64 = _new(G, 2 * M * M, dtype=f64)
        ops().data_layer_fwd(kind, D, M, L, R, Gtilde, log_ls, log_var, delta_F, Omega, hld, G, Lk, Kinv, Kinv64, hldK, info,
                             A, B, kq, W, KD, mean, q2, kl if meta["with_kl"] else None, ws64, engine, tc_ws, Kuu_e,
                             1 if prior is not None else 0)
        ctx.meta = meta
        ctx.engine = engine
        ctx.dims = (S, N)
        ctx.set_materialize_grads(False)  # no [L,M,M] zero fill for the gradient of the cached factor
        ctx.n_in = 11 if chained else (9 if Kuu is not None or Kuf is not None else 7)
        ctx.chained = chained
        ctx.save_for_backward(Gtilde, log_ls, log_var, delta_F, Osq_F, G, Omega, L64 if not chained else None, Kinv, Kinv64, A, B, W,
                              KD)
        info_all = torch.cat([info_O, info])
        ctx.mark_non_differentiable(Lk, Ltril, info_all)
        return mean, q2, kq, kl.to(f32).reshape(()), Lk, Ltril, info_all

    @staticmethod
    @on_device_of(1)
    def backward(ctx, mean_bar, q2_bar, kq_bar, kl_bar, _1, _2, _3):
        meta = ctx.meta
        Gtilde, log_ls, log_var, delta_F, Osq_F, G, Omega, L64, Kinv, Kinv64, A, B, W, KD = ctx.saved_tensors
        M, D = Gtilde.shape
        L = delta_F.shape[1]
        S, N = ctx.dims
        R = S * N
        dev = G
        use_kl = meta["with_kl"] and kl_bar is not None
        klb = _c(kl_bar.to(f32).reshape(1)) if use_kl else None
        mean_bar = _c(mean_bar) if mean_bar is not None else _zeros(dev, S, N, L)
        q2_bar = _c(q2_bar) if q2_bar is not None else _zeros(dev, S, N, L)
        kq_bar = _c(kq_bar) if kq_bar is not None else _zeros(dev, S, N)
        ext = meta["kind"] == _lib.KIND_EXTERNAL
        G_bar = _new(dev, S, N, D) if not ext else None
        Kuu_b = _new(dev, M, M) if ext else None
        acc_Gt, acc_hyp = _zeros(dev, M, D, dtype=f64), _zeros(dev, 2, dtype=f64)
        dlt_bar, Obar = _new(dev, M, L), _new(dev, L, M, M)
        q1bar = _new(dev, R)
        Abar, Cm = _new(dev, M, R), _new(dev, M, R)
        engine = ctx.engine
        H = _new(dev, _lib.feat_count(M), L)
        tc_ws = _lib.tc_workspace(M, R, L, dev) if engine in TC_ENGINES else None
        ws64 = _new(dev, 3 * M * M, dtype=f64)
        ops().data_layer_bwd(meta["kind"], D, M, L, R, Gtilde, log_ls, log_var, delta_F, Omega, G, Kinv, Kinv64, A, B, W, KD,
                             mean_bar, q2_bar, kq_bar, klb, G_bar, acc_Gt, acc_hyp, dlt_bar, Obar, q1bar, Abar, Cm, H, ws64,
                             engine, tc_ws, Kuu_b)
        del tc_ws
        hyp = acc_hyp.to(f32)
        if ctx.chained:  # OmegaChain turns (dLoss/dOmega, dLoss/dhld) into Osq_bar, on its own stream
            Osq_bar = None
            tail = (Obar, (-klb.to(f64)).expand(L) if use_kl else None)  # KL_F holds -hld_p for every gene
        else:
            coef = _c((-0.5 * klb).expand(L)) if use_kl else None
            Osq_bar = omega_grad(Osq_F, L64, Obar, coef, tc=(engine in TC_ENGINES))
            tail = ()
        if ext:  # C = dLoss/dK_uf; the covariance function's own arguments get their gradients through the caller's autograd
            return (None, None, None, hyp[1:2], dlt_bar, Osq_bar, None, Kuu_b, Cm) + tail
        return (None, acc_Gt.to(f32), hyp[0:1], hyp[1:2], dlt_bar, Osq_bar, G_bar) + ((None, None) if ctx.n_in >= 9 else ()) + tail


class SampleF(torch.autograd.Function):
    """Materialised sampling stage: F = mean + sqrt(kq + q2 + 2e-5) eps  (reference gpsa/models/vgpsa.py:197-204, :423-426).

    The kernel works IN PLACE on the buffers of `mean` and `q2` (outputs of DataLayerPre that nothing else reads): the
    returned F aliases `mean`, and the marginal variance saved for the backward aliases `q2`."""

    @staticmethod
    @on_device_of(1)
    def forward(ctx, mean, q2, kq, eps):
        S, N, L = mean.shape
        F, var = mean.detach(), q2.detach()
        if not (F.is_contiguous() and var.is_contiguous()):
            raise _lib.GPSALibraryError("SampleF expects the contiguous outputs of DataLayerPre")
        kq, eps = _c(kq.detach()), _c(eps.detach())
        ops().sample_fwd(S * N, L, kq, eps, F, var)
        ctx.save_for_backward(eps, var)
        return F

    @staticmethod
    @on_device_of(1)
    def backward(ctx, F_bar):
        eps, var = ctx.saved_tensors
        S, N, L = var.shape
        F_bar = _c(F_bar)
        q2_bar, kq_bar = _new(var, S, N, L), _new(var, S, N)
        ops().sample_bwd(S * N, L, F_bar, eps, var, q2_bar, kq_bar)
        return F_bar, q2_bar, kq_bar, None


def philox_normal(key, S, N, L, gene_off=0, samp_off=0):
    """eps [S,N,L] of the counter-based generator the fused sampling stage draws from in-kernel (same numbers)."""
    with torch.cuda.device(key.device):
        out = torch.empty(S, N, L, dtype=f32, device=key.device)
        ops().philox_normal(N, S, L, key, int(gene_off), int(samp_off), out)
    return out


class SampleNLL(torch.autograd.Function):
    """Fused sampling stage + Gaussian negative log-likelihood (reference gpsa/models/vgpsa.py:423-426, :532-538):
        -sum log N(Y; mean + sqrt(kq + q2 + 2e-5) eps, sigma) / S
    One pass over the two [S,N,L] buffers; eps comes from `eps` [S,N,L] if given, else from Philox keyed by
    (key, sample, spot, global gene).  F, eps and var are never stored: the buffers of `mean` and `q2` are overwritten
    IN PLACE with d(-LL)/dF and d(-LL)/dvar, which is all the backward needs.

    forward(meta, mean, q2, kq, Y [N,L], log_noise, eps | None, key | None) -> -LL (0-dim)
    meta = dict(gene_off=int, samp_off=int)
    """

    @staticmethod
    @on_device_of(1)
    def forward(ctx, meta, mean, q2, kq, Y, log_noise, eps, key):
        S, N, L = mean.shape
        U, Gu = mean.detach(), q2.detach()
        if not (U.is_contiguous() and Gu.is_contiguous()):
            raise _lib.GPSALibraryError("SampleNLL expects the contiguous outputs of DataLayerPre")
        Y, log_noise, kq = _c(Y.detach()), _c(log_noise.detach().reshape(1)), _c(kq.detach())
        if Y.shape != (N, L):
            raise ValueError(f"outputs have shape {tuple(Y.shape)}, F_samples imply {(N, L)}")
        eps = _c(eps.detach()) if eps is not None else None
        kqb = _new(U, S, N)
        acc = _zeros(U, 2, dtype=f64)  # [-LL, d(-LL)/dlog_noise]
        ops().sample_ll_fused(N, S, L, kq, Y, log_noise, eps, key, int(meta.get("gene_off", 0)),
                              int(meta.get("samp_off", 0)), U, Gu, kqb, acc[0:1], acc[1:2])
        ctx.save_for_backward(U, Gu, kqb, acc)
        ctx.used = False
        return acc[0].to(f32)

    @staticmethod
    @on_device_of(1)
    def backward(ctx, g):
        if ctx.used:
            raise RuntimeError("the fused sampling + likelihood stage overwrites its buffers in place: backward through "
                               "it a second time needs a new forward (set model.fused_ll = False for retain_graph use)")
        ctx.used = True
        U, Gu, kqb, acc = ctx.saved_tensors
        g = _c(g.to(f32).reshape(1))
        # loss.backward() hands down exactly 1: the kernels then return after one load
        for t in (U, Gu, kqb):
            ops().scale_if_not_one(t.numel(), g, t)
        return None, U, Gu, kqb, None, (acc[1] * g).to(f32).reshape(1), None, None


class LMCObserve(torch.autograd.Function):
    """Linear model of coregionalisation, materialised: F_obs [S,N,P] = F_lat [S,N,L] @ W [L,P]
    (reference gpsa/models/vgpsa.py:428-432), explicit forward and backward on the library's GEMM."""

    @staticmethod
    @on_device_of(1)
    def forward(ctx, F_lat, W):
        F_lat, W = _c(F_lat.detach()), _c(W.detach())
        S, N, L = F_lat.shape
        P = W.shape[1]
        if W.shape[0] != L:
            raise ValueError(f"W has shape {tuple(W.shape)}, latent samples have {L} outputs")
        out = _new(F_lat, S, N, P)
        ops().lmc_fwd(S * N, L, P, F_lat, W, out)
        ctx.save_for_backward(F_lat, W)
        return out

    @staticmethod
    @on_device_of(1)
    def backward(ctx, Fb):
        F_lat, W = ctx.saved_tensors
        S, N, L = F_lat.shape
        P = W.shape[1]
        Fb = _c(Fb)
        Flb, Wb = _new(F_lat, S, N, L), _new(W, L, P)
        ops().lmc_bwd(S * N, L, P, F_lat, W, Fb, Flb, Wb)
        return Flb, Wb


def lmc_fused_supported(L):
    return int(L) <= int(lib().gpsa_lmc_max_latent())


class LMCNLL(torch.autograd.Function):
    """LMC fused with the Gaussian negative log-likelihood: -sum log N(Y; F_lat @ W, sigma) / S without forming the
    [S,N,P] tensor (reference gpsa/models/vgpsa.py:428-432, :532-538).  forward(F_lat [S,N,L], W [L,P], Y [N,P], log_noise)."""

    @staticmethod
    @on_device_of(1)
    def forward(ctx, F_lat, W, Y, log_noise):
        F_lat, W, Y = _c(F_lat.detach()), _c(W.detach()), _c(Y.detach())
        log_noise = _c(log_noise.detach().reshape(1))
        S, N, L = F_lat.shape
        P = W.shape[1]
        if Y.shape != (N, P) or W.shape[0] != L:
            raise ValueError(f"outputs {tuple(Y.shape)} / loadings {tuple(W.shape)} do not match samples {(S, N, L)}")
        Flb, Wb = _new(F_lat, S, N, L), _new(W, L, P)
        acc = _zeros(F_lat, 2, dtype=f64)
        ops().lmc_ll_fused(N, S, L, P, F_lat, W, Y, log_noise, Flb, Wb, acc[0:1], acc[1:2])
        ctx.save_for_backward(Flb, Wb, acc)
        ctx.used = False
        return acc[0].to(f32)

    @staticmethod
    @on_device_of(1)
    def backward(ctx, g):
        if ctx.used:
            raise RuntimeError("the fused LMC likelihood keeps its gradients in place: a second backward needs a new forward")
        ctx.used = True
        Flb, Wb, acc = ctx.saved_tensors
        g = _c(g.to(f32).reshape(1))
        for t in (Flb, Wb):
            ops().scale_if_not_one(t.numel(), g, t)
        return Flb, Wb, None, (acc[1] * g).to(f32).reshape(1)


class DataLayer:
    """DataLayerPre followed by the materialised sampling stage: the reference's data layer as one callable.

    apply(meta, Gtilde, log_ls, log_var, delta_F, Omega_sqt_F, G [S,N,D], eps [S,N,L])
      -> (F_latent [S,N,L], KL_F, Kuu_chol_F, Omega_tril_F, info)
    """

    @staticmethod
    def apply(meta, Gtilde, log_ls, log_var, delta_F, Osq_F, G, eps, *kext):
        mean, q2, kq, kl, Lk, Ltril, info = DataLayerPre.apply(meta, Gtilde, log_ls, log_var, delta_F, Osq_F, G, *kext)
        return SampleF.apply(mean, q2, kq, eps), kl, Lk, Ltril, info


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
        ops().gaussian_ll_fwd(N, P, S, F, Y, log_noise, acc)
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
        ops().gaussian_ll_bwd(N, P, S, F, Y, log_noise, llb, F_bar, acc)
        return F_bar, None, acc.to(f32)
