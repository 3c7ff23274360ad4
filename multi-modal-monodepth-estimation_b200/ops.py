"""torch.autograd.Function wrappers around the C-ABI kernels.

Each Function keeps its tensors alive for the duration of the asynchronous launch (PyTorch's
caching allocator is stream-ordered on the current stream, which is the stream passed to the
kernels) and calls straight into libb200swin.so; nothing here computes on the host.
"""
from __future__ import annotations

import torch

from . import _lib as L


# ------------------------------------------------------------------------------ SiLog
class _SiLog(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, lambd):
        L.require_cuda(pred, target)
        lib = L.load()
        pred_c = pred.contiguous()
        tgt_c = target.contiguous()
        if tgt_c.dtype != torch.float32:
            tgt_c = tgt_c.float()
        if pred_c.shape != tgt_c.shape:
            raise ValueError(f"SiLog: pred {tuple(pred.shape)} and target {tuple(target.shape)} differ")
        n = pred_c.numel()
        with torch.cuda.device_of(pred_c):
            out = torch.empty(5, dtype=torch.float32, device=pred.device)      # [loss, stats(4)]
            ws_bytes = lib.b200swin_silog_workspace_bytes(n)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=pred.device)
            L.check(lib.b200swin_silog_fwd(pred_c.data_ptr(), L.dtype_code(pred_c), tgt_c.data_ptr(), n, lambd,
                                           out.data_ptr(), out.data_ptr() + 4, ws.data_ptr(), ws_bytes,
                                           L.stream_of(pred_c)), "silog_fwd")
        ctx.save_for_backward(pred_c, tgt_c, out)
        ctx.lambd = lambd
        ctx.pred_shape = pred.shape
        return out[0]

    @staticmethod
    def backward(ctx, gout):
        pred_c, tgt_c, out = ctx.saved_tensors
        lib = L.load()
        g = gout.contiguous().float()
        grad = torch.empty_like(pred_c)
        with torch.cuda.device_of(pred_c):
            L.check(lib.b200swin_silog_bwd(pred_c.data_ptr(), L.dtype_code(pred_c), tgt_c.data_ptr(), pred_c.numel(),
                                           ctx.lambd, out.data_ptr() + 4, g.data_ptr(), grad.data_ptr(),
                                           L.stream_of(pred_c)), "silog_bwd")
        return grad.view(ctx.pred_shape), None, None


def silog_loss(pred: torch.Tensor, target: torch.Tensor, lambd: float = 0.5) -> torch.Tensor:
    return _SiLog.apply(pred, target, float(lambd))


# ------------------------------------------------------------------------------ windows
def _window_move(x, B, H, W, C, ws, shift, gather: bool):
    lib = L.load()
    Hp, Wp = (H + ws - 1) // ws * ws, (W + ws - 1) // ws * ws
    nW = (Hp // ws) * (Wp // ws)
    x = x.contiguous()
    with torch.cuda.device_of(x):
        if gather:
            out = torch.empty((B * nW, ws * ws, C), dtype=x.dtype, device=x.device)
            fn, what = lib.b200swin_window_gather, "window_gather"
        else:
            out = torch.empty((B, H, W, C), dtype=x.dtype, device=x.device)
            fn, what = lib.b200swin_window_scatter, "window_scatter"
        L.check(fn(x.data_ptr(), out.data_ptr(), B, H, W, C, ws, shift, x.element_size(), L.stream_of(x)), what)
    return out


class _WindowGather(torch.autograd.Function):
    """x[B,H,W,C] -> [B*nW, ws*ws, C]: pad + roll(-shift) + partition in one pass."""

    @staticmethod
    def forward(ctx, x, ws, shift):
        L.require_cuda(x)
        B, H, W, C = x.shape
        ctx.geom = (B, H, W, C, ws, shift)
        return _window_move(x, B, H, W, C, ws, shift, True)

    @staticmethod
    def backward(ctx, g):
        B, H, W, C, ws, shift = ctx.geom
        return _window_move(g, B, H, W, C, ws, shift, False), None, None


class _WindowScatter(torch.autograd.Function):
    """win[B*nW, ws*ws, C] -> [B,H,W,C]: reverse + roll(+shift) + crop in one pass."""

    @staticmethod
    def forward(ctx, win, B, H, W, ws, shift):
        L.require_cuda(win)
        C = win.shape[-1]
        ctx.geom = (B, H, W, C, ws, shift)
        return _window_move(win, B, H, W, C, ws, shift, False)

    @staticmethod
    def backward(ctx, g):
        B, H, W, C, ws, shift = ctx.geom
        return _window_move(g, B, H, W, C, ws, shift, True), None, None, None, None, None


def window_gather(x, ws, shift=0):
    return _WindowGather.apply(x, int(ws), int(shift))


def window_scatter(win, B, H, W, ws, shift=0):
    return _WindowScatter.apply(win, int(B), int(H), int(W), int(ws), int(shift))


class _PatchMerge(torch.autograd.Function):
    """x[B,H,W,C] -> [B, ceil(H/2)*ceil(W/2), 4C]: zero-pad to even + the four strided slices + cat of
    PatchMerging.forward (swin_transformer_v2.py:660-672) as ONE gather; backward is the adjoint scatter."""

    @staticmethod
    def forward(ctx, x):
        L.require_cuda(x)
        lib = L.load()
        B, H, W, C = x.shape
        xc = x.contiguous()
        H2, W2 = (H + 1) // 2, (W + 1) // 2
        with torch.cuda.device_of(xc):
            out = torch.empty((B, H2 * W2, 4 * C), dtype=x.dtype, device=x.device)
            L.check(lib.b200swin_patch_merge(xc.data_ptr(), out.data_ptr(), B, H, W, C, xc.element_size(), 0,
                                             L.stream_of(xc)), "patch_merge")
        ctx.geom = (B, H, W, C)
        return out

    @staticmethod
    def backward(ctx, g):
        B, H, W, C = ctx.geom
        lib = L.load()
        gc = g.contiguous()
        with torch.cuda.device_of(gc):
            dx = torch.empty((B, H, W, C), dtype=g.dtype, device=g.device)
            L.check(lib.b200swin_patch_merge(gc.data_ptr(), dx.data_ptr(), B, H, W, C, gc.element_size(), 1,
                                             L.stream_of(gc)), "patch_merge(bwd)")
        return dx


def patch_merge(x):
    return _PatchMerge.apply(x)


def patchify(x, ph, pw, out_dtype):
    """x[B,Cin,H,W] (no gradient) -> cols[B*ceil(H/ph)*ceil(W/pw), Cin*ph*pw] in `out_dtype` (see b200swin_patchify)."""
    L.require_cuda(x)
    if x.requires_grad and torch.is_grad_enabled():
        raise RuntimeError("patchify: the input image must not require a gradient")
    lib = L.load()
    xc = x.contiguous()
    if xc.dtype not in (torch.float32, torch.bfloat16):
        xc = xc.float()
    B, Cin, H, W = xc.shape
    Hp, Wp = (H + ph - 1) // ph, (W + pw - 1) // pw
    with torch.cuda.device_of(xc):
        cols = torch.empty((B * Hp * Wp, Cin * ph * pw), dtype=out_dtype, device=x.device)
        L.check(lib.b200swin_patchify(xc.data_ptr(), L.dtype_code(xc), cols.data_ptr(), L.dtype_code(cols), B, Cin, H, W,
                                      ph, pw, L.stream_of(xc)), "patchify")
    return cols, Hp, Wp


def shift_mask(H, W, ws, shift, device):
    lib = L.load()
    Hp, Wp = (H + ws - 1) // ws * ws, (W + ws - 1) // ws * ws
    nW, N = (Hp // ws) * (Wp // ws), ws * ws
    out = torch.empty((nW, N, N), dtype=torch.float32, device=device)
    L.require_cuda(out)
    with torch.cuda.device_of(out):
        L.check(lib.b200swin_shift_mask(out.data_ptr(), H, W, ws, shift, L.stream_of(out)), "shift_mask")
    return out


# ------------------------------------------------------------------------------ LayerNorm (+residual)
def ln_colsum_supported(dtype: torch.dtype, C: int) -> bool:
    """Whether the LayerNorm backward can also emit the column sums of dx (bf16 streaming kernels only)."""
    return dtype == torch.bfloat16 and C % 8 == 0 and C <= 1536


class _LayerNormResidual(torch.autograd.Function):
    """y = residual + row_scale[b] * (LN(x) * gamma + beta); residual / row_scale optional.
    `producer_bias` is the bias of the Linear that produced x: it is NOT used in the forward (the GEMM epilogue
    already added it) -- it is an input only so that its gradient, the column sums of dx, can be returned from the
    LayerNorm backward kernel, which has dx in registers anyway (the caller passes bias.detach() to the Linear).

    stream32: bf16 activations with an fp32 residual stream (torch.autocast semantics: LayerNorm outputs and residual adds
    stay fp32).  `residual32` is the fp32 twin of `residual` (values only; gradients keep flowing through the bf16
    `residual`); the call returns (y, y32) with y = bf16(y32) and y32 marked non-differentiable."""

    @staticmethod
    def forward(ctx, x, residual, gamma, beta, row_scale, rows_per_scale, eps, producer_bias=None, residual32=None,
                stream32=False):
        L.require_cuda(x, residual, gamma, beta, row_scale, residual32)
        lib = L.load()
        C = x.shape[-1]
        xc = x.contiguous()
        rows = xc.numel() // C
        rc = None
        if residual is not None:
            rc = residual.contiguous()
            if rc.dtype != xc.dtype:
                rc = rc.to(xc.dtype)
        g32 = gamma.contiguous().float()
        b32 = beta.contiguous().float()
        rs = None if row_scale is None else row_scale.contiguous().float()
        y32 = None
        with torch.cuda.device_of(xc):
            y = torch.empty_like(xc)
            stats = torch.empty((2, rows), dtype=torch.float32, device=x.device)
            if stream32:
                if not stream32_supported(xc.dtype, C):
                    raise RuntimeError("layer_norm_residual: the fp32 residual stream needs bf16 activations, C % 8 == 0, C <= 1024")
                r32 = None
                if residual is not None:
                    r32 = (residual32 if residual32 is not None else rc.float()).detach().contiguous()
                    if r32.dtype != torch.float32 or r32.shape != xc.shape:
                        raise ValueError("layer_norm_residual: residual32 must be an fp32 tensor of x's shape")
                y32 = torch.empty(xc.shape, dtype=torch.float32, device=x.device)
                L.check(lib.b200swin_ln_fwd_stream32(xc.data_ptr(), L.ptr(r32), g32.data_ptr(), b32.data_ptr(), L.ptr(rs),
                                                     rows_per_scale, y.data_ptr(), y32.data_ptr(), stats[0].data_ptr(),
                                                     stats[1].data_ptr(), rows, C, eps, L.stream_of(xc)), "ln_fwd_stream32")
            else:
                L.check(lib.b200swin_ln_fwd(xc.data_ptr(), L.ptr(rc), g32.data_ptr(), b32.data_ptr(), L.ptr(rs),
                                            rows_per_scale, y.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(),
                                            rows, C, eps, L.dtype_code(xc), L.stream_of(xc)), "ln_fwd")
        ctx.save_for_backward(xc, g32, stats, rs)
        ctx.has_res = residual is not None
        ctx.rows_per_scale = rows_per_scale
        ctx.gdtype, ctx.bdtype = gamma.dtype, beta.dtype
        ctx.pbdtype = None
        ctx.stream32 = bool(stream32)
        if producer_bias is not None:
            if not ln_colsum_supported(xc.dtype, C):
                raise RuntimeError("layer_norm_residual: producer_bias needs bf16 activations with C % 8 == 0, C <= 1536")
            ctx.pbdtype = producer_bias.dtype
        if stream32:
            y32 = y32.view(x.shape)
            ctx.mark_non_differentiable(y32)
            # the twin never has a gradient: without this autograd hands the backward a zero-filled [T, C] fp32 tensor
            # for it (an 88 MB fill per LayerNorm at Swin-B stage 2)
            ctx.set_materialize_grads(False)
            return y.view(x.shape), y32
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy, _dy32=None):
        xc, g32, stats, rs = ctx.saved_tensors
        if dy is None:                      # only reachable with set_materialize_grads(False): y itself was unused
            dy = torch.zeros_like(xc)
        lib = L.load()
        C = xc.shape[-1]
        rows = xc.numel() // C
        dyc = dy.contiguous()
        if dyc.dtype != xc.dtype:
            dyc = dyc.to(xc.dtype)
        with torch.cuda.device_of(xc):
            dx = torch.empty_like(xc)
            dgb = torch.empty((3, C), dtype=torch.float32, device=xc.device)
            ws_bytes = lib.b200swin_ln_bwd_workspace_bytes(rows, C)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xc.device)
            want_cs = ctx.pbdtype is not None and ctx.needs_input_grad[7]
            L.check(lib.b200swin_ln_bwd(dyc.data_ptr(), xc.data_ptr(), g32.data_ptr(), stats[0].data_ptr(),
                                        stats[1].data_ptr(), L.ptr(rs), ctx.rows_per_scale, dx.data_ptr(),
                                        dgb[0].data_ptr(), dgb[1].data_ptr(), dgb[2].data_ptr() if want_cs else 0,
                                        rows, C, L.dtype_code(xc), ws.data_ptr(), ws_bytes, L.stream_of(xc)), "ln_bwd")
        dres = dyc.view(dy.shape) if ctx.has_res else None
        dpb = dgb[2].to(ctx.pbdtype) if want_cs else None
        return dx, dres, dgb[0].to(ctx.gdtype), dgb[1].to(ctx.bdtype), None, None, None, dpb, None, None


def stream32_supported(dtype: torch.dtype, C: int) -> bool:
    """Whether layer_norm_residual can keep an fp32 residual stream beside the bf16 activations."""
    return dtype == torch.bfloat16 and C % 8 == 0 and C <= 1024


def layer_norm_residual(x, gamma, beta, eps, residual=None, row_scale=None, rows_per_scale=1, producer_bias=None,
                        residual32=None, stream32=False):
    return _LayerNormResidual.apply(x, residual, gamma, beta, row_scale, int(rows_per_scale), float(eps), producer_bias,
                                    residual32, bool(stream32))


# The fp32 twin of a bf16 residual-stream tensor travels as an attribute of the tensor OBJECT the LayerNorm returned: the
# reference's block signature (x, mask_matrix) has no room for a second tensor.  Lost attributes (a caller that copies or
# views the tensor) only cost precision: the next LayerNorm then widens the bf16 tensor instead.
_TWIN = "_b200swin_fp32_stream"


def attach_stream(y16: torch.Tensor, y32: torch.Tensor) -> torch.Tensor:
    setattr(y16, _TWIN, y32)
    return y16


def stream_of_tensor(x: torch.Tensor):
    t = getattr(x, _TWIN, None)
    return t if (t is not None and t.shape == x.shape and t.device == x.device) else None


# ------------------------------------------------------------------------------ dense contractions
def compute_dtype(x: torch.Tensor) -> torch.dtype:
    """bf16 under torch.autocast('cuda', torch.bfloat16) (our definition of the reference's missing
    bf16 mode) or for bf16 inputs; otherwise fp32 (reference precision)."""
    if torch.is_autocast_enabled('cuda'):
        dt = torch.get_autocast_dtype('cuda')
        if dt != torch.bfloat16:
            raise TypeError("b200swin supports bfloat16 autocast only")
        return dt
    if x.dtype in (torch.float32, torch.bfloat16):
        return x.dtype
    raise TypeError(f"b200swin: unsupported activation dtype {x.dtype}")


class Operand:
    """A GEMM operand staged for the tensor cores: bf16 `hi` and, in fp32-accurate mode, the bf16
    residual `lo` (x ~ hi + lo)."""
    __slots__ = ("hi", "lo")

    def __init__(self, hi, lo=None):
        self.hi, self.lo = hi, lo


def stage_operand(t: torch.Tensor, exact: bool) -> Operand:
    """bf16 tensors pass through; fp32 tensors are cast (exact=False) or split hi/lo (exact=True)."""
    t = t.contiguous()
    if t.dtype == torch.bfloat16:
        if exact:
            raise TypeError("fp32-accurate GEMM mode needs float32 operands")
        return Operand(t)
    if t.dtype != torch.float32:
        raise TypeError(f"unsupported operand dtype {t.dtype}")
    lib = L.load()
    with torch.cuda.device_of(t):
        hi = torch.empty(t.shape, dtype=torch.bfloat16, device=t.device)
        lo = torch.empty_like(hi) if exact else None
        L.check(lib.b200swin_split_bf16(t.data_ptr(), hi.data_ptr(), L.ptr(lo), t.numel(), L.stream_of(t)),
                "split_bf16")
    return Operand(hi, lo)


_weight_cache: dict = {}
_staged_registry: dict = {}      # id(param) -> [weakref(param), bf16 view kept current by FusedAdamW, version seen]


def invalidate_weight_cache() -> None:
    """Forget every staged bf16 copy of a weight.  Call after changing parameters through ``p.data`` (EMA, clipping:
    such writes do not bump ``p._version``, which the cache keys on)."""
    _weight_cache.clear()
    for reg in _staged_registry.values():
        reg[2] = -1


def register_staged_weight(p: torch.Tensor, copy_bf16: torch.Tensor) -> None:
    """``copy_bf16`` is a bf16 tensor of p's shape that the caller keeps equal to ``p`` (optim.FusedAdamW writes it in
    the optimizer kernel); GEMMs read it instead of casting ``p`` every step."""
    import weakref
    key = id(p)
    _staged_registry[key] = [weakref.ref(p, lambda _r, k=key: _staged_registry.pop(k, None)), copy_bf16, p._version]


def stage_weight(w: torch.Tensor, exact: bool) -> Operand:
    """Stage a parameter once per value.  Parameters owned by ``optim.FusedAdamW`` come with a bf16 copy that the
    optimizer kernel refreshes; anything else is cast here and cached on (storage, version counter) -- except under
    CUDA-graph capture, where the cast must be part of every capture and nothing is cached."""
    if not exact:
        reg = _staged_registry.get(id(w))
        if reg is not None and reg[0]() is w:
            if reg[2] != w._version:                       # changed outside the optimizer (init, load_state_dict)
                wc = w.detach().contiguous()
                with torch.cuda.device_of(wc):
                    L.check(L.load().b200swin_split_bf16(wc.data_ptr(), reg[1].data_ptr(), 0, wc.numel(),
                                                         L.stream_of(wc)), "split_bf16")
                reg[2] = w._version
            return Operand(reg[1])
    if torch.cuda.is_current_stream_capturing():
        return stage_operand(w.detach(), exact)
    key = (w.data_ptr(), exact)      # data_ptr is stable for a live parameter; version catches in-place updates
    hit = _weight_cache.get(key)
    if hit is not None and hit[0] == w._version and hit[1] == tuple(w.shape) and hit[3]() is w:
        return hit[2]
    import weakref
    op = stage_operand(w.detach(), exact)
    if len(_weight_cache) > 1024:
        _weight_cache.clear()
    _weight_cache[key] = (w._version, tuple(w.shape), op, weakref.ref(w, lambda _r, k=key: _weight_cache.pop(k, None)))
    return op


def gemm(a: Operand, b: Operand, M: int, N: int, K: int, *, a_mn=False, b_mn=False, epilogue=L.EPI_NONE, bias=None,
         bias2=None, aux_in=None, aux_out=None, inv_norm=None, nH=0, out_dtype=torch.bfloat16, splits=1):
    """out[M,N] = epilogue(A . B^T) through b200swin_gemm_bf16 (see include/b200swin.h)."""
    lib = L.load()
    dev = a.hi.device
    with torch.cuda.device(dev):
        out = torch.empty((M, N), dtype=out_dtype, device=dev)
        ws, ws_bytes = None, 0
        if splits > 1:
            ws_bytes = lib.b200swin_gemm_workspace_bytes(M, N, splits)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        L.check(lib.b200swin_gemm_bf16(a.hi.data_ptr(), L.ptr(a.lo), int(a_mn), b.hi.data_ptr(), L.ptr(b.lo), int(b_mn),
                                       M, N, K, epilogue, L.ptr(bias), L.ptr(bias2), L.ptr(aux_in), L.ptr(aux_out),
                                       L.ptr(inv_norm), nH, out.data_ptr(), L.dtype_code(out), splits, L.ptr(ws),
                                       ws_bytes, L.stream_of(out)), "gemm_bf16")
    return out


def colsum(x: torch.Tensor, col0: int = 0, ncols: int | None = None, extra: torch.Tensor | None = None):
    """fp32 column sums of a 2-D [M, ld] tensor over columns [col0, col0+ncols)."""
    lib = L.load()
    M, ld = x.shape
    ncols = ld - col0 if ncols is None else ncols
    with torch.cuda.device_of(x):
        out = torch.empty(ncols, dtype=torch.float32, device=x.device)
        ws_bytes = lib.b200swin_colsum_workspace_bytes(M, ncols)
        ws = torch.empty(max(ws_bytes, 4), dtype=torch.uint8, device=x.device)
        L.check(lib.b200swin_colsum(x.data_ptr(), L.dtype_code(x), M, ld, col0, ncols, L.ptr(extra), out.data_ptr(),
                                    ws.data_ptr(), ws_bytes, L.stream_of(x)), "colsum")
    return out


def _f32(t):
    return None if t is None else t.detach().contiguous().float()


def _wgrad(dy_op: Operand, x_op: Operand, M: int, N: int, K: int) -> torch.Tensor:
    """dW[N,K] = dY^T . X with both operands read MN-major (as stored), split-K over the tokens."""
    lib = L.load()
    splits = lib.b200swin_gemm_splits(N, K, M)
    return gemm(dy_op, x_op, N, K, M, a_mn=True, b_mn=True, out_dtype=torch.float32, splits=splits)


def _dgrad_plus(dyo: Operand, wo: Operand, M: int, K: int, N: int, cd, dres):
    """dX[M,K] = dY . W (W read as stored) + dres, the addition fused into the GEMM epilogue when dres has the
    output's type and shape."""
    if dres is None:
        return gemm(dyo, wo, M, K, N, b_mn=True, out_dtype=cd)
    r = dres.reshape(M, K)
    if r.dtype != cd:
        r = r.to(cd)
    return gemm(dyo, wo, M, K, N, b_mn=True, epilogue=L.EPI_ADD, aux_in=r.contiguous(), out_dtype=cd)


class _Linear(torch.autograd.Function):
    """y = x @ W^T + b on tcgen05 (reference: nn.Linear at swin_transformer_v2.py:334 and others)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        L.require_cuda(x, weight, bias)
        cd = compute_dtype(x)
        exact = cd == torch.float32
        N, K = weight.shape
        x2 = x.reshape(-1, K)
        M = x2.shape[0]
        if cd == torch.bfloat16 and x2.dtype != torch.bfloat16:
            x2 = x2.to(torch.bfloat16)
        xo = stage_operand(x2, exact)
        wo = stage_weight(weight, exact)
        y = gemm(xo, wo, M, N, K, bias=_f32(bias), out_dtype=cd)
        ctx.save_for_backward(xo.hi, xo.lo, weight)
        ctx.has_bias = bias is not None
        ctx.cd, ctx.xdtype, ctx.xshape = cd, x.dtype, x.shape
        ctx.bdtype = bias.dtype if bias is not None else None
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        xhi, xlo, weight = ctx.saved_tensors
        exact = ctx.cd == torch.float32
        N, K = weight.shape
        dy2 = dy.reshape(-1, N)
        M = dy2.shape[0]
        if dy2.dtype != ctx.cd:
            dy2 = dy2.to(ctx.cd)
        dyo = stage_operand(dy2, exact)
        wo = stage_weight(weight, exact)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = gemm(dyo, wo, M, K, N, b_mn=True, out_dtype=ctx.cd).view(ctx.xshape)
            if dx.dtype != ctx.xdtype:
                dx = dx.to(ctx.xdtype)
        if ctx.needs_input_grad[1]:
            dw = _wgrad(dyo, Operand(xhi, xlo), M, N, K).to(weight.dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = colsum(dy2.contiguous()).to(ctx.bdtype)
        return dx, dw, db


def linear(x, weight, bias=None):
    return _Linear.apply(x, weight, bias)


class _Mlp(torch.autograd.Function):
    """fc2(GELU(fc1(x))) with the GELU fused into fc1's epilogue -- which also emits gelu'(pre-activation) for the
    backward -- and the multiplication by that derivative fused into fc2's dgrad epilogue
    (reference: Mlp.forward, swin_transformer_v2.py:76-89; exact-erf GELU).  relu=True: the ReLU feed-forward of
    Transformer_Encoder (models/cnn_transformer.py:193-195, :205-207) through the same two GEMMs."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, passthrough=False, relu=False, will_backward=True):
        # passthrough: also return x itself (as a view).  The caller routes its residual branch through that alias, so
        # the gradient of the residual arrives HERE and is added in the epilogue of the last dgrad GEMM instead of in
        # a separate elementwise pass by the autograd engine.
        L.require_cuda(x, w1, b1, w2, b2)
        cd = compute_dtype(x)
        exact = cd == torch.float32
        Hd, C = w1.shape
        Co = w2.shape[0]
        x2 = x.reshape(-1, C)
        M = x2.shape[0]
        if cd == torch.bfloat16 and x2.dtype != torch.bfloat16:
            x2 = x2.to(torch.bfloat16)
        xo = stage_operand(x2, exact)
        # the activation's derivative is written only when a backward will read it
        z = torch.empty((M, Hd), dtype=cd, device=x.device) if (will_backward and any(ctx.needs_input_grad)) else None
        h = gemm(xo, stage_weight(w1, exact), M, Hd, C, epilogue=L.EPI_RELU if relu else L.EPI_GELU, bias=_f32(b1),
                 aux_out=z, out_dtype=cd)
        ho = stage_operand(h, exact)
        y = gemm(ho, stage_weight(w2, exact), M, Co, Hd, bias=_f32(b2), out_dtype=cd)
        ctx.save_for_backward(xo.hi, xo.lo, z, ho.hi, ho.lo, w1, w2)
        ctx.cd, ctx.xdtype, ctx.xshape = cd, x.dtype, x.shape
        ctx.bd = (b1.dtype if b1 is not None else None, b2.dtype if b2 is not None else None)
        ctx.set_materialize_grads(False)
        y = y.view(*x.shape[:-1], Co)
        return (y, x.view_as(x)) if passthrough else y

    @staticmethod
    def backward(ctx, dy, dres=None):
        xhi, xlo, z, hhi, hlo, w1, w2 = ctx.saved_tensors
        if dy is None:                                   # only the alias was used downstream
            return dres, None, None, None, None, None, None, None
        exact = ctx.cd == torch.float32
        Hd, C = w1.shape
        Co = w2.shape[0]
        dy2 = dy.reshape(-1, Co)
        M = dy2.shape[0]
        if dy2.dtype != ctx.cd:
            dy2 = dy2.to(ctx.cd)
        dy2 = dy2.contiguous()
        dyo = stage_operand(dy2, exact)
        dw2 = _wgrad(dyo, Operand(hhi, hlo), M, Co, Hd).to(w2.dtype)
        db2 = colsum(dy2).to(ctx.bd[1]) if (ctx.bd[1] is not None and ctx.needs_input_grad[4]) else None
        dz = gemm(dyo, stage_weight(w2, exact), M, Hd, Co, b_mn=True, epilogue=L.EPI_DGELU, aux_in=z, out_dtype=ctx.cd)
        dzo = stage_operand(dz, exact)
        dw1 = _wgrad(dzo, Operand(xhi, xlo), M, Hd, C).to(w1.dtype)
        db1 = colsum(dz).to(ctx.bd[0]) if (ctx.bd[0] is not None and ctx.needs_input_grad[2]) else None
        dx = None
        if ctx.needs_input_grad[0]:
            dx = _dgrad_plus(dzo, stage_weight(w1, exact), M, C, Hd, ctx.cd, dres).view(ctx.xshape)
            if dx.dtype != ctx.xdtype:
                dx = dx.to(ctx.xdtype)
        return dx, dw1, db1, dw2, db2, None, None, None


def mlp(x, w1, b1, w2, b2, passthrough=False, relu=False):
    # grad mode is read HERE: inside Function.forward it is always off
    return _Mlp.apply(x, w1, b1, w2, b2, passthrough, relu, torch.is_grad_enabled())


class _QKV(torch.autograd.Function):
    """qkv = x @ Wqkv^T + (q_bias, 0, v_bias) with q and k L2-normalised per head in the GEMM epilogue
    (reference: swin_transformer_v2.py:283-293).  Returns (qkv_hat [T,3C], inv_norm [T,2,nH]).
    PRIVATE CONTRACT with _AttnCore: the gradient arriving for qkv_hat is already the gradient w.r.t. the
    UN-normalised q, k (the attention backward applies the F.normalize backward with inv_norm)."""

    @staticmethod
    def forward(ctx, x, weight, q_bias, v_bias, nH, passthrough=False):
        # passthrough: see _Mlp.forward
        L.require_cuda(x, weight, q_bias, v_bias)
        cd = compute_dtype(x)
        exact = cd == torch.float32
        N, K = weight.shape
        x2 = x.reshape(-1, K)
        M = x2.shape[0]
        if cd == torch.bfloat16 and x2.dtype != torch.bfloat16:
            x2 = x2.to(torch.bfloat16)
        xo = stage_operand(x2, exact)
        inv_norm = torch.empty((M, 2, nH), dtype=torch.float32, device=x.device)
        qkv = gemm(xo, stage_weight(weight, exact), M, N, K, epilogue=L.EPI_QKV, bias=_f32(q_bias), bias2=_f32(v_bias),
                   inv_norm=inv_norm, nH=nH, out_dtype=cd)
        ctx.save_for_backward(xo.hi, xo.lo, weight)
        ctx.cd, ctx.xdtype, ctx.xshape = cd, x.dtype, x.shape
        ctx.has_bias = q_bias is not None
        ctx.bdtype = q_bias.dtype if q_bias is not None else None
        ctx.mark_non_differentiable(inv_norm)
        ctx.set_materialize_grads(False)
        return (qkv, inv_norm, x.view_as(x)) if passthrough else (qkv, inv_norm)

    @staticmethod
    def backward(ctx, dqkv, _, dres=None):
        xhi, xlo, weight = ctx.saved_tensors
        if dqkv is None:
            return dres, None, None, None, None, None
        exact = ctx.cd == torch.float32
        N, K = weight.shape
        C = N // 3
        d2 = dqkv.reshape(-1, N)
        M = d2.shape[0]
        if d2.dtype != ctx.cd:
            d2 = d2.to(ctx.cd)
        d2 = d2.contiguous()
        do = stage_operand(d2, exact)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = _dgrad_plus(do, stage_weight(weight, exact), M, K, N, ctx.cd, dres).view(ctx.xshape)
            if dx.dtype != ctx.xdtype:
                dx = dx.to(ctx.xdtype)
        dw = _wgrad(do, Operand(xhi, xlo), M, N, K).to(weight.dtype)
        dqb = dvb = None
        if ctx.has_bias:
            cs = _take_colsum(dqkv)            # from the attention backward's epilogue when it provides them
            if cs is None:
                cs = colsum(d2)                # one pass over all 3C columns (two launches) instead of two over C each
            dqb = cs[:C].to(ctx.bdtype)
            dvb = cs[2 * C:].to(ctx.bdtype)
        return dx, dw, dqb, dvb, None, None


# Column sums of dqkv handed from _AttnCore.backward to _QKV.backward.  Autograd re-wraps the gradient (a view in
# between), so the hand-over is keyed by the data pointer; the entry keeps the gradient tensor alive until it is taken, so
# the address cannot be reused by another tensor meanwhile.  Unclaimed entries (a qkv that needs no gradient) are dropped.
_DQKV_COLSUM = {}


def _stash_colsum(dqkv, sums):
    if len(_DQKV_COLSUM) >= 8:
        _DQKV_COLSUM.clear()
    _DQKV_COLSUM[dqkv.data_ptr()] = (dqkv, sums)


def _take_colsum(dqkv):
    ent = _DQKV_COLSUM.pop(dqkv.data_ptr(), None)
    if ent is None or ent[0].shape.numel() != dqkv.shape.numel() or ent[0].dtype != dqkv.dtype:
        return None
    return ent[1]


def qkv_project(x, weight, q_bias, v_bias, nH, passthrough=False):
    return _QKV.apply(x, weight, q_bias, v_bias, int(nH), passthrough)


# ------------------------------------------------------------------------------ continuous position bias table
class _CpbTable(torch.autograd.Function):
    """table16[T,nH] = 16 sigmoid(relu(coords W0^T + b0) W2^T) and scale[nH] = exp(min(logit_scale, ln 100)) in one
    kernel each way (reference: rpe_mlp + sigmoid, swin_transformer_v2.py:304-313; logit_scale clamp + exp, :294)."""

    @staticmethod
    def forward(ctx, coords, w0, b0, w2, logit_scale):
        L.require_cuda(coords, w0, b0, w2, logit_scale)
        lib = L.load()
        c = coords.reshape(-1, 2).contiguous().float()
        w0c, b0c, w2c = w0.contiguous().float(), b0.contiguous().float(), w2.contiguous().float()
        lsc = logit_scale.reshape(-1).contiguous().float()
        T, HID, nH = c.shape[0], w0c.shape[0], w2c.shape[0]
        with torch.cuda.device_of(c):
            table = torch.empty((T, nH), dtype=torch.float32, device=c.device)
            scale = torch.empty((nH,), dtype=torch.float32, device=c.device)
            L.check(lib.b200swin_cpb_fwd(c.data_ptr(), w0c.data_ptr(), b0c.data_ptr(), w2c.data_ptr(), table.data_ptr(),
                                         lsc.data_ptr(), scale.data_ptr(), T, HID, nH, L.stream_of(c)), "cpb_fwd")
        ctx.save_for_backward(c, w0c, b0c, w2c, table, lsc)
        ctx.dtypes = (w0.dtype, b0.dtype, w2.dtype, logit_scale.dtype)
        ctx.ls_shape = logit_scale.shape
        ctx.set_materialize_grads(False)
        return table, scale

    @staticmethod
    def backward(ctx, dtable, dscale):
        c, w0c, b0c, w2c, table, lsc = ctx.saved_tensors
        lib = L.load()
        T, HID, nH = c.shape[0], w0c.shape[0], w2c.shape[0]
        with torch.cuda.device_of(c):
            dt = torch.zeros_like(table) if dtable is None else dtable.contiguous().float()
            ds = torch.zeros_like(lsc) if dscale is None else dscale.reshape(-1).contiguous().float()
            dw0, db0, dw2, dls = torch.empty_like(w0c), torch.empty_like(b0c), torch.empty_like(w2c), torch.empty_like(lsc)
            L.check(lib.b200swin_cpb_bwd(c.data_ptr(), w0c.data_ptr(), b0c.data_ptr(), w2c.data_ptr(), table.data_ptr(),
                                         dt.data_ptr(), dw0.data_ptr(), db0.data_ptr(), dw2.data_ptr(), lsc.data_ptr(),
                                         ds.data_ptr(), dls.data_ptr(), T, HID, nH, L.stream_of(c)), "cpb_bwd")
        d0, d1, d2, d3 = ctx.dtypes
        return None, dw0.to(d0), db0.to(d1), dw2.to(d2), dls.view(ctx.ls_shape).to(d3)


def cpb_table_and_scale(coords, w0, b0, w2, logit_scale):
    return _CpbTable.apply(coords, w0, b0, w2, logit_scale)


def cpb_table(coords, w0, b0, w2):
    nH = w2.shape[0]
    return _CpbTable.apply(coords, w0, b0, w2, torch.zeros(nH, device=w2.device))[0]


# ------------------------------------------------------------------------------ attention core
# "auto": tcgen05 kernels whenever they apply (bf16 storage, on-the-fly mask; any window up to 32x32: the single-tile
# kernels for windows 4/6/7/8/12, the KV-blocked kernels otherwise), else the fp32 CUDA-core kernels (fp32 tensors = the
# reference-precision mode, or an explicit mask tensor).  "simt" / "tc" force one implementation (tests, A/B timing).
ATTN_IMPL = {"mode": "auto", "bwd_mode": "auto"}


def _tc_applicable(dtype, ws, mask, backward=False):
    return dtype == torch.bfloat16 and mask is None and 1 <= ws <= 32


def _pick_impl(dtype, ws, mask, backward=False):
    mode = ATTN_IMPL["bwd_mode" if backward else "mode"]
    if mode == "simt":
        return 0
    if mode == "tc":
        return 1
    if mode == "flash":
        return 2
    if mode == "ws":                      # single-tile tcgen05 kernels (A/B timing against the warp-MMA kernels)
        return 3
    if mode == "mma":
        return 4
    return 1 if _tc_applicable(dtype, ws, mask, backward) else 0


class _AttnCore(torch.autograd.Function):
    """softmax(scale * q_hat k_hat^T + table16[rel] + mask) @ v over shifted windows, reading and writing the
    natural [B,H,W,*] layout (pad/roll/partition/reverse/crop and the shift mask are address math).
    Reference: swin_transformer_v2.py:295-328 + :429-463 + :874-892."""

    @staticmethod
    def forward(ctx, qkv, inv_norm, table16, scale, qpad, vpad, mask, geom, will_backward=True):
        B, H, W, C, nH, ws, shift = geom
        L.require_cuda(qkv, inv_norm, table16, scale, qpad, vpad, mask)
        lib = L.load()
        qkv = qkv.contiguous()
        t16 = table16.contiguous().float()
        sc = scale.contiguous().float()
        qp, vp, mk = _f32(qpad), _f32(vpad), _f32(mask)
        Hp, Wp = (H + ws - 1) // ws * ws, (W + ws - 1) // ws * ws
        nwin = B * (Hp // ws) * (Wp // ws)
        impl = _pick_impl(qkv.dtype, ws, mask)
        ctx.impl_bwd = _pick_impl(qkv.dtype, ws, mask, backward=True)
        if inv_norm is None:
            # un-normalised q, k (attn_type='normal'): the KV-blocked kernels (true row maximum) or the CUDA-core ones
            impl = ctx.impl_bwd = 2 if (impl != 0 and ctx.impl_bwd != 0) else 0
        nWm = mask.shape[0] if mask is not None else 0
        with torch.cuda.device_of(qkv):
            out = torch.empty((B, H, W, C), dtype=qkv.dtype, device=qkv.device)
            # bf16 residual of O for the backward's D = <dO, O> (tensor-core path, only when a backward will run)
            need_lo = impl != 0 and will_backward and any(ctx.needs_input_grad)
            out_lo = torch.empty_like(out) if need_lo else None
            lse = torch.empty((nwin, nH, ws * ws), dtype=torch.float32, device=qkv.device)
            L.check(lib.b200swin_attn_fwd(qkv.data_ptr(), out.data_ptr(), L.ptr(out_lo), lse.data_ptr(), t16.data_ptr(),
                                          sc.data_ptr(), L.ptr(qp), L.ptr(vp), L.ptr(mk), nWm, B, H, W, C, nH, ws, shift,
                                          L.dtype_code(qkv), impl, L.stream_of(qkv)), "attn_fwd")
        ctx.save_for_backward(qkv, out, out_lo, lse, inv_norm, t16, sc, qp, vp, mk)
        ctx.geom, ctx.impl, ctx.nWm = geom, impl, nWm
        ctx.dtypes = (table16.dtype, scale.dtype, vpad.dtype if vpad is not None else None)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, out_lo, lse, inv_norm, t16, sc, qp, vp, mk = ctx.saved_tensors
        B, H, W, C, nH, ws, shift = ctx.geom
        lib = L.load()
        dout = dout.contiguous()
        if dout.dtype != qkv.dtype:
            dout = dout.to(qkv.dtype)
        with torch.cuda.device_of(qkv):
            dqkv = torch.empty_like(qkv)
            # column sums of dq / dv (q_bias / v_bias gradients) straight from the backward's epilogue where the kernel
            # can: saves the qkv projection's backward a pass over dqkv
            want_cs = bool(lib.b200swin_attn_bwd_colsum_supported(ws, L.dtype_code(qkv), ctx.impl_bwd))
            acc = torch.zeros(t16.numel() + nH + C + (3 * C if want_cs else 0), dtype=torch.float32, device=qkv.device)
            dt16 = acc[:t16.numel()].view_as(t16)
            dsc = acc[t16.numel():t16.numel() + nH]
            dvp = acc[t16.numel() + nH:t16.numel() + nH + C]
            dcs = acc[t16.numel() + nH + C:] if want_cs else None
            ws_bytes = lib.b200swin_attn_bwd_workspace_bytes(B, H, W, nH, ws, L.dtype_code(qkv), ctx.impl_bwd)
            wsp = torch.empty(ws_bytes, dtype=torch.uint8, device=qkv.device) if ws_bytes else None
            L.check(lib.b200swin_attn_bwd(qkv.data_ptr(), out.data_ptr(), L.ptr(out_lo) if ctx.impl_bwd != 0 else 0,
                                          dout.data_ptr(), lse.data_ptr(),
                                          L.ptr(inv_norm), t16.data_ptr(), sc.data_ptr(), L.ptr(qp), L.ptr(vp),
                                          L.ptr(mk), ctx.nWm, dqkv.data_ptr(), dt16.data_ptr(), dsc.data_ptr(),
                                          dvp.data_ptr(), L.ptr(dcs), B, H, W, C, nH, ws, shift, L.dtype_code(qkv),
                                          ctx.impl_bwd, L.ptr(wsp), ws_bytes, L.stream_of(qkv)), "attn_bwd")
        if want_cs:
            _stash_colsum(dqkv, dcs)
        tdt, sdt, vdt = ctx.dtypes
        return (dqkv, None, dt16.to(tdt), dsc.view(sc.shape).to(sdt), None,
                dvp.to(vdt) if vdt is not None else None, None, None, None)


def attention_core(qkv, inv_norm, table16, scale, qpad, vpad, mask, B, H, W, C, nH, ws, shift):
    return _AttnCore.apply(qkv, inv_norm, table16, scale, qpad, vpad, mask, (B, H, W, C, nH, ws, shift),
                           torch.is_grad_enabled())


# ------------------------------------------------------------------------------ global multi-head attention
def _mha_rows(t: torch.Tensor, what: str) -> torch.Tensor:
    """[B, N, E] with unit column stride and B*N rows at ONE row stride (a column slice of a packed projection
    buffer qualifies); anything else is made contiguous."""
    if t.dim() != 3:
        raise ValueError(f"mha: {what} must be [batch, tokens, channels]")
    ok = (t.stride(2) == 1 and t.stride(1) % 8 == 0 and t.stride(0) == t.shape[1] * t.stride(1)
          and t.data_ptr() % 16 == 0)
    return t if ok else t.contiguous()


class _MhaCore(torch.autograd.Function):
    """out = softmax(scale * q k^T) v per (batch, head) over whole sequences -- the scaled-dot-product inside
    nn.MultiheadAttention (reference: Transformer_Encoder.forward, models/cnn_transformer.py:198-201).
    `packed` says how the projected operands arrive, so that their gradients leave in the same buffers (no slicing
    copies on either side):  'qkv' a = [B,N,3E];  'qk_v' a = [B,N,2E] (q | k), b = v;  'q_k_v' three tensors."""

    @staticmethod
    def forward(ctx, a, b, c, packed, nH, scale):
        L.require_cuda(a, b, c)
        lib = L.load()
        if packed == 'qkv':
            a = _mha_rows(a, 'qkv')
            E = a.shape[2] // 3
            q, k, v = a[..., :E], a[..., E:2 * E], a[..., 2 * E:]
        elif packed == 'qk_v':
            a, b = _mha_rows(a, 'qk'), _mha_rows(b, 'v')
            E = a.shape[2] // 2
            q, k, v = a[..., :E], a[..., E:], b
        else:
            a, b, c = _mha_rows(a, 'q'), _mha_rows(b, 'k'), _mha_rows(c, 'v')
            E = a.shape[2]
            q, k, v = a, b, c
        if not (q.dtype == k.dtype == v.dtype):
            raise TypeError("mha: q, k, v must share a dtype")
        if E % nH or E // nH not in (32, 64):
            raise NotImplementedError("mha: head_dim must be 32 or 64")
        B, Nq, Nk, hd = q.shape[0], q.shape[1], k.shape[1], E // nH
        if k.shape[0] != B or v.shape[0] != B or v.shape[1] != Nk or k.shape[2] != E or v.shape[2] != E:
            raise ValueError("mha: inconsistent q / k / v shapes")
        with torch.cuda.device_of(q):
            out = torch.empty((B, Nq, E), dtype=q.dtype, device=q.device)
            lse = torch.empty((B, nH, Nq), dtype=torch.float32, device=q.device)
            L.check(lib.b200swin_mha_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), q.stride(1), k.stride(1), v.stride(1),
                                         out.data_ptr(), E, lse.data_ptr(), B, Nq, Nk, nH, hd, float(scale),
                                         L.dtype_code(q), L.stream_of(q)), "mha_fwd")
        ctx.save_for_backward(a, b, c, out, lse)
        ctx.packed, ctx.nH, ctx.scale, ctx.E = packed, nH, float(scale), E
        ctx.mark_non_differentiable(lse)
        ctx.set_materialize_grads(False)
        return out, lse

    @staticmethod
    def backward(ctx, dout, _dlse=None):
        a, b, c, out, lse = ctx.saved_tensors
        lib = L.load()
        E, nH, packed = ctx.E, ctx.nH, ctx.packed
        if packed == 'qkv':
            q, k, v = a[..., :E], a[..., E:2 * E], a[..., 2 * E:]
        elif packed == 'qk_v':
            q, k, v = a[..., :E], a[..., E:], b
        else:
            q, k, v = a, b, c
        B, Nq, Nk, hd = q.shape[0], q.shape[1], k.shape[1], E // nH
        if dout is None:
            dout = torch.zeros_like(out)
        dout = dout.contiguous()
        if dout.dtype != q.dtype:
            dout = dout.to(q.dtype)
        with torch.cuda.device_of(q):
            if packed == 'qkv':
                da = torch.empty(a.shape, dtype=a.dtype, device=a.device)
                dq, dk, dv = da[..., :E], da[..., E:2 * E], da[..., 2 * E:]
                grads = (da, None, None)
            elif packed == 'qk_v':
                da = torch.empty(a.shape, dtype=a.dtype, device=a.device)
                dq, dk = da[..., :E], da[..., E:]
                dv = torch.empty(v.shape, dtype=v.dtype, device=v.device)
                grads = (da, dv, None)
            else:
                dq = torch.empty(q.shape, dtype=q.dtype, device=q.device)
                dk = torch.empty(k.shape, dtype=k.dtype, device=k.device)
                dv = torch.empty(v.shape, dtype=v.dtype, device=v.device)
                grads = (dq, dk, dv)
            ws_bytes = lib.b200swin_mha_bwd_workspace_bytes(B, Nq, nH)
            wsp = torch.empty(ws_bytes, dtype=torch.uint8, device=q.device)
            L.check(lib.b200swin_mha_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), q.stride(1), k.stride(1), v.stride(1),
                                         out.data_ptr(), E, dout.data_ptr(), E, lse.data_ptr(), dq.data_ptr(), dk.data_ptr(),
                                         dv.data_ptr(), dq.stride(1), dk.stride(1), dv.stride(1), B, Nq, Nk, nH, hd,
                                         ctx.scale, L.dtype_code(q), wsp.data_ptr(), ws_bytes, L.stream_of(q)), "mha_bwd")
        return grads + (None, None, None)


def mha_core(q, k, v, nH, scale=None, packed='q_k_v'):
    """(out [B,Nq,E], lse [B,nH,Nq]).  packed='qkv': q is the [B,N,3E] projection; 'qk_v': q is [B,N,2E], k is v."""
    if packed == 'qkv':
        E, args = q.shape[2] // 3, (q, None, None)
    elif packed == 'qk_v':
        E, args = q.shape[2] // 2, (q, k, None)
    else:
        E, args = q.shape[2], (q, k, v)
    if scale is None:
        scale = (E // nH) ** -0.5
    return _MhaCore.apply(*args, packed, nH, scale)


def mha_avg_weights(q, k, lse, nH, scale=None):
    """Head-averaged attention probabilities [B,Nq,Nk] (need_weights=True of nn.MultiheadAttention); no gradient."""
    L.require_cuda(q, k, lse)
    lib = L.load()
    q, k = _mha_rows(q.detach(), 'q'), _mha_rows(k.detach(), 'k')
    B, Nq, E = q.shape
    Nk = k.shape[1]
    if scale is None:
        scale = (E // nH) ** -0.5
    with torch.cuda.device_of(q):
        w = torch.empty((B, Nq, Nk), dtype=q.dtype, device=q.device)
        L.check(lib.b200swin_mha_avg_weights(q.data_ptr(), k.data_ptr(), q.stride(1), k.stride(1),
                                             lse.contiguous().data_ptr(), w.data_ptr(), B, Nq, Nk, nH, E // nH,
                                             float(scale), L.dtype_code(q), L.stream_of(q)), "mha_avg_weights")
    return w


# ------------------------------------------------------------------------------ depthwise 3x3 conv (ConvMlp.conv_proj)
class _DwConv3x3(torch.autograd.Function):
    """y = depthwise_conv3x3(x) on [B,H,W,C] tokens (reference: ConvMlp.conv_proj, swin_transformer_v2.py:98-111)."""

    @staticmethod
    def forward(ctx, x, weight):
        L.require_cuda(x, weight)
        lib = L.load()
        B, H, W, C = x.shape
        xc = x.contiguous()
        w32 = weight.detach().contiguous().float()
        with torch.cuda.device_of(xc):
            y = torch.empty_like(xc)
            L.check(lib.b200swin_dwconv3x3(xc.data_ptr(), w32.data_ptr(), y.data_ptr(), B, H, W, C, L.dtype_code(xc), 0,
                                           L.stream_of(xc)), "dwconv3x3")
        ctx.save_for_backward(xc, w32)
        ctx.wdtype = weight.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, w32 = ctx.saved_tensors
        lib = L.load()
        B, H, W, C = xc.shape
        dyc = dy.contiguous()
        if dyc.dtype != xc.dtype:
            dyc = dyc.to(xc.dtype)
        dx = dw = None
        with torch.cuda.device_of(xc):
            if ctx.needs_input_grad[0]:
                dx = torch.empty_like(xc)
                L.check(lib.b200swin_dwconv3x3(dyc.data_ptr(), w32.data_ptr(), dx.data_ptr(), B, H, W, C, L.dtype_code(xc), 1,
                                               L.stream_of(xc)), "dwconv3x3 (adjoint)")
            if ctx.needs_input_grad[1]:
                dw = torch.empty((C, 1, 3, 3), dtype=torch.float32, device=xc.device)
                ws_bytes = lib.b200swin_dwconv3x3_wgrad_workspace_bytes(B, H, W, C)
                wsp = torch.empty(ws_bytes, dtype=torch.uint8, device=xc.device)
                L.check(lib.b200swin_dwconv3x3_wgrad(xc.data_ptr(), dyc.data_ptr(), dw.data_ptr(), B, H, W, C,
                                                     L.dtype_code(xc), wsp.data_ptr(), ws_bytes, L.stream_of(xc)),
                        "dwconv3x3_wgrad")
                dw = dw.to(ctx.wdtype)
        return dx, dw


def dwconv3x3(x, weight):
    return _DwConv3x3.apply(x, weight)


# ------------------------------------------------------------------------------ post-norm transformer layer: LN(x + branch)
class _LayerNormSum(torch.autograd.Function):
    """(y32, y16) = LN(x + xadd) * gamma + beta in one kernel (reference: `x = v + attn; x = norm1(x)` and
    `x = x + ffn; x = norm2(x)`, models/cnn_transformer.py:202-203, :208-209).  x is the fp32 stream, xadd the branch output
    (fp32, or bf16 under autocast); y16 = bf16(y32) is the operand of the next GEMM (None unless asked for).  The backward
    normalises the saved sum; its dx is the gradient of x and of xadd alike."""

    @staticmethod
    def forward(ctx, x, xadd, gamma, beta, eps, want16, will_backward):
        L.require_cuda(x, xadd, gamma, beta)
        lib = L.load()
        C = x.shape[-1]
        xc = x.contiguous()
        if xc.dtype != torch.float32:
            xc = xc.float()
        ac = xadd.contiguous()
        rows = xc.numel() // C
        g32 = gamma.contiguous().float()
        b32 = beta.contiguous().float()
        keep = will_backward and any(ctx.needs_input_grad)
        with torch.cuda.device_of(xc):
            y = torch.empty_like(xc)
            y16 = torch.empty(xc.shape, dtype=torch.bfloat16, device=xc.device) if want16 else None
            xsum = torch.empty_like(xc) if keep else None
            stats = torch.empty((2, rows), dtype=torch.float32, device=xc.device)
            L.check(lib.b200swin_ln_fwd_sum(xc.data_ptr(), ac.data_ptr(), L.dtype_code(ac), g32.data_ptr(), b32.data_ptr(),
                                            y.data_ptr(), L.ptr(y16), L.ptr(xsum), stats[0].data_ptr(), stats[1].data_ptr(),
                                            rows, C, float(eps), L.stream_of(xc)), "ln_fwd_sum")
        ctx.save_for_backward(xsum, g32, stats)
        ctx.dtypes = (x.dtype, xadd.dtype, gamma.dtype, beta.dtype)
        ctx.set_materialize_grads(False)
        return y.view(x.shape), (None if y16 is None else y16.view(x.shape))

    @staticmethod
    def backward(ctx, dy32, dy16=None):
        xsum, g32, stats = ctx.saved_tensors
        lib = L.load()
        C = xsum.shape[-1]
        rows = xsum.numel() // C
        if dy32 is None and dy16 is None:
            return None, None, None, None, None, None, None
        dy = dy32.float() if dy32 is not None else None
        if dy16 is not None:
            dy = dy16.float() if dy is None else dy + dy16.float()
        dyc = dy.contiguous()
        with torch.cuda.device_of(xsum):
            dx = torch.empty_like(xsum)
            dgb = torch.empty((2, C), dtype=torch.float32, device=xsum.device)
            ws_bytes = lib.b200swin_ln_bwd_workspace_bytes(rows, C)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xsum.device)
            L.check(lib.b200swin_ln_bwd(dyc.data_ptr(), xsum.data_ptr(), g32.data_ptr(), stats[0].data_ptr(),
                                        stats[1].data_ptr(), 0, 1, dx.data_ptr(), dgb[0].data_ptr(), dgb[1].data_ptr(), 0,
                                        rows, C, L.F32, ws.data_ptr(), ws_bytes, L.stream_of(xsum)), "ln_bwd")
        xd, ad, gd, bd = ctx.dtypes
        return (dx if xd == torch.float32 else dx.to(xd), dx if ad == torch.float32 else dx.to(ad),
                dgb[0].to(gd), dgb[1].to(bd), None, None, None)


_TWIN16 = "_b200swin_bf16_twin"


def layer_norm_sum(x, xadd, gamma, beta, eps, want16=False):
    """(y32, y16 | None) = LN(x + xadd); y32 carries y16 as an attribute so that the next layer's GEMMs can read it."""
    y32, y16 = _LayerNormSum.apply(x, xadd, gamma, beta, float(eps), bool(want16), torch.is_grad_enabled())
    if y16 is not None:
        setattr(y32, _TWIN16, y16)
    return y32, y16


def bf16_twin_of(x: torch.Tensor):
    """The bf16 copy layer_norm_sum wrote beside an fp32 tensor (None when there is none or it does not match)."""
    t = getattr(x, _TWIN16, None)
    return t if (t is not None and t.shape == x.shape and t.device == x.device) else None
