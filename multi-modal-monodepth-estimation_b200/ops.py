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
class _LayerNormResidual(torch.autograd.Function):
    """y = residual + row_scale[b] * (LN(x) * gamma + beta); residual / row_scale optional."""

    @staticmethod
    def forward(ctx, x, residual, gamma, beta, row_scale, rows_per_scale, eps):
        L.require_cuda(x, residual, gamma, beta, row_scale)
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
        with torch.cuda.device_of(xc):
            y = torch.empty_like(xc)
            stats = torch.empty((2, rows), dtype=torch.float32, device=x.device)
            L.check(lib.b200swin_ln_fwd(xc.data_ptr(), L.ptr(rc), g32.data_ptr(), b32.data_ptr(), L.ptr(rs),
                                        rows_per_scale, y.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(),
                                        rows, C, eps, L.dtype_code(xc), L.stream_of(xc)), "ln_fwd")
        ctx.save_for_backward(xc, g32, stats, rs)
        ctx.has_res = residual is not None
        ctx.rows_per_scale = rows_per_scale
        ctx.gdtype, ctx.bdtype = gamma.dtype, beta.dtype
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        xc, g32, stats, rs = ctx.saved_tensors
        lib = L.load()
        C = xc.shape[-1]
        rows = xc.numel() // C
        dyc = dy.contiguous()
        if dyc.dtype != xc.dtype:
            dyc = dyc.to(xc.dtype)
        with torch.cuda.device_of(xc):
            dx = torch.empty_like(xc)
            dgb = torch.empty((2, C), dtype=torch.float32, device=xc.device)
            ws_bytes = lib.b200swin_ln_bwd_workspace_bytes(rows, C)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xc.device)
            L.check(lib.b200swin_ln_bwd(dyc.data_ptr(), xc.data_ptr(), g32.data_ptr(), stats[0].data_ptr(),
                                        stats[1].data_ptr(), L.ptr(rs), ctx.rows_per_scale, dx.data_ptr(),
                                        dgb[0].data_ptr(), dgb[1].data_ptr(), rows, C, L.dtype_code(xc),
                                        ws.data_ptr(), ws_bytes, L.stream_of(xc)), "ln_bwd")
        dres = dyc.view(dy.shape) if ctx.has_res else None
        return dx, dres, dgb[0].to(ctx.gdtype), dgb[1].to(ctx.bdtype), None, None, None


def layer_norm_residual(x, gamma, beta, eps, residual=None, row_scale=None, rows_per_scale=1):
    return _LayerNormResidual.apply(x, residual, gamma, beta, row_scale, int(rows_per_scale), float(eps))
