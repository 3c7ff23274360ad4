"""Optimizer step of the hot path: the reference's layer-decay parameter grouping + a fused multi-tensor AdamW.

Reference: ``models/optimizer.py`` -- ``get_num_layer_for_swin`` (:14-32) maps a parameter name to its depth,
``SwinLayerDecayOptimizerConstructor.add_params`` (:36-104) turns that into ~60 parameter groups
(``layer_<id>_{decay,no_decay}`` with ``lr_scale = layer_decay_rate ** (num_layers - id - 1)``) for
``torch.optim.AdamW``; ``train.py:195-203`` then rewrites every group's ``lr = current_lr * lr_scale`` each step.

Here the grouping is reproduced name for name (``layer_decay_param_groups``; a CPU test compares it with the reference's
constructor run through the mmcv shim), and ``FusedAdamW`` executes all groups in ONE kernel launch over flat fp32
buffers (``csrc/adamw.cu``): per-tensor ``lr_scale`` / ``weight_decay`` arrays, the step's base learning rate and the
step count in device memory (CUDA-graph capturable; a schedule rewrites one float), and the bf16 copies of the updated
weights -- the operands of the tcgen05 GEMMs -- written by the same pass.  ``param_groups`` keeps the reference's
contract (``lr``, ``lr_scale``, ``weight_decay``, ``param_names``), so the training loop's LR rewrite works unchanged.
"""
from __future__ import annotations

import torch

from . import _lib as L


# ------------------------------------------------------------------------------------------ grouping
def get_num_layer_for_swin(var_name: str, num_max_layer: int, layers_per_stage) -> int:
    """Depth of a parameter for layer-wise LR decay (reference models/optimizer.py:14-32; ``encoder.`` and
    ``backbone.`` prefixes are equivalent)."""
    if var_name.startswith("encoder"):
        var_name = var_name.replace("encoder", "backbone")
    if var_name in ("backbone.cls_token", "backbone.mask_token", "backbone.pos_embed", "backbone.absolute_pos_embed"):
        return 0
    if var_name.startswith("backbone.patch_embed"):
        return 0
    if var_name.startswith("backbone.layers"):
        parts = var_name.split(".")
        stage_id = int(parts[2])
        if parts[3] == "blocks":
            return int(parts[4]) + sum(layers_per_stage[:stage_id]) + 1
        if parts[3] == "downsample":
            return sum(layers_per_stage[:stage_id + 1])
        return None                      # the reference falls off its if-chain here as well
    return num_max_layer - 1


def layer_decay_param_groups(model: torch.nn.Module, base_lr: float, weight_decay: float, depths,
                             layer_decay_rate: float,
                             no_decay_names=("relative_position_bias_table", "rpe_mlp", "logit_scale")):
    """The parameter groups ``SwinLayerDecayOptimizerConstructor`` builds (reference models/optimizer.py:50-104,
    called from train.py:113-115), in the same order and with the same keys."""
    layers_per_stage = [int(d) for d in depths]
    for i in range(len(layers_per_stage) - 1):
        layers_per_stage[i] += 1                          # patch merging counts as a layer of its stage
    num_layers = sum(layers_per_stage) + 2                # + patch embed, + head
    groups: dict = {}
    for name, param in model.named_parameters():
        if not param.requires_grad:
            continue
        if len(param.shape) == 1 or name.endswith(".bias") or name in ("absolute_pos_embed",):
            kind, wd = "no_decay", 0.0
        else:
            kind, wd = "decay", weight_decay
            if any(nd in name for nd in no_decay_names):
                kind, wd = "no_decay", 0.0
        layer_id = get_num_layer_for_swin(name, num_layers, layers_per_stage)
        gname = "layer_%d_%s" % (layer_id, kind)
        if gname not in groups:
            scale = layer_decay_rate ** (num_layers - layer_id - 1)
            groups[gname] = {"weight_decay": wd, "params": [], "param_names": [], "lr_scale": scale,
                             "group_name": gname, "lr": scale * base_lr}
        groups[gname]["params"].append(param)
        groups[gname]["param_names"].append(name)
    return list(groups.values())


# ------------------------------------------------------------------------------------------ flat storage
class FlatParams:
    """Parameters, gradients (and optionally bf16 weight copies) of a model in flat buffers.  Every tensor starts on a
    chunk boundary (``b200swin_adamw_chunk()`` elements), so one optimizer launch and one all-reduce cover everything.
    The parameters are re-pointed at views of the flat buffer (``state_dict`` keys and values are unchanged)."""

    def __init__(self, params, bf16_copies: bool = True):
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("FlatParams: no trainable parameter")
        dev = params[0].device
        if any(p.device != dev or p.dtype != torch.float32 for p in params):
            raise RuntimeError("FlatParams: all parameters must be float32 tensors on one device")
        if bf16_copies and dev.type != "cuda":
            raise RuntimeError("FlatParams: bf16 weight copies feed the CUDA GEMMs; pass bf16_copies=False on the CPU "
                               "(layout / data-parallel host logic only -- the optimizer kernel has no CPU fallback)")
        self.chunk = int(L.load().b200swin_adamw_chunk())
        self.params = params
        self.offsets, off, chunk_tensor = [], 0, []
        for i, p in enumerate(params):
            self.offsets.append(off)
            n = (p.numel() + self.chunk - 1) // self.chunk
            chunk_tensor += [i] * n
            off += n * self.chunk
        self.total = off
        self.data = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.chunk_tensor = torch.tensor(chunk_tensor, dtype=torch.int32, device=dev)
        self.bf16 = torch.zeros(self.total, dtype=torch.bfloat16, device=dev) if bf16_copies else None
        self.grad_views, self.bf16_views = [], []
        with torch.no_grad():
            for p, o in zip(params, self.offsets):
                view = self.data[o:o + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view
                self.grad_views.append(self.grad[o:o + p.numel()].view_as(p))
        if bf16_copies:
            from . import ops
            self.bf16.copy_(self.data)
            for p, o in zip(params, self.offsets):
                v16 = self.bf16[o:o + p.numel()].view_as(p)
                self.bf16_views.append(v16)
                ops.register_staged_weight(p, v16)

    def pack_grads(self):
        """Gather the gradients autograd ASSIGNED to ``p.grad`` into the flat buffer (one multi-tensor copy); parameters
        without a gradient count as zero.  Returns the list of parameters that had none."""
        have = [(v, p.grad) for v, p in zip(self.grad_views, self.params) if p.grad is not None]
        missing = [v for v, p in zip(self.grad_views, self.params) if p.grad is None]
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        if missing:
            torch._foreach_zero_(missing)
        return missing

    def point_grads_at_flat(self):
        """Make ``p.grad`` the flat views (after the all-reduce wrote the averaged gradients there)."""
        for p, v in zip(self.params, self.grad_views):
            p.grad = v


class FusedAdamW(torch.optim.Optimizer):
    """``torch.optim.AdamW`` semantics, one launch for all parameter groups.  Accepts the same ``params`` argument
    (tensors or group dicts; extra keys such as ``lr_scale`` / ``param_names`` are kept).  ``flat``: share the flat
    buffers with a ``DataParallel`` wrapper.  Call ``set_lr(value)`` for a schedule under CUDA-graph replay; in eager
    mode rewriting ``param_groups[i]['lr']`` (train.py:202-203) is picked up by the next ``step()``."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, flat: FlatParams | None = None,
                 bf16_copies: bool = True):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        b = {tuple(g["betas"]) for g in self.param_groups}
        e = {float(g["eps"]) for g in self.param_groups}
        if len(b) != 1 or len(e) != 1:
            raise NotImplementedError("FusedAdamW: betas and eps must be the same for every group")
        self.betas, self.eps = b.pop(), e.pop()
        plist = [p for g in self.param_groups for p in g["params"] if p.requires_grad]
        self.flat = flat if flat is not None else FlatParams(plist, bf16_copies=bf16_copies)
        if [id(p) for p in self.flat.params] != [id(p) for p in plist]:
            raise ValueError("FusedAdamW: the shared FlatParams must hold the optimizer's parameters in group order")
        dev = self.flat.data.device
        L.require_cuda(self.flat.data)
        with torch.cuda.device(dev):
            self.exp_avg = torch.zeros_like(self.flat.data)
            self.exp_avg_sq = torch.zeros_like(self.flat.data)
            self.step_t = torch.zeros((), dtype=torch.float32, device=dev)
            self.lr_t = torch.zeros((), dtype=torch.float32, device=dev)
            self.lr_scale_t = torch.ones(len(plist), dtype=torch.float32, device=dev)
            self.wd_t = torch.zeros(len(plist), dtype=torch.float32, device=dev)
        self._uploaded = None
        self.grad_scale = 1.0
        self.grads_packed = False          # a DataParallel wrapper that already filled flat.grad sets this
        self._sync_hyper()

    def _per_tensor(self):
        lrs, wds = [], []
        for g in self.param_groups:
            n = sum(1 for p in g["params"] if p.requires_grad)
            lrs += [float(g["lr"])] * n
            wds += [float(g["weight_decay"])] * n
        return lrs, wds

    def _sync_hyper(self):
        """Upload per-tensor lr / weight decay when the host-side groups changed.  When every group's ``lr`` is
        ``base * lr_scale`` (the reference's schedule) only the scalar ``base`` moves."""
        lrs, wds = self._per_tensor()
        key = (tuple(lrs), tuple(wds))
        if key == self._uploaded:
            return
        scales = []
        for g in self.param_groups:
            n = sum(1 for p in g["params"] if p.requires_grad)
            scales += [float(g.get("lr_scale", 1.0))] * n
        base = None
        for lr, s in zip(lrs, scales):
            if s > 0:
                base = lr / s
                break
        uniform = base is not None and all(abs(lr - base * s) <= 1e-12 * max(1.0, abs(lr)) for lr, s in zip(lrs, scales))
        if not uniform:
            base, scales = 1.0, lrs
        self.lr_scale_t.copy_(torch.tensor(scales, dtype=torch.float32), non_blocking=True)
        self.wd_t.copy_(torch.tensor(wds, dtype=torch.float32), non_blocking=True)
        self.lr_t.fill_(base)
        self._uploaded = key

    def set_lr(self, base_lr: float):
        """Schedule hook that is safe between CUDA-graph replays: one device scalar, groups updated for bookkeeping."""
        for g in self.param_groups:
            g["lr"] = base_lr * float(g.get("lr_scale", 1.0))
        lrs, wds = self._per_tensor()
        if self._uploaded is not None and tuple(wds) == self._uploaded[1] and float(self.lr_scale_t.numel()) > 0:
            self.lr_t.fill_(base_lr)
            self._uploaded = (tuple(lrs), tuple(wds))
        else:
            self._sync_hyper()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if not torch.cuda.is_current_stream_capturing():
            self._sync_hyper()
        f = self.flat
        if not self.grads_packed:
            f.pack_grads()
        lib = L.load()
        with torch.cuda.device(f.data.device):
            self.step_t.add_(1.0)
            L.check(lib.b200swin_adamw_step(f.data.data_ptr(), f.grad.data_ptr(), self.exp_avg.data_ptr(),
                                            self.exp_avg_sq.data_ptr(), L.ptr(f.bf16), f.chunk_tensor.data_ptr(),
                                            self.lr_scale_t.data_ptr(), self.wd_t.data_ptr(), self.lr_t.data_ptr(),
                                            self.step_t.data_ptr(), self.betas[0], self.betas[1], self.eps,
                                            float(self.grad_scale), f.chunk_tensor.numel(),
                                            L.stream_of(f.data)), "adamw_step")
        return loss

    def zero_grad(self, set_to_none: bool = True):
        for p in self.flat.params:
            p.grad = None

    # flat state <-> the per-parameter layout of torch.optim.AdamW (load_model / save_model, utils/util.py:20-49)
    def state_dict(self):
        sd = super().state_dict()
        state = {}
        idx = 0
        for g in self.param_groups:
            for p in g["params"]:
                if p.requires_grad:
                    j = next(k for k, q in enumerate(self.flat.params) if q is p)
                    o = self.flat.offsets[j]
                    state[idx] = {"step": self.step_t.detach().clone(),
                                  "exp_avg": self.exp_avg[o:o + p.numel()].view_as(p).clone(),
                                  "exp_avg_sq": self.exp_avg_sq[o:o + p.numel()].view_as(p).clone()}
                idx += 1
        sd["state"] = state
        return sd

    def load_state_dict(self, state_dict):
        state = state_dict.get("state", {})
        groups = state_dict.get("param_groups")
        if groups is not None:
            for g, sg in zip(self.param_groups, groups):
                for k, v in sg.items():
                    if k != "params":
                        g[k] = v
        idx = 0
        with torch.no_grad():
            for g in self.param_groups:
                for p in g["params"]:
                    st = state.get(idx)
                    if st is not None and p.requires_grad:
                        j = next(k for k, q in enumerate(self.flat.params) if q is p)
                        o = self.flat.offsets[j]
                        self.exp_avg[o:o + p.numel()].view_as(p).copy_(st["exp_avg"])
                        self.exp_avg_sq[o:o + p.numel()].view_as(p).copy_(st["exp_avg_sq"])
                        self.step_t.fill_(float(st["step"]))
                    idx += 1
        self._uploaded = None
        self._sync_hyper()


def build_optimizers(model, cfg):
    """Drop-in for the reference's ``build_optimizers(model, cfg)`` call of train.py:113-115 for the AdamW +
    ``SwinLayerDecayOptimizerConstructor`` configuration (the only one the reference uses)."""
    if hasattr(model, "module"):
        model = model.module
    cfg = dict(cfg)
    if cfg.pop("type", "AdamW") != "AdamW":
        raise NotImplementedError("b200swin.build_optimizers: AdamW only")
    constructor = cfg.pop("constructor", None)
    pw = cfg.pop("paramwise_cfg", None) or {}
    lr, wd = cfg.pop("lr"), cfg.pop("weight_decay", 1e-2)
    if constructor == "SwinLayerDecayOptimizerConstructor":
        groups = layer_decay_param_groups(model, lr, wd, list(pw.get("num_layers")), pw.get("layer_decay_rate"),
                                          tuple(pw.get("no_decay_names", [])))
    elif constructor is None:
        groups = [{"params": [p for p in model.parameters() if p.requires_grad]}]
    else:
        raise NotImplementedError(f"b200swin.build_optimizers: constructor {constructor!r}")
    return FusedAdamW(groups, lr=lr, weight_decay=wd, **cfg)
