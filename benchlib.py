"""Shared pieces of bench.py: workload definitions (BASELINE.json configs), synthetic data, the
measured models, clock sampling and the roofline arithmetic.  Not part of the product package."""
from __future__ import annotations

import json
import math
import os
import statistics
import subprocess

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))

SWIN = {  # models/model.py:18-29
    "tiny": dict(embed_dim=96, num_heads=[3, 6, 12, 24], depths=[2, 2, 6, 2]),
    "base": dict(embed_dim=128, num_heads=[4, 8, 16, 32], depths=[2, 2, 18, 2]),
    "large": dict(embed_dim=192, num_heads=[6, 12, 24, 48], depths=[2, 2, 18, 2]),
}

# name -> description of one bench line (metric is always images/s = RGB frames through the encoder per second)
WORKLOADS = {
    # BASELINE config 2, the headline: SimMIM geometry (configs/config.yaml:56)
    "c2_ws12": dict(kind="train", size="base", windows=[12, 12, 12, 6], pre=[12, 12, 12, 6], img=(480, 480), pairs=24,
                    max_depth=10.0, invalid=0.05),
    "c2_ws24": dict(kind="train", size="base", windows=[24, 24, 24, 12], pre=[12, 12, 12, 6], img=(480, 480), pairs=24,
                    max_depth=10.0, invalid=0.05),
    # the reference's own default windows (configs/config.yaml:55)
    "c2_ws30": dict(kind="train", size="base", windows=[30, 30, 30, 15], pre=[12, 12, 12, 6], img=(480, 480), pairs=24,
                    max_depth=10.0, invalid=0.05),
    # BASELINE config 1 on the GPU (the reference arm runs the same thing on the CPU): Swin-T, default windows, 1 pair
    "c1_swinT": dict(kind="train", size="tiny", windows=[30, 30, 30, 15], pre=[30, 30, 30, 15], img=(480, 480), pairs=1,
                     max_depth=10.0, invalid=0.05),
    # KITTI- and VOID-shaped training steps (north_star: throughput on NYUv2-, KITTI- and VOID-shaped tensors)
    "kitti_train": dict(kind="train", size="base", windows=[12, 12, 12, 6], pre=[12, 12, 12, 6], img=(352, 1216),
                        pairs=12, max_depth=80.0, invalid=0.90),
    "void_train": dict(kind="train", size="base", windows=[12, 12, 12, 6], pre=[12, 12, 12, 6], img=(480, 640), pairs=16,
                       max_depth=10.0, invalid=0.30),
    # BASELINE config 4: Swin-L KITTI inference, CPB extrapolation 12 -> 24, batch sharded, no collective
    "c4_swinL_kitti_infer": dict(kind="infer", size="large", windows=[24, 24, 24, 12], pre=[12, 12, 12, 6],
                                 img=(352, 1216), frames=16, max_depth=80.0, invalid=0.90),
    # BASELINE config 5: attention half-block micro-benchmark vs the reference modules on the same GPU
    "c5_micro": dict(kind="micro"),
    # BASELINE config 3 (VOID 480 x 640, cnn_transformer path): the path contains no window attention -- only SiLog applies
    "c3_void_silog": dict(kind="silog", img=(480, 640), frames=64, max_depth=10.0, invalid=0.30),
    # ... and, with the global-attention kernels, its transformer encoder: 6 layers of global MHA + ReLU FFN
    # over the 30 x 40 = 1200 stride-16 tokens of a 480 x 640 frame, hidden 512 = 8 heads x 64 (models/model.py:73-95,
    # models/cnn_transformer.py:176-262), feed-forward 4096 (configs/config.yaml:63); batched inference
    "c3_void_encoder": dict(kind="mha", img=(480, 640), frames=32, hidden=512, heads=8, ff=4096, layers=6),
}


def encoder_cfg(w, drop_path_rate=0.3):
    s = SWIN[w["size"]]
    return dict(embed_dim=s["embed_dim"], depths=list(s["depths"]), num_heads=list(s["num_heads"]),
                window_size=list(w["windows"]), pretrain_window_size=list(w["pre"]),
                use_shift=[True, True, False, False], drop_path_rate=drop_path_rate)


def encoder_flops_per_frame(cfg, img):
    """Algorithmic forward FLOPs per frame: (GEMM part, attention-core part, attention-core bytes)."""
    gemm = attn = abytes = 0.0
    h, w = img[0] // 4, img[1] // 4
    for i, depth in enumerate(cfg["depths"]):
        C = cfg["embed_dim"] * 2 ** i
        ws = cfg["window_size"][i]
        T = h * w
        Tp = ((h + ws - 1) // ws * ws) * ((w + ws - 1) // ws * ws)
        N = ws * ws
        gemm += depth * (2 * T * C * 3 * C + 2 * T * C * C + 16 * T * C * C)
        attn += depth * (4 * Tp * N * C)
        abytes += depth * 8 * T * C
        if i < len(cfg["depths"]) - 1:
            h2, w2 = (h + 1) // 2, (w + 1) // 2
            gemm += 2 * (h2 * w2) * 4 * C * 2 * C
            h, w = h2, w2
    return gemm, attn, abytes


# ------------------------------------------------------------------------------------------ synthetic data
def make_batch(pairs, seed, img=(480, 480), max_depth=10.0, invalid=0.05, pin=False, pose=False):
    """RGB in [0,1) (ToTensor, no mean/std), depth GT U(0.5, max_depth) with a Bernoulli invalid mask set to 0,
    optionally random orthonormal R (9) and t (3) per direction."""
    g = torch.Generator().manual_seed(seed)
    H, W = img
    out = [torch.rand(pairs, 3, H, W, generator=g), torch.rand(pairs, 3, H, W, generator=g)]
    for _ in range(2):
        t = 0.5 + (max_depth - 0.5) * torch.rand(pairs, H, W, generator=g)
        out.append(torch.where(torch.rand(pairs, H, W, generator=g) < invalid, torch.zeros(()), t))
    if pose:
        for _ in range(2):
            q, _r = torch.linalg.qr(torch.randn(pairs, 3, 3, generator=g))
            out.append(q.reshape(pairs, 9).contiguous())
            out.append(torch.randn(pairs, 3, generator=g))
    if pin:
        out = [t.pin_memory() for t in out]
    return out


# ------------------------------------------------------------------------------------------ models
class EncoderReadout(torch.nn.Module):
    """Swin-V2 encoder (b200swin drop-in, or the reference's own when `reference_modules` is given) + pixel-shuffle depth
    read-out: Linear(8 * embed -> 32 * 32) per stride-32 token, sigmoid * max_depth (the reference decoders end the same
    way, models/decoder_v2.py:119).  The hot-path step of SURVEY.md section 8: everything in it runs in b200swin kernels."""

    def __init__(self, cfg, max_depth, reference_modules=None):
        super().__init__()
        self.max_depth = max_depth
        self.is_reference = reference_modules is not None
        if reference_modules is None:
            from b200swin.swin_transformer_v2 import SwinTransformerV2
            self.encoder = SwinTransformerV2(**cfg)
        else:
            import baseline
            self.encoder = baseline.quiet(reference_modules.swin.SwinTransformerV2, use_checkpoint=False, **cfg)
        self.encoder.init_weights(None)
        self.readout = torch.nn.Linear(cfg["embed_dim"] * 8, 32 * 32)
        torch.nn.init.normal_(self.readout.weight, std=0.02)
        torch.nn.init.zeros_(self.readout.bias)

    def forward(self, frame1, frame2=None):
        frames = frame1 if frame2 is None else torch.cat([frame1, frame2])     # models/model.py:116
        feat = self.encoder(frames)[0]                                         # [B, 8E, h, w] fp32 NCHW
        B, C, h, w = feat.shape
        tok = feat.permute(0, 2, 3, 1).reshape(B, h * w, C)
        if self.is_reference:
            d = torch.nn.functional.linear(tok, self.readout.weight, self.readout.bias)
        else:
            from b200swin import ops
            d = ops.linear(tok, self.readout.weight, self.readout.bias)
        d = d.view(B, h, w, 32, 32).permute(0, 1, 3, 2, 4).reshape(B, h * 32, w * 32)
        d = torch.sigmoid(d.float()) * self.max_depth
        return d if frame2 is None else d.chunk(2, dim=0)


def normalize_rot_vector_batched(rot_vector, method="svd"):
    """utils/util.py:5-17 without the per-sample Python loop / `.cuda()` allocations: the nearest rotation U V^T of every
    3x3 block.  "svd": one batched torch.linalg.svd (same algorithm, one cuSOLVER call instead of bs of them + bs host
    syncs).  "newton": the same polar factor by Higham's iteration X <- (X + X^-T) / 2 with the 3x3 inverse-transpose
    written out as cross products -- no cuSOLVER call, no host sync, CUDA-graph capturable."""
    bs = rot_vector.shape[0]
    m = rot_vector.reshape(bs, 3, 3).float()
    if method == "svd":
        u, _, vh = torch.linalg.svd(m, full_matrices=False)
        return (u @ vh).reshape(bs, 9).to(rot_vector.dtype)
    x = m
    for _ in range(12):
        r0, r1, r2 = x[:, 0], x[:, 1], x[:, 2]
        cof = torch.stack([torch.linalg.cross(r1, r2), torch.linalg.cross(r2, r0), torch.linalg.cross(r0, r1)], dim=1)
        det = (r0 * cof[:, 0]).sum(-1, keepdim=True).unsqueeze(-1)
        x = 0.5 * (x + cof / det)
    return x.reshape(bs, 9).to(rot_vector.dtype)


class FullDepthModel(torch.nn.Module):
    """the b200swin encoder behind the REFERENCE's decoder_v2 (depth + pose heads, cuDNN convs;
    models/decoder_v2.py, imported from the staged copy baseline/_ref) run under bf16 autocast in channels_last, with
    the batched rotation normalisation.  This is models/model.py's IDEDepth for model_scale 32 with the encoder swapped."""

    def __init__(self, cfg, max_depth, rot_method="svd"):
        super().__init__()
        import argparse
        import baseline
        from b200swin.swin_transformer_v2 import SwinTransformerV2
        ref = baseline.load()
        self.encoder = SwinTransformerV2(**cfg)
        self.encoder.init_weights(None)
        args = argparse.Namespace(max_depth=max_depth, num_deconv=3, num_filters=[32, 32, 32], deconv_kernels=[2, 2, 2],
                                  num_upscale_layer=2)                          # models/model.py:31-38
        E = cfg["embed_dim"]
        self.decoder = baseline.quiet(ref.decoder_v2.Decoder_v2, E * 8, E, args)
        self.decoder.init_weights()
        self.decoder = self.decoder.to(memory_format=torch.channels_last)
        ref.decoder_v2.normalize_rot_vector = lambda r: normalize_rot_vector_batched(r, rot_method)

    def forward(self, frame1, frame2):
        feats = self.encoder(torch.cat([frame1, frame2]))[0]
        f1, f2 = feats.contiguous(memory_format=torch.channels_last).chunk(2, dim=0)
        d1, r12, t12, d2, r21, t21 = self.decoder(f1, f2)
        return {"pred_d1": d1, "pred_d2": d2, "pred_r12": r12, "pred_r21": r21, "pred_t12": t12, "pred_t21": t21}


def full_step_loss(preds, batch, crit, lambda1=100.0, lambda2=100.0):
    """train.py:215-230 for decoder_v2."""
    _, _, d1, d2, R12, T12, R21, T21 = batch
    mse = torch.nn.functional.mse_loss
    loss_depth = (crit(preds["pred_d1"].squeeze(1).float(), d1) + crit(preds["pred_d2"].squeeze(1).float(), d2)) / 2
    loss_rot = (mse(preds["pred_r12"].float(), R12) + mse(preds["pred_r21"].float(), R21)) / 2
    loss_tr = (mse(preds["pred_t12"].float(), T12) + mse(preds["pred_t21"].float(), T21)) / 2
    return loss_depth + lambda1 * loss_rot + lambda2 * loss_tr


# ------------------------------------------------------------------------------------------ clocks / peaks
def clocks_sampler(path):
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    try:
        return subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                stdout=open(path, "w"), stderr=subprocess.DEVNULL)
    except Exception:
        return None


def clocks_summary(path, gpu_index):
    sm, mx, reasons = [], 0, set()
    try:
        for line in open(path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9 or f[0] != str(gpu_index):
                continue
            sm.append(float(f[1]))
            mx = max(mx, float(f[2]))
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
    except Exception:
        pass
    sm.sort()
    load = sm[len(sm) // 2:] if sm else []          # the first samples may precede the load: median over the upper half
    return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": mx or None,
            "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tf_sustained": d.get("bf16_tflops_sustained", 1404.8), "tf_burst": d.get("bf16_tflops", 1677.5),
                "hbm_gbs": d.get("hbm_gbs", 6449.4), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tf_sustained": 1400.0, "tf_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def committed_ncu(name):
    """A committed ncu --set full summary (profiles/<name>.json, produced by tools/ncu_summary.py), or None."""
    f = os.path.join(ROOT, "profiles", name + ".json")
    try:
        return json.load(open(f))
    except Exception:
        return None


def _mb(s):
    return float(str(s).split()[0]) * 1e6


def rooflines_from_events(events, steps, pk):
    """Per-kernel-family rooflines from the CUDA-event pairs recorded around every C-ABI call of the eager pass.
    GEMM: tensor roofline on 2*M*N*K (x3 in the fp32-accurate split mode).  Attention core: FLOPs 4*Tp*N*C forward /
    10*Tp*N*C backward, bytes 8*T*C / 16*T*C, MUFU floor = one exp2 per (row, key) at 16 per clock
    and SM.  LayerNorm: bytes moved.  Times are per step."""
    fam = {}

    def add(name, ms, flops=0.0, byts=0.0, exps=0.0):
        f = fam.setdefault(name, dict(ms=0.0, flops=0.0, bytes=0.0, exps=0.0, launches=0))
        f["ms"] += ms
        f["flops"] += flops
        f["bytes"] += byts
        f["exps"] += exps
        f["launches"] += 1

    for e0, e1, name, a in events:
        ms = e0.elapsed_time(e1)
        if name == "b200swin_gemm_bf16":
            fl = 2.0 * a[6] * a[7] * a[8] * (3 if a[1] else 1)
            add("gemm", ms, fl)
            kind = ("wgrad" if (a[2] and a[5]) else "dgrad" if a[5] else "fwd") + \
                   {0: "", 1: "+gelu", 2: "+qkvnorm", 3: "*gelu'", 4: "+residual"}[a[9]]
            add("gemm." + kind, ms, fl)
        elif name in ("b200swin_attn_fwd", "b200swin_attn_bwd"):
            off = 10 if name.endswith("fwd") else 17
            B, H, W, C, nH, ws = a[off:off + 6]
            Tp = B * ((H + ws - 1) // ws * ws) * ((W + ws - 1) // ws * ws)
            T, N = B * H * W, ws * ws
            if name.endswith("fwd"):
                add("attn_fwd", ms, 4.0 * Tp * N * C, 8.0 * T * C, float(Tp) * N * nH)
            else:
                # one exp2 per logit in the single-pass backward (windows 4 - 12), two in the KV-blocked two-pass backward
                add("attn_bwd", ms, 10.0 * Tp * N * C, 16.0 * T * C, float(Tp) * N * nH * (1.0 if ws in (4, 6, 7, 8, 12) else 2.0))
        elif name == "b200swin_ln_fwd":
            rows, C = a[9], a[10]
            es = 2 if a[12] == 1 else 4
            add("ln_fwd", ms, 0.0, rows * C * es * (3 if a[1] else 2))
        elif name == "b200swin_ln_bwd":
            rows, C = a[11], a[12]
            es = 2 if a[13] == 1 else 4
            add("ln_bwd", ms, 0.0, rows * C * es * 3)
        else:
            add(name.replace("b200swin_", ""), ms)
    out = {}
    sm_hz, n_sm = 1.9e9, 148
    for k, f in fam.items():
        ms = f["ms"] / steps
        d = {"ms_per_step": ms, "launches_per_step": f["launches"] / steps}
        if f["flops"]:
            d["tflops"] = f["flops"] / steps / (ms / 1e3) / 1e12
            d["frac_of_sustained_bf16_peak"] = d["tflops"] / pk["tf_sustained"]
        if f["bytes"]:
            d["algorithmic_gb_per_step"] = f["bytes"] / steps / 1e9
            d["gbs"] = f["bytes"] / steps / (ms / 1e3) / 1e9
            d["frac_of_hbm_peak"] = d["gbs"] / pk["hbm_gbs"]
        if f["exps"]:
            floor_ms = f["exps"] / steps / (16.0 * n_sm * sm_hz) * 1e3
            d["mufu_floor_ms_per_step"] = floor_ms
            d["frac_of_mufu_floor"] = floor_ms / ms if ms > 0 else None
        out[k] = d
    return out


def attn_evidence():
    """Tensor-pipe utilisation of the window-attention kernels from the committed ncu captures (profiles/*.json)."""
    out = {}
    for key, name in (("fwd", "r02_ncu_attn_mma_fwd_ws12"), ("bwd", "r02_ncu_attn_mma_bwd_ws12"),
                      ("fwd_single_tile_tcgen05", "r02_ncu_attn_fwd_ws12"), ("bwd_single_tile_tcgen05", "r02_ncu_attn_bwd_ws12"),
                      ("fwd_r01", "r01_ncu_attn"),
                      ("bwd_r01", "r01_ncu_attn_bwd"), ("fwd_kv_blocked_ws24", "r02_ncu_attn_flash_fwd_ws24"),
                      ("bwd_kv_blocked_ws24", "r02_ncu_attn_flash_bwd_ws24")):
        d = committed_ncu(name)
        if d is None:
            continue
        d = d if isinstance(d, dict) else d[0]
        out[key] = {"tensor_pipe_pct_active": d.get("tensor_pipe_pct_active"), "xu_pipe_pct": d.get("xu_pipe_pct"),
                    "source": f"profiles/{name}.json"}
    return out or None
