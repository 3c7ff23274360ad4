#!/usr/bin/env python
"""bench.py - train images/sec of the Swin-V2 hot path on B200 (BASELINE.json metric, configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--pairs P] [--dtype bf16|fp32]

A "step" is one pass of the hot path over one batch of synthetic NYUv2-shaped input:
Swin-V2-B encoder (embed 128, depths [2,2,18,2], heads [4,8,16,32], windows [12,12,12,6], 24 blocks) forward
and backward on 24 frame pairs (= 48 RGB frames of 480x480) per GPU under bf16 autocast with fp32 master
weights, a pixel-shuffle depth read-out (one more b200swin GEMM; the reference's decoder_v2 is a cuDNN conv
stack OUTSIDE the hot path and is not part of the step), the SiLog loss on both frames (forward + backward),
the NCCL gradient all-reduce (N > 1) and a fused AdamW step.  images/sec = frames through the encoder / s.

`value` times K steps with inputs resident in HBM (CUDA events, barrier + synchronize on both sides, max over
ranks).  `e2e` repeats the measurement through the public module API with HOST inputs: every step copies the
batch from pinned host memory and reads the loss back.  `roofline` is the aggregate of every tcgen05 GEMM
launch inside the timed region (CUDA events around each launch) against the measured sustained bf16 peak.
`cpu_baseline` / `--impl reference`: the CPU oracle (oracle/swin_ref.py, a restatement of the reference's PyTorch
path pinned by golden vectors; the reference itself is not on the GPU box) on a bounded sample of the same
workload with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG = dict(embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32], window_size=[12, 12, 12, 6],
           pretrain_window_size=[12, 12, 12, 6], use_shift=[True, True, False, False], drop_path_rate=0.3)
IMG = 480
MAX_DEPTH = 10.0
WORKLOAD = ("swin_v2_base_480x480_ws12_24pairs_per_gpu_train_step(encoder+pixelshuffle_readout+silog_x2+adamw; "
            "decoder_v2 outside hot path, not included)")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=24, help="frame pairs per GPU (BASELINE: 24)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--attn", default="auto", choices=["auto", "simt", "tc"])
    ap.add_argument("--cpu-sample-pairs", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="print per-entry-point CUDA-event times to stderr")
    ap.add_argument("--no-graph", action="store_true", help="time eager steps instead of CUDA-graph replays")
    return ap.parse_args()


def encoder_flops_per_frame():
    """Algorithmic forward FLOPs per frame (SURVEY.md section 8d formula), GEMM part and attention-core part."""
    gemm = attn = 0.0
    T = (IMG // 4) ** 2
    side = IMG // 4
    for i, depth in enumerate(CFG["depths"]):
        C = CFG["embed_dim"] * 2 ** i
        ws = CFG["window_size"][i]
        sp = (side + ws - 1) // ws * ws
        Tp, N = sp * sp, ws * ws
        gemm += depth * (2 * T * C * 3 * C + 2 * T * C * C + 16 * T * C * C)
        attn += depth * (4 * Tp * N * C)
        if i < len(CFG["depths"]) - 1:
            gemm += 2 * (T // 4) * 4 * C * 2 * C
            side = (side + 1) // 2
            T = side * side
    return gemm, attn


# ------------------------------------------------------------------------------------------ synthetic data
def make_batch(pairs, seed, device="cpu", pin=False):
    g = torch.Generator().manual_seed(seed)
    img1 = torch.rand(pairs, 3, IMG, IMG, generator=g)
    img2 = torch.rand(pairs, 3, IMG, IMG, generator=g)
    d = []
    for _ in range(2):
        t = 0.5 + (MAX_DEPTH - 0.5) * torch.rand(pairs, IMG, IMG, generator=g)
        t = torch.where(torch.rand(pairs, IMG, IMG, generator=g) < 0.05, torch.zeros(()), t)
        d.append(t)
    out = [img1, img2, d[0], d[1]]
    if pin:
        out = [t.pin_memory() for t in out]
    if device != "cpu":
        out = [t.to(device) for t in out]
    return out


# ------------------------------------------------------------------------------------------ b200 arm
class DepthModel(torch.nn.Module):
    """Swin-V2-B encoder (b200swin drop-in) + pixel-shuffle depth read-out: Linear(1024 -> 32*32) per stride-32
    token through the b200swin GEMM, sigmoid * max_depth (the reference decoders end the same way,
    models/decoder_v2.py:119)."""

    def __init__(self):
        super().__init__()
        from b200swin.swin_transformer_v2 import SwinTransformerV2
        self.encoder = SwinTransformerV2(**CFG)
        self.encoder.init_weights(None)
        self.readout = torch.nn.Linear(CFG["embed_dim"] * 8, 32 * 32)
        torch.nn.init.normal_(self.readout.weight, std=0.02)
        torch.nn.init.zeros_(self.readout.bias)

    def forward(self, frame1, frame2):
        from b200swin import ops
        frames = torch.cat([frame1, frame2])                     # models/model.py:116
        feat = self.encoder(frames)[0]                           # [2P, 1024, 15, 15] fp32 NCHW
        B, C, h, w = feat.shape
        tok = feat.permute(0, 2, 3, 1).reshape(B, h * w, C)
        d = ops.linear(tok, self.readout.weight, self.readout.bias)        # [2P, 225, 1024]
        d = d.view(B, h, w, 32, 32).permute(0, 1, 3, 2, 4).reshape(B, h * 32, w * 32)
        d = torch.sigmoid(d.float()) * MAX_DEPTH
        return d.chunk(2, dim=0)


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
    # the first samples may precede the load: median over the upper half
    sm.sort()
    load = sm[len(sm) // 2:] if sm else []
    return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": mx or None,
            "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1404.8), d.get("hbm_gbs", 6449.4), "measured"
    return 1400.0, 6650.0, "fallback"


def gemm_traffic():
    """DRAM bytes of one representative launch of the dominant kernel, from the committed ncu --set full capture."""
    f = os.path.join(ROOT, "profiles", "r01_ncu_gemm_plain_st2.json")
    try:
        d = json.load(open(f))
        rd, wr = float(d["dram_read"].split()[0]), float(d["dram_write"].split()[0])      # Mbyte
        M, N, K = 43200, 2048, 512
        return {"launch": "gemm_tc_kernel M=43200 N=2048 K=512 (stage-2 fc1 forward, plain epilogue)",
                "dram_bytes": (rd + wr) * 1e6, "algorithmic_bytes": 2.0 * (M * K + N * K + M * N),
                "source": "profiles/r01_ncu_gemm_plain_st2.json"}
    except Exception:
        return None


def run_b200(args):
    # Everything -- eager warm-up, the eager (roofline) pass, the graph capture and its replays -- runs on ONE
    # non-default stream: autograd binds each parameter's gradient accumulation to the stream of its first backward,
    # and a capture on any other stream would need cross-stream syncs that invalidate it.
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        _run_b200(args, stream)


def _run_b200(args, stream):
    import torch.distributed as dist
    from b200swin import SiLogLoss, _lib, ops
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops.ATTN_IMPL["mode"] = args.attn
    torch.manual_seed(0)                                   # identical weights on every rank
    model = DepthModel().to(dev)
    model.train()
    crit = SiLogLoss()
    params = [p for p in model.parameters() if p.requires_grad]
    # one flat fp32 gradient buffer for the data-parallel exchange: ONE NCCL all-reduce (average) over NVLink -- the only
    # collective of the path.  Autograd ASSIGNS fresh gradients (p.grad = None before the
    # backward); with N > 1 they are packed into the flat buffer by one multi-tensor copy.  (Pre-set .grad views made
    # autograd launch one accumulation kernel per parameter: ~330 launch-bound adds per step.)
    flat_grad = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
    flat_views, off = [], 0
    for p in params:
        flat_views.append(flat_grad[off:off + p.numel()].view_as(p))
        off += p.numel()
    opt = torch.optim.AdamW(params, lr=5e-4, weight_decay=0.05, fused=True, capturable=True)
    P = args.pairs
    host = make_batch(P, 1234 + rank, pin=True)
    statics = [[t.to(dev) for t in host] for _ in range(2)]   # two device-resident input sets (graph inputs);
    static = statics[0]                                       # the second lets the e2e loop prefetch the next batch
    use_amp = args.dtype == "bf16"

    def fwd_bwd(batch):
        img1, img2, d1, d2 = batch
        for p in params:
            p.grad = None
        with torch.autocast("cuda", torch.bfloat16, enabled=use_amp):
            p1, p2 = model(img1, img2)
        loss = (crit(p1, d1) + crit(p2, d2)) / 2                 # train.py:215-217
        loss.backward()
        if world > 1:
            pairs = [(v, p.grad) for v, p in zip(flat_views, params) if p.grad is not None]
            torch._foreach_copy_([v for v, _ in pairs], [g for _, g in pairs])
        return loss

    def finish():
        if world > 1:
            dist.all_reduce(flat_grad, op=dist.ReduceOp.AVG)
            for p, v in zip(params, flat_views):                 # the optimizer reads the averaged gradients
                if p.grad is not None:
                    p.grad = v
        opt.step()

    def step_eager(batch):
        loss = fwd_bwd(batch)
        finish()
        return loss

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    warm = max(args.warmup, 3)
    for _ in range(warm):
        loss = step_eager(static)
    torch.cuda.synchronize(dev)
    assert torch.isfinite(loss).item(), "non-finite loss in warm-up"

    # ---- eager pass: launch counts and live per-launch timing of the GEMM (CUDA events on the launching stream)
    _lib.reset_counters()
    _lib.TIMING.update(name="b200swin_gemm_bf16", events=[],
                       work=lambda a: 2.0 * a[6] * a[7] * a[8] * (3 if a[1] else 1))
    ms_eager = timed(lambda: step_eager(static), args.steps)
    _lib.TIMING["name"] = None
    launches = _lib.COUNTERS["launches"]
    calls = dict(_lib.COUNTERS["calls"])
    gemm_ms = sum(e0.elapsed_time(e1) for e0, e1, _, _ in _lib.TIMING["events"])
    gemm_flops = sum(w for _, _, w, _ in _lib.TIMING["events"])
    n_gemm = len(_lib.TIMING["events"])

    if args.breakdown:
        # diagnostic: CUDA-event time of every C-ABI entry point over one more step (not part of the JSON contract)
        _lib.reset_counters()
        _lib.TIMING.update(name="*", events=[], detail=True)
        ms1 = timed(lambda: step_eager(static), 1)
        _lib.TIMING["name"] = None
        _lib.TIMING["detail"] = False
        agg, gem = {}, {}
        for e0, e1, _, nm in _lib.TIMING["events"]:
            t = e0.elapsed_time(e1)
            if nm.startswith("gemm "):
                c = gem.setdefault(nm, [0, 0.0])
                c[0] += 1
                c[1] += t
                nm = "b200swin_gemm_bf16"
            agg[nm] = agg.get(nm, 0.0) + t
        if rank == 0:
            for k, (n, t) in sorted(gem.items(), key=lambda kv: -kv[1][1])[:40]:
                print(f"  {t:7.3f} ms  x{n:3d}  {k}", file=sys.stderr)
        if rank == 0:
            tot = sum(agg.values())
            print("breakdown (ms/step):", json.dumps({k: round(v, 2) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])}),
                  f"sum={tot:.1f} step={ms1:.1f} other(torch ops, gaps)={ms1 - tot:.1f}", file=sys.stderr)

    # ---- the product path: the whole step captured once in a CUDA graph (forward, backward and -- on one GPU -- the
    # optimizer), replayed per step; with N > 1 the all-reduce and the optimizer follow the replay eagerly.  (Capturing
    # the NCCL all-reduce inside the graph hung on this stack -- torch 2.11 / NCCL 2.28, two replays in flight -- and
    # stays opt-in: B200SWIN_BENCH_CAPTURE_ALLREDUCE=1.)
    graph = None
    graphs, g_losses = [], []
    full_capture = world == 1 or os.environ.get("B200SWIN_BENCH_CAPTURE_ALLREDUCE", "0") == "1"
    if not args.no_graph:
        def capture_all(with_tail):
            gs, ls = [], []
            for sset in statics:                           # one graph per input set, sharing one memory pool
                torch.cuda.synchronize(dev)
                ops._weight_cache.clear()                  # the bf16 staging of every weight must be part of the capture
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=stream, pool=(gs[0].pool() if gs else None)):
                    gl = fwd_bwd(sset)
                    if with_tail:
                        finish()
                gs.append(g)
                ls.append(gl)
            return gs, ls
        try:
            graphs, g_losses = capture_all(full_capture)
        except Exception as e:                             # noqa: BLE001 -- any capture failure of the collective
            if world == 1 or not full_capture:
                raise
            if rank == 0:
                print(f"[bench] all-reduce not capturable here ({type(e).__name__}); eager tail", file=sys.stderr)
            torch.cuda.synchronize(dev)
            full_capture = False
            graphs, g_losses = capture_all(False)
        graph = graphs[0]

    def step_set(k):
        if graph is None:
            return step_eager(statics[k])
        graphs[k].replay()
        if not full_capture:
            finish()
        return g_losses[k]

    step = lambda: step_set(0)
    for _ in range(2):
        step()
    clk_path = os.path.join(tempfile.gettempdir(), f"b200swin_clocks_{rank}.csv")
    sampler = clocks_sampler(clk_path) if rank == 0 else None
    ms = timed(step, args.steps)

    # ---- end to end through the public API with host buffers: per step H2D copy of the batch, the step, loss read-back
    # The copy of batch i+1 (pinned host -> the other input set, on a copy stream) overlaps the compute of batch i, as
    # a prefetching data loader would; every step still pays its own H2D copy and its own loss read-back inside the
    # timed region, and the first batch's copy is not hidden.
    copy_stream = torch.cuda.Stream()
    ev_copied = [torch.cuda.Event(), torch.cuda.Event()]
    ev_done = [torch.cuda.Event(), torch.cuda.Event()]

    def h2d(k, first_use):
        with torch.cuda.stream(copy_stream):
            if not first_use:
                copy_stream.wait_event(ev_done[k])         # the step that last read this input set has finished
            for d, h in zip(statics[k], host):
                d.copy_(h, non_blocking=True)
            ev_copied[k].record(copy_stream)

    # The loss of step i is copied to pinned host memory asynchronously and READ while step i + 1 runs (a training loop
    # that logs its loss does the same): every step's result still reaches the host inside the timed region, but the
    # device never idles waiting for the host to look at a number.
    loss_host = torch.empty(2, dtype=torch.float32).pin_memory()
    ev_loss = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_loop(n_steps):
        last = None
        h2d(0, True)
        for i in range(n_steps):
            k = i & 1
            if i + 1 < n_steps:
                h2d(k ^ 1, i == 0)
            stream.wait_event(ev_copied[k])
            loss_i = step_set(k)
            ev_done[k].record(stream)
            loss_host[k:k + 1].copy_(loss_i.detach().reshape(1), non_blocking=True)   # device -> host read of the result
            ev_loss[k].record(stream)
            if i > 0:
                ev_loss[k ^ 1].synchronize()
                last = float(loss_host[k ^ 1])
        ev_loss[(n_steps - 1) & 1].synchronize()
        last = float(loss_host[(n_steps - 1) & 1])
        return last
    torch.cuda.synchronize(dev)
    e2e_loop(2)
    torch.cuda.synchronize(dev)
    ms_e2e = timed(lambda: e2e_loop(args.steps), 1)
    if sampler is not None:
        sampler.terminate()
    frames = 2 * P * world
    value = frames * args.steps / (ms / 1e3)
    e2e = frames * args.steps / (ms_e2e / 1e3)

    if rank == 0:
        tf_peak, hbm_peak, how = peaks()
        gemm_f, attn_f = encoder_flops_per_frame()
        achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
        line = {
            "metric": "train images/sec, Swin-V2-B depth @480^2 (hot path: encoder fwd+bwd + SiLog + AdamW)",
            "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if use_amp else "f32(split-bf16 x3)", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "pairs_per_gpu": P, "frames_per_gpu": 2 * P, "windows": CFG["window_size"],
                       "attn_impl": args.attn, "parallelism": f"dp{world}",
                       "execution": ("cuda_graph_replay" + ("" if full_capture else "+eager_allreduce_adamw")) if graph is not None else "eager",
                       "l2": "inputs+activations >> 126 MB L2 (48x3x480x480 fp32 = 133 MB images alone)"},
            "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host),
                    "d2h_bytes_per_step": 4},
            "gpu_launches": launches,
            "launch_calls": calls,
            "eager_ms_per_step": ms_eager / args.steps,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s",
                         "frac": achieved / tf_peak, "traffic": gemm_traffic(), "kernel": "gemm_tc_kernel (all launches)",
                         "launches": n_gemm, "kernel_ms_per_step": gemm_ms / args.steps, "peak_source": how,
                         "timed_in": "eager pass of the same K steps (CUDA events around every launch)"},
            "model_flops": {"encoder_fwd_gflop_per_frame": (gemm_f + attn_f) / 1e9,
                            "step_tflops_achieved": 3 * (gemm_f + attn_f) * 2 * P / (ms / args.steps / 1e3) / 1e12},
            "clocks": clocks_summary(clk_path, local),
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_reference(args.cpu_sample_pairs, steps=1, warmup=0)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ reference arm (CPU)
def cpu_reference(pairs, steps, warmup):
    """The CPU oracle (port of the reference's PyTorch path) on a bounded sample: `pairs` frame pairs of the same
    workload, forward + backward + SiLog, fp32, all host threads."""
    from oracle import silog_ref, swin_ref
    from b200swin.swin_transformer_v2 import SwinTransformerV2
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    enc = SwinTransformerV2(**CFG)          # parameter container only (never executed on the CPU)
    enc.init_weights(None)
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and "relative_coords" not in k)
          for k, v in enc.state_dict().items()}
    readout_w = (torch.randn(1024, CFG["embed_dim"] * 8) * 0.02).requires_grad_(True)
    readout_b = torch.zeros(1024, requires_grad=True)
    img1, img2, d1, d2 = make_batch(pairs, 1234)

    leaves = [v for v in sd.values() if v.requires_grad] + [readout_w, readout_b]
    opt = torch.optim.AdamW(leaves, lr=5e-4, weight_decay=0.05)

    def one():
        opt.zero_grad(set_to_none=True)
        feat = swin_ref.swin_v2(torch.cat([img1, img2]), sd, CFG["embed_dim"], CFG["depths"], CFG["num_heads"],
                                CFG["window_size"], CFG["use_shift"], (3,))[0]
        B, C, h, w = feat.shape
        d = torch.nn.functional.linear(feat.permute(0, 2, 3, 1).reshape(B, h * w, C), readout_w, readout_b)
        d = torch.sigmoid(d.view(B, h, w, 32, 32).permute(0, 1, 3, 2, 4).reshape(B, h * 32, w * 32)) * MAX_DEPTH
        p1, p2 = d.chunk(2)
        loss = (silog_ref.silog_torch(p1, d1) + silog_ref.silog_torch(p2, d2)) / 2
        loss.backward()
        opt.step()
        return loss.item()

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return {"value": 2 * pairs * steps / dt, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"{steps} step(s) of {pairs} pair(s) ({2 * pairs} frames) 480x480, Swin-V2-B ws12 fwd+bwd+SiLog+AdamW, "
                      f"fp32 torch CPU, {dt:.1f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pairs = args.cpu_sample_pairs
    steps, warmup = max(1, min(args.steps, 2)), min(args.warmup, 1)
    t0 = time.perf_counter()
    base = cpu_reference(pairs, steps, warmup)
    line = {
        "impl": "reference",
        "metric": "train images/sec, Swin-V2-B depth @480^2 (hot path: encoder fwd+bwd + SiLog + AdamW)",
        "value": base["value"], "unit": "images/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": (time.perf_counter() - t0) * 1e3 / (steps + warmup), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_gpu": args.pairs, "frames_per_gpu": 2 * args.pairs,
                   "windows": CFG["window_size"], "parallelism": f"dp{args.gpus}", "execution": "cpu_reference_port",
                   "sample": f"each step = {pairs} pair(s) of the workload (bounded sample of the 24-pair batch)",
                   "note": "CPU oracle port of the reference PyTorch path on the host cores (the Python reference itself "
                           "cannot travel to the GPU box); requested steps/warmup "
                           f"({args.steps}/{args.warmup}) clamped to ({steps}/{warmup}) to bound the run"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
