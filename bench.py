#!/usr/bin/env python
"""bench.py - throughput of the Swin-V2 hot path on B200 (BASELINE.json metric: train images/sec, Swin-V2-B @480^2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME] [--dtype bf16|fp32]

Default workload `c2_ws12` = BASELINE configs[1]: one "step" is the hot path over one batch of synthetic NYUv2-shaped
input -- Swin-V2-B encoder (embed 128, depths [2,2,18,2], heads [4,8,16,32], windows [12,12,12,6]) forward + backward on
24 frame pairs (48 RGB frames of 480x480) per GPU under bf16 autocast with fp32 master weights, a pixel-shuffle depth
read-out (one more b200swin GEMM), the SiLog loss on both frames, the gradient all-reduce (N > 1; b200swin.DataParallel,
flat buffer, bucketed NCCL) and the fused multi-tensor AdamW on the reference's layer-decay groups.  images/sec = frames
through the encoder / s.  The JSON line also carries:

  e2e                   the same step through the public module API with HOST batches (pinned H2D copy + loss read-back
                        inside the timed region);
  roofline              every tcgen05 GEMM launch of the timed eager pass against the measured sustained bf16 peak;
  roofline_attn         the attention core, forward and backward separately (FLOPs, bytes, MUFU floor; live CUDA events)
                        + tensor-pipe % from the committed ncu captures;
  full_step             encoder + the REFERENCE's decoder_v2 (bf16 channels_last, batched rotation normalisation) +
                        SiLog + pose losses + all-reduce + AdamW: the whole training step of train.py (N = 1 and in SCALE);
  reference_cuda_eager  the reference's own PyTorch modules (baseline/_ref) on the same GPU, same step, eager;
  cpu_baseline          the reference's own modules on the host cores, bounded sample.

Other workloads (one JSON line each): c2_ws24, c2_ws30 (the reference's default windows), c1_swinT, kitti_train,
void_train, c4_swinL_kitti_infer (inference, batch-sharded, no collective), c5_micro (attention half-block micro-bench vs
the reference modules on the same GPU), c3_void_silog, c3_void_encoder (config 3's global-attention encoder layers).  `--impl reference` runs the UNMODIFIED reference (staged copy,
stock code path) on the host cores for the same metric and config.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import benchlib as BL  # noqa: E402

METRIC = "train images/sec, Swin-V2-B depth @480^2 (hot path: encoder fwd+bwd + SiLog + AdamW)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2_ws12", choices=sorted(BL.WORKLOADS))
    ap.add_argument("--pairs", type=int, default=0, help="frame pairs per GPU (default: the workload's, BASELINE: 24)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--attn", default="auto", choices=["auto", "simt", "tc", "flash", "ws", "mma"])
    ap.add_argument("--cpu-sample-pairs", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip full_step / reference_cuda_eager / cpu_baseline")
    ap.add_argument("--breakdown", action="store_true", help="print per-entry-point CUDA-event times to stderr")
    ap.add_argument("--no-graph", action="store_true", help="time eager steps instead of CUDA-graph replays")
    ap.add_argument("--buckets", type=int, default=4, help="all-reduce buckets of the flat gradient buffer")
    return ap.parse_args()


def env():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


# ================================================================================================ training step
class TrainRun:
    """One data-parallel training-step measurement: builds the model, the product DataParallel wrapper and FusedAdamW
    on shared flat buffers, captures forward + backward (+ optimizer on one GPU) in a CUDA graph, times replays."""

    def __init__(self, args, w, model, batch_host, loss_fn, depths, stream):
        import torch.distributed as dist
        from b200swin.data_parallel import DataParallel
        from b200swin.optim import FusedAdamW, layer_decay_param_groups
        self.args, self.w, self.stream = args, w, stream
        self.world, self.rank, self.local = env()
        self.dev = torch.device("cuda", self.local)
        self.dist = dist
        self.model = model.to(self.dev).train()
        self.loss_fn = loss_fn
        # the reference's optimizer: AdamW, lr 5e-4, wd 0.05, layer decay 0.9 (configs/config.yaml:16-19, train.py:113-115)
        groups = layer_decay_param_groups(self.model, 5e-4, 0.05, depths, 0.9)
        self.opt = FusedAdamW(groups, lr=5e-4, betas=(0.9, 0.999), weight_decay=0.05)
        self.dp = DataParallel(self.model, buckets=args.buckets, flat=self.opt.flat)
        self.opt.grads_packed = True                    # dp.reduce_gradients() fills the flat gradient buffer
        self.params = self.opt.flat.params
        self.host = batch_host
        self.statics = [[t.to(self.dev) for t in batch_host] for _ in range(2)]
        self.use_amp = args.dtype == "bf16"
        self.graphs, self.g_losses, self.full_capture = [], [], False

    def fwd_bwd(self, batch):
        for p in self.params:
            p.grad = None
        with torch.autocast("cuda", torch.bfloat16, enabled=self.use_amp):
            out = self.model(batch[0], batch[1])
        loss = self.loss_fn(out, batch)
        loss.backward()
        self.opt.flat.pack_grads()                      # one multi-tensor copy into the flat gradient buffer
        return loss

    def finish(self):
        if self.world > 1:
            self.dp.overlap = True                      # gradients are already packed: exchange only
            for lo, hi, _, _ in self.dp._bounds:
                self.dp._launch_bucket(lo, hi)
            self.dp._join()
            self.dp.overlap = False
        self.opt.step()

    def step_eager(self, batch):
        loss = self.fwd_bwd(batch)
        self.finish()
        return loss

    def barrier(self):
        if self.world > 1:
            self.dist.barrier(device_ids=[self.local])
        torch.cuda.synchronize(self.dev)

    def timed(self, fn, steps):
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    def capture(self):
        """forward + backward + gradient packing (+ the optimizer when there is no collective) as ONE graph per input set."""
        self.full_capture = self.world == 1
        gs, ls = [], []
        for sset in self.statics:
            torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self.stream, pool=(gs[0].pool() if gs else None)):
                gl = self.fwd_bwd(sset)
                if self.full_capture:
                    self.finish()
            gs.append(g)
            ls.append(gl)
        self.graphs, self.g_losses = gs, ls

    def step_set(self, k):
        if not self.graphs:
            return self.step_eager(self.statics[k])
        self.graphs[k].replay()
        if not self.full_capture:
            self.finish()
        return self.g_losses[k]

    def e2e(self, n_steps):
        """Public-API step with HOST batches: the copy of batch i+1 (pinned host -> the other input set, copy stream)
        overlaps the compute of batch i like a prefetching loader; every step pays its own H2D copy and its own loss
        read-back inside the timed region (the loss of step i is read while step i+1 runs)."""
        copy_stream = torch.cuda.Stream()
        ev_copied = [torch.cuda.Event(), torch.cuda.Event()]
        ev_done = [torch.cuda.Event(), torch.cuda.Event()]
        loss_host = torch.empty(2, dtype=torch.float32).pin_memory()
        ev_loss = [torch.cuda.Event(), torch.cuda.Event()]
        stream = self.stream

        def h2d(k, first_use):
            with torch.cuda.stream(copy_stream):
                if not first_use:
                    copy_stream.wait_event(ev_done[k])
                for d, h in zip(self.statics[k], self.host):
                    d.copy_(h, non_blocking=True)
                ev_copied[k].record(copy_stream)

        def loop(n):
            h2d(0, True)
            for i in range(n):
                k = i & 1
                if i + 1 < n:
                    h2d(k ^ 1, i == 0)
                stream.wait_event(ev_copied[k])
                loss_i = self.step_set(k)
                ev_done[k].record(stream)
                loss_host[k:k + 1].copy_(loss_i.detach().reshape(1), non_blocking=True)
                ev_loss[k].record(stream)
                if i > 0:
                    ev_loss[k ^ 1].synchronize()
                    float(loss_host[k ^ 1])
            ev_loss[(n - 1) & 1].synchronize()
            return float(loss_host[(n - 1) & 1])

        torch.cuda.synchronize(self.dev)
        loop(2)
        torch.cuda.synchronize(self.dev)
        return self.timed(lambda: loop(n_steps), 1)


def run_train(args, w, name):
    world, rank, local = env()
    torch.cuda.set_device(local)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        return _run_train(args, w, name, stream)


def _run_train(args, w, name, stream):
    import torch.distributed as dist
    from b200swin import SiLogLoss, _lib, ops
    world, rank, local = env()
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    ops.ATTN_IMPL["mode"] = ops.ATTN_IMPL["bwd_mode"] = args.attn
    P = args.pairs or w["pairs"]
    cfg = BL.encoder_cfg(w)
    crit = SiLogLoss()
    torch.manual_seed(0)                                   # identical weights on every rank
    model = BL.EncoderReadout(cfg, w["max_depth"])
    host = BL.make_batch(P, 1234 + rank, w["img"], w["max_depth"], w["invalid"], pin=True)

    def loss_fn(out, batch):
        return (crit(out[0], batch[2]) + crit(out[1], batch[3])) / 2                 # train.py:215-217

    run = TrainRun(args, w, model, host, loss_fn, cfg["depths"], stream)
    warm = max(args.warmup, 3)
    for _ in range(warm):
        loss = run.step_eager(run.statics[0])
    torch.cuda.synchronize(dev)
    assert torch.isfinite(loss).item(), "non-finite loss in warm-up"

    # ---- eager pass: launch counts + live CUDA-event timing of every C-ABI call (the rooflines)
    _lib.reset_counters()
    _lib.TIMING.update(name="*", events=[])
    ms_eager = run.timed(lambda: run.step_eager(run.statics[0]), args.steps)
    _lib.TIMING["name"] = None
    launches = _lib.COUNTERS["launches"]
    calls = dict(_lib.COUNTERS["calls"])
    pk = BL.peaks()
    fam = BL.rooflines_from_events(_lib.TIMING["events"], args.steps, pk)
    _lib.TIMING["events"] = []
    if args.breakdown and rank == 0:
        print("breakdown (ms/step): " + json.dumps({k: round(v["ms_per_step"], 3) for k, v in
                                                    sorted(fam.items(), key=lambda kv: -kv[1]["ms_per_step"])}),
              f"eager step {ms_eager / args.steps:.1f} ms", file=sys.stderr)

    if not args.no_graph:
        run.capture()
    for _ in range(2):
        run.step_set(0)
    clk_path = os.path.join(tempfile.gettempdir(), f"b200swin_clocks_{rank}.csv")
    sampler = BL.clocks_sampler(clk_path) if rank == 0 else None
    ms = run.timed(lambda: run.step_set(0), args.steps)
    ms_e2e = run.e2e(args.steps)
    if sampler is not None:
        sampler.terminate()
    frames = 2 * P * world
    value = frames * args.steps / (ms / 1e3)
    e2e = frames * args.steps / (ms_e2e / 1e3)

    line = None
    if rank == 0:
        gemm_f, attn_f, _ = BL.encoder_flops_per_frame(cfg, w["img"])
        g = fam.get("gemm", {})
        tr = BL.committed_ncu("r02_ncu_gemm_plain_st2")
        traffic = None
        if tr:
            M, N, K = 43200, 2048, 512
            traffic = {"launch": "gemm_tc_kernel M=43200 N=2048 K=512 (stage-2 fc1 forward)",
                       "dram_bytes": BL._mb(tr["dram_read"]) + BL._mb(tr["dram_write"]),
                       "algorithmic_bytes": 2.0 * (M * K + N * K + M * N), "source": "profiles/r02_ncu_gemm_plain_st2.json"}
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if run.use_amp else "f32(split-bf16 x3)", "data": "synthetic",
            "config": {"workload": f"{name}: swin_v2_{w['size']}_{w['img'][0]}x{w['img'][1]}_windows{w['windows']}_"
                                   f"{P}pairs_per_gpu_train_step(encoder+pixelshuffle_readout+silog_x2+layer_decay_adamw; "
                                   "decoder_v2 outside the hot path: see full_step)",
                       "pairs_per_gpu": P, "frames_per_gpu": 2 * P, "windows": w["windows"], "attn_impl": args.attn,
                       "parallelism": f"dp{world}", "allreduce_buckets": len(run.dp._bounds),
                       "execution": ("cuda_graph_replay" + ("" if run.full_capture else "+bucketed_allreduce+fused_adamw"))
                       if run.graphs else "eager",
                       "l2": "inputs + activations >> 126 MB L2 (the images alone are 133 MB)"},
            "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host),
                    "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "launch_calls": calls, "eager_ms_per_step": ms_eager / args.steps,
            "roofline": {"bound": "tensor", "achieved": g.get("tflops"), "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                         "frac": g.get("frac_of_sustained_bf16_peak"), "traffic": traffic,
                         "kernel": "gemm_tc_kernel (all launches)", "launches": int(g.get("launches_per_step", 0) * args.steps),
                         "kernel_ms_per_step": g.get("ms_per_step"), "peak_source": pk["source"],
                         "timed_in": "eager pass of the same K steps (CUDA events around every launch)"},
            "roofline_attn": {"fwd": fam.get("attn_fwd"), "bwd": fam.get("attn_bwd"), "ncu": BL.attn_evidence(),
                              "note": "FLOPs 4*Tp*N*C fwd / 10*Tp*N*C bwd, bytes 8*T*C / 16*T*C (SURVEY 8d); MUFU floor = one "
                                      "exp2 per (row,key) (two in the KV-blocked / recomputing backward) at 16/clk/SM"},
            "roofline_ln": {"fwd": fam.get("ln_fwd"), "bwd": fam.get("ln_bwd")},
            "kernel_ms_per_step": {k: round(v["ms_per_step"], 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms_per_step"])},
            "model_flops": {"encoder_fwd_gflop_per_frame": (gemm_f + attn_f) / 1e9,
                            "step_tflops_achieved": 3 * (gemm_f + attn_f) * 2 * P / (ms / args.steps / 1e3) / 1e12},
            "clocks": BL.clocks_summary(clk_path, local),
        }
        ev = BL.attn_evidence()
        if ev and "fwd" in ev:
            line["window_attn_tensor_pipe_pct"] = {"fwd": ev["fwd"]["tensor_pipe_pct_active"],
                                                   "bwd": (ev.get("bwd") or {}).get("tensor_pipe_pct_active"),
                                                   "source": "committed ncu --set full captures, see roofline_attn.ncu"}
    # free the hot-path model before the extras
    del run
    torch.cuda.empty_cache()
    extras = name == "c2_ws12" and not args.no_extras
    if extras:
        fs = full_step(args, w, stream)
        if rank == 0:
            line["full_step"] = fs
    if rank == 0 and extras and world == 1:
        try:
            line["reference_cuda_eager"] = reference_cuda_eager(w, dev)
        except Exception as e:                              # noqa: BLE001 -- the extra must never cost the headline
            line["reference_cuda_eager"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference(w, args.cpu_sample_pairs, steps=1, warmup=0)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier(device_ids=[local])
        dist.destroy_process_group()


# ================================================================================================ full step
def full_step(args, w, stream):
    """SURVEY.md section 8f-1 / 8d "full-step": b200swin encoder + the reference's decoder_v2 (depth + pose heads) under bf16
    autocast in channels_last, batched rotation normalisation, SiLog + pose MSE losses, all-reduce of the 282 M-parameter
    gradient, fused AdamW.  Captured in a CUDA graph when the decoder allows it (Newton polar factor instead of cuSOLVER),
    otherwise timed eagerly."""
    import baseline
    from b200swin import SiLogLoss
    world, rank, local = env()
    if not baseline.available():
        return {"unavailable": baseline.why_unavailable()}
    P = args.pairs or w["pairs"]
    cfg = BL.encoder_cfg(w)
    crit = SiLogLoss()
    torch.manual_seed(0)
    out = {}
    try:
        model = BL.FullDepthModel(cfg, w["max_depth"], rot_method="newton")
        host = BL.make_batch(P, 1234 + rank, w["img"], w["max_depth"], w["invalid"], pin=True, pose=True)
        run = TrainRun(args, w, model, host, lambda o, b: BL.full_step_loss(o, b, crit), cfg["depths"], stream)
        for _ in range(3):
            loss = run.step_eager(run.statics[0])
        torch.cuda.synchronize()
        assert torch.isfinite(loss).item(), "non-finite loss (full step)"
        execution = "eager"
        if not args.no_graph:
            try:
                run.capture()
                execution = "cuda_graph_replay" + ("" if run.full_capture else "+bucketed_allreduce+fused_adamw")
            except Exception as e:                          # noqa: BLE001 -- cuDNN / BatchNorm capture is best effort
                run.graphs = []
                torch.cuda.synchronize()
                execution = f"eager (graph capture failed: {type(e).__name__})"
        for _ in range(2):
            run.step_set(0)
        ms = run.timed(lambda: run.step_set(0), args.steps)
        frames = 2 * P * world
        nparam = sum(p.numel() for p in run.params)
        out = {"value": frames * args.steps / (ms / 1e3), "unit": "images/s", "ms_per_step": ms / args.steps,
               "execution": execution, "parameters": nparam, "allreduce_bytes": 4 * run.opt.flat.total if world > 1 else 0,
               "what": "b200swin encoder + reference decoder_v2 (bf16 autocast, channels_last, batched rot normalisation) + "
                       "SiLog x2 + pose MSE x4 + layer-decay fused AdamW" + (" + bucketed NCCL all-reduce" if world > 1 else "")}
        del run, model
    except Exception as e:                                  # noqa: BLE001
        out = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    torch.cuda.empty_cache()
    return out


# ================================================================================================ reference arms
def _ref_train_setup(w, pairs, device):
    import baseline
    ref = baseline.load()
    cfg = BL.encoder_cfg(w)
    torch.manual_seed(0)
    model = BL.EncoderReadout(cfg, w["max_depth"], reference_modules=ref).to(device).train()
    crit = ref.criterion.SiLogLoss()
    import copy
    opt = baseline.quiet(ref.optimizer.build_optimizers, _Enc(model), dict(
        type="AdamW", lr=5e-4, betas=(0.9, 0.999), weight_decay=0.05, constructor="SwinLayerDecayOptimizerConstructor",
        paramwise_cfg=dict(num_layers=copy.copy(cfg["depths"]), layer_decay_rate=0.9,
                           no_decay_names=["relative_position_bias_table", "rpe_mlp", "logit_scale"])))
    batch = [t.to(device) for t in BL.make_batch(pairs, 1234, w["img"], w["max_depth"], w["invalid"])]
    return model, crit, opt, batch


class _Enc(torch.nn.Module):
    """Gives the reference optimizer constructor the `encoder.` / other-name split of models/model.py."""

    def __init__(self, m):
        super().__init__()
        self.encoder = m.encoder
        self.decoder = m.readout


def _ref_step(model, crit, opt, batch, amp):
    opt.zero_grad()
    with torch.autocast(batch[0].device.type, torch.bfloat16, enabled=amp):
        p1, p2 = model(batch[0], batch[1])
    loss = (crit(p1, batch[2]) + crit(p2, batch[3])) / 2
    loss.backward()
    opt.step()
    return loss


def reference_cuda_eager(w, dev):
    """The reference's own PyTorch modules (unmodified, staged copy) on the same B200: same step (encoder + read-out +
    SiLog x2 + AdamW on the reference's layer-decay groups), eager, fp32 as the reference runs it and under bf16 autocast.
    Batch cut to 8 pairs: the reference materialises every [B_, nH, N, N] attention matrix in HBM."""
    import baseline
    if not baseline.available():
        return {"unavailable": baseline.why_unavailable()}
    out = {"pairs_per_step": 8, "what": "reference SwinTransformerV2 (baseline/_ref) + read-out + reference SiLog + "
                                        "torch AdamW on the reference's layer-decay groups, eager PyTorch on the same GPU"}
    for tag, amp in (("fp32", False), ("bf16_autocast", True)):
        model, crit, opt, batch = _ref_train_setup(w, 8, dev)
        for _ in range(2):
            loss = _ref_step(model, crit, opt, batch, amp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 3
        e0.record()
        for _ in range(n):
            loss = _ref_step(model, crit, opt, batch, amp)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        out[tag] = {"value": 16 / (ms / 1e3), "unit": "images/s", "ms_per_step": ms, "loss": float(loss)}
        del model, opt, batch
        torch.cuda.empty_cache()
    return out


def cpu_reference(w, pairs, steps, warmup):
    """The reference's OWN modules (staged copy baseline/_ref, stock code path) on the host cores: a bounded sample of the
    same workload -- `pairs` frame pairs, forward + backward + SiLog + AdamW, fp32, all host threads.  Falls back to the
    CPU oracle port when the staged copy is missing."""
    import baseline
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kind = "reference" if baseline.available() else "port"
    if kind == "reference":
        model, crit, opt, batch = _ref_train_setup(w, pairs, torch.device("cpu"))
        one = lambda: float(_ref_step(model, crit, opt, batch, False))              # noqa: E731
    else:
        one = _oracle_port_step(w, pairs)
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return {"value": 2 * pairs * steps / dt, "unit": "images/s", "cores": cores, "kind": kind, "ms_per_step": dt * 1e3 / steps,
            "sample": f"{steps} step(s) of {pairs} pair(s) ({2 * pairs} frames) {w['img'][0]}x{w['img'][1]}, Swin-V2-{w['size']} "
                      f"windows {w['windows']} fwd+bwd+SiLog+AdamW, fp32 torch CPU, {dt:.1f} s"}


def _oracle_port_step(w, pairs):
    from oracle import silog_ref, swin_ref
    from b200swin.swin_transformer_v2 import SwinTransformerV2
    cfg = BL.encoder_cfg(w)
    torch.manual_seed(0)
    enc = SwinTransformerV2(**cfg)                      # parameter container only (never executed on the CPU)
    enc.init_weights(None)
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and "relative_coords" not in k)
          for k, v in enc.state_dict().items()}
    rw = (torch.randn(1024, cfg["embed_dim"] * 8) * 0.02).requires_grad_(True)
    rb = torch.zeros(1024, requires_grad=True)
    img1, img2, d1, d2 = BL.make_batch(pairs, 1234, w["img"], w["max_depth"], w["invalid"])
    opt = torch.optim.AdamW([v for v in sd.values() if v.requires_grad] + [rw, rb], lr=5e-4, weight_decay=0.05)

    def one():
        opt.zero_grad(set_to_none=True)
        feat = swin_ref.swin_v2(torch.cat([img1, img2]), sd, cfg["embed_dim"], cfg["depths"], cfg["num_heads"],
                                cfg["window_size"], cfg["use_shift"], (3,))[0]
        B, C, h, ww = feat.shape
        d = torch.nn.functional.linear(feat.permute(0, 2, 3, 1).reshape(B, h * ww, C), rw, rb)
        d = torch.sigmoid(d.view(B, h, ww, 32, 32).permute(0, 1, 3, 2, 4).reshape(B, h * 32, ww * 32)) * w["max_depth"]
        p1, p2 = d.chunk(2)
        loss = (silog_ref.silog_torch(p1, d1) + silog_ref.silog_torch(p2, d2)) / 2
        loss.backward()
        opt.step()
        return loss.item()
    return one


def run_reference(args, w, name):
    world, rank, local = env()
    if rank != 0:
        return
    if w["kind"] != "train":
        print(json.dumps({"impl": "reference", "unavailable": f"workload {name}: the reference arm times training workloads"}))
        return
    pairs = min(args.cpu_sample_pairs, args.pairs or w["pairs"])
    steps, warmup = max(1, min(args.steps, 2)), min(args.warmup, 1)
    t0 = time.perf_counter()
    base = cpu_reference(w, pairs, steps, warmup)
    P = args.pairs or w["pairs"]
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": "images/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": base["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{name}: swin_v2_{w['size']}_{w['img'][0]}x{w['img'][1]}_windows{w['windows']}_"
                               f"{P}pairs_per_gpu_train_step(encoder+pixelshuffle_readout+silog_x2+layer_decay_adamw)",
                   "pairs_per_gpu": P, "frames_per_gpu": 2 * P, "windows": w["windows"], "parallelism": f"dp{args.gpus}",
                   "execution": "reference modules (baseline/_ref), CPU, all host threads" if base["kind"] == "reference"
                   else "cpu oracle port",
                   "sample": f"each step = {pairs} pair(s) of the workload (bounded sample of the {P}-pair batch)",
                   "note": f"requested steps/warmup ({args.steps}/{args.warmup}) clamped to ({steps}/{warmup}) to bound the run"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ================================================================================================ other workloads
def run_infer(args, w, name):
    """BASELINE config 4: batch-sharded inference, no collective.  Every rank runs its own frames; value = all frames / the
    slowest rank's time."""
    import torch.distributed as dist
    from b200swin import _lib, ops
    world, rank, local = env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops.ATTN_IMPL["mode"] = args.attn
    cfg = BL.encoder_cfg(w, drop_path_rate=0.0)
    torch.manual_seed(0)
    model = BL.EncoderReadout(cfg, w["max_depth"]).to(dev).eval()
    F = w["frames"]
    g = torch.Generator().manual_seed(1234 + rank)
    host = torch.rand(F, 3, *w["img"], generator=g).pin_memory()
    x = host.to(dev)
    amp = args.dtype == "bf16"
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream), torch.no_grad():
        def fwd(inp):
            with torch.autocast("cuda", torch.bfloat16, enabled=amp):
                return model(inp)
        for _ in range(max(args.warmup, 3)):
            y = fwd(x)
        torch.cuda.synchronize()
        _lib.reset_counters()
        _lib.TIMING.update(name="*", events=[])
        fwd(x)
        torch.cuda.synchronize()
        _lib.TIMING["name"] = None
        launches = _lib.COUNTERS["launches"]
        fam = BL.rooflines_from_events(_lib.TIMING["events"], 1, BL.peaks())
        graph = None
        if not args.no_graph:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                y = fwd(x)
        run = (lambda: graph.replay()) if graph is not None else (lambda: fwd(x))
        run()

        def timed(fn, n):
            if world > 1:
                dist.barrier(device_ids=[local])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            if world > 1:
                dist.barrier(device_ids=[local])
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = t.item()
            return ms
        clk_path = os.path.join(tempfile.gettempdir(), f"b200swin_clocks_{rank}.csv")
        sampler = BL.clocks_sampler(clk_path) if rank == 0 else None
        ms = timed(run, args.steps)
        out_host = torch.empty(y.shape, dtype=y.dtype).pin_memory()

        def e2e_once():
            x.copy_(host, non_blocking=True)
            run()
            out_host.copy_(y, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        e2e_once()
        ms_e2e = timed(e2e_once, args.steps)
        if sampler is not None:
            sampler.terminate()
    if rank == 0:
        gemm_f, attn_f, _ = BL.encoder_flops_per_frame(cfg, w["img"])
        print(json.dumps({
            "metric": "inference images/sec, Swin-V2-L depth 352x1216 (encoder + read-out), batch-sharded",
            "value": F * world * args.steps / (ms / 1e3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if amp else "f32(split-bf16 x3)", "data": "synthetic",
            "config": {"workload": f"{name}: swin_v2_{w['size']}_{w['img'][0]}x{w['img'][1]}_windows{w['windows']}_pretrain{w['pre']}_"
                                   f"{F}frames_per_gpu_inference(no collective)", "frames_per_gpu": F,
                       "parallelism": f"batch-sharded x{world}", "execution": "cuda_graph_replay" if graph else "eager"},
            "e2e": {"value": F * world * args.steps / (ms_e2e / 1e3), "unit": "images/s",
                    "h2d_bytes_per_step": host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * out_host.element_size()},
            "gpu_launches": launches,
            "kernel_ms_per_step": {k: round(v["ms_per_step"], 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms_per_step"])},
            "roofline": {"bound": "tensor", "achieved": fam.get("gemm", {}).get("tflops"), "peak": BL.peaks()["tf_sustained"],
                         "unit": "TFLOP/s", "frac": fam.get("gemm", {}).get("frac_of_sustained_bf16_peak"), "traffic": None,
                         "kernel": "gemm_tc_kernel (all launches)"},
            "roofline_attn": {"fwd": fam.get("attn_fwd")},
            "model_flops": {"encoder_fwd_gflop_per_frame": (gemm_f + attn_f) / 1e9,
                            "tflops_achieved": (gemm_f + attn_f) * F / (ms / args.steps / 1e3) / 1e12},
            "clocks": BL.clocks_summary(clk_path, local)}))
    if world > 1:
        dist.destroy_process_group()


def run_micro(args, name):
    """BASELINE config 5: the attention half-block (pad / roll / partition / WindowAttention / reverse / unroll / crop =
    `attn.attend` of the drop-in; the reference's SwinTransformerBlockPost attention half) for windows 8/12/16/24 and
    3..48 heads, shifted and unshifted, forward and forward+backward, b200swin (bf16 autocast) vs the reference modules
    (fp32 eager and bf16 autocast) on the SAME GPU.  Token grid ~1.4 M tokens at nH <= 6, scaled down with C."""
    import baseline
    from b200swin import swin_transformer_v2 as S
    from b200swin import _lib
    world, rank, local = env()
    if rank != 0:
        return
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    have_ref = baseline.available()
    ref = baseline.load() if have_ref else None
    rows = []

    def timeit(fn, n=5):
        fn(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for ws in (8, 12, 16, 24):
        for nH in (3, 4, 6, 8, 12, 16, 24, 32, 48):
            C = 32 * nH
            side = 96 if ws != 24 else 96
            B = max(1, int(round(1.4e6 * min(1.0, 192.0 / C) / (side * side))))
            for shift in (0, ws // 2):
                torch.manual_seed(0)
                mine = S.SwinTransformerBlockPost(dim=C, num_heads=nH, window_size=ws, shift_size=shift,
                                                  relative_coords_table_type="norm8_log_bylayer", rpe_output_type="sigmoid",
                                                  attn_type="cosine_mh", pretrain_window_size=ws).to(dev)
                mine.H = mine.W = side
                x = torch.randn(B, side * side, C, device=dev, requires_grad=True)
                cot = torch.randn(B, side * side, C, device=dev)

                def mine_fwd():
                    with torch.no_grad(), torch.autocast("cuda", torch.bfloat16):
                        return mine.attn.attend(x, B, side, side, shift)

                def mine_fb():
                    x.grad = None
                    with torch.autocast("cuda", torch.bfloat16):
                        y = mine.attn.attend(x, B, side, side, shift)
                    y.backward(cot.to(y.dtype))
                r = {"ws": ws, "nH": nH, "C": C, "tokens": B * side * side, "shift": shift,
                     "b200_fwd_ms": timeit(mine_fwd), "b200_fwdbwd_ms": timeit(mine_fb)}
                if have_ref:
                    rb = baseline.quiet(ref.swin.SwinTransformerBlockPost, dim=C, num_heads=nH, window_size=ws,
                                        shift_size=shift, relative_coords_table_type="norm8_log_bylayer",
                                        rpe_output_type="sigmoid", attn_type="cosine_mh", pretrain_window_size=ws).to(dev)
                    rb.H = rb.W = side
                    layer = baseline.quiet(ref.swin.BasicLayer, dim=C, depth=1, num_heads=nH, window_size=ws)
                    mask = _ref_mask(layer, side, side, ws, shift, dev) if shift else None

                    def ref_half(inp):                       # the attention half of SwinTransformerBlockPost.forward (:419-463)
                        xx = inp.view(B, side, side, C)
                        if shift:
                            xx = torch.roll(xx, shifts=(-shift, -shift), dims=(1, 2))
                        xw = ref.swin.window_partition(xx, ws).view(-1, ws * ws, C)
                        aw = rb.attn(xw, mask=mask).view(-1, ws, ws, C)
                        xx = ref.swin.window_reverse(aw, ws, side, side)
                        if shift:
                            xx = torch.roll(xx, shifts=(shift, shift), dims=(1, 2))
                        return xx.view(B, side * side, C)
                    for tag, amp in (("fp32", False), ("bf16", True)):
                        def ref_fwd():
                            with torch.no_grad(), torch.autocast("cuda", torch.bfloat16, enabled=amp):
                                return ref_half(x)

                        def ref_fb():
                            x.grad = None
                            with torch.autocast("cuda", torch.bfloat16, enabled=amp):
                                y = ref_half(x)
                            y.backward(cot.to(y.dtype))
                        try:
                            r[f"ref_{tag}_fwd_ms"] = timeit(ref_fwd, 3)
                            r[f"ref_{tag}_fwdbwd_ms"] = timeit(ref_fb, 3)
                        except torch.OutOfMemoryError:
                            r[f"ref_{tag}_fwd_ms"] = r[f"ref_{tag}_fwdbwd_ms"] = None
                            torch.cuda.empty_cache()
                    if r.get("ref_bf16_fwdbwd_ms"):
                        r["speedup_fwdbwd_vs_ref_bf16"] = r["ref_bf16_fwdbwd_ms"] / r["b200_fwdbwd_ms"]
                        r["speedup_fwd_vs_ref_bf16"] = r["ref_bf16_fwd_ms"] / r["b200_fwd_ms"]
                    del rb
                rows.append(r)
                del mine, x, cot
                torch.cuda.empty_cache()
    sp = [r["speedup_fwdbwd_vs_ref_bf16"] for r in rows if r.get("speedup_fwdbwd_vs_ref_bf16")]
    tok_s = sum(r["tokens"] for r in rows) / (sum(r["b200_fwdbwd_ms"] for r in rows) / 1e3)
    print(json.dumps({
        "metric": "attention half-block tokens/sec (fwd+bwd), windows 8/12/16/24 x heads 3..48, shifted and unshifted",
        "value": tok_s, "unit": "tokens/s", "n_gpus": 1, "steps": 5, "warmup": 2, "ms_per_step": None,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{name}: qkv GEMM + window attention core + proj GEMM on a 96x96 token grid, batch scaled to "
                               "~1.4 M tokens at nH <= 6", "reference": "reference SwinTransformerBlockPost attention half "
                               "(baseline/_ref), eager PyTorch on the same GPU" if have_ref else "unavailable"},
        "speedup_fwdbwd_vs_reference_bf16": {"min": min(sp) if sp else None, "median": sorted(sp)[len(sp) // 2] if sp else None,
                                             "max": max(sp) if sp else None},
        "rows": rows}))


def _ref_mask(layer, H, W, ws, shift, dev):
    """The reference's own mask construction (BasicLayer.forward, :874-892), run once on the CPU."""
    import numpy as np
    Hp = int(np.ceil(H / ws)) * ws
    Wp = int(np.ceil(W / ws)) * ws
    img_mask = torch.zeros((1, Hp, Wp, 1))
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img_mask[:, hs, wsl, :] = cnt
            cnt += 1
    import baseline
    mw = baseline.load().swin.window_partition(img_mask, ws).view(-1, ws * ws)
    am = mw.unsqueeze(1) - mw.unsqueeze(2)
    return am.masked_fill(am != 0, float(-100.0)).masked_fill(am == 0, float(0.0)).to(dev)


def run_silog(args, w, name):
    """BASELINE config 3 (VOID 480 x 640 through cnn_transformer): the path holds no window attention; the piece of the hot
    path that applies is the SiLog loss (forward + backward): HBM roofline, 8 B/px forward + 12 B/px backward."""
    import b200swin
    world, rank, local = env()
    if rank != 0:
        return
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    F = w["frames"]
    H, W = w["img"]
    g = torch.Generator().manual_seed(1234)
    tgt_h = (0.5 + (w["max_depth"] - 0.5) * torch.rand(F, H, W, generator=g))
    tgt_h = torch.where(torch.rand(F, H, W, generator=g) < w["invalid"], torch.zeros(()), tgt_h).pin_memory()
    pred = (torch.rand(F, H, W, generator=g) * w["max_depth"] + 0.1).to(dev).requires_grad_(True)
    tgt = tgt_h.to(dev)
    crit = b200swin.SiLogLoss()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def one():
        pred.grad = None
        loss = crit(pred, tgt)
        loss.backward()
        return loss
    for _ in range(max(args.warmup, 3)):
        one()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(args.steps):
        flush.zero_()                                     # L2 flush between timed iterations (256 MB > 126 MB)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        one()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / args.steps
    px = F * H * W
    pk = BL.peaks()
    gbs = 20.0 * px / (ms / 1e3) / 1e9
    # the reference's loss on the same GPU (boolean-index compaction + host sync)
    ref_ms = None
    import baseline
    if baseline.available():
        rc = baseline.load().criterion.SiLogLoss()
        p2 = pred.detach().clone().requires_grad_(True)

        def ref_one():
            p2.grad = None
            rc(p2, tgt).backward()
        ref_one(); ref_one()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ref_one()
        torch.cuda.synchronize()
        ref_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    print(json.dumps({
        "metric": "SiLog loss fwd+bwd images/sec, VOID-shaped 480x640", "value": F / (ms / 1e3), "unit": "images/s",
        "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{name}: SiLogLoss forward+backward on {F} x {H}x{W} fp32 maps, 30% invalid",
                   "l2": "256 MB flush write between timed iterations"},
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                     "traffic": None, "kernel": "silog_fwd + silog_final + silog_bwd (8 + 12 B/px incl. launch gaps)",
                     "peak_source": pk["source"]},
        "reference_cuda_eager": {"ms_per_step": ref_ms, "speedup": (ref_ms / ms) if ref_ms else None},
        "gpu_launches": 3 * args.steps}))


def run_mha(args, w, name):
    """BASELINE config 3's transformer encoder (the part of cnn_transformer that is attention): `layers` x
    Transformer_Encoder on [frames, 1200, 512] feature / position maps, batched inference under bf16 autocast, through the
    b200swin drop-in (global-attention kernels + tcgen05 GEMMs + LayerNorm kernels), beside the reference's own
    Transformer_Encoder modules on the same GPU (eager: nn.MultiheadAttention with need_weights=True as the reference
    calls it) and on the host cores."""
    import types
    from b200swin import _lib
    from b200swin.cnn_transformer import Transformer_Encoder
    world, rank, local = env()
    if rank != 0:
        return
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    F, E, nH, FF, NL = w["frames"], w["hidden"], w["heads"], w["ff"], w["layers"]
    N = (w["img"][0] // 16) * (w["img"][1] // 16)
    amp = args.dtype == "bf16"
    cargs = types.SimpleNamespace(transformer_ff_dim=FF)
    torch.manual_seed(0)
    layers = torch.nn.ModuleList([Transformer_Encoder(cargs, hidden_dim=E) for _ in range(NL)]).to(dev).eval()
    g = torch.Generator().manual_seed(1234)
    feat_h = torch.randn(F, N, E, generator=g).pin_memory()
    pos_h = torch.randn(F, N, E, generator=g).pin_memory()
    feat, pos = feat_h.to(dev), pos_h.to(dev)

    def fwd(mods, f, p):
        with torch.autocast("cuda", torch.bfloat16, enabled=amp):
            for m in mods:
                f = m(f, p)
        return f
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream), torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            y = fwd(layers, feat, pos)
        torch.cuda.synchronize()
        _lib.reset_counters()
        _lib.TIMING.update(name="*", events=[])
        fwd(layers, feat, pos)
        torch.cuda.synchronize()
        _lib.TIMING["name"] = None
        launches = _lib.COUNTERS["launches"]
        fam = {}
        for e0, e1, sym, _a in _lib.TIMING["events"]:
            fam[sym] = fam.get(sym, 0.0) + e0.elapsed_time(e1)
        graph = None
        if not args.no_graph:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                y = fwd(layers, feat, pos)
        run = (lambda: graph.replay()) if graph is not None else (lambda: fwd(layers, feat, pos))
        run()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def timed(fn, n):
            tot = 0.0
            for _ in range(n):
                flush.zero_()                         # L2 flush between timed iterations (256 MB > 126 MB)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            return tot / n
        clk_path = os.path.join(tempfile.gettempdir(), "b200swin_clocks_0.csv")
        sampler = BL.clocks_sampler(clk_path)
        ms = timed(run, args.steps)
        out_host = torch.empty(y.shape, dtype=y.dtype).pin_memory()

        def e2e_once():
            feat.copy_(feat_h, non_blocking=True)
            pos.copy_(pos_h, non_blocking=True)
            run()
            out_host.copy_(y, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        e2e_once()
        ms_e2e = timed(e2e_once, args.steps)
        sampler.terminate()
        # the reference's own layers on the same GPU, same weights
        ref = {}
        import baseline
        if baseline.available():
            R = baseline.load().cnn_transformer
            rl = torch.nn.ModuleList([R.Transformer_Encoder(cargs, hidden_dim=E) for _ in range(NL)]).to(dev).eval()
            rl.load_state_dict(layers.state_dict(), strict=True)
            for tag, a in (("fp32", False), ("bf16_autocast", True)):
                def ref_fwd():
                    f = feat
                    with torch.autocast("cuda", torch.bfloat16, enabled=a):
                        for m in rl:
                            f = m(f, pos)
                    return f
                yr = ref_fwd()
                ref_fwd()
                ref[tag] = {"ms_per_step": timed(ref_fwd, max(2, args.steps // 2))}
                ref[tag]["images_per_s"] = F / (ref[tag]["ms_per_step"] / 1e3)
                ref[tag]["speedup"] = ref[tag]["ms_per_step"] / ms
                if a == amp:
                    ref["max_abs_diff_vs_reference_same_precision"] = (y.float() - yr.float()).abs().max().item()
    # host cores: the reference's layers, one frame per step (bounded sample)
    cpu = None
    import baseline
    if baseline.available():
        R = baseline.load().cnn_transformer
        cl = torch.nn.ModuleList([R.Transformer_Encoder(cargs, hidden_dim=E) for _ in range(NL)]).eval()
        torch.set_num_threads(os.cpu_count() or 1)
        with torch.no_grad():
            f1, p1 = feat_h[:1].clone(), pos_h[:1].clone()
            def cpu_one():
                f = f1
                for m in cl:
                    f = m(f, p1)
                return f
            cpu_one()
            t0 = time.perf_counter()
            nrep = 3
            for _ in range(nrep):
                cpu_one()
            cpu_s = (time.perf_counter() - t0) / nrep
        cpu = {"value": 1.0 / cpu_s, "unit": "images/s", "cores": os.cpu_count(), "kind": "reference",
               "sample": f"{nrep} x 1 frame through the reference's {NL} Transformer_Encoder layers, fp32"}
    fl_attn = 4.0 * N * N * E
    fl_gemm = 2.0 * N * E * 3 * E + 2.0 * N * E * E + 4.0 * N * E * FF
    pk = BL.peaks()
    tf = (fl_attn + fl_gemm) * NL * F / (ms / 1e3) / 1e12
    attn_ms = fam.get("b200swin_mha_fwd", 0.0)
    print(json.dumps({
        "metric": "inference images/sec, cnn_transformer encoder layers (global MHA), VOID-shaped 480x640",
        "value": F / (ms / 1e3), "unit": "images/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if amp else "f32(split-bf16 x3 GEMMs, fp32 CUDA-core attention)", "data": "synthetic",
        "config": {"workload": f"{name}: {NL} x Transformer_Encoder(hidden {E}, {nH} heads x {E // nH}, ff {FF}) on {F} frames x "
                               f"{N} tokens ({w['img'][0]}x{w['img'][1]} / 16), inference",
                   "execution": "cuda_graph_replay" if graph else "eager",
                   "l2": "256 MB flush write between timed iterations"},
        "e2e": {"value": F / (ms_e2e / 1e3), "unit": "images/s", "h2d_bytes_per_step": 2 * feat_h.numel() * 4,
                "d2h_bytes_per_step": out_host.numel() * out_host.element_size()},
        "gpu_launches": launches,
        "kernel_ms_per_step": {k: round(v, 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])},
        "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                     "frac": tf / pk["tf_sustained"], "traffic": None,
                     "kernel": "whole layer stack (tcgen05 GEMMs + warp-MMA global attention)", "peak_source": pk["source"]},
        "roofline_attn": {"fwd": {"ms_per_step": attn_ms, "tflops": (fl_attn * NL * F / (attn_ms / 1e3) / 1e12) if attn_ms else None,
                                  "kernel": "gattn_fwd_mma_kernel<64> (mma.sync m16n8k16; legacy tensor pipe ~550 TF/s)"}},
        "reference_cuda_eager": ref, "cpu_baseline": cpu, "clocks": BL.clocks_summary(clk_path, local)}))


if __name__ == "__main__":
    a = parse()
    wl = BL.WORKLOADS[a.workload]
    if a.impl == "reference":
        run_reference(a, wl, a.workload)
    elif wl["kind"] == "train":
        run_train(a, wl, a.workload)
    elif wl["kind"] == "infer":
        run_infer(a, wl, a.workload)
    elif wl["kind"] == "micro":
        run_micro(a, a.workload)
    elif wl["kind"] == "mha":
        run_mha(a, wl, a.workload)
    else:
        run_silog(a, wl, a.workload)
