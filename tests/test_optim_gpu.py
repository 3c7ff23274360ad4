"""FusedAdamW (csrc/adamw.cu) against torch.optim.AdamW on the reference's layer-decay parameter groups
(models/optimizer.py:36-104) with the training loop's per-step learning-rate rewrite (train.py:195-203): parameters
within 1e-6 relative after 3 steps, bf16 weight copies equal to a cast of the fp32 masters, state_dict round trip."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(embed_dim=32, depths=[2, 2, 2], num_heads=[1, 2, 4], window_size=[4, 4, 2], pretrain_window_size=[4, 4, 2],
           use_shift=[True, True, False], drop_path_rate=0.0, out_indices=(2,))


class _Wrap(torch.nn.Module):
    def __init__(self):
        super().__init__()
        from b200swin.swin_transformer_v2 import SwinTransformerV2
        self.encoder = SwinTransformerV2(**CFG)
        self.encoder.init_weights(None)
        self.decoder = torch.nn.Linear(128, 7)


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def test_fused_adamw_matches_torch_adamw_over_three_steps_with_layer_decay_groups():
    from b200swin.optim import FusedAdamW, layer_decay_param_groups
    torch.manual_seed(0)
    m1 = _Wrap().cuda()
    m2 = copy.deepcopy(m1)
    g1 = layer_decay_param_groups(m1, 5e-4, 0.05, CFG["depths"], 0.8)
    g2 = layer_decay_param_groups(m2, 5e-4, 0.05, CFG["depths"], 0.8)
    assert len(g1) > 10
    fused = FusedAdamW(g1, lr=5e-4, betas=(0.9, 0.999), weight_decay=0.05)
    ref = torch.optim.AdamW(g2, lr=5e-4, betas=(0.9, 0.999), weight_decay=0.05)
    gen = torch.Generator(device="cuda").manual_seed(1)
    for step in range(3):
        cur = 5e-4 * (0.5 + 0.25 * step)                    # train.py:195-203: lr = current_lr * lr_scale per group
        for opt in (fused, ref):
            for g in opt.param_groups:
                g["lr"] = cur * g["lr_scale"]
        for p, q in zip(m1.parameters(), m2.parameters()):
            gr = torch.randn(p.shape, device="cuda", generator=gen) * (1.0 + step)
            p.grad, q.grad = gr.clone(), gr.clone()
        fused.step()
        ref.step()
    worst = max(_rel(p, q) for p, q in zip(m1.parameters(), m2.parameters()))
    assert worst < 1e-6, worst
    # the bf16 copies the GEMMs read are the cast of the updated masters
    for p, v16 in zip(fused.flat.params, fused.flat.bf16_views):
        assert torch.equal(v16, p.detach().bfloat16())
    # moments in torch's per-parameter layout
    sd, rsd = fused.state_dict(), ref.state_dict()
    for k in rsd["state"]:
        assert _rel(sd["state"][k]["exp_avg"], rsd["state"][k]["exp_avg"]) < 1e-6
        assert _rel(sd["state"][k]["exp_avg_sq"], rsd["state"][k]["exp_avg_sq"]) < 1e-6
        assert float(sd["state"][k]["step"]) == float(rsd["state"][k]["step"]) == 3.0


def test_set_lr_and_graph_capture_follow_a_schedule():
    from b200swin.optim import FusedAdamW
    torch.manual_seed(0)
    lin1, lin2 = torch.nn.Linear(40, 30).cuda(), torch.nn.Linear(40, 30).cuda()
    lin2.load_state_dict(lin1.state_dict())
    fused = FusedAdamW(lin1.parameters(), lr=1e-3, weight_decay=0.1)
    ref = torch.optim.AdamW(lin2.parameters(), lr=1e-3, weight_decay=0.1)
    grads = [torch.randn_like(p) for p in lin1.parameters()]
    for p, g in zip(lin1.parameters(), grads):
        p.grad = g.clone()
    fused.flat.pack_grads()
    fused.grads_packed = True
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fused.step()                                          # step 1, eager (also the warm-up before the capture)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):                   # capturing records the launch, it does not run it
            fused.step()
        g.replay()                                            # step 2 at the captured learning rate
        fused.set_lr(3e-3)                                    # the schedule moves ONE device scalar ...
        g.replay()                                            # ... and step 3 of the same graph follows it
    torch.cuda.synchronize()
    for lr in (1e-3, 1e-3, 3e-3):
        for gq in ref.param_groups:
            gq["lr"] = lr
        for p, gr in zip(lin2.parameters(), grads):
            p.grad = gr.clone()
        ref.step()
    for p, q in zip(lin1.parameters(), lin2.parameters()):
        assert _rel(p, q) < 1e-6


def test_gemm_reads_the_optimizer_owned_bf16_weight_and_follows_updates():
    """ops.linear on a FusedAdamW-owned weight reads the flat bf16 copy (no per-step cast) and sees every update,
    including one made outside the optimizer (load_state_dict)."""
    from b200swin import ops
    from b200swin.optim import FusedAdamW
    torch.manual_seed(0)
    lin = torch.nn.Linear(64, 48).cuda()
    opt = FusedAdamW(lin.parameters(), lr=1e-2, weight_decay=0.0)
    x = torch.randn(256, 64, device="cuda")

    def check():
        with torch.autocast("cuda", torch.bfloat16):
            y = ops.linear(x, lin.weight, lin.bias)
        ref = x.bfloat16().float() @ lin.weight.detach().bfloat16().float().t() + lin.bias.detach()
        assert _rel(y.float(), ref) < 1e-2
    check()
    assert ops.stage_weight(lin.weight, False).hi.data_ptr() == opt.flat.bf16_views[0].data_ptr()
    for p in lin.parameters():
        p.grad = torch.randn_like(p)
    opt.step()
    check()
    with torch.no_grad():
        lin.weight.copy_(torch.randn_like(lin.weight))
    check()
