import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200swin import ops
dev = "cuda"
torch.manual_seed(0)
def attempt(name, fn):
    ops._weight_cache.clear()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(); fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    ops._weight_cache.clear()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        print(name, "OK")
    except Exception as e:
        print(name, "FAILED:", str(e).splitlines()[0])
which = sys.argv[1]
T, C = 4096, 128
x = torch.randn(2, T // 2, C, device=dev, dtype=torch.bfloat16, requires_grad=True)
w = torch.randn(3 * C, C, device=dev, requires_grad=True); b = torch.randn(3 * C, device=dev, requires_grad=True)
if which == "linear":
    def f():
        y = ops.linear(x, w, b); y.float().sum().backward()
    attempt("linear fwd+bwd", f)
elif which == "ln":
    g_ = torch.ones(C, device=dev, requires_grad=True); be = torch.zeros(C, device=dev, requires_grad=True)
    def f():
        y = ops.layer_norm_residual(x, g_, be, 1e-6, residual=x); y.float().sum().backward()
    attempt("ln fwd+bwd", f)
elif which == "mlp":
    w1 = torch.randn(4 * C, C, device=dev, requires_grad=True); b1 = torch.randn(4 * C, device=dev, requires_grad=True)
    w2 = torch.randn(C, 4 * C, device=dev, requires_grad=True); b2 = torch.randn(C, device=dev, requires_grad=True)
    def f():
        y = ops.mlp(x, w1, b1, w2, b2); y.float().sum().backward()
    attempt("mlp fwd+bwd", f)
elif which == "attn":
    B, H, W, ws, nH = 2, 24, 24, 12, 4
    qkv = torch.randn(B, H, W, 3 * C, device=dev, dtype=torch.bfloat16, requires_grad=True)
    inv = torch.ones(B * H * W, 2, nH, device=dev)
    tab = torch.rand((2 * ws - 1) ** 2, nH, device=dev, requires_grad=True)
    sc = torch.full((nH,), 10.0, device=dev, requires_grad=True)
    def f():
        o = ops.attention_core(qkv, inv, tab, sc, None, None, None, B, H, W, C, nH, ws, 6); o.float().sum().backward()
    attempt("attention fwd+bwd", f)
elif which == "torchonly":
    lin = torch.nn.Linear(C, C).to(dev)
    def f():
        y = lin(x.float()); y.sum().backward()
    attempt("torch linear fwd+bwd", f)
if which == "qkv":
    qb = torch.randn(C, device=dev, requires_grad=True); vb = torch.randn(C, device=dev, requires_grad=True)
    def f():
        y, inv = ops.qkv_project(x, w, qb, vb, 4); y.float().sum().backward()
    attempt("qkv fwd+bwd", f)
elif which in ("conv", "merge", "layer", "enc"):
    from b200swin import swin_transformer_v2 as S
    import bench
    enc = S.SwinTransformerV2(**bench.CFG).to(dev).train()
    img = torch.rand(4, 3, 480, 480, device=dev)
    if which == "conv":
        def f():
            with torch.autocast("cuda", torch.bfloat16):
                y = enc.patch_embed(img)
            (y[0] if isinstance(y, tuple) else y).float().sum().backward()
        attempt("patch_embed fwd+bwd", f)
    elif which == "merge":
        pm = enc.layers[0].downsample
        xx = torch.randn(4, 120 * 120, 128, device=dev, dtype=torch.bfloat16, requires_grad=True)
        def f():
            with torch.autocast("cuda", torch.bfloat16):
                y = pm(xx, 120, 120)
            y.float().sum().backward()
        attempt("patch merging fwd+bwd", f)
    elif which == "layer":
        ly = enc.layers[0]
        xx = torch.randn(4, 120 * 120, 128, device=dev, dtype=torch.bfloat16, requires_grad=True)
        def f():
            with torch.autocast("cuda", torch.bfloat16):
                out = ly(xx, 120, 120)
            out[0].float().sum().backward()
        attempt("BasicLayer 0 fwd+bwd", f)
    else:
        def f():
            with torch.autocast("cuda", torch.bfloat16):
                out = enc(img)
            out[0].float().sum().backward()
        attempt("encoder fwd+bwd", f)
