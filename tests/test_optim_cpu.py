"""Host logic of the optimizer path on CPU: the layer-decay parameter grouping must equal what the REFERENCE's own
``SwinLayerDecayOptimizerConstructor`` (models/optimizer.py:36-104, run from the staged copy ``baseline/_ref`` through
the mmcv shim) builds for the same model -- including the reference's ``IDEDepth`` wrapped around the drop-in encoder --
and the flat parameter layout must keep ``state_dict`` intact."""
import argparse

import pytest
import torch

import baseline

needs_ref = pytest.mark.skipif(not baseline.available(), reason=baseline.why_unavailable() if not baseline.available() else "")

SWIN_T = dict(depths=[2, 2, 6, 2], window_size=[4, 4, 4, 2], pretrain_window_size=[4, 4, 4, 2], drop_path_rate=0.1,
              use_checkpoint=False, use_shift=[True, True, False, False])


def _idedepth_args():
    return argparse.Namespace(backbone="swin_tiny_v2", model_scale=32, decoder="decoder_v1", max_depth=10.0,
                              pretrained=None, **SWIN_T)


def test_get_num_layer_matches_the_reference_function():
    from b200swin.optim import get_num_layer_for_swin
    lps = [3, 3, 7, 2]
    cases = {"encoder.patch_embed.proj.weight": 0, "backbone.patch_embed.norm.bias": 0,
             "encoder.layers.0.blocks.1.attn.qkv.weight": 2, "encoder.layers.2.blocks.5.mlp.fc1.bias": 3 + 3 + 5 + 1,
             "encoder.layers.1.downsample.reduction.weight": 6, "decoder.deconv_layers.0.weight": 16,
             "encoder.norm3.weight": 16}
    for name, want in cases.items():
        assert get_num_layer_for_swin(name, 17, lps) == want, name
    if baseline.available():
        ref = baseline.load().optimizer.get_num_layer_for_swin
        for name in cases:
            assert ref(name, 17, lps) == get_num_layer_for_swin(name, 17, lps), name


@needs_ref
def test_reference_idedepth_accepts_the_drop_in_encoder_and_groups_identically():
    """models/model.py builds IDEDepth(encoder=SwinTransformerV2(...), decoder=...): swap the encoder for the drop-in
    built from the SAME keyword arguments (model.py:40-49), check the state_dict contract, then build the optimizer with
    the reference's constructor and with ours and compare group by group."""
    import copy
    from b200swin.optim import layer_decay_param_groups
    from b200swin.swin_transformer_v2 import SwinTransformerV2
    ref = baseline.load()
    args = _idedepth_args()
    model = baseline.quiet(ref.model.IDEDepth, args)
    ref_keys = list(model.state_dict().keys())
    enc = SwinTransformerV2(embed_dim=96, depths=args.depths, num_heads=[3, 6, 12, 24], window_size=args.window_size,
                            pretrain_window_size=args.pretrain_window_size, drop_path_rate=args.drop_path_rate,
                            use_checkpoint=args.use_checkpoint, use_shift=args.use_shift)
    enc.init_weights(pretrained=args.pretrained)
    model.encoder = enc
    assert list(model.state_dict().keys()) == ref_keys             # same names in the same order
    cfg = dict(type="AdamW", lr=5e-4, betas=(0.9, 0.999), weight_decay=0.05,
               constructor="SwinLayerDecayOptimizerConstructor",
               paramwise_cfg=dict(num_layers=copy.copy(args.depths), layer_decay_rate=0.9,
                                  no_decay_names=["relative_position_bias_table", "rpe_mlp", "logit_scale"]))
    ref_opt = baseline.quiet(ref.optimizer.build_optimizers, model, copy.deepcopy(cfg))       # train.py:113-115
    mine = layer_decay_param_groups(model, 5e-4, 0.05, args.depths, 0.9,
                                    ("relative_position_bias_table", "rpe_mlp", "logit_scale"))
    assert len(mine) == len(ref_opt.param_groups) and len(mine) > 20
    for g, r in zip(mine, ref_opt.param_groups):
        assert g["group_name"] == r["group_name"]
        assert g["param_names"] == r["param_names"]
        assert g["weight_decay"] == r["weight_decay"]
        assert abs(g["lr_scale"] - r["lr_scale"]) < 1e-15 and abs(g["lr"] - r["lr"]) < 1e-15
        assert [id(p) for p in g["params"]] == [id(p) for p in r["params"]]


def test_flat_params_keep_values_names_and_alignment():
    from b200swin.optim import FlatParams
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.LayerNorm(5), torch.nn.Linear(5, 3, bias=False))
    before = {k: v.clone() for k, v in net.state_dict().items()}
    flat = FlatParams(list(net.parameters()), bf16_copies=False)
    assert flat.total % flat.chunk == 0 and all(o % flat.chunk == 0 for o in flat.offsets)
    assert flat.chunk_tensor.tolist() == [0, 1, 2, 3, 4]            # one 1024-element chunk per small tensor
    for k, v in net.state_dict().items():
        assert torch.equal(v, before[k])
    for p in net.parameters():                                        # the parameters now live inside the flat buffer
        assert flat.data.data_ptr() <= p.data_ptr() < flat.data.data_ptr() + 4 * flat.total
    net(torch.randn(2, 7)).sum().backward()
    missing = flat.pack_grads()
    assert not missing
    for p, v in zip(flat.params, flat.grad_views):
        assert torch.equal(p.grad, v)
    with pytest.raises(RuntimeError):
        FlatParams(list(net.parameters()), bf16_copies=True)          # bf16 copies feed CUDA GEMMs: no CPU fallback
