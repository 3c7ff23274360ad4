"""mmcv.cnn: the five helpers models/decoder_v1.py:4, decoder_v2.py:4 and model.py:4 import."""
import torch.nn as nn


def build_conv_layer(cfg, *args, **kwargs):
    return nn.Conv2d(*args, **kwargs)


def build_norm_layer(cfg, num_features, postfix=""):
    return "bn" + str(postfix), nn.BatchNorm2d(num_features)


def build_upsample_layer(cfg, *args, **kwargs):
    typ = (cfg or {}).get("type", "deconv")
    if typ == "deconv":
        return nn.ConvTranspose2d(*args, **kwargs)
    return nn.Upsample(*args, **kwargs)


def constant_init(module, val, bias=0):
    if getattr(module, "weight", None) is not None:
        nn.init.constant_(module.weight, val)
    if getattr(module, "bias", None) is not None:
        nn.init.constant_(module.bias, bias)


def normal_init(module, mean=0, std=1, bias=0):
    if getattr(module, "weight", None) is not None:
        nn.init.normal_(module.weight, mean, std)
    if getattr(module, "bias", None) is not None:
        nn.init.constant_(module.bias, bias)
