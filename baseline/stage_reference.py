"""Stage a runnable copy of the UNMODIFIED reference under ``baseline/_ref/`` (git-ignored, shipped to the GPU box).

    python baseline/stage_reference.py [--src /root/reference] [--force]

The reference is plain Python with no build step; "installing" it means copying the modules of the measured
path next to import shims for the two packages this image lacks (``timm`` and ``mmcv``; ``baseline/shims``,
which are NOT reference code) and applying the one-line device fixes SURVEY.md section 8c lists -- each a literal
string substitution whose original is asserted to occur exactly as expected, so a changed reference fails loudly:

    models/swin_transformer_v2.py:294   .to('cuda:0')            -> .to(self.logit_scale.device)
    utils/util.py:12                    torch.eye(3).cuda()      -> torch.eye(3).to(rot_vector.device)
    models/cnn_transformer.py:171       dtype=torch.bool).cuda() -> dtype=torch.bool, device=x.device)
    models/cnn_transformer.py / resnet_only.py   pretrained=True -> pretrained=False   (no network for torchvision weights)

``models/checkpoint.py`` (mmcv checkpoint plumbing, 608 lines, unused when ``pretrained`` is empty) is replaced by the
two-function stub ``baseline/shims/ref_checkpoint_stub.py``.  Nothing under ``baseline/_ref`` is tracked by git; no
reference source enters the repository's history.  ``baseline.load()`` puts ``_ref`` and the shims on ``sys.path``.
"""
from __future__ import annotations

import argparse
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")

FILES = ["models/swin_transformer_v2.py", "models/decoder_v1.py", "models/decoder_v2.py", "models/model.py",
         "models/cnn_transformer.py", "models/resnet_only.py", "models/optimizer.py", "utils/criterion.py",
         "utils/metrics.py", "utils/util.py", "configs/config.yaml"]

# file -> [(old, new, expected occurrences)]
PATCHES = {
    "models/swin_transformer_v2.py": [(".to('cuda:0')", ".to(self.logit_scale.device)", 1)],
    "utils/util.py": [("torch.eye(3).cuda()", "torch.eye(3).to(rot_vector.device)", 1)],
    "models/cnn_transformer.py": [("dtype=torch.bool).cuda()", "dtype=torch.bool, device=x.device)", 1),
                                  ("pretrained=True", "pretrained=False", None)],
    "models/resnet_only.py": [("pretrained=True", "pretrained=False", None)],
}


def stage(src: str = "/root/reference", force: bool = False) -> str:
    marker = os.path.join(DST, ".staged")
    if os.path.exists(marker) and not force:
        return DST
    if not os.path.isfile(os.path.join(src, "models", "swin_transformer_v2.py")):
        raise FileNotFoundError(f"reference not found at {src}")
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for rel in FILES:
        with open(os.path.join(src, rel), "r") as f:
            text = f.read()
        for old, new, count in PATCHES.get(rel, []):
            n = text.count(old)
            if count is not None and n != count:
                raise RuntimeError(f"reference changed: {rel} holds {n} x {old!r}, expected {count}")
            text = text.replace(old, new)
        out = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        with open(out, "w") as f:
            f.write(text)
    for pkg in ("models", "utils"):
        open(os.path.join(DST, pkg, "__init__.py"), "w").close()
    shutil.copy(os.path.join(HERE, "shims", "ref_checkpoint_stub.py"), os.path.join(DST, "models", "checkpoint.py"))
    with open(marker, "w") as f:
        f.write(f"staged from {src}\n")
    return DST


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default=os.environ.get("B200SWIN_REFERENCE", "/root/reference"))
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    print(stage(a.src, a.force))
