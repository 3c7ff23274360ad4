"""CPU oracle for the Swin-V2 shifted-window attention + SiLog hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker (or as the
CPU baseline being reported), never as the thing measured or shipped.  The product
package (``multi-modal-monodepth-estimation_b200``, importable as ``b200swin``)
never imports this package and raises if its CUDA library is missing.

What is here
------------
* ``index_maps``  - numpy, integer-exact restatement of the reference's window
  partition / reverse / cyclic roll / pad-crop / shift-mask / relative-position-index
  maps (``models/swin_transformer_v2.py:120-147, 249-259, 429-463, 874-892``).
* ``swin_ref``    - functional torch-CPU restatement (fp32 or fp64) of
  ``WindowAttention.forward`` (``:275-336``), the post/pre-norm blocks (``:419-488``,
  ``:561-630``), ``Mlp`` (``:76-89``), ``LayerNormFP32`` (``:41-47``), ``PatchMerging``
  (``:648-678``), ``PatchEmbed`` (``:941-957``), ``BasicLayer.forward`` (``:866-908``)
  and ``SwinTransformerV2.forward`` (``:1251-1277``), driven by a reference
  ``state_dict``; plus the hand-derived backward of SURVEY appendix A.
* ``mha_ref``     - ``torch.nn.MultiheadAttention`` (batch_first, no masks) and the reference's
  ``Transformer_Encoder`` layer built on it (``models/cnn_transformer.py:176-216``); ``swin_ref`` also restates the
  ``attn_type='normal'`` / learned-bias-table / ``ConvMlp`` branches of the Swin file (``:92-117, 241-244, 296-298``).
* ``silog_ref``   - ``SiLogLoss`` (``utils/criterion.py:15-21``) with its closed-form
  gradient, and the ``eval_depth`` metrics (``utils/metrics.py:9-32``).

Pinning
-------
The reference ships no tests or golden vectors, so the oracle is
pinned against the reference ITSELF: ``tests/golden/make_golden.py`` imports the
unmodified reference modules from ``/root/reference`` in the build container (with
import shims for the absent ``timm``/``mmcv`` and the one-line device fix for
``swin_transformer_v2.py:294``), runs them on seeded inputs and commits the
input/weight/output/gradient tensors under ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` checks every oracle function against those files.
"""
