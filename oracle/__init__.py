"""CPU oracle for the jclip-b200 hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain PyTorch-CPU **fp32** restatement of the reference's
algorithm for the one hot path this repo accelerates (CLIP ViT-B/32
``encode_image`` with LoRA on the attention projections -> MTA mode seeking ->
cosine / LP++ logits -> top-5).  Every function cites the reference
``file:line`` it follows (paths relative to the reference checkout).

PARITY PINNED AGAINST THE REFERENCE'S SOURCE, NOT AGAINST JITTOR.  The reference ships no tests,
golden vectors or known-answer fixtures, and all of its arithmetic lives in the third-party
dependency ``jittor==1.3.8.5`` (requirements.txt:1), which is neither vendored in the reference nor
installable offline -- so the reference cannot be run as shipped.  What pins this oracle instead:

* ``oracle/jt_shim.py`` + ``oracle/make_golden.py``: the reference's *own* Python source
  (jclip/model.py and jclip/mha.py imported as modules; ``LoRALayer`` / ``LinearLoRA`` /
  ``PlainMultiheadAttentionLoRA`` / ``apply_lora`` / ``solve_mta`` / ``Channel_LP`` /
  ``logit_normalize`` cut out of test.py and ood.py by ``ast``) is executed unmodified on top of a
  small torch-backed stand-in for the Jittor ops it calls; its outputs are committed as
  ``tests/golden/ref_on_shim.npz`` and ``tests/test_oracle_golden.py`` checks the restatement here
  against them (tower <= 2e-5, MTA <= 1e-6).  That pins control flow, operator order, layouts and
  parameter plumbing against the reference text; it does NOT pin Jittor's kernels, whose assumed
  semantics (LayerNorm eps / biased variance, ``argsort`` return order, unbiased ``std``) are
  written down in oracle/jt_shim.py and DESIGN.md section 3.  Read "parity" in this repo with
  that caveat.
* the shipped LoRA checkpoint ``lora_weights1/lora_weights.pkl``: its schema (nesting, key names,
  shapes, dtypes, metadata) is recorded in ``tests/golden/lora_pickle_schema.json``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package, and only as the checker or
the reported CPU baseline.  Nothing under ``jittor-clip-fewshot_b200/`` imports it; the product
path fails loudly when the CUDA library is missing.
"""
from .vit import vit_encode_image, merge_lora_into_state_dict, layer_norm  # noqa: F401
from .text import text_encode                                               # noqa: F401
from .mta import solve_mta, solve_mta_logits, cdist, gaussian_kernel        # noqa: F401
from .head import channel_lp, logit_normalize, fuse_scores, topk_lowest_index_first, pipeline_image  # noqa: F401
