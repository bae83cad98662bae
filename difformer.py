"""Drop-in replacement for the reference's `difformer` module.

The reference harnesses do `from difformer import *` (node classification/parse.py:2) or
`from difformer import DIFFormer_v2` (physical particle/parse.py:3).  Put this repo on
PYTHONPATH ahead of the task directory and they pick up the B200 kernels unchanged.
"""
from difformer_b200.module import DIFFormer, DIFFormerConv, DIFFormer_v2, TransConv  # noqa: F401
from difformer_b200.ops import full_attention_conv, gcn_conv  # noqa: F401

__all__ = ["DIFFormer", "DIFFormerConv", "DIFFormer_v2", "TransConv", "full_attention_conv", "gcn_conv"]
