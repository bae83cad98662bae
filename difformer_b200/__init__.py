"""difformer_b200 -- B200 (sm_100a) kernels for DIFFormer's diffusion-attention propagation.

Drop-in for the reference's `difformer` module (see ../difformer.py): `DIFFormer`,
`DIFFormerConv`, `full_attention_conv`, `gcn_conv`, plus `DIFFormer_v2` / `TransConv`.
"""
from .ops import (full_attention_conv, gcn_conv, segmented_full_attention, GraphCSR, graph_csr,  # noqa: F401
                  simple_partials, simple_apply, simple_forward, set_simple_impl, set_fused_forward, set_projection_folding, subgraph)
from .module import DIFFormer, DIFFormerConv, DIFFormer_v2, TransConv, GraphedForward  # noqa: F401
from .sharded import (RowShardedAttention, RowShardComm, PartialsExchange, shard_rows, RowShard, shard_model,  # noqa: F401
                      gather_rows)

__version__ = "0.1.0"
