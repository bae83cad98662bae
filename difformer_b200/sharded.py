"""Row-sharded 'simple' propagation across GPUs: one process per GPU, nodes split
into contiguous row ranges, ONE all-reduce(sum) of the pass-1 partials per layer forward
(`dif_simple_partials_len()` floats = 67.6 KB at H=4, D=64, independent of N) and its mirror in
the backward.  No other collective exists on the path: 'sigmoid' and `gcn_conv` need all rows
and are replicas-only."""
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import ops


def shard_rows(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row range [begin, end) of `rank`: [r*N/G, (r+1)*N/G) with the remainder spread
    over the first ranks (sizes differ by at most one row)."""
    if world < 1 or not (0 <= rank < world) or n_total < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(n_total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allreduce_partials(partials: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum of the pass-1 partials over the process group (NCCL on GPUs; gloo in CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partials, op=dist.ReduceOp.SUM, group=group)
    return partials


class RowShardedAttention:
    """full_attention_conv(kernel='simple') on this rank's rows of a row-sharded graph."""

    def __init__(self, n_total: int, group=None):
        self.n_total, self.group = int(n_total), group

    def __call__(self, qs: torch.Tensor, ks: torch.Tensor, vs: torch.Tensor) -> torch.Tensor:
        return ops.full_attention_conv(qs, ks, vs, "simple", group=self.group, n_total=self.n_total)

    # the two passes separately (bench / overlap experiments)
    def reduce(self, qs, ks, vs) -> torch.Tensor:
        return allreduce_partials(ops.simple_partials(qs, ks, vs), self.group)

    def apply(self, qs, partials, hv: int, d: int) -> torch.Tensor:
        return ops.simple_apply(qs, partials, float(self.n_total), hv, d)
