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


class PartialsExchange:
    """Peer-mapped buffers + the LL-push NVLink all-reduce (`dif_comm_*`, csrc/comm.cu, common.cuh).

    Every rank pushes each element of its partials straight into its peers' buffers as a 64-bit word {call number |
    fp32} and polls the words the peers pushed into its own buffer; ranks are added in rank order (bit-identical
    result on every rank).  No NCCL call, no fences, no host sync.  Either fused into the tail of the pass-1 kernel
    (`fused_reduce`, tcgen05 shapes) or as a stand-alone kernel (`allreduce`: backward partials, other shapes)."""

    def __init__(self, length: int, group, device: torch.device):
        import ctypes
        from ._lib import check, lib
        self.len, self.group, self.device = int(length), group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self._lib, self._check = lib, check
        nbytes = int(lib.dif_comm_buffer_bytes(self.len))
        with torch.cuda.device(device):
            base = ctypes.c_void_p()
            check(lib.dif_comm_alloc(ctypes.byref(base), nbytes), "dif_comm_alloc")
            self.base = base.value
            handle = ctypes.create_string_buffer(64)
            check(lib.dif_comm_export(self.base, handle), "dif_comm_export")
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
            ptrs = []
            for r, h in enumerate(handles):
                if r == self.rank:
                    ptrs.append(self.base)
                else:
                    pp = ctypes.c_void_p()
                    check(lib.dif_comm_open(ctypes.create_string_buffer(h, 64), ctypes.byref(pp)), "dif_comm_open")
                    ptrs.append(pp.value)
        self.ptrs = ptrs
        self.c_ptrs = (ctypes.c_void_p * self.world)(*ptrs)
        self.seq = 0
        self._flag = ctypes.c_int32(0)
        # per-(shape) buffers of the fused call are cached: the library never allocates and Python should not either
        self._cache = {}
        dist.barrier(group=group)

    # ---- watchdog ---------------------------------------------------------------------------------------------
    def failed(self) -> bool:
        """Non-blocking: True once a kernel of this rank (that has already run) gave up waiting for a peer, or a peer told
        this rank that it gave up.  Reads a pinned host flag -- no device synchronisation."""
        self._check(self._lib.dif_comm_status(self.base, ctypes_byref(self._flag)), "dif_comm_status")
        return bool(self._flag.value)

    def raise_if_failed(self) -> None:
        if self.failed():
            raise RuntimeError("difformer_b200: the NVLink exchange of the partials timed out on some rank (a peer died, hung or "
                               "never launched): every result since then is invalid.  Quiesce all ranks, call "
                               "PartialsExchange.reset() on each of them (or fall back to the NCCL group) and retry.")

    def timed_out(self) -> bool:
        """Synchronises the device, then reports `failed()`."""
        torch.cuda.synchronize(self.device)
        return self.failed()

    def reset(self) -> None:
        """Clear the watchdog state.  Collective: every rank must call it, with no exchange in flight."""
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        with torch.cuda.device(self.device):
            self._check(self._lib.dif_comm_reset(self.base), "dif_comm_reset")
        dist.barrier(group=self.group)

    # ---- collectives --------------------------------------------------------------------------------------------
    def fused_reduce(self, qs, ks, vs):
        """Pass 1 + all-reduce in one kernel (dif_simple_reduce_allreduce).  Returns (partials, prepared) summed
        over all ranks, or None when the shape is not a tcgen05 shape (caller falls back to `allreduce`)."""
        from . import ops
        N, L, H, Hv, M, D = ops._shapes(qs, ks, vs)
        lib = self._lib
        if ops._SIMPLE_IMPL == ops._lib.DIF_IMPL_GENERIC or int(lib.dif_simple_prepared_bytes(H, Hv, M, D)) == 0:
            return None
        self.raise_if_failed()
        partials = torch.empty(self.len, dtype=torch.float32, device=self.device)
        prepared = torch.empty(int(lib.dif_simple_prepared_bytes(H, Hv, M, D)), dtype=torch.uint8, device=self.device)
        ws = ops.workspace(self.device, int(lib.dif_simple_workspace_bytes(N, H, Hv, M, D)))
        self.seq += 1
        with torch.cuda.device(self.device):
            self._check(lib.dif_simple_reduce_allreduce(qs.data_ptr(), ks.data_ptr(), vs.data_ptr(), N, H, Hv, M, D, partials.data_ptr(),
                                                        prepared.data_ptr(), ws.data_ptr(), ws.numel(), self.c_ptrs, self.rank, self.world,
                                                        self.seq, torch.cuda.current_stream(self.device).cuda_stream),
                        "dif_simple_reduce_allreduce")
        return partials, prepared

    def allreduce(self, src: torch.Tensor) -> torch.Tensor:
        """out = sum over ranks of `src` (fp32 [len], contiguous, this device): one small kernel, LL push over NVLink."""
        if src.numel() != self.len or src.dtype != torch.float32 or not src.is_contiguous():
            raise ValueError("allreduce: `src` must be a contiguous float32 tensor of the exchange's length")
        self.raise_if_failed()
        out = torch.empty(self.len, dtype=torch.float32, device=self.device)
        self.seq += 1
        with torch.cuda.device(self.device):
            self._check(self._lib.dif_comm_allreduce(self.c_ptrs, self.rank, self.world, self.len, self.seq, src.data_ptr(), out.data_ptr(),
                                                     torch.cuda.current_stream(self.device).cuda_stream), "dif_comm_allreduce")
        return out


def ctypes_byref(x):
    import ctypes
    return ctypes.byref(x)


class RowShardComm:
    """Process group + lazily created peer-mapped exchanges (one per payload length).  Pass it as
    `group=` to `full_attention_conv(..., 'simple')`: the partials are then all-reduced by the
    one-shot NVLink kernel instead of NCCL."""

    def __init__(self, group=None):
        self.group = group if group is not None else dist.group.WORLD
        self._ex = {}

    def exchange(self, length: int, device) -> PartialsExchange:
        key = (int(length), str(device))
        if key not in self._ex:
            self._ex[key] = PartialsExchange(length, self.group, torch.device(device))
        return self._ex[key]


class RowShardedAttention:
    """full_attention_conv(kernel='simple') on this rank's rows of a row-sharded graph."""

    def __init__(self, n_total: int, group=None, nvlink: bool = False):
        """nvlink=True: all-reduce through peer-mapped memory (one-shot kernel) instead of NCCL."""
        self.n_total = int(n_total)
        self.group = RowShardComm(group) if (nvlink and dist.is_initialized() and dist.get_world_size(group) > 1) else group

    def __call__(self, qs: torch.Tensor, ks: torch.Tensor, vs: torch.Tensor) -> torch.Tensor:
        return ops.full_attention_conv(qs, ks, vs, "simple", group=self.group, n_total=self.n_total)

    # the two passes separately (bench / overlap experiments)
    def reduce(self, qs, ks, vs) -> torch.Tensor:
        if isinstance(self.group, RowShardComm):
            fused = self.group.exchange(ops.lib.dif_simple_partials_len(qs.shape[1], vs.shape[1], qs.shape[2], vs.shape[2]), qs.device).fused_reduce(qs, ks, vs)
            if fused is not None:
                return fused[0]
            ex = self.group.exchange(ops.lib.dif_simple_partials_len(qs.shape[1], vs.shape[1], qs.shape[2], vs.shape[2]), qs.device)
            return ex.allreduce(ops.simple_partials(qs, ks, vs))
        return allreduce_partials(ops.simple_partials(qs, ks, vs), self.group)

    def apply(self, qs, partials, hv: int, d: int) -> torch.Tensor:
        return ops.simple_apply(qs, partials, float(self.n_total), hv, d)


# ----------------------------------------------------------------------------------------------------------------
# module-level row sharding: DIFFormer / DIFFormerConv built by the reference harness (parse_method), run on a row shard
# ----------------------------------------------------------------------------------------------------------------
class _GatherRows(torch.autograd.Function):
    """x_local [n_r, ...] on every rank -> x_all [n_total, ...] (rows in rank order); backward = reduce-scatter(sum) of the
    gradient rows to their owners.  Shards differ by at most one row (shard_rows): padded to equal chunks for the collective."""

    @staticmethod
    def forward(ctx, x, group, n_total):
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        sizes = [shard_rows(n_total, r, world) for r in range(world)]
        chunk = max(e - b for b, e in sizes)
        pad = x.new_zeros((chunk,) + tuple(x.shape[1:]))
        pad[:x.shape[0]] = x
        out = x.new_empty((world * chunk,) + tuple(x.shape[1:]))
        dist.all_gather_into_tensor(out, pad, group=group)
        ctx.group, ctx.sizes, ctx.chunk, ctx.rank = group, sizes, chunk, rank
        if all(e - b == chunk for b, e in sizes):
            return out
        return torch.cat([out[r * chunk:r * chunk + (e - b)] for r, (b, e) in enumerate(sizes)], 0)

    @staticmethod
    def backward(ctx, g):
        world, chunk = len(ctx.sizes), ctx.chunk
        g = g.contiguous()
        padded = g.new_zeros((world * chunk,) + tuple(g.shape[1:]))
        for r, (b, e) in enumerate(ctx.sizes):
            padded[r * chunk:r * chunk + (e - b)] = g[b:e]
        mine = g.new_empty((chunk,) + tuple(g.shape[1:]))
        dist.reduce_scatter_tensor(mine, padded, op=dist.ReduceOp.SUM, group=ctx.group)
        b, e = ctx.sizes[ctx.rank]
        return mine[:e - b], None, None


def gather_rows(x: torch.Tensor, group, n_total: int) -> torch.Tensor:
    """All-gather of the row shards (autograd-aware); identity without a multi-rank group."""
    if group is None or not dist.is_initialized() or dist.get_world_size(group) < 2:
        return x
    return _GatherRows.apply(x, group, int(n_total))


class RowShard:
    """What a `DIFFormer` needs to run on rows [begin, end) of an n_total-node graph (one process per GPU):
      * 'simple' attention: pass 1 on the local rows, ONE all-reduce of the 67.6 KB partials (fused into the kernel over
        NVLink with `nvlink=True`, NCCL otherwise), pass 2 on the local rows -- SURVEY.md 8e;
      * `gcn_conv`: the target rows are local, their neighbours are anywhere: the value rows are all-gathered (T bytes over
        NVLink) and the SpMM runs over this rank's rows of the CSR of the GLOBAL `edge_index` (replicated, built once and
        cached); backward = transposed SpMM + reduce-scatter -- SURVEY.md 8f-2.
    `edge_index` handed to the model stays the global edge list; `x` is the local row block."""

    def __init__(self, n_total: int, group=None, nvlink: bool = True):
        self.n_total = int(n_total)
        self.pg = group if group is not None else (dist.group.WORLD if dist.is_initialized() else None)
        self.world = dist.get_world_size(self.pg) if self.pg is not None else 1
        self.rank = dist.get_rank(self.pg) if self.pg is not None else 0
        self.begin, self.end = shard_rows(self.n_total, self.rank, self.world)
        # what ops.full_attention_conv takes as `group=`
        self.attn_group = (RowShardComm(self.pg) if nvlink else self.pg) if self.world > 1 else None


def shard_model(model, n_total: int, group=None, nvlink: bool = True) -> RowShard:
    """Switch a `DIFFormer` (e.g. the one `parse_method` built) to row-sharded propagation: every rank then calls
    `model(x_local, edge_index_global[, edge_weight])` with its rows [shard.begin, shard.end) of x and gets its rows of the
    logits.  Parameters are replicated: wrap the model in DDP (or all-reduce the gradients) as usual.  Returns the RowShard."""
    shard = RowShard(n_total, group, nvlink)
    model._row_shard = shard
    for conv in getattr(model, "convs", []):
        conv._row_shard = shard
    return shard
