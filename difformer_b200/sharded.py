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
