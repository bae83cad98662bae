"""CPU, world_size 2, gloo: the multi-GPU path of 'simple' is ONE all-reduce of the pass-1
partials.  The partials of each row shard come from the oracle here (no GPU); the
all-reduce goes through the product's `allreduce_partials`, packed in the C-ABI layout
[S | z | u | sum q^2 | sum k^2]; pass 2 on each shard must reproduce the unsharded result."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import difformer_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _pack(p):
    return torch.cat([p["S"].reshape(-1), p["z"].reshape(-1), p["u"].reshape(-1), p["sq"].reshape(1), p["sk"].reshape(1)]).contiguous()


def _unpack(flat, H, Hv, M, D, n):
    o = 0
    S = flat[o:o + H * M * D].reshape(H, M, D); o += H * M * D
    z = flat[o:o + H * M].reshape(H, M); o += H * M
    u = flat[o:o + Hv * D].reshape(Hv, D); o += Hv * D
    return {"S": S, "z": z, "u": u, "sq": flat[o], "sk": flat[o + 1], "n": torch.tensor(float(n))}


def _worker(rank, world, port, n, h, hv, d, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from difformer_b200.sharded import allreduce_partials, shard_rows
        q, k, v = O.synthetic_qkv(n, h, d, seed=5, hv=hv, adversarial=True, dtype=torch.float64)
        b, e = shard_rows(n, rank, world)
        flat = _pack(O.simple_partials(q[b:e], k[b:e], v[b:e]))
        assert flat.numel() == h * d * d + h * d + hv * d + 2
        allreduce_partials(flat)
        out = O.simple_apply(q[b:e], _unpack(flat, h, hv, d, d, n), n_total=n)
        torch.save({"span": (b, e), "out": out}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,h,hv,d", [(257, 4, 4, 16), (100, 2, 1, 8)])
def test_row_sharded_equals_unsharded(tmp_path, n, h, hv, d):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, n, h, hv, d, str(tmp_path)), nprocs=world, join=True)
    q, k, v = O.synthetic_qkv(n, h, d, seed=5, hv=hv, adversarial=True, dtype=torch.float64)
    want = O.simple_attention(q, k, v)
    got = torch.empty_like(want)
    for r in range(world):
        blob = torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))
        b, e = blob["span"]
        got[b:e] = blob["out"]
    assert O.rel_err(got, want) < 1e-12


def _gather_worker(rank, world, port, n, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from difformer_b200.sharded import RowShard, gather_rows, shard_model, shard_rows
        x = torch.arange(n * 6, dtype=torch.float64).reshape(n, 2, 3)
        b, e = shard_rows(n, rank, world)
        xl = x[b:e].clone().requires_grad_(True)
        full = gather_rows(xl, dist.group.WORLD, n)
        assert torch.equal(full.detach(), x)                       # rows arrive in rank order, padding removed
        w = torch.arange(n, dtype=torch.float64).reshape(n, 1, 1) + 1.0 + rank      # a different weighting on every rank
        (full * w).sum().backward()                                # d/dx_local = sum over ranks of their weights on my rows
        want = sum(torch.arange(n, dtype=torch.float64)[b:e] + 1.0 + r for r in range(world)).reshape(-1, 1, 1).expand(-1, 2, 3)
        assert torch.equal(xl.grad, want)
        # module plumbing: shard_model marks the model and every conv with the same RowShard
        import difformer
        m = difformer.DIFFormer(8, 16, 3, num_layers=2)
        sh = shard_model(m, n, dist.group.WORLD, nvlink=False)
        assert isinstance(sh, RowShard) and (sh.begin, sh.end) == (b, e) and all(c._row_shard is sh for c in m.convs)
        torch.save({"ok": True}, os.path.join(out_dir, f"g{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 11])
def test_gather_rows_and_its_gradient(tmp_path, n):
    """Row all-gather of the sharded gcn_conv: forward order / padding and backward = reduce-scatter(sum)."""
    world, port = 2, _free_port()
    mp.spawn(_gather_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.isfile(os.path.join(str(tmp_path), f"g{r}.pt")) for r in range(world))
