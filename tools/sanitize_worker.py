#!/usr/bin/env python
"""Small battery for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family once, at sizes a sanitized run
finishes in seconds.  Single process:  compute-sanitizer --tool racecheck python tools/sanitize_worker.py
Two ranks (cross-GPU LL exchange):     compute-sanitizer --target-processes all --tool memcheck \
                                         python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/sanitize_worker.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import difformer
from difformer_b200 import ops
from difformer_b200.sharded import RowShardedAttention, shard_rows
from oracle import difformer_oracle as O

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
worst = 0.0


def chk(name, got, want, tol=1e-3):
    global worst
    e = O.rel_err(got.cpu(), want)
    worst = max(worst, e)
    assert e < tol, (name, e)


if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    n, h, d = 3001, 4, 64
    q, k, v = O.synthetic_qkv(n, h, d, seed=3, adversarial=True)
    g = torch.randn(n, h, d, generator=torch.Generator().manual_seed(1))
    want = O.simple_attention(q.double(), k.double(), v.double())
    dq, dk, dv = O.simple_attention_backward(q.double(), k.double(), v.double(), g.double())
    b, e = shard_rows(n, rank, world)
    attn = RowShardedAttention(n, dist.group.WORLD, nvlink=True)
    for rep in range(2):
        qs, ks, vs = (t[b:e].to(dev).requires_grad_(True) for t in (q, k, v))
        out = attn(qs, ks, vs)
        out.backward(g[b:e].to(dev))
        chk("sharded out", out, want[b:e])
        chk("sharded dq", qs.grad, dq[b:e])
        chk("sharded dk", ks.grad, dk[b:e])
        chk("sharded dv", vs.grad, dv[b:e])
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
else:
    for (n, h, hv, d) in ((1500, 4, 4, 64), (700, 1, 1, 64), (333, 3, 1, 32)):
        q, k, v = O.synthetic_qkv(n, h, d, seed=n, hv=hv, adversarial=True)
        g = torch.randn(n, h, d, generator=torch.Generator().manual_seed(1))
        qs, ks, vs = (t.to(dev).requires_grad_(True) for t in (q, k, v))
        out = difformer.full_attention_conv(qs, ks, vs, "simple")
        out.backward(g.to(dev))
        chk("simple out", out, O.simple_attention(q.double(), k.double(), v.double()))
        dq, dk, dv = O.simple_attention_backward(q.double(), k.double(), v.double(), g.double())
        chk("simple dq", qs.grad, dq); chk("simple dk", ks.grad, dk); chk("simple dv", vs.grad, dv)
    n, h, d = 400, 2, 64
    q, k, v = O.synthetic_qkv(n, h, d, seed=9)
    q, k = q * 0.3, k * 0.3
    qs, ks, vs = (t.to(dev).requires_grad_(True) for t in (q, k, v))
    out = difformer.full_attention_conv(qs, ks, vs, "sigmoid")
    out.sum().backward()
    chk("sigmoid out", out, O.sigmoid_attention(q.double(), k.double(), v.double()))
    ei = O.synthetic_graph(n, 1500, seed=2).to(dev)
    chk("gcn", difformer.gcn_conv(v.to(dev), ei, None), O.gcn_conv(v.double(), ei.cpu(), None))
    nn_ = torch.tensor([3, 40, 1, 25, 64, 70, 17])
    tot = int(nn_.sum())
    q, k, v = O.synthetic_qkv(tot, 1, 64, seed=4)
    qs, ks, vs = (t.to(dev).requires_grad_(True) for t in (q, k, v))
    out = ops.segmented_full_attention(qs, ks, vs, "simple", nn_.to(dev))
    out.sum().backward()
    chk("segmented", out, O.segmented_simple_attention(q.double(), k.double(), v.double(), nn_))
    torch.cuda.synchronize()
print("sanitize worker rank", rank, "ok", worst, flush=True)
