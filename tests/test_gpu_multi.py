"""GPU, >= 2 devices: row-sharded 'simple' over NCCL == unsharded result (skipped on 1-GPU boxes)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["DIF_ROOT"])
from difformer_b200.sharded import RowShardedAttention, shard_rows
from difformer_b200 import ops
from oracle import difformer_oracle as O
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
n, h, d = 20011, 4, 64
q, k, v = O.synthetic_qkv(n, h, d, seed=11, adversarial=True)
b, e = shard_rows(n, rank, world)
want = O.simple_attention(q.double(), k.double(), v.double())
g = torch.randn(n, h, d, generator=torch.Generator().manual_seed(5))
dq, dk, dv = O.simple_attention_backward(q.double(), k.double(), v.double(), g.double())
outs = {}
for nvlink in (False, True):          # NCCL all-reduce, then the one-shot NVLink kernel over peer-mapped memory
    attn = RowShardedAttention(n, dist.group.WORLD, nvlink=nvlink)
    for rep in range(3):              # several calls: exercises the alternating slots / sequence numbers
        qs, ks, vs = (t[b:e].cuda().requires_grad_(True) for t in (q, k, v))
        out = attn(qs, ks, vs)
        out.backward(g[b:e].cuda())
        errs = [O.rel_err(out, want[b:e]), O.rel_err(qs.grad, dq[b:e]), O.rel_err(ks.grad, dk[b:e]), O.rel_err(vs.grad, dv[b:e])]
        assert max(errs) < 1e-3, (nvlink, rep, errs)
    outs[nvlink] = out.detach()
assert O.rel_err(outs[True], outs[False]) < 1e-5
# watchdog: rank 0 launches an exchange its peer never joins -> the kernel gives up after 30 s instead of hanging the
# GPU, and the status word says so
from difformer_b200.sharded import PartialsExchange
ex = PartialsExchange(1024, dist.group.WORLD, torch.device("cuda", rank))
assert not ex.timed_out()
slot = ex.next_slot()
slot.fill_(float(rank + 1))
if rank == 0:
    ex.allreduce(slot)
    assert ex.timed_out()
else:
    assert not ex.timed_out()
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok", errs)
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_row_sharded_nccl_matches_unsharded(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, DIF_ROOT=ROOT)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-3000:]
    assert res.stdout.count(" ok ") == 2
