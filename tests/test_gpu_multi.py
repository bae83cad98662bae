"""GPU, >= 2 devices: the row-sharded 'simple' path == the unsharded fp64 oracle, on 2 / 4 / 8 ranks (each world size is
skipped when the box has fewer GPUs).  Covers, at HEAD: the all-reduce fused into the pass-1 kernel tail
(`dif_simple_reduce_allreduce`, LL push over NVLink), the stand-alone `dif_comm_allreduce` (backward partials, generic
shapes), the NCCL baseline, row counts that do not divide by the world size, repeated calls (slot alternation), the
reduced partials themselves, and the watchdog."""
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
from difformer_b200.sharded import RowShardedAttention, RowShardComm, shard_rows
from difformer_b200 import ops
from oracle import difformer_oracle as O
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
worst = 0.0
# (n, h, hv, d): tcgen05 shapes (fused exchange) and shapes on the generic path (stand-alone exchange); n never divides
for (n, h, hv, d) in ((20011, 4, 4, 64), (4999, 1, 1, 64), (3001, 3, 3, 32), (2503, 2, 1, 64)):
    q, k, v = O.synthetic_qkv(n, h, d, seed=11 + n, adversarial=True)
    if hv != h:
        v = v[:, :1].contiguous()
    b, e = shard_rows(n, rank, world)
    want = O.simple_attention(q.double(), k.double(), v.double())
    wp = O.simple_partials(q.double(), k.double(), v.double())
    g = torch.randn(n, h, d, generator=torch.Generator().manual_seed(5))
    dq, dk, dv = O.simple_attention_backward(q.double(), k.double(), v.double(), g.double())
    outs = {}
    for nvlink in (False, True):          # NCCL all-reduce, then the LL-push exchange over peer-mapped memory
        attn = RowShardedAttention(n, dist.group.WORLD, nvlink=nvlink)
        for rep in range(3):              # several calls: exercises the alternating slots / call numbers
            qs, ks, vs = (t[b:e].to(dev).requires_grad_(True) for t in (q, k, v))
            out = attn(qs, ks, vs)
            out.backward(g[b:e].to(dev))
            errs = [O.rel_err(out, want[b:e]), O.rel_err(qs.grad, dq[b:e]), O.rel_err(ks.grad, dk[b:e]), O.rel_err(vs.grad, dv[b:e])]
            assert max(errs) < 1e-3, (n, h, hv, d, nvlink, rep, errs)
            worst = max(worst, max(errs))
        outs[nvlink] = out.detach()
        # the reduced partials themselves (S, z, u, sum q^2, sum k^2) against the oracle of the GLOBAL problem
        flat = attn.reduce(qs.detach(), ks.detach(), vs.detach()).double().cpu()
        nS, nz = h * d * d, h * d
        perr = [O.rel_err(flat[:nS].reshape(h, d, d), wp["S"]), O.rel_err(flat[nS:nS + nz].reshape(h, d), wp["z"]),
                O.rel_err(flat[nS + nz:nS + nz + hv * d].reshape(hv, d), wp["u"]),
                abs(float(flat[-2]) - float(wp["sq"])) / float(wp["sq"]), abs(float(flat[-1]) - float(wp["sk"])) / float(wp["sk"])]
        assert max(perr) < 1e-4, (n, h, hv, d, nvlink, perr)
        if nvlink:
            # bit-identical on every rank (fixed rank order of the sum)
            mine = attn.reduce(qs.detach(), ks.detach(), vs.detach())
            ref = mine.clone()
            dist.broadcast(ref, 0)
            assert torch.equal(mine, ref), "reduced partials differ between ranks"
            assert not attn.group.exchange(flat.numel(), dev).timed_out()
    assert O.rel_err(outs[True], outs[False]) < 1e-5
# module level (SURVEY 8b/8e/8f-2): a DIFFormer as parse_method builds it, switched to row-sharded propagation by shard_model --
# attention through the fused exchange, gcn_conv through the row all-gather -- against the same model run unsharded
import copy
import difformer
from difformer_b200.sharded import shard_model
torch.manual_seed(0)
n, cin = 3001, 24
x = torch.randn(n, cin)
ei = O.synthetic_graph(n, 9000, seed=4).to(dev)
ew = torch.rand(ei.shape[1], generator=torch.Generator().manual_seed(2)).to(dev)
m = difformer.DIFFormer(cin, 64, 5, num_layers=2, num_heads=4, kernel="simple", dropout=0.0, use_graph=True, use_source=True).to(dev)
ref = copy.deepcopy(m)
out_ref = ref(x.to(dev), ei, ew)
out_ref.square().sum().backward()
for nvlink in (False, True):
    ms = copy.deepcopy(m)
    sh = shard_model(ms, n, dist.group.WORLD, nvlink=nvlink)
    out = ms(x[sh.begin:sh.end].to(dev), ei, ew)
    assert O.rel_err(out, out_ref[sh.begin:sh.end]) < 1e-4, (nvlink, O.rel_err(out, out_ref[sh.begin:sh.end]))
    out.square().sum().backward()
    for (name, p), (_, pr) in zip(ms.named_parameters(), ref.named_parameters()):
        g = p.grad.clone()
        dist.all_reduce(g)                                   # what DDP would do with the replicated parameters
        # the gradients of the Wq / Wk biases are sums of dq / dk over all rows, which cancel almost completely (a constant shift of q or
        # k barely changes the normalised attention): both the sharded and the unsharded fp32 value carry a relative error of ~1e-3 there
        tol = 2e-2 if (name.endswith("bias") and (".Wk." in name or ".Wq." in name)) else 1e-3
        assert O.rel_err(g, pr.grad) < tol, (nvlink, name, O.rel_err(g, pr.grad))
# inference (no autograd): the sharded layers run with the projections folded into the propagation (SURVEY 8f-1 x 8e): Gram partials
# all-reduced in the kernel, mean_h V rows all-gathered for the SpMM -- against the unsharded model with the folding on and off
with torch.no_grad():
    ops.set_projection_folding(False)
    out_plain = ref.eval()(x.to(dev), ei, ew)
    ops.set_projection_folding(True)
    out_fold = ref(x.to(dev), ei, ew)
    assert O.rel_err(out_fold, out_plain) < 1e-4, ("folded vs explicit, unsharded", O.rel_err(out_fold, out_plain))
    for nvlink in (False, True):
        ms = copy.deepcopy(m).eval()
        sh = shard_model(ms, n, dist.group.WORLD, nvlink=nvlink)
        out = ms(x[sh.begin:sh.end].to(dev), ei, ew)
        assert O.rel_err(out, out_plain[sh.begin:sh.end]) < 1e-4, ("folded sharded", nvlink, O.rel_err(out, out_plain[sh.begin:sh.end]))
# kernel='sigmoid' row-sharded: the query rows are sharded, K and V all-gathered (autograd: reduce-scatter of dK, dV)
torch.manual_seed(1)
n, cin = 1501, 16
x = torch.randn(n, cin)
ei = O.synthetic_graph(n, 4000, seed=9).to(dev)
m = difformer.DIFFormer(cin, 64, 4, num_layers=2, num_heads=2, kernel="sigmoid", dropout=0.0, use_graph=True).to(dev)
ref = copy.deepcopy(m)
out_ref = ref(x.to(dev), ei)
out_ref.square().sum().backward()
ms = copy.deepcopy(m)
sh = shard_model(ms, n, dist.group.WORLD)
out = ms(x[sh.begin:sh.end].to(dev), ei)
assert O.rel_err(out, out_ref[sh.begin:sh.end]) < 1e-4, ("sigmoid sharded", O.rel_err(out, out_ref[sh.begin:sh.end]))
out.square().sum().backward()
for (name, p), (_, pr) in zip(ms.named_parameters(), ref.named_parameters()):
    g = p.grad.clone()
    dist.all_reduce(g)
    assert O.rel_err(g, pr.grad) < 2e-3, ("sigmoid sharded", name, O.rel_err(g, pr.grad))
# batched graphs (difformer-v2) sharded by whole graphs: only the two norms (forward) and (t_q, t_k) (backward) cross ranks
gen = torch.Generator().manual_seed(21)
nn_all = torch.randint(1, 90, (64 * world + 3,), generator=gen)
tot = int(nn_all.sum())
q, k, v = O.synthetic_qkv(tot, 1, 64, seed=8, adversarial=True)
g = torch.randn(tot, 1, 64, generator=gen)
want = O.segmented_simple_attention(q.double(), k.double(), v.double(), nn_all)
dq, dk, dv = O.segmented_simple_attention_backward(q.double(), k.double(), v.double(), nn_all, g.double())
gb, ge = shard_rows(nn_all.numel(), rank, world)            # graphs [gb, ge) live on this rank
rb, re_ = int(nn_all[:gb].sum()), int(nn_all[:ge].sum())
for seg_impl in ("auto", "tcgen05"):      # warp-per-graph kernels (graphs of up to 89 nodes), then the tensor-core tiles: two-phase backward either way
    ops.set_segmented_impl(seg_impl)
    qs, ks, vs = (t[rb:re_].to(dev).requires_grad_(True) for t in (q, k, v))
    out = ops.segmented_full_attention(qs, ks, vs, "simple", nn_all[gb:ge].to(dev), group=dist.group.WORLD)
    out.backward(g[rb:re_].to(dev))
    errs = [O.rel_err(out, want[rb:re_]), O.rel_err(qs.grad, dq[rb:re_]), O.rel_err(ks.grad, dk[rb:re_]), O.rel_err(vs.grad, dv[rb:re_])]
    assert max(errs) < 1e-3, ("segmented", seg_impl, errs)
ops.set_segmented_impl("auto")
dist.barrier()
if world == 2 and os.environ.get("DIF_TEST_WATCHDOG", "1") == "1":
    # watchdog: rank 0 launches an exchange its peer never joins -> the kernel gives up (DIF_COMM_TIMEOUT_MS) instead of
    # hanging the GPU, raises its host flag AND marks the peer's status word: the peer learns at its next call
    from difformer_b200.sharded import PartialsExchange
    ex = PartialsExchange(1024, dist.group.WORLD, dev)
    assert not ex.timed_out()
    src = torch.full((1024,), float(rank + 1), device=dev)
    if rank == 0:
        ex.allreduce(src)
        assert ex.timed_out()
        try:
            ex.allreduce(src)
            raise SystemExit("expected the failed exchange to raise")
        except RuntimeError:
            pass
    dist.barrier()
    if rank == 1:
        ex.allreduce(src)                 # sees rank 0's mark in its status word and raises its own host flag
        assert ex.timed_out()
    ex.reset()
    assert not ex.failed()
    ex.seq = 10                           # both ranks restart from the same call number
    got = ex.allreduce(src)
    assert torch.allclose(got, torch.full_like(got, 3.0)) and not ex.timed_out()
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok", worst)
"""


def _run(world, tmp_path, extra_env=None):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, DIF_ROOT=ROOT, DIF_COMM_TIMEOUT_MS="1500")
    env.update(extra_env or {})
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    if res.returncode != 0 or res.stdout.count(" ok ") != world:
        try:                                   # keep the workers' output where a remote run brings it back
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", f"multi_worker_world{world}.log"), "w") as f:
                f.write(res.stdout)
        except OSError:
            pass
    assert res.returncode == 0, res.stdout[-6000:]
    assert res.stdout.count(" ok ") == world


@pytest.mark.parametrize("world", [2, 4, 8])
def test_row_sharded_matches_unsharded(world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs >= {world} GPUs")
    _run(world, tmp_path)
