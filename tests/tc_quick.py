"""Quick tcgen05 bring-up check (run under a short timeout on the GPU box before the full suite)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from difformer_b200 import ops
from oracle import difformer_oracle as O

for h, n in ((4, 16), (4, 100), (4, 5000), (4, 132534), (1, 7), (1, 33), (1, 5000), (1, 132534), (2, 129), (2, 70000)):
    q, k, v = O.synthetic_qkv(n, h, 64, seed=n, adversarial=True)
    qg, kg, vg = q.cuda(), k.cuda(), v.cuda()
    ops.set_simple_impl("tcgen05")
    p_tc, prep = ops.simple_partials(qg, kg, vg, with_prepared=True)
    torch.cuda.synchronize()
    want = O.simple_partials(q.double(), k.double(), v.double())
    sz = h * 4096
    S = p_tc[:sz].reshape(h, 64, 64)
    print(h, n, "S", O.rel_err(S, want["S"]), "z", O.rel_err(p_tc[sz:sz + 64 * h].reshape(h, 64), want["z"]),
          "u", O.rel_err(p_tc[sz + 64 * h:sz + 128 * h].reshape(h, 64), want["u"]), "sq", float(p_tc[sz + 128 * h]) / float(want["sq"]) - 1,
          "sk", float(p_tc[sz + 128 * h + 1]) / float(want["sk"]) - 1, flush=True)
    assert O.rel_err(S, want["S"]) < 1e-4 and O.rel_err(p_tc[sz:sz + 64 * h].reshape(h, 64), want["z"]) < 1e-5
    o_tc = ops.simple_apply(qg, p_tc, float(n), h, 64)
    o_pr = ops.simple_apply(qg, p_tc, float(n), h, 64, prepared=prep)
    torch.cuda.synchronize()
    print(n, "out", O.rel_err(o_tc, O.simple_apply(q.double(), want)), "out(prepared)", O.rel_err(o_pr, O.simple_apply(q.double(), want)),
          "prepared vs prologue", O.rel_err(o_pr, o_tc), flush=True)
    assert prep is not None and O.rel_err(o_pr, O.simple_apply(q.double(), want)) < 1e-4
    only_s = p_tc.clone(); only_s[sz:sz + 128 * h] = 0
    qS = ops.simple_apply(qg, only_s, float(n), h, 64) * n
    _, parts = O.simple_apply(q.double(), want, return_parts=True)
    print(n, "qS", O.rel_err(qS, parts["qS"]), flush=True)
print("tc quick ok")
