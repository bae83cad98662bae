"""'sigmoid' forward: tcgen05 vs FFMA kernels -- parity against the fp64 oracle and timing (GPU box only)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import difformer_b200 as db
from difformer_b200 import ops
from oracle import difformer_oracle as O

dev = torch.device("cuda:0")


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


shapes = [(1, 1, 1, 1), (63, 200, 2, 2), (128, 128, 1, 1), (129, 257, 1, 1), (300, 1000, 4, 1), (2708, 2708, 1, 1), (2708, 2708, 4, 4), (10000, 10000, 1, 1)]
if os.environ.get("SIG_BIG"):
    shapes = [(10000, 10000, 1, 1)]
for (n, l, h, hv) in shapes:
    gen = torch.Generator().manual_seed(n + l)
    q = torch.randn(n, h, 64, generator=gen) * 0.3
    k = torch.randn(l, h, 64, generator=gen) * 0.3
    v = torch.randn(l, hv, 64, generator=gen)
    qd, kd, vd = q.to(dev), k.to(dev), v.to(dev)
    ref = O.sigmoid_attention(q.double(), k.double(), v.double()) if n * l <= 3000 * 3000 * 4 else None
    row = {"shape": [n, l, h, hv]}
    outs = {}
    for impl in os.environ.get("SIG_IMPLS", "generic,tcgen05").split(","):
        ops.set_sigmoid_impl(impl)
        out = db.full_attention_conv(qd, kd, vd, "sigmoid")
        torch.cuda.synchronize()
        outs[impl] = out
        if ref is not None:
            row[impl + "_err"] = float(O.rel_err(out.cpu(), ref))
        row[impl + "_us"] = round(timeit(lambda: db.full_attention_conv(qd, kd, vd, "sigmoid"), 10 if n > 5000 else 20), 1)
    if len(outs) == 2:
        row["tc_vs_generic"] = float(O.rel_err(outs["tcgen05"].cpu(), outs["generic"].cpu().double()))
        row["deterministic"] = bool(torch.equal(outs["tcgen05"], db.full_attention_conv(qd, kd, vd, "sigmoid")))
    print(json.dumps(row), flush=True)
ops.set_sigmoid_impl("auto")
