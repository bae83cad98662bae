"""Which part of the layer epilogue costs what: pass 2 on x (projected operands) with the epilogue pieces switched on one by one.
   python tools/layer_ablate.py            (CUDA-event times, 50 iterations each)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import difformer
from difformer_b200 import ops, projected
dev = torch.device("cuda", 0)
n, h, d = 132534, int(os.environ.get("H", "4")), 64
torch.manual_seed(11)
conv = difformer.DIFFormerConv(d, d, num_heads=h, kernel="simple", use_graph=True, use_weight=True).to(dev)
ln = torch.nn.LayerNorm(d).to(dev)
x, a0, a1 = (torch.randn(n, d, device=dev) for _ in range(3))
with torch.no_grad():
    vpart, nvec, vbar_part = projected.projected_operands(projected.gram(x), float(n), conv)
lnp = (ln.weight.detach(), ln.bias.detach(), ln.eps)


def timeit(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


cases = {
    "mode 0 (per-head output [N,H,64])": None,
    "mode 1, head mean only": ops.make_epilogue(1.0 / h, []),
    "mode 1 + 1 addend": ops.make_epilogue(1.0 / h, [(a0, 0.5)]),
    "mode 1 + 2 addends": ops.make_epilogue(1.0 / h, [(a0, 0.5), (a1, 0.5)]),
    "mode 1 + LayerNorm": ops.make_epilogue(1.0 / h, [], layer_norm=lnp),
    "mode 1 + 2 addends + LayerNorm": ops.make_epilogue(1.0 / h, [(a0, 0.5), (a1, 0.5)], layer_norm=lnp),
}
for name, ep in cases.items():
    print(f"{timeit(lambda: projected.apply(x, vpart, nvec, h, ep)):8.1f} us  {name}", flush=True)
print(f"{timeit(lambda: projected.head_mean_values(x, vbar_part, nvec, h)):8.1f} us  mean_h V (one head, mode 0)", flush=True)
print(f"{timeit(lambda: projected.gram(x)):8.1f} us  Gram pass", flush=True)
gp = projected.gram(x)
print(f"{timeit(lambda: projected.projected_operands(gp, float(n), conv)):8.1f} us  projection kernels", flush=True)
