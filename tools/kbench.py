#!/usr/bin/env python
"""Times the individual C-ABI calls of the 'simple' path at config A with CUDA events (kernel tuning aid).
   python tools/kbench.py [--n 132534] [--iters 200] [--impl auto]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from difformer_b200 import ops
from oracle import difformer_oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=132534)
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--impl", default="auto")
ap.add_argument("--tag", default="")
ap.add_argument("--h", type=int, default=4)
ap.add_argument("--noref", action="store_true")
ap.add_argument("--fused", action="store_true", help="also time the one-kernel forward (dif_simple_forward)")
ap.add_argument("--only-fused", action="store_true", help="time nothing but the one-kernel forward")
ap.add_argument("--dtype", default="f32", choices=["f32", "bf16", "f16"])
ap.add_argument("--membw", action="store_true", help="plain-torch read / write / copy bandwidth of this GPU")
a = ap.parse_args()
ops.set_simple_impl(a.impl)
H = a.h
q, k, v = (t.cuda() for t in O.synthetic_qkv(a.n, H, 64, seed=1))
T = a.n * H * 64 * 4


def timeit(fn, iters):
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


if a.membw:
    g = torch.empty(1 << 28, device="cuda").normal_(); g2 = torch.empty_like(g)
    t_w = timeit(lambda: g2.zero_(), 20)
    t_r = timeit(lambda: g.sum(), 20)
    t_c = timeit(lambda: g2.copy_(g), 20)
    print(f"1 GiB fp32: memset {g.numel()*4/t_w/1e3:5.0f} GB/s | sum (read) {g.numel()*4/t_r/1e3:5.0f} GB/s | copy {2*g.numel()*4/t_c/1e3:5.0f} GB/s (read+write)", flush=True)
    sys.exit(0)
if a.only_fused or a.dtype != "f32":
    dt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[a.dtype]
    q, k, v = q.to(dt), k.to(dt), v.to(dt)
    assert ops.simple_forward(q, k, v) is not None
    t_f = timeit(lambda: ops.simple_forward(q, k, v), a.iters)
    Tb = T * q.element_size() // 4
    print(f"{a.tag} dtype={a.dtype} H={H} n={a.n} one-kernel forward {t_f:7.1f} us  roofline(4T) {4*Tb/t_f/1e3/6571.2:5.3f}", flush=True)
    sys.exit(0)
part, prep = ops.simple_partials(q, k, v, with_prepared=True)
t_red = timeit(lambda: ops.simple_partials(q, k, v, with_prepared=True), a.iters)
t_app = timeit(lambda: ops.simple_apply(q, part, float(a.n), H, 64, prepared=prep), a.iters)


def op():
    pp, pr = ops.simple_partials(q, k, v, with_prepared=True)
    return ops.simple_apply(q, pp, float(a.n), H, 64, prepared=pr)


t_op = timeit(op, a.iters)
print(f"{a.tag} impl={a.impl} H={H} n={a.n} reduce+finalize {t_red:7.1f} us ({3*T/t_red/1e3:6.0f} GB/s)  apply {t_app:7.1f} us ({2*T/t_app/1e3:6.0f} GB/s)  "
      f"op {t_op:7.1f} us  roofline(4T) {4*T/t_op/1e3/6571.2:5.3f}", flush=True)

if a.fused:
    assert ops.simple_forward(q, k, v) is not None
    t_f = timeit(lambda: ops.simple_forward(q, k, v), a.iters)
    print(f"{a.tag} one-kernel forward {t_f:7.1f} us  roofline(4T) {4*T/t_f/1e3/6571.2:5.3f}", flush=True)
if a.noref:
    sys.exit(0)
# ---- plain-torch streaming references on the same box (what the memory system gives a trivial kernel)
big = torch.empty(3 * a.n * 256, device="cuda").normal_()
t_sum = timeit(lambda: big.sum(), 50)
src, dst = torch.empty(a.n * 256, device="cuda").normal_(), torch.empty(a.n * 256, device="cuda")
t_cp = timeit(lambda: dst.copy_(src), 50)
g = torch.empty(1 << 28, device="cuda").normal_(); g2 = torch.empty_like(g)
t_big = timeit(lambda: g2.copy_(g), 20)
print(f"  torch.sum over 3T: {t_sum:6.1f} us ({3*T/t_sum/1e3:5.0f} GB/s) | copy T->T: {t_cp:6.1f} us ({2*T/t_cp/1e3:5.0f} GB/s) | copy 1GiB: {2*g.numel()*4/t_big/1e3:5.0f} GB/s", flush=True)
