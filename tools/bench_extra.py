#!/usr/bin/env python
"""Secondary measurements of SURVEY.md 8(d): the other rows of the hot path at their BASELINE shapes.
Prints one JSON object per line (saved under profiles/).  CUDA events, warm-up, median-free mean."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import difformer
from difformer_b200 import ops
from oracle import difformer_oracle as O

dev = torch.device("cuda")


def timeit(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


def emit(**kw):
    print(json.dumps(kw), flush=True)


# ---- 1. 'sigmoid' at Cora shape (BASELINE configs[1])
n, h, d = 2708, 1, 64
q, k, v = (t.to(dev) for t in O.synthetic_qkv(n, h, d, seed=1))
q, k = q * 0.3, k * 0.3
t_f = timeit(lambda: difformer.full_attention_conv(q, k, v, "sigmoid"))
qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
def fb():
    o = difformer.full_attention_conv(qg, kg, vg, "sigmoid"); o.backward(torch.ones_like(o))
t_fb = timeit(fb, 20)
def ref_sig():
    p = torch.sigmoid(torch.einsum("nhm,lhm->nlh", q, k)); r = p.sum(1, keepdim=True); return torch.einsum("nlh,lhd->nhd", p / r, v)
t_ref = timeit(ref_sig)
emit(row="a-2 sigmoid fwd", shape="Cora N=2708 H=1 D=64", us=t_f, node_updates_per_s=n / t_f * 1e6, flops=4.0 * n * n * h * d,
     tflops=4.0 * n * n * h * d / t_f / 1e6, torch_gpu_chain_us=t_ref, fwd_bwd_us=t_fb)
n2 = 10000
q2, k2, v2 = (t.to(dev) for t in O.synthetic_qkv(n2, 1, 64, seed=2))
t_f2 = timeit(lambda: difformer.full_attention_conv(q2 * .3, k2 * .3, v2, "sigmoid"), 10, 2)
emit(row="a-2 sigmoid fwd", shape="N=10000 H=1 D=64 (main-batch.py mini-batch size)", us=t_f2, tflops=4.0 * n2 * n2 * 64 / t_f2 / 1e6)

# ---- 2. gcn_conv at config A with E = 16N random + N self loops
n, h, d = 132534, 4, 64
q, k, v = (t.to(dev) for t in O.synthetic_qkv(n, h, d, seed=3))
ei = O.synthetic_graph(n, 8 * n, seed=4).to(dev)
E = ei.shape[1]
t_build = timeit(lambda: ops.GraphCSR(ei, None, n), 5, 1)
csr = ops.graph_csr(ei, None, n)
t_spmm = timeit(lambda: ops.spmm(csr, v))
t_hm = timeit(lambda: ops.spmm(csr, v, head_mean=True))
gather = E * h * d * 4
emit(row="a-3 gcn_conv", shape=f"N={n} E={E} H=4 D=64", csr_build_us=t_build, spmm_us=t_spmm, spmm_head_mean_us=t_hm,
     gather_bytes=gather, spmm_gather_gbs=gather / t_spmm / 1e3, note="CSR is cached per edge_index; the reference rebuilds it every forward")

# ---- 3. fused propagation layer (attention + gcn + head mean + residual), no grad: what DIFFormerConv does after the Linears
prev = torch.randn(n, d, device=dev)
t_vbar = timeit(lambda: ops.head_mean(v))
vb = ops.head_mean(v)
t_spmm1 = timeit(lambda: ops.spmm(csr, vb.view(n, 1, d)))
emit(row="a-3 gcn_conv via mean_h(V)", shape=f"N={n} E={E}", head_mean_us=t_vbar, spmm_h1_us=t_spmm1)
def layer():
    vb_ = torch.empty((n, d), dtype=torch.float32, device=dev)
    part, prep = ops.simple_partials(q, k, v, with_prepared=True, vbar=vb_)
    g = ops.spmm(csr, vb_.view(n, 1, d)).view(n, d)
    ep = ops.make_epilogue(0.5 / h, [(g, 0.5), (prev, 0.5)])
    return ops.simple_apply(q, part, float(n), h, d, ep, prepared=prep)
t_layer = timeit(layer)
alg = n * (3 * h * d * 4 + 2 * d * 4) + E * 8 + (n + 1) * 4
emit(row="a-4/a-5 fused layer", shape=f"N={n} E={E} H=4 D=64", us=t_layer, node_updates_per_s=n / t_layer * 1e6,
     algorithmic_bytes=alg, hbm_frac_of_6571=alg / t_layer / 1e3 / 6571.2)

# ---- 4. 'simple' forward+backward at config A (backward = FFMA kernels)
qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
go = torch.randn(n, h, d, device=dev)
def fb2():
    o = difformer.full_attention_conv(qg, kg, vg, "simple"); o.backward(go)
t_fb2 = timeit(fb2, 20)
emit(row="a-1b simple fwd+bwd", shape=f"N={n} H=4 D=64", us=t_fb2, node_updates_per_s=n / t_fb2 * 1e6)

# ---- 5. batched graphs (BASELINE configs[4]): B=8192, n_g ~ U[10,40], H=1, D=64
gen = torch.Generator().manual_seed(5)
nn_ = torch.randint(10, 41, (8192,), generator=gen)
tot = int(nn_.sum())
qs, ks, vs = (t.to(dev) for t in O.synthetic_qkv(tot, 1, 64, seed=6))
nn_d = nn_.to(dev)
t_seg = timeit(lambda: ops.segmented_full_attention(qs, ks, vs, "simple", nn_d))
qsg, ksg, vsg = (t.clone().requires_grad_(True) for t in (qs, ks, vs))
gs = torch.randn(tot, 1, 64, device=dev)
def fb3():
    o = ops.segmented_full_attention(qsg, ksg, vsg, "simple", nn_d); o.backward(gs)
t_seg_fb = timeit(fb3, 20)
emit(row="a-6 segmented simple", shape=f"B=8192 sumN={tot} H=1 D=64", fwd_us=t_seg, fwd_bwd_us=t_seg_fb, node_updates_per_s=tot / t_seg * 1e6,
     note="reference pads with Python loops: 164 ms per make_batch call at B=8192 (SURVEY.md 3.4)")

# ---- 6. Cora-shaped model forward (BASELINE configs[0]/[1] shape), fused inference path
n, cin = 2708, 1433
x = torch.randn(n, cin, device=dev)
eic = O.synthetic_graph(n, 5278, seed=7).to(dev)
for kern in ("simple", "sigmoid"):
    m = difformer.DIFFormer(cin, 64, 7, num_layers=2, num_heads=1, kernel=kern, use_bn=True, use_residual=True, use_graph=True, use_weight=False).to(dev).eval()
    with torch.no_grad():
        t_m = timeit(lambda: m(x, eic), 50)
    row = dict(row="model forward", shape=f"Cora-like N=2708 C=1433 2 layers hidden 64 H=1 kernel={kern}", us=t_m)
    try:
        from difformer_b200 import GraphedForward
        with torch.no_grad():
            want = m(x, eic).clone()
        gf = GraphedForward(m, x, eic)
        row["cuda_graph_us"] = timeit(lambda: gf(x, eic), 50)
        row["cuda_graph_max_abs_diff"] = float((gf(x, eic) - want).abs().max())
    except Exception as exc:  # noqa: BLE001
        row["cuda_graph_error"] = repr(exc)[:300]
    emit(**row)
