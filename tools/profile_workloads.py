#!/usr/bin/env python
"""Runs every kernel family of the library a few times at its BASELINE shape so that ONE ncu invocation can capture them all:
  ncu --set full --clock-control none --import-source on -k regex:'<names>' -o gpurun_out/r2_prof_all python tools/profile_workloads.py
(config A for 'simple' fp32 / bf16 / backward / two-pass / fused layer, N = 10 000 and Cora for 'sigmoid', config 5 for the batched
graphs, E = 17 N for gcn_conv)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import difformer
from difformer_b200 import ops
from oracle import difformer_oracle as O

dev = torch.device("cuda")
reps = int(os.environ.get("DIF_PROFILE_REPS", "2"))
n, h, d = 132534, 4, 64
q, k, v = (t.to(dev) for t in O.synthetic_qkv(n, h, d, seed=3))
for _ in range(reps):
    ops.simple_forward(q, k, v)                                        # simple_fused_kernel<4>
    ops.simple_forward(q.bfloat16(), k.bfloat16(), v.bfloat16())       # simple_lp_kernel<4, Bf16>
    part, prep = ops.simple_partials(q, k, v, with_prepared=True)      # reduce_tma_kernel<4, false>
    ops.simple_apply(q, part, float(n), h, d, prepared=prep)           # apply_tc_kernel<0, 4>
q1, k1, v1 = (t.to(dev) for t in O.synthetic_qkv(n, 1, 128, seed=4))
for _ in range(reps):
    ops.simple_forward(q1, k1, v1)                                     # simple_fused_kernel<2, true> (hidden 128)
qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
go = torch.randn(n, h, d, device=dev)
for _ in range(reps):
    qg.grad = kg.grad = vg.grad = None
    difformer.full_attention_conv(qg, kg, vg, "simple").backward(go)   # reduce_tma_kernel<4, true>, bwd_apply_tc_kernel<0|1|2, 4>
# fused layer (mode-1 epilogue with LayerNorm) + gcn
ei = O.synthetic_graph(n, 8 * n, seed=4).to(dev)
csr = ops.graph_csr(ei, None, n)
prev = torch.randn(n, d, device=dev)
lnw, lnb = torch.ones(d, device=dev), torch.zeros(d, device=dev)
for _ in range(reps):
    vb_ = torch.empty((n, d), dtype=torch.float32, device=dev)
    part, prep = ops.simple_partials(q, k, v, with_prepared=True, vbar=vb_)
    g = ops.spmm(csr, vb_.view(n, 1, d)).view(n, d)                    # spmm_kernel
    ep = ops.make_epilogue(0.5 / h, [(g, 0.5), (prev, 0.5)], layer_norm=(lnw, lnb, 1e-5))
    ops.simple_apply(q, part, float(n), h, d, ep, prepared=prep)       # layer_tc_kernel<4, false> (gcn as an addend)
    ep = ops.make_epilogue(0.5 / h, [(prev, 0.5)], layer_norm=(lnw, lnb, 1e-5), gcn=(csr, vb_, 0.5))
    ops.simple_apply(q, part, float(n), h, d, ep, prepared=prep)       # layer_tc_kernel<4, false> with the gcn gather in the epilogue
    ops.spmm(csr, v)                                                   # spmm_kernel, all heads
# whole layer from x with the projections folded into the propagation
from difformer_b200 import module as M_
torch.manual_seed(11)
conv = difformer.DIFFormerConv(d, d, num_heads=h, kernel="simple", use_graph=True, use_weight=True).to(dev)
ln = torch.nn.LayerNorm(d).to(dev)
x = torch.randn(n, d, device=dev)
for _ in range(reps):
    with torch.no_grad():                                              # reduce_tma_kernel<1, false> (Gram mode), project_head_kernel,
        M_._conv_forward(conv, x, x, ei, None, x, False, residual=(0.5, prev), layer_norm=ln)   # project_finish_kernel, apply_tc_kernel<1, false>, spmm_kernel, layer_tc_kernel<4, true>
# sigmoid: N = 10 000 (main-batch.py mini-batch) forward + backward on tcgen05, then the FFMA kernels
n2 = 10000
q2, k2, v2 = (t.to(dev) for t in O.synthetic_qkv(n2, 1, 64, seed=2))
q2, k2 = (q2 * 0.3).requires_grad_(True), (k2 * 0.3).requires_grad_(True)
v2.requires_grad_(True)
for impl in ("tcgen05", "generic"):
    ops.set_sigmoid_impl(impl)
    for _ in range(reps):
        q2.grad = k2.grad = v2.grad = None
        o = difformer.full_attention_conv(q2, k2, v2, "sigmoid")       # sigmoid_fwd_tc_kernel / sigmoid_fwd_kernel
        o.backward(torch.ones_like(o))                                 # sigmoid_bwd_tc_kernel<false|true> / sigmoid_dq|dkv_kernel
ops.set_sigmoid_impl("auto")
# batched graphs (config 5): B = 8192, n_g ~ U[10, 40]
gen = torch.Generator().manual_seed(5)
nn_ = torch.randint(10, 41, (8192,), generator=gen)
tot = int(nn_.sum())
qs, ks, vs = (t.to(dev).requires_grad_(True) for t in O.synthetic_qkv(tot, 1, 64, seed=6))
nn_d = nn_.to(dev)
for _ in range(reps):
    qs.grad = ks.grad = vs.grad = None
    o = ops.segmented_full_attention(qs, ks, vs, "simple", nn_d)       # seg_fwd_tc_kernel (tensor cores)
    o.backward(torch.ones_like(o))                                     # seg_bwd_warp_kernel, seg_bwd_fixup_kernel
ops.set_segmented_impl("generic")
for _ in range(reps):
    with torch.no_grad():
        ops.segmented_full_attention(qs, ks, vs, "simple", nn_d)       # seg_fwd_warp_kernel
ops.set_segmented_impl("auto")
# generic FFMA 'simple' (H = 3: no tensor-core shape)
q3, k3, v3 = (t.to(dev) for t in O.synthetic_qkv(40000, 3, 32, seed=7))
for _ in range(reps):
    difformer.full_attention_conv(q3, k3, v3, "simple")                # reduce_kernel, finalize_kernel, apply_kernel
torch.cuda.synchronize()
print("profile workloads done")
