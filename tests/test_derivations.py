"""CPU: the closed forms the CUDA kernels implement, restated in a few lines of fp64 torch and checked against the
oracle (which is itself pinned to the reference by tests/golden/).  These are the derivations a reader needs to trust
before reading the kernels: csrc/segmented.cu (warp-per-graph direct form + t fix-up), csrc/sigmoid_sm100.cu
(paired reciprocal, hi/lo operand split, [hi | lo] accumulate) and csrc/sigmoid_bwd_sm100.cu (e P^2 form)."""
import math

import torch

from oracle import difformer_oracle as O


def _direct_segment_forward(q, k, v, c):
    w = 1.0 + c * (q @ k.T)                      # [n, n] weights of one graph, one head
    return (w @ v) / w.sum(1, keepdim=True), w


def test_segmented_direct_form_equals_s_form_forward_and_backward():
    """seg_fwd_warp_kernel / seg_bwd_warp_kernel + seg_bwd_fixup_kernel, per graph and head:
       out_n = sum_l w_nl v_l / d_n, w_nl = 1 + c q_n.k_l;   dw = (g.v - g.out)/d;   dq = c dw k - q t/|Q|^2, ...
       with ONE batch-wide t = sum dw c q.k."""
    gen = torch.Generator().manual_seed(0)
    n_nodes = torch.tensor([5, 17, 1, 40, 23])
    tot = int(n_nodes.sum())
    q, k, v = (t.double() for t in O.synthetic_qkv(tot, 2, 64, seed=3, adversarial=True))
    g = torch.randn(tot, 2, 64, generator=gen, dtype=torch.float64)
    sq, sk = (q * q).sum(), (k * k).sum()
    c = 1.0 / math.sqrt(float(sq) * float(sk))
    out = torch.empty_like(v)
    dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
    t = 0.0
    s = 0
    for n in n_nodes.tolist():
        for h in range(2):
            qq, kk, vv, gg = (x[s:s + n, h] for x in (q, k, v, g))
            o, w = _direct_segment_forward(qq, kk, vv, c)
            out[s:s + n, h] = o
            d = w.sum(1, keepdim=True)
            dw = (gg @ vv.T - (gg * o).sum(1, keepdim=True)) / d
            dq[s:s + n, h] = c * dw @ kk
            dk[s:s + n, h] = c * dw.T @ qq
            dv[s:s + n, h] = (w / d).T @ gg
            t += float((dw * c * (qq @ kk.T)).sum())
        s += n
    dq -= q * (t / sq)                            # the fix-up kernel: c = (|Q|^2 |K|^2)^-1/2 is batch-wide
    dk -= k * (t / sk)
    assert O.rel_err(out, O.segmented_simple_attention(q, k, v, n_nodes)) < 1e-12
    want = O.segmented_simple_attention_backward(q, k, v, n_nodes, g)
    for got, w64 in ((dq, want[0]), (dk, want[1]), (dv, want[2])):
        assert O.rel_err(got, w64) < 1e-10


def test_sigmoid_paired_reciprocal_and_e_p2_forms():
    """sigmoid_sm100.cu evaluates two sigmoids with one reciprocal; the backward draft uses P (1 - P) = e P^2."""
    s = torch.linspace(-42.9, 60.0, 2001, dtype=torch.float64)    # the clamp x <= 62 is s >= -42.97
    x = -s * math.log2(math.e)                    # the kernel's score: S = (-log2 e Q) K^T
    e = torch.exp2(torch.clamp(x, max=62.0))
    a, b = 1.0 + e[0::2][:1000], 1.0 + e[1::2][:1000]
    r = 1.0 / (a * b)
    p0, p1 = b * r, a * r
    assert torch.allclose(p0, torch.sigmoid(s[0::2][:1000]), rtol=1e-12, atol=0)
    assert torch.allclose(p1, torch.sigmoid(s[1::2][:1000]), rtol=1e-12, atol=0)
    assert torch.isfinite(a * b).all() and float((a * b).max()) < 3.4e38       # the clamp keeps the fp32 product finite
    p = 1.0 / (1.0 + e)
    assert torch.allclose(e * p * p, torch.sigmoid(s) * torch.sigmoid(-s), rtol=1e-12, atol=0)


def _bf16_split(x):
    hi = x.to(torch.bfloat16).to(torch.float32)
    lo = (x - hi).to(torch.bfloat16).to(torch.float32)
    return hi, lo


def test_bf16_hi_lo_split_products_reach_fp32_accuracy():
    """Three bf16 products hi.hi + lo.hi + hi.lo with fp32 accumulation (what the tcgen05 kernels issue) reproduce an
    fp32 GEMM to ~1e-5, while hi.hi alone (plain bf16) is ~3e-3: the reason every operand, P included, is split."""
    gen = torch.Generator().manual_seed(1)
    a = torch.randn(128, 64, generator=gen) * 0.3
    b = torch.randn(128, 64, generator=gen) * 0.3
    ref = a.double() @ b.double().T
    ah, al = _bf16_split(a)
    bh, bl = _bf16_split(b)
    three = (ah @ bh.T + al @ bh.T + ah @ bl.T).double()
    one = (ah @ bh.T).double()
    assert O.rel_err(three, ref) < 2e-5
    assert O.rel_err(one, ref) > 1e-3
    # [hi | lo] accumulate of the P V product: p_hi [v_hi | v_lo] + p_lo [v_hi | v_lo], halves added in the epilogue
    p = torch.rand(128, 128, generator=gen)
    v = torch.randn(128, 64, generator=gen)
    ph, pl = _bf16_split(p)
    vh, vl = _bf16_split(v)
    acc = ph @ torch.cat([vh, vl], 1) + pl @ torch.cat([vh, vl], 1)
    assert O.rel_err((acc[:, :64] + acc[:, 64:]).double(), p.double() @ v.double()) < 2e-5


def test_sigmoid_backward_closed_form():
    """dS = (g.v - g.out)/r * P (1 - P);  dV = (P/r)^T g;  dQ = dS K;  dK = dS^T Q   (sigmoid.cu and the tcgen05 draft)."""
    gen = torch.Generator().manual_seed(2)
    n, l = 37, 53
    q = torch.randn(n, 1, 64, generator=gen, dtype=torch.float64) * 0.3
    k = torch.randn(l, 1, 64, generator=gen, dtype=torch.float64) * 0.3
    v = torch.randn(l, 1, 64, generator=gen, dtype=torch.float64)
    g = torch.randn(n, 1, 64, generator=gen, dtype=torch.float64)
    p = torch.sigmoid(q[:, 0] @ k[:, 0].T)
    r = p.sum(1, keepdim=True)
    out = (p / r) @ v[:, 0]
    ds = (g[:, 0] @ v[:, 0].T - (g[:, 0] * out).sum(1, keepdim=True)) / r * p * (1 - p)
    want = O.sigmoid_attention_backward(q, k, v, g)
    assert O.rel_err((ds @ k[:, 0]).unsqueeze(1), want[0]) < 1e-12
    assert O.rel_err((ds.T @ q[:, 0]).unsqueeze(1), want[1]) < 1e-12
    assert O.rel_err(((p / r).T @ g[:, 0]).unsqueeze(1), want[2]) < 1e-12
