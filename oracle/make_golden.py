"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the UNMODIFIED reference
(`/root/reference/node classification/difformer.py`, `physical particle/difformer-v2.py`) under
`oracle/ref_shim.py`.  Run in the build container only:   python oracle/make_golden.py

The reference has no golden vectors of its own, so these fixtures -- outputs of
the reference code itself on seeded inputs -- are what pins both the oracle restatement and the
CUDA path.  torch 2.11.0 CPU, fp32, seeds listed per case.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_shim import load_reference_v1, load_reference_v2  # noqa: E402
from oracle.difformer_oracle import synthetic_graph, synthetic_qkv  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(d):
    return {k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def attention_cases(ref):
    cases = {}
    specs = [  # name, kernel, N, H, D, Hv, adversarial, seed
        ("simple_n161_h4_d64", "simple", 161, 4, 64, 4, False, 123),
        ("simple_n140_h4_d64_adv", "simple", 140, 4, 64, 4, True, 7),
        ("simple_n64_h1_d64", "simple", 64, 1, 64, 1, False, 11),
        ("simple_n129_h2_d32_hv1", "simple", 129, 2, 32, 1, True, 5),
        ("simple_n33_h3_d16", "simple", 33, 3, 16, 3, True, 9),
        ("sigmoid_n150_h1_d64", "sigmoid", 150, 1, 64, 1, False, 123),
        ("sigmoid_n101_h4_d64", "sigmoid", 101, 4, 64, 4, True, 3),
        ("sigmoid_n77_h2_d32_hv1", "sigmoid", 77, 2, 32, 1, False, 21),
    ]
    for name, kernel, n, h, d, hv, adv, seed in specs:
        q, k, v = synthetic_qkv(n, h, d, seed=seed, hv=hv, adversarial=adv)
        if kernel == "sigmoid":
            q, k = q * 0.3, k * 0.3
        q, k, v = (t.clone().requires_grad_(True) for t in (q, k, v))
        out = ref.full_attention_conv(q, k, v, kernel)
        g = torch.randn(out.shape, generator=torch.Generator().manual_seed(seed + 1000))
        out.backward(g)
        cases[name] = _np(dict(q=q, k=k, v=v, out=out, g=g, dq=q.grad, dk=k.grad, dv=v.grad))
        if kernel == "simple" and h == 1:  # :43 only broadcasts correctly for H == 1
            _, att = ref.full_attention_conv(q.detach(), k.detach(), v.detach(), kernel, output_attn=True)
            cases[name]["attn"] = att.numpy()
    return cases


def gcn_cases(ref):
    cases = {}
    # (a) undirected + self loops, no weights  (what main.py:72-79 feeds)
    n = 150
    ei = synthetic_graph(n, 400, seed=1)
    x = torch.randn(n, 4, 64, generator=torch.Generator().manual_seed(2))
    cases["gcn_undirected_selfloops"] = _np(dict(x=x, edge_index=ei, out=ref.gcn_conv(x, ei, None)))
    # (b) directed, duplicates, isolated nodes, sources with zero in-degree (-> inf -> 0), weights incl. 0/NaN/inf
    gen = torch.Generator().manual_seed(3)
    n = 97
    row = torch.randint(0, n, (500,), generator=gen)
    col = torch.randint(0, n // 2, (500,), generator=gen)      # upper half never a target: d=0 there
    row[:20], col[:20] = row[20:40], col[20:40]                # exact duplicates
    ei = torch.stack([row, col])
    w = torch.rand(500, generator=gen) * 2 - 0.5
    w[5], w[6], w[7] = 0.0, float("nan"), float("inf")
    x = torch.randn(n, 1, 32, generator=gen)
    cases["gcn_directed_weighted"] = _np(dict(x=x, edge_index=ei, edge_weight=w, out=ref.gcn_conv(x, ei, w)))
    cases["gcn_directed_unweighted"] = _np(dict(x=x, edge_index=ei, out=ref.gcn_conv(x, ei, None)))
    # (c) backward through gcn_conv wrt x
    xg = torch.randn(n, 2, 16, generator=gen).requires_grad_(True)
    out = ref.gcn_conv(xg, ei, w)
    g = torch.randn(out.shape, generator=gen)
    out.backward(g)
    cases["gcn_backward"] = _np(dict(x=xg, edge_index=ei, edge_weight=w, out=out, g=g, dx=xg.grad))
    return cases


def model_cases(ref):
    cases = {}
    specs = [  # name, ctor kwargs, N, C_in, C_out, n_pairs, edge weights?
        ("model_simple_cora_like", dict(num_layers=2, num_heads=1, kernel="simple", use_bn=True, use_residual=True,
                                        use_weight=False, use_graph=True), 120, 40, 7, 300, False),
        ("model_simple_h4_weight_source", dict(num_layers=2, num_heads=4, kernel="simple", use_bn=True, use_residual=True,
                                               use_weight=True, use_graph=True, graph_weight=0.3, use_source=True), 90, 24, 5, 200, True),
        ("model_simple_nograph_nobn", dict(num_layers=3, num_heads=2, kernel="simple", use_bn=False, use_residual=False,
                                           use_weight=True, use_graph=False, alpha=0.7), 75, 12, 3, 100, False),
        ("model_sigmoid_h2", dict(num_layers=2, num_heads=2, kernel="sigmoid", use_bn=True, use_residual=True,
                                  use_weight=True, use_graph=True), 83, 20, 4, 150, False),
    ]
    for name, kw, n, cin, cout, pairs, weighted in specs:
        torch.manual_seed(123)
        hid = 64
        m = ref.DIFFormer(cin, hid, cout, **kw)
        m.eval()
        gen = torch.Generator().manual_seed(77)
        x = torch.randn(n, cin, generator=gen)
        ei = synthetic_graph(n, pairs, seed=5)
        w = torch.rand(ei.shape[1], generator=gen) + 0.5 if weighted else None
        # train-mode-free backward: eval mode so dropout is the identity, grads are deterministic
        out = m(x, ei, w) if weighted else m(x, ei)
        loss = (out * torch.randn(out.shape, generator=gen)).sum()
        loss.backward()
        d = dict(x=x, edge_index=ei, out=out, hidden=hid, cin=cin, cout=cout)
        if weighted:
            d["edge_weight"] = w
        for k_, v_ in kw.items():
            d["cfg_" + k_] = v_
        for k_, v_ in m.state_dict().items():
            d["sd_" + k_] = v_
        for k_, p in m.named_parameters():
            if p.grad is not None:          # use_bn=False leaves the LayerNorms unused
                d["grad_" + k_] = p.grad
        cases[name] = _np(d)
    return cases


def v2_cases(ref2):
    cases = {}
    gen = torch.Generator().manual_seed(31)
    n_nodes = torch.tensor([5, 17, 1, 40, 23, 9])
    tot = int(n_nodes.sum())
    q, k, v = (torch.randn(tot, 1, 64, generator=gen) + 0.3 for _ in range(3))
    conv = ref2.TransConv(64, 64)
    q, k, v = (t.clone().requires_grad_(True) for t in (q, k, v))
    out = conv.full_attention(q, k, v, "simple", n_nodes)
    g = torch.randn(out.shape, generator=gen)
    out.backward(g)
    cases["v2_simple_segments"] = _np(dict(q=q, k=k, v=v, n_nodes=n_nodes, out=out, g=g, dq=q.grad, dk=k.grad, dv=v.grad))
    # multi-head variant (the class never uses it, the function supports it)
    q, k, v = (torch.randn(tot, 2, 32, generator=gen) for _ in range(3))
    cases["v2_simple_segments_h2"] = _np(dict(q=q, k=k, v=v, n_nodes=n_nodes,
                                              out=conv.full_attention(q, k, v, "simple", n_nodes)))
    # whole model, eval mode
    torch.manual_seed(123)
    m = ref2.DIFFormer_v2(16, 64, 3, num_layers=2, kernel="simple", use_graph=True)
    m.eval()
    x = torch.randn(tot, 16, generator=gen)
    # block-diagonal edges: ring inside each graph + self loops
    rows, cols, s = [], [], 0
    for n in n_nodes.tolist():
        idx = torch.arange(n) + s
        rows += [idx, idx.roll(1), idx]
        cols += [idx.roll(1), idx, idx]
        s += n
    ei = torch.stack([torch.cat(rows), torch.cat(cols)])
    out = m(x, ei, n_nodes)
    d = dict(x=x, edge_index=ei, n_nodes=n_nodes, out=out)
    for k_, v_ in m.state_dict().items():
        d["sd_" + k_] = v_
    cases["v2_model_simple"] = _np(d)
    # kernel='sigmoid' of the batched variant (difformer-v2.py:113-135): cross-graph attention between the nodes that share
    # a padded slot; appended last so that the random draws of the cases above stay what they were
    gen2 = torch.Generator().manual_seed(77)
    nn2 = torch.tensor([4, 9, 1, 12, 7])
    tot2 = int(nn2.sum())
    for name, h, d, scale in (("v2_sigmoid_segments", 1, 64, 0.3), ("v2_sigmoid_segments_h2", 2, 32, 0.5)):
        q, k, v = (torch.randn(tot2, h, d, generator=gen2) * s_ for s_ in (scale, scale, 1.0))
        q, k, v = (t.clone().requires_grad_(True) for t in (q, k, v))
        out = conv.full_attention(q, k, v, "sigmoid", nn2)
        g = torch.randn(out.shape, generator=gen2)
        out.backward(g)
        cases[name] = _np(dict(q=q, k=k, v=v, n_nodes=nn2, out=out, g=g, dq=q.grad, dk=k.grad, dv=v.grad))
    return cases


def main():
    os.makedirs(OUT, exist_ok=True)
    ref, ref2 = load_reference_v1(), load_reference_v2()
    groups = {"attention": attention_cases(ref), "gcn": gcn_cases(ref), "model": model_cases(ref),
              "v2": v2_cases(ref2)}
    for gname, cases in groups.items():
        flat = {}
        for cname, arrs in cases.items():
            for k, v in arrs.items():
                flat[f"{cname}/{k}"] = v
        path = os.path.join(OUT, f"{gname}.npz")
        np.savez_compressed(path, **flat)
        print(f"{path}: {len(cases)} cases, {os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
