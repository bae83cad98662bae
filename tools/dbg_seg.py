"""Timeline of CTA 0 of the tensor-core batched-graph forward:  DIF_SEG_DEBUG=1 python tools/dbg_seg.py"""
import os, sys
os.environ.setdefault("DIF_SEG_DEBUG", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from difformer_b200 import ops
from oracle import difformer_oracle as O
gen = torch.Generator().manual_seed(5)
nn_ = torch.randint(10, 41, (8192,), generator=gen)
tot = int(nn_.sum())
q, k, v = (t.cuda().requires_grad_(True) for t in O.synthetic_qkv(tot, 1, 64, seed=6))
nd = nn_.cuda()
for i in range(3):
    print(f"--- call {i}", file=sys.stderr, flush=True)
    o = ops.segmented_full_attention(q, k, v, "simple", nd)
    o.backward(torch.ones_like(o))
torch.cuda.synchronize()
