"""In-kernel timelines (DIF_TC_DEBUG_TIMES=1) of the projected layer's tcgen05 kernels: run under
   DIF_TC_DEBUG_TIMES=1 python tools/dbg_layer.py   (stderr: min/avg/max per stamp over the CTAs)."""
import os, sys
os.environ.setdefault("DIF_TC_DEBUG_TIMES", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import difformer
from difformer_b200 import ops, module as M_
from oracle import difformer_oracle as O
dev = torch.device("cuda", 0)
n, h, d = 132534, 4, 64
torch.manual_seed(11)
conv = difformer.DIFFormerConv(d, d, num_heads=h, kernel="simple", use_graph=True, use_weight=True).to(dev)
ln = torch.nn.LayerNorm(d).to(dev)
x, prev = torch.randn(n, d, device=dev), torch.randn(n, d, device=dev)
ei = O.synthetic_graph(n, 8 * n, seed=4).to(dev)
ops.set_projection_folding(True)
for i in range(3):
    print(f"--- iteration {i}", file=sys.stderr, flush=True)
    with torch.no_grad():
        M_._conv_forward(conv, x, x, ei, None, x, False, residual=(0.5, prev), layer_norm=ln)
torch.cuda.synchronize()
