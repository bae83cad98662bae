import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, difformer
from oracle import difformer_oracle as O
n, h = 132534, 4
q, k, v = (t.cuda().requires_grad_(True) for t in O.synthetic_qkv(n, h, 64, seed=1))
g = torch.randn(n, h, 64, device="cuda")
for _ in range(6):
    o = difformer.full_attention_conv(q, k, v, "simple"); o.backward(g)
torch.cuda.synchronize()
