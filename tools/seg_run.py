import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from difformer_b200 import ops
from oracle import difformer_oracle as O
gen = torch.Generator().manual_seed(5)
nn_ = torch.randint(10, 41, (8192,), generator=gen)
tot = int(nn_.sum())
qs, ks, vs = (t.cuda() for t in O.synthetic_qkv(tot, 1, 64, seed=6))
nn_d = nn_.cuda()
for _ in range(5):
    o = ops.segmented_full_attention(qs, ks, vs, "simple", nn_d)
torch.cuda.synchronize()
print(O.rel_err(o, O.segmented_simple_attention(qs.double().cpu(), ks.double().cpu(), vs.double().cpu(), nn_)))
