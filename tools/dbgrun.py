import sys, os
sys.path.insert(0, "/root/repo")
import torch
from difformer_b200 import ops
from oracle import difformer_oracle as O
q, k, v = (t.cuda() for t in O.synthetic_qkv(132534, 4, 64, seed=1))
for i in range(3):
    p, pr = ops.simple_partials(q, k, v, with_prepared=True)
    o = ops.simple_apply(q, p, 132534.0, 4, 64, prepared=pr)
torch.cuda.synchronize()
