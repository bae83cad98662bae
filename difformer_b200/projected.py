"""The Wq / Wk / Wv projections folded into the 'simple' propagation (node classification/difformer.py:115-140).

`DIFFormerConv.forward` computes Q = x Wq^T + bq, K = x Wk^T + bk, V = x Wv^T + bv from the SAME layer input x [N, C] (difformer.py
:195: `conv(x, x, ...)`) and hands the three [N, H, 64] tensors to `full_attention_conv`.  Everything pass 1 of 'simple' extracts from
K, V and Q is a row reduction, and row reductions of affine images of x follow from the Gram matrix G = X^T X and the column sums
s = X^T 1 of x itself (one pass over x: 256 B per node instead of 3 KB):

    S_h = K_h^T V_h = Wk_h G Wv_h^T + (Wk_h s) bv_h^T + bk_h (Wv_h s)^T + N bk_h bv_h^T         [M, D]
    z_h = Wk_h s + N bk_h          u_h = Wv_h s + N bv_h
    sum k^2 = <Wk G, Wk> + 2 bk.(Wk s) + N |bk|^2          (sum q^2 alike with Wq, bq)

and pass 2 needs q_h only inside the products q_h S_h and q_h . z_h, which are affine in x too:

    q_h S_h = x (Wq_h^T S_h) + bq_h^T S_h            q_h . z_h = x . (Wq_h^T z_h) + bq_h . z_h

So a layer runs as  [pass 1 on x: G, s]  ->  [dif_simple_project: a few 64 x 64 products per head, fp64]  ->  [pass 2 on x with the
projected operands, dif_simple_apply_projected]  and Q, K, V are never formed, written or read: 256 B (pass 1) + 256 B (pass 2) of
input per node instead of 3 KB + 1 KB, and the three Linear GEMMs disappear.  Inference path (no autograd), hidden = 64, H in
{1, 2, 4}; `use_weight=False` (V = x, one head) is the case Wv = I, bv = 0.
"""
import ctypes
from typing import Optional

import torch

from . import ops
from ._lib import check, lib

HID = 64


def supported(conv, query_input: torch.Tensor, source_input: torch.Tensor) -> bool:
    """Same input for Q and K/V, hidden size 64 in and out, a tensor-core head count, fp32 CUDA rows."""
    return (query_input is source_input and query_input.dim() == 2 and query_input.shape[1] == HID and conv.out_channels == HID
            and conv.num_heads in (1, 2, 4) and query_input.dtype == torch.float32 and query_input.is_cuda
            and ops._SIMPLE_IMPL != ops._lib.DIF_IMPL_GENERIC and conv.Wq.in_features == HID
            and conv.Wq.bias is not None and conv.Wk.bias is not None and (not conv.use_weight or conv.Wv.bias is not None)
            and (not query_input.is_contiguous() or query_input.data_ptr() % 32 == 0))


def gram(x: torch.Tensor, shard=None) -> torch.Tensor:
    """One pass over x [N, 64] -> the partials of (H = Hv = 1, 64, 64): [X^T X | X^T 1 | X^T 1 | sum x^2 | sum x^2] -- pass 1 of 'simple'
    with q = k = v = x (tcgen05 kernel in Gram mode, deterministic).  Additive over row shards: with `shard` (sharded.RowShard) the sum
    over the ranks is taken inside the same kernel over NVLink (RowShardComm) or by NCCL (plain process group)."""
    x3 = x.view(-1, 1, HID)
    if shard is None or shard.world < 2:
        return ops.simple_partials(x3, x3, x3)
    from . import sharded
    grp = shard.attn_group
    if isinstance(grp, sharded.RowShardComm):
        ex = grp.exchange(lib.dif_simple_partials_len(1, 1, HID, HID), x.device)
        fused = ex.fused_reduce(x3, x3, x3)
        return fused[0] if fused is not None else ex.allreduce(ops.simple_partials(x3, x3, x3))
    return sharded.allreduce_partials(ops.simple_partials(x3, x3, x3), grp)


def _w(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


def projected_operands(gram_partials: torch.Tensor, n_total: float, conv, with_values: bool = True):
    """The pass-2 operands of every head from the Gram partials and the layer's weights (dif_simple_project: fp64 arithmetic, two small
    launches): (vpartials fp32 in the partials layout of (H, Hv = H, 64, 64), nvec fp32 [H + 1], vbar_partials fp32 [4226] or None: the
    one-head pass-2 problem whose solution is mean_h V, with nvec[H:] as its denominator constant).
    The same algebra in torch fp64, which the tests check this against: oracle.difformer_oracle.projected_operands."""
    H = conv.num_heads
    dev = gram_partials.device
    Wq, bq, Wk, bk = _w(conv.Wq.weight), _w(conv.Wq.bias), _w(conv.Wk.weight), _w(conv.Wk.bias)
    Wv, bv = (_w(conv.Wv.weight), _w(conv.Wv.bias)) if conv.use_weight else (None, None)
    nv = H * HID * HID + 2 * H * HID + 2
    nb = (HID * HID + 2 * HID + 2) if with_values else 0
    buf = torch.empty(nv + nb + (-(nv + nb)) % 8 + H + 1, dtype=torch.float32, device=dev)      # one allocation for the outputs
    vpart, vbar_part, nvec = buf[:nv], (buf[nv:nv + nb] if with_values else None), buf[buf.numel() - (H + 1):]
    ws = ops.workspace(dev, lib.dif_simple_project_workspace_bytes(H))
    with torch.cuda.device(dev):
        check(lib.dif_simple_project(gram_partials.data_ptr(), Wq.data_ptr(), bq.data_ptr(), Wk.data_ptr(), bk.data_ptr(),
                                     None if Wv is None else Wv.data_ptr(), None if bv is None else bv.data_ptr(), float(n_total), H,
                                     vpart.data_ptr(), nvec.data_ptr(), None if vbar_part is None else vbar_part.data_ptr(), ws.data_ptr(),
                                     ws.numel(), ops._stream(gram_partials)), "dif_simple_project")
    return vpart, nvec, vbar_part


def value_operands(conv, device):
    """(vbar_partials, one): the one-head pass-2 problem whose solution is mean_h V = x wbar^T + bbar, from the weights alone
    (dif_simple_project_values) -- the value branch of a layer (mean_h V -> SpMM) does not depend on the Gram matrix."""
    key = (conv.Wv.weight.data_ptr(), conv.Wv.weight._version, conv.Wv.bias.data_ptr(), conv.Wv.bias._version, str(device))
    hit = getattr(conv, "_value_operands", None)
    if hit is not None and hit[0] == key and not torch.cuda.is_current_stream_capturing():
        return hit[1], hit[2]                  # inference: the weights do not change between calls (in-place updates bump _version)
    Wv, bv = _w(conv.Wv.weight), _w(conv.Wv.bias)
    buf = torch.empty(HID * HID + 2 * HID + 2 + 6 + 1, dtype=torch.float32, device=device)
    vbar_part, one = buf[:HID * HID + 2 * HID + 2], buf[-1:]
    with torch.cuda.device(device):
        check(lib.dif_simple_project_values(Wv.data_ptr(), bv.data_ptr(), conv.num_heads, vbar_part.data_ptr(), one.data_ptr(),
                                            torch.cuda.current_stream(device).cuda_stream), "dif_simple_project_values")
    if not torch.cuda.is_current_stream_capturing():
        object.__setattr__(conv, "_value_operands", (key, vbar_part, one))
    return vbar_part, one


def head_mean_values(x: torch.Tensor, vbar_part: torch.Tensor, nvec: torch.Tensor, H: int) -> torch.Tensor:
    """mean_h V = x wbar^T + bbar [N, 64] through the pass-2 kernel (one head, A = x): the input of the gcn term.  `nvec`: the [H + 1]
    vector of projected_operands (its last entry is the constant) with H, or the `one` of value_operands with H = 0."""
    return apply(x, vbar_part, nvec[H:], 1).view(-1, HID)


_SIDE_STREAMS: dict = {}


def side_stream(device) -> torch.cuda.Stream:
    """One extra stream per device for the value branch of the folded layer (created on first use, outside any graph capture)."""
    key = str(device)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


def apply(x: torch.Tensor, vpart: torch.Tensor, nvec: torch.Tensor, H: int, epilogue: Optional[ops.Epilogue] = None, keep=()) -> torch.Tensor:
    """Pass 2 on x with the projected operands -> [N, H, 64] (epilogue None) or [N, 64] (mode-1 layer epilogue)."""
    N = x.shape[0]
    fused = epilogue is not None and epilogue.mode == 1
    out = torch.empty((N, HID) if fused else (N, H, HID), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.dif_simple_apply_projected(x.data_ptr(), x.stride(0), vpart.data_ptr(), nvec.data_ptr(), N, H, out.data_ptr(),
                                             ctypes.byref(epilogue) if epilogue is not None else None, ops._stream(x)),
              "dif_simple_apply_projected")
    del keep
    return out


def attention(x: torch.Tensor, conv, n_total: Optional[float] = None) -> torch.Tensor:
    """full_attention_conv(Wq x, Wk x, Wv x, 'simple') -> [N, H, 64] without forming Q, K, V (no autograd)."""
    x = x.contiguous()
    vpart, nvec, _ = projected_operands(gram(x), float(x.shape[0] if n_total is None else n_total), conv)
    return apply(x, vpart, nvec, conv.num_heads)
