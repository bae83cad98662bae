"""nn.Module shell of the reference model, kept name-for-name so the reference harnesses and
checkpoints work unchanged:

  DIFFormerConv / DIFFormer   <- node classification/difformer.py:81-226
  TransConv / DIFFormer_v2    <- physical particle/difformer-v2.py:48-223

Submodule names (`convs.{i}.Wq/Wk/Wv`, `fcs.{0,1}`, `bns.{i}`) and ctor signatures are the
reference's; no extra parameters or persistent buffers are added, so `state_dict()` round-trips
with reference checkpoints (test_large_dataset.py:86-88).  Linear / LayerNorm / dropout stay
PyTorch (cuBLAS); the propagation between them runs in libdifformer_b200.so.

Two execution paths through a layer:
  * grad needed  : unfused ops (`full_attention_conv`, `gcn_conv`) with hand-written CUDA backward;
  * no grad      : pass 2 of 'simple' carries the whole layer epilogue (head mean, gcn term,
                   x_0, residual blend, difformer.py:129-140,200-201) -- one write of [N,D].
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops, projected


def _fusable(kernel, *tensors):
    if kernel != "simple":
        return False
    if not torch.is_grad_enabled():
        return True
    return not any(t is not None and t.requires_grad for t in tensors)


def _conv_forward_projected(conv, x, edge_index, edge_weight, x_0, residual, layer_norm, shard=None):
    """No-grad layer with the Wq / Wk / Wv projections folded into the propagation (projected.py): one pass over x for
    the Gram matrix, a few 64 x 64 products, one pass over x that writes the finished [N, 64] row (head mean, gcn term, x_0, residual
    blend, LayerNorm in the epilogue).  Q, K and V are never formed.
    Row-sharded (`shard`, sharded.shard_model): the Gram partials (4226 floats) are additive over the shards -- the same in-kernel NVLink
    all-reduce as the explicit path, 16 x smaller -- the operand algebra is replicated, mean_h V rows are all-gathered for the SpMM over
    this rank's target rows, pass 2 is local."""
    H = conv.num_heads
    x = ops._f32c(x)
    N = x.shape[0]
    n_total = N if shard is None else shard.n_total
    alpha = 1.0 if residual is None else float(residual[0])
    gw = conv.graph_weight
    w_attn, w_gcn = ((1.0 - gw), gw) if (conv.use_graph and gw > 0) else (1.0, 1.0)
    addends = []
    if conv.use_graph and shard is None:
        # single GPU: the value branch (mean_h V from the weights alone -> SpMM) runs on a side stream beside the Gram pass and the
        # operand algebra; the SpMM (no shared memory) shares the SMs with both.  Joined before pass 2.
        main, side = torch.cuda.current_stream(x.device), projected.side_stream(x.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            if conv.use_weight:
                vbar_part, one = projected.value_operands(conv, x.device)
                vbar = projected.head_mean_values(x, vbar_part, one, 0)
            else:
                vbar = x
            gmean = ops.spmm(ops.graph_csr(edge_index, edge_weight, N), vbar.view(N, 1, projected.HID)).view(N, projected.HID)
        vpart, nvec, _ = projected.projected_operands(projected.gram(x, None), float(n_total), conv, with_values=False)
        main.wait_stream(side)       # fork / join per layer: every buffer of one stream that the other touches is ordered by these two waits
        addends.append((gmean, alpha * w_gcn))
    else:
        vpart, nvec, vbar_part = projected.projected_operands(projected.gram(x, shard), float(n_total), conv, with_values=conv.use_graph and conv.use_weight)
    if conv.use_graph and shard is not None:
        # row-sharded: one stream (the in-kernel NVLink exchange of the Gram pass and the NCCL all-gather must not wait on each other)
        vbar = projected.head_mean_values(x, vbar_part, nvec, H) if conv.use_weight else x     # mean_h V [N, 64] (commutes with the SpMM)
        from .sharded import gather_rows
        if N != shard.end - shard.begin:
            raise ValueError(f"row shard expects {shard.end - shard.begin} local rows, got {N}")
        v_all = gather_rows(vbar, shard.pg, n_total)
        gmean = ops.spmm(ops.graph_csr(edge_index, edge_weight, n_total), v_all.view(n_total, 1, projected.HID),
                         rows=(shard.begin, shard.end)).view(N, projected.HID)
        addends.append((gmean, alpha * w_gcn))
    if getattr(conv, "use_source", False):
        addends.append((ops._f32c(x_0), alpha))
    if residual is not None:
        addends.append((ops._f32c(residual[1]), 1.0 - alpha))
    ln = None
    if (layer_norm is not None and residual is not None and layer_norm.elementwise_affine and layer_norm.bias is not None
            and tuple(layer_norm.normalized_shape) == (projected.HID,)):
        ln = (layer_norm.weight.detach().float().contiguous(), layer_norm.bias.detach().float().contiguous(), layer_norm.eps)
    ep = ops.make_epilogue(alpha * w_attn / H, addends, layer_norm=ln)
    out = projected.apply(x, vpart, nvec, H, ep, keep=(addends, ln))
    return out, None, (2 if ln is not None else 1) if residual is not None else 0


def _conv_forward(conv, query_input, source_input, edge_index, edge_weight, x_0, output_attn,
                  residual=None, n_nodes=None, layer_norm=None):
    """Shared body of DIFFormerConv.forward / TransConv.forward.

    residual = (alpha, prev) folds `alpha*x + (1-alpha)*prev` (difformer.py:200-201) into the fused
    epilogue; the return flag says whether it was applied.  layer_norm = the nn.LayerNorm that follows the layer
    (difformer.py:202-203): folded into the same epilogue when the tcgen05 kernel runs it (flag value 2)."""
    H, C = conv.num_heads, conv.out_channels
    segmented_ = n_nodes is not None
    shard_ = getattr(conv, "_row_shard", None)
    if shard_ is not None and shard_.world < 2:
        shard_ = None
    if (ops._PROJECTION_FOLDING and not output_attn and not segmented_ and conv.kernel == "simple"
            and _fusable("simple", query_input, source_input, x_0, None if residual is None else residual[1],
                         *[p_ for p_ in conv.parameters()])
            and projected.supported(conv, query_input, source_input)):
        return _conv_forward_projected(conv, query_input, edge_index, edge_weight, x_0, residual, layer_norm, shard_)
    query = conv.Wq(query_input).reshape(-1, H, C)
    key = conv.Wk(source_input).reshape(-1, H, C)
    if conv.use_weight:
        value = conv.Wv(source_input).reshape(-1, H, C)
    else:
        value = source_input.reshape(-1, 1, C)                      # difformer.py:120
    segmented = n_nodes is not None
    shard = getattr(conv, "_row_shard", None)             # sharded.shard_model: these rows are a shard of a larger graph
    if shard is not None and shard.world < 2:
        shard = None
    if shard is not None and (segmented or output_attn):
        raise NotImplementedError("row-sharded propagation covers full_attention_conv and gcn_conv; batched graphs shard by graph "
                                  "(ops.segmented_full_attention(..., group=)), attention maps need every row")
    use_source = getattr(conv, "use_source", False)
    gw = conv.graph_weight
    w_attn, w_gcn = ((1.0 - gw), gw) if (conv.use_graph and gw > 0) else (1.0, 1.0)

    if (not output_attn) and (not segmented) and shard is None and C <= ops._MAX_NATIVE_WIDTH and _fusable(conv.kernel, query, key, value, x_0,
                                                          None if residual is None else residual[1]):
        # ---- fused inference path: everything after the Linears is two kernels (+ one SpMM)
        q, k, v = ops._f32c(query), ops._f32c(key), ops._f32c(value)
        ops._need_cuda(q, k, v)
        N = q.shape[0]
        alpha = 1.0 if residual is None else float(residual[0])
        # pass 1 also emits mean_h(V) when a gcn term follows (V is streaming through the SM anyway)
        vbar = None
        if conv.use_graph and v.shape[1] > 1:
            vbar = torch.empty((N, v.shape[2]), dtype=torch.float32, device=v.device)
        partials, prepared = ops.simple_partials(q, k, v, with_prepared=True, vbar=vbar)
        addends, gcn = [], None
        if conv.use_graph:
            csr = ops.graph_csr(edge_index, edge_weight, N)
            # the head mean commutes with the SpMM: gather 256 B rows of mean_h(V) (L2-resident, T/H bytes)
            # instead of H x 256 B rows of V
            src = vbar if vbar is not None else v
            if (csr.max_degree is not None and csr.max_degree <= ops.GCN_EPILOGUE_MAX_DEGREE and v.shape[2] == 64
                    and ops.layer_tail_fusable(H, v.shape[1], C, v.shape[2])):
                gcn = (csr, src.reshape(N, v.shape[2]), alpha * w_gcn)      # gathered inside the pass-2 epilogue: never written to HBM
            else:
                gmean = ops.spmm(csr, src.view(N, 1, v.shape[2])).view(N, v.shape[2])
                addends.append((gmean, alpha * w_gcn))
        if use_source:
            addends.append((ops._f32c(x_0), alpha))
        if residual is not None:
            addends.append((ops._f32c(residual[1]), 1.0 - alpha))
        ln = None
        if (layer_norm is not None and residual is not None and layer_norm.elementwise_affine and layer_norm.bias is not None
                and tuple(layer_norm.normalized_shape) == (v.shape[2],) and ops.layer_tail_fusable(H, v.shape[1], C, v.shape[2])):
            ln = (layer_norm.weight.detach().float().contiguous(), layer_norm.bias.detach().float().contiguous(), layer_norm.eps)
        ep = ops.make_epilogue(alpha * w_attn / H, addends, layer_norm=ln, gcn=gcn)
        out = ops.simple_apply(q, partials, float(N), v.shape[1], v.shape[2], ep, keep=(addends, ln, gcn), prepared=prepared)
        return out, None, (2 if ln is not None else 1) if residual is not None else 0

    # ---- unfused path (training, sigmoid, batched graphs, attention visualisation)
    attn = None
    if segmented:
        attention_output = ops.segmented_full_attention(query, key, value, conv.kernel, n_nodes)
    elif output_attn:
        attention_output, attn = ops.full_attention_conv(query, key, value, conv.kernel, True)
    elif shard is not None and conv.kernel == "sigmoid":
        # every query row needs every key / value row: K and V are all-gathered over the process group (autograd: reduce-scatter of
        # dK, dV), the O(N L) work itself is sharded by query rows
        from .sharded import gather_rows
        attention_output = ops.full_attention_conv(query, gather_rows(key, shard.pg, shard.n_total),
                                                   gather_rows(value, shard.pg, shard.n_total), "sigmoid")
    elif shard is not None:
        attention_output = ops.full_attention_conv(query, key, value, conv.kernel, group=shard.attn_group, n_total=shard.n_total)
    else:
        attention_output = ops.full_attention_conv(query, key, value, conv.kernel)
    if conv.use_graph:
        g = ops.gcn_conv(value, edge_index, edge_weight, shard=shard)
        final_output = w_attn * attention_output + w_gcn * g if gw > 0 else attention_output + g
    else:
        final_output = attention_output
    final_output = final_output.mean(dim=1)
    if use_source:
        final_output = final_output + x_0
    return final_output, attn, False


class DIFFormerConv(nn.Module):
    """one DIFFormer layer (difformer.py:81-145)"""

    def __init__(self, in_channels, out_channels, num_heads, kernel='simple', use_graph=True, use_weight=True,
                 graph_weight=-1, use_source=False):
        super(DIFFormerConv, self).__init__()
        self.Wk = nn.Linear(in_channels, out_channels * num_heads)
        self.Wq = nn.Linear(in_channels, out_channels * num_heads)
        if use_weight:
            self.Wv = nn.Linear(in_channels, out_channels * num_heads)
        self.out_channels = out_channels
        self.num_heads = num_heads
        self.kernel = kernel
        self.use_graph = use_graph
        self.use_weight = use_weight
        self.graph_weight = graph_weight
        self.use_source = use_source

    def reset_parameters(self):
        self.Wk.reset_parameters()
        self.Wq.reset_parameters()
        if self.use_weight:
            self.Wv.reset_parameters()

    def forward(self, query_input, source_input, edge_index=None, edge_weight=None, x_0=None, output_attn=False,
                _residual=None, _layer_norm=None):
        out, attn, fused_res = _conv_forward(self, query_input, source_input, edge_index, edge_weight, x_0,
                                             output_attn, residual=_residual, layer_norm=_layer_norm)
        if _residual is not None:
            return out, fused_res
        return (out, attn) if output_attn else out


class DIFFormer(nn.Module):
    """DIFFormer model class (difformer.py:147-226)
    x: input node features [N, D]; edge_index: [2, E]; returns logits [N, C]"""

    def __init__(self, in_channels, hidden_channels, out_channels, num_layers=2, num_heads=1, kernel='simple',
                 alpha=0.5, dropout=0.5, use_bn=True, use_residual=True, use_weight=True, use_graph=True,
                 graph_weight=-1, use_source=False):
        super(DIFFormer, self).__init__()
        self.convs = nn.ModuleList()
        self.fcs = nn.ModuleList()
        self.fcs.append(nn.Linear(in_channels, hidden_channels))
        self.bns = nn.ModuleList()
        self.bns.append(nn.LayerNorm(hidden_channels))
        for i in range(num_layers):
            self.convs.append(DIFFormerConv(hidden_channels, hidden_channels, num_heads=num_heads, kernel=kernel,
                                            use_graph=use_graph, use_weight=use_weight, graph_weight=graph_weight,
                                            use_source=use_source))
            self.bns.append(nn.LayerNorm(hidden_channels))
        self.fcs.append(nn.Linear(hidden_channels, out_channels))
        self.dropout = dropout
        self.activation = F.relu
        self.use_bn = use_bn
        self.residual = use_residual
        self.alpha = alpha

    def reset_parameters(self):
        for conv in self.convs:
            conv.reset_parameters()
        for bn in self.bns:
            bn.reset_parameters()
        for fc in self.fcs:
            fc.reset_parameters()

    def forward(self, x, edge_index, edge_weight=None):
        layer_ = []
        x = self.fcs[0](x)
        if self.use_bn:
            x = self.bns[0](x)
        x = self.activation(x)
        x = F.dropout(x, p=self.dropout, training=self.training)
        layer_.append(x)
        for i, conv in enumerate(self.convs):
            res = (self.alpha, layer_[i]) if self.residual else (1.0, None)
            fused = 0
            if self.residual:
                # no-grad path: the residual blend AND the LayerNorm that follows go into the kernel's epilogue
                x, fused = conv(x, x, edge_index, edge_weight, layer_[0], _residual=res,
                                _layer_norm=self.bns[i + 1] if self.use_bn else None)
                if not fused:
                    x = self.alpha * x + (1 - self.alpha) * layer_[i]
            else:
                x = conv(x, x, edge_index, edge_weight, layer_[0])
            if self.use_bn and fused != 2:
                x = self.bns[i + 1](x)
            x = F.dropout(x, p=self.dropout, training=self.training)
            layer_.append(x)
        return self.fcs[-1](x)

    def get_attentions(self, x):
        layer_, attentions = [], []
        x = self.fcs[0](x)
        if self.use_bn:
            x = self.bns[0](x)
        x = self.activation(x)
        layer_.append(x)
        for i, conv in enumerate(self.convs):
            x, attn = conv(x, x, output_attn=True)
            attentions.append(attn)
            if self.residual:
                x = self.alpha * x + (1 - self.alpha) * layer_[i]
            if self.use_bn:
                x = self.bns[i + 1](x)
            layer_.append(x)
        return torch.stack(attentions, dim=0)


class TransConv(nn.Module):
    """batched-graph layer (difformer-v2.py:48-163); in_channels must equal out_channels when use_weight is False"""

    def __init__(self, in_channels, out_channels, num_heads=1, kernel='simple', use_graph=True, use_weight=True, graph_weight=-1):
        super().__init__()
        self.Wk = nn.Linear(in_channels, out_channels * num_heads)
        self.Wq = nn.Linear(in_channels, out_channels * num_heads)
        if use_weight:
            self.Wv = nn.Linear(in_channels, out_channels * num_heads)
        self.out_channels = out_channels
        self.num_heads = num_heads
        self.kernel = kernel
        self.use_graph = use_graph
        self.use_weight = use_weight
        self.graph_weight = graph_weight

    def reset_parameters(self):
        self.Wk.reset_parameters()
        self.Wq.reset_parameters()
        if self.use_weight:
            self.Wv.reset_parameters()

    def full_attention(self, qs, ks, vs, kernel, n_nodes):
        return ops.segmented_full_attention(qs, ks, vs, kernel, n_nodes)

    def forward(self, query_input, source_input, n_nodes, edge_index=None, edge_weight=None):
        # the reference only binds `value` under use_weight (difformer-v2.py:149-150) and fails with
        # UnboundLocalError otherwise; here use_weight=False uses the input like DIFFormerConv does
        out, _, _ = _conv_forward(self, query_input, source_input, edge_index, edge_weight, None, False, n_nodes=n_nodes)
        return out


class DIFFormer_v2(nn.Module):
    """difformer-v2.py:165-223; forward(x, edge_index, n_nodes)"""

    def __init__(self, in_channels, hidden_channels, out_channels, num_layers=2, kernel='simple', alpha=0.5, dropout=0.5,
                 use_bn=True, use_residual=True, use_weight=True, use_graph=True, graph_weight=-1):
        super().__init__()
        self.convs = nn.ModuleList()
        self.fcs = nn.ModuleList()
        self.fcs.append(nn.Linear(in_channels, hidden_channels))
        self.bns = nn.ModuleList()
        self.bns.append(nn.LayerNorm(hidden_channels))
        for i in range(num_layers):
            self.convs.append(TransConv(hidden_channels, hidden_channels, kernel=kernel, use_graph=use_graph,
                                        use_weight=use_weight, graph_weight=graph_weight))
            self.bns.append(nn.LayerNorm(hidden_channels))
        self.fcs.append(nn.Linear(hidden_channels, out_channels))
        self.dropout = dropout
        self.activation = F.relu
        self.use_bn = use_bn
        self.residual = use_residual
        self.alpha = alpha

    def reset_parameters(self):
        for conv in self.convs:
            conv.reset_parameters()
        for bn in self.bns:
            bn.reset_parameters()
        for fc in self.fcs:
            fc.reset_parameters()

    def forward(self, x, edge_index, n_nodes):
        layer_ = []
        x = self.fcs[0](x)
        if self.use_bn:
            x = self.bns[0](x)
        x = self.activation(x)
        x = F.dropout(x, p=self.dropout, training=self.training)
        layer_.append(x)
        for i, conv in enumerate(self.convs):
            x = conv(x, x, n_nodes, edge_index)
            if self.residual:
                x = self.alpha * x + (1 - self.alpha) * layer_[i]
            if self.use_bn:
                x = self.bns[i + 1](x)
            x = F.dropout(x, p=self.dropout, training=self.training)
            x = self.activation(x)
            layer_.append(x)
        x_out = self.fcs[-1](x)
        return F.dropout(x_out, p=self.dropout, training=self.training)


class GraphedForward:
    """CUDA-graph replay of `model(x, edge_index, ...)` for inference on a fixed graph (the small-N
    regimes -- Cora, the spatial-temporal snapshots with n <= 1068 -- are bound by kernel launches and Python, not by
    the GPU).  The model is put in eval mode, run a few times eagerly (builds and caches the CSR, sets kernel
    attributes), then captured once with `torch.no_grad()`; every call copies the new floating-point inputs into the
    captured buffers and replays the graph.

    Only floating-point tensors (node features, edge weights) may change between calls; integer tensors
    (`edge_index`, `n_nodes`) and all shapes are frozen at capture time -- build a new GraphedForward for a new graph.
    When an `edge_weight` is passed, the build of the normalised CSR values is part of the captured graph
    (`ops.graph_csr` bypasses its cache while capturing), so new edge weights take effect on every replay.
    The returned tensor is the captured output buffer: clone it if it must survive the next call."""

    def __init__(self, model: nn.Module, *example_args, warmup: int = 3):
        if not all(a.is_cuda for a in example_args if torch.is_tensor(a)):
            raise RuntimeError("GraphedForward: example arguments must be CUDA tensors")
        self.model = model.eval()
        self._static = [a.clone() if torch.is_tensor(a) and a.is_floating_point() else a for a in example_args]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):
                self.model(*self._static)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self._out = self.model(*self._static)

    def __call__(self, *args):
        if len(args) != len(self._static):
            raise ValueError(f"GraphedForward: expected {len(self._static)} arguments, got {len(args)}")
        for dst, src in zip(self._static, args):
            if torch.is_tensor(dst) and dst.is_floating_point():
                if src.shape != dst.shape:
                    raise ValueError(f"GraphedForward: shape {tuple(src.shape)} differs from the captured {tuple(dst.shape)}")
                if src.data_ptr() != dst.data_ptr():
                    dst.copy_(src)
        self.graph.replay()
        return self._out
