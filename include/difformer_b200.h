/*
 * difformer_b200 -- C ABI of the B200 (sm_100a) DIFFormer propagation kernels.
 *
 * The reference (qitianwu/DIFFormer) has no FFI: its hot path is ~70 lines of PyTorch in
 * `node classification/difformer.py:10-79` (+ the batched variant in
 * `physical particle/difformer-v2.py:71-111`).  This header is the boundary a maintainer binds
 * instead of those functions (ctypes stub: INTEGRATION.md).  Each entry point cites the
 * reference lines it replaces.
 *
 * Conventions (all entry points):
 *   - extern "C", plain pointers and sizes, no torch types.
 *   - every pointer is a DEVICE pointer owned by the caller (row-major, contiguous, fp32 unless
 *     stated); the library never allocates, never synchronises, never throws.
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream).
 *   - return 0 on success, DIF_EARG (-1) bad argument, DIF_EUNSUPPORTED (-2) shape not supported,
 *     DIF_ECUDA (-3) CUDA error; dif_last_error() returns a thread-local message.
 *   - scratch memory comes from the caller; size it with the matching *_workspace_bytes().
 *   - reentrant per (stream, workspace); results are deterministic (no float atomics anywhere).
 *
 * Tensor names follow the reference: qs[N,H,M], ks[L,H,M], vs[L,Hv,D] with Hv == H or Hv == 1
 * (use_weight=False, difformer.py:120), N == L (difformer.py:22,29).
 */
#ifndef DIFFORMER_B200_H
#define DIFFORMER_B200_H

#include <stdint.h>

#if defined(DIF_BUILD) && defined(__GNUC__)
#define DIF_API __attribute__((visibility("default")))
#else
#define DIF_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define DIF_OK 0
#define DIF_EARG (-1)
#define DIF_EUNSUPPORTED (-2)
#define DIF_ECUDA (-3)

/* kernel implementation selector for the 'simple' path */
/* element type of the node tensors (partials are always fp32) */
#define DIF_DTYPE_F32 0
#define DIF_DTYPE_BF16 1
#define DIF_DTYPE_F16 2

#define DIF_IMPL_AUTO 0     /* tcgen05 path when the shape qualifies, else generic */
#define DIF_IMPL_GENERIC 1  /* FFMA kernels, any H, M%4==0, D%4==0, M,D <= 128 */
#define DIF_IMPL_TCGEN05 2  /* tcgen05/TMEM kernels: M == D == 64, Hv == H, H in {1, 2, 4} */

DIF_API int dif_version(void);
DIF_API const char* dif_last_error(void);
/* 1 when the current device is compute capability 10.x (the only one this library targets) */
DIF_API int dif_device_supported(void);

/* ------------------------------------------------------------------------------------------
 * kernel='simple'  (full_attention_conv, difformer.py:18-39)
 *
 * Pass 1  dif_simple_reduce : row reductions of this rank's rows -> `partials`
 *           partials = [ S : H*M*D | z : H*M | u : Hv*D | sum q^2 | sum k^2 ]   (fp32, un-normalised)
 *           S[h,m,d] = sum_l k[l,h,m] v[l,hv,d]   (difformer.py:25)
 *           z[h,m]   = sum_l k[l,h,m]             (difformer.py:32-33)
 *           u[hv,d]  = sum_l v[l,hv,d]            (difformer.py:27-28)
 *           sums of squares replace torch.norm    (difformer.py:20-21)
 *         The partials are additive over row shards: multi-GPU = one all-reduce(sum) of this
 *         buffer between pass 1 and pass 2 (dif_simple_partials_len() floats, 67.6 KB at H=4,D=64).
 * Pass 2  dif_simple_apply  : out = (q^ S^ + u) / (q^ z^ + n_total)   (difformer.py:26,29,34-39)
 * ------------------------------------------------------------------------------------------ */
DIF_API int64_t dif_simple_partials_len(int H, int Hv, int M, int D);
DIF_API int64_t dif_simple_workspace_bytes(int64_t N, int H, int Hv, int M, int D);
/* `vbar` (optional, may be NULL): pass 1 also writes mean_h(V) as [N,D] (V is streaming through the SM anyway);
 * it feeds the gcn SpMM of the fused layer (the head mean commutes with the SpMM).  For Hv == 1 it is a copy-free
 * no-op on the tcgen05 path (V itself is mean_h(V)): callers should pass NULL then.
 * `prepared` (optional, may be NULL): when dif_simple_prepared_bytes() > 0 (tcgen05 shapes) pass 1 also writes the
 * pass-2 tensor-core operands derived from `partials` (bf16 hi/lo split, 128B-swizzled) into this caller-owned
 * buffer (16-byte aligned), which saves pass 2 its transposing prologue.  It is only valid for the exact
 * `partials` pass 1 produced: after an all-reduce (or any edit) of `partials` pass NULL to pass 2. */
DIF_API int64_t dif_simple_prepared_bytes(int H, int Hv, int M, int D);

DIF_API int dif_simple_reduce(const float* q, const float* k, const float* v,
                      int64_t N, int H, int Hv, int M, int D,
                      float* partials, void* prepared, float* vbar, void* workspace, int64_t workspace_bytes,
                      int impl, void* stream);

/* Epilogue of pass 2.
 *   mode 0: out[N,H,D] = attention output (what full_attention_conv returns).
 *   mode 1: layer epilogue fused (DIFFormerConv.forward difformer.py:129-140 + residual :200-201):
 *           out[N,D] = attn_scale * sum_h attn[n,h,:] + sum_j add_scale[j] * add[j][n,:]
 *           (attn_scale = alpha*w_attn/H; addends = head-meaned gcn term, x_0, previous layer) */
typedef struct {
    int mode;
    float attn_scale;
    int n_add;              /* 0..3 */
    const float* add[3];    /* each [N,D] or NULL */
    float add_scale[3];
    /* optional tail of the layer, applied to the finished [N,D] row (tcgen05 kernels only: DIF_EUNSUPPORTED on the FFMA path):
     * LayerNorm over D (difformer.py:202-203: weight / bias [D], eps) when ln_weight != NULL, then ReLU when relu != 0 */
    const float* ln_weight;
    const float* ln_bias;
    float ln_eps;
    int relu;
    /* optional gcn_conv term gathered INSIDE the epilogue (tcgen05 kernels only), so that it is never written to HBM:
     * out[n,:] += gcn_scale * sum over the CSR slots s of row n of gcn_val[s] * gcn_x[gcn_idx[s], :]
     * with gcn_x = mean_h(V) [N,D] (the `vbar` output of dif_simple_reduce; the head mean commutes with the SpMM) and
     * (gcn_rowptr, gcn_idx, gcn_val) the target-sorted CSR of dif_csr_build.  Each epilogue thread owns one output row and walks
     * its slots (L2-resident 256-byte rows): meant for graphs without very high-degree rows -- otherwise run dif_gcn_spmm and pass
     * its result as an addend. */
    const int32_t* gcn_rowptr;
    const int32_t* gcn_idx;
    const float* gcn_val;
    const float* gcn_x;
    float gcn_scale;
} dif_epilogue_t;

DIF_API int dif_simple_apply(const float* q, const float* partials, const void* prepared, double n_total,
                     int64_t N, int H, int Hv, int M, int D,
                     float* out, const dif_epilogue_t* epilogue,
                     int impl, void* stream);

/* The forward in ONE kernel (dif_simple_forward_workspace_bytes() > 0: M == D == 64 with Hv == H in {1, 2, 4}, any dtype; or one
 * head of M == D == 128, fp32 -- hidden_channels 128 of run.sh:43,70,75): a cooperative persistent launch runs
 * pass 1, the grid-wide deterministic sum of the partials (with `peer_bufs` != NULL and world > 1 also the cross-GPU
 * LL-push all-reduce, see dif_comm_* below) and pass 2 on the rows each CTA just streamed -- no second launch, the Q rows
 * of pass 2 are prefetched while the sum is in flight.  out[N,H,D] = full_attention_conv(q,k,v,'simple'); `partials`
 * receives the (all-reduced) pass-1 partials, which the backward needs.  n_total = global row count (= N unsharded).
 * `workspace` must be 128-byte aligned.  peer_bufs / rank / world / seq as in dif_simple_reduce_allreduce (NULL, 0, 1, 0
 * for a single GPU).  Other shapes return DIF_EUNSUPPORTED: call dif_simple_reduce + dif_simple_apply.
 * `dtype` = element type of q, k, v AND out: DIF_DTYPE_F32 or DIF_DTYPE_BF16 (the Linear outputs under bf16 autocast): the
 * bf16 kernel feeds the TMA-landed tiles straight to the tensor cores (no conversion pass, exact products, fp32 accumulation
 * and fp32 partials) and moves 2048 instead of 4096 algorithmic bytes per node at H = 4, D = 64.  DIF_DTYPE_F16 returns
 * DIF_EUNSUPPORTED: up-cast fp16 tensors and use the fp32 kernel (what difformer_b200.ops does). */
DIF_API int64_t dif_simple_forward_workspace_bytes(int64_t N, int H, int Hv, int M, int D);
DIF_API int dif_simple_forward(const void* q, const void* k, const void* v, int dtype,
                       int64_t N, int H, int Hv, int M, int D, double n_total,
                       float* partials, void* out, void* workspace, int64_t workspace_bytes,
                       void* const* peer_bufs, int rank, int world, unsigned long long seq, void* stream);

/* Pass 2 with the Wq projection folded into its operands (difformer.py:115-118).  The A operand is the layer input
 * x [N,64] itself (row stride ldx floats, shared by the H heads); `vpartials` holds, in the partials layout of (H, Hv = H, M = 64, D = 64),
 * S'_h = Wq_h^T S_h, z'_h = Wq_h^T z_h, u'_h = u_h + c bq_h^T S_h and (sum q^2, sum k^2); n_total_vec[h] = N + c bq_h . z_h (device).
 * out[n,h,:] = (c x_n S'_h + u'_h) / (c x_n . z'_h + n_total_vec[h]) = full_attention_conv(x Wq^T + bq, K, V, 'simple')[n,h,:]; with the
 * reductions of pass 1 taken from the Gram matrix X^T X (dif_simple_reduce(x, x, x, H = 1)) neither Q nor K nor V is ever written.
 * hidden = 64, H in {1, 2, 4}; the mode-1 epilogue (head mean, addends, LayerNorm) applies as in dif_simple_apply. */
DIF_API int dif_simple_apply_projected(const float* x, int64_t ldx, const float* vpartials, const float* n_total_vec,
                               int64_t N, int H, float* out, const dif_epilogue_t* epilogue, void* stream);

/* The operands of dif_simple_apply_projected from the Gram matrix of the layer input (replaces the three Linears of
 * difformer.py:115-120 together with pass 1 of full_attention_conv :18-39).  `gram_partials` = the output of dif_simple_reduce(x, x, x,
 * H = Hv = 1, M = D = 64): [X^T X | X^T 1 | X^T 1 | sum x^2 | sum x^2] (all-reduced over row shards when x is sharded; n_total = global
 * row count).  Wq/Wk/Wv: nn.Linear weights [H*64, 64] row-major with biases [H*64] (device, fp32); Wv = bv = NULL means V = x
 * (use_weight=False, difformer.py:120).  Writes vpartials [H*4096 + 2*H*64 + 2] and n_total_vec [H + 1] (see
 * dif_simple_apply_projected), and vbar_partials [4096 + 2*64 + 2]: the head mean of the value projection posed as a one-head pass-2
 * problem -- dif_simple_apply_projected(x, ldx, vbar_partials, n_total_vec + H, N, 1, vbar, NULL) writes mean_h V = x wbar^T + bbar,
 * the input of the gcn term (difformer.py:139; the head mean commutes with the SpMM).
 * vbar_partials may be NULL; dif_simple_project_values writes the same vbar_partials (and the denominator constant `one[0]` = 1) from
 * the weights alone, so the value branch of a layer (mean_h V -> SpMM) need not wait for the Gram matrix.
 * fp64 arithmetic (the 64 x 64 products on the FP64 tensor cores), deterministic, two small launches.
 * workspace: dif_simple_project_workspace_bytes(H), 8-byte aligned. */
DIF_API int64_t dif_simple_project_workspace_bytes(int H);
DIF_API int dif_simple_project(const float* gram_partials, const float* Wq, const float* bq, const float* Wk, const float* bk,
                       const float* Wv, const float* bv, double n_total, int H, float* vpartials, float* n_total_vec,
                       float* vbar_partials, void* workspace, int64_t workspace_bytes, void* stream);
DIF_API int dif_simple_project_values(const float* Wv, const float* bv, int H, float* vbar_partials, float* one, void* stream);

/* Backward of the 'simple' path (derived analytically; the reference uses autograd).
 *   bwd_partials = [ dS : H*M*D | dz : H*M | du : H*D | t_q | t_k ]  (raw, additive over shards;
 *   t_k is filled by dif_simple_bwd_apply after any all-reduce). `out` is the saved forward
 *   output [N,H,D], `g` = dL/dout. */
DIF_API int64_t dif_simple_bwd_partials_len(int H, int M, int D);
/* `rowscal`: dif_simple_bwd_rowscal_len() floats of caller-owned scratch ((1/den, dden) per (node, head)) that the
 * tcgen05 backward hands from its pass 1 to its pass 2; NULL (or len 0) selects the FFMA kernels. */
DIF_API int64_t dif_simple_bwd_rowscal_len(int64_t N, int H, int Hv, int M, int D);
DIF_API int dif_simple_bwd_reduce(const float* q, const float* g, const float* out, const float* partials,
                          double n_total, int64_t N, int H, int Hv, int M, int D,
                          float* bwd_partials, float* rowscal, void* workspace, int64_t workspace_bytes,
                          int impl, void* stream);
DIF_API int dif_simple_bwd_apply(const float* q, const float* k, const float* v, const float* g, const float* out,
                         const float* partials, float* bwd_partials, const float* rowscal, double n_total,
                         int64_t N, int H, int Hv, int M, int D,
                         float* dq, float* dk, float* dv, int impl, void* stream);

/* Batched-graph 'simple' (TransConv.full_attention, difformer-v2.py:80-111): rows are grouped in
 * B contiguous segments seg_ptr[0..B] (int32, device; seg_ptr[B] == N).  Normalisation by n_g per
 * graph (:107-109), Frobenius norms over the whole batch (:82-83).
 * norms = [sum q^2, sum k^2] (device, 2 floats) is produced by dif_sumsq2 (additive over shards). */
DIF_API int dif_sumsq2(const float* q, const float* k, int64_t count, float* norms,
               void* workspace, int64_t workspace_bytes, void* stream);
DIF_API int dif_segmented_simple_fwd(const float* q, const float* k, const float* v, const int32_t* seg_ptr,
                             int32_t B, const float* norms, int64_t N, int H, int Hv, int M, int D,
                             float* out, void* stream);
/* The same forward on the tensor cores (tcgen05): H = Hv = 1, M = D = 64, every graph <= max_nodes <= 128 rows.  Whole graphs are
 * packed into 128-row tiles that run as block-diagonal dense attention (scores Q K^T, weights mask o (1 + c s), W V), see
 * csrc/segmented_sm100.cu.  The packing is a device-side plan built once per batch layout from seg_ptr (no host loop, no data
 * movement): dif_segmented_plan_bytes(N, max_nodes) bytes (0: unsupported), 16-byte aligned, valid for any q / k / v of that layout.
 * q, k, v 32-byte aligned.  Results equal dif_segmented_simple_fwd to fp32 rounding. */
DIF_API int64_t dif_segmented_plan_bytes(int64_t N, int max_nodes);
DIF_API int dif_segmented_plan_build(const int32_t* seg_ptr, int32_t B, int64_t N, int max_nodes, void* plan, int64_t plan_bytes, void* stream);
DIF_API int dif_segmented_simple_fwd_tc(const float* q, const float* k, const float* v, const void* plan, int64_t plan_bytes,
                                const float* norms, int64_t N, int max_nodes, float* out, void* stream);
/* Backward on the same tiles (scores and dO V^T, weights, dV = W'^T dO, dS, dQ = dS K, dK = dS^T Q: five tensor-core products per
 * tile).  phase / workspace as dif_segmented_simple_bwd_phase: the batch-wide scalars end up at workspace float offset 2 B after
 * phase 1 (all-reduce them across ranks when the graphs are sharded), phase 2 applies them; phase 0 = both. */
DIF_API int dif_segmented_simple_bwd_tc(const float* q, const float* k, const float* v, const float* g, const float* out,
                                const void* plan, int64_t plan_bytes, const float* norms, int64_t N, int max_nodes, int32_t B,
                                float* dq, float* dk, float* dv, void* workspace, int64_t workspace_bytes, int phase, void* stream);
/* backward: `out` is the saved forward output, g = dL/dout; M in {16,32,64}, D <= 64.  M == D == 64: graphs of up to
 * 64 rows run one warp per graph (direct O(n^2) form), larger ones one CTA per graph; no atomics, deterministic. */
DIF_API int dif_segmented_simple_bwd(const float* q, const float* k, const float* v, const float* g, const float* out,
                             const int32_t* seg_ptr, int32_t B, const float* norms,
                             int64_t N, int H, int Hv, int M, int D,
                             float* dq, float* dk, float* dv,
                             void* workspace, int64_t workspace_bytes, void* stream);
/* Graphs sharded over ranks (whole graphs per rank): the backward needs the batch-wide scalars (t_q, t_k) summed over ALL ranks.
 * phase 1 leaves this rank's two sums at workspace float offset 2*B and returns; all-reduce those two floats; phase 2 (same
 * arguments, same workspace) finishes dq, dk, dv.  phase 0 = dif_segmented_simple_bwd. */
DIF_API int dif_segmented_simple_bwd_phase(const float* q, const float* k, const float* v, const float* g, const float* out,
                             const int32_t* seg_ptr, int32_t B, const float* norms,
                             int64_t N, int H, int Hv, int M, int D,
                             float* dq, float* dk, float* dv,
                             void* workspace, int64_t workspace_bytes, int phase, void* stream);
DIF_API int64_t dif_segmented_workspace_bytes(int32_t B);

/* ------------------------------------------------------------------------------------------
 * kernel='sigmoid'  (full_attention_conv, difformer.py:45-56): tiled, never materialises [N,L,H].
 *   out = (sigmoid(QK^T) / rowsum) V ; rowsum[N,H] is saved for the backward.
 *   Forward with M == D == 64 runs on tcgen05 (flash-style; Q, K, V and P = sigmoid(S) are all split into bf16
 *   hi + lo operands, fp32 accumulation, P and Q read by the MMA from tensor memory: ~5e-6 of the fp64 result; scores
 *   below -43 are clamped to -43, sigmoid = 2e-19); other shapes and the backward run the fp32 FFMA kernels.  The
 *   workspace holds the key-split partials and the bf16 operand images of K and V (dif_sigmoid_fwd_workspace_bytes).
 *   dif_sigmoid_set_impl(DIF_IMPL_GENERIC / _TCGEN05 / _AUTO) pins the forward path (process-wide); pinned to
 *   _TCGEN05, an unsupported shape returns DIF_EUNSUPPORTED instead of falling back.
 * ------------------------------------------------------------------------------------------ */
DIF_API int dif_sigmoid_set_impl(int impl);
DIF_API int64_t dif_sigmoid_fwd_workspace_bytes(int64_t N, int64_t L, int H, int Hv, int M, int D);
DIF_API int dif_sigmoid_fwd(const float* q, const float* k, const float* v,
                    int64_t N, int64_t L, int H, int Hv, int M, int D,
                    float* out, float* rowsum, void* workspace, int64_t workspace_bytes, void* stream);
DIF_API int dif_sigmoid_bwd(const float* q, const float* k, const float* v, const float* g, const float* out,
                    const float* rowsum, int64_t N, int64_t L, int H, int Hv, int M, int D,
                    float* dq, float* dk, float* dv, void* workspace, int64_t workspace_bytes, void* stream);
DIF_API int64_t dif_sigmoid_bwd_workspace_bytes(int64_t N, int64_t L, int H, int Hv, int M, int D);

/* ------------------------------------------------------------------------------------------
 * gcn_conv  (difformer.py:63-79)
 *   dif_csr_build: edge_index int64 [2,E] (row = source, col = target, difformer.py:65) ->
 *     target-sorted CSR (rowptr[N+1], src[E], val[E]) and its transpose (source-sorted, for the
 *     backward).  val_e = w_e * d[col]^-1/2 * d[row]^-1/2 with d = in-degree of `col` for both
 *     factors (:66-73), non-finite -> 0 (:74).  Stable in edge order => deterministic sums;
 *     duplicates are kept and summed (torch_sparse semantics).  perm[E] = original edge id of each
 *     CSR slot (bit-exact gather-index check).  Indices must be < 2^31.
 *   dif_gcn_spmm: out[c, :] = sum_{slots of c} val * x[src, :]   with F = Hx*D floats per row.
 *     head_mean != 0: out is [N,D] = mean over the Hx heads (used by the fused layer epilogue).
 * ------------------------------------------------------------------------------------------ */
DIF_API int64_t dif_csr_workspace_bytes(int64_t N, int64_t E);
DIF_API int dif_csr_build(const int64_t* edge_index, const float* edge_weight, int64_t N, int64_t E,
                  int32_t* rowptr, int32_t* src, float* val, int32_t* perm,
                  int32_t* rowptr_t, int32_t* dst_t, float* val_t,
                  void* workspace, int64_t workspace_bytes, void* stream);
DIF_API int dif_gcn_spmm(const float* x, const int32_t* rowptr, const int32_t* idx, const float* val,
                 int64_t N, int Hx, int D, int head_mean, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * One-shot NVLink all-reduce of the pass-1 partials (the path's only collective).
 * Every rank owns a peer-mappable buffer of dif_comm_buffer_bytes(len) bytes:
 *   [header: status word | pinned-host-flag pointer][LL region: u64 [2 slots][16 source ranks][len padded to 64]].
 * Protocol (LL push): a rank writes every element of its contribution into each peer's region as one 64-bit word
 * {call number | fp32} and polls the words its peers wrote into its own region -- data and flag travel together, one
 * NVLink traversal, no fences.  Ranks are added in rank order (bit-identical on every rank).
 *   dif_comm_alloc / _free      : the one place the library allocates (cudaMalloc: IPC-exportable, zero-filled; plus a
 *                                 64-byte pinned host flag for the watchdog)
 *   dif_comm_export / _open     : cudaIpc handle (64 bytes) out / peer pointer in (same node, NVLink peers)
 *   dif_comm_allreduce          : call number `seq` (1,2,3,... identical on all ranks): out[i] = sum over ranks of
 *                                 src[i], i < len (src, out: local device memory).  bufs[r] = pointer to rank r's
 *                                 buffer as mapped in this process (HOST array of `world` device pointers, world <= 16).
 *   dif_comm_status             : *timed_out = 1 once a kernel of this rank gave up waiting (30 s) for a peer, or was
 *                                 told by a peer that IT gave up: results since then are meaningless on every rank.
 *                                 Reads a pinned host flag: no device synchronisation (a timeout of a kernel that is
 *                                 still running shows up later).  dif_comm_reset clears it (device sync; all ranks
 *                                 must have quiesced and must reset before the next call).
 * ------------------------------------------------------------------------------------------ */
/* Pass 1 with the all-reduce fused into its tail (one kernel: compute + collective over peer memory): every
 * CTA exchanges "its" column slice of the partials with the peers (LL push, above) right after the local
 * cross-CTA sum.  Result: `partials` (and `prepared`) already hold the sum over all ranks.
 * tcgen05 shapes only (DIF_EUNSUPPORTED otherwise: use dif_simple_reduce + dif_comm_allreduce).  `seq` as in
 * dif_comm_allreduce; the two entry points may share buffers as long as seq keeps increasing. */
DIF_API int dif_simple_reduce_allreduce(const float* q, const float* k, const float* v,
                      int64_t N, int H, int Hv, int M, int D,
                      float* partials, void* prepared, void* workspace, int64_t workspace_bytes,
                      void* const* peer_bufs, int rank, int world, unsigned long long seq, void* stream);
DIF_API int64_t dif_comm_buffer_bytes(int64_t len);
DIF_API int dif_comm_alloc(void** ptr, int64_t bytes);
DIF_API int dif_comm_free(void* ptr);
DIF_API int dif_comm_export(void* ptr, void* handle64);
DIF_API int dif_comm_open(const void* handle64, void** peer_ptr);
DIF_API int dif_comm_close(void* peer_ptr);
DIF_API int dif_comm_status(const void* own_buf, int* timed_out);
DIF_API int dif_comm_reset(void* own_buf);
DIF_API int dif_comm_allreduce(void* const* bufs, int rank, int world, int64_t len, unsigned long long seq,
                               const float* src, float* out, void* stream);

/* mean over heads: x[N,Hx,D] -> out[N,D] */
DIF_API int dif_head_mean(const float* x, int64_t N, int Hx, int D, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIFFORMER_B200_H */
