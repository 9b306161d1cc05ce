/*
 * aero_gnn.h -- C-ABI of libaero_sm100.so: the B200 (sm_100a) implementation of the
 * MeshGraphNets message-passing hot path of cudagu/aero-gnn.
 *
 * The reference has no FFI: its boundary for this path is the Python nn.Module API
 * (SURVEY.md section 8b).  This header is therefore the *new* native boundary the Python
 * modules in aero_gnn_b200/models bind through ctypes; every entry point cites the
 * reference call site (file:line under the reference checkout) whose arithmetic it replaces.
 *
 * Conventions
 *   - every buffer is allocated and owned by the caller (PyTorch); pointers are raw device
 *     pointers; no function allocates, frees, or synchronises the device;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - sizes are int64_t, graph indices handed to kernels are int32_t (E, N < 2^31);
 *   - return value 0 = ok, otherwise an AERO_E* code; aero_last_error() gives the text
 *     (thread-local);
 *   - latent width is fixed at AERO_D = 128 (config.yaml:42 hidden_dim 128) for the fused
 *     block kernels; other widths are rejected with AERO_EUNSUPPORTED (no fallback).
 */
#ifndef AERO_GNN_H_
#define AERO_GNN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AERO_D 128

/* error codes */
#define AERO_OK 0
#define AERO_EINVAL 1        /* bad argument (null pointer, negative size, ...)           */
#define AERO_EUNSUPPORTED 2  /* dtype / width / activation outside the fused path         */
#define AERO_ECUDA 3         /* a CUDA runtime call or kernel launch failed               */
#define AERO_EWORKSPACE 4    /* workspace too small                                       */

/* storage dtype of latent rows */
#define AERO_F32 0
#define AERO_BF16 1

/* activations whose derivative is a function of the activation output
 * (mlp.py:37 getattr(F, activation_fn); mgnLayer.py:81 hard-codes ReLU for EdgeBlockSum) */
#define AERO_ACT_RELU 0
#define AERO_ACT_TANH 1
#define AERO_ACT_SIGMOID 2
#define AERO_ACT_ELU 3
#define AERO_ACT_LEAKY_RELU 4

/* which kernel family executes a block: SIMT = fp32 CUDA-core math (exact fp32 parity path),
 * UMMA = bf16 tcgen05 tensor-core tiles with fp32 accumulation in TMEM. */
#define AERO_PATH_SIMT 0
#define AERO_PATH_UMMA 1

const char* aero_last_error(void);
int aero_version(void);
/* 1 when the library was compiled with the tcgen05 (UMMA) forward kernels / backward kernels */
int aero_has_umma(void);
int aero_has_umma_bwd(void);
/* tcgen05 primitive self-test: one 128x128x128 bf16 GEMM, fp32 result c[128][128] (row-major inputs).
 *   mode 0: c = a  * b^T   (both operands K-major: forward Linear)
 *   mode 1: c = a  * b     (b MN-major: data gradient)
 *   mode 2: c = a^T * b    (both MN-major: weight gradient) */
int aero_umma_selftest(const void* a_bf16, const void* b_bf16, float* c, int mode, void* stream);
/* ------------------------------------------------------------------------------------------
 * Graph plan: receiver-CSR + sender-CSR, built once per mesh.
 * Replaces the per-step index work of mgnLayer.py:39-41 (row/col gathers), :101 and the
 * atomics of torch_scatter.scatter_add at mgnLayer.py:146.
 *
 *   edge_index  [2,E] int64 (row 0 = sender, row 1 = receiver), any order, duplicates and
 *               self-loops allowed (bsms_mgn.py:280-288 produces both).
 *   rowptr      [N+1]  receiver-CSR offsets
 *   perm        [E]    perm[k] = caller edge id stored at CSR slot k (stable: ascending id
 *                      inside one receiver, i.e. the order CPU scatter_add_ accumulates in)
 *   src, dst    [E]    sender / receiver of CSR slot k
 *   sptr        [N+1]  sender-CSR offsets
 *   sperm       [E]    CSR slots whose sender is n, ascending, for n = 0..N-1
 *   status      [4]    device int32: [0] = number of out-of-range indices found
 * ------------------------------------------------------------------------------------------ */
size_t aero_graph_plan_workspace_bytes(int64_t E, int64_t N);
int aero_graph_plan_build(const int64_t* edge_index, int64_t E, int64_t N,
                          int32_t* rowptr, int32_t* perm, int32_t* src, int32_t* dst,
                          int32_t* sptr, int32_t* sperm, int32_t* status,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Stable LSD radix sort of (key,value) pairs, `key_bits` low bits significant.
 * keys_out/vals_out receive the result; keys_in/vals_in are clobbered.  Building block of the
 * plan and of the bistride index kernels (replaces torch.argsort bsms_mgn.py:242 and the
 * sort inside torch.unique bsms_mgn.py:280). */
size_t aero_sort_pairs_workspace_bytes(int64_t n);
int aero_sort_pairs_u64(uint64_t* keys_in, int32_t* vals_in, uint64_t* keys_out, int32_t* vals_out,
                        int64_t n, int key_bits, void* workspace, size_t workspace_bytes, void* stream);

/* 64-bit content hash of a device buffer (plan-cache key); result written to out[0] (device). */
int aero_hash_u64(const void* data, int64_t nbytes, uint64_t* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Row gathers / segmented reductions (deterministic, no atomics).
 * ------------------------------------------------------------------------------------------ */
/* out[i,:] = in[idx[i],:]            (mgnLayer.py:40-41 node_attr[row]; bsms_mgn.py:306 unpool)
 * if add != NULL: out[i,:] = in[idx[i],:] + add[i,:]   (bsms_mgn.py:199-200 unpool + skip)
 * idx[i] < 0 reads a zero row (Unpool.forward of bistride_ops, pyc orig :102: non-selected fine rows stay 0) */
int aero_gather_rows(const void* in, const int32_t* idx, const void* add, void* out,
                     int64_t n_out, int64_t width, int dtype, void* stream);
/* out[n,:] = scale(n) * sum_{k in [ptr[n],ptr[n+1])} in[list ? list[k] : k, :]
 *   mean=0: scale = 1 (scatter_add, mgnLayer.py:146); mean=1: scale = 1/max(count,1)
 *   (scatter_mean, mgnLayer.py:144, bsms_mgn.py:265-272,283).  fp32 accumulation in list order.
 *   out_dtype may differ from in_dtype (fp32 aggregates of bf16 rows). */
int aero_segment_reduce(const void* in, const int32_t* ptr, const int32_t* list, void* out,
                        int64_t n_seg, int64_t width, int in_dtype, int out_dtype, int mean,
                        void* stream);
/* same, output rows ld_out elements apart (ld_out >= width): lets several reductions fill column blocks of one
 * matrix, e.g. the [N,256] gradient of the gathered projections P_s | P_d (autograd of mgnLayer.py:40-41,101). */
int aero_segment_reduce_ld(const void* in, const int32_t* ptr, const int32_t* list, void* out,
                           int64_t n_seg, int64_t width, int64_t ld_out, int in_dtype, int out_dtype,
                           int mean, void* stream);
/* backward of the mean/sum reduce w.r.t. `in`: g_in[i,:] = scale(seg[i]) * g_out[seg[i],:] */
int aero_segment_bcast(const void* g_out, const int32_t* seg_of_row, const int32_t* ptr,
                       void* g_in, int64_t n_rows, int64_t width, int dtype, int mean, void* stream);

/* Several strided 2-D copies with dtype conversion in ONE launch: packs the reference-named parameters of a processor
 * step (mgnLayer.py:26-30,68-91,124-132 -> nn.Linear / LayerNorm tensors) into the packed vectors below, and scatters
 * the packed gradients back into per-parameter buffers.  src == NULL fills the destination with zeros.  `segs` is a
 * HOST array (copied into the kernel's parameter space). */
#define AERO_MAX_COPY_SEGS 48
typedef struct aero_copy_seg {
  const void* src;
  void* dst;
  int64_t rows, cols;       /* `cols` contiguous elements per row                     */
  int64_t src_ld, dst_ld;   /* elements between consecutive rows                      */
  int32_t src_dtype, dst_dtype;   /* AERO_F32 | AERO_BF16                              */
} aero_copy_seg;
int aero_multi_copy(const aero_copy_seg* segs, int n_segs, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused MGN block (edge block or node block of one processor step).
 *
 * forward, per row r (an edge in receiver-CSR order, or a node):
 *     h0 = act( main[r] @ W_main^T + P[idx0[r], poff0:poff0+128] (+ P[idx1[r], poff1:...]) )
 *     h_l = act( h_{l-1} @ W_l^T + b_l )                    l = 1..L
 *     y   = h_L @ W_out^T + b_out ;  u = LayerNorm(y) * gamma + beta   (mlp.py:41-49, eps 1e-5)
 *     out[r] = resid[r] + u                                  (mgnLayer.py:205 / :211)
 *   and, when agg != NULL (edge block), agg[n] = sum of out[r] over the CSR rows of receiver n
 *   (mgnLayer.py:146), accumulated in CSR order in fp32.
 *
 *   The first Linear of the reference block (mgnLayer.py:26-30 EdgeBlock.mlp.layers[0],
 *   :97-103 EdgeBlockSum, :151-153 NodeBlock) is split as  [W_main | W_gathered...]: the
 *   gathered part is pre-projected per node into P (sum trick of mgnLayer.py:97-103), bias
 *   included, by the caller with a plain GEMM.
 *
 * Packed fp32 weights `w` (floats):  W_main[128*128] | W_1..W_L [L*128*128] | W_out[128*128]
 *                                    | b_1..b_L [L*128] | b_out[128] | gamma[128] | beta[128] | g_b0[128]
 * (g_b0: unused in `w`; in `g_w` it receives the gradient of the first Linear's bias = column sums of g_h0)
 * all matrices in nn.Linear layout [out][in].  aero_block_prepare turns them into the image
 * the chosen path reads (SIMT: fp32 + transposes; UMMA: bf16 swizzled shared-memory images).
 * ------------------------------------------------------------------------------------------ */
#define AERO_BLOCK_AGG_NO_CLEAR 1

typedef struct aero_block_desc {
  int32_t dtype;      /* AERO_F32 | AERO_BF16: storage of main/resid/out/P/g_* rows            */
  int32_t path;       /* AERO_PATH_SIMT | AERO_PATH_UMMA                                       */
  int32_t L;          /* number of hidden Linear(128,128) layers (config.yaml:48-49 -> 2)      */
  int32_t act;        /* AERO_ACT_*                                                            */
  int32_t use_ln;     /* LayerNorm on the block output (mlp.py:34-35)                          */
  int32_t main_f32;   /* 1: `main` (and g_main) rows are fp32 regardless of dtype (node block:
                         main = fp32 aggregate)                                                */
  int32_t has_resid_grad; /* bwd: 1 -> g_main = g_out + g_h0 @ W_main (edge: resid == main)    */
  int32_t flags;      /* AERO_BLOCK_*: bit 0 = forward does not clear `agg` first (a launch over a sub-range of
                         the rows of one mesh, e.g. boundary edges after the halo arrived, adds its receivers
                         to an aggregate an earlier launch over the other rows already started)               */
  int64_t rows;       /* E or N                                                                */
  int64_t n_nodes;    /* N (rows of P, agg)                                                    */
  int64_t ldp;        /* row stride of P in elements (384)                                     */
  int64_t poff0, poff1; /* column offsets inside a P row                                       */
  const void* main;   /* [rows,128]                                                            */
  const float* main_scale; /* optional [rows] fp32 multiplier on main rows ('mean': 1/deg)     */
  const void* resid;  /* [rows,128] (edge: == main; node: x); NULL = no residual (standalone block) */
  const void* P;      /* [n_nodes, ldp]                                                        */
  const int32_t* idx0;/* [rows] or NULL = identity                                             */
  const int32_t* idx1;/* [rows] or NULL = no second gathered term                              */
  const int32_t* rowptr; /* receiver CSR [n_nodes+1], needed when agg != NULL                  */
  const void* prepared;  /* output of aero_block_prepare                                       */
  void* out;          /* fwd: [rows,128]                                                       */
  float* agg;         /* fwd: optional [n_nodes,128] fp32                                      */
  /* backward only */
  const void* g_out;  /* [rows,128] gradient w.r.t. out                                        */
  const float* g_agg; /* optional [n_nodes,128] fp32, added as g_agg[idx1[r]] (edge block)     */
  void* g_main;       /* [rows,128] gradient w.r.t. main (may alias g_out)                     */
  void* g_h0;         /* [rows,128] gradient w.r.t. the first pre-activation                   */
  float* g_w;         /* packed like `w`; W_main slot is left untouched (caller: g_h0^T@main)  */
  void* workspace;
  size_t workspace_bytes;
  /* optional [rows,128] rows of the first hidden activation h_0 = act(main W_main^T + gathered P), latent dtype.
   * fwd: when non-NULL the kernel also stores h_0 there.  bwd: when non-NULL the kernel reads h_0 from there instead
   * of recomputing it (main, P and idx0 are then not read and P may be NULL; the trade is +256 B/row kept from
   * the forward for one GEMM, one gather and one epilogue less per backward tile).  AERO_PATH_UMMA only. */
  void* h0;
  /* fwd, optional, AERO_PATH_UMMA with main_f32: [rows,128] latent-dtype copy of the rows the first GEMM consumed,
   * i.e. round(main[r] * main_scale[r]) -- the node block's (scaled) aggregate in the latent dtype, which the caller's
   * weight-gradient GEMM  g_h0^T @ main  needs (saves a separate cast / scale pass over the fp32 aggregate). */
  void* main_lat;
  /* optional [rows,128] rows of the hidden activations H_1, H_2 (latent dtype; AERO_PATH_UMMA, L == 2, together with
   * h0).  fwd: stored when non-NULL.  bwd: read instead of recomputing the hidden layers (two GEMM phases and two
   * epilogues per tile less; a memory-for-time policy: +256 B/row each, kept per step by the caller). */
  void* h_hidden[2];
} aero_block_desc;

size_t aero_block_prepared_bytes(int L, int path);
int aero_block_prepare(const float* w, int L, int path, void* prepared, void* stream);
size_t aero_block_workspace_bytes(const aero_block_desc* d, int backward);
int aero_block_fwd(const aero_block_desc* d, void* stream);
int aero_block_bwd(const aero_block_desc* d, void* stream);
/* number of kernels the last fwd/bwd call on this thread launched (bench.py gpu_launches) */
int aero_last_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Bistride pooling index kernels (bit-exact integer work).
 *
 * aero_stride_pool_plan: bsms_mgn.py:231-262.  batch must be ascending.  Nodes of each graph
 * are ordered by pos[:,0] (ties: ascending node id, the CPU torch.argsort result), rank//stride
 * is the coarse id inside the graph, graphs are concatenated.
 *   posx        [N] float64 (caller widens pos[:,0] exactly) or NULL -> index order (:245)
 *   fine_to_coarse [N] int64 out;  coarse_batch [>= ceil-sum] int64 out (caller sizes it N)
 *   counts      [2] device int64 out: [0] = total coarse nodes
 * aero_coarsen_edges: bsms_mgn.py:274-288.  key = f2c[row]*max(Nc,1)+f2c[col]; unique sorted
 * keys -> coarse_edge_index, inverse map, and group lists for the edge-latent mean.
 *   coarse_edge_index [2,E] int64 out (first Ec columns valid, row stride E)
 *   inverse     [E] int64 out
 *   gptr        [E+1] int32 out, glist [E] int32 out: edges of coarse edge u =
 *               glist[gptr[u]..gptr[u+1]) ascending
 *   counts      [2] device int64: [0] = Ec
 * ------------------------------------------------------------------------------------------ */
size_t aero_stride_pool_workspace_bytes(int64_t N);
int aero_stride_pool_plan(const int64_t* batch, const double* posx, int64_t N, int64_t stride,
                          int64_t* fine_to_coarse, int64_t* coarse_batch, int64_t* counts,
                          void* workspace, size_t workspace_bytes, void* stream);
size_t aero_coarsen_edges_workspace_bytes(int64_t E);
int aero_coarsen_edges(const int64_t* edge_index, int64_t E, const int64_t* fine_to_coarse,
                       int64_t Nc, int64_t* coarse_edge_index, int64_t* inverse,
                       int32_t* gptr, int32_t* glist, int64_t* counts,
                       void* workspace, size_t workspace_bytes, void* stream);
/* group lists for a many-to-one int64 map (fine_to_coarse): members of group c ascending */
size_t aero_group_lists_workspace_bytes(int64_t n);
int aero_group_lists(const int64_t* group_of, int64_t n, int64_t n_groups,
                     int32_t* gptr, int32_t* glist, int32_t* group32,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * BFS-bistride pooling and WeightedEdgeConv (the reference's bistride_ops module, shipped only as
 * models/__pycache__/bistride_ops.cpython-311.pyc; "orig :NN" = first line of the code object in it).
 * Index kernels are bit-exact (last row).
 * ------------------------------------------------------------------------------------------ */
/* BistridePooling.bfs_distance (orig :21): dist[v] = number of sender->receiver hops from `start`, -1 when
 * unreachable.  sptr/sperm/dst = the sender-CSR of aero_graph_plan_build.  One kernel launch per BFS level; a call
 * runs levels [level_begin, level_begin+level_count) and writes status[0] = size of the next frontier (0 = finished),
 * status[1] = next level (device int64[2]).  level_begin == 0 initialises dist and the workspace; later calls must
 * pass the same workspace.  Replaces the reference's Python deque loop with one .item() per edge. */
size_t aero_bfs_levels_workspace_bytes(int64_t N);
int aero_bfs_levels(const int32_t* sptr, const int32_t* sperm, const int32_t* dst, int64_t N, int64_t E,
                    int64_t start, int64_t level_begin, int64_t level_count, int64_t* dist, int64_t* status,
                    void* workspace, size_t workspace_bytes, void* stream);
/* BistridePooling.select_bistride_nodes (orig :56) after the BFS: selected = ascending ids with even dist >= 0;
 * if fewer than 0.3*N of them, every reached node instead.  index_map[i] = rank of i in `selected` or -1
 * (MultiScaleGraphPreprocessor.create_multiscale_graph, bsms_mgn pyc orig :32).  counts[0] = number selected,
 * counts[1] = 1 when the fallback fired (device int64[2]). */
size_t aero_bistride_select_workspace_bytes(int64_t N);
int aero_bistride_select(const int64_t* dist, int64_t N, int64_t* selected, int64_t* index_map, int64_t* counts,
                         void* workspace, size_t workspace_bytes, void* stream);
/* coarse edges of create_multiscale_graph (bsms_mgn pyc orig :32): edges whose two endpoints are selected, renumbered
 * through index_map, self-loops dropped, caller order kept.  out_edge_index is [2,E] with row stride E (first
 * counts[0] columns valid); kept_ids (optional) = caller edge id of each kept edge; counts[1] = out-of-range ids. */
size_t aero_filter_edges_workspace_bytes(int64_t E);
int aero_filter_edges(const int64_t* edge_index, int64_t E, const int64_t* index_map, int64_t N,
                      int64_t* out_edge_index, int32_t* kept_ids, int64_t* counts,
                      void* workspace, size_t workspace_bytes, void* stream);

/* WeightedEdgeConv (orig :131-209):
 *     len_e = || pos[dst_e] - pos[src_e] ||                                   (compute_edge_weights, orig :152)
 *     w_e   = sigmoid( W2 relu( W1 [x[src_e] ; x[dst_e] ; len_e] + b1 ) + b2 )     hidden width 64
 *     out[n] = sum (or mean) over edges e with dst_e == n of  w_e * (x[src_e] Wt^T + bt)      (forward, orig :173)
 * The first Linear is split like the sum trick of mgnLayer.py:97-103: the caller pre-projects the node rows with one
 * plain GEMM into Q = x [W1_src ; W1_dst ; Wt]^T + [0 ; b1 ; bt], row = [A(64) | B(64) | T(out_dim)], so an edge needs
 * A[src] + B[dst] + w1_len * len.  With compute_w == 0 the weights are read from `w` (the weight reuse of the up pass,
 * BSMSGMP.forward bsms_mgn pyc orig :145) and Q holds only T (ldq >= out_dim).
 * Backward: dQ (same layout as Q) for the caller's GEMMs; compute_w: g_small = d(w1_len)[64] | d(W2)[64] | d(b2);
 * otherwise g_w = gradient of the given weights.  g_w_ext = gradient arriving at the returned weights (may be NULL).
 * Gradient w.r.t. pos is not produced.  Fixed accumulation order (receiver-CSR / sender-CSR), fp32 registers. */
typedef struct aero_wec_desc {
  int32_t dtype;      /* AERO_F32 | AERO_BF16: rows of Q, out, g_out, dQ and the weights w               */
  int32_t mean;       /* aggr == 'mean'                                                                   */
  int32_t compute_w;  /* 1: weights from the edge-weight MLP (written to w); 0: weights read from w       */
  int32_t pos_dim;
  int64_t N, E, out_dim, ldq;
  const void* Q;      /* [N, ldq]                                                                         */
  const float* pos;   /* [N, pos_dim] fp32 (compute_w)                                                    */
  const float* w1_len;/* [64] fp32: last input column of edge_weight_mlp.0.weight                         */
  const float* w2;    /* [64] fp32: edge_weight_mlp.2.weight                                              */
  const float* b2;    /* [1]  fp32 (device)                                                               */
  const int32_t *rowptr, *src, *dst, *perm, *sptr, *sperm;   /* graph plan                                */
  void* w;            /* [E] caller edge order                                                            */
  void* out;          /* fwd: [N, out_dim]                                                                */
  const void* g_out;  /* bwd: [N, out_dim]                                                                */
  const void* g_w_ext;/* bwd: optional [E] caller order                                                   */
  void* dQ;           /* bwd: [N, ldq]                                                                    */
  void* g_w;          /* bwd, compute_w == 0: [E] caller order                                            */
  float* g_small;     /* bwd, compute_w == 1: [129] fp32                                                  */
  void* workspace;
  size_t workspace_bytes;
} aero_wec_desc;
size_t aero_wec_workspace_bytes(const aero_wec_desc* d, int backward);
int aero_wec_fwd(const aero_wec_desc* d, void* stream);
int aero_wec_bwd(const aero_wec_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------
 * Weight-gradient reduction over rows (tcgen05, warp-specialised, fed by TMA):
 *     dW[128 a, 128] (fp32, contiguous) = A[rows, 128 a]^T  B[rows, 128]        a_panels = 1 or 2
 *   A, B bf16 row matrices (row strides lda / 128 elements).  The first Linear's weight gradient of a block
 *   (mgnLayer.py:97-103 backwards: g_h0^T e, g_h0n^T agg) and the projection weight gradients
 *   [g_P_s | g_P_d]^T x, g_h0n^T x.  Optionally (seg_out != NULL; rows in receiver-CSR order, dst = receiver of each
 *   row, rowptr = its CSR over n_nodes) the receiver sums of A's first 128 columns are taken from the same
 *   shared-memory tiles: seg_out[n, 0:128] (bf16, row stride seg_ld) = sum of A rows [rowptr[n], rowptr[n+1]) --
 *   the gradient of the receiver-side pre-projection P_d (mgnLayer.py:99), fp32 accumulation in CSR order.
 * ------------------------------------------------------------------------------------------ */
size_t aero_wgrad_workspace_bytes(int64_t rows, int a_panels, int with_seg);
int aero_wgrad(const void* A, int64_t lda, int a_panels, const void* B, int64_t rows, float* dW,
               const int32_t* dst, const int32_t* rowptr, int64_t n_nodes, void* seg_out, int64_t seg_ld,
               void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Row GEMMs of a processor step outside the fused blocks (tcgen05, warp-specialised, fed by TMA; bf16 rows):
 *     out[rows, 128 nb] = sum_{ka < na} A_ka[rows, 128] . Wtile(ka, nb_i)  (+ bias[128 nb])  (+ add[rows, 128 nb])
 *   a_blocks[ka] / a_ld[ka]: the K-blocks, each a [rows,128] column block of a row matrix (row stride a_ld[ka]).
 *   W: na * nb row tiles of 128 x 128 bf16, contiguous.  w_mn == 0: tile nb_i holds W[out, in] (K contiguous) --
 *   out = A W^T, nn.Linear's layout: the sum-trick pre-projection  P = x [W_s; W_d; W_nx]^T + b  (na = 1, nb = 3;
 *   mgnLayer.py:97-103 applied per node).  w_mn == 1: tile ka holds W[k, out] (outputs contiguous) -- out = A W:
 *   the gradient of the node latents through that projection,  g_x = [g_P_s | g_P_d | g_h0n] W + G_x  (na = 3,
 *   nb = 1), fp32 accumulation over K = 384 and one rounding.  Exactly one of na / nb may exceed 1 (na * nb <= 3).
 *   out / add rows may be column blocks of wider matrices (row strides out_ld / add_ld, multiples of 8).
 * ------------------------------------------------------------------------------------------ */
int aero_row_gemm(const void* const* a_blocks, const int64_t* a_ld, int na, const void* W, int w_mn, int nb,
                  const void* bias, const void* add, int64_t add_ld, void* out, int64_t out_ld, int64_t rows,
                  void* stream);

/* ------------------------------------------------------------------------------------------
 * The encoders' first Linear on raw features (mgn.py:123-124; mlp.py:40-44): K = in_features <= 16, 128 outputs.
 *   fwd: out[rows,128] = x[rows,K] W^T + b      (x row stride ldx; W [128,K], b [128] or NULL; all of `dtype`)
 *   bwd: dwb[128][K+1] (fp32) <- d(W)[c][k] = sum_r g[r][c] x[r][k]  and, in slot k = K, d(b)[c] = sum_r g[r][c]
 *        from ONE pass over the gradient rows g[rows,128]; deterministic.
 * ------------------------------------------------------------------------------------------ */
int aero_thin_linear_fwd(const void* x, int64_t ldx, const void* W, const void* b, void* out, int64_t rows, int K,
                         int dtype, void* stream);
size_t aero_thin_linear_workspace_bytes(int64_t rows, int K);
int aero_thin_linear_bwd(const void* g, const void* x, int64_t ldx, float* dwb, int64_t rows, int K, int dtype,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Training-step tail (utils.py:191-195, train.py:207-211, :222).
 *
 * aero_mse_loss_grad: loss[0] = loss_scale * sum (pred - target)^2 and grad = grad_scale * (pred - target) in one pass
 *   (nn.MSELoss mean reduction + its backward: loss_scale = 1/(rows*cols), grad_scale = 2/(rows*cols)); `pred` /
 *   `grad` rows may be strided (ld_* elements, e.g. the first `cols` columns of a 128-wide decoder tile); target is
 *   fp32 contiguous [rows, cols]; the loss stays on the device (no per-batch host sync); fixed-order reduction.
 * aero_adam_step: ONE launch updates every parameter tensor listed in the device-resident table `segs_device`
 *   (torch.optim.Adam semantics: g += weight_decay * w; m, v moments; w -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)).
 *   A segment with grad == NULL is skipped; master != NULL keeps an fp32 master copy of a bf16 parameter.
 *   steps_in / steps_out: device int32 [n_segs], the number of updates each tensor has had (torch keeps one step count
 *   per parameter); the launch reads steps_in and writes steps_out (two distinct arrays the caller swaps).
 * ------------------------------------------------------------------------------------------ */
typedef struct aero_adam_seg {
  void* param;        /* [n] p_dtype, updated in place            */
  const void* grad;   /* [n] g_dtype, or NULL (no gradient)       */
  float* m;           /* [n] fp32 first moment                    */
  float* v;           /* [n] fp32 second moment                   */
  float* master;      /* [n] fp32 master weights, or NULL         */
  int64_t n;
  int32_t p_dtype, g_dtype;
} aero_adam_seg;
size_t aero_mse_workspace_bytes(void);
int aero_mse_loss_grad(const void* pred, const float* target, void* grad, float* loss, int64_t rows, int64_t cols,
                       int64_t ld_pred, int64_t ld_grad, int dtype, float loss_scale, float grad_scale,
                       void* workspace, size_t workspace_bytes, void* stream);
int aero_adam_step(const aero_adam_seg* segs_device, int n_segs, int64_t max_elems, const int32_t* steps_in,
                   int32_t* steps_out, float lr, float beta1, float beta2, float eps, float weight_decay, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AERO_GNN_H_ */
