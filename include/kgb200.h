/* kgb200.h - C ABI of libkgb200.so: the B200 (sm_100a) message-passing hot path behind
 * keras-geometric's MessagePassing / GCNConv / GINConv / SAGEConv / GATv2Conv API.
 *
 * The reference (/root/reference, pure Python over keras.ops) has no FFI; the entry points
 * below are what a native plugin for its hot path binds.  Each one cites the reference code
 * it replaces (paths relative to /root/reference/src/keras_geometric/).
 *
 * Conventions
 *  - every function returns 0 (KGB_OK) or a negative KGB_ERR_* code; kgb_last_error() returns a
 *    thread-local message.  Nothing throws, allocates user-visible memory or synchronises.
 *  - every pointer is a BORROWED DEVICE pointer owned by the caller (outputs and workspaces
 *    included) unless its name ends in _host.  `device` is the CUDA ordinal the pointers live
 *    on, `stream` the cudaStream_t to launch on (0 = legacy default stream).
 *  - indices are int32, row pointers int64, data float32, row-major, leading dimension in
 *    elements.  Edge lists use the reference's COO convention: row 0 = source j,
 *    row 1 = target i; aggregation is segmented by target (layers/message_passing.py:191-212).
 */
#ifndef KGB200_H
#define KGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* kgb_stream_t; /* cudaStream_t */

enum {
  KGB_OK = 0,
  KGB_ERR_INVALID = -1,     /* bad argument */
  KGB_ERR_CUDA = -2,        /* a CUDA runtime call / launch failed */
  KGB_ERR_WORKSPACE = -3,   /* workspace too small */
  KGB_ERR_UNSUPPORTED = -4  /* shape outside the compiled kernel set */
};

enum { KGB_OP_SUM = 0, KGB_OP_MEAN = 1, KGB_OP_MAX = 2, KGB_OP_MIN = 3,
       KGB_OP_MAX_RAW = 4 /* keras segment_max itself: empty segment = -inf, no inf -> 0 rewrite */,
       KGB_OP_SQDEV = 5   /* second pass of the std aggregator (layers/aggregators.py:174-232): with the row means
                             in `addend`, out = sqrt(max(sum_k (x[col_k] - mean_row)^2 / max(count, 1e-8), 0)),
                             0 where count <= 1 */ };
enum { KGB_ACT_NONE = 0, KGB_ACT_RELU = 1 };

/* status bits written by kgb_csr_build into *status (device int32) */
enum { KGB_STATUS_OOB_INDEX = 1 };

int kgb_version(void);
const char* kgb_last_error(void);
/* number of kernels this library has launched in the process so far (all threads) */
int64_t kgb_launch_count(void);
/* SM count, compute capability and L2 size of `device` (host-side query, cached). */
int kgb_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, int64_t* l2_bytes);

/* ---------------------------------------------------------------------------------------
 * K1  COO -> CSR grouping (stable), degrees, optional appended self-loops.
 * Replaces: utils/main.py:8-16 (add_self_loops), the implicit per-target grouping inside
 * keras.ops.segment_sum/segment_max (aggregators.py:67-72,108,135), and the ones->segment_sum
 * degree count (utils/main.py:23-24, aggregators.py:66-69).
 *
 * Edges e in [0,E) come from edge_index (int32 [2,E], row 0 = src, row 1 = dst); edges
 * e in [E, E+n_loops) are the self-loops (e-E)->(e-E), appended last exactly like the
 * reference.  by_source = 0 groups by target (CSR used by the forward), 1 groups by source
 * (the transposed structure used by the backward).  Inside a segment edges keep their
 * original order (what a sequential scatter on the CPU visits), so
 *   perm   = argsort(key, stable)           int32 [E+n_loops]
 *   col    = other_endpoint[perm]           int32 [E+n_loops]
 *   deg    = bincount(key, n_seg)           int32 [n_seg]
 *   rowptr = exclusive_cumsum(deg)          int64 [n_seg+1]
 * bit-exactly.  Indices outside [0,n_seg) / [0,n_val) set KGB_STATUS_OOB_INDEX in *status
 * (the edge is clamped to 0 so nothing faults); the caller decides when to read status.
 * ------------------------------------------------------------------------------------- */
size_t kgb_csr_build_workspace_bytes(int64_t n_edges_total, int64_t n_seg);
int kgb_csr_build(int device, const int32_t* edge_index, int64_t E, int by_source,
                  int64_t n_seg, int64_t n_val, int64_t n_loops,
                  int64_t* rowptr, int32_t* col, int32_t* perm, int32_t* deg, int32_t* status,
                  void* ws, size_t ws_bytes, kgb_stream_t stream);

/* Hub table for load balance: every row with more than `threshold` edges is cut into chunks of
 * `chunk` edges that are reduced independently and merged in chunk order (deterministic).
 * counts[0] = n_hubs, counts[1] = n_chunks (device int32[2], zeroed by this call).
 * Capacities: hub_* arrays hold max_hubs entries, chunk_hub holds max_chunks entries, where
 * max_hubs = n_edges/threshold + 1 and max_chunks = n_edges/chunk + max_hubs + 1. */
int kgb_csr_hubs(int device, const int64_t* rowptr, int64_t n_seg, int32_t threshold, int32_t chunk,
                 int32_t* hub_row, int32_t* hub_chunk_base, int32_t* hub_nchunks, int32_t* chunk_hub,
                 int64_t max_hubs, int64_t max_chunks, int32_t* counts, kgb_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K2  GCN symmetric normalisation.  Replaces utils/main.py:20-33.
 *   dis[i] = (float(deg[i]) + 1e-12)^-0.5   (IEEE 1/sqrt, == torch.pow(.,-0.5) on the host)
 *   w[e]   = dis[dst[e]] * dis[src[e]]      for e in COO order incl. the n_loops self-loops
 * deg is the in-degree over targets including self-loops (from kgb_csr_build, by_source=0).
 * `w` may be NULL (only dis wanted).
 * ------------------------------------------------------------------------------------- */
int kgb_gcn_norm(int device, const int32_t* deg, int64_t n_nodes, const int32_t* edge_index,
                 int64_t E, int64_t n_loops, float* dis, float* w, kgb_stream_t stream);
/* out[r,f] = y[r,f] > 0 ? g[r,f] : 0 - backward of the ReLU fused into the gather / GEMM epilogues; out is dense [rows,F] */
int kgb_relu_bwd(int device, const float* g, int64_t ldg, const float* y, int64_t ldy, int64_t rows, int32_t F,
                 float* out, kgb_stream_t stream);
/* Same with the bias gradient in the same pass: out[r,f] = (y ? y[r,f] > 0 : true) ? g[r,f] : 0 and
 * partial[p, f] = sum over the rows of CTA p of out[r,f] (p < n_parts = kgb_colsum_parts(device, rows));
 * kgb_reduce_parts(partial, n_parts, F, bias_grad) adds them in order (deterministic).  y == NULL: no mask
 * (plain column sums; out may then be NULL too).  F % 4 == 0, F <= 1024, 16-byte aligned rows. */
int32_t kgb_colsum_parts(int device, int64_t rows);
int kgb_relu_bwd_colsum(int device, const float* g, int64_t ldg, const float* y, int64_t ldy, int64_t rows, int32_t F,
                        float* out, int64_t ldo, float* partial, int32_t n_parts, kgb_stream_t stream);
/* Softmax cross-entropy over integer labels (the training step's loss; one warp per row, C <= 1024):
 *   fwd: row_loss[r] = logsumexp(logits[r,:]) - logits[r, labels[r]]
 *   bwd: dlogits[r,c] = (softmax(logits[r,:])[c] - (c == labels[r])) * scale * grad_loss[0]   (grad_loss: device scalar) */
int kgb_softmax_xent_fwd(int device, const float* logits, int64_t ld, const int64_t* labels, int64_t rows, int32_t C,
                         float* row_loss, kgb_stream_t stream);
int kgb_softmax_xent_bwd(int device, const float* logits, int64_t ld, const int64_t* labels, int64_t rows, int32_t C,
                         const float* grad_loss, float scale, float* dlogits, int64_t ldd, kgb_stream_t stream);
/* Row-wise L2 normalisation, keras.ops.normalize(axis=-1, order=2) of SAGEConv(normalize=True)
 * (layers/sage_conv.py:432-433): y[r,:] = x[r,:] / max(||x[r,:]||_2, eps), norm[r] = ||x[r,:]||_2 (saved for the
 * backward).  bwd: gx = (g - y <g, y>) / max(norm, eps) where norm >= eps, g / eps otherwise.  Rows of up to 1024
 * floats are kept in registers (one read), wider rows are read twice. */
int kgb_l2_normalize(int device, const float* x, int64_t ldx, int64_t rows, int32_t F, float eps, float* y, int64_t ldy,
                     float* norm, kgb_stream_t stream);
int kgb_l2_normalize_bwd(int device, const float* g, int64_t ldg, const float* y, int64_t ldy, const float* norm,
                         int64_t rows, int32_t F, float eps, float* gx, int64_t ldgx, kgb_stream_t stream);
/* out[k] = in[perm[k]]  (bring COO-ordered edge weights into CSR / CSC slot order) */
int kgb_permute_f32(int device, const float* in, const int32_t* perm, int64_t n, float* out,
                    kgb_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K3/K4/K5/K9  destination-segmented gather-reduce (deterministic, no atomics).
 * Replaces: ops.take + Sum/Mean/Max/MinAggregator.aggregate (layers/message_passing.py:195-212,
 * layers/aggregators.py:56-167), the weighted GCN message (layers/gcn_conv.py:233-246) and its
 * bias update (:266-272).  With col = perm and x = messages it is the generic
 * Aggregator.aggregate(messages, target_idx, dim_size) (aggregators.py:24-39); over the
 * by_source structure it is the backward of sum/mean.
 *
 *   r        = row_ids ? row_ids[s] : s                      for slot s in [0, n_rows)
 *   red[s,:] = OP_{k in [rowptr[s], rowptr[s+1])}  edge_w[k] * src_scale[col[k]] * x[col[k], :]
 *   MEAN     : red / max(float(deg), 1e-8)                   (aggregators.py:77-81)
 *   MAX/MIN  : empty segment or +-inf result -> 0            (aggregators.py:108-112,162-167)
 *   out[r,:] = act( out_scale[s] * red + addend_scale * addend[r,:] + bias )
 * edge_w, src_scale, out_scale, addend, bias, row_ids, arg are optional (NULL).
 * MAX/MIN additionally write arg[r,f]: the source row id of the unique extremum, -1 when the
 * gradient is zero (empty / inf / nan), -2 when several edges tie (backward re-walks the row).
 * hub_* / partial: optional hub table from kgb_csr_hubs plus a float workspace of
 * kgb_gather_reduce_partial_bytes(); when absent every row is reduced by one lane group.
 * ------------------------------------------------------------------------------------- */
struct kgb_halo_push_args;
typedef struct kgb_gather_reduce_args {
  const float* x;          /* [n_src_rows, F] gathered matrix                  */
  int64_t ldx;
  int64_t n_src_rows;
  int32_t F;
  int32_t op;              /* KGB_OP_*                                          */
  const int64_t* rowptr;   /* [n_rows+1]                                        */
  const int32_t* col;      /* [nnz] row ids into x                              */
  int64_t n_rows;
  const int32_t* row_ids;  /* optional [n_rows] output row of each CSR row      */
  const float* edge_w;     /* optional [nnz] per-slot weight                    */
  const float* src_scale;  /* optional [n_src_rows] per-gathered-row factor     */
  const float* out_scale;  /* optional [n_rows] per-output-row factor           */
  const float* addend;     /* optional [*, F] added after the reduction         */
  int64_t ld_addend;
  float addend_scale;
  const float* bias;       /* optional [F]                                      */
  int32_t act;             /* KGB_ACT_*                                         */
  float* out;              /* [*, F]                                            */
  int64_t ldo;
  int32_t* arg;            /* optional [*, F] (MAX/MIN only), same ld as out    */
  /* hub table (all NULL/0 when unused) */
  const int32_t* hub_row;
  const int32_t* hub_chunk_base;
  const int32_t* hub_nchunks;
  const int32_t* chunk_hub;
  int32_t n_hubs;
  int32_t n_chunks;
  int32_t hub_threshold;
  int32_t hub_chunk;
  float* partial;          /* workspace, kgb_gather_reduce_partial_bytes()      */
  int32_t* work;           /* optional int32[2], ZERO on entry and left zero on exit: dynamic task
                              queue (load balance on skewed graphs).  One buffer must not be shared
                              by launches that can overlap in time.  NULL = static striding.     */
  const int32_t* unit_order; /* optional (used with `work`): a PERMUTATION of the
                              ceil(n_rows / kgb_gather_unit_rows()) row units, in the order the queue
                              hands them out.  Heaviest-first (by edge count, hub rows excluded) removes
                              the tail on power-law graphs; the result does not depend on it.   */
  /* ---- split source / split output (partitioned graphs: [owned rows | halo rows] without a concatenated copy) ---- */
  const float* x2;         /* optional: column ids >= n_split_src read row (id - n_split_src) of x2             */
  int64_t ldx2;
  int64_t n_split_src;     /* ignored when x2 is NULL                                                           */
  float* out2;             /* optional: output rows >= n_split_out go to row (r - n_split_out) of out2 ...      */
  int64_t ldo2;
  int64_t n_split_out;
  const struct kgb_halo_push_args* out2_push; /* ... or, when given, straight into the peers' windows: row
                              (r - n_split_out) is slot s of the push table (only slot_begin / dst / dst_row0 /
                              ldd / n_peers are read).  This is the fused "transposed gather + halo-gradient
                              exchange" kernel: the rows travel over NVLink while the rest is still being reduced.
                              Split-output rows take no addend / bias / activation / arg.                        */
  const int32_t* col_hot;  /* optional copy of `col` with bit 31 set for "hot" rows: those are loaded with an L2
                              evict_last hint, all others with evict_first.  Mark the most frequently gathered rows
                              that together fit in the L2 (power-law graphs: the hubs); the result does not depend on
                              it.  Measured on C4: +3 % at F = 256, a loss for narrower rows - off by default.        */
  /* ---- dropout of the gathered rows, fused (training only) ----
   * The reference drops the per-edge messages element-wise before they are weighted and reduced (GCNConv:
   * layers/gcn_conv.py:238-242, SAGEConv: layers/sage_conv.py:295-297) and materialises [E, F] tensors to do so.
   * Here element f of edge e survives when Philox4x32-10(counter = (edge id, f / 4), key = drop_seed) >= drop_p * 2^32
   * and is scaled by 1 / (1 - drop_p); nothing is stored - the transposed (backward) pass regenerates the same mask
   * from edge_id[] = original edge id of each slot (`perm` of the structure that is walked).  sum / mean only. */
  float drop_p;            /* 0 = no dropout                                                                      */
  uint64_t drop_seed;
  const int32_t* edge_id;  /* [nnz], required when drop_p > 0                                                      */
} kgb_gather_reduce_args;

/* The mask the kernels apply, for tests: out[e, f] = 1 / (1 - p) or 0 for edge ids edge_id[e] (NULL: e itself);
 * per_head != 0 selects the stream used by kgb_gatv2_* (one decision per edge and head).                         */
int kgb_dropout_mask(int device, const int32_t* edge_id, int64_t n_edges, int32_t F, int32_t per_head, float p,
                     uint64_t seed, float* out, kgb_stream_t stream);

/* rows per task-queue unit of kgb_gather_reduce (unit u covers rows [u*R, (u+1)*R)) */
int32_t kgb_gather_unit_rows(void);

size_t kgb_gather_reduce_partial_bytes(int32_t n_chunks, int32_t F, int32_t op);
int kgb_gather_reduce(int device, const kgb_gather_reduce_args* a, kgb_stream_t stream);

struct kgb_hub_table;
/* Backward of MAX/MIN (torch.scatter_reduce(amax) semantics: the gradient is split evenly
 * between tied maxima).  gx must be zero-initialised by the caller; accumulates
 *   gx[arg[r,f], f] += g[r,f]                       (unique extremum)
 *   gx[col[k],  f] += g[r,f] / n_ties               for every tied edge (arg == -2)
 * `out` is the forward result, x the forward input.  Same CSR as the forward.
 * Deterministic: several targets may select the same source entry, so the contributions are accumulated as 64-bit
 * FIXED-POINT integers (grid 2^-s chosen from max|g| and the row count so that no sum can overflow; integer addition
 * is associative, the order in which the atomics land cannot change a bit of the result) in the caller-provided
 * workspace `acc_ws` (kgb_gather_max_bwd_acc_bytes(n_src_rows, F), contents arbitrary on entry) and converted to
 * fp32 once per element at the end.  NaN / +-inf gradients bypass the grid (their sum is order-independent anyway). */
size_t kgb_gather_max_bwd_acc_bytes(int64_t n_src_rows, int32_t F);
int kgb_gather_max_bwd(int device, const float* g, int64_t ldg, const int32_t* arg,
                       const float* out, int64_t ldo, const float* x, int64_t ldx,
                       const int64_t* rowptr, const int32_t* col, const int32_t* row_ids,
                       int64_t n_rows, int32_t F, int32_t op, float* gx, int64_t ldgx,
                       int64_t n_src_rows, void* acc_ws,
                       const struct kgb_hub_table* hubs, kgb_stream_t stream);
/* `hubs` (optional) lets tied entries of hub rows be resolved chunk-parallel; its `partial` workspace must
 * hold kgb_gather_max_bwd_workspace_bytes(n_hubs, n_chunks, F) bytes. */
size_t kgb_gather_max_bwd_workspace_bytes(int32_t n_hubs, int32_t n_chunks, int32_t F);

/* out[r,:] = scale * src[idx[r],:]   (row gather: backward of the generic segment-sum w.r.t.
 * materialised messages, and the halo "pack" step of the partitioned path). idx may be NULL
 * (identity). */
int kgb_gather_rows(int device, const float* src, int64_t lds, const int32_t* idx, int64_t n_out,
                    int32_t F, float scale, float* out, int64_t ldo, kgb_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K7  halo exchange over peer memory (NVLink / NVSwitch), one process per GPU.
 * The reference is single-process (no counterpart); this is the exchange step the north star's partitioned path
 * needs before each layer's aggregation.  Every rank owns a WINDOW (plain cudaMalloc memory, exportable with CUDA
 * IPC); peers map it once (kgb_ipc_open) and kgb_halo_push then packs the requested feature rows and stores them
 * STRAIGHT INTO THE RECEIVERS' WINDOWS with 128-bit stores over NVLink - no staging buffer, no send/recv protocol.
 * Ordering between ranks (window free before the push / all pushes landed before the read) is the caller's
 * business (one tiny torch.distributed all-reduce on the same stream before and after).
 * ------------------------------------------------------------------------------------- */
#define KGB_MAX_PEERS 16
int kgb_window_alloc(int device, size_t bytes, void** ptr);                 /* cudaMalloc on `device`          */
int kgb_window_free(int device, void* ptr);
int kgb_ipc_export(int device, const void* ptr, unsigned char handle[64]);  /* cudaIpcGetMemHandle             */
int kgb_ipc_open(int device, const unsigned char handle[64], void** ptr);   /* map a peer's window, enables
                                                                                peer access from `device`       */
int kgb_ipc_close(int device, void* ptr);
typedef struct kgb_halo_push_args {
  const float* src;        /* [*, F] rows of this rank, leading dimension lds                              */
  int64_t lds;
  const int32_t* idx;      /* optional [n_slots]: slot s sends row idx[s]; NULL: slot s sends row s        */
  int32_t F;
  int32_t n_peers;         /* world size (<= KGB_MAX_PEERS)                                                 */
  int64_t slot_begin[KGB_MAX_PEERS + 1]; /* slots [slot_begin[p], slot_begin[p+1]) go to peer p            */
  float* dst[KGB_MAX_PEERS];             /* peer p's window region (device pointer valid on this device)   */
  int64_t dst_row0[KGB_MAX_PEERS];       /* first row of this rank's block inside peer p's region          */
  int64_t ldd;             /* leading dimension of the window rows (floats)                                 */
  int64_t slot_rot;        /* kgb_halo_push visits the slots in 256-slot chunks in a strided (golden-ratio)
                              order that starts at the chunk holding this slot: one rank's stores are spread over
                              all receivers at any moment, and with slot_rot = slot_begin[(rank + 1) % n_peers]
                              the ranks start at different receivers - no GPU is stored into by everyone at once */
} kgb_halo_push_args;
int kgb_halo_push(int device, const kgb_halo_push_args* a, kgb_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K6  fused GATv2 edge kernel.  Replaces layers/gatv2_conv.py:241-264 (_gatv2_propagate edge
 * part), :268-289 (_compute_attention), :291-311 (_softmax_by_target), :313-335
 * (_aggregate_messages) and the bias add of :337-352.
 *   s[k,h]   = sum_c att[h,c] * leaky_relu(hdst[i,h,c] + hsrc[col[k],h,c], slope)
 *   alpha    = exp(s - max_i) / (sum_i exp(s - max_i) + 1e-10)     per target i and head h
 *   out[i,h,:] = sum_k alpha[k,h] * hsrc[col[k],h,:]  (+ bias[h*C+c] when bias != NULL)
 * One pass over the CSR with an online max/sum; rowmax/rowden ([n_dst,H]) are saved for the
 * backward, which recomputes alpha instead of storing [nnz,H] tensors.
 * ------------------------------------------------------------------------------------- */
/* Hub table of the structure a GATv2 kernel walks (from kgb_csr_hubs) + its workspaces.  Optional:
 * NULL (or n_hubs == 0) makes one lane group reduce every row whole. */
typedef struct kgb_hub_table {
  const int32_t* hub_row;
  const int32_t* hub_chunk_base;
  const int32_t* hub_nchunks;
  const int32_t* chunk_hub;
  int32_t n_hubs;
  int32_t n_chunks;
  int32_t threshold;
  int32_t chunk;
  float* partial;   /* kgb_gatv2_partial_bytes(n_chunks, H, C) bytes                              */
  int32_t* work;    /* optional int32[64], zero on entry / left zero: dynamic task queue scratch */
  const int32_t* unit_order; /* optional (GATv2 kernels, with `work`): permutation of the
                       ceil(n_rows / kgb_gatv2_unit_rows()) row units, heaviest first               */
} kgb_hub_table;
int32_t kgb_gatv2_unit_rows(void);
size_t kgb_gatv2_partial_bytes(int32_t n_chunks, int32_t H, int32_t C);

/* Attention dropout (layers/gatv2_conv.py:252-253), fused: alpha of (edge e, head h) survives when
 * Philox4x32-10(counter = (edge id, h / 4), key = seed) >= p * 2^32 and is scaled by 1 / (1 - p); the softmax still
 * normalises over all edges.  edge_id = `perm` of the structure the call walks (CSR for fwd / bwd_dst, the transposed
 * one for bwd_src), so all three passes regenerate the same mask.  NULL or p == 0: no dropout. */
typedef struct kgb_gat_dropout {
  float p;
  uint64_t seed;
  const int32_t* edge_id;
} kgb_gat_dropout;

int kgb_gatv2_fwd(int device, const float* hsrc, const float* hdst, int64_t n_src, int64_t n_dst,
                  int32_t H, int32_t C, const float* att, float slope,
                  const int64_t* rowptr, const int32_t* col, const float* bias,
                  float* out, float* rowmax, float* rowden, const kgb_gat_dropout* drop,
                  const kgb_hub_table* hubs, kgb_stream_t stream);
/* Backward, pass 1 over the forward CSR (per target): g_hdst[i] (written), the per-(target, head) record
 * stat[i,h] = (rowmax, 1 / (rowden + 1e-10), r = sum_c g[i,h,c] * agg[i,h,c], 0) (written, workspace [n_dst,H,4],
 * 16-byte aligned: the three scalars the per-source pass needs per edge come from ONE 16-byte load) and the
 * attention-vector gradient partials g_att_part [n_parts, H*C] (n_parts from kgb_gatv2_bwd_parts()).
 * `agg` is the forward output; `bias` (optional) is the vector the forward fused into it (subtracted again). */
int kgb_gatv2_bwd_parts(int device, int64_t n_dst, int32_t H, int32_t C);
int kgb_gatv2_bwd_dst(int device, const float* g, const float* agg, const float* hsrc,
                      const float* hdst, int64_t n_src, int64_t n_dst, int32_t H, int32_t C,
                      const float* att, float slope, const int64_t* rowptr, const int32_t* col,
                      const float* rowmax, const float* rowden, const float* bias,
                      float* g_hdst, float* stat, float* g_att_part, int32_t n_parts, float* rec,
                      const kgb_gat_dropout* drop, const kgb_hub_table* hubs, kgb_stream_t stream);
/* Per-edge records (optional, `rec` above: [nnz, kgb_gatv2_rec_floats(H, C)] floats, one row per CSR slot, contents
 * arbitrary on entry): the per-target pass stores alpha_e * dropout_e and the logit gradient ds_e per head plus the
 * sign bits of z = h_i + h_j, and kgb_gatv2_bwd_src_rec computes the per-source gradient from them with ONE row
 * gather per edge (g_i) - no h rows, no logits, no exp - instead of kgb_gatv2_bwd_src's two gathers and three scalar
 * lookups.  slot_map[k'] = CSR slot of the edge in slot k' of the transposed structure.  kgb_gatv2_rec_floats
 * returns 0 for shapes without this path (more than one lane group per row: H * C / 4 > 32). */
int32_t kgb_gatv2_rec_floats(int32_t H, int32_t C);
int kgb_gatv2_bwd_src_rec(int device, const float* g, int64_t n_src, int64_t n_dst, int32_t H, int32_t C,
                          const float* att, float slope, const int64_t* colptr, const int32_t* row,
                          const int32_t* slot_map, const float* rec, const float* addend, float* g_hsrc,
                          const kgb_hub_table* hubs, kgb_stream_t stream);
/* Backward, pass 2 over the transposed structure (per source): g_hsrc[j] (written) = per-source gradient
 * (+ addend[j], optional [n_src, H*C]: on a square graph the per-target part g_hdst, saving a pass).
 * `stat` is the [n_dst,H,4] record array kgb_gatv2_bwd_dst wrote. */
int kgb_gatv2_bwd_src(int device, const float* g, const float* hsrc, const float* hdst,
                      int64_t n_src, int64_t n_dst, int32_t H, int32_t C, const float* att,
                      float slope, const int64_t* colptr, const int32_t* row,
                      const float* stat, const float* addend,
                      float* g_hsrc, const kgb_gat_dropout* drop, const kgb_hub_table* hubs, kgb_stream_t stream);
/* out[f] = sum_p part[p,f]  in fixed order (deterministic reduction of partials). */
int kgb_reduce_parts(int device, const float* part, int32_t n_parts, int32_t F, float* out,
                     kgb_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K8  dense node-feature transform on the tensor cores (tcgen05.mma, TMEM accumulators, TMA operands).
 * Replaces ops.matmul (layers/gcn_conv.py:233,335) and layers.Dense (layers/sage_conv.py:201-221,
 * layers/gin_conv.py:133-156, layers/gatv2_conv.py:95-101).
 * Every width goes through the hand-written tcgen05 kernels below (csrc/tc_gemm.cu): the host side pads the output
 * width to a multiple of 4 and loops 256-wide column slabs of the weights for wider layers; there is no library
 * (cuBLAS / CUTLASS) GEMM behind this ABI.
 * ------------------------------------------------------------------------------------- */

/* D[M,N] = A[M,K] * Wt[N,K]^T (+ C) (+ bias) (ReLU) with
 * tcgen05.mma.kind::tf32 and a 3xTF32 split (fp32-accurate), TMA operand staging, TMEM accumulators, one persistent
 * warp-specialised CTA per SM.  Wt must be pre-split by kgb_split_tf32 into wt_hi / wt_lo, each a dense
 * [kgb_linear_tc_rows(N), kgb_linear_tc2_k(K, 0)] fp32 matrix whose rows >= N and columns >= K are zero.
 * Any M >= 1 and any K (TMA zero-fills rows past M and columns past K); N <= 256 and N % 4 == 0 (pad the output),
 * 16-byte aligned pointers and leading dimensions that are multiples of 4 floats. */
int32_t kgb_linear_tc_rows(int32_t N);
/* hi = tf32-truncated W, lo = W - hi; written transposed ([cols, rows]) when transpose != 0 */
int kgb_split_tf32(int device, const float* w, int32_t rows, int32_t cols, int64_t ld, int32_t transpose,
                   float* hi, float* lo, kgb_stream_t stream);
int kgb_linear_tc(int device, const float* A, int64_t lda, int32_t M, int32_t K, const float* wt_hi,
                  const float* wt_lo, int32_t N, const float* C, int64_t ldc, const float* bias, int32_t act,
                  float* D, int64_t ldd, kgb_stream_t stream);
/* Same with an output leading dimension, to fill a column block of a wider (K-concatenated) split buffer. */
int kgb_split_tf32_ld(int device, const float* w, int32_t rows, int32_t cols, int64_t ld, int32_t transpose,
                      float* hi, float* lo, int64_t ldo, kgb_stream_t stream);
/* D = [A1 | A2] * [W1 ; W2] (+ C) (+ bias) (ReLU): two node-feature operands that share the row dimension are
 * multiplied in ONE pass (SAGEConv's lin_neigh(agg) + lin_self(x), sage_conv.py:411-433; the sum of two dX GEMMs).
 * wt_hi / wt_lo are [kgb_linear_tc_rows(N), kgb_linear_tc2_k(K1, K2)]: W1^T in columns [0, K1), zeros up to the next
 * multiple of 32, then W2^T (zero-padded to a multiple of 4 columns).  K2 == 0 (A2 NULL) is kgb_linear_tc, whose
 * row length kgb_linear_tc2_k(K, 0) is K rounded up to a multiple of 4. */
int32_t kgb_linear_tc2_k(int32_t K1, int32_t K2);
int kgb_linear_tc2(int device, const float* A1, int64_t lda1, int32_t K1, const float* A2, int64_t lda2, int32_t K2,
                   int32_t M, const float* wt_hi, const float* wt_lo, int32_t N, const float* C, int64_t ldc,
                   const float* bias, int32_t act, float* D, int64_t ldd, kgb_stream_t stream);

/* dW[Kx,N] = X[M,Kx]^T * G[M,N] with the same tcgen05 3xTF32 machinery (both operands MN-major, both split in
 * shared memory).  The node dimension is cut into kgb_linear_tc_dw_parts() contiguous slices, one per CTA; slice p
 * writes partials[p, :, :] and the caller adds the partials in order with kgb_reduce_parts (deterministic).
 * Needs M >= 1, Kx, N <= 256 and multiples of 4 (wider / ragged shapes: the caller loops 256-wide slabs over padded
 * operands), 16-byte aligned operands. */
int32_t kgb_linear_tc_dw_parts(int device, int64_t M);
int kgb_linear_tc_dw(int device, const float* X, int64_t ldx, const float* G, int64_t ldg, int32_t M, int32_t Kx,
                     int32_t N, float* partials, int32_t n_parts, kgb_stream_t stream);
/* Two weight gradients that share X in one pass: partials[p] = X^T [G1 | 0.. | G2] with G2 starting at column
 * ceil32(N1) (kgb_linear_tc_dw2_cols(N1, N2) columns in total, <= 256) - X is loaded and split once for both.
 * SAGEConv's lin_neigh / lin_self gradients when the layer aggregates after the transform (sage_conv.py:201-221). */
int32_t kgb_linear_tc_dw2_cols(int32_t N1, int32_t N2);
/* The mirror case, two feature operands that share G: partials[p] = [X1 | 0.. | X2]^T G with X2's rows starting at
 * row ceil32(Kx1) (kgb_linear_tc_dw2_cols(Kx1, Kx2) rows in total, <= 256) - G is loaded and split once.
 * SAGEConv's dW_neigh = agg^T g and dW_self = x^T g (sage_conv.py:411-433) when both inputs are at most 128 wide. */
int kgb_linear_tc_dw_x2(int device, const float* X1, int64_t ldx1, int32_t Kx1, const float* X2, int64_t ldx2, int32_t Kx2,
                        const float* G, int64_t ldg, int32_t N, int32_t M, float* partials, int32_t n_parts,
                        kgb_stream_t stream);
int kgb_linear_tc_dw2(int device, const float* X, int64_t ldx, const float* G1, int64_t ldg1, int32_t N1, const float* G2,
                      int64_t ldg2, int32_t N2, int32_t M, int32_t Kx, float* partials, int32_t n_parts,
                      kgb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* KGB200_H */
