/*
 * temp_b200 -- C ABI of the B200-native RGCN + GRU/BiGRU/attention forward path of TeMP.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.
 * Every entry point replaces a call the reference makes into third-party native code
 * (DGL 0.4.1 libdgl, cuBLAS, cuDNN) -- the reference line each one stands in for is cited on it.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host; the caller owns every buffer;
 *   - all floating point is fp32, all indices int32;
 *   - kernels are enqueued on the cudaStream_t passed in (as void*), never synchronise, never
 *     allocate; no global mutable state other than the thread-local last-error string;
 *   - every function returns TEMP_OK (0) or a negative TEMP_E* code; no exception crosses the ABI;
 *   - one CUDA device per process (the launchers cache per-kernel attributes and occupancy once per process), which is
 *     how the path is deployed: one process per GPU under torch.distributed.
 *
 * "packed rows": the nodes of all snapshot instances of a window batch are laid out back to back,
 * step-major (temp_b200/planner.py); a row index addresses one (instance, node).
 */
#ifndef TEMP_B200_H_
#define TEMP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TEMP_ABI_VERSION 20

#define TEMP_OK 0
#define TEMP_EINVAL (-1)   /* bad argument (null pointer, unsupported size, ...)      */
#define TEMP_ECUDA (-2)    /* a CUDA runtime call / launch failed; see last error     */
#define TEMP_EUNSUPPORTED (-3)

#define TEMP_MAX_D 256     /* embed_size == hidden_size upper bound (SIMT path and the 64-row tcgen05 kernels) */
#define TEMP_MAX_TERMS 3
#define TEMP_MAX_SCAN_STEPS 16
#define TEMP_MAX_PUSH_PEERS 16

#define TEMP_ACT_NONE 0
#define TEMP_ACT_RELU 1

#define TEMP_CELL_TORCH_GRU 0 /* torch.nn.GRU equations, reference models/RRGCN.py:75,84      */
#define TEMP_CELL_TYPE1 1     /* reference models/GRU_cell.py:18-31 (--type1)                 */

/* One dense term  acc += scale_row(A[gather(row)]) . W   of the fused layer kernel.
 *   a        [*, d]   source rows
 *   a_index  nullable: source row of packed row r is a[a_index[r]]; a negative index is a zero row.
 *            When null the source row is r itself.
 *   a_dt     nullable: per-row time gap; the gathered row is multiplied by
 *              exp(-a_dt[r] * inv_temperature)                     (decay_wb == null)
 *              exp(-max(decay_wb[0] * a_dt[r] + decay_wb[1], 0))   (learnable lambda)
 *            reference models/RRGCN.py:83, models/RGCN.py:106-107.
 *   w        [d, d] row-major (in x out), e.g. loop_weight / time_weight as stored.
 *   w_packed nullable: the same matrix in the tensor-core operand image written by temp_pack_weights
 *            (d % 4 == 0, d <= 256).  When every operand of a launch has its packed image the layer runs on the
 *            tcgen05 path (3xTF32 split, fp32-level accuracy: the 128-row tile kernel for d == 128 with 1x1 relation
 *            blocks, the 64-row tile kernel of tc_wide.cu for every other width and for 2x2 / 4x4 blocks -- e.g.
 *            d = 200, n_bases = 100 as the reference ships ICEWS05-15); otherwise on the fp32 SIMT path.   */
typedef struct {
  const float* a;
  const int32_t* a_index;
  const float* a_dt;
  const float* decay_wb;
  const float* w;
  const void* w_packed;
} TempDenseTerm;

/* Fused RGCN layer over packed rows [row0, row1):
 *   agg_v  = norm_v * sum_{e: dst_e = v} norm_v * blockdiag(W[rel_e]) . x[src_e]      (RGCN.py:91-104)
 *   out_v  = act( agg_v (or x_v when residual) + sum_t term_t(v) + h_bias )           (RGCN.py:53-70, 78-86)
 *   h_out  = out_v (+ time_embed[row_time[v]] if te_out)                              (RGCN.py:47-51, RRGCN.py:202-203)
 *   chain  = (out_v (+ te if te_chain)) . chain_w + chain_b                           (input half of the GRU:
 *            RRGCN.py:84 gi = x W_ih^T + b_ih; attention q/k/v: SARGCN.py:30-32)
 * Replaces: torch index_select + bmm + DGL update_all/fn.sum/apply (RGCN.py:92-104), torch.mm self
 * loop (RGCN.py:57), and the input GEMM of torch.nn.GRU.                                       */
typedef struct {
  int32_t row0, row1;
  int32_t d;                 /* feature width (in == out)                                        */
  /* graph part (row_ptr == null: no aggregation, e.g. forward_isolated)                         */
  const int32_t* row_ptr;    /* [rows+1] CSR by destination over packed rows (absolute offsets)  */
  const int32_t* e_src;      /* [E] feature row of the edge source inside x                       */
  const int32_t* e_rel;      /* [E]                                                               */
  const int32_t* e_dst;      /* [E] nullable: destination packed row of the edge (unused by the current kernels)   */
  const float* norm;         /* [rows] 1/in_degree (0 when in_degree is 0)                        */
  const float* x;            /* [*, d] source features for the aggregation                        */
  const float* weight;       /* [2*num_rels, n_bases*si*so] block-diagonal relation weights        */
  int32_t n_bases, si, so;
  int32_t residual;          /* 1: acc starts from the row's own term-0 input (RGCN.py:83)        */
  int32_t n_terms;
  TempDenseTerm terms[TEMP_MAX_TERMS];
  const float* h_bias;       /* nullable [d]                                                      */
  int32_t activation;
  const float* time_embed;   /* nullable [T, d]                                                   */
  const int32_t* row_time;   /* [rows] time_embed row per packed row (needed when te_* set)       */
  int32_t row_time_scalar;   /* used instead of row_time when row_time == null (isolated pass)    */
  int32_t te_out, te_chain;
  float* h_out;              /* nullable [rows, d]                                                */
  const float* chain_w;      /* nullable [d, chain_n] row-major (i.e. W_ih transposed)            */
  const float* chain_b;      /* nullable [chain_n]                                                */
  float* chain_out;          /* [rows, chain_ld]; columns [0, chain_n) are written                */
  int32_t chain_n, chain_ld;
  const void* chain_w_packed; /* nullable: chain_w in the temp_pack_weights image (d == 128 with 1x1 blocks: chain_n % 128 == 0) */
  float inv_temperature;
  float* agg_scratch;        /* nullable [>= row1 rows, d] scratch indexed by packed row: the tcgen05 path runs the
                                aggregation as its own HBM-bound launch and hands it over through this buffer      */
  /* Work lists of that launch (agg_lists != 0; otherwise every row of [row0, row1) is tested through row_ptr and
   * gets one warp): entries (packed row, first edge, end edge) of the rows of [row0, row1) with in-edges --
   * agg_rows: one warp per row; agg_heavy: high in-degree rows, one thread block per row (its 8 warps sum
   * contiguous edge chunks, the partial sums are added in chunk order: deterministic, no atomics).            */
  const int32_t* agg_rows;   /* [n_agg_rows, 3]  */
  const int32_t* agg_heavy;  /* [n_agg_heavy, 3] */
  int32_t n_agg_rows, n_agg_heavy;
  int32_t agg_lists;
  /* Snapshot-sharded forward (tcgen05 path): the chained output row r is stored to chain_peers[chain_owner[r]] instead
   * of chain_out -- the GPU that will scan the row's chain partition -- over NVLink peer memory, from the tile kernel's
   * chain epilogue (chain_peers: DEVICE array of peer-mapped bases of the same [rows, chain_ld] buffer, this rank's own
   * included; chain_owner: rank per packed row).  Null: chain_out as above.                                         */
  float* const* chain_peers;
  const int32_t* chain_owner;
  int32_t chain_world, reserved2;
} TempRgcnLayerArgs;

/* Recurrent half of the GRU for packed rows [row0, row1), fused with the gates:
 *   h0 = decay(state[prev_row[r]])  (zero when prev_row[r] < 0)                       (RRGCN.py:79-83)
 *   gh = h0 . whh_t + b_hh ; gates with gi (precomputed by the layer kernel's chain)   (Appendix A.3)
 *   out[r] (+)= h' (+ time_embed[row_time[r]])                                        (RRGCN.py:85, 202-203)
 * Replaces torch.nn.GRU / cuDNN RNN forward (RRGCN.py:84, BiRRGCN.py:34-43) and models/GRU_cell.py. */
typedef struct {
  int32_t row0, row1;
  int32_t d;
  const float* gi;           /* [rows, gi_ld]; this cell's input pre-activations start at gi_off   */
  int32_t gi_ld, gi_off;
  const float* state;        /* [*, d] rows addressed by prev_row                                 */
  const int32_t* prev_row;   /* [rows] or null (no previous state at all)                         */
  const float* dt;           /* [rows] nullable                                                   */
  const float* decay_wb;     /* nullable (learnable lambda: weight, bias)                         */
  float inv_temperature;
  const float* whh_t;        /* [d, 3d] row-major = weight_hh transposed                          */
  const void* whh_packed;    /* nullable: whh_t in the temp_pack_gru_weights image (d % 4 == 0, d <= 256) */
  const float* b_hh;         /* [3d]                                                              */
  int32_t cell_type;
  const float* time_embed;   /* nullable                                                          */
  const int32_t* row_time;
  int32_t row_time_scalar;
  int32_t accumulate;        /* 1: out += h' (second direction of the Bi centre step)             */
  float* out;                /* [rows, d]                                                         */
  int32_t out_index_is_row;  /* reserved, must be 1                                               */
  int32_t part_col;          /* column of this step in the scan's chain-partition table (see below) */
  int32_t push;              /* 1: the values this step writes are also stored to every peer (TempGruScanArgs.push_*) */
} TempGruArgs;

/* All GRU steps of a window in ONE launch (persistent CTAs, W_hh^T slices resident in shared memory
 * across steps).  Replaces the python time-step loop of pre_forward (DynamicRGCN.py:163-173) for the
 * serial half of the recurrence.
 *   parts  : chain-partition table [n_parts][part_stride][2] int32 (temp_b200/planner.py).  The recurrence
 *            couples a packed row only with the row of the same entity in the same batch item at the previous
 *            step (DynamicRGCN.py:35-54), so a batch item cut into entity-id ranges gives independent
 *            chains: entry (p, steps[s].part_col) = the packed-row range [lo, hi) of partition p at step s,
 *            at most part_rows rows (lo == hi: nothing).  The tcgen05 path (d == 128) gives every partition to
 *            one 4-CTA cluster (to one of its one or two pipelines when part_rows <= 48) and separates the steps of a
 *            partition by cluster-scope barriers; without the table (or d != 128) steps are separated by a
 *            grid-wide barrier (cooperative launch: gru_scan_tcw_kernel with the W_hh slices in tensor memory when every
 *            step has its packed image and d <= 224, else the fp32 SIMT gru_scan_kernel), or run as one
 *            gru_step_tcw_kernel launch per step (224 < d <= 256).
 *   barrier: 8 bytes of device memory for the grid-wide barrier (more with barrier_words, below), zero before the
 *            first use; the kernel leaves it zeroed.  Launches sharing one barrier word must not run concurrently. */
typedef struct {
  int32_t n_steps;
  int32_t n_parts;
  uint32_t* barrier;
  const int32_t* parts;
  int32_t part_stride;
  /* Fused all-gather of the final-layer states over NVLink peer memory (tcgen05 path): steps with `push` set store
   * every value they write to row r also to  push_bufs[k] + push_offset + (r - push_row0) * d  for all k <
   * push_world, where push_bufs is a DEVICE array of peer-mapped buffer bases (this rank's own buffer included), e.g.
   * from torch.distributed._symmetric_memory.  The caller separates the launch from the consumers on the other GPUs
   * with a cross-GPU barrier.  A "peer" may also be pinned host memory (UVA): the final states then reach the host
   * from inside the scan kernel, without a device-to-host copy after it.  push_bufs == null: no peer stores.       */
  int32_t push_world;
  float* const* push_bufs;
  int64_t push_offset;
  int32_t push_row0;
  int32_t part_rows;         /* upper bound of the rows of one partition step in `parts` (0: the legacy bound, 96).  Tables cut
                                at <= 48 rows whose steps all use ONE recurrent cell run on gru_scan_tm_kernel (W_hh in
                                tensor memory, one or two partition pipelines per CTA), the others on gru_scan_tc_kernel        */
  float* push_multicast;     /* nullable: NVLS multicast address of the same symmetric buffer -- one multimem.st per
                                value instead of one store per peer (the switch replicates it to every GPU); push_bufs then
                                lists only buffers OUTSIDE the multicast group (e.g. a pinned host buffer; push_world >= 0) */
  int32_t barrier_words;     /* zeroed 32-bit words behind `barrier` (0 or 2: the two barrier words only).  With
                                2 + n_steps * ceil(max step rows / 64) words or more, gru_scan_tcw_kernel (d != 128) replaces
                                the grid-wide barrier between steps by per-tile completion counters: a tile step waits only
                                for the tiles of the previous step that hold its prev_row rows.  Left zeroed by the kernel. */
  int32_t reserved3;
  TempGruArgs steps[TEMP_MAX_SCAN_STEPS];
} TempGruScanArgs;

/* Multi-head attention over the time axis for packed rows [row0, row1) (SARGCN.py:25-53):
 *   q, k_cur, v_cur = qkv[r, 0:d], [d:2d], [2d:3d]; history slot s of row r lives at
 *   kv_hist[slot_row[r*n_slots + s]] (k | v), negative = entity inactive (mask -1e10 -> weight 0);
 *   out[r, j*heads + head] = softmax_s(q.k_s / sqrt(dk) - decay_s) . v_s          (Appendix A.5)     */
typedef struct {
  int32_t row0, row1;
  int32_t d, heads;
  const float* qkv;          /* [rows, 3d]                                                        */
  const float* kv_hist;      /* [*, 2d]                                                           */
  const int32_t* slot_row;   /* [rows, n_slots]                                                   */
  int32_t n_slots;
  const float* tau;          /* [n_slots + 1] slot ages (last = current slot)                     */
  const float* decay_wb;     /* nullable                                                          */
  int32_t combine_max;       /* 1: out = max(out, result) (JK max over layers, SARGCN.py:117)     */
  float* out;                /* [rows, d]                                                         */
} TempAttnArgs;

/* out[i] = table[index[i]] rows, or zero rows for negative indices. */
typedef struct {
  int32_t n, d;
  const float* table;
  const int32_t* index;
  float* out;
} TempGatherArgs;

/* dst[index[i]] = src[i]  (scatter of active rows into the all-entity table, DynamicRGCN.py:62-63;
 * optional "+ te_row" broadcast used by the zero-history de-duplication of the isolated pass).     */
typedef struct {
  int32_t n, d;
  const float* src;
  const int32_t* src_index;  /* nullable: read src[src_index[i]]                                   */
  const int32_t* dst_index;  /* nullable: write dst[i]                                             */
  const float* add_row;      /* nullable [d] added to every written row                            */
  float* dst;
} TempScatterArgs;

#define TEMP_SCORE_DISTMULT 0   /* utils/scores.py:4-11   */
#define TEMP_SCORE_COMPLEX 1    /* utils/scores.py:27-44  */
#define TEMP_SCORE_TRANSE 2     /* utils/scores.py:46-55  */

/* Training-time link prediction of one target graph, fused:
 *   loss[p] = logsumexp_c score(p, c) - score(p, 0)     for the positive triples p = 0..n_pos-1
 * where candidate c of triple p is row cand[p * n_cand + c] of the all-entity table (column 0 = the true entity,
 * models/TKG_Module.py:202-213 with the all-zero labels of utils/CorrptTriples.py:40); corrupt_tail: the candidates
 * replace the object (score mode 'tail'), else the subject (mode 'head').  The mean over p is F.cross_entropy.
 * Replaces the [n_pos, n_cand, d] gather all_embeds_g[neg_samples], calc_score and F.cross_entropy -- the gather is
 * never materialised (770 MB per direction at 3000 positives x 501 candidates).  d % 32 == 0.               */
typedef struct {
  int32_t n_pos, n_cand, d;
  int32_t score_fn, corrupt_tail;
  const float* ent_embed;    /* [n_nodes, d] states of the target graph's nodes (local ids)                  */
  const float* rel_embeds;   /* [2 * num_rels, d]                                                            */
  const float* table;        /* [num_ents, d] all-entity table of the graph (get_all_embeds_Gt)              */
  const int64_t* triples;    /* [n_pos, 3] (subject, relation, object), local node ids                       */
  const int64_t* cand;       /* [n_pos, n_cand] global entity ids                                            */
  float* loss;               /* [n_pos]                                                                      */
} TempScoreLossArgs;

/* Backward of temp_score_loss_fwd (training through the fused scorer; the reference back-propagates through the
 * materialised [n_pos, n_cand, d] gather of models/TKG_Module.py:202-213):  given grad_loss[p] = dL / d loss[p], ADDS
 *   dL/d table[cand[p, c]]  +=  grad_loss[p] * (softmax_c(score(p, .)) - [c == 0]) * d score / d candidate row
 *   dL/d ent_embed[fixed_p], dL/d rel_embeds[r_p]  +=  the same through the triple's query vector
 * into the three caller-zeroed gradient buffers (vector float atomics: summation order is not fixed).  The scores are
 * recomputed from the inputs (nothing is saved by the forward).  Same shapes and restrictions as the forward.     */
typedef struct {
  int32_t n_pos, n_cand, d;
  int32_t score_fn, corrupt_tail;
  const float* ent_embed;
  const float* rel_embeds;
  const float* table;
  const int64_t* triples;
  const int64_t* cand;
  const float* grad_loss;    /* [n_pos]                                                                      */
  float* grad_ent_embed;     /* [n_nodes, d]      accumulated into                                           */
  float* grad_rel_embeds;    /* [2 * num_rels, d] accumulated into                                           */
  float* grad_table;         /* [num_ents, d]     accumulated into                                           */
} TempScoreLossBwdArgs;

/* Filtered ranking of one target graph's test triples against all entities:
 *   rank[q] = 1 + #{ j != target[q] :  v(q, j) > v(q, target[q])  or  (v(q, j) == v(q, target[q]) and j < target[q]) }
 * with v(q, j) = sigmoid(score(q, j)) for ordinary entities and sigmoid(-10e6) = 0 for the entities of the query's filter
 * list -- the position of the target in the stable descending sort of utils/evaluation.py:53-106
 * (perturb_and_get_rank: calc_score over all entities, torch.where(mask, -10e6, score), sigmoid, sort_and_rank;
 * "+ 1" of calc_metrics_single_graph line 49).  Replaces the [n_query, num_ents] score / mask / sort tensors; the filter
 * is a CSR list per query (mask_eval_set lines 82-99: known true answers of the same (timestamp, query), target
 * excluded, each id once) instead of a dense byte mask.  corrupt_tail: score mode 'tail' (candidates replace the
 * object), else 'head'.  d % 32 == 0, d <= 256.                                                             */
typedef struct {
  int32_t n_query, num_ents, d;
  int32_t score_fn, corrupt_tail;
  const float* ent_embed;    /* [n_nodes, d] states of the target graph's nodes (local ids)                  */
  const float* rel_embeds;   /* [2 * num_rels, d]                                                            */
  const float* table;        /* [num_ents, d] all-entity table of the graph (get_all_embeds_Gt)              */
  const int64_t* triples;    /* [n_query, 3] (subject, relation, object), local node ids                     */
  const int64_t* target;     /* [n_query] global id of the true entity on the corrupted side                 */
  const int32_t* filter_ptr; /* [n_query + 1] offsets into filter_ids (nullable: nothing filtered)           */
  const int32_t* filter_ids; /* global entity ids scored as -10e6                                            */
  int64_t* rank;             /* [n_query] 1-indexed                                                          */
} TempRankArgs;

enum { TEMP_OP_LAYER = 1, TEMP_OP_GRU = 2, TEMP_OP_ATTN = 3, TEMP_OP_GATHER = 4, TEMP_OP_SCATTER = 5,
       TEMP_OP_MEMCPY_H2D = 6, TEMP_OP_MEMCPY_D2H = 7, TEMP_OP_GRU_SCAN = 8 };

typedef struct {
  void* dst;
  const void* src;
  uint64_t bytes;
} TempCopyArgs;

/* One entry of a launch program (a static work list built once per window plan). */
typedef struct {
  int32_t kind;
  int32_t reserved;
  union {
    TempRgcnLayerArgs layer;
    TempGruArgs gru;
    TempGruScanArgs scan;
    TempAttnArgs attn;
    TempGatherArgs gather;
    TempScatterArgs scatter;
    TempCopyArgs copy;
  } u;
} TempOp;

/* ---- native window planner (temp_b200/csrc/planner.cpp) ------------------------------------------------------------
 * Host side: packs a batch of windows into the arrays above.  Replaces the per-step dgl.batch / history dictionary
 * work of models/DynamicRGCN.py:35-110, models/BiDynamicRGCN.py:17-100, models/TKG_Module.py:232-250.               */
typedef struct {
  int32_t time, n_nodes, n_edges;
  const int32_t* node_ids;   /* [n_nodes] global entity ids, ascending (utils/dataset.py:168)                     */
  const int32_t* row_ptr;    /* [n_nodes + 1] CSR by destination                                                  */
  const int32_t* csr_src;    /* [n_edges] local source node, edges of a destination in edge-id order               */
  const int32_t* csr_rel;    /* [n_edges]                                                                         */
  const float* norm;         /* [n_nodes] 1 / in_degree, 0 for in_degree 0 (utils/utils.py:74-79)                 */
} TempSnapshotView;

typedef struct {
  int32_t rows, edges, n_segments, n_instances, n_parts, n_agg_rows, n_agg_heavy, n_slots, batch, seq_len;
  int32_t scan_tile;   /* row bound of a partition step actually used (temp_plan_window's scan_tile < 0: chosen, 48 or 64) */
} TempPlanCounts;

enum { TEMP_PLAN_ENT_ID = 0, TEMP_PLAN_ROW_TIME, TEMP_PLAN_NORM, TEMP_PLAN_ROW_PTR, TEMP_PLAN_E_SRC, TEMP_PLAN_E_SRC_ENT,
       TEMP_PLAN_E_REL, TEMP_PLAN_PREV_A, TEMP_PLAN_DT_A, TEMP_PLAN_PREV_B, TEMP_PLAN_DT_B, TEMP_PLAN_SLOT_ROW,
       TEMP_PLAN_SCAN_PARTS, TEMP_PLAN_AGG_ROWS, TEMP_PLAN_AGG_HEAVY,
       TEMP_PLAN_INSTANCES /* int32 [n_instances][7] = item, step, dir (0 f, 1 b, 2 centre), time, row0, n, snapshot */,
       TEMP_PLAN_SEGMENTS  /* int32 [n_segments][6] = kind (0 hist_f, 1 hist_b, 2 final), step, row0, row1, inst0, inst1 */,
       TEMP_PLAN_LAST_F, TEMP_PLAN_LAST_B /* int32 [batch] instance of the last history step per item, -1 */,
       TEMP_PLAN_STEPS_F, TEMP_PLAN_STEPS_B /* int32 [seq_len - 1][batch] instance per (step, window row), -1 */ };

typedef struct TempPlan TempPlan;
/* snaps: every snapshot of the split in graph_dict order (ascending time); targets: snapshot index of each target
 * timestamp (any order).  Returns null on bad arguments.  The plan owns host memory until temp_plan_destroy.      */
TempPlan* temp_plan_window(const TempSnapshotView* snaps, int32_t n_snaps, const int32_t* targets, int32_t batch,
                           int32_t seq_len, int32_t bidirectional, int32_t attention, int32_t scan_tile,
                           int32_t heavy_degree);
void temp_plan_destroy(TempPlan* plan);
int temp_plan_counts(const TempPlan* plan, TempPlanCounts* out);
const void* temp_plan_array(const TempPlan* plan, int32_t which, int64_t* n_bytes);
/* Device blob of a plan: arrays TEMP_PLAN_ENT_ID .. TEMP_PLAN_AGG_HEAVY back to back, starts aligned to `align` bytes;
 * offsets[i] = -1 for arrays the plan does not have.  Returns the blob size (negative on bad arguments).           */
int64_t temp_plan_blob_layout(const TempPlan* plan, int64_t* offsets, int64_t* sizes, int32_t align);
int temp_plan_write_blob(const TempPlan* plan, uint8_t* dst, int32_t align);

/* Host side of the negative sampler (utils/CorrptTriples.py:61-85), continuing NumPy's global legacy generator: mt_key
 * [624] / mt_pos are the MT19937 state of np.random.get_state() and are advanced in place (hand them back with
 * np.random.set_state).  Slot s (tail corruption of triple s/2 for even s, head corruption for odd s) draws rounds of
 * `neg` candidates -- np.random.randint(num_entities, size=neg) -- until it holds `neg` of them outside its filter list
 * fids[fptr[s] .. fptr[s+1]); they are written to out_even / out_odd + (s/2) * out_stride.  sort_min_f: smallest
 * filter-list length for which numpy's in1d takes its merge-sort path (ceil(10 * neg ** 0.145)).  Returns the number of
 * rounds drawn, TEMP_EINVAL on bad arguments.  No CUDA.                                                            */
int64_t temp_negative_sample(uint32_t* mt_key, int32_t* mt_pos, int64_t num_entities, int32_t neg, const int64_t* fptr,
                             const int64_t* fids, int32_t n_slots, int32_t sort_min_f, int64_t* out_even, int64_t* out_odd,
                             int64_t out_stride);

int temp_abi_version(void);
const char* temp_last_error_string(void);
/* sm count, max dynamic shared memory per block, compute capability (major*10+minor) */
int temp_device_info(int32_t* sm_count, int32_t* max_smem, int32_t* cc);

int temp_rgcn_layer_fwd(const TempRgcnLayerArgs* args, void* stream);
/* The aggregation half alone (tcgen05 path; d % 4 == 0, d <= 256, 1x1 / 2x2 / 4x4 relation blocks): agg_scratch[v] =
 * norm_v^2 * sum_e blockdiag(W[rel_e]) x[src_e] for the rows of [row0, row1) with in-edges -- the DGL update_all / fn.sum /
 * apply of models/RGCN.py:91-104. */
int temp_rgcn_gather_fwd(const TempRgcnLayerArgs* args, void* stream);
int temp_gru_fwd(const TempGruArgs* args, void* stream);
int temp_gru_scan_fwd(const TempGruScanArgs* args, void* stream);
int temp_attention_fwd(const TempAttnArgs* args, void* stream);
int temp_gather_rows(const TempGatherArgs* args, void* stream);
int temp_scatter_rows(const TempScatterArgs* args, void* stream);
/* out[c, r] = in[r, c]  (weight preparation: weight_ih / weight_hh / q,k,v -> K-major-first)     */
int temp_score_loss_fwd(const TempScoreLossArgs* args, void* stream);
int temp_score_loss_bwd(const TempScoreLossBwdArgs* args, void* stream);
int temp_rank_filtered_fwd(const TempRankArgs* args, void* stream);
int temp_transpose(const float* in, int32_t rows, int32_t cols, float* out, int32_t out_ld, void* stream);
/* Tensor-core operand images.  A [k, n] row-major fp32 matrix (k % 4 == 0, k <= 256) is split into tf32 hi / lo
 * parts and stored, per 128 output features x 32 k, in the K-major SWIZZLE_128B shared-memory layout the tcgen05
 * kernels fetch with cp.async.bulk (temp_b200/csrc/tc_common.cuh); k is padded to whole 32-wide k-atoms and n to whole
 * 128-feature blocks with zeros (for k == 128, n % 128 == 0 nothing is padded).
 * temp_pack_gru_weights packs whh_t [d, 3 d] per block of 32 hidden columns (rows r|z|n|zero pad).
 * *_bytes return the image size, or a negative value for shapes without a tensor-core path.  */
int64_t temp_packed_weights_bytes(int32_t k, int32_t n);
int temp_pack_weights(const float* w_kn, int32_t k, int32_t n, void* packed, void* stream);
int64_t temp_packed_gru_bytes(int32_t d);
int temp_pack_gru_weights(const float* whh_t, int32_t d, void* packed, void* stream);
/* Runs ops[0..n) back to back on one stream (memcpy ops use cudaMemcpyAsync; host pointers must be
 * pinned for the copies to be asynchronous).  Returns the first failure.                          */
int temp_run_program(const TempOp* ops_host, int32_t n, void* stream);
/* Cross-GPU completion barrier of the fused all-gather (one tiny launch): stores `seq` into slot `rank` of every peer's
 * flag array (st.release.sys through the peer-mapped pointers flag_peers[0..world)), then waits until all `world` slots
 * of this GPU's own array flags_local hold a value >= seq (ld.acquire.sys).  seq must grow by one per step; the kernels
 * whose peer stores it publishes precede it on `stream`.                                                              */
int temp_peer_barrier(uint32_t* flags_local, uint32_t* const* flag_peers, int32_t world, int32_t rank, uint32_t seq,
                      void* stream);
/* The same program as ONE CUDA graph (copies, cluster launch and programmatic-dependent-launch edges included): captured
 * once on a private stream, launched into any stream with a single driver call.  The caller keeps every buffer the ops
 * point to alive and unchanged in address; run the program once with temp_run_program first (lazy kernel attributes). */
int temp_graph_create(const TempOp* ops_host, int32_t n, void** graph_exec_out);
int temp_graph_launch(void* graph_exec, void* stream);
int temp_graph_destroy(void* graph_exec);
/* Number of kernel launches temp_run_program(ops, n) issues (copies excluded); negative on a bad program. */
int temp_program_kernel_count(const TempOp* ops_host, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* TEMP_B200_H_ */
