/*
 * esr.h -- C ABI of libesr.so, the B200 (sm_100a) embedding-training hot path for the
 * ESRecsys trainers.
 *
 * The reference (BBischof/ESRecsys) is pure Python: it has NO FFI/plugin boundary.  Its hot path
 * runs through jax/flax/optax calls, so each entry point below cites the reference call it
 * replaces (paths relative to the reference root).  The reference-side binding a maintainer
 * would add is the ctypes stub shown in INTEGRATION.md.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; POD config structs start with `struct_size` so the
 *    struct can grow without breaking old callers.
 *  - Every pointer is DEVICE memory owned by the caller unless stated otherwise.  The library
 *    never allocates device memory, never keeps a pointer after return and never synchronises:
 *    all work is enqueued on the `stream` argument (a cudaStream_t passed as void*).
 *  - Return value: 0 (ESR_OK) or a negative ESR_E* code; esr_strerror() names it and
 *    esr_last_cuda_error() returns the CUDA error string behind the last ESR_ECUDA of the
 *    calling thread.  No C++ exception crosses this boundary.
 *  - Ids are NOT range-checked on the hot path (XLA clamps silently in the reference; we
 *    document "validated by the caller").  esr_check_ids_i32() is the debug validator.
 *  - Thread-compatible: no global mutable state besides the thread-local error string.
 */
#ifndef ESR_H_
#define ESR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ESR_VERSION 100 /* 0.1.0 */
#define ESR_MAX_PEERS 8 /* ranks of one NVSwitch domain addressed through peer pointers */

typedef void* esr_stream_t; /* cudaStream_t */

enum {
  ESR_OK = 0,
  ESR_EINVAL = -1,     /* bad shape / alignment / null pointer / unknown enum */
  ESR_EWORKSPACE = -2, /* workspace smaller than esr_*_workspace_bytes() */
  ESR_ECUDA = -3,      /* a CUDA call or launch failed; see esr_last_cuda_error() */
  ESR_ENOTSUP = -4,    /* valid request this build does not implement */
  ESR_ENOMEM = -5      /* a host decoder could not grow its line buffer (row valid, allocation failed) */
};

enum { ESR_OPT_ADAGRAD = 0, ESR_OPT_ADAM = 1, ESR_OPT_SGDM = 2 };
enum { ESR_BIAS_REFERENCE_BROADCAST = 0, ESR_BIAS_PER_PAIR = 1 };
enum { ESR_ROWS_UPDATE = 0, ESR_ROWS_EMIT_GRADS = 1 };
enum { ESR_IMPL_AUTO = 0, ESR_IMPL_LDG = 1, ESR_IMPL_TMA = 2 };
enum { ESR_LOSS_HINGE = 0, ESR_LOSS_SOFTMAX = 1 };

int esr_version(void);
const char* esr_strerror(int rc);
const char* esr_last_cuda_error(void);
/* sm count and compute capability of the current device. */
int esr_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------
 * Embedding table.  Replaces the flax param `{'embedding': f32[V,D]}` of nn.Embed
 * (wikipedia/models.py:16-19, spotify/models.py:30-31) plus its optimizer slot.
 *
 * Sparse (Adagrad) training keeps TWO row buffers and a per-row version byte: a step reads row r
 * from rows[ver[r]] and writes the updated row to rows[1-ver[r]]; versions of the touched rows
 * flip at the end of the step.  That makes the batch-synchronous update (every gradient taken at
 * the OLD parameters, exactly what XLA's scatter-add gives the reference) race-free without ever
 * materialising a gradient buffer: each touched row is read once and written once.
 * rows[1] == NULL and ver == NULL describe a plain dense table (dense Adam / SGD-momentum modes).
 * ------------------------------------------------------------------------------------------ */
typedef struct EsrTable {
  uint32_t struct_size;
  int32_t D;        /* row width in floats; D % 4 == 0, rows 16-byte aligned */
  int64_t V;        /* rows */
  float* rows[2];   /* [V*D] each */
  uint8_t* ver;     /* [V] or NULL */
  float* acc;       /* [V*D] Adagrad accumulator, or NULL */
  float* bias;      /* [V] GloVe bias (wikipedia/models.py:18-19), updated in place, or NULL */
  float* bias_acc;  /* [V] or NULL */
} EsrTable;

/* out[k, :] = current row ids[k].  Replaces jnp.take behind nn.Embed.__call__
 * (wikipedia/models.py:31-34, spotify/models.py:43-44). */
int esr_table_gather_f32(const EsrTable* t, const int32_t* ids, int64_t n, float* out,
                         esr_stream_t stream);
/* out[V*D] = dense current table (what state.params['..']['embedding'] holds in the reference). */
int esr_table_export_f32(const EsrTable* t, float* out, esr_stream_t stream);
/* out[b] = sum_d x[b,d] y[b,d]: jax.vmap(jnp.dot) (wikipedia/models.py:35-36), sum(a*b,-1) (pinterest/models.py:67-72). */
int esr_rowwise_dot_f32(const float* x, const float* y, int64_t B, int32_t D, float* out, esr_stream_t stream);
/* scores[v, t] = table[v] . queries[t], (V,T) row-major, T <= 64: Glove.score_all (wikipedia/models.py:40-55),
 * the product scan of find_top_k (pinterest/make_recommendations.py:57).  Streams the table once. */
int esr_score_all_f32(const EsrTable* t, const float* queries, int32_t T, float* scores, esr_stream_t stream);
/* Per-query ranking of a (V,T) row-major score matrix (the output of esr_score_all_f32): for every column t the
 * first k entries of the stable sort of the column, out_idx[r*T + t] = row of rank r (out_val likewise, optional).
 * descending == 0, k == V : jnp.argsort(scores, axis=0) of find_knn (wikipedia/train_cooccurence.py:91-97);
 * descending != 0         : jax.lax.top_k (spotify/train_spotify.py:120, pinterest/make_recommendations.py:64),
 * ties in index order in both cases. */
size_t esr_sort_cols_workspace_bytes(int64_t V);
int esr_sort_cols_f32(const float* scores, int64_t V, int32_t T, int32_t descending, int64_t k, int32_t* out_idx,
                      float* out_val, void* ws, size_t ws_bytes, esr_stream_t stream);
/* Fused retrieval (csrc/topk_scan.cu): one streaming pass over N candidate rows scores them against T queries and keeps
 * the running top-k per list in shared memory -- no (N, T) score matrix, no sort of N keys.  Candidate row n is
 * rows_a[(idx_a ? idx_a[n] : n) % mod_a] (mod_a == 0: no modulus; rows_a1 / ver: the double-buffered EsrTable form)
 * optionally concatenated with rows_b[idx_b[n]].
 *   max_over_queries == 0 : T lists, score_t(n) = row_n . queries[t]
 *       dump_knn (wikipedia/train_cooccurence.py:114-126; ties_high_index_first = 1: the tail of a stable ascending argsort),
 *       find_top_k (pinterest/make_recommendations.py:49-65; jax.lax.top_k, ties lower index first);
 *   max_over_queries == 1 : ONE list, score(n) = max_t(row_n . queries[t]) + boost * [idx_a[n] in ctx_a] + boost *
 *       [idx_b[n] in ctx_b] -- the neg_affinity of eval_step (spotify/models.py:78-80 with album_embed[album % 100000]
 *       and raw-id isin; spotify/train_spotify.py:113-131, top 500 of 2.26 M tracks).
 * out_idx[l * k + r] / out_val[l * k + r] = row / score of rank r of list l, best first (out_val may be NULL).
 * k <= 1024, T <= 64, (Da + Db) <= 512; ESR_ENOTSUP when the lists do not fit shared memory. */
typedef struct EsrTopkCfg {
  uint32_t struct_size;
  int32_t T;                /* queries */
  const float* rows_a;      /* [Va][Da] */
  const float* rows_a1;     /* second row buffer of an EsrTable, or NULL */
  const uint8_t* ver;       /* EsrTable.ver, or NULL */
  const float* rows_b;      /* [Vb][Db] or NULL */
  const int32_t* idx_a;     /* [N] or NULL (identity) */
  const int32_t* idx_b;     /* [N], required with rows_b */
  const float* queries;     /* [T][Da + Db] */
  const int32_t* ctx_a;     /* [n_ctx_a] raw ids compared with idx_a (max_over_queries only) */
  const int32_t* ctx_b;     /* [n_ctx_b] raw ids compared with idx_b */
  int64_t N;                /* candidate rows */
  int32_t Da, Db;
  int32_t mod_a;            /* row of A = idx_a[n] % mod_a when > 0 (spotify/models.py:37: album % 100000) */
  int32_t max_over_queries;
  int32_t n_ctx_a, n_ctx_b;
  float boost;              /* 0.1 in the reference */
  int32_t k;
  int32_t ties_high_index_first;
  int32_t reserved;
} EsrTopkCfg;
size_t esr_topk_workspace_bytes(int64_t N, int32_t D, int32_t T, int32_t max_over_queries, int32_t k);
int esr_topk_scan_f32(const EsrTopkCfg* cfg, int32_t* out_idx, float* out_val, void* ws, size_t ws_bytes,
                      esr_stream_t stream);
/* out[k] = uniform integer in [0, hi), k < n, from a counter-based stream keyed by (seed, step, k): the on-device
 * replacement of sample_negative's jax.random.randint(key, [n], 0, N - 1) (spotify/train_spotify.py:139-150; pass
 * hi = N - 1: the reference's upper bound is exclusive).  Bit-exact contract: oracle.index.sample_uniform. */
int esr_sample_uniform_i32(uint64_t seed, uint64_t step, int64_t n, int64_t hi, int32_t* out, esr_stream_t stream);
/* Debug validator: *n_bad (device int32) = number of ids outside [0, V). */
int esr_check_ids_i32(const int32_t* ids, int64_t n, int64_t V, int32_t* n_bad, esr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Index plan of one batch (integer bookkeeping only; depends on the ids, never on the table, so
 * it can be built for batch t+1 while batch t trains).  Slots: s in [0, 2B), key[s] = ids[s]
 * with ids the (2,B) int32 batch of wikipedia/cooccurrence_matrix.py:103-114 laid out flat
 * ([i ; j]: both roles index the same table, wikipedia/models.py:31-34).
 * Bit-exact contract: oracle/index.py (stable sort by row, unique rows, segment offsets).
 * Replaces the gather/scatter index handling XLA performs inside jnp.take and its VJP.
 * ------------------------------------------------------------------------------------------ */
typedef struct EsrPlan {
  uint32_t struct_size;
  int32_t key_bits;      /* number of significant key bits (ceil(log2 V)); 0 => 32 */
  int64_t n_slots;       /* 2B for GloVe */
  const int32_t* keys;   /* [n_slots] input */
  int32_t* sorted_keys;  /* [n_slots] */
  int32_t* perm;         /* [n_slots] sorted position -> slot, stable */
  int32_t* partner;      /* [n_slots] row id of the other role of the slot's pair, or NULL */
  int32_t* useg;         /* [n_slots] sorted position -> index into uniq */
  int32_t* uniq;         /* [n_slots] capacity; first *n_uniq valid */
  int32_t* seg_off;      /* [n_slots+1] capacity; first *n_uniq+1 valid */
  int32_t* n_uniq;       /* device scalar */
  const int32_t* n_valid; /* optional device scalar: number of REAL slots.  The row-sharded path builds its plan over a
                           * fixed-capacity slot array whose padding carries a key larger than every row id, so it sorts
                           * to the end; every consumer (plan, compact plan, prep, row pass, combine) then covers only the
                           * first *n_valid sorted slots.  NULL: all n_slots are real. */
  int32_t sort_impl;      /* ESR_SORT_*: which slot sort builds the plan (results are identical).  AUTO = WIDE. */
  int32_t reserved;
} EsrPlan;

/* Slot sort of the plan (results are identical).  WIDE is libesr's own LSD radix sort (csrc/index_plan.cu: 2048-slot
 * tiles over the whole GPU, two-level look-back, fused head pass): 47 us for 2^19 slots on an idle B200 against 75 us for
 * LIBRARY (cub::DeviceRadixSort + head count / scan / write kernels), and the faster one inside every trainer step
 * measured (single-GPU pipeline 146.8 vs 149 us, owner-routed step at 2 GPUs 265 vs 270 us, in-batch step 0.166 vs 0.200
 * ms; profiles/r2_plan_sort.md).  AUTO = WIDE.  Plans of up to 6144 slots (cub's single-tile kernel wins there) and of
 * more than 2^20 slots are always built by LIBRARY.  ESR_PLAN_SORT=own|cub in the environment overrides field and
 * thresholds (measurement control). */
enum { ESR_SORT_AUTO = 0, ESR_SORT_WIDE = 1, ESR_SORT_LIBRARY = 2 };

size_t esr_plan_workspace_bytes(int64_t n_slots);
int esr_plan_build_i32(const EsrPlan* plan, void* ws, size_t ws_bytes, esr_stream_t stream);
/* ids_out[perm[p]] = useg[p]: rewrites a batch's row ids as indices into its own unique-row list
 * (used by the row-sharded path, where the "table" of a step is the compact buffer of fetched rows). */
int esr_plan_remap_ids_i32(const EsrPlan* plan, int32_t* ids_out, esr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * GloVe step.  Replaces apply_model + update_model (wikipedia/train_cooccurence.py:71-101):
 * Glove.__call__ (wikipedia/models.py:21-38), glove_loss (:76-84), jax.value_and_grad (:86-87),
 * TrainState.apply_gradients (:101).
 * ------------------------------------------------------------------------------------------ */
typedef struct EsrGloveCfg {
  uint32_t struct_size;
  int32_t bias_mode;   /* ESR_BIAS_* ; REFERENCE_BROADCAST reproduces wikipedia/models.py:37 */
  int32_t rows_mode;   /* ESR_ROWS_UPDATE (fused sparse Adagrad) or ESR_ROWS_EMIT_GRADS */
  int32_t impl;        /* ESR_IMPL_* kernel variant of the row pass */
  int64_t B;           /* pairs in this (local) batch */
  int64_t B_global;    /* divisor of the loss (== B on one GPU; sum over ranks when sharded) */
  float lr;
  float eps;           /* Adagrad eps (optax default 1e-7) */
  float x_max;         /* 100.0  wikipedia/train_cooccurence.py:80 */
  float alpha;         /* 0.75   wikipedia/train_cooccurence.py:81 */
  int32_t chunk;       /* sorted slots per work item; 0 => default */
  int32_t reserved;
  const int32_t* emit_map; /* EMIT_GRADS only: gradient of unique row u is written to dE[emit_map[u]] /
                              db[emit_map[u]] (e.g. owner-bucket order of the sharded path); NULL => u */
  /* Optional peer scatter of the EMIT outputs (row-sharded path): when set, emit_map[u] encodes
   * owner << 27 | index and the gradient row is stored to emit_peers_dE[owner] + index*D (resp.
   * emit_peers_db[owner] + index) -- the owners' inboxes, written over NVLink by the row pass itself. */
  void* const* emit_peers_dE;
  void* const* emit_peers_db;
  int32_t n_emit_peers;
  int32_t row_blocks;  /* persistent row-pass grid; 0 => 2 CTAs per SM.  Fewer leaves SM room for the plan stream */
  /* Optional device-side loss log: esr_glove_finish also writes the step's loss to loss_log[*loss_step % loss_log_len] and
   * increments *loss_step -- the slot is chosen on the device, so a replayed CUDA graph logs every step without a copy
   * node on the step's stream. */
  float* loss_log;
  int32_t* loss_step;
  int32_t loss_log_len;
  int32_t reserved2;
  float* loss_host;    /* optional PINNED HOST mirror of loss_log (device-accessible pointer): the step's loss reaches the
                        * host as a 4-byte posted write of the finish kernel, readable after a stream synchronise */
} EsrGloveCfg;

/* Scalars block (device float[ESR_GLOVE_NSCAL]) shared by the three phases.  After
 * esr_glove_prep it holds LOCAL sums; a multi-GPU caller all-reduces [0..2] before
 * esr_glove_rows and [3..4] before esr_glove_finish. */
enum {
  ESR_SC_SUM_BS = 0,  /* sum_r bs_r,  bs_r = b[i_r] + b[j_r] */
  ESR_SC_SUM_BS2 = 1, /* sum_r bs_r^2 */
  ESR_SC_S0 = 2,      /* sum_c w_c */
  ESR_SC_S1 = 3,      /* sum_c w_c res_c */
  ESR_SC_S2 = 4,      /* sum_c w_c res_c^2   (per_pair: sum_c w_c (res_c - bs_c)^2) */
  ESR_SC_LOSS = 5,    /* written by esr_glove_finish */
  ESR_GLOVE_NSCAL = 8
};

size_t esr_glove_workspace_bytes(int64_t B, int32_t D, int32_t chunk);
/* Phase 1: per-slot w, t, bs, buffer versions; sums [0..2]. */
int esr_glove_prep_f32(const EsrTable* t, const EsrPlan* plan, const float* counts,
                       const EsrGloveCfg* cfg, float* scalars, void* ws, size_t ws_bytes,
                       esr_stream_t stream);
/* Phase 2: fused gather -> dot -> loss coefficient -> per-row segment sum -> (Adagrad row write |
 * emit dE[U,D]); sums [3..4].  dE may be NULL in UPDATE mode. */
int esr_glove_rows_f32(EsrTable* t, const EsrPlan* plan, const EsrGloveCfg* cfg, float* scalars,
                       float* dE, void* ws, size_t ws_bytes, esr_stream_t stream);
/* The two kernels of phase 2 separately (esr_glove_rows_f32 == main then combine): the row pass
 * proper, and the fixed-order combine of the partial sums of segments that straddle chunks. */
int esr_glove_rows_main_f32(EsrTable* t, const EsrPlan* plan, const EsrGloveCfg* cfg, float* scalars,
                            float* dE, void* ws, size_t ws_bytes, esr_stream_t stream);
int esr_glove_rows_combine_f32(EsrTable* t, const EsrPlan* plan, const EsrGloveCfg* cfg, float* scalars,
                               float* dE, void* ws, size_t ws_bytes, esr_stream_t stream);
/* Phase 3: bias gradient (-> Adagrad in place | emit db[U]), version flip, loss -> scalars[5]. */
int esr_glove_finish_f32(EsrTable* t, const EsrPlan* plan, const EsrGloveCfg* cfg, float* scalars,
                         float* db, void* ws, size_t ws_bytes, esr_stream_t stream);
/* All three phases back to back (single GPU). */
int esr_glove_step_f32(EsrTable* t, const EsrPlan* plan, const float* counts, const EsrGloveCfg* cfg,
                       float* scalars, float* dE, float* db, void* ws, size_t ws_bytes,
                       esr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Optimizer rules (optax semantics).
 * ------------------------------------------------------------------------------------------ */
/* Sparse Adagrad on the current rows, in place: rows uniq[0..*n_uniq) get gradient g[u,:].
 * North-star rule (BASELINE.json); also the owner-side update of the row-sharded path. */
int esr_sparse_adagrad_f32(EsrTable* t, const int32_t* uniq, const int32_t* n_uniq, int64_t cap,
                           const float* g, const float* gb, float lr, float eps, esr_stream_t stream);
/* dst[uniq[u], :] += / = g[u, :]  (dense gradient pytree of jax.value_and_grad, zero elsewhere). */
int esr_scatter_rows_f32(float* dst, int32_t D, const int32_t* uniq, const int32_t* n_uniq,
                         int64_t cap, const float* g, int32_t accumulate, esr_stream_t stream);
/* optax.adam(lr) over n elements -- wikipedia/train_cooccurence.py:171, :101;
 * pinterest/train_shop_the_look.py:175.  `count` is the step count AFTER this update (>=1).
 * Hyper-parameters are doubles because optax evaluates 1-b and 1-b^count in Python floats before the
 * f32 cast (1.f - 0.999f is off by 5e-5 relative). */
int esr_dense_adam_f32(float* p, const float* g, float* mu, float* nu, int64_t n, double lr, double b1,
                       double b2, double eps, int64_t count, esr_stream_t stream);
/* optax.sgd(lr, momentum) over n elements -- spotify/train_spotify.py:238-241, :110. */
int esr_dense_sgdm_f32(float* p, const float* g, float* trace, int64_t n, float lr, float momentum,
                       esr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Row-sharded table (north star; the reference has no distributed code -- SURVEY.md 0, 8(e)).
 * Ownership is cyclic: owner = row % n_ranks, local row = row / n_ranks (ids are frequency ranks,
 * wikipedia/make_dictionary.py:113-116).  Integer contract: oracle/index.py route_plan.
 * The exchanges themselves (NCCL all-to-all) are issued by the caller on the same stream.
 * ------------------------------------------------------------------------------------------ */
size_t esr_route_workspace_bytes(int64_t cap);
/* uniq[0..*n_uniq) sorted unique global rows of this rank's batch (EsrPlan.uniq).  Outputs:
 * order[k] = index into uniq of the k-th row in owner-bucket order (stable), send_local[k] = its
 * owner-local row id (payload of the id all-to-all), send_counts[r] = rows owned by rank r,
 * inv_order[order[k]] = k (optional, may be NULL: the EsrGloveCfg.emit_map of the peer-memory path). */
int esr_route_plan_i32(const int32_t* uniq, const int32_t* n_uniq, int64_t cap, int32_t n_ranks, int32_t* order,
                       int32_t* send_local, int32_t* send_counts, int32_t* inv_order, void* ws, size_t ws_bytes,
                       esr_stream_t stream);
/* Re-express a plan in unique-row indices (sorted_keys := useg, partner := unique index of the
 * partner row, uniq := 0..U-1) so a step can run on the compact table of fetched rows without a
 * second sort.  perm / useg / seg_off / n_uniq of the original plan stay valid.  scratch: [n_slots]. */
int esr_plan_compact_i32(const EsrPlan* plan, int32_t* sorted_keys, int32_t* partner, int32_t* uniq,
                         int32_t* scratch, esr_stream_t stream);
/* out[k] = src[ids[k]]  (bias lookups: jnp.take on the (V,1) bias table, wikipedia/models.py:32,34). */
int esr_gather_scalar_f32(const float* src, const int32_t* ids, int64_t n, float* out, esr_stream_t stream);
/* scatter == 0: out[k,:] = src[idx[k],:]; scatter != 0: out[idx[k],:] = src[k,:]; k < min(cap, *n_valid)
 * (n_valid may be NULL). */
int esr_permute_rows_f32(const float* src, const int32_t* idx, const int32_t* n_valid, int64_t cap, int32_t D,
                         int32_t scatter, float* out, esr_stream_t stream);
/* Owner side: plan built over the received owner-local ids; g_out[u,:] = sum of g_in[slot,:] over the
 * slots of unique row u in stable sorted order (deterministic); same for the optional bias grads. */
int esr_segment_sum_rows_f32(const EsrPlan* plan, int32_t D, const float* g_in, const float* gb_in, float* g_out,
                             float* gb_out, esr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Ranking losses of the Spotify and shop-the-look trainers (reference shapes).
 * ------------------------------------------------------------------------------------------ */
/* STLModel.__call__ scoring (pinterest/models.py:67-72) + train_step loss and gradient
 * (pinterest/train_shop_the_look.py:99-107): loss = (sum relu(1 + neg - pos) + reg * sum relu(||e|| - 1))
 * / batch_size over the three (B,D) embedding matrices.  row_ws: float[B] scratch. */
int esr_stl_triplet_f32(const float* scene, const float* pos, const float* neg, int64_t B, int32_t D,
                        float regularization, float batch_size, float* d_scene, float* d_pos, float* d_neg,
                        float* pos_score, float* neg_score, float* loss, float* row_ws, esr_stream_t stream);
/* SpotifyModel.__call__ (spotify/models.py:48-91) + train_step loss (spotify/train_spotify.py:91-105)
 * + jax.value_and_grad (:108-109) for a pack of playlists (one CTA each).  Ids are RAW (album ids are
 * taken mod VA for the lookup, isin compares raw ids).  Playlist e: nc context rows, next_off[e+1] -
 * next_off[e] next rows, o negatives; its stacked rows [ctx; next; neg] start at e*(nc+o) + next_off[e]
 * in dXa / dXr / album_rows / artist_rows / l2.  pos_aff / neg_aff / l2 may be NULL. */
int esr_spotify_fwd_bwd_f32(const float* album_table, int64_t VA, const float* artist_table, int32_t F,
                            int32_t n_playlists, int32_t nc, int32_t o, int32_t max_m, const int32_t* album_ctx,
                            const int32_t* artist_ctx, const int32_t* next_album, const int32_t* next_artist,
                            const int32_t* next_off, const int32_t* neg_album, const int32_t* neg_artist,
                            float regularization, float* loss, float* dXa, float* dXr, int32_t* album_rows,
                            int32_t* artist_rows, float* pos_aff, float* neg_aff, float* l2, esr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * In-batch-negative scoring on the tensor cores (tcgen05 / TMEM / TMA; csrc/inbatch_scores.cu).
 * Generalises the explicit-triplet scoring + hinge of the shop-the-look trainer
 * (pinterest/models.py:67-72, pinterest/train_shop_the_look.py:99-104) and the affinity hinge of the
 * Spotify trainer (spotify/train_spotify.py:91-94) to the B x B in-batch-negative form BASELINE.json
 * names for configs[2] and configs[3]: S = Q K^T with bf16 operands and fp32 accumulation; the positive
 * of query i is item i + diag_off, every other item is a negative.
 *   ESR_LOSS_HINGE   : loss = (1/b_norm) sum_i sum_{j != pos(i)} relu(margin + scale*S_ij - scale*S_i,pos(i))
 *   ESR_LOSS_SOFTMAX : loss = (1/b_norm) sum_i [logsumexp_j(scale*S_ij) - scale*S_i,pos(i)]
 * Q [Bq, D], K [Bk, D] fp32 row-major (gathered rows / tower outputs), rounded to bf16 (RNE) inside;
 * dQ, dK are the gradients with respect to those inputs (straight-through), fp32.  The fp32 score matrix
 * is never written to memory; the workspace holds one row chunk of the bf16 dL/dS matrix (hinge: exact
 * {0,1} mask, softmax: probabilities) that the backward contractions read straight back out of L2.  D in {64, 128, 192, 256}.  Contract: oracle/inbatch.py.
 * ------------------------------------------------------------------------------------------ */
typedef struct EsrInbatchCfg {
  uint32_t struct_size;
  int32_t loss_kind; /* ESR_LOSS_* */
  int64_t Bq;        /* queries (rows of Q) */
  int64_t Bk;        /* items (rows of K); == Bq on one GPU, the all-gathered batch when sharded */
  int64_t diag_off;  /* positive of query i is item i + diag_off (rank * Bq when sharded) */
  int32_t D;
  int32_t splits;    /* split-K of the two backward contractions; 0 => auto */
  float margin;      /* hinge margin (1.0 in the reference) */
  float scale;       /* score multiplier (1 / temperature); > 0 */
  float b_norm;      /* loss divisor: the GLOBAL number of queries */
  int32_t chunk_rows; /* query rows per pass; 0 => auto (one pass's bf16 dL/dS block stays resident in L2) */
} EsrInbatchCfg;

size_t esr_inbatch_workspace_bytes(const EsrInbatchCfg* cfg);
int esr_inbatch_fwd_bwd_bf16(const float* Q, const float* K, const EsrInbatchCfg* cfg, float* dQ, float* dK,
                             float* loss, void* ws, size_t ws_bytes, esr_stream_t stream);
/* Test / profiling aid, out10 = {byte offset of G (the LAST row chunk processed), ldG (elements), offset of diag,
 * offset of cnt (float [R][Bq]), R, offset of lse2 (log2 domain), n_chunks, offset of Qh, rows per chunk,
 * Sq * 100 + Sk}. */
int esr_inbatch_ws_layout(const EsrInbatchCfg* cfg, int64_t* out10);

/* ------------------------------------------------------------------------------------------
 * Native host runtime of the single-GPU training loop (csrc/pipeline.cu): the body of train_epoch
 * (wikipedia/train_cooccurence.py:103-112) as ONE call per step -- stage the batch on a copy stream, replay the plan
 * graph on a side stream and the step graph on the main stream, `depth` buffers in flight, events in between.
 * The object owns its three streams; the caller's buffers (staging ids / counts per parity, the loss scalar, a device
 * loss log, optionally a pinned host loss log) are registered once.  The stage graphs are captured from whatever the
 * calling thread launches on the stage's stream between capture_begin and capture_end (which: 0 plan, 1 step).
 * ------------------------------------------------------------------------------------------ */
typedef struct EsrPipeline EsrPipeline;
int esr_pipeline_create(int32_t depth, int32_t main_high_priority, EsrPipeline** out);
int esr_pipeline_streams(const EsrPipeline* p, esr_stream_t* copy, esr_stream_t* side, esr_stream_t* main_);
int esr_pipeline_set_buffers(EsrPipeline* p, void* const* ids_dev, void* const* counts_dev, size_t ids_bytes,
                             size_t counts_bytes, const float* loss_src, float* loss_log, int64_t loss_len,
                             float* loss_host);
int esr_pipeline_capture_begin(EsrPipeline* p, int32_t which);
int esr_pipeline_capture_end(EsrPipeline* p, int32_t which, int32_t k);
/* ids / counts: pinned host or device memory.  flags bit 0: also copy the step's loss to loss_host[step % loss_len]
 * (asynchronously, main stream); bit 1: the inputs were produced on caller_stream (NULL = the legacy default stream),
 * stage them behind it; bit 2: ids and counts are adjacent views of ONE allocation (one copy instead of two); bit 3: the
 * inputs are pinned host memory -- staged by a small kernel reading them over PCIe, which keeps the host->device copy
 * engine free for the parameter uploads of the graph launches. */
int esr_pipeline_submit(EsrPipeline* p, const void* ids, const void* counts, esr_stream_t caller_stream, int32_t flags,
                        int64_t* step_out);
/* Developer aid: timeline of the next n <= 64 steps; trace_read synchronises the device and returns, per traced step, the
 * boundaries {copy begin, copy end, plan begin, plan end, step begin, step end} in microseconds since the arming call. */
int esr_pipeline_trace(EsrPipeline* p, int32_t n);
int esr_pipeline_trace_read(EsrPipeline* p, float* out_us, int32_t* n_out);
/* Blocks the calling host thread until the staging copy of (already submitted) step `step` has completed, i.e. its host
 * batch may be overwritten -- the hand-shake of a loader thread that refills a ring of pinned batches. */
int esr_pipeline_wait_staged(const EsrPipeline* p, int64_t step);
int esr_pipeline_sync(const EsrPipeline* p);
int esr_pipeline_destroy(EsrPipeline* p);

/* ------------------------------------------------------------------------------------------
 * Row-sharded table over NVLink peer memory (device pointers of every rank's buffers, e.g. from a
 * symmetric-memory rendezvous; index = rank).  No host-known sizes, no NCCL on the data path; the
 * caller separates fetch / update phases with device barriers.  See csrc/peer_ops.cu.
 * ------------------------------------------------------------------------------------------ */
/* out[u,:] = shard_{uniq[u] % n}[uniq[u] / n, :], out_bias[u] likewise, u < *n_uniq: the lookup of
 * SURVEY.md 8(e) (index all-to-all + row all-to-all) as one gather over peer pointers. */
int esr_peer_gather_f32(const void* const* peer_rows, const void* const* peer_bias, int32_t n_ranks,
                        const int32_t* uniq, const int32_t* n_uniq, int64_t cap, int32_t D, float* out,
                        float* out_bias, esr_stream_t stream);
/* Owner-routed step: the same lookup for the REMOTE rows only -- a row this rank owns is read where it lives.
 * order / counts: the route plan of the unique rows (esr_route_plan_i32: bucket order, per-owner counts).  Row u lands in
 * out[u,:] / out_bias[u] (the fetch region behind the shard; see esr_plan_compact_owner_i32). */
int esr_peer_gather_remote_f32(const void* const* peer_rows, const void* const* peer_bias, int32_t n_ranks, int32_t me,
                               const int32_t* uniq, const int32_t* order, const int32_t* counts, int64_t cap, int32_t D,
                               float* out, float* out_bias, int32_t parts /* 1 rows | 2 biases */, esr_stream_t stream);
/* The plan re-expressed in ADDRESSES of the unified table [shard rows ; fetch region]: key of unique row u =
 * uniq[u] / n_ranks when this rank owns it, base + u otherwise (base = rows of the shard allocation); partner likewise.
 * scratch: n_slots ints. */
int esr_plan_compact_owner_i32(const EsrPlan* plan, int32_t n_ranks, int32_t me, int32_t base, int32_t* sorted_keys,
                               int32_t* partner, int32_t* scratch, esr_stream_t stream);
/* Owner `me`: from every source rank's published esr_route_plan_i32 outputs (send_counts[n],
 * send_local[]) copy the owner-local ids destined to me into recv_ids (source-major), write
 * src_meta[s] = {offset in recv_ids, count, displacement in source s's bucket order}, src_meta[3n] = total
 * (src_meta[3n+1] is the merge's owner-entry counter; src_meta holds 3n + 4 ints),
 * and slot_map[s*map_stride + x] = position of row x in source s's list (slot_map is all -1 on entry
 * and is restored to -1 by esr_peer_merge_adagrad_f32). */
/* Source side, after every rank published its route plan: emit_map[u] = owner << 27 | (offset of my
 * bucket inside owner's inbox + position of row u in that bucket), for the row pass's peer scatter.
 * err[0] is set to 1 if an index would exceed inbox_cap. */
int esr_peer_emit_plan_i32(const void* const* peer_counts, int32_t n_ranks, int32_t me, const int32_t* uniq,
                           const int32_t* n_uniq, int64_t cap, const int32_t* inv_order, int64_t inbox_cap,
                           int32_t* emit_map, int32_t* err, esr_stream_t stream);
int esr_peer_pull_ids_i32(const void* const* peer_counts, const void* const* peer_send_local, int32_t n_ranks,
                          int32_t me, int64_t recv_cap, int32_t* recv_ids, int32_t* src_meta, int32_t* slot_map,
                          int64_t map_stride, esr_stream_t stream);
/* Owner side: merge the gradients the sources scattered into my inbox (inbox_dE[recv_cap, D],
 * inbox_db[recv_cap], source-major like recv_ids; summed in source order, deterministic) and apply
 * optax.adagrad to the shard in place.  All loads are local. */
int esr_peer_merge_adagrad_f32(EsrTable* shard, const float* inbox_dE, const float* inbox_db,
                               int32_t n_ranks, const int32_t* recv_ids, const int32_t* src_meta, int32_t* slot_map,
                               int64_t map_stride, int32_t* desc /* scratch [recv_cap * (n_ranks + 2)], 8-byte aligned, recv_cap * n_ranks even */, int64_t recv_cap,
                               float lr, float eps, esr_stream_t stream);

/* All-reduce (sum) of count <= 6 floats across the ranks -- count == 0: a device barrier -- over symmetric peer memory,
 * as one single-warp kernel (no NCCL, CUDA-graph capturable): the batch-sum reductions SURVEY.md 8(e)(3) and the phase
 * barriers of the sharded step.  peer_sync[r] = rank r's sync block (esr_peer_sync_bytes() bytes, zeroed once before the
 * first call, same rendezvous as the other peer buffers); seq_counter = device uint32, zeroed once, advanced by every
 * call (every rank must make the same calls in the same order).  Sums are formed in rank order: bit-identical on every
 * rank.  in == out is allowed.  A peer that never arrives traps after 20 s instead of hanging the device. */
size_t esr_peer_sync_bytes(void);
int esr_peer_allreduce_f32(void* const* peer_sync, int32_t n_ranks, int32_t me, const float* in, float* out,
                           int32_t count, uint32_t* seq_counter, esr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Native decoders of the reference's record formats (HOST pointers).  Re-entrant;
 * allocation-free except for CooccurrenceRow lines longer than 16 KB decoded (heap line buffer, any
 * --max_row_size); a negative return is an ESR_E* code.
 * ------------------------------------------------------------------------------------------ */
/* Text of a *.cooccur.pb.b64 part after bz2 decompression: one base64 line per CooccurrenceRow
 * {1: index, 2: packed other_index, 3: packed float count}.  Writes the (i, j, count) triples in the order
 * CooccurrenceGenerator.get_item yields them (wikipedia/cooccurrence_matrix.py:62-83) and returns their number;
 * stops before a row that would exceed cap or at an incomplete last line; *consumed = bytes fully decoded. */
int64_t esr_decode_cooccur_b64(const char* text, size_t n_bytes, int32_t* out_i, int32_t* out_j, float* out_count,
                               int64_t cap, int64_t* n_rows, size_t* consumed);
/* TFRecord stream of tf.train.Example with int64_list features (spotify/input_pipeline.py:23-37,
 * spotify/make_training.py:102-112).  For key k the values of record r are vals[k][offs[k][r] .. offs[k][r+1]);
 * offs[k] has max_records + 1 entries.  Returns the number of records decoded; CRCs are not verified. */
int64_t esr_decode_tfrecord_int64(const uint8_t* data, size_t n_bytes, int32_t n_keys, const char* const* keys,
                                  int64_t* const* vals, const int64_t* val_cap, int64_t* const* offs,
                                  int64_t max_records, size_t* consumed);

/* Window shuffle of the host input pipeline (HOST pointers; get_shuffled_items, wikipedia/cooccurrence_matrix.py:80-87).
 * pi = a pseudo-random permutation of [0, n) fixed by `seed` (keyed multiply / xor-shift bijection, cycle-walked; never
 * stored);
 * dst_*[a] = src_*[pi(k0 + a)] for a in [0, m): elements [k0, k0 + m) of the shuffled window, gathered by `threads` host
 * threads straight into their destination (a pinned batch block).  k0 + m <= n. */
int esr_host_shuffle_gather(const int32_t* i, const int32_t* j, const float* c, int64_t n, uint64_t seed, int64_t k0,
                            int64_t m, int32_t* dst_i, int32_t* dst_j, float* dst_c, int32_t threads);

/* OWNER-COMPUTES pair routing (csrc/peer_ops.cu): a pair (i, j, x) is processed by the rank owning row i, so only the
 * unique partner rows j cross NVLink (4.3x fewer bytes on the bench stream at 8 ranks than keeping the pairs where they
 * arrived).  esr_peer_route_pairs_i32 (source side; depends on the ids only) partitions the (2,B) batch STABLY by
 * owner(i) = i % n straight into the owners' pair inboxes: peer_pair_rec[o] -> int4 [n][B] (region of source `me`: one
 * 16-byte record {i, j, count bits, 0} per pair -- a single NVLink store each), peer_pair_counts[o] -> int32 [n] (entry
 * `me` = pairs I send to o); my_counts[o] = the same numbers locally.  After a device barrier,
 * esr_peer_collect_pairs_i32 (owner side) concatenates the regions in source order into keys[2 * B_cap] ([i ; j] halves of
 * capacity B_cap, padding = pad_key, which must exceed every row id so that it sorts to the end), counts[B_cap] and
 * *n_valid = 2 m for EsrPlan.n_valid; err |= 2 if m > B_cap (the excess pairs are dropped: the caller must treat it as
 * fatal).  Bit-exact contract: oracle/index.py route_pairs / collect_pairs. */
size_t esr_peer_route_pairs_workspace_bytes(int64_t B);
int esr_peer_route_pairs_i32(const int32_t* ids, const float* counts, int64_t B, int32_t n_ranks, int32_t me,
                             void* const* peer_pair_rec, void* const* peer_pair_counts, int32_t* my_counts, void* ws,
                             size_t ws_bytes, esr_stream_t stream);
int esr_peer_collect_pairs_i32(const void* in_rec, const int32_t* in_counts, int32_t n_ranks, int64_t B, int64_t B_cap,
                               int32_t pad_key, int32_t* keys, float* counts, int32_t* n_valid, int32_t* err,
                               esr_stream_t stream);

/* The two halves of esr_peer_merge_adagrad_f32: resolve depends on the ids only (side stream, overlaps the row
 * pass); apply needs the gradients (after the device barrier).  desc: [recv_cap * (n_ranks + 2)] ints, 8-byte aligned --
 * per received entry the inbox row of every source naming its row, then one 8-byte record per OWNED row
 * {entry | several-sources flag << 31, owner-local row} that the merge streams. */
int esr_peer_resolve_i32(int32_t n_ranks, const int32_t* recv_ids, int32_t* src_meta, const int32_t* slot_map,
                         int64_t map_stride, int32_t* desc, int64_t recv_cap, esr_stream_t stream);
int esr_peer_apply_adagrad_f32(EsrTable* shard, const float* inbox_dE, const float* inbox_db, int32_t n_ranks,
                               const int32_t* recv_ids, const int32_t* src_meta, int32_t* slot_map, int64_t map_stride,
                               const int32_t* desc, int64_t recv_cap, float lr, float eps, esr_stream_t stream);
/* The same in two independent halves (parts: 1 = embedding rows, 2 = biases + slot_map restore; 3 = both): they touch
 * disjoint state, so the rows can start as soon as every rank's gradient ROWS have landed while the bias gradients are
 * still being formed. */
int esr_peer_apply_parts_f32(EsrTable* shard, const float* inbox_dE, const float* inbox_db, int32_t n_ranks,
                             const int32_t* recv_ids, const int32_t* src_meta, int32_t* slot_map, int64_t map_stride,
                             const int32_t* desc, int64_t recv_cap, float lr, float eps, int32_t parts, esr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ESR_H_ */
