// Fused GloVe step: gather -> dot -> loss coefficient -> per-row segment sum -> sparse Adagrad.
//
// Replaces apply_model + update_model of the reference (wikipedia/train_cooccurence.py:71-101):
//   Glove.__call__            wikipedia/models.py:21-38   (shared table for both roles, row-wise dot,
//                                                         the (B,)+(B,1)+(B,1) -> (B,B) broadcast)
//   glove_loss                wikipedia/train_cooccurence.py:76-84
//   jax.value_and_grad        wikipedia/train_cooccurence.py:86-87 (VJP of jnp.take = scatter-add)
//   TrainState.apply_gradients wikipedia/train_cooccurence.py:101
// Math: SURVEY.md App. A.1 (reference_broadcast) / A.2 (per_pair); CPU restatement: oracle/glove.py.
//
// Data movement (D = 128, R = 512 B per row):
//   The 2B slots of a batch are processed in SORTED row order, in chunks of `chunk` slots per warp.
//   A lane owns one float4 column of the row (NK float4 when D > 128).  For every slot the warp
//   reads the PARTNER row (random; an L2 hit whenever the partner is itself touched in this step),
//   for every segment head the SELF row, and for every segment that closes inside the chunk the
//   Adagrad accumulator row; it writes the NEW row into the table's other buffer (EsrTable
//   versioning) and the accumulator in place.  So each touched row is read once and written once
//   from HBM, with no gradient buffer in between: 4*U*R algorithmic bytes per step.
//   All S loads of a sub-batch are issued before the first use (S*3 independent 16-byte loads per
//   lane in flight) to cover HBM latency.
//   Segments that straddle a chunk boundary (Zipf head rows) leave per-chunk partial sums that
//   k_glove_combine adds in a fixed order, so the result is bit-reproducible run to run.
#include <algorithm>
#include <climits>
#include <cstddef>
#include <cstdint>

#include "esr_common.cuh"

namespace esr {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kPrepItems = 4;
constexpr int kCombineThreads = 512;
constexpr int kCombineWarps = kCombineThreads / 32;
constexpr int32_t kRowMask = 0x3fffffff;
constexpr int32_t kRoleBit = 0x40000000;  // slot is the i role: its pair counts towards the batch sums
constexpr int32_t kNoKey = 0x40000000;    // never a valid skv (bit 30 of a row id is 0)

struct __align__(16) SlotRec {
  int32_t code;  // partner row | role << 30 | partner version << 31
  float w;       // min(1, x/x_max)^alpha          wikipedia/train_cooccurence.py:79-81
  float t;       // log10(1 + x)                   wikipedia/train_cooccurence.py:82
  float bs;      // b[i] + b[j] of the slot's pair wikipedia/models.py:33-34,37
};

struct GloveWs {
  int32_t* skv;     // [n] sorted row | version << 31
  SlotRec* rec;     // [n]
  float* bsum;      // [n] per unique row: sum of bs (broadcast) or of g (per_pair) over its slots
  float* part;      // [nchunks][2][D] partial gradient sums of straddling segments
  float* parts;     // [nchunks][2]    partial scalar sums
  float* prep_blk;  // [prep_blocks][3]
  float* rows_blk;  // [row_blocks][2]
  int32_t* wl_count;  // [4 + heavy_cap]: light / heavy list lengths, work counter, then one ticket per heavy segment
  int32_t* wl_light;  // [nchunks]
  int32_t* wl_heavy;  // [nchunks]
  float* part2;       // [heavy_cap][kHeavySplit][D] second-level partial sums of the heavy segments
  float* parts2;      // [heavy_cap][kHeavySplit]
  int64_t heavy_cap;
  int64_t nchunks;
  int32_t prep_blocks;
  int32_t row_blocks;
  int32_t chunk;
};

// lanes that own one row in the group variant: 4 float4 per lane, rounded up to a power of two
int group_lanes(int D4) {
  int g = 1;
  while (g * 4 < D4 && g < 32) g <<= 1;
  return g;
}

constexpr int kMinAutoChunk = 16;
constexpr int kHeavyParts = 32;   // straddling segments with more partials than this are combined by several blocks
constexpr int kHeavySplit = 8;    // blocks per heavy segment

// chunk == 0: pick the chunk length that wastes the least of the last wave.  All resident groups
// advance in lock step through `rounds` chunks, so the pass costs rounds * (chunk + start-up).
int auto_chunk(int64_t n, int D4) {
  const int64_t resident = (int64_t)sm_count() * 2 * kWarps * (32 / group_lanes(D4));
  int best = 32;
  int64_t best_cost = INT64_MAX;
  for (int c = 32; c >= kMinAutoChunk; c -= 4) {
    const int64_t rounds = ceil_div(ceil_div(n, c), resident);
    const int64_t cost = rounds * (c + 2);
    if (cost < best_cost) {
      best_cost = cost;
      best = c;
    }
  }
  return best;
}

int norm_chunk(int32_t chunk, int64_t n, int D4) {
  if (chunk <= 0) return auto_chunk(n, D4);
  if (chunk > 32) return 32;
  return (chunk + 3) / 4 * 4;
}

size_t carve_ws(void* base, int64_t B, int32_t D, int32_t chunk_in, GloveWs* w) {
  const int64_t n = 2 * (B > 0 ? B : 1);
  const int chunk = norm_chunk(chunk_in, n, D / 4);
  Carver c(base);
  GloveWs t;
  t.chunk = chunk;
  t.nchunks = ceil_div(n, chunk);
  // capacity: with chunk == 0 the chunk length is chosen per device, so size for the smallest one
  const int64_t cap_chunks = chunk_in <= 0 ? ceil_div(n, kMinAutoChunk) : t.nchunks;
  t.prep_blocks = (int32_t)ceil_div(n, kThreads * kPrepItems);
  t.row_blocks = (int32_t)ceil_div(t.nchunks, kWarps);
  const int64_t cap_row_blocks = ceil_div(cap_chunks, kWarps);
  t.skv = c.take<int32_t>(n);
  t.rec = c.take<SlotRec>(n);
  t.bsum = c.take<float>(n);
  t.part = c.take<float>(cap_chunks * 2 * D);
  t.parts = c.take<float>(cap_chunks * 2);
  t.prep_blk = c.take<float>((size_t)t.prep_blocks * 3);
  t.rows_blk = c.take<float>((size_t)cap_row_blocks * 2);
  t.heavy_cap = n / ((int64_t)kHeavyParts * kMinAutoChunk) + 2;  // a heavy segment holds > kHeavyParts * chunk slots
  t.wl_count = c.take<int32_t>(4 + t.heavy_cap);
  t.wl_light = c.take<int32_t>(cap_chunks);
  t.wl_heavy = c.take<int32_t>(cap_chunks);
  t.part2 = c.take<float>((size_t)t.heavy_cap * kHeavySplit * D);
  t.parts2 = c.take<float>((size_t)t.heavy_cap * kHeavySplit);
  if (w) *w = t;
  return c.off;
}

// Fixed-order reduction of NV-wide block partials into scalars[off .. off+NV) (double accumulate).
template <int NV>
__device__ void reduce_partials(const float* __restrict__ blk, int nblk, float* __restrict__ out) {
  __shared__ double sh[kCombineThreads * NV];
  double acc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) acc[k] = 0.0;
  for (int b = threadIdx.x; b < nblk; b += blockDim.x)
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] += (double)__ldcg(blk + b * NV + k);
#pragma unroll
  for (int k = 0; k < NV; ++k) sh[threadIdx.x * NV + k] = acc[k];
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o)
#pragma unroll
      for (int k = 0; k < NV; ++k) sh[threadIdx.x * NV + k] += sh[(threadIdx.x + o) * NV + k];
    __syncthreads();
  }
  if (threadIdx.x == 0)
#pragma unroll
    for (int k = 0; k < NV; ++k) out[k] = (float)sh[k];
}

// ---------------------------------------------------------------------------------------------
// Phase 1: slot records + batch sums that do not depend on the dots.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_glove_prep(const int32_t* __restrict__ sk, const int32_t* __restrict__ perm,
                                                         const int32_t* __restrict__ partner,
                                                         const float* __restrict__ counts,
                                                         const float* __restrict__ bias, const uint8_t* __restrict__ ver,
                                                         int64_t n_cap, const int32_t* __restrict__ n_valid, int64_t B,
                                                         float x_max, float alpha,
                                                         int32_t* __restrict__ skv, SlotRec* __restrict__ rec,
                                                         float* __restrict__ blk, float* __restrict__ scalars,
                                                         int32_t* __restrict__ wl_count, int32_t wl_words) {
  __shared__ float red[32 * 3];
  __shared__ bool last_block;
  const int64_t n = n_valid ? min(n_cap, (int64_t)__ldg(n_valid)) : n_cap;
  // the row pass's work lists / work counter / heavy-segment tickets start every step at zero: cleared here rather than by a
  // memset node in front of the row pass (one launch less on the step's critical path)
  for (int32_t k = blockIdx.x * kThreads + threadIdx.x; k < wl_words; k += gridDim.x * kThreads) wl_count[k] = 0;
  // kPrepItems slots per thread, every load of a stage issued before the first use: the kernel is a chain of
  // dependent gathers (perm -> counts, row -> bias / version), so its time is (latency x stages), not bytes.
  const int64_t p0 = blockIdx.x * (int64_t)(kThreads * kPrepItems) + threadIdx.x;
  float v[3] = {0.f, 0.f, 0.f};
  int32_t row[kPrepItems], q[kPrepItems], s[kPrepItems];
#pragma unroll
  for (int k = 0; k < kPrepItems; ++k) {
    const int64_t p = p0 + (int64_t)k * kThreads;
    row[k] = q[k] = s[k] = 0;
    if (p < n) {
      row[k] = sk[p];
      q[k] = partner[p];
      s[k] = perm[p];
    }
  }
  float x[kPrepItems], br[kPrepItems], bq[kPrepItems];
  int32_t vr[kPrepItems], vq[kPrepItems];
#pragma unroll
  for (int k = 0; k < kPrepItems; ++k) {
    const int64_t p = p0 + (int64_t)k * kThreads;
    x[k] = br[k] = bq[k] = 0.f;
    vr[k] = vq[k] = 0;
    if (p < n) {
      x[k] = counts[s[k] < B ? s[k] : s[k] - B];
      br[k] = bias[row[k]];
      bq[k] = bias[q[k]];
      if (ver) {
        vr[k] = ver[row[k]];
        vq[k] = ver[q[k]];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kPrepItems; ++k) {
    const int64_t p = p0 + (int64_t)k * kThreads;
    if (p < n) {
      const bool role_i = s[k] < B;
      SlotRec r;
      r.w = powf(fminf(1.f, x[k] / x_max), alpha);
      r.t = log10f(1.f + x[k]);
      // b[i] + b[j]: fp add commutes, so both roles of a pair hold the same bits
      r.bs = br[k] + bq[k];
      r.code = (int32_t)((uint32_t)q[k] | (role_i ? (uint32_t)kRoleBit : 0u) | ((uint32_t)vq[k] << 31));
      rec[p] = r;
      skv[p] = (int32_t)((uint32_t)row[k] | ((uint32_t)vr[k] << 31));
      if (role_i) {
        v[0] += r.bs;
        v[1] = fmaf(r.bs, r.bs, v[1]);
        v[2] += r.w;
      }
    }
  }
  block_sum<3>(v, red);
  if (threadIdx.x == 0) {
    blk[blockIdx.x * 3 + 0] = v[0];
    blk[blockIdx.x * 3 + 1] = v[1];
    blk[blockIdx.x * 3 + 2] = v[2];
    // the last block to arrive reduces every block's partials in index order (deterministic); the ticket
    // lives in the unused tail of the scalars block, zeroed by the caller's memset
    __threadfence();
    last_block = atomicAdd(reinterpret_cast<unsigned*>(scalars + ESR_GLOVE_NSCAL - 1), 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last_block) {
    __threadfence();
    reduce_partials<3>(blk, gridDim.x, scalars + ESR_SC_SUM_BS);
  }
}

// ---------------------------------------------------------------------------------------------
// Phase 2: the row pass.
// ---------------------------------------------------------------------------------------------
template <int NK>
struct Row {
  float4 v[NK];
};

template <int NK>
__device__ __forceinline__ void row_zero(Row<NK>& r) {
#pragma unroll
  for (int k = 0; k < NK; ++k) r.v[k] = f4_zero();
}

// lane owns float4 columns lane, lane+32, ...; columns >= D4 read as zero and are never written
template <int NK, bool STREAM>
__device__ __forceinline__ void row_load(Row<NK>& r, const float4* __restrict__ p, int lane, int D4) {
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const int c = k * 32 + lane;
    if (c < D4) r.v[k] = STREAM ? ld_stream(p + c) : ld_keep(p + c);
    else r.v[k] = f4_zero();
  }
}

template <int NK, bool STREAM>
__device__ __forceinline__ void row_store(const Row<NK>& r, float4* __restrict__ p, int lane, int D4) {
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const int c = k * 32 + lane;
    if (c < D4) {
      if (STREAM) st_stream(p + c, r.v[k]);
      else p[c] = r.v[k];
    }
  }
}

template <int NK>
__device__ __forceinline__ void row_add(Row<NK>& acc, const Row<NK>& x) {
#pragma unroll
  for (int k = 0; k < NK; ++k) f4_add(acc.v[k], x.v[k]);
}

// EMIT destinations: with emit_peers the code emit_map[u] = owner << kEmitShift | index selects one of
// up to ESR_MAX_PEERS base pointers (the owners' gradient inboxes, written over NVLink).
constexpr int kEmitShift = 27;
struct EmitPeers {
  float* dE[ESR_MAX_PEERS];
  float* db[ESR_MAX_PEERS];
  int32_t on;
};

struct RowsArgs {
  const float* rows[2];
  float* wrows[2];
  float* acc;
  const int32_t* skv;
  const SlotRec* rec;
  const int32_t* useg;
  const int32_t* seg_off;
  const float* scalars;  // [0..2] already global
  float* bsum;
  float* part;
  float* parts;
  float* rows_blk;
  int32_t* wl_count;  // [0] light, [1] heavy work-list lengths
  int32_t* work_counter;  // dynamic work distribution of the persistent row pass
  int32_t* wl_light;  // head chunks of straddling segments with <= kHeavyParts partials
  int32_t* wl_heavy;
  int32_t* tickets;   // [heavy_cap] arrival counters of the heavy segments' blocks (zeroed with wl_count)
  float* part2;
  float* parts2;
  float* dE;
  const int32_t* emit_map;  // EMIT: gradient of unique row u goes to dE[emit_map[u]] (NULL: dE[u])
  EmitPeers peers;
  int64_t n;
  int64_t nchunks;
  const int32_t* n_valid;  // optional (EsrPlan.n_valid): only the first *n_valid sorted slots are real
  int32_t interleave;      // persistent async row pass: alternate work items from both ends of the sorted stream
  int32_t D4;
  int32_t chunk;
  int32_t per_pair;
  int32_t emit;
  float c2B;      // -2 / B_global
  float inv_B;    // 1 / B_global
  float lr, eps;
};


// Slots / chunks the pass really covers: the host-side capacity clamped by the device-side count of the plan
// (row-sharded path: the pairs a rank receives are only counted on the device).
__device__ __forceinline__ int64_t eff_n(const RowsArgs& a) {
  return a.n_valid ? min(a.n, (int64_t)__ldg(a.n_valid)) : a.n;
}
__device__ __forceinline__ int64_t eff_chunks(const RowsArgs& a, int64_t n) { return (n + a.chunk - 1) / a.chunk; }

// Close a finished segment: Adagrad row write (UPDATE) or gradient emit (EMIT).
template <int NK>
__device__ __forceinline__ void close_segment(const RowsArgs& a, int32_t key, int64_t u, const Row<NK>& self,
                                              const Row<NK>& accrow, const Row<NK>& grad, float bacc, int lane) {
  if (a.emit) {
    int64_t e = a.emit_map ? a.emit_map[u] : u;
    float* base = a.dE;
    if (a.peers.on) {
      base = a.peers.dE[e >> kEmitShift];
      e &= (1 << kEmitShift) - 1;
    }
    row_store<NK, true>(grad, reinterpret_cast<float4*>(base) + e * a.D4, lane, a.D4);
  } else {
    const int64_t row = key & kRowMask;
    const int v = (key >> 31) & 1;
    Row<NK> p = self, ac = accrow;
#pragma unroll
    for (int k = 0; k < NK; ++k) adagrad4(p.v[k], ac.v[k], grad.v[k], a.lr, a.eps);
    row_store<NK, true>(p, reinterpret_cast<float4*>(a.wrows[1 - v]) + row * a.D4, lane, a.D4);
    row_store<NK, true>(ac, reinterpret_cast<float4*>(a.acc) + row * a.D4, lane, a.D4);
  }
  if (lane == 0) a.bsum[u] = bacc;
}

// Per-chunk bookkeeping shared by both kernel variants.
struct ChunkMeta {
  int cnt;
  int32_t kl;          // lane's own skv (kNoKey beyond cnt)
  unsigned head_mask;  // slot is the first of its segment (true head)
  unsigned end_mask;   // slot is the last of its segment
  SlotRec r;           // lane's own record
  int64_t u_first;     // unique-row index of slot 0
};

__device__ __forceinline__ ChunkMeta load_chunk(const RowsArgs& a, int64_t c, int lane) {
  ChunkMeta m;
  const int64_t n = eff_n(a);
  const int64_t p0 = c * a.chunk;
  m.cnt = (int)min((int64_t)a.chunk, n - p0);
  m.kl = lane < m.cnt ? a.skv[p0 + lane] : kNoKey;
  const int32_t kprev0 = p0 > 0 ? a.skv[p0 - 1] : kNoKey;
  const int32_t knextN = p0 + m.cnt < n ? a.skv[p0 + m.cnt] : kNoKey;
  int32_t kp = __shfl_up_sync(FULL, m.kl, 1);
  if (lane == 0) kp = kprev0;
  int32_t kn = __shfl_down_sync(FULL, m.kl, 1);
  if (lane == m.cnt - 1) kn = knextN;
  m.head_mask = __ballot_sync(FULL, lane < m.cnt && m.kl != kp);
  m.end_mask = __ballot_sync(FULL, lane < m.cnt && m.kl != kn);
  m.r.code = 0; m.r.w = 0.f; m.r.t = 0.f; m.r.bs = 0.f;
  if (lane < m.cnt) m.r = a.rec[p0 + lane];
  m.u_first = a.useg[p0];
  return m;
}

template <int NK>
struct SegState {
  Row<NK> cur, grad;
  float bacc;
  bool started_here;
  int32_t cur_key;
};

// One slot: dot with the segment's self row, loss coefficient, gradient accumulate, and -- at
// the end of a segment or of the chunk -- the row update or the partial-sum hand-off.
template <int NK>
__device__ __forceinline__ void consume_slot(const RowsArgs& a, const ChunkMeta& m, int64_t c, int s, SegState<NK>& st,
                                             const Row<NK>& P, const Row<NK>& Sf, const Row<NK>& A, int32_t code,
                                             int32_t key, float w, float t, float bs, float mbs, float (&sums)[2],
                                             int lane) {
  const bool is_head = (m.head_mask >> s) & 1u;
  if (s == 0 || is_head) {
    st.cur = Sf;
    st.cur_key = key;
    row_zero(st.grad);
    st.bacc = 0.f;
    st.started_here = is_head;
  }
  float d = 0.f;
#pragma unroll
  for (int k = 0; k < NK; ++k) d += f4_dot(st.cur.v[k], P.v[k]);
  const float dot = warp_sum(d);
  const float res = t - dot;
  const float rr = a.per_pair ? res - bs : res;
  const float g = a.c2B * w * (a.per_pair ? rr : res - mbs);
  if (code & kRoleBit) {
    sums[0] = fmaf(w, res, sums[0]);
    sums[1] = fmaf(w * rr, rr, sums[1]);
  }
#pragma unroll
  for (int k = 0; k < NK; ++k) f4_fma(st.grad.v[k], g, P.v[k]);
  st.bacc += a.per_pair ? g : bs;
  const bool is_end = (m.end_mask >> s) & 1u;
  if (is_end) {
    if (st.started_here) {
      const int64_t u = m.u_first + __popc(m.head_mask & ((2u << s) - 1u) & ~1u);
      close_segment<NK>(a, st.cur_key, u, st.cur, A, st.grad, st.bacc, lane);
    } else {  // leading partial: the segment started in an earlier chunk
      row_store<NK, false>(st.grad, reinterpret_cast<float4*>(a.part) + (c * 2 + 0) * a.D4, lane, a.D4);
      if (lane == 0) a.parts[c * 2 + 0] = st.bacc;
    }
  } else if (s == m.cnt - 1) {  // the chunk ends inside a segment
    const int slot = st.started_here ? 1 : 0;
    row_store<NK, false>(st.grad, reinterpret_cast<float4*>(a.part) + (c * 2 + slot) * a.D4, lane, a.D4);
    if (lane == 0) {
      a.parts[c * 2 + slot] = st.bacc;
      if (slot == 1) {  // this chunk holds the head of a straddling segment: queue its combine
        const int64_t u = m.u_first + __popc(m.head_mask & ~1u);
        const int64_t np = (a.seg_off[u + 1] - 1) / a.chunk - c + 1;
        const int heavy = np > kHeavyParts;
        const int e = atomicAdd(a.wl_count + heavy, 1);
        (heavy ? a.wl_heavy : a.wl_light)[e] = (int32_t)c;
      }
    }
  }
}

__device__ __forceinline__ void store_block_sums(const RowsArgs& a, float (&sums)[2], float* red, int lane) {
  // per-block S1 / S2 partials (lane 0 of each warp carries the warp's lane-uniform value)
  float v[2] = {lane == 0 ? sums[0] : 0.f, lane == 0 ? sums[1] : 0.f};
  block_sum<2>(v, red);
  if (threadIdx.x == 0) {
    a.rows_blk[blockIdx.x * 2 + 0] = v[0];
    a.rows_blk[blockIdx.x * 2 + 1] = v[1];
  }
}

// LDG variant: one chunk per warp, S slots' rows held in registers between issue and use.
template <int NK, int S, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_glove_rows(const RowsArgs a) {
  __shared__ float red[32 * 2];
  const int lane = threadIdx.x & 31;
  const int64_t c = blockIdx.x * (int64_t)kWarps + (threadIdx.x >> 5);
  float sums[2] = {0.f, 0.f};  // S1, S2 contributions (lane-uniform)

  if (c < eff_chunks(a, eff_n(a))) {
    const ChunkMeta m = load_chunk(a, c, lane);
    const float mbs = a.scalars[ESR_SC_SUM_BS] * a.inv_B;
    SegState<NK> st;
    row_zero(st.cur);
    row_zero(st.grad);
    st.bacc = 0.f;
    st.started_here = false;
    st.cur_key = kNoKey;

    for (int sb = 0; sb < m.cnt; sb += S) {
      Row<NK> P[S], Sf[S], A[S];
      int32_t code[S], key[S];
      float w[S], t[S], bs[S];
      // ---- issue every load of the sub-batch before the first use ----
#pragma unroll
      for (int j = 0; j < S; ++j) {
        const int s = sb + j;
        const int src = s < 32 ? s : 31;
        code[j] = __shfl_sync(FULL, m.r.code, src);
        w[j] = __shfl_sync(FULL, m.r.w, src);
        t[j] = __shfl_sync(FULL, m.r.t, src);
        bs[j] = __shfl_sync(FULL, m.r.bs, src);
        key[j] = __shfl_sync(FULL, m.kl, src);
        if (s < m.cnt) {
          const int64_t q = code[j] & kRowMask;
          const int qv = (code[j] >> 31) & 1;
          row_load<NK, false>(P[j], reinterpret_cast<const float4*>(a.rows[qv]) + q * a.D4, lane, a.D4);
          const int64_t row = key[j] & kRowMask;
          if ((s == 0) || ((m.head_mask >> s) & 1u)) {
            const int v = (key[j] >> 31) & 1;
            row_load<NK, false>(Sf[j], reinterpret_cast<const float4*>(a.rows[v]) + row * a.D4, lane, a.D4);
          }
          // The accumulator is only used when the segment also started in this chunk; a segment that
          // closes here but started earlier is rare (one per straddling row).
          if (!a.emit && ((m.end_mask >> s) & 1u))
            row_load<NK, true>(A[j], reinterpret_cast<const float4*>(a.acc) + row * a.D4, lane, a.D4);
        }
      }
      // ---- consume in slot order ----
#pragma unroll
      for (int j = 0; j < S; ++j) {
        const int s = sb + j;
        if (s < m.cnt)
          consume_slot<NK>(a, m, c, s, st, P[j], Sf[j], A[j], code[j], key[j], w[j], t[j], bs[j], mbs, sums, lane);
      }
    }
  }
  store_block_sums(a, sums, red, lane);
}

// ---------------------------------------------------------------------------------------------
// Phase 2, group variant (default): a row is owned by a GROUP of G lanes (NV float4 per lane,
// D <= 16*G*NV... i.e. D4 <= G*NV), and the 32/G groups of a warp walk 32/G consecutive chunks in
// lock step, so every warp instruction advances 32/G slots.  The LDG-per-warp variant above spends
// ~230 warp instructions per slot and is issue-bound (ncu: 58-78 % issue-active at 4.7-5.1 TB/s);
// here the per-slot bookkeeping (keys, record, addresses, flags) is shared by 32/G slots.
// Lane gl of a group owns float4 columns gl, gl+G, gl+2G, ... so each load instruction of a group
// covers G*16 contiguous bytes.
// ---------------------------------------------------------------------------------------------
template <int G, int NV>
__device__ __forceinline__ void grow_load(Row<NV>& r, const float4* __restrict__ p, int gl, int D4, bool stream) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = k * G + gl;
    if (c < D4) r.v[k] = stream ? ld_stream(p + c) : ld_keep(p + c);
    else r.v[k] = f4_zero();
  }
}
template <int G, int NV>
__device__ __forceinline__ void grow_store(const Row<NV>& r, float4* __restrict__ p, int gl, int D4, bool stream) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = k * G + gl;
    if (c < D4) {
      if (stream) st_stream(p + c, r.v[k]);
      else p[c] = r.v[k];
    }
  }
}

// Per-group staging of a chunk's metadata in shared memory: keys[0] is the key before the chunk,
// keys[1 + s] slot s, keys[1 + cnt] the key after it.  Filled with coalesced loads once per chunk, so
// the slot loop never waits on a dependent global load for its bookkeeping.
struct __align__(16) GroupMeta {
  SlotRec rec[32];
  int32_t keys[36];
};

template <int G, int NV, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_glove_rows_grp(const RowsArgs a) {
  extern __shared__ __align__(16) unsigned char meta_raw[];
  __shared__ float red[32 * 2];
  constexpr int GP = 32 / G;  // chunks per warp
  const int lane = threadIdx.x & 31;
  const int gl = lane % G, grp = lane / G;
  GroupMeta& gm = reinterpret_cast<GroupMeta*>(meta_raw)[(threadIdx.x >> 5) * GP + grp];
  // Blocks are scheduled in index order; the sorted stream ends with the cold rows (singleton
  // segments: three DRAM rows per slot), the expensive chunks.  Walk the chunks from the END so the
  // long blocks start first and the cheap hot-row chunks fill the tail (longest-processing-time first).
  const int64_t n_eff = eff_n(a);
  const int64_t c = eff_chunks(a, n_eff) - 1 - ((blockIdx.x * (int64_t)kWarps + (threadIdx.x >> 5)) * GP + grp);
  const uint32_t D4 = (uint32_t)a.D4;
  const int64_t p0 = c * a.chunk;
  const int cnt = c >= 0 ? (int)min((int64_t)a.chunk, n_eff - p0) : 0;
  const float mbs = a.scalars[ESR_SC_SUM_BS] * a.inv_B;
  const float4* const rows0 = reinterpret_cast<const float4*>(a.rows[0]);
  const float4* const rows1 = reinterpret_cast<const float4*>(a.rows[1]);
  const float4* const accp = reinterpret_cast<const float4*>(a.acc);
  float sums[2] = {0.f, 0.f};  // group-uniform

  // ---- stage the chunk's metadata (lane-parallel, coalesced) ----
  for (int s = gl; s < cnt; s += G) {
    gm.keys[1 + s] = a.skv[p0 + s];
    gm.rec[s] = a.rec[p0 + s];
  }
  if (gl == 0 && cnt > 0) {
    gm.keys[0] = p0 > 0 ? a.skv[p0 - 1] : kNoKey;
    gm.keys[1 + cnt] = p0 + cnt < n_eff ? a.skv[p0 + cnt] : kNoKey;
    gm.keys[2 + cnt] = kNoKey;
  }
  __syncwarp();

  Row<NV> cur, grad;
  row_zero(cur);
  row_zero(grad);
  float bacc = 0.f;
  bool started_here = false;
  int64_t u = 0;
  // The record runs two slots ahead and the PARTNER row one slot ahead of the consumer, so a
  // group always has the next partner row in flight while it works on the current slot (on the
  // Zipf stream most slots need nothing else).
  int32_t key_prev = kNoKey, key_cur = kNoKey, key_next = kNoKey;
  SlotRec rec, rec1;  // slots s and s+1
  rec.code = 0; rec.w = 0.f; rec.t = 0.f; rec.bs = 0.f;
  rec1 = rec;
  Row<NV> Pn;  // partner row of slot s (prefetched during slot s-1)
  row_zero(Pn);
  if (cnt > 0) {
    key_prev = gm.keys[0];
    key_cur = gm.keys[1];
    rec = gm.rec[0];
    if (cnt > 1) rec1 = gm.rec[1];
    u = a.useg[p0];
    const uint32_t q = (uint32_t)(rec.code & kRowMask);
    grow_load<G, NV>(Pn, ((rec.code >> 31) & 1 ? rows1 : rows0) + (uint64_t)q * D4, gl, a.D4, false);
  }
  for (int s = 0; s < a.chunk; ++s) {
    const bool active = s < cnt;
    SlotRec rec2 = rec1;
    Row<NV> P = Pn, A;
    bool is_head = false, is_end = false;
    if (active) {
      key_next = gm.keys[2 + s];
      if (s + 2 < cnt) rec2 = gm.rec[s + 2];
      is_head = key_cur != key_prev;
      is_end = key_cur != key_next;
      const uint32_t row = (uint32_t)(key_cur & kRowMask);
      if (s == 0 || is_head) {
        grow_load<G, NV>(cur, ((key_cur >> 31) & 1 ? rows1 : rows0) + (uint64_t)row * D4, gl, a.D4, false);
        row_zero(grad);
        bacc = 0.f;
        started_here = is_head;
        if (is_head && s > 0) ++u;
      }
      if (!a.emit && is_end && started_here) grow_load<G, NV>(A, accp + (uint64_t)row * D4, gl, a.D4, true);
      if (s + 1 < cnt) {
        const uint32_t q = (uint32_t)(rec1.code & kRowMask);
        grow_load<G, NV>(Pn, ((rec1.code >> 31) & 1 ? rows1 : rows0) + (uint64_t)q * D4, gl, a.D4, false);
      }
    }
    float d = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) d += f4_dot(cur.v[k], P.v[k]);
    const float dot = group_sum<G>(d);
    if (active) {
      const float res = rec.t - dot;
      const float rr = a.per_pair ? res - rec.bs : res;
      const float g = a.c2B * rec.w * (a.per_pair ? rr : res - mbs);
      if (rec.code & kRoleBit) {
        sums[0] = fmaf(rec.w, res, sums[0]);
        sums[1] = fmaf(rec.w * rr, rr, sums[1]);
      }
#pragma unroll
      for (int k = 0; k < NV; ++k) f4_fma(grad.v[k], g, P.v[k]);
      bacc += a.per_pair ? g : rec.bs;
      if (is_end && started_here) {
        if (a.emit) {
          uint64_t e = a.emit_map ? (uint64_t)a.emit_map[u] : (uint64_t)u;
          float* base = a.dE;
          if (a.peers.on) {
            base = a.peers.dE[e >> kEmitShift];
            e &= (1u << kEmitShift) - 1u;
          }
          grow_store<G, NV>(grad, reinterpret_cast<float4*>(base) + e * D4, gl, a.D4, true);
        } else {
          const uint32_t row = (uint32_t)(key_cur & kRowMask);
          const int v = (key_cur >> 31) & 1;
#pragma unroll
          for (int k = 0; k < NV; ++k) adagrad4(cur.v[k], A.v[k], grad.v[k], a.lr, a.eps);
          grow_store<G, NV>(cur, reinterpret_cast<float4*>(a.wrows[1 - v]) + (uint64_t)row * D4, gl, a.D4, true);
          grow_store<G, NV>(A, reinterpret_cast<float4*>(a.acc) + (uint64_t)row * D4, gl, a.D4, true);
        }
        if (gl == 0) a.bsum[u] = bacc;
      } else if (is_end || s == cnt - 1) {
        // leading partial (segment started in an earlier chunk) or the chunk ends inside a segment
        const int slot = (!is_end && started_here) ? 1 : 0;
        grow_store<G, NV>(grad, reinterpret_cast<float4*>(a.part) + (uint64_t)(c * 2 + slot) * D4, gl, a.D4, false);
        if (gl == 0) {
          a.parts[c * 2 + slot] = bacc;
          if (slot == 1) {  // this chunk holds the head of a straddling segment: queue its combine
            const int64_t np = (a.seg_off[u + 1] - 1) / a.chunk - c + 1;
            const int heavy = np > kHeavyParts;
            const int e = atomicAdd(a.wl_count + heavy, 1);
            (heavy ? a.wl_heavy : a.wl_light)[e] = (int32_t)c;
          }
        }
      }
    }
    key_prev = key_cur;
    key_cur = key_next;
    rec = rec1;
    rec1 = rec2;
  }
  // per-block S1 / S2 partials: lane 0 of every group carries its group's value
  float v[2] = {gl == 0 ? sums[0] : 0.f, gl == 0 ? sums[1] : 0.f};
  block_sum<2>(v, red);
  if (threadIdx.x == 0) {
    a.rows_blk[blockIdx.x * 2 + 0] = v[0];
    a.rows_blk[blockIdx.x * 2 + 1] = v[1];
  }
}

// ---------------------------------------------------------------------------------------------
// Phase 2, group variant with ASYNC STAGING (default for D >= 128): same work split as
// k_glove_rows_grp, but the rows travel global -> shared with cp.async (LDGSTS), every lane
// copying and later reading back its own 16-byte columns, so rows in flight cost no registers
// and the prefetch runs ahead of the consumer:
//   partner row of slot s+2, self + accumulator rows of slot s+1 are in flight while slot s is
//   consumed (ncu on the register variant: 45 % of the stall samples sat on the first use of a
//   row that was issued only one slot earlier).
// Per group: 3 partner buffers + 1 self + 1 accumulator buffer of R bytes.  Two commit groups per
// slot, oldest first: {self, acc of s+1}, {partner of s+2}; cp.async.wait_group 1 at the top of
// slot s therefore guarantees partner(s), self(s), acc(s) and leaves partner(s+1) in flight.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool keep) {
  if (keep) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

template <int G, int NV>
__device__ __forceinline__ void grow_copy_async(uint32_t dst, const float4* __restrict__ src, int gl, int D4, bool keep) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = k * G + gl;
    if (c < D4) cp_async16(dst + 16u * c, src + c, keep);
  }
}
template <int G, int NV>
__device__ __forceinline__ void grow_from_smem(Row<NV>& r, const float4* sp, int gl, int D4) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = k * G + gl;
    r.v[k] = c < D4 ? sp[c] : f4_zero();
  }
}

// FULLD: D4 == G * NV, every lane column exists -> no per-column bounds checks in the hot loop
template <int G, int NV, bool FULLD>
__device__ __forceinline__ void grow_copy_async_f(uint32_t dst, const float4* __restrict__ src, int gl, int D4, bool keep) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = k * G + gl;
    if (FULLD || c < D4) cp_async16(dst + 16u * c, src + c, keep);
  }
}
template <int G, int NV, bool FULLD>
__device__ __forceinline__ void grow_from_smem_f(Row<NV>& r, const float4* sp, int gl, int D4) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = k * G + gl;
    r.v[k] = (FULLD || c < D4) ? sp[c] : f4_zero();
  }
}
template <int G, int NV, bool FULLD>
__device__ __forceinline__ void grow_store_f(const Row<NV>& r, float4* __restrict__ p, int gl, int D4, bool stream) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = k * G + gl;
    if (FULLD || c < D4) {
      if (stream) st_stream(p + c, r.v[k]);
      else p[c] = r.v[k];
    }
  }
}

template <int G, int NV, int MINB, bool FULLD>
__global__ void __launch_bounds__(kThreads, MINB) k_glove_rows_grp_async(const RowsArgs a) {
  extern __shared__ __align__(16) unsigned char dyn_raw[];
  __shared__ float red[32 * 2];
  constexpr int GP = 32 / G;  // chunks per warp
  constexpr int NB = 5;  // row buffers per group: P0 P1 P2 S A
  const int lane = threadIdx.x & 31;
  const int gl = lane % G, grp = lane / G;
  const int gidx = (threadIdx.x >> 5) * GP + grp;
  GroupMeta& gm = reinterpret_cast<GroupMeta*>(dyn_raw)[gidx];
  const uint32_t D4 = (uint32_t)a.D4;
  const uint32_t RB = D4 * 16u;
  unsigned char* bufs = dyn_raw + (size_t)kWarps * GP * sizeof(GroupMeta) + (size_t)gidx * NB * RB;
  const uint32_t bufs_u32 = (uint32_t)__cvta_generic_to_shared(bufs);
  const float mbs = a.scalars[ESR_SC_SUM_BS] * a.inv_B;
  const float4* const rows0 = reinterpret_cast<const float4*>(a.rows[0]);
  const float4* const rows1 = reinterpret_cast<const float4*>(a.rows[1]);
  const float4* const accp = reinterpret_cast<const float4*>(a.acc);
  float sums[2] = {0.f, 0.f};
  // Persistent warps pull work items (GP consecutive chunks) from a counter.  Items are handed out from
  // the END of the sorted stream: the cold rows (singleton segments, three DRAM rows per slot) are the
  // long items, so they start first and the cheap hot-row chunks fill in behind them (LPT order), and a
  // warp that drew cheap items simply draws more of them.
  const int64_t n_eff = eff_n(a), nchunks = eff_chunks(a, n_eff);
  const int64_t nitems = (nchunks + GP - 1) / GP;
  // a.interleave: even draws come from the END of the sorted stream (cold rows), odd draws from its FRONT (the Zipf head:
  // partner reads only, served by L2), so DRAM-bound and L2-bound items are in flight together instead of one after the other
  const int64_t n_end = a.interleave ? (nitems + 1) / 2 : nitems;
  const int64_t front_limit = nchunks - n_end * GP;  // chunks below this index belong to the front draws
  for (;;) {
  int64_t item = 0;
  if (lane == 0) item = atomicAdd(a.work_counter, 1);
  item = __shfl_sync(FULL, item, 0);
  if (item >= nitems) break;
  int64_t c;
  bool valid;
  if (!a.interleave || (item & 1) == 0) {
    c = nchunks - 1 - ((a.interleave ? item >> 1 : item) * GP + grp);
    valid = c >= 0 && (!a.interleave || c >= front_limit);
  } else {
    c = (item >> 1) * GP + grp;
    valid = c < front_limit;
  }
  const int64_t p0 = c * a.chunk;
  const int cnt = valid ? (int)min((int64_t)a.chunk, n_eff - p0) : 0;
  __syncwarp();  // the previous item's readers are done with gm
  for (int s = gl; s < cnt; s += G) {
    gm.keys[1 + s] = a.skv[p0 + s];
    gm.rec[s] = a.rec[p0 + s];
  }
  if (gl == 0 && cnt > 0) {
    gm.keys[0] = p0 > 0 ? a.skv[p0 - 1] : kNoKey;
    gm.keys[1 + cnt] = p0 + cnt < n_eff ? a.skv[p0 + cnt] : kNoKey;
    gm.keys[2 + cnt] = kNoKey;
    gm.keys[3 + cnt] = kNoKey;
  }
  __syncwarp();

  // issue helpers ------------------------------------------------------------------------------
  auto issue_partner = [&](int s) {  // partner row of slot s -> P[s % 3]
    if (s < cnt) {
      const int32_t code = gm.rec[s].code;
      const uint32_t q = (uint32_t)(code & kRowMask);
      grow_copy_async_f<G, NV, FULLD>(bufs_u32 + (uint32_t)(s % 3) * RB, ((code >> 31) & 1 ? rows1 : rows0) + (uint64_t)q * D4,
                                      gl, a.D4, true);
    }
  };
  // self / accumulator rows of slot s (self: chunk start or segment head; acc: a segment that both
  // starts and ends inside this chunk closes at s).  `started` = the segment containing s started here.
  auto issue_self_acc = [&](int s, bool started_if_not_head) {
    if (s < cnt) {
      const int32_t k0 = gm.keys[s], k1 = gm.keys[1 + s], k2 = gm.keys[2 + s];
      const bool head = k1 != k0, end = k1 != k2;
      const uint32_t row = (uint32_t)(k1 & kRowMask);
      if (s == 0 || head)
        grow_copy_async_f<G, NV, FULLD>(bufs_u32 + 3u * RB, ((k1 >> 31) & 1 ? rows1 : rows0) + (uint64_t)row * D4, gl, a.D4, true);
      if (!a.emit && end && (head || started_if_not_head))
        grow_copy_async_f<G, NV, FULLD>(bufs_u32 + 4u * RB, accp + (uint64_t)row * D4, gl, a.D4, false);
    }
  };
  Row<NV> cur, grad;
  row_zero(cur);
  row_zero(grad);
  float bacc = 0.f;
  bool started_here = false;
  int64_t u = cnt > 0 ? a.useg[p0] : 0;
  // prologue: {self, acc of slot 0} {partner 0} {nothing} {partner 1}
  issue_self_acc(0, false);
  cp_async_commit();
  issue_partner(0);
  cp_async_commit();
  cp_async_commit();
  issue_partner(1);
  cp_async_commit();

  // EMIT: destination of the gradient row of the segment in flight, loaded one segment ahead (it sat as a dependent
  // global load in front of every gradient store)
  const bool emit_mapped = a.emit && a.emit_map != nullptr;
  uint32_t e_cur = (emit_mapped && cnt > 0) ? (uint32_t)a.emit_map[u] : 0u, e_nx = 0u;
  for (int s = 0; s < a.chunk; ++s) {
    const bool active = s < cnt;
    cp_async_wait1();  // partner(s), self(s), acc(s) have landed; partner(s+1) may still be in flight
    Row<NV> P, A;   // P, A, rec are only consumed when the slot is active (the group_sum must still run warp-wide)
    bool is_head = false, is_end = false;
    int32_t key_cur = kNoKey;
    SlotRec rec;
    if (active) {
      const int32_t k0 = gm.keys[s], k2 = gm.keys[2 + s];
      key_cur = gm.keys[1 + s];
      rec = gm.rec[s];
      is_head = key_cur != k0;
      is_end = key_cur != k2;
      grow_from_smem_f<G, NV, FULLD>(P, reinterpret_cast<const float4*>(bufs + (size_t)(s % 3) * RB), gl, a.D4);
      if (s == 0 || is_head) {
        grow_from_smem_f<G, NV, FULLD>(cur, reinterpret_cast<const float4*>(bufs + 3 * (size_t)RB), gl, a.D4);
        row_zero(grad);
        bacc = 0.f;
        started_here = is_head;
        if (is_head && s > 0) {
          ++u;
          e_cur = e_nx;
        }
      }
      if (emit_mapped && is_end && s + 1 < cnt) e_nx = (uint32_t)a.emit_map[u + 1];  // the next segment starts at s + 1
      if (!a.emit && is_end && started_here)
        grow_from_smem_f<G, NV, FULLD>(A, reinterpret_cast<const float4*>(bufs + 4 * (size_t)RB), gl, a.D4);
    }
    // every buffer read above is private to the lane that filled it: refill without a barrier
    issue_self_acc(s + 1, started_here && !is_end);
    cp_async_commit();
    issue_partner(s + 2);
    cp_async_commit();

    float d0 = 0.f, d1 = 0.f;  // packed pair accumulators (FFMA2): the same tree in both roles of a pair
#pragma unroll
    for (int k = 0; k < NV; ++k) f4_dot2(d0, d1, cur.v[k], P.v[k]);
    const float dot = group_sum<G>(d0 + d1);
    if (active) {
      const float res = rec.t - dot;
      const float rr = a.per_pair ? res - rec.bs : res;
      const float g = a.c2B * rec.w * (a.per_pair ? rr : res - mbs);
      if (rec.code & kRoleBit) {
        sums[0] = fmaf(rec.w, res, sums[0]);
        sums[1] = fmaf(rec.w * rr, rr, sums[1]);
      }
#pragma unroll
      for (int k = 0; k < NV; ++k) f4_fma2(grad.v[k], g, P.v[k]);
      bacc += a.per_pair ? g : rec.bs;
      if (is_end && started_here) {
        if (a.emit) {
          uint64_t e = a.emit_map ? (uint64_t)e_cur : (uint64_t)u;
          float* base = a.dE;
          if (a.peers.on) {
            base = a.peers.dE[e >> kEmitShift];
            e &= (1u << kEmitShift) - 1u;
          }
          grow_store_f<G, NV, FULLD>(grad, reinterpret_cast<float4*>(base) + e * D4, gl, a.D4, true);
        } else {
          const uint32_t row = (uint32_t)(key_cur & kRowMask);
          const int v = (key_cur >> 31) & 1;
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            if (a.eps > 0.f) adagrad4_packed(cur.v[k], A.v[k], grad.v[k], a.lr, a.eps);
            else adagrad4(cur.v[k], A.v[k], grad.v[k], a.lr, a.eps);
          }
          grow_store_f<G, NV, FULLD>(cur, reinterpret_cast<float4*>(a.wrows[1 - v]) + (uint64_t)row * D4, gl, a.D4, true);
          grow_store_f<G, NV, FULLD>(A, reinterpret_cast<float4*>(a.acc) + (uint64_t)row * D4, gl, a.D4, true);
        }
        if (gl == 0) a.bsum[u] = bacc;
      } else if (is_end || s == cnt - 1) {
        const int slot = (!is_end && started_here) ? 1 : 0;
        grow_store_f<G, NV, FULLD>(grad, reinterpret_cast<float4*>(a.part) + (uint64_t)(c * 2 + slot) * D4, gl, a.D4, false);
        if (gl == 0) {
          a.parts[c * 2 + slot] = bacc;
          if (slot == 1) {
            const int64_t np = (a.seg_off[u + 1] - 1) / a.chunk - c + 1;
            const int heavy = np > kHeavyParts;
            const int e = atomicAdd(a.wl_count + heavy, 1);
            (heavy ? a.wl_heavy : a.wl_light)[e] = (int32_t)c;
          }
        }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  }  // work loop
  float v[2] = {gl == 0 ? sums[0] : 0.f, gl == 0 ? sums[1] : 0.f};
  block_sum<2>(v, red);
  if (threadIdx.x == 0) {
    a.rows_blk[blockIdx.x * 2 + 0] = v[0];
    a.rows_blk[blockIdx.x * 2 + 1] = v[1];
  }
}

// ---------------------------------------------------------------------------------------------
// Phase 2, TMA variant: rows are staged into shared memory by the bulk-copy engine
// (cp.async.bulk global -> shared, completion on an mbarrier; SASS: UBLKCP).  Every warp owns a
// private ring of NS stages; a stage holds the three rows one slot can need (partner, self,
// accumulator).  The in-flight data lives in shared memory instead of registers, so a warp keeps
// NS slots (up to 3*NS rows) outstanding while it consumes, and the low register count lets
// 32 warps per SM stay resident.  The kernel is persistent: warps stride over the chunks.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded spin: a lost arrival traps (a launch error the host sees) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!ok && spin > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                              uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}

template <int NK>
__device__ __forceinline__ void row_from_smem(Row<NK>& r, const float4* sp, int lane, int D4) {
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const int c = k * 32 + lane;
    r.v[k] = c < D4 ? sp[c] : f4_zero();
  }
}

constexpr int kTmaStages = 4;
constexpr int tma_min_blocks(int nk) { return nk == 1 ? 4 : (nk == 2 ? 2 : 1); }

template <int NK>
__global__ void __launch_bounds__(kThreads, tma_min_blocks(NK)) k_glove_rows_tma(const RowsArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bars[kWarps][kTmaStages];
  __shared__ float red[32 * 2];
  constexpr int NS = kTmaStages;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t RB = (uint32_t)a.D4 * 16u;  // row bytes
  const uint32_t stage_bytes = 3u * RB;
  unsigned char* my = smem_raw + (size_t)wid * NS * stage_bytes;
  const uint32_t my_u32 = smem_u32(my);
  const uint32_t bar0 = smem_u32(&bars[wid][0]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NS; ++i) mbar_init(bar0 + 8u * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  uint64_t pol_stream;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
  uint32_t phase = 0;  // bit i = parity the next wait on stage i expects
  float sums[2] = {0.f, 0.f};
  const float mbs = a.scalars[ESR_SC_SUM_BS] * a.inv_B;
  const int64_t stride = (int64_t)gridDim.x * kWarps;

  const int64_t nchunks_eff = eff_chunks(a, eff_n(a));
  for (int64_t c = blockIdx.x * (int64_t)kWarps + wid; c < nchunks_eff; c += stride) {
    const ChunkMeta m = load_chunk(a, c, lane);
    // lane s issues the copies of slot s (it holds the slot's key and record)
    auto issue = [&](int s) {
      if (lane == s) {
        const int stg = s % NS;
        const uint32_t bar = bar0 + 8u * stg;
        const uint32_t dst = my_u32 + (uint32_t)stg * stage_bytes;
        const bool need_self = (s == 0) || ((m.head_mask >> s) & 1u);
        const bool need_acc = !a.emit && ((m.end_mask >> s) & 1u);
        mbar_expect_tx(bar, RB * (1u + (need_self ? 1u : 0u) + (need_acc ? 1u : 0u)));
        const int64_t q = m.r.code & kRowMask;
        const int qv = (m.r.code >> 31) & 1;
        bulk_g2s(dst, a.rows[qv] + q * (int64_t)a.D4 * 4, RB, bar);
        const int64_t row = m.kl & kRowMask;
        if (need_self) bulk_g2s(dst + RB, a.rows[(m.kl >> 31) & 1] + row * (int64_t)a.D4 * 4, RB, bar);
        if (need_acc) bulk_g2s_hint(dst + 2u * RB, a.acc + row * (int64_t)a.D4 * 4, RB, bar, pol_stream);
      }
    };
#pragma unroll
    for (int s = 0; s < NS; ++s)
      if (s < m.cnt) issue(s);

    SegState<NK> st;
    row_zero(st.cur);
    row_zero(st.grad);
    st.bacc = 0.f;
    st.started_here = false;
    st.cur_key = kNoKey;

    for (int s = 0; s < m.cnt; ++s) {
      const int stg = s % NS;
      const int32_t code = __shfl_sync(FULL, m.r.code, s);
      const float w = __shfl_sync(FULL, m.r.w, s);
      const float t = __shfl_sync(FULL, m.r.t, s);
      const float bs = __shfl_sync(FULL, m.r.bs, s);
      const int32_t key = __shfl_sync(FULL, m.kl, s);
      mbar_wait(bar0 + 8u * stg, (phase >> stg) & 1u);
      phase ^= 1u << stg;
      const float4* sp = reinterpret_cast<const float4*>(my + (size_t)stg * stage_bytes);
      Row<NK> P, Sf, A;
      row_from_smem<NK>(P, sp, lane, a.D4);
      if (s == 0 || ((m.head_mask >> s) & 1u)) row_from_smem<NK>(Sf, sp + a.D4, lane, a.D4);
      if (!a.emit && ((m.end_mask >> s) & 1u)) row_from_smem<NK>(A, sp + 2 * a.D4, lane, a.D4);
      __syncwarp();  // every lane has read the stage: it can be refilled
      if (s + NS < m.cnt) issue(s + NS);
      consume_slot<NK>(a, m, c, s, st, P, Sf, A, code, key, w, t, bs, mbs, sums, lane);
    }
  }
  store_block_sums(a, sums, red, lane);
}

// ---------------------------------------------------------------------------------------------
// Phase 2, group variant with a per-group FIFO of bulk copies (ESR_IMPL_AUTO + cfg.reserved == 2; A/B candidate
// for the default).  Same work split as k_glove_rows_grp_async, different staging: every row a chunk needs
// (per slot, in order: self row at a segment head, accumulator row at a segment end, partner row) goes through
// ONE ring of kFifoBufs generic row buffers per group, filled by cp.async.bulk (one instruction per 512-byte row,
// issued by the group's first lane, completion on a per-buffer mbarrier).  The producer side runs ahead as far as
// the ring allows (up to kFifoAhead slots), so a group that only needs partner rows (most slots of a Zipf
// stream) keeps 4 rows in flight instead of 1-2: the async variant's fixed {self, acc, partner} buffers left
// ~70 % of the staged bytes idle at its wait point (measured: 2800 clk per slot-step, latency-bound).
// ---------------------------------------------------------------------------------------------
constexpr int kFifoBufs = 5;
constexpr int kFifoAhead = 4;

template <int G, int NV, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_glove_rows_grp_fifo(const RowsArgs a) {
  extern __shared__ __align__(16) unsigned char dyn_raw[];
  __shared__ float red[32 * 2];
  constexpr int GP = 32 / G;
  constexpr int NB = kFifoBufs;
  const int lane = threadIdx.x & 31;
  const int gl = lane % G, grp = lane / G;
  const int gidx = (threadIdx.x >> 5) * GP + grp;
  constexpr int kGroups = kWarps * GP;
  GroupMeta& gm = reinterpret_cast<GroupMeta*>(dyn_raw)[gidx];
  const uint32_t D4 = (uint32_t)a.D4;
  const uint32_t RB = D4 * 16u;
  unsigned char* bufs = dyn_raw + (size_t)kGroups * sizeof(GroupMeta) + (size_t)gidx * NB * RB;
  const uint32_t bufs_u32 = smem_u32(bufs);
  uint64_t* bars = reinterpret_cast<uint64_t*>(dyn_raw + (size_t)kGroups * sizeof(GroupMeta) + (size_t)kGroups * NB * RB) + gidx * NB;
  const uint32_t bar0 = smem_u32(bars);
  if (gl == 0) {
#pragma unroll
    for (int i = 0; i < NB; ++i) mbar_init(bar0 + 8u * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  uint64_t pol_stream;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
  uint32_t phase = 0;  // bit b = parity the next wait on buffer b expects
  const float mbs = a.scalars[ESR_SC_SUM_BS] * a.inv_B;
  const float* const rows0 = a.rows[0];
  const float* const rows1 = a.rows[1];
  float sums[2] = {0.f, 0.f};
  const int64_t n_eff = eff_n(a), nchunks = eff_chunks(a, n_eff);
  const int64_t nitems = (nchunks + GP - 1) / GP;
  for (;;) {
    int64_t item = 0;
    if (lane == 0) item = atomicAdd(a.work_counter, 1);
    item = __shfl_sync(FULL, item, 0);
    if (item >= nitems) break;
    const int64_t c = nchunks - 1 - (item * GP + grp);
    const int64_t p0 = c * a.chunk;
    const int cnt = c >= 0 ? (int)min((int64_t)a.chunk, n_eff - p0) : 0;
    __syncwarp();  // the previous item's readers are done with gm
    for (int s = gl; s < cnt; s += G) {
      gm.keys[1 + s] = a.skv[p0 + s];
      gm.rec[s] = a.rec[p0 + s];
    }
    if (gl == 0 && cnt > 0) {
      gm.keys[0] = p0 > 0 ? a.skv[p0 - 1] : kNoKey;
      gm.keys[1 + cnt] = p0 + cnt < n_eff ? a.skv[p0 + cnt] : kNoKey;
      gm.keys[2 + cnt] = kNoKey;
      gm.keys[3 + cnt] = kNoKey;
    }
    __syncwarp();
    // index of the first true head of the chunk: slots before it continue a segment started in an earlier chunk
    int first_head = cnt;
    for (int s = 0; s < cnt; ++s)
      if (gm.keys[1 + s] != gm.keys[s]) {
        first_head = s;
        break;
      }

    // ---- producer side: rows of slot t into the ring, FIFO order {self, acc, partner} ----
    int pb = 0, cb = 0, issued = 0;  // buffers issued / consumed (this chunk), slots issued
    auto slot_needs = [&](int t, bool& need_self, bool& need_acc) {
      const int32_t k0 = gm.keys[t], k1 = gm.keys[1 + t], k2 = gm.keys[2 + t];
      need_self = (t == 0) || (k1 != k0);
      need_acc = !a.emit && (k1 != k2) && t >= first_head;
    };
    auto put = [&](const float* src, bool stream) {
      const int b = pb % NB;
      if (gl == 0) {
        mbar_expect_tx(bar0 + 8u * b, RB);
        if (stream) bulk_g2s_hint(bufs_u32 + (uint32_t)b * RB, src, RB, bar0 + 8u * b, pol_stream);
        else bulk_g2s(bufs_u32 + (uint32_t)b * RB, src, RB, bar0 + 8u * b);
      }
      ++pb;
    };
    auto top_up = [&](int s) {
      while (issued < cnt && issued <= s + kFifoAhead) {
        bool ns, na;
        slot_needs(issued, ns, na);
        if (pb - cb + 1 + (ns ? 1 : 0) + (na ? 1 : 0) > NB) break;
        const int32_t k1 = gm.keys[1 + issued];
        const uint32_t row = (uint32_t)(k1 & kRowMask);
        if (ns) put(((k1 >> 31) & 1 ? rows1 : rows0) + (uint64_t)row * a.D4 * 4, false);
        if (na) put(a.acc + (uint64_t)row * a.D4 * 4, true);
        const int32_t code = gm.rec[issued].code;
        put(((code >> 31) & 1 ? rows1 : rows0) + (uint64_t)(uint32_t)(code & kRowMask) * a.D4 * 4, false);
        ++issued;
      }
    };
    // ---- consumer side: next buffer of the ring -> registers ----
    auto take = [&](Row<NV>& r) {
      const int b = cb % NB;
      mbar_wait(bar0 + 8u * b, (phase >> b) & 1u);
      phase ^= 1u << b;
      grow_from_smem<G, NV>(r, reinterpret_cast<const float4*>(bufs + (size_t)b * RB), gl, a.D4);
      ++cb;
    };

    Row<NV> cur, grad;
    row_zero(cur);
    row_zero(grad);
    float bacc = 0.f;
    bool started_here = false;
    int64_t u = cnt > 0 ? a.useg[p0] : 0;
    for (int s = 0; s < a.chunk; ++s) {
      const bool active = s < cnt;
      __syncwarp();  // every lane has read the buffers released in the previous slot: they may be refilled
      if (active) top_up(s);
      Row<NV> P, A;
      row_zero(P);
      bool is_head = false, is_end = false;
      int32_t key_cur = kNoKey;
      SlotRec rec;
      rec.code = 0; rec.w = 0.f; rec.t = 0.f; rec.bs = 0.f;
      if (active) {
        bool ns, na;
        slot_needs(s, ns, na);
        key_cur = gm.keys[1 + s];
        rec = gm.rec[s];
        is_head = key_cur != gm.keys[s];
        is_end = key_cur != gm.keys[2 + s];
        if (ns) {
          take(cur);
          row_zero(grad);
          bacc = 0.f;
          started_here = is_head;
          if (is_head && s > 0) ++u;
        }
        if (na) take(A);
        take(P);
      }
      float d = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) d += f4_dot(cur.v[k], P.v[k]);
      const float dot = group_sum<G>(d);
      if (active) {
        const float res = rec.t - dot;
        const float rr = a.per_pair ? res - rec.bs : res;
        const float g = a.c2B * rec.w * (a.per_pair ? rr : res - mbs);
        if (rec.code & kRoleBit) {
          sums[0] = fmaf(rec.w, res, sums[0]);
          sums[1] = fmaf(rec.w * rr, rr, sums[1]);
        }
#pragma unroll
        for (int k = 0; k < NV; ++k) f4_fma(grad.v[k], g, P.v[k]);
        bacc += a.per_pair ? g : rec.bs;
        if (is_end && started_here) {
          if (a.emit) {
            uint64_t e = a.emit_map ? (uint64_t)a.emit_map[u] : (uint64_t)u;
            float* base = a.dE;
            if (a.peers.on) {
              base = a.peers.dE[e >> kEmitShift];
              e &= (1u << kEmitShift) - 1u;
            }
            grow_store<G, NV>(grad, reinterpret_cast<float4*>(base) + e * D4, gl, a.D4, true);
          } else {
            const uint32_t row = (uint32_t)(key_cur & kRowMask);
            const int v = (key_cur >> 31) & 1;
#pragma unroll
            for (int k = 0; k < NV; ++k) adagrad4(cur.v[k], A.v[k], grad.v[k], a.lr, a.eps);
            grow_store<G, NV>(cur, reinterpret_cast<float4*>(a.wrows[1 - v]) + (uint64_t)row * D4, gl, a.D4, true);
            grow_store<G, NV>(A, reinterpret_cast<float4*>(a.acc) + (uint64_t)row * D4, gl, a.D4, true);
          }
          if (gl == 0) a.bsum[u] = bacc;
        } else if (is_end || s == cnt - 1) {
          const int slot = (!is_end && started_here) ? 1 : 0;
          grow_store<G, NV>(grad, reinterpret_cast<float4*>(a.part) + (uint64_t)(c * 2 + slot) * D4, gl, a.D4, false);
          if (gl == 0) {
            a.parts[c * 2 + slot] = bacc;
            if (slot == 1) {
              const int64_t np = (a.seg_off[u + 1] - 1) / a.chunk - c + 1;
              const int heavy = np > kHeavyParts;
              const int e = atomicAdd(a.wl_count + heavy, 1);
              (heavy ? a.wl_heavy : a.wl_light)[e] = (int32_t)c;
            }
          }
        }
      }
    }
  }  // work loop
  float v[2] = {gl == 0 ? sums[0] : 0.f, gl == 0 ? sums[1] : 0.f};
  block_sum<2>(v, red);
  if (threadIdx.x == 0) {
    a.rows_blk[blockIdx.x * 2 + 0] = v[0];
    a.rows_blk[blockIdx.x * 2 + 1] = v[1];
  }
}

// ---------------------------------------------------------------------------------------------
// Combine: adds the per-chunk partial sums of every segment that straddles chunk boundaries, in
// chunk order with a fixed tree (deterministic), and closes the segment.  The row pass queued the
// head chunk of each such segment: short ones (<= kHeavyParts partials) are taken one per warp,
// the Zipf-head rows (thousands of partials) one per block.  Block 0 first reduces the S1/S2
// block partials into scalars[3..4].
// ---------------------------------------------------------------------------------------------
struct Straddler {
  int32_t key;
  int64_t u, np;
};

__device__ __forceinline__ Straddler straddler_of(const RowsArgs& a, int64_t c) {
  Straddler s;
  const int64_t pl = min((c + 1) * (int64_t)a.chunk, eff_n(a)) - 1;
  s.key = a.skv[pl];
  s.u = a.useg[pl];
  s.np = (a.seg_off[s.u + 1] - 1) / a.chunk - c + 1;
  return s;
}

// partial m of the segment whose head chunk is c: m == 0 -> part[c][1]; m >= 1 -> part[c+m][0]
__device__ __forceinline__ int64_t part_index(int64_t c, int64_t m) { return m == 0 ? c * 2 + 1 : (c + m) * 2; }

template <int NK, int UN>
__device__ __forceinline__ void sum_parts(const RowsArgs& a, int64_t c, int64_t m0, int64_t m1, Row<NK>& sum,
                                          float& ssum, int lane) {
  for (int64_t m = m0; m < m1; m += UN) {
    Row<NK> x[UN];
    float xs[UN];
#pragma unroll
    for (int j = 0; j < UN; ++j) {
      if (m + j < m1) {
        const int64_t idx = part_index(c, m + j);
        row_load<NK, true>(x[j], reinterpret_cast<const float4*>(a.part) + idx * a.D4, lane, a.D4);
        xs[j] = a.parts[idx];
      }
    }
#pragma unroll
    for (int j = 0; j < UN; ++j) {
      if (m + j < m1) {
        row_add<NK>(sum, x[j]);
        ssum += xs[j];
      }
    }
  }
}

template <int NK>
__device__ __forceinline__ void close_straddler(const RowsArgs& a, const Straddler& s, const Row<NK>& sum, float bacc,
                                                int lane) {
  Row<NK> self, accrow;
  row_zero(self);
  row_zero(accrow);
  if (!a.emit) {
    const int64_t row = s.key & kRowMask;
    const int v = (s.key >> 31) & 1;
    row_load<NK, false>(self, reinterpret_cast<const float4*>(a.rows[v]) + row * a.D4, lane, a.D4);
    row_load<NK, true>(accrow, reinterpret_cast<const float4*>(a.acc) + row * a.D4, lane, a.D4);
  }
  close_segment<NK>(a, s.key, s.u, self, accrow, sum, bacc, lane);
}

template <int NK>
__global__ void __launch_bounds__(kCombineThreads) k_glove_combine(const RowsArgs a, int row_blocks,
                                                                   float* __restrict__ scalars) {
  constexpr int UN = NK == 1 ? 8 : 4;
  __shared__ float4 sh[kCombineWarps][NK * 32];
  __shared__ float shs[kCombineWarps];
  if (blockIdx.x == 0) reduce_partials<2>(a.rows_blk, row_blocks, scalars + ESR_SC_S1);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // light work list: one warp per straddling segment
  const int nl = a.wl_count[0];
  for (int e = blockIdx.x * kCombineWarps + wid; e < nl; e += gridDim.x * kCombineWarps) {
    const int64_t c = a.wl_light[e];
    const Straddler s = straddler_of(a, c);
    Row<NK> sum;
    row_zero(sum);
    float bacc = 0.f;
    sum_parts<NK, UN>(a, c, 0, s.np, sum, bacc, lane);
    close_straddler<NK>(a, s, sum, bacc, lane);
  }
  // heavy work list: kHeavySplit blocks per segment, each sums a contiguous range of the partials (warps
  // take contiguous sub-ranges) into a second-level partial; the LAST block to arrive adds those in
  // range order and closes the segment -- the summation tree is fixed, whichever block that is.
  __shared__ int ticket_sh;
  const int nh = a.wl_count[1];
  for (int task = blockIdx.x; task < nh * kHeavySplit; task += gridDim.x) {
    const int e = task / kHeavySplit, sp = task % kHeavySplit;
    const int64_t c = a.wl_heavy[e];
    const Straddler s = straddler_of(a, c);
    const int64_t per_split = ceil_div(s.np, (int64_t)kHeavySplit);
    const int64_t r0 = min(s.np, sp * per_split), r1 = min(s.np, r0 + per_split);
    const int64_t per = ceil_div(r1 - r0, (int64_t)kCombineWarps);
    const int64_t m0 = min(r1, r0 + wid * per), m1 = min(r1, m0 + per);
    Row<NK> sum;
    row_zero(sum);
    float ssum = 0.f;
    sum_parts<NK, UN>(a, c, m0, m1, sum, ssum, lane);
#pragma unroll
    for (int k = 0; k < NK; ++k) sh[wid][k * 32 + lane] = sum.v[k];
    if (lane == 0) shs[wid] = ssum;
    __syncthreads();
    if (wid == 0) {
      float bacc = shs[0];
      for (int w = 1; w < kCombineWarps; ++w) {
#pragma unroll
        for (int k = 0; k < NK; ++k) f4_add(sum.v[k], sh[w][k * 32 + lane]);
        bacc += shs[w];
      }
      float4* p2 = reinterpret_cast<float4*>(a.part2) + ((int64_t)e * kHeavySplit + sp) * a.D4;
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        const int col = k * 32 + lane;
        if (col < a.D4) __stcg(p2 + col, sum.v[k]);
      }
      if (lane == 0) {
        __stcg(a.parts2 + (int64_t)e * kHeavySplit + sp, bacc);
        __threadfence();
        ticket_sh = atomicAdd(a.tickets + e, 1);
      }
      __syncwarp();
      if (ticket_sh == kHeavySplit - 1) {
        __threadfence();
        Row<NK> tot;
        row_zero(tot);
        float btot = 0.f;
        for (int q = 0; q < kHeavySplit; ++q) {
          const float4* src = reinterpret_cast<const float4*>(a.part2) + ((int64_t)e * kHeavySplit + q) * a.D4;
#pragma unroll
          for (int k = 0; k < NK; ++k) {
            const int col = k * 32 + lane;
            if (col < a.D4) f4_add(tot.v[k], __ldcg(src + col));
          }
          btot += __ldcg(a.parts2 + (int64_t)e * kHeavySplit + q);
        }
        close_straddler<NK>(a, s, tot, btot, lane);
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// Phase 3: bias gradient + Adagrad, version flip, loss.
// ---------------------------------------------------------------------------------------------
struct FinishArgs {
  const int32_t* uniq;
  const int32_t* seg_off;
  const int32_t* n_uniq;
  const float* bsum;
  float* scalars;
  float* bias;
  float* bias_acc;
  uint8_t* ver;
  float* db;
  const int32_t* emit_map;
  EmitPeers peers;
  int32_t per_pair;
  int32_t emit;
  float B;  // B_global
  float lr, eps;
  float* loss_log;
  int32_t* loss_step;
  int32_t loss_log_len;
  float* loss_host;
};

__global__ void __launch_bounds__(kThreads) k_glove_finish(const FinishArgs a) {
  const int64_t u = blockIdx.x * (int64_t)kThreads + threadIdx.x;
  const float S0 = a.scalars[ESR_SC_S0], S1 = a.scalars[ESR_SC_S1];
  if (u < *a.n_uniq) {
    const float nslots = (float)(a.seg_off[u + 1] - a.seg_off[u]);
    // App. A.1: db[v] = sum_slots -(2/B^2) (S1 - bs_r S0) = -(2/B^2) (n_v S1 - S0 sum bs_r); A.2: sum g
    const float gb = a.per_pair ? a.bsum[u] : (-2.f / (a.B * a.B)) * (nslots * S1 - S0 * a.bsum[u]);
    if (a.emit) {
      int64_t e = a.emit_map ? a.emit_map[u] : u;
      float* base = a.db;
      if (a.peers.on) {
        base = a.peers.db[e >> kEmitShift];
        e &= (1 << kEmitShift) - 1;
      }
      base[e] = gb;
    } else {
      const int64_t row = a.uniq[u];
      float p = a.bias[row], ac = a.bias_acc[row];
      adagrad1(p, ac, gb, a.lr, a.eps);
      a.bias[row] = p;
      a.bias_acc[row] = ac;
      if (a.ver) a.ver[row] ^= 1;
    }
  }
  if (u == 0) {
    const float S2 = a.scalars[ESR_SC_S2];
    float loss;
    if (a.per_pair) {
      loss = S2 / a.B;
    } else {
      const float mbs = a.scalars[ESR_SC_SUM_BS] / a.B, mbs2 = a.scalars[ESR_SC_SUM_BS2] / a.B;
      loss = (S2 - 2.f * mbs * S1 + mbs2 * S0) / a.B;  // App. A.1 closed form of the (B,B) mean
    }
    a.scalars[ESR_SC_LOSS] = loss;
    if (a.loss_log != nullptr) {  // device-side slot: a replayed graph logs every step
      const int32_t t = *a.loss_step;
      a.loss_log[t % a.loss_log_len] = loss;
      if (a.loss_host != nullptr) a.loss_host[t % a.loss_log_len] = loss;  // pinned host mirror (posted PCIe write)
      *a.loss_step = t + 1;
    }
  }
}

bool glove_table_ok(const EsrTable* t, bool update) {
  if (t == nullptr || t->struct_size < sizeof(EsrTable) || t->D <= 0 || (t->D % 4) != 0 || t->D > 512) return false;
  if (t->V <= 0 || t->V > (int64_t)kRowMask) return false;
  if (!t->rows[0] || (reinterpret_cast<uintptr_t>(t->rows[0]) % 16) != 0 || !t->bias) return false;
  if (t->ver && !t->rows[1]) return false;
  if (update) {
    // the batch-synchronous in-place update needs the second buffer + versions
    if (!t->rows[1] || !t->ver || !t->acc || !t->bias_acc) return false;
    if ((reinterpret_cast<uintptr_t>(t->rows[1]) % 16) != 0 || (reinterpret_cast<uintptr_t>(t->acc) % 16) != 0) return false;
  }
  return true;
}

bool cfg_ok(const EsrGloveCfg* cfg, const EsrPlan* plan) {
  if (!cfg || cfg->struct_size < sizeof(EsrGloveCfg) || !plan || plan->struct_size < sizeof(EsrPlan)) return false;
  // B is a CAPACITY when the plan carries n_valid (row-sharded path: the pairs a rank processes are counted on the device),
  // so only then may it exceed the global normaliser
  if (cfg->B < 0 || plan->n_slots != 2 * cfg->B || (cfg->B > 0 && cfg->B_global <= 0)) return false;
  if (plan->n_valid == nullptr && cfg->B_global < cfg->B) return false;
  if (cfg->bias_mode != ESR_BIAS_REFERENCE_BROADCAST && cfg->bias_mode != ESR_BIAS_PER_PAIR) return false;
  if (cfg->rows_mode != ESR_ROWS_UPDATE && cfg->rows_mode != ESR_ROWS_EMIT_GRADS) return false;
  return true;
}


void load_emit_peers(const EsrGloveCfg* cfg, EmitPeers* p) {
  p->on = 0;
  for (int i = 0; i < ESR_MAX_PEERS; ++i) p->dE[i] = p->db[i] = nullptr;
  if (cfg->emit_map && cfg->emit_peers_dE && cfg->emit_peers_db && cfg->n_emit_peers > 0) {
    p->on = 1;
    for (int i = 0; i < cfg->n_emit_peers && i < ESR_MAX_PEERS; ++i) {
      p->dE[i] = static_cast<float*>(cfg->emit_peers_dE[i]);
      p->db[i] = static_cast<float*>(cfg->emit_peers_db[i]);
    }
  }
}

RowsArgs make_rows_args(const EsrTable* t, const EsrPlan* plan, const EsrGloveCfg* cfg, const GloveWs& w,
                        const float* scalars, float* dE) {
  RowsArgs a;
  a.rows[0] = t->rows[0];
  a.rows[1] = t->rows[1] ? t->rows[1] : t->rows[0];
  a.wrows[0] = t->rows[0];
  a.wrows[1] = t->rows[1];
  a.acc = t->acc;
  a.skv = w.skv;
  a.rec = w.rec;
  a.useg = plan->useg;
  a.seg_off = plan->seg_off;
  a.wl_count = w.wl_count;
  a.work_counter = w.wl_count + 2;
  a.wl_light = w.wl_light;
  a.wl_heavy = w.wl_heavy;
  a.tickets = w.wl_count + 4;
  a.part2 = w.part2;
  a.parts2 = w.parts2;
  a.nchunks = w.nchunks;
  a.scalars = scalars;
  a.bsum = w.bsum;
  a.part = w.part;
  a.parts = w.parts;
  a.rows_blk = w.rows_blk;
  a.dE = dE;
  a.emit_map = cfg->emit_map;
  load_emit_peers(cfg, &a.peers);
  a.n = plan->n_slots;
  a.n_valid = plan->n_valid;
  a.interleave = cfg->reserved == 7 ? 0 : 1;  // default since round 2 (148.5 vs 157.1 us per step); 7 = the old end-first order (A/B)
  a.D4 = t->D / 4;
  a.chunk = w.chunk;
  a.per_pair = cfg->bias_mode == ESR_BIAS_PER_PAIR;
  a.emit = cfg->rows_mode == ESR_ROWS_EMIT_GRADS;
  a.c2B = -2.f / (float)cfg->B_global;
  a.inv_B = 1.f / (float)cfg->B_global;
  a.lr = cfg->lr;
  a.eps = cfg->eps;
  return a;
}

}  // namespace
}  // namespace esr

using namespace esr;

extern "C" size_t esr_glove_workspace_bytes(int64_t B, int32_t D, int32_t chunk) {
  if (B < 0 || D <= 0) return 0;
  return carve_ws(nullptr, B, D, chunk, nullptr) + 256;
}

extern "C" int esr_glove_prep_f32(const EsrTable* t, const EsrPlan* plan, const float* counts, const EsrGloveCfg* cfg,
                                  float* scalars, void* ws, size_t ws_bytes, esr_stream_t stream_) {
  ESR_RANGE("esr_glove_prep_f32");
  ESR_REQUIRE(cfg_ok(cfg, plan) && glove_table_ok(t, cfg->rows_mode == ESR_ROWS_UPDATE) && scalars != nullptr);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ESR_CUDA(cudaMemsetAsync(scalars, 0, sizeof(float) * ESR_GLOVE_NSCAL, stream));
  if (cfg->B == 0) return ESR_OK;
  ESR_REQUIRE(counts && ws && plan->sorted_keys && plan->perm && plan->partner);
  if (ws_bytes < esr_glove_workspace_bytes(cfg->B, t->D, cfg->chunk)) return ESR_EWORKSPACE;
  GloveWs w;
  carve_ws(ws, cfg->B, t->D, cfg->chunk, &w);
  const int64_t n = plan->n_slots;
  k_glove_prep<<<w.prep_blocks, kThreads, 0, stream>>>(plan->sorted_keys, plan->perm, plan->partner, counts, t->bias,
                                                       t->ver, n, plan->n_valid, cfg->B, cfg->x_max, cfg->alpha, w.skv, w.rec,
                                                       w.prep_blk, scalars, w.wl_count, (int32_t)(4 + w.heavy_cap));
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

template <int G, int NV, int NKC>
static int launch_rows_grp(const RowsArgs& a, const GloveWs& w, float* scalars, int phases, cudaStream_t stream,
                           bool use_async = false, int grid_override = 0, bool fifo = false) {
  constexpr int GP = 32 / G;
  int row_blocks = (int)ceil_div(w.nchunks, (int64_t)kWarps * GP);
  if (use_async) {  // persistent, work-stealing: 2 CTAs per SM unless the caller leaves room for a concurrent stream
    const int64_t cap = grid_override > 0 ? grid_override : 2 * (int64_t)sm_count();
    row_blocks = (int)std::min<int64_t>(row_blocks, cap);
  }
  if ((phases & 1) && use_async && fifo) {
    const size_t smem = (size_t)kWarps * GP * (sizeof(GroupMeta) + kFifoBufs * ((size_t)a.D4 * 16 + 8));
    static SmemOptIn configured_fifo;  // per <G, NV> instantiation
    if (smem > 48 * 1024 && configured_fifo.raise(smem))
      ESR_CUDA(cudaFuncSetAttribute(k_glove_rows_grp_fifo<G, NV, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_glove_rows_grp_fifo<G, NV, 2><<<row_blocks, kThreads, smem, stream>>>(a);
    ESR_LAUNCH_CHECK();
  } else if ((phases & 1) && use_async) {
    const size_t smem = (size_t)kWarps * GP * (sizeof(GroupMeta) + 5 * (size_t)a.D4 * 16);
    static SmemOptIn configured;  // per <G, NV> instantiation
    if (smem > 48 * 1024 && configured.raise(smem)) {
      ESR_CUDA(cudaFuncSetAttribute(k_glove_rows_grp_async<G, NV, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
      ESR_CUDA(cudaFuncSetAttribute(k_glove_rows_grp_async<G, NV, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
    }
    if (a.D4 == G * NV) k_glove_rows_grp_async<G, NV, 2, true><<<row_blocks, kThreads, smem, stream>>>(a);
    else k_glove_rows_grp_async<G, NV, 2, false><<<row_blocks, kThreads, smem, stream>>>(a);
    ESR_LAUNCH_CHECK();
  } else if (phases & 1) {
    constexpr size_t smem = (size_t)kWarps * GP * sizeof(GroupMeta);
    static SmemOptIn configured;  // per <G, NV> instantiation
    if (smem > 48 * 1024 && configured.raise(smem))
      ESR_CUDA(cudaFuncSetAttribute(k_glove_rows_grp<G, NV, (NV <= 2 ? 3 : 2)>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_glove_rows_grp<G, NV, (NV <= 2 ? 3 : 2)><<<row_blocks, kThreads, smem, stream>>>(a);
    ESR_LAUNCH_CHECK();
  }
  if (phases & 2) {
    k_glove_combine<NKC><<<2 * sm_count(), kCombineThreads, 0, stream>>>(a, row_blocks, scalars);
    ESR_LAUNCH_CHECK();
  }
  return ESR_OK;
}

template <int NK, int S, int MINB>
static int launch_rows(const RowsArgs& a, const GloveWs& w, float* scalars, bool tma, int phases, cudaStream_t stream) {
  int row_blocks = w.row_blocks;
  if (tma) {
    const size_t smem = (size_t)kWarps * kTmaStages * 3 * a.D4 * 16;
    static SmemOptIn configured;  // per NK instantiation: largest dynamic smem opted in
    if (configured.raise(smem))
      ESR_CUDA(cudaFuncSetAttribute(k_glove_rows_tma<NK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(tma_min_blocks(NK), (size_t)(220 * 1024) / (smem + 1024)));
    row_blocks = (int)std::min<int64_t>(w.row_blocks, (int64_t)sm_count() * per_sm);
    if (phases & 1) k_glove_rows_tma<NK><<<row_blocks, kThreads, smem, stream>>>(a);
  } else {
    if (phases & 1) k_glove_rows<NK, S, MINB><<<row_blocks, kThreads, 0, stream>>>(a);
  }
  ESR_LAUNCH_CHECK();
  if (phases & 2) {
    k_glove_combine<NK><<<2 * sm_count(), kCombineThreads, 0, stream>>>(a, row_blocks, scalars);
    ESR_LAUNCH_CHECK();
  }
  return ESR_OK;
}

static int glove_rows_impl(EsrTable* t, const EsrPlan* plan, const EsrGloveCfg* cfg, float* scalars, float* dE, void* ws,
                           size_t ws_bytes, int phases, esr_stream_t stream_) {
  ESR_REQUIRE(cfg_ok(cfg, plan) && glove_table_ok(t, cfg->rows_mode == ESR_ROWS_UPDATE) && scalars != nullptr);
  if (cfg->B == 0) return ESR_OK;
  const bool emit = cfg->rows_mode == ESR_ROWS_EMIT_GRADS;
  ESR_REQUIRE(ws && plan->useg && plan->seg_off);
  ESR_REQUIRE(!emit || cfg->emit_peers_dE != nullptr || (dE != nullptr && (reinterpret_cast<uintptr_t>(dE) % 16) == 0));
  if (cfg->impl != ESR_IMPL_AUTO && cfg->impl != ESR_IMPL_LDG && cfg->impl != ESR_IMPL_TMA) return ESR_EINVAL;
  const bool tma = cfg->impl == ESR_IMPL_TMA;
  if (ws_bytes < esr_glove_workspace_bytes(cfg->B, t->D, cfg->chunk)) return ESR_EWORKSPACE;
  GloveWs w;
  carve_ws(ws, cfg->B, t->D, cfg->chunk, &w);
  const RowsArgs a = make_rows_args(t, plan, cfg, w, scalars, dE);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int nk = (a.D4 + 31) / 32;
  if (cfg->impl == ESR_IMPL_AUTO) {
    // group variant: G lanes x 4 float4 per row (G = D/16 rounded up to a power of two)
    const int d4 = a.D4;
    if (d4 <= 4) return launch_rows_grp<1, 4, 1>(a, w, scalars, phases, stream);
    if (d4 <= 8) return launch_rows_grp<2, 4, 1>(a, w, scalars, phases, stream);
    if (d4 <= 16) return launch_rows_grp<4, 4, 1>(a, w, scalars, phases, stream);
    const bool as = cfg->reserved != 1;  // reserved == 1: keep rows in registers (A/B probe)
    const int go = cfg->row_blocks;
    const bool ff = cfg->reserved == 2;  // reserved == 2: bulk-copy FIFO staging (A/B probe)
    if (d4 <= 32) return launch_rows_grp<8, 4, 1>(a, w, scalars, phases, stream, as && d4 > 16, go, ff);
    if (d4 <= 64) return launch_rows_grp<16, 4, 2>(a, w, scalars, phases, stream, as, go, ff);
    if (d4 <= 96) return launch_rows_grp<32, 3, 3>(a, w, scalars, phases, stream, as, go, ff);
    return launch_rows_grp<32, 4, 4>(a, w, scalars, phases, stream, as, go, ff);
  }
  switch (nk) {
    case 1:
      return launch_rows<1, 2, 4>(a, w, scalars, tma, phases, stream);
    case 2: return launch_rows<2, 2, 1>(a, w, scalars, tma, phases, stream);
    case 3: return launch_rows<3, 1, 1>(a, w, scalars, tma, phases, stream);
    case 4: return launch_rows<4, 1, 1>(a, w, scalars, tma, phases, stream);
    default: return ESR_EINVAL;
  }
}

extern "C" int esr_glove_rows_f32(EsrTable* t, const EsrPlan* plan, const EsrGloveCfg* cfg, float* scalars, float* dE,
                                  void* ws, size_t ws_bytes, esr_stream_t stream) {
  ESR_RANGE("esr_glove_rows_f32");
  return glove_rows_impl(t, plan, cfg, scalars, dE, ws, ws_bytes, 3, stream);
}
extern "C" int esr_glove_rows_main_f32(EsrTable* t, const EsrPlan* plan, const EsrGloveCfg* cfg, float* scalars,
                                       float* dE, void* ws, size_t ws_bytes, esr_stream_t stream) {
  ESR_RANGE("esr_glove_rows_main_f32");
  return glove_rows_impl(t, plan, cfg, scalars, dE, ws, ws_bytes, 1, stream);
}
extern "C" int esr_glove_rows_combine_f32(EsrTable* t, const EsrPlan* plan, const EsrGloveCfg* cfg, float* scalars,
                                          float* dE, void* ws, size_t ws_bytes, esr_stream_t stream) {
  ESR_RANGE("esr_glove_rows_combine_f32");
  return glove_rows_impl(t, plan, cfg, scalars, dE, ws, ws_bytes, 2, stream);
}

extern "C" int esr_glove_finish_f32(EsrTable* t, const EsrPlan* plan, const EsrGloveCfg* cfg, float* scalars, float* db,
                                    void* ws, size_t ws_bytes, esr_stream_t stream_) {
  ESR_RANGE("esr_glove_finish_f32");
  ESR_REQUIRE(cfg_ok(cfg, plan) && glove_table_ok(t, cfg->rows_mode == ESR_ROWS_UPDATE) && scalars != nullptr);
  const bool emit = cfg->rows_mode == ESR_ROWS_EMIT_GRADS;
  if (cfg->B == 0) return ESR_OK;
  ESR_REQUIRE(ws && plan->uniq && plan->seg_off && plan->n_uniq && (!emit || db != nullptr || cfg->emit_peers_db != nullptr));
  if (ws_bytes < esr_glove_workspace_bytes(cfg->B, t->D, cfg->chunk)) return ESR_EWORKSPACE;
  GloveWs w;
  carve_ws(ws, cfg->B, t->D, cfg->chunk, &w);
  FinishArgs a;
  a.uniq = plan->uniq;
  a.seg_off = plan->seg_off;
  a.n_uniq = plan->n_uniq;
  a.bsum = w.bsum;
  a.scalars = scalars;
  a.bias = t->bias;
  a.bias_acc = t->bias_acc;
  a.ver = t->ver;
  a.db = db;
  a.emit_map = cfg->emit_map;
  load_emit_peers(cfg, &a.peers);
  a.per_pair = cfg->bias_mode == ESR_BIAS_PER_PAIR;
  a.emit = emit;
  a.B = (float)cfg->B_global;
  a.lr = cfg->lr;
  a.eps = cfg->eps;
  const bool has_log = cfg->struct_size >= offsetof(EsrGloveCfg, reserved2) && cfg->loss_log && cfg->loss_step && cfg->loss_log_len > 0;
  a.loss_log = has_log ? cfg->loss_log : nullptr;
  a.loss_step = has_log ? cfg->loss_step : nullptr;
  a.loss_log_len = has_log ? cfg->loss_log_len : 0;
  a.loss_host = (has_log && cfg->struct_size >= sizeof(EsrGloveCfg)) ? cfg->loss_host : nullptr;
  k_glove_finish<<<(unsigned)ceil_div(plan->n_slots, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(a);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

extern "C" int esr_glove_step_f32(EsrTable* t, const EsrPlan* plan, const float* counts, const EsrGloveCfg* cfg,
                                  float* scalars, float* dE, float* db, void* ws, size_t ws_bytes, esr_stream_t stream) {
  ESR_RANGE("esr_glove_step_f32");
  int rc = esr_glove_prep_f32(t, plan, counts, cfg, scalars, ws, ws_bytes, stream);
  if (rc != ESR_OK) return rc;
  rc = esr_glove_rows_f32(t, plan, cfg, scalars, dE, ws, ws_bytes, stream);
  if (rc != ESR_OK) return rc;
  return esr_glove_finish_f32(t, plan, cfg, scalars, db, ws, ws_bytes, stream);
}
