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
#include "esr_common.cuh"

namespace esr {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kCombineThreads = 128;
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
  int64_t nchunks;
  int32_t prep_blocks;
  int32_t row_blocks;
  int32_t chunk;
};

int norm_chunk(int32_t chunk) {
  if (chunk <= 0) return 32;
  if (chunk > 32) return 32;
  return (chunk + 3) / 4 * 4;
}

size_t carve_ws(void* base, int64_t B, int32_t D, int32_t chunk_in, GloveWs* w) {
  const int64_t n = 2 * (B > 0 ? B : 1);
  const int chunk = norm_chunk(chunk_in);
  Carver c(base);
  GloveWs t;
  t.chunk = chunk;
  t.nchunks = ceil_div(n, chunk);
  t.prep_blocks = (int32_t)ceil_div(n, kThreads);
  t.row_blocks = (int32_t)ceil_div(t.nchunks, kWarps);
  t.skv = c.take<int32_t>(n);
  t.rec = c.take<SlotRec>(n);
  t.bsum = c.take<float>(n);
  t.part = c.take<float>(t.nchunks * 2 * D);
  t.parts = c.take<float>(t.nchunks * 2);
  t.prep_blk = c.take<float>((size_t)t.prep_blocks * 3);
  t.rows_blk = c.take<float>((size_t)t.row_blocks * 2);
  if (w) *w = t;
  return c.off;
}

// ---------------------------------------------------------------------------------------------
// Phase 1: slot records + batch sums that do not depend on the dots.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_glove_prep(const int32_t* __restrict__ sk, const int32_t* __restrict__ perm,
                                                         const int32_t* __restrict__ partner,
                                                         const float* __restrict__ counts,
                                                         const float* __restrict__ bias, const uint8_t* __restrict__ ver,
                                                         int64_t n, int64_t B, float x_max, float alpha,
                                                         int32_t* __restrict__ skv, SlotRec* __restrict__ rec,
                                                         float* __restrict__ blk) {
  __shared__ float red[32 * 3];
  const int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x;
  float v[3] = {0.f, 0.f, 0.f};
  if (p < n) {
    const int32_t row = sk[p], q = partner[p];
    const int64_t s = perm[p];
    const bool role_i = s < B;
    const float x = counts[role_i ? s : s - B];
    SlotRec r;
    r.w = powf(fminf(1.f, x / x_max), alpha);
    r.t = log10f(1.f + x);
    // b[i] + b[j]: fp add commutes, so both roles of a pair hold the same bits
    r.bs = role_i ? bias[row] + bias[q] : bias[q] + bias[row];
    const int32_t vr = ver ? ver[row] : 0, vq = ver ? ver[q] : 0;
    r.code = (int32_t)((uint32_t)q | (role_i ? (uint32_t)kRoleBit : 0u) | ((uint32_t)vq << 31));
    rec[p] = r;
    skv[p] = (int32_t)((uint32_t)row | ((uint32_t)vr << 31));
    if (role_i) {
      v[0] = r.bs;
      v[1] = r.bs * r.bs;
      v[2] = r.w;
    }
  }
  block_sum<3>(v, red);
  if (threadIdx.x == 0) {
    blk[blockIdx.x * 3 + 0] = v[0];
    blk[blockIdx.x * 3 + 1] = v[1];
    blk[blockIdx.x * 3 + 2] = v[2];
  }
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
    for (int k = 0; k < NV; ++k) acc[k] += (double)blk[b * NV + k];
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

__global__ void __launch_bounds__(kCombineThreads) k_glove_prep_reduce(const float* __restrict__ blk, int nblk,
                                                                       float* __restrict__ scalars) {
  reduce_partials<3>(blk, nblk, scalars + ESR_SC_SUM_BS);
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

struct RowsArgs {
  const float* rows[2];
  float* wrows[2];
  float* acc;
  const int32_t* skv;
  const SlotRec* rec;
  const int32_t* useg;
  const float* scalars;  // [0..2] already global
  float* bsum;
  float* part;
  float* parts;
  float* rows_blk;
  float* dE;
  int64_t n;
  int32_t D4;
  int32_t chunk;
  int32_t per_pair;
  int32_t emit;
  float c2B;      // -2 / B_global
  float inv_B;    // 1 / B_global
  float lr, eps;
};

// Close a finished segment: Adagrad row write (UPDATE) or gradient emit (EMIT).
template <int NK>
__device__ __forceinline__ void close_segment(const RowsArgs& a, int32_t key, int64_t u, const Row<NK>& self,
                                              const Row<NK>& accrow, const Row<NK>& grad, float bacc, int lane) {
  if (a.emit) {
    row_store<NK, true>(grad, reinterpret_cast<float4*>(a.dE) + u * a.D4, lane, a.D4);
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

template <int NK, int S>
__global__ void __launch_bounds__(kThreads) k_glove_rows(const RowsArgs a) {
  __shared__ float red[32 * 2];
  const int lane = threadIdx.x & 31;
  const int64_t c = blockIdx.x * (int64_t)kWarps + (threadIdx.x >> 5);
  const int64_t p0 = c * a.chunk;
  float sums[2] = {0.f, 0.f};  // S1, S2 contributions (lane-uniform)

  if (p0 < a.n) {
    const int cnt = (int)min((int64_t)a.chunk, a.n - p0);
    // keys of the chunk (+ one on each side) -> head / end masks
    const int32_t kl = lane < cnt ? a.skv[p0 + lane] : kNoKey;
    const int32_t kprev0 = p0 > 0 ? a.skv[p0 - 1] : kNoKey;
    const int32_t knextN = p0 + cnt < a.n ? a.skv[p0 + cnt] : kNoKey;
    int32_t kp = __shfl_up_sync(FULL, kl, 1);
    if (lane == 0) kp = kprev0;
    int32_t kn = __shfl_down_sync(FULL, kl, 1);
    if (lane == cnt - 1) kn = knextN;
    const unsigned head_mask = __ballot_sync(FULL, lane < cnt && kl != kp);
    const unsigned end_mask = __ballot_sync(FULL, lane < cnt && kl != kn);
    SlotRec r;
    r.code = 0; r.w = 0.f; r.t = 0.f; r.bs = 0.f;
    if (lane < cnt) r = a.rec[p0 + lane];
    const int64_t u_first = a.useg[p0];
    const float mbs = a.scalars[ESR_SC_SUM_BS] * a.inv_B;

    Row<NK> cur, grad;
    row_zero(cur);
    row_zero(grad);
    float bacc = 0.f;
    bool started_here = false;
    int32_t cur_key = kNoKey;

    for (int sb = 0; sb < cnt; sb += S) {
      Row<NK> P[S], Sf[S], A[S];
      int32_t code[S], key[S];
      float w[S], t[S], bs[S];
      // ---- issue every load of the sub-batch before the first use ----
#pragma unroll
      for (int j = 0; j < S; ++j) {
        const int s = sb + j;
        const int src = s < 32 ? s : 31;
        code[j] = __shfl_sync(FULL, r.code, src);
        w[j] = __shfl_sync(FULL, r.w, src);
        t[j] = __shfl_sync(FULL, r.t, src);
        bs[j] = __shfl_sync(FULL, r.bs, src);
        key[j] = __shfl_sync(FULL, kl, src);
        if (s < cnt) {
          const int64_t q = code[j] & kRowMask;
          const int qv = (code[j] >> 31) & 1;
          row_load<NK, false>(P[j], reinterpret_cast<const float4*>(a.rows[qv]) + q * a.D4, lane, a.D4);
          const bool need_self = (s == 0) || ((head_mask >> s) & 1u);
          if (need_self) {
            const int64_t row = key[j] & kRowMask;
            const int v = (key[j] >> 31) & 1;
            row_load<NK, false>(Sf[j], reinterpret_cast<const float4*>(a.rows[v]) + row * a.D4, lane, a.D4);
          }
          if (!a.emit && ((end_mask >> s) & 1u)) {
            // the accumulator is only used if the segment also started in this chunk; a segment
            // that closes here but started earlier is rare (one per straddling row), so the
            // unconditional prefetch costs nothing measurable
            const int64_t row = key[j] & kRowMask;
            row_load<NK, true>(A[j], reinterpret_cast<const float4*>(a.acc) + row * a.D4, lane, a.D4);
          }
        }
      }
      // ---- consume in slot order ----
#pragma unroll
      for (int j = 0; j < S; ++j) {
        const int s = sb + j;
        if (s < cnt) {
          const bool is_head = (head_mask >> s) & 1u;
          if (s == 0 || is_head) {
            cur = Sf[j];
            cur_key = key[j];
            row_zero(grad);
            bacc = 0.f;
            started_here = is_head;
          }
          float d = 0.f;
#pragma unroll
          for (int k = 0; k < NK; ++k) d += f4_dot(cur.v[k], P[j].v[k]);
          const float dot = warp_sum(d);
          const float res = t[j] - dot;
          const float rr = a.per_pair ? res - bs[j] : res;
          const float g = a.c2B * w[j] * (a.per_pair ? rr : res - mbs);
          if (code[j] & kRoleBit) {
            sums[0] = fmaf(w[j], res, sums[0]);
            sums[1] = fmaf(w[j] * rr, rr, sums[1]);
          }
#pragma unroll
          for (int k = 0; k < NK; ++k) f4_fma(grad.v[k], g, P[j].v[k]);
          bacc += a.per_pair ? g : bs[j];
          const bool is_end = (end_mask >> s) & 1u;
          if (is_end) {
            if (started_here) {
              const int64_t u = u_first + __popc(head_mask & ((2u << s) - 1u) & ~1u);
              close_segment<NK>(a, cur_key, u, cur, A[j], grad, bacc, lane);
            } else {  // leading partial: the segment started in an earlier chunk
              row_store<NK, false>(grad, reinterpret_cast<float4*>(a.part) + (c * 2 + 0) * a.D4, lane, a.D4);
              if (lane == 0) a.parts[c * 2 + 0] = bacc;
            }
          } else if (s == cnt - 1) {  // the chunk ends inside a segment
            const int slot = started_here ? 1 : 0;
            row_store<NK, false>(grad, reinterpret_cast<float4*>(a.part) + (c * 2 + slot) * a.D4, lane, a.D4);
            if (lane == 0) a.parts[c * 2 + slot] = bacc;
          }
        }
      }
    }
  }
  // per-block S1 / S2 partials (lane 0 of each warp carries the warp's value)
  float v[2] = {lane == 0 ? sums[0] : 0.f, lane == 0 ? sums[1] : 0.f};
  block_sum<2>(v, red);
  if (threadIdx.x == 0) {
    a.rows_blk[blockIdx.x * 2 + 0] = v[0];
    a.rows_blk[blockIdx.x * 2 + 1] = v[1];
  }
}

// One block per chunk; the block whose chunk holds the HEAD of a straddling segment adds the
// segment's partials in chunk order (fixed tree => deterministic) and closes the segment.
// The extra last block reduces the S1/S2 block partials into scalars[3..4].
template <int NK>
__global__ void __launch_bounds__(kCombineThreads) k_glove_combine(const RowsArgs a, const int32_t* __restrict__ seg_off,
                                                                   int64_t nchunks, int row_blocks,
                                                                   float* __restrict__ scalars) {
  if (blockIdx.x == nchunks) {
    reduce_partials<2>(a.rows_blk, row_blocks, scalars + ESR_SC_S1);
    return;
  }
  const int64_t c = blockIdx.x;
  const int64_t pl = min((c + 1) * (int64_t)a.chunk, a.n) - 1;
  if (pl >= a.n - 1) return;
  const int32_t key = a.skv[pl];
  if (key != a.skv[pl + 1]) return;
  const int64_t u = a.useg[pl];
  const int64_t s0 = seg_off[u];
  if (s0 < c * (int64_t)a.chunk) return;  // an earlier chunk holds the head
  const int64_t s1 = seg_off[u + 1];
  const int64_t c1 = (s1 - 1) / a.chunk;
  // partial list: index 0 = part[c][1]; index m >= 1 = part[c+m][0]; total np = c1 - c + 1
  const int64_t np = c1 - c + 1;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t per = ceil_div(np, (int64_t)kCombineWarps);
  const int64_t m0 = wid * per, m1 = min(np, m0 + per);
  Row<NK> sum;
  row_zero(sum);
  float ssum = 0.f;
  for (int64_t m = m0; m < m1; ++m) {
    const int64_t idx = m == 0 ? c * 2 + 1 : (c + m) * 2;
    Row<NK> x;
    row_load<NK, true>(x, reinterpret_cast<const float4*>(a.part) + idx * a.D4, lane, a.D4);
#pragma unroll
    for (int k = 0; k < NK; ++k) f4_add(sum.v[k], x.v[k]);
    ssum += a.parts[idx];
  }
  __shared__ float4 sh[kCombineWarps][NK * 32];
  __shared__ float shs[kCombineWarps];
#pragma unroll
  for (int k = 0; k < NK; ++k) sh[wid][k * 32 + lane] = sum.v[k];
  if (lane == 0) shs[wid] = ssum;
  __syncthreads();
  if (wid != 0) return;
  float bacc = shs[0];
#pragma unroll
  for (int w = 1; w < kCombineWarps; ++w) {
#pragma unroll
    for (int k = 0; k < NK; ++k) f4_add(sum.v[k], sh[w][k * 32 + lane]);
    bacc += shs[w];
  }
  Row<NK> self, accrow;
  row_zero(self);
  row_zero(accrow);
  if (!a.emit) {
    const int64_t row = key & kRowMask;
    const int v = (key >> 31) & 1;
    row_load<NK, false>(self, reinterpret_cast<const float4*>(a.rows[v]) + row * a.D4, lane, a.D4);
    row_load<NK, true>(accrow, reinterpret_cast<const float4*>(a.acc) + row * a.D4, lane, a.D4);
  }
  close_segment<NK>(a, key, u, self, accrow, sum, bacc, lane);
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
  int32_t per_pair;
  int32_t emit;
  float B;  // B_global
  float lr, eps;
};

__global__ void __launch_bounds__(kThreads) k_glove_finish(const FinishArgs a) {
  const int64_t u = blockIdx.x * (int64_t)kThreads + threadIdx.x;
  const float S0 = a.scalars[ESR_SC_S0], S1 = a.scalars[ESR_SC_S1];
  if (u < *a.n_uniq) {
    const float nslots = (float)(a.seg_off[u + 1] - a.seg_off[u]);
    // App. A.1: db[v] = sum_slots -(2/B^2) (S1 - bs_r S0) = -(2/B^2) (n_v S1 - S0 sum bs_r); A.2: sum g
    const float gb = a.per_pair ? a.bsum[u] : (-2.f / (a.B * a.B)) * (nslots * S1 - S0 * a.bsum[u]);
    if (a.emit) {
      a.db[u] = gb;
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
  if (cfg->B < 0 || plan->n_slots != 2 * cfg->B || cfg->B_global < cfg->B || (cfg->B > 0 && cfg->B_global <= 0)) return false;
  if (cfg->bias_mode != ESR_BIAS_REFERENCE_BROADCAST && cfg->bias_mode != ESR_BIAS_PER_PAIR) return false;
  if (cfg->rows_mode != ESR_ROWS_UPDATE && cfg->rows_mode != ESR_ROWS_EMIT_GRADS) return false;
  return true;
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
  a.scalars = scalars;
  a.bsum = w.bsum;
  a.part = w.part;
  a.parts = w.parts;
  a.rows_blk = w.rows_blk;
  a.dE = dE;
  a.n = plan->n_slots;
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
                                                       t->ver, n, cfg->B, cfg->x_max, cfg->alpha, w.skv, w.rec, w.prep_blk);
  ESR_LAUNCH_CHECK();
  k_glove_prep_reduce<<<1, kCombineThreads, 0, stream>>>(w.prep_blk, w.prep_blocks, scalars);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

template <int NK, int S>
static int launch_rows(const RowsArgs& a, const GloveWs& w, const EsrPlan* plan, float* scalars, cudaStream_t stream) {
  k_glove_rows<NK, S><<<w.row_blocks, kThreads, 0, stream>>>(a);
  ESR_LAUNCH_CHECK();
  k_glove_combine<NK><<<(unsigned)(w.nchunks + 1), kCombineThreads, 0, stream>>>(a, plan->seg_off, w.nchunks, w.row_blocks,
                                                                               scalars);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

extern "C" int esr_glove_rows_f32(EsrTable* t, const EsrPlan* plan, const EsrGloveCfg* cfg, float* scalars, float* dE,
                                  void* ws, size_t ws_bytes, esr_stream_t stream_) {
  ESR_REQUIRE(cfg_ok(cfg, plan) && glove_table_ok(t, cfg->rows_mode == ESR_ROWS_UPDATE) && scalars != nullptr);
  if (cfg->B == 0) return ESR_OK;
  const bool emit = cfg->rows_mode == ESR_ROWS_EMIT_GRADS;
  ESR_REQUIRE(ws && plan->useg && plan->seg_off);
  ESR_REQUIRE(!emit || (dE != nullptr && (reinterpret_cast<uintptr_t>(dE) % 16) == 0));
  if (cfg->impl != ESR_IMPL_AUTO && cfg->impl != ESR_IMPL_LDG) return ESR_ENOTSUP;
  if (ws_bytes < esr_glove_workspace_bytes(cfg->B, t->D, cfg->chunk)) return ESR_EWORKSPACE;
  GloveWs w;
  carve_ws(ws, cfg->B, t->D, cfg->chunk, &w);
  const RowsArgs a = make_rows_args(t, plan, cfg, w, scalars, dE);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int nk = (a.D4 + 31) / 32;
  switch (nk) {
    case 1: return launch_rows<1, 4>(a, w, plan, scalars, stream);
    case 2: return launch_rows<2, 2>(a, w, plan, scalars, stream);
    case 3: return launch_rows<3, 1>(a, w, plan, scalars, stream);
    case 4: return launch_rows<4, 1>(a, w, plan, scalars, stream);
    default: return ESR_EINVAL;
  }
}

extern "C" int esr_glove_finish_f32(EsrTable* t, const EsrPlan* plan, const EsrGloveCfg* cfg, float* scalars, float* db,
                                    void* ws, size_t ws_bytes, esr_stream_t stream_) {
  ESR_REQUIRE(cfg_ok(cfg, plan) && glove_table_ok(t, cfg->rows_mode == ESR_ROWS_UPDATE) && scalars != nullptr);
  const bool emit = cfg->rows_mode == ESR_ROWS_EMIT_GRADS;
  if (cfg->B == 0) return ESR_OK;
  ESR_REQUIRE(ws && plan->uniq && plan->seg_off && plan->n_uniq && (!emit || db != nullptr));
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
  a.per_pair = cfg->bias_mode == ESR_BIAS_PER_PAIR;
  a.emit = emit;
  a.B = (float)cfg->B_global;
  a.lr = cfg->lr;
  a.eps = cfg->eps;
  k_glove_finish<<<(unsigned)ceil_div(plan->n_slots, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(a);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

extern "C" int esr_glove_step_f32(EsrTable* t, const EsrPlan* plan, const float* counts, const EsrGloveCfg* cfg,
                                  float* scalars, float* dE, float* db, void* ws, size_t ws_bytes, esr_stream_t stream) {
  int rc = esr_glove_prep_f32(t, plan, counts, cfg, scalars, ws, ws_bytes, stream);
  if (rc != ESR_OK) return rc;
  rc = esr_glove_rows_f32(t, plan, cfg, scalars, dE, ws, ws_bytes, stream);
  if (rc != ESR_OK) return rc;
  return esr_glove_finish_f32(t, plan, cfg, scalars, db, ws, ws_bytes, stream);
}
