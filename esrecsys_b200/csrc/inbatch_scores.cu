// In-batch-negative scoring on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
// The one dense contraction of the path (BASELINE.json north star):
//   S = Q K^T  (Bq x Bk, bf16 operands, fp32 accumulate in TMEM), the positive of query i is item
//   i + diag_off, every other item of the batch is a negative.  It generalises the reference's
//   explicit-triplet scoring (pinterest/models.py:67-72 + train_shop_the_look.py:99-104: row-wise dot,
//   hinge(1 + neg - pos)) to B x B in-batch negatives:
//     hinge   : L = (1/Bn) sum_i sum_{j != pos(i)} relu(margin + s S_ij - s S_i,pos(i))
//     softmax : L = (1/Bn) sum_i [ logsumexp_j (s S_ij) - s S_i,pos(i) ]        (no reference oracle, D6)
//   and returns dL/dQ, dL/dK (straight-through to the fp32 inputs).
//
// The fp32 score matrix NEVER reaches HBM: the loss and dL/dS are computed in the epilogue straight
// out of TMEM and only G = dL/dS * Bn / s is stored, as bf16 (hinge: an exact {0,1} mask; softmax:
// the probabilities), Bq x Bk x 2 bytes.  The two backward contractions dQ = G K, dK = G^T Q read G
// back through TMA -- dQ with G as a K-major operand, dK with the SAME buffer as an MN-major operand
// (no transposed copy), the K~ / Q~ operands MN-major as well -- and the diagonal / scale terms are
// applied in fp32 by the finish kernel, so the hinge backward is exact up to fp32 summation order.
//
// Kernels
//   k_inbatch_cast   fp32 -> bf16 (RNE) of Q and K; diag_i = Q~_i . K~_pos(i) in fp32
//   k_inbatch_scores persistent per (i-block, j-range): Q tile resident in smem, K tiles through a
//                    TMA/mbarrier ring, one elected thread issues tcgen05.mma (128x128x16, cta_group::1)
//                    into a double-buffered TMEM accumulator, 8 epilogue warps tcgen05.ld the tile and
//                    apply the loss.  MODE 0 hinge, 1 softmax row statistics, 2 softmax probabilities.
//   k_inbatch_lse    merges the per-j-range (max, sum) pairs into logsumexp_i
//   k_inbatch_bwd    tcgen05 GEMM, one 128 x D tile per CTA (grid.y: dQ | dK, grid.z: split-K)
// The query rows can be processed in CHUNKS (cfg.chunk_rows; automatic above 512 MB of G): the G buffer then
// holds one chunk and is reused, which bounds the workspace for very large batches.  Measured on B200 at
// B = 8192: L2-sized chunks (33 MB) were SLOWER than one pass (166 vs 122 us) -- the backward kernel is bound
// by operand-fetch latency per k-block, not by HBM -- so one pass is the default.
//   k_inbatch_finish fixed-order split-K sum, diagonal terms, scale; loss reduction (deterministic)
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "esr_common.cuh"

namespace esr {
namespace {

constexpr int kTile = 128;                         // score tile edge = UMMA M = UMMA N
constexpr int kBK = 64;                            // bf16 per 128-byte swizzle row = one k-block
constexpr uint32_t kKBlkBytes = kTile * kBK * 2;   // 16 KB: [128 rows][128 B], SWIZZLE_128B, K-major
constexpr uint32_t kAtomBytes = kBK * kBK * 2;     // 8 KB: [64 k-rows][128 B] MN-major atom column
constexpr int kIbThreads = 384;                    // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4..11 epilogue
constexpr int kBwdThreads = 256;                   // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4..7 epilogue
constexpr int kMaxJS = 16;
constexpr int64_t kChunkBytes = 512ll << 20;       // cap of one row chunk of G (workspace bound; chunk_rows overrides)
constexpr int kMaxSplit = 8;
constexpr int kSmCountPlan = 148;                  // B200; only steers the work split

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded spin: a lost arrival traps (launch error the host sees) instead of hanging the GPU.
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
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 2-D tiled TMA load (SASS UTMALDG): box at element coords (x = inner, y = outer) -> smem, completion on bar.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int32_t x, int32_t y, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(x), "r"(y)
               : "memory");
}
// 2-D tiled TMA store (SASS UTMASTG): smem box -> global at element coords (x, y); rows/columns outside the
// tensor's extent are clipped.  Bulk-group completion.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int32_t x, int32_t y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tm)),
               "r"(src), "r"(x), "r"(y)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the smem source of every committed store group has been read (the staging buffer may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

// TMEM allocation (whole warp), result (lane 0 / column base) written to smem.
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, issued by ONE thread for the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive once every tcgen05.mma issued so far by this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets row (lane base + t).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
      "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA instruction descriptor, kind::f16: fp32 accumulate, bf16 A and B (cute::UMMA::InstrDescriptor bit layout).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// UMMA shared-memory matrix descriptor, SWIZZLE_128B, sm_100 version bit (cute::UMMA::SmemDescriptor).
//  K-major operand [rows][64 bf16]: SBO = 1024 B between 8-row groups, LBO unused (one swizzle atom along K).
//  MN-major operand, atom = 64 MN-elements (128 B) x 8 k-rows: SBO = 1024 B between 8-row k groups,
//  LBO = bytes between 64-element MN atoms.
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
         (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ------------------------------------------------------------------------------------------------
// Workspace
// ------------------------------------------------------------------------------------------------
struct IbPlan {
  int Bq, Bk, D, off;
  int n_ib, n_jb;
  int Bc, n_ic, n_chunks;     // row chunk: Bc rows (n_ic i-blocks); G holds ONE chunk and is reused
  int JS, j_per;              // score kernel grid (n_ic, JS); each CTA walks j_per j-blocks
  int R;                      // column ranges the per-row statistics are kept for: 2 * JS (two epilogue halves)
  int Sq, Sk;                 // split-K of dQ (over items) and of dK (over the chunk's queries)
  int64_t ldG;
};

struct IbWs {
  __nv_bfloat16* Qh;
  __nv_bfloat16* Kh;
  __nv_bfloat16* G;
  float* diag;
  float* cnt;      // [R][Bq]  hinge: active negatives of row i inside column range r (exact small integers)
  float2* stats;   // [R][Bq]  softmax: (running max, sum of exp2) of row i inside column range r, log2 domain
  float* lse2;     // [Bq]     log2-domain logsumexp
  float* lossp;    // [n_ib * JS]
  float* partQ;    // [Sq][Bq][D]
  float* partK;    // [n_chunks * Sk][Bk][D]
};

IbPlan make_plan(const EsrInbatchCfg* c) {
  IbPlan p;
  p.Bq = (int)c->Bq;
  p.Bk = (int)c->Bk;
  p.D = c->D;
  p.off = (int)c->diag_off;
  p.n_ib = (int)ceil_div(p.Bq, kTile);
  p.n_jb = (int)ceil_div(p.Bk, kTile);
  p.ldG = (int64_t)p.n_jb * kTile;
  // row chunks: one chunk of G (Bc x ldG bf16) is written by the score kernel and read straight back by
  // the two backward contractions, so it lives in L2; chunk_rows > 0 overrides (>= Bq: a single chunk).
  int64_t bc = c->chunk_rows > 0 ? c->chunk_rows : kChunkBytes / (p.ldG * 2);
  bc = bc / kTile * kTile;
  bc = bc < kTile ? kTile : bc;
  const int64_t nch = ceil_div((int64_t)p.n_ib * kTile, bc);
  p.n_ic = (int)ceil_div(p.n_ib, nch);  // balanced chunks
  p.Bc = p.n_ic * kTile;
  p.n_chunks = (int)ceil_div(p.n_ib, p.n_ic);
  int js = kSmCountPlan / p.n_ic;
  js = js < 1 ? 1 : js;
  js = js > p.n_jb ? p.n_jb : js;
  js = js > kMaxJS ? kMaxJS : js;
  p.j_per = (int)ceil_div(p.n_jb, js);
  p.JS = (int)ceil_div(p.n_jb, p.j_per);
  p.R = 2 * p.JS;
  // backward: dQ has n_ic tiles per chunk with ceil(Bk/64) k-blocks, dK has n_jb tiles with Bc/64 k-blocks.
  // Pick the k-blocks per CTA so that both kinds of CTA run about the same length and one wave fills the SMs.
  const int nkb_q = (int)ceil_div(p.Bk, kBK), nkb_k = (int)ceil_div(p.Bc, kBK);
  int best = nkb_q > nkb_k ? nkb_q : nkb_k;
  for (int t = best; t >= 1; --t) {
    int64_t sq = ceil_div(nkb_q, t), sk = ceil_div(nkb_k, t);
    sq = sq > kMaxSplit ? kMaxSplit : sq;
    sk = sk > kMaxSplit ? kMaxSplit : sk;
    if ((int64_t)p.n_ic * sq + (int64_t)p.n_jb * sk > kSmCountPlan) break;
    best = t;
  }
  p.Sq = (int)ceil_div(nkb_q, best);
  p.Sk = (int)ceil_div(nkb_k, best);
  if (c->splits > 0) p.Sq = p.Sk = c->splits;
  p.Sq = p.Sq > kMaxSplit ? kMaxSplit : p.Sq;
  p.Sk = p.Sk > kMaxSplit ? kMaxSplit : p.Sk;
  return p;
}

size_t carve_ib(void* base, const IbPlan& p, IbWs* w) {
  Carver c(base);
  IbWs t;
  t.Qh = c.take<__nv_bfloat16>((size_t)p.Bq * p.D);
  t.Kh = c.take<__nv_bfloat16>((size_t)p.Bk * p.D);
  t.G = c.take<__nv_bfloat16>((size_t)p.Bc * p.ldG);
  t.diag = c.take<float>(p.Bq);
  t.cnt = c.take<float>((size_t)p.R * p.Bq);
  t.stats = c.take<float2>((size_t)p.R * p.Bq);
  t.lse2 = c.take<float>(p.Bq);
  t.lossp = c.take<float>((size_t)p.n_ib * p.JS);
  t.partQ = c.take<float>((size_t)p.Sq * p.Bq * p.D);
  t.partK = c.take<float>((size_t)p.n_chunks * p.Sk * p.Bk * p.D);
  if (w) *w = t;
  return c.off;
}

// ------------------------------------------------------------------------------------------------
// k_inbatch_cast: one warp per row of Q (and of K): bf16 copies + the positive's score.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_inbatch_cast(const float* __restrict__ Q, const float* __restrict__ K, int Bq, int Bk,
                                                      int D, int off, __nv_bfloat16* __restrict__ Qh,
                                                      __nv_bfloat16* __restrict__ Kh, float* __restrict__ diag) {
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nmax = Bq > Bk ? Bq : Bk;
  if (row >= nmax) return;
  const int D4 = D >> 2;
  if (row < Bk) {
    for (int c = lane; c < D4; c += 32) {
      const float4 v = reinterpret_cast<const float4*>(K + row * D)[c];
      uint2 o;
      o.x = pack_bf16(v.x, v.y);
      o.y = pack_bf16(v.z, v.w);
      reinterpret_cast<uint2*>(Kh + row * D)[c] = o;
    }
  }
  if (row < Bq) {
    const int64_t pj = row + off;
    const bool has_pos = pj >= 0 && pj < Bk;
    float d = 0.f;
    for (int c = lane; c < D4; c += 32) {
      const float4 v = reinterpret_cast<const float4*>(Q + row * D)[c];
      const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&a);
      o.y = *reinterpret_cast<const uint32_t*>(&b);
      reinterpret_cast<uint2*>(Qh + row * D)[c] = o;
      if (has_pos) {
        const float4 k = reinterpret_cast<const float4*>(K + pj * D)[c];
        const __nv_bfloat162 ka = __floats2bfloat162_rn(k.x, k.y), kb = __floats2bfloat162_rn(k.z, k.w);
        d = fmaf(__low2float(a), __low2float(ka), d);
        d = fmaf(__high2float(a), __high2float(ka), d);
        d = fmaf(__low2float(b), __low2float(kb), d);
        d = fmaf(__high2float(b), __high2float(kb), d);
      }
    }
    d = warp_sum(d);
    if (lane == 0) diag[row] = d;
  }
}

// ------------------------------------------------------------------------------------------------
// k_inbatch_scores
//   grid (i-blocks of the chunk, JS column ranges); 384 threads:
//   warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4..11 epilogue -- warp w owns
//   TMEM lanes 32*(w%4).. (one thread per score row) and columns 64*half.. of every tile, half = (w-4)/4.
//   The epilogue is the bound (128x128 elementwise cells per 128x128x128 MMA), so it is written to
//   ~5 instructions per cell with a warp-uniform fast path for tiles that touch neither the diagonal nor
//   the right edge.
// ------------------------------------------------------------------------------------------------
struct ScoreArgs {
  const float* diag;
  __nv_bfloat16* G;   // this chunk's rows: G[(i - i_base) * ldG + j]
  int64_t ldG;
  float* cnt;
  float2* stats;
  const float* lse2;
  float* lossp;
  int Bq, Bk, off, n_jb, j_per, ib0;  // ib0: first i-block of the chunk
  float margin, scale;
  int debug;  // profiling probes (ESR_IB_DEBUG): 1 skip epilogue math, 2 skip TMEM loads, 4 skip MMAs, 8 skip TMA
};

constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int KB, int NS, int MODE>
__global__ void __launch_bounds__(kIbThreads, 1)
    k_inbatch_scores(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmGst, const ScoreArgs a) {
  extern __shared__ unsigned char ib_smem_raw[];
  __shared__ __align__(8) uint64_t bars[1 + 2 * NS + 4];
  __shared__ uint32_t tmem_slot;
  __shared__ float red[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(ib_smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms are 1024-byte aligned
  const uint32_t q_smem = base;
  const uint32_t k_smem = base + KB * kKBlkBytes;
  constexpr uint32_t kStageBytes = KB * kKBlkBytes;
  // per epilogue warp: a [32 rows][128 B] SWIZZLE_128B staging block of its 32 x 64 bf16 piece of G
  const uint32_t stg_smem = k_smem + NS * kStageBytes + (uint32_t)(warp >= 4 ? warp - 4 : 0) * 4096u;
  const uint32_t q_bar = smem_u32(&bars[0]);
  const uint32_t full0 = smem_u32(&bars[1]), empty0 = smem_u32(&bars[1 + NS]);
  const uint32_t tfull0 = smem_u32(&bars[1 + 2 * NS]), tempty0 = smem_u32(&bars[1 + 2 * NS + 2]);
  const int ib = a.ib0 + blockIdx.x, js = blockIdx.y;
  const int jb0 = js * a.j_per;
  const int jb1 = min(a.n_jb, jb0 + a.j_per);
  // Every CTA of a column range walks the same K tiles; started together they would all hit the same L2
  // lines at the same time (measured: 3500 clk per tile, LTS 65 % busy on a 2 MB operand).  Rotate the
  // starting tile per i-block so concurrent CTAs read different tiles.
  const int nt = jb1 - jb0;
  const int rot = nt > 0 ? (int)((blockIdx.x * 5u) % (unsigned)nt) : 0;
#define ESR_JB(it) (jb0 + ((it) + rot >= nt ? (it) + rot - nt : (it) + rot))

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    if (MODE != 1) tma_prefetch_desc(&tmGst);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_bar, 1);
    for (int s = 0; s < NS; ++s) {
      mbar_init(full0 + 8u * s, 1);
      mbar_init(empty0 + 8u * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull0 + 8u * s, 1);
      mbar_init(tempty0 + 8u * s, 8);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(&tmem_slot), 256);  // two 128-column fp32 accumulators
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_expect_tx(q_bar, kStageBytes);
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(q_smem + kb * kKBlkBytes, &tmQ, kb * kBK, ib * kTile, q_bar);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < nt; ++it) {
        const int jb = ESR_JB(it);
        mbar_wait(empty0 + 8u * stage, phase ^ 1u);
        if (a.debug & 8) {
          mbar_arrive(full0 + 8u * stage);
        } else {
          mbar_expect_tx(full0 + 8u * stage, kStageBytes);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb)
            tma_load_2d(k_smem + stage * kStageBytes + kb * kKBlkBytes, &tmK, kb * kBK, jb * kTile, full0 + 8u * stage);
        }
        if (++stage == NS) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kTile, kTile, 0, 0);
      mbar_wait(q_bar, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < nt; ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tempty0 + 8u * acc, acc_phase ^ 1u);
        mbar_wait(full0 + 8u * stage, phase);
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < ((a.debug & 4) ? 0 : KB); ++kb) {
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t ad = make_sdesc(q_smem + kb * kKBlkBytes + k * 32u, 0u, 1024u);
            const uint64_t bd = make_sdesc(k_smem + stage * kStageBytes + kb * kKBlkBytes + k * 32u, 0u, 1024u);
            umma_bf16(tmem_base + (uint32_t)acc * kTile, ad, bd, idesc, (kb | k) ? 1u : 0u);
          }
        }
        umma_commit(empty0 + 8u * stage);  // K stage free once these MMAs have read it
        umma_commit(tfull0 + 8u * acc);    // accumulator ready for the epilogue
        if (++stage == NS) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: thread = one score row (TMEM lane) x one 64-column half of every tile =====
    // TMEM -> register reads run at ~64 B/clk per SM (a 128x128 fp32 tile = 1024 clk, twice its MMA time), so
    // they are software-pipelined: the two 32-column loads of tile t+1 are in flight while tile t is consumed.
    const int q = warp & 3, half = (warp - 4) >> 2;
    const int r = q * 32 + lane;
    const int i = ib * kTile + r;
    const bool rv = i < a.Bq;
    const int jpos = i + a.off;
    const int i_first = ib * kTile + q * 32;  // first row of this warp
    // MODE 0: h = scale*S + c0, c0 = margin - scale*S_ii.  MODE 1/2 work in the log2 domain: t = S * (scale*log2 e).
    const float c0 = (rv && MODE == 0) ? a.margin - a.scale * a.diag[i] : 0.f;
    const float sl2 = a.scale * kLog2e;
    const float nlse2 = (rv && MODE == 2) ? -a.lse2[i] : 0.f;
    float ls0 = 0.f, ls1 = 0.f, cn0 = 0.f, cn1 = 0.f;
    float mx = -INFINITY, sm = 0.f;
    const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 64);

    // one 32-column block of scores (registers v) at columns jbase..jbase+31
    auto consume = [&](const uint32_t (&v)[32], int jbase) {
      if (a.debug & 1) return;
      // warp-uniform: does this 32x32 block touch the right edge, or (hinge) the positives' diagonal?
      const bool edge = jbase + 32 > a.Bk;
      const bool ondiag = MODE == 0 && (i_first + a.off < jbase + 32) && (i_first + 31 + a.off >= jbase);
      if (MODE == 0) {
        uint32_t pk[16];
        if (!edge && !ondiag) {
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float h0 = fmaf(a.scale, __uint_as_float(v[e]), c0), h1 = fmaf(a.scale, __uint_as_float(v[e + 1]), c0);
            ls0 += fmaxf(h0, 0.f);
            ls1 += fmaxf(h1, 0.f);
            // 1.0f where h > 0 else 0.0f (|h| is never in (0, 2^-100)); bf16(1.0) is the upper half of 1.0f
            const float m0 = __saturatef(h0 * 1.2676506e30f), m1 = __saturatef(h1 * 1.2676506e30f);
            cn0 += m0;
            cn1 += m1;
            pk[e >> 1] = __byte_perm(__float_as_uint(m0), __float_as_uint(m1), 0x7632);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            uint32_t bits = 0;
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              const int j = jbase + e + h2;
              const float h = fmaf(a.scale, __uint_as_float(v[e + h2]), c0);
              const bool m = j < a.Bk && j != jpos && h > 0.f;
              ls0 += m ? h : 0.f;
              cn0 += m ? 1.f : 0.f;
              bits |= m ? (0x3F80u << (16 * h2)) : 0u;
            }
            pk[e >> 1] = bits;
          }
        }
        // Direct stores would be 16 B per lane into 32 different rows: 32 LSU wavefronts per instruction
        // (measured: 2000 clk per tile, the whole kernel).  Stage the warp's 32 x 64 block in smem in the
        // TMA 128-byte swizzle (conflict-free st.shared.v4), one TMA store per tile writes it out.
        {
          const int c0 = ((jbase >> 5) & 1) * 4;  // first 16-byte chunk of this 32-column block inside the 128-byte row
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            st_shared_v4(stg_smem + (uint32_t)lane * 128u + (uint32_t)(((c0 + k4) ^ (lane & 7)) * 16),
                         make_uint4(pk[4 * k4], pk[4 * k4 + 1], pk[4 * k4 + 2], pk[4 * k4 + 3]));
        }
      } else if (MODE == 1) {
        float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float t0 = sl2 * __uint_as_float(v[e]), t1 = sl2 * __uint_as_float(v[e + 1]);
          cm0 = (!edge || jbase + e < a.Bk) ? fmaxf(cm0, t0) : cm0;
          cm1 = (!edge || jbase + e + 1 < a.Bk) ? fmaxf(cm1, t1) : cm1;
        }
        const float cm = fmaxf(cm0, cm1);
        if (cm > -INFINITY) {
          const float nm = fmaxf(mx, cm);
          float ad0 = 0.f, ad1 = 0.f;
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float x0 = ex2_approx(fmaf(sl2, __uint_as_float(v[e]), -nm));
            const float x1 = ex2_approx(fmaf(sl2, __uint_as_float(v[e + 1]), -nm));
            ad0 += (!edge || jbase + e < a.Bk) ? x0 : 0.f;
            ad1 += (!edge || jbase + e + 1 < a.Bk) ? x1 : 0.f;
          }
          sm = sm * ex2_approx(mx - nm) + (ad0 + ad1);
          mx = nm;
        }
      } else {
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float p0 = ex2_approx(fmaf(sl2, __uint_as_float(v[e]), nlse2));
          float p1 = ex2_approx(fmaf(sl2, __uint_as_float(v[e + 1]), nlse2));
          if (edge) {
            p0 = jbase + e < a.Bk ? p0 : 0.f;
            p1 = jbase + e + 1 < a.Bk ? p1 : 0.f;
          }
          pk[e >> 1] = pack_bf16(p0, p1);
        }
        // Direct stores would be 16 B per lane into 32 different rows: 32 LSU wavefronts per instruction
        // (measured: 2000 clk per tile, the whole kernel).  Stage the warp's 32 x 64 block in smem in the
        // TMA 128-byte swizzle (conflict-free st.shared.v4), one TMA store per tile writes it out.
        {
          const int c0 = ((jbase >> 5) & 1) * 4;  // first 16-byte chunk of this 32-column block inside the 128-byte row
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            st_shared_v4(stg_smem + (uint32_t)lane * 128u + (uint32_t)(((c0 + k4) ^ (lane & 7)) * 16),
                         make_uint4(pk[4 * k4], pk[4 * k4 + 1], pk[4 * k4 + 2], pk[4 * k4 + 3]));
        }
      }
    };
    // issue the two loads of tile `it` (accumulator it & 1) once the MMA warp has committed it
    auto issue = [&](int it, uint32_t (&va)[32], uint32_t (&vb)[32]) {
      const int acc = it & 1;
      mbar_wait(tfull0 + 8u * acc, (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      if (a.debug & 2) return;
      tmem_ld32(tcol + (uint32_t)(acc * kTile), va);
      tmem_ld32(tcol + (uint32_t)(acc * kTile + 32), vb);
    };
    // the loads of tile `it` have landed: hand the accumulator back to the MMA warp
    auto landed = [&](int it) {
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8u * (it & 1));
    };
    // staging buffer hand-over around the two consume() calls of a tile
    auto stage_begin = [&]() {
      if (MODE == 1 || (a.debug & 1)) return;
      if (lane == 0) tma_store_wait_read();  // the previous tile's store has read the buffer
      __syncwarp();
    };
    auto stage_end = [&](int jb) {
      if (MODE == 1 || (a.debug & 1)) return;
      fence_async_smem();  // generic-proxy writes -> visible to the TMA engine
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&tmGst, stg_smem, jb * kTile + half * 64, (int)blockIdx.x * kTile + q * 32);
        tma_store_commit();
      }
    };
    uint32_t a0[32], a1[32], b0[32], b1[32];
    if (nt > 0) issue(0, a0, a1);
#pragma unroll 1
    for (int it = 0; it < nt; it += 2) {
      landed(it);
      if (it + 1 < nt) issue(it + 1, b0, b1);
      stage_begin();
      consume(a0, ESR_JB(it) * kTile + half * 64);
      consume(a1, ESR_JB(it) * kTile + half * 64 + 32);
      stage_end(ESR_JB(it));
      if (it + 1 < nt) {
        landed(it + 1);
        if (it + 2 < nt) issue(it + 2, a0, a1);
        stage_begin();
        consume(b0, ESR_JB(it + 1) * kTile + half * 64);
        consume(b1, ESR_JB(it + 1) * kTile + half * 64 + 32);
        stage_end(ESR_JB(it + 1));
      }
    }
    if (MODE != 1 && lane == 0) tma_store_wait_all();  // global writes complete before the kernel ends
    const int range = js * 2 + half;
    if (MODE == 0) {
      if (rv) a.cnt[(int64_t)range * a.Bq + i] = cn0 + cn1;
      const float ws = warp_sum(rv ? ls0 + ls1 : 0.f);
      if (lane == 0) red[warp - 4] = ws;
      asm volatile("bar.sync 1, 256;" ::: "memory");  // the eight epilogue warps only
      if (warp == 4 && lane == 0)
        a.lossp[ib * gridDim.y + js] = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
    } else if (MODE == 1) {
      if (rv) a.stats[(int64_t)range * a.Bq + i] = make_float2(mx, sm);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, 256);
#undef ESR_JB
}

// log2-domain logsumexp_i from the per-range statistics (fixed order).
__global__ void __launch_bounds__(256) k_inbatch_lse(const float2* __restrict__ stats, int R, int Bq, float* __restrict__ lse2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Bq) return;
  float M = -INFINITY;
  for (int s = 0; s < R; ++s) M = fmaxf(M, stats[(int64_t)s * Bq + i].x);
  float t = 0.f;
  for (int s = 0; s < R; ++s) {
    const float2 v = stats[(int64_t)s * Bq + i];
    t += v.y * exp2f(v.x - M);
  }
  lse2[i] = M + log2f(t);
}

// ------------------------------------------------------------------------------------------------
// k_inbatch_bwd: C[128 x D] (split-K partial) = A[128 x Kr] * X[Kr x D] for one row chunk of G
//   which == 0 : dQ rows i of the chunk   A = G    K-major  box {64 j, 128 i}     X = K~ rows j  MN-major
//                split sp < Sq over the items j
//   which == 1 : dK rows j (all items)    A = G^T  MN-major 2 boxes {64 j, 64 i}  X = Q~ rows i  MN-major
//                split sp < Sk over the chunk's queries; partial index chunk * Sk + sp
// ------------------------------------------------------------------------------------------------
struct BwdArgs {
  float* partQ;
  float* partK;
  int Bq, Bk, Sq, Sk;
  int i0, rows, chunk;  // chunk = query rows [i0, i0 + rows)
};

template <int NA, int NS>
__global__ void __launch_bounds__(kBwdThreads, 1)
    k_inbatch_bwd(const __grid_constant__ CUtensorMap tmGk, const __grid_constant__ CUtensorMap tmGmn,
                  const __grid_constant__ CUtensorMap tmKb, const __grid_constant__ CUtensorMap tmQb, const BwdArgs a) {
  extern __shared__ unsigned char ib_smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * NS + 1];
  __shared__ uint32_t tmem_slot;
  constexpr int D = NA * 64;
  constexpr uint32_t kCols = D <= 128 ? 128u : 256u;
  constexpr uint32_t kStageBytes = kKBlkBytes + NA * kAtomBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x, which = blockIdx.y, sp = blockIdx.z;
  const int M = which ? a.Bk : a.rows;      // output rows of this GEMM
  const int Kr = which ? a.rows : a.Bk;     // reduction length
  if (mt * kTile >= M || sp >= (which ? a.Sk : a.Sq)) return;  // uniform for the CTA, before any barrier or allocation
  const int nkb = (Kr + kBK - 1) / kBK;
  const int per = (nkb + (which ? a.Sk : a.Sq) - 1) / (which ? a.Sk : a.Sq);
  const int kb0 = sp * per, kb1 = min(nkb, kb0 + per);
  // CTAs of different m-tiles share the X operand: rotate the k-block order per m-tile so they do not
  // all fetch the same L2 lines at the same time (the sum order stays fixed per CTA => deterministic)
  const int nk = kb1 - kb0;
  const int rot = nk > 0 ? (int)((mt * 7u) % (unsigned)nk) : 0;
#define ESR_KB(t) (kb0 + ((t) + rot >= nk ? (t) + rot - nk : (t) + rot))
  const uint32_t base = (smem_u32(ib_smem_raw) + 1023u) & ~1023u;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[NS]), done_bar = smem_u32(&bars[2 * NS]);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(which ? &tmGmn : &tmGk);
    tma_prefetch_desc(which ? &tmQb : &tmKb);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(full0 + 8u * s, 1);
      mbar_init(empty0 + 8u * s, 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(&tmem_slot), kCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < nk; ++t) {
        const int kb = ESR_KB(t);
        mbar_wait(empty0 + 8u * stage, phase ^ 1u);
        const uint32_t fb = full0 + 8u * stage;
        const uint32_t As = base + stage * kStageBytes, Bs = As + kKBlkBytes;
        mbar_expect_tx(fb, kStageBytes);
        if (which == 0) {
          tma_load_2d(As, &tmGk, kb * kBK, mt * kTile, fb);   // G rows are chunk-local
#pragma unroll
          for (int n = 0; n < NA; ++n) tma_load_2d(Bs + n * kAtomBytes, &tmKb, n * 64, kb * kBK, fb);
        } else {
          tma_load_2d(As, &tmGmn, mt * kTile, kb * kBK, fb);
          tma_load_2d(As + kAtomBytes, &tmGmn, mt * kTile + 64, kb * kBK, fb);
#pragma unroll
          for (int n = 0; n < NA; ++n) tma_load_2d(Bs + n * kAtomBytes, &tmQb, n * 64, a.i0 + kb * kBK, fb);
        }
        if (++stage == NS) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && kb0 < kb1) {
      const uint32_t idesc = which ? make_idesc(kTile, D, 1, 1) : make_idesc(kTile, D, 0, 1);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < nk; ++t) {
        mbar_wait(full0 + 8u * stage, phase);
        tc_fence_after();
        const uint32_t As = base + stage * kStageBytes, Bs = As + kKBlkBytes;
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          const uint64_t ad = which ? make_sdesc(As + k * 2048u, kAtomBytes, 1024u) : make_sdesc(As + k * 32u, 0u, 1024u);
          const uint64_t bd = make_sdesc(Bs + k * 2048u, kAtomBytes, 1024u);
          umma_bf16(tmem_base, ad, bd, idesc, (t > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(empty0 + 8u * stage);
        if (++stage == NS) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit(done_bar);
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int row = mt * kTile + q * 32 + lane;  // local output row
    float* out = which ? a.partK + ((int64_t)(a.chunk * a.Sk + sp) * a.Bk + row) * D
                       : a.partQ + ((int64_t)sp * a.Bq + a.i0 + row) * D;
    const bool has = kb0 < kb1;
    if (has) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int ch = 0; ch < D / 32; ++ch) {
      uint32_t v[32];
      if (has) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = 0u;
      }
      if (row < M) {
        float4* dst = reinterpret_cast<float4*>(out + ch * 32);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          dst[e] = make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]),
                               __uint_as_float(v[4 * e + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, kCols);
#undef ESR_KB
}

// ------------------------------------------------------------------------------------------------
// k_inbatch_finish
// ------------------------------------------------------------------------------------------------
struct FinishIbArgs {
  const float* partQ;
  const float* partK;
  const __nv_bfloat16* Qh;
  const __nv_bfloat16* Kh;
  const float* cnt;
  const float* lossp;
  const float* lse2;
  const float* diag;
  float* dQ;
  float* dK;
  float* loss;
  int Bq, Bk, D, off, Sq, SkAll, R, nlossp, softmax;
  float coef;    // scale / B_norm
  float inv_bn;  // 1 / B_norm
  float scale;
};

__device__ __forceinline__ float4 bf16x4_to_f4(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x), b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}

__global__ void __launch_bounds__(256) k_inbatch_finish(const FinishIbArgs a) {
  const int D4 = a.D >> 2;
  const int64_t nQ = (int64_t)a.Bq * D4, nK = (int64_t)a.Bk * D4;
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx < nQ + nK) {
    const bool isK = idx >= nQ;
    const int64_t e = isK ? idx - nQ : idx;
    const int64_t row = e / D4;
    const int c4 = (int)(e % D4);
    const int M = isK ? a.Bk : a.Bq;
    const int S = isK ? a.SkAll : a.Sq;
    const float* part = isK ? a.partK : a.partQ;
    float4 s = f4_zero();
    for (int sp = 0; sp < S; ++sp) f4_add(s, __ldcs(reinterpret_cast<const float4*>(part + ((int64_t)sp * M + row) * a.D) + c4));
    // diagonal term: query i and its positive item pos(i) = i + off
    const int64_t i = isK ? row - a.off : row;
    const int64_t other = isK ? i : row + a.off;  // row of the OTHER matrix
    const bool has = isK ? (i >= 0 && i < a.Bq) : (other >= 0 && other < a.Bk);
    if (has) {
      float w = 1.f;
      if (!a.softmax) {
        w = 0.f;
        for (int r = 0; r < a.R; ++r) w += a.cnt[(int64_t)r * a.Bq + i];
      }
      const float4 o = bf16x4_to_f4((isK ? a.Qh : a.Kh) + other * a.D + 4 * c4);
      s.x = fmaf(-w, o.x, s.x);
      s.y = fmaf(-w, o.y, s.y);
      s.z = fmaf(-w, o.z, s.z);
      s.w = fmaf(-w, o.w, s.w);
    }
    s.x *= a.coef; s.y *= a.coef; s.z *= a.coef; s.w *= a.coef;
    reinterpret_cast<float4*>((isK ? a.dK : a.dQ) + row * a.D)[c4] = s;
  }
  if (blockIdx.x == 0) {  // loss: fixed-order double accumulation
    __shared__ double sh[256];
    double acc = 0.0;
    if (a.softmax) {
      for (int i = threadIdx.x; i < a.Bq; i += blockDim.x) {
        const int64_t pj = (int64_t)i + a.off;
        const float pos = (pj >= 0 && pj < a.Bk) ? a.scale * a.diag[i] : 0.f;
        acc += (double)(a.lse2[i] * 0.6931471805599453f - pos);
      }
    } else {
      for (int i = threadIdx.x; i < a.nlossp; i += blockDim.x) acc += (double)a.lossp[i];
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) *a.loss = (float)(sh[0] * (double)a.inv_bn);
  }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled encode_fn() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
  }
  return fn;
}

// bf16 row-major [outer][ld] tensor, logical extent {inner, outer}; box {bi, bo}; SWIZZLE_128B; OOB reads as zero.
bool make_tmap(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t bi, uint32_t bo) {
  PFN_cuTensorMapEncodeTiled fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {inner, outer};
  const cuuint64_t strides[1] = {ld * 2};
  const cuuint32_t box[2] = {bi, bo};
  const cuuint32_t es[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool ib_cfg_ok(const EsrInbatchCfg* c) {
  if (!c || c->struct_size < sizeof(EsrInbatchCfg)) return false;
  if (c->Bq <= 0 || c->Bk <= 0 || c->Bq > (1 << 20) || c->Bk > (1 << 20)) return false;
  if (c->D != 64 && c->D != 128 && c->D != 192 && c->D != 256) return false;
  if (c->loss_kind != ESR_LOSS_HINGE && c->loss_kind != ESR_LOSS_SOFTMAX) return false;
  if (!(c->b_norm > 0.f) || !(c->scale > 0.f) || c->chunk_rows < 0 || c->splits < 0) return false;
  return true;
}

template <typename F>
int set_smem(F* fn, size_t bytes) {
  ESR_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return ESR_OK;
}

template <int KB, int NS>
int launch_scores(int mode, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tg, const ScoreArgs& sa, dim3 grid,
                  cudaStream_t st) {
  const size_t smem = (size_t)(1 + NS) * KB * kKBlkBytes + 8 * 4096 + 1024;
  int rc;
  if (mode == 0) {
    if ((rc = set_smem(k_inbatch_scores<KB, NS, 0>, smem)) != ESR_OK) return rc;
    k_inbatch_scores<KB, NS, 0><<<grid, kIbThreads, smem, st>>>(tq, tk, tg, sa);
  } else if (mode == 1) {
    if ((rc = set_smem(k_inbatch_scores<KB, NS, 1>, smem)) != ESR_OK) return rc;
    k_inbatch_scores<KB, NS, 1><<<grid, kIbThreads, smem, st>>>(tq, tk, tg, sa);
  } else {
    if ((rc = set_smem(k_inbatch_scores<KB, NS, 2>, smem)) != ESR_OK) return rc;
    k_inbatch_scores<KB, NS, 2><<<grid, kIbThreads, smem, st>>>(tq, tk, tg, sa);
  }
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

int launch_scores_d(int D, int mode, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tg, const ScoreArgs& sa,
                    dim3 grid, cudaStream_t st) {
  switch (D) {
    case 64: return launch_scores<1, 4>(mode, tq, tk, tg, sa, grid, st);
    case 128: return launch_scores<2, 4>(mode, tq, tk, tg, sa, grid, st);
    case 192: return launch_scores<3, 3>(mode, tq, tk, tg, sa, grid, st);
    case 256: return launch_scores<4, 2>(mode, tq, tk, tg, sa, grid, st);
    default: return ESR_EINVAL;
  }
}

template <int NA, int NS>
int launch_bwd(const CUtensorMap& gk, const CUtensorMap& gmn, const CUtensorMap& kb, const CUtensorMap& qb, const BwdArgs& ba,
               dim3 grid, cudaStream_t st) {
  const size_t smem = (size_t)NS * (kKBlkBytes + NA * kAtomBytes) + 1024;
  const int rc = set_smem(k_inbatch_bwd<NA, NS>, smem);
  if (rc != ESR_OK) return rc;
  k_inbatch_bwd<NA, NS><<<grid, kBwdThreads, smem, st>>>(gk, gmn, kb, qb, ba);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

}  // namespace
}  // namespace esr

using namespace esr;

extern "C" size_t esr_inbatch_workspace_bytes(const EsrInbatchCfg* cfg) {
  if (!ib_cfg_ok(cfg)) return 0;
  return carve_ib(nullptr, make_plan(cfg), nullptr) + 256;
}

extern "C" int esr_inbatch_ws_layout(const EsrInbatchCfg* cfg, int64_t* out /* [10] */) {
  ESR_REQUIRE(ib_cfg_ok(cfg) && out);
  const IbPlan p = make_plan(cfg);
  IbWs w;
  carve_ib(nullptr, p, &w);
  out[0] = (int64_t)reinterpret_cast<uintptr_t>(w.G);     // byte offset of G (bf16 [Bc][ldG], LAST chunk processed)
  out[1] = p.ldG;
  out[2] = (int64_t)reinterpret_cast<uintptr_t>(w.diag);  // float [Bq]
  out[3] = (int64_t)reinterpret_cast<uintptr_t>(w.cnt);   // float [R][Bq]
  out[4] = p.R;
  out[5] = (int64_t)reinterpret_cast<uintptr_t>(w.lse2);  // float [Bq], log2 domain
  out[6] = p.n_chunks;
  out[7] = (int64_t)reinterpret_cast<uintptr_t>(w.Qh);    // bf16 [Bq][D]
  out[8] = p.Bc;
  out[9] = p.Sq * 100 + p.Sk;
  return ESR_OK;
}

extern "C" int esr_inbatch_fwd_bwd_bf16(const float* Q, const float* K, const EsrInbatchCfg* cfg, float* dQ, float* dK,
                                        float* loss, void* ws, size_t ws_bytes, esr_stream_t stream_) {
  ESR_RANGE("esr_inbatch_fwd_bwd_bf16");
  ESR_REQUIRE(ib_cfg_ok(cfg) && Q && K && dQ && dK && loss && ws);
  ESR_REQUIRE((reinterpret_cast<uintptr_t>(Q) % 16) == 0 && (reinterpret_cast<uintptr_t>(K) % 16) == 0 &&
              (reinterpret_cast<uintptr_t>(dQ) % 16) == 0 && (reinterpret_cast<uintptr_t>(dK) % 16) == 0 &&
              (reinterpret_cast<uintptr_t>(ws) % 256) == 0);
  if (ws_bytes < esr_inbatch_workspace_bytes(cfg)) return ESR_EWORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const IbPlan p = make_plan(cfg);
  IbWs w;
  carve_ib(ws, p, &w);
  const int softmax = cfg->loss_kind == ESR_LOSS_SOFTMAX;

  CUtensorMap tmQ, tmK, tmKb, tmQb;
  if (!(make_tmap(&tmQ, w.Qh, p.D, p.Bq, p.D, kBK, kTile) && make_tmap(&tmK, w.Kh, p.D, p.Bk, p.D, kBK, kTile) &&
        make_tmap(&tmKb, w.Kh, p.D, p.Bk, p.D, 64, kBK) && make_tmap(&tmQb, w.Qh, p.D, p.Bq, p.D, 64, kBK)))
    return ESR_ENOTSUP;

  const int nmax = p.Bq > p.Bk ? p.Bq : p.Bk;
  k_inbatch_cast<<<(unsigned)ceil_div(nmax, 8), 256, 0, st>>>(Q, K, p.Bq, p.Bk, p.D, p.off, w.Qh, w.Kh, w.diag);
  ESR_LAUNCH_CHECK();

  ScoreArgs sa;
  sa.diag = w.diag;
  sa.G = w.G;
  sa.ldG = p.ldG;
  sa.cnt = w.cnt;
  sa.stats = w.stats;
  sa.lse2 = w.lse2;
  sa.lossp = w.lossp;
  sa.Bq = p.Bq;
  sa.Bk = p.Bk;
  sa.off = p.off;
  sa.n_jb = p.n_jb;
  sa.j_per = p.j_per;
  sa.ib0 = 0;
  sa.margin = cfg->margin;
  sa.scale = cfg->scale;
  {
    const char* dbg = getenv("ESR_IB_DEBUG");
    sa.debug = dbg ? atoi(dbg) : 0;
  }
  int rc;
  if (softmax) {  // row statistics of the whole batch first (no G traffic), then logsumexp
    if ((rc = launch_scores_d(p.D, 1, tmQ, tmK, tmQ /* unused in this mode */, sa, dim3(p.n_ib, p.JS), st)) != ESR_OK) return rc;
    k_inbatch_lse<<<(unsigned)ceil_div(p.Bq, 256), 256, 0, st>>>(w.stats, p.R, p.Bq, w.lse2);
    ESR_LAUNCH_CHECK();
  }

  BwdArgs ba;
  ba.partQ = w.partQ;
  ba.partK = w.partK;
  ba.Bq = p.Bq;
  ba.Bk = p.Bk;
  ba.Sq = p.Sq;
  ba.Sk = p.Sk;
  for (int c = 0; c < p.n_chunks; ++c) {
    const int ib0 = c * p.n_ic;
    const int n_ic = p.n_ib - ib0 < p.n_ic ? p.n_ib - ib0 : p.n_ic;
    const int i0 = ib0 * kTile;
    const int rows = p.Bq - i0 < n_ic * kTile ? p.Bq - i0 : n_ic * kTile;
    sa.ib0 = ib0;
    CUtensorMap tmGk, tmGmn, tmGst;  // this chunk's rows only: rows beyond `rows` read as zero / are not written
    if (!(make_tmap(&tmGk, w.G, p.Bk, rows, p.ldG, kBK, kTile) && make_tmap(&tmGmn, w.G, p.Bk, rows, p.ldG, 64, kBK) &&
          make_tmap(&tmGst, w.G, p.Bk, rows, p.ldG, 64, 32)))
      return ESR_ENOTSUP;
    if ((rc = launch_scores_d(p.D, softmax ? 2 : 0, tmQ, tmK, tmGst, sa, dim3(n_ic, p.JS), st)) != ESR_OK) return rc;
    ba.i0 = i0;
    ba.rows = rows;
    ba.chunk = c;
    const dim3 bgrid(n_ic > p.n_jb ? n_ic : p.n_jb, 2, p.Sq > p.Sk ? p.Sq : p.Sk);
    switch (p.D) {
      case 64: rc = launch_bwd<1, 6>(tmGk, tmGmn, tmKb, tmQb, ba, bgrid, st); break;
      case 128: rc = launch_bwd<2, 6>(tmGk, tmGmn, tmKb, tmQb, ba, bgrid, st); break;
      case 192: rc = launch_bwd<3, 4>(tmGk, tmGmn, tmKb, tmQb, ba, bgrid, st); break;
      default: rc = launch_bwd<4, 4>(tmGk, tmGmn, tmKb, tmQb, ba, bgrid, st); break;
    }
    if (rc != ESR_OK) return rc;
  }

  FinishIbArgs fa;
  fa.partQ = w.partQ;
  fa.partK = w.partK;
  fa.Qh = w.Qh;
  fa.Kh = w.Kh;
  fa.cnt = w.cnt;
  fa.lossp = w.lossp;
  fa.lse2 = w.lse2;
  fa.diag = w.diag;
  fa.dQ = dQ;
  fa.dK = dK;
  fa.loss = loss;
  fa.Bq = p.Bq;
  fa.Bk = p.Bk;
  fa.D = p.D;
  fa.off = p.off;
  fa.Sq = p.Sq;
  fa.SkAll = p.n_chunks * p.Sk;
  fa.R = p.R;
  fa.nlossp = p.n_ib * p.JS;
  fa.softmax = softmax;
  fa.coef = cfg->scale / cfg->b_norm;
  fa.inv_bn = 1.f / cfg->b_norm;
  fa.scale = cfg->scale;
  const int64_t nel = ((int64_t)p.Bq + p.Bk) * (p.D / 4);
  k_inbatch_finish<<<(unsigned)ceil_div(nel, 256), 256, 0, st>>>(fa);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}
