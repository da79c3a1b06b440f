// Ranking-loss steps of the Spotify and shop-the-look trainers (reference shapes, CUDA cores).
//
//  esr_stl_triplet_f32      <- STLModel.__call__ scoring (pinterest/models.py:67-72) + train_step loss
//                              (pinterest/train_shop_the_look.py:99-104) + its gradient (:106-107)
//  esr_spotify_fwd_bwd_f32  <- SpotifyModel.get_embeddings / __call__ (spotify/models.py:33-91) +
//                              train_step loss (spotify/train_spotify.py:91-105) + jax.value_and_grad
//                              (:108-109), for a PACK of playlists per launch (one CTA each; the
//                              reference runs one playlist per step and re-jits for every length m)
// Math and VJP conventions (max/min split ties equally, relu'(0) = 0): SURVEY.md App. A.3/A.4;
// CPU restatement: oracle/spotify.py, oracle/stl.py.  Every output element is produced by exactly
// one thread in a fixed order, so results are bit-reproducible.
#include "esr_common.cuh"

namespace esr {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// ---------------------------------------------------------------------------------------------
// Shop-the-look triplet: one warp per batch row.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float reg_term(float nrm) { return fmaxf(nrm - 1.f, 0.f); }

__global__ void __launch_bounds__(kThreads) k_stl_triplet(const float* __restrict__ S, const float* __restrict__ P,
                                                          const float* __restrict__ N, int64_t B, int D, float reg,
                                                          float inv_bs, float* __restrict__ dS, float* __restrict__ dP,
                                                          float* __restrict__ dN, float* __restrict__ pos_score,
                                                          float* __restrict__ neg_score, float* __restrict__ row_loss) {
  const int lane = threadIdx.x & 31;
  const int64_t b = blockIdx.x * (int64_t)kWarps + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* s = S + b * D;
  const float* p = P + b * D;
  const float* n = N + b * D;
  float sp = 0.f, sn = 0.f, ss = 0.f, pp = 0.f, nn = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float x = s[c], y = p[c], z = n[c];
    sp = fmaf(x, y, sp);
    sn = fmaf(x, z, sn);
    ss = fmaf(x, x, ss);
    pp = fmaf(y, y, pp);
    nn = fmaf(z, z, nn);
  }
  sp = warp_sum(sp); sn = warp_sum(sn); ss = warp_sum(ss); pp = warp_sum(pp); nn = warp_sum(nn);
  const float ns = sqrtf(ss), np_ = sqrtf(pp), nn_ = sqrtf(nn);
  const float hinge = 1.f + sn - sp;
  const float a = hinge > 0.f ? 1.f : 0.f;
  // d relu(||e|| - 1) / de = e / ||e|| when ||e|| > 1
  const float rs = (ns - 1.f) > 0.f ? reg / ns : 0.f;
  const float rp = (np_ - 1.f) > 0.f ? reg / np_ : 0.f;
  const float rn = (nn_ - 1.f) > 0.f ? reg / nn_ : 0.f;
  for (int c = lane; c < D; c += 32) {
    const float x = s[c], y = p[c], z = n[c];
    dS[b * D + c] = (a * (z - y) + rs * x) * inv_bs;
    dP[b * D + c] = (-a * x + rp * y) * inv_bs;
    dN[b * D + c] = (a * x + rn * z) * inv_bs;
  }
  if (lane == 0) {
    if (pos_score) pos_score[b] = sp;
    if (neg_score) neg_score[b] = sn;
    row_loss[b] = fmaxf(hinge, 0.f) + reg * (reg_term(ns) + reg_term(np_) + reg_term(nn_));
  }
}

// out[0] = scale * sum(v[0..n)) with a fixed summation tree (single block).
__global__ void __launch_bounds__(1024) k_sum_scaled(const float* __restrict__ v, int64_t n, float scale,
                                                     float* __restrict__ out) {
  __shared__ double sh[1024];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) acc += (double)v[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)(sh[0] * (double)scale);
}

// ---------------------------------------------------------------------------------------------
// Spotify: one CTA per playlist.  Shared memory holds the stacked embeddings X = [ctx; next; neg]
// (row stride FD+1 so a lane-per-row dot is bank-conflict free), the score matrix, and the
// per-row cotangents.
// ---------------------------------------------------------------------------------------------
struct SpotifyArgs {
  const float* A;   // album table  [VA, F]
  const float* R;   // artist table [VR, F]
  int64_t VA;
  int F;
  int nc;           // context rows per playlist (5: spotify/input_pipeline.py:24-26)
  int o;            // negatives per playlist (64: spotify/train_spotify.py:60)
  const int32_t* album_ctx;   // [P*nc] raw ids
  const int32_t* artist_ctx;  // [P*nc]
  const int32_t* next_album;  // [sum m]
  const int32_t* next_artist;
  const int32_t* next_off;    // [P+1]
  const int32_t* neg_album;   // [P*o]
  const int32_t* neg_artist;
  float reg;
  float* loss;         // [P]
  float* dXa;          // [T, F] gradient wrt the album half of every stacked row; T = P*(nc+o) + sum m,
  float* dXr;          // [T, F] artist half; playlist e starts at row e*(nc+o) + next_off[e]
  int32_t* album_rows; // [T] album id mod VA
  int32_t* artist_rows;
  float* pos_aff;      // [sum m] or NULL
  float* neg_aff;      // [P*o] or NULL
  float* l2;           // [T] or NULL
};

constexpr int kMaxColsPerLane = 4;  // 2F <= 128

__global__ void __launch_bounds__(kThreads) k_spotify(const SpotifyArgs a) {
  extern __shared__ float sm[];
  const int e = blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nc = a.nc, o = a.o, F = a.F, FD = 2 * F, LD = FD + 1;
  const int m0 = a.next_off[e], m = a.next_off[e + 1] - m0;
  const int n = nc + m + o;
  const int64_t row_base = (int64_t)e * (nc + o) + m0;
  float* X = sm;                      // [n][LD]
  float* Sc = X + (size_t)n * LD;     // [(m+o)][nc] scores vs context
  float* aff = Sc + (size_t)(m + o) * nc;  // [m+o] affinity incl. boosts
  float* dvec = aff + (m + o);        // [m+o] cotangent of the affinity
  float* mx = dvec + (m + o);         // [m+o] row max of Sc
  float* ties = mx + (m + o);         // [m+o] number of arg-max ties
  float* l2s = ties + (m + o);        // [n]
  float* wsum = l2s + n;              // [kWarps] per-warp loss partials
  __shared__ float stat[8];

  // ---- phase 0: gather [ctx; next; neg] (spotify/models.py:33-46) ----
  for (int r = wid; r < n; r += kWarps) {
    int32_t alb, art;
    if (r < nc) { alb = a.album_ctx[e * nc + r]; art = a.artist_ctx[e * nc + r]; }
    else if (r < nc + m) { alb = a.next_album[m0 + r - nc]; art = a.next_artist[m0 + r - nc]; }
    else { alb = a.neg_album[e * o + r - nc - m]; art = a.neg_artist[e * o + r - nc - m]; }
    const int64_t arow = (int64_t)alb % a.VA;
    float ss = 0.f;
    for (int c = lane; c < FD; c += 32) {
      const float v = c < F ? a.A[arow * F + c] : a.R[(int64_t)art * F + (c - F)];
      X[r * LD + c] = v;
      ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    if (lane == 0) {
      const float nrm = sqrtf(ss);
      l2s[r] = nrm;
      a.album_rows[row_base + r] = (int32_t)arow;
      a.artist_rows[row_base + r] = art;
      if (a.l2) a.l2[row_base + r] = nrm;
    }
  }
  __syncthreads();

  // ---- phase 1: scores of next / neg rows against the context, max + ties + boosts (:74-80) ----
  for (int idx = threadIdx.x; idx < (m + o) * nc; idx += kThreads) {
    const int q = idx / nc, k = idx - q * nc;
    const float* x = X + (nc + q) * LD;
    const float* c = X + k * LD;
    float d = 0.f;
    for (int t = 0; t < FD; ++t) d = fmaf(x[t], c[t], d);
    Sc[idx] = d;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < m + o; q += kThreads) {
    float best = Sc[q * nc];
    for (int k = 1; k < nc; ++k) best = fmaxf(best, Sc[q * nc + k]);
    int nt = 0;
    for (int k = 0; k < nc; ++k) nt += Sc[q * nc + k] == best;
    int32_t alb, art;
    if (q < m) { alb = a.next_album[m0 + q]; art = a.next_artist[m0 + q]; }
    else { alb = a.neg_album[e * o + q - m]; art = a.neg_artist[e * o + q - m]; }
    bool in_alb = false, in_art = false;  // isin on the RAW ids (models.py:75-76)
    for (int k = 0; k < nc; ++k) {
      in_alb |= a.album_ctx[e * nc + k] == alb;
      in_art |= a.artist_ctx[e * nc + k] == art;
    }
    float v = best;
    v = v + 0.1f * (in_alb ? 1.f : 0.f);
    v = v + 0.1f * (in_art ? 1.f : 0.f);
    aff[q] = v;
    mx[q] = best;
    ties[q] = (float)nt;
    if (q < m) { if (a.pos_aff) a.pos_aff[m0 + q] = v; }
    else if (a.neg_aff) a.neg_aff[e * o + q - m] = v;
  }
  __syncthreads();

  // ---- phase 2: batch statistics (train_spotify.py:91-97), warp 0, fixed order ----
  if (wid == 0) {
    float sp = 0.f, sn = 0.f, mn = INFINITY, mxn = -INFINITY;
    for (int q = lane; q < m; q += 32) { sp += aff[q]; mn = fminf(mn, aff[q]); }
    for (int q = m + lane; q < m + o; q += 32) { sn += aff[q]; mxn = fmaxf(mxn, aff[q]); }
    sp = warp_sum(sp);
    sn = warp_sum(sn);
    for (int off = 16; off > 0; off >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(FULL, mn, off));
      mxn = fmaxf(mxn, __shfl_xor_sync(FULL, mxn, off));
    }
    int cmin = 0, cmax = 0;
    for (int q = lane; q < m; q += 32) cmin += aff[q] == mn;
    for (int q = m + lane; q < m + o; q += 32) cmax += aff[q] == mxn;
    for (int off = 16; off > 0; off >>= 1) {
      cmin += __shfl_xor_sync(FULL, cmin, off);
      cmax += __shfl_xor_sync(FULL, cmax, off);
    }
    if (lane == 0) {
      const float mean_trip = 1.f + sn / (float)o - sp / (float)m;
      const float ext_trip = 1.f + mxn - mn;
      stat[0] = mean_trip > 0.f ? 1.f : 0.f;
      stat[1] = ext_trip > 0.f ? 1.f : 0.f;
      stat[2] = mn;
      stat[3] = mxn;
      stat[4] = (float)cmin;
      stat[5] = (float)cmax;
      stat[6] = fmaxf(mean_trip, 0.f) + fmaxf(ext_trip, 0.f);
    }
  }
  __syncthreads();
  {
    const float h1 = stat[0], h2 = stat[1], mn = stat[2], mxn = stat[3], cmin = stat[4], cmax = stat[5];
    for (int q = threadIdx.x; q < m + o; q += kThreads) {
      float d;
      if (q < m) d = -h1 / (float)m - (h2 != 0.f && aff[q] == mn ? 1.f / cmin : 0.f);
      else d = h1 / (float)o + (h2 != 0.f && aff[q] == mxn ? 1.f / cmax : 0.f);
      dvec[q] = d;
    }
  }
  __syncthreads();

  // ---- phase 3: gradient rows + self-affinity / norm losses, one warp per stacked row ----
  float lsum = 0.f;  // this warp's share of the gram hinge sums (already scaled by 1/n_block^2) + reg
  for (int r = wid; r < n; r += kWarps) {
    float acc[kMaxColsPerLane];
#pragma unroll
    for (int k = 0; k < kMaxColsPerLane; ++k) acc[k] = 0.f;
    const float* xr = X + r * LD;
    // (a) max-over-context affinity (ties share the cotangent equally)
    if (r < nc) {
      for (int q = 0; q < m + o; ++q) {
        if (Sc[q * nc + r] == mx[q]) {
          const float w = dvec[q] / ties[q];
          const float* xq = X + (nc + q) * LD;
#pragma unroll
          for (int k = 0; k < kMaxColsPerLane; ++k) {
            const int c = lane + 32 * k;
            if (c < FD) acc[k] = fmaf(w, xq[c], acc[k]);
          }
        }
      }
    } else {
      const int q = r - nc;
      for (int k2 = 0; k2 < nc; ++k2) {
        if (Sc[q * nc + k2] == mx[q]) {
          const float w = dvec[q] / ties[q];
          const float* xc = X + k2 * LD;
#pragma unroll
          for (int k = 0; k < kMaxColsPerLane; ++k) {
            const int c = lane + 32 * k;
            if (c < FD) acc[k] = fmaf(w, xc[c], acc[k]);
          }
        }
      }
    }
    // (b) self-affinity of the row's own block (:85-87, :99-101): G over ALL ordered pairs of the block
    int b0, bn;
    float sign, thr;
    if (r < nc) { b0 = 0; bn = nc; sign = -1.f; thr = 0.5f; }
    else if (r < nc + m) { b0 = nc; bn = m; sign = -1.f; thr = 0.5f; }
    else { b0 = nc + m; bn = o; sign = 1.f; thr = 0.f; }
    const float coef = 2.f * sign / ((float)bn * (float)bn);
    float hsum = 0.f;
    for (int j0 = 0; j0 < bn; j0 += 32) {
      const int j = j0 + lane;
      float g = 0.f;
      bool on = false;
      if (j < bn) {
        const float* xj = X + (b0 + j) * LD;
        for (int t = 0; t < FD; ++t) g = fmaf(xj[t], xr[t], g);
        // ctx / next: relu(0.5 - G); neg: relu(G)
        const float hv = sign < 0.f ? thr - g : g;
        on = hv > 0.f;
        if (on) hsum += hv;
      }
      unsigned mask = __ballot_sync(FULL, on);
      while (mask) {
        const int jj = j0 + __ffs(mask) - 1;
        mask &= mask - 1;
        const float* xj = X + (b0 + jj) * LD;
#pragma unroll
        for (int k = 0; k < kMaxColsPerLane; ++k) {
          const int c = lane + 32 * k;
          if (c < FD) acc[k] = fmaf(coef, xj[c], acc[k]);
        }
      }
    }
    hsum = warp_sum(hsum);
    lsum += hsum / ((float)bn * (float)bn);
    // (c) norm regulariser sum relu(l2 - reg) (:103)
    const float nrm = l2s[r];
    if (nrm - a.reg > 0.f) {
      lsum += nrm - a.reg;
      const float inv = 1.f / nrm;
#pragma unroll
      for (int k = 0; k < kMaxColsPerLane; ++k) {
        const int c = lane + 32 * k;
        if (c < FD) acc[k] = fmaf(xr[c], inv, acc[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < kMaxColsPerLane; ++k) {
      const int c = lane + 32 * k;
      if (c < F) a.dXa[(row_base + r) * F + c] = acc[k];
      else if (c < FD) a.dXr[(row_base + r) * F + (c - F)] = acc[k];
    }
  }
  if (lane == 0) wsum[wid] = lsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = stat[6];
    for (int w = 0; w < kWarps; ++w) tot += wsum[w];
    a.loss[e] = tot;
  }
}

size_t spotify_smem(int nc, int m, int o, int F) {
  const size_t n = (size_t)nc + m + o;
  const size_t q = (size_t)m + o;
  return sizeof(float) * (n * (2 * F + 1) + q * nc + 4 * q + n + kWarps) + 64;
}

}  // namespace
}  // namespace esr

using namespace esr;

extern "C" int esr_stl_triplet_f32(const float* scene, const float* pos, const float* neg, int64_t B, int32_t D,
                                   float regularization, float batch_size, float* d_scene, float* d_pos, float* d_neg,
                                   float* pos_score, float* neg_score, float* loss, float* row_ws, esr_stream_t stream_) {
  ESR_RANGE("esr_stl_triplet_f32");
  ESR_REQUIRE(B >= 0 && D > 0 && batch_size > 0.f && loss != nullptr);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (B == 0) {
    ESR_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), stream));
    return ESR_OK;
  }
  ESR_REQUIRE(scene && pos && neg && d_scene && d_pos && d_neg && row_ws);
  k_stl_triplet<<<(unsigned)ceil_div(B, kWarps), kThreads, 0, stream>>>(scene, pos, neg, B, D, regularization,
                                                                        1.f / batch_size, d_scene, d_pos, d_neg, pos_score,
                                                                        neg_score, row_ws);
  ESR_LAUNCH_CHECK();
  k_sum_scaled<<<1, 1024, 0, stream>>>(row_ws, B, 1.f / batch_size, loss);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

extern "C" int esr_spotify_fwd_bwd_f32(const float* album_table, int64_t VA, const float* artist_table, int32_t F,
                                       int32_t n_playlists, int32_t nc, int32_t o, int32_t max_m,
                                       const int32_t* album_ctx, const int32_t* artist_ctx, const int32_t* next_album,
                                       const int32_t* next_artist, const int32_t* next_off, const int32_t* neg_album,
                                       const int32_t* neg_artist, float regularization, float* loss, float* dXa,
                                       float* dXr, int32_t* album_rows, int32_t* artist_rows, float* pos_aff, float* neg_aff, float* l2,
                                       esr_stream_t stream_) {
  ESR_RANGE("esr_spotify_fwd_bwd_f32");
  ESR_REQUIRE(n_playlists >= 0 && F > 0 && 2 * F <= 32 * kMaxColsPerLane && nc >= 1 && o >= 1 && max_m >= 1 && VA > 0);
  if (n_playlists == 0) return ESR_OK;
  ESR_REQUIRE(album_table && artist_table && album_ctx && artist_ctx && next_album && next_artist && next_off &&
              neg_album && neg_artist && loss && dXa && dXr && album_rows && artist_rows);
  const size_t smem = spotify_smem(nc, max_m, o, F);
  if (smem > 220 * 1024) return ESR_ENOTSUP;  // playlist too long for one CTA's shared memory
  static SmemOptIn configured;
  if (configured.raise(smem)) ESR_CUDA(cudaFuncSetAttribute(k_spotify, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SpotifyArgs a;
  a.A = album_table; a.R = artist_table; a.VA = VA; a.F = F; a.nc = nc; a.o = o;
  a.album_ctx = album_ctx; a.artist_ctx = artist_ctx; a.next_album = next_album; a.next_artist = next_artist;
  a.next_off = next_off; a.neg_album = neg_album; a.neg_artist = neg_artist; a.reg = regularization;
  a.loss = loss; a.dXa = dXa; a.dXr = dXr; a.album_rows = album_rows; a.artist_rows = artist_rows;
  a.pos_aff = pos_aff; a.neg_aff = neg_aff; a.l2 = l2;
  k_spotify<<<n_playlists, kThreads, smem, static_cast<cudaStream_t>(stream_)>>>(a);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}
