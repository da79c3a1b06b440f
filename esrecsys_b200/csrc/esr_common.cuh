// Shared device/host helpers for libesr (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>
#include <stdio.h>

#include "esr.h"

namespace esr {

// ---------------------------------------------------------------------------------------------
// Error plumbing: no exception crosses the C ABI; the CUDA error string is kept per thread.
// ---------------------------------------------------------------------------------------------
void set_cuda_error(cudaError_t e, const char* what, const char* file, int line);

#define ESR_CUDA(call)                                              \
  do {                                                              \
    cudaError_t e__ = (call);                                       \
    if (e__ != cudaSuccess) {                                       \
      ::esr::set_cuda_error(e__, #call, __FILE__, __LINE__);        \
      return ESR_ECUDA;                                             \
    }                                                               \
  } while (0)

#define ESR_LAUNCH_CHECK() ESR_CUDA(cudaPeekAtLastError())

#define ESR_REQUIRE(cond) \
  do {                    \
    if (!(cond)) return ESR_EINVAL; \
  } while (0)

int sm_count();  // of the current device (cached per device)

// NVTX range over a C-ABI entry point (header-only NVTX 3: a no-op unless a profiler injects itself), so that a
// timeline shows the phases of a step -- plan / prep / rows / finish, route / gather / merge -- by name.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
#define ESR_RANGE(name) ::esr::NvtxRange esr_nvtx_range__(name)

// Largest dynamic shared-memory opt-in already made for ONE kernel, per device (cudaFuncSetAttribute is a per-device
// setting; a process normally drives one GPU, but nothing here relies on it):
//   static SmemOptIn seen;  if (smem > 48 * 1024 && seen.raise(smem)) cudaFuncSetAttribute(kernel, ..., smem);
struct SmemOptIn {
  static constexpr int kMaxDevices = 64;
  size_t v[kMaxDevices] = {};
  bool raise(size_t smem) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return true;  // unknown: always opt in
    if (smem <= v[dev]) return false;
    v[dev] = smem;
    return true;
  }
};

#ifdef __CUDACC__
#define ESR_HD __host__ __device__
#else
#define ESR_HD
#endif
ESR_HD static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
ESR_HD static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Carves a caller-provided workspace into 256-byte aligned pieces.
struct Carver {
  char* base;
  size_t off;
  explicit Carver(void* p) : base(static_cast<char*>(p)), off(0) {}
  template <typename T>
  T* take(size_t n) {
    T* p = reinterpret_cast<T*>(base + off);
    off += align_up(n * sizeof(T), 256);
    return p;
  }
};

// ---------------------------------------------------------------------------------------------
// Device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

constexpr unsigned FULL = 0xffffffffu;

// Sum over a power-of-two group of G lanes (xor butterfly: every lane gets the total, and the
// summation tree is fixed, so the result is bit-reproducible).
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

__device__ __forceinline__ float warp_sum(float v) { return group_sum<32>(v); }

// Block-wide sum of NV values for blockDim.x <= 1024 (fixed tree => deterministic). Result valid in thread 0.
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* smem /* [32*NV] */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) smem[wid * NV + k] = v[k];
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      float x = lane < nw ? smem[lane * NV + k] : 0.f;
      v[k] = warp_sum(x);
    }
  }
  __syncthreads();
}

// 128-bit loads/stores with cache policy.
//  ld_keep   : table rows that other work items of the same step will re-read (default caching).
//  ld_stream : read-once data (accumulator rows, partials): evict-first in L2, not kept in L1.
//  st_stream : written-once data (new rows, accumulator): evict-first.
__device__ __forceinline__ float4 ld_keep(const float4* p) { return __ldg(p); }
__device__ __forceinline__ float4 ld_stream(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float4* p, float4 v) { __stcs(p, v); }

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float f4_dot(float4 a, float4 b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ void f4_fma(float4& acc, float s, float4 v) {
  acc.x = fmaf(s, v.x, acc.x);
  acc.y = fmaf(s, v.y, acc.y);
  acc.z = fmaf(s, v.z, acc.z);
  acc.w = fmaf(s, v.w, acc.w);
}
__device__ __forceinline__ void f4_add(float4& acc, float4 v) {
  acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
}

// Packed fp32 pairs (Blackwell FFMA2 / FADD2 / FMUL2: fma.rn.f32x2 ...): each half is an ordinary IEEE op, so a packed
// sequence is bit-identical to the same scalar sequence; it halves the issue slots of the row pass, which is
// issue-bound on duplicate-heavy streams.
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {  // d += a * b
  asm("{\n .reg .b64 ra, rb, rc;\n mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n mov.b64 rc, {%0, %1};\n"
      " fma.rn.f32x2 rc, ra, rb, rc;\n mov.b64 {%0, %1}, rc;\n}"
      : "+f"(d0), "+f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fmul2(float& d0, float& d1, float a0, float a1, float b0, float b1) {  // d = a * b
  asm("{\n .reg .b64 ra, rb, rc;\n mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n mul.rn.f32x2 rc, ra, rb;\n mov.b64 {%0, %1}, rc;\n}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1, float b0, float b1) {  // d = a + b
  asm("{\n .reg .b64 ra, rb, rc;\n mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n add.rn.f32x2 rc, ra, rb;\n mov.b64 {%0, %1}, rc;\n}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
// dot of two float4 pairs into a 2-lane accumulator (acc0 gets x,z products, acc1 gets y,w)
__device__ __forceinline__ void f4_dot2(float& acc0, float& acc1, float4 a, float4 b) {
  ffma2(acc0, acc1, a.x, a.y, b.x, b.y);
  ffma2(acc0, acc1, a.z, a.w, b.z, b.w);
}
__device__ __forceinline__ void f4_fma2(float4& acc, float s, float4 v) {  // same bits as f4_fma
  ffma2(acc.x, acc.y, s, s, v.x, v.y);
  ffma2(acc.z, acc.w, s, s, v.z, v.w);
}
// optax.adagrad on two elements; same bits as two adagrad1 calls when eps > 0 (a == 0 then implies g == 0 and the
// update is exactly 0 either way, so the a > 0 guard of adagrad1 is only needed for eps == 0)
__device__ __forceinline__ void adagrad2(float& p0, float& p1, float& a0, float& a1, float g0, float g1, float lr, float eps) {
  ffma2(a0, a1, g0, g1, g0, g1);
  float t0, t1;
  fadd2(t0, t1, a0, a1, eps, eps);
  const float i0 = rsqrtf(t0), i1 = rsqrtf(t1);
  float s0, s1;
  fmul2(s0, s1, -lr, -lr, g0, g1);
  ffma2(p0, p1, s0, s1, i0, i1);
}
__device__ __forceinline__ void adagrad4_packed(float4& p, float4& a, float4 g, float lr, float eps) {
  adagrad2(p.x, p.y, a.x, a.y, g.x, g.y, lr, eps);
  adagrad2(p.z, p.w, a.z, a.w, g.z, g.w, lr, eps);
}

// optax.adagrad: a += g^2 ; p -= lr * g * rsqrt(a + eps)   (0 where a == 0)
__device__ __forceinline__ void adagrad1(float& p, float& a, float g, float lr, float eps) {
  a = fmaf(g, g, a);
  const float inv = a > 0.f ? rsqrtf(a + eps) : 0.f;
  p = fmaf(-lr * g, inv, p);
}
__device__ __forceinline__ void adagrad4(float4& p, float4& a, float4 g, float lr, float eps) {
  adagrad1(p.x, a.x, g.x, lr, eps);
  adagrad1(p.y, a.y, g.y, lr, eps);
  adagrad1(p.z, a.z, g.z, lr, eps);
  adagrad1(p.w, a.w, g.w, lr, eps);
}

#endif  // __CUDACC__

}  // namespace esr
