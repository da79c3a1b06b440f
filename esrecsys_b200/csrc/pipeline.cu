// Native host runtime of the single-GPU training loop: the body of train_epoch (wikipedia/train_cooccurence.py:103-112:
// next(train_it) -> apply_model -> update_model) as ONE C call per step.
//
// A step is three stream-ordered stages on three streams owned by this object
//   copy stream : stage the batch (H2D from pinned host memory on a copy engine, or D2D)
//   side stream : index plan of the batch (ids only)                       -- a captured CUDA graph per buffer parity
//   main stream : prep -> rows -> combine -> finish on the table           -- a captured CUDA graph per buffer parity
// with `depth` staging / plan buffers in flight (3: batch t+2 uploads while plan(t+1) is built and step(t) trains).
// Round 1 drove this from Python through torch (stream context managers, tensor.copy_, torch events, CUDAGraph.replay):
// ~14 framework calls, 170-200 us of host time per step -- more than the 157 us the GPU needs, so the end-to-end rate
// from host batches was HOST-bound (1.35 G pairs/s against 1.67 G device-timed).  Here a step is ~12 CUDA runtime calls.
// The kernels are captured while the caller launches them through the usual libesr entry points on these streams
// (esr_pipeline_capture_begin / _end bracket the launches), so no kernel knowledge lives here.
#include <new>

#include "esr_common.cuh"

namespace {
constexpr int kMaxDepth = 4;
constexpr uint32_t kMagic = 0x45535250u;  // "ESRP"
}  // namespace

struct EsrPipeline {
  uint32_t magic;
  int depth;
  cudaStream_t copy, side, main;
  cudaEvent_t ev_in, ev_copy[kMaxDepth], ev_plan[kMaxDepth], ev_done[kMaxDepth];
  cudaGraphExec_t g_plan[kMaxDepth], g_step[kMaxDepth];
  void* ids_dev[kMaxDepth];
  void* counts_dev[kMaxDepth];
  size_t ids_bytes, counts_bytes;
  const float* loss_src;
  float* loss_log;
  float* loss_host;
  int64_t loss_len;
  int64_t t;
  int capturing;  // -1 none, 0 plan, 1 step
};

using namespace esr;

static bool ok(const EsrPipeline* p) { return p != nullptr && p->magic == kMagic; }

extern "C" int esr_pipeline_create(int32_t depth, int32_t main_high_priority, EsrPipeline** out) {
  ESR_REQUIRE(out != nullptr && depth >= 1 && depth <= kMaxDepth);
  EsrPipeline* p = new (std::nothrow) EsrPipeline();
  if (p == nullptr) return ESR_ENOMEM;
  p->magic = kMagic;
  p->depth = depth;
  p->capturing = -1;
  int lo = 0, hi = 0;
  ESR_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  ESR_CUDA(cudaStreamCreateWithPriority(&p->copy, cudaStreamNonBlocking, lo));
  ESR_CUDA(cudaStreamCreateWithPriority(&p->side, cudaStreamNonBlocking, lo));
  ESR_CUDA(cudaStreamCreateWithPriority(&p->main, cudaStreamNonBlocking, main_high_priority ? hi : lo));
  ESR_CUDA(cudaEventCreateWithFlags(&p->ev_in, cudaEventDisableTiming));
  for (int k = 0; k < kMaxDepth; ++k) {
    ESR_CUDA(cudaEventCreateWithFlags(&p->ev_copy[k], cudaEventDisableTiming));
    ESR_CUDA(cudaEventCreateWithFlags(&p->ev_plan[k], cudaEventDisableTiming));
    ESR_CUDA(cudaEventCreateWithFlags(&p->ev_done[k], cudaEventDisableTiming));
    p->g_plan[k] = p->g_step[k] = nullptr;
    p->ids_dev[k] = p->counts_dev[k] = nullptr;
  }
  *out = p;
  return ESR_OK;
}

extern "C" int esr_pipeline_streams(const EsrPipeline* p, esr_stream_t* copy, esr_stream_t* side, esr_stream_t* main_) {
  ESR_REQUIRE(ok(p));
  if (copy) *copy = p->copy;
  if (side) *side = p->side;
  if (main_) *main_ = p->main;
  return ESR_OK;
}

extern "C" int esr_pipeline_set_buffers(EsrPipeline* p, void* const* ids_dev, void* const* counts_dev, size_t ids_bytes,
                                        size_t counts_bytes, const float* loss_src, float* loss_log, int64_t loss_len,
                                        float* loss_host) {
  ESR_REQUIRE(ok(p) && ids_dev && counts_dev && loss_src && loss_log && loss_len > 0);
  for (int k = 0; k < p->depth; ++k) {
    ESR_REQUIRE(ids_dev[k] != nullptr && counts_dev[k] != nullptr);
    p->ids_dev[k] = ids_dev[k];
    p->counts_dev[k] = counts_dev[k];
  }
  p->ids_bytes = ids_bytes;
  p->counts_bytes = counts_bytes;
  p->loss_src = loss_src;
  p->loss_log = loss_log;
  p->loss_len = loss_len;
  p->loss_host = loss_host;
  return ESR_OK;
}

// which: 0 = the plan stage (side stream), 1 = the step stage (main stream).  Everything the calling thread launches on
// that stream until esr_pipeline_capture_end becomes the stage's graph for buffer parity k.
extern "C" int esr_pipeline_capture_begin(EsrPipeline* p, int32_t which) {
  ESR_REQUIRE(ok(p) && (which == 0 || which == 1) && p->capturing < 0);
  ESR_CUDA(cudaStreamBeginCapture(which == 0 ? p->side : p->main, cudaStreamCaptureModeThreadLocal));
  p->capturing = which;
  return ESR_OK;
}

extern "C" int esr_pipeline_capture_end(EsrPipeline* p, int32_t which, int32_t k) {
  ESR_REQUIRE(ok(p) && which == p->capturing && k >= 0 && k < p->depth);
  cudaGraph_t g = nullptr;
  p->capturing = -1;
  ESR_CUDA(cudaStreamEndCapture(which == 0 ? p->side : p->main, &g));
  cudaGraphExec_t* slot = which == 0 ? &p->g_plan[k] : &p->g_step[k];
  if (*slot) {
    cudaGraphExecDestroy(*slot);
    *slot = nullptr;
  }
  const cudaError_t e = cudaGraphInstantiate(slot, g, 0);
  cudaGraphDestroy(g);
  ESR_CUDA(e);
  return ESR_OK;
}

// One training step.  ids / counts: pinned host or device memory of ids_bytes / counts_bytes.  flags: bit 0 = also copy
// the step's loss to the pinned host log; bit 1 = the inputs were produced on caller_stream (NULL = the legacy default
// stream): stage them behind it.  Returns the step number in *step_out.
extern "C" int esr_pipeline_submit(EsrPipeline* p, const void* ids, const void* counts, esr_stream_t caller_stream,
                                   int32_t flags, int64_t* step_out) {
  const bool read_loss = (flags & 1) != 0;
  ESR_REQUIRE(ok(p) && ids && counts && p->capturing < 0);
  const int k = (int)(p->t % p->depth);
  ESR_REQUIRE(p->g_plan[k] != nullptr && p->g_step[k] != nullptr && p->ids_dev[k] != nullptr);
  if (flags & 2) {
    ESR_CUDA(cudaEventRecord(p->ev_in, static_cast<cudaStream_t>(caller_stream)));
    ESR_CUDA(cudaStreamWaitEvent(p->copy, p->ev_in, 0));
  }
  ESR_CUDA(cudaStreamWaitEvent(p->copy, p->ev_done[k], 0));  // staging / plan buffers k are free (step t - depth done)
  ESR_CUDA(cudaMemcpyAsync(p->ids_dev[k], ids, p->ids_bytes, cudaMemcpyDefault, p->copy));
  ESR_CUDA(cudaMemcpyAsync(p->counts_dev[k], counts, p->counts_bytes, cudaMemcpyDefault, p->copy));
  ESR_CUDA(cudaEventRecord(p->ev_copy[k], p->copy));
  ESR_CUDA(cudaStreamWaitEvent(p->side, p->ev_copy[k], 0));
  ESR_CUDA(cudaGraphLaunch(p->g_plan[k], p->side));
  ESR_CUDA(cudaEventRecord(p->ev_plan[k], p->side));
  ESR_CUDA(cudaStreamWaitEvent(p->main, p->ev_plan[k], 0));
  ESR_CUDA(cudaGraphLaunch(p->g_step[k], p->main));
  const int64_t slot = p->t % p->loss_len;
  ESR_CUDA(cudaMemcpyAsync(p->loss_log + slot, p->loss_src, sizeof(float), cudaMemcpyDeviceToDevice, p->main));
  if (read_loss && p->loss_host)
    ESR_CUDA(cudaMemcpyAsync(p->loss_host + slot, p->loss_log + slot, sizeof(float), cudaMemcpyDeviceToHost, p->main));
  ESR_CUDA(cudaEventRecord(p->ev_done[k], p->main));
  if (step_out) *step_out = p->t;
  p->t += 1;
  return ESR_OK;
}

extern "C" int esr_pipeline_sync(const EsrPipeline* p) {
  ESR_REQUIRE(ok(p));
  ESR_CUDA(cudaStreamSynchronize(p->copy));
  ESR_CUDA(cudaStreamSynchronize(p->side));
  ESR_CUDA(cudaStreamSynchronize(p->main));
  return ESR_OK;
}

extern "C" int esr_pipeline_destroy(EsrPipeline* p) {
  if (p == nullptr) return ESR_OK;
  ESR_REQUIRE(ok(p));
  cudaStreamSynchronize(p->copy);
  cudaStreamSynchronize(p->side);
  cudaStreamSynchronize(p->main);
  for (int k = 0; k < kMaxDepth; ++k) {
    if (p->g_plan[k]) cudaGraphExecDestroy(p->g_plan[k]);
    if (p->g_step[k]) cudaGraphExecDestroy(p->g_step[k]);
    cudaEventDestroy(p->ev_copy[k]);
    cudaEventDestroy(p->ev_plan[k]);
    cudaEventDestroy(p->ev_done[k]);
  }
  cudaEventDestroy(p->ev_in);
  cudaStreamDestroy(p->copy);
  cudaStreamDestroy(p->side);
  cudaStreamDestroy(p->main);
  p->magic = 0;
  delete p;
  return ESR_OK;
}
