// C-ABI plumbing: error state, version, launch counter, GEMM dispatch.
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <mutex>
#include <unordered_set>

#include "common.cuh"

namespace mmt {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
const uint64_t* g_step_ctr = nullptr;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_status(cudaError_t e, const char* what) {
  set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return (int)e;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize must be set once per (kernel, device): a per-device table instead of a
// process-global latch, so a second GPU driven from the same process is configured too.
int ensure_dynamic_smem(const void* fn, size_t bytes, const char* what) {
  static std::mutex mu;
  static std::unordered_set<uint64_t> done;
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t key = (uint64_t)(uintptr_t)fn * 1315423911ull + (uint64_t)dev;
  std::lock_guard<std::mutex> lk(mu);
  if (done.count(key)) return 0;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return cuda_status(e, what);
  done.insert(key);
  return 0;
}

int gemm_simt(const mmt_gemm_desc& d, cudaStream_t stream);
int gemm_tc(const mmt_gemm_desc& d, cudaStream_t stream);
int gemm_tc_pair_bf16(const mmt_gemm_desc& d, cudaStream_t stream);

}  // namespace mmt

extern "C" {

int mmt_version(void) { return 100; }

int mmt_last_error(char* buf, size_t len) {
  if (!buf || len == 0) return MMT_E_ARG;
  strncpy(buf, mmt::g_err, len - 1);
  buf[len - 1] = 0;
  return 0;
}

int64_t mmt_launch_count(void) { return mmt::g_launches.load(); }

int mmt_set_step_counter(const uint64_t* dev_counter) {
  mmt::g_step_ctr = dev_counter;
  return 0;
}

int mmt_gemm(const mmt_gemm_desc* d, void* stream) {
  MMT_ARG_CHECK(d != nullptr, MMT_E_ARG, "mmt_gemm: null descriptor");
  MMT_ARG_CHECK(d->A && d->B && d->C, MMT_E_ARG, "mmt_gemm: null operand");
  MMT_ARG_CHECK(d->M >= 0 && d->N >= 0 && d->K >= 0 && d->batch >= 1 && d->batch_inner >= 1,
                MMT_E_SHAPE, "mmt_gemm: bad shape M=%d N=%d K=%d batch=%d/%d", d->M, d->N, d->K,
                d->batch, d->batch_inner);
  MMT_ARG_CHECK(d->epilogue >= MMT_EPI_NONE && d->epilogue <= MMT_EPI_DGELU, MMT_E_ARG,
                "mmt_gemm: bad epilogue %d", d->epilogue);
  MMT_ARG_CHECK(d->epilogue == MMT_EPI_NONE || d->aux != nullptr, MMT_E_ARG,
                "mmt_gemm: epilogue %d needs aux", d->epilogue);
  MMT_ARG_CHECK(d->batch <= 65535, MMT_E_SHAPE, "mmt_gemm: batch %d > 65535", d->batch);
  MMT_ARG_CHECK(d->colsum == nullptr || d->precision != MMT_PREC_FP32, MMT_E_UNSUPPORTED,
                "mmt_gemm: the fused column-sum epilogue exists on the tensor-core paths only");
  if (d->M == 0 || d->N == 0) return 0;
  if (d->precision == MMT_PREC_TF32) return mmt::gemm_tc(*d, (cudaStream_t)stream);
  if (d->precision == MMT_PREC_BF16) return mmt::gemm_tc_pair_bf16(*d, (cudaStream_t)stream);
  MMT_ARG_CHECK(d->precision == MMT_PREC_FP32, MMT_E_ARG, "mmt_gemm: bad precision %d",
                d->precision);
  return mmt::gemm_simt(*d, (cudaStream_t)stream);
}

}  // extern "C"
