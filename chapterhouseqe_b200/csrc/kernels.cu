// Ahead-of-time build of the generic (bytecode-interpreting) kernels + host launch helpers.
// The device code itself lives in device_code.cuh so that jit.cpp can hand the very same source
// to NVRTC with a program baked in as constants.
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <mutex>
#include <utility>

#include "device_code.cuh"

namespace chdb {

// Per-warp staging slice: a warp's rows of the widest fixed-width output; with Utf8 outputs the rebuilt
// offsets sit in front and the value bytes of short strings (or the long-string row tables) behind them.
size_t filter_project_stage_bytes(int max_out_width, int64_t avg_utf8_len) {
  const size_t rows = kTileRows / kWarps;
  size_t stage = rows * (size_t)(max_out_width < 4 ? 4 : max_out_width) + 32;
  if (avg_utf8_len >= 0) {
    const size_t off_stage = rows * 4 + 16;
    size_t strs = rows * (size_t)avg_utf8_len + 64;           // a warp slice of short strings, staged
    if (strs > 8 * 1024) strs = 0;                            // long strings go global -> global
    const size_t tables = (rows + 4) * 4 + rows * 4 + 16;     // ...and need the per-row tables instead
    const size_t want = off_stage + (strs > tables ? strs : tables);
    if (want > stage) stage = want;
  }
  return (stage + 15) & ~(size_t)15;
}

size_t filter_project_smem_bytes(size_t stage_bytes, bool has_utf8_out) {
  (void)has_utf8_out;
  return (size_t)kWarps * (stage_bytes + (size_t)(kTileRows / kWarps + 64));
}

cudaError_t launch_filter_project(const KernelParams& p, bool has64, size_t dyn_smem, cudaStream_t stream) {
  auto k32 = filter_project_kernel<uint32_t, kQuadsPerThread>;
  auto k64 = filter_project_kernel<uint64_t, kQuadsPerThread>;
  auto kern = has64 ? k64 : k32;
  if (dyn_smem > 48 * 1024) {   // opt in to the large window once per (kernel, device)
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> granted;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(mu);
    size_t& have = granted[{(const void*)kern, dev}];
    if (have < dyn_smem) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) return e;
      have = 200 * 1024;
    }
  }
  kern<<<dim3((unsigned)p.num_tiles), dim3(kThreads), dyn_smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace chdb
