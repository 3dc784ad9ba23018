// Ahead-of-time build of the generic (bytecode-interpreting) kernels + host launch helpers.
// The device code itself lives in device_code.cuh so that jit.cpp can hand the very same source
// to NVRTC with a program baked in as constants.
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_code.cuh"

namespace chdb {

// Per-warp staging slice: 256 rows of the widest fixed-width output, or of short Utf8 values.
size_t filter_project_stage_bytes(int max_out_width, int64_t avg_utf8_len) {
  const size_t rows = kTileRows / kWarps;
  size_t stage = rows * (size_t)(max_out_width < 4 ? 4 : max_out_width) + 32;
  if (avg_utf8_len > 0) {   // room for a warp slice of short strings (the staged Utf8 path)
    size_t want = rows * (size_t)avg_utf8_len + 64;
    if (want > 6 * 1024) want = 6 * 1024;
    if (want > stage) stage = want;
  }
  return (stage + 15) & ~(size_t)15;
}

size_t filter_project_smem_bytes(size_t stage_bytes, bool has_utf8_out) {
  const size_t rows = kTileRows / kWarps;
  size_t total = (size_t)kWarps * (stage_bytes + (size_t)kWarpBitStage);
  if (has_utf8_out) total += (size_t)kWarps * ((rows + 4) * 4 + rows * 4) + 16;
  return total;
}

cudaError_t launch_filter_project(const KernelParams& p, bool has64, size_t dyn_smem, cudaStream_t stream) {
  auto k32 = filter_project_kernel<uint32_t, kQuadsPerThread>;
  auto k64 = filter_project_kernel<uint64_t, kQuadsPerThread>;
  auto kern = has64 ? k64 : k32;
  if (dyn_smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem);
    if (e != cudaSuccess) return e;
  }
  kern<<<dim3((unsigned)p.num_tiles), dim3(kThreads), dyn_smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace chdb
