// Ahead-of-time build of the generic (bytecode-interpreting) kernels + host launch helpers.
// The device code itself lives in device_code.cuh so that jit.cpp can hand the very same source
// to NVRTC with a program baked in as constants.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "device_code.cuh"

namespace chdb {

namespace {
constexpr size_t kSmemPerSm = 228 * 1024;        // B200: 228 KB per SM, 1 KB of it reserved per resident CTA
constexpr size_t kSmemPerCtaMax = 227 * 1024;
inline size_t up16(size_t x) { return (x + 15) & ~(size_t)15; }
}  // namespace

size_t filter_project_static_smem() { return sizeof(SharedState) + 256; }   // + what the compiler adds (alignment, its own slots)

// Decides what is staged.  Every buffer of every slot is a candidate; when one stage of everything
// does not leave room for a ring of at least two stages, the largest buffers are read from global
// memory instead (the producer then prefetches their slices into L2).  Utf8 value bytes are staged
// only when the values are short (long values are copied global -> global by whole warps).
StagePlan plan_stages(KernelParams& kp, const int64_t* avg_utf8) {
  struct Buf { int slot; int kind; size_t bytes; };   // kind 0: validity, 1: offsets, 2: values
  std::vector<Buf> bufs;
  for (int s = 0; s < kp.n_in; s++) {
    kp.stage[s] = StageSlot{kNotStaged, kNotStaged, kNotStaged, 0};
    const ColumnDesc& c = kp.in[s];
    if (c.validity != nullptr) bufs.push_back({s, 0, (size_t)kTileRows / 8});
    if (c.type == T_UTF8) {
      bufs.push_back({s, 1, (size_t)(kTileRows + 4) * 4});
      const int64_t avg = avg_utf8 ? avg_utf8[s] : -1;
      if (avg >= 0 && avg <= 32) bufs.push_back({s, 2, up16((size_t)kTileRows * (size_t)avg * 5 / 4 + 64)});
    } else {
      bufs.push_back({s, 2, c.width ? (size_t)kTileRows * c.width : (size_t)kTileRows / 8});
    }
  }
  const size_t fixed_dyn = (size_t)kWriterGroups * 2 * kp.n_bits * kBitWords * 4 +
                           (kp.long_strings ? (size_t)kWriterWarps * 2 * (kWarpRows + 4) * 4 : 0);
  const size_t fixed = fixed_dyn + filter_project_static_smem() + 1024;
  auto stage_bytes = [&]() {
    size_t t = 0;
    for (auto& b : bufs) t += up16(b.bytes);
    return (t + 127) & ~(size_t)127;
  };
  struct Shape { int ctas, stages; };
  std::vector<Shape> shapes;
  for (int c = kMinCtasPerSm; c >= 1; c--)
    for (int st : {5, 4, 3}) shapes.push_back(Shape{c, st});
  shapes.push_back(Shape{1, 2});
  if (const char* e = std::getenv("CHDB_SHAPE")) {   // experiments: "ctas,stages" tried first
    int c = 0, st = 0;
    if (std::sscanf(e, "%d,%d", &c, &st) == 2 && c >= 1 && c <= 4 && st >= 2 && st <= kMaxStages) shapes.insert(shapes.begin(), Shape{c, st});
  }
  Shape pick{0, 0};
  while (true) {
    const size_t sb = stage_bytes();
    for (const Shape& sh : shapes) {
      const size_t per_cta = fixed + (size_t)sh.stages * sb;
      if (per_cta * sh.ctas <= kSmemPerSm && per_cta - 1024 <= kSmemPerCtaMax) { pick = sh; break; }
    }
    if (pick.ctas || bufs.empty()) break;
    auto big = std::max_element(bufs.begin(), bufs.end(), [](const Buf& a, const Buf& b) { return a.bytes < b.bytes; });
    bufs.erase(big);
  }
  if (!pick.ctas) pick = Shape{1, 2};
  size_t off = 0;
  for (auto& b : bufs) {
    StageSlot& sl = kp.stage[b.slot];
    if (b.kind == 0) sl.validity = (uint32_t)off;
    else if (b.kind == 1) sl.offsets = (uint32_t)off;
    else { sl.values = (uint32_t)off; sl.values_cap = (uint32_t)b.bytes; }
    off += up16(b.bytes);
  }
  kp.stage_bytes = (int32_t)stage_bytes();
  kp.n_stages = pick.stages;
  StagePlan plan;
  plan.dyn_smem = (size_t)pick.stages * (size_t)kp.stage_bytes + fixed_dyn;
  plan.ctas_per_sm = pick.ctas;
  return plan;
}

cudaError_t launch_filter_project(const KernelParams& p, bool has64, const StagePlan& plan, int sm_count, cudaStream_t stream) {
  auto k32 = filter_project_kernel<uint32_t, kQuadsPerThread>;
  auto k64 = filter_project_kernel<uint64_t, kQuadsPerThread>;
  auto kern = has64 ? k64 : k32;
  {   // opt in to the large dynamic window once per (kernel, device); static + dynamic may pass 48 KB for any plan
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> granted;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(mu);
    size_t& have = granted[{(const void*)kern, dev}];
    if (have < plan.dyn_smem) {
      cudaFuncAttributes fa;
      cudaError_t e = cudaFuncGetAttributes(&fa, kern);
      if (e != cudaSuccess) return e;
      const size_t want = kSmemPerCtaMax - fa.sharedSizeBytes;
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want);
      if (e != cudaSuccess) return e;
      have = want;
    }
  }
  const int64_t resident = (int64_t)plan.ctas_per_sm * sm_count;
  const unsigned grid = (unsigned)std::min<int64_t>(p.num_tiles, resident);
  kern<<<dim3(grid), dim3(kThreads), plan.dyn_smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace chdb
