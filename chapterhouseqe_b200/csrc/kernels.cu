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
// static shared memory of the streaming kernels + what the compiler adds (alignment, its own slots)
constexpr size_t kStaticSmem = sizeof(SharedState) + 256;
}  // namespace

// Decides what is staged.  Every buffer the kernel uses is a candidate; when one stage of everything
// does not leave room for a ring of at least two stages, the largest buffers are read from global
// memory instead (the producer then prefetches their slices into L2).  Utf8 value bytes are staged
// only when the values are short (long values are copied global -> global by whole warps).
StagePlan plan_stages(const KernelParams& kp, KernelStage& st, const int64_t* avg_utf8, bool gather) {
  struct Buf { int slot; int kind; size_t bytes; };   // kind 0: validity, 1: offsets, 2: values
  std::vector<Buf> bufs;
  for (int s = 0; s < kp.n_in; s++) {
    st.slot[s] = StageSlot{kNotStaged, kNotStaged, kNotStaged, 0};
    const ColumnDesc& c = kp.in[s];
    const uint8_t use = st.use[s];
    if ((use & USE_VALIDITY) && c.validity != nullptr) bufs.push_back({s, 0, (size_t)kTileRows / 8});
    if (c.type == T_UTF8) {
      if (use & USE_OFFSETS) bufs.push_back({s, 1, (size_t)(kTileRows + 4) * 4});
      const int64_t avg = avg_utf8 ? avg_utf8[s] : -1;
      if ((use & USE_VALUES) && avg >= 0 && avg <= 32) bufs.push_back({s, 2, up16((size_t)kTileRows * (size_t)avg * 5 / 4 + 64)});
    } else if (use & USE_VALUES) {
      bufs.push_back({s, 2, c.width ? (size_t)kTileRows * c.width : (size_t)kTileRows / 8});
    }
  }
  const bool has_pred = kp.pred_end > kp.pred_begin;
  const size_t extras = gather && has_pred ? (size_t)kTileRows / 8 + (size_t)(1 + kp.n_utf8) * kSlices * 8 : 0;
  const size_t fixed_dyn = gather ? (size_t)kComputeWarps * kp.n_bits * kBitWords * 4 +
                                        (kp.long_strings ? (size_t)kComputeWarps * 2 * (kWarpRows + 4) * 4 : 0)
                                  : 0;
  const size_t fixed = fixed_dyn + kStaticSmem + 1024;
  auto stage_bytes = [&]() {
    size_t t = extras;
    for (auto& b : bufs) t += up16(b.bytes);
    return (t + 127) & ~(size_t)127;
  };
  struct Shape { int ctas, stages; };
  std::vector<Shape> shapes;
  for (int c = kMinCtasPerSm; c >= 1; c--)
    for (int n = kMaxStages; n > kComputeGroups; n--) shapes.push_back(Shape{c, n});   // the parity waits need a ring deeper than the groups
  if (const char* e = std::getenv(gather ? "CHDB_SHAPE" : "CHDB_SHAPE_SELECT")) {   // experiments: "ctas,stages" tried first
    int c = 0, n = 0;
    if (std::sscanf(e, "%d,%d", &c, &n) == 2 && c >= 1 && c <= 4 && n > kComputeGroups && n <= kMaxStages) shapes.insert(shapes.begin(), Shape{c, n});
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
  if (!pick.ctas) pick = Shape{1, kComputeGroups + 1};
  size_t off = 0;
  for (auto& b : bufs) {
    StageSlot& sl = st.slot[b.slot];
    if (b.kind == 0) sl.validity = (uint32_t)off;
    else if (b.kind == 1) sl.offsets = (uint32_t)off;
    else { sl.values = (uint32_t)off; sl.values_cap = (uint32_t)b.bytes; }
    off += up16(b.bytes);
  }
  st.sel_off = (uint32_t)off;
  st.prefix_off = (uint32_t)(off + kTileRows / 8);
  st.stage_bytes = (int32_t)stage_bytes();
  st.n_stages = pick.stages;
  StagePlan plan;
  plan.dyn_smem = (size_t)pick.stages * (size_t)st.stage_bytes + fixed_dyn;
  plan.ctas_per_sm = pick.ctas;
  return plan;
}

bool pdl_enabled();

cudaError_t launch_streaming(const void* kernel, const KernelParams& p, const KernelStage& st, const StagePlan& plan, int sm_count,
                             size_t* granted, bool after_kernel, cudaStream_t stream) {
  if (*granted == 0) {   // opt in to the large dynamic window once per kernel; static + dynamic may pass 48 KB for any plan
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
    if (e != cudaSuccess) return e;
    const size_t want = kSmemPerCtaMax - fa.sharedSizeBytes;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want);
    if (e != cudaSuccess) return e;
    *granted = want;
  }
  const int64_t resident = (int64_t)plan.ctas_per_sm * sm_count;
  const unsigned grid = (unsigned)std::min<int64_t>(p.num_tiles, resident);
  void* args[] = {const_cast<KernelParams*>(&p), const_cast<KernelStage*>(&st)};
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = plan.dyn_smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  // only a kernel that follows another kernel of the same batch (and waits for it with griddepcontrol.wait) may start
  // early; the select kernel follows the workspace memset and must not overtake it
  cfg.numAttrs = after_kernel && pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelExC(&cfg, kernel, args);
}

bool pdl_enabled() {
  static const bool on = [] { const char* e = std::getenv("CHDB_PDL"); return !(e && *e == '0'); }();
  return on;
}

namespace {
size_t* granted_slot(const void* kern) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> granted;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> g(mu);
  return &granted[{kern, dev}];
}
}  // namespace

cudaError_t launch_select(const KernelParams& p, const KernelStage& st, bool has64, const StagePlan& plan, int sm_count, cudaStream_t stream) {
  const void* kern = has64 ? (const void*)select_kernel<uint64_t> : (const void*)select_kernel<uint32_t>;
  return launch_streaming(kern, p, st, plan, sm_count, granted_slot(kern), false, stream);
}

cudaError_t launch_gather(const KernelParams& p, const KernelStage& st, bool has64, const StagePlan& plan, int sm_count, cudaStream_t stream) {
  const void* kern = has64 ? (const void*)gather_kernel<uint64_t> : (const void*)gather_kernel<uint32_t>;
  return launch_streaming(kern, p, st, plan, sm_count, granted_slot(kern), p.pred_end > p.pred_begin, stream);
}

cudaError_t launch_scan(const KernelParams& p, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)p.num_chunks, (unsigned)(1 + p.n_utf8));
  cfg.blockDim = dim3(kScanThreads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  void* args[] = {const_cast<KernelParams*>(&p)};
  return cudaLaunchKernelExC(&cfg, (const void*)scan_kernel, args);
}

}  // namespace chdb
