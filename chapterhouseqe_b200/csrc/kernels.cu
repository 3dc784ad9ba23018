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

namespace chdb {
constexpr int kMinCtasPerSm = 4;   // register budget the interpreter kernels are compiled for
}
#include "device_code.cuh"

namespace chdb {

namespace {
constexpr size_t kSmemPerSm = 228 * 1024;        // B200: 228 KB per SM, 1 KB of it reserved per resident CTA
constexpr size_t kSmemPerCtaMax = 227 * 1024;
constexpr size_t kStaticSmem = 256;              // s_full, s_nulls, s_last + what the compiler adds
inline size_t up16(size_t x) { return (x + 15) & ~(size_t)15; }
inline size_t up128(size_t x) { return (x + 127) & ~(size_t)127; }
}  // namespace

// Decides what is staged.  Every buffer the kernel uses is a candidate; when the tile's slice of everything
// would leave room for fewer than kWantCtas CTAs per SM, the largest buffers are read from global memory instead
// (the loading warp then prefetches their slices into L2).  Utf8 value bytes are staged only when the values
// are short (long values are copied global -> global by whole warps).
int plan_tile(const KernelParams& kp, TilePlan& tp, const int64_t* avg_utf8, bool many, bool stage) {
  struct Buf { int slot; int kind; size_t bytes; };   // kind 0: validity, 1: offsets, 2: values
  std::vector<Buf> bufs;
  for (int s = 0; s < kp.n_in; s++) {
    tp.slot[s] = StageSlot{kNotStaged, kNotStaged, kNotStaged, 0};
    if (!stage) continue;
    const ColumnDesc& c = kp.in[s];
    const uint8_t use = tp.use[s];
    if ((use & USE_VALIDITY) && c.validity != nullptr) bufs.push_back({s, 0, (size_t)kTileRows / 8});
    if (c.type == T_UTF8) {
      if (use & USE_OFFSETS) bufs.push_back({s, 1, (size_t)(kTileRows + 4) * 4});
      const int64_t avg = avg_utf8 ? avg_utf8[s] : -1;
      if ((use & USE_VALUES) && avg >= 0 && avg <= 32) bufs.push_back({s, 2, up16((size_t)kTileRows * (size_t)avg * 5 / 4 + 64)});
    } else if (use & USE_VALUES) {
      bufs.push_back({s, 2, c.width ? (size_t)kTileRows * c.width : (size_t)kTileRows / 8});
    }
  }
  const int nq = 1 + kp.n_utf8;
  // the tables behind the stage
  size_t tables = 0;
  auto table = [&](size_t bytes) { const size_t at = tables; tables += up16(bytes); return at; };
  const size_t t_cols = table((size_t)std::max(kp.n_in, 1) * sizeof(ColumnDesc));
  const size_t t_cnt = table((size_t)nq * kTileSlices * 4);
  const size_t t_pre = table((size_t)nq * kTileSlices * 8);
  const size_t t_tot = table((size_t)nq * 8);
  const size_t t_bits = table((size_t)kWarps * kp.n_bits * kBitWords * 4);
  const size_t t_ltab = table(kp.long_strings ? (size_t)kWarps * 2 * (kWarpRows + 4) * 4 : 0);
  const size_t t_pext = table(kp.n_bits > 0 ? 256 : 0);
  const size_t t_params = table(many ? sizeof(KernelParams) : 0);
  auto stage_bytes = [&]() {
    size_t t = 0;
    for (auto& b : bufs) t += up16(b.bytes);
    return up128(t);
  };
  int want = 4;   // fewer resident tiles than this and the loads / look-backs stop overlapping
  if (const char* e = std::getenv("CHDB_WANT_CTAS")) want = std::max(1, std::min(16, std::atoi(e)));
  auto ctas_for = [&](size_t dyn) { return (int)std::min<size_t>(16, kSmemPerSm / (dyn + kStaticSmem + 1024)); };
  while (!bufs.empty()) {
    const size_t dyn = stage_bytes() + tables;
    if (dyn + kStaticSmem <= kSmemPerCtaMax && ctas_for(dyn) >= want) break;
    auto big = std::max_element(bufs.begin(), bufs.end(), [](const Buf& a, const Buf& b) { return a.bytes < b.bytes; });
    bufs.erase(big);
  }
  size_t off = 0;
  for (auto& b : bufs) {
    StageSlot& sl = tp.slot[b.slot];
    if (b.kind == 0) sl.validity = (uint32_t)off;
    else if (b.kind == 1) sl.offsets = (uint32_t)off;
    else { sl.values = (uint32_t)off; sl.values_cap = (uint32_t)b.bytes; }
    off += up16(b.bytes);
  }
  const size_t sb = stage_bytes();
  tp.cols_off = (uint32_t)(sb + t_cols);
  tp.cnt_off = (uint32_t)(sb + t_cnt);
  tp.pre_off = (uint32_t)(sb + t_pre);
  tp.tot_off = (uint32_t)(sb + t_tot);
  tp.bits_off = (uint32_t)(sb + t_bits);
  tp.ltab_off = (uint32_t)(sb + t_ltab);
  tp.pext_off = (uint32_t)(sb + t_pext);
  tp.params_off = (uint32_t)(sb + t_params);
  tp.dyn_smem = (uint32_t)(sb + tables);
  return std::max(1, ctas_for(tp.dyn_smem));
}

namespace {
bool pdl_enabled() {
  static const bool on = [] { const char* e = std::getenv("CHDB_PDL"); return !(e && *e == '0'); }();
  return on;
}
// the large dynamic shared-memory window is opted in to once per (kernel, device)
cudaError_t grant_smem(const void* kern) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, bool> granted;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> g(mu);
  bool& ok = granted[{kern, dev}];
  if (ok) return cudaSuccess;
  cudaFuncAttributes fa;
  e = cudaFuncGetAttributes(&fa, kern);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemPerCtaMax - fa.sharedSizeBytes));
  if (e != cudaSuccess) return e;
  ok = true;
  return cudaSuccess;
}
}  // namespace

cudaError_t launch_stream_kernel(const void* kernel, const KernelParams& p, const TilePlan& tp, unsigned grid, cudaStream_t stream) {
  cudaError_t e = grant_smem(kernel);
  if (e != cudaSuccess) return e;
  void* args[] = {const_cast<KernelParams*>(&p), const_cast<TilePlan*>(&tp)};
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = tp.dyn_smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;   // (always follows launch_zero's kernel)
  return cudaLaunchKernelExC(&cfg, kernel, args);
}

cudaError_t launch_stream(const KernelParams& p, const TilePlan& tp, bool has64, int mode, unsigned grid, cudaStream_t stream) {
  const bool many = p.many != nullptr;
  const void* kern;
  if (mode == kSelect) kern = has64 ? (const void*)stream_kernel<uint64_t, false, kSelect> : (const void*)stream_kernel<uint32_t, false, kSelect>;
  else if (mode == kGather) kern = has64 ? (const void*)stream_kernel<uint64_t, false, kGather> : (const void*)stream_kernel<uint32_t, false, kGather>;
  else kern = has64 ? (many ? (const void*)stream_kernel<uint64_t, true, kFused> : (const void*)stream_kernel<uint64_t, false, kFused>)
                    : (many ? (const void*)stream_kernel<uint32_t, true, kFused> : (const void*)stream_kernel<uint32_t, false, kFused>);
  return launch_stream_kernel(kern, p, tp, grid, stream);
}

cudaError_t launch_zero(void* p, size_t bytes, cudaStream_t stream) {
  const size_t n16 = bytes / 16;
  const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>(592, (n16 + kZeroThreads - 1) / kZeroThreads));
  zero_kernel<<<grid, kZeroThreads, 0, stream>>>((uint4*)p, n16);
  return cudaGetLastError();
}

}  // namespace chdb
