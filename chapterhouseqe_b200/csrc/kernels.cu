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
constexpr int kMinCtasPerSm = 2;   // register budget the interpreter kernels are compiled for
}
#include "device_code.cuh"

namespace chdb {

namespace {
constexpr size_t kSmemPerSm = 228 * 1024;        // B200: 228 KB per SM, 1 KB of it reserved per resident CTA
constexpr size_t kSmemPerCtaMax = 227 * 1024;
constexpr size_t kStaticSmem = 512;              // mbarriers, s_tot, s_last + what the compiler adds
inline size_t up16(size_t x) { return (x + 15) & ~(size_t)15; }
inline size_t up128(size_t x) { return (x + 127) & ~(size_t)127; }
}  // namespace

// Decides what is staged and how deep the ring is.  Every buffer the kernel uses is a candidate; when the tile's
// slice of everything would not leave room for the wanted CTAs per SM with a ring of the minimum depth, the
// largest buffers are read from global memory instead (the producer warp then prefetches their slices into L2).
// Utf8 value bytes are staged only when the values are short (long values are copied global -> global by whole warps).
void plan_tile(const KernelParams& kp, TilePlan& tp, const int64_t* avg_utf8, bool many) {
  struct Buf { int slot; int kind; size_t bytes; };   // kind 0: validity, 1: offsets, 2: values
  std::vector<Buf> bufs;
  for (int s = 0; s < kp.n_in; s++) {
    tp.slot[s] = StageSlot{kNotStaged, kNotStaged, kNotStaged, 0};
    const ColumnDesc& c = kp.in[s];
    const uint8_t use = tp.use[s];
    if ((use & USE_VALIDITY) && c.validity != nullptr) bufs.push_back({s, 0, (size_t)kTileRows / 8});
    if (c.type == T_UTF8) {
      if (use & USE_OFFSETS) bufs.push_back({s, 1, (size_t)(kTileRows + 4) * 4});
      const int64_t avg = avg_utf8 ? avg_utf8[s] : -1;
      if ((use & USE_VALUES) && avg >= 0 && avg <= 32) bufs.push_back({s, 2, up16((size_t)kTileRows * (size_t)avg + std::max<size_t>((size_t)kTileRows * (size_t)avg / 32, 128) + 32)});
    } else if (use & USE_VALUES) {
      bufs.push_back({s, 2, c.width ? (size_t)kTileRows * c.width : (size_t)kTileRows / 8});
    }
  }
  const bool has_pred = kp.pred_end > kp.pred_begin;
  const int nq = 1 + kp.n_utf8;
  const size_t sctx_stride = up16(sizeof(StageCtx) + (size_t)std::max(kp.n_in, 1) * sizeof(ColumnDesc) +
                                  (many ? sizeof(BatchHeader) + (size_t)std::max(kp.n_out, 1) * sizeof(OutDesc) : 0));
  auto stage_bytes = [&]() {
    size_t t = 0;
    for (auto& b : bufs) t += up16(b.bytes);
    return up128(t);
  };
  int want = 2;   // resident CTAs per SM: two pipelines hide each other's bubbles; more cost registers
  if (const char* e = std::getenv("CHDB_WANT_CTAS")) want = std::max(1, std::min(4, std::atoi(e)));
  int min_stages = has_pred ? 3 : 2;   // A(i) and B(i-1) hold two stages; at least one more is being filled
  if (const char* e = std::getenv("CHDB_MIN_STAGES")) min_stages = std::max(min_stages, std::min(kMaxStages, std::atoi(e)));
  int stages = min_stages;
  size_t fixed = 0;
  while (true) {
    // the tables behind the ring (their size depends on the ring depth only through the per-stage ones)
    auto fixed_for = [&](int st) {
      size_t t = 0;
      t += up16((size_t)st * nq * kTileSlices * 4);
      t += up16((size_t)st * nq * kTileSlices * 8);
      t += up16(has_pred ? (size_t)st * kComputeWarps * 32 : 0);
      t += up16((size_t)(many ? st : 1) * kMaxOutCols * 4);
      t += up16((size_t)kComputeWarps * kp.n_bits * kBitWords * 4);
      t += up16(kp.long_strings ? (size_t)kComputeWarps * 2 * (kWarpRows + 4) * 4 : 0);
      t += up16(kp.n_bits > 0 ? 256 : 0);
      return t;
    };
    const size_t avail = std::min(kSmemPerCtaMax, kSmemPerSm / (size_t)want - 1024) - kStaticSmem;
    const size_t per_stage = stage_bytes() + sctx_stride;
    int fit = kMaxStages;
    while (fit > 0 && (size_t)fit * per_stage + fixed_for(fit) > avail) fit--;
    if (fit >= min_stages || bufs.empty()) {
      stages = std::max(min_stages, std::min(fit, bufs.empty() ? min_stages : kMaxStages));
      fixed = fixed_for(stages);
      break;
    }
    auto big = std::max_element(bufs.begin(), bufs.end(), [](const Buf& a, const Buf& b) { return a.bytes < b.bytes; });
    bufs.erase(big);
  }
  if (const char* e = std::getenv("CHDB_MAX_STAGES")) stages = std::max(min_stages, std::min(stages, std::atoi(e)));
  size_t off = 0;
  for (auto& b : bufs) {
    StageSlot& sl = tp.slot[b.slot];
    if (b.kind == 0) sl.validity = (uint32_t)off;
    else if (b.kind == 1) sl.offsets = (uint32_t)off;
    else { sl.values = (uint32_t)off; sl.values_cap = (uint32_t)b.bytes; }
    off += up16(b.bytes);
  }
  tp.stages = (uint32_t)stages;
  tp.stage_bytes = (uint32_t)stage_bytes();
  size_t at = (size_t)stages * tp.stage_bytes;
  auto table = [&](size_t bytes) { const size_t here = at; at += up16(bytes); return (uint32_t)here; };
  tp.sctx_off = table((size_t)stages * sctx_stride);
  tp.sctx_stride = (uint32_t)sctx_stride;
  tp.cnt_off = table((size_t)stages * nq * kTileSlices * 4);
  tp.pre_off = table((size_t)stages * nq * kTileSlices * 8);
  tp.sel_off = table(has_pred ? (size_t)stages * kComputeWarps * 32 : 0);
  tp.nulls_off = table((size_t)(many ? stages : 1) * kMaxOutCols * 4);
  tp.bits_off = table((size_t)kComputeWarps * kp.n_bits * kBitWords * 4);
  tp.ltab_off = table(kp.long_strings ? (size_t)kComputeWarps * 2 * (kWarpRows + 4) * 4 : 0);
  tp.pext_off = table(kp.n_bits > 0 ? 256 : 0);
  tp.dyn_smem = (uint32_t)at;
  (void)fixed;
  tp.ctas_per_sm = (uint32_t)std::max<size_t>(1, std::min<size_t>((size_t)want, kSmemPerSm / (at + kStaticSmem + 1024)));
  if (const char* e = std::getenv("CHDB_PLAN_VERBOSE")) {
    if (*e == '1') {
      int staged = 0, unstaged = 0;
      for (int s = 0; s < kp.n_in; s++) {
        const StageSlot& sl = tp.slot[s];
        staged += (sl.values != kNotStaged) + (sl.validity != kNotStaged) + (sl.offsets != kNotStaged);
        unstaged += ((tp.use[s] & USE_VALUES) && sl.values == kNotStaged) + ((tp.use[s] & USE_OFFSETS) && kp.in[s].type == T_UTF8 && sl.offsets == kNotStaged);
      }
      std::fprintf(stderr, "[chdb plan] stages %u x %u B, tables %zu B, dyn smem %u B, %u CTAs/SM, %d staged buffers, %d read from global\n",
                   tp.stages, tp.stage_bytes, at - (size_t)stages * tp.stage_bytes, tp.dyn_smem, tp.ctas_per_sm, staged, unstaged);
    }
  }
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
  // The plan counts on the whole 228 KB of the SM being shared memory (two CTAs of ~114 KB each): without this the
  // driver may pick a smaller carve-out, one CTA per SM becomes resident and the launch takes twice as long.
  e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  ok = true;
  return cudaSuccess;
}
}  // namespace

cudaError_t launch_stream_kernel(const void* kernel, const KernelParams& p, const TilePlan& tp, int sm_count, cudaStream_t stream) {
  cudaError_t e = grant_smem(kernel);
  if (e != cudaSuccess) return e;
  void* args[] = {const_cast<KernelParams*>(&p), const_cast<TilePlan*>(&tp)};
  cudaLaunchConfig_t cfg = {};
  // persistent: every CTA is resident (tickets hand the tiles to whoever runs); never more CTAs than tiles
  const unsigned resident = (unsigned)std::max(1, sm_count) * std::max(1u, tp.ctas_per_sm);
  cfg.gridDim = dim3(std::max(1u, std::min(resident, (unsigned)std::max(p.total_tiles, 1))));
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

cudaError_t launch_stream(const KernelParams& p, const TilePlan& tp, bool has64, int sm_count, cudaStream_t stream) {
  const bool many = p.many != nullptr;
  const void* kern = has64 ? (many ? (const void*)stream_kernel<uint64_t, true> : (const void*)stream_kernel<uint64_t, false>)
                           : (many ? (const void*)stream_kernel<uint32_t, true> : (const void*)stream_kernel<uint32_t, false>);
  return launch_stream_kernel(kern, p, tp, sm_count, stream);
}

cudaError_t launch_zero(void* p, size_t bytes, cudaStream_t stream) {
  const size_t n16 = bytes / 16;
  const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>(592, (n16 + kZeroThreads - 1) / kZeroThreads));
  zero_kernel<<<grid, kZeroThreads, 0, stream>>>((uint4*)p, n16);
  return cudaGetLastError();
}

}  // namespace chdb
