// Host runtime behind the C ABI: contexts, device-resident batches, Arrow C Data Interface
// import/export, and the executor that turns a compiled Program + an input batch into one launch
// of the fused kernel.  Mirrors the call contract of the reference's record_utils functions
// (filter_record.rs:21-39, record_projection.rs:16-76): pure functions of (batch, expression).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <condition_variable>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "jit.hpp"
#include "kernels.cuh"
#include "program.hpp"

namespace chdb {

#define CUDA_CHECK(expr)                                                                        \
  do {                                                                                          \
    cudaError_t err__ = (expr);                                                                 \
    if (err__ != cudaSuccess)                                                                   \
      throw Error(CHDB_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(err__) + " at " #expr); \
  } while (0)

static inline size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline size_t bitmap_bytes(int64_t n) { return (size_t)((n + 7) / 8); }
constexpr size_t kPad = 64;   // every device buffer is readable this far past its logical end

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
struct CtxCore {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::atomic<int64_t> launches{0};
  std::atomic<int64_t> jit_launches{0};   // launches that ran an NVRTC-specialised kernel
  std::atomic<int64_t> alloc_misses{0};   // device allocations that went to cudaMallocAsync (block cache misses)
  int sm_count = 148;                     // persistent grid = CTAs per SM x SMs
  std::mutex mu;
  std::vector<void*> pinned_free;   // small pinned blocks for count read-back
  static constexpr size_t kPinnedBlock = 1024;
  // size-class cache of pinned host buffers: downloaded batches land in page-locked memory so the
  // D2H copies run at PCIe speed; ArrowArray.release hands the blocks back here.
  std::map<size_t, std::vector<void*>> host_free;
  // Device blocks released by finished batches, by size: steady-state batches of one shape reuse them
  // without going back to cudaMallocAsync / cudaFreeAsync (which cost a dozen driver calls per batch and
  // now and then take the allocator's millisecond slow path).  Reuse is safe because everything this
  // library does with a block is ordered on the ctx's one stream.
  // Each cached block remembers how many launch sets had been enqueued when it was released (see launch_set: a
  // select kernel that runs on the second stream may only touch blocks released before the previous launch set).
  std::map<size_t, std::vector<std::pair<void*, int64_t>>> dev_free;
  size_t dev_cached = 0;
  std::vector<cudaEvent_t> events_free;   // recycled: creating / destroying an event is a driver resource call
  // Select / gather overlap (launch_set): launch sets are numbered; before_last[k & 3] fires when everything enqueued on
  // `stream` before launch set k's LAST kernel has completed; select_done[k & 3] when its select kernel (second stream) has.
  cudaStream_t aux_stream = nullptr;
  cudaStream_t copy_stream = nullptr;   // Parquet decode of several row groups: the next row group's H2D copy (parquet.inc)
  cudaEvent_t select_done[4] = {nullptr, nullptr, nullptr, nullptr}, before_last[4] = {nullptr, nullptr, nullptr, nullptr};
  std::atomic<int64_t> launch_seq{0};
  std::atomic<int64_t> overlapped{0};     // launch sets whose select kernel ran on aux_stream
  static constexpr size_t kDevCacheCap = (size_t)24 << 30;   // beyond this, released blocks go back to the pool

  ~CtxCore() {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    if (aux_stream) cudaStreamSynchronize(aux_stream);
    for (auto& kv : dev_free)
      for (auto& p : kv.second) cudaFreeAsync(p.first, stream);
    for (int i = 0; i < 4; i++) {
      if (select_done[i]) cudaEventDestroy(select_done[i]);
      if (before_last[i]) cudaEventDestroy(before_last[i]);
    }
    if (aux_stream) cudaStreamDestroy(aux_stream);
    if (copy_stream) { cudaStreamSynchronize(copy_stream); cudaStreamDestroy(copy_stream); }
    for (cudaEvent_t e : events_free) cudaEventDestroy(e);
    if (stream) cudaStreamSynchronize(stream);
    for (void* p : pinned_free) cudaFreeHost(p);
    for (auto& kv : host_free)
      for (void* p : kv.second) cudaFreeHost(p);
    if (stream) cudaStreamDestroy(stream);
  }
  static size_t host_class(size_t bytes) {
    size_t c = 4096;
    while (c < bytes) c <<= 1;
    return c;
  }
  void* host_get(size_t bytes, size_t* cls) {
    *cls = host_class(bytes + 64);
    {
      std::lock_guard<std::mutex> g(mu);
      auto it = host_free.find(*cls);
      if (it != host_free.end() && !it->second.empty()) {
        void* p = it->second.back();
        it->second.pop_back();
        return p;
      }
    }
    void* p = nullptr;
    if (cudaSetDevice(device) != cudaSuccess || cudaHostAlloc(&p, *cls, cudaHostAllocPortable) != cudaSuccess) {
      cudaGetLastError();
      throw Error(CHDB_ERR_CUDA, "cudaHostAlloc failed for a " + std::to_string(*cls) + "-byte output buffer");
    }
    return p;
  }
  void host_put(void* p, size_t cls) {
    std::lock_guard<std::mutex> g(mu);
    host_free[cls].push_back(p);
  }
  void* pinned_get() {
    {
      std::lock_guard<std::mutex> g(mu);
      if (!pinned_free.empty()) {
        void* p = pinned_free.back();
        pinned_free.pop_back();
        return p;
      }
    }
    void* p = nullptr;
    CUDA_CHECK(cudaMallocHost(&p, kPinnedBlock));
    return p;
  }
  void pinned_put(void* p) {
    std::lock_guard<std::mutex> g(mu);
    pinned_free.push_back(p);
  }
  cudaEvent_t event_get() {
    {
      std::lock_guard<std::mutex> g(mu);
      if (!events_free.empty()) {
        cudaEvent_t e = events_free.back();
        events_free.pop_back();
        return e;
      }
    }
    cudaEvent_t e = nullptr;
    CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return e;
  }
  void event_put(cudaEvent_t e) {
    std::lock_guard<std::mutex> g(mu);
    events_free.push_back(e);
  }
};
using Core = std::shared_ptr<CtxCore>;

struct DevBuf {
  Core core;
  void* ptr = nullptr;
  size_t bytes = 0;
  ~DevBuf() {
    if (!ptr) return;
    {
      std::lock_guard<std::mutex> g(core->mu);
      if (core->dev_cached + bytes <= CtxCore::kDevCacheCap) {
        core->dev_free[bytes].push_back({ptr, core->launch_seq.load()});
        core->dev_cached += bytes;
        return;
      }
    }
    cudaSetDevice(core->device);
    cudaFreeAsync(ptr, core->stream);
  }
};
using Buf = std::shared_ptr<DevBuf>;

static Buf dev_alloc(const Core& core, size_t bytes) {
  auto b = std::make_shared<DevBuf>();
  b->core = core;
  b->bytes = round_up(bytes + kPad, 256);
  {
    std::lock_guard<std::mutex> g(core->mu);
    auto it = core->dev_free.find(b->bytes);
    if (it != core->dev_free.end() && !it->second.empty()) {
      b->ptr = it->second.back().first;
      it->second.pop_back();
      core->dev_cached -= b->bytes;
      return b;
    }
  }
  core->alloc_misses++;
  CUDA_CHECK(cudaMallocAsync(&b->ptr, b->bytes, core->stream));
  return b;
}

// A cached block that was released before launch set `max_seq + 1` was enqueued: everything that ever touched it was
// enqueued before that launch set.  Null when the cache holds none (the caller then stays on the one stream).
static Buf dev_alloc_released_by(const Core& core, size_t bytes, int64_t max_seq) {
  const size_t want = round_up(bytes + kPad, 256);
  std::lock_guard<std::mutex> g(core->mu);
  auto it = core->dev_free.find(want);
  if (it == core->dev_free.end()) return nullptr;
  auto& v = it->second;
  for (size_t i = 0; i < v.size(); i++) {   // (oldest first)
    if (v[i].second > max_seq) continue;
    auto b = std::make_shared<DevBuf>();
    b->core = core;
    b->bytes = want;
    b->ptr = v[i].first;
    v.erase(v.begin() + (std::ptrdiff_t)i);
    core->dev_cached -= want;
    return b;
  }
  return nullptr;
}

// One launch set (zero kernel + stream kernel): what its output batches share.  The counts of the run are mirrored
// into pinned host memory by the kernel's last CTA; `done` fires when the launch has finished.
struct LaunchShared {
  Core core;
  cudaEvent_t done = nullptr;
  bool waited = false;
  Buf workspace;              // counts, look-back descriptors, tickets, bit-packed outputs
  Buf params;                 // many-batch launches: per-batch records + tile table
  void* host = nullptr;       // pinned: count mirrors (+ staging of the records)
  size_t host_cls = 0;
  ~LaunchShared() {
    if (done) core->event_put(done);
    if (host) core->host_put(host, host_cls);
  }
  void wait() {
    if (waited) return;
    CUDA_CHECK(cudaSetDevice(core->device));
    CUDA_CHECK(cudaEventSynchronize(done));
    waited = true;
  }
  bool ready() {
    if (waited) return true;
    CUDA_CHECK(cudaSetDevice(core->device));
    const cudaError_t e = cudaEventQuery(done);
    if (e == cudaErrorNotReady) return false;
    CUDA_CHECK(e);
    waited = true;
    return true;
  }
};
// One batch's share of it.
struct RunResult {
  std::shared_ptr<LaunchShared> launch;
  uint64_t* host = nullptr;   // [n_counts] counts then [n_counts] = error word
  int n_counts = 0;
  void wait() { launch->wait(); }
};

struct DeviceColumn {
  InputColumn meta;
  const void* values = nullptr;      // kernels index these (Utf8: virtual base so that offsets apply directly)
  const uint8_t* validity = nullptr;
  const int32_t* offsets = nullptr;
  Buf values_buf, validity_buf, offsets_buf;   // ownership (null when wrapped / borrowed)
  int64_t null_count = 0;            // -1: in result->host[count_index]; -2: unknown (count on download)
  int count_index = -1;
  int64_t value_bytes = 0;           // Utf8: bytes in [first_offset, ..); -1: result->host[bytes_index]; -2: read offsets
  int bytes_index = -1;
  int64_t first_offset = 0;          // Utf8: offsets[0]; -2: read offsets
  bool nullable_per_batch = false;   // project_record: Field.nullable = null_count != 0
};

}  // namespace chdb

struct chdb_ctx {
  chdb::Core core;
};

struct chdb_program {
  std::unique_ptr<chdb::Program> p;
};

struct chdb_device_batch {
  std::atomic<int> refs{1};          // chdb_device_batch_retain / _release
  chdb::Core core;
  int64_t num_rows = 0;              // -1: result->host[0]
  std::vector<chdb::DeviceColumn> cols;
  std::shared_ptr<chdb::RunResult> result;
  // core->launch_seq when the batch was created: whatever fills its buffers was enqueued on the ctx stream before the
  // launch set of that number (-1: unknown; such a batch is never read from the second stream)
  int64_t born = -1;
};

namespace chdb {

static void resolve(chdb_device_batch* b) {
  if (!b->result) return;
  b->result->wait();
  const uint64_t* h = b->result->host;
  if (b->num_rows == -1) b->num_rows = (int64_t)h[0];
  for (auto& c : b->cols) {
    if (c.null_count == -1) c.null_count = (int64_t)h[c.count_index];
    if (c.value_bytes == -1) c.value_bytes = (int64_t)h[c.bytes_index];
  }
}

static void check_run_error(const chdb_device_batch* b) {
  if (!b->result) return;
  b->result->wait();
  const uint64_t word = b->result->host[b->result->n_counts];
  if (word == 0) return;
  const uint64_t packed = ~word;
  const int32_t code = (int32_t)(packed & 0xFF);
  const long long row = (long long)((packed >> 8) & 0xFFFFFFFFFFFFull);
  if (code == CHDB_ERR_DIVIDE_BY_ZERO)
    throw Error(code, "Divide by zero error (row " + std::to_string(row) + ")");
  throw Error(code, "Arithmetic overflow: Overflow happened on row " + std::to_string(row));
}

// ------------------------------------------------------------------------------------------
// Arrow C Data Interface: import (upload)
// ------------------------------------------------------------------------------------------
static int64_t popcount_bits(const uint8_t* bits, int64_t bit_off, int64_t n) {
  int64_t c = 0;
  for (int64_t i = 0; i < n; i++) {
    int64_t b = bit_off + i;
    c += (bits[b >> 3] >> (b & 7)) & 1;
  }
  return c;
}

// Copies n bits starting at bit_off into a byte-aligned device bitmap.
static Buf upload_bits(const Core& core, const uint8_t* bits, int64_t bit_off, int64_t n) {
  Buf out = dev_alloc(core, bitmap_bytes(n));
  if (n == 0) return out;
  if ((bit_off & 7) == 0) {
    CUDA_CHECK(cudaMemcpyAsync(out->ptr, bits + (bit_off >> 3), bitmap_bytes(n), cudaMemcpyHostToDevice, core->stream));
  } else {
    std::vector<uint8_t> tmp(bitmap_bytes(n), 0);
    for (int64_t i = 0; i < n; i++) {
      int64_t b = bit_off + i;
      if ((bits[b >> 3] >> (b & 7)) & 1) tmp[i >> 3] |= (uint8_t)(1u << (i & 7));
    }
    CUDA_CHECK(cudaMemcpyAsync(out->ptr, tmp.data(), tmp.size(), cudaMemcpyHostToDevice, core->stream));
    CUDA_CHECK(cudaStreamSynchronize(core->stream));  // tmp dies at scope end
  }
  return out;
}

static std::unique_ptr<chdb_device_batch> upload_batch(const Core& core, const ::ArrowArray* in, const ::ArrowSchema* schema) {
  if (!in) throw Error(CHDB_ERR_INVALID_ARGUMENT, "null ArrowArray");
  std::vector<InputColumn> cols = parse_schema(schema);
  if ((int64_t)cols.size() != in->n_children)
    throw Error(CHDB_ERR_INVALID_ARGUMENT, "ArrowArray / ArrowSchema child count mismatch");
  CUDA_CHECK(cudaSetDevice(core->device));
  std::unique_ptr<chdb_device_batch> b(new chdb_device_batch);
  b->core = core;
  b->born = core->launch_seq.load();
  b->num_rows = in->length;
  const int64_t n = in->length;
  for (size_t ci = 0; ci < cols.size(); ci++) {
    const ::ArrowArray* ch = in->children[ci];
    DeviceColumn dc;
    dc.meta = cols[ci];
    if (!dc.meta.supported)
      throw Error(CHDB_ERR_NOT_IMPLEMENTED, "column '" + dc.meta.name + "': Arrow format '" + dc.meta.format + "' is not supported");
    if (ch->length < n)
      throw Error(CHDB_ERR_INVALID_ARGUMENT, "child array shorter than the batch");
    const int64_t off = ch->offset + in->offset;
    const uint8_t* vbits = ch->n_buffers > 0 ? (const uint8_t*)ch->buffers[0] : nullptr;
    int64_t nulls = ch->null_count;
    if (vbits == nullptr) nulls = 0;
    else if (nulls < 0 || ch->length != n) nulls = n - popcount_bits(vbits, off, n);
    dc.null_count = nulls;
    if (vbits != nullptr && nulls > 0) {
      dc.validity_buf = upload_bits(core, vbits, off, n);
      dc.validity = (const uint8_t*)dc.validity_buf->ptr;
    }
    if (dc.meta.type == T_BOOL) {
      dc.values_buf = upload_bits(core, (const uint8_t*)ch->buffers[1], off, n);
      dc.values = dc.values_buf->ptr;
    } else if (dc.meta.type == T_UTF8) {
      const int32_t* o = (const int32_t*)ch->buffers[1];
      static const int32_t zero_off[1] = {0};
      if (o == nullptr) o = zero_off; else o += off;
      const int64_t first = n ? o[0] : 0, last = n ? o[n] : 0;
      dc.offsets_buf = dev_alloc(core, (size_t)(n + 1) * 4);
      CUDA_CHECK(cudaMemcpyAsync(dc.offsets_buf->ptr, o, (size_t)(n ? n + 1 : 1) * 4, cudaMemcpyHostToDevice, core->stream));
      dc.offsets = (const int32_t*)dc.offsets_buf->ptr;
      // value bytes are placed so that (first % 16) is preserved: 4-byte aligned sources stay aligned
      const size_t lead = (size_t)(first & 15);
      dc.values_buf = dev_alloc(core, lead + (size_t)(last - first));
      if (last > first)
        CUDA_CHECK(cudaMemcpyAsync((uint8_t*)dc.values_buf->ptr + lead, (const uint8_t*)ch->buffers[2] + first,
                                   (size_t)(last - first), cudaMemcpyHostToDevice, core->stream));
      dc.values = (const uint8_t*)dc.values_buf->ptr + lead - first;   // virtual base
      dc.first_offset = first;
      dc.value_bytes = last - first;
    } else {
      const size_t w = (size_t)dc.meta.width;
      dc.values_buf = dev_alloc(core, (size_t)n * w);
      if (n)
        CUDA_CHECK(cudaMemcpyAsync(dc.values_buf->ptr, (const uint8_t*)ch->buffers[1] + (size_t)off * w, (size_t)n * w,
                                   cudaMemcpyHostToDevice, core->stream));
      dc.values = dc.values_buf->ptr;
    }
    b->cols.push_back(std::move(dc));
  }
  return b;
}

// ------------------------------------------------------------------------------------------
// Arrow C Data Interface: export (download)
// ------------------------------------------------------------------------------------------
struct ExportedArray {
  Core core;
  std::vector<std::pair<void*, size_t>> owned;   // pinned blocks (pointer, size class) from core->host_get
  std::vector<const void*> buffers;
  std::vector<::ArrowArray> child_storage;
  std::vector<::ArrowArray*> children;
};
static void release_array(::ArrowArray* a) {
  if (!a || !a->release) return;
  auto* ex = (ExportedArray*)a->private_data;
  for (auto& c : ex->child_storage)
    if (c.release) c.release(&c);
  for (auto& p : ex->owned) ex->core->host_put(p.first, p.second);
  delete ex;
  a->release = nullptr;
}
struct ExportedSchema {
  std::string format, name;
  std::vector<::ArrowSchema> child_storage;
  std::vector<::ArrowSchema*> children;
};
static void release_schema(::ArrowSchema* s) {
  if (!s || !s->release) return;
  auto* ex = (ExportedSchema*)s->private_data;
  for (auto& c : ex->child_storage)
    if (c.release) c.release(&c);
  delete ex;
  s->release = nullptr;
}
static void make_schema(::ArrowSchema* out, const std::string& format, const std::string& name, int64_t flags, size_t n_children) {
  auto* ex = new ExportedSchema;
  ex->format = format;
  ex->name = name;
  ex->child_storage.resize(n_children);
  for (auto& c : ex->child_storage) { std::memset(&c, 0, sizeof(c)); ex->children.push_back(&c); }
  std::memset(out, 0, sizeof(*out));
  out->format = ex->format.c_str();
  out->name = ex->name.c_str();
  out->flags = flags;
  out->n_children = (int64_t)n_children;
  out->children = n_children ? ex->children.data() : nullptr;
  out->release = release_schema;
  out->private_data = ex;
}
static void* host_alloc(ExportedArray* ex, size_t bytes) {
  size_t cls = 0;
  void* p = ex->core->host_get(bytes, &cls);
  ex->owned.push_back({p, cls});
  return p;
}

// Value bytes of a Utf8 column; wrapped columns learn it from their offsets on first use (two 4-byte reads).
static int64_t utf8_value_bytes(const DeviceColumn& c, int64_t n) {
  if (c.value_bytes != -3) return c.value_bytes;
  DeviceColumn& m = const_cast<DeviceColumn&>(c);
  int32_t ends[2] = {0, 0};
  CUDA_CHECK(cudaMemcpy(&ends[0], c.offsets, 4, cudaMemcpyDeviceToHost));
  CUDA_CHECK(cudaMemcpy(&ends[1], c.offsets + n, 4, cudaMemcpyDeviceToHost));
  m.first_offset = ends[0];
  m.value_bytes = ends[1] - ends[0];
  return m.value_bytes;
}

// A download in two steps, so that a caller can do something else (or poll) while the copies run:
// download_begin() needs the run's counts (it waits for the launch unless it has finished) and enqueues every
// device -> pinned host copy on the ctx stream; download_end() runs after those copies have completed.
struct Download {
  std::unique_ptr<ExportedArray> top;
  std::vector<ExportedArray*> exs;
  std::vector<int64_t> first;   // Utf8: offsets[0] of the copied offsets (rebased to 0 in download_end)
  int64_t n = 0;
};

static void download_begin(chdb_device_batch* b, Download& d) {
  const Core& core = b->core;
  CUDA_CHECK(cudaSetDevice(core->device));
  check_run_error(b);
  resolve(b);
  const int64_t n = d.n = b->num_rows;
  d.top.reset(new ExportedArray);
  ExportedArray* top = d.top.get();
  top->core = core;
  top->child_storage.resize(b->cols.size());
  for (auto& c : top->child_storage) std::memset(&c, 0, sizeof(c));
  d.exs.assign(b->cols.size(), nullptr);
  d.first.assign(b->cols.size(), 0);
  for (size_t ci = 0; ci < b->cols.size(); ci++) {
    ExportedArray* ex = d.exs[ci] = new ExportedArray;
    ex->core = core;
    ::ArrowArray& a = top->child_storage[ci];
    a.private_data = ex;
    a.release = release_array;
    DeviceColumn& c = b->cols[ci];
    void* hv = nullptr;
    uint8_t* hval = nullptr;
    if (c.validity != nullptr && c.null_count != 0 && n) {
      hval = (uint8_t*)host_alloc(ex, bitmap_bytes(n));
      CUDA_CHECK(cudaMemcpyAsync(hval, c.validity, bitmap_bytes(n), cudaMemcpyDeviceToHost, core->stream));
    }
    if (c.meta.type == T_UTF8) {
      int32_t* ho = (int32_t*)host_alloc(ex, (size_t)(n + 1) * 4);
      ho[0] = 0;
      if (n) CUDA_CHECK(cudaMemcpyAsync(ho, c.offsets, (size_t)(n + 1) * 4, cudaMemcpyDeviceToHost, core->stream));
      int64_t first = 0, last = 0;
      if (n) {
        if (utf8_value_bytes(c, n) >= 0 && c.first_offset >= 0) {
          first = c.first_offset;
          last = first + c.value_bytes;
        } else {   // a sliced view: the extent is in the offsets themselves
          CUDA_CHECK(cudaStreamSynchronize(core->stream));
          first = ho[0];
          last = ho[n];
        }
      }
      d.first[ci] = first;
      hv = host_alloc(ex, (size_t)(last - first));
      if (last > first)
        CUDA_CHECK(cudaMemcpyAsync(hv, (const uint8_t*)c.values + first, (size_t)(last - first), cudaMemcpyDeviceToHost, core->stream));
      ex->buffers = {hval, ho, hv};
    } else if (c.meta.type == T_BOOL) {
      hv = host_alloc(ex, bitmap_bytes(n));
      if (n) CUDA_CHECK(cudaMemcpyAsync(hv, c.values, bitmap_bytes(n), cudaMemcpyDeviceToHost, core->stream));
      ex->buffers = {hval, hv};
    } else {
      const size_t w = (size_t)c.meta.width;
      hv = host_alloc(ex, (size_t)n * w);
      if (n) CUDA_CHECK(cudaMemcpyAsync(hv, c.values, (size_t)n * w, cudaMemcpyDeviceToHost, core->stream));
      ex->buffers = {hval, hv};
    }
  }
}

static void download_end(chdb_device_batch* b, Download& d, ::ArrowArray* out, ::ArrowSchema* out_schema) {
  const int64_t n = d.n;
  ExportedArray* top = d.top.get();
  make_schema(out_schema, "+s", "", 0, b->cols.size());
  for (size_t ci = 0; ci < b->cols.size(); ci++) {
    DeviceColumn& c = b->cols[ci];
    ExportedArray* ex = d.exs[ci];
    ::ArrowArray& a = top->child_storage[ci];
    if (c.meta.type == T_UTF8 && d.first[ci] != 0) {
      int32_t* ho = (int32_t*)ex->buffers[1];
      for (int64_t i = 0; i <= n; i++) ho[i] -= (int32_t)d.first[ci];
    }
    int64_t nulls = c.null_count;
    if (nulls < 0) {  // unknown (sliced view): count what was copied
      const uint8_t* hval = (const uint8_t*)ex->buffers[0];
      nulls = hval ? n - popcount_bits(hval, 0, n) : 0;
    }
    if (nulls == 0) ex->buffers[0] = nullptr;  // arrow-select drops the bitmap when no nulls survive
    a.length = n;
    a.null_count = nulls;
    a.offset = 0;
    a.n_buffers = (int64_t)ex->buffers.size();
    a.buffers = ex->buffers.data();
    a.n_children = 0;
    int64_t flags = c.meta.flags & ARROW_FLAG_NULLABLE;
    if (c.nullable_per_batch) flags = nulls != 0 ? ARROW_FLAG_NULLABLE : 0;
    make_schema(out_schema->children[ci], c.meta.format, c.meta.name, flags, 0);
    top->children.push_back(&a);
  }
  std::memset(out, 0, sizeof(*out));
  out->length = n;
  out->null_count = 0;
  out->n_buffers = 1;
  top->buffers = {nullptr};
  out->buffers = top->buffers.data();
  out->n_children = (int64_t)top->children.size();
  out->children = top->children.data();
  out->release = release_array;
  out->private_data = d.top.release();
}

static void download_batch(chdb_device_batch* b, ::ArrowArray* out, ::ArrowSchema* out_schema) {
  Download d;
  download_begin(b, d);
  CUDA_CHECK(cudaStreamSynchronize(b->core->stream));
  download_end(b, d, out, out_schema);
}

}  // namespace chdb

// A host-batch call in flight (chdb_filter_record_async): 0 kernels running, 1 device -> host copies running,
// 2 ready, 3 result taken.
struct chdb_pending {
  chdb::Core core;
  std::unique_ptr<chdb_device_batch> dev_in, dev_out;
  chdb::Download dl;
  int stage = 0;
  cudaEvent_t copied = nullptr;
};

namespace chdb {

// ------------------------------------------------------------------------------------------
// executor
// ------------------------------------------------------------------------------------------
static std::unique_ptr<chdb_device_batch> view_rows(const chdb_device_batch* in, int64_t n) {
  std::unique_ptr<chdb_device_batch> v(new chdb_device_batch);
  v->core = in->core;
  v->born = in->born;
  v->num_rows = n;
  v->cols = in->cols;
  for (auto& c : v->cols) {
    if (n != in->num_rows) {
      if (c.validity) c.null_count = -2;
      c.value_bytes = -2;
    }
  }
  return v;
}

static void check_schema(const Program& p, const chdb_device_batch* in) {
  if (p.schema.size() != in->cols.size())
    throw Error(CHDB_ERR_INVALID_ARGUMENT, "batch has " + std::to_string(in->cols.size()) + " columns, program was compiled for " +
                                               std::to_string(p.schema.size()));
  for (size_t i = 0; i < p.schema.size(); i++)
    if (p.schema[i].format != in->cols[i].meta.format || p.schema[i].name != in->cols[i].meta.name)
      throw Error(CHDB_ERR_INVALID_ARGUMENT, "batch schema differs from the schema the program was compiled for (column " +
                                                 std::to_string(i) + ")");
}

static DeviceColumn const_column(const Core& core, const OutputColumn& o) {
  DeviceColumn dc;
  dc.meta.name = o.name;
  dc.meta.format = o.format;
  dc.meta.type = o.type;
  dc.meta.width = o.width;
  dc.nullable_per_batch = true;
  if (o.type == T_UTF8) {
    int32_t offs[2] = {0, (int32_t)o.str.size()};
    dc.offsets_buf = dev_alloc(core, 8);
    CUDA_CHECK(cudaMemcpyAsync(dc.offsets_buf->ptr, offs, 8, cudaMemcpyHostToDevice, core->stream));
    dc.values_buf = dev_alloc(core, o.str.size());
    if (!o.str.empty())
      CUDA_CHECK(cudaMemcpyAsync(dc.values_buf->ptr, o.str.data(), o.str.size(), cudaMemcpyHostToDevice, core->stream));
    dc.offsets = (const int32_t*)dc.offsets_buf->ptr;
    dc.value_bytes = (int64_t)o.str.size();
  } else {
    uint64_t v = o.imm;
    dc.values_buf = dev_alloc(core, 8);
    CUDA_CHECK(cudaMemcpyAsync(dc.values_buf->ptr, &v, 8, cudaMemcpyHostToDevice, core->stream));
  }
  dc.values = dc.values_buf->ptr;
  CUDA_CHECK(cudaStreamSynchronize(core->stream));  // locals above
  return dc;
}

// Output buffers of one launch over many batches come out of one slab (one allocation per launch, not per column
// and batch).
struct Slab {
  Buf buf;
  size_t used = 0;
  void* take(size_t bytes) {
    const size_t at = used;
    used += round_up(bytes + kPad, 256);
    if (used > buf->bytes) throw Error(CHDB_ERR_CUDA, "internal: output slab overflow");
    return (uint8_t*)buf->ptr + at;
  }
};

// Everything execute() decides about one batch before anything is launched.
struct ZeroReq { size_t col; int ko; bool validity; size_t bytes; size_t at; };
struct Prepared {
  const chdb_device_batch* in = nullptr;
  std::unique_ptr<chdb_device_batch> view;
  std::unique_ptr<chdb_device_batch> out;
  bool launch = false, compact = false, any_const = false, all_const = false;
  int64_t n = 0;
  int ko = 0, n_utf8 = 0, n_counts = 1;
  int64_t avg_utf8 = -1;   // mean value length of the longest Utf8 output (sizes the staging area); -1: none
  uint64_t shape = 0;      // which input slots / outputs carry validity: batches of one launch must agree
  std::vector<ZeroReq> zero_reqs;
  ColumnDesc in_desc[kMaxInCols];
  OutDesc out_desc[kMaxOutCols];
};

// Bytes of output capacity prepare() will ask the allocator for (upper bound), for sizing a slab.
static size_t output_capacity(const Program& p, const chdb_device_batch* in) {
  const int64_t n = in->num_rows < 0 ? 0 : in->num_rows;
  size_t t = 0;
  for (auto& o : p.outputs) {
    if (o.kind == OutputColumn::CONST) continue;
    if (o.type == T_UTF8) {
      const int64_t vb = o.in_col >= 0 ? std::max<int64_t>(in->cols[o.in_col].value_bytes, 0) : 0;
      t += round_up((size_t)vb + kPad, 256) + round_up((size_t)(n + 1) * 4 + kPad, 256);
    } else if (o.type != T_BOOL) {
      t += round_up((size_t)n * (size_t)o.width + kPad, 256);
    }
  }
  return t + 1024;
}

static void prepare(chdb_ctx* ctx, const Program& p, const chdb_device_batch* in_orig, Slab* slab, Prepared& P) {
  const Core& core = ctx->core;
  if (in_orig->core->device != core->device)
    throw Error(CHDB_ERR_INVALID_ARGUMENT, "batch lives on another device than the ctx");
  check_schema(p, in_orig);
  if (in_orig->num_rows < 0) resolve(const_cast<chdb_device_batch*>(in_orig));
  if (in_orig->result) check_run_error(in_orig);

  const chdb_device_batch* in = in_orig;
  bool use_pred = p.has_pred;
  if (p.requires_single_row && in->num_rows != 1) throw Error(p.single_row_code, p.single_row_msg);
  if (p.has_pred && p.pred_const) {
    // arrow-select does not broadcast a len-1 mask: [true] keeps row 0, [false] keeps nothing
    if (in->num_rows < 1)
      throw Error(CHDB_ERR_INVALID_ARGUMENT,
                  "Invalid argument error: Filter predicate of length 1 is larger than target array of length 0");
    P.view = view_rows(in, p.pred_const_value ? 1 : 0);
    in = P.view.get();
    use_pred = false;
  }
  P.in = in;
  const int64_t n = P.n = in->num_rows;
  auto alloc = [&](size_t bytes, Buf* keep) -> void* {
    if (slab) { *keep = slab->buf; return slab->take(bytes); }
    *keep = dev_alloc(core, bytes);
    return (*keep)->ptr;
  };

  P.out.reset(new chdb_device_batch);
  chdb_device_batch* out = P.out.get();
  out->core = core;
  out->born = core->launch_seq.load();   // (outputs of a launch set: renumbered after it, see execute)

  // len-1 constant outputs (record_projection.rs:73: RecordBatch::try_new checks equal lengths)
  P.any_const = false;
  P.all_const = !p.outputs.empty();
  for (auto& o : p.outputs) {
    P.any_const |= o.kind == OutputColumn::CONST;
    P.all_const &= o.kind == OutputColumn::CONST;
  }
  if (P.any_const && !P.all_const && !use_pred && n != 1)   // (under a predicate the surviving rows count: checked after the launch)
    throw Error(CHDB_ERR_INVALID_ARGUMENT, "Invalid argument error: all columns in a record batch must have the same length");

  // which outputs go through the kernel
  const bool compact = P.compact = use_pred;
  bool any_kernel_out = false;
  for (auto& o : p.outputs) any_kernel_out |= o.kind == OutputColumn::EXPR || (o.kind == OutputColumn::PASS && compact);
  const bool launch = P.launch = n > 0 && (any_kernel_out || compact);

  std::memset(P.in_desc, 0, p.slot_to_col.size() * sizeof(ColumnDesc));
  std::memset(P.out_desc, 0, p.outputs.size() * sizeof(OutDesc));
  int n_utf8 = 0, n_counts = 1;
  if (launch) {
    for (size_t s = 0; s < p.slot_to_col.size(); s++) {
      const DeviceColumn& c = in->cols[p.slot_to_col[s]];
      P.in_desc[s].values = c.values;
      P.in_desc[s].validity = c.validity;
      P.in_desc[s].offsets = c.offsets;
      P.in_desc[s].type = c.meta.type;
      P.in_desc[s].width = (uint8_t)c.meta.width;
      if (c.validity) P.shape |= 1ull << s;
      // TMA bulk copies move 16-byte units
      if ((((uintptr_t)c.values | (uintptr_t)c.validity | (uintptr_t)c.offsets) & 15u) != 0)
        throw Error(CHDB_ERR_INVALID_ARGUMENT, "device buffers must be 16-byte aligned");
    }
    for (auto& o : p.outputs)
      if ((o.kind == OutputColumn::EXPR || (o.kind == OutputColumn::PASS && compact)) && o.type == T_UTF8) n_utf8++;
    n_counts = 1 + n_utf8;
  }

  auto slot_has_validity = [&](int slot) { return in->cols[p.slot_to_col[slot]].validity != nullptr; };
  auto expr_may_be_null = [&](const OutputColumn& o) {
    for (int i = o.begin; i < o.end; i++) {
      const Instr& ins = p.instrs[i];
      if (ins.src == SRC_COL && slot_has_validity(ins.slot)) return true;
      if (ins.op == OP_CMP_UTF8) {
        if (ins.slot != 0xFF && slot_has_validity(ins.slot)) return true;
        const uint8_t sb = (uint8_t)(ins.imm >> 56);
        if (sb != 0xFF && slot_has_validity(sb)) return true;
      }
    }
    return false;
  };

  int utf8_seen = 0, ko = 0;
  // Bit-packed outputs (Boolean values, validity bitmaps) are merged into with atomicOr at slice edges, so they start
  // zeroed: they live, with the look-back descriptors and the counts, in ONE region that one small kernel zeroes.
  out->cols.reserve(p.outputs.size());
  for (size_t k = 0; k < p.outputs.size(); k++) {
    const OutputColumn& o = p.outputs[k];
    DeviceColumn dc;
    const bool through_kernel = launch && (o.kind == OutputColumn::EXPR || (o.kind == OutputColumn::PASS && compact));
    if (o.kind == OutputColumn::CONST) {
      dc = const_column(core, o);
    } else if (o.kind == OutputColumn::PASS && !through_kernel) {
      dc = in->cols[o.in_col];   // shared buffers, like the reference's Arc clone
      dc.meta.name = o.name;
      dc.nullable_per_batch = !o.keep_declared_nullable;
    } else {
      dc.meta.name = o.name;
      dc.meta.format = o.format;
      dc.meta.type = o.type;
      dc.meta.width = o.width;
      dc.meta.flags = o.declared_nullable ? ARROW_FLAG_NULLABLE : 0;
      dc.nullable_per_batch = !o.keep_declared_nullable;
      if (!through_kernel) {  // n == 0: empty column of the right type
        dc.values_buf = dev_alloc(core, 16);
        dc.values = dc.values_buf->ptr;
        if (o.type == T_UTF8) {
          dc.offsets_buf = dev_alloc(core, 4);
          CUDA_CHECK(cudaMemsetAsync(dc.offsets_buf->ptr, 0, 4, core->stream));
          dc.offsets = (const int32_t*)dc.offsets_buf->ptr;
        }
      } else {
        OutDesc& od = P.out_desc[ko];
        od.kind = o.kind == OutputColumn::EXPR ? OUT_EXPR : OUT_PASS;
        od.type = o.type;
        od.width = (uint8_t)o.width;
        od.slot = (uint8_t)(o.slot >= 0 ? o.slot : 0);
        od.begin = (uint8_t)o.begin;
        od.end = (uint8_t)o.end;
        od.utf8_index = 0xFF;
        bool nullable = false;
        if (o.kind == OutputColumn::PASS) {
          if (o.slot < 0) throw Error(CHDB_ERR_INVALID_ARGUMENT, "internal: pass-through column without a kernel slot");
          nullable = in->cols[o.in_col].validity != nullptr;
        } else {
          nullable = expr_may_be_null(o);
        }
        if (o.type == T_UTF8) {
          const DeviceColumn& src = in->cols[o.in_col];
          const int64_t vb = utf8_value_bytes(src, n);
          if (vb < 0) throw Error(CHDB_ERR_NOT_IMPLEMENTED, "filtering a sliced Utf8 view");
          P.avg_utf8 = std::max<int64_t>(P.avg_utf8, (vb + n - 1) / n);
          dc.values = alloc((size_t)vb, &dc.values_buf);
          dc.offsets = (const int32_t*)alloc((size_t)(n + 1) * 4, &dc.offsets_buf);
          od.offsets = (int32_t*)dc.offsets;
          od.utf8_index = (uint8_t)utf8_seen;
          dc.value_bytes = -1;
          dc.bytes_index = 1 + utf8_seen;
          utf8_seen++;
          od.values = (void*)dc.values;
        } else if (o.type == T_BOOL) {
          P.zero_reqs.push_back({k, ko, false, round_up(bitmap_bytes(n), 4), 0});
        } else {
          dc.values = alloc((size_t)n * (size_t)o.width, &dc.values_buf);
          od.values = (void*)dc.values;
        }
        if (nullable) {
          P.zero_reqs.push_back({k, ko, true, round_up(bitmap_bytes(n), 4), 0});
          od.count_index = (uint8_t)n_counts;
          dc.null_count = -1;
          dc.count_index = n_counts;
          n_counts++;
          P.shape |= 1ull << (32 + ko);
        }
        ko++;
      }
    }
    out->cols.push_back(std::move(dc));
  }
  P.ko = ko;
  P.n_utf8 = n_utf8;
  P.n_counts = n_counts;
  if (!launch) {
    if (P.any_const && !P.all_const && compact)   // n == 0: nothing survives next to len-1 literal columns
      throw Error(CHDB_ERR_INVALID_ARGUMENT, "Invalid argument error: all columns in a record batch must have the same length");
    out->num_rows = P.all_const ? 1 : n;
    if (compact && !P.all_const) out->num_rows = 0;   // n == 0
  }
}

// Bytes of the zeroed per-batch workspace: counts | error word | done || look-back descriptors || bit-packed outputs
static size_t workspace_layout(Prepared& P, size_t* ws_counts_out, size_t* ws_desc_out) {
  const int64_t num_tiles = (P.n + kTileRows - 1) / kTileRows;
  const int nq = 1 + P.n_utf8;
  const size_t ws_counts = round_up((size_t)(P.n_counts + 2) * 8, 128);
  const size_t num_groups = (size_t)(num_tiles + kGroupTiles - 1) / kGroupTiles;   // (two-launch form: group totals behind the tile totals)
  const size_t ws_desc = P.compact ? round_up((size_t)nq * ((size_t)num_tiles + num_groups) * 8, 128) : 0;
  size_t total = ws_counts + ws_desc;
  for (auto& z : P.zero_reqs) { z.at = total; total += round_up(z.bytes, 128); }
  *ws_counts_out = ws_counts;
  *ws_desc_out = ws_desc;
  return total;
}

// Points the batch's bit-packed outputs and its header at its part of the workspace.
static void bind_workspace(Prepared& P, const std::shared_ptr<LaunchShared>& ls, uint8_t* ws, size_t ws_counts, BatchHeader& bh,
                           uint64_t* host_counts, int32_t first_tile) {
  for (auto& z : P.zero_reqs) {
    DeviceColumn& dc = P.out->cols[z.col];
    OutDesc& od = P.out_desc[z.ko];
    if (z.validity) {
      dc.validity_buf = ls->workspace;
      dc.validity = ws + z.at;
      od.validity = ws + z.at;
    } else {
      dc.values_buf = ls->workspace;
      dc.values = ws + z.at;
      od.values = ws + z.at;
    }
  }
  std::memset(&bh, 0, sizeof(bh));
  bh.num_rows = P.n;
  bh.counts = (uint64_t*)ws;
  bh.error_word = (uint64_t*)ws + P.n_counts;
  bh.done = (uint32_t*)((uint64_t*)ws + P.n_counts + 1);
  bh.desc = (uint64_t*)(ws + ws_counts);
  bh.host_counts = host_counts;
  bh.num_tiles = (int32_t)((P.n + kTileRows - 1) / kTileRows);
  bh.first_tile = first_tile;
  auto res = std::make_shared<RunResult>();
  res->launch = ls;
  res->host = host_counts;
  res->n_counts = P.n_counts;
  P.out->result = res;
  P.out->num_rows = P.compact ? -1 : P.n;
}

// The program part of the kernel parameters + the plan, from the first batch of a launch.
static int fill_program_params(const Program& p, const Prepared& P, KernelParams& kp, TilePlan& tp, bool many) {
  kp.n_in = (int)p.slot_to_col.size();
  kp.n_out = P.ko;
  kp.n_utf8 = P.n_utf8;
  kp.n_counts = P.n_counts;
  kp.pred_begin = P.compact ? p.pred_begin : 0;
  kp.pred_end = P.compact ? p.pred_end : 0;
  std::memcpy(kp.in, P.in_desc, p.slot_to_col.size() * sizeof(ColumnDesc));
  std::memcpy(kp.out, P.out_desc, (size_t)P.ko * sizeof(OutDesc));
  std::memcpy(kp.instrs, p.instrs.data(), p.instrs.size() * sizeof(Instr));
  std::memcpy(kp.strpool, p.strpool.data(), p.strpool.size());
  // bit-packed outputs assembled in shared memory: Boolean values and validity bitmaps
  kp.n_bits = 0;
  for (int k = 0; k < P.ko; k++) kp.n_bits += (kp.out[k].type == T_BOOL ? 1 : 0) + (kp.out[k].validity != nullptr ? 1 : 0);
  kp.long_strings = P.avg_utf8 > 16 ? 1 : 0;
  int64_t slot_avg[kMaxInCols];
  for (size_t s = 0; s < p.slot_to_col.size(); s++) {
    const DeviceColumn& c = P.in->cols[p.slot_to_col[s]];
    const int64_t vb = c.meta.type == T_UTF8 ? utf8_value_bytes(c, P.n) : -1;
    slot_avg[s] = vb >= 0 ? (vb + P.n - 1) / P.n : -1;
  }
  // which buffers the kernel touches
  std::memset(&tp, 0, sizeof(tp));
  auto mark_instrs = [&](int begin, int end) {
    for (int i = begin; i < end; i++) {
      const Instr& ins = kp.instrs[i];
      if (ins.op == OP_CMP_UTF8) {
        if (ins.slot != 0xFF) tp.use[ins.slot] |= USE_VALUES | USE_VALIDITY | USE_OFFSETS;
        const uint8_t sb = (uint8_t)(ins.imm >> 56);
        if (sb != 0xFF) tp.use[sb] |= USE_VALUES | USE_VALIDITY | USE_OFFSETS;
      } else if (ins.src == SRC_COL) {
        tp.use[ins.slot] |= USE_VALUES | USE_VALIDITY;
      }
    }
  };
  mark_instrs(kp.pred_begin, kp.pred_end);
  for (int i = kp.pred_begin; i < kp.pred_end; i++)
    if (kp.instrs[i].op == OP_CMP_UTF8) tp.pred_reads_utf8 = 1;
  for (int k = 0; k < P.ko; k++) {
    const OutDesc& od = kp.out[k];
    if (od.kind == OUT_EXPR) mark_instrs(od.begin, od.end);
    else tp.use[od.slot] |= USE_VALUES | USE_VALIDITY | USE_OFFSETS;
  }
  // CHDB_STAGE=0 (experiment): nothing is staged, the lanes read their rows from global memory (the loading warp
  // still prefetches the tile into L2)
  static const bool stage = [] { const char* e = std::getenv("CHDB_STAGE"); return !(e && *e == '0'); }();
  return plan_tile(kp, tp, slot_avg, many, stage);
}

// Zero kernel + stream kernel.  Long scans run the same device code specialised for this program by NVRTC
// (jit.cpp); short ones, or boxes without NVRTC, run the bytecode interpreter kernels.
// Single-batch launches with a predicate run as two kernels (select, then gather: device_code.cuh StreamMode) -- no CTA
// holds a staged tile while it waits for another CTA.  Small batches keep the single fused launch (one launch less on
// a latency-bound call).  CHDB_SPLIT = 0 | never : always fused;  always : two kernels whatever the size (tests);
// unset | auto : two kernels from kSplitAutoRows rows on.
constexpr int64_t kSplitAutoRows = 1 << 16;
// CHDB_OVERLAP=0: every kernel on the ctx's one stream
static bool overlap_enabled() {
  static const bool on = [] { const char* e = std::getenv("CHDB_OVERLAP"); return !(e && *e == '0'); }();
  return on;
}
static bool split_wanted(int64_t rows) {
  const char* e = std::getenv("CHDB_SPLIT");
  if (e && (*e == '0' || !std::strcmp(e, "never"))) return false;
  if (e && !std::strcmp(e, "always")) return true;
  return rows >= kSplitAutoRows;
}

// Two-launch form: what the select kernel can take over from the gather kernel (kernels.cuh KernelParams).
static void set_split_flags(KernelParams& kp) {
  kp.early_counts = 1;
  for (int k = 0; k < kp.n_out; k++)
    if (kp.out[k].kind != OUT_PASS) kp.early_counts = 0;
}

static void launch_set(const Core& core, const Program& p, const KernelParams& kp_in, const TilePlan& tp, int ctas_per_sm, unsigned grid,
                       void* ws, size_t ws_bytes, int64_t rows, const std::shared_ptr<LaunchShared>& ls, bool overlap = false) {
  KernelParams kp_traced;
  const KernelParams* kpp = &kp_in;
  Buf trace_buf;
  const char* trace_path = std::getenv("CHDB_TRACE");   // debugging aid: per-tile phase time stamps of this launch -> file
  const size_t trace_bytes = (size_t)8192 * 8 * 8;
  if (trace_path && *trace_path) {
    trace_buf = dev_alloc(core, trace_bytes);
    CUDA_CHECK(cudaMemsetAsync(trace_buf->ptr, 0, trace_bytes, core->stream));
    kp_traced = kp_in;
    kp_traced.trace = (uint64_t*)trace_buf->ptr;
    kpp = &kp_traced;
  }
  const KernelParams& kp = *kpp;
  // Launch set k.  With `overlap`, its zero and select kernels go to the ctx's second stream, where they wait for
  // before_last[k-1] only: they run next to launch set k-1's gather kernel (both are bandwidth-hungry but neither
  // saturates HBM alone), and the gather kernel of this set waits for them on the ctx stream.  execute() grants
  // `overlap` only when nothing those two kernels touch can still be in use by launch set k-1's last kernel.
  if (!core->before_last[0]) {
    for (int i = 0; i < 4; i++) CUDA_CHECK(cudaEventCreateWithFlags(&core->before_last[i], cudaEventDisableTiming));
  }
  const int64_t seq = core->launch_seq.load();
  cudaStream_t sel_stream = core->stream;
  if (overlap) {
    if (!core->aux_stream) {
      CUDA_CHECK(cudaStreamCreateWithFlags(&core->aux_stream, cudaStreamNonBlocking));
      for (int i = 0; i < 4; i++) CUDA_CHECK(cudaEventCreateWithFlags(&core->select_done[i], cudaEventDisableTiming));
    }
    sel_stream = core->aux_stream;
    core->overlapped++;
    CUDA_CHECK(cudaStreamWaitEvent(sel_stream, core->before_last[(seq - 1) & 3], 0));
  }
  CUDA_CHECK(launch_zero(ws, ws_bytes, sel_stream));
  core->launches++;
  const JitKernel* jk = nullptr;
  const JitMode jm = jit_mode();
  if (jm == JitMode::Always || (jm == JitMode::Auto && rows >= kJitAutoRows)) {
    std::string why;
    jk = jit_get(kp, p.has64, std::min(ctas_per_sm, 8), &why);
  }
  const bool split = kp.b.selbits != nullptr;
  if (split) {
    // select: nothing staged, only the small tables in shared memory
    TilePlan tps = tp;
    plan_tile(kp, tps, nullptr, false, false);
    const cudaError_t se = jk ? jit_launch_stream(jk, kp, tps, 1, grid, sel_stream) : launch_stream(kp, tps, p.has64, 1, grid, sel_stream);
    core->launches++;
    if (jk) core->jit_launches++;
    if (se != cudaSuccess) throw Error(CHDB_ERR_CUDA, std::string("kernel launch failed: ") + cudaGetErrorString(se));
    if (overlap) {
      CUDA_CHECK(cudaEventRecord(core->select_done[seq & 3], sel_stream));
      CUDA_CHECK(cudaStreamWaitEvent(core->stream, core->select_done[seq & 3], 0));
    }
  }
  CUDA_CHECK(cudaEventRecord(core->before_last[seq & 3], core->stream));
  core->launch_seq++;
  const int mode = split ? 2 : 0;
  const cudaError_t le = jk ? jit_launch_stream(jk, kp, tp, mode, grid, core->stream) : launch_stream(kp, tp, p.has64, mode, grid, core->stream);
  core->launches++;
  if (jk) core->jit_launches++;
  if (le != cudaSuccess) throw Error(CHDB_ERR_CUDA, std::string("kernel launch failed: ") + cudaGetErrorString(le));
  CUDA_CHECK(cudaEventRecord(ls->done, core->stream));   // the counts arrive in pinned memory with the kernel's last CTA
  if (trace_buf) {
    std::vector<uint64_t> h(trace_bytes / 8);
    CUDA_CHECK(cudaMemcpyAsync(h.data(), trace_buf->ptr, trace_bytes, cudaMemcpyDeviceToHost, core->stream));
    CUDA_CHECK(cudaStreamSynchronize(core->stream));
    if (FILE* f = std::fopen(trace_path, "wb")) { std::fwrite(h.data(), 8, h.size(), f); std::fclose(f); }
  }
}

static void after_launch(Prepared& P) {
  chdb_device_batch* out = P.out.get();
  if (P.all_const) {
    out->num_rows = 1;   // record_projection.rs:73 over len-1 arrays, whatever the filter kept (the launch only raises the predicate's errors)
  } else if (P.any_const && P.compact) {
    // literal columns are len 1: RecordBatch::try_new accepts them only next to exactly one surviving row
    check_run_error(out);
    resolve(out);
    if (out->num_rows != 1)
      throw Error(CHDB_ERR_INVALID_ARGUMENT, "Invalid argument error: all columns in a record batch must have the same length");
  }
}

// CHDB_HOST_TIMING=1: where the host spends its time inside execute() (debugging aid): calls that take longer than
// 300 us report their phases on stderr.
struct HostClock {
  bool on;
  std::chrono::steady_clock::time_point t0, t;
  double ph[5] = {0, 0, 0, 0, 0};
  int64_t misses0 = 0;
  const CtxCore* core = nullptr;
  HostClock() {
    static const bool enabled = [] { const char* e = std::getenv("CHDB_HOST_TIMING"); return e && *e == '1'; }();
    on = enabled;
    if (on) t0 = t = std::chrono::steady_clock::now();
  }
  void lap(int i) {
    if (!on) return;
    const auto now = std::chrono::steady_clock::now();
    ph[i] += std::chrono::duration<double, std::micro>(now - t).count();
    t = now;
  }
  ~HostClock() {
    if (!on) return;
    const double total = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    if (total > 300)
      std::fprintf(stderr, "[chdb host] execute %.0f us: prepare %.0f, workspace %.0f, params+plan %.0f, launches %.0f, after %.0f; %lld block cache misses\n",
                   total, ph[0], ph[1], ph[2], ph[3], ph[4], core ? (long long)(core->alloc_misses.load() - misses0) : -1ll);
  }
};

static std::unique_ptr<chdb_device_batch> execute(chdb_ctx* ctx, const Program& p, const chdb_device_batch* in_orig) {
  HostClock hc;
  const Core& core = ctx->core;
  hc.core = core.get();
  hc.misses0 = core->alloc_misses.load();
  CUDA_CHECK(cudaSetDevice(core->device));
  Prepared P;
  prepare(ctx, p, in_orig, nullptr, P);
  if (!P.launch) return std::move(P.out);
  hc.lap(0);

  size_t ws_counts, ws_desc;
  const size_t ws_total = workspace_layout(P, &ws_counts, &ws_desc);
  // two-launch form: the selection bitmap (whole tiles, every word written by the select kernel) sits behind the zeroed part
  const bool split = P.compact && split_wanted(P.n);
  const size_t sel_bytes = split ? (size_t)((P.n + kTileRows - 1) / kTileRows) * (kTileRows / 8) : 0;
  auto ls = std::make_shared<LaunchShared>();
  ls->core = core;
  // select next to the previous launch set's gather (launch_set): only when the input was complete, in ctx-stream
  // order, before that launch set was enqueued, and with a workspace block nothing younger has touched
  const int64_t seq = core->launch_seq.load();
  bool overlap = false;
  if (split && overlap_enabled() && seq >= 1 && in_orig->born >= 0 && in_orig->born <= seq - 1) {
    ls->workspace = dev_alloc_released_by(core, ws_total + sel_bytes, seq - 1);
    overlap = ls->workspace != nullptr;
  }
  if (!ls->workspace) ls->workspace = dev_alloc(core, ws_total + sel_bytes);
  ls->host = core->host_get((size_t)(P.n_counts + 1) * 8, &ls->host_cls);
  ls->done = core->event_get();
  uint8_t* ws = (uint8_t*)ls->workspace->ptr;
  KernelParams kp;
  std::memset(&kp, 0, sizeof(kp));
  bind_workspace(P, ls, ws, ws_counts, kp.b, (uint64_t*)ls->host, 0);
  if (split) kp.b.selbits = (uint32_t*)(ws + ws_total);
  hc.lap(1);
  TilePlan tp;
  const int ctas = fill_program_params(p, P, kp, tp, false);
  if (split) set_split_flags(kp);
  hc.lap(2);
  launch_set(core, p, kp, tp, ctas, (unsigned)kp.b.num_tiles, ws, ws_total, P.n, ls, overlap);
  P.out->born = core->launch_seq.load();   // complete once this launch set's last kernel is
  hc.lap(3);
  after_launch(P);
  hc.lap(4);
  return std::move(P.out);
}

// One launch set over many batches of one schema (the reference's native 10 000-row records,
// physical_planner.rs:323): outputs stay one batch per input, so the record_id <-> rec_<id>.parquet mapping
// (materialize_files_task.rs:119) is unchanged.  Batches whose validity layout differs from the first one's, and
// batches that need no kernel, are run one by one.
static void execute_many(chdb_ctx* ctx, const Program& p, const chdb_device_batch* const* ins, int32_t count, chdb_device_batch** outs) {
  const Core& core = ctx->core;
  CUDA_CHECK(cudaSetDevice(core->device));
  std::vector<std::unique_ptr<chdb_device_batch>> done((size_t)count);
  // one slab for every batch's output buffers
  size_t cap = 0;
  for (int32_t i = 0; i < count; i++) {
    if (!ins[i]) throw Error(CHDB_ERR_INVALID_ARGUMENT, "null batch");
    check_schema(p, ins[i]);
    if (ins[i]->num_rows < 0) resolve(const_cast<chdb_device_batch*>(ins[i]));
    for (auto& c : ins[i]->cols)
      if (c.meta.type == T_UTF8) utf8_value_bytes(c, ins[i]->num_rows);
    cap += output_capacity(p, ins[i]);
  }
  Slab slab;
  slab.buf = dev_alloc(core, cap);
  std::vector<Prepared> preps((size_t)count);
  std::vector<int32_t> group;
  uint64_t shape = 0;
  int64_t total_rows = 0;
  for (int32_t i = 0; i < count; i++) {
    Prepared& P = preps[(size_t)i];
    prepare(ctx, p, ins[i], &slab, P);
    if (!P.launch) { done[(size_t)i] = std::move(P.out); continue; }
    if (group.empty()) shape = P.shape;
    if (P.shape != shape || P.any_const) {   // another validity layout (or literal columns): not part of the shared launch
      done[(size_t)i] = execute(ctx, p, ins[i]);
      continue;
    }
    group.push_back(i);
    total_rows += P.n;
  }
  if (!group.empty()) {
    const int nb = (int)group.size();
    const Prepared& P0 = preps[(size_t)group[0]];
    const int n_in = (int)p.slot_to_col.size(), n_out = P0.ko, per = P0.n_counts + 1;
    const size_t stride = sizeof(BatchHeader) + (size_t)n_in * sizeof(ColumnDesc) + (size_t)n_out * sizeof(OutDesc);
    // workspace, per batch: counts, done, descriptors, bit-packed outputs
    std::vector<size_t> ws_at((size_t)nb), ws_counts((size_t)nb);
    size_t ws_total = 0;
    int64_t tiles = 0;
    for (int g = 0; g < nb; g++) {
      Prepared& P = preps[(size_t)group[(size_t)g]];
      size_t wd;
      ws_at[(size_t)g] = ws_total;
      ws_total += workspace_layout(P, &ws_counts[(size_t)g], &wd);
      tiles += (P.n + kTileRows - 1) / kTileRows;
    }
    if (tiles > INT32_MAX) throw Error(CHDB_ERR_NOT_IMPLEMENTED, "too many tiles in one launch");
    auto ls = std::make_shared<LaunchShared>();
    ls->core = core;
    ls->workspace = dev_alloc(core, ws_total);
    const size_t rec_bytes = round_up((size_t)nb * stride, 16), tab_bytes = round_up((size_t)tiles * 4, 16);
    const size_t host_counts_bytes = round_up((size_t)nb * (size_t)per * 8, 16);
    ls->host = core->host_get(host_counts_bytes + rec_bytes + tab_bytes, &ls->host_cls);
    ls->params = dev_alloc(core, rec_bytes + tab_bytes);
    ls->done = core->event_get();
    uint8_t* ws = (uint8_t*)ls->workspace->ptr;
    uint8_t* stage = (uint8_t*)ls->host + host_counts_bytes;   // pinned staging of the records and the tile table
    int32_t* tab = (int32_t*)(stage + rec_bytes);
    int32_t first_tile = 0;
    for (int g = 0; g < nb; g++) {
      Prepared& P = preps[(size_t)group[(size_t)g]];
      uint8_t* rec = stage + (size_t)g * stride;
      BatchHeader bh;
      bind_workspace(P, ls, ws + ws_at[(size_t)g], ws_counts[(size_t)g], bh, (uint64_t*)ls->host + (size_t)g * (size_t)per, first_tile);
      std::memcpy(rec, &bh, sizeof(bh));
      std::memcpy(rec + sizeof(bh), P.in_desc, (size_t)n_in * sizeof(ColumnDesc));
      std::memcpy(rec + sizeof(bh) + (size_t)n_in * sizeof(ColumnDesc), P.out_desc, (size_t)n_out * sizeof(OutDesc));
      for (int32_t t = 0; t < bh.num_tiles; t++) tab[first_tile + t] = g;
      first_tile += bh.num_tiles;
    }
    CUDA_CHECK(cudaMemcpyAsync(ls->params->ptr, stage, rec_bytes + tab_bytes, cudaMemcpyHostToDevice, core->stream));
    KernelParams kp;
    std::memset(&kp, 0, sizeof(kp));
    std::memcpy(&kp.b, stage, sizeof(BatchHeader));   // (batch 0's; the kernel reads every batch's own from its record)
    kp.many = (const uint8_t*)ls->params->ptr;
    kp.many_tile_batch = (const int32_t*)((const uint8_t*)ls->params->ptr + rec_bytes);
    kp.many_batches = nb;
    kp.many_stride = (int32_t)stride;
    TilePlan tp;
    Prepared& F = preps[(size_t)group[0]];
    const int ctas = fill_program_params(p, F, kp, tp, true);
    launch_set(core, p, kp, tp, ctas, (unsigned)tiles, ws, ws_total, total_rows, ls);
    for (int g = 0; g < nb; g++) done[(size_t)group[(size_t)g]] = std::move(preps[(size_t)group[(size_t)g]].out);
  }
  for (int32_t i = 0; i < count; i++) {
    done[(size_t)i]->born = core->launch_seq.load();
    outs[i] = done[(size_t)i].release();
  }
}

// Kernel parameter block with the *shape* execute() would produce for a batch whose nullable
// columns all carry validity bitmaps (pointers stay null): enough to build the specialised source.
static void shape_params(const Program& p, KernelParams& kp) {
  std::memset(&kp, 0, sizeof(kp));
  const bool compact = p.has_pred && !p.pred_const;
  kp.n_in = (int)p.slot_to_col.size();
  for (size_t s = 0; s < p.slot_to_col.size(); s++) {
    kp.in[s].type = p.schema[p.slot_to_col[s]].type;
    kp.in[s].width = (uint8_t)p.schema[p.slot_to_col[s]].width;
  }
  auto slot_nullable = [&](int slot) { return (p.schema[p.slot_to_col[slot]].flags & ARROW_FLAG_NULLABLE) != 0; };
  int ko = 0, n_utf8 = 0, n_counts = 1;
  for (auto& o : p.outputs)
    if ((o.kind == OutputColumn::EXPR || (o.kind == OutputColumn::PASS && compact)) && o.type == T_UTF8) n_utf8++;
  n_counts = 1 + n_utf8;
  int utf8_seen = 0;
  for (auto& o : p.outputs) {
    if (!(o.kind == OutputColumn::EXPR || (o.kind == OutputColumn::PASS && compact))) continue;
    OutDesc& od = kp.out[ko++];
    od.kind = o.kind == OutputColumn::EXPR ? OUT_EXPR : OUT_PASS;
    od.type = o.type;
    od.width = (uint8_t)o.width;
    od.slot = (uint8_t)(o.slot >= 0 ? o.slot : 0);
    od.begin = (uint8_t)o.begin;
    od.end = (uint8_t)o.end;
    od.utf8_index = o.type == T_UTF8 ? (uint8_t)utf8_seen++ : 0xFF;
    bool nullable = false;
    if (o.kind == OutputColumn::PASS) nullable = (p.schema[o.in_col].flags & ARROW_FLAG_NULLABLE) != 0;
    else
      for (int i = o.begin; i < o.end; i++)
        if (p.instrs[i].src == SRC_COL && slot_nullable(p.instrs[i].slot)) nullable = true;
    if (nullable) od.count_index = (uint8_t)n_counts++;
  }
  kp.n_out = ko;
  kp.n_utf8 = n_utf8;
  kp.pred_begin = compact ? p.pred_begin : 0;
  kp.pred_end = compact ? p.pred_end : 0;
  std::memcpy(kp.instrs, p.instrs.data(), p.instrs.size() * sizeof(Instr));
  std::memcpy(kp.strpool, p.strpool.data(), p.strpool.size());
  if (compact) set_split_flags(kp);   // (the shape large batches run with)
}

static_assert(kErrArithmeticOverflow == CHDB_ERR_ARITHMETIC_OVERFLOW && kErrDivideByZero == CHDB_ERR_DIVIDE_BY_ZERO,
              "device error codes must match include/chdb_gpu.h");

}  // namespace chdb

// ==========================================================================================
// C ABI
// ==========================================================================================
using namespace chdb;

extern "C" {

const char* chdb_code_name(int32_t code) {
  static const char* names[] = {"Ok", "ValueTypeNotImplemented", "ExpressionTypeNotImplemented", "BinaryOperatorNotImplemented",
                                "BinaryOperatinCastFailed", "FailedToParseAsAnInteger", "FailedToParseAsAFloat", "ColumnNotFound",
                                "IdentifierNotFound", "UnsupportedTypeCoersion", "CastToBooleanArrayFailedForArrayType",
                                "NotImplemented", "ArithmeticOverflow", "DivideByZero", "ComputeError", "InvalidArgumentError",
                                "Cuda", "BadJson", "Panic"};
  return code >= 0 && code <= CHDB_ERR_PANIC ? names[code] : "Unknown";
}
const char* chdb_version(void) { return "chdb-gpu 0.1.0"; }
const char* chdb_compiled_arch(void) { return "sm_100a"; }
uint32_t chdb_set_sql_extensions(uint32_t mask) { return chdb::g_sql_extensions.exchange(mask & 3u); }
uint32_t chdb_get_sql_extensions(void) { return chdb::g_sql_extensions.load(); }

int32_t chdb_ctx_create(int32_t device, chdb_ctx** out, chdb_status* st) {
  return guarded(st, [&] {
    if (!out) throw Error(CHDB_ERR_INVALID_ARGUMENT, "out is null");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
      throw Error(CHDB_ERR_CUDA, std::string("no CUDA device available (this library has no CPU fallback): ") +
                                     (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    if (device < 0 || device >= count) throw Error(CHDB_ERR_INVALID_ARGUMENT, "device index out of range");
    CUDA_CHECK(cudaSetDevice(device));
    auto core = std::make_shared<CtxCore>();
    core->device = device;
    CUDA_CHECK(cudaStreamCreateWithFlags(&core->stream, cudaStreamNonBlocking));
    int sms = 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) core->sm_count = sms;
    cudaMemPool_t pool;
    CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t threshold = UINT64_MAX;   // keep freed blocks cached: steady-state batches never hit cudaMalloc
    CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
    std::unique_ptr<chdb_ctx> c(new chdb_ctx);
    c->core = core;
    *out = c.release();
  });
}
void chdb_ctx_destroy(chdb_ctx* ctx) { delete ctx; }
void* chdb_ctx_stream(chdb_ctx* ctx) { return ctx ? (void*)ctx->core->stream : nullptr; }
int32_t chdb_ctx_device(chdb_ctx* ctx) { return ctx ? ctx->core->device : -1; }
int32_t chdb_ctx_synchronize(chdb_ctx* ctx, chdb_status* st) {
  return guarded(st, [&] {
    if (!ctx) throw Error(CHDB_ERR_INVALID_ARGUMENT, "ctx is null");
    CUDA_CHECK(cudaSetDevice(ctx->core->device));
    CUDA_CHECK(cudaStreamSynchronize(ctx->core->stream));
  });
}
int64_t chdb_ctx_launch_count(chdb_ctx* ctx) { return ctx ? ctx->core->launches.load() : 0; }
int64_t chdb_ctx_jit_launch_count(chdb_ctx* ctx) { return ctx ? ctx->core->jit_launches.load() : 0; }
int64_t chdb_ctx_alloc_miss_count(chdb_ctx* ctx) { return ctx ? ctx->core->alloc_misses.load() : 0; }
int64_t chdb_ctx_overlapped_count(chdb_ctx* ctx) { return ctx ? ctx->core->overlapped.load() : 0; }

int32_t chdb_jit_available(char* why, size_t cap) {
  std::string reason;
  const bool ok = jit_available(&reason) && jit_mode() != JitMode::Never;
  if (why && cap) std::snprintf(why, cap, "%s", ok ? "" : (reason.empty() ? "disabled by CHDB_JIT" : reason.c_str()));
  return ok ? 1 : 0;
}
size_t chdb_program_jit_source(const chdb_program* prog, char* buf, size_t cap) {
  if (!prog) return 0;
  KernelParams kp;
  shape_params(*prog->p, kp);
  std::string s = jit_prologue(kp, prog->p->has64, 2);
  if (buf && cap) {
    size_t n = std::min(cap - 1, s.size());
    std::memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return s.size();
}
int32_t chdb_program_jit_check(const chdb_program* prog, int64_t* cubin_bytes, char* log, size_t cap, chdb_status* st) {
  return guarded(st, [&] {
    if (!prog) throw Error(CHDB_ERR_INVALID_ARGUMENT, "program is null");
    KernelParams kp;
    shape_params(*prog->p, kp);
    std::string l;
    std::vector<char> cubin = jit_compile_offline(kp, prog->p->has64, &l);
    if (cubin_bytes) *cubin_bytes = (int64_t)cubin.size();
    if (const char* dump = std::getenv("CHDB_JIT_DUMP")) {   // for cuobjdump -sass (profiles/)
      if (FILE* f = std::fopen(dump, "wb")) { std::fwrite(cubin.data(), 1, cubin.size(), f); std::fclose(f); }
    }
    if (log && cap) std::snprintf(log, cap, "%s", l.c_str());
  });
}

static int32_t compile_into(chdb_program** out, chdb_status* st, Program::Mode mode, const char* expr, const char* items,
                            const struct ArrowSchema* schema, const char* aliases) {
  return guarded(st, [&] {
    if (!out) throw Error(CHDB_ERR_INVALID_ARGUMENT, "out is null");
    *out = nullptr;
    std::unique_ptr<chdb_program> h(new chdb_program);
    h->p = compile_program(mode, expr, items, schema, aliases);
    *out = h.release();
  });
}
int32_t chdb_program_compile_filter(const char* expr_json, const struct ArrowSchema* in_schema, const char* table_aliases_json,
                                    chdb_program** out, chdb_status* st) {
  return compile_into(out, st, Program::FILTER, expr_json, nullptr, in_schema, table_aliases_json);
}
int32_t chdb_program_compile_project(const char* select_items_json, const struct ArrowSchema* in_schema,
                                     const char* table_aliases_json, chdb_program** out, chdb_status* st) {
  return compile_into(out, st, Program::PROJECT, nullptr, select_items_json, in_schema, table_aliases_json);
}
int32_t chdb_program_compile_filter_project(const char* expr_json, const char* select_items_json,
                                            const struct ArrowSchema* in_schema, const char* table_aliases_json,
                                            chdb_program** out, chdb_status* st) {
  return compile_into(out, st, Program::FILTER_PROJECT, expr_json, select_items_json, in_schema, table_aliases_json);
}
void chdb_program_release(chdb_program* prog) { delete prog; }
size_t chdb_program_disassemble(const chdb_program* prog, char* buf, size_t cap) {
  if (!prog) return 0;
  std::string s = prog->p->disassemble();
  if (buf && cap) {
    size_t n = std::min(cap - 1, s.size());
    std::memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return s.size();
}
int32_t chdb_program_num_instructions(const chdb_program* prog) { return prog ? (int32_t)prog->p->instrs.size() : 0; }

int32_t chdb_upload(chdb_ctx* ctx, const struct ArrowArray* in, const struct ArrowSchema* in_schema, chdb_device_batch** out,
                    chdb_status* st) {
  return guarded(st, [&] {
    if (!ctx || !out) throw Error(CHDB_ERR_INVALID_ARGUMENT, "ctx / out is null");
    *out = nullptr;
    auto b = upload_batch(ctx->core, in, in_schema);
    CUDA_CHECK(cudaStreamSynchronize(ctx->core->stream));  // inputs are only borrowed for the call
    *out = b.release();
  });
}

int32_t chdb_device_batch_wrap(chdb_ctx* ctx, const struct ArrowSchema* schema, int64_t num_rows, const void* const* values,
                               const void* const* validity, const void* const* offsets, chdb_device_batch** out,
                               chdb_status* st) {
  return guarded(st, [&] {
    if (!ctx || !out || !values) throw Error(CHDB_ERR_INVALID_ARGUMENT, "ctx / out / values is null");
    *out = nullptr;
    std::vector<InputColumn> cols = parse_schema(schema);
    std::unique_ptr<chdb_device_batch> b(new chdb_device_batch);
    b->core = ctx->core;
    b->born = ctx->core->launch_seq.load();   // (the caller's buffers are ready in ctx-stream order)
    b->num_rows = num_rows;
    for (size_t i = 0; i < cols.size(); i++) {
      DeviceColumn dc;
      dc.meta = cols[i];
      if (!dc.meta.supported) throw Error(CHDB_ERR_NOT_IMPLEMENTED, "column '" + dc.meta.name + "': unsupported Arrow format");
      dc.values = values[i];
      dc.validity = validity ? (const uint8_t*)validity[i] : nullptr;
      dc.offsets = offsets ? (const int32_t*)offsets[i] : nullptr;
      if (((uintptr_t)dc.values | (uintptr_t)dc.validity | (uintptr_t)dc.offsets) & 15u)
        throw Error(CHDB_ERR_INVALID_ARGUMENT, "wrapped device buffers must be 16-byte aligned");
      dc.null_count = dc.validity ? -2 : 0;
      if (dc.meta.type == T_UTF8) {
        if (!dc.offsets) throw Error(CHDB_ERR_INVALID_ARGUMENT, "Utf8 column without offsets");
        dc.first_offset = -3;   // learnt from the offsets on first use (utf8_value_bytes): wrapping never touches the device
        dc.value_bytes = -3;
      }
      b->cols.push_back(std::move(dc));
    }
    *out = b.release();
  });
}

int32_t chdb_run_device(chdb_ctx* ctx, const chdb_program* prog, const chdb_device_batch* in, chdb_device_batch** out,
                        chdb_status* st) {
  return guarded(st, [&] {
    if (!ctx || !prog || !in || !out) throw Error(CHDB_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    *out = execute(ctx, *prog->p, in).release();
  });
}

int32_t chdb_run_device_many(chdb_ctx* ctx, const chdb_program* prog, const chdb_device_batch* const* in, int32_t count,
                             chdb_device_batch** out, chdb_status* st) {
  return guarded(st, [&] {
    if (!ctx || !prog || (count > 0 && (!in || !out)) || count < 0) throw Error(CHDB_ERR_INVALID_ARGUMENT, "null argument");
    for (int32_t i = 0; i < count; i++) out[i] = nullptr;
    if (count == 0) return;
    execute_many(ctx, *prog->p, in, count, out);
  });
}

int32_t chdb_device_batch_ready(chdb_ctx* ctx, const chdb_device_batch* b, chdb_status* st) {
  (void)ctx;
  int32_t ready = -1;
  guarded(st, [&] {
    if (!b) throw Error(CHDB_ERR_INVALID_ARGUMENT, "batch is null");
    ready = (!b->result || b->result->launch->ready()) ? 1 : 0;
  });
  return ready;
}

int32_t chdb_device_batch_status(chdb_ctx* ctx, const chdb_device_batch* b, chdb_status* st) {
  (void)ctx;
  return guarded(st, [&] {
    if (!b) throw Error(CHDB_ERR_INVALID_ARGUMENT, "batch is null");
    check_run_error(b);
  });
}
int64_t chdb_device_batch_num_rows(chdb_ctx* ctx, const chdb_device_batch* b, chdb_status* st) {
  (void)ctx;
  int64_t n = -1;
  guarded(st, [&] {
    if (!b) throw Error(CHDB_ERR_INVALID_ARGUMENT, "batch is null");
    resolve(const_cast<chdb_device_batch*>(b));
    n = b->num_rows;
  });
  return n;
}
int32_t chdb_device_batch_num_columns(const chdb_device_batch* b) { return b ? (int32_t)b->cols.size() : 0; }
int32_t chdb_device_batch_column(chdb_ctx* ctx, const chdb_device_batch* b, int32_t col, const void** values, int64_t* values_bytes,
                                 const void** validity, int64_t* validity_bytes, const void** offsets, int64_t* offsets_bytes,
                                 chdb_status* st) {
  (void)ctx;
  return guarded(st, [&] {
    if (!b || col < 0 || col >= (int32_t)b->cols.size()) throw Error(CHDB_ERR_INVALID_ARGUMENT, "bad batch / column index");
    check_run_error(b);
    resolve(const_cast<chdb_device_batch*>(b));
    const DeviceColumn& c = b->cols[col];
    const int64_t n = b->num_rows;
    const bool has_validity = c.validity != nullptr && c.null_count != 0;
    if (validity) *validity = has_validity ? c.validity : nullptr;
    if (validity_bytes) *validity_bytes = has_validity ? (int64_t)bitmap_bytes(n) : 0;
    if (offsets) *offsets = c.meta.type == T_UTF8 ? c.offsets : nullptr;
    if (offsets_bytes) *offsets_bytes = c.meta.type == T_UTF8 ? (n + 1) * 4 : 0;
    if (c.meta.type == T_UTF8) {
      CUDA_CHECK(cudaSetDevice(b->core->device));
      if (utf8_value_bytes(c, n) < 0 || c.first_offset < 0) throw Error(CHDB_ERR_NOT_IMPLEMENTED, "column buffers of a sliced Utf8 view");
      if (values) *values = (const uint8_t*)c.values + c.first_offset;
      if (values_bytes) *values_bytes = c.value_bytes;
    } else {
      if (values) *values = c.values;
      if (values_bytes) *values_bytes = c.meta.type == T_BOOL ? (int64_t)bitmap_bytes(n) : n * c.meta.width;
    }
  });
}
int64_t chdb_device_batch_nbytes(chdb_ctx* ctx, const chdb_device_batch* b, chdb_status* st) {
  (void)ctx;
  int64_t total = -1;
  guarded(st, [&] {
    if (!b) throw Error(CHDB_ERR_INVALID_ARGUMENT, "batch is null");
    resolve(const_cast<chdb_device_batch*>(b));
    int64_t t = 0;
    const int64_t n = b->num_rows;
    for (auto& c : b->cols) {
      if (c.validity) t += (int64_t)bitmap_bytes(n);
      if (c.meta.type == T_UTF8) t += (n + 1) * 4 + std::max<int64_t>(utf8_value_bytes(c, n), 0);
      else if (c.meta.type == T_BOOL) t += (int64_t)bitmap_bytes(n);
      else t += n * c.meta.width;
    }
    total = t;
  });
  return total;
}
int32_t chdb_download(chdb_ctx* ctx, const chdb_device_batch* b, struct ArrowArray* out, struct ArrowSchema* out_schema,
                      chdb_status* st) {
  (void)ctx;
  return guarded(st, [&] {
    if (!b || !out || !out_schema) throw Error(CHDB_ERR_INVALID_ARGUMENT, "null argument");
    download_batch(const_cast<chdb_device_batch*>(b), out, out_schema);
  });
}
void chdb_device_batch_retain(chdb_device_batch* b) {
  if (b) b->refs.fetch_add(1, std::memory_order_relaxed);
}
void chdb_device_batch_release(chdb_device_batch* b) {
  if (b && b->refs.fetch_sub(1, std::memory_order_acq_rel) == 1) delete b;
}
void chdb_device_batch_release_many(chdb_device_batch* const* batches, int32_t count) {
  if (!batches) return;
  for (int32_t i = 0; i < count; i++) chdb_device_batch_release(batches[i]);
}

// ---- device-resident record pool ---------------------------------------------------------------
// What a GPU-aware exchange keeps instead of `RecordPool.records: HashMap<u64, Arc<RecordBatch>>`
// (exchange_operator.rs:566-777): records stay in HBM between read_files, filter and materialize, are handed out
// by reference, dropped once every consumer operator has completed them (:727-733), and -- the memory management
// the reference lists as a TODO (DEV_NOTES.md:133-140) -- spilled to pinned host memory, least recently used
// first, when the pool holds more than its byte budget; a spilled record is uploaded again when it is asked for.
struct chdb_record_pool {
  struct Entry {
    chdb_device_batch* dev = nullptr;          // resident copy (the pool's own reference), or nullptr when spilled
    struct ArrowArray host_array;              // spilled copy (release != nullptr when present)
    struct ArrowSchema host_schema;
    int32_t consumers = 1;
    int64_t held = 0;                          // device bytes this record keeps alive
    uint64_t tick = 0;                         // last add / get
  };
  chdb_ctx* ctx = nullptr;
  int64_t budget = 0;
  std::mutex mu;
  std::map<uint64_t, Entry> records;
  std::map<const DevBuf*, int> bufs;           // device blocks held by resident records (shared blocks count once)
  int64_t device_bytes = 0, spilled_bytes = 0, spilled_records = 0;
  uint64_t clock = 0;
};

namespace chdb {
static void pool_account(chdb_record_pool* p, const chdb_device_batch* b, int sign, int64_t* held) {
  int64_t delta = 0;
  auto touch = [&](const Buf& buf) {
    if (!buf) return;
    int& n = p->bufs[buf.get()];
    if (sign > 0) { if (n++ == 0) delta += (int64_t)buf->bytes; }
    else if (--n == 0) { delta += (int64_t)buf->bytes; p->bufs.erase(buf.get()); }
  };
  for (auto& c : b->cols) { touch(c.values_buf); touch(c.validity_buf); touch(c.offsets_buf); }
  p->device_bytes += sign * delta;
  if (held) *held = delta;
}
static void pool_drop_host(chdb_record_pool::Entry& e) {
  if (e.host_array.release) e.host_array.release(&e.host_array);
  if (e.host_schema.release) e.host_schema.release(&e.host_schema);
  e.host_array.release = nullptr;
  e.host_schema.release = nullptr;
}
// Spills least-recently-used records nobody else references until the pool is within budget (or nothing is left to spill).
static void pool_enforce_budget(chdb_record_pool* p, uint64_t keep_id) {
  while (p->budget > 0 && p->device_bytes > p->budget) {
    chdb_record_pool::Entry* victim = nullptr;
    for (auto& kv : p->records) {
      chdb_record_pool::Entry& e = kv.second;
      if (!e.dev || kv.first == keep_id || e.dev->refs.load() > 1) continue;   // in use by a consumer: stays
      if (!victim || e.tick < victim->tick) victim = &e;
    }
    if (!victim) return;
    download_batch(victim->dev, &victim->host_array, &victim->host_schema);
    int64_t freed = 0;
    pool_account(p, victim->dev, -1, &freed);
    p->spilled_bytes += freed;
    p->spilled_records++;
    chdb_device_batch_release(victim->dev);
    victim->dev = nullptr;
  }
}
}  // namespace chdb

int32_t chdb_record_pool_create(chdb_ctx* ctx, int64_t budget_bytes, chdb_record_pool** out, chdb_status* st) {
  return guarded(st, [&] {
    if (!ctx || !out) throw Error(CHDB_ERR_INVALID_ARGUMENT, "ctx / out is null");
    auto* p = new chdb_record_pool;
    p->ctx = ctx;
    p->budget = budget_bytes;
    *out = p;
  });
}
void chdb_record_pool_destroy(chdb_record_pool* pool) {
  if (!pool) return;
  for (auto& kv : pool->records) {
    if (kv.second.dev) chdb_device_batch_release(kv.second.dev);
    pool_drop_host(kv.second);
  }
  delete pool;
}
int32_t chdb_record_pool_add(chdb_record_pool* pool, uint64_t record_id, chdb_device_batch* batch, int32_t consumers, chdb_status* st) {
  return guarded(st, [&] {
    if (!pool || !batch) throw Error(CHDB_ERR_INVALID_ARGUMENT, "pool / batch is null");
    if (batch->core->device != pool->ctx->core->device) throw Error(CHDB_ERR_INVALID_ARGUMENT, "batch lives on another device than the pool");
    std::lock_guard<std::mutex> g(pool->mu);
    if (pool->records.count(record_id)) throw Error(CHDB_ERR_INVALID_ARGUMENT, "record id already in the pool");
    chdb_record_pool::Entry& e = pool->records[record_id];
    std::memset(&e.host_array, 0, sizeof(e.host_array));
    std::memset(&e.host_schema, 0, sizeof(e.host_schema));
    chdb_device_batch_retain(batch);
    e.dev = batch;
    e.consumers = std::max(consumers, 1);
    e.tick = ++pool->clock;
    pool_account(pool, batch, +1, &e.held);
    pool_enforce_budget(pool, record_id);
  });
}
int32_t chdb_record_pool_get(chdb_record_pool* pool, uint64_t record_id, chdb_device_batch** out, chdb_status* st) {
  return guarded(st, [&] {
    if (!pool || !out) throw Error(CHDB_ERR_INVALID_ARGUMENT, "pool / out is null");
    *out = nullptr;
    std::lock_guard<std::mutex> g(pool->mu);
    auto it = pool->records.find(record_id);
    if (it == pool->records.end()) throw Error(CHDB_ERR_INVALID_ARGUMENT, "record id not in the pool");
    chdb_record_pool::Entry& e = it->second;
    e.tick = ++pool->clock;
    if (!e.dev) {   // spilled: bring it back
      CUDA_CHECK(cudaSetDevice(pool->ctx->core->device));
      auto b = upload_batch(pool->ctx->core, &e.host_array, &e.host_schema);
      CUDA_CHECK(cudaStreamSynchronize(pool->ctx->core->stream));   // the host copy is released next
      pool_drop_host(e);
      e.dev = b.release();
      pool_account(pool, e.dev, +1, &e.held);
      pool->spilled_records--;
      pool_enforce_budget(pool, record_id);
    }
    chdb_device_batch_retain(e.dev);
    *out = e.dev;
  });
}
int32_t chdb_record_pool_complete(chdb_record_pool* pool, uint64_t record_id, chdb_status* st) {
  return guarded(st, [&] {
    if (!pool) throw Error(CHDB_ERR_INVALID_ARGUMENT, "pool is null");
    std::lock_guard<std::mutex> g(pool->mu);
    auto it = pool->records.find(record_id);
    if (it == pool->records.end()) throw Error(CHDB_ERR_INVALID_ARGUMENT, "record id not in the pool");
    chdb_record_pool::Entry& e = it->second;
    if (--e.consumers > 0) return;
    if (e.dev) {
      pool_account(pool, e.dev, -1, nullptr);
      chdb_device_batch_release(e.dev);
    } else {
      pool->spilled_records--;
    }
    pool_drop_host(e);
    pool->records.erase(it);
  });
}
void chdb_record_pool_stats(chdb_record_pool* pool, int64_t* records, int64_t* device_bytes, int64_t* spilled_records,
                            int64_t* spilled_bytes) {
  if (!pool) return;
  std::lock_guard<std::mutex> g(pool->mu);
  if (records) *records = (int64_t)pool->records.size();
  if (device_bytes) *device_bytes = pool->device_bytes;
  if (spilled_records) *spilled_records = pool->spilled_records;
  if (spilled_bytes) *spilled_bytes = pool->spilled_bytes;
}

int32_t chdb_device_batches_pack(chdb_ctx* ctx, const chdb_device_batch* const* batches, int32_t count, void* dst, int64_t capacity,
                                 int64_t* sizes_out, int64_t* total_bytes, chdb_status* st) {
  return guarded(st, [&] {
    if (!ctx || (count > 0 && !batches) || count < 0) throw Error(CHDB_ERR_INVALID_ARGUMENT, "null argument");
    CUDA_CHECK(cudaSetDevice(ctx->core->device));
    int64_t at = 0, si = 0;
    auto piece = [&](const void* p, int64_t bytes) {
      if (sizes_out) sizes_out[si] = bytes;
      si++;
      if (bytes <= 0) return;
      if (dst) {
        if (at + bytes > capacity) throw Error(CHDB_ERR_INVALID_ARGUMENT, "pack: destination too small");
        CUDA_CHECK(cudaMemcpyAsync((uint8_t*)dst + at, p, (size_t)bytes, cudaMemcpyDeviceToDevice, ctx->core->stream));
      }
      at += (bytes + 255) / 256 * 256;
    };
    for (int32_t i = 0; i < count; i++) {
      chdb_device_batch* b = const_cast<chdb_device_batch*>(batches[i]);
      if (!b) throw Error(CHDB_ERR_INVALID_ARGUMENT, "null batch");
      if (b->core->device != ctx->core->device) throw Error(CHDB_ERR_INVALID_ARGUMENT, "batch lives on another device than the ctx");
      check_run_error(b);
      resolve(b);
      const int64_t n = b->num_rows;
      for (auto& c : b->cols) {
        const bool has_validity = c.validity != nullptr && c.null_count != 0;
        piece(c.validity, has_validity ? (int64_t)bitmap_bytes(n) : 0);
        if (c.meta.type == T_UTF8) {
          if (utf8_value_bytes(c, n) < 0 || c.first_offset < 0) throw Error(CHDB_ERR_NOT_IMPLEMENTED, "pack of a sliced Utf8 view");
          piece(c.offsets, (n + 1) * 4);
          piece((const uint8_t*)c.values + c.first_offset, c.value_bytes);
        } else {
          piece(nullptr, 0);
          piece(c.values, c.meta.type == T_BOOL ? (int64_t)bitmap_bytes(n) : n * c.meta.width);
        }
      }
    }
    if (total_bytes) *total_bytes = at;
  });
}

int32_t chdb_peer_copy(chdb_ctx* dst_ctx, chdb_ctx* src_ctx, const chdb_device_batch* src, chdb_device_batch** out,
                       chdb_status* st) {
  return guarded(st, [&] {
    if (!dst_ctx || !src_ctx || !src || !out) throw Error(CHDB_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    check_run_error(src);
    resolve(const_cast<chdb_device_batch*>(src));
    const Core& dc = dst_ctx->core;
    const Core& sc = src->core;
    const int sdev = sc->device;
    CUDA_CHECK(cudaSetDevice(dc->device));
    if (sdev != dc->device) {   // direct NVLink path (without peer access the copy is staged through the host)
      int can = 0;
      CUDA_CHECK(cudaDeviceCanAccessPeer(&can, dc->device, sdev));
      if (can) {
        const cudaError_t pe = cudaDeviceEnablePeerAccess(sdev, 0);
        if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) CUDA_CHECK(pe);
        cudaGetLastError();
      }
    }
    // the copies read what the source ctx's stream produced: order them behind it
    {
      CUDA_CHECK(cudaSetDevice(sdev));
      cudaEvent_t ev = sc->event_get();
      CUDA_CHECK(cudaEventRecord(ev, sc->stream));
      CUDA_CHECK(cudaSetDevice(dc->device));
      CUDA_CHECK(cudaStreamWaitEvent(dc->stream, ev, 0));
      sc->event_put(ev);
    }
    const int64_t n = src->num_rows;
    std::unique_ptr<chdb_device_batch> b(new chdb_device_batch);
    b->core = dc;
    b->born = dc->launch_seq.load();
    b->num_rows = n;
    auto copy = [&](const void* p, size_t bytes) {
      Buf buf = dev_alloc(dc, bytes);
      if (bytes) CUDA_CHECK(cudaMemcpyPeerAsync(buf->ptr, dc->device, p, sdev, bytes, dc->stream));
      return buf;
    };
    for (auto& c : src->cols) {
      DeviceColumn o = c;
      o.values_buf.reset(); o.validity_buf.reset(); o.offsets_buf.reset();
      if (c.validity && c.null_count != 0) {
        o.validity_buf = copy(c.validity, bitmap_bytes(n));
        o.validity = (const uint8_t*)o.validity_buf->ptr;
      } else {
        o.validity = nullptr;
        o.null_count = 0;
      }
      if (c.meta.type == T_UTF8) {
        if (utf8_value_bytes(c, n) < 0) throw Error(CHDB_ERR_NOT_IMPLEMENTED, "peer copy of a sliced Utf8 view");
        o.offsets_buf = copy(c.offsets, (size_t)(n + 1) * 4);
        o.offsets = (const int32_t*)o.offsets_buf->ptr;
        const size_t lead = (size_t)(c.first_offset & 15);
        Buf vb = dev_alloc(dc, lead + (size_t)c.value_bytes);
        if (c.value_bytes)
          CUDA_CHECK(cudaMemcpyPeerAsync((uint8_t*)vb->ptr + lead, dc->device, (const uint8_t*)c.values + c.first_offset, sdev,
                                         (size_t)c.value_bytes, dc->stream));
        o.values_buf = vb;
        o.values = (const uint8_t*)vb->ptr + lead - c.first_offset;
      } else {
        const size_t bytes = c.meta.type == T_BOOL ? bitmap_bytes(n) : (size_t)n * (size_t)c.meta.width;
        o.values_buf = copy(c.values, bytes);
        o.values = o.values_buf->ptr;
      }
      b->cols.push_back(std::move(o));
    }
    // A released source block goes back to the source ctx's cache and may be handed to that ctx's next launch:
    // make the source stream wait until the copies have read it.
    {
      cudaEvent_t ev = dc->event_get();
      CUDA_CHECK(cudaEventRecord(ev, dc->stream));
      CUDA_CHECK(cudaSetDevice(sdev));
      CUDA_CHECK(cudaStreamWaitEvent(sc->stream, ev, 0));
      CUDA_CHECK(cudaSetDevice(dc->device));
      dc->event_put(ev);
    }
    *out = b.release();
  });
}

// ---- host batches: upload -> run -> download ------------------------------------------------
static void run_host(chdb_ctx* ctx, const Program& p, const struct ArrowArray* in, const struct ArrowSchema* in_schema,
                     struct ArrowArray* out, struct ArrowSchema* out_schema) {
  if (!ctx || !in || !in_schema || !out || !out_schema) throw Error(CHDB_ERR_INVALID_ARGUMENT, "null argument");
  auto dev_in = upload_batch(ctx->core, in, in_schema);
  auto dev_out = execute(ctx, p, dev_in.get());
  download_batch(dev_out.get(), out, out_schema);   // synchronises; inputs stay borrowed until here
}

int32_t chdb_filter_record(chdb_ctx* ctx, const chdb_program* prog, const struct ArrowArray* in, const struct ArrowSchema* in_schema,
                           struct ArrowArray* out, struct ArrowSchema* out_schema, chdb_status* st) {
  return guarded(st, [&] {
    if (!prog) throw Error(CHDB_ERR_INVALID_ARGUMENT, "program is null");
    run_host(ctx, *prog->p, in, in_schema, out, out_schema);
  });
}
int32_t chdb_project_record(chdb_ctx* ctx, const chdb_program* prog, const struct ArrowArray* in,
                            const struct ArrowSchema* in_schema, struct ArrowArray* out, struct ArrowSchema* out_schema,
                            chdb_status* st) {
  return chdb_filter_record(ctx, prog, in, in_schema, out, out_schema, st);
}
// ---- the same without parking the calling thread (filter_task.rs:86-125 runs inside a tokio task) ----
int32_t chdb_filter_record_async(chdb_ctx* ctx, const chdb_program* prog, const struct ArrowArray* in,
                                 const struct ArrowSchema* in_schema, chdb_pending** out, chdb_status* st) {
  return guarded(st, [&] {
    if (!ctx || !prog || !in || !in_schema || !out) throw Error(CHDB_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    std::unique_ptr<chdb_pending> pd(new chdb_pending);
    pd->core = ctx->core;
    pd->dev_in = upload_batch(ctx->core, in, in_schema);          // H2D copies enqueued; `in` stays borrowed until ready
    pd->dev_out = execute(ctx, *prog->p, pd->dev_in.get());       // kernels enqueued
    *out = pd.release();
  });
}
int32_t chdb_project_record_async(chdb_ctx* ctx, const chdb_program* prog, const struct ArrowArray* in,
                                  const struct ArrowSchema* in_schema, chdb_pending** out, chdb_status* st) {
  return chdb_filter_record_async(ctx, prog, in, in_schema, out, st);
}

// 1: the result can be taken with chdb_pending_result; 0: not yet (never blocks); negative codes: see *st.
int32_t chdb_poll(chdb_pending* pd, chdb_status* st) {
  int32_t state = -1;
  const int32_t rc = guarded(st, [&] {
    if (!pd) throw Error(CHDB_ERR_INVALID_ARGUMENT, "pending is null");
    CUDA_CHECK(cudaSetDevice(pd->core->device));
    if (pd->stage == 0) {   // kernels running
      if (pd->dev_out->result && !pd->dev_out->result->launch->ready()) { state = 0; return; }
      if (!pd->dev_out->result) {   // nothing was launched (shared pass-through columns): wait for the uploads only
        if (!pd->copied) { pd->copied = pd->core->event_get(); CUDA_CHECK(cudaEventRecord(pd->copied, pd->core->stream)); }
      }
      download_begin(pd->dev_out.get(), pd->dl);   // counts are known: sizes the host buffers, enqueues the D2H copies
      if (pd->copied) pd->core->event_put(pd->copied);
      pd->copied = pd->core->event_get();
      CUDA_CHECK(cudaEventRecord(pd->copied, pd->core->stream));
      pd->stage = 1;
    }
    if (pd->stage == 1) {
      const cudaError_t e = cudaEventQuery(pd->copied);
      if (e == cudaErrorNotReady) { state = 0; return; }
      CUDA_CHECK(e);
      pd->stage = 2;
    }
    state = 1;
  });
  return rc == 0 ? state : -rc;
}
int32_t chdb_pending_result(chdb_pending* pd, struct ArrowArray* out, struct ArrowSchema* out_schema, chdb_status* st) {
  return guarded(st, [&] {
    if (!pd || !out || !out_schema) throw Error(CHDB_ERR_INVALID_ARGUMENT, "null argument");
    if (pd->stage == 3) throw Error(CHDB_ERR_INVALID_ARGUMENT, "result already taken");
    CUDA_CHECK(cudaSetDevice(pd->core->device));
    if (pd->stage == 0) download_begin(pd->dev_out.get(), pd->dl);   // (blocking use: waits for the launch)
    if (pd->stage < 2) CUDA_CHECK(cudaStreamSynchronize(pd->core->stream));
    download_end(pd->dev_out.get(), pd->dl, out, out_schema);
    pd->stage = 3;
  });
}
void chdb_pending_release(chdb_pending* pd) {
  if (!pd) return;
  cudaSetDevice(pd->core->device);
  if (pd->stage < 2) cudaStreamSynchronize(pd->core->stream);   // copies into / out of caller and pool memory must not outlive this
  if (pd->copied) pd->core->event_put(pd->copied);
  delete pd;
}

int32_t chdb_filter_record_expr(chdb_ctx* ctx, const struct ArrowArray* in, const struct ArrowSchema* in_schema,
                                const char* table_aliases_json, const char* expr_json, struct ArrowArray* out,
                                struct ArrowSchema* out_schema, chdb_status* st) {
  return guarded(st, [&] {
    auto p = compile_program(Program::FILTER, expr_json, nullptr, in_schema, table_aliases_json);
    run_host(ctx, *p, in, in_schema, out, out_schema);
  });
}
int32_t chdb_project_record_items(chdb_ctx* ctx, const char* select_items_json, const struct ArrowArray* in,
                                  const struct ArrowSchema* in_schema, const char* table_aliases_json, struct ArrowArray* out,
                                  struct ArrowSchema* out_schema, chdb_status* st) {
  return guarded(st, [&] {
    auto p = compile_program(Program::PROJECT, nullptr, select_items_json, in_schema, table_aliases_json);
    run_host(ctx, *p, in, in_schema, out, out_schema);
  });
}
int32_t chdb_compute_value(chdb_ctx* ctx, const struct ArrowArray* in, const struct ArrowSchema* in_schema,
                           const char* table_aliases_json, const char* expr_json, struct ArrowArray* out,
                           struct ArrowSchema* out_schema, int32_t* is_scalar, chdb_status* st) {
  return guarded(st, [&] {
    bool scalar = false;
    auto p = compile_value(expr_json, in_schema, table_aliases_json, &scalar);
    if (is_scalar) *is_scalar = scalar ? 1 : 0;
    // compute_value has no RecordBatch::try_new step: a len-1 result stays len 1 whatever the batch length
    if (p->outputs.size() == 1 && p->outputs[0].kind == OutputColumn::CONST) {
      if (!ctx) throw Error(CHDB_ERR_INVALID_ARGUMENT, "ctx is null");
      if (p->requires_single_row && in->length != 1) throw Error(p->single_row_code, p->single_row_msg);
      CUDA_CHECK(cudaSetDevice(ctx->core->device));
      chdb_device_batch b;
      b.core = ctx->core;
      b.num_rows = 1;
      b.cols.push_back(const_column(ctx->core, p->outputs[0]));
      download_batch(&b, out, out_schema);
      return;
    }
    run_host(ctx, *p, in, in_schema, out, out_schema);
  });
}

}  // extern "C"

#include "parquet.inc"
#include "parquet_encode.inc"
