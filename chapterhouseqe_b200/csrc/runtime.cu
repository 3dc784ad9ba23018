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
#include <string>
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
  std::map<size_t, std::vector<void*>> dev_free;
  size_t dev_cached = 0;
  std::vector<cudaEvent_t> events_free;   // recycled: creating / destroying an event is a driver resource call
  static constexpr size_t kDevCacheCap = (size_t)24 << 30;   // beyond this, released blocks go back to the pool

  ~CtxCore() {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    for (auto& kv : dev_free)
      for (void* p : kv.second) cudaFreeAsync(p, stream);
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
        core->dev_free[bytes].push_back(ptr);
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
      b->ptr = it->second.back();
      it->second.pop_back();
      core->dev_cached -= b->bytes;
      return b;
    }
  }
  CUDA_CHECK(cudaMallocAsync(&b->ptr, b->bytes, core->stream));
  return b;
}

// Counts of one kernel run, read back asynchronously into pinned memory.
struct RunResult {
  Core core;
  uint64_t* host = nullptr;   // [n_counts] counts then [n_counts] = error word
  int n_counts = 0;
  cudaEvent_t done = nullptr;
  bool waited = false;
  Buf workspace;
  ~RunResult() {
    if (done) core->event_put(done);
    if (host) core->pinned_put(host);
  }
  void wait() {
    if (waited) return;
    CUDA_CHECK(cudaSetDevice(core->device));
    CUDA_CHECK(cudaEventSynchronize(done));
    waited = true;
  }
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
  chdb::Core core;
  int64_t num_rows = 0;              // -1: result->host[0]
  std::vector<chdb::DeviceColumn> cols;
  std::shared_ptr<chdb::RunResult> result;
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

static void download_batch(chdb_device_batch* b, ::ArrowArray* out, ::ArrowSchema* out_schema) {
  const Core& core = b->core;
  CUDA_CHECK(cudaSetDevice(core->device));
  check_run_error(b);
  resolve(b);
  const int64_t n = b->num_rows;
  auto* top = new ExportedArray;
  top->core = core;
  std::unique_ptr<ExportedArray> top_guard(top);
  top->child_storage.resize(b->cols.size());
  for (auto& c : top->child_storage) std::memset(&c, 0, sizeof(c));
  // pass 1: offsets (needed to size Utf8 value copies of shared / sliced columns)
  std::vector<int32_t*> host_offsets(b->cols.size(), nullptr);
  std::vector<ExportedArray*> exs(b->cols.size(), nullptr);
  for (size_t ci = 0; ci < b->cols.size(); ci++) {
    exs[ci] = new ExportedArray;
    exs[ci]->core = core;
    ::ArrowArray& a = top->child_storage[ci];
    a.private_data = exs[ci];
    a.release = release_array;
    DeviceColumn& c = b->cols[ci];
    if (c.meta.type == T_UTF8) {
      int32_t* ho = (int32_t*)host_alloc(exs[ci], (size_t)(n + 1) * 4);
      ho[0] = 0;
      if (n) CUDA_CHECK(cudaMemcpyAsync(ho, c.offsets, (size_t)(n + 1) * 4, cudaMemcpyDeviceToHost, core->stream));
      host_offsets[ci] = ho;
    }
  }
  CUDA_CHECK(cudaStreamSynchronize(core->stream));
  // pass 2: values + validity
  for (size_t ci = 0; ci < b->cols.size(); ci++) {
    DeviceColumn& c = b->cols[ci];
    ExportedArray* ex = exs[ci];
    void* hv = nullptr;
    uint8_t* hval = nullptr;
    if (c.validity != nullptr && c.null_count != 0 && n) {
      hval = (uint8_t*)host_alloc(ex, bitmap_bytes(n));
      CUDA_CHECK(cudaMemcpyAsync(hval, c.validity, bitmap_bytes(n), cudaMemcpyDeviceToHost, core->stream));
    }
    if (c.meta.type == T_UTF8) {
      int32_t* ho = host_offsets[ci];
      const int64_t first = n ? ho[0] : 0, last = n ? ho[n] : 0;
      hv = host_alloc(ex, (size_t)(last - first));
      if (last > first)
        CUDA_CHECK(cudaMemcpyAsync(hv, (const uint8_t*)c.values + first, (size_t)(last - first), cudaMemcpyDeviceToHost, core->stream));
      if (first != 0)
        for (int64_t i = 0; i <= n; i++) ho[i] -= (int32_t)first;
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
  CUDA_CHECK(cudaStreamSynchronize(core->stream));
  make_schema(out_schema, "+s", "", 0, b->cols.size());
  for (size_t ci = 0; ci < b->cols.size(); ci++) {
    DeviceColumn& c = b->cols[ci];
    ExportedArray* ex = exs[ci];
    ::ArrowArray& a = top->child_storage[ci];
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
  out->private_data = top_guard.release();
}

// ------------------------------------------------------------------------------------------
// executor
// ------------------------------------------------------------------------------------------
static std::unique_ptr<chdb_device_batch> view_rows(const chdb_device_batch* in, int64_t n) {
  std::unique_ptr<chdb_device_batch> v(new chdb_device_batch);
  v->core = in->core;
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

// CHDB_HOST_TIMING=1: where the host spends its time inside execute() (debugging aid): calls that take
// longer than 300 us report their phases on stderr.
struct HostClock {
  bool on;
  std::chrono::steady_clock::time_point t0, t;
  double ph[4] = {0, 0, 0, 0};
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
    if (total > 300) std::fprintf(stderr, "[chdb host] execute %.0f us: allocate outputs %.0f, workspace+zero %.0f, plan+launch %.0f, event %.0f\n",
                                  total, ph[0], ph[1], ph[2], ph[3]);
  }
};

static std::unique_ptr<chdb_device_batch> execute(chdb_ctx* ctx, const Program& p, const chdb_device_batch* in_orig) {
  HostClock hc;
  const Core& core = ctx->core;
  CUDA_CHECK(cudaSetDevice(core->device));
  if (in_orig->core->device != core->device)
    throw Error(CHDB_ERR_INVALID_ARGUMENT, "batch lives on another device than the ctx");
  check_schema(p, in_orig);
  if (in_orig->num_rows < 0) resolve(const_cast<chdb_device_batch*>(in_orig));
  if (in_orig->result) check_run_error(in_orig);

  const chdb_device_batch* in = in_orig;
  std::unique_ptr<chdb_device_batch> view;
  bool use_pred = p.has_pred;
  if (p.requires_single_row && in->num_rows != 1) throw Error(p.single_row_code, p.single_row_msg);
  if (p.has_pred && p.pred_const) {
    // arrow-select does not broadcast a len-1 mask: [true] keeps row 0, [false] keeps nothing
    if (in->num_rows < 1)
      throw Error(CHDB_ERR_INVALID_ARGUMENT,
                  "Invalid argument error: Filter predicate of length 1 is larger than target array of length 0");
    view = view_rows(in, p.pred_const_value ? 1 : 0);
    in = view.get();
    use_pred = false;
  }
  const int64_t n = in->num_rows;

  std::unique_ptr<chdb_device_batch> out(new chdb_device_batch);
  out->core = core;

  // len-1 constant outputs (record_projection.rs:73: RecordBatch::try_new checks equal lengths)
  bool any_const = false, all_const = !p.outputs.empty();
  for (auto& o : p.outputs) {
    any_const |= o.kind == OutputColumn::CONST;
    all_const &= o.kind == OutputColumn::CONST;
  }
  if (any_const && !all_const && !use_pred && n != 1)   // (under a predicate the surviving rows count: checked after the launch)
    throw Error(CHDB_ERR_INVALID_ARGUMENT, "Invalid argument error: all columns in a record batch must have the same length");

  // which outputs go through the kernel
  const bool compact = use_pred;
  std::vector<int> kernel_outs;
  for (size_t k = 0; k < p.outputs.size(); k++) {
    const OutputColumn& o = p.outputs[k];
    if (o.kind == OutputColumn::EXPR || (o.kind == OutputColumn::PASS && compact)) kernel_outs.push_back((int)k);
  }
  const bool launch = n > 0 && (!kernel_outs.empty() || compact);

  KernelParams kp;
  std::memset(&kp, 0, sizeof(kp));
  int n_utf8 = 0, n_counts = 1;
  if (launch) {
    for (size_t s = 0; s < p.slot_to_col.size(); s++) {
      const DeviceColumn& c = in->cols[p.slot_to_col[s]];
      kp.in[s].values = c.values;
      kp.in[s].validity = c.validity;
      kp.in[s].offsets = c.offsets;
      kp.in[s].type = c.meta.type;
      kp.in[s].width = (uint8_t)c.meta.width;
    }
    kp.n_in = (int)p.slot_to_col.size();
    for (int k : kernel_outs)
      if (p.outputs[k].type == T_UTF8) n_utf8++;
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
  int64_t avg_utf8 = -1;  // mean value length of the longest Utf8 output (sizes the staging area); -1: none
  // Bit-packed outputs (Boolean values, validity bitmaps) are merged into with atomicOr at slice edges, so they start
  // zeroed: they live, with the look-back descriptors and the counts, in ONE region that one small kernel zeroes.
  struct ZeroReq { size_t col; int ko; bool validity; size_t bytes; size_t at; };
  std::vector<ZeroReq> zero_reqs;
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
        OutDesc& od = kp.out[ko];
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
          int64_t vb = src.value_bytes;
          if (vb < 0) throw Error(CHDB_ERR_NOT_IMPLEMENTED, "filtering a sliced Utf8 view");
          avg_utf8 = std::max<int64_t>(avg_utf8, (vb + n - 1) / n);
          dc.values_buf = dev_alloc(core, (size_t)vb);
          dc.offsets_buf = dev_alloc(core, (size_t)(n + 1) * 4);
          dc.offsets = (const int32_t*)dc.offsets_buf->ptr;
          od.offsets = (int32_t*)dc.offsets_buf->ptr;
          od.utf8_index = (uint8_t)utf8_seen;
          dc.value_bytes = -1;
          dc.bytes_index = 1 + utf8_seen;
          utf8_seen++;
          dc.values = dc.values_buf->ptr;
          od.values = dc.values_buf->ptr;
        } else if (o.type == T_BOOL) {
          zero_reqs.push_back({k, ko, false, round_up(bitmap_bytes(n), 4), 0});
        } else {
          dc.values_buf = dev_alloc(core, (size_t)n * (size_t)o.width);
          dc.values = dc.values_buf->ptr;
          od.values = dc.values_buf->ptr;
        }
        if (nullable) {
          zero_reqs.push_back({k, ko, true, round_up(bitmap_bytes(n), 4), 0});
          od.count_index = (uint8_t)n_counts;
          dc.null_count = -1;
          dc.count_index = n_counts;
          n_counts++;
        }
        ko++;
      }
    }
    out->cols.push_back(std::move(dc));
  }

  if (!launch) {
    if (any_const && !all_const && compact)   // n == 0: nothing survives next to len-1 literal columns
      throw Error(CHDB_ERR_INVALID_ARGUMENT, "Invalid argument error: all columns in a record batch must have the same length");
    out->num_rows = all_const ? 1 : n;
    if (compact && !all_const) out->num_rows = 0;   // n == 0
    return out;
  }

  hc.lap(0);
  // ---- one launch of the stream kernel (preceded by the kernel that zeroes its workspace) ----
  const int64_t num_tiles = (n + kTileRows - 1) / kTileRows;
  if (num_tiles > INT32_MAX) throw Error(CHDB_ERR_NOT_IMPLEMENTED, "batch too large");
  const int nq = 1 + n_utf8;
  // zeroed region: counts[n_counts] | error word | done | pad || look-back descriptors || bit-packed outputs
  const size_t ws_counts = round_up((size_t)(n_counts + 2) * 8, 128);
  const size_t ws_desc = compact ? round_up((size_t)nq * (size_t)num_tiles * 8, 128) : 0;
  size_t ws_total = ws_counts + ws_desc;
  for (auto& z : zero_reqs) { z.at = ws_total; ws_total += round_up(z.bytes, 128); }
  auto res = std::make_shared<RunResult>();
  res->core = core;
  res->n_counts = n_counts;
  res->workspace = dev_alloc(core, ws_total);
  if ((size_t)(n_counts + 1) * 8 > CtxCore::kPinnedBlock) throw Error(CHDB_ERR_NOT_IMPLEMENTED, "too many counted outputs");
  res->host = (uint64_t*)core->pinned_get();
  res->done = core->event_get();
  uint8_t* ws = (uint8_t*)res->workspace->ptr;
  for (auto& z : zero_reqs) {
    DeviceColumn& dc = out->cols[z.col];
    OutDesc& od = kp.out[z.ko];
    if (z.validity) {
      dc.validity_buf = res->workspace;
      dc.validity = ws + z.at;
      od.validity = ws + z.at;
    } else {
      dc.values_buf = res->workspace;
      dc.values = ws + z.at;
      od.values = ws + z.at;
    }
  }
  CUDA_CHECK(launch_zero(ws, ws_total, core->stream));
  core->launches++;

  hc.lap(1);
  kp.b.num_rows = n;
  kp.b.counts = (uint64_t*)ws;
  kp.b.error_word = (uint64_t*)ws + n_counts;
  kp.b.done = (uint32_t*)((uint64_t*)ws + n_counts + 1);
  kp.b.desc = (uint64_t*)(ws + ws_counts);
  kp.b.host_counts = res->host;
  kp.b.num_tiles = (int32_t)num_tiles;
  kp.b.first_tile = 0;
  kp.n_counts = n_counts;
  kp.n_out = ko;
  kp.n_utf8 = n_utf8;
  kp.pred_begin = compact ? p.pred_begin : 0;
  kp.pred_end = compact ? p.pred_end : 0;
  std::memcpy(kp.instrs, p.instrs.data(), p.instrs.size() * sizeof(Instr));
  std::memcpy(kp.strpool, p.strpool.data(), p.strpool.size());
  // bit-packed outputs assembled in shared memory: Boolean values and validity bitmaps
  kp.n_bits = 0;
  for (int k = 0; k < ko; k++) kp.n_bits += (kp.out[k].type == T_BOOL ? 1 : 0) + (kp.out[k].validity != nullptr ? 1 : 0);
  kp.long_strings = avg_utf8 > 16 ? 1 : 0;
  int64_t slot_avg[kMaxInCols];
  for (size_t s = 0; s < p.slot_to_col.size(); s++) {
    const DeviceColumn& c = in->cols[p.slot_to_col[s]];
    slot_avg[s] = c.meta.type == T_UTF8 && c.value_bytes >= 0 ? (c.value_bytes + n - 1) / n : -1;
    // TMA bulk copies move 16-byte units
    if ((((uintptr_t)c.values | (uintptr_t)c.validity | (uintptr_t)c.offsets) & 15u) != 0)
      throw Error(CHDB_ERR_INVALID_ARGUMENT, "device buffers must be 16-byte aligned");
  }
  // which buffers the kernel touches
  TilePlan tp;
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
  for (int k = 0; k < ko; k++) {
    const OutDesc& od = kp.out[k];
    if (od.kind == OUT_EXPR) mark_instrs(od.begin, od.end);
    else tp.use[od.slot] |= USE_VALUES | USE_VALIDITY | USE_OFFSETS;
  }
  const int ctas_per_sm = plan_tile(kp, tp, slot_avg, false);

  // Long scans run the same device code specialised for this program by NVRTC (jit.cpp); short
  // batches, or boxes without NVRTC, run the bytecode interpreter kernels.
  cudaError_t le = cudaSuccess;
  const JitKernel* jk = nullptr;
  const JitMode jm = jit_mode();
  if (jm == JitMode::Always || (jm == JitMode::Auto && n >= kJitAutoRows)) {
    std::string why;
    jk = jit_get(kp, p.has64, std::min(ctas_per_sm, 8), &why);
  }
  le = jk ? jit_launch_stream(jk, kp, tp, (unsigned)num_tiles, core->stream)
          : launch_stream(kp, tp, p.has64, (unsigned)num_tiles, core->stream);
  core->launches++;
  if (jk) core->jit_launches++;
  if (le != cudaSuccess) throw Error(CHDB_ERR_CUDA, std::string("kernel launch failed: ") + cudaGetErrorString(le));
  hc.lap(2);
  CUDA_CHECK(cudaEventRecord(res->done, core->stream));   // the counts arrive in res->host with the kernel's last CTA
  out->result = res;
  out->num_rows = compact ? -1 : n;
  if (all_const) {
    out->num_rows = 1;   // record_projection.rs:73 over len-1 arrays, whatever the filter kept (the launch only raises the predicate's errors)
  } else if (any_const && compact) {
    // literal columns are len 1: RecordBatch::try_new accepts them only next to exactly one surviving row
    check_run_error(out.get());
    resolve(out.get());
    if (out->num_rows != 1)
      throw Error(CHDB_ERR_INVALID_ARGUMENT, "Invalid argument error: all columns in a record batch must have the same length");
  }
  hc.lap(3);
  return out;
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
  std::string s = jit_prologue(kp, prog->p->has64, 5);
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
        int32_t ends[2] = {0, 0};
        CUDA_CHECK(cudaSetDevice(ctx->core->device));
        CUDA_CHECK(cudaMemcpy(&ends[0], dc.offsets, 4, cudaMemcpyDeviceToHost));
        CUDA_CHECK(cudaMemcpy(&ends[1], dc.offsets + num_rows, 4, cudaMemcpyDeviceToHost));
        dc.first_offset = ends[0];
        dc.value_bytes = ends[1] - ends[0];
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
      if (c.value_bytes < 0 || c.first_offset < 0) throw Error(CHDB_ERR_NOT_IMPLEMENTED, "column buffers of a sliced Utf8 view");
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
      if (c.meta.type == T_UTF8) t += (n + 1) * 4 + std::max<int64_t>(c.value_bytes, 0);
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
void chdb_device_batch_release(chdb_device_batch* b) { delete b; }

int32_t chdb_peer_copy(chdb_ctx* dst_ctx, chdb_ctx* src_ctx, const chdb_device_batch* src, chdb_device_batch** out,
                       chdb_status* st) {
  return guarded(st, [&] {
    if (!dst_ctx || !src_ctx || !src || !out) throw Error(CHDB_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    check_run_error(src);
    resolve(const_cast<chdb_device_batch*>(src));
    const Core& dc = dst_ctx->core;
    const int sdev = src->core->device;
    CUDA_CHECK(cudaSetDevice(dc->device));
    const int64_t n = src->num_rows;
    std::unique_ptr<chdb_device_batch> b(new chdb_device_batch);
    b->core = dc;
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
        if (c.value_bytes < 0) throw Error(CHDB_ERR_NOT_IMPLEMENTED, "peer copy of a sliced Utf8 view");
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
