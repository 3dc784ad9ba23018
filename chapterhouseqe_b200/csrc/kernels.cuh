// Kernel parameter block of the select -> scan -> gather kernels.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#endif

#include "bytecode.h"

namespace chdb {

// A tile is kSlices slices of 128 rows; a slice is the unit one warp works on (4 rows per lane).
// Both streaming kernels are persistent: kProducerWarps TMA producer warps feed a ring of shared-memory stages,
// kComputeGroups groups of kSlices compute warps drain it (group g takes the CTA's tiles g, g + kComputeGroups, ...).
#ifndef CHDB_SLICES
#define CHDB_SLICES 8
#endif
#ifndef CHDB_COMPUTE_GROUPS
#define CHDB_COMPUTE_GROUPS 3
#endif
constexpr int kSlices = CHDB_SLICES;
constexpr int kComputeGroups = CHDB_COMPUTE_GROUPS;
constexpr int kComputeWarps = kComputeGroups * kSlices;
constexpr int kWarpRows = 128;                        // rows of one slice
constexpr int kTileRows = kSlices * kWarpRows;        // 1024 rows per tile
constexpr int kProducerWarps = 3;           // one per kind of buffer: validity bitmaps, Utf8 offsets, values
constexpr int kThreads = (kComputeWarps + kProducerWarps) * 32;
constexpr int kMinCtasPerSm = kThreads <= 512 ? 2 : 1;       // register budget the kernels are compiled for
constexpr int kMaxStages = 8;               // depth of the shared-memory input ring
constexpr int kMaxQuantities = 1 + kMaxOutCols;              // scanned quantities: rows + bytes per Utf8 output
constexpr int kBitWords = kWarpRows / 32 + 2;                // words of one slice's bit-packed output stage
constexpr uint32_t kNotStaged = 0xFFFFFFFFu;
constexpr int kScanThreads = 256;
constexpr int kScanChunk = kScanThreads * 8;  // slices per CTA of the scan kernel

struct ColumnDesc {          // one input column slot (32 bytes)
  const void* values;        // fixed width: values; Boolean: bit-packed values; Utf8: value bytes
  const uint8_t* validity;   // LSB-first bitmap or nullptr (no nulls)
  const int32_t* offsets;    // Utf8 only: int32[num_rows + 1]
  uint8_t type;              // TypeId
  uint8_t width;             // bytes per value (0 for Boolean / Utf8)
  uint8_t pad[6];
};

// Where a column's slice of one tile sits inside a shared-memory stage (byte offsets from the stage
// base, 16-byte aligned), or kNotStaged when the kernel reads that buffer from global memory (or not
// at all).
struct StageSlot {
  uint32_t values;           // fixed width: kTileRows * width bytes; Boolean: kTileRows / 8; Utf8: values_cap bytes
  uint32_t validity;         // kTileRows / 8 bytes
  uint32_t offsets;          // Utf8: (kTileRows + 4) * 4 bytes
  uint32_t values_cap;       // Utf8: capacity for the tile's value bytes (a tile that needs more reads them from global)
};
enum SlotUse : uint8_t { USE_VALUES = 1, USE_VALIDITY = 2, USE_OFFSETS = 4 };

// One streaming kernel's view of the input: which buffers it touches and where they are staged.
struct KernelStage {
  int32_t n_stages;              // depth of the input ring (2 .. kMaxStages)
  int32_t stage_bytes;           // bytes of one stage (multiple of 128)
  uint32_t sel_off;              // gather: the tile's selection bits (kTileRows / 8 bytes), then ...
  uint32_t prefix_off;           // ... its slice prefixes: [quantity][kSlices] uint64
  uint8_t use[kMaxInCols];       // SlotUse mask per input slot
  StageSlot slot[kMaxInCols];
};

enum OutKind : uint8_t { OUT_PASS = 0, OUT_EXPR = 1 };

struct OutDesc {             // one output column that goes through the kernel (32 bytes)
  void* values;
  uint8_t* validity;         // nullptr when the output cannot contain nulls for this batch
  int32_t* offsets;          // Utf8 only
  uint8_t kind;              // OutKind
  uint8_t type;              // TypeId
  uint8_t width;
  uint8_t slot;              // OUT_PASS: input column slot
  uint8_t begin, end;        // OUT_EXPR: instruction range
  uint8_t utf8_index;        // OUT_PASS Utf8: which byte-count scan quantity (0..), else 0xFF
  uint8_t count_index;       // index into counts[] for this column's null count
};

// counts[] layout (uint64 each): [0] output rows, [1 .. 1+n_utf8) output value bytes per Utf8
// output, [1+n_utf8 ..) null count per kernel output, last: error word.
struct KernelParams {
  int64_t num_rows;
  int64_t num_slices;        // ceil(num_rows / 128)
  int64_t slice_pitch;       // num_slices rounded up to a whole number of tiles
  uint32_t* sel_bits;        // select -> gather: one bit per row (kTileRows / 8 bytes per tile, every tile complete)
  uint32_t* slice_counts;    // select -> scan: [quantity][slice_pitch] selected rows / selected value bytes per slice
  uint64_t* slice_prefix;    // scan -> gather: [quantity][slice_pitch] exclusive prefixes
  uint64_t* chunk_desc;      // scan: [quantity][num_chunks] decoupled look-back descriptors (zeroed)
  uint64_t* counts;          // see above (zeroed)
  uint64_t* error_word;      // zeroed; atomicMax(~packed)
  uint64_t* timing;          // debug (CHDB_PHASE_TIMING=1): 16 cycle counters summed over warps, or nullptr
  int32_t num_tiles, num_chunks;
  int32_t n_in, n_out, n_utf8;
  int32_t pred_begin, pred_end;  // pred_begin == pred_end: no predicate (every row is kept; gather only)
  int32_t n_bits;                // bit-packed outputs (Boolean values + validity bitmaps)
  int32_t long_strings;          // 1: per-warp row tables for the chunk-centric long-string copy are allocated
  ColumnDesc in[kMaxInCols];
  OutDesc out[kMaxOutCols];
  Instr instrs[kMaxInstr];
  char strpool[kStrPoolBytes];
};
static_assert(sizeof(KernelParams) + sizeof(KernelStage) <= 4096, "kernel parameters must fit 4 KB");

#ifndef __CUDACC_RTC__
// Fills `st` for one of the streaming kernels (st.use[] set by the caller; needs kp.in[], kp.n_*):
// decides which buffers are staged in shared memory.  avg_utf8[s]: mean value length of Utf8 slot s
// (or < 0).  Returns the dynamic shared memory the launch needs and the CTAs per SM it was sized for.
struct StagePlan { size_t dyn_smem; int ctas_per_sm; };
StagePlan plan_stages(const KernelParams& kp, KernelStage& st, const int64_t* avg_utf8, bool gather);
// has64: the program touches 64-bit types (selects the 64-bit accumulator container).
cudaError_t launch_select(const KernelParams& p, const KernelStage& st, bool has64, const StagePlan& plan, int sm_count, cudaStream_t stream);
cudaError_t launch_scan(const KernelParams& p, cudaStream_t stream);
cudaError_t launch_gather(const KernelParams& p, const KernelStage& st, bool has64, const StagePlan& plan, int sm_count, cudaStream_t stream);
// shared by the ahead-of-time and the run-time compiled kernels
cudaError_t launch_streaming(const void* kernel, const KernelParams& p, const KernelStage& st, const StagePlan& plan, int sm_count,
                             size_t* granted, bool after_kernel, cudaStream_t stream);
#endif

}  // namespace chdb
