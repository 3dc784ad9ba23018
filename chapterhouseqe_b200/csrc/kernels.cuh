// Kernel parameter block of the fused evaluate -> scan -> compact kernel.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#endif

#include "bytecode.h"

namespace chdb {

// A tile is kConsumerWarps slices of 128 rows.  CTA = kWriterGroups groups of kConsumerWarps writer warps
// (phase B; group g takes the CTA's tiles g, g + kWriterGroups, ...) + kConsumerWarps selector warps
// (phase A) + one TMA producer warp + kLookbackWarps look-back warps (warp j takes tiles j, j + kLookbackWarps, ...).
#ifndef CHDB_CONSUMER_WARPS
#define CHDB_CONSUMER_WARPS 8
#endif
#ifndef CHDB_WRITER_GROUPS
#define CHDB_WRITER_GROUPS 2
#endif
constexpr int kConsumerWarps = CHDB_CONSUMER_WARPS;
constexpr int kWriterGroups = CHDB_WRITER_GROUPS;
constexpr int kWriterWarps = kWriterGroups * kConsumerWarps;
constexpr int kQuadsPerThread = 1;          // each thread owns 4 consecutive rows per tile
constexpr int kWarpRows = 32 * 4 * kQuadsPerThread;          // rows of one warp slice
constexpr int kTileRows = kConsumerWarps * kWarpRows;        // 1024 rows per tile
constexpr int kLookbackWarps = 3;            // look-backs in flight per CTA (one L2 round trip each, ~2 us under load)
constexpr int kThreads = (kWriterWarps + kConsumerWarps + 1 + kLookbackWarps) * 32;
constexpr int kMinCtasPerSm = kThreads <= 512 ? 2 : 1;       // register budget the kernels are compiled for
constexpr int kMaxStages = 6;               // depth of the shared-memory input ring
constexpr int kMaxQuantities = 1 + kMaxOutCols;              // scanned quantities: rows + bytes per Utf8 output
constexpr int kBitWords = kTileRows / 32 + 2;                // words of one bit-packed output stage
constexpr uint32_t kNotStaged = 0xFFFFFFFFu;

struct ColumnDesc {          // one input column slot (32 bytes)
  const void* values;        // fixed width: values; Boolean: bit-packed values; Utf8: value bytes
  const uint8_t* validity;   // LSB-first bitmap or nullptr (no nulls)
  const int32_t* offsets;    // Utf8 only: int32[num_rows + 1]
  uint8_t type;              // TypeId
  uint8_t width;             // bytes per value (0 for Boolean / Utf8)
  uint8_t pad[6];
};

// Where a column's slice of one tile sits inside a shared-memory stage (byte offsets from the stage
// base, 16-byte aligned), or kNotStaged when the kernel reads that buffer from global memory.
struct StageSlot {
  uint32_t values;           // fixed width: kTileRows * width bytes; Boolean: kTileRows / 8; Utf8: values_cap bytes
  uint32_t validity;         // kTileRows / 8 bytes
  uint32_t offsets;          // Utf8: (kTileRows + 4) * 4 bytes
  uint32_t values_cap;       // Utf8: capacity for the tile's value bytes (a tile that needs more reads them from global)
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
  uint64_t* tile_desc;       // [(1 + n_utf8)][num_tiles] decoupled look-back descriptors (zeroed)
  uint32_t* ticket;          // dynamic tile id counter (zeroed)
  uint64_t* counts;          // see above (zeroed)
  uint64_t* error_word;      // zeroed; atomicMax(~packed)
  uint64_t* timing;          // debug (CHDB_PHASE_TIMING=1): 16 cycle counters summed over warps, or nullptr
  int32_t num_tiles;
  int32_t n_in, n_out, n_utf8;
  int32_t pred_begin, pred_end;  // pred_begin == pred_end: no predicate (every row is kept)
  int32_t n_stages;              // depth of the input ring (2 .. kMaxStages)
  int32_t stage_bytes;           // bytes of one stage (multiple of 128)
  int32_t n_bits;                // bit-packed outputs (Boolean values + validity bitmaps)
  int32_t long_strings;          // 1: per-warp row tables for the chunk-centric long-string copy are allocated
  ColumnDesc in[kMaxInCols];
  StageSlot stage[kMaxInCols];
  OutDesc out[kMaxOutCols];
  Instr instrs[kMaxInstr];
  char strpool[kStrPoolBytes];
};
static_assert(sizeof(KernelParams) <= 4096, "KernelParams must fit the 4 KB kernel parameter space");

#ifndef __CUDACC_RTC__
// Fills kp.stage[], kp.n_stages, kp.stage_bytes (needs kp.in[], kp.out[], kp.n_*): decides which
// buffers are staged in shared memory.  avg_utf8[s]: mean value length of Utf8 slot s (or < 0).
// Returns the dynamic shared memory the launch needs and the CTAs per SM it was sized for.
struct StagePlan { size_t dyn_smem; int ctas_per_sm; };
StagePlan plan_stages(KernelParams& kp, const int64_t* avg_utf8);
size_t filter_project_static_smem();   // the kernel's static shared memory (barriers, per-stage tile control blocks)
// has64: the program touches 64-bit types (selects the 64-bit accumulator container).
cudaError_t launch_filter_project(const KernelParams& p, bool has64, const StagePlan& plan, int sm_count, cudaStream_t stream);
#endif

}  // namespace chdb
