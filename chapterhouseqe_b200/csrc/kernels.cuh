// Kernel parameter block of the fused evaluate -> scan -> compact kernel.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#endif

#include "bytecode.h"

namespace chdb {

constexpr int kThreads = 256;               // 8 warps per CTA
constexpr int kWarps = kThreads / 32;
constexpr int kQuadsPerThread = 4;          // each thread owns QPT groups of 4 consecutive rows
constexpr int kTileRows = kThreads * 4 * kQuadsPerThread;   // 4096 rows per tile, 512 per warp

struct ColumnDesc {          // one input column slot (32 bytes)
  const void* values;        // fixed width: values; Boolean: bit-packed values; Utf8: value bytes
  const uint8_t* validity;   // LSB-first bitmap or nullptr (no nulls)
  const int32_t* offsets;    // Utf8 only: int32[num_rows + 1]
  uint8_t type;              // TypeId
  uint8_t width;             // bytes per value (0 for Boolean / Utf8)
  uint8_t pad[6];
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
  uint64_t* timing;          // debug (CHDB_PHASE_TIMING): 8 globaltimer stamps per tile, or nullptr
  int32_t num_tiles;
  int32_t n_in, n_out, n_utf8;
  int32_t pred_begin, pred_end;  // pred_begin == pred_end: no predicate (every row is kept)
  int32_t stage_bytes;           // bytes of ONE warp's output staging slice in dynamic shared memory
  int32_t prefetch_tiles;        // L2 prefetch distance in tiles (about one wave of resident CTAs)
  ColumnDesc in[kMaxInCols];
  OutDesc out[kMaxOutCols];
  Instr instrs[kMaxInstr];
  char strpool[kStrPoolBytes];
};
static_assert(sizeof(KernelParams) <= 4096, "KernelParams must fit the 4 KB kernel parameter space");

#ifndef __CUDACC_RTC__
// has64: the program touches 64-bit types (selects the 64-bit accumulator container).
cudaError_t launch_filter_project(const KernelParams& p, bool has64, size_t dyn_smem, cudaStream_t stream);
// Per-warp output staging slice: 256 rows of the widest fixed-width output, or of short Utf8 values.
size_t filter_project_stage_bytes(int max_out_width, int64_t avg_utf8_len);
size_t filter_project_smem_bytes(size_t stage_bytes, bool has_utf8_out);
#endif

}  // namespace chdb
