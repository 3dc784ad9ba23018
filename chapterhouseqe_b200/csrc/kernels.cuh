// Kernel parameter block of the fused filter / project / compact kernel ("stream" kernel).
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#endif

#include "bytecode.h"

namespace chdb {

// The stream kernel is persistent: a few CTAs per SM, each a software pipeline over tiles it draws from a
// ticket counter.  A tile is kTileSlices slices of 128 rows; a slice is the unit one compute warp works on
// (4 consecutive rows per lane, so every column access is a 128-bit load); compute warp w owns slice w of
// every tile of its CTA.  Besides the compute warps a CTA has one producer warp (tickets, TMA bulk loads
// into a ring of shared-memory stages) and kScanWarps scan warps (decoupled look-back).
#ifndef CHDB_COMPUTE_WARPS
#define CHDB_COMPUTE_WARPS 8
#endif
constexpr int kComputeWarps = CHDB_COMPUTE_WARPS;
constexpr int kTileSlices = kComputeWarps;
constexpr int kWarpRows = 128;                        // rows of one slice
constexpr int kTileRows = kTileSlices * kWarpRows;    // 1024 rows per tile
#ifndef CHDB_SCAN_WARPS
#define CHDB_SCAN_WARPS 2
#endif
constexpr int kScanWarps = CHDB_SCAN_WARPS;           // look-backs of consecutive tiles run concurrently, one per scan warp
constexpr int kProducerWarp = kComputeWarps, kScanWarp = kComputeWarps + 1;   // scan warps: kScanWarp .. kScanWarp + kScanWarps - 1
constexpr int kThreads = (kComputeWarps + 1 + kScanWarps) * 32;
constexpr int kMaxStages = 8;
constexpr int kTraceIters = 32;
constexpr int kMaxQuantities = 1 + kMaxOutCols;       // scanned quantities: rows + bytes per Utf8 output
constexpr int kBitWords = kWarpRows / 32 + 2;         // words of one slice's bit-packed output stage
constexpr uint32_t kNotStaged = 0xFFFFFFFFu;
constexpr int kZeroThreads = 256;
static_assert(kTileSlices <= 32, "a warp scans the tile's slice counts in one go");

struct ColumnDesc {          // one input column slot (32 bytes)
  const void* values;        // fixed width: values; Boolean: bit-packed values; Utf8: value bytes
  const uint8_t* validity;   // LSB-first bitmap or nullptr (no nulls)
  const int32_t* offsets;    // Utf8 only: int32[num_rows + 1]
  uint8_t type;              // TypeId
  uint8_t width;             // bytes per value (0 for Boolean / Utf8)
  uint8_t pad[6];
};

// Where a column's slice of the tile sits inside a ring stage (byte offsets from the stage base, 16-byte
// aligned), or kNotStaged when the kernel reads that buffer from global memory (or not at all).
struct StageSlot {
  uint32_t values;           // fixed width: kTileRows * width bytes; Boolean: kTileRows / 8; Utf8: values_cap bytes
  uint32_t validity;         // kTileRows / 8 bytes
  uint32_t offsets;          // Utf8: (kTileRows + 4) * 4 bytes
  uint32_t values_cap;       // Utf8: capacity for the tile's value bytes (a tile that needs more reads them from global)
};
enum SlotUse : uint8_t { USE_VALUES = 1, USE_VALIDITY = 2, USE_OFFSETS = 4 };

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

// What differs from batch to batch (pointers and sizes).  A launch over many small batches reads one of
// these per batch from global memory (packed: header, then n_in ColumnDesc, then n_out OutDesc).
struct BatchHeader {
  int64_t num_rows;
  uint64_t* desc;            // [descriptor group][num_tiles] decoupled look-back descriptors (zeroed)
  uint64_t* counts;          // see below (zeroed)
  uint64_t* error_word;      // zeroed; atomicMax(~packed)
  uint64_t* host_counts;     // pinned host mirror of counts[] + error word, written by the last CTA to finish
  int32_t num_tiles;
  int32_t first_tile;        // MANY: ticket of this batch's tile 0
  int64_t pad[2];
};
static_assert(sizeof(BatchHeader) == 64, "BatchHeader is copied in 16-byte units");

// What a ring stage holds besides the staged buffers: which tile it is, and the tile's view of the batch.
struct StageCtx {
  int64_t tile;              // tile index inside its batch; -1: the ticket counter ran out (no more work)
  int64_t row0;
  int32_t rows;
  int32_t batch;             // MANY: batch index
  int64_t pad;
  // followed by ColumnDesc cols[n_in] (pointers biased so that indexing with the ABSOLUTE row lands in the
  // stage, or in global memory), and with MANY by the batch's BatchHeader and OutDesc out[n_out]
};
static_assert(sizeof(StageCtx) == 32, "StageCtx layout");

// The CTA's dynamic shared memory: the ring of stages followed by the small tables.
struct TilePlan {
  uint32_t stages;               // ring depth
  uint32_t stage_bytes;          // staged buffers of one stage (multiple of 128)
  uint32_t sctx_off, sctx_stride;  // StageCtx [+ cols + header + outs] per stage
  uint32_t cnt_off;              // uint32[stage][quantity][kTileSlices]: per-slice counts
  uint32_t pre_off;              // uint64[stage][quantity][kTileSlices]: per-slice exclusive prefixes (batch-wide)
  uint32_t sel_off;              // uint8[stage][compute warp][lane]: the lanes' selection bits between A and B
  uint32_t nulls_off;            // uint32[stages][kMaxOutCols]: NULLs written per output (per stage with MANY)
  uint32_t bits_off;             // uint32[compute warp][n_bits][kBitWords]: per-warp bit stages
  uint32_t ltab_off;             // long strings: per-warp row tables
  uint32_t pext_off;             // uint8[256] bit-compaction table (only with bit-packed outputs)
  uint32_t dyn_smem;             // total
  uint32_t ctas_per_sm;          // resident CTAs per SM the plan leaves room for
  uint8_t use[kMaxInCols];       // SlotUse mask per input slot
  StageSlot slot[kMaxInCols];
};

// counts[] layout (uint64 each): [0] output rows, [1 .. 1+n_utf8) output value bytes per Utf8
// output, [1+n_utf8 ..) null count per kernel output, last: error word.
struct KernelParams {
  BatchHeader b;
  uint32_t* tickets;             // zeroed: next tile to hand out
  uint32_t* done;                // zeroed: CTAs that have finished
  int32_t total_tiles;           // over all batches of the launch
  int32_t n_in, n_out, n_utf8;
  int32_t pred_begin, pred_end;  // pred_begin == pred_end: no predicate (every row is kept)
  int32_t n_bits;                // bit-packed outputs (Boolean values + validity bitmaps)
  int32_t long_strings;          // 1: per-warp row tables for the chunk-centric long-string copy are allocated
  int32_t n_counts;              // entries of counts[] before the error word
  // MANY (one launch over several batches of one schema and shape): packed per-batch records in global memory
  int32_t many_batches, many_stride;
  uint64_t* trace;               // debugging aid (CHDB_TRACE): [cta][kTraceIters][8] clock64() stamps of the pipeline, or nullptr
  const uint8_t* many;           // nullptr: single batch (everything is in this block)
  const int32_t* many_tile_batch;  // [total_tiles] batch index of every ticket
  ColumnDesc in[kMaxInCols];
  OutDesc out[kMaxOutCols];
  Instr instrs[kMaxInstr];
  char strpool[kStrPoolBytes];
};
static_assert(sizeof(KernelParams) + sizeof(TilePlan) <= 4096, "kernel parameters must fit 4 KB");

// Look-back descriptors pack two scanned quantities into one 64-bit word: {flag:2 | odd quantity:31 | even
// quantity:31}; Arrow's int32 offsets bound both (rows and Utf8 bytes of one batch stay below 2^31).
CHDB_HD constexpr int desc_groups(int quantities) { return (quantities + 1) / 2; }

#ifndef __CUDACC_RTC__
// Fills `tp` (tp.use[] set by the caller; needs kp.in[], kp.n_*): decides which buffers are staged in
// shared memory and how deep the ring is.  avg_utf8[s]: mean value length of Utf8 slot s (or < 0).
void plan_tile(const KernelParams& kp, TilePlan& tp, const int64_t* avg_utf8, bool many);
// has64: the program touches 64-bit types (selects the 64-bit accumulator container).
cudaError_t launch_stream(const KernelParams& p, const TilePlan& tp, bool has64, int sm_count, cudaStream_t stream);
// Zeroes `bytes` (a multiple of 16) at p; the stream kernel that follows is launched as its programmatic dependent.
cudaError_t launch_zero(void* p, size_t bytes, cudaStream_t stream);
// shared by the ahead-of-time and the run-time compiled kernels
cudaError_t launch_stream_kernel(const void* kernel, const KernelParams& p, const TilePlan& tp, int sm_count, cudaStream_t stream);
#endif

}  // namespace chdb
