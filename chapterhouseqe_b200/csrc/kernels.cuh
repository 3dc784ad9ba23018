// Kernel parameter block of the fused filter / project / compact kernel ("stream" kernel).
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#endif

#include "bytecode.h"

namespace chdb {

// One CTA works on one tile.  A tile is kTileSlices slices of 128 rows; a slice is the unit one warp
// works on at a time (4 consecutive rows per lane, so every column access is a 128-bit load); warp w
// owns slices [w * kSpw, (w + 1) * kSpw) of its tile.  Many small CTAs (instead of one persistent CTA
// per SM) let the hardware hide one tile's load latency behind its neighbours' work.  Measured on C2
// (4 warps x 2 slices = 1024-row tiles: 83.9 us per 4M-row batch; 4 x 1 = 512-row tiles: 77.7 us; 8 x 1:
// 78.9 us): the default is 512-row tiles, 8-10 of them resident per SM.
#ifndef CHDB_WARPS
#define CHDB_WARPS 4
#endif
#ifndef CHDB_SPW
#define CHDB_SPW 1
#endif
constexpr int kWarps = CHDB_WARPS;
constexpr int kSpw = CHDB_SPW;                        // slices per warp
constexpr int kTileSlices = kWarps * kSpw;
constexpr int kWarpRows = 128;                        // rows of one slice
constexpr int kTileRows = kTileSlices * kWarpRows;    // 512 rows per tile
constexpr int kThreads = kWarps * 32;
constexpr int kMaxQuantities = 1 + kMaxOutCols;       // scanned quantities: rows + bytes per Utf8 output
constexpr int kBitWords = kWarpRows / 32 + 2;         // words of one slice's bit-packed output stage
constexpr uint32_t kNotStaged = 0xFFFFFFFFu;
constexpr int kZeroThreads = 256;
constexpr int kGroupTiles = 64;                       // two-launch form: tile totals are also summed per group of tiles
static_assert(kTileSlices <= 32, "a warp scans the tile's slice counts in one go");

struct ColumnDesc {          // one input column slot (32 bytes)
  const void* values;        // fixed width: values; Boolean: bit-packed values; Utf8: value bytes
  const uint8_t* validity;   // LSB-first bitmap or nullptr (no nulls)
  const int32_t* offsets;    // Utf8 only: int32[num_rows + 1]
  uint8_t type;              // TypeId
  uint8_t width;             // bytes per value (0 for Boolean / Utf8)
  uint8_t pad[6];
};

// Where a column's slice of the tile sits inside the CTA's shared-memory stage (byte offsets from the
// stage base, 16-byte aligned), or kNotStaged when the kernel reads that buffer from global memory
// (or not at all).
struct StageSlot {
  uint32_t values;           // fixed width: kTileRows * width bytes; Boolean: kTileRows / 8; Utf8: values_cap bytes
  uint32_t validity;         // kTileRows / 8 bytes
  uint32_t offsets;          // Utf8: (kTileRows + 4) * 4 bytes
  uint32_t values_cap;       // Utf8: capacity for the tile's value bytes (a tile that needs more reads them from global)
};
enum SlotUse : uint8_t { USE_VALUES = 1, USE_VALIDITY = 2, USE_OFFSETS = 4 };

// The CTA's dynamic shared memory: the stage (the tile's slice of every staged buffer, brought in by TMA
// bulk copies) followed by the small per-tile tables.
struct TilePlan {
  uint32_t cols_off;             // ColumnDesc[n_in]: the input columns as seen by this tile (pointers biased so that
                                 // indexing with the ABSOLUTE row lands in the stage, or in global memory)
  uint32_t cnt_off;              // uint32[quantity][kTileSlices]: per-slice counts
  uint32_t pre_off;              // uint64[quantity][kTileSlices]: per-slice exclusive prefixes (batch-wide)
  uint32_t tot_off;              // uint64[quantity]: batch totals (valid in the last tile)
  uint32_t bits_off;             // uint32[warp][n_bits][kBitWords]: per-warp bit stages
  uint32_t ltab_off;             // long strings: per-warp row tables
  uint32_t pext_off;             // uint8[256] bit-compaction table (only with bit-packed outputs)
  uint32_t params_off;           // MANY: this tile's KernelParams copy
  uint32_t dyn_smem;             // total
  uint32_t pred_reads_utf8;      // the predicate compares strings: it waits for the Utf8 value bytes too
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

// What differs from batch to batch (pointers and sizes).  A launch over many small batches reads one of
// these per batch from global memory (packed: header, then n_in ColumnDesc, then n_out OutDesc).
struct BatchHeader {
  int64_t num_rows;
  uint64_t* desc;            // [quantity][num_tiles] decoupled look-back descriptors (zeroed); two-launch form: plain tile
                             // totals, followed by [quantity][ceil(num_tiles / kGroupTiles)] group totals
  uint64_t* counts;          // see below (zeroed)
  uint64_t* error_word;      // zeroed; atomicMax(~packed)
  uint64_t* host_counts;     // pinned host mirror of counts[] + error word, written by the last CTA to finish
  uint32_t* done;            // zeroed; CTAs that have finished
  uint32_t* selbits;         // select -> gather: one selection bit per row, LSB-first, whole tiles (two-launch form only)
  int32_t num_tiles;
  int32_t first_tile;        // MANY: blockIdx.x of this batch's tile 0
};

// counts[] layout (uint64 each): [0] output rows, [1 .. 1+n_utf8) output value bytes per Utf8
// output, [1+n_utf8 ..) null count per kernel output, last: error word.
struct KernelParams {
  BatchHeader b;
  int32_t n_in, n_out, n_utf8;
  int32_t pred_begin, pred_end;  // pred_begin == pred_end: no predicate (every row is kept)
  int32_t n_bits;                // bit-packed outputs (Boolean values + validity bitmaps)
  int32_t long_strings;          // 1: per-warp row tables for the chunk-centric long-string copy are allocated
  int32_t n_counts;              // entries of counts[] before the error word
  int32_t early_counts;          // two-launch form, pass-through outputs only: the select kernel counts NULLs and publishes the totals
  // MANY (one launch over several batches of one schema and shape): packed per-batch records in global memory
  uint64_t* trace;               // debugging aid (CHDB_TRACE): [tile][8] clock64() stamps of the CTA's phases, or nullptr
  const uint8_t* many;           // nullptr: single batch (everything is in this block)
  const int32_t* many_tile_batch;  // [grid] batch index of every tile
  int32_t many_batches, many_stride;
  ColumnDesc in[kMaxInCols];
  OutDesc out[kMaxOutCols];
  Instr instrs[kMaxInstr];
  char strpool[kStrPoolBytes];
};
static_assert(sizeof(KernelParams) + sizeof(TilePlan) <= 4096, "kernel parameters must fit 4 KB");

#ifndef __CUDACC_RTC__
// Fills `tp` (tp.use[] set by the caller; needs kp.in[], kp.n_*): decides which buffers are staged in
// shared memory.  avg_utf8[s]: mean value length of Utf8 slot s (or < 0).  Returns the CTAs per SM the
// plan leaves room for.
// stage = false: the plan of the select kernel (nothing staged, only the small tables).
int plan_tile(const KernelParams& kp, TilePlan& tp, const int64_t* avg_utf8, bool many, bool stage = true);
// has64: the program touches 64-bit types (selects the 64-bit accumulator container).  mode: StreamMode (0 fused, 1 select, 2 gather).
cudaError_t launch_stream(const KernelParams& p, const TilePlan& tp, bool has64, int mode, unsigned grid, cudaStream_t stream);
// Zeroes `bytes` (a multiple of 16) at p; the stream kernel that follows is launched as its programmatic dependent.
cudaError_t launch_zero(void* p, size_t bytes, cudaStream_t stream);
// shared by the ahead-of-time and the run-time compiled kernels
cudaError_t launch_stream_kernel(const void* kernel, const KernelParams& p, const TilePlan& tp, unsigned grid, cudaStream_t stream);
#endif

}  // namespace chdb
