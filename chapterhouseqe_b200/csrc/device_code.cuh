// Fused WHERE-evaluation -> selection -> decoupled-look-back scan -> compaction kernel (sm_100a).
//
// One CTA owns one tile of kTileRows rows.  Per tile:
//   0. the tile's slice of every input column is prefetched into L2 (one 128-byte line per thread
//      and iteration), so the dependent loads below find their data on chip
//   1. every thread runs the predicate bytecode over its rows (128-bit coalesced column loads,
//      accumulator in registers) and gets a selection mask                     [compute_value.rs]
//   2. warp shuffles rank the selected rows inside each warp's 256-row slice; for each Utf8 output
//      the selected value bytes are summed the same way; warp totals meet in shared memory
//   3. a decoupled look-back over 64-bit {flag | value} tile descriptors turns the tile totals
//      into exclusive prefixes (rows, and bytes per Utf8 output)               [filter_record.rs:37]
//   4. from here on every WARP works alone (no block barriers): per output column it stages its
//      selected values in its private slice of shared memory at the destination's 16-byte phase
//      and writes them with aligned 16-byte stores; validity and Boolean bits are staged one byte
//      per row and packed 32 at a time; short Utf8 values are staged the same way, long ones are
//      produced output-chunk-centric.
// HBM traffic is therefore each referenced input byte once and each output byte once.
// Projection expressions are evaluated in step 4 under the selection mask, so checked-integer
// errors are raised for surviving rows only (the reference projects after filtering).
//
// Accumulator convention: for 8/16/32-bit integers and Float32 only the low 32 bits of the
// container are meaningful (integers sign-/zero-extended to 32 bits); 64-bit types use all of it.
//
// Build with -fmad=false: float results must be the IEEE single operations arrow-rs performs.
#pragma once
#include "kernels.cuh"

namespace chdb {
namespace {


constexpr uint32_t FULL = 0xFFFFFFFFu;
constexpr int kWarpBitStage = kTileRows / kWarps + 64;   // per warp: one byte per output row + word-alignment slack

// ------------------------------------------------------------------------------------------
// program access: at run time from the kernel parameters (generic interpreter), or at compile
// time when jit.cpp hands this file to NVRTC behind a generated prologue that defines
//   namespace chdb_jit { kInstrs[], kPredBegin, kPredEnd, kNumIn, kColType[], kColWidth[],
//                        kNumOut, kOutMeta[], kNumUtf8 }
// -- then every dispatch below folds away and the loads of a whole expression are scheduled
// together (the bytecode is "interpreted" by the compiler).
// ------------------------------------------------------------------------------------------
#ifdef CHDB_JIT
#define CHDB_STATIC_UNROLL _Pragma("unroll")
#define CHDB_N_IN chdb_jit::kNumIn
#define CHDB_N_OUT chdb_jit::kNumOut
#define CHDB_N_UTF8 chdb_jit::kNumUtf8
#define CHDB_PRED_BEGIN chdb_jit::kPredBegin
#define CHDB_PRED_END chdb_jit::kPredEnd
#define CHDB_COL_TYPE(P, s) chdb_jit::kColType[s]
#define CHDB_COL_WIDTH(P, s) chdb_jit::kColWidth[s]
#define CHDB_OUT_META(P, k) chdb_jit::kOutMeta[k]
#else
#define CHDB_STATIC_UNROLL _Pragma("unroll 1")
#define CHDB_N_IN P.n_in
#define CHDB_N_OUT P.n_out
#define CHDB_N_UTF8 P.n_utf8
#define CHDB_PRED_BEGIN P.pred_begin
#define CHDB_PRED_END P.pred_end
#define CHDB_COL_TYPE(P, s) P.in[s].type
#define CHDB_COL_WIDTH(P, s) P.in[s].width
// the eight small fields of OutDesc arrive as one 64-bit constant load
#define CHDB_OUT_META(P, k) (reinterpret_cast<const uint64_t*>(&P.out[k])[3])
#endif

template <typename V> struct Cont;
template <> struct Cont<uint32_t> { static constexpr bool k64 = false; };
template <> struct Cont<uint64_t> { static constexpr bool k64 = true; };

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
__device__ __noinline__ void report_error(const KernelParams& P, uint32_t order, int64_t row, uint32_t code) {
  unsigned long long packed = ((unsigned long long)order << 56) | (((unsigned long long)row & 0xFFFFFFFFFFFFull) << 8) | code;
  atomicMax((unsigned long long*)P.error_word, ~packed);
}

// bad / divz: per-thread row masks of failing rows (already restricted to evaluated rows)
template <int QPT>
__device__ __forceinline__ void report_rows(const KernelParams& P, const Instr& in, uint32_t ovf, uint32_t divz,
                                            const int64_t (&qbase)[QPT]) {
  const uint32_t any = ovf | divz;
  if (any) {
    const int j = __ffs(any) - 1;   // rows ascend with j inside a thread
    int64_t row = qbase[0];
#pragma unroll
    for (int q = 1; q < QPT; q++)
      if ((j >> 2) == q) row = qbase[q];   // static indexing keeps qbase in registers
    report_error(P, in.order, row + (j & 3), ((divz >> j) & 1u) ? kErrDivideByZero : kErrArithmeticOverflow);
  }
}

// ------------------------------------------------------------------------------------------
// loads
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// phase stamps of one tile (debugging aid, enabled by CHDB_PHASE_TIMING=1 in the environment)
#define CHDB_STAMP(i) do { if (P.timing != nullptr && (threadIdx.x & 31) == 0) { if ((i) < 4) { if (threadIdx.x == 0) P.timing[(size_t)tile * 8 + (i)] = global_ns(); } else atomicMax((unsigned long long*)&P.timing[(size_t)tile * 8 + (i)], (unsigned long long)global_ns()); } } while (0)

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int QPT>
__device__ __forceinline__ uint32_t load_bits(const uint8_t* __restrict__ bits, const int64_t (&qbase)[QPT], uint32_t need) {
  if (bits == nullptr) return FULL;
  uint32_t m = 0;
#pragma unroll
  for (int q = 0; q < QPT; q++) {
    if ((need >> (4 * q)) & 0xFu) {
      const uint32_t byte = __ldg(bits + (qbase[q] >> 3));
      m |= ((byte >> (uint32_t)(qbase[q] & 4)) & 0xFu) << (4 * q);
    }
  }
  return m;
}

// Column values of the thread's rows in accumulator form (see the convention above).
template <typename V, int QPT>
__device__ __forceinline__ void fetch_col(const ColumnDesc& c, uint8_t from_type, const int64_t (&qbase)[QPT], uint32_t need,
                                          V (&b)[4 * QPT]) {
  const uint8_t* __restrict__ base = (const uint8_t*)c.values;
#define CHDB_SKIP_QUAD(q) if (!((need >> (4 * (q))) & 0xFu)) { b[4 * (q)] = 0; b[4 * (q) + 1] = 0; b[4 * (q) + 2] = 0; b[4 * (q) + 3] = 0; continue; }
  switch (from_type) {
    case T_I32: case T_U32: case T_F32:
#pragma unroll
      for (int q = 0; q < QPT; q++) {
        CHDB_SKIP_QUAD(q)
        const uint4 x = __ldg((const uint4*)(base + qbase[q] * 4));
        b[4 * q + 0] = x.x; b[4 * q + 1] = x.y; b[4 * q + 2] = x.z; b[4 * q + 3] = x.w;
      }
      break;
    case T_I64: case T_U64: case T_F64:
      if constexpr (Cont<V>::k64) {
#pragma unroll
        for (int q = 0; q < QPT; q++) {
          CHDB_SKIP_QUAD(q)
          const uint4 x = __ldg((const uint4*)(base + qbase[q] * 8));
          const uint4 y = __ldg((const uint4*)(base + qbase[q] * 8 + 16));
          b[4 * q + 0] = x.x | ((uint64_t)x.y << 32); b[4 * q + 1] = x.z | ((uint64_t)x.w << 32);
          b[4 * q + 2] = y.x | ((uint64_t)y.y << 32); b[4 * q + 3] = y.z | ((uint64_t)y.w << 32);
        }
      }
      break;
    case T_I16: case T_U16:
#pragma unroll
      for (int q = 0; q < QPT; q++) {
        CHDB_SKIP_QUAD(q)
        const uint2 x = __ldg((const uint2*)(base + qbase[q] * 2));
        if (from_type == T_I16) {
          b[4 * q + 0] = (uint32_t)(int32_t)(int16_t)(x.x & 0xFFFFu); b[4 * q + 1] = (uint32_t)(int32_t)(int16_t)(x.x >> 16);
          b[4 * q + 2] = (uint32_t)(int32_t)(int16_t)(x.y & 0xFFFFu); b[4 * q + 3] = (uint32_t)(int32_t)(int16_t)(x.y >> 16);
        } else {
          b[4 * q + 0] = x.x & 0xFFFFu; b[4 * q + 1] = x.x >> 16; b[4 * q + 2] = x.y & 0xFFFFu; b[4 * q + 3] = x.y >> 16;
        }
      }
      break;
    default:  // T_I8 / T_U8
#pragma unroll
      for (int q = 0; q < QPT; q++) {
        CHDB_SKIP_QUAD(q)
        const uint32_t x = __ldg((const uint32_t*)(base + qbase[q]));
        if (from_type == T_I8) {
          b[4 * q + 0] = (uint32_t)(int32_t)(int8_t)(x & 0xFFu); b[4 * q + 1] = (uint32_t)(int32_t)(int8_t)((x >> 8) & 0xFFu);
          b[4 * q + 2] = (uint32_t)(int32_t)(int8_t)((x >> 16) & 0xFFu); b[4 * q + 3] = (uint32_t)(int32_t)(int8_t)(x >> 24);
        } else {
          b[4 * q + 0] = x & 0xFFu; b[4 * q + 1] = (x >> 8) & 0xFFu; b[4 * q + 2] = (x >> 16) & 0xFFu; b[4 * q + 3] = x >> 24;
        }
      }
      break;
  }
}

#undef CHDB_SKIP_QUAD

// ------------------------------------------------------------------------------------------
// casts (arrow-cast on the coercion lattice; int -> float is round-to-nearest-even)
// ------------------------------------------------------------------------------------------
template <typename V, int R>
__device__ __forceinline__ void cast_vals(V (&a)[R], uint8_t from, uint8_t to) {
  const TypeClass fc = type_class(from), tc = type_class(to);
  if (fc == tc) return;
  if (tc == C_F32) {
    if (fc == C_SINT) {
#pragma unroll
      for (int j = 0; j < R; j++) a[j] = (V)__float_as_uint((float)(int32_t)(uint32_t)a[j]);
    } else if (fc == C_UINT) {
#pragma unroll
      for (int j = 0; j < R; j++) a[j] = (V)__float_as_uint((float)(uint32_t)a[j]);
    } else if (fc == C_S64) {
#pragma unroll
      for (int j = 0; j < R; j++) a[j] = (V)__float_as_uint((float)(int64_t)a[j]);
    } else if (fc == C_U64) {
#pragma unroll
      for (int j = 0; j < R; j++) a[j] = (V)__float_as_uint((float)(uint64_t)a[j]);
    }
    return;
  }
  if constexpr (Cont<V>::k64) {
    if (tc == C_F64) {
      if (fc == C_SINT) {
#pragma unroll
        for (int j = 0; j < R; j++) a[j] = (V)__double_as_longlong((double)(int32_t)(uint32_t)a[j]);
      } else if (fc == C_UINT) {
#pragma unroll
        for (int j = 0; j < R; j++) a[j] = (V)__double_as_longlong((double)(uint32_t)a[j]);
      } else if (fc == C_S64) {
#pragma unroll
        for (int j = 0; j < R; j++) a[j] = (V)__double_as_longlong((double)(int64_t)a[j]);
      } else if (fc == C_U64) {
#pragma unroll
        for (int j = 0; j < R; j++) a[j] = (V)__double_as_longlong((double)(uint64_t)a[j]);
      } else if (fc == C_F32) {
#pragma unroll
        for (int j = 0; j < R; j++) {
          const uint32_t u = (uint32_t)a[j];
          const float x = __uint_as_float(u);
          uint64_t r;
          if (x != x)  // keep sign and payload, quiet (x86 cvtss2sd)
            r = ((uint64_t)(u & 0x80000000u) << 32) | 0x7FF8000000000000ull | ((uint64_t)(u & 0x007FFFFFu) << 29);
          else
            r = (uint64_t)__double_as_longlong((double)x);
          a[j] = (V)r;
        }
      }
    } else if (tc == C_S64 || tc == C_U64) {   // widening from a 32-bit container
      if (fc == C_SINT) {
#pragma unroll
        for (int j = 0; j < R; j++) a[j] = (V)(int64_t)(int32_t)(uint32_t)a[j];
      } else if (fc == C_UINT) {
#pragma unroll
        for (int j = 0; j < R; j++) a[j] = (V)(uint32_t)a[j];
      }
    }
  }
}

template <typename V, int R>
__device__ __forceinline__ uint32_t tobool_vals(const V (&a)[R], uint8_t t) {
  const TypeClass c = type_class(t);
  uint32_t m = 0;
  if (c == C_F32) {
#pragma unroll
    for (int j = 0; j < R; j++) m |= (__uint_as_float((uint32_t)a[j]) != 0.0f ? 1u : 0u) << j;   // NaN -> true, -0.0 -> false
  } else if (c == C_F64) {
#pragma unroll
    for (int j = 0; j < R; j++) m |= (__longlong_as_double((long long)(uint64_t)a[j]) != 0.0 ? 1u : 0u) << j;
  } else if (c == C_S64 || c == C_U64) {
#pragma unroll
    for (int j = 0; j < R; j++) m |= (a[j] != 0 ? 1u : 0u) << j;
  } else {
#pragma unroll
    for (int j = 0; j < R; j++) m |= ((uint32_t)a[j] != 0 ? 1u : 0u) << j;
  }
  return m;
}

// ------------------------------------------------------------------------------------------
// arithmetic (arrow-arith numeric.rs: checked integers on valid slots, IEEE floats everywhere)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float nanfix32(float r, float x, float y) {
  if (r != r) {
    uint32_t bits;
    if (x != x) bits = __float_as_uint(x) | 0x00400000u;
    else if (y != y) bits = __float_as_uint(y) | 0x00400000u;
    else bits = 0xFFC00000u;  // x86 default NaN has the sign bit set
    r = __uint_as_float(bits);
  }
  return r;
}
__device__ __forceinline__ double nanfix64(double r, double x, double y) {
  if (r != r) {
    unsigned long long bits;
    if (x != x) bits = (unsigned long long)__double_as_longlong(x) | 0x0008000000000000ull;
    else if (y != y) bits = (unsigned long long)__double_as_longlong(y) | 0x0008000000000000ull;
    else bits = 0xFFF8000000000000ull;
    r = __longlong_as_double((long long)bits);
  }
  return r;
}

// Rare, slow scalar paths stay out of line so the unrolled row loops around them remain small
// (and the accumulator arrays are only ever indexed statically, i.e. stay in registers).
// flags: 1 = overflow, 2 = divide by zero.
struct Slow64 { uint64_t r; uint32_t flags; };
struct Slow32 { int32_t r; uint32_t flags; };
__device__ __noinline__ Slow64 slow_i64(uint32_t op, int64_t x, int64_t y) {
  uint32_t fl = 0, *flags = &fl;
  int64_t r = 0;
  if (op == OP_DIV || op == OP_REM) {
    if (y == 0) *flags |= 2u;
    else if (x == INT64_MIN && y == -1) *flags |= 1u;
    else r = op == OP_DIV ? x / y : x % y;
  } else if (op == OP_ADD) {
    r = (int64_t)((uint64_t)x + (uint64_t)y);
    if (((x ^ r) & (y ^ r)) < 0) *flags |= 1u;
  } else if (op == OP_MUL) {
    r = (int64_t)((uint64_t)x * (uint64_t)y);
    if (__mul64hi(x, y) != (r >> 63)) *flags |= 1u;
  } else {
    r = (int64_t)((uint64_t)x - (uint64_t)y);
    if (((x ^ y) & (x ^ r)) < 0) *flags |= 1u;
  }
  return Slow64{(uint64_t)r, fl};
}
__device__ __noinline__ Slow64 slow_u64(uint32_t op, uint64_t x, uint64_t y) {
  uint32_t fl = 0, *flags = &fl;
  uint64_t r = 0;
  if (op == OP_DIV || op == OP_REM) {
    if (y == 0) *flags |= 2u;
    else r = op == OP_DIV ? x / y : x % y;
  } else if (op == OP_ADD) {
    r = x + y;
    if (r < x) *flags |= 1u;
  } else if (op == OP_MUL) {
    r = x * y;
    if (__umul64hi(x, y) != 0) *flags |= 1u;
  } else {
    r = x - y;
    if (x < y) *flags |= 1u;
  }
  return Slow64{r, fl};
}
// 8/16-bit integers: exact in 64 bits, then range-checked against [lo, hi]
__device__ __noinline__ Slow32 slow_narrow(uint32_t op, int32_t x32, int32_t y32, int32_t lo, int32_t hi) {
  uint32_t fl = 0, *flags = &fl;
  const int64_t x = x32, y = y32;
  int64_t r = 0;
  if (op == OP_DIV || op == OP_REM) {
    if (y == 0) *flags |= 2u;
    else if (lo < 0 && x == lo && y == -1) *flags |= 1u;
    else r = op == OP_DIV ? x / y : x % y;
  } else {
    r = op == OP_ADD ? x + y : op == OP_MUL ? x * y : x - y;
    if (r < lo || r > hi) { *flags |= 1u; r = 0; }
  }
  return Slow32{(int32_t)r, fl};
}

// IMM: the operand is the instruction's immediate (uniform); SWAP: operand is the LEFT side.
// Rows that are null or filtered out may hold anything afterwards: arrow leaves them unobservable.
template <bool IMM, bool SWAP, typename V, int QPT>
__device__ __forceinline__ void arith(const KernelParams& P, const Instr& in, V (&a)[4 * QPT], uint32_t& av, const V (&b)[4 * QPT],
                                      uint32_t bv, uint32_t active, const int64_t (&qbase)[QPT]) {
  constexpr int R = 4 * QPT;
  const uint8_t op = in.op, t = in.type;
  const uint32_t valid = av & bv;
  av = valid;
  const uint32_t m = valid & active;   // fallible ops are only *checked* on valid, live rows
  const V immv = (V)in.imm;
#define CHDB_B(j) (IMM ? immv : b[j])
#define CHDB_X(j) (SWAP ? CHDB_B(j) : a[j])
#define CHDB_Y(j) (SWAP ? a[j] : CHDB_B(j))
  uint32_t ovf = 0, divz = 0;
  if (t == T_I32) {
    if (op == OP_ADD) {
#pragma unroll
      for (int j = 0; j < R; j++) {
        const int32_t x = (int32_t)(uint32_t)CHDB_X(j), y = (int32_t)(uint32_t)CHDB_Y(j);
        const int32_t r = (int32_t)((uint32_t)x + (uint32_t)y);
        ovf |= ((uint32_t)((x ^ r) & (y ^ r)) >> 31) << j;
        a[j] = (V)(uint32_t)r;
      }
    } else if (op == OP_MUL) {
#pragma unroll
      for (int j = 0; j < R; j++) {
        const int32_t x = (int32_t)(uint32_t)CHDB_X(j), y = (int32_t)(uint32_t)CHDB_Y(j);
        const int64_t p = (int64_t)x * (int64_t)y;
        const int32_t r = (int32_t)p;
        ovf |= (p != (int64_t)r ? 1u : 0u) << j;
        a[j] = (V)(uint32_t)r;
      }
    } else if (op == OP_SUB) {
#pragma unroll
      for (int j = 0; j < R; j++) {
        const int32_t x = (int32_t)(uint32_t)CHDB_X(j), y = (int32_t)(uint32_t)CHDB_Y(j);
        const int32_t r = (int32_t)((uint32_t)x - (uint32_t)y);
        ovf |= ((uint32_t)((x ^ y) & (x ^ r)) >> 31) << j;
        a[j] = (V)(uint32_t)r;
      }
    } else {
      const int32_t d = (int32_t)(uint32_t)in.imm;
      if (IMM && !SWAP && d > 0 && (d & (d - 1)) == 0) {   // divisor 2^k: no error is possible
        const int k = __ffs(d) - 1;
#pragma unroll
        for (int j = 0; j < R; j++) {
          const int32_t x = (int32_t)(uint32_t)a[j];
          const int32_t q = (x + ((x >> 31) & (d - 1))) >> k;   // truncating division
          a[j] = (V)(uint32_t)(op == OP_DIV ? q : x - (q << k));
        }
      } else {
#pragma unroll
        for (int j = 0; j < R; j++) {
          const int32_t x = (int32_t)(uint32_t)CHDB_X(j), y = (int32_t)(uint32_t)CHDB_Y(j);
          const bool z = y == 0, o = x == INT32_MIN && y == -1;
          divz |= (z ? 1u : 0u) << j;
          ovf |= (o ? 1u : 0u) << j;
          const int32_t ys = (z || o) ? 1 : y;
          a[j] = (V)(uint32_t)(op == OP_DIV ? x / ys : x % ys);
        }
      }
    }
  } else if (t == T_F32) {
#pragma unroll
    for (int j = 0; j < R; j++) {
      const float x = __uint_as_float((uint32_t)CHDB_X(j)), y = __uint_as_float((uint32_t)CHDB_Y(j));
      float r;
      switch (op) {
        case OP_ADD: r = __fadd_rn(x, y); break;
        case OP_MUL: r = __fmul_rn(x, y); break;
        case OP_DIV: r = __fdiv_rn(x, y); break;
        case OP_REM: r = fmodf(x, y); break;
        default: r = __fsub_rn(x, y); break;
      }
      a[j] = (V)__float_as_uint(nanfix32(r, x, y));
    }
  } else if (t == T_U32) {
    const uint32_t d = (uint32_t)in.imm;
    if ((op == OP_DIV || op == OP_REM) && IMM && !SWAP && d != 0 && (d & (d - 1)) == 0) {
      const int k = __ffs((int)d) - 1;
#pragma unroll
      for (int j = 0; j < R; j++) {
        const uint32_t x = (uint32_t)a[j];
        a[j] = (V)(op == OP_DIV ? x >> k : x & (d - 1));
      }
    } else {
#pragma unroll
      for (int j = 0; j < R; j++) {
        const uint32_t x = (uint32_t)CHDB_X(j), y = (uint32_t)CHDB_Y(j);
        uint32_t r;
        if (op == OP_ADD) { r = x + y; ovf |= (r < x ? 1u : 0u) << j; }
        else if (op == OP_MUL) { const uint64_t p = (uint64_t)x * y; r = (uint32_t)p; ovf |= ((p >> 32) != 0 ? 1u : 0u) << j; }
        else if (op == OP_SUB) { r = x - y; ovf |= (x < y ? 1u : 0u) << j; }
        else { const bool z = y == 0; divz |= (z ? 1u : 0u) << j; const uint32_t ys = z ? 1u : y; r = op == OP_DIV ? x / ys : x % ys; }
        a[j] = (V)r;
      }
    }
  } else if (t == T_I8 || t == T_I16 || t == T_U8 || t == T_U16) {
    const int32_t lo = t == T_I8 ? -128 : t == T_I16 ? -32768 : 0;
    const int32_t hi = t == T_I8 ? 127 : t == T_I16 ? 32767 : t == T_U8 ? 255 : 65535;
#pragma unroll
    for (int j = 0; j < R; j++) {
      const Slow32 sr = slow_narrow(op, (int32_t)(uint32_t)CHDB_X(j), (int32_t)(uint32_t)CHDB_Y(j), lo, hi);
      a[j] = (V)(uint32_t)sr.r;
      ovf |= (sr.flags & 1u) << j;
      divz |= (sr.flags >> 1) << j;
    }
  } else {
    if constexpr (Cont<V>::k64) {
      if (t == T_F64) {
#pragma unroll
        for (int j = 0; j < R; j++) {
          const double x = __longlong_as_double((long long)CHDB_X(j)), y = __longlong_as_double((long long)CHDB_Y(j));
          double r;
          switch (op) {
            case OP_ADD: r = __dadd_rn(x, y); break;
            case OP_MUL: r = __dmul_rn(x, y); break;
            case OP_DIV: r = __ddiv_rn(x, y); break;
            case OP_REM: r = fmod(x, y); break;
            default: r = __dsub_rn(x, y); break;
          }
          a[j] = (V)__double_as_longlong(nanfix64(r, x, y));
        }
      } else if (t == T_I64) {
#pragma unroll
        for (int j = 0; j < R; j++) {
          const Slow64 sr = slow_i64(op, (int64_t)CHDB_X(j), (int64_t)CHDB_Y(j));
          a[j] = (V)sr.r;
          ovf |= (sr.flags & 1u) << j;
          divz |= (sr.flags >> 1) << j;
        }
      } else {  // T_U64
#pragma unroll
        for (int j = 0; j < R; j++) {
          const Slow64 sr = slow_u64(op, (uint64_t)CHDB_X(j), (uint64_t)CHDB_Y(j));
          a[j] = (V)sr.r;
          ovf |= (sr.flags & 1u) << j;
          divz |= (sr.flags >> 1) << j;
        }
      }
    }
  }
#undef CHDB_B
#undef CHDB_X
#undef CHDB_Y
  report_rows<QPT>(P, in, ovf & m, divz & m, qbase);
}

// ------------------------------------------------------------------------------------------
// comparisons (arrow-ord cmp.rs: natural integer order, IEEE-754 totalOrder for floats)
//   eq(a,b) | lt(a,b) | gt(a,b) = lt(b,a);  ne / ge / le are their complements
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int32_t total_key32(uint32_t u) {
  int32_t k = (int32_t)u;
  return k ^ (int32_t)(((uint32_t)(k >> 31)) >> 1);
}
__device__ __forceinline__ int64_t total_key64(uint64_t u) {
  int64_t k = (int64_t)u;
  return k ^ (int64_t)(((uint64_t)(k >> 63)) >> 1);
}

#define CHDB_CMP_LOOP(XT, XEXPR, YEXPR)                                                   \
  if (mode == 0) {                                                                        \
    _Pragma("unroll") for (int j = 0; j < R; j++) { const XT x = XEXPR, y = YEXPR; r |= (x == y ? 1u : 0u) << j; } \
  } else if (mode == 1) {                                                                 \
    _Pragma("unroll") for (int j = 0; j < R; j++) { const XT x = XEXPR, y = YEXPR; r |= (x < y ? 1u : 0u) << j; }  \
  } else {                                                                                \
    _Pragma("unroll") for (int j = 0; j < R; j++) { const XT x = XEXPR, y = YEXPR; r |= (y < x ? 1u : 0u) << j; }  \
  }

template <bool IMM, typename V, int R>
__device__ __forceinline__ uint32_t compare(const Instr& in, const V (&a)[R], uint32_t am, const V (&b)[R], uint32_t bm) {
  const uint8_t kind = in.aux;
  const TypeClass tc = type_class(in.type);
  const int mode = (kind == CMP_EQ || kind == CMP_NE) ? 0 : (kind == CMP_LT || kind == CMP_GE) ? 1 : 2;
  const bool negate = kind == CMP_NE || kind == CMP_GE || kind == CMP_LE;
  const V immv = (V)in.imm;
#define CHDB_B(j) (IMM ? immv : b[j])
  uint32_t r = 0;
  switch (tc) {
    case C_BOOL:  // false < true
      r = mode == 0 ? ~(am ^ bm) : mode == 1 ? (~am & bm) : (am & ~bm);
      break;
    case C_SINT: CHDB_CMP_LOOP(int32_t, (int32_t)(uint32_t)a[j], (int32_t)(uint32_t)CHDB_B(j)) break;
    case C_UINT: CHDB_CMP_LOOP(uint32_t, (uint32_t)a[j], (uint32_t)CHDB_B(j)) break;
    case C_F32:
      if (mode == 0) {  // bitwise: NaN == NaN with equal payloads, -0.0 != +0.0
        CHDB_CMP_LOOP(uint32_t, (uint32_t)a[j], (uint32_t)CHDB_B(j))
      } else {
        CHDB_CMP_LOOP(int32_t, total_key32((uint32_t)a[j]), total_key32((uint32_t)CHDB_B(j)))
      }
      break;
    default:
      if constexpr (Cont<V>::k64) {
        if (tc == C_S64) { CHDB_CMP_LOOP(int64_t, (int64_t)a[j], (int64_t)CHDB_B(j)) }
        else if (tc == C_U64) { CHDB_CMP_LOOP(uint64_t, (uint64_t)a[j], (uint64_t)CHDB_B(j)) }
        else if (mode == 0) { CHDB_CMP_LOOP(uint64_t, (uint64_t)a[j], (uint64_t)CHDB_B(j)) }
        else { CHDB_CMP_LOOP(int64_t, total_key64((uint64_t)a[j]), total_key64((uint64_t)CHDB_B(j))) }
      }
      break;
  }
#undef CHDB_B
  return negate ? ~r : r;
}

// Utf8: bytewise lexicographic; operands are columns or a literal from the string pool.
template <int QPT> struct QuadBases { int64_t v[QPT]; };

template <int QPT>
__device__ __noinline__ uint32_t cmp_utf8(const KernelParams& P, const Instr in, const QuadBases<QPT> qb, uint32_t inrange,
                                          const uint8_t* s_pool, uint32_t* valid_out) {
  const int64_t (&qbase)[QPT] = qb.v;
  uint32_t valid;
  const uint32_t slot_a = in.slot, slot_b = (uint32_t)(in.imm >> 56);
  const uint32_t pool_off = (uint32_t)in.imm, pool_len = (uint32_t)(in.imm >> 32) & 0xFFFFFFu;
  valid = FULL;
  if (slot_a != 0xFFu) valid &= load_bits<QPT>(P.in[slot_a].validity, qbase, inrange);
  if (slot_b != 0xFFu) valid &= load_bits<QPT>(P.in[slot_b].validity, qbase, inrange);
  uint32_t lt = 0, eq = 0;
#pragma unroll 1
  for (int j = 0; j < 4 * QPT; j++) {
    if (!((inrange >> j) & 1u)) continue;
    const int64_t row = qbase[j >> 2] + (j & 3);
    const uint8_t *pa, *pb;
    int la, lb;
    if (slot_a != 0xFFu) {
      const int32_t* off = P.in[slot_a].offsets;
      const int o0 = __ldg(off + row), o1 = __ldg(off + row + 1);
      pa = (const uint8_t*)P.in[slot_a].values + o0;
      la = o1 - o0;
    } else {
      pa = s_pool + pool_off;
      la = (int)pool_len;
    }
    if (slot_b != 0xFFu) {
      const int32_t* off = P.in[slot_b].offsets;
      const int o0 = __ldg(off + row), o1 = __ldg(off + row + 1);
      pb = (const uint8_t*)P.in[slot_b].values + o0;
      lb = o1 - o0;
    } else {
      pb = s_pool + pool_off;
      lb = (int)pool_len;
    }
    const int n = la < lb ? la : lb;
    int c = 0;
    for (int k = 0; k < n; k++) {
      const int x = pa[k], y = pb[k];
      if (x != y) { c = x < y ? -1 : 1; break; }
    }
    if (c == 0) c = la < lb ? -1 : (la > lb ? 1 : 0);
    lt |= (c < 0 ? 1u : 0u) << j;
    eq |= (c == 0 ? 1u : 0u) << j;
  }
  *valid_out = valid;
  switch (in.aux) {
    case CMP_EQ: return eq;
    case CMP_NE: return ~eq;
    case CMP_LT: return lt;
    case CMP_LE: return lt | eq;
    case CMP_GT: return ~(lt | eq);
    default: return ~lt;
  }
}


// ------------------------------------------------------------------------------------------
// the interpreter: accumulator in registers, one operand per instruction.
// Every handler updates the accumulator in place and keeps its operand array local to its own
// scope, so nothing but (acc, accm, accv) is carried around the dispatch loop.  It runs on QI
// quads (4 * QI rows per thread) at a time.
// ------------------------------------------------------------------------------------------
template <typename V, int QI>
struct Spill {
  V v[kMaxSpill][4 * QI];
  uint32_t m[kMaxSpill], valid[kMaxSpill];
};

// Fetches the operand of `in` (column or spill slot; immediates are handled by the IMM templates).
template <typename V, int QI>
__device__ __forceinline__ void fetch_operand(const KernelParams& P, const Instr& in, const int64_t (&qbase)[QI], uint32_t inrange,
                                              const Spill<V, QI>& stk, V (&b)[4 * QI], uint32_t& bm, uint32_t& bv) {
  constexpr int R = 4 * QI;
  if (in.src == SRC_COL) {
    const ColumnDesc& c = P.in[in.slot];
    bv = load_bits<QI>(c.validity, qbase, inrange);
    if (CHDB_COL_TYPE(P, in.slot) == T_BOOL) {
      bm = load_bits<QI>((const uint8_t*)c.values, qbase, inrange);
#pragma unroll
      for (int j = 0; j < R; j++) b[j] = 0;
    } else {
      fetch_col<V, QI>(c, in.from_type, qbase, inrange, b);
      if (in.type == T_BOOL) bm = tobool_vals<V, R>(b, in.from_type);
      else { bm = 0; cast_vals<V, R>(b, in.from_type, in.type); }
    }
  } else {  // SRC_STK
    if (in.type != T_BOOL) {   // Boolean spills only carry the two masks
#pragma unroll
      for (int j = 0; j < R; j++) b[j] = stk.v[in.slot][j];
    } else {
#pragma unroll
      for (int j = 0; j < R; j++) b[j] = 0;
    }
    bm = stk.m[in.slot];
    bv = stk.valid[in.slot];
  }
}

// Executes one instruction on the accumulator.  `in` is a run-time value in the generic kernel and
// a compile-time constant under CHDB_JIT (everything below then folds to the one handler).
template <typename V, int QI>
__device__ __forceinline__ void exec_instr(const KernelParams& P, const Instr in, const int64_t (&qbase)[QI], uint32_t inrange,
                                           uint32_t active, const uint8_t* s_pool, Spill<V, QI>& stk, V (&acc)[4 * QI],
                                           uint32_t& accm, uint32_t& accv) {
  constexpr int R = 4 * QI;
  const bool imm = in.src == SRC_IMM;
  switch (in.op) {
    case OP_LOAD:
      if (imm) {
#pragma unroll
        for (int j = 0; j < R; j++) acc[j] = (V)in.imm;
        accm = in.imm ? FULL : 0u;
        accv = FULL;
      } else {
        fetch_operand<V, QI>(P, in, qbase, inrange, stk, acc, accm, accv);   // straight into the accumulator
      }
      break;
    case OP_CAST: cast_vals<V, R>(acc, in.from_type, in.type); break;
    case OP_ADD: case OP_MUL: case OP_DIV: case OP_REM: case OP_SUB:
      if (imm) {
        if (in.flags & OPF_SWAP) arith<true, true, V, QI>(P, in, acc, accv, acc, FULL, active, qbase);
        else arith<true, false, V, QI>(P, in, acc, accv, acc, FULL, active, qbase);
      } else {
        V b[R];
        uint32_t bm, bv;
        fetch_operand<V, QI>(P, in, qbase, inrange, stk, b, bm, bv);
        if (in.flags & OPF_SWAP) arith<false, true, V, QI>(P, in, acc, accv, b, bv, active, qbase);
        else arith<false, false, V, QI>(P, in, acc, accv, b, bv, active, qbase);
      }
      break;
    case OP_CMP:
      if (imm) {
        accm = compare<true, V, R>(in, acc, accm, acc, in.imm ? FULL : 0u);
      } else {
        V b[R];
        uint32_t bm, bv;
        fetch_operand<V, QI>(P, in, qbase, inrange, stk, b, bm, bv);
        accm = compare<false, V, R>(in, acc, accm, b, bm);
        accv &= bv;
      }
      break;
    case OP_TOBOOL: accm = tobool_vals<V, R>(acc, in.type); break;
    case OP_AND: case OP_OR: {   // non-Kleene: null if either side is null
      uint32_t bm = in.imm ? FULL : 0u, bv = FULL;
      if (in.src == SRC_STK) {
        bm = stk.m[in.slot];
        bv = stk.valid[in.slot];
      } else if (in.src == SRC_COL) {
        V b[R];
        fetch_operand<V, QI>(P, in, qbase, inrange, stk, b, bm, bv);
      }
      accm = in.op == OP_AND ? (accm & bm) : (accm | bm);
      accv &= bv;
      break;
    }
    case OP_PUSH:
      if (in.type != T_BOOL) {
#pragma unroll
        for (int j = 0; j < R; j++) stk.v[in.slot][j] = acc[j];
      }
      stk.m[in.slot] = accm;
      stk.valid[in.slot] = accv;
      break;
    case OP_CMP_UTF8: {
      QuadBases<QI> qb;
#pragma unroll
      for (int q = 0; q < QI; q++) qb.v[q] = qbase[q];
      uint32_t v = FULL;
      accm = cmp_utf8<QI>(P, in, qb, inrange, s_pool, &v);
      accv = v;
      break;
    }
    default: break;
  }
}

#ifdef CHDB_JIT
template <typename V, int QI, int PC, int END>
__device__ __forceinline__ void run_range(const KernelParams& P, const int64_t (&qbase)[QI], uint32_t inrange, uint32_t active,
                                          const uint8_t* s_pool, Spill<V, QI>& stk, V (&acc)[4 * QI], uint32_t& accm,
                                          uint32_t& accv) {
  if constexpr (PC < END) {
    constexpr Instr in = chdb_jit::kInstrs[PC];
    exec_instr<V, QI>(P, in, qbase, inrange, active, s_pool, stk, acc, accm, accv);
    run_range<V, QI, PC + 1, END>(P, qbase, inrange, active, s_pool, stk, acc, accm, accv);
  }
}
#endif

// BEGIN/END >= 0: instruction range known at compile time (CHDB_JIT); otherwise [begin, end).
template <typename V, int QI, int BEGIN = -1, int END = -1>
__device__ __forceinline__ void run_program(const KernelParams& P, int begin, int end, const int64_t (&qbase)[QI], uint32_t inrange,
                                            uint32_t active, const uint8_t* s_pool, V (&acc)[4 * QI], uint32_t& accm,
                                            uint32_t& accv) {
  constexpr int R = 4 * QI;
  Spill<V, QI> stk;
#pragma unroll
  for (int j = 0; j < R; j++) acc[j] = 0;
  accm = 0;
  accv = FULL;
#ifdef CHDB_JIT
  if constexpr (BEGIN >= 0) {
    run_range<V, QI, BEGIN, END>(P, qbase, inrange, active, s_pool, stk, acc, accm, accv);
    return;
  }
#endif
#pragma unroll 1
  for (int pc = begin; pc < end; pc++) {
    // one 16-byte instruction = two 64-bit constant-bank loads, fields peeled off with shifts
    const uint2 w = *reinterpret_cast<const uint2*>(&P.instrs[pc]);
    Instr in;
    in.op = (uint8_t)w.x; in.type = (uint8_t)(w.x >> 8); in.src = (uint8_t)(w.x >> 16); in.flags = (uint8_t)(w.x >> 24);
    in.slot = (uint8_t)w.y; in.from_type = (uint8_t)(w.y >> 8); in.aux = (uint8_t)(w.y >> 16); in.order = (uint8_t)(w.y >> 24);
    in.imm = P.instrs[pc].imm;
    exec_instr<V, QI>(P, in, qbase, inrange, active, s_pool, stk, acc, accm, accv);
  }
}

// ------------------------------------------------------------------------------------------
// scans
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, int lane, uint32_t& total) {
  uint32_t x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t y = __shfl_up_sync(FULL, x, d);
    if (lane >= d) x += y;
  }
  total = __shfl_sync(FULL, x, 31);
  return x - v;
}

__device__ __forceinline__ uint64_t warp_sum64(uint64_t v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
  return v;
}

// Decoupled look-back (Merrill & Garland) on packed {flag:2 | value:62} descriptors; executed by
// one full warp.  Tiles are numbered by an atomic ticket, so every predecessor is already
// resident or finished and the spin always terminates.
constexpr uint64_t kFlagAgg = 1ull << 62, kFlagPrefix = 2ull << 62, kValueMask = (1ull << 62) - 1;
__device__ __forceinline__ uint64_t lookback(uint64_t* desc, uint32_t tile, uint64_t agg, int lane) {
  volatile uint64_t* d = desc;
  if (tile == 0) {
    if (lane == 0) d[0] = kFlagPrefix | agg;
    return 0;
  }
  if (lane == 0) d[tile] = kFlagAgg | agg;
  uint64_t excl = 0;
  int64_t base = (int64_t)tile - 1;
  while (true) {
    // each lane inspects 4 consecutive predecessors (nearest first): 128 tiles per hop
    uint64_t part = 0;
    bool found = false;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int64_t idx = base - (lane * 4 + j);
      uint64_t v = kFlagPrefix;  // tiles "before 0" contribute an inclusive prefix of 0
      if (idx >= 0 && !found) {
        do { v = d[idx]; } while ((v >> 62) == 0);
      }
      if (!found) {
        part += v & kValueMask;
        found = (v >> 62) == 2;
      }
    }
    const uint32_t pm = __ballot_sync(FULL, found);
    if (pm) {
      const int first = __ffs(pm) - 1;  // lane holding the nearest predecessor that already knows its prefix
      excl += warp_sum64(lane <= first ? part : 0);
      break;
    }
    excl += warp_sum64(part);
    base -= 128;
  }
  if (lane == 0) d[tile] = kFlagPrefix | (excl + agg);
  return excl;
}

// ------------------------------------------------------------------------------------------
// output staging -- per WARP.  After the tile prefixes are known every warp gathers its own 256
// rows on its own: it stages the selected values of one column in its private slice of shared
// memory at the destination's 16-byte phase and writes them with aligned 16-byte stores, with
// only __syncwarp() in between, so no warp ever waits for another one in this phase.
// Shared memory is addressed with explicit 32-bit shared-window addresses (computed once per
// kernel) so the hot loops are plain LDS / STS with register bases.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { uint64_t a;
  asm("cvta.to.shared.u64 %0, %1;" : "=l"(a) : "l"(p));
  return (uint32_t)a; }
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)v) : "memory"); }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t a, uint32_t lo, uint32_t hi) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
// Asynchronous global -> shared copies (LDGSTS): the data of the next output column streams into the
// warp's second buffer while the current column is being compacted.
__device__ __forceinline__ void cp_async16(uint32_t dst_s, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_s), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// nbytes is rounded up to 16 (buffers are padded); src and dst_s must be 16-byte aligned
__device__ __forceinline__ void async_copy_slice(uint32_t dst_s, const uint8_t* src, uint32_t nbytes, int lane) {
  for (uint32_t b = (uint32_t)lane * 16u; b < nbytes; b += 512u) cp_async16(dst_s + b, src + b);
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t a) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
  return v;
}

template <int W>
__device__ __forceinline__ void copy_elem_s2g(uint8_t* g, uint32_t s) {   // one W-byte element, shared -> global
  if (W == 4) { *(uint32_t*)g = lds32(s); }
  else if (W == 8) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(s) : "memory"); *(uint2*)g = v; }
  else if (W == 2) { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(s) : "memory"); *(uint16_t*)g = v; }
  else { *g = (uint8_t)lds8(s); }
}

// stage[mis, mis + nbytes) -> gdst_aligned[mis, ...), by one warp: aligned 16-byte stores in the
// middle; the (at most two) 16-byte chunks shared with the neighbours are written element-wise,
// one element per lane.
template <int W>
__device__ __forceinline__ void warp_writeout(uint32_t stage_s, uint8_t* gdst_aligned, uint32_t mis, uint32_t nbytes, int lane) {
  const uint32_t end = mis + nbytes;
  const uint32_t first_full = (mis + 15u) >> 4, end_full = end >> 4;   // full chunks: [first_full, end_full)
  uint32_t c = first_full + lane;
  for (; c + 32 < end_full; c += 64) {   // two independent 16-byte copies per trip
    const uint4 a = lds128(stage_s + (c << 4)), b = lds128(stage_s + ((c + 32) << 4));
    *(uint4*)(gdst_aligned + (c << 4)) = a;
    *(uint4*)(gdst_aligned + ((c + 32) << 4)) = b;
  }
  if (c < end_full) *(uint4*)(gdst_aligned + (c << 4)) = lds128(stage_s + (c << 4));
  constexpr uint32_t EPC = 16 / W;   // elements per chunk
  if (mis != 0) {                    // head chunk 0: lanes 0 .. EPC-1 take one element each
    const uint32_t b = (uint32_t)lane * W;
    if ((uint32_t)lane < EPC && b >= mis && b < end) copy_elem_s2g<W>(gdst_aligned + b, stage_s + b);
  }
  if ((end & 15u) != 0 && (end_full > 0 || mis == 0)) {   // tail chunk (unless it is also the head): lanes 16 .. 16+EPC-1
    const uint32_t b = (end_full << 4) + ((uint32_t)lane - 16u) * W;
    if ((uint32_t)lane >= 16u && (uint32_t)lane < 16u + EPC && b < end) copy_elem_s2g<W>(gdst_aligned + b, stage_s + b);
  }
}

// What a warp knows about its slice of the output once the prefixes are in.
template <int QPT>
struct WarpOut {
  uint64_t base;          // global output row index of the warp's first selected row
  uint32_t count;         // selected rows of this warp
  uint32_t rank[QPT];     // warp-local rank of the first selected row of each of this lane's quads
  uint32_t sel;           // selection bits of this lane's rows (4 per quad)
};

// Stages the selected rows of quad q of a W-byte pass-through column (the load was issued by the caller).
template <int W>
__device__ __forceinline__ void stage_quad(uint32_t a, uint32_t s4, const uint4& x, const uint4& y) {
  if (W == 4) {
    if (s4 & 1u) { sts32(a, x.x); a += 4; }
    if (s4 & 2u) { sts32(a, x.y); a += 4; }
    if (s4 & 4u) { sts32(a, x.z); a += 4; }
    if (s4 & 8u) { sts32(a, x.w); }
  } else if (W == 8) {
    if (s4 & 1u) { sts64(a, x.x, x.y); a += 8; }
    if (s4 & 2u) { sts64(a, x.z, x.w); a += 8; }
    if (s4 & 4u) { sts64(a, y.x, y.y); a += 8; }
    if (s4 & 8u) { sts64(a, y.z, y.w); }
  } else if (W == 2) {
    if (s4 & 1u) { sts16(a, x.x & 0xFFFFu); a += 2; }
    if (s4 & 2u) { sts16(a, x.x >> 16); a += 2; }
    if (s4 & 4u) { sts16(a, x.y & 0xFFFFu); a += 2; }
    if (s4 & 8u) { sts16(a, x.y >> 16); }
  } else {
    if (s4 & 1u) { sts8(a, x.x & 0xFFu); a += 1; }
    if (s4 & 2u) { sts8(a, (x.x >> 8) & 0xFFu); a += 1; }
    if (s4 & 4u) { sts8(a, (x.x >> 16) & 0xFFu); a += 1; }
    if (s4 & 8u) { sts8(a, x.x >> 24); }
  }
}

// Gathers the selected rows of a W-byte pass-through column: loads two quads at a time (in flight
// together), stages them, one write-out for the whole warp slice.
template <int W, int QPT>
__device__ __forceinline__ void gather_fixed(const uint8_t* __restrict__ src, uint8_t* dst, int64_t row_base, const WarpOut<QPT>& wo,
                                             uint32_t stage_s, int lane) {
  const uint32_t mis = (uint32_t)((wo.base * W) & 15u);
#pragma unroll
  for (int g = 0; g < QPT; g += 2) {
    uint4 x[2], y[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int q = g + h;
      x[h] = make_uint4(0, 0, 0, 0);
      y[h] = make_uint4(0, 0, 0, 0);
      if (q >= QPT || !((wo.sel >> (4 * q)) & 0xFu)) continue;
      const int64_t r = row_base + q * 128;
      if (W == 4) {
        x[h] = __ldg((const uint4*)(src + r * 4));
      } else if (W == 8) {
        x[h] = __ldg((const uint4*)(src + r * 8));
        y[h] = __ldg((const uint4*)(src + r * 8 + 16));
      } else if (W == 2) {
        const uint2 t = __ldg((const uint2*)(src + r * 2));
        x[h].x = t.x; x[h].y = t.y;
      } else {
        x[h].x = __ldg((const uint32_t*)(src + r));
      }
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int q = g + h;
      if (q >= QPT) continue;
      const uint32_t s4 = (wo.sel >> (4 * q)) & 0xFu;
      if (s4) stage_quad<W>(stage_s + mis + wo.rank[q] * W, s4, x[h], y[h]);
    }
  }
  __syncwarp();
  warp_writeout<W>(stage_s, dst + ((wo.base * W) & ~15ull), mis, wo.count * W, lane);
  __syncwarp();
}

// Stages four values that already sit in registers (one quad of a projection expression / of the
// rebuilt Utf8 offsets); the caller writes the slice out once all quads are staged.
template <int W, typename E>
__device__ __forceinline__ void stage_regs(uint32_t a, uint32_t s4, const E (&e)[4]) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    if ((s4 >> i) & 1u) {
      if (W == 4) sts32(a, (uint32_t)e[i]);
      else if (W == 8) sts64(a, (uint32_t)e[i], (uint32_t)((uint64_t)e[i] >> 32));
      else if (W == 2) sts16(a, (uint32_t)e[i] & 0xFFFFu);
      else sts8(a, (uint32_t)e[i] & 0xFFu);
      a += W;
    }
  }
}

// Compacts one bit per row (validity or Boolean values) into gbits at bit offset wo.base:
// every selected row drops its bit as one byte at its rank, then one lane per output word packs
// 32 bytes with eight multiplies.  gbits is zero-initialised; the (at most two) words shared with
// neighbouring warps are merged with atomicOr.
template <int QPT>
__device__ __forceinline__ void compact_bits(uint32_t bits, const WarpOut<QPT>& wo, uint32_t bstage_s, uint32_t* gbits, int lane) {
  const uint32_t o = (uint32_t)(wo.base & 31);
#pragma unroll
  for (int q = 0; q < QPT; q++) {
    uint32_t a = bstage_s + o + wo.rank[q];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int j = 4 * q + i;
      if ((wo.sel >> j) & 1u) { sts8(a, (bits >> j) & 1u); a++; }
    }
  }
  __syncwarp();
  const uint32_t end = o + wo.count;
  const uint32_t nwords = (end + 31u) >> 5;     // <= 9 for a 256-row warp slice
  const uint64_t g0 = wo.base >> 5;
  for (uint32_t k = lane; k < nwords; k += 32) {
    const uint4 lo4 = lds128(bstage_s + 32 * k), hi4 = lds128(bstage_s + 32 * k + 16);
    const uint32_t w[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
    uint32_t word = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) word |= (((w[i] & 0x01010101u) * 0x01020408u) >> 24 & 0xFu) << (4 * i);
    // bytes outside [o, end) of the first / last word are stale: mask them off
    const uint32_t lo = 32 * k < o ? o - 32 * k : 0, hi = 32 * k + 32 > end ? end - 32 * k : 32;
    const uint32_t mask = (hi >= 32 ? FULL : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
    word &= mask;
    if (mask == FULL) gbits[g0 + k] = word;
    else if (word) atomicOr(&gbits[g0 + k], word);
  }
  __syncwarp();
}

__device__ __forceinline__ void add_count(uint64_t* slot, uint32_t mine, int lane) {
  const uint32_t s = __reduce_add_sync(FULL, mine);
  if (lane == 0 && s) atomicAdd((unsigned long long*)slot, (unsigned long long)s);
}

// 16 / 4 bytes from an arbitrarily aligned global address (buffers are padded, so the aligned
// words around it are always readable).
__device__ __forceinline__ uint4 load16_unaligned(const uint8_t* p) {
  const uintptr_t a = (uintptr_t)p;
  if ((a & 15u) == 0) return __ldg((const uint4*)p);
  const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(a & 3u) * 8u;
  const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3);
  if (sh == 0) return make_uint4(w0, w1, w2, w3);
  const uint32_t w4 = __ldg(w + 4);
  return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
                    __funnelshift_r(w3, w4, sh));
}
__device__ __forceinline__ uint32_t load4_unaligned(const uint8_t* p) {
  const uintptr_t a = (uintptr_t)p;
  const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(a & 3u) * 8u;
  const uint32_t w0 = __ldg(w);
  if (sh == 0) return w0;
  return __funnelshift_r(w0, __ldg(w + 1), sh);
}

// One selected row's value bytes -> the shared-memory stage (the short-string path).
__device__ __forceinline__ void copy_row_g2s(const uint8_t* sp, uint32_t da, uint32_t n) {
  if ((((uint32_t)(uintptr_t)sp | da) & 3u) == 0) {
    const uint32_t nw = n >> 2;
    if (nw <= 4) {   // up to 16 bytes: straight-line
      uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
      if (nw > 0) w0 = __ldg((const uint32_t*)sp);
      if (nw > 1) w1 = __ldg((const uint32_t*)sp + 1);
      if (nw > 2) w2 = __ldg((const uint32_t*)sp + 2);
      if (nw > 3) w3 = __ldg((const uint32_t*)sp + 3);
      if (nw > 0) sts32(da, w0);
      if (nw > 1) sts32(da + 4, w1);
      if (nw > 2) sts32(da + 8, w2);
      if (nw > 3) sts32(da + 12, w3);
    } else {
#pragma unroll 1
      for (uint32_t i = 0; i < nw; i++) sts32(da + 4 * i, __ldg((const uint32_t*)sp + i));
    }
    const uint32_t done = nw << 2;
#pragma unroll 1
    for (uint32_t i = done; i < n; i++) sts8(da + i, __ldg(sp + i));
  } else {
#pragma unroll 1
    for (uint32_t i = 0; i < n; i++) sts8(da + i, __ldg(sp + i));
  }
}

// Long strings: the warp's output byte range is produced chunk-centric -- each lane builds aligned
// 16-byte output chunks, finding the source row of a byte by binary search over the warp's
// selected rows (s_oo: warp-local output byte offsets, s_src: source byte offsets).
__device__ __noinline__ void copy_long_strings(const uint8_t* __restrict__ sv, uint8_t* gal, uint32_t mis, uint32_t nbytes, uint32_t nrows,
                                               const uint32_t* s_oo, const int32_t* s_src, int lane) {
  const uint32_t end = mis + nbytes;
  const uint32_t nchunks = (end + 15u) >> 4;
#pragma unroll 1
  for (uint32_t ch = lane; ch < nchunks; ch += 32) {
    const uint32_t lo = ch << 4, hi = lo + 16;
    const uint32_t s = lo > mis ? lo : mis, t = hi < end ? hi : end;
    if (s >= t) continue;
    const uint32_t x = s - mis;  // warp-local output byte index of the first byte produced
    uint32_t lo_r = 0, hi_r = nrows;  // first r in (0, nrows] with s_oo[r] > x
    while (lo_r < hi_r) {
      const uint32_t mid = (lo_r + hi_r) >> 1;
      if (s_oo[mid] > x) hi_r = mid; else lo_r = mid + 1;
    }
    uint32_t r = lo_r - 1;  // row holding byte x (never an empty string)
    const bool full = (t - s) == 16u;
    if (full && x + 16u <= s_oo[r + 1]) {
      *(uint4*)(gal + lo) = load16_unaligned(sv + s_src[r] + (x - s_oo[r]));
      continue;
    }
    uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
#pragma unroll 1
    for (uint32_t b = s; b < t;) {
      const uint32_t xb = b - mis;
      while (xb >= s_oo[r + 1]) r++;
      const uint8_t* sp = sv + s_src[r] + (xb - s_oo[r]);
      uint32_t piece, step;
      if (((b & 3u) == 0) && b + 4 <= t && xb + 4 <= s_oo[r + 1]) { piece = load4_unaligned(sp); step = 4; }
      else { piece = (uint32_t)__ldg(sp) << (8u * (b & 3u)); step = 1; }
      const uint32_t wi = (b - lo) >> 2;
      if (wi == 0) w0 |= piece; else if (wi == 1) w1 |= piece; else if (wi == 2) w2 |= piece; else w3 |= piece;
      b += step;
    }
    if (full) {
      *(uint4*)(gal + lo) = make_uint4(w0, w1, w2, w3);
    } else {
      for (uint32_t b = s; b < t; b++) {
        const uint32_t wi = (b - lo) >> 2;
        const uint32_t word = wi == 0 ? w0 : wi == 1 ? w1 : wi == 2 ? w2 : w3;
        gal[b] = (uint8_t)(word >> (8u * (b & 3u)));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
// What every lane carries through the per-warp gather phase.
template <int QPT>
struct LaneCtx {
  int64_t row_base;       // first row of the lane's quad 0; quad q starts at row_base + q * 128
  uint32_t inrange;       // rows that exist (tail tile), 4 bits per quad
  uint32_t stage_s, bstage_s, wstage_bytes;
  int lane, warp;
};

template <int QPT>
__device__ __forceinline__ uint32_t load_bits_all(const uint8_t* __restrict__ bits, int64_t row_base, uint32_t need) {
  if (bits == nullptr) return FULL;
  uint32_t m = 0;
#pragma unroll
  for (int q = 0; q < QPT; q++) {
    if ((need >> (4 * q)) & 0xFu) {
      const int64_t r = row_base + q * 128;
      const uint32_t byte = __ldg(bits + (r >> 3));
      m |= ((byte >> (uint32_t)(r & 4)) & 0xFu) << (4 * q);
    }
  }
  return m;
}

// Gathers output column k for this warp.  `meta` packs the eight small OutDesc fields; under
// CHDB_JIT it is a compile-time constant and BEGIN/END carry the expression's instruction range.
// `vpre`: validity bits of a pass-through column, loaded before the look-back so that their latency
// hides behind it.
template <typename V, int QPT, int BEGIN = -1, int END = -1>
__device__ __forceinline__ void exec_output(const KernelParams& P, const int k, const uint64_t meta, const uint32_t vpre,
                                            const LaneCtx<QPT>& L, const WarpOut<QPT>& wo, const uint32_t (*s_wtot)[kWarps],
                                            const uint64_t* s_excl, const uint8_t* s_pool) {
  const uint32_t o_kind = (uint32_t)meta & 0xFFu, o_type = (uint32_t)(meta >> 8) & 0xFFu, o_width = (uint32_t)(meta >> 16) & 0xFFu;
  const uint32_t o_slot = (uint32_t)(meta >> 24) & 0xFFu, o_begin = (uint32_t)(meta >> 32) & 0xFFu, o_end = (uint32_t)(meta >> 40) & 0xFFu;
  const uint32_t o_utf8 = (uint32_t)(meta >> 48) & 0xFFu, o_count = (uint32_t)(meta >> 56);
  uint8_t* const o_values = (uint8_t*)P.out[k].values;
  uint8_t* const o_validity = P.out[k].validity;
  const uint32_t sel = wo.sel, stage_s = L.stage_s, bstage_s = L.bstage_s;
  const int lane = L.lane;
  uint32_t vbits = FULL;  // validity of this output for the lane's rows
  if (o_kind == OUT_EXPR) {
    uint32_t accm = 0;
    vbits = 0;
    const uint32_t W = o_type == T_BOOL ? 0u : o_width;
    const uint32_t mis = (uint32_t)((wo.base * W) & 15u);
#ifdef CHDB_JIT
#pragma unroll
#else
#pragma unroll 1
#endif
    for (int q = 0; q < QPT; q++) {
      const int64_t qb[1] = {L.row_base + q * 128};
      const uint32_t in4 = (L.inrange >> (4 * q)) & 0xFu, sel4 = (sel >> (4 * q)) & 0xFu;
      V a4[4];
      uint32_t m4, v4;
      // `sel` as the active mask: checked arithmetic only sees rows that survived the filter
      run_program<V, 1, BEGIN, END>(P, (int)o_begin, (int)o_end, qb, in4, sel4, s_pool, a4, m4, v4);
      accm |= (m4 & 0xFu) << (4 * q);
      vbits |= (v4 & 0xFu) << (4 * q);
      uint32_t rk = wo.rank[0];
#pragma unroll
      for (int t = 1; t < QPT; t++)
        if (t == q) rk = wo.rank[t];     // static indexing keeps the ranks in registers
      const uint32_t a = stage_s + mis + rk * W;
      if (W == 4) stage_regs<4, V>(a, sel4, a4);
      else if (W == 8) stage_regs<8, V>(a, sel4, a4);
      else if (W == 2) stage_regs<2, V>(a, sel4, a4);
      else if (W == 1) stage_regs<1, V>(a, sel4, a4);
    }
    if (o_type == T_BOOL) {
      compact_bits<QPT>(accm, wo, bstage_s, (uint32_t*)o_values, lane);
    } else {
      __syncwarp();
      uint8_t* gd = o_values + ((wo.base * W) & ~15ull);
      if (W == 4) warp_writeout<4>(stage_s, gd, mis, wo.count * 4, lane);
      else if (W == 8) warp_writeout<8>(stage_s, gd, mis, wo.count * 8, lane);
      else if (W == 2) warp_writeout<2>(stage_s, gd, mis, wo.count * 2, lane);
      else warp_writeout<1>(stage_s, gd, mis, wo.count, lane);
      __syncwarp();
    }
  } else {
    const ColumnDesc& c = P.in[o_slot];
    vbits = vpre;
    if (o_type == T_BOOL) {
      const uint32_t vals = load_bits_all<QPT>((const uint8_t*)c.values, L.row_base, sel);
      compact_bits<QPT>(vals, wo, bstage_s, (uint32_t*)o_values, lane);
    } else if (o_type == T_UTF8) {
      // -- offsets: running sum of the selected lengths, restarted at 0 for the output.
      //    The warp's stage holds the rebuilt offsets in its first kWarpRows * 4 + 16 bytes and the
      //    value bytes (short strings) behind them, so both are written out once per warp slice.
      constexpr uint32_t kOffStage = (uint32_t)(32 * 4 * QPT) * 4u + 16u;
      const int32_t* __restrict__ off = c.offsets;
      const uint8_t* __restrict__ sv = (const uint8_t*)c.values;
      uint32_t bytes_before = 0;
#pragma unroll
      for (int w = 0; w < kWarps; w++)
        if (w < L.warp) bytes_before += s_wtot[1 + o_utf8][w];
      const uint32_t warp_bytes = s_wtot[1 + o_utf8][L.warp];
      const uint64_t byte_base = s_excl[1 + o_utf8] + bytes_before;   // output byte offset of this warp's first value
      const uint32_t omis = (uint32_t)((wo.base * 4) & 15u);
      const uint32_t mis = (uint32_t)(byte_base & 15u);
      const uint32_t str_s = stage_s + kOffStage;
      const bool staged = mis + warp_bytes + 16u <= L.wstage_bytes - kOffStage;   // short strings fit the stage
      uint32_t* s_oo = (uint32_t*)__cvta_shared_to_generic(str_s);                // long strings: per-row tables instead
      int32_t* s_src = (int32_t*)(s_oo + 32 * 4 * QPT + 4);
      uint32_t run = 0;   // bytes of the quads handled so far
#pragma unroll
      for (int q = 0; q < QPT; q++) {
        const uint32_t s4 = (sel >> (4 * q)) & 0xFu;
        uint32_t len[4] = {0u, 0u, 0u, 0u}, boff[4];
        int32_t src[4] = {0, 0, 0, 0};
        if (s4) {
          const int4 a = __ldg((const int4*)(off + L.row_base + q * 128));
          const int a4 = __ldg(off + L.row_base + q * 128 + 4);
          src[0] = a.x; src[1] = a.y; src[2] = a.z; src[3] = a.w;
          if (s4 & 1u) len[0] = (uint32_t)(a.y - a.x);
          if (s4 & 2u) len[1] = (uint32_t)(a.z - a.y);
          if (s4 & 4u) len[2] = (uint32_t)(a.w - a.z);
          if (s4 & 8u) len[3] = (uint32_t)(a4 - a.w);
        }
        uint32_t tot;
        uint32_t bo = run + warp_excl_scan(len[0] + len[1] + len[2] + len[3], lane, tot);
        run += tot;
        uint32_t newoff[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          boff[i] = bo;                                   // warp-local output byte offset of the row
          newoff[i] = (uint32_t)byte_base + bo;
          bo += len[i];
        }
        stage_regs<4, uint32_t>(stage_s + omis + wo.rank[q] * 4, s4, newoff);
        if (staged) {
#pragma unroll
          for (int i = 0; i < 4; i++)
            if (((s4 >> i) & 1u) && len[i] != 0) copy_row_g2s(sv + src[i], str_s + mis + boff[i], len[i]);
        } else {
          uint32_t r = wo.rank[q];
#pragma unroll
          for (int i = 0; i < 4; i++)
            if ((s4 >> i) & 1u) { s_oo[r] = boff[i]; s_src[r] = src[i]; r++; }
        }
      }
      if (!staged && lane == 0) s_oo[wo.count] = warp_bytes;
      __syncwarp();
      warp_writeout<4>(stage_s, (uint8_t*)P.out[k].offsets + ((wo.base * 4) & ~15ull), omis, wo.count * 4, lane);
      uint8_t* gal = o_values + (byte_base - mis);
      if (staged) warp_writeout<1>(str_s, gal, mis, warp_bytes, lane);
      else copy_long_strings(sv, gal, mis, warp_bytes, wo.count, s_oo, s_src, lane);
      __syncwarp();
    } else if (o_width == 16) {
      const uint4* __restrict__ src = (const uint4*)c.values;
#pragma unroll
      for (int q = 0; q < QPT; q++) {
        uint32_t a = stage_s + wo.rank[q] * 16;
#pragma unroll
        for (int i = 0; i < 4; i++)
          if ((sel >> (4 * q + i)) & 1u) { sts128(a, __ldg(src + L.row_base + q * 128 + i)); a += 16; }
      }
      __syncwarp();
      uint4* dst = (uint4*)o_values + wo.base;
      for (uint32_t r = lane; r < wo.count; r += 32) dst[r] = lds128(stage_s + r * 16);
      __syncwarp();
    } else if (o_width == 4) {
      gather_fixed<4, QPT>((const uint8_t*)c.values, o_values, L.row_base, wo, stage_s, lane);
    } else if (o_width == 8) {
      gather_fixed<8, QPT>((const uint8_t*)c.values, o_values, L.row_base, wo, stage_s, lane);
    } else if (o_width == 2) {
      gather_fixed<2, QPT>((const uint8_t*)c.values, o_values, L.row_base, wo, stage_s, lane);
    } else {
      gather_fixed<1, QPT>((const uint8_t*)c.values, o_values, L.row_base, wo, stage_s, lane);
    }
  }
  if (o_validity != nullptr) {
    compact_bits<QPT>(vbits, wo, bstage_s, (uint32_t*)o_validity, lane);
    add_count(P.counts + o_count, (uint32_t)__popc(sel & ~vbits), lane);
  }
}

#ifdef CHDB_JIT
template <typename V, int QPT, int K, int N>
__device__ __forceinline__ void outputs_range(const KernelParams& P, const uint32_t (&vpre)[kMaxOutCols], const LaneCtx<QPT>& L,
                                              const WarpOut<QPT>& wo, const uint32_t (*s_wtot)[kWarps], const uint64_t* s_excl,
                                              const uint8_t* s_pool) {
  if constexpr (K < N) {
    constexpr uint64_t meta = chdb_jit::kOutMeta[K];
    exec_output<V, QPT, (int)((meta >> 32) & 0xFFu), (int)((meta >> 40) & 0xFFu)>(P, K, meta, vpre[K], L, wo, s_wtot, s_excl, s_pool);
    outputs_range<V, QPT, K + 1, N>(P, vpre, L, wo, s_wtot, s_excl, s_pool);
  }
}
#endif

// Pulls one tile's slice of every input buffer towards L2 (one 128-byte line per thread).
__device__ __forceinline__ void prefetch_tile(const KernelParams& P, int64_t row0, int32_t tile_rows, int tid) {
  CHDB_STATIC_UNROLL
  for (int s = 0; s < CHDB_N_IN; s++) {
    const ColumnDesc& c = P.in[s];
    const uint32_t ctype = CHDB_COL_TYPE(P, s);
    const int w = ctype == T_UTF8 ? 4 : (int)CHDB_COL_WIDTH(P, s);
    const uint8_t* v = (const uint8_t*)(ctype == T_UTF8 ? (const void*)c.offsets : c.values);
    if (w > 0) {
      for (int b = tid * 128; b < tile_rows * w; b += kThreads * 128) prefetch_l2(v + row0 * w + b);
    } else if (tid < 4 && tid * 1024 < tile_rows) {
      prefetch_l2(v + (row0 >> 3) + tid * 128);   // Boolean values: rows / 8 bytes
    }
    if (c.validity != nullptr && tid >= 32 && tid < 36 && (tid - 32) * 1024 < tile_rows)
      prefetch_l2(c.validity + (row0 >> 3) + (tid - 32) * 128);
  }
}

template <typename V, int QPT>
__device__ __forceinline__ void filter_project_body(const KernelParams& P) {
  constexpr int R = 4 * QPT;
  constexpr int T = kThreads * R;
  constexpr int WR = 32 * R;                    // rows per warp slice
  static_assert(R <= 32, "selection masks are 32-bit");
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ uint32_t s_tile;
  __shared__ uint32_t s_wtot[1 + kMaxOutCols][kWarps];   // per-warp totals: [0] rows, [1 + u] bytes of Utf8 output u
  __shared__ uint64_t s_excl[1 + kMaxOutCols];           // tile prefixes from the look-back
  __shared__ uint8_t s_pool[kStrPoolBytes];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool has_pred = CHDB_PRED_END > CHDB_PRED_BEGIN;

  LaneCtx<QPT> L;
  L.lane = lane;
  L.warp = warp;
  L.wstage_bytes = (uint32_t)P.stage_bytes;                                           // per warp
  L.stage_s = smem_u32(smem) + warp * L.wstage_bytes;                                 // this warp's staging slice
  L.bstage_s = smem_u32(smem) + kWarps * L.wstage_bytes + warp * kWarpBitStage;       // one byte per output row

  if (tid == 0) s_tile = has_pred ? atomicAdd(P.ticket, 1u) : blockIdx.x;
  if (tid < kStrPoolBytes) s_pool[tid] = (uint8_t)P.strpool[tid];
  __syncthreads();
  const uint32_t tile = s_tile;
  const int64_t row0 = (int64_t)tile * T;
  const int32_t tile_rows = (int32_t)(row0 + T < P.num_rows ? T : P.num_rows - row0);
  CHDB_STAMP(0);

  // ---- 0. prefetch: the tile that will start about one wave from now, so that by then its
  //         columns wait in L2; the first wave has nobody to do that for it and fetches its own ----
  {
    const int64_t ahead = (int64_t)tile + P.prefetch_tiles;
    if (P.prefetch_tiles > 0 && ahead < P.num_tiles) {
      const int64_t r0 = ahead * T;
      prefetch_tile(P, r0, (int32_t)(r0 + T < P.num_rows ? T : P.num_rows - r0), tid);
    }
    if ((int64_t)tile < P.prefetch_tiles || P.prefetch_tiles <= 0) prefetch_tile(P, row0, tile_rows, tid);
  }

  // rows of this lane: quad q covers rows  row0 + warp * WR + q * 128 + lane * 4 .. + 3
  L.row_base = row0 + warp * WR + lane * 4;
  L.inrange = 0;
#pragma unroll
  for (int q = 0; q < QPT; q++) {
    const int left = tile_rows - (warp * WR + q * 128 + lane * 4);
    L.inrange |= (left >= 4 ? 0xFu : left <= 0 ? 0u : ((1u << left) - 1u)) << (4 * q);
  }
  const uint32_t inrange = L.inrange;

  // ---- 1. predicate -> selection mask (one quad at a time: small register footprint) ---------
  uint32_t sel = inrange;
  if (has_pred) {
    sel = 0;
#ifdef CHDB_JIT
#pragma unroll
#else
#pragma unroll 1
#endif
    for (int q = 0; q < QPT; q++) {
      const int64_t qb[1] = {L.row_base + q * 128};
      const uint32_t in4 = (inrange >> (4 * q)) & 0xFu;
      V acc[4];
      uint32_t accm, accv;
#ifdef CHDB_JIT
      run_program<V, 1, chdb_jit::kPredBegin, chdb_jit::kPredEnd>(P, 0, 0, qb, in4, in4, s_pool, acc, accm, accv);
#else
      run_program<V, 1>(P, P.pred_begin, P.pred_end, qb, in4, in4, s_pool, acc, accm, accv);
#endif
      sel |= (accm & accv & in4) << (4 * q);  // NULL predicate rows are dropped (arrow-select filter)
    }
  }

  CHDB_STAMP(1);
  // ---- 2. rank the selected rows inside the warp; publish the warp totals --------------------
  WarpOut<QPT> wo;
  wo.sel = sel;
  uint32_t warp_rows = 0;
#pragma unroll
  for (int q = 0; q < QPT; q++) {
    uint32_t tot;
    wo.rank[q] = warp_rows + warp_excl_scan((uint32_t)__popc((sel >> (4 * q)) & 0xFu), lane, tot);
    warp_rows += tot;
  }
  wo.count = warp_rows;
  if (lane == 0) s_wtot[0][warp] = warp_rows;

  // selected value bytes per Utf8 output (and a prefetch of exactly those bytes)
  CHDB_STATIC_UNROLL
  for (int k = 0; k < CHDB_N_OUT; k++) {
    const uint64_t meta = CHDB_OUT_META(P, k);
    const uint32_t o_utf8 = (uint32_t)(meta >> 48) & 0xFFu, o_slot = (uint32_t)(meta >> 24) & 0xFFu;
    if (o_utf8 == 0xFFu) continue;   // uniform branch
    const int32_t* __restrict__ off = P.in[o_slot].offsets;
    const uint8_t* __restrict__ sv = (const uint8_t*)P.in[o_slot].values;
    uint32_t bytes = 0;
#pragma unroll
    for (int q = 0; q < QPT; q++) {
      const uint32_t s4 = (sel >> (4 * q)) & 0xFu;
      if (s4) {
        const int4 a = __ldg((const int4*)(off + L.row_base + q * 128));
        const int a4 = __ldg(off + L.row_base + q * 128 + 4);
        if (s4 & 1u) bytes += (uint32_t)(a.y - a.x);
        if (s4 & 2u) bytes += (uint32_t)(a.z - a.y);
        if (s4 & 4u) bytes += (uint32_t)(a.w - a.z);
        if (s4 & 8u) bytes += (uint32_t)(a4 - a.w);
        const uintptr_t pa = (uintptr_t)(sv + a.x), pe = (uintptr_t)(sv + a4);
        prefetch_l2((const void*)pa);
        for (uintptr_t line = (pa + 128) & ~(uintptr_t)127; line < pe; line += 128) prefetch_l2((const void*)line);
      }
    }
    const uint32_t wbytes = __reduce_add_sync(FULL, bytes);
    if (lane == 0) s_wtot[1 + o_utf8][warp] = wbytes;
  }
  // validity bits of the pass-through outputs: requested now, consumed after the look-back
  uint32_t vpre[kMaxOutCols];
  CHDB_STATIC_UNROLL
  for (int k = 0; k < CHDB_N_OUT; k++) {
    const uint64_t meta = CHDB_OUT_META(P, k);
    vpre[k] = FULL;
    if (((uint32_t)meta & 0xFFu) == OUT_PASS) vpre[k] = load_bits_all<QPT>(P.in[(uint32_t)(meta >> 24) & 0xFFu].validity, L.row_base, sel);
  }
  __syncthreads();
  CHDB_STAMP(2);

  // ---- 3. tile prefixes: warp qi runs the look-back for quantity qi ---------------------------
  const int nq = 1 + CHDB_N_UTF8;
  if (has_pred) {
    for (int qi = warp; qi < nq; qi += kWarps) {
      uint64_t agg = 0;
#pragma unroll
      for (int w = 0; w < kWarps; w++) agg += s_wtot[qi][w];
      const uint64_t excl = lookback(P.tile_desc + (size_t)qi * P.num_tiles, tile, agg, lane);
      if (lane == 0) {
        s_excl[qi] = excl;
        if (tile == (uint32_t)P.num_tiles - 1) P.counts[qi] = excl + agg;  // totals
      }
    }
  } else if (tid == 0) {
    s_excl[0] = (uint64_t)row0;
    if (tile == (uint32_t)P.num_tiles - 1) P.counts[0] = (uint64_t)P.num_rows;
  }
  __syncthreads();
  CHDB_STAMP(3);
  // from here on every warp works alone
  uint32_t rows_before = 0, tile_count = 0;
#pragma unroll
  for (int w = 0; w < kWarps; w++) {
    const uint32_t t = s_wtot[0][w];
    if (w < warp) rows_before += t;
    tile_count += t;
  }
  wo.base = s_excl[0] + rows_before;
  if (tile == (uint32_t)P.num_tiles - 1 && tid < CHDB_N_OUT) {
    // closing Utf8 offset: offsets[total_rows] = total_bytes (also covers an empty result)
    const OutDesc& o = P.out[tid];
    if (o.utf8_index != 0xFFu) {
      uint64_t tb = s_excl[1 + o.utf8_index];
      for (int w = 0; w < kWarps; w++) tb += s_wtot[1 + o.utf8_index][w];
      o.offsets[s_excl[0] + tile_count] = (int32_t)tb;
    }
  }
  if (wo.count == 0) return;  // warp-uniform; no block-wide barrier follows

  // ---- 4. gather every output column (per warp) ------------------------------------------------
#ifdef CHDB_JIT
  outputs_range<V, QPT, 0, chdb_jit::kNumOut>(P, vpre, L, wo, s_wtot, s_excl, s_pool);
#else
#pragma unroll 1
  for (int k = 0; k < P.n_out; k++) exec_output<V, QPT>(P, k, CHDB_OUT_META(P, k), vpre[k], L, wo, s_wtot, s_excl, s_pool);
#endif
  CHDB_STAMP(4);
}

}  // namespace

#ifndef CHDB_JIT
template <typename V, int QPT>
__global__ void __launch_bounds__(kThreads, 3) filter_project_kernel(const __grid_constant__ KernelParams P) {
  filter_project_body<V, QPT>(P);
}
#endif

}  // namespace chdb

#ifdef CHDB_JIT
// one specialised kernel per NVRTC module, found by its unmangled name
extern "C" __global__ void __launch_bounds__(chdb::kThreads, CHDB_JIT_MIN_BLOCKS) chdb_jit_kernel(const __grid_constant__ chdb::KernelParams P) {
  chdb::filter_project_body<chdb_jit::Container, chdb::kQuadsPerThread>(P);
}
#endif
