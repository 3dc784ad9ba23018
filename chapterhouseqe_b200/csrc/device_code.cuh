// Filter / projection / compaction on sm_100a: one fused single-pass kernel per batch.
//
// A CTA owns one tile of 1024 rows.  Warp 0 brings the tile's slice of every column the program touches
// into shared memory with TMA bulk copies; every warp then evaluates the predicate bytecode on its
// slices (accumulator in registers, 128-bit shared-memory loads; compute_value.rs), keeps the
// selection bits in registers and posts the slice counts; one warp per scanned quantity (selected
// rows, selected value bytes per Utf8 output) publishes the tile's aggregate and obtains the tile's
// batch-wide exclusive prefix by decoupled look-back over its predecessors' descriptors; finally every
// warp writes the selected rows of its slices straight to their final positions (neighbouring lanes
// hit neighbouring addresses); validity / Boolean bits are assembled per warp in shared memory and
// written as words; Utf8 offsets restart at 0; projection expressions are evaluated under the
// selection mask, so checked-integer errors are raised for surviving rows only
// (filter_record.rs:37, record_projection.rs).
//
// Every input byte is read from HBM once and every output byte written once.  Latency (the tile's
// loads, the look-back's L2 round trips) is hidden the way GPUs hide latency: by 5-6 other tiles
// resident on the same SM, each in a different phase.  (Round 1 ran one persistent CTA per SM, first
// with a per-tile look-back on its critical path -- 29 % of HBM peak -- then as three kernels with a
// second read of the predicate columns -- 45 %.)
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
#define CHDB_LONG_STRINGS(P) (chdb_jit::kLongStrings != 0)
#define CHDB_OUT_HAS_VALIDITY(P, k) (chdb_jit::kOutHasValidity[k] != 0)
#define CHDB_EARLY_COUNTS(P) (chdb_jit::kEarlyCounts != 0)
#define CHDB_IN_HAS_VALIDITY(P, s, ptr) (chdb_jit::kInHasValidity[s] != 0)
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
#define CHDB_LONG_STRINGS(P) (P.long_strings != 0)
#define CHDB_OUT_HAS_VALIDITY(P, k) (P.out[k].validity != nullptr)
#define CHDB_EARLY_COUNTS(P) (P.early_counts != 0)
#define CHDB_IN_HAS_VALIDITY(P, s, ptr) ((ptr) != nullptr)
#endif

template <typename V> struct Cont;
template <> struct Cont<uint32_t> { static constexpr bool k64 = false; };
template <> struct Cont<uint64_t> { static constexpr bool k64 = true; };

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
__device__ __noinline__ void report_error(const KernelParams& P, uint32_t order, int64_t row, uint32_t code) {
  unsigned long long packed = ((unsigned long long)order << 56) | (((unsigned long long)row & 0xFFFFFFFFFFFFull) << 8) | code;
  atomicMax((unsigned long long*)P.b.error_word, ~packed);
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

template <int QPT>
__device__ __forceinline__ uint32_t load_bits_nn(const uint8_t* __restrict__ bits, const int64_t (&qbase)[QPT], uint32_t need) {
  uint32_t m = 0;
#pragma unroll
  for (int q = 0; q < QPT; q++) {
    if ((need >> (4 * q)) & 0xFu) {
      const uint32_t byte = bits[qbase[q] >> 3];
      m |= ((byte >> (uint32_t)(qbase[q] & 4)) & 0xFu) << (4 * q);
    }
  }
  return m;
}
template <int QPT>
__device__ __forceinline__ uint32_t load_bits(const uint8_t* __restrict__ bits, const int64_t (&qbase)[QPT], uint32_t need) {
  if (bits == nullptr) return FULL;
  return load_bits_nn<QPT>(bits, qbase, need);
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
        const uint4 x = *(const uint4*)(base + qbase[q] * 4);
        b[4 * q + 0] = x.x; b[4 * q + 1] = x.y; b[4 * q + 2] = x.z; b[4 * q + 3] = x.w;
      }
      break;
    case T_I64: case T_U64: case T_F64:
      if constexpr (Cont<V>::k64) {
#pragma unroll
        for (int q = 0; q < QPT; q++) {
          CHDB_SKIP_QUAD(q)
          const uint4 x = *(const uint4*)(base + qbase[q] * 8);
          const uint4 y = *(const uint4*)(base + qbase[q] * 8 + 16);
          b[4 * q + 0] = x.x | ((uint64_t)x.y << 32); b[4 * q + 1] = x.z | ((uint64_t)x.w << 32);
          b[4 * q + 2] = y.x | ((uint64_t)y.y << 32); b[4 * q + 3] = y.z | ((uint64_t)y.w << 32);
        }
      }
      break;
    case T_I16: case T_U16:
#pragma unroll
      for (int q = 0; q < QPT; q++) {
        CHDB_SKIP_QUAD(q)
        const uint2 x = *(const uint2*)(base + qbase[q] * 2);
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
        const uint32_t x = *(const uint32_t*)(base + qbase[q]);
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
__device__ __noinline__ uint32_t cmp_utf8(const ColumnDesc* cols, const Instr in, const QuadBases<QPT> qb, uint32_t inrange,
                                          const uint8_t* s_pool, uint32_t* valid_out) {
  const int64_t (&qbase)[QPT] = qb.v;
  uint32_t valid;
  const uint32_t slot_a = in.slot, slot_b = (uint32_t)(in.imm >> 56);
  const uint32_t pool_off = (uint32_t)in.imm, pool_len = (uint32_t)(in.imm >> 32) & 0xFFFFFFu;
  valid = FULL;
  if (slot_a != 0xFFu) valid &= load_bits<QPT>(cols[slot_a].validity, qbase, inrange);
  if (slot_b != 0xFFu) valid &= load_bits<QPT>(cols[slot_b].validity, qbase, inrange);
  uint32_t lt = 0, eq = 0;
#pragma unroll 1
  for (int j = 0; j < 4 * QPT; j++) {
    if (!((inrange >> j) & 1u)) continue;
    const int64_t row = qbase[j >> 2] + (j & 3);
    const uint8_t *pa, *pb;
    int la, lb;
    if (slot_a != 0xFFu) {
      const int32_t* off = cols[slot_a].offsets;
      const int o0 = off[row], o1 = off[row + 1];
      pa = (const uint8_t*)cols[slot_a].values + o0;
      la = o1 - o0;
    } else {
      pa = s_pool + pool_off;
      la = (int)pool_len;
    }
    if (slot_b != 0xFFu) {
      const int32_t* off = cols[slot_b].offsets;
      const int o0 = off[row], o1 = off[row + 1];
      pb = (const uint8_t*)cols[slot_b].values + o0;
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
__device__ __forceinline__ void fetch_operand(const KernelParams& P, const ColumnDesc* cols, const Instr& in, const int64_t (&qbase)[QI], uint32_t inrange,
                                              const Spill<V, QI>& stk, V (&b)[4 * QI], uint32_t& bm, uint32_t& bv) {
  constexpr int R = 4 * QI;
  if (in.src == SRC_COL) {
    const ColumnDesc& c = cols[in.slot];
    // (known at compile time in the specialised kernels: no branch between this load and the ones before it)
    bv = CHDB_IN_HAS_VALIDITY(P, in.slot, c.validity) ? load_bits_nn<QI>(c.validity, qbase, inrange) : FULL;
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
__device__ __forceinline__ void exec_instr(const KernelParams& P, const ColumnDesc* cols, const Instr in, const int64_t (&qbase)[QI], uint32_t inrange,
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
        fetch_operand<V, QI>(P, cols, in, qbase, inrange, stk, acc, accm, accv);   // straight into the accumulator
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
        fetch_operand<V, QI>(P, cols, in, qbase, inrange, stk, b, bm, bv);
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
        fetch_operand<V, QI>(P, cols, in, qbase, inrange, stk, b, bm, bv);
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
        fetch_operand<V, QI>(P, cols, in, qbase, inrange, stk, b, bm, bv);
      }
      if (in.flags & OPF_KLEENE) {   // and_kleene / or_kleene: a false (true) side decides an AND (OR) whatever the other is
        const uint32_t at = accm & accv, bt = bm & bv, af = ~accm & accv, bf = ~bm & bv;
        if (in.op == OP_AND) { accm = at & bt; accv = (accv & bv) | af | bf; }
        else { accm = at | bt; accv = (accv & bv) | at | bt; }
      } else {
        accm = in.op == OP_AND ? (accm & bm) : (accm | bm);
        accv &= bv;
      }
      break;
    }
    case OP_NEG:
#pragma unroll
      for (int j = 0; j < R; j++) acc[j] ^= in.type == T_F32 ? (V)0x80000000u : (V)(1ull << (Cont<V>::k64 ? 63 : 31));
      break;
    case OP_NOT: accm = ~accm; break;
    case OP_ISNULL: {
      uint32_t v = accv;
      if (in.src == SRC_COL) v = load_bits<QI>(cols[in.slot].validity, qbase, inrange);
      accm = (in.flags & OPF_NEGATE) ? v : ~v;
      accv = FULL;
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
      accm = cmp_utf8<QI>(cols, in, qb, inrange, s_pool, &v);
      accv = v;
      break;
    }
    default: break;
  }
}

#ifdef CHDB_JIT
template <typename V, int QI, int PC, int END>
__device__ __forceinline__ void run_range(const KernelParams& P, const ColumnDesc* cols, const int64_t (&qbase)[QI], uint32_t inrange, uint32_t active,
                                          const uint8_t* s_pool, Spill<V, QI>& stk, V (&acc)[4 * QI], uint32_t& accm,
                                          uint32_t& accv) {
  if constexpr (PC < END) {
    constexpr Instr in = chdb_jit::kInstrs[PC];
    exec_instr<V, QI>(P, cols, in, qbase, inrange, active, s_pool, stk, acc, accm, accv);
    run_range<V, QI, PC + 1, END>(P, cols, qbase, inrange, active, s_pool, stk, acc, accm, accv);
  }
}
#endif

// BEGIN/END >= 0: instruction range known at compile time (CHDB_JIT); otherwise [begin, end).
template <typename V, int QI, int BEGIN = -1, int END = -1>
__device__ __forceinline__ void run_program(const KernelParams& P, const ColumnDesc* cols, int begin, int end, const int64_t (&qbase)[QI], uint32_t inrange,
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
    run_range<V, QI, BEGIN, END>(P, cols, qbase, inrange, active, s_pool, stk, acc, accm, accv);
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
    exec_instr<V, QI>(P, cols, in, qbase, inrange, active, s_pool, stk, acc, accm, accv);
  }
}

// ------------------------------------------------------------------------------------------
// scans
// ------------------------------------------------------------------------------------------
// Waits are bounded: a wait that lasts about a second is a bug (or a lost dependency), and a diagnostic plus a
// trapped launch (the host sees a CUDA error) beats a hung stream.
__device__ __noinline__ void wait_timed_out(const char* what, uint32_t a, uint32_t b) {
#ifndef CHDB_JIT   // (the run-time specialised kernels only trap: a printf call site costs them registers on the hot path)
  printf("[chdb] %s timed out: block %d warp %d (%u, %u)\n", what, (int)blockIdx.x, (int)(threadIdx.x >> 5), a, b);
#endif
  __trap();
}
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

// Decoupled look-back (Merrill & Garland) on packed {flag:2 | value:62} descriptors.
// Aggregates are published with a fire-and-forget red.max: {PREFIX | v} > {AGG | v} > 0 as unsigned
// numbers, so a descriptor can only move forward whatever order the updates reach L2 in, and the
// publisher needs no fence (the word carries its own flag; nothing else is read through it).
constexpr uint64_t kFlagAgg = 1ull << 62, kFlagPrefix = 2ull << 62, kValueMask = (1ull << 62) - 1;
__device__ __forceinline__ void publish_descriptor(uint64_t* d, uint64_t flag, uint64_t value) {
  const uint64_t v = flag | value;
  asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(d), "l"(v) : "memory");
}
// Relaxed, gpu-scope loads served by L2 (not `volatile`: strong system-scope loads complete one at a time).
__device__ __forceinline__ uint64_t load_descriptor(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// Exclusive prefix of chunk `chunk`, by one full warp; the chunk's own aggregate is already public.
// A hop reads kLookBackRows rows of 32 consecutive descriptors (lane l of row j: predecessor j * 32 + l), so one
// load instruction touches 256 contiguous bytes.  Every resident CTA reads the same few hundred descriptors at
// about the same time: what bounds a look-back under load is the number of sectors requested from the two or three
// L2 slices holding them (lane-major indexing -- 32 sectors per instruction -- measured 3-5x slower).
constexpr int kLookBackRows = 6;   // 192 predecessors per hop, all loads of a hop in flight together
__device__ __forceinline__ uint64_t lookback(uint64_t* d, uint32_t chunk, uint64_t agg, int lane) {
  if (chunk == 0) return 0;
  uint64_t part = 0;   // this lane's share of the sum of everything nearer than the nearest known prefix
  int64_t base = (int64_t)chunk - 1;
  bool done = false;
  while (!done) {
    uint64_t v[kLookBackRows];
#pragma unroll
    for (int j = 0; j < kLookBackRows; j++) {
      const int64_t idx = base - (j * 32 + lane);
      v[j] = 2ull << 62;   // chunks "before 0" contribute an inclusive prefix of 0
      if (idx >= 0) v[j] = load_descriptor(d + idx);
    }
#pragma unroll
    for (int j = 0; j < kLookBackRows; j++) {
      if (!done) {
        const int64_t idx = base - (j * 32 + lane);
        uint32_t spins = 0;
        while (true) {   // every descriptor of the row must at least carry its aggregate
          const bool missing = (v[j] >> 62) == 0;
          if (!__any_sync(FULL, missing)) break;
          if (missing) {
            __nanosleep(40);
            v[j] = load_descriptor(d + idx);
            if (++spins > (1u << 22)) wait_timed_out("look-back", chunk, (uint32_t)idx);
          }
        }
        const uint32_t pm = __ballot_sync(FULL, (v[j] >> 62) == 2);
        if (pm) {
          const int first = __ffs(pm) - 1;  // lane holding the nearest predecessor that already knows its prefix
          if (lane <= first) part += v[j] & kValueMask;
          done = true;
        } else {
          part += v[j] & kValueMask;
        }
      }
    }
    base -= 32 * kLookBackRows;
  }
  const uint64_t excl = warp_sum64(part);
  if (lane == 0) publish_descriptor(d + chunk, kFlagPrefix, excl + agg);
  return excl;
}


// ------------------------------------------------------------------------------------------
// shared-memory plumbing: mbarriers, TMA bulk copies, programmatic dependent launch
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { uint64_t a;
  asm("cvta.to.shared.u64 %0, %1;" : "=l"(a) : "l"(p));
  return (uint32_t)a; }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(64);
    if (++spins > (1u << 22)) wait_timed_out("mbarrier wait", bar, parity);
  }
}
// global -> shared bulk copy (TMA, 1-D); dst, src and bytes are multiples of 16
__device__ __forceinline__ void tma_load(uint32_t dst_s, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_s), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// Programmatic dependent launch: the stream kernel is launched as the programmatic dependent of the small
// kernel that zeroes its workspace, so its CTAs are scheduled (and run their prologue) while that kernel --
// and the previous batch's stream kernel before it -- drain; grid_dependency_wait() returns once the
// predecessor has completed and its writes are visible.
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// [sel4 << 4 | bits4] -> the bits of bits4 at the set positions of sel4, packed (a 4-bit pext), four entries per word;
// every CTA with bit-packed outputs copies it into shared memory.
__device__ const uint32_t kPext4Words[64] = {
    0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x01000100u, 0x01000100u, 0x01000100u, 0x01000100u,
    0x01010000u, 0x01010000u, 0x01010000u, 0x01010000u, 0x03020100u, 0x03020100u, 0x03020100u, 0x03020100u,
    0x00000000u, 0x01010101u, 0x00000000u, 0x01010101u, 0x01000100u, 0x03020302u, 0x01000100u, 0x03020302u,
    0x01010000u, 0x03030202u, 0x01010000u, 0x03030202u, 0x03020100u, 0x07060504u, 0x03020100u, 0x07060504u,
    0x00000000u, 0x00000000u, 0x01010101u, 0x01010101u, 0x01000100u, 0x01000100u, 0x03020302u, 0x03020302u,
    0x01010000u, 0x01010000u, 0x03030202u, 0x03030202u, 0x03020100u, 0x03020100u, 0x07060504u, 0x07060504u,
    0x00000000u, 0x01010101u, 0x02020202u, 0x03030303u, 0x01000100u, 0x03020302u, 0x05040504u, 0x07060706u,
    0x01010000u, 0x03030202u, 0x05050404u, 0x07070606u, 0x03020100u, 0x07060504u, 0x0b0a0908u, 0x0f0e0d0cu};

// What the output code needs from the CTA's shared memory besides the tile itself.
struct TileShared {
  const uint8_t* pool;      // string literals (kernel parameter space)
  const uint8_t* pext4;     // [sel4 << 4 | bits4] -> the selected bits, packed
  uint32_t* nulls;          // [n_out] NULLs written by this CTA
};

// ------------------------------------------------------------------------------------------
// Loading the tile: warp 0, lane s for input slot s.  The tile's slice of every staged buffer goes into
// the stage with a TMA bulk copy (completion is counted in bytes on the `full` mbarrier) and the
// column's biased pointer into the tile's column table; buffers that are used but not staged get a
// bulk L2 prefetch.  Utf8 value bytes start at offsets[row0]: those copies are issued in a second
// step (the two bounds are a global load away), after the fixed-size ones are already in flight.
// ------------------------------------------------------------------------------------------
// What load_tile's second step (the Utf8 value bytes) needs from its first.
struct TileValues {
  const uint8_t* values;   // the lane's Utf8 column values (nullptr: the lane has none to load)
  uint32_t o0, o1;         // offsets[row0], offsets[row0 + tile_rows]: requested in step 1, first used in step 2
  uint32_t cap, vso;
};
__device__ __forceinline__ TileValues load_tile(const KernelParams& P, const TilePlan& TP, ColumnDesc* cols, uint8_t* stage, uint32_t full,
                                                int64_t row0, uint32_t tile_rows, int lane, int warp) {
  const uint32_t ss = smem_u32(stage);
  const uint32_t bits_bytes = (((tile_rows + 7u) >> 3) + 15u) & ~15u;
  uint32_t nb[3] = {0, 0, 0}, so[3] = {0, 0, 0}, pf[3] = {0, 0, 0};   // 0: validity, 1: offsets, 2: values
  const uint8_t* src[3] = {nullptr, nullptr, nullptr};
  bool utf8_values = false;
  uint32_t cap = 0, vso = 0;
  ColumnDesc c;
  c.values = nullptr; c.validity = nullptr; c.offsets = nullptr; c.type = 0; c.width = 0;
  // (every warp of the CTA issues the copies of its share of the slots -- one bulk copy costs the issuing warp
  //  ~100 cycles, and nothing else can start before the last one is on its way; each warp arrives once on `full`)
  if (lane < CHDB_N_IN && (lane % kWarps) == warp) {
    c = P.in[lane];
    const StageSlot sl = TP.slot[lane];
    const uint32_t use = TP.use[lane];
    ColumnDesc t = c;   // this tile's view
    if ((use & USE_VALIDITY) && c.validity != nullptr) {
      src[0] = c.validity + (row0 >> 3);
      if (sl.validity != kNotStaged) { nb[0] = bits_bytes; so[0] = sl.validity; t.validity = stage + sl.validity - (row0 >> 3); }
      else pf[0] = bits_bytes;
    }
    if (c.type == T_UTF8) {
      if (use & USE_OFFSETS) {
        src[1] = (const uint8_t*)(c.offsets + row0);
        const uint32_t bytes = ((tile_rows + 1u) * 4u + 15u) & ~15u;
        if (sl.offsets != kNotStaged) { nb[1] = bytes; so[1] = sl.offsets; t.offsets = (const int32_t*)(stage + sl.offsets) - row0; }
        else pf[1] = bytes;
      }
      if (use & USE_VALUES) { utf8_values = true; cap = sl.values != kNotStaged ? sl.values_cap : 0u; vso = sl.values; }
    } else if (use & USE_VALUES) {
      const uint32_t w = c.width;
      const uint32_t bytes = w ? (tile_rows * w + 15u) & ~15u : bits_bytes;
      const int64_t adv = w ? row0 * (int64_t)w : (row0 >> 3);
      src[2] = (const uint8_t*)c.values + adv;
      if (sl.values != kNotStaged) { nb[2] = bytes; so[2] = sl.values; t.values = stage + sl.values - adv; }
      else pf[2] = bytes;
    }
    cols[lane] = t;
  }
  // Utf8 bounds: in flight while the fixed-size copies are issued
  uint32_t o0 = 0, o1 = 0;
  if (utf8_values) { o0 = (uint32_t)c.offsets[row0]; o1 = (uint32_t)c.offsets[row0 + tile_rows]; }
  const uint32_t tx = __reduce_add_sync(FULL, nb[0] + nb[1] + nb[2]);
  __syncwarp();   // the column table (but for the Utf8 value pointers) is complete before the arrival that publishes it
  if (lane == 0) mbar_arrive_expect_tx(full, tx);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 3; i++) {
    if (nb[i]) tma_load(ss + so[i], src[i], nb[i], full);
    if (pf[i]) tma_prefetch_l2(src[i], pf[i]);
  }
  TileValues tv;
  tv.values = utf8_values ? (const uint8_t*)c.values : nullptr;
  tv.o0 = o0; tv.o1 = o1; tv.cap = cap; tv.vso = vso;
  return tv;
}
// Step 2 (the same warp; the rest of the CTA need not wait for it): the Utf8 value bytes, whose bounds were a global
// round trip away.  The value pointers it puts into the column table are published by the arrival on `full_values`.
__device__ __forceinline__ void load_tile_values(const TileValues& tv, ColumnDesc* cols, uint8_t* stage, uint32_t full_values, int lane) {
  const uint32_t ss = smem_u32(stage);
  uint32_t vbytes = 0;
  const uint8_t* vsrc = nullptr;
  const uint32_t vso = tv.vso;
  if (tv.values != nullptr) {
    const uint32_t lo = tv.o0 & ~15u, len = (tv.o1 - lo + 15u) & ~15u;
    vsrc = tv.values + lo;
    if (tv.cap != 0 && len <= tv.cap) { vbytes = len; cols[lane].values = stage + vso - lo; }
    else if (len) tma_prefetch_l2(vsrc, len);
  }
  // The Utf8 value bytes (their bounds were a global load away) complete a second barrier: the predicate does not
  // wait for them, only the stores do (and a predicate that compares strings).
  const uint32_t tx2 = __reduce_add_sync(FULL, vbytes);
  __syncwarp();
  if (lane == 0) mbar_arrive_expect_tx(full_values, tx2);
  __syncwarp();
  if (vbytes) tma_load(ss + vso, vsrc, vbytes, full_values);
}
// ------------------------------------------------------------------------------------------
// gather: writing the selected rows
// ------------------------------------------------------------------------------------------
// What a lane knows about its rows of the current slice.
struct LaneCtx {
  int64_t row_base;       // absolute row of the lane's first row
  uint32_t row32;         // the same in 32 bits (Arrow arrays hold < 2^31 rows): one IMAD.WIDE per address
  uint32_t inrange;       // rows that exist (tail tile), 4 bits
  uint32_t sel;           // selected rows, 4 bits
  uint32_t rank;          // slice-local rank of the lane's first selected row
  uint32_t count;         // selected rows of the slice
  uint32_t obase;         // output row of the slice's first selected row (< 2^31, like every position below)
  const uint64_t* prefix; // the tile's slice prefixes [quantity][kTileSlices], or nullptr without a predicate
  int lane, slice;
  int wid;                // compute warp index in the CTA (its private bit stage / long-string tables)
};

// Drops the selected bits of the lane's rows into the warp's bit stage (zero-initialised): the bit of
// output row obase + r sits at stage bit (obase & 31) + r.
__device__ __forceinline__ void put_bits(uint32_t bits, const LaneCtx& L, uint32_t* sb, const uint8_t* pext4) {
  const uint32_t s4 = L.sel, b4 = bits & 0xFu & s4;
  if (b4) {
    const uint32_t c = pext4[(s4 << 4) | b4];
    const uint32_t p = (L.obase & 31u) + L.rank, sh = p & 31u;
    atomicOr(&sb[p >> 5], c << sh);
    if (sh > 28u && (c >> (32u - sh))) atomicOr(&sb[(p >> 5) + 1], c >> (32u - sh));
  }
}

__device__ __forceinline__ uint32_t load_bits4(const uint8_t* __restrict__ bits, uint32_t r, uint32_t need) {
  if (bits == nullptr) return FULL;
  if (!need) return 0;
  return ((uint32_t)bits[r >> 3] >> (uint32_t)(r & 4)) & 0xFu;
}

// Stores the selected elements of one quad at consecutive output positions.
template <typename E>
__device__ __forceinline__ void store_sel(E* base, uint32_t o, uint32_t s4, const E& e0, const E& e1, const E& e2, const E& e3) {
  const uint32_t i1 = o + (s4 & 1u), i2 = i1 + ((s4 >> 1) & 1u), i3 = i2 + ((s4 >> 2) & 1u);
  if (s4 & 1u) base[o] = e0;
  if (s4 & 2u) base[i1] = e1;
  if (s4 & 4u) base[i2] = e2;
  if (s4 & 8u) base[i3] = e3;
}

// 16 / 4 bytes from an arbitrarily aligned address (buffers are padded, so the aligned words around
// it are always readable).
__device__ __forceinline__ uint4 load16_unaligned(const uint8_t* p) {
  const uintptr_t a = (uintptr_t)p;
  if ((a & 15u) == 0) return *(const uint4*)p;
  const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(a & 3u) * 8u;
  const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
  if (sh == 0) return make_uint4(w0, w1, w2, w3);
  const uint32_t w4 = w[4];
  return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
                    __funnelshift_r(w3, w4, sh));
}
__device__ __forceinline__ uint32_t load4_unaligned(const uint8_t* p) {
  const uintptr_t a = (uintptr_t)p;
  const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(a & 3u) * 8u;
  const uint32_t w0 = w[0];
  if (sh == 0) return w0;
  return __funnelshift_r(w0, w[1], sh);
}

// One short value: by one thread, in the widest unit source, destination and length allow.  Only the common
// case (one aligned 8-byte unit) is inlined at the call sites: eight inlined copies of the loops below were a
// quarter of the kernel's code.
__device__ __noinline__ void copy_value_slow(uint8_t* dst, const uint8_t* src, uint32_t n) {
  const uint32_t a = (uint32_t)(uintptr_t)dst | (uint32_t)(uintptr_t)src | n;
  if ((a & 7u) == 0) {
#pragma unroll 1
    for (uint32_t i = 0; i < n; i += 8) *(uint2*)(dst + i) = *(const uint2*)(src + i);
  } else if ((a & 3u) == 0) {
#pragma unroll 1
    for (uint32_t i = 0; i < n; i += 4) *(uint32_t*)(dst + i) = *(const uint32_t*)(src + i);
  } else {
#pragma unroll 1
    for (uint32_t i = 0; i < n; i++) dst[i] = src[i];
  }
}
__device__ __forceinline__ void copy_value(uint8_t* dst, const uint8_t* src, uint32_t n) {
  const uint32_t a = (uint32_t)(uintptr_t)dst | (uint32_t)(uintptr_t)src;
  if (n == 8 && (a & 7u) == 0) *(uint2*)dst = *(const uint2*)src;
  else copy_value_slow(dst, src, n);
}

// Long strings: the warp's output byte range is produced chunk-centric -- each lane builds aligned 16-byte output
// chunks (s_oo: warp-local output byte offsets of the slice's selected rows, s_src: their source byte offsets).
// A chunk is assembled from PIECES, one per source row it touches: for each piece the 16 source bytes that would line
// up with the whole chunk are fetched with one unaligned 16-byte read and merged under a byte mask -- a chunk inside
// one row is one read and one store, a chunk that straddles two rows is two reads and one store (the first version
// built straddling chunks byte by byte: 42 % of the C4 gather kernel's stall samples).  The row of a lane's next
// chunk is found by walking on from the previous one when rows are long (a lane's chunks are 512 bytes apart), by
// binary search otherwise.
// kLowBytes16[n]: the n low bytes of a 16-byte chunk set (n in [0, 16]).  A table in constant memory: computing the
// masks with shifts was 29 % of the C4 gather kernel's instructions (profiles/r2g_c4_gather_ncu_summary.txt).
__constant__ uint4 kLowBytes16[17] = {
  {0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u},
  {0x000000FFu, 0x00000000u, 0x00000000u, 0x00000000u},
  {0x0000FFFFu, 0x00000000u, 0x00000000u, 0x00000000u},
  {0x00FFFFFFu, 0x00000000u, 0x00000000u, 0x00000000u},
  {0xFFFFFFFFu, 0x00000000u, 0x00000000u, 0x00000000u},
  {0xFFFFFFFFu, 0x000000FFu, 0x00000000u, 0x00000000u},
  {0xFFFFFFFFu, 0x0000FFFFu, 0x00000000u, 0x00000000u},
  {0xFFFFFFFFu, 0x00FFFFFFu, 0x00000000u, 0x00000000u},
  {0xFFFFFFFFu, 0xFFFFFFFFu, 0x00000000u, 0x00000000u},
  {0xFFFFFFFFu, 0xFFFFFFFFu, 0x000000FFu, 0x00000000u},
  {0xFFFFFFFFu, 0xFFFFFFFFu, 0x0000FFFFu, 0x00000000u},
  {0xFFFFFFFFu, 0xFFFFFFFFu, 0x00FFFFFFu, 0x00000000u},
  {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x00000000u},
  {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x000000FFu},
  {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x0000FFFFu},
  {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x00FFFFFFu},
  {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}};
__device__ __forceinline__ uint4 byte_range_mask(uint32_t a, uint32_t z) {   // bytes [a, z) of a 16-byte chunk, a < z <= 16
  const uint4 lo = kLowBytes16[a], hi = kLowBytes16[z];
  return make_uint4(hi.x & ~lo.x, hi.y & ~lo.y, hi.z & ~lo.z, hi.w & ~lo.w);
}
// the first version's assembly of one chunk, byte / word at a time: kept for the one case the piece reads cannot
// serve (a piece whose lined-up source address would lie before the start of the value buffer)
__device__ __noinline__ uint4 assemble_chunk_bytes(const uint8_t* __restrict__ sv, uint32_t mis, uint32_t lo, uint32_t s, uint32_t t,
                                                   uint32_t r, const uint32_t* s_oo, const int32_t* s_src) {
  uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
#pragma unroll 1
  for (uint32_t b = s; b < t;) {
    const uint32_t xb = b - mis;
    while (xb >= s_oo[r + 1]) r++;
    const uint8_t* sp = sv + s_src[r] + (xb - s_oo[r]);
    uint32_t piece, step;
    if (((b & 3u) == 0) && b + 4 <= t && xb + 4 <= s_oo[r + 1]) { piece = load4_unaligned(sp); step = 4; }
    else { piece = (uint32_t)*sp << (8u * (b & 3u)); step = 1; }
    const uint32_t wi = (b - lo) >> 2;
    if (wi == 0) w0 |= piece; else if (wi == 1) w1 |= piece; else if (wi == 2) w2 |= piece; else w3 |= piece;
    b += step;
  }
  return make_uint4(w0, w1, w2, w3);
}
// (kChunksInFlight chunks per lane and iteration: the first piece's reads of all of them are issued before any is
// looked at.  Measured on C4 / C4H (100-byte strings, 100 % / 50 % selected): 1 chunk 169.2 / 123.0 us per 2M-row batch,
// 2 chunks 164.2 / 102.3 us, 4 chunks 219.4 / 134.8 us)
constexpr int kChunksInFlight = 2;
__device__ __noinline__ void copy_long_strings(const uint8_t* __restrict__ sv, uint8_t* gal, uint32_t mis, uint32_t nbytes, uint32_t nrows,
                                               const uint32_t* s_oo, const int32_t* s_src, int lane) {
  const uint32_t end = mis + nbytes;
  const uint32_t nchunks = (end + 15u) >> 4;
  const bool walk = nbytes >= 64u * nrows;   // (uniform) long rows: the next chunk's row is a few rows on
  uint32_t r = 0;
  bool first = true;
#pragma unroll 1
  for (uint32_t ch0 = lane; ch0 < nchunks; ch0 += 32 * kChunksInFlight) {
    uint32_t rs[kChunksInFlight], w[kChunksInFlight][5], shv[kChunksInFlight];
    uint32_t live = 0, fb = 0;   // bit u: chunk u produces bytes / needs the byte-wise fallback
    // ---- 1. the row of each chunk's first byte, and the reads of its first piece ----
#pragma unroll
    for (int u = 0; u < kChunksInFlight; u++) {
      const uint32_t ch = ch0 + 32u * u;
      const uint32_t lo = ch << 4, hi = lo + 16;
      const uint32_t s = lo > mis ? lo : mis, t = hi < end ? hi : end;
      rs[u] = r;
      shv[u] = 0;
#pragma unroll
      for (int i = 0; i < 5; i++) w[u][i] = 0;
      if (ch >= nchunks || s >= t) continue;
      const uint32_t x = s - mis;  // warp-local output byte index of the first byte produced
      if (walk && !first) {
        while (s_oo[r + 1] <= x) r++;
      } else {
        uint32_t lo_r = 0, hi_r = nrows;  // first r in (0, nrows] with s_oo[r] > x
        while (lo_r < hi_r) {
          const uint32_t mid = (lo_r + hi_r) >> 1;
          if (s_oo[mid] > x) hi_r = mid; else lo_r = mid + 1;
        }
        r = lo_r - 1;  // row holding byte x (empty strings are skipped by the search)
      }
      first = false;
      rs[u] = r;
      live |= 1u << u;
      // the source address that lines up with output byte `lo`
      const int64_t vsrc = (int64_t)s_src[r] + (int64_t)lo - (int64_t)(s_oo[r] + mis);
      if (vsrc < 0) { fb |= 1u << u; continue; }
      const uintptr_t a = (uintptr_t)(sv + vsrc);
      const uint32_t* q = (const uint32_t*)(a & ~(uintptr_t)3);
      shv[u] = (uint32_t)(a & 3u) * 8u;
#pragma unroll
      for (int i = 0; i < 5; i++) w[u][i] = q[i];   // (buffers are padded: the fifth word is always readable)
    }
    // ---- 2. merge the pieces, store ----
#pragma unroll
    for (int u = 0; u < kChunksInFlight; u++) {
      if (!((live >> u) & 1u)) continue;
      const uint32_t ch = ch0 + 32u * u;
      const uint32_t lo = ch << 4, hi = lo + 16;
      const uint32_t s = lo > mis ? lo : mis, t = hi < end ? hi : end;
      uint4 acc;
      if ((fb >> u) & 1u) {
        acc = assemble_chunk_bytes(sv, mis, lo, s, t, rs[u], s_oo, s_src);
      } else {
        uint32_t rr = rs[u];
        const uint32_t row_hi = s_oo[rr + 1] + mis;
        uint32_t e = t < row_hi ? t : row_hi;
        acc = make_uint4(__funnelshift_r(w[u][0], w[u][1], shv[u]), __funnelshift_r(w[u][1], w[u][2], shv[u]),
                         __funnelshift_r(w[u][2], w[u][3], shv[u]), __funnelshift_r(w[u][3], w[u][4], shv[u]));
        if (!(s == lo && e == hi)) {
          const uint4 m = byte_range_mask(s - lo, e - lo);
          acc.x &= m.x; acc.y &= m.y; acc.z &= m.z; acc.w &= m.w;
          uint32_t b = e;
#pragma unroll 1
          while (b < t) {   // further pieces: the rows that follow inside this chunk
            do { rr++; } while (s_oo[rr + 1] + mis <= b);   // (skips empty strings)
            const uint32_t row_lo2 = s_oo[rr] + mis, row_hi2 = s_oo[rr + 1] + mis;
            e = t < row_hi2 ? t : row_hi2;
            const int64_t vsrc = (int64_t)s_src[rr] + (int64_t)lo - (int64_t)row_lo2;
            if (vsrc < 0) { acc = assemble_chunk_bytes(sv, mis, lo, s, t, rs[u], s_oo, s_src); break; }
            const uint4 v = load16_unaligned(sv + vsrc);
            const uint4 m2 = byte_range_mask(b - lo, e - lo);
            acc.x |= v.x & m2.x; acc.y |= v.y & m2.y; acc.z |= v.z & m2.z; acc.w |= v.w & m2.w;
            b = e;
          }
        }
      }
      if ((t - s) == 16u) {
        *(uint4*)(gal + lo) = acc;
      } else {
        for (uint32_t bb = s; bb < t; bb++) {
          const uint32_t wi = (bb - lo) >> 2;
          const uint32_t word = wi == 0 ? acc.x : wi == 1 ? acc.y : wi == 2 ? acc.z : acc.w;
          gal[bb] = (uint8_t)(word >> (8u * (bb & 3u)));
        }
      }
    }
  }
}

// An output column is handled in two steps so that, with the program known at compile time
// (CHDB_JIT), the loads of ALL pass-through columns are issued before the first store: the
// shared-memory latencies overlap instead of adding up column by column.
struct OutRegs {
  uint4 x, y;        // the lane's 4 values (widths 1..8), or x = 4 Utf8 offsets
  uint32_t z;        // 5th Utf8 offset, or Boolean value bits
  uint32_t vbits;    // validity bits of the lane's rows
};

// `meta` packs the eight small OutDesc fields; under CHDB_JIT it is a compile-time constant.
__device__ __forceinline__ void load_output(const KernelParams& P, const ColumnDesc* cols, const int k, const uint64_t meta,
                                            const LaneCtx& L, OutRegs& R) {
  const uint32_t o_kind = (uint32_t)meta & 0xFFu, o_type = (uint32_t)(meta >> 8) & 0xFFu, o_width = (uint32_t)(meta >> 16) & 0xFFu;
  const uint32_t o_slot = (uint32_t)(meta >> 24) & 0xFFu;
  // Only what the column's type needs is loaded, and nothing is cleared first: fields the stores never
  // look at stay undefined (clearing ten registers per output cost more instructions than the loads).
  R.vbits = FULL;
  if (o_kind != OUT_PASS) return;
  const ColumnDesc& c = cols[o_slot];
  const uint32_t sel = L.sel;
  const uint32_t r = L.row32;
  if (CHDB_OUT_HAS_VALIDITY(P, k)) R.vbits = load_bits4(c.validity, r, sel);
  if (!L.inrange) return;   // (a tail tile's missing rows: nothing to read, nothing will be stored)
  const uint8_t* src = (const uint8_t*)c.values;
  if (o_type == T_BOOL) {
    R.z = load_bits4(src, r, sel);
  } else if (o_type == T_UTF8) {
    const int4 a = *(const int4*)(c.offsets + (size_t)r);
    R.x = make_uint4((uint32_t)a.x, (uint32_t)a.y, (uint32_t)a.z, (uint32_t)a.w);
    R.z = (uint32_t)c.offsets[(size_t)r + 4];
  } else if (o_width == 4) {
    R.x = *(const uint4*)(src + (size_t)r * 4);
  } else if (o_width == 8) {
    R.x = *(const uint4*)(src + (size_t)r * 8);
    R.y = *(const uint4*)(src + (size_t)r * 8 + 16);
  } else if (o_width == 2) {
    const uint2 t = *(const uint2*)(src + (size_t)r * 2);
    R.x.x = t.x; R.x.y = t.y;
  } else if (o_width == 1) {
    R.x.x = *(const uint32_t*)(src + r);
  }
}

// Writes output column k for this lane's rows.  BEGIN/END: the expression's instruction range when known at
// compile time.  kb: running index of the bit-packed outputs (Boolean values, validity bitmaps) in the bit stage.
template <typename V, int BEGIN = -1, int END = -1>
__device__ __forceinline__ void store_output(const KernelParams& P, const ColumnDesc* cols, const int k, const uint64_t meta,
                                             const LaneCtx& L, const OutRegs& R, const TileShared& sh, uint32_t* bitstage,
                                             uint32_t* ltab, int& kb) {
  const uint32_t o_kind = (uint32_t)meta & 0xFFu, o_type = (uint32_t)(meta >> 8) & 0xFFu, o_width = (uint32_t)(meta >> 16) & 0xFFu;
  const uint32_t o_slot = (uint32_t)(meta >> 24) & 0xFFu, o_begin = (uint32_t)(meta >> 32) & 0xFFu, o_end = (uint32_t)(meta >> 40) & 0xFFu;
  const uint32_t o_utf8 = (uint32_t)(meta >> 48) & 0xFFu;
  uint8_t* const o_values = (uint8_t*)P.out[k].values;
  const bool o_has_validity = CHDB_OUT_HAS_VALIDITY(P, k);
  const uint32_t sel = L.sel;
  const int lane = L.lane;
  const uint32_t o = L.obase + L.rank;   // output row of the lane's first selected row
  uint32_t vbits = R.vbits;  // validity of this output for the lane's rows
  if (o_kind == OUT_EXPR) {
    uint32_t accm = 0;
    vbits = 0;
    if (sel) {
      const int64_t qb[1] = {L.row_base};
      V a4[4];
      // `sel` as the active mask: checked arithmetic only sees rows that survived the filter
      run_program<V, 1, BEGIN, END>(P, cols, (int)o_begin, (int)o_end, qb, L.inrange, sel, sh.pool, a4, accm, vbits);
      if (o_type == T_BOOL) {}
      else if (o_width == 4) store_sel<uint32_t>((uint32_t*)o_values, o, sel, (uint32_t)a4[0], (uint32_t)a4[1], (uint32_t)a4[2], (uint32_t)a4[3]);
      else if (o_width == 8) store_sel<uint64_t>((uint64_t*)o_values, o, sel, (uint64_t)a4[0], (uint64_t)a4[1], (uint64_t)a4[2], (uint64_t)a4[3]);
      else if (o_width == 2) store_sel<uint16_t>((uint16_t*)o_values, o, sel, (uint16_t)a4[0], (uint16_t)a4[1], (uint16_t)a4[2], (uint16_t)a4[3]);
      else store_sel<uint8_t>((uint8_t*)o_values, o, sel, (uint8_t)a4[0], (uint8_t)a4[1], (uint8_t)a4[2], (uint8_t)a4[3]);
    }
    if (o_type == T_BOOL) { put_bits(accm, L, bitstage + kb * kBitWords, sh.pext4); kb++; }
  } else {
    const ColumnDesc& c = cols[o_slot];
    if (o_type == T_BOOL) {
      put_bits(R.z, L, bitstage + kb * kBitWords, sh.pext4);
      kb++;
    } else if (o_type == T_UTF8) {
      // offsets: running sum of the selected lengths, restarted at 0 for the output
      const uint8_t* sv = (const uint8_t*)c.values;
      const uint32_t byte_base = (uint32_t)L.prefix[(1 + o_utf8) * kTileSlices + L.slice];   // output byte offset of this slice's first value
      int32_t* const o_off = P.out[k].offsets;
      const int32_t o5[5] = {(int32_t)R.x.x, (int32_t)R.x.y, (int32_t)R.x.z, (int32_t)R.x.w, (int32_t)R.z};
      uint32_t len[4];
#pragma unroll
      for (int i = 0; i < 4; i++) len[i] = ((sel >> i) & 1u) ? (uint32_t)(o5[i + 1] - o5[i]) : 0u;
      uint32_t slice_bytes;
      uint32_t bo = warp_excl_scan(len[0] + len[1] + len[2] + len[3], lane, slice_bytes);   // slice-local output byte offset
      // long values are copied by the whole warp, chunk-centric; short ones by the lane that owns the row
      const bool chunked = CHDB_LONG_STRINGS(P) && slice_bytes > 24u * L.count;
      uint32_t* s_oo = ltab + L.wid * (2 * (kWarpRows + 4));
      int32_t* s_src = (int32_t*)(s_oo + kWarpRows + 4);
      uint32_t oi = o;
      uint32_t wr = L.rank;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        if ((sel >> i) & 1u) {
          o_off[oi] = (int32_t)(byte_base + bo);
          oi++;
          if (chunked) { s_oo[wr] = bo; s_src[wr] = o5[i]; wr++; }
          else if (len[i]) copy_value(o_values + (byte_base + bo), sv + (uint32_t)o5[i], len[i]);
          bo += len[i];
        }
      }
      if (chunked) {
        if (lane == 0) s_oo[L.count] = slice_bytes;
        __syncwarp();
        const uint32_t mis = byte_base & 15u;
        copy_long_strings(sv, o_values + (byte_base - mis), mis, slice_bytes, L.count, s_oo, s_src, lane);
        __syncwarp();
      }
    } else if (sel) {
      uint8_t* const vb = o_values;
      const uint32_t vo = o;
      if (o_width == 4) {
        store_sel<uint32_t>((uint32_t*)vb, vo, sel, R.x.x, R.x.y, R.x.z, R.x.w);
      } else if (o_width == 8) {
        store_sel<uint2>((uint2*)vb, vo, sel, make_uint2(R.x.x, R.x.y), make_uint2(R.x.z, R.x.w), make_uint2(R.y.x, R.y.y),
                         make_uint2(R.y.z, R.y.w));
      } else if (o_width == 16) {
        const uint4* s16 = (const uint4*)c.values + L.row32;
        uint4* d = (uint4*)o_values + o;
#pragma unroll
        for (int i = 0; i < 4; i++)
          if ((sel >> i) & 1u) { *d = s16[i]; d++; }
      } else if (o_width == 2) {
        store_sel<uint16_t>((uint16_t*)vb, vo, sel, (uint16_t)R.x.x, (uint16_t)(R.x.x >> 16), (uint16_t)R.x.y, (uint16_t)(R.x.y >> 16));
      } else {
        const uint32_t x = R.x.x;
        store_sel<uint8_t>(vb, vo, sel, (uint8_t)x, (uint8_t)(x >> 8), (uint8_t)(x >> 16), (uint8_t)(x >> 24));
      }
    }
  }
  if (o_has_validity) {
    // nulls are the rare case: the stage collects the NULL bits (most lanes have nothing to add)
    const uint32_t nulls = sel & ~vbits & 0xFu;
    if (nulls) put_bits(nulls, L, bitstage + kb * kBitWords, sh.pext4);
    if (!CHDB_EARLY_COUNTS(P)) {   // (two-launch form with pass-through outputs only: the select kernel has counted them)
      const uint32_t slice_nulls = __reduce_add_sync(FULL, (uint32_t)__popc(nulls));
      if (lane == 0 && slice_nulls) atomicAdd(&sh.nulls[k], slice_nulls);
    }
    kb++;
  }
}

#ifdef CHDB_JIT
template <int K, int N>
__device__ __forceinline__ void load_outputs_range(const KernelParams& P, const ColumnDesc* cols, const LaneCtx& L, OutRegs (&R)[N > 0 ? N : 1]) {
  if constexpr (K < N) {
    load_output(P, cols, K, chdb_jit::kOutMeta[K], L, R[K]);
    load_outputs_range<K + 1, N>(P, cols, L, R);
  }
}
template <typename V, int K, int N>
__device__ __forceinline__ void store_outputs_range(const KernelParams& P, const ColumnDesc* cols, const LaneCtx& L, const OutRegs (&R)[N > 0 ? N : 1],
                                                    const TileShared& sh, uint32_t* bitstage, uint32_t* ltab, int& kb) {
  if constexpr (K < N) {
    constexpr uint64_t meta = chdb_jit::kOutMeta[K];
    store_output<V, (int)((meta >> 32) & 0xFFu), (int)((meta >> 40) & 0xFFu)>(P, cols, K, meta, L, R[K], sh, bitstage, ltab, kb);
    store_outputs_range<V, K + 1, N>(P, cols, L, R, sh, bitstage, ltab, kb);
  }
}
#endif

// The warp's bit stage -> global bitmaps.  Stage bit (obase & 31) + r belongs to output row obase + r;
// whole words are stored, the (at most two) words shared with neighbouring slices are merged with
// atomicOr (the bitmaps are zero-initialised).  The stage is left zeroed for the warp's next slice.
__device__ __forceinline__ void flush_bits(const KernelParams& P, uint32_t* bitstage, uint32_t obase, uint32_t count, int lane) {
  const uint32_t o = obase & 31u, end = o + count;
  const uint32_t nwords = (end + 31u) >> 5;
  const uint32_t g0 = obase >> 5;
  int kb = 0;
  CHDB_STATIC_UNROLL
  for (int k = 0; k < CHDB_N_OUT; k++) {
    const uint64_t meta = CHDB_OUT_META(P, k);
    const bool is_bool = ((uint32_t)(meta >> 8) & 0xFFu) == T_BOOL;
    uint8_t* const validity = P.out[k].validity;
#pragma unroll
    for (int which = 0; which < 2; which++) {   // 0: Boolean values, 1: validity
      if (which == 0 ? !is_bool : !CHDB_OUT_HAS_VALIDITY(P, k)) continue;
      uint32_t* sb = bitstage + kb * kBitWords;
      uint32_t* g = (uint32_t*)(which == 0 ? (uint8_t*)P.out[k].values : validity);
      if (lane < kBitWords) {
        const uint32_t w = (uint32_t)lane;
        uint32_t word = sb[w];
        sb[w] = 0;
        if (w < nwords) {
          const uint32_t lo = w == 0 ? o : 0u, hi = 32 * w + 32 > end ? end - 32 * w : 32u;
          const uint32_t mask = (hi >= 32 ? FULL : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
          if (which == 1) word = ~word;
          word &= mask;
          if (mask == FULL) g[g0 + w] = word;
          else if (word) atomicOr(&g[g0 + w], word);
        }
      }
      kb++;
    }
  }
}


// ------------------------------------------------------------------------------------------
// The stream kernel: one CTA per tile of kTileRows rows.
//   1. warp 0 brings the tile's slice of every column the program touches into shared memory (TMA);
//   2. every warp evaluates the predicate on its slices: selection bits and ranks stay in registers, the
//      slice counts (selected rows; selected value bytes per Utf8 output) go to shared memory;
//   3. warp q turns quantity q's slice counts into batch-wide exclusive prefixes: it publishes the tile's
//      aggregate and walks back over its predecessors' descriptors (decoupled look-back) -- the other
//      resident CTAs of the SM keep the memory system busy meanwhile;
//   4. every warp writes the selected rows of its slices to their final positions.
// Without a predicate steps 2-3 fall away (every row is kept, output row = input row).
// MANY: one launch over several batches of one schema: the CTA first fetches its batch's record.
// ------------------------------------------------------------------------------------------
// CHDB_TRACE: per tile, clock64() at: 0 CTA ready to load, 1 tile landed, 2 predicate done, 3 look-back done (warp 0),
// 4 Utf8 values landed, 5 rows stored, 6 CTA done; slot 7: the SM the CTA ran on.
__device__ __forceinline__ void trace_event(const KernelParams& P, int ev) {
  if (P.trace != nullptr && blockIdx.x < 8192) P.trace[(size_t)blockIdx.x * 8 + ev] = (uint64_t)clock64();
}

// MODE: kFused -- everything in one launch (steps 1-4 above);
//       kSelect / kGather -- the same work as two launches without any CTA waiting on another while it holds a tile:
//         select: no staging (every predicate byte is read once, by one thread: plain 128-bit global loads), steps 2-3:
//                 selection bits -> P.b.selbits, tile aggregates -> look-back -> every descriptor ends up as the tile's
//                 inclusive prefix; the CTAs that wait in the look-back hold registers only, and nothing follows the wait;
//         gather: steps 1 and 4 with the selection bits read back (128 B per tile) and the tile's exclusive prefix taken
//                 from its predecessor's descriptor.  Tiles are visited last-to-first: what select read last is still
//                 in L2 when gather starts.
enum StreamMode : int { kFused = 0, kSelect = 1, kGather = 2 };

template <typename V, bool MANY, int MODE>
__device__ __forceinline__ void stream_body(const KernelParams& PP, const TilePlan& TP) {
  static_assert(MODE == kFused || !MANY, "many-batch launches run the fused kernel");
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t s_full, s_full_values;
  __shared__ uint32_t s_nulls[kMaxOutCols];
  __shared__ uint32_t s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int64_t tile = blockIdx.x;
  if (MODE == kGather) tile = (int64_t)gridDim.x - 1 - (int64_t)blockIdx.x;
  if (MANY) {
    // this tile's parameter block: the program part from the launch parameters, the batch part from its record
    KernelParams* mine = (KernelParams*)(smem + TP.params_off);
    const int32_t k = PP.many_tile_batch[blockIdx.x];   // (one entry per tile: the batch it belongs to)
    const uint8_t* rec = PP.many + (size_t)k * (size_t)PP.many_stride;
    uint32_t* d32 = (uint32_t*)mine;
    const uint32_t* s32 = (const uint32_t*)&PP;
    for (int i = tid; i < (int)(sizeof(KernelParams) / 4); i += kThreads) d32[i] = s32[i];
    __syncthreads();
    const int n_in = PP.n_in, n_out = PP.n_out;
    const uint32_t* r32 = (const uint32_t*)rec;
    uint32_t* hb = (uint32_t*)&mine->b;
    for (int i = tid; i < (int)(sizeof(BatchHeader) / 4); i += kThreads) hb[i] = r32[i];
    uint32_t* hi = (uint32_t*)mine->in;
    for (int i = tid; i < n_in * 8; i += kThreads) hi[i] = r32[sizeof(BatchHeader) / 4 + i];
    uint32_t* ho = (uint32_t*)mine->out;
    for (int i = tid; i < n_out * 8; i += kThreads) ho[i] = r32[sizeof(BatchHeader) / 4 + n_in * 8 + i];
    __syncthreads();
    tile -= mine->b.first_tile;
  }
  const KernelParams& P = MANY ? *(const KernelParams*)(smem + TP.params_off) : PP;
  const bool has_pred = CHDB_PRED_END > CHDB_PRED_BEGIN;
  const int nq = 1 + CHDB_N_UTF8;
  uint8_t* const stage = smem;
  // (select reads the columns where they are: the descriptors come straight from the kernel parameters)
  ColumnDesc* const cols_s = (ColumnDesc*)(smem + TP.cols_off);
  const ColumnDesc* const cols = MODE == kSelect ? P.in : cols_s;
  uint32_t* const s_cnt = (uint32_t*)(smem + TP.cnt_off);
  uint64_t* const s_pre = (uint64_t*)(smem + TP.pre_off);
  uint64_t* const s_tot = (uint64_t*)(smem + TP.tot_off);
  uint32_t* const bitstages = (uint32_t*)(smem + TP.bits_off);
  uint32_t* const ltab = (uint32_t*)(smem + TP.ltab_off);
  uint8_t* const pext4 = smem + TP.pext_off;
  const uint32_t full = smem_u32(&s_full), full_values = smem_u32(&s_full_values);
  static_assert(kThreads >= 64, "one thread per word of the bit-compaction table");
  const int64_t row0 = tile * kTileRows;
  const int32_t tile_rows = (int32_t)(row0 + kTileRows < P.b.num_rows ? kTileRows : P.b.num_rows - row0);
  const bool last_tile = tile == (int64_t)P.b.num_tiles - 1;
  const uint64_t t_entry = (MODE == kGather && P.trace != nullptr) ? (uint64_t)clock64() : 0;
  // the bit-compaction table: requested now, stored to shared memory once its load has had time to complete
  // (nothing on the way to the tile's loads waits for a global round trip)
  uint32_t pext_word = 0;
  const bool pext_mine = MODE != kSelect && P.n_bits > 0 && tid < 64;
  if (pext_mine) pext_word = kPext4Words[tid];
  if (MODE != kSelect && tid == 0) {
    mbar_init(full, kWarps);          // (one arrival per warp: load_tile)
    mbar_init(full_values, kWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  // gather: the tile's loads are the first thing the CTA does.  They do not depend on the select kernel (the launch
  // chain starts with a kernel that is NOT a programmatic dependent, so everything older on the stream -- whatever
  // produced the inputs -- has completed).
  TileValues tv;
  if (MODE == kGather) {
    __syncthreads();   // (the barriers are initialised)
    tv = load_tile(P, TP, cols_s, stage, full, row0, (uint32_t)tile_rows, lane, warp);
  }
  if (MODE != kSelect && P.n_bits > 0) {
    if (MODE != kGather && pext_mine) ((uint32_t*)pext4)[tid] = pext_word;
#pragma unroll 1
    for (int i = tid; i < kWarps * P.n_bits * kBitWords; i += kThreads) bitstages[i] = 0;
  }
  if (tid < kMaxOutCols) s_nulls[tid] = 0;
  __syncthreads();
  grid_launch_dependents();
  grid_dependency_wait();     // the zeroed workspace; inputs an earlier kernel on this stream may still be writing
  if (tid == 0) {
    trace_event(P, 0);
    if (P.trace != nullptr && blockIdx.x < 8192) {
      uint32_t smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
      P.trace[(size_t)blockIdx.x * 8 + 7] = MODE == kGather ? t_entry : (uint64_t)smid;
    }
  }
  if (MODE == kFused) tv = load_tile(P, TP, cols_s, stage, full, row0, (uint32_t)tile_rows, lane, warp);
  // gather: the lane's selection bits of each of its slices (written by the select kernel; served by L2) -- requested
  // here, looked at after the prefix sums below have been requested too (one round trip, not two)
  uint32_t selq[kSpw];
  if (MODE == kGather && has_pred) {
#pragma unroll
    for (int j = 0; j < kSpw; j++) {
      const int64_t r = row0 + (warp * kSpw + j) * kWarpRows + lane * 4;
      selq[j] = (uint32_t)__ldcg((const uint8_t*)P.b.selbits + (r >> 3));
    }
  }
  // (the Utf8 bounds have had a barrier's worth of time to arrive; only this warp waits for them)
  if (MODE != kSelect) load_tile_values(tv, cols_s, stage, full_values, lane);
  if (MODE == kGather && CHDB_EARLY_COUNTS(P) && blockIdx.x == 0 && warp == kWarps - 1) {
    // The select kernel has completed: batch totals = the sums of its group totals; the closing Utf8 offset
    // (offsets[total_rows] = total_bytes, also covers an empty result); counts, NULL counts and the error word go to
    // the pinned host mirror.  One warp of one CTA, while its tile lands.
    const size_t nt = (size_t)P.b.num_tiles, ng = (nt + kGroupTiles - 1) / kGroupTiles;
    uint64_t rows = 0;
#pragma unroll 1
    for (int q = 0; q < nq; q++) {
      const uint64_t* gt = P.b.desc + (size_t)nq * nt + (size_t)q * ng;
      uint64_t part = 0;
#pragma unroll 1
      for (size_t i = lane; i < ng; i += 32) part += load_descriptor(gt + i);
      part = warp_sum64(part);   // (every lane holds the total)
      if (q == 0) rows = part;
      if (lane == 0) { P.b.counts[q] = part; P.b.host_counts[q] = part; }
      if (q > 0 && lane < CHDB_N_OUT) {
        const OutDesc& o = P.out[lane];
        if (o.utf8_index == (uint8_t)(q - 1)) o.offsets[rows] = (int32_t)part;
      }
    }
#pragma unroll 1
    for (int i = nq + lane; i <= P.n_counts; i += 32) P.b.host_counts[i] = load_descriptor((const uint64_t*)P.b.counts + i);
    __syncwarp();
  }
  if (MODE == kGather && has_pred) {
    // the tile's exclusive prefixes: the totals of the tile groups before its own + of the tiles before it inside its
    // group (left by the select kernel); the loads are in flight while the tile lands
    for (int q = warp; q < nq; q += kWarps) {
      const size_t nt = (size_t)P.b.num_tiles, ng = (nt + kGroupTiles - 1) / kGroupTiles;
      const int64_t g = tile / kGroupTiles;
      const uint64_t* tt = P.b.desc + (size_t)q * nt;
      const uint64_t* gt = P.b.desc + (size_t)nq * nt + (size_t)q * ng;
      uint64_t part = 0;
#pragma unroll 2
      for (int64_t i = lane; i < g; i += 32) part += load_descriptor(gt + i);
#pragma unroll 2
      for (int64_t i = g * kGroupTiles + lane; i < tile; i += 32) part += load_descriptor(tt + i);
      part = warp_sum64(part);
      if (lane == 0) s_tot[q] = part;
    }
  }
  if (MODE == kGather && pext_mine) ((uint32_t*)pext4)[tid] = pext_word;   // (visible after the barrier of step 3)
  if (MODE == kGather && has_pred) {
#pragma unroll
    for (int j = 0; j < kSpw; j++) {
      const int64_t r = row0 + (warp * kSpw + j) * kWarpRows + lane * 4;
      selq[j] = (selq[j] >> (uint32_t)(r & 4)) & 0xFu;
    }
  }
  if (MODE != kSelect) {
    mbar_wait(full, 0);
    if (TP.pred_reads_utf8 && MODE == kFused) mbar_wait(full_values, 0);
  }
  if (tid == 0) trace_event(P, 1);

  TileShared sh;
  sh.pool = (const uint8_t*)P.strpool;
  sh.pext4 = pext4;
  sh.nulls = s_nulls;

  // ---- 2 (select). every load of the warp's kSpw slices is issued before anything waits for one of them: the predicate
  //      runs on all of the lane's quads at once, the Utf8 offsets are requested alongside its columns ----
  if constexpr (MODE == kSelect) {
    int64_t qb[kSpw];
    uint32_t in_all = 0;
#pragma unroll
    for (int j = 0; j < kSpw; j++) {
      const int slice = warp * kSpw + j;
      qb[j] = row0 + slice * kWarpRows + lane * 4;
      const int left = tile_rows - (slice * kWarpRows + lane * 4);
      in_all |= (left >= 4 ? 0xFu : left <= 0 ? 0u : ((1u << left) - 1u)) << (4 * j);
    }
#ifdef CHDB_JIT
    int4 oa[chdb_jit::kNumUtf8 > 0 ? chdb_jit::kNumUtf8 : 1][kSpw];
    int ob[chdb_jit::kNumUtf8 > 0 ? chdb_jit::kNumUtf8 : 1][kSpw];
#pragma unroll
    for (int k = 0; k < chdb_jit::kNumOut; k++) {
      const uint64_t meta = chdb_jit::kOutMeta[k];
      const uint32_t o_utf8 = (uint32_t)(meta >> 48) & 0xFFu, o_slot = (uint32_t)(meta >> 24) & 0xFFu;
      if (o_utf8 == 0xFFu) continue;
      const int32_t* off = cols[o_slot].offsets;
#pragma unroll
      for (int j = 0; j < kSpw; j++) {
        oa[o_utf8][j] = make_int4(0, 0, 0, 0);
        ob[o_utf8][j] = 0;
        if ((in_all >> (4 * j)) & 0xFu) { oa[o_utf8][j] = *(const int4*)(off + qb[j]); ob[o_utf8][j] = off[qb[j] + 4]; }
      }
    }
#endif
    V acc[4 * kSpw];
    uint32_t accm, accv;
#ifdef CHDB_JIT
    run_program<V, kSpw, chdb_jit::kPredBegin, chdb_jit::kPredEnd>(P, cols, 0, 0, qb, in_all, in_all, sh.pool, acc, accm, accv);
#else
    run_program<V, kSpw>(P, cols, P.pred_begin, P.pred_end, qb, in_all, in_all, sh.pool, acc, accm, accv);
#endif
    const uint32_t sel_all = accm & accv & in_all;   // NULL predicate rows are dropped (arrow-select filter)
#pragma unroll
    for (int j = 0; j < kSpw; j++) {
      const int slice = warp * kSpw + j;
      const uint32_t sel4 = (sel_all >> (4 * j)) & 0xFu;
      const uint32_t wrows = __reduce_add_sync(FULL, (uint32_t)__popc(sel4));
      if (lane == 0) s_cnt[slice] = wrows;
      // selected value bytes per Utf8 output
      CHDB_STATIC_UNROLL
      for (int k = 0; k < CHDB_N_OUT; k++) {
        const uint64_t meta = CHDB_OUT_META(P, k);
        const uint32_t o_utf8 = (uint32_t)(meta >> 48) & 0xFFu;
        if (o_utf8 == 0xFFu) continue;   // uniform branch
        uint32_t bytes = 0;
#ifdef CHDB_JIT
        const int4 a = oa[o_utf8][j];
        const int a4 = ob[o_utf8][j];
#else
        int4 a = make_int4(0, 0, 0, 0);
        int a4 = 0;
        if (sel4) {
          const int32_t* off = cols[(uint32_t)(meta >> 24) & 0xFFu].offsets;
          a = *(const int4*)(off + qb[j]);
          a4 = off[qb[j] + 4];
        }
#endif
        if (sel4 & 1u) bytes += (uint32_t)(a.y - a.x);
        if (sel4 & 2u) bytes += (uint32_t)(a.z - a.y);
        if (sel4 & 4u) bytes += (uint32_t)(a.w - a.z);
        if (sel4 & 8u) bytes += (uint32_t)(a4 - a.w);
        const uint32_t wbytes = __reduce_add_sync(FULL, bytes);
        if (lane == 0) s_cnt[(1 + o_utf8) * kTileSlices + slice] = wbytes;
      }
      // the selection bits of the slice, LSB-first, one 32-bit word per 8 lanes
      uint32_t w = sel4;
      w |= __shfl_down_sync(FULL, w, 1) << 4;
      w |= __shfl_down_sync(FULL, w, 2) << 8;
      w |= __shfl_down_sync(FULL, w, 4) << 16;
      if ((lane & 7) == 0) P.b.selbits[tile * (kTileRows / 32) + slice * (kWarpRows / 32) + (lane >> 3)] = w;
    }
    // With pass-through outputs only, everything the host wants to know is known here: the NULLs among the selected
    // rows of every nullable output are counted now, and the first gather CTA publishes the batch totals before it
    // starts on its tile -- the gather kernel otherwise moves data and nothing else (no counting, no fence, no
    // last-CTA protocol at its end).
    if (CHDB_EARLY_COUNTS(P)) {
      CHDB_STATIC_UNROLL
      for (int k = 0; k < CHDB_N_OUT; k++) {
        if (!CHDB_OUT_HAS_VALIDITY(P, k)) continue;
        const uint64_t meta = CHDB_OUT_META(P, k);
        const uint8_t* vb = cols[(uint32_t)(meta >> 24) & 0xFFu].validity;
        uint32_t nulls = 0;
#pragma unroll
        for (int j = 0; j < kSpw; j++) {
          const uint32_t sel4 = (sel_all >> (4 * j)) & 0xFu;
          nulls += (uint32_t)__popc(sel4 & ~load_bits4(vb, (uint32_t)qb[j], sel4) & 0xFu);
        }
        nulls = __reduce_add_sync(FULL, nulls);
        if (lane == 0 && nulls) atomicAdd(&s_nulls[k], nulls);
      }
    }
    // tile totals: plain counts for the gather kernel's prefix sums, and the tile group's running totals
    __syncthreads();
    const size_t nt = (size_t)P.b.num_tiles, ng = (nt + kGroupTiles - 1) / kGroupTiles;
    for (int q = warp; q < nq; q += kWarps) {
      const uint32_t c = lane < kTileSlices ? s_cnt[q * kTileSlices + lane] : 0u;
      const uint32_t agg = __reduce_add_sync(FULL, c);
      if (lane == 0) {
        P.b.desc[(size_t)q * nt + (size_t)tile] = agg;
        asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(P.b.desc + (size_t)nq * nt + (size_t)q * ng + (size_t)(tile / kGroupTiles)), "l"((uint64_t)agg) : "memory");
      }
    }
    // (fire-and-forget: the gather kernel only starts once this kernel has completed, and its first CTA publishes
    //  the batch totals -- no fence, no last-CTA protocol here)
    if (CHDB_EARLY_COUNTS(P) && warp == 0 && lane < CHDB_N_OUT && P.out[lane].validity != nullptr && s_nulls[lane] != 0)
      asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(P.b.counts + P.out[lane].count_index), "l"((uint64_t)s_nulls[lane]) : "memory");
    return;
  }

  // ---- 2. predicate -> selection bits, ranks, slice counts ----
  uint32_t sels = 0;        // 4 selection bits per slice of this warp
  uint32_t ranks[kSpw];     // slice-local rank of the lane's first selected row
#pragma unroll
  for (int j = 0; j < kSpw; j++) {
    const int slice = warp * kSpw + j;
    const int64_t qb[1] = {row0 + slice * kWarpRows + lane * 4};
    const int left = tile_rows - (slice * kWarpRows + lane * 4);
    const uint32_t in4 = left >= 4 ? 0xFu : left <= 0 ? 0u : ((1u << left) - 1u);
    uint32_t sel4 = in4;
    if (MODE == kGather) {
      if (has_pred) sel4 = selq[j] & in4;
    } else if (has_pred) {
      V acc[4];
      uint32_t accm, accv;
#ifdef CHDB_JIT
      run_program<V, 1, chdb_jit::kPredBegin, chdb_jit::kPredEnd>(P, cols, 0, 0, qb, in4, in4, sh.pool, acc, accm, accv);
#else
      run_program<V, 1>(P, cols, P.pred_begin, P.pred_end, qb, in4, in4, sh.pool, acc, accm, accv);
#endif
      sel4 = accm & accv & in4;   // NULL predicate rows are dropped (arrow-select filter)
    }
    uint32_t wrows;
    ranks[j] = warp_excl_scan((uint32_t)__popc(sel4), lane, wrows);
    sels |= sel4 << (4 * j);
    if (lane == 0) s_cnt[slice] = wrows;
    if (has_pred) {
      // selected value bytes per Utf8 output
      CHDB_STATIC_UNROLL
      for (int k = 0; k < CHDB_N_OUT; k++) {
        const uint64_t meta = CHDB_OUT_META(P, k);
        const uint32_t o_utf8 = (uint32_t)(meta >> 48) & 0xFFu, o_slot = (uint32_t)(meta >> 24) & 0xFFu;
        if (o_utf8 == 0xFFu) continue;   // uniform branch
        const int32_t* off = cols[o_slot].offsets;
        uint32_t bytes = 0;
        if (sel4) {
          const int4 a = *(const int4*)(off + qb[0]);
          const int a4 = off[qb[0] + 4];
          if (sel4 & 1u) bytes += (uint32_t)(a.y - a.x);
          if (sel4 & 2u) bytes += (uint32_t)(a.z - a.y);
          if (sel4 & 4u) bytes += (uint32_t)(a.w - a.z);
          if (sel4 & 8u) bytes += (uint32_t)(a4 - a.w);
        }
        const uint32_t wbytes = __reduce_add_sync(FULL, bytes);
        if (lane == 0) s_cnt[(1 + o_utf8) * kTileSlices + slice] = wbytes;
      }
    }
  }

  // ---- 3. slice counts -> batch-wide exclusive prefixes ----
  if (has_pred) {
    __syncthreads();
    if (tid == 0) trace_event(P, 2);
    for (int q = warp; q < nq; q += kWarps) {
      const uint32_t c = lane < kTileSlices ? s_cnt[q * kTileSlices + lane] : 0u;
      uint32_t agg;
      const uint32_t before = warp_excl_scan(c, lane, agg);
      uint64_t* desc = P.b.desc + (size_t)q * (size_t)P.b.num_tiles;
      if (MODE == kGather) {
        const uint64_t excl = s_tot[q];   // (summed by this warp before the tile landed)
        __syncwarp();
        if (lane < kTileSlices) s_pre[q * kTileSlices + lane] = excl + before;
        if (lane == 0) s_tot[q] = excl + agg;
        continue;
      }
      if (lane == 0) publish_descriptor(desc + tile, tile == 0 ? kFlagPrefix : kFlagAgg, agg);
      const uint64_t excl = lookback(desc, (uint32_t)tile, agg, lane);
      if (lane < kTileSlices) s_pre[q * kTileSlices + lane] = excl + before;
      if (lane == 0) s_tot[q] = excl + agg;
      if (q == 0 && lane == 0) trace_event(P, 3);
    }
    __syncthreads();
    if (last_tile && !(MODE == kGather && CHDB_EARLY_COUNTS(P))) {
      // batch totals, and the closing Utf8 offset: offsets[total_rows] = total_bytes (also covers an empty result)
      if (tid < nq) P.b.counts[tid] = s_tot[tid];
      if (tid < CHDB_N_OUT) {
        const OutDesc& o = P.out[tid];
        if (o.utf8_index != 0xFFu) o.offsets[s_tot[0]] = (int32_t)s_tot[1 + o.utf8_index];
      }
    }
  }

  // ---- 4. the selected rows, to their final positions ----
  if (!TP.pred_reads_utf8 || MODE == kGather) mbar_wait(full_values, 0);
  if (tid == 0) trace_event(P, 4);
  uint32_t* const bitstage = bitstages + warp * P.n_bits * kBitWords;
#pragma unroll
  for (int j = 0; j < kSpw; j++) {
    const int slice = warp * kSpw + j;
    if (slice * kWarpRows >= tile_rows) break;   // (tail tile)
    LaneCtx L;
    L.lane = lane;
    L.slice = slice;
    L.wid = warp;
    L.row_base = row0 + slice * kWarpRows + lane * 4;
    L.row32 = (uint32_t)L.row_base;
    {
      const int left = tile_rows - (slice * kWarpRows + lane * 4);
      L.inrange = left >= 4 ? 0xFu : left <= 0 ? 0u : ((1u << left) - 1u);
    }
    L.sel = (sels >> (4 * j)) & 0xFu;
    L.rank = ranks[j];
    L.count = s_cnt[slice];
    if (has_pred) {
      L.prefix = s_pre;
      L.obase = (uint32_t)s_pre[slice];
    } else {
      L.prefix = nullptr;
      L.obase = (uint32_t)(row0 + slice * kWarpRows);
    }
    int kb = 0;
#ifdef CHDB_JIT
    {
      OutRegs R[chdb_jit::kNumOut > 0 ? chdb_jit::kNumOut : 1];
      load_outputs_range<0, chdb_jit::kNumOut>(P, cols, L, R);
      store_outputs_range<V, 0, chdb_jit::kNumOut>(P, cols, L, R, sh, bitstage, ltab, kb);
    }
#else
#pragma unroll 1
    for (int k = 0; k < P.n_out; k++) {
      OutRegs R;
      const uint64_t meta = CHDB_OUT_META(P, k);
      load_output(P, cols, k, meta, L, R);
      store_output<V>(P, cols, k, meta, L, R, sh, bitstage, ltab, kb);
    }
#endif
    if (P.n_bits > 0) {
      __syncwarp();
      flush_bits(P, bitstage, L.obase, L.count, lane);
      __syncwarp();
    }
  }

  if (MODE == kGather && CHDB_EARLY_COUNTS(P)) {   // (the select kernel did the bookkeeping)
    if (P.trace != nullptr) { __syncthreads(); if (tid == 0) { trace_event(P, 3); trace_event(P, 5); trace_event(P, 6); } }
    return;
  }

  // ---- null counts; the last CTA to finish mirrors the counts into pinned host memory (warp 0 only: the other warps
  //      are done once their rows are stored) ----
  __syncthreads();
  if (tid == 0) trace_event(P, 5);
  if (warp != 0) return;
  if (lane < CHDB_N_OUT && P.out[lane].validity != nullptr && s_nulls[lane] != 0)
    atomicAdd((unsigned long long*)(P.b.counts + P.out[lane].count_index), (unsigned long long)s_nulls[lane]);
  __syncwarp();
  uint32_t last = 0;
  if (lane == 0) {
    __threadfence();
    last = atomicAdd(P.b.done, 1u) == (uint32_t)P.b.num_tiles - 1u ? 1u : 0u;
  }
  last = __shfl_sync(FULL, last, 0);
  if (last) {
    __threadfence();
#pragma unroll 1
    for (int i = lane; i <= P.n_counts; i += 32) P.b.host_counts[i] = __ldcg((const unsigned long long*)P.b.counts + i);
  }
  if (tid == 0) trace_event(P, 6);
}

}  // namespace

// Zeroes the workspace of one launch (look-back descriptors, counts, error word, bit-packed outputs).
__device__ __forceinline__ void zero_body(uint4* p, size_t n16) {
  grid_launch_dependents();
  const size_t step = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += step) p[i] = make_uint4(0, 0, 0, 0);
}

#ifndef CHDB_JIT
template <typename V, bool MANY, int MODE>
__global__ void __launch_bounds__(kThreads, kMinCtasPerSm) stream_kernel(const __grid_constant__ KernelParams P, const __grid_constant__ TilePlan TP) {
  stream_body<V, MANY, MODE>(P, TP);
}
__global__ void __launch_bounds__(kZeroThreads) zero_kernel(uint4* p, size_t n16) { zero_body(p, n16); }
#endif

}  // namespace chdb

#ifdef CHDB_JIT
// the specialised kernels of one NVRTC module, found by their unmangled names
extern "C" __global__ void __launch_bounds__(chdb::kThreads, CHDB_JIT_MIN_BLOCKS) chdb_jit_stream(const __grid_constant__ chdb::KernelParams P, const __grid_constant__ chdb::TilePlan TP) {
  chdb::stream_body<chdb_jit::Container, false, chdb::kFused>(P, TP);
}
#ifndef CHDB_JIT_SELECT_MIN_BLOCKS
#define CHDB_JIT_SELECT_MIN_BLOCKS 8
#endif
extern "C" __global__ void __launch_bounds__(chdb::kThreads, CHDB_JIT_SELECT_MIN_BLOCKS) chdb_jit_select(const __grid_constant__ chdb::KernelParams P, const __grid_constant__ chdb::TilePlan TP) {
  chdb::stream_body<chdb_jit::Container, false, chdb::kSelect>(P, TP);
}
extern "C" __global__ void __launch_bounds__(chdb::kThreads, CHDB_JIT_MIN_BLOCKS) chdb_jit_gather(const __grid_constant__ chdb::KernelParams P, const __grid_constant__ chdb::TilePlan TP) {
  chdb::stream_body<chdb_jit::Container, false, chdb::kGather>(P, TP);
}
extern "C" __global__ void __launch_bounds__(chdb::kThreads, CHDB_JIT_MIN_BLOCKS) chdb_jit_stream_many(const __grid_constant__ chdb::KernelParams P, const __grid_constant__ chdb::TilePlan TP) {
  chdb::stream_body<chdb_jit::Container, true, chdb::kFused>(P, TP);
}
#endif
