// Fused WHERE-evaluation -> selection -> decoupled-look-back scan -> compaction kernel (sm_100a).
//
// One CTA owns one tile of kTileRows rows.  Per tile:
//   1. every thread runs the predicate bytecode over its rows (128-bit coalesced column loads,
//      accumulator in registers) and gets a selection mask                     [compute_value.rs]
//   2. warp shuffles + one shared-memory pass rank the selected rows; for each Utf8 output the
//      selected value bytes are summed the same way
//   3. a decoupled look-back over 64-bit {flag | value} tile descriptors turns the tile totals
//      into exclusive prefixes (rows, and bytes per Utf8 output)               [filter_record.rs:37]
//   4. every output column is gathered: values are staged in shared memory at the same
//      16-byte phase as their destination and written with full 16-byte stores; validity and
//      Boolean bits are packed with REDUX.OR; Utf8 bytes are produced output-chunk-centric so
//      every global store is an aligned 16-byte store.
// HBM traffic is therefore each referenced input byte once and each output byte once.
// Projection expressions are evaluated in step 4 under the selection mask, so checked-integer
// errors are raised for surviving rows only (the reference projects after filtering).
//
// Build with -fmad=false: float results must be the IEEE single operations arrow-rs performs.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/chdb_gpu.h"
#include "kernels.cuh"

namespace chdb {
namespace {

constexpr uint32_t FULL = 0xFFFFFFFFu;
constexpr int kBitStageWords = kTileRows / 32 + 8;

template <typename V> struct Cont;
template <> struct Cont<uint32_t> { using S = int32_t; static constexpr bool k64 = false; };
template <> struct Cont<uint64_t> { using S = int64_t; static constexpr bool k64 = true; };

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void report_error(const KernelParams& P, uint32_t order, int64_t row, uint32_t code) {
  unsigned long long packed = ((unsigned long long)order << 56) | (((unsigned long long)row & 0xFFFFFFFFFFFFull) << 8) | code;
  atomicMax((unsigned long long*)P.error_word, ~packed);
}

// ------------------------------------------------------------------------------------------
// loads
// ------------------------------------------------------------------------------------------
template <int QPT>
__device__ __forceinline__ uint32_t load_bits(const uint8_t* __restrict__ bits, const int64_t (&qbase)[QPT], uint32_t need) {
  if (bits == nullptr) return FULL;
  uint32_t m = 0;
#pragma unroll
  for (int q = 0; q < QPT; q++) {
    if ((need >> (4 * q)) & 0xFu) {
      uint32_t byte = __ldg(bits + (qbase[q] >> 3));
      m |= ((byte >> (uint32_t)(qbase[q] & 4)) & 0xFu) << (4 * q);
    }
  }
  return m;
}

// Raw little-endian values of the thread's rows, zero-extended to 64 bits (width 1/2/4/8).
template <int QPT>
__device__ __forceinline__ void fetch_raw(const ColumnDesc& c, const int64_t (&qbase)[QPT], uint32_t need,
                                          uint64_t (&e)[4 * QPT]) {
  const uint8_t* __restrict__ base = (const uint8_t*)c.values;
  const int width = c.width;
#pragma unroll
  for (int j = 0; j < 4 * QPT; j++) e[j] = 0;
#pragma unroll
  for (int q = 0; q < QPT; q++) {
    if (!((need >> (4 * q)) & 0xFu)) continue;
    const int64_t r = qbase[q];
    if (width == 4) {
      uint4 x = __ldg((const uint4*)(base + r * 4));
      e[4 * q + 0] = x.x; e[4 * q + 1] = x.y; e[4 * q + 2] = x.z; e[4 * q + 3] = x.w;
    } else if (width == 8) {
      uint4 x = __ldg((const uint4*)(base + r * 8));
      uint4 y = __ldg((const uint4*)(base + r * 8 + 16));
      e[4 * q + 0] = x.x | ((uint64_t)x.y << 32); e[4 * q + 1] = x.z | ((uint64_t)x.w << 32);
      e[4 * q + 2] = y.x | ((uint64_t)y.y << 32); e[4 * q + 3] = y.z | ((uint64_t)y.w << 32);
    } else if (width == 2) {
      uint2 x = __ldg((const uint2*)(base + r * 2));
      e[4 * q + 0] = x.x & 0xFFFFu; e[4 * q + 1] = x.x >> 16; e[4 * q + 2] = x.y & 0xFFFFu; e[4 * q + 3] = x.y >> 16;
    } else {
      uint32_t x = __ldg((const uint32_t*)(base + r));
      e[4 * q + 0] = x & 0xFFu; e[4 * q + 1] = (x >> 8) & 0xFFu; e[4 * q + 2] = (x >> 16) & 0xFFu; e[4 * q + 3] = x >> 24;
    }
  }
}

// Canonical accumulator form: integers sign-/zero-extended to the whole container.
template <typename V>
__device__ __forceinline__ V extend(uint64_t raw, uint8_t t) {
  using S = typename Cont<V>::S;
  switch (t) {
    case T_I8: return (V)(S)(int8_t)raw;
    case T_I16: return (V)(S)(int16_t)raw;
    case T_I32: return (V)(S)(int32_t)raw;
    default: return (V)raw;
  }
}

// ------------------------------------------------------------------------------------------
// casts (arrow-cast on the coercion lattice; int -> float is round-to-nearest-even)
// ------------------------------------------------------------------------------------------
template <typename V, int R>
__device__ __forceinline__ void cast_vals(V (&a)[R], uint8_t from, uint8_t to) {
  using S = typename Cont<V>::S;
  const TypeClass fc = type_class(from), tc = type_class(to);
  if (fc == tc) return;
  if (tc == C_F32) {
    if (fc == C_SINT || fc == C_S64) {
#pragma unroll
      for (int j = 0; j < R; j++) a[j] = (V)__float_as_uint((float)(S)a[j]);
    } else if (fc == C_UINT || fc == C_U64) {
#pragma unroll
      for (int j = 0; j < R; j++) a[j] = (V)__float_as_uint((float)a[j]);
    }
  } else if (tc == C_F64) {
    if constexpr (Cont<V>::k64) {
      if (fc == C_SINT || fc == C_S64) {
#pragma unroll
        for (int j = 0; j < R; j++) a[j] = (V)__double_as_longlong((double)(int64_t)a[j]);
      } else if (fc == C_UINT || fc == C_U64) {
#pragma unroll
        for (int j = 0; j < R; j++) a[j] = (V)__double_as_longlong((double)(uint64_t)a[j]);
      } else if (fc == C_F32) {
#pragma unroll
        for (int j = 0; j < R; j++) {
          uint32_t u = (uint32_t)a[j];
          float x = __uint_as_float(u);
          uint64_t r;
          if (x != x)  // keep sign and payload, quiet (x86 cvtss2sd)
            r = ((uint64_t)(u & 0x80000000u) << 32) | 0x7FF8000000000000ull | ((uint64_t)(u & 0x007FFFFFu) << 29);
          else
            r = (uint64_t)__double_as_longlong((double)x);
          a[j] = (V)r;
        }
      }
    }
  }
  // integer -> integer widening: the canonical container already holds the value
}

template <typename V, int R>
__device__ __forceinline__ uint32_t tobool_vals(const V (&a)[R], uint8_t t) {
  const TypeClass c = type_class(t);
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < R; j++) {
    bool b;
    if (c == C_F32) b = __uint_as_float((uint32_t)a[j]) != 0.0f;   // NaN -> true, -0.0 -> false
    else if (c == C_F64) b = __longlong_as_double((long long)(uint64_t)a[j]) != 0.0;
    else b = a[j] != 0;
    m |= (b ? 1u : 0u) << j;
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

template <typename V>
__device__ __forceinline__ int64_t narrow_i64(V x, bool is_signed) {
  if constexpr (Cont<V>::k64) return (int64_t)x;
  else return is_signed ? (int64_t)(int32_t)x : (int64_t)(uint32_t)x;
}

template <typename V, int QPT>
__device__ __forceinline__ void arith(const KernelParams& P, const Instr& in, V (&a)[4 * QPT], uint32_t& av, const V (&b)[4 * QPT],
                                      uint32_t bv, uint32_t active, const int64_t (&qbase)[QPT]) {
  constexpr int R = 4 * QPT;
  const bool swap = (in.flags & OPF_SWAP) != 0;
  const uint8_t op = in.op;
  const TypeClass tc = type_class(in.type);
  const uint32_t valid = av & bv;
  av = valid;
  const uint32_t m = valid & active;  // fallible ops run on valid, live rows only; other slots hold 0
#define CHDB_ROW(j) (qbase[(j) >> 2] + ((j) & 3))
  if (tc == C_SINT || tc == C_UINT) {
    const bool sgn = tc == C_SINT;
    int64_t lo, hi;
    switch (in.type) {
      case T_I8: lo = -128; hi = 127; break;
      case T_I16: lo = -32768; hi = 32767; break;
      case T_I32: lo = -2147483648ll; hi = 2147483647ll; break;
      case T_U8: lo = 0; hi = 255; break;
      case T_U16: lo = 0; hi = 65535; break;
      default: lo = 0; hi = 4294967295ll; break;
    }
#pragma unroll
    for (int j = 0; j < R; j++) {
      int64_t r = 0;
      if ((m >> j) & 1u) {
        const int64_t x = narrow_i64<V>(swap ? b[j] : a[j], sgn), y = narrow_i64<V>(swap ? a[j] : b[j], sgn);
        if (op == OP_DIV || op == OP_REM) {
          if (y == 0) {
            report_error(P, in.order, CHDB_ROW(j), CHDB_ERR_DIVIDE_BY_ZERO);
          } else if (sgn && x == lo && y == -1) {
            report_error(P, in.order, CHDB_ROW(j), CHDB_ERR_ARITHMETIC_OVERFLOW);
          } else if (sgn) {
            r = op == OP_DIV ? (int64_t)((int32_t)x / (int32_t)y) : (int64_t)((int32_t)x % (int32_t)y);
          } else {
            r = op == OP_DIV ? (int64_t)((uint32_t)x / (uint32_t)y) : (int64_t)((uint32_t)x % (uint32_t)y);
          }
        } else {
          r = op == OP_ADD ? x + y : op == OP_MUL ? x * y : x - y;
          if (r < lo || r > hi) {
            report_error(P, in.order, CHDB_ROW(j), CHDB_ERR_ARITHMETIC_OVERFLOW);
            r = 0;
          }
        }
      }
      a[j] = (V)r;
    }
  } else if (tc == C_F32) {
#pragma unroll
    for (int j = 0; j < R; j++) {
      const float x = __uint_as_float((uint32_t)(swap ? b[j] : a[j])), y = __uint_as_float((uint32_t)(swap ? a[j] : b[j]));
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
  } else {
    if constexpr (Cont<V>::k64) {
      if (tc == C_F64) {
#pragma unroll
        for (int j = 0; j < R; j++) {
          const double x = __longlong_as_double((long long)(swap ? b[j] : a[j])), y = __longlong_as_double((long long)(swap ? a[j] : b[j]));
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
      } else if (tc == C_S64) {
#pragma unroll
        for (int j = 0; j < R; j++) {
          int64_t r = 0;
          if ((m >> j) & 1u) {
            const int64_t x = (int64_t)(swap ? b[j] : a[j]), y = (int64_t)(swap ? a[j] : b[j]);
            bool ovf = false;
            if (op == OP_DIV || op == OP_REM) {
              if (y == 0) report_error(P, in.order, CHDB_ROW(j), CHDB_ERR_DIVIDE_BY_ZERO);
              else if (x == INT64_MIN && y == -1) ovf = true;
              else r = op == OP_DIV ? x / y : x % y;
            } else if (op == OP_ADD) {
              r = (int64_t)((uint64_t)x + (uint64_t)y);
              ovf = ((x ^ r) & (y ^ r)) < 0;
            } else if (op == OP_MUL) {
              r = (int64_t)((uint64_t)x * (uint64_t)y);
              ovf = __mul64hi(x, y) != (r >> 63);
            } else {
              r = (int64_t)((uint64_t)x - (uint64_t)y);
              ovf = ((x ^ y) & (x ^ r)) < 0;
            }
            if (ovf) {
              report_error(P, in.order, CHDB_ROW(j), CHDB_ERR_ARITHMETIC_OVERFLOW);
              r = 0;
            }
          }
          a[j] = (V)r;
        }
      } else {  // C_U64
#pragma unroll
        for (int j = 0; j < R; j++) {
          uint64_t r = 0;
          if ((m >> j) & 1u) {
            const uint64_t x = (uint64_t)(swap ? b[j] : a[j]), y = (uint64_t)(swap ? a[j] : b[j]);
            bool ovf = false;
            if (op == OP_DIV || op == OP_REM) {
              if (y == 0) report_error(P, in.order, CHDB_ROW(j), CHDB_ERR_DIVIDE_BY_ZERO);
              else r = op == OP_DIV ? x / y : x % y;
            } else if (op == OP_ADD) {
              r = x + y;
              ovf = r < x;
            } else if (op == OP_MUL) {
              r = x * y;
              ovf = __umul64hi(x, y) != 0;
            } else {
              r = x - y;
              ovf = x < y;
            }
            if (ovf) {
              report_error(P, in.order, CHDB_ROW(j), CHDB_ERR_ARITHMETIC_OVERFLOW);
              r = 0;
            }
          }
          a[j] = (V)r;
        }
      }
    }
  }
#undef CHDB_ROW
}

// ------------------------------------------------------------------------------------------
// comparisons (arrow-ord cmp.rs: natural integer order, IEEE-754 totalOrder for floats)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cmp_select(uint8_t kind, uint32_t lt, uint32_t eq) {
  switch (kind) {
    case CMP_EQ: return eq;
    case CMP_NE: return ~eq;
    case CMP_LT: return lt;
    case CMP_LE: return lt | eq;
    case CMP_GT: return ~(lt | eq);
    default: return ~lt;
  }
}

template <typename V, int R>
__device__ __forceinline__ uint32_t compare(const Instr& in, const V (&a)[R], uint32_t am, const V (&b)[R], uint32_t bm) {
  using S = typename Cont<V>::S;
  const TypeClass tc = type_class(in.type);
  uint32_t lt = 0, eq = 0;
  if (tc == C_BOOL) {  // false < true
    lt = ~am & bm;
    eq = ~(am ^ bm);
  } else {
#pragma unroll
    for (int j = 0; j < R; j++) {
      bool l, e;
      if (tc == C_SINT) {
        l = (S)a[j] < (S)b[j];
        e = a[j] == b[j];
      } else if (tc == C_F32) {
        int32_t x = (int32_t)(uint32_t)a[j], y = (int32_t)(uint32_t)b[j];
        e = x == y;  // bitwise: NaN == NaN with equal payloads, -0.0 != +0.0
        x ^= (int32_t)(((uint32_t)(x >> 31)) >> 1);
        y ^= (int32_t)(((uint32_t)(y >> 31)) >> 1);
        l = x < y;
      } else if (tc == C_S64) {
        l = (int64_t)a[j] < (int64_t)b[j];
        e = a[j] == b[j];
      } else if (tc == C_F64) {
        int64_t x = (int64_t)a[j], y = (int64_t)b[j];
        e = x == y;
        x ^= (int64_t)(((uint64_t)(x >> 63)) >> 1);
        y ^= (int64_t)(((uint64_t)(y >> 63)) >> 1);
        l = x < y;
      } else {  // C_UINT, C_U64: zero-extended containers compare unsigned
        l = a[j] < b[j];
        e = a[j] == b[j];
      }
      lt |= (l ? 1u : 0u) << j;
      eq |= (e ? 1u : 0u) << j;
    }
  }
  return cmp_select(in.aux, lt, eq);
}

// Utf8: bytewise lexicographic; operands are columns or a literal from the string pool.
template <int QPT>
__device__ __forceinline__ uint32_t cmp_utf8(const KernelParams& P, const Instr& in, const int64_t (&qbase)[QPT], uint32_t inrange,
                                             const uint8_t* s_pool, uint32_t& valid) {
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
      int o0 = __ldg(off + row), o1 = __ldg(off + row + 1);
      pa = (const uint8_t*)P.in[slot_a].values + o0;
      la = o1 - o0;
    } else {
      pa = s_pool + pool_off;
      la = (int)pool_len;
    }
    if (slot_b != 0xFFu) {
      const int32_t* off = P.in[slot_b].offsets;
      int o0 = __ldg(off + row), o1 = __ldg(off + row + 1);
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
  return cmp_select(in.aux, lt, eq);
}

// ------------------------------------------------------------------------------------------
// the interpreter: accumulator in registers, one operand per instruction
// ------------------------------------------------------------------------------------------
template <typename V, int QPT>
__device__ __forceinline__ void run_program(const KernelParams& P, int begin, int end, const int64_t (&qbase)[QPT], uint32_t inrange,
                                            uint32_t active, const uint8_t* s_pool, V (&acc)[4 * QPT], uint32_t& accm,
                                            uint32_t& accv) {
  constexpr int R = 4 * QPT;
  V stk[kMaxSpill][R];
  uint32_t stkm[kMaxSpill], stkv[kMaxSpill];
#pragma unroll
  for (int j = 0; j < R; j++) acc[j] = 0;
  accm = 0;
  accv = FULL;
#pragma unroll 1
  for (int pc = begin; pc < end; pc++) {
    const Instr in = P.instrs[pc];
    V b[R];
    uint32_t bm = 0, bv = FULL;
#pragma unroll
    for (int j = 0; j < R; j++) b[j] = 0;
    if (in.src == SRC_IMM) {
#pragma unroll
      for (int j = 0; j < R; j++) b[j] = (V)in.imm;
      bm = in.imm ? FULL : 0u;
    } else if (in.src == SRC_STK) {
      if (in.type != T_BOOL) {   // Boolean spills only carry the two masks
#pragma unroll
        for (int j = 0; j < R; j++) b[j] = stk[in.slot][j];
      }
      bm = stkm[in.slot];
      bv = stkv[in.slot];
    } else if (in.src == SRC_COL) {
      const ColumnDesc& c = P.in[in.slot];
      bv = load_bits<QPT>(c.validity, qbase, inrange);
      if (c.type == T_BOOL) {
        bm = load_bits<QPT>((const uint8_t*)c.values, qbase, inrange);
      } else {
        uint64_t raw[R];
        fetch_raw<QPT>(c, qbase, inrange, raw);
#pragma unroll
        for (int j = 0; j < R; j++) b[j] = extend<V>(raw[j], in.from_type);
        if (in.type == T_BOOL) bm = tobool_vals<V, R>(b, in.from_type);
        else cast_vals<V, R>(b, in.from_type, in.type);
      }
    }
    switch (in.op) {
      case OP_LOAD:
#pragma unroll
        for (int j = 0; j < R; j++) acc[j] = b[j];
        accm = bm;
        accv = bv;
        break;
      case OP_CAST: cast_vals<V, R>(acc, in.from_type, in.type); break;
      case OP_ADD: case OP_MUL: case OP_DIV: case OP_REM: case OP_SUB:
        arith<V, QPT>(P, in, acc, accv, b, bv, active, qbase);
        break;
      case OP_CMP:
        accm = compare<V, R>(in, acc, accm, b, bm);
        accv &= bv;
        break;
      case OP_TOBOOL: accm = tobool_vals<V, R>(acc, in.type); break;
      case OP_AND: accm &= bm; accv &= bv; break;   // non-Kleene: null if either side is null
      case OP_OR: accm |= bm; accv &= bv; break;
      case OP_PUSH:
        if (in.type != T_BOOL) {
#pragma unroll
          for (int j = 0; j < R; j++) stk[in.slot][j] = acc[j];
        }
        stkm[in.slot] = accm;
        stkv[in.slot] = accv;
        break;
      case OP_CMP_UTF8: accm = cmp_utf8<QPT>(P, in, qbase, inrange, s_pool, accv); break;
      default: break;
    }
  }
}

// ------------------------------------------------------------------------------------------
// scans
// ------------------------------------------------------------------------------------------
// Exclusive scan of per-(thread, quad) values in output order (quad group, warp, lane).
template <int QPT>
__device__ __forceinline__ void block_scan(const uint32_t (&val)[QPT], uint32_t (&excl)[QPT], uint32_t& total,
                                           uint32_t (*s_w)[kWarps], int lane, int warp) {
  uint32_t incl[QPT];
#pragma unroll
  for (int q = 0; q < QPT; q++) {
    uint32_t x = val[q];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t y = __shfl_up_sync(FULL, x, d);
      if (lane >= d) x += y;
    }
    incl[q] = x;
    if (lane == 31) s_w[q][warp] = x;
  }
  __syncthreads();
  uint32_t run = 0;
#pragma unroll
  for (int q = 0; q < QPT; q++) {
#pragma unroll
    for (int w = 0; w < kWarps; w++) {
      if (w == warp) excl[q] = run + incl[q] - val[q];
      run += s_w[q][w];
    }
  }
  total = run;
  __syncthreads();
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
    const int64_t idx = base - lane;
    uint64_t v = kFlagPrefix;  // tiles "before 0" contribute an inclusive prefix of 0
    if (idx >= 0) {
      do { v = d[idx]; } while ((v >> 62) == 0);
    }
    const uint32_t pm = __ballot_sync(FULL, (v >> 62) == 2);
    const uint64_t val = v & kValueMask;
    if (pm) {
      const int first = __ffs(pm) - 1;  // nearest predecessor that already knows its prefix
      excl += warp_sum64(lane <= first ? val : 0);
      break;
    }
    excl += warp_sum64(val);
    base -= 32;
  }
  if (lane == 0) d[tile] = kFlagPrefix | (excl + agg);
  return excl;
}

// ------------------------------------------------------------------------------------------
// output staging
// ------------------------------------------------------------------------------------------
template <int QPT>
__device__ __forceinline__ void stage_scatter(uint8_t* stage, uint32_t mis, int W, const uint64_t (&e)[4 * QPT], uint32_t sel,
                                              const uint32_t (&rank0)[QPT]) {
#pragma unroll
  for (int q = 0; q < QPT; q++) {
    uint32_t r = rank0[q];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int j = 4 * q + i;
      if ((sel >> j) & 1u) {
        uint8_t* p = stage + mis + (size_t)r * W;
        if (W == 4) *(uint32_t*)p = (uint32_t)e[j];
        else if (W == 8) *(uint64_t*)p = e[j];
        else if (W == 2) *(uint16_t*)p = (uint16_t)e[j];
        else *p = (uint8_t)e[j];
        r++;
      }
    }
  }
}

// stage[mis, mis + nbytes) -> gdst_aligned[mis, ...): aligned 16-byte stores in the middle,
// element stores in the (at most two) chunks shared with neighbouring tiles.
__device__ __forceinline__ void stage_writeout(const uint8_t* stage, uint8_t* gdst_aligned, uint32_t mis, uint32_t nbytes, int W,
                                               int tid) {
  const uint32_t end = mis + nbytes;
  const uint32_t nchunks = (end + 15u) >> 4;
  for (uint32_t c = tid; c < nchunks; c += kThreads) {
    const uint32_t lo = c << 4, hi = lo + 16;
    if (lo >= mis && hi <= end) {
      *(uint4*)(gdst_aligned + lo) = *(const uint4*)(stage + lo);
    } else {
      const uint32_t s = lo > mis ? lo : mis, t = hi < end ? hi : end;
      for (uint32_t b = s; b < t; b += W) {
        if (W == 4) *(uint32_t*)(gdst_aligned + b) = *(const uint32_t*)(stage + b);
        else if (W == 8) *(uint64_t*)(gdst_aligned + b) = *(const uint64_t*)(stage + b);
        else if (W == 2) *(uint16_t*)(gdst_aligned + b) = *(const uint16_t*)(stage + b);
        else gdst_aligned[b] = stage[b];
      }
    }
  }
}

// Compacts one bit per row (validity or Boolean values) into gbits at bit offset tile_prefix.
// gbits is zero-initialised; words shared with neighbouring tiles are merged with atomicOr.
template <int QPT>
__device__ __forceinline__ void compact_bits(uint32_t bits, uint32_t sel, const uint32_t (&rank0)[QPT], uint64_t tile_prefix,
                                             uint32_t tile_count, uint32_t* bitstage, uint32_t* gbits, int tid, int lane) {
  const uint32_t o = (uint32_t)(tile_prefix & 31);
  const uint32_t nwords = (o + tile_count + 31u) >> 5;
  for (uint32_t k = tid; k < nwords; k += kThreads) bitstage[k] = 0;
  __syncthreads();
#pragma unroll
  for (int q = 0; q < QPT; q++) {
    const uint32_t s4 = (sel >> (4 * q)) & 0xFu, b4 = (bits >> (4 * q)) & 0xFu;
    uint32_t cb = 0, n = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if ((s4 >> i) & 1u) {
        cb |= ((b4 >> i) & 1u) << n;
        n++;
      }
    }
    const uint32_t pos = o + rank0[q];
    const uint32_t w0 = __shfl_sync(FULL, pos, 0) >> 5;   // first staging word this warp touches
    const uint32_t rel = pos - (w0 << 5);                  // < 32 + 128
    const uint64_t v = (uint64_t)cb << (rel & 31u);
    const uint32_t wi = rel >> 5;
#pragma unroll
    for (uint32_t k = 0; k < 6; k++) {
      const uint32_t contrib = (wi == k) ? (uint32_t)v : ((wi + 1 == k) ? (uint32_t)(v >> 32) : 0u);
      const uint32_t word = __reduce_or_sync(FULL, contrib);
      if (lane == (int)k && word) atomicOr(&bitstage[w0 + k], word);
    }
  }
  __syncthreads();
  const uint64_t g0 = tile_prefix >> 5;
  for (uint32_t k = tid; k < nwords; k += kThreads) {
    const uint32_t w = bitstage[k];
    const uint64_t bit_lo = (g0 + k) << 5;
    const bool full = bit_lo >= tile_prefix && bit_lo + 32 <= tile_prefix + tile_count;
    if (full) gbits[g0 + k] = w;
    else if (w) atomicOr(&gbits[g0 + k], w);
  }
  __syncthreads();
}

__device__ __forceinline__ void add_count(uint64_t* slot, uint32_t mine, int lane) {
  const uint32_t s = __reduce_add_sync(FULL, mine);
  if (lane == 0 && s) atomicAdd((unsigned long long*)slot, (unsigned long long)s);
}

// 16 bytes from an arbitrarily aligned global address (buffers are padded, so the aligned words
// around it are always readable).
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

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
template <typename V, int QPT>
__global__ void __launch_bounds__(kThreads) filter_project_kernel(const __grid_constant__ KernelParams P) {
  constexpr int R = 4 * QPT;
  constexpr int T = kThreads * R;
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ uint32_t s_tile;
  __shared__ uint32_t s_w[QPT][kWarps];
  __shared__ uint64_t s_agg[1 + kMaxOutCols];
  __shared__ uint64_t s_excl[1 + kMaxOutCols];
  __shared__ uint8_t s_pool[kStrPoolBytes];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool has_pred = P.pred_end > P.pred_begin;

  uint8_t* stage = smem;
  uint32_t* bitstage = (uint32_t*)(smem + P.stage_bytes);
  uint32_t* s_oo = bitstage + kBitStageWords;   // [T + 1] tile-local output byte offsets (Utf8)
  int32_t* s_src = (int32_t*)(s_oo + T + 4);    // [T] source byte offsets (Utf8)

  if (tid == 0) s_tile = has_pred ? atomicAdd(P.ticket, 1u) : blockIdx.x;
  if (tid < kStrPoolBytes) s_pool[tid] = (uint8_t)P.strpool[tid];
  __syncthreads();
  const uint32_t tile = s_tile;
  const int64_t row0 = (int64_t)tile * T;

  int64_t qbase[QPT];
  uint32_t inrange = 0;
#pragma unroll
  for (int q = 0; q < QPT; q++) {
    qbase[q] = row0 + (int64_t)q * (kThreads * 4) + warp * 128 + lane * 4;
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (qbase[q] + i < P.num_rows) inrange |= 1u << (4 * q + i);
  }

  // ---- 1. predicate -> selection mask -----------------------------------------------------
  uint32_t sel = inrange;
  if (has_pred) {
    V acc[R];
    uint32_t accm, accv;
    run_program<V, QPT>(P, P.pred_begin, P.pred_end, qbase, inrange, inrange, s_pool, acc, accm, accv);
    sel = accm & accv & inrange;  // NULL predicate rows are dropped (arrow-select filter)
  }

  // ---- 2. rank the selected rows ------------------------------------------------------------
  uint32_t cnt[QPT], rank0[QPT], tile_count;
#pragma unroll
  for (int q = 0; q < QPT; q++) cnt[q] = __popc((sel >> (4 * q)) & 0xFu);
  block_scan<QPT>(cnt, rank0, tile_count, s_w, lane, warp);
  if (tid == 0) s_agg[0] = tile_count;

  // selected value bytes per Utf8 output
  for (int k = 0; k < P.n_out; k++) {
    const OutDesc& o = P.out[k];
    if (o.utf8_index == 0xFFu) continue;   // uniform branch
    const int32_t* __restrict__ off = P.in[o.slot].offsets;
    uint32_t bytes[QPT], bexcl[QPT], btotal;
#pragma unroll
    for (int q = 0; q < QPT; q++) {
      bytes[q] = 0;
      const uint32_t s4 = (sel >> (4 * q)) & 0xFu;
      if (s4) {
        const int4 a = __ldg((const int4*)(off + qbase[q]));
        const int a4 = __ldg(off + qbase[q] + 4);
        if (s4 & 1u) bytes[q] += (uint32_t)(a.y - a.x);
        if (s4 & 2u) bytes[q] += (uint32_t)(a.z - a.y);
        if (s4 & 4u) bytes[q] += (uint32_t)(a.w - a.z);
        if (s4 & 8u) bytes[q] += (uint32_t)(a4 - a.w);
      }
    }
    block_scan<QPT>(bytes, bexcl, btotal, s_w, lane, warp);
    if (tid == 0) s_agg[1 + o.utf8_index] = btotal;
  }
  __syncthreads();

  // ---- 3. tile prefixes ---------------------------------------------------------------------
  const int nq = 1 + P.n_utf8;
  if (has_pred) {
    for (int qi = warp; qi < nq; qi += kWarps) {
      const uint64_t agg = s_agg[qi];
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
  const uint64_t tile_prefix = s_excl[0];
  if (tile == (uint32_t)P.num_tiles - 1 && tid < P.n_out) {
    // closing Utf8 offset: offsets[total_rows] = total_bytes (also covers an empty result)
    const OutDesc& o = P.out[tid];
    if (o.utf8_index != 0xFFu)
      o.offsets[tile_prefix + tile_count] = (int32_t)(s_excl[1 + o.utf8_index] + s_agg[1 + o.utf8_index]);
  }
  if (tile_count == 0) return;  // block-uniform

  // ---- 4. gather every output column --------------------------------------------------------
  for (int k = 0; k < P.n_out; k++) {
    const OutDesc o = P.out[k];
    uint32_t vbits = FULL;  // validity of this output for the thread's rows
    if (o.kind == OUT_EXPR) {
      V acc[R];
      uint32_t accm, accv;
      // `sel` as the active mask: checked arithmetic only sees rows that survived the filter
      run_program<V, QPT>(P, o.begin, o.end, qbase, inrange, sel, s_pool, acc, accm, accv);
      vbits = accv;
      if (o.type == T_BOOL) {
        compact_bits<QPT>(accm, sel, rank0, tile_prefix, tile_count, bitstage, (uint32_t*)o.values, tid, lane);
      } else {
        uint64_t e[R];
#pragma unroll
        for (int j = 0; j < R; j++) e[j] = (uint64_t)acc[j];
        const uint32_t mis = (uint32_t)((tile_prefix * o.width) & 15u);
        stage_scatter<QPT>(stage, mis, o.width, e, sel, rank0);
        __syncthreads();
        stage_writeout(stage, (uint8_t*)o.values + ((tile_prefix * o.width) & ~15ull), mis, tile_count * o.width, o.width, tid);
        __syncthreads();
      }
    } else {
      const ColumnDesc& c = P.in[o.slot];
      vbits = load_bits<QPT>(c.validity, qbase, sel);
      if (o.type == T_BOOL) {
        const uint32_t vals = load_bits<QPT>((const uint8_t*)c.values, qbase, sel);
        compact_bits<QPT>(vals, sel, rank0, tile_prefix, tile_count, bitstage, (uint32_t*)o.values, tid, lane);
      } else if (o.type == T_UTF8) {
        // -- offsets: running sum of the selected lengths, restarted at 0 for the output --
        const int32_t* __restrict__ off = c.offsets;
        const uint64_t byte_prefix = s_excl[1 + o.utf8_index];
        const uint32_t tile_bytes = (uint32_t)s_agg[1 + o.utf8_index];
        uint32_t len[R], bytes[QPT], bexcl[QPT], btotal;
        int32_t src[R];
#pragma unroll
        for (int q = 0; q < QPT; q++) {
          bytes[q] = 0;
#pragma unroll
          for (int i = 0; i < 4; i++) { len[4 * q + i] = 0; src[4 * q + i] = 0; }
          const uint32_t s4 = (sel >> (4 * q)) & 0xFu;
          if (s4) {
            const int4 a = __ldg((const int4*)(off + qbase[q]));
            const int a4 = __ldg(off + qbase[q] + 4);
            src[4 * q + 0] = a.x; src[4 * q + 1] = a.y; src[4 * q + 2] = a.z; src[4 * q + 3] = a.w;
            if (s4 & 1u) len[4 * q + 0] = (uint32_t)(a.y - a.x);
            if (s4 & 2u) len[4 * q + 1] = (uint32_t)(a.z - a.y);
            if (s4 & 4u) len[4 * q + 2] = (uint32_t)(a.w - a.z);
            if (s4 & 8u) len[4 * q + 3] = (uint32_t)(a4 - a.w);
            bytes[q] = len[4 * q] + len[4 * q + 1] + len[4 * q + 2] + len[4 * q + 3];
          }
        }
        block_scan<QPT>(bytes, bexcl, btotal, s_w, lane, warp);
#pragma unroll
        for (int q = 0; q < QPT; q++) {
          uint32_t r = rank0[q], bo = bexcl[q];
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int j = 4 * q + i;
            if ((sel >> j) & 1u) {
              s_oo[r] = bo;
              s_src[r] = src[j];
              bo += len[j];
              r++;
            }
          }
        }
        if (tid == 0) s_oo[tile_count] = tile_bytes;
        __syncthreads();
        for (uint32_t r = tid; r < tile_count; r += kThreads) o.offsets[tile_prefix + r] = (int32_t)(byte_prefix + s_oo[r]);
        // -- value bytes: each thread produces aligned 16-byte output chunks --
        const uint8_t* __restrict__ sv = (const uint8_t*)c.values;
        const uint32_t mis = (uint32_t)(byte_prefix & 15u);
        uint8_t* gal = (uint8_t*)o.values + (byte_prefix - mis);
        const uint32_t end = mis + tile_bytes;
        const uint32_t nchunks = (end + 15u) >> 4;
        for (uint32_t ch = tid; ch < nchunks; ch += kThreads) {
          const uint32_t lo = ch << 4, hi = lo + 16;
          const uint32_t s = lo > mis ? lo : mis, t = hi < end ? hi : end;
          if (s >= t) continue;
          const uint32_t x = s - mis;  // tile-local output byte index of the first byte produced
          uint32_t lo_r = 0, hi_r = tile_count;  // first r in (0, count] with s_oo[r] > x
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
          uint32_t words[4] = {0u, 0u, 0u, 0u};
          for (uint32_t b = s; b < t;) {
            const uint32_t xb = b - mis;
            while (xb >= s_oo[r + 1]) r++;
            const uint8_t* sp = sv + s_src[r] + (xb - s_oo[r]);
            if (((b & 3u) == 0) && b + 4 <= t && xb + 4 <= s_oo[r + 1]) {
              words[(b - lo) >> 2] = load4_unaligned(sp);
              b += 4;
            } else {
              words[(b - lo) >> 2] |= (uint32_t)__ldg(sp) << (8u * (b & 3u));
              b += 1;
            }
          }
          if (full) {
            *(uint4*)(gal + lo) = make_uint4(words[0], words[1], words[2], words[3]);
          } else {
            for (uint32_t b = s; b < t; b++) gal[b] = (uint8_t)(words[(b - lo) >> 2] >> (8u * (b & 3u)));
          }
        }
        __syncthreads();
      } else if (o.width == 16) {
        const uint4* __restrict__ src = (const uint4*)c.values;
        uint4* st = (uint4*)stage;
#pragma unroll
        for (int q = 0; q < QPT; q++) {
          uint32_t r = rank0[q];
#pragma unroll
          for (int i = 0; i < 4; i++)
            if ((sel >> (4 * q + i)) & 1u) st[r++] = __ldg(src + qbase[q] + i);
        }
        __syncthreads();
        uint4* dst = (uint4*)o.values + tile_prefix;
        for (uint32_t r = tid; r < tile_count; r += kThreads) dst[r] = st[r];
        __syncthreads();
      } else {
        uint64_t e[R];
        fetch_raw<QPT>(c, qbase, sel, e);
        const uint32_t mis = (uint32_t)((tile_prefix * o.width) & 15u);
        stage_scatter<QPT>(stage, mis, o.width, e, sel, rank0);
        __syncthreads();
        stage_writeout(stage, (uint8_t*)o.values + ((tile_prefix * o.width) & ~15ull), mis, tile_count * o.width, o.width, tid);
        __syncthreads();
      }
    }
    if (o.validity != nullptr) {
      compact_bits<QPT>(vbits, sel, rank0, tile_prefix, tile_count, bitstage, (uint32_t*)o.validity, tid, lane);
      add_count(P.counts + o.count_index, (uint32_t)__popc(sel & ~vbits), lane);
    }
  }
}

}  // namespace

size_t filter_project_smem_bytes(int max_out_width, bool has_utf8_out) {
  size_t stage = (size_t)kTileRows * (size_t)(max_out_width < 4 ? 4 : max_out_width) + 32;
  stage = (stage + 15) & ~(size_t)15;
  size_t total = stage + (size_t)kBitStageWords * 4;
  if (has_utf8_out) total += (size_t)(kTileRows + 4) * 4 + (size_t)kTileRows * 4 + 16;
  return total;
}

cudaError_t launch_filter_project(const KernelParams& p, bool has64, size_t dyn_smem, cudaStream_t stream) {
  auto k32 = filter_project_kernel<uint32_t, kQuadsPerThread>;
  auto k64 = filter_project_kernel<uint64_t, kQuadsPerThread>;
  auto kern = has64 ? k64 : k32;
  if (dyn_smem > 48 * 1024) {  // only 16-byte (decimal128) columns need more than the default window
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem);
    if (e != cudaSuccess) return e;
  }
  kern<<<dim3((unsigned)p.num_tiles), dim3(kThreads), dyn_smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace chdb
