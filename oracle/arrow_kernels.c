/*
 * oracle/arrow_kernels.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the arrow-rs 53 compute kernels that the
 * reference's hot path delegates to.  The reference calls them from
 *   src/handlers/operator_handler/operators/record_utils/compute_value.rs
 *     :72-73,:95-96   compute::cast(x, Boolean)        -> ora_cast_to_bool
 *     :80, :103       compute::and / compute::or       -> ora_bitmap_and / ora_bitmap_or
 *     :119,:128,:137,:146  numeric::{add,div,mul,rem}  -> ora_arith
 *     :155-:203       cmp::{eq,neq,gt,gt_eq,lt,lt_eq}  -> ora_cmp / ora_cmp_utf8 / ora_cmp_bool
 *     :440,:446       compute::cast(x, common_type)    -> ora_cast
 *   src/handlers/operator_handler/operators/record_utils/filter_record.rs
 *     :37             compute::filter_record_batch     -> ora_filter_*
 *
 * The `arrow` crate (arrow = "53.1", Cargo.toml:42) is NOT vendored under
 * /root/reference, so these functions restate its published semantics
 * (SURVEY.md section 8a rows a3-a8):
 *   - integer add/mul/div/rem are CHECKED (ArithmeticOverflow / DivideByZero),
 *     evaluated only on valid slots, null slots hold 0;
 *   - float add/mul/div/rem are plain IEEE-754, evaluated on every slot;
 *   - float comparisons use IEEE-754 totalOrder (eq is bitwise equality);
 *   - cast int->float is round-to-nearest-even, numeric->bool is x != 0;
 *   - filter keeps row order, rebuilds Utf8 offsets from 0.
 *
 * NaN results: an x86-64 host (what arrow-rs runs on beside a B200) produces
 * the "default NaN" with the SIGN BIT SET (0xFFC00000 / 0xFFF8000000000000)
 * for invalid operations and otherwise returns the first NaN source operand,
 * quieted.  totalOrder makes the sign of a NaN observable (-NaN sorts below
 * everything), so the rule is written out explicitly here (ora_nanfix_*) and
 * in the CUDA kernels rather than left to the code generator.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this file's shared object.
 *
 * Build: gcc -O3 -march=native -ffp-contract=off -fno-fast-math -shared -fPIC
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

enum {
  ORA_BOOL = 0, ORA_I8 = 1, ORA_I16 = 2, ORA_I32 = 3, ORA_I64 = 4,
  ORA_U8 = 5, ORA_U16 = 6, ORA_U32 = 7, ORA_U64 = 8, ORA_F32 = 9, ORA_F64 = 10,
  ORA_UTF8 = 11
};
enum { ORA_ADD = 0, ORA_MUL = 1, ORA_DIV = 2, ORA_REM = 3, ORA_SUB = 4 };
enum { ORA_EQ = 0, ORA_NE = 1, ORA_LT = 2, ORA_LE = 3, ORA_GT = 4, ORA_GE = 5 };
enum { ORA_OK = 0, ORA_ERR_OVERFLOW = 1, ORA_ERR_DIVZERO = 2, ORA_ERR_BADARG = 3 };

static inline int getbit(const uint8_t *b, int64_t i) { return (b[i >> 3] >> (i & 7)) & 1; }
static inline void setbit(uint8_t *b, int64_t i) { b[i >> 3] |= (uint8_t)(1u << (i & 7)); }

/* ------------------------------------------------------------------ */
/* NaN rule (see header)                                               */
/* ------------------------------------------------------------------ */
static inline float f32_from_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f32_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline double f64_from_bits(uint64_t u) { double f; memcpy(&f, &u, 8); return f; }
static inline uint64_t f64_bits(double f) { uint64_t u; memcpy(&u, &f, 8); return u; }

static inline float ora_nanfix_f32(float r, float a, float b) {
  if (r != r) {
    if (a != a) return f32_from_bits(f32_bits(a) | 0x00400000u);
    if (b != b) return f32_from_bits(f32_bits(b) | 0x00400000u);
    return f32_from_bits(0xFFC00000u);
  }
  return r;
}
static inline double ora_nanfix_f64(double r, double a, double b) {
  if (r != r) {
    if (a != a) return f64_from_bits(f64_bits(a) | 0x0008000000000000ull);
    if (b != b) return f64_from_bits(f64_bits(b) | 0x0008000000000000ull);
    return f64_from_bits(0xFFF8000000000000ull);
  }
  return r;
}

/* ------------------------------------------------------------------ */
/* bitmaps                                                             */
/* ------------------------------------------------------------------ */
int64_t ora_popcount(int64_t n, const uint8_t *bits) {
  int64_t c = 0, i = 0;
  int64_t nb = n >> 3;
  for (; i + 8 <= nb; i += 8) { uint64_t w; memcpy(&w, bits + i, 8); c += __builtin_popcountll(w); }
  for (; i < nb; i++) c += __builtin_popcount(bits[i]);
  int rem = (int)(n & 7);
  if (rem) c += __builtin_popcount(bits[nb] & ((1u << rem) - 1u));
  return c;
}

/* arrow-arith boolean::and / boolean::or on the VALUE bitmaps
 * (validity is combined separately with ora_bitmap_and). */
void ora_bitmap_and(int64_t n, const uint8_t *a, const uint8_t *b, uint8_t *out) {
  int64_t nb = (n + 7) >> 3;
  for (int64_t i = 0; i < nb; i++) out[i] = a[i] & b[i];
}
void ora_bitmap_or(int64_t n, const uint8_t *a, const uint8_t *b, uint8_t *out) {
  int64_t nb = (n + 7) >> 3;
  for (int64_t i = 0; i < nb; i++) out[i] = a[i] | b[i];
}

/* ------------------------------------------------------------------ */
/* arithmetic: arrow-arith numeric::{add,mul,div,rem} (+ sub)          */
/*  sa/sb: stride 0 = scalar (Datum is_scalar), 1 = array              */
/*  valid: union validity bitmap of both operands or NULL              */
/* ------------------------------------------------------------------ */
#define INT_LOOP(T, BODY)                                                     \
  for (int64_t i = 0; i < n; i++) {                                            \
    if (valid && !getbit(valid, i)) { out[i] = 0; continue; }                  \
    T x = a[i * sa], y = b[i * sb], r = 0;                                     \
    BODY;                                                                      \
    out[i] = r;                                                                \
  }

#define DEF_INT_ARITH(T, NAME, SIGNED, TMIN)                                   \
  static int arith_##NAME(int op, int64_t n, const T *a, int sa, const T *b,   \
                          int sb, const uint8_t *valid, T *out,                \
                          int64_t *err_row) {                                  \
    switch (op) {                                                              \
    case ORA_ADD:                                                              \
      INT_LOOP(T, if (__builtin_add_overflow(x, y, &r)) {                      \
        *err_row = i; return ORA_ERR_OVERFLOW; })                              \
      return ORA_OK;                                                           \
    case ORA_SUB:                                                              \
      INT_LOOP(T, if (__builtin_sub_overflow(x, y, &r)) {                      \
        *err_row = i; return ORA_ERR_OVERFLOW; })                              \
      return ORA_OK;                                                           \
    case ORA_MUL:                                                              \
      INT_LOOP(T, if (__builtin_mul_overflow(x, y, &r)) {                      \
        *err_row = i; return ORA_ERR_OVERFLOW; })                              \
      return ORA_OK;                                                           \
    case ORA_DIV:                                                              \
      INT_LOOP(T, if (y == 0) { *err_row = i; return ORA_ERR_DIVZERO; }        \
               if (SIGNED && x == (T)(TMIN) && y == (T)-1) {                   \
                 *err_row = i; return ORA_ERR_OVERFLOW; }                      \
               r = (T)(x / y);)                                                \
      return ORA_OK;                                                           \
    case ORA_REM:                                                              \
      INT_LOOP(T, if (y == 0) { *err_row = i; return ORA_ERR_DIVZERO; }        \
               if (SIGNED && x == (T)(TMIN) && y == (T)-1) {                   \
                 *err_row = i; return ORA_ERR_OVERFLOW; }                      \
               r = (T)(x % y);)                                                \
      return ORA_OK;                                                           \
    }                                                                          \
    return ORA_ERR_BADARG;                                                     \
  }

DEF_INT_ARITH(int8_t, i8, 1, INT8_MIN)
DEF_INT_ARITH(int16_t, i16, 1, INT16_MIN)
DEF_INT_ARITH(int32_t, i32, 1, INT32_MIN)
DEF_INT_ARITH(int64_t, i64, 1, INT64_MIN)
DEF_INT_ARITH(uint8_t, u8, 0, 0)
DEF_INT_ARITH(uint16_t, u16, 0, 0)
DEF_INT_ARITH(uint32_t, u32, 0, 0)
DEF_INT_ARITH(uint64_t, u64, 0, 0)

#define DEF_FLT_ARITH(T, NAME, FMOD, FIX)                                      \
  static int arith_##NAME(int op, int64_t n, const T *a, int sa, const T *b,   \
                          int sb, T *out) {                                    \
    switch (op) {                                                              \
    case ORA_ADD:                                                              \
      for (int64_t i = 0; i < n; i++) { T x = a[i * sa], y = b[i * sb];        \
        out[i] = FIX(x + y, x, y); }                                           \
      return ORA_OK;                                                           \
    case ORA_SUB:                                                              \
      for (int64_t i = 0; i < n; i++) { T x = a[i * sa], y = b[i * sb];        \
        out[i] = FIX(x - y, x, y); }                                           \
      return ORA_OK;                                                           \
    case ORA_MUL:                                                              \
      for (int64_t i = 0; i < n; i++) { T x = a[i * sa], y = b[i * sb];        \
        out[i] = FIX(x * y, x, y); }                                           \
      return ORA_OK;                                                           \
    case ORA_DIV:                                                              \
      for (int64_t i = 0; i < n; i++) { T x = a[i * sa], y = b[i * sb];        \
        out[i] = FIX(x / y, x, y); }                                           \
      return ORA_OK;                                                           \
    case ORA_REM:                                                              \
      for (int64_t i = 0; i < n; i++) { T x = a[i * sa], y = b[i * sb];        \
        out[i] = FIX(FMOD(x, y), x, y); }                                      \
      return ORA_OK;                                                           \
    }                                                                          \
    return ORA_ERR_BADARG;                                                     \
  }

DEF_FLT_ARITH(float, f32, fmodf, ora_nanfix_f32)
DEF_FLT_ARITH(double, f64, fmod, ora_nanfix_f64)

/* out has n elements unless both operands are scalars (then n == 1). */
int ora_arith(int op, int type, int64_t n, const void *a, int a_is_scalar,
              const void *b, int b_is_scalar, const uint8_t *valid, void *out,
              int64_t *err_row) {
  int sa = a_is_scalar ? 0 : 1, sb = b_is_scalar ? 0 : 1;
  *err_row = -1;
  switch (type) {
  case ORA_I8: return arith_i8(op, n, a, sa, b, sb, valid, out, err_row);
  case ORA_I16: return arith_i16(op, n, a, sa, b, sb, valid, out, err_row);
  case ORA_I32: return arith_i32(op, n, a, sa, b, sb, valid, out, err_row);
  case ORA_I64: return arith_i64(op, n, a, sa, b, sb, valid, out, err_row);
  case ORA_U8: return arith_u8(op, n, a, sa, b, sb, valid, out, err_row);
  case ORA_U16: return arith_u16(op, n, a, sa, b, sb, valid, out, err_row);
  case ORA_U32: return arith_u32(op, n, a, sa, b, sb, valid, out, err_row);
  case ORA_U64: return arith_u64(op, n, a, sa, b, sb, valid, out, err_row);
  case ORA_F32: return arith_f32(op, n, a, sa, b, sb, out);
  case ORA_F64: return arith_f64(op, n, a, sa, b, sb, out);
  }
  return ORA_ERR_BADARG;
}

/* ------------------------------------------------------------------ */
/* comparisons: arrow-ord cmp::{eq,neq,lt,lt_eq,gt,gt_eq}              */
/* ------------------------------------------------------------------ */
static inline int32_t total_key_f32(float f) {
  int32_t k = (int32_t)f32_bits(f);
  k ^= (int32_t)(((uint32_t)(k >> 31)) >> 1);
  return k;
}
static inline int64_t total_key_f64(double f) {
  int64_t k = (int64_t)f64_bits(f);
  k ^= (int64_t)(((uint64_t)(k >> 63)) >> 1);
  return k;
}

/* is_eq / is_lt as arrow-rs ArrowNativeTypeOp defines them */
#define EQ_INT(x, y) ((x) == (y))
#define LT_INT(x, y) ((x) < (y))
#define EQ_F32(x, y) (f32_bits(x) == f32_bits(y))
#define LT_F32(x, y) (total_key_f32(x) < total_key_f32(y))
#define EQ_F64(x, y) (f64_bits(x) == f64_bits(y))
#define LT_F64(x, y) (total_key_f64(x) < total_key_f64(y))

#define CMP_LOOP(T, EXPR)                                                      \
  for (int64_t i = 0; i < n; i++) { T x = a[i * sa], y = b[i * sb];            \
    if (EXPR) setbit(out, i); }

#define DEF_CMP(T, NAME, EQ, LT)                                               \
  static void cmp_##NAME(int op, int64_t n, const T *a, int sa, const T *b,    \
                         int sb, uint8_t *out) {                               \
    switch (op) {                                                              \
    case ORA_EQ: CMP_LOOP(T, EQ(x, y)) break;                                  \
    case ORA_NE: CMP_LOOP(T, !EQ(x, y)) break;                                 \
    case ORA_LT: CMP_LOOP(T, LT(x, y)) break;                                  \
    case ORA_LE: CMP_LOOP(T, !LT(y, x)) break;                                 \
    case ORA_GT: CMP_LOOP(T, LT(y, x)) break;                                  \
    case ORA_GE: CMP_LOOP(T, !LT(x, y)) break;                                 \
    }                                                                          \
  }

DEF_CMP(int8_t, i8, EQ_INT, LT_INT)
DEF_CMP(int16_t, i16, EQ_INT, LT_INT)
DEF_CMP(int32_t, i32, EQ_INT, LT_INT)
DEF_CMP(int64_t, i64, EQ_INT, LT_INT)
DEF_CMP(uint8_t, u8, EQ_INT, LT_INT)
DEF_CMP(uint16_t, u16, EQ_INT, LT_INT)
DEF_CMP(uint32_t, u32, EQ_INT, LT_INT)
DEF_CMP(uint64_t, u64, EQ_INT, LT_INT)
DEF_CMP(float, f32, EQ_F32, LT_F32)
DEF_CMP(double, f64, EQ_F64, LT_F64)

/* out: value bitmap of ceil(n/8) bytes, zeroed here. */
int ora_cmp(int op, int type, int64_t n, const void *a, int a_is_scalar,
            const void *b, int b_is_scalar, uint8_t *out) {
  int sa = a_is_scalar ? 0 : 1, sb = b_is_scalar ? 0 : 1;
  memset(out, 0, (size_t)((n + 7) >> 3));
  switch (type) {
  case ORA_I8: cmp_i8(op, n, a, sa, b, sb, out); return ORA_OK;
  case ORA_I16: cmp_i16(op, n, a, sa, b, sb, out); return ORA_OK;
  case ORA_I32: cmp_i32(op, n, a, sa, b, sb, out); return ORA_OK;
  case ORA_I64: cmp_i64(op, n, a, sa, b, sb, out); return ORA_OK;
  case ORA_U8: cmp_u8(op, n, a, sa, b, sb, out); return ORA_OK;
  case ORA_U16: cmp_u16(op, n, a, sa, b, sb, out); return ORA_OK;
  case ORA_U32: cmp_u32(op, n, a, sa, b, sb, out); return ORA_OK;
  case ORA_U64: cmp_u64(op, n, a, sa, b, sb, out); return ORA_OK;
  case ORA_F32: cmp_f32(op, n, a, sa, b, sb, out); return ORA_OK;
  case ORA_F64: cmp_f64(op, n, a, sa, b, sb, out); return ORA_OK;
  }
  return ORA_ERR_BADARG;
}

/* Boolean operands are bit-packed; false < true. */
int ora_cmp_bool(int op, int64_t n, const uint8_t *a, int a_is_scalar,
                 const uint8_t *b, int b_is_scalar, uint8_t *out) {
  memset(out, 0, (size_t)((n + 7) >> 3));
  for (int64_t i = 0; i < n; i++) {
    int x = getbit(a, a_is_scalar ? 0 : i), y = getbit(b, b_is_scalar ? 0 : i), r = 0;
    switch (op) {
    case ORA_EQ: r = x == y; break;
    case ORA_NE: r = x != y; break;
    case ORA_LT: r = x < y; break;
    case ORA_LE: r = x <= y; break;
    case ORA_GT: r = x > y; break;
    case ORA_GE: r = x >= y; break;
    }
    if (r) setbit(out, i);
  }
  return ORA_OK;
}

/* Utf8: bytewise lexicographic order (arrow-ord compares &[u8]). */
static inline int bytes_cmp(const uint8_t *p, int64_t lp, const uint8_t *q, int64_t lq) {
  int64_t m = lp < lq ? lp : lq;
  int c = m ? memcmp(p, q, (size_t)m) : 0;
  if (c) return c < 0 ? -1 : 1;
  return lp < lq ? -1 : (lp > lq ? 1 : 0);
}
int ora_cmp_utf8(int op, int64_t n, const int32_t *ao, const uint8_t *ad,
                 int a_is_scalar, const int32_t *bo, const uint8_t *bd,
                 int b_is_scalar, uint8_t *out) {
  memset(out, 0, (size_t)((n + 7) >> 3));
  for (int64_t i = 0; i < n; i++) {
    int64_t ia = a_is_scalar ? 0 : i, ib = b_is_scalar ? 0 : i;
    int c = bytes_cmp(ad + ao[ia], ao[ia + 1] - ao[ia], bd + bo[ib], bo[ib + 1] - bo[ib]);
    int r = 0;
    switch (op) {
    case ORA_EQ: r = c == 0; break;
    case ORA_NE: r = c != 0; break;
    case ORA_LT: r = c < 0; break;
    case ORA_LE: r = c <= 0; break;
    case ORA_GT: r = c > 0; break;
    case ORA_GE: r = c >= 0; break;
    }
    if (r) setbit(out, i);
  }
  return ORA_OK;
}

/* ------------------------------------------------------------------ */
/* casts: arrow-cast cast() with default CastOptions{safe:true}        */
/* ------------------------------------------------------------------ */
static inline double cvt_f32_f64(float x) {
  if (x != x) { /* x86 cvtss2sd: keep sign and payload, quiet */
    uint32_t u = f32_bits(x);
    uint64_t v = ((uint64_t)(u & 0x80000000u) << 32) | 0x7FF8000000000000ull |
                 ((uint64_t)(u & 0x007FFFFFu) << 29);
    return f64_from_bits(v);
  }
  return (double)x;
}

#define CAST_LOOP(FROM, TO) { const FROM *s = in; TO *d = out;                 \
    for (int64_t i = 0; i < n; i++) { d[i] = (TO)s[i]; }                       \
    return ORA_OK; }

#define CAST_FROM(FROM)                                                        \
  switch (to) {                                                                \
  case ORA_I8: CAST_LOOP(FROM, int8_t)                                         \
  case ORA_I16: CAST_LOOP(FROM, int16_t)                                       \
  case ORA_I32: CAST_LOOP(FROM, int32_t)                                       \
  case ORA_I64: CAST_LOOP(FROM, int64_t)                                       \
  case ORA_U8: CAST_LOOP(FROM, uint8_t)                                        \
  case ORA_U16: CAST_LOOP(FROM, uint16_t)                                      \
  case ORA_U32: CAST_LOOP(FROM, uint32_t)                                      \
  case ORA_U64: CAST_LOOP(FROM, uint64_t)                                      \
  case ORA_F32: CAST_LOOP(FROM, float)                                         \
  case ORA_F64: CAST_LOOP(FROM, double)                                        \
  }                                                                            \
  return ORA_ERR_BADARG;

/* Only value-preserving (widening) int casts, int->float (RNE, what Rust
 * `as` and C both do) and f32->f64 are reachable from cast_to_common_type
 * (compute_value.rs:350-461); narrowing casts are rejected here. */
int ora_cast(int from, int to, int64_t n, const void *in, void *out) {
  static const int width[] = {0, 1, 2, 4, 8, 1, 2, 4, 8, 4, 8};
  if (from < ORA_I8 || from > ORA_F64 || to < ORA_I8 || to > ORA_F64) return ORA_ERR_BADARG;
  int from_f = from >= ORA_F32, to_f = to >= ORA_F32;
  if (from_f && !to_f) return ORA_ERR_BADARG;
  if (from_f && to_f && width[to] < width[from]) return ORA_ERR_BADARG;
  if (!from_f && !to_f && width[to] < width[from]) return ORA_ERR_BADARG;
  if (from == ORA_F32 && to == ORA_F64) {
    const float *s = in; double *d = out;
    for (int64_t i = 0; i < n; i++) d[i] = cvt_f32_f64(s[i]);
    return ORA_OK;
  }
  switch (from) {
  case ORA_I8: CAST_FROM(int8_t)
  case ORA_I16: CAST_FROM(int16_t)
  case ORA_I32: CAST_FROM(int32_t)
  case ORA_I64: CAST_FROM(int64_t)
  case ORA_U8: CAST_FROM(uint8_t)
  case ORA_U16: CAST_FROM(uint16_t)
  case ORA_U32: CAST_FROM(uint32_t)
  case ORA_U64: CAST_FROM(uint64_t)
  case ORA_F32: CAST_FROM(float)
  case ORA_F64: CAST_FROM(double)
  }
  return ORA_ERR_BADARG;
}

/* numeric -> Boolean: value != 0 (NaN -> true, -0.0 -> false) */
#define TOBOOL_LOOP(T) { const T *s = in;                                      \
    for (int64_t i = 0; i < n; i++) { if (s[i] != (T)0) setbit(out, i); }      \
    return ORA_OK; }
int ora_cast_to_bool(int from, int64_t n, const void *in, uint8_t *out) {
  memset(out, 0, (size_t)((n + 7) >> 3));
  switch (from) {
  case ORA_I8: TOBOOL_LOOP(int8_t)
  case ORA_I16: TOBOOL_LOOP(int16_t)
  case ORA_I32: TOBOOL_LOOP(int32_t)
  case ORA_I64: TOBOOL_LOOP(int64_t)
  case ORA_U8: TOBOOL_LOOP(uint8_t)
  case ORA_U16: TOBOOL_LOOP(uint16_t)
  case ORA_U32: TOBOOL_LOOP(uint32_t)
  case ORA_U64: TOBOOL_LOOP(uint64_t)
  case ORA_F32: TOBOOL_LOOP(float)
  case ORA_F64: TOBOOL_LOOP(double)
  }
  return ORA_ERR_BADARG;
}

/* ------------------------------------------------------------------ */
/* filter: arrow-select filter::filter_record_batch, per column        */
/*  sel = predicate.values & predicate.validity (NULL predicate = drop) */
/* ------------------------------------------------------------------ */
#define FOR_EACH_SET_BIT(sel, n, IDX, BODY)                                    \
  {                                                                            \
    int64_t nwords_ = (n + 63) >> 6;                                           \
    for (int64_t w_ = 0; w_ < nwords_; w_++) {                                 \
      uint64_t bits_ = 0;                                                      \
      int64_t nb_ = ((n + 7) >> 3) - w_ * 8;                                   \
      memcpy(&bits_, sel + w_ * 8, (size_t)(nb_ >= 8 ? 8 : nb_));              \
      if (w_ == nwords_ - 1 && (n & 63)) bits_ &= (~0ull) >> (64 - (n & 63));  \
      while (bits_) {                                                          \
        int64_t IDX = w_ * 64 + __builtin_ctzll(bits_);                        \
        bits_ &= bits_ - 1;                                                    \
        BODY;                                                                  \
      }                                                                        \
    }                                                                          \
  }

int64_t ora_filter_fixed(int width, int64_t n, const uint8_t *sel, const void *in, void *out) {
  int64_t k = 0;
  const uint8_t *s = in; uint8_t *d = out;
  switch (width) {
  case 1: FOR_EACH_SET_BIT(sel, n, i, d[k++] = s[i]) break;
  case 2: FOR_EACH_SET_BIT(sel, n, i, ((uint16_t *)d)[k++] = ((const uint16_t *)s)[i]) break;
  case 4: FOR_EACH_SET_BIT(sel, n, i, ((uint32_t *)d)[k++] = ((const uint32_t *)s)[i]) break;
  case 8: FOR_EACH_SET_BIT(sel, n, i, ((uint64_t *)d)[k++] = ((const uint64_t *)s)[i]) break;
  default: FOR_EACH_SET_BIT(sel, n, i, { memcpy(d + k * width, s + i * width, (size_t)width); k++; }) break;
  }
  return k;
}

/* bit-gather (validity bitmaps and Boolean value buffers);
 * returns the number of SET bits written; out zeroed here (capacity n bits). */
int64_t ora_filter_bits(int64_t n, const uint8_t *sel, const uint8_t *in, uint8_t *out) {
  int64_t k = 0, set = 0;
  memset(out, 0, (size_t)((n + 7) >> 3));
  FOR_EACH_SET_BIT(sel, n, i, { if (getbit(in, i)) { setbit(out, k); set++; } k++; })
  return set;
}

/* Utf8: new offsets start at 0; returns total value bytes written. */
int64_t ora_filter_utf8(int64_t n, const uint8_t *sel, const int32_t *in_off,
                        const uint8_t *in_data, int32_t *out_off, uint8_t *out_data) {
  int64_t k = 0; int64_t pos = 0;
  out_off[0] = 0;
  FOR_EACH_SET_BIT(sel, n, i, {
    int32_t len = in_off[i + 1] - in_off[i];
    memcpy(out_data + pos, in_data + in_off[i], (size_t)len);
    pos += len; out_off[++k] = (int32_t)pos; })
  return pos;
}
