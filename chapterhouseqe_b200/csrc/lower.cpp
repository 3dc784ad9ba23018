// Lowering: sqlparser 0.52 `Expr` / `SelectItem` (serde JSON) + Arrow schema -> register bytecode.
//
// Restates the *decisions* the reference makes while walking the tree
// (record_utils/compute_value.rs:57-344): node coverage, literal typing (:219-265), column
// resolution (:266-337), the coercion lattice (:350-431, :433-461), scalar tracking
// (ArrayDatum, :34-55) and the naming / nullability rules of record_projection.rs:16-76 --
// but instead of materialising an array per node it emits accumulator-machine code that
// one fused kernel interprets.  Scalar-only subtrees are folded here with the same
// arithmetic the kernels use (checked integers, IEEE floats, totalOrder compares).
#include <atomic>
#include <cerrno>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <sstream>

#include "json.hpp"
#include "program.hpp"

namespace chdb {

// ------------------------------------------------------------------------------------------
// type tables
// ------------------------------------------------------------------------------------------
const char* type_arrow_name(uint8_t t) {
  static const char* n[] = {"Boolean", "Int8", "Int16", "Int32", "Int64", "UInt8", "UInt16",
                            "UInt32", "UInt64", "Float32", "Float64", "Utf8", "Opaque"};
  return t <= T_OPAQUE ? n[t] : "?";
}
const char* type_format(uint8_t t) {
  static const char* f[] = {"b", "c", "s", "i", "l", "C", "S", "I", "L", "f", "g", "u"};
  return t <= T_UTF8 ? f[t] : "?";
}

static InputColumn column_from_format(const char* fmt) {
  InputColumn c;
  c.format = fmt ? fmt : "";
  const std::string& f = c.format;
  auto fixed = [&](TypeId t, int w) { c.type = t; c.width = w; };
  if (f == "b") c.type = T_BOOL;
  else if (f == "c") fixed(T_I8, 1);
  else if (f == "C") fixed(T_U8, 1);
  else if (f == "s") fixed(T_I16, 2);
  else if (f == "S") fixed(T_U16, 2);
  else if (f == "i") fixed(T_I32, 4);
  else if (f == "I") fixed(T_U32, 4);
  else if (f == "l") fixed(T_I64, 8);
  else if (f == "L") fixed(T_U64, 8);
  else if (f == "f") fixed(T_F32, 4);
  else if (f == "g") fixed(T_F64, 8);
  else if (f == "u") c.type = T_UTF8;
  // fixed-width types this library can carry through a filter but does not compute on
  else if (f == "e") fixed(T_OPAQUE, 2);
  else if (f == "tdD" || f == "tts" || f == "ttm") fixed(T_OPAQUE, 4);
  else if (f == "tdm" || f == "ttu" || f == "ttn" || f.rfind("ts", 0) == 0 || f.rfind("tD", 0) == 0) fixed(T_OPAQUE, 8);
  else if (f.rfind("d:", 0) == 0 && f.find(',', 2) != std::string::npos && f.find(',', f.find(',', 2) + 1) == std::string::npos)
    fixed(T_OPAQUE, 16);  // decimal128
  else { c.type = T_OPAQUE; c.supported = false; }
  return c;
}

std::vector<InputColumn> parse_schema(const ::ArrowSchema* schema) {
  if (!schema || !schema->format) throw Error(CHDB_ERR_INVALID_ARGUMENT, "null ArrowSchema");
  if (std::strcmp(schema->format, "+s") != 0)
    throw Error(CHDB_ERR_INVALID_ARGUMENT, "record batches must be exported as a struct array (format \"+s\")");
  std::vector<InputColumn> cols;
  for (int64_t i = 0; i < schema->n_children; i++) {
    const ::ArrowSchema* ch = schema->children[i];
    InputColumn c = column_from_format(ch->format);
    c.name = ch->name ? ch->name : "";
    c.flags = ch->flags;
    if (ch->dictionary) c.supported = false;
    cols.push_back(std::move(c));
  }
  return cols;
}

int Program::slot_for(int col) {
  for (size_t s = 0; s < slot_to_col.size(); s++)
    if (slot_to_col[s] == col) return (int)s;
  if ((int)slot_to_col.size() >= kMaxInCols)
    throw Error(CHDB_ERR_NOT_IMPLEMENTED, "more than 24 input columns referenced by one program");
  if (!schema[col].supported)
    throw Error(CHDB_ERR_NOT_IMPLEMENTED, "column '" + schema[col].name + "' has an Arrow layout (" + schema[col].format +
                                              ") this library cannot move");
  slot_to_col.push_back(col);
  return (int)slot_to_col.size() - 1;
}

// ------------------------------------------------------------------------------------------
// host-side scalar arithmetic (identical semantics to the kernels; used for constant folding)
// ------------------------------------------------------------------------------------------
static inline float f32_of(uint64_t v) { uint32_t u = (uint32_t)v; float f; std::memcpy(&f, &u, 4); return f; }
static inline uint64_t bits_of(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline double f64_of(uint64_t v) { double f; std::memcpy(&f, &v, 8); return f; }
static inline uint64_t bits_of(double f) { uint64_t u; std::memcpy(&u, &f, 8); return u; }

static float nanfix(float r, float a, float b) {
  if (r != r) {
    if (a != a) return f32_of(bits_of(a) | 0x00400000u);
    if (b != b) return f32_of(bits_of(b) | 0x00400000u);
    return f32_of(0xFFC00000u);
  }
  return r;
}
static double nanfix(double r, double a, double b) {
  if (r != r) {
    if (a != a) return f64_of(bits_of(a) | 0x0008000000000000ull);
    if (b != b) return f64_of(bits_of(b) | 0x0008000000000000ull);
    return f64_of(0xFFF8000000000000ull);
  }
  return r;
}

struct IntRange { __int128 lo, hi; };
static IntRange int_range(uint8_t t) {
  switch (t) {
    case T_I8: return {INT8_MIN, INT8_MAX};
    case T_I16: return {INT16_MIN, INT16_MAX};
    case T_I32: return {INT32_MIN, INT32_MAX};
    case T_I64: return {INT64_MIN, INT64_MAX};
    case T_U8: return {0, UINT8_MAX};
    case T_U16: return {0, UINT16_MAX};
    case T_U32: return {0, UINT32_MAX};
    default: return {0, (__int128)UINT64_MAX};
  }
}
static bool is_signed_int(uint8_t t) { return t >= T_I8 && t <= T_I64; }
static bool is_unsigned_int(uint8_t t) { return t >= T_U8 && t <= T_U64; }
static bool is_int(uint8_t t) { return t >= T_I8 && t <= T_U64; }
static bool is_float(uint8_t t) { return t == T_F32 || t == T_F64; }
static __int128 int_of(uint8_t t, uint64_t v) { return is_signed_int(t) ? (__int128)(int64_t)v : (__int128)v; }

static uint64_t fold_arith(uint8_t op, uint8_t t, uint64_t a, uint64_t b) {
  const char* sym = op == OP_ADD ? "+" : op == OP_MUL ? "*" : op == OP_DIV ? "/" : op == OP_REM ? "%" : "-";
  if (is_int(t)) {
    __int128 x = int_of(t, a), y = int_of(t, b), r = 0;
    IntRange rg = int_range(t);
    if ((op == OP_DIV || op == OP_REM) && y == 0) throw Error(CHDB_ERR_DIVIDE_BY_ZERO, "Divide by zero error");
    switch (op) {
      case OP_ADD: r = x + y; break;
      case OP_SUB: r = x - y; break;
      case OP_MUL: r = x * y; break;
      case OP_DIV: r = x / y; break;
      case OP_REM: r = (x == rg.lo && y == -1 && is_signed_int(t)) ? rg.hi + 1 : x % y; break;
    }
    if (r < rg.lo || r > rg.hi)
      throw Error(CHDB_ERR_ARITHMETIC_OVERFLOW, std::string("Arithmetic overflow: Overflow happened on: ") +
                                                    std::to_string((long long)x) + " " + sym + " " + std::to_string((long long)y));
    return (uint64_t)(int64_t)r;
  }
  if (t == T_F32) {
    float x = f32_of(a), y = f32_of(b), r = 0;
    switch (op) {
      case OP_ADD: r = x + y; break;
      case OP_SUB: r = x - y; break;
      case OP_MUL: r = x * y; break;
      case OP_DIV: r = x / y; break;
      case OP_REM: r = std::fmod(x, y); break;
    }
    return bits_of(nanfix(r, x, y));
  }
  if (t == T_F64) {
    double x = f64_of(a), y = f64_of(b), r = 0;
    switch (op) {
      case OP_ADD: r = x + y; break;
      case OP_SUB: r = x - y; break;
      case OP_MUL: r = x * y; break;
      case OP_DIV: r = x / y; break;
      case OP_REM: r = std::fmod(x, y); break;
    }
    return bits_of(nanfix(r, x, y));
  }
  throw Error(CHDB_ERR_INVALID_ARGUMENT, std::string("Invalid arithmetic operation: ") + type_arrow_name(t) + " " + sym + " " + type_arrow_name(t));
}

static int64_t total_key32(uint64_t v) { int32_t k = (int32_t)(uint32_t)v; k ^= (int32_t)(((uint32_t)(k >> 31)) >> 1); return k; }
static int64_t total_key64(uint64_t v) { int64_t k = (int64_t)v; k ^= (int64_t)(((uint64_t)(k >> 63)) >> 1); return k; }

static bool apply_cmp(uint8_t kind, bool lt, bool eq, bool gt) {
  switch (kind) {
    case CMP_EQ: return eq;
    case CMP_NE: return !eq;
    case CMP_LT: return lt;
    case CMP_LE: return !gt;
    case CMP_GT: return gt;
    default: return !lt;
  }
}
static bool fold_cmp(uint8_t kind, uint8_t t, uint64_t a, uint64_t b) {
  bool lt, eq, gt;
  if (t == T_F32) { int64_t x = total_key32(a), y = total_key32(b); lt = x < y; gt = x > y; eq = (uint32_t)a == (uint32_t)b; }
  else if (t == T_F64) { int64_t x = total_key64(a), y = total_key64(b); lt = x < y; gt = x > y; eq = a == b; }
  else { __int128 x = int_of(t, a), y = int_of(t, b); lt = x < y; gt = x > y; eq = x == y; }
  return apply_cmp(kind, lt, eq, gt);
}
static bool fold_cmp_str(uint8_t kind, const std::string& a, const std::string& b) {
  int c = a.compare(b);  // char_traits<char>::compare == memcmp: bytewise, then length
  size_t m = std::min(a.size(), b.size());
  int mc = m ? std::memcmp(a.data(), b.data(), m) : 0;
  c = mc ? mc : (a.size() < b.size() ? -1 : a.size() > b.size() ? 1 : 0);
  return apply_cmp(kind, c < 0, c == 0, c > 0);
}

// arrow cast on the lattice: widening ints are value-preserving; int -> float is RNE.
static uint64_t fold_cast(uint8_t from, uint8_t to, uint64_t v) {
  if (from == to) return v;
  if (is_int(from) && is_int(to)) return v;  // canonical container already holds the value
  if (is_int(from) && to == T_F32) return bits_of(is_signed_int(from) ? (float)(int64_t)v : (float)v);
  if (is_int(from) && to == T_F64) return bits_of(is_signed_int(from) ? (double)(int64_t)v : (double)v);
  if (from == T_F32 && to == T_F64) {
    float x = f32_of(v);
    if (x != x) {
      uint32_t u = (uint32_t)v;
      return ((uint64_t)(u & 0x80000000u) << 32) | 0x7FF8000000000000ull | ((uint64_t)(u & 0x007FFFFFu) << 29);
    }
    return bits_of((double)x);
  }
  throw Error(CHDB_ERR_NOT_IMPLEMENTED, std::string("cast ") + type_arrow_name(from) + " -> " + type_arrow_name(to));
}
static bool fold_tobool(uint8_t t, uint64_t v) {
  if (t == T_BOOL) return v != 0;
  if (t == T_F32) return f32_of(v) != 0.0f;
  if (t == T_F64) return f64_of(v) != 0.0;
  return v != 0;
}

// ------------------------------------------------------------------------------------------
// typed IR
// ------------------------------------------------------------------------------------------
struct TNode {
  enum K { CONST, COL, CAST, TOBOOL, ARITH, CMP, AND, OR, NEG, NOT, ISNULL } k = CONST;
  uint8_t type = T_NONE;   // result type
  bool is_scalar = false;  // ArrayDatum.is_scalar
  bool len1 = false;       // array of length 1 whatever the batch length
  uint64_t imm = 0;
  std::string str;
  int col = -1;
  uint8_t op = 0;          // ARITH: Opcode; CMP: CmpKind; AND / OR: 1 = Kleene; ISNULL: 1 = IS NOT NULL
  uint8_t order = 0;
  std::unique_ptr<TNode> l, r;
};
using TP = std::unique_ptr<TNode>;

// SQL nodes beyond what the reference's compute_value accepts; off unless the host asks for them
// (chdb_set_sql_extensions), so by default every such node returns the reference's own error.
std::atomic<uint32_t> g_sql_extensions{0};
constexpr uint32_t kExtOperators = 1;   // binary -, unary - / +, NOT, IS [NOT] NULL
constexpr uint32_t kExtKleene = 2;      // AND / OR in SQL three-valued logic (arrow and_kleene / or_kleene)

struct Lowering {
  Program& prog;
  const uint32_t ext = g_sql_extensions.load();
  std::vector<std::vector<std::string>> aliases;
  bool aliases_given = false;
  int next_order = 0;
  int depth = 0;

  explicit Lowering(Program& p) : prog(p) {}

  // ---- column resolution (compute_value.rs:266-337) ----
  int column_by_name(const std::string& name) {
    for (size_t i = 0; i < prog.schema.size(); i++)
      if (prog.schema[i].name == name) return (int)i;  // arrow: first match
    return -1;
  }
  TP make_col(int idx) {
    const InputColumn& c = prog.schema[idx];
    if (c.type == T_OPAQUE)
      throw Error(CHDB_ERR_NOT_IMPLEMENTED, "expressions over Arrow type '" + c.format + "' (column '" + c.name + "')");
    TP n(new TNode);
    n->k = TNode::COL;
    n->type = c.type;
    n->col = idx;
    return n;
  }
  static std::string ident_value(const Json& id) {
    const Json* v = id.get("value");
    if (!v || v->kind != Json::String) throw Error(CHDB_ERR_BAD_JSON, "Ident without a string 'value'");
    return v->str;
  }

  // ---- literal typing (compute_value.rs:219-265) ----
  static bool rust_float_syntax(const std::string& s) {
    size_t i = 0, n = s.size();
    if (i < n && (s[i] == '+' || s[i] == '-')) i++;
    size_t d0 = i;
    while (i < n && isdigit((unsigned char)s[i])) i++;
    size_t int_digits = i - d0, frac_digits = 0;
    if (i < n && s[i] == '.') {
      i++;
      size_t f0 = i;
      while (i < n && isdigit((unsigned char)s[i])) i++;
      frac_digits = i - f0;
    }
    if (int_digits + frac_digits == 0) return false;
    if (i < n && (s[i] == 'e' || s[i] == 'E')) {
      i++;
      if (i < n && (s[i] == '+' || s[i] == '-')) i++;
      size_t e0 = i;
      while (i < n && isdigit((unsigned char)s[i])) i++;
      if (i == e0) return false;
    }
    return i == n;
  }
  static bool rust_int_syntax(const std::string& s) {
    size_t i = 0, n = s.size();
    if (i < n && (s[i] == '+' || s[i] == '-')) i++;
    if (i == n) return false;
    for (; i < n; i++)
      if (!isdigit((unsigned char)s[i])) return false;
    return true;
  }
  TP make_const(uint8_t t, uint64_t imm) {
    TP n(new TNode);
    n->k = TNode::CONST;
    n->type = t;
    n->imm = imm;
    n->is_scalar = n->len1 = true;
    return n;
  }
  TP literal(const Json& val) {
    auto not_impl = [&]() -> Error {
      std::string what = val.kind == Json::String ? val.str : (val.kind == Json::Object && !val.obj.empty() ? val.obj[0].first : "?");
      return Error(CHDB_ERR_VALUE_TYPE_NOT_IMPLEMENTED, "value type not implemented: " + what);
    };
    if (val.kind != Json::Object || val.obj.size() != 1) throw not_impl();
    const std::string& tag = val.obj[0].first;
    const Json& body = val.obj[0].second;
    if (tag == "Number") {
      if (body.kind != Json::Array || body.arr.size() != 2 || body.arr[0].kind != Json::String)
        throw Error(CHDB_ERR_BAD_JSON, "Value::Number must be [string, bool]");
      const std::string& s = body.arr[0].str;
      if (body.arr[1].kind == Json::Bool && body.arr[1].b) throw not_impl();  // is_long
      if (s.find('.') != std::string::npos) {
        // f32 first; Rust's parser accepts every well-formed decimal (overflow -> inf), so the
        // f64 branch of the reference is unreachable for them.
        if (!rust_float_syntax(s)) throw Error(CHDB_ERR_FAILED_TO_PARSE_AS_A_FLOAT, "failed to parse " + s + " as a float");
        float f = std::strtof(s.c_str(), nullptr);  // glibc: correctly rounded, like Rust
        return make_const(T_F32, bits_of(f));
      }
      if (rust_int_syntax(s)) {
        errno = 0;
        long long v = std::strtoll(s.c_str(), nullptr, 10);
        if (errno == 0) {
          if (v >= INT32_MIN && v <= INT32_MAX) return make_const(T_I32, (uint64_t)(int64_t)v);
          return make_const(T_I64, (uint64_t)(int64_t)v);
        }
      }
      throw Error(CHDB_ERR_FAILED_TO_PARSE_AS_AN_INTEGER, "failed to parse " + s + " as an integer");
    }
    if (tag == "Boolean") {
      if (body.kind != Json::Bool) throw Error(CHDB_ERR_BAD_JSON, "Value::Boolean must be a bool");
      return make_const(T_BOOL, body.b ? 1 : 0);
    }
    if (tag == "SingleQuotedString") {
      if (body.kind != Json::String) throw Error(CHDB_ERR_BAD_JSON, "Value::SingleQuotedString must be a string");
      TP n = make_const(T_UTF8, 0);
      n->str = body.str;
      return n;
    }
    throw not_impl();
  }

  // ---- coercion lattice (compute_value.rs:350-431) ----
  static uint8_t common_type(uint8_t l, uint8_t r) {
    if (l == r) return l;
    auto sidx = [](uint8_t t) { return t - T_I8; };
    auto uidx = [](uint8_t t) { return t - T_U8; };
    if (is_signed_int(l) && is_signed_int(r)) return std::max(l, r);
    if (is_unsigned_int(l) && is_unsigned_int(r)) return std::max(l, r);
    for (int pass = 0; pass < 2; pass++) {
      uint8_t u = pass ? r : l, s = pass ? l : r;
      if (is_unsigned_int(u) && is_signed_int(s)) {
        if (sidx(s) > uidx(u)) return s;  // (uN, i2N or wider) -> the signed type
        break;
      }
    }
    if ((l == T_F32 && r == T_F64) || (l == T_F64 && r == T_F32)) return T_F64;
    for (int pass = 0; pass < 2; pass++) {
      uint8_t f = pass ? r : l, i = pass ? l : r;
      if (f == T_F32 && is_int(i) && type_width(i) <= 4) return T_F32;
      if (f == T_F64 && is_int(i)) return T_F64;
    }
    throw Error(CHDB_ERR_UNSUPPORTED_TYPE_COERSION, std::string("unsupported type coersion for operation between types ") +
                                                        type_arrow_name(l) + " and " + type_arrow_name(l));
    // (the reference's message prints the left type twice: "{0} and {0}", compute_value.rs:30)
  }
  TP cast_to(TP n, uint8_t to) {
    if (n->type == to) return n;
    if (n->k == TNode::CONST) {
      n->imm = fold_cast(n->type, to, n->imm);
      n->type = to;
      return n;
    }
    TP c(new TNode);
    c->k = TNode::CAST;
    c->type = to;
    c->is_scalar = n->is_scalar;
    c->len1 = n->len1;
    c->l = std::move(n);
    return c;
  }
  TP to_bool(TP n) {
    if (n->type == T_BOOL) return n;
    if (n->type == T_UTF8)  // arrow would parse 'true'/'false' strings; no reference query does this
      throw Error(CHDB_ERR_NOT_IMPLEMENTED, "not implemented: cast Utf8 to Boolean");
    if (n->k == TNode::CONST) {
      n->imm = fold_tobool(n->type, n->imm) ? 1 : 0;
      n->type = T_BOOL;
      return n;
    }
    TP c(new TNode);
    c->k = TNode::TOBOOL;
    c->type = T_BOOL;
    c->is_scalar = n->is_scalar;
    c->len1 = n->len1;
    c->l = std::move(n);
    return c;
  }
  void length_mismatch(int32_t code, const std::string& msg) {
    if (!prog.requires_single_row) {
      prog.requires_single_row = true;
      prog.single_row_code = code;
      prog.single_row_msg = msg;
    }
  }

  // ---- the tree walk (compute_value.rs:57-344) ----
  TP check(const Json& e) {
    if (++depth > 200) throw Error(CHDB_ERR_NOT_IMPLEMENTED, "expression nesting too deep");
    TP out = check_inner(e);
    --depth;
    return out;
  }
  TP check_inner(const Json& e) {
    std::string tag;
    const Json* body = nullptr;
    if (e.kind == Json::Object && e.obj.size() == 1) {
      tag = e.obj[0].first;
      body = &e.obj[0].second;
    } else if (e.kind == Json::String) {
      tag = e.str;
    } else {
      throw Error(CHDB_ERR_BAD_JSON, "an Expr must be an externally tagged enum");
    }
    if (tag == "Nested") return check(*body);
    if (tag == "Value") return literal(*body);
    if (tag == "Identifier") {
      std::string name = ident_value(*body);
      int idx = column_by_name(name);
      if (idx < 0) throw Error(CHDB_ERR_COLUMN_NOT_FOUND, "column not found: " + name);
      return make_col(idx);
    }
    if (tag == "CompoundIdentifier") {
      if (body->kind != Json::Array) throw Error(CHDB_ERR_BAD_JSON, "CompoundIdentifier must be an array");
      std::string joined;
      for (size_t i = 0; i < body->arr.size(); i++) joined += (i ? "." : "") + ident_value(body->arr[i]);
      if (body->arr.size() == 1) {
        int idx = column_by_name(joined);
        if (idx < 0) throw Error(CHDB_ERR_COLUMN_NOT_FOUND, "column not found: " + joined);
        return make_col(idx);
      }
      if (body->arr.size() == 2) {
        std::string alias = ident_value(body->arr[0]), name = ident_value(body->arr[1]);
        for (size_t i = 0; i < prog.schema.size(); i++) {
          if (prog.schema[i].name != name) continue;
          if (i >= aliases.size())
            throw Error(CHDB_ERR_PANIC, "table aliases vec has incorrect length");  // .expect() at :293
          for (auto& a : aliases[i])
            if (a == alias) return make_col((int)i);
        }
      }
      throw Error(CHDB_ERR_IDENTIFIER_NOT_FOUND, "identifier not found: \"" + joined + "\"");
    }
    if (tag == "BinaryOp") {
      const Json *lj = body->get("left"), *rj = body->get("right"), *oj = body->get("op");
      if (!lj || !rj || !oj) throw Error(CHDB_ERR_BAD_JSON, "BinaryOp needs left, op, right");
      TP l = check(*lj);
      TP r = check(*rj);
      std::string op = oj->kind == Json::String ? oj->str : (oj->kind == Json::Object && !oj->obj.empty() ? oj->obj[0].first : "?");
      return binary(op, std::move(l), std::move(r));
    }
    if ((ext & kExtOperators) && tag == "UnaryOp") {
      const Json *xj = body->get("expr"), *oj = body->get("op");
      if (!xj || !oj || oj->kind != Json::String) throw Error(CHDB_ERR_BAD_JSON, "UnaryOp needs op, expr");
      return unary(oj->str, check(*xj));
    }
    if ((ext & kExtOperators) && (tag == "IsNull" || tag == "IsNotNull")) return is_null(check(*body), tag == "IsNotNull");
    throw Error(CHDB_ERR_EXPRESSION_TYPE_NOT_IMPLEMENTED, "expression type not implemented: " + tag);
  }

  // ---- extension nodes: arrow-rs semantics of the kernels a Rust implementation would call ----
  TP unary(const std::string& op, TP x) {
    if (op == "Plus") return x;
    if (op == "Not") {   // compute::not over the Boolean cast, validity kept
      x = to_bool(std::move(x));
      if (x->k == TNode::CONST) { x->imm ^= 1; return x; }
      TP n(new TNode);
      n->k = TNode::NOT;
      n->type = T_BOOL;
      n->is_scalar = x->is_scalar;
      n->len1 = x->len1;
      n->l = std::move(x);
      return n;
    }
    if (op != "Minus") throw Error(CHDB_ERR_EXPRESSION_TYPE_NOT_IMPLEMENTED, "expression type not implemented: UnaryOp " + op);
    const uint8_t t = x->type;   // numeric::neg: checked on signed integers, sign flip on floats, no unsigned form
    if (!(is_signed_int(t) || t == T_F32 || t == T_F64))
      throw Error(CHDB_ERR_INVALID_ARGUMENT, std::string("Invalid argument error: Invalid arithmetic operation: -") + type_arrow_name(t));
    if (is_signed_int(t)) {
      const uint8_t order = (uint8_t)std::min(next_order++, 254);
      if (x->k == TNode::CONST) {
        TP c = make_const(t, fold_arith(OP_SUB, t, 0, x->imm));
        c->is_scalar = x->is_scalar;
        return c;
      }
      TP n(new TNode);
      n->k = TNode::ARITH;
      n->type = t;
      n->op = OP_SUB;
      n->order = order;
      n->is_scalar = x->is_scalar;
      n->len1 = x->len1;
      n->l = make_const(t, 0);
      n->r = std::move(x);
      return n;
    }
    if (x->k == TNode::CONST) { x->imm ^= t == T_F32 ? 0x80000000ull : (1ull << 63); return x; }
    TP n(new TNode);
    n->k = TNode::NEG;
    n->type = t;
    n->is_scalar = x->is_scalar;
    n->len1 = x->len1;
    n->l = std::move(x);
    return n;
  }
  TP is_null(TP x, bool negate) {   // arrow is_null / is_not_null: never null itself
    if (x->k == TNode::CONST) {     // literals are never null
      TP c = make_const(T_BOOL, negate ? 1 : 0);
      c->is_scalar = x->is_scalar;
      return c;
    }
    TP n(new TNode);
    n->k = TNode::ISNULL;
    n->type = T_BOOL;
    n->op = negate ? 1 : 0;
    n->is_scalar = x->is_scalar;
    n->len1 = x->len1;
    n->l = std::move(x);
    return n;
  }

  TP binary(const std::string& op, TP l, TP r) {
    if (op == "And" || op == "Or") {
      l = to_bool(std::move(l));
      r = to_bool(std::move(r));
      // arrow-arith binary_boolean_kernel: lengths must be equal, scalar-ness is ignored
      if (l->len1 != r->len1)
        length_mismatch(CHDB_ERR_COMPUTE_ERROR, "Compute error: Cannot perform bitwise operation on arrays of different length");
      TP n(new TNode);
      n->k = op == "And" ? TNode::AND : TNode::OR;
      n->op = (ext & kExtKleene) ? 1 : 0;
      n->type = T_BOOL;
      n->is_scalar = false;  // new_binary_op over bare BooleanArrays (compute_value.rs:81-85)
      n->len1 = l->len1 && r->len1;
      if (l->k == TNode::CONST && r->k == TNode::CONST) {
        TP c = make_const(T_BOOL, op == "And" ? (l->imm & r->imm) : (l->imm | r->imm));
        c->is_scalar = false;
        return c;
      }
      n->l = std::move(l);
      n->r = std::move(r);
      return n;
    }
    uint8_t arith = op == "Plus" ? OP_ADD : op == "Multiply" ? OP_MUL : op == "Divide" ? OP_DIV : op == "Modulo" ? OP_REM
                    : (op == "Minus" && (ext & kExtOperators)) ? OP_SUB : 0xFF;
    int cmp = op == "Eq" ? CMP_EQ : op == "NotEq" ? CMP_NE : op == "Lt" ? CMP_LT : op == "LtEq" ? CMP_LE
              : op == "Gt" ? CMP_GT : op == "GtEq" ? CMP_GE : -1;
    if (arith == 0xFF && cmp < 0)
      throw Error(CHDB_ERR_BINARY_OPERATOR_NOT_IMPLEMENTED, "binary operator not implemented: " + op);

    uint8_t t = common_type(l->type, r->type);  // cast_to_common_type (:433-461)
    l = cast_to(std::move(l), t);
    r = cast_to(std::move(r), t);
    bool both_scalar = l->is_scalar && r->is_scalar;
    bool len1 = both_scalar ? true : l->is_scalar ? r->len1 : r->is_scalar ? l->len1 : (l->len1 && r->len1);
    if (!l->is_scalar && !r->is_scalar && l->len1 != r->len1) {
      if (arith != 0xFF)
        length_mismatch(CHDB_ERR_COMPUTE_ERROR, "Compute error: Cannot perform binary operation on arrays of different length");
      else
        length_mismatch(CHDB_ERR_INVALID_ARGUMENT, "Invalid argument error: Cannot compare arrays of different lengths");
      len1 = false;
    }
    if (arith != 0xFF) {
      if (t == T_BOOL || t == T_UTF8)
        throw Error(CHDB_ERR_INVALID_ARGUMENT, std::string("Invalid argument error: Invalid arithmetic operation: ") +
                                                   type_arrow_name(t) + " " + op + " " + type_arrow_name(t));
      uint8_t order = (uint8_t)std::min(next_order++, 254);
      if (l->k == TNode::CONST && r->k == TNode::CONST) {
        TP c = make_const(t, fold_arith(arith, t, l->imm, r->imm));
        c->is_scalar = both_scalar;
        return c;
      }
      TP n(new TNode);
      n->k = TNode::ARITH;
      n->type = t;
      n->op = arith;
      n->order = order;
      n->is_scalar = both_scalar;
      n->len1 = len1;
      n->l = std::move(l);
      n->r = std::move(r);
      return n;
    }
    if (l->k == TNode::CONST && r->k == TNode::CONST) {
      bool v = t == T_UTF8 ? fold_cmp_str((uint8_t)cmp, l->str, r->str)
               : t == T_BOOL ? apply_cmp((uint8_t)cmp, l->imm < r->imm, l->imm == r->imm, l->imm > r->imm)
                             : fold_cmp((uint8_t)cmp, t, l->imm, r->imm);
      TP c = make_const(T_BOOL, v ? 1 : 0);
      c->is_scalar = both_scalar;
      return c;
    }
    TP n(new TNode);
    n->k = TNode::CMP;
    n->type = T_BOOL;
    n->op = (uint8_t)cmp;
    n->is_scalar = both_scalar;
    n->len1 = len1;
    n->l = std::move(l);
    n->r = std::move(r);
    return n;
  }

  // ---- code emission ----
  struct Operand {
    uint8_t src = SRC_NONE, slot = 0, from_type = T_NONE, type = T_NONE;
    uint64_t imm = 0;
  };
  bool as_operand(const TNode* n, Operand& o) {
    if (n->type == T_UTF8) return false;
    if (n->k == TNode::CONST) {
      o.src = SRC_IMM; o.imm = n->imm; o.type = o.from_type = n->type;
      return true;
    }
    const TNode* c = n;
    uint8_t to = n->type;
    if ((n->k == TNode::CAST || n->k == TNode::TOBOOL) && n->l->k == TNode::COL) c = n->l.get();
    if (c->k != TNode::COL) return false;
    o.src = SRC_COL; o.slot = (uint8_t)prog.slot_for(c->col); o.from_type = c->type; o.type = to;
    return true;
  }
  int need(const TNode* n) {
    Operand o;
    if (as_operand(n, o)) return 0;
    if (n->k == TNode::CAST || n->k == TNode::TOBOOL || n->k == TNode::NEG || n->k == TNode::NOT) return need(n->l.get());
    if (n->k == TNode::ISNULL) return n->l->k == TNode::COL ? 0 : need(n->l.get());
    if (n->k == TNode::CMP && n->l->type == T_UTF8) return 0;
    int a = need(n->l.get()), b = need(n->r.get());
    Operand ol, orr;
    if (as_operand(n->r.get(), orr)) return a;
    if (as_operand(n->l.get(), ol)) return b;
    return std::min(std::max(a, b + 1), std::max(b, a + 1));
  }
  void push(Instr in) {
    if ((int)prog.instrs.size() >= kMaxInstr)
      throw Error(CHDB_ERR_NOT_IMPLEMENTED, "expression lowers to more than 64 instructions");
    if (type_is_64(in.type) || ((in.src == SRC_COL || in.op == OP_CAST) && type_is_64(in.from_type))) prog.has64 = true;
    prog.instrs.push_back(in);
  }
  static Instr mk(uint8_t op, uint8_t type) {
    Instr in;
    std::memset(&in, 0, sizeof(in));
    in.op = op;
    in.type = type;
    in.from_type = type;
    return in;
  }
  void set_operand(Instr& in, const Operand& o) {
    in.src = o.src; in.slot = o.slot; in.from_type = o.from_type; in.imm = o.imm;
  }
  static uint8_t mirror(uint8_t k) {
    switch (k) {
      case CMP_LT: return CMP_GT;
      case CMP_LE: return CMP_GE;
      case CMP_GT: return CMP_LT;
      case CMP_GE: return CMP_LE;
      default: return k;
    }
  }
  uint8_t utf8_slot(const TNode* n, uint64_t& pool) {
    if (n->k == TNode::COL) return (uint8_t)prog.slot_for(n->col);
    if (prog.strpool.size() + n->str.size() > (size_t)kStrPoolBytes)
      throw Error(CHDB_ERR_NOT_IMPLEMENTED, "string literals longer than 256 bytes in total");
    pool = (uint64_t)prog.strpool.size() | ((uint64_t)n->str.size() << 32);
    prog.strpool += n->str;
    return 0xFF;
  }
  void emit(const TNode* n, int spill) {
    Operand o;
    if (as_operand(n, o)) {
      Instr in = mk(OP_LOAD, o.type);
      set_operand(in, o);
      push(in);
      return;
    }
    switch (n->k) {
      case TNode::CAST: {
        emit(n->l.get(), spill);
        uint8_t from = n->l->type, to = n->type;
        // 8/16/32-bit integers share one 32-bit accumulator form: widening among them is a no-op
        if (is_int(from) && is_int(to) && type_is_64(from) == type_is_64(to)) return;
        Instr in = mk(OP_CAST, to);
        in.from_type = from;
        push(in);
        return;
      }
      case TNode::TOBOOL: {
        emit(n->l.get(), spill);
        push(mk(OP_TOBOOL, n->l->type));
        return;
      }
      case TNode::NEG:
        emit(n->l.get(), spill);
        push(mk(OP_NEG, n->type));
        return;
      case TNode::NOT:
        emit(n->l.get(), spill);
        push(mk(OP_NOT, T_BOOL));
        return;
      case TNode::ISNULL: {
        Instr in = mk(OP_ISNULL, T_BOOL);
        if (n->op) in.flags |= OPF_NEGATE;
        if (n->l->k == TNode::COL) {   // only the column's validity bitmap is read (any type, Utf8 included)
          in.src = SRC_COL;
          in.slot = (uint8_t)prog.slot_for(n->l->col);
          in.from_type = T_BOOL;
        } else {
          emit(n->l.get(), spill);
        }
        push(in);
        return;
      }
      case TNode::CMP:
        if (n->l->type == T_UTF8) {
          Instr in = mk(OP_CMP_UTF8, T_UTF8);
          in.aux = n->op;
          uint64_t pool = 0;
          in.slot = utf8_slot(n->l.get(), pool);
          uint8_t b = utf8_slot(n->r.get(), pool);
          in.imm = ((uint64_t)b << 56) | pool;
          push(in);
          return;
        }
        break;
      default: break;
    }
    // binary node: ARITH / CMP / AND / OR
    uint8_t opc = n->k == TNode::ARITH ? n->op : n->k == TNode::CMP ? (uint8_t)OP_CMP : n->k == TNode::AND ? (uint8_t)OP_AND : (uint8_t)OP_OR;
    uint8_t optype = n->k == TNode::ARITH ? n->type : n->k == TNode::CMP ? n->l->type : (uint8_t)T_BOOL;
    Instr in = mk(opc, optype);
    in.order = n->order;
    in.aux = n->k == TNode::CMP ? n->op : 0;
    if ((n->k == TNode::AND || n->k == TNode::OR) && n->op) in.flags |= OPF_KLEENE;
    const TNode *l = n->l.get(), *r = n->r.get();
    Operand ol, orr;
    bool lo = as_operand(l, ol), ro = as_operand(r, orr);
    auto swapped = [&]() {  // acc holds the RIGHT child, operand is the LEFT one
      if (n->k == TNode::CMP) in.aux = mirror(in.aux);
      else if (n->k == TNode::ARITH) {
        // integer + and * commute exactly; floats keep source order (which NaN operand propagates)
        const bool commutes = (n->op == OP_ADD || n->op == OP_MUL) && is_int(n->type);
        if (!commutes) in.flags |= OPF_SWAP;
      }
    };
    if (ro && !(lo && ol.src == SRC_IMM && orr.src != SRC_IMM)) {
      emit(l, spill);
      set_operand(in, orr);
    } else if (lo) {
      emit(r, spill);
      set_operand(in, ol);
      swapped();
    } else {
      if (spill >= kMaxSpill) throw Error(CHDB_ERR_NOT_IMPLEMENTED, "expression needs more than 6 spill slots");
      prog.max_spill = std::max(prog.max_spill, spill + 1);
      bool left_first = need(l) >= need(r);
      const TNode *first = left_first ? l : r, *second = left_first ? r : l;
      emit(first, spill);
      Instr p = mk(OP_PUSH, first->type);
      p.slot = (uint8_t)spill;
      push(p);
      emit(second, spill + 1);
      in.src = SRC_STK;
      in.slot = (uint8_t)spill;
      in.from_type = optype;
      if (left_first) swapped();  // spill slot holds the left child
    }
    push(in);
  }
  // emits code for a typed tree; returns [begin, end)
  std::pair<int, int> emit_expr(const TNode* n) {
    int b = (int)prog.instrs.size();
    emit(n, 0);
    return {b, (int)prog.instrs.size()};
  }
};

// ------------------------------------------------------------------------------------------
// program assembly
// ------------------------------------------------------------------------------------------
static void parse_aliases(const char* aliases_json, Lowering& L, size_t ncols) {
  if (!aliases_json || !*aliases_json) {
    L.aliases.assign(ncols, {});
    return;
  }
  Json a = JsonParser(aliases_json).parse();
  if (a.kind != Json::Array) throw Error(CHDB_ERR_BAD_JSON, "table_aliases must be a JSON array of arrays");
  for (auto& col : a.arr) {
    if (col.kind != Json::Array) throw Error(CHDB_ERR_BAD_JSON, "table_aliases must be a JSON array of arrays");
    std::vector<std::string> names;
    for (auto& s : col.arr) names.push_back(s.str);
    L.aliases.push_back(std::move(names));
  }
}

static OutputColumn pass_column(Program& p, int col, bool keep_declared) {
  const InputColumn& c = p.schema[col];
  OutputColumn o;
  o.kind = OutputColumn::PASS;
  o.name = c.name;
  o.type = c.type;
  o.format = c.format;
  o.width = c.width;
  o.in_col = col;
  o.slot = -1;  // assigned by the caller when the column has to go through the kernel
  o.keep_declared_nullable = keep_declared;
  o.declared_nullable = (c.flags & ARROW_FLAG_NULLABLE) != 0;
  return o;
}

static void add_expr_output(Lowering& L, Program& p, const Json& expr, const std::string& name) {
  TP t = L.check(expr);
  OutputColumn o;
  o.name = name;
  o.type = (TypeId)t->type;
  o.format = type_format(t->type);
  o.width = type_width(t->type);
  if (t->k == TNode::COL) {
    o = pass_column(p, t->col, false);
    o.name = name;
  } else if (t->k == TNode::CONST) {
    o.kind = OutputColumn::CONST;
    o.imm = t->imm;
    o.str = t->str;
  } else {
    o.kind = OutputColumn::EXPR;
    auto range = L.emit_expr(t.get());
    o.begin = range.first;
    o.end = range.second;
  }
  p.outputs.push_back(std::move(o));
}

static void lower_select_items(Lowering& L, Program& p, const Json& items) {
  if (items.kind != Json::Array) throw Error(CHDB_ERR_BAD_JSON, "select items must be a JSON array of SelectItem");
  size_t unnamed_idx = 0;
  for (const Json& item : items.arr) {
    std::string tag;
    const Json* body = nullptr;
    if (item.kind == Json::Object && item.obj.size() == 1) {
      tag = item.obj[0].first;
      body = &item.obj[0].second;
    } else if (item.kind == Json::String) {
      tag = item.str;
    } else {
      throw Error(CHDB_ERR_BAD_JSON, "a SelectItem must be an externally tagged enum");
    }
    if (tag == "Wildcard") {  // record_projection.rs:27-32: clone every field and array
      for (size_t c = 0; c < p.schema.size(); c++) p.outputs.push_back(pass_column(p, (int)c, true));
    } else if (tag == "QualifiedWildcard") {  // :33-38
      throw Error(CHDB_ERR_NOT_IMPLEMENTED, "not implemented: SelectItem::QualifiedWildcard");
    } else if (tag == "UnnamedExpr") {  // :39-59
      std::string name = "unnamed_" + std::to_string(unnamed_idx);
      if (body->kind == Json::Object && body->obj.size() == 1 && body->obj[0].first == "Identifier")
        name = Lowering::ident_value(body->obj[0].second);
      add_expr_output(L, p, *body, name);
      unnamed_idx++;
    } else if (tag == "ExprWithAlias") {  // :60-68
      const Json *e = body->get("expr"), *a = body->get("alias");
      if (!e || !a) throw Error(CHDB_ERR_BAD_JSON, "ExprWithAlias needs expr and alias");
      add_expr_output(L, p, *e, Lowering::ident_value(*a));
    } else {
      throw Error(CHDB_ERR_NOT_IMPLEMENTED, "not implemented: SelectItem::" + tag);
    }
  }
}

static void lower_predicate(Lowering& L, Program& p, const Json& expr) {
  TP t = L.check(expr);
  if (t->type != T_BOOL)  // filter_record.rs:27-35
    throw Error(CHDB_ERR_CAST_TO_BOOLEAN_ARRAY_FAILED,
                std::string("cast to boolean array failed for array type: ") + type_arrow_name(t->type));
  p.has_pred = true;
  if (t->k == TNode::CONST) {
    p.pred_const = true;
    p.pred_const_value = t->imm != 0;
    return;
  }
  auto range = L.emit_expr(t.get());
  p.pred_begin = range.first;
  p.pred_end = range.second;
}

std::unique_ptr<Program> compile_program(Program::Mode mode, const char* expr_json, const char* items_json,
                                         const ::ArrowSchema* schema, const char* aliases_json) {
  std::unique_ptr<Program> p(new Program);
  p->mode = mode;
  p->schema = parse_schema(schema);
  Lowering L(*p);
  parse_aliases(aliases_json, L, p->schema.size());
  if (mode != Program::PROJECT) {
    if (!expr_json) throw Error(CHDB_ERR_INVALID_ARGUMENT, "expr_json is null");
    lower_predicate(L, *p, JsonParser(expr_json).parse());
  }
  if (mode == Program::FILTER) {
    for (size_t c = 0; c < p->schema.size(); c++) p->outputs.push_back(pass_column(*p, (int)c, true));
  } else {
    if (!items_json) throw Error(CHDB_ERR_INVALID_ARGUMENT, "select_items_json is null");
    lower_select_items(L, *p, JsonParser(items_json).parse());
  }
  if ((int)p->outputs.size() > kMaxOutCols)
    throw Error(CHDB_ERR_NOT_IMPLEMENTED, "more than 24 output columns");
  // Columns that must travel through the kernel: everything when rows are compacted; in a pure
  // projection PASS columns are shared with the input (the reference clones the Arc).
  for (auto& o : p->outputs)
    if (o.kind == OutputColumn::PASS && p->has_pred && !p->pred_const) o.slot = p->slot_for(o.in_col);
  for (auto& o : p->outputs)
    if (type_is_64(o.type) && o.kind != OutputColumn::CONST) p->has64 = p->has64 || o.kind == OutputColumn::EXPR;
  return p;
}

std::unique_ptr<Program> compile_value(const char* expr_json, const ::ArrowSchema* schema, const char* aliases_json,
                                       bool* is_scalar) {
  std::unique_ptr<Program> p(new Program);
  p->mode = Program::PROJECT;
  p->schema = parse_schema(schema);
  Lowering L(*p);
  parse_aliases(aliases_json, L, p->schema.size());
  if (!expr_json) throw Error(CHDB_ERR_INVALID_ARGUMENT, "expr_json is null");
  Json e = JsonParser(expr_json).parse();
  TP t = L.check(e);
  if (is_scalar) *is_scalar = t->is_scalar;
  OutputColumn o;
  o.name = "value";
  o.type = (TypeId)t->type;
  o.format = type_format(t->type);
  o.width = type_width(t->type);
  if (t->k == TNode::COL) {
    o = pass_column(*p, t->col, false);
    o.name = "value";
  } else if (t->k == TNode::CONST) {
    o.kind = OutputColumn::CONST;
    o.imm = t->imm;
    o.str = t->str;
  } else {
    o.kind = OutputColumn::EXPR;
    auto range = L.emit_expr(t.get());
    o.begin = range.first;
    o.end = range.second;
  }
  p->outputs.push_back(std::move(o));
  return p;
}

// ------------------------------------------------------------------------------------------
// disassembler
// ------------------------------------------------------------------------------------------
std::string Program::disassemble() const {
  static const char* opn[] = {"load", "cast", "add", "mul", "div", "rem", "sub", "cmp", "tobool", "and", "or", "push", "cmp_utf8", "end", "neg", "not", "isnull"};
  static const char* cmpn[] = {"eq", "ne", "lt", "le", "gt", "ge"};
  std::ostringstream os;
  os << "mode=" << (mode == FILTER ? "filter" : mode == PROJECT ? "project" : "filter_project")
     << " instrs=" << instrs.size() << " slots=" << slot_to_col.size() << " spill=" << max_spill
     << " container=" << (has64 ? "u64" : "u32") << (requires_single_row ? " requires_single_row" : "") << "\n";
  for (size_t s = 0; s < slot_to_col.size(); s++)
    os << "  slot " << s << " = column " << slot_to_col[s] << " '" << schema[slot_to_col[s]].name << "' "
       << type_arrow_name(schema[slot_to_col[s]].type) << "\n";
  auto dump = [&](int b, int e) {
    for (int i = b; i < e; i++) {
      const Instr& in = instrs[i];
      os << "    " << i << ": " << opn[in.op] << ((in.flags & OPF_KLEENE) ? "_kleene" : "") << ((in.flags & OPF_NEGATE) ? "_not" : "");
      if (in.op == OP_CMP || in.op == OP_CMP_UTF8) os << "." << cmpn[in.aux];
      os << "." << type_arrow_name(in.type);
      if (in.flags & OPF_SWAP) os << " swap";
      if (in.op == OP_CMP_UTF8) {
        auto side = [&](uint8_t slot) {
          if (slot != 0xFF) os << " col" << (int)slot;
          else os << " '" << strpool.substr((uint32_t)in.imm, (in.imm >> 32) & 0xFFFFFF) << "'";
        };
        side(in.slot);
        side((uint8_t)(in.imm >> 56));
      } else if (in.op == OP_CAST) {
        os << " from " << type_arrow_name(in.from_type);
      } else if (in.op == OP_PUSH) {
        os << " spill" << (int)in.slot;
      } else if (in.src == SRC_COL) {
        os << " col" << (int)in.slot;
        if (in.from_type != in.type) os << " (" << type_arrow_name(in.from_type) << ")";
      } else if (in.src == SRC_IMM) {
        os << " imm=0x" << std::hex << in.imm << std::dec;
      } else if (in.src == SRC_STK) {
        os << " spill" << (int)in.slot;
      }
      if (in.op >= OP_ADD && in.op <= OP_SUB) os << " order=" << (int)in.order;
      os << "\n";
    }
  };
  if (has_pred) {
    if (pred_const) os << "  predicate: len-1 constant " << (pred_const_value ? "true" : "false") << "\n";
    else { os << "  predicate:\n"; dump(pred_begin, pred_end); }
  }
  for (size_t k = 0; k < outputs.size(); k++) {
    const OutputColumn& o = outputs[k];
    os << "  out " << k << " '" << o.name << "' " << type_arrow_name(o.type);
    if (o.kind == OutputColumn::PASS) os << " = column " << o.in_col << (o.slot >= 0 ? " (compacted)" : " (shared)") << "\n";
    else if (o.kind == OutputColumn::CONST) os << " = len-1 constant\n";
    else { os << " =\n"; dump(o.begin, o.end); }
  }
  return os.str();
}

}  // namespace chdb
