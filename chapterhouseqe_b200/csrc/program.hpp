// Compiled program: the planner's expression tree(s) lowered for one input schema.
#pragma once
#include <atomic>
#include <memory>
#include <string>
#include <vector>

#include "bytecode.h"
#include "errors.hpp"

namespace chdb {

struct InputColumn {
  std::string name;
  std::string format;   // Arrow C format string, kept verbatim for pass-through
  int64_t flags = 0;    // ARROW_FLAG_NULLABLE as declared
  TypeId type = T_NONE;
  int width = 0;        // bytes per value for fixed-width types (bool = 0, utf8 = 0)
  bool supported = true;  // layout this library can move (fixed width, bool, utf8)
};

struct OutputColumn {
  enum Kind { PASS, EXPR, CONST } kind = PASS;
  std::string name;
  TypeId type = T_NONE;
  std::string format;
  int width = 0;
  int in_col = -1;           // PASS: index into the input schema
  int slot = -1;             // PASS: kernel column slot
  int begin = 0, end = 0;    // EXPR: instruction range
  bool keep_declared_nullable = false;  // filter / wildcard: clone the input field
  bool declared_nullable = false;
  uint64_t imm = 0;          // CONST: len-1 value
  std::string str;
};

struct Program {
  enum Mode { FILTER = 0, PROJECT = 1, FILTER_PROJECT = 2 } mode = FILTER;
  std::vector<InputColumn> schema;
  std::vector<int> slot_to_col;   // kernel column slot -> input schema index
  std::vector<Instr> instrs;
  std::string strpool;
  bool has_pred = false;
  int pred_begin = 0, pred_end = 0;
  bool pred_const = false;        // len-1 predicate: arrow-select does not broadcast it
  bool pred_const_value = false;
  std::vector<OutputColumn> outputs;
  bool requires_single_row = false;  // some node mixes a len-N array with a len-1 non-scalar array
  int32_t single_row_code = CHDB_OK;
  std::string single_row_msg;
  bool has64 = false;
  int max_spill = 0;

  int slot_for(int col);          // allocates on first use
  std::string disassemble() const;
};

// sqlparser-serde JSON + Arrow schema -> Program.  Throws chdb::Error.
std::vector<InputColumn> parse_schema(const ::ArrowSchema* schema);
std::unique_ptr<Program> compile_program(Program::Mode mode, const char* expr_json, const char* items_json,
                                         const ::ArrowSchema* schema, const char* aliases_json);
// compute_value(): one output named "value"; *is_scalar as ArrayDatum.is_scalar.
std::unique_ptr<Program> compile_value(const char* expr_json, const ::ArrowSchema* schema,
                                       const char* aliases_json, bool* is_scalar);

// chdb_set_sql_extensions: bit 0 = binary -, unary - / +, NOT, IS [NOT] NULL; bit 1 = Kleene AND / OR (read at compile time)
extern std::atomic<uint32_t> g_sql_extensions;

const char* type_arrow_name(uint8_t t);   // "Int32", "Float32", ...
const char* type_format(uint8_t t);       // Arrow C format string

}  // namespace chdb
