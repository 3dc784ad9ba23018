// Register bytecode shared by the host lowering (lower.cpp) and the device interpreter
// (kernels.cu).  The machine is an accumulator machine: one typed accumulator holding the
// thread's rows in real registers, one operand fetched per instruction (column, immediate
// or spill slot), and a small spill stack used only when both children of a node are
// non-leaf expressions.
#pragma once
#ifdef __CUDACC_RTC__
// NVRTC (run-time specialisation, see jit.cpp) has no system headers: spell the basics out.
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef unsigned long long uintptr_t;
typedef unsigned long size_t;
#define INT32_MIN (-2147483647 - 1)
#define INT64_MIN (-9223372036854775807ll - 1)
#else
#include <stdint.h>
#endif

#ifdef __CUDACC__
#define CHDB_HD __host__ __device__
#else
#define CHDB_HD
#endif

namespace chdb {

// Arrow types the evaluator computes on.  Numbering is part of the kernel ABI.
enum TypeId : uint8_t {
  T_BOOL = 0, T_I8 = 1, T_I16 = 2, T_I32 = 3, T_I64 = 4,
  T_U8 = 5, T_U16 = 6, T_U32 = 7, T_U64 = 8, T_F32 = 9, T_F64 = 10, T_UTF8 = 11,
  T_OPAQUE = 12,  // fixed-width pass-through only (width in ColumnDesc)
  T_NONE = 255
};

// How a value sits in the accumulator: every integer type is sign- or zero-extended to the
// full container, so widening integer casts are no-ops and only the class matters.
enum TypeClass : uint8_t { C_BOOL = 0, C_SINT = 1, C_UINT = 2, C_S64 = 3, C_U64 = 4, C_F32 = 5, C_F64 = 6, C_UTF8 = 7 };

CHDB_HD inline TypeClass type_class(uint8_t t) {
  switch (t) {
    case T_BOOL: return C_BOOL;
    case T_I8: case T_I16: case T_I32: return C_SINT;
    case T_U8: case T_U16: case T_U32: return C_UINT;
    case T_I64: return C_S64;
    case T_U64: return C_U64;
    case T_F32: return C_F32;
    case T_F64: return C_F64;
    default: return C_UTF8;
  }
}
CHDB_HD inline int type_width(uint8_t t) {
  switch (t) {
    case T_I8: case T_U8: return 1;
    case T_I16: case T_U16: return 2;
    case T_I32: case T_U32: case T_F32: return 4;
    case T_I64: case T_U64: case T_F64: return 8;
    default: return 0;
  }
}
CHDB_HD inline bool type_is_64(uint8_t t) { return t == T_I64 || t == T_U64 || t == T_F64; }

enum Opcode : uint8_t {
  OP_LOAD = 0,    // acc = operand
  OP_CAST = 1,    // acc = cast(acc: from_type -> type)
  OP_ADD = 2,     // acc = acc + operand   (OPF_SWAP: operand + acc)
  OP_MUL = 3,
  OP_DIV = 4,
  OP_REM = 5,
  OP_SUB = 6,     // extension (the reference rejects Minus); never emitted by default
  OP_CMP = 7,     // acc(bool) = acc <cmp> operand; cmp kind in `aux`
  OP_TOBOOL = 8,  // acc(bool) = acc != 0          (arrow cast numeric -> Boolean)
  OP_AND = 9,     // acc(bool) = acc & operand, validity = both valid (non-Kleene)
  OP_OR = 10,
  OP_PUSH = 11,   // spill[slot] = acc
  OP_CMP_UTF8 = 12,  // acc(bool) = utf8 operand A <cmp> utf8 operand B (columns / pool strings)
  OP_END = 13,
  // SQL extensions (chdb_set_sql_extensions; the reference rejects these nodes, so none is emitted by default)
  OP_NEG = 14,    // acc = -acc for floats (sign flip; integers lower to checked 0 - x)
  OP_NOT = 15,    // acc(bool) = !acc, validity kept (arrow compute::not)
  OP_ISNULL = 16  // acc(bool) = validity of acc / of the column operand is clear (OPF_NEGATE: IS NOT NULL); never null
};

enum CmpKind : uint8_t { CMP_EQ = 0, CMP_NE = 1, CMP_LT = 2, CMP_LE = 3, CMP_GT = 4, CMP_GE = 5 };
enum SrcKind : uint8_t { SRC_NONE = 0, SRC_COL = 1, SRC_IMM = 2, SRC_STK = 3 };
enum InstrFlags : uint8_t { OPF_SWAP = 1, OPF_KLEENE = 2 /* OP_AND / OP_OR: SQL three-valued logic */, OPF_NEGATE = 4 };

// 16 bytes; lives in kernel parameter (constant) space, read with uniform loads.
struct Instr {
  uint8_t op;
  uint8_t type;       // type the operation computes in (result type for OP_CAST / OP_LOAD)
  uint8_t src;        // SrcKind of the operand
  uint8_t flags;
  uint8_t slot;       // SRC_COL: input column slot; SRC_STK / OP_PUSH: spill slot; OP_CMP_UTF8: column A or 0xFF
  uint8_t from_type;  // SRC_COL: stored type of the column (cast to `type` on fetch); OP_CAST: source type
  uint8_t aux;        // OP_CMP / OP_CMP_UTF8: CmpKind
  uint8_t order;      // reference post-order index of the node (error priority)
  uint64_t imm;       // SRC_IMM: value bits in `type`; OP_CMP_UTF8: see UTF8 packing below
};
static_assert(sizeof(Instr) == 16, "Instr must be 16 bytes");

// OP_CMP_UTF8 operand packing in imm: [63:56] column B slot or 0xFF; when an operand is a
// literal its bytes sit in the string pool: [31:0] pool offset, [55:32] length (either A or B,
// never both -- literal-vs-literal is folded on the host).
constexpr int kMaxInstr = 64;
constexpr int kMaxInCols = 24;    // referenced + pass-through input column slots
constexpr int kMaxOutCols = 24;
constexpr int kMaxSpill = 6;
constexpr int kStrPoolBytes = 256;
// (sized so that the whole KernelParams block stays under the classic 4 KB parameter limit)

// chdb_code values the kernels can raise (static_assert'ed against include/chdb_gpu.h in runtime.cu)
constexpr uint32_t kErrArithmeticOverflow = 12;
constexpr uint32_t kErrDivideByZero = 13;

// Device error word: atomicMax of ~packed, packed = [63:56] order, [55:8] row, [7:0] chdb_code;
// 0 = no error, larger = earlier in the reference's evaluation order.

}  // namespace chdb
