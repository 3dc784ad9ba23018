"""SQL text -> the serde-JSON form of sqlparser 0.52 `Expr` / `SelectItem`.

The reference plans queries with `sqlparser::Parser::parse_sql(GenericDialect, ..)`
(src/planner/logical_planner.rs:241-248) and ships the resulting `Expr` /
`Vec<SelectItem>` between workers as serde JSON
(src/handlers/message_handler/messages/query.rs:424-431; Cargo.toml:29 enables
sqlparser's `serde` feature).  That JSON is what the C-ABI consumes.  There is no
Rust toolchain in this image, so this module re-creates the same trees for the
subset of SQL the sample queries use, so tests and benchmarks can be written as
SQL like the reference's own (record_utils/test_compute_value.rs:127-175).

Covers: SELECT list (`*`, `t.*`, expr, expr AS alias), FROM (kept as text, with
an optional table alias), WHERE; expressions with sqlparser's precedences
(OR 5 < AND 10 < NOT 15 < comparison 20 < + - 30 < * / % 40 < unary), literals
(numbers, 'strings', TRUE/FALSE, NULL), identifiers (`a`, `a.b`, "quoted"),
parentheses (-> `Nested`), IS [NOT] NULL, function calls (opaque).
"""
from __future__ import annotations

import json
import re

_TOKEN = re.compile(r"""
    (?P<ws>\s+|--[^\n]*)
  | (?P<num>(\d+\.\d*|\.\d+|\d+)([eE][+-]?\d+)?L?)
  | (?P<str>'(?:[^']|'')*')
  | (?P<qid>"(?:[^"]|"")*")
  | (?P<id>[A-Za-z_][A-Za-z_0-9$]*)
  | (?P<op><>|!=|>=|<=|==|=>|\|\||[-+*/%=<>(),.;])
""", re.X)

_WILDCARD_OPTS = {"opt_ilike": None, "opt_exclude": None, "opt_except": None, "opt_replace": None, "opt_rename": None}
_CMP = {"=": "Eq", "==": "Eq", "<>": "NotEq", "!=": "NotEq", "<": "Lt", "<=": "LtEq", ">": "Gt", ">=": "GtEq"}


class SqlParseError(ValueError):
    pass


def _tokenize(sql: str):
    pos, out = 0, []
    while pos < len(sql):
        m = _TOKEN.match(sql, pos)
        if not m:
            raise SqlParseError(f"unexpected character {sql[pos]!r} at {pos}")
        pos = m.end()
        kind = m.lastgroup
        if kind == "ws":
            continue
        out.append((kind, m.group(kind)))
    out.append(("eof", ""))
    return out


def ident(value: str, quote_style=None) -> dict:
    return {"value": value, "quote_style": quote_style}


class _Parser:
    def __init__(self, sql: str):
        self.toks = _tokenize(sql)
        self.i = 0

    # -- token helpers ---------------------------------------------------------
    def peek(self, k=0):
        return self.toks[min(self.i + k, len(self.toks) - 1)]

    def next(self):
        t = self.toks[self.i]
        self.i += 1
        return t

    def is_kw(self, word: str, k=0) -> bool:
        kind, text = self.peek(k)
        return kind == "id" and text.upper() == word

    def accept_kw(self, word: str) -> bool:
        if self.is_kw(word):
            self.i += 1
            return True
        return False

    def accept_op(self, op: str) -> bool:
        kind, text = self.peek()
        if kind == "op" and text == op:
            self.i += 1
            return True
        return False

    def expect_op(self, op: str):
        if not self.accept_op(op):
            raise SqlParseError(f"expected {op!r}, found {self.peek()[1]!r}")

    # -- expressions (precedence climbing with sqlparser's numbers) ------------
    def next_precedence(self) -> int:
        kind, text = self.peek()
        if kind == "id":
            up = text.upper()
            if up == "OR":
                return 5
            if up == "AND":
                return 10
            if up == "IS":
                return 17
            return 0
        if kind == "op":
            if text in _CMP:
                return 20
            if text in ("+", "-"):
                return 30
            if text in ("*", "/", "%"):
                return 40
            if text == "||":
                return 30
        return 0

    def parse_expr(self, precedence: int = 0):
        expr = self.parse_prefix()
        while True:
            nxt = self.next_precedence()
            if precedence >= nxt:
                return expr
            expr = self.parse_infix(expr, nxt)

    def parse_infix(self, left, precedence: int):
        kind, text = self.next()
        if kind == "id":
            up = text.upper()
            if up in ("AND", "OR"):
                right = self.parse_expr(precedence)
                return {"BinaryOp": {"left": left, "op": "And" if up == "AND" else "Or", "right": right}}
            if up == "IS":
                neg = self.accept_kw("NOT")
                if self.accept_kw("NULL"):
                    return {"IsNotNull" if neg else "IsNull": left}
                raise SqlParseError("expected NULL after IS [NOT]")
        op = _CMP.get(text) or {"+": "Plus", "-": "Minus", "*": "Multiply", "/": "Divide", "%": "Modulo",
                                "||": "StringConcat"}.get(text)
        if op is None:
            raise SqlParseError(f"unexpected infix token {text!r}")
        right = self.parse_expr(precedence)
        return {"BinaryOp": {"left": left, "op": op, "right": right}}

    def parse_prefix(self):
        kind, text = self.next()
        if kind == "num":
            if text.endswith("L"):
                return {"Value": {"Number": [text[:-1], True]}}
            return {"Value": {"Number": [text, False]}}
        if kind == "str":
            return {"Value": {"SingleQuotedString": text[1:-1].replace("''", "'")}}
        if kind == "op" and text == "(":
            inner = self.parse_expr()
            self.expect_op(")")
            return {"Nested": inner}
        if kind == "op" and text in ("-", "+"):
            # sqlparser: unary +/- binds at MulDiv precedence (40)
            return {"UnaryOp": {"op": "Minus" if text == "-" else "Plus", "expr": self.parse_expr(40)}}
        if kind == "id":
            up = text.upper()
            if up == "TRUE":
                return {"Value": {"Boolean": True}}
            if up == "FALSE":
                return {"Value": {"Boolean": False}}
            if up == "NULL":
                return {"Value": "Null"}
            if up == "NOT":
                return {"UnaryOp": {"op": "Not", "expr": self.parse_expr(15)}}
            return self.parse_identifier_chain(ident(text))
        if kind == "qid":
            return self.parse_identifier_chain(ident(text[1:-1].replace('""', '"'), '"'))
        raise SqlParseError(f"unexpected token {text!r}")

    def parse_identifier_chain(self, first: dict):
        parts = [first]
        while self.peek() == ("op", ".") and self.peek(1)[0] in ("id", "qid"):
            self.next()
            kind, text = self.next()
            parts.append(ident(text) if kind == "id" else ident(text[1:-1], '"'))
        if self.peek() == ("op", "("):
            # function call: the hot path rejects it (ExpressionTypeNotImplemented); keep it opaque
            depth, args_text = 0, []
            while True:
                kind, text = self.next()
                if kind == "eof":
                    raise SqlParseError("unterminated function call")
                args_text.append(text)
                if (kind, text) == ("op", "("):
                    depth += 1
                elif (kind, text) == ("op", ")"):
                    depth -= 1
                    if depth == 0:
                        break
            return {"Function": {"name": parts, "args_text": " ".join(args_text)}}
        if len(parts) == 1:
            return {"Identifier": parts[0]}
        return {"CompoundIdentifier": parts}

    # -- SELECT ----------------------------------------------------------------
    def parse_select_item(self):
        if self.accept_op("*"):
            return {"Wildcard": dict(_WILDCARD_OPTS)}
        # t.* / a.b.*
        j, names = 0, []
        while self.peek(j)[0] in ("id", "qid") and self.peek(j + 1) == ("op", "."):
            names.append(self.peek(j))
            if self.peek(j + 2) == ("op", "*"):
                self.i += j + 3
                obj = [ident(t) if k == "id" else ident(t[1:-1], '"') for k, t in names]
                return {"QualifiedWildcard": [obj, dict(_WILDCARD_OPTS)]}
            j += 2
        expr = self.parse_expr()
        if self.accept_kw("AS"):
            kind, text = self.next()
            if kind not in ("id", "qid"):
                raise SqlParseError("expected alias after AS")
            return {"ExprWithAlias": {"expr": expr, "alias": ident(text) if kind == "id" else ident(text[1:-1], '"')}}
        kind, text = self.peek()
        if kind == "id" and text.upper() not in ("FROM", "WHERE"):
            self.next()
            return {"ExprWithAlias": {"expr": expr, "alias": ident(text)}}
        return {"UnnamedExpr": expr}

    def parse_select(self):
        if not self.accept_kw("SELECT"):
            raise SqlParseError("expected SELECT")
        items = [self.parse_select_item()]
        while self.accept_op(","):
            items.append(self.parse_select_item())
        from_text, alias = None, None
        if self.accept_kw("FROM"):
            start, depth = self.i, 0
            while True:
                kind, text = self.peek()
                if kind == "eof" or (depth == 0 and (self.is_kw("WHERE") or (kind, text) == ("op", ";"))):
                    break
                if (kind, text) == ("op", "("):
                    depth += 1
                if (kind, text) == ("op", ")"):
                    depth -= 1
                self.next()
            toks = self.toks[start:self.i]
            # `read_files(...) [AS] alias`
            if len(toks) >= 2 and toks[-1][0] == "id" and toks[-2] == ("op", ")"):
                alias = toks[-1][1]
                toks = toks[:-1]
            elif len(toks) >= 3 and toks[-1][0] == "id" and toks[-2][0] == "id" and toks[-2][1].upper() == "AS":
                alias = toks[-1][1]
                toks = toks[:-2]
            from_text = "".join(t for _, t in toks)
        where = self.parse_expr() if self.accept_kw("WHERE") else None
        self.accept_op(";")
        if self.peek()[0] != "eof":
            raise SqlParseError(f"unexpected trailing token {self.peek()[1]!r}")
        return {"projection": items, "selection": where, "from": from_text, "alias": alias}


def parse_expr(sql: str) -> dict:
    """Parse one SQL expression into sqlparser-serde form (a python dict)."""
    p = _Parser(sql)
    e = p.parse_expr()
    if p.peek()[0] != "eof":
        raise SqlParseError(f"unexpected trailing token {p.peek()[1]!r}")
    return e


def parse_select(sql: str) -> dict:
    """Parse `SELECT items FROM ... [WHERE expr]`; returns dict(projection, selection, from, alias)."""
    return _Parser(sql).parse_select()


def split_statements(sql_text: str) -> list[str]:
    """Split a sample_queries/*.sql file into its statements (comments dropped)."""
    lines = [ln for ln in sql_text.splitlines() if not ln.strip().startswith("--")]
    return [s.strip() for s in "\n".join(lines).split(";") if s.strip()]


def to_json(tree) -> str:
    return json.dumps(tree, separators=(",", ":"))
