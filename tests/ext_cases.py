"""Inputs for the SQL-extension tests: nodes the reference's compute_value rejects and this library
accepts when chdb_set_sql_extensions asks for them."""
import numpy as np
import pyarrow as pa


def table(n: int, seed: int = 0) -> pa.RecordBatch:
    rng = np.random.default_rng(seed)

    def mask(p):
        return rng.random(n) < p if n else None

    f = rng.normal(0, 10, n).astype(np.float32)
    d = rng.normal(0, 1, n)
    if n > 8:   # specials: NaN of both signs, +-0, infinities
        f[:6] = np.array([np.nan, -np.nan, 0.0, -0.0, np.inf, -np.inf], dtype=np.float32)
        d[2:8] = np.array([np.nan, -np.nan, 0.0, -0.0, np.inf, -np.inf])
    return pa.RecordBatch.from_arrays(
        [pa.array(np.arange(n, dtype=np.int32)),
         pa.array(rng.integers(-1000, 1000, n).astype(np.int32), mask=mask(0.2)),
         pa.array(rng.integers(-2**40, 2**40, n, dtype=np.int64), mask=mask(0.1)),
         pa.array(rng.integers(-100, 100, n).astype(np.int16), mask=mask(0.3)),
         pa.array(rng.integers(0, 1000, n).astype(np.uint32)),
         pa.array(f, mask=mask(0.15)),
         pa.array(d, mask=mask(0.05)),
         pa.array(rng.random(n) < 0.5, mask=mask(0.25)),
         pa.array(rng.random(n) < 0.3, mask=mask(0.25)),
         pa.array(np.char.add("s", rng.integers(0, 50, n).astype(str)), mask=mask(0.4))],
        names=["id", "a", "k", "h", "u", "f", "d", "p", "q", "s"])


# (sql, extension mask): 1 = operators, 2 = Kleene AND / OR
VALUE_CASES = [
    ("a - 7", 1), ("7 - a", 1), ("k - a", 1), ("f - 1.5", 1), ("d - f", 1), ("h - h", 1), ("u - u", 1), ("id - a - k", 1),
    ("-a", 1), ("-k", 1), ("-h", 1), ("-f", 1), ("-d", 1), ("+a", 1), ("-(a + 1) * 2", 1), ("- -a", 1), ("-5", 1), ("-2.5", 1),
    ("not p", 1), ("not (a > 0)", 1), ("not a", 1), ("not (p and q)", 1), ("not true", 1),
    ("a is null", 1), ("a is not null", 1), ("s is null", 1), ("s is not null", 1), ("p is null", 1), ("(a + k) is null", 1),
    ("(f > 0.0) is not null", 1), ("id is null", 1), ("5 is null", 1),
    ("p and q", 2), ("p or q", 2), ("(a > 0) and q", 2), ("p or (f > 0.0)", 2), ("(a is null) or q", 3),
    ("(not p) or (a is null and q)", 3), ("a - 1 > 0 and s is not null", 3), ("not (p or q) and -a < 3", 3),
]
FILTER_CASES = [
    ("a is not null and a - 500 < 0", 1), ("not p", 1), ("s is null or -f > 2.0", 1), ("p or q", 2), ("p and (q or a > 0)", 2),
    ("not (p and q) or k - a > 0", 3), ("d is null", 1), ("id - 1000 > 0 and not (s = 's7')", 1),
]
ERROR_CASES = [   # (sql, mask, error kind)
    ("-u", 1, "InvalidArgumentError"), ("a - s", 1, "UnsupportedTypeCoersion"), ("-s", 1, "InvalidArgumentError"),
    ("a - 7", 0, "BinaryOperatorNotImplemented"), ("-a", 0, "ExpressionTypeNotImplemented"), ("not p", 0, "ExpressionTypeNotImplemented"),
    ("a is null", 0, "ExpressionTypeNotImplemented"), ("a - 7", 2, "BinaryOperatorNotImplemented"),
]
