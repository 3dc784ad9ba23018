"""Golden vectors transcribed from the reference's own unit tests
(/root/reference/src/handlers/operator_handler/operators/record_utils/test_*.rs).

Each case: (name, citation, schema [(name, dtype, nullable)], columns {.. python lists ..},
table_aliases, kind, query, expected).  `kind`:
  "value"   -> compute_value(expr)   expected = (dtype, [values])
  "filter"  -> filter_record(expr)   expected = {col: [values]}
Shared by the oracle tests (CPU) and the CUDA parity tests (GPU) so both are pinned to
the same reference-authored answers.
"""

GOLDEN = [
    ("test_add", "test_compute_value.rs:12-36",
     [("cost", "int32", False)], {"cost": [10, 20, 30, 40, 50]}, [],
     "value", "cost + 7", ("int32", [17, 27, 37, 47, 57])),
    ("test_eq", "test_compute_value.rs:39-63",
     [("cost", "int32", False)], {"cost": [10, 20, 30, 40, 50]}, [],
     "value", "cost = 20", ("bool", [False, True, False, False, False])),
    ("test_and_scalar_with_array", "test_compute_value.rs:66-89",
     [("is_good", "bool", False)], {"is_good": [True, False, True, False, True]}, [],
     "value", "is_good = true", ("bool", [True, False, True, False, True])),
    ("test_and_array_with_array", "test_compute_value.rs:92-124",
     [("is_tall", "bool", False), ("is_rich", "bool", False)],
     {"is_tall": [True, False, True, False, True], "is_rich": [False, True, True, True, False]}, [],
     "value", "is_tall = is_rich", ("bool", [False, False, True, False, False])),
    ("test_complex_expression", "test_compute_value.rs:127-175",
     [("a", "float32", False), ("b", "float32", False), ("c", "float32", False)],
     {"a": [0., 1., 2., 3.], "b": [1., 1., 3., 2.], "c": [0., 0., 1., 2.]}, [],
     "value", "a+1.0/(2.0+c)*b", ("float32", [0.5, 1.5, 3.0, 3.5])),
    ("test_string_equals_array_with_scalar", "test_compute_value.rs:178-201",
     [("text", "utf8", False)], {"text": ["hello", ", ", "world", "!"]}, [],
     "value", "text = 'world'", ("bool", [False, False, True, False])),
    ("test_string_not_equals_array_with_scalar", "test_compute_value.rs:204-227",
     [("text", "utf8", False)], {"text": ["hello", ", ", "world", "!"]}, [],
     "value", "text <> 'world'", ("bool", [True, True, False, True])),
    ("test_filter_simple_record", "test_filter_record.rs:12-39",
     [("cost", "int32", False)], {"cost": [10, 20, 30, 40, 50]}, [[]],
     "filter", "cost < 30", {"cost": [10, 20]}),
]

# test_table_alias (test_compute_value.rs:230-272): three columns all named "text"
TABLE_ALIAS_CASE = dict(
    schema=[("text", "utf8", False), ("text", "utf8", False), ("text", "utf8", False)],
    columns=[["hello", ", ", "world", "!"], ["a", "b", "c", "d"], ["hello", ", ", "world", "!"]],
    table_aliases=[["table_a"], ["table_b"], ["table_c"]],
    query="table_b.text = 'c'",
    expected=("bool", [False, False, True, False]),
)

# test_arrow_compute_behavior.rs:111-126 (u32 -> f32 round-to-nearest-even)
U32_TO_F32 = ([1, 1000, 16_777_216, 16_777_217, 4_294_967_295],
              [1.0, 1000.0, 16777216.0, 16777216.0, 4294967296.0])

# sample_queries/*.sql of the reference (statement text only; data is random in the reference)
SAMPLE_QUERIES = {
    "simple_q1": "select * from read_files('sample_data/simple/*.parquet') where id < 25",
    "simple_q2": "select * from read_files('sample_data/simple_wide_string/*.parquet') where id > 25",
    "simple_q3": "select id, value2 from read_files('sample_data/simple/*.parquet') where id < 75",
    "simple_q4": """select id, value1, id + 10.0 as id_plus_10, (value2 + 10) / 100 as value2,
                    1.0 / id as value3, 1.0 / (id * id) as value4, id * id as value5
                    from read_files('sample_data/simple/*.parquet') where id > 25 + 0.0""",
    "simple_q5": "select * from read_files('sample_data/simple/*.parquet') where id % 2 = 0",
    "readme": "select * from read_files('simple/*.parquet') where value2 > 10.0",
}
