"""Writes tests/golden/reference_vectors.json: the golden vectors of the reference's own unit tests for this path
(record_utils/test_compute_value.rs, test_filter_record.rs, test_arrow_compute_behavior.rs) as plain data.

The reference is Rust and cannot be built in this image, so the vectors are transcribed by hand in tests/refcases.py
(each with its file:line citation); this script only serialises them, so that non-Python checkers (tests/cabi, a Rust
maintainer's own test) can read the same answers.  Usage: python tests/golden/make_reference_vectors.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refcases  # noqa: E402

cases = []
for name, cite, schema, cols, aliases, kind, query, expected in refcases.GOLDEN:
    cases.append({"name": name, "reference": cite, "schema": [list(f) for f in schema], "columns": cols,
                  "table_aliases": aliases, "kind": kind, "query": query,
                  "expected": ({"dtype": expected[0], "values": expected[1]} if kind == "value" else expected)})
t = refcases.TABLE_ALIAS_CASE
cases.append({"name": "test_table_alias", "reference": "test_compute_value.rs:230-272", "schema": [list(f) for f in t["schema"]],
              "columns_in_order": t["columns"], "table_aliases": t["table_aliases"], "kind": "value", "query": t["query"],
              "expected": {"dtype": t["expected"][0], "values": t["expected"][1]}})
doc = {"source": "alekLukanen/ChapterhouseQE src/handlers/operator_handler/operators/record_utils/test_*.rs (transcribed)",
       "cases": cases,
       "u32_to_f32_round_to_nearest_even": {"reference": "test_arrow_compute_behavior.rs:111-126",
                                            "input": refcases.U32_TO_F32[0], "expected": refcases.U32_TO_F32[1]}}
with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:
    json.dump(doc, f, indent=1)
print(len(cases), "cases")
