"""chapterhouseqe_b200 -- B200-native filter / projection / compaction for ChapterhouseDB.

Python host-side mirror of the reference's `record_utils` module
(src/handlers/operator_handler/operators/record_utils/mod.rs:13-15):

    filter_record(rec, table_aliases, expr)        -> RecordBatch   (filter_record.rs:21-39)
    project_record(fields, record, table_aliases)  -> RecordBatch   (record_projection.rs:16-76)
    compute_value(rec, table_aliases, expr)        -> ArrayDatum    (compute_value.rs:57-344)

Everything goes through the C ABI declared in include/chdb_gpu.h (libchdb_gpu.so: hand-written
sm_100a CUDA kernels).  There is no CPU fallback: without the built library or without a CUDA
device these calls raise.
"""
from .api import (EXT_KLEENE, EXT_OPERATORS, sql_extensions,  # noqa: F401
                  ArrayDatum, ChdbError, Context, DeviceBatch, DeviceBatchList, ParquetFile, ParquetImage, Pending, Program, RecordPool,  # noqa: F401
                  compute_value,
                  default_context, encode_parquet, filter_project_record, filter_record, get_record_table_aliases, jit_available,
                  lib_path, load_library, project_record)
from . import sqlparser_lite  # noqa: F401

__all__ = ["EXT_KLEENE", "EXT_OPERATORS", "sql_extensions", "ArrayDatum", "ChdbError", "Context", "DeviceBatch", "DeviceBatchList", "ParquetFile", "ParquetImage", "Pending", "Program", "RecordPool", "compute_value",
           "default_context", "encode_parquet", "filter_project_record", "filter_record", "get_record_table_aliases", "jit_available",
           "lib_path", "load_library", "project_record", "sqlparser_lite"]
