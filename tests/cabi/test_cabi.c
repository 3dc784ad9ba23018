/* test_cabi.c -- drives the C ABI of include/chdb_gpu.h without Python: what a Rust / C host does.
 *
 * Builds a 10 000-row RecordBatch (id Int32, value2 Float32 with nulls) by hand as Arrow C Data Interface
 * structs, then checks
 *   1. chdb_filter_record_expr (the reference's one-shot signature, filter_record.rs:21-25) against a scalar C loop;
 *   2. a compiled program through upload -> chdb_run_device -> chdb_download;
 *   3. chdb_run_device_many over 16 records, one output per record;
 *   4. chdb_filter_record_async + chdb_poll (never blocks) + chdb_pending_result;
 *   5. the error convention: ColumnNotFound at compile time, DivideByZero raised on the device.
 * Exit code 0 = all checks passed; 77 = no CUDA device (the library has no CPU fallback).
 * Compiled by __graft_entry__.build() (gcc, links libchdb_gpu.so); run by tests/test_gpu_cabi.py.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "chdb_gpu.h"

#define N 10000
#define CHECK(cond, ...)                                 \
  do {                                                   \
    if (!(cond)) {                                       \
      fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); \
      fprintf(stderr, __VA_ARGS__);                      \
      fprintf(stderr, "\n");                             \
      exit(1);                                           \
    }                                                    \
  } while (0)

static void noop_release_schema(struct ArrowSchema* s) { s->release = NULL; }
static void noop_release_array(struct ArrowArray* a) { a->release = NULL; }

/* sqlparser 0.52 serde JSON of:  id % 2 = 0 and value2 > 10.0   /   id / 0 = 1 */
static const char* kPredicate =
    "{\"BinaryOp\":{\"left\":{\"BinaryOp\":{\"left\":{\"BinaryOp\":{\"left\":{\"Identifier\":{\"value\":\"id\",\"quote_style\":null}},"
    "\"op\":\"Modulo\",\"right\":{\"Value\":{\"Number\":[\"2\",false]}}}},\"op\":\"Eq\",\"right\":{\"Value\":{\"Number\":[\"0\",false]}}}},"
    "\"op\":\"And\",\"right\":{\"BinaryOp\":{\"left\":{\"Identifier\":{\"value\":\"value2\",\"quote_style\":null}},\"op\":\"Gt\","
    "\"right\":{\"Value\":{\"Number\":[\"10.0\",false]}}}}}}";
static const char* kDivZero =
    "{\"BinaryOp\":{\"left\":{\"BinaryOp\":{\"left\":{\"Identifier\":{\"value\":\"id\",\"quote_style\":null}},\"op\":\"Divide\","
    "\"right\":{\"Value\":{\"Number\":[\"0\",false]}}}},\"op\":\"Eq\",\"right\":{\"Value\":{\"Number\":[\"1\",false]}}}}";
static const char* kMissing =
    "{\"BinaryOp\":{\"left\":{\"Identifier\":{\"value\":\"nope\",\"quote_style\":null}},\"op\":\"Eq\",\"right\":{\"Value\":{\"Number\":[\"1\",false]}}}}";

struct Batch {
  int32_t* id;
  float* v2;
  uint8_t* v2_valid;
  struct ArrowSchema schema, fields[2];
  struct ArrowSchema* field_ptrs[2];
  struct ArrowArray array, cols[2];
  struct ArrowArray* col_ptrs[2];
  const void* top_buffers[1];
  const void* id_buffers[2];
  const void* v2_buffers[2];
};

static void make_batch(struct Batch* b, int n, int first_id) {
  memset(b, 0, sizeof(*b));
  b->id = (int32_t*)calloc((size_t)n + 16, 4);
  b->v2 = (float*)calloc((size_t)n + 16, 4);
  b->v2_valid = (uint8_t*)calloc((size_t)n / 8 + 16, 1);
  int64_t nulls = 0;
  for (int i = 0; i < n; i++) {
    b->id[i] = first_id + i;
    b->v2[i] = (float)((i * 37) % 1000) / 10.0f;
    if (i % 11 != 3) b->v2_valid[i >> 3] |= (uint8_t)(1u << (i & 7));
    else nulls++;
  }
  const char* names[2] = {"id", "value2"};
  const char* formats[2] = {"i", "f"};
  for (int c = 0; c < 2; c++) {
    b->fields[c].format = formats[c];
    b->fields[c].name = names[c];
    b->fields[c].flags = c == 1 ? ARROW_FLAG_NULLABLE : 0;
    b->fields[c].release = noop_release_schema;
    b->field_ptrs[c] = &b->fields[c];
    b->cols[c].length = n;
    b->cols[c].n_buffers = 2;
    b->cols[c].release = noop_release_array;
    b->col_ptrs[c] = &b->cols[c];
  }
  b->id_buffers[0] = NULL;
  b->id_buffers[1] = b->id;
  b->v2_buffers[0] = b->v2_valid;
  b->v2_buffers[1] = b->v2;
  b->cols[0].buffers = b->id_buffers;
  b->cols[0].null_count = 0;
  b->cols[1].buffers = b->v2_buffers;
  b->cols[1].null_count = nulls;
  b->schema.format = "+s";
  b->schema.name = "";
  b->schema.n_children = 2;
  b->schema.children = b->field_ptrs;
  b->schema.release = noop_release_schema;
  b->top_buffers[0] = NULL;
  b->array.length = n;
  b->array.n_buffers = 1;
  b->array.buffers = b->top_buffers;
  b->array.n_children = 2;
  b->array.children = b->col_ptrs;
  b->array.release = noop_release_array;
}

static void free_batch(struct Batch* b) {
  free(b->id);
  free(b->v2);
  free(b->v2_valid);
}

/* the reference's semantics in scalar C: keep rows where (id % 2 == 0) && valid(value2) && value2 > 10.0f */
static int expected_rows(const struct Batch* b, int n, int32_t* ids, float* v2s) {
  int k = 0;
  for (int i = 0; i < n; i++) {
    const int valid = (b->v2_valid[i >> 3] >> (i & 7)) & 1;
    if (b->id[i] % 2 == 0 && valid && b->v2[i] > 10.0f) {
      ids[k] = b->id[i];
      v2s[k] = b->v2[i];
      k++;
    }
  }
  return k;
}

static void check_output(const char* what, struct ArrowArray* out, struct ArrowSchema* os, const struct Batch* in, int n) {
  int32_t* ids = (int32_t*)malloc((size_t)n * 4);
  float* v2s = (float*)malloc((size_t)n * 4);
  const int k = expected_rows(in, n, ids, v2s);
  CHECK(out->length == k, "%s: %lld rows, expected %d", what, (long long)out->length, k);
  CHECK(out->n_children == 2 && os->n_children == 2, "%s: column count", what);
  CHECK(strcmp(os->children[0]->name, "id") == 0 && strcmp(os->children[1]->name, "value2") == 0, "%s: names", what);
  CHECK(strcmp(os->children[0]->format, "i") == 0 && strcmp(os->children[1]->format, "f") == 0, "%s: formats", what);
  CHECK((os->children[1]->flags & ARROW_FLAG_NULLABLE) != 0 && (os->children[0]->flags & ARROW_FLAG_NULLABLE) == 0, "%s: nullability", what);
  const int32_t* got_id = (const int32_t*)out->children[0]->buffers[1];
  const float* got_v2 = (const float*)out->children[1]->buffers[1];
  CHECK(out->children[1]->null_count == 0, "%s: a NULL value2 survived value2 > 10.0", what);
  CHECK(k == 0 || memcmp(got_id, ids, (size_t)k * 4) == 0, "%s: id column differs", what);
  CHECK(k == 0 || memcmp(got_v2, v2s, (size_t)k * 4) == 0, "%s: value2 column differs (bitwise)", what);
  free(ids);
  free(v2s);
  out->release(out);
  os->release(os);
}

int main(void) {
  chdb_status st;
  chdb_ctx* ctx = NULL;
  int rc = chdb_ctx_create(0, &ctx, &st);
  if (rc == CHDB_ERR_CUDA) {
    fprintf(stderr, "no CUDA device: %s\n", st.message);
    return 77;
  }
  CHECK(rc == 0, "ctx_create: %s", st.message);
  CHECK(strcmp(chdb_compiled_arch(), "sm_100a") == 0, "arch %s", chdb_compiled_arch());

  struct Batch b;
  make_batch(&b, N, 0);
  struct ArrowArray out;
  struct ArrowSchema os;

  /* 1. the reference's one-shot signature */
  rc = chdb_filter_record_expr(ctx, &b.array, &b.schema, "[[],[]]", kPredicate, &out, &os, &st);
  CHECK(rc == 0, "filter_record_expr: %s: %s", chdb_code_name(rc), st.message);
  check_output("filter_record_expr", &out, &os, &b, N);

  /* 2. compiled program, device-resident batches */
  chdb_program* prog = NULL;
  rc = chdb_program_compile_filter(kPredicate, &b.schema, NULL, &prog, &st);
  CHECK(rc == 0, "compile_filter: %s", st.message);
  CHECK(chdb_program_num_instructions(prog) > 0, "empty program");
  chdb_device_batch *din = NULL, *dout = NULL;
  CHECK(chdb_upload(ctx, &b.array, &b.schema, &din, &st) == 0, "upload: %s", st.message);
  const long long launches0 = chdb_ctx_launch_count(ctx);
  CHECK(chdb_run_device(ctx, prog, din, &dout, &st) == 0, "run_device: %s", st.message);
  {
    /* zero kernel + select + gather (or zero + the fused stream kernel with CHDB_SPLIT=0) */
    const long long nl = chdb_ctx_launch_count(ctx) - launches0;
    CHECK(nl == 3 || nl == 2, "expected the zero kernel + the select and gather kernels, got %lld launches", nl);
  }
  CHECK(chdb_download(ctx, dout, &out, &os, &st) == 0, "download: %s", st.message);
  check_output("run_device", &out, &os, &b, N);
  chdb_device_batch_release(dout);

  /* 3. many records in one launch set */
  enum { R = 16 };
  struct Batch recs[R];
  chdb_device_batch* dins[R];
  chdb_device_batch* douts[R];
  for (int r = 0; r < R; r++) {
    make_batch(&recs[r], N - 123 * r, 1000000 * r);
    CHECK(chdb_upload(ctx, &recs[r].array, &recs[r].schema, &dins[r], &st) == 0, "upload %d: %s", r, st.message);
  }
  const long long launches1 = chdb_ctx_launch_count(ctx);
  CHECK(chdb_run_device_many(ctx, prog, (const chdb_device_batch* const*)dins, R, douts, &st) == 0, "run_device_many: %s", st.message);
  CHECK(chdb_ctx_launch_count(ctx) == launches1 + 2, "16 records must share one launch set");
  for (int r = 0; r < R; r++) {
    CHECK(chdb_download(ctx, douts[r], &out, &os, &st) == 0, "download %d: %s", r, st.message);
    check_output("run_device_many", &out, &os, &recs[r], N - 123 * r);
    chdb_device_batch_release(douts[r]);
    chdb_device_batch_release(dins[r]);
    free_batch(&recs[r]);
  }

  /* 4. non-blocking host call */
  chdb_pending* pend = NULL;
  CHECK(chdb_filter_record_async(ctx, prog, &b.array, &b.schema, &pend, &st) == 0, "filter_record_async: %s", st.message);
  long polls = 0;
  for (;;) {
    const int p = chdb_poll(pend, &st);
    CHECK(p >= 0, "poll: %s", st.message);
    if (p == 1) break;
    polls++;
    CHECK(polls < 200000000, "poll never became ready");
  }
  CHECK(chdb_pending_result(pend, &out, &os, &st) == 0, "pending_result: %s", st.message);
  check_output("filter_record_async", &out, &os, &b, N);
  chdb_pending_release(pend);

  /* 4b. materialize on the device: three filtered records -> one Parquet image (a row group per record with a 1-row limit),
   *     read back through the device decoder and compared with what the filter must have kept */
  {
    enum { M = 3 };
    struct Batch mrecs[M];
    chdb_device_batch *mins[M], *mouts[M], *back[M];
    for (int r = 0; r < M; r++) {
      make_batch(&mrecs[r], N - 77 * r, 500 * r);
      CHECK(chdb_upload(ctx, &mrecs[r].array, &mrecs[r].schema, &mins[r], &st) == 0, "upload %d: %s", r, st.message);
    }
    CHECK(chdb_run_device_many(ctx, prog, (const chdb_device_batch* const*)mins, M, mouts, &st) == 0, "run_device_many: %s", st.message);
    void* image = NULL;
    int64_t image_len = 0;
    int32_t consumed = 0, groups = 0;
    CHECK(chdb_parquet_encode(ctx, (const chdb_device_batch* const*)mouts, M, 1, 0, &image, &image_len, &consumed, &groups, &st) == 0,
          "parquet_encode: %s", st.message);
    CHECK(consumed == M && groups == M && image_len > 12 && memcmp(image, "PAR1", 4) == 0 &&
              memcmp((const char*)image + image_len - 4, "PAR1", 4) == 0,
          "parquet image: consumed %d, row groups %d, %lld bytes", consumed, groups, (long long)image_len);
    chdb_parquet* pf = NULL;
    CHECK(chdb_parquet_open(image, image_len, &pf, &st) == 0, "parquet_open of the written image: %s", st.message);
    CHECK(chdb_parquet_num_row_groups(pf) == M && chdb_parquet_num_columns(pf) == 2, "written image: row groups / columns");
    CHECK(chdb_parquet_decode_row_groups(ctx, pf, 0, M, back, &st) == 0, "decode of the written image: %s", st.message);
    for (int r = 0; r < M; r++) {
      CHECK(chdb_download(ctx, back[r], &out, &os, &st) == 0, "download %d: %s", r, st.message);
      check_output("encode -> decode", &out, &os, &mrecs[r], N - 77 * r);
      chdb_device_batch_release(back[r]);
      chdb_device_batch_release(mouts[r]);
      chdb_device_batch_release(mins[r]);
      free_batch(&mrecs[r]);
    }
    chdb_parquet_close(pf);
    chdb_parquet_image_free(image);
  }

  /* 5. errors: the reference's kinds */
  chdb_program* bad = NULL;
  rc = chdb_program_compile_filter(kMissing, &b.schema, NULL, &bad, &st);
  CHECK(rc == CHDB_ERR_COLUMN_NOT_FOUND, "expected ColumnNotFound, got %s", chdb_code_name(rc));
  rc = chdb_filter_record_expr(ctx, &b.array, &b.schema, NULL, kDivZero, &out, &os, &st);
  CHECK(rc == CHDB_ERR_DIVIDE_BY_ZERO, "expected DivideByZero, got %s (%s)", chdb_code_name(rc), st.message);

  chdb_device_batch_release(din);
  chdb_program_release(prog);
  chdb_ctx_destroy(ctx);
  free_batch(&b);
  printf("test_cabi: all checks passed (%ld polls before the async result was ready)\n", polls);
  return 0;
}
