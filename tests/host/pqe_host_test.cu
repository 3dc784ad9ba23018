// Host-only check of the Parquet encoder's Thrift writer and file layout (csrc/parquet_encode.inc: lay_out / footer_bytes): builds a
// two-row-group file by hand -- page payloads packed on the host exactly as pqe_write packs them on the device -- so that the CPU
// test suite can read it back with pyarrow and with this library's own footer reader.  No kernel is launched.
#include "../../chapterhouseqe_b200/csrc/runtime.cu"
#include <fstream>
using namespace chdb; using namespace chdb::pqe;
int main(int argc, char** argv) {
  const char* out_path = argc > 1 ? argv[1] : "pqe_host_test.parquet";
  const int64_t n = 21;  // not a multiple of 8
  std::vector<ColumnPlan> cols(3);
  cols[0].type = T_I32; cols[0].name = "id"; cols[0].physical = 1; cols[0].w_out = 4;
  cols[1].type = T_F64; cols[1].name = "d"; cols[1].physical = 5; cols[1].w_out = 8; cols[1].optional = true;
  cols[2].type = T_UTF8; cols[2].name = "s"; cols[2].physical = 6; cols[2].converted = 0; cols[2].optional = true;
  std::vector<GroupPlan> groups(2);
  std::vector<std::vector<std::vector<uint8_t>>> payload(2);
  for (int gi = 0; gi < 2; gi++) {
    auto& g = groups[gi]; g.rows = 0; g.pages.assign(3, {});
    payload[gi].resize(3);
    for (int b = 0; b < 2; b++) {   // two batches (pages) per group
      g.rows += n;
      Page p0; p0.n = n; p0.values_bytes = n * 4;
      for (int i = 0; i < n; i++) { int32_t v = gi * 1000 + b * 100 + i; payload[gi][0].insert(payload[gi][0].end(), (uint8_t*)&v, (uint8_t*)&v + 4); }
      g.pages[0].push_back(p0);
      Page p1; p1.n = n; p1.levels_bitmap = true;   // every third row null
      std::vector<uint8_t> bm((n + 7) / 8, 0); int valid = 0;
      for (int i = 0; i < n; i++) if (i % 3) { bm[i / 8] |= 1 << (i % 8); valid++; }
      p1.values_bytes = valid * 8;
      payload[gi][1].insert(payload[gi][1].end(), bm.begin(), bm.end());
      for (int i = 0; i < n; i++) if (i % 3) { double v = i * 0.5 + b; payload[gi][1].insert(payload[gi][1].end(), (uint8_t*)&v, (uint8_t*)&v + 8); }
      g.pages[1].push_back(p1);
      Page p2; p2.n = n; p2.levels_bitmap = false;   // optional, all valid: RLE run
      std::vector<uint8_t> sv;
      for (int i = 0; i < n; i++) { std::string t = "row" + std::to_string(i * 7 + b); uint32_t l = t.size(); sv.insert(sv.end(), (uint8_t*)&l, (uint8_t*)&l + 4); sv.insert(sv.end(), t.begin(), t.end()); }
      p2.values_bytes = sv.size();
      payload[gi][2].insert(payload[gi][2].end(), sv.begin(), sv.end());
      g.pages[2].push_back(p2);
    }
  }
  std::vector<EncJob> jobs; std::vector<int64_t> totals;
  const int64_t at = lay_out(cols, groups, jobs, totals);
  size_t written = 0;
  auto footer = footer_bytes(cols, groups, &written);
  std::vector<uint8_t> img(at + footer.size() + 8, 0);
  memcpy(img.data(), "PAR1", 4);
  for (int gi = 0; gi < 2; gi++) for (int c = 0; c < 3; c++) {
    size_t used = 0;
    for (auto& pg : groups[gi].pages[c]) {
      memcpy(img.data() + pg.at, pg.head.data(), pg.head.size());
      const size_t body = pg.levels_payload + pg.values_bytes;
      memcpy(img.data() + pg.at + pg.head.size(), payload[gi][c].data() + used, body);
      used += body;
    }
  }
  memcpy(img.data() + at, footer.data(), footer.size());
  uint32_t fl = footer.size(); memcpy(img.data() + at + footer.size(), &fl, 4);
  memcpy(img.data() + img.size() - 4, "PAR1", 4);
  std::ofstream(out_path, std::ios::binary).write((const char*)img.data(), img.size());
  // our own footer reader
  auto meta = chdb::pq::read_footer(img.data(), (int64_t)img.size());
  printf("own reader: %zu row groups, %lld rows, %zu cols\n", meta.rgs.size(), (long long)meta.num_rows, meta.cols.size());
  return 0;
}
