// Minimal JSON reader for the serde_json output of sqlparser's AST (objects, arrays, strings,
// numbers, true/false/null).  Header-only; throws chdb::Error(CHDB_ERR_BAD_JSON).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "errors.hpp"

namespace chdb {

struct Json {
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  bool b = false;
  double num = 0;
  std::string str;                                  // String, or the literal text of a Number
  std::vector<Json> arr;                            // Array
  std::vector<std::pair<std::string, Json>> obj;    // Object, insertion order kept

  const Json* get(const std::string& key) const {
    if (kind != Object) return nullptr;
    for (auto& kv : obj)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  bool is_null() const { return kind == Null; }
};

class JsonParser {
 public:
  explicit JsonParser(const char* text) : p_(text ? text : "") {}
  Json parse() {
    Json v = value();
    ws();
    if (*p_) fail("trailing characters");
    return v;
  }

 private:
  const char* p_;
  [[noreturn]] void fail(const char* what) { throw Error(CHDB_ERR_BAD_JSON, std::string("bad JSON: ") + what); }
  void ws() {
    while (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r') ++p_;
  }
  Json value() {
    ws();
    Json v;
    switch (*p_) {
      case '{': {
        ++p_;
        v.kind = Json::Object;
        ws();
        if (*p_ == '}') { ++p_; return v; }
        for (;;) {
          ws();
          if (*p_ != '"') fail("expected object key");
          std::string k = string();
          ws();
          if (*p_ != ':') fail("expected ':'");
          ++p_;
          v.obj.emplace_back(std::move(k), value());
          ws();
          if (*p_ == ',') { ++p_; continue; }
          if (*p_ == '}') { ++p_; return v; }
          fail("expected ',' or '}'");
        }
      }
      case '[': {
        ++p_;
        v.kind = Json::Array;
        ws();
        if (*p_ == ']') { ++p_; return v; }
        for (;;) {
          v.arr.push_back(value());
          ws();
          if (*p_ == ',') { ++p_; continue; }
          if (*p_ == ']') { ++p_; return v; }
          fail("expected ',' or ']'");
        }
      }
      case '"':
        v.kind = Json::String;
        v.str = string();
        return v;
      case 't':
        if (p_[1] == 'r' && p_[2] == 'u' && p_[3] == 'e') { p_ += 4; v.kind = Json::Bool; v.b = true; return v; }
        fail("bad literal");
      case 'f':
        if (p_[1] == 'a' && p_[2] == 'l' && p_[3] == 's' && p_[4] == 'e') { p_ += 5; v.kind = Json::Bool; return v; }
        fail("bad literal");
      case 'n':
        if (p_[1] == 'u' && p_[2] == 'l' && p_[3] == 'l') { p_ += 4; return v; }
        fail("bad literal");
      default: {
        const char* s = p_;
        char* end = nullptr;
        v.num = std::strtod(s, &end);
        if (end == s) fail("unexpected character");
        v.kind = Json::Number;
        v.str.assign(s, (const char*)end);
        p_ = end;
        return v;
      }
    }
  }
  static void utf8_append(std::string& out, uint32_t cp) {
    if (cp < 0x80) out += char(cp);
    else if (cp < 0x800) { out += char(0xC0 | (cp >> 6)); out += char(0x80 | (cp & 0x3F)); }
    else if (cp < 0x10000) { out += char(0xE0 | (cp >> 12)); out += char(0x80 | ((cp >> 6) & 0x3F)); out += char(0x80 | (cp & 0x3F)); }
    else { out += char(0xF0 | (cp >> 18)); out += char(0x80 | ((cp >> 12) & 0x3F)); out += char(0x80 | ((cp >> 6) & 0x3F)); out += char(0x80 | (cp & 0x3F)); }
  }
  uint32_t hex4() {
    uint32_t v = 0;
    for (int i = 0; i < 4; i++) {
      char c = *p_++;
      v <<= 4;
      if (c >= '0' && c <= '9') v |= c - '0';
      else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
      else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
      else fail("bad \\u escape");
    }
    return v;
  }
  std::string string() {
    std::string out;
    ++p_;  // opening quote
    while (*p_ && *p_ != '"') {
      if (*p_ == '\\') {
        ++p_;
        switch (*p_++) {
          case '"': out += '"'; break;
          case '\\': out += '\\'; break;
          case '/': out += '/'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'n': out += '\n'; break;
          case 'r': out += '\r'; break;
          case 't': out += '\t'; break;
          case 'u': {
            uint32_t cp = hex4();
            if (cp >= 0xD800 && cp < 0xDC00 && p_[0] == '\\' && p_[1] == 'u') {
              p_ += 2;
              uint32_t lo = hex4();
              cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
            }
            utf8_append(out, cp);
            break;
          }
          default: fail("bad escape");
        }
      } else {
        out += *p_++;
      }
    }
    if (*p_ != '"') fail("unterminated string");
    ++p_;
    return out;
  }
};

}  // namespace chdb
