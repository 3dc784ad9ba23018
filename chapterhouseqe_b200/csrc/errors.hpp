#pragma once
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

#include "../../include/chdb_gpu.h"

namespace chdb {

// Internal exception; converted to (code, chdb_status) at the C boundary.
struct Error : std::runtime_error {
  int32_t code;
  Error(int32_t c, const std::string& msg) : std::runtime_error(msg), code(c) {}
};

inline int32_t set_status(chdb_status* st, int32_t code, const char* msg) {
  if (st) {
    st->code = code;
    std::snprintf(st->message, sizeof(st->message), "%s", msg ? msg : "");
  }
  return code;
}
inline int32_t set_ok(chdb_status* st) { return set_status(st, CHDB_OK, ""); }

// Runs f(), mapping exceptions to status codes.  Nothing may unwind across the C ABI.
template <class F>
int32_t guarded(chdb_status* st, F&& f) {
  try {
    f();
    return set_ok(st);
  } catch (const Error& e) {
    return set_status(st, e.code, e.what());
  } catch (const std::bad_alloc&) {
    return set_status(st, CHDB_ERR_CUDA, "out of host memory");
  } catch (const std::exception& e) {
    return set_status(st, CHDB_ERR_INVALID_ARGUMENT, e.what());
  }
}

}  // namespace chdb
