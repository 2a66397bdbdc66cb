// No C++ exception may cross the C ABI (include/ruvgnn.h: "never panics/aborts across the ABI").
// Every int32_t entry point wraps its body in RUVGNN_ABI_TRY ... RUVGNN_ABI_CATCH: a failed host
// allocation becomes RUVGNN_ERR_OOM, anything else RUVGNN_ERR_INVALID_INPUT with the message.
#pragma once
#include <exception>
#include <new>
#include <string>

#include "../../include/ruvgnn.h"

namespace ruvgnn {
int32_t fail(int32_t code, const std::string &msg);
}

#define RUVGNN_ABI_TRY try {
#define RUVGNN_ABI_CATCH                                                                      \
  }                                                                                           \
  catch (const std::bad_alloc &) {                                                            \
    return ::ruvgnn::fail(RUVGNN_ERR_OOM, "host allocation failed");                          \
  }                                                                                           \
  catch (const std::exception &e) {                                                           \
    return ::ruvgnn::fail(RUVGNN_ERR_INVALID_INPUT, std::string("internal error: ") + e.what()); \
  }                                                                                           \
  catch (...) {                                                                               \
    return ::ruvgnn::fail(RUVGNN_ERR_INVALID_INPUT, "internal error");                        \
  }
