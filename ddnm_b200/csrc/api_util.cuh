// try/catch wrappers for the C ABI
#pragma once
#include <string>

#include "common.cuh"

namespace ddnm {
extern thread_local std::string g_last_error;
}

#define DDNM_API_BEGIN try {
#define DDNM_API_END                                        \
  return 0;                                                 \
  }                                                         \
  catch (const std::exception& e) {                         \
    ::ddnm::g_last_error = e.what();                        \
    return 1;                                               \
  }                                                         \
  catch (...) {                                             \
    ::ddnm::g_last_error = "unknown C++ exception";         \
    return 2;                                               \
  }
