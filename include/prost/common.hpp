// prost/common.hpp -- shared plumbing of the C++ drop-in layer.
//
// The classes in include/prost/ keep the names and signatures of tum-vision/prost's C++ API
// (reference: include/prost/*.hpp) but are thin RAII wrappers over the C ABI in prost_b200.h;
// all device work happens in libprost_b200.so.  Only T = float is supported (north star).
#ifndef PROST_COMMON_HPP_
#define PROST_COMMON_HPP_

#include <sys/types.h>

#include <cstddef>
#include <cstdint>
#include <functional>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>

#include "prost/exception.hpp"
#include "prost_b200.h"

namespace prost {

using std::cout;
using std::endl;
using std::function;
using std::list;
using std::map;
using std::shared_ptr;
using std::string;
using std::stringstream;
using std::vector;

inline string get_version() { return pb_version(); }

namespace detail {

template <typename T>
struct require_float {
  static_assert(std::is_same<T, float>::value,
                "prost_b200 implements the float (`real`) instantiation of the prost API only");
};

inline void check(int status) {
  if (status != PB_OK) throw Exception(pb_last_error());
}

// One process-wide context on the selected GPU (reference: cudaSetDevice(current_gpu_device) in
// matlab/+prost/private/prost.cpp:56,69).  set_gpu() before the first object is created.
struct ContextHolder {
  pb_context* ctx = nullptr;
  int device = 0;
  ~ContextHolder() { if (ctx) pb_context_destroy(ctx); }
};
inline ContextHolder& holder() {
  static ContextHolder h;
  return h;
}
inline pb_context* context() {
  ContextHolder& h = holder();
  if (!h.ctx) check(pb_context_create(h.device, nullptr, &h.ctx));
  return h.ctx;
}

}  // namespace detail

/// Selects the GPU used by subsequently created objects (mex command "set_gpu", prost.cpp:299-303).
inline void set_gpu(int device) {
  detail::ContextHolder& h = detail::holder();
  if (h.ctx && h.device != device) { pb_context_destroy(h.ctx); h.ctx = nullptr; }
  h.device = device;
}

}  // namespace prost

#endif
