// host/compat: tbb::global_control stand-in (caps the worker count of compat parallel_for). TEST INFRASTRUCTURE ONLY.
#pragma once
#include "parallel_for.h"

namespace tbb {
class global_control {
  public:
    enum parameter { max_allowed_parallelism, thread_stack_size };
    global_control(parameter p, std::size_t value) {
        if (p == max_allowed_parallelism && value >= 1) compat_detail::parallelism_cap() = static_cast<int>(value);
    }
};
} // namespace tbb
