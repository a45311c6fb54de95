// Kernel-launch accounting: `++g_kernel_launches` at every launch site counts the launch for the process and for the context
// whose C-ABI call is running on this host thread (vpin_kernel_launches(ctx) reports the latter; bench.py's gpu_launches).
#pragma once
#include <atomic>
#include <cstdint>

namespace vpin {

struct LaunchCounter {
  std::atomic<uint64_t> total{0};
  static std::atomic<uint64_t> *&current() {  // the running context's counter (set on entry to every C-ABI call)
    static thread_local std::atomic<uint64_t> *p = nullptr;
    return p;
  }
  LaunchCounter &operator++() {
    total.fetch_add(1, std::memory_order_relaxed);
    if (std::atomic<uint64_t> *c = current()) c->fetch_add(1, std::memory_order_relaxed);
    return *this;
  }
  uint64_t load() const { return total.load(); }
};
extern LaunchCounter g_kernel_launches;

}  // namespace vpin
