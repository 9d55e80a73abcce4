// Library-wide state of the C ABI: last-error text and the launch counter.
#include <stdarg.h>

#include <atomic>
#include <mutex>

#include "common.cuh"

namespace hcf {

static std::mutex g_err_mu;
static char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  std::lock_guard<std::mutex> lk(g_err_mu);
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace hcf

extern "C" int hcf_abi_version(void) { return HCF_ABI_VERSION; }

extern "C" const char* hcf_last_error(void) {
  static thread_local char copy[512];
  std::lock_guard<std::mutex> lk(hcf::g_err_mu);
  snprintf(copy, sizeof(copy), "%s", hcf::g_err);
  return copy;
}

extern "C" uint64_t hcf_launch_count(void) { return hcf::g_launches.load(); }
extern "C" void hcf_launch_count_reset(void) { hcf::g_launches.store(0); }
