// Library-wide state of libopd_b200.so (error string, launch counter, version).
#include "opd_common.h"

namespace opd {
std::string& last_error_ref() {
  static thread_local std::string s;
  return s;
}
std::atomic<int64_t> g_launches{0};
}  // namespace opd

extern "C" {
int opd_version(void) { return OPD_ABI_VERSION; }
const char* opd_last_error(void) { return opd::last_error_ref().c_str(); }
int64_t opd_launch_count(void) { return opd::g_launches.load(); }
}
