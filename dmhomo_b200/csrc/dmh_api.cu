// Library-level entry points and process-wide state of libdmhomo.
#include "dmh_common.cuh"

namespace dmh {
thread_local char g_last_error[512] = "";
std::atomic<uint64_t> g_launches{0};
thread_local const char* g_last_kernel = "";
}  // namespace dmh

extern "C" int dmh_version(void) { return DMH_ABI_VERSION; }
extern "C" const char* dmh_last_error_string(void) { return dmh::g_last_error; }
extern "C" uint64_t dmh_launch_count(void) { return dmh::g_launches.load(std::memory_order_relaxed); }
extern "C" const char* dmh_last_kernel_name(void) { return dmh::g_last_kernel; }
