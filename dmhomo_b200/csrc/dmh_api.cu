// Library-level entry points and process-wide state of libdmhomo.
#include "dmh_common.cuh"

#include <cstring>

namespace dmh {
thread_local char g_last_error[512] = "";
std::atomic<uint64_t> g_launches{0};
thread_local const char* g_last_kernel = "";
Tuning& tuning() {
  static Tuning t;
  return t;
}
}  // namespace dmh

extern "C" int dmh_version(void) { return DMH_ABI_VERSION; }
extern "C" const char* dmh_last_error_string(void) { return dmh::g_last_error; }
extern "C" uint64_t dmh_launch_count(void) { return dmh::g_launches.load(std::memory_order_relaxed); }
extern "C" const char* dmh_last_kernel_name(void) { return dmh::g_last_kernel; }

namespace {
int* tuning_slot(const char* key) {
  if (!key) return nullptr;
  dmh::Tuning& t = dmh::tuning();
  if (!strcmp(key, "tile")) return &t.tile;
  if (!strcmp(key, "tile_interior")) return &t.tile_interior;
  if (!strcmp(key, "tile_flow")) return &t.tile_flow;
  if (!strcmp(key, "tile_wide")) return &t.tile_wide;
  if (!strcmp(key, "channels")) return &t.channels;
  if (!strcmp(key, "tile_pair_major")) return &t.tile_pair_major;
  if (!strcmp(key, "tile_dyn")) return &t.tile_dyn;
  if (!strcmp(key, "tile_chunk")) return &t.tile_chunk;
  return nullptr;
}
}  // namespace
extern "C" int dmh_set_tuning(const char* key, int value) {
  int* p = tuning_slot(key);
  if (!p) return dmh::fail(DMH_EINVAL, "set_tuning: unknown key '%s'", key ? key : "(null)");
  *p = value;
  return DMH_OK;
}
extern "C" int dmh_get_tuning(const char* key, int* value) {
  int* p = tuning_slot(key);
  if (!p || !value) return dmh::fail(DMH_EINVAL, "get_tuning: unknown key '%s'", key ? key : "(null)");
  *value = *p;
  return DMH_OK;
}
