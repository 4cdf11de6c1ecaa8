// Error string + device helpers shared by every translation unit.
#include "common.cuh"
#include "../../include/ifcb_b200.h"
#include <cstring>

namespace ifcb {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

int sm_count() {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  cached = n;
  return n;
}
}  // namespace ifcb

extern "C" int ifcb_abi_version(void) { return IFCB_B200_ABI_VERSION; }
extern "C" const char* ifcb_last_error(void) { return ifcb::get_error(); }
extern "C" int ifcb_sm_count(void) { return ifcb::sm_count(); }
