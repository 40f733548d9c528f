// Library-wide state: last-error string, launch counter, device queries.
#include "common.cuh"
#include <stdlib.h>
#include <string.h>

namespace boa {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool thin_passes() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BOA_B200_THIN");
    v = (e && atoi(e) != 0) ? 1 : 0;
  }
  return v == 1;
}

int sm_count() {
  static int n = 0;
  if (n) return n;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int v = 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
  n = v;
  return n;
}

}  // namespace boa

extern "C" const char* boa_last_error(void) { return boa::g_err; }
extern "C" int boa_abi_version(void) { return 1; }
extern "C" uint64_t boa_kernel_launch_count(void) { return boa::g_launches.load(); }
