// Library-level entry points: version, error string, device check.
#include <cstdlib>

#include "common.cuh"

namespace ef {

char* last_error_buf() {
  static thread_local char buf[512] = "";
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

static unsigned long long g_launches = 0;  // kernels launched by this library in this process (bench.py: gpu_launches)

int check_launch(const char* what) {
  __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED);
  const cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return EF_OK;
  return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
}

static int g_pdl = -1;  // -1 = not decided yet (EF_PDL environment variable, default on)
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("EF_PDL");
    g_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl != 0;
}

}  // namespace ef

extern "C" int ef_debug_pdl(int on) {  // programmatic dependent launch of the forward kernels of a model step (default on)
  ef::g_pdl = on ? 1 : 0;
  return EF_OK;
}

extern "C" int ef_version(void) { return EF_VERSION; }
extern "C" uint64_t ef_launch_count(void) { return __atomic_load_n(&ef::g_launches, __ATOMIC_RELAXED); }
extern "C" const char* ef_last_error(void) { return ef::last_error_buf(); }

extern "C" int ef_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return ef::fail(-1, "ef_device_ok: no CUDA device");
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return ef::fail(-1, "ef_device_ok: attribute query failed");
  return major == 10 ? 1 : 0;
}
