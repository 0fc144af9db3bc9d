// Library-level entry points of libfemb200: ABI version, error text, launch counter, device check.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "femb_common.cuh"

namespace femb {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char* what) {
  const cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return FEMB_E_CUDA;
  }
  return FEMB_OK;
}

}  // namespace femb

extern "C" {

int femb_abi_version(void) { return FEMB_ABI_VERSION; }

const char* femb_last_error(void) { return femb::g_err; }

long long femb_launch_count(void) { return femb::g_launches.load(std::memory_order_relaxed); }

int femb_check_device(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
    cudaGetLastError();
    femb::set_error("femb_check_device: no CUDA device %d", device);
    return FEMB_E_NODEVICE;
  }
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, device) != cudaSuccess) {
    cudaGetLastError();
    femb::set_error("femb_check_device: cannot query device %d", device);
    return FEMB_E_CUDA;
  }
  if (p.major != 10) {
    femb::set_error("femb_check_device: device %d is sm_%d%d, libfemb200 is built for sm_100a only", device, p.major, p.minor);
    return FEMB_E_NODEVICE;
  }
  return FEMB_OK;
}

}  // extern "C"
