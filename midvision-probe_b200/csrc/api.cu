// api.cu -- library-level entry points: version, error string, device query.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void mv_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int mv_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  return dev % MV_MAX_DEVICES;
}

int mv_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

extern "C" {

int mv_version(void) { return 100; }  // 0.1.0

const char* mv_last_error(void) { return g_err; }

int mv_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, int* l2_bytes) {
  int v = 0;
  if (sm_count) {
    MV_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
    *sm_count = v;
  }
  if (cc_major) {
    MV_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, device));
    *cc_major = v;
  }
  if (cc_minor) {
    MV_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, device));
    *cc_minor = v;
  }
  if (l2_bytes) {
    MV_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, device));
    *l2_bytes = v;
  }
  return MV_OK;
}

}  // extern "C"
