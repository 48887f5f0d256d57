// libkgb200: error reporting, device selection, version.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace kgb {

static thread_local char g_err[512] = "";
static thread_local int g_dev = -1;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int use_device(int device) {
  if (device < 0) {
    set_error("negative device ordinal %d", device);
    return KGB_ERR_INVALID;
  }
  if (g_dev != device) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
      set_error("cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
      return KGB_ERR_CUDA;
    }
    g_dev = device;
  }
  return KGB_OK;
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

static int g_sm[64];
static int g_sm_ok[64];

int sm_count(int device) {
  if (device >= 0 && device < 64 && g_sm_ok[device]) return g_sm[device];
  int n = 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) n = 148;
  if (device >= 0 && device < 64) {
    g_sm[device] = n;
    g_sm_ok[device] = 1;
  }
  return n;
}

}  // namespace kgb

extern "C" {

int kgb_version(void) { return 204; /* ABI version: keras_geometric_b200/_lib.py:ABI_VERSION must match */ }

const char* kgb_last_error(void) { return kgb::g_err; }

int64_t kgb_launch_count(void) { return (int64_t)__atomic_load_n(&kgb::g_launches, __ATOMIC_RELAXED); }

int kgb_window_alloc(int device, size_t bytes, void** ptr) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(ptr != nullptr && bytes > 0, "bad arguments");
  KGB_CHECK_CUDA(cudaMalloc(ptr, bytes));   // plain cudaMalloc: exportable with cudaIpcGetMemHandle
  return KGB_OK;
}

int kgb_window_free(int device, void* ptr) {
  KGB_USE_DEVICE(device);
  if (ptr) KGB_CHECK_CUDA(cudaFree(ptr));
  return KGB_OK;
}

int kgb_ipc_export(int device, const void* ptr, unsigned char handle[64]) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(ptr != nullptr && handle != nullptr, "NULL pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  cudaIpcMemHandle_t h;
  KGB_CHECK_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
  memcpy(handle, &h, 64);
  return KGB_OK;
}

int kgb_ipc_open(int device, const unsigned char handle[64], void** ptr) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(ptr != nullptr && handle != nullptr, "NULL pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  KGB_CHECK_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return KGB_OK;
}

int kgb_ipc_close(int device, void* ptr) {
  KGB_USE_DEVICE(device);
  if (ptr) KGB_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return KGB_OK;
}

int kgb_device_info(int device, int* sms, int* cc_major, int* cc_minor, int64_t* l2_bytes) {
  KGB_USE_DEVICE(device);
  int v = 0;
  KGB_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
  if (sms) *sms = v;
  KGB_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, device));
  if (cc_major) *cc_major = v;
  KGB_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, device));
  if (cc_minor) *cc_minor = v;
  KGB_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, device));
  if (l2_bytes) *l2_bytes = v;
  return KGB_OK;
}

}  // extern "C"
