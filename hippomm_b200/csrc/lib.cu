// Library plumbing: error strings, device checks.
#include "common.cuh"

#include <string.h>

namespace hippo {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

hippo_status cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return HIPPO_E_CUDA;
}

struct DevInfo { int valid; int major; int minor; int sms; };
static DevInfo g_dev[64];

static hippo_status dev_info(DevInfo** out) {
  int dev = 0;
  HIPPO_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) { set_error("device ordinal %d out of range", dev); return HIPPO_E_BADARG; }
  DevInfo* d = &g_dev[dev];
  if (!d->valid) {
    HIPPO_CUDA(cudaDeviceGetAttribute(&d->major, cudaDevAttrComputeCapabilityMajor, dev));
    HIPPO_CUDA(cudaDeviceGetAttribute(&d->minor, cudaDevAttrComputeCapabilityMinor, dev));
    HIPPO_CUDA(cudaDeviceGetAttribute(&d->sms, cudaDevAttrMultiProcessorCount, dev));
    d->valid = 1;
  }
  *out = d;
  return HIPPO_OK;
}

hippo_status check_arch() {
  DevInfo* d;
  hippo_status s = dev_info(&d);
  if (s != HIPPO_OK) return s;
  if (d->major != 10) {
    set_error("libhippo_b200 is built for sm_100a only; current device is sm_%d%d (no fallback path)",
              d->major, d->minor);
    return HIPPO_E_ARCH;
  }
  return HIPPO_OK;
}

int sm_count() {
  DevInfo* d;
  if (dev_info(&d) != HIPPO_OK) return 0;
  return d->sms;
}

}  // namespace hippo

extern "C" {

int32_t hippo_abi_version(void) { return HIPPO_ABI_VERSION; }
const char* hippo_last_error(void) { return hippo::g_err; }
hippo_status hippo_device_check(void) { return hippo::check_arch(); }
int32_t hippo_sm_count(void) { return hippo::sm_count(); }

}  // extern "C"
