// capi.cu -- ABI housekeeping for include/mrx.h: version, last-error string, device props.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace mrx {

static thread_local char g_last_error[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

}  // namespace mrx

extern "C" int mrx_abi_version(void) { return MRX_ABI_VERSION; }

extern "C" const char *mrx_last_error(void) { return mrx::g_last_error; }

extern "C" int mrx_device_props(int device, int *sm_count, int *cc_major, int *cc_minor,
                                int *max_smem_optin) {
  MRX_CHECK_ARG(sm_count && cc_major && cc_minor && max_smem_optin,
                "mrx_device_props: null pointer");
  MRX_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, device));
  MRX_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, device));
  MRX_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, device));
  MRX_CUDA(cudaDeviceGetAttribute(max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin,
                                  device));
  return MRX_OK;
}
