// capi.cu -- error plumbing and device check for libcmflow_b200.
#include <stdarg.h>
#include <string.h>

#include "cmf_common.cuh"

static thread_local char g_err[512] = "";

void cmf_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *cmf_last_error(void) { return g_err; }
extern "C" const char *cmf_version(void) { return "cmflow_b200 0.1.0 sm_100a"; }

extern "C" int cmf_device_check(void) {
    int dev = 0;
    CMF_CUDA(cudaGetDevice(&dev));
    int major = 0;
    CMF_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) {
        cmf_set_error("cmf_device_check: device %d has compute capability %d.x; this library is built for sm_100a only", dev, major);
        return CMF_ERR_STATE;
    }
    return CMF_OK;
}
