// Host-side plumbing of libgenpose_b200.so: thread-local error text and launch accounting.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace gp {

static thread_local char t_err[512] = "";
static thread_local long long t_launches = 0;

char *last_error_buf() { return t_err; }

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { t_launches += n; }

int num_sms() {
    static thread_local int cached_dev = -1, cached_sms = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        cached_dev = dev;
        cached_sms = sms;
    }
    return cached_sms;
}

}  // namespace gp

extern "C" int gp_version(void) { return 1; }
extern "C" const char *gp_last_error(void) { return gp::last_error_buf(); }
extern "C" long long gp_launch_count(void) { return gp::t_launches; }
extern "C" void gp_launch_count_reset(void) { gp::t_launches = 0; }

extern "C" int gp_zero(void *dst, size_t bytes, gp_stream_t s) {
    if (bytes == 0) return GP_OK;
    GP_REQUIRE(dst, "gp_zero: null pointer");
    GP_CUDA(cudaMemsetAsync(dst, 0, bytes, gp::as_stream(s)));
    return GP_OK;
}
