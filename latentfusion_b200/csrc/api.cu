// C-ABI glue: version / error reporting / precision dispatch for lf_conv_fwd.
#include "common.cuh"

#include <stdlib.h>
#include <string.h>

namespace lf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
        cached_dev = dev;
    }
    return cached;
}

static const char* kOptName[OPT_COUNT] = {"LFB200_TC_DC", "LFB200_TC_DEBUG", "LFB200_TC_NO_DUAL", "LFB200_RESAMPLE_KC",
                                           "LFB200_RESAMPLE_W", "LFB200_RESAMPLE_BRICK", "LFB200_BWDCAM", "LFB200_TC_NO_TRI"};
static int g_opt[OPT_COUNT];
static bool g_opt_loaded = false;

static void load_options() {
    if (g_opt_loaded) return;
    for (int i = 0; i < OPT_COUNT; ++i) {
        const char* e = getenv(kOptName[i]);
        g_opt[i] = (e && e[0]) ? atoi(e) : 0;
    }
    g_opt_loaded = true;
}

int option(Option o) { load_options(); return g_opt[o]; }

int conv_fp32_launch(const lf_conv_desc* d, const float* x, const float* w, const float* bias, float* y,
                     float* rnorm, cudaStream_t st);
int conv_tc_supported(const lf_conv_desc* d);
int conv_tc_launch(const lf_conv_desc* d, const float* x, const float* w, const float* bias, float* y,
                   float* rnorm, cudaStream_t st);

}  // namespace lf

extern "C" const char* lf_version(void) { return "lfb200 0.1.0 (sm_100a)"; }
extern "C" const char* lf_last_error(void) { return lf::g_err; }
extern "C" int lf_sm_count(void) { return lf::sm_count(); }
// debug/tuning switches (the LFB200_* environment variables, read once at first use): explicit override; not thread-safe
extern "C" int lf_set_option(const char* name, int value) {
    lf::load_options();
    for (int i = 0; i < lf::OPT_COUNT; ++i)
        if (name != nullptr && strcmp(name, lf::kOptName[i]) == 0) { lf::g_opt[i] = value; return LF_OK; }
    lf::set_error("lf_set_option: unknown option %s", name ? name : "(null)");
    return LF_EINVAL;
}

extern "C" int lf_conv_fwd(const lf_conv_desc* desc, const float* x, const float* w, const float* bias,
                           float* y, float* rnorm, void* stream) {
    if (desc == nullptr) { lf::set_error("conv: null descriptor"); return LF_EINVAL; }
    if (desc->precision != 0) {
        if (!lf::conv_tc_supported(desc)) {
            lf::set_error("conv: tcgen05 path does not support this shape (ndim=%d k=%d cin=%d cout=%d)",
                          desc->ndim, desc->k, desc->cin, desc->cout);
            return LF_EUNSUPPORTED;
        }
        return lf::conv_tc_launch(desc, x, w, bias, y, rnorm, (cudaStream_t)stream);
    }
    return lf::conv_fp32_launch(desc, x, w, bias, y, rnorm, (cudaStream_t)stream);
}
