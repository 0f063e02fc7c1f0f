// api.cu -- C-ABI plumbing of libgq.so: version, error text, format registry, search parameter table.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

#include <atomic>

namespace {
thread_local char g_err[512] = "";
std::atomic<long> g_launches{0};
}

void gq_count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" long gq_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void gq_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// num[i] = fp32(rmin + rdelta*i + maxq), evaluated in double exactly like the Python expression
// `self.rmin + self.rdelta * i + self.maxq` of the reference (quant_utils.py:241).
void gq_fill_search_params(SearchParams &sp, int maxq, double rmin, double rdelta, int nstep) {
    sp.nstep = nstep;
    for (int i = 0; i < 64; ++i) sp.num[i] = (float)(rmin + rdelta * (double)i + (double)maxq);
}

extern "C" int gq_abi_version(void) { return GQ_ABI_VERSION; }

extern "C" const char *gq_last_error(void) { return g_err; }

extern "C" int gq_format_info(int qtype, int out7[7]) {
    FmtInfo f;
    if (!out7 || !gq_fmt_info(qtype, f)) {
        gq_set_error("gq_format_info: unknown q_type %d", qtype);
        return GQ_ERR_INVALID;
    }
    out7[0] = f.bits; out7[1] = f.qmin; out7[2] = f.qmax; out7[3] = f.smq; out7[4] = f.gs; out7[5] = f.asym; out7[6] = f.ts;
    return GQ_OK;
}

extern "C" int gq_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
