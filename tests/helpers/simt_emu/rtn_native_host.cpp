// TEST-ONLY: gptq_gguf_toolkit_b200/csrc/rtn_native.cuh (the experimental native-arithmetic RTN kernel, not yet run on a GPU) on
// the SIMT emulator, with the CUDA intrinsics from tests/helpers/host_shim.  RtnParams / RtnSmem are rtn.cu's own (rtn_structs.cuh).
#define SIMT_EMU 1
#define GQ_HOST_SHIM 1
#include "simt_emu.h"
#include "host_shim_intrinsics.h"
#include "rtn_native.cuh"

namespace {
constexpr int R = 32, NT = 256;
#include "rtn_structs.cuh"          // rtn.cu's own RtnParams / RtnSmem
RtnSmem g_sm;       // one block at a time: the block's shared memory

template <int QT, int RND> void run(const RtnParams &p) {
    simt::launch(dim3((p.d_row + R - 1) / R, p.nsb), dim3(NT), [&]() { rtn_native_body<QT, RND, R, NT, RtnParams, RtnSmem>(p, g_sm); });
}
template <int RND> int dispatch(int qtype, const RtnParams &p) {
    switch (qtype) {
    case GQ_Q2_K: run<GQ_Q2_K, RND>(p); return 0;
    case GQ_Q3_K: run<GQ_Q3_K, RND>(p); return 0;
    case GQ_Q4_K: run<GQ_Q4_K, RND>(p); return 0;
    case GQ_Q5_K: run<GQ_Q5_K, RND>(p); return 0;
    case GQ_Q6_K: run<GQ_Q6_K, RND>(p); return 0;
    }
    return -1;
}
template <int QT> void run32(const RtnParams &p) {
    simt::launch(dim3((p.d_row + R - 1) / R, p.nsb), dim3(NT), [&]() { rtn_body<QT, R, NT, RtnParams, RtnSmem>(p, g_sm); });
}
}  // namespace

// the SHIPPED kernel body (rtn_kernel = rtn_body<...>): fp32 weight, fp32 dequantised output
extern "C" int run_rtn_fp32(int qtype, const float *W, int d_row, int d_col, double rmin, double rdelta, int nstep, uint8_t *qweight,
                            uint16_t *d, uint16_t *dmin, uint8_t *sq, uint8_t *zq, uint8_t *packed, float *wdeq) {
    const int bits = qtype == GQ_Q2_K ? 2 : qtype == GQ_Q3_K ? 3 : qtype == GQ_Q4_K ? 4 : qtype == GQ_Q5_K ? 5 : 6;
    const int gs = (qtype == GQ_Q4_K || qtype == GQ_Q5_K) ? 32 : 16;
    RtnParams p;
    p.W = W; p.w_dtype = GQ_F32; p.ld_in = d_col; p.d_row = d_row; p.nsb = d_col / 256;
    p.sp.nstep = nstep;
    for (int i = 0; i <= nstep && i < 64; ++i) p.sp.num[i] = (float)(rmin + rdelta * (double)i + (double)((1 << bits) - 1));
    p.d = d; p.dmin = dmin; p.d_stride = p.nsb; p.sq = sq; p.zq = zq; p.sq_stride = d_col / gs;
    p.qweight = qweight; p.packed = packed; p.wdeq = wdeq; p.wdeq_dtype = GQ_F32; p.flags = nullptr;
    switch (qtype) {
    case GQ_Q2_K: run32<GQ_Q2_K>(p); return 0;
    case GQ_Q3_K: run32<GQ_Q3_K>(p); return 0;
    case GQ_Q4_K: run32<GQ_Q4_K>(p); return 0;
    case GQ_Q5_K: run32<GQ_Q5_K>(p); return 0;
    case GQ_Q6_K: run32<GQ_Q6_K>(p); return 0;
    }
    return -1;
}

// W: 16-bit weight (w_dtype 2 = bf16, 1 = fp16) as raw bits; wdeq: 16-bit output of the same dtype
extern "C" int run_rtn_native(int w_dtype, int qtype, const uint16_t *W, int d_row, int d_col, double rmin, double rdelta, int nstep,
                              uint8_t *qweight, uint16_t *d, uint16_t *dmin, uint8_t *sq, uint8_t *zq, uint8_t *packed, uint16_t *wdeq) {
    const int bits = qtype == GQ_Q2_K ? 2 : qtype == GQ_Q3_K ? 3 : qtype == GQ_Q4_K ? 4 : qtype == GQ_Q5_K ? 5 : 6;
    const int gs = (qtype == GQ_Q4_K || qtype == GQ_Q5_K) ? 32 : 16;
    RtnParams p;
    p.W = W; p.w_dtype = w_dtype; p.ld_in = d_col; p.d_row = d_row; p.nsb = d_col / 256;
    p.sp.nstep = nstep;
    for (int i = 0; i <= nstep && i < 64; ++i) p.sp.num[i] = (float)(rmin + rdelta * (double)i + (double)((1 << bits) - 1));
    p.d = d; p.dmin = dmin; p.d_stride = p.nsb; p.sq = sq; p.zq = zq; p.sq_stride = d_col / gs;
    p.qweight = qweight; p.packed = packed; p.wdeq = wdeq; p.wdeq_dtype = w_dtype; p.flags = nullptr;
    return w_dtype == GQ_BF16 ? dispatch<GQ_RND_BF16>(qtype, p) : dispatch<GQ_RND_F16>(qtype, p);
}
