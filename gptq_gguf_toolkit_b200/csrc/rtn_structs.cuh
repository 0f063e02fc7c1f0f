// rtn_structs.cuh -- parameter block and shared-memory layout of the RTN kernels (csrc/rtn.cu).  Included INSIDE rtn.cu's
// anonymous namespace after `constexpr int R = 32; constexpr int NT = 256;` (and by tests/helpers/simt_emu the same way): not a
// stand-alone header.
struct RtnParams {
    const void *W;       // (d_row, *) of w_dtype, row stride ld_in elements
    int w_dtype;
    long ld_in;
    int d_row, nsb;
    SearchParams sp;
    // metadata outputs: d/dmin at [row*d_stride + sb], sq/zq at [row*sq_stride + sb*GPR + g]
    uint16_t *d, *dmin;
    long d_stride;
    uint8_t *sq, *zq;
    long sq_stride;
    // optional full-matrix outputs (row stride = nsb*256 elements)
    uint8_t *qweight;
    uint8_t *packed;
    void *wdeq;
    int wdeq_dtype;
    uint32_t *flags;
};

struct __align__(16) RtnSmem {
    float Wt[R * 256];
    uint8_t codes[R * 256];
    float gsc[R * 16];
    float gzr[R * 16];
    RowScales<R> rs;
};
