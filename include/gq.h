/*
 * gq.h -- C ABI of libgq.so: the B200 (sm_100a) GPTQ -> GGUF K-quant hot path.
 *
 * This is the drop-in boundary for the per-layer hot path of IST-DASLab/gptq-gguf-toolkit
 * (reference paths below are relative to quant/gptq/src/ of that repository).  The reference
 * has no FFI: its "operator API" is the Python protocol of class GPTQ (gptq.py:28-324), the
 * stateless numerics in quant_utils.py and the byte packers in packing_utils.py.  Each entry
 * point here replaces one of those, with plain pointers and sizes only (no torch types).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch tensors: pass data_ptr());
 *     the library keeps no reference past the call and never allocates tensor-sized memory;
 *   - all work is enqueued on the caller's stream (`stream` = cudaStream_t as void*), there is no
 *     hidden synchronisation and no host read-back: data-dependent failures (non positive definite
 *     Hessian) are resolved ON DEVICE exactly like the reference resolves them on the host, and are
 *     reported through a device-side int the caller may read whenever it likes;
 *   - return value: GQ_OK or a gq_status error; gq_last_error() gives a thread-local message;
 *     nothing throws or exits across the ABI;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns GQ_ERR_CUDA.
 *   - matrices are row-major.  W is (d_row, d_col) like nn.Linear.weight; H and U are (d_col, d_col);
 *     U is the UPPER Cholesky factor of inv(H) stored row-major (the reference's torch tensor is
 *     column-major, gptq.py:320 -- convert knowingly when feeding reference fixtures).
 *   - K-quant metadata layout (same as the reference's five tensors, gptq.py:295):
 *       qweight (d_row, d_col)            u8 codes (Q2/4/5_K) or i8 codes (Q3/6_K)
 *       d, dmin (d_row, d_col/256)        fp16 bit patterns ("super_group_scale/zero")
 *       sq, zq  (d_row, d_col/group_size) u8 / i8 ("group_scale_quant/group_zero_quant")
 *       packed  (d_row, d_col/256*type_size) raw GGUF block bytes (packing_utils.py)
 */
#ifndef GQ_H_
#define GQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GQ_ABI_VERSION 1
#define GQ_API __attribute__((visibility("default")))

typedef enum {
    GQ_OK = 0,
    GQ_ERR_INVALID = 1,      /* bad argument (shape, alignment, null pointer, unknown q_type) */
    GQ_ERR_CUDA = 2,         /* CUDA runtime error or no device */
    GQ_ERR_UNSUPPORTED = 3,  /* valid in the reference, not implemented here (see message) */
    GQ_ERR_WORKSPACE = 4     /* workspace too small; call the matching *_workspace_bytes */
} gq_status;

/* GGML type ids, identical to the reference's GGMLQuantizationType (quant_utils.py:11-16) */
typedef enum { GQ_Q2_K = 10, GQ_Q3_K = 11, GQ_Q4_K = 12, GQ_Q5_K = 13, GQ_Q6_K = 14 } gq_qtype;

typedef enum { GQ_F32 = 0, GQ_F16 = 1, GQ_BF16 = 2 } gq_dtype;

/* GQ_MODE_EXACT reproduces the reference's CPU arithmetic bit for bit given the same (W, U).
 * GQ_MODE_FAST runs the rank-k update on tcgen05 tensor cores (split-fp16 operands, fp32 accumulation); not bit-identical.
 * The exact arithmetic has two bit-identical schedules: ONE left-looking launch per layer (every 32-row CTA applies all
 * earlier blocks to its own tile), and a right-looking one (per 256-column super-block a panel launch + an exact FFMA
 * update of the whole trailing part by all SMs) that wins when a launch has few rows and many columns (a row slice of
 * down_proj on one of several GPUs).  GQ_MODE_EXACT picks by a cost model (environment GQ_EXACT_SCHEDULE=left|right
 * overrides it); GQ_MODE_EXACT_LEFT / GQ_MODE_EXACT_RIGHT force one (tests compare both with the goldens). */
typedef enum { GQ_MODE_EXACT = 0, GQ_MODE_FAST = 1, GQ_MODE_EXACT_LEFT = 2, GQ_MODE_EXACT_RIGHT = 3 } gq_mode;

typedef void *gq_stream_t;

GQ_API int gq_abi_version(void);
GQ_API const char *gq_last_error(void);

/* Format registry -- replaces GGML_QUANT_SIZES (quant_utils.py:19-26).
 * out7 = {bits, qmin, qmax, scale_maxq, group_size, asymmetric(0/1), gguf type_size in bytes}. */
GQ_API int gq_format_info(int qtype, int out7[7]);

/* Number of CUDA devices visible (0 when there is none; never an error). */
GQ_API int gq_device_count(void);

/* Number of CUDA kernels this library has launched in this process so far (bench.py's gpu_launches). */
GQ_API long gq_launch_count(void);

/* H <- beta*H + alpha * X^T X   -- replaces GPTQ.update's addmm_ (gptq.py:110-112).
 * X: (n_tok, d_col) of x_dtype, row-major, contiguous.  H: (d_col, d_col) fp32, kept fully symmetric.
 * 16-bit inputs (bf16 and fp16, d_col a multiple of 256) take the tcgen05 path (products exact in fp32, fp32 accumulation in
 * TMEM); fp32 inputs the fp32 SIMT path.  workspace: gq_hessian_workspace_bytes() bytes (may be 0). */
GQ_API size_t gq_hessian_workspace_bytes(long n_tok, int d_col, int x_dtype);
GQ_API int gq_hessian_update(float *H, const void *X, long n_tok, int d_col, int x_dtype, float beta,
                      float alpha, void *workspace, size_t ws_bytes, gq_stream_t stream);

/* Dead-channel fix -- replaces quantization_pre_step (gptq.py:134-141):
 * for every i with H[i,i] == 0: H[i,i] = 1 and W[:,i] = 0.  W: fp32 working copy. */
GQ_API int gq_pre_step(float *H, float *W, int d_row, int d_col, gq_stream_t stream);

/* U <- upper Cholesky factor of inv(H + damp*I) -- replaces GPTQ._prepare + inv_sym
 * (gptq.py:305-324, linalg_utils.py:8-12).  Steps, all on device: columns of W that are all zero get
 * their H row/col zeroed and diag 1; damp = rel_damp * mean(diag H) added to the diagonal (H is modified
 * in place, as in the reference); U = R^-1 where H = R R^T, R upper (equals chol(inv(H), upper)).
 * If H is not positive definite, U is set to the identity (gptq.py:321-323) and *not_pd_flag (device
 * int, may be NULL) is set to 1, else 0. */
GQ_API size_t gq_prepare_workspace_bytes(int d_col);
GQ_API int gq_prepare(float *H, const float *W, int d_row, int d_col, float rel_damp, float *U_out,
               void *workspace, size_t ws_bytes, int *not_pd_flag, gq_stream_t stream);

/* The column-blocked quantise -> error -> rank-k-update loop -- replaces GPTQ.step (gptq.py:146-295)
 * for act_order = static_groups = False (see gq_gptq_quantize_ex for those); one call per layer (the number of kernel
 * launches behind it depends on the schedule, see gq_mode).
 *   W     (d_row, d_col) fp32 working copy; CLOBBERED (holds the propagated errors on return).
 *   U     from gq_prepare.
 *   rmin, rdelta, nstep: K-quant search parameters (quant_utils.py:66-68), doubles like Python floats.
 *   outputs qweight,d,sq,dmin,zq as described at the top; for Q3_K/Q6_K dmin and zq are written as 0.
 *   packed (optional, may be NULL): GGUF block bytes, fused bit-pack (packing_utils.py:33-326).
 *   wdeq   (optional): dequantised weights (d_row, d_col) of wdeq_dtype == dequantize_linear_weight(...)
 *          .to(dtype) (quant_utils.py:277-310, quantizer.py:257-264).
 *   search_flags (optional): 2 u32 per super-block; bit i of [2s] = search candidate i saw a group with
 *          D > eps, bit i of [2s+1] = candidate i was accepted by a group.  The reference skips a
 *          candidate for the whole call when no group is valid (quant_utils.py:250-252); this library
 *          never skips, so (flags[2s+1] & ~flags[2s]) != 0 marks the (degenerate) inputs on which the
 *          two can differ.  Must be zero-initialised by the caller.
 * block_size: 128 (the reference's run_quant.sh default: the fused kernels) or 32 / 64 / 256 (a plain right-looking schedule with
 * the same arithmetic, exact modes only, no act_order); other values return GQ_ERR_UNSUPPORTED.
 * mode GQ_MODE_FAST needs gq_gptq_workspace_bytes() of scratch (the exact modes need none): the rank-k updates between
 * 256-column super-blocks then run as tcgen05 split-fp16 GEMMs (fp32-class accuracy, NOT bit-identical to the reference).
 * GQ_MODE_EXACT / _LEFT / _RIGHT give bit-identical outputs (see gq_mode); they differ in the number of launches. */
GQ_API size_t gq_gptq_workspace_bytes(int d_row, int d_col, int mode);
GQ_API int gq_gptq_quantize(float *W, const float *U, int d_row, int d_col, int qtype, int block_size,
                     double rmin, double rdelta, int nstep, int mode,
                     void *qweight, uint16_t *d, void *sq, uint16_t *dmin, void *zq,
                     uint8_t *packed, void *wdeq, int wdeq_dtype, uint32_t *search_flags,
                     void *workspace, size_t ws_bytes, gq_stream_t stream);

/* gq_gptq_quantize with the reference's two optional variants of GPTQ.step (gptq.py:184-216, 233-238, 273-277;
 * exact modes only; Q3_K ignores both, gptq.py:204-206):
 *   static_groups = 0  scales searched per super-block on the live weights (gq_gptq_quantize);
 *                 = 1  all scales / zeros searched up front on W as passed (gptq.py:184-196);
 *                 = 2  d, dmin, sq, zq already hold them on entry (the caller ran gq_get_scale_and_zero on the 256-column
 *                      slabs of the UN-permuted weights -- needed for act_order, where W below is permuted).
 *   perm (device int32[d_col], may be NULL) = act_order: W and U are in PERMUTED column order (W[:, perm], U from
 *        H[perm][:, perm], perm = argsort(diag H, descending)); loop column c is original column perm[c] and uses that
 *        column's group scales.  Requires static_groups = 2.  qweight comes out in LOOP order (the caller un-permutes it:
 *        out[:, perm[c]] = qweight[:, c]); packed and wdeq must be NULL (use gq_pack / gq_dequantize afterwards). */
GQ_API int gq_gptq_quantize_ex(float *W, const float *U, int d_row, int d_col, int qtype, int block_size,
                        double rmin, double rdelta, int nstep, int mode, int static_groups, const int *perm,
                        void *qweight, uint16_t *d, void *sq, uint16_t *dmin, void *zq,
                        uint8_t *packed, void *wdeq, int wdeq_dtype, uint32_t *search_flags,
                        void *workspace, size_t ws_bytes, gq_stream_t stream);

/* Kernel-level timing of gq_gptq_quantize for benchmarks: when enabled, CUDA events are recorded around every launch of
 * the fused search / column-loop kernel (kind 0) and of the tcgen05 rank-k GEMM of GQ_MODE_FAST (kind 1).
 * gq_profile_read synchronises on the recorded events, returns summed milliseconds and launch counts per kind, and clears. */
GQ_API void gq_profile_enable(int on);
GQ_API int gq_profile_read(float ms[2], int counts[2]);
/* Same with kind 2 = the operand preparation (fp16 hi/lo split) launches of GQ_MODE_FAST's GEMMs as a third entry. */
GQ_API int gq_profile_read3(float ms[3], int counts[3]);

/* RTN K-quant without a Hessian -- replaces Quantizer._quant_non_block_module
 * (quantizer.py:278-330).  W: (d_row, d_col) of w_dtype, read-only; arithmetic is fp32. */
GQ_API int gq_rtn_quantize(const void *W, int w_dtype, int d_row, int d_col, int qtype,
                    double rmin, double rdelta, int nstep,
                    void *qweight, uint16_t *d, void *sq, uint16_t *dmin, void *zq,
                    uint8_t *packed, void *wdeq, int wdeq_dtype, gq_stream_t stream);

/* gq_rtn_quantize with the scale search in the arithmetic of the weight's own dtype, as the reference does it -- the driver's
 * default for embed_tokens / lm_head of a 16-bit model (validated on B200 against the reference's own outputs,
 * tests/golden/rtn_bf16.npz / rtn_f16.npz, tests/test_zz_gpu_rtn_native.py):
 * (quantizer.py:303-305 passes the weight un-widened): for GQ_BF16 / GQ_F16 every op of the search rounds to that dtype; the final
 * quantize() is fp32 as in the reference.  GQ_F32 falls through to gq_rtn_quantize.  Same arguments and outputs. */
GQ_API int gq_rtn_quantize_native(const void *W, int w_dtype, int d_row, int d_col, int qtype,
                    double rmin, double rdelta, int nstep,
                    void *qweight, uint16_t *d, void *sq, uint16_t *dmin, void *zq,
                    uint8_t *packed, void *wdeq, int wdeq_dtype, gq_stream_t stream);

/* Scale / min search of one super-block column -- replaces quant_utils.Quantizer.get_scale_and_zero
 * (quant_utils.py:90-145).  x: (rows, 256) fp32 with row stride x_stride (elements).
 * d,dmin: one fp16 per row (stride d_stride elements); sq,zq: 256/group_size codes per row
 * (stride sq_stride). */
GQ_API int gq_get_scale_and_zero(const float *x, long x_stride, int rows, int qtype,
                          double rmin, double rdelta, int nstep,
                          uint16_t *d, uint16_t *dmin, long d_stride, void *sq, void *zq, long sq_stride,
                          uint32_t *search_flags, gq_stream_t stream);

/* dequantize_linear_weight (quant_utils.py:277-310): out (d_row, d_col) of out_dtype. */
GQ_API int gq_dequantize(int qtype, const void *qweight, const uint16_t *d, const void *sq,
                  const uint16_t *dmin, const void *zq, int d_row, int d_col,
                  void *out, int out_dtype, gq_stream_t stream);

/* pack_Q2K ... pack_Q6K (packing_utils.py:33-326); inputs are not modified. */
GQ_API int gq_pack(int qtype, const void *qweight, const uint16_t *d, const void *sq,
            const uint16_t *dmin, const void *zq, int d_row, int d_col,
            uint8_t *out, gq_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GQ_H_ */
