/*
 * gq_fwd.h -- fused element-wise pieces of the calibration block forwards (libgq.so, csrc/fwd_ops.cu).
 *
 * NOT part of the reference's hot path: the block forwards belong to the caller's model.  These three
 * entry points replace groups of eager PyTorch kernels inside HF's Llama block with one pass over HBM each, keeping
 * HF's rounding points (see the header of fwd_ops.cu); gptq_gguf_toolkit_b200/fused_forward.py installs them after
 * checking each against the module it replaces.  16-bit activations (GQ_BF16 / GQ_F16) only, device pointers,
 * caller's stream, status codes as in gq.h.
 */
#ifndef GQ_FWD_H_
#define GQ_FWD_H_
#include "gq.h"
#ifdef __cplusplus
extern "C" {
#endif
/* out[r, :] = w * rn16(x[r, :] * rsqrt(mean(x[r, :]^2) + eps))   (transformers LlamaRMSNorm.forward) */
GQ_API int gq_fwd_rmsnorm(const void *x, const void *w, void *out, long rows, int dim, float eps, int dtype, gq_stream_t stream);
/* out = silu(gate) * up   (transformers LlamaMLP.forward: act_fn(gate_proj(x)) * up_proj(x)) */
GQ_API int gq_fwd_silu_mul(const void *gate, const void *up, void *out, long n, int dtype, gq_stream_t stream);
/* out = x * cos + rotate_half(x) * sin for x of logical shape (B, H, L, hd) with element strides (sb, sh, sl, 1)
 * (transformers apply_rotary_pos_emb); cos / sin: (B or 1, L, hd) contiguous. */
GQ_API int gq_fwd_rope(const void *x, void *out, const void *cos, const void *sin, int B, int H, int L, int hd, long sb, long sh,
                long sl, long osb, long osh, long osl, int cos_batched, int dtype, gq_stream_t stream);
#ifdef __cplusplus
}
#endif
#endif
