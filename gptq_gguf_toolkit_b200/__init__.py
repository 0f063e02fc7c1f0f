"""gptq_gguf_toolkit_b200 -- B200-native GPTQ -> GGUF K-quant hot path (libgq, include/gq.h) behind the
Python surface of IST-DASLab/gptq-gguf-toolkit's quant/gptq stage."""
from .quant_utils import GGML_QUANT_SIZES, GGMLQuantizationType, QuantizationScale, dequantize_linear_weight  # noqa: F401

__all__ = ["GGML_QUANT_SIZES", "GGMLQuantizationType", "QuantizationScale", "dequantize_linear_weight"]
