"""flexq_b200: B200-native (sm_100a) W6A6/W6A8 quantized-linear hot path of FlexQ.

Public surface mirrors the reference's: ``QuantLinear`` / ``UniformAffineQuantizer``
(algorithm/flexq_quantize) on the python side and the C ABI in include/flexq_b200.h.
"""
from . import capi  # noqa: F401
from .quantizer import UniformAffineQuantizer  # noqa: F401
from .int_linear import QuantLinear  # noqa: F401
from .int_llama_layer import QuantLlamaAttention, QuantLlamaDecoderLayer, QuantLlamaMLP, quantize_llama  # noqa: F401

__all__ = ["capi", "UniformAffineQuantizer", "QuantLinear", "QuantLlamaMLP", "QuantLlamaAttention", "QuantLlamaDecoderLayer",
           "quantize_llama"]
