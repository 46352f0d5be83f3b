"""rtkv — B200-native real-time prefill KV-cache compression.

Drop-in for the reference's compression package: same class names, method
signatures and return structures; the work is done by sm_100a CUDA kernels
behind the C ABI in include/rtkv.h (librtkv.so).  No CPU fallback.
"""
from .config import CompressionConfig
from .dynamic_quantization import DynamicPrecisionQuantizer
from .selective_propagation import SelectiveTokenPropagator
from .token_importance import LayerWiseImportanceTracker, PromptGuidedImportanceScorer
from .unified_compressor import CompressionHook, PackedKV, RealTimePrefillCompressor, UnifiedCompressor

__all__ = ["CompressionConfig", "PackedKV", "DynamicPrecisionQuantizer", "SelectiveTokenPropagator",
           "LayerWiseImportanceTracker", "PromptGuidedImportanceScorer", "CompressionHook",
           "RealTimePrefillCompressor", "UnifiedCompressor"]
