"""`src.compression.unified_compressor` of the reference, served by rtkv."""
from rtkv.unified_compressor import CompressionHook, RealTimePrefillCompressor, UnifiedCompressor  # noqa: F401
