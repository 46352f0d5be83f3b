"""`src.compression.dynamic_quantization` of the reference, served by rtkv."""
from rtkv.dynamic_quantization import DynamicPrecisionQuantizer  # noqa: F401
