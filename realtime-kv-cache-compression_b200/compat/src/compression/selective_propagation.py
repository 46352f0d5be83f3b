"""`src.compression.selective_propagation` of the reference, served by rtkv."""
from rtkv.selective_propagation import SelectiveTokenPropagator  # noqa: F401
