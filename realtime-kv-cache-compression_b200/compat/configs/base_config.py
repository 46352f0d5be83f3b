"""`configs.base_config` of the reference, served by rtkv."""
from rtkv.config import CompressionConfig  # noqa: F401
