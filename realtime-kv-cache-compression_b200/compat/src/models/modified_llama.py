"""`src.models.modified_llama` of the reference, served by rtkv."""
from rtkv.llama_hooks import (CompressedLlamaAttention, CompressedLlamaDecoderLayer,  # noqa: F401
                              CompressedLlamaForCausalLM, create_compressed_llama_model)
