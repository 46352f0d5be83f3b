import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "realtime-kv-cache-compression_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files if k != "meta"}
    d["meta"] = json.loads(bytes(z["meta"]).decode())
    return d


LAYER_CASES = ["layer_f32_s64", "layer_f32_b2_s96", "layer_f32_s256_l31", "layer_f32_s700_888",
               "layer_bf16_s128", "layer_f16_s128", "layer_f32_tiny_fallback", "layer_f16_16_8_4_nohigh"]
STAGE_CASES = ["stage_f32_8_4_2", "stage_f32_16_8_4", "stage_bf16_8_4_2", "stage_f16_8_4_2"]


@pytest.fixture(scope="session")
def built_lib():
    """librtkv.so, built on demand (nvcc cross-compiles without a GPU)."""
    sys.path.insert(0, PKG)
    import build as rtkv_build
    return rtkv_build.build()
