"""Pins the oracle's quantiser against the REAL reference on the edge rows the GPU tests use
(tests/test_parity_gpu.py: stress_rows, fuzz_rows): wide dynamic range, zeros of both signs, denormals, a
constant row that rounds to -0.0, non-finite rows, fp16 rows whose x/scale overflows the half range.
Container only (imports /root/reference read-only, like pin_against_reference.py).  Writes the tally to
tests/golden/PINNING_EDGE.json.      Usage: python oracle/pin_edge_rows.py
"""
import importlib.util
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
import oracle  # noqa: E402
from configs.base_config import CompressionConfig  # noqa: E402  (the reference's)
from src.compression.dynamic_quantization import DynamicPrecisionQuantizer  # noqa: E402

TDT = {oracle.F32: torch.float32, oracle.BF16: torch.bfloat16, oracle.F16: torch.float16}


def load_row_generators():
    sys.path.insert(0, os.path.join(ROOT, "tests"))      # the test module imports its conftest
    spec = importlib.util.spec_from_file_location("tp", os.path.join(ROOT, "tests", "test_parity_gpu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.stress_rows, mod.fuzz_rows


def same(a, b):
    """bit-equal floats, any NaN == any NaN"""
    return (a == b) & (np.signbit(a) == np.signbit(b)) | (np.isnan(a) & np.isnan(b))


def main():
    stress_rows, fuzz_rows = load_row_generators()
    rng = np.random.default_rng(77)
    rows = np.stack(stress_rows(rng, 256) + list(fuzz_rows(np.random.default_rng(1000), 400, 256)))
    tally = {"rows": int(rows.shape[0]), "cases": []}
    for dt in (oracle.F32, oracle.BF16, oracle.F16):
        for bits in ((2, 4, 8), (3, 6, 12)):
            x = torch.from_numpy(rows[None]).to(TDT[dt])
            S = x.shape[1]
            labels = np.random.default_rng(5).integers(0, 3, (1, S)).astype(np.uint8)
            cfg = CompressionConfig(low_precision_bits=bits[0], medium_precision_bits=bits[1], high_precision_bits=bits[2])
            q = DynamicPrecisionQuantizer(cfg)
            with torch.no_grad():
                ref_k, _, _ = q.apply_mixed_precision_quantization(x, x.clone(), torch.from_numpy(labels.astype(np.int64)))
            want = ref_k.float().numpy()[0]
            got_bits, _, _ = oracle.fake_quant_rows(oracle.from_torch(x)[0], dt, labels, bits)
            got = oracle.to_torch(got_bits, dt).float().numpy()[0]
            ok = same(got, want)
            bad_rows = np.unique(np.nonzero(~ok)[0])
            tally["cases"].append({"dtype": int(dt), "bits": list(bits), "elements": int(ok.size),
                                   "mismatches": int((~ok).sum()), "bad_rows": bad_rows[:10].tolist(),
                                   "nan_elements": int(np.isnan(want).sum()),
                                   "negative_zero_elements": int(((want == 0) & np.signbit(want)).sum())})
            print(tally["cases"][-1])
            if dt == oracle.F16 and bits == (2, 4, 8):      # the most edge-rich case travels as a golden fixture
                np.savez_compressed(os.path.join(ROOT, "tests", "golden", "edge_rows_f16.npz"),
                                    want=oracle.from_torch(ref_k)[0][0], labels=labels[0], bits=np.array(bits))
    tally["torch"] = torch.__version__
    with open(os.path.join(ROOT, "tests", "golden", "PINNING_EDGE.json"), "w") as f:
        json.dump(tally, f, indent=1)
    assert all(c["mismatches"] == 0 for c in tally["cases"]), "oracle differs from the reference on edge rows"


if __name__ == "__main__":
    main()
