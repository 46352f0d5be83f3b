#!/usr/bin/env python
"""Install the UNMODIFIED reference hot path into baseline/_ref/ (git-ignored; it still travels to the GPU box
with the gpurun snapshot, like the built .so files).

The reference has no build metadata (`python -m pip install --no-index --target baseline/_ref /root/reference` fails
with "Neither 'setup.py' nor 'pyproject.toml' found") and is run with PYTHONPATH=<repo root> (README.md:76), so the
install is a verbatim copy of the five files of the path, keeping their layout:
    configs/base_config.py
    src/compression/{token_importance,dynamic_quantization,selective_propagation,unified_compressor}.py
plus the reference's three unit-test files of the path (tests/test_{compression,importance_scoring,quantization}.py), which
tests/test_reference_suite_gpu.py runs unmodified against the drop-in (compat/) on the GPU box.
Nothing under baseline/_ref/ is product code: bench.py imports it only to TIME the reference's own CPU path
(`cpu_baseline.reference_python`), tests/ never need it (the GPU box has no /root/reference; parity fixtures made
from the reference live in tests/golden/).  A sha256 manifest proves the copies are unmodified."""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ["configs/base_config.py", "src/compression/token_importance.py", "src/compression/dynamic_quantization.py",
         "src/compression/selective_propagation.py", "src/compression/unified_compressor.py"]
# the reference's own unit tests of the path: run UNMODIFIED against the import-compatible tree (compat/) by
# tests/test_reference_suite_gpu.py.  Optional: `available()` only asks for FILES.
TEST_FILES = ["tests/test_compression.py", "tests/test_importance_scoring.py", "tests/test_quantization.py"]


def install(src_root="/root/reference", quiet=False):
    if not os.path.isdir(src_root):
        return False
    manifest = {}
    for rel in FILES + TEST_FILES:
        src, dst = os.path.join(src_root, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": src_root, "sha256": manifest}, f, indent=1)
    if not quiet:
        print(f"installed {len(FILES) + len(TEST_FILES)} reference files into {DEST}")
    return True


def available():
    return all(os.path.exists(os.path.join(DEST, rel)) for rel in FILES)


if __name__ == "__main__":
    sys.exit(0 if install(*(sys.argv[1:2])) else 1)
