"""The reference's OWN unit tests of the path, unmodified, against the drop-in: `baseline/_ref/tests/test_{compression,
importance_scoring,quantization}.py` (installed by baseline/install_ref.py from /root/reference; git-ignored, travels
with the snapshot) run with the import-compatible tree `compat/` first on PYTHONPATH, so that `configs.base_config` and
`src.compression.*` resolve to rtkv.  Expected outcome = the reference's own on the same torch: 17 passed and
ONE stale failure, `test_normalize_attention_scores` (it asserts a zero mean for what the code min-max normalises,
tests/test_importance_scoring.py:74-78 of the reference)."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = os.path.join(ROOT, "baseline", "_ref", "tests")
PKG = os.path.join(ROOT, "realtime-kv-cache-compression_b200")
FILES = ["test_compression.py", "test_importance_scoring.py", "test_quantization.py"]

pytestmark = pytest.mark.gpu


def test_reference_unit_tests_pass_against_the_drop_in():
    missing = [f for f in FILES if not os.path.exists(os.path.join(REF_TESTS, f))]
    if missing:
        pytest.skip("baseline/_ref/tests is not installed (python baseline/install_ref.py needs /root/reference)")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(PKG, "compat"), PKG])
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "-rf"] + [os.path.join(REF_TESTS, f) for f in FILES],
                       cwd=REF_TESTS, env=env, capture_output=True, text=True, timeout=900)
    out = r.stdout + r.stderr
    m = re.search(r"(?:(\d+) failed, )?(\d+) passed", out)
    assert m, out[-3000:]
    failed, passed = int(m.group(1) or 0), int(m.group(2))
    failures = re.findall(r"^FAILED (\S+)", out, flags=re.M)
    assert passed == 17 and failed == 1, out[-3000:]
    assert len(failures) == 1 and failures[0].endswith("TestPromptGuidedImportanceScorer::test_normalize_attention_scores"), failures
    # it really was the drop-in that ran: the compat tree re-exports rtkv classes
    probe = subprocess.run([sys.executable, "-c", "import src.compression.unified_compressor as m; print(m.RealTimePrefillCompressor.__module__)"],
                           env=env, capture_output=True, text=True, cwd=REF_TESTS)
    assert probe.stdout.strip().startswith("rtkv."), probe.stdout + probe.stderr
