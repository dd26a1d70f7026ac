"""A bounded, fixed-seed slice of the randomised parity soak (tools/soak_parity.py) inside `pytest -m gpu`: random sizes,
contents and parameters through the C ABI, bit for bit against live cv2 (ORB, kNN, BGR2GRAY), the C++ restatement of the
reference (detector, matcher, undistortImage), the batched / pipelined sequence path against the single-image calls, and
findEssentialMat against the oracle loop AND cv2 (masks exact, E under the condition-aware bound stated in the tool)."""
import importlib.util
import json
import os

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed0", [2, 3])
def test_soak_slice_has_no_mismatch(gpu_ctx, seed0):
    pytest.importorskip("cv2")
    spec = importlib.util.spec_from_file_location("soak_parity", os.path.join(ROOT, "tools", "soak_parity.py"))
    soak = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(soak)
    rec = soak.run(budget=40.0, seed0=seed0, max_cases=120, context=gpu_ctx)
    assert sum(rec["cases"].values()) >= 30, rec
    assert rec["mismatches"] == 0, json.dumps(rec["details"])[:2000]
