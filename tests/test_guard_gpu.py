"""Own bounds check (compute-sanitizer is closed on the GPU pool): the unit workload that touches every kernel runs with
guard-banded device buffers, and no kernel may have written outside the buffer it was handed (profiles/guard_check.py)."""
import importlib.util
import os

import pytest

pytestmark = pytest.mark.gpu


def test_no_kernel_writes_outside_its_buffers():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("guard_check", os.path.join(root, "profiles", "guard_check.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.main() == 0
