"""Randomised differential test of the C ABI against the oracle (scripts/fuzz_gpu.py): random small graphs, widths, walkers,
summation-order flags, fused vectors, padding workspaces, strides, sum and max -- every case bit for bit where the library
says it sums in CSR order, within 1e-4 elsewhere.  3000 cases ran clean on a B200 in round 2; the suite keeps 300."""
import importlib.util
import os

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [1, 2])
def test_random_calls_match_the_oracle(pkg, oracle, seed):
    spec = importlib.util.spec_from_file_location("fuzz_gpu", os.path.join(ROOT, "scripts", "fuzz_gpu.py"))
    fuzz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fuzz)
    lines = []
    bad = fuzz.run(150, seed, log=lines.append)
    assert bad == 0, "\n".join(lines)
