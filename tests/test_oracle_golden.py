"""The CPU oracle (oracle/restatement.py) must reproduce every golden scenario recorded from the reference's own
source (tests/golden/*.pt, written by oracle/make_golden.py). This is what pins the oracle."""
import glob
import os

import pytest
import torch

from oracle.make_golden import check_oracle, scenarios

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.pt")))


def test_golden_files_cover_all_scenarios():
    names = {os.path.splitext(os.path.basename(p))[0] for p in GOLDEN}
    assert names == set(scenarios().keys())


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_oracle_reproduces_reference(path):
    torch.set_num_threads(1)
    g = torch.load(path, weights_only=False)
    check_oracle(g)
