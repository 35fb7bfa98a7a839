"""The committed fixtures under tests/golden/ still agree with the oracle (CPU), so a fixture and
the code that made it cannot drift apart unnoticed."""
import os

import numpy as np
import pytest

from oracle import etkf

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("tag", ["c1", "c2", "c3s", "c5s"])
def test_analysis_fixture_matches_sparse_oracle(tag):
    z = np.load(os.path.join(GOLD, "analysis_cases.npz"))
    X, idx, y, mask, Xa = (z["%s_%s" % (tag, n)] for n in ("X", "idx", "y", "mask", "Xa"))
    rho, sd = z[tag + "_par"]
    got = etkf.analysis_sparse(X, y, idx, mask / sd ** 2, rho)
    assert np.max(np.abs(got - Xa)) / np.max(np.abs(Xa)) < 1e-11


def test_converged_fixture_is_in_the_converged_regime():
    z = np.load(os.path.join(GOLD, "etkf_converged_dp5.npz"))
    assert z["rmse_dp5"].shape == (200,)
    assert 0.15 < z["rmse_dp5"][50:].mean() < 0.25        # textbook value ~0.2, far below sigma_obs = 1
