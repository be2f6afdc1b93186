"""Pins the oracle's DGPADM restatement (src/expokit/dgpadm.f:171-339) against
scipy.linalg.expm on random banded upper-Hessenberg H for m+2 in 12..102 (SURVEY 8c
pin 3)."""
import numpy as np
import pytest
from scipy.linalg import expm

import oracle


def krylov_like_H(n, rng, scale):
    H = np.zeros((n, n))
    for j in range(n - 2):
        H[j, j] = -abs(rng.standard_normal()) * scale
        H[j + 1, j] = abs(rng.standard_normal()) * scale
        if j > 0:
            H[j - 1, j] = rng.standard_normal() * scale
    H[n - 1, n - 2] = 1.0
    return H


@pytest.mark.parametrize("n", [12, 22, 32, 52, 77, 102])
@pytest.mark.parametrize("t", [1e-3, 0.3, 4.0])
def test_dgpadm_matches_scipy(n, t):
    rng = np.random.default_rng(n)
    H = krylov_like_H(n, rng, 5.0)
    E, ns, hnorm = oracle.dgpadm(H, t)
    R = expm(t * H)
    assert np.abs(E - R).max() <= 1e-11 * max(1.0, np.abs(R).max())
    assert abs(hnorm - abs(t * np.abs(H).sum(axis=1).max())) <= 1e-14 * hnorm
    assert ns == max(0, int(np.log(hnorm) / np.log(2.0)) + 2)


def test_dgpadm_leading_block():
    # the FSP-shrink path calls DGPADM with m = M+1 and ldh = M+2 (KrylovSolver.f90:438,490)
    rng = np.random.default_rng(3)
    H = krylov_like_H(20, rng, 2.0)
    E, _, _ = oracle.dgpadm(H, 0.5, m=19)
    assert np.abs(E - expm(0.5 * H[:19, :19])).max() < 1e-12


def test_dgpadm_null_matrix_is_an_error():
    with pytest.raises(RuntimeError):
        oracle.dgpadm(np.zeros((5, 5)), 1.0)
