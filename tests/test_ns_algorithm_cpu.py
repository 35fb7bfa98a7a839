"""The algorithm of the Newton-Schulz K4 kernels (csrc/etkf_ns.cu, csrc/etkf_ns_device.cuh), restated
in NumPy and checked against the oracle's eigendecomposition form on CPU: deflation of the known
eigenvector, Gershgorin / norm bounds, interval-scaled coupled iteration in the form the kernels use
(Y stored transposed, Z <- T Z and Y^T <- T Y^T, i.e. M = Z Y -> T M T^T), padding, the final assembly
of T.  It pins the constants the kernels use (stop at 1 - l < 1e-12, fallback above a cond bound of
1e7) and shows the ordering that must NOT be used (Y <- T Y: M -> T Z T Y is no congruence)."""
import numpy as np
import pytest

from oracle import etkf as o_etkf

MAX_COND = 1e7


def ns_transform(C, g, k, rho, KP, ordering="congruence"):
    shift = (k - 1) / rho
    delta_k = max(np.trace(C) / ((k - 1) * k), 0.0) if k > 1 else 0.0
    Ad = C + delta_k + shift * np.eye(k)                        # A' = A + delta U
    rows = np.abs(Ad).sum(axis=1)
    s = min(np.linalg.norm(Ad, 'fro'), rows.max())
    lower = max(shift, (2 * np.diag(Ad) - rows).min())           # Gershgorin, or A' >= A >= shift I
    if not (np.isfinite(s) and s > 0 and s <= MAX_COND * lower):
        return None, 0                                           # the kernels fall back to Jacobi
    Z = np.eye(KP)
    Yt = np.eye(KP)
    Yt[:k, :k] = (Ad / s).T
    l = np.sqrt(lower / s) * (1 - 1e-6)
    I = np.eye(KP)
    for it in range(1, 41):
        al2 = 3.0 / (1.0 + l + l * l)
        al = np.sqrt(al2)
        c1, c2 = 1.5 * al, -0.5 * al * al2
        T = c1 * I + c2 * (Z @ Yt.T)                             # row strips T[R,:] = c1 I + c2 Z[R,:] Y
        Z = T @ Z
        if ordering == "congruence":
            Yt = T @ Yt                                          # Y <- Y T^T: M -> T M T^T
        else:
            Yt = Yt @ T.T                                        # Y <- T Y: M -> T Z T Y
        l = 0.5 * al * l * (3.0 - al2 * l * l)
        if 1.0 - l < 1e-12:
            break
    Zr = Z[:k, :k]
    u = Zr @ g
    wa = Zr @ u / s
    cs = Zr.sum(axis=0)
    T = 1.0 / k + (wa - wa.mean())[:, None] + np.sqrt((k - 1) / s) * (Zr - cs[None, :] / k)
    return T, it


def stats(k, scale, seed, n_y=None):
    rng = np.random.default_rng(seed)
    n_y = n_y or max(3, 3 * k)
    Yp = rng.standard_normal((k, n_y)) * np.logspace(0, scale, n_y)
    Yp -= Yp.mean(axis=0)
    return Yp @ Yp.T, Yp @ rng.standard_normal(n_y)


@pytest.mark.parametrize("k,KP", [(3, 16), (20, 24), (32, 32), (40, 64), (64, 64), (100, 128), (128, 128)])
@pytest.mark.parametrize("rho,scale", [(1.0, 0.0), (1.3, 1.0), (1.0, 2.0), (1.05, 3.0)])
def test_ns_algorithm_matches_eigendecomposition(k, KP, rho, scale):
    C, g = stats(k, scale, seed=17 * k + int(10 * scale))
    T_ref = o_etkf.transform_from_stats(C, g, k, rho)
    T, its = ns_transform(C, g, k, rho, KP)
    assert T is not None and 0 < its < 30
    assert np.max(np.abs(T - T_ref)) / np.max(np.abs(T_ref)) < 1e-11


def test_deflation_and_gershgorin_cut_the_step_count():
    """C3-like statistics (many observations, independent members): the nonzero eigenvalues of C are
    tightly clustered, the known eigenvalue (k-1)/rho is far below them; deflating it and bounding
    the rest with Gershgorin halves the number of steps."""
    rng = np.random.default_rng(0)
    k, n_y = 64, 19661
    Yp = rng.standard_normal((k, n_y))
    Yp -= Yp.mean(axis=0)
    C, g = Yp @ Yp.T, Yp @ rng.standard_normal(n_y)
    T, its = ns_transform(C, g, k, 1.0, 64)
    assert its <= 5
    shift = k - 1.0
    A = C + shift * np.eye(k)
    s = min(np.linalg.norm(A, 'fro'), np.abs(A).sum(axis=1).max())
    assert s / shift > 300                                       # what the plain interval would be
    assert np.max(np.abs(T - o_etkf.transform_from_stats(C, g, k, 1.0))) < 1e-12


def test_fallback_decision_and_nonfinite():
    C, g = stats(64, 5.0, seed=1)
    assert ns_transform(C, g, 64, 1.0, 64)[0] is None           # cond bound > 1e7 -> Jacobi
    C, g = stats(64, 0.0, seed=2)
    C[3, 4] = np.nan
    assert ns_transform(C, g, 64, 1.0, 64)[0] is None


def test_left_multiplying_y_is_unstable():
    """Y <- T Y (with Z <- T Z) loses the coupling at moderate condition numbers: M = Z Y becomes
    T Z T Y instead of the congruence T M T^T the kernels' ordering gives."""
    C, g = stats(64, 3.0, seed=5)
    T_ref = o_etkf.transform_from_stats(C, g, 64, 1.0)
    good, _ = ns_transform(C, g, 64, 1.0, 64, ordering="congruence")
    with np.errstate(all="ignore"):
        bad, _ = ns_transform(C, g, 64, 1.0, 64, ordering="TY")
    err = lambda T: np.max(np.abs(T - T_ref)) / np.max(np.abs(T_ref))
    assert err(good) < 1e-11
    assert not (err(bad) < 1e-6)
