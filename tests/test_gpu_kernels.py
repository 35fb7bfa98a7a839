"""GPU parity tests: every kernel is called through the C ABI (ctypes) and compared with the
CPU oracle on the same seeded inputs.  Tolerances (from BASELINE.json's north star):
  * observation gather / indexing: bit-exact;
  * single analysis: max|dXa| / max|Xa| <= 1e-10 (FP64);
  * RK4 forecast vs oracle RK4 (same integrator, different rounding/FMA order): <= 1e-11 relative
    over one window;
  * multi-cycle deterministic GPU vs oracle-RK4: <= 1e-9 over the first cycles."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import l96 as o_l96, etkf as o_etkf, windows as o_win
from dataassimbench_b200 import _native


@pytest.fixture(scope="module")
def H():
    h = _native.Handle(0)
    yield h
    h.close()


def dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def rel_err(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def make_ensemble(N, k, seed=0, spinup=200):
    rng = np.random.default_rng(seed)
    x = 8.0 + rng.standard_normal(N)
    x = o_l96.rk4_forecast(x, spinup, 0.01)
    return x + rng.standard_normal((k, N)), x


# ------------------------------------------------------------------------- Lorenz-63 members
@pytest.mark.parametrize("k,steps,dt", [(1, 1, 0.01), (7, 25, 0.01), (1000, 100, 0.01), (300, 40, 0.005)])
def test_l63_forecast_matches_oracle_rk4(H, k, steps, dt):
    from oracle import l63 as o_l63
    rng = np.random.default_rng(k + steps)
    X = o_l63.X0_DEFAULT + rng.standard_normal((k, 3))
    last, traj = o_l63.rk4_forecast(X, steps, dt, return_traj=True)
    x_in = dev(X)
    x_out = torch.empty_like(x_in)
    tr = torch.empty((steps, k, 3), dtype=torch.float64, device='cuda')
    H.l63_forecast(x_in, x_out, tr, k, steps, dt)
    H.sync(); torch.cuda.synchronize()
    # rounding differences (FMA contraction) are amplified by the chaotic dynamics: one ulp of noise per
    # step ends at 2e-13 over 1000 members x 100 steps in the oracle itself; 1e-10 leaves room and is
    # still five orders below the method's own truncation error
    assert rel_err(x_out.cpu().numpy(), last) < 1e-10
    assert rel_err(tr.cpu().numpy(), traj[1:]) < 1e-10
    assert np.array_equal(tr[-1].cpu().numpy(), x_out.cpu().numpy())
    # in place, no trajectory, non-default parameters
    par = dict(sigma=9.0, rho=30.0, beta=2.5)
    H.l63_forecast(x_in, x_in, None, k, steps, dt, **par)
    torch.cuda.synchronize()
    assert rel_err(x_in.cpu().numpy(), o_l63.rk4_forecast(X, steps, dt, **par)) < 1e-10


# ------------------------------------------------------------------------------------- K1
@pytest.mark.parametrize("N,k,steps,dt", [
    (5, 8, 10, 0.05), (36, 10, 20, 0.01), (40, 20, 20, 0.01), (1000, 3, 7, 0.01),
    (4096, 16, 20, 0.01), (65536, 4, 20, 0.01), (5000, 2, 70, 0.01), (301, 5, 1, 0.01),
    (4, 2, 3, 0.01)])
def test_forecast_matches_oracle_rk4(H, N, k, steps, dt):
    X, _ = make_ensemble(N, k, seed=N + k, spinup=50)
    ref, ref_traj = o_l96.rk4_forecast(X, steps, dt, return_traj=True)
    x_in = dev(X)
    x_out = torch.empty_like(x_in)
    traj = torch.full((steps, k, N), float('nan'), dtype=torch.float64, device='cuda')
    H.l96_forecast(x_in, x_out, traj, N, k, steps, dt, 8.0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    out = x_out.cpu().numpy()
    assert np.all(np.isfinite(out))
    assert rel_err(out, ref) < 1e-11
    assert rel_err(traj.cpu().numpy(), ref_traj[1:]) < 1e-11
    # without trajectory
    x_out2 = torch.empty_like(x_in)
    H.l96_forecast(x_in, x_out2, None, N, k, steps, dt, 8.0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(x_out, x_out2)


def test_forecast_propagates_nonfinite(H):
    """kernels must propagate inf/NaN faithfully, never mask them."""
    N, k = 256, 2
    X, _ = make_ensemble(N, k, seed=3, spinup=10)
    X[1, 100] = np.inf
    x_in = dev(X)
    x_out = torch.empty_like(x_in)
    H.l96_forecast(x_in, x_out, None, N, k, 5, 0.01, 8.0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    out = x_out.cpu().numpy()
    assert np.all(np.isfinite(out[0]))
    assert not np.all(np.isfinite(out[1]))


def test_forecast_linearity_free_property_full_size(H):
    """C3 size: translation equivariance of the periodic stencil (size-independent property):
    rolling the input along the grid rolls the output, bit for bit."""
    N, k, S = 65536, 64, 20
    rng = np.random.default_rng(7)
    X = 3.0 * rng.standard_normal((k, N)) + 2.0
    x_in = dev(X)
    x_out = torch.empty_like(x_in)
    st = torch.cuda.current_stream().cuda_stream
    H.l96_forecast(x_in, x_out, None, N, k, S, 0.01, 8.0, st)
    shift = 12345
    x_in_r = torch.roll(x_in, shift, dims=1).contiguous()
    x_out_r = torch.empty_like(x_in)
    H.l96_forecast(x_in_r, x_out_r, None, N, k, S, 0.01, 8.0, st)
    torch.cuda.synchronize()
    assert torch.equal(torch.roll(x_out, shift, dims=1), x_out_r)
    # spot-check 2 members against the oracle
    ref = o_l96.rk4_forecast(X[:2], S, 0.01)
    assert rel_err(x_out[:2].cpu().numpy(), ref) < 1e-11


@pytest.mark.parametrize("shape", ["16128", "16256", "17128", "12192", "384", "192"])
@pytest.mark.parametrize("N,k,steps", [(20000, 3, 20), (4099, 2, 7)])
def test_forecast_cta_shapes_agree(H, shape, N, k, steps, monkeypatch):
    """Every CTA shape the autotuner may pick (8 or 16 points per thread, 2-3 or 1 CTA per SM) runs the
    same per-point arithmetic: identical bits, and the oracle tolerance."""
    X, _ = make_ensemble(N, k, seed=N, spinup=50)
    ref = o_l96.rk4_forecast(X, steps, 0.01)
    x_in = dev(X)
    st = torch.cuda.current_stream().cuda_stream
    base = torch.empty_like(x_in)
    H.l96_forecast(x_in, base, None, N, k, steps, 0.01, 8.0, st)
    monkeypatch.setenv("DAB_L96_T", shape)
    out = torch.empty_like(x_in)
    H.l96_forecast(x_in, out, None, N, k, steps, 0.01, 8.0, st)
    torch.cuda.synchronize()
    assert rel_err(out.cpu().numpy(), ref) < 1e-11
    assert torch.equal(out, base)


def run_window_through_cycler(H, X, S, forced=None, monkeypatch=None, dt=0.01):
    """One analysis-free cycle (no observation in the window: T = inflation only, rho = 1 -> identity) through the
    cycler, i.e. through the extended-layout path the tuner serves: returns the forecast of X."""
    k, N = X.shape
    padded = np.zeros((1, 1), dtype=np.int64)                  # the reference's "no observation time" slot
    cfg = _native.CycleConfig(system_dim=N, ensemble_dim=k, n_steps=S, model_dt=dt, forcing=8.0, rho=1.0,
                              obs_error_sd=1.0, n_obs_times=1, n_obs=1, stationary=1, n_cycles=1, t_pad=1,
                              rank=0, world=1)
    if forced is not None:
        monkeypatch.setenv(*forced)
    H.cycler_setup(cfg, np.zeros((1, 1)), np.zeros((1, 1), dtype=np.int64), padded)
    if forced is not None:
        monkeypatch.delenv(forced[0])
    H.cycler_set_state(np.ascontiguousarray(X), True)
    H.cycler_run(0, 1)
    H.sync()
    out = np.empty((k, N))
    H.cycler_get_state(out, True)
    return out, H.cycler_forecast_shape()[0]


@pytest.mark.parametrize("N,k,S", [(40000, 8, 20), (65536, 16, 20), (30000, 12, 7), (9000, 64, 32), (262144, 6, 20)])
def test_forecast_chained_kernel_is_bit_identical(H, N, k, S, monkeypatch):
    """The chained kernel (l96_forecast_v6.cu: tiles take their left boundary from the previous tile's
    history instead of recomputing a halo) against the first design, through the cycler: identical bits, and
    the oracle tolerance on a few members."""
    X, _ = make_ensemble(N, k, seed=N + S, spinup=50)
    X = X + 0.0
    # all-masked window: the analysis is Xb T with T = U + sqrt(rho)(I - U) = I at rho = 1 up to rounding, so
    # compare the two forecast kernels on the analysis the cycler itself produced (same T, same bits)
    base, code0 = run_window_through_cycler(H, X, S, ("DAB_L96_V6", "0"), monkeypatch)
    out, code5 = run_window_through_cycler(H, X, S, ("DAB_L96_V6", "1"), monkeypatch)
    assert code5 == 60000 and code0 < 60000
    assert np.array_equal(out, base)
    out2, code6 = run_window_through_cycler(H, X, S, ("DAB_L96_V6", "2"), monkeypatch)      # 6 groups x 64 threads
    assert code6 == 60001 and np.array_equal(out2, base)
    ref = o_l96.rk4_forecast(X[:2], S, 0.01)
    assert rel_err(out[:2], ref) < 1e-9          # the identity transform adds its own rounding


# ------------------------------------------------------------------------------------- K2
@pytest.mark.parametrize("N,k,n_y", [(36, 10, 18), (40, 20, 20), (65536, 64, 19661), (1000, 7, 1), (50, 3, 0)])
def test_gather_bit_exact(H, N, k, n_y):
    rng = np.random.default_rng(N + n_y)
    X = rng.standard_normal((k, N))
    idx = rng.choice(N, n_y, replace=False, shuffle=False).astype(np.int64)   # unsorted, like Observer
    mask = (rng.random(n_y) > 0.2).astype(np.uint8)
    Xd = dev(X)
    Yb = torch.empty((k, max(n_y, 1)), dtype=torch.float64, device='cuda')[:, :n_y].contiguous()
    st = torch.cuda.current_stream().cuda_stream
    H.obs_gather(Xd, dev(idx) if n_y else None, None, Yb if n_y else None, N, k, n_y, st)
    torch.cuda.synchronize()
    assert np.array_equal(Yb.cpu().numpy(), o_etkf.obs_gather(X, idx))
    if n_y:
        H.obs_gather(Xd, dev(idx), dev(mask), Yb, N, k, n_y, st)
        torch.cuda.synchronize()
        assert np.array_equal(Yb.cpu().numpy(), o_etkf.obs_gather(X, idx) * mask)


# -------------------------------------------------------------------------------- K3 / K4
def analysis_inputs(N, k, n_y, seed, sd=1.0, masked=0.1, spread=1.0):
    rng = np.random.default_rng(seed)
    truth = 3.0 * rng.standard_normal(N) + 2.0
    X = truth + spread * rng.standard_normal((k, N))
    idx = rng.choice(N, n_y, replace=(n_y > N), shuffle=False).astype(np.int64)
    y = truth[idx] + sd * rng.standard_normal(n_y)
    mask = (rng.random(n_y) >= masked).astype(np.uint8)
    return X, y, idx, mask


@pytest.mark.parametrize("N,k,n_y", [(36, 10, 18), (40, 20, 20), (2000, 64, 600), (1500, 128, 1500),
                                     (500, 33, 77), (100, 2, 5), (65536, 64, 19661)])
def test_stats_match_oracle(H, N, k, n_y):
    X, y, idx, mask = analysis_inputs(N, k, n_y, seed=n_y)
    sd = 0.7
    C, g = o_etkf.contraction(X, y, idx, mask / sd ** 2)
    stats = torch.empty(k * k + k, dtype=torch.float64, device='cuda')
    H.etkf_stats(dev(X), dev(y), dev(idx), dev(mask), sd, N, k, n_y, stats,
                 torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    s = stats.cpu().numpy()
    assert rel_err(s[:k * k].reshape(k, k), C) < 1e-12
    assert np.max(np.abs(s[k * k:] - g)) / np.max(np.abs(g)) < 1e-12


@pytest.mark.parametrize("k", [2, 3, 8, 10, 20, 33, 64, 100, 128])
@pytest.mark.parametrize("rho", [1.0, 1.3])
def test_transform_matrix_matches_oracle(H, k, rho):
    rng = np.random.default_rng(k)
    n_y = max(3, 3 * k)
    Yp = rng.standard_normal((k, n_y)) * np.logspace(0, 2, n_y)     # spread of scales
    Yp -= Yp.mean(axis=0)
    C = Yp @ Yp.T
    g = Yp @ rng.standard_normal(n_y)
    T_ref = o_etkf.transform_from_stats(C, g, k, rho)
    stats = dev(np.concatenate([C.ravel(), g]))
    T = torch.empty((k, k), dtype=torch.float64, device='cuda')
    lam = torch.empty(k, dtype=torch.float64, device='cuda')
    sweeps = torch.zeros(1, dtype=torch.int32, device='cuda')
    H.etkf_transform_matrix(stats, k, rho, False, T, lam, sweeps, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert 0 < int(sweeps.item()) < 30
    lam_ref = np.linalg.eigvalsh((k - 1) / rho * np.eye(k) + C)
    assert np.max(np.abs(np.sort(lam.cpu().numpy()) - lam_ref)) / lam_ref.max() < 1e-13
    assert rel_err(T.cpu().numpy(), T_ref) < 1e-11


def _random_stats(k, rho, scale, seed):
    rng = np.random.default_rng(seed)
    n_y = max(3, 3 * k)
    Yp = rng.standard_normal((k, n_y)) * np.logspace(0, scale, n_y)
    Yp -= Yp.mean(axis=0)
    C = Yp @ Yp.T
    g = Yp @ rng.standard_normal(n_y)
    return C, g


@pytest.mark.parametrize("k", [2, 3, 8, 10, 20, 32, 33, 40, 64, 65, 100, 128])
@pytest.mark.parametrize("rho,scale", [(1.0, 0.0), (1.3, 1.0), (1.0, 2.0), (1.05, 3.0)])
def test_transform_matrix_newton_schulz_path(H, k, rho, scale):
    """No eigenvalue request: the Newton-Schulz kernels -- on a thread-block cluster for 32 < k <= 128
    (csrc/etkf_ns.cu), inside one CTA for k <= 32 (csrc/etkf_ns_device.cuh).  Same tolerance as the
    eigensolver path, and agreement with it."""
    C, g = _random_stats(k, rho, scale, seed=1000 * k + int(10 * scale))
    T_ref = o_etkf.transform_from_stats(C, g, k, rho)
    stats = dev(np.concatenate([C.ravel(), g]))
    st = torch.cuda.current_stream().cuda_stream
    T = torch.full((k, k), float('nan'), dtype=torch.float64, device='cuda')
    steps = torch.zeros(1, dtype=torch.int32, device='cuda')
    H.set_eig_method(H.EIG_AUTO)
    H.etkf_transform_matrix(stats, k, rho, False, T, None, steps, st)
    torch.cuda.synchronize()
    cond = np.linalg.cond((k - 1) / rho * np.eye(k) + C)
    assert 0 < int(steps.item()) < 30, (int(steps.item()), cond)
    assert rel_err(T.cpu().numpy(), T_ref) < 1e-11, cond
    T_j = torch.empty_like(T)
    H.set_eig_method(H.EIG_JACOBI)
    try:
        H.etkf_transform_matrix(stats, k, rho, False, T_j, None, None, st)
        torch.cuda.synchronize()
    finally:
        H.set_eig_method(H.EIG_AUTO)
    assert rel_err(T.cpu().numpy(), T_j.cpu().numpy()) < 1e-11
    # bit-reproducible from call to call (the sharded cycler replicates K4 on every rank)
    T2 = torch.empty_like(T)
    H.etkf_transform_matrix(stats, k, rho, False, T2, None, None, st)
    torch.cuda.synchronize()
    assert torch.equal(T, T2)


@pytest.mark.parametrize("k", [20, 64, 128])
def test_transform_matrix_newton_schulz_falls_back(H, k):
    """Ill-conditioned (bound on cond(A) > 1e7) or non-finite statistics: the Jacobi eigensolver
    enqueued behind the Newton-Schulz kernel takes over; the result keeps the oracle tolerance."""
    rho = 1.0
    C, g = _random_stats(k, rho, 4.0, seed=k)
    assert np.linalg.cond((k - 1) / rho * np.eye(k) + C) > 3e7      # cond <= the kernel's bound
    T_ref = o_etkf.transform_from_stats(C, g, k, rho)
    st = torch.cuda.current_stream().cuda_stream
    T = torch.full((k, k), float('nan'), dtype=torch.float64, device='cuda')
    sweeps = torch.zeros(1, dtype=torch.int32, device='cuda')
    H.etkf_transform_matrix(dev(np.concatenate([C.ravel(), g])), k, rho, False, T, None, sweeps, st)
    torch.cuda.synchronize()
    assert int(sweeps.item()) > 0
    assert rel_err(T.cpu().numpy(), T_ref) < 1e-10                  # eigh itself is ~cond*eps here
    # a well-conditioned problem right after it: the fallback flag is cleared again
    C, g = _random_stats(k, rho, 0.0, seed=k + 1)
    T_ref = o_etkf.transform_from_stats(C, g, k, rho)
    H.etkf_transform_matrix(dev(np.concatenate([C.ravel(), g])), k, rho, False, T, None, None, st)
    torch.cuda.synchronize()
    assert rel_err(T.cpu().numpy(), T_ref) < 1e-11
    bad = np.concatenate([C.ravel(), g])
    bad[5] = np.nan
    T.zero_()
    H.etkf_transform_matrix(dev(bad), k, rho, False, T, None, None, st)
    torch.cuda.synchronize()
    assert not np.all(np.isfinite(T.cpu().numpy()))


def test_transform_matrix_empty_R_and_no_obs(H):
    k = 12
    T = torch.empty((k, k), dtype=torch.float64, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    H.etkf_transform_matrix(None, k, 1.0, True, T, None, None, st)
    torch.cuda.synchronize()
    assert np.allclose(T.cpu().numpy(), 1.0 / k)
    # zero statistics (all rows masked): T = U + sqrt(rho) (I - U)
    stats = torch.zeros(k * k + k, dtype=torch.float64, device='cuda')
    H.etkf_transform_matrix(stats, k, 1.21, False, T, None, None, st)
    torch.cuda.synchronize()
    U = np.ones((k, k)) / k
    assert np.allclose(T.cpu().numpy(), U + 1.1 * (np.eye(k) - U), atol=1e-14)


# ------------------------------------------------------------------------- single analysis
@pytest.mark.parametrize("N,k,n_y,rho", [(36, 10, 18, 1.0), (40, 20, 20, 1.1), (2000, 64, 600, 1.0),
                                         (1500, 128, 1500, 1.05), (5, 8, 6, 1.0)])
def test_analysis_matches_literal_reference_formula(H, N, k, n_y, rho):
    """P2 of SURVEY 8d: against the oracle's LITERAL restatement of _compute_analysis
    (dense H, dense R, pinv, sqrtm) at every shape the literal formula can hold."""
    X, y, idx, mask = analysis_inputs(N, k, n_y, seed=N * 7 + k)
    if n_y > N:
        idx = idx % N
    sd = 1.5
    Hm = o_etkf.calc_default_H(n_y, N, idx)
    Hm = np.where(mask, Hm.T, 0).T
    ref = o_etkf.analysis_literal(X.T, y, Hm, o_etkf.calc_default_R(n_y, sd), rho).T
    Xa = torch.empty((k, N), dtype=torch.float64, device='cuda')
    H.etkf_analysis(dev(X), Xa, dev(y), dev(idx), dev(mask), sd, rho, N, k, n_y,
                    torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rel_err(Xa.cpu().numpy(), ref) < 1e-10


@pytest.mark.parametrize("N,k,n_y", [(65536, 64, 19661), (131072, 128, 131072)])
def test_analysis_matches_sparse_oracle_large(H, N, k, n_y):
    X, y, idx, mask = analysis_inputs(N, k, n_y, seed=5, masked=0.0)
    ref = o_etkf.analysis_sparse(X, y, idx, mask.astype(np.float64), 1.0)
    Xa = torch.empty((k, N), dtype=torch.float64, device='cuda')
    H.etkf_analysis(dev(X), Xa, dev(y), dev(idx), None, 1.0, 1.0, N, k, n_y,
                    torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rel_err(Xa.cpu().numpy(), ref) < 1e-10


@pytest.mark.parametrize("N,k,frac", [(1048576, 128, 1.0), (4194304, 64, 0.3)])
def test_analysis_full_size_c5_c4(H, N, k, frac):
    """BASELINE configs[4] and [3] at FULL size (contract_kernel<16>, > 2^31-byte ensembles): one analysis against
    the sparse oracle, every element compared."""
    rng = np.random.default_rng(N % 1000 + k)
    truth = 3.0 * rng.standard_normal(N) + 2.0
    X = truth + rng.standard_normal((k, N))
    n_y = int(round(frac * N))
    idx = np.sort(rng.choice(N, n_y, replace=False, shuffle=False)) if n_y == N else rng.choice(N, n_y, replace=False, shuffle=False)
    idx = idx.astype(np.int64)
    y = truth[idx] + rng.standard_normal(n_y)
    ref = o_etkf.analysis_sparse(X, y, idx, np.ones(n_y), 1.0)
    Xd = dev(X)
    Xa = torch.empty((k, N), dtype=torch.float64, device='cuda')
    H.etkf_analysis(Xd, Xa, dev(y), dev(idx), None, 1.0, 1.0, N, k, n_y, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = Xa.cpu().numpy()
    del Xd, Xa
    err = 0.0
    for j in range(k):                                   # row by row: no second 2 GB temporary
        err = max(err, float(np.max(np.abs(got[j] - ref[j]))))
    assert err / float(np.max(np.abs(ref))) < 1e-10


def test_cycler_ten_cycles_full_c3_matches_oracle_rk4(H):
    """BASELINE configs[2] at full size (N = 65 536, k = 64, 19 661 stationary observations, S = 20): ten cycles
    against the oracle cycling with the same RK4 -- analyses of the first cycles to 1e-9, the last one within the
    growth the chaotic dynamics allow (errors double every ~2 cycles)."""
    N, k, n_obs, n_cycles, S, sd, rho = 65536, 64, 19661, 10, 20, 1.0, 1.0
    case = cycler_case(N, k, n_obs, n_cycles, S, sd, rho, True, seed=11)
    out, final, _ = run_gpu_cycler(H, case, N, k, n_cycles, sd, rho, True)
    X = case['X0'].copy()
    w = np.full(n_obs, 1.0 / sd ** 2)
    errs = []
    for c in range(n_cycles):
        Xa = o_etkf.analysis_sparse(X, case['vals'][c], case['sysidx'][c], w, rho)
        errs.append(rel_err(out[c], Xa))
        X = o_l96.rk4_forecast(Xa, S, case['dt'])
    assert np.all(np.isfinite(out))
    assert max(errs[:4]) < 1e-10, errs
    assert max(errs) < 1e-8, errs
    assert rel_err(final, X) < 1e-8


def test_analysis_empty_obs_collapses_to_mean(H):
    N, k = 64, 6
    X, *_ = analysis_inputs(N, k, 4, seed=1)
    Xa = torch.empty((k, N), dtype=torch.float64, device='cuda')
    H.etkf_analysis(dev(X), Xa, None, None, None, 1.0, 1.0, N, k, 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.allclose(Xa.cpu().numpy(), np.tile(X.mean(axis=0), (k, 1)), atol=1e-13)


# ------------------------------------------------------------------------------ the cycler
def cycler_case(N, k, n_obs, n_cycles, S, sd, rho, stationary, seed, obs_every=None, dt=0.01):
    """Synthetic experiment following SURVEY 8d; returns oracle inputs."""
    rng = np.random.default_rng(seed)
    X0, truth0 = make_ensemble(N, k, seed=seed, spinup=300)
    window = S * dt
    obs_every = obs_every or S
    n_truth = n_cycles * S + 1
    _, truth = o_l96.rk4_forecast(truth0, n_truth - 1, dt, return_traj=True)
    t_all = np.arange(n_truth) * dt
    obs_tidx = np.arange(0, n_truth, obs_every)
    obs_times = t_all[obs_tidx]
    if stationary:
        locs = rng.choice(N, n_obs, replace=False, shuffle=False)
        sysidx = np.tile(locs, (len(obs_tidx), 1)).astype(np.int64)
        vals = truth[obs_tidx][:, locs] + sd * rng.standard_normal((len(obs_tidx), n_obs))
    else:
        counts = rng.integers(max(1, n_obs // 2), n_obs + 1, size=len(obs_tidx))
        width = int(counts.max())
        sysidx = np.zeros((len(obs_tidx), width), dtype=np.int64)
        vals = np.full((len(obs_tidx), width), np.nan)
        for i, (ti, cnt) in enumerate(zip(obs_tidx, counts)):
            l = rng.choice(N, cnt, replace=False, shuffle=False)
            sysidx[i, :cnt] = l
            vals[i, :cnt] = truth[ti, l] + sd * rng.standard_normal(cnt)
    return dict(X0=X0, truth=truth, obs_times=obs_times, sysidx=sysidx, vals=vals, window=window,
                start=0.0, dt=dt, S=S)


def run_gpu_cycler(H, case, N, k, n_cycles, sd, rho, stationary, return_traj=False):
    S, dt = case['S'], case['dt']
    times = o_win.get_all_times(case['start'], case['window'], n_cycles)
    filt = o_win.get_obs_indices(times, case['obs_times'], case['window'])
    padded = np.ascontiguousarray(o_win.pad_time_indices(filt))
    cfg = _native.CycleConfig(system_dim=N, ensemble_dim=k, n_steps=S, model_dt=dt, forcing=8.0, rho=rho,
                              obs_error_sd=sd, n_obs_times=case['vals'].shape[0], n_obs=case['vals'].shape[1],
                              stationary=int(stationary), n_cycles=n_cycles, t_pad=padded.shape[1],
                              rank=0, world=1)
    vals = np.ascontiguousarray(case['vals'])
    sysidx = np.ascontiguousarray(case['sysidx'])
    H.cycler_setup(cfg, vals, sysidx, padded)
    H.cycler_set_state(np.ascontiguousarray(case['X0']), True)
    out = torch.empty((n_cycles, k, N), dtype=torch.float64).pin_memory()
    traj = torch.empty((n_cycles, S, k, N), dtype=torch.float64, device='cuda') if return_traj else None
    H.cycler_run(0, n_cycles, out, traj)
    H.sync()
    final = np.empty((k, N))
    H.cycler_get_state(final, True)
    return out.numpy().copy(), final, (traj.cpu().numpy() if return_traj else None)


@pytest.mark.parametrize("N,k,n_obs,n_cycles,S,stationary,obs_every", [
    (36, 10, 18, 20, 20, True, None),
    (40, 20, 20, 20, 20, False, None),
    (40, 20, 40, 40, 5, True, None),          # converged regime
    (256, 8, 64, 10, 10, False, 5),           # two obs times per window (T_pad = 2)
    (2048, 64, 600, 5, 20, True, None),
])
def test_cycler_matches_oracle_rk4(H, N, k, n_obs, n_cycles, S, stationary, obs_every):
    sd, rho = 1.0, 1.05
    case = cycler_case(N, k, n_obs, n_cycles, S, sd, rho, stationary, seed=N + k, obs_every=obs_every)
    fc = lambda x, n: o_l96.generate_rk4(x, n, case['dt'])
    ref = o_etkf.cycle(case['X0'], case['start'], case['vals'], case['obs_times'], case['sysidx'],
                       n_cycles, sd, fc, case['dt'], stationary_observers=stationary,
                       analysis_window=case['window'], return_forecast=True, rho=rho, literal=False)
    out, final, traj = run_gpu_cycler(H, case, N, k, n_cycles, sd, rho, stationary, return_traj=True)
    assert np.all(np.isfinite(out))
    # analyses = cycle_timestep 0 of every cycle
    assert rel_err(out, ref[:, :, 0, :]) < 1e-9
    # forecast steps 1..S-1 (the reference drops the last point, which is the next background)
    gpu_traj = traj.transpose(0, 2, 1, 3)                       # [cycle, member, step, N]
    assert rel_err(gpu_traj[:, :, :S - 1, :], ref[:, :, 1:, :]) < 1e-9


@pytest.mark.parametrize("N,k,n_obs", [(36, 10, 18), (6000, 64, 1800), (3000, 128, 3000)])
def test_cycler_dependent_launch_is_bit_identical(H, N, k, n_obs):
    """Programmatic dependent launch (dab_set_pdl) only moves where a kernel waits for its
    predecessor: the analyses of a multi-cycle run must not change by a single bit, run after run."""
    n_cycles, S, sd, rho = 10, 20, 1.0, 1.03
    case = cycler_case(N, k, n_obs, n_cycles, S, sd, rho, True, seed=N + 3 * k)
    try:
        H.set_pdl(False)
        base, base_final, _ = run_gpu_cycler(H, case, N, k, n_cycles, sd, rho, True)
        H.set_pdl(True)
        for _ in range(3):
            out, final, _ = run_gpu_cycler(H, case, N, k, n_cycles, sd, rho, True)
            assert np.array_equal(out, base)
            assert np.array_equal(final, base_final)
    finally:
        H.set_pdl(True)
    assert np.all(np.isfinite(base))


def test_cycler_long_window_chunked(H):
    """n_steps above the per-launch limit exercises the chunked forecast path."""
    N, k, n_cycles, S = 512, 6, 3, 70
    case = cycler_case(N, k, 100, n_cycles, S, 1.0, 1.0, True, seed=11)
    fc = lambda x, n: o_l96.generate_rk4(x, n, case['dt'])
    ref = o_etkf.cycle(case['X0'], 0.0, case['vals'], case['obs_times'], case['sysidx'], n_cycles, 1.0, fc,
                       case['dt'], analysis_window=case['window'], return_forecast=False, literal=False)
    out, _, _ = run_gpu_cycler(H, case, N, k, n_cycles, 1.0, 1.0, True)
    assert rel_err(out, ref) < 1e-9


def test_cycle_host_entry_point(H):
    """dab_etkf_cycle_host_f64: host buffers in, host buffers out (the e2e call)."""
    N, k, n_cycles, S = 1024, 16, 5, 20
    sd, rho = 1.0, 1.0
    case = cycler_case(N, k, 300, n_cycles, S, sd, rho, True, seed=2)
    fc = lambda x, n: o_l96.generate_rk4(x, n, case['dt'])
    ref = o_etkf.cycle(case['X0'], 0.0, case['vals'], case['obs_times'], case['sysidx'], n_cycles, sd, fc,
                       case['dt'], analysis_window=case['window'], return_forecast=False, rho=rho, literal=False)
    times = o_win.get_all_times(0.0, case['window'], n_cycles)
    padded = np.ascontiguousarray(o_win.pad_time_indices(o_win.get_obs_indices(times, case['obs_times'], case['window'])))
    cfg = _native.CycleConfig(system_dim=N, ensemble_dim=k, n_steps=S, model_dt=case['dt'], forcing=8.0, rho=rho,
                              obs_error_sd=sd, n_obs_times=case['vals'].shape[0], n_obs=case['vals'].shape[1],
                              stationary=1, n_cycles=n_cycles, t_pad=padded.shape[1], rank=0, world=1)
    out = np.empty((n_cycles, k, N))
    H.etkf_cycle_host(cfg, np.ascontiguousarray(case['X0']), np.ascontiguousarray(case['vals']),
                      np.ascontiguousarray(case['sysidx']), padded, out)
    assert rel_err(out, ref) < 1e-9
    assert H.kernel_launches() == 5 * n_cycles


@pytest.mark.parametrize("N,k", [(4096, 16), (1000, 33)])
def test_cycler_scores_analyses_on_device(H, N, k):
    """dab_cycler_set_truth: the sum of squared errors of the ensemble-mean analysis against a
    device-resident truth, per cycle, equals dabench.metrics-style mse of the returned analyses."""
    from dataassimbench_b200 import metrics
    n_cycles, S, sd, rho = 5, 20, 1.0, 1.02
    case = cycler_case(N, k, N // 3, n_cycles, S, sd, rho, True, seed=N)
    times = o_win.get_all_times(case['start'], case['window'], n_cycles)
    padded = np.ascontiguousarray(o_win.pad_time_indices(o_win.get_obs_indices(times, case['obs_times'], case['window'])))
    cfg = _native.CycleConfig(system_dim=N, ensemble_dim=k, n_steps=S, model_dt=case['dt'], forcing=8.0, rho=rho,
                              obs_error_sd=sd, n_obs_times=case['vals'].shape[0], n_obs=case['vals'].shape[1],
                              stationary=1, n_cycles=n_cycles, t_pad=padded.shape[1], rank=0, world=1)
    H.cycler_setup(cfg, np.ascontiguousarray(case['vals']), np.ascontiguousarray(case['sysidx']), padded)
    H.cycler_set_state(np.ascontiguousarray(case['X0']), True)
    truth = np.ascontiguousarray(case['truth'][::S][:n_cycles])            # truth at the analysis times
    truth_dev = dev(truth)
    sse = torch.full((n_cycles,), float('nan'), dtype=torch.float64, device='cuda')
    H.cycler_set_truth(truth_dev, sse)
    out = torch.empty((n_cycles, k, N), dtype=torch.float64).pin_memory()
    H.cycler_run(0, n_cycles, out)
    H.sync()
    H.cycler_set_truth(None, None)
    got = metrics.rmse_from_sse(sse.cpu().numpy(), N)
    ref = np.array([metrics.rmse(out[c].numpy().mean(axis=0), truth[c]) for c in range(n_cycles)])
    assert np.max(np.abs(got - ref) / ref) < 1e-12


@pytest.mark.parametrize("forced", [("DAB_L96_V6", "1"), ("DAB_L96_V6", "2"), ("DAB_L96_V3", "20"), ("DAB_L96_T", "256"), None])
@pytest.mark.parametrize("N,k,n_obs", [(8192, 16, 2500), (40000, 64, 40000), (65536, 8, 3)])
def test_cycler_compact_yb_matches_gather(H, N, k, n_obs, forced, monkeypatch):
    """Stationary networks: the forecast kernels also write the observed columns into a compact Yb and the next
    contraction reads that (etkf_contract_dense.cu) instead of gathering from the state.  Same experiment with the
    route switched off (DAB_YB_EMIT=0, the default: measured slower at C3, profiles/yb_emit_r02.md): the summation
    order of the statistics differs, nothing else."""
    n_cycles, S, sd, rho = 4, 20, 1.0, 1.02
    case = cycler_case(N, k, n_obs, n_cycles, S, sd, rho, True, seed=N + k + n_obs)
    if forced is not None:
        if forced[0] != "DAB_L96_V6":
            monkeypatch.setenv("DAB_L96_V6", "0")
        if forced[0] == "DAB_L96_T":
            monkeypatch.setenv("DAB_L96_V3", "0")
        monkeypatch.setenv(*forced)
    monkeypatch.setenv("DAB_YB_EMIT", "0")
    base, base_final, _ = run_gpu_cycler(H, case, N, k, n_cycles, sd, rho, True)
    assert H.cycler_compact_contractions() == 0
    monkeypatch.setenv("DAB_YB_EMIT", "1")
    out, final, _ = run_gpu_cycler(H, case, N, k, n_cycles, sd, rho, True)
    # the first background came from set_state; networks of < 64 unsorted rows keep their order and the gather
    assert H.cycler_compact_contractions() >= (n_cycles - 2 if n_obs >= 64 else 0)
    assert np.all(np.isfinite(base))
    assert rel_err(out, base) < 1e-11 and rel_err(final, base_final) < 1e-11
