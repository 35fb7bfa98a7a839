"""Surface-QG generator, CPU side: the oracle against the one numerical golden the reference holds for SQGTurb
(tests/observer_base_test.py:240-261), and the package's host mirror (tables, helpers, host rhs) against the oracle."""
import os

import numpy as np
import pytest

from oracle import sqgturb as o_sqg

DEFAULT_PV = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'dataassimbench_b200',
                          '_suppl_data', 'sqgturb_57600steps.npy')


def reference_observer_draws(N, n_times, seed=99, p_time=0.3, p_loc=0.01, sd=25.):
    """The order in which Observer.observe consumes default_rng(99) for a (level, x, y) state
    (dabench/observer/_observer.py:192-234, :331-348): times, count, one `choice` per coordinate WITH replacement
    (more than one non-time coordinate, :223-226), errors."""
    rng = np.random.default_rng(seed)
    tsel = np.where(rng.binomial(1, p=p_time, size=n_times).astype(bool))[0]
    count = int(np.sum(rng.binomial(1, p=p_loc, size=2 * N * N)))
    lev = rng.choice(np.arange(2), size=count, replace=True, shuffle=False)
    xx = rng.choice(np.arange(N), size=count, replace=True, shuffle=False)
    yy = rng.choice(np.arange(N), size=count, replace=True, shuffle=False)
    err = rng.normal(loc=0.0, scale=sd, size=(1, len(tsel), count))
    return tsel, count, lev, xx, yy, err


def test_oracle_reproduces_the_reference_sqgturb_observer_golden():
    pv = np.load(DEFAULT_PV)
    assert pv.shape == (2, 96, 96) and pv.dtype == np.float32
    values, times = o_sqg.generate(pv, 10, precision='single')          # SQGTurb().generate(n_steps=10)
    assert values.shape == (11, 2, 96, 96)
    tsel, count, lev, xx, yy, err = reference_observer_draws(96, 11)
    assert count == 204 and len(tsel) == 2                               # obs_vec['pv'].shape == (2, 204)
    assert np.allclose(times[tsel], [2700., 9000.])
    assert lev[0] * 96 * 96 + xx[0] * 96 + yy[0] == 10130                # system_index[0, 0, 0]
    assert err[0, 1, 187] == pytest.approx(25.714826330507496)
    observed = values[tsel[0]][lev[44], xx[44], yy[44]] + err[0, 0, 44]
    assert observed == pytest.approx(2128.20749725)                      # rel 1e-6, as the reference asserts it
    # the double-precision tables give the same trajectory to float32 accuracy
    v64, _ = o_sqg.generate(pv, 10, precision='double')
    assert v64[tsel[0]][lev[44], xx[44], yy[44]] + err[0, 0, 44] == pytest.approx(2128.20749725)


def test_oracle_scan_quirk_and_shapes():
    rng = np.random.default_rng(1)
    pv = rng.standard_normal((2, 16, 16)) * 100
    T = o_sqg.tables(16, 'double')
    x0 = np.fft.rfft2(pv)
    values, times = o_sqg.generate(pv, 3, precision='double')
    assert values.shape == (4, 2, 16, 16) and np.allclose(times, [0, 900, 1800, 2700])
    one = o_sqg.rk4_step(T, x0)
    assert np.allclose(values[0], np.fft.irfft2(one), rtol=0, atol=1e-9)  # row 0 is ONE step past x0
    # the truncated Jacobian keeps no Nyquist column and the tendency is Hermitian-consistent
    d = o_sqg.rhs(T, x0)
    assert d.shape == (2, 16, 9)
    with pytest.raises(IndexError):
        o_sqg.tables(16, 'double', symmetric=False)


@pytest.mark.parametrize("precision", ['single', 'double'])
@pytest.mark.parametrize("r", [0.0, 1.0e-6])
def test_host_mirror_tables_and_rhs_match_the_oracle(precision, r):
    from dataassimbench_b200.data import SQGTurb
    rng = np.random.default_rng(2)
    ft = np.float32 if precision == 'single' else np.float64
    pv = (rng.standard_normal((2, 32, 32)) * 500).astype(ft)
    g = SQGTurb(pv=pv, precision=precision, r=r, delta_t=600)
    T = o_sqg.tables(32, precision, r=r, delta_t=600)
    assert g.system_dim == 2048 and g.original_dim == (2, 32, 32) and g.x0.shape == (2 * 32 * 17,)
    for name, mine in [('Hovermu', g.Hovermu), ('tanhmu', g.tanhmu), ('sinhmu', g.sinhmu), ('ksqlsq', g.ksqlsq),
                       ('hyperdiff', g.hyperdiff), ('pvspec_eq', g.pvspec_eq), ('ik_pad', g.ik_pad),
                       ('il_pad', g.il_pad), ('ik', g.ik), ('il', g.il), ('pvbar', g.pvbar)]:
        assert mine.dtype == T[name].dtype, name
        assert np.array_equal(mine, T[name]), name
    assert g.ekman == T['ekman']
    x = g.map1dto2d(g.x0)
    assert np.array_equal(g.rhs(x), o_sqg.rhs(T, x))
    assert np.array_equal(g.x0_gridded, g.ifft2(x))
    assert g.fft2_2dto1d(pv).shape == g.x0.shape


def test_host_mirror_argument_errors():
    from dataassimbench_b200.data import SQGTurb
    with pytest.raises(ValueError, match='1st dim of pv should be 2'):
        SQGTurb(pv=np.zeros((3, 8, 8)))
    with pytest.raises(ValueError, match='must be even'):
        SQGTurb(pv=np.zeros((2, 7, 7)))
    with pytest.raises(ValueError, match='Precision'):
        SQGTurb(pv=np.zeros((2, 8, 8)), precision='half')
    with pytest.raises(IndexError):
        SQGTurb(pv=np.zeros((2, 8, 8)), symmetric=False)
    g = SQGTurb()
    assert g.N == 96 and g.system_dim == 18432 and g.x0.dtype == np.complex64


def test_observer_on_a_gridded_state_draws_like_the_reference():
    """Observer on `pv[time, level, x, y]` (the reference's test_obs_sqgturb, tests/observer_base_test.py:240-261):
    the seeded draws -- times, count, one `choice` per coordinate with replacement, errors -- reproduce the
    reference's goldens; the values are the flattened state at those indices plus the errors."""
    import dataassimbench_b200 as dab
    from dataassimbench_b200 import _xr
    rng = np.random.default_rng(0)
    pv = rng.standard_normal((11, 2, 96, 96))
    ds = _xr.new_dataset({'pv': (('time', 'level', 'x', 'y'), pv)},
                         coords={'time': np.arange(11) * 900.0, 'level': np.arange(2), 'x': np.arange(96),
                                 'y': np.arange(96)}, attrs={'system_dim': 18432, 'delta_t': 900.0})
    obs = dab.observer.Observer(ds, random_time_density=0.3, random_location_density=0.01, error_sd=25.).observe()
    assert obs['pv'].shape == (2, 204)
    assert np.allclose(np.asarray(obs['time'].values), [2700., 9000.])
    sysidx = np.asarray(obs['system_index'].values)
    errors = np.asarray(obs['errors'].values)
    assert sysidx.shape == (1, 2, 204) and sysidx[0, 0, 0] == 10130
    assert errors[0, 1, 187] == pytest.approx(25.714826330507496)
    vals = np.asarray(obs['pv'].values)
    flat = pv.reshape(11, -1)
    assert vals[1, 123] == pytest.approx(flat[10, sysidx[0, 1, 123]] + errors[0, 1, 123])
    assert np.array_equal(vals, flat[[3, 10]][:, sysidx[0, 0]] + errors[0])
    tsel, count, lev, xx, yy, err = reference_observer_draws(96, 11)
    assert np.array_equal(sysidx[0, 0], lev * 96 * 96 + xx * 96 + yy) and np.array_equal(errors, err)


def test_dab_accessor_flatten_and_split():
    """`ds.dab.flatten()` / `ds.dab.split_train_val_test()` (dabench/data/_xarray_accessor.py:23-51) on the Dataset
    the generators return."""
    from dataassimbench_b200 import _xr
    pv = np.arange(5 * 2 * 4 * 4, dtype=np.float64).reshape(5, 2, 4, 4)
    ds = _xr.new_dataset({'pv': (('time', 'level', 'x', 'y'), pv)},
                         coords={'time': np.arange(5) * 900.0, 'level': np.arange(2), 'x': np.arange(4), 'y': np.arange(4)},
                         attrs={'system_dim': 32})
    flat = ds.dab.flatten()
    assert flat.shape == (5, 32) and np.array_equal(np.asarray(flat.values), pv.reshape(5, -1))
    one = ds.isel(time=2).dab.flatten()
    assert one.shape == (32,) or one.shape == (1, 32)
    a, b, c = ds.dab.split_train_val_test([2, 2, 1])
    assert a.sizes['time'] == 2 and b.sizes['time'] == 2 and c.sizes['time'] == 1
    a, b = ds.dab.split_train_val_test([0.6, 0.4])
    assert a.sizes['time'] == 3 and b.sizes['time'] == 2
    x = _xr.new_dataset({'x': (('time', 'index'), np.ones((3, 7)))}, coords={'time': np.arange(3.0), 'index': np.arange(7)})
    assert x.dab.flatten().shape == (3, 7)
