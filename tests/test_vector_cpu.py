"""`dataassimbench_b200.vector` against the behaviour of `dabench.vector` (dabench/vector/_vector.py:38-151,
_state_vector.py:36-99, _obs_vector.py:48-185), which cannot be imported here (it imports jax): the cases below
restate what those lines do."""
import numpy as np
import pytest

from dataassimbench_b200.vector import StateVector, ObsVector


def make_state(T=6, N=4):
    return StateVector(system_dim=N, delta_t=0.1, values=np.arange(T * N, dtype=float).reshape(T, N),
                       times=0.1 * np.arange(T))


def test_state_vector_defaults_and_views():
    sv = StateVector(system_dim=6, original_dim=(2, 3), values=np.arange(12.0).reshape(2, 6))
    assert np.array_equal(sv.times, [0, 1]) and sv.time_dim == 2           # _vector.py:28-35
    assert np.array_equal(sv.xi, sv.values[-1])                            # _state_vector.py:66-70
    assert sv.xi_gridded.shape == (2, 3) and sv.values_gridded.shape == (2, 2, 3)
    assert str(sv) == f'Current State = {sv.xi}, Timesteps = 2'
    sv.xi = [1, 2, 3, 4, 5, 6]
    assert isinstance(sv.xi, np.ndarray) and sv.xi_gridded[1, 2] == 6
    empty = StateVector(system_dim=3)
    assert empty.values is None and empty.times is None and empty.time_dim is None
    assert empty.xi is None and empty.xi_gridded is None and empty.values_gridded is None
    assert empty.original_dim == 3                                          # _state_vector.py:46-49
    with pytest.raises(AttributeError, match='does not contain any data values'):
        empty[0]


def test_vector_getitem_returns_independent_copies():
    sv = make_state()
    a = sv[1:5:2]
    assert np.array_equal(a.values, sv.values[1:5:2]) and np.allclose(a.times, [0.1, 0.3]) and a.time_dim == 2
    a.values[0, 0] = -1.0
    assert sv.values[1, 0] == 4.0                                           # deep copy (_vector.py:44)
    one = sv[3]
    assert one.time_dim == 1 and np.array_equal(one.values, sv.values[3])   # _vector.py:54-56
    picked = sv[np.array([0, 5])]
    assert picked.time_dim == 2 and np.array_equal(picked.values, sv.values[[0, 5]])
    masked = sv[np.array([True, False, False, True, False, False])]
    assert np.array_equal(masked.values, sv.values[[0, 3]])
    assert sv.time_dim == 6 and sv.values.shape == (6, 4)


@pytest.mark.parametrize("kw,expect", [
    (dict(), [1, 2, 3]),                                        # inclusive both ends
    (dict(inclusive=False), [2]),
    (dict(start_inclusive=True), [1, 2]),                       # naming one end switches the other off
    (dict(end_inclusive=True), [2, 3]),
    (dict(start_inclusive=True, end_inclusive=True), [1, 2, 3]),
    (dict(start_inclusive=False, end_inclusive=False), [2]),
])
def test_filter_times_end_point_rules(kw, expect):
    sv = make_state()
    # 0.1 and 0.30000000000000004 are "equal" to the end points through isclose(rtol=0) (_vector.py:128-147)
    out = sv.filter_times(0.1, 0.3, **kw)
    assert np.array_equal(out.values, sv.values[expect])
    assert out.time_dim == len(expect)


def test_filter_times_open_ends_and_datetimes():
    sv = make_state()
    assert sv.filter_times(None, 0.2).time_dim == 3
    assert sv.filter_times(0.35, None).time_dim == 2
    assert sv.filter_times(None, None).time_dim == 6
    t = np.array(['2020-01-01', '2020-01-02', '2020-01-03'], dtype='datetime64[D]')
    dv = StateVector(system_dim=1, values=np.arange(3.0)[:, None], times=t)
    out = dv.filter_times(np.datetime64('2020-01-02'), np.datetime64('2020-01-03'), start_inclusive=False,
                          end_inclusive=True)
    assert np.array_equal(out.times, t[2:])


def test_obs_vector_bookkeeping_and_slicing():
    vals = np.array([1.0, 2.0, 3.0, 4.0])
    ov = ObsVector(values=vals, times=[0.0, 0.0, 0.5, 0.5], location_indices=np.array([3, 7, 3, 7]),
                   errors=[0.1, -0.1, 0.2, -0.2], coords=np.array([[3.0], [7.0], [3.0], [7.0]]),
                   error_sd=0.3, error_dist='normal')
    assert ov.num_obs == 4 and np.array_equal(ov.obs_dims, [1, 1, 1, 1])    # _obs_vector.py:138-147
    assert ov.stationary_observers is True and ov.error_sd == 0.3
    late = ov.filter_times(0.5, 0.5)
    assert late.num_obs == 2 and np.array_equal(late.values, [3.0, 4.0])
    assert np.array_equal(late.location_indices, [3, 7]) and np.array_equal(late.errors, [0.2, -0.2])
    assert late.coords.shape == (2, 1)
    head = ov[:3]
    assert head.num_obs == 3 and np.array_equal(head.obs_dims, [1, 1, 1])
    head.errors[0] = 9.0
    assert ov.errors[0] == 0.1
    # two values per observation, and a ragged object array
    ov2 = ObsVector(values=np.ones((5, 2)))
    assert ov2.num_obs == 5 and np.array_equal(ov2.obs_dims, [2] * 5)
    assert ov2.times is None              # the base constructor runs before the values are set (_obs_vector.py:71-75)
    ragged = np.empty(2, dtype=object)
    ragged[0], ragged[1] = np.ones(3), np.ones(1)
    assert np.array_equal(ObsVector(values=ragged).obs_dims, [3, 1])
    with pytest.warns(UserWarning, match='does not match dimensions'):
        ObsVector(values=np.ones(3), obs_dims=[2, 2, 2])
    empty = ObsVector(obs_dims=[1, 1])
    assert empty.values is None and empty.obs_dims == [1, 1] and empty.errors is None and empty.coords is None
    with pytest.raises(AttributeError, match='does not contain any data values'):
        empty[:1]


def test_setters_convert_to_arrays():
    sv = make_state()
    sv.values = [[1.0, 2.0, 3.0, 4.0]]
    sv.times = [7]
    assert isinstance(sv.values, np.ndarray) and isinstance(sv.times, np.ndarray)
    ov = ObsVector(values=[1.0, 2.0])
    ov.errors = [0.5, 0.5]
    ov.coords = [[0], [1]]
    assert isinstance(ov.errors, np.ndarray) and ov.coords.shape == (2, 1)
    ov.values = [1.0, 2.0, 3.0]
    assert ov.num_obs == 3
