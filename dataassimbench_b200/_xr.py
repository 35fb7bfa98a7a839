"""Dataset container used at the API boundary.

The reference passes `xarray.Dataset` objects in and out of `ETKF.cycle()`.  xarray is not
installed in this image, so when it cannot be imported a small stand-in with the same spelling
for the handful of operations the hot path and its tests touch is used instead
(SURVEY.md 7.1 step 0).  When xarray IS importable, real `xarray.Dataset` objects are accepted
and returned.  Nothing here computes on the hot path: it only names arrays.
"""
import numpy as np

try:                                    # pragma: no cover - not available in this image
    import xarray as _xarray
    HAVE_XARRAY = True
except Exception:                        # ImportError or a broken install
    _xarray = None
    HAVE_XARRAY = False


class MiniDataArray:
    def __init__(self, dims, data, coords=None, attrs=None, name=None):
        self.dims = tuple(dims)
        self.data = data if hasattr(data, "shape") else np.asarray(data)
        if len(self.dims) != self.data.ndim:
            raise ValueError("dims %r do not match data of shape %r" % (self.dims, self.data.shape))
        self.coords = dict(coords or {})
        self.attrs = dict(attrs or {})
        self.name = name

    @property
    def values(self):
        return np.asarray(self.data)

    @property
    def shape(self):
        return tuple(self.data.shape)

    @property
    def sizes(self):
        return dict(zip(self.dims, self.data.shape))

    @property
    def size(self):
        return int(np.prod(self.data.shape))

    def __array__(self, dtype=None, copy=None):
        a = np.asarray(self.data)
        return a.astype(dtype) if dtype is not None else a

    def __getitem__(self, key):
        return np.asarray(self.data)[key]

    def __len__(self):
        return self.data.shape[0]

    def isel(self, **idx):
        return _isel_array(self, idx)

    def mean(self, dim=None):
        if dim is None:
            return float(np.mean(self.data))
        ax = self.dims.index(dim)
        return MiniDataArray([d for d in self.dims if d != dim], np.mean(self.data, axis=ax),
                             {k: v for k, v in self.coords.items() if k != dim}, self.attrs, self.name)

    def transpose(self, *dims):
        dims = _expand_ellipsis(dims, self.dims)
        return MiniDataArray(dims, np.transpose(self.data, [self.dims.index(d) for d in dims]),
                             self.coords, self.attrs, self.name)


def _expand_ellipsis(dims, all_dims):
    if Ellipsis in dims:
        i = dims.index(Ellipsis)
        rest = [d for d in all_dims if d not in dims]
        dims = tuple(dims[:i]) + tuple(rest) + tuple(dims[i + 1:])
    return tuple(dims)


def _isel_array(arr, idx):
    data, dims, coords = arr.data, list(arr.dims), dict(arr.coords)
    for d, i in idx.items():
        if d not in dims:
            continue
        ax = dims.index(d)
        data = np.take(data, i, axis=ax) if not isinstance(i, slice) else data[(slice(None),) * ax + (i,)]
        if d in coords:
            c = np.asarray(coords[d])
            coords[d] = c[i] if c.ndim else c
        if np.ndim(i) == 0 and not isinstance(i, slice):
            dims.pop(ax)
    return MiniDataArray(dims, data, coords, arr.attrs, arr.name)


class MiniDataset:
    """`xarray.Dataset` look-alike: data_vars {name: (dims, array)}, coords, attrs."""

    def __init__(self, data_vars=None, coords=None, attrs=None):
        self._vars = {}
        self.coords = {k: (v if isinstance(v, np.ndarray) else np.asarray(v))
                       for k, v in (coords or {}).items()}
        self.attrs = dict(attrs or {})
        for name, spec in (data_vars or {}).items():
            self._set(name, spec)

    def _set(self, name, spec):
        if isinstance(spec, MiniDataArray):
            self._vars[name] = MiniDataArray(spec.dims, spec.data, None, spec.attrs, name)
        else:
            dims, data = spec[0], spec[1]
            self._vars[name] = MiniDataArray(tuple(dims), data, None, None, name)

    # ---- mapping-ish surface
    @property
    def data_vars(self):
        return dict(self._vars)

    def __iter__(self):
        return iter(self._vars)

    def __contains__(self, k):
        return k in self._vars or k in self.coords

    def __getitem__(self, key):
        if key in self._vars:
            v = self._vars[key]
            return MiniDataArray(v.dims, v.data, {d: self.coords[d] for d in v.dims if d in self.coords},
                                 v.attrs, key)
        if key in self.coords:
            c = self.coords[key]
            return MiniDataArray((key,) if np.ndim(c) else (), c, None, None, key)
        raise KeyError(key)

    def __setitem__(self, key, spec):
        self._set(key, spec)

    def __getattr__(self, name):
        # xarray exposes attrs / coords / variables as attributes (obs_vector.error_sd, ds.time, ...)
        if name.startswith("_"):
            raise AttributeError(name)
        attrs = self.__dict__.get("attrs", {})
        if name in attrs:
            return attrs[name]
        if name in self.__dict__.get("_vars", {}) or name in self.__dict__.get("coords", {}):
            return self[name]
        raise AttributeError(name)

    @property
    def sizes(self):
        out = {}
        for v in self._vars.values():
            out.update(v.sizes)
        for k, c in self.coords.items():
            if np.ndim(c) == 1 and k not in out:
                out[k] = len(c)
        return out

    @property
    def dims(self):
        return self.sizes

    def copy(self):
        return MiniDataset({k: (v.dims, v.data) for k, v in self._vars.items()}, self.coords, self.attrs)

    def assign(self, **kw):
        out = self.copy()
        for k, spec in kw.items():
            if isinstance(spec, (tuple, list)) and len(spec) == 2 and not np.isscalar(spec[0]) \
                    and all(isinstance(d, str) for d in spec[0]):
                out._set(k, spec)
            elif isinstance(spec, MiniDataArray):
                out._set(k, spec)
            else:
                out._set(k, ((), np.asarray(spec)))
        return out

    def assign_attrs(self, **kw):
        out = self.copy()
        out.attrs.update(kw)
        return out

    def assign_coords(self, coords=None, **kw):
        out = self.copy()
        for k, v in dict(coords or {}, **kw).items():
            out.coords[k] = np.asarray(v)
        return out

    def drop_vars(self, names):
        names = [names] if isinstance(names, str) else list(names)
        out = self.copy()
        for n in names:
            out._vars.pop(n, None)
            out.coords.pop(n, None)
        return out

    def isel(self, **idx):
        out = MiniDataset(attrs=self.attrs)
        for k, v in self._vars.items():
            a = _isel_array(v, idx)
            out._vars[k] = MiniDataArray(a.dims, a.data, None, v.attrs, k)
        for k, c in self.coords.items():
            if k in idx and np.ndim(c) == 1:
                i = idx[k]
                out.coords[k] = c[i]
            else:
                out.coords[k] = c
        return out

    def mean(self, dim=None):
        out = MiniDataset(attrs=self.attrs, coords={k: c for k, c in self.coords.items() if k != dim})
        for k, v in self._vars.items():
            m = v.mean(dim) if (dim is None or dim in v.dims) else v
            out._vars[k] = m if isinstance(m, MiniDataArray) else MiniDataArray((), np.asarray(m))
        return out

    def transpose(self, *dims):
        out = self.copy()
        for k, v in self._vars.items():
            out._vars[k] = v.transpose(*[d for d in _expand_ellipsis(dims, v.dims) if d in v.dims])
        return out

    def stack(self, **kw):
        """stack(new=[a, b]): merge dims a, b (in that order) into a trailing dim `new`."""
        out = self.copy()
        for new, olds in kw.items():
            for k, v in list(out._vars.items()):
                if not all(o in v.dims for o in olds):
                    continue
                keep = [d for d in v.dims if d not in olds]
                t = v.transpose(*(keep + list(olds)))
                shape = [t.data.shape[i] for i in range(len(keep))] + [-1]
                out._vars[k] = MiniDataArray(keep + [new], np.reshape(np.asarray(t.data), shape), None, v.attrs, k)
        return out

    def fillna(self, value):
        out = self.copy()
        for k, v in out._vars.items():
            a = np.asarray(v.data)
            if np.issubdtype(a.dtype, np.floating):
                out._vars[k] = MiniDataArray(v.dims, np.where(np.isnan(a), value, a), None, v.attrs, k)
        return out

    @property
    def dab(self):
        """The reference's `ds.dab` accessor (dabench/data/_xarray_accessor.py:23-51)."""
        return _DabAccessor(self)

    def __repr__(self):
        body = ", ".join("%s%s" % (k, tuple(v.dims)) for k, v in self._vars.items())
        return "<MiniDataset %s | coords: %s | attrs: %s>" % (body, list(self.coords), list(self.attrs))


class _DabAccessor:
    """`Dataset.dab`: `flatten()` stacks every non-time dimension of every variable into one trailing `system`
    dimension (`to_stacked_array('system', ['time'])` in the reference), `split_train_val_test()` cuts the time
    axis into consecutive pieces given as counts or as fractions."""

    def __init__(self, ds):
        self._ds = ds

    def flatten(self):
        ds = self._ds
        has_time = 'time' in ds.coords
        cols = []
        for v in ds._vars.values():
            a = np.asarray(v.data)
            if has_time and 'time' in v.dims:
                a = np.moveaxis(a, v.dims.index('time'), 0)
                cols.append(a.reshape(a.shape[0], -1))
            else:
                cols.append(a.reshape(1, -1) if has_time else a.reshape(-1))
        if has_time:
            n_t = max(c.shape[0] for c in cols)
            cols = [np.broadcast_to(c, (n_t, c.shape[1])) for c in cols]
            return MiniDataArray(('time', 'system'), np.concatenate(cols, axis=1), {'time': ds.coords['time']}, ds.attrs)
        return MiniDataArray(('system',), np.concatenate(cols), None, ds.attrs)

    def split_train_val_test(self, split_lengths):
        ds = self._ds
        n_t = ds.sizes['time']
        lengths = np.array(split_lengths)
        if (lengths > 1.0).sum() == 0:                       # fractions of the time axis
            lengths = np.round(lengths * n_t).astype(int)
        total = int(np.sum(lengths))
        if total != n_t:
            import warnings
            warnings.warn("Specified split lengths ({}) {} Xarray object's time dimension ({}).".format(
                split_lengths, 'exceed' if total > n_t else 'are shorter than', n_t))
        out, start = [], 0
        for n in lengths:
            # the reference never advances its start index (:47-50): every piece begins at 0; kept
            out.append(ds.isel(time=slice(start, start + int(n))))
        return tuple(out)


if HAVE_XARRAY:                          # pragma: no cover - not available in this image
    try:
        @_xarray.register_dataset_accessor('dab')
        class _XarrayDabAccessor:
            def __init__(self, obj):
                self._obj = obj

            def flatten(self):
                return self._obj.to_stacked_array('system', ['time'] if 'time' in self._obj.coords else [])

            def split_train_val_test(self, split_lengths):
                lengths = np.array(split_lengths)
                if (lengths > 1.0).sum() == 0:
                    lengths = np.round(lengths * self._obj.sizes['time']).astype(int)
                return tuple(self._obj.isel(time=slice(0, int(n))) for n in lengths)
    except Exception:
        pass


def new_dataset(data_vars, coords=None, attrs=None):
    """Build a Dataset: real xarray when importable, the stand-in otherwise."""
    if HAVE_XARRAY:
        return _xarray.Dataset(data_vars, coords=coords, attrs=attrs)
    return MiniDataset(data_vars, coords, attrs)


def var_array(ds, name):
    """NumPy view of a data variable of either Dataset flavour."""
    v = ds[name]
    return np.asarray(v.values if hasattr(v, "values") else v)


def var_data(ds, name):
    """The array object behind a data variable as it was stored (a NumPy array, or a device tensor
    kept by the stand-in Dataset); no conversion."""
    v = ds[name]
    return v.data if hasattr(v, "data") else v


def var_dims(ds, name):
    return tuple(ds[name].dims)


def get_attr(ds, name, default=None):
    return ds.attrs.get(name, default)
