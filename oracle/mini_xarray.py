"""TEST INFRASTRUCTURE ONLY -- a minimal labelled array: the part of ``xarray.DataArray`` that the reference's
post-processing methods touch (``ElectromagneticFieldData.flux / dot / outer_dot / pol_fraction / symmetry_expanded``,
``ModeData.overlap_sort``, ``ModeSolver._colocate_data / _normalize_modes / _grid_correction``).

xarray is not in this image, so ``oracle/ref_post.py`` executes the reference's own method bodies over THIS
class instead.  It is not a port of xarray: it implements, from xarray's documented behaviour, exactly the operations
those method bodies use --

  * arithmetic and numpy ufuncs broadcast BY DIMENSION NAME, coordinates aligned by exact label ("inner" join, xarray's
    default for arithmetic); in-place operators write into the existing buffer (``field /= scaling`` is seen by the owner);
  * ``interp``: linear, one axis after the other, through ``scipy.interpolate.interp1d(bounds_error=False)`` -- the very call
    xarray makes for 1-D linear interpolation -- NaN outside the data unless ``kwargs`` carries a ``fill_value``;
  * ``sum`` skips NaN (xarray's ``skipna`` default for floating-point data);
  * ``sel(method="nearest")``, ``isel`` (integer -> the dimension is dropped, list / array -> kept), outer indexing with a
    dict key (``a[{dim: inds}] *= v``), ``squeeze``, ``assign_coords``, ``rename``, ``expand_dims``, ``get_axis_num``.

Anything else raises.  Nothing under ``tidy3d_b200/`` may import this module.
"""
from __future__ import annotations

import numpy as np
from scipy.interpolate import interp1d


class _Coords(dict):
    """``da.coords``: dimension name -> 1-D DataArray of labels."""

    def to_index(self):  # pragma: no cover - not used by the reference paths
        raise NotImplementedError


def _as_label_array(v):
    if isinstance(v, DataArray):
        v = v.values
    return np.asarray(v)


class DataArray:
    __array_priority__ = 100.0

    def __init__(self, data=None, coords=None, dims=None, name=None, attrs=None):
        if isinstance(data, DataArray):
            coords = coords if coords is not None else {k: v.values for k, v in data.coords.items()}
            dims = dims if dims is not None else data.dims
            data = data.values
        self.values = np.asarray(data)
        if dims is None:
            if coords is None:
                if self.values.ndim:
                    raise ValueError("mini_xarray: dims or coords required")
                dims = ()
            else:
                dims = tuple(coords.keys())
        if isinstance(dims, str):
            dims = (dims,)
        self.dims = tuple(dims)
        if len(self.dims) != self.values.ndim:
            raise ValueError(f"mini_xarray: {self.values.ndim}-d data with dims {self.dims}")
        self._labels = {}
        for k, v in (coords or {}).items():
            lab = _as_label_array(v)
            if lab.ndim == 0:
                continue  # scalar (non-index) coordinates are dropped
            if k not in self.dims:
                raise ValueError(f"mini_xarray: coordinate {k!r} is not a dimension of {self.dims}")
            if lab.shape != (self.values.shape[self.dims.index(k)],):
                raise ValueError(f"mini_xarray: coordinate {k!r} of length {lab.shape} on axis of {self.values.shape[self.dims.index(k)]}")
            self._labels[k] = lab
        self.name = name
        self.attrs = dict(attrs or {})

    # ---- basic properties --------------------------------------------------------------------------------------------
    @property
    def data(self):
        return self.values

    @data.setter
    def data(self, v):
        v = np.asarray(v)
        if v.shape != self.values.shape:
            raise ValueError("mini_xarray: replacement data must keep the shape")
        self.values = v

    @property
    def coords(self):
        return _Coords((k, type(self)(self._labels[k], dims=(k,))) for k in self.dims if k in self._labels)

    @property
    def sizes(self):
        return dict(zip(self.dims, self.values.shape))

    shape = property(lambda self: self.values.shape)
    dtype = property(lambda self: self.values.dtype)
    ndim = property(lambda self: self.values.ndim)
    size = property(lambda self: self.values.size)
    real = property(lambda self: self._like(self.values.real))
    imag = property(lambda self: self._like(self.values.imag))
    T = property(lambda self: type(self)(self.values.T, self._labels, self.dims[::-1]))

    def __len__(self):
        return len(self.values)

    def __iter__(self):
        return iter(self.values)

    def __getattr__(self, name):  # da.f, da.mode_index, da.x ...
        labels = self.__dict__.get("_labels", {})
        if name in labels:
            return type(self)(labels[name], coords={name: labels[name]}, dims=(name,))
        raise AttributeError(f"mini_xarray.DataArray has no attribute {name!r}")

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.values, dtype=dtype)

    def __repr__(self):
        return f"<mini DataArray {dict(self.sizes)} {self.dtype}>"

    def __float__(self):
        return float(self.values)

    def __complex__(self):
        return complex(self.values)

    def __bool__(self):
        return bool(self.values)

    def _like(self, values):
        return type(self)(values, self._labels, self.dims)

    def copy(self, deep=True, data=None):
        vals = (np.array(self.values, copy=True) if deep else self.values) if data is None else np.asarray(data)
        return type(self)(vals, {k: np.array(v, copy=True) for k, v in self._labels.items()}, self.dims)

    def to_numpy(self):
        return self.values

    def astype(self, dtype, **_):
        return self._like(self.values.astype(dtype))

    def conj(self):
        return self._like(np.conj(self.values))

    conjugate = conj

    def item(self):
        return self.values.item()

    def get_axis_num(self, dim):
        return self.dims.index(dim)

    # ---- arithmetic: broadcasting by dimension name, inner join on labels ---------------------------------------------
    @staticmethod
    def _align(a, b):
        """Restrict both operands to the labels they share along common labelled dimensions."""
        for d in a.dims:
            if d in b.dims and d in a._labels and d in b._labels:
                la, lb = a._labels[d], b._labels[d]
                if la.shape == lb.shape and np.array_equal(la, lb):
                    continue
                common = [v for v in la if v in set(lb.tolist())]
                ia = [int(np.where(la == v)[0][0]) for v in common]
                ib = [int(np.where(lb == v)[0][0]) for v in common]
                a, b = a.isel(**{d: ia}), b.isel(**{d: ib})
        return a, b

    def _binary(self, other, op, reflexive=False):
        if isinstance(other, DataArray):
            a, b = self._align(self, other)
            dims = list(a.dims) + [d for d in b.dims if d not in a.dims]
            for d in dims:
                if d in a.dims and d in b.dims and a.sizes[d] != b.sizes[d]:
                    raise ValueError(f"mini_xarray: dimension {d!r} of sizes {a.sizes[d]} and {b.sizes[d]} without labels to align")

            def spread(x):
                order = [x.dims.index(d) for d in dims if d in x.dims]
                v = np.transpose(x.values, order)
                shape = [x.sizes[d] if d in x.dims else 1 for d in dims]
                return v.reshape(shape)

            av, bv = spread(a), spread(b)
            labels = dict(b._labels)
            labels.update(a._labels)
            res = op(bv, av) if reflexive else op(av, bv)
            return type(self)(res, {k: v for k, v in labels.items() if k in dims}, dims)
        o = np.asarray(other)
        if o.ndim > self.values.ndim:
            raise ValueError("mini_xarray: cannot broadcast a larger unlabelled array")
        res = op(o, self.values) if reflexive else op(self.values, o)
        if res.shape != self.values.shape:
            raise ValueError("mini_xarray: unlabelled operand changed the shape")
        return self._like(res)

    def _inplace(self, other, op):
        res = self._binary(other, op)
        if res.dims != self.dims or res.shape != self.shape:
            raise ValueError("mini_xarray: in-place operation would change dimensions")
        if np.can_cast(res.dtype, self.values.dtype, casting="same_kind"):
            self.values[...] = res.values  # the owner of this array sees the update
        else:
            self.values = res.values
        return self

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or kwargs.get("out") is not None:
            return NotImplemented
        kwargs.pop("out", None)
        if len(inputs) == 1:
            return self._like(ufunc(self.values, **kwargs))
        if len(inputs) == 2:
            a, b = inputs
            if a is self:
                return self._binary(b, lambda x, y: ufunc(x, y, **kwargs))
            return self._binary(a, lambda x, y: ufunc(x, y, **kwargs), reflexive=True)
        return NotImplemented

    def __neg__(self):
        return self._like(-self.values)

    def __abs__(self):
        return self._like(np.abs(self.values))

    def __pos__(self):
        return self


def _install_operators():
    import operator as O

    for nm, fn in (("add", O.add), ("sub", O.sub), ("mul", O.mul), ("truediv", O.truediv), ("pow", O.pow),
                   ("lt", O.lt), ("le", O.le), ("gt", O.gt), ("ge", O.ge), ("eq", O.eq), ("ne", O.ne)):
        setattr(DataArray, f"__{nm}__", (lambda f: lambda self, other: self._binary(other, f))(fn))
    for nm, fn in (("add", O.add), ("sub", O.sub), ("mul", O.mul), ("truediv", O.truediv), ("pow", O.pow)):
        setattr(DataArray, f"__r{nm}__", (lambda f: lambda self, other: self._binary(other, f, reflexive=True))(fn))
        setattr(DataArray, f"__i{nm}__", (lambda f: lambda self, other: self._inplace(other, f))(fn))
    DataArray.__hash__ = None


_install_operators()


def _indexer(self, key):
    """(tuple of per-axis indices for np.ix_-style OUTER indexing, kept dims, kept labels)."""
    idx, dims, labels = [], [], {}
    for ax, d in enumerate(self.dims):
        if d in key:
            k = key[d]
            if isinstance(k, DataArray):
                k = k.values
            if isinstance(k, slice):
                k = np.arange(self.values.shape[ax])[k]
            k = np.asarray(k)
            if k.dtype == bool:
                k = np.nonzero(k)[0]
            if k.ndim == 0:
                idx.append(int(k))
                continue
            idx.append(k.astype(int))
        else:
            idx.append(np.arange(self.values.shape[ax]))
        dims.append(d)
        if d in self._labels:
            labels[d] = self._labels[d][idx[-1]]
    unknown = set(key) - set(self.dims)
    if unknown:
        raise KeyError(f"mini_xarray: {unknown} are not dimensions of {self.dims}")
    return idx, dims, labels


def _outer(idx):
    """Broadcastable index tuple for outer (orthogonal) indexing, integers dropping their axis."""
    arrays = [i for i in idx if not isinstance(i, int)]
    grids = iter(np.ix_(*arrays)) if arrays else iter(())
    return tuple(i if isinstance(i, int) else next(grids) for i in idx)


def isel(self, indexers=None, drop=False, **kw):
    key = dict(indexers or {}, **kw)
    idx, dims, labels = _indexer(self, key)
    return type(self)(self.values[_outer(idx)], labels, dims)


def getitem(self, key):
    if isinstance(key, dict):
        return self.isel(**key)
    if isinstance(key, str):
        return self.coords[key]
    raise TypeError("mini_xarray: only dict keys are supported for indexing")


def setitem(self, key, value):
    if not isinstance(key, dict):
        raise TypeError("mini_xarray: only dict keys are supported for assignment")
    idx, dims, _ = _indexer(self, key)
    if isinstance(value, DataArray):
        if tuple(value.dims) != tuple(dims):
            raise ValueError("mini_xarray: assigned array has other dimensions")
        value = value.values
    self.values[_outer(idx)] = value


def sel(self, indexers=None, method=None, drop=False, **kw):
    key = dict(indexers or {}, **kw)
    pos = {}
    for d, want in key.items():
        lab = self._labels[d]
        w = _as_label_array(want)
        scalar = w.ndim == 0
        w = np.atleast_1d(w)
        if method == "nearest":
            dist = np.abs(lab[None, :].astype(float) - w[:, None].astype(float))
            # pandas: "tied distances are broken by preferring the larger index value"
            ind = dist.shape[1] - 1 - np.argmin(dist[:, ::-1], axis=1)
        elif method is None:
            ind = np.array([int(np.where(lab == v)[0][0]) for v in w])
        else:
            raise NotImplementedError(method)
        pos[d] = int(ind[0]) if scalar else ind
    return self.isel(**pos)


def assign_coords(self, coords=None, **kw):
    new = dict(self._labels)
    for k, v in dict(coords or {}, **kw).items():
        new[k] = _as_label_array(v)
    return type(self)(self.values, new, self.dims)


def squeeze(self, dim=None, drop=False, axis=None):
    if dim is None:
        dims = [d for d in self.dims if self.sizes[d] == 1]
    else:
        dims = [dim] if isinstance(dim, str) else list(dim)
    for d in dims:
        if self.sizes[d] != 1:
            raise ValueError(f"mini_xarray: cannot squeeze dimension {d!r} of size {self.sizes[d]}")
    return self.isel(**{d: 0 for d in dims})


def rename(self, new_name_or_name_dict=None, **names):
    m = dict(new_name_or_name_dict or {}, **names)
    return type(self)(self.values, {m.get(k, k): v for k, v in self._labels.items()}, tuple(m.get(d, d) for d in self.dims))


def expand_dims(self, dim=None, axis=None):
    if not isinstance(dim, dict) or len(dim) != 1:
        raise NotImplementedError("mini_xarray: expand_dims(dim={name: labels}, axis=k)")
    (name, lab), = dim.items()
    lab = _as_label_array(lab)
    axis = 0 if axis is None else axis
    vals = np.repeat(np.expand_dims(self.values, axis), lab.size, axis=axis)
    dims = list(self.dims)
    dims.insert(axis, name)
    return type(self)(vals, dict(self._labels, **{name: lab}), dims)


def _sum(self, dim=None, skipna=None, **_):
    dims = list(self.dims) if dim is None else [dim] if isinstance(dim, str) else list(dim)
    axes = tuple(self.dims.index(d) for d in dims)
    fn = np.nansum if (self.values.dtype.kind in "fc" and skipna is not False) else np.sum
    keep = [d for d in self.dims if d not in dims]
    return type(self)(fn(self.values, axis=axes), {k: v for k, v in self._labels.items() if k in keep}, keep)


def interp(self, coords=None, method="linear", assume_sorted=False, kwargs=None, **coords_kwargs):
    if method != "linear":
        raise NotImplementedError(method)
    want = dict(coords or {}, **coords_kwargs)
    out = self
    for d, new in want.items():  # orthogonal indexers: one 1-D interpolation per dimension (xarray decomposes the same way)
        ax = out.dims.index(d)
        x = out._labels[d].astype(float)
        new = _as_label_array(new).astype(float)
        y = out.values
        if not assume_sorted:
            order = np.argsort(x)
            x, y = x[order], np.take(y, order, axis=ax)
        opts = dict(bounds_error=False)
        opts.update(kwargs or {})
        vals = interp1d(x, y, kind="linear", axis=ax, assume_sorted=True, **opts)(new)
        labels = dict(out._labels)
        dims = list(out.dims)
        if new.ndim == 0:
            labels.pop(d)
            dims.pop(ax)
        else:
            labels[d] = new
        out = type(self)(vals, labels, dims)
    return out


for _nm, _fn in (("isel", isel), ("__getitem__", getitem), ("__setitem__", setitem), ("sel", sel), ("assign_coords", assign_coords),
                 ("squeeze", squeeze), ("rename", rename), ("expand_dims", expand_dims), ("sum", _sum), ("interp", interp)):
    setattr(DataArray, _nm, _fn)


class Dataset:
    """``xr.Dataset(data_vars={...})``: attribute / key access to the named arrays, nothing else."""

    def __init__(self, data_vars=None, coords=None, attrs=None):
        self.data_vars = dict(data_vars or {})

    def __getattr__(self, name):
        dv = self.__dict__.get("data_vars", {})
        if name in dv:
            return dv[name]
        raise AttributeError(name)

    def __getitem__(self, name):
        return self.data_vars[name]
