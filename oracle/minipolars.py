"""TEST INFRASTRUCTURE ONLY -- the few polars calls the reference's scoring path makes, on numpy columns.

polars is not installed in this image, so the reference functions that take a polars frame could not run here and
their glue had to be restated (oracle/restate.py).  This module implements, with the obvious semantics, exactly the
calls those functions make, so that the UNMODIFIED reference code runs on a `DataFrame` of numpy columns:

    motif_model_contig / motif_model_bin / get_parent_scores   find_motifs_bin.py:1265-1331, 1382-1433
        frame.filter(pl.col(c) <cmp> v), frame[c].to_numpy()
    find_best_candidates / MotifSearcher                       find_motifs_bin.py:607-834, 853-1182
        + get_column(c).unique() / .to_list(), is_empty()
    filter_pileup / filter_pileup_minimummod_frequency         dataload.py:191-226
        + with_columns([(col + "_" + col).alias(n)]), group_by(c).agg(name=expr), pl.count(), (expr > v).sum(),
          expr / expr, is_in, drop
    merge_motifs_in_df                                         find_motifs_bin.py:1436-1537
        + `for key, df in frame.group_by(a, b)`, is_in(...).not_(), DataFrame({column: scalar}), concat
          (with `pl.DataFrame` bound to this module's DataFrame and the MotifSearchResult wrapper bypassed by the test)

`install(nm)` puts `col`, `count`, `lit` on the stub `polars` module that oracle/ref_shim.py registered (the reference
modules hold that module as `pl`) and on the names imported with `from polars import col`.  Everything else of polars
stays a placeholder.  Semantics that polars leaves open are fixed here and stated: `unique()` keeps first-appearance
order (polars' default order is unspecified), `group_by` groups appear in first-appearance order.

Nothing in nanomotif_b200/ may import this module.
"""
from __future__ import annotations

import numpy as np


def _arr(v):
    if isinstance(v, np.ndarray):
        a = v
    elif isinstance(v, (list, tuple)):
        if v and not isinstance(v[0], (int, float, bool, str, np.generic)):
            a = np.empty(len(v), dtype=object)  # arbitrary objects (e.g. models): no array coercion
            a[:] = list(v)
        else:
            a = np.asarray(v)
    else:  # a scalar: a one-row column (pl.DataFrame({"motif": "GATC", "model": model, ...}))
        a = np.empty(1, dtype=object)
        a[0] = v
        if isinstance(v, (int, float, bool, np.generic)) and not isinstance(v, str):
            a = np.asarray([v])
    if a.dtype.kind in "US":
        a = a.astype(object)
    return a


def _plain(v):
    """str subclasses (the reference's Motif) compare by their text inside a frame, as in polars."""
    return str.__str__(v) if isinstance(v, str) else v


class Series:
    def __init__(self, values, name=""):
        self.values, self.name = _arr(values), name

    def to_numpy(self):
        return self.values

    def to_list(self):
        return self.values.tolist()

    def unique(self):
        _, first = np.unique(self.values.astype(str) if self.values.dtype == object else self.values, return_index=True)
        return Series(self.values[np.sort(first)], self.name)

    def __len__(self):
        return len(self.values)

    def __iter__(self):
        return iter(self.values.tolist())

    def __getitem__(self, i):
        v = self.values[i]
        return Series(v, self.name) if isinstance(v, np.ndarray) else (v.item() if hasattr(v, "item") else v)


class Expr:
    """A column expression: `fn(frame)` -> numpy array (or a scalar for aggregations)."""

    def __init__(self, fn, name=None):
        self.fn, self.name = fn, name

    def _bin(self, other, op):
        o = other.fn if isinstance(other, Expr) else (lambda df, _v=other: _v)
        return Expr(lambda df: op(self.fn(df), o(df)), self.name)

    def __eq__(self, other):  # noqa: D105 - expression builder, like polars
        return self._bin(other, lambda a, b: a == b)

    def __ne__(self, other):
        return self._bin(other, lambda a, b: a != b)

    def __ge__(self, other):
        return self._bin(other, lambda a, b: a >= b)

    def __le__(self, other):
        return self._bin(other, lambda a, b: a <= b)

    def __gt__(self, other):
        return self._bin(other, lambda a, b: a > b)

    def __lt__(self, other):
        return self._bin(other, lambda a, b: a < b)

    def __and__(self, other):
        return self._bin(other, lambda a, b: a & b)

    def __or__(self, other):
        return self._bin(other, lambda a, b: a | b)

    def __add__(self, other):
        return self._bin(other, lambda a, b: a + b)

    def __truediv__(self, other):
        return self._bin(other, lambda a, b: np.asarray(a, dtype=np.float64) / np.asarray(b, dtype=np.float64))

    def __invert__(self):
        return Expr(lambda df: ~self.fn(df), self.name)

    not_ = __invert__
    __hash__ = None

    def is_in(self, values):
        vals = {_plain(v) for v in (values.to_list() if isinstance(values, Series) else values)}
        return Expr(lambda df: np.fromiter((_plain(x) in vals for x in np.asarray(self.fn(df)).tolist()), dtype=bool,
                                           count=len(self.fn(df))), self.name)

    def sum(self):
        return Expr(lambda df: np.asarray(self.fn(df)).sum(), self.name)

    def alias(self, name):
        return Expr(self.fn, name)


def col(name: str) -> Expr:
    return Expr(lambda df: df._cols[name], name)


def lit(value) -> Expr:
    return Expr(lambda df: value)


def count() -> Expr:
    return Expr(lambda df: df.height, "count")


class GroupBy:
    def __init__(self, frame, keys):
        self.frame, self.keys = frame, keys

    def _groups(self):
        f = self.frame
        key_cols = [f._cols[k] for k in self.keys]
        tags = np.array(["\x1f".join(str(c[i]) for c in key_cols) for i in range(f.height)], dtype=object)
        _, first, inv = np.unique(tags.astype(str), return_index=True, return_inverse=True)
        for g in np.argsort(first):  # groups in first-appearance order
            yield np.flatnonzero(inv == g)

    def __iter__(self):
        """((key values), sub-frame) per group, as `for (a, b), df in frame.group_by("a", "b")`."""
        f = self.frame
        for rows in self._groups():
            key = tuple(_plain(f._cols[k][rows[0]]) if not hasattr(f._cols[k][rows[0]], "item") else f._cols[k][rows[0]].item()
                        for k in self.keys)
            yield key, DataFrame({k: v[rows] for k, v in f._cols.items()})

    def agg(self, *exprs, **named):
        f = self.frame
        out = {k: [] for k in self.keys}
        specs = [(e.name, e) for e in exprs] + list(named.items())
        for name, _ in specs:
            out[name] = []
        for rows in self._groups():
            sub = DataFrame({k: v[rows] for k, v in f._cols.items()})
            for k in self.keys:
                out[k].append(f._cols[k][rows[0]])
            for name, e in specs:
                out[name].append(e.fn(sub))
        return DataFrame({k: _arr(v) if len(v) else np.zeros(0) for k, v in out.items()})


class DataFrame:
    """Columns as numpy arrays (strings as object arrays)."""

    def __init__(self, data=None):
        self._cols = {k: _arr(v) for k, v in (data or {}).items()}
        n = {len(v) for v in self._cols.values()}
        if len(n) > 1:
            raise ValueError("columns differ in length")

    @property
    def height(self) -> int:
        return len(next(iter(self._cols.values()))) if self._cols else 0

    @property
    def columns(self) -> list:
        return list(self._cols)

    @property
    def shape(self):
        return (self.height, len(self._cols))

    def is_empty(self) -> bool:
        return self.height == 0

    def __len__(self):
        return self.height

    def filter(self, *exprs):
        keep = np.ones(self.height, dtype=bool)
        for e in exprs:
            keep &= np.asarray(e.fn(self), dtype=bool)
        return DataFrame({k: v[keep] for k, v in self._cols.items()})

    def get_column(self, name) -> Series:
        return Series(self._cols[name], name)

    def __getitem__(self, name) -> Series:
        return self.get_column(name)

    def with_columns(self, *exprs, **named):
        cols = dict(self._cols)
        flat = [e for x in exprs for e in (x if isinstance(x, (list, tuple)) else [x])]
        for e in flat:
            cols[e.name] = _arr(e.fn(self))
        for name, e in named.items():
            cols[name] = _arr(e.fn(self))
        return DataFrame(cols)

    def drop(self, *names):
        flat = [n for x in names for n in (x if isinstance(x, (list, tuple)) else [x])]
        return DataFrame({k: v for k, v in self._cols.items() if k not in flat})

    def group_by(self, *keys):
        return GroupBy(self, [k for x in keys for k in (x if isinstance(x, (list, tuple)) else [x])])

    def sort(self, name):
        order = np.argsort(self._cols[name], kind="stable")
        return DataFrame({k: v[order] for k, v in self._cols.items()})

    def to_dict(self) -> dict:
        return dict(self._cols)


def concat(frames, **_kwargs):
    """Rows of frames with the same columns, in order (pl.concat with its default vertical strategy)."""
    frames = list(frames)
    cols = frames[0].columns
    for f in frames:
        if set(f.columns) != set(cols):
            raise ValueError(f"concat: columns differ: {f.columns} vs {cols}")
    out = {}
    for c in cols:
        parts = [f._cols[c] for f in frames]
        if any(p.dtype == object for p in parts):
            parts = [p.astype(object) for p in parts]
        out[c] = np.concatenate(parts) if parts else np.zeros(0)
    return DataFrame(out)


def install(nm) -> None:
    """Make `pl.col` / `pl.count` / `pl.lit` and the `col` imported by name resolve to this module inside the loaded
    reference package `nm` (oracle/ref_shim.load_reference())."""
    import sys

    pl = sys.modules["polars"]
    for name, obj in (("col", col), ("count", count), ("lit", lit), ("concat", concat)):
        setattr(pl, name, obj)
    for mod in (nm.find_motifs_bin, nm.dataload):
        if hasattr(mod, "col") or "col" in vars(mod):
            mod.col = col
