"""A small R-formula front end: just enough of `model.frame` / `model.matrix` for the
designs mincLm, vertexLm and anatLm build (R/minc_voxel_statistics.R:773-812).

Supported on the right-hand side: `+`, `-1` / `+0` (no intercept), `:` and `*` between
columns, numeric columns, and factors (string / categorical / bool columns) coded with R's
default treatment contrasts (first level in sorted order is the baseline;
`drop.unused.levels = TRUE` as mincLm sets, :777).  Column names follow R:
"(Intercept)", "SexM", "age", "SexM:age".

This is host logic only -- the design matrix it returns goes through the C ABI.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np


class FormulaError(ValueError):
    pass


def split_formula(formula: str) -> Tuple[str, str]:
    if "~" not in formula:
        raise FormulaError("formula needs a '~'")
    lhs, rhs = formula.split("~", 1)
    return lhs.strip(), rhs.strip()


def _is_factor(col) -> bool:
    a = np.asarray(col)
    return a.dtype.kind in "OUSb" or str(getattr(col, "dtype", "")) == "category"


def _levels(col) -> List:
    dt = getattr(col, "dtype", None)
    if str(dt) == "category":
        cats = list(col.cat.categories)
        present = set(np.asarray(col).tolist())
        return [c for c in cats if c in present]            # drop.unused.levels = TRUE
    a = np.asarray(col)
    if a.dtype.kind == "b":
        return [False, True] if a.any() and not a.all() else sorted(set(a.tolist()))
    return sorted(set(a.tolist()))


def _term_columns(name: str, col, intercept_absorbed: bool):
    """Columns contributed by one variable: [(label, values)]."""
    if _is_factor(col):
        lv = _levels(col)
        a = np.asarray(col)
        use = lv[1:] if intercept_absorbed else lv
        return [(f"{name}{l if not isinstance(l, (bool, np.bool_)) else str(bool(l)).upper()}",
                 (a == l).astype(np.float64)) for l in use]
    a = np.asarray(col, dtype=np.float64)
    return [(name, a)]


def parse_rhs(rhs: str) -> Tuple[bool, List[Tuple[str, ...]]]:
    """-> (has_intercept, ordered list of terms, each a tuple of variable names)."""
    rhs = rhs.replace(" ", "")
    if not rhs:
        raise FormulaError("empty right-hand side")
    intercept = True
    terms: List[Tuple[str, ...]] = []
    for piece in filter(None, rhs.replace("-", "+-").split("+")):
        if piece in ("0", "-1"):
            intercept = False
            continue
        if piece == "1":
            continue
        if "-" in piece or "(" in piece or "^" in piece or "/" in piece or "|" in piece:
            raise FormulaError(f"unsupported formula syntax in term '{piece}'")
        if "*" in piece:
            vars_ = piece.split("*")
            # a*b = a + b + a:b (and the higher-order products for more factors)
            n = len(vars_)
            for order in range(1, n + 1):
                from itertools import combinations
                for combo in combinations(vars_, order):
                    if tuple(combo) not in terms:
                        terms.append(tuple(combo))
        else:
            t = tuple(piece.split(":"))
            if t not in terms:
                terms.append(t)
    # R orders terms by interaction order (main effects first)
    terms.sort(key=len)
    return intercept, terms


def model_matrix(rhs: str, data, rows: Sequence[int] | None = None):
    """Returns (X [n x p] float64, column names, assign vector as model.matrix's attr)."""
    intercept, terms = parse_rhs(rhs)
    get = (lambda k: data[k]) if not hasattr(data, "iloc") else (lambda k: data[k])
    n = len(data) if rows is None else len(rows)

    def col(name):
        try:
            c = get(name)
        except Exception as e:  # noqa: BLE001
            raise FormulaError(f"variable '{name}' not found in data") from e
        if rows is not None:
            c = c.iloc[list(rows)] if hasattr(c, "iloc") else np.asarray(c)[list(rows)]
        return c

    cols: List[np.ndarray] = []
    names: List[str] = []
    assign: List[int] = []
    if intercept:
        cols.append(np.ones(n))
        names.append("(Intercept)")
        assign.append(0)
    absorbed = intercept
    for ti, term in enumerate(terms, start=1):
        parts = []
        for v in term:
            c = col(v)
            # a factor is coded with full dummies only when nothing has absorbed its constant yet
            full = _is_factor(c) and not absorbed and len(term) == 1
            parts.append(_term_columns(v, c, intercept_absorbed=not full))
            if full:
                absorbed = True
        # Kronecker product of the parts' columns, first variable varying fastest (as R does)
        combos = [([], np.ones(n))]
        for p in parts:
            combos = [(lbl + [l2], val * v2) for (l2, v2) in p for (lbl, val) in combos]
        for lbl, val in combos:
            cols.append(val)
            names.append(":".join(lbl))
            assign.append(ti)
    if not cols:
        raise FormulaError("model has no columns")
    X = np.column_stack(cols).astype(np.float64)
    return X, names, assign
