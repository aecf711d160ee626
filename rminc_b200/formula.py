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


def _term_columns(name: str, col, contrasts: bool):
    """Columns contributed by one variable of a term: [(label, values)].  A factor is coded by treatment contrasts
    (baseline level dropped) or by a full set of indicator columns."""
    if _is_factor(col):
        lv = _levels(col)
        a = np.asarray(col)
        use = lv[1:] if contrasts else lv
        return [(f"{name}{l if not isinstance(l, (bool, np.bool_)) else str(bool(l)).upper()}",
                 (a == l).astype(np.float64)) for l in use]
    a = np.asarray(col, dtype=np.float64)
    return [(name, a)]


def _variables_in_order(rhs: str) -> List[str]:
    """Variable names in order of first appearance (R orders the variables of an interaction label that way)."""
    out: List[str] = []
    for tok in rhs.replace(" ", "").replace("-", "+").replace("*", "+").replace(":", "+").split("+"):
        if tok and tok not in ("0", "1") and tok not in out:
            out.append(tok)
    return out


def parse_rhs(rhs: str) -> Tuple[bool, List[Tuple[str, ...]]]:
    """-> (has_intercept, ordered list of terms, each a tuple of variable names in order of first appearance)."""
    rhs = rhs.replace(" ", "")
    if not rhs:
        raise FormulaError("empty right-hand side")
    order = {v: i for i, v in enumerate(_variables_in_order(rhs))}
    intercept = True
    terms: List[Tuple[str, ...]] = []

    def add(vars_):
        t = tuple(sorted(set(vars_), key=lambda v: order[v]))
        if t not in terms:
            terms.append(t)

    for piece in filter(None, rhs.replace("-", "+-").split("+")):
        if piece in ("0", "-1"):
            intercept = False
            continue
        if piece == "1":
            continue
        if "-" in piece or "(" in piece or "^" in piece or "/" in piece or "|" in piece:
            raise FormulaError(f"unsupported formula syntax in term '{piece}'")
        if "*" in piece:
            from itertools import combinations
            # a*b = a + b + a:b (and the higher-order products for more factors); a factor of the product may itself
            # be an interaction (a*b:c)
            vars_ = [tuple(f.split(":")) for f in piece.split("*")]
            for k in range(1, len(vars_) + 1):
                for combo in combinations(vars_, k):
                    add([v for part in combo for v in part])
        else:
            add(piece.split(":"))
    # R orders terms by interaction order (main effects first), keeping formula order within an order
    terms.sort(key=len)
    return intercept, terms


def model_matrix(rhs: str, data, rows: Sequence[int] | None = None):
    """Returns (X [n x p] float64, column names, assign vector as model.matrix's attr).

    Factor coding follows R's model.matrix (src/library/stats/src/model.c, TermCode; Chambers & Hastie, Statistical
    Models in S, p. 38): a factor f of a term T is coded by contrasts when T without f is empty or contained in an
    EARLIER term of the model, and by a full set of indicator columns otherwise -- `~ Sex:age` gives SexF:age and
    SexM:age, `~ age + Sex:age` gives age and age:SexM.  Without an intercept the first factor met (terms in order,
    variables in order) is switched to indicator coding."""
    intercept, terms = parse_rhs(rhs)
    cache: Dict[str, object] = {}

    def col(name):
        if name in cache:
            return cache[name]
        try:
            c = data[name]
        except Exception as e:  # noqa: BLE001
            raise FormulaError(f"variable '{name}' not found in data") from e
        if rows is not None:
            c = c.iloc[list(rows)] if hasattr(c, "iloc") else np.asarray(c)[list(rows)]
        cache[name] = c
        return c

    if rows is not None:
        n = len(rows)
    elif terms:
        n = len(np.asarray(col(terms[0][0])))
    else:
        n = len(data)
    # TermCode: 1 = contrasts, 2 = indicator (dummy) coding, per (term, factor variable)
    code: Dict[Tuple[int, str], int] = {}
    for ti, term in enumerate(terms):
        for v in term:
            if not _is_factor(col(v)):
                continue
            rest = set(term) - {v}
            marginal = not rest or any(rest <= set(prev) for prev in terms[:ti])
            code[(ti, v)] = 1 if marginal else 2
    if not intercept:
        first = next(((ti, v) for ti, term in enumerate(terms) for v in term if _is_factor(col(v))), None)
        if first is not None:
            code[first] = 2

    cols: List[np.ndarray] = []
    names: List[str] = []
    assign: List[int] = []
    if intercept:
        cols.append(np.ones(n))
        names.append("(Intercept)")
        assign.append(0)
    for ti, term in enumerate(terms):
        parts = [_term_columns(v, col(v), contrasts=code.get((ti, v), 1) == 1) for v in term]
        # Kronecker product of the parts' columns, first variable varying fastest (as R does)
        combos = [([], np.ones(n))]
        for p in parts:
            combos = [(lbl + [l2], val * v2) for (l2, v2) in p for (lbl, val) in combos]
        for lbl, val in combos:
            cols.append(val)
            names.append(":".join(lbl))
            assign.append(ti + 1)
    if not cols:
        raise FormulaError("model has no columns")
    X = np.column_stack(cols).astype(np.float64)
    return X, names, assign
