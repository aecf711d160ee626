"""The R-formula front end against known model.matrix output (host logic only, no GPU).

The expected column sets are what R's model.matrix returns for these formulas: the coding rule is TermCode's
(src/library/stats/src/model.c; Statistical Models in S, p. 38) -- a factor of a term is coded by contrasts when the
term without it is empty or contained in an earlier term, by a full set of indicators otherwise; without an intercept
the first factor met is switched to indicators.  mincLm / vertexLm / anatLm / mincAnova build their design with it
(R/minc_voxel_statistics.R:773-812)."""
import numpy as np
import pytest

from rminc_b200.formula import FormulaError, model_matrix, parse_rhs

DATA = {"Sex": np.array(list("FMFMFM")), "age": np.arange(6.0) + 1, "Group": np.array(list("AABBAB")),
        "flag": np.array([True, False, True, True, False, False])}

CASES = [
    ("Sex", ["(Intercept)", "SexM"], [0, 1]),
    ("Sex + age", ["(Intercept)", "SexM", "age"], [0, 1, 2]),
    ("Sex - 1", ["SexF", "SexM"], [1, 1]),
    ("0 + Sex", ["SexF", "SexM"], [1, 1]),
    ("Sex + Group - 1", ["SexF", "SexM", "GroupB"], [1, 1, 2]),
    ("age + Sex - 1", ["age", "SexF", "SexM"], [1, 2, 2]),
    # interactions whose margins are absent keep the factor's full indicator coding (the advisor's case)
    ("Sex:age", ["(Intercept)", "SexF:age", "SexM:age"], [0, 1, 1]),
    ("Sex + Sex:age", ["(Intercept)", "SexM", "SexF:age", "SexM:age"], [0, 1, 2, 2]),
    # ... and contrasts once the margin is in the model
    ("age + Sex:age", ["(Intercept)", "age", "age:SexM"], [0, 1, 2]),
    ("Sex*age", ["(Intercept)", "SexM", "age", "SexM:age"], [0, 1, 2, 3]),
    ("age*Sex", ["(Intercept)", "age", "SexM", "age:SexM"], [0, 1, 2, 3]),
    ("Sex*Group", ["(Intercept)", "SexM", "GroupB", "SexM:GroupB"], [0, 1, 2, 3]),
    ("Sex:Group", ["(Intercept)", "SexF:GroupA", "SexM:GroupA", "SexF:GroupB", "SexM:GroupB"], [0, 1, 1, 1, 1]),
    ("flag", ["(Intercept)", "flagTRUE"], [0, 1]),
]


@pytest.mark.parametrize("rhs,names,assign", CASES)
def test_model_matrix_columns_follow_r(rhs, names, assign):
    X, got_names, got_assign = model_matrix(rhs, DATA)
    assert got_names == names
    assert got_assign == assign
    assert X.shape == (6, len(names))


def test_model_matrix_values():
    X, names, _ = model_matrix("Sex:age", DATA)
    m = (DATA["Sex"] == "M").astype(float)
    np.testing.assert_array_equal(X, np.column_stack([np.ones(6), (1 - m) * DATA["age"], m * DATA["age"]]))
    X, names, _ = model_matrix("age + Sex:age", DATA)
    np.testing.assert_array_equal(X, np.column_stack([np.ones(6), DATA["age"], m * DATA["age"]]))
    # rows= subsets the frame and unused levels are dropped (drop.unused.levels = TRUE, :777)
    X, names, _ = model_matrix("Group", DATA, rows=[0, 1, 4])
    assert names == ["(Intercept)"] and X.shape == (3, 1)


def test_term_order_and_errors():
    assert parse_rhs("b:a + a")[1] == [("a",), ("b", "a")]          # main effects first; labels in order of appearance
    assert parse_rhs("a*b*c")[1] == [("a",), ("b",), ("c",), ("a", "b"), ("a", "c"), ("b", "c"), ("a", "b", "c")]
    with pytest.raises(FormulaError):
        model_matrix("missing", DATA)
    with pytest.raises(FormulaError):
        model_matrix("(age + Sex)^2", DATA)
