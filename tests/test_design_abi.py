"""CPU-side tests of the product library: the C-ABI surface and the host design factorisation.

No compute entry point can run here (no GPU): they must fail loudly with RMG_ERR_NO_DEVICE.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, have_gpu
from helpers import design_cfg1, design_cfg2, design_cov


def header_symbols():
    src = open(os.path.join(ROOT, "include", "rminc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rmg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    from rminc_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 13
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/rminc_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == syms
    assert lib.rmg_version() >= 100 and lib.rmg_max_p() == 32


def test_library_is_torch_free(lib):
    import subprocess
    from rminc_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "libR" not in out and "python" not in out


def test_design_factor_matches_oracle(lib, oracle):
    from rminc_b200 import core
    rng = np.random.default_rng(2)
    designs = [design_cfg1(20), design_cfg2(500), design_cov(2000, 7), design_cov(12, 6), np.ones((9, 1)),
               rng.normal(size=(30, 5)),                       # no intercept
               design_cov(64, 12), design_cov(100, 32)]
    for X in designs:
        n, p = X.shape
        d, o = core.design_factor(X), oracle.design_probe(X)
        assert d["info"] == 0 and o["info"] == 0
        assert list(d["pivot"]) == list(o["pivot"]) and d["rank"] == o["rank"] == p
        assert np.array_equal(d["cdiag"], o["c"])              # same operation order -> same bits
        Q, Rinv = d["Q"], d["Rinv"]
        assert np.allclose(Q.T @ Q, np.eye(p), atol=1e-13)
        assert np.allclose(Q @ np.linalg.inv(Rinv), X[:, d["pivot"]], rtol=1e-12, atol=1e-12 * np.abs(X).max())
        assert np.allclose(Rinv @ o["R"], np.eye(p), atol=1e-10)


def test_design_factor_rank_deficient_and_singular(lib, oracle):
    from rminc_b200 import core
    rng = np.random.default_rng(3)
    n = 40
    g = (np.arange(n) % 2).astype(float)
    age = np.round(rng.normal(60, 10, n), 1)
    for X in (np.column_stack([np.ones(n), g, 1 - g, age]), np.column_stack([np.ones(n), g, g, age])):
        d, o = core.design_factor(X), oracle.design_probe(X)
        assert list(d["pivot"]) == list(o["pivot"]) == [0, 1, 3, 2]
        assert d["rank"] == o["rank"] == 3 and d["info"] == 0
        assert np.array_equal(d["cdiag"], o["c"])
        assert (d["Q"][:, 3] == 0).all() and (d["Rinv"][3] == 0).all() and (d["Rinv"][:, 3] == 0).all()
    Xz = np.column_stack([np.ones(n), g, np.zeros(n), age])
    d = core.design_factor(Xz)
    assert d["info"] == 4 == oracle.design_probe(Xz)["info"]
    # the error text is the reference's (src/modelling_functions.c:554)
    err = C.create_string_buffer(256)
    Xf = np.asfortranarray(Xz)
    rc = lib.rmg_design_factor(Xf.ctypes.data, n, 4, None, None, None, None, None, None, err, 256)
    assert rc == 2 and err.value.decode() == "element (4, 4) is zero, so the inverse cannot be computed"


def test_argument_validation(lib):
    from rminc_b200 import core
    X = design_cfg1(20)
    with pytest.raises(ValueError):
        core.lm_fit(np.zeros((10, 19)), X)
    with pytest.raises(ValueError):
        core.lm_fit(np.zeros((10, 20)), X, mask=np.ones(9))
    with pytest.raises(core.RmincGpuError) as e:
        core.lm_fit(np.zeros((10, 20)), np.ones((20, 33)))
    assert e.value.code == 1
    with pytest.raises(core.RmincGpuError) as e:          # non-finite design
        core.lm_fit(np.zeros((10, 20)), np.full((20, 2), np.nan))
    assert e.value.code == 1
    with pytest.raises(core.SingularDesignError):         # singular design fails before any GPU work
        core.lm_fit(np.zeros((10, 20)), np.column_stack([np.ones(20), np.zeros(20)]))
    with pytest.raises(ValueError):
        core.randomize(np.zeros((10, 20)), X, np.full((20, 3), 20), [4])


@pytest.mark.skipif(have_gpu(), reason="only meaningful on a machine without a CUDA device")
def test_no_cpu_fallback(lib):
    """Without a GPU every compute entry point refuses; nothing silently computes on the host."""
    from rminc_b200 import core
    X = design_cfg1(20)
    Y = np.random.default_rng(0).normal(size=(64, 20))
    assert core.device_count() == 0
    for call in (lambda: core.lm_fit(Y, X), lambda: core.vertex_lm(Y, X), lambda: core.lm_fit(Y, X, n_gpus=1),
                 lambda: core.randomize(Y, X, np.arange(20)[:, None], [4]),
                 lambda: core.randomize(Y, X, np.arange(20)[:, None], [4], n_gpus=2)):
        with pytest.raises(core.RmincGpuError) as e:
            call()
        assert e.value.code == 4 and "no CPU fallback" in str(e.value)
    # ... and so does every entry point of the rows next to the path
    F = np.asfortranarray(Y.astype(np.float32))
    idx = np.arange(20)[:, None]
    adj = [[1]] + [[v - 1, v + 1] for v in range(1, 63)] + [[62]]
    more = {
        "randomize_stored": lambda: core.randomize_stored(F, X, idx, [4]),
        "randomize_stored n_gpus": lambda: core.randomize_stored(F, X, idx, [4], n_gpus=2),
        "volume_tfce": lambda: core.volume_tfce(Y[:, 0], (4, 4, 4)),
        "randomize_volume_tfce": lambda: core.randomize_volume_tfce(Y, X, idx, [4], (4, 4, 4)),
        "randomize_volume_tfce n_gpus": lambda: core.randomize_volume_tfce(Y, X, idx, [4], (4, 4, 4), n_gpus=2),
        "graph_tfce": lambda: core.graph_tfce(Y[:, 0], adj),
        "randomize_tfce": lambda: core.randomize_tfce(Y, X, idx, [4], adj),
        "randomize_tfce n_gpus": lambda: core.randomize_tfce(Y, X, idx, [4], adj, n_gpus=2),
        "lm_fit_stored": lambda: core.lm_fit_stored(F, X),
        "lm_fit_stored n_gpus": lambda: core.lm_fit_stored(F, X, n_gpus=2),
        "lm_fit_cov": lambda: core.lm_fit_cov(Y, Y[:, ::-1].copy(), X),
        "anova_fit": lambda: core.anova_fit(Y, X, [0, 1]),
        "vertex_wlm": lambda: core.vertex_wlm(Y, X, np.ones(20)),
        "fdr": lambda: core.fdr(Y[:, 0], core.STAT_T, 18),
    }
    for name, call in more.items():
        with pytest.raises(core.RmincGpuError) as e:
            call()
        assert e.value.code == 4 and "no CUDA device" in str(e.value), name
    # the one host-only routine works without a device
    ptr, nb = core.neighbour_list(3, 2, 2, 6)
    assert ptr[-1] == nb.size == 2 * (2 * 2 * 2 + 3 * 1 * 2 + 3 * 2 * 1)


def test_ctypes_signatures_match_the_header():
    """A ctypes call with the wrong number or width of arguments corrupts the call silently: every prototype of
    include/rminc_b200.h is parsed and compared, parameter by parameter, with rminc_b200/_lib.py's SIGNATURES."""
    from rminc_b200 import _lib
    src = open(os.path.join(ROOT, "include", "rminc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = re.findall(r"\b(int|double|int64_t|void|void \*)\s*(rmg_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src)
    assert len(protos) == len(_lib.SIGNATURES)

    def kind(ctype):
        if ctype is None:
            return "void"
        if ctype in (C.c_void_p, C.c_char_p) or isinstance(ctype, type(C.POINTER(C.c_int))) and hasattr(ctype, "contents"):
            return "ptr"
        return {C.c_int: "int", C.c_int64: "int64", C.c_double: "double", C.c_size_t: "size_t"}[ctype]

    for ret, name, params in protos:
        res, args = _lib.SIGNATURES[name]
        assert kind(res) == {"int": "int", "double": "double", "int64_t": "int64", "void": "void", "void *": "ptr"}[ret], name
        plist = [p.strip() for p in params.split(",") if p.strip() and p.strip() != "void"]
        assert len(plist) == len(args), f"{name}: header has {len(plist)} parameters, ctypes {len(args)}"
        for i, (decl, a) in enumerate(zip(plist, args)):
            if "*" in decl:
                want = "ptr"
            else:
                base = decl.replace("const ", "").split()[0]
                want = {"int": "int", "int32_t": "int", "int64_t": "int64", "double": "double", "size_t": "size_t"}[base]
            assert kind(a) == want, f"{name}, parameter {i + 1} ({decl}): ctypes says {kind(a)}"


def test_pvalue_function_matches_textbook_and_scipy(lib):
    """rmg_pvalue (the host build of the code the FDR kernel runs): pt2 / pf through the incomplete beta function."""
    from scipy import stats as S
    from rminc_b200 import core
    assert core.pvalue(2.228139, core.STAT_T, 10) == pytest.approx(0.05, rel=2e-6)       # t_.975(10)
    assert core.pvalue(4.964603, core.STAT_F, 1, 10) == pytest.approx(0.05, rel=2e-6)     # F_.95(1, 10)
    assert core.pvalue(1.0, core.STAT_T, 1) == pytest.approx(0.5, rel=1e-14)              # Cauchy
    for df in (1, 3, 8, 50, 496, 1993):
        for t in (1e-3, 0.5, 1, 2, 5, 12, 40, 1e3):
            assert core.pvalue(t, core.STAT_T, df) == pytest.approx(2 * S.t.sf(t, df), rel=1e-11)
            assert core.pvalue(-t, core.STAT_T, df) == core.pvalue(t, core.STAT_T, df)
    for d1, d2 in ((1, 8), (3, 496), (6, 1993), (10, 3)):
        for f in (1e-6, 0.5, 2, 10, 100, 1e5):
            assert core.pvalue(f, core.STAT_F, d1, d2) == pytest.approx(S.f.sf(f, d1, d2), rel=1e-11)
    assert np.isnan(core.pvalue(float("nan"), core.STAT_T, 5)) and core.pvalue(float("inf"), core.STAT_T, 5) == 0.0
    assert core.pvalue(0.0, core.STAT_F, 2, 9) == 1.0 and core.pvalue(-1.0, core.STAT_F, 2, 9) == 1.0


def test_product_does_not_reference_the_oracle():
    """rminc_b200/ must never import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "rminc_b200")
    for dp, _, files in os.walk(pkg):
        if "build" in dp.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".c", ".R")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert not re.search(r"\boracle\b|liboracle|orc_", txt), f"{f} mentions the oracle"


def test_rshim_compiles_against_r_api_declarations():
    """rminc_b200/csrc/rshim.c (the .Call layer a maintainer adds to RMINC) cannot be built here --
    R's headers are absent -- so at least syntax-check it against the declaration-only stand-ins in
    oracle/rstub and the real include/rminc_b200.h."""
    import subprocess
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "oracle", "rstub"),
                        "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "rminc_b200", "csrc", "rshim.c")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = open(os.path.join(ROOT, "rminc_b200", "csrc", "rshim.c")).read()
    for name in ("rmincgpu_vertex_lm_loop", "rmincgpu_lm", "rmincgpu_randomize"):
        assert name in src and name in open(os.path.join(ROOT, "rminc_b200", "R", "rminc_gpu.R")).read()


def _split_top_level_args(text):
    """Split 'a, f(b, c), d' at commas outside parentheses / brackets / quotes."""
    args, depth, cur, quote = [], 0, [], None
    for ch in text:
        if quote:
            cur.append(ch)
            if ch == quote:
                quote = None
            continue
        if ch in "\"'":
            quote = ch
            cur.append(ch)
        elif ch in "([{":
            depth += 1
            cur.append(ch)
        elif ch in ")]}":
            depth -= 1
            cur.append(ch)
        elif ch == "," and depth == 0:
            args.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    if "".join(cur).strip():
        args.append("".join(cur).strip())
    return args


def test_r_wrappers_match_the_registered_routines():
    """R cannot run here, so the binding is checked statically: every .Call("name", ...) in rminc_b200/R/rminc_gpu.R
    names a routine of rshim.c's R_CallMethodDef table (the analogue of src/RMINC_init.c:57-82) and passes exactly the
    registered number of arguments, and every registered routine is defined with that many SEXP parameters."""
    import re
    csrc = open(os.path.join(ROOT, "rminc_b200", "csrc", "rshim.c")).read()
    rsrc = "\n".join(re.sub(r"#.*", "", open(os.path.join(ROOT, "rminc_b200", "R", f)).read())          # comments out
                     for f in ("rminc_gpu.R", os.path.join("tests", "test_rminc_gpu.R")))
    table = {m.group(1): int(m.group(2)) for m in re.finditer(r'\{"(rmincgpu_\w+)",\s*\(DL_FUNC\)&\1,\s*(\d+)\}', csrc)}
    assert len(table) >= 13
    for name, arity in table.items():
        m = re.search(r"SEXP\s+" + name + r"\s*\(([^)]*)\)", csrc)
        assert m, f"{name} is registered but not defined"
        assert len([a for a in m.group(1).split(",") if a.strip() and a.strip() != "void"]) == arity, f"{name}: definition vs registered arity {arity}"
    calls = 0
    for m in re.finditer(r'\.Call\(', rsrc):
        depth, i = 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(rsrc[i], 0)
            i += 1
        args = _split_top_level_args(rsrc[m.end():i - 1])
        name = args[0].strip("\"'")
        if any(a.replace(" ", "") == 'PACKAGE="RMINC"' for a in args):
            continue                                  # the maintainer's test file also calls RMINC's own routines
        assert name in table, f".Call to unregistered routine {name}"
        passed = [a for a in args[1:] if not a.startswith("PACKAGE")]
        assert len(passed) == table[name], f".Call({name}) passes {len(passed)} arguments, {table[name]} registered"
        calls += 1
    assert calls >= 12

