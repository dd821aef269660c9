"""CPU test (no GPU): libmoihgp.so loads and exports every symbol include/moihgp_b200.h declares.
No compute entry point is called here."""
import os
import re

from conftest import ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "moihgp_b200.h")).read()
    names = set(re.findall(r"\b(moihgp_cuda_[a-z_A-Z0-9]+)\s*\(", src))
    legacy = re.findall(r"gp##XX##_([a-z0-9_]+)\(", src)
    for xx in ("32", "52"):
        names.update("gp%s_%s" % (xx, n) for n in legacy)
    return names


def test_header_symbols_are_exported():
    from multioutputihgp_b200 import _lib
    _lib.build()
    lib = _lib.load()
    decl = _declared_symbols()
    assert len(decl) == 26 + len(_lib.CUDA_NAMES), sorted(decl)
    assert decl == set(_lib.ALL_SYMBOLS)
    for s in decl:
        assert hasattr(lib, s), s


def test_reference_pywrapper_symbol_set():
    """The 13 x 2 legacy symbols are exactly the ones the reference's pywrapper.py binds (pywrapper.py:28-83)."""
    from multioutputihgp_b200 import _lib
    want = {"new", "del", "step1", "step2", "step3", "step4", "update", "lik1", "lik2", "get_params", "igp_dim", "num_param", "num_igp_param"}
    assert set(_lib.LEGACY_NAMES) == want


def test_package_import_needs_no_gpu():
    import multioutputihgp_b200 as m
    assert {"MOIHGP", "MOIHGPOnlineLearning", "MOIHGPSequences"} <= set(m.__all__)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under multioutputihgp_b200/ may import, load or link it."""
    pkg = os.path.join(ROOT, "multioutputihgp_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                for pat in (r"liboracle", r"\boracle_[a-z0-9_]+\s*\(", r"from\s+oracle", r"import\s+oracle", r"oracle\.binding", r"oracle/"):
                    assert not re.search(pat, txt), (pat, os.path.join(dp, f))
