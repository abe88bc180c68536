"""R is not installed in the build container: compile-check the .Call shim against a minimal stub of
the R C API so that signature drift between rpkg/src/shim.c and include/plaidgpu.h is caught here."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_shim_compiles_against_header():
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "rstub"),
                        "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "rpkg", "src", "shim.c")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_r_wrappers_keep_reference_signatures():
    src = open(os.path.join(ROOT, "rpkg", "R", "plaid.R")).read()
    for sig in ['plaid <- function(X, matG, stats = c("mean", "sum"), chunk = NULL, normalize = TRUE)',
                "normalize_medians <- function(x, ignore.zero = NULL)",
                'sparse_colranks <- function(X, signed = FALSE, ties.method = "average")',
                "colranks <- function(X, sparse = NULL, signed = FALSE, keep.zero = FALSE,",
                "replaid.scse <- function(X, matG, removeLog2 = NULL, scoreMean = FALSE)",
                "replaid.sing <- function(X, matG)", "replaid.ssgsea <- function(X, matG, alpha = 0)",
                "replaid.ucell <- function(X, matG, rmax = 1500)",
                "replaid.aucell <- function(X, matG, aucMaxRank = ceiling(0.05 * nrow(X)))",
                'replaid.gsva <- function(X, matG, tau = 0, rowtf = c("z", "ecdf")[1])']:
        assert sig in src, sig
    assert '[plaid] ERROR. No overlapping features.' in src


def test_r_package_exports_the_reference_namespace():
    """every export of the reference's NAMESPACE (bigomics/plaid NAMESPACE:3-16) is exported and defined here"""
    ns = open(os.path.join(ROOT, "rpkg", "NAMESPACE")).read()
    code = open(os.path.join(ROOT, "rpkg", "R", "plaid.R")).read() + open(os.path.join(ROOT, "rpkg", "R", "gmt.R")).read()
    for name in ["colranks", "gmt2mat", "mat2gmt", "normalize_medians", "plaid", "plaid.test", "read.gmt", "replaid.aucell",
                 "replaid.gsva", "replaid.scse", "replaid.sing", "replaid.ssgsea", "replaid.ucell", "write.gmt"]:
        assert f"export({name})" in ns, name
        assert f"\n{name} <- function(" in "\n" + code, name
    assert 'getOption("plaid.gpus"' in code and "plaidgpu_score_multi" in open(os.path.join(ROOT, "rpkg", "src", "shim.c")).read()
