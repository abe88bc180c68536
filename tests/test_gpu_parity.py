"""GPU parity tests proper (-m gpu): the CUDA path, reached through the C ABI (ctypes binding of
include/plaidgpu.h behind plaid_b200's R-mirroring functions), against the CPU oracle on the
same inputs.  Bars (BASELINE.json north_star): ranks bit-exact (incl. averaged ties); scores
within 1e-6 relative.  Every test runs twice (fixture `precision`):
  * "fp64"  plaidgpu_opts.exact_fp64 = 1: every add in fp64 -> tolerance 1e-11 (1e-9 for the
            pow()/exp2()-based scorers);
  * "tc"    the default: the block of high-degree rows (sparse X) / every row (dense X) goes through
            the tcgen05 int8 fixed-point pass (tc_kernels.cu): per-column 30-bit fixed point with exact
            integer accumulation, |error| <= 2^-30 max_g|x_gj| per term -> tolerance 2e-8 relative to the
            score scale (observed <= 1.6e-9).  Inputs too small for the pass take the fp64 kernels."""
import numpy as np
import pytest
import scipy.sparse as sp

import plaid_b200 as pb
from oracle import plaid_oracle as O
from plaid_b200 import synth

from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-11
TC_TOL = 2e-8


@pytest.fixture(autouse=True, params=["tc", "fp64"])
def precision(request):
    from plaid_b200 import api
    old = api.EXACT_FP64
    api.EXACT_FP64 = request.param == "fp64"
    yield request.param
    api.EXACT_FP64 = old


def tol(exact):
    """the tolerance of the all-fp64 path, or the fixed-point bound of the tensor-core pass"""
    from plaid_b200 import api
    return exact if api.EXACT_FP64 else max(exact, TC_TOL)


def _named(X, xr, xc, G, gr, gc):
    return (pb.NamedMatrix(X, xr, xc), pb.NamedMatrix(G, gr, gc), O.Named(X, xr, xc), O.Named(G, gr, gc))


# ---- the reference's own bundled fixture -----------------------------------------------------
def test_fixture_plaid_all_variants(fixture_mats, golden, gpu_ctx):
    Xg, Gg, _, _ = _named(*fixture_mats)
    r = pb.plaid(Xg, Gg, ctx=gpu_ctx)
    assert r.mat.shape == (50, 50) and r.rownames[0] == fixture_mats[5][0]
    assert rel_err(r.mat, golden["plaid_mean_norm"]) < tol(TOL)
    assert rel_err(pb.plaid(Xg, Gg, normalize=False, ctx=gpu_ctx).mat, golden["plaid_mean_raw"]) < tol(TOL)
    assert rel_err(pb.plaid(Xg, Gg, stats="sum", normalize=False, ctx=gpu_ctx).mat, golden["plaid_sum_raw"]) < tol(TOL)


def test_fixture_scorers(fixture_mats, golden, gpu_ctx):
    Xg, Gg, _, _ = _named(*fixture_mats)
    assert rel_err(pb.replaid_scse(Xg, Gg, ctx=gpu_ctx).mat, golden["scse_default"]) < tol(1e-9)
    assert rel_err(pb.replaid_scse(Xg, Gg, removeLog2=False, scoreMean=True, ctx=gpu_ctx).mat,
                   golden["scse_mean_nolog"]) < tol(TOL)
    assert rel_err(pb.replaid_sing(Xg, Gg, ctx=gpu_ctx).mat, golden["sing"]) < tol(TOL)
    assert rel_err(pb.replaid_ssgsea(Xg, Gg, ctx=gpu_ctx).mat, golden["ssgsea_a0"]) < tol(TOL)
    assert rel_err(pb.replaid_ssgsea(Xg, Gg, alpha=0.25, ctx=gpu_ctx).mat, golden["ssgsea_a025"]) < tol(1e-9)
    assert rel_err(pb.replaid_ucell(Xg, Gg, ctx=gpu_ctx).mat, golden["ucell"]) < tol(TOL)
    assert rel_err(pb.replaid_aucell(Xg, Gg, ctx=gpu_ctx).mat, golden["aucell"]) < tol(TOL)


def test_fixture_ranks_bit_exact(fixture_mats, golden, gpu_ctx):
    X = fixture_mats[0]
    r = pb.sparse_colranks(X, ctx=gpu_ctx)
    assert np.array_equal(r.indices, X.indices) and np.array_equal(r.indptr, X.indptr)
    assert np.array_equal(r.data, golden["sparse_colranks_avg"])
    assert np.array_equal(pb.sparse_colranks(X, signed=True, ties_method="min", ctx=gpu_ctx).data,
                          golden["sparse_colranks_min_signed"])
    assert np.array_equal(pb.colranks(X, ctx=gpu_ctx), golden["colranks_dense_avg"])
    assert np.array_equal(pb.colranks(X, ties_method="min", ctx=gpu_ctx), golden["colranks_dense_min"])


# ---- synthetic configs (BASELINE.json) at sizes the oracle finishes in seconds ----------------
@pytest.fixture(scope="module")
def c1_like():
    """C1 at its stated size (BASELINE.json configs[0]): pbmc3k-shaped sparse X 13,714 x 2,700 x 50
    hallmark-sized sets; rows matched by NAME with shuffled, partially overlapping names."""
    P, N, S = 13714, 2700, 50
    X = synth.sparse_x_numpy(P, N, seed=synth.SEED0 + 0)
    G = synth.genesets_numpy(4386, S, seed=synth.SEED0 + 100, size_cap=(32, 200))
    rng = np.random.default_rng(5)
    gnames = [f"SYM{k}" for k in range(4386)]
    xnames = np.array(gnames + [f"OTHER{k}" for k in range(P - 4386)])
    xnames = list(xnames[rng.permutation(P)])
    return X, xnames, [f"c{k}" for k in range(N)], G, gnames, synth.set_names(S)


def test_c1_plaid_sparse(c1_like, gpu_ctx):
    Xg, Gg, Xo, Go = _named(*c1_like)
    for stats in ["mean", "sum"]:
        for norm in [False, True]:
            got = pb.plaid(Xg, Gg, stats=stats, normalize=norm, ctx=gpu_ctx).mat
            assert rel_err(got, O.plaid(Xo, Go, stats=stats, normalize=norm).mat) < tol(TOL)


def test_multi_tile_many_sets(gpu_ctx):
    """S large enough for several shared-memory tiles + a ragged last tile."""
    P, N, S = 3000, 64, 9001
    X = synth.sparse_x_numpy(P, N, seed=21)
    G = synth.genesets_numpy(P, S, seed=22, size_cap=(5, 300))
    names = synth.gene_names(P)
    Xg, Gg, Xo, Go = _named(X, names, None, G, names, None)
    for norm in [False, True]:
        assert rel_err(pb.plaid(Xg, Gg, normalize=norm, ctx=gpu_ctx).mat, O.plaid(Xo, Go, normalize=norm).mat) < tol(TOL)
    info = gpu_ctx.plan_info()
    assert info["n_tiles"] >= 3 and info["n_tiles"] * info["tile_sets"] >= S


def test_tail_pass_tiles_chunks_and_signs(gpu_ctx, monkeypatch):
    """The tail pass (tail_kernels.cu: rows outside the tensor-core block, gene-major cell tiles of 1,056 columns,
    integer lo / hi accumulators): several tiles with a ragged last one, one tile per chunk and all tiles in one
    chunk (bit-identical), stored negative values (borrows through the high word) and a rank scorer whose
    per-entry terms change sign; all against the oracle."""
    from plaid_b200 import api
    P, N, S = 3000, 2300, 2500
    X = synth.sparse_x_numpy(P, N, seed=41).tocsc()
    rng = np.random.default_rng(42)
    X.data = X.data * np.where(rng.random(X.nnz) < 0.3, -1.0, 1.0)  # 30 % negative stored entries
    G = synth.genesets_numpy(P, S, seed=43, size_cap=(5, 400))
    names = synth.gene_names(P)
    Xg, Gg, Xo, Go = _named(X, names, None, G, names, None)
    outs = {}
    for tiles in ("1", "64"):
        monkeypatch.setenv("PLAIDGPU_TAIL_TILES", tiles)
        a = pb.plaid(Xg, Gg, normalize=False, ctx=gpu_ctx).mat
        b = pb.replaid_sing(Xg, Gg, ctx=gpu_ctx).mat
        c = pb.plaid(Xg, Gg, ctx=gpu_ctx).mat
        outs[tiles] = (a, b, c)
    info = gpu_ctx.plan_info()
    if not api.EXACT_FP64:
        assert info["tc_rows"] > 0 and info["tail_rows"] > 0 and N > 2 * info["tail_tile_cells"]
    for u, v in zip(outs["1"], outs["64"]):
        assert np.array_equal(u, v)
    assert rel_err(outs["1"][0], O.plaid(Xo, Go, normalize=False).mat) < tol(TOL)
    assert rel_err(outs["1"][1], O.replaid_sing(Xo, Go).mat) < tol(TOL)
    assert rel_err(outs["1"][2], O.plaid(Xo, Go).mat) < tol(TOL)


def test_c2_dense_bulk(gpu_ctx):
    P, N, S = 2000, 48, 3000
    X = synth.dense_x_numpy(P, N, seed=synth.SEED0 + 1)
    G = synth.genesets_numpy(P, S, seed=synth.SEED0 + 101, size_cap=(5, 500))
    names = synth.gene_names(P)
    Xg, Gg, Xo, Go = _named(X, names, None, G, names, None)
    assert rel_err(pb.plaid(Xg, Gg, ctx=gpu_ctx).mat, O.plaid(Xo, Go).mat) < tol(TOL)
    assert rel_err(pb.plaid(Xg, Gg, stats="sum", normalize=False, ctx=gpu_ctx).mat,
                   O.plaid(Xo, Go, stats="sum", normalize=False).mat) < tol(TOL)


def test_c2_dense_bulk_full_size(gpu_ctx):
    """C2 at its stated size (BASELINE.json configs[1]): dense 20,000 x 1,000 with 30,000 sets; the default
    plaid() against the oracle on all 30,000 sets, the un-normalised sums on 2,500 sampled sets."""
    P, N, S = 20000, 1000, 30000
    X = synth.dense_x_numpy(P, N, seed=synth.SEED0 + 1)
    Gp, Gi = synth.genesets_torch(P, S, seed=synth.SEED0 + 3, device="cuda:0")  # the bench's collection
    G = sp.csc_matrix((np.ones(Gi.size), Gi, Gp), shape=(P, S))
    names = synth.gene_names(P)
    Xg, Gg, Xo, Go = _named(X, names, None, G, names, None)
    assert rel_err(pb.plaid(Xg, Gg, ctx=gpu_ctx).mat, O.plaid(Xo, Go).mat) < tol(TOL)
    pick = np.sort(np.random.default_rng(3).choice(S, size=2500, replace=False))
    got = pb.plaid(Xg, Gg, stats="sum", normalize=False, ctx=gpu_ctx).mat[pick]
    assert rel_err(got, O.plaid(Xo, O.Named(G[:, pick], names), stats="sum", normalize=False).mat) < tol(TOL)


def test_c3_rank_scorers_sparse(gpu_ctx):
    P, N, S = 4000, 96, 700
    X = synth.sparse_x_numpy(P, N, seed=synth.SEED0 + 2)
    G = synth.genesets_numpy(P, S, seed=synth.SEED0 + 102, size_cap=(5, 400))
    names = synth.gene_names(P)
    Xg, Gg, Xo, Go = _named(X, names, None, G, names, None)
    assert rel_err(pb.replaid_ssgsea(Xg, Gg, ctx=gpu_ctx).mat, O.replaid_ssgsea(Xo, Go).mat) < tol(TOL)
    assert rel_err(pb.replaid_sing(Xg, Gg, ctx=gpu_ctx).mat, O.replaid_sing(Xo, Go).mat) < tol(TOL)
    assert rel_err(pb.replaid_ucell(Xg, Gg, rmax=300, ctx=gpu_ctx).mat, O.replaid_ucell(Xo, Go, rmax=300).mat) < tol(TOL)
    assert rel_err(pb.replaid_aucell(Xg, Gg, ctx=gpu_ctx).mat, O.replaid_aucell(Xo, Go).mat) < tol(TOL)
    assert rel_err(pb.replaid_scse(Xg, Gg, ctx=gpu_ctx).mat, O.replaid_scse(Xo, Go).mat) < tol(1e-9)


def test_rank_scorers_dense_input(gpu_ctx):
    P, N, S = 1500, 20, 200
    X = synth.dense_x_numpy(P, N, seed=31)
    X[::9] = np.round(X[::9])  # ties
    G = synth.genesets_numpy(P, S, seed=32, size_cap=(5, 200))
    names = synth.gene_names(P)
    Xg, Gg, Xo, Go = _named(X, names, None, G, names, None)
    assert rel_err(pb.replaid_sing(Xg, Gg, ctx=gpu_ctx).mat, O.replaid_sing(Xo, Go).mat) < tol(TOL)
    assert rel_err(pb.replaid_ssgsea(Xg, Gg, ctx=gpu_ctx).mat, O.replaid_ssgsea(Xo, Go).mat) < tol(TOL)
    assert rel_err(pb.replaid_ucell(Xg, Gg, ctx=gpu_ctx).mat, O.replaid_ucell(Xo, Go).mat) < tol(TOL)
    assert rel_err(pb.replaid_aucell(Xg, Gg, ctx=gpu_ctx).mat, O.replaid_aucell(Xo, Go).mat) < tol(TOL)
    assert rel_err(pb.replaid_scse(Xg, Gg, ctx=gpu_ctx).mat, O.replaid_scse(Xo, Go).mat) < tol(1e-9)


def test_gsva_z(fixture_mats, golden, gpu_ctx):
    """replaid.gsva(rowtf="z"): row z-transform across samples, signed dense ranks, plaid (R/plaid.R:338-363)."""
    Xg, Gg, _, _ = _named(*fixture_mats)
    assert rel_err(pb.replaid_gsva(Xg, Gg, ctx=gpu_ctx).mat, golden["gsva_z"]) < tol(1e-9)  # sparse input, densified
    P, N, S = 1200, 40, 150
    X = synth.dense_x_numpy(P, N, seed=35)
    X[5] = 3.25  # constant row: sd = 0 -> z = 0 for every sample (tie group at zero)
    X[6] = X[7]  # duplicated rows -> tied z -> averaged ranks
    G = synth.genesets_numpy(P, S, seed=36, size_cap=(5, 150))
    names = synth.gene_names(P)
    Xg, Gg, Xo, Go = _named(X, names, None, G, names, None)
    assert rel_err(pb.replaid_gsva(Xg, Gg, ctx=gpu_ctx).mat, O.replaid_gsva(Xo, Go).mat) < tol(1e-9)
    assert rel_err(pb.replaid_gsva(Xg, Gg, tau=0.5, ctx=gpu_ctx).mat, O.replaid_gsva(Xo, Go, tau=0.5).mat) < tol(1e-9)
    with pytest.raises(ValueError):
        pb.replaid_gsva(Xg, Gg, rowtf="nope", ctx=gpu_ctx)
    # rowtf = "ecdf": per-gene ECDF across samples (ties: every tied sample gets the fraction <= its value)
    X[9] = np.round(X[9])
    Xg, Gg, Xo, Go = _named(X, names, None, G, names, None)
    assert rel_err(pb.replaid_gsva(Xg, Gg, rowtf="ecdf", ctx=gpu_ctx).mat, O.replaid_gsva(Xo, Go, rowtf="ecdf").mat) < tol(1e-9)


def test_wide_dynamic_range_columns_keep_their_zero_pattern(gpu_ctx):
    """a column whose entries span more than 2^31 cannot be held in the 30-bit per-column fixed point: its smallest
    entries would become exact zeros, and zeros carry meaning (normalize_medians drops them, R/plaid.R:557-566) — the
    library must notice and score the call in fp64 (found by tools/fuzz_paths.py: 1.5e-3 off without the guard)"""
    rng = np.random.default_rng(35)
    P, N, S = 3000, 49, 1500
    X = sp.random(P, N, density=0.01, format="csc", random_state=3, data_rvs=lambda n: rng.lognormal(0.5, 0.8, n))
    X.data *= 10.0 ** rng.integers(-6, 7, size=X.data.size)
    G = synth.genesets_numpy(P, S, seed=36, size_cap=(3, 700))
    names = synth.gene_names(P)
    Xg, Gg, Xo, Go = _named(X, names, None, G, names, None)
    want = O.plaid(Xo, Go).mat
    got = pb.plaid(Xg, Gg, ctx=gpu_ctx).mat
    assert rel_err(got, want) < tol(TOL)
    raw_w, raw_g = O.plaid(Xo, Go, normalize=False).mat, pb.plaid(Xg, Gg, normalize=False, ctx=gpu_ctx).mat
    assert np.array_equal(raw_g == 0, raw_w == 0)          # the zero pattern of the scores is the reference's
    D = X.toarray()                                         # the same through the dense-input path
    assert rel_err(pb.plaid(pb.NamedMatrix(D, names), Gg, ctx=gpu_ctx).mat, O.plaid(O.Named(D, names), Go).mat) < tol(TOL)


# ---- ranking: bit-exact, adversarial -----------------------------------------------------------
@pytest.mark.parametrize("ties", ["average", "min", "max"])
@pytest.mark.parametrize("signed", [False, True])
def test_colranks_bit_exact(ties, signed, gpu_ctx):
    P, N = 700, 40
    X = synth.sparse_x_numpy(P, N, seed=41, density=0.15).tolil()
    X[:, 3] = 0  # empty column
    X[:5, 4] = 1.25  # all-tied column
    X = sp.csc_matrix(X)
    X.data[::5] *= -1.0
    X.data[::13] = 0.0  # explicit stored zeros
    X.data[7] = -0.0
    got = pb.sparse_colranks(X, signed=signed, ties_method=ties, ctx=gpu_ctx)
    assert np.array_equal(got.data, O.sparse_colranks(X, signed=signed, ties_method=ties).data)
    assert np.array_equal(pb.colranks(X, signed=signed, ties_method=ties, ctx=gpu_ctx),
                          O.colranks(X, signed=signed, ties_method=ties))
    D = X.toarray()
    assert np.array_equal(pb.colranks(D, signed=signed, ties_method=ties, ctx=gpu_ctx),
                          O.colranks(D, signed=signed, ties_method=ties))


def test_colranks_nan_and_large_column(gpu_ctx):
    rng = np.random.default_rng(9)
    D = np.round(rng.normal(size=(20000, 3)), 1)  # bulk-sized column (P = 20k), heavy ties
    D[5, 0] = np.nan
    got = pb.colranks(D, ctx=gpu_ctx)
    want = O.colranks(D)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)])


def test_colranks_counting_and_sorting_paths_agree(gpu_ctx):
    """k_rank ranks a column by COUNTING its distinct values when there are at most 448 of them and by sorting
    otherwise; columns on both sides of that limit (447 / 448 / 449 / 700 distinct values, 12 values, all distinct,
    negatives + stored zeros + NaN, empty) in one launch must all be bit-exact, for every ranking variant"""
    rng = np.random.default_rng(21)
    P = 6000
    cols = []
    for nd in (447, 448, 449, 700, 12, None, 3):
        n = 2500
        rows = np.sort(rng.choice(P, size=n, replace=False))
        if nd is None:
            vals = rng.normal(size=n)  # all distinct
        else:
            pool = np.round(rng.normal(size=4 * nd), 6)
            pool = np.unique(pool)[:nd]
            assert pool.size == nd
            vals = np.concatenate([pool, rng.choice(pool, size=n - nd)])  # every pool value at least once
            rng.shuffle(vals)
        cols.append((rows, vals))
    rows, vals = cols[-1]
    vals[:50] = 0.0        # stored zeros
    vals[50:60] = np.nan   # NaN entries
    vals[60:200] *= -1.0
    cols.append((np.zeros(0, dtype=np.int64), np.zeros(0)))  # an empty column
    indptr = np.concatenate([[0], np.cumsum([len(r) for r, _ in cols])])
    X = sp.csc_matrix((np.concatenate([v for _, v in cols]), np.concatenate([r for r, _ in cols]), indptr),
                      shape=(P, len(cols)))

    def same(a, b):
        assert np.array_equal(np.isnan(a), np.isnan(b))
        assert np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])

    for ties in ("average", "min", "max"):
        for signed in (False, True):
            same(pb.sparse_colranks(X, signed=signed, ties_method=ties, ctx=gpu_ctx).data,
                 O.sparse_colranks(X, signed=signed, ties_method=ties).data)
            same(pb.colranks(X, signed=signed, ties_method=ties, ctx=gpu_ctx), O.colranks(X, signed=signed, ties_method=ties))
    D = X[:, :5].toarray()
    D[D == 0] = 1.5  # dense input, 6000-row columns: few distinct values -> counting path on dense columns
    same(pb.colranks(D, ctx=gpu_ctx), O.colranks(D))


def test_colranks_order_of_appearance_ties(gpu_ctx):
    """ties.method first / last (base::rank via sparse_colranks, matrixStats::colRanks on dense input) and dense
    (matrixStats only): bit-exact vs the oracle incl. signed ranks, NaN, -0, short and long columns; the methods R does
    not offer on a branch are errors here too"""
    rng = np.random.default_rng(44)
    P = 3000
    D = np.round(rng.normal(size=(P, 6)), 1)          # heavy ties
    D[:, 1] = rng.normal(size=P)                       # all distinct
    D[::50, 2] = np.nan
    D[:, 3] = np.where(rng.random(P) < 0.5, 0.0, -0.0)
    D[:, 4] = 7.0
    Dl = np.round(rng.normal(size=(20000, 2)), 2)      # bulk-sized columns

    def same(a, b):
        assert np.array_equal(np.isnan(a), np.isnan(b))
        assert np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])

    for ties in ("first", "last", "dense"):
        for signed in (False, True):
            same(pb.colranks(D, signed=signed, ties_method=ties, ctx=gpu_ctx), O.colranks(D, signed=signed, ties_method=ties))
        same(pb.colranks(Dl, ties_method=ties, ctx=gpu_ctx), O.colranks(Dl, ties_method=ties))
    X = synth.sparse_x_numpy(2000, 30, seed=45, density=0.2)
    X.data[::7] *= -1.0
    X.data[::11] = 0.0
    for ties in ("first", "last"):
        for signed in (False, True):
            same(pb.sparse_colranks(X, signed=signed, ties_method=ties, ctx=gpu_ctx).data,
                 O.sparse_colranks(X, signed=signed, ties_method=ties).data)
            same(pb.colranks(X, keep_zero=True, signed=signed, ties_method=ties, ctx=gpu_ctx).data,
                 O.sparse_colranks(X, signed=signed, ties_method=ties).data)
    with pytest.raises(ValueError):
        pb.sparse_colranks(X, ties_method="dense", ctx=gpu_ctx)      # base::rank has no "dense"
    with pytest.raises(ValueError):
        pb.colranks(X, ties_method="random", ctx=gpu_ctx)            # R's RNG
    from plaid_b200 import _lib as L
    with pytest.raises(L.PlaidGpuError):
        pb.colranks(X, ties_method="first", ctx=gpu_ctx)             # sparseMatrixStats::colRanks: max / average / min


def test_colranks_bucket_path_many_distinct_values(gpu_ctx):
    """columns with many distinct values are ranked by the splitter / bucket path of k_rank (no sort): dense bulk
    columns of 20,000 distinct values, mixtures of large tie classes and distinct values, signed ranks, dense
    semantics with negatives where the zero group falls on / between splitters, NaN, long sparse columns"""
    rng = np.random.default_rng(33)
    P = 20000
    D = rng.normal(size=(P, 9))
    D[:, 1] = np.abs(D[:, 1]); D[rng.random(P) < 0.2, 1] = 0.0         # bulk column with 20 % zeros (a big tie class)
    D[:, 2] = np.where(rng.random(P) < 0.5, np.round(D[:, 2], 1), D[:, 2])  # half heavily tied, half distinct
    D[::97, 3] = np.nan
    D[:, 4] = np.exp(8.0 * D[:, 4])                                     # skewed over many binades
    D[:, 5] = 1.0 + rng.integers(0, 3000, size=P) * 2.0 ** -45          # 3000 classes differing in the low bits only
    D[:, 6] = np.sort(D[:, 6])                                          # sorted input: the strided sample is exact quantiles
    D[:, 7] = -np.abs(D[:, 7])
    D[:, 8] = np.concatenate([np.full(P - 300, 2.5), rng.normal(size=300)])  # one class holds 98.5 %

    def same(a, b):
        assert np.array_equal(np.isnan(a), np.isnan(b))
        assert np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])

    for ties in ("average", "min", "max"):
        for signed in (False, True):
            same(pb.colranks(D, signed=signed, ties_method=ties, ctx=gpu_ctx), O.colranks(D, signed=signed, ties_method=ties))
    # sparse columns of ~5,000 distinct stored values each, with negatives, stored zeros and NaN: sparse_colranks and
    # the dense-semantics ranks (implicit zeros form one group between the negative and the positive entries)
    n = 5000
    cols = []
    for k in range(6):
        rows = np.sort(rng.choice(P, size=n, replace=False))
        vals = rng.normal(size=n)
        if k == 1: vals = np.abs(vals)
        if k == 2: vals[:40] = 0.0
        if k == 3: vals[:7] = np.nan
        if k == 4: vals = -np.abs(vals)
        if k == 5: vals[::2] = np.round(vals[::2], 2)
        cols.append((rows, vals))
    indptr = np.concatenate([[0], np.cumsum([len(r) for r, _ in cols])])
    X = sp.csc_matrix((np.concatenate([v for _, v in cols]), np.concatenate([r for r, _ in cols]), indptr), shape=(P, len(cols)))
    for ties in ("average", "min", "max"):
        for signed in (False, True):
            same(pb.sparse_colranks(X, signed=signed, ties_method=ties, ctx=gpu_ctx).data,
                 O.sparse_colranks(X, signed=signed, ties_method=ties).data)
            same(pb.colranks(X, signed=signed, ties_method=ties, ctx=gpu_ctx), O.colranks(X, signed=signed, ties_method=ties))


# ---- normalize_medians ---------------------------------------------------------------------------
def test_normalize_medians_edge_cases(gpu_ctx):
    rng = np.random.default_rng(4)
    x = rng.normal(size=(501, 33))
    assert rel_err(pb.normalize_medians(x, ctx=gpu_ctx), O.normalize_medians(x)) < 1e-13
    y = np.abs(x)
    y[rng.random(y.shape) < 0.4] = 0.0
    y[:, 2] = 0.0  # all-zero column -> median NA -> 0
    y[:, 3] = 7.0  # constant column: every radix digit is tied
    y[10, 5] = np.nan
    assert rel_err(pb.normalize_medians(y, ctx=gpu_ctx), O.normalize_medians(y)) < 1e-13
    for iz in [False, True]:
        assert rel_err(pb.normalize_medians(y, ignore_zero=iz, ctx=gpu_ctx), O.normalize_medians(y, ignore_zero=iz)) < 1e-13
    z = rng.normal(size=(4000, 5)) * 10.0 ** rng.integers(-200, 200, size=(4000, 5))  # wild exponents
    assert rel_err(pb.normalize_medians(z, ctx=gpu_ctx), O.normalize_medians(z)) < 1e-13
    e = rng.normal(size=(64, 2))  # even count: mean of the two middles
    assert rel_err(pb.normalize_medians(e, ctx=gpu_ctx), O.normalize_medians(e)) < 1e-13


def test_normalize_medians_long_columns_single_pass_path(gpu_ctx):
    """columns >= 4096 rows go through the sampled single-pass kernel (+ exact fallback): adversarial shapes"""
    rng = np.random.default_rng(14)
    S, N = 9000, 24
    x = np.abs(rng.normal(size=(S, N))) + 0.1
    x[rng.random(x.shape) < 0.3] = 0.0          # many zeros -> the two medians differ a lot
    x[:, 1] = 0.0                                # all-zero column
    x[:, 2] = 4.5                                # constant column (every candidate ties)
    x[:, 3] = np.round(x[:, 3], 1)               # heavy ties around the median
    x[:, 4] = 0.0; x[:40, 4] = rng.random(40)    # almost all zero: few non-zeros
    x[::7, 5] = np.nan                           # NaNs dropped
    x[:, 6] = rng.normal(size=S) * 10.0 ** rng.integers(-250, 250, size=S)  # wild exponents, negatives
    x[:, 7] = np.sort(x[:, 7])                   # sorted column: a strided sample is still fine
    x[:, 8] = np.where(np.arange(S) % 2 == 0, 1.0, 2.0)  # two values, even count -> mean of middles
    assert rel_err(pb.normalize_medians(x, ctx=gpu_ctx), O.normalize_medians(x)) < 1e-13
    for iz in (False, True):
        assert rel_err(pb.normalize_medians(x, ignore_zero=iz, ctx=gpu_ctx), O.normalize_medians(x, ignore_zero=iz)) < 1e-13
    y = rng.normal(size=(30000, 6))              # C4-length columns, no zeros, negatives
    assert rel_err(pb.normalize_medians(y, ctx=gpu_ctx), O.normalize_medians(y)) < 1e-13
    # single-median kernel (ignore.zero given): its bracket and filter live in FP32 — cases where float rounding,
    # float range or the zero block could matter
    S = 30000
    w = rng.normal(size=(S, 14))
    w[rng.random(w.shape) < 0.3] = 0.0           # zeros INSIDE the bracket of the non-zero median (centred scores)
    w[:, 1] = 1.0 + rng.integers(0, 50, size=S) * 2.0 ** -40   # distinct doubles that round to ONE float
    w[:, 2] = rng.normal(size=S) * 1e-60         # every value rounds to +-0 in float
    w[:, 3] = rng.normal(size=S) * 1e200         # every value rounds to +-inf in float
    w[::3, 4] = np.inf; w[1::3, 4] = -np.inf     # infinities on both sides
    w[:, 5] = np.abs(w[:, 5]); w[:, 5][rng.random(S) < 0.5] = 0.0; w[::11, 5] = 1e-50  # zeros + tiny non-zeros
    w[:, 6] = 0.0; w[:30, 6] = rng.random(30)    # fewer than 48 valid sample points when the zeros are dropped
    w[:, 7] = -np.abs(w[:, 7]) - 1.0             # all negative
    w[:, 8] = np.floor(rng.random(S) * 3.0)      # three values, one of them zero
    w[:, 9] = rng.random(S) * 2.0 ** -140        # float-subnormal range
    w[5, 10] = np.nan                            # a single NaN -> exact fallback kernel
    w[:, 11] = np.where(rng.random(S) < 0.5, 0.0, -0.0)   # signed zeros only
    w[:, 12] = 3.0 + rng.normal(size=S) * 1e-12  # narrow spread around a float
    # each column paired with its half (medians m and m / 2 exactly): out = x - m + 0.75 m exposes m at its own scale
    for k in range(w.shape[1]):
        pair = np.column_stack([w[:, k], 0.5 * w[:, k]])
        for iz in (False, True):
            g, o = pb.normalize_medians(pair, ignore_zero=iz, ctx=gpu_ctx), O.normalize_medians(pair, ignore_zero=iz)
            fin = np.isfinite(o)
            assert np.array_equal(np.isfinite(g), fin) and np.array_equal(g[~fin], o[~fin], equal_nan=True), (k, iz)
            scale = np.max(np.abs(o[fin])) if fin.any() else 0.0
            assert np.max(np.abs(g[fin] - o[fin]), initial=0.0) <= 1e-13 * scale, (k, iz)
    with np.errstate(invalid="ignore"):
        for iz in (False, True):  # all columns in one launch (columns 3 and 4 overflow / poison mean(med): left out)
            sel = [c for c in range(w.shape[1]) if c not in (3, 4)]
            assert rel_err(pb.normalize_medians(w[:, sel], ignore_zero=iz, ctx=gpu_ctx),
                           O.normalize_medians(w[:, sel], ignore_zero=iz)) < 1e-13


# ---- reference edge cases ------------------------------------------------------------------------
def test_no_overlap_returns_none(gpu_ctx):
    X = synth.sparse_x_numpy(50, 4, seed=1, density=0.3)
    G = synth.genesets_numpy(50, 5, seed=2, size_cap=(3, 10))
    assert pb.plaid(pb.NamedMatrix(X, [f"a{k}" for k in range(50)]), pb.NamedMatrix(G, [f"b{k}" for k in range(50)]),
                    ctx=gpu_ctx) is None


def test_vector_input_single_sample_and_empty_set(gpu_ctx):
    P, S = 200, 10
    G = synth.genesets_numpy(P, S, seed=3, size_cap=(3, 20)).tolil()
    G[:, 4] = 0  # empty gene set -> score 0
    G = sp.csc_matrix(G)
    names = synth.gene_names(P)
    v = np.random.default_rng(6).normal(size=P)
    got = pb.plaid(pb.NamedMatrix(v, names), pb.NamedMatrix(G, names), normalize=False, ctx=gpu_ctx).mat
    want = O.plaid(O.Named(v, names), O.Named(G, names), normalize=False).mat
    assert got.shape == (S, 1) and got[4, 0] == 0.0
    assert rel_err(got, want) < tol(TOL)


def test_nan_propagates_only_to_sets_with_that_gene(gpu_ctx):
    P, N, S = 300, 6, 40
    X = synth.sparse_x_numpy(P, N, seed=51, density=0.3)
    X.data[3] = np.nan
    G = synth.genesets_numpy(P, S, seed=52, size_cap=(3, 40))
    names = synth.gene_names(P)
    got = pb.plaid(pb.NamedMatrix(X, names), pb.NamedMatrix(G, names), normalize=False, ctx=gpu_ctx).mat
    want = O.plaid(O.Named(X, names), O.Named(G, names), normalize=False).mat
    assert np.isnan(want).any() and not np.isnan(want).all()
    assert rel_err(got, want) < tol(TOL)


def test_chunked_crossprod_matches(gpu_ctx):
    X = synth.sparse_x_numpy(500, 30, seed=61, density=0.2)
    G = synth.genesets_numpy(500, 70, seed=62, size_cap=(3, 60))
    Gs = G @ sp.diags(1.0 / (1e-8 + np.asarray(G.sum(0)).ravel()))
    assert rel_err(pb.chunked_crossprod(Gs, X, ctx=gpu_ctx), O.chunked_crossprod(sp.csc_matrix(Gs), X)) < tol(TOL)
    assert rel_err(pb.chunked_crossprod(G, X, chunk=7, ctx=gpu_ctx), O.chunked_crossprod(G, X, chunk=7)) < tol(TOL)


def test_plaid_test_statistics(gpu_ctx):
    """plaid.test (R/plaid.R:392-474): GPU reductions + host distribution functions vs the oracle restatement"""
    from plaid_b200 import _lib as L
    P, N, S = 900, 60, 120
    X = synth.sparse_x_numpy(P, N, seed=111, density=0.3)
    G = synth.genesets_numpy(P, S, seed=112, size_cap=(5, 120))
    names = synth.gene_names(P)
    y = (np.random.default_rng(3).random(N) < 0.4).astype(int)
    sets = synth.set_names(S)
    got = pb.plaid_test(pb.NamedMatrix(X, names), y, pb.NamedMatrix(G, names, sets), ctx=gpu_ctx)
    want = O.plaid_test(O.Named(X, names, None), y, O.Named(G, names, sets))
    assert got[1] == want[1] and sorted(got[2]) == sorted(want[2])
    # rows are sorted by p.meta; align by set name (near-equal p-values may swap places at the 1e-15 level)
    go, wo = np.argsort(got[2]), np.argsort(want[2])
    assert rel_err(got[0][go], want[0][wo]) < 1e-8
    assert np.all(np.diff(got[0][:, got[1].index("p.meta")]) >= 0)
    gm = pb.group_moments(np.arange(12.0).reshape(3, 4), [0, 1, 1, 0], ctx=gpu_ctx)
    assert np.array_equal(gm, [[3, 11, 19], [9, 65, 185], [3, 11, 19], [5, 61, 181]])
    # f1 as specified: the reductions fused onto the scoring call (scores stay on the device, normalisation applied
    # in registers) are the reductions of the materialised normalised matrix, bit for bit — plaid and a rank scorer
    Xn, Gn = pb.NamedMatrix(X, names), pb.NamedMatrix(G, names, sets)
    for kw, full in ((dict(), pb.plaid(Xn, Gn, ctx=gpu_ctx)),
                     (dict(normalize=0), pb.plaid(Xn, Gn, normalize=False, ctx=gpu_ctx)),
                     (dict(scorer=L.UCELL, rmax=200.0), pb.replaid_ucell(Xn, Gn, rmax=200, ctx=gpu_ctx))):
        fused = pb.score_group_moments(Xn, Gn, y, ctx=gpu_ctx, **kw)
        assert np.array_equal(fused, pb.group_moments(full.mat, y, ctx=gpu_ctx))
    o_full = O.plaid(O.Named(X, names), O.Named(G, names)).mat
    fused = pb.score_group_moments(Xn, Gn, y, ctx=gpu_ctx)
    want_gm = np.stack([o_full[:, y == 0].sum(1), (o_full[:, y == 0] ** 2).sum(1), o_full[:, y == 1].sum(1), (o_full[:, y == 1] ** 2).sum(1)])
    assert rel_err(fused, want_gm) < tol(1e-10)


def test_degenerate_shapes(gpu_ctx):
    """N = 1, S = 1, empty matrix, empty columns, a set of ALL genes, S above 2^16 (reference benchmark: 61k sets)"""
    P = 400
    names = synth.gene_names(P)
    X = synth.sparse_x_numpy(P, 9, seed=101, density=0.2).tolil()
    X[:, 2] = 0
    X[:, 8] = 0
    X = sp.csc_matrix(X)
    G1 = sp.csc_matrix(np.ones((P, 1)))                        # one set holding every gene
    for G in (G1, synth.genesets_numpy(P, 3, seed=102, size_cap=(3, 50))):
        for norm in (False, True):
            got = pb.plaid(pb.NamedMatrix(X, names), pb.NamedMatrix(G, names), normalize=norm, ctx=gpu_ctx).mat
            assert rel_err(got, O.plaid(O.Named(X, names), O.Named(G, names), normalize=norm).mat) < tol(TOL)
    one = pb.plaid(pb.NamedMatrix(X[:, :1], names), pb.NamedMatrix(G1, names), ctx=gpu_ctx).mat
    assert one.shape == (1, 1)
    Z = sp.csc_matrix((P, 5))                                   # nothing stored at all
    Gz = synth.genesets_numpy(P, 40, seed=103, size_cap=(3, 50))
    got = pb.plaid(pb.NamedMatrix(Z, names), pb.NamedMatrix(Gz, names), ctx=gpu_ctx).mat
    assert rel_err(got, O.plaid(O.Named(Z, names), O.Named(Gz, names)).mat) < tol(TOL)
    assert np.array_equal(pb.colranks(Z, ctx=gpu_ctx), O.colranks(Z))
    assert rel_err(pb.replaid_ucell(pb.NamedMatrix(X, names), pb.NamedMatrix(Gz, names), ctx=gpu_ctx).mat,
                   O.replaid_ucell(O.Named(X, names), O.Named(Gz, names)).mat) < tol(TOL)
    # more sets than fit 16 bits (the reference benchmarks 61,459 sets): tiles + gather block + 32-bit offsets
    P2, N2, S2 = 3000, 40, 70001
    X2 = synth.sparse_x_numpy(P2, N2, seed=104)
    G2 = synth.genesets_numpy(P2, 2000, seed=105, size_cap=(5, 100))
    G2 = sp.hstack([G2] * 36).tocsc()[:, :S2]
    n2 = synth.gene_names(P2)
    got = pb.plaid(pb.NamedMatrix(X2, n2), pb.NamedMatrix(G2, n2), ctx=gpu_ctx).mat
    assert got.shape == (S2, N2)
    assert rel_err(got, O.plaid(O.Named(X2, n2), O.Named(G2, n2)).mat) < tol(TOL)


def test_sharded_protocol_is_shard_count_invariant():
    """column shards on separate contexts (begin / compute / finish with exchanged scalars) give exactly
    the single-context result, for 2 and 3 ragged shards (multi-GPU invariance)"""
    from plaid_b200 import _lib as L, sharded
    from plaid_b200.api import _matrix_struct, _opts
    P, N, S = 1200, 101, 1500
    X = synth.sparse_x_numpy(P, N, seed=91)
    G = synth.genesets_numpy(P, S, seed=92, size_cap=(5, 200))
    names = synth.gene_names(P)
    rowmap = pb.make_rowmap(names, names)
    ctxs = [pb.Context(0) for _ in range(3)]
    for c in ctxs:
        c.set_genesets(G)
    for scorer, kw in [(L.PLAID, dict(normalize=1)), (L.UCELL, dict(rmax=150.0)), (L.SSGSEA, dict(alpha=0.0)),
                       (L.SCSE, dict())]:
        whole = np.empty((S, N), order="F")
        keep = []
        M = _matrix_struct(X, keep)
        o = _opts(ctxs[0].lib, scorer=scorer, out_location=L.HOST, **kw)
        ctxs[0].check(ctxs[0].lib.plaidgpu_score(ctxs[0].h, M, rowmap.ctypes.data, o, whole.ctypes.data))
        for world in (2, 3):
            spans = [sharded.shard_columns(N, world, r) for r in range(world)]
            outs = [np.empty((S, hi - lo), order="F") for lo, hi in spans]
            mats = [_matrix_struct(X[:, lo:hi], keep) for lo, hi in spans]
            opts = [_opts(ctxs[0].lib, scorer=scorer, out_location=L.HOST, **kw) for _ in spans]
            sharded.score_multi(ctxs[:world], mats, rowmap, opts, [a.ctypes.data for a in outs], [hi - lo for lo, hi in spans])
            assert np.array_equal(np.concatenate(outs, axis=1), whole)


def test_c_level_multi_context_entry_and_pageable_ring(monkeypatch):
    """plaidgpu_score_multi (one host thread per context, blocks written straight into the caller's matrix) equals
    the one-context result bit for bit for 2 and 3 contexts, sparse and dense input, every scorer family; the
    pageable-destination path (pinned ring + copy threads) equals the direct copy into pinned memory."""
    import torch
    P, N, S = 1500, 333, 1800
    X = synth.sparse_x_numpy(P, N, seed=93)
    D = synth.dense_x_numpy(P, 97, seed=94)
    G = synth.genesets_numpy(P, S, seed=95, size_cap=(5, 200))
    names = synth.gene_names(P)
    Gn = pb.NamedMatrix(G, names)
    ctxs = [pb.Context(0) for _ in range(3)]
    calls = [("plaid", lambda c, M: pb.plaid(M, Gn, ctx=c)), ("plaid raw sum", lambda c, M: pb.plaid(M, Gn, stats="sum", normalize=False, ctx=c)),
             ("ssgsea", lambda c, M: pb.replaid_ssgsea(M, Gn, ctx=c)), ("ucell", lambda c, M: pb.replaid_ucell(M, Gn, rmax=200, ctx=c)),
             ("scse", lambda c, M: pb.replaid_scse(M, Gn, ctx=c)), ("gsva", lambda c, M: pb.replaid_gsva(M, Gn, ctx=c))]
    for label, f in calls:
        for mat in (X, D):
            if label == "gsva" and mat is X:
                continue
            M = pb.NamedMatrix(mat, names)
            whole = f(ctxs[0], M).mat
            for world in (2, 3):
                got = f(ctxs[:world], M).mat
                if label == "gsva":  # row sums are added per shard: z agrees to the last bits only
                    assert rel_err(got, whole) < tol(1e-9), (label, world)
                else:
                    assert np.array_equal(got, whole), (label, world)
    # pageable destination (numpy malloc) vs pinned destination: same bits, several ring blocks (64 MB each)
    Nbig = 5000
    Xb = synth.sparse_x_numpy(P, Nbig, seed=96)
    a = pb.plaid(pb.NamedMatrix(Xb, names), Gn, ctx=ctxs[0]).mat                   # pageable: ring path (72 MB)
    pinned = torch.empty(S * Nbig, dtype=torch.float64).pin_memory()
    b = pb.plaid(pb.NamedMatrix(Xb, names), Gn, ctx=ctxs[0], out=pinned.numpy().reshape((S, Nbig), order="F")).mat
    assert np.array_equal(a, np.asarray(b))
    monkeypatch.setenv("PLAIDGPU_NO_RING", "1")
    c2 = pb.plaid(pb.NamedMatrix(Xb, names), Gn, ctx=ctxs[0]).mat
    assert np.array_equal(a, c2)


def test_early_shipping_to_pinned_host_output(monkeypatch):
    """Pinned host destination: column chunks leave as RAW scores while later chunks are still being scored, and
    plaidgpu_score_finish fixes them up on the host (api.cu host_fixup_column = k_fixup operation by operation).
    The result must equal the ordinary path bit for bit: plaid() normalised (host fix-up with medians),
    replaid.ucell (alpha / beta fix-up), replaid.sing (no normalisation: every chunk is final when it leaves)."""
    import torch
    from plaid_b200 import api
    P, N, S = 1500, 4000, 1800
    X = synth.sparse_x_numpy(P, N, seed=97)
    G = synth.genesets_numpy(P, S, seed=98, size_cap=(5, 200))
    names = synth.gene_names(P)
    Xn, Gn = pb.NamedMatrix(X, names), pb.NamedMatrix(G, names)
    ctx = pb.Context(0)
    monkeypatch.setenv("PLAIDGPU_TAIL_TILES", "1")   # chunks of 1,056 columns
    monkeypatch.setenv("PLAIDGPU_H2D_PIECE", "20000")  # plaid(): X crosses PCIe in ~20 pieces, chunks wait for theirs only
    calls = [lambda **kw: pb.plaid(Xn, Gn, ctx=ctx, **kw), lambda **kw: pb.replaid_ucell(Xn, Gn, rmax=200, ctx=ctx, **kw),
             lambda **kw: pb.replaid_sing(Xn, Gn, ctx=ctx, **kw)]
    for f in calls:
        monkeypatch.setenv("PLAIDGPU_NO_EARLY", "1")
        want = f().mat.copy()
        monkeypatch.delenv("PLAIDGPU_NO_EARLY")
        monkeypatch.setenv("PLAIDGPU_EARLY_FRAC", "0.6")  # two of the four chunks leave early
        pinned = torch.empty(S * N, dtype=torch.float64).pin_memory()
        got = f(out=pinned.numpy().reshape((S, N), order="F")).mat
        assert np.array_equal(np.asarray(got), want)
        monkeypatch.delenv("PLAIDGPU_EARLY_FRAC")
    ctx.close()


def test_median_choice_across_shards():
    """normalize_medians uses ONE median, picked by min(x) == 0 over ALL shards (R/plaid.R:556-557); each shard
    computes up front only the median its own minimum predicts and the other one on demand.  Long columns
    (S >= 4096: the single-pass statistics kernel) in every constellation: zeros everywhere; zeros in shard 1 and
    negatives in shard 2 (shard 1 predicted wrong); no zeros at all; negatives everywhere; explicit ignore.zero."""
    from plaid_b200 import _lib as L, sharded
    from plaid_b200.api import _matrix_struct, _opts
    P, N, S = 1500, 64, 4300
    names = synth.gene_names(P)
    G = synth.genesets_numpy(P, S, seed=52, size_cap=(5, 120))
    rowmap = pb.make_rowmap(names, names)
    Xs = synth.sparse_x_numpy(P, N, seed=51)                      # scores >= 0 with exact zeros
    Xneg = Xs.copy().tolil()
    Xneg[:, N // 2:] = -Xs[:, N // 2:].toarray()                  # second half: negative scores
    Xneg = Xneg.tocsc()
    Dpos = np.abs(synth.dense_x_numpy(P, N, seed=53)) + 0.5       # every score > 0
    Dall = -Dpos
    ctxs = [pb.Context(0) for _ in range(2)]
    for c in ctxs:
        c.set_genesets(G)
    Go = O.Named(G, names, [f"s{k}" for k in range(S)])
    cases = [("zeros", Xs, -1, True), ("zeros+negatives", Xneg, -1, False), ("positive", Dpos, -1, False),
             ("negative", Dall, -1, False), ("forced on", Xneg, 1, True), ("forced off", Xs, 0, False)]
    for label, X, izopt, iz_expected in cases:
        keep = []
        whole = np.empty((S, N), order="F")
        o = _opts(ctxs[0].lib, scorer=L.PLAID, out_location=L.HOST, normalize=1, ignore_zero=izopt)
        ctxs[0].check(ctxs[0].lib.plaidgpu_score(ctxs[0].h, _matrix_struct(X, keep), rowmap.ctypes.data, o, whole.ctypes.data))
        raw = O.plaid(O.Named(X, names, [f"c{k}" for k in range(N)]), Go, normalize=False).mat
        want = O.normalize_medians(raw, ignore_zero=None if izopt < 0 else bool(izopt))
        assert rel_err(whole, want) < tol(1e-11), label
        spans = [sharded.shard_columns(N, 2, r) for r in range(2)]
        outs = [np.empty((S, hi - lo), order="F") for lo, hi in spans]
        mats = [_matrix_struct(X[:, lo:hi], keep) for lo, hi in spans]
        opts = [_opts(ctxs[0].lib, scorer=L.PLAID, out_location=L.HOST, normalize=1, ignore_zero=izopt) for _ in spans]
        scal = sharded.score_multi(ctxs, mats, rowmap, opts, [a.ctypes.data for a in outs], [hi - lo for lo, hi in spans])
        assert [bool(s.ignore_zero) for s in scal] == [iz_expected] * 2, label
        assert np.array_equal(np.concatenate(outs, axis=1), whole), label
        # the full pair is still available on request (computed on demand), and agrees with numpy
        ma, mz = np.empty(spans[0][1]), np.empty(spans[0][1])
        tmp = np.empty((S, spans[0][1]), order="F")
        loc = L.Scalars()
        c0 = ctxs[0]
        c0.check(c0.lib.plaidgpu_score_begin(c0.h, mats[0], rowmap.ctypes.data, opts[0], loc))
        c0.check(c0.lib.plaidgpu_score_compute(c0.h, loc, tmp.ctypes.data))
        c0.check(c0.lib.plaidgpu_get_col_medians(c0.h, ma.ctypes.data, mz.ctypes.data))
        r0 = raw[:, :spans[0][1]]
        assert np.allclose(ma, np.median(r0, axis=0), rtol=tol(1e-12), atol=0), label
        z = np.where(r0 == 0, np.nan, r0)
        with np.errstate(all="ignore"):
            mzo = np.nanmedian(z, axis=0)
        assert np.allclose(mz, np.where(np.isnan(mzo), 0.0, mzo), rtol=tol(1e-12), atol=0), label


def test_gsva_on_column_shards():
    """replaid.gsva on 2 and 3 ragged column shards (threads, one context each): rowtf "ecdf" re-partitions
    the dense shards into row blocks (all-to-all), ranks each gene across ALL samples (plaidgpu_row_ecdf)
    and sends them back -> bit-identical to the one-shard call; rowtf "z" all-reduces the row sums
    (R/plaid.R:343-346)."""
    import threading
    from plaid_b200 import _lib as L, sharded
    from plaid_b200.api import _opts
    P, N, S = 900, 53, 700
    Draw = synth.dense_x_numpy(P, N, seed=61)
    Dtie = np.round(Draw, 1)  # ties across samples and genes; only for ecdf: with rounded data some values EQUAL
    # their row mean, z is then +-1e-15 noise whose sign (hence signed rank) depends on the summation order
    G = synth.genesets_numpy(P, S, seed=62, size_cap=(5, 150))
    names = synth.gene_names(P)
    rowmap = pb.make_rowmap(names, names)
    ctxs = [pb.Context(0) for _ in range(3)]
    for c in ctxs:
        c.set_genesets(G)
    Go = O.Named(G, names, [f"s{k}" for k in range(S)])
    for rowtf, tau in (("ecdf", 0.0), ("z", 0.0), ("ecdf", 0.5)):
        D = Dtie if rowtf == "ecdf" else Draw
        Xo = O.Named(D, names, [f"c{k}" for k in range(N)])
        whole = pb.replaid_gsva(pb.NamedMatrix(D, names), pb.NamedMatrix(G, names), tau=tau, rowtf=rowtf, ctx=ctxs[0]).mat
        assert rel_err(whole, O.replaid_gsva(Xo, Go, tau=tau, rowtf=rowtf).mat) < tol(1e-9)
        for world in (2, 3):
            comms = sharded.ThreadComm.group(world)
            spans = [sharded.shard_columns(N, world, r) for r in range(world)]
            outs = [np.empty((S, hi - lo), order="F") for lo, hi in spans]
            errs = []

            def run(r):
                try:
                    lo, hi = spans[r]
                    o = _opts(ctxs[r].lib, scorer=L.GSVA, out_location=L.HOST, tau=tau)
                    sharded.gsva_shard(ctxs[r], comms[r], D[:, lo:hi], rowmap, o, outs[r].ctypes.data, rowtf=rowtf)
                except Exception as e:  # pragma: no cover
                    errs.append(e)
                    comms[r]._s["barrier"].abort()

            ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
            for t in ts:
                t.start()
            for t in ts:
                t.join(timeout=120)
            assert not errs, errs
            got = np.concatenate(outs, axis=1)
            if rowtf == "ecdf":
                assert np.array_equal(got, whole)
            else:  # row sums are added per shard: last-bit differences in z may move a rank by a tie
                assert rel_err(got, whole) < tol(1e-9), (rowtf, tau, world, rel_err(got, whole), float(np.abs(got - whole).max()))


def test_column_chunked_host_path_is_bit_identical(monkeypatch):
    """outputs larger than the device budget are scored in column chunks (two passes when normalised);
    forced here with a tiny budget: results must equal the one-pass results bit for bit"""
    P, N, S = 1500, 203, 1100
    X = synth.sparse_x_numpy(P, N, seed=81)
    D = synth.dense_x_numpy(P, 67, seed=82)
    G = synth.genesets_numpy(P, S, seed=83, size_cap=(5, 200))
    names = synth.gene_names(P)
    Gn = pb.NamedMatrix(G, names)
    ctx = pb.Context(0)
    calls = [lambda c: pb.plaid(pb.NamedMatrix(X, names), Gn, ctx=c).mat,
             lambda c: pb.plaid(pb.NamedMatrix(X, names), Gn, stats="sum", normalize=False, ctx=c).mat,
             lambda c: pb.plaid(pb.NamedMatrix(D, names), Gn, ctx=c).mat,
             lambda c: pb.replaid_ucell(pb.NamedMatrix(X, names), Gn, rmax=200, ctx=c).mat,
             lambda c: pb.replaid_ssgsea(pb.NamedMatrix(X, names), Gn, alpha=0.25, ctx=c).mat,
             lambda c: pb.replaid_scse(pb.NamedMatrix(X, names), Gn, ctx=c).mat]
    whole = [f(ctx) for f in calls]
    monkeypatch.setenv("PLAIDGPU_MAX_OUT_BYTES", str(S * 8 * 40))  # 40 columns per chunk (rounded to 32)
    parts = [f(ctx) for f in calls]
    for a, b in zip(whole, parts):
        assert np.array_equal(a, b)
    assert rel_err(whole[0], O.plaid(O.Named(X, names), O.Named(G, names)).mat) < tol(TOL)


# ---- size-independent properties at larger sizes (no oracle needed) ---------------------------------
def test_properties_linearity_and_column_independence(gpu_ctx):
    P, N, S = 20000, 512, 30000
    X = synth.sparse_x_numpy(P, N, seed=71)
    G = synth.genesets_numpy(P, 2000, seed=72)
    G = sp.hstack([G] * 15).tocsc()[:, :S]  # 30k sets (repeated blocks), several tiles
    names = synth.gene_names(P)
    Gn = pb.NamedMatrix(G, names)
    a = pb.plaid(pb.NamedMatrix(X, names), Gn, stats="sum", normalize=False, ctx=gpu_ctx).mat
    assert a.shape == (S, N)
    # repeated set blocks give the same rows (the order of the fp64 adds may differ between tiles)
    assert np.allclose(a[:2000], a[2000:4000], rtol=1e-13, atol=0)
    # column independence: scoring a column subset gives the same bits (sharding invariant)
    sub = pb.plaid(pb.NamedMatrix(X[:, 100:228], names), Gn, stats="sum", normalize=False, ctx=gpu_ctx).mat
    assert np.array_equal(sub, a[:, 100:228])
    # linearity in X (power-of-two scaling is exact in fp64)
    b = pb.plaid(pb.NamedMatrix(X * 4.0, names), Gn, stats="sum", normalize=False, ctx=gpu_ctx).mat
    assert np.array_equal(b, 4.0 * a)
    # sum of all set scores == sum over genes of x * degree (checksum of checksums)
    deg = np.asarray(G.sum(1)).ravel()
    chk = np.asarray(X.T @ deg).ravel()
    assert np.allclose(a.sum(0), chk, rtol=tol(1e-12))
    # median-normalised output: every column has the same median (mean of medians), zeros ignored
    n = pb.plaid(pb.NamedMatrix(X, names), Gn, ctx=gpu_ctx).mat
    raw = pb.plaid(pb.NamedMatrix(X, names), Gn, normalize=False, ctx=gpu_ctx).mat
    z = raw.copy()
    z[z == 0] = np.nan
    med = np.nanmedian(z, axis=0)
    assert np.allclose(n, raw - med[None, :] + med.mean(), rtol=0, atol=1e-12)


def test_score_to_file_is_bit_identical(tmp_path, monkeypatch):
    """tiled egress (scope row f4): the result streamed to a .npy / raw file in column tiles equals the
    in-memory result bit for bit, one tile or many (forced with a small tile budget), normalised or not"""
    from plaid_b200 import _lib as L
    P, N, S = 1300, 157, 900
    X = synth.sparse_x_numpy(P, N, seed=71)
    G = synth.genesets_numpy(P, S, seed=72, size_cap=(5, 150))
    names = synth.gene_names(P)
    Xn, Gn = pb.NamedMatrix(X, names), pb.NamedMatrix(G, names)
    ctx = pb.Context(0)
    cases = [(dict(), lambda: pb.plaid(Xn, Gn, ctx=ctx).mat),
             (dict(normalize=0, stats_mean=0), lambda: pb.plaid(Xn, Gn, stats="sum", normalize=False, ctx=ctx).mat),
             (dict(scorer=L.UCELL, rmax=120.0), lambda: pb.replaid_ucell(Xn, Gn, rmax=120, ctx=ctx).mat)]
    whole = [f() for _, f in cases]
    for budget in (None, S * 8 * 40):  # one tile; 40-column tiles (32 after rounding) -> 5 tiles, ragged tail
        if budget is not None:
            monkeypatch.setenv("PLAIDGPU_MAX_OUT_BYTES", str(budget))
        for (kw, _), ref in zip(cases, whole):
            path = str(tmp_path / "scores.npy")
            assert pb.score_to_file(Xn, Gn, path, ctx=ctx, **kw) == (S, N)
            got = np.load(path, mmap_mode="r")
            assert got.shape == (S, N) and got.flags.f_contiguous
            assert np.array_equal(got, ref)
        raw = str(tmp_path / "scores.bin")
        pb.score_to_file(Xn, Gn, raw, fmt="raw", ctx=ctx)
        assert np.array_equal(np.fromfile(raw).reshape((S, N), order="F"), whole[0])
    monkeypatch.delenv("PLAIDGPU_MAX_OUT_BYTES")
    with pytest.raises(L.PlaidGpuError, match="cannot create"):
        pb.score_to_file(Xn, Gn, str(tmp_path / "no_such_dir" / "x.npy"), ctx=ctx)


def test_full_c4_shard_size_properties():
    """BASELINE config C4, one GPU's shard (20,000 genes x 125,000 cells x 30,000 sets, 30 GB of scores,
    inputs generated in HBM like bench.py): the oracle cannot run at this size, so the result is checked through
    size-independent properties — column independence (bit-exact re-scoring of sampled columns on their own),
    the degree checksum of every column, exact medians of every column (device sort), and the normalisation
    identity out = raw - med_j + mean(med)."""
    import torch
    from plaid_b200 import _lib as L
    P, N, S = 20000, 125000, 30000
    dev = "cuda:0"
    Gp, Gi = synth.genesets_torch(P, S, seed=synth.SEED0 + 3, device=dev)
    G = sp.csc_matrix((np.ones(Gi.size), Gi, Gp), shape=(P, S))
    xp, xi, xx = synth.sparse_x_torch(P, N, seed=synth.SEED0 + 1003, device=dev)
    names = synth.gene_names(P)
    Gn = pb.NamedMatrix(G, names)
    ctx = pb.Context(0)
    Xd = pb.NamedMatrix(pb.DeviceCSC(xp, xi, xx, (P, N)), names)
    raw = torch.empty(S * N, dtype=torch.float64, device=dev)
    out = torch.empty(S * N, dtype=torch.float64, device=dev)
    assert pb.plaid(Xd, Gn, normalize=False, ctx=ctx, out=raw) is not None
    assert pb.plaid(Xd, Gn, ctx=ctx, out=out) is not None
    raw2, out2 = raw.view(N, S), out.view(N, S)  # row j = column j of the S x N column-major result

    # checksum of checksums: sum_s raw[s, j] * n_s == sum_g x[g, j] * degree(g)   (raw is the set MEAN)
    ns = torch.from_numpy(np.asarray(G.sum(0)).ravel()).to(dev)
    deg = torch.from_numpy(np.asarray(G.sum(1)).ravel()).to(dev)
    colid = torch.repeat_interleave(torch.arange(N, device=dev), (xp[1:] - xp[:-1]).long())
    want = torch.zeros(N, dtype=torch.float64, device=dev).index_add_(0, colid, xx * deg[xi.long()])
    got = torch.empty(N, dtype=torch.float64, device=dev)
    for j0 in range(0, N, 8192):
        got[j0:j0 + 8192] = raw2[j0:j0 + 8192] @ (ns + 1e-8)
    assert torch.allclose(got, want, rtol=tol(1e-11), atol=0)

    # exact medians of every column (zeros dropped: min(raw) == 0 here), by a device sort
    assert float(raw.min()) == 0.0
    med = torch.empty(N, dtype=torch.float64, device=dev)
    for j0 in range(0, N, 4096):
        blk = raw2[j0:j0 + 4096]
        srt = torch.where(blk == 0, torch.full_like(blk, float("inf")), blk).sort(dim=1).values
        cnt = (blk != 0).sum(dim=1)
        lo = srt.gather(1, ((cnt - 1) // 2).clamp(min=0)[:, None])[:, 0]
        hi = srt.gather(1, (cnt // 2).clamp(max=S - 1)[:, None])[:, 0]
        med[j0:j0 + 4096] = torch.where(cnt > 0, (lo + hi) / 2, torch.zeros_like(lo))
    c = float(med.mean())
    worst = 0.0
    for j0 in range(0, N, 8192):
        worst = max(worst, float((out2[j0:j0 + 8192] - (raw2[j0:j0 + 8192] - med[j0:j0 + 8192, None] + c)).abs().max()))
    assert worst < 1e-12  # mean(med): R sums in long double, torch in fp64 pairwise

    # column independence at full size: sampled columns re-scored alone give the same bits
    rng = np.random.default_rng(5)
    cols = np.sort(rng.choice(N, size=96, replace=False))
    p_h = xp.cpu().numpy()
    parts_i, parts_x, pp = [], [], [0]
    for j in cols:
        parts_i.append(xi[p_h[j]:p_h[j + 1]])
        parts_x.append(xx[p_h[j]:p_h[j + 1]])
        pp.append(pp[-1] + int(p_h[j + 1] - p_h[j]))
    sub = pb.NamedMatrix(pb.DeviceCSC(torch.tensor(pp, dtype=torch.int32, device=dev), torch.cat(parts_i), torch.cat(parts_x),
                                      (P, len(cols))), names)
    sub_out = torch.empty(S * len(cols), dtype=torch.float64, device=dev)
    pb.plaid(sub, Gn, normalize=False, ctx=ctx, out=sub_out)
    assert torch.equal(sub_out.view(len(cols), S), raw2[torch.from_numpy(cols).to(dev)])
    # the same for a rank scorer (C5-shaped): replaid.sing has no cross-column scalar, so the sampled
    # columns — ranked and scored alone — must again give the same bits
    assert pb.replaid_sing(Xd, Gn, ctx=ctx, out=out) is not None
    pb.replaid_sing(sub, Gn, ctx=ctx, out=sub_out)
    assert torch.equal(sub_out.view(len(cols), S), out2[torch.from_numpy(cols).to(dev)])
    assert float(out.min()) >= -0.5 and float(out.max()) <= 0.5  # r / nrow(X) - 0.5

    # ---- the ORACLE at the benchmarked shape: the 96 sampled columns of this very shard, scored on the CPU with
    # the restated reference (R/plaid.R:60-87, 213-309) and compared with the columns of the full-size runs above
    # (same plan: 27 scatter tiles of 1,120 sets, tensor-core block / 704-row gather block)
    Xs = sp.csc_matrix((torch.cat(parts_x).cpu().numpy(), torch.cat(parts_i).cpu().numpy(), np.asarray(pp, dtype=np.int32)),
                       shape=(P, len(cols)))
    Xo, Go = O.Named(Xs, names), O.Named(G, names)
    cidx = torch.from_numpy(cols).to(dev)
    o_raw = O.plaid(Xo, Go, normalize=False).mat                         # S x 96
    assert rel_err(raw2[cidx].cpu().numpy().T, o_raw) < tol(1e-11)
    # normalised: medians / mean(med) are global over the 125,000 columns -> taken from the full GPU run, whose
    # medians are themselves checked against the oracle's own column medians here
    z = o_raw.copy()
    z[z == 0] = np.nan
    o_med = np.nanmedian(z, axis=0)
    assert np.allclose(med[cidx].cpu().numpy(), o_med, rtol=tol(1e-12), atol=0)
    pb.plaid(Xd, Gn, ctx=ctx, out=out)
    o_norm = o_raw - med[cidx].cpu().numpy()[None, :] + c
    assert rel_err(out2[cidx].cpu().numpy().T, o_norm) < tol(1e-11)
    # replaid.sing of the full shard (no cross-column scalar), and the scorers whose transform uses the global
    # max rank on the sampled columns scored on their own (the oracle sees the same 96 columns)
    pb.replaid_sing(Xd, Gn, ctx=ctx, out=out)
    assert rel_err(out2[cidx].cpu().numpy().T, O.replaid_sing(Xo, Go).mat) < tol(1e-11)
    for fn, ofn in ((pb.replaid_ssgsea, O.replaid_ssgsea), (pb.replaid_ucell, O.replaid_ucell),
                    (pb.replaid_aucell, O.replaid_aucell)):
        fn(sub, Gn, ctx=ctx, out=sub_out)
        assert rel_err(sub_out.view(len(cols), S).cpu().numpy().T, ofn(Xo, Go).mat) < tol(1e-11), fn.__name__
    del raw, out
    torch.cuda.empty_cache()
