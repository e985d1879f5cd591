"""CPU tests of the oracle (test infrastructure): the restated solvers against independent
implementations / optimality conditions, the model glue against the reference's own Cython
(oracle/_ref, when built) and against the committed golden fixtures it produced."""
import os

import numpy as np
import pytest
import scipy.optimize

from amico_b200 import synth
from oracle import oracle as orc
from oracle import ref_runner

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_nnls_matches_scipy_lawson_hanson():
    rng = np.random.default_rng(0)
    for m, n in ((20, 8), (50, 30), (33, 11), (10, 25)):
        for _ in range(20):
            A = rng.uniform(0, 1, (m, n))
            y = rng.uniform(0, 1, m)
            x, rn = orc.nnls(A, y)
            xs, rs = scipy.optimize.nnls(A, y)
            assert np.allclose(x, xs, atol=1e-9)
            assert abs(rn - rs) < 1e-9


def test_nnls_kkt_on_noddi_dictionary():
    P = synth.make_problem(2, n_vox=40)
    lut = synth.lut_index_numpy(P.DIRs, P.htable)
    for i in range(40):
        A = synth.dictionary_for_direction("NODDI", P.KERNELS, int(lut[i]))
        y = P.y[i].astype(np.float64)
        x, _ = orc.nnls(A, y)
        w = A.T @ (y - A @ x)
        assert (x >= 0).all()
        assert w.max() < 1e-8            # dual feasibility
        assert np.abs(w[x > 0]).max() < 1e-8  # complementary slackness


@pytest.mark.parametrize("l1,l2", [(0.0, 1e-3), (0.5, 1e-3), (0.0, 4.0), (0.2, 0.0)])
def test_lasso_kkt(l1, l2):
    rng = np.random.default_rng(1)
    for m, n in ((33, 11), (90, 40), (300, 26)):
        A = rng.uniform(0, 1, (m, n))
        y = rng.uniform(0, 1, m)
        x = orc.lasso(A, y, l1, l2)
        g = A.T @ (y - A @ x) - max(l2, 1e-10) * x - l1
        assert (x >= 0).all()
        assert g[x == 0].max(initial=-1) < 1e-7
        if (x > 0).any():
            assert np.abs(g[x > 0]).max() < 1e-7


def test_lasso_equals_augmented_nnls():
    """enet(A, y, 0, l2) == nnls([A; sqrt(l2) I], [y; 0])."""
    rng = np.random.default_rng(2)
    A = rng.uniform(0, 1, (40, 12))
    y = rng.uniform(0, 1, 40)
    l2 = 0.05
    x = orc.lasso(A, y, 0.0, l2)
    xa, _ = orc.nnls(np.vstack([A, np.sqrt(l2) * np.eye(12)]), np.concatenate([y, np.zeros(12)]))
    assert np.allclose(x, xa, atol=1e-9)


def test_lasso_support_cap_when_m_lt_n():
    """SPAMS truncates the path at L = min(m, n) atoms (binds for SANDI: m = 4 < n = 15)."""
    P = synth.make_problem(4, n_vox=200)
    A = synth.dictionary_for_direction("SANDI", P.KERNELS, 0)
    for i in range(200):
        x = orc.lasso(A, P.y[i].astype(np.float64), 0.0, 5e-3)
        assert (x > 0).sum() <= A.shape[0]


def test_lut_index_against_numpy_restatement_and_edges():
    P = synth.make_problem(1, n_vox=4)
    rng = np.random.default_rng(3)
    v = rng.standard_normal((200000, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    edge = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1], [0, -0.0, 1], [0, 0, 0]], dtype=np.float64)
    d = np.vstack([edge, v])
    assert np.array_equal(orc.lut_indices(d, P.htable), synth.lut_index_numpy(d, P.htable))
    bad = np.array([[np.nan, 0.1, 0.2], [0.0, 1.0, 0.0]])
    assert orc.lut_indices(bad, P.htable)[0] == -1
    with pytest.raises(RuntimeError, match="index out of bounds"):
        orc.fit("FreeWater", P.y[:2], bad, P.htable, P.KERNELS, P.params, 0.0, 1e-3)


def test_htable_semantics_match_reference_table():
    ref_dir = "/root/reference/amico/directions"
    if not os.path.isdir(ref_dir):
        pytest.skip("reference not mounted")
    dirs = np.fromfile(os.path.join(ref_dir, "ndirs=500.bin"), dtype=np.float64).reshape(-1, 3)
    ht = np.fromfile(os.path.join(ref_dir, "htable_ndirs=500.bin"), dtype=np.int16)
    mine = synth.build_htable(dirs)
    # nearest-direction semantics: identical except at numerical ties between two equally near directions
    assert (mine == ht).mean() > 0.995
    ang = np.deg2rad(np.arange(181.0))
    T, Pp = np.meshgrid(ang, ang, indexing="ij")
    v = np.stack([np.sin(T) * np.cos(Pp), np.sin(T) * np.sin(Pp), np.cos(T)], -1).reshape(-1, 3)
    dots = np.abs(v @ dirs.T)
    diff = np.nonzero(mine != ht)[0]
    assert np.allclose(dots[diff, mine[diff]], dots[diff, ht[diff]], atol=1e-6)


CASES = [("freewater_cfg1", 1, "FreeWater", 512, None), ("freewater_mouse", 1, "FreeWaterMouse", 256, 11),
         ("noddi_cfg2", 2, "NODDI", 384, None), ("sandi_cfg4", 4, "SANDI", 512, None), ("czb_cfg5", 5, "CylinderZeppelinBall", 320, None)]


@pytest.mark.parametrize("name,cfg,model,n_vox,seed", CASES)
def test_oracle_reproduces_golden_reference_outputs(name, cfg, model, n_vox, seed):
    """The fixtures are outputs of daducci/AMICO's unmodified Cython fit (tests/golden/make_golden.py)."""
    import hashlib
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    P = synth.make_problem(cfg, n_vox=n_vox, model=model, seed=seed)
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(P.y).tobytes())
    if P.DIRs is not None:
        h.update(np.ascontiguousarray(P.DIRs).tobytes())
    for k in sorted(P.KERNELS):
        if k != "model":
            h.update(np.ascontiguousarray(P.KERNELS[k]).tobytes())
    if h.hexdigest() != str(g["input_sha256"]):
        pytest.skip("synthetic generator is not bit-reproducible on this host (libm differences): fixture inputs differ")
    res = orc.fit_problem(P, rmse=True, nrmse=True, extra=model in ("NODDI", "FreeWater", "FreeWaterMouse"), nthreads=2)
    for k in ("estimates", "rmse", "nrmse", "estimates_mod", "y_corrected"):
        if k in g.files:
            assert np.array_equal(res[k], g[k]), k


@pytest.mark.parametrize("cfg,model", [(1, "FreeWater"), (2, "NODDI"), (4, "SANDI"), (5, "CylinderZeppelinBall")])
def test_oracle_equals_reference_glue(cfg, model):
    if not ref_runner.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    P = synth.make_problem(cfg, n_vox=640, model=model, seed=77)
    a = orc.fit_problem(P, rmse=True, nrmse=True, nthreads=3)
    b = ref_runner.fit_problem(P, nthreads=2, rmse=True, nrmse=True)
    for k in ("estimates", "rmse", "nrmse"):
        assert np.array_equal(a[k], b[k]), k


def test_chunking_is_invisible():
    P = synth.make_problem(2, n_vox=101)
    a = orc.fit_problem(P, nthreads=1)["estimates"]
    b = orc.fit_problem(P, nthreads=7)["estimates"]
    assert np.array_equal(a, b)


# ----------------------------------------------------------------------------------------------- the reference's own LUT tables
def _refdirs_problem(ndirs=500):
    g = np.load(os.path.join(GOLDEN, f"noddi_refdirs{ndirs}.npz"))
    return g, synth.make_problem(2, n_vox=384, seed=77, lut_dirs=g["lut_dirs"], htable=g["htable"])


def test_reference_hash_table_semantics():
    """amico/directions/htable_ndirs=500.bin (fixture): entry [theta_deg * 181 + phi_deg] is the LUT direction nearest (max |dot|)
    to that whole-degree direction -- exactly what synth.build_htable computes, so synthetic tables have the real semantics."""
    g = np.load(os.path.join(GOLDEN, "noddi_refdirs500.npz"))
    dirs, ht = g["lut_dirs"], g["htable"]
    assert dirs.shape == (500, 3) and ht.shape == (181 * 181,) and ht.dtype == np.int16
    np.testing.assert_allclose(np.linalg.norm(dirs, axis=1), 1.0, atol=1e-12)
    assert (dirs[:, 1] >= 0).all()                      # half sphere y >= 0: why dir_to_lut_idx flips (amico/lut.pyx:335-338)
    assert np.array_equal(synth.build_htable(dirs), ht)


@pytest.mark.parametrize("ndirs", [500, 1, 5000])
def test_oracle_reproduces_golden_on_reference_direction_set(ndirs):
    """NODDI on the reference's REAL direction sets and hash tables (500, 1 and 5000 directions): outputs of the reference's
    Cython fit."""
    import hashlib
    g, P = _refdirs_problem(ndirs)
    assert g["lut_dirs"].shape == (ndirs, 3)
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(P.y).tobytes())
    h.update(np.ascontiguousarray(P.DIRs).tobytes())
    for k in sorted(P.KERNELS):
        if k != "model":
            h.update(np.ascontiguousarray(P.KERNELS[k]).tobytes())
    if h.hexdigest() != str(g["input_sha256"]):
        pytest.skip("synthetic generator is not bit-reproducible on this host (libm differences): fixture inputs differ")
    res = orc.fit_problem(P, rmse=True, nthreads=2)
    assert np.array_equal(res["estimates"], g["estimates"]) and np.array_equal(res["rmse"], g["rmse"])
