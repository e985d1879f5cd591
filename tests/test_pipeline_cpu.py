"""CPU checks of the pre-processing / DTI oracle (oracle/pipeline.py) -- the checker the GPU pipeline tests use."""
import numpy as np
import pytest

from amico_b200 import synth
from amico_b200.evaluation import dti_design_matrix
from oracle import pipeline as opl


def test_preprocess_oracle_hand_values():
    sch = synth.Scheme(np.array([[0, 0, 0, 0], [1, 0, 0, 1000], [0, 0, 0, 0], [0, 1, 0, 1000]], dtype=float))
    dwi = np.zeros((1, 1, 3, 4), dtype=np.float32)
    dwi[0, 0, 0] = [100, 50, 300, -20]   # mean b0 = 200
    dwi[0, 0, 1] = [0, 5, 0, 7]          # mean b0 = 0 -> norm factor 0
    dwi[0, 0, 2] = [10, 5, 30, 2]
    mask = np.array([[[1, 1, 0]]], dtype=np.uint8)
    r = opl.preprocess(dwi, sch, mask)
    assert r["y"].shape == (2, 4)
    np.testing.assert_array_equal(r["vox_idx"], [0, 1])
    np.testing.assert_allclose(r["y"][0], [0.5, 0.25, 1.5, 0.0])  # negative value clamped
    np.testing.assert_array_equal(r["y"][1], 0.0)
    np.testing.assert_allclose(r["mean_b0s"].ravel(), [200, 0, 20])
    m = opl.preprocess(dwi, sch, mask, doMergeB0=True)
    np.testing.assert_allclose(m["y"][0], [1.0, 0.25, 0.0])
    with pytest.raises(FloatingPointError):
        bad = dwi.copy()
        bad[0, 0, 2, 1] = np.nan
        opl.preprocess(bad, sch, mask)
    ok = opl.preprocess(bad, sch, None, replace_bad_voxels=0)
    assert np.isfinite(ok["y"]).all()


def test_preprocess_oracle_mean_is_sequential_float32():
    # the GPU kernel relies on this: numpy reduces the fancy-indexed copy sequentially in index order, in float32
    rng = np.random.default_rng(3)
    P, dwi, mask = synth.make_raw_volume(2, (4, 3, 5), seed=1)
    sch = P.scheme
    r = opl.preprocess(dwi, sch, mask, doNormalizeSignal=True)
    flat = dwi.reshape(-1, sch.nS)
    s = flat[:, sch.b0_idx[0]].copy()
    for i in sch.b0_idx[1:]:
        s = (s + flat[:, i]).astype(np.float32)
    mb = (s / np.float32(sch.b0_count)).astype(np.float32)
    np.testing.assert_array_equal(r["mean_b0s"].ravel(), mb)
    nf = (np.float32(1) / mb).astype(np.float32)
    want = (flat * nf[:, None]).astype(np.float32)[mask.ravel() == 1].astype(np.float64)
    want[want < 0] = 0
    np.testing.assert_array_equal(r["y"], want)
    del rng


def test_directional_average_oracle_aliasing():
    # dir_avg_img is a view of the first n_shells+1 volumes: a shell that contains volume 0..n_shells reads the
    # already averaged value (reference behaviour, amico/core.py:234-251)
    tab = np.array([[1, 0, 0, 1000], [0, 0, 0, 0], [0, 1, 0, 1000], [0, 0, 1, 2000], [1, 0, 0, 2000]], dtype=float)
    sch = synth.Scheme(tab)
    dwi = np.array([[[[8, 4, 2, 6, 10]]]], dtype=np.float32)
    r = opl.preprocess(dwi, sch, None, doNormalizeSignal=False, doDirectionalAverage=True)
    # b0 mean = 4 -> volume 0 := 4; shell b=1000 = mean(vol0 (now 4), vol2 = 2) = 3; shell b=2000 = mean(6, 10) = 8
    np.testing.assert_array_equal(r["y"], [[4.0, 3.0, 8.0]])


def test_dti_oracle_recovers_tensor_direction():
    rng = np.random.default_rng(0)
    sch = synth.make_scheme(2)
    n = 200
    v = rng.standard_normal((n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    g = sch.raw[:, :3]
    b = sch.b
    lam1, lam2 = 1.7e-3, 0.3e-3
    adc = lam2 + (lam1 - lam2) * (g @ v.T) ** 2          # (nS, n)
    y = np.exp(-b[:, None] * adc).T
    d = opl.dti_directions(y, sch)
    assert np.abs(np.abs((d * v).sum(1)) - 1).max() < 1e-9
    # merged-b0 gradient table (core.py:432-433)
    ym = np.concatenate([y[:, sch.b0_idx].mean(1, keepdims=True), y[:, sch.dwi_idx]], axis=1)
    dm = opl.dti_directions(ym, sch, doMergeB0=True)
    assert np.abs(np.abs((dm * v).sum(1)) - 1).max() < 1e-9


def test_design_matrix_matches_oracle():
    sch = synth.make_scheme(5)
    np.testing.assert_array_equal(dti_design_matrix(sch.b, sch.raw[:, :3]), opl.dti_design_matrix(sch.b, sch.raw[:, :3]))


def test_scatter_oracle():
    vol = opl.scatter_maps(np.array([[1.0, 2.0], [3.0, 4.0]]), [4, 1], 6)
    assert vol.dtype == np.float32 and vol.shape == (6, 2)
    np.testing.assert_array_equal(vol[[1, 4]], [[3, 4], [1, 2]])
    assert vol.sum() == 10


def test_real_sh_basis_addition_theorem():
    """The SH basis restated from dipy's definition: sum_m Y_lm(g) Y_lm(d) = (2l+1)/(4 pi) P_l(g.d) for every even l <= 12
    (holds for any orthonormal real basis of each order), and the m = 0 column is the zonal harmonic the reference's
    rotation relies on (amico/lut.pyx:129-138: idx_m0 = (l*l + l + 2)/2 - 1)."""
    from numpy.polynomial import legendre as npleg
    from amico_b200 import lut
    rng = np.random.default_rng(0)
    g = rng.standard_normal((40, 3)); g /= np.linalg.norm(g, axis=1, keepdims=True)
    d = rng.standard_normal((9, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    Yg, m, l = lut.real_sh_descoteaux(12, *lut.cart2sphere(*g.T)[1:])
    Yd, _, _ = lut.real_sh_descoteaux(12, *lut.cart2sphere(*d.T)[1:])
    assert Yg.shape == (40, 91)
    for ll in range(0, 13, 2):
        sel = l == ll
        c = np.zeros(ll + 1); c[ll] = 1
        np.testing.assert_allclose(Yg[:, sel] @ Yd[:, sel].T, (2 * ll + 1) / (4 * np.pi) * npleg.legval(g @ d.T, c), atol=1e-12)
        i0 = (ll * ll + ll + 2) // 2 - 1
        assert m[i0] == 0 and l[i0] == ll
        np.testing.assert_allclose(Yg[:, i0], np.sqrt((2 * ll + 1) / (4 * np.pi)) * npleg.legval(g[:, 2], c), atol=1e-12)


def test_aux_structures_resample_layout():
    from amico_b200 import lut
    sch = synth.make_scheme(2)
    idx_out, Ylm = lut.aux_structures_resample(sch, 12)
    assert idx_out.dtype == np.int32 and Ylm.dtype == np.float32 and Ylm.shape == (sch.dwi_count, 91 * 2)
    np.testing.assert_array_equal(np.sort(idx_out), sch.dwi_idx)
    n0 = len(sch.shells[0]["idx"])
    assert (Ylm[:n0, 91:] == 0).all() and (Ylm[n0:, :91] == 0).all()


def test_resample_oracle_exact_vs_float32_statement():
    rng = np.random.default_rng(2)
    from amico_b200 import lut
    sch = synth.make_scheme(1)
    idx_out, Ylm = lut.aux_structures_resample(sch, 12)
    lm = (rng.standard_normal((6, Ylm.shape[1])) * 0.3).astype(np.float32)
    a = opl.resample_kernel(lm, sch.nS, idx_out, Ylm, False, 6)
    b = opl.resample_kernel_exact(lm, sch.nS, idx_out, Ylm, False, 6)
    assert a.dtype == np.float32 and (a[:, sch.b0_idx] == 1).all()
    np.testing.assert_allclose(a, b, atol=2e-6)


def test_oracle_dti_wls_sanity():
    """The WLS restatement: on a noise-free single-tensor signal both fits return the generating direction; with noise they differ."""
    sch = synth.make_scheme(2)
    b, g = sch.b, sch.raw[:, :3]
    e = np.array([0.6, 0.48, 0.64])
    D = 0.3e-3 * np.eye(3) + 1.4e-3 * np.outer(e, e)
    s = np.exp(-b * np.einsum("ij,jk,ik->i", g, D, g))[None, :]
    for method in ("OLS", "WLS"):
        d = opl.dti_directions(s, sch, method=method)[0]
        assert abs(abs(d @ e) - 1.0) < 1e-10, method
    rng = np.random.default_rng(0)
    noisy = np.abs(s + 0.05 * rng.standard_normal((50, len(b))))
    a, w = opl.dti_directions(noisy, sch, method="OLS"), opl.dti_directions(noisy, sch, method="WLS")
    assert (1.0 - np.abs((a * w).sum(1))).max() > 1e-6
    assert (np.abs((w * e).sum(1)) > 0.9).all()
