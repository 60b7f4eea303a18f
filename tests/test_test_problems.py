"""The reference's test-problem classes (Deblurring2D, Tomography) re-hosted on the GPU operators.

CPU part: the analytic phantoms equal the oracle's (and, when /root/reference is mounted, the reference's own
phantoms.py bit for bit).  GPU part: the demo call sequence of the reference - forward_Op / gen_true / gen_data /
add_noise / solver - against the oracle's statements of Deblurring2D.py:66-147 and the CT line model.
"""
import os

import numpy as np
import pytest

import trips_oracle as O


def test_phantoms_match_oracle_and_reference():
    import trips_b200.test_problems as TP

    for n in (16, 33, 64):
        assert np.array_equal(TP.shepp_logan(n), O.shepp_logan(n))
    ref_file = "/root/reference/trips/utilities/phantoms.py"
    if os.path.exists(ref_file):  # container only; the GPU box has no reference tree
        import importlib.util

        spec = importlib.util.spec_from_file_location("_ref_phantoms", ref_file)
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        for n in (16, 33, 64):
            assert np.array_equal(TP.shepp_logan(n), ref.shepp_logan(n))
            assert np.array_equal(TP.smooth(n), ref.smooth(n))


def test_gen_true_argument_errors_mirror_the_reference():
    import trips_b200.test_problems as TP

    with pytest.raises(TypeError, match="dimension of the image is not specified"):
        TP.Deblurring2D().gen_true("satellite")
    with pytest.raises(ValueError, match="does not exist"):
        TP.Deblurring2D().gen_true("no_such_image", nx=8, ny=8)
    with pytest.raises(TypeError, match="dimension of the image is not specified"):
        TP.Tomography().gen_true("smooth")
    with pytest.raises(TypeError, match="valid test problem"):
        TP.Tomography().gen_true("no_such_phantom", nx=8, ny=8)
    x, nx, ny = TP.Tomography().gen_true("smooth", nx=8, ny=8)
    assert x.shape == (64, 1) and (nx, ny) == (8, 8)


def test_seeded_noise_has_the_requested_level():
    import trips_b200.test_problems as TP

    D = TP.Deblurring2D(seed=7)
    D.nx, D.ny = 6, 5
    b = np.arange(30.0).reshape((-1, 1))
    bm, delta = D.add_noise(b, "Gaussian", 0.05)
    ref_b, ref_delta = O.add_noise(b, 0.05, np.random.default_rng(7))
    assert bm.shape == (6, 5)
    assert np.array_equal(bm.reshape((-1, 1)), ref_b) and delta == float(ref_delta)
    assert abs(delta / np.linalg.norm(b) - 0.05) < 1e-15


@pytest.mark.gpu
def test_deblurring_demo_sequence_matches_oracle():
    import trips_b200 as tb

    nx = ny = 48
    for crime in (False, True):
        D = tb.Deblurring2D(CommitCrime=crime, seed=11)
        A = D.forward_Op([9, 9], (1.5, 1.5), nx, ny)
        PSF, centre = D.Gauss([9, 9], (1.5, 1.5))
        assert np.array_equal(PSF, O.gauss_psf([9, 9], (1.5, 1.5))) and tuple(centre) == (4, 4)
        x = D.gen_true("shepp_logan")
        b = D.gen_data(x)
        want = (O.blur_operator(PSF, nx, ny) @ x.reshape((-1, 1))) if crime else O.blur_data(x, PSF, nx, ny)
        assert np.array_equal(b, want)  # ndimage tap order and rounding reproduced exactly
        bm, delta = D.add_noise(b, "Gaussian", 0.01)
        with O.reductions("exact"):
            x_ref, info_ref = O.Hybrid_LSQR(O.blur_operator(PSF, nx, ny), bm.reshape((-1, 1)), n_iter=20,
                                            regparam="dp", delta=delta, x_true=x.reshape((-1, 1)))
        x_gpu, info = tb.Hybrid_LSQR(A, bm.reshape((-1, 1)), n_iter=20, regparam="dp", delta=delta,
                                     x_true=x.reshape((-1, 1)))
        assert np.linalg.norm(x_gpu - x_ref) / np.linalg.norm(x_ref) <= 1e-10


@pytest.mark.gpu
def test_tomography_demo_sequence_matches_oracle():
    import trips_b200 as tb

    nx = ny = 32
    views = 24
    T = tb.Tomography(CommitCrime=False, seed=3, geometry="parallel")
    x_true, _, _ = T.gen_true("smooth", nx=nx, ny=ny)
    A, b_true, p, q, AforMatrixOperation = T.gen_data(x_true, nx, ny, views)
    assert (p, q) == (views, int(np.sqrt(2) * nx)) and AforMatrixOperation is A
    theta = np.linspace(0, np.pi, views, endpoint=False)
    A_mis = O.ct_matrix(nx, theta + 1e-8)
    assert np.array_equal(b_true, A_mis @ x_true)  # data from the slightly rotated geometry (Tomography.py:62-65,159)
    assert not np.array_equal(b_true, O.ct_matrix(nx, theta) @ x_true)
    b, delta = T.add_noise(b_true, "Gaussian", 0.02)
    assert b.shape == (views, q)
    with O.reductions("exact"):
        x_ref, _ = O.CGLS(O.ct_matrix(nx, theta), b.reshape((-1, 1)), np.zeros((nx * ny, 1)), 30, 0.0)[:2]
    x_gpu, _ = tb.CGLS(A, b.reshape((-1, 1)), np.zeros((nx * ny, 1)), 30, 0.0)[:2]
    assert np.linalg.norm(x_gpu - x_ref) / np.linalg.norm(x_ref) <= 1e-10
    # CommitCrime=True: two return values, data from the operator itself
    T2 = tb.Tomography(CommitCrime=True, geometry="parallel", layout="implicit")
    ops = T2.forward_Op(nx, ny, views)
    assert len(ops) == 2
    assert np.array_equal(T2.gen_data(x_true, nx, ny, views)[1], O.ct_matrix(nx, theta) @ x_true)
    # the reference's own geometry (default): flat-detector fan beam, source at 3 nx, detector at nx, bins 4/3 wide
    T3 = tb.Tomography(CommitCrime=True, seed=3)
    A3, b3, p3, q3, _ = T3.gen_data(x_true, nx, ny, views)
    F = O.ct_matrix(nx, theta, fan=O.fan_geometry(nx))
    assert isinstance(A3, tb.FanBeamCT) and np.array_equal(b3, F @ x_true)
    bm, delta = T3.add_noise(b3, "Gaussian", 0.02)
    with O.reductions("exact"):
        x_ref, _ = O.Hybrid_LSQR(F, bm.reshape((-1, 1)), n_iter=15, regparam="dp", delta=delta)
    x_gpu, _ = tb.Hybrid_LSQR(A3, bm.reshape((-1, 1)), n_iter=15, regparam="dp", delta=delta)
    assert np.linalg.norm(x_gpu - x_ref) / np.linalg.norm(x_ref) <= 1e-10
