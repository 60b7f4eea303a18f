"""TEST INFRASTRUCTURE - generates tests/golden/*.npz by running the REAL reference (via oracle/ref_loader.py).

Run in the build container only:   python oracle/make_golden.py
The reference ships no golden vectors or tests of its own, so these are outputs of the
reference's functions on small seeded problems; both the oracle (CPU, everywhere) and the CUDA path (GPU box) are
checked against them.  Inputs are stored with the outputs so nothing has to be regenerated at test time.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402
import trips_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def csr_fields(A, prefix):
    A = sp.csr_matrix(A)
    A.sort_indices()
    return {prefix + "_indptr": A.indptr.astype(np.int64), prefix + "_indices": A.indices.astype(np.int32),
            prefix + "_data": A.data, prefix + "_shape": np.array(A.shape)}


def stack_history(info, idx):
    return np.stack([np.asarray(info["xHistory"][i]).reshape(-1) for i in idx], axis=1)


def main():
    ref = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(2022)

    # ---- CT 24 x 24, 16 views -------------------------------------------------------------------------------
    nx, views = 24, 16
    A = O.ct_matrix(nx, O.ct_angles(views))
    x_true = O.shepp_logan(nx).reshape((-1, 1))
    b, delta = O.add_noise(A @ x_true, 0.01, rng)
    out = dict(nx=nx, views=views, x_true=x_true, b=b, delta=delta, **csr_fields(A, "A"))

    U = b / np.linalg.norm(b)
    B = np.empty(1)
    V = np.empty((A.shape[1], 1))
    for _ in range(8):
        U, B, V = ref.decompositions.golub_kahan_update(A, U, B, V)
    out.update(gk_U=U, gk_B=B, gk_V=V)

    x, info = ref.CGLS(A, b, np.zeros((A.shape[1], 1)), 20, 0, x_true=x_true)
    out.update(cgls_x=x, cgls_relres=np.array(info["relResidual"]), cgls_relerr=np.array(info["relError"]))

    for tag, rp, kw in (("fix", 1e-2, {}), ("dp", "dp", {"delta": float(delta)}), ("gcv", "gcv", {})):
        x, info = ref.Hybrid_LSQR(A, b, n_iter=20, regparam=rp, x_true=x_true, **kw)
        out.update({f"hlsqr_{tag}_x": x, f"hlsqr_{tag}_lam": np.array(info["regParam_history"], dtype=float),
                    f"hlsqr_{tag}_relerr": np.array(info["relError"]),
                    f"hlsqr_{tag}_hist": stack_history(info, (0, 9, 18))})

    M = (A.T @ A).tocsr()
    rhs = A.T @ b
    for tag, rp, kw in (("fix", 1e-2, {}), ("dp", "dp", {"delta": float(np.linalg.norm(A.T @ (b - A @ x_true)))})):
        x, info = ref.Hybrid_GMRES(M, rhs, 15, regparam=rp, x_true=x_true, **kw)
        out.update({f"hgmres_{tag}_x": x, f"hgmres_{tag}_lam": np.array(info["regParam_history"], dtype=float)})
    out.update(hgmres_dp_delta=float(np.linalg.norm(A.T @ (b - A @ x_true))))

    L = O.first_derivative_2d(nx, nx)
    for tag, rp, kw in (("fix", 1e-1, {}), ("dp", "dp", {"delta": float(delta)}), ("gcv", "gcv", {})):
        x, info = ref.GKS(A, b, L, projection_dim=3, n_iter=15, regparam=rp, **kw)
        out.update({f"gks_{tag}_x": x, f"gks_{tag}_lam": np.array(info["regParam_history"], dtype=float),
                    f"gks_{tag}_res": np.array(info["Residual"])})
        x, info = ref.MMGKS(A, b, L, pnorm=2, qnorm=1, projection_dim=3, n_iter=15, regparam=rp, **kw)
        out.update({f"mmgks_{tag}_x": x, f"mmgks_{tag}_lam": np.array(info["regParam_history"], dtype=float),
                    f"mmgks_{tag}_res": np.array(info["Residual"])})
    x, info = ref.MMGKS(A, b, L, pnorm=1.5, qnorm=0.8, projection_dim=2, n_iter=10, regparam=1e-1, epsilon=0.05)
    out.update(mmgks_pq_x=x, mmgks_pq_res=np.array(info["Residual"]))
    np.savez_compressed(os.path.join(OUT, "ct24.npz"), **out)

    # ---- deblurring 32 x 32, Gaussian PSF 7 x 7 ---------------------------------------------------------------
    n = 32
    D = ref.Deblurring2D(CommitCrime=False)
    Aop = D.forward_Op((7, 7), (2, 2), n, n)
    PSF = O.gauss_psf((7, 7), (2, 2))
    img = O.shepp_logan(n)
    x_true = img.reshape((-1, 1))
    b_true = D.gen_data(x_true)
    b, delta = O.add_noise(b_true, 0.01, rng)
    probe = rng.standard_normal((n * n, 1))
    out = dict(n=n, PSF=PSF, x_true=x_true, b_true=b_true, b=b, delta=delta, probe=probe,
               fwd_probe=Aop @ probe, adj_probe=Aop.T @ probe)
    x, info = ref.Hybrid_LSQR(Aop, b, n_iter=15, regparam="dp", delta=float(delta))
    out.update(hlsqr_dp_x=x, hlsqr_dp_lam=np.array(info["regParam_history"], dtype=float))
    x, info = ref.Hybrid_GMRES(Aop, b, 15, regparam=1e-3)
    out.update(hgmres_fix_x=x)
    L = O.first_derivative_2d(n, n)
    x, info = ref.MMGKS(Aop, b, L, pnorm=2, qnorm=1, projection_dim=3, n_iter=12, regparam="dp", delta=float(delta))
    out.update(mmgks_dp_x=x, mmgks_dp_lam=np.array(info["regParam_history"], dtype=float))
    x, info = ref.GKS(Aop, b, L, projection_dim=3, n_iter=12, regparam=1e-2)
    out.update(gks_fix_x=x)
    np.savez_compressed(os.path.join(OUT, "deblur32.npz"), **out)

    # ---- parameter rules on a fixed projected problem ----------------------------------------------------------
    k = 6
    Bm = np.diag(rng.uniform(0.5, 2.0, k)) + np.diag(rng.uniform(0.1, 1.0, k - 1), -1)
    Bm = np.vstack((Bm, np.zeros((1, k))))
    Bm[k, k - 1] = 0.3
    bhat = np.zeros(k + 1)
    bhat[0] = 3.0
    Q, s, _ = np.linalg.svd(Bm, full_matrices=False)
    lam_std = ref.gcv.generalized_crossvalidation(Q, np.diag(s), np.eye(k), bhat)
    lam_mod = ref.gcv.generalized_crossvalidation(Q, np.diag(s), np.eye(k), bhat, variant="modified", fullsize=500)
    Qf = np.linalg.qr(rng.standard_normal((40, k + 1)))[0]
    bfull = Qf @ bhat.reshape(-1, 1) + 0.01 * rng.standard_normal((40, 1))
    lam_dp = ref.dp.discrepancy_principle(Qf, Bm, ref.pylops.Identity(k), bfull, delta=0.5)
    RL = np.triu(rng.standard_normal((k, k))) + 2 * np.eye(k)
    RA = np.triu(rng.standard_normal((k, k))) + 2 * np.eye(k)
    Qk = Qf[:, :k]
    lam_dp_L = ref.dp.discrepancy_principle(Qk, RA, RL, bfull, delta=0.8)
    lam_gcv_L = ref.gcv.generalized_crossvalidation(Qk, RA, RL, bfull)
    # L-curve rule (trips/utilities/reg_param/l_curve.py) on the same projected problems
    bp = Qk.T @ bfull
    lc_grid = np.array([1e-6, 1e-3, 0.1, 1.5])
    lc_kappa = np.array([ref.l_curve.curvature(l, RA, RL, bp) for l in lc_grid])
    lam_lc_L = ref.l_curve.l_curve(RA, RL, bp)
    lam_lc_I = ref.l_curve.l_curve(np.diag(s), np.eye(k), Q.T @ bhat.reshape(-1, 1))
    np.savez_compressed(os.path.join(OUT, "regparam.npz"), B=Bm, bhat=bhat, Qf=Qf, bfull=bfull, RA=RA, RL=RL,
                        lam_std=lam_std, lam_mod=lam_mod, lam_dp=lam_dp, lam_dp_L=lam_dp_L, lam_gcv_L=lam_gcv_L,
                        lc_grid=lc_grid, lc_kappa=lc_kappa, lam_lc_L=lam_lc_L, lam_lc_I=lam_lc_I)
    # ---- fan-beam CT 20 x 20, 12 views (the reference's own ASTRA geometry) + the L-curve rule inside the solvers ---
    nxf, vf = 20, 12
    Af = O.ct_matrix(nxf, O.ct_angles(vf), fan=O.fan_geometry(nxf))
    xf = O.shepp_logan(nxf).reshape((-1, 1))
    bf, deltaf = O.add_noise(Af @ xf, 0.01, rng)
    outf = dict(nx=nxf, views=vf, x_true=xf, b=bf, delta=deltaf, **csr_fields(Af, "A"))
    x, info = ref.CGLS(Af, bf, np.zeros((Af.shape[1], 1)), 15, 0, x_true=xf)
    outf.update(cgls_x=x, cgls_relerr=np.array(info["relError"]))
    x, info = ref.Hybrid_LSQR(Af, bf, n_iter=12, regparam="l_curve", x_true=xf)
    outf.update(hlsqr_lc_x=x, hlsqr_lc_lam=np.array(info["regParam_history"], dtype=float))
    Lf = O.first_derivative_2d(nxf, nxf)
    x, info = ref.GKS(Af, bf, Lf, projection_dim=3, n_iter=8, regparam="l_curve")
    outf.update(gks_lc_x=x, gks_lc_lam=np.array(info["regParam_history"], dtype=float))
    x, info = ref.MMGKS(Af, bf, Lf, pnorm=2, qnorm=1, projection_dim=3, n_iter=8, regparam="l_curve")
    outf.update(mmgks_lc_x=x, mmgks_lc_lam=np.array(info["regParam_history"], dtype=float))
    np.savez_compressed(os.path.join(OUT, "ctfan20.npz"), **outf)

    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
