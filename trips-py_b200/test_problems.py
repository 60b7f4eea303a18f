"""Test-problem classes with the reference's method names, producing GPU operators.

Mirrors the call patterns of trips/test_problems/Deblurring2D.py and trips/test_problems/Tomography.py (the demos'
`Deblur.forward_Op(...)`, `gen_data`, `add_noise`, ... sequence), so a demo script switches by changing its import.
Only the problem SET-UP lives here (host-side, done once); the operators it returns are the CUDA-backed ones of
`trips_b200.operators`.

Deliberate differences, all forced by the environment:
  * Tomography(geometry='fan') - the default - is the reference's geometry: flat-detector fan beam, line model, source at
    3 nx, detector at nx, bins of width 4/3 (`astra.create_proj_geom('fanflat', ...)`, `'line_fanflat'`,
    Tomography.py:57-67), built by this package's FanBeamCT because ASTRA is not part of the reference tree (its rotation
    sense / axis orientation follow ASTRA's documented conventions, unpinned).  geometry='parallel' gives the
    parallel-beam operator, whose matrix-free layout is the fast path (`layout='implicit'`).
  * `seed=` is honoured (the reference pops it and never uses it: noise there is unseeded, Tomography.py:43,206).
  * `gen_true` offers the deterministic analytic phantoms (shepp_logan, smooth) and, like the reference, images from
    `./data/image_data/<name>.mat` when that file exists; the random phantoms and downloads are not reproduced.
"""
from os.path import exists

import numpy as np
import torch

from .kernels import F64
from .operators import (FanBeamCT, ParallelBeamCT, PSFBlur2D, ct_angles, ct_num_detectors, default_device, gauss_psf,
                        to_device_vector)


def shepp_logan(n):
    """Modified Shepp-Logan phantom: ten ellipses on [-1,1]^2 sampled with spacing 2/(n-1), negatives clipped
    (the construction of trips/utilities/phantoms.py:18-60)."""
    table = ((1.0, .69, .92, 0, 0, 0), (-.8, .6624, .8740, 0, -.0184, 0), (-.2, .1100, .3100, .22, 0, -18),
             (-.2, .1600, .4100, -.22, 0, 18), (.1, .2100, .2500, 0, .35, 0), (.1, .0460, .0460, 0, .1, 0),
             (.1, .0460, .0460, 0, -.1, 0), (.1, .0460, .0230, -.08, -.605, 0), (.1, .0230, .0230, 0, -.606, 0),
             (.1, .0230, .0460, .06, -.605, 0))
    g = (np.arange(n) - (n - 1) / 2) / ((n - 1) / 2)
    X, Y = np.meshgrid(g, -g)
    img = np.zeros((n, n))
    for amp, a, b, x0, y0, phi in table:
        ph = phi * np.pi / 180
        xr = (X - x0) * np.cos(ph) + (Y - y0) * np.sin(ph)
        yr = (Y - y0) * np.cos(ph) - (X - x0) * np.sin(ph)
        img[(xr ** 2) / a ** 2 + (yr ** 2) / b ** 2 <= 1] += amp
    img[img < 0] = 0
    return img


def smooth(n, p=4):
    """Sum of four anisotropic Gaussians, normalised to max 1 (trips/utilities/phantoms.py:103-118)."""
    idx = np.arange(n)
    I, J = np.meshgrid(idx, idx, indexing="xy")
    sigma = 0.25 * n
    centres = np.array([[0.6, 0.6], [0.5, 0.3], [0.2, 0.7], [0.8, 0.2]]) * n
    amps = (1, 0.5, 0.7, 0.9)
    img = np.zeros((n, n))
    for i in range(p):
        img += amps[i] * np.exp(-(I - centres[i, 0]) ** 2 / (1.2 * sigma) ** 2 - (J - centres[i, 1]) ** 2 / sigma ** 2)
    return img / img.max()


_PHANTOMS = {"shepp_logan": shepp_logan, "smooth": smooth}


def _load_mat_image(name):
    import scipy.io as spio

    path = f"./data/image_data/{name}.mat"
    if not exists(path):
        raise FileNotFoundError(f"{path} not found (the reference ships these images with its demos; copy demos/data "
                                "next to your script) - or use one of the analytic phantoms: " + ", ".join(_PHANTOMS))
    X = spio.loadmat(path)["x_true"]
    if X.ndim == 3:
        X = 0.4 * X[:, :, 0] + 0.4 * X[:, :, 1] + 0.1 * X[:, :, 2]
    return np.asarray(X, dtype=np.float64)


class _Problem:
    def __init__(self, **kwargs):
        seed = kwargs.pop("seed", None)
        self._rng = np.random.default_rng(seed) if seed is not None else None
        self.nx = None
        self.ny = None
        self.CommitCrime = kwargs["CommitCrime"] if ("CommitCrime" in kwargs) else False
        self.device = kwargs.get("device")

    def _randn(self, n):
        return self._rng.standard_normal(n) if self._rng is not None else np.random.randn(n)

    def _gaussian_noise(self, b_true, noise_level):
        """e = noise_level * ||b|| / ||n|| * n ; delta = ||e||   (Tomography.py:204-212, Deblurring2D.py:142-147)."""
        b_true = np.asarray(b_true, dtype=np.float64).reshape((-1, 1))
        noise = self._randn(b_true.shape[0]).reshape((-1, 1))
        e = noise_level * np.linalg.norm(b_true) / np.linalg.norm(noise) * noise
        return b_true + e, float(np.linalg.norm(e))


class Deblurring2D(_Problem):
    """trips/test_problems/Deblurring2D.py:41-160 on the GPU operators."""

    def Gauss(self, PSFdim, PSFspread):
        self.m, self.n = PSFdim[0], PSFdim[1]
        self.dim, self.spread = PSFdim, PSFspread
        PSF = gauss_psf(PSFdim, PSFspread)
        mm, nn = np.where(PSF == PSF.max())
        return PSF, np.array([mm[0], nn[0]]).astype(int)

    def forward_Op(self, dim, spread, nx, ny):
        self.nx, self.ny = nx, ny
        PSF, _ = self.Gauss(dim, spread)
        return PSFBlur2D(PSF, nx, ny, device=self.device)

    def gen_true(self, im, **kwargs):
        if self.nx is None or self.ny is None:
            if ("nx" in kwargs) and ("ny" in kwargs):
                self.nx, self.ny = kwargs["nx"], kwargs["ny"]
            else:
                raise TypeError("The dimension of the image is not specified. You can input nx and ny as gen_true(im, nx, ny) or first define the forward operator through A = Deblur.forward_Op([11,11], 0.7, nx, ny) ")
        if im in _PHANTOMS:
            return _PHANTOMS[im](self.nx)
        if im in ["satellite", "hubble", "h_im", "shape"]:
            image = _load_mat_image(im)
            if image.shape != (self.nx, self.ny):
                raise ValueError(f"{im}.mat is {image.shape}; resizing is not provided - build the operator for that size")
            return image
        raise ValueError("The image you requested does not exist! Specify the right name. Options are 'satellite', 'hubble', 'h_im")

    def gen_data(self, x):
        """Blurred data.  CommitCrime == False: blur the image embedded in a zero-padded 2x frame with zero boundary
        conditions and crop (Deblurring2D.py:121-133); True: the operator itself (reflect)."""
        dev = torch.device(self.device) if self.device is not None else default_device()
        PSF, _ = self.Gauss(self.dim, self.spread)
        if self.CommitCrime:
            return PSFBlur2D(PSF, self.nx, self.ny, device=dev) @ np.asarray(x, dtype=np.float64).reshape((-1, 1))
        nxb, nyb = 2 * self.nx, 2 * self.ny
        px, py = self.nx // 2, self.ny // 2
        big = torch.zeros((nxb, nyb), dtype=F64, device=dev)
        big[px:px + self.nx, py:py + self.ny] = to_device_vector(x, dev).reshape(self.nx, self.ny)
        op = PSFBlur2D(PSF, nxb, nyb, device=dev, mode="constant")
        blurred = op.apply_dev(big.reshape(-1).contiguous()).reshape(nxb, nyb)
        return blurred[px:px + self.nx, py:py + self.ny].contiguous().cpu().numpy().reshape((-1, 1))

    def add_noise(self, b_true, opt, noise_level):
        if opt == "Gaussian":
            b_meas, delta = self._gaussian_noise(b_true, noise_level)
            return b_meas.reshape((self.nx, self.ny)), delta
        raise NotImplementedError("only opt='Gaussian' is provided (Poisson / Laplace draws are host-side one-liners)")


class Tomography(_Problem):
    """trips/test_problems/Tomography.py:41-227 (forward_Op :78-88, gen_data :153-168) on this package's CT operators
    (geometry='fan' | 'parallel', layout= passed to the operator; see the module docstring)."""

    def __init__(self, **kwargs):
        self.geometry = kwargs.pop("geometry", "fan")
        self.layout = kwargs.pop("layout", "auto")
        if self.geometry not in ("fan", "parallel"):
            raise ValueError("geometry must be 'fan' (the reference's) or 'parallel'")
        super().__init__(**kwargs)

    def _operator(self, nx, ny, views, angles):
        cls = FanBeamCT if self.geometry == "fan" else ParallelBeamCT
        return cls(nx, views, ny=ny, angles=angles, device=self.device, layout=self.layout)

    def forward_Op(self, nx, ny, views):
        self.nx, self.ny, self.views = nx, ny, views
        self.p, self.q = ct_num_detectors(nx), views
        self.theta = ct_angles(views)
        A = self._operator(nx, ny, views, self.theta)
        self.A = A
        if self.CommitCrime is False:
            self.A_mis = self._operator(nx, ny, views, self.theta + 1e-8)  # :62-65,74-75
            return A, A, self.A_mis
        return A, A

    def gen_true(self, test_problem, **kwargs):
        if self.nx is None or self.ny is None:
            if ("nx" in kwargs) and ("ny" in kwargs):
                self.nx, self.ny = kwargs["nx"], kwargs["ny"]
            else:
                raise TypeError("The dimension of the image is not specified. You can input nx and ny as (x_true, nx, ny) = Tomo.gen_true(testproblem, nx = nx, ny = ny) or first define the forward operator")
        if test_problem in _PHANTOMS:
            x_true = _PHANTOMS[test_problem](self.nx)
            return x_true.reshape((-1, 1)), self.nx, self.ny
        raise TypeError("You must enter a valid test problem! Options here are: " + ", ".join(_PHANTOMS)
                        + " (the reference's random phantoms and saved data sets are not reproduced)")

    def gen_data(self, x, nx, ny, views):
        """(A, b, p, q, AforMatrixOperation) as Tomography.py:153-168; data from the slightly rotated geometry unless
        CommitCrime."""
        ops = self.forward_Op(nx, ny, views)
        A = ops[0]
        src = A if self.CommitCrime else ops[2]
        b = src @ np.asarray(x, dtype=np.float64).reshape((-1, 1))
        self.p = views
        self.q = int(b.shape[0] / views)
        return A, b, self.p, self.q, ops[1]

    def add_noise(self, b_true, opt, noise_level):
        if opt == "Gaussian":
            b_meas, delta = self._gaussian_noise(b_true, noise_level)
            return b_meas.reshape((self.views, -1)), delta
        raise NotImplementedError("only opt='Gaussian' is provided")
