"""`SVT` estimator and `mixed_noise_model` — the Python surface of the reference (pguresvt/svt.py:14-337),
re-implemented over the B200 C ABI.  Constructor arguments, defaults, validation order, error messages,
dtype dispatch, array orders and result attributes follow the reference so that its tests and notebooks run
unchanged; the numerics happen in libpguresvt_b200.so.
"""
import numpy as np

from ._pguresvt import pguresvt_d, pguresvt_f, pguresvt_u8, pguresvt_u16
from ._pguresvt import reversed_axes_copy as _reversed_axes_copy

_ENTRY_POINTS = {
    np.dtype("uint8"): pguresvt_u8,
    np.dtype("uint16"): pguresvt_u16,
    np.dtype("float32"): pguresvt_f,
    np.dtype("float64"): pguresvt_d,
}


def _is_power_of_two(n):
    return n != 0 and (n & (n - 1)) == 0


def mixed_noise_model(X, alpha=1.0, mu=0.0, sigma=0.0, random_state=None):
    """Corrupt X with Poisson-Gaussian noise  Y = alpha * Poisson(X / alpha) + N(mu, sigma^2)  on the
    max-normalised data, scaled back afterwards (reference: svt.py:14-74; draws Poisson first, then normal,
    from a legacy numpy RandomState so the reference's golden values reproduce).

    alpha in (0, 1] is the detector gain, mu the offset, sigma >= 0 the Gaussian read-out noise;
    random_state may be None, an int seed or a numpy RandomState.
    """
    if alpha <= 0.0 or alpha > 1.0:
        raise ValueError("alpha should be in range [0, 1]")
    if sigma < 0.0:
        raise ValueError("sigma should be >= 0.0")
    rng = random_state if isinstance(random_state, np.random.RandomState) else np.random.RandomState(random_state)

    Xf = X.astype(float)
    peak = Xf.max()
    Xf /= peak
    shot = alpha * rng.poisson(Xf / alpha)
    Y = shot + mu + sigma * rng.normal(size=Xf.shape)
    Y *= peak
    return Y


class SVT:
    """Singular value thresholding denoiser for image sequences (PGURE-SVT; Furnival, Leary & Midgley,
    Ultramicroscopy 178 (2017) 112–124).

    Parameters (names and defaults as in the reference, svt.py:172-191)
    ----------
    trajectory_length : odd int, frames per Casorati matrix (15)
    patch_size : int, patch side in pixels (4)
    patch_overlap : int, stride of the patch grid (1)
    motion_estimation : bool, ARPS motion compensation of patch trajectories (True)
    motion_window : odd int > 1, ARPS search neighbourhood in pixels (7)
    motion_filter : int or None, radius of the median prefilter used for motion estimation (5)
    optimize_pgure : bool, choose lambda per frame by minimising the PGURE risk (True)
    lambda1 : float or None, the threshold (optimize_pgure=False) or the start of the search
    exponential_weighting : bool, exponentially weighted thresholds (True)
    noise_method : int 1..4, variant of the mu/sigma estimate (4)
    noise_alpha, noise_mu, noise_sigma : float or None, known noise parameters; None = estimate
    tol : float, relative stopping tolerance of the lambda search (1e-7)
    max_iter : int, evaluation budget of the lambda search (500)
    n_jobs : int or None, accepted for compatibility; the GPU path does not depend on it
    random_seed : int or None, seed of the PGURE perturbations

    Attributes after ``denoise``: ``Y_`` (same shape as the input, C-contiguous float64), ``lambda1s_``,
    ``noise_alphas_``, ``noise_mus_``, ``noise_sigmas_`` (one value per frame).
    """

    def __init__(
        self,
        trajectory_length=15,
        patch_size=4,
        patch_overlap=1,
        motion_estimation=True,
        motion_window=7,
        motion_filter=5,
        optimize_pgure=True,
        lambda1=None,
        exponential_weighting=True,
        noise_method=4,
        noise_alpha=None,
        noise_mu=None,
        noise_sigma=None,
        tol=1e-7,
        max_iter=500,
        n_jobs=None,
        random_seed=None,
    ):
        self.trajectory_length = trajectory_length
        self.patch_size = patch_size
        self.patch_overlap = patch_overlap
        self.motion_estimation = motion_estimation
        self.motion_window = motion_window
        self.motion_filter = motion_filter
        self.optimize_pgure = optimize_pgure
        self.lambda1 = lambda1
        self.exponential_weighting = exponential_weighting
        self.noise_method = noise_method
        self.noise_alpha = noise_alpha
        self.noise_mu = noise_mu
        self.noise_sigma = noise_sigma
        self.tol = tol
        self.max_iter = max_iter
        self.n_jobs = n_jobs
        self.random_seed = random_seed

        self.Y_ = None

    @staticmethod
    def _or_sentinel(value):
        # the C++ side encodes "not given" as -1 (svt.py:222-228)
        return -1 if value is None else value

    def _check_arguments(self, X):
        """Validate arguments in the reference's order with the reference's messages (svt.py:212-271)."""
        if X.min() < 0.0:
            raise ValueError(
                "Negative values found in data. PGURE-SVT "
                "requires strictly non-negative image data."
            )

        self.lambda1_ = self._or_sentinel(self.lambda1)
        self.motion_filter_ = self._or_sentinel(self.motion_filter)
        self.noise_alpha_ = self._or_sentinel(self.noise_alpha)
        self.noise_mu_ = self._or_sentinel(self.noise_mu)
        self.noise_sigma_ = self._or_sentinel(self.noise_sigma)
        self.n_jobs_ = self._or_sentinel(self.n_jobs)
        self.random_seed_ = self._or_sentinel(self.random_seed)

        if self.patch_overlap > self.patch_size:
            raise ValueError(
                f"Invalid patch_overlap parameter: got {self.patch_overlap}, "
                f"should not be greater than patch_size ({self.patch_size})"
            )

        if self.trajectory_length % 2 == 0 or self.trajectory_length < 1:
            raise ValueError(
                f"Invalid trajectory_length parameter: got {self.trajectory_length},"
                "but expected a positive, odd-valued integer"
            )

        if self.motion_estimation:
            if self.motion_window < 2 or self.motion_window % 2 == 0:
                raise ValueError(
                    f"Invalid motion_window parameter: got {self.motion_window}, "
                    "should be greater a positive, odd-valued integer > 1 pixel."
                )
            if not isinstance(self.motion_filter_, int):
                raise ValueError(
                    f"Invalid motion_filter parameter: got {type(self.motion_filter)}, "
                    "should be an integer number of pixels or None."
                )

        if not self.optimize_pgure and (self.lambda1 is None or self.lambda1 < 0.0):
            raise ValueError(
                f"Invalid lambda1 parameter: got {self.lambda1}, "
                "should be a float >= 0.0 if optimize_pgure is None."
            )

        unknown_noise = any(v < 0.0 for v in (self.noise_alpha_, self.noise_mu_, self.noise_sigma_))
        if unknown_noise:
            if X.shape[0] != X.shape[1]:
                raise ValueError(f"Quadtree noise estimation requires square images, got {X.shape}")
            if not _is_power_of_two(X.shape[0]):
                raise ValueError("Quadtree noise estimation requires image dimensions 2^N")

    def denoise(self, X):
        """Denoise the sequence X of shape (rows, cols, frames); returns self (svt.py:273-337)."""
        self._check_arguments(X)

        X_dtype = getattr(X, "dtype", None)
        if X_dtype not in _ENTRY_POINTS:
            raise TypeError(
                f"Invalid dtype: got {X_dtype}, but only {list(_ENTRY_POINTS.keys())} are supported"
            )

        if not X.flags.f_contiguous:
            X = np.asfortranarray(X, dtype=X_dtype)

        Xd, estimates, _ = _ENTRY_POINTS[X_dtype](
            input_images=X,
            trajectory_length=self.trajectory_length,
            patch_size=self.patch_size,
            patch_overlap=self.patch_overlap,
            motion_estimation=self.motion_estimation,
            motion_window=self.motion_window,
            motion_filter=self.motion_filter_,
            optimize_pgure=self.optimize_pgure,
            lambda1=self.lambda1_,
            exponential_weighting=self.exponential_weighting,
            noise_method=self.noise_method,
            noise_alpha=self.noise_alpha_,
            noise_mu=self.noise_mu_,
            noise_sigma=self.noise_sigma_,
            tol=self.tol,
            max_iter=self.max_iter,
            n_jobs=self.n_jobs_,
            random_seed=self.random_seed_,
        )

        # bridge output is (frames, cols, rows): transpose back to the caller's axis order, C-contiguous
        self.Y_ = _reversed_axes_copy(Xd)
        self.lambda1s_, self.noise_alphas_, self.noise_mus_, self.noise_sigmas_ = (estimates[i] for i in range(4))
        return self
