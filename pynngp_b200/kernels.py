"""Covariance specifications the engine can evaluate on device.

The reference takes an arbitrary Python callable ``cov`` (pyNNGP/nngp.py:12) and never calls it on
coordinates; a callable cannot run inside a CUDA kernel and this package has no CPU fallback, so
``cov`` is one of the parametric families below (or ``None`` / a name / a tuple, see ``parse``).

    C(a, b) = sigma2 * rho(phi * ||s_a - s_b||)  (a != b),   C(a, a) = sigma2 + tau2 + eps_a^2
"""
from __future__ import annotations

from dataclasses import dataclass

KERNEL_IDS = {"exponential": 0, "matern32": 1, "matern52": 2}


@dataclass
class Kernel:
    name: str = "exponential"
    sigma2: float | None = None
    phi: float | None = None
    tau2: float | None = None

    @property
    def kernel_id(self) -> int:
        return KERNEL_IDS[self.name]

    def params(self, sigma2=None, phi=None, tau2=None):
        vals = [sigma2 if sigma2 is not None else self.sigma2,
                phi if phi is not None else self.phi,
                tau2 if tau2 is not None else self.tau2]
        if any(v is None for v in vals):
            raise ValueError("sigma2, phi and tau2 must be given (in the call or in the kernel spec)")
        return [float(vals[0]), float(vals[1]), float(vals[2]), 0.0]


def Exponential(sigma2=None, phi=None, tau2=None):
    return Kernel("exponential", sigma2, phi, tau2)


def Matern(nu=1.5, sigma2=None, phi=None, tau2=None):
    nu = float(nu)
    name = {0.5: "exponential", 1.5: "matern32", 2.5: "matern52"}.get(nu)
    if name is None:
        raise ValueError("Matern smoothness must be 0.5, 1.5 or 2.5 (general nu needs Bessel K_nu: not built)")
    return Kernel(name, sigma2, phi, tau2)


def parse(cov) -> Kernel:
    """None | Kernel | 'exponential' | 'matern32' | 'matern52' | ('matern', nu) | ('exponential',)."""
    if cov is None:
        return Kernel("exponential")
    if isinstance(cov, Kernel):
        return cov
    if isinstance(cov, str):
        key = cov.lower().replace("é", "e").replace("_", "").replace("-", "")
        key = {"exp": "exponential", "matern": "matern32", "matern1.5": "matern32", "matern2.5": "matern52",
               "matern3/2": "matern32", "matern5/2": "matern52"}.get(key, key)
        if key not in KERNEL_IDS:
            raise ValueError(f"unknown covariance family {cov!r}")
        return Kernel(key)
    if isinstance(cov, tuple) and cov and isinstance(cov[0], str):
        fam = cov[0].lower()
        if fam.startswith("exp"):
            return Kernel("exponential")
        if fam.startswith("mat"):
            return Matern(cov[1] if len(cov) > 1 else 1.5)
        raise ValueError(f"unknown covariance family {cov!r}")
    if callable(cov):
        raise TypeError(
            "cov must be a parametric kernel spec (pynngp_b200.kernels.Exponential / Matern, a name or a "
            "tuple): an arbitrary Python callable cannot run on the GPU and there is no CPU fallback")
    raise TypeError(f"cannot interpret cov={cov!r}")
