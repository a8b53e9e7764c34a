"""Mualem-van Genuchten soil hydraulics needed on the host at setup time and by the
post-processing getters.  Follows src/module_KPT.jl of LWFBrook90.jl v0.9.9:
  KPT_SOILPAR_Mvg1d constructor   :333-406
  derive_auxiliary_SOILVAR        :432-443
  FK_MvG                          :495-500
  FPSIM_MvG                       :590-594
  FTheta_MvG                      :683-685
  FWETNES                         :706-722
The device and the oracle carry their own versions of the per-step functions.
"""
from __future__ import annotations

import numpy as np

RHOWG = 0.00981  # src/module_CONSTANTS.jl:13


def FK_MvG(W, KSAT, l, n):
    A = np.maximum(W, 1.0e-6)
    return KSAT * A ** l * (1 - (1 - A ** (n / (n - 1))) ** (1 - 1 / n)) ** 2


def FPSIM_MvG(W, alpha, n, m):
    return 9.81 * (-1 / alpha) * (np.maximum(W, 1.0e-6) ** (-1 / m) - 1) ** (1 / n)


def FTheta_MvG(W, THSAT, thr):
    return W * (THSAT - thr) + thr


def FWETNES(psim, alpha, n, m):
    psim = np.asarray(psim, dtype=float)
    W = np.ones_like(psim)
    neg = psim <= 0
    W[neg] = (1 + (-alpha[neg] * psim[neg] / 9.81) ** n[neg]) ** (-m[neg])
    return W


def _bisect_wetf(Kfc, KSAT, l, n):
    """Root of FK_MvG(W) - Kfc on (0,1); bisection down to adjacent floats, which is what
    Roots.find_zero(f, (0.0, 1.0), Bisection()) delivers (src/module_KPT.jl:359-361)."""
    f = lambda x: -Kfc + float(FK_MvG(np.float64(x), KSAT, l, n))
    a, b = 0.0, 1.0
    fa, fb = f(a), f(b)
    if fa == 0.0:
        return a
    if fb == 0.0:
        return b
    if not (fa < 0.0 < fb or fb < 0.0 < fa):
        raise ValueError(f"Computed invalid p_WETF (Kfc={Kfc}, Ksat={KSAT}, l={l}, n={n})")
    for _ in range(2000):
        mid = 0.5 * (a + b)
        if mid <= a or mid >= b:
            break
        fm = f(mid)
        if fm == 0.0:
            return mid
        if (fm < 0) == (fa < 0):
            a, fa = mid, fm
        else:
            b, fb = mid, fm
    return a if abs(fa) <= abs(fb) else b


class SoilPar:
    """Discretised column with MvG parameters and derived fields
    (reference: KPT_SOILPAR_Mvg1d, src/module_KPT.jl:310-407)."""

    def __init__(self, THICK, STONEF, THSAT, KSAT, alpha, n, m, l, thr, Kfc=2.0):
        f = lambda x: np.array(x, dtype=float)
        self.p_THICK, self.p_STONEF, self.p_THSAT = f(THICK), f(STONEF), f(THSAT)
        self.p_KSAT, self.p_MvGalpha, self.p_MvGn = f(KSAT), f(alpha), f(n)
        self.p_MvGm, self.p_MvGl, self.p_thr = f(m), f(l), f(thr)
        N = self.NLAYER = len(self.p_THICK)
        self.p_SWATMAX = self.p_THICK * self.p_THSAT * (1.0 - self.p_STONEF)
        psig = np.empty(N)
        for i in range(N):
            if i == 0:
                psig[0] = -RHOWG * self.p_THICK[0] / 2
            else:
                psig[i] = psig[i - 1] - RHOWG * ((self.p_THICK[i - 1] + self.p_THICK[i]) / 2)
        self.p_PSIG = psig
        Kfc = np.broadcast_to(np.asarray(Kfc, dtype=float), (N,))
        self.p_WETF = np.array([_bisect_wetf(Kfc[i], self.p_KSAT[i], self.p_MvGl[i], self.p_MvGn[i])
                                for i in range(N)])
        self.p_PSIF = FPSIM_MvG(self.p_WETF, self.p_MvGalpha, self.p_MvGn, self.p_MvGm)
        self.p_THETAF = FTheta_MvG(self.p_WETF, self.p_THSAT, self.p_thr)
        self.dz_m_cells = self.p_THICK / 1000
        dzc = self.dz_m_cells
        centres = (np.cumsum(dzc) + np.cumsum(np.concatenate([[0.0], dzc[:-1]]))) / 2
        self.dz_m = np.diff(centres)

    def soilvar(self, SWATI):
        """(WETNES, PSIM, PSITI, theta, KK) from SWATI (reference :432-443)."""
        S = np.asarray(SWATI, dtype=float)
        W = (self.p_THSAT * S / self.p_SWATMAX - self.p_thr) / (self.p_THSAT - self.p_thr)
        W = np.minimum(1.0, W)
        psim = FPSIM_MvG(W, self.p_MvGalpha, self.p_MvGn, self.p_MvGm)
        theta = FTheta_MvG(W, self.p_THSAT, self.p_thr)
        kk = FK_MvG(W, self.p_KSAT, self.p_MvGl, self.p_MvGn)
        return W, psim, psim + self.p_PSIG, theta, kk
