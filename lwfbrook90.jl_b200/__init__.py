"""lwfbrook90.jl_b200 — B200-native batched integrator behind the LWFBrook90.jl API.

Host-side mirror of the reference's user API for the one accelerated path
(loadSPAC / setup / remakeSPAC -> simulate_ -> getters).  All numerics of the hot path run in
hand-written sm_100a CUDA kernels behind the C ABI of include/lwfb90.h; there is no CPU fallback.
"""
from .inputs import SPAC, loadSPAC
from .setup import DiscretizedSPAC, remakeSPAC, setup
from .discretize import Rootden_beta_

__all__ = ["SPAC", "DiscretizedSPAC", "loadSPAC", "setup", "remakeSPAC", "Rootden_beta_"]


def __getattr__(name):
    # lazy: importing the API module loads the CUDA shared library
    if name in ("simulate_", "simulate", "simulate_ensemble", "Ensemble", "ODESolution", "get_soil_",
                "get_delta", "get_δ", "get_amounts", "get_theta", "get_psi", "get_deltasoil",
                "get_water_partitioning", "get_daily_soilFluxes", "capi", "api", "ensemble"):
        import importlib
        if name in ("capi", "api", "ensemble"):
            return importlib.import_module(f"{__name__}.{name}")
        api = importlib.import_module(f"{__name__}.api")
        return getattr(api, name)
    raise AttributeError(name)
