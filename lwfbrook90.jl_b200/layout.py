"""Layout contract shared with the C ABI.

The enums in ``include/lwfb90.h`` are the single source of truth for the flat per-member
parameter blocks, the forcing table and the accumulator order.  This module parses that header
so the Python host side can never drift from the C side.
"""
from __future__ import annotations

import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "lwfb90.h")


def _parse_enums(path: str) -> dict:
    txt = open(path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    out = {}
    for body in re.findall(r"enum\s*\{(.*?)\}", txt, flags=re.S):
        val = -1
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                name, v = [s.strip() for s in item.split("=")]
                val = int(v, 0)
            else:
                name = item
                val += 1
            out[name] = val
    return out


_E = _parse_enums(HEADER)


class _NS:
    def __init__(self, prefix: str, stop):
        self._names = []
        for k, v in _E.items():
            if k.startswith(prefix) and k != stop:
                short = k[len(prefix):]
                setattr(self, short, v)
                self._names.append((v, short))
        self._names.sort()
        self.count = _E[stop] if stop else len(self._names)

    def names(self):
        return [n for _, n in self._names]


F = _NS("LWFB_F_", "LWFB_F")        # forcing columns
S = _NS("LWFB_S_", "LWFB_PS")       # per-member scalars
L = _NS("LWFB_L_", "LWFB_PL")       # per-member per-layer vectors
A = _NS("LWFB_A_", "LWFB_NACCUM")   # accumulators
C = _NS("LWFB_C_", "LWFB_NCOUNTERS")
SC = _NS("LWFB_SC_", "LWFB_NSCORES")
ST = _NS("LWFB_ST_", None)
WP = _NS("LWFB_WP_", "LWFB_NWP")     # water-partitioning columns
SV = _NS("LWFB_SV_", "LWFB_NSV")     # soil variables of get_soil_

NF = _E["LWFB_F"]
NPS = _E["LWFB_PS"]
NPL = _E["LWFB_PL"]
NACCUM = _E["LWFB_NACCUM"]
NCOUNTERS = _E["LWFB_NCOUNTERS"]
NSCORES = _E["LWFB_NSCORES"]
NWP = _E["LWFB_NWP"]
NSV = _E["LWFB_NSV"]

ACCUM_NAMES = (
    "cum_d_prec", "cum_d_rfal", "cum_d_sfal", "cum_d_rint", "cum_d_sint", "cum_d_rsno",
    "cum_d_rnet", "cum_d_smlt", "cum_d_evap", "cum_d_tran", "cum_d_irvp", "cum_d_isvp",
    "cum_d_slvp", "cum_d_snvp", "cum_d_pint", "cum_d_ptran", "cum_d_pslvp",
    "flow", "seep", "srfl", "slfl", "byfl", "dsfl", "gwfl", "vrfln",
    "cum_d_rthr", "cum_d_sthr",
    "StorageSWAT", "StorageWATER", "BALERD_SWAT", "BALERD_total")
assert len(ACCUM_NAMES) == NACCUM


def state_len(N: int, iso: bool) -> int:
    return 9 * N + 51 if iso else 5 * N + 39


class StateIndex:
    """Offsets of the reference's ComponentArray fields inside one state record
    (reference: src/func_DiffEq_definition_u0.jl:42-73)."""

    def __init__(self, N: int, iso: bool):
        self.N, self.iso = N, iso
        k = 3 if iso else 1
        o = 0
        self.SWATI_mm = slice(o, o + N)
        self.SWATI_d18O = slice(o + N, o + 2 * N) if iso else None
        self.SWATI_d2H = slice(o + 2 * N, o + 3 * N) if iso else None
        o += k * N
        for name in ("GWAT", "INTS", "INTR", "SNOW", "RWU", "XYLEM"):
            setattr(self, name + "_mm", o)
            setattr(self, name + "_d18O", o + 1 if iso else None)
            setattr(self, name + "_d2H", o + 2 if iso else None)
            o += k
        self.CC = o
        self.SNOWLQ = o + 1
        o += 2
        self.TRANI_mm = slice(o, o + N)
        self.TRANI_d18O = slice(o + N, o + 2 * N) if iso else None
        self.TRANI_d2H = slice(o + 2 * N, o + 3 * N) if iso else None
        o += k * N
        self.aux_theta = slice(o, o + N)
        self.aux_psi = slice(o + N, o + 2 * N)
        self.aux_K = slice(o + 2 * N, o + 3 * N)
        o += 3 * N
        self.accum = slice(o, o + NACCUM)
        self.accum0 = o
        o += NACCUM
        self.L = o
        assert o == state_len(N, iso)

    def acc(self, name: str) -> int:
        return self.accum0 + ACCUM_NAMES.index(name)
