"""Generates the table of lwfb_log_core (include/lwfb90_elem.h): for c_i = 1 + i/128, i = -38..53,
invc_i = RN(1/c_i) and logc_i = -log(invc_i) split in hi (32 significant bits, so that k*LN2_HI + hi is exact)
and lo.  High precision from Python's decimal module (60 digits)."""
import struct
from decimal import Decimal, getcontext
getcontext().prec = 60
def d2hex(x):
    return float(x).hex()
rows = []
for i in range(-38, 54):
    c = Decimal(1) + Decimal(i) / Decimal(128)
    invc = float(Decimal(1) / c)                    # correctly rounded by float(Decimal)
    if i == 0:
        invc = 1.0
    logc = -(Decimal(invc).ln())                    # exact value for the double invc
    hi = float(logc)
    # keep 32 significant bits in hi
    bits = struct.unpack("<q", struct.pack("<d", hi))[0]
    bits &= ~((1 << 21) - 1)
    hi = struct.unpack("<d", struct.pack("<q", bits))[0]
    lo = float(logc - Decimal(hi))
    rows.append((invc, hi, lo))
print("#define LWFB_LOG_TAB_I0 38")
print("#define LWFB_LOG_TAB_N %d" % len(rows))
print("#define LWFB_LOG_TABLE { \\")
for invc, hi, lo in rows:
    print("  %s, %s, %s, \\" % (d2hex(invc), d2hex(hi), d2hex(lo)))
print("}")
