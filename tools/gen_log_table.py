"""Generates the table of lwfb_log_core (include/lwfb90_elem.h).

Reduction: with OFF = bits(1.0) - 75.5 * 2^45, tmp = bits(x) - OFF gives k = tmp >> 52 (x = 2^k z, z in [0.70508, 1.41016))
and the table index i = (tmp >> 45) & 127: 128 intervals of 2^45 bit patterns (2^-8 wide below 1, 2^-7 wide above).
Interval 75 has 1.0 at the centre of its bit patterns and uses c = 1 exactly (invc = 1, logc = 0), so that log x keeps its
relative accuracy near x = 1; every other interval uses its mid point.  Per interval: invc = RN(1/c) and logc = -log(invc)
for that double, split in hi (32 significant bits, so that k * LN2_HI + hi is exact) and lo.
High precision from Python's decimal module (60 digits)."""
import struct
from decimal import Decimal, getcontext
getcontext().prec = 60
OFF = 0x3FF0000000000000 - 151 * (1 << 44)          # 75.5 * 2^45
assert OFF == 0x3FE6900000000000


def val(bits):
    return Decimal(struct.unpack("<d", struct.pack("<q", bits))[0])


def znorm(bits):
    """value of the normalised z for a pattern in [OFF, OFF + 2^52)"""
    k = (bits - OFF) >> 52
    return val(bits - (k << 52))


rows = []
for i in range(128):
    lo_b, hi_b = OFF + i * (1 << 45), OFF + (i + 1) * (1 << 45)
    zl = val(lo_b)
    zh = val(hi_b) if hi_b < OFF + (1 << 52) else val(OFF) * 2
    if zh < zl:                                      # (the interval that crosses the exponent boundary at 1.0)
        zh = val(hi_b)
    c = Decimal(1) if i == 75 else (zl + zh) / 2
    if i == 75:
        assert zl < 1 < zh
    invc = float(Decimal(1) / c)
    logc = -(Decimal(invc).ln())
    hi = float(logc)
    bits = struct.unpack("<q", struct.pack("<d", hi))[0]
    bits &= ~((1 << 21) - 1)
    hi = struct.unpack("<d", struct.pack("<q", bits))[0]
    lo = float(logc - Decimal(hi))
    rmax = max(abs(zl * Decimal(invc) - 1), abs(zh * Decimal(invc) - 1))
    rows.append((invc, hi, lo, float(rmax)))
print("/* max |r| over the table: %.6f */" % max(r[3] for r in rows))
print("#define LWFB_LOG_OFF_HI 0x%08x" % (OFF >> 32))
print("#define LWFB_LOG_TAB_N %d" % len(rows))
print("#define LWFB_LOG_TABLE { \\")
for invc, hi, lo, _ in rows:
    print("  %s, %s, %s, \\" % (float(invc).hex(), float(hi).hex(), float(lo).hex()))
print("}")
