"""Generates the table of lwfb_exp_main (include/lwfb90_elem.h): T[j] = 2^(j/64), j = 0..63, as the double nearest to it
and the relative tail (2^(j/64) - hi) / hi, so that hi * (1 + tail) carries ~100 bits.  mpmath, 60 digits."""
import mpmath as mp
mp.mp.dps = 60
N = 64
print("#define LWFB_EXP_TAB_N %d" % N)
print("#define LWFB_EXP_TABLE { \\")
for j in range(N):
    v = mp.power(2, mp.mpf(j) / N)
    hi = float(v)
    tail = float((v - mp.mpf(hi)) / mp.mpf(hi))
    print("  %s, %s, \\" % (hi.hex(), tail.hex()))
print("}")
ln2 = mp.log(2)
print("/* 64/ln2 = %s */" % float(N / ln2).hex())
