"""Host proof that the FP64 construction of profiles/microbench/fp64_offload.cu yields the exact high word of x * M
(M = 0xD2511F53, a Philox multiplier): every fma / add result is shown to be exactly representable by comparing the
rounded double with the exact rational, for the corner cases and 2e5 random x.  (The construction is exact; it is the
FP64 instructions' issue cost on B200 that refutes the idea -- profiles/r02_fp64_offload.txt.)"""
import random
import struct
from fractions import Fraction as F


def as_double(hi, lo):
    return struct.unpack("<d", struct.pack("<II", lo, hi))[0]


def lo_word(d):
    return struct.unpack("<II", struct.pack("<d", d))[0]


M = 0xD2511F53
Mh_s, Ml_s = (M >> 16) * 2.0 ** -16, (M & 0xFFFF) * 2.0 ** -32      # exact doubles
C = 2.0 ** 52 + 2.0 ** 20
random.seed(1)
for x in [0, 1, 0xFFFFFFFF, 0x80000000, 0x7FFFFFFF] + [random.getrandbits(32) for _ in range(200000)]:
    lo = (x * M) & 0xFFFFFFFF
    xd = float(F(as_double(0x43300000, x)) - F(2.0 ** 52))
    assert xd == x
    vneg = as_double(0xC1300000, lo)                                  # -(2^20 + lo 2^-32)
    u = float(F(xd) * F(Ml_s) + F(vneg))
    assert F(u) == F(xd) * F(Ml_s) + F(vneg)                          # first fma: exactly representable
    s = float(F(xd) * F(Mh_s) + F(u))
    assert F(s) == F(xd) * F(Mh_s) + F(u)                             # second fma: hi - 2^20, an integer
    assert lo_word(float(F(s) + F(C))) == (x * M) >> 32
print("exact for all inputs tried")
