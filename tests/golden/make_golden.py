#!/usr/bin/env python
"""Regenerates tests/golden/curand_host.json from cuRAND's OWN headers compiled for the host
(oracle/curand_host.cpp -> oracle/_ref/libcurand_host.so, CUDA 12.9 / cuRAND 10.3.10).

The reference (edo01/NMCH) has no tests or fixtures of its own; its arithmetic
lives in the cuRAND device API, so these vectors are the known-answer pins for the oracle's
restatement of that API.  Run in the build container (needs /usr/local/cuda/include):

    make -C oracle _ref/libcurand_host.so && python tests/golden/make_golden.py
"""
import ctypes as C
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libcurand_host.so"))
ULL = C.c_ulonglong

SEEDS = [1234, 0, 0xDEADBEEFCAFEF00D]
SUBSEQS = [0, 1, 2, 3, 4, 5, 15, 16, 255, 262143, 2**24 - 1, 2**24 + 12345, 2**30 + 5, 2**31 - 1, 2**33 + 7]


def u32(kind, seed, subseq, offset, n):
    out = (C.c_uint32 * n)()
    lib.crh_u32(kind, ULL(seed), ULL(subseq), ULL(offset), n, out)
    return list(out)


def xorwow_init(seed, subseq, offset):
    d = C.c_uint32()
    v = (C.c_uint32 * 5)()
    lib.crh_xorwow_init(ULL(seed), ULL(subseq), ULL(offset), C.byref(d), v)
    return [d.value] + list(v)


def mrg_init(seed, subseq, offset):
    st = (C.c_uint32 * 6)()
    lib.crh_mrg_init(ULL(seed), ULL(subseq), ULL(offset), st)
    return list(st)


def state_after(seed, subseq, n):
    d = C.c_uint32()
    v = (C.c_uint32 * 5)()
    lib.crh_xorwow_state_after_normal2(ULL(seed), ULL(subseq), n, C.byref(d), v)
    return [d.value] + list(v)


def normal2(kind, seed, subseq, n):
    out = (C.c_float * (2 * n))()
    lib.crh_normal2(kind, ULL(seed), ULL(subseq), n, out)
    return [float(x) for x in out]


def poisson(kind, seed, subseq, lam, n):
    out = (C.c_uint * n)()
    lib.crh_poisson(kind, ULL(seed), ULL(subseq), C.c_double(lam), n, out)
    return list(out)


def mixed(kind, seed, subseq, n):
    out = (C.c_double * n)()
    lib.crh_mixed(kind, ULL(seed), ULL(subseq), n, out)
    return [float(x) for x in out]


g = {"generator": "cuRAND 10.3.10 headers (CUDA 12.9) compiled for host, oracle/curand_host.cpp",
     "xorwow_init": [], "mrg_init": [], "u32": [], "xorwow_after_normal2": [], "normal2": [], "poisson": [], "mixed": []}
for seed in SEEDS:
    for ss in SUBSEQS:
        g["xorwow_init"].append({"seed": seed, "subseq": ss, "offset": 0, "state": xorwow_init(seed, ss, 0)})
    for off in [1, 2, 3, 4, 5, 1000, 2**20 + 3, 2**40 + 11]:
        g["xorwow_init"].append({"seed": seed, "subseq": 7, "offset": off, "state": xorwow_init(seed, 7, off)})
for seed in SEEDS:
    for ss in [0, 1, 2, 5, 262143, 2**24 + 12345, 2**33 + 7, 2**50 + 3]:
        for off in [0, 1, 7, 2**40 + 11]:
            g["mrg_init"].append({"seed": seed, "subseq": ss, "offset": off, "state": mrg_init(seed, ss, off)})
for kind in (0, 1, 2):
    for seed in SEEDS[:2]:
        for ss in [0, 1, 5, 262143, 2**24 + 12345, 2**33 + 7]:
            for off in [0, 1, 2, 3, 6, 4001]:
                g["u32"].append({"kind": kind, "seed": seed, "subseq": ss, "offset": off,
                                 "out": u32(kind, seed, ss, off, 12)})
for n in [1, 10, 1000]:
    for ss in [0, 1, 262143]:
        g["xorwow_after_normal2"].append({"seed": 1234, "subseq": ss, "n": n, "state": state_after(1234, ss, n)})
for kind in (0, 1, 2):
    for ss in [0, 1, 77]:
        g["normal2"].append({"kind": kind, "seed": 1234, "subseq": ss, "out": normal2(kind, 1234, ss, 16)})
for kind in (0, 1, 2):
    for lam in [0.5, 7.25, 63.9, 64.0, 150.0, 2222.2, 3999.0, 4000.5, 25000.0]:
        g["poisson"].append({"kind": kind, "seed": 1234, "subseq": 3, "lambda": lam,
                             "out": poisson(kind, 1234, 3, lam, 64)})
for kind in (0, 1, 2):
    g["mixed"].append({"kind": kind, "seed": 99, "subseq": 11, "out": mixed(kind, 99, 11, 60)})

with open(os.path.join(HERE, "curand_host.json"), "w") as f:
    json.dump(g, f, separators=(",", ":"))
print("wrote", os.path.join(HERE, "curand_host.json"))
