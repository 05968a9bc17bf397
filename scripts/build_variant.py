"""Tuning builds: recompile ONE source of the library with extra -D flags and link a variant library of the same ABI
(nmch_b200/variants/libnmch_b200_<tag>.so; load it with NMCH_B200_LIB=<path>).  Not product code.
    python scripts/build_variant.py em_kernels.cu minb8 -DNMCHB_EM_MINB=8"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nmch_b200 import _build  # noqa: E402


def main():
    src, tag, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
    _build.build()
    out_dir = os.path.join(_build.PKG, "variants")
    os.makedirs(out_dir, exist_ok=True)
    obj = os.path.join(out_dir, f"{src.replace('.cu', '')}_{tag}.o")
    subprocess.run([_build.NVCC, *_build.ARCH, *_build.COMMON, *flags, "-Xptxas", "-v", "-I", os.path.join(ROOT, "include"),
                    "-c", os.path.join(_build.CSRC, src), "-o", obj], check=True)
    objs = [obj if s == src else os.path.join(_build.PKG, "build", s.replace(".cu", ".o")) for s in _build.CU_SOURCES]
    lib = os.path.join(out_dir, f"libnmch_b200_{tag}.so")
    subprocess.run([_build.NVCC, *_build.ARCH, "-shared", "-o", lib, *objs, "-ldl"], check=True)
    print(lib)


if __name__ == "__main__":
    main()
