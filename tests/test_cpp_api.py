"""The C++ method API is source-compatible with reference-style user code (reference README.md:60-93) and links
with plain g++ (no nvcc, no curand headers on the user side)."""
import os
import subprocess
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

USER_CODE = textwrap.dedent(r"""
    #include "NMCH/methods/NMCH_FE.hpp"
    #include "NMCH/methods/NMCH_EM.hpp"
    #include <cstring>
    using namespace nmch::methods;

    template <typename M> static int drive(int NTPB, int NB, int N)
    {
        M nmch(NTPB, NB, 1.0f, 1.0f, 0.1f, 0.0f, 0.5f, -0.7f, 0.1f, 0.3f, N);
        nmch.init(1234ULL);
        nmch.compute();
        nmch.set_k(1.0f); nmch.set_theta(0.2f); nmch.set_sigma(0.4f);
        nmch.compute();
        nmch.print_stats();
        float e = nmch.get_strike_price(), e2 = nmch.get_price_squared(), t = nmch.get_execution_time(), err = nmch.get_err();
        nmch.finalize();
        return (e > 0 && e2 > 0 && t >= 0 && err > 0) ? 0 : 3;
    }

    int main(int argc, char **argv)
    {
        if (argc > 1 && !strcmp(argv[1], "types")) return 0;      // compile/link check only
        int rc = 0;
        rc |= drive<NMCH_FE_K1_MM<curandStateXORWOW_t>>(128, 8, 50);
        rc |= drive<NMCH_FE_K2_MM<curandStatePhilox4_32_10_t>>(128, 8, 50);
        rc |= drive<NMCH_FE_K3_MM<curandStateMRG32k3a_t>>(128, 8, 50);
        rc |= drive<NMCH_FE_K2_PHILOX_MM>(128, 8, 50);
        rc |= drive<NMCH_FE_K1_PgM<curandStateXORWOW_t>>(128, 8, 50);
        rc |= drive<NMCH_FE_K1_PiM<curandStateXORWOW_t>>(128, 8, 50);
        rc |= drive<NMCH_EM_K1_MM<curandStateXORWOW_t>>(128, 8, 50);
        rc |= drive<NMCH_EM_K2_MM<curandStatePhilox4_32_10_t>>(128, 8, 50);
        rc |= drive<NMCH_EM_K3_MM<curandStateXORWOW_t>>(128, 8, 50);
        return rc;
    }
""")


@pytest.fixture(scope="module")
def user_exe(tmp_path_factory):
    from nmch_b200 import _build
    _build.build(only_if_missing=True)
    d = tmp_path_factory.mktemp("user")
    src = d / "user.cpp"
    src.write_text(USER_CODE)
    exe = d / "user"
    api = [os.path.join(ROOT, "src", "NMCH", "methods", f) for f in ("NMCH.cpp", "NMCH_FE.cpp", "NMCH_EM.cpp")]
    api.append(os.path.join(ROOT, "src", "NMCH", "utils", "utils.cpp"))
    pkg = os.path.join(ROOT, "nmch_b200")
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), str(src), *api, "-o",
                    str(exe), "-L", pkg, "-lnmch_b200", f"-Wl,-rpath,{pkg}"], check=True)
    return str(exe)


def test_reference_style_user_code_compiles_and_links(user_exe):
    assert subprocess.run([user_exe, "types"]).returncode == 0


@pytest.mark.gpu
def test_all_reference_class_names_run(user_exe):
    r = subprocess.run([user_exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    assert r.stdout.count("METHOD: FORWARD-EULER") == 6 and r.stdout.count("METHOD: EXACT-METHOD") == 3


def test_heston_pricer_in_utils_matches_oracle(tmp_path):
    from oracle import oracle as o
    src = tmp_path / "h.cpp"
    src.write_text('#include "NMCH/utils/utils.hpp"\n#include <cstdio>\nint main(){printf("%.12f %.12f %.9f\\n",'
                   'nmch::utils::heston_call(1,1,0.1,0,0.5,0.1,0.3,-0.7,1), nmch::utils::heston_call(1,1,0.1,0,2.08,0.108,1.0,-0.7,1),'
                   'nmch::utils::NP(0.15));}')
    exe = tmp_path / "h"
    pkg = os.path.join(ROOT, "nmch_b200")
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src),
                    os.path.join(ROOT, "src", "NMCH", "utils", "utils.cpp"), "-o", str(exe), "-L", pkg, "-lnmch_b200",
                    f"-Wl,-rpath,{pkg}"], check=True)
    a, b, c = map(float, subprocess.run([str(exe)], capture_output=True, text=True).stdout.split())
    assert abs(a - 0.1197325094) < 2e-9 and abs(b - 0.1104934558) < 2e-9
    assert abs(c - o.lib().orc_NP(0.15)) < 1e-9          # printed with 9 decimals
