"""The C++ host mirror (include/moihgp_b200/moihgp.hpp): compile check on CPU, behaviour against the oracle on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
CPP = os.path.join(ROOT, "tests", "cpp")


def test_header_compiles_with_reference_vector_type():
    """RegressionObjective / OnlineObjective instantiate with an Eigen::VectorXd-shaped type: the reference's functor
    signature (moihgp_regression.h:34, moihgp_online.h:40)."""
    subprocess.check_call(["g++", "-std=c++11", "-fsyntax-only", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "oracle", "eigen_shim"),
                           "-I" + os.path.join(ROOT, "include"), os.path.join(CPP, "compile_with_eigen_types.cpp")])


def test_header_is_plain_c_abi_only():
    """The host mirror depends on nothing but the C ABI header (no torch, no CUDA headers, no oracle)."""
    src = open(os.path.join(ROOT, "include", "moihgp_b200", "moihgp.hpp")).read()
    includes = [l.split()[1] for l in src.splitlines() if l.startswith("#include")]
    assert set(includes) <= {"<cstddef>", "<list>", "<stdexcept>", "<string>", "<vector>", '"../moihgp_b200.h"'}, includes


@pytest.mark.gpu
def test_cpp_host_api_matches_oracle(cuda_lib):
    exe = os.path.join(CPP, "_build", "test_host_api")
    subprocess.check_call(["make", "-C", CPP], stdout=subprocess.DEVNULL)      # no-op when the binary built by build() is current
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0 and r.stdout.strip().endswith("OK")
