"""nohuman_b200/csrc/nh_math.h (the integer building blocks the kernels use)
compiled for the host and checked against the oracle: fmix64, reverse
complement, branch-free ASCII->2-bit packing with ambiguity bits, the exact
invariant-divisor modulo that replaces `hc % capacity`, l-mer extraction."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_math_matches_oracle(oracle, tmp_path):
    exe = str(tmp_path / "host_math_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe,
                           os.path.join(ROOT, "tests", "native", "host_math_check.cc"),
                           "-L", os.path.join(ROOT, "oracle"), "-lk2oracle",
                           "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr
