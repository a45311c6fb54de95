"""CPU: the prover's host-side 64-bit arithmetic (vpin_b200/csrc/host_fast.hpp — F_p in 5x51 limbs, fixed-base ristretto255
multiplication, RFC 9496 encoding — and the 4x64 Montgomery multiplication in fl.cuh) against the oracle's independent code."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_fast_matches_oracle(tmp_path):
    exe = str(tmp_path / "test_host_fast")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_host_fast.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "host_fast: ok" in out.stdout, out.stdout + out.stderr
