"""Host-side (CPU) check of the __host__ __device__ arithmetic the kernels are built from:
csrc/fe128.cuh (Montgomery field ops) and csrc/blake2b.cuh (leaf encoder, compression)
are compiled with g++ and compared with the C oracle on 200k seeded cases.
The device PTX variants of the same functions are covered by the -m gpu parity tests."""
import os
import subprocess

from oracle import cbind

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_check(tmp_path):
    cbind.build()
    exe = str(tmp_path / "host_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", os.path.join(ROOT, "tests", "host_check.cpp"),
                           "-o", exe, "-L" + os.path.join(ROOT, "oracle"), "-lzkoracle",
                           "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host_check: ok" in out.stdout
