"""include/zk_impl.hpp - the C++ host side of the drop-in (the crate's API over the C ABI) - driven by
the reference's own #[test] functions restated in tests/cpp/reference_tests.cpp.

* not gpu: the test program is built twice - (a) against the real libzkb200.so only: it must link
  (every entry point the mirror binds exists) and its host-only groups must pass; (b) with
  tests/cpp/oracle_backend.cpp interposing the GPU entry points with the C oracle, so the mirror's
  HOST logic (packing, error mapping, proof-stream assembly, FRI::verify) is exercised without a GPU.
* gpu: build (a) runs everything on the B200.
In both, the FRI proof the C++ mirror produces for the reference's fri::tests::verify input must equal
the Python oracle's proof byte for byte (compared through BLAKE2b-512 of the proof)."""
import hashlib
import os
import re
import subprocess

import pytest

from oracle import cbind, field as F, fastfri, proof_stream as PS
from oracle.fri import FRI as OFRI

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp")
LIBDIR = os.path.join(ROOT, "zk_stark_tutor_b200", "lib")
ORADIR = os.path.join(ROOT, "oracle")


def _build(tmp_path, with_oracle_backend):
    assert os.path.exists(os.path.join(LIBDIR, "libzkb200.so")), "libzkb200.so is not built (run __graft_entry__.build())"
    exe = str(tmp_path / ("reference_tests_cpu" if with_oracle_backend else "reference_tests"))
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(SRC, "reference_tests.cpp")]
    if with_oracle_backend:
        cbind.build()
        cmd += [os.path.join(SRC, "oracle_backend.cpp"), "-L" + ORADIR, "-lzkoracle", "-Wl,-rpath," + ORADIR]
    cmd += ["-o", exe, "-L" + LIBDIR, "-lzkb200", "-Wl,-rpath," + LIBDIR]
    subprocess.check_call(cmd)
    return exe


def _run(exe, mode):
    out = subprocess.run([exe, mode], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "test result: ok" in out.stdout, out.stdout
    assert "FAILED" not in out.stdout
    return out.stdout


def _oracle_fri_proof_hash():
    """fri.rs:450-500: degree-63 polynomial with coefficients 0..63 on the 256-point domain <omega>, ef 4, 17 tests."""
    n, ef, ncc = 256, 4, 17
    w = F.primitive_nth_root(n)
    coeffs = cbind.to_arr(list(range(64)))
    cw = cbind.coset_lde(w, n, 1, coeffs)                  # offset 1: plain evaluation on <omega>, as the reference's test does
    ps = PS.IndependentProofStream()
    fastfri.prove(OFRI(F.GENERATOR, w, n, ef, ncc), cw, ps)
    proof = ps.digest()
    return len(proof), hashlib.blake2b(proof).hexdigest()


def _check_proof(stdout):
    m = re.search(r"fri::tests::verify proof: (\d+) bytes, blake2b512 = ([0-9a-f]{128})", stdout)
    assert m, stdout
    size, digest = _oracle_fri_proof_hash()
    assert (int(m.group(1)), m.group(2)) == (size, digest), "C++ mirror's FRI proof differs from the oracle's"


def test_cpp_mirror_links_and_host_groups_pass(tmp_path):
    out = _run(_build(tmp_path, False), "host")
    assert "17 passed; 0 failed" in out


def test_cpp_mirror_host_logic_on_oracle_backend(tmp_path):
    out = _run(_build(tmp_path, True), "all")
    assert "34 passed; 0 failed" in out
    _check_proof(out)


@pytest.mark.gpu
def test_cpp_mirror_reference_tests_on_gpu(tmp_path):
    out = _run(_build(tmp_path, False), "all")
    assert "34 passed; 0 failed" in out
    _check_proof(out)


def _build_example(tmp_path, with_oracle_backend):
    exe = str(tmp_path / "lde_fri")
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "lde_fri.cpp")]
    if with_oracle_backend:
        cbind.build()
        cmd += [os.path.join(SRC, "oracle_backend.cpp"), "-L" + ORADIR, "-lzkoracle", "-Wl,-rpath," + ORADIR]
    cmd += ["-o", exe, "-L" + LIBDIR, "-lzkb200", "-Wl,-rpath," + LIBDIR]
    subprocess.check_call(cmd)
    return exe


def test_cpp_example_on_oracle_backend(tmp_path):
    """examples/lde_fri.cpp (LDE -> commit -> open -> FRI prove -> verify through the C++ mirror) builds and accepts its own proof"""
    out = subprocess.run([_build_example(tmp_path, True), "10"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "verify: ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_example_on_gpu(tmp_path):
    out = subprocess.run([_build_example(tmp_path, False), "16"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "verify: ok" in out.stdout, out.stdout + out.stderr
