"""Builds and runs the C++ host-side mirror (include/ruvgnn.hpp) test program, which reads like
the reference's own unit tests (tests/cpp/test_layer.cpp)."""
import subprocess
from pathlib import Path

import pytest

from ruvector_b200 import _lib

ROOT = Path(__file__).resolve().parents[1]


def build_exe(tmp_path):
    exe = tmp_path / "test_layer"
    lib = _lib.lib_path()
    cmd = ["g++", "-std=c++17", "-O1", "-I", str(ROOT / "include"), str(ROOT / "tests" / "cpp" / "test_layer.cpp"),
           "-o", str(exe), str(lib), f"-Wl,-rpath,{lib.parent}"]
    subprocess.run(cmd, check=True)
    return exe


def test_cpp_constructor_and_error_behaviour(built, tmp_path):
    res = subprocess.run([str(build_exe(tmp_path)), "config"], capture_output=True, text=True, timeout=120)
    print(res.stdout, res.stderr)
    assert res.returncode == 0 and res.stdout.strip().endswith("OK")


@pytest.mark.gpu
def test_cpp_forward_tests(built, tmp_path):
    res = subprocess.run([str(build_exe(tmp_path)), "all"], capture_output=True, text=True, timeout=300)
    print(res.stdout, res.stderr)
    assert res.returncode == 0 and res.stdout.strip().endswith("OK")


def build_c_exe(tmp_path):
    exe = tmp_path / "test_abi"
    lib = _lib.lib_path()
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", str(ROOT / "include"),
                    str(ROOT / "tests" / "c" / "test_abi.c"), "-o", str(exe), str(lib), f"-Wl,-rpath,{lib.parent}", "-lm"],
                   check=True)
    return exe


def test_plain_c_caller_of_the_abi(built, tmp_path):
    """tests/c/test_abi.c: the boundary as a cgo / Rust FFI / JNI binding sees it (C99, no C++)."""
    res = subprocess.run([str(build_c_exe(tmp_path))], capture_output=True, text=True, timeout=120)
    print(res.stdout, res.stderr)
    assert res.returncode == 0 and res.stdout.strip().endswith("OK")


@pytest.mark.gpu
def test_plain_c_caller_forward(built, tmp_path):
    res = subprocess.run([str(build_c_exe(tmp_path)), "gpu"], capture_output=True, text=True, timeout=300)
    print(res.stdout, res.stderr)
    assert res.returncode == 0 and res.stdout.strip().endswith("OK")


def build_c_dist_exe(tmp_path):
    exe = tmp_path / "test_dist"
    lib = _lib.lib_path()
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", str(ROOT / "include"),
                    str(ROOT / "tests" / "c" / "test_dist.c"), "-o", str(exe), str(lib), f"-Wl,-rpath,{lib.parent}",
                    "-lm", "-lpthread"], check=True)
    return exe


def test_plain_c_dist_caller_builds_and_reports_missing_devices(built, tmp_path):
    """tests/c/test_dist.c links against the dist entries of the ABI; without (enough) devices it
    says so and exits 77 instead of computing anything on the host."""
    res = subprocess.run([str(build_c_dist_exe(tmp_path))], capture_output=True, text=True, timeout=120)
    print(res.stdout, res.stderr)
    assert res.returncode in (0, 77)


@pytest.mark.gpu
def test_plain_c_dist_caller_two_gpus_one_process(built, tmp_path):
    """One process, one host thread per GPU (ncclCommInitAll inside the library), both data planes,
    exact mode bit-identical to the single-GPU run."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    res = subprocess.run([str(build_c_dist_exe(tmp_path)), "2"], capture_output=True, text=True, timeout=300)
    print(res.stdout, res.stderr)
    assert res.returncode == 0 and res.stdout.strip().endswith("OK")
