"""Runs csrc/fft2_core.cuh's pass code on the CPU (threads emulated) against a float64 DFT."""

import os
import shutil
import subprocess

import pytest


def test_fft_core_on_host(repo_root, tmp_path):
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    exe = tmp_path / "host_fft_harness"
    subprocess.run(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", str(exe),
                    os.path.join(repo_root, "tests", "host_fft_harness.cu")], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert "OK" in out and "FAIL" not in out
