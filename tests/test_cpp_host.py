"""C++ host layer (mbpol_openmm_plugin_b200/host/mbpol_host.h) and the reference-style C++ known-answer tests."""
import os
import subprocess

import pytest

from conftest import ROOT, has_cuda


def build_exe(tmp_path):
    from mbpol_openmm_plugin_b200 import _lib
    exe = str(tmp_path / "TestB200MBPolForces")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", "-O1",
                           os.path.join(ROOT, "tests", "cpp", "TestB200MBPolForces.cpp"),
                           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "mbpol_openmm_plugin_b200", "host"),
                           "-L" + libdir, "-lmbpol_b200", "-Wl,-rpath," + libdir, "-o", exe])
    return exe


@pytest.mark.skipif(has_cuda(), reason="no-GPU behaviour")
def test_cpp_tests_compile_and_fail_loudly_without_gpu(tmp_path):
    r = subprocess.run([build_exe(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_cpp_known_answer_tests(tmp_path):
    r = subprocess.run([build_exe(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "Done" in r.stdout, r.stdout + r.stderr
