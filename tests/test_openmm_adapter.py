"""OpenMM CUDA-platform adapter (openmm_adapter/B200MBPolKernels.{h,cpp}) compiled against the stand-in OpenMM headers of
openmm_adapter/shim and driven through a fake CudaContext with real device buffers.

CPU: the adapter and its test driver compile and link (and fail loudly without a GPU).
GPU: water256_bulk through plugin registration -> KernelFactory -> initialize -> execute of BOTH kernels, with a shuffled
atom order, O/H/H/M particles, double4 and real4 posq, the padded int64 fixed-point force buffer pre-filled by "other
forces": forces 1e-6 kJ/mol/nm and energy 1e-9 against the reference fixture in double precision; one tuple-list build
per evaluation."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, has_cuda, load_golden

CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build_exe(tmp_path):
    from mbpol_openmm_plugin_b200 import _lib
    exe = str(tmp_path / "TestB200OpenMMAdapter")
    libdir = os.path.dirname(_lib.LIB_PATH)
    ad = os.path.join(ROOT, "openmm_adapter")
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", "-O1",
                           os.path.join(ROOT, "tests", "cpp", "TestB200OpenMMAdapter.cpp"),
                           os.path.join(ad, "B200MBPolKernels.cpp"),
                           "-I" + ad, "-I" + os.path.join(ad, "shim"), "-I" + os.path.join(ROOT, "include"),
                           "-I" + os.path.join(CUDA, "include"), "-L" + libdir, "-lmbpol_b200", "-Wl,-rpath," + libdir,
                           "-L" + os.path.join(CUDA, "lib64"), "-lcudart", "-Wl,-rpath," + os.path.join(CUDA, "lib64"),
                           "-o", exe])
    return exe


def write_input(path, g):
    with open(path, "wb") as f:
        f.write(struct.pack("<ii", len(g["pos_mol"]), int(bool(g["pbc"]))))
        f.write(np.asarray(g["box"], dtype="<f8").tobytes())
        f.write(struct.pack("<dd", float(g["cut2"]), float(g["cut3"])))
        f.write(np.ascontiguousarray(g["pos_mol"], dtype="<f8").tobytes())


def test_adapter_compiles_against_the_shim(tmp_path):
    exe = build_exe(tmp_path)
    assert os.path.exists(exe)
    if not has_cuda():
        g = load_golden("water14_xml")
        write_input(str(tmp_path / "in.bin"), g)
        r = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin"), "double"], capture_output=True, text=True)
        assert r.returncode != 0 and "Done" not in r.stdout          # no GPU: an error, never a silent fallback


@pytest.mark.gpu
@pytest.mark.parametrize("case,precision", [("water256_bulk", "double"), ("water14_xml", "double"), ("water256_bulk", "single")])
def test_adapter_reproduces_the_reference_through_openmm_buffers(case, precision, tmp_path):
    g = load_golden(case)
    n = len(g["pos_mol"])
    exe = build_exe(tmp_path)
    write_input(str(tmp_path / "in.bin"), g)
    r = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin"), precision], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0 and "Done" in r.stdout, r.stdout + r.stderr
    raw = open(str(tmp_path / "out.bin"), "rb").read()
    e = struct.unpack_from("<d", raw, 0)[0]
    f = np.frombuffer(raw, dtype="<f8", count=12 * n, offset=8).reshape(n, 4, 3)
    f3 = np.frombuffer(raw, dtype="<f8", count=12 * n, offset=8 + 96 * n).reshape(n, 4, 3)
    builds, ev1, ev2 = struct.unpack_from("<iii", raw, 8 + 192 * n)
    assert (builds, ev1, ev2) == (1, 1, 2)              # one gather + list build serves both kernels of an evaluation
    assert np.all(f[:, 3] == 0.0) and np.all(f3[:, 3] == 0.0)        # M sites
    want = (g["f2"] + g["f3"]).reshape(n, 3, 3)
    want3 = g["f3"].reshape(n, 3, 3)
    e_want = float(g["e2"]) + 2.0 * float(g["e3"])      # evaluation 1: 2B + 3B, evaluation 2: the 3B group alone
    if precision == "double":
        # the fixed-point buffer quantises to 2^-32 kJ/mol/nm = 2.3e-10
        assert np.abs(f[:, :3] - want).max() < 1e-6 and np.abs(f3[:, :3] - want3).max() < 1e-6
        assert abs(e - e_want) <= 1e-9 * abs(e_want)
    else:
        # real4 positions carry ~1e-7 relative error: the path is exercised, parity is not claimed
        assert np.abs(f[:, :3] - want).max() < 2e-2 * np.abs(want).max()
        assert abs(e - e_want) <= 1e-3 * abs(e_want)
