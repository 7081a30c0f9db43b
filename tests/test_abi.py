"""The C-ABI library loads without a GPU and exports every symbol include/mbpol_b200.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, has_cuda


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mbpol_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mbpol_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from mbpol_openmm_plugin_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mbpol_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names


def test_no_libcuda_link_dependency():
    import subprocess
    from mbpol_openmm_plugin_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out          # driver entry points are resolved at run time (pip_module.cu)


def test_work_model_constants():
    from mbpol_openmm_plugin_b200 import _lib
    L = _lib.load()
    assert _lib.FLOP_PER_DIMER == 30601 and _lib.FLOP_PER_TRIMER == 111818
    assert 15000 < L.mbpol_pip_instruction_count(2) < 40000
    assert 60000 < L.mbpol_pip_instruction_count(3) < 100000
    assert L.mbpol_pip_instruction_count(4) == -1


@pytest.mark.skipif(has_cuda(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    """No CPU fallback: creating a context without a CUDA device is an error, not a silent downgrade."""
    from mbpol_openmm_plugin_b200 import MBPolContext, MBPolError, MBPolTwoBodyForce
    f = MBPolTwoBodyForce()
    f.addParticle([0, 1, 2])
    f.addParticle([3, 4, 5])
    with pytest.raises(MBPolError) as e:
        MBPolContext(6, [f])
    assert e.value.code == -2 and "no CUDA device" in str(e.value)


def test_product_does_not_import_oracle():
    """The product path must never route through oracle/ (checked on the sources)."""
    pkg = os.path.join(ROOT, "mbpol_openmm_plugin_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".h", ".cpp")):
                src = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "refbind" not in src and "liboracle" not in src and "libmbpol_ref" not in src, fn


def test_plain_c_client_compiles_and_links(tmp_path):
    """examples/capi_demo.c: the ABI is usable from plain C (no C++/torch types in the signatures)"""
    import subprocess
    from mbpol_openmm_plugin_b200 import _lib
    exe = str(tmp_path / "capi_demo")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", os.path.join(ROOT, "examples", "capi_demo.c"),
                           "-I" + os.path.join(ROOT, "include"), "-L" + libdir, "-lmbpol_b200",
                           "-Wl,-rpath," + libdir, "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    if has_cuda():
        assert r.returncode == 0 and "E_2B = 6.142078" in r.stdout
    else:
        assert r.returncode == 2 and "no CUDA device" in r.stderr


def test_two_gpu_c_client_compiles_and_runs_where_two_gpus_exist(tmp_path):
    """examples/capi_demo_2gpu.c: NCCL collectives behind the C ABI from plain C, one thread per GPU (exit code 77 = fewer
    than two GPUs on this machine; with two it must reproduce the single-GPU energies and forces)"""
    import subprocess
    from mbpol_openmm_plugin_b200 import _lib
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    exe = str(tmp_path / "capi_demo_2gpu")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", os.path.join(ROOT, "examples", "capi_demo_2gpu.c"),
                           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(cuda, "include"), "-L" + libdir,
                           "-lmbpol_b200", "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-lpthread", "-lm",
                           "-Wl,-rpath," + libdir, "-Wl,-rpath," + os.path.join(cuda, "lib64"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode in (0, 77), r.stdout + r.stderr
    if r.returncode == 0:
        assert "Done" in r.stdout
