import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False


def pytest_collection_modifyitems(config, items):
    if has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if "poly_values" not in p)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.fixture(scope="session")
def ref():
    from oracle.refbind import Ref, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return Ref()


@pytest.fixture(scope="session")
def port():
    from oracle.refbind import Port, have_port
    if not have_port():
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
    return Port()


@pytest.fixture(scope="session")
def checker():
    """strongest available oracle: the reference's own code, else the C port"""
    from oracle.refbind import best
    return best()


def make_context(n_mol, box, pbc, cut2=0.65, cut3=0.45, stride=3, which=(2, 3), **kw):
    from mbpol_openmm_plugin_b200 import MBPolContext, MBPolThreeBodyForce, MBPolTwoBodyForce
    from mbpol_openmm_plugin_b200.systems import particle_indices
    forces = []
    for order, cls, cut in ((2, MBPolTwoBodyForce, cut2), (3, MBPolThreeBodyForce, cut3)):
        if order not in which:
            continue
        f = cls()
        f.setCutoff(cut)
        f.setNonbondedMethod(f.CutoffPeriodic if pbc else f.CutoffNonPeriodic)
        for row in particle_indices(n_mol, stride):
            f.addParticle(row.tolist())
        forces.append(f)
    if 1 in which:
        from mbpol_openmm_plugin_b200 import MBPolOneBodyForce
        f = MBPolOneBodyForce()
        f.setNonbondedMethod(f.Periodic if pbc else f.NonPeriodic)
        for row in particle_indices(n_mol, stride):
            f.addOneBody(row.tolist())
        forces.insert(0, f)
    return MBPolContext(n_mol * stride, forces, box_nm=list(box) if pbc else None, **kw)
