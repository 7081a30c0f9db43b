"""Host-side mirror of the plugin API: same names, defaults and error behaviour as the reference's Force classes."""
import pytest

from mbpol_openmm_plugin_b200 import (B200CalcMBPolThreeBodyForceKernel, B200CalcMBPolTwoBodyForceKernel, MBPolError,
                                      MBPolThreeBodyForce, MBPolTwoBodyForce)
from mbpol_openmm_plugin_b200 import api


@pytest.mark.parametrize("cls", [MBPolTwoBodyForce, MBPolThreeBodyForce])
def test_force_defaults_and_accessors(cls):
    f = cls()
    # openmmapi/src/MBPolTwoBodyForce.cpp:43 / MBPolThreebodyForce.cpp:42
    assert f.getNonbondedMethod() == cls.CutoffNonPeriodic == 2
    assert f.getCutoff() == 1.0e10
    assert (cls.NoCutoff, cls.CutoffPeriodic) == (0, 1)
    assert f.addParticle([0, 1, 2]) == 0
    assert f.addParticle([4, 5, 6]) == 1
    assert f.getNumMolecules() == f.getNumParticles() == 2
    assert f.getParticleParameters(1) == [4, 5, 6]
    f.setParticleParameters(1, [8, 9, 10])
    assert f.getParticleParameters(1) == [8, 9, 10]
    f.setCutoff(0.65)
    f.setNonbondedMethod(cls.CutoffPeriodic)
    assert (f.getCutoff(), f.getNonbondedMethod()) == (0.65, 1)
    with pytest.raises(ValueError):
        f.setNonbondedMethod(7)


def test_kernel_names():
    # openmmapi/include/openmm/mbpolKernels.h:137-139, 176-178
    assert B200CalcMBPolTwoBodyForceKernel.Name() == "CalcMBPolTwoBodyForce"
    assert B200CalcMBPolThreeBodyForceKernel.Name() == "CalcMBPolThreeBodyForce"


def test_cutoff_larger_than_half_box_is_rejected_at_context_creation():
    # MBPolTwoBodyForceImpl.cpp:60-66 -- checked before any device work
    f = MBPolTwoBodyForce()
    f.setNonbondedMethod(f.CutoffPeriodic)
    f.setCutoff(0.65)
    f.addParticle([0, 1, 2])
    f.addParticle([3, 4, 5])
    with pytest.raises(MBPolError) as e:
        api.MBPolContext(6, [f], box_nm=[1.2, 1.2, 1.2])
    assert e.value.code == -3
    with pytest.raises(MBPolError):
        api.MBPolContext(6, [f])            # periodic without a box
    with pytest.raises(MBPolError):
        api.MBPolContext(6, [])             # no forces


def test_mismatched_molecule_lists_rejected():
    f2, f3 = MBPolTwoBodyForce(), MBPolThreeBodyForce()
    f2.addParticle([0, 1, 2])
    f3.addParticle([3, 4, 5])
    with pytest.raises(MBPolError):
        api.MBPolContext(6, [f2, f3])
