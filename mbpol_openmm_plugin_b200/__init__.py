"""mbpol_openmm_plugin_b200 -- B200-native (sm_100a, fp64) MB-pol 2B+3B many-body path behind the plugin's own API.

See DESIGN.md for the path and its boundary; include/mbpol_b200.h for the C ABI; INTEGRATION.md for the OpenMM side.
"""
from .api import (B200CalcCustomDispersionForceKernel, CustomDispersionForce, B200CalcMBPolOneBodyForceKernel, B200CalcMBPolThreeBodyForceKernel, B200CalcMBPolTwoBodyForceKernel, MBPolContext, MBPolError,
                  MBPolOneBodyForce, MBPolThreeBodyForce, MBPolTwoBodyForce, eval_tuples, measure_fp64_peak, pip_eval)

__all__ = ["CustomDispersionForce", "B200CalcCustomDispersionForceKernel", "MBPolOneBodyForce", "B200CalcMBPolOneBodyForceKernel", "MBPolTwoBodyForce", "MBPolThreeBodyForce", "MBPolContext", "MBPolError",
           "B200CalcMBPolTwoBodyForceKernel", "B200CalcMBPolThreeBodyForceKernel", "pip_eval", "eval_tuples",
           "measure_fp64_peak"]
