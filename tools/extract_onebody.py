"""Extract the one-body (Partridge-Schwenke) surface parameters of the reference into a data file.

    python tools/extract_onebody.py [/root/reference] -> mbpol_openmm_plugin_b200/data/onebody.npz

Source: platforms/reference/src/mbpol_interaction_constants.h:4-22 (unit constants), :24-366 (c5zA, cbasis, ccore,
crest: 245 expansion coefficients each), :368-375 (reoh, b1, roh, alphaoh, deohA, phh1A, phh2), :377-424 (idx1..3:
exponents+1 of the 245 terms).  These are published physical parameters (Partridge & Schwenke, JCP 106, 4618 (1997),
as re-fitted for MB-pol); only numbers are taken, the evaluation code is ours (csrc/onebody.cu).
The derived table c5z = f5z*c5zA + fbasis*cbasis + fcore*ccore + frest*crest is evaluated here in IEEE double in
the reference's operation order (MBPolReferenceOneBodyForce.cpp:108-110) so that the GPU kernel starts from bit-identical
coefficients.
"""
import os
import re
import sys

import numpy as np

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
txt = open(os.path.join(ref, "platforms/reference/src/mbpol_interaction_constants.h")).read()
txt = re.sub(r"//[^\n]*", "", txt)


def array(name):
    m = re.search(r"\b%s\[245\]\s*=\s*\{([^}]*)\}" % name, txt)
    vals = [v for v in re.split(r"[\s,]+", m.group(1).strip()) if v]
    assert len(vals) == 245, (name, len(vals))
    return vals


def scalar(name):
    return float(re.search(r"const double %s\s*=\s*([-+0-9.eE]+)\s*;" % name, txt).group(1))


out = {k: np.array([float(v) for v in array(k)]) for k in ("c5zA", "cbasis", "ccore", "crest")}
for k in ("idx1", "idx2", "idx3"):
    out[k] = np.array([int(v) for v in array(k)], dtype=np.int32)
for k in ("reoh", "b1", "roh", "alphaoh", "deohA", "phh1A", "phh2", "Eh_J", "Na", "c0", "h_Js"):
    out[k] = np.float64(scalar(k))
# MBPolReferenceOneBodyForce.cpp:77-80
f5z, fbasis, fcore, frest = 0.999677885, 0.15860145369897, -1.6351695982132, 1.0
c5z = np.empty(245)
for i in range(245):
    c5z[i] = ((f5z * float(out["c5zA"][i]) + fbasis * float(out["cbasis"][i])) + fcore * float(out["ccore"][i])) \
             + frest * float(out["crest"][i])
out["c5z"] = c5z
out["f5z"] = np.float64(f5z)
# mbpol_interaction_constants.h:7-8,20,22
kcal_J = 4184.0
Eh_kcalmol = float(out["Eh_J"]) * float(out["Na"]) / kcal_J
Eh_cm1 = 1.0e-2 * float(out["Eh_J"]) / (float(out["c0"]) * float(out["h_Js"]))
out["cm1_kcalmol"] = np.float64(Eh_kcalmol / Eh_cm1)
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mbpol_openmm_plugin_b200", "data", "onebody.npz")
np.savez_compressed(dst, **out)
print(dst, {k: (v.shape if getattr(v, "shape", ()) else float(v)) for k, v in out.items()})
