"""TEST INFRASTRUCTURE ONLY -- CPU restatement of MB-pol's TT-damped C6 dispersion.

The reference does not implement this term itself: python/mbpol.xml:37-108 builds a stock OpenMM CustomNonbondedForce

    E = sum over atom pairs (i < j, different residues, no M site) within the cutoff of   -C6 tt6(r) / r^6
    tt6(r) = 1 - exp(-d6 r) sum_{k=0..6} (d6 r)^k / k!                 (Tang-Toennies damping)

with the C6 / d6 tables of mbpol.xml:53-82 (kJ/mol nm^6, nm^-1), CutoffNonPeriodic / CutoffPeriodic on the ATOM-ATOM
distance (minimum image per atom pair), no switching function and no long-range correction (OpenMM defaults).  OpenMM is
a third-party dependency that is absent here, so its published expression semantics are restated in numpy
and pinned to the reference's own known answer: python/tests/TestCustomDispersion.py:14 (water3.pdb, -6.84471477
kcal/mol, tolerance 0.01) -- tests/test_dispersion.py."""
from __future__ import annotations

import numpy as np

# mbpol.xml:53-82, atom classes O = 0, H = 1
C6 = np.array([[9.92951990e-4, 3.49345451e-4], [3.49345451e-4, 8.40715638e-5]])
D6 = np.array([[9.29548582e+01, 9.77520243e+01], [9.77520243e+01, 9.40647517e+01]])
FACT = np.array([1.0, 1.0, 2.0, 6.0, 24.0, 120.0, 720.0])


def pair_terms(r, c6, d6):
    """energy and dE/dr of one atom pair class at distances r (nm)"""
    x = d6 * r
    s = sum(x ** k / FACT[k] for k in range(7))
    ex = np.exp(-x)
    tt = 1.0 - s * ex
    dtt = d6 * ex * x ** 6 / 720.0
    r6 = r ** 6
    return -c6 * tt / r6, -c6 * (dtt / r6 - 6.0 * tt / (r6 * r))


def eval_dispersion(pos, idx, box=None, pbc=False, cutoff=1e10):
    """pos: (n_particles, 3) nm; idx: (n_mol, 3) particle indices {O, H1, H2}.  Returns E (kJ/mol), F (kJ/mol/nm)."""
    pos = np.asarray(pos, dtype=np.float64)
    idx = np.asarray(idx)
    n = len(idx)
    atoms = idx.reshape(-1)
    cls = np.tile([0, 1, 1], n)
    mol = np.repeat(np.arange(n), 3)
    p = pos[atoms]
    f = np.zeros_like(p)
    e = 0.0
    box = None if box is None else np.asarray(box, dtype=np.float64)
    for a in range(len(atoms)):
        sel = np.nonzero(mol > mol[a])[0]                    # every pair of atoms of different molecules once
        if len(sel) == 0:
            continue
        d = p[sel] - p[a]
        if pbc:
            d -= box * np.floor(d / box + 0.5)
        r = np.sqrt((d * d).sum(axis=1))
        keep = r <= cutoff
        sel, d, r = sel[keep], d[keep], r[keep]
        c6, d6 = C6[cls[a], cls[sel]], D6[cls[a], cls[sel]]
        en, de = pair_terms(r, c6, d6)
        e += en.sum()
        g = (de / r)[:, None] * d                            # dE/d(p_j)
        f[sel] -= g
        f[a] += g.sum(axis=0)
    out = np.zeros_like(pos)
    np.add.at(out, atoms, f)
    return float(e), out
