"""Velocity-Verlet NVE driver for the 2B+3B path (BASELINE config 2: water256_bulk, 100 steps, dt = 0.2 fs as in
python/example_ini/mbpol_bulk_nve.ini:11; masses of python/mbpol.xml:3-4).  The integrator is plumbing (a few torch /
numpy element-wise ops); forces come from a callable so the same loop drives the CUDA path and the CPU oracle.
OpenMM units: nm, ps, amu, kJ/mol  ->  a = F / m directly."""
from __future__ import annotations

import numpy as np

MASS_O, MASS_H = 15.99943, 1.007947
KB = 0.0083144626          # kJ/mol/K


def masses(n_mol):
    """(3 n_mol, 1) numpy array of atomic masses in amu (O, H, H per molecule)"""
    return np.tile([MASS_O, MASS_H, MASS_H], n_mol)[:, None].copy()


def maxwell_boltzmann(n_mol, temperature=300.0, seed=1234):
    rng = np.random.default_rng(seed)
    m = np.tile([MASS_O, MASS_H, MASS_H], n_mol)[:, None]
    v = rng.normal(size=(3 * n_mol, 3)) * np.sqrt(KB * temperature / m)
    v -= (v * m).sum(axis=0) / m.sum()                      # zero total momentum
    return v


def run(force_fn, pos, vel, m, n_steps, dt_ps=0.0002):
    """force_fn(pos) -> (potential energy, forces).  Returns final pos, vel and the list of (E_pot, E_kin) per step."""
    e, f = force_fn(pos)
    log = [(float(e), float(0.5 * (m * vel * vel).sum()))]
    for _ in range(n_steps):
        vel = vel + (0.5 * dt_ps) * f / m
        pos = pos + dt_ps * vel
        e, f = force_fn(pos)
        vel = vel + (0.5 * dt_ps) * f / m
        log.append((float(e), float(0.5 * (m * vel * vel).sum())))
    return pos, vel, log
