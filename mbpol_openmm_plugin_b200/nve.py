"""Velocity-Verlet NVE drivers for the 2B+3B path: `run` is the host loop (forces from any callable, so that the same loop
drives the CUDA path and the CPU oracle), `run_device` the on-device integrator of the C ABI (mbpol_nve_run: positions,
velocities, forces and the energy log never leave HBM during the run).

Velocity-Verlet NVE driver for the 2B+3B path (BASELINE config 2: water256_bulk, 100 steps, dt = 0.2 fs as in
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


MSITE_WEIGHTS = (0.573293118, 0.213353441, 0.213353441)          # python/mbpol.xml:14 (average3)


def run_device(context, pos, vel, m, n_steps, dt_ps=0.0002, msite_index=None):
    """On-device NVE: `context` is an MBPolContext, pos / vel float64 CUDA tensors (n_particles x 3, updated IN PLACE), m a
    CUDA tensor of n_particles masses (0 = virtual site).  msite_index: per-molecule particle index of the M site (placed
    after every drift) or None.  Returns the (n_steps + 1) x 2 array of (E_pot, E_kin)."""
    import ctypes as C

    import torch

    from . import _lib
    from .api import DO_1B, MBPolOneBodyForce
    L = _lib.load()
    assert pos.is_cuda and vel.is_cuda and m.is_cuda and pos.dtype == vel.dtype == m.dtype == torch.float64
    assert pos.is_contiguous() and vel.is_contiguous() and m.numel() == context.n_particles
    flags = (context._flags & 3) | (_lib.NVE_ONEBODY if context._flags & DO_1B else 0)
    cfg = _lib.NveConfig()
    cfg.dt_ps = float(dt_ps)
    cfg.flags = flags
    cfg.onebody_periodic = int(context._f1 is not None and context._f1.getNonbondedMethod() == MBPolOneBodyForce.Periodic)
    keep = None
    if msite_index is not None:
        keep = np.ascontiguousarray(msite_index, dtype=np.int32)
        cfg.msite_index = keep.ctypes.data_as(C.POINTER(C.c_int32))
    for i, w in enumerate(MSITE_WEIGHTS):
        cfg.msite_weights[i] = w
    log = np.zeros((n_steps + 1, 2))
    _lib.check(L.mbpol_nve_run(context.engine.h, pos.data_ptr(), vel.data_ptr(), m.contiguous().data_ptr(), C.byref(cfg),
                               int(n_steps), log.ctypes.data, torch.cuda.current_stream().cuda_stream), context.engine.h)
    return log
