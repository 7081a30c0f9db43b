"""BASELINE config 2: water256_bulk, periodic, 2B+3B only, 100-step velocity-Verlet NVE (dt 0.2 fs, 300 K, seed 1234):
the CUDA path and the CPU oracle are driven through the same integrator loop and must stay on the same trajectory."""
import numpy as np
import pytest

from conftest import load_golden, make_context

pytestmark = pytest.mark.gpu


def test_nve_100_steps_tracks_the_oracle(checker):
    import torch
    from oracle.refbind import canonical
    from mbpol_openmm_plugin_b200 import nve
    g = load_golden("water256_bulk")
    n, box = len(g["pos_mol"]), g["box"]
    steps = 100 if checker.kind == "reference" else 4        # the C port needs ~7 s per evaluation
    pos0 = g["pos_mol"].reshape(-1, 3).copy()
    vel0 = nve.maxwell_boltzmann(n)
    idx = np.arange(3 * n).reshape(n, 3)

    ctx = make_context(n, box, True)

    def gpu_force(p):
        ctx.setPositions(p)
        e = ctx.getPotentialEnergy()
        return e, ctx.getForces()

    def cpu_force(p):
        posO = np.ascontiguousarray(p[0::3])
        pairs = canonical(checker.pairs(posO, box, True, 0.65))        # lists rebuilt every step, like the reference
        trips = canonical(checker.triplets(posO, box, True, 0.45))
        e2, f2 = checker.eval_2b(p, idx, pairs, box, True)[:2]
        e3, f3 = checker.eval_3b(p, idx, trips, box, True)[:2]
        return e2 + e3, f2 + f3

    pg, vg, log_g = nve.run(gpu_force, torch.as_tensor(pos0).cuda(), torch.as_tensor(vel0).cuda(),
                            torch.as_tensor(nve.masses(n)).cuda(), steps)
    pc, vc, log_c = nve.run(cpu_force, pos0, vel0, nve.masses(n), steps)
    log_g, log_c = np.array(log_g), np.array(log_c)
    assert np.abs(pg.cpu().numpy() - pc).max() < 1e-10                  # nm, after `steps` steps
    assert np.abs(vg.cpu().numpy() - vc).max() < 1e-8                   # nm/ps
    assert np.abs(log_g[:, 0] - log_c[:, 0]).max() <= 1e-9 * np.abs(log_c[:, 0]).max()
    # Total energy is NOT asserted to be conserved: the Reference platform's chain-triplet list drops a trimer the moment
    # d(I,J) or d(J,K) crosses 4.5 A even when the other two switches are still non-zero (SURVEY.md finding 2), so the
    # reference's own 2B+3B surface has small discontinuities.  What must hold is that both implementations see the same
    # total energy along the same trajectory.
    assert np.abs(log_g.sum(axis=1) - log_c.sum(axis=1)).max() <= 1e-9 * np.abs(log_c.sum(axis=1)).max()


def test_device_integrator_tracks_the_oracle(checker):
    """mbpol_nve_run: the same 100 steps with positions, velocities, forces and the energy log resident on the device
    (O/H/H/M particles, M sites placed after every drift) against the oracle-driven host loop."""
    import torch
    from oracle.refbind import canonical
    from mbpol_openmm_plugin_b200 import nve
    from mbpol_openmm_plugin_b200.systems import with_virtual_sites
    g = load_golden("water256_bulk")
    n, box = len(g["pos_mol"]), g["box"]
    steps = 100 if checker.kind == "reference" else 4
    pos0 = g["pos_mol"].reshape(-1, 3).copy()
    vel0 = nve.maxwell_boltzmann(n)
    idx = np.arange(3 * n).reshape(n, 3)

    def cpu_force(p):
        posO = np.ascontiguousarray(p[0::3])
        pairs = canonical(checker.pairs(posO, box, True, 0.65))
        trips = canonical(checker.triplets(posO, box, True, 0.45))
        e2, f2 = checker.eval_2b(p, idx, pairs, box, True)[:2]
        e3, f3 = checker.eval_3b(p, idx, trips, box, True)[:2]
        return e2 + e3, f2 + f3

    pc, vc, log_c = nve.run(cpu_force, pos0, vel0, nve.masses(n), steps)
    log_c = np.array(log_c)
    for skin in (0.0, 0.05):
        ctx = make_context(n, box, True, stride=4)
        ctx.engine.set_skin(skin)
        pos4 = torch.as_tensor(with_virtual_sites(g["pos_mol"]).reshape(-1, 3)).cuda()
        pos4[3::4] = 0.0                                              # M sites start anywhere: the integrator places them
        vel4 = torch.zeros_like(pos4)
        vel4.view(n, 4, 3)[:, :3] = torch.as_tensor(vel0.reshape(n, 3, 3)).cuda()
        m4 = torch.zeros(4 * n, dtype=torch.float64, device="cuda")
        m4.view(n, 4)[:, :3] = torch.as_tensor(nve.masses(n).reshape(n, 3)).cuda()
        log_g = nve.run_device(ctx, pos4, vel4, m4, steps, msite_index=np.arange(n) * 4 + 3)
        pg = pos4.view(n, 4, 3).cpu().numpy()
        assert np.abs(pg[:, :3].reshape(-1, 3) - pc).max() < 1e-10                       # nm, after `steps` steps
        assert np.abs(vel4.view(n, 4, 3)[:, :3].reshape(-1, 3).cpu().numpy() - vc).max() < 1e-8
        want_m = sum(w * pg[:, k] for k, w in enumerate(nve.MSITE_WEIGHTS))
        assert np.abs(pg[:, 3] - want_m).max() < 1e-14 and np.all(vel4.view(n, 4, 3)[:, 3].cpu().numpy() == 0.0)
        assert np.abs(log_g[:, 0] - log_c[:, 0]).max() <= 1e-9 * np.abs(log_c[:, 0]).max()
        assert np.abs(log_g.sum(axis=1) - log_c.sum(axis=1)).max() <= 1e-9 * np.abs(log_c.sum(axis=1)).max()
        st = ctx.engine.stats()
        assert (st["n_row_builds"] == steps + 1) if skin == 0.0 else (st["n_row_builds"] < steps // 2)
