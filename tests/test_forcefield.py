"""PDB + mbpol.xml front-end: parsing on CPU, evaluation on GPU."""
import os

import numpy as np
import pytest

from conftest import load_golden


def write_pdb(path, pos_mol, box=None):
    from mbpol_openmm_plugin_b200.systems import with_virtual_sites
    p = with_virtual_sites(pos_mol) * 10.0
    with open(path, "w") as f:
        if box is not None:
            f.write("CRYST1%9.3f%9.3f%9.3f  90.00  90.00  90.00 P 1           1\n" % tuple(10 * np.asarray(box)))
        for i, xyz in enumerate(p):
            name = ("O", "H1", "H2", "M")[i % 4]
            f.write("HETATM%5d %-4s HOH  %4d    %8.3f%8.3f%8.3f  1.00  0.00\n" % (i + 1, " " + name, i // 4 + 1, *xyz))
        f.write("END\n")


def test_xml_and_pdb_parsing(tmp_path):
    from mbpol_openmm_plugin_b200 import forcefield as ff
    g = load_golden("water14_xml")
    pdb_path = str(tmp_path / "w14.pdb")
    write_pdb(pdb_path, g["pos_mol"])
    pdb = ff.PDBFile(pdb_path)
    assert len(pdb.atoms) == 56 and pdb.water_indices()[1] == [4, 5, 6]
    field = ff.ForceField()
    assert field.cutoffs == {"MBPolTwoBodyForce": 0.65, "MBPolThreeBodyForce": 0.45}          # python/mbpol.xml:31-36
    f1, f2, f3 = field.createForces(pdb, ff.PME)
    assert field.one_body and f1.getNumOneBodys() == 14 and f1.getNonbondedMethod() == f1.Periodic   # python/mbpol.py:46-49
    assert field.createForces(pdb, ff.NoCutoff)[0].getNonbondedMethod() == f1.NonPeriodic
    assert f2.getNonbondedMethod() == f2.CutoffPeriodic and f3.getCutoff() == 0.45 and f2.getNumMolecules() == 14
    with pytest.raises(ValueError):
        field.createForces(pdb, "Ewald")
    assert not field.dispersion                       # the reduced XML shipped here has no dispersion script
    ref_xml = "/root/reference/python/mbpol.xml"
    if os.path.exists(ref_xml):                       # the reference's full force-field file parses too
        full = ff.ForceField(ref_xml)
        assert full.cutoffs == field.cutoffs and full.dispersion
        forces = full.createForces(pdb, ff.PME, nonbondedCutoff=0.9)
        d = forces[-1]
        assert type(d).__name__ == "CustomDispersionForce" and d.getCutoffDistance() == 0.9
        assert d.getNonbondedMethod() == d.CutoffPeriodic and d.getNumMolecules() == 14


@pytest.mark.gpu
def test_context_from_pdb(tmp_path):
    from mbpol_openmm_plugin_b200 import forcefield as ff
    g = load_golden("water14_xml")
    pdb_path = str(tmp_path / "w14.pdb")
    write_pdb(pdb_path, g["pos_mol"])
    ctx = ff.ForceField().createContext(ff.PDBFile(pdb_path), ff.CutoffNonPeriodic)
    e = ctx.getPotentialEnergy()
    # PDB coordinates carry 3 decimals in Angstrom: compare with a fresh evaluation of the same rounded coordinates
    pdb = ff.PDBFile(pdb_path)
    assert len(ctx.getPairs()) >= 70 and np.isfinite(e)
    f = ctx.getForces().reshape(14, 4, 3)
    assert np.all(f[:, 3] == 0.0)
    many = ctx.energies["CalcMBPolTwoBodyForce"] + ctx.energies["CalcMBPolThreeBodyForce"]
    # every term against the oracle on the rounded PDB coordinates themselves: lists bit-exact, E 1e-9, F 1e-6
    from oracle.refbind import best, canonical
    o = best()
    widx = np.asarray(pdb.water_indices())
    posO = pdb.positions[widx[:, 0]].copy()
    pairs, trips = canonical(o.pairs(posO, None, False, 0.65)), canonical(o.triplets(posO, None, False, 0.45))
    assert np.array_equal(ctx.getPairs(), pairs) and np.array_equal(ctx.getTriplets(), trips)
    e2, f2 = o.eval_2b(pdb.positions, widx, pairs, None, False)[:2]
    e3, f3 = o.eval_3b(pdb.positions, widx, trips, None, False)[:2]
    e1, f1 = o.eval_1b(pdb.positions, widx, None, False)
    assert abs(ctx.energies["CalcMBPolTwoBodyForce"] - e2) <= 1e-9 * abs(e2)
    assert abs(ctx.energies["CalcMBPolThreeBodyForce"] - e3) <= 1e-9 * abs(e3)
    assert abs(ctx.energies["CalcMBPolOneBodyForce"] - e1) <= 1e-9 * abs(e1)
    assert np.abs(ctx.getForces() - (f1 + f2 + f3)).max() < 1e-6
    assert abs(e - (many + e1)) < 1e-9
