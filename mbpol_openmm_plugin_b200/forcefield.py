"""ForceField front-end for the 2B+3B path without OpenMM: PDB + mbpol.xml-style parameters -> forces -> context.

Mirrors what `app.ForceField("mbpol.xml").createSystem(pdb.topology, nonbondedMethod=...)` does for these two forces
(python/mbpol.py:81-225): the cutoff is the element's `cutoff_nm` attribute (NOT nonbondedCutoff, mbpol.py:99,131), the
nonbonded method maps NoCutoff / CutoffNonPeriodic / CutoffPeriodic / PME onto the plugin enum (mbpol.py:118-121), and one
`addParticle([O, H1, H2])` is issued per HOH residue (mbpol.py:138-150).  Both the reduced XML shipped here
(data/mbpol_2b3b.xml) and the reference's full python/mbpol.xml parse: only these force elements are read.
An `<MBPolOneBodyForce>` element adds the one-body term (python/mbpol.py:12-79: NoCutoff / CutoffNonPeriodic ->
NonPeriodic, CutoffPeriodic / PME -> Periodic, one addOneBody([O, H1, H2]) per water).  A `<Script>` that builds the
TT-damped dispersion CustomNonbondedForce (python/mbpol.xml:37-108) adds a CustomDispersionForce with
`nonbondedCutoff` as its cutoff distance and the method map of mbpol.xml:88-90 (PME -> CutoffPeriodic)."""
from __future__ import annotations

import os
import xml.etree.ElementTree as ET

import numpy as np

from .api import CustomDispersionForce, MBPolContext, MBPolOneBodyForce, MBPolThreeBodyForce, MBPolTwoBodyForce

DEFAULT_XML = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "mbpol_2b3b.xml")

NoCutoff, CutoffNonPeriodic, CutoffPeriodic, PME = "NoCutoff", "CutoffNonPeriodic", "CutoffPeriodic", "PME"
_METHOD = {NoCutoff: 0, CutoffNonPeriodic: 2, CutoffPeriodic: 1, PME: 1}


class PDBFile:
    """Minimal PDB reader: atoms (name, residue name, residue id) and positions in nm; CRYST1 box if present."""

    def __init__(self, path):
        self.atoms, pos, self.box_nm = [], [], None
        for line in open(path):
            if line.startswith("CRYST1"):
                self.box_nm = [float(line[6:15]) * 0.1, float(line[15:24]) * 0.1, float(line[24:33]) * 0.1]
            elif line.startswith(("ATOM", "HETATM")):
                self.atoms.append((line[12:16].strip(), line[17:20].strip(), (line[21], line[22:27])))
                pos.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
        self.positions = np.asarray(pos, dtype=np.float64) * 0.1

    def water_indices(self):
        """[O, H1, H2] particle indices per HOH residue, in file order"""
        res, order = {}, []
        for i, (name, resname, key) in enumerate(self.atoms):
            if resname not in ("HOH", "WAT"):
                continue
            if key not in res:
                res[key] = {}
                order.append(key)
            res[key][name] = i
        return [[res[k]["O"], res[k]["H1"], res[k]["H2"]] for k in order]


class ForceField:
    def __init__(self, xml_path=DEFAULT_XML):
        root = ET.parse(xml_path).getroot()
        self.cutoffs = {}
        for tag in ("MBPolTwoBodyForce", "MBPolThreeBodyForce"):
            el = root.find(tag)
            if el is not None:
                self.cutoffs[tag] = float(el.attrib["cutoff_nm"])
        self.one_body = root.find("MBPolOneBodyForce") is not None
        self.dispersion = any("tt6" in (el.text or "") and "CustomNonbondedForce" in (el.text or "")
                              for el in root.findall("Script"))
        if not self.cutoffs and not self.one_body:
            raise ValueError("no MBPolOneBodyForce / MBPolTwoBodyForce / MBPolThreeBodyForce element in " + xml_path)

    def createForces(self, pdb, nonbondedMethod=CutoffNonPeriodic, nonbondedCutoff=1.0):
        """nonbondedCutoff (nm; OpenMM's createSystem default is 1 nm) only reaches the dispersion force: the MB-pol forces
        take their cutoffs from the XML (python/mbpol.py:99,131)"""
        if nonbondedMethod not in _METHOD:
            raise ValueError("Illegal nonbonded method for MBPolTwoBodyForce")
        forces = []
        if self.one_body:
            f = MBPolOneBodyForce()
            f.setNonbondedMethod(f.Periodic if _METHOD[nonbondedMethod] == 1 else f.NonPeriodic)
            for ohh in pdb.water_indices():
                f.addOneBody(ohh)
            forces.append(f)
        for tag, cls in (("MBPolTwoBodyForce", MBPolTwoBodyForce), ("MBPolThreeBodyForce", MBPolThreeBodyForce)):
            if tag not in self.cutoffs:
                continue
            f = cls()
            f.setCutoff(self.cutoffs[tag])
            f.setNonbondedMethod(_METHOD[nonbondedMethod])
            for ohh in pdb.water_indices():
                f.addParticle(ohh)
            forces.append(f)
        if self.dispersion:
            d = CustomDispersionForce()
            d.setNonbondedMethod({NoCutoff: d.NoCutoff, CutoffNonPeriodic: d.CutoffNonPeriodic, CutoffPeriodic: d.CutoffPeriodic,
                                  PME: d.CutoffPeriodic}[nonbondedMethod])
            d.setCutoffDistance(nonbondedCutoff)
            for ohh in pdb.water_indices():
                d.addMolecule(ohh)
            forces.append(d)
        return forces

    def createContext(self, pdb, nonbondedMethod=CutoffNonPeriodic, box_nm=None, nonbondedCutoff=1.0, **kw):
        box = box_nm if box_nm is not None else pdb.box_nm
        periodic = _METHOD[nonbondedMethod] == 1
        ctx = MBPolContext(len(pdb.atoms), self.createForces(pdb, nonbondedMethod, nonbondedCutoff),
                           box_nm=box if periodic else None, **kw)
        ctx.setPositions(pdb.positions)
        return ctx
