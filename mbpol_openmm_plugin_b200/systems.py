"""Water-box inputs for the 2B+3B path: PDB reader and the synthetic liquid boxes of BASELINE.json / SURVEY.md 8d.

The synthetic boxes tile the 512-water liquid cell the reference's own benchmark script uses
(python/pdb/H2O_1.pdb, box 2.48343697144598 nm, python/utils/run_benchmark.py:37), stored here as
data/water512_cell.npz (O,H1,H2 coordinates in nm, M sites dropped):  2^3 -> 4,096, 4^3 -> 32,768, 8^3 -> 262,144 waters.
A rigid per-molecule Gaussian displacement (sigma 0.005 nm, seed 20260101 + N) breaks the replica degeneracy.
"""
from __future__ import annotations

import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
CELL512 = os.path.join(PKG, "data", "water512_cell.npz")


def read_pdb_waters(path):
    """O,H1,H2 of every HOH residue (M sites dropped): array [n_mol, 3, 3] in nm (PDB is in Angstrom)."""
    mol = {}
    order = []
    for line in open(path):
        if not line.startswith(("ATOM", "HETATM")):
            continue
        name = line[12:16].strip()
        key = (line[21], line[22:27])
        if key not in mol:
            mol[key] = {}
            order.append(key)
        mol[key][name] = [float(line[30:38]), float(line[38:46]), float(line[46:54])]
    out = np.array([[mol[k]["O"], mol[k]["H1"], mol[k]["H2"]] for k in order], dtype=np.float64)
    return out * 0.1


def particle_indices(n_mol, stride=3):
    """{O,H1,H2} particle indices for molecules laid out with `stride` particles each (4 = O,H1,H2,M as in mbpol.xml)."""
    base = np.arange(n_mol, dtype=np.int32)[:, None] * stride
    return base + np.arange(3, dtype=np.int32)[None, :]


def with_virtual_sites(pos_mol):
    """[n_mol,3,3] -> particle array [4*n_mol,3] with an M site per molecule (weights of python/mbpol.xml:14)."""
    O, H1, H2 = pos_mol[:, 0], pos_mol[:, 1], pos_mol[:, 2]
    w = 0.573293118
    M = w * O + 0.5 * (1 - w) * (H1 + H2)
    return np.concatenate([pos_mol, M[:, None, :]], axis=1).reshape(-1, 3)


def synthetic_box(tiles, sigma_nm=0.005):
    """(pos_mol [n,3,3] nm, box [3] nm) for a tiles^3 replication of the 512-water cell; molecule order = tile-major."""
    d = np.load(CELL512)
    cell, L = d["pos_mol"], float(d["box_nm"])
    shifts = np.array([[i, j, k] for i in range(tiles) for j in range(tiles) for k in range(tiles)], dtype=np.float64) * L
    pos = (cell[None, :, :, :] + shifts[:, None, None, :]).reshape(-1, 3, 3)
    n = pos.shape[0]
    rng = np.random.default_rng(20260101 + n)
    pos = pos + rng.normal(0.0, sigma_nm, size=(n, 1, 3))
    return np.ascontiguousarray(pos), np.array([L * tiles] * 3)


WORKLOADS = {          # name -> tiles
    "synthetic_4096": 2,
    "synthetic_32768": 4,
    "synthetic_262144": 8,
}
