"""Shared helpers for the parity tests: oracle <-> engine plumbing."""
import numpy as np

from gomc_b200 import synth
from oracle import pyoracle as po


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(float(np.max(np.abs(b))) if b.size else 0.0, 1e-300)
    return float(np.max(np.abs(a - b))) / scale if a.size else 0.0


def box_atoms(s):
    return np.arange(s.n_atoms, dtype=np.int32)


def box_mols(s):
    return np.arange(s.n_mols, dtype=np.int32)


def oracle_for(s):
    return po.Oracle.from_system(s)


def random_move(s, rng, m, amp):
    """Rigid displacement of molecule m by up to amp per axis, wrapped."""
    sl = slice(s.mol_start[m], s.mol_start[m + 1])
    d = rng.uniform(-amp, amp, size=3)
    if getattr(s, "cell_basis", None) is not None:   # wrap in unslant coordinates
        r = np.stack([s.x[sl], s.y[sl], s.z[sl]], 1) + d
        u = np.mod(r @ s.cell_basis_inv, s.axis)
        u = np.minimum(u, np.nextafter(s.axis, 0))
        r = u @ s.cell_basis
        return r[:, 0].copy(), r[:, 1].copy(), r[:, 2].copy()
    nx = np.mod(s.x[sl] + d[0], s.axis[0])
    ny = np.mod(s.y[sl] + d[1], s.axis[1])
    nz = np.mod(s.z[sl] + d[2], s.axis[2])
    return nx, ny, nz


SMALL_SYSTEMS = {
    "spce_small": lambda: synth.make_spce(216, r_cut=7.0),            # 3 cells/axis: generic PBC path
    "spce_mid": lambda: synth.make_spce(1000, r_cut=7.5),             # 4 cells/axis: shifted path
    "argon": lambda: synth.make_argon(864, r_cut=8.0),
    "mixture_std": lambda: synth.make_mixture(),
    "mixture_shift": lambda: synth.make_mixture(vdw_kind=synth.VDW_SHIFT),
    "mixture_switch": lambda: synth.make_mixture(vdw_kind=synth.VDW_SWITCH, r_switch=6.5),
    "spce_triclinic": lambda: synth.make_spce(
        343, r_cut=6.0, cell_vectors=synth.triclinic_cell(24.5, (85.0, 70.0, 100.0))),
    "mixture_exp6": lambda: synth.make_mixture(vdw_kind=synth.VDW_EXP6, n_b_exp=16.0,
                                               du_eps=12.0, du_sigma=1.2),
    "mixture_martini": lambda: synth.make_mixture(vdw_kind=synth.VDW_SWITCH, r_switch=6.0,
                                                  martini=True, ewald=False, n_b_exp=12.0),
}
