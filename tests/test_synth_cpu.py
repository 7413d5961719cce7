"""CPU tests of the host logic: the synthetic-system generator derives the
force-field tables and Ewald constants exactly as the reference does (pinned
by the golden dumps, which hold what the reference itself derived from the
files the generator wrote)."""
import glob
import os

import numpy as np
import pytest

from gomc_b200 import synth
from oracle.make_golden import CASES

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", sorted(CASES))
def test_generator_matches_what_reference_parsed(name):
    d = dict(np.load(os.path.join(GOLD, name + ".npz")))
    s = CASES[name][0]()
    assert s.n_atoms == d["nAtoms"][0] and s.n_mols == d["nMols"][0]
    assert np.array_equal(s.mol_start, d["molStart"])
    assert np.array_equal(s.mol, d["particleMol"])
    assert np.array_equal(s.charge, d["particleCharge"])
    assert np.array_equal(s.kind, d["particleKind"])
    assert np.allclose(s.axis, d["box0.axis"], rtol=1e-15)
    sig, eps, nn = s.ff.tables()
    if s.ff.is_martini:   # CHARMM-unit round trip of the parameter file: 1e-14, not bit-exact
        assert np.allclose(sig, d["ff.sigmaSq"], rtol=1e-13, atol=0)
        assert np.allclose(eps, d["ff.epsilon_cn"], rtol=1e-13, atol=0)
    else:
        assert np.array_equal(sig, d["ff.sigmaSq"])          # FFParticle::Blend
        assert np.array_equal(eps, d["ff.epsilon_cn"])
    assert np.array_equal(nn, d["ff.n"])
    if s.ff.vdw_kind == synth.VDW_EXP6:   # Brent roots: the reference works in float
        r_min, exp_c, r_max_sq = s.ff.exp6_tables()
        assert np.allclose(r_min, d["ff.rMin"], rtol=1e-6)
        assert np.allclose(exp_c, d["ff.expConst"], rtol=1e-13)
        assert np.allclose(r_max_sq, d["ff.rMaxSq"], rtol=1e-5)
    assert s.ff.alpha == d["ff.alpha"][0]                 # Forcefield.cpp:80
    assert s.ff.recip_rcut == d["ff.recip_rcut"][0]       # Forcefield.cpp:82
    # the reference re-wraps whole molecules on load; atoms stay congruent mod L
    if s.cell_basis is not None:
        assert np.allclose(s.cell_basis.reshape(-1), d["box0.cellBasis"], rtol=0, atol=1e-15)
        assert np.allclose(s.cell_basis_inv.reshape(-1), d["box0.cellBasisInv"], rtol=0, atol=1e-14)
        return
    for c, arr in zip("xyz", (s.x, s.y, s.z)):
        diff = np.abs(d[f"coords.{c}"] - arr)
        L = s.axis[0]
        assert np.all((diff < 1e-9) | (np.abs(diff - L) < 1e-9))


def test_sizes_of_baseline_configs():
    """Problem sizes of BASELINE.md section 3 (computed, not allocated)."""
    import math
    for n_mol, kmax_expected in ((10000, 25), (33334, 37)):
        L = round((n_mol / 0.0334) ** (1 / 3), 3)
        recip_rcut = -2.0 * math.log(1e-5) / 10.0
        assert int(recip_rcut * L / (2 * math.pi)) + 1 == kmax_expected


def test_pentane_generator_and_two_box_writer(tmp_path):
    """TraPPE-UA n-pentane (BASELINE configs[2]) and the GEMC two-box input files the
    GOMC-on-engine runner feeds to the reference executables."""
    import numpy as np
    s = synth.make_pentane(64, L=30.0, charged=True)
    assert s.n_atoms == 5 * 64 and abs(float(np.sum(s.charge))) < 1e-12
    m = 7
    sl = slice(s.mol_start[m], s.mol_start[m + 1])
    r = np.stack([s.x[sl], s.y[sl], s.z[sl]], 1)
    d = r[1:] - r[:-1]
    d -= s.axis * np.round(d / s.axis)
    assert np.allclose(np.linalg.norm(d, axis=1), 1.54, atol=2e-3)      # rounded to 1e-3 A
    cosang = [-(d[i] @ d[i + 1]) / 1.54 ** 2 for i in range(3)]
    assert np.allclose(np.degrees(np.arccos(cosang)), 114.0, atol=0.3)
    v = synth.make_pentane(8, L=40.0, seed=77, charged=True)
    synth.write_gomc_inputs(s, str(tmp_path), multiparticle=False, run_steps=100, second=v)
    conf = (tmp_path / "in.conf").read_text()
    for needle in ("GEMC NVT", "Coordinates 1 box1.pdb", "Structure 1 box1.psf", "SwapFreq",
                   "RegrowthFreq", "CellBasisVector1 1 40.0", "RcutCoulomb 1"):
        assert needle in conf, needle
    psf = (tmp_path / "box1.psf").read_text()
    assert "%8d !NATOM" % 40 in psf and "%8d !NPHI" % 16 in psf
    par = (tmp_path / "par.inp").read_text()
    assert "CH3\tCH2\tCH2\tCH2" in par
    # single-box output is unchanged by the two-box support
    synth.write_gomc_inputs(s, str(tmp_path / "one"))
    assert "GEMC" not in (tmp_path / "one" / "in.conf").read_text()
