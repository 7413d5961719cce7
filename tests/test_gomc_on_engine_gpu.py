"""GOMC ITSELF on the engine: the unmodified reference move loop, PRNG, CBMC and I/O linked
against integration/gomc_shim.cu + libgomc_b200.so (oracle/_ref/GOMC_B200_NVT) versus the
reference's own CPU executable (oracle/_ref/GOMC_CPU_NVT, +p1) on the same input and seed.

Bar = the reference's own regression bar (test/Run_Examples.py:124-160, byte-identical PDB)
plus north_star's "identical accept/reject over a fixed-seed run": identical acceptance
counters and per-step energies to 1e-9 for 10^4 single-molecule steps.  The MultiParticle run
reports the first step at which the trajectories part (forces enter the trial displacement,
SURVEY.md section 7 hard part 2) and must stay identical for at least the first 200 steps."""
import os

import pytest

from integration import run_parity

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _need(ens="NVT"):
    for exe in (f"GOMC_CPU_{ens}", f"GOMC_B200_{ens}"):
        if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", exe)):
            pytest.skip(f"oracle/_ref/{exe} not built (make -f integration/Makefile)")


def test_nvt_translate_rotate_identical_trajectory():
    _need()
    r = run_parity.compare(mols=343, steps=10000, mp=False)
    print(r)
    assert r["steps_printed"] >= 10000
    assert r["first_divergent_step"] is None, r
    assert r["counters_identical"], r
    assert r["pdb_identical"], r


def test_nvt_multiparticle_trajectory():
    _need()
    r = run_parity.compare(mols=343, steps=1500, mp=True)
    print(r)
    assert r["steps_printed"] >= 1500
    assert r["first_divergent_step"] is None or r["first_divergent_step"] > 200, r


def test_gemc_pentane_cbmc_swaps():
    """BASELINE configs[2] in small: GEMC-NVT of TraPPE-UA n-pentane (with partial charges so
    that the reciprocal deltas are not identically zero), liquid + vapour box, translate /
    rotate / CBMC regrowth / CBMC molecule transfer.  Every accepted swap goes through
    SwapDestRecip + SwapSourceRecip + SwapCorrection x2 + SwapSelf + UpdateRecip on both boxes
    of the engine, with molecules changing box."""
    _need("GEMC")
    r = run_parity.compare(mols=120, steps=3000, ens="GEMC", rcut=10.0)
    print(r)
    assert r["steps_printed"] >= 2 * 3000          # both boxes, every step
    assert r["first_divergent_step"] is None, r
    assert r["counters_identical"], r
    assert r["pdb_identical"], r
