"""Parity at the BASELINE config sizes against the UNMODIFIED reference.

tests/golden/full_*.npz (oracle/make_golden_full.py) hold, for the synthetic boxes of
BASELINE.json at their full sizes, the reference's own BoxInter energies and its
BoxReciprocalSums values on ~256 k-vectors spread over the whole k list.  The boxes are
rebuilt here by the same deterministic generator (coordinate checksum asserted).

  -m gpu      : the CUDA path through the C ABI -- pair sweep and EVERY structure-factor
                algorithm (FP64 MMA, INT8 tensor cores incl. its five-slice form that only
                large boxes reach, non-uniform FFT) -- held to the reference at 1e-9
  -m "not gpu": the C oracle against the same dumps (bit-exact sums, BoxInter to 1e-12)
"""
import os

import numpy as np
import pytest

from gomc_b200 import synth
from tests.helpers import box_atoms, box_mols

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-9

CASES = {
    "full_cfg1_argon4000": lambda: synth.make_argon(4000),
    "full_cfg2_spce10k": lambda: synth.make_spce(10000),
    "full_cfg4_spce100k": lambda: synth.make_spce(33334),
    "full_cfg5_electrolyte1m": lambda: synth.make_electrolyte(),
}


def _checksum(s):
    i = np.arange(s.n_atoms)
    return np.array([s.x.sum(), s.y.sum(), s.z.sum(), (s.x * (i % 97 + 1)).sum(),
                     (s.y * (i % 89 + 1)).sum(), (s.z * (i % 83 + 1)).sum()])


def _load(name):
    path = os.path.join(GOLD, name + ".npz")
    if not os.path.exists(path):
        pytest.skip(f"{name}.npz not generated")
    d = dict(np.load(path))
    s = CASES[name]()
    assert s.n_atoms == int(d["nAtoms"][0])
    assert np.array_equal(_checksum(s), d["coords.checksum"]), "generator drifted from the fixture"
    assert np.array_equal(np.asarray(s.axis, dtype=float), d["box0.axis"])
    return d, s


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_full_size_gpu_vs_reference(name):
    from gomc_b200 import engine as eng
    d, s = _load(name)
    e = eng.Engine.from_system(s)
    try:
        lj, re = e.box_inter(0)
        assert abs(lj - d["BoxInter.inter"][0]) <= TOL * abs(d["BoxInter.inter"][0])
        if s.ff.electrostatic:
            assert abs(re - d["BoxInter.real"][0]) <= TOL * abs(d["BoxInter.real"][0])
        if "slab.index" not in d:
            return
        idx = d["slab.index"]
        nk = int(d["box0.nk"][0])
        assert e.nk == nk
        kx, ky, kz, hs, pf = e.get_kvectors(0, eng.K_REF | eng.K_DEVICE, nk)
        for a, b in ((kx, "kx"), (ky, "ky"), (kz, "kz"), (pf, "prefact")):
            assert np.array_equal(a[idx], d["slab." + b]), b
        refR, refI = d["slab.sumRnew"], d["slab.sumInew"]
        scale = max(np.max(np.abs(refR)), np.max(np.abs(refI)))
        # FP64 MMA, INT8 tensor cores, non-uniform FFT (the default)
        for algo in (2, 3, 5):
            e.set_recip_algo(algo)
            e.mark_coords_changed()
            e.box_reciprocal_sums(0)
            gR, gI = e.get_recip_sums(0, eng.SUM_NEW, nk)
            assert np.max(np.abs(gR[idx] - refR)) <= TOL * scale, f"algo {algo} Re"
            assert np.max(np.abs(gI[idx] - refI)) <= TOL * scale, f"algo {algo} Im"
        e.set_recip_algo(4)
    finally:
        e.close()


@pytest.mark.parametrize("name", ["full_cfg1_argon4000", "full_cfg2_spce10k", "full_cfg4_spce100k"])
def test_full_size_oracle_vs_reference(name):
    from oracle import pyoracle as po
    d, s = _load(name)
    o = po.Oracle.from_system(s)
    if s.n_atoms <= 30000:      # the oracle's serial sweep: seconds at these sizes
        lj, re = o.box_inter(s.x, s.y, s.z, s.kind, s.mol, s.charge, box_atoms(s))
        # the reference ran its OpenMP reduction over all host threads: summation order differs
        assert abs(lj - d["BoxInter.inter"][0]) <= 1e-12 * abs(d["BoxInter.inter"][0])
        if s.ff.electrostatic:
            assert abs(re - d["BoxInter.real"][0]) <= 1e-12 * abs(d["BoxInter.real"][0])
    if "slab.index" in d:
        kx, ky, kz, hs, pf, _ = o.recip_init_orth()
        idx = d["slab.index"]
        assert len(kx) == int(d["box0.nk"][0])
        for a, b in ((kx, "kx"), (ky, "ky"), (kz, "kz"), (pf, "prefact")):
            assert np.array_equal(a[idx], d["slab." + b]), b
        sR, sI = o.box_recip_sums(box_mols(s), s.mol_start, s.x, s.y, s.z, s.charge,
                                  kx[idx].copy(), ky[idx].copy(), kz[idx].copy())
        # same molecule-outer loop order per k: bit-exact
        assert np.array_equal(sR, d["slab.sumRnew"])
        assert np.array_equal(sI, d["slab.sumInew"])
