"""EwaldCached (CachedFourier true, BASELINE configs[1]) evidence.

The engine does not replicate the reference's per-molecule cos/sin cache (9.9 GB at configs[1],
DESIGN.md section 1 a13): its host-mirror class EwaldCached maps RestoreMol / exgMolCache /
backupMolCache to no-ops over the same kernels as Ewald.  tests/golden/cached_*.npz are dumps of
the reference's CACHED class (oracle/make_golden.py cached) driven through the same sequence as
the plain fixtures -- MolReciprocal as ref - cached + new with RestoreMol after every rejection,
SwapDestRecip handing the cached rows to SwapSourceRecip, MultiParticle with backupMolCache /
exgMolCache.  Shown here: (1) the reference's two classes return the same bits for every array
of that sequence, so one implementation can stand for both; (2) the oracle and (3) the engine
reproduce the cached dumps."""
import glob
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.helpers import rel_err

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CACHED = sorted(glob.glob(os.path.join(GOLD, "cached_*.npz")))
TOL = 1e-9


def _xyz(d, key):
    return [d[f"{key}.{c}"] for c in "xyz"]


def test_have_cached_goldens():
    assert len(CACHED) >= 2


@pytest.mark.parametrize("path", CACHED, ids=[os.path.basename(p)[:-4] for p in CACHED])
def test_cached_class_returns_the_plain_class_bits(path):
    c = dict(np.load(path))
    p = dict(np.load(os.path.join(GOLD, os.path.basename(path)[len("cached_"):])))
    shared = [k for k in c if k in p]
    # everything the cached class accepts is there: single-molecule deltas, swaps, MP
    for need in ("box0.move.dRecip", "box0.SwapDestRecip", "box0.SwapSourceRecip",
                 "box0.SwapCorrection.new", "box0.BoxReciprocal", "box0.mpDisplace.wRatio"):
        assert need in shared, need
    assert len(shared) >= 200
    for k in shared:
        assert np.array_equal(c[k], p[k]), k


@pytest.mark.parametrize("path", CACHED, ids=[os.path.basename(p)[:-4] for p in CACHED])
def test_oracle_reproduces_cached_dumps(path):
    d = dict(np.load(path))
    po.set_threads(1)
    o = po.Oracle.from_dump(d, 0)
    x, y, z = _xyz(d, "coords")
    kx, ky, kz, hs, pf, _ = o.recip_init_orth()
    ms, bm = d["molStart"], d["box0.mols"]
    q = d["particleCharge"]
    sR, sI = o.box_recip_sums(bm, ms, x, y, z, q, kx, ky, kz)
    assert np.array_equal(sR, d["box0.BoxReciprocalSums.sumRnew"])
    assert np.array_equal(sI, d["box0.BoxReciprocalSums.sumInew"])
    st = d["box0.move.start"]
    for t, m in enumerate(d["box0.move.mol"]):
        nx, ny, nz = (d[f"box0.move.{c}"][st[t]:st[t + 1]] for c in "xyz")
        sl = slice(ms[m], ms[m + 1])
        e, sRn, sIn = o.mol_reciprocal(q[sl], (x[sl], y[sl], z[sl]), (nx, ny, nz), kx, ky, kz, pf,
                                       d["box0.sumRref"], d["box0.sumIref"])
        assert e - d["box0.sysPotRef.recip"][0] == d["box0.move.dRecip"][t]     # bit-exact


@pytest.mark.gpu
@pytest.mark.parametrize("path", CACHED, ids=[os.path.basename(p)[:-4] for p in CACHED])
def test_engine_reproduces_cached_dumps(path):
    """The engine against the CACHED reference class: every trial is followed by the no-op
    that stands for RestoreMol (the state must be untouched: the same trial again returns the
    same bits), one accepted move goes through UpdateRecip and back."""
    from gomc_b200 import engine as eng
    from tests.test_golden_gpu import engine_from_dump
    d = dict(np.load(path))
    e = engine_from_dump(d)
    try:
        ms = d["molStart"]
        x, y, z = _xyz(d, "coords")
        st = d["box0.move.start"]
        ref_recip = d["box0.sysPotRef.recip"][0]
        for t, m in enumerate(d["box0.move.mol"]):
            nx, ny, nz = (d[f"box0.move.{c}"][st[t]:st[t + 1]] for c in "xyz")
            en = e.mol_reciprocal(0, int(m), nx, ny, nz)
            assert abs(en - (d["box0.move.dRecip"][t] + ref_recip)) <= TOL * abs(ref_recip)
            # rejected -> RestoreMol(m) in the reference, nothing here; state unchanged
            assert e.mol_reciprocal(0, int(m), nx, ny, nz) == en
        # accept the first non-overlapping trial, then move the molecule back: the reference
        # sums must be where they started (UpdateRecip twice)
        t = int(np.flatnonzero(d["box0.move.overlap"] == 0)[0])
        m = int(d["box0.move.mol"][t])
        sl = slice(ms[m], ms[m + 1])
        nx, ny, nz = (d[f"box0.move.{c}"][st[t]:st[t + 1]] for c in "xyz")
        e.mol_reciprocal(0, m, nx, ny, nz)
        e.set_molecule_coords(m, nx, ny, nz)
        e.update_recip(0)
        back = e.mol_reciprocal(0, m, x[sl], y[sl], z[sl])
        assert abs(back - ref_recip) <= TOL * abs(ref_recip)
        e.set_molecule_coords(m, x[sl], y[sl], z[sl])
        e.update_recip(0)
        m = int(d["box0.swap.mol"][0])
        nc = _xyz(d, "box0.swap.newCoords")
        sl = slice(ms[m], ms[m + 1])
        dest = e.swap_reciprocal(0, m, *nc, True)
        src = e.swap_reciprocal(0, m, x[sl], y[sl], z[sl], False)
        assert abs(dest - (d["box0.SwapDestRecip"][0] + ref_recip)) <= TOL * abs(ref_recip)
        assert abs(src - (d["box0.SwapSourceRecip"][0] + ref_recip)) <= TOL * abs(ref_recip)
    finally:
        e.close()
