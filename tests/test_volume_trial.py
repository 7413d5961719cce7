"""NPT volume trials against the UNMODIFIED reference's VolumeTransfer move object.

tests/golden/npt_spce343.npz (oracle/make_golden.py npt -> gomc_probe_NPT `volume`) holds one
ACCEPTED and one REJECTED volume trial: the scaled coordinates the move produced, the k list
RecipInit(newDim) built, sumRnew/sumInew of BoxReciprocalSetup, the energies of
VolumeTransfer::CalcEn (src/moves/VolumeTransfer.h:139-198), and -- after the accept branch
(UpdateRecip + UpdateRecipVec, :255-260) and after the reject branch (:263-269) -- a
single-molecule MolReciprocal / MoleculeInter that shows which state the Ewald object is in.

  -m "not gpu": the C oracle reproduces the k lists and sums bit for bit
  -m gpu      : the engine's call sequence of SURVEY.md section 8b rule 6
                recip_init -> box_reciprocal_setup -> box_inter -> box_reciprocal(new)
                -> update_recip + update_recip_vec | reject, through the C ABI
"""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "npt_spce343.npz")
TOL = 1e-9


def _xyz(d, key):
    return [d[f"{key}.{c}"] for c in "xyz"]


@pytest.fixture(scope="module")
def npt():
    if not os.path.exists(GOLD):
        pytest.skip("npt_spce343.npz not generated")
    return dict(np.load(GOLD))


def _oracle(d, axis, volume):
    """The reference's newDim: cbrt-scaled axes and the stored volume (oldVolume + delta)."""
    from oracle import pyoracle as po
    dd = dict(d)
    dd["box0.axis"] = np.asarray(axis, dtype=np.float64)
    o = po.Oracle.from_dump(dd)
    o.p.volume = float(volume)
    return o


def test_volume_trial_oracle_bit_exact(npt):
    d = npt
    mols = d["box0.mols"]
    for t in (0, 1):
        tag = f"trial{t}"
        o = _oracle(d, d[tag + ".newAxis"], d[tag + ".newVolume"][0])
        kx, ky, kz, hs, pf, _ = o.recip_init_orth()
        assert len(kx) == int(d[tag + ".nk"][0])
        for a, n in ((kx, "kx"), (ky, "ky"), (kz, "kz"), (hs, "hsqr"), (pf, "prefact")):
            assert np.array_equal(a, d[f"{tag}.{n}"]), n
        x, y, z = _xyz(d, tag + ".newCoords")
        sR, sI = o.box_recip_sums(mols, d["molStart"], x, y, z, d["particleCharge"], kx, ky, kz)
        assert np.array_equal(sR, d[tag + ".sumRnew"]) and np.array_equal(sI, d[tag + ".sumInew"])
        assert o.box_reciprocal(sR, sI, pf) == d[tag + ".recip"][0]
        ba = np.arange(int(d["nAtoms"][0]), dtype=np.int32)
        lj, re = o.box_inter(x, y, z, d["particleKind"], d["particleMol"], d["particleCharge"], ba)
        assert abs(lj - d[tag + ".inter"][0]) <= 1e-13 * abs(lj)
        assert abs(re - d[tag + ".real"][0]) <= 1e-13 * abs(re)


@pytest.mark.gpu
def test_volume_trial_state_machine_gpu(npt):
    from gomc_b200 import engine as eng
    d = npt
    e = eng.Engine(1)
    try:
        e.init_forcefield(d["ff.sigmaSq"], d["ff.epsilon_cn"], d["ff.n"], int(d["ff.vdwKind"][0]),
                          int(d["ff.kindCount"][0]), float(d["ff.rCut"][0]),
                          d["ff.rCutCoulomb"][:1], float(d["ff.rCutLow"][0]),
                          float(d["ff.rswitch"][0]), d["ff.alpha"][:1], int(d["ff.ewald"][0]),
                          int(d["ff.electrostatic"][0]))
        e.init_topology(d["particleKind"], d["particleMol"], d["particleCharge"], d["molStart"])
        e.set_box_molecules(0, d["box0.mols"])
        e.set_box_axes(0, d["box0.axis"])
        e.set_coords(*_xyz(d, "coords"))
        e.set_com(*_xyz(d, "com"))
        # Ewald::Init with the NPT head-room (RecipCountInit, excess 1.25 -> imageTotal)
        e.init_ewald(int(d["box0.imageTotal"][0]), d["ff.recip_rcut"][:1])
        n0, _ = e.recip_init(0, d["box0.axis"])
        e.box_reciprocal_setup(0)
        e.set_recip_ref(0)
        assert n0 == int(d["box0.nk"][0])

        def probe_move(tag):
            m = int(d[tag + ".mol"][0])
            pos = _xyz(d, tag + ".newPos")
            lj, re, ov = e.molecule_inter(0, m, *pos)
            en = e.mol_reciprocal(0, m, *pos)
            ref = d[tag + ".MolReciprocal"][0] + d[tag + ".sysPotRef.recip"][0]
            assert ov == bool(d[tag + ".overlap"][0])
            assert abs(lj - d[tag + ".dLJ"][0]) <= TOL * max(abs(lj), 1.0)
            assert abs(re - d[tag + ".dReal"][0]) <= TOL * max(abs(re), 1.0)
            assert abs(en - ref) <= TOL * abs(ref)

        probe_move("state0")
        cur_axis, cur_xyz, cur_com = d["box0.axis"], _xyz(d, "coords"), _xyz(d, "com")
        for t in (0, 1):
            tag = f"trial{t}"
            nk = int(d[tag + ".nk"][0])
            # VolumeTransfer::CalcEn on the scaled box (the structure factor algorithm is the
            # engine's default; every algorithm on the accepted trial)
            e.set_box_axes(0, d[tag + ".newAxis"])
            e.set_coords(*_xyz(d, tag + ".newCoords"))
            e.set_com(*_xyz(d, tag + ".newCOM"))
            n, _ = e.recip_init(0, d[tag + ".newAxis"], d[tag + ".newVolume"][0])
            assert n == nk
            for a, name in zip(e.get_kvectors(0, eng.K_NEW | eng.K_DEVICE, nk),
                               ("kx", "ky", "kz", "hsqr", "prefact")):
                assert np.array_equal(a, d[f"{tag}.{name}"]), name
            scale = max(np.max(np.abs(d[tag + ".sumRnew"])), np.max(np.abs(d[tag + ".sumInew"])))
            for algo in ((0, 1, 2, 3, 5, 4) if t == 0 else (4,)):
                e.set_recip_algo(algo)
                e.mark_coords_changed()
                en = e.box_reciprocal_setup(0)
                gR, gI = e.get_recip_sums(0, eng.SUM_NEW, nk)
                assert np.max(np.abs(gR - d[tag + ".sumRnew"])) <= TOL * scale, algo
                assert np.max(np.abs(gI - d[tag + ".sumInew"])) <= TOL * scale, algo
                assert abs(en - d[tag + ".recip"][0]) <= TOL * abs(en), algo
            lj, re = e.box_inter(0)
            assert abs(lj - d[tag + ".inter"][0]) <= TOL * abs(lj)
            assert abs(re - d[tag + ".real"][0]) <= TOL * abs(re)
            en = e.box_reciprocal(0, True)
            assert abs(en - d[tag + ".recip"][0]) <= TOL * abs(en)
            # the reference k set and sums are untouched by the trial (rule 1)
            rR, rI = e.get_recip_sums(0, eng.SUM_REF, n0)
            if t == 0:
                assert np.max(np.abs(rR - d["box0.sumRref"])) <= TOL * scale
                assert np.array_equal(e.get_kvectors(0, eng.K_REF, n0)[0], d["box0.kx"])
                # accept: VolumeTransfer::Accept, :243-260
                e.update_recip(0)
                e.update_recip_vec(0)
                cur_axis, cur_xyz, cur_com = (d[tag + ".newAxis"], _xyz(d, tag + ".newCoords"),
                                              _xyz(d, tag + ".newCOM"))
                n0 = nk
                en = e.box_reciprocal(0, False)
                assert abs(en - d[tag + ".after.BoxReciprocal"][0]) <= TOL * abs(en)
                assert np.array_equal(e.get_kvectors(0, eng.K_REF | eng.K_DEVICE, nk)[4],
                                      d[tag + ".prefact"])
            else:
                # reject: the box, coordinates and COMs go back; nothing is called on the Ewald
                # object (non-cached), and the next delta must see the accepted state of trial 0
                e.set_box_axes(0, cur_axis)
                e.set_coords(*cur_xyz)
                e.set_com(*cur_com)
                assert np.array_equal(e.get_kvectors(0, eng.K_REF, n0)[0], d["trial0.kx"])
            probe_move(tag + ".after")
        # a full recomputation on the final state reproduces the accepted trial's energy
        en = e.box_reciprocal_sums(0)
        assert abs(en - d["trial0.recip"][0]) <= TOL * abs(en)
    finally:
        e.close()
