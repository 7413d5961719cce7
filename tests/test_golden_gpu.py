"""GPU parity against the reference ITSELF: the CUDA engine (through the C ABI)
versus tests/golden/*.npz, i.e. outputs of the unmodified GOMC CPU build on the
same inputs.  Tolerance 1e-9 relative (north_star); flags exact."""
import glob
import os

import numpy as np
import pytest

from gomc_b200 import engine as eng
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-9
# (full_*: full-size slab pins, tests/test_full_size.py; npt_* / pentane_*: own test modules)
GOLDEN = sorted(g for g in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(g).startswith(("full_", "npt_", "pentane_", "cached_")))


def _xyz(d, key):
    return [d[f"{key}.{c}"] for c in "xyz"]


def engine_from_dump(d):
    """An engine initialised the way GOMC initialises its GPU state, from a probe dump."""
    e = eng.Engine(1)
    e.init_forcefield(d["ff.sigmaSq"], d["ff.epsilon_cn"], d["ff.n"], int(d["ff.vdwKind"][0]),
                      int(d["ff.kindCount"][0]), float(d["ff.rCut"][0]), d["ff.rCutCoulomb"][:1],
                      float(d["ff.rCutLow"][0]), float(d["ff.rswitch"][0]), d["ff.alpha"][:1],
                      int(d["ff.ewald"][0]), int(d["ff.electrostatic"][0]),
                      is_martini=int(d["ff.isMartini"][0]),
                      dielectric=float(d["ff.dielectric"][0]) if "ff.dielectric" in d else 1.0)
    if "ff.rMin" in d:
        e.init_exp6(d["ff.rMin"], d["ff.expConst"], d["ff.rMaxSq"])
    e.init_topology(d["particleKind"], d["particleMol"], d["particleCharge"], d["molStart"])
    e.set_box_molecules(0, d["box0.mols"])
    if int(d["box0.orthogonal"][0]):
        e.set_box_axes(0, d["box0.axis"])
    else:
        e.set_box_cell_basis(0, d["box0.cellBasis"], d["box0.cellBasisInv"], d["box0.axis"])
    e.set_coords(*_xyz(d, "coords"))
    e.set_com(*_xyz(d, "com"))
    if "lambda.params" in d:        # fractional molecule (free energy / NeMTMC state)
        lp = d["lambda.params"]
        e.init_softcore(lp[3], lp[4], int(lp[5]), int(lp[6]))
        e.update_lambda(0, int(lp[0]), int(lp[7]), lp[1], lp[2])
    e.nk = 0
    if d["ff.ewald"][0]:
        e.nk = e.setup_ewald(d["box0.axis"], d["ff.recip_rcut"][:1])
    return e


@pytest.fixture(scope="module", params=GOLDEN, ids=[os.path.basename(g)[:-4] for g in GOLDEN])
def gold(request):
    d = dict(np.load(request.param))
    e = engine_from_dump(d)
    yield d, e
    e.close()


def test_pair_energies_and_forces(gold):
    d, e = gold
    lj, re = e.box_inter(0)
    assert abs(lj - d["box0.BoxInter.inter"][0]) <= TOL * abs(lj)
    assert abs(re - d["box0.BoxInter.real"][0]) <= TOL * max(abs(re), 1e-300)
    lj, re = e.box_force(0)
    assert abs(lj - d["box0.BoxForce.inter"][0]) <= TOL * abs(lj)
    gF, gM = e.get_forces(eng.ATOM_FORCE), e.get_forces(eng.MOL_FORCE)
    for i, c in enumerate("xyz"):
        assert rel_err(gF[i], d[f"box0.BoxForce.atomForce.{c}"]) <= TOL
        assert rel_err(gM[i], d[f"box0.BoxForce.molForce.{c}"]) <= TOL


def test_moves_and_trials(gold):
    d, e = gold
    st = d["box0.move.start"]
    for t, m in enumerate(d["box0.move.mol"]):
        nx, ny, nz = (d[f"box0.move.{c}"][st[t]:st[t + 1]] for c in "xyz")
        lj, re, ov = e.molecule_inter(0, int(m), nx, ny, nz)
        assert ov == bool(d["box0.move.overlap"][t])
        assert abs(lj - d["box0.move.dLJ"][t]) <= TOL * max(abs(lj), 1.0)
        assert abs(re - d["box0.move.dReal"][t]) <= TOL * max(abs(re), 1.0)
        if d["ff.ewald"][0]:
            en = e.mol_reciprocal(0, int(m), nx, ny, nz)
            ref = d["box0.move.dRecip"][t] + d["box0.sysPotRef.recip"][0]
            assert abs(en - ref) <= TOL * abs(ref)
    m = int(d["box0.swap.mol"][0])
    en, re, ov = e.particle_inter(0, m, 0, *_xyz(d, "box0.ParticleInter.trialPos"))
    assert np.array_equal(ov.astype(np.int32), d["box0.ParticleInter.overlap"])
    assert rel_err(en, d["box0.ParticleInter.en"]) <= TOL


def test_reciprocal(gold):
    d, e = gold
    if not d["ff.ewald"][0]:
        pytest.skip("no Ewald")
    nk = int(d["box0.nk"][0])
    assert e.nk == nk
    for a, name in zip(e.get_kvectors(0, eng.K_REF, nk), ("kx", "ky", "kz", "hsqr", "prefact")):
        assert np.array_equal(a, d["box0." + name]), name
    for a, name in zip(e.get_kvectors(0, eng.K_REF | eng.K_DEVICE, nk),
                       ("kx", "ky", "kz", "hsqr", "prefact")):
        assert np.array_equal(a, d["box0." + name]), "device " + name
    # per-term, SIMT factorised, FP64 MMA, INT8 tensor cores, non-uniform FFT
    for algo in (0, 1, 2, 3, 5):
        e.set_recip_algo(algo)
        en = e.box_reciprocal_sums(0)
        gR, gI = e.get_recip_sums(0, eng.SUM_NEW, nk)
        scale = max(np.max(np.abs(d["box0.sumRref"])), np.max(np.abs(d["box0.sumIref"])))
        assert np.max(np.abs(gR - d["box0.sumRref"])) <= TOL * scale
        assert np.max(np.abs(gI - d["box0.sumIref"])) <= TOL * scale
        assert abs(en - d["box0.BoxReciprocal"][0]) <= TOL * abs(en)
    e.set_recip_algo(4)
    sf, co = e.box_self_correction(0)
    assert abs(sf - d["box0.BoxSelf"][0]) <= TOL * abs(sf)
    assert abs(co - d["box0.MolCorrection.sum"][0]) <= TOL * abs(co)
    m = int(d["box0.swap.mol"][0])
    ref = d["box0.sysPotRef.recip"][0]
    en = e.swap_reciprocal(0, m, *_xyz(d, "box0.swap.newCoords"), 1)
    assert abs(en - (d["box0.SwapDestRecip"][0] + ref)) <= TOL * abs(en)
    ms = d["molStart"]
    old = [a[ms[m]:ms[m + 1]] for a in _xyz(d, "coords")]
    en = e.swap_reciprocal(0, m, *old, 0)
    assert abs(en - (d["box0.SwapSourceRecip"][0] + ref)) <= TOL * abs(en)
    # swap corrections against the reference's numbers
    co, se = e.swap_correction(0, m, *_xyz(d, "box0.swap.newCoords"))
    assert abs(co - d["box0.SwapCorrection.new"][0]) <= TOL * abs(co)
    assert abs(se - d["box0.SwapSelf"][0]) <= TOL * abs(se)
    co, _ = e.swap_correction(0, m, *old)
    assert abs(co - d["box0.SwapCorrection.old"][0]) <= TOL * abs(co)


def test_virial(gold):
    d, e = gold
    if "box0.Virial.interTens" not in d:
        pytest.skip("fixture predates the virial dump")
    vT, rT = e.box_inter_virial(0)
    assert rel_err(vT, d["box0.Virial.interTens"]) <= TOL
    if d["ff.electrostatic"][0]:
        assert rel_err(rT, d["box0.Virial.realTens"]) <= TOL
    if d["ff.ewald"][0]:
        e.box_reciprocal_sums(0)
        e.set_recip_ref(0)
        for algo in (0, 2, 5):
            e.set_recip_algo(algo)
            wT = e.virial_reciprocal(0)
            assert rel_err(wT, d["box0.Virial.recipTens"]) <= TOL
        e.set_recip_algo(4)


@pytest.mark.parametrize("kind", ["mpDisplace", "mpRotate"])
def test_multiparticle_move(gold, kind):
    """Device MultiParticle step against the reference move object: same variates
    (Philox4x64-10), trial coordinates, energies on them, acceptance weight, reject."""
    d, e = gold
    pre = f"box0.{kind}."
    if pre + "params" not in d:
        pytest.skip("no such move in this fixture")
    if not int(d["box0.orthogonal"][0]):
        L = None          # positions compared directly (no wrap ambiguity expected)
    tmax, rmax, lbeta, step, seed, key = d[pre + "params"]
    rot = kind == "mpRotate"
    vmax = rmax if rot else tmax
    bm = d["box0.mols"]
    ewald = bool(d["ff.ewald"][0])
    # reference forces / torques of the current positions (MultiParticle::Prep)
    if ewald:
        e.box_reciprocal_sums(0)
        e.set_recip_ref(0)
        e.copy_recip(0)
        e.box_force_reciprocal(0)
    lj0, re0 = e.box_force(0)
    e.calculate_torque(0)
    e.mp_transform(0, int(rot), vmax, lbeta, int(step), int(key), int(seed))
    k, inr = e.mp_get_trial(e.n_mols)
    assert np.array_equal(inr[bm], d[pre + "inForceRange"][bm])
    for c, a in zip("xyz", k):
        assert rel_err(a[bm], d[pre + "k." + c][bm]) <= TOL
    e.mp_select(1)
    ax = d["box0.axis"]
    for c, a, cm, L in zip("xyz", e.get_coords(), e.get_com(), ax):
        dx = a - d[pre + "newMolsPos." + c]
        dc = cm - d[pre + "newCOMs." + c]
        if int(d["box0.orthogonal"][0]):       # a point on the box face may wrap either way
            dx -= L * np.round(dx / L)
            dc -= L * np.round(dc / L)
        assert np.max(np.abs(dx)) <= TOL * L and np.max(np.abs(dc)) <= TOL * L
    # MultiParticle::CalcEn on the trial set
    if ewald:
        rc = e.box_reciprocal_sums(0)
    lj, re = e.box_force(0)
    if ewald:
        e.box_force_reciprocal(0)
    e.calculate_torque(0)
    want = d[pre + "newEnergy"]
    assert abs(lj - want[0]) <= TOL * abs(want[0])
    assert abs(re - want[1]) <= TOL * max(abs(want[1]), 1e-300)
    if ewald:
        assert abs(rc - want[2]) <= TOL * abs(want[2])
    w = e.mp_coeff(0, int(rot), vmax, lbeta)
    assert abs(w - d[pre + "wRatio"][0]) <= 1e-8 * abs(d[pre + "wRatio"][0])
    # reject: reference set back, energies as before
    e.mp_select(0)
    lj1, re1 = e.box_force(0)
    assert (lj1, re1) == (lj0, re0)
    for c, a in zip("xyz", e.get_coords()):
        assert np.array_equal(a, d["coords." + c])


@pytest.mark.parametrize("kind", ["bmDisplace", "bmRotate"])
def test_brownian_multiparticle_move(gold, kind):
    d, e = gold
    pre = f"box0.{kind}."
    if pre + "params" not in d:
        pytest.skip("no such move in this fixture")
    tmax, rmax, beta, step, seed, key = d[pre + "params"]
    rot = kind == "bmRotate"
    vmax = rmax if rot else tmax
    bm = d["box0.mols"]
    ewald = bool(d["ff.ewald"][0])
    if ewald:
        e.box_reciprocal_sums(0)
        e.set_recip_ref(0)
        e.copy_recip(0)
        e.box_force_reciprocal(0)
    e.box_force(0)
    e.calculate_torque(0)
    e.bm_transform(0, int(rot), vmax, beta, int(step), int(key), int(seed))
    k, _ = e.mp_get_trial(e.n_mols)
    for c, a in zip("xyz", k):
        assert rel_err(a[bm], d[pre + "k." + c][bm]) <= TOL
    e.mp_select(1)
    try:
        _check_bm_trial(d, e, pre, rot, vmax, beta, bm, ewald)
    finally:
        e.mp_select(0)                        # never leave the module fixture on the trial set
    for c, a in zip("xyz", e.get_coords()):
        assert np.array_equal(a, d["coords." + c])


def _check_bm_trial(d, e, pre, rot, vmax, beta, bm, ewald):
    orth = bool(int(d["box0.orthogonal"][0]))
    for c, a, cm, L in zip("xyz", e.get_coords(), e.get_com(), d["box0.axis"]):
        dx = a - d[pre + "newMolsPos." + c]
        dc = cm - d[pre + "newCOMs." + c]
        if orth:
            dx -= L * np.round(dx / L)
            dc -= L * np.round(dc / L)
        # displacements scale with force * BETA * max: compare relative to their size
        scale = max(L, float(np.max(np.abs(d[pre + "k." + c][bm]))))
        assert np.max(np.abs(dx)) <= TOL * scale and np.max(np.abs(dc)) <= TOL * scale
    if pre + "wRatio" in d:
        if ewald:
            e.box_reciprocal_sums(0)
        e.box_force(0)
        if ewald:
            e.box_force_reciprocal(0)
        e.calculate_torque(0)
        w, want = e.bm_coeff(0, int(rot), vmax, beta), d[pre + "wRatio"][0]
        kmax = max(float(np.max(np.abs(d[pre + "k." + c][bm]))) for c in "xyz")
        if not np.isfinite(want):
            assert w == want                  # EXP6 overlap: -inf on both sides
        elif kmax < 1e3:
            assert abs(w - want) <= 1e-8 * max(abs(want), 1.0)
        # else: the start configuration's forces are astronomically large (overlapping
        # Martini beads) and a rotation by ~1e11 rad is ill-conditioned in any arithmetic


def test_exchange_and_lambda_reciprocal(gold):
    """MolExchangeReciprocal (two chained calls), ChangeLambdaRecip, ChangeRecip."""
    from tests.test_oracle_golden import exchange_weights
    d, e = gold
    if not d["ff.ewald"][0] or "box0.exchange.mols" not in d:
        pytest.skip("no Ewald")
    nk = int(d["box0.nk"][0])
    e.box_reciprocal_sums(0)
    e.set_recip_ref(0)
    ref = d["box0.sysPotRef.recip"][0]
    ms, q = d["molStart"], d["particleCharge"]
    xyz = _xyz(d, "coords")
    lp = d.get("lambda.params")
    coef = (lambda m: float(np.sqrt(lp[2])) if lp is not None and int(m) == int(lp[0]) else 1.0)
    calls = exchange_weights(q, ms, d["box0.exchange.mols"],
                             [_xyz(d, "box0.exchange0.newCoords"),
                              _xyz(d, "box0.exchange1.newCoords")], xyz, coef)
    for c, (w, cx) in enumerate(calls):
        en = e.mol_exchange_reciprocal(0, w, *cx, first_call=(c == 0))
        want = d["box0.exchange.dRecip"][c] + ref
        assert abs(en - want) <= TOL * abs(want)
    gR, gI = e.get_recip_sums(0, eng.SUM_NEW, nk)
    scale = max(np.max(np.abs(d["box0.sumRref"])), np.max(np.abs(d["box0.sumIref"])))
    assert np.max(np.abs(gR - d["box0.exchange.sumRnew"])) <= TOL * scale
    assert np.max(np.abs(gI - d["box0.exchange.sumInew"])) <= TOL * scale
    m = int(d["box0.changeLambda.mol"][0])
    mc = [a[ms[m]:ms[m + 1]] for a in xyz]
    en = e.change_lambda_mol_reciprocal(0, m, *mc, np.sqrt(0.85) - np.sqrt(0.3))
    want = d["box0.changeLambda.dRecip"][0] + ref
    assert abs(en - want) <= TOL * abs(want)
    gR, _ = e.get_recip_sums(0, eng.SUM_NEW, nk)
    assert np.max(np.abs(gR - d["box0.changeLambda.sumRnew"])) <= TOL * scale
    er = e.change_recip(0, m, d["box0.changeRecip.lambda"], 2)
    want = d["box0.changeRecip.dRecip"] + ref
    assert np.max(np.abs(er - want)) <= TOL * np.max(np.abs(want))
    if "box0.changeSelf.dSelf" in d:        # ChangeSelf / ChangeCorrection
        lam = d["box0.changeRecip.lambda"]
        es, ec = e.change_self_correction(0, m)
        assert rel_err((lam - lam[2]) * es, d["box0.changeSelf.dSelf"]) <= TOL
        assert rel_err((lam - lam[2]) * ec, d["box0.changeCorrection.dCorrection"]) <= TOL
    # the reference sums are never written by trial deltas
    rR, _ = e.get_recip_sums(0, eng.SUM_REF, nk)
    assert np.max(np.abs(rR - d["box0.sumRref"])) <= TOL * scale


def test_reciprocal_force_and_torque(gold):
    d, e = gold
    if not d["ff.ewald"][0]:
        pytest.skip("no Ewald")
    e.copy_recip(0)
    e.box_force(0)
    e.box_force_reciprocal(0)
    e.calculate_torque(0)
    gR, gM, gT = (e.get_forces(w) for w in (eng.ATOM_FORCE_REC, eng.MOL_FORCE_REC,
                                            eng.MOL_TORQUE))
    for i, c in enumerate("xyz"):
        assert rel_err(gR[i], d[f"box0.BoxForceReciprocal.atomForceRec.{c}"]) <= TOL
        assert rel_err(gM[i], d[f"box0.BoxForceReciprocal.molForceRec.{c}"]) <= TOL
        assert rel_err(gT[i], d[f"box0.CalculateTorque.molTorque.{c}"]) <= TOL


def test_cbmc_growth_step(gold):
    """CBMC growth of the last site of a chain molecule (pentane fixtures): ParticleInter on a
    batch of trial positions of site `part` and ParticleNonbonded against the sites that are
    already built (src/CalculateEnergy.cpp:689-785), against the reference's dumps."""
    d, e = gold
    if "box0.grow.part" not in d:
        pytest.skip("no chain molecule")
    ms = d["molStart"]
    kind, q = d["particleKind"], d["particleCharge"]
    x, y, z = _xyz(d, "coords")
    m, part = int(d["box0.grow.mol"][0]), int(d["box0.grow.part"][0])
    partners = d["box0.grow.partners"].astype(int) + int(ms[m])
    a = int(ms[m]) + part
    tp = _xyz(d, "box0.grow.trialPos")
    nb = e.particle_nonbonded(0, kind[a], q[a], kind[partners], q[partners], x[partners],
                              y[partners], z[partners], *tp)
    assert rel_err(nb, d["box0.grow.ParticleNonbonded"]) <= TOL
    # the call increments: a second call on the same buffer doubles it
    en, re, ov = e.particle_inter(0, m, part, *tp)
    assert np.array_equal(ov.astype(np.int32), d["box0.grow.ParticleInter.overlap"])
    assert rel_err(en, d["box0.grow.ParticleInter.en"]) <= TOL
    if np.any(d["box0.grow.ParticleInter.real"] != 0):
        assert rel_err(re, d["box0.grow.ParticleInter.real"]) <= TOL
