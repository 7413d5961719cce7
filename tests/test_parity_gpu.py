"""GPU parity tests: the CUDA engine (through the C ABI) against the CPU oracle
on the same seeded inputs.  Tolerance: relative 1e-9 (BASELINE.json north_star),
written as TOL below; integer/flag outputs must match exactly."""
import numpy as np
import pytest

from gomc_b200 import synth
from gomc_b200 import engine as eng
from tests.helpers import (SMALL_SYSTEMS, box_atoms, box_mols, oracle_for, random_move,
                           rel_err)

pytestmark = pytest.mark.gpu
TOL = 1e-9

NAMES = list(SMALL_SYSTEMS)


@pytest.fixture(scope="module", params=NAMES)
def case(request):
    s = SMALL_SYSTEMS[request.param]()
    e = eng.Engine.from_system(s)
    o = oracle_for(s)
    yield s, e, o
    e.close()


def test_box_inter(case):
    s, e, o = case
    lj, re = e.box_inter(0)
    olj, ore = o.box_inter(s.x, s.y, s.z, s.kind, s.mol, s.charge, box_atoms(s))
    assert abs(lj - olj) <= TOL * abs(olj)
    assert abs(re - ore) <= TOL * max(abs(ore), 1e-300)
    # determinism: two calls, identical bits
    assert e.box_inter(0) == (lj, re)


def test_box_force(case):
    s, e, o = case
    lj, re = e.box_force(0)
    olj, ore, aF, mF = o.box_force(s.x, s.y, s.z, s.kind, s.mol, s.charge, box_atoms(s),
                                   s.n_mols)
    assert abs(lj - olj) <= TOL * abs(olj)
    assert abs(re - ore) <= TOL * max(abs(ore), 1e-300)
    gF = e.get_forces(eng.ATOM_FORCE)
    gM = e.get_forces(eng.MOL_FORCE)
    for c in range(3):
        assert rel_err(gF[c], aF[c]) <= TOL
        assert rel_err(gM[c], mF[c]) <= TOL
    # Newton's third law: total force vanishes (to rounding of the sum)
    tot = max(abs(float(np.sum(gF[c]))) for c in range(3))
    assert tot <= 1e-9 * max(float(np.max(np.abs(gF[0]))), 1.0) * s.n_atoms ** 0.5


def test_molecule_inter(case):
    s, e, o = case
    rng = np.random.default_rng(11)
    for t in range(6):
        m = int(rng.integers(s.n_mols))
        amp = 0.4 if t % 2 == 0 else 0.45 * float(min(s.axis))
        nx, ny, nz = random_move(s, rng, m, amp)
        if t == 5:  # forced overlap: land on another molecule's first atom
            other = (m + 1) % s.n_mols
            a0 = s.mol_start[other]
            sl = slice(s.mol_start[m], s.mol_start[m + 1])
            nx = s.x[sl] - s.x[sl][0] + s.x[a0] + 0.3
            ny = s.y[sl] - s.y[sl][0] + s.y[a0] + 0.2
            nz = s.z[sl] - s.z[sl][0] + s.z[a0] + 0.1
            if s.cell_basis is None:
                nx, ny, nz = (np.mod(v, s.axis[k]) for k, v in enumerate((nx, ny, nz)))
            else:
                u = np.mod(np.stack([nx, ny, nz], 1) @ s.cell_basis_inv, s.axis)
                nx, ny, nz = (np.minimum(u, np.nextafter(s.axis, 0)) @ s.cell_basis).T.copy()
        lj, re, ov = e.molecule_inter(0, m, nx, ny, nz)
        # the one-launch trial (trial.cuh) agrees with the separate calls
        flj, fre, fov, frc = e.molecule_trial(0, m, nx, ny, nz)
        assert (flj, fre, fov) == (lj, re, ov)
        if _ewald(s):
            rc = e.mol_reciprocal(0, m, nx, ny, nz)
            assert abs(frc - rc) <= 1e-13 * abs(rc)
        ba = box_atoms(s)
        ba = ba[(ba < s.mol_start[m]) | (ba >= s.mol_start[m + 1])]
        olj, ore, oov = o.molecule_inter(s.x, s.y, s.z, s.kind, s.mol, s.charge, ba, m,
                                         s.mol_start[m], s.mol_start[m + 1] - s.mol_start[m],
                                         nx, ny, nz)
        assert ov == oov
        assert abs(lj - olj) <= TOL * max(abs(olj), 1.0)
        assert abs(re - ore) <= TOL * max(abs(ore), 1.0)
    assert ov  # the forced-overlap move must have been flagged


def test_particle_inter(case):
    s, e, o = case
    rng = np.random.default_rng(5)
    m = int(rng.integers(s.n_mols))
    trials = 7
    tx, ty, tz = (rng.uniform(0, s.axis[d], trials) for d in range(3))
    if s.cell_basis is not None:   # uniform in the slanted cell
        tx, ty, tz = (np.stack([tx, ty, tz], 1) @ s.cell_basis).T.copy()
    en, re, ov = e.particle_inter(0, m, 0, tx, ty, tz)
    ba = box_atoms(s)
    ba = ba[(ba < s.mol_start[m]) | (ba >= s.mol_start[m + 1])]
    a0 = s.mol_start[m]
    oen, ore, oov = o.particle_inter(s.x, s.y, s.z, s.kind, s.mol, s.charge, ba, m,
                                     s.kind[a0], s.charge[a0], tx, ty, tz)
    assert np.array_equal(ov, oov)
    assert rel_err(en, oen) <= TOL
    if np.any(ore != 0):
        assert rel_err(re, ore) <= TOL


def _ewald(s):
    return s.ff.ewald and s.ff.electrostatic


def test_kvectors_bit_exact(case):
    s, e, o = case
    if not _ewald(s):
        pytest.skip("no Ewald")
    kx, ky, kz, hs, pf, kmax = o.recip_init_orth()
    assert e.nk == len(kx)
    g = e.get_kvectors(0, eng.K_REF, e.nk)
    for a, b in zip(g, (kx, ky, kz, hs, pf)):
        assert np.array_equal(a, b)          # index-compatible with the host list
    # the copy the kernels read (kx, ky, kz, |k|^2 regenerated on the device from the row
    # table) carries the same bits
    for which in (eng.K_REF, eng.K_NEW):
        for a, b in zip(e.get_kvectors(0, which | eng.K_DEVICE, e.nk), (kx, ky, kz, hs, pf)):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("algo", [0, 1, 2, 3, 5])
def test_box_reciprocal_sums(case, algo):
    s, e, o = case
    if not _ewald(s):
        pytest.skip("no Ewald")
    kx, ky, kz, hs, pf, kmax = o.recip_init_orth()
    sR, sI = o.box_recip_sums(box_mols(s), s.mol_start, s.x, s.y, s.z, s.charge, kx, ky, kz)
    eo = o.box_reciprocal(sR, sI, pf)
    e.set_recip_algo(algo)
    en = e.box_reciprocal_sums(0)
    gR, gI = e.get_recip_sums(0, eng.SUM_NEW, e.nk)
    e.set_recip_algo(4)
    scale = max(np.max(np.abs(sR)), np.max(np.abs(sI)))
    assert np.max(np.abs(gR - sR)) <= TOL * scale
    assert np.max(np.abs(gI - sI)) <= TOL * scale
    assert abs(en - eo) <= TOL * abs(eo)
    assert abs(e.box_reciprocal(0, False) - eo) <= TOL * abs(eo)


def test_recip_algo_auto_is_nufft(case):
    """Algorithm 4 (the default) is the non-uniform FFT (algorithm 5) for an orthogonal box and
    the per-term kernel for a slanted one: same bits as the explicitly selected kernel.  The
    non-uniform FFT itself agrees with the FP64-MMA sum to 1e-11 of max |S| (window error
    ~1e-13), far inside the 1e-9 bar, and the INT8 kernel to 1e-10."""
    s, e, o = case
    if not _ewald(s):
        pytest.skip("no Ewald")

    def sums(algo):
        e.set_recip_algo(algo)
        e.mark_coords_changed()
        en = e.box_reciprocal_sums(0)
        return (en,) + tuple(e.get_recip_sums(0, eng.SUM_NEW, e.nk))

    try:
        slanted = getattr(s, "cell_basis", None) is not None
        fp64, i8, auto, nf = sums(2), sums(3), sums(4), sums(5)
        for a, b in zip(auto, sums(0) if slanted else nf):
            assert np.array_equal(a, b)
        assert abs(i8[0] - fp64[0]) <= 1e-10 * abs(fp64[0])
        scale = max(np.max(np.abs(fp64[1])), np.max(np.abs(fp64[2])))
        assert abs(nf[0] - fp64[0]) <= 1e-11 * abs(fp64[0])
        assert np.max(np.abs(nf[1] - fp64[1])) <= 1e-11 * scale
        assert np.max(np.abs(nf[2] - fp64[2])) <= 1e-11 * scale
        # bit-reproducible: gather spreading in sorted order, no atomics
        again = sums(5)
        for a, b in zip(again, nf):
            assert np.array_equal(a, b)
    finally:
        e.set_recip_algo(4)


def test_mol_and_swap_reciprocal(case):
    s, e, o = case
    if not _ewald(s):
        pytest.skip("no Ewald")
    kx, ky, kz, hs, pf, kmax = o.recip_init_orth()
    sR, sI = o.box_recip_sums(box_mols(s), s.mol_start, s.x, s.y, s.z, s.charge, kx, ky, kz)
    rng = np.random.default_rng(3)
    for t in range(3):
        m = int(rng.integers(s.n_mols))
        sl = slice(s.mol_start[m], s.mol_start[m + 1])
        nx, ny, nz = random_move(s, rng, m, 1.5)
        en = e.mol_reciprocal(0, m, nx, ny, nz)
        oe, oR, oI = o.mol_reciprocal(s.charge[sl], (s.x[sl], s.y[sl], s.z[sl]), (nx, ny, nz),
                                      kx, ky, kz, pf, sR, sI)
        assert abs(en - oe) <= TOL * abs(oe)
        gR, gI = e.get_recip_sums(0, eng.SUM_NEW, e.nk)
        assert rel_err(gR, oR) <= TOL and rel_err(gI, oI) <= TOL
        for insert in (1, 0):
            en = e.swap_reciprocal(0, m, nx, ny, nz, insert)
            oe, oR, oI = o.swap_recip(insert, s.charge[sl], (nx, ny, nz), kx, ky, kz, pf, sR, sI)
            assert abs(en - oe) <= TOL * abs(oe)
            # SwapCorrection / SwapSelf, alone and fused with the reciprocal delta
            oc, osf = o.swap_correction(s.charge[sl], (nx, ny, nz)), o.swap_self(s.charge[sl])
            co, se = e.swap_correction(0, m, nx, ny, nz)
            assert abs(co - oc) <= TOL * abs(oc) and abs(se - osf) <= TOL * abs(osf)
            ft = e.swap_trial(0, m, nx, ny, nz, insert)
            assert all(abs(a - b) <= 1e-13 * abs(b) for a, b in zip(ft, (en, co, se)))
            gR, gI = e.get_recip_sums(0, eng.SUM_NEW, e.nk)
            assert rel_err(gR, oR) <= TOL and rel_err(gI, oI) <= TOL
    # the reference sums must be untouched by trial moves (state machine rule 1)
    rR, rI = e.get_recip_sums(0, eng.SUM_REF, e.nk)
    assert rel_err(rR, sR) <= TOL and rel_err(rI, sI) <= TOL


@pytest.mark.parametrize("algo", [2, 5])   # FP64-MMA GEMM; type-2 non-uniform FFT
def test_force_reciprocal_and_torque(case, algo):
    s, e, o = case
    if not _ewald(s):
        pytest.skip("no Ewald")
    kx, ky, kz, hs, pf, kmax = o.recip_init_orth()
    sR, sI = o.box_recip_sums(box_mols(s), s.mol_start, s.x, s.y, s.z, s.charge, kx, ky, kz)
    e.copy_recip(0)
    e.box_force(0)
    e.set_recip_algo(algo)
    try:
        e.box_force_reciprocal(0)
    finally:
        e.set_recip_algo(4)
    e.calculate_torque(0)
    rF, mR = o.box_force_reciprocal(box_mols(s), s.mol_start, s.x, s.y, s.z, s.charge, kx, ky,
                                    kz, pf, sR, sI, s.n_mols)
    _, _, aF, _ = o.box_force(s.x, s.y, s.z, s.kind, s.mol, s.charge, box_atoms(s), s.n_mols)
    tq = o.calculate_torque(box_mols(s), s.mol_start, s.x, s.y, s.z, s.com(), aF, rF, s.n_mols)
    gR = e.get_forces(eng.ATOM_FORCE_REC)
    gM = e.get_forces(eng.MOL_FORCE_REC)
    gT = e.get_forces(eng.MOL_TORQUE)
    for c in range(3):
        assert rel_err(gR[c], rF[c]) <= TOL
        assert rel_err(gM[c], mR[c]) <= TOL
        assert rel_err(gT[c], tq[c]) <= TOL


def test_self_correction(case):
    s, e, o = case
    if not _ewald(s):
        pytest.skip("no Ewald")
    sf, co = e.box_self_correction(0)
    osf = o.box_self(box_mols(s), s.mol_start, s.charge)
    oco = o.box_correction(box_mols(s), s.mol_start, s.x, s.y, s.z, s.charge)
    assert abs(sf - osf) <= TOL * abs(osf)
    assert abs(co - oco) <= TOL * max(abs(oco), 1e-300)


def test_accept_state_machine(case):
    """Accepted move: set_molecule_coords + UpdateRecip must leave the engine in
    the state a full recomputation gives (SURVEY.md section 8b contract)."""
    s, e, o = case
    if not _ewald(s):
        pytest.skip("no Ewald")
    rng = np.random.default_rng(21)
    m = int(rng.integers(s.n_mols))
    sl = slice(s.mol_start[m], s.mol_start[m + 1])
    nx, ny, nz = random_move(s, rng, m, 0.8)
    e_new = e.mol_reciprocal(0, m, nx, ny, nz)
    old = (s.x[sl].copy(), s.y[sl].copy(), s.z[sl].copy())
    e.set_molecule_coords(m, nx, ny, nz)
    e.update_recip(0)
    full = e.box_reciprocal_sums(0)        # recompute from scratch into "new"
    assert abs(full - e_new) <= 1e-9 * abs(full)
    x2, y2, z2 = s.x.copy(), s.y.copy(), s.z.copy()
    x2[sl], y2[sl], z2[sl] = nx, ny, nz
    olj, ore = o.box_inter(x2, y2, z2, s.kind, s.mol, s.charge, box_atoms(s))
    lj, re = e.box_inter(0)
    assert abs(lj - olj) <= TOL * abs(olj) and abs(re - ore) <= TOL * abs(ore)
    # undo for the other tests of this module
    e.set_molecule_coords(m, *old)
    e.box_reciprocal_sums(0)
    e.update_recip(0)


@pytest.mark.parametrize("algo", [2, 3])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_partials_sum_to_full(case, world, algo):
    """Multi-GPU sharding, emulated on one GPU: the partial energies of the ranks
    add up to the unsharded result and every S(k) is owned by exactly one rank
    (default FP64-MMA structure factor and the opt-in INT8 one)."""
    s, e, o = case
    e.set_recip_algo(algo)
    lj0, re0, rc0 = e.call_full_box_energy(0)
    fR, fI = (e.get_recip_sums(0, eng.SUM_NEW, e.nk) if _ewald(s) else (None, None))
    lj = re = rc = 0.0
    owned = np.zeros(e.nk, dtype=np.int32)
    try:
        for r in range(world):
            e.set_shard(r, world)
            a, b, c = e.call_full_box_energy(0)
            lj, re, rc = lj + a, re + b, rc + c
            if _ewald(s):
                pR, pI = e.get_recip_sums(0, eng.SUM_NEW, e.nk)
                mine = (pR != 0.0) | (pI != 0.0)
                owned += mine
                # same k, different atom-slab split: equal up to summation order
                scale = max(np.max(np.abs(fR)), np.max(np.abs(fI)))
                if mine.any():          # a box with fewer INT8 tiles than ranks leaves some idle
                    assert np.max(np.abs(pR[mine] - fR[mine])) <= 1e-12 * scale
                    assert np.max(np.abs(pI[mine] - fI[mine])) <= 1e-12 * scale
    finally:
        e.set_shard(0, 1)
        e.set_recip_algo(4)
    assert abs(lj - lj0) <= 1e-12 * abs(lj0)
    assert abs(re - re0) <= 1e-12 * max(abs(re0), 1.0)
    if _ewald(s):
        assert abs(rc - rc0) <= 1e-12 * abs(rc0)
        assert owned.max() <= 1
        # k-vectors with S(k) == 0 exactly are indistinguishable from "not owned"
        assert np.count_nonzero(owned == 0) <= np.count_nonzero((fR == 0.0) & (fI == 0.0))
    e.call_full_box_energy(0)


def test_int8_structure_factor_large_box():
    """The INT8 tensor-core structure factor (recip algorithm 3) on the 100k-atom box of
    BASELINE configs[3]: five-slice path (c range > 32), many atom chunks and tiles.  No CPU
    oracle finishes this size in seconds, so the FP64 MMA kernel -- itself held to the oracle
    and to the reference's dumps on the small systems -- is the comparison; 1e-9 as everywhere."""
    from gomc_b200 import synth
    s = synth.make_spce(33334)
    e = eng.Engine.from_system(s)
    try:
        out = {}
        for algo in (2, 3):
            e.set_recip_algo(algo)
            e.mark_coords_changed()
            en = e.box_reciprocal_sums(0)
            out[algo] = (en,) + tuple(e.get_recip_sums(0, eng.SUM_NEW, e.nk))
        scale = max(np.max(np.abs(out[2][1])), np.max(np.abs(out[2][2])))
        assert abs(out[3][0] - out[2][0]) <= TOL * abs(out[2][0])
        assert np.max(np.abs(out[3][1] - out[2][1])) <= TOL * scale
        assert np.max(np.abs(out[3][2] - out[2][2])) <= TOL * scale
        # deterministic: integer accumulation, fixed-order FP64 reduction
        e.mark_coords_changed()
        assert e.box_reciprocal_sums(0) == out[3][0]
    finally:
        e.close()


def test_literal_reciprocal_dropins(case):
    """The host-array forms of the reciprocal seam (what the reference's Call*GPU functions
    receive: k list, point charges, the moved molecule's old/new coordinates) against the
    resident-state entry points and the oracle."""
    s, e, o = case
    if not _ewald(s) or getattr(s, "cell_basis", None) is not None:
        pytest.skip("orthogonal Ewald boxes")
    kx, ky, kz, hs, pf, kmax = o.recip_init_orth()
    nk = len(kx)
    sR, sI = o.box_recip_sums(box_mols(s), s.mol_start, s.x, s.y, s.z, s.charge, kx, ky, kz)
    scale = max(np.max(np.abs(sR)), np.max(np.abs(sI)))
    # host k list in (new set), explicit charges in, sums out
    e.set_kvectors(0, kx, ky, kz, hs, pf)
    for a, b in zip(e.get_kvectors(0, eng.K_NEW | eng.K_DEVICE, nk), (kx, ky, kz, hs, pf)):
        assert np.array_equal(a, b)
    for algo in (0, 2, 5):
        e.set_recip_algo(algo)
        en, gR, gI = e.call_box_reciprocal_points(0, True, s.x, s.y, s.z, s.charge, nk)
        assert np.max(np.abs(gR - sR)) <= TOL * scale and np.max(np.abs(gI - sI)) <= TOL * scale
        assert abs(en - o.box_reciprocal(sR, sI, pf)) <= TOL * abs(en)
    e.set_recip_algo(4)
    e.set_recip_ref(0)
    # a shuffled point order gives the same sums (only the summation order changes)
    perm = np.random.default_rng(5).permutation(s.n_atoms)
    en2, gR2, gI2 = e.call_box_reciprocal_points(0, False, s.x[perm], s.y[perm], s.z[perm],
                                                 s.charge[perm], nk)
    assert np.max(np.abs(gR2 - sR)) <= TOL * scale
    # moved molecule by explicit coordinates == by index on the resident state
    rng = np.random.default_rng(9)
    m = int(rng.integers(s.n_mols))
    sl = slice(s.mol_start[m], s.mol_start[m + 1])
    nx, ny, nz = random_move(s, rng, m, 1.0)
    e.box_reciprocal_sums(0)
    e.update_recip(0)          # resident-atom sums as the reference state
    want = e.mol_reciprocal(0, m, nx, ny, nz)
    wR, wI = e.get_recip_sums(0, eng.SUM_NEW, nk)
    en, gR, gI = e.call_mol_reciprocal(0, s.charge[sl], (s.x[sl], s.y[sl], s.z[sl]),
                                       (nx, ny, nz), nk)
    assert en == want and np.array_equal(gR, wR) and np.array_equal(gI, wI)
    for insert in (1, 0):
        want = e.swap_reciprocal(0, m, nx, ny, nz, insert)
        en, gR, gI = e.call_swap_reciprocal(0, s.charge[sl], (nx, ny, nz), insert, nk)
        assert en == want
    # host sums -> device (CallMolExchangeReciprocalGPU)
    e.set_recip_sums(0, eng.SUM_NEW, sR * 0.5, sI * 0.25)
    bR, bI = e.get_recip_sums(0, eng.SUM_NEW, nk)
    assert np.array_equal(bR, sR * 0.5) and np.array_equal(bI, sI * 0.25)
    # restore the module's state
    e.box_reciprocal_sums(0)
    e.update_recip(0)


# ---- the two pair-sweep kernels (gomcb200_set_pair_algo) ---------------------------------
def _pair_outputs(e, s):
    en = e.box_inter(0)
    fe = e.box_force(0)
    f = [np.array(c) for c in e.get_forces(eng.ATOM_FORCE)]
    e.set_com(*s.com())
    vir = np.concatenate(e.box_inter_virial(0))
    return en, fe, f, vir


@pytest.mark.parametrize("name", ["spce_small", "spce_mid", "argon", "mixture_shift",
                                  "mixture_exp6", "mixture_martini"])
def test_pair_kernels_agree(name):
    """k_pair_box2 (TMA staging, FP32 filter, tabulated Ewald real-space terms) against the
    first kernel on the same coordinates: the filter never decides InRcut, so the pair sets
    are identical and energies / forces / virial agree to rounding (1e-12), far inside TOL."""
    s = SMALL_SYSTEMS[name]()
    e = eng.Engine.from_system(s)
    try:
        e.set_pair_algo(0)
        a = _pair_outputs(e, s)
        e.set_pair_algo(1)
        b = _pair_outputs(e, s)
        for k in (0, 1):
            for x, y in zip(a[k], b[k]):
                assert abs(x - y) <= 1e-12 * max(abs(x), 1.0), (name, k, x, y)
        for c in range(3):
            assert rel_err(b[2][c], a[2][c]) <= 1e-12
        assert rel_err(b[3], a[3]) <= 1e-12
        assert _pair_outputs(e, s)[0] == b[0]      # bit-reproducible
    finally:
        e.close()


@pytest.mark.parametrize("dense_cells,per_cell", [(3, 300), (1, 600)])
def test_pair_sweep_dense_cells(dense_cells, per_cell):
    """Uneven boxes: a few cells far above the average population.  300 atoms in three
    neighbouring cells exceeds one staging pass of k_pair_box2 (several passes, the self range
    re-staged); 600 atoms in one cell exceeds what it can stage at all, so the kernel queued
    behind it as the fallback must take over (device-side gate, no host round trip)."""
    s = synth.make_argon(4000)                       # L = 57.3, 5 cells of 11.5 A per axis
    rng = np.random.default_rng(3)
    cs = float(s.axis[0]) / 5.0
    moved = rng.permutation(s.n_atoms)[:dense_cells * per_cell]
    for c in range(dense_cells):
        idx = moved[c * per_cell:(c + 1) * per_cell]
        lo = np.array([1.0 + c, 2.0, 2.0]) * cs
        # a jittered sub-lattice inside the cell keeps the pairs apart (r > 1 A)
        g = int(np.ceil(per_cell ** (1 / 3)))
        pts = np.stack(np.meshgrid(*[np.arange(g)] * 3, indexing="ij"), -1).reshape(-1, 3)
        pts = (pts[:per_cell] + 0.5 + rng.uniform(-0.1, 0.1, (per_cell, 3))) * (cs / g)
        s.x[idx], s.y[idx], s.z[idx] = (lo + pts).T
    e = eng.Engine.from_system(s)
    o = oracle_for(s)
    try:
        olj, ore, aF, mF = o.box_force(s.x, s.y, s.z, s.kind, s.mol, s.charge, box_atoms(s),
                                       s.n_mols)
        res = {}
        for algo in (0, 1):
            e.set_pair_algo(algo)
            lj, _ = e.box_inter(0)
            flj, _ = e.box_force(0)
            f = [np.array(c) for c in e.get_forces(eng.ATOM_FORCE)]
            assert abs(lj - olj) <= TOL * abs(olj)
            assert abs(flj - olj) <= TOL * abs(olj)
            for c in range(3):
                assert rel_err(f[c], aF[c]) <= TOL
            res[algo] = (lj, flj)
        if dense_cells == 1:     # energy sweep fell back: the very same kernel ran
            assert res[0][0] == res[1][0]
    finally:
        e.close()


def test_full_box_energy_graph_replays():
    """gomcb200_call_full_box_energy runs on two streams and, from the fourth call on identical
    inputs, as a replayed CUDA graph.  Every call -- warm-up, capture, replays, and the calls
    after something the graph bakes in has changed (reference/new sum buffers exchanged by
    UpdateRecip, another structure-factor algorithm, a fractional molecule) -- must return what
    the separate entry points return on the same coordinates."""
    s = SMALL_SYSTEMS["spce_mid"]()
    e = eng.Engine.from_system(s)
    rng = np.random.default_rng(21)
    try:
        def check(tag):
            x = np.mod(s.x + rng.uniform(-0.05, 0.05, s.n_atoms), s.axis[0])
            y = np.mod(s.y + rng.uniform(-0.05, 0.05, s.n_atoms), s.axis[1])
            z = np.mod(s.z + rng.uniform(-0.05, 0.05, s.n_atoms), s.axis[2])
            lj, re, rc = e.call_full_box_energy(0, x, y, z)
            lj2, re2 = e.box_inter(0)
            rc2 = e.box_reciprocal_sums(0)
            for a, b in ((lj, lj2), (re, re2), (rc, rc2)):
                assert abs(a - b) <= 1e-12 * max(abs(b), 1.0), (tag, a, b)
        for i in range(7):                 # plain, plain, plain, capture, replay ...
            check(("replay", i))
        e.update_recip(0)                  # new <-> ref sums exchanged: other buffers
        for i in range(5):
            check(("after update_recip", i))
        e.set_recip_algo(2)                # direct sum instead of the FFT
        for i in range(5):
            check(("algo 2", i))
        e.set_recip_algo(4)
        e.init_softcore(0.5, 3.0, 2, 1)
        e.update_lambda(0, 5, int(s.mol_kind[5]), 0.6, 0.4)   # fractional molecule
        for i in range(5):
            check(("lambda", i))
    finally:
        e.close()
