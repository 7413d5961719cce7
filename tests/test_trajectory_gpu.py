"""GPU tests of the move-level contract (north_star: identical accept/reject
sequences over a fixed-seed run) and of two-box (GEMC-style) bookkeeping.

A short NVT Monte-Carlo run of single-molecule translations is driven twice with
the same MT19937 stream (numpy RandomState is MT19937, like GOMC's PRNG): once with
energies from the CUDA engine through the C ABI, once with the CPU oracle.  The
Metropolis rule is the reference's (src/moves/Translate.h:97-105):
accept iff  u < exp(-beta * (dLJ + dReal + dRecip))  and no overlap."""
import numpy as np
import pytest

from gomc_b200 import engine as eng, synth
from tests.helpers import box_atoms, box_mols, oracle_for

pytestmark = pytest.mark.gpu


def _metropolis_run(s, steps, seed, use_gpu):
    rs = np.random.RandomState(seed)
    beta = 1.0 / 298.0                  # energies are in K
    x, y, z = s.x.copy(), s.y.copy(), s.z.copy()
    o = oracle_for(s)
    e = eng.Engine.from_system(s) if use_gpu else None
    ew = s.ff.ewald and s.ff.electrostatic
    if ew:
        kx, ky, kz, hs, pf, _ = o.recip_init_orth()
        sR, sI = o.box_recip_sums(box_mols(s), s.mol_start, x, y, z, s.charge, kx, ky, kz)
        e_recip = e.box_reciprocal(0, False) if use_gpu else o.box_reciprocal(sR, sI, pf)
    else:
        e_recip = 0.0
    decisions, deltas = [], []
    for _ in range(steps):
        m = int(rs.randint(s.n_mols))
        d = (rs.random_sample(3) - 0.5) * 0.6
        u = rs.random_sample()
        sl = slice(s.mol_start[m], s.mol_start[m + 1])
        nx = np.mod(x[sl] + d[0], s.axis[0])
        ny = np.mod(y[sl] + d[1], s.axis[1])
        nz = np.mod(z[sl] + d[2], s.axis[2])
        if use_gpu:
            dlj, dre, ov = e.molecule_inter(0, m, nx, ny, nz)
            e_new = e.mol_reciprocal(0, m, nx, ny, nz) if (ew and not ov) else e_recip
        else:
            ba = box_atoms(s)
            ba = ba[(ba < sl.start) | (ba >= sl.stop)]
            dlj, dre, ov = o.molecule_inter(x, y, z, s.kind, s.mol, s.charge, ba, m, sl.start,
                                            sl.stop - sl.start, nx, ny, nz)
            if ew and not ov:
                e_new, nR, nI = o.mol_reciprocal(s.charge[sl], (x[sl], y[sl], z[sl]),
                                                 (nx, ny, nz), kx, ky, kz, pf, sR, sI)
            else:
                e_new = e_recip
        dE = dlj + dre + (e_new - e_recip)
        acc = (not ov) and (u < np.exp(-beta * dE))
        decisions.append(bool(acc))
        deltas.append(dE)
        if acc:
            x[sl], y[sl], z[sl] = nx, ny, nz
            e_recip = e_new
            if use_gpu:
                e.set_molecule_coords(m, nx, ny, nz, [nx[0], ny[0], nz[0]])
                if ew:
                    e.update_recip(0)
            elif ew:
                sR, sI = nR, nI
    if use_gpu:
        e.close()
    return decisions, np.array(deltas), (x, y, z)


@pytest.mark.parametrize("name", ["spce", "argon"])
def test_accept_reject_sequence_matches_cpu(name):
    s = synth.make_spce(216, r_cut=7.0) if name == "spce" else synth.make_argon(500, r_cut=8.0)
    steps = 150
    dec_g, dE_g, xyz_g = _metropolis_run(s, steps, 123, True)
    dec_c, dE_c, xyz_c = _metropolis_run(s, steps, 123, False)
    assert dec_g == dec_c                       # identical accept / reject sequence
    assert 5 < sum(dec_g) < steps               # the run exercised both branches
    finite = np.isfinite(dE_c) & (np.abs(dE_c) < 1e12)
    assert np.max(np.abs(dE_g[finite] - dE_c[finite]) / np.maximum(np.abs(dE_c[finite]), 1.0)) <= 1e-9
    for a, b in zip(xyz_g, xyz_c):
        assert np.array_equal(a, b)             # same trajectory, bit for bit


def test_two_boxes_are_independent():
    """GEMC bookkeeping: two boxes in one engine (BOX_TOTAL = 2), global atom arrays,
    per-box molecule lists, per-box axes / k-vectors / structure factors; moving a
    molecule between the box lists changes both boxes as the oracle predicts."""
    a = synth.make_spce(125, r_cut=6.5, seed=3)
    b = synth.make_spce(64, r_cut=6.5, seed=4, density=0.02)
    ff = a.ff
    sig, eps, nn = ff.tables()
    n_a, n_m = a.n_atoms, a.n_mols
    x = np.concatenate([a.x, b.x]); y = np.concatenate([a.y, b.y]); z = np.concatenate([a.z, b.z])
    kind = np.concatenate([a.kind, b.kind]); q = np.concatenate([a.charge, b.charge])
    mol = np.concatenate([a.mol, b.mol + n_m]).astype(np.int32)
    ms = np.concatenate([a.mol_start, b.mol_start[1:] + n_a]).astype(np.int32)
    e = eng.Engine(2)
    e.init_forcefield(sig, eps, nn, ff.vdw_kind, len(ff.type_names), ff.r_cut,
                      [ff.r_cut_coulomb] * 2, ff.r_cut_low, 0.0, [ff.alpha] * 2, True, True)
    e.init_topology(kind, mol, q, ms)
    mols0 = list(range(n_m)); mols1 = list(range(n_m, n_m + b.n_mols))
    e.set_box_molecules(0, mols0); e.set_box_molecules(1, mols1)
    e.set_box_axes(0, a.axis); e.set_box_axes(1, b.axis)
    e.set_coords(x, y, z)
    e.init_ewald(0, [ff.recip_rcut] * 2)
    total = max(e.recip_count(0, a.axis), e.recip_count(1, b.axis))
    e.init_ewald(total, [ff.recip_rcut] * 2)
    oa, ob = oracle_for(a), oracle_for(b)
    for box, (s_, o_, mols) in enumerate(((a, oa, mols0), (b, ob, mols1))):
        nk, _ = e.recip_init(box, s_.axis)
        rc = e.box_reciprocal_setup(box)
        e.set_recip_ref(box)
        kx, ky, kz, hs, pf, _ = o_.recip_init_orth()
        sR, sI = o_.box_recip_sums(box_mols(s_), s_.mol_start, s_.x, s_.y, s_.z, s_.charge,
                                   kx, ky, kz)
        assert nk == len(kx)
        assert abs(rc - o_.box_reciprocal(sR, sI, pf)) <= 1e-9 * abs(rc)
        lj, re = e.box_inter(box)
        olj, ore = o_.box_inter(s_.x, s_.y, s_.z, s_.kind, s_.mol, s_.charge, box_atoms(s_))
        assert abs(lj - olj) <= 1e-9 * abs(olj) and abs(re - ore) <= 1e-9 * abs(ore)
    # transfer the last molecule of box 0 into box 1 at a new position (swap move accepted)
    m = n_m - 1
    sl = slice(ms[m], ms[m + 1])
    new = (np.mod(x[sl] * 0.5 + 3.0, b.axis[0]), np.mod(y[sl] * 0.5 + 2.0, b.axis[1]),
           np.mod(z[sl] * 0.5 + 1.0, b.axis[2]))
    e_dst = e.swap_reciprocal(1, m, *new, 1)           # SwapDestRecip in box 1
    e_src = e.swap_reciprocal(0, m, x[sl], y[sl], z[sl], 0)   # SwapSourceRecip in box 0
    e.update_recip(0); e.update_recip(1)
    e.set_molecule_coords(m, *new)
    e.set_box_molecules(0, mols0[:-1]); e.set_box_molecules(1, mols1 + [m])
    assert abs(e.box_reciprocal_sums(1) - e_dst) <= 1e-9 * abs(e_dst)
    assert abs(e.box_reciprocal_sums(0) - e_src) <= 1e-9 * abs(e_src)
    x2, y2, z2 = x.copy(), y.copy(), z.copy()
    x2[sl], y2[sl], z2[sl] = new
    ba1 = np.concatenate([np.arange(n_a, len(x)), np.arange(sl.start, sl.stop)]).astype(np.int32)
    olj, ore = ob.box_inter(x2, y2, z2, kind, mol, q, ba1)
    lj, re = e.box_inter(1)
    assert abs(lj - olj) <= 1e-9 * abs(olj) and abs(re - ore) <= 1e-9 * abs(ore)
    e.close()
