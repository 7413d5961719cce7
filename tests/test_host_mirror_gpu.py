"""GPU test of the C++ host mirror (gomc_b200/host/GomcB200.h): a compiled driver
uses the reference-named classes (CalculateEnergy, EwaldCached, NoEwald) the way
System::Init and Translate::CalcEn/Accept do; results are checked against the
oracle replaying the same accepted-move sequence."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

from gomc_b200 import synth
from tests.helpers import box_atoms, box_mols, oracle_for, random_move

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "gomc_b200", "host", "host_mirror_test")
TOL = 1e-9


def _write(path, s, moves):
    sig, eps, nn = s.ff.tables()
    with open(path, "wb") as f:
        f.write(struct.pack("<6i", s.n_atoms, s.n_mols, len(s.ff.type_names), s.ff.vdw_kind,
                            int(s.ff.ewald), len(moves)))
        f.write(np.array([s.ff.r_cut, s.ff.r_cut_coulomb, s.ff.r_cut_low, s.ff.r_switch,
                          s.ff.alpha, s.ff.recip_rcut, *s.axis], dtype="<f8").tobytes())
        for a in (sig, eps, nn, s.x, s.y, s.z, s.charge, *s.com()):
            f.write(np.asarray(a, dtype="<f8").tobytes())
        for a in (s.kind, s.mol, s.mol_start):
            f.write(np.asarray(a, dtype="<i4").tobytes())
        for m, (nx, ny, nz) in moves:
            f.write(struct.pack("<i", m))
            for a in (nx, ny, nz):
                f.write(np.asarray(a, dtype="<f8").tobytes())


@pytest.mark.parametrize("name", ["spce", "argon"])
def test_host_mirror_move_sequence(tmp_path, name):
    assert os.path.exists(EXE), "run __graft_entry__.build()"
    s = synth.make_spce(343, r_cut=8.0) if name == "spce" else synth.make_argon(500, r_cut=8.0)
    o = oracle_for(s)
    rng = np.random.default_rng(17)
    x, y, z = s.x.copy(), s.y.copy(), s.z.copy()
    moves = []
    for t in range(8):
        m = int(rng.integers(s.n_mols))
        s.x, s.y, s.z = x, y, z          # random_move reads the current coordinates
        moves.append((m, random_move(s, rng, m, 0.35)))
        # the driver accepts every non-overlapping move: mirror that below
        sl = slice(s.mol_start[m], s.mol_start[m + 1])
        ba = box_atoms(s)
        ba = ba[(ba < s.mol_start[m]) | (ba >= s.mol_start[m + 1])]
        _, _, ov = o.molecule_inter(x, y, z, s.kind, s.mol, s.charge, ba, m, s.mol_start[m],
                                    sl.stop - sl.start, *moves[-1][1])
        if not ov:
            x, y, z = x.copy(), y.copy(), z.copy()
            x[sl], y[sl], z[sl] = moves[-1][1]
    s = synth.make_spce(343, r_cut=8.0) if name == "spce" else synth.make_argon(500, r_cut=8.0)
    inp = tmp_path / "in.bin"
    _write(str(inp), s, moves)
    out = subprocess.run([EXE, str(inp)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    r = json.loads(out.stdout)

    # initial SystemTotal pieces
    lj, re = o.box_inter(s.x, s.y, s.z, s.kind, s.mol, s.charge, box_atoms(s))
    assert abs(r["inter"] - lj) <= TOL * abs(lj)
    if s.ff.ewald:
        kx, ky, kz, hs, pf, _ = o.recip_init_orth()
        sR, sI = o.box_recip_sums(box_mols(s), s.mol_start, s.x, s.y, s.z, s.charge, kx, ky, kz)
        assert abs(r["real"] - re) <= TOL * abs(re)
        assert abs(r["recip"] - o.box_reciprocal(sR, sI, pf)) <= TOL * abs(r["recip"])
        assert abs(r["self"] - o.box_self(box_mols(s), s.mol_start, s.charge)) <= TOL * abs(r["self"])
        co = o.box_correction(box_mols(s), s.mol_start, s.x, s.y, s.z, s.charge)
        assert abs(r["correction"] - co) <= TOL * abs(co)
    else:
        assert r["recip"] == 0.0 and r["self"] == 0.0

    # replay the trial / accept sequence with the oracle
    x, y, z = s.x.copy(), s.y.copy(), s.z.copy()
    for (m, (nx, ny, nz)), g in zip(moves, r["moves"]):
        sl = slice(s.mol_start[m], s.mol_start[m + 1])
        ba = box_atoms(s)
        ba = ba[(ba < s.mol_start[m]) | (ba >= s.mol_start[m + 1])]
        dlj, dre, ov = o.molecule_inter(x, y, z, s.kind, s.mol, s.charge, ba, m,
                                        s.mol_start[m], sl.stop - sl.start, nx, ny, nz)
        assert g["mol"] == m and bool(g["overlap"]) == ov
        assert abs(g["dLJ"] - dlj) <= TOL * max(abs(dlj), 1.0)
        assert abs(g["dReal"] - dre) <= TOL * max(abs(dre), 1.0)
        if s.ff.ewald:
            assert abs(g["swapCorr"] - o.swap_correction(s.charge[sl], (nx, ny, nz))) \
                <= TOL * abs(g["swapCorr"])
            assert abs(g["swapSelf"] - o.swap_self(s.charge[sl])) <= TOL * abs(g["swapSelf"])
            if not ov:
                sR0, sI0 = o.box_recip_sums(box_mols(s), s.mol_start, x, y, z, s.charge, kx, ky, kz)
                e_new, _, _ = o.mol_reciprocal(s.charge[sl], (x[sl], y[sl], z[sl]), (nx, ny, nz),
                                               kx, ky, kz, pf, sR0, sI0)
                d_ref = e_new - o.box_reciprocal(sR0, sI0, pf)
                assert abs(g["dRecip"] - d_ref) <= 1e-9 * abs(e_new)
        if not ov:
            x, y, z = x.copy(), y.copy(), z.copy()
            x[sl], y[sl], z[sl] = nx, ny, nz
    # running sums stay consistent with a recomputation (RecalculateAndCheck, far tighter
    # than the reference's 1e-3)
    for k in ("inter", "real", "recip"):
        a, b = r["running"][k], r["recomputed"][k]
        assert abs(a - b) <= 1e-9 * max(abs(b), 1.0), k
    # the MultiParticle displacement step replayed with the oracle on the final coordinates
    g = r["mp"]
    bm, ba = box_mols(s), box_atoms(s)
    beta, lam, tmax = 1.0 / 300.0, 0.5, 0.02
    _, _, aF, mF = o.box_force(x, y, z, s.kind, s.mol, s.charge, ba, s.n_mols)
    mR = None
    if s.ff.ewald:
        sRc, sIc = o.box_recip_sums(bm, s.mol_start, x, y, z, s.charge, kx, ky, kz)
        _, mR = o.box_force_reciprocal(bm, s.mol_start, x, y, z, s.charge, kx, ky, kz, pf,
                                       sRc, sIc, s.n_mols)
    else:
        mR = [np.zeros(s.n_mols) for _ in range(3)]
    com = [np.zeros(s.n_mols) for _ in range(3)]
    k, inr, new, _ = o.mp_transform(0, bm, s.mol_start, (x, y, z), com, mF, mR, tmax, lam, beta,
                                    77, 123, 0)
    ljn, ren, _, mFn = o.box_force(*new, s.kind, s.mol, s.charge, ba, s.n_mols)
    assert abs(g["inter"] - ljn) <= TOL * abs(ljn)
    mRn = [np.zeros(s.n_mols) for _ in range(3)]
    if s.ff.ewald:
        assert abs(g["real"] - ren) <= TOL * abs(ren)
        sRn, sIn = o.box_recip_sums(bm, s.mol_start, *new, s.charge, kx, ky, kz)
        assert abs(g["recip"] - o.box_reciprocal(sRn, sIn, pf)) <= TOL * abs(g["recip"])
        _, mRn = o.box_force_reciprocal(bm, s.mol_start, *new, s.charge, kx, ky, kz, pf,
                                        sRn, sIn, s.n_mols)
    w = o.mp_coeff(bm, inr, mF, mR, mFn, mRn, k, tmax, lam, beta)
    assert abs(g["w"] - w) <= 1e-8 * abs(w)
    assert abs(g["inter_after_reject"] - r["recomputed"]["inter"]) <= 1e-12 * abs(g["inter"])
    lj2, re2 = o.box_inter(x, y, z, s.kind, s.mol, s.charge, box_atoms(s))
    assert abs(r["recomputed"]["inter"] - lj2) <= TOL * abs(lj2)


MGPU = os.path.join(ROOT, "gomc_b200", "host", "multi_gpu_test")


def _n_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
        return sum(1 for ln in out.splitlines() if ln.startswith("GPU "))
    except OSError:
        return 0


@pytest.mark.parametrize("name,world", [("spce", 2), ("spce", 4), ("argon", 2)])
def test_one_process_many_gpus(tmp_path, name, world):
    """ONE C++ process, one engine + host thread per GPU, gomcb200_set_comm: the collective
    lives inside the engine (NCCL on the engine streams), so a single-process host such as
    GOMC reaches every GPU through include/gomc_b200.h.  Every rank must return the complete
    energies, equal to the single-GPU ones (1e-12: the partial sums associate differently)
    and to the oracle (TOL).  Needs `world` GPUs on the box (gpurun --gpus N)."""
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    assert os.path.exists(MGPU), "run __graft_entry__.build()"
    # 4 096 molecules -> 128-point fine grid: 8 bricks along x, divisible by 2 and 4 ranks
    s = synth.make_spce(4096, r_cut=9.0) if name == "spce" else synth.make_argon(4000)
    p = tmp_path / "sys.bin"
    _write(p, s, [])
    r = subprocess.run([MGPU, str(p), str(world)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    o = oracle_for(s)
    olj, ore = o.box_inter(s.x, s.y, s.z, s.kind, s.mol, s.charge, box_atoms(s))
    orc = 0.0
    if s.ff.ewald:
        kx, ky, kz, hs, pf, _ = o.recip_init_orth()
        sR, sI = o.box_recip_sums(box_mols(s), s.mol_start, s.x, s.y, s.z, s.charge, kx, ky, kz)
        orc = o.box_reciprocal(sR, sI, pf)
    ref = (olj, ore, orc)
    for c in range(3):
        assert abs(res["single"][c] - ref[c]) <= TOL * max(abs(ref[c]), 1e-300)
    for rank in res["ranks"]:
        for c in range(3):
            assert abs(rank[c] - res["single"][c]) <= 1e-12 * max(abs(res["single"][c]), 1.0)
            assert abs(rank[c] - ref[c]) <= TOL * max(abs(ref[c]), 1e-300)
