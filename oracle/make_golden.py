"""TEST INFRASTRUCTURE: regenerates tests/golden/*.npz by running the UNMODIFIED
reference (oracle/_ref/gomc_probe_NVT, built by oracle/ref_build.mk from
/root/reference) on small synthetic systems written by gomc_b200.synth.

Run in the build container only (needs /root/reference):
    python oracle/make_golden.py
The probe runs with OMP_NUM_THREADS=1 so the reference's summation order is
deterministic.  Each .npz holds the inputs as the reference parsed them
(coordinates, kinds, charges, tables, k-vectors) and the outputs of every
hot-path function.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gomc_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

CASES = {
    # name: (system factory, nMoves, seed)
    "spce100_rc7": (lambda: synth.make_spce(100, r_cut=7.0), 6, 7),
    "spce343_rc8": (lambda: synth.make_spce(343, r_cut=8.0, r_cut_coulomb=8.0), 6, 11),
    "argon256": (lambda: synth.make_argon(256, r_cut=7.0), 4, 3),
    "mixture_std": (lambda: synth.make_mixture(), 6, 5),
    "mixture_shift": (lambda: synth.make_mixture(vdw_kind=synth.VDW_SHIFT), 4, 5),
    "mixture_switch": (lambda: synth.make_mixture(vdw_kind=synth.VDW_SWITCH, r_switch=6.5), 4, 5),
    "mixture_n13": (lambda: synth.make_mixture(n_b_exp=13.0, seed=9), 4, 5),
    # EXP6: n is the exp-6 alpha (geometric mixing, truncated to uint in the energy);
    # every kind has a non-zero sigma (rMin = 0 gives NaN forces in the reference too)
    "mixture_exp6": (lambda: synth.make_mixture(vdw_kind=synth.VDW_EXP6, n_b_exp=16.0,
                                                du_eps=12.0, du_sigma=1.2), 4, 5),
    # Martini switch (ParaTypeMARTINI + Potential SWITCH), Ewald off -> switched Coulomb
    "mixture_martini": (lambda: synth.make_mixture(vdw_kind=synth.VDW_SWITCH, r_switch=6.0,
                                                   martini=True, ewald=False, n_b_exp=12.0), 4, 5),
    # non-orthogonal cell (80/75/70 degrees): BoxDimensionsNonOrth + RecipInitNonOrth
    "spce216_triclinic": (lambda: synth.make_spce(
        216, r_cut=6.0, cell_vectors=synth.triclinic_cell(20.494)), 6, 7),
    "mixture_martini_ewald": (lambda: synth.make_mixture(vdw_kind=synth.VDW_SWITCH, r_switch=6.0,
                                                         martini=True, ewald=True, seed=4, n_b_exp=12.0), 4, 5),
    # BASELINE configs[2]'s molecule: TraPPE-UA n-pentane (five united atoms, flexible), the
    # all-neutral alkane (reciprocal deltas = the degenerate zero) and a variant with partial
    # charges; the probe also grows the last site CBMC-style (ParticleInter +
    # ParticleNonbonded on the same trial positions)
    "pentane150": (lambda: synth.make_pentane(150), 6, 5),
    "pentane150q": (lambda: synth.make_pentane(150, charged=True), 6, 5),
    # one fractional molecule (free-energy / NeMTMC state): soft-core LJ, with and without
    # soft-core Coulomb.  Extra probe arguments:
    #   pick, lambdaVDW, lambdaCoulomb, sc_alpha, sc_sigma, sc_power, sc_coul
    "spce100_lambda": (lambda: synth.make_spce(100, r_cut=7.0), 6, 7,
                       (17, 0.6, 0.35, 0.5, 3.0, 2, 0)),
    # molecule 0 of kind 0: the only case where the reference's BoxSelf scales the
    # fractional molecule (it passes the kind index where a molecule index is expected)
    "spce100_lambda_mol0": (lambda: synth.make_spce(100, r_cut=7.0), 4, 7,
                            (0, 0.3, 0.55, 0.5, 3.0, 2, 1)),
    "mixture_shift_lambda_sccoul": (lambda: synth.make_mixture(vdw_kind=synth.VDW_SHIFT), 4, 5,
                                    (3, 0.45, 0.7, 0.5, 3.0, 2, 1)),
    "mixture_switch_lambda": (lambda: synth.make_mixture(vdw_kind=synth.VDW_SWITCH,
                                                         r_switch=6.5), 4, 5,
                              (40, 0.8, 0.5, 0.5, 3.0, 2, 1)),
}


def make_cached():
    """CachedFourier true (EwaldCached, BASELINE configs[1]): the same systems and probe
    sequence as spce100_rc7 / spce343_rc8 through the reference's CACHED reciprocal class --
    MolReciprocal as ref - cached + new with RestoreMol on rejection, SwapDest/SourceRecip
    handing the cached rows over, MultiParticle with backupMolCache / exgMolCache.  The probe
    skips what EwaldCached refuses (MolExchangeReciprocal, ChangeLambdaRecip, ChangeRecip)."""
    probe = os.path.join(ROOT, "oracle", "_ref", "gomc_probe_NVT")
    env = dict(os.environ, OMP_NUM_THREADS="1")
    for name in ("spce100_rc7", "spce343_rc8"):
        make, n_moves, seed = CASES[name][:3]
        s = make()
        with tempfile.TemporaryDirectory() as d:
            synth.write_gomc_inputs(s, d, cached_fourier=True)
            log = subprocess.run([probe, "golden", "in.conf", "dump.bin", str(n_moves), str(seed)],
                                 cwd=d, env=env, capture_output=True, text=True)
            if log.returncode != 0 or "Cache Ewald Fourier                Active" not in log.stdout:
                print(log.stdout[-3000:], log.stderr[-2000:])
                raise SystemExit(f"cached probe failed on {name}")
            dump = po.read_dump(os.path.join(d, "dump.bin"))
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cached_" + name + ".npz"), **dump)
        print(f"cached_{name}: {s.n_atoms} atoms, {len(dump)} arrays")


def make_npt():
    """NPT volume trials (one accepted, one rejected) through the reference's own
    VolumeTransfer object: tests/golden/npt_spce343.npz (probe mode `volume`)."""
    probe = os.path.join(ROOT, "oracle", "_ref", "gomc_probe_NPT")
    if not os.path.exists(probe):
        subprocess.check_call(["make", "-s", "-f", "oracle/ref_build.mk", "ENS_LIST=NPT", "-j8"],
                              cwd=ROOT)
    s = synth.make_spce(343, r_cut=8.0, r_cut_coulomb=8.0)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    with tempfile.TemporaryDirectory() as d:
        synth.write_gomc_inputs(s, d, npt=True)
        log = subprocess.run([probe, "volume", "in.conf", "dump.bin", "260.0", "-340.0"],
                             cwd=d, env=env, capture_output=True, text=True)
        if log.returncode != 0:
            print(log.stdout[-3000:], log.stderr[-2000:])
            raise SystemExit("NPT probe failed")
        dump = po.read_dump(os.path.join(d, "dump.bin"))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "npt_spce343.npz"), **dump)
    print(f"npt_spce343: {s.n_atoms} atoms, {len(dump)} arrays, nk {int(dump['box0.nk'][0])} -> "
          f"{int(dump['trial0.nk'][0])} (accepted), {int(dump['trial1.nk'][0])} (rejected)")


def main():
    if sys.argv[1:] == ["npt"]:
        return make_npt()
    if sys.argv[1:] == ["cached"]:
        return make_cached()
    probe = os.path.join(ROOT, "oracle", "_ref", "gomc_probe_NVT")
    if not os.path.exists(probe):
        subprocess.check_call(["make", "-s", "-f", "oracle/ref_build.mk", "ENS_LIST=NVT", "-j8"],
                              cwd=ROOT)
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    only = set(sys.argv[1:])
    for name, case in CASES.items():
        if only and name not in only:
            continue
        make, n_moves, seed = case[:3]
        extra = [str(v) for v in case[3]] if len(case) > 3 else []
        s = make()
        with tempfile.TemporaryDirectory() as d:
            synth.write_gomc_inputs(s, d)
            log = subprocess.run([probe, "golden", "in.conf", "dump.bin", str(n_moves), str(seed)]
                                 + extra, cwd=d, env=env, capture_output=True, text=True)
            if log.returncode != 0:
                print(log.stdout[-3000:], log.stderr[-2000:])
                raise SystemExit(f"probe failed on {name}")
            dump = po.read_dump(os.path.join(d, "dump.bin"))
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **dump)
        print(f"{name}: {s.n_atoms} atoms, {len(dump)} arrays, "
              f"nk={int(dump['box0.nk'][0]) if 'box0.nk' in dump else 0}")


if __name__ == "__main__":
    main()
