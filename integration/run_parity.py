"""GOMC itself on the gomc_b200 engine versus the reference's own CPU build.

Runs oracle/_ref/GOMC_CPU_<ENS> (the unmodified reference, +p1) and
oracle/_ref/GOMC_B200_<ENS> (the same unmodified sources linked against
integration/gomc_shim.cu + libgomc_b200.so instead of src/GPU/*.cu) on the same input and
seed, and compares what the reference's own regression harness compares
(test/Run_Examples.py:124-160: the output PDB, byte for byte) plus the per-step energies the
reference prints (builds without NDEBUG print every step) and the acceptance counters.

    python integration/run_parity.py [--mols 343] [--steps 10000] [--mp] [--json out.json]
Needs a GPU (the B200 executable) -- test infrastructure, not product code.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gomc_b200 import synth  # noqa: E402

STEP_RE = re.compile(r"^Step (\d+): Box (\d) Energies")
NUM_RE = re.compile(r"[-+]?\d+\.\d+(?:[eE][-+]?\d+)?")


def parse(log):
    """{(step, box): [Total, IntraB, IntraNB, Inter, LRC, TotalElectric, Real, Recip, Self, Corr]},
    MOVE_/ENER_ lines."""
    steps, counters = {}, []
    lines = log.splitlines()
    for i, ln in enumerate(lines):
        m = STEP_RE.match(ln)
        if m and i + 2 < len(lines):
            vals = [float(v) for v in NUM_RE.findall(lines[i + 1] + " " + lines[i + 2])]
            steps[(int(m.group(1)), int(m.group(2)))] = vals
        elif ln.startswith(("MOVE_", "ENER_", "STAT_")):
            counters.append(ln.split())
    return steps, counters


def run(exe, workdir, threads=1, env_extra=None):
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    env.update(env_extra or {})
    t0 = time.time()
    r = subprocess.run([exe, f"+p{threads}", "in.conf"], cwd=workdir, env=env,
                       capture_output=True, text=True)
    return r.returncode, r.stdout + r.stderr, time.time() - t0


def compare(mols=343, steps=10000, mp=False, ens="NVT", rcut=8.0, tol=1e-9, env_extra=None,
            charged=True):
    """ens "GEMC": BASELINE configs[2] in small -- TraPPE-UA n-pentane, a liquid and a vapour
    box, translate / rotate / CBMC regrowth / CBMC molecule swaps (MoleculeTransfer:
    SwapDestRecip, SwapSourceRecip, SwapCorrection x2, SwapSelf through the engine)."""
    cpu = os.path.join(ROOT, "oracle", "_ref", f"GOMC_CPU_{ens}")
    b200 = os.path.join(ROOT, "oracle", "_ref", f"GOMC_B200_{ens}")
    for exe in (cpu, b200):
        if not os.path.exists(exe):
            raise FileNotFoundError(exe + f" (make -f integration/Makefile ENS={ens})")
    second = None
    if ens == "GEMC":
        s = synth.make_pentane(mols, L=round((mols / 0.0038) ** (1 / 3), 3), charged=charged,
                               r_cut=rcut)
        second = synth.make_pentane(max(mols // 5, 8), L=50.0, seed=77, charged=charged,
                                    r_cut=rcut)
        out = {"system": f"TraPPE-UA n-pentane{' (partial charges)' if charged else ''}: "
                         f"{s.n_mols} molecules in box 0 (L {float(s.axis[0])}), {second.n_mols} "
                         f"in box 1 (L 50), Rcut {rcut}",
               "steps": steps, "ensemble": "GEMC-NVT",
               "moves": "translate 0.5 / rotate 0.2 / CBMC regrowth 0.1 / CBMC swap 0.2"}
    else:
        s = synth.make_spce(mols, r_cut=rcut, r_cut_coulomb=rcut)
        out = {"system": f"SPC/E {mols} molecules, Rcut {rcut}", "steps": steps, "ensemble": ens,
               "moves": "translate 0.4 / rotate 0.4 / MultiParticle 0.2" if mp
                        else "translate 0.6 / rotate 0.4"}
    logs = {}
    with tempfile.TemporaryDirectory() as d:
        for tag, exe in (("cpu", cpu), ("b200", b200)):
            wd = os.path.join(d, tag)
            synth.write_gomc_inputs(s, wd, multiparticle=mp, run_steps=steps, second=second)
            conf = open(os.path.join(wd, "in.conf")).read()
            conf = conf.replace("RestartFreq false 1000", f"RestartFreq true {steps}")
            open(os.path.join(wd, "in.conf"), "w").write(conf)
            rc, log, secs = run(exe, wd, env_extra=env_extra if tag == "b200" else None)
            if rc != 0:
                raise RuntimeError(f"{tag} run failed:\n{log[-3000:]}")
            logs[tag] = log
            out[tag + "_seconds"] = round(secs, 2)
            out[tag + "_pdb"] = b"".join(
                open(os.path.join(wd, f"out_BOX_{b}_restart.pdb"), "rb").read()
                for b in range(2 if second is not None else 1))
    sc, cc = parse(logs["cpu"])
    sb, cb = parse(logs["b200"])
    out["steps_printed"] = len(sc)
    first_div, max_rel = None, 0.0
    for key in sorted(sc):
        if key not in sb:
            first_div = first_div or key[0]
            continue
        for a, b in zip(sc[key], sb[key]):
            rel = abs(a - b) / max(abs(a), 1.0)
            max_rel = max(max_rel, rel) if first_div is None else max_rel
            if rel > tol and first_div is None:
                first_div = key[0]
    out["first_divergent_step"] = first_div
    out["max_rel_energy_diff_before_divergence"] = max_rel
    out["counters_identical"] = cc == cb
    out["final_counters_cpu"] = [c for c in cc if c[0].startswith("MOVE_")][-1:]
    out["final_counters_b200"] = [c for c in cb if c[0].startswith("MOVE_")][-1:]
    out["pdb_identical"] = out.pop("cpu_pdb") == out.pop("b200_pdb")
    out["engine_banner"] = [ln for ln in logs["b200"].splitlines() if "GPU" in ln][:3]
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--mols", type=int, default=343)
    ap.add_argument("--steps", type=int, default=10000)
    ap.add_argument("--mp", action="store_true")
    ap.add_argument("--ens", default="NVT", choices=["NVT", "GEMC"])
    ap.add_argument("--cpu-only", action="store_true", help="smoke-run the CPU executable only")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    res = compare(a.mols, a.steps, a.mp, ens=a.ens)
    print(json.dumps(res, indent=1))
    if a.json:
        json.dump(res, open(a.json, "w"), indent=1)
