#!/usr/bin/env python
"""bench.py -- the judged benchmark of the GOMC B200 energy/force engine.

Metric (BASELINE.json): full-box energy evaluations per second, one evaluation
= BoxInter + BoxReciprocalSums + BoxReciprocal, on
the synthetic SPC/E box BASELINE's target is quoted on (100 002 atoms,
Rcut = RcutCoulomb = 10 A, Tolerance 1e-5 -> 102 978 k-vectors).

  python bench.py --gpus N --steps K --warmup W [--workload spce100k|spce10k|argon4k|electrolyte1m]
                  [--recip-algo 5|2|3]  # structure factor: non-uniform FFT (default) / FP64-DMMA /
                                        # INT8 direct sums
                  [--pair-algo 1|0]     # pair sweep: k_pair_box2 (default) / the first kernel
  python bench.py --impl reference ...   # the reference's own CPU path on the host cores

One JSON line on stdout (rank 0).  `value` is timed with inputs resident in
HBM (cell binning + packing of the coordinates included every step); `e2e`
goes through the host-buffer C-ABI call (pinned host coordinates in, three
doubles out, copies inside the timed region).  Multi-GPU (torchrun): cell slabs
and FFT x-slabs are sharded, coordinates replicated; the engine's own NCCL
communicator all-gathers the pruned FFT slabs and all-reduces the three
energies on the engine's stream; strong scaling (the box is fixed).  Every
line of the default workload also carries `cfg5`: the same step on the 1 M-atom
electrolyte box at this N.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "full-box energy evals/s (BoxInter+BoxReciprocal)"
UNIT = "evals/s"

WORKLOADS = {
    # name: (factory kwargs, description)
    "spce100k": dict(kind="spce", n_mols=33334),
    "spce10k": dict(kind="spce", n_mols=10000),
    "argon4k": dict(kind="argon", n_atoms=4000),
    "electrolyte1m": dict(kind="electrolyte"),
}


def make_system(name):
    from gomc_b200 import synth
    w = WORKLOADS[name]
    if w["kind"] == "spce":
        return synth.make_spce(w["n_mols"])
    if w["kind"] == "argon":
        return synth.make_argon(w["n_atoms"])
    return synth.make_electrolyte()


def workload_config(name, s, nk, extra=None):
    cfg = {"workload": f"{name}: NVT SPC/E-type synthetic box, {s.n_atoms} atoms, "
                       f"L={float(s.axis[0])} A, Rcut={s.ff.r_cut} RcutCoulomb={s.ff.r_cut_coulomb} "
                       f"Tolerance={s.ff.tolerance}, k-vectors={nk}",
           "atoms": s.n_atoms, "molecules": s.n_mols, "k_vectors": nk,
           "step": "BoxInter + BoxReciprocalSums + BoxReciprocal (one full-box LJ+Ewald energy)",
           "l2": "flushed between timed steps (256 MiB write)"}
    if extra:
        cfg.update(extra)
    return cfg


# --------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "power_w_max": float(max(pw)), "samples": len(sm), "reasons": sorted(reasons)}


def fp64_peak():
    """Measured FP64 FMA peak of this GPU (tools/fp64_peak, built by build())."""
    exe = os.path.join(ROOT, "tools", "fp64_peak")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout
        return json.loads(out.strip().splitlines()[-1])
    except Exception as ex:  # noqa: BLE001
        return {"error": str(ex)}


# --------------------------------------------------------------------------
def cpu_baseline_port(s, budget_s=12.0):
    """Oracle (C port of the reference algorithm, OpenMP) on the host cores, on a
    bounded sample: BoxInter on the full box + the structure factor on a k-slab,
    extrapolated linearly in the number of k-vectors."""
    from oracle import pyoracle as po
    o = po.Oracle.from_system(s)
    cores = po.max_threads()
    ba = np.arange(s.n_atoms, dtype=np.int32)
    bm = np.arange(s.n_mols, dtype=np.int32)
    t0 = time.perf_counter()
    o.box_inter(s.x, s.y, s.z, s.kind, s.mol, s.charge, ba)
    t_inter = time.perf_counter() - t0
    nk, t_slab, slab = 0, 0.0, 0
    if s.ff.ewald and s.ff.electrostatic:
        kx, ky, kz, hs, pf, _ = o.recip_init_orth()
        nk = len(kx)
        probe = max(8, nk // 2048)
        t0 = time.perf_counter()
        o.box_recip_sums(bm, s.mol_start, s.x, s.y, s.z, s.charge, kx, ky, kz, 0, probe)
        per_k = (time.perf_counter() - t0) / probe
        slab = int(min(nk, max(probe, budget_s / max(per_k, 1e-9))))
        t0 = time.perf_counter()
        o.box_recip_sums(bm, s.mol_start, s.x, s.y, s.z, s.charge, kx, ky, kz, 0, slab)
        t_slab = time.perf_counter() - t0
    t_full = t_inter + (t_slab * nk / slab if slab else 0.0)
    return {"value": 1.0 / t_full, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"ESTIMATED from a k-slab: BoxInter on the full box ({t_inter:.2f} s) + "
                      f"BoxReciprocalSums on {slab} of {nk} k-vectors ({t_slab:.2f} s), "
                      "extrapolated linearly in k"}


def reference_cpu_sample(s, reps=5):
    """The reference's own CPU implementation of the step on the host cores (the unmodified
    GOMC CPU build, oracle/_ref/gomc_probe_NVT, OpenMP over all cores) on a bounded sample:
    BoxInter over the full box + BoxReciprocalSums / BoxReciprocal on a k-slab sized for ~1.5 s,
    extrapolated linearly in the number of k-vectors.  None when the probe is not built."""
    probe = os.path.join(ROOT, "oracle", "_ref", "gomc_probe_NVT")
    if not os.path.exists(probe):
        return None
    from gomc_b200 import synth
    from oracle import pyoracle as po
    cores = os.cpu_count() or 1
    nk_est = 0
    if s.ff.ewald:
        nk_est = len(po.Oracle.from_system(s).recip_init_orth()[0])
    k_frac = max(1, int(nk_est * s.n_atoms * 40e-9 / cores / 1.5)) if nk_est else 1
    with tempfile.TemporaryDirectory() as d:
        synth.write_gomc_inputs(s, d)
        env = dict(os.environ, OMP_NUM_THREADS=str(cores))
        r = subprocess.run([probe, "time", "in.conf", "dump.bin", str(k_frac), str(reps)],
                           cwd=d, env=env, capture_output=True, text=True)
        if r.returncode != 0:
            return {"error": "reference probe failed: " + r.stderr[-200:]}
        dmp = po.read_dump(os.path.join(d, "dump.bin"))
    nk_full, nk_slab = int(dmp["time.nkFull"][0]), int(dmp["time.nkSlab"][0])
    t_inter = float(np.median(dmp["time.BoxInter"]))
    t_sums = float(np.median(dmp["time.BoxReciprocalSums.slab"])) if nk_full else 0.0
    t_rec = float(np.median(dmp["time.BoxReciprocal.slab"])) if nk_full else 0.0
    scale = nk_full / nk_slab if nk_slab else 0.0
    t_full = t_inter + (t_sums + t_rec) * scale
    return {"value": 1.0 / t_full, "unit": UNIT, "cores": int(dmp["threads"][0]),
            "kind": "reference",
            "sample": f"ESTIMATED from a k-slab: unmodified GOMC CPU build (oracle/_ref): BoxInter "
                      f"full box ({t_inter:.3f} s) + BoxReciprocalSums/BoxReciprocal on {nk_slab} "
                      f"of {nk_full} k-vectors ({t_sums:.3f} s), extrapolated linearly in k; "
                      f"median of {reps} steps",
            "nk": nk_full}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation on the host cores
    (oracle/_ref probe when it was built, else the C port)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    s = make_system(args.workload)
    from oracle import pyoracle as po
    nk_est = 0
    if s.ff.ewald:
        nk_est = len(po.Oracle.from_system(s).recip_init_orth()[0])
    steps, warm = args.steps, args.warmup
    line = {"metric": METRIC, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, s, nk_est)}
    cb = reference_cpu_sample(s, reps=max(1, min(steps, 8)))
    if cb is not None and "error" in cb:
        line["unavailable"] = cb["error"]
        print(json.dumps(line))
        return 0
    if cb is None:
        cb = cpu_baseline_port(s, budget_s=8.0)
    cb.pop("nk", None)
    val = cb["value"]
    line.update({"value": val, "ms_per_step": 1e3 / val,
                 "cpu_baseline": cb,
                 "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0,
                         "d2h_bytes_per_step": 0},
                 "gpu_launches": 0})
    print(json.dumps(line))
    return 0


def reference_gpu_build(s, reps=5, force=True):
    """The reference's own shipped GPU build (src/GPU, compiled unmodified for sm_100 by
    oracle/ref_build.mk `gpu`) timed on this GPU: BoxInter + BoxReciprocalSums +
    BoxReciprocal over the FULL k list, host wall clock inside the probe as SURVEY.md
    section 8(d) prescribes.  A comparator, not an oracle.  None when not built."""
    probe = os.path.join(ROOT, "oracle", "_ref", "gomc_probe_GPU_NVT")
    if not os.path.exists(probe):
        return None
    from gomc_b200 import synth
    from oracle import pyoracle as po
    with tempfile.TemporaryDirectory() as d:
        synth.write_gomc_inputs(s, d)
        cmd = [probe, "time", "in.conf", "dump.bin", "1", str(reps)] + (["force"] if force else [])
        try:
            r = subprocess.run(cmd, cwd=d, capture_output=True, text=True, timeout=600)
        except subprocess.TimeoutExpired:
            return {"unavailable": "reference GPU probe timed out"}
        if r.returncode != 0:
            return {"unavailable": "reference GPU probe failed: " + (r.stderr or r.stdout)[-300:]}
        dmp = po.read_dump(os.path.join(d, "dump.bin"))
    med = lambda k: float(np.median(dmp[k])) if k in dmp and len(dmp[k]) else 0.0
    t_inter, t_sums, t_rec = med("time.BoxInter"), med("time.BoxReciprocalSums.slab"), \
        med("time.BoxReciprocal.slab")
    t = t_inter + t_sums + t_rec
    out = {"value": 1.0 / t, "unit": UNIT, "ms_per_step": 1e3 * t,
           "BoxInter_ms": 1e3 * t_inter, "BoxReciprocalSums_ms": 1e3 * t_sums,
           "BoxReciprocal_ms": 1e3 * t_rec,
           "energies": {"lj": float(dmp["time.lj"][0]), "real": float(dmp["time.real"][0]),
                        "recip": float(dmp["time.recipSlab"][0])},
           "what": "unmodified GOMC GPU build (CallBoxInterGPU + CallBoxReciprocalSumsGPU), "
                   f"same box, host buffers in, median of {reps}"}
    if force:
        # (the GPU probe runs BoxForceReciprocal over the full k list)
        out["BoxForce_ms"] = 1e3 * med("time.BoxForce")
        out["BoxForceReciprocal_ms"] = 1e3 * med("time.BoxForceReciprocal.slab8")
        out["multiparticle_ms"] = (out["BoxReciprocalSums_ms"] + out["BoxForce_ms"] +
                                   out["BoxForceReciprocal_ms"])
    return out


# --------------------------------------------------------------------------
def small_box_extras(eng, device, flush, reps=200):
    """E1/E3/E4 of SURVEY.md section 8(d) on BASELINE.json configs[1] (SPC/E, 10 000
    molecules, nk = 30 782): the single-molecule and swap trials are latency-bound,
    so they are wall-clock per call through the C ABI, host buffers in, scalars out."""
    import torch
    s = make_system("spce10k")
    e = eng.Engine.from_system(s, device=device)
    rng = np.random.default_rng(5)

    def trial_coords(m):
        sl = slice(s.mol_start[m], s.mol_start[m + 1])
        d = rng.uniform(-0.5, 0.5, 3)
        return s.x[sl] + d[0], s.y[sl] + d[1], s.z[sl] + d[2]

    def wall(fn, n, flush_each=False):
        t = []
        for i in range(n):
            if flush_each:
                flush.zero_()
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn(i)
            t.append((time.perf_counter() - t0) * 1e3)
        return float(np.mean(t))

    def e1(_):
        e.L.gomcb200_mark_coords_changed(e.h)
        e.call_full_box_energy(0)
    for i in range(3):
        e1(i)
    ms_e1 = wall(e1, 20, True)
    e.call_full_box_energy(0)
    e.set_recip_ref(0)
    mols = [int(m) for m in rng.integers(0, s.n_mols, reps + 10)]
    dp = C.POINTER(C.c_double)
    keep = []

    def ptrs(arrs):  # ctypes pointers made once: the loop times the C ABI, not numpy
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in arrs]
        keep.append(arrs)
        return [a.ctypes.data_as(dp) for a in arrs]
    moves = [ptrs(trial_coords(m)) for m in mols]
    olds = [ptrs((s.x[s.mol_start[m]:s.mol_start[m + 1]], s.y[s.mol_start[m]:s.mol_start[m + 1]],
                  s.z[s.mol_start[m]:s.mol_start[m + 1]])) for m in mols]
    lj, re, rc, co, se = (C.c_double() for _ in range(5))
    ov = C.c_int()
    r = [C.byref(v) for v in (lj, re, ov, rc, co, se)]
    Lib, h = e.L, e.h

    def e3(i):  # MoleculeInter + MolReciprocal, accept every other trial (UpdateRecip)
        m = mols[i]
        if Lib.gomcb200_molecule_trial(h, 0, m, *moves[i], r[0], r[1], r[2], r[3]):
            raise RuntimeError(Lib.gomcb200_last_error().decode())
        if i & 1:
            Lib.gomcb200_set_molecule_coords(h, m, *moves[i], None)
            Lib.gomcb200_update_recip(h, 0)
    for i in range(reps, reps + 10):
        e3(i)
    ms_e3 = wall(e3, reps)

    def e4(i):  # SwapDestRecip + SwapSourceRecip + SwapCorrection x2 (+ SwapSelf)
        m = mols[i]
        if (Lib.gomcb200_swap_trial(h, 0, m, *moves[i], 1, r[3], r[4], r[5]) or
                Lib.gomcb200_swap_trial(h, 0, m, *olds[i], 0, r[3], r[4], r[5])):
            raise RuntimeError(Lib.gomcb200_last_error().decode())
    for i in range(reps, reps + 10):
        e4(i)
    ms_e4 = wall(e4, reps)
    nk = e.nk
    e.close()
    return {"workload": "spce10k (BASELINE configs[1])", "n_atoms": int(s.n_atoms), "nk": int(nk),
            "E1_full_box_evals_per_s": 1e3 / ms_e1, "E1_ms": ms_e1,
            "E3_single_molecule_trials_per_s": 1e3 / ms_e3, "E3_ms": ms_e3,
            "E4_swap_trials_per_s": 1e3 / ms_e4, "E4_ms": ms_e4,
            "timing": "wall clock per call incl. H2D of the trial molecule and D2H of the scalars"}


def share_unique_id(eng, rank, world, dist, torch):
    """rank 0 makes the NCCL unique id of the engines' own communicator; the process group of
    the launcher only carries its 128 bytes to the other ranks."""
    if world == 1:
        return None
    t = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(eng.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, 0)
    return bytes(t.cpu().numpy().tobytes())


def time_full_box(e, world, dist, torch, flush, px, py, pz, steps, host):
    """`steps` timed evaluations.  Device time per step = CUDA events on the engine's stream
    around the whole call (binning, both sweeps, the in-engine NCCL all-reduce); wall time is
    taken next to it.  Returns (device ms, structure-factor stage ms, wall ms, energies)."""
    out = [C.c_double(), C.c_double(), C.c_double()]
    outp = [C.byref(o) for o in out]
    dev, dom, wall = [], [], []
    for _ in range(steps):
        flush.zero_()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        if host:
            rc = e.L.gomcb200_call_full_box_energy(e.h, 0, px, py, pz, *outp)
        else:
            e.L.gomcb200_mark_coords_changed(e.h)
            rc = e.L.gomcb200_call_full_box_energy(e.h, 0, None, None, None, *outp)
        if rc:
            raise RuntimeError(e.L.gomcb200_last_error().decode())
        torch.cuda.synchronize()
        wall.append((time.perf_counter() - t0) * 1e3)
        a, b = e.last_timing()
        dev.append(a)
        dom.append(b)
    return np.array(dev), np.array(dom), np.array(wall), [o.value for o in out]


def cfg5_record(eng, local, rank, world, dist, torch, flush, steps):
    """The box north_star's multi-GPU target is quoted on (1 M-atom electrolyte, 1.05 M
    k-vectors), same step, at this N: rides in every line so that the driver's scaling run
    carries the curve."""
    s = make_system("electrolyte1m")
    e = eng.Engine.from_system(s, device=local)
    e.set_comm(share_unique_id(eng, rank, world, dist, torch), rank, world)
    e.enable_timing(True)
    for _ in range(3):
        time_full_box(e, world, dist, torch, flush, None, None, None, 1, False)
    dev, dom, wall, en = time_full_box(e, world, dist, torch, flush, None, None, None, steps, False)
    nk = e.nk
    # the structure-factor stage alone (inside the step it shares the GPU with the pair sweep)
    t = []
    for _ in range(3):
        flush.zero_()
        torch.cuda.synchronize()
        e.mark_coords_changed()
        e.box_reciprocal_sums(0)
        t.append(e.last_timing()[1])
    dom = np.array(t)
    # E2 on the same box: the energy/force work of one MultiParticle trial
    e.set_com(*s.com())

    def mp_step():
        e.L.gomcb200_mark_coords_changed(e.h)
        e.box_reciprocal_sums(0)
        e.box_force(0)
        e.box_force_reciprocal(0)
        e.calculate_torque(0)
        e.get_forces(eng.MOL_TORQUE, 0, 1)
    mp_step()
    t_mp = []
    for _ in range(3):
        flush.zero_()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        mp_step()
        t_mp.append((time.perf_counter() - t0) * 1e3)
    e.close()
    return s, nk, float(np.mean(dev)), float(np.mean(wall)), float(np.mean(dom)), en, \
        float(np.mean(t_mp))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="spce100k", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--recip-algo", type=int, default=None, choices=[2, 3, 5],
                    help="structure-factor algorithm of the timed step: 5 = non-uniform FFT "
                         "(the engine's default for orthogonal boxes), 2 = FP64 DMMA direct "
                         "sum, 3 = INT8 tcgen05 byte-sliced direct sum")
    ap.add_argument("--pair-algo", type=int, default=None, choices=[0, 1],
                    help="pair-sweep kernel: 1 = k_pair_box2 (default), 0 = the first kernel")
    ap.add_argument("--no-ref-gpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from gomc_b200 import engine as eng, shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: gomc_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    s = make_system(args.workload)
    e = eng.Engine.from_system(s, device=local)
    # the engine's own NCCL communicator: cells and FFT slabs sharded, the three energies
    # all-reduced on the engine's stream (no collective in this script's timed region)
    e.set_comm(share_unique_id(eng, rank, world, dist, torch), rank, world)
    e.enable_timing(True)
    nk = e.nk
    n_charged = int(np.count_nonzero(np.abs(s.charge) >= 1e-9))
    if args.recip_algo is not None:
        e.set_recip_algo(args.recip_algo)
    if args.pair_algo is not None:
        e.set_pair_algo(args.pair_algo)
    recip_algo = args.recip_algo if args.recip_algo is not None else 5
    ewald = bool(s.ff.ewald and s.ff.electrostatic)

    # pinned host coordinates for the e2e leg
    hx, hy, hz = (torch.from_numpy(a.copy()).pin_memory() for a in (s.x, s.y, s.z))
    dp = C.POINTER(C.c_double)
    px, py, pz = (C.cast(t.data_ptr(), dp) for t in (hx, hy, hz))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def run(host, k):
        return time_full_box(e, world, dist, torch, flush, px, py, pz, k, host)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        run(False, 1)
        run(True, 1)
    torch.cuda.synchronize()
    # nvidia-smi needs ~1 s to deliver its first sample: keep the GPU under the same
    # load until it does, so that the clocks are those of the timed region
    def load_until_samples(n, max_s):
        # same step, untimed, until rank 0 holds n samples; every rank runs the same number
        # of steps (the step ends in a collective), rank 0 decides
        t_wait = time.perf_counter()
        while True:
            go = int(rank == 0 and len(sampler.lines) < n and time.perf_counter() - t_wait < max_s)
            if world > 1:
                flag = torch.tensor([go], device="cuda")
                dist.broadcast(flag, 0)
                go = int(flag.item())
            if not go:
                return
            run(False, 1)

    load_until_samples(1, 5.0)
    sampler.lines.clear()
    l0 = e.launch_count()
    dev_ms, dom_ms, wall_ms, en_res = run(False, args.steps)
    l1 = e.launch_count()
    dev_ms_h, _, wall_ms_h, en_host = run(True, args.steps)
    # a short timed region (sharded runs take tens of ms) can fall between two 100 ms
    # nvidia-smi samples: keep the same load running until three samples exist
    load_until_samples(3, 3.0)
    clocks = sampler.stop() if rank == 0 else None

    # ---- the dominant kernels timed alone (CUDA events on the engine's stream) ----------
    # pair sweep: second call on unchanged coordinates = the sweep kernel + its final reduce
    def alone(fn, k):
        fn()
        t = []
        for _ in range(k):
            flush.zero_()
            torch.cuda.synchronize()
            fn()
            t.append(e.last_timing()[0])
        return float(np.mean(t))
    pair_ms = alone(lambda: e.box_inter(0), max(5, args.steps // 2))
    # structure-factor stage alone (inside the step it runs on a second stream next to the
    # pair sweep): BoxReciprocalSums on re-packed coordinates, events around the stage
    recip_stage_ms = 0.0
    if ewald:
        t = []
        for _ in range(max(5, args.steps // 2)):
            flush.zero_()
            torch.cuda.synchronize()
            e.mark_coords_changed()
            e.box_reciprocal_sums(0)
            t.append(e.last_timing()[1])
        recip_stage_ms = float(np.mean(t))

    # secondary metric of BASELINE.json: MultiParticle moves/s = the energy/force work of
    # MultiParticle::CalcEn (src/moves/MultiParticle.h:414-441): BoxReciprocalSums + BoxForce
    # + BoxReciprocal + BoxForceReciprocal + torque, coordinates resident (single GPU)
    mp_ms = mp_move_ms = None
    if ewald and s.n_mols > 0:
        def mp_step():
            e.L.gomcb200_mark_coords_changed(e.h)
            e.box_reciprocal_sums(0)
            e.box_force(0)
            e.box_force_reciprocal(0)
            e.calculate_torque(0)
            e.get_forces(eng.MOL_TORQUE, 0, 1)      # forces stay on the device; sync point
        for _ in range(3):
            mp_step()
        t_mp = []
        for _ in range(max(5, args.steps // 2)):
            flush.zero_()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            mp_step()
            t_mp.append((time.perf_counter() - t0) * 1e3)
        mp_ms = shard.max_over_ranks(float(np.mean(t_mp)), world, "cuda")
        # whole MultiParticle move on the device: trial transform (Philox), CalcEn on the
        # trial set, acceptance weight, reject (pointer exchange back)
        e.set_com(*s.com())
        e.copy_recip(0)
        mp_step()                                    # reference forces / torques

        def mp_move(i):
            e.mp_transform(0, i & 1, 0.02, 0.5 / 300.0, 1000 + i, 0, 123)
            e.mp_select(1)
            e.box_reciprocal_sums(0)
            e.box_force(0)
            e.box_force_reciprocal(0)
            e.calculate_torque(0)
            w = e.mp_coeff(0, i & 1, 0.02, 0.5 / 300.0)
            e.mp_select(0)
            return w
        for i in range(2):
            mp_move(i)
        t_mv = []
        for i in range(max(5, args.steps // 2)):
            flush.zero_()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            mp_move(i)
            t_mv.append((time.perf_counter() - t0) * 1e3)
        mp_move_ms = shard.max_over_ranks(float(np.mean(t_mv)), world, "cuda")
    # the direct-sum structure-factor kernels of round 1 next to the default (same step)
    direct = None
    if world == 1 and ewald and recip_algo == 5 and not args.no_extras:
        direct = {}
        for algo, name in ((2, "fp64_dmma"), (3, "int8_tcgen05")):
            e.set_recip_algo(algo)
            for _ in range(2):
                run(False, 1)
            d3, m3, _, en3 = run(False, max(3, args.steps // 4))
            direct[name] = {"ms_per_step": float(np.mean(d3)),
                            "structure_factor_stage_ms": float(np.mean(m3)),
                            "recip_rel_diff_vs_default": abs(en3[2] - en_res[2]) / abs(en_res[2])}
        e.set_recip_algo(4)
        direct["what"] = ("gomcb200_set_recip_algo 2 / 3: the N x nk direct sums on FP64 DMMA and on "
                          "tcgen05 kind::i8 (DESIGN.md 4.1, 4.1b); the default is the non-uniform FFT")
    extras = None
    if world == 1 and args.workload == "spce100k" and not args.no_extras:
        extras = small_box_extras(eng, local, flush)
    e.close()
    cfg5 = None
    if args.workload == "spce100k" and not args.no_extras:
        s5, nk5, dev5, wall5, dom5, en5, mp5 = cfg5_record(eng, local, rank, world, dist, torch,
                                                           flush, max(3, min(args.steps, 5)))
        dev5, wall5, mp5 = (shard.max_over_ranks(v, world, "cuda") for v in (dev5, wall5, mp5))
        cfg5 = {"workload": f"electrolyte1m: {s5.n_atoms} atoms, L={float(s5.axis[0])} A, "
                            f"k-vectors={nk5} (BASELINE configs[4])",
                "n_gpus": world, "ms_per_step": dev5, "value": 1e3 / dev5, "unit": UNIT,
                "wall_ms_per_step": wall5, "structure_factor_stage_ms": dom5,
                "multiparticle_ms_per_step": mp5,
                "energies": {"lj": en5[0], "real": en5[1], "recip": en5[2]},
                "timing": "CUDA events on the engine stream, max over ranks; coordinates "
                          "resident, re-binned every step"}

    # One clock at every N: CUDA events on the engine's stream around the whole call (the
    # all-reduce is on that stream), max over ranks.  Wall time between synchronisations is
    # reported next to it.
    ms_res = shard.max_over_ranks(float(np.mean(dev_ms)), world, "cuda")
    ms_res_wall = shard.max_over_ranks(float(np.mean(wall_ms)), world, "cuda")
    ms_e2e = shard.max_over_ranks(float(np.mean(wall_ms_h)), world, "cuda")
    pair_ms = shard.max_over_ranks(pair_ms, world, "cuda")
    recip_stage_ms = shard.max_over_ranks(recip_stage_ms, world, "cuda")

    if rank == 0:
        peak = fp64_peak()
        dfma_tf, dmma_tf = peak.get("dfma_tflops"), peak.get("dmma_tflops")
        # pairs inside rc x ~60 flop + candidate tests x 11 flop
        vol = float(np.prod(s.axis))
        rho = s.n_atoms / vol
        rc = max(s.ff.r_cut, s.ff.r_cut_coulomb)
        p_in = 0.5 * s.n_atoms * (4.0 / 3.0) * np.pi * rc ** 3 * rho
        cells = [max(int(a // rc), 3) for a in s.axis]
        p_cand = 0.5 * 27.0 * (s.n_atoms / float(np.prod(cells))) * s.n_atoms
        pair_flops = (60.0 * p_in + 11.0 * p_cand) / world
        pair_ach = pair_flops / (pair_ms * 1e-3) * 1e-12
        pair_bytes = 40.0 * s.n_atoms / world
        traffic = None
        tj = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tj):
            traffic = json.load(open(tj)).get("k_pair_box2_dram_bytes")
        peaks = {}
        pj = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pj):
            peaks = json.load(open(pj))
        hbm_peak = peaks.get("hbm_gbs")   # driver-written copy bandwidth of this pool's B200s
        sf_flops = 4.0 * n_charged * nk
        line = {
            "metric": METRIC, "value": 1e3 / ms_res, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res,
            "wall_ms_per_step": ms_res_wall,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            # identical to the reference arm's `config` (the driver compares the two)
            "config": workload_config(args.workload, s, nk),
            "engine": {
                "parallelism": f"cell slabs + FFT x-slabs sharded over {world} GPU(s), coordinates "
                               "replicated, NCCL inside the engine (all-gather of the pruned FFT "
                               "slabs, all-reduce of 3 energies)",
                "recip_algo": recip_algo, "pair_algo": 1 if args.pair_algo is None else args.pair_algo,
                "timing": "CUDA events on the engine stream around the whole call, max over ranks"},
            "e2e": {"value": 1e3 / ms_e2e, "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": 3 * 8 * s.n_atoms, "d2h_bytes_per_step": 24,
                    "timing": "wall clock around gomcb200_call_full_box_energy, pinned host "
                              "coordinates in, three doubles out"},
            "gpu_launches": int(l1 - l0),
            "clocks": clocks,
            # the dominant kernel of the step: the pair sweep (43 % of the device time)
            "roofline": {"bound": "fp64",
                         "kernel": "k_pair_box2<VDW_STD, MODE_ENERGY> (+ its 4 us final reduce): "
                                   "BoxInter on unchanged coordinates, timed alone",
                         "achieved": pair_ach, "peak": dfma_tf, "unit": "TFLOP/s",
                         "frac": (pair_ach / dfma_tf) if dfma_tf else None,
                         "traffic": traffic,
                         "algorithmic_flops_per_launch": pair_flops,
                         "algorithmic": "SURVEY.md 8(d): 60 flop x pairs inside rc "
                                        f"({p_in:.3g}) + 11 flop x candidate pairs ({p_cand:.3g})",
                         "kernel_ms": pair_ms,
                         "peak_source": "tools/fp64_peak DFMA stream measured in this run "
                                        "(MEASURED_PEAKS.json has no FP64 entry); the kernel is "
                                        "bound by shared-memory wavefronts and issue slots, not "
                                        "by the FP64 pipe (profiles/r2_pair2_ncu.txt)",
                         "hbm": {"achieved_GBps": pair_bytes / (pair_ms * 1e-3) * 1e-9,
                                 "peak_GBps": hbm_peak,
                                 "algorithmic_bytes_per_launch": pair_bytes,
                                 "note": "40 B/atom once: the pair path is not HBM-bound by "
                                         "construction (SURVEY.md 8d)"}},
            "structure_factor": None if not ewald else {
                "stage_ms": recip_stage_ms,
                "algorithm": {5: "non-uniform FFT (type 1): DMMA spread, pruned FP64 FFT",
                              2: "direct sum, FP64 DMMA", 3: "direct sum, INT8 tcgen05"}[recip_algo],
                "direct_sum_equivalent_TFLOPs": sf_flops / (recip_stage_ms * 1e-3) * 1e-12,
                "direct_sum_flops": sf_flops,
                "dfma_peak_TFLOPs": dfma_tf, "dmma_peak_TFLOPs": dmma_tf,
                "note": "4 flop x charged atoms x k-vectors is the minimum of the DIRECT sum "
                        "(SURVEY.md 8d); the FFT path does ~20x less arithmetic, so this figure "
                        "is a speed ratio against that algorithm, not a pipe utilisation"},
            "multiparticle": (None if mp_ms is None else {
                "value": 1e3 / mp_ms, "unit": "MP energy/force evaluations per s",
                "ms_per_step": mp_ms,
                "step": "BoxReciprocalSums + BoxForce + BoxReciprocal + BoxForceReciprocal + "
                        "CalculateTorque (MultiParticle::CalcEn), wall clock incl. launches, max "
                        "over ranks; N > 1: cell-slab BoxForce + all-reduce of the atom forces, "
                        "slab-sharded structure factor, reciprocal force replicated",
                "full_move_ms": mp_move_ms,
                "full_move": "device trial transform + CalcEn on the trial set + GetCoeff + "
                             "reject, coordinates never leave the GPU"}),
            "direct_sum_kernels": direct,
            "small_box": extras,
            "cfg5": cfg5,
            "energies": {"lj": en_res[0], "real": en_res[1], "recip": en_res[2],
                         "host_path_identical": en_res == en_host},
        }
        if not args.no_cpu_baseline and world == 1:
            # the same estimator as `--impl reference` (the unmodified reference on the host
            # cores); the C port only where the reference probe is not built
            cb = reference_cpu_sample(s, reps=3)
            if cb is None or "error" in cb:
                cb = cpu_baseline_port(s)
            cb.pop("nk", None)
            line["cpu_baseline"] = cb
        if not args.no_ref_gpu and world == 1:
            line["reference_gpu_build"] = reference_gpu_build(s)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
