"""A/B of the two pair-sweep kernels (gomcb200_set_pair_algo 0 / 1): energies, forces, virial
on the same coordinates, and device time per call."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gomc_b200 import synth, engine as eng

def timeit(f, n=10):
    f(); f()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return round(min(ts) * 1e3, 4)

def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))

systems = []
for a in sys.argv[1:] or ["spce1000", "argon4000", "spce10000", "spce33334"]:
    systems.append(a)
for name in systems:
    if name.startswith("spce"):
        s = synth.make_spce(int(name[4:]))
    elif name.startswith("argon"):
        s = synth.make_argon(int(name[5:]))
    elif name == "electrolyte":
        s = synth.make_electrolyte()
    else:
        raise SystemExit(name)
    e = eng.Engine.from_system(s)
    out = {"system": name, "atoms": int(s.n_atoms)}
    res = {}
    for algo in (0, 1):
        e.set_pair_algo(algo)
        en = e.box_inter(0)
        fe = e.box_force(0)
        f = [np.array(c) for c in e.get_forces(eng.ATOM_FORCE)]
        mf = [np.array(c) for c in e.get_forces(eng.MOL_FORCE)]
        e.set_com(*s.com())
        vir = e.box_inter_virial(0) if hasattr(e, "box_inter_virial") else None
        res[algo] = (en, fe, f, mf, vir)
        out[f"t_inter_{algo}"] = timeit(lambda: e.box_inter(0))
        out[f"t_force_{algo}"] = timeit(lambda: e.box_force(0))
    a, b = res[0], res[1]
    out["inter_rel"] = [abs(x - y) / max(abs(x), 1e-300) for x, y in zip(a[0][:2], b[0][:2])]
    out["force_en_rel"] = [abs(x - y) / max(abs(x), 1e-300) for x, y in zip(a[1][:2], b[1][:2])]
    out["force_rel"] = max(rel(x, y) for x, y in zip(b[2], a[2]))
    out["molforce_rel"] = max(rel(x, y) for x, y in zip(b[3], a[3]))
    if a[4] is not None:
        out["virial_rel"] = rel(np.array(b[4], dtype=float).ravel(), np.array(a[4], dtype=float).ravel())
    out["energies"] = list(b[0][:2])
    print(json.dumps(out), flush=True)
    e.close()
