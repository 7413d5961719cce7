"""Non-uniform FFT structure factor / reciprocal force vs the FP64-MMA kernels: agreement and
device time per system size (development check, not the judged bench)."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gomc_b200 import synth, engine as eng


def timeit(f, n=5):
    f()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return round(min(ts) * 1e3, 4)


for nm in [int(a) for a in sys.argv[1:]] or [1000, 10000, 33334]:
    s = synth.make_spce(nm)
    e = eng.Engine.from_system(s)
    out = {"mols": nm, "atoms": s.n_atoms, "nk": e.nk}
    res = {}
    for algo in (2, 5):
        e.set_recip_algo(algo)
        e.mark_coords_changed()
        en = e.box_reciprocal_sums(0)
        res[algo] = (en,) + tuple(e.get_recip_sums(0, eng.SUM_NEW, e.nk))
        out[f"recip_ms_algo{algo}"] = timeit(lambda: (e.mark_coords_changed(), e.box_reciprocal_sums(0)))
        e.copy_recip(0)
        e.box_force_reciprocal(0)
        res[(algo, "f")] = e.get_forces(eng.ATOM_FORCE_REC)
        out[f"force_ms_algo{algo}"] = timeit(
            lambda: (e.box_force_reciprocal(0), e.get_forces(eng.MOL_FORCE_REC, 0, 1)), 3)
    scale = max(np.max(np.abs(res[2][1])), np.max(np.abs(res[2][2])))
    out["energy_rel"] = abs(res[5][0] - res[2][0]) / abs(res[2][0])
    out["sum_err_over_max"] = max(np.max(np.abs(res[5][1] - res[2][1])),
                                  np.max(np.abs(res[5][2] - res[2][2]))) / scale
    fs = max(np.max(np.abs(c)) for c in res[(2, "f")])
    out["force_err_over_max"] = max(np.max(np.abs(a - b)) for a, b in
                                    zip(res[(5, "f")], res[(2, "f")])) / fs
    e.set_recip_algo(4)
    out["full_ms"] = timeit(lambda: e.call_full_box_energy(0, s.x, s.y, s.z))
    out["box_inter_ms"] = timeit(lambda: e.box_inter(0))
    print(json.dumps(out), flush=True)
    e.close()
