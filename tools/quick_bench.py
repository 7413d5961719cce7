"""Quick device timing of the individual entry points (not the judged bench)."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gomc_b200 import synth, engine as eng

def timeit(f, n=5):
    f()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3, float(np.median(ts)) * 1e3

for nm in [int(a) for a in sys.argv[1:]] or [10000, 33334]:
    t0 = time.perf_counter()
    s = synth.make_spce(nm)
    t1 = time.perf_counter()
    e = eng.Engine.from_system(s)
    t2 = time.perf_counter()
    print(f"spce{nm}: atoms {s.n_atoms} L {s.axis[0]} nk {e.nk} gen {t1-t0:.1f}s setup {t2-t1:.2f}s", flush=True)
    e.enable_timing(True)
    r = {}
    r["box_inter"] = timeit(lambda: e.box_inter(0)); 
    r["box_force"] = timeit(lambda: e.box_force(0))
    r["recip_mma"] = timeit(lambda: e.box_reciprocal_sums(0)); r["recip_mma_dev"] = e.last_timing()
    e.set_recip_algo(1); r["recip_fact"] = timeit(lambda: e.box_reciprocal_sums(0)); r["recip_fact_dev"] = e.last_timing(); e.set_recip_algo(2)
    if nm <= 10000:
        e.set_recip_algo(0)
        r["recip_direct"] = timeit(lambda: e.box_reciprocal_sums(0), 2); r["recip_direct_dev"] = e.last_timing()
        e.set_recip_algo(2)
    r["full_resident"] = timeit(lambda: e.call_full_box_energy(0)); r["full_resident_dev"] = e.last_timing()
    r["full_host"] = timeit(lambda: e.call_full_box_energy(0, s.x, s.y, s.z))
    m = nm // 2
    sl = slice(s.mol_start[m], s.mol_start[m+1])
    nx, ny, nz = s.x[sl] + 0.3, s.y[sl] + 0.2, s.z[sl] - 0.1
    r["molecule_inter"] = timeit(lambda: e.molecule_inter(0, m, nx, ny, nz), 20)
    r["mol_reciprocal"] = timeit(lambda: e.mol_reciprocal(0, m, nx, ny, nz), 20)
    r["swap_trial"] = timeit(lambda: e.swap_trial(0, m, nx, ny, nz, 1), 20)
    r["molecule_trial"] = timeit(lambda: e.molecule_trial(0, m, nx, ny, nz), 20)
    e.copy_recip(0)
    r["force_recip_mma"] = timeit(lambda: (e.box_force_reciprocal(0), e.get_forces(eng.MOL_FORCE_REC, 0, 1)), 3)
    if nm <= 10000:
        e.set_recip_algo(0)
        r["force_recip_direct"] = timeit(lambda: (e.box_force_reciprocal(0), e.get_forces(eng.MOL_FORCE_REC, 0, 1)), 2)
        e.set_recip_algo(2)
    print(json.dumps({k: [round(x, 4) for x in v] for k, v in r.items()}), flush=True)
    print("energies", e.call_full_box_energy(0), flush=True)
    e.close()
