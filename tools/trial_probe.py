"""Launch a few single-molecule and swap trials on the spce10k box (ncu target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gomc_b200 import engine as eng, synth

s = synth.make_spce(10000, seed=123)
e = eng.Engine.from_system(s)
e.call_full_box_energy(0)
e.set_recip_ref(0)
rng = np.random.default_rng(1)
for t in range(4):
    m = int(rng.integers(s.n_mols))
    sl = slice(s.mol_start[m], s.mol_start[m + 1])
    d = rng.uniform(-0.5, 0.5, 3)
    nx, ny, nz = s.x[sl] + d[0], s.y[sl] + d[1], s.z[sl] + d[2]
    print(e.molecule_trial(0, m, nx, ny, nz))
    print(e.swap_trial(0, m, nx, ny, nz, 1))
e.close()
