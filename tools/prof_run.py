"""Small driver for ncu captures: one system, a few calls of the hot entry points."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gomc_b200 import synth, engine as eng
nm = int(sys.argv[1]) if len(sys.argv) > 1 else 33334
what = sys.argv[2] if len(sys.argv) > 2 else "full"
s = synth.make_spce(nm)
e = eng.Engine.from_system(s)
for _ in range(3):
    if what in ("full", "inter"): e.box_inter(0)
    if what in ("full", "force"): e.box_force(0)
    if what in ("full", "recip"): e.box_reciprocal_sums(0)
e.close()
