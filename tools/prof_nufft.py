"""ncu driver: a few structure-factor / reciprocal-force evaluations with the non-uniform FFT."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gomc_b200 import synth, engine as eng
nm = int(sys.argv[1]) if len(sys.argv) > 1 else 33334
s = synth.make_spce(nm)
e = eng.Engine.from_system(s)
for _ in range(2):
    e.mark_coords_changed()
    e.box_reciprocal_sums(0)
    e.copy_recip(0)
    e.box_force_reciprocal(0)
e.close()
