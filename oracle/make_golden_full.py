"""TEST INFRASTRUCTURE: full-size parity pins for the BASELINE configs.

Runs the UNMODIFIED reference (oracle/_ref/gomc_probe_NVT `slab` mode, built by
oracle/ref_build.mk from /root/reference) on the synthetic boxes of BASELINE.json at their
FULL sizes and commits only what a test needs (a few KB per box):
  * BoxInter LJ / real-space energy of the whole box,
  * sumRnew / sumInew of Ewald::BoxReciprocalSums on ~256 k-vectors spread over the whole
    k list (every x slab and column range, both ends, random picks),
  * a checksum of the coordinates as the reference parsed them -- the generator
    (gomc_b200.synth) is deterministic, so tests rebuild the box and compare the checksum.

Run in the build container only (needs /root/reference):
    python oracle/make_golden_full.py [name ...]
"""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gomc_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

# BASELINE.json configs[0], [1], [3], [4] (configs[2], the pentane GEMC pair, has its own
# fixture: tests/golden/pentane_gemc.npz)
CASES = {
    "full_cfg1_argon4000": lambda: synth.make_argon(4000),
    "full_cfg2_spce10k": lambda: synth.make_spce(10000),
    "full_cfg4_spce100k": lambda: synth.make_spce(33334),
    "full_cfg5_electrolyte1m": lambda: synth.make_electrolyte(),
}


def checksum(s):
    i = np.arange(s.n_atoms)
    return np.array([s.x.sum(), s.y.sum(), s.z.sum(), (s.x * (i % 97 + 1)).sum(),
                     (s.y * (i % 89 + 1)).sum(), (s.z * (i % 83 + 1)).sum()])


def main():
    probe = os.path.join(ROOT, "oracle", "_ref", "gomc_probe_NVT")
    out_dir = os.path.join(ROOT, "tests", "golden")
    env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1))
    only = set(sys.argv[1:])
    for name, make in CASES.items():
        if only and name not in only:
            continue
        t0 = time.time()
        s = make()
        with tempfile.TemporaryDirectory() as d:
            synth.write_gomc_inputs(s, d)
            log = subprocess.run([probe, "slab", "in.conf", "dump.bin", "256", "11", "coords"],
                                 cwd=d, env=env, capture_output=True, text=True)
            if log.returncode != 0:
                print(log.stdout[-3000:], log.stderr[-2000:])
                raise SystemExit(f"probe failed on {name}")
            dump = po.read_dump(os.path.join(d, "dump.bin"))
        # the reference parsed exactly the generator's coordinates (3-decimal PDB fields)
        for c, a in zip("xyz", (s.x, s.y, s.z)):
            assert np.array_equal(dump.pop("coords." + c), a), f"{name}: coords.{c} differ"
        assert int(dump["nAtoms"][0]) == s.n_atoms
        # order-independent comparison: the probe sums serially, numpy pairwise
        assert np.allclose(dump["coords.checksum"], checksum(s), rtol=1e-12)
        dump["coords.checksum"] = checksum(s)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **dump)
        print(f"{name}: {s.n_atoms} atoms, nk={int(dump['box0.nk'][0]) if 'box0.nk' in dump else 0}, "
              f"{len(dump.get('slab.index', []))} pinned k-vectors, {time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    main()
