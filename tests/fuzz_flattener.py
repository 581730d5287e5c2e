"""Offline fuzz of the host-side flattener on the CPU model (not collected by pytest; run it by hand):

    python tests/fuzz_flattener.py [first_seed last_seed]        # GSDF_RXY=1 / GSDF_NO_GUARDS=1 select the variants,
                                                                 # FUZZ_RICH=1 adds threads, nuts, line sets, bounds wrappers,
                                                                 # FUZZ_HOST_INTERP=1 runs interp.cuh itself (tests/hostinterp.py)
                                                                 # instead of the numpy model (about 10x faster)

For every seed it builds random 3-D and 2-D trees (tests/shapes.py::random_trees, depth 3 and 5), flattens them, runs the
program on tests/progsim.py with small tiles (so that guards fire often) and compares with the oracle's evaluation of the
tree bit for bit. This is how the box-guard defect of DESIGN.md section 2 was found. ~40 ms per tree."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import hostinterp  # noqa: E402
import progsim  # noqa: E402
import shapes  # noqa: E402
from gsdf_b200 import gsdf  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    s0, s1 = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (100, 130)
    O.build()
    M = progsim.Math(O)
    bad = n = 0
    t0 = time.time()
    for dim in (3, 2):
        for seed in range(s0, s1):
            bld = gsdf.Builder()
            for depth in (3, 5):
                for name, s in shapes.random_trees(bld, seed, 12, dim, depth=depth, rich=os.environ.get('FUZZ_RICH') is not None):
                    f = bld.flatten(s)
                    P = progsim.Program(f["blob"], f["aux"])
                    if not P.supported():
                        continue
                    pos = shapes.sample_points(s)
                    if len(pos) > int(os.environ.get("FUZZ_MAXPTS", "20000")):
                        pos = pos[::len(pos) // int(os.environ.get("FUZZ_MAXPTS", "20000")) + 1]
                    t = O.Tree.from_shader(s)
                    want = t.eval2(pos) if s.is2d else t.eval3(pos)
                    n += 1
                    try:
                        if os.environ.get("FUZZ_HOST_INTERP"):   # the device interpreter's own source, compiled for the host
                            got = hostinterp.run(f, pos, "GSDF_RXY" if os.environ.get("GSDF_RXY", "0") not in ("", "0") else None)
                        else:
                            got = progsim.run(P, pos, M, tile=int(os.environ.get("FUZZ_TILE", "256")))
                    except AssertionError as e:
                        print("ASSERT", dim, seed, depth, name, e)
                        bad += 1
                        continue
                    nb = int(((got.view(np.uint32) != want.view(np.uint32)) & ~(np.isnan(got) & np.isnan(want))).sum())
                    if nb:
                        bad += 1
                        print("MISMATCH dim=%d seed=%d depth=%d %s: %d of %d points, %d instructions" % (dim, seed, depth, name, nb, len(pos), f["ninstr"]))
    print("trees %d, bad %d, %.0f s" % (n, bad, time.time() - t0))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
