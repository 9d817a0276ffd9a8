#!/usr/bin/env python
"""Small scan through every kernel family (one warp, 2 / 3 / 4 / 6 / 8 warps, two-block groups; score and trace),
checked against the oracle -- sized to run under compute-sanitizer:

    compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_probe.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import __graft_entry__ as ge
import orc
from common import oracle_twin, plan7_profile_inputs, random_seq, ref_paths


def main():
    pkg = ge.load_pkg()
    o = orc.Oracle(double=False)
    cfg = pkg.protein_cfg(pkg.ENTRY_DIST_OCCUPANCY, 0.01)
    rng = np.random.default_rng(17)
    sizes = [int(x) for x in sys.argv[1:]] or [20, 100, 150, 200, 300, 400, 520, 600, 700, 1100, 1300, 2100, 3100]
    db = pkg.Db(0)
    twins = []
    for i, M in enumerate(sizes):
        p = pkg.ProteinProfile.build(*plan7_profile_inputs(rng, M), cfg, "SAN%02d" % i)
        db.add(p)
        twins.append(oracle_twin(o, p, 0.01))
    db.commit()
    seqs = [random_seq(rng, n) for n in (7, 41, 64, 97)]
    res = db.scan(seqs, lrt_threshold=-1e30)  # every pair is traced
    ref = o.scan(twins, seqs, thr=-1e30, flavour=1)
    assert np.array_equal(res.alt_loglik, ref["alt"]) and np.array_equal(res.null_loglik, ref["null"])
    want = ref_paths(ref, len(sizes))
    for i in range(res.nhits):
        s, p, path = res.hit_at(i)
        assert path == want[(s, p)], (s, p)
    print("sanitize probe ok: %d pairs, shapes %s" % (res.nhits, sorted({pkg.kernel_shape(m) for m in sizes})))


if __name__ == "__main__":
    main()
