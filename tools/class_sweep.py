#!/usr/bin/env python
"""Measurement: score-kernel throughput of every kernel class (one core length per line).

    python tools/class_sweep.py [M ...]            # default: a representative length per class
    python tools/class_sweep.py --table            # every row of the class table (dcp_classes.h), forced, at capacity

For each core length M it builds a few identical-length profiles, scans the same reads (scores only)
and prints cells/s and padded-node-rows/s -- the table behind dcpgpu's shard cost model
(dcp_cost.c) and DESIGN.md's per-class numbers.  Output: one JSON object per line.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

import __graft_entry__ as ge
from common import plan7_profile_inputs
from concurrent.futures import ThreadPoolExecutor

DEFAULT = [20, 32, 50, 64, 80, 96, 110, 128, 144, 160, 176, 192, 200, 224, 240, 256, 300, 320, 350, 384, 420, 448, 480,
           512, 544, 576, 620, 672, 720, 768, 900, 1024, 1150, 1280, 1400, 1536, 1700, 1792, 1900, 2048, 2300, 2560,
           2800, 3072, 3300, 3584, 3800, 4096]


def table_rows():
    import re
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "deciphon-old_b200", "csrc",
                            "dcp_classes.h")).read()
    body = src[src.index("#define DCP_CLASS_TABLE"):]
    rows = [tuple(int(x) for x in m.groups()) for m in re.finditer(r"X\((\d+), (\d+), (\d+), (\d+)\)", body)]
    if "--unmeasured" in sys.argv:  # only the candidates whose rate is still 0
        rows = [r for r in rows if r[3] == 0]
    return [r[:3] for r in rows]


def main():
    forced = []
    if sys.argv[1:2] == ["--table"]:
        forced = table_rows()
        sizes = [(tw * 32 if tw else 16) * q for tw, q, bps in forced]
    else:
        sizes = [int(x) for x in sys.argv[1:]] or DEFAULT
    L = int(os.environ.get("L", 1500))
    target_cells = float(os.environ.get("CELLS", 1.5e11))  # ~0.3 s per point at 500 GCUPS
    reps = int(os.environ.get("REPS", 2))
    pkg = ge.load_pkg()
    cfg = pkg.protein_cfg(pkg.ENTRY_DIST_OCCUPANCY, 0.01)
    rng = np.random.default_rng(5)
    for idx, M in enumerate(sizes):
        if forced:
            os.environ["DCPGPU_FORCE_SHAPE"] = "%d,%d,%d" % forced[idx]
        w, q, b = pkg.kernel_shape(M)
        nprof = max(4, min(64, 16384 // M))
        nreads = int(max(64, target_cells / (nprof * M * L)))
        nreads = (nreads + 3) // 4 * 4
        models = [plan7_profile_inputs(rng, M) for _ in range(nprof)]
        with ThreadPoolExecutor(16) as ex:
            profs = list(ex.map(lambda m: pkg.ProteinProfile.build(*m, cfg, "S"), models))
        db = pkg.Db(0)
        for p in profs:
            db.add(p)
        db.commit()
        reads = ["".join("ACGT"[i] for i in rng.integers(0, 4, L)) for _ in range(nreads)]
        st = db.stage(reads)
        best = None
        for _ in range(reps + 1):
            r = db.scan_resident(st, want_paths=False)
            t = r.timing
            if best is None or t.score_ms < best:
                best = t.score_ms
            cells = t.alt_cells
            del r
        padded = pkg.kernel_padded_width(M)
        print(json.dumps({"M": M, "warps": forced[idx][0] if forced else w, "q": q, "blocks": b, "bps": forced[idx][2] if forced else None, "padded": padded, "nprof": nprof, "nreads": nreads,
                          "L": L, "score_ms": round(best, 3), "gcups": round(cells / best / 1e6, 1),
                          "padded_gnodes_per_s": round(cells / M * padded / best / 1e6, 1)}), flush=True)
        del st, db


if __name__ == "__main__":
    main()
