#!/usr/bin/env python
"""Experiment: L2 tile size of the score kernel's work order (DCPGPU_TILE_MB) vs kernel time.
Run plain for times, or under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:k_score`
for DRAM bytes per launch (one launch per tile size, in the order given)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import __graft_entry__ as ge
from concurrent.futures import ThreadPoolExecutor

tiles = [float(x) for x in (sys.argv[1:] or ["64", "32", "16", "8"])]
P, R = int(os.environ.get("P", 1000)), int(os.environ.get("R", 10000))
pkg = ge.load_pkg()
models = bench.gen_models(P, bench.CORE, 1)
reads = bench.gen_reads(models, R, bench.READ_LEN, 2)
cfg = pkg.protein_cfg(pkg.ENTRY_DIST_OCCUPANCY, 0.01)
with ThreadPoolExecutor(16) as ex:
    profs = list(ex.map(lambda i: pkg.ProteinProfile.build(*models[i], cfg, "SYN%06d" % i), range(P)))
db = pkg.Db(0)
for p in profs:
    db.add(p)
db.commit()
staged = db.stage(reads)
reps = int(os.environ.get("REPS", 1))
for t in tiles:
    os.environ["DCPGPU_TILE_MB"] = str(t)
    for _ in range(reps):
        r = db.scan_resident(staged)
        tm = r.timing
        print("tile_mb=%g score_ms=%.1f gcups=%.1f" % (t, tm.score_ms, tm.alt_cells / tm.score_ms / 1e6), flush=True)
        del r
