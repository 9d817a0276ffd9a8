#!/usr/bin/env python
"""bench.py -- Viterbi GCUPS of the scan hot path (BASELINE.json metric), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N = 1 (the headline, BASELINE.json configs[1], the largest configuration quoted on one GPU):
  1 000 synthetic Pfam-shaped profiles of core length 200  x  10 000 synthetic 1 kbp frameshifted coding reads,
  multi_hits on, LRT threshold 10, traceback + product paths for the hits.  A "step" is one pass of the hot path
  (null + alt Viterbi, LRT filter, traceback of hits) over the whole batch.  `value` times the pass with sequences
  and profiles already resident in HBM (dcpgpu_scan_resident); `e2e` times the reference-facing call dcpgpu_scan
  with HOST buffers (sequence H2D, hit/path D2H inside the timed region).  The line also carries a `secondary`
  block: the shapes of configs[2..4] run once each in the same invocation, with hit checksums and oracle samples.

N > 1 (torchrun, one rank per GPU; BASELINE.json configs[2], "profiles sharded across 1/2/4/8 B200"):
  ONE FIXED Pfam-A-sized database (20 000 profiles, clipped log-normal lengths 50..2000) x 1 000 reads of 1.5 kbp.
  The database is sharded over the ranks by modelled cost (dcpgpu_shard_profiles), every rank scans all reads
  against its shard, rank 0 gathers the hit records (NCCL) and merges them by (sequence, profile).  Total work is
  fixed => "scaling": "strong".  There is no collective on the data path; what limits scaling is shard balance,
  reported per rank (`busy_ms`).  The same database on one GPU is the N = 1 line's `secondary.config3_fixed_db`
  (same hit checksum), so strong-scaling efficiency can be read from driver-run lines alone.

--impl reference times the CPU oracle (the reference itself cannot be built offline) on the host
cores for the same metric, on a bounded sample of the same workload; it never loads libdcpgpu.so.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CORE = 200
READ_LEN = 1000
OPS_PER_CELL = 33    # 18 FADD + 15 two-input max
INSTR_PER_CELL = 27  # 18 FADD + 9 FMNMX3
DB_PROFILES = 20000  # the fixed config-3-shaped database of the N > 1 runs
DB_READS = 1000
DB_READ_LEN = 1500


# ----------------------------------------------------------------------------------------------------------------
# synthetic workloads (deterministic, per-profile seeds so that a rank can build only its shard)
# ----------------------------------------------------------------------------------------------------------------
def pfam_sizes(n, seed):
    """Clipped log-normal core lengths on [50, 2000], median ~130, mean ~175."""
    rng = np.random.default_rng(seed)
    return np.clip(np.exp(rng.normal(np.log(130), 0.75, n)), 50, 2000).astype(int)


def model_of(seed, i, M):
    from common import plan7_profile_inputs
    return plan7_profile_inputs(np.random.default_rng([seed, i]), int(M))


def gen_models(n, M, seed, workload="config2"):
    rng = np.random.default_rng(seed)
    if workload in ("pfam", "short"):
        sizes = pfam_sizes(n, seed)
    elif workload == "long":
        sizes = rng.integers(2800, 3200, n)
    else:
        sizes = [M] * n
    from common import plan7_profile_inputs
    return [plan7_profile_inputs(rng, int(m)) for m in sizes]


_CODON_TAB = None


def read_from(rng, match_lp, L, indel=0.02, sub=0.01):
    """A frameshifted coding read: a codon path drawn from the profile, per-base indels and substitutions,
    cut or padded with random nucleotides to L."""
    global _CODON_TAB
    from common import AMINO, CODONS_OF
    if _CODON_TAB is None:
        tab = [np.array([["ACGT".index(c) for c in cod] for cod in CODONS_OF[a]], np.uint8) for a in AMINO]
        _CODON_TAB = (tab, np.array([len(t) for t in tab]))
    codon_tab, ncod = _CODON_TAB
    M = match_lp.shape[0]
    p = np.exp(match_lp)
    p /= p.sum(1, keepdims=True)
    aa = (p.cumsum(1) > rng.random((M, 1))).argmax(1)
    pick = (rng.random(M) * ncod[aa]).astype(int)
    core = np.concatenate([codon_tab[a][k] for a, k in zip(aa, pick)])
    u = rng.random(core.size)
    keep = u >= indel / 2                       # deletions
    ins = (u >= indel / 2) & (u < indel)        # insertions after the base
    subm = (u >= indel) & (u < indel + sub)
    core = np.where(subm, rng.integers(0, 4, core.size), core)
    core = np.repeat(core, keep.astype(int) + ins.astype(int))
    if core.size >= L:
        s = rng.integers(0, core.size - L + 1)
        seq = core[s:s + L]
    else:
        pad = L - core.size
        left = rng.integers(0, pad + 1)
        seq = np.concatenate([rng.integers(0, 4, left), core, rng.integers(0, 4, pad - left)])
    return np.frombuffer(b"ACGT", np.uint8)[seq.astype(np.uint8)].tobytes()


def gen_reads(models, nreads, L, seed, indel=0.02, sub=0.01):
    rng = np.random.default_rng(seed)
    return [read_from(rng, models[rng.integers(0, len(models))][1], L, indel, sub) for _ in range(nreads)]


def gen_reads_db(sizes, model_seed, nreads, L, seed, indel=0.02, sub=0.01):
    """Reads drawn from random profiles of a database given by (sizes, model_seed); only those models are built."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(nreads):
        i = int(rng.integers(0, len(sizes)))
        out.append(read_from(rng, model_of(model_seed, i, sizes[i])[1], L, indel, sub))
    return out


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_ev = index, [], threading.Event()

    def run(self):
        while not self.stop_ev.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in o.strip().split(",")])
            except Exception:
                pass
            self.stop_ev.wait(0.2)

    def summary(self):
        self.stop_ev.set()
        self.join()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of thread_run + imm Viterbi, timed on the host cores (never touches libdcpgpu.so)
# ----------------------------------------------------------------------------------------------------------------
def cpu_scan_sample(models, reads, budget_s, generic, what):
    """Time the oracle on all host cores over a bounded sample: `models` are (null_lp, match_lp, trans) inputs,
    built with the oracle's own model builder (orc_profile_build)."""
    import orc
    o = orc.Oracle(double=False)
    cores = os.cpu_count() or 1
    o.set_threads(cores)  # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core
    twins = [o.build(m[1].shape[0], orc.ENTRY_OCCUPANCY, 0.01, m[0], m[1], m[2]) for m in models]
    used = min(len(twins), cores)
    flav = 0 if generic else 1
    # calibrate on one read, then size the sample for ~budget_s
    t0 = time.perf_counter()
    o.scan(twins, reads[:1], thr=10.0, flavour=flav, want_paths=True)
    dt = max(time.perf_counter() - t0, 1e-3)
    nreads = int(max(1, min(len(reads), budget_s / dt)))
    t0 = time.perf_counter()
    ref = o.scan(twins, reads[:nreads], thr=10.0, flavour=flav, want_paths=True)
    dt = time.perf_counter() - t0
    assert ref["rc"] == 0
    cells = sum(len(r) for r in reads[:nreads]) * sum(m[1].shape[0] for m in models)
    return {"gcups": cells / dt / 1e9, "pairs_per_s": nreads * len(twins) / dt, "seconds": dt, "cores": used,
            "sample": "%d profiles (%s) x %d reads of %d nt, %s oracle, OpenMP static over profiles" % (
                len(twins), what, nreads, len(reads[0]), "generic-interpreter (imm-shaped)" if generic else
                "specialised-recurrence")}


def cpu_workload(strong, ncores, nprofiles=1000):
    """(models, reads, description) of the CPU sample: the FIRST profiles and the FIRST reads of the very workload
    the GPU arm scans (config 2 for N = 1, the fixed config-3-shaped database for N > 1), so the sample has the
    workload's own hit rate (reads are drawn from the whole database, not from the sampled profiles)."""
    n = min(max(ncores, 8), 128)
    if strong:
        sizes = pfam_sizes(DB_PROFILES, 1)
        models = [model_of(1, i, sizes[i]) for i in range(n)]
        reads = gen_reads_db(sizes, 1, 128, DB_READ_LEN, 2)
        return models, reads, "Pfam lengths 50..2000, mean %d" % int(np.mean(sizes[:n]))
    models = gen_models(nprofiles, CORE, 1)
    return models[:n], gen_reads(models, 256, READ_LEN, 2), "M=%d" % CORE


# ----------------------------------------------------------------------------------------------------------------
# helpers of the GPU arm
# ----------------------------------------------------------------------------------------------------------------
def build_db(pkg, device, items, seed_names="SYN"):
    """items: [(global index, (null_lp, match_lp, trans))] -> committed Db.  Profiles are built on all host cores
    in chunks, so the host never holds more than one chunk of 5.4 KB-per-node tables besides the database's own."""
    from concurrent.futures import ThreadPoolExecutor
    cfg = pkg.protein_cfg(pkg.ENTRY_DIST_OCCUPANCY, 0.01)
    db = pkg.Db(device)
    nthr = min(32, os.cpu_count() or 8)
    with ThreadPoolExecutor(nthr) as ex:
        for a in range(0, len(items), 256):
            chunk = items[a:a + 256]
            profs = list(ex.map(lambda it: pkg.ProteinProfile.build(*it[1](), cfg, "%s%06d" % (seed_names, it[0])), chunk))
            for p in profs:
                db.add(p)
            db.profiles = []  # the database holds its own copies
            del profs
    db.commit()
    return db


def hits_digest(seq, gprof, alt, null, nsteps, steps):
    """Order-sensitive digest of a merged hit list: identical across device counts iff the merged lists are."""
    h = hashlib.sha256()
    for a in (seq.astype(np.uint32), gprof.astype(np.uint32), alt.astype(np.float32), null.astype(np.float32),
              nsteps.astype(np.uint32), steps.astype(np.uint16)):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


def result_hit_arrays(res, mine):
    """(seq, global profile, alt, null, nsteps, steps) of one rank's result; `mine`: local -> global profile."""
    seq, prof, alt, null, ns = res.hits()
    return seq, np.asarray(mine, np.uint32)[prof], alt, null, ns, res.steps()


def merge_hit_arrays(parts):
    """Concatenate per-rank hit arrays and order them by (sequence, global profile), paths included."""
    seq = np.concatenate([p[0] for p in parts])
    prof = np.concatenate([p[1] for p in parts])
    alt = np.concatenate([p[2] for p in parts])
    null = np.concatenate([p[3] for p in parts])
    ns = np.concatenate([p[4] for p in parts])
    steps = np.concatenate([p[5] for p in parts])
    order = np.lexsort((prof, seq))
    starts = np.concatenate([[0], np.cumsum(ns)[:-1]]).astype(np.int64)
    if len(order):
        idx = np.concatenate([np.arange(starts[i], starts[i] + ns[i]) for i in order]) if ns.sum() else np.zeros(0, np.int64)
    else:
        idx = np.zeros(0, np.int64)
    return seq[order], prof[order], alt[order], null[order], ns[order], steps[idx]


def gather_hit_arrays(dist, torch, arrays, world, rank):
    """All ranks' hit arrays on rank 0, through NCCL: one padded uint8 all_gather per step."""
    blobs = [np.ascontiguousarray(a).view(np.uint8).reshape(-1) for a in arrays]
    sizes = torch.tensor([b.size for b in blobs], dtype=torch.int64, device="cuda")
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = [s.cpu().numpy() for s in all_sizes]
    cap = int(max(s.sum() for s in all_sizes))
    mine = torch.zeros(max(cap, 1), dtype=torch.uint8, device="cuda")
    flat = np.concatenate(blobs) if blobs else np.zeros(0, np.uint8)
    mine[:flat.size] = torch.from_numpy(flat).cuda()
    got = [torch.zeros_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, got, dst=0)
    if rank != 0:
        return None
    dts = [np.uint32, np.uint32, np.float32, np.float32, np.uint32, np.uint16]
    parts = []
    for r in range(world):
        buf = got[r].cpu().numpy()
        off, arrs = 0, []
        for k, dt in enumerate(dts):
            n = int(all_sizes[r][k])
            a = buf[off:off + n].view(dt)
            arrs.append(a.reshape(-1, 2) if k == 5 else a)
            off += n
        parts.append(tuple(arrs))
    return parts


def oracle_sample_check(pkg, res, prof_inputs, reads, pairs):
    """Checker: the GPU's alt/null log-likelihoods and hit flags of a few (sequence, profile) pairs against the
    oracle's specialised recurrence, bit for bit.  prof_inputs(p) -> (null_lp, match_lp, trans) of local profile p."""
    import orc
    from common import oracle_twin
    o = orc.Oracle(double=False)
    cfg = pkg.protein_cfg(pkg.ENTRY_DIST_OCCUPANCY, 0.01)
    alt, null, hit = res.alt_loglik, res.null_loglik, res.hit
    ok = 0
    for s, p in pairs:
        prof = pkg.ProteinProfile.build(*prof_inputs(p), cfg, "chk")
        tw = oracle_twin(o, prof, 0.01)
        r = o.scan([tw], [reads[s]], thr=10.0, flavour=1, want_paths=False)
        same = (np.float32(r["alt"][0, 0]).tobytes() == np.float32(alt[s, p]).tobytes()
                and np.float32(r["null"][0, 0]).tobytes() == np.float32(null[s, p]).tobytes()
                and int(r["hit"][0, 0]) == int(hit[s, p]))
        ok += int(same)
    return {"pairs": len(pairs), "bit_equal": ok}


def run_secondary(pkg, device, alu, peak_ops, args):
    """Shapes of BASELINE configs[2..4] on one GPU, one timed pass each (after one untimed pass): GCUPS, the score
    kernels' fraction of the measured issue peak, a digest of the merged hit list and an oracle sample check."""
    out = {}
    rng = np.random.default_rng(11)

    def one(name, db, reads, inputs_of, mine, note):
        db.scan(reads)  # untimed pass of the same size: the database's memory pool reaches its working size
        t0 = time.perf_counter()
        res = db.scan(reads)
        wall = time.perf_counter() - t0
        t = res.timing
        arr = result_hit_arrays(res, mine)
        nprof = db.nprofiles
        # sample: the first hits (true positives) and random pairs (mostly misses)
        hs, hp = res.hits()[0], res.hits()[1]
        pairs = [(int(hs[i]), int(hp[i])) for i in np.linspace(0, max(len(hs) - 1, 0), min(8, len(hs))).astype(int)]
        pairs += [(int(rng.integers(0, len(reads))), int(rng.integers(0, nprof))) for _ in range(8)]
        chk = oracle_sample_check(pkg, res, inputs_of, reads, pairs) if not args.no_cpu else None
        cells = float(t.alt_cells)
        out[name] = {"workload": note, "gcups": cells / (t.total_ms * 1e-3) / 1e9, "e2e_gcups": cells / wall / 1e9,
                     "pairs_per_s": len(reads) * nprof / (t.total_ms * 1e-3), "score_ms": t.score_ms,
                     "trace_ms": t.trace_ms, "total_ms": t.total_ms, "launches": int(t.launches),
                     "score_frac_of_measured_issue_peak": cells * OPS_PER_CELL / (t.score_ms * 1e-3) / peak_ops,
                     "hits": int(res.nhits), "hits_digest": hits_digest(*arr), "oracle_sample": chk}
        del res

    # config 3 shape: the fixed database the N > 1 runs shard, here on one GPU
    sizes = pfam_sizes(args.db_profiles, 1)
    t0 = time.perf_counter()
    db = build_db(pkg, device, [(i, (lambda i=i: model_of(1, i, sizes[i]))) for i in range(len(sizes))], "PFS")
    build_s = time.perf_counter() - t0
    reads = gen_reads_db(sizes, 1, args.db_reads, DB_READ_LEN, 2)
    one("config3_fixed_db", db, reads, lambda p: model_of(1, p, sizes[p]), np.arange(len(sizes)),
        "configs[2] shape, the database of the N > 1 runs: %d profiles (Pfam lengths 50..2000, %d nodes) x %d reads of "
        "%d nt on ONE GPU" % (len(sizes), int(sizes.sum()), len(reads), DB_READ_LEN))
    out["config3_fixed_db"]["db_build_s"] = build_s
    out["config3_fixed_db"]["db_device_gb"] = db.device_bytes / 1e9
    # config 5 shape: short Illumina-like reads against the same database, full traceback of the hits
    short = gen_reads_db(sizes, 1, 4 * args.db_reads, 150, 3, indel=0.0002, sub=0.005)
    one("config5_short_reads", db, short, lambda p: model_of(1, p, sizes[p]), np.arange(len(sizes)),
        "configs[4] shape: %d reads of 150 nt (substitutions 0.5 %%, indels 0.02 %%) x the same %d profiles, traceback "
        "of all hits" % (len(short), len(sizes)))
    del db
    # config 4 shape: long profiles x long contigs
    lrng = np.random.default_rng(4)
    lsizes = lrng.integers(2800, 3200, 8)
    ldb = build_db(pkg, device, [(i, (lambda i=i: model_of(4, i, lsizes[i]))) for i in range(len(lsizes))], "LNG")
    contigs = gen_reads_db(lsizes, 4, 200, 10000, 5)
    one("config4_long", ldb, contigs, lambda p: model_of(4, p, lsizes[p]), np.arange(len(lsizes)),
        "configs[3] shape: %d profiles of ~3000 nodes x %d contigs of 10 kbp, traceback of all hits" % (
            len(lsizes), len(contigs)))
    del ldb
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--profiles", type=int, default=1000, help="config 2: profiles")
    ap.add_argument("--reads", type=int, default=10000, help="config 2: reads")
    ap.add_argument("--core", type=int, default=CORE, help="profile core length (experiments; the metric is quoted at 200)")
    ap.add_argument("--workload", default=None, choices=["config2", "fixeddb", "pfam", "long", "short"],
                    help="default: config2 on one GPU, fixeddb (the config-3-shaped database, strong scaling) on several; "
                         "pfam / long / short: scaled shapes of configs 3 / 4 / 5 for experiments")
    ap.add_argument("--db-profiles", type=int, default=DB_PROFILES, help="fixeddb: profiles of the database")
    ap.add_argument("--db-reads", type=int, default=DB_READS, help="fixeddb: reads of 1.5 kbp")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    a = ap.parse_args()
    globals()["CORE"] = a.core

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    workload = a.workload or ("fixeddb" if max(world, a.gpus) > 1 else "config2")
    strong = workload == "fixeddb"
    if workload in ("pfam", "long", "short"):
        globals()["READ_LEN"] = {"pfam": 1500, "long": 10000, "short": 150}[workload]

    if strong:
        config = {"workload": "configs[2] shape: ONE fixed database of %d synthetic profiles (Pfam length distribution, clipped "
                              "log-normal 50..2000) x %d synthetic %d nt frameshifted coding reads, multi_hits, LRT>=10, "
                              "traceback of hits; profiles sharded over the GPUs by modelled cost, hits merged by "
                              "(sequence, profile) on rank 0" % (a.db_profiles, a.db_reads, DB_READ_LEN),
                  "db_profiles": a.db_profiles, "reads": a.db_reads, "read_length": DB_READ_LEN,
                  "sharding": "profiles by modelled cost (padded width / measured class rate), no collective on the data path",
                  "l2": "emission tables %.1f GB in total >> 126 MB L2" % (pfam_sizes(a.db_profiles, 1).sum() * 5456 / 1e9),
                  "e2e_returns": "hit rows only (thread_run emits products for hits; per-pair scores stay on the device)"}
    else:
        config = {"workload": ("configs[1]" if workload == "config2" else "shape of " + workload) +
                              ": %d synthetic profiles (core length %d) x %d synthetic %d nt frameshifted coding reads, "
                              "multi_hits, LRT>=10, traceback of hits" % (a.profiles, CORE, a.reads, READ_LEN),
                  "profiles_per_gpu": a.profiles, "reads": a.reads, "core_length": CORE, "read_length": READ_LEN,
                  "sharding": "one GPU", "l2": "emission tables %.2f GB per GPU >> 126 MB L2" % (a.profiles * 1364 * 256 * 4 / 1e9),
                  "e2e_returns": "hit rows only (thread_run emits products for hits; per-pair scores stay on the device)"}

    if a.impl == "reference":
        if rank != 0:
            return
        ncores = os.cpu_count() or 8
        models, reads, what = cpu_workload(strong, ncores, a.profiles)
        per = max(2.0, a.cpu_budget / max(1, a.steps))
        vals = [cpu_scan_sample(models, reads, per, generic=False, what=what) for _ in range(a.warmup + a.steps)]
        vals = vals[a.warmup:] or vals
        gen = cpu_scan_sample(models, reads, min(10.0, a.cpu_budget), generic=True, what=what)
        g = float(np.mean([v["gcups"] for v in vals]))
        line = {"impl": "reference", "metric": "viterbi_gcups", "value": g, "unit": "GCUPS", "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": float(np.mean([v["seconds"] for v in vals]) * 1e3),
                "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config, "pairs_per_s": float(np.mean([v["pairs_per_s"] for v in vals])),
                "cpu_baseline": {"value": g, "unit": "GCUPS", "cores": vals[0]["cores"], "kind": "port",
                                 "sample": vals[0]["sample"],
                                 "flavours": {"specialised_recurrence_gcups": g, "generic_interpreter_gcups": gen["gcups"],
                                              "note": "value = the faster flavour; the generic interpreter has imm's "
                                                      "algorithmic shape (state and transition lists, DP matrix reused "
                                                      "per thread); both hold the profiles in RAM, the reference "
                                                      "re-reads them from disk per pair"}},
                "e2e": {"value": g, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    pkg = ge.load_pkg()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)

    t_setup = time.perf_counter()
    if strong:
        sizes = pfam_sizes(a.db_profiles, 1)
        shard = pkg.shard_profiles(sizes, world)
        mine = np.nonzero(shard == rank)[0]
        db = build_db(pkg, local, [(int(i), (lambda i=i: model_of(1, int(i), sizes[i]))) for i in mine], "PFS")
        reads = gen_reads_db(sizes, 1, a.db_reads, DB_READ_LEN, 2)
        cost = np.array([pkg.profile_cost(m) for m in sizes])
        modelled = np.bincount(shard, weights=cost, minlength=world)
        my_inputs = lambda p: model_of(1, int(mine[p]), sizes[mine[p]])
    else:
        models = gen_models(a.profiles, CORE, 1, workload)
        mine = np.arange(len(models))
        db = build_db(pkg, local, [(i, (lambda i=i: models[i])) for i in range(len(models))])
        reads = gen_reads(models, a.reads, READ_LEN, 2)
        my_inputs = lambda p: models[p]
    staged = db.stage(reads)
    setup_s = time.perf_counter() - t_setup

    alu = pkg.microbench_alu(local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value) ----
    for _ in range(a.warmup):
        db.scan_resident(staged)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    score_ms, total_ms, trace_ms, prep_ms, launches, cells, nhits = [], [], [], [], 0, 0, 0
    for _ in range(a.steps):
        r = db.scan_resident(staged)
        t = r.timing
        score_ms.append(t.score_ms), total_ms.append(t.total_ms), trace_ms.append(t.trace_ms), prep_ms.append(t.prep_ms)
        launches += t.launches
        cells = t.alt_cells
        nhits = r.nhits
        del r
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.summary()
    dev_s = sum(total_ms) / 1e3  # CUDA-event time of the K passes on the engine's stream

    # ---- end to end through the C-ABI call with host buffers (N > 1: + gather of the hit records, merge) ----
    gather_bytes = 0

    def e2e_step():
        """One end-to-end pass: C-ABI call with host buffers, hit records of all ranks on rank 0, merged."""
        nonlocal gather_bytes
        r = db.scan(reads)
        arrays = result_hit_arrays(r, mine)
        out = arrays
        if world > 1:
            parts = gather_hit_arrays(dist, torch, arrays, world, rank)
            gather_bytes = sum(x.nbytes for x in arrays)
            out = merge_hit_arrays(parts) if rank == 0 else None
        return r, out

    e2e_step()  # untimed: first use of the host path and of the NCCL gather (communicator set-up)
    barrier()
    t1 = time.perf_counter()
    h2d = d2h = 0
    e2e_steps, merged = [], None
    for _ in range(a.steps):
        ts = time.perf_counter()
        r, merged = e2e_step()
        h2d, d2h = r.timing.h2d_bytes, r.timing.d2h_bytes
        e2e_steps.append([round((time.perf_counter() - ts) * 1e3, 2), round(r.timing.total_ms, 2)])
        last = r
    barrier()
    e2e_s = time.perf_counter() - t1

    red = torch.tensor([dev_s, e2e_s, float(cells), float(len(reads) * len(mine)), wall], dtype=torch.float64, device="cuda")
    busy = [dev_s / a.steps * 1e3]
    if world > 1:
        mx = red.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = red.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        allb = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(allb, red[:1].clone())
        busy = [float(b.item()) / a.steps * 1e3 for b in allb]
        dev_s, e2e_s, wall = mx[0].item(), mx[1].item(), mx[4].item()
        tot_cells, tot_pairs = sm[2].item(), sm[3].item()
    else:
        tot_cells, tot_pairs = float(cells), float(len(reads) * len(mine))
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    gcups = tot_cells * a.steps / dev_s / 1e9
    e2e_gcups = tot_cells * a.steps / e2e_s / 1e9
    k_ms = float(np.mean(score_ms))
    k_gcups = cells / (k_ms * 1e-3) / 1e9  # rank 0's score kernels alone
    peak_ops = alu["mix_ginst"] * 1e9 * OPS_PER_CELL / INSTR_PER_CELL  # lane-ops/s at the measured mix issue rate
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    hard_ops = alu["sms"] * 128 * sm_mhz * 1e6 * OPS_PER_CELL / INSTR_PER_CELL  # one lane-instruction per lane and clock
    achieved_ops = cells * OPS_PER_CELL / (k_ms * 1e-3)
    # algorithmic HBM bytes of one score pass: tables once (M+2 emission tables, 8 transition scores per node) +
    # row records once + one score per pair
    if strong:
        my_sizes = sizes[mine]
    else:
        my_sizes = np.array([m[1].shape[0] for m in models])
    alg_bytes = (float(((my_sizes + 2) * 1364 * 4 + 8 * (my_sizes + 1) * 4).sum())
                 + sum(len(x) for x in reads) * 66 + len(reads) * len(mine) * 4)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum of one k_score launch (ncu --set full), if captured
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if not strong and tj.get("profiles_per_gpu") == a.profiles and tj.get("reads") == a.reads and tj.get("core_length") == CORE:
            traffic = tj["dram_bytes_per_launch"]
    except Exception:
        pass
    line = {
        "metric": "viterbi_gcups", "value": gcups, "unit": "GCUPS", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": dev_s / a.steps * 1e3, "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
        "pairs_per_s": tot_pairs * a.steps / dev_s, "hits_per_step_rank0": nhits,
        "e2e": {"value": e2e_gcups, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "pairs_per_s": tot_pairs * a.steps / e2e_s,
                "includes": "sequence H2D, scan, hit/path D2H" + (", NCCL gather of %d B of hit records per rank, merge on rank 0"
                                                                  % gather_bytes if world > 1 else "")},
        "gpu_launches": int(launches),
        "phases_ms_rank0": {"prep": float(np.mean(prep_ms)), "score": k_ms, "trace": float(np.mean(trace_ms)),
                            "total": float(np.mean(total_ms))},
        "wall_s_timed_region": wall, "setup_s": setup_s, "e2e_steps_ms_wall_vs_device": e2e_steps,
        "clocks": clocks,
        "merged_hits": {"count": int(len(merged[0])), "digest": hits_digest(*merged)},
        "roofline": {"bound": "fp32-alu-issue", "kernel": "k_score<8>" if not strong else "k_score* (all classes of rank 0's shard)",
                     "achieved": achieved_ops / 1e12, "peak": peak_ops / 1e12,
                     "unit": "TFLOP/s", "frac": achieved_ops / peak_ops, "frac_hard": achieved_ops / hard_ops,
                     "peak_hard": hard_ops / 1e12, "traffic": traffic,
                     "kernel_gcups": k_gcups, "kernel_ms": k_ms,
                     "peak_source": "measured live: dcpgpu_microbench_alu 2:1 FADD:FMNMX3 mix = %.0f G lane-instr/s "
                                    "(FADD %.0f, FMNMX3 %.0f), x33/27 ops per instruction; peak_hard = %d SMs x 128 lanes x "
                                    "%.0f MHz x33/27 (no mix, no dual-issue limits); inside the microbenchmark the SM "
                                    "clock was %.0f MHz (mix) / %.0f MHz (FADD): %.1f / %.1f lane-instr per SM and clock "
                                    "of 128" % (
                                        alu["mix_ginst"], alu["fadd_ginst"], alu["fmnmx3_ginst"], alu["sms"], sm_mhz,
                                        alu["mix_clock_mhz"], alu["fadd_clock_mhz"],
                                        alu["mix_ginst"] * 1e3 / (alu["sms"] * max(alu["mix_clock_mhz"], 1.0)),
                                        alu["fadd_ginst"] * 1e3 / (alu["sms"] * max(alu["fadd_clock_mhz"], 1.0))),
                     "hbm": {"bound": "hbm", "achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650"}},
    }
    if strong:
        line["shards"] = {"busy_ms_per_step": [round(b, 2) for b in busy],
                          "balance": sum(busy) / (len(busy) * max(busy)),
                          "imbalance_ms": max(busy) - sum(busy) / len(busy),
                          "modelled_cost_share": [round(float(x / modelled.sum()), 4) for x in modelled],
                          "profiles_per_rank": [int((shard == r).sum()) for r in range(world)],
                          "limiter": "shard balance: value = total cells / slowest rank; one-GPU time of the same database "
                                     "~ sum of busy_ms (also measured directly: the N = 1 line's secondary.config3_fixed_db)"}
    if not a.no_cpu:
        cm, cr, what = cpu_workload(strong, os.cpu_count() or 8, a.profiles)
        c = cpu_scan_sample(cm, cr, a.cpu_budget, generic=False, what=what)
        line["cpu_baseline"] = {"value": c["gcups"], "unit": "GCUPS", "cores": c["cores"], "kind": "port",
                                "sample": c["sample"], "pairs_per_s": c["pairs_per_s"]}
        # checker: a few pairs of the timed workload against the oracle, bit for bit
        rng = np.random.default_rng(3)
        hs, hp = last.hits()[0], last.hits()[1]
        pairs = [(int(hs[i]), int(hp[i])) for i in np.linspace(0, max(len(hs) - 1, 0), min(8, len(hs))).astype(int)]
        pairs += [(int(rng.integers(0, len(reads))), int(rng.integers(0, len(mine)))) for _ in range(8)]
        line["oracle_sample"] = oracle_sample_check(pkg, last, my_inputs, reads, pairs)
    del last, staged, db
    if world == 1 and not strong and workload == "config2" and not a.no_secondary:
        line["secondary"] = run_secondary(pkg, local, alu, peak_ops, a)
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
