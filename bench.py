#!/usr/bin/env python
"""bench.py -- Viterbi GCUPS of the scan hot path (BASELINE.json metric), one process per GPU.

Workload (BASELINE.json configs[1], the largest single-GPU configuration):
  1 000 synthetic Pfam-shaped profiles of core length 200  x  10 000 synthetic 1 kbp frameshifted
  coding reads, multi_hits on, LRT threshold 10, traceback + product paths for the hits.
A "step" is one pass of the hot path (null + alt Viterbi, LRT filter, traceback of hits) over the
whole batch.  `value` times the pass with sequences and profiles already resident in HBM
(dcpgpu_scan_resident); `e2e` times the reference-facing call dcpgpu_scan with HOST buffers
(sequence H2D, hit/path D2H inside the timed region).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--profiles P] [--reads R]

With N > 1 (torchrun, one rank per GPU) profiles are sharded across ranks by cumulative core
length (dcpgpu_shard_profiles), every rank scans all reads against its shard, no collective on
the data path; the per-GPU work is kept fixed (P profiles per rank) => "scaling": "weak".
--impl reference times the CPU oracle (the reference itself cannot be built offline)
on the host cores for the same metric, on a bounded sample of the same workload.
"""
import argparse
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
OPS_PER_CELL = 33  # 18 FADD + 15 two-input max
INSTR_PER_CELL = 27  # 18 FADD + 9 FMNMX3


WORKLOAD = "config2"


def gen_models(n, M, seed):
    from common import plan7_profile_inputs
    rng = np.random.default_rng(seed)
    if WORKLOAD in ("pfam", "short"):  # clipped log-normal, median ~130
        sizes = np.clip(np.exp(rng.normal(np.log(130), 0.75, n)), 50, 2000).astype(int)
    elif WORKLOAD == "long":
        sizes = rng.integers(2800, 3200, n)
    else:
        sizes = [M] * n
    return [plan7_profile_inputs(rng, int(m)) for m in sizes]


def gen_reads(models, nreads, L, seed):
    """Frameshifted coding reads: a codon path drawn from a random profile, indels 2 %, substitutions 1 %."""
    from common import AMINO, CODONS_OF
    rng = np.random.default_rng(seed)
    codon_tab = [np.array([[ "ACGT".index(c) for c in cod] for cod in CODONS_OF[a]], np.uint8) for a in AMINO]
    ncod = np.array([len(t) for t in codon_tab])
    out = []
    for r in range(nreads):
        ma = models[rng.integers(0, len(models))][1]
        M = ma.shape[0]
        p = np.exp(ma)
        p /= p.sum(1, keepdims=True)
        aa = (p.cumsum(1) > rng.random((M, 1))).argmax(1)
        pick = (rng.random(M) * ncod[aa]).astype(int)
        core = np.concatenate([codon_tab[a][k] for a, k in zip(aa, pick)])
        u = rng.random(core.size)
        keep = u >= 0.01                       # deletions
        ins = (u >= 0.01) & (u < 0.02)         # insertions after the base
        sub = (u >= 0.02) & (u < 0.03)
        core = np.where(sub, rng.integers(0, 4, core.size), core)
        rep = keep.astype(int) + ins.astype(int)
        core = np.repeat(core, rep)
        if core.size >= L:
            s = rng.integers(0, core.size - L + 1)
            seq = core[s:s + L]
        else:
            pad = L - core.size
            left = rng.integers(0, pad + 1)
            seq = np.concatenate([rng.integers(0, 4, left), core, rng.integers(0, 4, pad - left)])
        out.append(np.frombuffer(b"ACGT", np.uint8)[seq.astype(np.uint8)].tobytes())
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


def cpu_scan_sample(pkg, models, reads, budget_s, generic):
    """Time the oracle (port of thread_run + imm Viterbi) on all host cores over a bounded sample."""
    import orc
    from common import oracle_twin
    o = orc.Oracle(double=False)
    cores = os.cpu_count() or 1
    o.set_threads(cores)  # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core
    nprof = min(len(models), max(cores, 8))
    used = min(nprof, cores)
    cfg = pkg.protein_cfg(pkg.ENTRY_DIST_OCCUPANCY, 0.01)
    twins = []
    for i in range(nprof):
        p = pkg.ProteinProfile.build(*models[i], cfg, "P%d" % i)
        twins.append(oracle_twin(o, p, 0.01))
    # calibrate on one read per profile, then size the sample for ~budget_s
    t0 = time.perf_counter()
    o.scan(twins, reads[:1], thr=10.0, flavour=0 if generic else 1, want_paths=True)
    dt = max(time.perf_counter() - t0, 1e-3)
    nreads = int(max(1, min(len(reads), budget_s / dt)))
    t0 = time.perf_counter()
    ref = o.scan(twins, reads[:nreads], thr=10.0, flavour=0 if generic else 1, want_paths=True)
    dt = time.perf_counter() - t0
    assert ref["rc"] == 0
    cells = sum(len(r) for r in reads[:nreads]) * sum(m[1].shape[0] for m in models[:nprof])
    return {"gcups": cells / dt / 1e9, "pairs_per_s": nreads * nprof / dt, "seconds": dt, "cores": used,
            "sample": "%d profiles (M=%d) x %d reads of %d nt, %s oracle, OpenMP static over profiles" % (
                nprof, CORE, nreads, READ_LEN, "generic-interpreter" if generic else "specialised-recurrence")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--profiles", type=int, default=1000, help="profiles per GPU")
    ap.add_argument("--reads", type=int, default=10000)
    ap.add_argument("--core", type=int, default=CORE, help="profile core length (experiments; the metric is quoted at 200)")
    ap.add_argument("--workload", default="config2", choices=["config2", "pfam", "long", "short"],
                    help="config2 (default, the headline metric) or a scaled-down shape of BASELINE configs 3/4/5: "
                         "pfam = Pfam length distribution 50..2000 x 1.5 kbp reads, long = core ~3000 x 10 kbp contigs, "
                         "short = Pfam lengths x 150 bp reads (secondary numbers for DESIGN.md, not the bench line)")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    globals()["CORE"] = a.core
    globals()["WORKLOAD"] = a.workload
    if a.workload != "config2":
        globals()["READ_LEN"] = {"pfam": 1500, "long": 10000, "short": 150}[a.workload]

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import __graft_entry__ as ge
    pkg = ge.load_pkg()
    config = {"workload": ("configs[1]" if a.workload == "config2" else "shape of " + a.workload) + ": %d synthetic profiles (core length %d) per GPU x %d synthetic %d nt frameshifted "
                          "coding reads, multi_hits, LRT>=10, traceback of hits" % (a.profiles, CORE, a.reads, READ_LEN),
              "profiles_per_gpu": a.profiles, "reads": a.reads, "core_length": CORE, "read_length": READ_LEN,
              "sharding": "profiles by cumulative core length, no collective", "l2": "emission tables %.2f GB per GPU >> 126 MB L2"
              % (a.profiles * 1364 * 256 * 4 / 1e9)}

    if a.impl == "reference":
        if rank != 0:
            return
        models = gen_models(min(max(os.cpu_count() or 8, 8), 128), CORE, 1)
        reads = gen_reads(models, 256, READ_LEN, 2)
        vals = []
        for _ in range(a.warmup + a.steps):
            r = cpu_scan_sample(pkg, models, reads, max(2.0, a.cpu_budget / max(1, a.steps)), generic=True)
            vals.append(r)
        vals = vals[a.warmup:] or vals
        g = float(np.mean([v["gcups"] for v in vals]))
        line = {"impl": "reference", "metric": "viterbi_gcups", "value": g, "unit": "GCUPS", "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": float(np.mean([v["seconds"] for v in vals]) * 1e3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "pairs_per_s": float(np.mean([v["pairs_per_s"] for v in vals])),
                "cpu_baseline": {"value": g, "unit": "GCUPS", "cores": vals[0]["cores"], "kind": "port",
                                 "sample": vals[0]["sample"]},
                "e2e": {"value": g, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)

    # every rank generates the same world*P models, shards them by cumulative core length, keeps its shard
    t_setup = time.perf_counter()
    models = gen_models(a.profiles * world, CORE, 1)
    shard = pkg.shard_profiles([m[1].shape[0] for m in models], world)
    mine = [i for i in range(len(models)) if shard[i] == rank]
    reads = gen_reads(models, a.reads, READ_LEN, 2)
    cfg = pkg.protein_cfg(pkg.ENTRY_DIST_OCCUPANCY, 0.01)
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(min(32, os.cpu_count() or 8)) as ex:
        profs = list(ex.map(lambda i: pkg.ProteinProfile.build(*models[i], cfg, "SYN%06d" % i), mine))
    db = pkg.Db(local)
    for p in profs:
        db.add(p)
    db.commit()
    del profs
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

    # ---- end to end through the C-ABI call with host buffers ----
    db.scan(reads[:64])
    barrier()
    t1 = time.perf_counter()
    h2d = d2h = 0
    e2e_steps = []
    for _ in range(a.steps):
        ts = time.perf_counter()
        r = db.scan(reads)
        h2d, d2h = r.timing.h2d_bytes, r.timing.d2h_bytes
        _ = r.nhits
        e2e_steps.append([round((time.perf_counter() - ts) * 1e3, 2), round(r.timing.total_ms, 2)])
        del r
    barrier()
    e2e_s = time.perf_counter() - t1

    red = torch.tensor([dev_s, e2e_s, float(cells), float(len(reads) * len(mine)), wall], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = red.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = red.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_s, e2e_s, wall = mx[0].item(), mx[1].item(), mx[4].item()
        tot_cells, tot_pairs = sm[2].item(), sm[3].item()
    else:
        tot_cells, tot_pairs = float(cells), float(len(reads) * len(mine))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    gcups = tot_cells * a.steps / dev_s / 1e9
    e2e_gcups = tot_cells * a.steps / e2e_s / 1e9
    k_ms = float(np.mean(score_ms))
    k_gcups = cells / (k_ms * 1e-3) / 1e9  # rank 0's score kernel alone
    peak_ops = alu["mix_ginst"] * 1e9 * OPS_PER_CELL / INSTR_PER_CELL  # lane-ops/s at the measured mix issue rate
    achieved_ops = cells * OPS_PER_CELL / (k_ms * 1e-3)
    # algorithmic HBM bytes of one score launch: tables once + row records once per profile + outputs
    # tables (M+2 emission tables, 8 transition scores per node) + row records + one score per pair
    alg_bytes = (sum((models[i][1].shape[0] + 2) * 1364 * 4 + 8 * (models[i][1].shape[0] + 1) * 4 for i in mine)
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
        if tj.get("profiles_per_gpu") == a.profiles and tj.get("reads") == a.reads and tj.get("core_length") == CORE:
            traffic = tj["dram_bytes_per_launch"]
    except Exception:
        pass
    line = {
        "metric": "viterbi_gcups", "value": gcups, "unit": "GCUPS", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": dev_s / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config,
        "pairs_per_s": tot_pairs * a.steps / dev_s, "hits_per_step_rank0": nhits,
        "e2e": {"value": e2e_gcups, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "pairs_per_s": tot_pairs * a.steps / e2e_s},
        "gpu_launches": int(launches),
        "phases_ms_rank0": {"prep": float(np.mean(prep_ms)), "score": k_ms, "trace": float(np.mean(trace_ms)),
                            "total": float(np.mean(total_ms))},
        "wall_s_timed_region": wall, "setup_s": setup_s, "e2e_steps_ms_wall_vs_device": e2e_steps,
        "clocks": clocks,
        "roofline": {"bound": "fp32-alu-issue", "kernel": "k_score<8>", "achieved": achieved_ops / 1e12, "peak": peak_ops / 1e12,
                     "unit": "TFLOP/s", "frac": achieved_ops / peak_ops, "traffic": traffic,
                     "kernel_gcups": k_gcups, "kernel_ms": k_ms,
                     "peak_source": "measured live: dcpgpu_microbench_alu 2:1 FADD:FMNMX3 mix = %.0f G lane-instr/s "
                                    "(FADD %.0f, FMNMX3 %.0f), x33/27 ops per instruction" % (
                                        alu["mix_ginst"], alu["fadd_ginst"], alu["fmnmx3_ginst"]),
                     "hbm": {"bound": "hbm", "achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650"}},
    }
    if not a.no_cpu and world >= 1:
        c = cpu_scan_sample(pkg, models[:min(max(os.cpu_count() or 8, 8), 128)], reads[:256], a.cpu_budget, generic=False)
        line["cpu_baseline"] = {"value": c["gcups"], "unit": "GCUPS", "cores": c["cores"], "kind": "port",
                                "sample": c["sample"], "pairs_per_s": c["pairs_per_s"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
