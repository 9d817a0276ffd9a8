"""N > 1 host logic on CPU: world size 2 over gloo.  Each rank takes its shard of the profiles
(dcpgpu_shard_profiles), scores it -- with the ORACLE standing in for the GPU, there is none here --
and rank 0 merges the hits; the result must equal the single-process scan of all profiles."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["OMP_NUM_THREADS"] = "1"
    import __graft_entry__ as ge
    import orc
    from common import random_seq, ref_paths
    pkg = ge.load_pkg()
    import importlib
    sharding = importlib.import_module("deciphon_old_b200.sharding")
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    o = orc.Oracle(double=False)
    sizes = [5, 40, 7, 33, 12, 64, 3, 20]
    profs = [o.sample(100 + i, M, orc.ENTRY_OCCUPANCY, 0.01) for i, M in enumerate(sizes)]
    rng = np.random.default_rng(0)
    seqs = [random_seq(rng, n) for n in (30, 61, 90)]
    mine = sharding.shard_indices(pkg, sizes, world, rank)
    ref = o.scan([profs[i] for i in mine], seqs, thr=-1e30, flavour=1)
    paths = ref_paths(ref, len(mine))
    hits = [(s, mine[p], float(ref["alt"][s, p]), float(ref["null"][s, p]), paths[(s, p)])
            for s in range(len(seqs)) for p in range(len(mine)) if ref["hit"][s, p]]
    gathered = [None] * world
    dist.all_gather_object(gathered, hits)
    owners = [None] * world
    dist.all_gather_object(owners, mine)
    if rank == 0:
        merged = sharding.merge_hits(gathered)
        full = o.scan(profs, seqs, thr=-1e30, flavour=1)
        fp = ref_paths(full, len(sizes))
        want = [(s, p, float(full["alt"][s, p]), float(full["null"][s, p]), fp[(s, p)])
                for s in range(len(seqs)) for p in range(len(sizes)) if full["hit"][s, p]]
        loads = [sum(pkg.profile_cost(sizes[i]) for i in own) for own in owners]
        q.put((merged == want, sorted(sum(owners, [])) == list(range(len(sizes))), loads))
    dist.barrier()
    dist.destroy_process_group()


def test_axis_choice(pkg):
    import importlib
    sharding = importlib.import_module("deciphon_old_b200.sharding")
    rng = np.random.default_rng(0)
    many = np.clip(np.exp(rng.normal(np.log(130), 0.7, 2000)), 50, 2000).astype(int)
    axis, shard = sharding.plan(pkg, many, [1500] * 1000, 8)
    assert axis == "profiles" and len(shard) == 2000 and shard.max() == 7
    # config 4: few long profiles of unequal length over 8 GPUs do not balance -> shard the contigs instead,
    # by cumulative length (dcpgpu_shard_sequences): equal lengths split evenly, ragged ones by nucleotides
    few = [3000] * 3 + [2000] * 2
    axis, bounds = sharding.plan(pkg, few, [10000] * 10001, 8)
    assert axis == "sequences" and bounds[0] == 0 and bounds[-1] == 10001 and len(bounds) == 9
    assert all(1250 <= b - a <= 1251 for a, b in zip(bounds[:-1], bounds[1:]))
    lens = [100] * 90 + [9100]
    b2 = pkg.shard_sequences(lens, 2)
    assert list(b2) == [0, 90, 91]  # 9000 nt | 9100 nt
    assert list(pkg.shard_sequences([5, 5, 5], 8))[-1] == 3  # more shards than sequences: empty tails
    assert sharding.plan(pkg, few, [10] * 10, 1)[0] == "profiles"


def test_two_rank_shard_and_merge():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    same, complete, loads = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same and complete
    import __graft_entry__ as ge
    assert abs(loads[0] - loads[1]) <= ge.load_pkg().profile_cost(64) + 1e-9  # LPT balance: within one largest profile
