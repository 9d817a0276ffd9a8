"""Host-side sharding helpers for one-process-per-GPU scans (no collective on the data path).

Profiles are partitioned by cumulative core length (dcpgpu_shard_profiles, LPT greedy); every rank
scans all sequences against its shard; hits come back tagged with the GLOBAL profile index and are
merged by (sequence, profile) order -- the order a single-GPU scan returns them in.
"""
import numpy as np


def shard_indices(pkg, core_sizes, world, rank):
    """Global profile indices owned by `rank` (ascending)."""
    shard = pkg.shard_profiles(core_sizes, world)
    return [i for i in range(len(core_sizes)) if shard[i] == rank]


def local_hits(result, mine):
    """[(seq, global_prof, alt, null, path)] of one rank's dcpgpu result; `mine` maps local -> global."""
    out = []
    alt, null = result.alt_loglik, result.null_loglik
    for i in range(result.nhits):
        s, p, path = result.hit_at(i)
        out.append((s, mine[p], float(alt[s, p]), float(null[s, p]), path))
    return out


def merge_hits(per_rank):
    """Concatenate the ranks' hit lists and order them by (sequence, global profile)."""
    merged = [h for hits in per_rank for h in hits]
    merged.sort(key=lambda h: (h[0], h[1]))
    return merged


def plan(pkg, core_sizes, nseqs, world):
    """Choose the axis to shard over: profiles by cumulative core length when that balances
    (imbalance <= 10 %), otherwise sequences (every rank holds all profiles and scans a contiguous slice).
    Returns ("profiles", shard_of_profile) or ("sequences", [(lo, hi)] per rank)."""
    sizes = np.asarray(core_sizes, np.int64)
    if world <= 1:
        return "profiles", np.zeros(len(sizes), np.uint32)
    shard = pkg.shard_profiles(sizes, world)
    loads = np.bincount(shard, weights=sizes, minlength=world)
    if len(sizes) >= world and loads.max() <= 1.10 * loads.mean():
        return "profiles", shard
    per = -(-nseqs // world)
    return "sequences", [(min(r * per, nseqs), min((r + 1) * per, nseqs)) for r in range(world)]
