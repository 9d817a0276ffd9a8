"""Host-side helpers for ONE-PROCESS-PER-GPU scans (torchrun; bench.py --gpus N), no collective on the data path.

Inside one process the library does all of this itself (dcpgpu_mdb_*, csrc/dcp_multi.cpp).  Across processes
the same rules are applied through the same C functions: profiles are partitioned by modelled cost
(dcpgpu_shard_profiles), or -- when few long profiles do not balance -- sequences by cumulative length
(dcpgpu_shard_sequences); every rank scans its part; hits come back tagged with GLOBAL indices and are merged
by (sequence, profile), the order a single-GPU scan returns them in.
"""
import numpy as np


def plan(pkg, core_sizes, seq_lens, world):
    """Axis rule of dcpgpu_mdb_commit(AUTO): ("profiles", shard_of_profile) when the modelled shard costs balance
    within 10 %, else ("sequences", bounds) with world + 1 bounds of contiguous sequence ranges."""
    sizes = np.asarray(core_sizes, np.int64)
    if world <= 1:
        return "profiles", np.zeros(len(sizes), np.uint32)
    shard = pkg.shard_profiles(sizes, world)
    cost = np.array([pkg.profile_cost(m) for m in sizes])
    loads = np.bincount(shard, weights=cost, minlength=world)
    if len(sizes) >= world and loads.max() * world <= 1.10 * loads.sum():
        return "profiles", shard
    return "sequences", pkg.shard_sequences(seq_lens, world)


def shard_indices(pkg, core_sizes, world, rank):
    """Global profile indices owned by `rank` on the profile axis (ascending)."""
    shard = pkg.shard_profiles(core_sizes, world)
    return [i for i in range(len(core_sizes)) if shard[i] == rank]


def local_hits(result, mine, seq0=0):
    """[(seq, global_prof, alt, null, path)] of one rank's result; `mine` maps local -> global profile,
    seq0 is the global index of the rank's first sequence (sequence axis)."""
    out = []
    _, _, alt, null, _ = result.hits()
    for i in range(result.nhits):
        s, p, path = result.hit_at(i)
        out.append((seq0 + s, mine[p], float(alt[i]), float(null[i]), path))
    return out


def merge_hits(per_rank):
    """Concatenate the ranks' hit lists and order them by (sequence, global profile)."""
    merged = [h for hits in per_rank for h in hits]
    merged.sort(key=lambda h: (h[0], h[1]))
    return merged
