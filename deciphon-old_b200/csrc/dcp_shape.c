/*
 * dcp_shape.c -- how a profile maps onto the GPU, what it costs there, and how profiles are
 * partitioned over devices.
 *
 * The reference partitions profiles over its OpenMP threads by equal COUNT
 * (src/db/profile_reader.c:54-72, xmath_partition_size in include/deciphon/core/xmath.h:24-30) and
 * pays for it with a barrier per sequence (scan.c:239-250).  Here a shard's load is the modelled
 * score-pass time of its profiles: padded width of the profile's kernel class divided by the measured
 * padded-node rate of that class (tools/class_sweep.py on a B200, profiles/r02_class_sweep.jsonl).
 * Cost per node varies by more than 3x between classes, so balancing nominal core length (round 1)
 * leaves the devices that drew the long profiles behind.
 */
#include "dcp_internal.h"

#include <stdlib.h>

#include "dcp_classes.h"

#include <stdio.h>
#include <string.h>

#define ROW(TW, Q, BPS, RATE) {TW, Q, BPS, RATE},
static const struct dcp_class kClasses[] = {DCP_CLASS_TABLE(ROW)};
#undef ROW
enum { kNumClasses = sizeof kClasses / sizeof kClasses[0] };
_Static_assert((int)kNumClasses <= (int)DCP_MAX_CLASSES, "raise DCP_MAX_CLASSES");

/* nodes a class holds: tw warps of 32 lanes, or (tw = 0) the 16 lanes of a half-warp, q nodes per lane */
static unsigned class_capacity(struct dcp_class const *c) { return (c->tw ? c->tw * 32 : 16) * c->q; }

unsigned dcp_num_classes(void) { return kNumClasses; }
struct dcp_class const *dcp_class_at(unsigned cls) { return cls < kNumClasses ? &kClasses[cls] : NULL; }

/*
 * Profile length -> kernel class: the class that holds the profile and minimises padded width / rate.
 * DCPGPU_FORCE_SHAPE="tw,q,bps" (experiments: tools/class_sweep.py) forces a class for every profile it holds.
 */
unsigned dcp_kernel_class(unsigned M)
{
    char const *force = getenv("DCPGPU_FORCE_SHAPE");
    if (force)
    {
        unsigned tw = 0, q = 0, bps = 0;
        if (sscanf(force, "%u,%u,%u", &tw, &q, &bps) == 3)
            for (unsigned c = 0; c < kNumClasses; ++c)
                if (kClasses[c].tw == tw && kClasses[c].q == q && kClasses[c].bps == bps && class_capacity(&kClasses[c]) >= M) return c;
    }
    unsigned best = kNumClasses;
    double best_cost = 0.0;
    for (unsigned c = 0; c < kNumClasses; ++c)
    {
        const unsigned cap = class_capacity(&kClasses[c]);
        if (cap < M || kClasses[c].rate <= 0.0) continue;
        const double cost = (double)cap / kClasses[c].rate;
        if (best == kNumClasses || cost < best_cost) best = c, best_cost = cost;
    }
    return best; /* the table always holds 4096 nodes: best < kNumClasses for every valid M */
}

enum rc dcpgpu_kernel_shape(unsigned core_size, unsigned *warps, unsigned *nodes_per_lane, unsigned *blocks)
{
    if (core_size == 0 || core_size > DCP_PROTEIN_MODEL_CORE_SIZE_MAX)
        return dcp_error(RC_EINVAL, "core size out of range");
    struct dcp_class const *c = &kClasses[dcp_kernel_class(core_size)];
    if (warps) *warps = c->tw ? c->tw : 1; /* half-warp classes: one warp (shared by two pairs) */
    if (nodes_per_lane) *nodes_per_lane = c->q;
    if (blocks) *blocks = c->tw > (unsigned)DCP_MAX_W ? 2 : 1;
    return RC_OK;
}

double dcp_profile_cost(unsigned core_size)
{
    if (core_size == 0) return 0.0;
    if (core_size > DCP_PROTEIN_MODEL_CORE_SIZE_MAX) core_size = DCP_PROTEIN_MODEL_CORE_SIZE_MAX;
    struct dcp_class const *c = &kClasses[dcp_kernel_class(core_size)];
    return (double)class_capacity(c) / (c->rate > 0.0 ? c->rate : 300.0);
}

unsigned dcpgpu_kernel_padded_width(unsigned core_size)
{
    if (core_size == 0 || core_size > DCP_PROTEIN_MODEL_CORE_SIZE_MAX) return 0;
    return class_capacity(&kClasses[dcp_kernel_class(core_size)]);
}

double dcpgpu_profile_cost(unsigned core_size) { return dcp_profile_cost(core_size); }

struct by_cost
{
    double cost;
    unsigned idx;
};

static int cmp_by_cost(void const *a, void const *b)
{
    struct by_cost const *x = a, *y = b;
    if (x->cost != y->cost) return x->cost > y->cost ? -1 : 1; /* descending cost */
    return x->idx < y->idx ? -1 : x->idx > y->idx;              /* stable in profile order */
}

/* Longest-processing-time-first partition by modelled cost; replaces the equal-count split of
 * profile_reader.c:54-72.  Deterministic: every process of a multi-process scan computes the same map. */
enum rc dcpgpu_shard_profiles(unsigned nprofiles, unsigned const *core_sizes, unsigned nshards, unsigned *shard_of)
{
    if (nshards == 0) return dcp_error(RC_EINVAL, "nshards must be positive");
    struct by_cost *order = malloc((nprofiles ? nprofiles : 1) * sizeof *order);
    double *load = calloc(nshards, sizeof *load);
    if (!order || !load)
    {
        free(order), free(load);
        return dcp_error(RC_ENOMEM, "alloc shard tables");
    }
    for (unsigned i = 0; i < nprofiles; ++i) order[i].cost = dcp_profile_cost(core_sizes[i]), order[i].idx = i;
    qsort(order, nprofiles, sizeof *order, cmp_by_cost);
    for (unsigned r = 0; r < nprofiles; ++r)
    {
        unsigned best = 0;
        for (unsigned s = 1; s < nshards; ++s)
            if (load[s] < load[best]) best = s;
        shard_of[order[r].idx] = best;
        load[best] += order[r].cost;
    }
    free(order), free(load);
    return RC_OK;
}

/* Contiguous ranges of sequences with about equal nucleotide totals: bounds[0..nshards], bounds[s]..bounds[s+1]
 * is shard s (the sequence-axis split for databases of few long profiles). */
enum rc dcpgpu_shard_sequences(unsigned nseqs, unsigned const *lens, unsigned nshards, unsigned *bounds)
{
    if (nshards == 0) return dcp_error(RC_EINVAL, "nshards must be positive");
    unsigned long long total = 0, acc = 0;
    for (unsigned i = 0; i < nseqs; ++i) total += lens[i];
    unsigned s = 0;
    bounds[0] = 0;
    for (unsigned i = 0; i < nseqs && s + 1 < nshards; ++i)
    {
        /* boundary s+1 goes where the running total is closest to its share of the nucleotides: before or
         * after sequence i */
        const unsigned long long after = acc + lens[i];
        while (s + 1 < nshards && after * nshards >= total * (unsigned long long)(s + 1))
        {
            const unsigned long long want = total * (unsigned long long)(s + 1); /* scaled by nshards */
            const int before_is_closer = want - acc * nshards < after * nshards - want;
            const unsigned cut = (before_is_closer && i > bounds[s]) ? i : i + 1;
            bounds[++s] = cut;
        }
        acc = after;
    }
    while (s < nshards) bounds[++s] = nseqs;
    return RC_OK;
}
