/* dcp_score_sw.cu -- launches of the one-warp-per-pair score kernels k_score<Q> (kernel classes 1..8). */
#include "dcp_classes.h"
#include "dcp_score.cuh"

#include <cstdlib>

namespace
{
#ifndef DCP_STAGED_VARIANTS
#define DCP_STAGED_VARIANTS 0 /* 1 (variant builds, tools/build_variant.sh): also compile the shared-memory staging experiments */
#endif
#if DCP_STAGED_VARIANTS
template <int Q, int MODE>
cudaError_t launch_score_staged(int nblocks, cudaStream_t st, const ScoreArgs &a)
{
    /* experiments: the 4/5-nt emission lines through a per-warp shared ring, filled two rows ahead by bulk copies +
     * mbarriers (MODE 1, DCPGPU_TMA=1) or by per-lane cp.async (MODE 2, DCPGPU_TMA=2) */
    const int warps = score_warps(Q), LINE = 32 * (Q <= 4 ? 4 : 8);
    const size_t smem = (size_t)warps * 4 * LINE * sizeof(float) + (size_t)warps * 2 * sizeof(uint64_t);
    cudaFuncSetAttribute(k_score<Q, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_score<Q, MODE><<<nblocks, warps * 32, smem, st>>>(a.emis, a.trans, a.metas, a.class_profs, a.n_class, a.seqs, a.nseq,
                                                        a.total_recs, a.rows, a.wcodes, a.spec, a.alt, a.nprof, a.counter,
                                                        a.seq_tile);
    return cudaGetLastError();
}
#endif
template <int Q>
cudaError_t launch_score(int nblocks, cudaStream_t st, const ScoreArgs &a)
{
#if DCP_STAGED_VARIANTS
    /* measured, both slower than the product path: M = 200: bulk copies 372 vs 557 GCUPS (round 1); cp.async 559 vs 581
     * at 8 nodes per lane and 545 vs 576 at 7 (round 2) -- DESIGN.md 6.1 */
    static const int staged = getenv("DCPGPU_TMA") ? atoi(getenv("DCPGPU_TMA")) : 0;
    if (staged == 1) return launch_score_staged<Q, 1>(nblocks, st, a);
    if (staged == 2) return launch_score_staged<Q, 2>(nblocks, st, a);
#endif
    /* no shared memory: give the whole unified array to L1 (emission lines, row records) */
    cudaFuncSetAttribute(k_score<Q, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
    k_score<Q, 0><<<nblocks, score_warps(Q) * 32, 0, st>>>(a.emis, a.trans, a.metas, a.class_profs, a.n_class,
                                                           a.seqs, a.nseq, a.total_recs, a.rows, a.wcodes, a.spec,
                                                           a.alt, a.nprof, a.counter, a.seq_tile);
    return cudaGetLastError();
}
/* two pairs per warp, 16 lanes each (class table rows with TW = 0) */
template <int Q>
cudaError_t launch_score_h(int nblocks, cudaStream_t st, const ScoreArgs &a)
{
    cudaFuncSetAttribute(k_score_h<Q>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
    k_score_h<Q><<<nblocks, score_warps(Q) * 32, 0, st>>>(a.emis, a.trans, a.metas, a.class_profs, a.n_class, a.seqs, a.nseq,
                                                         a.total_recs, a.rows, a.wcodes, a.spec, a.alt, a.nprof,
                                                         a.counter, a.seq_tile);
    return cudaGetLastError();
}
} // namespace

cudaError_t dcp_launch_score(const dcp_class &c, int sm_count, cudaStream_t st, const ScoreArgs &a)
{
    const int nblocks = sm_count; /* persistent: one block per SM (as many warps as the register file holds) */
#define X(TW, Q, BPS, RATE)                                                                                 \
    if (TW <= 1 && c.tw == TW && c.q == Q)                                                                  \
    {                                                                                                       \
        static_assert(TW > 1 || BPS == score_warps(Q), "class table: warps per block of k_score<Q>");       \
        if (TW == 1) return launch_score<(TW == 1 ? Q : 1)>(nblocks, st, a);                                \
        return launch_score_h<(TW == 0 ? Q : 4)>(nblocks, st, a);                                           \
    }
    DCP_CLASS_TABLE(X)
#undef X
    return cudaErrorInvalidValue;
}
