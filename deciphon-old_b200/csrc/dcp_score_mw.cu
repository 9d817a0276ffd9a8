/* dcp_score_mw.cu -- launches of the warp-group score kernels k_score_mw<W, CL, Q, BPS> (class table rows with
 * more than one warp per pair). */
#include "dcp_classes.h"
#include "dcp_score_mw.cuh"

namespace
{
template <int TW, int Q, int BPS>
cudaError_t launch_mw(int sm_count, cudaStream_t st, const ScoreArgs &a)
{
    if constexpr (TW > 1)
    {
        constexpr int CL = TW > kMaxW ? 2 : 1, W = TW / CL;
        /* persistent grid: BPS blocks per SM; two-block groups: one cluster per pair of SMs */
        const unsigned blocks = CL == 2 ? (unsigned)(sm_count / 2 * 2) : (unsigned)(sm_count * BPS);
        return launch_group(k_score_mw<W, CL, Q, BPS>, CL, blocks, W * 32, st, a.emis, a.trans, a.metas,
                            a.class_profs, a.n_class, a.seqs, a.nseq, a.total_recs, a.rows, a.wcodes, a.spec, a.alt,
                            a.nprof, a.counter, a.seq_tile);
    }
    else
        return cudaErrorInvalidValue; /* one-warp classes: dcp_score_sw.cu */
}
} // namespace

cudaError_t dcp_launch_score_mw(const dcp_class &c, int sm_count, cudaStream_t st, const ScoreArgs &a)
{
#define X(TW, Q, BPS, RATE)                                                                                 \
    if (TW > 1 && c.tw == TW && c.q == Q && c.bps == BPS) return launch_mw<TW, Q, BPS>(sm_count, st, a);
    DCP_CLASS_TABLE(X)
#undef X
    return cudaErrorInvalidValue;
}
