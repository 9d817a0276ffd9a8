/*
 * dcp_trace.cu -- traceback pass for hits only (imm_prod.path of imm_dp_viterbi on the alt dp,
 * consumed by prod_fwrite, src/server/prod.c:153-181).
 *
 * k_trace<Q> re-runs the alt recurrence of one (sequence, profile) hit with the same warp /
 * lane / register layout and the same fp32 operation order as k_score<Q> (so T[L] is bit-equal),
 * and records for every state and row which incoming (transition, source length) won.  The
 * winner is imm's: first maximum, strict '>', over incoming transitions in canonical order and,
 * inside one transition, over the source's emission length ascending -- evaluated on the rounded
 * sums fl(fl(Tin + e) + t), exactly like a start-position interpreter would.
 *
 * Canonical incoming order (imm's own order is not recoverable from the reference tree):
 *   M_k: B, M_{k-1}, I_{k-1}, D_{k-1}    I_k: M_k, I_k    D_k: M_{k-1}, D_{k-1}
 *   E: M_1, M_2, D_2, M_3, D_3, ...      N: S, N    B: S, N, J, E    J: E, J   C: E, C   T: E, C
 *
 * Backpointers: one uint16 per (row, node) [mcode:4 | icode:4 | dcode:3] stored as
 * [row][sub-node][lane] (64-byte coalesced stores) and one uint32 per row for the specials
 * [ecode:15 | n:3 | b:4 | j:3 | c:3 | t:3].
 */
#include "dcp_classes.h"
#include "dcp_kernels.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

namespace
{

struct TraceJob
{
    uint32_t seq, prof;
    uint64_t cell_off; /* uint16 units */
    uint64_t row_off;  /* uint32 units */
};

__device__ __forceinline__ void first_max5(const float (&s)[5], float t, int base, float &best, int &code)
{
#pragma unroll
    for (int l = 0; l < 5; ++l)
    {
        float v = s[l] + t;
        if (v > best) best = v, code = base + l;
    }
}


/*
 * The trace kernels run one copy of the row code and rotate the five-row rings with register moves (95 moves against
 * ~1700 instructions of a row).  The score kernels' five-fold unrolling (static ring slots, no moves) made the trace
 * kernels 9 x 1700 instructions long, and the instruction cache misses were their top stall (ncu: no_instruction
 * 3.0 of 7.3 cycles per issue, profiles/r01_k_trace_mw_long_ncu.txt).
 */
template <int Q>
__device__ __forceinline__ void ring_rotate(float (&tm)[5][Q], float (&ti)[5][Q], float (&tn)[5], float (&tj)[5],
                                            float (&tc)[5])
{
#pragma unroll
    for (int i = 0; i < Q; ++i)
    {
        float a = tm[0][i], b = ti[0][i];
#pragma unroll
        for (int s = 0; s < 4; ++s) tm[s][i] = tm[s + 1][i], ti[s][i] = ti[s + 1][i];
        tm[4][i] = a, ti[4][i] = b;
    }
    float a = tn[0], b = tj[0], c = tc[0];
#pragma unroll
    for (int s = 0; s < 4; ++s) tn[s] = tn[s + 1], tj[s] = tj[s + 1], tc[s] = tc[s + 1];
    tn[4] = a, tj[4] = b, tc[4] = c;
}

/* LN = lanes per hit: 32, or 16 for the half-warp classes (the trace kernel then runs the hit twice, on both halves
 * of the warp: identical values, identical stores -- traceback is not worth a second code path) */
template <int Q, int R, int LN>
__device__ __forceinline__ void trace_row(float (&tm)[5][Q], float (&ti)[5][Q], float (&tn)[5], float (&tj)[5],
                                          float (&tc)[5], const NodeParams<Q> &p,
                                          const float *__restrict__ emis_lane, const RowRec *__restrict__ rec,
                                          uint32_t wcode, int lane, const float *__restrict__ sp, uint16_t *__restrict__ cell_bp,
                                          uint32_t *__restrict__ row_bp, float &T_out)
{
    constexpr int QP = Q <= 4 ? 4 : 8;
    constexpr int S[5] = {(R + 4) % 5, (R + 3) % 5, (R + 2) % 5, (R + 1) % 5, R};
    const float NN = sp[0], CC = sp[1], JJ = sp[2], NB = sp[3], CT = sp[4], JB = sp[5];
    const float ET = sp[9], ECC = sp[10], EB = sp[11], EJJ = sp[12];
    struct { float eI[5], eN[5]; } in;
    load_row_insert(rec, in.eI);
    load_row_special(rec, in.eN);
    uint32_t code[5];
    codes_of(wcode, code);
    float em[5][Q];
    load_emis<Q, LN * QP, LN * 4>(em, emis_lane, code);

    /* W[j-l][X][l] for every emitting state: the five candidate sums per state */
    float sM[Q][5], sI[Q][5], vm[Q], vi[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i)
    {
#pragma unroll
        for (int l = 0; l < 5; ++l)
        {
            sM[i][l] = tm[S[l]][i] + em[l][i];
            sI[i][l] = ti[S[l]][i] + in.eI[l];
        }
        vm[i] = fmaxf(max3(sM[i][0], sM[i][1], sM[i][2]), fmaxf(sM[i][3], sM[i][4]));
        vi[i] = fmaxf(max3(sI[i][0], sI[i][1], sI[i][2]), fmaxf(sI[i][3], sI[i][4]));
    }
    float sN[5], sJ[5], sC[5];
#pragma unroll
    for (int l = 0; l < 5; ++l)
    {
        sN[l] = tn[S[l]] + in.eN[l];
        sJ[l] = tj[S[l]] + in.eN[l];
        sC[l] = tc[S[l]] + in.eN[l];
    }
    /* sums of the node to the left of this lane's first node */
    float pM0[5], pI0[5];
#pragma unroll
    for (int l = 0; l < 5; ++l)
    {
        pM0[l] = __shfl_up_sync(FULL, sM[Q - 1][l], 1, LN);
        pI0[l] = __shfl_up_sync(FULL, sI[Q - 1][l], 1, LN);
        if (lane == 0) pM0[l] = NEG_INF, pI0[l] = NEG_INF;
    }

    /* D chain (order: M_{k-1} by length, then D_{k-1}) */
    float d[Q];
    int dcode[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i)
    {
        float best = NEG_INF;
        int code = 0;
        if (i == 0)
            first_max5(pM0, p.MD[0], 0, best, code);
        else
        {
            first_max5(sM[i - 1], p.MD[i], 0, best, code);
            float x = d[i - 1] + p.DD[i];
            if (x > best) best = x, code = 5;
        }
        d[i] = best, dcode[i] = code;
    }
    float din;
    for (;;)
    {
        float old = d[Q - 1];
        din = __shfl_up_sync(FULL, old, 1, LN);
        if (lane == 0) din = NEG_INF;
        float x = din;
#pragma unroll
        for (int i = 0; i < Q; ++i)
        {
            x = x + p.DD[i];
            if (x > d[i]) d[i] = x, dcode[i] = 5;
            x = d[i];
        }
        if (!__any_sync(FULL, d[Q - 1] > old)) break;
    }

    /* E: first max over M_1, M_2, D_2, ... ; lanes hold increasing k */
    float ebest = NEG_INF;
    int ecode = 0;
#pragma unroll
    for (int i = 0; i < Q; ++i)
    {
        int k0 = lane * Q + i; /* k - 1 */
        float zero = 0.0f;
        first_max5(sM[i], zero, k0 * 6, ebest, ecode);
        if (k0 >= 1)
        {
            float v = d[i] + zero;
            if (v > ebest) ebest = v, ecode = k0 * 6 + 5;
        }
    }
    float E = warp_max(ebest);
    unsigned who = __ballot_sync(FULL, ebest == E);
    ecode = __shfl_sync(FULL, ecode, who ? __ffs(who) - 1 : 0);

    /* specials */
    float best;
    int ncode = 0, bcode = 0, jcode = 0, ccode = 0, tcode = 0;
    best = NEG_INF;
    first_max5(sN, NN, 1, best, ncode);
    float tinN = best;
    best = NEG_INF;
    first_max5(sN, NB, 1, best, bcode);
    first_max5(sJ, JB, 6, best, bcode);
    {
        float v = E + EB;
        if (v > best) best = v, bcode = 11;
    }
    float B = best;
    best = E + EJJ, jcode = 0;
    first_max5(sJ, JJ, 1, best, jcode);
    float tinJ = best;
    best = E + ECC, ccode = 0;
    first_max5(sC, CC, 1, best, ccode);
    float tinC = best;
    best = E + ET, tcode = 0;
    first_max5(sC, CT, 1, best, tcode);
    T_out = best;
    tn[R] = tinN, tj[R] = tinJ, tc[R] = tinC;
    if (lane == 0)
        *row_bp = (uint32_t)ecode | (uint32_t)ncode << 15 | (uint32_t)bcode << 18 | (uint32_t)jcode << 22 |
                  (uint32_t)ccode << 25 | (uint32_t)tcode << 28;

    /* core Tin + backpointers */
#pragma unroll
    for (int i = 0; i < Q; ++i)
    {
        float mb = B + p.ent[i];
        int mcode = 0;
        if (i == 0)
        {
            first_max5(pM0, p.MM[0], 1, mb, mcode);
            first_max5(pI0, p.IM[0], 6, mb, mcode);
            float v = din + p.DM[0];
            if (v > mb) mb = v, mcode = 11;
        }
        else
        {
            first_max5(sM[i - 1], p.MM[i], 1, mb, mcode);
            first_max5(sI[i - 1], p.IM[i], 6, mb, mcode);
            float v = d[i - 1] + p.DM[i];
            if (v > mb) mb = v, mcode = 11;
        }
        float ib = NEG_INF;
        int icode = 0;
        first_max5(sM[i], p.MI[i], 0, ib, icode);
        first_max5(sI[i], p.II[i], 5, ib, icode);
        tm[R][i] = mb;
        ti[R][i] = ib;
        cell_bp[i * LN + lane] = (uint16_t)(mcode | icode << 4 | dcode[i] << 8);
    }
}

template <int Q, int LN>
__global__ void __launch_bounds__(128) k_trace(const float *__restrict__ emis, const float *__restrict__ trans,
                                               const ProfMeta *__restrict__ metas,
                                               const SeqMeta *__restrict__ seqs, uint64_t total_rows,
                                               const RowRec *__restrict__ rows, const uint16_t *__restrict__ wcodes,
                                               const float *__restrict__ spec, const TraceJob *__restrict__ jobs, uint32_t njobs,
                                               uint16_t *__restrict__ cell_bp, uint32_t *__restrict__ row_bp,
                                               float *__restrict__ alt_out)
{
    const int lane = threadIdx.x & (LN - 1); /* index inside the hit's lanes; with LN = 16 both halves run the same hit */
    uint32_t job = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (job >= njobs) return;
    TraceJob tj_ = jobs[job];
    ProfMeta pm = metas[tj_.prof];
    SeqMeta sm = seqs[tj_.seq];
    NodeParams<Q> p;
    load_params<Q>(p, trans + pm.trans_off, LN * Q, lane * Q);
    const float *emis_lane = emis + pm.emis_off + lane * 4;
    const RowRec *r = rows + (size_t)pm.null_id * total_rows + sm.rec_off + 1; /* record of row 1 */
    const uint16_t *wc = wcodes + sm.rec_off; /* wc[j] = window of row j */
    const float *sp = spec + (size_t)tj_.seq * 16;
    uint16_t *cb = cell_bp + tj_.cell_off;
    uint32_t *rb = row_bp + tj_.row_off;
    const uint32_t L = sm.len;

    float tm[5][Q], ti[5][Q], tn[5], tjr[5], tc[5];
#pragma unroll
    for (int s = 0; s < 5; ++s)
    {
        tn[s] = tjr[s] = tc[s] = NEG_INF;
#pragma unroll
        for (int i = 0; i < Q; ++i) tm[s][i] = NEG_INF, ti[s][i] = NEG_INF;
    }
    /* row 0: S starts; N <- S, B <- S, M_k <- B; codes are all 0 */
    const float NN = sp[0], NB = sp[3];
#pragma unroll
    for (int i = 0; i < Q; ++i)
    {
        tm[4][i] = NB + p.ent[i];
        cb[i * LN + lane] = 0;
    }
    tn[4] = NN;
    if (lane == 0) rb[0] = 0;

    float T = NEG_INF;
    uint32_t j = 1;
    constexpr uint32_t CS = Q * LN; /* cell backpointers per row */
#pragma unroll 1
    for (; j <= L; ++j)
    {
        trace_row<Q, 0, LN>(tm, ti, tn, tjr, tc, p, emis_lane, r + (j - 1), wc[j], lane, sp, cb + (size_t)j * CS, rb + j, T);
        ring_rotate<Q>(tm, ti, tn, tjr, tc);
    }
    if (lane == 0) alt_out[job] = T;
}

/* ----------------------------------------------------------------------------------------- */
/* traceback pass for hits on profiles of 257..2048 nodes: W warps (one block) per hit        */
/* ----------------------------------------------------------------------------------------- */
struct MwTraceShared
{
    alignas(16) float xch[2][kMaxGroupWarps][12]; /* 2-block groups: Group::exchange buffers ... */
    unsigned long long xbar[2];                   /* ... and their mbarriers */
    float pM[2][kMaxGroupWarps][5], pI[2][kMaxGroupWarps][5]; /* the five sums of each warp's last node */
    float d_last[2][kMaxGroupWarps];
    float e_best[2][kMaxGroupWarps];
    int e_code[2][kMaxGroupWarps];
    int flag[2][2];
};

/* cell backpointers of a row: [sub-node][warp][lane] */
template <int W, int CL, int R, int Q>
__device__ __forceinline__ void trace_row_mw(float (&tm)[5][Q], float (&ti)[5][Q], float (&tn)[5], float (&tj)[5],
                                             float (&tc)[5], const NodeParams<Q> &p,
                                             const float *__restrict__ emis_lane, const RowRec *__restrict__ rec,
                                             uint32_t wcode, int warp, int lane, int par,
                                             Group<CL, MwTraceShared> &grp,
                                             const float *__restrict__ sp, uint16_t *__restrict__ cell_bp,
                                             uint32_t *__restrict__ row_bp, float &T_out)
{
    constexpr int TW = W * CL; /* `warp` is the warp's index in the whole group, 0..TW-1 */
    constexpr int ROW = 256 * TW;
    constexpr int S[5] = {(R + 4) % 5, (R + 3) % 5, (R + 2) % 5, (R + 1) % 5, R};
    MwTraceShared &sh = *grp.me;
    const float NN = sp[0], CC = sp[1], JJ = sp[2], NB = sp[3], CT = sp[4], JB = sp[5];
    const float ET = sp[9], ECC = sp[10], EB = sp[11], EJJ = sp[12];
    struct { float eI[5], eN[5]; } in;
    load_row_insert(rec, in.eI);
    load_row_special(rec, in.eN);
    uint32_t code[5];
    codes_of(wcode, code);
    float em[5][Q];
    load_emis<Q, ROW>(em, emis_lane, code);

    float sM[Q][5], sI[Q][5];
#pragma unroll
    for (int i = 0; i < Q; ++i)
#pragma unroll
        for (int l = 0; l < 5; ++l)
        {
            sM[i][l] = tm[S[l]][i] + em[l][i];
            sI[i][l] = ti[S[l]][i] + in.eI[l];
        }
    float sN[5], sJ[5], sC[5];
#pragma unroll
    for (int l = 0; l < 5; ++l)
    {
        sN[l] = tn[S[l]] + in.eN[l];
        sJ[l] = tj[S[l]] + in.eN[l];
        sC[l] = tc[S[l]] + in.eN[l];
    }
    float pM0[5], pI0[5];
#pragma unroll
    for (int l = 0; l < 5; ++l)
    {
        pM0[l] = __shfl_up_sync(FULL, sM[Q - 1][l], 1);
        pI0[l] = __shfl_up_sync(FULL, sI[Q - 1][l], 1);
        if (lane == 0) pM0[l] = NEG_INF, pI0[l] = NEG_INF; /* the left warp's values arrive with barrier A */
    }

    /* D chain (order: M_{k-1} by length, then D_{k-1}), first inside the warp: its first node has no source yet */
    float d[Q];
    int dcode[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i)
    {
        float best = NEG_INF;
        int c = 0;
        if (i == 0)
            first_max5(pM0, p.MD[0], 0, best, c);
        else
        {
            first_max5(sM[i - 1], p.MD[i], 0, best, c);
            float x = d[i - 1] + p.DD[i];
            if (x > best) best = x, c = 5;
        }
        d[i] = best, dcode[i] = c;
    }
    float din = NEG_INF;
    for (;;)
    {
        float old = d[Q - 1];
        din = __shfl_up_sync(FULL, old, 1);
        if (lane == 0) din = NEG_INF;
        float x = din;
#pragma unroll
        for (int i = 0; i < Q; ++i)
        {
            x = x + p.DD[i];
            if (x > d[i]) d[i] = x, dcode[i] = 5;
            x = d[i];
        }
        if (!__any_sync(FULL, d[Q - 1] > old)) break;
    }
#ifndef DCP_CLUSTER_XCH
#define DCP_CLUSTER_XCH 1
#endif
    float E, ew;
    int ecode;
    if constexpr (CL == 2 && DCP_CLUSTER_XCH)
    {
        /* A: the five sums of each warp's last node and the end of its local D chain, one exchange (no cluster barrier) */
        const float pay_a[12] = {sM[Q - 1][0], sM[Q - 1][1], sM[Q - 1][2], sM[Q - 1][3], sM[Q - 1][4], sI[Q - 1][0],
                                 sI[Q - 1][1], sI[Q - 1][2], sI[Q - 1][3], sI[Q - 1][4], d[Q - 1],     0.0f};
        int s = grp.exchange(warp, lane, pay_a);
        if (lane == 0 && warp)
        {
#pragma unroll
            for (int l = 0; l < 5; ++l) pM0[l] = sh.xch[s][warp - 1][l], pI0[l] = sh.xch[s][warp - 1][5 + l];
            first_max5(pM0, p.MD[0], 0, d[0], dcode[0]);
        }
        float din0 = warp ? sh.xch[s][warp - 1][10] : NEG_INF;
        for (;;)
        {
            const float before = __shfl_sync(FULL, d[Q - 1], 31);
            for (;;)
            {
                float old = d[Q - 1];
                din = __shfl_up_sync(FULL, old, 1);
                if (lane == 0) din = din0;
                float x = din;
#pragma unroll
                for (int i = 0; i < Q; ++i)
                {
                    x = x + p.DD[i];
                    if (x > d[i]) d[i] = x, dcode[i] = 5;
                    x = d[i];
                }
                if (!__any_sync(FULL, d[Q - 1] > old)) break;
            }
            /* E candidates of this warp with the D values of this round (final when no warp's last D rose) */
            float ebest = NEG_INF;
            ecode = 0;
#pragma unroll
            for (int i = 0; i < Q; ++i)
            {
                int k0 = warp * 32 * Q + lane * Q + i; /* k - 1 */
                float zero = 0.0f;
                first_max5(sM[i], zero, k0 * 6, ebest, ecode);
                if (k0 >= 1)
                {
                    float v = d[i] + zero;
                    if (v > ebest) ebest = v, ecode = k0 * 6 + 5;
                }
            }
            ew = warp_max(ebest);
            unsigned who = __ballot_sync(FULL, ebest == ew);
            ecode = __shfl_sync(FULL, ecode, who ? __ffs(who) - 1 : 0);
            const float pay_c[12] = {d[Q - 1] > before ? 1.0f : 0.0f, d[Q - 1], ew, __int_as_float(ecode), 0.0f, 0.0f, 0.0f,
                                     0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
            s = grp.exchange(warp, lane, pay_c); /* C */
            float rose = sh.xch[s][0][0];
#pragma unroll
            for (int w = 1; w < TW; ++w) rose = fmaxf(rose, sh.xch[s][w][0]);
            if (rose == 0.0f) break;
            din0 = warp ? sh.xch[s][warp - 1][1] : NEG_INF;
        }
        E = sh.xch[s][0][2];
        ecode = __float_as_int(sh.xch[s][0][3]);
#pragma unroll
        for (int w = 1; w < TW; ++w)
            if (sh.xch[s][w][2] > E) E = sh.xch[s][w][2], ecode = __float_as_int(sh.xch[s][w][3]);
    }
    else
    {
        if (lane == 31)
        {
#pragma unroll
            for (int l = 0; l < 5; ++l)
            {
                GRP_PUT(grp, pM[par][warp][l], sM[Q - 1][l]);
                GRP_PUT(grp, pI[par][warp][l], sI[Q - 1][l]);
            }
            GRP_PUT(grp, d_last[0][warp], d[Q - 1]);
        }
        grp.sync(); /* A: the five sums of each warp's last node and the end of its local D chain */
        if (lane == 0 && warp)
        {
#pragma unroll
            for (int l = 0; l < 5; ++l) pM0[l] = sh.pM[par][warp - 1][l], pI0[l] = sh.pI[par][warp - 1][l];
            /* M_{k-1} -> D_k candidates of the warp's first node come first in the order; its D_{k-1} follows below */
            first_max5(pM0, p.MD[0], 0, d[0], dcode[0]);
        }

        /* carries between warps, lazily; the per-warp E candidates ride on the same barrier (C): when no warp's last D
         * rose, every D was final and so are the candidates published in that round */
        for (int round = 0;; ++round)
        {
            const int b = round & 1;
            if (round > 0)
            {
                if (lane == 31) GRP_PUT(grp, d_last[b][warp], d[Q - 1]);
                grp.sync(); /* B */
            }
            const float din0 = warp ? sh.d_last[b][warp - 1] : NEG_INF;
            const float before = __shfl_sync(FULL, d[Q - 1], 31);
            for (;;)
            {
                float old = d[Q - 1];
                din = __shfl_up_sync(FULL, old, 1);
                if (lane == 0) din = din0;
                float x = din;
#pragma unroll
                for (int i = 0; i < Q; ++i)
                {
                    x = x + p.DD[i];
                    if (x > d[i]) d[i] = x, dcode[i] = 5;
                    x = d[i];
                }
                if (!__any_sync(FULL, d[Q - 1] > old)) break;
            }
            const float after = __shfl_sync(FULL, d[Q - 1], 31);

            /* E: first max over M_1, M_2, D_2, ... ; warps and lanes hold increasing k */
            float ebest = NEG_INF;
            ecode = 0;
#pragma unroll
            for (int i = 0; i < Q; ++i)
            {
                int k0 = warp * 32 * Q + lane * Q + i; /* k - 1 */
                float zero = 0.0f;
                first_max5(sM[i], zero, k0 * 6, ebest, ecode);
                if (k0 >= 1)
                {
                    float v = d[i] + zero;
                    if (v > ebest) ebest = v, ecode = k0 * 6 + 5;
                }
            }
            ew = warp_max(ebest);
            unsigned who = __ballot_sync(FULL, ebest == ew);
            ecode = __shfl_sync(FULL, ecode, who ? __ffs(who) - 1 : 0);
            if (lane == 0)
            {
                GRP_PUT(grp, e_best[par][warp], ew);
                GRP_PUT(grp, e_code[par][warp], ecode);
            }
            /* round 0 already used the left warp's local chain end; a warp whose last D rose has to be re-read */
            if (!grp.any(after > before, sh.flag, CL == 2 ? grp.peer->flag : sh.flag, b)) break; /* C */
        }
        E = sh.e_best[par][0];
        ecode = sh.e_code[par][0];
#pragma unroll
        for (int w = 1; w < TW; ++w)
            if (sh.e_best[par][w] > E) E = sh.e_best[par][w], ecode = sh.e_code[par][w];
    }

    /* specials: every warp keeps its own copy of the N/J/C rings (identical values) */
    float best;
    int ncode = 0, bcode = 0, jcode = 0, ccode = 0, tcode = 0;
    best = NEG_INF;
    first_max5(sN, NN, 1, best, ncode);
    float tinN = best;
    best = NEG_INF;
    first_max5(sN, NB, 1, best, bcode);
    first_max5(sJ, JB, 6, best, bcode);
    {
        float v = E + EB;
        if (v > best) best = v, bcode = 11;
    }
    float B = best;
    best = E + EJJ, jcode = 0;
    first_max5(sJ, JJ, 1, best, jcode);
    float tinJ = best;
    best = E + ECC, ccode = 0;
    first_max5(sC, CC, 1, best, ccode);
    float tinC = best;
    best = E + ET, tcode = 0;
    first_max5(sC, CT, 1, best, tcode);
    T_out = best;
    tn[R] = tinN, tj[R] = tinJ, tc[R] = tinC;
    if (warp == 0 && lane == 0)
        *row_bp = (uint32_t)ecode | (uint32_t)ncode << 15 | (uint32_t)bcode << 18 | (uint32_t)jcode << 22 |
                  (uint32_t)ccode << 25 | (uint32_t)tcode << 28;

#pragma unroll
    for (int i = 0; i < Q; ++i)
    {
        float mb = B + p.ent[i];
        int mcode = 0;
        if (i == 0)
        {
            first_max5(pM0, p.MM[0], 1, mb, mcode);
            first_max5(pI0, p.IM[0], 6, mb, mcode);
            float v = din + p.DM[0];
            if (v > mb) mb = v, mcode = 11;
        }
        else
        {
            first_max5(sM[i - 1], p.MM[i], 1, mb, mcode);
            first_max5(sI[i - 1], p.IM[i], 6, mb, mcode);
            float v = d[i - 1] + p.DM[i];
            if (v > mb) mb = v, mcode = 11;
        }
        float ib = NEG_INF;
        int icode = 0;
        first_max5(sM[i], p.MI[i], 0, ib, icode);
        first_max5(sI[i], p.II[i], 5, ib, icode);
        tm[R][i] = mb;
        ti[R][i] = ib;
        cell_bp[i * (32 * TW) + warp * 32 + lane] = (uint16_t)(mcode | icode << 4 | dcode[i] << 8);
    }
}

template <int W, int CL, int Q>
__global__ void __launch_bounds__(W * 32) k_trace_mw(const float *__restrict__ emis, const float *__restrict__ trans,
                                                     const ProfMeta *__restrict__ metas,
                                                     const SeqMeta *__restrict__ seqs, uint64_t total_rows,
                                                     const RowRec *__restrict__ rows,
                                                     const uint16_t *__restrict__ wcodes,
                                                     const float *__restrict__ spec,
                                                     const TraceJob *__restrict__ jobs, uint32_t njobs,
                                                     uint16_t *__restrict__ cell_bp, uint32_t *__restrict__ row_bp,
                                                     float *__restrict__ alt_out)
{
    constexpr int TW = W * CL;
    __shared__ MwTraceShared sh;
    Group<CL, MwTraceShared> grp;
    grp.init(&sh);
    if constexpr (CL == 2) grp.exchange_init(W);
    const int lane = threadIdx.x & 31, warp = grp.rank * W + (threadIdx.x >> 5);
    const uint32_t job = blockIdx.x / CL;
    if (job >= njobs) return;
    TraceJob tj_ = jobs[job];
    ProfMeta pm = metas[tj_.prof];
    SeqMeta sm = seqs[tj_.seq];
    NodeParams<Q> p;
    load_params<Q>(p, trans + pm.trans_off, 32 * Q * TW, warp * 32 * Q + lane * Q);
    const float *emis_lane = emis + pm.emis_off + warp * 256 + lane * 4;
    const RowRec *r = rows + (size_t)pm.null_id * total_rows + sm.rec_off; /* r[j] = record of row j */
    const uint16_t *wc = wcodes + sm.rec_off;
    const float *sp = spec + (size_t)tj_.seq * 16;
    uint16_t *cb = cell_bp + tj_.cell_off;
    uint32_t *rb = row_bp + tj_.row_off;
    const uint32_t L = sm.len;
    constexpr uint32_t CS = Q * 32 * TW;

    float tm[5][Q], ti[5][Q], tn[5], tjr[5], tc[5];
#pragma unroll
    for (int s = 0; s < 5; ++s)
    {
        tn[s] = tjr[s] = tc[s] = NEG_INF;
#pragma unroll
        for (int i = 0; i < Q; ++i) tm[s][i] = NEG_INF, ti[s][i] = NEG_INF;
    }
    const float NN = sp[0], NB = sp[3];
#pragma unroll
    for (int i = 0; i < Q; ++i)
    {
        tm[4][i] = NB + p.ent[i];
        cb[i * (32 * TW) + warp * 32 + lane] = 0;
    }
    tn[4] = NN;
    if (warp == 0 && lane == 0) rb[0] = 0;
    if (CL == 2) grp.sync(); /* both blocks' shared memory exists before the first remote store */

    float T = NEG_INF;
    uint32_t j = 1;
#define TR_ARGS(jj) r + (jj), wc[(jj)], warp, lane, (int)((jj)&1u), grp, sp, cb + (size_t)(jj) * CS, rb + (jj), T
#pragma unroll 1
    for (; j <= L; ++j)
    {
        trace_row_mw<W, CL, 0, Q>(tm, ti, tn, tjr, tc, p, emis_lane, TR_ARGS(j));
        ring_rotate<Q>(tm, ti, tn, tjr, tc);
    }
#undef TR_ARGS
    if (warp == 0 && lane == 0) alt_out[job] = T;
    if (CL == 2) grp.sync(); /* no block may exit while its peer can still store into its shared memory */
}

/*
 * Follow the backpointers from (T, row L) to S.  mode 0: count steps; mode 1: write them
 * (reversed, then flipped in place) at steps + step_off[job].
 */
enum { W_S, W_N, W_B, W_E, W_J, W_C, W_T, W_M, W_I, W_D };

__global__ void k_walk(const ProfMeta *__restrict__ metas, const SeqMeta *__restrict__ seqs,
                       const TraceJob *__restrict__ jobs, uint32_t njobs, const uint16_t *__restrict__ cell_bp,
                       const uint32_t *__restrict__ row_bp, int mode, uint32_t *__restrict__ nsteps,
                       const uint64_t *__restrict__ step_off, dcp_step *__restrict__ steps,
                       uint32_t *__restrict__ errors)
{
    uint32_t job = blockIdx.x * blockDim.x + threadIdx.x;
    if (job >= njobs) return;
    TraceJob tj = jobs[job];
    const uint32_t Q = metas[tj.prof].Q, W = metas[tj.prof].W, LN = metas[tj.prof].LN;
    const uint32_t CS = Q * LN * W;
    const uint16_t *cb = cell_bp + tj.cell_off;
    const uint32_t *rb = row_bp + tj.row_off;
    const uint32_t L = seqs[tj.seq].len;
    dcp_step *out = mode ? steps + step_off[job] : nullptr;
    const uint32_t limit = mode ? nsteps[job] : 0xffffffffu;

    int st = W_T;
    uint32_t k = 0, r = L, len = 0, n = 0;
    bool bad = false;
    for (;;)
    {
        uint16_t id;
        switch (st)
        {
        case W_M: id = (uint16_t)(PROTEIN_MATCH_STATE | k); break;
        case W_I: id = (uint16_t)(PROTEIN_INSERT_STATE | k); break;
        case W_D: id = (uint16_t)(PROTEIN_DELETE_STATE | k); break;
        case W_S: id = PROTEIN_S_STATE; break;
        case W_N: id = PROTEIN_N_STATE; break;
        case W_B: id = PROTEIN_B_STATE; break;
        case W_E: id = PROTEIN_E_STATE; break;
        case W_J: id = PROTEIN_J_STATE; break;
        case W_C: id = PROTEIN_C_STATE; break;
        default: id = PROTEIN_T_STATE; break;
        }
        if (mode)
        {
            if (n >= limit) { bad = true; break; }
            out[n].state_id = id, out[n].seqlen = (uint8_t)len;
        }
        n++;
        if (st == W_S) break;
        if (n > 0x7ffffff0u) { bad = true; break; }
        uint32_t rw = rb[r];
        uint32_t src_len = 0;
        int nst = st;
        uint32_t nk = k;
        if (st == W_M || st == W_I || st == W_D)
        {
            uint32_t node = k - 1;
            uint16_t c = cb[(size_t)r * CS + (node % Q) * (LN * W) + node / Q]; /* [sub][warp][lane] */
            if (st == W_M)
            {
                uint32_t m = c & 15;
                if (m == 0) nst = W_B;
                else if (m <= 5) nst = W_M, nk = k - 1, src_len = m;
                else if (m <= 10) nst = W_I, nk = k - 1, src_len = m - 5;
                else nst = W_D, nk = k - 1;
            }
            else if (st == W_I)
            {
                uint32_t m = (c >> 4) & 15;
                if (m <= 4) nst = W_M, src_len = m + 1;
                else nst = W_I, src_len = m - 4;
            }
            else
            {
                uint32_t m = (c >> 8) & 7;
                if (m <= 4) nst = W_M, nk = k - 1, src_len = m + 1;
                else nst = W_D, nk = k - 1;
            }
            if (nk == 0 && nst != W_B) { bad = true; break; }
        }
        else if (st == W_T)
        {
            uint32_t m = (rw >> 28) & 7;
            if (m == 0) nst = W_E; else nst = W_C, src_len = m;
        }
        else if (st == W_C)
        {
            uint32_t m = (rw >> 25) & 7;
            if (m == 0) nst = W_E; else nst = W_C, src_len = m;
        }
        else if (st == W_J)
        {
            uint32_t m = (rw >> 22) & 7;
            if (m == 0) nst = W_E; else nst = W_J, src_len = m;
        }
        else if (st == W_B)
        {
            uint32_t m = (rw >> 18) & 15;
            if (m == 0) nst = W_S;
            else if (m <= 5) nst = W_N, src_len = m;
            else if (m <= 10) nst = W_J, src_len = m - 5;
            else nst = W_E;
        }
        else if (st == W_N)
        {
            uint32_t m = (rw >> 15) & 7;
            if (m == 0) nst = W_S; else nst = W_N, src_len = m;
        }
        else /* W_E */
        {
            uint32_t m = rw & 0x7fff;
            nk = m / 6 + 1;
            uint32_t sub = m % 6;
            if (sub <= 4) nst = W_M, src_len = sub + 1; else nst = W_D;
        }
        if (src_len > r) { bad = true; break; }
        if (nst == W_S && r != 0) { bad = true; break; }
        r -= src_len;
        st = nst, k = nk, len = src_len;
    }
    if (bad)
    {
        atomicAdd(errors, 1u);
        if (!mode) nsteps[job] = 0;
        return;
    }
    if (!mode)
        nsteps[job] = n;
    else
        for (uint32_t a = 0, b = n - 1; a < b; ++a, --b)
        {
            dcp_step t = out[a];
            out[a] = out[b], out[b] = t;
        }
}

template <int Q, int LN>
void launch_trace(cudaStream_t st, uint32_t njobs, const dcpgpu_db *db, const dcpgpu_seqs *sq, const RowRec *rows,
                  const uint16_t *wcodes, const float *spec, const TraceJob *jobs, uint16_t *cell_bp, uint32_t *row_bp, float *alt)
{
    k_trace<Q, LN><<<(njobs + 3) / 4, 128, 0, st>>>(db->d_emis, db->d_trans, db->d_metas, sq->d_metas, sq->total + sq->nseq, rows,
                                                wcodes, spec, jobs, njobs, cell_bp, row_bp, alt);
}

/* one warp per hit (TW = 1) or a group of TW warps, one block or the two blocks of a cluster */
template <int TW, int Q>
void launch_trace_class(cudaStream_t st, uint32_t njobs, const dcpgpu_db *db, const dcpgpu_seqs *sq, const RowRec *rows,
                        const uint16_t *wcodes, const float *spec, const TraceJob *jobs, uint16_t *cell_bp,
                        uint32_t *row_bp, float *alt)
{
    if constexpr (TW <= 1)
        launch_trace<Q, TW == 1 ? 32 : 16>(st, njobs, db, sq, rows, wcodes, spec, jobs, cell_bp, row_bp, alt);
    else
    {
        constexpr int CL = TW > kMaxW ? 2 : 1, W = TW / CL;
        launch_group(k_trace_mw<W, CL, Q>, CL, njobs * CL, W * 32, st, db->d_emis, db->d_trans, db->d_metas, sq->d_metas,
                     sq->total + sq->nseq, rows, wcodes, spec, jobs, njobs, cell_bp, row_bp, alt);
    }
}

} // namespace

enum rc dcp_trace_hits(dcpgpu_db *db, dcpgpu_seqs *sq, dcpgpu_result *res, const RowRec *d_rows,
                       const uint16_t *d_wcodes, const float *d_spec, uint64_t *launches)
{
    cudaStream_t st = db->stream;
    const size_t nhits = res->hits.size();
    /* backpointer budget per batch */
    size_t free_b = 0, total_b = 0;
    CU_TRY(cudaMemGetInfo(&free_b, &total_b));
    const size_t budget = std::min<size_t>(std::max<size_t>(free_b / 2, (size_t)64 << 20), (size_t)16 << 30);

    const std::vector<float> &score_alt = res->hit_alt; /* score pass result of each hit */

    size_t done = 0;
    while (done < nhits)
    {
        /* take hits while their backpointers fit the budget; group the batch by class */
        std::vector<TraceJob> jobs;
        size_t cells = 0, rowsz = 0, end = done;
        while (end < nhits)
        {
            const HitRec &h = res->hits[end];
            uint32_t QW = db->metas[h.prof].Q * db->metas[h.prof].W * db->metas[h.prof].LN; /* padded nodes */
            size_t L1 = (size_t)sq->metas[h.seq].len + 1;
            size_t need = L1 * QW * sizeof(uint16_t) + L1 * sizeof(uint32_t);
            if (!jobs.empty() && (cells * 2 + rowsz * 4 + need > budget)) break;
            jobs.push_back({h.seq, h.prof, cells, rowsz});
            cells += L1 * QW;
            rowsz += L1;
            ++end;
        }
        const uint32_t nj = (uint32_t)jobs.size();
        /* order jobs by class so each launch sees a contiguous range */
        std::vector<uint32_t> order(nj);
        for (uint32_t i = 0; i < nj; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
            return db->metas[jobs[a].prof].cls < db->metas[jobs[b].prof].cls;
        });
        std::vector<TraceJob> sorted(nj);
        for (uint32_t i = 0; i < nj; ++i) sorted[i] = jobs[order[i]];

        DevBuf b_jobs, b_cells, b_rows, b_alt, b_n, b_off, b_err, b_steps;
        CU_TRY(b_jobs.alloc(nj * sizeof(TraceJob), db));
        CU_TRY(b_cells.alloc(cells * sizeof(uint16_t), db));
        CU_TRY(b_rows.alloc(rowsz * sizeof(uint32_t), db));
        CU_TRY(b_alt.alloc(nj * sizeof(float), db));
        CU_TRY(b_n.alloc(nj * sizeof(uint32_t), db));
        CU_TRY(b_off.alloc(nj * sizeof(uint64_t), db));
        CU_TRY(b_err.alloc(sizeof(uint32_t), db));
        CU_TRY(cudaMemcpyAsync(b_jobs.p, sorted.data(), nj * sizeof(TraceJob), cudaMemcpyHostToDevice, st));
        CU_TRY(cudaMemsetAsync(b_err.p, 0, sizeof(uint32_t), st));
        /* one launch per kernel class, side by side on the main and side streams: a class often holds a handful of
         * hits, and a trace launch lasts as long as its longest sequence however few hits it has */
        StreamFan fan(db);
        CU_TRY(fan.fork());
        for (uint32_t a = 0; a < nj;)
        {
            cudaStream_t st = fan.next();
            uint32_t cls = db->metas[sorted[a].prof].cls, b = a;
            while (b < nj && db->metas[sorted[b].prof].cls == cls) ++b;
            {
                const dcp_class &kc = *dcp_class_at(cls);
                bool launched = false;
                /* the trace kernels depend on (warps per pair, nodes per lane) only, not on the occupancy variant */
#define X(TW, Q, BPS, RATE)                                                                                       \
    if (!launched && kc.tw == TW && kc.q == Q)                                                                    \
    {                                                                                                             \
        launch_trace_class<TW, Q>(st, b - a, db, sq, d_rows, d_wcodes, d_spec, b_jobs.as<TraceJob>() + a,         \
                                  b_cells.as<uint16_t>(), b_rows.as<uint32_t>(), b_alt.as<float>() + a);          \
        launched = true;                                                                                          \
    }
                DCP_CLASS_TABLE(X)
#undef X
                if (!launched) return dcp_error(RC_EFAIL, "no trace kernel for this kernel class");
            }
            (*launches)++;
            a = b;
        }
        CU_TRY(fan.join());
        CU_TRY(cudaGetLastError());
        k_walk<<<(nj + 63) / 64, 64, 0, st>>>(db->d_metas, sq->d_metas, b_jobs.as<TraceJob>(), nj,
                                              b_cells.as<uint16_t>(), b_rows.as<uint32_t>(), 0, b_n.as<uint32_t>(),
                                              nullptr, nullptr, b_err.as<uint32_t>());
        (*launches)++;
        std::vector<uint32_t> ns(nj);
        std::vector<float> talt(nj);
        uint32_t nerr = 0;
        CU_TRY(cudaMemcpyAsync(ns.data(), b_n.p, nj * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(talt.data(), b_alt.p, nj * sizeof(float), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(&nerr, b_err.p, sizeof nerr, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        if (nerr) return dcp_error(RC_EFAIL, "traceback walked off the DP matrix");
        std::vector<uint64_t> off(nj);
        uint64_t tot = 0;
        for (uint32_t i = 0; i < nj; ++i) off[i] = tot, tot += ns[i];
        for (uint32_t i = 0; i < nj; ++i)
        {
            size_t hit = done + order[i];
            if (memcmp(&talt[i], &score_alt[hit], sizeof(float)) != 0)
                return dcp_error(RC_EFAIL, "trace pass and score pass disagree on the alt log-likelihood");
        }
        CU_TRY(b_steps.alloc(std::max<uint64_t>(tot, 1) * sizeof(dcp_step), db));
        CU_TRY(cudaMemcpyAsync(b_off.p, off.data(), nj * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        k_walk<<<(nj + 63) / 64, 64, 0, st>>>(db->d_metas, sq->d_metas, b_jobs.as<TraceJob>(), nj,
                                              b_cells.as<uint16_t>(), b_rows.as<uint32_t>(), 1, b_n.as<uint32_t>(),
                                              b_off.as<uint64_t>(), b_steps.as<dcp_step>(), b_err.as<uint32_t>());
        (*launches)++;
        std::vector<dcp_step> got(tot);
        CU_TRY(cudaMemcpyAsync(got.data(), b_steps.p, tot * sizeof(dcp_step), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(&nerr, b_err.p, sizeof nerr, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        if (nerr) return dcp_error(RC_EFAIL, "traceback walked off the DP matrix");
        /* append in hit order */
        std::vector<uint32_t> inv(nj);
        for (uint32_t i = 0; i < nj; ++i) inv[order[i]] = i;
        for (uint32_t h = 0; h < nj; ++h)
        {
            uint32_t i = inv[h];
            HitRec &hr = res->hits[done + h];
            hr.step_off = res->steps.size();
            hr.nsteps = ns[i];
            res->steps.insert(res->steps.end(), got.begin() + off[i], got.begin() + off[i] + ns[i]);
        }
        done = end;
    }
    return RC_OK;
}
