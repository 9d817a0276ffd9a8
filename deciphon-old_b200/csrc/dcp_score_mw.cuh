/*
 * dcp_score_mw.cuh -- alt Viterbi score pass for profiles of 257..4096 nodes: a group of warps (one block, or
 * the two blocks of a cluster) per (sequence, profile) pair.
 */
#ifndef DCP_SCORE_MW_CUH
#define DCP_SCORE_MW_CUH
#include "dcp_score.cuh"

namespace
{
/* ----------------------------------------------------------------------------------------- */
/* alt Viterbi, score pass, profiles of 257..4096 nodes: a group of warps per pair            */
/* ----------------------------------------------------------------------------------------- */
/*
 * Same recurrence, same fp32 operation order and the same lane layout as k_score<Q>; node
 * k-1 = gwarp * 32 Q + lane * Q + sub.  257..2048 nodes: W warps of one block (CL = 1);
 * 2049..4096 nodes: W warps in each block of a 2-block cluster (CL = 2, 255 registers x 16 warps do
 * not fit one SM), exchanging through distributed shared memory.  What a single warp exchanges with
 * shuffles is exchanged between warps through shared memory, at most two rendezvous per row:
 *   A   V_M / V_I / D of each warp's last node (D from the warp-local chain), per-warp max of V_M (-> E),
 *       V_N / V_J / V_C of warp 0
 *   C   group-wide OR: did any warp's last D rise when the left neighbour's values came in?
 *       (if so every warp re-reads its left neighbour's new D and repeats -- exact lazy propagation, as inside a
 *       warp).  Not needed with two warps, and skipped when the carry bound below shows that no D can rise.
 * Which (warps, nodes per lane, resident blocks) shapes exist and what they achieve: dcp_classes.h.
 */
/*
 * Carry bound of one warp of a group (lane w holds warp w's): S = an upper bound of the sum of the D->D scores of
 * the warp's nodes after its first (host, dcp_engine.cu), md0 / dd0 = M->D and D->D into its first node.  A carry
 * that enters warp w cannot lift the warp's last D above  max(V_M(left) + md0, D(left) + dd0) + S: when that is
 * below the end of the warp's own chain for every warp, the D values exchanged at A are final and the second
 * rendezvous of the row (C) is skipped -- every warp evaluates the same test on the same exchanged values, so no
 * communication is needed to agree on it.  Otherwise the exact lazy rounds run as before.
 */
struct CarryBound
{
    float S, md0, dd0;
};
#ifndef DCP_CARRY_BOUND
#define DCP_CARRY_BOUND 1
#endif

/* vm_left / d_left / d_own: V_M and local D end of warp w - 1, local D end of warp w, for w = clamp(lane, 1, TW - 1)
 * (every lane evaluates, lanes outside 1..TW-1 are masked: no divergence on the row's critical path) */
template <int TW>
__device__ __forceinline__ bool carry_cannot_rise(const CarryBound &cb, int lane, float vm_left, float d_left,
                                                  float d_own)
{
    const float cin = fmaxf(vm_left + cb.md0, d_left + cb.dd0);
    /* the chain is rounded at every node: 2^-13 relative slack covers 256 roundings of 2^-24 four times over */
    const float bound = cin + cb.S + (9.8e-4f + 1.22e-4f * (fabsf(cin) + fabsf(cb.S)));
    const bool bad = (lane >= 1) & (lane < TW) & (cin > NEG_INF) & (cb.S > NEG_INF) & !(bound <= d_own);
    return !__any_sync(FULL, bad);
}

struct MwShared
{
    alignas(16) float4 rec[2][kMaxGroupWarps]; /* mw_row: per-warp record of the row, by row parity */
    alignas(16) float v_spec[2][4];            /* V_N, V_J, V_C of the row */
    float d_loc[2][kMaxGroupWarps]; /* ends of the warps' own D chains, by row parity */
    float vm_last[2][kMaxGroupWarps], vi_last[2][kMaxGroupWarps], e_warp[2][kMaxGroupWarps];
    float d_last[2][kMaxGroupWarps];
    int flag[2][2];
    unsigned long long item;
    alignas(16) float xch[2][kMaxGroupWarps][8]; /* 2-block groups: Group::exchange buffers ... */
    unsigned long long xbar[2];                  /* ... and their mbarriers */
};

/* One-block groups: skipping the second barrier gains 5 % where 12 warps are resident (168 registers: (3,6,4) 339 ->
 * 357, (4,5,3) 343 -> 361 G padded cells/s) and loses 6..9 % in the 255-register classes ((4,8,2) 468 -> 427, (8,8,1)
 * 417 -> 394: the three bound registers push the row into spills) -- so it is on for the former only. */
#ifndef DCP_CARRY_BOUND_CL1
#define DCP_CARRY_BOUND_CL1(W, BPS) ((W) * (BPS) > 8)
#endif
/*
 * One DP row of a warp group, laid out for the instruction scheduler.  ptxas schedules inside basic blocks, and lazy
 * carry loops are blocks of their own that hold nothing but one dependent add/max chain (round 1's layout: issue slots
 * 43 % busy, `wait` the top stall).  Here the first carry round on either side of the rendezvous is straight-line code
 * (a further round is a rare loop after it), and the work that does not depend on the D chain -- Tin_I, the M->M /
 * I->M part of Tin_M before the rendezvous, the B->M part after it -- sits in the same block, so it fills the chain's
 * latency.  Only maxima are re-associated, every sum is still (V_src + t): bit-identical to the sequential order.
 */
template <int W, int CL, int R, int Q, int BPS, class Tap = NoTap>
__device__ __forceinline__ void mw_row(float (&tm)[5][Q], float (&ti)[5][Q], float (&tx)[5],
                                        const NodeParams<Q> &p, RowState<Q> &rs,
                                        const float *__restrict__ emis_lane, const RowRec *__restrict__ rec_next,
                                        const uint16_t *__restrict__ w_next2, int gw, int lane, int par,
                                        Group<CL, MwShared> &grp, float NB, float JB, float EB, float cE, float cX,
                                        const CarryBound &cb, float &E_out, float &vC_out, Tap *tap = nullptr)
{
    constexpr int TW = W * CL;
    constexpr int ROW = 256 * TW;
    constexpr int S1 = (R + 4) % 5, S2 = (R + 3) % 5, S3 = (R + 2) % 5, S4 = (R + 1) % 5, S5 = R;
    MwShared &sh = *grp.me;

    float vm[Q], vi[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i)
        vm[i] = fmaxf(max3(tm[S1][i] + rs.em[0][i], tm[S2][i] + rs.em[1][i], tm[S3][i] + rs.em[2][i]),
                      fmaxf(tm[S4][i] + rs.em[3][i], tm[S5][i] + rs.em[4][i]));
#pragma unroll
    for (int i = 0; i < Q; ++i)
        vi[i] = fmaxf(max3(ti[S1][i] + rs.eI[0], ti[S2][i] + rs.eI[1], ti[S3][i] + rs.eI[2]),
                      fmaxf(ti[S4][i] + rs.eI[3], ti[S5][i] + rs.eI[4]));
    float vx = fmaxf(max3(tx[S1] + rs.eN[0], tx[S2] + rs.eN[1], tx[S3] + rs.eN[2]),
                     fmaxf(tx[S4] + rs.eN[3], tx[S5] + rs.eN[4]));

    uint32_t code[5];
    codes_of(rs.w1, code);
    load_emis_part<Q, 3, 5, ROW, 128, emis256(TW, Q)>(rs.em, emis_lane, code);
    load_row_insert(rec_next, rs.eI);
    if (gw == 0 && lane < 3) load_row_special(rec_next, rs.eN);
    rs.w1 = rs.w2;
    rs.w2 = __ldg(w_next2);

    float eloc = vm[0];
#pragma unroll
    for (int i = 1; i < Q; ++i) eloc = fmaxf(eloc, vm[i]);
    const float ew = warp_max(eloc);
    float vm_prev = __shfl_up_sync(FULL, vm[Q - 1], 1);
    float vi_prev = __shfl_up_sync(FULL, vi[Q - 1], 1);
    if (lane == 0) vm_prev = NEG_INF, vi_prev = NEG_INF; /* the left warp's values arrive with the rendezvous */
    load_emis_part<Q, 0, 3, ROW, 128, emis256(TW, Q)>(rs.em, emis_lane, code);

    /* D chain inside the warp (nothing from the warp to the left yet), first carry round straight-line */
    float d[Q];
    d[0] = vm_prev + p.MD[0];
#pragma unroll
    for (int i = 1; i < Q; ++i) d[i] = fmaxf(vm[i - 1] + p.MD[i], d[i - 1] + p.DD[i]);
    float old = d[Q - 1];
    float din = __shfl_up_sync(FULL, old, 1);
    if (lane == 0) din = NEG_INF;
    {
        float x = din;
#pragma unroll
        for (int i = 0; i < Q; ++i)
        {
            x = x + p.DD[i];
            d[i] = fmaxf(d[i], x);
            x = d[i];
        }
    }
    bool more = __any_sync(FULL, d[Q - 1] > old);
    /* independent of the D chain: Tin_I, and the part of Tin_M that comes from this warp's own M and I (slot R's old
     * content, row j-5, is dead) */
#pragma unroll
    for (int i = 0; i < Q; ++i)
    {
        const float pm = i == 0 ? vm_prev : vm[i - 1];
        const float pi = i == 0 ? vi_prev : vi[i - 1];
        tm[R][i] = fmaxf(pm + p.MM[i], pi + p.IM[i]);
        ti[R][i] = fmaxf(vm[i] + p.MI[i], vi[i] + p.II[i]);
    }
    while (more)
    {
        old = d[Q - 1];
        din = __shfl_up_sync(FULL, old, 1);
        if (lane == 0) din = NEG_INF;
        float x = din;
#pragma unroll
        for (int i = 0; i < Q; ++i)
        {
            x = x + p.DD[i];
            d[i] = fmaxf(d[i], x);
            x = d[i];
        }
        more = __any_sync(FULL, d[Q - 1] > old);
    }

    /* rendezvous A */
    float E, vN, vJ, vC, din0;
    int xs = 0;
    bool final_d = TW == 2;
    if constexpr (CL == 2)
    {
        const float xN = __shfl_sync(FULL, vx, 0), xJ = __shfl_sync(FULL, vx, 1), xC = __shfl_sync(FULL, vx, 2);
        const float pay_a[8] = {vm[Q - 1], vi[Q - 1], ew, d[Q - 1], xN, xJ, xC, 0.0f};
        xs = grp.exchange(gw, lane, pay_a);
        const float4 *x = reinterpret_cast<const float4 *>(sh.xch[xs]); /* two float4 per warp */
        /* lane w holds warp w's record; E is one redux over the per-warp maxima instead of TW loads and maxima */
        const float4 mine = x[2 * min(lane, TW - 1)];
        E = warp_max(lane < TW ? mine.z : NEG_INF);
        const float4 left = x[2 * max(gw - 1, 0)], spec = x[1];
        if (lane == 0 && gw) vm_prev = left.x, vi_prev = left.y;
        vN = spec.x, vJ = spec.y, vC = spec.z;
        din0 = gw ? left.w : NEG_INF;
        if (DCP_CARRY_BOUND)
        {
            const float4 lw = x[2 * (min(max(lane, 1), TW - 1) - 1)];
            final_d = carry_cannot_rise<TW>(cb, lane, lw.x, lw.w, mine.w);
        }
    }
    else
    {
        if constexpr (Q < 8)
        {
            /* one 16-byte record per warp: {V_M, V_I of its last node, its maximum of V_M, the end of its own D chain};
             * lane w reads warp w's record and E is one redux over the per-warp maxima.  +4..9 % below 8 nodes per lane;
             * at 8 per lane (255 registers, no slack) it costs 4..8 %, so those classes keep the scalar form below */
            if (lane == 31) sh.rec[par][gw] = make_float4(vm[Q - 1], vi[Q - 1], ew, d[Q - 1]);
            if (gw == 0 && lane < 3) sh.v_spec[par][lane] = vx;
            __syncthreads();
            const float4 mine = sh.rec[par][min(lane, TW - 1)];
            E = warp_max(lane < TW ? mine.z : NEG_INF);
            const float4 left = sh.rec[par][max(gw - 1, 0)];
            const float4 spec = *reinterpret_cast<const float4 *>(sh.v_spec[par]);
            if (lane == 0 && gw) vm_prev = left.x, vi_prev = left.y;
            vN = spec.x, vJ = spec.y, vC = spec.z;
            din0 = gw ? left.w : NEG_INF;
            if (DCP_CARRY_BOUND_CL1(W, BPS) && TW > 2)
            {
                const float4 lw = sh.rec[par][min(max(lane, 1), TW - 1) - 1];
                final_d = carry_cannot_rise<TW>(cb, lane, lw.x, lw.w, mine.w);
            }
        }
        else
        {
            if (lane == 31)
            {
                sh.vm_last[par][gw] = vm[Q - 1], sh.vi_last[par][gw] = vi[Q - 1];
                sh.e_warp[par][gw] = ew, sh.d_loc[par][gw] = d[Q - 1];
            }
            if (gw == 0 && lane < 3) sh.v_spec[par][lane] = vx;
            __syncthreads();
            if (lane == 0 && gw) vm_prev = sh.vm_last[par][gw - 1], vi_prev = sh.vi_last[par][gw - 1];
            E = sh.e_warp[par][0];
#pragma unroll
            for (int w = 1; w < TW; ++w) E = fmaxf(E, sh.e_warp[par][w]);
            vN = sh.v_spec[par][0], vJ = sh.v_spec[par][1], vC = sh.v_spec[par][2];
            din0 = gw ? sh.d_loc[par][gw - 1] : NEG_INF;
            if (DCP_CARRY_BOUND_CL1(W, BPS) && TW > 2)
            {
                const int w = min(max(lane, 1), TW - 1);
                final_d = carry_cannot_rise<TW>(cb, lane, sh.vm_last[par][w - 1], sh.d_loc[par][w - 1], sh.d_loc[par][w]);
            }
        }
    }

    /* the carry from the warp to the left enters at lane 0: first round straight-line */
    const float before = __shfl_sync(FULL, d[Q - 1], 31);
    old = d[Q - 1];
    din = __shfl_up_sync(FULL, old, 1);
    if (lane == 0) din = din0;
    {
        float x = lane == 0 ? fmaxf(vm_prev + p.MD[0], din0 + p.DD[0]) : din + p.DD[0];
        d[0] = fmaxf(d[0], x);
        x = d[0];
#pragma unroll
        for (int i = 1; i < Q; ++i)
        {
            x = x + p.DD[i];
            d[i] = fmaxf(d[i], x);
            x = d[i];
        }
    }
    more = __any_sync(FULL, d[Q - 1] > old);
    /* independent of the carry: B, the specials, the B->M part of Tin_M, lane 0's M->M / I->M from the left warp */
    const float B = max3(vN + NB, vJ + JB, E + EB);
    tx[R] = fmaxf(E + cE, vx + cX);
    if (lane == 0) tm[R][0] = fmaxf(vm_prev + p.MM[0], vi_prev + p.IM[0]);
#pragma unroll
    for (int i = 0; i < Q; ++i) tm[R][i] = fmaxf(tm[R][i], B + p.ent[i]);
    while (more)
    {
        old = d[Q - 1];
        din = __shfl_up_sync(FULL, old, 1);
        if (lane == 0) din = din0;
        float x = lane == 0 ? fmaxf(vm_prev + p.MD[0], din0 + p.DD[0]) : din + p.DD[0];
        d[0] = fmaxf(d[0], x);
        x = d[0];
#pragma unroll
        for (int i = 1; i < Q; ++i)
        {
            x = x + p.DD[i];
            d[i] = fmaxf(d[i], x);
            x = d[i];
        }
        more = __any_sync(FULL, d[Q - 1] > old);
    }
    /* more than two warps: a warp whose last D rose has to be re-read by its right neighbour (rare; exact lazy rounds) */
    if (!final_d)
    {
        float bef = before;
        for (int round = 0;; ++round)
        {
            const float after = __shfl_sync(FULL, d[Q - 1], 31);
            if constexpr (CL == 2)
            {
                const float pay_c[8] = {after > bef ? 1.0f : 0.0f, d[Q - 1], 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
                xs = grp.exchange(gw, lane, pay_c);
                float rose = sh.xch[xs][0][0];
#pragma unroll
                for (int w = 1; w < TW; ++w) rose = fmaxf(rose, sh.xch[xs][w][0]);
                if (rose == 0.0f) break;
                din0 = gw ? sh.xch[xs][gw - 1][1] : NEG_INF;
            }
            else
            {
                const int b = round & 1;
                /* publish the new end first: if any warp rose, everybody re-reads its left neighbour after the OR */
                if (lane == 31) sh.d_last[b][gw] = d[Q - 1];
                if (!__syncthreads_or(after > bef)) break;
                din0 = gw ? sh.d_last[b][gw - 1] : NEG_INF;
            }
            bef = after;
            for (;;)
            {
                old = d[Q - 1];
                din = __shfl_up_sync(FULL, old, 1);
                if (lane == 0) din = din0;
                float x = lane == 0 ? fmaxf(vm_prev + p.MD[0], din0 + p.DD[0]) : din + p.DD[0];
                d[0] = fmaxf(d[0], x);
                x = d[0];
#pragma unroll
                for (int i = 1; i < Q; ++i)
                {
                    x = x + p.DD[i];
                    d[i] = fmaxf(d[i], x);
                    x = d[i];
                }
                if (!__any_sync(FULL, d[Q - 1] > old)) break;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < Q; ++i)
    {
        const float pd = i == 0 ? din : d[i - 1];
        tm[R][i] = fmaxf(tm[R][i], pd + p.DM[i]);
    }
    tap_row(tap, d, B);
    E_out = E;
    vC_out = vC;
}

/* BPS = resident blocks per SM the kernel is compiled for (class table): W * BPS warps per SM, i.e. 255 registers
 * a thread for 8 warps, 168 for 12 */
template <int W, int CL, int Q, int BPS>
__global__ void __launch_bounds__(W * 32, CL == 2 ? 1 : BPS)
k_score_mw(const float *__restrict__ emis, const float *__restrict__ trans, const ProfMeta *__restrict__ metas,
           const uint32_t *__restrict__ class_profs, uint32_t n_class_profs, const SeqMeta *__restrict__ seqs,
           uint32_t nseq, uint64_t total_recs, const RowRec *__restrict__ rows,
           const uint16_t *__restrict__ wcodes, const float *__restrict__ spec, float *__restrict__ alt_out,
           uint32_t nprof, unsigned long long *__restrict__ counter, uint32_t seq_tile)
{
    constexpr int TW = W * CL;
    constexpr int ROW = 256 * TW;
    __shared__ MwShared sh;
    Group<CL, MwShared> grp;
    grp.init(&sh);
    const int lane = threadIdx.x & 31, gw = grp.rank * W + (threadIdx.x >> 5);
    const unsigned long long n_items = (unsigned long long)n_class_profs * nseq;
    if (CL == 2) grp.sync(); /* both blocks' shared memory exists before the first remote store */
    if constexpr (CL == 2) grp.exchange_init(W);
    for (;;)
    {
        if (grp.rank == 0 && threadIdx.x == 0)
        {
            unsigned long long it = atomicAdd(counter, 1ULL);
            GRP_PUT(grp, item, it);
        }
        grp.sync();
        const unsigned long long item = sh.item;
        grp.sync();
        if (item >= n_items) break;
        /* (sequence tile, profile, sequence in tile), as in k_score */
        const unsigned long long per_full_tile = (unsigned long long)seq_tile * n_class_profs;
        const uint32_t tile = (uint32_t)(item / per_full_tile);
        const unsigned long long in_tile = item - (unsigned long long)tile * per_full_tile;
        const uint32_t seqs_here = min(seq_tile, nseq - tile * seq_tile);
        const uint32_t prof = class_profs[in_tile / seqs_here];
        const uint32_t s = tile * seq_tile + (uint32_t)(in_tile % seqs_here);
        const ProfMeta pm = metas[prof];
        NodeParams<Q> p;
        load_params<Q>(p, trans + pm.trans_off, 32 * Q * TW, gw * 32 * Q + lane * Q);
        CarryBound cb = {NEG_INF, NEG_INF, NEG_INF};
        if (TW > 2 && lane < TW)
        {
            const float *b = trans + pm.trans_off + 8 * (32 * Q * TW); /* [3][TW] after the eight parameter arrays */
            cb.S = __ldg(b + lane), cb.md0 = __ldg(b + TW + lane), cb.dd0 = __ldg(b + 2 * TW + lane);
        }
        const float *emis_lane = emis + pm.emis_off + gw * 256 + lane * emis_lane_stride(TW, Q);
        const SeqMeta sm = seqs[s];
        const RowRec *recs = rows + (size_t)pm.null_id * total_recs + sm.rec_off;
        const uint16_t *wc = wcodes + sm.rec_off;
        const float *sp = spec + (size_t)s * 16;
        const uint32_t L = sm.len;

        const float NN = sp[0], CC = sp[1], JJ = sp[2], NB = sp[3], CT = sp[4], JB = sp[5];
        const float ET = sp[9], ECC = sp[10], EB = sp[11], EJJ = sp[12];
        const float cE = lane == 0 ? NEG_INF : (lane == 1 ? EJJ : ECC);
        const float cX = lane == 0 ? NN : (lane == 1 ? JJ : CC);

        float tm[5][Q], ti[5][Q], tx[5];
#pragma unroll
        for (int r = 0; r < 5; ++r)
        {
            tx[r] = NEG_INF;
#pragma unroll
            for (int i = 0; i < Q; ++i) tm[r][i] = NEG_INF, ti[r][i] = NEG_INF;
        }
#pragma unroll
        for (int i = 0; i < Q; ++i) tm[4][i] = NB + p.ent[i];
        tx[4] = (gw == 0 && lane == 0) ? NN : NEG_INF;

        RowState<Q> rs;
#pragma unroll
        for (int l = 0; l < 5; ++l) rs.eN[l] = NEG_INF;
        {
            uint32_t code[5];
            codes_of(__ldg(wc + 1), code);
            load_emis<Q, ROW, 128, emis256(TW, Q)>(rs.em, emis_lane, code);
        }
        load_row_insert(recs + 1, rs.eI);
        if (gw == 0 && lane < 3) load_row_special(recs + 1, rs.eN);
        rs.w1 = __ldg(wc + min(2u, L));
        rs.w2 = __ldg(wc + min(3u, L));

        float E = NEG_INF, vC = NEG_INF;
        uint32_t j = 1;
#define MW_ROW(WW, CC, RR, QQ) mw_row<WW, CC, RR, QQ, BPS>
#define MW_ARGS(jj) recs + min((uint32_t)(jj) + 1u, L), wc + min((uint32_t)(jj) + 3u, L), gw, lane, (int)((jj)&1u), grp
        for (; j + 4 <= L; j += 5)
        {
            MW_ROW(W, CL, 0, Q)(tm, ti, tx, p, rs, emis_lane, MW_ARGS(j), NB, JB, EB, cE, cX, cb, E, vC);
            MW_ROW(W, CL, 1, Q)(tm, ti, tx, p, rs, emis_lane, MW_ARGS(j + 1), NB, JB, EB, cE, cX, cb, E, vC);
            MW_ROW(W, CL, 2, Q)(tm, ti, tx, p, rs, emis_lane, MW_ARGS(j + 2), NB, JB, EB, cE, cX, cb, E, vC);
            MW_ROW(W, CL, 3, Q)(tm, ti, tx, p, rs, emis_lane, MW_ARGS(j + 3), NB, JB, EB, cE, cX, cb, E, vC);
            MW_ROW(W, CL, 4, Q)(tm, ti, tx, p, rs, emis_lane, MW_ARGS(j + 4), NB, JB, EB, cE, cX, cb, E, vC);
        }
        if (j <= L) MW_ROW(W, CL, 0, Q)(tm, ti, tx, p, rs, emis_lane, MW_ARGS(j), NB, JB, EB, cE, cX, cb, E, vC);
        if (j + 1 <= L) MW_ROW(W, CL, 1, Q)(tm, ti, tx, p, rs, emis_lane, MW_ARGS(j + 1), NB, JB, EB, cE, cX, cb, E, vC);
        if (j + 2 <= L) MW_ROW(W, CL, 2, Q)(tm, ti, tx, p, rs, emis_lane, MW_ARGS(j + 2), NB, JB, EB, cE, cX, cb, E, vC);
        if (j + 3 <= L) MW_ROW(W, CL, 3, Q)(tm, ti, tx, p, rs, emis_lane, MW_ARGS(j + 3), NB, JB, EB, cE, cX, cb, E, vC);
#undef MW_ARGS
#undef MW_ROW
        if (gw == 0 && lane == 0) alt_out[(size_t)s * nprof + prof] = fmaxf(E + ET, vC + CT);
    }
}

} // namespace
#endif
