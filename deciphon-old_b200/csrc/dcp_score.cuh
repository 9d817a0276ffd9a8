/*
 * dcp_score.cuh -- alt Viterbi score pass for profiles of up to 256 nodes: one warp per (sequence, profile) pair
 * (imm_dp_viterbi on the alt dp, src/server/scan_thread.c:117; recurrence: DESIGN.md section 3).
 */
#ifndef DCP_SCORE_CUH
#define DCP_SCORE_CUH
#ifndef DCP_NOTAIL_MAXQ
#define DCP_NOTAIL_MAXQ 7 /* up to here: whole five-row groups, no tail copies of the row code */
#endif
#include "dcp_kernels.cuh"

#include <type_traits>

namespace
{
/* ----------------------------------------------------------------------------------------- */
/* alt Viterbi, score pass                                                                   */
/* ----------------------------------------------------------------------------------------- */
/* ---- TMA (cp.async.bulk) + mbarrier helpers for the streamed 4/5-nt emission lines ---- */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
/* one bulk copy global -> shared (SASS UBLKCP), completion counted in bytes on `bar` */
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; ++spin)
    {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
        if (spin > (1u << 26)) __trap(); /* a lost bulk copy must not hang the GPU */
    }
}

/* What the traceback pass (dcp_trace.cu) needs from a row besides the ring: the row's final D values of this lane's
 * nodes and B.  The score kernels pass no tap (NoTap: nothing is generated). */
struct NoTap
{
};
template <int Q>
struct RowTap
{
    float d[Q], B;
};
template <class Tap, int Q>
__device__ __forceinline__ void tap_row(Tap *tap, const float (&d)[Q], float B)
{
    if constexpr (!std::is_same<Tap, NoTap>::value)
    {
#pragma unroll
        for (int i = 0; i < Q; ++i) tap->d[i] = d[i];
        tap->B = B;
    }
}

/* loads in flight for the next row(s) */
template <int Q>
struct RowState
{
    float em[5][Q];    /* match emissions of the row about to be processed */
    float eI[5], eN[5];
    uint32_t w1;       /* window of the row after it (addresses of the next emission loads) */
    uint32_t w2;       /* window two rows ahead, in flight */
    uint32_t w3;       /* TMA variant: window three rows ahead */
};

/*
 * One DP row.  R = ring slot this row writes ((j-1) % 5); the slot holding row j-l is
 * (R - l + 5) % 5, so slot R still holds row j-5 while it is read.
 * Lanes 0,1,2 also carry the N, J, C special states (tx ring); cE/cX are their lane-specific
 * E->X and X->X scores.  Returns E[j] and this lane's V_X[j].
 *
 * Software pipeline (no load is consumed in the row that issues it):
 *   rs.em            row j's match emissions, issued during row j-1
 *   rs.eI / rs.eN    row j's shared emissions, issued early in row j-1
 *   rs.w1            window of row j+1, loaded during row j-1: addresses of row j+1's emission loads
 *   rs.w2            window of row j+2, loaded here
 */
/* per-warp staging of the streamed lines: [2 stages][4-nt line, 5-nt line][32 * QP floats] + 2 mbarriers */
struct TmaCtx
{
    float *ring;
    uint64_t *bar;
    uint32_t g; /* rows issued so far by this warp: stage = g & 1, phase parity = (g >> 1) & 1 */
};

/* MODE: how the streamed 4- and 5-nt emission lines of a row reach the registers.  0: LDG.128 one row ahead (the product
 * path).  1: bulk copies (cp.async.bulk, SASS UBLKCP) into a per-warp shared ring two rows ahead, one mbarrier per stage.
 * 2: per-lane 16-byte asynchronous copies (cp.async.cg, SASS LDGSTS) into the same ring two rows ahead: no registers
 * are held across the row and the copy is issued where the source puts it. */
template <int Q, int R, int MODE, class Tap = NoTap>
__device__ __forceinline__ void score_row(float (&tm)[5][Q], float (&ti)[5][Q], float (&tx)[5],
                                          const NodeParams<Q> &p, RowState<Q> &rs,
                                          const float *__restrict__ emis_lane,
                                          const RowRec *__restrict__ rec_next,
                                          const uint16_t *__restrict__ w_next2, int lane, float NB, float JB,
                                          float EB, float cE, float cX, float &E_out, float &vx_out, TmaCtx &tc,
                                          Tap *tap = nullptr)
{
    constexpr int S1 = (R + 4) % 5, S2 = (R + 3) % 5, S3 = (R + 2) % 5, S4 = (R + 1) % 5, S5 = R;
    constexpr int QP = Q <= 4 ? 4 : 8, LINE = 32 * QP;
    static_assert(MODE == 0 || !emis256(1, Q), "the shared-memory staging experiments read the split line layout: build with -DDCP_EMIS256_ALL=0");

    if constexpr (MODE != 0)
    {
        /* this row's 4- and 5-nt lines were copied into the stage two rows ago */
        const uint32_t st = tc.g & 1u;
        if constexpr (MODE == 1) mbar_wait(tc.bar + st, (tc.g >> 1) & 1u);
        else asm volatile("cp.async.wait_group 1;" ::: "memory"); /* all but the group of the previous row */
#pragma unroll
        for (int l = 3; l < 5; ++l)
        {
            const float4 *line = reinterpret_cast<const float4 *>(tc.ring + (st * 2 + (l - 3)) * LINE);
            float4 a = line[lane];
            float t[8] = {a.x, a.y, a.z, a.w, 0.f, 0.f, 0.f, 0.f};
            if (Q > 4)
            {
                float4 b = line[32 + lane];
                t[4] = b.x, t[5] = b.y, t[6] = b.z, t[7] = b.w;
            }
#pragma unroll
            for (int i = 0; i < Q; ++i) rs.em[l][i] = t[i];
        }
    }

    float vm[Q], vi[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i)
        vm[i] = fmaxf(max3(tm[S1][i] + rs.em[0][i], tm[S2][i] + rs.em[1][i], tm[S3][i] + rs.em[2][i]),
                      fmaxf(tm[S4][i] + rs.em[3][i], tm[S5][i] + rs.em[4][i]));

#pragma unroll
    for (int i = 0; i < Q; ++i)
        vi[i] = fmaxf(max3(ti[S1][i] + rs.eI[0], ti[S2][i] + rs.eI[1], ti[S3][i] + rs.eI[2]),
                      fmaxf(ti[S4][i] + rs.eI[3], ti[S5][i] + rs.eI[4]));
    /* special state carried by this lane */
    float vx = fmaxf(max3(tx[S1] + rs.eN[0], tx[S2] + rs.eN[1], tx[S3] + rs.eN[2]),
                     fmaxf(tx[S4] + rs.eN[3], tx[S5] + rs.eN[4]));

    /* issue row j+1, part 1: the 4- and 5-nt lines (256 and 1024 codes: the likely L1 misses),
     * the shared emissions (their registers were just consumed) and the window two rows ahead */
    uint32_t code[5];
    codes_of(rs.w1, code);
    if constexpr (MODE == 2)
    {
        /* the stage is consumed (vm above used its values): refill it with the lines of row j+2, every lane its own
         * 16-byte quads (it reads back only what it copied itself: no barrier) */
        const uint32_t st = tc.g & 1u;
#pragma unroll
        for (int l = 3; l < 5; ++l)
        {
            const uint32_t cd = l == 3 ? 84u + (rs.w2 & 255u) : 340u + (rs.w2 & 1023u);
            const float *src = emis_lane + (size_t)cd * LINE;
            const uint32_t dst = smem_u32(tc.ring + (st * 2 + (l - 3)) * LINE + lane * 4);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            if (Q > 4) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 512u), "l"(src + 128) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        tc.g++;
        rs.w1 = rs.w2;
        rs.w2 = rs.w3;
        rs.w3 = __ldg(w_next2);
    }
    else if constexpr (MODE == 1)
    {
        /* the stage is consumed (vm above used its values): refill it with the lines of row j+2 */
        __syncwarp();
        if (lane == 0)
        {
            const uint32_t st = tc.g & 1u;
            const float *base = emis_lane; /* lane 0: start of the profile's table */
            mbar_expect_tx(tc.bar + st, 2 * LINE * 4);
            tma_load_1d(tc.ring + (st * 2 + 0) * LINE, base + (size_t)(84u + (rs.w2 & 255u)) * LINE, LINE * 4, tc.bar + st);
            tma_load_1d(tc.ring + (st * 2 + 1) * LINE, base + (size_t)(340u + (rs.w2 & 1023u)) * LINE, LINE * 4, tc.bar + st);
        }
        tc.g++;
        rs.w1 = rs.w2;
        rs.w2 = rs.w3;
        rs.w3 = __ldg(w_next2);
    }
    else
    {
        load_emis_part<Q, 3, 5, 32 * QP, 128, emis256(1, Q)>(rs.em, emis_lane, code);
        rs.w1 = rs.w2;
        rs.w2 = __ldg(w_next2);
    }
    load_row_insert(rec_next, rs.eI);
    if (lane < 3) load_row_special(rec_next, rs.eN);

    /* E[j]: every M_k -> E is 0 and D_k <= max V_M because MD, DD <= 0 (checked at commit) */
    float eloc = vm[0];
#pragma unroll
    for (int i = 1; i < Q; ++i) eloc = fmaxf(eloc, vm[i]);
    float E = warp_max(eloc);

    /* node k0-1 lives in the previous lane */
    float vm_prev = __shfl_up_sync(FULL, vm[Q - 1], 1);
    if (lane == 0) vm_prev = NEG_INF;

    /* D chain: local pass with no carry-in, then exact lazy propagation across lanes */
    float d[Q];
    d[0] = vm_prev + p.MD[0];
#pragma unroll
    for (int i = 1; i < Q; ++i) d[i] = fmaxf(vm[i - 1] + p.MD[i], d[i - 1] + p.DD[i]);
    /* first propagation round: unconditional and straight-line, so that the compiler can fill its
     * dependency stalls with the independent work that follows */
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

    float vi_prev = __shfl_up_sync(FULL, vi[Q - 1], 1);
    if (lane == 0) vi_prev = NEG_INF;

    /* issue row j+1, part 2: the short lines (L1 resident) */
    load_emis_part<Q, 0, 3, 32 * QP, 128, emis256(1, Q)>(rs.em, emis_lane, code);

    /* B[j] = max(V_N + NB, V_J + JB, E + (EJ+JB)) */
    float vN = __shfl_sync(FULL, vx, 0);
    float vJ = __shfl_sync(FULL, vx, 1);
    float B = max3(vN + NB, vJ + JB, E + EB);
    tx[R] = fmaxf(E + cE, vx + cX);

    /* everything of Tin that does not involve D (slot R's old content, row j-5, is dead) */
#pragma unroll
    for (int i = 0; i < Q; ++i)
    {
        float pm = i == 0 ? vm_prev : vm[i - 1];
        float pi = i == 0 ? vi_prev : vi[i - 1];
        tm[R][i] = max3(B + p.ent[i], pm + p.MM[i], pi + p.IM[i]);
        ti[R][i] = fmaxf(vm[i] + p.MI[i], vi[i] + p.II[i]);
    }
    /* a carry that crossed a whole lane keeps propagating (rare) */
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
#pragma unroll
    for (int i = 0; i < Q; ++i)
    {
        float pd = i == 0 ? din : d[i - 1];
        tm[R][i] = fmaxf(tm[R][i], pd + p.DM[i]);
    }
    tap_row(tap, d, B);
    E_out = E;
    vx_out = vx;
}

/* recs / wc = record and window of row 0 of this sequence (L+1 of each) */
template <int Q, int MODE>
__device__ __forceinline__ float score_pair(const NodeParams<Q> &p, const float *__restrict__ emis_lane,
                                            const RowRec *__restrict__ recs, const uint16_t *__restrict__ wc,
                                            uint32_t L, const float *__restrict__ sp, int lane, TmaCtx &tc)
{
    const float NN = sp[0], CC = sp[1], JJ = sp[2], NB = sp[3], CT = sp[4], JB = sp[5];
    const float ET = sp[9], ECC = sp[10], EB = sp[11], EJJ = sp[12];
    const float cE = lane == 0 ? NEG_INF : (lane == 1 ? EJJ : ECC);
    const float cX = lane == 0 ? NN : (lane == 1 ? JJ : CC);

    float tm[5][Q], ti[5][Q], tx[5];
#pragma unroll
    for (int s = 0; s < 5; ++s)
    {
        tx[s] = NEG_INF;
#pragma unroll
        for (int i = 0; i < Q; ++i) tm[s][i] = NEG_INF, ti[s][i] = NEG_INF;
    }
    /* row 0: S = 0, B[0] = NB, Tin_N[0] = NN, Tin_Mk[0] = B[0] + entry_k */
#pragma unroll
    for (int i = 0; i < Q; ++i) tm[4][i] = NB + p.ent[i];
    tx[4] = lane == 0 ? NN : NEG_INF;

    /* pipeline prologue: row 1's loads, windows of rows 2 and 3 */
    RowState<Q> rs;
#pragma unroll
    for (int l = 0; l < 5; ++l) rs.eN[l] = NEG_INF;
    {
        uint32_t code[5];
        codes_of(__ldg(wc + 1), code);
        load_emis<Q, 32 * (Q <= 4 ? 4 : 8), 128, emis256(1, Q)>(rs.em, emis_lane, code);
    }
    load_row_insert(recs + 1, rs.eI);
    if (lane < 3) load_row_special(recs + 1, rs.eN);
    rs.w1 = __ldg(wc + min(2u, L));
    rs.w2 = __ldg(wc + min(3u, L));
    rs.w3 = 0;
    if constexpr (MODE == 2)
    {
        constexpr int LINE = 32 * (Q <= 4 ? 4 : 8);
        const uint32_t wa = __ldg(wc + 1), wb = rs.w1;
#pragma unroll
        for (int k = 0; k < 2; ++k)
        {
            const uint32_t st = (tc.g + k) & 1u, w = k ? wb : wa;
#pragma unroll
            for (int l = 3; l < 5; ++l)
            {
                const uint32_t cd = l == 3 ? 84u + (w & 255u) : 340u + (w & 1023u);
                const float *src = emis_lane + (size_t)cd * LINE;
                const uint32_t dst = smem_u32(tc.ring + (st * 2 + (l - 3)) * LINE + lane * 4);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                if (Q > 4) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 512u), "l"(src + 128) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        rs.w1 = __ldg(wc + min(2u, L));
        rs.w2 = __ldg(wc + min(3u, L));
        rs.w3 = __ldg(wc + min(4u, L));
    }
    else if constexpr (MODE == 1)
    {
        /* the streamed lines of rows 1 and 2 go into the two stages; from here on every row refills the
         * stage it has just consumed with the lines of the row two ahead */
        constexpr int LINE = 32 * (Q <= 4 ? 4 : 8);
        const uint32_t wa = __ldg(wc + 1), wb = rs.w1;
        __syncwarp();
        if (lane == 0)
        {
#pragma unroll
            for (int k = 0; k < 2; ++k)
            {
                const uint32_t st = (tc.g + k) & 1u, w = k ? wb : wa;
                mbar_expect_tx(tc.bar + st, 2 * LINE * 4);
                tma_load_1d(tc.ring + (st * 2 + 0) * LINE, emis_lane + (size_t)(84u + (w & 255u)) * LINE, LINE * 4, tc.bar + st);
                tma_load_1d(tc.ring + (st * 2 + 1) * LINE, emis_lane + (size_t)(340u + (w & 1023u)) * LINE, LINE * 4, tc.bar + st);
            }
        }
        /* row j refills with row j+2: its window must be in w2 when row j runs */
        rs.w1 = __ldg(wc + min(2u, L)); /* window of row 2 (addresses of row 2's short lines) */
        rs.w2 = __ldg(wc + min(3u, L)); /* window of row 3: refilled by row 1 */
        rs.w3 = __ldg(wc + min(4u, L));
    }

    float E = NEG_INF, vx = NEG_INF;
    uint32_t j = 1;
#define ROW_ARGS(jj) recs + min((uint32_t)(jj) + 1u, L), wc + min((uint32_t)(jj) + (MODE ? 4u : 3u), L)
    if constexpr (Q <= DCP_NOTAIL_MAXQ)
    {
    /* always whole groups of five rows (one copy of the row code in the instruction cache): rows past L
     * recompute on clamped inputs and are ignored; E and V_X of row L are latched when they pass */
    float E_L = NEG_INF, vx_L = NEG_INF;
#define LATCH(jj)                                                                                              \
    if ((jj) == L) E_L = E, vx_L = vx;
    for (; j <= L; j += 5)
    {
        score_row<Q, 0, MODE>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j), lane, NB, JB, EB, cE, cX, E, vx, tc);
        LATCH(j)
        score_row<Q, 1, MODE>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j + 1), lane, NB, JB, EB, cE, cX, E, vx, tc);
        LATCH(j + 1)
        score_row<Q, 2, MODE>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j + 2), lane, NB, JB, EB, cE, cX, E, vx, tc);
        LATCH(j + 2)
        score_row<Q, 3, MODE>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j + 3), lane, NB, JB, EB, cE, cX, E, vx, tc);
        LATCH(j + 3)
        score_row<Q, 4, MODE>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j + 4), lane, NB, JB, EB, cE, cX, E, vx, tc);
        LATCH(j + 4)
    }
#undef LATCH
    E = E_L, vx = vx_L;
    }
    else
    {
    for (; j + 4 <= L; j += 5)
    {
        score_row<Q, 0, MODE>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j), lane, NB, JB, EB, cE, cX, E, vx, tc);
        score_row<Q, 1, MODE>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j + 1), lane, NB, JB, EB, cE, cX, E, vx, tc);
        score_row<Q, 2, MODE>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j + 2), lane, NB, JB, EB, cE, cX, E, vx, tc);
        score_row<Q, 3, MODE>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j + 3), lane, NB, JB, EB, cE, cX, E, vx, tc);
        score_row<Q, 4, MODE>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j + 4), lane, NB, JB, EB, cE, cX, E, vx, tc);
    }
    if (j <= L) score_row<Q, 0, MODE>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j), lane, NB, JB, EB, cE, cX, E, vx, tc);
    if (j + 1 <= L) score_row<Q, 1, MODE>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j + 1), lane, NB, JB, EB, cE, cX, E, vx, tc);
    if (j + 2 <= L) score_row<Q, 2, MODE>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j + 2), lane, NB, JB, EB, cE, cX, E, vx, tc);
    if (j + 3 <= L) score_row<Q, 3, MODE>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j + 3), lane, NB, JB, EB, cE, cX, E, vx, tc);
    }
#undef ROW_ARGS
    if constexpr (MODE == 2) asm volatile("cp.async.wait_group 0;" ::: "memory"); /* rows past the last were requested too */
    if constexpr (MODE == 1)
    {
        /* rows L+1 and L+2 were requested too (clamped windows): drain them so the stages are free again */
        mbar_wait(tc.bar + (tc.g & 1u), (tc.g >> 1) & 1u);
        tc.g++;
        mbar_wait(tc.bar + (tc.g & 1u), (tc.g >> 1) & 1u);
        tc.g++;
        __syncwarp();
    }
    /* T[L] = max(E[L] + (EC+CT), V_C[L] + CT); V_C lives in lane 2 */
    float vC = __shfl_sync(FULL, vx, 2);
    return fmaxf(E + ET, vC + CT);
}

template <int Q, int MODE>
__global__ void __launch_bounds__(score_warps(Q) * 32, 1)
k_score(const float *__restrict__ emis, const float *__restrict__ trans, const ProfMeta *__restrict__ metas,
        const uint32_t *__restrict__ class_profs, uint32_t n_class_profs, const SeqMeta *__restrict__ seqs,
        uint32_t nseq, uint64_t total_recs, const RowRec *__restrict__ rows, const uint16_t *__restrict__ wcodes,
        const float *__restrict__ spec, float *__restrict__ alt_out, uint32_t nprof,
        unsigned long long *__restrict__ counter, uint32_t seq_tile)
{
    const int lane = threadIdx.x & 31;
    TmaCtx tc = {nullptr, nullptr, 0};
    if constexpr (MODE != 0)
    {
        extern __shared__ __align__(128) unsigned char smem_raw[];
        constexpr int LINE = 32 * (Q <= 4 ? 4 : 8);
        const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        tc.ring = reinterpret_cast<float *>(smem_raw) + (size_t)warp * 4 * LINE;
        tc.bar = reinterpret_cast<uint64_t *>(smem_raw + (size_t)nwarps * 4 * LINE * sizeof(float)) + warp * 2;
        if (MODE == 1 && lane == 0)
        {
            mbar_init(tc.bar, 1);
            mbar_init(tc.bar + 1, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
    }
    /*
     * Work items in tile order: (sequence tile, profile, chunk of kSeqChunk sequences).  All warps
     * of the GPU walk the items in order, so at any time they share a handful of profiles (their
     * emission lines are hot in L1/L2) and one tile of sequences (seq_tile sequences' row records,
     * sized by the host to stay in L2 while every profile passes over them).
     */
    const uint32_t nchunks = (nseq + kSeqChunk - 1) / kSeqChunk;
    const uint32_t tile_chunks = seq_tile / kSeqChunk;
    const unsigned long long n_items = (unsigned long long)n_class_profs * nchunks;
    for (;;)
    {
        unsigned long long item = 0;
        if (lane == 0) item = atomicAdd(counter, 1ULL);
        item = __shfl_sync(FULL, item, 0);
        if (item >= n_items) break;
        /* item -> (tile, profile, chunk in tile); the last tile may be short */
        const unsigned long long per_full_tile = (unsigned long long)tile_chunks * n_class_profs;
        const uint32_t tile = (uint32_t)(item / per_full_tile);
        const unsigned long long in_tile = item - (unsigned long long)tile * per_full_tile;
        const uint32_t chunks_here = min(tile_chunks, nchunks - tile * tile_chunks);
        uint32_t pi = (uint32_t)(in_tile / chunks_here);
        uint32_t ci = tile * tile_chunks + (uint32_t)(in_tile % chunks_here);
        uint32_t prof = class_profs[pi];
        ProfMeta pm = metas[prof];
        NodeParams<Q> p;
        load_params<Q>(p, trans + pm.trans_off, 32 * Q, lane * Q);
        const float *emis_lane = emis + pm.emis_off + lane * emis_lane_stride(1, Q);
        const RowRec *rows_t = rows + (size_t)pm.null_id * total_recs;
        uint32_t s_end = min(nseq, (ci + 1) * kSeqChunk);
        for (uint32_t s = ci * kSeqChunk; s < s_end; ++s)
        {
            SeqMeta sm = seqs[s];
            float T = score_pair<Q, MODE>(p, emis_lane, rows_t + sm.rec_off, wcodes + sm.rec_off, sm.len,
                                         spec + (size_t)s * 16, lane, tc);
            if (lane == 0) alt_out[(size_t)s * nprof + prof] = T;
        }
    }
}


/* ----------------------------------------------------------------------------------------- */
/* profiles of up to 128 nodes: TWO (sequence, profile) pairs per warp, 16 lanes each          */
/* ----------------------------------------------------------------------------------------- */
/*
 * A warp's row costs ~78 instructions whatever it computes (window codes and addresses, row-record loads, shuffles,
 * the E reduction, N/J/C/B) plus ~30 per node and lane; a 50-node profile in a whole warp pays the fixed part for 2
 * nodes per lane.  Here the two halves of a warp run two sequences against the same profile in lock step: the same
 * instruction stream as k_score<Q> -- shuffles and the E reduction confined to 16 lanes, row records, window codes,
 * lengths and length-dependent specials per half -- so the fixed part is shared by two pairs and the padded width of a
 * pair is 16 Q instead of 32 Q'.  Bit-identical arithmetic (the lane layout of a pair is the same consecutive-nodes
 * layout, the D chain the same exact lazy propagation over 16 lanes).
 */
/* maximum over the 16 lanes of this lane's half.  redux.sync with a per-half member mask is compiled into a divergent
 * loop over the distinct masks (REDUX + branches: the half-warp kernels ran 1.5x slower per row with it); two
 * whole-warp reductions with the other half masked out by -inf are straight-line */
__device__ __forceinline__ float half_max(float x, int half)
{
    const float lo = warp_max(half ? NEG_INF : x), hi = warp_max(half ? x : NEG_INF);
    return half ? hi : lo;
}

template <int Q, int R, class Tap = NoTap>
__device__ __forceinline__ void score_row_h(float (&tm)[5][Q], float (&ti)[5][Q], float (&tx)[5],
                                            const NodeParams<Q> &p, RowState<Q> &rs,
                                            const float *__restrict__ emis_lane,
                                            const RowRec *__restrict__ rec_next,
                                            const uint16_t *__restrict__ w_next2, int hl, int half, float NB,
                                            float JB, float EB, float cE, float cX, float &E_out, float &vx_out,
                                            Tap *tap = nullptr)
{
    constexpr int S1 = (R + 4) % 5, S2 = (R + 3) % 5, S3 = (R + 2) % 5, S4 = (R + 1) % 5, S5 = R;
    constexpr int QP = Q <= 4 ? 4 : 8, ROW = 16 * QP, HOFF = 64;

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
    load_emis_part<Q, 3, 5, ROW, HOFF, emis256(0, Q)>(rs.em, emis_lane, code);
    rs.w1 = rs.w2;
    rs.w2 = __ldg(w_next2);
    load_row_insert(rec_next, rs.eI);
    if (hl < 3) load_row_special(rec_next, rs.eN);

    float eloc = vm[0];
#pragma unroll
    for (int i = 1; i < Q; ++i) eloc = fmaxf(eloc, vm[i]);
    const float E = half_max(eloc, half);

    float vm_prev = __shfl_up_sync(FULL, vm[Q - 1], 1, 16);
    if (hl == 0) vm_prev = NEG_INF;

    float d[Q];
    d[0] = vm_prev + p.MD[0];
#pragma unroll
    for (int i = 1; i < Q; ++i) d[i] = fmaxf(vm[i - 1] + p.MD[i], d[i - 1] + p.DD[i]);
    float old = d[Q - 1];
    float din = __shfl_up_sync(FULL, old, 1, 16);
    if (hl == 0) din = NEG_INF;
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

    float vi_prev = __shfl_up_sync(FULL, vi[Q - 1], 1, 16);
    if (hl == 0) vi_prev = NEG_INF;

    load_emis_part<Q, 0, 3, ROW, HOFF, emis256(0, Q)>(rs.em, emis_lane, code);

    const float vN = __shfl_sync(FULL, vx, 0, 16);
    const float vJ = __shfl_sync(FULL, vx, 1, 16);
    const float B = max3(vN + NB, vJ + JB, E + EB);
    tx[R] = fmaxf(E + cE, vx + cX);

#pragma unroll
    for (int i = 0; i < Q; ++i)
    {
        const float pm = i == 0 ? vm_prev : vm[i - 1];
        const float pi = i == 0 ? vi_prev : vi[i - 1];
        tm[R][i] = max3(B + p.ent[i], pm + p.MM[i], pi + p.IM[i]);
        ti[R][i] = fmaxf(vm[i] + p.MI[i], vi[i] + p.II[i]);
    }
    while (more)
    {
        old = d[Q - 1];
        din = __shfl_up_sync(FULL, old, 1, 16);
        if (hl == 0) din = NEG_INF;
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
#pragma unroll
    for (int i = 0; i < Q; ++i)
    {
        const float pd = i == 0 ? din : d[i - 1];
        tm[R][i] = fmaxf(tm[R][i], pd + p.DM[i]);
    }
    tap_row(tap, d, B);
    E_out = E;
    vx_out = vx;
}

/* this lane's half: recs / wc / L / sp of its own sequence; Lmax = the longer of the warp's two sequences */
template <int Q>
__device__ __forceinline__ float score_pair_h(const NodeParams<Q> &p, const float *__restrict__ emis_lane,
                                              const RowRec *__restrict__ recs, const uint16_t *__restrict__ wc,
                                              uint32_t L, uint32_t Lmax, const float *__restrict__ sp, int hl,
                                              int half)
{
    constexpr int QP = Q <= 4 ? 4 : 8;
    const float NN = sp[0], CC = sp[1], JJ = sp[2], NB = sp[3], CT = sp[4], JB = sp[5];
    const float ET = sp[9], ECC = sp[10], EB = sp[11], EJJ = sp[12];
    const float cE = hl == 0 ? NEG_INF : (hl == 1 ? EJJ : ECC);
    const float cX = hl == 0 ? NN : (hl == 1 ? JJ : CC);

    float tm[5][Q], ti[5][Q], tx[5];
#pragma unroll
    for (int s = 0; s < 5; ++s)
    {
        tx[s] = NEG_INF;
#pragma unroll
        for (int i = 0; i < Q; ++i) tm[s][i] = NEG_INF, ti[s][i] = NEG_INF;
    }
#pragma unroll
    for (int i = 0; i < Q; ++i) tm[4][i] = NB + p.ent[i];
    tx[4] = hl == 0 ? NN : NEG_INF;

    RowState<Q> rs;
#pragma unroll
    for (int l = 0; l < 5; ++l) rs.eN[l] = NEG_INF;
    {
        uint32_t code[5];
        codes_of(__ldg(wc + 1), code);
        load_emis<Q, 16 * QP, 64, emis256(0, Q)>(rs.em, emis_lane, code);
    }
    load_row_insert(recs + 1, rs.eI);
    if (hl < 3) load_row_special(recs + 1, rs.eN);
    rs.w1 = __ldg(wc + min(2u, L));
    rs.w2 = __ldg(wc + min(3u, L));
    rs.w3 = 0;

    /* whole groups of five rows up to the longer sequence; rows past this half's own L recompute on clamped inputs
     * and are ignored, E and V_X of row L are latched when they pass */
    float E = NEG_INF, vx = NEG_INF, E_L = NEG_INF, vx_L = NEG_INF;
#define ROW_ARGS_H(jj) recs + min((uint32_t)(jj) + 1u, L), wc + min((uint32_t)(jj) + 3u, L)
#define LATCH_H(jj) /* selects, not a branch: L differs between the halves */                                   \
    {                                                                                                          \
        const bool at = (jj) == L;                                                                             \
        E_L = at ? E : E_L, vx_L = at ? vx : vx_L;                                                             \
    }
    for (uint32_t j = 1; j <= Lmax; j += 5)
    {
        score_row_h<Q, 0>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS_H(j), hl, half, NB, JB, EB, cE, cX, E, vx);
        LATCH_H(j)
        score_row_h<Q, 1>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS_H(j + 1), hl, half, NB, JB, EB, cE, cX, E, vx);
        LATCH_H(j + 1)
        score_row_h<Q, 2>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS_H(j + 2), hl, half, NB, JB, EB, cE, cX, E, vx);
        LATCH_H(j + 2)
        score_row_h<Q, 3>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS_H(j + 3), hl, half, NB, JB, EB, cE, cX, E, vx);
        LATCH_H(j + 3)
        score_row_h<Q, 4>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS_H(j + 4), hl, half, NB, JB, EB, cE, cX, E, vx);
        LATCH_H(j + 4)
    }
#undef ROW_ARGS_H
#undef LATCH_H
    const float vC = __shfl_sync(FULL, vx_L, 2, 16);
    return fmaxf(E_L + ET, vC + CT);
}

template <int Q>
__global__ void __launch_bounds__(score_warps(Q) * 32, 1)
k_score_h(const float *__restrict__ emis, const float *__restrict__ trans, const ProfMeta *__restrict__ metas,
          const uint32_t *__restrict__ class_profs, uint32_t n_class_profs, const SeqMeta *__restrict__ seqs,
          uint32_t nseq, uint64_t total_recs, const RowRec *__restrict__ rows, const uint16_t *__restrict__ wcodes,
          const float *__restrict__ spec, float *__restrict__ alt_out, uint32_t nprof,
          unsigned long long *__restrict__ counter, uint32_t seq_tile)
{
    const int lane = threadIdx.x & 31, hl = lane & 15, half = lane >> 4;
    /* work items as in k_score: (sequence tile, profile, chunk of kSeqChunk sequences); a chunk is taken two
     * sequences at a time, one per half-warp */
    const uint32_t nchunks = (nseq + kSeqChunk - 1) / kSeqChunk;
    const uint32_t tile_chunks = seq_tile / kSeqChunk;
    const unsigned long long n_items = (unsigned long long)n_class_profs * nchunks;
    for (;;)
    {
        unsigned long long item = 0;
        if (lane == 0) item = atomicAdd(counter, 1ULL);
        item = __shfl_sync(FULL, item, 0);
        if (item >= n_items) break;
        const unsigned long long per_full_tile = (unsigned long long)tile_chunks * n_class_profs;
        const uint32_t tile = (uint32_t)(item / per_full_tile);
        const unsigned long long in_tile = item - (unsigned long long)tile * per_full_tile;
        const uint32_t chunks_here = min(tile_chunks, nchunks - tile * tile_chunks);
        const uint32_t pi = (uint32_t)(in_tile / chunks_here);
        const uint32_t ci = tile * tile_chunks + (uint32_t)(in_tile % chunks_here);
        const uint32_t prof = class_profs[pi];
        const ProfMeta pm = metas[prof];
        NodeParams<Q> p;
        load_params<Q>(p, trans + pm.trans_off, 16 * Q, hl * Q);
        const float *emis_lane = emis + pm.emis_off + hl * emis_lane_stride(0, Q);
        const RowRec *rows_t = rows + (size_t)pm.null_id * total_recs;
        const uint32_t s_end = min(nseq, (ci + 1) * kSeqChunk);
        for (uint32_t s0 = ci * kSeqChunk; s0 < s_end; s0 += 2)
        {
            const uint32_t s = min(s0 + (uint32_t)half, s_end - 1); /* odd tail: the upper half repeats the last one */
            const SeqMeta sm = seqs[s];
            /* the row loop's trip count, from warp-uniform loads (a shuffle would make the loop look divergent to the
             * compiler, which then guards every shuffle in it with BRA.DIV) */
            const uint32_t Lmax = max(seqs[s0].len, seqs[min(s0 + 1, s_end - 1)].len);
            const float T = score_pair_h<Q>(p, emis_lane, rows_t + sm.rec_off, wcodes + sm.rec_off, sm.len, Lmax,
                                            spec + (size_t)s * 16, hl, half);
            if (hl == 0 && s0 + (uint32_t)half < s_end) alt_out[(size_t)s * nprof + prof] = T;
        }
    }
}

} // namespace
#endif
