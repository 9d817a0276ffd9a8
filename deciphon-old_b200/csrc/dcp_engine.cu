/*
 * dcp_engine.cu -- the sm_100a scan engine behind include/dcpgpu.h (Part 2).
 *
 * Replaces, for a whole batch of (sequence, profile) pairs, what the reference does per pair
 * in thread_run (src/server/scan_thread.c:86-135): protein_profile_setup, imm_dp_viterbi on the
 * null and alt DPs, xmath_lrt + threshold and the traceback that feeds prod_fwrite.
 *
 * Kernels (all fp32 max-plus, no tensor cores):
 *   k_rows       per (null table, sequence row): window codes + N/J/C/R and insert emissions
 *   k_null       1-state null Viterbi per (sequence, null table)
 *   k_score<Q>   alt Viterbi score, one warp per pair, Q consecutive core nodes per lane,
 *                last five rows of Tin_M/Tin_I in registers, D chain resolved exactly by lazy
 *                warp-shuffle propagation, E by redux.sync.max.f32 (CREDUX)
 *   k_lrt        LRT filter -> hit flags / hit list
 *   k_trace<Q>   same DP for hits only, with imm's (transition, source length) first-max
 *                tie-break and packed 11-bit backpointers per cell
 *   k_walk       backpointers -> (state_id, seqlen) steps
 *
 * Emission tables live in HBM transposed: [profile][code 0..1363][lane][QP] so that the 32 lanes
 * of a warp read one contiguous 32*QP*4-byte line per (row, length) with LDG.128.
 */
#ifndef DCP_NOTAIL_MAXQ
#define DCP_NOTAIL_MAXQ 6
#endif
#include "dcp_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>


namespace
{

/* ----------------------------------------------------------------------------------------- */
/* prep kernels                                                                              */
/* ----------------------------------------------------------------------------------------- */
__global__ void k_rows(const uint8_t *__restrict__ bases, const SeqMeta *__restrict__ seqs, uint32_t nseq,
                       const float *__restrict__ null_tabs, const float *__restrict__ ins_tab,
                       uint32_t n_null, uint64_t total_recs, RowRec *__restrict__ rows,
                       uint16_t *__restrict__ wcodes)
{
    uint32_t s = blockIdx.x;
    if (s >= nseq) return;
    SeqMeta sm = seqs[s];
    const uint8_t *b = bases + sm.row_off;
    for (uint32_t j = threadIdx.x; j <= sm.len; j += blockDim.x)
    {
        /* seq[j-l] is the l-th least significant base pair of the window */
        uint32_t w = 0;
#pragma unroll
        for (int l = 1; l <= 5; ++l)
            if (j >= (uint32_t)l) w |= (uint32_t)b[j - l] << (2 * (l - 1));
        wcodes[sm.rec_off + j] = (uint16_t)w;
        uint32_t code[5];
        codes_of(w, code);
        for (uint32_t t = 0; t < n_null; ++t)
        {
            RowRec r;
#pragma unroll
            for (int l = 0; l < 5; ++l)
            {
                r.eN[l] = null_tabs[(size_t)t * kTab + code[l]];
                r.eI[l] = ins_tab[code[l]];
            }
            r.pad0[0] = r.pad0[1] = r.pad0[2] = r.pad1[0] = r.pad1[1] = r.pad1[2] = 0.0f;
            rows[(size_t)t * total_recs + sm.rec_off + j] = r;
        }
    }
}

/* imm_dp_viterbi on the null dp (scan_thread.c:115): V_R[j] = max_l Tin_R[j-l] + e_R(seq[j-l:j]) */
__global__ void k_null(const SeqMeta *__restrict__ seqs, uint32_t nseq, uint32_t n_null, uint64_t total_recs,
                       const RowRec *__restrict__ rows, const float *__restrict__ spec,
                       float *__restrict__ null_out)
{
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= nseq * n_null) return;
    uint32_t s = tid / n_null, t = tid % n_null;
    SeqMeta sm = seqs[s];
    const float RR = spec[(size_t)s * 16 + 6];
    const RowRec *r = rows + (size_t)t * total_recs + sm.rec_off;
    /* ring of Tin_R for the last five rows; Tin_R[0] = start lprob 0 */
    float tin[5] = {NEG_INF, NEG_INF, NEG_INF, NEG_INF, 0.0f};
    float v = NEG_INF;
    for (uint32_t j = 1; j <= sm.len; ++j)
    {
        float e[5];
        load_row_special(r + j, e);
        v = fmaxf(fmaxf(fmaxf(tin[4] + e[0], tin[3] + e[1]), fmaxf(tin[2] + e[2], tin[1] + e[3])), tin[0] + e[4]);
        tin[0] = tin[1], tin[1] = tin[2], tin[2] = tin[3], tin[3] = tin[4];
        tin[4] = v + RR;
    }
    null_out[(size_t)s * n_null + t] = v;
}

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

template <int Q, int R, bool TMA>
__device__ __forceinline__ void score_row(float (&tm)[5][Q], float (&ti)[5][Q], float (&tx)[5],
                                          const NodeParams<Q> &p, RowState<Q> &rs,
                                          const float *__restrict__ emis_lane,
                                          const RowRec *__restrict__ rec_next,
                                          const uint16_t *__restrict__ w_next2, int lane, float NB, float JB,
                                          float EB, float cE, float cX, float &E_out, float &vx_out, TmaCtx &tc)
{
    constexpr int S1 = (R + 4) % 5, S2 = (R + 3) % 5, S3 = (R + 2) % 5, S4 = (R + 1) % 5, S5 = R;
    constexpr int QP = Q <= 4 ? 4 : 8, LINE = 32 * QP;

    if constexpr (TMA)
    {
        /* this row's 4- and 5-nt lines were bulk-copied into the stage two rows ago */
        const uint32_t st = tc.g & 1u;
        mbar_wait(tc.bar + st, (tc.g >> 1) & 1u);
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
    if constexpr (TMA)
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
        load_emis_part<Q, 3, 5>(rs.em, emis_lane, code);
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
    load_emis_part<Q, 0, 3>(rs.em, emis_lane, code);

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
    E_out = E;
    vx_out = vx;
}

/* recs / wc = record and window of row 0 of this sequence (L+1 of each) */
template <int Q, bool TMA>
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
        load_emis<Q>(rs.em, emis_lane, code);
    }
    load_row_insert(recs + 1, rs.eI);
    if (lane < 3) load_row_special(recs + 1, rs.eN);
    rs.w1 = __ldg(wc + min(2u, L));
    rs.w2 = __ldg(wc + min(3u, L));
    rs.w3 = 0;
    if constexpr (TMA)
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
#define ROW_ARGS(jj) recs + min((uint32_t)(jj) + 1u, L), wc + min((uint32_t)(jj) + (TMA ? 4u : 3u), L)
    if constexpr (Q <= DCP_NOTAIL_MAXQ)
    {
    /* always whole groups of five rows (one copy of the row code in the instruction cache): rows past L
     * recompute on clamped inputs and are ignored; E and V_X of row L are latched when they pass */
    float E_L = NEG_INF, vx_L = NEG_INF;
#define LATCH(jj)                                                                                              \
    if ((jj) == L) E_L = E, vx_L = vx;
    for (; j <= L; j += 5)
    {
        score_row<Q, 0, TMA>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j), lane, NB, JB, EB, cE, cX, E, vx, tc);
        LATCH(j)
        score_row<Q, 1, TMA>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j + 1), lane, NB, JB, EB, cE, cX, E, vx, tc);
        LATCH(j + 1)
        score_row<Q, 2, TMA>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j + 2), lane, NB, JB, EB, cE, cX, E, vx, tc);
        LATCH(j + 2)
        score_row<Q, 3, TMA>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j + 3), lane, NB, JB, EB, cE, cX, E, vx, tc);
        LATCH(j + 3)
        score_row<Q, 4, TMA>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j + 4), lane, NB, JB, EB, cE, cX, E, vx, tc);
        LATCH(j + 4)
    }
#undef LATCH
    E = E_L, vx = vx_L;
    }
    else
    {
    for (; j + 4 <= L; j += 5)
    {
        score_row<Q, 0, TMA>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j), lane, NB, JB, EB, cE, cX, E, vx, tc);
        score_row<Q, 1, TMA>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j + 1), lane, NB, JB, EB, cE, cX, E, vx, tc);
        score_row<Q, 2, TMA>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j + 2), lane, NB, JB, EB, cE, cX, E, vx, tc);
        score_row<Q, 3, TMA>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j + 3), lane, NB, JB, EB, cE, cX, E, vx, tc);
        score_row<Q, 4, TMA>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j + 4), lane, NB, JB, EB, cE, cX, E, vx, tc);
    }
    if (j <= L) score_row<Q, 0, TMA>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j), lane, NB, JB, EB, cE, cX, E, vx, tc);
    if (j + 1 <= L) score_row<Q, 1, TMA>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j + 1), lane, NB, JB, EB, cE, cX, E, vx, tc);
    if (j + 2 <= L) score_row<Q, 2, TMA>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j + 2), lane, NB, JB, EB, cE, cX, E, vx, tc);
    if (j + 3 <= L) score_row<Q, 3, TMA>(tm, ti, tx, p, rs, emis_lane, ROW_ARGS(j + 3), lane, NB, JB, EB, cE, cX, E, vx, tc);
    }
#undef ROW_ARGS
    if constexpr (TMA)
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

template <int Q, bool TMA>
__global__ void __launch_bounds__(score_warps(Q) * 32, 1)
k_score(const float *__restrict__ emis, const float *__restrict__ trans, const ProfMeta *__restrict__ metas,
        const uint32_t *__restrict__ class_profs, uint32_t n_class_profs, const SeqMeta *__restrict__ seqs,
        uint32_t nseq, uint64_t total_recs, const RowRec *__restrict__ rows, const uint16_t *__restrict__ wcodes,
        const float *__restrict__ spec, float *__restrict__ alt_out, uint32_t nprof,
        unsigned long long *__restrict__ counter, uint32_t seq_tile)
{
    const int lane = threadIdx.x & 31;
    TmaCtx tc = {nullptr, nullptr, 0};
    if constexpr (TMA)
    {
        extern __shared__ __align__(128) unsigned char smem_raw[];
        constexpr int LINE = 32 * (Q <= 4 ? 4 : 8);
        const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        tc.ring = reinterpret_cast<float *>(smem_raw) + (size_t)warp * 4 * LINE;
        tc.bar = reinterpret_cast<uint64_t *>(smem_raw + (size_t)nwarps * 4 * LINE * sizeof(float)) + warp * 2;
        if (lane == 0)
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
        const float *emis_lane = emis + pm.emis_off + lane * 4;
        const RowRec *rows_t = rows + (size_t)pm.null_id * total_recs;
        uint32_t s_end = min(nseq, (ci + 1) * kSeqChunk);
        for (uint32_t s = ci * kSeqChunk; s < s_end; ++s)
        {
            SeqMeta sm = seqs[s];
            float T = score_pair<Q, TMA>(p, emis_lane, rows_t + sm.rec_off, wcodes + sm.rec_off, sm.len,
                                         spec + (size_t)s * 16, lane, tc);
            if (lane == 0) alt_out[(size_t)s * nprof + prof] = T;
        }
    }
}

/* ----------------------------------------------------------------------------------------- */
/* alt Viterbi, score pass, profiles of 257..4096 nodes: a group of warps per pair            */
/* ----------------------------------------------------------------------------------------- */
/*
 * Same recurrence, same fp32 operation order and the same lane layout as k_score<8>; node
 * k-1 = gwarp * 256 + lane * 8 + sub.  257..2048 nodes: W warps of one block (CL = 1);
 * 2049..4096 nodes: W warps in each block of a 2-block cluster (CL = 2, 255 registers x 16 warps do
 * not fit one SM), exchanging through distributed shared memory.  What a single warp exchanges with
 * shuffles is exchanged between warps through shared memory, two group barriers per row:
 *   A   V_M / V_I / D of each warp's last node (D from the warp-local chain), per-warp max of V_M (-> E),
 *       V_N / V_J / V_C of warp 0
 *   C   group-wide OR: did any warp's last D rise when the left neighbour's values came in?
 *       (if so publish the new D, barrier B, and repeat -- exact lazy propagation, as inside a warp)
 * Measured alternatives that were not faster: keeping 8 warps per SM with 5 or 6 nodes per lane
 * (M = 600: 252 vs 267 GCUPS) -- the barriers, not the occupancy, bound these kernels.
 */
struct MwShared
{
    float vm_last[2][kMaxGroupWarps], vi_last[2][kMaxGroupWarps], e_warp[2][kMaxGroupWarps];
    float d_last[2][kMaxGroupWarps];
    float v_spec[2][4]; /* V_N, V_J, V_C of the row */
    int flag[2][2];
    unsigned long long item;
    alignas(16) float xch[2][kMaxGroupWarps][8]; /* 2-block groups: Group::exchange buffers ... */
    unsigned long long xbar[2];                  /* ... and their mbarriers */
};

template <int W, int CL, int R, int Q>
__device__ __forceinline__ void mw_row(float (&tm)[5][Q], float (&ti)[5][Q], float (&tx)[5],
                                       const NodeParams<Q> &p, RowState<Q> &rs,
                                       const float *__restrict__ emis_lane, const RowRec *__restrict__ rec_next,
                                       const uint16_t *__restrict__ w_next2, int gw, int lane, int par,
                                       Group<CL, MwShared> &grp, float NB, float JB, float EB, float cE, float cX,
                                       float &E_out, float &vC_out)
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
    /* N, J, C live in lanes 0..2 of the group's first warp */
    float vx = fmaxf(max3(tx[S1] + rs.eN[0], tx[S2] + rs.eN[1], tx[S3] + rs.eN[2]),
                     fmaxf(tx[S4] + rs.eN[3], tx[S5] + rs.eN[4]));

    /* next row's loads (same software pipeline as the single-warp kernel) */
    uint32_t code[5];
    codes_of(rs.w1, code);
    load_emis_part<Q, 3, 5, ROW>(rs.em, emis_lane, code);
    load_row_insert(rec_next, rs.eI);
    if (gw == 0 && lane < 3) load_row_special(rec_next, rs.eN);
    rs.w1 = rs.w2;
    rs.w2 = __ldg(w_next2);

    float eloc = vm[0];
#pragma unroll
    for (int i = 1; i < Q; ++i) eloc = fmaxf(eloc, vm[i]);
    float ew = warp_max(eloc);
    float vm_prev = __shfl_up_sync(FULL, vm[Q - 1], 1);
    float vi_prev = __shfl_up_sync(FULL, vi[Q - 1], 1);
    load_emis_part<Q, 0, 3, ROW>(rs.em, emis_lane, code);

    /* D chain inside the warp, nothing from the warp to the left yet: the warp's first node starts at -inf
     * (its M->D and D->D sources both live in the left warp and arrive together after barrier A) */
    float d[Q];
    d[0] = lane == 0 ? NEG_INF : vm_prev + p.MD[0];
#pragma unroll
    for (int i = 1; i < Q; ++i) d[i] = fmaxf(vm[i - 1] + p.MD[i], d[i - 1] + p.DD[i]);
    float din;
    for (;;)
    {
        float old = d[Q - 1];
        din = __shfl_up_sync(FULL, old, 1);
        float x = lane == 0 ? NEG_INF : din + p.DD[0];
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
#ifndef DCP_CLUSTER_XCH
#define DCP_CLUSTER_XCH 1
#endif
    float E, vN, vJ, vC;
    if constexpr (CL == 2 && DCP_CLUSTER_XCH)
    {
        /* A: boundary values, per-warp maxima, the local D chains' ends and the specials, one exchange */
        const float xN = __shfl_sync(FULL, vx, 0), xJ = __shfl_sync(FULL, vx, 1), xC = __shfl_sync(FULL, vx, 2);
        const float pay_a[8] = {vm[Q - 1], vi[Q - 1], ew, d[Q - 1], xN, xJ, xC, 0.0f};
        int s = grp.exchange(gw, lane, pay_a);
        {
            const float(*x)[8] = sh.xch[s];
            if (lane == 0)
            {
                vm_prev = gw ? x[gw - 1][0] : NEG_INF;
                vi_prev = gw ? x[gw - 1][1] : NEG_INF;
            }
            E = x[0][2];
#pragma unroll
            for (int w = 1; w < TW; ++w) E = fmaxf(E, x[w][2]);
            vN = x[0][4], vJ = x[0][5], vC = x[0][6];
        }
        float din0 = gw ? sh.xch[s][gw - 1][3] : NEG_INF;
        /* carries between warps, lazily: every round ends with an exchange of (did my last D rise, my last D) */
        for (;;)
        {
            const float before = __shfl_sync(FULL, d[Q - 1], 31);
            for (;;)
            {
                float old = d[Q - 1];
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
            const float pay_c[8] = {d[Q - 1] > before ? 1.0f : 0.0f, d[Q - 1], 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
            s = grp.exchange(gw, lane, pay_c);
            float rose = sh.xch[s][0][0];
#pragma unroll
            for (int w = 1; w < TW; ++w) rose = fmaxf(rose, sh.xch[s][w][0]);
            if (rose == 0.0f) break; /* C */
            din0 = gw ? sh.xch[s][gw - 1][1] : NEG_INF;
        }
    }
    else
    {
        if (lane == 31)
        {
            GRP_PUT(grp, vm_last[par][gw], vm[Q - 1]);
            GRP_PUT(grp, vi_last[par][gw], vi[Q - 1]);
            GRP_PUT(grp, e_warp[par][gw], ew);
            GRP_PUT(grp, d_last[0][gw], d[Q - 1]);
        }
        if (gw == 0 && lane < 3) GRP_PUT(grp, v_spec[par][lane], vx);
        grp.sync(); /* A: boundary values, per-warp maxima, specials and the local D chains' ends */

        if (lane == 0)
        {
            vm_prev = gw ? sh.vm_last[par][gw - 1] : NEG_INF;
            vi_prev = gw ? sh.vi_last[par][gw - 1] : NEG_INF;
        }
        E = sh.e_warp[par][0];
#pragma unroll
        for (int w = 1; w < TW; ++w) E = fmaxf(E, sh.e_warp[par][w]);
        vN = sh.v_spec[par][0], vJ = sh.v_spec[par][1], vC = sh.v_spec[par][2];

        /* carries between warps: D of the warp's first node = max(V_M(left) + MD, D(left) + DD), then lazily on */
        float din0 = NEG_INF; /* D of the last node of the warp to the left */
        for (int round = 0;; ++round)
        {
            const int b = round & 1;
            if (round > 0)
            {
                if (lane == 31) GRP_PUT(grp, d_last[b][gw], d[Q - 1]);
                grp.sync(); /* B */
            }
            din0 = gw ? sh.d_last[b][gw - 1] : NEG_INF;
            const float before = __shfl_sync(FULL, d[Q - 1], 31);
            for (;;)
            {
                float old = d[Q - 1];
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
            /* two warps: the left warp has no carry-in, so the D it published at A was final, and nobody reads
             * the right warp's -- no second round can happen and barrier C is not needed (M = 512: 457 -> 520 GCUPS) */
            if (TW == 2) break;
            const float after = __shfl_sync(FULL, d[Q - 1], 31);
            if (!grp.any(after > before, sh.flag, CL == 2 ? grp.peer->flag : sh.flag, b)) break; /* C */
        }
    }

    float B = max3(vN + NB, vJ + JB, E + EB);
    tx[R] = fmaxf(E + cE, vx + cX);
#pragma unroll
    for (int i = 0; i < Q; ++i)
    {
        float pm = i == 0 ? vm_prev : vm[i - 1];
        float pi = i == 0 ? vi_prev : vi[i - 1];
        float pd = i == 0 ? din : d[i - 1];
        tm[R][i] = fmaxf(fmaxf(B + p.ent[i], pm + p.MM[i]), fmaxf(pi + p.IM[i], pd + p.DM[i]));
        ti[R][i] = fmaxf(vm[i] + p.MI[i], vi[i] + p.II[i]);
    }
    E_out = E;
    vC_out = vC;
}

template <int W, int CL, int Q = 8>
__global__ void __launch_bounds__(W * 32, CL == 2 ? 1 : 8 / W)
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
        const float *emis_lane = emis + pm.emis_off + gw * 256 + lane * 4;
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
            load_emis<Q, ROW>(rs.em, emis_lane, code);
        }
        load_row_insert(recs + 1, rs.eI);
        if (gw == 0 && lane < 3) load_row_special(recs + 1, rs.eN);
        rs.w1 = __ldg(wc + min(2u, L));
        rs.w2 = __ldg(wc + min(3u, L));

        float E = NEG_INF, vC = NEG_INF;
        uint32_t j = 1;
#define MW_ARGS(jj) recs + min((uint32_t)(jj) + 1u, L), wc + min((uint32_t)(jj) + 3u, L), gw, lane, (int)((jj)&1u), grp
        for (; j + 4 <= L; j += 5)
        {
            mw_row<W, CL, 0, Q>(tm, ti, tx, p, rs, emis_lane, MW_ARGS(j), NB, JB, EB, cE, cX, E, vC);
            mw_row<W, CL, 1, Q>(tm, ti, tx, p, rs, emis_lane, MW_ARGS(j + 1), NB, JB, EB, cE, cX, E, vC);
            mw_row<W, CL, 2, Q>(tm, ti, tx, p, rs, emis_lane, MW_ARGS(j + 2), NB, JB, EB, cE, cX, E, vC);
            mw_row<W, CL, 3, Q>(tm, ti, tx, p, rs, emis_lane, MW_ARGS(j + 3), NB, JB, EB, cE, cX, E, vC);
            mw_row<W, CL, 4, Q>(tm, ti, tx, p, rs, emis_lane, MW_ARGS(j + 4), NB, JB, EB, cE, cX, E, vC);
        }
        if (j <= L) mw_row<W, CL, 0, Q>(tm, ti, tx, p, rs, emis_lane, MW_ARGS(j), NB, JB, EB, cE, cX, E, vC);
        if (j + 1 <= L) mw_row<W, CL, 1, Q>(tm, ti, tx, p, rs, emis_lane, MW_ARGS(j + 1), NB, JB, EB, cE, cX, E, vC);
        if (j + 2 <= L) mw_row<W, CL, 2, Q>(tm, ti, tx, p, rs, emis_lane, MW_ARGS(j + 2), NB, JB, EB, cE, cX, E, vC);
        if (j + 3 <= L) mw_row<W, CL, 3, Q>(tm, ti, tx, p, rs, emis_lane, MW_ARGS(j + 3), NB, JB, EB, cE, cX, E, vC);
#undef MW_ARGS
        if (gw == 0 && lane == 0) alt_out[(size_t)s * nprof + prof] = fmaxf(E + ET, vC + CT);
    }
}

/*
 * Match emission tables, host layout [node][code] -> device layout [code][warp][half][lane][4]
 * (node k-1 = warp * 32 Q + lane * Q + sub sits in half sub/4, float sub%4 of its lane; pads are -inf).
 */
__global__ void k_layout(const float *__restrict__ raw, float *__restrict__ out, uint32_t M, uint32_t Q, uint32_t QP,
                         uint32_t W)
{
    const uint32_t ROW = 32 * QP * W;
    const size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= (size_t)kTab * ROW) return;
    const uint32_t code = (uint32_t)(x / ROW), pos = (uint32_t)(x % ROW);
    const uint32_t warp = pos / (32 * QP), r = pos % (32 * QP);
    const uint32_t half = r / 128, lane = (r % 128) / 4, sub = half * 4 + (r % 4);
    float v = NEG_INF;
    if (sub < Q)
    {
        const uint32_t k = warp * 32 * Q + lane * Q + sub;
        if (k < M) v = raw[(size_t)k * kTab + code];
    }
    out[x] = v;
}

/* xmath_lrt + threshold (scan_thread.c:121-123): hit iff finite and not (lrt < threshold) */
__global__ void k_lrt(const float *__restrict__ alt, const float *__restrict__ null_by_tab,
                      const ProfMeta *__restrict__ metas, uint32_t nseq, uint32_t nprof, uint32_t n_null,
                      double thr, uint8_t *__restrict__ hit, unsigned long long *__restrict__ nhits)
{
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = (size_t)nseq * nprof;
    if (idx >= n) return;
    uint32_t s = (uint32_t)(idx / nprof), pr = (uint32_t)(idx % nprof);
    float nl = null_by_tab[(size_t)s * n_null + metas[pr].null_id];
    float lrt = -2.0f * (nl - alt[idx]);
    bool h = isfinite(lrt) && !((double)lrt < thr);
    hit[idx] = h ? 1 : 0;
    if (h) atomicAdd(nhits, 1ULL);
}

__global__ void k_collect(const uint8_t *__restrict__ hit, size_t n, unsigned long long *__restrict__ cursor,
                          unsigned long long *__restrict__ list)
{
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    if (hit[idx]) list[atomicAdd(cursor, 1ULL)] = idx;
}


__global__ void k_gather(const unsigned long long *__restrict__ list, size_t nhits, const float *__restrict__ alt,
                         const float *__restrict__ null_by_tab, const ProfMeta *__restrict__ metas, uint32_t nprof,
                         uint32_t n_null, float *__restrict__ out_alt, float *__restrict__ out_null)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nhits) return;
    unsigned long long idx = list[i];
    uint32_t s = (uint32_t)(idx / nprof), pr = (uint32_t)(idx % nprof);
    out_alt[i] = alt[idx];
    out_null[i] = null_by_tab[(size_t)s * n_null + metas[pr].null_id];
}

template <int Q>
void launch_score(int nblocks, cudaStream_t st, const float *emis, const float *trans, const ProfMeta *metas,
                  const uint32_t *class_profs, uint32_t n_class, const SeqMeta *seqs, uint32_t nseq,
                  uint64_t total_rows, const RowRec *rows, const uint16_t *wcodes, const float *spec, float *alt,
                  uint32_t nprof,
                  unsigned long long *counter, uint32_t seq_tile)
{
    static const bool use_tma = getenv("DCPGPU_TMA") && atoi(getenv("DCPGPU_TMA")) != 0;
    if (use_tma)
    {
        /* experiment: 4/5-nt emission lines through cp.async.bulk + mbarrier into a per-warp shared ring */
        const int warps = score_warps(Q), LINE = 32 * (Q <= 4 ? 4 : 8);
        const size_t smem = (size_t)warps * 4 * LINE * sizeof(float) + (size_t)warps * 2 * sizeof(uint64_t);
        cudaFuncSetAttribute(k_score<Q, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_score<Q, true><<<nblocks, warps * 32, smem, st>>>(emis, trans, metas, class_profs, n_class, seqs, nseq,
                                                            total_rows, rows, wcodes, spec, alt, nprof, counter,
                                                            seq_tile);
        return;
    }
    /* no shared memory: give the whole unified array to L1 (emission lines, row records) */
    cudaFuncSetAttribute(k_score<Q, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
    k_score<Q, false><<<nblocks, score_warps(Q) * 32, 0, st>>>(emis, trans, metas, class_profs, n_class, seqs, nseq,
                                                        total_rows, rows, wcodes, spec, alt, nprof, counter, seq_tile);
}

} // namespace

/* ----------------------------------------------------------------------------------------- */
/* host objects                                                                              */
/* ----------------------------------------------------------------------------------------- */
static protein_profile *profile_clone(protein_profile const *src)
{
    protein_profile *p = (protein_profile *)calloc(1, sizeof *p);
    if (!p) return nullptr;
    *p = *src;
    unsigned n = src->core_size;
    p->consensus = (char *)malloc(n + 1);
    p->match_ndists = (dcp_nuclt_dist *)malloc(n * sizeof *p->match_ndists);
    p->match_emission = (float *)malloc((size_t)n * kTab * sizeof(float));
    p->trans = (protein_trans *)malloc((n + 1) * sizeof *p->trans);
    p->entry = (float *)malloc(n * sizeof(float));
    if (!p->consensus || !p->match_ndists || !p->match_emission || !p->trans || !p->entry)
    {
        protein_profile_del(p);
        return nullptr;
    }
    memcpy(p->consensus, src->consensus, n + 1);
    memcpy(p->match_ndists, src->match_ndists, n * sizeof *p->match_ndists);
    memcpy(p->match_emission, src->match_emission, (size_t)n * kTab * sizeof(float));
    memcpy(p->trans, src->trans, (n + 1) * sizeof *p->trans);
    memcpy(p->entry, src->entry, n * sizeof(float));
    return p;
}

extern "C" enum rc dcpgpu_db_new(struct dcpgpu_db **out, int device)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return dcp_error(RC_EFAIL, "no CUDA device: the scan engine has no CPU fallback");
    if (device < 0 || device >= ndev) return dcp_error(RC_EINVAL, "no such CUDA device");
    CU_TRY(cudaSetDevice(device));
    dcpgpu_db *db = new (std::nothrow) dcpgpu_db;
    if (!db) return dcp_error(RC_ENOMEM, "alloc db");
    db->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) db->sm_count = prop.multiProcessorCount;
    {
        /* keep freed scratch blocks cached in the pool between scans */
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess)
        {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    if (cudaStreamCreateWithFlags(&db->stream, cudaStreamNonBlocking) != cudaSuccess)
    {
        delete db;
        return dcp_error(RC_EFAIL, "cudaStreamCreate failed");
    }
    *out = db;
    return RC_OK;
}

static enum rc db_take(struct dcpgpu_db *db, struct protein_profile *prof, bool owned);

extern "C" enum rc dcpgpu_db_add(struct dcpgpu_db *db, struct protein_profile const *prof)
{
    return db_take(db, const_cast<protein_profile *>(prof), false);
}

/* same as dcpgpu_db_add, but the database takes ownership of `prof` (no copy); used by the press */
enum rc dcp_db_adopt(struct dcpgpu_db *db, struct protein_profile *prof) { return db_take(db, prof, true); }

static enum rc db_take(struct dcpgpu_db *db, struct protein_profile *prof, bool owned)
{
    if (db->committed) return dcp_error(RC_EFAIL, "database already committed");
    if (prof->core_size == 0) return dcp_error(RC_EINVAL, "profile has not been absorbed");
    if (prof->core_size > DCP_PROTEIN_MODEL_CORE_SIZE_MAX)
        return dcp_error(RC_EINVAL, "profile is too long"); /* limits.h:11, protein_profile.c:58 */
    if (db->epsilon >= 0.0f && db->epsilon != prof->cfg.epsilon)
        return dcp_error(RC_EINVAL, "all profiles of a database share one epsilon");
    /* the score pass takes E[j] = max_k V_Mk[j]; that needs delete scores to be log-probabilities */
    for (unsigned i = 0; i <= prof->core_size; ++i)
        if (prof->trans[i].MD > 0.0f || prof->trans[i].DD > 0.0f)
            return dcp_error(RC_EINVAL, "MD/DD transition scores must be <= 0");
    protein_profile *copy = owned ? prof : profile_clone(prof);
    if (!copy) return dcp_error(RC_ENOMEM, "clone profile");
    db->epsilon = prof->cfg.epsilon;
    uint32_t id = UINT32_MAX;
    for (size_t t = 0; t < db->null_tabs.size(); ++t)
        if (!memcmp(db->null_tabs[t].data(), prof->null_emission, kTab * sizeof(float)))
        {
            id = (uint32_t)t;
            break;
        }
    if (id == UINT32_MAX)
    {
        id = (uint32_t)db->null_tabs.size();
        db->null_tabs.emplace_back(prof->null_emission, prof->null_emission + kTab);
    }
    db->profs.push_back(copy);
    db->null_id.push_back(id);
    return RC_OK;
}

#ifndef DCP_W2Q7
#define DCP_W2Q7 1
#endif
/* profile length -> (nodes per lane, warps per pair, kernel class) */
static void kernel_shape(uint32_t M, uint32_t &Q, uint32_t &W, uint32_t &cls)
{
    /* nodes per lane.  193..224 nodes would fit 7 per lane, but k_score<7> does not fit the
     * register file without spilling in its straight-line form; 8 per lane (7 idle lanes) measured
     * faster: 515 vs 489 GCUPS at M = 200.  Above 256 nodes W warps share one pair, 8 nodes per lane. */
    Q = (M + 31) / 32, W = 1;
    if (Q == 7) Q = 8;
    cls = Q;
    if (Q > 8)
    {
        Q = 8, W = (M + 255) / 256;
        if (W > kMaxW) W = (W + 1) / 2 * 2; /* two blocks of W/2 warps (cluster) */
        cls = kMaxQ + W;
        /* two / three warps with 6 or 7 nodes per lane instead of 8: fewer padded nodes for 257..448 and 513..672
         * (M = 350: 357 -> 418 GCUPS; 5 nodes per lane for 257..320 gains only 2 %, 6 per lane is better there too) */
        if (M <= 384) Q = 6, cls = kClsW2Q6;
        else if (DCP_W2Q7 && M <= 448) Q = 7, cls = kClsW2Q7;
        else if (M > 512 && M <= 576) Q = 6, cls = kClsW3Q6;
        else if (M > 576 && M <= 672) Q = 7, cls = kClsW3Q7;
    }
}

extern "C" enum rc dcpgpu_kernel_shape(unsigned core_size, unsigned *warps, unsigned *nodes_per_lane, unsigned *blocks)
{
    if (core_size == 0 || core_size > DCP_PROTEIN_MODEL_CORE_SIZE_MAX)
        return dcp_error(RC_EINVAL, "core size out of range");
    uint32_t Q, W, cls;
    kernel_shape(core_size, Q, W, cls);
    if (warps) *warps = W;
    if (nodes_per_lane) *nodes_per_lane = Q;
    if (blocks) *blocks = W > (uint32_t)kMaxW ? 2 : 1;
    return RC_OK;
}

extern "C" enum rc dcpgpu_db_commit(struct dcpgpu_db *db)
{
    if (db->committed) return dcp_error(RC_EFAIL, "database already committed");
    if (db->profs.empty()) return dcp_error(RC_EINVAL, "database is empty");
    CU_TRY(cudaSetDevice(db->device));
    size_t nprof = db->profs.size();
    db->metas.resize(nprof);
    uint64_t emis_floats = 0, trans_floats = 0;
    for (size_t i = 0; i < nprof; ++i)
    {
        uint32_t M = db->profs[i]->core_size;
        uint32_t Q, W, cls;
        kernel_shape(M, Q, W, cls);
        uint32_t QP = Q <= 4 ? 4 : 8;
        ProfMeta &m = db->metas[i];
        m.M = M, m.Q = Q, m.QP = QP, m.null_id = db->null_id[i];
        m.W = W, m.cls = cls;
        m.emis_off = emis_floats;
        m.trans_off = trans_floats;
        emis_floats += (uint64_t)kTab * 32 * QP * W;
        trans_floats += (uint64_t)8 * 32 * Q * W;
        db->class_list[m.cls].push_back((uint32_t)i);
    }
    CU_TRY(cudaMalloc(&db->d_emis, emis_floats * sizeof(float)));
    CU_TRY(cudaMalloc(&db->d_trans, trans_floats * sizeof(float)));
    CU_TRY(cudaMalloc(&db->d_metas, nprof * sizeof(ProfMeta)));
    CU_TRY(cudaMalloc(&db->d_null_tabs, db->null_tabs.size() * kTab * sizeof(float)));
    CU_TRY(cudaMalloc(&db->d_ins_tab, kTab * sizeof(float)));
    db->device_bytes = (emis_floats + trans_floats + (db->null_tabs.size() + 1) * kTab) * sizeof(float) +
                       nprof * sizeof(ProfMeta);

    /* Upload the tables as the host holds them ([node][code], contiguous) through two pinned buffers and let
     * k_layout transpose them into the kernels' [code][warp][half][lane][4] layout on the device. */
    const size_t raw_max = (size_t)DCP_PROTEIN_MODEL_CORE_SIZE_MAX * kTab;
    float *stage[2] = {nullptr, nullptr};
    float *d_raw[2] = {nullptr, nullptr};
    cudaEvent_t freed[2];
    for (int b = 0; b < 2; ++b)
    {
        CU_TRY(cudaMallocHost(&stage[b], raw_max * sizeof(float)));
        CU_TRY(cudaMalloc(&d_raw[b], raw_max * sizeof(float)));
        CU_TRY(cudaEventCreateWithFlags(&freed[b], cudaEventDisableTiming));
    }
    std::vector<float> tr_all(trans_floats, NEG_INF);
    for (size_t i = 0; i < nprof; ++i)
    {
        const ProfMeta &m = db->metas[i];
        const protein_profile *p = db->profs[i];
        const int b = (int)(i & 1);
        if (i >= 2) CU_TRY(cudaEventSynchronize(freed[b])); /* the copy out of stage[b] has completed */
        const size_t raw = (size_t)m.M * kTab;
        memcpy(stage[b], p->match_emission, raw * sizeof(float));
        CU_TRY(cudaMemcpyAsync(d_raw[b], stage[b], raw * sizeof(float), cudaMemcpyHostToDevice, db->stream));
        CU_TRY(cudaEventRecord(freed[b], db->stream));
        const uint32_t ROW = 32 * m.QP * m.W;
        const size_t out = (size_t)kTab * ROW;
        k_layout<<<(unsigned)((out + 255) / 256), 256, 0, db->stream>>>(d_raw[b], db->d_emis + m.emis_off, m.M, m.Q,
                                                                         m.QP, m.W);
        const uint32_t NP = 32 * m.Q * m.W;
        float *tr = tr_all.data() + m.trans_off;
        for (uint32_t k = 1; k <= m.M; ++k) /* node k, slot k-1 */
        {
            uint32_t n = k - 1;
            if (k >= 2)
            {
                const protein_trans &t = p->trans[k - 1];
                tr[0 * NP + n] = t.MM, tr[1 * NP + n] = t.IM, tr[2 * NP + n] = t.DM;
                tr[3 * NP + n] = t.MD, tr[4 * NP + n] = t.DD;
            }
            if (k <= m.M - 1)
            {
                const protein_trans &t = p->trans[k];
                tr[5 * NP + n] = t.MI, tr[6 * NP + n] = t.II;
            }
            tr[7 * NP + n] = p->entry[k - 1];
        }
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(db->stream));
    CU_TRY(cudaMemcpy(db->d_trans, tr_all.data(), trans_floats * sizeof(float), cudaMemcpyHostToDevice));
    for (int b = 0; b < 2; ++b)
    {
        cudaFreeHost(stage[b]);
        cudaFree(d_raw[b]);
        cudaEventDestroy(freed[b]);
    }
    CU_TRY(cudaMemcpy(db->d_metas, db->metas.data(), nprof * sizeof(ProfMeta), cudaMemcpyHostToDevice));
    for (size_t t = 0; t < db->null_tabs.size(); ++t)
        CU_TRY(cudaMemcpy(db->d_null_tabs + t * kTab, db->null_tabs[t].data(), kTab * sizeof(float),
                          cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(db->d_ins_tab, db->profs[0]->insert_emission, kTab * sizeof(float), cudaMemcpyHostToDevice));
    for (int q = 1; q <= kNumClasses; ++q)
        if (!db->class_list[q].empty())
        {
            CU_TRY(cudaMalloc(&db->d_class[q], db->class_list[q].size() * sizeof(uint32_t)));
            CU_TRY(cudaMemcpy(db->d_class[q], db->class_list[q].data(), db->class_list[q].size() * sizeof(uint32_t),
                              cudaMemcpyHostToDevice));
        }
    /* the host copies stay for decode and product rows; those need the nucleotide distributions, transitions
     * and names, not the 5.4 KB per node of match emissions that now live in HBM */
    for (auto *p : db->profs)
    {
        free(p->match_emission);
        p->match_emission = nullptr;
    }
    db->committed = true;
    return RC_OK;
}

extern "C" unsigned dcpgpu_db_nprofiles(struct dcpgpu_db const *db) { return (unsigned)db->profs.size(); }
extern "C" char const *dcpgpu_db_accession(struct dcpgpu_db const *db, unsigned i)
{
    return i < db->profs.size() ? db->profs[i]->accession : nullptr;
}
extern "C" unsigned dcpgpu_db_core_size(struct dcpgpu_db const *db, unsigned i)
{
    return i < db->profs.size() ? db->profs[i]->core_size : 0;
}
extern "C" uint64_t dcpgpu_db_device_bytes(struct dcpgpu_db const *db) { return db->device_bytes; }

extern "C" void dcpgpu_db_del(struct dcpgpu_db *db)
{
    if (!db) return;
    cudaSetDevice(db->device);
    for (auto *p : db->profs) protein_profile_del(p);
    cudaFree(db->d_emis), cudaFree(db->d_trans), cudaFree(db->d_metas);
    cudaFree(db->d_null_tabs), cudaFree(db->d_ins_tab);
    for (int q = 0; q <= kNumClasses; ++q) cudaFree(db->d_class[q]);
    if (db->h_stage) cudaFreeHost(db->h_stage);
    if (db->stream)
    {
        cudaStreamSynchronize(db->stream);
        cudaStreamDestroy(db->stream);
    }
    delete db;
}

const protein_profile *dcp_db_profile(struct dcpgpu_db const *db, unsigned i) { return db->profs[i]; }

extern "C" enum rc dcpgpu_seqs_new(struct dcpgpu_seqs **out, struct dcpgpu_db *db, unsigned nseqs,
                                   char const *const *seqs, unsigned const *lens)
{
    if (!db->committed) return dcp_error(RC_EFAIL, "commit the database first");
    if (nseqs == 0) return dcp_error(RC_EINVAL, "no sequences");
    CU_TRY(cudaSetDevice(db->device));
    dcpgpu_seqs *sq = new (std::nothrow) dcpgpu_seqs;
    if (!sq) return dcp_error(RC_ENOMEM, "alloc seqs");
    sq->db = db;
    sq->nseq = nseqs;
    sq->metas.resize(nseqs);
    uint64_t total = 0;
    for (unsigned i = 0; i < nseqs; ++i)
    {
        if (lens[i] == 0)
        {
            delete sq;
            return dcp_error(RC_EINVAL, "sequence cannot be empty"); /* protein_profile.c:158 */
        }
        sq->metas[i].len = lens[i], sq->metas[i].pad = 0, sq->metas[i].row_off = total;
        sq->metas[i].rec_off = total + i; /* L+1 records per sequence */
        total += lens[i];
    }
    sq->total = total;
    /* pinned staging buffer, cached in the db and grown on demand */
    if (db->h_stage_cap < total)
    {
        if (db->h_stage) cudaFreeHost(db->h_stage);
        db->h_stage = nullptr, db->h_stage_cap = 0;
        size_t cap = total + total / 4 + 4096;
        if (cudaMallocHost(&db->h_stage, cap) != cudaSuccess)
        {
            delete sq;
            return dcp_error(RC_ENOMEM, "pinned staging for sequences");
        }
        db->h_stage_cap = cap;
    }
    uint8_t *h = (uint8_t *)db->h_stage;
    static const int8_t lut_init = 0;
    (void)lut_init;
    int8_t lut[256];
    memset(lut, -1, sizeof lut);
    lut['A'] = 0, lut['C'] = 1, lut['G'] = 2, lut['T'] = 3;
    bool bad = false;
    for (unsigned i = 0; i < nseqs && !bad; ++i)
    {
        uint8_t *dst = h + sq->metas[i].row_off;
        const unsigned char *src = (const unsigned char *)seqs[i];
        for (unsigned j = 0; j < lens[i]; ++j)
        {
            int8_t c = lut[src[j]];
            if (c < 0)
            {
                bad = true;
                break;
            }
            dst[j] = (uint8_t)c;
        }
    }
    if (bad)
    {
        delete sq;
        return dcp_error(RC_EINVAL, "sequence symbol outside ACGT");
    }
    cudaError_t e = cudaMallocAsync(&sq->d_bases, total, db->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&sq->d_metas, nseqs * sizeof(SeqMeta), db->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(sq->d_bases, h, total, cudaMemcpyHostToDevice, db->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(sq->d_metas, sq->metas.data(), nseqs * sizeof(SeqMeta), cudaMemcpyHostToDevice, db->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(db->stream); /* staging buffer is reusable again */
    if (e != cudaSuccess)
    {
        dcp_set_error(cudaGetErrorString(e));
        dcpgpu_seqs_del(sq);
        return RC_EFAIL;
    }
    sq->h2d_bytes = total + nseqs * sizeof(SeqMeta);
    *out = sq;
    return RC_OK;
}

extern "C" void dcpgpu_seqs_del(struct dcpgpu_seqs *sq)
{
    if (!sq) return;
    cudaSetDevice(sq->db->device);
    cudaFreeAsync(sq->d_bases, sq->db->stream), cudaFreeAsync(sq->d_metas, sq->db->stream);
    delete sq;
}


extern "C" enum rc dcpgpu_scan_resident(struct dcpgpu_db *db, struct dcpgpu_seqs *sq,
                                        struct dcpgpu_params const *prm, struct dcpgpu_result **out)
{
    if (!db->committed || sq->db != db) return dcp_error(RC_EINVAL, "sequences were staged for another database");
    CU_TRY(cudaSetDevice(db->device));
    cudaStream_t st = db->stream;
    const uint32_t nseq = sq->nseq, nprof = (uint32_t)db->profs.size(), n_null = (uint32_t)db->null_tabs.size();
    const size_t npairs = (size_t)nseq * nprof;

    dcpgpu_result *res = new (std::nothrow) dcpgpu_result;
    if (!res) return dcp_error(RC_ENOMEM, "alloc result");
    res->db = db, res->nseq = nseq, res->nprof = nprof, res->n_null = n_null;
    struct Guard
    {
        dcpgpu_result *r;
        ~Guard() { if (r) dcpgpu_result_del(r); }
    } guard{res};

    /* protein_profile_setup per sequence length (host libm, exactly once per distinct L) */
    std::vector<float> spec((size_t)nseq * 16, 0.0f);
    {
        std::vector<std::pair<uint32_t, uint32_t>> byL(nseq);
        for (uint32_t s = 0; s < nseq; ++s) byL[s] = {sq->metas[s].len, s};
        std::sort(byL.begin(), byL.end());
        float x[13];
        uint32_t last = 0;
        for (auto &e : byL)
        {
            if (e.first != last) dcp_specials(e.first, prm->multi_hits, prm->hmmer3_compat, x), last = e.first;
            memcpy(&spec[(size_t)e.second * 16], x, sizeof x);
        }
    }
    DevBuf b_spec, b_rows, b_wcodes, b_counter, b_nhits;
    CU_TRY(b_spec.alloc(spec.size() * sizeof(float), st));
    const uint64_t total_recs = sq->total + nseq;
    CU_TRY(b_rows.alloc((size_t)n_null * total_recs * sizeof(RowRec), st));
    CU_TRY(b_wcodes.alloc(total_recs * sizeof(uint16_t), st));
    CU_TRY(b_counter.alloc((kNumClasses + 1) * sizeof(unsigned long long), st));
    CU_TRY(b_nhits.alloc(2 * sizeof(unsigned long long), st));
    CU_TRY(cudaMallocAsync(&res->d_alt, npairs * sizeof(float), st));
    CU_TRY(cudaMallocAsync(&res->d_null, (size_t)nseq * n_null * sizeof(float), st));
    CU_TRY(cudaMallocAsync(&res->d_hit, npairs, st));

    struct Events
    {
        cudaEvent_t e[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        ~Events()
        {
            for (auto x : e)
                if (x) cudaEventDestroy(x);
        }
    } events;
    cudaEvent_t(&ev)[5] = events.e;
    for (auto &e : ev) CU_TRY(cudaEventCreate(&e));
    uint64_t launches = 0;

    CU_TRY(cudaEventRecord(ev[0], st));
    CU_TRY(cudaMemcpyAsync(b_spec.p, spec.data(), spec.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemsetAsync(b_counter.p, 0, (kNumClasses + 1) * sizeof(unsigned long long), st));
    CU_TRY(cudaMemsetAsync(b_nhits.p, 0, 2 * sizeof(unsigned long long), st));
    k_rows<<<nseq, 128, 0, st>>>(sq->d_bases, sq->d_metas, nseq, db->d_null_tabs, db->d_ins_tab, n_null, total_recs,
                                 b_rows.as<RowRec>(), b_wcodes.as<uint16_t>());
    k_null<<<(nseq * n_null + 127) / 128, 128, 0, st>>>(sq->d_metas, nseq, n_null, total_recs, b_rows.as<RowRec>(),
                                                        b_spec.as<float>(), res->d_null);
    launches += 2;
    CU_TRY(cudaEventRecord(ev[1], st));

    const int nblocks = db->sm_count; /* persistent: one 8-warp block per SM (255 regs/thread) */
    /* sequences per L2 tile: their row records (64 B per row and null table) stay L2-resident while every
     * profile passes.  32 MB measured best (DRAM bytes per config-2 launch: 64 MB 145 GB, 48 MB 78 GB,
     * 32 MB 34 GB, 16 MB 74 GB; profiles/r01_tile_sweep_dram.csv): data read by all SMs is held in both L2
     * partitions, so about half of the 126 MB is usable for it. */
    uint32_t seq_tile;
    {
        const double rec_bytes_per_seq = (double)total_recs / nseq * sizeof(RowRec) * n_null;
        double tile_mb = 32.0;
        if (const char *e = getenv("DCPGPU_TILE_MB")) tile_mb = std::max(0.25, atof(e)); /* experiment knob */
        double t = (tile_mb * 1024 * 1024) / rec_bytes_per_seq;
        seq_tile = (uint32_t)std::min<double>(std::max<double>(t, kSeqChunk), 1 << 20);
        seq_tile = std::max<uint32_t>(kSeqChunk, seq_tile / kSeqChunk * kSeqChunk);
    }
    uint64_t cells = 0;
    for (int q = 1; q <= kMaxQ; ++q)
    {
        if (db->class_list[q].empty()) continue;
        uint32_t n_class = (uint32_t)db->class_list[q].size();
        unsigned long long *ctr = b_counter.as<unsigned long long>() + q;
#define LAUNCH(QQ)                                                                                         \
    case QQ:                                                                                               \
        launch_score<QQ>(nblocks, st, db->d_emis, db->d_trans, db->d_metas, db->d_class[q], n_class,       \
                         sq->d_metas, nseq, total_recs, b_rows.as<RowRec>(), b_wcodes.as<uint16_t>(), b_spec.as<float>(), \
                         res->d_alt,                                                                       \
                         nprof, ctr, seq_tile);                                                                      \
        break;
        switch (q)
        {
            LAUNCH(1) LAUNCH(2) LAUNCH(3) LAUNCH(4) LAUNCH(5) LAUNCH(6) LAUNCH(7) LAUNCH(8)
        }
#undef LAUNCH
        launches++;
        for (uint32_t id : db->class_list[q]) cells += (uint64_t)db->metas[id].M * sq->total;
    }
    for (int q = kMaxQ + 2; q <= kNumClasses; ++q)
    {
        if (db->class_list[q].empty()) continue;
        uint32_t n_class = (uint32_t)db->class_list[q].size();
        unsigned long long *ctr = b_counter.as<unsigned long long>() + q;
        /* warps per pair: 2..8 one block, 10/12/14/16 two blocks */
        const int tw = q >= kClsW3Q6 ? 3 : q > kMaxQ + kMaxGroupWarps ? 2 : q - kMaxQ;
        const int cl = tw > kMaxW ? 2 : 1, w = tw / cl;
        const unsigned blocks = cl == 2 ? (unsigned)(db->sm_count / 2 * 2) : (unsigned)(db->sm_count * (8 / w));
        cudaError_t le = cudaErrorInvalidValue;
#define MW_LAUNCH(WW, CC, ...)                                                                                \
    le = launch_group(k_score_mw<WW, CC, ##__VA_ARGS__>, CC, blocks, WW * 32, st, db->d_emis, db->d_trans, db->d_metas, \
                      db->d_class[q], n_class, sq->d_metas, nseq, total_recs, b_rows.as<RowRec>(),            \
                      b_wcodes.as<uint16_t>(), b_spec.as<float>(), res->d_alt, nprof, ctr, seq_tile)
        switch (q > kMaxQ + kMaxGroupWarps ? -q : tw)
        {
        case -kClsW2Q6: MW_LAUNCH(2, 1, 6); break;
        case -kClsW2Q7: MW_LAUNCH(2, 1, 7); break;
        case -kClsW3Q6: MW_LAUNCH(3, 1, 6); break;
        case -kClsW3Q7: MW_LAUNCH(3, 1, 7); break;
        case 2: MW_LAUNCH(2, 1); break;
        case 3: MW_LAUNCH(3, 1); break;
        case 4: MW_LAUNCH(4, 1); break;
        case 5: MW_LAUNCH(5, 1); break;
        case 6: MW_LAUNCH(6, 1); break;
        case 7: MW_LAUNCH(7, 1); break;
        case 8: MW_LAUNCH(8, 1); break;
        case 10: MW_LAUNCH(5, 2); break;
        case 12: MW_LAUNCH(6, 2); break;
        case 14: MW_LAUNCH(7, 2); break;
        case 16: MW_LAUNCH(8, 2); break;
        }
#undef MW_LAUNCH
        CU_TRY(le);
        launches++;
        for (uint32_t id : db->class_list[q]) cells += (uint64_t)db->metas[id].M * sq->total;
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaEventRecord(ev[2], st));

    k_lrt<<<(unsigned)((npairs + 255) / 256), 256, 0, st>>>(res->d_alt, res->d_null, db->d_metas, nseq, nprof, n_null,
                                                            prm->lrt_threshold, res->d_hit,
                                                            b_nhits.as<unsigned long long>());
    launches++;
    unsigned long long nhits = 0;
    CU_TRY(cudaMemcpyAsync(&nhits, b_nhits.p, sizeof nhits, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    uint64_t d2h = sizeof nhits;

    if (nhits)
    {
        DevBuf b_list;
        CU_TRY(b_list.alloc(nhits * sizeof(unsigned long long), st));
        k_collect<<<(unsigned)((npairs + 255) / 256), 256, 0, st>>>(res->d_hit, npairs,
                                                                    b_nhits.as<unsigned long long>() + 1,
                                                                    b_list.as<unsigned long long>());
        launches++;
        std::vector<unsigned long long> list(nhits);
        CU_TRY(cudaMemcpyAsync(list.data(), b_list.p, nhits * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        d2h += nhits * sizeof(unsigned long long);
        std::sort(list.begin(), list.end()); /* (sequence, profile) order, independent of scheduling */
        DevBuf b_ga, b_gn;
        CU_TRY(b_ga.alloc(nhits * sizeof(float), st));
        CU_TRY(b_gn.alloc(nhits * sizeof(float), st));
        CU_TRY(cudaMemcpyAsync(b_list.p, list.data(), nhits * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
        k_gather<<<(unsigned)((nhits + 255) / 256), 256, 0, st>>>(b_list.as<unsigned long long>(), nhits, res->d_alt,
                                                                  res->d_null, db->d_metas, nprof, n_null,
                                                                  b_ga.as<float>(), b_gn.as<float>());
        launches++;
        res->hit_alt.resize(nhits), res->hit_null.resize(nhits);
        CU_TRY(cudaMemcpyAsync(res->hit_alt.data(), b_ga.p, nhits * sizeof(float), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(res->hit_null.data(), b_gn.p, nhits * sizeof(float), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        d2h += nhits * 2 * sizeof(float);
        res->hits.resize(nhits);
        for (size_t i = 0; i < nhits; ++i)
        {
            res->hits[i].seq = (uint32_t)(list[i] / nprof);
            res->hits[i].prof = (uint32_t)(list[i] % nprof);
            res->hits[i].step_off = 0, res->hits[i].nsteps = 0;
        }
    }
    CU_TRY(cudaEventRecord(ev[3], st));
    if (nhits && prm->want_paths)
    {
        enum rc rc = dcp_trace_hits(db, sq, res, b_rows.as<RowRec>(), b_wcodes.as<uint16_t>(), b_spec.as<float>(),
                                    &launches);
        if (rc) return rc;
        res->have_paths = true;
    }
    CU_TRY(cudaEventRecord(ev[4], st));
    CU_TRY(cudaStreamSynchronize(st));
    CU_TRY(cudaGetLastError());

    dcpgpu_timing &t = res->timing;
    cudaEventElapsedTime(&t.prep_ms, ev[0], ev[1]);
    cudaEventElapsedTime(&t.score_ms, ev[1], ev[2]);
    cudaEventElapsedTime(&t.trace_ms, ev[3], ev[4]);
    cudaEventElapsedTime(&t.total_ms, ev[0], ev[4]);
    t.launches = launches;
    t.alt_cells = cells;
    t.h2d_bytes = spec.size() * sizeof(float);
    t.d2h_bytes = d2h + res->steps.size() * sizeof(dcp_step) + res->hits.size() * 8;
    guard.r = nullptr;
    *out = res;
    return RC_OK;
}

extern "C" enum rc dcpgpu_scan(struct dcpgpu_db *db, unsigned nseqs, char const *const *seqs, unsigned const *lens,
                               struct dcpgpu_params const *prm, struct dcpgpu_result **out)
{
    dcpgpu_seqs *sq = nullptr;
    enum rc rc = dcpgpu_seqs_new(&sq, db, nseqs, seqs, lens);
    if (rc) return rc;
    rc = dcpgpu_scan_resident(db, sq, prm, out);
    if (!rc) (*out)->timing.h2d_bytes += sq->h2d_bytes;
    dcpgpu_seqs_del(sq);
    return rc;
}

static enum rc result_fetch(dcpgpu_result *r)
{
    if (r->fetched) return RC_OK;
    CU_TRY(cudaSetDevice(r->db->device));
    size_t n = (size_t)r->nseq * r->nprof;
    r->alt.resize(n), r->null_ll.resize(n), r->hit.resize(n);
    std::vector<float> nt((size_t)r->nseq * r->n_null);
    CU_TRY(cudaMemcpy(r->alt.data(), r->d_alt, n * sizeof(float), cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(nt.data(), r->d_null, nt.size() * sizeof(float), cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(r->hit.data(), r->d_hit, n, cudaMemcpyDeviceToHost));
    for (uint32_t s = 0; s < r->nseq; ++s)
        for (uint32_t p = 0; p < r->nprof; ++p)
            r->null_ll[(size_t)s * r->nprof + p] = nt[(size_t)s * r->n_null + r->db->null_id[p]];
    r->fetched = true;
    return RC_OK;
}

extern "C" unsigned dcpgpu_result_nseqs(struct dcpgpu_result const *r) { return r->nseq; }
extern "C" unsigned dcpgpu_result_nprofiles(struct dcpgpu_result const *r) { return r->nprof; }
extern "C" float const *dcpgpu_result_null_loglik(struct dcpgpu_result const *r)
{
    return result_fetch(const_cast<dcpgpu_result *>(r)) ? nullptr : r->null_ll.data();
}
extern "C" float const *dcpgpu_result_alt_loglik(struct dcpgpu_result const *r)
{
    return result_fetch(const_cast<dcpgpu_result *>(r)) ? nullptr : r->alt.data();
}
extern "C" uint8_t const *dcpgpu_result_hit(struct dcpgpu_result const *r)
{
    return result_fetch(const_cast<dcpgpu_result *>(r)) ? nullptr : r->hit.data();
}
extern "C" uint64_t dcpgpu_result_nhits(struct dcpgpu_result const *r) { return r->hits.size(); }

extern "C" enum rc dcpgpu_result_hit_at(struct dcpgpu_result const *r, uint64_t i, unsigned *seq_idx,
                                        unsigned *prof_idx, struct dcp_step const **steps, unsigned *nsteps)
{
    if (i >= r->hits.size()) return dcp_error(RC_EINVAL, "hit index out of range");
    const HitRec &h = r->hits[i];
    if (seq_idx) *seq_idx = h.seq;
    if (prof_idx) *prof_idx = h.prof;
    if (steps) *steps = r->have_paths ? r->steps.data() + h.step_off : nullptr;
    if (nsteps) *nsteps = r->have_paths ? h.nsteps : 0;
    return RC_OK;
}

extern "C" void dcpgpu_result_timing(struct dcpgpu_result const *r, struct dcpgpu_timing *t) { *t = r->timing; }

extern "C" void dcpgpu_result_del(struct dcpgpu_result *r)
{
    if (!r) return;
    cudaSetDevice(r->db->device);
    cudaStream_t st = r->db->stream;
    cudaFreeAsync(r->d_alt, st), cudaFreeAsync(r->d_null, st), cudaFreeAsync(r->d_hit, st);
    delete r;
}
