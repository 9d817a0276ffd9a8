/*
 * dcp_trace.cu -- traceback pass for hits only (imm_prod.path of imm_dp_viterbi on the alt dp,
 * consumed by prod_fwrite, src/server/prod.c:153-181).
 *
 * The reference keeps one backpointer per (row, state) of the whole DP matrix of a task (imm_task, reused per
 * thread: src/server/scan_thread.c:40-55).  Here a hit is traced by row checkpointing, with the score kernels' own
 * row code (score_row / score_row_h / mw_row: same lane layout, same fp32 operation order, so every value is bit-equal
 * to the score pass's; T[L], which falls out of the walk's first step, is compared with it on the host):
 *
 *   forward   rows 1..L; before every segment of C rows the five-row ring of Tin_M / Tin_I / Tin_N,J,C -- all the
 *             state the recurrence carries -- is stored: 40 B per node and C rows instead of a backpointer per cell
 *   backward  segment by segment from the last: reload the ring, recompute the segment's rows, this time storing
 *             Tin_M, Tin_I and D of every cell and E, B, Tin_N/J/C of every row into a per-group scratch
 *             ((C + 5) rows, reused by every segment and every hit), then walk the path through the segment.  No
 *             backpointer is ever computed for a cell the path does not visit.  The cells are first stored for a band
 *             of nodes only -- the lanes just below the node the walk enters the segment at: a path moves down one
 *             node per match or delete step, about C / 3 nodes per segment -- so the scratch lines in use stay in L2
 *             and the stores do not evict the emission tables; a walk that needs a cell outside the band (a long
 *             run of deletions, or E, whose source can be any node) has the segment recomputed once more with
 *             every cell stored.
 *   walk      one warp per hit, one lane per candidate: at state s and row r the lanes evaluate the incoming
 *             (transition, source length) candidates of Tin_s[r] from the stored values -- fl(fl(Tin_src[r-l] +
 *             e_src(seq[r-l:r])) + t), the sums a start-position interpreter forms -- and the first lane that
 *             attains the maximum wins.  That is imm's rule: first maximum, strict '>', over incoming transitions in
 *             canonical order and, inside one transition, over the source's emission length ascending.
 *
 * Canonical incoming order (imm's own order is not recoverable from the reference tree):
 *   M_k: B, M_{k-1}, I_{k-1}, D_{k-1}    I_k: M_k, I_k    D_k: M_{k-1}, D_{k-1}
 *   E: M_1, M_2, D_2, M_3, D_3, ...      N: S, N    B: S, N, J, E    J: E, J   C: E, C   T: E, C
 * (D_k <= max_{k' < k} V_M(k') because M->D and D->D scores are <= 0 -- checked at commit -- and fp32 addition of a
 * non-positive number never rounds upwards, so under strict '>' a D state never wins E: E's source is the first
 * M_k, and inside it the first length, whose sum equals E.)
 */
#include "dcp_classes.h"
#include "dcp_score_mw.cuh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace
{

struct TraceJob
{
    uint32_t seq, prof;
    uint64_t ck_off;   /* floats into the checkpoint buffer */
    uint64_t step_off; /* first of `cap` steps of this hit in the raw step buffer (filled from the end) */
    uint32_t cap, pad;
};

struct TraceArgs
{
    const float *emis, *trans;
    const ProfMeta *metas;
    const SeqMeta *seqs;
    uint64_t total_recs;
    const RowRec *rows;
    const uint16_t *wcodes;
    const float *spec;
    const TraceJob *jobs;
    uint32_t njobs, C;             /* rows per segment (a multiple of 5) */
    unsigned long long *counter;   /* job queue cursor of this launch (zeroed) */
    float *ckpt;                   /* ring checkpoints of all jobs */
    float *scratch;                /* per resident group: (C + 5) rows of cells and row records */
    dcp_step *steps_raw;
    uint32_t *nsteps;              /* per job; 0 on failure */
    float *alt_out;                /* per job: T[L] as the walk's first step forms it (compared with the score pass) */
    uint32_t *errors;              /* [0] walks that left the DP matrix, [1] step buffers that were too small */
    unsigned long long *prof;      /* DCP_TRACE_PROF builds: cycles in forward rows, backward rows, walks, the rest of
                                    * a hit; walk steps; segments recomputed with every cell; segment passes */
};

#ifndef DCP_TRACE_PROF
#define DCP_TRACE_PROF 0
#endif
#define PROF_CLK() (DCP_TRACE_PROF ? clock64() : 0ll)

/* shape of a kernel class as the trace kernel sees it: TW = 0 the half-warp classes (16 lanes per hit; the trace kernel
 * runs the same hit on both halves of a warp: identical values, identical stores), 1 one warp, >= 2 a group of warps */
template <int TW, int Q>
struct Shape
{
    static constexpr int LN = TW == 0 ? 16 : 32;         /* lanes per warp-unit of a hit */
    static constexpr int NW = TW == 0 ? 1 : TW;          /* warps per hit */
    static constexpr int CL = NW > kMaxW ? 2 : 1;        /* blocks per hit */
    static constexpr int W = NW / CL;                    /* warps per block and hit */
    static constexpr int QP = Q <= 4 ? 4 : 8;
    static constexpr int NT = LN * NW;                   /* threads that hold distinct nodes */
    static constexpr int MP = NT * QP;                   /* padded node slots = floats per emission line */
    static constexpr int NP = NT * Q;                    /* stride of the transition arrays */
    static constexpr int BLOCK = TW <= 1 ? 128 : W * 32; /* TW <= 1: four independent warps per block */
    static constexpr int MINB = TW <= 1 ? 2 : (CL == 2 ? 1 : (8 / W > 0 ? 8 / W : 1));
    static constexpr int CK = 10 * MP + 32;              /* floats per ring checkpoint: [slot 5][M, I][MP] + [slot 5][4] */
    static constexpr int RR = 8;                         /* floats per row record of the scratch: E, B, Tin_N, Tin_J, Tin_C */
    static_assert(TW <= 1 || Q > 4, "warp groups use eight-float lanes");
    __host__ __device__ static size_t scratch_floats(uint32_t C) { return (size_t)(C + 5) * (3 * MP + RR); }
};

/* a lane's Q values as one or two 16-byte stores: [quad][thread][4], so a warp's store is contiguous */
template <int Q, int NT>
__device__ __forceinline__ void store_q(float *dst, const float (&v)[Q])
{
    float4 a = make_float4(v[0], Q > 1 ? v[1 % Q] : 0.f, Q > 2 ? v[2 % Q] : 0.f, Q > 3 ? v[3 % Q] : 0.f);
    *reinterpret_cast<float4 *>(dst) = a;
    if (Q > 4)
    {
        float4 b = make_float4(v[4 % Q], Q > 5 ? v[5 % Q] : 0.f, Q > 6 ? v[6 % Q] : 0.f, Q > 7 ? v[7 % Q] : 0.f);
        *reinterpret_cast<float4 *>(dst + NT * 4) = b;
    }
}
template <int Q, int NT>
__device__ __forceinline__ void load_q(const float *src, float (&v)[Q])
{
    const float4 a = __ldcg(reinterpret_cast<const float4 *>(src));
    const float t[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int i = 0; i < (Q < 4 ? Q : 4); ++i) v[i] = t[i];
    if (Q > 4)
    {
        const float4 b = __ldcg(reinterpret_cast<const float4 *>(src + NT * 4));
        const float u[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 4; i < Q; ++i) v[i % Q] = u[i - 4];
    }
}
/* slot of node n (0-based) in a cell array laid out by store_q */
template <int Q, int NT>
__device__ __forceinline__ uint32_t cell_slot(uint32_t n)
{
    const uint32_t t = n / Q, sub = n % Q;
    return (sub >> 2) * (NT * 4) + t * 4 + (sub & 3);
}

/* who this thread is inside its hit */
struct Who
{
    int lane;  /* lane of the warp-unit: 0..LN-1 */
    int half;  /* TW = 0: which half of the warp */
    int gw;    /* warp of the group */
    int t;     /* gw * LN + lane: owner of nodes t * Q .. t * Q + Q - 1 */
    int xs;    /* 0..2: this thread carries N / J / C; 3: it writes E and B of the row; else -1 */
};

/* length-dependent special scores: what a row needs (the walk reads the rest from the sequence's record when a
 * candidate uses it) */
struct RowSpecials
{
    float NB, JB, EB, cE, cX;
};
/* ----------------------------------------------------------------------------------------- */
/* rows j0 + 1 .. min(L, j0 + C) from the ring as it stands after row j0                       */
/* ----------------------------------------------------------------------------------------- */
template <int TW, int Q, int R>
__device__ __forceinline__ void one_row(float (&tm)[5][Q], float (&ti)[5][Q], float (&tx)[5], const NodeParams<Q> &p,
                                        RowState<Q> &rs, const float *__restrict__ emis_lane,
                                        const RowRec *__restrict__ recs, const uint16_t *__restrict__ wc, uint32_t L,
                                        uint32_t jj, uint32_t j0, const RowSpecials &k, const Who &me,
                                        Group<Shape<TW, Q>::CL, MwShared> &grp, const CarryBound &cb, bool store,
                                        bool cells_on, float *__restrict__ cells, float *__restrict__ rowrec)
{
    using S = Shape<TW, Q>;
    const RowRec *rn = recs + min(jj + 1u, L);
    const uint16_t *wn = wc + min(jj + 3u, L);
    RowTap<Q> tap;
    float E, vC;
    if constexpr (TW == 0)
        score_row_h<Q, R>(tm, ti, tx, p, rs, emis_lane, rn, wn, me.lane, me.half, k.NB, k.JB, k.EB, k.cE, k.cX, E, vC, &tap);
    else if constexpr (TW == 1)
    {
        TmaCtx tc = {nullptr, nullptr, 0};
        score_row<Q, R, false>(tm, ti, tx, p, rs, emis_lane, rn, wn, me.lane, k.NB, k.JB, k.EB, k.cE, k.cX, E, vC, tc, &tap);
    }
    else
        mw_row<S::W, S::CL, R, Q, 1>(tm, ti, tx, p, rs, emis_lane, rn, wn, me.gw, me.lane, (int)(jj & 1u), grp, k.NB, k.JB,
                                     k.EB, k.cE, k.cX, cb, E, vC, &tap);
    (void)vC; /* T[L] comes out of the walk's first step, whose candidates are exactly T's */
    if (store)
    {
        const uint32_t slot = jj - j0 + 4u;
        if (cells_on)
        {
            float *c = cells + (size_t)slot * (3 * S::MP);
            store_q<Q, S::NT>(c, tm[R]);
            store_q<Q, S::NT>(c + S::MP, ti[R]);
            store_q<Q, S::NT>(c + 2 * S::MP, tap.d);
        }
        if (me.xs >= 0)
        {
            float *rr = rowrec + (size_t)slot * S::RR;
            if (me.xs < 3)
                rr[2 + me.xs] = tx[R];
            else
                rr[0] = E, rr[1] = tap.B;
        }
    }
}

template <int TW, int Q>
__device__ __forceinline__ void run_rows(float (&tm)[5][Q], float (&ti)[5][Q], float (&tx)[5], const NodeParams<Q> &p,
                                         const float *__restrict__ emis_lane, const RowRec *__restrict__ recs,
                                         const uint16_t *__restrict__ wc, uint32_t L, uint32_t j0, uint32_t C,
                                         const RowSpecials &k, const Who &me, Group<Shape<TW, Q>::CL, MwShared> &grp,
                                         const CarryBound &cb, bool store, bool cells_on,
                                         float *__restrict__ cells, float *__restrict__ rowrec)
{
    using S = Shape<TW, Q>;
    /* pipeline prologue of the score kernels, at row j0 + 1 */
    RowState<Q> rs;
#pragma unroll
    for (int l = 0; l < 5; ++l) rs.eN[l] = NEG_INF;
    {
        uint32_t code[5];
        codes_of(__ldg(wc + j0 + 1), code);
        load_emis<Q, S::MP, S::LN * 4, emis256(TW, Q)>(rs.em, emis_lane, code);
    }
    load_row_insert(recs + j0 + 1, rs.eI);
    if (me.xs >= 0 && me.xs < 3) load_row_special(recs + j0 + 1, rs.eN);
    rs.w1 = __ldg(wc + min(j0 + 2u, L));
    rs.w2 = __ldg(wc + min(j0 + 3u, L));
    rs.w3 = 0;
    /* whole groups of five rows (static ring slots); rows past L recompute on clamped inputs and are ignored */
    const uint32_t j1 = min(L, j0 + C);
#define ONE(RR, jj) one_row<TW, Q, RR>(tm, ti, tx, p, rs, emis_lane, recs, wc, L, (jj), j0, k, me, grp, cb, store, cells_on, cells, rowrec)
#pragma unroll 1
    for (uint32_t j = j0 + 1; j <= j1; j += 5)
    {
        ONE(0, j);
        ONE(1, j + 1);
        ONE(2, j + 2);
        ONE(3, j + 3);
        ONE(4, j + 4);
    }
#undef ONE
}

/* ----------------------------------------------------------------------------------------- */
/* the walk                                                                                   */
/* ----------------------------------------------------------------------------------------- */
enum { W_S, W_N, W_B, W_E, W_J, W_C, W_T, W_M, W_I, W_D };

struct Walker
{
    int st;
    uint32_t k, r, len, n;
    bool bad, over, done;
    bool need_full; /* the walk stopped at a state whose candidates lie outside the stored band */
};

/* threads whose cells the backward pass stores: lo..hi (none when hi < lo), or all of them */
struct Band
{
    int lo, hi;
    bool full;
    __device__ __forceinline__ bool has(int t) const { return full || (t >= lo && t <= hi); }
};
/* band for a walk that enters a segment in state (st, k): the thread of node k and the `lanes` - 1 threads below it */
template <int Q>
__device__ __forceinline__ Band band_of(int st, uint32_t k, int lanes)
{
    Band b;
    b.full = false, b.lo = 1, b.hi = 0;
    if (st == W_M || st == W_I || st == W_D)
    {
        b.hi = (int)((k - 1u) / Q);
        b.lo = max(0, b.hi - lanes + 1);
    }
    return b;
}

__device__ __forceinline__ uint16_t state_id_of(int st, uint32_t k)
{
    switch (st)
    {
    case W_M: return (uint16_t)(PROTEIN_MATCH_STATE | k);
    case W_I: return (uint16_t)(PROTEIN_INSERT_STATE | k);
    case W_D: return (uint16_t)(PROTEIN_DELETE_STATE | k);
    case W_S: return PROTEIN_S_STATE;
    case W_N: return PROTEIN_N_STATE;
    case W_B: return PROTEIN_B_STATE;
    case W_E: return PROTEIN_E_STATE;
    case W_J: return PROTEIN_J_STATE;
    case W_C: return PROTEIN_C_STATE;
    default: return PROTEIN_T_STATE;
    }
}

/*
 * Candidates of Tin_s[r] by state and index, in the canonical order (header of this file), 16 bits each:
 *   kind:3 (0 none, WK_M / WK_I / WK_D a cell of the source node, WK_ROW a field of the row record) | source length:3 |
 *   source node is this node (1) or the one before (0):1 | transition: index of the per-node parameter array (core)
 *   or of the sequence's specials record:4 | row-record field:3 | needs a node before this one:1 | core:1
 */
enum { WK_NONE, WK_M, WK_I, WK_D, WK_ROW };
__device__ __forceinline__ uint32_t walk_entry(int st, int c)
{
    auto cellc = [](uint32_t kind, uint32_t l, uint32_t same, uint32_t par, uint32_t need) {
        return kind | l << 3 | same << 6 | par << 7 | need << 14 | 1u << 15;
    };
    auto rowc = [](uint32_t field, uint32_t l, uint32_t spec, uint32_t core) {
        return (uint32_t)WK_ROW | l << 3 | spec << 7 | field << 11 | core << 15;
    };
    /* specials record (dcp_specials): 0 NN, 1 CC, 2 JJ, 3 NB, 4 CT, 5 JB, 9 E->T, 10 E->C, 11 E->B, 12 E->J;
     * parameter arrays: 0 MM, 1 IM, 2 DM, 3 MD, 4 DD (into this node), 5 MI, 6 II (own), 7 entry */
    const bool lo = c >= 1 && c <= 5, hi = c >= 6 && c <= 10;
    const uint32_t l = lo ? (uint32_t)c : hi ? (uint32_t)c - 5u : 0u;
    switch (st)
    {
    case W_M:
        if (c == 0) return rowc(1, 0, 7, 1); /* B[r] + entry */
        if (lo) return cellc(WK_M, l, 0, 0, 1);
        if (hi) return cellc(WK_I, l, 0, 1, 1);
        if (c == 11) return cellc(WK_D, 0, 0, 2, 1);
        return 0;
    case W_I:
        if (c < 5) return cellc(WK_M, (uint32_t)c + 1u, 1, 5, 0);
        if (c < 10) return cellc(WK_I, (uint32_t)c - 4u, 1, 6, 0);
        return 0;
    case W_D:
        if (c < 5) return cellc(WK_M, (uint32_t)c + 1u, 0, 3, 1);
        if (c == 5) return cellc(WK_D, 0, 0, 4, 1);
        return 0;
    case W_T: return c == 0 ? rowc(0, 0, 9, 0) : lo ? rowc(4, l, 4, 0) : 0;
    case W_C: return c == 0 ? rowc(0, 0, 10, 0) : lo ? rowc(4, l, 1, 0) : 0;
    case W_J: return c == 0 ? rowc(0, 0, 12, 0) : lo ? rowc(3, l, 2, 0) : 0;
    case W_N: return lo ? rowc(2, l, 0, 0) : 0;
    case W_B: return lo ? rowc(2, l, 3, 0) : hi ? rowc(3, l, 5, 0) : c == 11 ? rowc(0, 0, 11, 0) : 0;
    default: return 0;
    }
}
constexpr int kWalkTab = 10 * 16; /* [state][candidate] */

/* frame-table code of seq[r-l:r] from the packed window of row r */
__device__ __forceinline__ uint32_t code_of_len(uint32_t w, uint32_t l)
{
    const uint32_t off = l == 1 ? 0u : l == 2 ? 4u : l == 3 ? 20u : l == 4 ? 84u : 340u;
    return off + (w & ((1u << (2u * l)) - 1u));
}

/*
 * Walk while the current row lies in the segment above row j0 (rows j0 + 1 .. ; scratch slot of row x = x - j0 + 4,
 * slots 0..4 = the ring rows j0 - 4 .. j0).  All 32 lanes of the warp run this with identical walker state; lane c
 * evaluates candidate c.
 */
template <int TW, int Q>
__device__ __forceinline__ void walk_segment(Walker &w, uint32_t j0, uint32_t M, const float *__restrict__ cells,
                                             const float *__restrict__ rowrec, const float *__restrict__ emis,
                                             const float *__restrict__ tr, const RowRec *__restrict__ recs,
                                             const uint16_t *__restrict__ wc, const float *__restrict__ spv,
                                             dcp_step *__restrict__ out, uint32_t cap, int c, const Band &band,
                                             float *__restrict__ alt_out, const uint16_t *tab)
{
    using S = Shape<TW, Q>;
    auto cell = [&](int which, uint32_t n, uint32_t row) -> float {
        return __ldcg(cells + ((size_t)(row - j0 + 4u) * 3 + which) * S::MP + cell_slot<Q, S::NT>(n));
    };
    auto rowv = [&](int idx, uint32_t row) -> float { return __ldcg(rowrec + (size_t)(row - j0 + 4u) * S::RR + idx); };
    auto emis_m = [&](uint32_t n, uint32_t code) -> float {
        const uint32_t t = n / Q, sub = n % Q, wp = t / S::LN, ln = t % S::LN;
        return __ldg(emis + (size_t)code * S::MP + wp * (S::LN * S::QP) + emis_pos(ln, sub, S::LN, TW, Q));
    };

    auto stored = [&](uint32_t n) -> bool { return band.has((int)(n / Q)); };
    w.need_full = false;
    /* window of the current row; the windows of the five rows below it are loaded while a step's candidates are
     * evaluated, so that the next step's emission addresses do not wait for a load of their own */
    uint32_t win = __ldg(wc + w.r);
    while (!w.done && !w.bad && !w.over && (w.r > j0 || j0 == 0))
    {
        /* are the cells this state's candidates read stored?  (before the step is emitted: the walk resumes here) */
        if (w.r != 0 && !band.full)
        {
            const uint32_t n = w.k - 1u;
            const bool ok = w.st == W_E   ? false
                            : w.st == W_M ? (n == 0 || stored(n - 1))
                            : w.st == W_I ? stored(n)
                            : w.st == W_D ? (n == 0 || stored(n - 1))
                                          : true;
            if (!ok)
            {
                w.need_full = true;
                break;
            }
        }
        /* arrive: emit this state's step (the path is produced backwards: the buffer fills from its end) */
        if (w.n >= cap)
        {
            w.over = true;
            break;
        }
        if (c == 0)
        {
            dcp_step s;
            s.state_id = state_id_of(w.st, w.k), s.seqlen = (uint8_t)w.len;
            out[cap - 1u - w.n] = s;
        }
        w.n++;
        if (w.st == W_S)
        {
            w.done = true;
            break;
        }
        const uint32_t r = w.r;
        const uint32_t wnext = __ldg(wc + (r - min((uint32_t)c, min(r, 5u)))); /* lane c: row r - c */
        int nst = w.st;
        uint32_t nk = w.k, src_len = 0;
        if (r == 0)
        {
            /* row 0: S -> N, S -> B, B -> M_k are the only finite entries */
            if (w.st == W_M) nst = W_B;
            else if (w.st == W_N || w.st == W_B) nst = W_S;
            else w.bad = true;
        }
        else if (w.st == W_E)
        {
            /* first M_k in node order, then the first length, whose sum equals E[r] */
            const float E = rowv(0, r);
            bool found = false;
            for (uint32_t nb = 0; nb < M && !found; nb += 32)
            {
                const uint32_t n = nb + (uint32_t)c;
                int lhit = 0;
                if (n < M)
                {
#pragma unroll
                    for (int l = 5; l >= 1; --l)
                    {
                        const float s = cell(0, n, r - l) + emis_m(n, code_of_len(win, l));
                        if (s == E) lhit = l;
                    }
                }
                const unsigned who = __ballot_sync(FULL, lhit != 0);
                if (who)
                {
                    const int src = __ffs(who) - 1;
                    nk = nb + (uint32_t)src + 1u;
                    src_len = (uint32_t)__shfl_sync(FULL, lhit, src);
                    nst = W_M;
                    found = true;
                }
            }
            if (!found) w.bad = true;
        }
        else
        {
            /*
             * Lane c describes candidate c -- where its source value, its emission and its transition score live --
             * from a table entry and selects, then every lane loads and adds at once: (a + e) + t, e = 0 for a mute
             * source (B, E, D).  No branch depends on the lane: branches per candidate kind serialise their loads (three
             * L2 round trips per step instead of one), and a divergent description costs more than the arithmetic.
             */
            const RowRec *rec = recs + r;
            const uint32_t n = w.k - 1u; /* core states: this node */
            const uint32_t ent = tab[w.st * 16 + (c & 15)] & (c < 16 ? 0xffffu : 0u);
            const uint32_t kind = ent & 7u, l = (ent >> 3) & 7u, dn = (ent >> 6) & 1u, tix = (ent >> 7) & 15u;
            const uint32_t idx = (ent >> 11) & 7u;
            const bool need_n1 = (ent >> 14) & 1u, core = (ent >> 15) & 1u;
            const bool valid = kind != 0 && !(need_n1 && n == 0);
            const uint32_t src = valid && kind <= WK_D ? n - 1u + dn : 0u;
            const uint32_t slot = valid ? r - l - j0 + 4u : 4u;
            const float *pa = kind <= WK_D ? cells + ((size_t)slot * 3 + (kind - 1u)) * S::MP + cell_slot<Q, S::NT>(src)
                                           : rowrec + (size_t)slot * S::RR + idx;
            if (!valid) pa = rowrec;
            const uint32_t t_ = src / Q, sub = src % Q, wp = t_ / S::LN, ln = t_ % S::LN;
            const uint32_t lc = l ? l : 1u;
            const float *pe_m = emis + (size_t)code_of_len(win, lc) * S::MP + wp * (S::LN * S::QP) + emis_pos(ln, sub, S::LN, TW, Q);
            const float *pe_r = (kind == WK_I ? rec->eI : rec->eN) + (lc - 1u);
            const bool emits = valid && l != 0 && kind != WK_D;
            const float *pe = kind == WK_M ? pe_m : pe_r;
            const float *pt = core ? tr + tix * S::NP + n : spv + tix;
            const float va = __ldcg(pa);
            const float ve = emits ? __ldg(pe) : 0.0f;
            const float vt = valid ? __ldg(pt) : 0.0f;
            const float v = valid ? (va + ve) + vt : NEG_INF;
            const float best = warp_max(v);
            /* the candidates of T at row L are T[L] = max(E[L] + (EC+CT), V_C[L] + CT) itself: the alt log-likelihood
             * of this pass, compared with the score pass's on the host */
            if (w.st == W_T && c == 0) *alt_out = best;
            const unsigned who = __ballot_sync(FULL, v == best);
            const int code = who ? __ffs(who) - 1 : 0;
            /* the winner's entry says what the source is */
            const uint32_t went = __shfl_sync(FULL, valid ? ent : 0u, code);
            const uint32_t wkind = went & 7u;
            src_len = (went >> 3) & 7u;
            if (wkind == 0)
                w.bad = true; /* nothing finite leads here (or S below row 0) */
            else if (wkind <= WK_D)
            {
                nst = wkind == WK_M ? W_M : wkind == WK_I ? W_I : W_D;
                nk = w.k - 1u + ((went >> 6) & 1u);
                if (nk == 0) w.bad = true;
            }
            else
            {
                const uint32_t f = (went >> 11) & 7u; /* row-record field of the source: E, B, Tin_N, Tin_J, Tin_C */
                nst = f == 0 ? W_E : f == 1 ? W_B : f == 2 ? W_N : f == 3 ? W_J : W_C;
            }
        }
        if (src_len > r) w.bad = true;
        if (nst == W_S && r != 0) w.bad = true;
        if (w.bad) break;
        win = __shfl_sync(FULL, wnext, (int)src_len);
        w.r = r - src_len;
        w.st = nst, w.k = nk, w.len = src_len;
    }
}

/*
 * The walk of one segment pass as a real call.  Everything it needs -- the hit, its profile and sequence, a dozen
 * pointers, the walker -- is re-derived here from the job index and from shared memory, so none of it is live in the
 * caller's row loop (inlined, the walk's arguments and the walker's state cost the rows ~20 registers at a budget of 255
 * with no slack: spills in every row; forward rows ran 25..60 % slower than the score kernel's).
 */
struct WalkShared
{
    Walker w;
    int band_lo, band_hi, again; /* the band the next pass stores; again: recompute this segment with every cell */
};

template <int TW, int Q>
__device__ __noinline__ void walk_pass(const TraceArgs *__restrict__ a, WalkShared *ws, WalkShared *peer_ws,
                                       const uint16_t *tab, uint32_t job, uint32_t j0, float *scr)
{
    using S = Shape<TW, Q>;
    const int c = threadIdx.x & 31;
    const TraceJob tj = a->jobs[job];
    const ProfMeta pm = a->metas[tj.prof];
    const SeqMeta sm = a->seqs[tj.seq];
    const float *rowrec = scr + (size_t)(a->C + 5) * (3 * S::MP);
    Walker w = ws->w;
    Band band;
    band.lo = ws->band_lo, band.hi = ws->band_hi, band.full = ws->again != 0; /* what this pass stored */
    walk_segment<TW, Q>(w, j0, pm.M, scr, rowrec, a->emis + pm.emis_off, a->trans + pm.trans_off,
                        a->rows + (size_t)pm.null_id * a->total_recs + sm.rec_off, a->wcodes + sm.rec_off,
                        a->spec + (size_t)tj.seq * 16, a->steps_raw + tj.step_off, tj.cap, c, band, a->alt_out + job, tab);
    const int band_lanes = (int)((a->C / 3u + 8u) / Q + 2u);
    const Band next = band_of<Q>(w.st, w.k, band_lanes);
    __syncwarp();
    if (c == 0)
    {
        ws->w = w;
        ws->band_lo = next.lo, ws->band_hi = next.hi, ws->again = w.need_full;
        if (peer_ws) peer_ws->band_lo = next.lo, peer_ws->band_hi = next.hi, peer_ws->again = w.need_full;
    }
    __syncwarp();
}

/* ----------------------------------------------------------------------------------------- */
/* one hit after the other: forward pass with ring checkpoints, then segments backwards        */
/* ----------------------------------------------------------------------------------------- */
template <int TW, int Q>
__global__ void __launch_bounds__(Shape<TW, Q>::BLOCK, Shape<TW, Q>::MINB) k_trace(const TraceArgs a)
{
    using S = Shape<TW, Q>;
    constexpr int CL = S::CL;
    __shared__ MwShared sh;
    __shared__ TraceArgs sh_args;                    /* for walk_pass: a called function cannot see the kernel's parameters */
    __shared__ WalkShared sh_walk[TW <= 1 ? 4 : 1];  /* walker state and its verdict, per hit in flight in this block */
    Group<CL, MwShared> grp;
    grp.init(&sh);
    __shared__ uint16_t sh_tab[kWalkTab];            /* the walk's candidate table */
    if (threadIdx.x == 0) sh_args = a;
    for (int i = threadIdx.x; i < kWalkTab; i += blockDim.x) sh_tab[i] = (uint16_t)walk_entry(i / 16, i % 16);
    __syncthreads();
    WalkShared *const ws = sh_walk + (TW <= 1 ? (threadIdx.x >> 5) : 0);
    WalkShared *peer_ws = nullptr;
    if constexpr (CL == 2) peer_ws = cg::this_cluster().map_shared_rank(ws, grp.rank ^ 1);
    Who me;
    const int wlane = threadIdx.x & 31;
    uint32_t group; /* index of this hit-at-a-time unit on the device: owner of one scratch area */
    if constexpr (TW <= 1)
    {
        me.lane = wlane & (S::LN - 1), me.half = TW == 0 ? wlane >> 4 : 0, me.gw = 0;
        group = blockIdx.x * (S::BLOCK / 32) + (threadIdx.x >> 5);
    }
    else
    {
        me.lane = wlane, me.half = 0, me.gw = grp.rank * S::W + (int)(threadIdx.x >> 5);
        group = blockIdx.x / CL;
        if (CL == 2) grp.sync(); /* both blocks' shared memory exists before the first remote store */
        if constexpr (CL == 2) grp.exchange_init(S::W);
    }
    me.t = me.gw * S::LN + me.lane;
    me.xs = (me.gw == 0 && me.lane < 4) ? me.lane : -1;
    const bool walker = me.gw == 0; /* the warp that walks (TW = 0: both halves, as one warp of 32 candidates) */
    float *const scr = a.scratch + (size_t)group * S::scratch_floats(a.C);
    float *const cells = scr + (size_t)me.t * 4;                  /* this thread's first quad of row slot 0 */
    float *const rowrec = scr + (size_t)(a.C + 5) * (3 * S::MP);  /* row records after the cell rows */
    const uint32_t C = a.C;

    for (;;)
    {
        uint32_t job;
        if constexpr (TW <= 1)
        {
            unsigned long long it = 0;
            if (wlane == 0) it = atomicAdd(a.counter, 1ULL);
            job = (uint32_t)min(__shfl_sync(FULL, it, 0), (unsigned long long)a.njobs);
        }
        else
        {
            if (grp.rank == 0 && threadIdx.x == 0)
            {
                unsigned long long it = atomicAdd(a.counter, 1ULL);
                GRP_PUT(grp, item, it);
            }
            grp.sync();
            job = (uint32_t)min(sh.item, (unsigned long long)a.njobs);
            grp.sync();
        }
        if (job >= a.njobs) break;
        const TraceJob tj = a.jobs[job];
        const ProfMeta pm = a.metas[tj.prof];
        const SeqMeta sm = a.seqs[tj.seq];
        const uint32_t L = sm.len;
        NodeParams<Q> p;
        load_params<Q>(p, a.trans + pm.trans_off, S::NP, me.t * Q);
        CarryBound cb = {NEG_INF, NEG_INF, NEG_INF};
        if (S::NW > 2 && me.lane < S::NW)
        {
            const float *b = a.trans + pm.trans_off + 8 * S::NP; /* [3][NW] after the eight parameter arrays */
            cb.S = __ldg(b + me.lane), cb.md0 = __ldg(b + S::NW + me.lane), cb.dd0 = __ldg(b + 2 * S::NW + me.lane);
        }
        const float *emis_prof = a.emis + pm.emis_off;
        const float *emis_lane = emis_prof + me.gw * (S::LN * S::QP) + me.lane * emis_lane_stride(TW, Q);
        const RowRec *recs = a.rows + (size_t)pm.null_id * a.total_recs + sm.rec_off;
        const uint16_t *wc = a.wcodes + sm.rec_off;
        const float *spv = a.spec + (size_t)tj.seq * 16;
        RowSpecials k;
        k.NB = spv[3], k.JB = spv[5], k.EB = spv[11];
        k.cE = me.xs == 0 ? NEG_INF : (me.xs == 1 ? spv[12] : spv[10]); /* E->J + J->J, E->C + C->C */
        k.cX = me.xs == 0 ? spv[0] : (me.xs == 1 ? spv[2] : spv[1]);    /* N->N, J->J, C->C */
        const uint32_t nseg = (L + C - 1) / C;
        float *const ck = a.ckpt + tj.ck_off + (size_t)me.t * 4;

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
        for (int i = 0; i < Q; ++i) tm[4][i] = k.NB + p.ent[i];
        tx[4] = me.xs == 0 ? spv[0] : NEG_INF;

        /*
         * One loop, one copy of the row code (ten copies -- five ring rotations, twice -- thrashed the instruction
         * cache: no_instruction was the top stall of the rows): forward over the segments, checkpointing the ring
         * before each; then backward, reloading the ring, recomputing the segment with cells stored, and walking.
         */
        long long pc_job = PROF_CLK(), pc_fwd = 0, pc_bwd = 0, pc_walk = 0, pc_full = 0, pc_pass = 0;
        if (wlane == 0 && (walker || CL == 2))
        {
            /* the walk starts at T in row L; no cells are stored while it is in a special state */
            Walker w0;
            w0.st = W_T, w0.k = 0, w0.r = L, w0.len = 0, w0.n = 0, w0.bad = false, w0.over = false, w0.done = false;
            w0.need_full = false;
            ws->w = w0, ws->band_lo = 1, ws->band_hi = 0, ws->again = 0;
        }
        Band band;
        band.lo = 1, band.hi = 0, band.full = false;
        bool back = false;
        uint32_t s = 0;
#pragma unroll 1
        for (;;)
        {
            const uint32_t j0 = s * C;
            float *c = ck + (size_t)s * S::CK;
            float *cx = a.ckpt + tj.ck_off + (size_t)s * S::CK + 10 * S::MP;
            const bool cells_on = back && band.has(me.t);
            if (!back)
            {
#pragma unroll
                for (int q = 0; q < 5; ++q)
                {
                    store_q<Q, S::NT>(c + (q * 2 + 0) * S::MP, tm[q]);
                    store_q<Q, S::NT>(c + (q * 2 + 1) * S::MP, ti[q]);
                }
                if (me.xs >= 0 && me.xs < 3)
                {
#pragma unroll
                    for (int q = 0; q < 5; ++q) cx[q * 4 + me.xs] = tx[q];
                }
            }
            else
            {
#pragma unroll
                for (int q = 0; q < 5; ++q)
                {
                    load_q<Q, S::NT>(c + (q * 2 + 0) * S::MP, tm[q]);
                    load_q<Q, S::NT>(c + (q * 2 + 1) * S::MP, ti[q]);
                    tx[q] = NEG_INF;
                }
                if (me.xs >= 0 && me.xs < 3)
                {
#pragma unroll
                    for (int q = 0; q < 5; ++q) tx[q] = __ldcg(cx + q * 4 + me.xs);
                }
                /* the ring rows j0 - 4 .. j0 are slots 0 .. 4 of the scratch: the walk reads up to five rows back */
#pragma unroll
                for (int q = 0; q < 5; ++q)
                {
                    float *d = cells + (size_t)q * (3 * S::MP);
                    if (cells_on)
                    {
                        store_q<Q, S::NT>(d, tm[q]);
                        store_q<Q, S::NT>(d + S::MP, ti[q]);
                    }
                    if (me.xs >= 0 && me.xs < 3) rowrec[q * S::RR + 2 + me.xs] = tx[q];
                }
            }
            const long long pc0 = PROF_CLK();
            run_rows<TW, Q>(tm, ti, tx, p, emis_lane, recs, wc, L, j0, C, k, me, grp, cb, back, cells_on, cells, rowrec);
            if (back) pc_bwd += PROF_CLK() - pc0, pc_pass++, pc_full += band.full;
            else pc_fwd += PROF_CLK() - pc0;
            if (!back)
            {
                if (s + 1 < nseg)
                {
                    ++s;
                    continue;
                }
                back = true; /* row L has passed: the last segment again, this time for the walk */
                /* warp groups: the shared row buffers alternate by row parity, and the backward pass restarts at other rows */
                if constexpr (TW > 1) grp.sync();
                continue;
            }
            if constexpr (TW <= 1) __syncwarp();
            else grp.sync();
            if (walker)
            {
                const long long pc1 = PROF_CLK();
                walk_pass<TW, Q>(&sh_args, ws, peer_ws, sh_tab, job, j0, scr);
                pc_walk += PROF_CLK() - pc1;
            }
            if constexpr (TW <= 1) __syncwarp();
            else grp.sync();
            band.lo = ws->band_lo, band.hi = ws->band_hi, band.full = ws->again != 0;
            if (band.full) continue; /* once more with every cell: the walk left the band */
            if (s == 0) break;
            --s;
        }
        if (walker && wlane == 0)
        {
            const Walker w = ws->w;
            const bool ok = w.done && !w.bad && !w.over;
            a.nsteps[job] = ok ? w.n : 0u;
            if (!ok) atomicAdd(a.errors + (w.over ? 1 : 0), 1u);
            if (DCP_TRACE_PROF)
            {
                atomicAdd(a.prof + 0, (unsigned long long)pc_fwd), atomicAdd(a.prof + 1, (unsigned long long)pc_bwd);
                atomicAdd(a.prof + 2, (unsigned long long)pc_walk);
                atomicAdd(a.prof + 3, (unsigned long long)(PROF_CLK() - pc_job));
                atomicAdd(a.prof + 4, (unsigned long long)w.n), atomicAdd(a.prof + 5, (unsigned long long)pc_full);
                atomicAdd(a.prof + 6, (unsigned long long)pc_pass), atomicAdd(a.prof + 7, 1ull);
            }
        }
        if constexpr (TW <= 1) __syncwarp(); /* the next hit's walker is written by lane 0 */
    }
    if constexpr (CL == 2) grp.sync(); /* no block may exit while its peer can still store into its shared memory */
}

/* dense copy of the paths: hit i owns nsteps[i] steps at the end of its raw buffer */
__global__ void k_pack(const TraceJob *__restrict__ jobs, uint32_t njobs, const uint32_t *__restrict__ nsteps,
                       const uint64_t *__restrict__ off, const dcp_step *__restrict__ raw, dcp_step *__restrict__ out)
{
    const uint32_t job = blockIdx.x;
    if (job >= njobs) return;
    const uint32_t n = nsteps[job];
    const dcp_step *src = raw + jobs[job].step_off + (jobs[job].cap - n);
    dcp_step *dst = out + off[job];
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

template <int TW, int Q>
cudaError_t launch_trace(cudaStream_t st, int sm_count, TraceArgs a, float *scratch_pool, size_t *scratch_used,
                         bool dry)
{
    using S = Shape<TW, Q>;
    /* resident groups: as many as the device holds at once, no more than there are hits */
    const uint32_t per_sm = TW <= 1 ? (uint32_t)(S::MINB * S::BLOCK / 32) : (uint32_t)S::MINB;
    uint32_t groups = S::CL == 2 ? (uint32_t)(sm_count / 2) : (uint32_t)sm_count * per_sm;
    groups = std::min(groups, a.njobs);
    if (TW <= 1) groups = (groups + (S::BLOCK / 32) - 1) / (S::BLOCK / 32) * (S::BLOCK / 32); /* whole blocks of warps */
    const size_t need = (size_t)groups * S::scratch_floats(a.C);
    if (dry)
    {
        *scratch_used += need;
        return cudaSuccess;
    }
    a.scratch = scratch_pool + *scratch_used;
    *scratch_used += need;
    /* little or no shared memory: give the unified array to L1 (emission lines, row records), as the score kernels do */
    cudaFuncSetAttribute(k_trace<TW, Q>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
    unsigned blocks;
    if (TW <= 1) blocks = groups / (S::BLOCK / 32);
    else blocks = groups * S::CL;
    return launch_group(k_trace<TW, Q>, S::CL, blocks, S::BLOCK, st, a);
}

uint32_t segment_rows(uint32_t Lmax)
{
    /* about sqrt(L) rows, a multiple of five: the checkpoints (40 B per node and segment) and the scratch ((C + 5)
     * rows) both stay small, and a segment is long enough to amortise its fixed cost (ring reload, pipeline prologue,
     * two group barriers; 2 sqrt(L) measured the same on 1 kbp reads -- 27 vs 29 ms for 10 000 hits -- and
     * worse on 10 kbp contigs, 103 vs 96 ms for 200 hits: the band widens with C).  DCPGPU_TRACE_SEG overrides (tests,
     * experiments). */
    static const int forced = getenv("DCPGPU_TRACE_SEG") ? atoi(getenv("DCPGPU_TRACE_SEG")) : 0;
    if (forced > 0) return (uint32_t)std::min(1000, (forced + 4) / 5 * 5);
    uint32_t c = (uint32_t)std::ceil(std::sqrt((double)Lmax) / 5.0) * 5u;
    return std::min(250u, std::max(20u, c));
}

} // namespace

enum rc dcp_trace_hits(dcpgpu_db *db, dcpgpu_seqs *sq, dcpgpu_result *res, const RowRec *d_rows,
                       const uint16_t *d_wcodes, const float *d_spec, uint64_t *launches)
{
    cudaStream_t st = db->stream;
    const size_t nhits = res->hits.size();
    /* DCPGPU_TRACE_TIMES=1: host-side timeline of the pass on stderr (where the time between the kernels goes) */
    static const bool times = getenv("DCPGPU_TRACE_TIMES") && atoi(getenv("DCPGPU_TRACE_TIMES")) != 0;
    const auto t0 = std::chrono::steady_clock::now();
    auto stamp = [&](const char *what) {
        if (times)
            fprintf(stderr, "[dcp_trace] %8.3f ms  %s\n",
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), what);
    };
    /* free memory = what was free after commit less what the pool has handed out (no cudaMemGetInfo here: it
     * took 20 ms of a 50 ms pass) */
    uint64_t pool_used = 0;
    cudaMemPoolGetAttribute(db->pool, cudaMemPoolAttrUsedMemCurrent, &pool_used);
    const size_t free_b = db->free_at_commit > pool_used ? db->free_at_commit - (size_t)pool_used : 0;
    const size_t budget = std::min<size_t>(std::max<size_t>(free_b / 2, (size_t)64 << 20), (size_t)16 << 30);
    const std::vector<float> &score_alt = res->hit_alt; /* score pass result of each hit */

    /* checkpoint floats of a hit, given its class's segment length */
    auto ck_floats = [&](const HitRec &h, uint32_t C) -> size_t {
        const ProfMeta &m = db->metas[h.prof];
        const size_t MP = (size_t)m.LN * m.W * m.QP, L = sq->metas[h.seq].len;
        return ((L + C - 1) / C) * (10 * MP + 32);
    };
    /* steps a hit's buffer holds: generous for ordinary paths; a path that does not fit (many passes through the
     * core) has its batch retried with eight times the room, up to the most a path can have.
     * DCPGPU_TRACE_CAP overrides the first guess (tests). */
    static const int cap0 = getenv("DCPGPU_TRACE_CAP") ? atoi(getenv("DCPGPU_TRACE_CAP")) : 0;
    auto step_cap = [&](const HitRec &h, uint32_t mul) -> uint64_t {
        const uint64_t L = sq->metas[h.seq].len, M = db->metas[h.prof].M;
        const uint64_t first = cap0 > 0 ? (uint64_t)cap0 : 2 * (L + M) + 64;
        return std::min<uint64_t>(first * mul, (L + 2) * (M + 3) + 8);
    };

    size_t done = 0;
    uint32_t cap_mul = 1;
    while (done < nhits)
    {
        /* segment length per class from the longest sequence among ALL remaining hits of the class would need a
         * pass of its own; the longest sequence of the batch is close enough and keeps this a single loop */
        uint32_t Lmax = 1;
        for (size_t i = done; i < nhits; ++i) Lmax = std::max(Lmax, sq->metas[res->hits[i].seq].len);
        const uint32_t C = segment_rows(Lmax);

        /* take hits while their checkpoints and step buffers fit the budget; group the batch by class */
        std::vector<TraceJob> jobs;
        size_t ckf = 0, rawsteps = 0, end = done;
        while (end < nhits)
        {
            const HitRec &h = res->hits[end];
            const size_t need_ck = ck_floats(h, C);
            const uint64_t cap = step_cap(h, cap_mul);
            if (cap > 0xfffffff0ull) return dcp_error(RC_EFAIL, "path buffer of a hit exceeds 2^32 steps");
            if (!jobs.empty() && ((ckf + need_ck) * 4 + (rawsteps + cap) * sizeof(dcp_step) > budget)) break;
            jobs.push_back({h.seq, h.prof, ckf, rawsteps, (uint32_t)cap, 0});
            ckf += need_ck, rawsteps += cap;
            ++end;
        }
        const uint32_t nj = (uint32_t)jobs.size();
        stamp("batch chosen");
        /* order jobs by class so each launch sees a contiguous range; inside a class the longest sequences first (by
         * power-of-two bucket: the tail of a launch is one hit long), and inside a bucket by profile: hits of one profile
         * are then in flight together and share its emission tables in L2 (10 000 hits on 1000 profiles: forward rows
         * 1620 -> 1260 cycles with the tables resident) */
        auto bucket = [&](uint32_t j) { return 31 - __builtin_clz(std::max(1u, sq->metas[jobs[j].seq].len)); };
        std::vector<uint32_t> order(nj);
        for (uint32_t i = 0; i < nj; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
            const uint32_t cx = db->metas[jobs[x].prof].cls, cy = db->metas[jobs[y].prof].cls;
            if (cx != cy) return cx < cy;
            const int bx = bucket(x), by = bucket(y);
            if (bx != by) return bx > by;
            return jobs[x].prof < jobs[y].prof;
        });
        std::vector<TraceJob> sorted(nj);
        for (uint32_t i = 0; i < nj; ++i) sorted[i] = jobs[order[i]];

        /* class ranges */
        struct Range { uint32_t a, b, cls; };
        std::vector<Range> ranges;
        for (uint32_t x = 0; x < nj;)
        {
            uint32_t cls = db->metas[sorted[x].prof].cls, y = x;
            while (y < nj && db->metas[sorted[y].prof].cls == cls) ++y;
            ranges.push_back({x, y, cls});
            x = y;
        }

        TraceArgs base = {};
        base.emis = db->d_emis, base.trans = db->d_trans, base.metas = db->d_metas, base.seqs = sq->d_metas;
        base.total_recs = sq->total + sq->nseq, base.rows = d_rows, base.wcodes = d_wcodes, base.spec = d_spec;
        base.C = C;
        /* the trace kernels depend on (warps per pair, nodes per lane) only, not on the occupancy variant */
        auto launch_range = [&](const Range &r, cudaStream_t s, TraceArgs a, float *pool, size_t *used, bool dry) -> int {
            const dcp_class &kc = *dcp_class_at(r.cls);
            a.njobs = r.b - r.a;
            bool launched = false;
            cudaError_t e = cudaSuccess;
#define X(TW, Q, BPS, RATE)                                                                                       \
    if (!launched && kc.tw == TW && kc.q == Q)                                                                    \
    {                                                                                                             \
        e = launch_trace<TW, Q>(s, db->sm_count, a, pool, used, dry);                                             \
        launched = true;                                                                                          \
    }
            DCP_CLASS_TABLE(X)
#undef X
            if (!launched) return -1;
            return e == cudaSuccess ? 0 : 1;
        };
        size_t scratch_floats = 0;
        for (const Range &r : ranges)
            if (launch_range(r, nullptr, base, nullptr, &scratch_floats, true) < 0)
                return dcp_error(RC_EFAIL, "no trace kernel for this kernel class");

        stamp("jobs sorted");
        DevBuf b_jobs, b_ck, b_scr, b_raw, b_alt, b_n, b_off, b_err, b_cnt, b_steps;
        CU_TRY(b_jobs.alloc(nj * sizeof(TraceJob), db));
        CU_TRY(b_ck.alloc(ckf * sizeof(float), db));
        CU_TRY(b_scr.alloc(scratch_floats * sizeof(float), db));
        CU_TRY(b_raw.alloc(rawsteps * sizeof(dcp_step), db));
        CU_TRY(b_alt.alloc(nj * sizeof(float), db));
        CU_TRY(b_n.alloc(nj * sizeof(uint32_t), db));
        CU_TRY(b_off.alloc(nj * sizeof(uint64_t), db));
        CU_TRY(b_err.alloc(2 * sizeof(uint32_t) + 8 * sizeof(unsigned long long), db));
        CU_TRY(b_cnt.alloc(ranges.size() * sizeof(unsigned long long), db));
        stamp("buffers allocated");
        CU_TRY(cudaMemcpyAsync(b_jobs.p, sorted.data(), nj * sizeof(TraceJob), cudaMemcpyHostToDevice, st));
        CU_TRY(cudaMemsetAsync(b_err.p, 0, 2 * sizeof(uint32_t) + 8 * sizeof(unsigned long long), st));
        CU_TRY(cudaMemsetAsync(b_cnt.p, 0, ranges.size() * sizeof(unsigned long long), st));
        base.ckpt = b_ck.as<float>(), base.steps_raw = b_raw.as<dcp_step>(), base.errors = b_err.as<uint32_t>();
        base.prof = reinterpret_cast<unsigned long long *>(b_err.as<uint32_t>() + 2);
        /* one launch per kernel class, side by side on the main and side streams: a class often holds a handful of
         * hits, and a trace launch lasts as long as its longest sequence however few hits it has */
        StreamFan fan(db);
        CU_TRY(fan.fork());
        size_t used = 0;
        for (size_t ri = 0; ri < ranges.size(); ++ri)
        {
            const Range &r = ranges[ri];
            TraceArgs a = base;
            a.jobs = b_jobs.as<TraceJob>() + r.a, a.nsteps = b_n.as<uint32_t>() + r.a, a.alt_out = b_alt.as<float>() + r.a;
            a.counter = b_cnt.as<unsigned long long>() + ri;
            if (launch_range(r, fan.next(), a, b_scr.as<float>(), &used, false) != 0)
            {
                CU_TRY(cudaGetLastError());
                return dcp_error(RC_EFAIL, "trace kernel launch failed");
            }
            (*launches)++;
        }
        CU_TRY(fan.join());
        CU_TRY(cudaGetLastError());
        stamp("kernels queued");
        std::vector<uint32_t> ns(nj);
        std::vector<float> talt(nj);
        uint32_t nerr[2] = {0, 0};
        CU_TRY(cudaMemcpyAsync(ns.data(), b_n.p, nj * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(talt.data(), b_alt.p, nj * sizeof(float), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(nerr, b_err.p, sizeof nerr, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        stamp("kernels done, counts on the host");
        if (DCP_TRACE_PROF)
        {
            unsigned long long pr[8];
            CU_TRY(cudaMemcpy(pr, base.prof, sizeof pr, cudaMemcpyDeviceToHost));
            const double n = (double)std::max(1ull, pr[7]);
            fprintf(stderr, "[dcp_trace prof] hits %llu: per hit kcycles fwd rows %.0f, bwd rows %.0f, walk %.0f, whole hit %.0f; "
                            "steps %.0f, segment passes %.1f of which with every cell %.1f\n",
                    pr[7], pr[0] / n / 1e3, pr[1] / n / 1e3, pr[2] / n / 1e3, pr[3] / n / 1e3, pr[4] / n, pr[6] / n, pr[5] / n);
        }
        if (nerr[0]) return dcp_error(RC_EFAIL, "traceback walked off the DP matrix");
        if (nerr[1])
        {
            /* a path longer than 2 (L + M) + 64 steps (many passes through the core): retry the batch with room */
            if (cap_mul >= 4096) return dcp_error(RC_EFAIL, "traceback path does not fit its buffer");
            cap_mul *= 8;
            continue;
        }
        for (uint32_t i = 0; i < nj; ++i)
        {
            size_t hit = done + order[i];
            if (memcmp(&talt[i], &score_alt[hit], sizeof(float)) != 0)
                return dcp_error(RC_EFAIL, "trace pass and score pass disagree on the alt log-likelihood");
        }
        /* the paths are packed on the device in hit order and land in the result's step array in one copy */
        std::vector<uint32_t> inv(nj);
        for (uint32_t i = 0; i < nj; ++i) inv[order[i]] = i;
        std::vector<uint64_t> off(nj);
        uint64_t tot = 0;
        const size_t steps0 = res->steps.size();
        for (uint32_t h = 0; h < nj; ++h)
        {
            const uint32_t i = inv[h];
            HitRec &hr = res->hits[done + h];
            hr.step_off = steps0 + tot, hr.nsteps = ns[i];
            off[i] = tot, tot += ns[i];
        }
        res->steps.resize(steps0 + tot);
        CU_TRY(b_steps.alloc(std::max<uint64_t>(tot, 1) * sizeof(dcp_step), db));
        CU_TRY(cudaMemcpyAsync(b_off.p, off.data(), nj * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        k_pack<<<nj, 128, 0, st>>>(b_jobs.as<TraceJob>(), nj, b_n.as<uint32_t>(), b_off.as<uint64_t>(),
                                   b_raw.as<dcp_step>(), b_steps.as<dcp_step>());
        (*launches)++;
        CU_TRY(cudaMemcpyAsync(res->steps.data() + steps0, b_steps.p, tot * sizeof(dcp_step), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        CU_TRY(cudaGetLastError());
        stamp("paths on the host");
        done = end;
        cap_mul = 1;
        stamp("paths appended");
    }
    return RC_OK;
}
