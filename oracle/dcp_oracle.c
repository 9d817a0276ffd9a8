/*
 * oracle/dcp_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference's scan hot path (deciphon-old + the imm 2.0.3
 * DP library it links, which is NOT in /root/reference and cannot be built here).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/--impl reference
 * legs may load this library, and only as the checker / reported CPU baseline.
 * The product (deciphon-old_b200/csrc) never links or calls anything in oracle/.
 *
 * Parity pinning: the golden vectors of /root/reference/test/protein_profile.c
 * (null loglik :41, alt loglik uniform :65 / occupancy :157, path shapes :43-77,
 * ten decoded codons :83-102) are reproduced by this file -- see
 * tests/test_oracle_kat.py.  Tie-break order between incoming transitions, the
 * codon iteration order of decode and the abc name string are NOT pinned by any
 * runnable reference test ("parity unpinned" for those three items).
 *
 * Build flavours: -DORC_DOUBLE => DP arithmetic in double (the reference's
 * IMM_DOUBLE_PRECISION=On CI leg); default => float (reference default).
 * Model parameters (codon tables, frame emission tables, entry distribution,
 * length-dependent specials) are always evaluated in double and rounded ONCE to
 * the DP type.
 *
 * Each function cites the reference file:line it follows (paths relative to
 * /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef ORC_DOUBLE
typedef double ofloat;
#else
typedef float ofloat;
#endif

#define NEGINF (-INFINITY)
#define NTAB 1364 /* 4 + 16 + 64 + 256 + 1024 strings of 1..5 nt */
#define NAMINO 20
#define NTRANS 7

/* include/deciphon/core/rc.h:4-15 */
enum { RC_OK, RC_END, RC_EFAIL, RC_EINVAL, RC_EIO, RC_ENOMEM, RC_EPARSE };

/* include/deciphon/model/protein_state.h:7-21 */
enum {
    ST_MATCH = 0 << 14,
    ST_INSERT = 1 << 14,
    ST_DELETE = 2 << 14,
    ST_EXT = 3 << 14,
    ST_R = ST_EXT | 0,
    ST_S = ST_EXT | 1,
    ST_N = ST_EXT | 2,
    ST_B = ST_EXT | 3,
    ST_E = ST_EXT | 4,
    ST_J = ST_EXT | 5,
    ST_C = ST_EXT | 6,
    ST_T = ST_EXT | 7,
};

static const int tab_off[6] = {0, 0, 4, 20, 84, 340};

/* ------------------------------------------------------------------ */
/* A.1 primitives: imm_rnd (splitmix64 seeding + xoshiro256+), lprob helpers */
/* ------------------------------------------------------------------ */
struct rnd { uint64_t s[4]; };

static uint64_t splitmix_next(uint64_t *x)
{
    uint64_t z = (*x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

static struct rnd rnd_init(uint64_t seed)
{
    struct rnd r;
    for (int i = 0; i < 4; ++i) r.s[i] = splitmix_next(&seed);
    return r;
}

static uint64_t rnd_u64(struct rnd *r)
{
    uint64_t *s = r->s;
    uint64_t res = s[0] + s[3];
    uint64_t t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = (s[3] << 45) | (s[3] >> 19);
    return res;
}

static double rnd_dbl(struct rnd *r) { return (double)(rnd_u64(r) >> 11) * 0x1.0p-53; }

static double lse2(double a, double b)
{
    if (a == NEGINF) return b;
    if (b == NEGINF) return a;
    double m = a > b ? a : b;
    double d = a > b ? b - a : a - b;
    return m + log1p(exp(d));
}

static double lse_n(int n, const double *x)
{
    double m = NEGINF;
    for (int i = 0; i < n; ++i)
        if (x[i] > m) m = x[i];
    if (m == NEGINF) return NEGINF;
    double s = 0;
    for (int i = 0; i < n; ++i) s += exp(x[i] - m);
    return m + log(s);
}

/* imm_lprob_sample + imm_lprob_normalize as used by protein_profile.c:268-269 */
static void lprob_sample(struct rnd *r, int n, double *a)
{
    for (int i = 0; i < n; ++i) a[i] = log(rnd_dbl(r));
}
static void lprob_normalize(int n, double *a)
{
    double z = lse_n(n, a);
    for (int i = 0; i < n; ++i) a[i] -= z;
}

/* ------------------------------------------------------------------ */
/* A.2 genetic code (NCBI table 1) and setup_nuclt_dist                 */
/* ------------------------------------------------------------------ */
static const char AMINO[] = "ACDEFGHIKLMNPQRSTVWY"; /* protein_h3reader.c:83-102 */
static const char NUC[] = "ACGT";
static const char GC_AA[] = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";

static int nuc_idx(char c)
{
    switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    }
    return -1;
}

/* amino acid letter of codon (a,b,c in ACGT indices) */
static char gc_aa(int a, int b, int c)
{
    static const int to_tcag[4] = {2, 1, 3, 0}; /* A,C,G,T -> position in TCAG */
    return GC_AA[to_tcag[a] * 16 + to_tcag[b] * 4 + to_tcag[c]];
}

struct ndist {
    double nucltp[4];   /* imm_nuclt_lprob */
    double codonm[125]; /* imm_codon_marg, index a*25+b*5+c, 4 = any */
};

/* src/model/protein_model.c:342-408 (codon_lprob, nuclt_lprob, setup_nuclt_dist) */
static void setup_nuclt_dist(struct ndist *d, const double amino_lprobs[NAMINO])
{
    int count[128] = {0};
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b)
            for (int c = 0; c < 4; ++c) count[(int)gc_aa(a, b, c)]++;

    double aalp[128];
    for (int i = 0; i < 128; ++i) aalp[i] = NEGINF; /* stops ('*') stay -inf */
    for (int i = 0; i < NAMINO; ++i)
        aalp[(int)AMINO[i]] = amino_lprobs[i] - log((double)count[(int)AMINO[i]]);

    double cod[64];
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b)
            for (int c = 0; c < 4; ++c) cod[a * 16 + b * 4 + c] = aalp[(int)gc_aa(a, b, c)];
    lprob_normalize(64, cod); /* imm_codon_lprob_normalize, :404 */

    /* nuclt_lprob, :342-359 */
    double acc[4][64 * 3];
    int nacc[4] = {0, 0, 0, 0};
    const double l3 = log(3.0);
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b)
            for (int c = 0; c < 4; ++c) {
                double v = cod[a * 16 + b * 4 + c] - l3;
                acc[a][nacc[a]++] = v;
                acc[b][nacc[b]++] = v;
                acc[c][nacc[c]++] = v;
            }
    for (int x = 0; x < 4; ++x) d->nucltp[x] = lse_n(nacc[x], acc[x]);

    /* imm_codon_marg, :407 */
    for (int a = 0; a < 5; ++a)
        for (int b = 0; b < 5; ++b)
            for (int c = 0; c < 5; ++c) {
                double t[64];
                int n = 0;
                for (int x = 0; x < 4; ++x)
                    for (int y = 0; y < 4; ++y)
                        for (int z = 0; z < 4; ++z)
                            if ((a == 4 || a == x) && (b == 4 || b == y) && (c == 4 || c == z))
                                t[n++] = cod[x * 16 + y * 4 + z];
                d->codonm[a * 25 + b * 5 + c] = lse_n(n, t);
            }
}

/* ------------------------------------------------------------------ */
/* A.4 imm_frame_state emission: lprob of a 1..5-nt string              */
/* ------------------------------------------------------------------ */
#define MG(a, b, c) (mg[(a)*25 + (b)*5 + (c)])
#define ANY 4

static double frame_lprob(const int *z, int n, const double nucltp[4], const double mg[125],
                          double eps)
{
    const double le = log(eps), l1 = log(1 - eps);
    const double *B = nucltp;
    double t[32];
    int k = 0;
    if (n == 1) {
        t[0] = MG(z[0], ANY, ANY);
        t[1] = MG(ANY, z[0], ANY);
        t[2] = MG(ANY, ANY, z[0]);
        return 2 * le + 2 * l1 - log(3.0) + lse_n(3, t);
    }
    if (n == 2) {
        t[0] = MG(ANY, z[0], z[1]);
        t[1] = MG(z[0], ANY, z[1]);
        t[2] = MG(z[0], z[1], ANY);
        double v0 = log(2.0) + le + 3 * l1 - log(3.0) + lse_n(3, t);
        t[0] = B[z[1]] + MG(z[0], ANY, ANY);
        t[1] = B[z[1]] + MG(ANY, z[0], ANY);
        t[2] = B[z[1]] + MG(ANY, ANY, z[0]);
        t[3] = B[z[0]] + MG(z[1], ANY, ANY);
        t[4] = B[z[0]] + MG(ANY, z[1], ANY);
        t[5] = B[z[0]] + MG(ANY, ANY, z[1]);
        double v1 = 3 * le + l1 - log(3.0) + lse_n(6, t);
        return lse2(v0, v1);
    }
    if (n == 3) {
        double v0 = 4 * l1 + MG(z[0], z[1], z[2]);
        k = 0;
        for (int i = 0; i < 3; ++i) {
            int r[2], m = 0;
            for (int j = 0; j < 3; ++j)
                if (j != i) r[m++] = z[j];
            t[k++] = B[z[i]] + MG(ANY, r[0], r[1]);
            t[k++] = B[z[i]] + MG(r[0], ANY, r[1]);
            t[k++] = B[z[i]] + MG(r[0], r[1], ANY);
        }
        double v1 = log(4.0) + 2 * le + 2 * l1 - log(9.0) + lse_n(9, t);
        k = 0;
        for (int i = 0; i < 3; ++i) {
            double o = 0;
            for (int j = 0; j < 3; ++j)
                if (j != i) o += B[z[j]];
            t[k++] = o + MG(z[i], ANY, ANY);
            t[k++] = o + MG(ANY, z[i], ANY);
            t[k++] = o + MG(ANY, ANY, z[i]);
        }
        double v2 = 4 * le - log(9.0) + lse_n(9, t);
        double v[3] = {v0, v1, v2};
        return lse_n(3, v);
    }
    if (n == 4) {
        k = 0;
        for (int i = 0; i < 4; ++i) {
            int r[3], m = 0;
            for (int j = 0; j < 4; ++j)
                if (j != i) r[m++] = z[j];
            t[k++] = B[z[i]] + MG(r[0], r[1], r[2]);
        }
        double v0 = le + 3 * l1 - log(2.0) + lse_n(4, t);
        k = 0;
        for (int i = 0; i < 4; ++i)
            for (int j = i + 1; j < 4; ++j) {
                int r[2], m = 0;
                for (int q = 0; q < 4; ++q)
                    if (q != i && q != j) r[m++] = z[q];
                double o = B[z[i]] + B[z[j]];
                t[k++] = o + MG(ANY, r[0], r[1]);
                t[k++] = o + MG(r[0], ANY, r[1]);
                t[k++] = o + MG(r[0], r[1], ANY);
            }
        double v1 = 3 * le + l1 - log(9.0) + lse_n(18, t);
        return lse2(v0, v1);
    }
    /* n == 5 */
    k = 0;
    for (int i = 0; i < 5; ++i)
        for (int j = i + 1; j < 5; ++j) {
            int r[3], m = 0;
            for (int q = 0; q < 5; ++q)
                if (q != i && q != j) r[m++] = z[q];
            t[k++] = B[z[i]] + B[z[j]] + MG(r[0], r[1], r[2]);
        }
    return 2 * le + 2 * l1 - log(10.0) + lse_n(10, t);
}

/* tabulate all 1364 strings; code = tab_off[len] + base-4 value, first base most significant */
static void frame_table(const struct ndist *d, double eps, ofloat out[NTAB])
{
    for (int n = 1; n <= 5; ++n) {
        int cnt = 1 << (2 * n);
        for (int v = 0; v < cnt; ++v) {
            int z[5];
            for (int i = 0; i < n; ++i) z[i] = (v >> (2 * (n - 1 - i))) & 3;
            out[tab_off[n] + v] = (ofloat)frame_lprob(z, n, d->nucltp, d->codonm, eps);
        }
    }
}

/* ------------------------------------------------------------------ */
/* profile                                                              */
/* ------------------------------------------------------------------ */
struct orc_profile {
    int M;
    int entry_dist; /* 1 = uniform, 2 = occupancy (include/deciphon/model/entry_dist.h) */
    double eps;
    struct ndist null_nd, ins_nd, *match_nd; /* [M] */
    ofloat emN[NTAB], emI[NTAB], *emM;       /* emM[M][NTAB] */
    ofloat *trans;                            /* [(M+1)][7]: MM MI MD IM II DM DD */
    ofloat *entry;                            /* [M] B->M_k */
};

void orc_profile_del(struct orc_profile *p)
{
    if (!p) return;
    free(p->match_nd);
    free(p->emM);
    free(p->trans);
    free(p->entry);
    free(p);
}

static struct orc_profile *profile_alloc(int M)
{
    struct orc_profile *p = calloc(1, sizeof *p);
    p->M = M;
    p->match_nd = calloc(M, sizeof *p->match_nd);
    p->emM = malloc(sizeof(ofloat) * NTAB * (size_t)M);
    p->trans = malloc(sizeof(ofloat) * NTRANS * (size_t)(M + 1));
    p->entry = malloc(sizeof(ofloat) * (size_t)M);
    return p;
}

/* src/model/protein_model.c:258-283 (calculate_occupancy), :410-439 (setup_entry_trans) */
static void entry_scores(int M, int entry_dist, const double *tr /* [(M+1)][7] */, ofloat *entry)
{
    if (entry_dist == 1) {
        double Md = (double)M;
        double cost = log(2.0 / (Md * (Md + 1))) * Md; /* sic, :414-415 */
        for (int i = 0; i < M; ++i) entry[i] = (ofloat)cost;
        return;
    }
    if (M < 1) return;
    double *locc = malloc(sizeof(double) * (size_t)M);
    const double *t = tr;
    locc[0] = lse2(t[1] /*MI*/, t[0] /*MM*/);
    for (int i = 1; i < M; ++i) {
        t += NTRANS;
        double v0 = locc[i - 1] + lse2(t[0], t[1]);
        double v1 = log1p(-exp(locc[i - 1])) + t[5] /*DM*/;
        locc[i] = lse2(v0, v1);
    }
    double logZ = NEGINF;
    for (int i = 0; i < M; ++i) logZ = lse2(logZ, locc[i] + log((double)(M - i)));
    for (int i = 0; i < M; ++i) entry[i] = (ofloat)(locc[i] - logZ);
    free(locc);
}

/*
 * Build from model-level inputs, as protein_model_init/add_node/add_trans do
 * (src/model/protein_model.c:49-96,105-137).  In the float flavour the inputs are
 * first rounded to float, as the reference's imm_float arguments would be.
 */
struct orc_profile *orc_profile_build(int M, int entry_dist, double eps, const double *null_lprobs,
                                      const double *match_lprobs /* [M][20] */,
                                      const double *trans /* [(M+1)][7] */)
{
    struct orc_profile *p = profile_alloc(M);
    p->entry_dist = entry_dist;
    p->eps = eps;
    double nl[NAMINO], zero[NAMINO] = {0};
    for (int i = 0; i < NAMINO; ++i) nl[i] = (double)(ofloat)null_lprobs[i];
    setup_nuclt_dist(&p->null_nd, nl);   /* :122 */
    setup_nuclt_dist(&p->ins_nd, zero);  /* :126-127 */
    frame_table(&p->null_nd, eps, p->emN);
    frame_table(&p->ins_nd, eps, p->emI);
    for (int k = 0; k < M; ++k) {
        double lodds[NAMINO];
        for (int i = 0; i < NAMINO; ++i) /* :60-62, in imm_float */
            lodds[i] = (double)(ofloat)((ofloat)match_lprobs[k * NAMINO + i] - (ofloat)nl[i]);
        setup_nuclt_dist(&p->match_nd[k], lodds);
        frame_table(&p->match_nd[k], eps, p->emM + (size_t)k * NTAB);
    }
    double *tr = malloc(sizeof(double) * NTRANS * (M + 1));
    for (int i = 0; i < NTRANS * (M + 1); ++i) {
        p->trans[i] = (ofloat)trans[i];
        tr[i] = (double)p->trans[i];
    }
    entry_scores(M, entry_dist, tr, p->entry);
    free(tr);
    return p;
}

/* src/model/protein_profile.c:259-304 (protein_profile_sample) */
struct orc_profile *orc_profile_sample(unsigned seed, int M, int entry_dist, double eps)
{
    struct rnd r = rnd_init(seed);
    double null_lp[NAMINO];
    lprob_sample(&r, NAMINO, null_lp);
    lprob_normalize(NAMINO, null_lp);
    double *match = malloc(sizeof(double) * NAMINO * M);
    for (int k = 0; k < M; ++k) {
        lprob_sample(&r, NAMINO, match + k * NAMINO);
        lprob_normalize(NAMINO, match + k * NAMINO);
    }
    double *tr = malloc(sizeof(double) * NTRANS * (M + 1));
    for (int i = 0; i <= M; ++i) {
        double *t = tr + i * NTRANS;
        lprob_sample(&r, NTRANS, t);
        if (i == 0) t[6] = NEGINF;
        if (i == M) {
            t[2] = NEGINF;
            t[6] = NEGINF;
        }
        lprob_normalize(NTRANS, t);
    }
    struct orc_profile *p = orc_profile_build(M, entry_dist, eps, null_lp, match, tr);
    free(match);
    free(tr);
    return p;
}

/* the three model-level input arrays of a sampled profile (to feed the product's builder) */
void orc_sample_inputs(unsigned seed, int M, double *null_lp, double *match, double *tr)
{
    struct rnd r = rnd_init(seed);
    lprob_sample(&r, NAMINO, null_lp);
    lprob_normalize(NAMINO, null_lp);
    for (int k = 0; k < M; ++k) {
        lprob_sample(&r, NAMINO, match + k * NAMINO);
        lprob_normalize(NAMINO, match + k * NAMINO);
    }
    for (int i = 0; i <= M; ++i) {
        double *t = tr + i * NTRANS;
        lprob_sample(&r, NTRANS, t);
        if (i == 0) t[6] = NEGINF;
        if (i == M) {
            t[2] = NEGINF;
            t[6] = NEGINF;
        }
        lprob_normalize(NTRANS, t);
    }
}

/* Import DP-level numbers produced elsewhere (the product's host builder), so that DP
 * parity can be checked on bit-identical inputs.  Arrays are float32/float64 by flavour. */
struct orc_profile *orc_profile_import(int M, double eps, const ofloat *emM, const ofloat *emI,
                                       const ofloat *emN, const ofloat *trans, const ofloat *entry,
                                       const double *null_nd /* 4+125 */,
                                       const double *ins_nd /* 4+125 */,
                                       const double *match_nd /* [M][4+125] */)
{
    struct orc_profile *p = profile_alloc(M);
    p->eps = eps;
    memcpy(p->emM, emM, sizeof(ofloat) * NTAB * (size_t)M);
    memcpy(p->emI, emI, sizeof(ofloat) * NTAB);
    memcpy(p->emN, emN, sizeof(ofloat) * NTAB);
    memcpy(p->trans, trans, sizeof(ofloat) * NTRANS * (size_t)(M + 1));
    memcpy(p->entry, entry, sizeof(ofloat) * (size_t)M);
    if (null_nd) {
        memcpy(p->null_nd.nucltp, null_nd, 4 * sizeof(double));
        memcpy(p->null_nd.codonm, null_nd + 4, 125 * sizeof(double));
    }
    if (ins_nd) {
        memcpy(p->ins_nd.nucltp, ins_nd, 4 * sizeof(double));
        memcpy(p->ins_nd.codonm, ins_nd + 4, 125 * sizeof(double));
    }
    if (match_nd)
        for (int k = 0; k < M; ++k) {
            memcpy(p->match_nd[k].nucltp, match_nd + (size_t)k * 129, 4 * sizeof(double));
            memcpy(p->match_nd[k].codonm, match_nd + (size_t)k * 129 + 4, 125 * sizeof(double));
        }
    return p;
}

int orc_profile_M(const struct orc_profile *p) { return p->M; }
int orc_float_size(void) { return (int)sizeof(ofloat); }
const ofloat *orc_profile_emM(const struct orc_profile *p) { return p->emM; }
const ofloat *orc_profile_emI(const struct orc_profile *p) { return p->emI; }
const ofloat *orc_profile_emN(const struct orc_profile *p) { return p->emN; }
const ofloat *orc_profile_trans(const struct orc_profile *p) { return p->trans; }
const ofloat *orc_profile_entry(const struct orc_profile *p) { return p->entry; }
void orc_profile_ndist(const struct orc_profile *p, int which /* -2 null, -1 insert, k>=0 match */,
                       double out[129])
{
    const struct ndist *d = which == -2 ? &p->null_nd : which == -1 ? &p->ins_nd : &p->match_nd[which];
    memcpy(out, d->nucltp, 4 * sizeof(double));
    memcpy(out + 4, d->codonm, 125 * sizeof(double));
}

/* ------------------------------------------------------------------ */
/* protein_profile_setup: length-dependent specials                    */
/* src/model/protein_profile.c:155-216                                 */
/* ------------------------------------------------------------------ */
struct xtrans {
    ofloat NN, CC, JJ, NB, CT, JB, RR, EJ, EC;
    ofloat ET, ECC, EB, EJJ; /* the pre-summed E->T, E->C, E->B, E->J of :206-212 */
};

static int specials(unsigned L_, int multi_hits, int hmmer3_compat, struct xtrans *t)
{
    if (L_ == 0) return RC_EINVAL; /* :158 */
    double L = (double)(ofloat)L_;
    double q = 0.0, log_q = NEGINF;
    if (multi_hits) {
        q = 0.5;
        log_q = log(0.5);
    }
    double lp = log(L) - log(L + 2 + q / (1 - q));
    double l1p = log(2 + q / (1 - q)) - log(L + 2 + q / (1 - q));
    double lr = log(L) - log(L + 1);
    t->NN = t->CC = t->JJ = (ofloat)lp;
    t->NB = t->CT = t->JB = (ofloat)l1p;
    t->RR = (ofloat)lr;
    t->EJ = (ofloat)log_q;
    t->EC = (ofloat)log(1 - q);
    if (hmmer3_compat) t->NN = t->CC = t->JJ = (ofloat)0;
    t->ET = t->EC + t->CT; /* imm_float additions, :206-212 */
    t->ECC = t->EC + t->CC;
    t->EB = t->EJ + t->JB;
    t->EJJ = t->EJ + t->JJ;
    return RC_OK;
}

int orc_specials(unsigned L, int multi_hits, int hmmer3_compat, ofloat out[13])
{
    struct xtrans t;
    int rc = specials(L, multi_hits, hmmer3_compat, &t);
    if (rc) return rc;
    ofloat v[13] = {t.NN, t.CC, t.JJ, t.NB, t.CT, t.JB, t.RR, t.EJ, t.EC, t.ET, t.ECC, t.EB, t.EJJ};
    memcpy(out, v, sizeof v);
    return RC_OK;
}

/* encode ASCII ACGT -> 0..3; returns RC_EINVAL on any other symbol */
static int encode_seq(const char *seq, int L, uint8_t *out)
{
    for (int i = 0; i < L; ++i) {
        int b = nuc_idx(seq[i]);
        if (b < 0) return RC_EINVAL;
        out[i] = (uint8_t)b;
    }
    return RC_OK;
}

static inline int code_of(const uint8_t *s, int start, int len)
{
    int v = 0;
    for (int i = 0; i < len; ++i) v = (v << 2) | s[start + i];
    return tab_off[len] + v;
}

/* ------------------------------------------------------------------ */
/* A.5 generic interpreter, imm's start-position form                  */
/* W[r][s][l] = Tin(s,r) + e_s(seq[r:r+l]); Tin = first max over        */
/* (incoming transition order, source length ascending), strict '>'.   */
/* ------------------------------------------------------------------ */
struct gtrans { int src; ofloat score; };
struct gstate {
    int id;        /* protein_state id */
    int minlen, maxlen; /* 0,0 mute; 1,5 frame */
    const ofloat *em;   /* NTAB table or NULL */
    int ninc;
    struct gtrans inc[4];
    int *xinc_src; ofloat *xinc_score; int nxinc; /* E: long incoming list */
};

struct gmodel {
    int n;
    struct gstate *st;
    int *order; /* per-row evaluation order */
    int start, end;
};

static void gstate_init(struct gstate *s, int id, int frame, const ofloat *em)
{
    memset(s, 0, sizeof *s);
    s->id = id;
    s->minlen = frame ? 1 : 0;
    s->maxlen = frame ? 5 : 0;
    s->em = em;
}
static void ginc(struct gstate *s, int src, ofloat sc)
{
    s->inc[s->ninc].src = src;
    s->inc[s->ninc].score = sc;
    s->ninc++;
}

enum { GS_S, GS_N, GS_B, GS_E, GS_J, GS_C, GS_T, GS_CORE };
#define GM(k) (GS_CORE + 3 * ((k)-1) + 0) /* k = 1..M */
#define GI(k) (GS_CORE + 3 * ((k)-1) + 1)
#define GD(k) (GS_CORE + 3 * ((k)-1) + 2)

/*
 * Alt model graph: src/model/protein_model.c:316-340 (specials topology), :460-500 (core),
 * :410-458 (entry/exit).  CANONICAL incoming order (imm's own order is unpinned):
 *   N: S,N   B: S,N,J,E   J: E,J   C: E,C   T: E,C
 *   M_k: B, M_{k-1}, I_{k-1}, D_{k-1}    I_k: M_k, I_k    D_k: M_{k-1}, D_{k-1}
 *   E: M_1, M_2, D_2, M_3, D_3, ...
 */
static struct gmodel *gmodel_alt(const struct orc_profile *p, const struct xtrans *x)
{
    int M = p->M;
    struct gmodel *g = calloc(1, sizeof *g);
    g->n = GS_CORE + 3 * M;
    g->st = calloc(g->n, sizeof *g->st);
    gstate_init(&g->st[GS_S], ST_S, 0, NULL);
    gstate_init(&g->st[GS_N], ST_N, 1, p->emN);
    gstate_init(&g->st[GS_B], ST_B, 0, NULL);
    gstate_init(&g->st[GS_E], ST_E, 0, NULL);
    gstate_init(&g->st[GS_J], ST_J, 1, p->emN);
    gstate_init(&g->st[GS_C], ST_C, 1, p->emN);
    gstate_init(&g->st[GS_T], ST_T, 0, NULL);
    ginc(&g->st[GS_N], GS_S, x->NN);
    ginc(&g->st[GS_N], GS_N, x->NN);
    ginc(&g->st[GS_B], GS_S, x->NB);
    ginc(&g->st[GS_B], GS_N, x->NB);
    ginc(&g->st[GS_B], GS_J, x->JB);
    ginc(&g->st[GS_B], GS_E, x->EB);
    ginc(&g->st[GS_J], GS_E, x->EJJ);
    ginc(&g->st[GS_J], GS_J, x->JJ);
    ginc(&g->st[GS_C], GS_E, x->ECC);
    ginc(&g->st[GS_C], GS_C, x->CC);
    ginc(&g->st[GS_T], GS_E, x->ET);
    ginc(&g->st[GS_T], GS_C, x->CT);
    for (int k = 1; k <= M; ++k) {
        gstate_init(&g->st[GM(k)], ST_MATCH | k, 1, p->emM + (size_t)(k - 1) * NTAB);
        gstate_init(&g->st[GI(k)], ST_INSERT | k, 1, p->emI);
        gstate_init(&g->st[GD(k)], ST_DELETE | k, 0, NULL);
        ginc(&g->st[GM(k)], GS_B, p->entry[k - 1]);
        if (k >= 2) {
            const ofloat *t = p->trans + (size_t)(k - 1) * NTRANS; /* trans[k-1] links k-1 -> k */
            ginc(&g->st[GM(k)], GM(k - 1), t[0]);
            ginc(&g->st[GM(k)], GI(k - 1), t[3]);
            ginc(&g->st[GM(k)], GD(k - 1), t[5]);
            ginc(&g->st[GD(k)], GM(k - 1), t[2]);
            ginc(&g->st[GD(k)], GD(k - 1), t[6]);
        }
        if (k <= M - 1) {
            const ofloat *t = p->trans + (size_t)k * NTRANS; /* trans[k] holds M_k->I_k, I_k->I_k */
            ginc(&g->st[GI(k)], GM(k), t[1]);
            ginc(&g->st[GI(k)], GI(k), t[4]);
        }
    }
    struct gstate *E = &g->st[GS_E];
    E->nxinc = 2 * M - 1;
    E->xinc_src = malloc(sizeof(int) * E->nxinc);
    E->xinc_score = malloc(sizeof(ofloat) * E->nxinc);
    int n = 0;
    for (int k = 1; k <= M; ++k) {
        E->xinc_src[n] = GM(k);
        E->xinc_score[n++] = (ofloat)0;
        if (k >= 2) {
            E->xinc_src[n] = GD(k);
            E->xinc_score[n++] = (ofloat)0;
        }
    }
    /* per-row order: every mute source before its same-row dependants */
    g->order = malloc(sizeof(int) * g->n);
    n = 0;
    g->order[n++] = GS_S;
    for (int k = 1; k <= M; ++k) g->order[n++] = GD(k);
    g->order[n++] = GS_E;
    g->order[n++] = GS_J;
    g->order[n++] = GS_C;
    g->order[n++] = GS_N;
    g->order[n++] = GS_B;
    g->order[n++] = GS_T;
    for (int k = 1; k <= M; ++k) {
        g->order[n++] = GM(k);
        g->order[n++] = GI(k);
    }
    g->start = GS_S;
    g->end = GS_T;
    return g;
}

/* imm_dp_change_trans x12 (protein_profile.c:190-214): the length-dependent scores of an existing graph */
static void gmodel_alt_specials(struct gmodel *g, const struct xtrans *x)
{
    g->st[GS_N].inc[0].score = x->NN, g->st[GS_N].inc[1].score = x->NN;
    g->st[GS_B].inc[0].score = x->NB, g->st[GS_B].inc[1].score = x->NB;
    g->st[GS_B].inc[2].score = x->JB, g->st[GS_B].inc[3].score = x->EB;
    g->st[GS_J].inc[0].score = x->EJJ, g->st[GS_J].inc[1].score = x->JJ;
    g->st[GS_C].inc[0].score = x->ECC, g->st[GS_C].inc[1].score = x->CC;
    g->st[GS_T].inc[0].score = x->ET, g->st[GS_T].inc[1].score = x->CT;
}

static struct gmodel *gmodel_null(const struct orc_profile *p, const struct xtrans *x)
{
    struct gmodel *g = calloc(1, sizeof *g);
    g->n = 1;
    g->st = calloc(1, sizeof *g->st);
    gstate_init(&g->st[0], ST_R, 1, p->emN);
    ginc(&g->st[0], 0, x->RR);
    g->order = calloc(1, sizeof(int));
    g->start = g->end = 0;
    return g;
}

static void gmodel_del(struct gmodel *g)
{
    for (int i = 0; i < g->n; ++i) {
        free(g->st[i].xinc_src);
        free(g->st[i].xinc_score);
    }
    free(g->st);
    free(g->order);
    free(g);
}

/* returns RC; fills loglik and (optionally) the path */
static int gviterbi(const struct gmodel *g, const uint8_t *seq, int L, ofloat *loglik,
                    uint16_t *step_state, uint8_t *step_len, int *nsteps, int max_steps)
{
    int n = g->n;
    size_t rows = (size_t)L + 1;
    /* DP matrix and backpointers are per-thread and grow-only, like the imm_task a scan thread keeps and
     * resets between pairs (scan_thread.c:40-55); every entry that is read has been written in the same call
     * (W[r'][s][l] is read at row r' + l <= L and written at row r' iff r' + l <= L), so nothing is cleared */
    static _Thread_local struct { ofloat *W; int32_t *bt; uint8_t *bl; size_t cap; } scr;
    if (scr.cap < rows * n) {
        free(scr.W); free(scr.bt); free(scr.bl);
        scr.cap = rows * n + rows * n / 4;
        scr.W = malloc(sizeof(ofloat) * scr.cap * 6);
        scr.bt = malloc(sizeof(int32_t) * scr.cap);
        scr.bl = malloc(scr.cap);
        if (!scr.W || !scr.bt || !scr.bl) {
            free(scr.W); free(scr.bt); free(scr.bl);
            scr.W = NULL, scr.bt = NULL, scr.bl = NULL, scr.cap = 0;
            return RC_ENOMEM;
        }
    }
    ofloat *W = scr.W;
    int32_t *bt = scr.bt;
    uint8_t *bl = scr.bl;
#define WW(r, s, l) W[((size_t)(r)*n + (s)) * 6 + (l)]
    for (int r = 0; r <= L; ++r) {
        for (int oi = 0; oi < n; ++oi) {
            int s = g->order[oi];
            const struct gstate *st = &g->st[s];
            ofloat best = NEGINF;
            int32_t b_t = -2;
            uint8_t b_l = 0;
            if (s == g->start && r == 0) {
                best = (ofloat)0;
                b_t = -1;
            }
            int ninc = st->nxinc ? st->nxinc : st->ninc;
            for (int ti = 0; ti < ninc; ++ti) {
                int src = st->nxinc ? st->xinc_src[ti] : st->inc[ti].src;
                ofloat sc = st->nxinc ? st->xinc_score[ti] : st->inc[ti].score;
                const struct gstate *ss = &g->st[src];
                for (int l = ss->minlen; l <= ss->maxlen; ++l) {
                    if (r - l < 0) break;
                    ofloat v = WW(r - l, src, l) + sc;
                    if (v > best) {
                        best = v;
                        b_t = ti;
                        b_l = (uint8_t)l;
                    }
                }
            }
            bt[(size_t)r * n + s] = b_t;
            bl[(size_t)r * n + s] = b_l;
            for (int l = st->minlen; l <= st->maxlen; ++l) {
                if (r + l > L) break;
                ofloat e = l == 0 ? (ofloat)0 : st->em[code_of(seq, r, l)];
                WW(r, s, l) = best + e;
            }
        }
    }
    /* end: imm takes max over the end state's lengths at rows L-l */
    const struct gstate *es = &g->st[g->end];
    ofloat best = NEGINF;
    int bestl = -1;
    for (int l = es->minlen; l <= es->maxlen; ++l) {
        if (L - l < 0) break;
        ofloat v = WW(L - l, g->end, l);
        if (v > best) {
            best = v;
            bestl = l;
        }
    }
    *loglik = best;
    int rc = RC_OK;
    if (nsteps) {
        *nsteps = 0;
        if (bestl < 0) {
            rc = RC_EFAIL; /* no finite path */
        } else {
            int s = g->end, r = L - bestl, l = bestl, k = 0;
            for (;;) {
                if (k >= max_steps) { rc = RC_ENOMEM; break; }
                step_state[k] = (uint16_t)g->st[s].id;
                step_len[k] = (uint8_t)l;
                k++;
                int32_t t = bt[(size_t)r * n + s];
                if (t == -1) break;
                if (t == -2) { rc = RC_EFAIL; break; }
                const struct gstate *st = &g->st[s];
                int src = st->nxinc ? st->xinc_src[t] : st->inc[t].src;
                l = bl[(size_t)r * n + s];
                r -= l;
                s = src;
            }
            for (int i = 0; i < k / 2; ++i) {
                uint16_t a = step_state[i]; step_state[i] = step_state[k - 1 - i]; step_state[k - 1 - i] = a;
                uint8_t b = step_len[i]; step_len[i] = step_len[k - 1 - i]; step_len[k - 1 - i] = b;
            }
            *nsteps = k;
        }
    }
#undef WW
    return rc;
}

/* imm_dp_viterbi on the null dp (scan_thread.c:115) */
int orc_viterbi_null(const struct orc_profile *p, const char *seq, int L, int multi_hits,
                     int hmmer3_compat, ofloat *loglik, uint16_t *step_state, uint8_t *step_len,
                     int *nsteps, int max_steps)
{
    struct xtrans x;
    int rc = specials((unsigned)L, multi_hits, hmmer3_compat, &x);
    if (rc) return rc;
    uint8_t *s = malloc(L);
    if ((rc = encode_seq(seq, L, s))) { free(s); return rc; }
    struct gmodel *g = gmodel_null(p, &x);
    rc = gviterbi(g, s, L, loglik, step_state, step_len, nsteps, max_steps);
    gmodel_del(g);
    free(s);
    return rc;
}

/* imm_dp_viterbi on the alt dp (scan_thread.c:117), generic interpreter */
int orc_viterbi_alt(const struct orc_profile *p, const char *seq, int L, int multi_hits,
                    int hmmer3_compat, ofloat *loglik, uint16_t *step_state, uint8_t *step_len,
                    int *nsteps, int max_steps)
{
    struct xtrans x;
    int rc = specials((unsigned)L, multi_hits, hmmer3_compat, &x);
    if (rc) return rc;
    uint8_t *s = malloc(L);
    if ((rc = encode_seq(seq, L, s))) { free(s); return rc; }
    struct gmodel *g = gmodel_alt(p, &x);
    rc = gviterbi(g, s, L, loglik, step_state, step_len, nsteps, max_steps);
    gmodel_del(g);
    free(s);
    return rc;
}

/* ------------------------------------------------------------------ */
/* A.5 specialised recurrence (end-position form), scores only          */
/* ------------------------------------------------------------------ */
static inline ofloat fmx(ofloat a, ofloat b) { return a > b ? a : b; }

static ofloat null_score_fast(const struct orc_profile *p, const struct xtrans *x, const uint8_t *s,
                              int L)
{
    ofloat *V = malloc(sizeof(ofloat) * (L + 1));
    V[0] = NEGINF;
    for (int j = 1; j <= L; ++j) {
        ofloat best = NEGINF;
        for (int l = 1; l <= 5 && l <= j; ++l) {
            ofloat tin = (j - l == 0) ? (ofloat)0 : V[j - l] + x->RR;
            ofloat v = tin + p->emN[code_of(s, j - l, l)];
            best = fmx(best, v);
        }
        V[j] = best;
    }
    ofloat r = V[L];
    free(V);
    return r;
}

static ofloat alt_score_fast(const struct orc_profile *p, const struct xtrans *x, const uint8_t *s,
                             int L)
{
    int M = p->M;
    /* rings of the last 5 rows of Tin for M/I (per node) and N/J/C */
    ofloat *TM = malloc(sizeof(ofloat) * 5 * (M + 1));
    ofloat *TI = malloc(sizeof(ofloat) * 5 * (M + 1));
    ofloat *VM = malloc(sizeof(ofloat) * (M + 1));
    ofloat *VI = malloc(sizeof(ofloat) * (M + 1));
    ofloat *D = malloc(sizeof(ofloat) * (M + 1));
    ofloat TN[5], TJ[5], TC[5];
    for (int i = 0; i < 5 * (M + 1); ++i) TM[i] = TI[i] = NEGINF;
    for (int i = 0; i < 5; ++i) TN[i] = TJ[i] = TC[i] = NEGINF;
    ofloat T = NEGINF;
    const ofloat *tr = p->trans;
    for (int j = 0; j <= L; ++j) {
        int code[6] = {0};
        int lmax = j < 5 ? j : 5;
        for (int l = 1; l <= lmax; ++l) code[l] = code_of(s, j - l, l);
        ofloat VN = NEGINF, VJ = NEGINF, VC = NEGINF;
        for (int l = 1; l <= lmax; ++l) {
            int slot = (j - l) % 5;
            ofloat e = p->emN[code[l]];
            VN = fmx(VN, TN[slot] + e);
            VJ = fmx(VJ, TJ[slot] + e);
            VC = fmx(VC, TC[slot] + e);
        }
        ofloat E = NEGINF;
        VM[0] = VI[0] = D[0] = NEGINF;
        for (int k = 1; k <= M; ++k) {
            ofloat vm = NEGINF, vi = NEGINF;
            const ofloat *em = p->emM + (size_t)(k - 1) * NTAB;
            for (int l = 1; l <= lmax; ++l) {
                int slot = (j - l) % 5;
                vm = fmx(vm, TM[slot * (M + 1) + k] + em[code[l]]);
                vi = fmx(vi, TI[slot * (M + 1) + k] + p->emI[code[l]]);
            }
            VM[k] = vm;
            VI[k] = vi;
            if (k >= 2) {
                const ofloat *t = tr + (size_t)(k - 1) * NTRANS;
                D[k] = fmx(VM[k - 1] + t[2], D[k - 1] + t[6]);
            } else
                D[k] = NEGINF;
            E = fmx(E, vm);
            if (k >= 2) E = fmx(E, D[k]);
        }
        ofloat S = j == 0 ? (ofloat)0 : NEGINF;
        ofloat B = fmx(fmx(S + x->NB, VN + x->NB), fmx(VJ + x->JB, E + x->EB));
        int slot = j % 5;
        TN[slot] = fmx(S + x->NN, VN + x->NN);
        TJ[slot] = fmx(E + x->EJJ, VJ + x->JJ);
        TC[slot] = fmx(E + x->ECC, VC + x->CC);
        T = fmx(E + x->ET, VC + x->CT);
        for (int k = 1; k <= M; ++k) {
            ofloat tm = B + p->entry[k - 1];
            if (k >= 2) {
                const ofloat *t = tr + (size_t)(k - 1) * NTRANS;
                tm = fmx(tm, VM[k - 1] + t[0]);
                tm = fmx(tm, VI[k - 1] + t[3]);
                tm = fmx(tm, D[k - 1] + t[5]);
            }
            TM[slot * (M + 1) + k] = tm;
            if (k <= M - 1) {
                const ofloat *t = tr + (size_t)k * NTRANS;
                TI[slot * (M + 1) + k] = fmx(VM[k] + t[1], VI[k] + t[4]);
            } else
                TI[slot * (M + 1) + k] = NEGINF;
        }
    }
    free(TM); free(TI); free(VM); free(VI); free(D);
    return T;
}

int orc_scores_fast(const struct orc_profile *p, const char *seq, int L, int multi_hits,
                    int hmmer3_compat, ofloat *null_ll, ofloat *alt_ll)
{
    struct xtrans x;
    int rc = specials((unsigned)L, multi_hits, hmmer3_compat, &x);
    if (rc) return rc;
    uint8_t *s = malloc(L);
    if ((rc = encode_seq(seq, L, s))) { free(s); return rc; }
    *null_ll = null_score_fast(p, &x, s, L);
    *alt_ll = alt_score_fast(p, &x, s, L);
    free(s);
    return RC_OK;
}

/* ------------------------------------------------------------------ */
/* decode: protein_profile_decode (src/model/protein_profile.c:306-331) */
/* -> imm_frame_cond_decode; argmax over codons in ACGT-lexicographic   */
/* order, strict '>' (imm's tie rule is unpinned).                      */
/* ------------------------------------------------------------------ */
int orc_decode(const struct orc_profile *p, unsigned state_id, const char *frag, int len,
               char codon_out[3], char *amino_out)
{
    const struct ndist *d;
    unsigned msb = state_id & (3u << 14);
    if (msb == ST_INSERT)
        d = &p->ins_nd;
    else if (msb == ST_MATCH)
        d = &p->match_nd[(state_id & 0x3fff) - 1];
    else
        d = &p->null_nd;
    int z[5];
    if (len < 1 || len > 5) return RC_EINVAL;
    for (int i = 0; i < len; ++i) {
        z[i] = nuc_idx(frag[i]);
        if (z[i] < 0) return RC_EINVAL;
    }
    double best = NEGINF;
    int arg = -1;
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b)
            for (int c = 0; c < 4; ++c) {
                double clp = d->codonm[a * 25 + b * 5 + c];
                double mg[125];
                for (int x = 0; x < 5; ++x)
                    for (int y = 0; y < 5; ++y)
                        for (int w = 0; w < 5; ++w)
                            mg[x * 25 + y * 5 + w] =
                                ((x == 4 || x == a) && (y == 4 || y == b) && (w == 4 || w == c)) ? clp
                                                                                                : NEGINF;
                double v = frame_lprob(z, len, d->nucltp, mg, p->eps);
                if (v > best) {
                    best = v;
                    arg = a * 16 + b * 4 + c;
                }
            }
    if (arg < 0) return RC_EINVAL; /* NaN / no finite codon: protein_profile.c:327-328 */
    int a = arg >> 4, b = (arg >> 2) & 3, c = arg & 3;
    codon_out[0] = NUC[a];
    codon_out[1] = NUC[b];
    codon_out[2] = NUC[c];
    *amino_out = gc_aa(a, b, c); /* imm_gc_decode(1, codon), protein_match.c:44 */
    return RC_OK;
}

/* protein_state_name: src/model/protein_state.c:5-39 */
int orc_state_name(unsigned id, char *name)
{
    unsigned msb = id & (3u << 14);
    if (msb == ST_EXT) {
        static const char x[] = "RSNBEJCT";
        name[0] = x[id & 7];
        name[1] = 0;
        return 1;
    }
    name[0] = msb == ST_MATCH ? 'M' : msb == ST_INSERT ? 'I' : 'D';
    return 1 + sprintf(name + 1, "%u", (unsigned)(id & 0x3fff));
}

static int is_mute(unsigned id)
{
    unsigned msb = id & (3u << 14);
    if (msb == ST_EXT) return id == ST_S || id == ST_B || id == ST_E || id == ST_T;
    return msb == ST_DELETE;
}

/*
 * Product row: src/server/prod.c:13-41 (write_begin), :153-181 (prod_fwrite),
 * src/server/protein_match.c:21-56.  Returns bytes written (excluding NUL) or -1.
 */
long orc_product_row(const struct orc_profile *p, long scan_id, long seq_id, const char *accession,
                     double alt_ll, double null_ll, const char *seq, const uint16_t *step_state,
                     const uint8_t *step_len, int nsteps, char *out, long cap)
{
    long n = 0;
#define EMIT(...)                                                                                  \
    do {                                                                                           \
        int w_ = snprintf(out + n, (size_t)(cap - n), __VA_ARGS__);                                \
        if (w_ < 0 || w_ >= cap - n) return -1;                                                    \
        n += w_;                                                                                   \
    } while (0)
    EMIT("%ld\t%ld\t%s\t%s\t%.17g\t%.17g\t%s\t%s\t", scan_id, seq_id, accession, "dna", alt_ll,
         null_ll, "protein", "0.1.0");
    int start = 0;
    for (int i = 0; i < nsteps; ++i) {
        if (i > 0) EMIT(";");
        char name[8], codon[4] = {0}, amino[2] = {0};
        orc_state_name(step_state[i], name);
        int len = step_len[i];
        if (!is_mute(step_state[i])) {
            if (orc_decode(p, step_state[i], seq + start, len, codon, amino)) return -1;
        }
        EMIT("%.*s,%s,%s,%s", len, seq + start, name, codon, amino);
        start += len;
    }
    EMIT("\n");
#undef EMIT
    return n;
}

#ifdef _OPENMP
#include <omp.h>
#endif
/* number of OpenMP threads orc_scan uses (NUM_THREADS of the reference's .env, cli_server.c:40-53) */
int orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* ------------------------------------------------------------------ */
/* scan: restatement of thread_run (src/server/scan_thread.c:86-135)    */
/* over a batch, OpenMP over profiles like scan.c:239-250.               */
/* flavour 0 = generic interpreter for null+alt (imm's algorithmic      */
/* shape), 1 = specialised recurrence for scores, generic only for hits */
/* ------------------------------------------------------------------ */
int orc_scan(int nprof, struct orc_profile *const *profs, int nseq, const char *const *seqs,
             const int *lens, int multi_hits, int hmmer3_compat, double lrt_threshold, int flavour,
             int want_paths, ofloat *null_ll /* [nseq*nprof] */, ofloat *alt_ll, uint8_t *hit,
             int *path_off /* [nseq*nprof+1] or NULL */, uint16_t *step_state, uint8_t *step_len,
             long step_cap)
{
    int rc_all = RC_OK;
    long npairs = (long)nseq * nprof;
    int **pstate = NULL;
    /* paths are collected per pair then concatenated in (seq, profile) order */
    uint16_t **pst = calloc(npairs, sizeof *pst);
    uint8_t **pln = calloc(npairs, sizeof *pln);
    int *pn = calloc(npairs, sizeof *pn);
    (void)pstate;
#pragma omp parallel for schedule(static, 1) collapse(1)
    for (int pi = 0; pi < nprof; ++pi) {
        const struct orc_profile *p = profs[pi];
        /* flavour 0: the profile's two DPs are compiled once (the reference unpacks them once per pair from
         * disk, scan_thread.c:99; holding them in RAM is a deviation in the CPU's favour) and only the
         * length-dependent transitions change per sequence, as protein_profile_setup does */
        struct gmodel *g_alt = NULL, *g_null = NULL;
        uint8_t *enc = NULL;
        int enc_cap = 0;
        for (int si = 0; si < nseq; ++si) {
            long idx = (long)si * nprof + pi;
            int L = lens[si];
            ofloat nl = NEGINF, al = NEGINF;
            int rc;
            int maxst = L + 8;
            uint16_t *ss = NULL;
            uint8_t *sl = NULL;
            int ns = 0;
            if (flavour == 0) {
                ss = malloc(sizeof(uint16_t) * maxst);
                sl = malloc(maxst);
                struct xtrans x;
                rc = specials((unsigned)L, multi_hits, hmmer3_compat, &x);
                if (!rc && L > enc_cap) {
                    free(enc);
                    enc = malloc(L), enc_cap = L;
                }
                if (!rc) rc = encode_seq(seqs[si], L, enc);
                if (!rc) {
                    if (!g_alt) g_alt = gmodel_alt(p, &x), g_null = gmodel_null(p, &x);
                    gmodel_alt_specials(g_alt, &x);
                    g_null->st[0].inc[0].score = x.RR;
                    rc = gviterbi(g_null, enc, L, &nl, NULL, NULL, NULL, 0);
                }
                if (!rc) rc = gviterbi(g_alt, enc, L, &al, ss, sl, &ns, maxst);
            } else {
                rc = orc_scores_fast(p, seqs[si], L, multi_hits, hmmer3_compat, &nl, &al);
            }
            if (rc) {
#pragma omp critical
                rc_all = rc;
                free(ss); free(sl);
                continue;
            }
            null_ll[idx] = nl;
            alt_ll[idx] = al;
            ofloat lrt = (ofloat)-2 * (nl - al); /* xmath.h:32-43 */
            int h = isfinite((double)lrt) && !((double)lrt < lrt_threshold); /* scan_thread.c:123 */
            hit[idx] = (uint8_t)h;
            if (h && want_paths) {
                if (flavour != 0) {
                    ss = malloc(sizeof(uint16_t) * maxst);
                    sl = malloc(maxst);
                    ofloat al2;
                    rc = orc_viterbi_alt(p, seqs[si], L, multi_hits, hmmer3_compat, &al2, ss, sl, &ns, maxst);
                    if (rc || al2 != al) {
#pragma omp critical
                        rc_all = rc ? rc : RC_EFAIL;
                    }
                }
                pst[idx] = ss;
                pln[idx] = sl;
                pn[idx] = ns;
            } else {
                free(ss); free(sl);
            }
        }
        if (g_alt) gmodel_del(g_alt);
        if (g_null) gmodel_del(g_null);
        free(enc);
    }
    if (path_off) {
        long off = 0;
        for (long i = 0; i < npairs; ++i) {
            path_off[i] = (int)off;
            if (pn[i]) {
                if (off + pn[i] > step_cap) { rc_all = RC_ENOMEM; break; }
                memcpy(step_state + off, pst[i], sizeof(uint16_t) * pn[i]);
                memcpy(step_len + off, pln[i], pn[i]);
                off += pn[i];
            }
        }
        path_off[npairs] = (int)off;
    }
    for (long i = 0; i < npairs; ++i) { free(pst[i]); free(pln[i]); }
    free(pst); free(pln); free(pn);
    return rc_all;
}
