/*
 * dcp_model.c -- host side of the model layer (plain C).
 *
 * Mirrors src/model/protein_model.c, protein_profile.c, protein_state.c of the
 * reference: amino-acid profile -> codon tables -> frame-state emission tables and
 * the Plan7-shaped transition set that the sm_100a kernels consume.  The reference
 * delegates the table maths to imm (imm_codon_marg, imm_frame_state, imm_hmm_reset_dp);
 * here it is done directly: model parameters are evaluated in double precision,
 * in the probability domain where that is cheaper, and rounded once to fp32
 * (imm_float of the reference's default build).
 */
#include "dcp_internal.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static const char amino_syms[] = "ACDEFGHIKLMNPQRSTVWY"; /* protein_h3reader.c:83-102 */
static const char nuclt_syms[] = "ACGT";
/* NCBI translation table 1 with bases ordered TCAG at each codon position */
static const char gc1_tcag[] = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";

int dcp_nuclt_index(char c)
{
    const char *p = c ? strchr(nuclt_syms, c) : NULL;
    return p ? (int)(p - nuclt_syms) : -1;
}

char dcp_gc_decode(int a, int b, int c)
{
    static const int rank_tcag[4] = {2, 1, 3, 0};
    return gc1_tcag[16 * rank_tcag[a] + 4 * rank_tcag[b] + rank_tcag[c]];
}

/* ---- log-probability helpers (imm_lprob_add / imm_lprob_normalize) ---- */
static double logaddexp(double x, double y)
{
    if (x == -INFINITY) return y;
    if (y == -INFINITY) return x;
    double hi = x > y ? x : y, lo = x > y ? y : x;
    return hi + log1p(exp(lo - hi));
}

static double logsumexp(unsigned n, double const *v)
{
    double top = -INFINITY;
    for (unsigned i = 0; i < n; ++i)
        if (v[i] > top) top = v[i];
    if (top == -INFINITY) return top;
    double acc = 0.0;
    for (unsigned i = 0; i < n; ++i) acc += exp(v[i] - top);
    return top + log(acc);
}

/* ---- setup_nuclt_dist (protein_model.c:342-408) ---- */
void dcp_nuclt_dist_setup(struct dcp_nuclt_dist *nd, double const amino_lprobs[DCP_AMINO_SIZE])
{
    /* codon_lprob (:361-394): split each amino acid evenly over its codons, stops get zero */
    unsigned ncodons[128] = {0};
    for (int i = 0; i < 64; ++i) ncodons[(unsigned char)gc1_tcag[i]]++;
    double by_amino[128];
    for (int i = 0; i < 128; ++i) by_amino[i] = -INFINITY;
    for (int i = 0; i < DCP_AMINO_SIZE; ++i)
    {
        unsigned char aa = (unsigned char)amino_syms[i];
        by_amino[aa] = amino_lprobs[i] - log((double)ncodons[aa]);
    }
    double codon[64];
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b)
            for (int c = 0; c < 4; ++c)
                codon[16 * a + 4 * b + c] = by_amino[(unsigned char)dcp_gc_decode(a, b, c)];
    double z = logsumexp(64, codon); /* imm_codon_lprob_normalize (:404) */
    for (int i = 0; i < 64; ++i) codon[i] -= z;

    /* nuclt_lprob (:342-359): base frequency over the three codon positions */
    double pile[4][192];
    unsigned fill[4] = {0};
    double const third = log(3.0);
    for (int i = 0; i < 64; ++i)
    {
        int pos[3] = {i >> 4, (i >> 2) & 3, i & 3};
        for (int k = 0; k < 3; ++k) pile[pos[k]][fill[pos[k]]++] = codon[i] - third;
    }
    for (int x = 0; x < 4; ++x) nd->nucltp[x] = logsumexp(fill[x], pile[x]);

    /* imm_codon_marg (:407): 5x5x5 with index 4 = "any" */
    for (int idx = 0; idx < 125; ++idx)
    {
        int pa = idx / 25, pb = (idx / 5) % 5, pc = idx % 5;
        double terms[64];
        unsigned n = 0;
        for (int i = 0; i < 64; ++i)
        {
            int a = i >> 4, b = (i >> 2) & 3, c = i & 3;
            if ((pa == 4 || pa == a) && (pb == 4 || pb == b) && (pc == 4 || pc == c))
                terms[n++] = codon[i];
        }
        nd->codonm[idx] = logsumexp(n, terms);
    }
}

/* ---- imm_frame_state emission tables, probability domain ---- */
struct frame_ctx
{
    double B[4];     /* base probabilities */
    double m[125];   /* codon marginals, probabilities */
    double one[4];   /* m(x,_,_) + m(_,x,_) + m(_,_,x) */
    double two[16];  /* m(_,x,y) + m(x,_,y) + m(x,y,_) */
    double c1, c2a, c2b, c3a, c3b, c3c, c4a, c4b, c5; /* epsilon polynomials */
};

#define MM_(a, b, c) (fc->m[25 * (a) + 5 * (b) + (c)])

static void frame_ctx_init(struct frame_ctx *fc, struct dcp_nuclt_dist const *nd, double eps)
{
    for (int x = 0; x < 4; ++x) fc->B[x] = exp(nd->nucltp[x]);
    for (int i = 0; i < 125; ++i) fc->m[i] = exp(nd->codonm[i]);
    for (int x = 0; x < 4; ++x) fc->one[x] = MM_(x, 4, 4) + MM_(4, x, 4) + MM_(4, 4, x);
    for (int x = 0; x < 4; ++x)
        for (int y = 0; y < 4; ++y) fc->two[4 * x + y] = MM_(4, x, y) + MM_(x, 4, y) + MM_(x, y, 4);
    double e = eps, f = 1.0 - eps;
    fc->c1 = e * e * f * f / 3.0;
    fc->c2a = 2.0 * e * f * f * f / 3.0;
    fc->c2b = e * e * e * f / 3.0;
    fc->c3a = f * f * f * f;
    fc->c3b = 4.0 * e * e * f * f / 9.0;
    fc->c3c = e * e * e * e / 9.0;
    fc->c4a = e * f * f * f / 2.0;
    fc->c4b = e * e * e * f / 9.0;
    fc->c5 = e * e * f * f / 10.0;
}

/* probability of emitting the n-nt string z */
static double frame_prob(struct frame_ctx const *fc, int const *z, int n)
{
    double const *B = fc->B;
    switch (n)
    {
    case 1:
        return fc->c1 * fc->one[z[0]];
    case 2:
        return fc->c2a * fc->two[4 * z[0] + z[1]] +
               fc->c2b * (B[z[1]] * fc->one[z[0]] + B[z[0]] * fc->one[z[1]]);
    case 3: {
        double keep = MM_(z[0], z[1], z[2]);
        double sub = B[z[0]] * fc->two[4 * z[1] + z[2]] + B[z[1]] * fc->two[4 * z[0] + z[2]] +
                     B[z[2]] * fc->two[4 * z[0] + z[1]];
        double rare = B[z[1]] * B[z[2]] * fc->one[z[0]] + B[z[0]] * B[z[2]] * fc->one[z[1]] +
                      B[z[0]] * B[z[1]] * fc->one[z[2]];
        return fc->c3a * keep + fc->c3b * sub + fc->c3c * rare;
    }
    case 4: {
        /* one inserted base i: codon = the other three; two inserted bases i<j plus one deletion */
        static const unsigned char pair4[6][4] = {{0, 1, 2, 3}, {0, 2, 1, 3}, {0, 3, 1, 2},
                                                  {1, 2, 0, 3}, {1, 3, 0, 2}, {2, 3, 0, 1}};
        double ins1 = B[z[0]] * MM_(z[1], z[2], z[3]) + B[z[1]] * MM_(z[0], z[2], z[3]) +
                      B[z[2]] * MM_(z[0], z[1], z[3]) + B[z[3]] * MM_(z[0], z[1], z[2]);
        double mix = 0.0;
        for (int q = 0; q < 6; ++q)
            mix += B[z[pair4[q][0]]] * B[z[pair4[q][1]]] * fc->two[4 * z[pair4[q][2]] + z[pair4[q][3]]];
        return fc->c4a * ins1 + fc->c4b * mix;
    }
    default: {
        static const unsigned char pair5[10][5] = {{0, 1, 2, 3, 4}, {0, 2, 1, 3, 4}, {0, 3, 1, 2, 4}, {0, 4, 1, 2, 3},
                                                   {1, 2, 0, 3, 4}, {1, 3, 0, 2, 4}, {1, 4, 0, 2, 3}, {2, 3, 0, 1, 4},
                                                   {2, 4, 0, 1, 3}, {3, 4, 0, 1, 2}};
        double ins2 = 0.0;
        for (int q = 0; q < 10; ++q)
            ins2 += B[z[pair5[q][0]]] * B[z[pair5[q][1]]] * MM_(z[pair5[q][2]], z[pair5[q][3]], z[pair5[q][4]]);
        return fc->c5 * ins2;
    }
    }
}

static const unsigned frame_offset[6] = {0, 0, 4, 20, 84, 340};

unsigned dcp_frame_code(unsigned len, unsigned packed) { return frame_offset[len] + packed; }

void dcp_frame_table(struct dcp_nuclt_dist const *nd, double eps, float out[DCP_FRAME_TABLE_SIZE])
{
    struct frame_ctx fc;
    frame_ctx_init(&fc, nd, eps);
    for (int n = 1; n <= 5; ++n)
    {
        unsigned count = 1u << (2 * n);
        for (unsigned v = 0; v < count; ++v)
        {
            int z[5];
            for (int i = 0; i < n; ++i) z[i] = (int)(v >> (2 * (n - 1 - i))) & 3;
            out[frame_offset[n] + v] = (float)log(frame_prob(&fc, z, n));
        }
    }
}

/* ---- the same emission in the log domain; used by decode where exact ties matter ---- */
static double frame_lprob_log(double const nucltp[4], double const mg[125], double eps, int const *z,
                              int n)
{
#define LM(a, b, c) (mg[25 * (a) + 5 * (b) + (c)])
    double const le = log(eps), lf = log(1.0 - eps);
    double t[18];
    unsigned k = 0;
    if (n == 1)
    {
        t[0] = LM(z[0], 4, 4), t[1] = LM(4, z[0], 4), t[2] = LM(4, 4, z[0]);
        return 2 * le + 2 * lf - log(3.0) + logsumexp(3, t);
    }
    if (n == 2)
    {
        t[0] = LM(4, z[0], z[1]), t[1] = LM(z[0], 4, z[1]), t[2] = LM(z[0], z[1], 4);
        double p = log(2.0) + le + 3 * lf - log(3.0) + logsumexp(3, t);
        t[0] = nucltp[z[1]] + LM(z[0], 4, 4);
        t[1] = nucltp[z[1]] + LM(4, z[0], 4);
        t[2] = nucltp[z[1]] + LM(4, 4, z[0]);
        t[3] = nucltp[z[0]] + LM(z[1], 4, 4);
        t[4] = nucltp[z[0]] + LM(4, z[1], 4);
        t[5] = nucltp[z[0]] + LM(4, 4, z[1]);
        double q = 3 * le + lf - log(3.0) + logsumexp(6, t);
        return logaddexp(p, q);
    }
    if (n == 3)
    {
        double v[3];
        v[0] = 4 * lf + LM(z[0], z[1], z[2]);
        for (int i = 0; i < 3; ++i)
        {
            int r0 = z[i == 0 ? 1 : 0], r1 = z[i == 2 ? 1 : 2];
            t[k++] = nucltp[z[i]] + LM(4, r0, r1);
            t[k++] = nucltp[z[i]] + LM(r0, 4, r1);
            t[k++] = nucltp[z[i]] + LM(r0, r1, 4);
        }
        v[1] = log(4.0) + 2 * le + 2 * lf - log(9.0) + logsumexp(9, t);
        k = 0;
        for (int i = 0; i < 3; ++i)
        {
            double others = 0.0;
            for (int j = 0; j < 3; ++j)
                if (j != i) others += nucltp[z[j]];
            t[k++] = others + LM(z[i], 4, 4);
            t[k++] = others + LM(4, z[i], 4);
            t[k++] = others + LM(4, 4, z[i]);
        }
        v[2] = 4 * le - log(9.0) + logsumexp(9, t);
        return logsumexp(3, v);
    }
    if (n == 4)
    {
        for (int i = 0; i < 4; ++i)
        {
            int r[3], c = 0;
            for (int j = 0; j < 4; ++j)
                if (j != i) r[c++] = z[j];
            t[k++] = nucltp[z[i]] + LM(r[0], r[1], r[2]);
        }
        double p = le + 3 * lf - log(2.0) + logsumexp(4, t);
        k = 0;
        for (int i = 0; i < 4; ++i)
            for (int j = i + 1; j < 4; ++j)
            {
                int r[2], c = 0;
                for (int q = 0; q < 4; ++q)
                    if (q != i && q != j) r[c++] = z[q];
                double w = nucltp[z[i]] + nucltp[z[j]];
                t[k++] = w + LM(4, r[0], r[1]);
                t[k++] = w + LM(r[0], 4, r[1]);
                t[k++] = w + LM(r[0], r[1], 4);
            }
        double q = 3 * le + lf - log(9.0) + logsumexp(18, t);
        return logaddexp(p, q);
    }
    for (int i = 0; i < 5; ++i)
        for (int j = i + 1; j < 5; ++j)
        {
            int r[3], c = 0;
            for (int q = 0; q < 5; ++q)
                if (q != i && q != j) r[c++] = z[q];
            t[k++] = nucltp[z[i]] + nucltp[z[j]] + LM(r[0], r[1], r[2]);
        }
    return 2 * le + 2 * lf - log(10.0) + logsumexp(10, t);
#undef LM
}

/* ---- imm_rnd: splitmix64-seeded xoshiro256+ ---- */
struct xoshiro
{
    uint64_t w[4];
};

static struct xoshiro xoshiro_seed(uint64_t seed)
{
    struct xoshiro g;
    for (int i = 0; i < 4; ++i)
    {
        seed += 0x9e3779b97f4a7c15ULL;
        uint64_t v = seed;
        v = (v ^ (v >> 30)) * 0xbf58476d1ce4e5b9ULL;
        v = (v ^ (v >> 27)) * 0x94d049bb133111ebULL;
        g.w[i] = v ^ (v >> 31);
    }
    return g;
}

static double xoshiro_unit(struct xoshiro *g)
{
    uint64_t out = g->w[0] + g->w[3];
    uint64_t sh = g->w[1] << 17;
    g->w[2] ^= g->w[0];
    g->w[3] ^= g->w[1];
    g->w[1] ^= g->w[2];
    g->w[0] ^= g->w[3];
    g->w[2] ^= sh;
    g->w[3] = (g->w[3] << 45) | (g->w[3] >> 19);
    return (double)(out >> 11) * (1.0 / 9007199254740992.0);
}

/* ---- protein_model ---- */
struct protein_model *protein_model_new(struct protein_cfg cfg, float const null_lprobs[DCP_AMINO_SIZE])
{
    if (!(cfg.epsilon >= 0.0f && cfg.epsilon <= 1.0f)) /* protein_cfg.h:18 assert */
    {
        dcp_set_error("epsilon outside [0, 1]");
        return NULL;
    }
    struct protein_model *m = calloc(1, sizeof *m);
    if (!m) return NULL;
    m->cfg = cfg;
    memcpy(m->null_lprobs, null_lprobs, sizeof m->null_lprobs);
    double nl[DCP_AMINO_SIZE], zeros[DCP_AMINO_SIZE] = {0};
    for (int i = 0; i < DCP_AMINO_SIZE; ++i) nl[i] = (double)null_lprobs[i];
    dcp_nuclt_dist_setup(&m->null_ndist, nl);      /* protein_model.c:122 */
    dcp_nuclt_dist_setup(&m->insert_ndist, zeros); /* protein_model.c:126-127 */
    m->node_idx = m->trans_idx = UINT32_MAX;
    return m;
}

enum rc protein_model_setup(struct protein_model *m, unsigned core_size)
{
    if (core_size == 0) return dcp_error(RC_EINVAL, "`core_size` cannot be zero");
    if (core_size > DCP_PROTEIN_MODEL_CORE_SIZE_MAX) return dcp_error(RC_EINVAL, "`core_size` is too big");
    void *a = realloc(m->match_ndists, core_size * sizeof *m->match_ndists);
    if (!a) return dcp_error(RC_ENOMEM, "failed to alloc nodes");
    m->match_ndists = a;
    void *b = realloc(m->trans, (core_size + 1) * sizeof *m->trans);
    if (!b) return dcp_error(RC_ENOMEM, "failed to alloc trans");
    m->trans = b;
    m->core_size = core_size;
    m->consensus[core_size] = '\0';
    m->node_idx = 0;
    m->trans_idx = 0;
    return RC_OK;
}

enum rc protein_model_add_node(struct protein_model *m, float const lprobs[DCP_AMINO_SIZE], char consensus)
{
    if (m->core_size == 0) return dcp_error(RC_EFAIL, "must call protein_model_setup first");
    if (m->node_idx == m->core_size) return dcp_error(RC_EFAIL, "reached limit of nodes");
    m->consensus[m->node_idx] = consensus;
    double lodds[DCP_AMINO_SIZE];
    for (int i = 0; i < DCP_AMINO_SIZE; ++i) /* imm_float subtraction, protein_model.c:60-62 */
        lodds[i] = (double)(float)(lprobs[i] - m->null_lprobs[i]);
    dcp_nuclt_dist_setup(&m->match_ndists[m->node_idx], lodds);
    m->node_idx++;
    return RC_OK;
}

enum rc protein_model_add_trans(struct protein_model *m, struct protein_trans trans)
{
    if (m->core_size == 0) return dcp_error(RC_EFAIL, "must call protein_model_setup first");
    if (m->trans_idx == m->core_size + 1) return dcp_error(RC_EFAIL, "reached limit of transitions");
    m->trans[m->trans_idx++] = trans;
    return RC_OK;
}

void protein_model_del(struct protein_model *m)
{
    if (!m) return;
    free(m->match_ndists);
    free(m->trans);
    free(m);
}

static bool model_complete(struct protein_model const *m)
{
    return m->core_size > 0 && m->node_idx == m->core_size && m->trans_idx == m->core_size + 1;
}

/* ---- protein_profile ---- */
struct protein_profile *protein_profile_new(char const *accession, struct protein_cfg cfg)
{
    if (!(cfg.epsilon >= 0.0f && cfg.epsilon <= 1.0f))
    {
        dcp_set_error("epsilon outside [0, 1]");
        return NULL;
    }
    struct protein_profile *p = calloc(1, sizeof *p);
    if (!p) return NULL;
    p->cfg = cfg;
    strncpy(p->accession, accession ? accession : "", DCP_PROFILE_ACC_SIZE - 1);
    return p;
}

void protein_profile_del(struct protein_profile *p)
{
    if (!p) return;
    free(p->match_ndists);
    free(p->match_emission);
    free(p->trans);
    free(p->entry);
    free(p->consensus);
    free(p);
}

/* calculate_occupancy + setup_entry_trans (protein_model.c:258-283,410-439) */
static void entry_distribution(struct protein_model const *m, float *entry)
{
    unsigned n = m->core_size;
    if (m->cfg.entry_dist == ENTRY_DIST_UNIFORM)
    {
        double M = (double)n;
        float cost = (float)(log(2.0 / (M * (M + 1.0))) * M); /* sic: multiplied by M, :414-415 */
        for (unsigned i = 0; i < n; ++i) entry[i] = cost;
        return;
    }
    double *occ = malloc(n * sizeof *occ);
    struct protein_trans const *t = m->trans;
    occ[0] = logaddexp((double)t->MI, (double)t->MM);
    for (unsigned i = 1; i < n; ++i)
    {
        ++t;
        double stay = occ[i - 1] + logaddexp((double)t->MM, (double)t->MI);
        double back = log1p(-exp(occ[i - 1])) + (double)t->DM;
        occ[i] = logaddexp(stay, back);
    }
    double logZ = -INFINITY;
    for (unsigned i = 0; i < n; ++i) logZ = logaddexp(logZ, occ[i] + log((double)(n - i)));
    for (unsigned i = 0; i < n; ++i) entry[i] = (float)(occ[i] - logZ);
    free(occ);
}

enum rc protein_profile_absorb(struct protein_profile *p, struct protein_model const *m)
{
    if (!model_complete(m)) return dcp_error(RC_EFAIL, "model is not complete");
    if (m->cfg.entry_dist != ENTRY_DIST_UNIFORM && m->cfg.entry_dist != ENTRY_DIST_OCCUPANCY)
        return dcp_error(RC_EINVAL, "unknown entry distribution");
    unsigned n = m->core_size;
    p->cfg = m->cfg;
    p->core_size = n;
    free(p->match_ndists);
    free(p->match_emission);
    free(p->trans);
    free(p->entry);
    free(p->consensus);
    p->match_ndists = malloc(n * sizeof *p->match_ndists);
    p->match_emission = malloc((size_t)n * DCP_FRAME_TABLE_SIZE * sizeof(float));
    p->trans = malloc((n + 1) * sizeof *p->trans);
    p->entry = malloc(n * sizeof(float));
    p->consensus = malloc(n + 1);
    if (!p->match_ndists || !p->match_emission || !p->trans || !p->entry || !p->consensus)
        return dcp_error(RC_ENOMEM, "alloc profile tables");
    memcpy(p->consensus, m->consensus, n + 1);
    memcpy(p->match_ndists, m->match_ndists, n * sizeof *p->match_ndists);
    memcpy(p->trans, m->trans, (n + 1) * sizeof *p->trans);
    p->null_ndist = m->null_ndist;
    p->insert_ndist = m->insert_ndist;
    double eps = (double)m->cfg.epsilon;
    dcp_frame_table(&p->null_ndist, eps, p->null_emission);
    dcp_frame_table(&p->insert_ndist, eps, p->insert_emission);
    for (unsigned k = 0; k < n; ++k)
        dcp_frame_table(&p->match_ndists[k], eps, p->match_emission + (size_t)k * DCP_FRAME_TABLE_SIZE);
    entry_distribution(m, p->entry);
    return RC_OK;
}

enum rc protein_profile_sample(struct protein_profile *p, unsigned seed, unsigned core_size)
{
    if (core_size < 2) return dcp_error(RC_EINVAL, "core_size must be >= 2"); /* assert at :262 */
    struct xoshiro g = xoshiro_seed(seed);
    double tmp[DCP_AMINO_SIZE];
    float lp[DCP_AMINO_SIZE];
    for (int i = 0; i < DCP_AMINO_SIZE; ++i) tmp[i] = log(xoshiro_unit(&g));
    double z = logsumexp(DCP_AMINO_SIZE, tmp);
    for (int i = 0; i < DCP_AMINO_SIZE; ++i) lp[i] = (float)(tmp[i] - z);

    struct protein_model *m = protein_model_new(p->cfg, lp);
    if (!m) return dcp_error(RC_ENOMEM, "alloc model");
    enum rc rc = protein_model_setup(m, core_size);
    for (unsigned k = 0; !rc && k < core_size; ++k)
    {
        for (int i = 0; i < DCP_AMINO_SIZE; ++i) tmp[i] = log(xoshiro_unit(&g));
        z = logsumexp(DCP_AMINO_SIZE, tmp);
        for (int i = 0; i < DCP_AMINO_SIZE; ++i) lp[i] = (float)(tmp[i] - z);
        rc = protein_model_add_node(m, lp, '-');
    }
    for (unsigned k = 0; !rc && k <= core_size; ++k)
    {
        double t[PROTEIN_TRANS_SIZE];
        for (int i = 0; i < PROTEIN_TRANS_SIZE; ++i) t[i] = log(xoshiro_unit(&g));
        if (k == 0) t[6] = -INFINITY;
        if (k == core_size) t[2] = t[6] = -INFINITY;
        z = logsumexp(PROTEIN_TRANS_SIZE, t);
        struct protein_trans tr;
        for (int i = 0; i < PROTEIN_TRANS_SIZE; ++i) tr.data[i] = (float)(t[i] - z);
        rc = protein_model_add_trans(m, tr);
    }
    if (!rc) rc = protein_profile_absorb(p, m);
    protein_model_del(m);
    return rc;
}

/* The loop body of protein_h3reader_next (src/model/protein_h3reader.c:18-72) for arrays that are
 * already in memory: trans[0], then per node (match lprobs, trans[k]); then absorb. */
enum rc protein_profile_build(struct protein_profile *p, unsigned core_size,
                              float const null_lprobs[DCP_AMINO_SIZE], float const *match_lprobs,
                              float const *trans, char const *consensus)
{
    struct protein_model *m = protein_model_new(p->cfg, null_lprobs);
    if (!m) return dcp_error(RC_ENOMEM, "alloc model");
    enum rc rc = protein_model_setup(m, core_size);
    struct protein_trans t;
    if (!rc)
    {
        memcpy(t.data, trans, sizeof t.data);
        rc = protein_model_add_trans(m, t);
    }
    for (unsigned k = 0; !rc && k < core_size; ++k)
    {
        rc = protein_model_add_node(m, match_lprobs + (size_t)k * DCP_AMINO_SIZE, consensus ? consensus[k] : '-');
        if (rc) break;
        memcpy(t.data, trans + (size_t)(k + 1) * PROTEIN_TRANS_SIZE, sizeof t.data);
        rc = protein_model_add_trans(m, t);
    }
    if (!rc) rc = protein_profile_absorb(p, m);
    protein_model_del(m);
    return rc;
}

void dcp_specials(unsigned seq_size, bool multi_hits, bool hmmer3_compat, float x[13])
{
    double L = (double)(float)seq_size;
    double q = multi_hits ? 0.5 : 0.0;
    double log_q = multi_hits ? log(0.5) : -INFINITY;
    double denom = log(L + 2.0 + q / (1.0 - q));
    float loop = (float)(log(L) - denom);
    float move = (float)(log(2.0 + q / (1.0 - q)) - denom);
    float NN = hmmer3_compat ? 0.0f : loop;
    x[0] = NN, x[1] = NN, x[2] = NN;        /* NN CC JJ */
    x[3] = move, x[4] = move, x[5] = move;  /* NB CT JB */
    x[6] = (float)(log(L) - log(L + 1.0));  /* RR */
    x[7] = (float)log_q;                    /* EJ */
    x[8] = (float)log(1.0 - q);             /* EC */
    x[9] = x[8] + x[4];                     /* E->T = EC + CT, protein_profile.c:206 */
    x[10] = x[8] + x[1];                    /* E->C = EC + CC */
    x[11] = x[7] + x[5];                    /* E->B = EJ + JB */
    x[12] = x[7] + x[2];                    /* E->J = EJ + JJ */
}

enum rc protein_profile_setup(struct protein_profile *p, unsigned seq_size, bool multi_hits,
                              bool hmmer3_compat, float out13[13])
{
    (void)p;
    if (seq_size == 0) return dcp_error(RC_EINVAL, "sequence cannot be empty"); /* :158 */
    float x[13];
    dcp_specials(seq_size, multi_hits, hmmer3_compat, x);
    if (out13) memcpy(out13, x, sizeof x);
    return RC_OK;
}

bool protein_state_is_mute(unsigned id)
{
    unsigned msb = id & (3u << 14);
    if (msb == PROTEIN_EXT_STATE)
        return id == PROTEIN_S_STATE || id == PROTEIN_B_STATE || id == PROTEIN_E_STATE || id == PROTEIN_T_STATE;
    return msb == PROTEIN_DELETE_STATE;
}

unsigned protein_state_name(unsigned id, char name[DCP_STATE_NAME_SIZE])
{
    unsigned msb = id & (3u << 14);
    if (msb == PROTEIN_EXT_STATE)
    {
        name[0] = "RSNBEJCT"[id & 7];
        name[1] = '\0';
        return 1;
    }
    name[0] = msb == PROTEIN_MATCH_STATE ? 'M' : msb == PROTEIN_INSERT_STATE ? 'I' : 'D';
    unsigned v = id & 0x3fff, len = 1;
    char digits[6];
    unsigned nd = 0;
    do
    {
        digits[nd++] = (char)('0' + v % 10);
        v /= 10;
    } while (v);
    while (nd) name[len++] = digits[--nd];
    name[len] = '\0';
    return len;
}

float xmath_lrt_f32(float null_loglik, float alt_loglik) { return -2 * (null_loglik - alt_loglik); }

enum rc protein_profile_decode(struct protein_profile const *p, char const *frag, unsigned frag_size,
                               unsigned state_id, char codon[3], char *amino)
{
    if (protein_state_is_mute(state_id)) return dcp_error(RC_EINVAL, "mute states emit nothing");
    if (frag_size < 1 || frag_size > 5) return dcp_error(RC_EINVAL, "fragment must have 1..5 nucleotides");
    unsigned msb = state_id & (3u << 14);
    struct dcp_nuclt_dist const *nd = &p->null_ndist;
    if (msb == PROTEIN_INSERT_STATE) nd = &p->insert_ndist;
    else if (msb == PROTEIN_MATCH_STATE)
    {
        unsigned k = (state_id & 0x3fff) - 1;
        if (k >= p->core_size) return dcp_error(RC_EINVAL, "state outside the profile");
        nd = &p->match_ndists[k];
    }
    int z[5];
    for (unsigned i = 0; i < frag_size; ++i)
        if ((z[i] = dcp_nuclt_index(frag[i])) < 0) return dcp_error(RC_EINVAL, "failed to decode sequence");
    double eps = (double)p->cfg.epsilon;
    double top = -INFINITY;
    int arg = -1;
    for (int cdn = 0; cdn < 64; ++cdn)
    {
        int a = cdn >> 4, b = (cdn >> 2) & 3, c = cdn & 3;
        double lp = nd->codonm[25 * a + 5 * b + c];
        double onehot[125];
        for (int i = 0; i < 125; ++i)
        {
            int pa = i / 25, pb = (i / 5) % 5, pc = i % 5;
            onehot[i] = ((pa == 4 || pa == a) && (pb == 4 || pb == b) && (pc == 4 || pc == c)) ? lp : -INFINITY;
        }
        double joint = frame_lprob_log(nd->nucltp, onehot, eps, z, (int)frag_size);
        if (joint > top) top = joint, arg = cdn;
    }
    if (arg < 0) return dcp_error(RC_EINVAL, "failed to decode sequence"); /* :327-328 */
    int a = arg >> 4, b = (arg >> 2) & 3, c = arg & 3;
    codon[0] = nuclt_syms[a], codon[1] = nuclt_syms[b], codon[2] = nuclt_syms[c];
    if (amino) *amino = dcp_gc_decode(a, b, c);
    return RC_OK;
}

unsigned protein_profile_core_size(struct protein_profile const *p) { return p->core_size; }
char const *protein_profile_accession(struct protein_profile const *p) { return p->accession; }
float const *protein_profile_match_emission(struct protein_profile const *p) { return p->match_emission; }
float const *protein_profile_insert_emission(struct protein_profile const *p) { return p->insert_emission; }
float const *protein_profile_null_emission(struct protein_profile const *p) { return p->null_emission; }
float const *protein_profile_trans(struct protein_profile const *p) { return (float const *)p->trans; }
float const *protein_profile_entry(struct protein_profile const *p) { return p->entry; }

enum rc protein_profile_nuclt_dist(struct protein_profile const *p, int which, double out[129])
{
    struct dcp_nuclt_dist const *nd;
    if (which == -2) nd = &p->null_ndist;
    else if (which == -1) nd = &p->insert_ndist;
    else if (which >= 0 && (unsigned)which < p->core_size) nd = &p->match_ndists[which];
    else return dcp_error(RC_EINVAL, "no such nuclt_dist");
    memcpy(out, nd->nucltp, sizeof nd->nucltp);
    memcpy(out + 4, nd->codonm, sizeof nd->codonm);
    return RC_OK;
}
