/*
 * dcp_h3reader.c -- HMMER3 ASCII profile reader and in-memory press.
 *
 * Mirrors src/model/protein_h3reader.c:18-103 (protein_h3reader_init/next/del) and the loop of
 * hmm_press (src/server/hmm.c:120-178).  The reference parses the file with hmmer-reader 0.1.3
 * (external, not in its tree); this is a direct reader of the published HMMER3/f text format:
 *   header lines (NAME, ACC, LENG, ...), "HMM" line + transition header line, optional COMPO line,
 *   node 0: insert emissions + transitions, then per node: match line ("k  20 scores  MAP CONS RF MM CS"),
 *   insert emissions, 7 transitions (m->m m->i m->d i->m i->i d->m d->d), terminated by "//".
 * Scores are -ln(p); '*' means p = 0.  As in the reference, insert emissions are ignored and the
 * background is HMMER3's Swiss-Prot 50.8 amino frequencies (protein_h3reader.c:79-103).
 */
#include "dcp_internal.h"

#include <ctype.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { LINE_MAX_ = 8192 };

struct protein_h3reader
{
    FILE *fp;
    struct protein_cfg cfg;
    float null_lprobs[DCP_AMINO_SIZE];
    struct protein_model *model;
    char name[64], acc[DCP_PROFILE_ACC_SIZE];
    unsigned leng;
    char line[LINE_MAX_];
    unsigned lineno;
};

/* HMMER3 background frequencies, Swiss-Prot 50.8, order ACDEFGHIKLMNPQRSTVWY (protein_h3reader.c:79-103) */
static const double swissprot_bg[DCP_AMINO_SIZE] = {
    0.0787945, 0.0151600, 0.0535222, 0.0668298, 0.0397062, 0.0695071, 0.0229198, 0.0590092, 0.0594422, 0.0963728,
    0.0237718, 0.0414386, 0.0482904, 0.0395639, 0.0540978, 0.0683364, 0.0540687, 0.0673417, 0.0114135, 0.0304133};

struct protein_h3reader *protein_h3reader_new(struct protein_cfg cfg, FILE *fp)
{
    struct protein_h3reader *r = calloc(1, sizeof *r);
    if (!r) return NULL;
    r->fp = fp;
    r->cfg = cfg;
    for (int i = 0; i < DCP_AMINO_SIZE; ++i) r->null_lprobs[i] = (float)log(swissprot_bg[i]);
    r->model = protein_model_new(cfg, r->null_lprobs);
    if (!r->model)
    {
        free(r);
        return NULL;
    }
    return r;
}

void protein_h3reader_del(struct protein_h3reader *r)
{
    if (!r) return;
    protein_model_del(r->model);
    free(r);
}

struct protein_model const *protein_h3reader_model(struct protein_h3reader const *r) { return r->model; }
char const *protein_h3reader_accession(struct protein_h3reader const *r) { return r->acc; }
char const *protein_h3reader_name(struct protein_h3reader const *r) { return r->name; }

static bool next_line(struct protein_h3reader *r)
{
    if (!fgets(r->line, sizeof r->line, r->fp)) return false;
    r->lineno++;
    size_t n = strlen(r->line);
    while (n && (r->line[n - 1] == '\n' || r->line[n - 1] == '\r')) r->line[--n] = '\0';
    return true;
}

/* -ln p -> ln p; '*' -> -inf.  Returns NULL on a malformed token. */
static char *take_score(char *p, float *out)
{
    while (*p == ' ' || *p == '\t') ++p;
    if (*p == '*')
    {
        *out = -INFINITY;
        return p + 1;
    }
    char *end = NULL;
    double v = strtod(p, &end);
    if (end == p) return NULL;
    *out = (float)(-v);
    return end;
}

static enum rc parse_error(struct protein_h3reader *r, char const *what)
{
    char msg[160];
    snprintf(msg, sizeof msg, "HMMER3 parse error at line %u: %s", r->lineno, what);
    return dcp_error(RC_EPARSE, msg);
}

static enum rc read_trans(struct protein_h3reader *r, struct protein_trans *t)
{
    if (!next_line(r)) return parse_error(r, "missing transition line");
    char *p = r->line;
    for (int i = 0; i < PROTEIN_TRANS_SIZE; ++i)
        if (!(p = take_score(p, &t->data[i]))) return parse_error(r, "bad transition score");
    return RC_OK;
}

enum rc protein_h3reader_next(struct protein_h3reader *r)
{
    /* header */
    bool seen = false;
    r->name[0] = r->acc[0] = '\0';
    r->leng = 0;
    for (;;)
    {
        if (!next_line(r)) return seen ? parse_error(r, "unexpected end of file in header") : RC_END;
        char *p = r->line;
        while (*p == ' ') ++p;
        if (!*p) continue;
        if (!seen)
        {
            if (strncmp(p, "HMMER3/", 7) != 0) return parse_error(r, "expected a HMMER3/ header");
            seen = true;
            continue;
        }
        if (!strncmp(p, "NAME", 4) && isspace((unsigned char)p[4])) sscanf(p + 4, " %63s", r->name);
        else if (!strncmp(p, "ACC", 3) && isspace((unsigned char)p[3])) sscanf(p + 3, " %31s", r->acc);
        else if (!strncmp(p, "LENG", 4) && isspace((unsigned char)p[4])) r->leng = (unsigned)strtoul(p + 4, NULL, 10);
        else if (!strncmp(p, "ALPH", 4))
        {
            char alph[32] = {0};
            sscanf(p + 4, " %31s", alph);
            if (strcmp(alph, "amino") != 0) return parse_error(r, "only amino-acid profiles are supported");
        }
        else if (!strncmp(p, "HMM", 3) && (p[3] == ' ' || p[3] == '\t'))
            break;
    }
    if (r->leng == 0) return parse_error(r, "LENG missing or zero");
    if (!r->acc[0])
    {
        /* no ACC line: the name stands in, cut to PROFILE_ACC_SIZE (limits.h:13) */
        memcpy(r->acc, r->name, sizeof r->acc - 1);
        r->acc[sizeof r->acc - 1] = '\0';
    }
    if (!next_line(r)) return parse_error(r, "missing transition header"); /* m->m m->i ... */

    enum rc rc = protein_model_setup(r->model, r->leng);
    if (rc) return rc;

    /* node 0: optional COMPO, insert emissions (ignored), transitions */
    if (!next_line(r)) return parse_error(r, "missing node 0");
    char *p = r->line;
    while (*p == ' ') ++p;
    if (!strncmp(p, "COMPO", 5))
        if (!next_line(r)) return parse_error(r, "missing node 0 insert line");
    struct protein_trans t;
    if ((rc = read_trans(r, &t))) return rc;
    if ((rc = protein_model_add_trans(r->model, t))) return rc;

    for (unsigned k = 1; k <= r->leng; ++k)
    {
        if (!next_line(r)) return parse_error(r, "missing match line");
        p = r->line;
        char *end = NULL;
        unsigned long idx = strtoul(p, &end, 10);
        if (end == p || idx != k) return parse_error(r, "unexpected node index");
        p = end;
        float match[DCP_AMINO_SIZE];
        for (int i = 0; i < DCP_AMINO_SIZE; ++i)
            if (!(p = take_score(p, &match[i]))) return parse_error(r, "bad match score");
        /* trailing fields: MAP CONS RF MM CS; the consensus residue is the second */
        char cons = '-';
        char f0[32] = {0}, f1[32] = {0};
        if (sscanf(p, " %31s %31s", f0, f1) == 2) cons = f1[0];
        if ((rc = protein_model_add_node(r->model, match, cons))) return rc;
        if (!next_line(r)) return parse_error(r, "missing insert line"); /* ignored, protein_model.c:126-127 */
        if ((rc = read_trans(r, &t))) return rc;
        if ((rc = protein_model_add_trans(r->model, t))) return rc;
    }
    if (!next_line(r)) return parse_error(r, "missing // terminator");
    p = r->line;
    while (*p == ' ') ++p;
    if (strncmp(p, "//", 2) != 0) return parse_error(r, "expected //");
    return RC_OK;
}
