/*
 * dcp-scan -- local, file-driven scan: HMMER3 profiles x FASTA nucleotide sequences -> product TSV.
 *
 * Stands in for the REST-driven scan_run (src/server/scan.c:215-269), which needs a live
 * deciphon-sched: press the .hmm in memory (hmm_press, src/server/hmm.c:120-178), scan every
 * sequence against every profile on the GPU, write prod_fclose's header and one prod_fwrite row
 * per hit (src/server/prod.c:106-181).  Sequence ids are 1-based file order.  With --devices the database is
 * sharded over several GPUs (dcpgpu_mdb_*, the counterpart of scan_run's thread partitions, scan.c:239-250);
 * the output is byte-identical whatever the device count.
 *
 *   dcp-scan [--single-hit] [--hmmer3-compat] [--lrt X] [--epsilon E] [--uniform-entry]
 *            [--device N | --devices A,B,...] [--axis auto|profiles|sequences] [--scan-id N] [--batch N]
 *            profiles.{hmm,dcp} sequences.fasta > products.tsv
 *   dcp-scan --press [--epsilon E] [--uniform-entry] profiles.hmm database.dcp       (hmm_press to a file)
 */
#include "dcpgpu.h"

#include <ctype.h>
#include <stdlib.h>
#include <string.h>

struct seqs
{
    char **seq;
    unsigned *len;
    int64_t *id;
    unsigned n, cap;
};

static int push_seq(struct seqs *s, char *buf, unsigned len, int64_t id)
{
    if (s->n == s->cap)
    {
        unsigned cap = s->cap ? 2 * s->cap : 1024;
        char **a = realloc(s->seq, cap * sizeof *a);
        unsigned *b = a ? realloc(s->len, cap * sizeof *b) : NULL;
        int64_t *c = b ? realloc(s->id, cap * sizeof *c) : NULL;
        if (a) s->seq = a;
        if (b) s->len = b;
        if (c) s->id = c;
        if (!a || !b || !c) return -1;
        s->cap = cap;
    }
    s->seq[s->n] = buf, s->len[s->n] = len, s->id[s->n] = id;
    s->n++;
    return 0;
}

/* reads up to `max` records; returns 1 if more may follow, 0 at end of file, -1 on error */
static int read_fasta(FILE *fp, struct seqs *s, unsigned max, int64_t *next_id, int *pending)
{
    char *cur = NULL;
    size_t len = 0, cap = 0;
    int c, have = *pending;
    *pending = 0;
    for (;;)
    {
        c = fgetc(fp);
        if (c == '>' || c == EOF)
        {
            if (have)
            {
                if (push_seq(s, cur ? cur : calloc(1, 1), (unsigned)len, (*next_id)++)) return -1;
                cur = NULL, len = cap = 0, have = 0;
            }
            if (c == EOF) return 0;
            while ((c = fgetc(fp)) != EOF && c != '\n') {}
            have = 1;
            if (s->n >= max)
            {
                *pending = 1;
                return 1;
            }
            continue;
        }
        if (isspace(c)) continue;
        if (!have) return -1; /* sequence data before any header */
        if (len + 2 > cap)
        {
            cap = cap ? 2 * cap : 4096;
            char *t = realloc(cur, cap);
            if (!t) return -1;
            cur = t;
        }
        cur[len++] = (char)toupper(c);
        cur[len] = '\0';
    }
}

static int ends_with(char const *s, char const *suffix)
{
    size_t n = strlen(s), m = strlen(suffix);
    return n >= m && !strcmp(s + n - m, suffix);
}

/* hmm_press (src/server/hmm.c:120-178) to a local .dcp file; no GPU involved */
static int press(char const *hmm_path, char const *dcp_path, struct protein_cfg cfg)
{
    FILE *hmm = fopen(hmm_path, "r"), *out = fopen(dcp_path, "wb");
    if (!hmm || !out)
    {
        fprintf(stderr, "dcp-scan: cannot open %s\n", hmm ? dcp_path : hmm_path);
        return 1;
    }
    struct protein_h3reader *rd = protein_h3reader_new(cfg, hmm);
    struct protein_db_writer *w = protein_db_writer_open(out, cfg);
    enum rc rc = (rd && w) ? RC_OK : RC_ENOMEM;
    unsigned n = 0;
    while (!rc && (rc = protein_h3reader_next(rd)) == RC_OK)
    {
        struct protein_profile *p = protein_profile_new(protein_h3reader_accession(rd), cfg);
        rc = p ? protein_profile_absorb(p, protein_h3reader_model(rd)) : RC_ENOMEM;
        if (!rc) rc = protein_db_writer_pack_profile(w, p);
        protein_profile_del(p);
        if (!rc) ++n;
    }
    if (rc == RC_END) rc = RC_OK;
    enum rc rc2 = w ? protein_db_writer_close(w, rc == RC_OK) : RC_OK;
    protein_h3reader_del(rd);
    fclose(hmm), fclose(out);
    if (rc || rc2)
    {
        fprintf(stderr, "dcp-scan: %s (rc=%d)\n", dcpgpu_last_error(), (int)(rc ? rc : rc2));
        return 1;
    }
    fprintf(stderr, "dcp-scan: pressed %u profiles into %s\n", n, dcp_path);
    return 0;
}

/* load a pressed database: protein_db_reader_open + profile_reader_next per profile */
static enum rc load_dcp(struct dcpgpu_db *db, FILE *fp, unsigned *nprof)
{
    struct protein_db_reader *rd = NULL;
    enum rc rc = protein_db_reader_open(&rd, fp);
    if (rc) return rc;
    struct protein_profile *p = NULL;
    while ((rc = protein_db_reader_next(rd, &p)) == RC_OK)
    {
        rc = dcpgpu_db_add(db, p);
        protein_profile_del(p);
        if (rc) break;
        ++*nprof;
    }
    protein_db_reader_close(rd);
    return rc == RC_END ? RC_OK : rc;
}

static void usage(void)
{
    fputs("usage: dcp-scan [--single-hit] [--hmmer3-compat] [--lrt X] [--epsilon E] [--uniform-entry]\n"
          "                [--device N | --devices A,B,...] [--axis auto|profiles|sequences] [--scan-id N] [--batch N]\n"
          "                profiles.{hmm,dcp} sequences.fasta > products.tsv\n"
          "       dcp-scan --press [--epsilon E] [--uniform-entry] profiles.hmm database.dcp\n",
          stderr);
}

int main(int argc, char **argv)
{
    struct dcpgpu_params prm = {.multi_hits = true, .hmmer3_compat = false, .lrt_threshold = 10.0, .want_paths = true};
    struct protein_cfg cfg = {ENTRY_DIST_OCCUPANCY, 0.01f}; /* PROTEIN_CFG_DEFAULT */
    int devices[64] = {0};
    unsigned ndev = 1;
    enum dcpgpu_axis axis = DCPGPU_AXIS_AUTO;
    int64_t scan_id = 1;
    unsigned batch = 65536;
    int do_press = 0;
    int i = 1;
    for (; i < argc && argv[i][0] == '-' && argv[i][1] == '-'; ++i)
    {
        if (!strcmp(argv[i], "--press")) do_press = 1;
        else if (!strcmp(argv[i], "--single-hit")) prm.multi_hits = false;
        else if (!strcmp(argv[i], "--hmmer3-compat")) prm.hmmer3_compat = true;
        else if (!strcmp(argv[i], "--uniform-entry")) cfg.entry_dist = ENTRY_DIST_UNIFORM;
        else if (!strcmp(argv[i], "--lrt") && i + 1 < argc) prm.lrt_threshold = atof(argv[++i]);
        else if (!strcmp(argv[i], "--epsilon") && i + 1 < argc) cfg.epsilon = (float)atof(argv[++i]);
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) devices[0] = atoi(argv[++i]), ndev = 1;
        else if (!strcmp(argv[i], "--devices") && i + 1 < argc)
        {
            ndev = 0;
            for (char *t = strtok(argv[++i], ","); t && ndev < 64; t = strtok(NULL, ",")) devices[ndev++] = atoi(t);
            if (ndev == 0)
            {
                usage();
                return 2;
            }
        }
        else if (!strcmp(argv[i], "--axis") && i + 1 < argc)
        {
            ++i;
            if (!strcmp(argv[i], "profiles")) axis = DCPGPU_AXIS_PROFILES;
            else if (!strcmp(argv[i], "sequences")) axis = DCPGPU_AXIS_SEQUENCES;
            else axis = DCPGPU_AXIS_AUTO;
        }
        else if (!strcmp(argv[i], "--scan-id") && i + 1 < argc) scan_id = atoll(argv[++i]);
        else if (!strcmp(argv[i], "--batch") && i + 1 < argc) batch = (unsigned)atoi(argv[++i]);
        else
        {
            usage();
            return 2;
        }
    }
    if (argc - i != 2 || batch == 0)
    {
        usage();
        return 2;
    }
    if (do_press) return press(argv[i], argv[i + 1], cfg);
    FILE *hmm = fopen(argv[i], ends_with(argv[i], ".dcp") ? "rb" : "r");
    FILE *fa = fopen(argv[i + 1], "r");
    if (!hmm || !fa)
    {
        fprintf(stderr, "dcp-scan: cannot open %s\n", hmm ? argv[i + 1] : argv[i]);
        return 1;
    }
    /* one device: a plain database; several: a multi-device one whose view takes the profiles */
    struct dcpgpu_db *db = NULL;
    struct dcpgpu_mdb *mdb = NULL;
    enum rc rc = ndev > 1 ? dcpgpu_mdb_new(&mdb, ndev, devices) : dcpgpu_db_new(&db, devices[0]);
    if (!rc && mdb) db = dcpgpu_mdb_view(mdb);
    unsigned nprof = 0;
    if (!rc) rc = ends_with(argv[i], ".dcp") ? load_dcp(db, hmm, &nprof) : dcpgpu_press_hmm(db, hmm, cfg, &nprof);
    if (!rc) rc = mdb ? dcpgpu_mdb_commit(mdb, axis) : dcpgpu_db_commit(db);
    fclose(hmm);
    if (rc)
    {
        fprintf(stderr, "dcp-scan: %s (rc=%d)\n", dcpgpu_last_error(), (int)rc);
        return 1;
    }
    dcpgpu_prod_fwrite_header(stdout);
    if (mdb)
        fprintf(stderr, "dcp-scan: %u devices, %s axis, modelled shard imbalance %.3f\n", ndev,
                dcpgpu_mdb_axis(mdb) == DCPGPU_AXIS_PROFILES ? "profile" : "sequence", dcpgpu_mdb_imbalance(mdb));
    /* --batch bounds the host memory of a FASTA chunk; the library tiles a chunk further by device memory */
    int64_t next_id = 1;
    int pending = 0, more = 1;
    uint64_t total_hits = 0, total_seqs = 0;
    while (more > 0)
    {
        struct seqs s = {0};
        more = read_fasta(fa, &s, batch, &next_id, &pending);
        if (more < 0)
        {
            fputs("dcp-scan: malformed FASTA or out of memory\n", stderr);
            return 1;
        }
        if (s.n)
        {
            struct dcpgpu_result *res = NULL;
            rc = mdb ? dcpgpu_mdb_scan(mdb, s.n, (char const *const *)s.seq, s.len, &prm, &res)
                     : dcpgpu_scan(db, s.n, (char const *const *)s.seq, s.len, &prm, &res);
            if (!rc) rc = dcpgpu_prod_fwrite(res, db, stdout, scan_id, s.id, s.n, (char const *const *)s.seq);
            if (rc)
            {
                fprintf(stderr, "dcp-scan: %s (rc=%d)\n", dcpgpu_last_error(), (int)rc);
                return 1;
            }
            total_hits += dcpgpu_result_nhits(res);
            total_seqs += s.n;
            dcpgpu_result_del(res);
        }
        for (unsigned k = 0; k < s.n; ++k) free(s.seq[k]);
        free(s.seq), free(s.len), free(s.id);
    }
    fclose(fa);
    fprintf(stderr, "dcp-scan: %u profiles x %llu sequences, %llu hits\n", nprof, (unsigned long long)total_seqs,
            (unsigned long long)total_hits);
    if (mdb) dcpgpu_mdb_del(mdb);
    else dcpgpu_db_del(db);
    return 0;
}
