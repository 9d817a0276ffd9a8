/*
 * dcp_db.c -- the .dcp database container (MessagePack), writer and reader.
 *
 * Follows the layout the reference's code writes (src/db/writer.c:95-117,153-175,
 * src/db/protein_writer.c:56-96, src/model/protein_profile.c:338-400, src/model/nuclt_dist.c:5-23):
 *   map(2) { "header":   map(8) { magic_number=0xC6F1 (reference: 0xC6F0), profile_typeid=2, float_size=4, entry_dist,
 *                                 epsilon, abc, amino, profile_sizes },
 *            "profiles": array(N) of map(16) { accession, null, alt, core_size, consensus,
 *                                 R, S, N, B, E, J, C, T, null_ndist, alt_insert_ndist, alt_match_ndist } }
 * with keys checked positionally (src/core/expect.c:11-22).
 *
 * Deviation (documented in DESIGN.md): the reference stores imm-defined blobs under "abc", "amino",
 * "null" and "alt" (imm_abc_pack / imm_dp_pack) whose byte layout lives in imm 2.0.3, which is not in the
 * reference tree.  Here those four values are self-describing instead: the alphabets are strings, and
 * "null"/"alt" carry the explicit DP-level arrays the kernels consume (emission tables, the 7 transition
 * scores per node, entry scores).  profile_sizes is a plain array of uint32.  Files written here are read
 * back bit-exactly by this reader; they are not interchangeable with files pressed by the C reference, and
 * say so themselves: magic_number is 0xC6F1 here, and a file with the reference's 0xC6F0 is refused with
 * RC_EPARSE and a message naming the imm blobs.
 */
#include "dcp_internal.h"

#include <stdlib.h>
#include <string.h>

/* 0xC6F0 is the reference's magic (include/deciphon/db/types.h:11).  Files written here carry explicit arrays where
 * the reference stores imm blobs, so they get their own magic: each reader rejects the other's files at the first
 * header key instead of failing somewhere inside a profile. */
enum { DCP_MAGIC_REFERENCE = 0xC6F0, DCP_MAGIC = 0xC6F1, DCP_PROFILE_PROTEIN = 2 };

/* ---------------------------- MessagePack subset ---------------------------- */
static bool put(FILE *fp, void const *p, size_t n) { return fwrite(p, 1, n, fp) == n; }
static bool put_u8(FILE *fp, unsigned v)
{
    unsigned char b = (unsigned char)v;
    return put(fp, &b, 1);
}
static bool put_be(FILE *fp, uint64_t v, int bytes)
{
    unsigned char b[8];
    for (int i = 0; i < bytes; ++i) b[i] = (unsigned char)(v >> (8 * (bytes - 1 - i)));
    return put(fp, b, (size_t)bytes);
}
static bool w_uint(FILE *fp, uint64_t v)
{
    if (v < 128) return put_u8(fp, (unsigned)v);
    if (v <= 0xff) return put_u8(fp, 0xcc) && put_be(fp, v, 1);
    if (v <= 0xffff) return put_u8(fp, 0xcd) && put_be(fp, v, 2);
    if (v <= 0xffffffffu) return put_u8(fp, 0xce) && put_be(fp, v, 4);
    return put_u8(fp, 0xcf) && put_be(fp, v, 8);
}
static bool w_str(FILE *fp, char const *s)
{
    size_t n = strlen(s);
    bool ok = n < 32 ? put_u8(fp, 0xa0 | (unsigned)n)
                     : n <= 0xff ? (put_u8(fp, 0xd9) && put_be(fp, n, 1)) : (put_u8(fp, 0xda) && put_be(fp, n, 2));
    return ok && n <= 0xffff && put(fp, s, n);
}
static bool w_map(FILE *fp, unsigned n) { return n < 16 ? put_u8(fp, 0x80 | n) : (put_u8(fp, 0xde) && put_be(fp, n, 2)); }
static bool w_array(FILE *fp, uint32_t n)
{
    if (n < 16) return put_u8(fp, 0x90 | n);
    if (n <= 0xffff) return put_u8(fp, 0xdc) && put_be(fp, n, 2);
    return put_u8(fp, 0xdd) && put_be(fp, n, 4);
}
static bool w_bin(FILE *fp, void const *p, size_t n)
{
    bool ok = n <= 0xff ? (put_u8(fp, 0xc4) && put_be(fp, n, 1))
                        : n <= 0xffff ? (put_u8(fp, 0xc5) && put_be(fp, n, 2)) : (put_u8(fp, 0xc6) && put_be(fp, n, 4));
    return ok && n <= 0xffffffffu && put(fp, p, n);
}
static bool w_f32(FILE *fp, float v)
{
    uint32_t u;
    memcpy(&u, &v, 4);
    return put_u8(fp, 0xca) && put_be(fp, u, 4);
}

static bool get(FILE *fp, void *p, size_t n) { return fread(p, 1, n, fp) == n; }
static bool get_be(FILE *fp, int bytes, uint64_t *v)
{
    unsigned char b[8];
    if (!get(fp, b, (size_t)bytes)) return false;
    *v = 0;
    for (int i = 0; i < bytes; ++i) *v = (*v << 8) | b[i];
    return true;
}
static bool r_uint(FILE *fp, uint64_t *v)
{
    unsigned char t;
    if (!get(fp, &t, 1)) return false;
    if (t < 128) return *v = t, true;
    if (t == 0xcc) return get_be(fp, 1, v);
    if (t == 0xcd) return get_be(fp, 2, v);
    if (t == 0xce) return get_be(fp, 4, v);
    if (t == 0xcf) return get_be(fp, 8, v);
    return false;
}
static bool r_str(FILE *fp, char *out, size_t cap)
{
    unsigned char t;
    uint64_t n;
    if (!get(fp, &t, 1)) return false;
    if ((t & 0xe0) == 0xa0) n = t & 0x1f;
    else if (t == 0xd9) { if (!get_be(fp, 1, &n)) return false; }
    else if (t == 0xda) { if (!get_be(fp, 2, &n)) return false; }
    else return false;
    if (n + 1 > cap) return false;
    if (!get(fp, out, (size_t)n)) return false;
    out[n] = '\0';
    return true;
}
static bool r_map(FILE *fp, unsigned *n)
{
    unsigned char t;
    uint64_t v;
    if (!get(fp, &t, 1)) return false;
    if ((t & 0xf0) == 0x80) return *n = t & 0x0f, true;
    if (t == 0xde && get_be(fp, 2, &v)) return *n = (unsigned)v, true;
    return false;
}
static bool r_array(FILE *fp, uint32_t *n)
{
    unsigned char t;
    uint64_t v;
    if (!get(fp, &t, 1)) return false;
    if ((t & 0xf0) == 0x90) return *n = t & 0x0f, true;
    if (t == 0xdc && get_be(fp, 2, &v)) return *n = (uint32_t)v, true;
    if (t == 0xdd && get_be(fp, 4, &v)) return *n = (uint32_t)v, true;
    return false;
}
static bool r_bin(FILE *fp, void *out, size_t expect)
{
    unsigned char t;
    uint64_t n;
    if (!get(fp, &t, 1)) return false;
    int bytes = t == 0xc4 ? 1 : t == 0xc5 ? 2 : t == 0xc6 ? 4 : 0;
    if (!bytes || !get_be(fp, bytes, &n) || n != expect) return false;
    return get(fp, out, expect);
}
static bool r_f32(FILE *fp, float *v)
{
    unsigned char t;
    uint64_t u;
    if (!get(fp, &t, 1) || t != 0xca || !get_be(fp, 4, &u)) return false;
    uint32_t w = (uint32_t)u;
    memcpy(v, &w, 4);
    return true;
}
/* expect_map_key (src/core/expect.c:11-22) */
static bool expect_key(FILE *fp, char const *key)
{
    char buf[32];
    return r_str(fp, buf, sizeof buf) && !strcmp(buf, key);
}

/* ---------------------------------- writer ---------------------------------- */
struct protein_db_writer
{
    FILE *fp, *tmp_profiles; /* profiles are spooled, the header needs their sizes first (writer.c:24-36) */
    struct protein_cfg cfg;
    unsigned nprofiles, cap;
    uint32_t *sizes;
};

struct protein_db_writer *protein_db_writer_open(FILE *fp, struct protein_cfg cfg)
{
    struct protein_db_writer *w = calloc(1, sizeof *w);
    if (!w) return NULL;
    w->fp = fp;
    w->cfg = cfg;
    w->tmp_profiles = tmpfile();
    if (!w->tmp_profiles)
    {
        free(w);
        dcp_set_error("create tmpfile");
        return NULL;
    }
    return w;
}

static bool w_ndist(FILE *fp, struct dcp_nuclt_dist const *nd)
{
    /* nuclt_dist_pack (nuclt_dist.c:5-13): array(2) [nuclt_lprob, codon_marg] */
    return w_array(fp, 2) && w_bin(fp, nd->nucltp, sizeof nd->nucltp) && w_bin(fp, nd->codonm, sizeof nd->codonm);
}

/* protein_profile_pack (protein_profile.c:338-400) */
enum rc protein_db_writer_pack_profile(struct protein_db_writer *w, struct protein_profile const *p)
{
    if (p->core_size == 0) return dcp_error(RC_EINVAL, "profile has not been absorbed");
    if (p->cfg.epsilon != w->cfg.epsilon || p->cfg.entry_dist != w->cfg.entry_dist)
        return dcp_error(RC_EINVAL, "profile configuration differs from the database header");
    FILE *fp = w->tmp_profiles;
    long start = ftell(fp);
    unsigned M = p->core_size;
    bool ok = w_map(fp, 16);
    ok = ok && w_str(fp, "accession") && w_str(fp, p->accession);
    ok = ok && w_str(fp, "null") && w_map(fp, 1) && w_str(fp, "emission") &&
         w_bin(fp, p->null_emission, sizeof p->null_emission);
    ok = ok && w_str(fp, "alt") && w_map(fp, 4) && w_str(fp, "trans") &&
         w_bin(fp, p->trans, (size_t)(M + 1) * sizeof *p->trans) && w_str(fp, "entry") &&
         w_bin(fp, p->entry, (size_t)M * sizeof(float)) && w_str(fp, "insert_emission") &&
         w_bin(fp, p->insert_emission, sizeof p->insert_emission) && w_str(fp, "match_emission") &&
         w_bin(fp, p->match_emission, (size_t)M * DCP_FRAME_TABLE_SIZE * sizeof(float));
    ok = ok && w_str(fp, "core_size") && w_uint(fp, M);
    ok = ok && w_str(fp, "consensus") && w_str(fp, p->consensus);
    /* state indices of the compiled DPs: null has R only; alt in canonical order S N B E J C T */
    static char const *const names[8] = {"R", "S", "N", "B", "E", "J", "C", "T"};
    static unsigned const index[8] = {0, 0, 1, 2, 3, 4, 5, 6};
    for (int i = 0; ok && i < 8; ++i) ok = w_str(fp, names[i]) && w_uint(fp, index[i]);
    ok = ok && w_str(fp, "null_ndist") && w_ndist(fp, &p->null_ndist);
    ok = ok && w_str(fp, "alt_insert_ndist") && w_ndist(fp, &p->insert_ndist);
    ok = ok && w_str(fp, "alt_match_ndist") && w_array(fp, M);
    for (unsigned k = 0; ok && k < M; ++k) ok = w_ndist(fp, &p->match_ndists[k]);
    if (!ok) return dcp_error(RC_EIO, "write profile");
    if (w->nprofiles == w->cap)
    {
        unsigned cap = w->cap ? 2 * w->cap : 256;
        uint32_t *s = realloc(w->sizes, cap * sizeof *s);
        if (!s) return dcp_error(RC_ENOMEM, "profile sizes");
        w->sizes = s, w->cap = cap;
    }
    w->sizes[w->nprofiles++] = (uint32_t)(ftell(fp) - start);
    return RC_OK;
}

/* db_writer_close (writer.c:95-117): root map, header (+ profile_sizes), profiles */
enum rc protein_db_writer_close(struct protein_db_writer *w, bool successfully)
{
    enum rc rc = RC_OK;
    if (successfully)
    {
        FILE *fp = w->fp;
        bool ok = w_map(fp, 2) && w_str(fp, "header") && w_map(fp, 8);
        ok = ok && w_str(fp, "magic_number") && w_uint(fp, DCP_MAGIC);
        ok = ok && w_str(fp, "profile_typeid") && w_uint(fp, DCP_PROFILE_PROTEIN);
        ok = ok && w_str(fp, "float_size") && w_uint(fp, sizeof(float));
        ok = ok && w_str(fp, "entry_dist") && w_uint(fp, (unsigned)w->cfg.entry_dist);
        ok = ok && w_str(fp, "epsilon") && w_f32(fp, w->cfg.epsilon);
        ok = ok && w_str(fp, "abc") && w_str(fp, "ACGT");                    /* imm_dna_iupac symbols */
        ok = ok && w_str(fp, "amino") && w_str(fp, "ACDEFGHIKLMNPQRSTVWY"); /* imm_amino_iupac symbols */
        ok = ok && w_str(fp, "profile_sizes") && w_array(fp, w->nprofiles);
        for (unsigned i = 0; ok && i < w->nprofiles; ++i) ok = w_uint(fp, w->sizes[i]);
        ok = ok && w_str(fp, "profiles") && w_array(fp, w->nprofiles);
        rewind(w->tmp_profiles);
        char buf[1 << 16];
        size_t n;
        while (ok && (n = fread(buf, 1, sizeof buf, w->tmp_profiles)) > 0) ok = put(fp, buf, n);
        if (!ok || fflush(fp)) rc = dcp_error(RC_EIO, "write database");
    }
    fclose(w->tmp_profiles);
    free(w->sizes);
    free(w);
    return rc;
}

/* ---------------------------------- reader ---------------------------------- */
struct protein_db_reader
{
    FILE *fp;
    struct protein_cfg cfg;
    unsigned nprofiles, next;
    uint32_t *sizes;
};

/* db_reader_open + protein_db_reader_open (src/db/reader.c:25-79, src/db/protein_reader.c:40-82) */
enum rc protein_db_reader_open(struct protein_db_reader **out, FILE *fp)
{
    struct protein_db_reader *r = calloc(1, sizeof *r);
    if (!r) return dcp_error(RC_ENOMEM, "alloc reader");
    r->fp = fp;
    unsigned n = 0;
    uint64_t v = 0;
    char sym[32];
    uint32_t count = 0;
    enum rc rc = RC_EPARSE;
    char const *why = "bad database header";
    if (!r_map(fp, &n) || n != 2 || !expect_key(fp, "header") || !r_map(fp, &n) || n != 8) goto fail;
    if (!expect_key(fp, "magic_number") || !r_uint(fp, &v)) goto fail;
    if (v == DCP_MAGIC_REFERENCE)
    {
        why = "database pressed by the C reference (imm_abc / imm_dp blobs): not supported, press the .hmm again with "
              "dcp-scan --press";
        goto fail;
    }
    if (v != DCP_MAGIC) { why = "wrong magic number"; goto fail; }
    if (!expect_key(fp, "profile_typeid") || !r_uint(fp, &v) || v != DCP_PROFILE_PROTEIN) { why = "not a protein database"; goto fail; }
    if (!expect_key(fp, "float_size") || !r_uint(fp, &v) || v != sizeof(float)) { why = "float_size must be 4"; goto fail; }
    if (!expect_key(fp, "entry_dist") || !r_uint(fp, &v)) goto fail;
    r->cfg.entry_dist = (enum entry_dist)v;
    if (!expect_key(fp, "epsilon") || !r_f32(fp, &r->cfg.epsilon)) goto fail;
    if (!expect_key(fp, "abc") || !r_str(fp, sym, sizeof sym) || strcmp(sym, "ACGT")) { why = "unsupported nucleotide alphabet"; goto fail; }
    if (!expect_key(fp, "amino") || !r_str(fp, sym, sizeof sym) || strcmp(sym, "ACDEFGHIKLMNPQRSTVWY")) { why = "unsupported amino alphabet"; goto fail; }
    if (!expect_key(fp, "profile_sizes") || !r_array(fp, &count)) goto fail;
    r->sizes = malloc((count ? count : 1) * sizeof *r->sizes);
    if (!r->sizes) { rc = RC_ENOMEM; why = "profile sizes"; goto fail; }
    for (uint32_t i = 0; i < count; ++i)
    {
        if (!r_uint(fp, &v)) goto fail;
        r->sizes[i] = (uint32_t)v;
    }
    uint32_t again = 0;
    if (!expect_key(fp, "profiles") || !r_array(fp, &again) || again != count) goto fail;
    r->nprofiles = count;
    *out = r;
    return RC_OK;
fail:
    free(r->sizes);
    free(r);
    return dcp_error(rc, why);
}

unsigned protein_db_reader_nprofiles(struct protein_db_reader const *r) { return r->nprofiles; }
struct protein_cfg protein_db_reader_cfg(struct protein_db_reader const *r) { return r->cfg; }
uint32_t protein_db_reader_profile_size(struct protein_db_reader const *r, unsigned i)
{
    return i < r->nprofiles ? r->sizes[i] : 0;
}

static bool r_ndist(FILE *fp, struct dcp_nuclt_dist *nd)
{
    uint32_t n;
    return r_array(fp, &n) && n == 2 && r_bin(fp, nd->nucltp, sizeof nd->nucltp) && r_bin(fp, nd->codonm, sizeof nd->codonm);
}

/* profile_reader_next + protein_profile unpack (profile_reader.c:156-168, protein_profile.c:38-117):
 * RC_END after the last profile.  The caller owns *out (protein_profile_del). */
enum rc protein_db_reader_next(struct protein_db_reader *r, struct protein_profile **out)
{
    if (r->next >= r->nprofiles) return RC_END;
    FILE *fp = r->fp;
    long start = ftell(fp);
    struct protein_profile *p = protein_profile_new("", r->cfg);
    if (!p) return dcp_error(RC_ENOMEM, "alloc profile");
    unsigned n = 0;
    uint64_t v = 0;
    char const *why = "malformed profile";
    if (!r_map(fp, &n) || n != 16) goto fail;
    if (!expect_key(fp, "accession") || !r_str(fp, p->accession, sizeof p->accession)) goto fail;
    if (!expect_key(fp, "null") || !r_map(fp, &n) || n != 1 || !expect_key(fp, "emission") ||
        !r_bin(fp, p->null_emission, sizeof p->null_emission))
        goto fail;
    /* "alt" comes before core_size in the reference's key order: read its sizes from the bin headers */
    if (!expect_key(fp, "alt") || !r_map(fp, &n) || n != 4 || !expect_key(fp, "trans")) goto fail;
    {
        unsigned char t;
        uint64_t len = 0;
        if (!get(fp, &t, 1)) goto fail;
        int bytes = t == 0xc4 ? 1 : t == 0xc5 ? 2 : t == 0xc6 ? 4 : 0;
        if (!bytes || !get_be(fp, bytes, &len) || len % sizeof(struct protein_trans) || len < 2 * sizeof(struct protein_trans))
            goto fail;
        unsigned M = (unsigned)(len / sizeof(struct protein_trans)) - 1;
        if (M > DCP_PROTEIN_MODEL_CORE_SIZE_MAX) { why = "profile is too long"; goto fail; } /* protein_profile.c:58 */
        p->core_size = M;
        p->trans = malloc(len);
        p->entry = malloc(M * sizeof(float));
        p->match_emission = malloc((size_t)M * DCP_FRAME_TABLE_SIZE * sizeof(float));
        p->match_ndists = malloc(M * sizeof *p->match_ndists);
        p->consensus = malloc(M + 1);
        if (!p->trans || !p->entry || !p->match_emission || !p->match_ndists || !p->consensus) { why = "alloc profile tables"; goto fail; }
        if (!get(fp, p->trans, len)) goto fail;
        if (!expect_key(fp, "entry") || !r_bin(fp, p->entry, M * sizeof(float))) goto fail;
        if (!expect_key(fp, "insert_emission") || !r_bin(fp, p->insert_emission, sizeof p->insert_emission)) goto fail;
        if (!expect_key(fp, "match_emission") || !r_bin(fp, p->match_emission, (size_t)M * DCP_FRAME_TABLE_SIZE * sizeof(float)))
            goto fail;
        if (!expect_key(fp, "core_size") || !r_uint(fp, &v) || v != M) goto fail;
        if (!expect_key(fp, "consensus") || !r_str(fp, p->consensus, M + 1)) goto fail;
    }
    static char const *const names[8] = {"R", "S", "N", "B", "E", "J", "C", "T"};
    for (int i = 0; i < 8; ++i)
        if (!expect_key(fp, names[i]) || !r_uint(fp, &v)) goto fail;
    if (!expect_key(fp, "null_ndist") || !r_ndist(fp, &p->null_ndist)) goto fail;
    if (!expect_key(fp, "alt_insert_ndist") || !r_ndist(fp, &p->insert_ndist)) goto fail;
    uint32_t cnt = 0;
    if (!expect_key(fp, "alt_match_ndist") || !r_array(fp, &cnt) || cnt != p->core_size) goto fail;
    for (unsigned k = 0; k < p->core_size; ++k)
        if (!r_ndist(fp, &p->match_ndists[k])) goto fail;
    if ((uint32_t)(ftell(fp) - start) != r->sizes[r->next]) { why = "profile size differs from the header"; goto fail; }
    r->next++;
    *out = p;
    return RC_OK;
fail:
    protein_profile_del(p);
    return dcp_error(RC_EPARSE, why);
}

void protein_db_reader_close(struct protein_db_reader *r)
{
    if (!r) return;
    free(r->sizes);
    free(r);
}
