/*
 * dcp_products.cpp -- product rows (include/dcpgpu.h Part 3).
 *
 * prod_fwrite + write_begin (src/server/prod.c:13-41,153-181) and
 * protein_match_write_func (src/server/protein_match.c:21-56): one TSV row per hit,
 *   scan_id seq_id profile_name abc_name alt_loglik null_loglik profile_typeid version match
 * with match = steps joined by ';', each step "frag,state,codon,amino" (mute steps print
 * empty frag/codon/amino).  abc_name is imm_abc_typeid_name(IMM_DNA) = "dna",
 * profile_typeid "protein" (src/model/profile_types.c), version 0.1.0 (CMakeLists.txt:5).
 */
#include "dcp_engine.h"

#include <cinttypes>
#include <cstdio>
#include <cstring>
#include <string>

static const char kHeader[] = "scan_id\tseq_id\tprofile_name\tabc_name\talt_loglik\tnull_loglik\tprofile_typeid\tversion\tmatch\n";

extern "C" enum rc dcpgpu_prod_fwrite_header(FILE *fp)
{
    return fputs(kHeader, fp) < 0 ? dcp_error(RC_EIO, "failed to write header") : RC_OK;
}

static enum rc build_row(const dcpgpu_result *r, const dcpgpu_db *db, uint64_t hit, int64_t scan_id, int64_t seq_id,
                         const char *seq, std::string &row)
{
    if (!r->have_paths) return dcp_error(RC_EINVAL, "scan ran without want_paths");
    if (hit >= r->hits.size()) return dcp_error(RC_EINVAL, "hit index out of range");
    const HitRec &h = r->hits[hit];
    const protein_profile *prof = dcp_db_profile(db, h.prof);
    char head[256];
    /* %.17g of the imm_float promoted to double, prod.c:29-31 */
    snprintf(head, sizeof head, "%" PRId64 "\t%" PRId64 "\t%s\t%s\t%.17g\t%.17g\t%s\t%s\t", scan_id, seq_id,
             prof->accession, "dna", (double)r->hit_alt[hit], (double)r->hit_null[hit], "protein", "0.1.0");
    row.assign(head);
    const dcp_step *steps = r->steps.data() + h.step_off;
    unsigned start = 0;
    for (unsigned i = 0; i < h.nsteps; ++i)
    {
        if (i > 0) row.push_back(';'); /* prod.c:168-171 */
        char name[DCP_STATE_NAME_SIZE];
        protein_state_name(steps[i].state_id, name);
        char codon[4] = {0}, amino[2] = {0};
        unsigned len = steps[i].seqlen;
        if (!protein_state_is_mute(steps[i].state_id))
        {
            enum rc rc = protein_profile_decode(prof, seq + start, len, steps[i].state_id, codon, amino);
            if (rc) return dcp_error(RC_EIO, "failed to write match"); /* protein_match.c:54-55 */
        }
        row.append(seq + start, len);
        row.push_back(',');
        row.append(name);
        row.push_back(',');
        row.append(codon);
        row.push_back(',');
        row.append(amino);
        start += len;
    }
    row.push_back('\n');
    return RC_OK;
}

extern "C" long dcpgpu_prod_row(struct dcpgpu_result const *r, struct dcpgpu_db const *db, uint64_t hit,
                                int64_t scan_id, int64_t seq_id, char const *seq, char *out, long cap)
{
    std::string row;
    if (build_row(r, db, hit, scan_id, seq_id, seq, row)) return -1;
    if ((long)row.size() + 1 > cap) return -1;
    memcpy(out, row.c_str(), row.size() + 1);
    return (long)row.size();
}

extern "C" enum rc dcpgpu_prod_fwrite(struct dcpgpu_result const *r, struct dcpgpu_db const *db, FILE *fp,
                                      int64_t scan_id, int64_t const *seq_ids, unsigned nseqs,
                                      char const *const *seqs)
{
    if (nseqs != r->nseq) return dcp_error(RC_EINVAL, "sequence count differs from the scan");
    std::string row;
    for (uint64_t i = 0; i < r->hits.size(); ++i)
    {
        unsigned s = r->hits[i].seq;
        enum rc rc = build_row(r, db, i, scan_id, seq_ids ? seq_ids[s] : (int64_t)s, seqs[s], row);
        if (rc) return rc;
        if (fwrite(row.data(), 1, row.size(), fp) != row.size()) return dcp_error(RC_EIO, "failed to write prod");
    }
    return RC_OK;
}
