/*
 * dcp_press.cpp -- hmm_press without the REST plumbing (src/server/hmm.c:120-178):
 * read every profile of a HMMER3 file, absorb it (protein_profile_absorb) and add it to a
 * device database, accession taken from the ACC field (hmm.c:39).
 */
#include "dcp_engine.h"

extern "C" enum rc dcpgpu_press_hmm(struct dcpgpu_db *db, FILE *hmm, struct protein_cfg cfg, unsigned *nprofiles)
{
    struct protein_h3reader *rd = protein_h3reader_new(cfg, hmm);
    if (!rd) return dcp_error(RC_ENOMEM, "alloc h3reader");
    unsigned n = 0;
    enum rc rc;
    while ((rc = protein_h3reader_next(rd)) == RC_OK)
    {
        struct protein_profile *p = protein_profile_new(protein_h3reader_accession(rd), cfg);
        if (!p)
        {
            rc = dcp_error(RC_ENOMEM, "alloc profile");
            break;
        }
        rc = protein_profile_absorb(p, protein_h3reader_model(rd));
        if (!rc) rc = dcpgpu_db_add(db, p);
        protein_profile_del(p);
        if (rc) break;
        ++n;
    }
    protein_h3reader_del(rd);
    if (nprofiles) *nprofiles = n;
    return rc == RC_END ? RC_OK : rc;
}
