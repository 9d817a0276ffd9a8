/*
 * dcp_press.cpp -- hmm_press without the REST plumbing (src/server/hmm.c:120-178):
 * read every profile of a HMMER3 file, absorb it (protein_profile_absorb) and add it to a
 * device database, accession taken from the ACC field (hmm.c:39).
 *
 * The reference presses single-threaded (hmm.c:122 ignores num_threads).  Parsing stays sequential
 * here too, but the expensive step -- absorb = (M + 2) frame emission tables of 1364 entries each -- runs
 * on all host cores, a bounded batch of models at a time, and profiles are added in file order.
 */
#include "dcp_engine.h"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace
{
struct Job
{
    protein_model *model = nullptr; /* private copy of what absorb reads */
    protein_profile *prof = nullptr;
    enum rc rc = RC_OK;
};

protein_model *model_snapshot(protein_model const *m)
{
    protein_model *c = (protein_model *)malloc(sizeof *c);
    if (!c) return nullptr;
    *c = *m;
    c->match_ndists = (dcp_nuclt_dist *)malloc(m->core_size * sizeof *c->match_ndists);
    c->trans = (protein_trans *)malloc((m->core_size + 1) * sizeof *c->trans);
    if (!c->match_ndists || !c->trans)
    {
        free(c->match_ndists), free(c->trans), free(c);
        return nullptr;
    }
    memcpy(c->match_ndists, m->match_ndists, m->core_size * sizeof *c->match_ndists);
    memcpy(c->trans, m->trans, (m->core_size + 1) * sizeof *c->trans);
    return c;
}
} // namespace

extern "C" enum rc dcpgpu_press_hmm(struct dcpgpu_db *db, FILE *hmm, struct protein_cfg cfg, unsigned *nprofiles)
{
    struct protein_h3reader *rd = protein_h3reader_new(cfg, hmm);
    if (!rd) return dcp_error(RC_ENOMEM, "alloc h3reader");
    const unsigned nthreads = std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
    const size_t batch = (size_t)nthreads * 4;
    unsigned n = 0;
    enum rc rc = RC_OK;
    bool eof = false;
    while (!eof && rc == RC_OK)
    {
        std::vector<Job> jobs;
        while (jobs.size() < batch)
        {
            enum rc r = protein_h3reader_next(rd);
            if (r == RC_END)
            {
                eof = true;
                break;
            }
            if (r)
            {
                rc = r;
                break;
            }
            Job j;
            j.model = model_snapshot(protein_h3reader_model(rd));
            j.prof = protein_profile_new(protein_h3reader_accession(rd), cfg);
            if (!j.model || !j.prof) rc = dcp_error(RC_ENOMEM, "alloc profile");
            jobs.push_back(j);
            if (rc) break;
        }
        if (rc == RC_OK && !jobs.empty())
        {
            std::atomic<size_t> next{0};
            auto work = [&]() {
                for (size_t i; (i = next.fetch_add(1)) < jobs.size();)
                    jobs[i].rc = protein_profile_absorb(jobs[i].prof, jobs[i].model);
            };
            std::vector<std::thread> pool;
            for (unsigned t = 1; t < std::min<size_t>(nthreads, jobs.size()); ++t) pool.emplace_back(work);
            work();
            for (auto &t : pool) t.join();
        }
        for (auto &j : jobs)
        {
            if (rc == RC_OK && j.rc) rc = dcp_error(j.rc, "failed to absorb a profile");
            if (rc == RC_OK)
            {
                rc = dcp_db_adopt(db, j.prof); /* the database owns it now */
                if (rc == RC_OK) j.prof = nullptr, ++n;
            }
            protein_profile_del(j.prof);
            protein_model_del(j.model);
        }
    }
    protein_h3reader_del(rd);
    if (nprofiles) *nprofiles = n;
    return rc;
}
