/*
 * dcp_multi.cpp -- scans larger than one launch: batches tiled by device memory, and a database
 * sharded over several GPUs of one box (include/dcpgpu.h, "multi-device").
 *
 * What it replaces in the reference: scan_run's `omp parallel for` over profile partitions
 * (src/server/scan.c:239-250), the partitioning itself (src/db/profile_reader.c:54-72) and the ordered
 * concatenation of the per-thread product files (src/server/prod.c:106-145).  Here a partition is a GPU:
 *
 *   profile axis   profiles are split by modelled cost (dcp_shape.c), each device keeps its shard in HBM,
 *                  every device scans all sequences; the per-device hit lists are merged by
 *                  (sequence, global profile) -- the order a single device returns them in.
 *   sequence axis  few long profiles (config 4: 50 profiles over 8 GPUs) do not balance; then every device
 *                  holds the whole database and scans a contiguous range of the sequences, split by
 *                  cumulative length.
 *
 * One host thread per device drives its stream; there is no collective and no peer copy on the data path
 * (every pair is independent).  Results are deterministic and independent of the device count.
 */
#include "dcp_engine.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

enum rc dcp_scan_once(struct dcpgpu_db *db, unsigned nseqs, char const *const *seqs, unsigned const *lens,
                      struct dcpgpu_params const *prm, struct dcpgpu_result **out);

namespace
{
struct Part
{
    dcpgpu_result *res = nullptr;
    const std::vector<uint32_t> *gprofs = nullptr; /* local profile -> global profile; nullptr = identity */
    uint32_t seq0 = 0;
    int device = 0;
};

void drop(std::vector<Part> &parts)
{
    for (auto &p : parts) dcpgpu_result_del(p.res);
    parts.clear();
}

/* device bytes one scan of S sequences / T nucleotides needs besides the database (dcp_engine.cu) */
size_t scan_bytes(const dcpgpu_db *db, size_t S, size_t T)
{
    const size_t P = db->profs.size(), n_null = db->null_tabs.size(), recs = T + S;
    return n_null * recs * sizeof(RowRec) + recs * 2 + T + S * (sizeof(SeqMeta) + 64 + n_null * 4) + P * S * 5;
}

/*
 * Scan sequences [lo, hi) on one device, in as many launches as the device's free memory asks for: row
 * records are 64 B per nucleotide and null table, pair scores 5 B per (sequence, profile), and a batch is
 * capped at 2^30 pairs.  Appends one Part per launch.
 */
enum rc scan_range(dcpgpu_db *db, const std::vector<uint32_t> *gprofs, uint32_t lo, uint32_t hi,
                   char const *const *seqs, unsigned const *lens, dcpgpu_params const *prm, std::vector<Part> &out)
{
    CU_TRY(cudaSetDevice(db->device));
    /* what was free after commit less what the pool has handed out (cudaMemGetInfo costs up to 20 ms per call with a
     * multi-GB pool, dcp_trace.cu) */
    uint64_t used = 0;
    cudaMemPoolGetAttribute(db->pool, cudaMemPoolAttrUsedMemCurrent, &used);
    const size_t avail = db->free_at_commit > used ? db->free_at_commit - (size_t)used : 0;
    /* 40 % of what is free, less 64 MB for the small buffers of a scan; the traceback pass sizes itself later */
    size_t budget = std::max<size_t>((size_t)(0.4 * (double)avail), (size_t)320 << 20) - ((size_t)64 << 20);
    if (const char *e = getenv("DCPGPU_SCAN_BUDGET_KB")) budget = (size_t)std::max(1.0, atof(e)) << 10; /* test knob */
    const size_t max_pairs = (size_t)1 << 30, P = db->profs.size();
    uint32_t a = lo;
    while (a < hi)
    {
        size_t S = 0, T = 0;
        uint32_t b = a;
        while (b < hi)
        {
            const size_t S1 = S + 1, T1 = T + lens[b];
            if (S > 0 && (scan_bytes(db, S1, T1) > budget || S1 * P > max_pairs)) break;
            S = S1, T = T1, ++b;
        }
        dcpgpu_result *res = nullptr;
        enum rc rc = dcp_scan_once(db, b - a, seqs + a, lens + a, prm, &res);
        if (rc) return rc;
        out.push_back({res, gprofs, a, db->device}); /* dcpgpu_scan_resident has fired prm->progress */
        a = b;
    }
    return RC_OK;
}

/*
 * Merge per-launch results into one dcpgpu_result over `nseq` sequences x `nprof` (global) profiles.
 * Hits are ordered by (sequence, global profile) -- prod_fclose concatenates the thread files in thread
 * order (prod.c:119-141), which for one sequence at a time is exactly this order.  `parallel`: the parts
 * ran side by side on different devices (phase times are the maximum), otherwise one after the other (sums).
 */
enum rc merge_parts(dcpgpu_db *view, std::vector<Part> &parts, uint32_t nseq, uint32_t nprof, bool want_paths,
                    bool parallel, dcpgpu_result **out)
{
    if (parts.size() == 1 && !parts[0].gprofs && parts[0].seq0 == 0 && parts[0].res->db == view)
    {
        *out = parts[0].res; /* one launch on the database itself: nothing to merge */
        parts.clear();
        return RC_OK;
    }
    dcpgpu_result *r = new (std::nothrow) dcpgpu_result;
    if (!r) return dcp_error(RC_ENOMEM, "alloc result");
    r->db = view, r->nseq = nseq, r->nprof = nprof, r->n_null = 0;
    struct Key
    {
        uint32_t seq, prof, part;
        uint64_t idx;
    };
    std::vector<Key> keys;
    size_t nsteps = 0;
    for (uint32_t k = 0; k < parts.size(); ++k)
    {
        const dcpgpu_result *q = parts[k].res;
        for (uint64_t i = 0; i < q->hits.size(); ++i)
        {
            const HitRec &h = q->hits[i];
            keys.push_back({parts[k].seq0 + h.seq, parts[k].gprofs ? (*parts[k].gprofs)[h.prof] : h.prof, k, i});
            nsteps += h.nsteps;
        }
    }
    std::sort(keys.begin(), keys.end(), [](const Key &a, const Key &b) {
        return a.seq != b.seq ? a.seq < b.seq : a.prof < b.prof;
    });
    r->hits.resize(keys.size()), r->hit_alt.resize(keys.size()), r->hit_null.resize(keys.size());
    r->steps.reserve(nsteps);
    for (size_t i = 0; i < keys.size(); ++i)
    {
        const dcpgpu_result *q = parts[keys[i].part].res;
        const HitRec &h = q->hits[keys[i].idx];
        r->hits[i] = {keys[i].seq, keys[i].prof, r->steps.size(), h.nsteps};
        r->hit_alt[i] = q->hit_alt[keys[i].idx], r->hit_null[i] = q->hit_null[keys[i].idx];
        if (q->have_paths) r->steps.insert(r->steps.end(), q->steps.begin() + h.step_off, q->steps.begin() + h.step_off + h.nsteps);
    }
    r->have_paths = want_paths;
    dcpgpu_timing &t = r->timing;
    for (auto &p : parts)
    {
        const dcpgpu_timing &u = p.res->timing;
        if (parallel)
        {
            t.prep_ms = std::max(t.prep_ms, u.prep_ms), t.score_ms = std::max(t.score_ms, u.score_ms);
            t.trace_ms = std::max(t.trace_ms, u.trace_ms), t.total_ms = std::max(t.total_ms, u.total_ms);
        }
        else
            t.prep_ms += u.prep_ms, t.score_ms += u.score_ms, t.trace_ms += u.trace_ms, t.total_ms += u.total_ms;
        t.launches += u.launches, t.alt_cells += u.alt_cells, t.h2d_bytes += u.h2d_bytes, t.d2h_bytes += u.d2h_bytes;
        /* the hit lists now live in the merged result; the parts keep only what a lazy fetch of the pair matrices needs */
        std::vector<HitRec>().swap(p.res->hits);
        std::vector<dcp_step>().swap(p.res->steps);
        std::vector<float>().swap(p.res->hit_alt), std::vector<float>().swap(p.res->hit_null);
        r->parts.push_back(p.res), r->part_profs.push_back(p.gprofs), r->part_seq0.push_back(p.seq0);
        r->part_device.push_back(p.device);
    }
    parts.clear();
    *out = r;
    return RC_OK;
}

/* per-device sequential launches of one device are summed before the devices are compared */
void fold_device_timing(std::vector<Part> &parts, size_t first)
{
    if (parts.size() - first < 2) return;
    dcpgpu_timing &t = parts[first].res->timing;
    for (size_t k = first + 1; k < parts.size(); ++k)
    {
        dcpgpu_timing &u = parts[k].res->timing;
        t.prep_ms += u.prep_ms, t.score_ms += u.score_ms, t.trace_ms += u.trace_ms, t.total_ms += u.total_ms;
        u.prep_ms = u.score_ms = u.trace_ms = u.total_ms = 0.0f;
    }
}
} // namespace

/* thread_run's loop over a batch from host buffers (scan_thread.c:86-135), tiled by device memory */
extern "C" enum rc dcpgpu_scan(struct dcpgpu_db *db, unsigned nseqs, char const *const *seqs, unsigned const *lens,
                               struct dcpgpu_params const *prm, struct dcpgpu_result **out)
{
    if (db->device < 0) return dcp_error(RC_EINVAL, "a multi-device view is scanned through dcpgpu_mdb_scan");
    if (!db->committed) return dcp_error(RC_EFAIL, "commit the database first");
    if (nseqs == 0) return dcp_error(RC_EINVAL, "no sequences");
    std::vector<Part> parts;
    enum rc rc = scan_range(db, nullptr, 0, nseqs, seqs, lens, prm, parts);
    if (!rc) rc = merge_parts(db, parts, nseqs, (uint32_t)db->profs.size(), prm->want_paths, false, out);
    drop(parts);
    return rc;
}

/* ----------------------------------------------------------------------------------------- */
/* multi-device database                                                                     */
/* ----------------------------------------------------------------------------------------- */
struct dcpgpu_mdb
{
    dcpgpu_db *view = nullptr;           /* every profile, global order; owns them */
    std::vector<dcpgpu_db *> shards;     /* one per device */
    std::vector<std::vector<uint32_t>> gprofs; /* per device: global index of its local profiles (ascending) */
    std::vector<uint32_t> shard_of;      /* profile axis: device index of each profile */
    enum dcpgpu_axis axis = DCPGPU_AXIS_AUTO;
    bool committed = false;
    double imbalance = 0.0; /* modelled: max shard cost / mean shard cost */
};

extern "C" enum rc dcpgpu_mdb_new(struct dcpgpu_mdb **out, unsigned ndevices, int const *devices)
{
    if (ndevices == 0) return dcp_error(RC_EINVAL, "no devices");
    dcpgpu_mdb *m = new (std::nothrow) dcpgpu_mdb;
    if (!m) return dcp_error(RC_ENOMEM, "alloc mdb");
    m->view = dcp_db_new_host();
    enum rc rc = m->view ? RC_OK : dcp_error(RC_ENOMEM, "alloc mdb");
    for (unsigned d = 0; d < ndevices && !rc; ++d)
    {
        /* a device may be listed more than once: its shards then share that GPU (each with its own stream and
         * host thread; the kernels never wait on one another) -- how the N-way path is tested on a one-GPU box */
        dcpgpu_db *db = nullptr;
        if (!rc) rc = dcpgpu_db_new(&db, devices[d]);
        if (!rc) db->owns_profs = false, m->shards.push_back(db);
    }
    if (rc)
    {
        dcpgpu_mdb_del(m);
        return rc;
    }
    m->gprofs.resize(ndevices);
    *out = m;
    return RC_OK;
}

extern "C" void dcpgpu_mdb_del(struct dcpgpu_mdb *m)
{
    if (!m) return;
    for (auto *db : m->shards) dcpgpu_db_del(db);
    dcpgpu_db_del(m->view);
    delete m;
}

extern "C" struct dcpgpu_db *dcpgpu_mdb_view(struct dcpgpu_mdb *m) { return m->view; }
extern "C" unsigned dcpgpu_mdb_ndevices(struct dcpgpu_mdb const *m) { return (unsigned)m->shards.size(); }
extern "C" unsigned dcpgpu_mdb_nprofiles(struct dcpgpu_mdb const *m) { return (unsigned)m->view->profs.size(); }
extern "C" enum dcpgpu_axis dcpgpu_mdb_axis(struct dcpgpu_mdb const *m) { return m->axis; }
extern "C" double dcpgpu_mdb_imbalance(struct dcpgpu_mdb const *m) { return m->imbalance; }
extern "C" int dcpgpu_mdb_device_of(struct dcpgpu_mdb const *m, unsigned profile)
{
    if (!m->committed || profile >= m->view->profs.size()) return -1;
    return m->axis == DCPGPU_AXIS_PROFILES ? m->shards[m->shard_of[profile]]->device : -1;
}
extern "C" uint64_t dcpgpu_mdb_device_bytes(struct dcpgpu_mdb const *m, unsigned shard)
{
    return shard < m->shards.size() ? m->shards[shard]->device_bytes : 0;
}

extern "C" enum rc dcpgpu_mdb_add(struct dcpgpu_mdb *m, struct protein_profile const *prof)
{
    if (m->committed) return dcp_error(RC_EFAIL, "database already committed");
    return dcpgpu_db_add(m->view, prof);
}

extern "C" enum rc dcpgpu_mdb_commit(struct dcpgpu_mdb *m, enum dcpgpu_axis axis)
{
    if (m->committed) return dcp_error(RC_EFAIL, "database already committed");
    const size_t nprof = m->view->profs.size(), ndev = m->shards.size();
    if (nprof == 0) return dcp_error(RC_EINVAL, "database is empty");
    std::vector<unsigned> sizes(nprof);
    for (size_t i = 0; i < nprof; ++i) sizes[i] = m->view->profs[i]->core_size;
    m->shard_of.assign(nprof, 0);
    enum rc rc = dcpgpu_shard_profiles((unsigned)nprof, sizes.data(), (unsigned)ndev, m->shard_of.data());
    if (rc) return rc;
    std::vector<double> load(ndev, 0.0);
    double total = 0.0;
    for (size_t i = 0; i < nprof; ++i)
    {
        const double c = dcp_profile_cost(sizes[i]);
        load[m->shard_of[i]] += c, total += c;
    }
    const double worst = *std::max_element(load.begin(), load.end());
    const double profile_imbalance = worst * (double)ndev / total;
    if (axis == DCPGPU_AXIS_AUTO)
        /* pick the axis with the better balance; sequences split to within one sequence */
        axis = (ndev > 1 && (nprof < ndev || profile_imbalance > 1.10)) ? DCPGPU_AXIS_SEQUENCES : DCPGPU_AXIS_PROFILES;
    m->axis = axis;
    m->imbalance = axis == DCPGPU_AXIS_PROFILES ? profile_imbalance : 1.0;
    for (size_t d = 0; d < ndev; ++d) m->gprofs[d].clear();
    for (size_t i = 0; i < nprof; ++i)
        for (size_t d = 0; d < ndev; ++d)
            if (axis == DCPGPU_AXIS_SEQUENCES || m->shard_of[i] == d)
            {
                rc = dcp_db_borrow(m->shards[d], m->view->profs[i]);
                if (rc) return rc;
                m->gprofs[d].push_back((uint32_t)i);
            }
    /* upload the shards side by side: one host thread per device */
    std::vector<enum rc> rcs(ndev, RC_OK);
    std::vector<std::string> errs(ndev);
    std::vector<std::thread> pool;
    for (size_t d = 0; d < ndev; ++d)
        pool.emplace_back([&, d]() {
            dcpgpu_db *db = m->shards[d];
            if (db->profs.empty()) return; /* more devices than profiles: this one stays idle */
            db->keep_host_tables = true;   /* the view owns the profiles */
            rcs[d] = dcpgpu_db_commit(db);
            if (rcs[d]) errs[d] = dcpgpu_last_error();
        });
    for (auto &t : pool) t.join();
    for (size_t d = 0; d < ndev; ++d)
        if (rcs[d]) return dcp_error(rcs[d], errs[d].c_str());
    for (auto *p : m->view->profs)
    {
        free(p->match_emission);
        p->match_emission = nullptr;
    }
    m->committed = true;
    return RC_OK;
}

extern "C" enum rc dcpgpu_mdb_scan(struct dcpgpu_mdb *m, unsigned nseqs, char const *const *seqs,
                                   unsigned const *lens, struct dcpgpu_params const *prm, struct dcpgpu_result **out)
{
    if (!m->committed) return dcp_error(RC_EFAIL, "commit the database first");
    if (nseqs == 0) return dcp_error(RC_EINVAL, "no sequences");
    const size_t ndev = m->shards.size();
    std::vector<unsigned> bounds(ndev + 1, 0);
    if (m->axis == DCPGPU_AXIS_SEQUENCES)
    {
        enum rc rc = dcpgpu_shard_sequences(nseqs, lens, (unsigned)ndev, bounds.data());
        if (rc) return rc;
    }
    std::vector<std::vector<Part>> per_dev(ndev);
    std::vector<enum rc> rcs(ndev, RC_OK);
    std::vector<std::string> errs(ndev);
    /* the progress callback is fired from the device threads: serialise it like progress_consume's omp critical
     * (src/core/progress.c:82) */
    struct Gate
    {
        dcpgpu_params const *prm;
        std::mutex mu;
    } gate{prm, {}};
    dcpgpu_params local = *prm;
    local.user = &gate;
    local.progress = prm->progress ? +[](void *g, uint64_t pairs) {
        Gate *x = (Gate *)g;
        std::lock_guard<std::mutex> lk(x->mu);
        x->prm->progress(x->prm->user, pairs);
    } : nullptr;
    std::vector<std::thread> pool;
    for (size_t d = 0; d < ndev; ++d)
        pool.emplace_back([&, d]() {
            dcpgpu_db *db = m->shards[d];
            if (!db->committed) return;
            uint32_t lo = 0, hi = nseqs;
            if (m->axis == DCPGPU_AXIS_SEQUENCES) lo = bounds[d], hi = bounds[d + 1];
            if (lo >= hi) return;
            rcs[d] = scan_range(db, m->axis == DCPGPU_AXIS_SEQUENCES ? nullptr : &m->gprofs[d], lo, hi, seqs, lens,
                                &local, per_dev[d]);
            if (rcs[d]) errs[d] = dcpgpu_last_error();
        });
    for (auto &t : pool) t.join();
    std::vector<Part> parts;
    enum rc rc = RC_OK;
    for (size_t d = 0; d < ndev; ++d)
    {
        if (rcs[d] && !rc) rc = dcp_error(rcs[d], errs[d].c_str());
        fold_device_timing(per_dev[d], 0);
        for (auto &p : per_dev[d]) parts.push_back(p);
    }
    if (!rc) rc = merge_parts(m->view, parts, nseqs, (uint32_t)m->view->profs.size(), prm->want_paths, true, out);
    drop(parts);
    return rc;
}
