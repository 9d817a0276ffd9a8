/*
 * dcp_engine.cu -- the sm_100a scan engine behind include/dcpgpu.h (Part 2).
 *
 * Replaces, for a whole batch of (sequence, profile) pairs, what the reference does per pair
 * in thread_run (src/server/scan_thread.c:86-135): protein_profile_setup, imm_dp_viterbi on the
 * null and alt DPs, xmath_lrt + threshold and the traceback that feeds prod_fwrite.
 *
 * Kernels (all fp32 max-plus, no tensor cores):
 *   k_rows       per (null table, sequence row): window codes + N/J/C/R and insert emissions
 *   k_null       1-state null Viterbi per (sequence, null table)
 *   k_score<Q>   alt Viterbi score, one warp per pair, Q consecutive core nodes per lane,
 *                last five rows of Tin_M/Tin_I in registers, D chain resolved exactly by lazy
 *                warp-shuffle propagation, E by redux.sync.max.f32 (CREDUX)
 *   k_lrt        LRT filter -> hit flags / hit list
 *   k_trace<Q>   same DP for hits only, with imm's (transition, source length) first-max
 *                tie-break and packed 11-bit backpointers per cell
 *   k_walk       backpointers -> (state_id, seqlen) steps
 *
 * Emission tables live in HBM transposed: [profile][code 0..1363][lane][QP] so that the 32 lanes
 * of a warp read one contiguous 32*QP*4-byte line per (row, length) with LDG.128.
 */
#include "dcp_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>


namespace
{

/* ----------------------------------------------------------------------------------------- */
/* prep kernels                                                                              */
/* ----------------------------------------------------------------------------------------- */
__global__ void k_rows(const uint8_t *__restrict__ bases, const SeqMeta *__restrict__ seqs, uint32_t nseq,
                       const float *__restrict__ null_tabs, const float *__restrict__ ins_tab,
                       uint32_t n_null, uint64_t total_recs, RowRec *__restrict__ rows,
                       uint16_t *__restrict__ wcodes)
{
    uint32_t s = blockIdx.x;
    if (s >= nseq) return;
    SeqMeta sm = seqs[s];
    const uint8_t *b = bases + sm.row_off;
    for (uint32_t j = threadIdx.x; j <= sm.len; j += blockDim.x)
    {
        /* seq[j-l] is the l-th least significant base pair of the window */
        uint32_t w = 0;
#pragma unroll
        for (int l = 1; l <= 5; ++l)
            if (j >= (uint32_t)l) w |= (uint32_t)b[j - l] << (2 * (l - 1));
        wcodes[sm.rec_off + j] = (uint16_t)w;
        uint32_t code[5];
        codes_of(w, code);
        for (uint32_t t = 0; t < n_null; ++t)
        {
            RowRec r;
#pragma unroll
            for (int l = 0; l < 5; ++l)
            {
                r.eN[l] = null_tabs[(size_t)t * kTab + code[l]];
                r.eI[l] = ins_tab[code[l]];
            }
            r.pad0[0] = r.pad0[1] = r.pad0[2] = r.pad1[0] = r.pad1[1] = r.pad1[2] = 0.0f;
            rows[(size_t)t * total_recs + sm.rec_off + j] = r;
        }
    }
}

/* imm_dp_viterbi on the null dp (scan_thread.c:115): V_R[j] = max_l Tin_R[j-l] + e_R(seq[j-l:j]) */
__global__ void k_null(const SeqMeta *__restrict__ seqs, uint32_t nseq, uint32_t n_null, uint64_t total_recs,
                       const RowRec *__restrict__ rows, const float *__restrict__ spec,
                       float *__restrict__ null_out)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)nseq * n_null) return;
    const uint32_t s = (uint32_t)(tid / n_null), t = (uint32_t)(tid % n_null);
    SeqMeta sm = seqs[s];
    const float RR = spec[(size_t)s * 16 + 6];
    const RowRec *r = rows + (size_t)t * total_recs + sm.rec_off;
    /* ring of Tin_R for the last five rows; Tin_R[0] = start lprob 0 */
    float tin[5] = {NEG_INF, NEG_INF, NEG_INF, NEG_INF, 0.0f};
    float v = NEG_INF;
    for (uint32_t j = 1; j <= sm.len; ++j)
    {
        float e[5];
        load_row_special(r + j, e);
        v = fmaxf(fmaxf(fmaxf(tin[4] + e[0], tin[3] + e[1]), fmaxf(tin[2] + e[2], tin[1] + e[3])), tin[0] + e[4]);
        tin[0] = tin[1], tin[1] = tin[2], tin[2] = tin[3], tin[3] = tin[4];
        tin[4] = v + RR;
    }
    null_out[(size_t)s * n_null + t] = v;
}

/*
 * Match emission tables, host layout [node][code] -> device layout [code][warp][half][lane][4] or [code][warp][lane][8]
 * (node k-1 = warp * 32 Q + lane * Q + sub sits in half sub/4, float sub%4 of its lane; pads are -inf).
 */
__global__ void k_layout(const float *__restrict__ raw, float *__restrict__ out, uint32_t M, uint32_t Q, uint32_t QP,
                         uint32_t W, uint32_t LN, bool whole)
{
    /* LN lanes per pair (32, or 16 for the half-warp classes): a line is [warp][half][LN lanes][4] (or, `whole`, [warp][LN lanes][8]) */
    const uint32_t ROW = LN * QP * W;
    const size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= (size_t)kTab * ROW) return;
    const uint32_t code = (uint32_t)(x / ROW), pos = (uint32_t)(x % ROW);
    const uint32_t warp = pos / (LN * QP), r = pos % (LN * QP);
    uint32_t half = r / (LN * 4), lane = (r % (LN * 4)) / 4, sub = half * 4 + (r % 4);
    if (whole) lane = r / QP, sub = r % QP; /* [lane][8]: one 256-bit load per lane (dcp_kernels.cuh: emis256) */
    float v = NEG_INF;
    if (sub < Q)
    {
        const uint32_t k = warp * LN * Q + lane * Q + sub;
        if (k < M) v = raw[(size_t)k * kTab + code];
    }
    out[x] = v;
}

/* xmath_lrt + threshold (scan_thread.c:121-123): hit iff finite and not (lrt < threshold) */
__global__ void k_lrt(const float *__restrict__ alt, const float *__restrict__ null_by_tab,
                      const ProfMeta *__restrict__ metas, uint32_t nseq, uint32_t nprof, uint32_t n_null,
                      double thr, uint8_t *__restrict__ hit, unsigned long long *__restrict__ nhits)
{
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = (size_t)nseq * nprof;
    if (idx >= n) return;
    uint32_t s = (uint32_t)(idx / nprof), pr = (uint32_t)(idx % nprof);
    float nl = null_by_tab[(size_t)s * n_null + metas[pr].null_id];
    float lrt = -2.0f * (nl - alt[idx]);
    bool h = isfinite(lrt) && !((double)lrt < thr);
    hit[idx] = h ? 1 : 0;
    if (h) atomicAdd(nhits, 1ULL);
}

__global__ void k_collect(const uint8_t *__restrict__ hit, size_t n, unsigned long long *__restrict__ cursor,
                          unsigned long long *__restrict__ list)
{
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    if (hit[idx]) list[atomicAdd(cursor, 1ULL)] = idx;
}


__global__ void k_gather(const unsigned long long *__restrict__ list, size_t nhits, const float *__restrict__ alt,
                         const float *__restrict__ null_by_tab, const ProfMeta *__restrict__ metas, uint32_t nprof,
                         uint32_t n_null, float *__restrict__ out_alt, float *__restrict__ out_null)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nhits) return;
    unsigned long long idx = list[i];
    uint32_t s = (uint32_t)(idx / nprof), pr = (uint32_t)(idx % nprof);
    out_alt[i] = alt[idx];
    out_null[i] = null_by_tab[(size_t)s * n_null + metas[pr].null_id];
}

} // namespace

/* ----------------------------------------------------------------------------------------- */
/* host objects                                                                              */
/* ----------------------------------------------------------------------------------------- */
static protein_profile *profile_clone(protein_profile const *src)
{
    protein_profile *p = (protein_profile *)calloc(1, sizeof *p);
    if (!p) return nullptr;
    *p = *src;
    unsigned n = src->core_size;
    p->consensus = (char *)malloc(n + 1);
    p->match_ndists = (dcp_nuclt_dist *)malloc(n * sizeof *p->match_ndists);
    p->match_emission = (float *)malloc((size_t)n * kTab * sizeof(float));
    p->trans = (protein_trans *)malloc((n + 1) * sizeof *p->trans);
    p->entry = (float *)malloc(n * sizeof(float));
    if (!p->consensus || !p->match_ndists || !p->match_emission || !p->trans || !p->entry)
    {
        protein_profile_del(p);
        return nullptr;
    }
    memcpy(p->consensus, src->consensus, n + 1);
    memcpy(p->match_ndists, src->match_ndists, n * sizeof *p->match_ndists);
    memcpy(p->match_emission, src->match_emission, (size_t)n * kTab * sizeof(float));
    memcpy(p->trans, src->trans, (n + 1) * sizeof *p->trans);
    memcpy(p->entry, src->entry, n * sizeof(float));
    return p;
}

extern "C" enum rc dcpgpu_db_new(struct dcpgpu_db **out, int device)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return dcp_error(RC_EFAIL, "no CUDA device: the scan engine has no CPU fallback");
    if (device < 0 || device >= ndev) return dcp_error(RC_EINVAL, "no such CUDA device");
    CU_TRY(cudaSetDevice(device));
    dcpgpu_db *db = new (std::nothrow) dcpgpu_db;
    if (!db) return dcp_error(RC_ENOMEM, "alloc db");
    db->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) db->sm_count = prop.multiProcessorCount;
    /* the engine's own memory pool: freed scratch blocks stay cached in it between scans, and the device's
     * default pool (shared with whatever else lives in the process) keeps its own release policy */
    cudaMemPoolProps pp = {};
    pp.allocType = cudaMemAllocationTypePinned;
    pp.handleTypes = cudaMemHandleTypeNone;
    pp.location.type = cudaMemLocationTypeDevice;
    pp.location.id = device;
    cudaError_t e = cudaMemPoolCreate(&db->pool, &pp);
    if (e == cudaSuccess)
    {
        uint64_t keep = UINT64_MAX;
        e = cudaMemPoolSetAttribute(db->pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&db->stream, cudaStreamNonBlocking);
    for (int i = 0; i < dcpgpu_db::kSide && e == cudaSuccess; ++i)
    {
        e = cudaStreamCreateWithFlags(&db->side[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&db->join_ev[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&db->fork_ev, cudaEventDisableTiming);
    if (e != cudaSuccess)
    {
        dcp_set_error(cudaGetErrorString(e));
        db->owns_profs = false;
        dcpgpu_db_del(db);
        return RC_EFAIL;
    }
    *out = db;
    return RC_OK;
}

/* the global profile list of a multi-device database: no device, never committed */
struct dcpgpu_db *dcp_db_new_host(void)
{
    dcpgpu_db *db = new (std::nothrow) dcpgpu_db;
    if (db) db->device = -1;
    return db;
}

enum take_mode { TAKE_COPY, TAKE_ADOPT, TAKE_BORROW };
static enum rc db_take(struct dcpgpu_db *db, struct protein_profile *prof, take_mode mode);

extern "C" enum rc dcpgpu_db_add(struct dcpgpu_db *db, struct protein_profile const *prof)
{
    return db_take(db, const_cast<protein_profile *>(prof), TAKE_COPY);
}

/* same as dcpgpu_db_add, but the database takes ownership of `prof` (no copy); used by the press */
enum rc dcp_db_adopt(struct dcpgpu_db *db, struct protein_profile *prof) { return db_take(db, prof, TAKE_ADOPT); }
/* no copy and no ownership: a device shard of a dcpgpu_mdb refers to the profiles of the mdb's view */
enum rc dcp_db_borrow(struct dcpgpu_db *db, struct protein_profile *prof) { return db_take(db, prof, TAKE_BORROW); }

static enum rc db_take(struct dcpgpu_db *db, struct protein_profile *prof, take_mode mode)
{
    const bool owned = mode != TAKE_COPY;
    if (db->committed) return dcp_error(RC_EFAIL, "database already committed");
    if (prof->core_size == 0) return dcp_error(RC_EINVAL, "profile has not been absorbed");
    if (prof->core_size > DCP_PROTEIN_MODEL_CORE_SIZE_MAX)
        return dcp_error(RC_EINVAL, "profile is too long"); /* limits.h:11, protein_profile.c:58 */
    if (db->epsilon >= 0.0f && db->epsilon != prof->cfg.epsilon)
        return dcp_error(RC_EINVAL, "all profiles of a database share one epsilon");
    if (!prof->match_emission) return dcp_error(RC_EINVAL, "profile tables were already released");
    /* every I_k shares one insert table per database (protein_model.c:126-127: zero log-odds + the db-wide
     * epsilon); a profile unpacked from a file that disagrees would be scored with the wrong one */
    if (!db->profs.empty() &&
        memcmp(db->profs[0]->insert_emission, prof->insert_emission, kTab * sizeof(float)) != 0)
        return dcp_error(RC_EINVAL, "all profiles of a database share one insert emission table");
    /* the score pass takes E[j] = max_k V_Mk[j]; that needs delete scores to be log-probabilities */
    for (unsigned i = 0; i <= prof->core_size; ++i)
        if (prof->trans[i].MD > 0.0f || prof->trans[i].DD > 0.0f)
            return dcp_error(RC_EINVAL, "MD/DD transition scores must be <= 0");
    protein_profile *copy = owned ? prof : profile_clone(prof);
    if (!copy) return dcp_error(RC_ENOMEM, "clone profile");
    db->epsilon = prof->cfg.epsilon;
    uint32_t id = UINT32_MAX;
    for (size_t t = 0; t < db->null_tabs.size(); ++t)
        if (!memcmp(db->null_tabs[t].data(), prof->null_emission, kTab * sizeof(float)))
        {
            id = (uint32_t)t;
            break;
        }
    if (id == UINT32_MAX)
    {
        id = (uint32_t)db->null_tabs.size();
        db->null_tabs.emplace_back(prof->null_emission, prof->null_emission + kTab);
    }
    db->profs.push_back(copy);
    db->null_id.push_back(id);
    return RC_OK;
}

static void kernel_shape(uint32_t M, uint32_t &Q, uint32_t &W, uint32_t &cls, uint32_t &LN)
{
    cls = dcp_kernel_class(M);
    const dcp_class *c = dcp_class_at(cls);
    Q = c->q, W = c->tw ? c->tw : 1, LN = c->tw ? 32 : 16; /* tw = 0: two pairs per warp, 16 lanes each */
}

namespace
{
/* staging of dcpgpu_db_commit: two pinned host buffers + two device buffers, released on every exit path */
struct CommitStage
{
    float *host[2] = {nullptr, nullptr};
    float *dev[2] = {nullptr, nullptr};
    cudaEvent_t freed[2] = {nullptr, nullptr};
    ~CommitStage()
    {
        for (int b = 0; b < 2; ++b)
        {
            if (host[b]) cudaFreeHost(host[b]);
            if (dev[b]) cudaFree(dev[b]);
            if (freed[b]) cudaEventDestroy(freed[b]);
        }
    }
};

void db_release_device(dcpgpu_db *db)
{
    cudaFree(db->d_emis), cudaFree(db->d_trans), cudaFree(db->d_metas);
    cudaFree(db->d_null_tabs), cudaFree(db->d_ins_tab);
    db->d_emis = db->d_trans = db->d_null_tabs = db->d_ins_tab = nullptr;
    db->d_metas = nullptr;
    for (int q = 0; q < kMaxClasses; ++q)
    {
        cudaFree(db->d_class[q]);
        db->d_class[q] = nullptr;
        db->class_list[q].clear();
    }
    db->metas.clear();
    db->device_bytes = 0;
}

enum rc db_commit(dcpgpu_db *db)
{
    size_t nprof = db->profs.size();
    db->metas.resize(nprof);
    uint64_t emis_floats = 0, trans_floats = 0;
    uint32_t max_M = 1;
    for (size_t i = 0; i < nprof; ++i)
    {
        uint32_t M = db->profs[i]->core_size;
        uint32_t Q, W, cls, LN;
        kernel_shape(M, Q, W, cls, LN);
        uint32_t QP = Q <= 4 ? 4 : 8;
        ProfMeta &m = db->metas[i];
        m.M = M, m.Q = Q, m.QP = QP, m.null_id = db->null_id[i];
        m.W = W, m.cls = cls, m.LN = LN, m.pad = 0;
        m.emis_off = emis_floats;
        m.trans_off = trans_floats;
        emis_floats += (uint64_t)kTab * LN * QP * W;
        trans_floats += (uint64_t)8 * LN * Q * W + 3 * W; /* + the carry bounds of the W warps (dcp_score_mw.cuh) */
        db->class_list[m.cls].push_back((uint32_t)i);
        max_M = std::max(max_M, M);
    }
    CU_TRY(cudaMalloc(&db->d_emis, emis_floats * sizeof(float)));
    CU_TRY(cudaMalloc(&db->d_trans, trans_floats * sizeof(float)));
    CU_TRY(cudaMalloc(&db->d_metas, nprof * sizeof(ProfMeta)));
    CU_TRY(cudaMalloc(&db->d_null_tabs, db->null_tabs.size() * kTab * sizeof(float)));
    CU_TRY(cudaMalloc(&db->d_ins_tab, kTab * sizeof(float)));
    db->device_bytes = (emis_floats + trans_floats + (db->null_tabs.size() + 1) * kTab) * sizeof(float) +
                       nprof * sizeof(ProfMeta);

    /* Upload the tables as the host holds them ([node][code], contiguous) through two pinned buffers and let
     * k_layout transpose them into the layout the profile's kernel class reads (dcp_kernels.cuh: emis256) on the device. */
    const size_t raw_max = (size_t)max_M * kTab;
    CommitStage stg;
    for (int b = 0; b < 2; ++b)
    {
        CU_TRY(cudaMallocHost(&stg.host[b], raw_max * sizeof(float)));
        CU_TRY(cudaMalloc(&stg.dev[b], raw_max * sizeof(float)));
        CU_TRY(cudaEventCreateWithFlags(&stg.freed[b], cudaEventDisableTiming));
    }
    std::vector<float> tr_all(trans_floats, NEG_INF);
    for (size_t i = 0; i < nprof; ++i)
    {
        const ProfMeta &m = db->metas[i];
        const protein_profile *p = db->profs[i];
        const int b = (int)(i & 1);
        if (i >= 2) CU_TRY(cudaEventSynchronize(stg.freed[b])); /* the copy out of host[b] has completed */
        const size_t raw = (size_t)m.M * kTab;
        memcpy(stg.host[b], p->match_emission, raw * sizeof(float));
        CU_TRY(cudaMemcpyAsync(stg.dev[b], stg.host[b], raw * sizeof(float), cudaMemcpyHostToDevice, db->stream));
        CU_TRY(cudaEventRecord(stg.freed[b], db->stream));
        const uint32_t ROW = m.LN * m.QP * m.W;
        const size_t out = (size_t)kTab * ROW;
        k_layout<<<(unsigned)((out + 255) / 256), 256, 0, db->stream>>>(stg.dev[b], db->d_emis + m.emis_off, m.M, m.Q,
                                                                         m.QP, m.W, m.LN,
                                                                         emis256(m.LN == 16 ? 0 : (int)m.W, (int)m.Q));
        const uint32_t NP = m.LN * m.Q * m.W;
        float *tr = tr_all.data() + m.trans_off;
        for (uint32_t k = 1; k <= m.M; ++k) /* node k, slot k-1 */
        {
            uint32_t n = k - 1;
            if (k >= 2)
            {
                const protein_trans &t = p->trans[k - 1];
                tr[0 * NP + n] = t.MM, tr[1 * NP + n] = t.IM, tr[2 * NP + n] = t.DM;
                tr[3 * NP + n] = t.MD, tr[4 * NP + n] = t.DD;
            }
            if (k <= m.M - 1)
            {
                const protein_trans &t = p->trans[k];
                tr[5 * NP + n] = t.MI, tr[6 * NP + n] = t.II;
            }
            tr[7 * NP + n] = p->entry[k - 1];
        }
        /* carry bound per warp: S = sum of D->D over the warp's nodes after its first, rounded up; the M->D and
         * D->D scores into its first node */
        float *cbnd = tr + 8 * NP;
        for (uint32_t w = 0; w < m.W; ++w)
        {
            const uint32_t n0 = w * m.LN * m.Q;
            double sum = 0.0;
            for (uint32_t n = n0 + 1; n < n0 + m.LN * m.Q; ++n) sum += (double)tr[4 * NP + n];
            cbnd[w] = std::isinf(sum) ? NEG_INF : std::nextafterf((float)sum, INFINITY);
            cbnd[m.W + w] = tr[3 * NP + n0], cbnd[2 * m.W + w] = tr[4 * NP + n0];
        }
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(db->stream));
    CU_TRY(cudaMemcpy(db->d_trans, tr_all.data(), trans_floats * sizeof(float), cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(db->d_metas, db->metas.data(), nprof * sizeof(ProfMeta), cudaMemcpyHostToDevice));
    for (size_t t = 0; t < db->null_tabs.size(); ++t)
        CU_TRY(cudaMemcpy(db->d_null_tabs + t * kTab, db->null_tabs[t].data(), kTab * sizeof(float),
                          cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(db->d_ins_tab, db->profs[0]->insert_emission, kTab * sizeof(float), cudaMemcpyHostToDevice));
    for (int q = 0; q < kMaxClasses; ++q)
        if (!db->class_list[q].empty())
        {
            CU_TRY(cudaMalloc(&db->d_class[q], db->class_list[q].size() * sizeof(uint32_t)));
            CU_TRY(cudaMemcpy(db->d_class[q], db->class_list[q].data(), db->class_list[q].size() * sizeof(uint32_t),
                              cudaMemcpyHostToDevice));
        }
    return RC_OK;
}
} // namespace

extern "C" enum rc dcpgpu_db_commit(struct dcpgpu_db *db)
{
    if (db->device < 0) return dcp_error(RC_EINVAL, "a multi-device view is committed through dcpgpu_mdb_commit");
    if (db->committed) return dcp_error(RC_EFAIL, "database already committed");
    if (db->profs.empty()) return dcp_error(RC_EINVAL, "database is empty");
    CU_TRY(cudaSetDevice(db->device));
    enum rc rc = db_commit(db);
    if (rc)
    {
        /* leave the database as it was before the call (a retry starts from clean class lists) */
        db_release_device(db);
        return rc;
    }
    /* the host copies stay for decode and product rows; those need the nucleotide distributions, transitions
     * and names, not the 5.4 KB per node of match emissions that now live in HBM */
    if (!db->keep_host_tables)
        for (auto *p : db->profs)
        {
            free(p->match_emission);
            p->match_emission = nullptr;
        }
    {
        /* what the device has left once the tables are resident: the budget of later passes is taken from this and
         * the pool's own counters (cudaMemGetInfo on the scan path costs up to 20 ms with a multi-GB pool) */
        size_t free_b = 0, total_b = 0;
        CU_TRY(cudaMemGetInfo(&free_b, &total_b));
        db->free_at_commit = free_b;
    }
    db->committed = true;
    return RC_OK;
}

extern "C" unsigned dcpgpu_db_nprofiles(struct dcpgpu_db const *db) { return (unsigned)db->profs.size(); }
extern "C" char const *dcpgpu_db_accession(struct dcpgpu_db const *db, unsigned i)
{
    return i < db->profs.size() ? db->profs[i]->accession : nullptr;
}
extern "C" unsigned dcpgpu_db_core_size(struct dcpgpu_db const *db, unsigned i)
{
    return i < db->profs.size() ? db->profs[i]->core_size : 0;
}
extern "C" uint64_t dcpgpu_db_device_bytes(struct dcpgpu_db const *db) { return db->device_bytes; }

extern "C" void dcpgpu_db_del(struct dcpgpu_db *db)
{
    if (!db) return;
    if (db->owns_profs)
        for (auto *p : db->profs) protein_profile_del(p);
    if (db->device >= 0)
    {
        cudaSetDevice(db->device);
        if (db->stream) cudaStreamSynchronize(db->stream);
        db_release_device(db);
        if (db->h_stage) cudaFreeHost(db->h_stage);
        if (db->stream) cudaStreamDestroy(db->stream);
        for (int i = 0; i < dcpgpu_db::kSide; ++i)
        {
            if (db->side[i]) cudaStreamSynchronize(db->side[i]), cudaStreamDestroy(db->side[i]);
            if (db->join_ev[i]) cudaEventDestroy(db->join_ev[i]);
        }
        if (db->fork_ev) cudaEventDestroy(db->fork_ev);
        if (db->pool) cudaMemPoolDestroy(db->pool);
    }
    delete db;
}

const protein_profile *dcp_db_profile(struct dcpgpu_db const *db, unsigned i) { return db->profs[i]; }

extern "C" enum rc dcpgpu_seqs_new(struct dcpgpu_seqs **out, struct dcpgpu_db *db, unsigned nseqs,
                                   char const *const *seqs, unsigned const *lens)
{
    if (!db->committed) return dcp_error(RC_EFAIL, "commit the database first");
    if (nseqs == 0) return dcp_error(RC_EINVAL, "no sequences");
    CU_TRY(cudaSetDevice(db->device));
    dcpgpu_seqs *sq = new (std::nothrow) dcpgpu_seqs;
    if (!sq) return dcp_error(RC_ENOMEM, "alloc seqs");
    sq->db = db;
    sq->nseq = nseqs;
    sq->metas.resize(nseqs);
    uint64_t total = 0;
    for (unsigned i = 0; i < nseqs; ++i)
    {
        if (lens[i] == 0)
        {
            delete sq;
            return dcp_error(RC_EINVAL, "sequence cannot be empty"); /* protein_profile.c:158 */
        }
        sq->metas[i].len = lens[i], sq->metas[i].pad = 0, sq->metas[i].row_off = total;
        sq->metas[i].rec_off = total + i; /* L+1 records per sequence */
        total += lens[i];
    }
    sq->total = total;
    /* pinned staging buffer, cached in the db and grown on demand */
    if (db->h_stage_cap < total)
    {
        if (db->h_stage) cudaFreeHost(db->h_stage);
        db->h_stage = nullptr, db->h_stage_cap = 0;
        size_t cap = total + total / 4 + 4096;
        if (cudaMallocHost(&db->h_stage, cap) != cudaSuccess)
        {
            delete sq;
            return dcp_error(RC_ENOMEM, "pinned staging for sequences");
        }
        db->h_stage_cap = cap;
    }
    uint8_t *h = (uint8_t *)db->h_stage;
    static const int8_t lut_init = 0;
    (void)lut_init;
    int8_t lut[256];
    memset(lut, -1, sizeof lut);
    lut['A'] = 0, lut['C'] = 1, lut['G'] = 2, lut['T'] = 3;
    bool bad = false;
    for (unsigned i = 0; i < nseqs && !bad; ++i)
    {
        uint8_t *dst = h + sq->metas[i].row_off;
        const unsigned char *src = (const unsigned char *)seqs[i];
        for (unsigned j = 0; j < lens[i]; ++j)
        {
            int8_t c = lut[src[j]];
            if (c < 0)
            {
                bad = true;
                break;
            }
            dst[j] = (uint8_t)c;
        }
    }
    if (bad)
    {
        delete sq;
        return dcp_error(RC_EINVAL, "sequence symbol outside ACGT");
    }
    cudaError_t e = cudaMallocFromPoolAsync(&sq->d_bases, total, db->pool, db->stream);
    if (e == cudaSuccess) e = cudaMallocFromPoolAsync(&sq->d_metas, nseqs * sizeof(SeqMeta), db->pool, db->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(sq->d_bases, h, total, cudaMemcpyHostToDevice, db->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(sq->d_metas, sq->metas.data(), nseqs * sizeof(SeqMeta), cudaMemcpyHostToDevice, db->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(db->stream); /* staging buffer is reusable again */
    if (e != cudaSuccess)
    {
        dcp_set_error(cudaGetErrorString(e));
        dcpgpu_seqs_del(sq);
        return RC_EFAIL;
    }
    sq->h2d_bytes = total + nseqs * sizeof(SeqMeta);
    *out = sq;
    return RC_OK;
}

extern "C" void dcpgpu_seqs_del(struct dcpgpu_seqs *sq)
{
    if (!sq) return;
    cudaSetDevice(sq->db->device);
    cudaFreeAsync(sq->d_bases, sq->db->stream), cudaFreeAsync(sq->d_metas, sq->db->stream);
    delete sq;
}


extern "C" enum rc dcpgpu_scan_resident(struct dcpgpu_db *db, struct dcpgpu_seqs *sq,
                                        struct dcpgpu_params const *prm, struct dcpgpu_result **out)
{
    if (!db->committed || sq->db != db) return dcp_error(RC_EINVAL, "sequences were staged for another database");
    CU_TRY(cudaSetDevice(db->device));
    cudaStream_t st = db->stream;
    const uint32_t nseq = sq->nseq, nprof = (uint32_t)db->profs.size(), n_null = (uint32_t)db->null_tabs.size();
    const size_t npairs = (size_t)nseq * nprof;

    dcpgpu_result *res = new (std::nothrow) dcpgpu_result;
    if (!res) return dcp_error(RC_ENOMEM, "alloc result");
    res->db = db, res->nseq = nseq, res->nprof = nprof, res->n_null = n_null;
    struct Guard
    {
        dcpgpu_result *r;
        ~Guard() { if (r) dcpgpu_result_del(r); }
    } guard{res};

    /* protein_profile_setup per sequence length (host libm, exactly once per distinct L) */
    std::vector<float> spec((size_t)nseq * 16, 0.0f);
    {
        std::vector<std::pair<uint32_t, uint32_t>> byL(nseq);
        for (uint32_t s = 0; s < nseq; ++s) byL[s] = {sq->metas[s].len, s};
        std::sort(byL.begin(), byL.end());
        float x[13];
        uint32_t last = 0;
        for (auto &e : byL)
        {
            if (e.first != last) dcp_specials(e.first, prm->multi_hits, prm->hmmer3_compat, x), last = e.first;
            memcpy(&spec[(size_t)e.second * 16], x, sizeof x);
        }
    }
    DevBuf b_spec, b_rows, b_wcodes, b_counter, b_nhits;
    CU_TRY(b_spec.alloc(spec.size() * sizeof(float), db));
    const uint64_t total_recs = sq->total + nseq;
    CU_TRY(b_rows.alloc((size_t)n_null * total_recs * sizeof(RowRec), db));
    CU_TRY(b_wcodes.alloc(total_recs * sizeof(uint16_t), db));
    CU_TRY(b_counter.alloc(kMaxClasses * sizeof(unsigned long long), db));
    CU_TRY(b_nhits.alloc(2 * sizeof(unsigned long long), db));
    CU_TRY(cudaMallocFromPoolAsync(&res->d_alt, npairs * sizeof(float), db->pool, st));
    CU_TRY(cudaMallocFromPoolAsync(&res->d_null, (size_t)nseq * n_null * sizeof(float), db->pool, st));
    CU_TRY(cudaMallocFromPoolAsync(&res->d_hit, npairs, db->pool, st));

    struct Events
    {
        cudaEvent_t e[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        ~Events()
        {
            for (auto x : e)
                if (x) cudaEventDestroy(x);
        }
    } events;
    cudaEvent_t(&ev)[5] = events.e;
    for (auto &e : ev) CU_TRY(cudaEventCreate(&e));
    uint64_t launches = 0;

    CU_TRY(cudaEventRecord(ev[0], st));
    CU_TRY(cudaMemcpyAsync(b_spec.p, spec.data(), spec.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemsetAsync(b_counter.p, 0, kMaxClasses * sizeof(unsigned long long), st));
    CU_TRY(cudaMemsetAsync(b_nhits.p, 0, 2 * sizeof(unsigned long long), st));
    k_rows<<<nseq, 128, 0, st>>>(sq->d_bases, sq->d_metas, nseq, db->d_null_tabs, db->d_ins_tab, n_null, total_recs,
                                 b_rows.as<RowRec>(), b_wcodes.as<uint16_t>());
    k_null<<<(unsigned)(((size_t)nseq * n_null + 127) / 128), 128, 0, st>>>(sq->d_metas, nseq, n_null, total_recs, b_rows.as<RowRec>(),
                                                        b_spec.as<float>(), res->d_null);
    launches += 2;
    CU_TRY(cudaEventRecord(ev[1], st));

    /* sequences per L2 tile: their row records (64 B per row and null table) stay L2-resident while every
     * profile passes.  32 MB measured best (DRAM bytes per config-2 launch: 64 MB 145 GB, 48 MB 78 GB,
     * 32 MB 34 GB, 16 MB 74 GB; profiles/r01_tile_sweep_dram.csv): data read by all SMs is held in both L2
     * partitions, so about half of the 126 MB is usable for it. */
    uint32_t seq_tile;
    {
        const double rec_bytes_per_seq = (double)total_recs / nseq * sizeof(RowRec) * n_null;
        double tile_mb = 32.0;
        if (const char *e = getenv("DCPGPU_TILE_MB")) tile_mb = std::max(0.25, atof(e)); /* experiment knob */
        double t = (tile_mb * 1024 * 1024) / rec_bytes_per_seq;
        seq_tile = (uint32_t)std::min<double>(std::max<double>(t, kSeqChunk), 1 << 20);
        seq_tile = std::max<uint32_t>(kSeqChunk, seq_tile / kSeqChunk * kSeqChunk);
    }
    uint64_t cells = 0;
    ScoreArgs sa = {db->d_emis, db->d_trans, db->d_metas, nullptr, 0, sq->d_metas, nseq, total_recs,
                    b_rows.as<RowRec>(), b_wcodes.as<uint16_t>(), b_spec.as<float>(), res->d_alt, nprof, nullptr, seq_tile};
    /* one persistent launch per kernel class, spread over the main and the side streams: every launch sizes its grid
     * to fill the GPU, so they still run one after the other, but a class's blocks start as the previous class's
     * blocks run out of work (the tail of a launch is up to one work item long) */
    StreamFan fan(db);
    static const bool fan_out = !(getenv("DCPGPU_ONE_STREAM") && atoi(getenv("DCPGPU_ONE_STREAM")) != 0);
    if (fan_out) CU_TRY(fan.fork());
    for (int q = 0; q < kMaxClasses; ++q)
    {
        if (db->class_list[q].empty()) continue;
        const dcp_class &kc = *dcp_class_at(q);
        sa.class_profs = db->d_class[q], sa.n_class = (uint32_t)db->class_list[q].size();
        sa.counter = b_counter.as<unsigned long long>() + q;
        cudaStream_t cs = fan_out ? fan.next() : st;
        CU_TRY(kc.tw <= 1 ? dcp_launch_score(kc, db->sm_count, cs, sa) : dcp_launch_score_mw(kc, db->sm_count, cs, sa));
        launches++;
        for (uint32_t id : db->class_list[q]) cells += (uint64_t)db->metas[id].M * sq->total;
    }
    if (fan_out) CU_TRY(fan.join());
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaEventRecord(ev[2], st));

    k_lrt<<<(unsigned)((npairs + 255) / 256), 256, 0, st>>>(res->d_alt, res->d_null, db->d_metas, nseq, nprof, n_null,
                                                            prm->lrt_threshold, res->d_hit,
                                                            b_nhits.as<unsigned long long>());
    launches++;
    unsigned long long nhits = 0;
    CU_TRY(cudaMemcpyAsync(&nhits, b_nhits.p, sizeof nhits, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    uint64_t d2h = sizeof nhits;

    if (nhits)
    {
        DevBuf b_list;
        CU_TRY(b_list.alloc(nhits * sizeof(unsigned long long), db));
        k_collect<<<(unsigned)((npairs + 255) / 256), 256, 0, st>>>(res->d_hit, npairs,
                                                                    b_nhits.as<unsigned long long>() + 1,
                                                                    b_list.as<unsigned long long>());
        launches++;
        std::vector<unsigned long long> list(nhits);
        CU_TRY(cudaMemcpyAsync(list.data(), b_list.p, nhits * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        d2h += nhits * sizeof(unsigned long long);
        std::sort(list.begin(), list.end()); /* (sequence, profile) order, independent of scheduling */
        DevBuf b_ga, b_gn;
        CU_TRY(b_ga.alloc(nhits * sizeof(float), db));
        CU_TRY(b_gn.alloc(nhits * sizeof(float), db));
        CU_TRY(cudaMemcpyAsync(b_list.p, list.data(), nhits * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
        k_gather<<<(unsigned)((nhits + 255) / 256), 256, 0, st>>>(b_list.as<unsigned long long>(), nhits, res->d_alt,
                                                                  res->d_null, db->d_metas, nprof, n_null,
                                                                  b_ga.as<float>(), b_gn.as<float>());
        launches++;
        res->hit_alt.resize(nhits), res->hit_null.resize(nhits);
        CU_TRY(cudaMemcpyAsync(res->hit_alt.data(), b_ga.p, nhits * sizeof(float), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(res->hit_null.data(), b_gn.p, nhits * sizeof(float), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        d2h += nhits * 2 * sizeof(float);
        res->hits.resize(nhits);
        for (size_t i = 0; i < nhits; ++i)
        {
            res->hits[i].seq = (uint32_t)(list[i] / nprof);
            res->hits[i].prof = (uint32_t)(list[i] % nprof);
            res->hits[i].step_off = 0, res->hits[i].nsteps = 0;
        }
    }
    CU_TRY(cudaEventRecord(ev[3], st));
    if (nhits && prm->want_paths)
    {
        enum rc rc = dcp_trace_hits(db, sq, res, b_rows.as<RowRec>(), b_wcodes.as<uint16_t>(), b_spec.as<float>(),
                                    &launches);
        if (rc) return rc;
        res->have_paths = true;
    }
    CU_TRY(cudaEventRecord(ev[4], st));
    CU_TRY(cudaStreamSynchronize(st));
    CU_TRY(cudaGetLastError());

    dcpgpu_timing &t = res->timing;
    cudaEventElapsedTime(&t.prep_ms, ev[0], ev[1]);
    cudaEventElapsedTime(&t.score_ms, ev[1], ev[2]);
    cudaEventElapsedTime(&t.trace_ms, ev[3], ev[4]);
    cudaEventElapsedTime(&t.total_ms, ev[0], ev[4]);
    t.launches = launches;
    t.alt_cells = cells;
    t.h2d_bytes = spec.size() * sizeof(float);
    t.d2h_bytes = d2h + res->steps.size() * sizeof(dcp_step) + res->hits.size() * 8;
    guard.r = nullptr;
    *out = res;
    if (prm->progress) prm->progress(prm->user, (uint64_t)npairs); /* progress_consume, scan_thread.c:120 */
    return RC_OK;
}

/* one launch set over one batch from host buffers: stage the sequences, scan, release them (dcpgpu_scan tiles
 * large batches over this, dcp_multi.cpp) */
enum rc dcp_scan_once(struct dcpgpu_db *db, unsigned nseqs, char const *const *seqs, unsigned const *lens,
                      struct dcpgpu_params const *prm, struct dcpgpu_result **out)
{
    dcpgpu_seqs *sq = nullptr;
    enum rc rc = dcpgpu_seqs_new(&sq, db, nseqs, seqs, lens);
    if (rc) return rc;
    rc = dcpgpu_scan_resident(db, sq, prm, out);
    if (!rc) (*out)->timing.h2d_bytes += sq->h2d_bytes;
    dcpgpu_seqs_del(sq);
    return rc;
}

enum rc dcp_result_fetch(dcpgpu_result *r)
{
    if (r->fetched) return RC_OK;
    size_t n = (size_t)r->nseq * r->nprof;
    if (!r->parts.empty())
    {
        /* merged result: scatter every part's pair arrays into the global (sequence, profile) matrix */
        r->alt.assign(n, NEG_INF), r->null_ll.assign(n, NEG_INF), r->hit.assign(n, 0);
        for (size_t k = 0; k < r->parts.size(); ++k)
        {
            dcpgpu_result *q = r->parts[k];
            enum rc rc = dcp_result_fetch(q);
            if (rc) return rc;
            const std::vector<uint32_t> *gp = r->part_profs[k];
            for (uint32_t s = 0; s < q->nseq; ++s)
                for (uint32_t p = 0; p < q->nprof; ++p)
                {
                    const size_t src = (size_t)s * q->nprof + p;
                    const size_t dst = (size_t)(r->part_seq0[k] + s) * r->nprof + (gp ? (*gp)[p] : p);
                    r->alt[dst] = q->alt[src], r->null_ll[dst] = q->null_ll[src], r->hit[dst] = q->hit[src];
                }
            /* the part's own copies are no longer needed */
            std::vector<float>().swap(q->alt), std::vector<float>().swap(q->null_ll), std::vector<uint8_t>().swap(q->hit);
            q->fetched = false;
        }
        r->fetched = true;
        return RC_OK;
    }
    CU_TRY(cudaSetDevice(r->db->device));
    r->alt.resize(n), r->null_ll.resize(n), r->hit.resize(n);
    std::vector<float> nt((size_t)r->nseq * r->n_null);
    CU_TRY(cudaMemcpy(r->alt.data(), r->d_alt, n * sizeof(float), cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(nt.data(), r->d_null, nt.size() * sizeof(float), cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(r->hit.data(), r->d_hit, n, cudaMemcpyDeviceToHost));
    for (uint32_t s = 0; s < r->nseq; ++s)
        for (uint32_t p = 0; p < r->nprof; ++p)
            r->null_ll[(size_t)s * r->nprof + p] = nt[(size_t)s * r->n_null + r->db->null_id[p]];
    r->fetched = true;
    return RC_OK;
}
static enum rc result_fetch(dcpgpu_result *r) { return dcp_result_fetch(r); }

extern "C" unsigned dcpgpu_result_nseqs(struct dcpgpu_result const *r) { return r->nseq; }
extern "C" unsigned dcpgpu_result_nprofiles(struct dcpgpu_result const *r) { return r->nprof; }
extern "C" float const *dcpgpu_result_null_loglik(struct dcpgpu_result const *r)
{
    return result_fetch(const_cast<dcpgpu_result *>(r)) ? nullptr : r->null_ll.data();
}
extern "C" float const *dcpgpu_result_alt_loglik(struct dcpgpu_result const *r)
{
    return result_fetch(const_cast<dcpgpu_result *>(r)) ? nullptr : r->alt.data();
}
extern "C" uint8_t const *dcpgpu_result_hit(struct dcpgpu_result const *r)
{
    return result_fetch(const_cast<dcpgpu_result *>(r)) ? nullptr : r->hit.data();
}
extern "C" uint64_t dcpgpu_result_nhits(struct dcpgpu_result const *r) { return r->hits.size(); }

extern "C" enum rc dcpgpu_result_hit_at(struct dcpgpu_result const *r, uint64_t i, unsigned *seq_idx,
                                        unsigned *prof_idx, struct dcp_step const **steps, unsigned *nsteps)
{
    if (i >= r->hits.size()) return dcp_error(RC_EINVAL, "hit index out of range");
    const HitRec &h = r->hits[i];
    if (seq_idx) *seq_idx = h.seq;
    if (prof_idx) *prof_idx = h.prof;
    if (steps) *steps = r->have_paths ? r->steps.data() + h.step_off : nullptr;
    if (nsteps) *nsteps = r->have_paths ? h.nsteps : 0;
    return RC_OK;
}

extern "C" enum rc dcpgpu_result_hits(struct dcpgpu_result const *r, unsigned *seq_idx, unsigned *prof_idx, float *alt,
                                      float *null_ll, unsigned *nsteps)
{
    for (size_t i = 0; i < r->hits.size(); ++i)
    {
        if (seq_idx) seq_idx[i] = r->hits[i].seq;
        if (prof_idx) prof_idx[i] = r->hits[i].prof;
        if (alt) alt[i] = r->hit_alt[i];
        if (null_ll) null_ll[i] = r->hit_null[i];
        if (nsteps) nsteps[i] = r->have_paths ? r->hits[i].nsteps : 0;
    }
    return RC_OK;
}

extern "C" uint64_t dcpgpu_result_steps(struct dcpgpu_result const *r, struct dcp_step const **steps)
{
    if (steps) *steps = r->have_paths ? r->steps.data() : nullptr;
    return r->have_paths ? r->steps.size() : 0;
}

extern "C" void dcpgpu_result_timing(struct dcpgpu_result const *r, struct dcpgpu_timing *t) { *t = r->timing; }

extern "C" void dcpgpu_result_del(struct dcpgpu_result *r)
{
    if (!r) return;
    for (auto *q : r->parts) dcpgpu_result_del(q);
    if (r->db && r->db->device >= 0 && (r->d_alt || r->d_null || r->d_hit))
    {
        cudaSetDevice(r->db->device);
        cudaStream_t st = r->db->stream;
        cudaFreeAsync(r->d_alt, st), cudaFreeAsync(r->d_null, st), cudaFreeAsync(r->d_hit, st);
    }
    delete r;
}

extern "C" unsigned dcpgpu_result_nparts(struct dcpgpu_result const *r) { return (unsigned)r->parts.size(); }

extern "C" enum rc dcpgpu_result_part_timing(struct dcpgpu_result const *r, unsigned part, int *device,
                                             struct dcpgpu_timing *t)
{
    if (part >= r->parts.size()) return dcp_error(RC_EINVAL, "no such part");
    if (device) *device = r->part_device[part];
    if (t) *t = r->parts[part]->timing;
    return RC_OK;
}
