/* dcp_engine.h -- types shared by the CUDA translation units of libdcpgpu (C++ only). */
#ifndef DCP_ENGINE_H
#define DCP_ENGINE_H

#include "dcp_internal.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <vector>

#define NEG_INF (-INFINITY)
#define FULL 0xffffffffu

constexpr int kTab = DCP_FRAME_TABLE_SIZE;
constexpr int kMaxQ = DCP_MAX_Q;          /* nodes per lane, single-warp classes cover M <= 256 */
constexpr int kMaxW = DCP_MAX_W;          /* warps per block in the multi-warp classes: M <= 8 * 256 = 2048 per block */
constexpr int kMaxGroupWarps = DCP_MAX_GROUP_WARPS; /* two blocks (a cluster) per pair above 2048 nodes: M <= 4096 */
constexpr int kMaxClasses = DCP_MAX_CLASSES; /* rows of the kernel class table (dcp_classes.h) */
constexpr int kWarpsPerBlock = 8; /* k_score block = 8 independent warps */
/* independent warps per SM in k_score<Q>: 8 nodes per lane need 254 registers (two warps per scheduler);
 * up to 4 nodes per lane the state fits 168 registers (three per scheduler), up to 2 it fits 128 (four).
 * Measured: M = 128: 550 vs 463 GCUPS, M = 64: 345 vs 269 with 12 instead of 8 warps. */
constexpr int score_warps(int Q) { return Q <= 2 ? 16 : Q <= 4 ? 12 : kWarpsPerBlock; } /* Q = 5 at 12 warps: no gain */
#ifndef DCP_SEQ_CHUNK
#define DCP_SEQ_CHUNK 4
#endif
constexpr int kSeqChunk = DCP_SEQ_CHUNK; /* sequences per work item */

#define CU_TRY(expr)                                                                           \
    do                                                                                         \
    {                                                                                          \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess)                                                                 \
        {                                                                                      \
            dcp_set_error(cudaGetErrorString(e_));                                             \
            return RC_EFAIL;                                                                   \
        }                                                                                      \
    } while (0)

/*
 * Per-row inputs of the DP besides the match emissions.  A sequence of L nucleotides owns L+1
 * records (rows 0..L).  wcode[row] (one uint16 per record, independent of the null table) is the
 * 2-bit packed window of the last five nucleotides ending at that row; the frame-table code of
 * seq[j-l:j] is frame_off[l] + (w & (4^l - 1)).  RowRec holds the emissions that are shared by
 * all core nodes: insert (eI) and N/J/C/R (eN), one record per (null table, row).
 */
struct __align__(16) RowRec
{
    float eI[5]; /* insert emission of seq[j-l:j], l = 1..5 */
    float pad0[3];
    float eN[5]; /* N/J/C/R emission of seq[j-l:j] */
    float pad1[3];
};
static_assert(sizeof(RowRec) == 64, "row record is one 64-byte line");

struct ProfMeta
{
    uint32_t M, Q, QP, null_id;
    uint32_t W, cls; /* warps per pair; kernel class (row of the class table) */
    uint32_t LN, pad; /* lanes per pair: 32, or 16 when two pairs share a warp (W = 1) */
    uint64_t emis_off;  /* floats into d_emis */
    uint64_t trans_off; /* floats into d_trans */
};

struct SeqMeta
{
    uint32_t len;
    uint32_t pad;
    uint64_t row_off; /* first base */
    uint64_t rec_off; /* record of row 0 (= row_off + sequence index) */
};


struct dcpgpu_db
{
    int device = 0; /* -1: host-only view (the global profile list of a multi-device database) */
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    /* side streams: the per-class launches of a scan are spread over them so that one class's tail (persistent
     * blocks running out of work, or a handful of hits to trace) overlaps the next class's start */
    static constexpr int kSide = 7;
    cudaStream_t side[kSide] = {};
    cudaEvent_t fork_ev = nullptr, join_ev[kSide] = {};
    cudaMemPool_t pool = nullptr; /* the engine's own stream-ordered pool (scratch of scans, results) */
    bool committed = false;
    bool owns_profs = true;        /* false: shard of a dcpgpu_mdb, profiles belong to its view */
    bool keep_host_tables = false; /* shard of a replicated database: the owner frees the host tables */
    float epsilon = -1.0f;
    std::vector<protein_profile *> profs; /* deep copies, host side (decode, products) */
    std::vector<std::vector<float>> null_tabs;
    std::vector<uint32_t> null_id;
    std::vector<ProfMeta> metas;
    std::vector<uint32_t> class_list[kMaxClasses]; /* profile ids by kernel class */
    float *d_emis = nullptr, *d_trans = nullptr, *d_null_tabs = nullptr, *d_ins_tab = nullptr;
    ProfMeta *d_metas = nullptr;
    uint32_t *d_class[kMaxClasses] = {nullptr};
    uint64_t device_bytes = 0;
    size_t free_at_commit = 0; /* free device memory right after commit (pool empty) */
    void *h_stage = nullptr; /* pinned staging for sequence uploads (grow-only) */
    size_t h_stage_cap = 0;
};

struct dcpgpu_seqs
{
    dcpgpu_db *db = nullptr;
    uint32_t nseq = 0;
    uint64_t total = 0; /* nucleotides == row records per null table */
    std::vector<SeqMeta> metas;
    uint8_t *d_bases = nullptr;
    SeqMeta *d_metas = nullptr;
    uint64_t h2d_bytes = 0;
};

struct HitRec
{
    uint32_t seq, prof;
    uint64_t step_off;
    uint32_t nsteps;
};

struct dcpgpu_result
{
    dcpgpu_db *db = nullptr;
    /* merged result of a multi-device scan: the per-device results it was built from (owned) and, for each,
     * the global profile index of its local profiles / the first global sequence index */
    std::vector<dcpgpu_result *> parts;
    std::vector<const std::vector<uint32_t> *> part_profs;
    std::vector<uint32_t> part_seq0;
    std::vector<int> part_device;
    uint32_t nseq = 0, nprof = 0, n_null = 0;
    float *d_alt = nullptr, *d_null = nullptr;
    uint8_t *d_hit = nullptr;
    bool fetched = false;
    std::vector<float> alt, null_ll;
    std::vector<uint8_t> hit;
    std::vector<HitRec> hits;
    std::vector<float> hit_alt, hit_null;
    std::vector<dcp_step> steps;
    bool have_paths = false;
    dcpgpu_timing timing = {};
};


/* Scratch device buffer of one scan: stream-ordered allocation from the database's own memory pool
 * (created at dcpgpu_db_new with its release threshold raised, so repeated scans reuse the same blocks
 * and never pay cudaMalloc/cudaFree; the device's default pool, which torch and others share, is left alone). */
struct DevBuf
{
    void *p = nullptr;
    cudaStream_t st = nullptr;
    cudaError_t alloc(size_t bytes, dcpgpu_db *db)
    {
        st = db->stream;
        return cudaMallocFromPoolAsync(&p, bytes ? bytes : 1, db->pool, st);
    }
    ~DevBuf()
    {
        if (p) cudaFreeAsync(p, st);
    }
    template <class T>
    T *as() { return (T *)p; }
};

/* arguments of a score-pass launch for one kernel class (dcp_score_sw.cu, dcp_score_mw.cu) */
struct ScoreArgs
{
    const float *emis, *trans;
    const ProfMeta *metas;
    const uint32_t *class_profs; /* profile ids of the class */
    uint32_t n_class;
    const SeqMeta *seqs;
    uint32_t nseq;
    uint64_t total_recs;
    const RowRec *rows;
    const uint16_t *wcodes;
    const float *spec;
    float *alt;
    uint32_t nprof;
    unsigned long long *counter; /* work queue cursor of this launch (zeroed) */
    uint32_t seq_tile;           /* sequences per L2 tile */
};
cudaError_t dcp_launch_score(const dcp_class &c, int sm_count, cudaStream_t st, const ScoreArgs &a);    /* tw = 1, 0 */
cudaError_t dcp_launch_score_mw(const dcp_class &c, int sm_count, cudaStream_t st, const ScoreArgs &a); /* tw > 1 */

/* Launches of a phase spread over the database's main stream and its side streams: fork() makes the side streams
 * wait for what the main stream has queued so far, next() hands out the streams round-robin, join() makes the main
 * stream wait for everything queued on the side streams. */
struct StreamFan
{
    dcpgpu_db *db;
    int n = 0;
    bool used[dcpgpu_db::kSide] = {};
    explicit StreamFan(dcpgpu_db *d) : db(d) {}
    cudaError_t fork()
    {
        cudaError_t e = cudaEventRecord(db->fork_ev, db->stream);
        for (int i = 0; i < dcpgpu_db::kSide && e == cudaSuccess; ++i) e = cudaStreamWaitEvent(db->side[i], db->fork_ev, 0);
        return e;
    }
    cudaStream_t next()
    {
        const int k = n++ % (dcpgpu_db::kSide + 1);
        if (k == 0) return db->stream;
        used[k - 1] = true;
        return db->side[k - 1];
    }
    cudaError_t join()
    {
        cudaError_t e = cudaSuccess;
        for (int i = 0; i < dcpgpu_db::kSide && e == cudaSuccess; ++i)
            if (used[i])
            {
                e = cudaEventRecord(db->join_ev[i], db->side[i]);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(db->stream, db->join_ev[i], 0);
            }
        return e;
    }
};

/* dcp_trace.cu */
enum rc dcp_trace_hits(dcpgpu_db *db, dcpgpu_seqs *sq, dcpgpu_result *res, const RowRec *d_rows,
                       const uint16_t *d_wcodes, const float *d_spec, uint64_t *launches);
const protein_profile *dcp_db_profile(struct dcpgpu_db const *db, unsigned i);
enum rc dcp_db_adopt(struct dcpgpu_db *db, struct protein_profile *prof); /* takes ownership, no copy */
enum rc dcp_db_borrow(struct dcpgpu_db *db, struct protein_profile *prof); /* no copy, no ownership (mdb shards) */
struct dcpgpu_db *dcp_db_new_host(void); /* device = -1: profile list only */
enum rc dcp_result_fetch(dcpgpu_result *r);

#endif
