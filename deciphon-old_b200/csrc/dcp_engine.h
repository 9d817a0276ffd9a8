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
constexpr int kMaxQ = 8;          /* nodes per lane, single-warp classes cover M <= 256 */
constexpr int kMaxW = 8;          /* warps per block in the multi-warp classes: M <= 8 * 256 = 2048 per block */
constexpr int kMaxGroupWarps = 16; /* two blocks (a cluster) per pair above 2048 nodes: M <= 4096 */
constexpr int kClsW2Q6 = kMaxQ + kMaxGroupWarps + 1; /* 257..384 nodes: two warps, 6 nodes per lane */
constexpr int kClsW2Q7 = kMaxQ + kMaxGroupWarps + 2; /* 385..448 nodes: two warps, 7 nodes per lane */
constexpr int kClsW3Q6 = kMaxQ + kMaxGroupWarps + 3; /* 513..576 nodes: three warps, 6 nodes per lane */
constexpr int kClsW3Q7 = kMaxQ + kMaxGroupWarps + 4; /* 577..672 nodes: three warps, 7 nodes per lane */
constexpr int kNumClasses = kClsW3Q7; /* class c: 1..8 = one warp, Q = c; 8 + TW = TW warps per pair, Q = 8
                                       * (TW = 2..8 one block; 10, 12, 14, 16 two blocks); then the two above */
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
    uint32_t W, cls; /* warps per pair; kernel class */
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
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool committed = false;
    float epsilon = -1.0f;
    std::vector<protein_profile *> profs; /* deep copies, host side (decode, products) */
    std::vector<std::vector<float>> null_tabs;
    std::vector<uint32_t> null_id;
    std::vector<ProfMeta> metas;
    std::vector<uint32_t> class_list[kNumClasses + 1]; /* profile ids by kernel class */
    float *d_emis = nullptr, *d_trans = nullptr, *d_null_tabs = nullptr, *d_ins_tab = nullptr;
    ProfMeta *d_metas = nullptr;
    uint32_t *d_class[kNumClasses + 1] = {nullptr};
    uint64_t device_bytes = 0;
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


/* Scratch device buffer of one scan: stream-ordered allocation from the device's default memory
 * pool (release threshold raised at dcpgpu_db_new, so repeated scans reuse the same blocks and
 * never pay cudaMalloc/cudaFree). */
struct DevBuf
{
    void *p = nullptr;
    cudaStream_t st = nullptr;
    cudaError_t alloc(size_t bytes, cudaStream_t stream)
    {
        st = stream;
        return cudaMallocAsync(&p, bytes ? bytes : 1, stream);
    }
    ~DevBuf()
    {
        if (p) cudaFreeAsync(p, st);
    }
    template <class T>
    T *as() { return (T *)p; }
};

/* dcp_trace.cu */
enum rc dcp_trace_hits(dcpgpu_db *db, dcpgpu_seqs *sq, dcpgpu_result *res, const RowRec *d_rows,
                       const uint16_t *d_wcodes, const float *d_spec, uint64_t *launches);
const protein_profile *dcp_db_profile(struct dcpgpu_db const *db, unsigned i);
enum rc dcp_db_adopt(struct dcpgpu_db *db, struct protein_profile *prof); /* takes ownership, no copy */

#endif
