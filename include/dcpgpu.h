/*
 * dcpgpu.h -- C ABI of the B200-native deciphon scan engine (libdcpgpu.so).
 *
 * Drop-in boundary for ONE path of EBI-Metagenomics/deciphon-old: the inner loop of
 * thread_run (src/server/scan_thread.c:86-135): per (sequence, profile) pair
 * protein_profile_setup -> imm_dp_viterbi(null) -> imm_dp_viterbi(alt) -> xmath_lrt ->
 * threshold -> path/product.  Everything is plain C: opaque handles, caller-owned
 * input buffers, int return codes equal to the reference's enum rc
 * (include/deciphon/core/rc.h:4-15).  No torch types, no global state.
 *
 * Part 1 mirrors the host-side model API the path is fed by (src/model), with the
 * reference's own function names and argument meaning; imm types are replaced by
 * plain arrays because imm is not part of this library.
 * Part 2 is the batch engine that replaces thread_run's loop and the imm DP.
 * Part 3 is the product writer (src/server/prod.c, src/server/protein_match.c).
 *
 * Paths cited below are relative to the reference repository root.
 */
#ifndef DCPGPU_H
#define DCPGPU_H

#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* include/deciphon/core/rc.h:4-15 -- same values, same meaning */
enum rc
{
    RC_OK,
    RC_END,
    RC_EFAIL,
    RC_EINVAL,
    RC_EIO,
    RC_ENOMEM,
    RC_EPARSE,
    RC_EAPI,
    RC_EHTTP,
};

/* include/deciphon/core/limits.h:4-13 */
enum
{
    DCP_PROFILE_ACC_SIZE = 32,
    DCP_PROTEIN_MODEL_CORE_SIZE_MAX = 4096,
    DCP_AMINO_SIZE = 20,  /* IMM_AMINO_SIZE, order ACDEFGHIKLMNPQRSTVWY */
    DCP_NUCLT_SIZE = 4,   /* IMM_NUCLT_SIZE, order ACGT */
    DCP_STATE_NAME_SIZE = 8,
    DCP_FRAME_TABLE_SIZE = 1364, /* 4+16+64+256+1024 strings of 1..5 nt per frame state */
};

/* include/deciphon/model/entry_dist.h:4-9 */
enum entry_dist
{
    ENTRY_DIST_NULL,
    ENTRY_DIST_UNIFORM,
    ENTRY_DIST_OCCUPANCY,
};

/* include/deciphon/model/protein_cfg.h:7-23 (imm_float = float, the reference default) */
struct protein_cfg
{
    enum entry_dist entry_dist;
    float epsilon;
};

/* include/deciphon/model/protein_trans.h:8-27 */
#define PROTEIN_TRANS_SIZE 7
struct protein_trans
{
    union
    {
        struct
        {
            float MM, MI, MD, IM, II, DM, DD;
        };
        float data[PROTEIN_TRANS_SIZE];
    };
};

/* include/deciphon/model/protein_state.h:7-55 -- 16-bit state ids */
enum protein_state_id
{
    PROTEIN_MATCH_STATE = (0 << 14),
    PROTEIN_INSERT_STATE = (1 << 14),
    PROTEIN_DELETE_STATE = (2 << 14),
    PROTEIN_EXT_STATE = (3 << 14),
    PROTEIN_R_STATE = (PROTEIN_EXT_STATE | 0),
    PROTEIN_S_STATE = (PROTEIN_EXT_STATE | 1),
    PROTEIN_N_STATE = (PROTEIN_EXT_STATE | 2),
    PROTEIN_B_STATE = (PROTEIN_EXT_STATE | 3),
    PROTEIN_E_STATE = (PROTEIN_EXT_STATE | 4),
    PROTEIN_J_STATE = (PROTEIN_EXT_STATE | 5),
    PROTEIN_C_STATE = (PROTEIN_EXT_STATE | 6),
    PROTEIN_T_STATE = (PROTEIN_EXT_STATE | 7),
};

/* imm_step (imm/path.h): one element of a decoded state path */
struct dcp_step
{
    uint16_t state_id;
    uint8_t seqlen;
};

/* ------------------------------------------------------------------------- */
/* Part 1 -- model (host, C).  src/model/protein_model.c, protein_profile.c  */
/* ------------------------------------------------------------------------- */
struct protein_model;   /* opaque; include/deciphon/model/protein_model.h:16-48 */
struct protein_profile; /* opaque; include/deciphon/model/protein_profile.h:12-43 */

/* protein_model_init (protein_model.c:105-137).  Returns NULL on ENOMEM. */
struct protein_model *protein_model_new(struct protein_cfg cfg,
                                        float const null_lprobs[DCP_AMINO_SIZE]);
/* protein_model_setup (protein_model.c:154-185): RC_EINVAL for 0 or > 4096 */
enum rc protein_model_setup(struct protein_model *, unsigned core_size);
/* protein_model_add_node (protein_model.c:49-82) */
enum rc protein_model_add_node(struct protein_model *, float const lprobs[DCP_AMINO_SIZE],
                               char consensus);
/* protein_model_add_trans (protein_model.c:84-96) */
enum rc protein_model_add_trans(struct protein_model *, struct protein_trans trans);
void protein_model_del(struct protein_model *);

/* protein_profile_init (protein_profile.c:136-153) */
struct protein_profile *protein_profile_new(char const *accession, struct protein_cfg cfg);
/* protein_profile_absorb (protein_profile.c:218-257): compiles the model into DP tables */
enum rc protein_profile_absorb(struct protein_profile *, struct protein_model const *);
/* protein_h3reader_next's loop body (src/model/protein_h3reader.c:18-72) over in-memory arrays:
 * trans[0], then (match_lprobs[k], trans[k+1]) per node, then absorb.  match_lprobs: [core_size][20],
 * trans: [core_size+1][7], consensus: core_size chars or NULL. */
enum rc protein_profile_build(struct protein_profile *, unsigned core_size,
                              float const null_lprobs[DCP_AMINO_SIZE], float const *match_lprobs,
                              float const *trans, char const *consensus);
/* protein_profile_sample (protein_profile.c:259-304): imm_rnd(seed) driven random profile */
enum rc protein_profile_sample(struct protein_profile *, unsigned seed, unsigned core_size);
/* protein_profile_setup (protein_profile.c:155-216): length-dependent special transitions.
 * RC_EINVAL when seq_size == 0.  out13 (may be NULL) receives
 * NN CC JJ NB CT JB RR EJ EC (E->T) (E->C) (E->B) (E->J). */
enum rc protein_profile_setup(struct protein_profile *, unsigned seq_size, bool multi_hits,
                              bool hmmer3_compat, float out13[13]);
/* protein_profile_decode (protein_profile.c:306-331): most likely codon of a 1..5-nt fragment
 * emitted by a non-mute state; frag is ASCII ACGT.  amino (may be NULL) = imm_gc_decode(1, codon). */
enum rc protein_profile_decode(struct protein_profile const *, char const *frag, unsigned frag_size,
                               unsigned state_id, char codon[3], char *amino);
void protein_profile_del(struct protein_profile *);

unsigned protein_profile_core_size(struct protein_profile const *);
char const *protein_profile_accession(struct protein_profile const *);
/* DP-level numbers (for parity checks against an external oracle).  Pointers stay valid
 * until the profile is deleted.  match_emission: [core_size][1364]; trans: [core_size+1][7]
 * in MM MI MD IM II DM DD order; entry: [core_size] (B->M_k). */
float const *protein_profile_match_emission(struct protein_profile const *);
float const *protein_profile_insert_emission(struct protein_profile const *);
float const *protein_profile_null_emission(struct protein_profile const *);
float const *protein_profile_trans(struct protein_profile const *);
float const *protein_profile_entry(struct protein_profile const *);
/* nuclt_dist (include/deciphon/model/nuclt_dist.h:7-11) as 4 nucltp + 125 codonm doubles.
 * which: -2 null, -1 insert, k >= 0 match node k. */
enum rc protein_profile_nuclt_dist(struct protein_profile const *, int which, double out[129]);

/* HMMER3 ASCII reader: protein_h3reader_init/next/del (src/model/protein_h3reader.c:18-103).
 * _next parses the next profile of the file into the reader's protein_model: RC_OK, RC_END at end of
 * file, RC_EPARSE on malformed input.  The background is HMMER3's Swiss-Prot 50.8 frequencies. */
struct protein_h3reader;
struct protein_h3reader *protein_h3reader_new(struct protein_cfg cfg, FILE *fp);
enum rc protein_h3reader_next(struct protein_h3reader *);
struct protein_model const *protein_h3reader_model(struct protein_h3reader const *);
char const *protein_h3reader_accession(struct protein_h3reader const *); /* ACC, else NAME */
char const *protein_h3reader_name(struct protein_h3reader const *);
void protein_h3reader_del(struct protein_h3reader *);

/* .dcp database container (MessagePack), src/db/writer.c, protein_writer.c, reader.c, protein_reader.c,
 * profile_reader.c and protein_profile_pack/unpack (protein_profile.c:38-117,338-400).  Same root/header/
 * profile maps and key order as the reference; the imm-defined blobs ("abc", "amino", "null", "alt") are
 * stored as explicit arrays instead (dcp_db.c header comment). */
struct protein_db_writer;
struct protein_db_reader;
struct protein_db_writer *protein_db_writer_open(FILE *fp, struct protein_cfg cfg);
enum rc protein_db_writer_pack_profile(struct protein_db_writer *, struct protein_profile const *);
enum rc protein_db_writer_close(struct protein_db_writer *, bool successfully); /* frees the writer */
enum rc protein_db_reader_open(struct protein_db_reader **out, FILE *fp);
unsigned protein_db_reader_nprofiles(struct protein_db_reader const *);
struct protein_cfg protein_db_reader_cfg(struct protein_db_reader const *);
uint32_t protein_db_reader_profile_size(struct protein_db_reader const *, unsigned i);
/* next profile of the file (caller frees with protein_profile_del); RC_END after the last */
enum rc protein_db_reader_next(struct protein_db_reader *, struct protein_profile **out);
void protein_db_reader_close(struct protein_db_reader *);

/* protein_state_name (src/model/protein_state.c:5-39); returns the name length */
unsigned protein_state_name(unsigned id, char name[DCP_STATE_NAME_SIZE]);
bool protein_state_is_mute(unsigned id);

/* xmath_lrt (include/deciphon/core/xmath.h:32-43) */
float xmath_lrt_f32(float null_loglik, float alt_loglik);

/* ------------------------------------------------------------------------- */
/* Part 2 -- engine.  Replaces thread_run's per-pair loop + imm_dp_viterbi.   */
/* ------------------------------------------------------------------------- */
struct dcpgpu_db;     /* profiles resident in HBM on one device */
struct dcpgpu_seqs;   /* a batch of sequences resident in HBM */
struct dcpgpu_result; /* scores, hit flags and hit paths of one scan */

/* struct scan_thread's knobs (src/server/scan_thread.h:16-18); threshold is 10.0 in scan.c:221 */
struct dcpgpu_params
{
    bool multi_hits;
    bool hmmer3_compat;
    double lrt_threshold;
    bool want_paths; /* false: scores + hit flags only (no traceback pass) */
    /* progress_consume (src/server/scan_thread.c:120, src/core/progress.c:70-90): called with the number of
     * (sequence, profile) pairs just finished -- once per launch set, i.e. per memory-sized tile of the batch
     * and per device; calls are serialised.  May be NULL. */
    void (*progress)(void *user, uint64_t pairs_consumed);
    void *user;
};

/* Create an empty database bound to CUDA device `device`.  RC_EFAIL if there is no usable
 * CUDA device (there is no CPU fallback). */
enum rc dcpgpu_db_new(struct dcpgpu_db **db, int device);
/* Append one profile (copied; the caller keeps ownership of `prof`).  All profiles of a db
 * must share epsilon (protein_db header, src/db/protein_writer.c:56-96). */
enum rc dcpgpu_db_add(struct dcpgpu_db *, struct protein_profile const *prof);
/* Lay the tables out for the kernels and upload them to HBM.  Call once, after the adds. */
enum rc dcpgpu_db_commit(struct dcpgpu_db *);
unsigned dcpgpu_db_nprofiles(struct dcpgpu_db const *);
uint64_t dcpgpu_db_device_bytes(struct dcpgpu_db const *);
void dcpgpu_db_del(struct dcpgpu_db *);

/* hmm_press without the REST plumbing (src/server/hmm.c:120-178): every profile of a HMMER3 file is
 * absorbed and added to `db` (call dcpgpu_db_commit afterwards).  PROTEIN_CFG_DEFAULT is
 * {ENTRY_DIST_OCCUPANCY, 0.01} (protein_cfg.h:13,22-23). */
enum rc dcpgpu_press_hmm(struct dcpgpu_db *, FILE *hmm, struct protein_cfg cfg, unsigned *nprofiles);
/* accession / core size of profile i of the database */
char const *dcpgpu_db_accession(struct dcpgpu_db const *, unsigned i);
unsigned dcpgpu_db_core_size(struct dcpgpu_db const *, unsigned i);

/* Stage sequences (ASCII ACGT, not NUL-terminated; lens[i] nucleotides each) on the db's
 * device.  RC_EINVAL for an empty sequence (protein_profile.c:158) or a non-ACGT symbol. */
enum rc dcpgpu_seqs_new(struct dcpgpu_seqs **out, struct dcpgpu_db *, unsigned nseqs,
                        char const *const *seqs, unsigned const *lens);
void dcpgpu_seqs_del(struct dcpgpu_seqs *);

/* Run null + alt Viterbi, LRT filter and (want_paths) traceback for every
 * (sequence, profile) pair; inputs already resident in HBM.  Results are ordered
 * sequence-major, profile-minor, independent of how the GPU scheduled the pairs. */
enum rc dcpgpu_scan_resident(struct dcpgpu_db *, struct dcpgpu_seqs *, struct dcpgpu_params const *,
                             struct dcpgpu_result **out);
/* Same, from host buffers: stages the sequences, scans, copies results back.  A batch whose row records and
 * pair scores would not fit the device's free memory is run as several launch sets (64 B per nucleotide and
 * null table, 5 B per pair, at most 2^30 pairs each) and the results are merged, so nseqs is not bounded by
 * HBM.  A non-ACGT symbol fails the whole batch with RC_EINVAL, as an imm_seq error fails the reference's
 * whole job (scan.c:229-231, 244-256). */
enum rc dcpgpu_scan(struct dcpgpu_db *, unsigned nseqs, char const *const *seqs,
                    unsigned const *lens, struct dcpgpu_params const *, struct dcpgpu_result **out);

unsigned dcpgpu_result_nseqs(struct dcpgpu_result const *);
unsigned dcpgpu_result_nprofiles(struct dcpgpu_result const *);
/* [nseqs * nprofiles], index seq * nprofiles + prof; prod.null_loglik / alt_loglik */
float const *dcpgpu_result_null_loglik(struct dcpgpu_result const *);
float const *dcpgpu_result_alt_loglik(struct dcpgpu_result const *);
uint8_t const *dcpgpu_result_hit(struct dcpgpu_result const *);
uint64_t dcpgpu_result_nhits(struct dcpgpu_result const *);
/* i-th hit in (sequence, profile) order: its pair and, when paths were requested, its path
 * (imm_path_nsteps / imm_path_step, prod.c:162-176). */
enum rc dcpgpu_result_hit_at(struct dcpgpu_result const *, uint64_t i, unsigned *seq_idx,
                             unsigned *prof_idx, struct dcp_step const **steps, unsigned *nsteps);
/* The whole hit list at once, (sequence, profile) order: nhits entries per array (any may be NULL); alt / null are
 * the fp32 log-likelihoods prod.c prints.  _steps: all paths back to back in hit order (hit i owns nsteps[i] of
 * them); the pointer stays valid until dcpgpu_result_del. */
enum rc dcpgpu_result_hits(struct dcpgpu_result const *, unsigned *seq_idx, unsigned *prof_idx, float *alt_loglik,
                           float *null_loglik, unsigned *nsteps);
uint64_t dcpgpu_result_steps(struct dcpgpu_result const *, struct dcp_step const **steps);
/* device time of the kernels of the last scan on this db, by phase (ms, CUDA events) */
struct dcpgpu_timing
{
    float prep_ms;   /* row records + null Viterbi */
    float score_ms;  /* alt Viterbi score pass (the hot kernel) */
    float trace_ms;  /* traceback pass for hits */
    float total_ms;  /* first launch to last completion, including copies issued by the scan */
    uint64_t launches;   /* kernels launched by the scan */
    uint64_t alt_cells;  /* sum over pairs of L * M */
    uint64_t h2d_bytes, d2h_bytes;
};
void dcpgpu_result_timing(struct dcpgpu_result const *, struct dcpgpu_timing *);
enum rc dcpgpu_result_part_timing(struct dcpgpu_result const *, unsigned part, int *device, struct dcpgpu_timing *);
void dcpgpu_result_del(struct dcpgpu_result *);

/* ------------------------------------------------------------------------- */
/* Part 2b -- multi-device.  Replaces scan_run's omp-parallel-for over profile  */
/* partitions (scan.c:239-250, profile_reader.c:54-72) and the ordered          */
/* concatenation of per-thread products (prod.c:106-145): a partition is a GPU. */
/* ------------------------------------------------------------------------- */
/* Longest-processing-time partition of profiles over `nshards` devices by MODELLED COST: padded width of the
 * profile's kernel class / measured rate of that class (replaces the equal-count partition of
 * src/db/profile_reader.c:54-72).  Deterministic. */
enum rc dcpgpu_shard_profiles(unsigned nprofiles, unsigned const *core_sizes, unsigned nshards,
                              unsigned *shard_of);
/* modelled score-pass time of one sequence row against a profile of `core_size` nodes (ns on one B200) */
double dcpgpu_profile_cost(unsigned core_size);
/* contiguous ranges of sequences with about equal nucleotide totals; bounds has nshards + 1 entries */
enum rc dcpgpu_shard_sequences(unsigned nseqs, unsigned const *lens, unsigned nshards, unsigned *bounds);

struct dcpgpu_mdb; /* a database sharded (or replicated) over several GPUs of one box */
enum dcpgpu_axis
{
    DCPGPU_AXIS_AUTO,      /* profiles when their modelled costs balance within 10 %, else sequences */
    DCPGPU_AXIS_PROFILES,  /* each device holds a cost-balanced shard of the profiles, scans all sequences */
    DCPGPU_AXIS_SEQUENCES, /* each device holds every profile, scans a contiguous range of the sequences */
};
/* devices: CUDA device ordinals, one shard each (an ordinal may repeat: its shards then share that GPU) */
enum rc dcpgpu_mdb_new(struct dcpgpu_mdb **out, unsigned ndevices, int const *devices);
enum rc dcpgpu_mdb_add(struct dcpgpu_mdb *, struct protein_profile const *prof);
/* shard, upload (one host thread per device) and release the host tables */
enum rc dcpgpu_mdb_commit(struct dcpgpu_mdb *, enum dcpgpu_axis axis);
/* every profile in global order: pass it to dcpgpu_press_hmm before the commit, and to dcpgpu_prod_* /
 * dcpgpu_db_accession with the results of dcpgpu_mdb_scan */
struct dcpgpu_db *dcpgpu_mdb_view(struct dcpgpu_mdb *);
unsigned dcpgpu_mdb_ndevices(struct dcpgpu_mdb const *);
unsigned dcpgpu_mdb_nprofiles(struct dcpgpu_mdb const *);
enum dcpgpu_axis dcpgpu_mdb_axis(struct dcpgpu_mdb const *);
double dcpgpu_mdb_imbalance(struct dcpgpu_mdb const *); /* modelled max / mean shard cost */
int dcpgpu_mdb_device_of(struct dcpgpu_mdb const *, unsigned profile); /* -1 when replicated */
uint64_t dcpgpu_mdb_device_bytes(struct dcpgpu_mdb const *, unsigned shard);
/* dcpgpu_scan over all devices; hits come back merged in (sequence, global profile) order, bit-identical to a
 * single-device scan of the same database whatever the device count or axis */
enum rc dcpgpu_mdb_scan(struct dcpgpu_mdb *, unsigned nseqs, char const *const *seqs, unsigned const *lens,
                        struct dcpgpu_params const *, struct dcpgpu_result **out); /* delete `out` before the mdb */
void dcpgpu_mdb_del(struct dcpgpu_mdb *);
/* launch sets a (merged) result was built from, and the device / phase times of each: the per-device busy
 * times of a multi-device scan (their spread is the shard imbalance) */
unsigned dcpgpu_result_nparts(struct dcpgpu_result const *);

/* How the engine maps a profile of `core_size` nodes (1..4096, limits.h:11) onto the GPU: `warps` per
 * (sequence, profile) pair, `nodes_per_lane`, and 1 or 2 thread blocks (a cluster) per pair; warps * 32 *
 * nodes_per_lane >= core_size is the padded width the kernels compute.  Pure function (no device needed). */
enum rc dcpgpu_kernel_shape(unsigned core_size, unsigned *warps, unsigned *nodes_per_lane, unsigned *blocks);
/* the padded width itself: warps * 32 * nodes_per_lane, or 16 * nodes_per_lane for profiles short enough that two
 * (sequence, profile) pairs share a warp, 16 lanes each (0 for an invalid core size) */
unsigned dcpgpu_kernel_padded_width(unsigned core_size);

/* ------------------------------------------------------------------------- */
/* Part 3 -- products.  src/server/prod.c:13-41,106-181, protein_match.c:21-56 */
/* ------------------------------------------------------------------------- */
/* prod_fclose's header line (prod.c:115-117) */
enum rc dcpgpu_prod_fwrite_header(FILE *fp);
/* One row per hit, in (sequence, profile) order -- prod_fwrite + protein_match_write_func.
 * seq_ids may be NULL (then the sequence index is used).  Requires want_paths. */
enum rc dcpgpu_prod_fwrite(struct dcpgpu_result const *, struct dcpgpu_db const *, FILE *fp,
                           int64_t scan_id, int64_t const *seq_ids, unsigned nseqs,
                           char const *const *seqs);
/* Row for a single hit into a caller buffer; returns length or -1 if it does not fit. */
long dcpgpu_prod_row(struct dcpgpu_result const *, struct dcpgpu_db const *, uint64_t hit,
                     int64_t scan_id, int64_t seq_id, char const *seq, char *out, long cap);

/* Measured FP32 issue peaks of `device` (the roofline denominators of the score kernel):
 * out[0] FADD, out[1] FMNMX3, out[2] the DP cell's 2:1 FADD:FMNMX3 mix, in 1e9 lane-instructions/s;
 * out[3] = SM count; out[4], out[5] = SM clock in MHz measured inside the mix / FADD kernel. */
enum rc dcpgpu_microbench_alu(int device, double out[6]);

char const *dcpgpu_last_error(void); /* thread-local message of the last failure */

#ifdef __cplusplus
}
#endif
#endif
