/* dcp_internal.h -- private structs shared by the host C files and the CUDA engine. */
#ifndef DCP_INTERNAL_H
#define DCP_INTERNAL_H

#include "dcpgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* include/deciphon/model/nuclt_dist.h:7-11, kept in double on the host (decode only) */
struct dcp_nuclt_dist
{
    double nucltp[4];
    double codonm[125]; /* [a][b][c], 4 = any */
};

/* include/deciphon/model/protein_model.h:16-48 without the imm_hmm objects */
struct protein_model
{
    struct protein_cfg cfg;
    unsigned core_size;
    char consensus[DCP_PROTEIN_MODEL_CORE_SIZE_MAX + 1];
    float null_lprobs[DCP_AMINO_SIZE];
    struct dcp_nuclt_dist null_ndist;
    struct dcp_nuclt_dist insert_ndist;
    unsigned node_idx;
    struct dcp_nuclt_dist *match_ndists; /* [core_size] */
    unsigned trans_idx;
    struct protein_trans *trans; /* [core_size + 1] */
};

/* include/deciphon/model/protein_profile.h:12-43; the two imm_dp objects are replaced by
 * the explicit DP-level tables the kernels consume */
struct protein_profile
{
    char accession[DCP_PROFILE_ACC_SIZE];
    struct protein_cfg cfg;
    unsigned core_size;
    char *consensus;
    struct dcp_nuclt_dist null_ndist;
    struct dcp_nuclt_dist insert_ndist;
    struct dcp_nuclt_dist *match_ndists;
    float null_emission[DCP_FRAME_TABLE_SIZE];   /* R, N, J, C */
    float insert_emission[DCP_FRAME_TABLE_SIZE]; /* every I_k */
    float *match_emission;                       /* [core_size][1364] */
    struct protein_trans *trans;                 /* [core_size + 1] */
    float *entry;                                /* [core_size]  B -> M_k */
};

void dcp_nuclt_dist_setup(struct dcp_nuclt_dist *, double const amino_lprobs[DCP_AMINO_SIZE]);
void dcp_frame_table(struct dcp_nuclt_dist const *, double eps, float out[DCP_FRAME_TABLE_SIZE]);
unsigned dcp_frame_code(unsigned len, unsigned packed);
void dcp_specials(unsigned seq_size, bool multi_hits, bool hmmer3_compat, float x[13]);
int dcp_nuclt_index(char c);
char dcp_gc_decode(int a, int b, int c);

/*
 * How the engine maps a profile onto the GPU (dcp_shape.c, dcp_classes.h): a kernel class is a row of the
 * class table -- warps per (sequence, profile) pair, nodes per lane, resident blocks per SM, measured rate.
 */
enum
{
    DCP_MAX_Q = 8,            /* nodes per lane; one warp holds M <= 256 */
    DCP_MAX_W = 8,            /* warps per block in the multi-warp classes: M <= 2048 per block */
    DCP_MAX_GROUP_WARPS = 16, /* two blocks (a cluster) per pair above 2048 nodes: M <= 4096 */
    DCP_MAX_CLASSES = 64,
};
struct dcp_class
{
    unsigned tw, q, bps;
    double rate;
};
unsigned dcp_num_classes(void);
struct dcp_class const *dcp_class_at(unsigned cls);
/* the class a profile of `core_size` nodes runs in (index into the table) */
unsigned dcp_kernel_class(unsigned core_size);
/* modelled score-pass time of one sequence row against a profile of `core_size` nodes, in ns of one whole
 * B200 (padded nodes / measured padded-node rate of the profile's kernel class); the shard weights */
double dcp_profile_cost(unsigned core_size);

/* error reporting: message kept per thread, code returned (logging.h:32-72 convention) */
void dcp_set_error(char const *msg);
enum rc dcp_error(enum rc rc, char const *msg);

#ifdef __cplusplus
}
#endif
#endif
