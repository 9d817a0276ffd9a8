/*
 * dcp_classes.h -- the kernel class table: every (warps per pair, nodes per lane, occupancy) shape the score
 * and trace kernels are instantiated for, with its measured rate.  One X-macro feeds the host-side chooser
 * and cost model (dcp_shape.c) and the CUDA launchers (dcp_score_sw.cu, dcp_score_mw.cu, dcp_trace.cu).
 *
 *   X(TW, Q, BPS, RATE)
 *   TW    warps per (sequence, profile) pair: 1 = k_score<Q> (one warp per pair); 2..8 = k_score_mw<TW, 1, Q, BPS>
 *         (a group of warps in one block); 10, 12, 14, 16 = k_score_mw<TW / 2, 2, Q, 1> (two blocks of a cluster)
 *   Q     core nodes per lane: the class holds profiles of up to TW * 32 * Q nodes
 *   BPS   resident blocks per SM the kernel is compiled for (TW = 1: warps per block, one block per SM).  The
 *         register file gives 8 warps per SM at 255 registers a thread, 12 at 168, 16 at 128.
 *   RATE  measured score-pass rate in 1e9 padded (row, node) cells per second on one B200
 *         (tools/class_sweep.py, profiles/r02_class_sweep*.jsonl); 0 = compiled in for experiments
 *         (DCPGPU_FORCE_SHAPE) but never chosen.
 *
 * A profile of M nodes runs in the class that minimises padded width / RATE among those that hold it
 * (dcp_kernel_shape); the same quotient is the profile's weight when shards are balanced (dcp_profile_cost).
 */
#ifndef DCP_CLASSES_H
#define DCP_CLASSES_H

#define DCP_CLASS_TABLE(X)                                                                                      \
    /* one warp per pair */                                                                                     \
    X(1, 1, 16, 206) X(1, 2, 16, 414) X(1, 3, 12, 441) X(1, 4, 12, 600)                                         \
    X(1, 5, 8, 576) X(1, 6, 8, 646) X(1, 7, 8, 0) X(1, 8, 8, 730)                                               \
    /* two warps */                                                                                             \
    X(2, 5, 6, 0) X(2, 6, 4, 463) X(2, 6, 6, 0) X(2, 7, 4, 480) X(2, 8, 4, 520)                                 \
    /* three warps */                                                                                           \
    X(3, 5, 4, 0) X(3, 6, 2, 295) X(3, 6, 4, 0) X(3, 7, 2, 315) X(3, 8, 2, 343)                                 \
    /* four to eight warps */                                                                                   \
    X(4, 5, 3, 0) X(4, 6, 3, 0) X(4, 8, 2, 431)                                                                 \
    X(5, 5, 2, 0) X(5, 6, 2, 0) X(5, 8, 1, 272)                                                                 \
    X(6, 5, 2, 0) X(6, 6, 2, 0) X(6, 8, 1, 320)                                                                 \
    X(7, 8, 1, 353) X(8, 8, 1, 387)                                                                             \
    /* two blocks of a cluster */                                                                               \
    X(10, 8, 1, 170) X(12, 8, 1, 195) X(14, 8, 1, 216) X(16, 8, 1, 230)

#endif
