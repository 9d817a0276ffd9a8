/*
 * dcp_classes.h -- the kernel class table: every (warps per pair, nodes per lane, occupancy) shape the score
 * and trace kernels are instantiated for, with its measured rate.  One X-macro feeds the host-side chooser
 * and cost model (dcp_shape.c) and the CUDA launchers (dcp_score_sw.cu, dcp_score_mw.cu, dcp_trace.cu).
 *
 *   X(TW, Q, BPS, RATE)
 *   TW    warps per (sequence, profile) pair: 0 = k_score_h<Q> (half a warp per pair, two pairs per warp);
 *         1 = k_score<Q> (one warp per pair); 2..8 = k_score_mw<TW, 1, Q, BPS>
 *         (a group of warps in one block); 10, 12, 14, 16 = k_score_mw<TW / 2, 2, Q, 1> (two blocks of a cluster)
 *   Q     core nodes per lane: the class holds profiles of up to TW * 32 * Q nodes (16 * Q for TW = 0)
 *   BPS   resident blocks per SM the kernel is compiled for (TW = 1: warps per block, one block per SM).  The
 *         register file gives 8 warps per SM at 255 registers a thread, 12 at 168, 16 at 128.
 *   RATE  measured score-pass rate in 1e9 padded (row, node) cells per second on one B200
 *         (tools/class_sweep.py, profiles/r02_class_sweep*.jsonl; the seven classes that read whole 32-byte emission
 *         lines per lane -- dcp_kernels.cuh: emis256 -- scaled by their measured gain, tools/r2_e256_probe.sh);
 *         0 = compiled in for experiments
 *         (DCPGPU_FORCE_SHAPE) but never chosen.
 *
 * A profile of M nodes runs in the class that minimises padded width / RATE among those that hold it
 * (dcp_kernel_shape); the same quotient is the profile's weight when shards are balanced (dcp_profile_cost).
 */
#ifndef DCP_CLASSES_H
#define DCP_CLASSES_H

#define DCP_CLASS_TABLE(X)                                                                                      \
    /* TW = 0: two pairs per warp, 16 lanes each (k_score_h<Q>): profiles of up to 16 Q nodes */                \
    X(0, 4, 12, 568) X(0, 5, 8, 584) X(0, 6, 8, 649) X(0, 8, 8, 715)                                            \
    /* one warp per pair, 8 resident warps per SM */                                                            \
    X(1, 5, 8, 573) X(1, 6, 8, 647) X(1, 7, 8, 0) X(1, 8, 8, 730)                                                          \
    /* two warps: 12 resident warps per SM with 5 nodes per lane (168 registers), else 8 */                     \
    X(2, 5, 6, 480) X(2, 6, 4, 485) X(2, 7, 4, 504) X(2, 8, 4, 582)                                             \
    /* three warps: 12 resident warps (6 at 255 registers leave the schedulers idle) */                         \
    X(3, 6, 4, 387)                                                                                             \
    /* four warps */                                                                                            \
    X(4, 5, 3, 393) X(4, 6, 2, 407) X(4, 7, 2, 436) X(4, 8, 2, 468)                                             \
    /* six and eight warps: 12 or 8 resident warps per SM; five or seven leave issue slots idle */                       \
    X(6, 6, 2, 359) X(8, 6, 1, 355) X(8, 7, 1, 391) X(8, 8, 1, 417)                             \
    /* two blocks of a cluster */                                                                               \
    X(16, 6, 1, 265) X(16, 8, 1, 308)

/* Measured and left out (rates before the straight-line row layout, which lifted every multi-warp class by 5..30 %;
 * each loses to a neighbour in padded width / rate):
 * (1,1,16: 193) (1,2,16: 406) (1,3,12: 437) (1,4,12: 596: profiles of up to 128 nodes now share a warp two by two) (1,7,8: 618 -- see DESIGN.md 6.2: ptxas sinks the next row's emission loads to mid-row,
 * long-scoreboard stalls 0.46 per issue against 0.07 at 8 nodes per lane) (2,5,4: 332) (2,6,6: 408,
 * spills) (3,5,4: 311) (3,6,2: 294) (3,7,2: 315) (3,8,2: 342) (4,5,2: 269) (4,6,3: 311, spills) (5,5,2: 315)
 * (5,6,2: 286) (6,5,2: 286) (6,8,1: 320) (7,8,1: 352) (8,5,1: 259) (5,8,1: 285 and 14,8: 266 with the new layout, 10,8: 197) (12,8: 195) (16,5: 171) (16,7: 211);
 * profiles/r02_class_sweep_table.jsonl */

#endif
