/* dcp_kernels.cuh -- device helpers shared by the score and trace kernels. */
#ifndef DCP_KERNELS_CUH
#define DCP_KERNELS_CUH

#include "dcp_engine.h"

template <int Q>
struct NodeParams
{
    /* incoming to node k from node k-1 (trans[k-1]) */
    float MM[Q], IM[Q], DM[Q], MD[Q], DD[Q];
    /* own insert loop (trans[k]) and entry B->M_k */
    float MI[Q], II[Q], ent[Q];
};

__device__ __forceinline__ float warp_max(float x)
{
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

/* tr = [8 params][NP node slots]; this lane owns slots first .. first+Q-1 */
template <int Q>
__device__ __forceinline__ void load_params(NodeParams<Q> &p, const float *__restrict__ tr, int NP, int first)
{
#pragma unroll
    for (int i = 0; i < Q; ++i)
    {
        int n = first + i;
        p.MM[i] = __ldg(tr + 0 * NP + n);
        p.IM[i] = __ldg(tr + 1 * NP + n);
        p.DM[i] = __ldg(tr + 2 * NP + n);
        p.MD[i] = __ldg(tr + 3 * NP + n);
        p.DD[i] = __ldg(tr + 4 * NP + n);
        p.MI[i] = __ldg(tr + 5 * NP + n);
        p.II[i] = __ldg(tr + 6 * NP + n);
        p.ent[i] = __ldg(tr + 7 * NP + n);
    }
}

/* frame-table codes of the windows ending at a row, from its packed 10-bit window */
__device__ __forceinline__ void codes_of(uint32_t w, uint32_t (&c)[5])
{
    c[0] = w & 3u;
    c[1] = 4u + (w & 15u);
    c[2] = 20u + (w & 63u);
    c[3] = 84u + (w & 255u);
    c[4] = 340u + (w & 1023u);
}

__device__ __forceinline__ void load_row_insert(const RowRec *__restrict__ r, float (&eI)[5])
{
    const float4 *q = reinterpret_cast<const float4 *>(r);
    float4 a = __ldg(q);
    float e4 = __ldg(reinterpret_cast<const float *>(q + 1));
    eI[0] = a.x, eI[1] = a.y, eI[2] = a.z, eI[3] = a.w, eI[4] = e4;
}

__device__ __forceinline__ void load_row_special(const RowRec *__restrict__ r, float (&eN)[5])
{
    const float4 *q = reinterpret_cast<const float4 *>(r) + 2;
    float4 a = __ldg(q);
    float e4 = __ldg(reinterpret_cast<const float *>(q + 1));
    eN[0] = a.x, eN[1] = a.y, eN[2] = a.z, eN[3] = a.w, eN[4] = e4;
}

/*
 * Match emissions of one row: for each of the five lengths one line of the transposed table,
 * [code][half][lane][4] -- every LDG.128 of a warp covers 512 contiguous bytes.
 * emis_lane = table base + lane * 4.  (L1 eviction hints -- evict_last on the 1-3 nt lines,
 * no_allocate on the 5 nt lines -- were measured and did not help: 381 vs 399 GCUPS.)
 */
#ifndef DCP_POLICY
#define DCP_POLICY 0
#endif
/*
 * Two layouts of an emission line.  Split: [code][warp][half][lane][4], a lane's first quad by one 128-bit load and
 * its second quad by a 4..16-byte load 512 bytes further on.  Whole: [code][warp][lane][8], a lane's eight floats by one
 * 256-bit load (SASS LDG.E.ENL2.256): five instead of ten loads per row.  Which one is faster is a property of the
 * kernel's schedule, not of the memory system -- measured per class (tools/r2_e256_probe.sh, kernel-only GCUPS, whole
 * against split): (2,8) +3.9 %, (2,5) +2.7 %, (3,6) +1.2 %, (0,8) +1.1 %, (4,6) +0.8 %, (1,8) +0.6 %, (1,6) +0.4 %;
 * (2,6) -2.9 %, (2,7) -3.0 %, (8,8) -1.8 %, (0,5) -1.7 %, (1,5) -1.2 %, (4,8) -0.9 % -- so the layout is chosen by
 * class (TW = warps per pair, 0 for the half-warp classes; Q = nodes per lane).  DCP_EMIS256_ALL = 0 / 1 forces
 * split / whole everywhere (variant builds).
 */
#ifndef DCP_EMIS256_ALL
#define DCP_EMIS256_ALL -1
#endif
__host__ __device__ constexpr bool emis256(int TW, int Q)
{
    return Q <= 4 ? false
           : DCP_EMIS256_ALL >= 0
               ? DCP_EMIS256_ALL != 0
               : (TW == 0 && Q == 8) || (TW == 1 && (Q == 6 || Q == 8)) || (TW == 2 && (Q == 5 || Q == 8)) ||
                     (TW == 3 && Q == 6) || (TW == 4 && Q == 6);
}
/* floats between consecutive lanes' data in a line, and the position of a lane's sub-node inside its warp-unit */
__host__ __device__ constexpr int emis_lane_stride(int TW, int Q) { return emis256(TW, Q) ? 8 : 4; }
__host__ __device__ inline uint32_t emis_pos(uint32_t lane, uint32_t sub, uint32_t LN, int TW, int Q)
{
    return emis256(TW, Q) ? lane * 8 + sub : (sub >> 2) * (LN * 4) + lane * 4 + (sub & 3);
}

/* L1 policy experiments for the emission lines, by window length l (0..4 = 1..5 nt):
 * 0 default everywhere; 1: 5-nt lines no_allocate; 2: 4- and 5-nt lines no_allocate;
 * 3: 1..3-nt lines evict_last, 5-nt no_allocate; 4: 1..3 evict_last, 4..5 no_allocate */
__device__ __forceinline__ float4 ldg4(const float *p, int L)
{
    float4 v;
    const bool keep = (DCP_POLICY == 3 || DCP_POLICY == 4) && L < 3;
    const bool stream = (DCP_POLICY == 1 && L == 4) || (DCP_POLICY == 2 && L >= 3) ||
                        (DCP_POLICY == 3 && L == 4) || (DCP_POLICY == 4 && L >= 3);
    if (keep)
        asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    else if (stream)
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    else
        v = __ldg(reinterpret_cast<const float4 *>(p));
    return v;
}

/* ROW = floats per table line of the profile; HOFF = floats between a lane's first and second quad (lanes per
 * pair x 4: 128 for a whole warp per pair, 64 for a half-warp per pair) */
template <int Q, int L0, int L1, int ROW = 32 * (Q <= 4 ? 4 : 8), int HOFF = 128, bool E256 = false>
__device__ __forceinline__ void load_emis_part(float (&em)[5][Q], const float *__restrict__ emis_lane,
                                               const uint32_t (&code)[5])
{
#pragma unroll
    for (int l = L0; l < L1; ++l)
    {
        const float *src = emis_lane + (size_t)code[l] * ROW;
        if (E256)
        {
            float t[8];
            asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                : "=f"(t[0]), "=f"(t[1]), "=f"(t[2]), "=f"(t[3]), "=f"(t[4]), "=f"(t[5]), "=f"(t[6]), "=f"(t[7])
                : "l"(src));
#pragma unroll
            for (int i = 0; i < Q; ++i) em[l][i] = t[i];
            continue;
        }
        if (Q >= 4)
        {
            float4 a = ldg4(src, l);
            em[l][0] = a.x, em[l][1] = a.y, em[l][2] = a.z, em[l][3] = a.w;
        }
        else
        {
            /* Q < 4: only the first Q floats of the lane's quad are real */
            float4 a = __ldg(reinterpret_cast<const float4 *>(src));
            float t[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int i = 0; i < Q; ++i) em[l][i] = t[i];
        }
        /* second half (+128 floats) holds nodes 4..Q-1 of the lane: load exactly Q-4 floats */
        if (Q == 8)
        {
            float4 b = ldg4(src + HOFF, l);
            em[l][4 % Q] = b.x, em[l][5 % Q] = b.y, em[l][6 % Q] = b.z, em[l][7 % Q] = b.w;
        }
        else if (Q == 7)
        {
#if DCP_Q7_LOAD8
            /* one 16-byte load (the lane's fourth float is padding).  volatile: ptxas otherwise narrows a v4 load
             * with a dead component into three 4-byte loads (25 instead of 15 loads per row) */
            float4 b;
            asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(src + HOFF));
            em[l][4 % Q] = b.x, em[l][5 % Q] = b.y, em[l][6 % Q] = b.z;
#else
            float2 b = __ldg(reinterpret_cast<const float2 *>(src + HOFF));
            float c = __ldg(src + HOFF + 2);
            em[l][4 % Q] = b.x, em[l][5 % Q] = b.y, em[l][6 % Q] = c;
#endif
        }
        else if (Q == 6)
        {
            float2 b = __ldg(reinterpret_cast<const float2 *>(src + HOFF));
            em[l][4 % Q] = b.x, em[l][5 % Q] = b.y;
        }
        else if (Q == 5)
        {
            em[l][4 % Q] = __ldg(src + HOFF);
        }
    }
}

template <int Q, int ROW = 32 * (Q <= 4 ? 4 : 8), int HOFF = 128, bool E256 = false>
__device__ __forceinline__ void load_emis(float (&em)[5][Q], const float *__restrict__ emis_lane,
                                          const uint32_t (&code)[5])
{
    load_emis_part<Q, 0, 5, ROW, HOFF, E256>(em, emis_lane, code);
}

/* ----------------------------------------------------------------------------------------- */
/* group of warps sharing one pair: W warps of one block (CL = 1) or of a 2-block cluster     */
/* ----------------------------------------------------------------------------------------- */
#include <cooperative_groups.h>
#include <cstddef>
namespace cg = cooperative_groups;


/*
 * Barrier / broadcast helpers.  With CL = 2 every value a warp publishes is stored into BOTH
 * blocks' shared memory (distributed shared memory, cluster.map_shared_rank) so that reads stay
 * local, and the block barrier becomes a cluster barrier.
 */
template <int CL, class Shared>
struct Group
{
    Shared *me, *peer;
    int rank;

    __device__ __forceinline__ void init(Shared *mine)
    {
        me = mine, peer = nullptr, rank = 0;
        if (CL == 2)
        {
            cg::cluster_group cl = cg::this_cluster();
            rank = (int)cl.block_rank();
            peer = cl.map_shared_rank(mine, rank ^ 1);
        }
    }
    __device__ __forceinline__ void sync()
    {
        if (CL == 2) cg::this_cluster().sync();
        else __syncthreads();
    }
    /*
     * 2-block groups, score kernel: barrier + all-to-all of a few floats per warp in one step, without the cluster
     * barrier and without a fence (barrier.cluster.arrive.release costs MEMBAR.ALL.GPU, which also drains the
     * emission loads in flight for the next row).  Lane 31 of every warp stores its 8 or 12 floats into its own block
     * (st.shared) and into the peer block (st.async through DSMEM, SASS STAS.128: the bytes are counted on the peer's
     * mbarrier as they land), then arrives on its own block's mbarrier expecting the bytes of its mirror warp.  The
     * phase completes when the block's W warps have arrived and all bytes of the peer's W warps have landed.  Two
     * mbarriers / buffers alternate; a block cannot be two exchanges ahead of its peer (each exchange needs the peer's
     * bytes), so bytes of exchange n + 2 never reach a barrier that is still in exchange n.
     * `Shared` provides `float xch[2][kMaxGroupWarps][8 or 12]` (16-byte aligned) and `unsigned long long xbar[2]`.
     */
    uint32_t me_s, peer_s, xstate; /* shared-window addresses of *me in this block / the peer; slot + parities */
    __device__ __forceinline__ void exchange_init(int block_warps)
    {
        me_s = (uint32_t)__cvta_generic_to_shared(me);
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_s) : "r"(me_s), "r"(rank ^ 1));
        xstate = 0;
        if (threadIdx.x == 0)
        {
            const uint32_t b = me_s + (uint32_t)offsetof(Shared, xbar);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(block_warps) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b + 8), "r"(block_warps) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        cg::this_cluster().sync();
    }
    /* the NF floats of v are taken from lane 31; returns the buffer index to read me->xch[..] from */
    static constexpr int NF = (int)(sizeof(((Shared *)nullptr)->xch[0][0]) / sizeof(float)); /* floats per warp: 8, 12 */
    __device__ __forceinline__ int exchange(int gw, int lane, const float (&v)[NF])
    {
        static_assert(NF % 4 == 0, "whole 16-byte stores");
        const uint32_t s = xstate & 1u;
        const uint32_t bar = me_s + (uint32_t)offsetof(Shared, xbar) + s * 8u;
        if (lane == 31)
        {
            const uint32_t off = (uint32_t)offsetof(Shared, xch) + (s * kMaxGroupWarps + (uint32_t)gw) * (NF * 4u);
            const uint32_t rbar = peer_s + (uint32_t)offsetof(Shared, xbar) + s * 8u;
#pragma unroll
            for (int q = 0; q < NF; q += 4)
            {
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(me_s + off + q * 4), "f"(v[q]), "f"(v[q + 1]),
                             "f"(v[q + 2]), "f"(v[q + 3])
                             : "memory");
                asm volatile(
                    "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                        peer_s + off + q * 4),
                    "r"(__float_as_uint(v[q])), "r"(__float_as_uint(v[q + 1])), "r"(__float_as_uint(v[q + 2])),
                    "r"(__float_as_uint(v[q + 3])), "r"(rbar)
                    : "memory");
            }
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(NF * 4u) : "memory");
        }
        const uint32_t parity = (xstate >> (1u + s)) & 1u;
        uint32_t ok = 0;
        for (uint32_t spin = 0; !ok; ++spin)
        {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                         " selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(ok)
                         : "r"(bar), "r"(parity)
                         : "memory");
            if (spin > (1u << 24)) __trap(); /* a lost store must not hang the GPU */
        }
        xstate ^= (2u << s) | 1u;
        return (int)s;
    }
    /* OR of a predicate over every thread of the group; `flag` = two ints per block, caller alternates slot */
    __device__ __forceinline__ bool any(bool pred, int (*flag)[2], int (*peer_flag)[2], int slot)
    {
        bool r = __syncthreads_or(pred);
        if (CL == 2)
        {
            if (threadIdx.x == 0) flag[slot][rank] = r, peer_flag[slot][rank] = r;
            cg::this_cluster().sync();
            r = flag[slot][0] | flag[slot][1];
        }
        return r;
    }
};
#define GRP_PUT(grp, field, value)                                                                  \
    do                                                                                              \
    {                                                                                               \
        (grp).me->field = (value);                                                                  \
        if (CL == 2) (grp).peer->field = (value);                                                   \
    } while (0)

/* launch helper: plain launch for one block per pair, cluster launch (2 blocks) above 2048 nodes */
template <class... Params, class... Args>
static cudaError_t launch_group(void (*kernel)(Params...), int cl, unsigned blocks, unsigned threads, cudaStream_t st,
                                Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks), cfg.blockDim = dim3(threads), cfg.dynamicSmemBytes = 0, cfg.stream = st;
    cudaLaunchAttribute attr[1];
    if (cl == 2)
    {
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr, cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<Params>(args)...);
}


#endif
