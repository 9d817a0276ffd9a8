/* dcp_kernels.cuh -- device helpers shared by the score and trace kernels. */
#ifndef DCP_KERNELS_CUH
#define DCP_KERNELS_CUH

#include "dcp_engine.h"

#ifndef DCP_HINTS
#define DCP_HINTS 0
#endif

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

template <int Q>
__device__ __forceinline__ void load_params(NodeParams<Q> &p, const float *__restrict__ tr, int lane)
{
    constexpr int NP = 32 * Q;
#pragma unroll
    for (int i = 0; i < Q; ++i)
    {
        int n = lane * Q + i;
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

/* 128-bit read-only loads with an L1 eviction hint (PTX ld.global.nc.L1::*) */
__device__ __forceinline__ float4 ldg_keep(const float4 *p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ldg_stream(const float4 *p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

/*
 * Lines of the 1-, 2- and 3-nt windows (84 codes, 84 KB per profile at QP = 8) are reused by
 * every row of every warp working on the profile: keep them in L1 (evict_last).  The 5-nt lines
 * (1024 codes, 1 MB) are a stream with almost no reuse inside one SM: do not allocate them in L1,
 * L2 serves them.  4-nt lines (256 codes) take the default policy.
 */
template <int Q, int L0, int L1>
__device__ __forceinline__ void load_emis_part(float (&em)[5][Q], const float *__restrict__ emis_lane,
                                               const uint32_t (&code)[5])
{
    constexpr int QP = Q <= 4 ? 4 : 8;
    constexpr int ROW = 32 * QP;
#pragma unroll
    for (int l = L0; l < L1; ++l)
    {
        const float4 *src = reinterpret_cast<const float4 *>(emis_lane + (size_t)code[l] * ROW);
#if DCP_HINTS
        float4 a = l < 3 ? ldg_keep(src) : (l == 4 ? ldg_stream(src) : __ldg(src));
#else
        float4 a = __ldg(src);
#endif
        float tmp[8];
        tmp[0] = a.x, tmp[1] = a.y, tmp[2] = a.z, tmp[3] = a.w;
        if (Q > 4)
        {
            /* second half: +128 floats */
#if DCP_HINTS
            float4 b = l < 3 ? ldg_keep(src + 32) : (l == 4 ? ldg_stream(src + 32) : __ldg(src + 32));
#else
            float4 b = __ldg(src + 32);
#endif
            tmp[4] = b.x, tmp[5] = b.y, tmp[6] = b.z, tmp[7] = b.w;
        }
#pragma unroll
        for (int i = 0; i < Q; ++i) em[l][i] = tmp[i];
    }
}

template <int Q>
__device__ __forceinline__ void load_emis(float (&em)[5][Q], const float *__restrict__ emis_lane,
                                          const uint32_t (&code)[5])
{
    load_emis_part<Q, 0, 5>(em, emis_lane, code);
}

#endif
