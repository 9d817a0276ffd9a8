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

struct RowIn
{
    float eN[5], eI[5];
    uint32_t code[5];
};

__device__ __forceinline__ RowIn load_row(const RowRec *__restrict__ r)
{
    const float4 *q = reinterpret_cast<const float4 *>(r);
    float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
    RowIn o;
    o.eN[0] = a.x, o.eN[1] = a.y, o.eN[2] = a.z, o.eN[3] = a.w, o.eN[4] = b.x;
    o.eI[0] = b.y, o.eI[1] = b.z, o.eI[2] = b.w, o.eI[3] = c.x, o.eI[4] = c.y;
    o.code[0] = __float_as_uint(c.z), o.code[1] = __float_as_uint(c.w);
    o.code[2] = __float_as_uint(d.x), o.code[3] = __float_as_uint(d.y), o.code[4] = __float_as_uint(d.z);
    return o;
}

template <int Q>
__device__ __forceinline__ void load_emis(float (&em)[5][Q], const float *__restrict__ emis_lane,
                                          const uint32_t (&code)[5])
{
    constexpr int QP = Q <= 4 ? 4 : 8;
    constexpr int ROW = 32 * QP;
#pragma unroll
    for (int l = 0; l < 5; ++l)
    {
        const float4 *src = reinterpret_cast<const float4 *>(emis_lane + (size_t)code[l] * ROW);
        float4 a = __ldg(src);
        float tmp[8];
        tmp[0] = a.x, tmp[1] = a.y, tmp[2] = a.z, tmp[3] = a.w;
        if (Q > 4)
        {
            float4 b = __ldg(src + 1);
            tmp[4] = b.x, tmp[5] = b.y, tmp[6] = b.z, tmp[7] = b.w;
        }
#pragma unroll
        for (int i = 0; i < Q; ++i) em[l][i] = tmp[i];
    }
}


#endif
