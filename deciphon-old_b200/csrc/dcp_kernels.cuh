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

/* everything one DP row needs besides the match emissions */
struct RowIn
{
    float eI[5];        /* insert emission of seq[j-l:j] */
    float eN[5];        /* N/J/C emission of seq[j-l:j] */
    uint32_t code[5];   /* frame-table code of seq[j-l:j] */
    uint32_t next[5];   /* the same for row j+1 (0 past the end) */
};

__device__ __forceinline__ void unpack_codes(uint32_t (&c)[5], uint32_t w0, uint32_t w1, uint32_t w2)
{
    c[0] = w0 & 0xffffu, c[1] = w0 >> 16, c[2] = w1 & 0xffffu, c[3] = w1 >> 16, c[4] = w2 & 0xffffu;
}

/* first 32 bytes of a record: what every lane needs (eI of this row, codes of the next) */
__device__ __forceinline__ void load_row_common(const RowRec *__restrict__ r, float (&eI)[5], uint32_t (&next)[5])
{
    const float4 *q = reinterpret_cast<const float4 *>(r);
    float4 a = __ldg(q), b = __ldg(q + 1);
    eI[0] = a.x, eI[1] = a.y, eI[2] = a.z, eI[3] = a.w, eI[4] = b.x;
    unpack_codes(next, __float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(b.w));
}

/* second 32 bytes: eN (used by the lanes that carry N, J, C) and this row's own codes */
__device__ __forceinline__ void load_row_special(const RowRec *__restrict__ r, float (&eN)[5])
{
    const float4 *q = reinterpret_cast<const float4 *>(r) + 2;
    float4 a = __ldg(q);
    float e4 = __ldg(reinterpret_cast<const float *>(q + 1));
    eN[0] = a.x, eN[1] = a.y, eN[2] = a.z, eN[3] = a.w, eN[4] = e4;
}

__device__ __forceinline__ RowIn load_row(const RowRec *__restrict__ r)
{
    RowIn o;
    load_row_common(r, o.eI, o.next);
    load_row_special(r, o.eN);
    const float4 *q = reinterpret_cast<const float4 *>(r) + 3;
    float4 d = __ldg(q);
    unpack_codes(o.code, __float_as_uint(d.y), __float_as_uint(d.z), __float_as_uint(d.w));
    return o;
}

/*
 * Match emissions of one row: for each of the five lengths one line of the transposed table,
 * [code][half][lane][4] -- every LDG.128 of a warp covers 512 contiguous bytes.
 * emis_lane = table base + lane * 4.
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
        float4 a = __ldg(src);
        float tmp[8];
        tmp[0] = a.x, tmp[1] = a.y, tmp[2] = a.z, tmp[3] = a.w;
        if (Q > 4)
        {
            float4 b = __ldg(src + 32); /* second half: +128 floats */
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
