/*
 * dcp_microbench.cu -- measured FP32 issue-rate peaks for the roofline of k_score.
 *
 * The DP cell is 18 FADD + 15 two-input max (9 FMNMX3 on sm_100a) with no multiply, so neither
 * the HBM copy bandwidth nor the tensor-core GEMM rate in MEASURED_PEAKS.json bounds it.  These
 * kernels measure, on the device the scan runs on, the rate at which independent FADD, FMNMX3
 * and the cell's own 18:9 FADD:FMNMX3 mix issue from registers.
 */
#include "dcp_engine.h"

namespace
{
constexpr int kChains = 9;   /* independent dependency chains per thread */
constexpr int kInner = 64;

/* mode 0: FADD only, 1: FMNMX3 only, 2: 18 FADD + 9 FMNMX3 per group (the DP cell's mix) */
/* clk[0..1]: SM cycles and nanoseconds thread 0 of block 0 spent in the loop: the SM clock the kernel really ran at */
template <int MODE>
__global__ void __launch_bounds__(256) k_alu(float *out, float seed, int iters, unsigned long long *clk)
{
    unsigned long long c0 = 0, t0 = 0;
    if (clk && blockIdx.x == 0 && threadIdx.x == 0)
    {
        c0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    }
    float a[kChains], b[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) a[c] = seed + threadIdx.x + c, b[c] = seed - c;
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int k = 0; k < kInner; ++k)
        {
#pragma unroll
            for (int c = 0; c < kChains; ++c)
            {
                if (MODE == 0)
                {
                    a[c] = a[c] + b[c];
                    b[c] = b[c] + seed;
                    a[c] = a[c] + seed;
                }
                else if (MODE == 1)
                {
                    a[c] = fmaxf(fmaxf(a[c], b[c]), seed);
                    b[c] = fmaxf(fmaxf(b[c], a[c]), -seed);
                    a[c] = fmaxf(fmaxf(a[c], seed), b[c]);
                }
                else
                {
                    float x = a[c] + b[c];
                    float y = b[c] + seed;
                    a[c] = fmaxf(fmaxf(x, y), a[c]);
                    b[c] = y;
                }
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += a[c] + b[c];
    if (s == 12345.678f) out[threadIdx.x] = s;
    if (clk && blockIdx.x == 0 && threadIdx.x == 0)
    {
        unsigned long long c1 = clock64(), t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        clk[0] = c1 - c0, clk[1] = t1 - t0;
    }
}

template <int MODE>
float run(cudaStream_t st, int sms, float *d_out, int iters, unsigned long long *d_clk)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    k_alu<MODE><<<sms * 8, 256, 0, st>>>(d_out, 1.0f, 4, nullptr); /* warm-up */
    cudaEventRecord(e0, st);
    k_alu<MODE><<<sms * 8, 256, 0, st>>>(d_out, 1.0f, iters, d_clk);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0), cudaEventDestroy(e1);
    return ms;
}
} // namespace

/* out[0..2] = warp-instruction issue rate in 1e9 lane-instructions/s for FADD, FMNMX3 and the
 * 2:1 FADD:FMNMX3 mix of the DP cell; out[3] = SM count; out[4] = SM clock in MHz measured inside the mix kernel
 * (clock64 against %globaltimer) -- out[2] / (out[3] * out[4] * 1e-3) is lane-instructions per SM and clock, to be
 * read against the 128 FP32 lanes of an SM; out[5] = the same clock for the FADD-only kernel. */
extern "C" enum rc dcpgpu_microbench_alu(int device, double out[6])
{
    CU_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    int sms = prop.multiProcessorCount;
    float *d_out = nullptr;
    unsigned long long *d_clk = nullptr, h_clk[4] = {0, 1, 0, 1};
    CU_TRY(cudaMalloc(&d_out, 256 * sizeof(float)));
    CU_TRY(cudaMalloc(&d_clk, 4 * sizeof(unsigned long long)));
    const int iters = 2000;
    const double threads = (double)sms * 8 * 256;
    float ms0 = run<0>(0, sms, d_out, iters, d_clk + 2);
    float ms1 = run<1>(0, sms, d_out, iters, nullptr);
    float ms2 = run<2>(0, sms, d_out, iters, d_clk);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpy(h_clk, d_clk, sizeof h_clk, cudaMemcpyDeviceToHost));
    cudaFree(d_out), cudaFree(d_clk);
    out[4] = h_clk[1] ? (double)h_clk[0] / (double)h_clk[1] * 1e3 : 0.0;
    out[5] = h_clk[3] ? (double)h_clk[2] / (double)h_clk[3] * 1e3 : 0.0;
    const double per_thread3 = (double)iters * kInner * kChains * 3.0;
    out[0] = threads * per_thread3 / (ms0 * 1e-3) / 1e9;
    out[1] = threads * per_thread3 / (ms1 * 1e-3) / 1e9;
    out[2] = threads * per_thread3 / (ms2 * 1e-3) / 1e9; /* 2 FADD + 1 FMNMX3 per chain step */
    out[3] = sms;
    return RC_OK;
}
