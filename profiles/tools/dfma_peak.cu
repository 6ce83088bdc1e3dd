// fp64 peak of the device, measured: independent DFMA chains in registers, no memory traffic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dfma_peak profiles/tools/dfma_peak.cu
//   /tmp/dfma_peak > gpurun_out/fp64_peak.json
// Prints one JSON line: DFMA thread-instructions per second (whole device), per SM per clock at
// the SM clock read around the run, and the same for a DADD/DMUL/DFMA mix.  The roofline line of
// bench.py (`roofline.fp64_frac`) divides the stage's fp64 thread-instructions per second by
// `dfma_per_s` ("builder must run a DFMA micro-benchmark").
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int CHAINS, bool MIX>
__global__ void __launch_bounds__(256) dfma_kernel(double *out, double a, double b, int iters)
{
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) x[c] = a + 1e-3 * (threadIdx.x + c);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            if (MIX && (c % 3) == 1) x[c] = x[c] * b;            // DMUL
            else if (MIX && (c % 3) == 2) x[c] = x[c] + a;        // DADD
            else x[c] = fma(x[c], b, a);                          // DFMA
        }
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += x[c];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;    // never true: keeps the chains alive
}

static int sm_clock_mhz()
{
    FILE *p = popen("nvidia-smi --id=0 --query-gpu=clocks.sm --format=csv,noheader,nounits", "r");
    if (!p) return 0;
    int mhz = 0;
    if (fscanf(p, "%d", &mhz) != 1) mhz = 0;
    pclose(p);
    return mhz;
}

template <int CHAINS, bool MIX>
static double run(int sms, int ctas_per_sm, int iters, double *out, int *mhz)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = sms * ctas_per_sm;
    dfma_kernel<CHAINS, MIX><<<grid, 256>>>(out, 1.0000001, 0.9999999, iters);      // warm-up
    cudaDeviceSynchronize();
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        dfma_kernel<CHAINS, MIX><<<grid, 256>>>(out, 1.0000001, 0.9999999, iters);
        if (rep == 2) *mhz = sm_clock_mhz();       // sampled while the kernel runs
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double rate = (double)grid * 256 * CHAINS * (double)iters / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    return best;
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out = nullptr;
    cudaMalloc(&out, sizeof(double) * 256 * sms * 8);
    const int iters = 1 << 17;        // ~0.1-0.3 s per launch: long enough for a clock sample
    int mhz8 = 0, mhz16 = 0, mhzm = 0;
    const double r8 = run<8, false>(sms, 8, iters, out, &mhz8);
    const double r16 = run<16, false>(sms, 4, iters, out, &mhz16);
    const double rm = run<12, true>(sms, 6, iters, out, &mhzm);
    const double best = r8 > r16 ? r8 : r16;
    const int mhz = r8 > r16 ? mhz8 : mhz16;
    printf("{\"sms\": %d, \"dfma_per_s\": %.6e, \"fp64_tflops\": %.3f, \"sm_mhz_during\": %d, "
           "\"dfma_per_clk_per_sm\": %.2f, \"mix_per_s\": %.6e, \"sm_mhz_during_mix\": %d, "
           "\"chains8_per_s\": %.6e, \"chains16_per_s\": %.6e, "
           "\"how\": \"independent DFMA chains in registers, 256-thread CTAs, best of 5 launches, CUDA events\"}\n",
           sms, best, 2.0 * best / 1e12, mhz, mhz > 0 ? best / (sms * (double)mhz * 1e6) : 0.0, rm, mhzm, r8, r16);
    cudaFree(out);
    return 0;
}
