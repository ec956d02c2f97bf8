// How do DMMA (mma.sync m8n8k4 f64) warps and scalar-FP64 warps share an SM sub-partition on B200?
// One CTA of 16 warps per SM; warp w sits on sub-partition w % 4.  Each warp is given a role by a 16-entry table:
//   'M' = DMMA stream (18 independent accumulators), 'F' = DFMA stream with ILP chains, '.' = idle.
// Every role runs a fixed amount of work; the kernel reports per-role clocks (max over warps of that role), so the
// slowdown of either stream under mixing is visible.  Diagnostic only (design input for the fused kernel).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/fp64_mix.cu -o tools/fp64_mix
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct Roles { char r[16]; int ilp; };

template <int ILP>
__device__ __forceinline__ double dfma_stream(int iters, double a, double b, double seed) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = seed + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    return s;
}

__global__ void __launch_bounds__(512, 1) mix_kernel(Roles roles, int mma_iters, int fma_instr, double a, double b,
                                                      double* out, long long* clocks) {
    const int warp = threadIdx.x >> 5;
    const char role = roles.r[warp];
    __shared__ long long s_clk[16];
    __syncthreads();
    const long long t0 = clock64();
    double s = 0;
    if (role == 'M') {
        double c[18][2];
#pragma unroll
        for (int i = 0; i < 18; ++i) { c[i][0] = 0; c[i][1] = 0; }
        const double fa = threadIdx.x * 1e-6, fb = 1.0 + threadIdx.x * 1e-7;
        for (int it = 0; it < mma_iters; ++it) {
#pragma unroll
            for (int i = 0; i < 18; ++i) dmma884(c[i][0], c[i][1], fa, fb);
        }
#pragma unroll
        for (int i = 0; i < 18; ++i) s += c[i][0] + c[i][1];
    } else if (role == 'F') {
        if (roles.ilp == 1) s = dfma_stream<1>(fma_instr, a, b, threadIdx.x);
        else if (roles.ilp == 2) s = dfma_stream<2>(fma_instr / 2, a, b, threadIdx.x);
        else if (roles.ilp == 4) s = dfma_stream<4>(fma_instr / 4, a, b, threadIdx.x);
        else s = dfma_stream<8>(fma_instr / 8, a, b, threadIdx.x);
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) s_clk[warp] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < 16) clocks[blockIdx.x * 16 + threadIdx.x] = s_clk[threadIdx.x];
}

int main(int argc, char** argv) {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* out; long long* clocks;
    CK(cudaMalloc(&out, sizeof(double) * sms * 512));
    CK(cudaMalloc(&clocks, sizeof(long long) * sms * 16));
    const int mma_iters = 2000;          // 18 DMMA each
    const int fma_instr = 16000;
    const char* cfgs[] = {
        "MMMMMMMMMMMM....", "MMMMMMMM........", "MMMM............", "............FFFF", "........FFFFFFFF",
        "MMMMMMMMMMMMFFFF", "MMMMMMMMFFFFFFFF", "MMMMFFFFFFFFFFFF",
        // sub-partition-partitioned: F warps only on sub-partition 3 (warps 3, 7, 11, 15)
        "MMMFMMMFMMMFMMMF", "MMM.MMM.MMM.MMM.", "...F...F...F...F",
        "MMMFMMMFMMM.MMM.", "MMMFMMM.MMM.MMM.",
    };
    printf("{\"mma_dmma_per_warp\": %d, \"fma_instr_per_warp\": %d, \"runs\": [\n", mma_iters * 18, fma_instr);
    bool first = true;
    for (const char* cfg : cfgs) {
        for (int ilp : {1, 2, 4, 8}) {
            if (!strchr(cfg, 'F') && ilp != 1) continue;
            Roles r; memcpy(r.r, cfg, 16); r.ilp = ilp;
            for (int rep = 0; rep < 2; ++rep) {
                mix_kernel<<<sms, 512>>>(r, mma_iters, fma_instr, 1.0000001, 1e-9, out, clocks);
                CK(cudaDeviceSynchronize());
            }
            static long long h[16 * 256];
            CK(cudaMemcpy(h, clocks, sizeof(long long) * sms * 16, cudaMemcpyDeviceToHost));
            long long mM = 0, mF = 0;
            for (int w = 0; w < 16; ++w) { if (cfg[w] == 'M' && h[w] > mM) mM = h[w]; if (cfg[w] == 'F' && h[w] > mF) mF = h[w]; }
            printf("%s  {\"cfg\": \"%s\", \"ilp\": %d, \"mma_clk\": %lld, \"clk_per_dmma_per_warp\": %.2f, \"fma_clk\": %lld, \"clk_per_dfma_per_warp\": %.2f}",
                   first ? "" : ",\n", cfg, ilp, mM, mM ? (double)mM / (mma_iters * 18) : 0.0, mF, mF ? (double)mF / fma_instr : 0.0);
            first = false;
        }
    }
    printf("\n]}\n");
    return 0;
}
