// fp64 peak microbenchmark for B200 (sm_100a): the roofline denominator for the fused
// regressor+Gram kernel is not in MEASURED_PEAKS.json (bf16 only), so measure it here:
//   (1) DFMA register-resident FMA chains,
//   (2) DMMA mma.sync m8n8k4 / m16n8k8 / m16n8k16 f64 with independent accumulators,
//   (3) cuBLAS DGEMM 8192^3 (burst = best of 10, sustained = back-to-back for ~3 s).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo tools/fp64_peak.cu -lcublas -o tools/fp64_peak
// Prints one JSON object on stdout.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

template <int CHAINS>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double acc[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NACC>
__global__ void __launch_bounds__(256) dmma884_kernel(double* out, int iters) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0; c[i][1] = 0; }
    double a = threadIdx.x * 1e-6, b = 1.0 + threadIdx.x * 1e-7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void __launch_bounds__(256) dmma1688_kernel(double* out, int iters) {
    double c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0;
    double a[4], b[2];
    for (int j = 0; j < 4; ++j) a[j] = threadIdx.x * 1e-6 + j;
    for (int j = 0; j < 2; ++j) b[j] = 1.0 + threadIdx.x * 1e-7 + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma1688(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void __launch_bounds__(256) dmma16816_kernel(double* out, int iters) {
    double c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0;
    double a[8], b[4];
    for (int j = 0; j < 8; ++j) a[j] = threadIdx.x * 1e-6 + j;
    for (int j = 0; j < 4; ++j) b[j] = 1.0 + threadIdx.x * 1e-7 + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma16816(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main(int argc, char** argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
    int sms = prop.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 256 * 4));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", prop.name, sms, prop.clockRate);

    // (1) DFMA
    {
        const int iters = 20000;
        for (int bps : {1, 2, 4, 8}) {
            int blocks = sms * bps;
            double ms = time_ms([&] { dfma_kernel<8><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double flops = 2.0 * 8 * iters * 256.0 * blocks;
            printf(", \"dfma_tflops_bps%d\": %.3f", bps, flops / ms * 1e-9);
        }
    }
    // (2) DMMA variants: blocks-per-SM sweep at 256 threads (8 warps)
    {
        const int iters = 4000;
        for (int bps : {1, 2, 4}) {
            int blocks = sms * bps;
            double ms = time_ms([&] { dmma884_kernel<16><<<blocks, 256>>>(out, iters); }, 5);
            double flops = 2.0 * 8 * 8 * 4 * 16.0 * iters * 8 * blocks;
            printf(", \"dmma_m8n8k4_tflops_bps%d\": %.3f", bps, flops / ms * 1e-9);
            ms = time_ms([&] { dmma1688_kernel<8><<<blocks, 256>>>(out, iters); }, 5);
            flops = 2.0 * 16 * 8 * 8 * 8.0 * iters * 8 * blocks;
            printf(", \"dmma_m16n8k8_tflops_bps%d\": %.3f", bps, flops / ms * 1e-9);
            ms = time_ms([&] { dmma16816_kernel<8><<<blocks, 256>>>(out, iters); }, 5);
            flops = 2.0 * 16 * 8 * 16 * 8.0 * iters * 8 * blocks;
            printf(", \"dmma_m16n8k16_tflops_bps%d\": %.3f", bps, flops / ms * 1e-9);
        }
        // single warp per SMSP (4 warps / SM) to see if one warp saturates the pipe
        {
            int blocks = sms;
            double ms = time_ms([&] { dmma884_kernel<16><<<blocks, 128>>>(out, iters); }, 5);
            double flops = 2.0 * 8 * 8 * 4 * 16.0 * iters * 4 * blocks;
            printf(", \"dmma_m8n8k4_tflops_4warps\": %.3f", flops / ms * 1e-9);
            ms = time_ms([&] { dmma884_kernel<4><<<blocks, 128>>>(out, iters); }, 5);
            flops = 2.0 * 8 * 8 * 4 * 4.0 * iters * 4 * blocks;
            printf(", \"dmma_m8n8k4_tflops_4warps_4acc\": %.3f", flops / ms * 1e-9);
        }
    }
    // (3) cuBLAS DGEMM
    {
        const int n = 8192;
        double *A, *B, *C;
        CK(cudaMalloc(&A, sizeof(double) * n * n)); CK(cudaMalloc(&B, sizeof(double) * n * n)); CK(cudaMalloc(&C, sizeof(double) * n * n));
        CK(cudaMemset(A, 0, sizeof(double) * n * n)); CK(cudaMemset(B, 0, sizeof(double) * n * n));
        cublasHandle_t h; cublasCreate(&h);
        double one = 1.0, zero = 0.0;
        auto gemm = [&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); };
        double ms = time_ms(gemm, 10);
        double flops = 2.0 * n * (double)n * n;
        printf(", \"dgemm8192_burst_tflops\": %.3f", flops / ms * 1e-9);
        // sustained: ~3 s back to back
        int reps = (int)(3000.0 / ms) + 1;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; ++i) gemm();
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float tot; CK(cudaEventElapsedTime(&tot, e0, e1));
        printf(", \"dgemm8192_sustained_tflops\": %.3f", flops * reps / tot * 1e-9);
        // SYRK-shaped: C(160x160) += A^T A with K = 18*65536 (tall-skinny, the shape of our contraction)
        cublasDestroy(h);
    }
    printf("}\n");
    return 0;
}
