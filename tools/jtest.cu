// diagnostic: jacobi_min_eig<6> on the device vs the host on the same matrix (the inlined, register-promoted form of this
// routine was miscompiled by nvcc 12.9 -O3 for sm_100a: device -0.480293150 vs host 0.487109436; see tsqr_kernels.cuh)
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/jtest tools/jtest.cu
#include <cstdio>
#include <cmath>
#include "../system_identification_b200/csrc/tsqr_kernels.cuh"
using namespace sysid;
__host__ __device__ void fill(double (&a)[6][6]) {
    double m = 17.7, h[3] = {0.3, -0.1, 2.0};
    double Ib[3][3] = {{1.2, 0.01, 0.02}, {0.01, 1.1, 0.03}, {0.02, 0.03, 0.5}};
    for (int i = 0; i < 6; ++i) for (int k = 0; k < 6; ++k) a[i][k] = 0.0;
    for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) a[i][k] = Ib[i][k];
    a[0][4] = -h[2]; a[0][5] = h[1]; a[1][3] = h[2]; a[1][5] = -h[0]; a[2][3] = -h[1]; a[2][4] = h[0];
    for (int i = 0; i < 3; ++i) for (int k = 3; k < 6; ++k) a[k][i] = a[i][k];
    for (int i = 0; i < 3; ++i) a[3 + i][3 + i] = m;
}

__global__ void k(double* out) {
    double a[6][6];
    fill(a);
    out[0] = jacobi_min_eig<6>(a, 6);
    for (int i = 0; i < 6; ++i) out[1 + i] = a[i][i];
    double b[4][4] = {{2, 1, 0, 0}, {1, 2, 1, 0}, {0, 1, 2, 1}, {0, 0, 1, 2}};
    out[7] = jacobi_min_eig<4>(b, 4);
}
int main() {
    double a[6][6]; fill(a);
    printf("host   %.9f\n", jacobi_min_eig<6>(a, 6));
    double* d; cudaMalloc(&d, 64); k<<<1, 1>>>(d); double h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("device %.9f  diag %.6f %.6f %.6f %.6f %.6f %.6f  4x4 %.9f (0.381966011)\n", h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
}
