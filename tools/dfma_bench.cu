// FP64 FMA throughput of the device (dependent chains of DFMA, 8 per thread): sizes the matrix-free operator
// application against the fp64 pipe.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dfma_bench.bin tools/dfma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, int n, double a) {
    double x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < n; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, 1e-9);
    double s = 0;
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    double* d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(double));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int n = 20000;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<<<148 * 8, 256>>>(d, n, 0.999999);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 148 * 8 * 256 * 8.0 * n;
        printf("dfma: %.3f ms, %.2f TFLOP/s fp64\n", ms, fl / ms * 1e-9);
    }
    return 0;
}
