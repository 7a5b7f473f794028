// Does mma.sync.m8n8k4.f64 accumulate like a chain of DFMAs in k order?  (The batch sums of the
// inner ICP loop have a fixed summation order that the CPU oracle reproduces bit for bit.)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cuda_runtime.h>

__global__ void k(const float* A, const float* B, double* D, int nchunks) {
    // A: [nchunks*4][8] (row-major: batch row, m), B: [nchunks*4][8] (batch row, n)
    const int lane = threadIdx.x;
    double c0 = 0.0, c1 = 0.0;
    for (int j = 0; j < nchunks; ++j) {
        const double a = (double)A[(j * 4 + lane % 4) * 8 + lane / 4];
        const double b = (double)B[(j * 4 + lane % 4) * 8 + lane / 4];
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    }
    D[(lane / 4) * 8 + (lane % 4) * 2] = c0;
    D[(lane / 4) * 8 + (lane % 4) * 2 + 1] = c1;
}

int main() {
    const int nchunks = 8, rows = nchunks * 4;
    float hA[rows * 8], hB[rows * 8];
    double hD[64];
    float *dA, *dB; double* dD;
    cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD));
    long bad_seq = 0, bad_pair = 0, total = 0;
    srand(1);
    for (int trial = 0; trial < 2000; ++trial) {
        for (int i = 0; i < rows * 8; ++i) {
            hA[i] = (float)((rand() / (double)RAND_MAX - 0.5) * pow(10.0, (rand() % 7) - 3));
            hB[i] = (float)((rand() / (double)RAND_MAX - 0.5) * pow(10.0, (rand() % 7) - 3));
        }
        cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
        cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
        k<<<1, 32>>>(dA, dB, dD, nchunks);
        cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
        for (int m = 0; m < 8; ++m)
            for (int n = 0; n < 8; ++n) {
                double seq = 0.0;
                for (int r = 0; r < rows; ++r) seq = fma((double)hA[r * 8 + m], (double)hB[r * 8 + n], seq);
                double pr = 0.0;   // hypothesis 2: the four products of a chunk summed pairwise, then added
                for (int j = 0; j < nchunks; ++j) {
                    double p[4];
                    for (int q = 0; q < 4; ++q) p[q] = (double)hA[(j * 4 + q) * 8 + m] * (double)hB[(j * 4 + q) * 8 + n];
                    pr = pr + ((p[0] + p[1]) + (p[2] + p[3]));
                }
                ++total;
                if (memcmp(&seq, &hD[m * 8 + n], 8)) ++bad_seq;
                if (memcmp(&pr, &hD[m * 8 + n], 8)) ++bad_pair;
            }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("dmma_order: %s; %ld results, mismatches vs sequential-FMA chain: %ld, vs pairwise: %ld\n",
           cudaGetErrorString(e), total, bad_seq, bad_pair);
    return 0;
}
