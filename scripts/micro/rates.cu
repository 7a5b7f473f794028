// rates.cu -- hardware rates that decide the layout of the inner-loop kernel (round 2):
//   (1) FP64 pipes: DFMA, DMMA.8x8x4, F2F.F64.F32 issue rates per SM
//   (2) grid-wide barrier cost: cooperative-groups grid.sync() against a counter + flag barrier, with and
//       without the per-iteration "every CTA re-reads G x 28 partial sums" step
//   (3) streaming reads of an L2-resident / DRAM-resident buffer by a persistent grid
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o rates.bin rates.cu ; run on a B200.
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

// ---- (1) FP64 pipes ------------------------------------------------------------------------------
template <int NACC>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = threadIdx.x + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) acc[k] = fma(a, acc[k], b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < NACC; ++k) s += acc[k];
    if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void dmma_kernel(double* out, int iters, double a, double b) {
    double c0[NACC], c1[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) { c0[k] = threadIdx.x; c1[k] = k; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < NACC; ++k)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[k]), "+d"(c1[k]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < NACC; ++k) s += c0[k] + c1[k];
    if (s == 123.456) out[0] = s;
}

__global__ void f2f_kernel(double* out, int iters, float seed) {
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = seed + threadIdx.x + k;
    double s = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            double d;
            asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(f[k]));
            long long bits = __double_as_longlong(d);
            f[k] = __int_as_float((int)(bits >> 32) ^ i);     // keeps the conversion alive, integer pipe
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) s += f[k];
    if (s == 123.456) out[0] = s;
}

// ---- (2) barriers ---------------------------------------------------------------------------------
__global__ void cg_barrier_kernel(int iters, double* part, double* sink, int reread) {
    cg::grid_group grid = cg::this_grid();
    double acc = 0;
    for (int i = 0; i < iters; ++i) {
        if (reread && threadIdx.x < 28) part[(size_t)blockIdx.x * 28 + threadIdx.x] = i + threadIdx.x;
        grid.sync();
        if (reread && threadIdx.x < 28) {
            double s = 0;
            for (unsigned g = 0; g < gridDim.x; ++g) s += __ldcg(part + (size_t)g * 28 + threadIdx.x);
            acc += s;
        }
    }
    if (acc == 123.456) sink[0] = acc;
}

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// counter barrier: every CTA adds 1 (release) after a __syncthreads, one thread spins (acquire) until the
// counter reaches (i + 1) * G, then a __syncthreads releases the CTA.  Monotonic counter, no reset.
// reread = 1: every CTA then sums the G x 28 partials in order with one lane per value (the latency chain of the
// real kernel); reread = 2: the partials are summed in chunks by all warps, then the chunk sums by one warp.
__global__ void ctr_barrier_kernel(int iters, unsigned* counter, double* part, double* sink, int reread) {
    __shared__ double s_chunk[32][28];
    double acc = 0;
    const unsigned G = gridDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int i = 0; i < iters; ++i) {
        // two alternating partial buffers: a CTA may run ahead by one iteration
        double* pbuf = part + (size_t)(i & 1) * G * 28;
        if (reread && threadIdx.x < 28) __stcg(pbuf + (size_t)blockIdx.x * 28 + threadIdx.x, (double)(i + threadIdx.x));
        __syncthreads();
        if (threadIdx.x == 0) {
            red_release(counter, 1u);
            const unsigned want = (unsigned)(i + 1) * G;
            while (ld_acquire(counter) < want) { }
        }
        __syncthreads();
        if (reread == 1) {
            if (threadIdx.x < 28) {
                double s = 0;
                for (unsigned g = 0; g < G; ++g) s += __ldcg(pbuf + (size_t)g * 28 + threadIdx.x);
                acc += s;
            }
        } else if (reread == 2) {
            const unsigned per = (G + nwarps - 1) / nwarps;
            if (lane < 28) {
                double s = 0;
                const unsigned g0 = warp * per, g1 = min(G, g0 + per);
                for (unsigned g = g0; g < g1; ++g) s += __ldcg(pbuf + (size_t)g * 28 + lane);
                s_chunk[warp][lane] = s;
            }
            __syncthreads();
            if (threadIdx.x < 28) {
                double s = 0;
                for (int w = 0; w < nwarps; ++w) s += s_chunk[w][threadIdx.x];
                acc += s;
            }
        }
    }
    if (acc == 123.456) sink[0] = acc;
}

// ---- (3) streaming reads ---------------------------------------------------------------------------
__global__ void stream_kernel(const float4* __restrict__ p, size_t n, int reps, float* sink) {
    float s = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride * 4) {
            float4 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { const size_t j = i + k * stride; v[k] = (j < n) ? __ldcg(p + j) : make_float4(0, 0, 0, 0); }
#pragma unroll
            for (int k = 0; k < 4; ++k) s += v[k].x + v[k].y + v[k].z + v[k].w;
        }
    }
    if (s == 123.456f) sink[0] = s;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s, %d SMs, nominal %d MHz\n", prop.name, sms, clk_khz / 1000);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    double* dout; CK(cudaMalloc(&dout, 1 << 20));
    unsigned* dctr; CK(cudaMalloc(&dctr, 256));

    // (1) FP64 rates: 4 CTAs x 256 threads per SM would exceed registers for large NACC; 2 x 256 is enough to fill the pipe
    {
        const int iters = 4096;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            dfma_kernel<28><<<sms * 2, 256>>>(dout, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            double ms = time_ms(e0, e1);
            double winstr = (double)sms * 2 * 8 * iters * 28;
            if (rep) printf("DFMA  : %.3f ms, %.2f warp-instr/us/SM = %.3f warp-DFMA/clk/SM @1.9GHz, %.1f TFLOP/s\n", ms,
                            winstr / sms / (ms * 1e3), winstr / sms / (ms * 1e-3) / 1.9e9, winstr * 64 / (ms * 1e-3) / 1e12);
            cudaEventRecord(e0);
            dmma_kernel<8><<<sms * 2, 256>>>(dout, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            ms = time_ms(e0, e1);
            winstr = (double)sms * 2 * 8 * iters * 8;
            if (rep) printf("DMMA  : %.3f ms, %.2f warp-instr/us/SM = %.3f DMMA.8x8x4/clk/SM @1.9GHz, %.1f TFLOP/s\n", ms,
                            winstr / sms / (ms * 1e3), winstr / sms / (ms * 1e-3) / 1.9e9, winstr * 512 / (ms * 1e-3) / 1e12);
            cudaEventRecord(e0);
            f2f_kernel<<<sms * 2, 256>>>(dout, iters, 1.5f);
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            ms = time_ms(e0, e1);
            winstr = (double)sms * 2 * 8 * iters * 8;
            if (rep) printf("F2F64 : %.3f ms, %.3f warp-cvt/clk/SM @1.9GHz (loop also carries 2 integer ops per cvt)\n", ms,
                            winstr / sms / (ms * 1e-3) / 1.9e9);
        }
    }

    // (2) barriers
    {
        double* part; CK(cudaMalloc(&part, (size_t)4096 * 28 * 8 * 2));
        const int iters = 2000;
        struct Geo { int ctas_per_sm, threads; } geos[] = {{1, 512}, {1, 256}, {1, 1024}, {3, 256}};
        for (Geo g : geos) {
            const int grid = sms * g.ctas_per_sm;
            for (int reread = 0; reread <= 2; ++reread) {
                float ms_cg = -1.f;
                if (reread < 2) {
                    int it = iters; void* args[] = {&it, &part, &dout, &reread};
                    for (int rep = 0; rep < 2; ++rep) {
                        cudaEventRecord(e0);
                        CK(cudaLaunchCooperativeKernel((void*)cg_barrier_kernel, dim3(grid), dim3(g.threads), args, 0, 0));
                        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
                        ms_cg = time_ms(e0, e1);
                    }
                }
                float ms_ctr = 0.f;
                for (int rep = 0; rep < 2; ++rep) {
                    CK(cudaMemset(dctr, 0, 256));
                    int it = iters; void* args[] = {&it, &dctr, &part, &dout, &reread};
                    cudaEventRecord(e0);
                    CK(cudaLaunchCooperativeKernel((void*)ctr_barrier_kernel, dim3(grid), dim3(g.threads), args, 0, 0));
                    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
                    ms_ctr = time_ms(e0, e1);
                }
                printf("barrier grid=%4d x %4d thr, reread=%d: grid.sync %.2f us/iter, counter barrier %.2f us/iter\n", grid,
                       g.threads, reread, ms_cg * 1e3 / iters, ms_ctr * 1e3 / iters);
            }
        }
    }

    // (3) streaming
    {
        const size_t sizes_mb[] = {16, 32, 48, 64, 96, 256, 1024};
        float4* buf; CK(cudaMalloc(&buf, (size_t)1024 << 20));
        CK(cudaMemset(buf, 0, (size_t)1024 << 20));
        for (size_t mb : sizes_mb) {
            const size_t n = (mb << 20) / 16;
            const int reps = (int)(4096 / mb) + 1;
            for (int geo = 0; geo < 2; ++geo) {
                const int grid = sms * (geo ? 4 : 2), thr = geo ? 256 : 512;
                float ms = 0;
                for (int rep = 0; rep < 2; ++rep) {
                    cudaEventRecord(e0);
                    stream_kernel<<<grid, thr>>>(buf, n, reps, (float*)dout);
                    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
                    ms = time_ms(e0, e1);
                }
                printf("stream %5zu MB x %3d reps, grid %d x %d: %.3f ms -> %.0f GB/s\n", mb, reps, grid, thr, ms,
                       (double)(mb << 20) * reps / (ms * 1e-3) / 1e9);
            }
        }
    }
    return 0;
}
