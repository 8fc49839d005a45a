// L2 / HBM bandwidth probe (B200): how fast can all SMs re-read a buffer that fits in L2, write one, and do both?
// Decides whether a two-pass CEM (second read of y from L2) can beat the HBM roofline.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/l2bw tools/l2bw.cu && tools/bin/l2bw
#include <cstdio>
#include <cuda_runtime.h>

__global__ void rd(const float4* __restrict__ p, size_t n, float* sink) {
    float4 acc = make_float4(0, 0, 0, 0);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p + i));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (acc.x + acc.y + acc.z + acc.w == 12345.678f) *sink = acc.x;
}
__global__ void wr(float4* __restrict__ p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}
__global__ void cp(const float4* __restrict__ a, float4* __restrict__ b, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
int main() {
    const size_t maxb = 512ull << 20;
    float4 *a, *b; float* sink;
    cudaMalloc(&a, maxb); cudaMalloc(&b, maxb); cudaMalloc(&sink, 4);
    cudaMemset(a, 0, maxb); cudaMemset(b, 0, maxb);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8, blk = 512;
    printf("size_MB  read_GB/s  write_GB/s  copy_GB/s(r+w)   [each: 1 warm pass, then 10 timed passes over the same buffer]\n");
    for (size_t mb : {8, 16, 32, 50, 64, 96, 128, 192, 256, 512}) {
        const size_t n = (mb << 20) / 16;
        float ms[3];
        for (int k = 0; k < 3; ++k) {
            auto run = [&]() { if (k == 0) rd<<<grid, blk>>>(a, n, sink); else if (k == 1) wr<<<grid, blk>>>(b, n); else cp<<<grid, blk>>>(a, b, n / 2); };
            run(); run();
            cudaEventRecord(e0);
            for (int r = 0; r < 10; ++r) run();
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms[k], e0, e1);
        }
        const double gb = (double)(mb << 20) / 1e9;
        printf("%6zu  %9.0f  %9.0f  %9.0f\n", mb, gb * 10 / (ms[0] * 1e-3), gb * 10 / (ms[1] * 1e-3), gb * 10 / (ms[2] * 1e-3));
    }
    // the CEM pattern: stream-read 50 MB cold (flushed), then re-read it (L2) while writing 50 MB
    {
        const size_t n = (50ull << 20) / 16;
        float t1 = 0, t2 = 0, t3 = 0;
        for (int r = 0; r < 5; ++r) {
            wr<<<grid, blk>>>(b + (256ull << 20) / 16, (256ull << 20) / 16);      // flush L2
            cudaEventRecord(e0); rd<<<grid, blk>>>(a, n, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float m; cudaEventElapsedTime(&m, e0, e1); t1 += m;
            cudaEventRecord(e0); cp<<<grid, blk>>>(a, b, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&m, e0, e1); t2 += m;
            wr<<<grid, blk>>>(b + (256ull << 20) / 16, (256ull << 20) / 16);
            cudaEventRecord(e0); cp<<<grid, blk>>>(a, b, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&m, e0, e1); t3 += m;
        }
        printf("CEM pattern, 50 MB: cold read %.1f us; then copy with the source in L2 %.1f us; cold copy %.1f us\n", t1 / 5 * 1e3, t2 / 5 * 1e3, t3 / 5 * 1e3);
    }
    return 0;
}
