// Latency microbenchmark: dependent chains of warp-min implementations (development aid).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(uint32_t* out, long long* cyc, int iters)
{
    uint32_t v = threadIdx.x * 2654435761u + 12345u;
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        uint32_t m;
        if (MODE == 0) m = __reduce_min_sync(0xffffffffu, v);
        else if (MODE == 1) {
            m = v;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
        } else if (MODE == 2) {  // packed 16x2 butterfly
            m = v;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = __vminu2(m, __shfl_xor_sync(0xffffffffu, m, o));
        } else {                  // plain dependent ALU op for reference
            m = __vminu2(v, acc + 0x00010001u);
        }
        acc += m;
        v = v * 1664525u + m;     // next input depends on the result
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[blockIdx.x] = t1 - t0; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main()
{
    uint32_t* d; long long* c;
    cudaMalloc(&d, 1 << 20); cudaMalloc(&c, 8 * 1024);
    const int iters = 4096;
    for (int warps = 1; warps <= 16; warps *= 4) {
        long long h[4];
        k<0><<<1, 32 * warps>>>(d, c, iters); cudaMemcpy(&h[0], c, 8, cudaMemcpyDeviceToHost);
        k<1><<<1, 32 * warps>>>(d, c, iters); cudaMemcpy(&h[1], c, 8, cudaMemcpyDeviceToHost);
        k<2><<<1, 32 * warps>>>(d, c, iters); cudaMemcpy(&h[2], c, 8, cudaMemcpyDeviceToHost);
        k<3><<<1, 32 * warps>>>(d, c, iters); cudaMemcpy(&h[3], c, 8, cudaMemcpyDeviceToHost);
        printf("warps/CTA %2d: cycles per dependent iteration: redux %.1f  shfl-butterfly32 %.1f  shfl-butterfly16x2 %.1f  alu-only %.1f\n",
               warps, double(h[0]) / iters, double(h[1]) / iters, double(h[2]) / iters, double(h[3]) / iters);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
