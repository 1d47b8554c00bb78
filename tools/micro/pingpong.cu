// Inter-SM message latency through L2: CTA 0 and CTA k bounce a counter with 8-byte {data, tag} records
// (st.relaxed.gpu / ld.relaxed.gpu, the protocol of k_vert's inter-strip exchange).  Prints cycles per one-way hop.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void st_rel(uint2* p, uint32_t a, uint32_t b) { asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory"); }
__device__ __forceinline__ uint2 ld_rel(const uint2* p) { uint2 v; asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_vol(uint2* p, uint32_t a, uint32_t b) { asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory"); }
__device__ __forceinline__ uint2 ld_vol(const uint2* p) { uint2 v; asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory"); return v; }

template <int MODE>
__global__ void k(uint2* buf, int partner, int iters, long long* out, int* smid)
{
    // buf[0..31]: written by CTA 0, buf[64..95]: written by the partner
    const int lane = threadIdx.x;
    if (blockIdx.x != 0 && blockIdx.x != partner) return;
    const bool first = blockIdx.x == 0;
    uint2* mine = buf + (first ? 0 : 64) + lane;
    const uint2* theirs = buf + (first ? 64 : 0) + lane;
    if (lane == 0) { int s; asm("mov.u32 %0, %%smid;" : "=r"(s)); smid[first ? 0 : 1] = s; }
    long long t0 = clock64();
    for (int i = 1; i <= iters; i++) {
        if (first) { if (MODE) st_vol(mine, i, i); else st_rel(mine, i, i); }
        while (true) { uint2 v = MODE ? ld_vol(theirs) : ld_rel(theirs); if (__all_sync(0xffffffffu, v.y == uint32_t(i))) break; }
        if (!first) { if (MODE) st_vol(mine, i, i); else st_rel(mine, i, i); }
    }
    long long t1 = clock64();
    if (lane == 0 && first) *out = t1 - t0;
}

int main()
{
    uint2* buf; long long* out; int* smid;
    cudaMalloc(&buf, 4096); cudaMalloc(&out, 8); cudaMalloc(&smid, 8);
    const int iters = 2000;
    for (int mode = 0; mode < 2; mode++)
        for (int partner : {1, 2, 37, 74, 75, 110, 147}) {
            cudaMemset(buf, 0, 4096);
            if (mode) k<1><<<148, 32>>>(buf, partner, iters, out, smid); else k<0><<<148, 32>>>(buf, partner, iters, out, smid);
            long long h; int s[2];
            cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost); cudaMemcpy(s, smid, 8, cudaMemcpyDeviceToHost);
            printf("%s CTA 0 (SM %3d) <-> CTA %3d (SM %3d): %.0f cycles per one-way hop\n", mode ? "volatile   " : "relaxed.gpu", s[0], partner, s[1], double(h) / iters / 2);
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
