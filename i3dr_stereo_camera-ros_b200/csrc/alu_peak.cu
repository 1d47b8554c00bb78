// Packed-int16 issue-peak microbenchmark: the ALU roofline denominator for the aggregation kernels.
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/b200sgm.h"

namespace {

constexpr int kIters = 4096;
constexpr int kIlp = 8;

// MODE 0: VIMNMX3.U16x2 (3-input min, 4 elementary ops), 1: VIMNMX.U16x2 (2 ops),
// 2: the aggregation mix per packed register: PRMT, add, min3, min, add3-ish (2 adds), min, add+min.
template <int MODE>
__global__ void __launch_bounds__(256) k_alu(uint32_t* out, uint32_t seed)
{
    uint32_t a[kIlp], b = seed * 0x10001u + threadIdx.x, c = seed ^ 0x00050003u;
#pragma unroll
    for (int i = 0; i < kIlp; i++) a[i] = threadIdx.x * 0x00010001u + i * 0x00030002u + seed;
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int i = 0; i < kIlp; i++) {
            if (MODE == 0) a[i] = __vimin3_u16x2(a[i], b, c + i) + 0x00010001u * 0;   // dependent chain per i
            else if (MODE == 1) a[i] = __vminu2(a[i], b + i);
            else {
                uint32_t q = __byte_perm(a[i], a[(i + 1) % kIlp], 0x5432) + 0x00080008u;
                uint32_t t = __vimin3_u16x2(a[i], q, b);
                t = __vminu2(t, c);
                uint32_t l = a[i] + t - (c & 0x00FF00FFu);
                b = __vminu2(b, l);
                a[i] = __vminu2(l + (q & 0x000F000Fu), 0x7FFF7FFFu);
            }
        }
        if (MODE != 2) { b += 0x00010001u; c ^= b; }
    }
    uint32_t r = b ^ c;
#pragma unroll
    for (int i = 0; i < kIlp; i++) r ^= a[i];
    if (r == 0xDEADBEEFu) out[0] = r;  // keep the work alive
}

template <int MODE>
double run(int sms, double ops_per_inner)
{
    uint32_t* d = nullptr;
    cudaMalloc(&d, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 8;
    k_alu<MODE><<<blocks, 256>>>(d, 1);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        k_alu<MODE><<<blocks, 256>>>(d, rep + 2);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    const double inner = double(blocks) * 256.0 * kIters * kIlp;
    return inner * ops_per_inner / (best * 1e-3) / 1e12;
}

}  // namespace

extern "C" int b200sgm_alu_peak(int device, double tera_ops[3])
{
    if (!tera_ops) return B200SGM_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return B200SGM_ECUDA;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    tera_ops[0] = run<0>(sms, 4.0);
    tera_ops[1] = run<1>(sms, 2.0);
    // mix: per packed register 2 cells x 9 elementary ops of the reference formulation (9 ops/cell/path)
    tera_ops[2] = run<2>(sms, 18.0);
    return cudaGetLastError() == cudaSuccess ? B200SGM_OK : B200SGM_ECUDA;
}
