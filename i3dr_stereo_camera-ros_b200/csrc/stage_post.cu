// Launchers of the post-processing stages (A.7, A.8); kernels in k_wta.cuh / k_post.cuh.
#include "stages.h"
#include "k_wta.cuh"
#include "k_post.cuh"

namespace b200sgm {

void launch_lrcheck(int16_t* disp, const uint32_t* disp2key, const Eff& e, cudaStream_t st)
{
    WtaGeom wg{e.W, e.H, e.W1, e.minX1, e.minD, e.D, e.Dp, e.uniq, e.d12, e.INVALID};
    dim3 block(256), grid((e.W1 + 255) / 256, e.H);
    k_lrcheck<<<grid, block, 0, st>>>(disp, disp2key, wg);
}

void launch_median3(const int16_t* src, int16_t* dst, int W, int H, cudaStream_t st)
{
    dim3 block(256), grid((W + 255) / 256, H);
    k_median3<<<grid, block, 0, st>>>(src, dst, W, H);
}

void launch_fill16(int16_t* p, int n, int16_t v, cudaStream_t st) { k_fill16<<<(n + 255) / 256, 256, 0, st>>>(p, n, v); }

cudaError_t launch_speckle(int16_t* img, int* label, int* parent, int* runlen, int* csize, int W, int H, int newVal, int maxSize, int maxDiff,
                           cudaStream_t st, int* launches)
{
    const int npix = W * H;
    cudaError_t ce;
    k_speckle_runs<<<H, 256, 0, st>>>(img, label, parent, runlen, csize, W, newVal, maxDiff);
    ++*launches; if ((ce = cudaGetLastError()) != cudaSuccess) return ce;
    dim3 block(256), grid((W + 255) / 256, H);
    k_speckle_vmerge<<<grid, block, 0, st>>>(img, label, parent, W, H, newVal, maxDiff);
    ++*launches; if ((ce = cudaGetLastError()) != cudaSuccess) return ce;
    k_speckle_size<<<(npix + 255) / 256, 256, 0, st>>>(label, parent, runlen, csize, npix, maxSize);
    ++*launches; if ((ce = cudaGetLastError()) != cudaSuccess) return ce;
    k_speckle_apply<<<(npix + 255) / 256, 256, 0, st>>>(img, label, parent, csize, npix, newVal, maxSize);
    ++*launches; return cudaGetLastError();
}

}  // namespace b200sgm
