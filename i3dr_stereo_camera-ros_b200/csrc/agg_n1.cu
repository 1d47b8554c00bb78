#include "agg.cuh"
template int launch_agg_n<1>(b200sgm_engine*, Lane&, const Eff&, cudaStream_t, int);
