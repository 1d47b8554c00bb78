#include "agg.cuh"
template int launch_agg_n<4>(b200sgm_engine*, Lane&, const Eff&, cudaStream_t, int);
