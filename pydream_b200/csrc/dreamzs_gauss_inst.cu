// One translation unit per chains-per-CTA variant of the dense-Gaussian kernel: -DDZ_TC=<chains per CTA>.
#include "dreamzs_gauss_kernel.cuh"
#define DZ_CAT2(a) dreamzs_launch_gauss_##a
#define DZ_CAT(a) DZ_CAT2(a)
int DZ_CAT(DZ_TC)(const dreamzs::StepParams &P, cudaStream_t stream) { return dreamzs::launch_gauss<DZ_TC>(P, stream); }

#if DZ_TC == 8
size_t dreamzs_launch_gauss_smem_bytes(const dreamzs_config &cfg, int TC) { return dreamzs::gauss_smem_bytes(cfg, TC); }
#endif
