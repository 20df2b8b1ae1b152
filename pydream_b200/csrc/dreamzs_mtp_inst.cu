// Point-parallel multi-try kernels: one translation unit per layout, compiled with -DDZ_G=<lanes per point> -DDZ_R=<chunk rounds>.
#include "dreamzs_mtp_kernel.cuh"
#define DZ_CAT2(a, b) dreamzs_launch_mtp_##a##_##b
#define DZ_CAT(a, b) DZ_CAT2(a, b)
int DZ_CAT(DZ_G, DZ_R)(const dreamzs::StepParams &P, cudaStream_t stream) {
  return dreamzs::launch_mtp<DZ_G, DZ_R>(P, stream);
}
