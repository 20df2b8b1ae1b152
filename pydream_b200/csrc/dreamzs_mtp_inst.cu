// Point-parallel multi-try kernel: one translation unit per lane-group width, compiled with -DDZ_G=<lanes per point>.
#include "dreamzs_mtp_kernel.cuh"
#define DZ_CAT2(a) dreamzs_launch_mtp_##a
#define DZ_CAT(a) DZ_CAT2(a)
int DZ_CAT(DZ_G)(const dreamzs::StepParams &P, size_t smem, cudaStream_t stream) {
  return dreamzs::launch_mtp<DZ_G>(P, smem, stream);
}
