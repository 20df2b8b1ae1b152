// Two-stage single-try step: one translation unit per kernel variant, compiled with -DDZ_G=<lanes per chain> -DDZ_R=<chunk rounds>.
#include "dreamzs_st2_kernel.cuh"
#define DZ_CAT2(a, b) dreamzs_launch_st2_##a##_##b
#define DZ_CAT(a, b) DZ_CAT2(a, b)
int DZ_CAT(DZ_G, DZ_R)(const dreamzs::StepParams &P, int threads, size_t smem, cudaStream_t stream) {
  return dreamzs::launch_st2<DZ_G, DZ_R>(P, threads, smem, stream);
}
