// b200sim_step2.cu -- instances + launcher of step2_kernel (see b200sim_step2.cuh).
#include "b200sim_step2.cuh"
#include "b200sim_step2.h"

namespace b200sim {

namespace {

template <typename T, int G, int THREADS>
int launch_inst(const Params<T>& P, int grid, int threads, size_t smem, cudaStream_t st, bool pdl) {
  auto kern = step2_kernel<T, G, THREADS>;
  if (threads > THREADS) return (int)cudaErrorInvalidConfiguration;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  if (pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, kern, P);
  }
  kern<<<grid, threads, smem, st>>>(P);
  return (int)cudaGetLastError();
}

}  // namespace

#ifndef B200SIM_S2_F8
#define B200SIM_S2_F8 384
#endif
#ifndef B200SIM_S2_F16
#define B200SIM_S2_F16 448
#endif
#ifndef B200SIM_S2_D8
#define B200SIM_S2_D8 192
#endif
#ifndef B200SIM_S2_D16
#define B200SIM_S2_D16 384
#endif

int step2_max_threads(size_t scalar_bytes, int G) {
  if (scalar_bytes == 4) return G == 8 ? B200SIM_S2_F8 : (G == 16 ? B200SIM_S2_F16 : 0);
  if (scalar_bytes == 8) return G == 8 ? B200SIM_S2_D8 : (G == 16 ? B200SIM_S2_D16 : 0);
  return 0;
}

size_t step2_env_words(size_t scalar_bytes, int nL, int nc, int G) { return env2_ws_words(scalar_bytes, nL, nc, G); }

template <>
int launch_step2<float>(const Params<float>& P, int G, int grid, int threads, size_t smem, cudaStream_t st, bool pdl) {
  if (G == 8) return launch_inst<float, 8, B200SIM_S2_F8>(P, grid, threads, smem, st, pdl);
  if (G == 16) return launch_inst<float, 16, B200SIM_S2_F16>(P, grid, threads, smem, st, pdl);
  return (int)cudaErrorInvalidConfiguration;
}

template <>
int launch_step2<double>(const Params<double>& P, int G, int grid, int threads, size_t smem, cudaStream_t st, bool pdl) {
  if (G == 8) return launch_inst<double, 8, B200SIM_S2_D8>(P, grid, threads, smem, st, pdl);
  if (G == 16) return launch_inst<double, 16, B200SIM_S2_D16>(P, grid, threads, smem, st, pdl);
  return (int)cudaErrorInvalidConfiguration;
}

}  // namespace b200sim
