// b200sim_step2.h -- launcher of the second-generation fused step kernel (b200sim_step2.cuh),
// compiled in its own translation unit (b200sim_step2.cu).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>

namespace b200sim {

template <typename T>
struct Params;

// threads the <T, G> instance is compiled for (its __launch_bounds__); 0: no such instance
int step2_max_threads(size_t scalar_bytes, int G);
// shared-memory words of T per environment
size_t step2_env_words(size_t scalar_bytes, int nL, int nc, int G);
// launch on `st` (optionally as a programmatic dependent launch); returns a cudaError_t value
template <typename T>
int launch_step2(const Params<T>& P, int G, int grid, int threads, size_t smem, cudaStream_t st, bool pdl);

}  // namespace b200sim
