// window_loop.cu -- the persistent window loop (k_window_loop of kernels.cuh) in a
// translation unit of its own: its own copies of the device functions, so that the
// loop's register pressure and stack frame never touch the per-window kernel
// (engine.cu), which the large configurations run.
#include <cuda_runtime.h>

#include "../../include/simian_b200.h"
#include "device_types.cuh"

#define SIMIAN_LOOP_TU 1
namespace simian_loop_tu {
#include "kernels.cuh"
} // namespace simian_loop_tu

// host entry points used by engine.cu
int simian_loop_blocks_per_sm(size_t smem_bytes) {
  int per_sm = 0;
  if (cudaFuncSetAttribute(simian_loop_tu::k_window_loop,
                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)smem_bytes) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, simian_loop_tu::k_window_loop,
                                                    SIMIAN_THREADS, smem_bytes) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return per_sm;
}

cudaError_t simian_loop_launch(const DevCfg &cfg, uint32_t win_ix, uint32_t max_windows,
                               size_t smem_bytes, cudaStream_t stream) {
  void *args[] = {(void *)&cfg, (void *)&win_ix, (void *)&max_windows};
  return cudaLaunchCooperativeKernel((const void *)simian_loop_tu::k_window_loop,
                                     dim3(cfg.n_blocks), dim3(SIMIAN_THREADS), args, smem_bytes,
                                     stream);
}

size_t simian_loop_smem_bytes(uint32_t epb) {
  return sizeof(simian_loop_tu::WindowShared) + (size_t)epb * sizeof(EntityCore);
}
