// Kernel launch helper: cudaLaunchKernelEx with the programmatic-stream-serialization attribute
// (PDL).  A kernel launched this way may begin while its predecessor in the stream is still
// running; it must execute ptx::pdl_wait() before touching memory the predecessor produces
// (all kernels of this library do).  Works under stream capture (programmatic graph edges).
//
// If the runtime rejects a programmatic launch (observed: cudaErrorInvalidValue for some
// launches issued from PyTorch's autograd worker thread), the launch is retried as a plain
// stream-ordered launch and counted in pdl_fallbacks().
#pragma once
#include <cuda_runtime.h>

#include <utility>

#include "bflc_kernels.h"

namespace bflc {

void note_pdl_fallback();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
  if (e == cudaErrorInvalidValue && cfg.numAttrs != 0) {
    (void)cudaGetLastError();  // clear, then retry without the attribute
    cfg.numAttrs = 0;
    note_pdl_fallback();
    e = cudaLaunchKernelEx(&cfg, kernel, args...);
  }
  return e;
}

}  // namespace bflc
