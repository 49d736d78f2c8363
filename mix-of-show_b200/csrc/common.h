// common.h — host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mos_sm100.h"

namespace mos {

int set_err(int code, const char* fmt, ...);  // stores thread-local message, returns code

#define MOS_CHECK_ARG(cond, ...)                                  \
  do {                                                            \
    if (!(cond)) return ::mos::set_err(MOS_EINVAL, __VA_ARGS__);  \
  } while (0)

#define MOS_CHECK_CUDA(expr)                                                                       \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::mos::set_err(MOS_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                            __FILE__, __LINE__);                                                   \
  } while (0)

#define MOS_CHECK_DTYPE(dt, who) \
  MOS_CHECK_ARG((dt) == MOS_DT_BF16 || (dt) == MOS_DT_F16, who ": act_dtype %d is neither MOS_DT_BF16 nor MOS_DT_F16", (int)(dt))

#define MOS_CHECK_LAUNCH() MOS_CHECK_CUDA(cudaGetLastError())

// Encode a tiled bf16/fp32 tensor map. dims/strides innermost first; strides[i] is the byte pitch of dim i+1.
// swizzle: 0 none, 1 = 32B, 2 = 64B, 3 = 128B. Returns 0 or MOS_E*.
int encode_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, int swizzle);

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
// Launch with programmatic stream serialization (PDL): the kernel must call pdl_wait() before touching global memory.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

}  // namespace mos
