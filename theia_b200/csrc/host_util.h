// Host-side helpers shared by the .cu files: error text, launch counter, SM count, TMA
// tensor-map encoding through the driver entry point (no link-time libcuda dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace theia {

int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);
int num_sms();
// CTAs per MMA the GEMM launcher uses for an N tile of bn and m_tiles 128-row tiles (gemm_tc.cu): 2 = CTA pair
int gemm_pair_mode(int bn, int m_tiles, int a_mode);
// rank-D bf16 tensor map, 128-byte (default) or 32-byte swizzle, zero OOB fill. dims[0] is the contiguous dim,
// strides_bytes has rank-1 entries (dims 1..rank-1).
int encode_tensor_map(CUtensorMap* tm, const void* ptr, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides = nullptr,
                      int swizzle_bytes = 128);  // 128 or 32

// Once-per-DEVICE latch (function attributes such as the dynamic shared-memory limit are per device: a process that
// drives several GPUs has to set them on each).  `if (flag.first_use()) { cudaFuncSetAttribute(...); }`
struct PerDeviceOnce {
  bool done[64] = {};
  bool first_use() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;  // unknown device: always (re)apply
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};

#define THEIA_CHECK_LAUNCH(what)                                                                   \
  do {                                                                                             \
    theia::count_launch();                                                                         \
    cudaError_t e__ = cudaGetLastError();                                                          \
    if (e__ != cudaSuccess) return theia::set_error(THEIA_ERR_CUDA, what ": %s", cudaGetErrorString(e__)); \
  } while (0)

}  // namespace theia
