// Host-side helpers shared by the translation units of libcdb200.so: error reporting, CUDA checks,
// TMA tensor-map construction (driver entry point resolved at run time, so the library has no
// link-time dependency on libcuda and loads on a machine without a driver).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <utility>

#include "../../include/cdb200.h"

namespace cdb {

// Thread-local message for cdb_last_error().
char* error_buffer();
int fail(int status, const char* fmt, ...);

#define CDB_CUDA_OK(expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      (void)cudaGetLastError(); /* clear the non-sticky error state */                      \
      return ::cdb::fail(CDB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                         __FILE__, __LINE__);                                               \
    }                                                                                       \
  } while (0)

// After every kernel launch: count it (cdb_launch_count) and surface launch errors.
#define CDB_LAUNCH_OK()                        \
  do {                                         \
    ::cdb::note_launch();                      \
    CDB_CUDA_OK(cudaGetLastError());           \
  } while (0)

#define CDB_REQUIRE(cond, status, ...)                  \
  do {                                                  \
    if (!(cond)) return ::cdb::fail(status, __VA_ARGS__); \
  } while (0)

// Device-side abort flag (set when a bounded mbarrier wait times out).
int* device_abort_flag_ptr();

// Encodes a tiled tensor map with 128-byte swizzle. dims/strides are given innermost first;
// strides[i] (bytes) is the stride of dimension i+1 (dimension 0 is contiguous).
int make_tmap(CUtensorMap* out, CUtensorMapDataType dt, int rank, void* base, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box);

// Same with the swizzle mode selectable: atom32 != 0 -> 128-byte swizzle with 32-byte atomicity (32-byte chunks
// XORed with the row index mod 4), the only shared-memory layout tcgen05 accepts for MN-major TF32 operands.
int make_tmap_swz(CUtensorMap* out, CUtensorMapDataType dt, int rank, void* base, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, int atom32);

int sm_count();
void note_launch();

// CDB_PDL=1 switches programmatic dependent launches on (off by default: see api.cu).
bool pdl_enabled();

// Launch with optional thread-block cluster (x dimension) and programmatic stream serialization (see ptx.cuh pdl_*).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                             int cluster_x, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cluster_x > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cluster_x;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl && pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }
inline int floor_div(int a, int b) {
  int q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
  return q;
}
inline int pos_mod(int a, int b) { return a - floor_div(a, b) * b; }

}  // namespace cdb
