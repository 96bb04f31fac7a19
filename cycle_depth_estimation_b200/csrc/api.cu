// Library plumbing: version, error strings, the device abort flag and tensor-map encoding.
#include <atomic>
#include <mutex>

#include <stdlib.h>

#include "common.cuh"

namespace cdb {

static thread_local char g_err[512] = "";

char* error_buffer() { return g_err; }

int fail(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return status;
}

static std::atomic<long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

__device__ int g_device_abort = 0;

bool pdl_enabled() {
  // measured on the CycleGAN step (graph replay, same box): 30.19 ms without, 30.42 ms with programmatic launches —
  // the early-resident successor keeps the SMs a finishing kernel frees from the weight-gradient stream — so OFF
  // unless CDB_PDL=1
  static const bool on = getenv("CDB_PDL") && atoi(getenv("CDB_PDL")) != 0;
  return on;
}

int* device_abort_flag_ptr() {
  static int* ptr = nullptr;
  if (!ptr) {
    void* p = nullptr;
    if (cudaGetSymbolAddress(&p, g_device_abort) == cudaSuccess) ptr = static_cast<int*>(p);
  }
  return ptr;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap(CUtensorMap* out, CUtensorMapDataType dt, int rank, void* base, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box) {
  return make_tmap_swz(out, dt, rank, base, dims, strides_bytes, box, 0);
}

int make_tmap_swz(CUtensorMap* out, CUtensorMapDataType dt, int rank, void* base, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, int atom32) {
  EncodeTiledFn fn = encode_tiled_fn();
  CDB_REQUIRE(fn != nullptr, CDB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(out, dt, static_cast<cuuint32_t>(rank), base, gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail(CDB_ERR_BAD_DESC,
                "cuTensorMapEncodeTiled failed (%d): rank %d base %p dims [%llu %llu %llu %llu %llu] "
                "strides [%llu %llu %llu %llu] box [%u %u %u %u %u]",
                (int)r, rank, base, (unsigned long long)dims[0],
                (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
                (unsigned long long)(rank > 3 ? dims[3] : 0), (unsigned long long)(rank > 4 ? dims[4] : 0),
                (unsigned long long)(rank > 1 ? strides_bytes[0] : 0),
                (unsigned long long)(rank > 2 ? strides_bytes[1] : 0),
                (unsigned long long)(rank > 3 ? strides_bytes[2] : 0),
                (unsigned long long)(rank > 4 ? strides_bytes[3] : 0), box[0], rank > 1 ? box[1] : 0,
                rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0);
  }
  return CDB_OK;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

}  // namespace cdb

extern "C" {

int cdb_version(void) { return 100; }

long long cdb_launch_count(void) { return cdb::g_launches.load(std::memory_order_relaxed); }

const char* cdb_last_error(void) { return cdb::error_buffer(); }

int cdb_device_abort_flag(void) {
  int* p = cdb::device_abort_flag_ptr();
  if (!p) return -1;
  int v = 0;
  if (cudaMemcpy(&v, p, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  if (v) {
    int zero = 0;
    cudaMemcpy(p, &zero, sizeof(int), cudaMemcpyHostToDevice);
  }
  return v;
}

}  // extern "C"
