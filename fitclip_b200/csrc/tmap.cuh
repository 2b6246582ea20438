// Tensor-map construction with a per-thread cache.  cuTensorMapEncodeTiled is a pure function of its arguments (it
// touches no device state), and one encoder pass builds 3-4 maps for each of ~600 launches from a handful of distinct
// (pointer, shape) tuples -- the arena of an fc_model never moves -- so the 128-byte descriptors are memoised in a small
// direct-mapped table keyed by every argument.  bf16, <= 3 dimensions, 128-byte swizzle, L2 promotion 256 B: the only
// kind of map this library uses.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "common.cuh"

namespace fc {

typedef CUresult (*TmapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline TmapEncodeTiledFn tmap_encode_fn() {
  static TmapEncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<TmapEncodeTiledFn>(sym);
  }
  return fn;
}

struct TmapKey {
  const void* base;
  uint64_t dims[3];
  uint64_t strides[2];  // bytes, dimensions 1 and 2
  uint32_t box[3];
  uint32_t rank;
  bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};

// rank 2 or 3; dims[0] is the contiguous dimension.  Returns FC_OK / FC_ERR_CUDA (message set).
inline int tmap_bf16_sw128(CUtensorMap* tm, const void* base, uint32_t rank, const uint64_t* dims, const uint64_t* strides,
                           const uint32_t* box) {
  struct Entry {
    TmapKey key;
    CUtensorMap map;
    bool valid;
  };
  constexpr int SLOTS = 512;
  static thread_local Entry* table = nullptr;
  if (!table) table = new Entry[SLOTS]();
  TmapKey key;
  memset(&key, 0, sizeof(key));  // padding bytes take part in the comparison
  key.base = base;
  key.rank = rank;
  for (uint32_t i = 0; i < rank; ++i) {
    key.dims[i] = dims[i];
    key.box[i] = box[i];
    if (i + 1 < rank) key.strides[i] = strides[i];
  }
  uint64_t h = 1469598103934665603ull;  // FNV-1a over the key
  const unsigned char* kb = reinterpret_cast<const unsigned char*>(&key);
  for (size_t i = 0; i < sizeof(key); ++i) h = (h ^ kb[i]) * 1099511628211ull;
  Entry& e = table[h % SLOTS];
  if (e.valid && e.key == key) {
    *tm = e.map;
    return FC_OK;
  }
  TmapEncodeTiledFn fn = tmap_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return FC_ERR_CUDA;
  }
  cuuint64_t d[3] = {dims[0], rank > 1 ? dims[1] : 1, rank > 2 ? dims[2] : 1};
  cuuint64_t st[2] = {strides[0], rank > 2 ? strides[1] : 0};
  cuuint32_t bx[3] = {box[0], rank > 1 ? box[1] : 1, rank > 2 ? box[2] : 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), d, st, bx, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) rank=%u dims=%llu x %llu x %llu box=%u x %u x %u",
              static_cast<int>(r), rank, static_cast<unsigned long long>(d[0]), static_cast<unsigned long long>(d[1]),
              static_cast<unsigned long long>(d[2]), bx[0], bx[1], bx[2]);
    return FC_ERR_CUDA;
  }
  e.key = key;
  e.map = *tm;
  e.valid = true;
  return FC_OK;
}

}  // namespace fc
