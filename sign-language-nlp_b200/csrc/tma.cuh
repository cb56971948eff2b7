// TMA (cp.async.bulk.tensor) + tcgen05 kind::tf32 helpers shared by the TMA-fed kernels
// (gemm_tma.cu, rnn_step_tc.cu): PTX wrappers, swizzled UMMA descriptors, cached tensor maps.
#pragma once
#include <cuda.h>

#include <unordered_map>

#include "tc05.cuh"

namespace slnlp {

constexpr int UMMA_K = 8;   // tf32: 8 elements = 32 bytes per instruction

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 128-byte-swizzled operand descriptor (cute::UMMA::SmemDescriptor): layout_type 2 = SWIZZLE_128B
// (16-byte chunks), 1 = SWIZZLE_128B_BASE32B (32-byte chunks: the only MN-major layout of 32-bit operands)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint64_t layout_type) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (layout_type << 61);
}
// c = f32, a = b = tf32, per-operand major bit (0 = K-major, 1 = MN-major)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}


// ---------------------------------------------------------------- host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

struct MapKey {
  const void* ptr;
  uint64_t d0, d1, d2, ld1, ld2;
  uint32_t b1;      // box rows; bit 31 = MN-major (32-byte swizzle atom)
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && ld1 == o.ld1 && ld2 == o.ld2 && b1 == o.b1;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    for (uint64_t v : {k.d0, k.d1, k.d2, k.ld1, k.ld2, (uint64_t)k.b1}) h = h * 1000003u ^ (size_t)v;
    return h;
  }
};

// fp32 tensor [d2][d1][d0] (d0 contiguous; row stride ld1 floats, slab stride ld2 floats), TMA box
// {32, b1, 1}.  mn_major selects the 32-byte-atom 128 B swizzle that 32-bit MN-major UMMA operands
// need; otherwise the plain 128 B swizzle of K-major operands.  Maps are cached per thread.
inline bool tensor_map3(const float* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld1, uint64_t ld2, uint32_t b1,
                        bool mn_major, CUtensorMap* out, bool rank3 = true) {
  thread_local std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  const MapKey key{ptr, d0, d1, d2, ld1, ld2, b1 | (mn_major ? 0x80000000u : 0u) | (rank3 ? 0x40000000u : 0u)};
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return true;
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  if (((uintptr_t)ptr & 15) || (ld1 & 3) || (ld2 & 3)) return false;
  const cuuint64_t gdim[3] = {d0, d1, d2};
  const cuuint64_t gstr[2] = {ld1 * sizeof(float), ld2 * sizeof(float)};
  const cuuint32_t box[3] = {32, b1, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const cuuint32_t rank = rank3 ? 3 : 2;
  if (fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<float*>(ptr), gdim, gstr, box, estr,
         CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *out);
  return true;
}
inline bool tensor_map(const float* ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b1, bool mn_major, CUtensorMap* out) {
  return tensor_map3(ptr, d0, d1, 1, ld, ld * d1, b1, mn_major, out, false);
}

// bf16 tensor [d2][d1][d0] (d0 contiguous; row stride ld1, slab stride ld2, in elements), TMA box {64, b1, 1},
// 128-byte swizzle (K-major and MN-major 16-bit UMMA operands both use it).  d2 = 0: a rank-2 map.
inline bool tensor_map3_bf16(const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld1, uint64_t ld2, uint32_t b1,
                             CUtensorMap* out) {
  thread_local std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  const MapKey key{ptr, d0, d1, d2, ld1, ld2, b1};
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return true;
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  if (((uintptr_t)ptr & 15) || (ld1 & 7) || (ld2 & 7)) return false;
  const cuuint64_t gdim[3] = {d0, d1, d2 ? d2 : 1};
  const cuuint64_t gstr[2] = {ld1 * 2, ld2 * 2};
  const cuuint32_t box[3] = {64, b1, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  if (fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, d2 ? 3 : 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *out);
  return true;
}
inline bool tensor_map_bf16(const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b1, CUtensorMap* out) {
  return tensor_map3_bf16(ptr, d0, d1, 0, ld, 0, b1, out);
}
// c = f32, a = b = bf16, per-operand major bit (0 = K-major, 1 = MN-major)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

}  // namespace slnlp
