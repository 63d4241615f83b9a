// tma.cuh -- Tensor Memory Accelerator plumbing for the packed kernels (sm_100a):
//   * host: tensor-map encoding through the driver entry point (no link-time libcuda dependency);
//   * device: bulk-tensor store / load PTX, bulk-group bookkeeping, mbarrier helpers.
// The kernels stage a tile of the STFT matrix S[b][f][p] in shared memory and let the TMA engine
// move it: the slice-lane layout makes every group of RL bins a dense [RL rows][TF slices] box.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace veils {
namespace tma {

using EncodeFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                              const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeFn encode_fn() {
    static EncodeFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeFn)p;
    }();
    return fn;
}

// Tensor map of S[b][f][p] (element size esz = 4 or 8 bytes, row pitch ldp elements) seen as
//   dim0 = p (size P), dim1 = d (size RL + 1, stride G ldp), dim2 = g (size G, stride ldp),
//   dim3 = b (size B, stride F ldp):  row f = g + G d  (F = G RL + 1; d = RL exists for g = 0 only
//   and is the Nyquist row -- the other (RL, g) coordinates are never touched).
// Box = [TF slices][rows][1][1]; coordinates (p, d, g, b).  (This dimension order is the one the
// load direction accepts: with g before d, tensor loads / prefetches trap as "illegal instruction"
// on this driver while stores work.)  Returns false when the layout cannot be described (alignment).
// Tensor LOADS need a 16-byte aligned start address in the innermost dimension (measured: an odd
// column of 8-byte elements traps as "illegal instruction"); stores from aligned tiles never hit it.
inline bool make_s_map(CUtensorMap *m, const void *base, int esz, int64_t P, int64_t ldp, int64_t F, int64_t B,
                       int G, int RL, int TF, int box_rows, int box_groups = 1) {
    EncodeFn fn = encode_fn();
    if (!fn) return false;
    if (((uintptr_t)base & 15) != 0 || (ldp * esz) % 16 != 0 || (TF * esz) % 16 != 0) return false;
    if (P <= 0 || B <= 0 || P >= (1LL << 32) || B >= (1LL << 32)) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)P, (cuuint64_t)(RL + 1), (cuuint64_t)G, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)(G * ldp * esz), (cuuint64_t)(ldp * esz), (cuuint64_t)(F * ldp * esz)};
    for (int i = 0; i < 3; ++i)
        if (strides[i] >= (1ULL << 40) || strides[i] % 16 != 0) return false;
    const cuuint32_t box[4] = {(cuuint32_t)TF, (cuuint32_t)box_rows, (cuuint32_t)box_groups, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUtensorMapDataType dt = esz == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32;
    const CUresult r = fn(m, dt, 4, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// 3-pass plans: last-pass group g = g1 + R1 g2 keeps its RL staged rows at work row g1 L1 + g2 L2 + d
// (L1 = RL G / R1 ... dense in the order g1, g2, d), so the whole tile is one 5-D box:
//   dim0 = p, dim1 = d (stride G ldp), dim2 = g2 (stride R1 ldp), dim3 = g1 (stride ldp), dim4 = b.
inline bool make_s_map5(CUtensorMap *m, const void *base, int esz, int64_t P, int64_t ldp, int64_t F, int64_t B,
                        int G, int RL, int R1, int TF) {
    EncodeFn fn = encode_fn();
    if (!fn) return false;
    if (((uintptr_t)base & 15) != 0 || (ldp * esz) % 16 != 0 || (TF * esz) % 16 != 0) return false;
    if (P <= 0 || B <= 0 || P >= (1LL << 32) || B >= (1LL << 32) || G % R1 != 0) return false;
    const int G2 = G / R1;
    if (RL > 256 || G2 > 256 || R1 > 256) return false;
    const cuuint64_t dims[5] = {(cuuint64_t)P, (cuuint64_t)RL, (cuuint64_t)G2, (cuuint64_t)R1, (cuuint64_t)B};
    const cuuint64_t strides[4] = {(cuuint64_t)(G * ldp * esz), (cuuint64_t)(R1 * ldp * esz), (cuuint64_t)(ldp * esz),
                                   (cuuint64_t)(F * ldp * esz)};
    for (int i = 0; i < 4; ++i)
        if (strides[i] >= (1ULL << 40) || strides[i] % 16 != 0) return false;
    const cuuint32_t box[5] = {(cuuint32_t)TF, (cuuint32_t)RL, (cuuint32_t)G2, (cuuint32_t)R1, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUtensorMapDataType dt = esz == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32;
    return fn(m, dt, 5, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Tensor map for LANDING a tile of S[b][f][p] (8-byte elements) in shared memory (wide_inv_kernel): the plain
// 3-D view dim0 = p (size P), dim1 = f (size F, stride ldp), dim2 = b (stride F ldp); box = [8 slices][rows][1].
// Needs a 16-byte aligned base and row pitch.  (4-D views with the smaller row stride in dim1 trap as
// "illegal instruction" in the load direction on this driver -- see make_s_map.)
inline bool make_land_map(CUtensorMap *m, const void *base, int64_t P, int64_t ldp, int64_t F, int64_t B, int box_rows) {
    EncodeFn fn = encode_fn();
    if (!fn) return false;
    if (((uintptr_t)base & 15) != 0 || (ldp * 8) % 16 != 0) return false;
    if (P <= 0 || B <= 0 || P >= (1LL << 31) || B >= (1LL << 31) || box_rows > 256) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)P, (cuuint64_t)F, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)(ldp * 8), (cuuint64_t)(F * ldp * 8)};
    for (int i = 0; i < 2; ++i)
        if (strides[i] >= (1ULL << 40) || strides[i] % 16 != 0) return false;
    const cuuint32_t box[3] = {8, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

#if defined(__CUDACC__)
__device__ __forceinline__ void prefetch_l2_3d(const CUtensorMap *m, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];"
                 ::"l"(m), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// global -> shared, 3-D box, completion on an mbarrier (complete_tx::bytes)
__device__ __forceinline__ void load_3d(const CUtensorMap *m, unsigned dst, unsigned bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void store_5d(const CUtensorMap *m, unsigned src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(m), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
// shared -> global, box at coordinates (c0, c1, c2, c3); completion tracked by the bulk group
#ifndef VEILS_L2_HINT
#define VEILS_L2_HINT 0          // experiment: 1 = evict_first on tensor stores, 2 = on tensor loads (profiles/r1c_ablation.txt)
#endif
__device__ __forceinline__ unsigned long long policy_evict_first() {
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void store_4d(const CUtensorMap *m, unsigned src, int c0, int c1, int c2, int c3) {
#if VEILS_L2_HINT & 1
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4, %5}], [%1], %6;"
                 ::"l"(m), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy_evict_first()) : "memory");
#else
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(m), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
#endif
}
// shared -> global, 1-D: `bytes` (multiple of 16) from a 16-byte aligned shared address to a
// 16-byte aligned global address
__device__ __forceinline__ void store_bulk(void *dst, unsigned src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
// global -> shared, 1-D: `bytes` (multiple of 16), 16-byte aligned addresses; completion on an mbarrier
__device__ __forceinline__ void load_bulk(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk stores of this thread have finished READING shared memory
__device__ __forceinline__ void wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory become visible to the async proxy (TMA)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void prefetch_map(const CUtensorMap *m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
// global -> shared, completion on an mbarrier (complete_tx::bytes)
__device__ __forceinline__ void load_4d(const CUtensorMap *m, unsigned dst, unsigned bar, int c0, int c1, int c2, int c3) {
#if VEILS_L2_HINT & 2
    asm volatile("cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
                 ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy_evict_first()) : "memory");
#else
    asm volatile("cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
#endif
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
#endif

}  // namespace tma
}  // namespace veils
