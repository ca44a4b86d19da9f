// Row gather through the TMA engine (tensor-map copies, SASS UTMALDG): cp.async.bulk.tensor.2d ... tile::gather4 fetches FOUR rows of
// a 2-D tensor, given by their row indices, into four consecutive 128-byte rows of shared memory and applies the 128-byte swizzle on
// the way -- so a gathered row block is directly a SWIZZLE_128B K-major tcgen05.mma operand.  One instruction moves 512 bytes; the
// LSU form (cp.async, 16 bytes per lane) needs 32 lane-requests for the same bytes and is bound by the requests an SM keeps in
// flight (profiles/probe_kernels.py).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace kpf {

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda at link time)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// Tensor map of a row-major [rows][row_elems] tensor of 16-bit elements (row pitch row_bytes) for gather4 copies of 64-element
// (128-byte) row segments with the 128-byte swizzle.  0 on success.
inline int make_row_gather_map(CUtensorMap* map, const void* base, unsigned long long rows, unsigned row_elems, unsigned long long row_bytes) {
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return KPF_ERR_UNSUPPORTED;
    const cuuint64_t dims[2] = {row_elems, rows};
    const cuuint64_t strides[1] = {row_bytes};
    const cuuint32_t box[2] = {64, 1};          // gather4: a box is one row segment; an instruction moves four of them
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : KPF_ERR_BAD_ARGUMENT;
}

// rows r0..r3, elements [col, col + 64) -> dst[0..512): four swizzled 128-byte rows; completes 512 bytes on `bar`
__device__ __forceinline__ void tma_gather4(void* smem_dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar))
                 : "memory");
}

// 64-bit shared-memory descriptor halves of a SWIZZLE_128B K-major operand: rows of 128 bytes (64 16-bit K elements), 8-row atoms of
// 1024 bytes (1024-byte aligned), `sbo` bytes between atoms along the rows; a K = 16 step advances the start address by 32 bytes.
__device__ __forceinline__ uint32_t desc_lo_sw128(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint32_t desc_hi_sw128(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29); }

}  // namespace kpf
