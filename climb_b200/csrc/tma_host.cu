// Host-side TMA descriptor (CUtensorMap) construction shared by the kernels. The driver entry point is
// fetched through the runtime (cudaGetDriverEntryPoint), so the library does not link libcuda.
#include "common.cuh"
#include "internal.h"

#include <cuda.h>

namespace climb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || sym == nullptr) {
        cudaGetLastError();
        return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(sym);
    return fn;
}

// bf16 tensor of `rank` dims (dims[0] contiguous), strides in ELEMENTS for dims 1.., 128B-swizzled boxes
int encode_tmap_bf16(void* map_out, const void* ptr, int rank, const long long* dims, const long long* strides_elems,
                     const int* box) {
    EncodeTiledFn fn = encode_fn();
    CLIMB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
    CLIMB_REQUIRE(rank >= 2 && rank <= 5, "encode_tmap: rank %d", rank);
    CLIMB_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA operand not 16-byte aligned");
    cuuint64_t d[5], st[4];
    cuuint32_t b[5], es[5];
    for (int i = 0; i < rank; ++i) {
        d[i] = static_cast<cuuint64_t>(dims[i]);
        b[i] = static_cast<cuuint32_t>(box[i]);
        es[i] = 1;
        if (i > 0) {
            CLIMB_REQUIRE((strides_elems[i - 1] * 2) % 16 == 0, "TMA stride %lld elements is not 16-byte aligned",
                          strides_elems[i - 1]);
            st[i - 1] = static_cast<cuuint64_t>(strides_elems[i - 1]) * 2;
        }
    }
    CUresult r = fn(static_cast<CUtensorMap*>(map_out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), d,
                    st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CLIMB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
    return 0;
}

}  // namespace climb
