// Host side of sx_tma.cuh: tensor-map encoding through the driver entry point (no link-time libcuda dependency).
#include "sx_tma.cuh"

namespace sx {

#ifndef SX_EMU
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int tma_encode(TmaMap* out, const void* base, size_t n0, size_t n1, size_t n2, size_t pitch1, size_t pitch2, int box0, int box1) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    SX_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    SX_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available from this driver");
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  SX_REQUIRE(((uintptr_t)base & 15) == 0 && box0 * 2 <= 256 && box1 <= 256, "tensor map: unsupported geometry");
  const cuuint64_t dims[3] = {(cuuint64_t)n0 * 2, (cuuint64_t)n1, (cuuint64_t)n2};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch1 * sizeof(cplx), (cuuint64_t)pitch2 * sizeof(cplx)};
  const cuuint32_t box[3] = {(cuuint32_t)box0 * 2, (cuuint32_t)box1, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult rc = encode(&out->m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(base), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SX_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (" + std::to_string((int)rc) + ")");
  return 0;
}
#else
int tma_encode(TmaMap* out, const void* base, size_t n0, size_t n1, size_t n2, size_t pitch1, size_t pitch2, int box0, int box1) {
  out->base = (char*)const_cast<void*>(base);
  out->dim[0] = n0 * 2; out->dim[1] = n1; out->dim[2] = n2;
  out->stride[0] = 8; out->stride[1] = pitch1 * sizeof(cplx); out->stride[2] = pitch2 * sizeof(cplx);
  out->box[0] = (unsigned)box0 * 2; out->box[1] = (unsigned)box1; out->box[2] = 1;
  return 0;
}
#endif

}  // namespace sx
