// Host-side helpers shared by every entry point of libbya.so: device check, SM count, TMA tensor-map encoding.
#include <cudaTypedefs.h>
#include "common.cuh"
#include "../../include/bya.h"

namespace bya_host {

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                     uint32_t box_inner, uint32_t box_rows, uint64_t batch, uint64_t batch_stride_bytes) {
  auto enc = get_encode();
  if (!enc) return BYA_ERR_DRIVER;
  // inner box of 128 B -> 128B swizzle (operand tiles), 64 B -> 64B swizzle (epilogue staging tiles)
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (row_stride_bytes & 15) || (batch_stride_bytes & 15) ||
      (box_inner * 2 != 128 && box_inner * 2 != 64) || box_rows > 256)
    return BYA_ERR_ALIGN;
  const CUtensorMapSwizzle swz = box_inner * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  const bool three_d = batch > 1;
  cuuint64_t dims[3] = {inner, rows, batch};
  cuuint64_t strides[2] = {row_stride_bytes, batch_stride_bytes};
  cuuint32_t box[3] = {box_inner, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, three_d ? 3 : 2, const_cast<void*>(base), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? BYA_OK : BYA_ERR_DRIVER;
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

}  // namespace bya_host

extern "C" int bya_abi_version(void) { return BYA_ABI_VERSION; }

extern "C" int bya_check_device(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return BYA_ERR_CUDA;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return BYA_ERR_CUDA;
  if (major != 10) return BYA_ERR_ARCH;
  if (!bya_host::get_encode()) return BYA_ERR_DRIVER;
  return BYA_OK;
}
