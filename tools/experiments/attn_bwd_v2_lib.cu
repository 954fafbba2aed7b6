// ROUND-2 PROTOTYPE wrapper: exposes the two-CTAs-per-SM attention backward of attention_bwd_v2.cuh through a C entry
// point with the signature of b200_attention_bwd (minus `out`: delta must already be filled, e.g. by the product call).
// build (from the repo root):
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -shared -Xcompiler -fPIC \
//        -o tools/experiments/libattn_v2.so tools/experiments/attn_bwd_v2_lib.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>

#include "attention_bwd_v2.cuh"

namespace {
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode3(CUtensorMap* m, const void* base, long long cols, long long T, long long B) {
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) return -1;
    fn = reinterpret_cast<EncodeFn>(fp);
  }
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(cols) * 2, static_cast<cuuint64_t>(cols) * T * 2};
  cuuint32_t box[3] = {64, 128, 1}, es[3] = {1, 1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fprintf(stderr, "attn_bwd_v2: cuTensorMapEncodeTiled failed (%d)\n", static_cast<int>(r));
  return r == CUDA_SUCCESS ? 0 : -1;
}
}  // namespace

extern "C" int b200x_attention_bwd_v2(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv,
                                      int B, int T, int H, float scale, void* stream) {
  if (T < 1 || T > 256) return -2;
  b200::AttnBwd2Params p;
  memset(&p, 0, sizeof(p));
  p.B = B, p.H = H, p.T = T;
  p.nblk = (T + 127) / 128;
  p.scale = scale;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.lse = lse;
  p.delta = delta;
  const long long HD = static_cast<long long>(H) * 64;
  if (encode3(&p.qkv_map, qkv, 3 * HD, T, B) || encode3(&p.do_map, dout, HD, T, B) || encode3(&p.dqkv_map, dqkv, 3 * HD, T, B))
    return -1;
  static bool cfg = false;
  if (!cfg) {
    if (cudaFuncSetAttribute(b200::attn_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             b200::kAttnBwd2SmemBytes) != cudaSuccess)
      return -3;
    cfg = true;
  }
  b200::attn_bwd2_kernel<<<B * H, 288, b200::kAttnBwd2SmemBytes, static_cast<cudaStream_t>(stream)>>>(p);
  return cudaPeekAtLastError() == cudaSuccess ? 0 : -4;
}
