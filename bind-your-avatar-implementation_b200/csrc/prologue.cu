// Data-movement and small elementwise kernels of the PER-GENERATION prologue: everything that is
// timestep-invariant — LocalFacialExtractor (models/router.py:157-193), AudioProjModel (models/audio_model.py:78-114), the
// face / router / audio K-V precompute (router.py:247-254, :377-383; audio_model.py:241-256) — runs on bya_gemm_bf16 /
// bya_layernorm_* / bya_attention_d64 plus the kernels here, so that no library (cuBLAS / ATen) kernel is left on the
// product path.  All of them are single passes over a few MB, HBM/L2-bound, 16-byte accesses where the layout allows.
#include "common.cuh"
#include "../../include/bya.h"

namespace bya {

// out[r, c] = bf16(src[r, c]); rows of src / out `lds` / `ldo` elements apart (cat / repeat / unfold / cast of the
// reference's prologue).  VEC: 8 bf16 (or 8 fp32 -> 8 bf16) per thread when pointers, strides and cols allow it.
template <bool SRC_F32, bool VEC>
__global__ void __launch_bounds__(256) copy2d_kernel(const void* __restrict__ src, long long lds, __nv_bfloat16* __restrict__ out,
                                                     long long ldo, int rows, int cols) {
  const int per_row = VEC ? cols / 8 : cols;
  const long long total = (long long)rows * per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = int(i / per_row), c = int(i % per_row);
    if (VEC) {
      uint4 o;
      if (SRC_F32) {
        const float4* s = reinterpret_cast<const float4*>(static_cast<const float*>(src) + r * lds) + 2 * c;
        const float4 a = s[0], b = s[1];
        o.x = pack_bf16x2(a.x, a.y), o.y = pack_bf16x2(a.z, a.w), o.z = pack_bf16x2(b.x, b.y), o.w = pack_bf16x2(b.z, b.w);
      } else {
        o = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(src) + r * lds)[c];
      }
      reinterpret_cast<uint4*>(out + r * ldo)[c] = o;
    } else {
      out[r * ldo + c] = SRC_F32 ? __float2bfloat16(static_cast<const float*>(src)[r * lds + c])
                                 : static_cast<const __nv_bfloat16*>(src)[r * lds + c];
    }
  }
}

// out = act(sum over the k-split slices of ws + bias) as bf16 (slices summed in index order: deterministic)
__global__ void __launch_bounds__(256) splitk_finalize_kernel(const float* __restrict__ ws, int splits, long long slice, long long ldw,
                                                              const __nv_bfloat16* __restrict__ bias, int act,
                                                              __nv_bfloat16* __restrict__ out, long long ldo, int rows, int cols) {
  const int per_row = cols / 4;
  const long long total = (long long)rows * per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = int(i / per_row), c = int(i % per_row) * 4;
    const float* w = ws + r * ldw + c;
    float x[4] = {0.f, 0.f, 0.f, 0.f};
    for (int sidx = 0; sidx < splits; ++sidx) {
      const float4 v = *reinterpret_cast<const float4*>(w + sidx * slice);
      x[0] += v.x, x[1] += v.y, x[2] += v.z, x[3] += v.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (bias) x[j] += __bfloat162float(bias[c + j]);
      if (act == GEMM_ACT_RELU) x[j] = fmaxf(x[j], 0.f);
      else if (act == GEMM_ACT_GELU_TANH) x[j] = gelu_tanh(x[j]);
      else if (act == GEMM_ACT_GELU_ERF) x[j] = gelu_erf(x[j]);
    }
    uint2 o;
    o.x = pack_bf16x2(x[0], x[1]);
    o.y = pack_bf16x2(x[2], x[3]);
    *reinterpret_cast<uint2*>(out + r * ldo + c) = o;
  }
}

// LayerNorm(D = NV*256, affine) + LeakyReLU, one warp per row (the mapping MLPs of LocalFacialExtractor, router.py:118-154)
template <int NV>
__global__ void __launch_bounds__(256) ln_leaky_kernel(const __nv_bfloat16* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ out,
                                                       int ldo, int rows, float eps, const __nv_bfloat16* __restrict__ gamma,
                                                       const __nv_bfloat16* __restrict__ beta, float slope) {
  constexpr int D = NV * 256;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const uint4* src = reinterpret_cast<const uint4*>(x + size_t(warp) * ldx);
  float v[NV * 8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const uint4 u = src[i * 32 + lane];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[i * 8 + 2 * j] = bf16_lo(w[j]);
      v[i * 8 + 2 * j + 1] = bf16_hi(w[j]);
      sum += v[i * 8 + 2 * j] + v[i * 8 + 2 * j + 1];
    }
  }
  const float mean = warp_sum(sum) * (1.f / D);
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < NV * 8; ++i) {
    const float d = v[i] - mean;
    var += d * d;
  }
  const float rstd = rsqrtf(warp_sum(var) * (1.f / D) + eps);
  uint4* dst = reinterpret_cast<uint4*>(out + size_t(warp) * ldo);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c0 = (i * 32 + lane) * 8;
    const uint4 g = *reinterpret_cast<const uint4*>(gamma + c0), b = *reinterpret_cast<const uint4*>(beta + c0);
    const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, bw[4] = {b.x, b.y, b.z, b.w};
    float y[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      y[2 * j] = (v[i * 8 + 2 * j] - mean) * rstd * bf16_lo(gw[j]) + bf16_lo(bw[j]);
      y[2 * j + 1] = (v[i * 8 + 2 * j + 1] - mean) * rstd * bf16_hi(gw[j]) + bf16_hi(bw[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = y[j] > 0.f ? y[j] : y[j] * slope;
    uint4 o;
    o.x = pack_bf16x2(y[0], y[1]), o.y = pack_bf16x2(y[2], y[3]), o.z = pack_bf16x2(y[4], y[5]), o.w = pack_bf16x2(y[6], y[7]);
    dst[i * 32 + lane] = o;
  }
}

// x [G*32, ldx] (k of head h at columns k_off + h*d, v at v_off + h*d)  ->  K [G][H][32][d], Vt [G][H][d][32].
// One block per (group, head): the 32 x d tile of k is copied row by row, v is transposed through shared memory.
__global__ void __launch_bounds__(256) kv_pack_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, int k_off, int v_off,
                                                      __nv_bfloat16* __restrict__ K, __nv_bfloat16* __restrict__ Vt, int heads, int d) {
  __shared__ __nv_bfloat16 tile[32][128 + 2];
  const int g = blockIdx.x / heads, h = blockIdx.x % heads;
  const __nv_bfloat16* src = x + (long long)g * 32 * ldx;
  __nv_bfloat16* kd = K + (size_t(g) * heads + h) * 32 * d;
  __nv_bfloat16* vd = Vt + (size_t(g) * heads + h) * 32 * d;
  for (int i = threadIdx.x; i < 32 * d; i += blockDim.x) {
    const int t = i / d, c = i % d;
    kd[i] = src[t * ldx + k_off + h * d + c];
    tile[t][c] = src[t * ldx + v_off + h * d + c];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * d; i += blockDim.x) {
    const int c = i / 32, t = i % 32;
    vd[i] = tile[t][c];
  }
}

// k [C*32, ldk] (columns h*d + e) -> mat [C*32*H, H*d]: row (c, tok*H + h), columns h*d.. = k[c*32+tok, h*d..], rest 0
__global__ void __launch_bounds__(256) router_keys_scatter_kernel(const __nv_bfloat16* __restrict__ k, long long ldk,
                                                                  __nv_bfloat16* __restrict__ mat, int heads, int d, long long rows) {
  const int width = heads * d;             // columns of mat
  const int vec_per_row = width / 8;
  const long long total = rows * vec_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / vec_per_row;
    const int col = int(i % vec_per_row) * 8;
    const int h = int(row % heads);
    const long long ctok = row / heads;    // c*32 + tok
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (col / d == h) o = *reinterpret_cast<const uint4*>(k + ctok * ldk + col);
    *reinterpret_cast<uint4*>(mat + row * width + col) = o;
  }
}

static int grid_for(long long work_items, int threads = 256) {
  long long nb = (work_items + threads - 1) / threads;
  const long long cap = (long long)bya_host::num_sms() * 16;
  if (nb > cap) nb = cap;
  return int(nb < 1 ? 1 : nb);
}

}  // namespace bya

using namespace bya;

extern "C" int bya_copy2d(void* stream, const void* src, long long lds, int src_f32, void* out, long long ldo, int rows, int cols) {
  // lds may be 0 (one source row broadcast to every output row) or smaller than cols (overlapping sliding windows)
  if (!src || !out || rows <= 0 || cols <= 0 || lds < 0 || ldo < cols) return BYA_ERR_SHAPE;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int esz = src_f32 ? 4 : 2;
  const bool vec = cols % 8 == 0 && ldo % 8 == 0 && (lds * esz) % 16 == 0 && !(reinterpret_cast<uintptr_t>(src) & 15) &&
                   !(reinterpret_cast<uintptr_t>(out) & 15);
  const long long items = (long long)rows * (vec ? cols / 8 : cols);
  const int g = grid_for(items);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  if (src_f32) {
    if (vec) copy2d_kernel<true, true><<<g, 256, 0, s>>>(src, lds, o, ldo, rows, cols);
    else copy2d_kernel<true, false><<<g, 256, 0, s>>>(src, lds, o, ldo, rows, cols);
  } else {
    if (vec) copy2d_kernel<false, true><<<g, 256, 0, s>>>(src, lds, o, ldo, rows, cols);
    else copy2d_kernel<false, false><<<g, 256, 0, s>>>(src, lds, o, ldo, rows, cols);
  }
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}

extern "C" int bya_memset_zero(void* stream, void* ptr, long long bytes) {
  if (!ptr || bytes < 0) return BYA_ERR_SHAPE;
  if (bytes == 0) return BYA_OK;
  return cudaMemsetAsync(ptr, 0, size_t(bytes), reinterpret_cast<cudaStream_t>(stream)) == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}

extern "C" int bya_splitk_finalize(void* stream, const float* ws, int splits, int ws_rows, long long ldw, int row0, const void* bias,
                                   int act, void* out, long long ldo, int rows, int cols) {
  if (!ws || !out || splits <= 0 || rows <= 0 || cols <= 0 || row0 < 0 || row0 + rows > ws_rows) return BYA_ERR_SHAPE;
  if (cols % 4 || ldw % 4 || ldo % 4 || ldw < cols || ldo < cols) return BYA_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(ws) & 15) || (reinterpret_cast<uintptr_t>(out) & 7)) return BYA_ERR_ALIGN;
  if (act < GEMM_ACT_NONE || act > GEMM_ACT_RELU) return BYA_ERR_SHAPE;
  splitk_finalize_kernel<<<grid_for((long long)rows * cols / 4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      ws + row0 * ldw, splits, (long long)ws_rows * ldw, ldw, static_cast<const __nv_bfloat16*>(bias), act,
      static_cast<__nv_bfloat16*>(out), ldo, rows, cols);
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}

extern "C" int bya_layernorm_leakyrelu(void* stream, const void* x, int ldx, void* out, int ldo, int rows, int dim, float eps,
                                       const void* gamma, const void* beta, float slope) {
  if (!x || !out || !gamma || !beta || rows <= 0) return BYA_ERR_SHAPE;
  if (ldx % 8 || ldo % 8) return BYA_ERR_ALIGN;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = (rows + 7) / 8;
  const __nv_bfloat16 *xp = static_cast<const __nv_bfloat16*>(x), *g = static_cast<const __nv_bfloat16*>(gamma),
                      *b = static_cast<const __nv_bfloat16*>(beta);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  switch (dim) {
    case 1024: ln_leaky_kernel<4><<<blocks, 256, 0, s>>>(xp, ldx, o, ldo, rows, eps, g, b, slope); break;
    case 512: ln_leaky_kernel<2><<<blocks, 256, 0, s>>>(xp, ldx, o, ldo, rows, eps, g, b, slope); break;
    case 2048: ln_leaky_kernel<8><<<blocks, 256, 0, s>>>(xp, ldx, o, ldo, rows, eps, g, b, slope); break;
    default: return BYA_ERR_SHAPE;
  }
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}

extern "C" int bya_kv_pack(void* stream, const void* x, long long ldx, int k_off, int v_off, void* K, void* Vt, int groups,
                           int heads, int head_dim) {
  if (!x || !K || !Vt || groups <= 0 || heads <= 0 || (head_dim != 64 && head_dim != 128)) return BYA_ERR_SHAPE;
  if (k_off < 0 || v_off < 0 || ldx < k_off + heads * head_dim || ldx < v_off + heads * head_dim) return BYA_ERR_SHAPE;
  kv_pack_kernel<<<groups * heads, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), ldx, k_off, v_off, static_cast<__nv_bfloat16*>(K), static_cast<__nv_bfloat16*>(Vt), heads,
      head_dim);
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}

extern "C" int bya_router_keys_scatter(void* stream, const void* k, long long ldk, void* mat, int chars, int heads, int head_dim) {
  if (!k || !mat || chars <= 0 || heads <= 0 || head_dim % 8 || ldk % 8 || ldk < heads * head_dim) return BYA_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(k) & 15) || (reinterpret_cast<uintptr_t>(mat) & 15)) return BYA_ERR_ALIGN;
  const long long rows = (long long)chars * 32 * heads;
  router_keys_scatter_kernel<<<grid_for(rows * (heads * head_dim / 8)), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(k), ldk, static_cast<__nv_bfloat16*>(mat), heads, head_dim, rows);
  return cudaGetLastError() == cudaSuccess ? BYA_OK : BYA_ERR_CUDA;
}
