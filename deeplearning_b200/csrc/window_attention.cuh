// Shifted-window multi-head self-attention (Swin), head_dim 32, 7x7 windows (49 tokens), on tcgen05 - forward and backward.
//
//   attn = softmax( scale * q k^T + relative_position_bias[h] (+ shift mask[window]) ) ;  out = attn v
//
// The cyclic shift (torch.roll), window_partition and window_reverse of the reference are folded into the addressing:
// a window's 49 tokens are gathered with 16-byte cp.async copies straight from the un-rolled qkv tensor [B,H,W,3C]
// (pixel ((wy*7+i+shift)%H, (wx*7+j+shift)%W)) into the 128B-swizzled operand layout, and the result rows are scattered
// back to the same pixels, so none of the four full-tensor permutation passes of the reference is executed.
// Per (batch, window, head): S = Q K^T (M=128 with 49 valid rows, N=64 keys, K=32) in TMEM, the soft-max threads own one
// query row each, P (bf16) goes through shared memory into O = P V. Scores never touch HBM; only the row log-sum-exp is kept.
// CTAs are persistent over (batch, window) pairs of ONE head so that the backward pass can accumulate the gradient of the
// relative-position bias in registers and flush it with one atomicAdd per element per CTA.
//
// Replaces WindowAttention.forward and the roll / window_partition / window_reverse / roll sequence of
// SwinTransformerBlock.forward (classification/swin_transformer/models/swin_transformer.py:118-149, :251-280), including the
// optional fused kernels of kernels/window_process (the --fused_window_process path).
#pragma once
#include "common.cuh"

namespace b200 {

struct WAttnParams {
  const __nv_bfloat16* qkv;   // [B][H][W][3*C]
  __nv_bfloat16* out;         // fwd: [B][H][W][C]
  const float* bias;          // dense [nH][49][49]
  const float* mask;          // [nW][49][49] (0 / -100) or null
  float* lse;                 // [B][nW][nH][49]
  int B, H, W, nH, shift;
  float scale;
  // backward only
  const __nv_bfloat16* o;     // forward output [B][H][W][C]
  const __nv_bfloat16* dout;  // [B][H][W][C]
  __nv_bfloat16* dqkv;        // [B][H][W][3*C]
  float* dbias;               // dense [nH][49][49], accumulated with atomics (zeroed by the caller)
};

constexpr int kWS = 7, kWT = 49;

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// pixel offset (in pixels) of token t of window (wy, wx) in the un-rolled image
__device__ __forceinline__ long long wattn_pixel(const WAttnParams& p, int b, int wy, int wx, int t) {
  const int i = t / kWS, j = t - i * kWS;
  int y = wy * kWS + i + p.shift, x = wx * kWS + j + p.shift;
  if (y >= p.H) y -= p.H;
  if (x >= p.W) x -= p.W;
  return (static_cast<long long>(b) * p.H + y) * p.W + x;
}

// Gather one 49 x 32 bf16 tile (64 B per token) into a [64][128B] 128B-swizzled tile (only the first 4 chunks of a row are used).
__device__ __forceinline__ void wattn_gather(uint32_t tile_s, const __nv_bfloat16* base, long long row_stride, int col0,
                                             const WAttnParams& p, int b, int wy, int wx, int tid, int nthreads) {
  for (int idx = tid; idx < kWT * 4; idx += nthreads) {
    const int t = idx >> 2, c = idx & 3;
    const long long pix = wattn_pixel(p, b, wy, wx, t);
    cp_async16(tile_s + t * 128 + ((c ^ (t & 7)) << 4), base + pix * row_stride + col0 + c * 8);
  }
}

constexpr int kWAttnFwdSmem = 3 * 8192 + 16384 + 256 + 1024;

__global__ void __launch_bounds__(160, 3) wattn_fwd_kernel(const WAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;            // [64][128B] (the M=128 MMA also reads the 64 rows that follow: the K tile, harmless)
  uint8_t* sK = smem + 8192;
  uint8_t* sV = smem + 16384;
  uint8_t* sP = smem + 24576;    // [128][128B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 24576 + 16384);
  uint64_t* bar_s = bars;
  uint64_t* bar_o = bars + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = p.nH * 32;
  const int nWy = p.H / kWS, nWx = p.W / kWS, nW = nWy * nWx;

  // zero the operand tiles once: pad rows (49..63) of K / V must be finite zeros for every item
  for (int i = threadIdx.x; i < (24576 + 16384) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp_idx == 4) {
    if (lane == 0) {
      mbar_init(bar_s, 1);
      mbar_init(bar_o, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<128>(tmem_ptr_smem);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t q_s = smem_u32(sQ), k_s = smem_u32(sK), v_s = smem_u32(sV), p_s = smem_u32(sP);
  const uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);   // S[128 x 64 keys] = Q K^T, both K-major, K = 32 (2 steps)
  const uint32_t idesc_o = make_idesc_bf16(128, 32, 0, 1);   // O[128 x 32] = P (K-major) V (MN-major), K = 64 keys

  const int head = blockIdx.x % p.nH;
  const int lanes = gridDim.x / p.nH;           // CTAs sharing this head
  const int total = p.B * nW;
  uint32_t phase = 0;
  for (int item = blockIdx.x / p.nH; item < total; item += lanes, phase ^= 1) {
    const int b = item / nW, win = item - b * nW;
    const int wy = win / nWx, wx = win - wy * nWx;
    // ---- gather Q, K, V of this (window, head)
    wattn_gather(q_s, p.qkv, 3 * C, head * 32, p, b, wy, wx, threadIdx.x, blockDim.x);
    wattn_gather(k_s, p.qkv, 3 * C, C + head * 32, p, b, wy, wx, threadIdx.x, blockDim.x);
    wattn_gather(v_s, p.qkv, 3 * C, 2 * C + head * 32, p, b, wy, wx, threadIdx.x, blockDim.x);
    cp_async_wait_all();
    fence_proxy_async_smem();
    __syncthreads();
    if (warp_idx == 4) {
      if (lane == 0) {
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 2; ++k)
          umma_f16(tmem_base, make_smem_desc_sw128(q_s + k * 32, 16, 1024), make_smem_desc_sw128(k_s + k * 32, 16, 1024),
                   idesc_s, k > 0 ? 1u : 0u);
        umma_commit(bar_s);
      }
    } else {
      // ---- soft-max: thread = query row
      const int row = warp_idx * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp_idx * 32) << 16);
      mbar_wait(bar_s, phase);
      tc_fence_after();
      uint32_t v[64];
      {
        uint32_t(&lo)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
        uint32_t(&hi)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
        tmem_ld_32x32(taddr, lo);
        tmem_ld_32x32(taddr + 32, hi);
        tmem_ld_wait();
      }
      float mx = -INFINITY;
      const bool valid = row < kWT;
      const float* brow = p.bias + (static_cast<long long>(head) * kWT + (valid ? row : 0)) * kWT;
      const float* mrow = p.mask ? p.mask + (static_cast<long long>(win) * kWT + (valid ? row : 0)) * kWT : nullptr;
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        float s = -INFINITY;
        if (j < kWT) {
          s = fmaf(__uint_as_float(v[j]), p.scale, __ldg(brow + j));
          if (mrow) s += __ldg(mrow + j);
        }
        v[j] = __float_as_uint(s);
        mx = fmaxf(mx, s);
      }
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 64; j += 2) {
        const float e0 = valid ? __expf(__uint_as_float(v[j]) - mx) : 0.f;
        const float e1 = valid ? __expf(__uint_as_float(v[j + 1]) - mx) : 0.f;
        const uint32_t w = pack_bf16x2(e0, e1);
        sum += bf16_lo(w) + bf16_hi(w);
        v[j >> 1] = w;  // in place: slot j/2 has already been consumed
      }
#pragma unroll
      for (int c = 0; c < 8; ++c)
        sts128(p_s + row * 128 + ((c ^ (row & 7)) << 4), v[c * 4], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]);
      if (valid) p.lse[((static_cast<long long>(b) * nW + win) * p.nH + head) * kWT + row] = mx + __logf(sum);
      tc_fence_before();
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (threadIdx.x == 0) {
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_f16(tmem_base + 64, make_smem_desc_sw128(p_s + ks * 32, 16, 1024),
                   make_smem_desc_sw128(v_s + ks * 2048, 8192, 1024), idesc_o, ks > 0 ? 1u : 0u);
        umma_commit(bar_o);
      }
      mbar_wait(bar_o, phase);
      tc_fence_after();
      uint32_t ov[32];
      tmem_ld_32x32(taddr + 64, ov);
      tmem_ld_wait();
      if (valid) {
        const float inv = 1.0f / sum;
        __nv_bfloat16* dst = p.out + wattn_pixel(p, b, wy, wx, row) * C + head * 32;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(ov[c * 8 + 0]) * inv, __uint_as_float(ov[c * 8 + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(ov[c * 8 + 2]) * inv, __uint_as_float(ov[c * 8 + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(ov[c * 8 + 4]) * inv, __uint_as_float(ov[c * 8 + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(ov[c * 8 + 6]) * inv, __uint_as_float(ov[c * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + c * 8) = w;
        }
      }
      tc_fence_before();
    }
    __syncthreads();  // tiles and TMEM are free for the next item
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 4) {
    tc_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Backward. Per (batch, window, head):
//   P = exp(scale*S + bias + mask - lse);  dP = dO V^T;  dS = P*(dP - delta), delta_i = <dO_i, O_i>;  dbias += dS
//   dV = P^T dO;  dQ = scale * dS K;  dK = scale * dS^T Q          (dS is stored pre-multiplied by scale)
constexpr int kWAttnBwdSmem = 5 * 8192 + 16384 + 256 + 1024;

__global__ void __launch_bounds__(160, 3) wattn_bwd_kernel(const WAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // order matters: the M=128 MMAs with a 64-row A tile read 64 further rows of whatever tile follows (finite data)
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 8192;
  uint8_t* sV = smem + 16384;
  uint8_t* sdO = smem + 24576;
  uint8_t* sO = smem + 32768;
  uint8_t* sP = smem + 40960;  // [128][128B]: P, then dS in place
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 40960 + 16384);
  uint64_t* bar_s = bars;
  uint64_t* bar_dp = bars + 1;
  uint64_t* bar_dq = bars + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 3);
  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = p.nH * 32;
  const int nWy = p.H / kWS, nWx = p.W / kWS, nW = nWy * nWx;

  for (int i = threadIdx.x; i < (40960 + 16384) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp_idx == 4) {
    if (lane == 0) {
      mbar_init(bar_s, 1);
      mbar_init(bar_dp, 1);
      mbar_init(bar_dq, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<128>(tmem_ptr_smem);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t q_s = smem_u32(sQ), k_s = smem_u32(sK), v_s = smem_u32(sV), do_s = smem_u32(sdO), o_s = smem_u32(sO),
                 p_s = smem_u32(sP);
  const uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);   // [128 q] x [64 keys], K = 32
  const uint32_t idesc_t = make_idesc_bf16(128, 32, 1, 1);   // A^T B: A, B both MN-major (P^T dO, dS^T Q), K = 64 query rows
  const uint32_t idesc_q = make_idesc_bf16(128, 32, 0, 1);   // dS (K-major) x K (MN-major), K = 64 keys
  constexpr uint32_t kColS = 0, kColDV = 64, kColDK = 96;

  const int head = blockIdx.x % p.nH;
  const int lanes = gridDim.x / p.nH;
  const int total = p.B * nW;
  const int row = (warp_idx & 3) * 32 + lane;
  const bool valid = warp_idx < 4 && row < kWT;
  float db[kWT];  // gradient of bias[head][row][:] accumulated over this CTA's items
#pragma unroll
  for (int j = 0; j < kWT; ++j) db[j] = 0.f;

  uint32_t phase = 0;
  for (int item = blockIdx.x / p.nH; item < total; item += lanes, phase ^= 1) {
    const int b = item / nW, win = item - b * nW;
    const int wy = win / nWx, wx = win - wy * nWx;
    wattn_gather(q_s, p.qkv, 3 * C, head * 32, p, b, wy, wx, threadIdx.x, blockDim.x);
    wattn_gather(k_s, p.qkv, 3 * C, C + head * 32, p, b, wy, wx, threadIdx.x, blockDim.x);
    wattn_gather(v_s, p.qkv, 3 * C, 2 * C + head * 32, p, b, wy, wx, threadIdx.x, blockDim.x);
    wattn_gather(do_s, p.dout, C, head * 32, p, b, wy, wx, threadIdx.x, blockDim.x);
    wattn_gather(o_s, p.o, C, head * 32, p, b, wy, wx, threadIdx.x, blockDim.x);
    cp_async_wait_all();
    fence_proxy_async_smem();
    __syncthreads();
    if (warp_idx == 4) {
      if (lane == 0) {
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 2; ++k)
          umma_f16(tmem_base + kColS, make_smem_desc_sw128(q_s + k * 32, 16, 1024),
                   make_smem_desc_sw128(k_s + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(bar_s);
      }
      __syncwarp();
    } else {
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp_idx * 32) << 16);
      const long long pix = wattn_pixel(p, b, wy, wx, valid ? row : 0);
      // delta_i = <dO_i, O_i> from the gathered tiles (64 B each)
      float delta = 0.f;
      if (valid) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t off = row * 128 + ((c ^ (row & 7)) << 4);
          float a[8], o[8];
          unpack8(*reinterpret_cast<const uint4*>(sdO + off), a);
          unpack8(*reinterpret_cast<const uint4*>(sO + off), o);
#pragma unroll
          for (int e = 0; e < 8; ++e) delta = fmaf(a[e], o[e], delta);
        }
      }
      const float lse = valid ? p.lse[((static_cast<long long>(b) * nW + win) * p.nH + head) * kWT + row] : 0.f;
      const float* brow = p.bias + (static_cast<long long>(head) * kWT + (valid ? row : 0)) * kWT;
      const float* mrow = p.mask ? p.mask + (static_cast<long long>(win) * kWT + (valid ? row : 0)) * kWT : nullptr;
      // ---- P
      mbar_wait(bar_s, phase);
      tc_fence_after();
      {
        uint32_t v[64];
        uint32_t(&lo)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
        uint32_t(&hi)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
        tmem_ld_32x32(taddr + kColS, lo);
        tmem_ld_32x32(taddr + kColS + 32, hi);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 64; j += 2) {
          float e0 = 0.f, e1 = 0.f;
          if (valid && j < kWT) {
            float s0 = fmaf(__uint_as_float(v[j]), p.scale, __ldg(brow + j));
            if (mrow) s0 += __ldg(mrow + j);
            e0 = __expf(s0 - lse);
          }
          if (valid && j + 1 < kWT) {
            float s1 = fmaf(__uint_as_float(v[j + 1]), p.scale, __ldg(brow + j + 1));
            if (mrow) s1 += __ldg(mrow + j + 1);
            e1 = __expf(s1 - lse);
          }
          v[j >> 1] = pack_bf16x2(e0, e1);
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
          sts128(p_s + row * 128 + ((c ^ (row & 7)) << 4), v[c * 4], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (threadIdx.x == 0) {
        tc_fence_after();
        // dP = dO V^T  (into the S columns)      and      dV = P^T dO
#pragma unroll
        for (int k = 0; k < 2; ++k)
          umma_f16(tmem_base + kColS, make_smem_desc_sw128(do_s + k * 32, 16, 1024),
                   make_smem_desc_sw128(v_s + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)  // 64 query rows = 4 steps of 16
          umma_f16(tmem_base + kColDV, make_smem_desc_sw128(p_s + ks * 2048, 8192, 1024),
                   make_smem_desc_sw128(do_s + ks * 2048, 8192, 1024), idesc_t, ks > 0 ? 1u : 0u);
        umma_commit(bar_dp);
      }
      // ---- dS (in place over P)
      mbar_wait(bar_dp, phase);
      tc_fence_after();
      {
        uint32_t v[64];
        uint32_t(&lo)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
        uint32_t(&hi)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
        tmem_ld_32x32(taddr + kColS, lo);
        tmem_ld_32x32(taddr + kColS + 32, hi);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t off = row * 128 + ((c ^ (row & 7)) << 4);
          float pv[8];
          unpack8(*reinterpret_cast<const uint4*>(sP + off), pv);  // P of this row (bf16, as the tensor core saw it)
          float d[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int j = c * 8 + e;
            d[e] = valid ? pv[e] * (__uint_as_float(v[j]) - delta) : 0.f;
            if (j < kWT) db[j] += d[e];
            d[e] *= p.scale;
          }
          sts128(p_s + off, pack_bf16x2(d[0], d[1]), pack_bf16x2(d[2], d[3]), pack_bf16x2(d[4], d[5]), pack_bf16x2(d[6], d[7]));
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (threadIdx.x == 0) {
        tc_fence_after();
        // dQ = dS K (S columns again) ;  dK = dS^T Q
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_f16(tmem_base + kColS, make_smem_desc_sw128(p_s + ks * 32, 16, 1024),
                   make_smem_desc_sw128(k_s + ks * 2048, 8192, 1024), idesc_q, ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_f16(tmem_base + kColDK, make_smem_desc_sw128(p_s + ks * 2048, 8192, 1024),
                   make_smem_desc_sw128(q_s + ks * 2048, 8192, 1024), idesc_t, ks > 0 ? 1u : 0u);
        umma_commit(bar_dq);
      }
      mbar_wait(bar_dq, phase);
      tc_fence_after();
      // ---- write dq (row = query), dk / dv (row = key) of this token
      uint32_t g[32];
      __nv_bfloat16* dst = p.dqkv + pix * 3 * C + head * 32;
#pragma unroll
      for (int which = 0; which < 3; ++which) {
        tmem_ld_32x32(taddr + (which == 0 ? kColS : (which == 1 ? kColDK : kColDV)), g);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(g[c * 8 + 0]), __uint_as_float(g[c * 8 + 1]));
            w.y = pack_bf16x2(__uint_as_float(g[c * 8 + 2]), __uint_as_float(g[c * 8 + 3]));
            w.z = pack_bf16x2(__uint_as_float(g[c * 8 + 4]), __uint_as_float(g[c * 8 + 5]));
            w.w = pack_bf16x2(__uint_as_float(g[c * 8 + 6]), __uint_as_float(g[c * 8 + 7]));
            *reinterpret_cast<uint4*>(dst + which * C + c * 8) = w;
          }
        }
      }
      tc_fence_before();
    }
    __syncthreads();
    tc_fence_after();
  }
  if (valid) {
    float* dbp = p.dbias + (static_cast<long long>(head) * kWT + row) * kWT;
#pragma unroll
    for (int j = 0; j < kWT; ++j) atomicAdd(dbp + j, db[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 4) {
    tc_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}

// bias[h][i][j] = table[index[i][j]][h]      (WindowAttention.forward :131-134)
__global__ void wattn_bias_gather_kernel(const float* __restrict__ table, const long long* __restrict__ index,
                                         float* __restrict__ bias, int nH) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nH * kWT * kWT) return;
  const int h = i / (kWT * kWT), ij = i - h * kWT * kWT;
  bias[i] = table[index[ij] * nH + h];
}
// dtable[index[i][j]][h] (+)= dbias[h][i][j]   (dtable zeroed / holding the running gradient)
__global__ void wattn_bias_scatter_kernel(const float* __restrict__ dbias, const long long* __restrict__ index,
                                          float* __restrict__ dtable, int nH) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nH * kWT * kWT) return;
  const int h = i / (kWT * kWT), ij = i - h * kWT * kWT;
  atomicAdd(dtable + index[ij] * nH + h, dbias[i]);
}

// ------------------------------------------------------------------------------------------------------------------
// Stand-alone permutations of kernels/window_process (the reference's only first-party CUDA; B2 seam), 16-byte vectors:
//   partition: out[b*nW + win][wy][wx][:] = in[b][(wh*ws + wy - shift) mod H][(ww*ws + wx - shift) mod W][:]
//   merge:     out[b][h][w][:] = in[b*nW + win(h', w')][h' % ws][w' % ws][:],  (h', w') = ((h - shift) mod H, (w - shift) mod W)
__global__ void window_partition_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int H, int W,
                                        int cvec, int shift, int ws) {
  const int nWx = W / ws, nWy = H / ws;
  const long long total = static_cast<long long>(B) * H * W * cvec;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cvec);
    long long t = i / cvec;
    const int wx = static_cast<int>(t % ws);
    t /= ws;
    const int wy = static_cast<int>(t % ws);
    t /= ws;
    const int ww = static_cast<int>(t % nWx);
    t /= nWx;
    const int wh = static_cast<int>(t % nWy);
    const long long b = t / nWy;
    int y = (wh * ws + wy - shift) % H, x = (ww * ws + wx - shift) % W;
    if (y < 0) y += H;
    if (x < 0) x += W;
    out[i] = __ldg(in + ((b * H + y) * W + x) * cvec + c);
  }
}
__global__ void window_merge_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int H, int W, int cvec,
                                    int shift, int ws) {
  const int nWx = W / ws, nWy = H / ws;
  const long long total = static_cast<long long>(B) * H * W * cvec;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cvec);
    long long t = i / cvec;
    const int w = static_cast<int>(t % W);
    t /= W;
    const int h = static_cast<int>(t % H);
    const long long b = t / H;
    int y = (h - shift) % H, x = (w - shift) % W;
    if (y < 0) y += H;
    if (x < 0) x += W;
    const long long win = (b * nWy + y / ws) * nWx + x / ws;
    out[i] = __ldg(in + ((win * ws + y % ws) * ws + x % ws) * cvec + c);
  }
}

}  // namespace b200
