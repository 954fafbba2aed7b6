// ViT multi-head attention backward for sm_100a (T <= 256 tokens, head dim 64), the backward of attn_fwd_kernel
// (attention.cuh); replaces autograd through classification/vision_transformer/vit_model.py:97-108.
//
//   S  = Q K^T                    -> P = exp(scale*S - lse)            (recomputed from the saved log-sum-exp, never in HBM)
//   dP = dO V^T                   -> dS = scale * P * (dP - delta),  delta_i = sum_d dO[i,d] O[i,d] (attn_delta_kernel)
//   dV += P^T dO ,  dK += dS^T Q ,  dQ = dS K
//
// One CTA per (batch, head) walks the keys in blocks of 128 so that TWO CTAs fit on an SM and overlap each other's
// MMA / soft-max / TMA phases (the first version owned all 512 TMEM columns and 230 KB per CTA, its S -> P -> dP -> dS -> dQ
// chain was strictly serial with the tensor pipe 14 % busy: 386 us per ViT-B/16 layer at bs 256; this one: 281 us):
//
//   for key block j (128 keys):      K_j, V_j resident (16 KB each), dK_j / dV_j accumulate in TMEM (64 + 64 columns)
//     for query block mb (128 rows): S = Q K_j^T (128 columns) -> P -> dP = dO V_j^T -> dS (in place over P)
//                                    dV_j += P^T dO, dK_j += dS^T Q, dQ_part = dS K_j
//
//   TMEM: 128 (S / dP / dQ_part) + 64 + 64 = 256 columns;  shared memory: K, V, Q, dO 16 KB each + P/dS 32 KB = 96 KB.
//   dQ of a query block is the sum over the key blocks: block 0 stores its part (bf16) through the normal dQ path, later
//   blocks TMA-load that part back (same CTA, so program order + wait_group make it visible), add their fp32 accumulator
//   and store the sum - deterministic, no atomics, 25 KB of L2-hot extra traffic per (batch, head).
//
// P and dS live in shared memory as bf16 in the key-blocked 128B-swizzled layout, which serves both as a K-major A operand
// (dQ = dS K) and as an MN-major A operand (P^T dO, dS^T Q) without any transpose; K, V, Q, dO tiles are likewise consumed
// in place as K-major or MN-major B operands.  Warp 4 = TMA producer + MMA issuer (one elected thread), warps 0-3 / 5-8 =
// soft-max warps (two per TMEM lane quadrant, the pair splits the 128 key columns of every row in half).
#pragma once
#include "attention.cuh"

namespace b200 {

struct alignas(64) AttnBwdParams {
  CUtensorMap qkv_map;   // qkv (3*H*64, T, B), box (64, 128, 1): Q, K and V tiles
  CUtensorMap do_map;    // dO (H*64, T, B), box (64, 128, 1)
  CUtensorMap dqkv_map;  // dqkv (3*H*64, T, B), box (64, 128, 1): stores, and loads of the partial dQ
  int B, H, T, nblk;     // nblk = ceil(T / 128) query blocks = key blocks
  float scale, scale_log2e;
  const float* lse;      // [B][H][T]
  const float* delta;    // [B][H][T]
};

constexpr int kAttnBwdSmemBytes = 4 * 16384 + 32768 + 256 + 1024;

__global__ void __launch_bounds__(288, 2) attn_bwd_kernel(const __grid_constant__ AttnBwdParams p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = smem + 16384;
  uint8_t* sQ = smem + 32768;    // also the landing buffer of the partial dQ (Q is dead once bar_dq fires)
  uint8_t* sdO = smem + 49152;   // also the dQ staging buffer (as in the product kernel)
  uint8_t* sP = smem + 65536;    // [2 key blocks of 64][128 q rows][64 keys] bf16: P, then dS in place; dK_j | dV_j staging
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 98304);
  uint64_t* bar_kv = bars + 0;      // K_j / V_j landed
  uint64_t* bar_q = bars + 1;       // Q / dO of the current query block landed
  uint64_t* bar_s = bars + 2;       // S in TMEM
  uint64_t* bar_p = bars + 3;       // P in smem (8 warp arrivals)
  uint64_t* bar_dp = bars + 4;      // dP in TMEM, dV MMAs retired
  uint64_t* bar_ds = bars + 5;      // dS in smem (8 warp arrivals)
  uint64_t* bar_dq = bars + 6;      // dQ_part in TMEM, dK MMAs retired -> Q / dO / P buffers reusable
  uint64_t* bar_free = bars + 7;    // soft-max warps have drained dQ_part (8 warp arrivals)
  uint64_t* bar_kvfree = bars + 8;  // dK_j / dV_j drained and stored (8 warp arrivals): K / V / accumulators reusable
  uint64_t* bar_part = bars + 9;    // partial dQ landed in sQ
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x % p.H;
  const int b = blockIdx.x / p.H;
  const int HD = p.H * 64;

  if (warp_idx == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&p.qkv_map);
      tma_prefetch_desc(&p.do_map);
      tma_prefetch_desc(&p.dqkv_map);
      mbar_init(bar_kv, 1);
      mbar_init(bar_q, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, 8);
      mbar_init(bar_dp, 1);
      mbar_init(bar_ds, 8);
      mbar_init(bar_dq, 1);
      mbar_init(bar_free, 8);
      mbar_init(bar_kvfree, 8);
      mbar_init(bar_part, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<256>(tmem_ptr_smem);   // two CTAs per SM share the 512 columns
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  constexpr uint32_t kColS = 0, kColDV = 128, kColDK = 192;

  if (warp_idx == 4) {
    if (lane == 0) {
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), q_addr = smem_u32(sQ), do_addr = smem_u32(sdO);
      const uint32_t p_addr = smem_u32(sP);
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);   // [128 q] x [128 keys], K-major both
      constexpr uint32_t idesc_t = make_idesc_bf16(128, 64, 1, 1);    // A^T B, both operands MN-major
      constexpr uint32_t idesc_q = make_idesc_bf16(128, 64, 0, 1);    // dS (K-major) x K (MN-major)
      int it = 0;
      for (int j = 0; j < p.nblk; ++j) {
        if (j > 0) mbar_wait(bar_kvfree, (j - 1) & 1);
        mbar_expect_tx(bar_kv, 2 * 16384);
        tma_load_3d(sK, &p.qkv_map, bar_kv, HD + h * 64, j * 128, b);
        tma_load_3d(sV, &p.qkv_map, bar_kv, 2 * HD + h * 64, j * 128, b);
        for (int mb = 0; mb < p.nblk; ++mb, ++it) {
          const uint32_t ph = it & 1;
          if (it > 0) mbar_wait(bar_free, (it - 1) & 1);   // previous dQ_part drained: S columns, Q / dO / P reusable
          mbar_expect_tx(bar_q, 2 * 16384);
          tma_load_3d(sQ, &p.qkv_map, bar_q, h * 64, mb * 128, b);
          tma_load_3d(sdO, &p.do_map, bar_q, h * 64, mb * 128, b);
          if (mb == 0) mbar_wait(bar_kv, j & 1);
          mbar_wait(bar_q, ph);
          tc_fence_after();
          // S = Q K_j^T
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + kColS, make_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                     make_smem_desc_sw128(k_addr + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
          umma_commit(bar_s);
          // P ready (S consumed): dP = dO V_j^T into the same columns, dV_j += P^T dO
          mbar_wait(bar_p, ph);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + kColS, make_smem_desc_sw128(do_addr + k * 32, 16, 1024),
                     make_smem_desc_sw128(v_addr + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
          for (int ks = 0; ks < 8; ++ks)   // 128 queries = 8 steps of 16
            umma_f16(tmem_base + kColDV, make_smem_desc_sw128(p_addr + ks * 2048, 16384, 1024),
                     make_smem_desc_sw128(do_addr + ks * 2048, 8192, 1024), idesc_t, (mb > 0 || ks > 0) ? 1u : 0u);
          umma_commit(bar_dp);
          // dS ready: dQ_part = dS K_j, dK_j += dS^T Q
          mbar_wait(bar_ds, ph);
          tc_fence_after();
          for (int ks = 0; ks < 8; ++ks)   // 128 keys = 8 steps of 16
            umma_f16(tmem_base + kColS, make_smem_desc_sw128(p_addr + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                     make_smem_desc_sw128(k_addr + ks * 2048, 8192, 1024), idesc_q, ks > 0 ? 1u : 0u);
          for (int ks = 0; ks < 8; ++ks)
            umma_f16(tmem_base + kColDK, make_smem_desc_sw128(p_addr + ks * 2048, 16384, 1024),
                     make_smem_desc_sw128(q_addr + ks * 2048, 8192, 1024), idesc_t, (mb > 0 || ks > 0) ? 1u : 0u);
          umma_commit(bar_dq);
        }
      }
    }
  } else {
    // 8 soft-max warps: two per TMEM lane quadrant; the pair splits the 128 key columns of every row in half
    const int quad = warp_idx & 3;
    const int pair = warp_idx > 4 ? 1 : 0;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const long long bh = static_cast<long long>(b) * p.H + h;
    int it = 0;
    uint32_t part_phase = 0;
    for (int j = 0; j < p.nblk; ++j) {
      for (int mb = 0; mb < p.nblk; ++mb, ++it) {
        const uint32_t ph = it & 1;
        const int t = mb * 128 + row;
        const bool valid = t < p.T;
        const float lse2 = valid ? p.lse[bh * p.T + t] * 1.4426950408889634f : INFINITY;
        const float delta = valid ? p.delta[bh * p.T + t] : 0.f;
        // ---- P = exp2(S*scale*log2e - lse*log2e), keys >= T masked
        mbar_wait(bar_s, ph);
        tc_fence_after();
#pragma unroll 1
        for (int c = pair * 2; c < pair * 2 + 2; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(lane_addr + kColS + c * 32, v);
          tmem_ld_wait();
          const int key0 = j * 128 + c * 32;            // first key of this chunk
          const bool crosses = key0 + 32 > p.T;         // warp-uniform
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float e[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) e[i] = attn_ex2(fmaf(__uint_as_float(v[g * 8 + i]), p.scale_log2e, -lse2));
            if (crosses) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (key0 + g * 8 + i >= p.T) e[i] = 0.f;
            }
            const int col = c * 32 + g * 8;
            *reinterpret_cast<uint4*>(sP + (col >> 6) * 16384 + row * 128 + ((((col & 63) >> 3) ^ (row & 7)) << 4)) = pack8(e);
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p);
        // ---- dS = scale * P * (dP - delta), in place (the dV MMAs that read P have retired when bar_dp fires)
        mbar_wait(bar_dp, ph);
        tc_fence_after();
#pragma unroll 1
        for (int c = pair * 2; c < pair * 2 + 2; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(lane_addr + kColS + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = c * 32 + g * 8;
            uint4* slot = reinterpret_cast<uint4*>(sP + (col >> 6) * 16384 + row * 128 + ((((col & 63) >> 3) ^ (row & 7)) << 4));
            float pv[8], e[8];
            unpack8(*slot, pv);
#pragma unroll
            for (int i = 0; i < 8; ++i) e[i] = p.scale * pv[i] * (__uint_as_float(v[g * 8 + i]) - delta);
            *slot = pack8(e);
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_ds);
        // ---- dQ_part (+ the parts of the previous key blocks) -> bf16 -> staging -> TMA store
        mbar_wait(bar_dq, ph);
        tc_fence_after();
        if (j > 0) {
          if (threadIdx.x == 0) {
            tma_store_wait_all<0>();   // this thread's earlier dQ stores are complete (not just read) before the reload
            mbar_expect_tx(bar_part, 16384);
            tma_load_3d(sQ, &p.dqkv_map, bar_part, h * 64, mb * 128, b);
          }
          mbar_wait(bar_part, part_phase);
          part_phase ^= 1;
        }
        uint8_t* stg = sdO;
        {
          const int c = pair;   // each warp of the pair drains one 32-column half of dQ_part
          uint32_t v[32];
          tmem_ld_32x32(lane_addr + kColS + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int off = row * 128 + (((c * 4 + g) ^ (row & 7)) << 4);
            float e[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) e[i] = __uint_as_float(v[g * 8 + i]);
            if (j > 0) {
              float prev[8];
              unpack8(*reinterpret_cast<const uint4*>(sQ + off), prev);
#pragma unroll
              for (int i = 0; i < 8; ++i) e[i] += prev[i];
            }
            *reinterpret_cast<uint4*>(stg + off) = pack8(e);
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        named_bar_sync(1, 256);
        if (threadIdx.x == 0) {
          tma_store_3d(&p.dqkv_map, stg, h * 64, mb * 128, b);
          tma_store_commit();
          tma_store_wait_read<0>();
        }
        named_bar_sync(1, 256);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_free);
      }
      // ---- dK_j, dV_j (TMEM lane = key) staged through the P region, which is free after the last bar_dq of this block
      for (int which = 0; which < 2; ++which) {   // 0: dK, 1: dV
        uint8_t* stg = sP + which * 16384;
        const uint32_t col0 = which == 0 ? kColDK : kColDV;
        const int c = pair;
        uint32_t v[32];
        tmem_ld_32x32(lane_addr + col0 + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) e[i] = __uint_as_float(v[g * 8 + i]);
          *reinterpret_cast<uint4*>(stg + row * 128 + (((c * 4 + g) ^ (row & 7)) << 4)) = pack8(e);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      named_bar_sync(1, 256);
      if (threadIdx.x == 0) {
        tma_store_3d(&p.dqkv_map, sP, HD + h * 64, j * 128, b);
        tma_store_3d(&p.dqkv_map, sP + 16384, 2 * HD + h * 64, j * 128, b);
        tma_store_commit();
        tma_store_wait_read<0>();
      }
      named_bar_sync(1, 256);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_kvfree);
    }
    if (threadIdx.x == 0) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 4) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace b200
