// Implicit-GEMM convolution / linear forward kernel for sm_100a.
//
//   D[pixel, n] = sum_{tap, c} A_tap[pixel (+tap offset), c] * Wt[n, tap*Cin + c]        (bf16 in, fp32 accumulate)
//
// * A (activations, NHWC bf16) is fetched by TMA as 4-D boxes (64 channels x b1 x b2 x b3 pixels = 128 rows of 128 B,
//   128B-swizzled) straight into the layout tcgen05.mma consumes (K-major). A filter tap is just a coordinate offset;
//   out-of-image pixels are zero-filled by the TMA unit, which is the convolution's zero padding. Stride-2 convolutions
//   pass up to four "phase" views (even/odd rows x cols) of the input as separate tensor maps.
// * B (weights, [Cout][taps*Cin] bf16) is a 2-D TMA box of BLOCK_N rows x 64 k.
// * One elected thread issues tcgen05.mma (M=128, N=BLOCK_N, K=16) into a double-buffered TMEM accumulator, so the
//   epilogue of tile i overlaps the main loop of tile i+1. The kernel is persistent: grid = min(tiles, #SM).
// * Epilogue warps: tcgen05.ld -> (+bias, activation, +residual) -> bf16 -> swizzled smem staging -> TMA store, plus
//   optional per-channel sum / sum-of-squares partials (train-mode BatchNorm statistics) per 128-row tile.
//
// Replaces the cuDNN/cuBLAS calls behind nn.Conv2d / nn.Linear on the reference's hot path
// (classification/resnet/models/networks.py:27-35,104-124; classification/vision_transformer/vit_model.py:95,109,127-133).
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int kMaxTaps = 16;

struct alignas(64) ConvGemmParams {
  CUtensorMap a_maps[4];
  CUtensorMap b_map;
  CUtensorMap d_map;
  int num_taps;
  int k_per_tap;         // Cin (elements of K per tap)
  int k_blocks_per_tap;  // ceil(Cin / 64)
  int tiles1, tiles2, tiles3;
  int box1, box2, box3;
  int dim1, dim2, dim3;  // output pixel extents (for row->address mapping of residual / fp32 output)
  int n_tiles;
  int N;
  int8_t tap_map[kMaxTaps];  // which activation view (phase) the tap reads
  int8_t tap_o1[kMaxTaps];   // pixel offset along dim1 (w)
  int8_t tap_o2[kMaxTaps];   // pixel offset along dim2 (h)
  int8_t tap_w[kMaxTaps];    // which k_per_tap-wide slice of the weight matrix the tap multiplies
  float* stats;             // [m_tiles][2][N] partial sums, or null
  const float* bias;        // [N] or null
  int act;                  // 0 none, 1 relu, 2 gelu(erf)
  const __nv_bfloat16* residual;  // bf16 tensor added in the epilogue (pixel strides rs1..rs3, in elements) or null
  long long rs1, rs2, rs3;
  uint32_t desc_lbo, desc_sbo;  // K-major smem descriptor strides (bytes): 16 / 1024
  float* out_f32;           // direct fp32 output ([pixels][ld_out]) or null -> bf16 TMA store
  long long ld_out;
};

template <int BLOCK_N>
struct ConvGemmCfg {
  static constexpr int BLOCK_M = 128;
  static constexpr int BLOCK_K = 64;
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGES = (BLOCK_N == 256) ? 3 : (BLOCK_N == 128 ? 5 : 6);
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGING_BYTES = 128 * 128;  // one 128x64 bf16 chunk
  static constexpr int STATS_BYTES = 4 * 2 * 64 * 4;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGING_BYTES + STATS_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512);
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

template <int BLOCK_N>
__global__ void __launch_bounds__(192, 1) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  using Cfg = ConvGemmCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint8_t* staging = smem + STAGES * Cfg::STAGE_BYTES;
  float* stats_smem = reinterpret_cast<float*>(staging + 2 * Cfg::STAGING_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stats_smem) + Cfg::STATS_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = p.tiles1 * p.tiles2 * p.tiles3;
  const int num_tiles = m_tiles * p.n_tiles;
  const int num_kb = p.num_taps * p.k_blocks_per_tap;

  if (warp_idx == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.a_maps[i]);
    tma_prefetch_desc(&p.b_map);
    tma_prefetch_desc(&p.d_map);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp_idx == 1) {
    tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr_smem);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        const int m_tile = tile / p.n_tiles;
        const int t1 = m_tile % p.tiles1;
        const int t2 = (m_tile / p.tiles1) % p.tiles2;
        const int t3 = m_tile / (p.tiles1 * p.tiles2);
        const int c1 = t1 * p.box1, c2 = t2 * p.box2, c3 = t3 * p.box3;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / p.k_blocks_per_tap;
          const int cb = kb - tap * p.k_blocks_per_tap;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = stage_base + stage * Cfg::STAGE_BYTES;
          uint8_t* b_dst = a_dst + Cfg::A_BYTES;
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_4d(a_dst, &p.a_maps[p.tap_map[tap]], &full_bar[stage], cb * 64, c1 + p.tap_o1[tap],
                      c2 + p.tap_o2[tap], c3);
          tma_load_2d(b_dst, &p.b_map, &full_bar[stage], p.tap_w[tap] * p.k_per_tap + cb * 64, n_tile * BLOCK_N);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(stage_base + stage * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = make_smem_desc_sw128(a_addr + k * 32, p.desc_lbo, p.desc_sbo);
            const uint64_t db = make_smem_desc_sw128(b_addr + k * 32, p.desc_lbo, p.desc_sbo);
            umma_f16(tmem_d, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== Epilogue (4 warps, 128 threads) =====================
    const int q = warp_idx & 3;          // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;       // row of the 128-row tile owned by this thread
    const int epi_tid = threadIdx.x - 64;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t store_counter = 0;  // counts issued TMA stores (selects the staging buffer)
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles;
      const int m_tile = tile / p.n_tiles;
      const int t1 = m_tile % p.tiles1;
      const int t2 = (m_tile / p.tiles1) % p.tiles2;
      const int t3 = m_tile / (p.tiles1 * p.tiles2);
      const int c1 = t1 * p.box1, c2 = t2 * p.box2, c3 = t3 * p.box3;
      // Row -> pixel mapping (used by residual / fp32 output paths).
      const int i1 = row % p.box1;
      const int i2 = (row / p.box1) % p.box2;
      const int i3 = row / (p.box1 * p.box2);
      const int p1 = c1 + i1, p2 = c2 + i2, p3 = c3 + i3;
      const bool row_ok = (p1 < p.dim1) && (p2 < p.dim2) && (p3 < p.dim3);
      const long long pix = (static_cast<long long>(p3) * p.dim2 + p2) * p.dim1 + p1;

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(q * 32) << 16);

#pragma unroll 1
      for (int ch = 0; ch < BLOCK_N / 64; ++ch) {
        const int n0 = n_tile * BLOCK_N + ch * 64;
        uint32_t v[2][32];
        tmem_ld_32x32(tmem_acc + ch * 64, v[0]);
        tmem_ld_32x32(tmem_acc + ch * 64 + 32, v[1]);
        tmem_ld_wait();
        if (ch == BLOCK_N / 64 - 1) {
          // all TMEM reads of this accumulator are done -> hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        if (n0 >= p.N) continue;  // fully out-of-range column chunk (N not a multiple of BLOCK_N)

        float f[64];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          f[j] = __uint_as_float(v[0][j]);
          f[32 + j] = __uint_as_float(v[1][j]);
        }
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 64; ++j) {
            const int n = n0 + j;
            f[j] += (n < p.N) ? __ldg(p.bias + n) : 0.0f;
          }
        }
        if (p.act == 1) {
#pragma unroll
          for (int j = 0; j < 64; ++j) f[j] = fmaxf(f[j], 0.0f);
        } else if (p.act == 2) {
#pragma unroll
          for (int j = 0; j < 64; ++j) f[j] = gelu_erf(f[j]);
        }
        if (p.residual != nullptr && row_ok) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.residual + p3 * p.rs3 + p2 * p.rs2 + p1 * p.rs1 + n0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (n0 + j * 8 < p.N) {
              const uint4 r = __ldg(rp + j);
              f[j * 8 + 0] += bf16_lo(r.x);
              f[j * 8 + 1] += bf16_hi(r.x);
              f[j * 8 + 2] += bf16_lo(r.y);
              f[j * 8 + 3] += bf16_hi(r.y);
              f[j * 8 + 4] += bf16_lo(r.z);
              f[j * 8 + 5] += bf16_hi(r.z);
              f[j * 8 + 6] += bf16_lo(r.w);
              f[j * 8 + 7] += bf16_hi(r.w);
            }
          }
        }

        if (p.out_f32 != nullptr) {
          if (row_ok) {
            float* op = p.out_f32 + pix * p.ld_out + n0;
#pragma unroll
            for (int j = 0; j < 64; ++j)
              if (n0 + j < p.N) op[j] = f[j];
          }
          continue;
        }

        if (!row_ok) {
          // rows of a partial pixel box: clipped by the TMA store, and must not pollute the BN statistics
#pragma unroll
          for (int j = 0; j < 64; ++j) f[j] = 0.f;
        }
        // ---- bf16 path: registers -> swizzled staging -> TMA store
        uint8_t* buf = staging + (store_counter & 1) * Cfg::STAGING_BYTES;
        ++store_counter;
        if (epi_tid == 0) tma_store_wait_read<1>();  // the store that used this buffer two chunks ago has drained
        named_bar_sync(1, 128);
        {
          uint8_t* rowp = buf + row * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 w;
            w.x = pack_bf16x2(f[j * 8 + 0], f[j * 8 + 1]);
            w.y = pack_bf16x2(f[j * 8 + 2], f[j * 8 + 3]);
            w.z = pack_bf16x2(f[j * 8 + 4], f[j * 8 + 5]);
            w.w = pack_bf16x2(f[j * 8 + 6], f[j * 8 + 7]);
            *reinterpret_cast<uint4*>(rowp + ((j ^ (row & 7)) << 4)) = w;
          }
        }
        if (p.stats != nullptr) {
          // Column sums over this warp's 32 rows, read back from the (bf16-rounded) staging tile:
          // lane l owns columns 2l, 2l+1; bank-conflict-free thanks to the 128B swizzle.
          __syncwarp();
          float s0 = 0.f, s1 = 0.f, ss0 = 0.f, ss1 = 0.f;
#pragma unroll 8
          for (int r = 0; r < 32; ++r) {
            const int rr = q * 32 + r;
            const uint32_t w =
                *reinterpret_cast<const uint32_t*>(buf + rr * 128 + ((((lane >> 2) ^ (rr & 7)) << 4) | ((lane & 3) << 2)));
            const float a = bf16_lo(w), b = bf16_hi(w);
            s0 += a;
            s1 += b;
            ss0 = fmaf(a, a, ss0);
            ss1 = fmaf(b, b, ss1);
          }
          float* sp = stats_smem + q * 128;  // [q][2][64]
          sp[2 * lane] = s0;
          sp[2 * lane + 1] = s1;
          sp[64 + 2 * lane] = ss0;
          sp[64 + 2 * lane + 1] = ss1;
        }
        fence_proxy_async_smem();
        named_bar_sync(2, 128);
        if (epi_tid == 0) {
          tma_store_4d(&p.d_map, buf, n0, c1, c2, c3);
          tma_store_commit();
        }
        if (p.stats != nullptr) {
          const int which = epi_tid >> 6;  // 0 = sum, 1 = sum of squares
          const int col = epi_tid & 63;
          if (n0 + col < p.N) {
            const float t = stats_smem[0 * 128 + which * 64 + col] + stats_smem[1 * 128 + which * 64 + col] +
                            stats_smem[2 * 128 + which * 64 + col] + stats_smem[3 * 128 + which * 64 + col];
            p.stats[(static_cast<long long>(m_tile) * 2 + which) * p.N + n0 + col] = t;
          }
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (epi_tid == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

}  // namespace b200
