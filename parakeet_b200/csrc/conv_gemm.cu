// pk_conv_gemm: channels-last Conv1D / Linear / batched matmul as an im2col-free tiled GEMM on tcgen05.
//
//   - persistent CTAs (one per SM) walking 128(time) x BLOCK_N(channel) output tiles, two TMEM accumulators so that the
//     epilogue of tile i overlaps the main loop of tile i+1; 192 threads:
//       warp 0   : TMA producer  (one elected lane)
//       warp 1   : TMEM allocator + tcgen05.mma issuer (one elected lane)
//       warps 2-5: epilogue (TMEM -> registers -> bias/act/residual/mask -> global), one output row per thread
//   - K loop over (tap, 64-channel chunk): each conv tap is just the same A tensor read `(tap - pad) * dil` rows
//     further along; TMA zero-fills rows outside the utterance, which is the conv's zero padding (no im2col).
//   - split-bf16 operands, 3 MMAs per K-step (hi*hi + lo*hi + hi*lo), fp32 accumulation in TMEM.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "pk_host.h"
#include "pk_sm100.cuh"

namespace pk {

constexpr int kBlockM = 128;
constexpr int kGemmThreads = 192;

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kABytes = kBlockM * kSwizzleBytes;       // one plane of one A chunk (16 KB)
  static constexpr int kBBytes = BLOCK_N * kSwizzleBytes;       // one plane of one B chunk
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;  // hi + lo
  static constexpr int kStages = (BLOCK_N >= 256) ? 2 : (BLOCK_N >= 128 ? 3 : 4);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr uint32_t kTmemCols = BLOCK_N < 32 ? 32 : BLOCK_N;
};

struct GemmKernelArgs {
  int m, n, k_chunks, taps, dil, pad, heads, batch;
  int tiles_m, total_tiles;       // persistent schedule (set by launch<>)
  int a_bmul, a_hmul, a_col0, a_colh;
  int b_bmul, b_hmul, b_col0, b_colh, b_tap_stride;
  float scale;
  const float* bias;
  int act;
  const float* residual;
  const int32_t* lens;
  float* y_f32;
  __nv_bfloat16* y_hi;
  __nv_bfloat16* y_lo;
  long long y_batch_stride, y_head_stride;
  int y_ld;
  int passes;
  // fused pair epilogues (pk_conv_gemm_ex): the tile covers n = 2 * epi_c columns, column c is paired with column epi_c + c
  int epi, epi_c;
  const float* e_res;            // GATE: fp32 (batch, m, e_res_ld) added before the gate, or NULL
  long long e_res_bs;
  int e_res_ld;
  float* e_state;                // WF_UPDATE: fp32 (batch, m, epi_c) running state / skip sum
  float* e_skip;
  int e_skip_init;
  __nv_bfloat16* e_buf_hi;       // WF_UPDATE: optional split planes (batch, m, e_buf_ld) receiving the new state at e_buf_col0
  __nv_bfloat16* e_buf_lo;
  int e_buf_ld, e_buf_col0;
};

// bias / activation / residual / row mask / stores for one 32-column chunk of one output row (v: the accumulators)
__device__ __forceinline__ void gemm_epilogue_chunk(const GemmKernelArgs& p, float (&v)[32], const int nb, const bool row_ok,
                                                    const bool row_live, const long long y_off) {
  if (row_ok && nb < p.n) {
    // bias: one batch of independent loads per 32-column chunk (a dependent load per element would serialise the
    // epilogue on global-memory latency), activation selected outside the element loops
    float bv[32];
    if (p.bias == nullptr) {
#pragma unroll
      for (int j = 0; j < 32; ++j) bv[j] = 0.f;
    } else if (nb + 32 <= p.n && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + nb);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b = __ldg(b4 + j);
        bv[4 * j] = b.x; bv[4 * j + 1] = b.y; bv[4 * j + 2] = b.z; bv[4 * j + 3] = b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) bv[j] = nb + j < p.n ? __ldg(p.bias + nb + j) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], p.scale, bv[j]);
    if (p.act == PK_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (p.act == PK_ACT_TANH) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
    }
    const bool full = (nb + 32 <= p.n) && ((p.y_ld & 7) == 0) && (((y_off + nb) & 7) == 0);
    if (p.residual != nullptr) {
      if (full) {
        const float4* r4 = reinterpret_cast<const float4*>(p.residual + y_off + nb);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 r = __ldg(r4 + j);
          v[4 * j + 0] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (nb + j < p.n) v[j] += __ldg(p.residual + y_off + nb + j);
      }
    }
    if (!row_live) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
    }
    if (full) {
      if (p.y_f32 != nullptr) {
        float4* o4 = reinterpret_cast<float4*>(p.y_f32 + y_off + nb);
#pragma unroll
        for (int j = 0; j < 8; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      if (p.y_hi != nullptr) {
        uint4* oh = reinterpret_cast<uint4*>(p.y_hi + y_off + nb);
        uint4* ol = reinterpret_cast<uint4*>(p.y_lo + y_off + nb);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 h, l;
          split8(v + 8 * j, h, l);
          oh[j] = h;
          ol[j] = l;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (nb + j < p.n) {
          if (p.y_f32 != nullptr) p.y_f32[y_off + nb + j] = v[j];
          if (p.y_hi != nullptr) {
            __nv_bfloat16 h, l;
            split_bf16(v[j], h, l);
            p.y_hi[y_off + nb + j] = h;
            p.y_lo[y_off + nb + j] = l;
          }
        }
      }
    }
  }
}

// Fused pair epilogues: one 32-column chunk of the first half (va: columns [nb, nb+32) of [0, C)) together with the
// matching chunk of the second half (vg: columns C + [nb, nb+32)); bias / scale applied to both.
//   PK_EPI_GATE      z = tanh(a + res_a) * sigmoid(g + res_g) -> split planes (batch, m, y_ld) at column nb
//                    (ResidualBlock gate of waveflow.py:277-281 fused into the dilated-conv GEMM)
//   PK_EPI_WF_UPDATE state += a; skip (=|+=) g; new state -> optional split planes
//                    (waveflow.py:282-294 `res, skip = split(out_proj(z))`, ResidualNet.add_input :386-392)
__device__ __forceinline__ void gemm_epilogue_pair(const GemmKernelArgs& p, float (&va)[32], float (&vg)[32], const int nb,
                                                   const int bz, const int row, const bool row_ok) {
  if (!row_ok) return;
  const int C = p.epi_c;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bg = ba;
    if (p.bias != nullptr) {
      ba = __ldg(reinterpret_cast<const float4*>(p.bias + nb) + j);
      bg = __ldg(reinterpret_cast<const float4*>(p.bias + C + nb) + j);
    }
    va[4 * j] = fmaf(va[4 * j], p.scale, ba.x); va[4 * j + 1] = fmaf(va[4 * j + 1], p.scale, ba.y);
    va[4 * j + 2] = fmaf(va[4 * j + 2], p.scale, ba.z); va[4 * j + 3] = fmaf(va[4 * j + 3], p.scale, ba.w);
    vg[4 * j] = fmaf(vg[4 * j], p.scale, bg.x); vg[4 * j + 1] = fmaf(vg[4 * j + 1], p.scale, bg.y);
    vg[4 * j + 2] = fmaf(vg[4 * j + 2], p.scale, bg.z); vg[4 * j + 3] = fmaf(vg[4 * j + 3], p.scale, bg.w);
  }
  const long long grow = static_cast<long long>(bz) * p.m + row;       // row index in (batch, m, .) tensors
  if (p.epi == PK_EPI_GATE) {
    if (p.e_res != nullptr) {
      const float4* ra = reinterpret_cast<const float4*>(p.e_res + bz * p.e_res_bs + static_cast<long long>(row) * p.e_res_ld + nb);
      const float4* rg = reinterpret_cast<const float4*>(p.e_res + bz * p.e_res_bs + static_cast<long long>(row) * p.e_res_ld + C + nb);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 x = __ldg(ra + j), y = __ldg(rg + j);
        va[4 * j] += x.x; va[4 * j + 1] += x.y; va[4 * j + 2] += x.z; va[4 * j + 3] += x.w;
        vg[4 * j] += y.x; vg[4 * j + 1] += y.y; vg[4 * j + 2] += y.z; vg[4 * j + 3] += y.w;
      }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) va[j] = tanhf(va[j]) * (1.f / (1.f + expf(-vg[j])));
    uint4* oh = reinterpret_cast<uint4*>(p.y_hi + grow * p.y_ld + nb);
    uint4* ol = reinterpret_cast<uint4*>(p.y_lo + grow * p.y_ld + nb);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 h, l;
      split8(va + 8 * j, h, l);
      oh[j] = h;
      ol[j] = l;
    }
  } else {   // PK_EPI_WF_UPDATE
    float4* st4 = reinterpret_cast<float4*>(p.e_state + grow * C + nb);
    float4* sk4 = reinterpret_cast<float4*>(p.e_skip + grow * C + nb);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 s = st4[j];
      va[4 * j] += s.x; va[4 * j + 1] += s.y; va[4 * j + 2] += s.z; va[4 * j + 3] += s.w;
      st4[j] = make_float4(va[4 * j], va[4 * j + 1], va[4 * j + 2], va[4 * j + 3]);
      float4 k = make_float4(vg[4 * j], vg[4 * j + 1], vg[4 * j + 2], vg[4 * j + 3]);
      if (!p.e_skip_init) {
        const float4 o = sk4[j];
        k.x += o.x; k.y += o.y; k.z += o.z; k.w += o.w;
      }
      sk4[j] = k;
    }
    if (p.e_buf_hi != nullptr) {
      uint4* oh = reinterpret_cast<uint4*>(p.e_buf_hi + grow * p.e_buf_ld + p.e_buf_col0 + nb);
      uint4* ol = reinterpret_cast<uint4*>(p.e_buf_lo + grow * p.e_buf_ld + p.e_buf_col0 + nb);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 h, l;
        split8(va + 8 * j, h, l);
        oh[j] = h;
        ol[j] = l;
      }
    }
  }
}

struct GemmTile {     // persistent tile schedule: (batch, head) fastest, then m-tile, then n-tile, so that the CTAs running
  int m0, n0, bz, hz; // at the same time share one n-tile of B (the weights stay hot in L2) AND the dead tiles of a ragged
};                    // batch (all utterances' m-tile k for large k) are consecutive indices, i.e. spread evenly over the
                      // grid-strided CTAs (with m fastest and 148 % tiles_m == 0 some CTAs would own only dead tiles)
__device__ __forceinline__ GemmTile gemm_tile(int tile, int tiles_m, int zdim, int heads, int block_n) {
  GemmTile t;
  const int z = tile % zdim;
  const int r = tile / zdim;
  const int mt = r % tiles_m;
  t.m0 = mt * kBlockM;
  t.n0 = (r / tiles_m) * block_n;
  t.bz = z / heads;
  t.hz = z % heads;
  return t;
}

// Ragged batches (`lens`): an m-tile that starts at or past its utterance's length holds no live row.  Every role skips it
// with the same test - no loads, no MMAs, no accumulator hand-over - and the epilogue warps just write its zero rows (the rows
// are the utterance's own zero padding for the next conv).  On LJSpeech-shaped batches (T ~ U{60..140}, padded to the longest)
// a third of the m-tiles are dead.
__device__ __forceinline__ bool gemm_tile_live(const GemmKernelArgs& p, int m0, int bz) {
  return p.lens == nullptr || m0 < __ldg(p.lens + bz);
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                 const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                 const GemmKernelArgs p) {
  using Cfg = GemmCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-B alignment for SWIZZLE_128B tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* acc_full = empty_bar + Cfg::kStages;     // [2] MMA issuer -> epilogue: accumulator of tile i is complete
  uint64_t* acc_empty = acc_full + 2;                // [2] epilogue -> MMA issuer: accumulator buffer drained
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_chunks = p.taps * p.k_chunks;
  const int zdim = p.batch * p.heads;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a_hi);
    tma_prefetch_desc(&tm_a_lo);
    tma_prefetch_desc(&tm_b_hi);
    tma_prefetch_desc(&tm_b_lo);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<2 * Cfg::kTmemCols>(tmem_base_slot);   // two accumulators: epilogue(i) overlaps mainloop(i+1)
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------ TMA producer ------------------------------
      const uint32_t tx_bytes = (p.passes == 3) ? Cfg::kStageBytes : (Cfg::kABytes + Cfg::kBBytes);
      uint32_t it = 0;                                 // running stage counter across tiles
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const GemmTile t = gemm_tile(tile, p.tiles_m, zdim, p.heads, BLOCK_N);
        if (!gemm_tile_live(p, t.m0, t.bz)) continue;
        const int a_batch = t.bz * p.a_bmul + t.hz * p.a_hmul;
        const int b_batch = t.bz * p.b_bmul + t.hz * p.b_hmul;
        const int a_col = p.a_col0 + t.hz * p.a_colh;
        const int b_col = p.b_col0 + t.hz * p.b_colh;
        for (int i = 0; i < num_chunks; ++i, ++it) {
          const int s = it % Cfg::kStages;
          const uint32_t ph = (it / Cfg::kStages) & 1;
          const int tap = i / p.k_chunks;
          const int kc = i % p.k_chunks;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* st = smem + s * Cfg::kStageBytes;
          mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
          const int a_row = t.m0 + (tap - p.pad) * p.dil;
          tma_load_3d(st, &tm_a_hi, &full_bar[s], a_col + kc * kChunkK, a_row, a_batch);
          tma_load_3d(st + 2 * Cfg::kABytes, &tm_b_hi, &full_bar[s], b_col + tap * p.b_tap_stride + kc * kChunkK, t.n0, b_batch);
          if (p.passes == 3) {
            tma_load_3d(st + Cfg::kABytes, &tm_a_lo, &full_bar[s], a_col + kc * kChunkK, a_row, a_batch);
            tma_load_3d(st + 2 * Cfg::kABytes + Cfg::kBBytes, &tm_b_lo, &full_bar[s],
                        b_col + tap * p.b_tap_stride + kc * kChunkK, t.n0, b_batch);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------ MMA issuer ------------------------------
      constexpr uint32_t idesc = make_idesc_bf16_f32(kBlockM, BLOCK_N);
      uint32_t it = 0;
      int lt = 0;                                      // tiles processed by this CTA
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const GemmTile t = gemm_tile(tile, p.tiles_m, zdim, p.heads, BLOCK_N);
        if (!gemm_tile_live(p, t.m0, t.bz)) continue;
        const int buf = lt & 1;
        mbar_wait(&acc_empty[buf], ((lt >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + buf * Cfg::kTmemCols;
        for (int i = 0; i < num_chunks; ++i, ++it) {
          const int s = it % Cfg::kStages;
          const uint32_t ph = (it / Cfg::kStages) & 1;
          mbar_wait(&full_bar[s], ph);
          tcgen05_fence_after();
          const uint32_t st = smem_u32(smem + s * Cfg::kStageBytes);
          const uint64_t a_hi = make_smem_desc_sw128(st);
          const uint64_t a_lo = make_smem_desc_sw128(st + Cfg::kABytes);
          const uint64_t b_hi = make_smem_desc_sw128(st + 2 * Cfg::kABytes);
          const uint64_t b_lo = make_smem_desc_sw128(st + 2 * Cfg::kABytes + Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < kChunkK / kUmmaK; ++k) {
            const uint64_t koff = static_cast<uint64_t>((k * kUmmaK * 2) >> 4);  // 32 B per K-step inside the 128B row
            umma_bf16(d_tmem, a_hi + koff, b_hi + koff, idesc, (i | k) != 0);
            if (p.passes == 3) {
              umma_bf16(d_tmem, a_lo + koff, b_hi + koff, idesc, 1);
              umma_bf16(d_tmem, a_hi + koff, b_lo + koff, idesc, 1);
            }
          }
          umma_commit(&empty_bar[s]);  // frees this smem stage when the MMAs above have read it
        }
        umma_commit(&acc_full[buf]);   // accumulator complete
        ++lt;
      }
    }
  } else {
    // ------------------------------ epilogue ------------------------------
    const int quarter = warp & 3;               // TMEM lane quarter this warp may access
    int lt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const GemmTile t = gemm_tile(tile, p.tiles_m, zdim, p.heads, BLOCK_N);
      const int buf = lt & 1;
      const int row = t.m0 + quarter * 32 + lane;   // output time step
      const bool row_ok = row < p.m;
      const bool row_live = row_ok && (p.lens == nullptr || row < __ldg(p.lens + t.bz));
      const long long y_off = t.bz * p.y_batch_stride + t.hz * p.y_head_stride + static_cast<long long>(row) * p.y_ld;
      if (!gemm_tile_live(p, t.m0, t.bz)) {        // dead tile: zero rows, nothing to wait for
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
          gemm_epilogue_chunk(p, v, t.n0 + c * 32, row_ok, false, y_off);
        }
        continue;
      }
      ++lt;
      mbar_wait(&acc_full[buf], ((lt - 1) >> 1) & 1);
      tcgen05_fence_after();
      const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * Cfg::kTmemCols;
      if (p.epi == PK_EPI_NONE) {
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          float v[32];
          __syncwarp();
          tmem_ld_32x32(t_acc + c * 32, v);
          tmem_ld_wait();
          if (c == BLOCK_N / 32 - 1) {            // last read of this accumulator: hand it back to the MMA issuer
            tcgen05_fence_before();
            mbar_arrive(&acc_empty[buf]);
          }
          gemm_epilogue_chunk(p, v, t.n0 + c * 32, row_ok, row_live, y_off);
        }
      } else {
        // fused pair epilogues: the tile holds columns [0, 2C); chunk c of the first half with chunk c of the second half
        const int hc = p.epi_c / 32;
#pragma unroll 1
        for (int c = 0; c < hc; ++c) {
          float va[32], vg[32];
          __syncwarp();
          tmem_ld_32x32(t_acc + c * 32, va);
          tmem_ld_32x32(t_acc + p.epi_c + c * 32, vg);
          tmem_ld_wait();
          if (c == hc - 1) {
            tcgen05_fence_before();
            mbar_arrive(&acc_empty[buf]);
          }
          gemm_epilogue_pair(p, va, vg, c * 32, t.bz, row, row_ok);
        }
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<2 * Cfg::kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2, clusters of 2) for wide outputs: one M = 256 x BLOCK_N tile per pair.
// Each CTA loads its own 128 rows of A and HALF of the B tile (BLOCK_N / 2 rows), so the L2 -> SM operand stream per CTA
// drops from 96 KB to 64 KB per K-chunk at BLOCK_N = 256 (the single-CTA kernel needs 62 B/clk/SM there against a
// 42 B/clk/SM share of the L2 throughput cap) and three 64 KB stages fit instead of two 96 KB ones.
// Leader CTA (cluster rank 0): issues every MMA / commit, owns full[s] (TMA bytes of both CTAs) and acc_empty[2]
// (warp-elected relaxed remote arrivals); both CTAs: producer, epilogue, local empty[s] / acc_full[2] (multicast commits).
// ---------------------------------------------------------------------------------------------------------------
template <int BLOCK_N>
struct GemmPairCfg {
  static constexpr int kABytes = kBlockM * kSwizzleBytes;             // one plane of this CTA's A chunk (16 KB)
  static constexpr int kBBytes = (BLOCK_N / 2) * kSwizzleBytes;       // one plane of this CTA's half of the B chunk
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  static constexpr int kStages = 3;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
  static constexpr uint32_t kTmemCols = BLOCK_N <= 128 ? 128 : 256;   // accumulator stride (2 buffers: power-of-two allocation)
};

template <int BLOCK_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
conv_gemm_pair_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                      const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                      const GemmKernelArgs p) {
  using Cfg = GemmPairCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem + Cfg::kStages * Cfg::kStageBytes;
  const uint32_t full_bar = bars;                              // [stages] (the leader's copy is live)
  const uint32_t empty_bar = full_bar + 8 * Cfg::kStages;      // [stages]
  const uint32_t acc_full = empty_bar + 8 * Cfg::kStages;      // [2]
  const uint32_t acc_empty = acc_full + 16;                    // [2] leader
  const uint32_t tmem_slot = acc_empty + 16;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_chunks = p.taps * p.k_chunks;
  const int zdim = p.batch * p.heads;
  const int pair_id = blockIdx.x >> 1, pair_step = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a_hi);
    tma_prefetch_desc(&tm_a_lo);
    tma_prefetch_desc(&tm_b_hi);
    tma_prefetch_desc(&tm_b_lo);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init_a(full_bar + 8 * s, 1);
      mbar_init_a(empty_bar + 8 * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init_a(acc_full + 8 * i, 1);
      mbar_init_a(acc_empty + 8 * i, 2 * 4);                   // 4 epilogue warps in each CTA, one elected arrival each
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm_a<2 * Cfg::kTmemCols>(tmem_slot);
  tcgen05_fence_before();
  cluster_sync();
  tcgen05_fence_after();
  const uint32_t tmem_base = lds_u32(tmem_slot);

  // tile schedule of the pair: (batch, head) fastest, then m-pair-tile (256 rows), then n-tile; p.tiles_m counts pair tiles
  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------ TMA producer (both CTAs) ------------------------------
      const uint32_t full_leader = mapa_shared(full_bar, 0);
      const uint32_t tx_bytes = 2 * ((p.passes == 3) ? Cfg::kStageBytes : (Cfg::kABytes + Cfg::kBBytes));   // both CTAs
      uint32_t it = 0;
      for (int tile = pair_id; tile < p.total_tiles; tile += pair_step) {
        const int z = tile % zdim;
        const int r = tile / zdim;
        const int mt = r % p.tiles_m;
        const int m0 = mt * 2 * kBlockM + static_cast<int>(rank) * kBlockM;
        const int n0 = (r / p.tiles_m) * BLOCK_N + static_cast<int>(rank) * (BLOCK_N / 2);
        const int bz = z / p.heads, hz = z % p.heads;
        if (!gemm_tile_live(p, mt * 2 * kBlockM, bz)) continue;      // the PAIR tile (256 rows) is the unit that is skipped
        const int a_batch = bz * p.a_bmul + hz * p.a_hmul;
        const int b_batch = bz * p.b_bmul + hz * p.b_hmul;
        const int a_col = p.a_col0 + hz * p.a_colh;
        const int b_col = p.b_col0 + hz * p.b_colh;
        for (int i = 0; i < num_chunks; ++i, ++it) {
          const int s = it % Cfg::kStages;
          const uint32_t ph = (it / Cfg::kStages) & 1;
          const int tap = i / p.k_chunks;
          const int kc = i % p.k_chunks;
          mbar_wait_a(empty_bar + 8 * s, ph ^ 1);
          const uint32_t st = smem + s * Cfg::kStageBytes;
          const uint32_t fb = full_leader + 8 * s;
          if (leader) mbar_arrive_expect_tx_a(full_bar + 8 * s, tx_bytes);
          const int a_row = m0 + (tap - p.pad) * p.dil;
          const int bc = b_col + tap * p.b_tap_stride + kc * kChunkK;
          tma_load_3d_2sm_a(st, &tm_a_hi, fb, a_col + kc * kChunkK, a_row, a_batch);
          tma_load_3d_2sm_a(st + 2 * Cfg::kABytes, &tm_b_hi, fb, bc, n0, b_batch);
          if (p.passes == 3) {
            tma_load_3d_2sm_a(st + Cfg::kABytes, &tm_a_lo, fb, a_col + kc * kChunkK, a_row, a_batch);
            tma_load_3d_2sm_a(st + 2 * Cfg::kABytes + Cfg::kBBytes, &tm_b_lo, fb, bc, n0, b_batch);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ------------------------------ MMA issuer (leader CTA) ------------------------------
      constexpr uint32_t idesc = make_idesc_bf16_f32(2 * kBlockM, BLOCK_N);
      uint32_t it = 0;
      int lt = 0;
      for (int tile = pair_id; tile < p.total_tiles; tile += pair_step) {
        if (!gemm_tile_live(p, ((tile / zdim) % p.tiles_m) * 2 * kBlockM, (tile % zdim) / p.heads)) continue;
        const int buf = lt & 1;
        mbar_wait_a(acc_empty + 8 * buf, ((lt >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + buf * Cfg::kTmemCols;
        for (int i = 0; i < num_chunks; ++i, ++it) {
          const int s = it % Cfg::kStages;
          const uint32_t ph = (it / Cfg::kStages) & 1;
          mbar_wait_a(full_bar + 8 * s, ph);
          tcgen05_fence_after();
          const uint32_t st = smem + s * Cfg::kStageBytes;
          const uint64_t a_hi = make_smem_desc_sw128(st);
          const uint64_t a_lo = make_smem_desc_sw128(st + Cfg::kABytes);
          const uint64_t b_hi = make_smem_desc_sw128(st + 2 * Cfg::kABytes);
          const uint64_t b_lo = make_smem_desc_sw128(st + 2 * Cfg::kABytes + Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < kChunkK / kUmmaK; ++k) {
            const uint64_t koff = static_cast<uint64_t>((k * kUmmaK * 2) >> 4);
            umma_bf16_2sm(d_tmem, a_hi + koff, b_hi + koff, idesc, (i | k) != 0);
            if (p.passes == 3) {
              umma_bf16_2sm(d_tmem, a_lo + koff, b_hi + koff, idesc, 1);
              umma_bf16_2sm(d_tmem, a_hi + koff, b_lo + koff, idesc, 1);
            }
          }
          umma_commit_2sm_a(empty_bar + 8 * s);
        }
        umma_commit_2sm_a(acc_full + 8 * buf);
        ++lt;
      }
    }
  } else {
    // ------------------------------ epilogue (both CTAs, own 128 rows, all BLOCK_N columns) ------------------------------
    const int quarter = warp & 3;
    const uint32_t acc_empty_l = mapa_shared(acc_empty, 0);
    int lt = 0;
    for (int tile = pair_id; tile < p.total_tiles; tile += pair_step) {
      const int z = tile % zdim;
      const int r = tile / zdim;
      const int mt = r % p.tiles_m;
      const int tn0 = (r / p.tiles_m) * BLOCK_N;
      const int bz = z / p.heads, hz = z % p.heads;
      const int buf = lt & 1;
      const int row = mt * 2 * kBlockM + static_cast<int>(rank) * kBlockM + quarter * 32 + lane;
      const bool row_ok = row < p.m;
      const bool row_live = row_ok && (p.lens == nullptr || row < __ldg(p.lens + bz));
      const long long y_off = bz * p.y_batch_stride + hz * p.y_head_stride + static_cast<long long>(row) * p.y_ld;
      if (!gemm_tile_live(p, mt * 2 * kBlockM, bz)) {
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
          gemm_epilogue_chunk(p, v, tn0 + c * 32, row_ok, false, y_off);
        }
        continue;
      }
      ++lt;
      mbar_wait_a(acc_full + 8 * buf, ((lt - 1) >> 1) & 1);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        float v[32];
        __syncwarp();
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * Cfg::kTmemCols + c * 32, v);
        tmem_ld_wait();
        if (c == BLOCK_N / 32 - 1) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed_a(acc_empty_l + 8 * buf);
        }
        gemm_epilogue_chunk(p, v, tn0 + c * 32, row_ok, row_live, y_off);
      }
    }
    tcgen05_fence_before();
  }
  cluster_sync();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_2sm<2 * Cfg::kTmemCols>(tmem_base);
  }
}

// fp32 SIMT evaluation of the same contract (debug / cross-check).
__global__ void conv_gemm_simt_kernel(const __nv_bfloat16* a_hi, const __nv_bfloat16* a_lo, const __nv_bfloat16* b_hi,
                                      const __nv_bfloat16* b_lo, pk_operand oa, pk_operand ob, GemmKernelArgs p, int k,
                                      int batch) {
  const long long total = static_cast<long long>(batch) * p.heads * p.m * p.n;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int n = idx % p.n;
    const int t = (idx / p.n) % p.m;
    const int z = idx / (static_cast<long long>(p.n) * p.m);
    const int b = z / p.heads, h = z % p.heads;
    const long long ab = (b * p.a_bmul + h * p.a_hmul) * oa.batch_stride;
    const long long bb = (b * p.b_bmul + h * p.b_hmul) * ob.batch_stride;
    const int ac = p.a_col0 + h * p.a_colh, bc = p.b_col0 + h * p.b_colh;
    float acc = 0.f;
    for (int tap = 0; tap < p.taps; ++tap) {
      const int r = t + (tap - p.pad) * p.dil;
      if (r < 0 || r >= oa.rows) continue;
      for (int kk = 0; kk < k; ++kk) {
        if (ac + kk >= oa.cols) break;
        const long long ai = ab + static_cast<long long>(r) * oa.ld + ac + kk;
        const int bcol = bc + tap * p.b_tap_stride + kk;
        if (bcol >= ob.cols || n >= ob.rows) continue;
        const long long bi = bb + static_cast<long long>(n) * ob.ld + bcol;
        float av = __bfloat162float(a_hi[ai]);
        float bv = __bfloat162float(b_hi[bi]);
        if (p.passes == 3) {
          av += __bfloat162float(a_lo[ai]);
          bv += __bfloat162float(b_lo[bi]);
        }
        acc = fmaf(av, bv, acc);
      }
    }
    float x = acc * p.scale;
    if (p.bias) x += p.bias[n];
    if (p.act == PK_ACT_RELU) x = fmaxf(x, 0.f);
    else if (p.act == PK_ACT_TANH) x = tanhf(x);
    const long long yo = b * p.y_batch_stride + h * p.y_head_stride + static_cast<long long>(t) * p.y_ld + n;
    if (p.residual) x += p.residual[yo];
    if (p.lens && t >= p.lens[b]) x = 0.f;
    if (p.y_f32) p.y_f32[yo] = x;
    if (p.y_hi) {
      __nv_bfloat16 hh, ll;
      split_bf16(x, hh, ll);
      p.y_hi[yo] = hh;
      p.y_lo[yo] = ll;
    }
  }
}

static int validate_common(const pk_conv_gemm_args* a) {
  PK_CHECK_ARG(a != nullptr, "args is NULL");
  PK_CHECK_ARG(a->a.hi && a->b.hi, "operand hi planes must be non-NULL");
  PK_CHECK_ARG(a->passes == 1 || a->passes == 3, "passes must be 1 or 3 (got %d)", a->passes);
  PK_CHECK_ARG(a->passes == 1 || (a->a.lo && a->b.lo), "passes=3 needs lo planes");
  PK_CHECK_ARG(a->batch > 0 && a->heads > 0 && a->m > 0 && a->n > 0 && a->k > 0, "batch/heads/m/n/k must be > 0");
  PK_CHECK_ARG(a->taps >= 1 && a->dil >= 1 && a->pad >= 0, "bad taps/dil/pad");
  PK_CHECK_ARG((a->a.ld % 8) == 0 && (a->b.ld % 8) == 0, "operand row strides must be multiples of 8 elements (16 B)");
  PK_CHECK_ARG((a->a.batch_stride % 8) == 0 && (a->b.batch_stride % 8) == 0, "operand batch strides must be multiples of 8");
  PK_CHECK_ARG((reinterpret_cast<uintptr_t>(a->a.hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->b.hi) & 15) == 0,
               "operand planes must be 16-byte aligned");
  PK_CHECK_ARG((a->y_hi == nullptr) == (a->y_lo == nullptr), "y_hi and y_lo must both be set or both NULL");
  PK_CHECK_ARG(a->act >= PK_ACT_NONE && a->act <= PK_ACT_TANH, "unknown activation %d", a->act);
  return PK_OK;
}

static int validate(const pk_conv_gemm_args* a) {
  int rc = validate_common(a);
  if (rc) return rc;
  PK_CHECK_ARG((a->y_f32 != nullptr) || (a->y_hi != nullptr), "no output requested");
  return PK_OK;
}

static int validate_epilogue(const pk_conv_gemm_args* a, const pk_gemm_epilogue* e) {
  int rc = validate_common(a);
  if (rc) return rc;
  PK_CHECK_ARG(e->mode == PK_EPI_GATE || e->mode == PK_EPI_WF_UPDATE, "unknown epilogue mode %d", e->mode);
  PK_CHECK_ARG(e->channels > 0 && (e->channels % 32) == 0 && a->n == 2 * e->channels && a->n <= 256,
               "fused epilogues need n == 2 * channels, channels %% 32 == 0, n <= 256 (n=%d channels=%d)", a->n, e->channels);
  PK_CHECK_ARG(a->heads == 1 && a->act == PK_ACT_NONE && a->residual == nullptr && a->lens == nullptr,
               "fused epilogues take heads == 1, no activation / residual / lens in the base arguments");
  PK_CHECK_ARG(a->bias == nullptr || (reinterpret_cast<uintptr_t>(a->bias) & 15) == 0, "bias must be 16-byte aligned");
  if (e->mode == PK_EPI_GATE) {
    PK_CHECK_ARG(a->y_hi && a->y_lo && (a->y_ld % 8) == 0, "GATE writes split planes (y_hi / y_lo), y_ld %% 8 == 0");
    PK_CHECK_ARG(e->residual == nullptr || ((e->residual_ld % 4) == 0 && (e->residual_batch_stride % 4) == 0 &&
                                            (reinterpret_cast<uintptr_t>(e->residual) & 15) == 0),
                 "GATE residual must be 16-byte aligned with strides %% 4 == 0");
  } else {
    PK_CHECK_ARG(e->state && e->skip, "WF_UPDATE needs state and skip");
    PK_CHECK_ARG((e->buf_hi == nullptr) == (e->buf_lo == nullptr), "buf_hi and buf_lo must both be set or both NULL");
    PK_CHECK_ARG(e->buf_hi == nullptr || ((e->buf_ld % 8) == 0 && (e->buf_col0 % 8) == 0), "buf_ld / buf_col0 must be multiples of 8");
  }
  return PK_OK;
}

static GemmKernelArgs to_kernel_args(const pk_conv_gemm_args* a) {
  GemmKernelArgs p;
  p.m = a->m; p.n = a->n; p.k_chunks = (a->k + kChunkK - 1) / kChunkK; p.taps = a->taps; p.dil = a->dil; p.pad = a->pad;
  p.heads = a->heads; p.batch = a->batch; p.tiles_m = 0; p.total_tiles = 0;
  p.a_bmul = a->a.bmul; p.a_hmul = a->a.hmul; p.a_col0 = a->a.col0; p.a_colh = a->a.colh;
  p.b_bmul = a->b.bmul; p.b_hmul = a->b.hmul; p.b_col0 = a->b.col0; p.b_colh = a->b.colh;
  p.b_tap_stride = p.k_chunks * kChunkK;
  p.scale = a->scale; p.bias = a->bias; p.act = a->act; p.residual = a->residual; p.lens = a->lens;
  p.y_f32 = a->y_f32; p.y_hi = static_cast<__nv_bfloat16*>(a->y_hi); p.y_lo = static_cast<__nv_bfloat16*>(a->y_lo);
  p.y_batch_stride = a->y_batch_stride; p.y_head_stride = a->y_head_stride; p.y_ld = a->y_ld;
  p.passes = a->passes;
  p.epi = PK_EPI_NONE; p.epi_c = 0; p.e_res = nullptr; p.e_res_bs = 0; p.e_res_ld = 0; p.e_state = nullptr; p.e_skip = nullptr;
  p.e_skip_init = 0; p.e_buf_hi = nullptr; p.e_buf_lo = nullptr; p.e_buf_ld = 0; p.e_buf_col0 = 0;
  return p;
}

template <int BLOCK_N>
static int launch(const pk_conv_gemm_args* a, cudaStream_t stream, const pk_gemm_epilogue* e = nullptr) {
  using Cfg = GemmCfg<BLOCK_N>;
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  int rc;
  if ((rc = encode_tmap_bf16_3d(&ta_hi, a->a.hi, a->a.cols, a->a.rows, a->a.batches, a->a.ld, a->a.batch_stride, kBlockM))) return rc;
  if ((rc = encode_tmap_bf16_3d(&tb_hi, a->b.hi, a->b.cols, a->b.rows, a->b.batches, a->b.ld, a->b.batch_stride, BLOCK_N))) return rc;
  if (a->passes == 3) {
    if ((rc = encode_tmap_bf16_3d(&ta_lo, a->a.lo, a->a.cols, a->a.rows, a->a.batches, a->a.ld, a->a.batch_stride, kBlockM))) return rc;
    if ((rc = encode_tmap_bf16_3d(&tb_lo, a->b.lo, a->b.cols, a->b.rows, a->b.batches, a->b.ld, a->b.batch_stride, BLOCK_N))) return rc;
  } else {
    ta_lo = ta_hi;
    tb_lo = tb_hi;
  }
  static bool attr_set = false;
  if (!attr_set) {
    PK_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  GemmKernelArgs p = to_kernel_args(a);
  if (e != nullptr) {
    p.epi = e->mode; p.epi_c = e->channels; p.e_res = e->residual; p.e_res_bs = e->residual_batch_stride; p.e_res_ld = e->residual_ld;
    p.e_state = e->state; p.e_skip = e->skip; p.e_skip_init = e->skip_init;
    p.e_buf_hi = static_cast<__nv_bfloat16*>(e->buf_hi); p.e_buf_lo = static_cast<__nv_bfloat16*>(e->buf_lo);
    p.e_buf_ld = e->buf_ld; p.e_buf_col0 = e->buf_col0;
  }
  p.tiles_m = (a->m + kBlockM - 1) / kBlockM;
  const long long total = static_cast<long long>(p.tiles_m) * ((a->n + BLOCK_N - 1) / BLOCK_N) * a->batch * a->heads;
  PK_CHECK_ARG(total < (1LL << 31), "too many output tiles");
  p.total_tiles = static_cast<int>(total);
  const int grid = static_cast<int>(std::min<long long>(total, sm_count()));   // persistent: one CTA per SM
  conv_gemm_kernel<BLOCK_N><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta_hi, ta_lo, tb_hi, tb_lo, p);
  PK_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return PK_OK;
}

static bool use_pair() {
  static const bool v = []() {
    const char* e = getenv("PK_GEMM_PAIR");
    return !(e && e[0] == '0') && sm_count() >= 2;
  }();
  return v;
}

template <int BLOCK_N>
static int launch_pair(const pk_conv_gemm_args* a, cudaStream_t stream) {
  using Cfg = GemmPairCfg<BLOCK_N>;
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  int rc;
  if ((rc = encode_tmap_bf16_3d(&ta_hi, a->a.hi, a->a.cols, a->a.rows, a->a.batches, a->a.ld, a->a.batch_stride, kBlockM))) return rc;
  if ((rc = encode_tmap_bf16_3d(&tb_hi, a->b.hi, a->b.cols, a->b.rows, a->b.batches, a->b.ld, a->b.batch_stride, BLOCK_N / 2))) return rc;
  if (a->passes == 3) {
    if ((rc = encode_tmap_bf16_3d(&ta_lo, a->a.lo, a->a.cols, a->a.rows, a->a.batches, a->a.ld, a->a.batch_stride, kBlockM))) return rc;
    if ((rc = encode_tmap_bf16_3d(&tb_lo, a->b.lo, a->b.cols, a->b.rows, a->b.batches, a->b.ld, a->b.batch_stride, BLOCK_N / 2))) return rc;
  } else {
    ta_lo = ta_hi;
    tb_lo = tb_hi;
  }
  static bool attr_set = false;
  if (!attr_set) {
    PK_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_pair_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  GemmKernelArgs p = to_kernel_args(a);
  p.tiles_m = (a->m + 2 * kBlockM - 1) / (2 * kBlockM);        // pair tiles of 256 rows
  const long long total = static_cast<long long>(p.tiles_m) * ((a->n + BLOCK_N - 1) / BLOCK_N) * a->batch * a->heads;
  PK_CHECK_ARG(total < (1LL << 31), "too many output tiles");
  p.total_tiles = static_cast<int>(total);
  const int grid = 2 * static_cast<int>(std::min<long long>(total, sm_count() / 2));
  conv_gemm_pair_kernel<BLOCK_N><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta_hi, ta_lo, tb_hi, tb_lo, p);
  PK_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return PK_OK;
}

}  // namespace pk

extern "C" int pk_conv_gemm(const pk_conv_gemm_args* args, pk_stream_t stream) {
  int rc = pk::validate(args);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // tile width with the fewest padded columns (ties -> wider tile)
  const int n = args->n;
  auto waste = [n](int bn) { return (n + bn - 1) / bn * bn - n; };
  if (n <= 64) return pk::launch<64>(args, s);
  if (waste(256) <= waste(128) && n > 128) {
    // wide outputs: CTA pairs (half of the B tile per CTA) unless PK_GEMM_PAIR=0 or the rows fit one 128-row tile
    if (pk::use_pair() && args->m > 128) return pk::launch_pair<256>(args, s);
    return pk::launch<256>(args, s);
  }
  // N = 384 / 1152 / ... : 192-wide pair tiles have no padded columns and a pair instruction of 96 clk of work
  if (pk::use_pair() && args->m > 128 && n >= 384 && waste(192) < waste(256) && waste(192) <= waste(128))
    return pk::launch_pair<192>(args, s);
  return pk::launch<128>(args, s);
}

extern "C" int pk_conv_gemm_ex(const pk_conv_gemm_args* args, const pk_gemm_epilogue* epi, pk_stream_t stream) {
  if (epi == nullptr || epi->mode == PK_EPI_NONE) return pk_conv_gemm(args, stream);
  int rc = pk::validate_epilogue(args, epi);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // the whole 2C-wide row must sit in one tile of the single-CTA kernel
  return args->n <= 128 ? pk::launch<128>(args, s, epi) : pk::launch<256>(args, s, epi);
}

extern "C" int pk_conv_gemm_simt(const pk_conv_gemm_args* args, pk_stream_t stream) {
  int rc = pk::validate(args);
  if (rc) return rc;
  const pk::GemmKernelArgs p = pk::to_kernel_args(args);
  const long long total = static_cast<long long>(args->batch) * args->heads * args->m * args->n;
  const int threads = 256;
  const int blocks = static_cast<int>(std::min<long long>((total + threads - 1) / threads, 148LL * 16));
  pk::conv_gemm_simt_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(args->a.hi), static_cast<const __nv_bfloat16*>(args->a.lo),
      static_cast<const __nv_bfloat16*>(args->b.hi), static_cast<const __nv_bfloat16*>(args->b.lo), args->a, args->b, p,
      args->k, args->batch);
  PK_CHECK_CUDA(cudaGetLastError());
  pk::count_launch();
  return PK_OK;
}
