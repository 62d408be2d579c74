// pk_waveflow_layer: one ResidualBlock.add_input of the WaveFlow inverse (reference parakeet/models/waveflow.py:248-285) as
// ONE CTA-pair kernel - the row-by-row autoregressive path spends its time here (8 flows x 15 rows x 8 layers launches).
//
//   a | g   = conv2d(ring of the last 3 rows, dilation 2^l along the width) + condition_proj(condition row) + biases
//   z       = tanh(a) * sigmoid(g)
//   res|skip= out_proj(z);   new row = row + res  -> next layer's ring slot (split planes);   skip (=|+=) skip
//
// It replaces, per layer, two pk_conv_gemm_ex launches (gate epilogue, wf_update epilogue), one fp32 read of the hoisted
// condition projections (2C floats per position) and the fp32 state read-modify-write.  Structure = pwg_fc.cu:
//   * M = 256 positions per pair tile (128 per CTA), cta_group::2 MMAs, N = 128 = a | g channels;
//   * GEMM1 has 11 K-chunks: 3 width taps x 3 ring slots of 64 channels + the 80 condition channels (64 + 16).  K = 656 per
//     output channel does not fit shared memory as a resident operand (168 KB per CTA with both planes), so every stage of
//     the 4-deep ring carries the A chunk (32 KB: hi | lo) AND this CTA's 64 output channels of the weight chunk
//     (16 KB: hi | lo, L2 resident: 360 KB per layer variant);
//   * z goes back into tensor memory over the accumulator columns it was computed from and is GEMM2's A operand;
//   * the residual add is a tensor-core pass of the centre-tap / newest-slot chunk with [0 | I] (no fp32 state tensor: the
//     running row lives in the ring slot as hi + lo, 16 mantissa bits, re-split after every layer);
//   * GEMM2's B operand is out_proj with its rows reordered to [skip | res] so that the accumulator halves line up with the
//     two groups of store warps.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "pk_host.h"
#include "pk_sm100.cuh"

namespace pk {
namespace wf {

constexpr int kC = 64;                                       // residual channels
constexpr int kG = 128;                                      // gate channels (a | g) == out_proj outputs (skip | res)
constexpr int kATile = 128 * kSwizzleBytes;                  // 16 KB: 128 positions x one 64-channel chunk of one plane
constexpr int kWTile = 64 * kSwizzleBytes;                   // 8 KB: 64 output channels x one chunk of one plane
constexpr int kStages = 4;
constexpr int kStageBytes = 2 * kATile + 2 * kWTile;         // A hi | A lo | W hi | W lo
constexpr int kChunks = 11;                                  // 9 conv chunks + 2 condition chunks
constexpr int kW1Cols = kChunks * kChunkK;                   // 704: row length of the packed GEMM1 weight
constexpr int kGateWarps = 4;
constexpr int kStoreWarps = 8;
constexpr int kFirstGateWarp = 4;
constexpr int kThreads = (kFirstGateWarp + kGateWarps + kStoreWarps) * 32;
constexpr int kSmem = kStages * kStageBytes + 2 * kWTile + kWTile + 1024 + 256;
static_assert(kSmem <= 227 * 1024, "shared memory budget");

struct LayerArgs {
  int batch, w, dil;
  int resid_chunk;              // 3 + slot: the centre-tap chunk of the newest row's ring slot
  int cond_ksteps_last;         // K-steps of the second condition chunk ((n_mels - 64 + 15) / 16)
  float gate_c[128];            // pre-scaled biases of the gate (see the host code)
  float out_b[128];             // out_proj bias in accumulator order: skip | res
  float k_a, k_g;
  float* skip;
  int skip_init;
  __nv_bfloat16* y_hi;          // next layer's ring planes (batch, w, y_ld), written at column y_col0; NULL on the last layer
  __nv_bfloat16* y_lo;
  int y_ld, y_col0;
  unsigned long long* prof;
};

#define PK_TICK(k)                                      \
  if (kProf) {                                          \
    const long long n_ = clock64();                     \
    tacc[k] += n_ - tlast;                              \
    tlast = n_;                                         \
  }

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float ra = a - __uint_as_float(hi << 16);
  const float rb = b - __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(ra, rb);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void st_global_v8(void* ptr, const uint32_t* w) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
               "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}

struct TileIter {   // 256-position tiles of the pair, per batch element; this CTA owns positions [m0 + 128 * rank, +128)
  int idx, step, tiles_per_b, total;
  __device__ TileIter(const LayerArgs& p)
      : idx(static_cast<int>(blockIdx.x >> 1) - static_cast<int>(gridDim.x >> 1)), step(gridDim.x >> 1),
        tiles_per_b((p.w + 255) >> 8), total(((p.w + 255) >> 8) * p.batch) {}
  __device__ bool next(int& b, int& m0) {
    idx += step;
    if (idx >= total) return false;
    b = idx / tiles_per_b;
    m0 = (idx % tiles_per_b) * 256;
    return true;
  }
};

template <bool kProf>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
waveflow_layer_kernel(const __grid_constant__ CUtensorMap tm_x,    // ring planes (batch, w, 3C): 4-D, both planes in one box
                      const __grid_constant__ CUtensorMap tm_c,    // condition row planes (batch, w, n_mels)
                      const __grid_constant__ CUtensorMap tm_w1,   // packed GEMM1 weight planes (128, 704), box = 64 rows
                      const __grid_constant__ CUtensorMap tm_w2_hi, const __grid_constant__ CUtensorMap tm_w2_lo,
                      const LayerArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w2 = smem + kStages * kStageBytes;            // [hi | lo] this CTA's 64 rows of out_proj
  const uint32_t ident = w2 + 2 * kWTile;                      // this CTA's 64 rows of [0 | I]
  const uint32_t bars = ident + kWTile;
  const uint32_t full_bar = bars;                              // [stages]   (leader's copy is the live one)
  const uint32_t empty_bar = full_bar + 8 * kStages;           // [stages]
  const uint32_t acc1_full = empty_bar + 8 * kStages;          // [2]
  const uint32_t acc2_full = acc1_full + 16;                   // [2]
  const uint32_t acc2_empty = acc2_full + 16;                  // [2] leader
  const uint32_t z_full = acc2_empty + 16;                     // [2] leader
  const uint32_t w_bar = z_full + 16;
  const uint32_t tmem_slot = w_bar + 8;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x); tma_prefetch_desc(&tm_c); tma_prefetch_desc(&tm_w1);
    tma_prefetch_desc(&tm_w2_hi); tma_prefetch_desc(&tm_w2_lo);
    for (int s = 0; s < kStages; ++s) { mbar_init_a(full_bar + 8 * s, 1); mbar_init_a(empty_bar + 8 * s, 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init_a(acc1_full + 8 * i, 1);
      mbar_init_a(acc2_full + 8 * i, 1); mbar_init_a(acc2_empty + 8 * i, 2 * kStoreWarps);
      mbar_init_a(z_full + 8 * i, 2 * kGateWarps);
    }
    mbar_init_a(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm_a<512>(tmem_slot);
  if (threadIdx.x >= 128 && threadIdx.x < 192) {
    // this CTA's half of [0 | I]: rank 0 supplies accumulator columns 0..63 (skip half: nothing added), rank 1 columns
    // 64..127 (row n = e_n: column 64 + n receives channel n of the newest row)
    const int n = threadIdx.x - 128;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      uint4 v = make_uint4(0, 0, 0, 0);
      if (rank == 1 && (n >> 3) == c) {
        const uint32_t one = (n & 1) ? 0x3f800000u : 0x00003f80u;
        const int wd = (n & 7) >> 1;
        v.x = wd == 0 ? one : 0; v.y = wd == 1 ? one : 0; v.z = wd == 2 ? one : 0; v.w = wd == 3 ? one : 0;
      }
      sts_u4(ident + n * kSwizzleBytes + ((c ^ (n & 7)) * 16), v);
    }
    fence_proxy_async_all();
  }
  tcgen05_fence_before();
  cluster_sync();                      // barriers of both CTAs are initialised before any remote arrive / TMA credit
  tcgen05_fence_after();
  if (warp == 0 && lane == 0) {
    mbar_arrive_expect_tx_a(w_bar, 2 * kWTile);
    tma_load_3d_a(w2, &tm_w2_hi, w_bar, 0, 64 * rank, 0);
    tma_load_3d_a(w2 + kWTile, &tm_w2_lo, w_bar, 0, 64 * rank, 0);
    mbar_wait_a(w_bar, 0);
  }
  cluster_sync();                      // both halves of out_proj are in place before the leader's first GEMM2
  const uint32_t tmem_base = lds_u32(tmem_slot);

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------ TMA producer (both CTAs: own positions, own weight rows) ------------------------------
      uint32_t it = 0;
      const uint32_t full_leader = mapa_shared(full_bar, 0);
      TileIter ti(p);
      int b, m0;
      while (ti.next(b, m0)) {
        const int row0 = m0 + 128 * static_cast<int>(rank);
        for (int j = 0; j < kChunks; ++j, ++it) {
          const int s = it % kStages;
          mbar_wait_a(empty_bar + 8 * s, ((it / kStages) & 1) ^ 1);
          const uint32_t st = smem + s * kStageBytes;
          const uint32_t fb = full_leader + 8 * s;
          if (leader) mbar_arrive_expect_tx_a(full_bar + 8 * s, 2 * kStageBytes);       // the chunks of both CTAs
          if (j < 9) {
            const int tap = j / 3, slot = j - 3 * tap;
            tma_load_4d_2sm_a(st, &tm_x, fb, slot * kC, row0 + (tap - 1) * p.dil, b, 0);   // rows outside [0, w) read as zero
          } else {
            tma_load_4d_2sm_a(st, &tm_c, fb, (j - 9) * kChunkK, row0, b, 0);               // columns >= n_mels read as zero
          }
          tma_load_4d_2sm_a(st + 2 * kATile, &tm_w1, fb, j * kChunkK, 64 * static_cast<int>(rank), 0, 0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ------------------------------ MMA issuer (leader CTA only) ------------------------------
      constexpr uint32_t idesc = make_idesc_bf16_f32(256, 128);
      uint32_t it = 0;
      long long tacc[4] = {0, 0, 0, 0};
      long long tlast = clock64();
      auto g1 = [&](int i) {
        // acc1(i & 1): its previous user is tile i-2, whose GEMM2 (the last reader - z lives in the accumulator's own columns)
        // was issued by this thread before this point; tcgen05.mma execute in issue order
        const int buf = i & 1;
        const uint32_t d = tmem_base + buf * 128;
        for (int j = 0; j < kChunks; ++j, ++it) {
          const int s = it % kStages;
          PK_TICK(0)
          mbar_wait_a(full_bar + 8 * s, (it / kStages) & 1);
          PK_TICK(1)
          tcgen05_fence_after();
          const uint32_t st = smem + s * kStageBytes;
          const uint64_t a_hi = make_smem_desc_sw128(st), a_lo = make_smem_desc_sw128(st + kATile);
          const uint64_t b_hi = make_smem_desc_sw128(st + 2 * kATile), b_lo = make_smem_desc_sw128(st + 2 * kATile + kWTile);
          const int ksteps = j == kChunks - 1 ? p.cond_ksteps_last : 4;
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t koff = static_cast<uint64_t>((k * kUmmaK * 2) >> 4);
            umma_bf16_2sm(d, a_hi + koff, b_hi + koff, idesc, !(j == 0 && k == 0));
            umma_bf16_2sm(d, a_lo + koff, b_hi + koff, idesc, 1);
            umma_bf16_2sm(d, a_hi + koff, b_lo + koff, idesc, 1);
          }
          if (j == p.resid_chunk) {
            // residual pass: acc2(i) = [0 | row_hi + row_lo] from the newest row's centre-tap tiles of both CTAs
            mbar_wait_a(acc2_empty + 8 * buf, ((i >> 1) & 1) ^ 1);
            PK_TICK(2)
            tcgen05_fence_after();
            const uint64_t b_id = make_smem_desc_sw128(ident);
            const uint32_t d2 = tmem_base + 256 + buf * 128;
            for (int k = 0; k < 4; ++k) {
              const uint64_t koff = static_cast<uint64_t>((k * kUmmaK * 2) >> 4);
              umma_bf16_2sm(d2, a_hi + koff, b_id + koff, idesc, k != 0);
              umma_bf16_2sm(d2, a_lo + koff, b_id + koff, idesc, 1);
            }
          }
          umma_commit_2sm_a(empty_bar + 8 * s);
        }
        umma_commit_2sm_a(acc1_full + 8 * buf);
      };
      auto g2 = [&](int i) {
        const int buf = i & 1;
        PK_TICK(0)
        mbar_wait_a(z_full + 8 * buf, (i >> 1) & 1);   // the gate warps of both CTAs wrote z over acc1(buf)
        PK_TICK(3)
        tcgen05_fence_after();
        // A from tensor memory: z_hi / z_lo of channels [32 h, 32 h + 32) sit in columns 32 h + [0, 16) / 32 h + [16, 32) of
        // acc1(buf), one 32-bit column per channel pair: K-step k (channels 16 k ..) starts at column 32 (k / 2) + 8 (k % 2)
        const uint32_t za = tmem_base + buf * 128;
        const uint32_t d2 = tmem_base + 256 + buf * 128;
        const uint64_t b_hi = make_smem_desc_sw128(w2), b_lo = make_smem_desc_sw128(w2 + kWTile);
        for (int k = 0; k < 4; ++k) {
          const uint64_t koff = static_cast<uint64_t>((k * kUmmaK * 2) >> 4);
          const uint32_t a_hi = za + 32 * (k >> 1) + 8 * (k & 1), a_lo = a_hi + 16;
          umma_bf16_2sm_ts(d2, a_hi, b_hi + koff, idesc, 1);   // on top of the residual pass
          umma_bf16_2sm_ts(d2, a_lo, b_hi + koff, idesc, 1);
          umma_bf16_2sm_ts(d2, a_hi, b_lo + koff, idesc, 1);
        }
        umma_commit_2sm_a(acc2_full + 8 * buf);
      };
      TileIter ti(p);
      int b, m0;
      int n_issued = 0, n_done = 0;
      bool have = ti.next(b, m0);
      if (have) g1(n_issued++);
      while (have) {
        const bool have_next = ti.next(b, m0);
        if (have_next) g1(n_issued++);
        g2(n_done++);
        have = have_next;
      }
      PK_TICK(0)
      if (kProf) {
        for (int k = 0; k < 4; ++k) atomicAdd(p.prof + k, static_cast<unsigned long long>(tacc[k]));
        atomicAdd(p.prof + 4, static_cast<unsigned long long>(n_done));
      }
    }
  } else if (warp < kFirstGateWarp) {
    // idle warps
  } else if (warp < kFirstGateWarp + kGateWarps) {
    // ------------------------------ gate warps (both CTAs, own TMEM lanes) ------------------------------
    const int quarter = warp & 3;
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t z_full_l = mapa_shared(z_full, 0);
    float k_a, k_g;
    asm volatile("mov.f32 %0, %2;\n\tmov.f32 %1, %3;" : "=f"(k_a), "=f"(k_g) : "f"(p.k_a), "f"(p.k_g));
    TileIter ti(p);
    int b, m0;
    for (int i = 0; ti.next(b, m0); ++i) {
      const int buf = i & 1;
      mbar_wait_a(acc1_full + 8 * buf, (i >> 1) & 1);
      tcgen05_fence_after();
      const uint32_t acc = tmem_base + lane_base + buf * 128;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float va[32], vb[32];
        uint32_t zw[32];                          // [0, 16): z_hi of channels 32 half + (0 .. 31), [16, 32): z_lo
        __syncwarp();
        tmem_ld_32x32(acc + half * 32, va);
        tmem_ld_32x32(acc + 64 + half * 32, vb);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float z[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            // tanh(a) sigmoid(g) = (1 - e1) / ((1 + e1) (1 + e2)), e1 = exp(-2a) (clamped: e1 * e2 must stay finite), e2 = exp(-g)
            const float e1 = ex2_approx(fminf(fmaf(va[j + e], k_a, p.gate_c[half * 32 + j + e]), 60.f));
            const float e2 = ex2_approx(fminf(fmaf(vb[j + e], k_g, p.gate_c[64 + half * 32 + j + e]), 60.f));
            const float t1 = 1.f + e1;
            z[e] = (1.f - e1) * rcp_approx(fmaf(t1, e2, t1));
          }
          split2(z[0], z[1], zw[j / 2], zw[16 + j / 2]);
          split2(z[2], z[3], zw[j / 2 + 1], zw[16 + j / 2 + 1]);
        }
        // over the a-columns this half has just been read from (the g-columns [64, 128) stay untouched until GEMM1 of tile i+2)
        tmem_st_32x32(acc + half * 32, zw);
      }
      tmem_st_wait();                    // z is in tensor memory
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed_a(z_full_l + 8 * buf);
    }
  } else {
    // ------------------------------ store warps (both CTAs) ------------------------------
    const int sw = warp - kFirstGateWarp - kGateWarps;
    const int quarter = warp & 3;
    const int half = sw >> 2;                 // 0: skip columns [0, 64), 1: new row columns [64, 128)
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t acc2_empty_l = mapa_shared(acc2_empty, 0);
    TileIter ti(p);
    int b, m0;
    for (int i = 0; ti.next(b, m0); ++i) {
      const int buf = i & 1;
      const int row = m0 + 128 * static_cast<int>(rank) + quarter * 32 + lane;
      const long long pos = static_cast<long long>(b) * p.w + row;
      mbar_wait_a(acc2_full + 8 * buf, (i >> 1) & 1);
      tcgen05_fence_after();
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        float v[32];
        __syncwarp();
        tmem_ld_32x32(tmem_base + lane_base + 256 + buf * 128 + half * 64 + pass * 32, v);
        tmem_ld_wait();
        if (pass == 1) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed_a(acc2_empty_l + 8 * buf);
        }
        const float* ob = p.out_b + half * 64 + pass * 32;
        if (row >= p.w) {
          // positions past the end of the row: nothing to store
        } else if (half == 0) {
          float* dst = p.skip + pos * kC + pass * 32;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 o = make_float4(v[4 * c] + ob[4 * c], v[4 * c + 1] + ob[4 * c + 1], v[4 * c + 2] + ob[4 * c + 2],
                                         v[4 * c + 3] + ob[4 * c + 3]);
            if (p.skip_init) {
              *reinterpret_cast<float4*>(dst + 4 * c) = o;
            } else {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * c), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w)
                           : "memory");
            }
          }
        } else if (p.y_hi != nullptr) {
          uint32_t oh[16], ol[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) split2(v[2 * e] + ob[2 * e], v[2 * e + 1] + ob[2 * e + 1], oh[e], ol[e]);
          const long long off = pos * p.y_ld + p.y_col0 + pass * 32;
          st_global_v8(p.y_hi + off, oh);
          st_global_v8(p.y_hi + off + 16, oh + 8);
          st_global_v8(p.y_lo + off, ol);
          st_global_v8(p.y_lo + off + 16, ol + 8);
        }
      }
    }
  }
  tcgen05_fence_before();
  cluster_sync();                      // neither CTA may free its TMEM / exit while the pair's MMAs can still touch it
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

}  // namespace wf
}  // namespace pk

extern "C" int pk_waveflow_layer(const pk_waveflow_layer_args* a, pk_stream_t stream) {
  using namespace pk;
  using namespace pk::wf;
  PK_CHECK_ARG(a != nullptr, "args is NULL");
  PK_CHECK_ARG(a->batch > 0 && a->width > 0 && a->dilation >= 1, "bad batch/width/dilation");
  PK_CHECK_ARG(a->channels == kC, "the fused WaveFlow layer is built for 64 residual channels (got %d)", a->channels);
  PK_CHECK_ARG(a->n_mels > 64 && a->n_mels <= 128 && (a->n_mels % 8) == 0, "n_mels must be in (64, 128], a multiple of 8");
  PK_CHECK_ARG(a->slot >= 0 && a->slot < 3, "slot must be 0..2");
  PK_CHECK_ARG(a->buf_hi && a->buf_lo && a->cond_hi && a->cond_lo && a->w1_hi && a->w1_lo && a->w2_hi && a->w2_lo && a->bias1 &&
               a->bias2 && a->skip, "NULL pointer in pk_waveflow_layer_args");
  PK_CHECK_ARG((a->next_hi == nullptr) == (a->next_lo == nullptr), "next_hi / next_lo: both or neither");
  PK_CHECK_ARG(a->next_hi != a->buf_hi, "the next layer's ring must not alias this layer's");
  PK_CHECK_ARG(a->cond_batch_stride >= static_cast<int64_t>(a->width) * a->n_mels && (a->cond_batch_stride % 8) == 0,
               "bad condition batch stride");
  PK_CHECK_ARG(sm_count() >= 2, "needs at least one SM pair");
  CUtensorMap tx, tc, tw1, tw2_hi, tw2_lo;
  int rc;
  const uint64_t W = a->width, B = a->batch;
  if ((rc = encode_tmap_bf16_planes(&tx, a->buf_hi, a->buf_lo, 3 * kC, W, B, 3 * kC, W * 3 * kC, 128))) return rc;
  if ((rc = encode_tmap_bf16_planes(&tc, a->cond_hi, a->cond_lo, a->n_mels, W, B, a->n_mels, a->cond_batch_stride, 128))) return rc;
  if ((rc = encode_tmap_bf16_planes(&tw1, a->w1_hi, a->w1_lo, kW1Cols, kG, 1, kW1Cols, 0, 64))) return rc;
  if ((rc = encode_tmap_bf16_3d(&tw2_hi, a->w2_hi, 64, 128, 1, 64, 0, 64))) return rc;
  if ((rc = encode_tmap_bf16_3d(&tw2_lo, a->w2_lo, 64, 128, 1, 64, 0, 64))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    PK_CHECK_CUDA(cudaFuncSetAttribute(waveflow_layer_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    PK_CHECK_CUDA(cudaFuncSetAttribute(waveflow_layer_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr_set = true;
  }
  LayerArgs p;
  p.batch = a->batch; p.w = a->width; p.dil = a->dilation;
  p.resid_chunk = 3 + a->slot;
  p.cond_ksteps_last = (a->n_mels - 64 + kUmmaK - 1) / kUmmaK;
  constexpr float kLog2e = 1.4426950408889634f;
  p.k_a = -2.f * kLog2e; p.k_g = -kLog2e;
  for (int i = 0; i < 64; ++i) {
    p.gate_c[i] = -2.f * kLog2e * a->bias1[i];
    p.gate_c[64 + i] = -kLog2e * a->bias1[64 + i];
  }
  for (int i = 0; i < 128; ++i) p.out_b[i] = a->bias2[i];
  p.skip = a->skip; p.skip_init = a->skip_init;
  p.y_hi = static_cast<__nv_bfloat16*>(a->next_hi); p.y_lo = static_cast<__nv_bfloat16*>(a->next_lo);
  p.y_ld = 3 * kC; p.y_col0 = a->slot * kC;
  p.prof = static_cast<unsigned long long*>(a->prof);
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int pair_tiles = ((a->width + 255) / 256) * a->batch;
  const int grid = 2 * std::min(pair_tiles, sm_count() / 2);
  if (p.prof != nullptr) {
    waveflow_layer_kernel<true><<<grid, kThreads, kSmem, st>>>(tx, tc, tw1, tw2_hi, tw2_lo, p);
  } else {
    waveflow_layer_kernel<false><<<grid, kThreads, kSmem, st>>>(tx, tc, tw1, tw2_hi, tw2_lo, p);
  }
  PK_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return PK_OK;
}
