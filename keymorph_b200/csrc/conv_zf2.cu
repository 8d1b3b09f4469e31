// z-folded AND 2-CTA tcgen05 implicit-GEMM 3x3x3 convolution for the Cout = 64 layers
// (keymorph/unet3d/buildingblocks.py:50-52; in the bench network 32->64 and 64->64 at 128^3, 64->64 at
// 64^3 and the 192->64 first decoder conv at 128^3, together 60 % of the backbone's FLOPs).
//
// It combines the two ideas that are measured separately in conv_zf.cu and conv_tc2.cu:
//   * dz folded into N (conv_zf.cu): one input plane z' feeds, with 9 (dx, dy) taps instead of 27,
//     the three output planes z'+1, z', z'-1, which live as a rotating ring of three 64-column blocks
//     in TMEM while the CTA walks along z.  N = 192 per MMA instead of 64: a third of the MMA
//     instructions and a third of the activation bytes read from shared memory per output.
//   * CTA pairs (conv_tc2.cu): N = 192 weight rows would cost 6 KB of shared-memory reads per MMA and
//     24 KB of TMA per tap and Cin chunk; with cta_group::2 each SM holds and reads only 96 of them
//     (A 4 KB + B 3 KB per MMA), and the weights are streamed (3 ring rotations x 27 taps do not fit).
// Work unit = a column of ONE 8(x) x 16(y) brick x a z segment of up to 64 planes (+2 halo planes);
// the two CTAs of a pair take x-adjacent bricks and step through the planes in lockstep (they share
// every weight tile, hence the ring rotation).  A pair interleaves two unit pairs (TMEM sets 0 / 1).
// Barrier protocol: exactly conv_tc2.cu (leader-side full / tempty, multicast empty / tfull).
#include <algorithm>
#include <type_traits>
#include "km_common.cuh"
#include "tc_ptx.cuh"

using namespace kmtc;

namespace {

constexpr int kThreads = 320;           // warp 0 TMA, warp 1 MMA (leader), warps 2..9 epilogue
constexpr int kEpiThreads = 256;
constexpr int kStagesMax = 12;

// compile-time shape of one instantiation: COUT output channels, KC input channels per chunk
template <int COUT, int KC, int MT>
struct Cfg {
  static constexpr int kMT = MT;                       // y-adjacent bricks per unit sharing every weight tile
  static constexpr int kBoxRows = 8 * (16 * MT + 2);   // 8 x-voxels by 16 MT + 2 y-rows
  static constexpr int kSets = (2 * MT * 3 * COUT <= 512) ? 2 : 1;   // TMEM accumulator sets (units in flight)
  static constexpr uint32_t kSetStride = kSets == 2 ? 256u : 0u;
  static constexpr int kCout = COUT;
  static constexpr int kKC = KC;
  static constexpr int kN3 = 3 * COUT;                 // MMA N
  static constexpr int kHalfRows = kN3 / 2;            // weight rows held by each CTA
  static constexpr int kRowBytes = KC * 2;             // 128 B (SWIZZLE_128B) or 64 B (SWIZZLE_64B)
  static constexpr int kSteps = KC / 16;
  static constexpr uint32_t kASub = kBoxRows * kRowBytes;                                  // multiple of 1 KB
  static constexpr uint32_t kBTile = kHalfRows * kRowBytes;                                // one dy slice
  static constexpr uint32_t kStage = kASub + 3 * kBTile;                                   // one (dx, chunk)
  static constexpr int kCols = COUT / 2;               // output channels per epilogue thread
};

struct Zf2Geom {
  int N, D, H, W, chunks;
  int mt, sets;
  int xpairs, tiles_y, zsegs, lz, punits;   // punits = N * zsegs * tiles_y * xpairs (unit pairs), lz planes each
  int flags;
  uint32_t off_scratch, off_stats, off_bars;
  int stages;
  int w_per_sample;   // GN-folded path: one packed weight set per sample (tmB's last dimension is 3 N)
  int chunks0;        // K chunks [0, chunks0) come from tmA, the rest from tmA1 (un-materialised concat)
};

constexpr int kBiasClasses = 36;   // z code (0..3) x y code (0..2) x x code (0..2), see conv_zf.cu

struct Unit {
  int n, x0, y0, zs, planes;
  bool valid;   // this CTA's brick lies inside the volume (odd tiles_x: the last pair is half empty)
};

__device__ __forceinline__ Unit decode_unit(const Zf2Geom& g, int pu, uint32_t rank) {
  Unit r;
  const int xp = pu % g.xpairs;
  pu /= g.xpairs;
  r.x0 = (2 * xp + (int)rank) * 8;
  r.y0 = (pu % g.tiles_y) * (16 * g.mt);
  pu /= g.tiles_y;
  r.zs = (pu % g.zsegs) * g.lz;
  r.n = pu / g.zsegs;
  r.planes = min(g.lz, g.D - r.zs) + 2;
  r.valid = r.x0 < g.W;
  return r;
}

// tile sequence of a CTA pair (identical for all roles of both CTAs): unit pairs pair + k * npairs,
// two at a time, interleaved plane by plane; images never mix inside one interleave group
template <typename F>
__device__ __forceinline__ void for_each_tile(const Zf2Geom& g, uint32_t rank, F&& fn) {
  uint32_t cnt[2] = {0u, 0u};
  const int pair = (int)blockIdx.x >> 1, npairs = (int)gridDim.x >> 1;
  const int upi = g.punits / g.N;
  for (int n = 0; n < g.N; ++n) {
    if (g.sets == 1) {   // one TMEM set: the unit pairs of this CTA pair one after the other
      for (int ua = pair; ua < upi; ua += npairs) {
        const Unit a = decode_unit(g, n * upi + ua, rank);
        for (int p = 0; p < a.planes; ++p) fn(std::integral_constant<uint32_t, 0u>{}, a, p, cnt[0]++);
      }
      continue;
    }
    for (int ua = pair; ua < upi; ua += 2 * npairs) {
      const int ub = ua + npairs;
      const Unit a = decode_unit(g, n * upi + ua, rank);
      Unit b = a;
      b.planes = 0;
      if (ub < upi) b = decode_unit(g, n * upi + ub, rank);
      const int pmax = max(a.planes, b.planes);
      for (int p = 0; p < pmax; ++p) {
        if (p < a.planes) fn(std::integral_constant<uint32_t, 0u>{}, a, p, cnt[0]++);
        if (p < b.planes) fn(std::integral_constant<uint32_t, 1u>{}, b, p, cnt[1]++);
      }
    }
  }
}


template <int COUT, int KC, int MT, bool POOL, bool F16, bool ADD = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_zf2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA1,
                const __grid_constant__ CUtensorMap tmB, const Zf2Geom g, uint16_t* __restrict__ out, uint16_t* __restrict__ pooled,
                float* __restrict__ stats, const float* __restrict__ bias_tab, const uint16_t* __restrict__ addend) {
  using C = Cfg<COUT, KC, MT>;
  constexpr int kMT = C::kMT;
  constexpr uint32_t kSetStride = C::kSetStride;
  constexpr int kCout = C::kCout, kKC = C::kKC, kN3 = C::kN3, kHalfRows = C::kHalfRows;
  constexpr int kRowBytes = C::kRowBytes, kSteps = C::kSteps, kCols = C::kCols;
  constexpr uint32_t kASub = C::kASub, kBTile = C::kBTile, kStage = C::kStage;
  constexpr uint32_t kLayout = kRowBytes == 128 ? 2u : (kRowBytes == 64 ? 4u : 6u);   // SWIZZLE_128B / 64B / 32B
  constexpr uint32_t kSbo = 8u * kRowBytes;                  // bytes between 8-row groups
  const int kStages = g.stages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_u32);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  const uint32_t bars = base + g.off_bars;   // full[S], empty[S], tfull[2], tempty[2]
  auto full_bar = [&](int s) { return bars + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bars + 8u * (uint32_t)(kStages + s); };
  auto tfull_bar = [&](uint32_t a) { return bars + 8u * (uint32_t)(2 * kStages + a); };
  auto tempty_bar = [&](uint32_t a) { return bars + 8u * (uint32_t)(2 * kStages + 2 + a); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sm + g.off_bars + 8u * (2 * kStages + 5));

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 2);            // both CTAs' producers
      mbar_init(empty_bar(s), 1);
    }
    for (uint32_t a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 2 * kEpiThreads);   // both CTAs' epilogues
    }
    fence_mbar_init();
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmA1);
    prefetch_tmap(&tmB);
  }
  __syncwarp();
  if (warp == 0) {
    tmem_alloc2(smem_u32(tmem_ptr_smem), 512);
    tmem_relinquish2();
  }
  {
    float* s_stats = reinterpret_cast<float*>(sm + g.off_stats);
    if (g.flags & KM_CONV_STATS)
      for (int i = threadIdx.x; i < g.N * kCout * 2; i += kThreads) s_stats[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int subs = 3 * g.chunks;   // (dx, chunk) sub-iterations per plane, one pipeline stage each

  if (warp == 0) {
    // =============================== TMA producer (both CTAs) ====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for_each_tile(g, rank, [&](auto, const Unit& u, int p, uint32_t) {
        const int z = u.zs - 1 + p;   // input plane; outside [0, D) -> TMA zero fill
        const int rot = p % 3 + (g.w_per_sample ? 3 * u.n : 0);
        for (int dx = 0; dx < 3; ++dx) {
          for (int ch = 0; ch < g.chunks; ++ch) {
            mbar_wait(empty_bar(s), ph ^ 1u);
            const uint32_t lead_full = mapa_u32(full_bar(s), 0);
            mbar_arrive_expect_tx_cluster(lead_full, kStage);
            const uint32_t dst = base + (uint32_t)s * kStage;
            // channel-concatenated input read in place: the first chunks0 chunks from tmA, the rest from tmA1
            if (ch < g.chunks0)
              tma2_load_5d(dst, &tmA, lead_full, ch * kKC, u.x0 + dx - 1, u.y0 - 1, z, u.n);
            else
              tma2_load_5d(dst, &tmA1, lead_full, (ch - g.chunks0) * kKC, u.x0 + dx - 1, u.y0 - 1, z, u.n);
            // weights (Cin, 3*Cout rows, dy, dx, rot): this CTA's 96 rows of all three dy slices
            tma2_load_5d(dst + kASub, &tmB, lead_full, ch * kKC, (int)rank * kHalfRows, 0, dx, rot);
            if (++s == kStages) {
              s = 0;
              ph ^= 1u;
            }
          }
        }
      });
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA only) ===============
    if (rank == 0) {
      const uint32_t issue = elect_one();
      constexpr uint32_t desc_hi = (kSbo >> 4) | (1u << 14) | (kLayout << 29);
      const uint32_t lo_flag = 1u << 16;
      const uint32_t base16 = ((base & 0x3FFFFu) >> 4) | lo_flag;
      const uint32_t idesc = umma_idesc_16(256, kN3, F16);
      int s = 0;
      uint32_t ph = 0;
      for_each_tile(g, rank, [&](auto set_c, const Unit&, int p, uint32_t cnt) {
        constexpr uint32_t set = decltype(set_c)::value;
        mbar_wait(tempty_bar(set), (cnt & 1u) ^ 1u);   // both CTAs drained + zeroed the previous plane
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + set * kSetStride;
        uint32_t accum = p == 0 ? 0u : 1u;   // a unit's first plane overwrites the whole ring
        for (int si = 0; si < subs; ++si) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t a16 = base16 + (uint32_t)s * (kStage >> 4);
          const uint32_t b16 = a16 + (kASub >> 4);
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
            for (int kk = 0; kk < kSteps; ++kk) {
#pragma unroll
              for (int m = 0; m < kMT; ++m)   // bricks sharing this weight tile
                umma2_16_pred(d_tmem + (uint32_t)(m * kN3), a16 + (uint32_t)(16 * m + dy) * (kSbo >> 4) + 2u * kk,
                                b16 + (uint32_t)dy * (kBTile >> 4) + 2u * kk, desc_hi, idesc, accum, issue);
              accum = 1u;
            }
          }
          umma2_commit_pred(empty_bar(s), issue);   // frees the stage in both CTAs
          if (++s == kStages) {
            s = 0;
            ph ^= 1u;
          }
        }
        umma2_commit_pred(tfull_bar(set), issue);
      });
    }
  } else {
    // =============================== epilogue (8 warps, both CTAs) ===============
    // A thread owns one voxel row of the brick (TMEM lane) and 32 of the 64 output channels; the
    // bf16 row piece goes straight to global memory (two 256-bit stores), the GroupNorm statistics
    // stay in registers until the image changes.
    const int q = warp & 3;
    const int row = q * 32 + lane;        // tx = row & 7, ty = row >> 3
    const int half = (warp - 2) >> 2;     // column half [kCols half, kCols half + kCols)
    const int et = half * 128 + row;
    float* s_stats = reinterpret_cast<float*>(sm + g.off_stats);     // [N][Cout][2]
    float* s_wred = reinterpret_cast<float*>(sm + g.off_scratch);    // [8 warps][2 kCols] flush scratch
    const bool do_relu = (g.flags & KM_CONV_RELU) != 0;
    const bool do_stats = (g.flags & KM_CONV_STATS) != 0;
    auto all_bar = [&]() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    const int tx = row & 7, ty = row >> 3;
    float ssum[kCols], ssq[kCols];
#pragma unroll
    for (int j = 0; j < kCols; ++j) ssum[j] = ssq[j] = 0.f;
    uint32_t zprev[POOL ? 2 : 1][POOL ? kMT : 1][kCols / 16][8];   // xy-pooled even planes in flight (pool mode)
#pragma unroll
    for (int a = 0; a < (POOL ? 2 : 1); ++a)
#pragma unroll
      for (int m = 0; m < (POOL ? kMT : 1); ++m)
#pragma unroll
        for (int b = 0; b < kCols / 16; ++b)
#pragma unroll
          for (int j = 0; j < 8; ++j) zprev[a][m][b][j] = 0u;
    int n_cur = -1;
    auto flush_stats = [&](int n) {
#pragma unroll
      for (int j = 0; j < kCols; ++j) {
        const float a = km_warp_sum(ssum[j]), b = km_warp_sum(ssq[j]);
        if (lane == 0) {
          s_wred[(warp - 2) * 2 * kCols + 2 * j] = a;
          s_wred[(warp - 2) * 2 * kCols + 2 * j + 1] = b;
        }
        ssum[j] = ssq[j] = 0.f;
      }
      all_bar();
      if (et < 4 * kCols) {   // thread -> (column half h, column, sum | sumsq)
        const int h = et / (2 * kCols), i = et % (2 * kCols);
        float a = 0.f;
        for (int w4 = 0; w4 < 4; ++w4) a += s_wred[(h * 4 + w4) * 2 * kCols + i];
        s_stats[((size_t)n * kCout + h * kCols + (i >> 1)) * 2 + (i & 1)] += a;
      }
      all_bar();
    };

    for_each_tile(g, rank, [&](auto set_c, const Unit& u, int p, uint32_t cnt) {
      constexpr uint32_t set = decltype(set_c)::value;
      const int zo = u.zs - 2 + p;
      const bool store = p >= 2 && zo < g.D;
      const uint32_t slot = (uint32_t)((p + 2) % 3);
      if (do_stats && u.n != n_cur) {
        if (n_cur >= 0) flush_stats(n_cur);
        n_cur = u.n;
      }
      const uint32_t lead_tempty = mapa_u32(tempty_bar(set), 0);
      mbar_wait(tfull_bar(set), cnt & 1u);
      tc_fence_after();
#pragma unroll
      for (int m = 0; m < kMT; ++m) {
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + set * kSetStride + (uint32_t)(m * kN3) +
                             slot * kCout + (uint32_t)(half * kCols);
      const int x2 = u.x0 + tx, y2 = u.y0 + 16 * m + ty;
      if (store) {   // warp-uniform: tcgen05.ld is a warp-collective operation
        uint32_t r[kCols / 16][16];
        // ADD: partial sums of the same layer computed elsewhere (conv_up2.cu: the upsampled half of a decoder's
        // concat, on the coarse lattice).  Prefetched into L1 before the TMEM read (no registers held), loaded
        // 16 channels at a time where they are consumed.
        const uint16_t* ap = nullptr;
        if (ADD && u.valid && x2 < g.W && y2 < g.H) {
          ap = addend + ((((size_t)u.n * g.D + zo) * g.H + y2) * g.W + x2) * kCout + half * kCols;
          asm volatile("prefetch.global.L1 [%0];" ::"l"(ap));
        }
#pragma unroll
        for (int b = 0; b < kCols / 16; ++b) tmem_ld16(taddr + 16u * b, r[b]);
        tmem_ld_wait();
#pragma unroll
        for (int b = 0; b < kCols / 16; ++b) tmem_st16_zero(taddr + 16u * b);   // fresh output plane z' + 2
        if (m == kMT - 1) {   // last brick drained: the set goes back to the issuer before the arithmetic
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive_cluster(lead_tempty);
        }
        const bool inside = u.valid && x2 < g.W && y2 < g.H;
        const size_t vox = (((size_t)u.n * g.D + zo) * g.H + y2) * g.W + x2;
        if (ADD && ap) {
#pragma unroll
          for (int b = 0; b < kCols / 16; ++b) {
            const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(ap) + 2 * b);
            const uint4 a1 = __ldg(reinterpret_cast<const uint4*>(ap) + 2 * b + 1);
            const uint32_t a8[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 ab = km_unpack2<F16>(a8[j]);
              r[b][2 * j] = __float_as_uint(__uint_as_float(r[b][2 * j]) + ab.x);
              r[b][2 * j + 1] = __float_as_uint(__uint_as_float(r[b][2 * j + 1]) + ab.y);
            }
          }
        }
        if (bias_tab) {   // GroupNorm shift of the folded norm: bias[sample][border class][cout] (L1-resident)
          const int cls = ((zo == 0 ? 1 : 0) | (zo == g.D - 1 ? 2 : 0)) * 9 +
                          (y2 == 0 ? 1 : (y2 == g.H - 1 ? 2 : 0)) * 3 + (x2 == 0 ? 1 : (x2 == g.W - 1 ? 2 : 0));
          const float4* bp = reinterpret_cast<const float4*>(bias_tab + ((size_t)u.n * kBiasClasses + cls) * kCout +
                                                             half * kCols);
#pragma unroll
          for (int hh = 0; hh < kCols / 16; ++hh)
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 bv = __ldg(bp + hh * 4 + j4);
              r[hh][4 * j4 + 0] = __float_as_uint(__uint_as_float(r[hh][4 * j4 + 0]) + bv.x);
              r[hh][4 * j4 + 1] = __float_as_uint(__uint_as_float(r[hh][4 * j4 + 1]) + bv.y);
              r[hh][4 * j4 + 2] = __float_as_uint(__uint_as_float(r[hh][4 * j4 + 2]) + bv.z);
              r[hh][4 * j4 + 3] = __float_as_uint(__uint_as_float(r[hh][4 * j4 + 3]) + bv.w);
            }
        }
#pragma unroll
        for (int hh = 0; hh < kCols / 16; ++hh) {
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float a = __uint_as_float(r[hh][2 * j]);
            float b = __uint_as_float(r[hh][2 * j + 1]);
            if (do_relu) {
              a = fmaxf(a, 0.f);
              b = fmaxf(b, 0.f);
            }
            pk[j] = km_pack2<F16>(a, b);
          }
          if (out && inside) st_global_v8(out + vox * kCout + half * kCols + 16 * hh, pk);
          bool acc_stats = inside && !(POOL && pooled);
          if (POOL && pooled) {
            // MaxPool3d(2) in registers (see conv_zf.cu): x / y neighbours are lanes ^1 / ^8, the z
            // neighbour is the previous plane of this unit; max commutes with the bf16 rounding
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              pk[j] = km_max2<F16>(pk[j], __shfl_xor_sync(0xffffffffu, pk[j], 1));
              pk[j] = km_max2<F16>(pk[j], __shfl_xor_sync(0xffffffffu, pk[j], 8));
            }
            if ((zo & 1) == 0) {
#pragma unroll
              for (int j = 0; j < 8; ++j) zprev[POOL ? set : 0][POOL ? m : 0][hh][j] = pk[j];
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) pk[j] = km_max2<F16>(pk[j], zprev[POOL ? set : 0][POOL ? m : 0][hh][j]);
              const int xp = x2 >> 1, yp = y2 >> 1, zp = zo >> 1;
              if (u.valid && ((tx | ty) & 1) == 0 && xp < (g.W >> 1) && yp < (g.H >> 1)) {
                const size_t pv = (((size_t)u.n * (g.D >> 1) + zp) * (g.H >> 1) + yp) * (g.W >> 1) + xp;
                st_global_v8(pooled + pv * kCout + half * kCols + 16 * hh, pk);
                acc_stats = true;   // statistics of the pooled map
              }
            }
          }
          if (do_stats && acc_stats) {   // statistics of the values actually stored (bf16-rounded)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 ab = km_unpack2<F16>(pk[j]);
              const float ar = ab.x, br = ab.y;
              ssum[16 * hh + 2 * j] += ar;
              ssq[16 * hh + 2 * j] = fmaf(ar, ar, ssq[16 * hh + 2 * j]);
              ssum[16 * hh + 2 * j + 1] += br;
              ssq[16 * hh + 2 * j + 1] = fmaf(br, br, ssq[16 * hh + 2 * j + 1]);
            }
          }
        }
      } else {
#pragma unroll
        for (int b = 0; b < kCols / 16; ++b) tmem_st16_zero(taddr + 16u * b);
        if (m == kMT - 1) {
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive_cluster(lead_tempty);
        }
      }
      }
    });

    if (do_stats) {
      if (n_cur >= 0) flush_stats(n_cur);
      float* dst = stats + (size_t)blockIdx.x * g.N * kCout * 2;
      for (int i = et; i < g.N * kCout * 2; i += kEpiThreads) dst[i] = s_stats[i];
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// fp32 (Cout, Cin, 3, 3, 3) -> bf16 [rot][dx][dy][j*Cout + cout][Cin], dz(j, rot) = (rot + 1 - j) mod 3
template <bool F16>
__global__ void pack_weights_zf2_kernel(const float* __restrict__ w, uint16_t* __restrict__ p, int Cout,
                                        int Cin) {
  const long long total = 27ll * 3 * Cout * Cin;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int ci = (int)(t % Cin);
    t /= Cin;
    const int co = (int)(t % Cout);
    t /= Cout;
    const int j = (int)(t % 3);
    t /= 3;
    const int dy = (int)(t % 3);
    t /= 3;
    const int dx = (int)(t % 3);
    const int r = (int)(t / 3);
    const int dz = (r + 1 - j + 3) % 3;
    p[i] = km_from_float<F16>(w[((size_t)co * Cin + ci) * 27 + dz * 9 + dy * 3 + dx]);
  }
}


}  // namespace

extern "C" int km_sm_count(void);

extern "C" int km_conv3d_zfold_pair_supported(int Cin, int Cout, int D, int H, int W) {
  const bool shape = (Cout == 64 && Cin % 32 == 0) || (Cout == 32 && (Cin % 32 == 0 || Cin == 16));
  return (shape && Cin >= 16 && Cin <= 512 && W >= 8 && H >= 16 && D >= 1) ? 1 : 0;
}

extern "C" int km_pack_weights_zfold_pair(const float* w, void* packed, int Cout, int Cin, km_stream_t stream) {
  KM_CHECK_ARG(w && packed && (Cout == 64 || Cout == 32) && Cin % 16 == 0 && Cin > 0,
               "km_pack_weights_zfold_pair: needs Cout in {32, 64}, Cin %% 16 == 0");
  if (km_operand_fp16())
    pack_weights_zf2_kernel<true><<<128, 256, 0, km_cs(stream)>>>(w, reinterpret_cast<uint16_t*>(packed), Cout, Cin);
  else
    pack_weights_zf2_kernel<false><<<128, 256, 0, km_cs(stream)>>>(w, reinterpret_cast<uint16_t*>(packed), Cout, Cin);
  KM_LAUNCH_OK("pack_weights_zf2_kernel");
  return KM_OK;
}

namespace {
int g_zf2_mt2 = 0;   // km_set_option(KM_OPT_ZF2_TWO_BRICKS)
template <int COUT, int KC, int MT, bool POOL>
int launch_zf2(const void* x, const void* x1, int Cin0, const void* wz, const float* bias_tab, const void* addend,
               void* out, void* pooled, float* stats, int N, int Cin, int D, int H, int W, int flags, cudaStream_t st) {
  using C = Cfg<COUT, KC, MT>;
  KM_CHECK_ARG(POOL || !pooled, "km_conv3d_zfold_pair: fused pooling is built for Cout = 32 only");
  Zf2Geom g;
  memset(&g, 0, sizeof(g));
  g.N = N; g.D = D; g.H = H; g.W = W;
  g.chunks = Cin / KC;
  g.flags = flags;
  g.w_per_sample = bias_tab ? 1 : 0;
  if (!x1) Cin0 = Cin;
  KM_CHECK_ARG(Cin0 > 0 && Cin0 <= Cin && Cin0 % KC == 0 && (Cin - Cin0) % KC == 0,
               "km_conv3d_zfold_pair: the two inputs must split the channels at a multiple of %d (got %d + %d)", KC,
               Cin0, Cin - Cin0);
  g.chunks0 = Cin0 / KC;
  const int tiles_x = (W + 7) / 8;
  g.xpairs = (tiles_x + 1) / 2;
  g.mt = MT;
  g.sets = C::kSets;
  g.tiles_y = (H + 16 * MT - 1) / (16 * MT);
  // z segment length: longer segments amortise the two halo planes, shorter ones balance the unit
  // pairs over the CTA pairs; pick the cheaper of 64 / 32 / 16 planes for this shape
  const int npairs_hw = km_sm_count() / 2;
  double best = 1e30;
  for (int lz = 64; lz >= 16; lz /= 2) {
    const long long zs = (D + lz - 1) / lz;
    const long long upi = zs * g.tiles_y * g.xpairs;
    const long long rounds = (upi + npairs_hw - 1) / npairs_hw;
    const double cost = (double)rounds * (double)(std::min(lz, D) + 2);   // planes walked by the busiest pair
    if (cost < best) {
      best = cost;
      g.lz = lz;
    }
  }
  g.zsegs = (D + g.lz - 1) / g.lz;
  const long long punits = (long long)N * g.zsegs * g.tiles_y * g.xpairs;
  KM_CHECK_ARG(punits < (1ll << 30), "km_conv3d_zfold_pair: too many units");
  g.punits = (int)punits;
  const uint32_t stats_bytes = (flags & KM_CONV_STATS) ? (uint32_t)N * COUT * 2u * 4u : 0u;
  const uint32_t fixed = 8u * 2u * C::kCols * 4u + stats_bytes + 8u * (2u * kStagesMax + 6u) + 16u + 64u;
  const uint32_t kSmemMax = 232448 - 1024;
  int stages = (int)((kSmemMax - fixed) / C::kStage);
  if (stages > kStagesMax) stages = kStagesMax;
  KM_CHECK_ARG(stages >= 2, "km_conv3d_zfold_pair: shared memory budget exceeded (batch too large)");
  g.stages = stages;
  uint32_t off = (uint32_t)stages * C::kStage;
  g.off_scratch = off; off += 8u * 2u * C::kCols * 4u;
  g.off_stats = off; off += stats_bytes;
  off = (off + 7u) & ~7u;
  g.off_bars = off; off += 8u * (2u * kStagesMax + 6u) + 16u;
  const uint32_t smem_bytes = off + 1024;
  KM_CHECK_ARG(smem_bytes <= 232448, "km_conv3d_zfold_pair: shared memory overflow (%u)", smem_bytes);

  PFN_encodeTiled encode = tensor_map_encoder();
  if (!encode) {
    km_set_error("km_conv3d_zfold_pair: cuTensorMapEncodeTiled unavailable");
    return KM_ECUDA;
  }
  const CUtensorMapSwizzle swz = C::kRowBytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : C::kRowBytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUtensorMap tmA, tmA1, tmB;
  for (int which = 0; which < 2; ++which) {
    const int Cc = which == 0 ? Cin0 : (x1 ? Cin - Cin0 : Cin0);
    const void* ptr = which == 0 ? x : (x1 ? x1 : x);
    cuuint64_t dims[5] = {(cuuint64_t)Cc, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)Cc * 2, (cuuint64_t)W * Cc * 2, (cuuint64_t)H * W * Cc * 2,
                             (cuuint64_t)D * H * W * Cc * 2};
    cuuint32_t box[5] = {(cuuint32_t)KC, 8, (cuuint32_t)(16 * MT + 2), 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(which == 0 ? &tmA : &tmA1, KM_TMAP_16, 5, const_cast<void*>(ptr), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      km_set_error("km_conv3d_zfold_pair: cuTensorMapEncodeTiled(A%d) failed with %d", which, (int)r);
      return KM_ECUDA;
    }
  }
  {
    // packed weights [rot][dx][dy][3*Cout rows][Cin] viewed as (Cin, rows, dy, dx, rot): one box = the
    // three dy slices of one (rot, dx) for half of the rows
    const cuuint64_t tile = (cuuint64_t)C::kN3 * Cin * 2;
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)C::kN3, 3, 3, (cuuint64_t)(bias_tab ? 3 * N : 3)};
    cuuint64_t strides[4] = {(cuuint64_t)Cin * 2, tile, 3 * tile, 9 * tile};
    cuuint32_t box[5] = {(cuuint32_t)KC, (cuuint32_t)C::kHalfRows, 3, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tmB, KM_TMAP_16, 5, const_cast<void*>(wz), dims, strides, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      km_set_error("km_conv3d_zfold_pair: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
      return KM_ECUDA;
    }
  }
  static unsigned long long attr_set = 0;   // one mask per instantiation
  if (km_first_use_on_device(&attr_set))
{
    KM_CUDA_OK(cudaFuncSetAttribute(conv_zf2_kernel<COUT, KC, MT, POOL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    KM_CUDA_OK(cudaFuncSetAttribute(conv_zf2_kernel<COUT, KC, MT, POOL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  }
  const int nsm = km_sm_count();
  const int upi = g.punits / N;
  int grid = nsm & ~1;
  if (grid / 2 > upi) grid = 2 * upi;
  if ((flags & KM_CONV_STATS) && grid < nsm)
    KM_CUDA_OK(cudaMemsetAsync(stats, 0, (size_t)nsm * N * COUT * 2 * sizeof(float), st));
  uint16_t* o16 = reinterpret_cast<uint16_t*>(out);
  uint16_t* p16 = reinterpret_cast<uint16_t*>(pooled);
  const uint16_t* a16 = reinterpret_cast<const uint16_t*>(addend);
  if constexpr (COUT == 64 && KC == 64 && MT == 1 && !POOL) {
    if (addend) {   // the only shape the split decoder layer uses
      static unsigned long long attr_add = 0;
      if (km_first_use_on_device(&attr_add)) {
        KM_CUDA_OK(cudaFuncSetAttribute(conv_zf2_kernel<COUT, KC, MT, POOL, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        KM_CUDA_OK(cudaFuncSetAttribute(conv_zf2_kernel<COUT, KC, MT, POOL, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
      }
      if (km_operand_fp16())
        conv_zf2_kernel<COUT, KC, MT, POOL, true, true><<<grid, kThreads, smem_bytes, st>>>(tmA, tmA1, tmB, g, o16, p16, stats, bias_tab, a16);
      else
        conv_zf2_kernel<COUT, KC, MT, POOL, false, true><<<grid, kThreads, smem_bytes, st>>>(tmA, tmA1, tmB, g, o16, p16, stats, bias_tab, a16);
      KM_LAUNCH_OK("conv_zf2_kernel");
      return KM_OK;
    }
  }
  KM_CHECK_ARG(!addend, "km_conv3d_zfold_pair: an addend needs Cout = 64 and Cin %% 64 == 0");
  if (km_operand_fp16())
    conv_zf2_kernel<COUT, KC, MT, POOL, true><<<grid, kThreads, smem_bytes, st>>>(tmA, tmA1, tmB, g, o16, p16, stats, bias_tab, a16);
  else
    conv_zf2_kernel<COUT, KC, MT, POOL, false><<<grid, kThreads, smem_bytes, st>>>(tmA, tmA1, tmB, g, o16, p16, stats, bias_tab, a16);
  KM_LAUNCH_OK("conv_zf2_kernel");
  return KM_OK;
}
}  // namespace

void km_zf2_set_two_bricks(int v) { g_zf2_mt2 = v ? 1 : 0; }

namespace {
int dispatch_zf2(const char* who, const void* x, const void* x1, int Cin0, const void* wz, const float* bias_tab,
                 const void* addend, void* out, void* pooled, float* stats, int N, int Cin, int Cout, int D, int H,
                 int W, int flags, km_stream_t stream) {
  KM_CHECK_ARG(x && wz && (out || pooled), "%s: null argument", who);
  KM_CHECK_ARG(km_conv3d_zfold_pair_supported(Cin, Cout, D, H, W), "%s: unsupported shape (Cin=%d Cout=%d H=%d W=%d)",
               who, Cin, Cout, H, W);
  KM_CHECK_ARG(N > 0, "%s: bad batch", who);
  KM_CHECK_ARG(!pooled || (D >= 2 && H >= 2 && W >= 2), "%s: volume too small to pool", who);
  KM_CHECK_ARG(!(flags & KM_CONV_STATS) || stats, "%s: KM_CONV_STATS needs stats", who);
  KM_CHECK_ARG(!(flags & KM_CONV_COM), "%s: KM_CONV_COM is not supported", who);
  KM_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)x1 & 15) == 0 && ((uintptr_t)wz & 15) == 0 &&
                   ((uintptr_t)out & 31) == 0 && ((uintptr_t)pooled & 31) == 0 && ((uintptr_t)addend & 31) == 0,
               "%s: pointers must be 16-byte (outputs: 32-byte) aligned", who);
  cudaStream_t st = km_cs(stream);
  if (Cin == 16) return launch_zf2<32, 16, 1, true>(x, x1, Cin0, wz, bias_tab, addend, out, pooled, stats, N, Cin, D, H, W, flags, st);
  const bool k64 = Cin % 64 == 0 && (!x1 || Cin0 % 64 == 0);
  if (Cout == 64 && k64) {
    // two bricks per unit halve the weight bytes per output (the kernel is bound by L2 -> SM traffic)
    // at the price of a single TMEM set; needs two brick rows
    if (H >= 32 && g_zf2_mt2 && !addend)
      return launch_zf2<64, 64, 2, false>(x, x1, Cin0, wz, bias_tab, addend, out, pooled, stats, N, Cin, D, H, W, flags, st);
    return launch_zf2<64, 64, 1, false>(x, x1, Cin0, wz, bias_tab, addend, out, pooled, stats, N, Cin, D, H, W, flags, st);
  }
  if (Cout == 64) return launch_zf2<64, 32, 1, false>(x, x1, Cin0, wz, bias_tab, addend, out, pooled, stats, N, Cin, D, H, W, flags, st);
  return k64 ? launch_zf2<32, 64, 1, true>(x, x1, Cin0, wz, bias_tab, addend, out, pooled, stats, N, Cin, D, H, W, flags, st)
             : launch_zf2<32, 32, 1, true>(x, x1, Cin0, wz, bias_tab, addend, out, pooled, stats, N, Cin, D, H, W, flags, st);
}
}  // namespace

extern "C" int km_conv3d_zfold_pair(const void* x, const void* wz, void* out, void* pooled, float* stats,
                                    int N, int Cin, int Cout, int D, int H, int W, int flags,
                                    km_stream_t stream) {
  return dispatch_zf2("km_conv3d_zfold_pair", x, nullptr, Cin, wz, nullptr, nullptr, out, pooled, stats, N, Cin, Cout,
                      D, H, W, flags, stream);
}

extern "C" size_t km_conv3d_zfold_pair_gn_workspace_bytes(int N, int Cin, int Cout) {
  const size_t wbytes = ((size_t)27 * 3 * Cout * Cin * 2 + 255) & ~(size_t)255;
  return (size_t)N * wbytes + (size_t)N * kBiasClasses * Cout * 4;
}

extern "C" int km_conv3d_zfold_pair_gn_cat(const void* x0, const void* x1, int Cin0, int Cin1, const float* w,
                                           const float* scale, const float* shift, void* out, void* pooled,
                                           float* stats, void* workspace, int N, int Cout, int D, int H, int W,
                                           int flags, km_stream_t stream) {
  const int Cin = Cin0 + (x1 ? Cin1 : 0);
  KM_CHECK_ARG(w && scale && shift && workspace && ((uintptr_t)workspace & 255) == 0,
               "km_conv3d_zfold_pair_gn: null / unaligned (256 B) argument");
  KM_CHECK_ARG(km_conv3d_zfold_pair_supported(Cin, Cout, D, H, W),
               "km_conv3d_zfold_pair_gn: unsupported shape (Cin=%d Cout=%d H=%d W=%d)", Cin, Cout, H, W);
  KM_CHECK_ARG(N > 0 && N <= 1024, "km_conv3d_zfold_pair_gn: bad batch");
  // the per-sample weight sets must be contiguous for the 5-D tensor map: (27 * 3 * Cout * Cin * 2) bytes each
  const size_t wbytes = (size_t)27 * 3 * Cout * Cin * 2;
  void* packed = workspace;
  float* bias = reinterpret_cast<float*>(static_cast<char*>(workspace) + (((size_t)N * wbytes + 255) & ~(size_t)255));
  const int rf = km_fold_gn(w, scale, shift, packed, bias, N, Cout, Cin, 1, stream);
  if (rf != KM_OK) return rf;
  return dispatch_zf2("km_conv3d_zfold_pair_gn", x0, x1, Cin0, workspace, bias, nullptr, out, pooled, stats, N, Cin,
                      Cout, D, H, W, flags, stream);
}

// The first Cs channels of a (Cs + Cu)-channel folded layer; `addend` holds the other channels' partial sums
// (km_conv3d_up2_gn).  The border-class bias table covers ALL channels, the packed weights the first Cs.
extern "C" int km_conv3d_zfold_pair_gn_add(const void* x, int Cs, int Cu, const float* w, const float* scale,
                                           const float* shift, const void* addend, void* out, float* stats,
                                           void* workspace, int N, int Cout, int D, int H, int W, int flags,
                                           km_stream_t stream) {
  KM_CHECK_ARG(w && scale && shift && addend && workspace && ((uintptr_t)workspace & 255) == 0,
               "km_conv3d_zfold_pair_gn_add: null / unaligned (256 B) argument");
  KM_CHECK_ARG(Cs > 0 && Cs % 64 == 0 && Cout == 64 && Cu > 0 && km_conv3d_zfold_pair_supported(Cs, Cout, D, H, W),
               "km_conv3d_zfold_pair_gn_add: unsupported shape (Cs=%d Cu=%d Cout=%d H=%d W=%d)", Cs, Cu, Cout, H, W);
  KM_CHECK_ARG(N > 0 && N <= 1024, "km_conv3d_zfold_pair_gn_add: bad batch");
  const size_t wbytes = (size_t)27 * 3 * Cout * Cs * 2;
  float* bias = reinterpret_cast<float*>(static_cast<char*>(workspace) + (((size_t)N * wbytes + 255) & ~(size_t)255));
  const int rf = km_fold_gn_part(w, scale, shift, workspace, bias, N, Cout, Cs + Cu, 0, Cs, 1, stream);
  if (rf != KM_OK) return rf;
  return dispatch_zf2("km_conv3d_zfold_pair_gn_add", x, nullptr, Cs, workspace, bias, addend, out, nullptr, stats, N,
                      Cs, Cout, D, H, W, flags, stream);
}

extern "C" int km_conv3d_zfold_pair_gn(const void* x, const float* w, const float* scale, const float* shift,
                                       void* out, void* pooled, float* stats, void* workspace, int N, int Cin,
                                       int Cout, int D, int H, int W, int flags, km_stream_t stream) {
  return km_conv3d_zfold_pair_gn_cat(x, nullptr, Cin, 0, w, scale, shift, out, pooled, stats, workspace, N, Cout, D,
                                     H, W, flags, stream);
}
