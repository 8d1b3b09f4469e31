// "z-folded" tcgen05 implicit-GEMM 3x3x3 convolution for the backbone's first tensor-core layer
// (Cin = 16 -> Cout = 32 at full resolution: keymorph/unet3d/buildingblocks.py:50-52, second
// SingleConv of encoder 0).
//
// Why a second kernel: with Cin = 16 every tcgen05.mma of the general kernel (conv_tc.cu) is a
// 128 x 32 x 16 sliver: 27 of them per 128 voxels, each re-reading a 4 KB activation slice from
// shared memory and each fed by its own piece of nine TMA boxes per brick group.  That layer ran
// at ~0.48 PFLOP/s.  Here the three dz taps are folded into the N dimension instead:
//
//     D[voxel(z'), (j, cout)] += A[voxel(z') shifted by (dy, dx), cin] . W[dz(j), dy, dx][cout][cin]
//
// i.e. one plane z' of the input contributes, with ONE pass over its data and 9 MMAs of N = 96
// per brick, to the three output planes z'+1 (dz=0), z' (dz=1) and z'-1 (dz=2).  The three column
// blocks j of the accumulator are a ring of output planes that lives in TMEM across the CTA's walk
// along z: block (p+2)%3 is complete after plane p, the epilogue drains it (ReLU, bf16, stats) and
// zeroes it (tcgen05.st) so that it becomes the fresh plane z'+2 of the next step; all MMAs after
// a unit's first plane accumulate.  Because the ring rotates, the weights are kept in the three
// row rotations r = p % 3 (dz(j, r) = (r + 1 - j) mod 3), all resident in shared memory (81 KB).
// Per plane: 3 TMA activation boxes (dx) instead of 9 (dz, dx), a third of the MMAs, a third of the
// shared-memory operand reads.
//
// A work unit is a column of bricks: 8(x) x 32(y) voxels (two 8x16 bricks sharing every weight
// slice, y taps through UMMA descriptor offsets exactly as MODE 2 of conv_tc.cu) x a z segment of
// up to 64 planes (+2 halo planes, zero-filled by TMA outside the volume).  A CTA interleaves two
// units plane by plane (TMEM sets 0 / 1) so that the epilogue of one overlaps the MMAs of the other.
// Roles (320 threads, one persistent CTA per SM): warp 0 = TMA producer, warp 1 = MMA issuer,
// warps 2..9 = epilogue.
#include <type_traits>
#include "km_common.cuh"
#include "tc_ptx.cuh"

using namespace kmtc;

namespace {

constexpr int kThreads = 320;
constexpr int kEpiThreads = 256;
constexpr int kKC = 16;                 // input channels (one K step)
constexpr int kCout = 32;
constexpr int kN3 = 3 * kCout;          // MMA N: three output planes x Cout
constexpr int kMT = 2;                  // bricks per unit (stacked in y)
constexpr int kRowBytes = kKC * 2;      // 32 B rows, SWIZZLE_32B
constexpr int kBoxRows = 8 * (16 * kMT + 2);
constexpr uint32_t kASub = 9216;        // 272 rows x 32 B = 8704, rounded up to 1 KB
constexpr uint32_t kAStage = 3 * kASub; // one plane: dx = 0, 1, 2
constexpr uint32_t kBTile = kN3 * kRowBytes;          // 3072 B
constexpr uint32_t kBBytes = 27 * kBTile;             // [r][dx][dy] = 82944 B
constexpr int kStages = 5;
constexpr int kLZ = 64;                 // z planes per unit

struct ZfGeom {
  int N, D, H, W;
  int n_base, n_count;          // samples handled by this launch
  int cta_per_sample;           // > 0 (GN-folded path): CTAs [k cps, (k+1) cps) own sample k and its weights
  int tiles_x, tiles_y, zsegs, units;
  int flags;
  uint32_t off_b, off_staging, off_stats, off_bias, off_bars;
  int stat_parts;
};

constexpr int kBiasClasses = 36;        // z code (0..3) x y code (0..2) x x code (0..2)

struct Unit {
  int n, x0, y0, zs, planes;   // planes = segment length + 2
};

__device__ __forceinline__ Unit decode_unit(const ZfGeom& g, int u) {
  Unit r;
  r.x0 = (u % g.tiles_x) * 8;
  u /= g.tiles_x;
  r.y0 = (u % g.tiles_y) * (16 * kMT);
  u /= g.tiles_y;
  r.zs = (u % g.zsegs) * kLZ;
  r.n = g.n_base + u / g.zsegs;
  r.planes = min(kLZ, g.D - r.zs) + 2;
  return r;
}

// The tile sequence of this CTA, identical for the three roles: units blockIdx.x + k*gridDim.x are
// taken in pairs and interleaved plane by plane; fn(set, unit, p, cnt) with cnt = tiles already
// done on that TMEM set.
template <typename F>
__device__ __forceinline__ void for_each_tile(const ZfGeom& g, F&& fn) {
  uint32_t cnt[2] = {0u, 0u};
  int G = (int)gridDim.x, first = (int)blockIdx.x, end = g.units;
  if (g.cta_per_sample > 0) {
    const int ups = g.zsegs * g.tiles_y * g.tiles_x, ns = (int)blockIdx.x / g.cta_per_sample;
    G = g.cta_per_sample;
    first = ns * ups + (int)blockIdx.x % g.cta_per_sample;
    end = (ns + 1) * ups;
  }
  for (int ua = first; ua < end; ua += 2 * G) {
    const int ub = ua + G;
    const Unit a = decode_unit(g, ua);
    Unit b = a;
    b.planes = 0;
    if (ub < end) b = decode_unit(g, ub);
    const int pmax = max(a.planes, b.planes);
    for (int p = 0; p < pmax; ++p) {
      if (p < a.planes) fn(std::integral_constant<uint32_t, 0u>{}, a, p, cnt[0]++);
      if (p < b.planes) fn(std::integral_constant<uint32_t, 1u>{}, b, p, cnt[1]++);
    }
  }
}


// BIAS: the input is the RAW (un-normalised) activation and tmB holds weights pre-multiplied by the
// sample's GroupNorm scale; the GroupNorm shift enters as bias_tab[class][cout], class = which of the
// 27 taps fall outside the volume for this voxel (zero padding applies to the NORMALISED input).
template <bool BIAS, bool F16>
__global__ void __launch_bounds__(kThreads, 1)
conv_zf_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const ZfGeom g, uint16_t* __restrict__ out, uint16_t* __restrict__ pooled,
               float* __restrict__ stats, const float* __restrict__ bias_tab) {
  constexpr uint32_t kLayout = 6u;                 // SWIZZLE_32B
  constexpr uint32_t kSbo = 8u * kRowBytes;        // 256 B between 8-row groups
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_u32);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t bars = base + g.off_bars;   // full[S], empty[S], tfull[2], tempty[2], wfull
  auto full_bar = [&](int s) { return bars + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bars + 8u * (uint32_t)(kStages + s); };
  auto tfull_bar = [&](uint32_t a) { return bars + 8u * (uint32_t)(2 * kStages + a); };
  auto tempty_bar = [&](uint32_t a) { return bars + 8u * (uint32_t)(2 * kStages + 2 + a); };
  const uint32_t w_bar = bars + 8u * (uint32_t)(2 * kStages + 4);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sm + g.off_bars + 8u * (2 * kStages + 5));

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (uint32_t a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kEpiThreads);
    }
    mbar_init(w_bar, 1);
    fence_mbar_init();
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  __syncwarp();
  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_ptr_smem), 512);
    tmem_relinquish();
  }
  {
    float* s_stats = reinterpret_cast<float*>(sm + g.off_stats);
    if (g.flags & KM_CONV_STATS)
      for (int i = threadIdx.x; i < g.N * kCout * 2; i += kThreads) s_stats[i] = 0.f;
    if (BIAS) {
      float* s_bias = reinterpret_cast<float*>(sm + g.off_bias);
      const float* src = bias_tab + (g.cta_per_sample > 0 ? (size_t)(blockIdx.x / g.cta_per_sample) * kBiasClasses * kCout : 0);
      for (int i = threadIdx.x; i < kBiasClasses * kCout; i += kThreads) s_bias[i] = src[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      mbar_arrive_expect_tx(w_bar, kBBytes);
      const int t0 = g.cta_per_sample > 0 ? 27 * ((int)blockIdx.x / g.cta_per_sample) : 0;   // this sample's weights
      for (int t = 0; t < 27; ++t) tma_load_3d(base + g.off_b + (uint32_t)t * kBTile, &tmB, w_bar, 0, 0, t0 + t);
      int s = 0;
      uint32_t ph = 0;
      for_each_tile(g, [&](auto, const Unit& u, int p, uint32_t) {
        const int z = u.zs - 1 + p;   // input plane; outside [0, D) -> TMA zero fill
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_arrive_expect_tx(full_bar(s), 3u * (uint32_t)(kBoxRows * kRowBytes));
        const uint32_t dst = base + (uint32_t)s * kAStage;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
          tma_load_5d(dst + (uint32_t)dx * kASub, &tmA, full_bar(s), 0, u.x0 + dx - 1, u.y0 - 1, z, u.n);
        if (++s == kStages) {
          s = 0;
          ph ^= 1u;
        }
      });
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    const uint32_t issue = elect_one();
    constexpr uint32_t desc_hi = (kSbo >> 4) | (1u << 14) | (kLayout << 29);
    const uint32_t lo_flag = 1u << 16;
    const uint32_t a_base16 = ((base & 0x3FFFFu) >> 4) | lo_flag;
    const uint32_t b_base16 = (((base + g.off_b) & 0x3FFFFu) >> 4) | lo_flag;
    const uint32_t idesc = umma_idesc_16(128, kN3, F16);
    int s = 0;
    uint32_t ph = 0;
    mbar_wait(w_bar, 0u);
    for_each_tile(g, [&](auto set_c, const Unit&, int p, uint32_t cnt) {
      constexpr uint32_t set = decltype(set_c)::value;
      mbar_wait(tempty_bar(set), (cnt & 1u) ^ 1u);   // previous plane of this set drained + zeroed
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + set * 256u;
      const uint32_t a16 = a_base16 + (uint32_t)s * (kAStage >> 4);
      const uint32_t b16 = b_base16 + (uint32_t)(p % 3) * 9u * (kBTile >> 4);
      uint32_t accum = p == 0 ? 0u : 1u;   // a unit's first plane overwrites the whole ring
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const uint32_t bb = b16 + (uint32_t)(dx * 3 + dy) * (kBTile >> 4);
#pragma unroll
          for (int m = 0; m < kMT; ++m) {
            // rows of brick m shifted by dy: (16 m + dy) atoms of 8 rows x 32 B
            const uint32_t aa = a16 + (uint32_t)dx * (kASub >> 4) + (uint32_t)(16 * m + dy) * (kSbo >> 4);
            umma_16_pred(d_tmem + (uint32_t)m * kN3, aa, bb, desc_hi, idesc, accum, issue);
          }
          accum = 1u;
        }
      }
      umma_commit_pred(empty_bar(s), issue);
      umma_commit_pred(tfull_bar(set), issue);
      if (++s == kStages) {
        s = 0;
        ph ^= 1u;
      }
    });
  } else {
    // =============================== epilogue (8 warps) ==========================
    // A thread owns one voxel row of each brick (TMEM lane) and 16 of the 32 output channels.
    // Nothing is staged: the bf16 row piece (32 B = one full sector) goes straight to global
    // memory, and the GroupNorm statistics of the stored values are accumulated per thread in
    // registers over all tiles of an image; they are reduced across the CTA only when the image
    // changes (fixed order: shuffles, then the four warps of a column half one after the other).
    const int q = warp & 3;               // TMEM lane quadrant
    const int row = q * 32 + lane;        // voxel within a brick: tx = row & 7, ty = row >> 3
    const int half = (warp - 2) >> 2;     // column half [16 half, 16 half + 16)
    const int et = half * 128 + row;
    float* s_stats = reinterpret_cast<float*>(sm + g.off_stats);     // [N][Cout][2]
    float* s_wred = reinterpret_cast<float*>(sm + g.off_staging);    // [8 warps][32] flush scratch
    const bool do_relu = (g.flags & KM_CONV_RELU) != 0;
    const bool do_stats = (g.flags & KM_CONV_STATS) != 0;
    auto all_bar = [&]() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    const int tx = row & 7, ty = row >> 3;
    float ssum[16], ssq[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) ssum[j] = ssq[j] = 0.f;
    uint32_t zprev[2][kMT][8];   // xy-pooled even plane of each unit in flight (pool mode)
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int m = 0; m < kMT; ++m)
#pragma unroll
        for (int j = 0; j < 8; ++j) zprev[a][m][j] = 0u;
    int n_cur = -1;
    auto flush_stats = [&](int n) {
      // per-thread sums -> s_stats[n][half*16 + j][2]
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float a = km_warp_sum(ssum[j]), b = km_warp_sum(ssq[j]);
        if (lane == 0) {
          s_wred[(warp - 2) * 32 + 2 * j] = a;
          s_wred[(warp - 2) * 32 + 2 * j + 1] = b;
        }
        ssum[j] = ssq[j] = 0.f;
      }
      all_bar();
      if (et < 64) {   // thread -> (column half h, column j, sum | sumsq)
        const int h = et >> 5, i = et & 31;
        float a = 0.f;
        for (int w4 = 0; w4 < 4; ++w4) a += s_wred[(h * 4 + w4) * 32 + i];
        s_stats[((size_t)n * kCout + h * 16 + (i >> 1)) * 2 + (i & 1)] += a;
      }
      all_bar();
    };

    for_each_tile(g, [&](auto set_c, const Unit& u, int p, uint32_t cnt) {
      constexpr uint32_t set = decltype(set_c)::value;
      const int zo = u.zs - 2 + p;                   // the output plane completed by this tile
      const bool store = p >= 2 && zo < g.D;         // (zo >= zs by construction)
      const uint32_t slot = (uint32_t)((p + 2) % 3);
      const int zcode = (zo == 0 ? 1 : 0) | (zo == g.D - 1 ? 2 : 0);
      if (do_stats && u.n != n_cur) {                // CTA-uniform: the tile sequence is shared
        if (n_cur >= 0) flush_stats(n_cur);
        n_cur = u.n;
      }
      mbar_wait(tfull_bar(set), cnt & 1u);
      tc_fence_after();
#pragma unroll
      for (int m = 0; m < kMT; ++m) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + set * 256u + (uint32_t)m * kN3 +
                               slot * kCout + (uint32_t)half * 16u;
        const int x2 = u.x0 + tx, y2 = u.y0 + 16 * m + ty;
        if (store) {   // warp-uniform: tcgen05.ld is a warp-collective operation
          uint32_t r[16];
          tmem_ld16(taddr, r);
          tmem_ld_wait();
          if (BIAS) {
            const int cls = zcode * 9 + (y2 == 0 ? 1 : (y2 == g.H - 1 ? 2 : 0)) * 3 +
                            (x2 == 0 ? 1 : (x2 == g.W - 1 ? 2 : 0));
            const float4* bp = reinterpret_cast<const float4*>(sm + g.off_bias) + cls * (kCout / 4) + half * 4;
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 bv = bp[j4];
              r[4 * j4 + 0] = __float_as_uint(__uint_as_float(r[4 * j4 + 0]) + bv.x);
              r[4 * j4 + 1] = __float_as_uint(__uint_as_float(r[4 * j4 + 1]) + bv.y);
              r[4 * j4 + 2] = __float_as_uint(__uint_as_float(r[4 * j4 + 2]) + bv.z);
              r[4 * j4 + 3] = __float_as_uint(__uint_as_float(r[4 * j4 + 3]) + bv.w);
            }
          }
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float a = __uint_as_float(r[2 * j]), b = __uint_as_float(r[2 * j + 1]);
            if (do_relu) {
              a = fmaxf(a, 0.f);
              b = fmaxf(b, 0.f);
            }
            pk[j] = km_pack2<F16>(a, b);
          }
          const bool inside = x2 < g.W && y2 < g.H;
          if (out && inside) {
            // one 256-bit store = one full 32-byte sector per lane (two 16-byte stores would be two
            // half-sector requests on the L1 -> crossbar path, which then bounds the kernel)
            const size_t vox = (((size_t)u.n * g.D + zo) * g.H + y2) * g.W + x2;
            st_global_v8(out + vox * kCout + half * 16, pk);
          }
          bool acc_stats = inside && !pooled;
          if (pooled) {
            // MaxPool3d(2) (keymorph/unet3d/buildingblocks.py:363,387) in registers: x / y
            // neighbours are lanes ^1 / ^8 of the same warp, the z neighbour is the previous plane
            // of this unit (kept per TMEM set); max commutes with the bf16 rounding
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              pk[j] = km_max2<F16>(pk[j], __shfl_xor_sync(0xffffffffu, pk[j], 1));
              pk[j] = km_max2<F16>(pk[j], __shfl_xor_sync(0xffffffffu, pk[j], 8));
            }
            if ((zo & 1) == 0) {
#pragma unroll
              for (int j = 0; j < 8; ++j) zprev[set][m][j] = pk[j];
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) pk[j] = km_max2<F16>(pk[j], zprev[set][m][j]);
              const int xp = x2 >> 1, yp = y2 >> 1, zp = zo >> 1;
              if (((tx | ty) & 1) == 0 && xp < (g.W >> 1) && yp < (g.H >> 1)) {
                const size_t vox = (((size_t)u.n * (g.D >> 1) + zp) * (g.H >> 1) + yp) * (g.W >> 1) + xp;
                st_global_v8(pooled + vox * kCout + half * 16, pk);
                acc_stats = true;   // statistics of the pooled map (what the next GroupNorm needs)
              }
            }
          }
          if (do_stats && acc_stats) {   // statistics of the values actually stored (bf16-rounded)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 ab = km_unpack2<F16>(pk[j]);
              const float ar = ab.x, br = ab.y;
              ssum[2 * j] += ar;
              ssq[2 * j] = fmaf(ar, ar, ssq[2 * j]);
              ssum[2 * j + 1] += br;
              ssq[2 * j + 1] = fmaf(br, br, ssq[2 * j + 1]);
            }
          }
        }
        tmem_st16_zero(taddr);   // the drained block becomes the fresh output plane z' + 2
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(tempty_bar(set));
    });

    if (do_stats) {
      if (n_cur >= 0) flush_stats(n_cur);
      float* dst = stats + (size_t)blockIdx.x * g.N * kCout * 2;
      for (int i = g.n_base * kCout * 2 + et; i < (g.n_base + g.n_count) * kCout * 2; i += kEpiThreads)
        dst[i] = s_stats[i];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// fp32 (Cout, Cin, 3, 3, 3) -> bf16 [r][dx][dy][j][Cout][Cin], dz(j, r) = (r + 1 - j) mod 3
template <bool F16>
__global__ void pack_weights_zf_kernel(const float* __restrict__ w, uint16_t* __restrict__ p, int Cout,
                                       int Cin) {
  const int total = 27 * 3 * Cout * Cin;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int t = i;
    const int ci = t % Cin;
    t /= Cin;
    const int co = t % Cout;
    t /= Cout;
    const int j = t % 3;
    t /= 3;
    const int dy = t % 3;
    t /= 3;
    const int dx = t % 3;
    const int r = t / 3;
    const int dz = (r + 1 - j + 3) % 3;
    p[i] = km_from_float<F16>(w[((size_t)co * Cin + ci) * 27 + dz * 9 + dy * 3 + dx]);
  }
}


inline uint32_t zf_round_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

}  // namespace

extern "C" int km_sm_count(void);

extern "C" int km_conv3d_zfold_supported(int Cin, int Cout, int D, int H, int W) {
  return (Cin == kKC && Cout == kCout && W >= 8 && H >= 16 && D >= 1) ? 1 : 0;
}

extern "C" size_t km_pack_weights_zfold_bytes(int Cout, int Cin) {
  return (size_t)27 * 3 * Cout * Cin * 2;
}

extern "C" int km_pack_weights_zfold(const float* w, void* packed, int Cout, int Cin, km_stream_t stream) {
  KM_CHECK_ARG(w && packed && Cout == kCout && Cin == kKC, "km_pack_weights_zfold: needs Cin=%d, Cout=%d", kKC, kCout);
  if (km_operand_fp16())
    pack_weights_zf_kernel<true><<<64, 256, 0, km_cs(stream)>>>(w, reinterpret_cast<uint16_t*>(packed), Cout, Cin);
  else
    pack_weights_zf_kernel<false><<<64, 256, 0, km_cs(stream)>>>(w, reinterpret_cast<uint16_t*>(packed), Cout, Cin);
  KM_LAUNCH_OK("pack_weights_zf_kernel");
  return KM_OK;
}

namespace {

// one launch over samples [n_base, n_base + n_count) of an N-sample tensor
// (per_sample: wz / bias_tab hold n_count consecutive weight sets / tables and the CTAs are partitioned by sample)
int launch_zf(const void* x, const void* wz, const float* bias_tab, void* out, void* pooled, float* stats, int N,
              int n_base, int n_count, int Cin, int D, int H, int W, int flags, bool per_sample,
              km_stream_t stream) {
  ZfGeom g;
  memset(&g, 0, sizeof(g));
  g.N = N; g.D = D; g.H = H; g.W = W;
  g.n_base = n_base; g.n_count = n_count;
  g.flags = flags;
  g.tiles_x = (W + 7) / 8;
  g.tiles_y = (H + 16 * kMT - 1) / (16 * kMT);
  g.zsegs = (D + kLZ - 1) / kLZ;
  const long long units = (long long)n_count * g.zsegs * g.tiles_y * g.tiles_x;
  KM_CHECK_ARG(units < (1ll << 30), "km_conv3d_zfold: too many units");
  g.units = (int)units;
  g.stat_parts = 1;
  const uint32_t stats_bytes = (flags & KM_CONV_STATS) ? (uint32_t)N * kCout * 2u * 4u : 0u;
  uint32_t off = (uint32_t)kStages * kAStage;
  g.off_b = off; off += kBBytes;
  g.off_staging = off; off += 8u * 32u * 4u;   // per-warp scratch of the statistics flush
  g.off_bias = off; off += bias_tab ? (uint32_t)kBiasClasses * kCout * 4u : 0u;
  g.off_stats = off; off += stats_bytes;
  off = zf_round_up(off, 8);
  g.off_bars = off; off += 8u * (2u * kStages + 6u) + 16u;
  const uint32_t smem_bytes = off + 1024;
  KM_CHECK_ARG(smem_bytes <= 232448, "km_conv3d_zfold: shared memory overflow (%u; batch too large)", smem_bytes);

  PFN_encodeTiled encode = tensor_map_encoder();
  if (!encode) {
    km_set_error("km_conv3d_zfold: cuTensorMapEncodeTiled unavailable");
    return KM_ECUDA;
  }
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2,
                             (cuuint64_t)D * H * W * Cin * 2};
    cuuint32_t box[5] = {(cuuint32_t)kKC, 8, (cuuint32_t)(16 * kMT + 2), 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tmA, KM_TMAP_16, 5, const_cast<void*>(x), dims, strides, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      km_set_error("km_conv3d_zfold: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
      return KM_ECUDA;
    }
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)kN3, (cuuint64_t)(per_sample ? 27 * n_count : 27)};
    cuuint64_t strides[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)kN3 * Cin * 2};
    cuuint32_t box[3] = {(cuuint32_t)kKC, (cuuint32_t)kN3, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&tmB, KM_TMAP_16, 3, const_cast<void*>(wz), dims, strides, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      km_set_error("km_conv3d_zfold: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
      return KM_ECUDA;
    }
  }
  static unsigned long long attr_set = 0;
  if (km_first_use_on_device(&attr_set)) {
    KM_CUDA_OK(cudaFuncSetAttribute(conv_zf_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    KM_CUDA_OK(cudaFuncSetAttribute(conv_zf_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    KM_CUDA_OK(cudaFuncSetAttribute(conv_zf_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    KM_CUDA_OK(cudaFuncSetAttribute(conv_zf_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  }
  const int nsm = km_sm_count();
  int grid = g.units < nsm ? g.units : nsm;
  if (per_sample) {
    const int ups = g.units / n_count;
    g.cta_per_sample = nsm / n_count < ups ? nsm / n_count : ups;
    grid = g.cta_per_sample * n_count;
  }
  if (grid < nsm && (flags & KM_CONV_STATS) && n_base == 0)   // partial slots of the CTAs that do not run
    KM_CUDA_OK(cudaMemsetAsync(stats, 0, (size_t)nsm * N * kCout * 2 * sizeof(float), km_cs(stream)));
  uint16_t* o16 = reinterpret_cast<uint16_t*>(out);
  uint16_t* p16 = reinterpret_cast<uint16_t*>(pooled);
  const bool f16 = km_operand_fp16() != 0;
#define KM_ZF(B, F) conv_zf_kernel<B, F><<<grid, kThreads, smem_bytes, km_cs(stream)>>>(tmA, tmB, g, o16, p16, stats, bias_tab)
  if (bias_tab) {
    if (f16) KM_ZF(true, true);
    else KM_ZF(true, false);
  } else {
    if (f16) KM_ZF(false, true);
    else KM_ZF(false, false);
  }
#undef KM_ZF
  KM_LAUNCH_OK("conv_zf_kernel");
  return KM_OK;
}

int check_zf_args(const char* who, const void* x, const void* w, void* out, void* pooled, float* stats, int N,
                  int Cin, int Cout, int D, int H, int W, int flags) {
  KM_CHECK_ARG(x && w && (out || pooled), "%s: null argument", who);
  KM_CHECK_ARG(!pooled || (D >= 2 && H >= 2 && W >= 2), "%s: volume too small to pool", who);
  KM_CHECK_ARG(((uintptr_t)pooled & 31) == 0 && ((uintptr_t)out & 31) == 0, "%s: outputs must be 32-byte aligned", who);
  KM_CHECK_ARG(km_conv3d_zfold_supported(Cin, Cout, D, H, W), "%s: unsupported shape (Cin=%d Cout=%d H=%d W=%d)", who,
               Cin, Cout, H, W);
  KM_CHECK_ARG(N > 0, "%s: bad batch", who);
  KM_CHECK_ARG(!(flags & KM_CONV_STATS) || stats, "%s: KM_CONV_STATS needs stats", who);
  KM_CHECK_ARG(!(flags & KM_CONV_COM), "%s: KM_CONV_COM is not supported", who);
  KM_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0, "%s: pointers must be 16-byte aligned", who);
  return KM_OK;
}

}  // namespace

extern "C" int km_conv3d_zfold(const void* x, const void* wz, void* out, void* pooled, float* stats,
                               int N, int Cin, int Cout, int D, int H, int W, int flags,
                               km_stream_t stream) {
  const int rc = check_zf_args("km_conv3d_zfold", x, wz, out, pooled, stats, N, Cin, Cout, D, H, W, flags);
  if (rc != KM_OK) return rc;
  return launch_zf(x, wz, nullptr, out, pooled, stats, N, 0, N, Cin, D, H, W, flags, false, stream);
}

extern "C" size_t km_conv3d_zfold_gn_workspace_bytes(int N) {
  return (size_t)N * (27 * 3 * kCout * kKC * 2 + kBiasClasses * kCout * 4);
}

extern "C" int km_conv3d_zfold_gn(const void* x, const float* w, const float* scale, const float* shift,
                                  void* out, void* pooled, float* stats, void* workspace, int N, int Cin,
                                  int Cout, int D, int H, int W, int flags, km_stream_t stream) {
  const int rc = check_zf_args("km_conv3d_zfold_gn", x, w, out, pooled, stats, N, Cin, Cout, D, H, W, flags);
  if (rc != KM_OK) return rc;
  KM_CHECK_ARG(scale && shift && workspace && ((uintptr_t)workspace & 15) == 0, "km_conv3d_zfold_gn: null / unaligned argument");
  const size_t wbytes = (size_t)27 * 3 * kCout * kKC * 2;
  void* packed = workspace;
  float* bias = reinterpret_cast<float*>(static_cast<char*>(workspace) + (size_t)N * wbytes);
  const int rf = km_fold_gn(w, scale, shift, packed, bias, N, Cout, Cin, 1, stream);
  if (rf != KM_OK) return rf;
  // one launch, the CTAs split evenly between the samples (each keeps its sample's weights resident);
  // batches larger than half the SM count fall back to one launch per sample
  if (2 * N <= km_sm_count())
    return launch_zf(x, workspace, bias, out, pooled, stats, N, 0, N, Cin, D, H, W, flags, true, stream);
  for (int n = 0; n < N; ++n) {
    const int r2 = launch_zf(x, static_cast<const char*>(workspace) + (size_t)n * wbytes,
                             bias + (size_t)n * kBiasClasses * kCout, out, pooled, stats, N, n, 1, Cin, D, H, W, flags,
                             false, stream);
    if (r2 != KM_OK) return r2;
  }
  return KM_OK;
}
