// The upsampled half of a decoder's first convolution, computed on the COARSE lattice
// (keymorph/unet3d/buildingblocks.py:409-445: x = cat(skip, F.interpolate(x, scale 2, 'nearest')) -> GN -> conv 3x3x3).
//
// A nearest-upsampled input takes the same value on 2x2x2 fine voxels, so for an output voxel of parity p
// (per axis) the three taps -1, 0, +1 read only TWO coarse voxels: (i-1, i) with weights (W-1, W0+W+1) for
// p = 0, (i, i+1) with (W-1+W0, W+1) for p = 1.  Per parity class the 27-tap convolution over the fine lattice
// becomes an 8-tap convolution over the coarse one with pre-summed weights: 8/27 of the MMA work, and the
// upsampled tensor (1 GB for a 128-channel 128^3 pair) is never written or read.
//
// Kernel = conv_zf2.cu's machinery (CTA pairs, cta_group::2 MMA with M = 256, TMA producer, TMEM ring walked
// along z) on that lattice:
//   * a unit is one (x, y) parity class of an 8(x) x 16(y) COARSE brick and a z segment; its 128 rows are the
//     fine voxels (2 i + px, 2 k + py) of the class, so one weight tile serves all rows
//   * z: one coarse input plane j feeds the FOUR fine output planes 2j-1 .. 2j+2 (weights W+1 | W0+W+1 |
//     W-1+W0 | W-1 along z): N = 4 x 64, a ring of four 64-column blocks per TMEM set, two rotations (j parity)
//   * per coarse plane: 2 (dx') x Cin/64 stages, each 2 (dy') x 4 MMAs of 256 x 256 x 16; afterwards the planes
//     2j-1 and 2j are complete and are drained (fp16/bf16 partial sums, no activation) to their strided fine
//     voxels; the skip half's kernel (conv_zf2.cu) adds them in its epilogue before bias / ReLU / statistics.
// The joint GroupNorm of the concat is folded in: the packer scales the summed weights by scale[n][Cs + c]; the
// shift term of ALL channels lives in the skip kernel's border-class bias table (conv_misc.cu km_fold_gn).
#include <algorithm>
#include <type_traits>
#include "km_common.cuh"
#include "tc_ptx.cuh"

using namespace kmtc;

namespace {

constexpr int kThreads = 320;           // warp 0 TMA, warp 1 MMA (leader), warps 2..9 epilogue
constexpr int kEpiThreads = 256;
constexpr int kStagesMax = 4;
constexpr int kCout = 64, kKC = 64, kN = 4 * kCout, kHalfRows = kN / 2;
constexpr int kRowBytes = kKC * 2, kSteps = kKC / 16, kCols = kCout / 2;
constexpr int kBoxRows = 8 * 17;                                   // 8 x-voxels by 16 + 1 y-rows
constexpr uint32_t kASub = kBoxRows * kRowBytes;                   // 17 KB
constexpr uint32_t kBTile = kHalfRows * kRowBytes;                 // one dy' slice: 16 KB
constexpr uint32_t kStage = kASub + 2 * kBTile;                    // one (dx', chunk)
constexpr uint32_t kSetStride = 256u;
constexpr uint32_t kSbo = 8u * kRowBytes;

struct Up2Geom {
  int N, Dc, Hc, Wc, chunks;
  int cblocks, cout_total;                 // output channels in blocks of 64 (one unit computes one block)
  int xpairs, tiles_y, zsegs, lz, units;   // units = N * zsegs * cblocks * 4 classes * tiles_y * xpairs (unit pairs)
  uint32_t off_bars;
  int stages;
};

struct Unit {
  int n, x0, y0, zs, planes, px, py, cb;
  bool valid;
};

__device__ __forceinline__ Unit decode_unit(const Up2Geom& g, int u, uint32_t rank) {
  Unit r;
  const int xp = u % g.xpairs;
  u /= g.xpairs;
  r.x0 = (2 * xp + (int)rank) * 8;
  r.y0 = (u % g.tiles_y) * 16;
  u /= g.tiles_y;
  r.px = u & 1;
  r.py = (u >> 1) & 1;
  u >>= 2;
  r.cb = u % g.cblocks;
  u /= g.cblocks;
  r.zs = (u % g.zsegs) * g.lz;
  r.n = u / g.zsegs;
  r.planes = min(g.lz, g.Dc - r.zs) + 2;   // coarse input planes zs-1 .. zs+lz
  r.valid = r.x0 < g.Wc;
  return r;
}

// two units at a time per CTA pair (TMEM sets 0 / 1), interleaved plane by plane
template <typename F>
__device__ __forceinline__ void for_each_tile(const Up2Geom& g, uint32_t rank, F&& fn) {
  uint32_t cnt[2] = {0u, 0u};
  const int pair = (int)blockIdx.x >> 1, npairs = (int)gridDim.x >> 1;
  for (int ua = pair; ua < g.units; ua += 2 * npairs) {
    const int ub = ua + npairs;
    const Unit a = decode_unit(g, ua, rank);
    Unit b = a;
    b.planes = 0;
    if (ub < g.units) b = decode_unit(g, ub, rank);
    const int pmax = max(a.planes, b.planes);
    for (int p = 0; p < pmax; ++p) {
      if (p < a.planes) fn(std::integral_constant<uint32_t, 0u>{}, a, p, cnt[0]++);
      if (p < b.planes) fn(std::integral_constant<uint32_t, 1u>{}, b, p, cnt[1]++);
    }
  }
}

template <bool F16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_up2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Up2Geom g,
                uint16_t* __restrict__ out) {
  constexpr uint32_t kLayout = 2u;   // SWIZZLE_128B
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
    prefetch_tmap(&tmB);
  }
  __syncwarp();
  if (warp == 0) {
    tmem_alloc2(smem_u32(tmem_ptr_smem), 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int subs = 2 * g.chunks;   // (dx', chunk) sub-iterations per coarse plane, one pipeline stage each

  if (warp == 0) {
    // =============================== TMA producer (both CTAs) ====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for_each_tile(g, rank, [&](auto, const Unit& u, int p, uint32_t) {
        const int j = u.zs - 1 + p;   // coarse input plane; outside [0, Dc) -> TMA zero fill
        const int wsel = ((((u.n * g.cblocks + u.cb) * 4 + u.py * 2 + u.px)) << 1) | (j & 1);
        for (int dx = 0; dx < 2; ++dx) {
          for (int ch = 0; ch < g.chunks; ++ch) {
            mbar_wait(empty_bar(s), ph ^ 1u);
            const uint32_t lead_full = mapa_u32(full_bar(s), 0);
            mbar_arrive_expect_tx_cluster(lead_full, kStage);
            const uint32_t dst = base + (uint32_t)s * kStage;
            tma2_load_5d(dst, &tmA, lead_full, ch * kKC, u.x0 + u.px - 1 + dx, u.y0 + u.py - 1, j, u.n);
            // weights (Cin, 256 rows, dy', dx', (n, class, rot)): this CTA's 128 rows of both dy' slices
            tma2_load_5d(dst + kASub, &tmB, lead_full, ch * kKC, (int)rank * kHalfRows, 0, dx, wsel);
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
      const uint32_t idesc = umma_idesc_16(256, kN, F16);
      int s = 0;
      uint32_t ph = 0;
      for_each_tile(g, rank, [&](auto set_c, const Unit&, int p, uint32_t cnt) {
        constexpr uint32_t set = decltype(set_c)::value;
        mbar_wait(tempty_bar(set), (cnt & 1u) ^ 1u);   // both CTAs drained + zeroed the completed planes
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + set * kSetStride;
        uint32_t accum = p == 0 ? 0u : 1u;   // a unit's first plane overwrites the whole ring
        for (int si = 0; si < subs; ++si) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t a16 = base16 + (uint32_t)s * (kStage >> 4);
          const uint32_t b16 = a16 + (kASub >> 4);
#pragma unroll
          for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
            for (int kk = 0; kk < kSteps; ++kk) {
              umma2_16_pred(d_tmem, a16 + (uint32_t)dy * (kSbo >> 4) + 2u * kk,
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
    // A thread owns one row of the brick (a coarse voxel of the class = one fine voxel) and 32 of the 64
    // output channels of the two planes that complete at this step.
    const int q = warp & 3;
    const int row = q * 32 + lane;        // tx = row & 7, ty = row >> 3
    const int half = (warp - 2) >> 2;
    const int tx = row & 7, ty = row >> 3;
    const int D = 2 * g.Dc, H = 2 * g.Hc, W = 2 * g.Wc;
    for_each_tile(g, rank, [&](auto set_c, const Unit& u, int p, uint32_t cnt) {
      constexpr uint32_t set = decltype(set_c)::value;
      const int j = u.zs - 1 + p;
      const int zlo = 2 * u.zs, zhi = min(2 * (u.zs + u.planes - 2), D);   // this unit's fine output planes
      const uint32_t lead_tempty = mapa_u32(tempty_bar(set), 0);
      mbar_wait(tfull_bar(set), cnt & 1u);
      tc_fence_after();
      uint32_t r[2][kCols / 16][16];
      bool store[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int zo = 2 * j - 1 + k;
        store[k] = zo >= zlo && zo < zhi;   // warp-uniform: tcgen05.ld is a warp-collective operation
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + set * kSetStride + (uint32_t)(zo & 3) * kCout +
                               (uint32_t)(half * kCols);
        if (store[k]) {
#pragma unroll
          for (int b = 0; b < kCols / 16; ++b) tmem_ld16(taddr + 16u * b, r[k][b]);
          tmem_ld_wait();
        }
#pragma unroll
        for (int b = 0; b < kCols / 16; ++b) tmem_st16_zero(taddr + 16u * b);   // becomes plane zo + 4
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive_cluster(lead_tempty);
      const int xc = u.x0 + tx, yc = u.y0 + ty;
      const bool inside = u.valid && xc < g.Wc && yc < g.Hc;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (!store[k] || !inside) continue;
        const int zo = 2 * j - 1 + k;
        const size_t vox = (((size_t)u.n * D + zo) * H + (2 * yc + u.py)) * W + (2 * xc + u.px);
#pragma unroll
        for (int hh = 0; hh < kCols / 16; ++hh) {
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            pk[i] = km_pack2<F16>(__uint_as_float(r[k][hh][2 * i]), __uint_as_float(r[k][hh][2 * i + 1]));
          st_global_v8(out + vox * g.cout_total + u.cb * kCout + half * kCols + 16 * hh, pk);
        }
      }
    });
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// fp32 (Cout, Cs + Cu, 3, 3, 3) and scale (N, Cs + Cu) -> 16-bit [n][cout block][class py px][rot][dx'][dy'][slot * 64 + cout][Cu]:
// the taps a coarse neighbour stands for, summed, times the GroupNorm scale of the channel.
//   x: fine taps dx in {px - 1 + 2 dx', px + 2 dx'} & [0, 2]  (same for y);
//   z: the slot holds fine plane 2j + d, d = ((slot - 2 rot + 1) & 3) - 1, fine taps dz in {1 - d, 2 - d} & [0, 2]
template <bool F16>
__global__ void __launch_bounds__(256)
pack_up2_kernel(const float* __restrict__ w, const float* __restrict__ scale, uint16_t* __restrict__ packed, int N,
                int Cs, int Cu, int cblocks) {
  // one thread per (cout, upsampled cin): its 27 taps are contiguous; the 2 x 2 x 2 tap sets per axis are the four
  // selectors {0}, {1,2}, {0,1}, {2} (index p * 2 + tap'), reduced separably into 64 sums
  const int Cin = Cs + Cu;
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= cblocks * kCout * Cu) return;
  const int co = t / Cu, ci = t % Cu;
  const float* src = w + ((size_t)co * Cin + Cs + ci) * 27;
  float wk[27];
#pragma unroll
  for (int k = 0; k < 27; ++k) wk[k] = src[k];
  float sx[3][3][4], sy[3][4][4], sz[4][4][4];
#pragma unroll
  for (int d = 0; d < 9; ++d) {
    const float w0 = wk[d * 3], w1 = wk[d * 3 + 1], w2 = wk[d * 3 + 2];
    sx[d / 3][d % 3][0] = w0;
    sx[d / 3][d % 3][1] = w1 + w2;
    sx[d / 3][d % 3][2] = w0 + w1;
    sx[d / 3][d % 3][3] = w2;
  }
#pragma unroll
  for (int dz = 0; dz < 3; ++dz)
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const float a0 = sx[dz][0][x], a1 = sx[dz][1][x], a2 = sx[dz][2][x];
      sy[dz][0][x] = a0;
      sy[dz][1][x] = a1 + a2;
      sy[dz][2][x] = a0 + a1;
      sy[dz][3][x] = a2;
    }
#pragma unroll
  for (int y = 0; y < 4; ++y)
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const float a0 = sy[0][y][x], a1 = sy[1][y][x], a2 = sy[2][y][x];
      sz[0][y][x] = a0;
      sz[1][y][x] = a1 + a2;
      sz[2][y][x] = a0 + a1;
      sz[3][y][x] = a2;
    }
  const int cb = co / kCout, co64 = co % kCout;
  const size_t per = (size_t)32 * kN * Cu * cblocks;   // [cout block][class][rot][dx'][dy'][slot][64][Cu]
  for (int n = 0; n < N; ++n) {
    const float sc = scale ? scale[(size_t)n * Cin + Cs + ci] : 1.f;
    uint16_t* dst = packed + (size_t)n * per + (size_t)cb * 32 * kN * Cu + (size_t)co64 * Cu + ci;
#pragma unroll
    for (int cls = 0; cls < 4; ++cls)
#pragma unroll
      for (int rot = 0; rot < 2; ++rot)
#pragma unroll
        for (int dxp = 0; dxp < 2; ++dxp)
#pragma unroll
          for (int dyp = 0; dyp < 2; ++dyp)
#pragma unroll
            for (int slot = 0; slot < 4; ++slot) {
              const int d = ((slot - 2 * rot + 1) & 3) - 1;              // fine plane 2j + d lives in this slot
              const int zsel = d == -1 ? 3 : (d == 0 ? 1 : (d == 1 ? 2 : 0));
              const float v = sz[zsel][(cls >> 1) * 2 + dyp][(cls & 1) * 2 + dxp] * sc;
              dst[(size_t)((((cls * 2 + rot) * 2 + dxp) * 2 + dyp) * 4 + slot) * kCout * Cu] = km_from_float<F16>(v);
            }
  }
}

}  // namespace

extern "C" int km_sm_count(void);

extern "C" int km_conv3d_up2_supported(int Cu, int Cout, int Dc, int Hc, int Wc) {
  return (Cout % kCout == 0 && Cout >= kCout && Cout <= 256 && Cu % kKC == 0 && Cu >= kKC && Cu <= 512 && Wc >= 8 &&
          Hc >= 16 && Dc >= 1) ? 1 : 0;
}

extern "C" size_t km_conv3d_up2_gn_workspace_bytes(int N, int Cu, int Cout) {
  return (size_t)N * 32 * kN * Cu * 2 * (size_t)((Cout + kCout - 1) / kCout);
}

extern "C" int km_conv3d_up2_gn(const void* xc, const float* w, const float* scale, int Cs, int Cu, void* out,
                                void* workspace, int N, int Cout, int Dc, int Hc, int Wc, km_stream_t stream) {
  KM_CHECK_ARG(xc && w && out && workspace, "km_conv3d_up2_gn: null argument");
  KM_CHECK_ARG(km_conv3d_up2_supported(Cu, Cout, Dc, Hc, Wc) && Cs >= 0,
               "km_conv3d_up2_gn: unsupported shape (Cu=%d Cout=%d coarse %dx%dx%d)", Cu, Cout, Dc, Hc, Wc);
  KM_CHECK_ARG(N > 0 && N <= 1024, "km_conv3d_up2_gn: bad batch");
  KM_CHECK_ARG(((uintptr_t)xc & 15) == 0 && ((uintptr_t)out & 31) == 0 && ((uintptr_t)workspace & 255) == 0,
               "km_conv3d_up2_gn: pointers must be 16-byte (output 32, workspace 256) aligned");
  cudaStream_t st = km_cs(stream);
  const int cblocks = Cout / kCout;
  if (km_operand_fp16())
    pack_up2_kernel<true><<<(Cout * Cu + 255) / 256, 256, 0, st>>>(w, scale, reinterpret_cast<uint16_t*>(workspace), N, Cs, Cu, cblocks);
  else
    pack_up2_kernel<false><<<(Cout * Cu + 255) / 256, 256, 0, st>>>(w, scale, reinterpret_cast<uint16_t*>(workspace), N, Cs, Cu, cblocks);
  KM_LAUNCH_OK("pack_up2_kernel");

  Up2Geom g;
  memset(&g, 0, sizeof(g));
  g.N = N; g.Dc = Dc; g.Hc = Hc; g.Wc = Wc;
  g.chunks = Cu / kKC;
  g.cblocks = cblocks;
  g.cout_total = Cout;
  const int tiles_x = (Wc + 7) / 8;
  g.xpairs = (tiles_x + 1) / 2;
  g.tiles_y = (Hc + 15) / 16;
  // z segment length: longer segments amortise the two halo planes, shorter ones balance the units over the
  // CTA pairs (all samples and classes share one unit list)
  const int npairs_hw = km_sm_count() / 2;
  double best = 1e30;
  for (int lz = 64; lz >= 8; lz /= 2) {
    const long long zs = (Dc + lz - 1) / lz;
    const long long units = (long long)N * zs * cblocks * 4 * g.tiles_y * g.xpairs;
    const long long rounds = (units + npairs_hw - 1) / npairs_hw;
    const double cost = (double)rounds * (double)(std::min(lz, Dc) + 2);   // planes walked by the busiest pair
    if (cost < best) {
      best = cost;
      g.lz = lz;
    }
  }
  g.zsegs = (Dc + g.lz - 1) / g.lz;
  const long long units = (long long)N * g.zsegs * cblocks * 4 * g.tiles_y * g.xpairs;
  KM_CHECK_ARG(units < (1ll << 30), "km_conv3d_up2_gn: too many units");
  g.units = (int)units;
  g.stages = kStagesMax;
  g.off_bars = (uint32_t)g.stages * kStage;
  const uint32_t smem_bytes = g.off_bars + 8u * (2u * kStagesMax + 6u) + 16u + 1024u;
  KM_CHECK_ARG(smem_bytes <= 232448, "km_conv3d_up2_gn: shared memory overflow (%u)", smem_bytes);

  PFN_encodeTiled encode = tensor_map_encoder();
  if (!encode) {
    km_set_error("km_conv3d_up2_gn: cuTensorMapEncodeTiled unavailable");
    return KM_ECUDA;
  }
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cu, (cuuint64_t)Wc, (cuuint64_t)Hc, (cuuint64_t)Dc, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)Cu * 2, (cuuint64_t)Wc * Cu * 2, (cuuint64_t)Hc * Wc * Cu * 2,
                             (cuuint64_t)Dc * Hc * Wc * Cu * 2};
    cuuint32_t box[5] = {(cuuint32_t)kKC, 8, 17, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tmA, KM_TMAP_16, 5, const_cast<void*>(xc), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      km_set_error("km_conv3d_up2_gn: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
      return KM_ECUDA;
    }
  }
  {
    // packed weights [n, cout block, class, rot][dx'][dy'][256 rows][Cu] viewed as (Cu, rows, dy', dx', the rest)
    const cuuint64_t tile = (cuuint64_t)kN * Cu * 2;
    cuuint64_t dims[5] = {(cuuint64_t)Cu, (cuuint64_t)kN, 2, 2, (cuuint64_t)(8 * N * cblocks)};
    cuuint64_t strides[4] = {(cuuint64_t)Cu * 2, tile, 2 * tile, 4 * tile};
    cuuint32_t box[5] = {(cuuint32_t)kKC, (cuuint32_t)kHalfRows, 2, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tmB, KM_TMAP_16, 5, workspace, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      km_set_error("km_conv3d_up2_gn: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
      return KM_ECUDA;
    }
  }
  static unsigned long long attr_set = 0;
  if (km_first_use_on_device(&attr_set)) {
    KM_CUDA_OK(cudaFuncSetAttribute(conv_up2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    KM_CUDA_OK(cudaFuncSetAttribute(conv_up2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  }
  const int nsm = km_sm_count();
  int grid = nsm & ~1;
  if (grid / 2 > g.units) grid = 2 * g.units;
  if (km_operand_fp16())
    conv_up2_kernel<true><<<grid, kThreads, smem_bytes, st>>>(tmA, tmB, g, reinterpret_cast<uint16_t*>(out));
  else
    conv_up2_kernel<false><<<grid, kThreads, smem_bytes, st>>>(tmA, tmB, g, reinterpret_cast<uint16_t*>(out));
  KM_LAUNCH_OK("conv_up2_kernel");
  return KM_OK;
}
