// 2-CTA ("CTA pair", cta_group::2) variant of the tcgen05 implicit-GEMM 3x3x3 convolution of
// conv_tc.cu (its MODE 2 data path: 8(x) x 16(y) bricks, y taps through UMMA descriptor offsets, up
// to 4 bricks per weight fetch, streaming weights, statistics epilogue).
//
// Reference call site: keymorph/unet3d/buildingblocks.py:50-52 (Conv3d k3 p1, no bias, + ReLU).
//
// Why: the Cout = 64 / 128 layers of the backbone are bound by shared-memory operand bandwidth (ncu:
// smem->tensor wavefronts 93 %): every K = 16 MMA re-reads (128 + BN) x 32 B.  With cta_group::2 the
// two SMs of a TPC execute ONE M = 256 MMA: each CTA supplies its own 128 activation rows (its own
// brick group) but only HALF of the weight rows (BN/2), so the weight bytes read from shared memory
// -- and fetched by TMA -- per SM halve: 6 -> 5 KB per MMA at BN = 64, 8 -> 6 KB at BN = 128.
//
// Protocol (all barriers live at identical shared-memory offsets in both CTAs):
//   full[s]    leader only: 2 arrivals (each CTA's producer announces its own bytes, the peer's
//              arrives remotely) + the transaction bytes of both CTAs' TMA loads (cta_group::2 loads
//              complete on the leader's barrier);
//   empty[s]   per CTA: released by the leader's tcgen05.commit, multicast to both CTAs;
//   tfull[a]   per CTA: accumulator ready, multicast commit;
//   tempty[a]  leader only: 512 arrivals (both CTAs' epilogue threads; the peer's arrive remotely).
// Only the leader (cluster rank 0) issues MMAs (two issuer warps); both CTAs run a TMA producer and an
// 8-warp epilogue on their own half of the accumulator.
#include "km_common.cuh"
#include "tc_ptx.cuh"

using namespace kmtc;

namespace {

constexpr int kTileM = 128;
constexpr int kThreads = 352;     // warp 0 TMA, warps 1 + 10 MMA (leader), warps 2..9 epilogue
constexpr int kEpiThreads = 256;
constexpr int kMaxStages = 12;

struct PairGeom {
  int N, D, H, W, Cin, Cout;
  int chunks;          // Cin / KC
  int tiles_x, tiles_y;
  int BN;              // = Cout (one output-channel block)
  int stages, sub, subiters;
  int mt;              // y-adjacent bricks per CTA that share one weight fetch
  int eb;              // bricks staged together per epilogue round
  int issuers;
  int flags;
  int stat_parts;
  uint32_t a_sub_bytes, b_sub_bytes, a_sub_stride, b_sub_stride, stage_stride;
  uint32_t off_staging, off_rowvalid, off_stats, off_bars;
  uint32_t staging_pitch;
  uint32_t idesc;
  int total_tiles;
  int w_per_sample;    // GN-folded path: one packed weight set per sample (tmB's last dimension is 3 N)
};

constexpr int kBiasClasses = 36;   // z code (0..3) x y code (0..2) x x code (0..2), see conv_zf.cu

struct TileCoord {
  int n, x0, y0, z0;
};
__device__ __forceinline__ TileCoord decode_tile(const PairGeom& g, int t) {
  TileCoord c;
  c.x0 = (t % g.tiles_x) * 8;
  t /= g.tiles_x;
  c.y0 = (t % g.tiles_y) * (16 * g.mt);
  t /= g.tiles_y;
  c.z0 = t % g.D;
  c.n = t / g.D;
  return c;
}

template <int KC, bool F16, bool ADD = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const PairGeom g, uint16_t* __restrict__ out, float* __restrict__ stats,
                const float* __restrict__ bias_tab, const uint16_t* __restrict__ addend) {
  constexpr int kRowBytes = KC * 2;
  constexpr int kSteps = KC / 16;
  constexpr uint32_t kLayout = kRowBytes == 128 ? 2u : (kRowBytes == 64 ? 4u : 6u);
  constexpr uint32_t kSbo = 8u * kRowBytes;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_u32);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = (int)blockIdx.x >> 1, npairs = (int)gridDim.x >> 1;
  const int stages = g.stages;
  const uint32_t stage_stride = g.stage_stride;

  const uint32_t bars = base + g.off_bars;  // full[stages], empty[stages], tfull[2], tempty[2]
  auto full_bar = [&](int s) { return bars + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bars + 8u * (uint32_t)(stages + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (uint32_t)(2 * stages + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (uint32_t)(2 * stages + 2 + a); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sm + g.off_bars + 8u * (2 * stages + 5));

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 2);                         // both CTAs' producers
      mbar_init(empty_bar(s), (uint32_t)g.issuers);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), (uint32_t)g.issuers);
      mbar_init(tempty_bar(a), 2 * kEpiThreads);         // both CTAs' epilogues
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
  {
    float* s_stats = reinterpret_cast<float*>(sm + g.off_stats);
    if (g.flags & KM_CONV_STATS)
      for (int i = threadIdx.x; i < g.stat_parts * g.N * g.Cout * 2; i += kThreads) s_stats[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers are initialised before anyone arrives on them remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int sub = g.sub;
  const int n_stage_iters = (g.subiters + sub - 1) / sub;
  const int last_nsub = g.subiters - (n_stage_iters - 1) * sub;
  const int hbn = g.BN / 2;   // weight rows held by each CTA

  if (warp == 0) {
    // =============================== TMA producer (both CTAs) ====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t sub_tx = g.a_sub_bytes + g.b_sub_bytes;
      for (int pt = pair; 2 * pt < g.total_tiles; pt += npairs) {
        const int tile = min(2 * pt + (int)rank, g.total_tiles - 1);   // an odd last pair: duplicate work
        const TileCoord tc = decode_tile(g, tile);
        int ch = 0, dz = -1, dx = -1;   // (dz, dx) group: the three dy taps share one activation box
        for (int si = 0; si < n_stage_iters; ++si) {
          const int nsub = (si == n_stage_iters - 1) ? last_nsub : sub;
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t lead_full = mapa_u32(full_bar(s), 0);
          mbar_arrive_expect_tx_cluster(lead_full, (uint32_t)nsub * sub_tx);
          uint32_t a_dst = base + (uint32_t)s * stage_stride;
          uint32_t b_dst = a_dst + (uint32_t)sub * g.a_sub_stride;
          for (int u = 0; u < nsub; ++u) {
            // activations: dims (C, W, H, D, N), box 8 x-voxels by 16*mt+2 y-rows
            tma2_load_5d(a_dst, &tmA, lead_full, ch * KC, tc.x0 + dx, tc.y0 - 1, tc.z0 + dz, tc.n);
            // weights: (Cin, Cout, dx, dy, dz) map, this CTA's half of the rows, all three dy slices
            tma2_load_5d(b_dst, &tmB, lead_full, ch * KC, (int)rank * hbn, dx + 1, 0,
                         dz + 1 + (g.w_per_sample ? 3 * tc.n : 0));
            a_dst += g.a_sub_stride;
            b_dst += g.b_sub_stride;
            if (++ch == g.chunks) {
              ch = 0;
              if (++dx == 2) {
                dx = -1;
                ++dz;
              }
            }
          }
          if (++s == stages) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1 || warp == 10) {
    // =============================== MMA issuers (leader CTA only) ==============
    if (rank == 0 && !(warp == 10 && g.issuers < 2)) {
      const uint32_t issue = elect_one();
      int s = 0;
      uint32_t ph = 0;
      uint32_t tcount = 0;
      constexpr uint32_t desc_hi = (kSbo >> 4) | (1u << 14) | (kLayout << 29);
      const uint32_t lo_flag = 1u << 16;
      const uint32_t a_tap16 = (8u * kRowBytes) >> 4;                 // one swizzle atom = one dy step
      const uint32_t b_tap16 = ((uint32_t)hbn * kRowBytes) >> 4;      // next dy slice of this CTA's rows
      const uint32_t a_sub16 = g.a_sub_stride >> 4, b_sub16 = g.b_sub_stride >> 4;
      const uint32_t base16 = ((base & 0x3FFFFu) >> 4) | lo_flag;
      const uint32_t stage16 = stage_stride >> 4;
      const uint32_t boff16 = ((uint32_t)sub * g.a_sub_stride) >> 4;
      const uint32_t idesc = g.idesc;
      const uint32_t bn = (uint32_t)g.BN;
      const int m_lo = (g.issuers == 2 && warp == 10) ? g.mt / 2 : 0;
      const int mt = g.issuers == 2 ? (warp == 10 ? g.mt - g.mt / 2 : g.mt / 2) : g.mt;
      for (int pt = pair; 2 * pt < g.total_tiles; pt += npairs, ++tcount) {
        const uint32_t acc = tcount & 1u;
        const uint32_t acc_ph = (tcount >> 1) & 1u;
        mbar_wait(tempty_bar((int)acc), acc_ph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (acc * (uint32_t)g.mt + (uint32_t)m_lo) * bn;
        uint32_t accum = 0;
        for (int si = 0; si < n_stage_iters; ++si) {
          const int nsub = (si == n_stage_iters - 1) ? last_nsub : sub;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          uint32_t a16 = base16 + (uint32_t)s * stage16;
          uint32_t b16 = a16 + boff16;
          a16 += (uint32_t)m_lo * 16u * a_tap16;
          for (int u = 0; u < nsub; ++u) {
            uint32_t am16 = a16, dm = d_tmem;
            for (int m = 0; m < mt; ++m) {   // bricks sharing this weight slice
#pragma unroll
              for (int t = 0; t < 3; ++t) {
#pragma unroll
                for (int kk = 0; kk < kSteps; ++kk) {
                  umma2_16_pred(dm, am16 + t * a_tap16 + 2u * kk, b16 + t * b_tap16 + 2u * kk, desc_hi, idesc,
                                  (t | kk) ? 1u : accum, issue);
                }
              }
              am16 += 16u * a_tap16;
              dm += bn;
            }
            accum = 1u;
            a16 += a_sub16;
            b16 += b_sub16;
          }
          umma2_commit_pred(empty_bar(s), issue);   // frees the stage in BOTH CTAs
          if (++s == stages) {
            s = 0;
            ph ^= 1u;
          }
        }
        umma2_commit_pred(tfull_bar((int)acc), issue);   // accumulators complete in both CTAs
      }
    }
  } else {
    // =============================== epilogue (8 warps, both CTAs) ===============
    const int q = warp & 3;              // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;       // accumulator row == voxel within the brick
    const int half = (warp - 2) >> 2;    // column group
    const int et = half * kTileM + row;
    uint8_t* staging = sm + g.off_staging;
    uint8_t* rowvalid = sm + g.off_rowvalid;                               // [eb][128]
    float* s_stats = reinterpret_cast<float*>(sm + g.off_stats);           // [parts][N][Cout][2]
    const bool do_relu = (g.flags & KM_CONV_RELU) != 0;
    const bool do_stats = (g.flags & KM_CONV_STATS) != 0;
    const int BN = g.BN;
    const uint32_t pitch = g.staging_pitch;
    const int cpr = BN / 8;
    const int parts = g.stat_parts;
    const int rows_per_part = kTileM / parts;
    auto all_bar = [&]() { asm volatile("bar.sync 3, 256;" ::: "memory"); };
    const int tx = row & 7, ty = row >> 3;

    uint32_t tcount = 0;
    for (int pt = pair; 2 * pt < g.total_tiles; pt += npairs, ++tcount) {
      const bool tile_valid = 2 * pt + (int)rank < g.total_tiles;
      const TileCoord tc0 = decode_tile(g, min(2 * pt + (int)rank, g.total_tiles - 1));
      const uint32_t acc = tcount & 1u;
      const uint32_t acc_ph = (tcount >> 1) & 1u;
      const uint32_t lead_tempty = mapa_u32(tempty_bar((int)acc), 0);
      mbar_wait(tfull_bar((int)acc), acc_ph);
      tc_fence_after();
      if (!tile_valid) {   // duplicate of the pair's other tile: nothing to store
        tc_fence_before();
        mbar_arrive_cluster(lead_tempty);
        continue;
      }
      int m_last = 0;
      while (m_last + 1 < g.mt && tc0.y0 + 16 * (m_last + 1) < g.H) ++m_last;
      const int eb = g.eb;
      for (int m = 0; m <= m_last; m += eb) {
        const int nb = min(eb, m_last + 1 - m);
        const int y0 = tc0.y0 + 16 * m;
        const bool last_sub = m + nb - 1 == m_last;
        uint32_t vmask = 0;
        for (int mb = 0; mb < nb; ++mb) {
          const bool v = (tc0.x0 + tx < g.W) && (y0 + 16 * mb + ty < g.H);
          vmask |= (v ? 1u : 0u) << mb;
          rowvalid[mb * kTileM + row] = v ? 1 : 0;
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) +
                               (acc * (uint32_t)g.mt + (uint32_t)m) * (uint32_t)BN;
        {
          const int nblk = BN / 16, split = (nblk + 1) / 2;
          const int b_lo = half == 0 ? 0 : split, b_hi = half == 0 ? split : nblk;
          for (int mb = 0; mb < nb; ++mb) {
            const bool vrow = ((vmask >> mb) & 1u) != 0;
            const uint32_t taddr_b = taddr + (uint32_t)(mb * BN);
            uint8_t* srow = staging + (size_t)(mb * kTileM + row) * pitch;
            const float4* brow = nullptr;
            if (bias_tab) {   // GroupNorm shift of the folded norm: bias[sample][border class][cout] (L1-resident)
              const int x2 = tc0.x0 + tx, y2 = y0 + 16 * mb + ty, z2 = tc0.z0;
              const int cls = ((z2 == 0 ? 1 : 0) | (z2 == g.D - 1 ? 2 : 0)) * 9 +
                              (y2 == 0 ? 1 : (y2 == g.H - 1 ? 2 : 0)) * 3 + (x2 == 0 ? 1 : (x2 == g.W - 1 ? 2 : 0));
              brow = reinterpret_cast<const float4*>(bias_tab + ((size_t)tc0.n * kBiasClasses + cls) * g.Cout);
            }
            // ADD: partial sums of the same layer from conv_up2.cu (the upsampled half of a decoder's concat)
            const uint4* arow = nullptr;
            if (ADD && vrow) {
              const int x2 = tc0.x0 + tx, y2 = y0 + 16 * mb + ty;
              arow = reinterpret_cast<const uint4*>(
                  addend + ((((size_t)tc0.n * g.D + tc0.z0) * g.H + y2) * g.W + x2) * g.Cout);
            }
            for (int blk = b_lo; blk < b_hi; ++blk) {
              const int c0 = blk * 16;
              uint32_t r[16];
              tmem_ld16(taddr_b + (uint32_t)c0, r);
              float4 bv[4];
              if (brow) {
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) bv[j4] = __ldg(brow + blk * 4 + j4);
              }
              uint4 a0 = make_uint4(0u, 0u, 0u, 0u), a1 = a0;
              if (ADD && arow) {
                a0 = __ldg(arow + 2 * blk);
                a1 = __ldg(arow + 2 * blk + 1);
              }
              tmem_ld_wait();
              if (ADD && arow) {
                const uint32_t a8[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float2 ab = km_unpack2<F16>(a8[j]);
                  r[2 * j] = __float_as_uint(__uint_as_float(r[2 * j]) + ab.x);
                  r[2 * j + 1] = __float_as_uint(__uint_as_float(r[2 * j + 1]) + ab.y);
                }
              }
              if (brow) {
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                  r[4 * j4 + 0] = __float_as_uint(__uint_as_float(r[4 * j4 + 0]) + bv[j4].x);
                  r[4 * j4 + 1] = __float_as_uint(__uint_as_float(r[4 * j4 + 1]) + bv[j4].y);
                  r[4 * j4 + 2] = __float_as_uint(__uint_as_float(r[4 * j4 + 2]) + bv[j4].z);
                  r[4 * j4 + 3] = __float_as_uint(__uint_as_float(r[4 * j4 + 3]) + bv[j4].w);
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
                pk[j] = vrow ? km_pack2<F16>(a, b) : 0u;   // rows outside the volume: zeros (stats need no mask)
              }
              uint4* dst = reinterpret_cast<uint4*>(srow + (size_t)c0 * 2);
              dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
          if (last_sub) {
            tc_fence_before();
            mbar_arrive_cluster(lead_tempty);
          }
          all_bar();
        }
        // ---- staged bf16 tile -> global (coalesced 16-byte chunks) + per-channel stats ----
        const int total_chunks = nb * kTileM * cpr;
        for (int id = et; id < total_chunks; id += kEpiThreads) {
          const int j = id % cpr;
          const int rr = id / cpr;
          const int mb = rr >> 7;
          const int r2 = rr & (kTileM - 1);
          if (!rowvalid[mb * kTileM + r2]) continue;
          const int x2 = tc0.x0 + (r2 & 7), y2 = y0 + 16 * mb + (r2 >> 3);
          const size_t vox = (((size_t)tc0.n * g.D + tc0.z0) * g.H + y2) * g.W + x2;
          *reinterpret_cast<uint4*>(out + vox * g.Cout + j * 8) =
              *reinterpret_cast<const uint4*>(staging + (size_t)(mb * kTileM + r2) * pitch + j * 16);
        }
        if (do_stats) {
          for (int id = et; id < parts * BN; id += kEpiThreads) {
            const int col = id % BN, part = id / BN;
            float s = 0.f, ss = 0.f;
            for (int mb = 0; mb < nb; ++mb) {
              const uint8_t* p = staging + (size_t)(mb * kTileM + part * rows_per_part) * pitch + (size_t)col * 2;
#pragma unroll 8
              for (int rr = 0; rr < rows_per_part; ++rr) {
                const float v = km_to_float<F16>(*reinterpret_cast<const uint16_t*>(p));
                p += pitch;
                s += v;
                ss = fmaf(v, v, ss);
              }
            }
            float* d = s_stats + (((size_t)part * g.N + tc0.n) * g.Cout + col) * 2;
            d[0] += s;
            d[1] += ss;
          }
        }
        all_bar();  // staging / rowvalid may be overwritten by the next round
      }
    }

    all_bar();
    if (do_stats) {
      float* dst = stats + (size_t)blockIdx.x * g.N * g.Cout * 2;
      const int n = g.N * g.Cout * 2;
      for (int i = et; i < n; i += kEpiThreads) {
        float a = 0.f;
        for (int p = 0; p < parts; ++p) a += s_stats[(size_t)p * n + i];
        dst[i] = a;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no remote arrive / multicast commit may target a CTA that has exited
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

inline uint32_t pair_round_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

typedef void (*PairKernel)(const CUtensorMap, const CUtensorMap, const PairGeom, uint16_t*, float*,
                           const float*, const uint16_t*);


}  // namespace

extern "C" int km_sm_count(void);

extern "C" int km_conv3d_tc_pair_supported(int Cin, int Cout, int D, int H, int W) {
  const bool cin_ok = Cin >= 32 && Cin % 32 == 0;
  return (cin_ok && (Cout == 64 || Cout == 128) && W >= 8 && H >= 32 && D >= 1) ? 1 : 0;
}

namespace {
int launch_pair(const void* x, const void* wp, const float* bias_tab, const void* addend, void* out, float* stats,
                int N, int Cin, int Cout, int D, int H, int W, int flags, km_stream_t stream);
}

extern "C" int km_conv3d_tc_pair(const void* x, const void* wp, void* out, float* stats, int N, int Cin,
                                 int Cout, int D, int H, int W, int flags, km_stream_t stream) {
  return launch_pair(x, wp, nullptr, nullptr, out, stats, N, Cin, Cout, D, H, W, flags, stream);
}

extern "C" size_t km_conv3d_tc_pair_gn_workspace_bytes(int N, int Cin, int Cout) {
  return (((size_t)N * 27 * Cout * Cin * 2 + 255) & ~(size_t)255) + (size_t)N * kBiasClasses * Cout * 4;
}

extern "C" int km_conv3d_tc_pair_gn(const void* x, const float* w, const float* scale, const float* shift,
                                    void* out, float* stats, void* workspace, int N, int Cin, int Cout, int D,
                                    int H, int W, int flags, km_stream_t stream) {
  KM_CHECK_ARG(w && scale && shift && workspace && ((uintptr_t)workspace & 255) == 0,
               "km_conv3d_tc_pair_gn: null / unaligned (256 B) argument");
  KM_CHECK_ARG(km_conv3d_tc_pair_supported(Cin, Cout, D, H, W),
               "km_conv3d_tc_pair_gn: unsupported shape (Cin=%d Cout=%d H=%d W=%d)", Cin, Cout, H, W);
  KM_CHECK_ARG(N > 0 && N <= 1024, "km_conv3d_tc_pair_gn: bad batch");
  const size_t wbytes = (size_t)27 * Cout * Cin * 2;
  void* packed = workspace;
  float* bias = reinterpret_cast<float*>(static_cast<char*>(workspace) + (((size_t)N * wbytes + 255) & ~(size_t)255));
  const int rf = km_fold_gn(w, scale, shift, packed, bias, N, Cout, Cin, 0, stream);
  if (rf != KM_OK) return rf;
  return launch_pair(x, workspace, bias, nullptr, out, stats, N, Cin, Cout, D, H, W, flags, stream);
}

// The first Cs channels of a (Cs + Cu)-channel folded layer; `addend` holds the other channels' partial sums
// (km_conv3d_up2_gn).  The border-class bias table covers ALL channels, the packed weights the first Cs.
extern "C" int km_conv3d_tc_pair_gn_add(const void* x, int Cs, int Cu, const float* w, const float* scale,
                                        const float* shift, const void* addend, void* out, float* stats,
                                        void* workspace, int N, int Cout, int D, int H, int W, int flags,
                                        km_stream_t stream) {
  KM_CHECK_ARG(w && scale && shift && addend && workspace && ((uintptr_t)workspace & 255) == 0 &&
                   ((uintptr_t)addend & 15) == 0,
               "km_conv3d_tc_pair_gn_add: null / unaligned argument");
  KM_CHECK_ARG(Cs > 0 && Cu > 0 && km_conv3d_tc_pair_supported(Cs, Cout, D, H, W),
               "km_conv3d_tc_pair_gn_add: unsupported shape (Cs=%d Cu=%d Cout=%d H=%d W=%d)", Cs, Cu, Cout, H, W);
  KM_CHECK_ARG(N > 0 && N <= 1024, "km_conv3d_tc_pair_gn_add: bad batch");
  const size_t wbytes = (size_t)27 * Cout * Cs * 2;
  float* bias = reinterpret_cast<float*>(static_cast<char*>(workspace) + (((size_t)N * wbytes + 255) & ~(size_t)255));
  const int rf = km_fold_gn_part(w, scale, shift, workspace, bias, N, Cout, Cs + Cu, 0, Cs, 0, stream);
  if (rf != KM_OK) return rf;
  return launch_pair(x, workspace, bias, addend, out, stats, N, Cs, Cout, D, H, W, flags, stream);
}

namespace {
int launch_pair(const void* x, const void* wp, const float* bias_tab, const void* addend, void* out, float* stats,
                int N, int Cin, int Cout, int D, int H, int W, int flags, km_stream_t stream) {
  KM_CHECK_ARG(x && wp && out, "km_conv3d_tc_pair: null argument");
  KM_CHECK_ARG(km_conv3d_tc_pair_supported(Cin, Cout, D, H, W),
               "km_conv3d_tc_pair: unsupported shape (Cin=%d Cout=%d H=%d W=%d)", Cin, Cout, H, W);
  KM_CHECK_ARG(N > 0, "km_conv3d_tc_pair: bad batch");
  KM_CHECK_ARG(!(flags & KM_CONV_STATS) || stats, "km_conv3d_tc_pair: KM_CONV_STATS needs stats");
  KM_CHECK_ARG(!(flags & KM_CONV_COM), "km_conv3d_tc_pair: KM_CONV_COM is not supported");
  KM_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)wp & 15) == 0 && ((uintptr_t)out & 15) == 0,
               "km_conv3d_tc_pair: pointers must be 16-byte aligned");
  PairGeom g;
  memset(&g, 0, sizeof(g));
  g.N = N; g.D = D; g.H = H; g.W = W; g.Cin = Cin; g.Cout = Cout;
  g.flags = flags;
  g.w_per_sample = bias_tab ? 1 : 0;
  const int kc = (Cin % 64 == 0) ? 64 : 32;
  g.chunks = Cin / kc;
  const int row_bytes = kc * 2;
  g.BN = Cout;
  int mt = 4;
  while (mt > 1 && (2 * mt * g.BN > 512 || 16 * (mt / 2) >= H)) mt /= 2;
  g.mt = mt;
  g.issuers = mt >= 2 ? 2 : 1;
  g.subiters = 9 * g.chunks;
  g.a_sub_bytes = (16u * mt + 2u) * 8u * row_bytes;
  g.b_sub_bytes = 3u * (uint32_t)(g.BN / 2) * row_bytes;
  g.a_sub_stride = pair_round_up(g.a_sub_bytes, 1024);
  g.b_sub_stride = pair_round_up(g.b_sub_bytes, 1024);
  g.tiles_x = (W + 7) / 8;
  g.tiles_y = (H + 16 * mt - 1) / (16 * mt);
  const long long tiles = (long long)N * D * g.tiles_y * g.tiles_x;
  KM_CHECK_ARG(tiles < (1ll << 30), "km_conv3d_tc_pair: too many tiles");
  g.total_tiles = (int)tiles;
  g.stat_parts = (g.BN <= kEpiThreads && kEpiThreads % g.BN == 0) ? kEpiThreads / g.BN : 1;
  g.idesc = umma_idesc_16(256, g.BN, km_operand_fp16() != 0);

  const uint32_t kSmemMax = 232448 - 1024;
  g.staging_pitch = (uint32_t)g.BN * 2 + 16;
  g.eb = g.mt;
  while (g.eb > 1 && (uint32_t)g.eb * kTileM * g.staging_pitch > 48u * 1024u) g.eb /= 2;
  const uint32_t staging_bytes = pair_round_up((uint32_t)g.eb * kTileM * g.staging_pitch, 16);
  const uint32_t rowvalid_bytes = 4u * kTileM;
  const uint32_t stats_bytes = (flags & KM_CONV_STATS) ? (uint32_t)g.stat_parts * N * Cout * 2u * 4u : 0;
  const uint32_t bars_bytes = 8u * (2u * kMaxStages + 6u) + 16u;
  const uint32_t fixed = staging_bytes + rowvalid_bytes + stats_bytes + bars_bytes + 64;
  const uint32_t unit = g.a_sub_stride + g.b_sub_stride;
  KM_CHECK_ARG(fixed + 2 * unit <= kSmemMax, "km_conv3d_tc_pair: shared memory budget exceeded (N=%d Cout=%d)", N, Cout);
  const uint32_t avail = kSmemMax - fixed;
  int sub = 1;
  while (sub < g.subiters && (uint32_t)(sub + 1) * unit <= 40u * 1024u && (uint32_t)(sub + 1) * unit * 3u <= avail) ++sub;
  g.sub = sub;
  g.stage_stride = (uint32_t)sub * unit;
  int stages = (int)(avail / g.stage_stride);
  if (stages > kMaxStages) stages = kMaxStages;
  KM_CHECK_ARG(stages >= 2, "km_conv3d_tc_pair: not enough shared memory for a 2-stage pipeline");
  g.stages = stages;
  uint32_t off = (uint32_t)stages * g.stage_stride;
  g.off_staging = off; off += staging_bytes;
  g.off_rowvalid = off; off += rowvalid_bytes;
  g.off_stats = off; off += stats_bytes;
  off = pair_round_up(off, 8);
  g.off_bars = off; off += bars_bytes;
  const uint32_t smem_bytes = off + 1024;
  KM_CHECK_ARG(smem_bytes <= 232448, "km_conv3d_tc_pair: shared memory overflow (%u)", smem_bytes);

  PFN_encodeTiled encode = tensor_map_encoder();
  if (!encode) {
    km_set_error("km_conv3d_tc_pair: cuTensorMapEncodeTiled unavailable");
    return KM_ECUDA;
  }
  const CUtensorMapSwizzle swz = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2,
                             (cuuint64_t)D * H * W * Cin * 2};
    cuuint32_t box[5] = {(cuuint32_t)kc, 8, (cuuint32_t)(16 * g.mt + 2), 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tmA, KM_TMAP_16, 5, const_cast<void*>(x), dims, strides, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      km_set_error("km_conv3d_tc_pair: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
      return KM_ECUDA;
    }
  }
  {
    // [tap = dz*9 + dy*3 + dx][Cout][Cin] viewed as (Cin, Cout, dx, dy, dz): one box = the three dy
    // slices of a (dz, dx) group for HALF of the output channels, landing as three [BN/2 x kc] tiles
    const cuuint64_t slice = (cuuint64_t)Cout * Cin * 2;
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)Cout, 3, 3, (cuuint64_t)(bias_tab ? 3 * N : 3)};
    cuuint64_t strides[4] = {(cuuint64_t)Cin * 2, slice, 3 * slice, 9 * slice};
    cuuint32_t box[5] = {(cuuint32_t)kc, (cuuint32_t)(g.BN / 2), 1, 3, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tmB, KM_TMAP_16, 5, const_cast<void*>(wp), dims, strides, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      km_set_error("km_conv3d_tc_pair: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
      return KM_ECUDA;
    }
  }
  const bool f16 = km_operand_fp16() != 0;
  KM_CHECK_ARG(!addend || kc == 64, "km_conv3d_tc_pair: an addend needs Cin %% 64 == 0");
  PairKernel kernel = addend ? (f16 ? conv_tc2_kernel<64, true, true> : conv_tc2_kernel<64, false, true>)
                      : kc == 64 ? (f16 ? conv_tc2_kernel<64, true> : conv_tc2_kernel<64, false>)
                                 : (f16 ? conv_tc2_kernel<32, true> : conv_tc2_kernel<32, false>);
  static unsigned long long attr_set[2][3] = {};   // per (operand type, Cin chunk | addend) instantiation
  if (km_first_use_on_device(&attr_set[f16 ? 1 : 0][addend ? 2 : (kc == 64)]))
    KM_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  const int nsm = km_sm_count();
  int grid = nsm & ~1;
  const int pairs_needed = (g.total_tiles + 1) / 2;
  if (grid / 2 > pairs_needed) grid = 2 * pairs_needed;
  if ((flags & KM_CONV_STATS) && grid < nsm)
    KM_CUDA_OK(cudaMemsetAsync(stats, 0, (size_t)nsm * N * Cout * 2 * sizeof(float), km_cs(stream)));
  kernel<<<grid, kThreads, smem_bytes, km_cs(stream)>>>(tmA, tmB, g, reinterpret_cast<uint16_t*>(out), stats,
                                                        bias_tab, reinterpret_cast<const uint16_t*>(addend));
  KM_LAUNCH_OK("conv_tc2_kernel");
  return KM_OK;
}
}  // namespace
