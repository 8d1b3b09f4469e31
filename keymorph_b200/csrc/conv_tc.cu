// tcgen05 implicit-GEMM 3-D convolution for the keypoint backbones.
//
// Replaces the cuDNN call sites keymorph/unet3d/buildingblocks.py:50-52 (Conv3d k3 p1, no bias),
// keymorph/layers.py:173-175 (Conv3d k3 p1 + bias) and keymorph/unet3d/model.py:99,389 (final
// 1x1x1 conv), the latter fused with keymorph/layers.py:92-134 (ReLU + centre of mass) so that the
// heat map never has to reach HBM.
//
// GEMM view:  D[128 voxels x BN] += A[128 voxels x KC] * B[BN x KC]^T  for every (tap, Cin chunk)
//   A = activation brick shifted by the tap offset, fetched by TMA from the bf16 NDHWC tensor
//       (out-of-bounds rows are zero-filled = conv padding);
//   B = weights [tap][Cout][Cin], K-major, fetched by TMA;
//   D = fp32 accumulator in TMEM, double buffered so the epilogue of tile i overlaps tile i+1.
// Roles (320 threads, one persistent CTA per SM): warp 0 = TMA producer, warp 1 = MMA issuer,
// warps 2..9 = epilogue in two column groups (tcgen05.ld -> bias/ReLU -> bf16 staging in smem ->
// coalesced stores, per-channel sum / sum-of-squares for the next GroupNorm, or centre-of-mass
// partials).
//
// Two data paths (template parameter MODE):
//   MODE 1 "x-halo reuse" (3x3x3, H >= 8, W >= 16): brick 16(x) x 8(y) x 1(z).  The tensor map
//     orders the dims (C, H, W, D, N) so that shared-memory rows run y-fastest; an x step is then
//     exactly one 8-row swizzle atom and the three dx taps read the SAME 18-column box through UMMA
//     descriptors whose start address differs by one atom: 9 activation boxes per Cin chunk
//     instead of 27.
//   MODE 0 generic (1x1x1, tiny volumes): one box per tap, brick = longest W-run x H x D.
// The single-thread producer / issuer loops are kept free of divisions and descriptor rebuilds:
// with one thread feeding the tensor core, instruction count per MMA is what limits throughput.
#include "km_common.cuh"
#include "tc_ptx.cuh"

using namespace kmtc;

namespace {

constexpr int kTileM = 128;
constexpr int kThreads = 352;     // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue, warp 10 second MMA issuer
constexpr int kEpiThreads = 256;
constexpr int kFstagePitch = 33;  // floats, conflict-free transposed access
constexpr int kMaxStages = 12;
int g_force_mode0 = 0;            // km_set_option(KM_OPT_CONV_FORCE_GENERIC): A/B the two data paths
int g_no_resident = 0;            // km_set_option(KM_OPT_CONV_NO_RESIDENT_WEIGHTS)
int g_max_mt = 4;                 // km_set_option(KM_OPT_CONV_MAX_BRICKS)
int g_no_epi_batch = 0;           // km_set_option(KM_OPT_CONV_NO_EPILOGUE_BATCH)
int g_halo_axis = 2;              // km_set_option(KM_OPT_CONV_HALO_AXIS): 1 = x-shift, 2 = y-shift
int g_two_issuers = 128;          // km_set_option(KM_OPT_CONV_TWO_ISSUERS): max BN that gets 2 issuer warps (0 = off)
int g_interleave = 0;             // km_set_option(KM_OPT_CONV_INTERLEAVE_BRICKS)

struct ConvGeom {
  int N, D, H, W, Cin, Cout;
  int taps;
  int chunks;          // Cin / KC
  int TW, TH, TD;      // output brick, TW*TH*TD == 128
  int tiles_x, tiles_y, tiles_z;
  int BN, n_blocks;    // output-channel block
  int stages;
  int flags;
  int has_out;
  int sub;             // sub-iterations (A box + B box) packed into one pipeline stage
  int subiters;        // sub-iterations per tile: taps*chunks (mode 0) or 9*chunks (mode 1)
  int stat_parts;      // row groups that accumulate channel statistics independently
  int b_resident;      // all weight slices stay in shared memory for the whole kernel
  int mt;              // mode 1: x-adjacent 16x8 bricks that share one weight fetch (1, 2 or 4)
  int eb;              // bricks staged together per epilogue round (<= mt)
  int interleave;      // issue order: bricks innermost (1) or taps innermost (0)
  int issuers;         // MMA issuer warps (1 or 2)
  uint32_t off_bres;
  uint32_t a_sub_bytes, b_sub_bytes;      // TMA bytes per sub-iteration
  uint32_t a_sub_stride, b_sub_stride;    // 1024-aligned slots inside a stage
  uint32_t stage_stride;
  uint32_t off_staging, off_fstage, off_rowinfo, off_stats, off_com, off_scratch, off_bias, off_bars;
  uint32_t staging_pitch;               // bytes per staged row (BN*2 + 16)
  uint32_t tmem_cols;
  uint32_t idesc;
  int total_tiles;
};

// accumulator row -> offset inside the output brick
template <int MODE>
__device__ __forceinline__ void row_to_voxel(const ConvGeom& g, int row, int& tx, int& ty, int& tz) {
  if (MODE == 0) {
    tx = row % g.TW;
    ty = (row / g.TW) % g.TH;
    tz = row / (g.TW * g.TH);
  } else if (MODE == 1) {  // rows are stored y-fastest: an x step is a whole 8-row swizzle atom
    ty = row & 7;
    tx = row >> 3;
    tz = 0;
  } else {                 // MODE 2: rows are stored x-fastest: a y step is a whole swizzle atom
    tx = row & 7;
    ty = row >> 3;
    tz = 0;
  }
}

struct TileCoord {
  int nb, n, x0, y0, z0;
};
__device__ __forceinline__ TileCoord decode_tile(const ConvGeom& g, int t) {
  TileCoord c;
  c.nb = t % g.n_blocks;
  t /= g.n_blocks;
  c.x0 = (t % g.tiles_x) * g.TW;
  t /= g.tiles_x;
  c.y0 = (t % g.tiles_y) * g.TH;
  t /= g.tiles_y;
  c.z0 = (t % g.tiles_z) * g.TD;
  t /= g.tiles_z;
  c.n = t;
  return c;
}

template <int KC, int MODE, bool F16>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const ConvGeom g, const float* __restrict__ bias, uint16_t* __restrict__ out,
               float* __restrict__ stats, float* __restrict__ com) {
  constexpr int kRowBytes = KC * 2;
  constexpr int kSteps = KC / 16;
  constexpr int kNtap = MODE >= 1 ? 3 : 1;
  constexpr uint32_t kLayout = kRowBytes == 128 ? 2u : (kRowBytes == 64 ? 4u : 6u);
  constexpr uint32_t kSbo = 8u * kRowBytes;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_u32);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = g.stages;
  const uint32_t stage_stride = g.stage_stride;

  const uint32_t bars = base + g.off_bars;  // full[stages], empty[stages], tfull[2], tempty[2]
  auto full_bar = [&](int s) { return bars + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bars + 8u * (uint32_t)(stages + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (uint32_t)(2 * stages + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (uint32_t)(2 * stages + 2 + a); };
  const uint32_t bres_bar = bars + 8u * (uint32_t)(2 * stages + 4);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sm + g.off_bars + 8u * (2 * stages + 5));

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), (uint32_t)g.issuers);     // one tcgen05.commit per issuer warp
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), (uint32_t)g.issuers);
      mbar_init(tempty_bar(a), kEpiThreads);
    }
    mbar_init(bres_bar, 1);
    fence_mbar_init();
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  __syncwarp();
  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_ptr_smem), g.tmem_cols);
    tmem_relinquish();
  }
  // zero the per-CTA accumulators
  {
    float* s_stats = reinterpret_cast<float*>(sm + g.off_stats);
    float* s_com = reinterpret_cast<float*>(sm + g.off_com);
    if (g.flags & KM_CONV_STATS)
      for (int i = threadIdx.x; i < g.stat_parts * g.N * g.Cout * 2; i += kThreads) s_stats[i] = 0.f;
    if (g.flags & KM_CONV_COM)
      for (int i = threadIdx.x; i < g.N * g.Cout * 4; i += kThreads) s_com[i] = 0.f;
    float* s_bias = reinterpret_cast<float*>(sm + g.off_bias);
    for (int i = threadIdx.x; i < g.Cout; i += kThreads) s_bias[i] = bias ? bias[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int sub = g.sub;
  const int n_stage_iters = (g.subiters + sub - 1) / sub;
  const int last_nsub = g.subiters - (n_stage_iters - 1) * sub;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const bool resident = g.b_resident != 0;
      const uint32_t sub_tx = g.a_sub_bytes + (resident ? 0u : g.b_sub_bytes);
      if (resident) {
        // weights: every (dz,dy) x Cin-chunk slice (3 dx taps each) is loaded once per CTA
        mbar_arrive_expect_tx(bres_bar, (uint32_t)g.subiters * g.b_sub_bytes);
        uint32_t dst = base + g.off_bres;
        for (int grp = 0; grp < g.subiters / g.chunks; ++grp)
          for (int ch = 0; ch < g.chunks; ++ch, dst += g.b_sub_stride)
            if (MODE == 2) tma_load_5d(dst, &tmB, bres_bar, ch * KC, 0, grp % 3, 0, grp / 3);
            else tma_load_3d(dst, &tmB, bres_bar, ch * KC, 0, MODE == 1 ? grp * 3 : grp);
      }
      for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(g, tile);
        const int bn0 = tc.nb * g.BN;
        int grp = 0, ch = 0;  // grp: tap (mode 0) or (dz,dy) pair (mode 1); ch: Cin chunk
        int dz = g.taps == 27 ? -1 : 0, dy = dz, dx = dz;
        for (int si = 0; si < n_stage_iters; ++si) {
          const int nsub = (si == n_stage_iters - 1) ? last_nsub : sub;
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_arrive_expect_tx(full_bar(s), (uint32_t)nsub * sub_tx);
          uint32_t a_dst = base + (uint32_t)s * stage_stride;
          uint32_t b_dst = a_dst + (uint32_t)sub * g.a_sub_stride;
          for (int u = 0; u < nsub; ++u) {
            if (MODE == 0) {
              tma_load_5d(a_dst, &tmA, full_bar(s), ch * KC, tc.x0 + dx, tc.y0 + dy, tc.z0 + dz,
                          tc.n);
              if (!resident) tma_load_3d(b_dst, &tmB, full_bar(s), ch * KC, bn0, grp);
            } else if (MODE == 1) {
              // tensor map dims are (C, H, W, D, N): rows land as h + 8 * w, 18 x-columns incl. halo
              tma_load_5d(a_dst, &tmA, full_bar(s), ch * KC, tc.y0 + dy, tc.x0 - 1, tc.z0 + dz,
                          tc.n);
              if (!resident) tma_load_3d(b_dst, &tmB, full_bar(s), ch * KC, bn0, grp * 3);
            } else {
              // MODE 2: dims (C, W, H, D, N), box 8 x-voxels (one atom, contiguous in global memory)
              // by 16*mt+2 y-rows; here `dy` counts the dx tap of this (dz, dx) group.  The weights
              // come through a 5-D map (Cin, Cout, dx, dy, dz): all three dy slices of (dz, dx).
              tma_load_5d(a_dst, &tmA, full_bar(s), ch * KC, tc.x0 + dy, tc.y0 - 1, tc.z0 + dz,
                          tc.n);
              if (!resident) tma_load_5d(b_dst, &tmB, full_bar(s), ch * KC, bn0, dy + 1, 0, dz + 1);
            }
            a_dst += g.a_sub_stride;
            b_dst += g.b_sub_stride;
            if (++ch == g.chunks) {
              ch = 0;
              ++grp;
              if (MODE == 0) {
                if (++dx == 2) {
                  dx = -1;
                  if (++dy == 2) {
                    dy = -1;
                    ++dz;
                  }
                }
              } else {
                if (++dy == 2) {
                  dy = -1;
                  ++dz;
                }
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
    // =============================== MMA issuer(s) ==============================
    // With g.issuers == 2 the bricks of a group are split between warp 1 and warp 10: one thread
    // cannot issue the short (N <= 64) MMAs of the narrow layers fast enough to keep the tensor
    // pipe busy.  Each issuer commits to the stage / accumulator barriers (arrival count 2).
    if (warp == 10 && g.issuers < 2) goto issuer_done;
    // The whole warp runs this loop with warp-uniform values (so that the compiler keeps the UMMA
    // descriptors in uniform registers); only the elected lane issues tcgen05.mma / commit.
    {
      const uint32_t issue = elect_one();   // 1 in exactly one lane
      int s = 0;
      uint32_t ph = 0;
      uint32_t tcount = 0;
      // UMMA smem descriptor: hi word is constant, lo word = (addr >> 4) | LBO(=1) << 16
      constexpr uint32_t desc_hi = (kSbo >> 4) | (1u << 14) | (kLayout << 29);
      const uint32_t lo_flag = 1u << 16;
      const uint32_t a_tap16 = (8u * kRowBytes) >> 4;                 // one swizzle atom = one tap step
      const uint32_t b_tap16 = ((uint32_t)g.BN * kRowBytes) >> 4;
      const uint32_t a_sub16 = g.a_sub_stride >> 4, b_sub16 = g.b_sub_stride >> 4;
      const uint32_t base16 = ((base & 0x3FFFFu) >> 4) | lo_flag;
      const uint32_t stage16 = stage_stride >> 4;
      const uint32_t boff16 = ((uint32_t)sub * g.a_sub_stride) >> 4;
      const uint32_t idesc = g.idesc;
      const bool resident = g.b_resident != 0;
      const uint32_t bres16 = (((base + g.off_bres) & 0x3FFFFu) >> 4) | lo_flag;
      const uint32_t bn = (uint32_t)g.BN;
      const bool g_interleave = g.interleave != 0;
      // this issuer's bricks [m_lo, m_lo + mt) of the group
      const int m_lo = (g.issuers == 2 && warp == 10) ? g.mt / 2 : 0;
      const int mt = g.issuers == 2 ? (warp == 10 ? g.mt - g.mt / 2 : g.mt / 2) : g.mt;
      if (resident) mbar_wait(bres_bar, 0u);
      for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t acc = tcount & 1u;
        const uint32_t acc_ph = (tcount >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_ph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (acc * (uint32_t)g.mt + (uint32_t)m_lo) * bn;
        uint32_t accum = 0;
        uint32_t bq16 = bres16;   // running weight slice (resident mode)
        for (int si = 0; si < n_stage_iters; ++si) {
          const int nsub = (si == n_stage_iters - 1) ? last_nsub : sub;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          uint32_t a16 = base16 + (uint32_t)s * stage16;
          uint32_t b16 = resident ? bq16 : a16 + boff16;
          a16 += (uint32_t)m_lo * 16u * a_tap16;
          for (int u = 0; u < nsub; ++u) {
            if (g_interleave) {
              // bricks innermost: consecutive MMAs accumulate into DIFFERENT TMEM tiles, so the
              // accumulator read-after-write latency of the tensor pipe overlaps across bricks
#pragma unroll
              for (int t = 0; t < kNtap; ++t) {
#pragma unroll
                for (int kk = 0; kk < kSteps; ++kk) {
                  uint32_t am16 = a16 + t * a_tap16 + 2u * kk, dm = d_tmem;
                  const uint32_t bb16 = b16 + t * b_tap16 + 2u * kk;
                  const uint32_t acc_flag = (t | kk) ? 1u : accum;
                  for (int m = 0; m < mt; ++m) {
                    umma_16_pred(dm, am16, bb16, desc_hi, idesc, acc_flag, issue);
                    am16 += 16u * a_tap16;
                    dm += bn;
                  }
                }
              }
            } else {
            uint32_t am16 = a16, dm = d_tmem;
            for (int m = 0; m < mt; ++m) {   // bricks sharing this weight slice
#pragma unroll
              for (int t = 0; t < kNtap; ++t) {
#pragma unroll
                for (int kk = 0; kk < kSteps; ++kk) {
                  umma_16_pred(dm, am16 + t * a_tap16 + 2u * kk, b16 + t * b_tap16 + 2u * kk,
                                 desc_hi, idesc, (t | kk) ? 1u : accum, issue);
                }
              }
              am16 += 16u * a_tap16;
              dm += bn;
            }
            }
            accum = 1u;
            a16 += a_sub16;
            b16 += b_sub16;
          }
          bq16 += (uint32_t)nsub * b_sub16;
          umma_commit_pred(empty_bar(s), issue);  // frees the smem stage when these MMAs have read it
          if (++s == stages) {
            s = 0;
            ph ^= 1u;
          }
        }
        umma_commit_pred(tfull_bar((int)acc), issue);  // accumulator complete -> epilogue
      }
    }
  issuer_done:;
  } else {
    // =============================== epilogue (8 warps) ==========================
    // Two groups of four warps; each warp reads its own TMEM lane quadrant (warp % 4), the groups
    // split the accumulator COLUMNS between them so that TMEM loads, conversion and the
    // centre-of-mass transposes of the two halves overlap.
    const int q = warp & 3;              // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;       // accumulator row == voxel within the brick
    const int half = (warp - 2) >> 2;    // 0: warps 2..5, 1: warps 6..9
    const int et = half * kTileM + row;  // epilogue thread id 0..255
    uint8_t* staging = sm + g.off_staging;
    float* fstage = reinterpret_cast<float*>(sm + g.off_fstage) + (size_t)half * 2 * kTileM * kFstagePitch;
    float* rowlin = reinterpret_cast<float*>(sm + g.off_rowinfo);          // [3][128] tz, ty, tx
    uint8_t* rowvalid = sm + g.off_rowinfo + 3 * kTileM * sizeof(float);   // [eb][128]
    float* s_stats = reinterpret_cast<float*>(sm + g.off_stats);           // [parts][N][Cout][2]
    float* s_com = reinterpret_cast<float*>(sm + g.off_com);
    float* scratch = reinterpret_cast<float*>(sm + g.off_scratch) + (size_t)half * 2 * 4 * 32 * 4;
    const float* s_bias = reinterpret_cast<const float*>(sm + g.off_bias);  // [Cout], zeros if none
    const float step_z = g.D > 1 ? 1.f / (float)(g.D - 1) : 0.f;
    const float step_y = g.H > 1 ? 1.f / (float)(g.H - 1) : 0.f;
    const float step_x = g.W > 1 ? 1.f / (float)(g.W - 1) : 0.f;
    const bool do_relu = (g.flags & KM_CONV_RELU) != 0;
    const bool do_stats = (g.flags & KM_CONV_STATS) != 0;
    const bool do_com = (g.flags & KM_CONV_COM) != 0;
    const bool has_out = g.has_out != 0;
    const int BN = g.BN;
    const uint32_t pitch = g.staging_pitch;
    const int cpr = BN / 8;  // 16-byte chunks per staged row
    const int parts = g.stat_parts;
    const int rows_per_part = kTileM / parts;
    auto half_bar = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory"); };
    auto all_bar = [&]() { asm volatile("bar.sync 3, 256;" ::: "memory"); };

    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++tcount) {
      const TileCoord tc0 = decode_tile(g, tile);
      const uint32_t acc = tcount & 1u;
      const uint32_t acc_ph = (tcount >> 1) & 1u;
      const int n0 = tc0.nb * BN;
      mbar_wait(tfull_bar((int)acc), acc_ph);
      tc_fence_after();
      int m_last = 0;   // last brick of the group that lies (partly) inside the volume
      if (MODE == 2) {
        while (m_last + 1 < g.mt && tc0.y0 + 16 * (m_last + 1) < g.H) ++m_last;
      } else {
        while (m_last + 1 < g.mt && tc0.x0 + 16 * (m_last + 1) < g.W) ++m_last;
      }
      const int eb = g.eb;
      for (int m = 0; m <= m_last; m += eb) {
      // one epilogue round: nb bricks staged together (fewer barriers per brick for narrow BN)
      const int nb = min(eb, m_last + 1 - m);
      TileCoord tc = tc0;
      if (MODE == 2) tc.y0 += 16 * m;
      else tc.x0 += 16 * m;
      const bool last_sub = m + nb - 1 == m_last;

      // voxel of this row (both groups write the same values)
      int tx, ty, tz;
      row_to_voxel<MODE>(g, row, tx, ty, tz);
      const int vz = tc.z0 + tz;
      uint32_t vmask = 0;
      for (int mb = 0; mb < nb; ++mb) {
        const int vx = tc.x0 + tx + (MODE == 2 ? 0 : 16 * mb);
        const int vy = tc.y0 + ty + (MODE == 2 ? 16 * mb : 0);
        const bool v = (vx < g.W) && (vy < g.H) && (vz < g.D);
        vmask |= (v ? 1u : 0u) << mb;
        rowvalid[mb * kTileM + row] = v ? 1 : 0;
      }
      const bool valid = (vmask & 1u) != 0;
      if (do_com) {
        rowlin[0 * kTileM + row] = (float)tz;   // brick-local offsets (narrow-brick CoM path)
        rowlin[1 * kTileM + row] = (float)ty;
        rowlin[2 * kTileM + row] = (float)tx;
      }

      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) +
                             (acc * (uint32_t)g.mt + (uint32_t)m) * (uint32_t)BN;

      if (!do_com) {
        const int nblk = BN / 16, split = (nblk + 1) / 2;
        const int b_lo = half == 0 ? 0 : split, b_hi = half == 0 ? split : nblk;
        for (int mb = 0; mb < nb; ++mb) {
          const bool vrow = ((vmask >> mb) & 1u) != 0;
          const uint32_t taddr_b = taddr + (uint32_t)(mb * BN);
          uint8_t* srow = staging + (size_t)(mb * kTileM + row) * pitch;
          for (int blk = b_lo; blk < b_hi; ++blk) {
            const int c0 = blk * 16;
            uint32_t r[16];
            tmem_ld16(taddr_b + (uint32_t)c0, r);
            tmem_ld_wait();
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float a = __uint_as_float(r[2 * j]) + s_bias[n0 + c0 + 2 * j];
              float b = __uint_as_float(r[2 * j + 1]) + s_bias[n0 + c0 + 2 * j + 1];
              if (do_relu) {
                a = fmaxf(a, 0.f);
                b = fmaxf(b, 0.f);
              }
              // rows outside the volume are staged as zeros so that the statistics need no mask
              pk[j] = vrow ? km_pack2<F16>(a, b) : 0u;
            }
            uint4* dst = reinterpret_cast<uint4*>(srow + (size_t)c0 * 2);
            dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
        if (last_sub) {
          tc_fence_before();
          mbar_arrive(tempty_bar((int)acc));
        }
        all_bar();
      } else {
        // final conv + centre of mass: 32-column chunks, transposed through fp32 smem.  Sums are
        // taken with brick-local voxel offsets (tx, ty, tz) and converted to linspace(0,1,n)
        // coordinates once per tile: lin(i) = i / (n - 1).
        const int nchunks = BN / 32, split = (nchunks + 1) / 2;
        const int c_lo = half == 0 ? 0 : split, c_hi = half == 0 ? split : nchunks;
        const bool wide = MODE == 0 && g.TW >= 32;   // a 32-row quarter is then an x-run: ty, tz constant
        const int col = row & 31, rq = row >> 5;
        const float q_tx = (float)((rq * 32) % g.TW);
        const float q_ty = (float)(((rq * 32) / g.TW) % g.TH);
        const float q_tz = (float)((rq * 32) / (g.TW * g.TH));
        auto combine = [&](int cc, int slot) {
          const float* sc = scratch + (size_t)slot * 4 * 32 * 4;
          float t[4];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            t[k] = ((sc[(0 * 32 + row) * 4 + k] + sc[(1 * 32 + row) * 4 + k]) +
                    sc[(2 * 32 + row) * 4 + k]) + sc[(3 * 32 + row) * 4 + k];
          float* dstc = s_com + ((size_t)tc.n * g.Cout + n0 + cc * 32 + row) * 4;
          dstc[0] += t[0];
          dstc[1] += step_z * fmaf((float)tc.z0, t[0], t[1]);
          dstc[2] += step_y * fmaf((float)tc.y0, t[0], t[2]);
          dstc[3] += step_x * fmaf((float)tc.x0, t[0], t[3]);
        };
        int it_c = 0;
        for (int cc = c_lo; cc < c_hi; ++cc, ++it_c) {
          const int c0 = cc * 32;
          float* fs = fstage + (size_t)(it_c & 1) * kTileM * kFstagePitch;
          uint32_t r0[16], r1[16];
          tmem_ld16(taddr + (uint32_t)c0, r0);
          tmem_ld16(taddr + (uint32_t)(c0 + 16), r1);
          tmem_ld_wait();
          if (cc == c_hi - 1 && last_sub) {
            tc_fence_before();
            mbar_arrive(tempty_bar((int)acc));
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float v0 = __uint_as_float(r0[j]) + s_bias[n0 + c0 + j];
            const float v1 = __uint_as_float(r1[j]) + s_bias[n0 + c0 + 16 + j];
            r0[j] = __float_as_uint(v0);
            r1[j] = __float_as_uint(v1);
            fs[row * kFstagePitch + j] = valid ? fmaxf(v0, 0.f) : 0.f;
            fs[row * kFstagePitch + 16 + j] = valid ? fmaxf(v1, 0.f) : 0.f;
          }
          if (has_out) {
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              pk[j] = km_pack2<F16>(__uint_as_float(r0[2 * j]), __uint_as_float(r0[2 * j + 1]));
            uint4* dst = reinterpret_cast<uint4*>(staging + (size_t)row * pitch + (size_t)c0 * 2);
            dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              pk[j] = km_pack2<F16>(__uint_as_float(r1[2 * j]), __uint_as_float(r1[2 * j + 1]));
            dst[2] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            dst[3] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
          half_bar();
          // combine the previous chunk's quarter partials (written before this barrier)
          if (it_c > 0 && row < 32) combine(cc - 1, (it_c - 1) & 1);
          // this thread: column `col`, rows [32*rq, +32)
          {
            float s0 = 0.f, sz, sy, sx;
            const float* fcol = fs + (size_t)(rq * 32) * kFstagePitch + col;
            if (wide) {
              float sxr = 0.f;
#pragma unroll
              for (int rr = 0; rr < 32; ++rr) {
                const float hv = fcol[rr * kFstagePitch];
                s0 += hv;
                sxr = fmaf(hv, (float)rr, sxr);
              }
              sx = fmaf(q_tx, s0, sxr);
              sy = q_ty * s0;
              sz = q_tz * s0;
            } else {
              sz = sy = sx = 0.f;
#pragma unroll 8
              for (int rr = 0; rr < 32; ++rr) {
                const int r2 = rq * 32 + rr;
                const float hv = fcol[rr * kFstagePitch];
                s0 += hv;
                sz = fmaf(hv, rowlin[0 * kTileM + r2], sz);
                sy = fmaf(hv, rowlin[1 * kTileM + r2], sy);
                sx = fmaf(hv, rowlin[2 * kTileM + r2], sx);
              }
            }
            float* sc = scratch + (size_t)(it_c & 1) * 4 * 32 * 4 + (size_t)(rq * 32 + col) * 4;
            sc[0] = s0;
            sc[1] = sz;
            sc[2] = sy;
            sc[3] = sx;
          }
        }
        if (c_hi > c_lo) {
          half_bar();
          if (row < 32) combine(c_hi - 1, (it_c - 1) & 1);
        } else if (last_sub) {
          tc_fence_before();
          mbar_arrive(tempty_bar((int)acc));   // this group had no columns in this tile
        }
        all_bar();
      }

      // ---- staged bf16 tile -> global (coalesced 16-byte chunks) + per-channel stats ----
      if (has_out) {
        const int total_chunks = nb * kTileM * cpr;
        for (int id = et; id < total_chunks; id += kEpiThreads) {
          const int j = id % cpr;
          const int rr = id / cpr;
          const int mb = rr >> 7;
          int r2 = rr & (kTileM - 1);
          if (MODE == 1) r2 = ((r2 & 15) << 3) | (r2 >> 4);   // walk x fastest for coalescing
          if (!rowvalid[mb * kTileM + r2]) continue;
          int tx2, ty2, tz2;
          row_to_voxel<MODE>(g, r2, tx2, ty2, tz2);
          const int x2 = tc.x0 + tx2 + (MODE == 2 ? 0 : 16 * mb);
          const int y2 = tc.y0 + ty2 + (MODE == 2 ? 16 * mb : 0), z2 = tc.z0 + tz2;
          const size_t vox = (((size_t)tc.n * g.D + z2) * g.H + y2) * g.W + x2;
          const uint4 v = *reinterpret_cast<const uint4*>(
              staging + (size_t)(mb * kTileM + r2) * pitch + j * 16);
          *reinterpret_cast<uint4*>(out + vox * g.Cout + n0 + j * 8) = v;
        }
      }
      if (do_stats) {
        // thread -> (column, row group); every (part, column) slot is owned by one thread
        for (int id = et; id < parts * BN; id += kEpiThreads) {
          const int col = id % BN, part = id / BN;
          float s = 0.f, ss = 0.f;
          for (int mb = 0; mb < nb; ++mb) {
            const uint8_t* p = staging + (size_t)(mb * kTileM + part * rows_per_part) * pitch +
                               (size_t)col * 2;
#pragma unroll 8
            for (int rr = 0; rr < rows_per_part; ++rr) {
              const float v = km_to_float<F16>(*reinterpret_cast<const uint16_t*>(p));
              p += pitch;
              s += v;
              ss = fmaf(v, v, ss);
            }
          }
          float* d = s_stats + (((size_t)part * g.N + tc.n) * g.Cout + n0 + col) * 2;
          d[0] += s;
          d[1] += ss;
        }
      }
      all_bar();  // staging / rowinfo may be overwritten by the next brick / tile
      }
    }

    // ---- per-CTA partials -> global ----
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
    if (do_com) {
      float* dst = com + (size_t)blockIdx.x * g.N * g.Cout * 4;
      for (int i = et; i < g.N * g.Cout * 4; i += kEpiThreads) dst[i] = s_com[i];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, g.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------
// weight repack: fp32 (Cout, Cin, taps) -> bf16 [tap][Cout][Cin]
template <bool F16>
__global__ void pack_weights_kernel(const float* __restrict__ w, uint16_t* __restrict__ p,
                                    int Cout, int Cin, int taps) {
  const long long total = (long long)taps * Cout * Cin;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const int co = (int)((i / Cin) % Cout);
    const int tap = (int)(i / ((long long)Cin * Cout));
    p[i] = km_from_float<F16>(w[((long long)co * Cin + ci) * taps + tap]);
  }
}


int g_sm_count = 0;
int sm_count() {
  if (g_sm_count == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      g_sm_count = 148;
  }
  return g_sm_count;
}

inline uint32_t round_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

typedef void (*ConvKernel)(const CUtensorMap, const CUtensorMap, const ConvGeom, const float*,
                           uint16_t*, float*, float*);

template <bool F16>
ConvKernel pick_kernel_t(int kc, int mode) {
  if (mode == 2) {
    if (kc == 64) return conv_tc_kernel<64, 2, F16>;
    if (kc == 32) return conv_tc_kernel<32, 2, F16>;
    return conv_tc_kernel<16, 2, F16>;
  }
  if (mode == 1) {
    if (kc == 64) return conv_tc_kernel<64, 1, F16>;
    if (kc == 32) return conv_tc_kernel<32, 1, F16>;
    return conv_tc_kernel<16, 1, F16>;
  }
  if (kc == 64) return conv_tc_kernel<64, 0, F16>;
  if (kc == 32) return conv_tc_kernel<32, 0, F16>;
  return conv_tc_kernel<16, 0, F16>;
}
ConvKernel pick_kernel(int kc, int mode) {
  return km_operand_fp16() ? pick_kernel_t<true>(kc, mode) : pick_kernel_t<false>(kc, mode);
}

}  // namespace

extern "C" int km_sm_count(void) { return sm_count(); }
void km_conv_set_force_generic(int v) { g_force_mode0 = v ? 1 : 0; }
void km_conv_set_no_resident(int v) { g_no_resident = v ? 1 : 0; }
void km_conv_set_max_mt(int v) { g_max_mt = v >= 4 ? 4 : (v >= 2 ? 2 : 1); }
void km_conv_set_no_epi_batch(int v) { g_no_epi_batch = v ? 1 : 0; }
void km_conv_set_halo_axis(int v) { g_halo_axis = v == 1 ? 1 : 2; }
void km_conv_set_interleave(int v) { g_interleave = v ? 1 : 0; }
void km_conv_set_two_issuers(int v) { g_two_issuers = v; }
extern "C" int km_conv_nparts(void) { return sm_count(); }

extern "C" int km_pack_weights(const float* w, void* packed, int Cout, int Cin, int taps,
                               km_stream_t stream) {
  KM_CHECK_ARG(w && packed && Cout > 0 && Cin > 0 && (taps == 27 || taps == 1),
               "km_pack_weights: bad arguments");
  const long long total = (long long)taps * Cout * Cin;
  const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  if (km_operand_fp16())
    pack_weights_kernel<true><<<blocks, 256, 0, km_cs(stream)>>>(w, reinterpret_cast<uint16_t*>(packed), Cout, Cin, taps);
  else
    pack_weights_kernel<false><<<blocks, 256, 0, km_cs(stream)>>>(w, reinterpret_cast<uint16_t*>(packed), Cout, Cin, taps);
  KM_LAUNCH_OK("pack_weights_kernel");
  return KM_OK;
}

extern "C" int km_conv3d_tc(const void* x, const void* wp, const float* bias, void* out,
                            float* stats, float* com, int N, int Cin, int Cout, int D, int H,
                            int W, int taps, int flags, km_stream_t stream) {
  KM_CHECK_ARG(x && wp, "km_conv3d_tc: null input");
  KM_CHECK_ARG(taps == 27 || taps == 1, "km_conv3d_tc: taps must be 27 or 1");
  KM_CHECK_ARG(N > 0 && D > 0 && H > 0 && W > 0, "km_conv3d_tc: bad shape");
  KM_CHECK_ARG(Cin % 16 == 0 && Cin >= 16, "km_conv3d_tc: Cin must be a multiple of 16 (got %d)",
               Cin);
  KM_CHECK_ARG(Cout % 16 == 0 && Cout >= 16, "km_conv3d_tc: Cout must be a multiple of 16 (got %d)",
               Cout);
  KM_CHECK_ARG(!(flags & KM_CONV_STATS) || stats, "km_conv3d_tc: KM_CONV_STATS needs stats");
  KM_CHECK_ARG(!(flags & KM_CONV_COM) || com, "km_conv3d_tc: KM_CONV_COM needs com");
  KM_CHECK_ARG(out || (flags & KM_CONV_COM), "km_conv3d_tc: out may only be NULL with KM_CONV_COM");
  KM_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)wp & 15) == 0 && ((uintptr_t)out & 15) == 0,
               "km_conv3d_tc: pointers must be 16-byte aligned");

  ConvGeom g;
  memset(&g, 0, sizeof(g));
  g.N = N; g.D = D; g.H = H; g.W = W; g.Cin = Cin; g.Cout = Cout;
  g.taps = taps;
  g.flags = flags;
  g.has_out = out ? 1 : 0;
  const int kc = (Cin % 64 == 0) ? 64 : ((Cin % 32 == 0) ? 32 : 16);
  g.chunks = Cin / kc;
  const int row_bytes = kc * 2;
  // halo-reuse paths: 2 = y is the shifted axis (atoms are x-runs, contiguous in global memory),
  // 1 = x is the shifted axis; 0 = generic
  int mode = 0;
  if (taps == 27 && !g_force_mode0) {
    if (g_halo_axis == 2 && W >= 8 && H >= 16) mode = 2;
    else if (H >= 8 && W >= 16) mode = 1;
    else if (W >= 8 && H >= 16) mode = 2;
  }
  const int gran = (flags & KM_CONV_COM) ? 32 : 16;
  KM_CHECK_ARG(Cout % gran == 0, "km_conv3d_tc: Cout must be a multiple of %d in this mode", gran);
  const int bn_cap = mode >= 1 ? 128 : 256;
  g.BN = 0;
  for (int bn = bn_cap; bn >= gran; bn -= gran)
    if (Cout % bn == 0) { g.BN = bn; break; }
  KM_CHECK_ARG(g.BN > 0, "km_conv3d_tc: no channel block for Cout=%d", Cout);
  g.n_blocks = Cout / g.BN;
  g.mt = 1;
  g.interleave = g_interleave;
  g.issuers = 1;
  if (mode >= 1) {
    // bricks that share one weight fetch: as many as TMEM (2 buffers x mt x BN columns) allows,
    // not more than the row / column holds; shared memory is checked below
    int mt = g_max_mt;
    const int extent = mode == 1 ? W : H;
    while (mt > 1 && (2 * mt * g.BN > 512 || 16 * (mt / 2) >= extent)) mt /= 2;
    while (mt > 1 && round_up((16u * mt + 2u) * 8u * row_bytes, 1024) +
                             round_up(3u * (uint32_t)g.BN * row_bytes, 1024) > 96u * 1024u)
      mt /= 2;
    g.mt = mt;
    g.issuers = (g_two_issuers && mt >= 2 && g.BN <= g_two_issuers) ? 2 : 1;
    if (mode == 1) { g.TW = 16 * mt; g.TH = 8; }
    else { g.TW = 8; g.TH = 16 * mt; }
    g.TD = 1;
    g.subiters = 9 * g.chunks;
    g.a_sub_bytes = (16u * mt + 2u) * 8u * row_bytes;
    g.b_sub_bytes = 3u * (uint32_t)g.BN * row_bytes;
  } else {
    // generic brick: as long a W-run as possible, then H, then D (power-of-two factors of 128).
    // Volumes with fewer than 128 voxels (deepest level of small inputs): the brick sticks out of
    // the tensor; TMA zero-fills the out-of-bounds part and the epilogue masks those rows.
    auto pow2_le = [](int v, int cap) { int p = 1; while (p * 2 <= v && p * 2 <= cap) p *= 2; return p; };
    g.TW = pow2_le(W, 128);
    g.TH = pow2_le(H, 128 / g.TW);
    g.TD = 128 / (g.TW * g.TH);
    g.subiters = taps * g.chunks;
    g.a_sub_bytes = (uint32_t)kTileM * row_bytes;
    g.b_sub_bytes = (uint32_t)g.BN * row_bytes;
  }
  g.a_sub_stride = round_up(g.a_sub_bytes, 1024);
  g.b_sub_stride = round_up(g.b_sub_bytes, 1024);
  g.tiles_x = (W + g.TW - 1) / g.TW;
  g.tiles_y = (H + g.TH - 1) / g.TH;
  g.tiles_z = (D + g.TD - 1) / g.TD;
  const long long tiles = (long long)N * g.tiles_z * g.tiles_y * g.tiles_x * g.n_blocks;
  KM_CHECK_ARG(tiles < (1ll << 31), "km_conv3d_tc: too many tiles");
  g.total_tiles = (int)tiles;
  g.stat_parts = (g.BN <= kEpiThreads && kEpiThreads % g.BN == 0) ? kEpiThreads / g.BN : 1;

  g.idesc = umma_idesc_16(kTileM, g.BN, km_operand_fp16() != 0);
  uint32_t cols = 32;
  while (cols < 2u * (uint32_t)(g.mt * g.BN)) cols *= 2;
  g.tmem_cols = cols;
  KM_CHECK_ARG(cols <= 512, "km_conv3d_tc: TMEM overflow");

  // shared-memory carve-up (offsets relative to the 1024-aligned base)
  const uint32_t kSmemMax = 232448 - 1024;  // 227 KB minus alignment slack
  g.staging_pitch = (uint32_t)g.BN * 2 + 16;
  // bricks staged per epilogue round: all bricks of a group when that costs <= 48 KB
  g.eb = 1;
  if (g.has_out && !(flags & KM_CONV_COM) && !g_no_epi_batch) {
    g.eb = g.mt;
    while (g.eb > 1 && (uint32_t)g.eb * kTileM * g.staging_pitch > 48u * 1024u) g.eb /= 2;
  }
  const uint32_t staging_bytes = g.has_out ? round_up((uint32_t)g.eb * kTileM * g.staging_pitch, 16) : 0;
  const uint32_t fstage_bytes = (flags & KM_CONV_COM) ? 2u * 2u * kTileM * kFstagePitch * 4u : 0;
  const uint32_t rowinfo_bytes = 3u * kTileM * 4u + 4u * kTileM;
  const uint32_t stats_bytes =
      (flags & KM_CONV_STATS) ? (uint32_t)g.stat_parts * N * Cout * 2u * 4u : 0;
  const uint32_t com_bytes = (flags & KM_CONV_COM) ? (uint32_t)N * Cout * 4u * 4u : 0;
  const uint32_t scratch_bytes = (flags & KM_CONV_COM) ? 2u * 2u * 4u * 32u * 4u * 4u : 0;
  const uint32_t bars_bytes = 8u * (2u * kMaxStages + 6u) + 16u;
  const uint32_t bias_bytes = (uint32_t)Cout * 4u;
  const uint32_t fixed = staging_bytes + fstage_bytes + rowinfo_bytes + stats_bytes + com_bytes +
                         scratch_bytes + bias_bytes + bars_bytes + 64;
  // small weight tensors (first layers) stay resident: the per-tile weight re-fetch is TMA-request bound
  const uint32_t bres_bytes = (uint32_t)g.subiters * g.b_sub_stride;
  g.b_resident = (mode >= 1 && g.n_blocks == 1 && bres_bytes <= 64u * 1024u && !g_no_resident) ? 1 : 0;
  const uint32_t unit = g.a_sub_stride + (g.b_resident ? 0u : g.b_sub_stride);
  KM_CHECK_ARG(fixed + (g.b_resident ? bres_bytes : 0u) + 2 * unit <= kSmemMax,
               "km_conv3d_tc: shared memory budget exceeded (N*Cout too large: N=%d Cout=%d)", N, Cout);
  const uint32_t avail = kSmemMax - fixed - (g.b_resident ? bres_bytes : 0u);
  // pack sub-iterations into a stage until it holds ~40 KB (fewer barrier round trips, more bytes
  // in flight per barrier) while keeping at least 3 stages
  int sub = 1;
  while (sub < g.subiters && (uint32_t)(sub + 1) * unit <= 40u * 1024u &&
         (uint32_t)(sub + 1) * unit * 3u <= avail)
    ++sub;
  g.sub = sub;
  g.stage_stride = (uint32_t)sub * unit;
  int stages = (int)(avail / g.stage_stride);
  if (stages > kMaxStages) stages = kMaxStages;
  KM_CHECK_ARG(stages >= 2, "km_conv3d_tc: not enough shared memory for a 2-stage pipeline");
  g.stages = stages;
  uint32_t off = (uint32_t)stages * g.stage_stride;
  g.off_bres = off;
  if (g.b_resident) off += bres_bytes;
  g.off_staging = off; off += staging_bytes;
  g.off_fstage = off; off += fstage_bytes;
  g.off_rowinfo = off; off += round_up(rowinfo_bytes, 16);
  g.off_stats = off; off += stats_bytes;
  g.off_com = off; off += com_bytes;
  g.off_scratch = off; off += scratch_bytes;
  g.off_bias = off; off += bias_bytes;
  off = round_up(off, 8);
  g.off_bars = off; off += bars_bytes;
  const uint32_t smem_bytes = off + 1024;
  KM_CHECK_ARG(smem_bytes <= 232448, "km_conv3d_tc: shared memory overflow (%u)", smem_bytes);

  PFN_encodeTiled encode = tensor_map_encoder();
  if (!encode) {
    km_set_error("km_conv3d_tc: cuTensorMapEncodeTiled unavailable");
    return KM_ECUDA;
  }
  const CUtensorMapSwizzle swz = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                   : CU_TENSOR_MAP_SWIZZLE_32B;
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2,
                             (cuuint64_t)H * W * Cin * 2, (cuuint64_t)D * H * W * Cin * 2};
    cuuint32_t box[5] = {(cuuint32_t)kc, (cuuint32_t)g.TW, (cuuint32_t)g.TH, (cuuint32_t)g.TD, 1};
    if (mode == 1) {
      // (C, H, W, D, N): y is the fastest spatial index in shared memory (8 rows = one swizzle atom)
      dims[1] = (cuuint64_t)H; dims[2] = (cuuint64_t)W;
      strides[0] = (cuuint64_t)W * Cin * 2; strides[1] = (cuuint64_t)Cin * 2;
      box[1] = 8; box[2] = (cuuint32_t)(16 * g.mt + 2); box[3] = 1;
    } else if (mode == 2) {
      box[1] = 8; box[2] = (cuuint32_t)(16 * g.mt + 2); box[3] = 1;   // natural (C, W, H, D, N) order
    }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tmA, KM_TMAP_16, 5, const_cast<void*>(x), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      km_set_error("km_conv3d_tc: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
      return KM_ECUDA;
    }
  }
  if (mode == 2) {
    // [tap = dz*9 + dy*3 + dx][Cout][Cin] viewed as (Cin, Cout, dx, dy, dz): one box = the three dy
    // slices of a (dz, dx) group, landing as three consecutive [BN x kc] tiles
    const cuuint64_t slice = (cuuint64_t)Cout * Cin * 2;
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)Cout, 3, 3, 3};
    cuuint64_t strides[4] = {(cuuint64_t)Cin * 2, slice, 3 * slice, 9 * slice};
    cuuint32_t box[5] = {(cuuint32_t)kc, (cuuint32_t)g.BN, 1, 3, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tmB, KM_TMAP_16, 5, const_cast<void*>(wp), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      km_set_error("km_conv3d_tc: cuTensorMapEncodeTiled(B5) failed with %d", (int)r);
      return KM_ECUDA;
    }
  } else {
    cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout, (cuuint64_t)taps};
    cuuint64_t strides[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)Cout * Cin * 2};
    cuuint32_t box[3] = {(cuuint32_t)kc, (cuuint32_t)g.BN, (cuuint32_t)(mode == 1 ? 3 : 1)};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&tmB, KM_TMAP_16, 3, const_cast<void*>(wp), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      km_set_error("km_conv3d_tc: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
      return KM_ECUDA;
    }
  }

  ConvKernel kernel = pick_kernel(kc, mode);
  static unsigned long long attr_set[2][3][3] = {};   // per (operand type, mode, Cin chunk) instantiation
  const int ki = kc == 64 ? 2 : (kc == 32 ? 1 : 0);
  if (km_first_use_on_device(&attr_set[km_operand_fp16() ? 1 : 0][mode][ki]))
    KM_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  const int nsm = sm_count();
  const int grid = g.total_tiles < nsm ? g.total_tiles : nsm;
  // partial slots of CTAs that are not launched must still be defined
  if (grid < nsm) {
    if (flags & KM_CONV_STATS)
      KM_CUDA_OK(cudaMemsetAsync(stats, 0, (size_t)nsm * N * Cout * 2 * sizeof(float), km_cs(stream)));
    if (flags & KM_CONV_COM)
      KM_CUDA_OK(cudaMemsetAsync(com, 0, (size_t)nsm * N * Cout * 4 * sizeof(float), km_cs(stream)));
  }
  kernel<<<grid, kThreads, smem_bytes, km_cs(stream)>>>(
      tmA, tmB, g, bias, reinterpret_cast<uint16_t*>(out), stats, com);
  KM_LAUNCH_OK("conv_tc_kernel");
  return KM_OK;
}
