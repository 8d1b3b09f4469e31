// Final 1x1x1 convolution fused with ReLU + centre of mass (+ power mass) on tcgen05, TRANSPOSED:
//
//     heat^T[channel, voxel] = W[channel, Cin] . X^T[Cin, voxel]
//
// Reference call sites: keymorph/unet3d/model.py:99,389 (final_conv), keymorph/layers.py:92-134
// (CenterOfMass3d: relu, marginal sums, linspace(0,1,n) weighted means), keymorph/model.py:95-109
// (weight_by_power: sum of relu(feat) per keypoint channel).
//
// Why transposed: the weights are the M operand (128 keypoint channels per MMA = the 128 TMEM lanes),
// the activation brick is the N operand (128 voxels = 128 TMEM columns).  Both are K-major in shared
// memory exactly as TMA delivers them (channels innermost), so nothing is physically transposed,
// but each epilogue thread now owns ONE keypoint channel and walks its voxels along TMEM columns:
// the four running sums [sum h, sum h*z, sum h*y, sum h*x] stay in registers for the whole kernel,
// no shared-memory transposes, no atomics, ~4 instructions per heat-map element.  The K x (S/2)^3
// heat map (2.1 GB fp32 per volume at K=256, S=256) never exists anywhere.
//
// Roles (576 threads, one persistent CTA per SM): warp 0 = TMA producer, warp 1 = MMA issuer,
// warps 2..17 = epilogue (two M-tiles of 128 channels x four TMEM lane quadrants x two halves of
// the brick's 128 voxel columns; the epilogue is the kernel's critical path, so it gets four
// warps per scheduler).  The two column halves write separate partial slots.
// TMEM: 2 buffers x 2 M-tiles x 128 columns = 512 columns; a work unit is (brick, pass) where pass
// p covers channels [256p, 256p+256).
#include "km_common.cuh"
#include "tc_ptx.cuh"

using namespace kmtc;

namespace {

constexpr int kThreads = 576;       // warp 0 TMA, warp 1 MMA, warps 2..17 epilogue
constexpr int kEpiThreads = 512;
constexpr int kBrick = 128;        // voxels per brick = MMA N
constexpr int kMaxStages = 12;
constexpr int kMaxPasses = 2;      // Cout <= 512

struct ComGeom {
  int N, D, H, W, Cin, Cout;
  int chunks;            // Cin / KC
  int TW, TH, TD;        // brick, TW*TH*TD == 128 (powers of two, x fastest)
  int tiles_x, tiles_y, tiles_z, tiles_per_img, total_tiles;
  int mtiles;            // Cout / 128
  int passes;            // ceil(mtiles / 2)
  int stages;
  uint32_t stage_bytes;  // chunks * 128 * KC * 2
  uint32_t off_w, off_tab, off_bars;
  uint32_t idesc;        // kind::f16 instruction descriptor (fp16 or bf16 operands)
};

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

template <int KC>
__global__ void __launch_bounds__(kThreads, 1)
com_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
              const ComGeom g, const float* __restrict__ bias, float* __restrict__ com) {
  constexpr int kRowBytes = KC * 2;
  constexpr int kSteps = KC / 16;
  constexpr uint32_t kLayout = kRowBytes == 128 ? 2u : (kRowBytes == 64 ? 4u : 6u);
  constexpr uint32_t kSbo = 8u * kRowBytes;
  constexpr uint32_t kTileBytes = 128u * kRowBytes;   // one [128 rows x KC] operand tile

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_u32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stages = g.stages;
  const uint32_t bars = base + g.off_bars;   // full[stages], empty[stages], tfull[2], tempty[2], wfull
  auto full_bar = [&](int s) { return bars + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bars + 8u * (uint32_t)(stages + s); };
  auto tfull_bar = [&](uint32_t a) { return bars + 8u * (uint32_t)(2 * stages + a); };
  auto tempty_bar = [&](uint32_t a) { return bars + 8u * (uint32_t)(2 * stages + 2 + a); };
  const uint32_t w_bar = bars + 8u * (uint32_t)(2 * stages + 4);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sm + g.off_bars + 8u * (2 * stages + 5));

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (uint32_t a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kEpiThreads);
    }
    mbar_init(w_bar, 1);
    fence_mbar_init();
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW);
  }
  __syncwarp();
  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_ptr_smem), 512);
    tmem_relinquish();
  }
  // brick-local voxel offsets of every column (generic epilogue path): (tx, ty, tz, 0)
  {
    float4* tab = reinterpret_cast<float4*>(sm + g.off_tab);
    for (int j = threadIdx.x; j < kBrick; j += kThreads)
      tab[j] = make_float4((float)(j % g.TW), (float)((j / g.TW) % g.TH), (float)(j / (g.TW * g.TH)), 0.f);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int passes = g.passes, mtiles = g.mtiles;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      // weights: resident for the whole kernel, tiles ordered [chunk][mtile]
      mbar_arrive_expect_tx(w_bar, (uint32_t)(g.chunks * mtiles) * kTileBytes);
      uint32_t dst = base + g.off_w;
      for (int ch = 0; ch < g.chunks; ++ch)
        for (int mt = 0; mt < mtiles; ++mt, dst += kTileBytes)
          tma_load_3d(dst, &tmW, w_bar, ch * KC, mt * 128, 0);
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
        int t = tile;
        const int x0 = (t % g.tiles_x) * g.TW;
        t /= g.tiles_x;
        const int y0 = (t % g.tiles_y) * g.TH;
        t /= g.tiles_y;
        const int z0 = (t % g.tiles_z) * g.TD;
        const int n = t / g.tiles_z;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_arrive_expect_tx(full_bar(s), g.stage_bytes);
        uint32_t a_dst = base + (uint32_t)s * g.stage_bytes;
        for (int ch = 0; ch < g.chunks; ++ch, a_dst += kTileBytes)
          tma_load_5d(a_dst, &tmX, full_bar(s), ch * KC, x0, y0, z0, n);
        if (++s == stages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    const uint32_t issue = elect_one();
    constexpr uint32_t desc_hi = (kSbo >> 4) | (1u << 14) | (kLayout << 29);
    const uint32_t lo_flag = 1u << 16;
    const uint32_t w16 = (((base + g.off_w) & 0x3FFFFu) >> 4) | lo_flag;
    const uint32_t x16_base = ((base & 0x3FFFFu) >> 4) | lo_flag;
    const uint32_t tile16 = kTileBytes >> 4;
    const uint32_t idesc = g.idesc;
    int s = 0;
    uint32_t ph = 0, ucount = 0;
    mbar_wait(w_bar, 0u);
    for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      const uint32_t x16 = x16_base + (uint32_t)s * (g.stage_bytes >> 4);
      for (int p = 0; p < passes; ++p, ++ucount) {
        const uint32_t buf = ucount & 1u, bph = (ucount >> 1) & 1u;
        mbar_wait(tempty_bar(buf), bph ^ 1u);
        tc_fence_after();
        for (int hm = 0; hm < 2; ++hm) {
          const int mt = 2 * p + hm;
          if (mt >= mtiles) break;
          const uint32_t d_tmem = tmem_base + buf * 256u + (uint32_t)hm * 128u;
          for (int ch = 0; ch < g.chunks; ++ch) {
            const uint32_t a16 = w16 + (uint32_t)(ch * mtiles + mt) * tile16;
            const uint32_t b16 = x16 + (uint32_t)ch * tile16;
#pragma unroll
            for (int kk = 0; kk < kSteps; ++kk)
              umma_16_pred(d_tmem, a16 + 2u * kk, b16 + 2u * kk, desc_hi, idesc,
                             (ch | kk) ? 1u : 0u, issue);
          }
        }
        umma_commit_pred(tfull_bar(buf), issue);   // accumulators of this pass complete -> epilogue
      }
      umma_commit_pred(empty_bar(s), issue);       // all passes have read the activation stage
      if (++s == stages) {
        s = 0;
        ph ^= 1u;
      }
    }
  } else {
    // =============================== epilogue (16 warps) =========================
    const int q = warp & 3;               // TMEM lane quadrant this warp may access
    const int ew = warp - 2;
    const int hm = (ew >> 2) & 1;         // M-tile of the pass handled by this warp
    const int ch2 = ew >> 3;              // half of the brick's voxel columns: [64 ch2, 64 ch2 + 64)
    const int crow = q * 32 + lane;       // channel within the M-tile == TMEM lane
    const float4* tab = reinterpret_cast<const float4*>(sm + g.off_tab);
    const float step_z = g.D > 1 ? 1.f / (float)(g.D - 1) : 0.f;
    const float step_y = g.H > 1 ? 1.f / (float)(g.H - 1) : 0.f;
    const float step_x = g.W > 1 ? 1.f / (float)(g.W - 1) : 0.f;
    const bool xrun = g.TW >= 32;         // a 32-column chunk is an x-run: ty, tz constant inside it
    float bch[kMaxPasses];
    float acc[kMaxPasses][4];
#pragma unroll
    for (int p = 0; p < kMaxPasses; ++p) {
      const int c = (2 * p + hm) * 128 + crow;
      bch[p] = (bias && c < g.Cout) ? __ldg(bias + c) : 0.f;
      acc[p][0] = acc[p][1] = acc[p][2] = acc[p][3] = 0.f;
    }
    auto flush = [&](int n) {
#pragma unroll
      for (int p = 0; p < kMaxPasses; ++p) {
        const int c = (2 * p + hm) * 128 + crow;
        if (p < passes && c < g.Cout) {
          *reinterpret_cast<float4*>(com + (((size_t)(2 * blockIdx.x + ch2) * g.N + n) * g.Cout + c) * 4) =
              make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]);
        }
        acc[p][0] = acc[p][1] = acc[p][2] = acc[p][3] = 0.f;
      }
    };
    uint32_t ucount = 0;
    int n_cur = -1;
    for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
      int t = tile;
      const int x0 = (t % g.tiles_x) * g.TW;
      t /= g.tiles_x;
      const int y0 = (t % g.tiles_y) * g.TH;
      t /= g.tiles_y;
      const int z0 = (t % g.tiles_z) * g.TD;
      const int n = t / g.tiles_z;
      if (n != n_cur) {
        if (n_cur >= 0) flush(n_cur);
        n_cur = n;
      }
      const bool full = x0 + g.TW <= g.W && y0 + g.TH <= g.H && z0 + g.TD <= g.D;
#pragma unroll
      for (int p = 0; p < kMaxPasses; ++p) {
        if (p >= passes) break;
        const uint32_t buf = ucount & 1u, bph = (ucount >> 1) & 1u;
        ++ucount;
        mbar_wait(tfull_bar(buf), bph);
        tc_fence_after();
        if (2 * p + hm < mtiles) {
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256u + (uint32_t)hm * 128u;
          const float b = bch[p];
          float t0 = 0.f, tzs = 0.f, tys = 0.f, txs = 0.f;   // brick-local sums
          uint32_t r[32];
#pragma unroll 1
          for (int cc = 0; cc < 2; ++cc) {
            const int ci = 2 * ch2 + cc;          // 32-column chunk of the brick
            tmem_ld32(taddr + 32u * (uint32_t)ci, r);
            tmem_ld_wait();
            if (xrun && full) {
              // four independent accumulator chains per sum
              float s0p[4] = {0.f, 0.f, 0.f, 0.f}, sxp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float h = fmaxf(__uint_as_float(r[j]) + b, 0.f);
                s0p[j & 3] += h;
                sxp[j & 3] = fmaf(h, (float)j, sxp[j & 3]);
              }
              const float s0 = (s0p[0] + s0p[1]) + (s0p[2] + s0p[3]);
              const float sx = (sxp[0] + sxp[1]) + (sxp[2] + sxp[3]);
              const float4 o = tab[ci * 32];   // offsets of the chunk's first voxel
              t0 += s0;
              txs += fmaf(o.x, s0, sx);
              tys = fmaf(o.y, s0, tys);
              tzs = fmaf(o.z, s0, tzs);
            } else {
              // bricks that stick out of the volume (TMA zero-fill would read as relu(bias)) and
              // narrow volumes: per-column offsets from the table, invalid voxels masked
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float4 o = tab[ci * 32 + j];
                const bool ok = x0 + (int)o.x < g.W && y0 + (int)o.y < g.H && z0 + (int)o.z < g.D;
                const float h = ok ? fmaxf(__uint_as_float(r[j]) + b, 0.f) : 0.f;
                t0 += h;
                txs = fmaf(h, o.x, txs);
                tys = fmaf(h, o.y, tys);
                tzs = fmaf(h, o.z, tzs);
              }
            }
          }
          // brick-local offsets -> linspace(0,1,n) coordinates: lin(i) = i / (n - 1)
          acc[p][0] += t0;
          acc[p][1] += step_z * fmaf((float)z0, t0, tzs);
          acc[p][2] += step_y * fmaf((float)y0, t0, tys);
          acc[p][3] += step_x * fmaf((float)x0, t0, txs);
        }
        tc_fence_before();
        mbar_arrive(tempty_bar(buf));
      }
    }
    if (n_cur >= 0) flush(n_cur);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


typedef void (*ComKernel)(const CUtensorMap, const CUtensorMap, const ComGeom, const float*, float*);

}  // namespace

extern "C" int km_sm_count(void);

extern "C" int km_conv1x1_com_nparts(void) { return 2 * km_sm_count(); }

extern "C" int km_conv1x1_com(const void* x, const void* wp, const float* bias, float* com, int N,
                              int Cin, int Cout, int D, int H, int W, km_stream_t stream) {
  KM_CHECK_ARG(x && wp && com, "km_conv1x1_com: null argument");
  KM_CHECK_ARG(N > 0 && D > 0 && H > 0 && W > 0, "km_conv1x1_com: bad shape");
  KM_CHECK_ARG(Cin % 16 == 0 && Cin >= 16 && Cin <= 256, "km_conv1x1_com: Cin must be a multiple of 16 in [16,256] (got %d)", Cin);
  KM_CHECK_ARG(Cout % 128 == 0 && Cout >= 128 && Cout <= 256 * kMaxPasses,
               "km_conv1x1_com: Cout must be a multiple of 128, at most %d (got %d; pad the weights)",
               256 * kMaxPasses, Cout);
  KM_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)wp & 15) == 0 && ((uintptr_t)com & 15) == 0,
               "km_conv1x1_com: pointers must be 16-byte aligned");
  ComGeom g;
  memset(&g, 0, sizeof(g));
  g.N = N; g.D = D; g.H = H; g.W = W; g.Cin = Cin; g.Cout = Cout;
  g.idesc = umma_idesc_16(128, kBrick, km_operand_fp16() != 0);
  const int kc = (Cin % 64 == 0) ? 64 : ((Cin % 32 == 0) ? 32 : 16);
  g.chunks = Cin / kc;
  const int row_bytes = kc * 2;
  auto pow2_le = [](int v, int cap) { int p = 1; while (p * 2 <= v && p * 2 <= cap) p *= 2; return p; };
  g.TW = pow2_le(W, kBrick);
  g.TH = pow2_le(H, kBrick / g.TW);
  g.TD = kBrick / (g.TW * g.TH);
  g.tiles_x = (W + g.TW - 1) / g.TW;
  g.tiles_y = (H + g.TH - 1) / g.TH;
  g.tiles_z = (D + g.TD - 1) / g.TD;
  g.tiles_per_img = g.tiles_x * g.tiles_y * g.tiles_z;
  const long long tiles = (long long)N * g.tiles_per_img;
  KM_CHECK_ARG(tiles < (1ll << 31), "km_conv1x1_com: too many tiles");
  g.total_tiles = (int)tiles;
  g.mtiles = Cout / 128;
  g.passes = (g.mtiles + 1) / 2;
  g.stage_bytes = (uint32_t)g.chunks * 128u * row_bytes;
  const uint32_t w_bytes = (uint32_t)g.chunks * g.mtiles * 128u * row_bytes;
  const uint32_t tab_bytes = kBrick * 16u;
  const uint32_t bars_bytes = 8u * (2u * kMaxStages + 6u) + 16u;
  const uint32_t kSmemMax = 232448 - 1024;
  KM_CHECK_ARG(w_bytes + tab_bytes + bars_bytes + 2 * g.stage_bytes <= kSmemMax,
               "km_conv1x1_com: weights do not fit in shared memory (Cin=%d Cout=%d)", Cin, Cout);
  int stages = (int)((kSmemMax - w_bytes - tab_bytes - bars_bytes) / g.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  g.stages = stages;
  uint32_t off = (uint32_t)stages * g.stage_bytes;
  g.off_w = off; off += w_bytes;
  g.off_tab = off; off += tab_bytes;
  g.off_bars = off; off += bars_bytes;
  const uint32_t smem_bytes = off + 1024;

  PFN_encodeTiled encode = tensor_map_encoder();
  if (!encode) {
    km_set_error("km_conv1x1_com: cuTensorMapEncodeTiled unavailable");
    return KM_ECUDA;
  }
  const CUtensorMapSwizzle swz = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                   : CU_TENSOR_MAP_SWIZZLE_32B;
  CUtensorMap tmX, tmW;
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2,
                             (cuuint64_t)D * H * W * Cin * 2};
    cuuint32_t box[5] = {(cuuint32_t)kc, (cuuint32_t)g.TW, (cuuint32_t)g.TH, (cuuint32_t)g.TD, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tmX, KM_TMAP_16, 5, const_cast<void*>(x), dims, strides,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      km_set_error("km_conv1x1_com: cuTensorMapEncodeTiled(X) failed with %d", (int)r);
      return KM_ECUDA;
    }
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout, 1};
    cuuint64_t strides[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)Cout * Cin * 2};
    cuuint32_t box[3] = {(cuuint32_t)kc, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&tmW, KM_TMAP_16, 3, const_cast<void*>(wp), dims, strides,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      km_set_error("km_conv1x1_com: cuTensorMapEncodeTiled(W) failed with %d", (int)r);
      return KM_ECUDA;
    }
  }
  ComKernel kernel = kc == 64 ? com_tc_kernel<64> : (kc == 32 ? com_tc_kernel<32> : com_tc_kernel<16>);
  static unsigned long long attr_set[3] = {0, 0, 0};
  const int ki = kc == 64 ? 2 : (kc == 32 ? 1 : 0);
  if (km_first_use_on_device(&attr_set[ki]))
    KM_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  const int nsm = km_sm_count();
  const int grid = g.total_tiles < nsm ? g.total_tiles : nsm;
  // a CTA only writes the (image, channel) slots of the images it worked on
  KM_CUDA_OK(cudaMemsetAsync(com, 0, (size_t)2 * nsm * N * Cout * 4 * sizeof(float), km_cs(stream)));
  kernel<<<grid, kThreads, smem_bytes, km_cs(stream)>>>(tmX, tmW, g, bias, com);
  KM_LAUNCH_OK("com_tc_kernel");
  return KM_OK;
}
