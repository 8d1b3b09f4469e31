// Stem convolution: 1-channel fp32 volume -> COUT (16 | 32) channels, 3x3x3, pad 1, as an implicit
// GEMM on the warp-level tensor-core path (mma.sync m16n8k8, TF32 operands, fp32 accumulate):
// M = 16 consecutive x voxels, N = COUT, K = 27 taps padded to 32.
//
// Reference call sites: keymorph/unet3d/buildingblocks.py:39-132 (first SingleConv of encoder 0:
// GroupNorm(1 group) -> Conv3d(1,16) -> ReLU, followed by the GroupNorm of the next SingleConv) and
// keymorph/layers.py:137-187 (ConvNet block 1: Conv3d(1,32,bias) -> InstanceNorm3d -> ReLU).
//
// The layer is run TWICE instead of being followed by a normalisation pass over its output: the
// volume is tiny (4 B/voxel) compared with the layer's output (2*COUT B/voxel), so
//   pass 1 (out == NULL): statistics of the pre-normalisation output only, nothing stored;
//   pass 2              : recompute, apply the following layer's normalisation (out_scale/out_shift),
//                         store the bf16 NDHWC tensor the tcgen05 convolution consumes.
// This removes one full write + read + write of the (N,D,H,W,COUT) tensor (1 GB at 256^3, COUT 16,
// two volumes) and one bf16 rounding.
//
// A operand: the CTA stages a 66 x 6 x 6 halo tile of the (input-normalised, zero-padded) volume in
// shared memory; every thread gathers its m16n8k8 A fragments straight from that tile (k = tap).
// B operand: the 27 x COUT weights live in registers as TF32 fragments for the whole kernel.
#include "km_common.cuh"

namespace {

typedef uint16_t act16;   // fp16 or bf16 bits (template parameter F16, km_common.cuh)

constexpr int kTX = 64, kTY = 4, kTZ = 4;            // outputs per CTA tile
constexpr int kHX = kTX + 2, kHY = kTY + 2, kHZ = kTZ + 2;
constexpr int kSX = 74;                               // smem row stride (words): 74 % 32 = 10 keeps
                                                      // the (x, dy) gathers of one LDS on distinct banks
constexpr int kHalo = kHZ * kHY * kHX;
constexpr int kPre = (kHalo + 255) / 256;

__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                         uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// grid (km_stem_nparts, N), 256 threads, 2 CTAs / SM
template <int COUT, bool F16>
__global__ void __launch_bounds__(256, 2)
conv_stem_mma_kernel(const float* __restrict__ x, const float* __restrict__ w,
                     const float* __restrict__ bias, const float* __restrict__ in_scale,
                     const float* __restrict__ in_shift, const float* __restrict__ out_scale,
                     const float* __restrict__ out_shift, act16* __restrict__ out,
                     float* __restrict__ stats, int N, int D, int H, int W, int relu_pre,
                     int relu_post) {
  constexpr int NT = COUT / 8;            // n-tiles of 8 channels
  constexpr int kStage = COUT * 2 + 16;   // bytes per voxel in the store-staging buffer (48 | 80)
  __shared__ float tile[kHZ * kHY * kSX];
  __shared__ __align__(16) unsigned char stage[8][16 * kStage];
  __shared__ float red[8][2 * COUT];
  __shared__ __align__(8) float cst[3 * COUT];

  const int n = blockIdx.y;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;   // mma "groupID" / "threadID_in_group"

  // ---- B fragments (weights) and per-thread channel constants ---------------------------------
  uint32_t bfrag[4][NT][2];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int k = ks * 8 + t + 4 * j;
        bfrag[ks][nt][j] = to_tf32(k < 27 ? __ldg(w + (nt * 8 + g) * 27 + k) : 0.f);
      }
  int aoff[4][2];   // smem offset of tap k = ks*8 + t + 4j relative to the output voxel's halo origin
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int k = ks * 8 + t + 4 * j;
      aoff[ks][j] = k < 27 ? ((k / 9) * kHY + (k / 3) % 3) * kSX + k % 3 : 0;
    }
  // per-channel epilogue constants: [bias | out_scale | out_shift], read back as float2 per n-tile
  for (int c = threadIdx.x; c < COUT; c += 256) {
    cst[c] = bias ? __ldg(bias + c) : 0.f;
    cst[COUT + c] = out_scale ? __ldg(out_scale + (size_t)n * COUT + c) : 1.f;
    cst[2 * COUT + c] = out_shift ? __ldg(out_shift + (size_t)n * COUT + c) : 0.f;
  }
  const float a_in = in_scale ? __ldg(in_scale + n) : 1.f;
  const float b_in = in_shift ? __ldg(in_shift + n) : 0.f;
  const float* xn = x + (size_t)n * D * H * W;
  act16* on = out ? out + (size_t)n * D * H * W * COUT : nullptr;

  const int tiles_x = (W + kTX - 1) / kTX, tiles_y = (H + kTY - 1) / kTY;
  const int tiles_z = (D + kTZ - 1) / kTZ;
  const int ntiles = tiles_x * tiles_y * tiles_z;

  float s[NT][2], ss[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) s[nt][0] = s[nt][1] = ss[nt][0] = ss[nt][1] = 0.f;

  // software pipeline: the halo of the NEXT tile is fetched into registers while the current one
  // is being multiplied (raw values + a validity mask travel in registers; the input normalisation
  // is applied when the tile is written to shared memory, so that nothing waits on the loads before
  // the MMA loop).  Thread -> halo elements: column hx = t & 63 of rows (t >> 6) + 4 j, j = 0..8
  // (row = hz * 6 + hy), plus the two extra columns 64 / 65 of row t >> 1 for t < 72: the index
  // arithmetic per element is a handful of instructions and the row offsets are immediates.
  constexpr int kRowsPerThread = (kHZ * kHY) / 4;   // 9
  static_assert(kHZ * kHY == 36 && kHX == 66, "halo mapping assumes a 66 x 6 x 6 tile");
  float pre[kRowsPerThread + 1];
  uint32_t pre_ok = 0;
  const int f_hx = threadIdx.x & 63, f_r0 = threadIdx.x >> 6;
  const int e_r = threadIdx.x >> 1, e_hx = 64 + (threadIdx.x & 1);   // extra columns (t < 72)
  const int HW = H * W;
  int ftl = blockIdx.x;   // tile being fetched, decoded incrementally (no divisions per tile)
  int ftx = ftl % tiles_x, fty = (ftl / tiles_x) % tiles_y, ftz = ftl / (tiles_x * tiles_y);
  const int stx = (int)gridDim.x % tiles_x, sty = ((int)gridDim.x / tiles_x) % tiles_y;
  const int stz = (int)gridDim.x / (tiles_x * tiles_y);
  int cx0 = 0, cy0 = 0, cz0 = 0;   // origin of the tile whose halo sits in `pre`
  auto fetch = [&]() {
    cx0 = ftx * kTX;
    cy0 = fty * kTY;
    cz0 = ftz * kTZ;
    const int gx = cx0 - 1 + f_hx;
    const bool okx = (unsigned)gx < (unsigned)W;
    pre_ok = 0;
#pragma unroll
    for (int j = 0; j < kRowsPerThread; ++j) {
      const int r = f_r0 + 4 * j;
      const int hz = (r * 43) >> 8, hy = r - 6 * hz;   // r / 6, r % 6 for r < 36
      const int gy = cy0 - 1 + hy, gz = cz0 - 1 + hz;
      pre[j] = 0.f;
      if (okx && (unsigned)gy < (unsigned)H && (unsigned)gz < (unsigned)D) {
        pre[j] = __ldg(xn + (gz * HW + gy * W + gx));
        pre_ok |= 1u << j;
      }
    }
    pre[kRowsPerThread] = 0.f;
    if (threadIdx.x < 72) {
      const int hz = (e_r * 43) >> 8, hy = e_r - 6 * hz;
      const int ex = cx0 - 1 + e_hx, gy = cy0 - 1 + hy, gz = cz0 - 1 + hz;
      if ((unsigned)ex < (unsigned)W && (unsigned)gy < (unsigned)H && (unsigned)gz < (unsigned)D) {
        pre[kRowsPerThread] = __ldg(xn + (gz * HW + gy * W + ex));
        pre_ok |= 1u << kRowsPerThread;
      }
    }
    // advance to this CTA's next tile (mixed-radix add with carries)
    ftl += gridDim.x;
    ftx += stx;
    int c = 0;
    if (ftx >= tiles_x) { ftx -= tiles_x; c = 1; }
    fty += sty + c;
    c = 0;
    if (fty >= tiles_y) { fty -= tiles_y; c = 1; }
    ftz += stz + c;
  };

  const int wy = wid & 3, wz = (wid >> 2) * 2;   // the warp's row (y) and first z plane in the tile
  unsigned char* my_stage = stage[wid];
  // shared-memory byte addresses of this thread's A-fragment taps, relative to the tile's x group 0
  // and z plane wz (tile independent: the loops below only add immediates)
  uint32_t aaddr[4][2];
  {
    const uint32_t rowbase = (uint32_t)__cvta_generic_to_shared(tile) + (uint32_t)(((wz * kHY + wy) * kSX + g) * 4);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
      for (int j = 0; j < 2; ++j) aaddr[ks][j] = rowbase + (uint32_t)aoff[ks][j] * 4u;
  }
  float* st_main = tile + f_r0 * kSX + f_hx;
  float* st_extra = tile + e_r * kSX + e_hx;

  if ((int)blockIdx.x < ntiles) fetch();
  for (int tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
    const int x0 = cx0, y0 = cy0, z0 = cz0;
    __syncthreads();   // previous tile fully consumed
#pragma unroll
    for (int j = 0; j < kRowsPerThread; ++j) {
      const float v = ((pre_ok >> j) & 1u) ? fmaf(a_in, pre[j], b_in) : 0.f;
      st_main[4 * j * kSX] = __uint_as_float(to_tf32(v));
    }
    if (threadIdx.x < 72) {
      const float v = ((pre_ok >> kRowsPerThread) & 1u) ? fmaf(a_in, pre[kRowsPerThread], b_in) : 0.f;
      *st_extra = __uint_as_float(to_tf32(v));
    }
    __syncthreads();
    if (tl + (int)gridDim.x < ntiles) fetch();

    const int gy = y0 + wy;
    if (gy >= H) continue;   // warp-uniform; the loop-top barriers are still reached by everyone
    // two m-tiles (the warp's two z planes, same 16 x positions) per step: 2*NT independent
    // accumulator chains keep the tensor pipe busy with only 4 warps per scheduler.  The four x
    // groups are fully unrolled so that every LDS is [register + immediate].
    const bool xfull = x0 + kTX <= W;   // no per-voxel masks needed
#pragma unroll
    for (int xi = 0; xi < 4; ++xi) {
      const int xg = xi * 16;
      if (x0 + xg < W) {   // warp-uniform
      float acc[2][NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float2 b2 = *reinterpret_cast<const float2*>(cst + nt * 8 + 2 * t);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          acc[u][nt][0] = acc[u][nt][2] = b2.x;
          acc[u][nt][1] = acc[u][nt][3] = b2.y;
        }
      }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t a[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int imm = (u * kHY * kSX + xg) * 4;
          a[u][0] = lds_u32(aaddr[ks][0] + imm);        // constant offsets fold into the LDS immediate
          a[u][1] = lds_u32(aaddr[ks][0] + imm + 32);
          a[u][2] = lds_u32(aaddr[ks][1] + imm);
          a[u][3] = lds_u32(aaddr[ks][1] + imm + 32);
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int u = 0; u < 2; ++u)
            mma_tf32(acc[u][nt], a[u][0], a[u][1], a[u][2], a[u][3], bfrag[ks][nt][0], bfrag[ks][nt][1]);
      }
      // rows of this thread: voxel x = x0 + xg + g (acc[..][0..1]) and + 8 (acc[..][2..3])
      const bool ok0 = xfull || x0 + xg + g < W, ok1 = xfull || x0 + xg + g + 8 < W;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int gz = z0 + wz + u;
        if (gz < D) {   // warp-uniform
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (relu_pre) acc[u][nt][j] = fmaxf(acc[u][nt][j], 0.f);
        if (stats) {
          if (xfull) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const float v0 = acc[u][nt][j], v1 = acc[u][nt][2 + j];
                s[nt][j] += v0 + v1;
                ss[nt][j] = fmaf(v0, v0, fmaf(v1, v1, ss[nt][j]));
              }
          } else {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const float v0 = ok0 ? acc[u][nt][j] : 0.f, v1 = ok1 ? acc[u][nt][2 + j] : 0.f;
                s[nt][j] += v0 + v1;
                ss[nt][j] = fmaf(v0, v0, fmaf(v1, v1, ss[nt][j]));
              }
          }
        }
        if (on) {
          // normalise for the next layer, round to bf16, transpose through the warp's staging
          // buffer so that every lane stores 16 contiguous bytes (8 channels of one voxel)
          __syncwarp();
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const float2 sc2 = *reinterpret_cast<const float2*>(cst + COUT + nt * 8 + 2 * t);
            const float2 sh2 = *reinterpret_cast<const float2*>(cst + 2 * COUT + nt * 8 + 2 * t);
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              o[j] = fmaf((j & 1) ? sc2.y : sc2.x, acc[u][nt][j], (j & 1) ? sh2.y : sh2.x);
              if (relu_post) o[j] = fmaxf(o[j], 0.f);
            }
            *reinterpret_cast<uint32_t*>(my_stage + g * kStage + nt * 16 + t * 4) = km_pack2<F16>(o[0], o[1]);
            *reinterpret_cast<uint32_t*>(my_stage + (g + 8) * kStage + nt * 16 + t * 4) = km_pack2<F16>(o[2], o[3]);
          }
          __syncwarp();
          act16* dst = on + ((size_t)(gz * HW + gy * W + x0 + xg)) * COUT;
#pragma unroll
          for (int i = 0; i < NT / 2; ++i) {
            const int c = lane + 32 * i;            // 16-byte chunk id inside the 16-voxel segment
            const int v = c / NT, q = c % NT;
            if (xfull || x0 + xg + v < W) {
              const uint4 val = *reinterpret_cast<const uint4*>(my_stage + v * kStage + q * 16);
              *reinterpret_cast<uint4*>(dst + (size_t)v * COUT + q * 8) = val;
            }
          }
        }
        }
      }
      }
    }
  }

  if (stats) {
    // lanes with equal t hold the same channels: reduce over g (xor 4, 8, 16), fixed order
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          s[nt][j] += __shfl_xor_sync(0xffffffffu, s[nt][j], o);
          ss[nt][j] += __shfl_xor_sync(0xffffffffu, ss[nt][j], o);
        }
        if (g == 0) {
          const int c = nt * 8 + 2 * t + j;
          red[wid][2 * c] = s[nt][j];
          red[wid][2 * c + 1] = ss[nt][j];
        }
      }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * COUT; i += 256) {
      float a = 0.f;
      for (int wv = 0; wv < 8; ++wv) a += red[wv][i];
      stats[((size_t)blockIdx.x * N + n) * COUT * 2 + i] = a;
    }
  }
}

}  // namespace

extern "C" int km_stem_nparts(int, int, int, int) { return 148 * 2; }

extern "C" int km_conv3d_stem(const float* x, const float* w, const float* bias,
                              const float* in_scale, const float* in_shift, const float* out_scale,
                              const float* out_shift, void* out, float* stats, int N, int Cout,
                              int D, int H, int W, int relu_pre, int relu_post, km_stream_t stream) {
  KM_CHECK_ARG(x && w && N > 0 && D > 0 && H > 0 && W > 0, "km_conv3d_stem: bad arguments");
  KM_CHECK_ARG(out || stats, "km_conv3d_stem: neither an output nor statistics requested");
  KM_CHECK_ARG(Cout == 16 || Cout == 32, "km_conv3d_stem: Cout must be 16 or 32 (got %d)", Cout);
  KM_CHECK_ARG((out_scale == nullptr) == (out_shift == nullptr),
               "km_conv3d_stem: out_scale and out_shift go together");
  KM_CHECK_ARG((long long)D * H * W < (1ll << 31), "km_conv3d_stem: volume too large");
  const dim3 grid(km_stem_nparts(N, D, H, W), N);
  act16* o = reinterpret_cast<act16*>(out);
  const bool f16 = km_operand_fp16() != 0;
#define KM_STEM(C, F)                                                                                    \
  conv_stem_mma_kernel<C, F><<<grid, 256, 0, km_cs(stream)>>>(x, w, bias, in_scale, in_shift, out_scale, \
                                                              out_shift, o, stats, N, D, H, W, relu_pre, relu_post)
  if (Cout == 16) {
    if (f16) KM_STEM(16, true);
    else KM_STEM(16, false);
  } else {
    if (f16) KM_STEM(32, true);
    else KM_STEM(32, false);
  }
#undef KM_STEM
  KM_LAUNCH_OK("conv_stem_mma_kernel");
  return KM_OK;
}
