// Tiled trilinear warp: output tiles of 16 x 8 x 8 voxels gather from a shared-memory copy of their
// pre-image, staged by ONE TMA box load per (tile, channel).
//
// Reference call sites: keymorph/utils.py:14-21 (align_img = F.grid_sample(bilinear, border,
// align_corners=False)), keymorph/transformations.py:37-79 (affine flow field), keymorph/loss_ops.py:9-63.
//
// Why: the direct-gather kernels of warp.cu issue 8 global loads per voxel whose 32 lanes fall on ~5 cache
// lines each under a rotation (ncu, profiles/r02_ncu_warp_tps_kernels_baseline.txt: 15 sectors and 5.3
// wavefronts per request, long-scoreboard stalls, 1.9 TB/s = 0.29 of HBM for the fused affine warp; 3.1 GB
// of DRAM reads for 2.1 GB of one-hot channels).  Here the moving volume reaches the SM through the TMA
// unit: the tile's sampling coordinates are computed first, the minimum corner of their floor() indices is
// block-reduced, one thread issues cp.async.bulk.tensor for the 28 x 20 x 20 box at that origin (44.8 KB,
// zero-filled outside the volume, completion on an mbarrier), and the eight corners of every voxel are
// shared-memory loads.  A voxel whose corners do not fit the box (strong local magnification / shear) falls
// back to the direct global gather, so any transform stays correct.  With C channels the coordinates and
// weights are computed once and reused for every channel's box.  3 CTAs per SM hide the TMA latency.
//
// EXACT = true reproduces ATen's grid_sampler_3d arithmetic (un-fused multiplies and adds in ATen's corner
// order) for km_grid_sample3d; EXACT = false is tri_sample_fast of warp.cu (FMA chain), bit-identical to the
// direct-gather kernels it replaces in km_warp_loss.
#include <climits>
#include "km_common.cuh"
#include "tc_ptx.cuh"

using namespace kmtc;

namespace {

constexpr int kTX = 16, kTY = 8, kTZ = 8;             // output tile (x fastest)
constexpr int kBX = 28, kBY = 20, kBZ = 20;           // pre-image box in the moving volume; its x origin is rounded
                                                      // down to a multiple of 4: the innermost TMA start
                                                      // coordinate must be 16-byte aligned (measured: any other
                                                      // value raises "illegal instruction", tools/probes/)
constexpr int kBoxBytes = kBX * kBY * kBZ * 4;        // 44800
constexpr int kThreads = 256;
constexpr int kVox = 4;                               // voxels per thread: z = (tid >> 7) + 2 k
constexpr int kMaxC = 32;                             // channels with per-warp loss slots in shared memory

int g_warp_tile = 1;   // km_set_option(KM_OPT_WARP_TILE)

__device__ __forceinline__ float src_index(float g, int size) {
  float v = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.f), (float)size), 1.f), 0.5f);
  return fminf((float)(size - 1), fmaxf(v, 0.f));
}

struct TileArgs {
  const float* mat;        // COORD affine: (N,3,4) rows in (z,y,x)
  const float* grid;       // COORD grid: (N,Do,Ho,Wo,3)
  const float* moving;     // (N,C,Di,Hi,Wi): fallback gathers
  const float* fixed;      // (N,C,Do,Ho,Wo) or null
  float* out;              // (N,C,Do,Ho,Wo) or null
  float* grid_out;         // (N,Do,Ho,Wo,3) or null (affine only)
  float* partials;         // [gridDim.x][N][C][4] or null
  int N, C, Di, Hi, Wi, Do, Ho, Wo;
  int tiles_x, tiles_y, tiles_z;
};

template <bool AFFINE, bool EXACT>
__global__ void __launch_bounds__(kThreads, 3)
warp_tile_kernel(const __grid_constant__ CUtensorMap tmM, const TileArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  // the TMA destination wants 128-byte alignment: align by hand (128 spare bytes are allocated)
  unsigned char* smem_raw = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
  float* box = reinterpret_cast<float*>(smem_raw);                       // [kBZ][kBY][kBX]
  float* s_stage = reinterpret_cast<float*>(smem_raw + kBoxBytes);       // [8 warps][96]
  float* s_red = s_stage + 8 * 96;                                       // [8 warps][kMaxC * 4]
  // per-voxel state (ix, iy, iz, box offset) lives in shared memory, so the per-voxel loops stay rolled and
  // the kernel fits 64 registers (4 CTAs / SM); one conflict-free LDS.128 per voxel and channel
  float4* s_vox = reinterpret_cast<float4*>(s_red + 8 * kMaxC * 4) + threadIdx.x;   // [kVox][kThreads]
  __shared__ int s_min[2][3];
  __shared__ __align__(8) unsigned long long s_bar;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = blockIdx.y;
  const int lx = tid & 15, ly = (tid >> 4) & 7, lz0 = tid >> 7;
  const uint32_t bar = smem_u32(&s_bar), box_u32 = smem_u32(box);
  const bool multi = a.C > 1;
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    for (int i = 0; i < 3; ++i) s_min[0][i] = s_min[1][i] = INT_MAX;
    prefetch_tmap(&tmM);
  }
  if (multi && a.partials)
    for (int i = lane; i < a.C * 4; i += 32) s_red[wid * kMaxC * 4 + i] = 0.f;
  __syncthreads();

  float m[12];
  if (AFFINE) {
#pragma unroll
    for (int i = 0; i < 12; ++i) m[i] = __ldg(a.mat + n * 12 + i);
  }
  const size_t in_vol = (size_t)a.Di * a.Hi * a.Wi, out_vol = (size_t)a.Do * a.Ho * a.Wo;
  const float* gn = AFFINE ? nullptr : a.grid + (size_t)n * out_vol * 3;
  float* gon = a.grid_out ? a.grid_out + (size_t)n * out_vol * 3 : nullptr;
  const int HiWi = a.Hi * a.Wi;
  float acc1[4] = {0.f, 0.f, 0.f, 0.f};
  uint32_t loads = 0;
  const int ntiles = a.tiles_x * a.tiles_y * a.tiles_z;
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int tx = tile % a.tiles_x, ty = (tile / a.tiles_x) % a.tiles_y, tz = tile / (a.tiles_x * a.tiles_y);
    const int x = tx * kTX + lx, y = ty * kTY + ly;
    const int pb = it & 1;
    const int vo0 = ((tz * kTZ + lz0) * a.Ho + y) * a.Wo + x, vo_step = 2 * a.Ho * a.Wo;
    // per voxel only the clamped source coordinates and (after the origin is known) the box offset stay in
    // registers; floor(), weights and corner indices are recomputed per channel (a few ALU ops next to 8 loads)
    // off >= 0: offset of corner (x0,y0,z0) inside the box; -1: direct global gather; -2: no voxel
    int mnx = INT_MAX, mny = INT_MAX, mnz = INT_MAX;
#pragma unroll 1
    for (int k = 0; k < kVox; ++k) {
      const int z = tz * kTZ + lz0 + 2 * k;
      const bool ok = x < a.Wo && y < a.Ho && z < a.Do;
      float gx = 0.f, gy = 0.f, gz = 0.f;
      if (AFFINE) {
        const float pz = km_linspace(-1.f, 1.f, a.Do, min(z, a.Do - 1)), py = km_linspace(-1.f, 1.f, a.Ho, min(y, a.Ho - 1));
        const float px = km_linspace(-1.f, 1.f, a.Wo, min(x, a.Wo - 1));
        gz = fmaf(m[0], pz, fmaf(m[1], py, fmaf(m[2], px, m[3])));
        gy = fmaf(m[4], pz, fmaf(m[5], py, fmaf(m[6], px, m[7])));
        gx = fmaf(m[8], pz, fmaf(m[9], py, fmaf(m[10], px, m[11])));
        if (gon) {
          // the 16 (x,y,z) triples of a tile row are 192 contiguous bytes: staged per warp (2 rows) and
          // written as 16-byte pieces by 24 lanes
          float* sg = s_stage + wid * 96;
          __syncwarp();
          sg[(lane >> 4) * 48 + lx * 3 + 0] = gx;
          sg[(lane >> 4) * 48 + lx * 3 + 1] = gy;
          sg[(lane >> 4) * 48 + lx * 3 + 2] = gz;
          __syncwarp();
          const bool row_full = tx * kTX + kTX <= a.Wo;
          if (row_full) {
            if (lane < 24) {
              const int row = lane / 12, q = lane - row * 12;
              const int yr = ty * kTY + (ly & ~1) + row;
              if (yr < a.Ho && z < a.Do) {
                const float4 v = *reinterpret_cast<const float4*>(sg + row * 48 + q * 4);
                *reinterpret_cast<float4*>(gon + (((size_t)z * a.Ho + yr) * a.Wo + tx * kTX) * 3 + q * 4) = v;
              }
            }
          } else if (ok) {
            float* gp = gon + (((size_t)z * a.Ho + y) * a.Wo + x) * 3;
            gp[0] = gx;
            gp[1] = gy;
            gp[2] = gz;
          }
        }
      } else if (ok) {
        const float* gp = gn + (((size_t)z * a.Ho + y) * a.Wo + x) * 3;
        gx = __ldg(gp);
        gy = __ldg(gp + 1);
        gz = __ldg(gp + 2);
      }
      const float sx = src_index(gx, a.Wi), sy = src_index(gy, a.Hi), sz = src_index(gz, a.Di);
      s_vox[k * kThreads] = make_float4(sx, sy, sz, __int_as_float(ok ? -1 : -2));
      if (ok) {
        mnx = min(mnx, (int)floorf(sx));
        mny = min(mny, (int)floorf(sy));
        mnz = min(mnz, (int)floorf(sz));
      }
    }
    mnx = __reduce_min_sync(0xffffffffu, mnx);
    mny = __reduce_min_sync(0xffffffffu, mny);
    mnz = __reduce_min_sync(0xffffffffu, mnz);
    if (lane == 0) {
      atomicMin(&s_min[pb][0], mnx);
      atomicMin(&s_min[pb][1], mny);
      atomicMin(&s_min[pb][2], mnz);
    }
    __syncthreads();   // origin complete; every thread has finished reading the previous tile's box
    const int bx = s_min[pb][0] & ~3, by = s_min[pb][1], bz = s_min[pb][2];
#pragma unroll 1
    for (int k = 0; k < kVox; ++k) {
      float4 sv = s_vox[k * kThreads];
      const int ox = (int)floorf(sv.x) - bx, oy = (int)floorf(sv.y) - by, oz = (int)floorf(sv.z) - bz;
      if (__float_as_int(sv.w) == -1 && ox + 1 < kBX && oy + 1 < kBY && oz + 1 < kBZ) {
        sv.w = __int_as_float((oz * kBY + oy) * kBX + ox);
        s_vox[k * kThreads] = sv;
      }
    }
    for (int c = 0; c < a.C; ++c) {
      const size_t ch = (size_t)n * a.C + c;
      if (c > 0) __syncthreads();   // the previous channel's box has been consumed
      if (tid == 0) {
        if (c == 0) {
          s_min[pb ^ 1][0] = INT_MAX;
          s_min[pb ^ 1][1] = INT_MAX;
          s_min[pb ^ 1][2] = INT_MAX;
        }
        mbar_arrive_expect_tx(bar, kBoxBytes);
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(box_u32), "l"(reinterpret_cast<uint64_t>(&tmM)), "r"(bar), "r"(bx), "r"(by), "r"(bz), "r"((int)ch)
            : "memory");
      }
      mbar_wait(bar, loads & 1u);
      ++loads;
      const float* vol = a.moving + ch * in_vol;
      const float* fix_c = a.fixed ? a.fixed + ch * out_vol : nullptr;
      float* out_c = a.out ? a.out + ch * out_vol : nullptr;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int k = 0; k < kVox; ++k) {
        const float4 sv = s_vox[k * kThreads];
        const int offk = __float_as_int(sv.w);
        if (offk == -2) continue;
        const float ixk = sv.x, iyk = sv.y, izk = sv.z;
        const float x0f = floorf(ixk), y0f = floorf(iyk), z0f = floorf(izk);
        float v[8];
        if (offk >= 0) {
          const float* p = box + offk;
          v[0] = p[0]; v[1] = p[1]; v[2] = p[kBX]; v[3] = p[kBX + 1];
          v[4] = p[kBY * kBX]; v[5] = p[kBY * kBX + 1]; v[6] = p[kBY * kBX + kBX]; v[7] = p[kBY * kBX + kBX + 1];
        } else {
          // direct gather; clamped corners carry weight exactly 0 (EXACT: ATen skips them, adding 0 is the same)
          const int x0 = (int)x0f, y0 = (int)y0f, z0 = (int)z0f;
          const int x1 = min(x0 + 1, a.Wi - 1), y1 = min(y0 + 1, a.Hi - 1), z1 = min(z0 + 1, a.Di - 1);
          const int r00 = z0 * HiWi + y0 * a.Wi, r01 = z0 * HiWi + y1 * a.Wi;
          const int r10 = z1 * HiWi + y0 * a.Wi, r11 = z1 * HiWi + y1 * a.Wi;
          v[0] = __ldg(vol + r00 + x0); v[1] = __ldg(vol + r00 + x1);
          v[2] = __ldg(vol + r01 + x0); v[3] = __ldg(vol + r01 + x1);
          v[4] = __ldg(vol + r10 + x0); v[5] = __ldg(vol + r10 + x1);
          v[6] = __ldg(vol + r11 + x0); v[7] = __ldg(vol + r11 + x1);
          if (EXACT) {   // a clamped (out-of-volume) corner must contribute +0 even when the sample is not finite
            if (x0 + 1 >= a.Wi) v[1] = v[3] = v[5] = v[7] = 0.f;
            if (y0 + 1 >= a.Hi) v[2] = v[3] = v[6] = v[7] = 0.f;
            if (z0 + 1 >= a.Di) v[4] = v[5] = v[6] = v[7] = 0.f;
          }
        }
        float r;
        if (EXACT) {
          // ATen grid_sampler_3d: weight = (x * y) * z and out += value * weight, un-fused, corner order
          // tnw, tne, tsw, tse, bnw, bne, bsw, bse
          const float wx1 = __fsub_rn(ixk, x0f), wx0 = __fsub_rn(x0f + 1.f, ixk);
          const float wy1 = __fsub_rn(iyk, y0f), wy0 = __fsub_rn(y0f + 1.f, iyk);
          const float wz1 = __fsub_rn(izk, z0f), wz0 = __fsub_rn(z0f + 1.f, izk);
          r = __fmul_rn(v[0], __fmul_rn(__fmul_rn(wx0, wy0), wz0));
          r = __fadd_rn(r, __fmul_rn(v[1], __fmul_rn(__fmul_rn(wx1, wy0), wz0)));
          r = __fadd_rn(r, __fmul_rn(v[2], __fmul_rn(__fmul_rn(wx0, wy1), wz0)));
          r = __fadd_rn(r, __fmul_rn(v[3], __fmul_rn(__fmul_rn(wx1, wy1), wz0)));
          r = __fadd_rn(r, __fmul_rn(v[4], __fmul_rn(__fmul_rn(wx0, wy0), wz1)));
          r = __fadd_rn(r, __fmul_rn(v[5], __fmul_rn(__fmul_rn(wx1, wy0), wz1)));
          r = __fadd_rn(r, __fmul_rn(v[6], __fmul_rn(__fmul_rn(wx0, wy1), wz1)));
          r = __fadd_rn(r, __fmul_rn(v[7], __fmul_rn(__fmul_rn(wx1, wy1), wz1)));
        } else {
          // tri_sample_fast of warp.cu, same operation order
          const float wx1 = ixk - x0f, wx0 = (x0f + 1.f) - ixk;
          const float wy1 = iyk - y0f, wy0 = (y0f + 1.f) - iyk;
          const float wz1 = izk - z0f, wz0 = (z0f + 1.f) - izk;
          const float w00 = wx0 * wy0, w10 = wx1 * wy0, w01 = wx0 * wy1, w11 = wx1 * wy1;
          r = v[0] * (w00 * wz0);
          r = fmaf(v[1], w10 * wz0, r);
          r = fmaf(v[2], w01 * wz0, r);
          r = fmaf(v[3], w11 * wz0, r);
          r = fmaf(v[4], w00 * wz1, r);
          r = fmaf(v[5], w10 * wz1, r);
          r = fmaf(v[6], w01 * wz1, r);
          r = fmaf(v[7], w11 * wz1, r);
        }
        const int idx = vo0 + k * vo_step;     // voxel offset inside the channel (< 2^31, checked on the host)
        if (out_c) out_c[idx] = r;
        if (fix_c) {
          const float fv = __ldg(fix_c + idx);
          const float d = r - fv;
          acc[0] = fmaf(d, d, acc[0]);
          acc[1] = fmaf(r, fv, acc[1]);
          acc[2] = fmaf(r, r, acc[2]);
          acc[3] = fmaf(fv, fv, acc[3]);
        }
      }
      if (a.fixed) {
        if (multi) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float s = km_warp_sum(acc[j]);
            if (lane == 0) s_red[wid * kMaxC * 4 + c * 4 + j] += s;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) acc1[j] += acc[j];
        }
      }
    }
  }
  if (a.partials) {
    if (!multi) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float s = km_warp_sum(acc1[j]);
        if (lane == 0) s_red[wid * kMaxC * 4 + j] = s;
      }
    }
    __syncthreads();
    for (int i = tid; i < a.C * 4; i += kThreads) {
      float s = 0.f;
      for (int wv = 0; wv < 8; ++wv) s += s_red[wv * kMaxC * 4 + i];
      a.partials[((size_t)blockIdx.x * a.N + n) * a.C * 4 + i] = s;
    }
  }
}

constexpr int kSmemBytes = 128 + kBoxBytes + 8 * 96 * 4 + 8 * kMaxC * 4 * 4 + kVox * kThreads * 16;   // box + grid staging + loss slots + voxel state

}  // namespace

void km_warp_set_tile(int v) { g_warp_tile = v ? 1 : 0; }

// can the tiled kernel take this call?  (bilinear only; W % 4: TMA strides are multiples of 16 bytes)
bool km_warp_tile_eligible(const float* moving, int C, int Di, int Hi, int Wi, int Do, int Ho, int Wo) {
  return g_warp_tile && C <= kMaxC && (Wi % 4) == 0 && (((uintptr_t)moving) & 15) == 0 &&
         (long long)Di * Hi * Wi < (1ll << 31) && (long long)Do * Ho * Wo < (1ll << 31) && Wi >= 4 &&
         (long long)Do * Ho * Wo >= 4096;
}

// coord: KM_COORD_AFFINE (mat) or KM_COORD_GRID (grid).  partials: [grid_x][N][C][4] floats with
// grid_x = KM_RED_BLOCKS (the caller reduces them), or null.
int km_warp_tile_launch(int coord, bool exact, const float* mat, const float* grid, const float* moving,
                        const float* fixed, float* out, float* grid_out, float* partials, int N, int C, int Di,
                        int Hi, int Wi, int Do, int Ho, int Wo, cudaStream_t st) {
  PFN_encodeTiled encode = tensor_map_encoder();
  if (!encode) {
    km_set_error("warp_tile: cuTensorMapEncodeTiled unavailable");
    return KM_ECUDA;
  }
  CUtensorMap tmM;
  cuuint64_t dims[4] = {(cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)Di, (cuuint64_t)N * C};
  cuuint64_t strides[3] = {(cuuint64_t)Wi * 4, (cuuint64_t)Hi * Wi * 4, (cuuint64_t)Di * Hi * Wi * 4};
  cuuint32_t bdim[4] = {kBX, kBY, kBZ, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode(&tmM, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(moving), dims, strides, bdim, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    km_set_error("warp_tile: cuTensorMapEncodeTiled failed with %d", (int)r);
    return KM_ECUDA;
  }
  TileArgs a;
  a.mat = mat; a.grid = grid; a.moving = moving; a.fixed = fixed; a.out = out; a.grid_out = grid_out;
  a.partials = partials;
  a.N = N; a.C = C; a.Di = Di; a.Hi = Hi; a.Wi = Wi; a.Do = Do; a.Ho = Ho; a.Wo = Wo;
  a.tiles_x = (Wo + kTX - 1) / kTX; a.tiles_y = (Ho + kTY - 1) / kTY; a.tiles_z = (Do + kTZ - 1) / kTZ;
  const dim3 g(KM_RED_BLOCKS, N);
  static unsigned long long attr_set = 0;
  if (km_first_use_on_device(&attr_set)) {
    KM_CUDA_OK(cudaFuncSetAttribute(warp_tile_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    KM_CUDA_OK(cudaFuncSetAttribute(warp_tile_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    KM_CUDA_OK(cudaFuncSetAttribute(warp_tile_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  }
  if (coord == KM_COORD_AFFINE)
    warp_tile_kernel<true, false><<<g, kThreads, kSmemBytes, st>>>(tmM, a);
  else if (exact)
    warp_tile_kernel<false, true><<<g, kThreads, kSmemBytes, st>>>(tmM, a);
  else
    warp_tile_kernel<false, false><<<g, kThreads, kSmemBytes, st>>>(tmM, a);
  KM_LAUNCH_OK("warp_tile_kernel");
  return KM_OK;
}
