// remap_lean2_kernel: the round-2 headline kernel -- bilinear, TMA-staged, planar 1 / 3 / 4
// channels, equi2equi / equi2pers / equi2cube -- built around three measurements of the
// round-1 lean kernel (profiles/r01_ncu_remap_bf16_v6_lean_final.txt: 147 instructions per
// pixel, L1 data pipe 82 %, 2.3 wavefronts per 2-byte tap load):
//
//  1. COORDINATES FROM A STRIP-WIDE NODE GRID.  A CTA walks a strip of up to 8 tiles
//     (64 x 8 pixels each) along x.  Before the first tile all 128 threads evaluate the
//     exact coordinate chain at the strip's NODES -- every 8th column of every row, plus one
//     column on either side -- into shared memory (fully occupied warps: 1/8 node per pixel
//     instead of the 1/4 of round 1, where every lane computed its own).  A pixel then gets
//     (ui, uj) from the four nodes around it on its own row by cubic Lagrange interpolation
//     with weights that are LANE CONSTANTS (a lane's pixels all sit at the same offset inside
//     their 8-pixel interval), as differences from the nearest node: 4 LDS.64 + 3 packed
//     subtractions + 3 FFMA2 per pixel, no shuffles, no weight loads.
//  2. PER-TILE HEADERS.  The same prologue reduces each tile's node hull to its TMA box
//     (origin, shape) and to three flags -- smooth (4th and 2nd differences of the nodes
//     small: the interpolation error is below 1e-3 px and no pixel leaves the node hull by
//     more than 1/4 px, so no per-pixel "inside the box" test is needed), interior (no clamp,
//     no wrap, no footprint shift can trigger) and straddling the seam.  The tile loop then
//     runs without bounding-box reductions, votes, clamps or wraps on its common path.
//  3. ONE BANK PER LANE.  A lane owns four pixels 16 apart (2-byte and 4-byte types) or two
//     pairs (uint8), so the 16 lanes of a row read 16 ADJACENT source pixels per tap
//     instruction: 32 (64) contiguous bytes instead of the 128 strided bytes of 4-pixel
//     lanes, and box pitches are chosen so that the warp's second lane row lands in the
//     other half of the banks (pitch * elt = 64 mod 128).
//
// Synchronisation: no __syncthreads in the tile loop.  The one box buffer is handed over with
// two mbarriers: `full` (TMA completion) and `empty` (one arrival per warp when it is done
// reading the tile); thread 0 waits for `empty` of the previous tile before it issues the
// next box, every other thread runs ahead into the next tile's coordinate arithmetic.
//
// Tiles that are not smooth (source poles), ragged tiles and the EXACT variant evaluate the
// exact chain on every pixel; tiles whose footprint fits no box (poles, seam) gather from
// global memory -- same values on every path (tests: test_staged_and_direct_kernels_agree).
#pragma once

#include "remap_kernels.cuh"

namespace eqb {

constexpr int kL2Warps = 4;
constexpr int kL2TileW = 64, kL2TileH = 8;
constexpr int kL2Step = 8;                         // node spacing along x
constexpr int kL2MaxIters = 8;                     // tiles per strip
constexpr int kL2Cols = kL2MaxIters * 8 + 3;       // node columns -1 .. 8 * iters + 1
constexpr int kL2Boxes = 13;
static_assert(kL2Boxes <= kTmaSlots, "tensor-map slots");

// Box buffer per CTA.  1- and 2-byte pixels: 22.5 KB, so that buffer + node grid + headers stay
// under the 27.5 KB that let eight CTAs share an SM.
EQB_HD constexpr int l2_smem_bytes(int elt) { return elt == 4 ? 32 * 1024 : 23040; }

// Box m: l2_box_w x l2_box_h source pixels x channels.  1- and 2-byte pixels: pitches of 96 /
// 160 / 224 pixels put consecutive rows 64 (mod 128) bytes apart for 2-byte pixels and 32 / 96
// for bytes; 4-byte pixels use 80 / 112 / 144 / 176 (64 mod 128).  Boxes 0-2 are small
// (upright and mildly rolled views: less L2 -> shared traffic); the others spend the whole
// budget in different aspect ratios (see box_w in remap_kernels.cuh for the reasoning).
EQB_HD constexpr int l2_box_w(int elt, int m) {
  const int w2 = m <= 3 ? 96 : m == 4 ? 160 : m == 5 ? 224 : m == 6 ? 128 : m == 7 ? 72 :
                 m == 8 ? 64 : m == 9 ? 56 : m == 10 ? 48 : m == 11 ? 256 : 192;
  const int w4 = m <= 3 ? 80 : m == 4 ? 112 : m == 5 ? 144 : m == 6 ? 176 : m == 7 ? 48 :
                 m == 8 ? 64 : m == 9 ? 96 : m == 10 ? 208 : m == 11 ? 240 : 128;
  const int w = elt == 4 ? w4 : w2;
  const int per16 = 16 / elt;
  return (w + per16 - 1) / per16 * per16;
}
EQB_HD constexpr int l2_box_h(int elt, int nc, int m) {
  const int budget = l2_smem_bytes(elt) / elt / nc;
  const int full = budget / l2_box_w(elt, m) < 128 ? budget / l2_box_w(elt, m) : 128;
  const int small = m == 0 ? 16 : m == 1 ? 24 : 32;
  return m < 3 && small < full ? small : full;
}

// Cubic Lagrange weights of the nodes at -8, 0, +8, +16 for the pixel at offset t (0..7)
// from the second one; w[t][1] is not stored (the weights sum to 1: u = n1 + sum w_j (n_j - n1)).
struct L2Weights { float w[8][3]; };
constexpr L2Weights make_l2_weights() {
  L2Weights r{};
  for (int t = 0; t < 8; ++t) {
    const double s = t / 8.0;
    r.w[t][0] = (float)(-s * (s - 1.0) * (s - 2.0) / 6.0);
    r.w[t][1] = (float)(-(s + 1.0) * s * (s - 2.0) / 2.0);
    r.w[t][2] = (float)((s + 1.0) * s * (s - 1.0) / 6.0);
  }
  return r;
}
static __device__ const L2Weights g_l2_weights = make_l2_weights();

enum : int { L2_SMOOTH = 1, L2_BORDER = 2, L2_STRADDLE = 4, L2_RAGGED = 8, L2_BACKGROUND = 16 };

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Rays of the cube faces as coded in torch_utils/grid.py:137-171 (F R B L U D).
__device__ __forceinline__ void cube_ray(int face, float rc, float rr, float& mx, float& my,
                                         float& mz) {
  switch (face) {
    case 0: mx = 0.5f; my = rc; mz = -rr; break;
    case 1: mx = -rc; my = 0.5f; mz = -rr; break;
    case 2: mx = -0.5f; my = -rc; mz = -rr; break;
    case 3: mx = rc; my = -0.5f; mz = -rr; break;
    case 4: mx = rr; my = rc; mz = 0.5f; break;
    default: mx = -rr; my = rc; mz = -0.5f; break;
  }
}

// Coordinates of a NODE: column x may lie up to 8 pixels outside the image (the map is
// continued smoothly: longitude is periodic, the pinhole grid and the cube-face ramp are
// linear).  Nodes only feed the interpolation and the boxes, never a pixel directly -- except
// where x is a pixel column, and there this is the exact chain.
template <int XF>
__device__ __forceinline__ void node_coords(const Params& p, const RowCtx<XF>& c, int x, int y,
                                            int cell, float& uj, float& ui) {
  if (XF == EQB_EQUI2EQUI) {
    const int xw = x < 0 ? x + p.dw : (x >= p.dw ? x - p.dw : x);
    coords<XF, false>(p, c, 0, xw, y, 0, 0, uj, ui);
  } else if (XF == EQB_EQUI2PERS) {
    coords<XF, false>(p, c, 0, x, y, 0, 0, uj, ui);
  } else {  // EQB_EQUI2CUBE
    const int w = p.w_face, lx = x - cell * w;
    float rcol;
    if (lx >= 0 && lx < w) {
      rcol = __ldg(p.tx + lx);
    } else {
      const float step = __ldg(p.tx + 1) - __ldg(p.tx);
      rcol = lx < 0 ? fmaf((float)lx, step, __ldg(p.tx))
                    : fmaf((float)(lx - (w - 1)), step, __ldg(p.tx + w - 1));
    }
    float mx, my, mz, X, Y, Z, phi, theta;
    cube_ray(cell, rcol, c.r0, mx, my, mz);
    rotate(c.A, mx, my, mz, X, Y, Z);
    spherical<false>(X, Y, Z, phi, theta);
    tail_minus_half<false>(phi, theta, p.shf, p.swf, uj, ui);
  }
}

// One tap of channel plane offset OFF (bytes, compile time) as fp32.
template <typename T, int OFF> __device__ __forceinline__ float l2_tap(unsigned addr) {
  if (sizeof(T) == 4) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));
    return v;
  }
  if (sizeof(T) == 1) {
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
    // 2^23 + b, minus 2^23: full-rate pipes instead of the conversion unit
    return __fsub_rn(__uint_as_float(v | 0x4B000000u), 8388608.0f);
  }
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1+%2];" : "=h"(v) : "r"(addr), "n"(OFF));
  if (!cuda::std::is_same<T, __half>::value) return __uint_as_float((unsigned)v << 16);  // bf16
  return __half2float(__ushort_as_half(v));
}

// Store of a lane's results for one channel.  P = 1: four pixels 16 apart; P = 2 (uint8): two
// pairs 32 apart.  `d` points at the lane's first pixel of this channel.
template <typename T>
__device__ __forceinline__ void l2_store(T* d, const float (&o)[4], unsigned flags) {
  if constexpr (sizeof(T) == 4) {
#pragma unroll
    for (int g = 0; g < 4; ++g) __stcs(reinterpret_cast<float*>(d) + 16 * g, o[g]);
  } else if constexpr (sizeof(T) == 2) {
    unsigned lo, hi;
    if constexpr (cuda::std::is_same<T, __half>::value) {
      const __half2 a = __floats2half2_rn(o[0], o[1]), b = __floats2half2_rn(o[2], o[3]);
      lo = *reinterpret_cast<const unsigned*>(&a);
      hi = *reinterpret_cast<const unsigned*>(&b);
    } else {
      const __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]);
      const __nv_bfloat162 b = __floats2bfloat162_rn(o[2], o[3]);
      lo = *reinterpret_cast<const unsigned*>(&a);
      hi = *reinterpret_cast<const unsigned*>(&b);
    }
    unsigned short* q = reinterpret_cast<unsigned short*>(d);
    __stcs(q, (unsigned short)lo);
    __stcs(q + 16, (unsigned short)(lo >> 16));
    __stcs(q + 32, (unsigned short)hi);
    __stcs(q + 48, (unsigned short)(hi >> 16));
  } else {
    // bilinear code values are convex combinations: floor == the reference's truncation
    const unsigned t0 = u8_floor_bits(o[0]), t1 = u8_floor_bits(o[1]);
    const unsigned t2 = u8_floor_bits(o[2]), t3 = u8_floor_bits(o[3]);
    unsigned short* q = reinterpret_cast<unsigned short*>(d);
    __stcs(q, (unsigned short)__byte_perm(t0, t1, 0x0040u));
    __stcs(q + 16, (unsigned short)__byte_perm(t2, t3, 0x0040u));
  }
}

// Blend + store of one channel from box M: pixel pairs (0, 1) and (2, 3) share FFMA2s.
template <typename T, int M, int NC, int CI>
__device__ __forceinline__ void l2_blend_channel(const unsigned (&a)[4],
                                                 const unsigned long long (&w_nw)[2],
                                                 const unsigned long long (&w_ne)[2],
                                                 const unsigned long long (&w_sw)[2],
                                                 const unsigned long long (&w_se)[2], T* dst,
                                                 long long dsc, unsigned flags) {
  constexpr int ELT = (int)sizeof(T);
  constexpr int ROW = l2_box_w(ELT, M) * ELT;
  constexpr int CH = CI * ROW * l2_box_h(ELT, NC, M);
  float o[4];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const unsigned a0 = a[2 * k], a1 = a[2 * k + 1];
    unsigned long long acc = mul2(pk2(l2_tap<T, CH>(a0), l2_tap<T, CH>(a1)), w_nw[k]);
    acc = fma2(pk2(l2_tap<T, CH + ELT>(a0), l2_tap<T, CH + ELT>(a1)), w_ne[k], acc);
    acc = fma2(pk2(l2_tap<T, CH + ROW>(a0), l2_tap<T, CH + ROW>(a1)), w_sw[k], acc);
    acc = fma2(pk2(l2_tap<T, CH + ROW + ELT>(a0), l2_tap<T, CH + ROW + ELT>(a1)), w_se[k], acc);
    unpk2(acc, o[2 * k], o[2 * k + 1]);
  }
  l2_store<T>(dst + CI * dsc, o, flags);
}

template <typename T, int M, int NC>
__device__ __forceinline__ void l2_blend(const unsigned (&a)[4], const float (&fx)[4],
                                         const float (&fy)[4], T* dst, long long dsc,
                                         unsigned flags) {
  const unsigned long long one = pk2(1.0f, 1.0f), mone = pk2(-1.0f, -1.0f);
  unsigned long long w_nw[2], w_ne[2], w_sw[2], w_se[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const unsigned long long bx = pk2(fx[2 * k], fx[2 * k + 1]), by = pk2(fy[2 * k], fy[2 * k + 1]);
    const unsigned long long ax = fma2(bx, mone, one), ay = fma2(by, mone, one);  // 1 - b, exact
    w_nw[k] = mul2(ax, ay);
    w_ne[k] = mul2(bx, ay);
    w_sw[k] = mul2(ax, by);
    w_se[k] = mul2(bx, by);
  }
  l2_blend_channel<T, M, NC, 0>(a, w_nw, w_ne, w_sw, w_se, dst, dsc, flags);
  if constexpr (NC > 1) l2_blend_channel<T, M, NC, 1>(a, w_nw, w_ne, w_sw, w_se, dst, dsc, flags);
  if constexpr (NC > 2) l2_blend_channel<T, M, NC, 2>(a, w_nw, w_ne, w_sw, w_se, dst, dsc, flags);
  if constexpr (NC > 3) l2_blend_channel<T, M, NC, 3>(a, w_nw, w_ne, w_sw, w_se, dst, dsc, flags);
}

template <int XF, typename T, int NC, bool EXACT>
__global__ void __launch_bounds__(kL2Warps * 32, EXACT || sizeof(T) == 4 ? 6 : 8)
    remap_lean2_kernel(const __grid_constant__ Params p) {
  constexpr int ELT = (int)sizeof(T);
  constexpr int PER16 = 16 / ELT;
  constexpr int P = ELT == 1 ? 2 : 1;  // adjacent pixels per group: a group is >= 2 bytes
  constexpr int G = 4 / P;             // groups per lane, 16 * P pixels apart
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(16) float2 nodes[kL2TileH][kL2Cols + 1];  // (ui, uj)
  __shared__ __align__(16) int4 headers[kL2MaxIters];
  __shared__ __align__(8) unsigned long long full_bar, empty_bar;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.z;
  const int tile_y0 = blockIdx.y * kL2TileH;
  const int strip_x0 = blockIdx.x * p.iters * kL2TileW;
  const int ntiles = min(p.iters, (p.dw - strip_x0 + kL2TileW - 1) / kL2TileW);
  int cell = 0;
  if (XF == EQB_EQUI2CUBE) cell = strip_x0 / p.w_face;  // (a strip never straddles two cells)
  const bool background = XF == EQB_EQUI2CUBE && cell >= 6;

  if (tid == 0) {
    mbar_init(&full_bar, 1);
    mbar_init(&empty_bar, kL2Warps);
  }

  // ---- prologue 1: the exact chain at the strip's nodes -------------------------------
  if (!background) {
    const int row = tid & (kL2TileH - 1);
    const int y = min(tile_y0 + row, p.dh - 1);
    RowCtx<XF> rc;
    row_setup<XF>(p, b, y, rc);
    const int ncols = ntiles * 8 + 3;
    for (int col = tid >> 3; col < ncols; col += kL2Warps * 32 / kL2TileH) {
      float uj, ui;
      node_coords<XF>(p, rc, strip_x0 + kL2Step * (col - 1), y, cell, uj, ui);
      nodes[row][col] = make_float2(ui, uj);
    }
  }
  __syncthreads();

  // ---- prologue 2: one header per tile (warp w: tiles w, w + 4) -------------------------
  {
    // lane m holds the shape of box m
    const int my_box = lane < kL2Boxes ? (p.tma_bw[lane & 15] | (p.tma_bh[lane & 15] << 16)) : 0;
    for (int k = warp; k < ntiles; k += kL2Warps) {
      int flags = 0;
      if (strip_x0 + (k + 1) * kL2TileW > p.dw || tile_y0 + kL2TileH > p.dh) flags |= L2_RAGGED;
      if (background) {
        if (lane == 0) headers[k] = make_int4(0, 0, -1, flags | L2_BACKGROUND);
        continue;
      }
      // window of five nodes centred on tile-local node column c = 1, 3, 5, 7 of row r
      const int r = lane & 7, c = 1 + 2 * (lane >> 3);
      const float2* np = &nodes[r][k * 8 + c - 1];  // (array index = node column + 1)
      float2 n[5];
#pragma unroll
      for (int i = 0; i < 5; ++i) n[i] = np[i];
      // hull of the tile's own node columns 0..8 (the windows also see columns -1 and 9,
      // which only matter for the seam: `alo / ahi` span all five)
      float ulo = n[1].x, uhi = n[1].x, vlo = n[1].y, vhi = n[1].y;
      float alo = fminf(n[0].x, n[4].x), ahi = fmaxf(n[0].x, n[4].x);
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        if ((i == 0 && c == 1) || (i == 4 && c == 7) || i == 1) continue;
        ulo = fminf(ulo, n[i].x); uhi = fmaxf(uhi, n[i].x);
        vlo = fminf(vlo, n[i].y); vhi = fmaxf(vhi, n[i].y);
      }
      alo = fminf(alo, ulo); ahi = fmaxf(ahi, uhi);
      // (coordinates are >= 0: their bit patterns order like integers)
      const float umin = __int_as_float(__reduce_min_sync(0xffffffffu, __float_as_int(ulo)));
      const float umax = __int_as_float(__reduce_max_sync(0xffffffffu, __float_as_int(uhi)));
      const float vmin = __int_as_float(__reduce_min_sync(0xffffffffu, __float_as_int(vlo)));
      const float vmax = __int_as_float(__reduce_max_sync(0xffffffffu, __float_as_int(vhi)));
      const float amin = __int_as_float(__reduce_min_sync(0xffffffffu, __float_as_int(alo)));
      const float amax = __int_as_float(__reduce_max_sync(0xffffffffu, __float_as_int(ahi)));
      // the tile's own footprint wraps around the seam (no single box holds it) / some node of
      // its interpolation stencils lies across the seam (differences must be unwrapped)
      const bool box_straddle = umax - umin >= 0.5f * p.swf;
      const bool straddle = amax - amin >= 0.5f * p.swf;
      // smoothness: 4th difference (the cubic's error is ~0.023 of it) and 2nd differences
      // (a pixel leaves the hull of its neighbouring nodes by at most 1/8 of them)
      float d[4][2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = i < 2 ? i : i + 1;
        d[i][0] = n[j].x - n[2].x;
        d[i][1] = n[j].y - n[2].y;
        if (straddle) {
          if (d[i][0] > 0.5f * p.swf) d[i][0] -= p.swf;
          if (d[i][0] < -0.5f * p.swf) d[i][0] += p.swf;
        }
      }
      bool ok = true;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float d4 = (d[0][q] + d[3][q]) - 4.0f * (d[1][q] + d[2][q]);
        const float d2a = d[1][q] + d[2][q];
        const float d2b = d[0][q] - 2.0f * d[1][q];
        const float d2c = d[3][q] - 2.0f * d[2][q];
        ok = ok && fabsf(d4) < 0.03f && fabsf(d2a) < 2.0f && fabsf(d2b) < 2.0f && fabsf(d2c) < 2.0f;
      }
      if (__all_sync(0xffffffffu, ok)) flags |= L2_SMOOTH;
      if (straddle) flags |= L2_STRADDLE | L2_BORDER;
      if (umin < 1.0f || umax > p.swm1f - 2.0f || vmin < 1.0f || vmax > p.shm1f - 2.0f)
        flags |= L2_BORDER;
      // box: top-left taps of the hull (with the footprint shift at the last column / row),
      // one more pixel right and down, margin 2
      const int bx0 = (int)fminf(floorf(fminf(umin, p.swm1f)), p.swm1f - 1.0f);
      const int bx1 = (int)fminf(floorf(fminf(umax, p.swm1f)), p.swm1f - 1.0f);
      const int by0 = (int)fminf(floorf(fminf(vmin, p.shm1f)), p.shm1f - 1.0f);
      const int by1 = (int)fminf(floorf(fminf(vmax, p.shm1f)), p.shm1f - 1.0f);
      const int gx0 = (bx0 - 2) & ~(PER16 - 1);  // 16-byte aligned (TMA requirement)
      const int gy0 = by0 - 2;
      const int need_w = bx1 + 4 - gx0, need_h = by1 + 4 - gy0;
      const unsigned fit = __ballot_sync(
          0xffffffffu, need_w <= (my_box & 0xffff) && need_h <= (my_box >> 16));
      const int m = (box_straddle || !p.use_tma) ? -1 : __ffs(fit) - 1;
      if (lane == 0) headers[k] = make_int4(gx0, gy0, m, flags);
    }
  }
  __syncthreads();

  // ---- tile loop ---------------------------------------------------------------------------
  const int lx = lane & 15, ly = lane >> 4;
  const int row = warp * 2 + ly;
  const int yq = tile_y0 + row;
  const int y = min(yq, p.dh - 1);
  const unsigned tile_s = smem_u32(smem_raw);
  const T* const src_b = static_cast<const T*>(p.src) + (long long)b * p.ssb;
  T* const dst_row = static_cast<T*>(p.dst) + (long long)b * p.dsb + (long long)y * p.dsh;
  const int ssh = (int)p.ssh;
  // interpolation weights of this lane's pixels (offset t inside their 8-pixel interval)
  unsigned long long wt[P][3];
#pragma unroll
  for (int q = 0; q < P; ++q) {
    const int t = (lx * P + q) & 7;
#pragma unroll
    for (int j = 0; j < 3; ++j) wt[q][j] = pk2(g_l2_weights.w[t][j], g_l2_weights.w[t][j]);
  }
  RowCtx<XF> rc;
  if (EXACT) row_setup<XF>(p, b, y, rc);
  // boxed tiles so far: phase of both mbarriers.  Only boxed tiles touch them -- a warp that
  // arrived on `empty` for tiles it ran through without a box could complete a phase while a
  // slower warp still reads the buffer.
  unsigned nbox = 0;

  for (int k = 0; k < ntiles; ++k) {
    const int4 hdr = headers[k];
    const int gx0 = hdr.x, gy0 = hdr.y, m = hdr.z, flags = hdr.w;
    const int tile_x0 = strip_x0 + k * kL2TileW;
    const bool boxed = m >= 0;
    if (tid == 0 && boxed) {
      // the buffer is free once every warp has arrived for the previous boxed tile
      if (nbox > 0) mbar_wait(&empty_bar, (nbox - 1) & 1u);
      const int bw = p.tma_bw[m], bh = p.tma_bh[m];
      mbar_expect_tx(&full_bar, (unsigned)(bw * bh * NC * ELT));
      tma_load_box(smem_raw, &p.tmaps[m], &full_bar, gx0, gy0, 0, p.tma_batched ? b : 0);
    }
    __syncwarp();

    // this lane's pixels: tile-local x of pixel j
    auto xloc = [&](int j) { return (j / P) * 16 * P + lx * P + (j % P); };
    T* dst = dst_row + tile_x0 + lx * P;
    if (XF == EQB_EQUI2CUBE) dst = dst_row + p.cell_off[cell] + (tile_x0 - cell * p.w_face) + lx * P;

    if (flags & L2_BACKGROUND) {  // dice background cell
      if (yq < p.dh) {
        const float z[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = 0; c < NC; ++c) l2_store<T>(dst + c * p.dsc, z, 0u);
      }
      continue;
    }

    // ---- coordinates ------------------------------------------------------------------------
    float ixa[4], iya[4];
    const bool interp = !EXACT && (flags & (L2_SMOOTH | L2_RAGGED)) == L2_SMOOTH;
    if (interp) {
      const unsigned long long mone = pk2(-1.0f, -1.0f);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const unsigned long long* np = reinterpret_cast<const unsigned long long*>(
            &nodes[row][k * 8 + ((g * 16 * P + lx * P) >> 3)]);
        const unsigned long long n0 = np[0], n1 = np[1], n2 = np[2], n3 = np[3];
        unsigned long long d0 = fma2(n1, mone, n0), d2 = fma2(n1, mone, n2),
                           d3 = fma2(n1, mone, n3);
        if (flags & L2_STRADDLE) {  // unwrap the longitude across the seam
          auto unwrap = [&](unsigned long long& d) {
            float di, dj;
            unpk2(d, di, dj);
            if (di > 0.5f * p.swf) di -= p.swf;
            if (di < -0.5f * p.swf) di += p.swf;
            d = pk2(di, dj);
          };
          unwrap(d0); unwrap(d2); unwrap(d3);
        }
#pragma unroll
        for (int q = 0; q < P; ++q) {
          const unsigned long long u =
              fma2(wt[q][2], d3, fma2(wt[q][1], d2, fma2(wt[q][0], d0, n1)));
          unpk2(u, ixa[g * P + q], iya[g * P + q]);
        }
      }
      if (flags & L2_BORDER) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (ixa[j] < 0.0f) ixa[j] += p.swf;
          if (ixa[j] >= p.swf) ixa[j] -= p.swf;
          // the normalise / un-normalise round trip is clamp(u, 0, size-1) up to rounding
          ixa[j] = fminf(ixa[j], p.swm1f);
          iya[j] = fminf(fmaxf(iya[j], 0.0f), p.shm1f);
        }
      }
    } else {
      if (!EXACT) row_setup<XF>(p, b, y, rc);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int x = min(tile_x0 + xloc(j), p.dw - 1);
        float uj, ui;
        coords<XF, false>(p, rc, b, x, y, cell, min(x - cell * p.w_face, p.w_face - 1), uj, ui);
        ixa[j] = to_index(ui, p.inv_wm1, p.swm1f);
        iya[j] = to_index(uj, p.inv_hm1, p.shm1f);
      }
    }

    // ---- per-pixel sampling state --------------------------------------------------------------
    // (interior smooth tiles: no footprint shift can trigger, every tap is inside the box)
    const bool plain = interp && !(flags & L2_BORDER);
    unsigned a[4];
    float fx[4], fy[4];
    bool inside = true;
    {
      float xf[4], yf[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        xf[j] = floorf(ixa[j]);
        yf[j] = floorf(iya[j]);
      }
      if (!plain) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // footprint shifted to x0 = W-2 at ix == W-1 (taps_setup)
          xf[j] = fminf(xf[j], p.swm1f - 1.0f);
          yf[j] = fminf(yf[j], p.shm1f - 1.0f);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        fx[j] = __fsub_rn(ixa[j], xf[j]);
        fy[j] = __fsub_rn(iya[j], yf[j]);
      }
      if (boxed) {
        const int pitch = p.tma_bw[m];
        const float pitchf = (float)pitch;
        // byte address = ((yf - gy0) * pitch + (xf - gx0)) * ELT + tile_s; yf * pitch + xf is an
        // exact integer below 2^24
        const int base = (int)tile_s - (gy0 * pitch + gx0) * ELT;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          a[j] = (unsigned)((int)fmaf(yf[j], pitchf, xf[j]) * ELT + base);
        if (!(flags & L2_SMOOTH) || (flags & L2_RAGGED)) {
          const float gx0f = (float)gx0, gy0f = (float)gy0;
          const float xlim = (float)(pitch - 2), ylim = (float)(p.tma_bh[m] - 2);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float xr = xf[j] - gx0f, yr = yf[j] - gy0f;
            inside = inside && xr >= 0.0f && xr <= xlim && yr >= 0.0f && yr <= ylim;
          }
        }
      }
      if (!boxed || !inside) {
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] = (unsigned)((int)yf[j] * ssh + (int)xf[j]);
      }
    }
    if (boxed) mbar_wait(&full_bar, nbox & 1u);

    // ---- blend and store ----------------------------------------------------------------------
    const bool full_tile = !(flags & L2_RAGGED);
    if (yq < p.dh) {
      if (boxed && inside && full_tile) {
        auto blend_m = [&](auto mc) {
          constexpr int M = decltype(mc)::value;
          l2_blend<T, M, NC>(a, fx, fy, dst, p.dsc, p.flags);
        };
        switch (m) {
          case 0: blend_m(cuda::std::integral_constant<int, 0>{}); break;
          case 1: blend_m(cuda::std::integral_constant<int, 1>{}); break;
          case 2: blend_m(cuda::std::integral_constant<int, 2>{}); break;
          case 3: blend_m(cuda::std::integral_constant<int, 3>{}); break;
          case 4: blend_m(cuda::std::integral_constant<int, 4>{}); break;
          case 5: blend_m(cuda::std::integral_constant<int, 5>{}); break;
          case 6: blend_m(cuda::std::integral_constant<int, 6>{}); break;
          case 7: blend_m(cuda::std::integral_constant<int, 7>{}); break;
          case 8: blend_m(cuda::std::integral_constant<int, 8>{}); break;
          case 9: blend_m(cuda::std::integral_constant<int, 9>{}); break;
          case 10: blend_m(cuda::std::integral_constant<int, 10>{}); break;
          case 11: blend_m(cuda::std::integral_constant<int, 11>{}); break;
          default: blend_m(cuda::std::integral_constant<int, 12>{}); break;
        }
      } else {
        // slow paths: taps from the box through generic loads (ragged tiles) or from global
        // memory (no box, or a tap outside it); element stores
        float w_nw[4], w_ne[4], w_sw[4], w_se[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float ax = __fsub_rn(1.0f, fx[j]), ay = __fsub_rn(1.0f, fy[j]);
          w_nw[j] = __fmul_rn(ax, ay);
          w_ne[j] = __fmul_rn(fx[j], ay);
          w_sw[j] = __fmul_rn(ax, fy[j]);
          w_se[j] = __fmul_rn(fx[j], fy[j]);
        }
        const bool from_box = boxed && inside;
        const int pitch = boxed ? p.tma_bw[m] : 0;
        const int plane = boxed ? pitch * p.tma_bh[m] : 0;
        for (int c = 0; c < NC; ++c) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (tile_x0 + xloc(j) >= p.dw) continue;
            float v;
            if (from_box) {
              const T* q0 = reinterpret_cast<const T*>(smem_raw) + (a[j] - tile_s) / ELT + c * plane;
              v = blend4(lds_f(q0), lds_f(q0 + 1), lds_f(q0 + pitch), lds_f(q0 + pitch + 1),
                         w_nw[j], w_ne[j], w_sw[j], w_se[j]);
            } else {
              const T* q0 = src_b + (long long)c * p.ssc + (int)a[j];
              float v_nw, v_ne, v_sw, v_se;
              if constexpr (ELT == 4) {
                v_nw = ld_f(q0); v_ne = ld_f(q0 + 1); v_sw = ld_f(q0 + ssh); v_se = ld_f(q0 + ssh + 1);
              } else {
                ldg_pair<T>(q0, v_nw, v_ne);
                ldg_pair<T>(q0 + ssh, v_sw, v_se);
              }
              v = blend4(v_nw, v_ne, v_sw, v_se, w_nw[j], w_ne[j], w_sw[j], w_se[j]);
            }
            const T o = OutCvt<T>::cvt(v, p.flags);
            Pack<T, 1>::st(dst + c * p.dsc + (xloc(j) - lx * P), &o);
          }
        }
      }
    }
    if (boxed) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar);
      ++nbox;
    }
  }
}

}  // namespace eqb
