// STATUS: EXPERIMENT, NOT BUILT (round 2).  Correct -- bit-identical to remap_lean2.cuh on every
// path (profiles/r02_experiments/lean3_32x32_tiles_check.jsonl) -- but not faster on any rotation
// family: 32 x 4K bf16 bench rotations 1.56-1.63 ms vs 1.47 ms, identity 1.24-1.31 vs 1.24 ms.
// What the captures say (profiles/r02_experiments/ncu_lean3_*.txt): the per-pixel instruction
// count does drop (127 with pixel pairs per lane) but the pair mapping costs 2.17 wavefronts per
// tap LDS, and with one pixel per lane (142 instr / px) the kernel lands where lean2 is -- both
// are co-limited by the shared-memory data pipe (~65-70 %) and the issue slots (~65-70 %), so a
// restructuring that trades one for the other does not move the time.  The 32 x 32 tiles also
// need 3072+ pixels of box capacity per channel to fit at bench rolls (tools/sim_tiles.py), which
// costs occupancy.  Kept for the record; to build it, copy this file and remap_l3.inl into
// equilib_b200/csrc/, add remap_l3_xf{0,1,2}.cu (#define EQB_XF n / #include "remap_l3.inl") and
// the can_lean3 / setup_tma_l3 dispatch of git history (commit "lean3 experiment").
//
// remap_lean3_kernel: bilinear, TMA-staged, planar 1 / 3 / 4 channels of 2- and 4-byte
// pixels, equi2equi / equi2pers / equi2cube.  The strip-node kernel (remap_lean2.cuh) with
// its per-tile overheads cut, after its ncu capture (profiles/r02_ncu_lean2_v2.txt: 141
// instructions per pixel of which ~25 % were not work -- 8.5 % spinning on the box barriers,
// 3 % re-deriving shared-memory addresses, 12 per pixel of per-tile bookkeeping amortised
// over only 4 pixels per thread):
//
//  1. 32 x 32 TILES, 8 PIXELS PER THREAD.  A work item is a 32 x 32 tile -- or, where its
//     footprint fits no box (rolled views at high source latitudes), its upper / lower
//     32 x 16 half.  128 threads walk an item in batches of 32 x 16 pixels (4 per thread), so
//     the per-item bookkeeping (header, pointers, box wait) is paid once per 8 pixels, and a
//     square tile needs the smallest box per pixel under any roll.
//  2. ROWS INTERLEAVED BY CHANNEL.  The tensor map orders the source as (W, C, H, B) --
//     strides need not be monotonic (measured: tools/probes/tma_align_probe.cu) -- so ONE
//     box load lands as [row][channel][x].  The byte offset of tap (row r, channel c) from a
//     pixel's address is (r * C + c) * pitch: it depends on the box PITCH only, not on its
//     height, so boxes of any height share code and the blend is specialised on 8 pitches
//     (48 .. 128 pixels) instead of 11 box shapes.
//  3. NO WAITING PRODUCER.  The box of item n + 2 is requested by whichever warp finishes item
//     n LAST (a shared-memory counter per buffer), into the buffer that warp has just seen
//     drained: no empty-barrier, nobody spins waiting to issue.  Consumers block in
//     mbarrier.try_wait (the hardware suspends the warp) without a sleep loop.
//  4. TAPS AS IN remap_lean2.cuh: the 16 lanes of a row read 16 ADJACENT source pixels per tap
//     instruction as aligned 32-bit words + one PRMT (8-9 words per half warp: conflict-free
//     under any roll).  Measured alternative, dropped: a lane owning two adjacent pixels
//     (32-bit stores, shared stencil) spreads a tap instruction over 32 distinct words of two
//     rows -- 2.17 wavefronts per LDS instead of 1.5, slower overall.
//
// Coordinates: as in remap_lean2.cuh -- the exact chain at the strip's nodes (every 8th
// column of every row), cubic Lagrange interpolation along x with lane-constant weights on
// smooth interior tiles; the exact chain on every pixel in the EXACT variant and on tiles
// that are not smooth (source poles), touch the source border or straddle the seam.  Tiles
// whose footprint fits no box gather from global memory.  Same values on every path.
#pragma once

#include "remap_lean2.cuh"

namespace eqb {

constexpr int kL3Warps = 4;
constexpr int kL3TileW = 32, kL3TileH = 32, kL3BatchH = 16;
constexpr int kL3Step = 8;                        // node spacing along x
constexpr int kL3MaxIters = 6;                    // tiles per strip
constexpr int kL3Cols = kL3MaxIters * 4 + 3;      // node columns -1 .. 4 * iters + 1
constexpr int kL3Bufs = 2;
constexpr int kL3BufBytes = 18 * 1024;
constexpr int kL3Pitches = 8;
constexpr int kL3Boxes = 11;
static_assert(kL3Boxes <= kTmaSlots, "tensor-map slots");

// Box pitches in pixels (multiples of 16 bytes for every element size).
EQB_HD constexpr int l3_pitch(int pi) {
  return pi == 0 ? 48 : pi == 1 ? 56 : pi == 2 ? 64 : pi == 3 ? 72 : pi == 4 ? 80 :
         pi == 5 ? 96 : pi == 6 ? 112 : 128;
}
// Box m: boxes 0-2 are the short versions of pitches 0-2 (upright 32-row tiles: less to
// fetch); boxes 3-10 spend the whole buffer at pitches 0-7.  pick = the first that fits.
EQB_HD constexpr int l3_box_pi(int m) { return m < 3 ? m : m - 3; }
EQB_HD constexpr int l3_box_h(int elt, int nc, int m) {
  const int cap = kL3BufBytes / elt / nc;
  const int full0 = cap / l3_pitch(l3_box_pi(m));
  const int full = full0 < 96 ? full0 : 96;
  return m < 3 ? (38 < full ? 38 : full) : full;
}

enum : int { L3_SMOOTH = 1, L3_BORDER = 2, L3_STRADDLE = 4, L3_RAGGED = 8, L3_BACKGROUND = 16 };

__device__ __forceinline__ void mbar_wait_blocking(unsigned bar_addr, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
      "@P1 bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar_addr), "r"(parity) : "memory");
}

// Blend + store of one channel: pixel pairs (0, 1) [upper row] and (2, 3) [lower row] share
// FFMA2s; pixels 0 / 2 are column lx, pixels 1 / 3 column lx + 16 (l2_store).
template <typename T, int PITCH, int NC, int CI>
__device__ __forceinline__ void l3_blend_channel(const L2Px (&q)[4],
                                                 const unsigned long long (&w_nw)[2],
                                                 const unsigned long long (&w_ne)[2],
                                                 const unsigned long long (&w_sw)[2],
                                                 const unsigned long long (&w_se)[2], T* d0,
                                                 T* d1, long long dsc) {
  constexpr int ELT = (int)sizeof(T);
  constexpr int CH = CI * PITCH * ELT;         // channel CI of the tap's row
  constexpr int ROW = NC * PITCH * ELT;        // the row below
  float o[4];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const L2Px &x = q[2 * k], &y = q[2 * k + 1];
    unsigned long long acc =
        mul2(pk2(l2_tap<T, CH>(x.a0, x.s0), l2_tap<T, CH>(y.a0, y.s0)), w_nw[k]);
    acc = fma2(pk2(l2_tap<T, CH>(x.a1, x.s1), l2_tap<T, CH>(y.a1, y.s1)), w_ne[k], acc);
    acc = fma2(pk2(l2_tap<T, CH + ROW>(x.a0, x.s0), l2_tap<T, CH + ROW>(y.a0, y.s0)), w_sw[k], acc);
    acc = fma2(pk2(l2_tap<T, CH + ROW>(x.a1, x.s1), l2_tap<T, CH + ROW>(y.a1, y.s1)), w_se[k], acc);
    unpk2(acc, o[2 * k], o[2 * k + 1]);
  }
  l2_store<T>(d0 + CI * dsc, d1 + CI * dsc, o, true);
}

template <typename T, int PITCH, int NC>
__device__ __forceinline__ void l3_blend(const L2Px (&q)[4], const float (&fx)[4],
                                         const float (&fy)[4], T* d0, T* d1, long long dsc) {
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
  l3_blend_channel<T, PITCH, NC, 0>(q, w_nw, w_ne, w_sw, w_se, d0, d1, dsc);
  if constexpr (NC > 1) l3_blend_channel<T, PITCH, NC, 1>(q, w_nw, w_ne, w_sw, w_se, d0, d1, dsc);
  if constexpr (NC > 2) l3_blend_channel<T, PITCH, NC, 2>(q, w_nw, w_ne, w_sw, w_se, d0, d1, dsc);
  if constexpr (NC > 3) l3_blend_channel<T, PITCH, NC, 3>(q, w_nw, w_ne, w_sw, w_se, d0, d1, dsc);
}

// One boxed, smooth, interior work item on the fast path: `nbatch` batches of 32 x 16 pixels.
// nrow = this lane's first stencil node (row of its pixel 0, interval of column lx); the
// batch below is kL3BatchH node rows further.  x0 / y0 = this lane's pixel 0 (EXACT only).
template <int XF, typename T, int PITCH, int NC, bool EXACT>
__device__ __forceinline__ void l3_fast_item(const Params& p, const float2* nrow,
                                             const unsigned long long (&wt)[3], int base, T* d0,
                                             long long dsh, long long dsc, int nbatch,
                                             unsigned bar, unsigned parity, int b, int x0, int y0,
                                             int cell) {
  constexpr int ELT = (int)sizeof(T);
  constexpr int NODE_ROW = kL3Cols + 1;  // float2 elements per node row
  const float rowf = (float)(PITCH * NC);
#pragma unroll 1
  for (int bi = 0; bi < nbatch; ++bi) {
    float ixa[4], iya[4];
    if constexpr (!EXACT) {
      const unsigned long long mone = pk2(-1.0f, -1.0f);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const unsigned long long* np = reinterpret_cast<const unsigned long long*>(
            nrow + (bi * kL3BatchH + 2 * (j >> 1)) * NODE_ROW + 2 * (j & 1));
        const unsigned long long n0 = np[0], n1 = np[1], n2 = np[2], n3 = np[3];
        const unsigned long long d0n = fma2(n1, mone, n0), d2n = fma2(n1, mone, n2),
                                 d3n = fma2(n1, mone, n3);
        const unsigned long long u = fma2(wt[2], d3n, fma2(wt[1], d2n, fma2(wt[0], d0n, n1)));
        unpk2(u, ixa[j], iya[j]);
      }
    } else {
      RowCtx<XF> rc;
#pragma unroll
      for (int gy = 0; gy < 2; ++gy) {
        const int y = y0 + bi * kL3BatchH + 2 * gy;
        row_setup<XF>(p, b, y, rc);
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          const int j = 2 * gy + sl, x = x0 + 16 * sl;
          float uj, ui;
          coords<XF, false>(p, rc, b, x, y, cell, x - cell * p.w_face, uj, ui);
          ixa[j] = to_index(ui, p.inv_wm1, p.swm1f);
          iya[j] = to_index(uj, p.inv_hm1, p.shm1f);
        }
      }
    }
    L2Px px[4];
    float fx[4], fy[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float xf = floorf(ixa[j]), yf = floorf(iya[j]);
      if (EXACT) {  // footprint shifted to x0 = W-2 at ix == W-1 (taps_setup)
        xf = fminf(xf, p.swm1f - 1.0f);
        yf = fminf(yf, p.shm1f - 1.0f);
      }
      fx[j] = __fsub_rn(ixa[j], xf);
      fy[j] = __fsub_rn(iya[j], yf);
      // yf * NC * pitch + xf is an exact integer below 2^24
      l2_px_setup<T>((unsigned)((int)fmaf(yf, rowf, xf) * ELT + base), px[j]);
    }
    if (bi == 0) mbar_wait_blocking(bar, parity);
    T* const d = d0 + bi * kL3BatchH * dsh;
    l3_blend<T, PITCH, NC>(px, fx, fy, d, d + 2 * dsh, dsc);
  }
}

template <int XF, typename T, int NC, bool EXACT>
__global__ void __launch_bounds__(kL3Warps * 32, EXACT || sizeof(T) == 4 ? 4 : 5)
    remap_lean3_kernel(const __grid_constant__ Params p) {
  constexpr int ELT = (int)sizeof(T);
  constexpr int PER16 = 16 / ELT;
  constexpr int BUF = kL3BufBytes;
  extern __shared__ __align__(1024) unsigned char smem_raw[];          // kL3Bufs box buffers
  __shared__ __align__(16) float2 nodes[kL3TileH][kL3Cols + 1];        // (ui, uj)
  __shared__ __align__(16) int4 items[2 * kL3MaxIters];                // gx0, gy0, box, flags
  __shared__ int boxed_slot[2 * kL3MaxIters + 1];                      // boxed item n -> slot
  __shared__ int n_boxed;
  __shared__ int drained[kL3Bufs];                                     // warps done with a buffer
  __shared__ __align__(8) unsigned long long full_bar[kL3Bufs];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.z;
  const int tile_y0 = blockIdx.y * kL3TileH;
  const int strip_x0 = blockIdx.x * p.iters * kL3TileW;
  const int ntiles = min(p.iters, (p.dw - strip_x0 + kL3TileW - 1) / kL3TileW);
  int cell = 0;
  if (XF == EQB_EQUI2CUBE) cell = strip_x0 / p.w_face;  // (a strip never straddles two cells)
  const bool background = XF == EQB_EQUI2CUBE && cell >= 6;
  const unsigned tile_s = smem_u32(smem_raw);
  const unsigned bar_s = smem_u32(&full_bar[0]);

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kL3Bufs; ++i) {
      mbar_init(&full_bar[i], 1);
      drained[i] = 0;
    }
  }

  // ---- prologue 1: the exact chain at the strip's nodes -------------------------------
  if (!background) {
    const int row = tid & (kL3TileH - 1);
    const int y = min(tile_y0 + row, p.dh - 1);
    RowCtx<XF> rc;
    row_setup<XF>(p, b, y, rc);
    const int ncols = ntiles * 4 + 3;
    for (int col = tid >> 5; col < ncols; col += kL3Warps) {
      float uj, ui;
      node_coords<XF>(p, rc, strip_x0 + kL3Step * (col - 1), y, cell, uj, ui);
      nodes[row][col] = make_float2(ui, uj);
    }
  }
  __syncthreads();

  // ---- prologue 2: the work items of every tile (warp w: tiles w, w + 4) -----------------
  // slots 2k, 2k + 1: the whole tile k (slot 2k + 1 absent), or its upper and lower halves
  {
    const int my_box = lane < kL3Boxes ? (p.tma_bw[lane & 15] | (p.tma_bh[lane & 15] << 16)) : 0;
    for (int k = warp; k < ntiles; k += kL3Warps) {
      int flags = 0;
      if (strip_x0 + (k + 1) * kL3TileW > p.dw || tile_y0 + kL3TileH > p.dh) flags |= L3_RAGGED;
      if (background) {
        if (lane == 0) {
          items[2 * k] = make_int4(0, 0, -1, flags | L3_BACKGROUND);
          items[2 * k + 1] = make_int4(0, 0, -2, 0);
        }
        continue;
      }
      // lane = row: the seven nodes -1 .. 5 around the tile's columns 0 .. 4
      const float2* np = &nodes[lane][k * 4];  // (array index = node column + 1)
      float2 n[7];
#pragma unroll
      for (int i = 0; i < 7; ++i) n[i] = np[i];
      // hull of the tile's own columns (1 .. 5 in n[]) and of all seven
      float ulo = n[1].x, uhi = n[1].x, vlo = n[1].y, vhi = n[1].y;
#pragma unroll
      for (int i = 2; i < 6; ++i) {
        ulo = fminf(ulo, n[i].x); uhi = fmaxf(uhi, n[i].x);
        vlo = fminf(vlo, n[i].y); vhi = fmaxf(vhi, n[i].y);
      }
      const float alo = fminf(fminf(ulo, n[0].x), n[6].x), ahi = fmaxf(fmaxf(uhi, n[0].x), n[6].x);
      // smoothness of the row: 4th and 2nd differences of the two five-node windows
      bool ok = true;
#pragma unroll
      for (int c0 = 0; c0 < 3; c0 += 2) {
        const float ax = n[c0].x - n[c0 + 2].x, bx = n[c0 + 1].x - n[c0 + 2].x,
                    cx = n[c0 + 3].x - n[c0 + 2].x, dx = n[c0 + 4].x - n[c0 + 2].x;
        const float ay = n[c0].y - n[c0 + 2].y, by = n[c0 + 1].y - n[c0 + 2].y,
                    cy = n[c0 + 3].y - n[c0 + 2].y, dy = n[c0 + 4].y - n[c0 + 2].y;
        const float d4x = (ax + dx) - 4.0f * (bx + cx), d4y = (ay + dy) - 4.0f * (by + cy);
        const float d2x = bx + cx, d2y = by + cy;
        ok = ok && fmaxf(fabsf(d4x), fabsf(d4y)) < 0.03f && fmaxf(fabsf(d2x), fabsf(d2y)) < 1.5f;
      }
      // per half (lanes 0-15 / 16-31), then the whole tile.  (coordinates are >= 0: their
      // bit patterns order like integers)
      const unsigned hm = lane < 16 ? 0x0000ffffu : 0xffff0000u;
      float umin = __int_as_float(__reduce_min_sync(hm, __float_as_int(ulo)));
      float umax = __int_as_float(__reduce_max_sync(hm, __float_as_int(uhi)));
      float vmin = __int_as_float(__reduce_min_sync(hm, __float_as_int(vlo)));
      float vmax = __int_as_float(__reduce_max_sync(hm, __float_as_int(vhi)));
      float amin = __int_as_float(__reduce_min_sync(hm, __float_as_int(alo)));
      float amax = __int_as_float(__reduce_max_sync(hm, __float_as_int(ahi)));
      const bool smooth_h = __all_sync(hm, ok);
      const float umin_t = fminf(umin, __shfl_xor_sync(0xffffffffu, umin, 16));
      const float umax_t = fmaxf(umax, __shfl_xor_sync(0xffffffffu, umax, 16));
      const float vmin_t = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, 16));
      const float vmax_t = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, 16));
      const float amin_t = fminf(amin, __shfl_xor_sync(0xffffffffu, amin, 16));
      const float amax_t = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 16));
      const bool smooth_t = __all_sync(0xffffffffu, ok);

      // header of a hull: box origin, first box that fits, flags
      auto header = [&](float u0, float u1, float v0, float v1, float a0, float a1, bool smooth,
                        int& gx0, int& gy0, int& fl) -> unsigned {
        fl = flags;
        // the footprint wraps around the seam (no single box holds it) / some node of an
        // interpolation stencil lies across the seam (such tiles take the exact chain)
        const bool box_straddle = u1 - u0 >= 0.5f * p.swf;
        if (a1 - a0 >= 0.5f * p.swf) fl |= L3_STRADDLE;
        if (smooth) fl |= L3_SMOOTH;
        if (u0 < 1.0f || u1 > p.swm1f - 2.0f || v0 < 1.0f || v1 > p.shm1f - 2.0f) fl |= L3_BORDER;
        // box: top-left taps of the hull (with the footprint shift at the last column / row),
        // one more pixel right and down, margin 2
        const int bx0 = (int)fminf(floorf(fminf(u0, p.swm1f)), p.swm1f - 1.0f);
        const int bx1 = (int)fminf(floorf(fminf(u1, p.swm1f)), p.swm1f - 1.0f);
        const int by0 = (int)fminf(floorf(fminf(v0, p.shm1f)), p.shm1f - 1.0f);
        const int by1 = (int)fminf(floorf(fminf(v1, p.shm1f)), p.shm1f - 1.0f);
        gx0 = (bx0 - 2) & ~(PER16 - 1);  // 16-byte aligned (TMA requirement, measured)
        gy0 = by0 - 2;
        const int need_w = bx1 + 4 - gx0, need_h = by1 + 4 - gy0;
        const unsigned fit = __ballot_sync(
            0xffffffffu, need_w <= (my_box & 0xffff) && need_h <= (my_box >> 16));
        return (box_straddle || !p.use_tma) ? 0u : fit;
      };
      int gx, gy, fl;
      const unsigned fit_t = header(umin_t, umax_t, vmin_t, vmax_t, amin_t, amax_t, smooth_t,
                                    gx, gy, fl);
      if (fit_t) {
        if (lane == 0) {
          items[2 * k] = make_int4(gx, gy, __ffs(fit_t) - 1, fl | (2 << 8));
          items[2 * k + 1] = make_int4(0, 0, -2, 0);
        }
      } else {
        // two halves: lanes 0-15 hold the upper half's hull, lanes 16-31 the lower one's.
        // (header() votes with the whole warp: evaluate each half's hull on all lanes)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int src_lane = 16 * h;
          const float u0 = __shfl_sync(0xffffffffu, umin, src_lane);
          const float u1 = __shfl_sync(0xffffffffu, umax, src_lane);
          const float v0 = __shfl_sync(0xffffffffu, vmin, src_lane);
          const float v1 = __shfl_sync(0xffffffffu, vmax, src_lane);
          const float a0 = __shfl_sync(0xffffffffu, amin, src_lane);
          const float a1 = __shfl_sync(0xffffffffu, amax, src_lane);
          const bool sm = __shfl_sync(0xffffffffu, (int)smooth_h, src_lane) != 0;
          const unsigned fit_h = header(u0, u1, v0, v1, a0, a1, sm, gx, gy, fl);
          if (lane == 0)
            items[2 * k + h] = make_int4(gx, gy, fit_h ? __ffs(fit_h) - 1 : -1,
                                         fl | (1 << 8) | (h << 10));
        }
      }
    }
  }
  __syncthreads();
  if (tid == 0) {  // boxed items in order; the first kL3Bufs boxes are requested right away
    int n = 0;
    for (int s = 0; s < 2 * ntiles; ++s)
      if (items[s].z >= 0) boxed_slot[n++] = s;
    n_boxed = n;
  }
  __syncthreads();

  // Request the box of boxed item n into buffer n % kL3Bufs.
  auto issue = [&](int n) {
    const int4 h = items[boxed_slot[n]];
    const unsigned buf = (unsigned)n % kL3Bufs;
    mbar_expect_tx(&full_bar[buf], (unsigned)(p.tma_bw[h.z] * p.tma_bh[h.z] * NC * ELT));
    tma_load_box(smem_raw + buf * BUF, &p.tmaps[h.z], &full_bar[buf], h.x, 0, h.y,
                 p.tma_batched ? b : 0);
  };
  if (tid == 0) {
    const int nb = n_boxed;
#pragma unroll
    for (int i = 0; i < kL3Bufs; ++i)
      if (i < nb) issue(i);
  }

  // ---- item loop ---------------------------------------------------------------------------
  // Batch = 32 x 16 pixels: lane (lx, ly) of warp w owns, in rows 4w + ly and 4w + ly + 2,
  // the pixels lx and lx + 16.  Pixel j: row j >> 1, column slot j & 1.
  const int lx = lane & 15, ly = lane >> 4;
  const int row0 = warp * 4 + ly;
  const T* const src_b = static_cast<const T*>(p.src) + (long long)b * p.ssb;
  T* const dst_b = static_cast<T*>(p.dst) + (long long)b * p.dsb;
  const int ssh = (int)p.ssh;
  const long long dsh = p.dsh, dsc = p.dsc;
  // interpolation weights of this lane's pixels (offset t = lx & 7 inside their 8-pixel
  // interval: the same for lx and lx + 16)
  unsigned long long wt[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) wt[j] = pk2(g_l2_weights.w[lx & 7][j], g_l2_weights.w[lx & 7][j]);
  const int iv = lx >> 3;  // interval of pixel lx inside the tile (pixel lx + 16: iv + 2)
  // this lane's first output element of the strip (row tile_y0 + row0, column strip_x0 + lx)
  T* const d_strip = XF == EQB_EQUI2CUBE
      ? dst_b + p.cell_off[cell] + (strip_x0 - cell * p.w_face) + lx
      : dst_b + strip_x0 + lx;

  int nbox = 0;  // boxed items seen so far
  for (int slot = 0; slot < 2 * ntiles; ++slot) {
    const int4 hdr = items[slot];
    const int m = hdr.z;
    if (m == -2) continue;
    const int gx0 = hdr.x, gy0 = hdr.y, flags = hdr.w & 0xff;
    const int nbatch = (flags & L3_BACKGROUND) ? 2 : (hdr.w >> 8) & 3;
    const int half = (hdr.w >> 10) & 1;
    const int k = slot >> 1;
    const bool boxed = m >= 0;
    const unsigned buf = (unsigned)nbox % kL3Bufs;
    const bool fast = boxed && (EXACT ? (flags & (L3_SMOOTH | L3_RAGGED)) == L3_SMOOTH
                                      : (flags & (L3_SMOOTH | L3_BORDER | L3_STRADDLE |
                                                  L3_RAGGED)) == L3_SMOOTH);

    if (fast) {
      // ======== fast path: smooth interior tile with a box (never ragged: no row / column
      // checks); one straight-line batch loop per box pitch ============================
      const int pitch = p.tma_bw[m];
      // byte address = ((yf - gy0) * NC * pitch + (xf - gx0)) * ELT + buffer
      const int base = (int)(tile_s + buf * BUF) - (gy0 * pitch * NC + gx0) * ELT;
      const int brow = half * kL3BatchH + row0;
      T* const d0 = d_strip + (long long)(tile_y0 + brow) * dsh + k * kL3TileW;
      const float2* const nrow = &nodes[brow][k * 4 + iv];
      const unsigned bar = bar_s + buf * 8u, parity = ((unsigned)nbox / kL3Bufs) & 1u;
      auto run_p = [&](auto pc) {
        constexpr int PITCH = l3_pitch(decltype(pc)::value);
        l3_fast_item<XF, T, PITCH, NC, EXACT>(p, nrow, wt, base, d0, dsh, dsc, nbatch, bar, parity,
                                              b, strip_x0 + k * kL3TileW + lx, tile_y0 + brow,
                                              cell);
      };
      switch (pitch) {
        case l3_pitch(0): run_p(cuda::std::integral_constant<int, 0>{}); break;
        case l3_pitch(1): run_p(cuda::std::integral_constant<int, 1>{}); break;
        case l3_pitch(2): run_p(cuda::std::integral_constant<int, 2>{}); break;
        case l3_pitch(3): run_p(cuda::std::integral_constant<int, 3>{}); break;
        case l3_pitch(4): run_p(cuda::std::integral_constant<int, 4>{}); break;
        case l3_pitch(5): run_p(cuda::std::integral_constant<int, 5>{}); break;
        case l3_pitch(6): run_p(cuda::std::integral_constant<int, 6>{}); break;
        default: run_p(cuda::std::integral_constant<int, 7>{}); break;
      }
    } else {
      // ======== general path: background fill, or the exact chain on every pixel with taps
      // from the box or from global memory ==========================================
      if (boxed) mbar_wait_blocking(bar_s + buf * 8u, ((unsigned)nbox / kL3Bufs) & 1u);
      const int tile_x0 = strip_x0 + k * kL3TileW;
      const int pitch = boxed ? p.tma_bw[m] : 0;
      const int bh = boxed ? p.tma_bh[m] : 0;
      const T* const box = reinterpret_cast<const T*>(smem_raw + buf * BUF);
#pragma unroll 1
      for (int bi = 0; bi < nbatch; ++bi) {
        const int y0q = tile_y0 + (half + bi) * kL3BatchH + row0;
        T* const d0 = d_strip + (long long)min(y0q, p.dh - 1) * dsh + k * kL3TileW;
        RowCtx<XF> rc;
#pragma unroll 1
        for (int gy = 0; gy < 2; ++gy) {
          const int yq = y0q + 2 * gy;
          if (yq >= p.dh) break;
          if (!(flags & L3_BACKGROUND)) row_setup<XF>(p, b, yq, rc);
#pragma unroll 1
          for (int sl = 0; sl < 2; ++sl) {
            const int x = tile_x0 + lx + 16 * sl;
            if (x >= p.dw) continue;
            T* d = d0 + (gy ? 2 * dsh : 0) + 16 * sl;
            if (flags & L3_BACKGROUND) {  // dice background cell
              const T z = OutCvt<T>::cvt(0.0f, p.flags);
              for (int c = 0; c < NC; ++c) Pack<T, 1>::st(d + c * dsc, &z);
              continue;
            }
            float uj, ui;
            coords<XF, false>(p, rc, b, x, yq, cell, x - cell * p.w_face, uj, ui);
            const float ix = to_index(ui, p.inv_wm1, p.swm1f), iy = to_index(uj, p.inv_hm1, p.shm1f);
            const float xf = fminf(floorf(ix), p.swm1f - 1.0f), yf = fminf(floorf(iy), p.shm1f - 1.0f);
            const float bx = __fsub_rn(ix, xf), by = __fsub_rn(iy, yf);
            const float ax = __fsub_rn(1.0f, bx), ay = __fsub_rn(1.0f, by);
            const float w_nw = __fmul_rn(ax, ay), w_ne = __fmul_rn(bx, ay);
            const float w_sw = __fmul_rn(ax, by), w_se = __fmul_rn(bx, by);
            const int xi = (int)xf, yi = (int)yf;
            const int xr = xi - gx0, yr = yi - gy0;
            const bool from_box = boxed && xr >= 0 && xr <= pitch - 2 && yr >= 0 && yr <= bh - 2;
            for (int c = 0; c < NC; ++c) {
              float v_nw, v_ne, v_sw, v_se;
              if (from_box) {
                const T* q0 = box + (yr * NC + c) * pitch + xr;
                v_nw = lds_f(q0); v_ne = lds_f(q0 + 1);
                v_sw = lds_f(q0 + NC * pitch); v_se = lds_f(q0 + NC * pitch + 1);
              } else {
                const T* q0 = src_b + (long long)c * p.ssc + (yi * ssh + xi);
                if constexpr (ELT == 4) {
                  v_nw = ld_f(q0); v_ne = ld_f(q0 + 1); v_sw = ld_f(q0 + ssh); v_se = ld_f(q0 + ssh + 1);
                } else {
                  ldg_pair<T>(q0, v_nw, v_ne);
                  ldg_pair<T>(q0 + ssh, v_sw, v_se);
                }
              }
              const T o = OutCvt<T>::cvt(blend4(v_nw, v_ne, v_sw, v_se, w_nw, w_ne, w_sw, w_se),
                                         p.flags);
              Pack<T, 1>::st(d + c * dsc, &o);
            }
          }
        }
      }
    }

    if (boxed) {
      // this warp is done with the buffer; the warp that is done LAST requests the box of
      // boxed item nbox + kL3Bufs into it
      __syncwarp();
      if (lane == 0) {
        const int seen = atomicAdd(&drained[buf], 1);
        if (seen == kL3Warps - 1) {
          drained[buf] = 0;
          if (nbox + kL3Bufs < n_boxed) issue(nbox + kL3Bufs);
        }
      }
      ++nbox;
    }
  }
}

}  // namespace eqb
