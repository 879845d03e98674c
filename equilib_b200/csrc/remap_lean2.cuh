// remap_lean2_kernel<XF, T, NC, EXACT, MODE>: the strip-node kernel -- TMA-staged, planar
// 1 / 3 / 4 channels.  Round 2 built it for the headline case (bilinear equi2equi / equi2pers /
// equi2cube); since then it also serves
//   MODE = BICUBIC   4 x 4 taps from the same boxes, packed row dot-products, in-kernel clamp;
//                    also pers2equi (fill strips decided from tile corners) and cube2equi from a
//                    horizon strip;
//   MODE = NEAREST   one tap per channel; 8/16-bit images round interpolated coordinates away
//                    from rounding ties and defer the rest to the exact chain (identical taps).
// Designed from ncu captures of its predecessors
// (profiles/r01_ncu_remap_bf16_v6_lean_final.txt: 147 instructions per pixel, L1 data pipe
// 82 %, 2.3 wavefronts per 2-byte tap; profiles/r02_ncu_lean2_v1.txt: 5.3x the source bytes
// fetched as TMA boxes, a fifth of all issued instructions spinning on the box barrier):
//
//  1. COORDINATES FROM A STRIP-WIDE NODE GRID.  A CTA walks a strip of up to 8 tiles
//     (32 x 16 pixels each) along x.  Before the first tile all 128 threads evaluate the
//     exact coordinate chain at the strip's NODES -- every 8th column of every row, plus one
//     column on either side -- into shared memory (fully occupied warps: 1/8 node per pixel
//     instead of the 1/4 of round 1, where every lane computed its own).  A pixel then gets
//     (ui, uj) from the four nodes around it on its own row by cubic Lagrange interpolation
//     with weights that are LANE CONSTANTS (a lane's pixels all sit at the same offset inside
//     their 8-pixel interval), as differences from the nearest node: 4 LDS.64 + 3 packed
//     subtractions + 3 FFMA2 per pixel, no shuffles, no weight loads.
//  2. PER-TILE HEADERS.  The same prologue reduces each tile's node hull to its TMA box
//     (origin, shape) and to flags -- smooth (4th and 2nd differences of the nodes small: the
//     interpolation error is below 1e-3 px and no pixel leaves the node hull by more than
//     1/4 px, so no per-pixel "inside the box" test is needed), border (a clamp, the wrap or
//     the footprint shift at the last column could trigger; bicubic: a tap would be
//     reflected), seam, fill (pers2equi), ties (nearest).  Smooth interior tiles with a box
//     take a straight-line fast path: no reductions, votes, clamps or wraps.
//  3. SQUARER TILES, TWO BOX BUFFERS.  A 32 x 16 tile of a view rolled by 0.3 rad needs a
//     48 x 30 box where the 64 x 8 tile of round 1 needed 96 x 40: less than half the L2 ->
//     shared traffic per pixel.  The box of the NEXT tile is requested while the current one
//     is blended (two buffers, full / empty mbarriers per buffer, no __syncthreads in the tile
//     loop, waits suspend in hardware instead of spinning on the issue slots).
//  4. ALIGNED-WORD TAPS.  The 16 lanes of a row read 16 ADJACENT source pixels per tap
//     instruction.  A 1- or 2-byte tap is fetched as the aligned 32-bit word that holds it
//     (lanes sharing a word read the same address: a broadcast, one wavefront) and moved
//     into place by one PRMT with a per-pixel selector -- the same two instructions as
//     LDS.U16 + shift, without the extra wavefronts of sub-word accesses (measured: 1.7 per
//     LDS.U16 when neighbouring lanes read the two halves of one word).  A bicubic tap row is
//     the two or three aligned words that hold its four taps plus funnel shifts.
//
// Tiles that are not smooth (source poles), border / seam / ragged tiles and the EXACT variant
// evaluate the exact chain on every pixel; tiles whose footprint fits no box (poles, seam)
// gather from global memory -- same values on every path
// (tests: test_staged_and_direct_kernels_agree, test_lean2_bicubic_equals_the_direct_kernel,
// test_lean2_nearest_equals_the_direct_kernel).
#pragma once

#include "remap_kernels.cuh"

namespace eqb {

#ifndef EQB_L2_BUF_KB
#define EQB_L2_BUF_KB 12
#endif
constexpr int kL2Warps = 4;
constexpr int kL2TileW = 32, kL2TileH = 16;
constexpr int kL2Step = 8;                          // node spacing along x
constexpr int kL2MaxIters = 8;                      // tiles per strip
constexpr int kL2Cols = kL2MaxIters * 4 + 3;        // node columns -1 .. 4 * iters + 1
constexpr int kL2Boxes = 11;
constexpr int kL2TieCap = 256;                      // deferred pixels per strip (nearest)
constexpr int kL2MinBoxH = kL2TileH + 5;            // an upright tile: 16 rows + 1 + margins
static_assert(kL2Boxes <= kTmaSlots, "tensor-map slots");

// One box buffer (a CTA holds two).
#ifndef EQB_L2_BUF_KB_F32
#define EQB_L2_BUF_KB_F32 18  // (measured 16 / 18 / 20 KB: 1.17 / 1.10 / 1.13 ms on 16 x 4K fp32, bench rotations)
#endif
// (bicubic: 18 KB for every type -- its 96 registers allow 5 CTAs per SM anyway, and the 4 x 4
// footprint of a tile is 3 pixels wider and taller: with 12 KB, 12 % of the tiles of the bench
// rotations and 25 % on rotations drawn from SO(3) fit no box; with 18 KB 5 % / 7 %
// (tools/sim_tiles.py))
#ifndef EQB_L2_BUF_KB_F32_CUBIC
#define EQB_L2_BUF_KB_F32_CUBIC 24  // (measured 18 / 24 / 32 KB: 2.75 / 2.07 / 2.11 ms on 16 x 4K fp32 bicubic, bench rotations)
#endif
EQB_HD constexpr int l2_buf_bytes(int elt, bool cubic = false) {
  return (elt == 4 && cubic) ? EQB_L2_BUF_KB_F32_CUBIC * 1024
         : (elt == 4 || cubic) ? EQB_L2_BUF_KB_F32 * 1024 : 12 * 1024;
}
// CTAs per SM the shared memory allows (two buffers + ~6 KB of nodes, headers, reserve)
EQB_HD constexpr int l2_ctas(int elt, bool cubic = false) {
  return 227 * 1024 / (2 * l2_buf_bytes(elt, cubic) + 6 * 1024);
}

// Box m: l2_box_w x l2_box_h source pixels x channels.  Boxes 0-1 are small (upright and
// mildly rolled views); the others spend the whole buffer in different aspect ratios: 48
// wide for rolls, 64-96 for the longitude stretch 1 / cos(lat) at high source latitudes,
// 40 / 32 for rolls beyond 45 degrees.  pick = the first that fits.
EQB_HD constexpr int l2_box_w(int elt, int m) {
  const int w = m <= 2 ? 48 : m == 3 ? 64 : m == 4 ? 56 : m == 5 ? 80 : m == 6 ? 72 :
                m == 7 ? 88 : m == 8 ? 96 : m == 9 ? 40 : 32;
  const int per16 = 16 / elt;
  return (w + per16 - 1) / per16 * per16;
}
EQB_HD constexpr int l2_box_h(int elt, int nc, int m, bool cubic = false) {
  const int budget = l2_buf_bytes(elt, cubic) / elt / nc;
  const int full = budget / l2_box_w(elt, m) < 128 ? budget / l2_box_w(elt, m) : 128;
  const int small = m == 0 ? 24 : 32;
  return m < 2 && small < full ? small : full;
}

// Cubic Lagrange weights of the nodes at -8, 0, +8, +16 for the pixel at offset t (0..7)
// from the second one; the weight of that node is not stored (the weights sum to 1:
// u = n1 + sum w_j (n_j - n1)).
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

enum : int { L2_SMOOTH = 1, L2_BORDER = 2, L2_STRADDLE = 4, L2_RAGGED = 8, L2_BACKGROUND = 16,
              L2_FILL = 32, L2_TIES = 64 };

// (barriers are addressed by their 32-bit shared-memory address, computed once per kernel:
// re-deriving it from a generic pointer at every use was 3 % of the issued instructions)
__device__ __forceinline__ void mbar_arrive(unsigned bar_addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
// Wait on a phase: try_wait suspends the warp in hardware up to the time hint, so the loop
// around it rarely iterates (the earlier probe + nanosleep(32) loop spent 8 % of the issued
// instructions spinning).
__device__ __forceinline__ void mbar_wait_backoff(unsigned bar_addr, unsigned parity) {
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

// Two nodes of one row at once (the packed chain, coords2).
template <int XF>
__device__ __forceinline__ void node_coords2(const Params& p, const RowCtx<XF>& c, int x0, int x1,
                                             int cell, float (&ui)[2], float (&uj)[2]) {
  if (XF == EQB_EQUI2EQUI) {
    const int xa = x0 < 0 ? x0 + p.dw : (x0 >= p.dw ? x0 - p.dw : x0);
    const int xb = x1 < 0 ? x1 + p.dw : (x1 >= p.dw ? x1 - p.dw : x1);
    coords2<XF, false>(p, c, xa, xb, 0, 0.f, 0.f, ui, uj);
  } else if (XF == EQB_EQUI2PERS) {
    coords2<XF, false>(p, c, x0, x1, 0, 0.f, 0.f, ui, uj);
  } else {  // EQB_EQUI2CUBE: the ramp continued linearly outside the face
    const int w = p.w_face;
    float rc[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int lx = (i ? x1 : x0) - cell * w;
      if (lx >= 0 && lx < w) {
        rc[i] = __ldg(p.tx + lx);
      } else {
        const float step = __ldg(p.tx + 1) - __ldg(p.tx);
        rc[i] = lx < 0 ? fmaf((float)lx, step, __ldg(p.tx))
                       : fmaf((float)(lx - (w - 1)), step, __ldg(p.tx + w - 1));
      }
    }
    coords2<XF, false>(p, c, x0, x1, cell, rc[0], rc[1], ui, uj);
  }
}

// pers2equi (XF == EQB_PERS2EQUI, bicubic, exact chain): a NODE carries the projective
// coordinates of its ray where it looks into the camera's half space -- they only feed the tile
// headers (hull, smoothness, box), never a pixel, so the quotient need not be the IEEE one --
// else a negative sentinel (such tiles take the general path), and a mask of the "outside the
// field of view" conditions it satisfies with the room pers2equi_run_surely_outside() asks of
// the two ends of an 8-pixel run: a tile whose nodes (every 8th column of every row) all share
// one condition is fill (out[mask] = 0, then the clamp: pers2equi/torch.py:232-251) -- 92 % of
// an 8K output at fov 90 -- and costs no divide, no tap and no per-pixel test.
__device__ __forceinline__ void pers_node(const Params& p, const RowCtx<EQB_PERS2EQUI>& c, int x,
                                          float& ui, float& uj, unsigned& mask) {
  const int xw = x < 0 ? x + p.dw : (x >= p.dw ? x - p.dw : x);
  const float2 cs = __ldg(reinterpret_cast<const float2*>(p.tx) + xw);
  const float dth = (float)kL2Step * (kTwoPi / (float)p.dw);
  const float room = 0.25f * dth * dth * fabsf(c.r0) *
                         (fabsf(c.A[0]) + fabsf(c.A[1]) + fabsf(c.A[3]) + fabsf(c.A[4]) +
                          (p.swf + p.shf) * (fabsf(c.A[6]) + fabsf(c.A[7]))) + 1e-3f;
  const float mx = c.r0 * cs.x, my = c.r0 * cs.y;
  const float X = c.A[0] * mx + c.A[1] * my + c.z0;
  const float Y = c.A[3] * mx + c.A[4] * my + c.z1;
  const float Z = c.A[6] * mx + c.A[7] * my + c.z2;
  const float wl = p.swf - 0.499f, hl = p.shf - 0.499f;
  mask = (Z < -room ? 1u : 0u) | (X + 0.501f * Z < -room ? 2u : 0u) |
         (wl * Z - X < -room ? 4u : 0u) | (Y + 0.501f * Z < -room ? 8u : 0u) |
         (hl * Z - Y < -room ? 16u : 0u);
  ui = uj = -1.0e4f;
  if (Z > 1e-6f) {
    const float rz = __frcp_rn(Z);
    ui = fminf(fmaxf(fmaf(X, rz, 0.5f), -1.0e4f), 1.0e5f);
    uj = fminf(fmaxf(fmaf(Y, rz, 0.5f), -1.0e4f), 1.0e5f);
  }
}

// ---- taps --------------------------------------------------------------------------------
// Sampling state of one pixel on the fast path: the word addresses of its left and right taps
// (row y0, channel 0) and the PRMT selectors that move the tap out of its word.
struct L2Px {
  unsigned a0, a1;  // shared-memory byte addresses
  unsigned s0, s1;  // PRMT selectors (fp16 / fp32: unused)
};

template <typename T>
__device__ __forceinline__ void l2_px_setup(unsigned a, L2Px& q) {
  if (sizeof(T) == 2 && !cuda::std::is_same<T, __half>::value) {
    // bf16 -> fp32 is "the 16 bits in the upper half": selector 0x1044 lifts the low half of
    // the word ([0, 0, b0, b1]), 0x3244 keeps the high half ([0, 0, b2, b3])
    q.a0 = a & ~3u;
    q.a1 = (a + 2u) & ~3u;
    q.s0 = (a & 2u) ? 0x3244u : 0x1044u;
    q.s1 = q.s0 ^ 0x2200u;
  } else if (sizeof(T) == 1) {
    // byte k of the word into byte 0 of 0x4B0000kk = 2^23 + k
    q.a0 = a & ~3u;
    q.a1 = (a + 1u) & ~3u;
    q.s0 = 0x7540u | (a & 3u);
    q.s1 = 0x7540u | ((a + 1u) & 3u);
  } else {
    q.a0 = a;
    q.a1 = a + (unsigned)sizeof(T);
    q.s0 = q.s1 = 0u;
  }
}

// The tap at compile-time byte offset OFF (row / channel) from word address `addr`.
template <typename T, int OFF>
__device__ __forceinline__ float l2_tap(unsigned addr, unsigned sel) {
  if (sizeof(T) == 4) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));
    return v;
  }
  if (cuda::std::is_same<T, __half>::value) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1+%2];" : "=h"(v) : "r"(addr), "n"(OFF));
    return __half2float(__ushort_as_half(v));
  }
  unsigned w;
  asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(w) : "r"(addr), "n"(OFF));
  if (sizeof(T) == 1)
    return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, sel)), 8388608.0f);
  return __uint_as_float(__byte_perm(w, 0u, sel));
}

// Store of a lane's four results of one channel: d0 / d1 point at its pixels of the upper /
// lower of its two rows.  2- and 4-byte types: pixels d[0], d[16]; uint8: the pair d[0..1].
template <typename T>
__device__ __forceinline__ void l2_store(T* d0, T* d1, const float (&o)[4], bool row1) {
  if constexpr (sizeof(T) == 4) {
    __stcs(reinterpret_cast<float*>(d0), o[0]);
    __stcs(reinterpret_cast<float*>(d0) + 16, o[1]);
    if (row1) {
      __stcs(reinterpret_cast<float*>(d1), o[2]);
      __stcs(reinterpret_cast<float*>(d1) + 16, o[3]);
    }
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
    unsigned short* q0 = reinterpret_cast<unsigned short*>(d0);
    unsigned short* q1 = reinterpret_cast<unsigned short*>(d1);
    __stcs(q0, (unsigned short)lo);
    __stcs(q0 + 16, (unsigned short)(lo >> 16));
    if (row1) {
      __stcs(q1, (unsigned short)hi);
      __stcs(q1 + 16, (unsigned short)(hi >> 16));
    }
  } else {
    // bilinear code values are convex combinations: floor == the reference's truncation
    const unsigned t0 = u8_floor_bits(o[0]), t1 = u8_floor_bits(o[1]);
    const unsigned t2 = u8_floor_bits(o[2]), t3 = u8_floor_bits(o[3]);
    __stcs(reinterpret_cast<unsigned short*>(d0), (unsigned short)__byte_perm(t0, t1, 0x0040u));
    if (row1)
      __stcs(reinterpret_cast<unsigned short*>(d1), (unsigned short)__byte_perm(t2, t3, 0x0040u));
  }
}

// Blend + store of one channel from box M: pixel pairs (0, 1) and (2, 3) share FFMA2s.
template <typename T, int M, int NC, int CI>
__device__ __forceinline__ void l2_blend_channel(const L2Px (&q)[4],
                                                 const unsigned long long (&w_nw)[2],
                                                 const unsigned long long (&w_ne)[2],
                                                 const unsigned long long (&w_sw)[2],
                                                 const unsigned long long (&w_se)[2], T* d0,
                                                 T* d1, long long dsc, bool row1) {
  constexpr int ELT = (int)sizeof(T);
  constexpr int ROW = l2_box_w(ELT, M) * ELT;
  constexpr int CH = CI * ROW * l2_box_h(ELT, NC, M);
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
  l2_store<T>(d0 + CI * dsc, d1 + CI * dsc, o, row1);
}

template <typename T, int M, int NC>
__device__ __forceinline__ void l2_blend(const L2Px (&q)[4], const float (&fx)[4],
                                         const float (&fy)[4], T* d0, T* d1, long long dsc,
                                         bool row1) {
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
  l2_blend_channel<T, M, NC, 0>(q, w_nw, w_ne, w_sw, w_se, d0, d1, dsc, row1);
  if constexpr (NC > 1) l2_blend_channel<T, M, NC, 1>(q, w_nw, w_ne, w_sw, w_se, d0, d1, dsc, row1);
  if constexpr (NC > 2) l2_blend_channel<T, M, NC, 2>(q, w_nw, w_ne, w_sw, w_se, d0, d1, dsc, row1);
  if constexpr (NC > 3) l2_blend_channel<T, M, NC, 3>(q, w_nw, w_ne, w_sw, w_se, d0, d1, dsc, row1);
}


// ---- bicubic (MODE == EQB_BICUBIC) ----------------------------------------------------------
// The 4 x 4 taps of a pixel come from the same box (hull of the node taps widened by one pixel
// to the left / top and two to the right / bottom); tiles within two pixels of the source
// border -- where ATen reflects taps -- take the general path.  Arithmetic and order are those
// of sample<..., Taps<EQB_BICUBIC>> (rows first, then columns; A = -0.75), two pixels per packed
// instruction, so every kernel of the library produces the same bits
// (test_staged_and_direct_kernels_agree).

// cubic_coeffs() for two pixels at once (t + 1 and 1 - t as x * 1 + 1 / x * -1 + 1: the same
// single rounding as the scalar add / subtract)
__device__ __forceinline__ void cubic_coeffs2(unsigned long long t, unsigned long long (&c)[4]) {
  const float A = -0.75f;
  const unsigned long long one = bc2(1.0f);
  unsigned long long x = fma2(t, one, one);
  c[0] = fma2(fma2(fma2(bc2(A), x, bc2(-5.0f * A)), x, bc2(8.0f * A)), x, bc2(-4.0f * A));
  c[1] = fma2(mul2(fma2(bc2(A + 2.0f), t, bc2(-(A + 3.0f))), t), t, one);
  x = fma2(t, bc2(-1.0f), one);
  c[2] = fma2(mul2(fma2(bc2(A + 2.0f), x, bc2(-(A + 3.0f))), x), x, one);
  x = fma2(x, one, one);
  c[3] = fma2(fma2(fma2(bc2(A), x, bc2(-5.0f * A)), x, bc2(8.0f * A)), x, bc2(-4.0f * A));
}

// Four horizontally adjacent taps starting at byte address `a` of the box (the box shape is a
// run-time value here, unlike in the bilinear blend: eleven unrolled copies of the 96-tap blend
// were 13 k instructions, and ncu showed half of all stall samples waiting for instruction
// fetch -- profiles/r02_experiments/ncu_lean2_bicubic_unrolled_boxes.txt).  1- and 2-byte
// pixels: the aligned 32-bit words that hold the taps (`aw` = a & ~3: lanes sharing a word
// broadcast) and one or two funnel shifts by `sh` bits.
template <typename T>
__device__ __forceinline__ void l2_taps4(unsigned aw, unsigned sh, float (&v)[4]) {
  if constexpr (sizeof(T) == 4) {
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[0]) : "r"(aw));
    asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(v[1]) : "r"(aw));
    asm volatile("ld.shared.f32 %0, [%1+8];" : "=f"(v[2]) : "r"(aw));
    asm volatile("ld.shared.f32 %0, [%1+12];" : "=f"(v[3]) : "r"(aw));
  } else if constexpr (sizeof(T) == 2) {
    unsigned w0, w1, w2;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(aw));
    asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(aw));
    asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(w2) : "r"(aw));
    const unsigned x0 = __funnelshift_r(w0, w1, sh), x1 = __funnelshift_r(w1, w2, sh);
    if constexpr (cuda::std::is_same<T, __half>::value) {
      const __half2 h0 = *reinterpret_cast<const __half2*>(&x0);
      const __half2 h1 = *reinterpret_cast<const __half2*>(&x1);
      v[0] = __low2float(h0); v[1] = __high2float(h0);
      v[2] = __low2float(h1); v[3] = __high2float(h1);
    } else {
      v[0] = __uint_as_float(x0 << 16); v[1] = __uint_as_float(x0 & 0xffff0000u);
      v[2] = __uint_as_float(x1 << 16); v[3] = __uint_as_float(x1 & 0xffff0000u);
    }
  } else {
    unsigned w0, w1;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(aw));
    asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(aw));
    const unsigned x = __funnelshift_r(w0, w1, sh);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      v[i] = __fsub_rn(__uint_as_float(__byte_perm(x, 0x4B000000u, 0x7540u | i)), 8388608.0f);
  }
}

// Two results of one channel: the lane's two pixels of ONE row (2- and 4-byte types: d[0] and
// d[16]; uint8: the adjacent pair d[0..1]), clamped, converted like OutCvt.
template <typename T>
__device__ __forceinline__ void l2_store_pair(T* d, float o0, float o1, bool clip, float lo,
                                              float hi, unsigned flags) {
  if (clip) {
    o0 = fminf(fmaxf(o0, lo), hi);
    o1 = fminf(fmaxf(o1, lo), hi);
  }
  const T t0 = OutCvt<T>::cvt(o0, flags), t1 = OutCvt<T>::cvt(o1, flags);
  if constexpr (sizeof(T) == 1) {
    const T t[2] = {t0, t1};
    Pack<T, 2>::st(d, t);
  } else {
    Pack<T, 1>::st(d, &t0);
    Pack<T, 1>::st(d + 16, &t1);
  }
}

// One channel of a pixel pair (x, y): 2 x 16 taps from the box rows at ax / ay + j * row bytes,
// rows then columns.
template <typename T>
__device__ __forceinline__ void l2_bicubic_channel(unsigned ax, unsigned shx, unsigned ay,
                                                   unsigned shy, unsigned row,
                                                   const unsigned long long (&cx)[4],
                                                   const unsigned long long (&cy)[4], float& ox,
                                                   float& oy) {
  unsigned long long acc = 0ull;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float vx[4], vy[4];
    l2_taps4<T>(ax + j * row, shx, vx);
    l2_taps4<T>(ay + j * row, shy, vy);
    unsigned long long r = mul2(pk2(vx[0], vy[0]), cx[0]);
    r = fma2(pk2(vx[1], vy[1]), cx[1], r);
    r = fma2(pk2(vx[2], vy[2]), cx[2], r);
    r = fma2(pk2(vx[3], vy[3]), cx[3], r);
    acc = j == 0 ? mul2(r, cy[0]) : fma2(r, cy[j], acc);
  }
  unpk2(acc, ox, oy);
}

// The lane's pixel pair of one row (pixels x, y; top-left taps at byte addresses ax, ay of
// channel 0) from a box with `row` bytes per row and `plane` bytes per channel, every channel.
template <typename T, int NC>
__device__ __forceinline__ void l2_bicubic_pair(unsigned ax, unsigned ay, float fxx, float fxy,
                                                float fyx, float fyy, unsigned row,
                                                unsigned plane, T* d, long long dsc, bool clip,
                                                float lo, float hi, unsigned flags) {
  unsigned long long cx[4], cy[4];
  cubic_coeffs2(pk2(fxx, fxy), cx);
  cubic_coeffs2(pk2(fyx, fyy), cy);
  constexpr unsigned MASK = sizeof(T) == 4 ? 0u : 3u;
  const unsigned shx = (ax & MASK) * 8u, shy = (ay & MASK) * 8u;
  ax &= ~MASK;
  ay &= ~MASK;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    float ox, oy;
    l2_bicubic_channel<T>(ax + c * plane, shx, ay + c * plane, shy, row, cx, cy, ox, oy);
    l2_store_pair<T>(d + c * dsc, ox, oy, clip, lo, hi, flags);
  }
}

// ---- nearest (MODE == EQB_NEAREST) ----------------------------------------------------------
// Distance from a rounding tie below which an INTERPOLATED coordinate does not decide the tap
// (the pixel gets the exact chain): 4e-3 px up to 4096-pixel sources -- the cubic's error on a
// smooth tile is < 1e-3 px (header test) and the measured worst case against the exact chain,
// rounding noise of both included, 1.7e-3 px at 4K -- growing with the source size like the
// fp32 resolution of a coordinate (4.9e-4 px at 4K, 9.8e-4 px at 8K).
__device__ __forceinline__ float l2_tie_guard(const Params& p) {
  return 4e-3f * fmaxf(1.0f, fmaxf(p.swf, p.shf) * (1.0f / 4096.0f));
}

// The tap is the source pixel at (rint(iy), rint(ix)) -- nearbyint, like ATen -- and the output
// element is that element, bit for bit: no conversion in either direction.  `a` = its byte
// address in the box (channel 0), `plane` the bytes of one channel of the box.
template <typename T> __device__ __forceinline__ T l2_lds_elem(unsigned a) {
  if constexpr (sizeof(T) == 4) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return *reinterpret_cast<const T*>(&v);
  } else if constexpr (sizeof(T) == 2) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return *reinterpret_cast<const T*>(&v);
  } else {
    unsigned short v;  // (no 8-bit register class: zero-extended into 16 bits)
    asm volatile("ld.shared.u8 %0, [%1];" : "=h"(v) : "r"(a));
    return (T)v;
  }
}

// The lane's two pixels of one row from byte addresses ax, ay: 2- and 4-byte types d[0] and
// d[16], uint8 the adjacent pair d[0..1].
template <typename T, int NC>
__device__ __forceinline__ void l2_nearest_pair(unsigned ax, unsigned ay, unsigned plane, T* d,
                                                long long dsc) {
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const T t0 = l2_lds_elem<T>(ax + c * plane), t1 = l2_lds_elem<T>(ay + c * plane);
    if constexpr (sizeof(T) == 1) {
      const T t[2] = {t0, t1};
      Pack<T, 2>::st(d + c * dsc, t);
    } else {
      Pack<T, 1>::st(d + c * dsc, &t0);
      Pack<T, 1>::st(d + c * dsc + 16, &t1);
    }
  }
}

// One pixel on the general path (tiles at the source border, poles, seam; footprints that fit
// no box): the 4 x 4 taps from the box where all of them lie inside it and none is reflected,
// else from global memory, each tap reflected about 0 and size-1 like ATen.  One copy of the
// code (not inlined: the general path is rare, the instruction cache is not).
template <typename T, int NC>
__device__ __noinline__ void l2_bicubic_general(const Params& p, float ix, float iy, bool boxed,
                                                const T* box, int pitch, int bh, int gx0, int gy0,
                                                const T* src_b, T* d, bool clip, float clip_lo,
                                                float clip_hi) {
  const int ssh = (int)p.ssh;
  const float xf = floorf(ix), yf = floorf(iy);
  float cx[4], cy[4];
  cubic_coeffs(__fsub_rn(ix, xf), cx);
  cubic_coeffs(__fsub_rn(iy, yf), cy);
  const int xi = (int)xf, yi = (int)yf;
  const int xr = xi - 1 - gx0, yr = yi - 1 - gy0;
  const bool interior = xi >= 1 && xi + 2 <= p.sw - 1 && yi >= 1 && yi + 2 <= p.sh - 1;
  const bool from_box =
      boxed && interior && xr >= 0 && xr + 3 <= pitch - 1 && yr >= 0 && yr + 3 <= bh - 1;
  int ox[4], oy[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ox[i] = from_box ? xr + i : reflect_clip(xi - 1 + i, p.sw);
    oy[i] = from_box ? (yr + i) * pitch : reflect_clip(yi - 1 + i, p.sh) * ssh;
  }
#pragma unroll 1
  for (int c = 0; c < NC; ++c) {
    const T* const sb = box + c * bh * pitch;
    const T* const gb = src_b + (long long)c * p.ssc;
    float v[4][4];
    if (from_box) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) v[r][i] = lds_f(sb + oy[r] + ox[i]);
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) v[r][i] = ld_f(gb + oy[r] + ox[i]);
    }
    float acc = 0.0f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float rs = __fmul_rn(v[r][0], cx[0]);
      rs = __fmaf_rn(v[r][1], cx[1], rs);
      rs = __fmaf_rn(v[r][2], cx[2], rs);
      rs = __fmaf_rn(v[r][3], cx[3], rs);
      acc = r == 0 ? __fmul_rn(rs, cy[0]) : __fmaf_rn(rs, cy[r], acc);
    }
    if (clip) acc = fminf(fmaxf(acc, clip_lo), clip_hi);
    const T o = OutCvt<T>::cvt(acc, p.flags);
    Pack<T, 1>::st(d + c * p.dsc, &o);
  }
}

template <int XF, typename T, int NC, bool EXACT, int MODE>
__global__ void __launch_bounds__(kL2Warps * 32,
                                  XF == EQB_PERS2EQUI ? 4 :  // (two scalar chains with IEEE divides live)
                                  MODE == EQB_BICUBIC ? (l2_ctas(sizeof(T), true) < 5 ? l2_ctas(sizeof(T), true) : 5)
                                  : sizeof(T) == 4    ? (l2_ctas(4) < 5 ? l2_ctas(4) : 5)
                                                      : (EXACT ? 5 : 7))
    remap_lean2_kernel(const __grid_constant__ Params p) {
  constexpr int ELT = (int)sizeof(T);
  constexpr bool CUBIC = MODE == EQB_BICUBIC;
  constexpr bool NEAR = MODE == EQB_NEAREST;  // (hull and boxes as for bilinear: the tap is
                                              // one of the two pixels around the coordinate)
  // taps of a pixel span [x - LO, x + HI] x [y - LO, y + HI] around its floored coordinate
  constexpr int LO = CUBIC ? 1 : 0, HI = CUBIC ? 2 : 1;
  constexpr int PER16 = 16 / ELT;
  constexpr int P = ELT == 1 ? 2 : 1;  // adjacent pixels per lane and row: a group is >= 2 bytes
  constexpr int BUF = l2_buf_bytes(ELT, MODE == EQB_BICUBIC);
  extern __shared__ __align__(1024) unsigned char smem_raw[];  // two box buffers
  __shared__ __align__(16) float2 nodes[kL2TileH][kL2Cols + 1];  // (ui, uj)
  __shared__ __align__(16) int4 headers[kL2MaxIters];
  __shared__ unsigned char nmask[XF == EQB_PERS2EQUI ? kL2TileH : 1][kL2Cols + 1];  // pers_node
  // nearest from interpolated coordinates: pixels deferred to the exact chain (see the fast path)
  __shared__ unsigned short tie_list[MODE == EQB_NEAREST && !EXACT ? kL2TieCap : 1];
  __shared__ int tie_count;
  __shared__ __align__(8) unsigned long long full_bar[2], empty_bar[2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.z;
  const int tile_y0 = blockIdx.y * kL2TileH;
  const int strip_x0 = blockIdx.x * p.iters * kL2TileW;
  const int ntiles = min(p.iters, (p.dw - strip_x0 + kL2TileW - 1) / kL2TileW);
  int cell = 0;
  if (XF == EQB_EQUI2CUBE) cell = strip_x0 / p.w_face;  // (a strip never straddles two cells)
  const bool background = XF == EQB_EQUI2CUBE && cell >= 6;

  if (tid == 0) {
    mbar_init(&full_bar[0], 1);
    mbar_init(&full_bar[1], 1);
    mbar_init(&empty_bar[0], kL2Warps);
    mbar_init(&empty_bar[1], kL2Warps);
    tie_count = 0;
  }

  // ---- pers2equi: a strip whose tiles are all outside the field of view (most of them: the
  // view covers 7.5 % of an 8K output at fov 90) needs no node, no header and no barrier but
  // this one: the corners of its tiles decide
  if constexpr (XF == EQB_PERS2EQUI && MODE == EQB_BICUBIC) {
    bool fill_strip = strip_x0 + ntiles * kL2TileW <= p.dw && tile_y0 + kL2TileH <= p.dh;
    if (fill_strip && tid < 32) {
      // lane 2k + j: corner column k (0 .. ntiles), row j (top / bottom) of the strip
      const int kc = lane >> 1, j = lane & 1;
      unsigned m = 0x1fu;
      if (kc <= ntiles)
        m = pers_corner_mask(p, b, strip_x0 + kc * kL2TileW,
                              min(tile_y0 + j * kL2TileH, p.dh - 1), kL2TileW, kL2TileH);
      // tile k: corners 2k, 2k + 1, 2k + 2, 2k + 3
      const unsigned m1 = __shfl_down_sync(0xffffffffu, m, 1);
      const unsigned m2 = __shfl_down_sync(0xffffffffu, m, 2);
      const unsigned m3 = __shfl_down_sync(0xffffffffu, m, 3);
      const bool tile_fill = (m & m1 & m2 & m3) != 0u;
      fill_strip = __all_sync(0xffffffffu, (lane & 1) || kc >= ntiles || tile_fill);
    }
    if (__syncthreads_and(fill_strip || tid >= 32)) {
      const bool clip0 = p.clip != nullptr;
      float lo0 = 0.f, hi0 = 0.f;
      if (clip0) { lo0 = __ldg(p.clip); hi0 = __ldg(p.clip + 1); }
      const int lx0 = lane & 15, ly0 = lane >> 4;
      constexpr int P0 = sizeof(T) == 1 ? 2 : 1;
      T* d = static_cast<T*>(p.dst) + (long long)b * p.dsb +
             (long long)(tile_y0 + warp * 4 + ly0) * p.dsh + strip_x0 + lx0 * P0;
      for (int k = 0; k < ntiles; ++k, d += kL2TileW)
        for (int c = 0; c < NC; ++c) {
          l2_store_pair<T>(d + c * p.dsc, 0.f, 0.f, clip0, lo0, hi0, p.flags);
          l2_store_pair<T>(d + c * p.dsc + 2 * p.dsh, 0.f, 0.f, clip0, lo0, hi0, p.flags);
        }
      return;
    }
  }

  // ---- prologue 1: the exact chain at the strip's nodes -------------------------------
  if (!background) {
    const int row = tid & (kL2TileH - 1);
    const int y = min(tile_y0 + row, p.dh - 1);
    RowCtx<XF> rc;
    row_setup<XF>(p, b, y, rc);
    const int ncols = ntiles * 4 + 3;
    constexpr int STEP = kL2Warps * 32 / kL2TileH;  // columns a thread advances by
    if constexpr (XF == EQB_PERS2EQUI) {
      for (int col = tid >> 4; col < ncols; col += STEP) {
        float ui, uj;
        unsigned mask;
        pers_node(p, rc, strip_x0 + kL2Step * (col - 1), ui, uj, mask);
        nodes[row][col] = make_float2(ui, uj);
        nmask[row][col] = (unsigned char)mask;
      }
    } else if constexpr (XF == EQB_CUBE2EQUI) {
      // coordinates from the 1-D tables (no transcendental); columns continue periodically
      for (int col = tid >> 4; col < ncols; col += STEP) {
        const int x = strip_x0 + kL2Step * (col - 1);
        const int xw = x < 0 ? x + p.dw : (x >= p.dw ? x - p.dw : x);
        float uj, ui;
        coords<XF, false>(p, rc, b, xw, y, 0, 0, uj, ui);
        nodes[row][col] = make_float2(ui, uj);
      }
    } else if constexpr (EXACT) {
      // two nodes per packed chain (coords2)
      for (int col = tid >> 4; col < ncols; col += 2 * STEP) {
        const int col2 = col + STEP < ncols ? col + STEP : col;
        float ui[2], uj[2];
        node_coords2<XF>(p, rc, strip_x0 + kL2Step * (col - 1), strip_x0 + kL2Step * (col2 - 1),
                         cell, ui, uj);
        nodes[row][col] = make_float2(ui[0], uj[0]);
        nodes[row][col2] = make_float2(ui[1], uj[1]);
      }
    } else {
      // (the packed chain is slower here: 1.41 vs 1.37 ms on the headline workload -- this
      // variant runs at 72 registers for 7 CTAs per SM)
      for (int col = tid >> 4; col < ncols; col += STEP) {
        float uj, ui;
        node_coords<XF>(p, rc, strip_x0 + kL2Step * (col - 1), y, cell, uj, ui);
        nodes[row][col] = make_float2(ui, uj);
      }
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
      // window of five nodes centred on tile-local node column c = 1 or 3 of row r: the
      // tile's own columns are 0..4, the windows also see -1 and 5 (interpolation stencils)
      const int r = lane & 15, c = 1 + 2 * (lane >> 4);
      const float2* np = &nodes[r][k * 4 + c - 1];  // (array index = node column + 1)
      float2 n[5];
#pragma unroll
      for (int i = 0; i < 5; ++i) n[i] = np[i];
      if constexpr (XF == EQB_PERS2EQUI) {
        // every node of the tile's own columns 0..4 outside the view by one common condition?
        const unsigned char* mp = &nmask[r][k * 4 + c - 1];
        const int o = c == 1 ? 1 : 0;
        const unsigned common = __reduce_and_sync(
            0xffffffffu, (unsigned)(mp[o] & mp[o + 1] & mp[o + 2] & mp[o + 3]));
        if (common != 0u && !(flags & L2_RAGGED)) {
          if (lane == 0) headers[k] = make_int4(0, 0, -1, flags | L2_FILL);
          continue;
        }
      }
      float ulo = fminf(fminf(n[1].x, n[2].x), n[3].x), uhi = fmaxf(fmaxf(n[1].x, n[2].x), n[3].x);
      float vlo = fminf(fminf(n[1].y, n[2].y), n[3].y), vhi = fmaxf(fmaxf(n[1].y, n[2].y), n[3].y);
      const float2 e = c == 1 ? n[4] : n[0];  // the end node that is a column of the tile
      ulo = fminf(ulo, e.x); uhi = fmaxf(uhi, e.x);
      vlo = fminf(vlo, e.y); vhi = fmaxf(vhi, e.y);
      const float2 f = c == 1 ? n[0] : n[4];  // the one outside
      const float alo = fminf(ulo, f.x), ahi = fmaxf(uhi, f.x);
      // (coordinates are >= 0: their bit patterns order like integers)
      const float umin = __int_as_float(__reduce_min_sync(0xffffffffu, __float_as_int(ulo)));
      const float umax = __int_as_float(__reduce_max_sync(0xffffffffu, __float_as_int(uhi)));
      const float vmin = __int_as_float(__reduce_min_sync(0xffffffffu, __float_as_int(vlo)));
      const float vmax = __int_as_float(__reduce_max_sync(0xffffffffu, __float_as_int(vhi)));
      const float amin = __int_as_float(__reduce_min_sync(0xffffffffu, __float_as_int(alo)));
      const float amax = __int_as_float(__reduce_max_sync(0xffffffffu, __float_as_int(ahi)));
      // the tile's own footprint wraps around the seam (no single box holds it) / some node of
      // its interpolation stencils lies across the seam (such tiles take the exact chain)
      const bool box_straddle = umax - umin >= 0.5f * p.swf;
      const bool straddle = amax - amin >= 0.5f * p.swf;
      // smoothness: 4th difference (the cubic's error is ~0.023 of it) and 2nd difference (a
      // pixel leaves the hull of its neighbouring nodes by at most 1/8 of it), from differences
      // against the centre node (no cancellation at coordinates of order 4096)
      bool ok = !straddle;
      {
        const float ax = n[0].x - n[2].x, bx = n[1].x - n[2].x, cx = n[3].x - n[2].x,
                    dx = n[4].x - n[2].x;
        const float ay = n[0].y - n[2].y, by = n[1].y - n[2].y, cy = n[3].y - n[2].y,
                    dy = n[4].y - n[2].y;
        const float d4x = (ax + dx) - 4.0f * (bx + cx), d4y = (ay + dy) - 4.0f * (by + cy);
        const float d2x = bx + cx, d2y = by + cy;
        ok = ok && fmaxf(fabsf(d4x), fabsf(d4y)) < 0.03f && fmaxf(fabsf(d2x), fabsf(d2y)) < 1.5f;
      }
      if (__all_sync(0xffffffffu, ok)) flags |= L2_SMOOTH;
      if constexpr (MODE == EQB_NEAREST && !EXACT) {
        // every node of the tile within the tie guard of a rounding tie (in x or in y): so is
        // (nearly) every pixel -- the tile takes the exact chain (see the fast path)
        bool nt = true;
        const float guard = l2_tie_guard(p);
#pragma unroll
        for (int i = 1; i <= 3; ++i) {
          const float tx = n[i].x - floorf(n[i].x), ty = n[i].y - floorf(n[i].y);
          nt = nt && (fabsf(tx - 0.5f) < guard || fabsf(ty - 0.5f) < guard);
        }
        if (__all_sync(0xffffffffu, nt)) flags |= L2_TIES;
      }
      if (straddle) flags |= L2_STRADDLE;
      // (bicubic: a tap within LO / HI pixels of the source border would be reflected)
      constexpr float BM = CUBIC ? 2.0f : 1.0f;
      if (umin < BM || umax > p.swm1f - (BM + 1.0f) || vmin < BM || vmax > p.shm1f - (BM + 1.0f))
        flags |= L2_BORDER;
      // box: top-left taps of the hull (with the footprint shift at the last column / row),
      // one more pixel right and down, margin 2
      const int bx0 = (int)fminf(floorf(fminf(umin, p.swm1f)), p.swm1f - 1.0f);
      const int bx1 = (int)fminf(floorf(fminf(umax, p.swm1f)), p.swm1f - 1.0f);
      const int by0 = (int)fminf(floorf(fminf(vmin, p.shm1f)), p.shm1f - 1.0f);
      const int by1 = (int)fminf(floorf(fminf(vmax, p.shm1f)), p.shm1f - 1.0f);
      const int gx0 = (bx0 - 2 - LO) & ~(PER16 - 1);  // 16-byte aligned (TMA requirement)
      const int gy0 = by0 - 2 - LO;
      const int need_w = bx1 + 3 + HI - gx0, need_h = by1 + 3 + HI - gy0;
      const unsigned fit = __ballot_sync(
          0xffffffffu, need_w <= (my_box & 0xffff) && need_h <= (my_box >> 16));
      const int m = (box_straddle || !p.use_tma) ? -1 : __ffs(fit) - 1;
      if (lane == 0) headers[k] = make_int4(gx0, gy0, m, flags);
    }
  }
  __syncthreads();

  // ---- tile loop ---------------------------------------------------------------------------
  // lane (lx, ly) of warp w owns, in rows 4w + ly and 4w + ly + 2, the pixels lx and lx + 16
  // (uint8: the pair 2 lx, 2 lx + 1).  Pixel j: row j >> 1, column slot j & 1.
  const int lx = lane & 15, ly = lane >> 4;
  const int row0 = warp * 4 + ly;
  const int y0q = tile_y0 + row0;
  const bool row0_ok = y0q < p.dh, row1_ok = y0q + 2 < p.dh;
  const unsigned tile_s = smem_u32(smem_raw);
  const unsigned full_s = smem_u32(&full_bar[0]), empty_s = smem_u32(&empty_bar[0]);
  const T* const src_b = static_cast<const T*>(p.src) + (long long)b * p.ssb;
  T* const dst_b = static_cast<T*>(p.dst) + (long long)b * p.dsb;
  const int ssh = (int)p.ssh;
  auto xloc = [&](int j) { return P == 1 ? lx + 16 * (j & 1) : 2 * lx + (j & 1); };
  // interpolation weights of this lane's pixels (offset t inside their 8-pixel interval)
  unsigned long long wt[P][3];
#pragma unroll
  for (int q = 0; q < P; ++q) {
    const int t = (lx * P + q) & 7;
#pragma unroll
    for (int j = 0; j < 3; ++j) wt[q][j] = pk2(g_l2_weights.w[t][j], g_l2_weights.w[t][j]);
  }

  // EXACT: the batch matrix and the row terms of the lane's two rows, loaded once per strip (in
  // the tile loop they were 22 dependent table loads per tile in front of every chain: the
  // fp32 kernel showed 31 % long-scoreboard stalls)
  RowCtx<XF> rows[2];
  if constexpr (EXACT) {
    row_setup<XF>(p, b, min(y0q, p.dh - 1), rows[0]);
    row_setup<XF>(p, b, min(y0q + 2, p.dh - 1), rows[1]);
  }

  // bicubic: the clamp of clip_output (the bilinear launches of this kernel carry none)
  const bool clip = CUBIC && p.clip != nullptr;
  float clip_lo = 0.f, clip_hi = 0.f;
  if (clip) { clip_lo = __ldg(p.clip); clip_hi = __ldg(p.clip + 1); }

  // The box of boxed tile n lives in buffer n & 1; phase parity of its barriers (n >> 1) & 1.
  // Thread 0 requests the box of the NEXT boxed tile while the current one is blended.
  auto issue = [&](int k, unsigned n) {  // thread 0 only
    const int4 h = headers[k];
    const unsigned buf = n & 1u;
    if (n >= 2) mbar_wait_backoff(empty_s + buf * 8u, ((n >> 1) - 1u) & 1u);
    mbar_expect_tx(&full_bar[buf], (unsigned)(p.tma_bw[h.z] * p.tma_bh[h.z] * NC * ELT));
    tma_load_box(smem_raw + buf * BUF, &p.tmaps[h.z], &full_bar[buf], h.x, h.y, 0,
                 p.tma_batched ? b : 0);
  };
  auto next_boxed = [&](int k) {
    for (++k; k < ntiles; ++k)
      if (headers[k].z >= 0) return k;
    return -1;
  };
  unsigned nbox = 0;
  if (tid == 0) {
    const int k0 = next_boxed(-1);
    if (k0 >= 0) issue(k0, 0);
  }

  for (int k = 0; k < ntiles; ++k) {
    const int4 hdr = headers[k];
    const int gx0 = hdr.x, gy0 = hdr.y, m = hdr.z, flags = hdr.w;
    const int tile_x0 = strip_x0 + k * kL2TileW;
    const bool boxed = m >= 0;
    T* d0 = dst_b + (long long)min(y0q, p.dh - 1) * p.dsh + tile_x0 + lx * P;
    if (XF == EQB_EQUI2CUBE)
      d0 = dst_b + (long long)min(y0q, p.dh - 1) * p.dsh + p.cell_off[cell] +
           (tile_x0 - cell * p.w_face) + lx * P;
    T* d1 = d0 + 2 * p.dsh;

    if (flags & L2_BACKGROUND) {  // dice background cell
      if (row0_ok) {
        const float z[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = 0; c < NC; ++c) l2_store<T>(d0 + c * p.dsc, d1 + c * p.dsc, z, row1_ok);
      }
      continue;
    }

    if constexpr (XF == EQB_PERS2EQUI) {
      if (flags & L2_FILL) {  // outside the field of view: out[mask] = 0, then the clamp
        if constexpr (CUBIC) {
          if (row0_ok)
            for (int c = 0; c < NC; ++c) {
              l2_store_pair<T>(d0 + c * p.dsc, 0.f, 0.f, clip, clip_lo, clip_hi, p.flags);
              if (row1_ok)
                l2_store_pair<T>(d1 + c * p.dsc, 0.f, 0.f, clip, clip_lo, clip_hi, p.flags);
            }
        }
        continue;
      }
    }

    const bool fast = boxed && (EXACT ? (flags & (L2_SMOOTH | L2_RAGGED | (CUBIC ? L2_BORDER : 0))) ==
                                            L2_SMOOTH
                                      : (flags & (L2_SMOOTH | L2_BORDER | L2_STRADDLE |
                                                  L2_RAGGED | L2_TIES)) == L2_SMOOTH);
    if (fast) {
      // ======== fast path: smooth tile with a box ==========================================
      float ixa[4], iya[4];
      if constexpr (!EXACT) {
        const unsigned long long mone = pk2(-1.0f, -1.0f);
#pragma unroll
        for (int g = 0; g < 4 / P; ++g) {
          // P == 1: pixel g; P == 2: the pair (2g, 2g + 1) shares its stencil
          const int r = row0 + 2 * ((g * P) >> 1);
          const int iv = P == 1 ? (lx >> 3) + 2 * (g & 1) : (lx >> 2);
          const unsigned long long* np =
              reinterpret_cast<const unsigned long long*>(&nodes[r][k * 4 + iv]);
          const unsigned long long n0 = np[0], n1 = np[1], n2 = np[2], n3 = np[3];
          const unsigned long long d0n = fma2(n1, mone, n0), d2n = fma2(n1, mone, n2),
                                   d3n = fma2(n1, mone, n3);
#pragma unroll
          for (int q = 0; q < P; ++q) {
            const unsigned long long u =
                fma2(wt[q][2], d3n, fma2(wt[q][1], d2n, fma2(wt[q][0], d0n, n1)));
            unpk2(u, ixa[g * P + q], iya[g * P + q]);
          }
        }
      } else {
#pragma unroll
        for (int gy = 0; gy < 2; ++gy) {
          const RowCtx<XF>& rc = rows[gy];
          const int xa = tile_x0 + xloc(2 * gy), xb = tile_x0 + xloc(2 * gy + 1);
          if constexpr (XF == EQB_PERS2EQUI || XF == EQB_CUBE2EQUI) {
            // scalar chains (pers2equi: every pixel of a smooth interior tile lies inside the
            // field of view)
            const int y = min(y0q + 2 * gy, p.dh - 1);
            float uj, ui;
            coords<XF, false>(p, rc, b, xa, y, 0, 0, uj, ui);
            ixa[2 * gy] = to_index(ui, p.inv_wm1, p.swm1f);
            iya[2 * gy] = to_index(uj, p.inv_hm1, p.shm1f);
            coords<XF, false>(p, rc, b, xb, y, 0, 0, uj, ui);
            ixa[2 * gy + 1] = to_index(ui, p.inv_wm1, p.swm1f);
            iya[2 * gy + 1] = to_index(uj, p.inv_hm1, p.shm1f);
            continue;
          }
          float rca = 0.f, rcb = 0.f;
          if (XF == EQB_EQUI2CUBE) {
            rca = __ldg(p.tx + (xa - cell * p.w_face));
            rcb = __ldg(p.tx + (xb - cell * p.w_face));
          }
          float i2[2], j2[2];
          coords2<XF, true>(p, rc, xa, xb, cell, rca, rcb, i2, j2);
          ixa[2 * gy] = i2[0]; ixa[2 * gy + 1] = i2[1];
          iya[2 * gy] = j2[0]; iya[2 * gy + 1] = j2[1];
        }
      }
      if constexpr (NEAR && !EXACT) {
        // ---- nearest from interpolated coordinates: the tap is a rounding decision, so every
        // pixel whose interpolated coordinate lies within l2_tie_guard() of a rounding tie --
        // more than twice the measured worst-case distance between the interpolated and the
        // exact coordinate (1.7e-3 px at 4K, test_fast_coordinates_error_in_pixels) -- gets the exact
        // chain; all others round like it (test_nearest_4k_tie_guard_equals_the_exact_chain).
        // Such pixels are 1.6 % of a typical view.  A warp that ran the chain for them at once
        // would still pay all of its ~150 instructions for one or two busy lanes (measured:
        // 1.50 ms against 1.64 for the exact kernel), so they are DEFERRED: their positions
        // go to a per-strip list in shared memory, the tile stores their provisional values
        // like everybody else's, and after the last tile all 128 threads evaluate the exact
        // chain for one listed pixel each and overwrite it.  Tiles whose NODES all sit on a
        // tie (identity, pure yaw: every row coordinate is y + 0.5) are flagged L2_TIES by the
        // prologue and never come here (general path, exact chain); a tile with more than 32
        // such pixels per warp, or a full list, takes the packed exact chain below.
        const float kTieGuard = l2_tie_guard(p);
        bool tie[4];
        unsigned mk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float tx = ixa[j] - floorf(ixa[j]), ty = iya[j] - floorf(iya[j]);
          tie[j] = fabsf(tx - 0.5f) < kTieGuard || fabsf(ty - 0.5f) < kTieGuard;
          mk[j] = __ballot_sync(0xffffffffu, tie[j]);
        }
        if (mk[0] | mk[1] | mk[2] | mk[3]) {  // (warp-uniform)
          const int c0 = __popc(mk[0]), c1 = c0 + __popc(mk[1]), c2 = c1 + __popc(mk[2]);
          const int total = c2 + __popc(mk[3]);
          int base = kL2TieCap;
          if (total <= 32) {
            if (lane == 0) base = atomicAdd(&tie_count, total);
            base = __shfl_sync(0xffffffffu, base, 0);
          }
          if (base + total <= kL2TieCap) {
            // deferred: (tile, row, column) of each such pixel
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (tie[j]) {
                const int off = j == 0 ? 0 : (j == 1 ? c0 : (j == 2 ? c1 : c2));
                const int idx = base + off + __popc(mk[j] & ((1u << lane) - 1u));
                tie_list[idx] = (unsigned short)((k << 9) | ((row0 + 2 * (j >> 1)) << 5) | xloc(j));
              }
          } else {
            // the list is full, or most of the tile sits on a tie: exact chain for the tile
            if (total <= 32 && base < kL2TieCap)  // (slots reserved above stay unused)
              for (int i = base + lane; i < kL2TieCap; i += 32) tie_list[i] = 0xffffu;
#pragma unroll 1
            for (int gy = 0; gy < 2; ++gy) {
              RowCtx<XF> rcx;
              row_setup<XF>(p, b, min(y0q + 2 * gy, p.dh - 1), rcx);
              const int xa = tile_x0 + xloc(2 * gy), xb = tile_x0 + xloc(2 * gy + 1);
              float rca = 0.f, rcb = 0.f;
              if (XF == EQB_EQUI2CUBE) {
                rca = __ldg(p.tx + (xa - cell * p.w_face));
                rcb = __ldg(p.tx + (xb - cell * p.w_face));
              }
              float i2[2], j2[2];
              coords2<XF, true>(p, rcx, xa, xb, cell, rca, rcb, i2, j2);
              ixa[2 * gy] = i2[0]; ixa[2 * gy + 1] = i2[1];
              iya[2 * gy] = j2[0]; iya[2 * gy + 1] = j2[1];
            }
          }
        }
      }
      L2Px px[4];
      unsigned pa[4];  // bicubic: byte address of the top-left tap (x - 1, y - 1), channel 0
      float fx[4], fy[4];
      {
        const int pitch = p.tma_bw[m];
        const float pitchf = (float)pitch;
        // byte address = ((yf - gy0) * pitch + (xf - gx0)) * ELT + buffer; yf * pitch + xf is
        // an exact integer below 2^24
        const int base =
            (int)(tile_s + (nbox & 1u) * BUF) - ((gy0 + LO) * pitch + gx0 + LO) * ELT;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if constexpr (NEAR) {  // nearbyint, like ATen (taps_setup)
            const int xi = __float2int_rn(ixa[j]), yi = __float2int_rn(iya[j]);
            pa[j] = (unsigned)((yi * pitch + xi) * ELT + base);
            continue;
          }
          float xf = floorf(ixa[j]), yf = floorf(iya[j]);
          if (EXACT && !CUBIC) {  // footprint shifted to x0 = W-2 at ix == W-1 (taps_setup)
            xf = fminf(xf, p.swm1f - 1.0f);
            yf = fminf(yf, p.shm1f - 1.0f);
          }
          fx[j] = __fsub_rn(ixa[j], xf);
          fy[j] = __fsub_rn(iya[j], yf);
          const unsigned a = (unsigned)((int)fmaf(yf, pitchf, xf) * ELT + base);
          if constexpr (CUBIC) pa[j] = a;
          else l2_px_setup<T>(a, px[j]);
        }
      }
      // request the next box, then wait for this one (requesting it before this tile's
      // coordinates instead changes nothing: 1.384 vs 1.371 ms on the headline workload)
      if (tid == 0) {
        const int k2 = next_boxed(k);
        if (k2 >= 0) issue(k2, nbox + 1);
      }
      __syncwarp();
      mbar_wait_backoff(full_s + (nbox & 1u) * 8u, (nbox >> 1) & 1u);
      if constexpr (NEAR) {
        if (row0_ok) {
          const unsigned plane = (unsigned)(p.tma_bw[m] * p.tma_bh[m] * ELT);
          l2_nearest_pair<T, NC>(pa[0], pa[1], plane, d0, p.dsc);
          if (row1_ok) l2_nearest_pair<T, NC>(pa[2], pa[3], plane, d1, p.dsc);
        }
      } else if constexpr (CUBIC) {
        if (row0_ok) {
          const unsigned row = (unsigned)(p.tma_bw[m] * ELT);
          const unsigned plane = row * (unsigned)p.tma_bh[m];
          l2_bicubic_pair<T, NC>(pa[0], pa[1], fx[0], fx[1], fy[0], fy[1], row, plane, d0, p.dsc,
                                 clip, clip_lo, clip_hi, p.flags);
          if (row1_ok)
            l2_bicubic_pair<T, NC>(pa[2], pa[3], fx[2], fx[3], fy[2], fy[3], row, plane, d1,
                                   p.dsc, clip, clip_lo, clip_hi, p.flags);
        }
      } else if (row0_ok) {
        auto blend_m = [&](auto mc) {
          constexpr int M = decltype(mc)::value;
          l2_blend<T, M, NC>(px, fx, fy, d0, d1, p.dsc, row1_ok);
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
          default: blend_m(cuda::std::integral_constant<int, 10>{}); break;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_s + (nbox & 1u) * 8u);
      ++nbox;
      continue;
    }

    // ======== general path: taps from the box where they lie inside it, else from global
    // memory.  Coordinates: interpolated like on the fast path where the tile is smooth and
    // interior (a footprint too wide for any box -- high source latitudes -- is usually still
    // smooth), the exact chain elsewhere (source poles, source border, seam, EXACT).
    if (boxed) {
      if (tid == 0) {
        const int k2 = next_boxed(k);
        if (k2 >= 0) issue(k2, nbox + 1);
      }
      __syncwarp();
      mbar_wait_backoff(full_s + (nbox & 1u) * 8u, (nbox >> 1) & 1u);
    }
    {
      const int pitch = boxed ? p.tma_bw[m] : 0;
      const int bh = boxed ? p.tma_bh[m] : 0;
      const T* const box = reinterpret_cast<const T*>(smem_raw + (nbox & 1u) * BUF);
      // (nearest: the exact chain here -- the tie guard lives on the fast path only)
      const bool interp =
          !EXACT && !NEAR && (flags & (L2_SMOOTH | L2_BORDER | L2_STRADDLE)) == L2_SMOOTH;
      float ixa[4], iya[4];
      bool inval[4] = {false, false, false, false};  // pers2equi: outside the field of view
      if (interp) {
        const unsigned long long mone = pk2(-1.0f, -1.0f);
#pragma unroll
        for (int g = 0; g < 4 / P; ++g) {
          const int r = row0 + 2 * ((g * P) >> 1);
          const int iv = P == 1 ? (lx >> 3) + 2 * (g & 1) : (lx >> 2);
          const unsigned long long* np =
              reinterpret_cast<const unsigned long long*>(&nodes[r][k * 4 + iv]);
          const unsigned long long n0 = np[0], n1 = np[1], n2 = np[2], n3 = np[3];
          const unsigned long long d0n = fma2(n1, mone, n0), d2n = fma2(n1, mone, n2),
                                   d3n = fma2(n1, mone, n3);
#pragma unroll
          for (int q = 0; q < P; ++q) {
            const unsigned long long u =
                fma2(wt[q][2], d3n, fma2(wt[q][1], d2n, fma2(wt[q][0], d0n, n1)));
            unpk2(u, ixa[g * P + q], iya[g * P + q]);
          }
        }
      } else {
        RowCtx<XF> rc;
#pragma unroll 1
        for (int gy = 0; gy < 2; ++gy) {
          const int yq = min(y0q + 2 * gy, p.dh - 1);
          row_setup<XF>(p, b, yq, rc);
          const int xa = min(tile_x0 + xloc(2 * gy), p.dw - 1);
          const int xb = min(tile_x0 + xloc(2 * gy + 1), p.dw - 1);
          if constexpr (XF == EQB_PERS2EQUI || XF == EQB_CUBE2EQUI) {
            float uj, ui;
            inval[2 * gy] = coords<XF, false>(p, rc, b, xa, yq, 0, 0, uj, ui);
            ixa[2 * gy] = to_index(ui, p.inv_wm1, p.swm1f);
            iya[2 * gy] = to_index(uj, p.inv_hm1, p.shm1f);
            inval[2 * gy + 1] = coords<XF, false>(p, rc, b, xb, yq, 0, 0, uj, ui);
            ixa[2 * gy + 1] = to_index(ui, p.inv_wm1, p.swm1f);
            iya[2 * gy + 1] = to_index(uj, p.inv_hm1, p.shm1f);
            continue;
          }
          float rca = 0.f, rcb = 0.f;
          if (XF == EQB_EQUI2CUBE) {
            rca = __ldg(p.tx + (xa - cell * p.w_face));
            rcb = __ldg(p.tx + (xb - cell * p.w_face));
          }
          float i2[2], j2[2];
          coords2<XF, true>(p, rc, xa, xb, cell, rca, rcb, i2, j2);
          ixa[2 * gy] = i2[0]; ixa[2 * gy + 1] = i2[1];
          iya[2 * gy] = j2[0]; iya[2 * gy + 1] = j2[1];
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int yq = y0q + 2 * (j >> 1), x = tile_x0 + xloc(j);
        if (yq >= p.dh || x >= p.dw) continue;
        const float ix = ixa[j], iy = iya[j];
        T* d = ((j >> 1) ? d1 : d0) + (xloc(j) - lx * P);
        if constexpr (CUBIC) {
          if (XF == EQB_PERS2EQUI && inval[j]) {
            const float fill = clip ? fminf(fmaxf(0.0f, clip_lo), clip_hi) : 0.0f;
            const T o = OutCvt<T>::cvt(fill, p.flags);
            for (int c = 0; c < NC; ++c) Pack<T, 1>::st(d + c * p.dsc, &o);
            continue;
          }
          l2_bicubic_general<T, NC>(p, ix, iy, boxed, box, pitch, bh, gx0, gy0, src_b, d, clip,
                                    clip_lo, clip_hi);
          continue;
        }
        if constexpr (NEAR) {
          const int xi = __float2int_rn(ix), yi = __float2int_rn(iy);
          const int xr = xi - gx0, yr = yi - gy0;
          const bool from_box = boxed && xr >= 0 && xr < pitch && yr >= 0 && yr < bh;
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const T o = from_box ? box[(c * bh + yr) * pitch + xr]
                                 : __ldg(src_b + (long long)c * p.ssc + (yi * ssh + xi));
            Pack<T, 1>::st(d + c * p.dsc, &o);
          }
          continue;
        }
        const float xf = fminf(floorf(ix), p.swm1f - 1.0f), yf = fminf(floorf(iy), p.shm1f - 1.0f);
        const float bx = __fsub_rn(ix, xf), by = __fsub_rn(iy, yf);
        const float ax = __fsub_rn(1.0f, bx), ay = __fsub_rn(1.0f, by);
        const float w_nw = __fmul_rn(ax, ay), w_ne = __fmul_rn(bx, ay);
        const float w_sw = __fmul_rn(ax, by), w_se = __fmul_rn(bx, by);
        const int xi = (int)xf, yi = (int)yf;
        const int xr = xi - gx0, yr = yi - gy0;
        const bool from_box = boxed && xr >= 0 && xr <= pitch - 2 && yr >= 0 && yr <= bh - 2;
        float v[NC][4];
        if (from_box) {
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const T* q0 = box + (c * bh + yr) * pitch + xr;
            v[c][0] = lds_f(q0); v[c][1] = lds_f(q0 + 1);
            v[c][2] = lds_f(q0 + pitch); v[c][3] = lds_f(q0 + pitch + 1);
          }
        } else {
#pragma unroll
          for (int c = 0; c < NC; ++c) {  // (all loads of the pixel in flight together)
            const T* q0 = src_b + (long long)c * p.ssc + (yi * ssh + xi);
            if constexpr (ELT == 4) {
              v[c][0] = ld_f(q0); v[c][1] = ld_f(q0 + 1);
              v[c][2] = ld_f(q0 + ssh); v[c][3] = ld_f(q0 + ssh + 1);
            } else {
              ldg_pair<T>(q0, v[c][0], v[c][1]);
              ldg_pair<T>(q0 + ssh, v[c][2], v[c][3]);
            }
          }
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const T o = OutCvt<T>::cvt(blend4(v[c][0], v[c][1], v[c][2], v[c][3], w_nw, w_ne, w_sw,
                                            w_se), p.flags);
          Pack<T, 1>::st(d + c * p.dsc, &o);
        }
      }
    }
    if (boxed) {
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_s + (nbox & 1u) * 8u);
      ++nbox;
    }
  }

  // ---- nearest from interpolated coordinates: the deferred pixels of the strip, one per
  // thread: exact chain, tap from global memory, overwrite (the barrier orders this store after
  // the provisional one of the tile loop)
  if constexpr (NEAR && !EXACT) {
    __syncthreads();
    const int n = min(tie_count, kL2TieCap);
    for (int i = tid; i < n; i += kL2Warps * 32) {
      const unsigned e = tie_list[i];
      if (e == 0xffffu) continue;
      const int yy = tile_y0 + (int)((e >> 5) & 15u);
      const int xx = strip_x0 + (int)(e >> 9) * kL2TileW + (int)(e & 31u);
      RowCtx<XF> rcx;
      row_setup<XF>(p, b, yy, rcx);
      float uj, ui;
      coords<XF, false>(p, rcx, b, xx, yy, cell, xx - cell * p.w_face, uj, ui);
      const int xi = __float2int_rn(to_index(ui, p.inv_wm1, p.swm1f));
      const int yi = __float2int_rn(to_index(uj, p.inv_hm1, p.shm1f));
      T* d = dst_b + (long long)yy * p.dsh + xx;
      if (XF == EQB_EQUI2CUBE)
        d = dst_b + (long long)yy * p.dsh + p.cell_off[cell] + (xx - cell * p.w_face);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const T o = __ldg(src_b + (long long)c * p.ssc + (yi * ssh + xi));
        Pack<T, 1>::st(d + c * p.dsc, &o);
      }
    }
  }
}

}  // namespace eqb
