// eqb_equi2cube2equi: equirect -> cubemap -> equirect in ONE launch.
//
// The reference runs equi2cube (equilib/equi2cube/torch.py:189-251) and then cube2equi
// (equilib/cube2equi/torch.py:324-357) and pays one full write and one full read of the
// cubemap in between (plus its horizon <-> dice / dict copies, cube2equi/torch.py:22-65,
// equi2cube/torch.py:58-76).  Here the cubemap never exists: an output pixel evaluates the
// cube2equi coordinates from the 1-D tables (bit-exact), and every cube pixel its sampler taps
// -- one for nearest, four for bilinear, resolved on the virtual horizon strip, so taps bleed
// into the neighbouring face in horizon order exactly as in the reference
// (cube2equi/torch.py:228-242) -- is produced on the fly by the equi2cube chain (ray of the
// face pixel, rotation, asin / atan2, sampling of the source equirect), rounded to the image
// type as the stored cubemap would have been.  Same operations in the same order as the two
// launches, so the result equals cube2equi(equi2cube(x)) of this library bit for bit
// (tests/test_gpu_parity.py::test_fused_cube_chain_equals_two_launches) -- the two cube pixels
// of a tap row share one packed chain (coords2) when they lie in one face.
//
// Cost: four coordinate chains per output pixel instead of 6 w^2 / (H W) = 0.75, but no
// cubemap traffic and no second launch; timed next to the two launches in DESIGN.md.
#include "remap_kernels.cuh"

namespace eqb {

struct ChainParams {
  Params e;  // stage 1, equi2cube: src = the equirect, tx = rng, mats; dst unused
  Params c;  // stage 2, cube2equi: tables, dst = the output equirect; src unused
};

template <typename T> __device__ __forceinline__ float as_stored(float v, unsigned flags) {
  return (float)OutCvt<T>::cvt(v, flags);
}
template <> __device__ __forceinline__ float as_stored<__half>(float v, unsigned) {
  return __half2float(__float2half_rn(v));
}
template <> __device__ __forceinline__ float as_stored<__nv_bfloat16>(float v, unsigned) {
  return __bfloat162float(__float2bfloat16_rn(v));
}

// Value of the cube pixel whose equi2cube sampler coordinates are (ix, iy), every channel,
// as the cubemap would have stored it.
template <int MODE, typename T>
__device__ __forceinline__ void cube_value(const Params& pe, const T* src_b, float ix, float iy,
                                           float (&v)[4]) {
  const int ssh = (int)pe.ssh;
  if (MODE == EQB_NEAREST) {
    const int off = __float2int_rn(iy) * ssh + __float2int_rn(ix);  // nearbyint (taps_setup)
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < pe.C) v[c] = ld_f(src_b + (long long)c * pe.ssc + off);
  } else {
    // taps_setup<BILINEAR>: footprint shifted to x0 = W-2 at ix == W-1
    const float x0 = fminf(floorf(ix), pe.swm1f - 1.0f), y0 = fminf(floorf(iy), pe.shm1f - 1.0f);
    const float ax = __fsub_rn(x0 + 1.0f, ix), bx = __fsub_rn(ix, x0);
    const float ay = __fsub_rn(y0 + 1.0f, iy), by = __fsub_rn(iy, y0);
    const float w_nw = __fmul_rn(ax, ay), w_ne = __fmul_rn(bx, ay);
    const float w_sw = __fmul_rn(ax, by), w_se = __fmul_rn(bx, by);
    const int off = (int)y0 * ssh + (int)x0;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < pe.C) {
        const T* q0 = src_b + (long long)c * pe.ssc + off;
        const T* q1 = q0 + ssh;
        v[c] = as_stored<T>(
            blend4(ld_f(q0), ld_f(q0 + 1), ld_f(q1), ld_f(q1 + 1), w_nw, w_ne, w_sw, w_se),
            pe.flags);
      }
  }
}

// Cube pixels (xa, y) and (xb, y) of the virtual horizon strip (columns 0 .. 6w-1).
template <int MODE, typename T>
__device__ __forceinline__ void cube_row_pair(const Params& pe, RowCtx<EQB_EQUI2CUBE>& re,
                                              const T* src_b, int xa, int xb, int y,
                                              float (&va)[4], float (&vb)[4]) {
  re.r0 = __ldg(pe.tx + y);
  const int fa = (int)(((float)xa + 0.5f) * pe.inv_wface);
  const int fb = (int)(((float)xb + 0.5f) * pe.inv_wface);
  const int la = xa - fa * pe.w_face, lb = xb - fb * pe.w_face;
  float ix[2], iy[2];
  if (fa == fb) {
    coords2<EQB_EQUI2CUBE, true>(pe, re, 0, 0, fa, __ldg(pe.tx + la), __ldg(pe.tx + lb), ix, iy);
  } else {  // the tap pair straddles two faces (bleed in horizon order)
    float uj, ui;
    coords<EQB_EQUI2CUBE, false>(pe, re, 0, 0, 0, fa, la, uj, ui);
    ix[0] = to_index(ui, pe.inv_wm1, pe.swm1f);
    iy[0] = to_index(uj, pe.inv_hm1, pe.shm1f);
    coords<EQB_EQUI2CUBE, false>(pe, re, 0, 0, 0, fb, lb, uj, ui);
    ix[1] = to_index(ui, pe.inv_wm1, pe.swm1f);
    iy[1] = to_index(uj, pe.inv_hm1, pe.shm1f);
  }
  cube_value<MODE, T>(pe, src_b, ix[0], iy[0], va);
  cube_value<MODE, T>(pe, src_b, ix[1], iy[1], vb);
}

template <int MODE, typename T>
__global__ void __launch_bounds__(256) chain_kernel(const __grid_constant__ ChainParams cp) {
  const Params& pe = cp.e;
  const Params& pc = cp.c;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int b = blockIdx.z;
  if (x >= pc.dw || y >= pc.dh) return;

  // stage 2 coordinates: cube2equi/torch.py:172-245 from the 1-D tables
  RowCtx<EQB_CUBE2EQUI> rc;
  row_setup<EQB_CUBE2EQUI>(pc, b, y, rc);
  float uj, ui;
  coords<EQB_CUBE2EQUI, false>(pc, rc, b, x, y, 0, 0, uj, ui);
  const float ix = to_index(ui, pc.inv_wm1, pc.swm1f);
  const float iy = to_index(uj, pc.inv_hm1, pc.shm1f);

  RowCtx<EQB_EQUI2CUBE> re;
  row_setup<EQB_EQUI2CUBE>(pe, b, 0, re);  // the batch matrix (r0 is set per cube row)
  const T* const src_b = static_cast<const T*>(pe.src) + (long long)b * pe.ssb;
  T* dst = static_cast<T*>(pc.dst) + (long long)b * pc.dsb + (long long)y * pc.dsh + x;

  float o[4] = {0.f, 0.f, 0.f, 0.f};
  if (MODE == EQB_NEAREST) {
    const int cx = __float2int_rn(ix), cy = __float2int_rn(iy);
    re.r0 = __ldg(pe.tx + cy);
    const int f = (int)(((float)cx + 0.5f) * pe.inv_wface);
    float vj, vi;
    coords<EQB_EQUI2CUBE, false>(pe, re, 0, 0, 0, f, cx - f * pe.w_face, vj, vi);
    cube_value<MODE, T>(pe, src_b, to_index(vi, pe.inv_wm1, pe.swm1f),
                        to_index(vj, pe.inv_hm1, pe.shm1f), o);
  } else {
    // taps_setup<BILINEAR, FACES>: x0 = floor(ix); the +1 taps are clamped to the strip (they
    // leave it only where their weight is exactly 0)
    const float x0 = floorf(ix), y0 = floorf(iy);
    const float ax = __fsub_rn(x0 + 1.0f, ix), bx = __fsub_rn(ix, x0);
    const float ay = __fsub_rn(y0 + 1.0f, iy), by = __fsub_rn(iy, y0);
    const float w_nw = __fmul_rn(ax, ay), w_ne = __fmul_rn(bx, ay);
    const float w_sw = __fmul_rn(ax, by), w_se = __fmul_rn(bx, by);
    const int xi0 = (int)x0, yi0 = (int)y0;
    const int xi1 = min(xi0 + 1, pc.sw - 1), yi1 = min(yi0 + 1, pc.sh - 1);
    float v_nw[4], v_ne[4], v_sw[4], v_se[4];
    cube_row_pair<MODE, T>(pe, re, src_b, xi0, xi1, yi0, v_nw, v_ne);
    cube_row_pair<MODE, T>(pe, re, src_b, xi0, xi1, yi1, v_sw, v_se);
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < pc.C) o[c] = blend4(v_nw[c], v_ne[c], v_sw[c], v_se[c], w_nw, w_ne, w_sw, w_se);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c)
    if (c < pc.C) {
      const T t = OutCvt<T>::cvt(o[c], pc.flags);
      Pack<T, 1>::st(dst + (long long)c * pc.dsc, &t);
    }
}

template <int MODE>
static cudaError_t launch_chain_mode(const ChainParams& cp, int dtype, cudaStream_t stream) {
  const dim3 grid((cp.c.dw + 31) / 32, (cp.c.dh + 7) / 8, cp.c.B);
  switch (dtype) {
    case EQB_U8: chain_kernel<MODE, uint8_t><<<grid, 256, 0, stream>>>(cp); break;
    case EQB_F16: chain_kernel<MODE, __half><<<grid, 256, 0, stream>>>(cp); break;
    case EQB_BF16: chain_kernel<MODE, __nv_bfloat16><<<grid, 256, 0, stream>>>(cp); break;
    case EQB_F32: chain_kernel<MODE, float><<<grid, 256, 0, stream>>>(cp); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_chain(const Params& e, const Params& c, int mode, int dtype,
                         cudaStream_t stream) {
  ChainParams cp;
  cp.e = e;
  cp.c = c;
  if (mode == EQB_NEAREST) return launch_chain_mode<EQB_NEAREST>(cp, dtype, stream);
  if (mode == EQB_BILINEAR) return launch_chain_mode<EQB_BILINEAR>(cp, dtype, stream);
  return cudaErrorNotSupported;
}

}  // namespace eqb
