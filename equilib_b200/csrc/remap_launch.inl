// Included by remap_xf<N>.cu with EQB_XF defined: instantiates the kernels of one
// transform and its mode/dtype/PX/warp-shape dispatch (one translation unit per
// transform so the six compile in parallel).
#include "remap_kernels.cuh"
#if EQB_XF <= 4
#include "remap_lean2.cuh"
#endif

namespace eqb {

// Pixels per thread (= store width) instantiated per dtype and mode; must agree
// with vec_px() in capi.cu.  Bicubic keeps one pixel per thread (16 taps of state).
template <typename T, int MODE> struct VecPX { static constexpr int value = 2; };
template <> struct VecPX<uint8_t, EQB_NEAREST> { static constexpr int value = 4; };
template <> struct VecPX<uint8_t, EQB_BILINEAR> { static constexpr int value = 4; };
template <typename T> struct VecPX<T, EQB_BICUBIC> { static constexpr int value = 1; };

template <int XF, int MODE, typename T, int PX, int WS>
static cudaError_t launch_one(Params p, cudaStream_t stream) {
  constexpr int LW = 32 >> WS, LH = 1 << WS;
  const int tile_w = LW * PX, tile_h = LH * kWarps;
  const long long tiles_x = (p.dw + tile_w - 1) / tile_w;
  const long long ctas1 = tiles_x * ((p.dh + tile_h - 1) / tile_h) * p.B;
  // walk several tiles per thread once the grid is large enough to fill the GPU
  // many times over; small launches keep one tile per thread for parallelism
  p.iters = ctas1 >= 148ll * 64 ? 4 : 1;
  const dim3 grid((unsigned)((tiles_x + p.iters - 1) / p.iters),
                  (unsigned)((p.dh + tile_h - 1) / tile_h), (unsigned)p.B);
  remap_kernel<XF, MODE, T, PX, WS><<<grid, kWarps * 32, 0, stream>>>(p);
  return cudaGetLastError();
}

template <int XF, int MODE, typename T, int PX>
static cudaError_t launch_ws(const Params& p, int ws, cudaStream_t stream) {
  switch (ws) {
    case 0: return launch_one<XF, MODE, T, PX, 0>(p, stream);
    case 1: return launch_one<XF, MODE, T, PX, 1>(p, stream);
    default: return launch_one<XF, MODE, T, PX, 2>(p, stream);
  }
}

template <int XF, int MODE, typename T>
static cudaError_t launch_px(const Params& p, int px, int ws, cudaStream_t stream) {
  if (px == 1) return launch_ws<XF, MODE, T, 1>(p, ws, stream);
  // cube2equi carries one offset and face id per tap: 4 uint8 pixels per thread spill
  // pers2equi keeps vector stores in bicubic mode (its output is mostly fill)
  constexpr int V = (XF == EQB_CUBE2EQUI && MODE == EQB_BILINEAR && sizeof(T) == 1) ? 2
                    : (XF == EQB_PERS2EQUI && MODE == EQB_BICUBIC)
                        ? VecPX<T, EQB_BILINEAR>::value
                        : VecPX<T, MODE>::value;
  return launch_ws<XF, MODE, T, V>(p, ws, stream);
}

template <int XF, int MODE>
static cudaError_t launch_dtype(const Params& p, int dtype, int px, int ws, cudaStream_t stream) {
  switch (dtype) {
    case EQB_U8: return launch_px<XF, MODE, uint8_t>(p, px, ws, stream);
    case EQB_F16: return launch_px<XF, MODE, __half>(p, px, ws, stream);
    case EQB_BF16: return launch_px<XF, MODE, __nv_bfloat16>(p, px, ws, stream);
    case EQB_F32: return launch_px<XF, MODE, float>(p, px, ws, stream);
  }
  return cudaErrorInvalidValue;
}

template <>
cudaError_t launch_remap<EQB_XF>(const Params& p, int mode, int dtype, int px, int ws,
                                 cudaStream_t stream) {
  switch (mode) {
    case EQB_NEAREST: return launch_dtype<EQB_XF, EQB_NEAREST>(p, dtype, px, ws, stream);
    case EQB_BILINEAR: return launch_dtype<EQB_XF, EQB_BILINEAR>(p, dtype, px, ws, stream);
    case EQB_BICUBIC: return launch_dtype<EQB_XF, EQB_BICUBIC>(p, dtype, px, ws, stream);
  }
  return cudaErrorInvalidValue;
}

// (the lean kernel is not instantiated for 4-byte pixels)
template <typename T> struct LeanType { using type = T; };
template <> struct LeanType<float> { using type = __nv_bfloat16; };

template <int XF, int MODE, typename T, int PX, bool FAST, bool CLIP>
static cudaError_t launch_staged_clip(Params p, cudaStream_t stream) {
  using G = StagedGeom<PX>;
  const int smem = staged_smem_bytes((int)sizeof(T));
  // the opt-in is per device (a process may drive several GPUs): set it on every launch
  // that needs more than the 48 KB default
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(remap_staged_kernel<XF, MODE, T, PX, FAST, CLIP>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  p.smem_elems = smem / (int)sizeof(T);
  const long long tiles_x = (p.dw + G::tile_w - 1) / G::tile_w;
  const long long tiles_y = (p.dh + G::tile_h - 1) / G::tile_h;
  p.iters = tiles_x * tiles_y * p.B >= 148ll * 64 ? 4 : 1;
  const dim3 grid((unsigned)((tiles_x + p.iters - 1) / p.iters), (unsigned)tiles_y, (unsigned)p.B);
  remap_staged_kernel<XF, MODE, T, PX, FAST, CLIP><<<grid, kStWarps * 32, smem, stream>>>(p);
  return cudaGetLastError();
}

// Lean variant of <BILINEAR, PX 4, FAST, no clip> for three channels of 1- or 2-byte pixels
// through TMA (the headline case).
template <int XF, typename T, int NC>
static cudaError_t launch_lean_nc(Params p, cudaStream_t stream) {
  const long long tiles_x = (p.dw + 63) / 64;
  const long long tiles_y = (p.dh + kLeanTileH - 1) / kLeanTileH;
  // tiles a CTA walks along x: 16 where that leaves >= 8 waves of CTAs in the grid (the
  // per-CTA setup -- tables, matrix, row terms -- is ~180 instructions per thread)
  const long long tiles = tiles_x * tiles_y * p.B;
  const long long wave = 148ll * 2048 / (kLeanWarps * 32);
  p.iters = tiles >= wave * 8 * 16 ? 16 : tiles >= wave * 16 ? 4 : 1;
  const dim3 grid((unsigned)((tiles_x + p.iters - 1) / p.iters), (unsigned)tiles_y, (unsigned)p.B);
  remap_lean_kernel<XF, T, 0, NC><<<grid, kLeanWarps * 32, lean_smem_bytes(), stream>>>(p);
  return cudaGetLastError();
}

template <int XF, typename T>
static cudaError_t launch_lean(const Params& p, cudaStream_t stream) {
  switch (p.C) {
    case 1: return launch_lean_nc<XF, T, 1>(p, stream);
    case 3: return launch_lean_nc<XF, T, 3>(p, stream);
    case 4: return launch_lean_nc<XF, T, 4>(p, stream);
  }
  return cudaErrorNotSupported;  // (capi.cu::can_lean)
}

template <int XF, int MODE, typename T, int PX, bool FAST>
static cudaError_t launch_staged_one(const Params& p, cudaStream_t stream) {
  if (p.clip) return launch_staged_clip<XF, MODE, T, PX, FAST, true>(p, stream);
  return launch_staged_clip<XF, MODE, T, PX, FAST, false>(p, stream);
}

// bilinear: PX per dtype, or the 4-pixel fast-coordinate variant; bicubic: exact
// coordinates, 2 pixels per lane (8 cubic weights per pixel are live in the blend)
template <int XF, typename T, int PX>
static cudaError_t launch_staged_t(const Params& p, int mode, bool fast, bool lean,
                                   cudaStream_t stream) {
  if (mode == EQB_BICUBIC) return launch_staged_one<XF, EQB_BICUBIC, T, 2, false>(p, stream);
#if EQB_XF <= 2  // transforms with a transcendental chain have a fast-coordinate variant
         // (enum values are invisible to the preprocessor: 0..2 = equi2equi/pers/cube)
  if (fast) {
    if (lean)
      return launch_lean<XF, typename LeanType<T>::type>(p, stream);
    return launch_staged_one<XF, EQB_BILINEAR, T, 4, true>(p, stream);
  }
#endif
  return launch_staged_one<XF, EQB_BILINEAR, T, PX, false>(p, stream);
}

template <>
cudaError_t launch_staged<EQB_XF>(const Params& p, int mode, int dtype, bool fast, bool lean,
                                  cudaStream_t stream) {
  switch (dtype) {
    case EQB_U8: return launch_staged_t<EQB_XF, uint8_t, 4>(p, mode, fast, lean, stream);
    case EQB_F16: return launch_staged_t<EQB_XF, __half, 2>(p, mode, fast, lean, stream);
    case EQB_BF16: return launch_staged_t<EQB_XF, __nv_bfloat16, 2>(p, mode, fast, lean, stream);
    case EQB_F32: return launch_staged_t<EQB_XF, float, 2>(p, mode, fast, lean, stream);
  }
  return cudaErrorInvalidValue;
}

// Channels-last uint8 RGB / RGBX (EQB_FLAG_CHANNELS_LAST): the lean kernel's IL variants.
template <int XF, int IL>
static cudaError_t launch_nhwc_il(Params p, cudaStream_t stream) {
  const long long tiles_x = (p.dw + 63) / 64;
  const long long tiles_y = (p.dh + kLeanTileH - 1) / kLeanTileH;
  const long long tiles = tiles_x * tiles_y * p.B;
  const long long wave = 148ll * 2048 / (kLeanWarps * 32);
  p.iters = tiles >= wave * 8 * 16 ? 16 : tiles >= wave * 16 ? 4 : 1;
  const dim3 grid((unsigned)((tiles_x + p.iters - 1) / p.iters), (unsigned)tiles_y, (unsigned)p.B);
  // (+16: the RGB tap loads are aligned words that may end past the last tap)
  remap_lean_kernel<XF, uint8_t, IL>
      <<<grid, kLeanWarps * 32, lean_smem_bytes() + 16, stream>>>(p);
  return cudaGetLastError();
}

template <>
cudaError_t launch_nhwc<EQB_XF>(const Params& p, cudaStream_t stream) {
#if EQB_XF <= 1  // equi2equi, equi2pers
  if (p.C == 4) return launch_nhwc_il<EQB_XF, 4>(p, stream);
  if (p.C == 3) return launch_nhwc_il<EQB_XF, 3>(p, stream);
#endif
  return cudaErrorNotSupported;
}

// ---- remap_lean2_kernel (remap_lean2.cuh) ---------------------------------------------------
#if EQB_XF <= 4
template <int XF, typename T, int NC, bool EXACT, int MODE>
static cudaError_t launch_lean2_one(Params p, cudaStream_t stream) {
  const long long tiles_x = (p.dw + kL2TileW - 1) / kL2TileW;
  const long long tiles_y = (p.dh + kL2TileH - 1) / kL2TileH;
  // tiles per strip: as many as leave >= 4 waves of CTAs (the node prologue amortises over
  // the strip); a strip must not straddle two cube cells
  const long long wave = 148ll * (EQB_L2_BUF_KB > 12 ? 6 : 7);
  int iters = kL2MaxIters;
  while (iters > 1 && ((tiles_x + iters - 1) / iters) * tiles_y * p.B < 4 * wave) iters >>= 1;
  if (XF == EQB_EQUI2CUBE)
    while (iters > 1 && p.w_face % (iters * kL2TileW)) iters >>= 1;
  p.iters = iters;
  const dim3 grid((unsigned)((tiles_x + iters - 1) / iters), (unsigned)tiles_y, (unsigned)p.B);
  // (+16: the bicubic tap rows are read as aligned words that may end past the last tap)
  constexpr int SMEM = 2 * l2_buf_bytes((int)sizeof(T), MODE == EQB_BICUBIC) + 16;
  if (SMEM > 48 * 1024)  // (the opt-in is per device: set it on every launch that needs it)
    cudaFuncSetAttribute(remap_lean2_kernel<XF, T, NC, EXACT, MODE>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  remap_lean2_kernel<XF, T, NC, EXACT, MODE>
      <<<grid, kL2Warps * 32, SMEM, stream>>>(p);
  return cudaGetLastError();
}

template <int XF, typename T, int MODE>
static cudaError_t launch_lean2_t(const Params& p, bool exact, cudaStream_t stream) {
  if constexpr (XF == EQB_PERS2EQUI) {  // bicubic, exact chain (no transcendental to save)
    if (MODE != EQB_BICUBIC || !exact) return cudaErrorNotSupported;
    switch (p.C) {
      case 1: return launch_lean2_one<XF, T, 1, true, EQB_BICUBIC>(p, stream);
      case 3: return launch_lean2_one<XF, T, 3, true, EQB_BICUBIC>(p, stream);
      case 4: return launch_lean2_one<XF, T, 4, true, EQB_BICUBIC>(p, stream);
    }
    return cudaErrorNotSupported;
  } else if constexpr (XF == EQB_CUBE2EQUI) {  // the tables ARE the exact coordinates
    if (!exact) return cudaErrorNotSupported;
    switch (p.C) {
      case 1: return launch_lean2_one<XF, T, 1, true, MODE>(p, stream);
      case 3: return launch_lean2_one<XF, T, 3, true, MODE>(p, stream);
      case 4: return launch_lean2_one<XF, T, 4, true, MODE>(p, stream);
    }
    return cudaErrorNotSupported;
  } else {
    switch (p.C * 2 + (exact ? 1 : 0)) {
      case 2: return launch_lean2_one<XF, T, 1, false, MODE>(p, stream);
      case 3: return launch_lean2_one<XF, T, 1, true, MODE>(p, stream);
      case 6: return launch_lean2_one<XF, T, 3, false, MODE>(p, stream);
      case 7: return launch_lean2_one<XF, T, 3, true, MODE>(p, stream);
      case 8: return launch_lean2_one<XF, T, 4, false, MODE>(p, stream);
      case 9: return launch_lean2_one<XF, T, 4, true, MODE>(p, stream);
    }
    return cudaErrorNotSupported;  // (capi.cu::can_lean2)
  }
}

template <int XF, int MODE>
static cudaError_t launch_lean2_m(const Params& p, int dtype, bool exact, cudaStream_t stream) {
  switch (dtype) {
    case EQB_U8: return launch_lean2_t<XF, uint8_t, MODE>(p, exact, stream);
    case EQB_F16: return launch_lean2_t<XF, __half, MODE>(p, exact, stream);
    case EQB_BF16: return launch_lean2_t<XF, __nv_bfloat16, MODE>(p, exact, stream);
    case EQB_F32: return launch_lean2_t<XF, float, MODE>(p, exact, stream);
  }
  return cudaErrorNotSupported;
}
#endif

template <>
cudaError_t launch_lean2<EQB_XF>(const Params& p, int dtype, bool exact, int mode,
                                 cudaStream_t stream) {
#if EQB_XF <= 2
  if (mode == EQB_BICUBIC) return launch_lean2_m<EQB_XF, EQB_BICUBIC>(p, dtype, exact, stream);
  if (mode == EQB_BILINEAR) return launch_lean2_m<EQB_XF, EQB_BILINEAR>(p, dtype, exact, stream);
  if (mode == EQB_NEAREST) return launch_lean2_m<EQB_XF, EQB_NEAREST>(p, dtype, exact, stream);
#elif EQB_XF == 3 || EQB_XF == 4
  if (mode == EQB_BICUBIC) return launch_lean2_m<EQB_XF, EQB_BICUBIC>(p, dtype, exact, stream);
#endif
  return cudaErrorNotSupported;
}

template <>
cudaError_t launch_coords<EQB_XF>(const Params& p, bool libm, cudaStream_t stream) {
  const dim3 grid((p.dw + 31) / 32, (p.dh + 7) / 8, p.B);
  if (libm) coords_kernel<EQB_XF, true><<<grid, 256, 0, stream>>>(p);
  else coords_kernel<EQB_XF, false><<<grid, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace eqb
