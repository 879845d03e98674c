// Included by remap_xf<N>.cu with EQB_XF defined: instantiates the kernels of one
// transform and its mode/dtype/PX dispatch (one translation unit per transform so
// the six compile in parallel).
#include "remap_kernels.cuh"

namespace eqb {

// Pixels per thread (= store width) instantiated per dtype and mode; must agree
// with vec_px() in capi.cu.  uint8 bicubic stays at 2 (4 would need 250 registers).
template <typename T, int MODE> struct VecPX { static constexpr int value = 2; };
template <> struct VecPX<uint8_t, EQB_NEAREST> { static constexpr int value = 4; };
template <> struct VecPX<uint8_t, EQB_BILINEAR> { static constexpr int value = 4; };

template <int XF, int MODE, typename T, int PX>
static cudaError_t launch_one(const Params& p, cudaStream_t stream) {
  const dim3 block(kTileW, kTileH);
  const dim3 grid((p.dw + kTileW * PX - 1) / (kTileW * PX), (p.dh + kTileH - 1) / kTileH, p.B);
  remap_kernel<XF, MODE, T, PX><<<grid, block, 0, stream>>>(p);
  return cudaGetLastError();
}

template <int XF, int MODE, typename T>
static cudaError_t launch_px(const Params& p, int px, cudaStream_t stream) {
  if (px == 1) return launch_one<XF, MODE, T, 1>(p, stream);
  return launch_one<XF, MODE, T, VecPX<T, MODE>::value>(p, stream);
}

template <int XF, int MODE>
static cudaError_t launch_dtype(const Params& p, int dtype, int px, cudaStream_t stream) {
  switch (dtype) {
    case EQB_U8: return launch_px<XF, MODE, uint8_t>(p, px, stream);
    case EQB_F16: return launch_px<XF, MODE, __half>(p, px, stream);
    case EQB_BF16: return launch_px<XF, MODE, __nv_bfloat16>(p, px, stream);
    case EQB_F32: return launch_px<XF, MODE, float>(p, px, stream);
  }
  return cudaErrorInvalidValue;
}

template <>
cudaError_t launch_remap<EQB_XF>(const Params& p, int mode, int dtype, int px,
                                 cudaStream_t stream) {
  switch (mode) {
    case EQB_NEAREST: return launch_dtype<EQB_XF, EQB_NEAREST>(p, dtype, px, stream);
    case EQB_BILINEAR: return launch_dtype<EQB_XF, EQB_BILINEAR>(p, dtype, px, stream);
    case EQB_BICUBIC: return launch_dtype<EQB_XF, EQB_BICUBIC>(p, dtype, px, stream);
  }
  return cudaErrorInvalidValue;
}

template <>
cudaError_t launch_coords<EQB_XF>(const Params& p, cudaStream_t stream) {
  const dim3 block(kTileW, kTileH);
  const dim3 grid((p.dw + kTileW - 1) / kTileW, (p.dh + kTileH - 1) / kTileH, p.B);
  coords_kernel<EQB_XF><<<grid, block, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace eqb
