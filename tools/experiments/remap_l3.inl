// Launchers of remap_lean3_kernel for transform EQB_XF (one translation unit per transform:
// remap_l3_xf*.cu, so that they compile in parallel with remap_xf*.cu).
#include "remap_lean3.cuh"

namespace eqb {

template <int XF, typename T, int NC, bool EXACT>
static cudaError_t launch_lean3_one(Params p, cudaStream_t stream) {
  const long long tiles_x = (p.dw + kL3TileW - 1) / kL3TileW;
  const long long tiles_y = (p.dh + kL3TileH - 1) / kL3TileH;
  // tiles per strip: as many as leave >= 4 waves of CTAs (the node prologue amortises over
  // the strip); a strip must not straddle two cube cells
  const long long wave = 148ll * 5;
  int iters = kL3MaxIters;
  while (iters > 1 && ((tiles_x + iters - 1) / iters) * tiles_y * p.B < 4 * wave) --iters;
  if (XF == EQB_EQUI2CUBE)
    while (iters > 1 && p.w_face % (iters * kL3TileW)) --iters;
  p.iters = iters;
  const dim3 grid((unsigned)((tiles_x + iters - 1) / iters), (unsigned)tiles_y, (unsigned)p.B);
  remap_lean3_kernel<XF, T, NC, EXACT>
      <<<grid, kL3Warps * 32, kL3Bufs * kL3BufBytes, stream>>>(p);
  return cudaGetLastError();
}

template <int XF, typename T>
static cudaError_t launch_lean3_t(const Params& p, bool exact, cudaStream_t stream) {
  switch (p.C * 2 + (exact ? 1 : 0)) {
    case 2: return launch_lean3_one<XF, T, 1, false>(p, stream);
    case 3: return launch_lean3_one<XF, T, 1, true>(p, stream);
    case 6: return launch_lean3_one<XF, T, 3, false>(p, stream);
    case 7: return launch_lean3_one<XF, T, 3, true>(p, stream);
    case 8: return launch_lean3_one<XF, T, 4, false>(p, stream);
    case 9: return launch_lean3_one<XF, T, 4, true>(p, stream);
  }
  return cudaErrorNotSupported;  // (capi.cu::can_lean3)
}

template <>
cudaError_t launch_lean3<EQB_XF>(const Params& p, int dtype, bool exact, cudaStream_t stream) {
  switch (dtype) {
    case EQB_F16: return launch_lean3_t<EQB_XF, __half>(p, exact, stream);
    case EQB_BF16: return launch_lean3_t<EQB_XF, __nv_bfloat16>(p, exact, stream);
    case EQB_F32: return launch_lean3_t<EQB_XF, float>(p, exact, stream);
  }
  return cudaErrorNotSupported;
}

}  // namespace eqb
