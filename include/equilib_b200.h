/*
 * equilib_b200 -- C ABI of the B200 (sm_100a) remap kernels.
 *
 * Drop-in boundary for equilib's torch remap path.  Every entry point is
 * `extern "C"`, takes plain pointers/sizes (no torch types), never allocates
 * device memory, never synchronises and never throws: it returns an
 * `eqb_status` and records a thread-local message readable with
 * `eqb_last_error()`.  All device pointers must belong to the CUDA device that
 * is current on the calling thread; `stream` is a `cudaStream_t` passed as
 * `void*` (NULL = legacy default stream).
 *
 * Reference interfaces replaced (paths in haruishi43/equilib v0.6.0):
 *   eqb_remap            <- the torch `run()` bodies: grid prep + matmul +
 *                           convert_grid + torch_grid_sample + clip, i.e.
 *                           equilib/equi2equi/torch.py:143-198,
 *                           equilib/equi2pers/torch.py:186-243,
 *                           equilib/equi2cube/torch.py:189-251,
 *                           equilib/cube2equi/torch.py:324-357,
 *                           equilib/pers2equi/torch.py:196-253
 *                           (no H x W x 2 grid is ever materialised);
 *                           with transform = EQB_GRID it replaces
 *                           equilib/grid_sample/torch/native.py:11-90
 *                           (`native()` -> F.grid_sample) for a caller-made grid.
 *   eqb_remap_coords     <- `convert_grid` of each transform, e.g.
 *                           equilib/equi2equi/torch.py:24-42 (test/debug aid).
 *   eqb_equi2cube2equi   <- equi2cube followed by cube2equi (the two `run()` bodies above
 *                           and the cubemap re-formatting between them,
 *                           equilib/equi2cube/torch.py:17-76,
 *                           equilib/cube2equi/torch.py:15-132) in one launch.
 *   eqb_minmax           <- `torch.min(src)`, `torch.max(src)` feeding
 *                           `torch.clip`, e.g. equilib/equi2equi/torch.py:197.
 *   eqb_rotation_compose <- `R_z @ R_y @ R_x` of
 *                           equilib/torch_utils/rotation.py:27-78 (host code).
 */
#ifndef EQUILIB_B200_H_
#define EQUILIB_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EQB_ABI_VERSION 1

typedef enum eqb_status {
  EQB_OK = 0,
  EQB_ERR_INVALID = -1,     /* bad descriptor (null pointer, size, enum value) */
  EQB_ERR_UNSUPPORTED = -2, /* valid request the library does not implement   */
  EQB_ERR_CUDA = -3         /* a CUDA runtime call failed                     */
} eqb_status;

typedef enum eqb_transform {
  EQB_EQUI2EQUI = 0,
  EQB_EQUI2PERS = 1,
  EQB_EQUI2CUBE = 2,
  EQB_CUBE2EQUI = 3,
  EQB_PERS2EQUI = 4,
  EQB_GRID = 5 /* sample with a caller-supplied pixel-unit (y, x) grid */
} eqb_transform;

typedef enum eqb_mode { EQB_NEAREST = 0, EQB_BILINEAR = 1, EQB_BICUBIC = 2 } eqb_mode;

typedef enum eqb_dtype { EQB_U8 = 0, EQB_F16 = 1, EQB_BF16 = 2, EQB_F32 = 3 } eqb_dtype;

/* flags */
#define EQB_FLAG_SATURATE_U8 1u /* uint8 store saturates instead of wrapping mod 256 */
#define EQB_FLAG_LIBM 4u        /* eqb_remap_coords only: CUDA math-library sqrt /    \
                                   div / atan2 instead of their inlined fast paths   */
#define EQB_FLAG_NO_STAGE 8u    /* never use the shared-memory staged kernel           */
#define EQB_FLAG_EXACT_COORDS 16u /* 16/8-bit images: exact coordinate chain on every     \
                                    pixel.  Default for these types (bilinear and bicubic  \
                                    on large launches): the chain on every                 \
                                    4th / 8th pixel of a row + cubic interpolation along  \
                                    x -- measured <= 1.7e-3 px (mean 1.8e-4 px) from the   \
                                    exact chain at 2048 x 4096, test bound 3e-3 px;        \
                                    DESIGN.md 4.3.  Nearest: interpolated coordinates are  \
                                    rounded only away from rounding ties, every other      \
                                    pixel gets the exact chain -- identical taps with or   \
                                    without this flag                     */
#define EQB_FLAG_FAST_COORDS 32u  /* fp32 images: opt in to the interpolated coordinates  */
#define EQB_FLAG_FORCE_STAGE 128u /* experiments: staged kernel wherever it is legal           */
#define EQB_FLAG_NO_TMA 64u       /* staged kernel: cp.async row chunks instead of TMA     */
#define EQB_FLAG_NO_LEAN 1024u    /* experiments: generic staged kernel where the lean     \
                                     variant would run                                     */
#define EQB_FLAG_NO_LEAN2 4096u   /* experiments: the round-1 lean kernel where the strip-node \
                                     kernel (remap_lean2.cuh) would run                    */
#define EQB_FLAG_COMPACT_MAP 32768u /* eqb_map_build / eqb_remap_mapped: 4-byte map entries   \
                                      (see eqb_map_build)                                    */
#define EQB_FLAG_FORCE_LEAN2 8192u /* experiments: strip-node kernel for uint8 too          */
#define EQB_FLAG_CHANNELS_LAST 2048u /* src and dst are channels-last (NHWC): channel stride  \
                                       1, pixel stride = channels; strides_h in elements.     \
                                       Served for uint8 RGB / RGBX bilinear equi2equi and     \
                                       equi2pers on large launches (the layout   \
                                       video decoders hand over; the reference converts HWC    \
                                       to CHW on the host, scripts/equi2pers_torch.py:29-44); \
                                       see eqb_remap_supports_channels_last                   */
/* bits 8..9: force the warp tile shape (0 = auto, 1 = 32x1, 2 = 16x2, 3 = 8x4 lanes) */
#define EQB_FLAG_WARP_SHAPE_SHIFT 8
#define EQB_FLAG_WARP_SHAPE(n) ((uint32_t)(n) << EQB_FLAG_WARP_SHAPE_SHIFT)

/*
 * One remap launch.  Images are planar NCHW views with unit inner stride;
 * strides are in ELEMENTS.  A zero batch stride (expanded batch) is allowed for
 * the source.
 *
 * Per transform:
 *   EQUI2EQUI  src = equirect (src_h x src_w), dst = equirect (dst_h x dst_w).
 *              mats = B x 9 (R).  tab_x = dst_w interleaved (cos(theta), sin(theta))
 *              pairs (8-byte aligned), tab_y = [cos(phi) | sin(phi)] (2*dst_h).
 *   EQUI2PERS  src = equirect, dst = perspective.  mats = B x 9 (R @ G).
 *   EQUI2CUBE  src = equirect, dst = cube canvas.  mats = B x 9 (R).
 *              tab_x = rng (w_face).  dst_h = w_face, dst_w = n_cells * w_face:
 *              the canvas is n_cells (6 or 12) square cells side by side; cell i
 *              lives at element offset cell_offset[i] inside a (b, c) plane of
 *              dst with row stride dst_stride_h; cells 0..5 are faces
 *              F,R,B,L,U,D, cells 6..11 (dice background) are zero-filled.
 *   CUBE2EQUI  src = cube, dst = equirect.  src_h = w_face, src_w = 6 * w_face
 *              (the virtual horizon strip the reference samples from); face i
 *              lives at cell_offset[i] inside a (b, c) plane of src with row
 *              stride src_stride_h.  If face_ptrs != NULL it is a device array
 *              of B x 6 base pointers (one C x w x w face each, addressed with
 *              src_stride_c / src_stride_h) and src / cell_offset are ignored.
 *              tab_x = [tanx | cosx | sin(theta) | cos(theta)] (4*dst_w),
 *              tab_y = [ntan | cu | cd] (3*dst_h), itab_x = [face | thr] (2*dst_w).
 *   PERS2EQUI  src = perspective, dst = equirect.  mats = B x 9 (G @ R);
 *              tables as EQUI2EQUI.
 *   GRID       grid = B x 2 x dst_h x dst_w fp32, (y, x) source pixel units,
 *              batch stride grid_stride_b (0 = shared by all batch items).
 */
typedef struct eqb_remap_desc {
  int32_t transform; /* eqb_transform */
  int32_t mode;      /* eqb_mode      */
  int32_t dtype;     /* eqb_dtype of src and dst */
  uint32_t flags;

  int32_t batch;
  int32_t channels;
  int32_t src_h, src_w;
  int32_t dst_h, dst_w;

  const void* src;
  int64_t src_stride_b, src_stride_c, src_stride_h;
  void* dst;
  int64_t dst_stride_b, dst_stride_c, dst_stride_h;

  const float* mats;     /* device, batch x 9, row-major */
  const float* tab_x;    /* device, per output column    */
  const float* tab_y;    /* device, per output row       */
  const int32_t* itab_x; /* device, per output column    */
  const float* grid;     /* device, EQB_GRID only        */
  int64_t grid_stride_b;
  const float* clip;     /* device float[2] {lo, hi}, or NULL for no clip */

  int32_t w_face;
  int32_t n_cells;
  int64_t cell_offset[12];
  const void* const* face_ptrs;
} eqb_remap_desc;

/* Plain planar image view, for eqb_minmax. */
typedef struct eqb_image {
  const void* data;
  int32_t dtype;
  int32_t batch, channels, height, width;
  int64_t stride_b, stride_c, stride_h;
} eqb_image;

/* ABI version of the loaded library (== EQB_ABI_VERSION it was built with). */
int32_t eqb_abi_version(void);

/* Message of the last failing call on this thread ("" if none). */
const char* eqb_last_error(void);

/* Fused ray generation + ray-to-pixel map + sampling + output policy. */
int32_t eqb_remap(const eqb_remap_desc* desc, void* stream);

/*
 * 1 when eqb_remap serves `desc` with EQB_FLAG_CHANNELS_LAST set (the flag itself need not
 * be set in `desc`), else 0: the caller then converts to the planar layout, as
 * equilib/equi2pers/base.py does for every input.  Host only, no launch.
 */
int32_t eqb_remap_supports_channels_last(const eqb_remap_desc* desc);

/*
 * Static-camera path (the reference's roadmap item "full-grid cache for the
 * fixed-rotation case", README.md:168; its equilib/_cache.py:21-40 caches the rays only).
 *
 * eqb_map_build packs coordinates -- the output of eqb_remap_coords for the same
 * descriptor -- into a sampling map: per output pixel 8 bytes (top-left tap x | y << 16,
 * bilinear fractions as round(f * 65535) x | y << 16) in `map` (batch x dst_h x dst_w
 * entries) and per EQB_MAP_TILE_W x EQB_MAP_TILE_H output tile 16 bytes in `heads`
 * (batch x ceil(dst_h / TILE_H) x ceil(dst_w / TILE_W) entries: origin and shape index of
 * the shared-memory box the tile's footprint fits, or -1).  desc->dtype and
 * desc->channels select the box table; src / dst are not touched.
 * With EQB_FLAG_COMPACT_MAP the entries are 4 bytes (the reference roadmap's "compact"
 * full-grid cache, README.md:168): the tap as an offset from the tile's box origin (x | y << 8)
 * and the fractions as round(f * 255) << 16 / << 24 -- coordinates quantised to 1/510 pixel;
 * tiles without a box have no entries and eqb_remap_mapped evaluates the exact chain for them
 * (mats / tables of the descriptor must then be those the map was built from).
 *
 * eqb_remap_mapped remaps desc->src into desc->dst from such a map: bilinear, planar
 * 1 / 3 / 4 channels of uint8 / fp16 / bf16 -- or, with EQB_FLAG_CHANNELS_LAST, uint8 RGB /
 * RGBX frames in NHWC layout -- no clamp, 16-byte aligned source rows, 4-pixel aligned
 * destination rows (EQB_ERR_UNSUPPORTED otherwise).  Results equal eqb_remap with
 * EQB_FLAG_EXACT_COORDS up to the 16-bit fractions.  coords_out of eqb_remap_coords is also
 * "used by the product path" here, once per set of rotations.
 */
#define EQB_MAP_TILE_W 64
#define EQB_MAP_TILE_H 8
int32_t eqb_map_build(const eqb_remap_desc* desc, const float* coords, void* map, void* heads,
                      void* stream);
int32_t eqb_remap_mapped(const eqb_remap_desc* desc, const void* map, const void* heads,
                         void* stream);

/*
 * Fused equirect -> cubemap -> equirect: the result of eqb_remap(to_cube)
 * followed by eqb_remap(to_equi) without the cubemap ever touching memory -- the reference
 * writes it, re-formats it and reads it back (equilib/equi2cube/torch.py:189-251,58-76,
 * equilib/cube2equi/torch.py:22-65,324-357).  `to_cube` is the EQUI2CUBE descriptor of the first
 * launch (src = the equirect, mats, tab_x = rng, w_face, n_cells = 6; its dst is ignored),
 * `to_equi` the CUBE2EQUI descriptor of the second (tables, dst = the output; its src is
 * ignored).  Both must agree in mode (nearest or bilinear), dtype, batch, channels (<= 4) and
 * w_face and carry no clamp.  Cube taps are resolved on the virtual horizon strip, so the
 * reference's cross-face tap bleed (cube2equi/torch.py:228-242) is reproduced; the intermediate
 * cube pixel is rounded to the image type like a stored cubemap; every cube pixel is produced
 * by the exact coordinate chain.
 */
int32_t eqb_equi2cube2equi(const eqb_remap_desc* to_cube, const eqb_remap_desc* to_equi,
                           void* stream);

/*
 * Writes the sampling coordinates of the transform (what the reference calls
 * the grid: B x 2 x dst_h x dst_w fp32, (uj, ui) in source pixel units, before
 * the sampler's normalisation) to `coords_out`.  src/dst/dtype are ignored.
 * Input of eqb_map_build; otherwise for tests and debugging.
 */
int32_t eqb_remap_coords(const eqb_remap_desc* desc, float* coords_out, void* stream);

/* Global min and max of an image batch -> out2[0], out2[1] (device floats). */
int32_t eqb_minmax(const eqb_image* img, float* out2, void* stream);

/*
 * HOST function.  trig = n x 6 floats (cos r, sin r, cos p, sin p, cos y, sin y;
 * the caller applies the z_down sign flip and the float64 -> float32 rounding);
 * out = n x 9 floats, R = (R_z @ R_y) @ R_x with fp32 fused multiply-add chains.
 */
int32_t eqb_rotation_compose(const float* trig, int32_t n, float* out);

/* Number of kernel launches issued by this library since load (all threads). */
int64_t eqb_launch_count(void);

/*
 * Identity of the build: the first 16 hex digits of sha256 over the sources, headers and
 * compiler flags the library was built from (equilib_b200/build.py::source_hash), "" for a
 * build made outside build.py.  build() rebuilds when it differs from the tree's.
 */
const char* eqb_source_hash(void);

#ifdef __cplusplus
}
#endif
#endif /* EQUILIB_B200_H_ */
