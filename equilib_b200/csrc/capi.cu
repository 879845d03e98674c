// extern "C" boundary of libequilib_b200.so (see include/equilib_b200.h).
// Validates descriptors, picks the store width, launches.  No allocation, no sync.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "remap_kernels.cuh"
#include "remap_lean2.cuh"

namespace eqb {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static size_t elt_size(int dtype) {
  switch (dtype) {
    case EQB_U8: return 1;
    case EQB_F16:
    case EQB_BF16: return 2;
    case EQB_F32: return 4;
  }
  return 0;
}

static int vec_px(int dtype, int mode, int transform) {
  if (mode == EQB_BICUBIC && transform != EQB_PERS2EQUI) return 1;  // see launch_px
  if (transform == EQB_CUBE2EQUI && mode == EQB_BILINEAR) return 2;  // see launch_px
  return dtype == EQB_U8 ? 4 : 2;
}

// Common validation + Params fill.  `need_images` is false for eqb_remap_coords.
static int fill(const eqb_remap_desc* d, bool need_images, Params& p) {
  if (!d) return fail(EQB_ERR_INVALID, "null descriptor");
  if (d->transform < EQB_EQUI2EQUI || d->transform > EQB_GRID)
    return fail(EQB_ERR_INVALID, "unknown transform %d", d->transform);
  if (d->batch <= 0 || d->batch > 65535)
    return fail(EQB_ERR_INVALID, "batch %d out of range [1, 65535]", d->batch);
  if (d->dst_h <= 0 || d->dst_w <= 0 || d->src_h < 2 || d->src_w < 2)
    return fail(EQB_ERR_INVALID, "bad sizes: src %dx%d (each must be >= 2), dst %dx%d", d->src_h,
                d->src_w, d->dst_h, d->dst_w);
  if (d->dst_h > 65535 * kWarps)
    return fail(EQB_ERR_INVALID, "dst_h %d too large", d->dst_h);
  std::memset(&p, 0, sizeof(p));
  p.B = d->batch;
  p.C = d->channels;
  p.sh = d->src_h;
  p.sw = d->src_w;
  p.dh = d->dst_h;
  p.dw = d->dst_w;
  p.mats = d->mats;
  p.tx = d->tab_x;
  p.ty = d->tab_y;
  p.itx = d->itab_x;
  p.grid = d->grid;
  p.gsb = d->grid_stride_b;
  p.clip = d->clip;
  p.inv_wm1 = 1.0f / (float)(d->src_w - 1);
  p.inv_hm1 = 1.0f / (float)(d->src_h - 1);
  p.shf = (float)d->src_h;
  p.swf = (float)d->src_w;
  p.shm1f = (float)(d->src_h - 1);
  p.swm1f = (float)(d->src_w - 1);
  p.inv_wface = d->w_face > 0 ? 1.0f / (float)d->w_face : 0.0f;
  p.one = 1.0f;
  p.face_maps = 0;
  p.xf = d->transform;
  p.iters = 1;
  p.w_face = d->w_face;
  p.n_cells = d->n_cells;
  p.flags = d->flags;
  p.face_ptrs = d->face_ptrs;
  for (int i = 0; i < 12; ++i) p.cell_off[i] = d->cell_offset[i];

  const int xf = d->transform;
  if (xf != EQB_CUBE2EQUI && xf != EQB_GRID && !d->mats)
    return fail(EQB_ERR_INVALID, "mats is required for transform %d", xf);
  if ((xf == EQB_EQUI2EQUI || xf == EQB_PERS2EQUI) && (!d->tab_x || !d->tab_y))
    return fail(EQB_ERR_INVALID, "tab_x/tab_y are required for transform %d", xf);
  if (xf == EQB_EQUI2CUBE) {
    if (!d->tab_x) return fail(EQB_ERR_INVALID, "tab_x (rng) is required for equi2cube");
    if (d->w_face <= 0 || (d->n_cells != 6 && d->n_cells != 12))
      return fail(EQB_ERR_INVALID, "equi2cube needs w_face > 0 and n_cells in {6, 12}");
    if (d->dst_h != d->w_face || d->dst_w != d->n_cells * d->w_face)
      return fail(EQB_ERR_INVALID, "equi2cube canvas must be w_face x n_cells*w_face");
  }
  if (xf == EQB_CUBE2EQUI) {
    if (!d->tab_x || !d->tab_y || !d->itab_x)
      return fail(EQB_ERR_INVALID, "tab_x/tab_y/itab_x are required for cube2equi");
    if (d->w_face < 2 || d->src_h != d->w_face || d->src_w != 6 * d->w_face)
      return fail(EQB_ERR_INVALID, "cube2equi source must be the w_face x 6*w_face strip");
  }
  if (xf == EQB_GRID && !d->grid) return fail(EQB_ERR_INVALID, "grid is required for EQB_GRID");

  if (need_images) {
    if (d->mode < EQB_NEAREST || d->mode > EQB_BICUBIC)
      return fail(EQB_ERR_INVALID, "unknown mode %d", d->mode);
    if (!elt_size(d->dtype)) return fail(EQB_ERR_INVALID, "unknown dtype %d", d->dtype);
    if (d->channels <= 0) return fail(EQB_ERR_INVALID, "channels must be positive");
    if (!d->dst) return fail(EQB_ERR_INVALID, "dst is null");
    if (!d->src && !(xf == EQB_CUBE2EQUI && d->face_ptrs))
      return fail(EQB_ERR_INVALID, "src is null");
    if (d->src_stride_h < (xf == EQB_CUBE2EQUI ? d->w_face : d->src_w))
      return fail(EQB_ERR_INVALID, "src_stride_h smaller than a row");
    {
      // in-plane source offsets are 32-bit in the kernels
      long long span = (long long)d->src_stride_h * d->src_h;
      if (xf == EQB_CUBE2EQUI && !d->face_ptrs)
        for (int i = 0; i < 6; ++i)
          if (d->cell_offset[i] + span > span) span = d->cell_offset[i] + span;
      if (span >= (1ll << 31))
        return fail(EQB_ERR_UNSUPPORTED, "a source plane spans %lld elements (limit 2^31)", span);
    }
    p.src = d->src;
    p.dst = d->dst;
    p.ssb = d->src_stride_b;
    p.ssc = d->src_stride_c;
    p.ssh = d->src_stride_h;
    p.dsb = d->dst_stride_b;
    p.dsc = d->dst_stride_c;
    p.dsh = d->dst_stride_h;
  }
  return EQB_OK;
}

// Widest store the destination layout allows: every row start of every plane (and
// cell) must be aligned to PX elements.
static int pick_px(const eqb_remap_desc* d) {
  const int v = vec_px(d->dtype, d->mode, d->transform);
  // (the vectorised bicubic pers2equi holds one result per channel and pixel in registers)
  if (d->mode == EQB_BICUBIC && d->transform == EQB_PERS2EQUI && d->channels > 4) return 1;
  const size_t bytes = v * elt_size(d->dtype);
  if (reinterpret_cast<uintptr_t>(d->dst) % bytes) return 1;
  if (d->dst_stride_b % v || d->dst_stride_c % v || d->dst_stride_h % v) return 1;
  if (d->transform == EQB_EQUI2CUBE) {
    if (d->w_face % v) return 1;
    for (int i = 0; i < d->n_cells; ++i)
      if (d->cell_offset[i] % v) return 1;
  }
  return v;
}

// Warp tile shape: 0 = 32x1, 1 = 16x2, 2 = 8x4 lanes.  Forced through the flags for
// experiments, otherwise picked per transform (DESIGN.md "warp tile").
static int pick_ws(const eqb_remap_desc* d) {
  const unsigned forced = (d->flags >> EQB_FLAG_WARP_SHAPE_SHIFT) & 3u;
  if (forced) return (int)forced - 1;
  return 1;
}

// The staged (shared-memory) kernel serves bilinear sampling from a plain image when
// its 16-byte cp.async row chunks are aligned and the launch is large enough to pay
// for the CTA-wide synchronisation; everything else takes the direct-gather kernel.
static bool dst_vector_ok(const eqb_remap_desc* d, int v) {
  const size_t bytes = v * elt_size(d->dtype);
  if (reinterpret_cast<uintptr_t>(d->dst) % bytes) return false;
  if (d->dst_stride_b % v || d->dst_stride_c % v || d->dst_stride_h % v) return false;
  if (d->transform == EQB_EQUI2CUBE) {
    if (d->w_face % v) return false;
    for (int i = 0; i < d->n_cells; ++i)
      if (d->cell_offset[i] % v) return false;
  }
  return true;
}

static bool tma_available();  // (cuTensorMapEncodeTiled resolved, below)

static bool cube_is_horizon(const eqb_remap_desc* d) {
  if (d->face_ptrs || d->src_stride_h < 6ll * d->w_face) return false;
  for (int i = 0; i < 6; ++i)
    if (d->cell_offset[i] != (long long)i * d->w_face) return false;
  return true;
}

static bool can_stage(const eqb_remap_desc* d) {
  if (d->flags & EQB_FLAG_NO_STAGE) return false;
  if (d->transform == EQB_CUBE2EQUI && !cube_is_horizon(d)) {
    // a true horizon strip is one plain image (its cross-face tap bleed is simply what the
    // memory holds); any other face canvas (dice, stacked faces) is staged face by face
    // through per-face tensor maps; separate face tensors (face_ptrs) gather directly
    const long long per16 = 16 / (long long)elt_size(d->dtype);
    if (d->face_ptrs || (d->flags & EQB_FLAG_NO_TMA) || !tma_available()) return false;
    if (d->w_face % per16 || d->w_face < 16) return false;
    for (int i = 0; i < 6; ++i)
      if (d->cell_offset[i] % per16) return false;
  }
  // pers2equi: most of the output lies outside the field of view and is a plain fill in
  // the direct kernel (no taps, no tile synchronisation): 1.7 ms vs 3.4 ms for 8 x 8K
  if (d->transform == EQB_PERS2EQUI) return false;
  if (d->mode != EQB_BILINEAR && d->mode != EQB_BICUBIC) return false;
  // store width of the staged kernel: 2 pixels (bicubic, 16/32-bit bilinear), 4 (uint8)
  const int v = d->mode == EQB_BICUBIC ? 2 : (d->dtype == EQB_U8 ? 4 : 2);
  if (!dst_vector_ok(d, v)) return false;
  const long long per16 = 16 / (long long)elt_size(d->dtype);
  if (reinterpret_cast<uintptr_t>(d->src) % 16) return false;
  if (d->src_stride_b % per16 || d->src_stride_c % per16 || d->src_stride_h % per16) return false;
  if (d->src_w > 65536 || d->src_h > 65536 || d->src_w < 4 || d->src_h < 4) return false;
  if (d->channels > 4) return false;
  return (long long)d->batch * d->dst_h * d->dst_w >= (1ll << 18);
}

// ---- TMA descriptors for the staged kernel ------------------------------------------
// cuTensorMapEncodeTiled lives in the driver library; it is fetched through the runtime
// so that this library links against cudart only.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) !=
            cudaSuccess || q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

static bool tma_available() { return encode_tiled() != nullptr; }

// ---- tensor-map cache ---------------------------------------------------------------------
// Encoding a tensor map costs ~1 us and a staged launch needs up to 13 of them; repeated
// launches on the same source (video frames in a reused buffer, BASELINE config 1's 6 us kernel)
// would spend more host time encoding than the kernel runs.  A tensor map is a pure function of
// (data type, base pointer, dims, strides, box), so encoded maps are kept in a small direct-
// mapped table keyed by exactly those values.  (All maps here: rank 4, element strides 1, no
// interleave / swizzle, 128-byte L2 promotion, zero fill.)
struct TmapKey {
  unsigned long long ptr, dims[4], strides[3];
  unsigned box[4];
  int dt;
  bool operator==(const TmapKey& o) const { return std::memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapSlot {
  TmapKey key;
  CUtensorMap map;
  bool used;
};
static std::mutex g_tmap_mutex;
static TmapSlot g_tmap_cache[512];

static CUresult encode_cached(CUtensorMap* out, CUtensorMapDataType dt, void* ptr,
                              const cuuint64_t* dims, const cuuint64_t* strides,
                              const cuuint32_t* box) {
  TmapKey k;
  std::memset(&k, 0, sizeof(k));
  k.ptr = reinterpret_cast<unsigned long long>(ptr);
  for (int i = 0; i < 4; ++i) { k.dims[i] = dims[i]; k.box[i] = box[i]; }
  for (int i = 0; i < 3; ++i) k.strides[i] = strides[i];
  k.dt = (int)dt;
  unsigned long long h = 1469598103934665603ull;
  const unsigned char* b = reinterpret_cast<const unsigned char*>(&k);
  for (size_t i = 0; i < sizeof(k); ++i) h = (h ^ b[i]) * 1099511628211ull;
  TmapSlot& slot = g_tmap_cache[h % 512];
  {
    std::lock_guard<std::mutex> lock(g_tmap_mutex);
    if (slot.used && slot.key == k) {
      *out = slot.map;
      return CUDA_SUCCESS;
    }
  }
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return CUDA_ERROR_NOT_SUPPORTED;
  const cuuint32_t ones[4] = {1, 1, 1, 1};
  const CUresult r = enc(out, dt, 4, ptr, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_SUCCESS) {
    std::lock_guard<std::mutex> lock(g_tmap_mutex);
    slot.key = k;
    slot.map = *out;
    slot.used = true;
  }
  return r;
}

// Tensor maps over the source viewed as (W, H, C, B): a wide box for near-upright views,
// a square one for views rotated by 45 degrees and one in between, all within the
// kernel's shared-memory budget.
static void setup_tma(const eqb_remap_desc* d, Params& p, bool lean) {
  p.use_tma = 0;
  if (d->flags & EQB_FLAG_NO_TMA) return;
  if (!encode_tiled()) return;
  const int elt = (int)elt_size(d->dtype);
  // box shapes: the compile-time table the kernels are specialised on
  int bw[kTmaBoxes], bh[kTmaBoxes];
  for (int m = 0; m < kTmaBoxes; ++m) {
    bw[m] = box_w(elt, m);
    bh[m] = lean ? lean_box_h(elt, m, d->channels) : box_h(elt, d->channels, m);
  }
  if (bh[3] < 20) return;  // (more than 4 channels of 4-byte pixels)
  CUtensorMapDataType dt = d->dtype == EQB_U8    ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                           : d->dtype == EQB_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                           : d->dtype == EQB_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                  : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const bool batched = d->src_stride_b > 0 && d->batch > 1;
  const cuuint64_t dims[4] = {(cuuint64_t)d->src_w, (cuuint64_t)d->src_h, (cuuint64_t)d->channels,
                              (cuuint64_t)(batched ? d->batch : 1)};
  const cuuint64_t cstride = (cuuint64_t)d->src_stride_c * elt;
  const cuuint64_t strides[3] = {(cuuint64_t)d->src_stride_h * elt,
                                 d->channels > 1 ? cstride : (cuuint64_t)d->src_stride_h * elt * d->src_h,
                                 batched ? (cuuint64_t)d->src_stride_b * elt
                                         : (d->channels > 1 ? cstride : (cuuint64_t)d->src_stride_h * elt * d->src_h) * d->channels};
  for (int i = 0; i < 3; ++i)
    if (strides[i] == 0 || strides[i] % 16) return;
  for (int m = 0; m < kTmaBoxes; ++m) {
    const cuuint32_t box[4] = {(cuuint32_t)bw[m], (cuuint32_t)bh[m], (cuuint32_t)d->channels, 1};
    if (encode_cached(&p.tmaps[m], dt, const_cast<void*>(d->src), dims, strides, box) !=
        CUDA_SUCCESS)
      return;
    p.tma_bw[m] = bw[m];
    p.tma_bh[m] = bh[m];
  }
  for (int m = kTmaBoxes; m < 16; ++m) p.tma_bw[m] = p.tma_bh[m] = 0;
  p.tma_batched = batched ? 1 : 0;
  p.use_tma = 1;
}

// cube2equi from a face canvas that is not a horizon strip: kFaceBoxes tensor maps per face
// over the face viewed as (w, w, C, B) at src + cell_offset[f]; tmaps[f * kFaceBoxes + m].
static void setup_tma_faces(const eqb_remap_desc* d, Params& p) {
  p.use_tma = 0;
  p.face_maps = 0;
  if (!encode_tiled()) return;
  const int elt = (int)elt_size(d->dtype);
  const int per16 = 16 / elt;
  CUtensorMapDataType dt = d->dtype == EQB_U8    ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                           : d->dtype == EQB_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                           : d->dtype == EQB_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                  : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const bool batched = d->src_stride_b > 0 && d->batch > 1;
  const cuuint64_t w = (cuuint64_t)d->w_face;
  const cuuint64_t dims[4] = {w, w, (cuuint64_t)d->channels, (cuuint64_t)(batched ? d->batch : 1)};
  const cuuint64_t row = (cuuint64_t)d->src_stride_h * elt;
  const cuuint64_t cstride = d->channels > 1 ? (cuuint64_t)d->src_stride_c * elt : row * w;
  const cuuint64_t strides[3] = {row, cstride,
                                 batched ? (cuuint64_t)d->src_stride_b * elt
                                         : cstride * d->channels};
  for (int i = 0; i < 3; ++i)
    if (strides[i] == 0 || strides[i] % 16) return;
  const int budget = staged_smem_bytes(elt) / elt / d->channels;
  const int widths[kFaceBoxes] = {128, 96, 64};
  for (int m = 0; m < 16; ++m) p.tma_bw[m] = p.tma_bh[m] = 0;
  for (int m = 0; m < kFaceBoxes; ++m) {
    const int bw = (widths[m] + per16 - 1) / per16 * per16;
    const int bh = budget / bw < 128 ? budget / bw : 128;
    if (bh < 8) return;
    const cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)d->channels, 1};
    for (int f = 0; f < 6; ++f) {
      void* base = const_cast<char*>(static_cast<const char*>(d->src)) +
                   (long long)d->cell_offset[f] * elt;
      if (reinterpret_cast<uintptr_t>(base) % 16) return;
      if (encode_cached(&p.tmaps[f * kFaceBoxes + m], dt, base, dims, strides, box) !=
          CUDA_SUCCESS)
        return;
    }
    p.tma_bw[m] = bw;
    p.tma_bh[m] = bh;
  }
  p.tma_batched = batched ? 1 : 0;
  p.use_tma = 1;
  p.face_maps = 1;
}

// Fast coordinates (exact chain on every 4th pixel + cubic interpolation) are the
// default for 16- and 8-bit images, opt-in for fp32, and need 4-pixel vector stores.
static bool can_fast(const eqb_remap_desc* d) {
  if (d->transform > EQB_EQUI2CUBE || d->mode != EQB_BILINEAR) return false;
  if (d->flags & EQB_FLAG_EXACT_COORDS) return false;
  if (d->dtype == EQB_F32 && !(d->flags & EQB_FLAG_FAST_COORDS)) return false;
  if (!dst_vector_ok(d, 4)) return false;
  // a 16-lane row segment (64 pixels) must not straddle two faces
  if (d->transform == EQB_EQUI2CUBE && d->w_face % 64) return false;
  return true;
}

// The lean variant of the fast-coordinate kernel: 1, 3 or 4 channels of 1- or 2-byte pixels,
// no clamp (remap_lean_kernel; its boxes differ, so the choice is made before the
// tensor maps are encoded).
static bool can_lean(const eqb_remap_desc* d) {
  return can_fast(d) && (d->channels == 1 || d->channels == 3 || d->channels == 4) &&
         d->dtype != EQB_F32 && !d->clip &&
         !(d->flags & (EQB_FLAG_NO_LEAN | EQB_FLAG_NO_TMA));
}

// Channels-last uint8 (EQB_FLAG_CHANNELS_LAST): served by the channels-last variant of the
// lean kernel only -- RGB / RGBX, bilinear, fast coordinates, equi2equi / equi2pers, TMA.
static bool can_nhwc(const eqb_remap_desc* d) {
  if (d->dtype != EQB_U8 || (d->channels != 3 && d->channels != 4)) return false;
  if (d->mode != EQB_BILINEAR || d->transform > EQB_EQUI2PERS) return false;
  if (d->flags & (EQB_FLAG_EXACT_COORDS | EQB_FLAG_NO_STAGE | EQB_FLAG_NO_TMA)) return false;
  if (d->clip || !encode_tiled()) return false;
  const long long c = d->channels;
  if (d->src_stride_c != 1 || d->dst_stride_c != 1) return false;
  if (d->src_stride_h < c * d->src_w || d->dst_stride_h < c * d->dst_w) return false;
  // source rows as 32-bit words under 16-byte aligned strides (TMA)
  if (reinterpret_cast<uintptr_t>(d->src) % 16 || d->src_stride_h % 16 || d->src_stride_b % 16)
    return false;
  if ((c * d->src_w) % 4) return false;
  // 4-pixel vector stores: 16 (RGBX) or 12 (RGB) bytes per lane
  const long long va = c == 4 ? 16 : 4;
  if (reinterpret_cast<uintptr_t>(d->dst) % va || d->dst_stride_h % va || d->dst_stride_b % va)
    return false;
  if (d->src_w > 65536 || d->src_h > 65536 || d->src_w < 4 || d->src_h < 4) return false;
  if ((long long)d->src_stride_h * d->src_h >= (1ll << 31)) return false;
  return (long long)d->batch * d->dst_h * d->dst_w >= (1ll << 18);
}

// Tensor maps for the channels-last kernel: a source row is W*C/4 32-bit words.
static void setup_tma_nhwc(const eqb_remap_desc* d, Params& p) {
  p.use_tma = 0;
  if (!encode_tiled()) return;
  const int c = d->channels;
  const bool batched = d->src_stride_b > 0 && d->batch > 1;
  const cuuint64_t row = (cuuint64_t)d->src_stride_h;  // bytes (uint8)
  const cuuint64_t dims[4] = {(cuuint64_t)d->src_w * c / 4, (cuuint64_t)d->src_h, 1,
                              (cuuint64_t)(batched ? d->batch : 1)};
  const cuuint64_t strides[3] = {row, row * d->src_h,
                                 batched ? (cuuint64_t)d->src_stride_b : row * d->src_h};
  for (int m = 0; m < 16; ++m) p.tma_bw[m] = p.tma_bh[m] = 0;
  for (int m = 0; m < kTmaBoxes; ++m) {
    const int bw = box_w(1, m), bh = lean_box_h(1, m, c);
    if (bw * c / 4 > 256) continue;  // (box dimensions are limited to 256 elements)
    const cuuint32_t box[4] = {(cuuint32_t)(bw * c / 4), (cuuint32_t)bh, 1, 1};
    if (encode_cached(&p.tmaps[m], CU_TENSOR_MAP_DATA_TYPE_UINT32, const_cast<void*>(d->src),
                      dims, strides, box) != CUDA_SUCCESS)
      return;
    p.tma_bw[m] = bw;
    p.tma_bh[m] = bh;
  }
  p.tma_batched = batched ? 1 : 0;
  p.use_tma = 1;
}

// The strip-node kernel (remap_lean2.cuh): planar 1 / 3 / 4 channels, equi2equi / equi2pers /
// equi2cube, TMA.  Bilinear without a clamp: fp32 images run it with the exact chain on every
// pixel unless EQB_FLAG_FAST_COORDS is set; 8/16-bit images interpolate between the strip's
// nodes unless EQB_FLAG_EXACT_COORDS is set.  Bicubic (with or without a clamp) and nearest: the
// same coordinate policy, every dtype including uint8 (nearest rounds interpolated coordinates
// only away from rounding ties: its taps are those of the exact chain either way).
static bool can_lean2(const eqb_remap_desc* d) {
  if (d->flags & (EQB_FLAG_NO_STAGE | EQB_FLAG_NO_TMA | EQB_FLAG_NO_LEAN | EQB_FLAG_NO_LEAN2 |
                  EQB_FLAG_CHANNELS_LAST))
    return false;
  const bool cubic = d->mode == EQB_BICUBIC;
  const bool nearest = d->mode == EQB_NEAREST && d->transform <= EQB_EQUI2CUBE;
  // pers2equi: bicubic only (fill tiles cost no divide, no tap; the in-view ones read boxes);
  // cube2equi: bicubic from a true horizon strip, which is one plain image (measured on 16 x 4K
  // from w_face 1024: bicubic fp32 2.74 -> 2.42 ms, bf16 2.56 -> 1.90, uint8 2.65 -> 2.06;
  // bilinear 1.25 -> 1.21 / 1.03 -> 1.04: stays on the generic staged kernel, like every other
  // canvas, which that kernel stages face by face)
  if (d->transform == EQB_CUBE2EQUI) {
    if (!cubic || !cube_is_horizon(d)) return false;
  } else if (d->transform > EQB_EQUI2CUBE && !(d->transform == EQB_PERS2EQUI && cubic)) {
    return false;
  }
  // (nearest: a clamp cannot change a copied element; the host launches none)
  if (!cubic && !nearest && (d->mode != EQB_BILINEAR || d->clip)) return false;
  if (nearest && d->clip) return false;
  if (d->channels != 1 && d->channels != 3 && d->channels != 4) return false;
  if (!encode_tiled()) return false;
  const long long elt = (long long)elt_size(d->dtype), per16 = 16 / elt;
  if (reinterpret_cast<uintptr_t>(d->src) % 16) return false;
  if (d->src_stride_b % per16 || d->src_stride_c % per16 || d->src_stride_h % per16) return false;
  if (d->src_w > 65536 || d->src_h > 65536 || d->src_w < 16 || d->src_h < 16) return false;
  if (elt == 1) {
    // measured (gpurun_out/check_lean2.log): 2-byte stores of pixel pairs make the strip-node
    // kernel slower than the round-1 lean kernel on uint8 BILINEAR (1.71 vs 1.40 ms on the bench
    // rotations); it stays there unless asked for (EQB_FLAG_FORCE_LEAN2, experiments)
    if (!cubic && !nearest && !(d->flags & EQB_FLAG_FORCE_LEAN2)) return false;
    if (reinterpret_cast<uintptr_t>(d->dst) % 2) return false;
    if (d->dst_stride_b % 2 || d->dst_stride_c % 2 || d->dst_stride_h % 2) return false;
  }
  if (d->transform == EQB_EQUI2CUBE) {
    if (d->w_face % 64) return false;  // a 64-pixel tile must not straddle two cells
    for (int i = 0; i < d->n_cells; ++i)
      if (d->cell_offset[i] % 2) return false;
  }
  return (long long)d->batch * d->dst_h * d->dst_w >= (1ll << 18);
}

static void setup_tma_l2(const eqb_remap_desc* d, Params& p) {
  p.use_tma = 0;
  if (!encode_tiled()) return;
  const int elt = (int)elt_size(d->dtype);
  CUtensorMapDataType dt = d->dtype == EQB_U8    ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                           : d->dtype == EQB_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                           : d->dtype == EQB_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                  : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const bool batched = d->src_stride_b > 0 && d->batch > 1;
  const cuuint64_t dims[4] = {(cuuint64_t)d->src_w, (cuuint64_t)d->src_h, (cuuint64_t)d->channels,
                              (cuuint64_t)(batched ? d->batch : 1)};
  const cuuint64_t row = (cuuint64_t)d->src_stride_h * elt;
  const cuuint64_t cstride = d->channels > 1 ? (cuuint64_t)d->src_stride_c * elt : row * d->src_h;
  const cuuint64_t strides[3] = {row, cstride,
                                 batched ? (cuuint64_t)d->src_stride_b * elt
                                         : cstride * d->channels};
  for (int i = 0; i < 3; ++i)
    if (strides[i] == 0 || strides[i] % 16) return;
  for (int m = 0; m < 16; ++m) p.tma_bw[m] = p.tma_bh[m] = 0;
  for (int m = 0; m < kL2Boxes; ++m) {
    const int bw = l2_box_w(elt, m);
    const int bh = l2_box_h(elt, d->channels, m, d->mode == EQB_BICUBIC);
    if (bw > 256 || bh < kL2MinBoxH) continue;
    const cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)d->channels, 1};
    if (encode_cached(&p.tmaps[m], dt, const_cast<void*>(d->src), dims, strides, box) !=
        CUDA_SUCCESS)
      return;
    p.tma_bw[m] = bw;
    p.tma_bh[m] = bh;
  }
  p.tma_batched = batched ? 1 : 0;
  p.use_tma = 1;
}

template <int XF> static cudaError_t go(Params& p, const eqb_remap_desc* d, int px,
                                        cudaStream_t s) {
  if (can_lean2(d)) {
    setup_tma_l2(d, p);
    if (p.use_tma) {
      const bool exact = (d->flags & EQB_FLAG_EXACT_COORDS) || XF == EQB_PERS2EQUI ||
                         XF == EQB_CUBE2EQUI ||
                         (d->dtype == EQB_F32 && !(d->flags & EQB_FLAG_FAST_COORDS));
      return launch_lean2<XF>(p, d->dtype, exact, d->mode, s);
    }
  }
  if (d->flags & EQB_FLAG_CHANNELS_LAST) {
    // (eqb_remap has checked can_nhwc)
    setup_tma_nhwc(d, p);
    if (!p.use_tma) return cudaErrorNotSupported;
    return launch_nhwc<XF>(p, s);
  }
  if (XF == EQB_CUBE2EQUI && can_stage(d) && !cube_is_horizon(d)) {
    setup_tma_faces(d, p);
    if (p.use_tma) return launch_staged<XF>(p, d->mode, d->dtype, false, false, s);
    return launch_remap<XF>(p, d->mode, d->dtype, px, pick_ws(d), s);
  }
  if (can_stage(d)) {
    bool lean = can_lean(d);
    setup_tma(d, p, lean);
    if (lean && !p.use_tma) {  // no encoder / unaligned strides: generic boxes
      lean = false;
      setup_tma(d, p, false);
    }
    return launch_staged<XF>(p, d->mode, d->dtype, can_fast(d), lean, s);
  }
  return launch_remap<XF>(p, d->mode, d->dtype, px, pick_ws(d), s);
}

// ---------------------------------------------------------------------------
// min / max reduction (clip_output)
// ---------------------------------------------------------------------------

__device__ __forceinline__ unsigned enc(float f) {  // order-preserving float -> uint
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec(unsigned e) {
  return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e);
}

__global__ void minmax_init(unsigned* out) {
  out[0] = 0xffffffffu;
  out[1] = 0u;
}
__global__ void minmax_fini(unsigned* out) {
  const float lo = dec(out[0]), hi = dec(out[1]);
  reinterpret_cast<float*>(out)[0] = lo;
  reinterpret_cast<float*>(out)[1] = hi;
}

// One row-segment per thread iteration; rows are contiguous, so consecutive threads
// read consecutive elements.  Grid-stride over (plane-row, x-chunk).
template <typename T>
__global__ void __launch_bounds__(256) minmax_kernel(const T* __restrict__ data, int B, int C,
                                                     int H, int W, long long sb, long long sc,
                                                     long long sh, unsigned* out, bool vec) {
  float lo = INFINITY, hi = -INFINITY;
  const long long rows = (long long)B * C * H;
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const int y = (int)(r % H);
    const long long bc = r / H;
    const int c = (int)(bc % C);
    const long long b = bc / C;
    const T* row = data + b * sb + c * sc + y * sh;
    if (vec) {
      // 16 bytes per thread and load (rows are 16-byte aligned, W a multiple of 16 / sizeof(T)):
      // the scalar loop below is issue-bound at a fifth of the HBM rate
      constexpr int PER = 16 / (int)sizeof(T);
      const uint4* row4 = reinterpret_cast<const uint4*>(row);
      for (int x = threadIdx.x; x < W / PER; x += blockDim.x) {
        const uint4 q = __ldg(row4 + x);
        const T* e = reinterpret_cast<const T*>(&q);
#pragma unroll
        for (int i = 0; i < PER; ++i) {
          float v;
          if constexpr (sizeof(T) == 4) v = __uint_as_float(reinterpret_cast<const unsigned*>(&q)[i]);
          else if constexpr (sizeof(T) == 1) v = (float)reinterpret_cast<const unsigned char*>(&q)[i];
          else if constexpr (cuda::std::is_same<T, __half>::value) v = __half2float(e[i]);
          else v = __bfloat162float(e[i]);
          lo = fminf(lo, v);
          hi = fmaxf(hi, v);
        }
      }
    } else {
      for (int x = threadIdx.x; x < W; x += blockDim.x) {
        const float v = ld_f(row + x);
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __shared__ float slo[8], shi[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { slo[warp] = lo; shi[warp] = hi; }
  __syncthreads();
  if (warp == 0) {
    lo = lane < 8 ? slo[lane] : INFINITY;
    hi = lane < 8 ? shi[lane] : -INFINITY;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) {
      atomicMin(out, enc(lo));
      atomicMax(out + 1, enc(hi));
    }
  }
}

}  // namespace eqb

using namespace eqb;

extern "C" {

int32_t eqb_abi_version(void) { return EQB_ABI_VERSION; }

const char* eqb_last_error(void) { return g_err; }

int64_t eqb_launch_count(void) { return g_launches.load(); }

#define EQB_STR2(x) #x
#define EQB_STR(x) EQB_STR2(x)
#ifndef EQB_SRC_HASH
#define EQB_SRC_HASH
#endif
// ("EQB_SRC_HASH=<digest>" is also what build.py::built_hash looks for in the file)
const char* eqb_source_hash(void) {
  static const char tagged[] = "EQB_SRC_HASH=" EQB_STR(EQB_SRC_HASH);
  return tagged + 13;
}

int32_t eqb_remap(const eqb_remap_desc* d, void* stream) {
  Params p;
  const int rc = fill(d, true, p);
  if (rc != EQB_OK) return rc;
  if ((d->flags & EQB_FLAG_CHANNELS_LAST) && !can_nhwc(d))
    return fail(EQB_ERR_UNSUPPORTED,
                "channels-last layout is served for uint8 RGB/RGBX bilinear equi2equi/equi2pers "
                "on large aligned launches only (eqb_remap_supports_channels_last)");
  const int px = pick_px(d);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  switch (d->transform) {
    case EQB_EQUI2EQUI: e = go<EQB_EQUI2EQUI>(p, d, px, s); break;
    case EQB_EQUI2PERS: e = go<EQB_EQUI2PERS>(p, d, px, s); break;
    case EQB_EQUI2CUBE: e = go<EQB_EQUI2CUBE>(p, d, px, s); break;
    case EQB_CUBE2EQUI: e = go<EQB_CUBE2EQUI>(p, d, px, s); break;
    case EQB_PERS2EQUI: e = go<EQB_PERS2EQUI>(p, d, px, s); break;
    default: e = go<EQB_GRID>(p, d, px, s); break;
  }
  if (e != cudaSuccess)
    return fail(EQB_ERR_CUDA, "remap launch failed: %s", cudaGetErrorString(e));
  g_launches.fetch_add(1);
  g_err[0] = 0;
  return EQB_OK;
}

int32_t eqb_remap_supports_channels_last(const eqb_remap_desc* d) {
  Params p;
  if (fill(d, true, p) != EQB_OK) {
    g_err[0] = 0;
    return 0;
  }
  return can_nhwc(d) ? 1 : 0;
}

static_assert(EQB_MAP_TILE_W == 64 && EQB_MAP_TILE_H == kLeanTileH, "map tile == lean tile");

static int map_box_table(const eqb_remap_desc* d, Params& p) {
  const int elt = (int)elt_size(d->dtype);
  if (!elt || elt > 2) return fail(EQB_ERR_UNSUPPORTED, "static maps: uint8 / fp16 / bf16 images");
  if (d->channels != 1 && d->channels != 3 && d->channels != 4)
    return fail(EQB_ERR_UNSUPPORTED, "static maps: 1, 3 or 4 channels");
  if (d->src_w > 65536 || d->src_h > 65536)
    return fail(EQB_ERR_UNSUPPORTED, "static maps: source sides up to 65536");
  for (int m = 0; m < 16; ++m) p.tma_bw[m] = p.tma_bh[m] = 0;
  for (int m = 0; m < kTmaBoxes; ++m) {
    p.tma_bw[m] = box_w(elt, m);
    p.tma_bh[m] = lean_box_h(elt, m, d->channels);
  }
  return EQB_OK;
}

int32_t eqb_map_build(const eqb_remap_desc* d, const float* coords, void* map, void* heads,
                      void* stream) {
  Params p;
  int rc = fill(d, false, p);
  if (rc != EQB_OK) return rc;
  if (!coords || !map || !heads) return fail(EQB_ERR_INVALID, "null argument");
  rc = map_box_table(d, p);
  if (rc != EQB_OK) return rc;
  const cudaError_t e = launch_map_pack(p, coords, map, heads, (int)elt_size(d->dtype),
                                        static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess)
    return fail(EQB_ERR_CUDA, "map build launch failed: %s", cudaGetErrorString(e));
  g_launches.fetch_add(1);
  g_err[0] = 0;
  return EQB_OK;
}

int32_t eqb_remap_mapped(const eqb_remap_desc* d, const void* map, const void* heads,
                         void* stream) {
  Params p;
  int rc = fill(d, true, p);
  if (rc != EQB_OK) return rc;
  if (!map || !heads) return fail(EQB_ERR_INVALID, "null argument");
  rc = map_box_table(d, p);
  if (rc != EQB_OK) return rc;
  if (d->mode != EQB_BILINEAR || d->clip)
    return fail(EQB_ERR_UNSUPPORTED, "static maps: bilinear, no clamp");
  if (d->dst_w % 4) return fail(EQB_ERR_UNSUPPORTED, "static maps: output width % 4 == 0");
  if (d->flags & EQB_FLAG_CHANNELS_LAST) {
    // uint8 RGB / RGBX frames in the decoder's layout: the alignment rules of can_nhwc
    const long long c = d->channels, va = c == 4 ? 16 : 4;
    if (d->dtype != EQB_U8 || (c != 3 && c != 4) || d->src_stride_c != 1 || d->dst_stride_c != 1 ||
        d->src_stride_h < c * d->src_w || d->dst_stride_h < c * d->dst_w ||
        reinterpret_cast<uintptr_t>(d->src) % 16 || d->src_stride_h % 16 || d->src_stride_b % 16 ||
        (c * d->src_w) % 4 || reinterpret_cast<uintptr_t>(d->dst) % va || d->dst_stride_h % va ||
        d->dst_stride_b % va || d->src_w < 4 || d->src_h < 4)
      return fail(EQB_ERR_UNSUPPORTED,
                  "static maps, channels-last: uint8 RGB / RGBX with 16-byte aligned source rows");
    setup_tma_nhwc(d, p);
  } else {
    const long long per16 = 16 / (long long)elt_size(d->dtype);
    if (reinterpret_cast<uintptr_t>(d->src) % 16 || d->src_stride_b % per16 ||
        d->src_stride_c % per16 || d->src_stride_h % per16 || d->src_w < 4 || d->src_h < 4)
      return fail(EQB_ERR_UNSUPPORTED, "static maps: 16-byte aligned source rows and planes");
    if (!dst_vector_ok(d, 4))
      return fail(EQB_ERR_UNSUPPORTED, "static maps: 4-pixel aligned destination rows");
    setup_tma(d, p, true);
  }
  if (!p.use_tma) return fail(EQB_ERR_UNSUPPORTED, "static maps: no tensor-map encoder");
  const cudaError_t e = launch_mapped(p, d->dtype, map, heads, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess)
    return fail(EQB_ERR_CUDA, "mapped remap launch failed: %s", cudaGetErrorString(e));
  g_launches.fetch_add(1);
  g_err[0] = 0;
  return EQB_OK;
}

int32_t eqb_equi2cube2equi(const eqb_remap_desc* to_cube, const eqb_remap_desc* to_equi,
                           void* stream) {
  if (!to_cube || !to_equi) return fail(EQB_ERR_INVALID, "null descriptor");
  if (to_cube->transform != EQB_EQUI2CUBE || to_equi->transform != EQB_CUBE2EQUI)
    return fail(EQB_ERR_INVALID, "descriptors must be EQUI2CUBE followed by CUBE2EQUI");
  Params pe, pc;
  int rc = fill(to_cube, false, pe);
  if (rc != EQB_OK) return rc;
  rc = fill(to_equi, false, pc);
  if (rc != EQB_OK) return rc;
  const eqb_remap_desc *a = to_cube, *b = to_equi;
  if (a->mode != b->mode || a->dtype != b->dtype || a->batch != b->batch ||
      a->channels != b->channels || a->w_face != b->w_face)
    return fail(EQB_ERR_INVALID, "the two stages differ in mode, dtype, batch, channels or w_face");
  if (a->mode != EQB_NEAREST && a->mode != EQB_BILINEAR)
    return fail(EQB_ERR_UNSUPPORTED, "fused cube chain: nearest and bilinear");
  if (a->clip || b->clip)
    return fail(EQB_ERR_UNSUPPORTED, "fused cube chain: no clamp (run the two launches)");
  if (!elt_size(a->dtype)) return fail(EQB_ERR_INVALID, "unknown dtype %d", a->dtype);
  if (a->channels <= 0 || a->channels > 4)
    return fail(EQB_ERR_UNSUPPORTED, "fused cube chain: 1 to 4 channels");
  if (!a->src || !b->dst) return fail(EQB_ERR_INVALID, "source / destination is null");
  if (a->src_stride_h < a->src_w) return fail(EQB_ERR_INVALID, "src_stride_h smaller than a row");
  if ((long long)a->src_stride_h * a->src_h >= (1ll << 31))
    return fail(EQB_ERR_UNSUPPORTED, "a source plane spans more than 2^31 elements");
  pe.src = a->src;
  pe.ssb = a->src_stride_b;
  pe.ssc = a->src_stride_c;
  pe.ssh = a->src_stride_h;
  pc.dst = b->dst;
  pc.dsb = b->dst_stride_b;
  pc.dsc = b->dst_stride_c;
  pc.dsh = b->dst_stride_h;
  const cudaError_t e =
      launch_chain(pe, pc, a->mode, a->dtype, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess)
    return fail(EQB_ERR_CUDA, "fused cube chain launch failed: %s", cudaGetErrorString(e));
  g_launches.fetch_add(1);
  g_err[0] = 0;
  return EQB_OK;
}

int32_t eqb_remap_coords(const eqb_remap_desc* d, float* coords_out, void* stream) {
  Params p;
  const int rc = fill(d, false, p);
  if (rc != EQB_OK) return rc;
  if (!coords_out) return fail(EQB_ERR_INVALID, "coords_out is null");
  p.coords_out = coords_out;
  const bool libm = (d->flags & EQB_FLAG_LIBM) != 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  switch (d->transform) {
    case EQB_EQUI2EQUI: e = launch_coords<EQB_EQUI2EQUI>(p, libm, s); break;
    case EQB_EQUI2PERS: e = launch_coords<EQB_EQUI2PERS>(p, libm, s); break;
    case EQB_EQUI2CUBE: e = launch_coords<EQB_EQUI2CUBE>(p, libm, s); break;
    case EQB_CUBE2EQUI: e = launch_coords<EQB_CUBE2EQUI>(p, libm, s); break;
    case EQB_PERS2EQUI: e = launch_coords<EQB_PERS2EQUI>(p, libm, s); break;
    default: e = launch_coords<EQB_GRID>(p, libm, s); break;
  }
  if (e != cudaSuccess)
    return fail(EQB_ERR_CUDA, "coords launch failed: %s", cudaGetErrorString(e));
  g_launches.fetch_add(1);
  g_err[0] = 0;
  return EQB_OK;
}

int32_t eqb_minmax(const eqb_image* im, float* out2, void* stream) {
  if (!im || !im->data || !out2) return fail(EQB_ERR_INVALID, "null argument");
  if (im->batch <= 0 || im->channels <= 0 || im->height <= 0 || im->width <= 0)
    return fail(EQB_ERR_INVALID, "empty image");
  if (!elt_size(im->dtype)) return fail(EQB_ERR_INVALID, "unknown dtype %d", im->dtype);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  unsigned* o = reinterpret_cast<unsigned*>(out2);
  const long long rows = (long long)im->batch * im->channels * im->height;
  const int blocks = (int)(rows < 148 * 16 ? rows : 148 * 16);
  const long long per16 = 16 / (long long)elt_size(im->dtype);
  const bool vec = reinterpret_cast<uintptr_t>(im->data) % 16 == 0 && im->width % per16 == 0 &&
                   im->stride_b % per16 == 0 && im->stride_c % per16 == 0 &&
                   im->stride_h % per16 == 0;
  minmax_init<<<1, 1, 0, s>>>(o);
#define EQB_MM(T)                                                                          \
  minmax_kernel<T><<<blocks, 256, 0, s>>>(static_cast<const T*>(im->data), im->batch,     \
                                          im->channels, im->height, im->width,            \
                                          im->stride_b, im->stride_c, im->stride_h, o, vec)
  switch (im->dtype) {
    case EQB_U8: EQB_MM(uint8_t); break;
    case EQB_F16: EQB_MM(__half); break;
    case EQB_BF16: EQB_MM(__nv_bfloat16); break;
    default: EQB_MM(float); break;
  }
#undef EQB_MM
  minmax_fini<<<1, 1, 0, s>>>(o);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return fail(EQB_ERR_CUDA, "minmax launch failed: %s", cudaGetErrorString(e));
  g_launches.fetch_add(3);
  g_err[0] = 0;
  return EQB_OK;
}

int32_t eqb_rotation_compose(const float* trig, int32_t n, float* out) {
  if (!trig || !out || n < 0) return fail(EQB_ERR_INVALID, "null argument");
  // C = A @ B for 3x3 fp32 as the FMA chain k = 0, 1, 2 that torch's 2-D mm
  // produces for equilib/torch_utils/rotation.py:77 (SURVEY probe B15).
  auto mm = [](const float* A, const float* B, float* C) {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        float acc = A[i * 3 + 0] * B[0 * 3 + j];
        acc = fmaf(A[i * 3 + 1], B[1 * 3 + j], acc);
        acc = fmaf(A[i * 3 + 2], B[2 * 3 + j], acc);
        C[i * 3 + j] = acc;
      }
  };
  for (int k = 0; k < n; ++k) {
    const float* t = trig + (size_t)k * 6;
    const float cr = t[0], sr = t[1], cp = t[2], sp = t[3], cy = t[4], sy = t[5];
    const float Rx[9] = {1.f, 0.f, 0.f, 0.f, cr, -sr, 0.f, sr, cr};
    const float Ry[9] = {cp, 0.f, sp, 0.f, 1.f, 0.f, -sp, 0.f, cp};
    const float Rz[9] = {cy, -sy, 0.f, sy, cy, 0.f, 0.f, 0.f, 1.f};
    float T[9];
    mm(Rz, Ry, T);
    mm(T, Rx, out + (size_t)k * 9);
  }
  return EQB_OK;
}

}  // extern "C"
