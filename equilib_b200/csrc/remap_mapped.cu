// Static-camera path: the sampling map of a fixed set of rotations is built
// once -- per output pixel the top-left tap (x, y) and the two bilinear fractions, 8 bytes;
// per 64 x 8 tile the shared-memory box its footprint fits -- and every later frame is
// remapped from it: no ray, no asin / atan2, no interpolation, no bounding box.  The
// reference's counterpart is its grid cache (equilib/_cache.py:21-40), which keeps the
// rays but recomputes rotation, spherical mapping and normalisation every call; its
// README.md:168 lists the full-map cache as future work.
//
//   eqb_map_build   coordinates (eqb_remap_coords output) -> map + tile headers
//   eqb_remap_mapped   frames + map -> frames   (bilinear, 1 / 3 / 4 planar channels of
//                      1- or 2-byte pixels, the lean kernel's blend)
#include "remap_kernels.cuh"

namespace eqb {

constexpr unsigned kFracOne = 65535u;  // fractions are stored as round(f * 65535)

// ---- build -------------------------------------------------------------------------
// One CTA per 64 x 8 tile, lanes laid out as in the lean kernel (16 x 2 lanes per warp, a
// lane owns 4 consecutive pixels).  Coordinates go through the same to_index / floor /
// footprint shift as the remap kernels, so the taps are the ones the exact path uses.
// COMPACT: 4 bytes per pixel -- the tap as an offset from the tile's box origin (x: 8 bits,
// y: 8 bits; boxes are at most 256 x 128) and the two fractions as round(f * 255) -- written
// after the tile's box is known.  Tiles without a box store nothing usable: the remap
// kernel evaluates the exact chain for them (source poles, the seam: a few per cent).
template <bool COMPACT>
__global__ void __launch_bounds__(kLeanWarps * 32)
    map_pack_kernel(const __grid_constant__ Params p, const float* __restrict__ coords,
                    void* __restrict__ map_out, int4* __restrict__ heads, int per16) {
  uint2* const map = static_cast<uint2*>(map_out);
  __shared__ int bbox[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lx = lane % kStLaneW, ly = lane / kStLaneW;
  const int b = blockIdx.z;
  const int yq = blockIdx.y * kLeanTileH + warp * kStLaneH + ly;
  const int x0 = blockIdx.x * 64 + lx * 4;
  if (tid == 0) { bbox[0] = 0x7fffffff; bbox[1] = 0x7fffffff; bbox[2] = -1; bbox[3] = -1; }
  __syncthreads();
  int nx = 0x7fffffff, ny = 0x7fffffff, mx = -1, my = -1;
  int cxi[4] = {0, 0, 0, 0}, cyi[4] = {0, 0, 0, 0};
  unsigned cq[4] = {0, 0, 0, 0};
  if (yq < p.dh) {
    const long long plane = (long long)p.dh * p.dw;
    const float* cj = coords + (long long)b * 2 * plane + (long long)yq * p.dw;
    const float* ci = cj + plane;
    uint2* mrow = map + ((long long)b * p.dh + yq) * p.dw;
    for (int i = 0; i < 4; ++i) {
      const int x = x0 + i;
      if (x >= p.dw) break;
      const float ix = to_index(__ldg(ci + x), p.inv_wm1, p.swm1f);
      const float iy = to_index(__ldg(cj + x), p.inv_hm1, p.shm1f);
      const float xf = fminf(floorf(ix), p.swm1f - 1.0f), yf = fminf(floorf(iy), p.shm1f - 1.0f);
      const float fx = __fsub_rn(ix, xf), fy = __fsub_rn(iy, yf);
      const int xi = (int)xf, yi = (int)yf;
      if (COMPACT) {
        cxi[i] = xi; cyi[i] = yi;
        cq[i] = (unsigned)__float2int_rn(fx * 255.0f) | ((unsigned)__float2int_rn(fy * 255.0f) << 8);
      } else {
        const unsigned qx = (unsigned)__float2int_rn(fx * (float)kFracOne);
        const unsigned qy = (unsigned)__float2int_rn(fy * (float)kFracOne);
        mrow[x] = make_uint2((unsigned)xi | ((unsigned)yi << 16), qx | (qy << 16));
      }
      nx = min(nx, xi); ny = min(ny, yi); mx = max(mx, xi); my = max(my, yi);
    }
  }
  nx = __reduce_min_sync(0xffffffffu, nx);
  ny = __reduce_min_sync(0xffffffffu, ny);
  mx = __reduce_max_sync(0xffffffffu, mx);
  my = __reduce_max_sync(0xffffffffu, my);
  if (lane == 0) {
    atomicMin(&bbox[0], nx); atomicMin(&bbox[1], ny);
    atomicMax(&bbox[2], mx); atomicMax(&bbox[3], my);
  }
  __syncthreads();
  if (tid == 0) {
    // the footprint is exact here (every tap is known): no margin
    const int gx0 = bbox[0] & ~(per16 - 1), gy0 = bbox[1];
    const int need_w = bbox[2] + 2 - gx0, need_h = bbox[3] + 2 - gy0;
    int m = -1;
    if (bbox[2] >= 0 && bbox[2] - bbox[0] < (p.sw >> 1))
      for (int k = kTmaBoxes - 1; k >= 0; --k)
        if (need_w <= p.tma_bw[k] && need_h <= p.tma_bh[k]) m = k;
    const long long tile =
        ((long long)b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    heads[tile] = make_int4(gx0, gy0, m, 0);
    bbox[0] = gx0;
    bbox[1] = gy0;
    bbox[2] = m;
  }
  if (COMPACT) {
    __syncthreads();
    if (yq < p.dh && bbox[2] >= 0) {
      unsigned* crow = static_cast<unsigned*>(map_out) + ((long long)b * p.dh + yq) * p.dw;
      for (int i = 0; i < 4 && x0 + i < p.dw; ++i)
        crow[x0 + i] = (unsigned)(cxi[i] - bbox[0]) | ((unsigned)(cyi[i] - bbox[1]) << 8) |
                       (cq[i] << 16);
    }
  }
}

// ---- remap from the map ------------------------------------------------------------------
// IL = 0: NC planar channels.  IL = 3 / 4: channels-last uint8 RGB / RGBX (NC == IL), the
// layout of decoded video frames, read and written in place (see nhwc_blend).
template <typename T, int NC, int IL = 0, bool CM = false>
__global__ void __launch_bounds__(kLeanWarps * 32, 32 / kLeanWarps)
    remap_mapped_kernel(const __grid_constant__ Params p, const void* __restrict__ map_in,
                        const int4* __restrict__ heads, int tiles_x) {
  const uint2* const map = static_cast<const uint2*>(map_in);
  constexpr int PX = 4;
  constexpr int ELT = (int)sizeof(T);
  constexpr int PXE = IL ? IL : 1;  // elements from a pixel to its right neighbour
  static_assert(IL == 0 || (ELT == 1 && NC == IL), "channels-last: uint8 RGB / RGBX");
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long tma_bar;
  unsigned tma_phase = 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) mbar_init(&tma_bar, 1);
  __syncthreads();
  const int lx = lane % kStLaneW, ly = lane / kStLaneW;
  const int b = blockIdx.z;
  const int yq = blockIdx.y * kLeanTileH + warp * kStLaneH + ly;
  const int y = min(yq, p.dh - 1);
  const unsigned tile_s = smem_u32(smem_raw);
  const T* const src_b = static_cast<const T*>(p.src) + (long long)b * p.ssb;
  T* const dst_row = static_cast<T*>(p.dst) + (long long)b * p.dsb + (long long)y * p.dsh;
  const uint2* const map_row = map + ((long long)b * p.dh + y) * p.dw;
  const unsigned* const cmap_row =
      static_cast<const unsigned*>(map_in) + ((long long)b * p.dh + y) * p.dw;
  const int4* const head_row = heads + ((long long)b * gridDim.y + blockIdx.y) * tiles_x;
  const int ssh = (int)p.ssh;
  const float q = 1.0f / (float)kFracOne;

  for (int it = 0; it < p.iters; ++it) {
    const int tx = blockIdx.x * p.iters + it;
    if (tx >= tiles_x) break;  // uniform
    const int x0 = tx * 64 + lx * PX;
    const bool live = yq < p.dh && x0 < p.dw;
    const int4 hd = __ldg(head_row + tx);
    const int gx0 = hd.x, gy0 = hd.y, m = hd.z;
    const bool boxed = m >= 0;
    const int pitch = p.tma_bw[boxed ? m : 0], box_h = p.tma_bh[boxed ? m : 0];
    __syncthreads();  // every warp is done reading the previous tile's box
    if (boxed && tid == 0) {
      mbar_expect_tx(&tma_bar, (unsigned)(pitch * box_h * NC * ELT));
      // (channels-last maps are rows of 32-bit words)
      tma_load_box(smem_raw, &p.tmaps[m], &tma_bar, IL ? gx0 * IL / 4 : gx0, gy0, 0,
                   p.tma_batched ? b : 0);
    }
    unsigned a[PX];
    float fx[PX], fy[PX];
    if constexpr (CM) {
      if (boxed) {
        unsigned e[PX];
        if (live && x0 + PX <= p.dw) {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(cmap_row + x0));
          e[0] = v.x; e[1] = v.y; e[2] = v.z; e[3] = v.w;
        } else {
#pragma unroll
          for (int i = 0; i < PX; ++i) e[i] = __ldg(cmap_row + min(x0 + i, p.dw - 1));
        }
#pragma unroll
        for (int i = 0; i < PX; ++i) {
          const int xr = (int)(e[i] & 0xffu), yr = (int)((e[i] >> 8) & 0xffu);
          fx[i] = (float)((e[i] >> 16) & 0xffu) * (1.0f / 255.0f);
          fy[i] = (float)(e[i] >> 24) * (1.0f / 255.0f);
          a[i] = tile_s + (unsigned)((yr * pitch + xr) * (PXE * ELT));
        }
      } else {
        // no box, no map entry: the exact chain (source poles, the seam)
#pragma unroll
        for (int i = 0; i < PX; ++i) {
          const int x = min(x0 + i, p.dw - 1);
          float uj, ui;
          if (p.xf == EQB_EQUI2EQUI) {
            RowCtx<EQB_EQUI2EQUI> rc;
            row_setup<EQB_EQUI2EQUI>(p, b, y, rc);
            coords<EQB_EQUI2EQUI, false>(p, rc, b, x, y, 0, 0, uj, ui);
          } else {
            RowCtx<EQB_EQUI2PERS> rc;
            row_setup<EQB_EQUI2PERS>(p, b, y, rc);
            coords<EQB_EQUI2PERS, false>(p, rc, b, x, y, 0, 0, uj, ui);
          }
          const float ix = to_index(ui, p.inv_wm1, p.swm1f), iy = to_index(uj, p.inv_hm1, p.shm1f);
          const float xf = fminf(floorf(ix), p.swm1f - 1.0f), yf = fminf(floorf(iy), p.shm1f - 1.0f);
          fx[i] = __fsub_rn(ix, xf);
          fy[i] = __fsub_rn(iy, yf);
          a[i] = (unsigned)((int)yf * ssh + (int)xf * PXE);
        }
      }
    } else {
      uint2 e[PX];
      if (live && x0 + PX <= p.dw) {
        const uint4 e01 = __ldg(reinterpret_cast<const uint4*>(map_row + x0));
        const uint4 e23 = __ldg(reinterpret_cast<const uint4*>(map_row + x0) + 1);
        e[0] = make_uint2(e01.x, e01.y); e[1] = make_uint2(e01.z, e01.w);
        e[2] = make_uint2(e23.x, e23.y); e[3] = make_uint2(e23.z, e23.w);
      } else {
#pragma unroll
        for (int i = 0; i < PX; ++i) e[i] = __ldg(map_row + min(x0 + i, p.dw - 1));
      }
#pragma unroll
      for (int i = 0; i < PX; ++i) {
        const int xi = (int)(e[i].x & 0xffffu), yi = (int)(e[i].x >> 16);
        fx[i] = (float)(e[i].y & 0xffffu) * q;
        fy[i] = (float)(e[i].y >> 16) * q;
        a[i] = boxed ? tile_s + (unsigned)(((yi - gy0) * pitch + (xi - gx0)) * (PXE * ELT))
                     : (unsigned)(yi * ssh + xi * PXE);
      }
    }
    if (boxed) {
      mbar_wait(&tma_bar, tma_phase);
      tma_phase ^= 1u;
    }
    if (!live) continue;
    T* dst = dst_row + x0 * PXE;
    const int npx = min(PX, p.dw - x0);
    if (boxed && npx == PX) {
      auto blend_m = [&](auto mc) {
        constexpr int M = decltype(mc)::value;
        if constexpr (IL != 0) nhwc_blend<IL, M>(a, fx, fy, dst, p.flags);
        else lean_blend<T, M, NC>(a, fx, fy, dst, p.dsc, p.flags);
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
        default: blend_m(cuda::std::integral_constant<int, 11>{}); break;
      }
      continue;
    }
    // ragged right edge (taps from the box, element loads) or no box (global gathers)
    float w_nw[PX], w_ne[PX], w_sw[PX], w_se[PX];
#pragma unroll
    for (int i = 0; i < PX; ++i) {
      const float ax = __fsub_rn(1.0f, fx[i]), ay = __fsub_rn(1.0f, fy[i]);
      w_nw[i] = __fmul_rn(ax, ay);
      w_ne[i] = __fmul_rn(fx[i], ay);
      w_sw[i] = __fmul_rn(ax, fy[i]);
      w_se[i] = __fmul_rn(fx[i], fy[i]);
    }
    const T* src = src_b;
    const T* sm = reinterpret_cast<const T*>(smem_raw);
    for (int c = 0; c < NC; ++c, src += (IL ? 1 : p.ssc), sm += (IL ? 1 : pitch * box_h)) {
      float o[PX];
#pragma unroll
      for (int i = 0; i < PX; ++i) {
        o[i] = 0.0f;
        if (i >= npx) continue;
        if (boxed) {
          const T* q0 = sm + (a[i] - tile_s) / ELT;
          const T* q1 = q0 + pitch * PXE;
          o[i] = blend4(lds_f(q0), lds_f(q0 + PXE), lds_f(q1), lds_f(q1 + PXE), w_nw[i],
                        w_ne[i], w_sw[i], w_se[i]);
        } else if (IL == 0) {
          const T* q0 = src + (int)a[i];
          float v_nw, v_ne, v_sw, v_se;
          ldg_pair<T>(q0, v_nw, v_ne);
          ldg_pair<T>(q0 + ssh, v_sw, v_se);
          o[i] = blend4(v_nw, v_ne, v_sw, v_se, w_nw[i], w_ne[i], w_sw[i], w_se[i]);
        } else {
          const T* q0 = src + (int)a[i];
          const T* q1 = q0 + ssh;
          o[i] = blend4(ld_f(q0), ld_f(q0 + PXE), ld_f(q1), ld_f(q1 + PXE), w_nw[i], w_ne[i],
                        w_sw[i], w_se[i]);
        }
      }
      if (IL) {
#pragma unroll
        for (int i = 0; i < PX; ++i) {
          const T v = OutCvt<T>::cvt(o[i], p.flags);
          if (i < npx) Pack<T, 1>::st(dst + i * PXE + c, &v);
        }
      } else {
        emit<T, PX>(dst + c * p.dsc, o, npx, p.flags);
      }
    }
  }
}

cudaError_t launch_map_pack(const Params& p, const float* coords, void* map, void* heads,
                            int elt, cudaStream_t stream) {
  const dim3 grid((unsigned)((p.dw + 63) / 64), (unsigned)((p.dh + kLeanTileH - 1) / kLeanTileH),
                  (unsigned)p.B);
  if (p.flags & EQB_FLAG_COMPACT_MAP)
    map_pack_kernel<true><<<grid, kLeanWarps * 32, 0, stream>>>(p, coords, map,
                                                                static_cast<int4*>(heads), 16 / elt);
  else
    map_pack_kernel<false><<<grid, kLeanWarps * 32, 0, stream>>>(
        p, coords, map, static_cast<int4*>(heads), 16 / elt);
  return cudaGetLastError();
}

template <typename T, int NC, int IL = 0>
static cudaError_t launch_mapped_nc(Params p, const void* map, const void* heads,
                                    cudaStream_t stream) {
  const int tiles_x = (p.dw + 63) / 64;
  const long long tiles_y = (p.dh + kLeanTileH - 1) / kLeanTileH;
  const long long tiles = tiles_x * tiles_y * p.B;
  const long long wave = 148ll * 2048 / (kLeanWarps * 32);
  p.iters = tiles >= wave * 8 * 16 ? 16 : tiles >= wave * 16 ? 4 : 1;
  const dim3 grid((unsigned)((tiles_x + p.iters - 1) / p.iters), (unsigned)tiles_y, (unsigned)p.B);
  // (+16: the channels-last RGB tap loads are aligned words that may end past the last tap)
  if (p.flags & EQB_FLAG_COMPACT_MAP)
    remap_mapped_kernel<T, NC, IL, true><<<grid, kLeanWarps * 32, lean_smem_bytes() + 16, stream>>>(
        p, map, static_cast<const int4*>(heads), tiles_x);
  else
    remap_mapped_kernel<T, NC, IL, false><<<grid, kLeanWarps * 32, lean_smem_bytes() + 16, stream>>>(
        p, map, static_cast<const int4*>(heads), tiles_x);
  return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_mapped_t(const Params& p, const void* map, const void* heads,
                                   cudaStream_t stream) {
  switch (p.C) {
    case 1: return launch_mapped_nc<T, 1>(p, map, heads, stream);
    case 3: return launch_mapped_nc<T, 3>(p, map, heads, stream);
    case 4: return launch_mapped_nc<T, 4>(p, map, heads, stream);
  }
  return cudaErrorNotSupported;
}

cudaError_t launch_mapped(const Params& p, int dtype, const void* map, const void* heads,
                          cudaStream_t stream) {
  if (p.flags & EQB_FLAG_CHANNELS_LAST) {
    if (dtype == EQB_U8 && p.C == 3) return launch_mapped_nc<uint8_t, 3, 3>(p, map, heads, stream);
    if (dtype == EQB_U8 && p.C == 4) return launch_mapped_nc<uint8_t, 4, 4>(p, map, heads, stream);
    return cudaErrorNotSupported;
  }
  switch (dtype) {
    case EQB_U8: return launch_mapped_t<uint8_t>(p, map, heads, stream);
    case EQB_F16: return launch_mapped_t<__half>(p, map, heads, stream);
    case EQB_BF16: return launch_mapped_t<__nv_bfloat16>(p, map, heads, stream);
  }
  return cudaErrorNotSupported;
}

}  // namespace eqb
