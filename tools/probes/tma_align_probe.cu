// Probe (round 2): which start coordinates does a tiled TMA box load accept on B200?
//  (a) innermost coordinate not a multiple of 16 bytes (odd pixel columns) for 1/2/4-byte types
//  (b) negative / out-of-range coordinates (zero fill)
//  (c) a tensor map whose dimension order is (W, C, H): strides not monotonic
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_align_probe tma_align_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int ELT>
__global__ void k_load(const __grid_constant__ CUtensorMap map, int x, int y, int c, int box_elems,
                       unsigned char* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(box_elems * ELT) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(smem)), "l"(&map), "r"(smem_u32(&bar)), "r"(x), "r"(y), "r"(c) : "memory");
  }
  unsigned ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
  }
  for (int i = threadIdx.x; i < box_elems * ELT; i += blockDim.x) out[i] = smem[i];
}

template <int ELT> int run(EncodeTiledFn enc, CUtensorMapDataType dt, const char* name) {
  const int W = 512, H = 64, C = 3;
  std::vector<unsigned char> h((size_t)W * H * C * ELT);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)((i * 131u + (i >> 8) * 7u) & 0xff);
  unsigned char *d, *dout;
  CK(cudaMalloc(&d, h.size()));
  CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice));
  const int bw = 48, bh = 8;
  CK(cudaMalloc(&dout, bw * bh * 3 * ELT));
  // (W, H, C) order and (W, C, H) order
  for (int order = 0; order < 2; ++order) {
    CUtensorMap map;
    cuuint64_t dims[3], strides[2];
    cuuint32_t box[3], ones[3] = {1, 1, 1};
    if (order == 0) {
      dims[0] = W; dims[1] = H; dims[2] = C;
      strides[0] = (cuuint64_t)W * ELT; strides[1] = (cuuint64_t)W * H * ELT;
      box[0] = bw; box[1] = bh; box[2] = 1;
    } else {
      dims[0] = W; dims[1] = C; dims[2] = H;
      strides[0] = (cuuint64_t)W * H * ELT; strides[1] = (cuuint64_t)W * ELT;
      box[0] = bw; box[1] = C; box[2] = bh;
    }
    CUresult r = enc(&map, dt, 3, d, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("%s order %s: encode -> %d\n", name, order ? "(W,C,H)" : "(W,H,C)", (int)r);
    if (r != CUDA_SUCCESS) continue;
    const int xs[] = {0, 8, 16, -8, -16, W - 24, W - 8, 32, 64, 128};
    for (int xi = 0; xi < 10; ++xi) {
      const int x = xs[xi], y = 5;
      const int elems = order == 0 ? bw * bh : bw * C * bh;
      CK(cudaMemset(dout, 0xEE, bw * bh * 3 * ELT));
      if (order == 0) k_load<ELT><<<1, 128, 32 * 1024>>>(map, x, y, 1, elems, dout);
      else k_load<ELT><<<1, 128, 32 * 1024>>>(map, x, 0, y, elems, dout);
      cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("  x=%d: FAILED: %s\n", x, cudaGetErrorString(e));
        return 2;  // (the context is dead after an illegal instruction)
      }
      std::vector<unsigned char> o(elems * ELT);
      CK(cudaMemcpy(o.data(), dout, o.size(), cudaMemcpyDeviceToHost));
      size_t bad = 0;
      for (int row = 0; row < (order == 0 ? bh : bh * C); ++row) {
        int yy, cc;
        if (order == 0) { yy = y + row; cc = 1; } else { yy = y + row / C; cc = row % C; }
        for (int col = 0; col < bw; ++col)
          for (int b = 0; b < ELT; ++b) {
            const int sx = x + col;
            unsigned char want = 0;
            if (sx >= 0 && sx < W) want = h[(((size_t)cc * H + yy) * W + sx) * ELT + b];
            bad += o[((size_t)row * bw + col) * ELT + b] != want;
          }
      }
      printf("  x=%d: ok, mismatching bytes %zu   got %02x %02x %02x %02x .. want(row0) %02x %02x %02x %02x\n", x, bad, o[0], o[1], o[2], o[3],
             (x >= 0 ? h[(((size_t)1 * H + y) * W + x) * ELT + 0] : 0), (x >= 0 ? h[(((size_t)1 * H + y) * W + x) * ELT + 1] : 0),
             (x >= 0 ? h[(((size_t)1 * H + y) * W + x) * ELT + 2] : 0), (x >= 0 ? h[(((size_t)1 * H + y) * W + x) * ELT + 3] : 0));
    }
  }
  cudaFree(d); cudaFree(dout);
  return 0;
}

int main() {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)f;
  int rc = run<2>(enc, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, "bf16");
  if (rc) return rc;
  rc = run<4>(enc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, "f32");
  if (rc) return rc;
  rc = run<1>(enc, CU_TENSOR_MAP_DATA_TYPE_UINT8, "u8");
  return rc;
}
