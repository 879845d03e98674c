// Probe (round 2): does tex2Dgather work on pitch-linear 2-D textures on B200, and how fast
// is a gather-based bilinear remap compared with the shared-memory tap scheme?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tex_gather_probe tex_gather_probe.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int W = 4096, H = 2048, C = 3;

struct Aff { float a, b, c, d, e, f; };  // sx = a x + b y + c ; sy = d x + e y + f

template <int MODE, typename TEXEL>
__global__ void __launch_bounds__(256) k_remap(const cudaTextureObject_t* texs, const TEXEL* src,
                                               __nv_bfloat16* dst, Aff m, int frames) {
  const int b = blockIdx.z;
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int x0 = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4;
  const cudaTextureObject_t tex = texs[b];
  const TEXEL* s = src + (size_t)b * C * H * W;
  __nv_bfloat16* d = dst + (size_t)b * C * H * W + (size_t)y * W + x0;
  float o[C][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float sx = m.a * (x0 + i) + m.b * y + m.c, sy = m.d * (x0 + i) + m.e * y + m.f;
    sx = sx - floorf(sx * (1.0f / W)) * W;
    sx = fminf(fmaxf(sx, 0.f), W - 1.001f);
    sy = fminf(fmaxf(sy, 0.f), H - 1.001f);
    const float xf = floorf(sx), yf = floorf(sy);
    const float bx = sx - xf, by = sy - yf, ax = 1.f - bx, ay = 1.f - by;
    const float wnw = ax * ay, wne = bx * ay, wsw = ax * by, wse = bx * by;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float vnw, vne, vsw, vse;
      if (MODE == 0) {  // gather
        if (sizeof(TEXEL) == 2) {
          const ushort4 g = tex2Dgather<ushort4>(tex, xf + 1.0f, yf + 1.0f + (float)(c * H), 0);
          vsw = __uint_as_float((unsigned)g.x << 16); vse = __uint_as_float((unsigned)g.y << 16);
          vne = __uint_as_float((unsigned)g.z << 16); vnw = __uint_as_float((unsigned)g.w << 16);
        } else if (sizeof(TEXEL) == 4) {
          const float4 g = tex2Dgather<float4>(tex, xf + 1.0f, yf + 1.0f + (float)(c * H), 0);
          vsw = g.x; vse = g.y; vne = g.z; vnw = g.w;
        } else {
          const uchar4 g = tex2Dgather<uchar4>(tex, xf + 1.0f, yf + 1.0f + (float)(c * H), 0);
          vsw = g.x; vse = g.y; vne = g.z; vnw = g.w;
        }
      } else if (MODE == 1) {  // 4 point fetches
        const float tx = xf + 0.5f, ty = yf + 0.5f + (float)(c * H);
        if (sizeof(TEXEL) == 2) {
          vnw = __uint_as_float((unsigned)tex2D<unsigned short>(tex, tx, ty) << 16);
          vne = __uint_as_float((unsigned)tex2D<unsigned short>(tex, tx + 1, ty) << 16);
          vsw = __uint_as_float((unsigned)tex2D<unsigned short>(tex, tx, ty + 1) << 16);
          vse = __uint_as_float((unsigned)tex2D<unsigned short>(tex, tx + 1, ty + 1) << 16);
        } else if (sizeof(TEXEL) == 4) {
          vnw = tex2D<float>(tex, tx, ty); vne = tex2D<float>(tex, tx + 1, ty);
          vsw = tex2D<float>(tex, tx, ty + 1); vse = tex2D<float>(tex, tx + 1, ty + 1);
        } else {
          vnw = tex2D<unsigned char>(tex, tx, ty); vne = tex2D<unsigned char>(tex, tx + 1, ty);
          vsw = tex2D<unsigned char>(tex, tx, ty + 1); vse = tex2D<unsigned char>(tex, tx + 1, ty + 1);
        }
      } else {  // direct global gather
        const TEXEL* q = s + ((size_t)c * H + (int)yf) * W + (int)xf;
        if (sizeof(TEXEL) == 2) {
          vnw = __uint_as_float((unsigned)__ldg((const unsigned short*)q) << 16);
          vne = __uint_as_float((unsigned)__ldg((const unsigned short*)q + 1) << 16);
          vsw = __uint_as_float((unsigned)__ldg((const unsigned short*)q + W) << 16);
          vse = __uint_as_float((unsigned)__ldg((const unsigned short*)q + W + 1) << 16);
        } else if (sizeof(TEXEL) == 4) {
          vnw = __ldg((const float*)q); vne = __ldg((const float*)q + 1);
          vsw = __ldg((const float*)q + W); vse = __ldg((const float*)q + W + 1);
        } else {
          vnw = __ldg((const unsigned char*)q); vne = __ldg((const unsigned char*)q + 1);
          vsw = __ldg((const unsigned char*)q + W); vse = __ldg((const unsigned char*)q + W + 1);
        }
      }
      o[c][i] = fmaf(vse, wse, fmaf(vsw, wsw, fmaf(vne, wne, vnw * wnw)));
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const __nv_bfloat162 p0 = __floats2bfloat162_rn(o[c][0], o[c][1]);
    const __nv_bfloat162 p1 = __floats2bfloat162_rn(o[c][2], o[c][3]);
    uint2 u;
    u.x = *reinterpret_cast<const unsigned*>(&p0);
    u.y = *reinterpret_cast<const unsigned*>(&p1);
    __stcs(reinterpret_cast<uint2*>(d + (size_t)c * H * W), u);
  }
}

template <typename TEXEL> cudaChannelFormatDesc chan();
template <> cudaChannelFormatDesc chan<unsigned short>() { return cudaCreateChannelDesc<unsigned short>(); }
template <> cudaChannelFormatDesc chan<float>() { return cudaCreateChannelDesc<float>(); }
template <> cudaChannelFormatDesc chan<unsigned char>() { return cudaCreateChannelDesc<unsigned char>(); }

template <typename TEXEL> void run(const char* name, int frames) {
  const size_t n = (size_t)frames * C * H * W;
  TEXEL* src; __nv_bfloat16* dst[3];
  CK(cudaMalloc(&src, n * sizeof(TEXEL)));
  for (int i = 0; i < 3; ++i) CK(cudaMalloc(&dst[i], n * 2));
  {  // deterministic fill: value = small integer pattern valid in every format
    std::vector<TEXEL> h((size_t)C * H * W);
    for (size_t i = 0; i < h.size(); ++i) {
      const unsigned v = (unsigned)((i * 2654435761u) >> 20) & 0x7f;
      if (sizeof(TEXEL) == 2) { __nv_bfloat16 t = __float2bfloat16((float)v); h[i] = *(TEXEL*)&t; }
      else if (sizeof(TEXEL) == 4) { float t = (float)v; h[i] = *(TEXEL*)&t; }
      else h[i] = (TEXEL)v;
    }
    for (int b = 0; b < frames; ++b)
      CK(cudaMemcpy(src + (size_t)b * C * H * W, h.data(), h.size() * sizeof(TEXEL), cudaMemcpyHostToDevice));
  }
  std::vector<cudaTextureObject_t> texs(frames);
  for (int b = 0; b < frames; ++b) {
    cudaResourceDesc rd{}; rd.resType = cudaResourceTypePitch2D;
    rd.res.pitch2D.devPtr = src + (size_t)b * C * H * W;
    rd.res.pitch2D.desc = chan<TEXEL>();
    rd.res.pitch2D.width = W; rd.res.pitch2D.height = (size_t)C * H;
    rd.res.pitch2D.pitchInBytes = (size_t)W * sizeof(TEXEL);
    cudaTextureDesc td{}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
    CK(cudaCreateTextureObject(&texs[b], &rd, &td, nullptr));
  }
  cudaTextureObject_t* dtex; CK(cudaMalloc(&dtex, frames * sizeof(cudaTextureObject_t)));
  CK(cudaMemcpy(dtex, texs.data(), frames * sizeof(cudaTextureObject_t), cudaMemcpyHostToDevice));

  const float angs[3] = {0.0f, 0.15f, 0.6f};
  for (int ai = 0; ai < 3; ++ai) {
    const float al = angs[ai];
    Aff m{cosf(al), -sinf(al), 300.3f, sinf(al), cosf(al), 10.7f - 500.f * sinf(al)};
    dim3 grid(W / 128, H / 8, frames), block(256);
    float ms[3];
    for (int mode = 0; mode < 3; ++mode) {
      cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      for (int it = 0; it < 3; ++it) {
        if (mode == 0) k_remap<0, TEXEL><<<grid, block>>>(dtex, src, dst[0], m, frames);
        if (mode == 1) k_remap<1, TEXEL><<<grid, block>>>(dtex, src, dst[1], m, frames);
        if (mode == 2) k_remap<2, TEXEL><<<grid, block>>>(dtex, src, dst[2], m, frames);
      }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s mode %d: launch failed: %s\n", name, mode, cudaGetErrorString(e)); cudaGetLastError(); ms[mode] = -1; continue; }
      CK(cudaEventRecord(e0));
      const int reps = 10;
      for (int it = 0; it < reps; ++it) {
        if (mode == 0) k_remap<0, TEXEL><<<grid, block>>>(dtex, src, dst[0], m, frames);
        if (mode == 1) k_remap<1, TEXEL><<<grid, block>>>(dtex, src, dst[1], m, frames);
        if (mode == 2) k_remap<2, TEXEL><<<grid, block>>>(dtex, src, dst[2], m, frames);
      }
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms[mode], e0, e1)); ms[mode] /= reps;
    }
    // correctness: gather and point fetches against the direct gather
    std::vector<__nv_bfloat16> h0((size_t)C * H * W), h1(h0.size()), h2(h0.size());
    CK(cudaMemcpy(h0.data(), dst[0], h0.size() * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h1.data(), dst[1], h0.size() * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h2.data(), dst[2], h0.size() * 2, cudaMemcpyDeviceToHost));
    size_t bad0 = 0, bad1 = 0;
    for (size_t i = 0; i < h0.size(); ++i) {
      bad0 += *(unsigned short*)&h0[i] != *(unsigned short*)&h2[i];
      bad1 += *(unsigned short*)&h1[i] != *(unsigned short*)&h2[i];
    }
    printf("%s frames=%d angle=%.2f: gather %.3f ms (mismatch %zu)  point4 %.3f ms (mismatch %zu)  ldg %.3f ms   [px=%.1f M]\n",
           name, frames, al, ms[0], bad0, ms[1], bad1, ms[2], (double)frames * H * W / 1e6);
    fflush(stdout);
  }
  for (auto t : texs) cudaDestroyTextureObject(t);
  cudaFree(src); for (int i = 0; i < 3; ++i) cudaFree(dst[i]); cudaFree(dtex);
}

int main() {
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  printf("device %s, texturePitchAlignment %zu, maxTexture2DLinear %d x %d x %d, maxTexture2DGather %d x %d\n", pr.name,
         pr.texturePitchAlignment, pr.maxTexture2DLinear[0], pr.maxTexture2DLinear[1], pr.maxTexture2DLinear[2],
         pr.maxTexture2DGather[0], pr.maxTexture2DGather[1]);
  run<unsigned short>("bf16(u16)", 32);
  run<float>("f32", 16);
  run<unsigned char>("u8", 32);
  return 0;
}
