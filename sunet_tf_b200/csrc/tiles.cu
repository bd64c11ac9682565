// Any-resolution tile pipeline on the device (demo_any_resolution.py:35-52 and :125-139):
// centre the image on a zero canvas of side X = ceil(max(h,w)/k)*k, cut row-major k x k tiles with the given stride,
// and after the per-tile forwards overlap-add them, divide by the cover count, crop and clamp.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sunet_b200.h"
#include "error.h"

namespace sunet {

struct TileGeom {
  int h, w, k, stride, X, n, oy, ox;
};
static int make_geom(int h, int w, int k, int stride, TileGeom* g) {
  if (h <= 0 || w <= 0 || k <= 0 || stride <= 0 || stride > k) return fail(SUNET_E_SHAPE, "tiles: bad geometry h=%d w=%d k=%d stride=%d", h, w, k, stride);
  const int mx = h > w ? h : w;
  g->h = h; g->w = w; g->k = k; g->stride = stride;
  g->X = (mx + k - 1) / k * k;                 // :38
  if ((g->X - k) % stride) return fail(SUNET_E_SHAPE, "tiles: canvas %d is not covered by kernel %d / stride %d", g->X, k, stride);
  g->n = (g->X - k) / stride + 1;              // unfold count per axis
  g->oy = (g->X - h) / 2; g->ox = (g->X - w) / 2;  // :42
  return 0;
}

__global__ void tiles_extract_kernel(const float* __restrict__ img, int C, TileGeom g, float* __restrict__ tiles, int first, int64_t total) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = static_cast<int>(i % g.k), y = static_cast<int>((i / g.k) % g.k);
  const int c = static_cast<int>((i / (static_cast<int64_t>(g.k) * g.k)) % C);
  const int t = first + static_cast<int>(i / (static_cast<int64_t>(g.k) * g.k * C));
  const int Y = (t / g.n) * g.stride + y - g.oy, Xc = (t % g.n) * g.stride + x - g.ox;
  float v = 0.f;
  if (Y >= 0 && Y < g.h && Xc >= 0 && Xc < g.w) v = __ldg(img + (static_cast<int64_t>(c) * g.h + Y) * g.w + Xc);
  tiles[i] = v;
}

// gather form (no atomics): each canvas pixel sums the tiles of [first, first+count) that cover it
__global__ void tiles_fold_kernel(const float* __restrict__ tiles, int C, TileGeom g, int first, int count, float* __restrict__ acc,
                                  int64_t total) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int Xc = static_cast<int>(i % g.X), Y = static_cast<int>((i / g.X) % g.X);
  const int c = static_cast<int>(i / (static_cast<int64_t>(g.X) * g.X));
  const int ti_hi = min(g.n - 1, Y / g.stride), tj_hi = min(g.n - 1, Xc / g.stride);
  const int ti_lo = max(0, (Y - g.k + g.stride) / g.stride), tj_lo = max(0, (Xc - g.k + g.stride) / g.stride);
  float sum = 0.f;
  for (int ti = ti_lo; ti <= ti_hi; ++ti)
    for (int tj = tj_lo; tj <= tj_hi; ++tj) {
      const int t = ti * g.n + tj;
      if (t < first || t >= first + count) continue;
      const int y = Y - ti * g.stride, x = Xc - tj * g.stride;
      sum += __ldg(tiles + ((static_cast<int64_t>(t - first) * C + c) * g.k + y) * g.k + x);
    }
  acc[i] += sum;
}

__global__ void tiles_finish_kernel(const float* __restrict__ acc, int C, TileGeom g, float* __restrict__ out, int64_t total) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = static_cast<int>(i % g.w), y = static_cast<int>((i / g.w) % g.h);
  const int c = static_cast<int>(i / (static_cast<int64_t>(g.w) * g.h));
  const int Y = y + g.oy, Xc = x + g.ox;
  const int cy = min(g.n - 1, Y / g.stride) - max(0, (Y - g.k + g.stride) / g.stride) + 1;
  const int cx = min(g.n - 1, Xc / g.stride) - max(0, (Xc - g.k + g.stride) / g.stride) + 1;
  const float v = acc[(static_cast<int64_t>(c) * g.X + Y) * g.X + Xc] / static_cast<float>(cy * cx);
  out[i] = fminf(fmaxf(v, 0.f), 1.f);
}

}  // namespace sunet

using namespace sunet;
extern "C" {

int sunet_tiles_extract(const float* img, int chans, int h, int w, int kernel, int stride, float* tiles, int first, int count, void* stream) {
  TileGeom g;
  SUNET_TRY(make_geom(h, w, kernel, stride, &g));
  if (first < 0 || count < 0 || first + count > g.n * g.n) return fail(SUNET_E_SHAPE, "tiles: range [%d,%d) outside %d tiles", first, first + count, g.n * g.n);
  const int64_t total = static_cast<int64_t>(count) * chans * kernel * kernel;
  if (total == 0) return 0;
  tiles_extract_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(img, chans, g, tiles, first, total);
  SUNET_CHECK_LAUNCH();
  return 0;
}

int sunet_tiles_fold(const float* tiles, int chans, int h, int w, int kernel, int stride, int first, int count, float* acc, void* stream) {
  TileGeom g;
  SUNET_TRY(make_geom(h, w, kernel, stride, &g));
  if (first < 0 || count < 0 || first + count > g.n * g.n) return fail(SUNET_E_SHAPE, "tiles: range [%d,%d) outside %d tiles", first, first + count, g.n * g.n);
  const int64_t total = static_cast<int64_t>(chans) * g.X * g.X;
  if (count == 0) return 0;
  tiles_fold_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(tiles, chans, g, first, count, acc, total);
  SUNET_CHECK_LAUNCH();
  return 0;
}

int sunet_tiles_finish(const float* acc, int chans, int h, int w, int kernel, int stride, float* out, void* stream) {
  TileGeom g;
  SUNET_TRY(make_geom(h, w, kernel, stride, &g));
  const int64_t total = static_cast<int64_t>(chans) * h * w;
  tiles_finish_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(acc, chans, g, out, total);
  SUNET_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
