// TEST INFRASTRUCTURE — stand-in for cvd/convolution.h: convolveGaussian(float image, sigma) as the
// oracle restates it (separable FIR of radius ceil(3 sigma), float taps normalised to unit sum,
// replicated borders, rows then columns, float accumulation from the centre outwards).
#pragma once
#include <cmath>
#include <vector>
#include "image.h"
namespace CVD {
inline void convolveGaussian(BasicImage<float>& im, double sigma) {
  const int w = im.size().x, h = im.size().y;
  const int ks = (int)std::ceil(3.0 * sigma);
  std::vector<float> taps(ks + 1);
  float ksum = 0.f;
  for (int i = 1; i <= ks; i++) ksum += (taps[i] = (float)std::exp(-i * i / (2 * sigma * sigma)));
  taps[0] = 1.f;
  ksum = ksum * 2 + taps[0];
  const double factor = 1.0 / ksum;
  for (int i = 0; i <= ks; i++) taps[i] = (float)(taps[i] * factor);
  auto cl = [](int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); };
  std::vector<float> hrow((size_t)w * h);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      float a = im[y][x] * taps[0];
      for (int k = 1; k <= ks; k++) a += (im[y][cl(x - k, w - 1)] + im[y][cl(x + k, w - 1)]) * taps[k];
      hrow[(size_t)y * w + x] = a;
    }
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      float a = hrow[(size_t)y * w + x] * taps[0];
      for (int k = 1; k <= ks; k++) a += (hrow[(size_t)cl(y - k, h - 1) * w + x] + hrow[(size_t)cl(y + k, h - 1) * w + x]) * taps[k];
      im[y][x] = a;
    }
}
}  // namespace CVD
