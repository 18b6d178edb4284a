// TEST INFRASTRUCTURE — stand-in for cvd/vision.h, restating (from libCVD's documented behaviour, as
// SURVEY.md §8c lists it) the two calls the reference makes: halfSample and transform.
#pragma once
#include <TooN/TooN.h>
#include "utility.h"
namespace CVD {
// out(x, y) = mean of the 2x2 block, integer division (KeyFrame.cc:27, ImageProcess.cc:288)
inline void halfSample(const BasicImage<byte>& in, BasicImage<byte>& out) {
  const ImageRef sz = out.size();
  for (int y = 0; y < sz.y; y++) {
    const byte* r0 = in[2 * y];
    const byte* r1 = in[2 * y + 1];
    byte* o = out[y];
    for (int x = 0; x < sz.x; x++) o[x] = (byte)((r0[2 * x] + r0[2 * x + 1] + r1[2 * x] + r1[2 * x + 1]) / 4);
  }
}
template <class T> inline Image<T> halfSample(const BasicImage<T>& in) { Image<T> out(in.size() / 2); halfSample(in, out); return out; }

// bilinear sample evaluated in double and converted to the pixel type (byte: truncation)
template <class T> inline void sample(const BasicImage<T>& im, double x, double y, T& result) {
  const int lx = (int)x, ly = (int)y;
  x -= lx; y -= ly;
  const T* r0 = im[ly] + lx;
  const T* r1 = im[ly + 1] + lx;
  result = static_cast<T>((1 - y) * ((1 - x) * r0[0] + x * r0[1]) + y * ((1 - x) * r1[0] + x * r1[1]));
}
// out(p) = in(inOrig + M (p - outOrig)); the source position is accumulated across / down; pixels
// whose source lies outside [0, W-1) x [0, H-1) get defaultValue and are counted
template <class T, class M, class V1, class V2> inline int transform(const BasicImage<T>& in, BasicImage<T>& out, const M& m, const V1& inOrig, const V2& outOrig, const T defaultValue = T()) {
  const int w = out.size().x, h = out.size().y, iw = in.size().x, ih = in.size().y;
  const TooN::Vector<2> across = m.T()[0];
  const TooN::Vector<2> down = m.T()[1];
  const TooN::Vector<2> p0 = inOrig - m * outOrig;
  TooN::Vector<2> p = p0;
  const TooN::Vector<2> carriage_return = down - w * across;
  const double x_bound = iw - 1, y_bound = ih - 1;
  int count = 0;
  for (int i = 0; i < h; ++i, p += carriage_return)
    for (int j = 0; j < w; ++j, p += across) {
      if (0 <= p[0] && 0 <= p[1] && p[0] < x_bound && p[1] < y_bound) sample(in, p[0], p[1], out[i][j]);
      else { out[i][j] = defaultValue; ++count; }
    }
  return count;
}
}  // namespace CVD
