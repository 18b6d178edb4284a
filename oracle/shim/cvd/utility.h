// TEST INFRASTRUCTURE — stand-in for cvd/utility.h: copy(in, out, size, begin, dst).
#pragma once
#include "image.h"
namespace CVD {
template <class S, class D> inline void copy(const BasicImage<S>& in, BasicImage<D>& out, ImageRef size = ImageRef(-1, -1), ImageRef begin = ImageRef(), ImageRef dst = ImageRef()) {
  if (size.x == -1 && size.y == -1) size = in.size();
  for (int y = 0; y < size.y; y++)
    for (int x = 0; x < size.x; x++) out[ImageRef(dst.x + x, dst.y + y)] = (D)in[ImageRef(begin.x + x, begin.y + y)];
}
}  // namespace CVD
