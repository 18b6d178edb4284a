// TEST INFRASTRUCTURE — stand-in for cvd/vector_image_ref.h: vec(), ir() (truncation), ir_rounded().
#pragma once
#include <TooN/TooN.h>
#include "image_ref.h"
namespace CVD {
inline TooN::Vector<2> vec(const ImageRef& r) { TooN::Vector<2> v; v[0] = r.x; v[1] = r.y; return v; }
template <class V> inline ImageRef ir(const V& v) { return ImageRef((int)v[0], (int)v[1]); }
template <class V> inline ImageRef ir_rounded(const V& v) {
  return ImageRef((int)(v[0] > 0.0 ? v[0] + 0.5 : v[0] - 0.5), (int)(v[1] > 0.0 ? v[1] + 0.5 : v[1] - 0.5));
}
}  // namespace CVD
