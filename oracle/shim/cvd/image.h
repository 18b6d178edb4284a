// TEST INFRASTRUCTURE — stand-in for cvd/image.h: BasicImage (non-owning view) and Image
// (reference-counted buffer: copies are shallow, resize() detaches, as in libCVD).
#pragma once
#include <cstring>
#include <memory>
#include <vector>
#include "byte.h"
#include "image_ref.h"
namespace CVD {
template <class T> class BasicImage {
 public:
  BasicImage() : my_data(nullptr), my_size(0, 0), my_stride(0) {}
  BasicImage(T* d, const ImageRef& sz, int stride = -1) : my_data(d), my_size(sz), my_stride(stride < 0 ? sz.x : stride) {}
  virtual ~BasicImage() {}
  ImageRef size() const { return my_size; }
  int row_stride() const { return my_stride; }
  int totalsize() const { return my_size.x * my_size.y; }
  T* data() { return my_data; }
  const T* data() const { return my_data; }
  T& operator[](const ImageRef& p) { return my_data[p.y * my_stride + p.x]; }
  const T& operator[](const ImageRef& p) const { return my_data[p.y * my_stride + p.x]; }
  T* operator[](int row) { return my_data + row * my_stride; }
  const T* operator[](int row) const { return my_data + row * my_stride; }
  bool in_image(const ImageRef& p) const { return p.x >= 0 && p.y >= 0 && p.x < my_size.x && p.y < my_size.y; }
  bool in_image_with_border(const ImageRef& p, int b) const { return p.x >= b && p.y >= b && p.x < my_size.x - b && p.y < my_size.y - b; }
  void fill(const T& v) { for (int y = 0; y < my_size.y; y++) for (int x = 0; x < my_size.x; x++) my_data[y * my_stride + x] = v; }
  void zero() { for (int y = 0; y < my_size.y; y++) std::memset((void*)(my_data + y * my_stride), 0, sizeof(T) * my_size.x); }
  T* begin() { return my_data; }
  T* end() { return my_data + totalsize(); }
 protected:
  T* my_data;
  ImageRef my_size;
  int my_stride;
};
template <class T> class SubImage : public BasicImage<T> {
 public:
  using BasicImage<T>::BasicImage;
};
template <class T> class Image : public BasicImage<T> {
 public:
  Image() {}
  explicit Image(const ImageRef& sz) { resize(sz); }
  Image(const ImageRef& sz, const T& v) { resize(sz); this->fill(v); }
  void resize(const ImageRef& sz) {
    buf = std::make_shared<std::vector<T>>((size_t)(sz.x > 0 ? sz.x : 0) * (sz.y > 0 ? sz.y : 0));
    this->my_data = buf->data(); this->my_size = sz; this->my_stride = sz.x;
  }
  void resize(const ImageRef& sz, const T& v) { resize(sz); this->fill(v); }
  Image copy_from_me() const { Image r(this->my_size); if (buf) *r.buf = *buf; return r; }
  void make_unique() { if (buf && buf.use_count() > 1) { *this = copy_from_me(); } }
 private:
  std::shared_ptr<std::vector<T>> buf;
};
}  // namespace CVD
