// TEST INFRASTRUCTURE — stand-in for cvd/image_ref.h.
#pragma once
#include <iostream>
namespace CVD {
struct ImageRef {
  int x, y;
  ImageRef() : x(0), y(0) {}
  ImageRef(int x_, int y_) : x(x_), y(y_) {}
  bool operator==(const ImageRef& o) const { return x == o.x && y == o.y; }
  bool operator!=(const ImageRef& o) const { return !(*this == o); }
  bool operator<(const ImageRef& o) const { return y < o.y || (y == o.y && x < o.x); }
  ImageRef operator+(const ImageRef& o) const { return ImageRef(x + o.x, y + o.y); }
  ImageRef operator-(const ImageRef& o) const { return ImageRef(x - o.x, y - o.y); }
  ImageRef operator-() const { return ImageRef(-x, -y); }
  ImageRef operator*(int k) const { return ImageRef(x * k, y * k); }
  ImageRef operator/(int k) const { return ImageRef(x / k, y / k); }
  ImageRef& operator+=(const ImageRef& o) { x += o.x; y += o.y; return *this; }
  ImageRef& operator-=(const ImageRef& o) { x -= o.x; y -= o.y; return *this; }
  ImageRef& operator*=(int k) { x *= k; y *= k; return *this; }
  ImageRef& operator/=(int k) { x /= k; y /= k; return *this; }
  int& operator[](int i) { return i == 0 ? x : y; }
  int operator[](int i) const { return i == 0 ? x : y; }
  unsigned int mag_squared() const { return (unsigned int)(x * x + y * y); }
  int area() const { return x * y; }
  // raster scan: advance x, wrap to the next row; false (and reset to 0,0) after the last pixel
  bool next(const ImageRef& max) { if (++x >= max.x) { x = 0; if (++y >= max.y) { y = 0; return false; } } return true; }
  bool next(const ImageRef& min, const ImageRef& max) { if (++x >= max.x) { x = min.x; if (++y >= max.y) { y = min.y; return false; } } return true; }
  void home() { x = 0; y = 0; }
};
inline ImageRef operator*(int k, const ImageRef& r) { return ImageRef(r.x * k, r.y * k); }
inline std::ostream& operator<<(std::ostream& os, const ImageRef& r) { return os << "[" << r.x << " " << r.y << "]"; }
}  // namespace CVD
