// TEST INFRASTRUCTURE — stand-in for cvd/fast_corner.h.  fast_corner_detect_10 by the segment-test
// DEFINITION (>= 10 contiguous of the 16 ring pixels all > p + b or all < p - b, 3-pixel border,
// raster order) — deliberately the brute-force form, independent of the oracle's and the product's
// optimised detectors; fast_nonmax = FAST-9 score by bisection + 8-neighbour non-maximum suppression.
#pragma once
#include <vector>
#include "image.h"
namespace CVD {
namespace fast_detail {
static const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
inline bool is_corner(const BasicImage<byte>& im, int x, int y, int b, int arc) {
  const int p = im[y][x];
  {  // any arc of >= 9 contiguous ring pixels contains at least two of the four compass pixels
    int nb = 0, nd = 0;
    for (int i = 0; i < 16; i += 4) { const int v = im[y + dy[i]][x + dx[i]]; nb += v > p + b; nd += v < p - b; }
    if (nb < 2 && nd < 2) return false;
  }
  int brighter = 0, darker = 0;  // bit i set: ring pixel i is brighter / darker
  for (int i = 0; i < 16; i++) {
    const int v = im[y + dy[i]][x + dx[i]];
    if (v > p + b) brighter |= 1 << i;
    if (v < p - b) darker |= 1 << i;
  }
  for (int s = 0; s < 16; s++) {
    int mask = 0;
    for (int k = 0; k < arc; k++) mask |= 1 << ((s + k) & 15);
    if ((brighter & mask) == mask || (darker & mask) == mask) return true;
  }
  return false;
}
}  // namespace fast_detail
inline void fast_corner_detect_10(const BasicImage<byte>& im, std::vector<ImageRef>& corners, int barrier) {
  for (int y = 3; y < im.size().y - 3; y++)
    for (int x = 3; x < im.size().x - 3; x++)
      if (fast_detail::is_corner(im, x, y, barrier, 10)) corners.push_back(ImageRef(x, y));
}
inline int fast_corner_score_9(const BasicImage<byte>& im, const ImageRef& c, int barrier) {
  int bmin = barrier, bmax = 255, b = (bmax + bmin) / 2;
  for (;;) {
    if (fast_detail::is_corner(im, c.x, c.y, b, 9)) bmin = b; else bmax = b;
    if (bmin == bmax - 1 || bmin == bmax) return bmin;
    b = (bmin + bmax) / 2;
  }
}
inline void fast_nonmax(const BasicImage<byte>& im, const std::vector<ImageRef>& corners, int barrier, std::vector<ImageRef>& max_corners) {
  const int w = im.size().x, h = im.size().y;
  std::vector<int> smap((size_t)w * h, -1), score(corners.size());
  for (size_t i = 0; i < corners.size(); i++) smap[(size_t)corners[i].y * w + corners[i].x] = score[i] = fast_corner_score_9(im, corners[i], barrier);
  for (size_t i = 0; i < corners.size(); i++) {
    bool keep = true;
    for (int oy = -1; oy <= 1 && keep; oy++)
      for (int ox = -1; ox <= 1; ox++) {
        const int x = corners[i].x + ox, y = corners[i].y + oy;
        if ((!ox && !oy) || x < 0 || y < 0 || x >= w || y >= h) continue;
        if (smap[(size_t)y * w + x] > score[i]) { keep = false; break; }
      }
    if (keep) max_corners.push_back(corners[i]);
  }
}
}  // namespace CVD
