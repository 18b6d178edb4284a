// TEST INFRASTRUCTURE — CPU oracle for PTAM path T (per-frame tracker).  NOT part of the product.
// PINNED against the reference's own Tracker.cc / MapMaker.cc / PatchFinder.cc / KeyFrame.cc /
// ImageProcess.cc compiled in place (oracle/_ref, Makefile.ref): Tracker::TrackFrame sequences,
// ReFindInSingleKeyFrame and MakeKeyFrame_Rest bit-identical (tests/test_ref_pin_tracker.py).
// Restated, not pinned (the libraries are absent from this image; there are no upstream tests or
// golden vectors): the arithmetic inside libCVD / TooN themselves — see oracle_math.h and DESIGN.md §2.
// Single-threaded, faithful loop order.
//
// Restates:  KeyFrame::MakeKeyFrame_Lite        src/KeyFrame.cc:18-54
//            CVD::halfSample / fast_corner_detect_10 / transform / sample   (libCVD 20150407)
//            PatchFinder (all five steps)        src/PatchFinder.cc:52-318
//            ImageProcess::ZMSSDAtPoint          src/ImageProcess.cc:130-163
//            TrackerData                         include/Tracker.h:41-146
//            Tracker::TrackMap / SearchForPoints / CalcPoseUpdate / motion model / quality
//                                                src/Tracker.cc:442-698,867-1107
// The exported orc_tracker_* functions have the signatures of ptam_tracker_* in
// include/ptam_b200.h so that tests drive oracle and product with the same code.
#include "oracle_math.h"
#include "../include/ptam_b200.h"
#include <chrono>
#include <cstdio>
#include <string>

namespace orc {

struct IRef { int x, y; };

struct Level {
  int w = 0, h = 0;
  std::vector<uint8_t> im;
  std::vector<IRef> corners;
  std::vector<int> lut;
  std::vector<IRef> max_corners;           // vMaxCorners (KeyFrame.h:62)
  std::vector<IRef> cand_pos;              // vCandidates[i].irLevelPos (KeyFrame.h:36-41,77)
  std::vector<double> cand_score;          // vCandidates[i].dSTScore
  const uint8_t* row(int y) const { return im.data() + (size_t)y * w; }
  bool in_image_with_border(int x, int y, int b) const { return x >= b && y >= b && x < w - b && y < h - b; }
};

struct KeyFrame { Level lev[PTAM_LEVELS]; };

// CVD::halfSample (called KeyFrame.cc:27): truncating mean of each 2x2 block; out = in/2.
static void half_sample(const Level& in, Level& out) {
  out.w = in.w / 2; out.h = in.h / 2;
  out.im.assign((size_t)out.w * out.h, 0);
  for (int y = 0; y < out.h; y++) {
    const uint8_t* t = in.row(2 * y);
    const uint8_t* b = in.row(2 * y + 1);
    uint8_t* o = out.im.data() + (size_t)y * out.w;
    for (int x = 0; x < out.w; x++) o[x] = (uint8_t)((t[2 * x] + t[2 * x + 1] + b[2 * x] + b[2 * x + 1]) / 4);
  }
}

// CVD::fast_corner_detect_10 by definition (called KeyFrame.cc:35-42): corner iff >= 10 contiguous
// ring pixels are all > p+t or all < p-t; 3-pixel border; raster order.
static inline bool has_run10(unsigned m) {
  m |= m << 16;  // circular
  unsigned a = m & (m >> 1);
  a &= a >> 2;          // runs >= 4
  a &= a >> 4;          // runs >= 8
  a &= m >> 8;          // runs >= 9
  a &= m >> 9;          // runs >= 10
  return (a & 0xFFFFu) != 0;
}
static void fast10(const Level& L, std::vector<IRef>& corners, int t) {
  static const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
  static const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
  corners.clear();
  int off[16];
  for (int k = 0; k < 16; k++) off[k] = dy[k] * L.w + dx[k];
  for (int y = 3; y < L.h - 3; y++) {
    const uint8_t* r = L.row(y);
    for (int x = 3; x < L.w - 3; x++) {
      const uint8_t* c = r + x;
      const int cb = *c + t, c_b = *c - t;
      // any 10-arc contains at least one of ring[0], ring[8]
      const int p0 = c[off[0]], p8 = c[off[8]];
      if (!(p0 > cb || p8 > cb || p0 < c_b || p8 < c_b)) continue;
      unsigned mb = 0, md = 0;
      for (int k = 0; k < 16; k++) {
        const int v = c[off[k]];
        mb |= (unsigned)(v > cb) << k;
        md |= (unsigned)(v < c_b) << k;
      }
      if (has_run10(mb) || has_run10(md)) corners.push_back({x, y});
    }
  }
}

// CPU cost split of TrackFrame for the bench's C1 record (BASELINE.md "Configs" row C1): seconds spent in
// 0 image copy + halfSample, 1 FAST-10 + row LUT, 2 small blurry image + rotation estimator, 3 SearchForPoints
// (warp, template, ZMSSD search, sub-pixel), 4 CalcPoseUpdate; 5 = frames.  Per thread.
static thread_local double g_split[6] = {0, 0, 0, 0, 0, 0};
struct SplitTimer {
  int k;
  std::chrono::steady_clock::time_point t0;
  explicit SplitTimer(int k_) : k(k_), t0(std::chrono::steady_clock::now()) {}
  ~SplitTimer() { g_split[k] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

static void make_keyframe_lite(KeyFrame& kf, const uint8_t* im, int w, int h, int stride, bool detect = true) {
  static const int thr[4] = {10, 15, 15, 10};
  {
    SplitTimer tm(0);
    kf.lev[0].w = w; kf.lev[0].h = h;
    kf.lev[0].im.resize((size_t)w * h);
    for (int y = 0; y < h; y++) std::memcpy(kf.lev[0].im.data() + (size_t)y * w, im + (size_t)y * stride, w);
    for (int i = 1; i < PTAM_LEVELS; i++) half_sample(kf.lev[i - 1], kf.lev[i]);
  }
  SplitTimer tm(1);
  for (int i = 0; i < PTAM_LEVELS; i++) {
    Level& lev = kf.lev[i];
    lev.corners.clear();
    lev.lut.clear();
    if (!detect) continue;
    fast10(lev, lev.corners, thr[i]);
    unsigned v = 0;
    for (int y = 0; y < lev.h; y++) {
      while (v < lev.corners.size() && y > lev.corners[v].y) v++;
      lev.lut.push_back((int)v);
    }
  }
}

// ImageProcess::ZMSSDAtPoint, 8x8 (ImageProcess.cc:130-163).
static int zmssd_at_point(const Level& L, IRef ir, const uint8_t* tmpl, int tsum, int tsumsq, int max_ssd) {
  if (!L.in_image_with_border(ir.x, ir.y, 4)) return max_ssd + 1;
  int isq = 0, isum = 0, cross = 0;
  for (int r = 0; r < 8; r++) {
    const uint8_t* ip = L.row(ir.y - 4 + r) + (ir.x - 4);
    const uint8_t* tp = tmpl + 8 * r;
    for (int c = 0; c < 8; c++) {
      int n = ip[c];
      isum += n; isq += n * n; cross += n * tp[c];
    }
  }
  int SA = tsum, SB = isum;
  return ((2 * SA * SB - SA * SA - SB * SB) / 64 + isq + tsumsq - 2 * cross);
}

struct MapPoint {
  double world[3], right[3], down[3];
  int src_kf, src_level;
  IRef center;
  int outliers = 0, inliers = 0;
};

// PatchFinder + TrackerData merged (one Finder per map point, Tracker.h:46-47).
struct TData {
  // PatchFinder state
  uint8_t tmpl[64];
  int tsum = 0, tsumsq = 0;
  bool has_template = false;      // mpLastTemplateMapPoint == &p
  double last_warp[4] = {9999.9, 0, 0, 9999.9};
  bool template_bad = false;
  double warp_inv[4];
  int search_level = 0;
  float jac[36][2];
  double hinv[9];
  double subpix[2], coarse[2], mean_diff;
  // TrackerData
  double v3cam[3], implane[2], v2image[2], derivs[4];
  bool in_image = false, in_pvs = false, searched = false, found = false, did_subpix = false;
  int n_search_level = -1;
  double v2found[2] = {0, 0};
  double sqrt_inv_noise = 1;
  double err[2];
  double J[12];  // 2x6 row-major
  bool unit_projected = false;  // unit entry: the point projected into the frame, so warp_inv is this call's
};

static inline double level_zero_pos(double p, int l) { return (p + 0.5) * (1 << l) - 0.5; }
static inline double level_n_pos(double p, int l) { return (p + 0.5) / (1 << l) - 0.5; }

// ---------------------------------------------------------------------------------------------
// KeyFrame::MakeKeyFrame_Rest (KeyFrame.cc:61-82), SURVEY 8f rank 2: fast_nonmax(im, vCorners, 10,
// vMaxCorners) and the Shi-Tomasi candidates.  (Its last two lines build the relocaliser's
// SmallBlurryImage, which is sbi_make / sbi_make_jacs above.)
// libCVD pieces restated (library arithmetic: not pinned): fast_nonmax = fast_corner_score_9 followed by
// nonmax_suppression.  The score is found by bisection on the threshold b in [barrier, 255): the
// largest b at which the pixel is still a FAST-9 corner (>= 9 contiguous ring pixels all > p + b or
// all < p - b); a corner is suppressed when one of its 8 neighbours is a corner with a strictly
// greater score (equal scores both survive); output keeps raster order.
// ---------------------------------------------------------------------------------------------
static bool is_fast_corner_n(const Level& L, int x, int y, int b, int arc) {
  static const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
  static const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
  const int p = L.row(y)[x];
  for (int start = 0; start < 16; start++) {
    bool allb = true, alld = true;
    for (int k = 0; k < arc && (allb || alld); k++) {
      const int v = L.row(y + dy[(start + k) & 15])[x + dx[(start + k) & 15]];
      if (!(v > p + b)) allb = false;
      if (!(v < p - b)) alld = false;
    }
    if (allb || alld) return true;
  }
  return false;
}
static int fast_corner_score_9(const Level& L, IRef c, int barrier) {
  int bmin = barrier, bmax = 255, b = (bmax + bmin) / 2;
  for (;;) {
    if (is_fast_corner_n(L, c.x, c.y, b, 9)) bmin = b; else bmax = b;
    if (bmin == bmax - 1 || bmin == bmax) return bmin;
    b = (bmin + bmax) / 2;
  }
}
// ImageProcess::ShiTomasiScoreAtPoint (ImageProcess.cc:20-47)
static double shi_tomasi_score(const Level& L, int half, IRef c) {
  double dXX = 0, dYY = 0, dXY = 0;
  for (int y = c.y - half; y <= c.y + half; y++)
    for (int x = c.x - half; x <= c.x + half; x++) {
      const double dx = (double)L.row(y)[x + 1] - (double)L.row(y)[x - 1];
      const double dy = (double)L.row(y + 1)[x] - (double)L.row(y - 1)[x];
      dXX += dx * dx; dYY += dy * dy; dXY += dx * dy;
    }
  const int n = (2 * half + 1) * (2 * half + 1);
  dXX = dXX / (2.0 * n); dYY = dYY / (2.0 * n); dXY = dXY / (2.0 * n);
  return 0.5 * (dXX + dYY - std::sqrt((dXX + dYY) * (dXX + dYY) - 4 * (dXX * dYY - dXY * dXY)));
}
static void make_keyframe_rest(KeyFrame& kf, double min_st_score) {
  for (int l = 0; l < PTAM_LEVELS; l++) {
    Level& L = kf.lev[l];
    const size_t n = L.corners.size();
    std::vector<int> score(n);
    std::vector<int> smap((size_t)L.w * L.h, -1);
    for (size_t i = 0; i < n; i++) {
      score[i] = fast_corner_score_9(L, L.corners[i], 10);
      smap[(size_t)L.corners[i].y * L.w + L.corners[i].x] = score[i];
    }
    L.max_corners.clear(); L.cand_pos.clear(); L.cand_score.clear();
    for (size_t i = 0; i < n; i++) {
      const IRef c = L.corners[i];
      bool keep = true;
      for (int dy = -1; dy <= 1 && keep; dy++)
        for (int dx = -1; dx <= 1; dx++) {
          if (!dx && !dy) continue;
          const int x = c.x + dx, y = c.y + dy;
          if (x < 0 || y < 0 || x >= L.w || y >= L.h) continue;
          if (smap[(size_t)y * L.w + x] > score[i]) { keep = false; break; }
        }
      if (!keep) continue;
      L.max_corners.push_back(c);
      if (!L.in_image_with_border(c.x, c.y, 10)) continue;
      const double st = shi_tomasi_score(L, 3, c);
      if (st > min_st_score) { L.cand_pos.push_back(c); L.cand_score.push_back(st); }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// SmallBlurryImage + rotation estimator: SmallBlurryImage::MakeFromKF
// (ImageProcess.cc:279-304), ImageProcess::MakeJacs (:170-191), IteratePosRelToTarget (:313-412),
// SE3fromSE2 (:421-473), CalcSBIRotation (:482-494); used by Tracker::TrackFrame (Tracker.cc:95-108)
// and PredictPoseWithMotionModel (:1012-1029).
// libCVD pieces restated (library arithmetic: not pinned): halfSample as above; convolveGaussian(float image,
// sigma) as a separable FIR of radius ceil(3 sigma), taps exp(-i^2 / 2 sigma^2) normalised to unit
// sum and rounded to float, replicated borders, float accumulation  centre, then (a + b) * tap
// outwards, rows first then columns; CVD::transform with bilinear sample() evaluated in double and
// rounded to float, positions accumulated across / down, default value -9e20f outside.
// ---------------------------------------------------------------------------------------------
struct SBI {
  int w = 0, h = 0;
  std::vector<float> tmpl;    // mimTemplate
  std::vector<float> jx, jy;  // mimImageJacs (float differences, read as double)
  bool valid = false;
};

static void sbi_gaussian_taps(double sigma, float* taps /*ksize+1*/, int& ksize) {
  ksize = (int)std::ceil(3.0 * sigma);
  float ksum = 0.f;
  for (int i = 1; i <= ksize; i++) ksum += (taps[i] = (float)std::exp(-i * i / (2 * sigma * sigma)));
  taps[0] = 1.f;
  ksum = ksum * 2 + taps[0];
  const double factor = 1.0 / ksum;
  for (int i = 0; i <= ksize; i++) taps[i] = (float)(taps[i] * factor);
}

static void sbi_make(const KeyFrame& kf, double blur, SBI& o) {
  const Level& L3 = kf.lev[3];
  o.w = (L3.w / 2); o.h = (L3.h / 2);
  // mirSize = aLevels[3].im.size() / 2 and halfSample(aLevels[3].im, mimSmall)
  Level small;
  half_sample(L3, small);
  const int n = o.w * o.h;
  unsigned sum = 0;
  for (int i = 0; i < n; i++) sum += small.im[i];
  const float mean = ((float)sum) / n;
  std::vector<float> t(n), hrow(n);
  for (int i = 0; i < n; i++) t[i] = small.im[i] - mean;
  float taps[16];
  int ks;
  sbi_gaussian_taps(blur, taps, ks);
  auto cl = [](int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); };
  for (int y = 0; y < o.h; y++)
    for (int x = 0; x < o.w; x++) {
      float a = t[y * o.w + x] * taps[0];
      for (int k = 1; k <= ks; k++) a += (t[y * o.w + cl(x - k, o.w - 1)] + t[y * o.w + cl(x + k, o.w - 1)]) * taps[k];
      hrow[y * o.w + x] = a;
    }
  o.tmpl.assign(n, 0.f);
  for (int y = 0; y < o.h; y++)
    for (int x = 0; x < o.w; x++) {
      float a = hrow[y * o.w + x] * taps[0];
      for (int k = 1; k <= ks; k++) a += (hrow[cl(y - k, o.h - 1) * o.w + x] + hrow[cl(y + k, o.h - 1) * o.w + x]) * taps[k];
      o.tmpl[y * o.w + x] = a;
    }
  o.jx.clear(); o.jy.clear();
  o.valid = true;
}

static void sbi_make_jacs(SBI& s) {  // ImageProcess::MakeJacs
  const int n = s.w * s.h;
  s.jx.assign(n, 0.f); s.jy.assign(n, 0.f);
  for (int y = 1; y < s.h - 1; y++)
    for (int x = 1; x < s.w - 1; x++) {
      s.jx[y * s.w + x] = s.tmpl[y * s.w + x + 1] - s.tmpl[y * s.w + x - 1];
      s.jy[y * s.w + x] = s.tmpl[(y + 1) * s.w + x] - s.tmpl[(y - 1) * s.w + x];
    }
}

struct SE2 { double c = 1, s = 0, t[2] = {0, 0}; };  // rotation [[c,-s],[s,c]] + translation
static SE2 se2_mul(const SE2& a, const SE2& b) {
  SE2 r;
  r.c = a.c * b.c - a.s * b.s; r.s = a.s * b.c + a.c * b.s;
  r.t[0] = a.t[0] + (a.c * b.t[0] - a.s * b.t[1]);
  r.t[1] = a.t[1] + (a.s * b.t[0] + a.c * b.t[1]);
  return r;
}

// IteratePosRelToTarget: ESM alignment of `cur` against `other` (which has its jacs made)
static SE2 sbi_iterate(const SBI& cur, const SBI& other, int n_its, double& final_score) {
  const int w = cur.w, h = cur.h;
  const int cx = w / 2, cy = h / 2;
  SE2 c2c;
  std::vector<float> warped(w * h);
  double mean_offset = 0.0;
  final_score = 0.0;
  for (int it = 0; it < n_its; it++) {
    final_score = 0.0;
    double acc[4] = {0, 0, 0, 0}, tri[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    // se2XForm = se2WfromC * se2CtoC * se2WfromC.inverse()
    SE2 wfc; wfc.t[0] = cx; wfc.t[1] = cy;
    SE2 wfc_inv; wfc_inv.t[0] = -(double)cx; wfc_inv.t[1] = -(double)cy;
    const SE2 xf = se2_mul(se2_mul(wfc, c2c), wfc_inv);
    // CVD::transform(mimTemplate, imWarped, R, t, zero, -9e20f)
    {
      const double across[2] = {xf.c, xf.s}, down[2] = {-xf.s, xf.c};
      double pp[2] = {xf.t[0], xf.t[1]};
      const double cr[2] = {down[0] - w * across[0], down[1] - w * across[1]};
      const double xb = w - 1, yb = h - 1;
      for (int i = 0; i < h; i++, pp[0] += cr[0], pp[1] += cr[1])
        for (int j = 0; j < w; j++, pp[0] += across[0], pp[1] += across[1]) {
          if (0 <= pp[0] && 0 <= pp[1] && pp[0] < xb && pp[1] < yb) {
            double x = pp[0], y = pp[1];
            const int lx = (int)x, ly = (int)y;
            x -= lx; y -= ly;
            const float* r0 = &cur.tmpl[ly * w + lx];
            const float* r1 = r0 + w;
            warped[i * w + j] = (float)((1 - y) * ((1 - x) * r0[0] + x * r0[1]) + y * ((1 - x) * r1[0] + x * r1[1]));
          } else warped[i * w + j] = -9e20f;
        }
    }
    for (int y = 1; y < h - 1; y++)
      for (int x = 1; x < w - 1; x++) {
        const float l = warped[y * w + x - 1], r = warped[y * w + x + 1], u = warped[(y - 1) * w + x], dn = warped[(y + 1) * w + x];
        const float here = warped[y * w + x];
        if (l + r + u + dn + here < -9999.9) continue;
        const double g0 = r - l, g1 = dn - u;
        const double sg0 = 0.25 * (g0 + (double)other.jx[y * w + x]), sg1 = 0.25 * (g1 + (double)other.jy[y * w + x]);
        double J[4];
        J[0] = sg0; J[1] = sg1; J[2] = -(y - cy) * sg0 + (x - cx) * sg1; J[3] = 1.0;
        const double diff = here - other.tmpl[y * w + x] + mean_offset;
        final_score += diff * diff;
        for (int q = 0; q < 4; q++) acc[q] += diff * J[q];
        tri[0] += J[0] * J[0]; tri[1] += J[1] * J[0]; tri[2] += J[1] * J[1];
        tri[3] += J[2] * J[0]; tri[4] += J[2] * J[1]; tri[5] += J[2] * J[2];
        tri[6] += J[0]; tri[7] += J[1]; tri[8] += J[2]; tri[9] += 1.0;
      }
    double M[16], upd[4];
    int v = 0;
    for (int j = 0; j < 4; j++)
      for (int i = 0; i <= j; i++) M[4 * j + i] = M[4 * i + j] = tri[v++];
    ldlt_factor(M, 4, 4);
    ldlt_backsub(M, 4, 4, acc, upd);
    SE2 u;
    u.t[0] = -upd[0]; u.t[1] = -upd[1];
    u.c = std::cos(-upd[2]); u.s = std::sin(-upd[2]);
    c2c = se2_mul(c2c, u);
    mean_offset -= upd[3];
  }
  return c2c;
}

// SE3fromSE2: the camera rotation that produces the image-plane rotation (returns so3 as a matrix)
static void sbi_so3_from_se2(const SE2& se2, const Camera& cam_small, int w, int h, double* R /*9*/) {
  const double cx = w / 2, cy = h / 2;
  double turned[2][2], orig[2][3];
  const double off[2] = {5.0, -5.0};
  for (int i = 0; i < 2; i++) {
    turned[i][0] = cx + (se2.c * off[i] + se2.t[0]);
    turned[i][1] = cy + (se2.s * off[i] + se2.t[1]);
    const double im[2] = {cx + off[i], cy};
    double c2[2];
    cam_small.unproject(im, c2);
    orig[i][0] = c2[0]; orig[i][1] = c2[1]; orig[i][2] = 1.0;
  }
  for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  for (int it = 0; it < 3; it++) {
    double C[9] = {10, 0, 0, 0, 10, 0, 0, 0, 10}, b[3] = {0, 0, 0};  // WLS<3>, add_prior(10)
    for (int i = 0; i < 2; i++) {
      double vc[3];
      for (int r = 0; r < 3; r++) vc[r] = R[3 * r] * orig[i][0] + R[3 * r + 1] * orig[i][1] + R[3 * r + 2] * orig[i][2];
      const double ip[2] = {vc[0] / vc[2], vc[1] / vc[2]};
      const Camera::Proj q = cam_small.project(ip);
      const double err[2] = {turned[i][0] - q.im[0], turned[i][1] - q.im[1]};
      double dv[4];
      cam_small.derivs(q, dv);
      const double ooz = 1.0 / vc[2];
      // SO3 generator_field(m, v3Cam): m=0: (0,-z,y), 1: (z,0,-x), 2: (-y,x,0)
      const double gen[3][3] = {{0, -vc[2], vc[1]}, {vc[2], 0, -vc[0]}, {-vc[1], vc[0], 0}};
      double Jr[2][3];
      for (int m = 0; m < 3; m++) {
        const double a0 = (gen[m][0] - vc[0] * gen[m][2] * ooz) * ooz, a1 = (gen[m][1] - vc[1] * gen[m][2] * ooz) * ooz;
        Jr[0][m] = dv[0] * a0 + dv[1] * a1;
        Jr[1][m] = dv[2] * a0 + dv[3] * a1;
      }
      for (int r = 0; r < 2; r++)
        for (int a = 0; a < 3; a++) {
          for (int c = 0; c < 3; c++) C[3 * a + c] += Jr[r][a] * Jr[r][c];
          b[a] += err[r] * Jr[r][a];
        }
    }
    double x[3];
    ldlt_factor(C, 3, 3);
    ldlt_backsub(C, 3, 3, b, x);
    double E[9], N[9];
    so3_exp(x, E);
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) N[3 * r + c] = E[3 * r] * R[c] + E[3 * r + 1] * R[3 + c] + E[3 * r + 2] * R[6 + c];
    for (int i = 0; i < 9; i++) R[i] = N[i];
  }
}

struct Tracker {
  Camera cam;
  Camera cam_small;  // ATANCamera at the SmallBlurryImage size (SE3fromSE2, ImageProcess.cc:423)
  int W, H, S;
  ptam_tracker_params prm;
  std::vector<KeyFrame> store;
  // relocaliser (Relocaliser.cc): pose and SmallBlurryImage (blur 2.5, KeyFrame.cc:80-81) of every stored keyframe
  std::vector<SE3> kf_pose;
  std::vector<char> kf_has_pose;
  std::vector<SBI> kf_sbi;
  bool relocaliser_on() const {
    if (store.empty()) return false;
    for (char c : kf_has_pose) if (!c) return false;
    return true;
  }
  struct Stream {
    std::vector<MapPoint> pts;
    std::vector<TData> td;
    KeyFrame cur;
    ptam_tracker_state st;
    SE3 pose, start;
    int attempted[4], found[4];
    bool did_coarse = false;
    std::vector<int> iter_set;
    std::vector<int> unit_set;   // the points the last patch_search searched (unit entry)
    ptam_track_result res;
    SBI sbi_this, sbi_last;      // mpSBIThisFrame / mpSBILastFrame (Tracker.cc:97-108)
    double sbi_rot[3] = {0, 0, 0}, sbi_score = 0;
  };
  std::vector<Stream> streams;
  std::string err;

  // TrackerData::Project (Tracker.h:70-86)
  void project(TData& d, const MapPoint& p, const SE3& pose, Camera::Proj& q) {
    d.in_image = false;
    pose.apply(p.world, d.v3cam);
    if (d.v3cam[2] < 0.001) return;
    d.implane[0] = d.v3cam[0] / d.v3cam[2];
    d.implane[1] = d.v3cam[1] / d.v3cam[2];
    if (d.implane[0] * d.implane[0] + d.implane[1] * d.implane[1] > cam.largest_radius * cam.largest_radius) return;
    q = cam.project(d.implane);
    d.v2image[0] = q.im[0]; d.v2image[1] = q.im[1];
    if (q.invalid) return;
    if (d.v2image[0] < 0 || d.v2image[1] < 0 || d.v2image[0] > W || d.v2image[1] > H) return;
    d.in_image = true;
  }
  // TrackerData::ProjectAndDerivs (Tracker.h:89-94).  NB: the reference calls
  // Cam.GetProjectionDerivs() whenever bFound, even if Project bailed out early; the camera then
  // still holds the state of its most recent Project() call.  `last` carries that state.
  void project_and_derivs(TData& d, const MapPoint& p, const SE3& pose, Camera::Proj& last) {
    project(d, p, pose, last);
    if (d.found) cam.derivs(last, d.derivs);
  }
  // PatchFinder::CalcSearchLevelAndWarpMatrix (PatchFinder.cc:52-84)
  int calc_search_level(TData& d, const MapPoint& p, const SE3& pose) {
    double v3cam[3], mr[3], md[3];
    pose.apply(p.world, v3cam);
    const double ooz = 1.0 / v3cam[2];
    pose.rotate(p.right, mr);
    pose.rotate(p.down, md);
    double a0 = (mr[0] - v3cam[0] * mr[2] * ooz), a1 = (mr[1] - v3cam[1] * mr[2] * ooz);
    double c0 = (d.derivs[0] * a0 + d.derivs[1] * a1) * ooz, c1 = (d.derivs[2] * a0 + d.derivs[3] * a1) * ooz;
    d.warp_inv[0] = c0; d.warp_inv[2] = c1;  // column 0
    a0 = (md[0] - v3cam[0] * md[2] * ooz); a1 = (md[1] - v3cam[1] * md[2] * ooz);
    c0 = (d.derivs[0] * a0 + d.derivs[1] * a1) * ooz; c1 = (d.derivs[2] * a0 + d.derivs[3] * a1) * ooz;
    d.warp_inv[1] = c0; d.warp_inv[3] = c1;  // column 1
    double det = d.warp_inv[0] * d.warp_inv[3] - d.warp_inv[1] * d.warp_inv[2];
    d.search_level = 0;
    while (det > 3 && d.search_level < PTAM_LEVELS - 1) { d.search_level++; det *= 0.25; }
    if (det > 3 || det < 0.25) { d.template_bad = true; return -1; }
    return d.search_level;
  }
  // PatchFinder::MakeTemplateCoarseCont (PatchFinder.cc:98-127) + CVD::transform/sample.
  void make_template(TData& d, const MapPoint& p) {
    const double* m = d.warp_inv;
    const double det = m[0] * m[3] - m[2] * m[1];
    const double idet = 1.0 / det;
    const int sc = 1 << d.search_level;
    double m2[4];  // M2Inverse (Tools.h:54-65) * LevelScale
    m2[0] = (m[3] * idet) * sc; m2[3] = (m[0] * idet) * sc;
    m2[2] = (-m[2] * idet) * sc; m2[1] = (-m[1] * idet) * sc;
    bool refresh = !d.has_template;
    for (int i = 0; !refresh && i < 2; i++) {
      double d0 = m2[i] - d.last_warp[i], d1 = m2[2 + i] - d.last_warp[2 + i];
      if (d0 * d0 + d1 * d1 > 0.07 * 0.07) refresh = true;
    }
    if (!refresh) return;
    const Level& src = store[p.src_kf].lev[p.src_level];
    // CVD::transform(in, out, M, inOrig, outOrig): p = inOrig + M (out - outOrig), accumulated
    // incrementally across/down exactly like libCVD (vision.h).
    const double across[2] = {m2[0], m2[2]}, downv[2] = {m2[1], m2[3]};
    double pp[2] = {p.center.x - (m2[0] * 4.0 + m2[1] * 4.0), p.center.y - (m2[2] * 4.0 + m2[3] * 4.0)};
    const double cr[2] = {downv[0] - 8 * across[0], downv[1] - 8 * across[1]};
    const double xb = src.w - 1, yb = src.h - 1;
    int outside = 0;
    for (int i = 0; i < 8; i++, pp[0] += cr[0], pp[1] += cr[1])
      for (int j = 0; j < 8; j++, pp[0] += across[0], pp[1] += across[1]) {
        if (0 <= pp[0] && 0 <= pp[1] && pp[0] < xb && pp[1] < yb) {
          double x = pp[0], y = pp[1];
          const int lx = (int)x, ly = (int)y;
          x -= lx; y -= ly;
          const uint8_t* r0 = src.row(ly) + lx;
          const uint8_t* r1 = src.row(ly + 1) + lx;
          double v = (1 - y) * ((1 - x) * r0[0] + x * r0[1]) + y * ((1 - x) * r1[0] + x * r1[1]);
          d.tmpl[8 * i + j] = (uint8_t)v;
        } else {
          d.tmpl[8 * i + j] = 0;
          ++outside;
        }
      }
    d.template_bad = outside != 0;
    int s = 0, ss = 0;
    for (int k = 0; k < 64; k++) { int b = d.tmpl[k]; s += b; ss += b * b; }
    d.tsum = s; d.tsumsq = ss;
    d.has_template = true;
    for (int k = 0; k < 4; k++) d.last_warp[k] = m2[k];
  }
  // PatchFinder::FindPatchCoarse (PatchFinder.cc:160-211)
  bool find_patch_coarse(TData& d, IRef pos, const KeyFrame& kf, unsigned range) {
    const int max_ssd = 8 * 8 * 500;
    const int sc = 1 << d.search_level;
    pos.x /= sc; pos.y /= sc;
    range = (range + sc - 1) / sc;
    int top = pos.y - (int)range, bot1 = pos.y + (int)range + 1;
    const int left = pos.x - (int)range, right = pos.x + (int)range;
    const Level& L = kf.lev[d.search_level];
    if (top < 0) top = 0;
    if (top >= L.h) return false;
    if (bot1 <= 0) return false;
    IRef best{0, 0};
    int best_ssd = max_ssd + 1;
    size_t i = L.lut[top];
    size_t i_end = bot1 >= L.h ? L.corners.size() : (size_t)L.lut[bot1];
    for (; i < i_end; i++) {
      const IRef c = L.corners[i];
      if (c.x < left || c.x > right) continue;
      const int ddx = pos.x - c.x, ddy = pos.y - c.y;
      if ((unsigned)(ddx * ddx + ddy * ddy) > range * range) continue;
      int ssd = zmssd_at_point(L, c, d.tmpl, d.tsum, d.tsumsq, max_ssd);
      if (ssd < best_ssd) { best = c; best_ssd = ssd; }
    }
    if (best_ssd < max_ssd) {
      d.coarse[0] = level_zero_pos((double)best.x, d.search_level);
      d.coarse[1] = level_zero_pos((double)best.y, d.search_level);
      return true;
    }
    return false;
  }
  // PatchFinder::MakeSubPixTemplate (PatchFinder.cc:219-240)
  void make_subpix_template(TData& d) {
    double H[9] = {0};
    for (int x = 1; x < 7; x++)
      for (int y = 1; y < 7; y++) {
        double g[3];
        g[0] = 0.5 * (d.tmpl[8 * y + x + 1] - d.tmpl[8 * y + x - 1]);
        g[1] = 0.5 * (d.tmpl[8 * (y + 1) + x] - d.tmpl[8 * (y - 1) + x]);
        g[2] = 1.0;
        d.jac[(y - 1) * 6 + (x - 1)][0] = (float)g[0];
        d.jac[(y - 1) * 6 + (x - 1)][1] = (float)g[1];
        for (int r = 0; r < 3; r++)
          for (int c = 0; c < 3; c++) H[3 * r + c] += g[r] * g[c];
      }
    ldlt_inverse(H, 3, d.hinv);
    d.subpix[0] = d.coarse[0]; d.subpix[1] = d.coarse[1];
    d.mean_diff = 0.0;
  }
  // PatchFinder::IterateSubPix (PatchFinder.cc:275-318)
  double iterate_subpix(TData& d, const KeyFrame& kf) {
    const int l = d.search_level;
    const double cx = level_n_pos(d.subpix[0], l), cy = level_n_pos(d.subpix[1], l);
    const Level& L = kf.lev[l];
    const int rx = (int)(cx > 0.0 ? cx + 0.5 : cx - 0.5), ry = (int)(cy > 0.0 ? cy + 0.5 : cy - 0.5);
    if (!L.in_image_with_border(rx, ry, 5)) return -1.0;
    const double bx = cx - 4, by = cy - 4;
    const double dX = bx - std::floor(bx), dY = by - std::floor(by);
    const float fTL = (float)((1.0 - dX) * (1.0 - dY));
    const float fTR = (float)((dX) * (1.0 - dY));
    const float fBL = (float)((1.0 - dX) * (dY));
    const float fBR = (float)((dX) * (dY));
    double acc[3] = {0, 0, 0};
    const int ibx = (int)bx, iby = (int)by;
    for (int y = 1; y < 7; y++) {
      const uint8_t* tl = L.row(iby + y) + ibx + 1;
      for (int x = 1; x < 7; x++) {
        float fp = fTL * tl[0] + fTR * tl[1] + fBL * tl[L.w] + fBR * tl[L.w + 1];
        tl++;
        double diff = fp - d.tmpl[8 * y + x] + d.mean_diff;
        acc[0] += diff * d.jac[(y - 1) * 6 + (x - 1)][0];
        acc[1] += diff * d.jac[(y - 1) * 6 + (x - 1)][1];
        acc[2] += diff;
      }
    }
    double u[3];
    for (int r = 0; r < 3; r++) u[r] = d.hinv[3 * r] * acc[0] + d.hinv[3 * r + 1] * acc[1] + d.hinv[3 * r + 2] * acc[2];
    d.subpix[0] -= u[0] * (1 << l);
    d.subpix[1] -= u[1] * (1 << l);
    d.mean_diff -= u[2];
    return u[0] * u[0] + u[1] * u[1];
  }
  bool iterate_subpix_to_convergence(TData& d, const KeyFrame& kf, int max_its) {
    for (int it = 0; it < max_its; it++) {
      double u2 = iterate_subpix(d, kf);
      if (u2 < 0) return false;
      if (u2 < 0.03 * 0.03) return true;
    }
    return false;
  }
  // Tracker::SearchForPoints (Tracker.cc:867-912)
  unsigned search_for_points(Stream& s, const std::vector<int>& v, unsigned range, int subpix_its) {
    SplitTimer tm(3);
    unsigned nfound = 0;
    for (int idx : v) {
      TData& d = s.td[idx];
      make_template(d, s.pts[idx]);
      if (d.template_bad) { d.in_image = d.found = false; continue; }
      s.attempted[d.search_level]++;
      bool f = find_patch_coarse(d, IRef{(int)d.v2image[0], (int)d.v2image[1]}, s.cur, range);
      d.searched = true;
      if (!f) { d.found = false; continue; }
      d.found = true;
      d.sqrt_inv_noise = 1.0 / (1 << d.search_level);
      nfound++;
      s.found[d.search_level]++;
      if (subpix_its > 0) {
        d.did_subpix = true;
        make_subpix_template(d);
        if (!iterate_subpix_to_convergence(d, s.cur, subpix_its)) {
          d.found = false; nfound--; s.found[d.search_level]--;
          continue;
        }
        d.v2found[0] = d.subpix[0]; d.v2found[1] = d.subpix[1];
      } else {
        d.v2found[0] = d.coarse[0]; d.v2found[1] = d.coarse[1];
        d.did_subpix = false;
      }
    }
    return nfound;
  }
  // TrackerData::CalcJacobian (Tracker.h:125-136)
  void calc_jacobian(TData& d) {
    const double ooz = 1.0 / d.v3cam[2];
    const double X = d.v3cam[0], Y = d.v3cam[1], Z = d.v3cam[2];
    // SE3<>::generator_field(m, (X,Y,Z,1)): translations e_m; rotations (0,-Z,Y),(Z,0,-X),(-Y,X,0)
    const double g[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, -Z, Y}, {Z, 0, -X}, {-Y, X, 0}};
    for (int m = 0; m < 6; m++) {
      double a0 = (g[m][0] - X * g[m][2] * ooz) * ooz;
      double a1 = (g[m][1] - Y * g[m][2] * ooz) * ooz;
      d.J[m] = d.derivs[0] * a0 + d.derivs[1] * a1;
      d.J[6 + m] = d.derivs[2] * a0 + d.derivs[3] * a1;
    }
  }
  // Tracker::CalcPoseUpdate (Tracker.cc:928-1005) with TooN WLS<6>.
  void calc_pose_update(Stream& s, const std::vector<int>& v, double override_sigma, bool mark, double* mu) {
    SplitTimer tm(4);
    const int est = prm.mestimator;
    std::vector<double> e2;
    for (int idx : v) {
      TData& d = s.td[idx];
      if (!d.found) continue;
      d.err[0] = d.sqrt_inv_noise * (d.v2found[0] - d.v2image[0]);
      d.err[1] = d.sqrt_inv_noise * (d.v2found[1] - d.v2image[1]);
      e2.push_back(d.err[0] * d.err[0] + d.err[1] * d.err[1]);
    }
    for (int i = 0; i < 6; i++) mu[i] = 0;
    if (e2.empty()) return;
    double sigma2 = override_sigma > 0 ? override_sigma : mest_find_sigma_squared(e2, est);
    double C[36] = {0}, b[6] = {0};
    for (int i = 0; i < 6; i++) C[7 * i] += 100.0;  // add_prior
    for (int idx : v) {
      TData& d = s.td[idx];
      if (!d.found) continue;
      const double es = d.err[0] * d.err[0] + d.err[1] * d.err[1];
      const double wgt = mest_weight(es, sigma2, est);
      if (wgt == 0.0) { if (mark) s.pts[idx].outliers++; continue; }
      else if (mark) s.pts[idx].inliers++;
      for (int r = 0; r < 2; r++) {  // add_mJ(m, J, w): C += (J w) J^T ; b += m (J w)
        double Jr[6], Jw[6];
        for (int k = 0; k < 6; k++) { Jr[k] = d.sqrt_inv_noise * d.J[6 * r + k]; Jw[k] = Jr[k] * wgt; }
        for (int i = 0; i < 6; i++)
          for (int j = 0; j < 6; j++) C[6 * i + j] += Jw[i] * Jr[j];
        for (int i = 0; i < 6; i++) b[i] += d.err[r] * Jw[i];
      }
    }
    ldlt_factor(C, 6, 6);
    ldlt_backsub(C, 6, 6, b, mu);
  }

  // Tracker::TrackMap (Tracker.cc:442-698); std::random_shuffle == identity permutation.
  void track_map(Stream& s) {
    for (int i = 0; i < 4; i++) s.attempted[i] = s.found[i] = 0;
    std::vector<int> pvs[4];
    Camera::Proj last{};
    for (size_t i = 0; i < s.pts.size(); i++) {
      TData& d = s.td[i];
      d.in_pvs = false;
      d.n_search_level = -1;
      project(d, s.pts[i], s.pose, last);
      if (!d.in_image) continue;
      cam.derivs(last, d.derivs);
      d.n_search_level = calc_search_level(d, s.pts[i], s.pose);
      if (d.n_search_level == -1) continue;
      d.searched = false; d.found = false;
      d.in_pvs = true;
      pvs[d.n_search_level].push_back((int)i);
    }
    for (int l = 0; l < 4; l++) s.res.n_pvs[l] = (int)pvs[l].size();
    std::vector<int> next, iter;
    unsigned coarse_max = prm.coarse_max, coarse_range = prm.coarse_range;
    s.did_coarse = false;
    bool try_coarse = true;
    if (prm.disable_coarse || s.st.msd_scaled_velocity_magnitude < prm.coarse_min_velocity || coarse_max == 0) try_coarse = false;
    if (s.st.just_recovered_so_use_coarse) {
      try_coarse = true; coarse_max *= 2; coarse_range *= 2; s.st.just_recovered_so_use_coarse = 0;
    }
    s.res.n_coarse = 0;
    if (try_coarse && pvs[3].size() + pvs[2].size() > (unsigned)prm.coarse_min) {
      if (pvs[3].size() <= coarse_max) { next = pvs[3]; pvs[3].clear(); }
      else {
        for (unsigned i = 0; i < coarse_max; i++) next.push_back(pvs[3][i]);
        pvs[3].erase(pvs[3].begin(), pvs[3].begin() + coarse_max);
      }
      if (next.size() < coarse_max) {
        unsigned more = coarse_max - (unsigned)next.size();
        if (pvs[2].size() <= more) { next = pvs[2]; pvs[2].clear(); }  // sic: assignment, as in Tracker.cc:533
        else {
          for (unsigned i = 0; i < more; i++) next.push_back(pvs[2][i]);
          pvs[2].erase(pvs[2].begin(), pvs[2].begin() + more);
        }
      }
      s.res.n_coarse = (int)next.size();
      unsigned nf = search_for_points(s, next, coarse_range, prm.coarse_subpix_its);
      iter = next;
      if (nf >= (unsigned)prm.coarse_min) {
        s.did_coarse = true;
        for (int it = 0; it < 10; it++) {
          if (it != 0)
            for (int idx : iter) if (s.td[idx].found) project_and_derivs(s.td[idx], s.pts[idx], s.pose, last);
          for (int idx : iter) if (s.td[idx].found) calc_jacobian(s.td[idx]);
          double mu[6];
          calc_pose_update(s, iter, it > 5 ? 1.0 : 0.0, false, mu);
          s.pose = se3_mul(se3_exp(mu), s.pose);
        }
      }
    }
    const unsigned fine_range = s.did_coarse ? 5 : 10;
    {
      for (int idx : pvs[3]) project_and_derivs(s.td[idx], s.pts[idx], s.pose, last);
      search_for_points(s, pvs[3], fine_range, 8);
      for (int idx : pvs[3]) iter.push_back(idx);
      s.res.n_level3 = (int)pvs[3].size();
    }
    {
      next.clear();
      for (int l = 2; l >= 0; l--) for (int idx : pvs[l]) next.push_back(idx);
      int use = prm.max_patches_per_frame - (int)iter.size();
      if (use < 0) use = 0;
      if ((int)next.size() > use) next.resize(use);  // random_shuffle == identity, then chop
      if (s.did_coarse)
        for (int idx : next) project_and_derivs(s.td[idx], s.pts[idx], s.pose, last);
      search_for_points(s, next, fine_range, 0);
      for (int idx : next) iter.push_back(idx);
      s.res.n_fine = (int)next.size();
    }
    double last_update[6] = {0, 0, 0, 0, 0, 0};
    for (int it = 0; it < 10; it++) {
      const bool nonlin = it == 0 || it == 4 || it == 9;
      if (it != 0) {
        if (nonlin) {
          for (int idx : iter) if (s.td[idx].found) project_and_derivs(s.td[idx], s.pts[idx], s.pose, last);
        } else {
          for (int idx : iter) {
            TData& d = s.td[idx];
            if (!d.found) continue;
            for (int r = 0; r < 2; r++) {  // LinearUpdate (Tracker.h:139-142)
              double a = 0;
              for (int k = 0; k < 6; k++) a += d.J[6 * r + k] * last_update[k];
              d.v2image[r] += a;
            }
          }
        }
      }
      if (nonlin) for (int idx : iter) if (s.td[idx].found) calc_jacobian(s.td[idx]);
      double mu[6];
      calc_pose_update(s, iter, it > 5 ? 16.0 : 0.0, it == 9, mu);
      s.pose = se3_mul(se3_exp(mu), s.pose);
      for (int k = 0; k < 6; k++) last_update[k] = mu[k];
    }
    s.iter_set = iter;
    // scene depth (Tracker.cc:680-697)
    double sum = 0, sumsq = 0; int n = 0;
    for (int idx : iter) if (s.td[idx].found) { double z = s.td[idx].v3cam[2]; sum += z; sumsq += z * z; n++; }
    if (n > 20) {
      s.st.scene_depth_mean = sum / n;
      s.st.scene_depth_sigma = std::sqrt((sumsq / n) - s.st.scene_depth_mean * s.st.scene_depth_mean);
    }
  }

  // MapMaker::ReFindInSingleKeyFrame / ReFind_Common (MapMaker.cc:943-1040): every map
  // point of the stream is looked for in the keyframe that is the stream's current frame, whose pose
  // is given.  The map maker's PatchFinder is its own (static, MapMaker.cc:977): consecutive calls are
  // for different points, so the template is always re-made; the return value of
  // CalcSearchLevelAndWarpMatrix is ignored there, only TemplateBad() after MakeTemplateCoarseCont
  // (= source pixels outside the source image) rejects.  A point that is not found is one the
  // reference puts into sNeverRetryKFs.
  void refind(Stream& s, const SE3& pose) {
    for (int i = 0; i < 4; i++) s.attempted[i] = s.found[i] = 0;
    Camera::Proj last{};
    for (size_t i = 0; i < s.pts.size(); i++) {
      TData& d = s.td[i];
      d.in_pvs = false; d.searched = false; d.found = false; d.did_subpix = false;
      d.n_search_level = -1;
      d.has_template = false; d.template_bad = false;
      project(d, s.pts[i], pose, last);
      if (!d.in_image) continue;
      cam.derivs(last, d.derivs);
      calc_search_level(d, s.pts[i], pose);
      d.template_bad = false;
      d.n_search_level = d.search_level;
      d.in_pvs = true;
      make_template(d, s.pts[i]);
      if (d.template_bad) continue;
      s.attempted[d.search_level]++;
      const bool f = find_patch_coarse(d, IRef{(int)d.v2image[0], (int)d.v2image[1]}, s.cur, 4);  // very tight search radius
      d.searched = true;
      if (!f) continue;
      d.found = true;
      s.found[d.search_level]++;
      if (d.search_level > 0) {
        d.did_subpix = true;
        make_subpix_template(d);
        iterate_subpix_to_convergence(d, s.cur, 8);  // result not checked (MapMaker.cc:1005-1007)
        d.v2found[0] = d.subpix[0]; d.v2found[1] = d.subpix[1];
      } else {
        d.v2found[0] = d.coarse[0]; d.v2found[1] = d.coarse[1];
      }
    }
  }

  // Unit entry (ptam_patch_search_batch): PatchFinder steps 1-5 (PatchFinder.h:54-98) for every map point
  // against the stream's current frame, driven exactly as Tracker::SearchForPoints drives them
  // (Tracker.cc:867-912): TrackerData::Project, GetProjectionDerivs, CalcSearchLevelAndWarpMatrix (a negative
  // level rejects the point), MakeTemplateCoarseCont with the per-point template cache, FindPatchCoarse at the
  // given range, and MakeSubPixTemplate + IterateSubPixToConvergence when subpix_its > 0.
  void patch_search(Stream& s, const SE3& pose, unsigned range, int subpix_its) {
    for (int i = 0; i < 4; i++) s.attempted[i] = s.found[i] = 0;
    Camera::Proj last{};
    std::vector<int> v;
    for (size_t i = 0; i < s.pts.size(); i++) {
      TData& d = s.td[i];
      d.in_pvs = false; d.searched = false; d.found = false; d.did_subpix = false;
      d.n_search_level = -1;
      d.unit_projected = false;
      project(d, s.pts[i], pose, last);
      if (!d.in_image) continue;
      d.unit_projected = true;
      cam.derivs(last, d.derivs);
      d.n_search_level = calc_search_level(d, s.pts[i], pose);
      if (d.n_search_level == -1) continue;
      d.in_pvs = true;
      v.push_back((int)i);
    }
    search_for_points(s, v, range, subpix_its);
    s.unit_set = v;
  }
  // Unit entry (ptam_pose_update): CalcJacobian (Tracker.h:125-136) + one Tracker::CalcPoseUpdate
  // (Tracker.cc:928-1005) over the points the last patch_search found.
  int pose_update(Stream& s, double override_sigma, bool mark, double* mu) {
    int nf = 0;
    for (int idx : s.unit_set)
      if (s.td[idx].found) { calc_jacobian(s.td[idx]); nf++; }
    calc_pose_update(s, s.unit_set, override_sigma, mark, mu);
    return nf;
  }

  // MapMaker::AddPointEpipolar (MapMaker.cc:529-688; SURVEY 8f rank 3, second half) up to the sub-pixel
  // position in the target keyframe: epipolar segment of the candidate's view ray in the target's z=1
  // plane, scan of ALL target corners of that level against it, un-warped 8x8 template
  // (MakeTemplateCoarseNoWarp), first minimum ZMSSD <= mnMaxSSD, sub-pixel refinement (10 iterations, must
  // converge).  The target keyframe is the stream's current frame; the source a stored keyframe.
  // The triangulation that follows (4x4 SVD per accepted point, MapMaker.cc:176-187) is left to the caller.
  struct EpiLine { bool ok; double normal[2], along[2], norm_dist, min_len, max_len; };
  EpiLine epipolar_line(int level, IRef cand, const SE3& src, double dmean, double dsigma, const SE3& tgt, double wiggle) const {
    EpiLine e{};
    const double root[2] = {level_zero_pos((double)cand.x, level), level_zero_pos((double)cand.y, level)};
    double uc[2];
    cam.unproject(root, uc);
    double ray[3] = {uc[0], uc[1], 1.0};
    const double nr = std::sqrt(ray[0] * ray[0] + ray[1] * ray[1] + ray[2] * ray[2]);
    for (int i = 0; i < 3; i++) ray[i] = ray[i] / nr;
    double ray_w[3], dirn[3];
    for (int i = 0; i < 3; i++) ray_w[i] = src.R[i] * ray[0] + src.R[3 + i] * ray[1] + src.R[6 + i] * ray[2];  // R^T ray
    tgt.rotate(ray_w, dirn);
    const double d_start = std::max(wiggle, dmean - dsigma), d_end = std::min(40 * wiggle, dmean + dsigma);
    double c_w[3], c_t[3];
    for (int i = 0; i < 3; i++) c_w[i] = -(src.R[i] * src.t[0] + src.R[3 + i] * src.t[1] + src.R[6 + i] * src.t[2]);
    tgt.apply(c_w, c_t);
    double rs[3], re[3];
    for (int i = 0; i < 3; i++) { rs[i] = c_t[i] + d_start * dirn[i]; re[i] = c_t[i] + d_end * dirn[i]; }
    if (re[2] <= rs[2]) return e;
    if (re[2] <= 0.0) return e;
    if (rs[2] <= 0.0) {
      const double k = 0.001 - rs[2] / dirn[2];
      for (int i = 0; i < 3; i++) rs[i] += dirn[i] * k;
    }
    const double A[2] = {rs[0] / rs[2], rs[1] / rs[2]}, B[2] = {re[0] / re[2], re[1] / re[2]};
    double al[2] = {A[0] - B[0], A[1] - B[1]};
    if (al[0] * al[0] + al[1] * al[1] < 1e-8) return e;
    const double na = std::sqrt(al[0] * al[0] + al[1] * al[1]);
    al[0] = al[0] / na; al[1] = al[1] / na;
    e.along[0] = al[0]; e.along[1] = al[1];
    e.normal[0] = al[1]; e.normal[1] = -al[0];
    e.norm_dist = A[0] * e.normal[0] + A[1] * e.normal[1];
    if (std::fabs(e.norm_dist) > cam.largest_radius) return e;
    const double la = al[0] * A[0] + al[1] * A[1], lb = al[0] * B[0] + al[1] * B[1];
    e.min_len = std::min(la, lb) - 0.05;
    e.max_len = std::max(la, lb) + 0.05;
    if (e.min_len < -2.0) e.min_len = -2.0;
    if (e.max_len < -2.0) e.max_len = -2.0;
    if (e.min_len > 2.0) e.min_len = 2.0;
    if (e.max_len > 2.0) e.max_len = 2.0;
    e.ok = true;
    return e;
  }
  void epipolar_search(Stream& s, int level, int src_kf, const SE3& src, double dmean, double dsigma, const SE3& tgt, double wiggle,
                       int n, const int32_t* cand, int32_t* found, int32_t* best_corner, double* sub_pos) {
    const Level& SL = store[src_kf].lev[level];
    const Level& TL = s.cur.lev[level];
    const int max_ssd = 8 * 8 * 500;
    // vImplaneCorners: UnProject of the (truncated) level-zero position of every target corner (MapMaker.cc:608-614)
    std::vector<double> implane(2 * TL.corners.size());
    for (size_t i = 0; i < TL.corners.size(); i++) {
      const double p[2] = {(double)(int)level_zero_pos((double)TL.corners[i].x, level), (double)(int)level_zero_pos((double)TL.corners[i].y, level)};
      cam.unproject(p, &implane[2 * i]);
    }
    const double max_dist = cam.one_pixel_dist * (4.0 + 1.0 * (1 << level));
    const double max_dist_sq = max_dist * max_dist;
    for (int c = 0; c < n; c++) {
      found[c] = 0; best_corner[c] = -1; sub_pos[2 * c] = sub_pos[2 * c + 1] = 0;
      const IRef cp{cand[2 * c], cand[2 * c + 1]};
      const EpiLine e = epipolar_line(level, cp, src, dmean, dsigma, tgt, wiggle);
      if (!e.ok) continue;
      TData d;
      d.search_level = level;
      if (!SL.in_image_with_border(cp.x, cp.y, 5)) continue;  // MakeTemplateCoarseNoWarp: template bad
      int ts = 0, tss = 0;
      for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) { const int b = SL.row(cp.y - 4 + y)[cp.x - 4 + x]; d.tmpl[8 * y + x] = (uint8_t)b; ts += b; tss += b * b; }
      d.tsum = ts; d.tsumsq = tss;
      int best = -1, best_ssd = max_ssd + 1;
      for (size_t i = 0; i < TL.corners.size(); i++) {
        const double* im = &implane[2 * i];
        const double dd = e.norm_dist - (im[0] * e.normal[0] + im[1] * e.normal[1]);
        if (dd * dd > max_dist_sq) continue;
        const double al = im[0] * e.along[0] + im[1] * e.along[1];
        if (al < e.min_len) continue;
        if (al > e.max_len) continue;
        const int ssd = zmssd_at_point(TL, TL.corners[i], d.tmpl, d.tsum, d.tsumsq, max_ssd);
        if (ssd < best_ssd) { best = (int)i; best_ssd = ssd; }
      }
      if (best == -1) continue;
      best_corner[c] = best;
      d.coarse[0] = level_zero_pos((double)TL.corners[best].x, level);
      d.coarse[1] = level_zero_pos((double)TL.corners[best].y, level);
      make_subpix_template(d);
      if (!iterate_subpix_to_convergence(d, s.cur, 10)) continue;
      found[c] = 1;
      sub_pos[2 * c] = d.subpix[0]; sub_pos[2 * c + 1] = d.subpix[1];
    }
  }

  void track_frame(Stream& s, const uint8_t* im, int stride) {
    make_keyframe_lite(s.cur, im, W, H, stride);
    g_split[5] += 1.0;
    // Update the small images for the rotation estimator (Tracker.cc:95-108)
    {
      SplitTimer tm(2);
      if (!s.sbi_this.valid) { sbi_make(s.cur, prm.rotation_estimator_blur, s.sbi_this); s.sbi_last = s.sbi_this; }
      else { s.sbi_last = s.sbi_this; sbi_make(s.cur, prm.rotation_estimator_blur, s.sbi_this); }
    }
    s.st.frame++;
    s.pose = SE3::from12(s.st.se3_cam_from_world);
    s.res.recovery = 0; s.res.reloc_keyframe = -1; s.res.reloc_score = 0.0;
    if (relocaliser_on() && s.st.lost_frames >= 3) {
      // Tracker.cc:170-178: AttemptRecovery (Tracker.cc:196-207, Relocaliser.cc:12-38), then TrackMap and
      // AssessTrackingQuality only — no motion model on this frame
      SBI cur;
      sbi_make(s.cur, 2.5, cur);  // kCurrent.pSBI = new SmallBlurryImage(kCurrent): default blur
      double best_score = 99999999999999.9;
      int best = -1;
      for (size_t i = 0; i < store.size(); i++) {
        double ssd = 0.0;  // SSDofImgs (ImageProcess.cc:88-105)
        for (int k = 0; k < cur.w * cur.h; k++) { const double dd = cur.tmpl[k] - kf_sbi[i].tmpl[k]; ssd += dd * dd; }
        if (ssd < best_score) { best_score = ssd; best = (int)i; }
      }
      double score = 0.0;
      const SE2 se2 = sbi_iterate(cur, kf_sbi[best], 6, score);
      SE3 rot;
      sbi_so3_from_se2(se2, cam_small, cur.w, cur.h, rot.R);
      const SE3 best_pose = se3_mul(rot, kf_pose[best]);
      s.res.reloc_keyframe = best; s.res.reloc_score = score;
      ptam_track_result& r = s.res;
      if (!(score < 9e6)) {  // Reloc2.MaxScore
        s.res.recovery = 2;
        s.pose.to12(r.se3_cam_from_world);
        r.scene_depth_mean = s.st.scene_depth_mean; r.scene_depth_sigma = s.st.scene_depth_sigma;
        for (int i = 0; i < 4; i++) { r.meas_attempted[i] = r.meas_found[i] = 0; r.n_pvs[i] = 0; r.n_corners[i] = (int)s.cur.lev[i].corners.size(); }
        r.did_coarse = 0; r.n_coarse = r.n_level3 = r.n_fine = 0; r.quality_needs_kf_distance = 0;
        r.tracking_quality = s.st.tracking_quality; r.n_candidates = 0;
        return;
      }
      s.res.recovery = 1;
      s.pose = s.start = best_pose;
      for (int k = 0; k < 6; k++) s.st.velocity[k] = 0.0;
      s.st.just_recovered_so_use_coarse = 1;
      track_map(s);
      assess_and_report(s);
      return;
    }
    // PredictPoseWithMotionModel (Tracker.cc:1012-1029)
    s.start = s.pose;
    double vpred[6];
    for (int k = 0; k < 6; k++) vpred[k] = s.st.velocity[k];
    {
      SplitTimer tm(2);
      sbi_make_jacs(s.sbi_last);
      const SE2 se2 = sbi_iterate(s.sbi_this, s.sbi_last, 6, s.sbi_score);  // CalcSBIRotation, nIterations = 6
      SE3 rot;
      sbi_so3_from_se2(se2, cam_small, s.sbi_this.w, s.sbi_this.h, rot.R);
      double ln6[6];
      se3_ln(rot, ln6);
      for (int k = 0; k < 3; k++) s.sbi_rot[k] = ln6[3 + k];
    }
    if (prm.use_rotation_estimator) {
      for (int k = 0; k < 3; k++) vpred[3 + k] = s.sbi_rot[k];
      vpred[0] = 0.0; vpred[1] = 0.0;
    }
    s.pose = se3_mul(se3_exp(vpred), s.start);
    track_map(s);
    // UpdateMotionModel (Tracker.cc:1035-1056)
    double motion[6];
    se3_ln(se3_mul(s.pose, se3_inverse(s.start)), motion);
    if (prm.use_constant_velocity) for (int k = 0; k < 6; k++) s.st.velocity[k] = motion[k];
    else for (int k = 0; k < 6; k++) s.st.velocity[k] = 0.9 * (0.5 * motion[k] + 0.5 * s.st.velocity[k]);
    double v6[6];
    for (int k = 0; k < 6; k++) v6[k] = s.st.velocity[k];
    for (int k = 0; k < 3; k++) v6[k] *= 1.0 / s.st.scene_depth_mean;
    double m = 0;
    for (int k = 0; k < 6; k++) m += v6[k] * v6[k];
    s.st.msd_scaled_velocity_magnitude = std::sqrt(m);
    assess_and_report(s);
  }
  void assess_and_report(Stream& s) {
    // AssessTrackingQuality (Tracker.cc:1062-1107)
    int ta = 0, tf = 0, la = 0, lf = 0;
    for (int i = 0; i < 4; i++) {
      ta += s.attempted[i]; tf += s.found[i];
      if (i >= 2) { la += s.attempted[i]; lf += s.found[i]; }
    }
    s.res.quality_needs_kf_distance = 0;
    if (tf == 0 || ta == 0) s.st.tracking_quality = 0;
    else {
      double tfrac = (double)tf / ta;
      double lfrac = la > 10 ? (double)lf / la : tfrac;
      if (tfrac > prm.quality_good) s.st.tracking_quality = 2;
      else if (lfrac < prm.quality_lost) s.st.tracking_quality = 0;
      else s.res.quality_needs_kf_distance = 1;
    }
    if (s.st.tracking_quality == 0) s.st.lost_frames++; else s.st.lost_frames = 0;
    s.pose.to12(s.st.se3_cam_from_world);
    // result
    ptam_track_result& r = s.res;
    s.pose.to12(r.se3_cam_from_world);
    r.scene_depth_mean = s.st.scene_depth_mean; r.scene_depth_sigma = s.st.scene_depth_sigma;
    for (int i = 0; i < 4; i++) {
      r.meas_attempted[i] = s.attempted[i]; r.meas_found[i] = s.found[i];
      r.n_corners[i] = (int)s.cur.lev[i].corners.size();
    }
    r.did_coarse = s.did_coarse;
    r.tracking_quality = s.st.tracking_quality;
    r.n_candidates = 0;  // product-only diagnostic (roofline accounting); not part of the parity contract
  }
};

}  // namespace orc

using orc::Tracker;

extern "C" {

void orc_tracker_default_params(ptam_tracker_params* p) {
  p->coarse_min = 20; p->coarse_max = 60; p->coarse_range = 30; p->coarse_subpix_its = 8;
  p->disable_coarse = 0; p->max_patches_per_frame = 1000; p->mestimator = 0; p->use_constant_velocity = 1;
  p->coarse_min_velocity = 0.006; p->quality_good = 0.3; p->quality_lost = 0.13;
  p->use_rotation_estimator = 1; p->reserved0 = 0; p->rotation_estimator_blur = 0.75;  // Tracker.cc:95-96
}

void* orc_tracker_create(int, const double* cam_params, int width, int height, int n_streams, const ptam_tracker_params* params) {
  Tracker* t = new Tracker;
  t->cam.init(cam_params, width, height);
  t->cam_small.init(cam_params, ((width / 8) / 2), ((height / 8) / 2));
  t->W = width; t->H = height; t->S = n_streams;
  if (params) t->prm = *params; else orc_tracker_default_params(&t->prm);
  t->streams.resize(n_streams);
  for (auto& s : t->streams) {
    std::memset(&s.st, 0, sizeof(s.st));
    std::memset(&s.res, 0, sizeof(s.res));
    orc::SE3().to12(s.st.se3_cam_from_world);
    s.st.tracking_quality = 2;
    s.st.scene_depth_mean = 1.0; s.st.scene_depth_sigma = 1.0;
  }
  return t;
}
void orc_tracker_destroy(void* t) { delete (Tracker*)t; }
const char* orc_tracker_last_error(const void* t) { return ((const Tracker*)t)->err.c_str(); }

int orc_tracker_add_keyframe(void* tp, const uint8_t* image, int stride) {
  Tracker* t = (Tracker*)tp;
  t->store.emplace_back();
  orc::make_keyframe_lite(t->store.back(), image, t->W, t->H, stride, false);
  t->kf_pose.emplace_back(); t->kf_has_pose.push_back(0);
  t->kf_sbi.emplace_back();
  orc::sbi_make(t->store.back(), 2.5, t->kf_sbi.back());   // MakeKeyFrame_Rest: pSBI = new SmallBlurryImage(*this); MakeJacs
  orc::sbi_make_jacs(t->kf_sbi.back());
  return (int)t->store.size() - 1;
}
int orc_tracker_set_keyframe_pose(void* tp, int kf, const double* se3) {
  Tracker* t = (Tracker*)tp;
  if (kf < 0 || kf >= (int)t->store.size()) return PTAM_ERR_INVALID;
  t->kf_pose[kf] = orc::SE3::from12(se3); t->kf_has_pose[kf] = 1;
  return PTAM_OK;
}

int orc_tracker_set_map(void* tp, int stream, int n, const double* world, const double* right, const double* down,
                        const int32_t* src_kf, const int32_t* src_level, const int32_t* center) {
  Tracker* t = (Tracker*)tp;
  if (stream < 0 || stream >= t->S) return PTAM_ERR_INVALID;
  auto& s = t->streams[stream];
  s.pts.assign(n, orc::MapPoint());
  s.td.assign(n, orc::TData());
  for (int i = 0; i < n; i++) {
    if (src_kf[i] < 0 || src_kf[i] >= (int)t->store.size() || src_level[i] < 0 || src_level[i] >= PTAM_LEVELS) return PTAM_ERR_INVALID;
    for (int k = 0; k < 3; k++) { s.pts[i].world[k] = world[3 * i + k]; s.pts[i].right[k] = right[3 * i + k]; s.pts[i].down[k] = down[3 * i + k]; }
    s.pts[i].src_kf = src_kf[i]; s.pts[i].src_level = src_level[i];
    s.pts[i].center = {center[2 * i], center[2 * i + 1]};
  }
  return 0;
}
int orc_tracker_set_state(void* tp, int stream, const ptam_tracker_state* st) {
  Tracker* t = (Tracker*)tp;
  if (stream < 0 || stream >= t->S) return PTAM_ERR_INVALID;
  t->streams[stream].st = *st; return 0;
}
int orc_tracker_get_state(void* tp, int stream, ptam_tracker_state* st) {
  Tracker* t = (Tracker*)tp;
  if (stream < 0 || stream >= t->S) return PTAM_ERR_INVALID;
  *st = t->streams[stream].st; return 0;
}
int orc_tracker_make_keyframes(void* tp, const uint8_t* const* images, int stride) {
  Tracker* t = (Tracker*)tp;
  for (int s = 0; s < t->S; s++) orc::make_keyframe_lite(t->streams[s].cur, images[s], t->W, t->H, stride);
  return 0;
}
int orc_tracker_track_frames(void* tp, const uint8_t* const* images, int stride, ptam_track_result* results) {
  Tracker* t = (Tracker*)tp;
  for (int s = 0; s < t->S; s++) {
    t->track_frame(t->streams[s], images[s], stride);
    if (results) results[s] = t->streams[s].res;
  }
  return 0;
}
int orc_tracker_synchronize(void*) { return 0; }
// test / bench infrastructure only (no product counterpart): this thread's CPU cost split, see g_split
void orc_tracker_cpu_split(double out[6], int reset) {
  for (int k = 0; k < 6; k++) { out[k] = orc::g_split[k]; if (reset) orc::g_split[k] = 0.0; }
}
int orc_tracker_level_size(const void* tp, int level, int* w, int* h) {
  const Tracker* t = (const Tracker*)tp;
  int ww = t->W, hh = t->H;
  for (int l = 0; l < level; l++) { ww /= 2; hh /= 2; }
  *w = ww; *h = hh; return 0;
}
int orc_tracker_get_level(void* tp, int stream, int level, uint8_t* pixels, int32_t* corners_xy, int cap, int32_t* row_lut) {
  Tracker* t = (Tracker*)tp;
  const orc::Level& L = t->streams[stream].cur.lev[level];
  if (pixels) std::memcpy(pixels, L.im.data(), L.im.size());
  if (corners_xy)
    for (size_t i = 0; i < L.corners.size() && (int)i < cap; i++) { corners_xy[2 * i] = L.corners[i].x; corners_xy[2 * i + 1] = L.corners[i].y; }
  if (row_lut) for (size_t i = 0; i < L.lut.size(); i++) row_lut[i] = L.lut[i];
  return (int)L.corners.size();
}
int orc_tracker_refind_in_keyframes(void* tp, const uint8_t* const* images, int stride, const double* se3) {
  Tracker* t = (Tracker*)tp;
  for (int s = 0; s < t->S; s++) {
    orc::make_keyframe_lite(t->streams[s].cur, images[s], t->W, t->H, stride);
    t->refind(t->streams[s], orc::SE3::from12(se3 + 12 * s));
  }
  return PTAM_OK;
}
int orc_patch_search_batch(void* tp, const double* se3, unsigned range, int subpix_its) {
  Tracker* t = (Tracker*)tp;
  if (!se3 || subpix_its < 0) return PTAM_ERR_INVALID;
  for (int s = 0; s < t->S; s++) {
    if (t->streams[s].cur.lev[0].im.empty()) return PTAM_ERR_INVALID;
    t->patch_search(t->streams[s], orc::SE3::from12(se3 + 12 * s), range, subpix_its);
  }
  return PTAM_OK;
}
int orc_patch_get_results(void* tp, int stream, int32_t* level, double* warp_inverse, int32_t* template_bad, int32_t* found,
                          double* pos, int32_t* subpix_converged) {
  Tracker* t = (Tracker*)tp;
  if (stream < 0 || stream >= t->S) return PTAM_ERR_INVALID;
  auto& s = t->streams[stream];
  for (size_t i = 0; i < s.pts.size(); i++) {
    const orc::TData& d = s.td[i];
    const bool fnd = d.in_pvs && d.found;
    if (level) level[i] = d.n_search_level;
    if (warp_inverse) for (int q = 0; q < 4; q++) warp_inverse[4 * i + q] = d.unit_projected ? d.warp_inv[q] : 0.0;
    if (template_bad) template_bad[i] = (d.unit_projected && d.template_bad) ? 1 : 0;
    if (found) found[i] = fnd ? 1 : 0;
    if (pos) { pos[2 * i] = fnd ? d.v2found[0] : 0.0; pos[2 * i + 1] = fnd ? d.v2found[1] : 0.0; }
    if (subpix_converged) subpix_converged[i] = (fnd && d.did_subpix) ? 1 : 0;
  }
  return (int)s.pts.size();
}
int orc_pose_update(void* tp, double override_sigma_squared, int mark_outliers, double* mu6, int32_t* n_found) {
  Tracker* t = (Tracker*)tp;
  for (int s = 0; s < t->S; s++) {
    double mu[6];
    const int nf = t->pose_update(t->streams[s], override_sigma_squared, mark_outliers != 0, mu);
    if (mu6) for (int k = 0; k < 6; k++) mu6[6 * s + k] = mu[k];
    if (n_found) n_found[s] = nf;
  }
  return PTAM_OK;
}
int orc_tracker_epipolar_search(void* tp, int stream, int level, int src_kf, const double* src_se3, double src_depth_mean,
                                double src_depth_sigma, const double* target_se3, double wiggle_scale, int n_cand,
                                const int32_t* cand_xy, int32_t* found, int32_t* best_corner, double* sub_pos) {
  Tracker* t = (Tracker*)tp;
  if (stream < 0 || stream >= t->S || level < 0 || level >= PTAM_LEVELS || src_kf < 0 || src_kf >= (int)t->store.size()) return PTAM_ERR_INVALID;
  t->epipolar_search(t->streams[stream], level, src_kf, orc::SE3::from12(src_se3), src_depth_mean, src_depth_sigma,
                     orc::SE3::from12(target_se3), wiggle_scale, n_cand, cand_xy, found, best_corner, sub_pos);
  return PTAM_OK;
}
int orc_tracker_keyframe_rest(void* tp, int stream, double min_shi_tomasi_score) {
  Tracker* t = (Tracker*)tp;
  if (stream < 0 || stream >= t->S) return PTAM_ERR_INVALID;
  orc::make_keyframe_rest(t->streams[stream].cur, min_shi_tomasi_score);
  return PTAM_OK;
}
int orc_tracker_get_level_rest(void* tp, int stream, int level, int32_t* max_xy, int max_cap, int32_t* cand_xy, double* cand_score,
                               int cand_cap, int* n_cand) {
  Tracker* t = (Tracker*)tp;
  if (stream < 0 || stream >= t->S || level < 0 || level >= PTAM_LEVELS) return PTAM_ERR_INVALID;
  const orc::Level& L = t->streams[stream].cur.lev[level];
  if (max_xy)
    for (size_t i = 0; i < L.max_corners.size() && (int)i < max_cap; i++) { max_xy[2 * i] = L.max_corners[i].x; max_xy[2 * i + 1] = L.max_corners[i].y; }
  for (size_t i = 0; i < L.cand_pos.size() && (int)i < cand_cap; i++) {
    if (cand_xy) { cand_xy[2 * i] = L.cand_pos[i].x; cand_xy[2 * i + 1] = L.cand_pos[i].y; }
    if (cand_score) cand_score[i] = L.cand_score[i];
  }
  if (n_cand) *n_cand = (int)L.cand_pos.size();
  return (int)L.max_corners.size();
}
int orc_tracker_get_points(void* tp, int stream, int32_t* flags, int32_t* level, double* v2_found, double* v2_image,
                           int32_t* outl, int32_t* inl) {
  Tracker* t = (Tracker*)tp;
  auto& s = t->streams[stream];
  for (size_t i = 0; i < s.pts.size(); i++) {
    const orc::TData& d = s.td[i];
    if (flags) {
      int f = 0;
      if (d.in_pvs) {
        f |= PTAM_PT_IN_PVS;
        if (d.in_image) f |= PTAM_PT_IN_IMAGE;
        if (d.searched) f |= PTAM_PT_SEARCHED;
        if (d.found) f |= PTAM_PT_FOUND;
        if (d.found && d.did_subpix) f |= PTAM_PT_SUBPIX;
      }
      if (d.template_bad) f |= PTAM_PT_TEMPLATE_BAD;
      flags[i] = f;
    }
    if (level) level[i] = d.n_search_level;
    if (v2_found) { v2_found[2 * i] = d.found && d.in_pvs ? d.v2found[0] : 0; v2_found[2 * i + 1] = d.found && d.in_pvs ? d.v2found[1] : 0; }
    if (v2_image) { v2_image[2 * i] = d.in_pvs ? d.v2image[0] : 0; v2_image[2 * i + 1] = d.in_pvs ? d.v2image[1] : 0; }
    if (outl) outl[i] = s.pts[i].outliers;
    if (inl) inl[i] = s.pts[i].inliers;
  }
  return (int)s.pts.size();
}
int orc_tracker_get_templates(void* tp, int stream, uint8_t* tmpl, int32_t* sums) {
  Tracker* t = (Tracker*)tp;
  auto& s = t->streams[stream];
  for (size_t i = 0; i < s.pts.size(); i++) {
    const orc::TData& d = s.td[i];
    if (tmpl) { if (d.has_template) std::memcpy(tmpl + 64 * i, d.tmpl, 64); else std::memset(tmpl + 64 * i, 0, 64); }
    if (sums) { sums[2 * i] = d.has_template ? d.tsum : 0; sums[2 * i + 1] = d.has_template ? d.tsumsq : 0; }
  }
  return (int)s.pts.size();
}
int orc_tracker_get_iteration_set(void* tp, int stream, int32_t* idx, int cap) {
  Tracker* t = (Tracker*)tp;
  auto& s = t->streams[stream];
  for (size_t i = 0; i < s.iter_set.size() && (int)i < cap; i++) idx[i] = s.iter_set[i];
  return (int)s.iter_set.size();
}

// ---- unit-level helpers used only by tests -------------------------------------------------
double orc_atan(double x) { return orc::spec_atan(x); }
int orc_tracker_get_sbi(void* tp, int stream, float* tmpl, int cap, double* rot3, double* score) {
  Tracker* t = (Tracker*)tp;
  if (stream < 0 || stream >= t->S) return PTAM_ERR_INVALID;
  const auto& s = t->streams[stream];
  const int n = s.sbi_this.w * s.sbi_this.h;
  if (tmpl) for (int i = 0; i < n && i < cap; i++) tmpl[i] = s.sbi_this.tmpl[i];
  if (rot3) for (int k = 0; k < 3; k++) rot3[k] = s.sbi_rot[k];
  if (score) *score = s.sbi_score;
  return n;
}
void orc_se3_exp(const double* mu, double* out12) { orc::se3_exp(mu).to12(out12); }
void orc_se3_ln(const double* in12, double* out6) { orc::se3_ln(orc::SE3::from12(in12), out6); }
void orc_cam_project(const double* params, int w, int h, const double* cam_xy, double* im_xy, double* derivs4, int* invalid) {
  orc::Camera c; c.init(params, w, h);
  auto q = c.project(cam_xy);
  im_xy[0] = q.im[0]; im_xy[1] = q.im[1];
  if (derivs4) c.derivs(q, derivs4);
  if (invalid) *invalid = q.invalid;
}
void orc_cam_unproject(const double* params, int w, int h, const double* im_xy, double* cam_xy) {
  orc::Camera c; c.init(params, w, h);
  c.unproject(im_xy, cam_xy);
}
double orc_cam_largest_radius(const double* params, int w, int h) { orc::Camera c; c.init(params, w, h); return c.largest_radius; }
int orc_zmssd(const uint8_t* im, int w, int h, int x, int y, const uint8_t* tmpl) {
  orc::Level L; L.w = w; L.h = h; L.im.assign(im, im + (size_t)w * h);
  int s = 0, ss = 0;
  for (int k = 0; k < 64; k++) { s += tmpl[k]; ss += tmpl[k] * tmpl[k]; }
  return orc::zmssd_at_point(L, orc::IRef{x, y}, tmpl, s, ss, 32000);
}
// brute-force FAST-10 straight from the definition (independent of fast10's shortcut): for tests
int orc_fast10_bruteforce(const uint8_t* im, int w, int h, int t, int32_t* xy, int cap) {
  static const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
  static const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
  int n = 0;
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++) {
      int p = im[y * w + x];
      bool corner = false;
      for (int start = 0; start < 16 && !corner; start++) {
        bool allb = true, alld = true;
        for (int k = 0; k < 10; k++) {
          int v = im[(y + dy[(start + k) & 15]) * w + x + dx[(start + k) & 15]];
          if (!(v > p + t)) allb = false;
          if (!(v < p - t)) alld = false;
        }
        corner = allb || alld;
      }
      if (corner) { if (n < cap) { xy[2 * n] = x; xy[2 * n + 1] = y; } n++; }
    }
  return n;
}
}  // extern "C"
