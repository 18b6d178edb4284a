// sm_100a kernels of path T (per-frame tracker).  Compiled with -fmad=false: the integer outputs
// (template bytes, search levels, patch offsets) depend on IEEE-exact f64 intermediates.
//
//   k_pyramid      CVD::halfSample x3 (KeyFrame.cc:25-28), one 64x64 L0 tile per CTA
//   k_fast         fast_corner_detect_10 on all four levels (KeyFrame.cc:35-42): smem-staged tiles,
//                  4 pixels per thread from 32-bit words, corner bitmask per row
//   k_compact      raster-ordered corner lists + row LUT from the bitmasks (KeyFrame.cc:46-52)
//   k_pvs_select   motion-model prediction, TrackerData::Project + GetProjectionDerivs +
//                  CalcSearchLevelAndWarpMatrix per map point, PVS lists, coarse/fine selection
//                  (Tracker.cc:454-611, PatchFinder.cc:52-84)
//   k_search       warp per point: MakeTemplateCoarseCont + FindPatchCoarse/ZMSSD + sub-pixel
//                  (Tracker.cc:867-912, PatchFinder.cc:98-318, ImageProcess.cc:130-163)
//   k_pose         one CTA per stream: the ten Gauss-Newton iterations of CalcPoseUpdate with the
//                  Tukey sigma by radix select, then motion model + quality (Tracker.cc:552-568,
//                  614-643,928-1107)
#pragma once
#include "common.cuh"
#include "../../include/ptam_b200.h"

namespace ptam {

struct LevelDesc {
  int w, h, pitch;      // pitch of the library-owned image buffer for this level
  int nwords;           // 32-pixel mask words per row
  int corner_cap;       // capacity of the corner list
  int lut_off;          // offset (ints) into the per-stream LUT buffer
  size_t img_off;       // byte offset inside a pyramid buffer
  size_t corner_off;    // offset (int2) inside the per-stream corner buffer
  size_t mask_off;      // offset (words) inside the per-stream mask buffer
  int tiles_x, tiles_y, tile_base;  // k_fast tiling (tile_base counts from level 1: level 0 has its own launch)
};

struct Geom {
  LevelDesc lev[kLevels];
  size_t pyr_bytes;      // one pyramid (L0..L3)
  size_t corner_stride;  // int2 per stream
  int lut_stride;        // ints per stream
  size_t mask_stride;    // words per stream
  int fast_tiles;        // k_fast2 tiles per frame over levels 1..3
  int thresholds[kLevels];
};

// per-stream persistent tracker state (ptam_tracker_state) + per-frame control block
struct StreamCtl {
  ptam_tracker_state st;
  double start_pose[12];  // mse3StartPos
  double pose[12];        // working pose (mse3CamFromWorld)
  int n_pvs[kLevels];
  int attempted[kLevels], found[kLevels];
  int n_coarse, n_l3, n_fine;
  int try_coarse, coarse_range, did_coarse;
  int n_corners[kLevels];
  int res_n_corners[kLevels];  // n_corners of THIS frame for the results (k_compact of the next batch may already run beside the fine pose)
  int needs_kf_distance;
  int n_cand;  // ZMSSD windows evaluated this frame (coarse + fine), for the roofline accounting
  // SmallBlurryImage rotation estimator (Tracker.cc:95-108,1012-1029)
  int sbi_valid, sbi_idx;   // a previous small image exists / which of the two buffers holds it
  double sbi_rot[3], sbi_score;
  // relocaliser (Tracker.cc:170-178): what k_reloc decided for this frame
  int frame_mode;           // 0 normal; 1 relocalised, TrackMap + quality from the recovered pose; 2 relocalisation failed
  int reloc_kf;             // Relocaliser::mnBest
  double reloc_score;       // final ESM score against that keyframe
};

// frame-scoped point flags (low byte) and persistent flags (high bits)
enum : int {
  F_IN_IMAGE = 1, F_IN_PVS = 2, F_SEARCHED = 4, F_FOUND = 8, F_SUBPIX = 16,
  F_TEMPLATE_BAD = 32,   // persistent: PatchFinder::mbTemplateBad
  F_HAS_TEMPLATE = 64,   // persistent: mpLastTemplateMapPoint == &p
  F_REFRESH = 128        // frame-scoped: k_search_prep decided that the coarse template must be re-warped
};

struct PointArrays {
  // map (inputs)
  const double* world; const double* right; const double* down;
  const int* src_kf; const int* src_level; const int2* center;
  // template cache
  uint8_t* tmpl; int* tsum; int* tsumsq; double* last_warp; double* m2;
  int4* geo;      // per point, 2 x int4: (posx, posy, r, -), (i0, i1, -, -) of FindPatchCoarse, written by k_search_prep
  // frame state
  int* flags; int* level; int* search_level;
  double* v3cam; double* v2image; double* derivs; double* warp_inv;
  double* v2found; double* sqrt_inv_noise; double* J;
  int* outliers; int* inliers;
  int* pvs;       // [S][4][cap]
  int* iter_idx;  // [S][cap]
  double* e2;     // [S][cap] scratch
  int cap;        // points per stream (stride)
};

struct FrameSrc {
  const uint8_t* l0;     // level-0 pixels of stream 0
  size_t stream_pitch;   // bytes between streams
  int pitch;             // bytes between rows
};

struct SbiDev {
  int w, h, n, ks;       // small image size ((W/8)/2 x (H/8)/2), Gaussian radius
  float taps[8];         // normalised taps, computed on the host (ImageProcess.cc:303, sigma = RotationEstimatorBlur)
  float* tmpl;           // [2][S][n] mimTemplate of this / the previous frame
  CamModel cam_small;    // the camera at the small image size (ImageProcess.cc:423)
};

struct TrackerDev {
  Geom g;
  CamModel cam;
  SbiDev sbi;
  ptam_tracker_params prm;
  FrameSrc src;
  uint8_t* pyr;          // [S] pyramids (levels 1..3 used; level 0 when frames come from the host)
  int2* corners;         // [S] corner lists
  int* lut;              // [S] row LUTs
  uint32_t* mask;        // [S] corner bitmasks
  int* ncorn;            // [S][kLevels] corners per level, counted by k_fast2 (zeroed before it)
  const uint8_t* const* kf_ptrs;  // stored keyframe pyramids
  int n_kf;
  StreamCtl* ctl;        // [S]
  const int* pt_count;   // [S]
  PointArrays p;
  int S;
  int s_off;                   // first stream of this launch (0 unless a batch is cut into groups)
  int pose_ws_smem;            // k_pose keeps the per-point working set in (dynamic) shared memory
  int mode;                    // 0: TrackFrame; 1: MapMaker::ReFindInSingleKeyFrame (MapMaker.cc:943-1040);
                               // 2: PatchFinder unit entry (ptam_patch_search_batch / ptam_pose_update)
  const double* refind_pose;   // [S][12] keyframe poses for mode 1, the caller's poses for mode 2
  unsigned unit_range;         // mode 2: FindPatchCoarse range (level-zero pixels)
  int unit_subpix_its;         // mode 2: IterateSubPixToConvergence budget, 0 = no sub-pixel step
  double unit_override_sigma;  // mode 2 pose update: dOverrideSigma (0 = M-estimator sigma)
  int unit_mark;               // mode 2 pose update: bMarkOutliers
  double* unit_mu;             // mode 2 pose update: [S][6] v6Update, [S] found counts behind them
  int* unit_nfound;
  // relocaliser (Relocaliser.cc:12-38): on once every stored keyframe has a pose
  int reloc_on;
  size_t kf_sbi_off;           // byte offset of the keyframe's SmallBlurryImage (blur 2.5) in its buffer
  const double* kf_pose;       // [n_kf][12] KeyFrame::se3CfromW
  float taps25[12];            // Gaussian taps for sigma = 2.5 (SmallBlurryImage's default blur)
  int ks25;
};

constexpr int kF2W = 128, kF2H = 48;   // k_fast2 tile

// Level geometry of a w x h frame (pyramid, masks, corner lists, k_fast2 tiling); host side.
inline void make_geom(Geom& g, int w, int h) {
  const int thr[4] = {10, 15, 15, 10};  // KeyFrame.cc:35-42
  size_t img = 0, cor = 0, msk = 0;
  int lut_o = 0, tiles = 0, lw = w, lh = h;
  for (int l = 0; l < kLevels; l++) {
    LevelDesc& L = g.lev[l];
    L.w = lw; L.h = lh; L.pitch = (lw + 15) & ~15;
    L.nwords = (lw + 31) / 32;
    L.corner_cap = (lw > 6 ? lw - 6 : 0) * (lh > 6 ? lh - 6 : 0);
    L.img_off = img; img += (size_t)L.pitch * lh; img = (img + 255) & ~(size_t)255;
    L.corner_off = cor; cor += L.corner_cap;
    L.lut_off = lut_o; lut_o += lh;
    L.mask_off = msk; msk += (size_t)L.nwords * lh;
    L.tiles_x = (lw + kF2W - 1) / kF2W; L.tiles_y = (lh + kF2H - 1) / kF2H;
    L.tile_base = l == 0 ? 0 : tiles;
    if (l > 0) tiles += L.tiles_x * L.tiles_y;
    g.thresholds[l] = thr[l];
    lw /= 2; lh /= 2;
  }
  g.pyr_bytes = img; g.corner_stride = cor; g.lut_stride = lut_o; g.mask_stride = msk; g.fast_tiles = tiles;
}

PTAM_DEV const uint8_t* level_image(const TrackerDev& d, int s, int l, int& pitch) {
  if (l == 0) { pitch = d.src.pitch; return d.src.l0 + (size_t)s * d.src.stream_pitch; }
  pitch = d.g.lev[l].pitch;
  return d.pyr + (size_t)s * d.g.pyr_bytes + d.g.lev[l].img_off;
}

// =============================================================================================
// k_pyramid — one CTA = 64x64 level-0 tile -> 32x32 L1, 16x16 L2, 8x8 L3.  256 threads, each
// reads a 4x4 block of L0 (four 32-bit loads when rows are 4-byte aligned).
// Truncating mean at every level, exactly CVD::halfSample.
// =============================================================================================
__global__ void __launch_bounds__(256) k_pyramid(TrackerDev d) {
  __shared__ uint8_t s2[16][16];
  const int s = blockIdx.z;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int x0 = blockIdx.x * 64 + tx * 4, y0 = blockIdx.y * 64 + ty * 4;
  const LevelDesc &L0 = d.g.lev[0], &L1 = d.g.lev[1], &L2 = d.g.lev[2], &L3 = d.g.lev[3];
  int p0;
  const uint8_t* im0 = level_image(d, s, 0, p0);
  uint8_t* base = d.pyr + (size_t)s * d.g.pyr_bytes;
  uint8_t* im1 = base + L1.img_off;
  uint8_t* im2 = base + L2.img_off;
  uint8_t* im3 = base + L3.img_off;
  const bool aligned = ((p0 & 3) == 0) && ((reinterpret_cast<uintptr_t>(im0) & 3) == 0);
  unsigned rows[4] = {0, 0, 0, 0};
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const int y = y0 + r;
    if (y < L0.h && x0 < L0.w) {
      const uint8_t* p = im0 + (size_t)y * p0 + x0;
      if (aligned && x0 + 3 < L0.w) rows[r] = __ldg(reinterpret_cast<const unsigned*>(p));
      else {
        unsigned v = 0;
        for (int k = 0; k < 4; k++) if (x0 + k < L0.w) v |= (unsigned)__ldg(p + k) << (8 * k);
        rows[r] = v;
      }
    }
  }
  // 2x2 L1 pixels
  int l1[2][2];
#pragma unroll
  for (int j = 0; j < 2; j++)
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const unsigned a = rows[2 * j] >> (16 * i), b = rows[2 * j + 1] >> (16 * i);
      l1[j][i] = (int)((a & 255) + ((a >> 8) & 255) + (b & 255) + ((b >> 8) & 255)) / 4;
    }
  const int x1 = x0 >> 1, y1 = y0 >> 1;
#pragma unroll
  for (int j = 0; j < 2; j++)
#pragma unroll
    for (int i = 0; i < 2; i++)
      if (x1 + i < L1.w && y1 + j < L1.h) im1[(size_t)(y1 + j) * L1.pitch + x1 + i] = (uint8_t)l1[j][i];
  const int v2 = (l1[0][0] + l1[0][1] + l1[1][0] + l1[1][1]) / 4;
  const int x2 = x0 >> 2, y2 = y0 >> 2;
  if (x2 < L2.w && y2 < L2.h) im2[(size_t)y2 * L2.pitch + x2] = (uint8_t)v2;
  s2[ty][tx] = (uint8_t)v2;
  __syncthreads();
  if (threadIdx.x < 64) {
    const int ax = threadIdx.x & 7, ay = threadIdx.x >> 3;
    const int x3 = blockIdx.x * 8 + ax, y3 = blockIdx.y * 8 + ay;
    if (x3 < L3.w && y3 < L3.h) {
      const int v3 = (s2[2 * ay][2 * ax] + s2[2 * ay][2 * ax + 1] + s2[2 * ay + 1][2 * ax] + s2[2 * ay + 1][2 * ax + 1]) / 4;
      im3[(size_t)y3 * L3.pitch + x3] = (uint8_t)v3;
    }
  }
}

PTAM_DEV bool run10(unsigned m) {
  m |= m << 16;
  unsigned a = m & (m >> 1);
  a &= a >> 2;
  a &= a >> 4;
  a &= m >> 8;
  a &= m >> 9;
  return (a & 0xFFFFu) != 0;
}

// =============================================================================================
// k_fast2 — pyramid + FAST-10, the per-frame form of KeyFrame::MakeKeyFrame_Lite's image work
// (KeyFrame.cc:25-42).  Two launches: <true> runs on level 0 and ALSO writes levels 1..3 (every CTA
// half-samples its own 128x48 tile three times: 8x8 level-0 blocks, one thread each, truncating mean
// at every level exactly as CVD::halfSample), <false> runs FAST on levels 1..3.
// Tile = 128 px x 48 rows per CTA (256 threads), staged in shared memory with a 3-row halo by 16-byte
// loads.  Warp w owns rows 6w..6w+5, lane l the pixels 4l..4l+3 of each row.  Comparisons run two pixels
// per 32-bit register in 16-bit lanes: bit 15 of  x + (0x8000 - p - t - 1)  is set iff x > p + t,  bit 15
// of  (0x8000 + p - t - 1) - x  iff x < p - t;  no lane can carry or borrow into its neighbour.
//   A. compass test on every pixel.  Any arc of >= 10 ring pixels contains two ADJACENT compass points
//      (ring 0/4/8/12 = below/right/above/left): (below|above) & (right|left), all brighter than p+t or
//      all darker than p-t.  The aligned word of a row is unpacked once and serves as the centre of row
//      y, "above" of row y+3 and "below" of row y-3.  ~9 % of level-0 pixels pass (every edge does), in
//      ~13 % of the 4-pixel words; those words go to a per-warp list (ballot + popc, no barrier).
//   B. the same warp tests its listed words for an ANTIPODAL pair: an arc of >= 10 contains five
//      consecutive even ring positions j .. j+8, hence both ends of one of the four diameters
//      (0,8) (4,12) (2,10) (6,14).  A straight edge never has both ends of a diameter on its far side, so
//      what is left (~2 % of level-0 pixels; 1.4 % are corners) goes to a CTA-wide list with the polarity.
//   C. every thread runs the 16-pixel ring test (>= 10 contiguous, strict) of ONE candidate in ONE
//      polarity: sign of  v * sg + c0  with (sg, c0) = (-1, p + t) for "brighter", (+1, t - p) for "darker".
// Output: one bit per pixel.  Raster order is restored by k_compact.
// =============================================================================================
constexpr int kF2SW = kF2W + 32;   // staged bytes per row: x0-16 .. x0+143 (ten 16-byte chunks)
constexpr int kF2SR = kF2H + 6;    // staged rows: y0-3 .. y0+50

PTAM_DEV bool ring10(const uint8_t* c, int sg, int c0) {
  unsigned m = 0;
  const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
  const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
#pragma unroll
  for (int j = 0; j < 16; j++) {
    const int v = c[dy[j] * kF2SW + dx[j]];
    m = __funnelshift_l((unsigned)(v * sg + c0), m, 1);  // shifts in the sign bit
  }
  return run10(m);
}

PTAM_DEV unsigned hsum2(unsigned a) { return (a & 0x00ff00ffu) + ((a >> 8) & 0x00ff00ffu); }  // [b0+b1, b2+b3] in 16-bit lanes
PTAM_DEV unsigned lds32(const uint8_t* p) { return *reinterpret_cast<const unsigned*>(p); }

template <bool kPyr, int kVarA = 1, int kEmit = 0, int kMinB = 6>
__global__ void __launch_bounds__(256, kMinB) k_fast2(TrackerDev d) {
  __shared__ __align__(16) uint8_t tile[kF2SR * kF2SW];
  __shared__ unsigned word_list[8][6 * 32];       // per warp: words that passed A: row - 6w | lane << 8 | pass bits (7,15,23,31)
  __shared__ uint16_t cand_list[kF2W * kF2H];     // pixels that passed B: row << 7 | x | polarity << 13
  __shared__ unsigned out_mask[kF2H][4];
  __shared__ int cand_count;
  int l = 0, tx, ty, s;
  if (kPyr) { tx = blockIdx.x; ty = blockIdx.y; s = blockIdx.z + d.s_off; }
  else {
    s = blockIdx.y + d.s_off;
    l = 1;
#pragma unroll
    for (int k = 2; k < kLevels; k++) if ((int)blockIdx.x >= d.g.lev[k].tile_base) l = k;
    const int t = blockIdx.x - d.g.lev[l].tile_base;
    ty = t / d.g.lev[l].tiles_x; tx = t - ty * d.g.lev[l].tiles_x;
  }
  const LevelDesc& L = d.g.lev[l];
  const int x0 = tx * kF2W, y0 = ty * kF2H;
  int pitch;
  const uint8_t* im = level_image(d, s, l, pitch);
  const bool al16 = ((pitch & 15) == 0) && ((reinterpret_cast<uintptr_t>(im) & 15) == 0);
  if (threadIdx.x < kF2H * 4) (&out_mask[0][0])[threadIdx.x] = 0u;
  if (threadIdx.x == 0) cand_count = 0;
  // ---- stage rows y0-3 .. y0+50, bytes x0-16 .. x0+143 (zero outside the image): thread -> one of the
  // ten 16-byte column chunks and rows r0, r0 + 25, r0 + 50, so the column tests are done once per thread
  if (threadIdx.x < 25 * (kF2SW / 16)) {
    const int r0 = threadIdx.x / (kF2SW / 16), c = threadIdx.x - r0 * (kF2SW / 16);
    const int x = x0 - 16 + 16 * c;
    const int xmode = (x + 15 < 0 || x >= L.w) ? 0 : ((al16 && x >= 0 && x + 15 < L.w) ? 1 : 2);
#pragma unroll
    for (int rr = 0; rr < 3; rr++) {
      const int r = r0 + 25 * rr;
      if (r >= kF2SR) break;
      const int y = y0 - 3 + r;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (xmode && y >= 0 && y < L.h) {
        const uint8_t* p = im + (size_t)y * pitch + x;
        if (xmode == 1) v = __ldg(reinterpret_cast<const uint4*>(p));
        else {
          unsigned w4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
          for (int k = 0; k < 16; k++)
            if (x + k >= 0 && x + k < L.w) w4[k >> 2] |= (unsigned)__ldg(p + k) << (8 * (k & 3));
          v = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
      }
      *reinterpret_cast<uint4*>(&tile[r * kF2SW + 16 * c]) = v;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ---- levels 1..3 of this tile: thread -> one 8x8 level-0 block -> 4x4, 2x2, 1 pixel(s)
  if (kPyr && threadIdx.x < (kF2W / 8) * (kF2H / 8)) {
    const int bx = threadIdx.x & 15, by = threadIdx.x >> 4;
    const LevelDesc &L1 = d.g.lev[1], &L2 = d.g.lev[2], &L3 = d.g.lev[3];
    uint8_t* base = d.pyr + (size_t)s * d.g.pyr_bytes;
    const int x1 = (x0 >> 1) + 4 * bx, y1 = (y0 >> 1) + 4 * by;
    unsigned l1a[4], l1b[4];  // level-1 pixels (0,1) and (2,3) of row j in 16-bit lanes
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint2 a = *reinterpret_cast<const uint2*>(&tile[(3 + 8 * by + 2 * j) * kF2SW + 16 + 8 * bx]);
      const uint2 b = *reinterpret_cast<const uint2*>(&tile[(4 + 8 * by + 2 * j) * kF2SW + 16 + 8 * bx]);
      l1a[j] = ((hsum2(a.x) + hsum2(b.x)) >> 2) & 0x00ff00ffu;
      l1b[j] = ((hsum2(a.y) + hsum2(b.y)) >> 2) & 0x00ff00ffu;
      const unsigned w4 = __byte_perm(l1a[j], l1b[j], 0x6420);
      if (y1 + j < L1.h) {
        uint8_t* o = base + L1.img_off + (size_t)(y1 + j) * L1.pitch + x1;
        if (x1 + 3 < L1.w) *reinterpret_cast<unsigned*>(o) = w4;
        else
          for (int k = 0; k < 3; k++) if (x1 + k < L1.w) o[k] = (uint8_t)(w4 >> (8 * k));
      }
    }
    const int x2 = (x0 >> 2) + 2 * bx, y2 = (y0 >> 2) + 2 * by;
    unsigned v2[2][2];
#pragma unroll
    for (int m = 0; m < 2; m++) {
      const unsigned ua = l1a[2 * m] + l1a[2 * m + 1], ub = l1b[2 * m] + l1b[2 * m + 1];
      v2[m][0] = ((ua + (ua >> 16)) & 0xffffu) >> 2;
      v2[m][1] = ((ub + (ub >> 16)) & 0xffffu) >> 2;
      if (y2 + m < L2.h) {
        uint8_t* o = base + L2.img_off + (size_t)(y2 + m) * L2.pitch + x2;
        if (x2 + 1 < L2.w) *reinterpret_cast<uint16_t*>(o) = (uint16_t)(v2[m][0] | (v2[m][1] << 8));
        else if (x2 < L2.w) o[0] = (uint8_t)v2[m][0];
      }
    }
    const int x3 = (x0 >> 3) + bx, y3 = (y0 >> 3) + by;
    if (x3 < L3.w && y3 < L3.h)
      base[L3.img_off + (size_t)y3 * L3.pitch + x3] = (uint8_t)((v2[0][0] + v2[0][1] + v2[1][0] + v2[1][1]) >> 2);
  }
  const int thr = min(max(d.g.thresholds[l], 0), 255);
  const unsigned K = 0x80008000u - (unsigned)(thr + 1) * 0x00010001u;
  // ---- stage A: compass test, 4 pixels x 6 rows per thread
  unsigned vm = 0;  // bit 8k+7: pixel k of this thread can be a corner (3-pixel border of the image)
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int xx = x0 + 4 * lane + k;
    if (xx >= 3 && xx < L.w - 3) vm |= 0x80u << (8 * k);
  }
  const uint8_t* tcol = &tile[(6 * warp) * kF2SW + 16 + 4 * lane];
  int n_words = 0;  // entries in this warp's list (warp-uniform)
  const unsigned lt = (1u << lane) - 1u;
  const unsigned ebase = (unsigned)lane << 8;
  if (kVarA == 1) {
    // byte lanes, four pixels per register, sign-agnostic: |x - p| > t for (below|above) & (right|left).
    // |x - p| by VABSDIFF4; d > t per byte without carries between bytes: for t < 128 bit 7 of
    // ((d & 0x7f) + (0x7f - t)) | d, for t >= 128 bit 7 of ((d & 0x7f) + (0x7f - (t & 0x7f))) & d.
    unsigned cwr[12];
#pragma unroll
    for (int r = 0; r < 12; r++) cwr[r] = lds32(tcol + r * kF2SW);
    const unsigned m7 = 0x7f7f7f7fu, c = (unsigned)(0x7f - (thr & 0x7f)) * 0x01010101u;
    const int i_lo = max(0, 3 - (y0 + 6 * warp)), i_hi = min(6, L.h - 3 - (y0 + 6 * warp));  // rows that can hold corners
#define PTAM_STAGE_A(COMBINE)                                                                                   \
    _Pragma("unroll") for (int i = 0; i < 6; i++) {                                                             \
      const unsigned w0 = lds32(tcol + (i + 3) * kF2SW - 4), w2 = lds32(tcol + (i + 3) * kF2SW + 4);            \
      const unsigned P = cwr[i + 3];                                                                            \
      const unsigned dS = __vabsdiffu4(cwr[i + 6], P), dN = __vabsdiffu4(cwr[i], P);                            \
      const unsigned dE = __vabsdiffu4(__byte_perm(P, w2, 0x6543), P), dW = __vabsdiffu4(__byte_perm(w0, P, 0x4321), P); \
      const unsigned sS = (dS & m7) + c, sN = (dN & m7) + c, sE = (dE & m7) + c, sW = (dW & m7) + c;            \
      unsigned z = (COMBINE) & vm;                                                                              \
      if (i < i_lo || i >= i_hi) z = 0;                    /* warp-uniform */                                   \
      const unsigned votes = __ballot_sync(kFull, z != 0);                                                      \
      if (z) word_list[warp][n_words + __popc(votes & lt)] = z | ebase | i;                                     \
      n_words += __popc(votes);                                                                                 \
    }
    if (thr < 128) { PTAM_STAGE_A((sS | dS | sN | dN) & (sE | dE | sW | dW)) }
    else { PTAM_STAGE_A(((sS & dS) | (sN & dN)) & ((sE & dE) | (sW & dW))) }
#undef PTAM_STAGE_A
  } else {
    unsigned lo[12], hi[12];  // aligned words of staged rows 6w .. 6w+11, pixels (0,1) and (2,3) in 16-bit lanes
#pragma unroll
    for (int r = 0; r < 12; r++) {
      const unsigned cw = lds32(tcol + r * kF2SW);
      lo[r] = __byte_perm(cw, 0u, 0x4140); hi[r] = __byte_perm(cw, 0u, 0x4342);
    }
#pragma unroll
    for (int i = 0; i < 6; i++) {
      const int y = y0 + 6 * warp + i;
      const unsigned w0 = lds32(tcol + (i + 3) * kF2SW - 4), w2 = lds32(tcol + (i + 3) * kF2SW + 4);
      const unsigned Plo = lo[i + 3], Phi = hi[i + 3];
      const unsigned lft_lo = __byte_perm(w0, 0u, 0x4241), lft_hi = __byte_perm(w0, Plo, 0x5453);   // pixels x-3 .. x
      const unsigned rgt_lo = __byte_perm(Phi, w2, 0x1412), rgt_hi = __byte_perm(w2, 0u, 0x4241);   // pixels x+3 .. x+6
      const unsigned Bl = K - Plo, Bh = K - Phi, Dl = Plo + K, Dh = Phi + K;
      const unsigned r_lo = (((lo[i + 6] + Bl) | (lo[i] + Bl)) & ((rgt_lo + Bl) | (lft_lo + Bl))) |
                            (((Dl - lo[i + 6]) | (Dl - lo[i])) & ((Dl - rgt_lo) | (Dl - lft_lo)));
      const unsigned r_hi = (((hi[i + 6] + Bh) | (hi[i] + Bh)) & ((rgt_hi + Bh) | (lft_hi + Bh))) |
                            (((Dh - hi[i + 6]) | (Dh - hi[i])) & ((Dh - rgt_hi) | (Dh - lft_hi)));
      unsigned z = __byte_perm(r_lo, r_hi, 0x7531) & vm;   // bits 15 / 31 of the lanes -> bit 7 of bytes 0..3
      if (y < 3 || y >= L.h - 3) z = 0;                    // warp-uniform
      const unsigned votes = __ballot_sync(kFull, z != 0);
      if (z) word_list[warp][n_words + __popc(votes & lt)] = z | ebase | i;
      n_words += __popc(votes);
    }
  }
  __syncwarp();
  // ---- stage B: antipodal-pair test of the listed words (same warp), survivors to the CTA-wide pixel list
  for (int wb = 0; wb < n_words; wb += 32) {
    const int wi = wb + lane;
    unsigned zb = 0, zd = 0, e0 = 0;
    if (wi < n_words) {
      const unsigned e = word_list[warp][wi];
      const int ry = 6 * warp + (int)(e & 7u), wl = (e >> 8) & 31;
      e0 = (unsigned)((ry << 7) + 4 * wl);
      const uint8_t* tc = &tile[(ry + 3) * kF2SW + 16 + 4 * wl];
      const unsigned cw = lds32(tc), w0 = lds32(tc - 4), w2 = lds32(tc + 4);
      const unsigned cS = lds32(tc + 3 * kF2SW), cN = lds32(tc - 3 * kF2SW);
      const unsigned a0 = lds32(tc + 2 * kF2SW - 4), a1 = lds32(tc + 2 * kF2SW), a2 = lds32(tc + 2 * kF2SW + 4);   // row y+2
      const unsigned b0 = lds32(tc - 2 * kF2SW - 4), b1 = lds32(tc - 2 * kF2SW), b2 = lds32(tc - 2 * kF2SW + 4);   // row y-2
      const unsigned Plo = __byte_perm(cw, 0u, 0x4140), Phi = __byte_perm(cw, 0u, 0x4342);
      const unsigned Bl = K - Plo, Bh = K - Phi, Dl = Plo + K, Dh = Phi + K;
      // ring pixels of the four pixels, (0,1) and (2,3) in 16-bit lanes
      const unsigned S_lo = __byte_perm(cS, 0u, 0x4140), S_hi = __byte_perm(cS, 0u, 0x4342);
      const unsigned N_lo = __byte_perm(cN, 0u, 0x4140), N_hi = __byte_perm(cN, 0u, 0x4342);
      const unsigned W_lo = __byte_perm(w0, 0u, 0x4241), W_hi = __byte_perm(w0, Plo, 0x5453);
      const unsigned E_lo = __byte_perm(Phi, w2, 0x1412), E_hi = __byte_perm(w2, 0u, 0x4241);
      const unsigned SW_lo = __byte_perm(a0, 0u, 0x4342), SW_hi = __byte_perm(a1, 0u, 0x4140);   // (x-2, y+2)
      const unsigned SE_lo = __byte_perm(a1, 0u, 0x4342), SE_hi = __byte_perm(a2, 0u, 0x4140);   // (x+2, y+2)
      const unsigned NW_lo = __byte_perm(b0, 0u, 0x4342), NW_hi = __byte_perm(b1, 0u, 0x4140);   // (x-2, y-2)
      const unsigned NE_lo = __byte_perm(b1, 0u, 0x4342), NE_hi = __byte_perm(b2, 0u, 0x4140);   // (x+2, y-2)
#define PTAM_AB(op, C, S_, N_, E_, W_, SE_, NW_, NE_, SW_)                                             \
      ((((op(C, S_)) | (op(C, N_))) & ((op(C, E_)) | (op(C, W_)))) &                                  \
       (((op(C, S_)) & (op(C, N_))) | ((op(C, E_)) & (op(C, W_))) | ((op(C, SE_)) & (op(C, NW_))) | ((op(C, NE_)) & (op(C, SW_)))))
#define PTAM_BR(C, X) ((X) + (C))
#define PTAM_DK(C, X) ((C) - (X))
      const unsigned rb_lo = PTAM_AB(PTAM_BR, Bl, S_lo, N_lo, E_lo, W_lo, SE_lo, NW_lo, NE_lo, SW_lo);
      const unsigned rb_hi = PTAM_AB(PTAM_BR, Bh, S_hi, N_hi, E_hi, W_hi, SE_hi, NW_hi, NE_hi, SW_hi);
      const unsigned rd_lo = PTAM_AB(PTAM_DK, Dl, S_lo, N_lo, E_lo, W_lo, SE_lo, NW_lo, NE_lo, SW_lo);
      const unsigned rd_hi = PTAM_AB(PTAM_DK, Dh, S_hi, N_hi, E_hi, W_hi, SE_hi, NW_hi, NE_hi, SW_hi);
#undef PTAM_AB
#undef PTAM_BR
#undef PTAM_DK
      zb = __byte_perm(rb_lo, rb_hi, 0x7531) & e & 0x80808080u;
      zd = __byte_perm(rd_lo, rd_hi, 0x7531) & e & 0x80808080u;
    }
    const unsigned both = zb | zd;
    if (kEmit == 1) {
      // survivors of the warp's 32 words to the pixel list: one ballot per pixel slot, one atomic per round
      const unsigned v0 = __ballot_sync(kFull, both & 0x80u), v1 = __ballot_sync(kFull, both & 0x8000u);
      const unsigned v2 = __ballot_sync(kFull, both & 0x800000u), v3 = __ballot_sync(kFull, both & 0x80000000u);
      const int n0 = __popc(v0), n1 = __popc(v1), n2 = __popc(v2), n3 = __popc(v3);
      int o = 0;
      if (lane == 0 && n0 + n1 + n2 + n3) o = atomicAdd(&cand_count, n0 + n1 + n2 + n3);
      o = __shfl_sync(kFull, o, 0);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const unsigned vk = k == 0 ? v0 : k == 1 ? v1 : k == 2 ? v2 : v3;
        const int b = 8 * k + 7;
        if ((both >> b) & 1u)
          cand_list[o + __popc(vk & lt)] = (uint16_t)((e0 + k) | (((zb >> b) & 1u) << 13) | (((zd >> b) & 1u) << 14));
        o += k == 0 ? n0 : k == 1 ? n1 : k == 2 ? n2 : n3;
      }
    } else if (both) {
      unsigned m = both;
      int o = atomicAdd(&cand_count, __popc(m));
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        cand_list[o++] = (uint16_t)((e0 + (b >> 3)) | (((zb >> b) & 1u) << 13) | (((zd >> b) & 1u) << 14));
      }
    }
  }
  __syncthreads();
  // ---- stage C: ring test, one candidate per thread
  const int n_cand = cand_count;
  for (int ci = threadIdx.x; ci < n_cand; ci += 256) {
    const unsigned e = cand_list[ci];
    const int pos = e & 0x1fff, pol = e >> 13;
    const int ry = pos >> 7, px = pos & 127;
    const uint8_t* c = &tile[(ry + 3) * kF2SW + 16 + px];
    const int p = *c;
    const int sg = (pol & 1) ? -1 : 1, c0 = (pol & 1) ? p + thr : thr - p;
    bool corner = ring10(c, sg, c0);
    if (pol == 3 && !corner) corner = ring10(c, 1, thr - p);  // rare: both polarities passed B
    if (corner) atomicOr(&out_mask[ry][px >> 5], 1u << (px & 31));
  }
  __syncthreads();
  int n_here = 0;
  if (threadIdx.x < kF2H * 4) {
    const int ry = threadIdx.x >> 2, wq = threadIdx.x & 3;
    const int y = y0 + ry, word = (x0 >> 5) + wq;
    if (y < L.h && word < L.nwords) {
      d.mask[(size_t)s * d.g.mask_stride + L.mask_off + (size_t)y * L.nwords + word] = out_mask[ry][wq];
      n_here = __popc(out_mask[ry][wq]);
    }
  }
  if (threadIdx.x < 192) {  // corners of this tile (warps 0..5 hold the mask words) -> the level's counter
    n_here = __reduce_add_sync(kFull, n_here);
    if ((threadIdx.x & 31) == 0 && n_here) atomicAdd(&d.ncorn[s * kLevels + l], n_here);
  }
}

// =============================================================================================
// k_compact — one CTA per (level, stream).  The corner mask of a level is one contiguous array of
// words (rows x nwords), so raster order is word order: a block-wide exclusive scan of the popcounts,
// 1024 words per round, gives every word the index of its first corner; the word that starts row y
// also holds LUT[y] = number of corners in rows < y (KeyFrame.cc:46-52).
// =============================================================================================
__global__ void __launch_bounds__(1024) k_compact(TrackerDev d) {
  __shared__ int wsum[32];
  const int l = blockIdx.x, s = blockIdx.y + d.s_off;
  const LevelDesc& L = d.g.lev[l];
  const uint32_t* mask = d.mask + (size_t)s * d.g.mask_stride + L.mask_off;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nw = L.nwords, total = L.h * nw;
  int* lut = d.lut + (size_t)s * d.g.lut_stride + L.lut_off;
  int2* out = d.corners + (size_t)s * d.g.corner_stride + L.corner_off;
  int carry = 0;  // corners in the words of earlier rounds (the same in every thread)
  for (int base = 0; base < total; base += 1024) {
    const int w = base + (int)threadIdx.x;
    unsigned m = w < total ? mask[w] : 0u;
    const int c = __popc(m);
    int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int ws = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(kFull, ws, o);
        if (lane >= o) ws += v;
      }
      wsum[lane] = ws;
    }
    __syncthreads();
    int o = carry + (warp ? wsum[warp - 1] : 0) + inc - c;
    carry += wsum[31];
    if (w < total) {
      const int y = w / nw, wi = w - y * nw;
      if (wi == 0) lut[y] = o;
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        out[o++] = make_int2(wi * 32 + b, y);
      }
    }
    __syncthreads();  // wsum is rewritten by the next round
  }
  if (threadIdx.x == 0) d.ctl[s].n_corners[l] = carry;
}

// =============================================================================================
// k_sbi — one CTA per stream: SmallBlurryImage::MakeFromKF of the current frame (halfSample of level 3,
// zero mean, Gaussian blur; ImageProcess.cc:279-304), MakeJacs of the previous one (:170-191), six
// ESM iterations of IteratePosRelToTarget (:313-412) and SE3fromSE2 (:421-473).  Result: the so3
// rotation vector PredictPoseWithMotionModel substitutes for the rotational velocity.
// Float arithmetic follows the specification in oracle/oracle_tracker.cpp operation by operation
// (no FMA contraction), so the small image is bit-identical; the ESM sums are reduced in parallel
// (different order), and warp positions are evaluated directly instead of accumulated.
// =============================================================================================
struct Se2 { double c, s, tx, ty; };
PTAM_DEV Se2 se2_mul(const Se2& a, const Se2& b) {
  Se2 r;
  r.c = a.c * b.c - a.s * b.s; r.s = a.s * b.c + a.c * b.s;
  r.tx = a.tx + (a.c * b.tx - a.s * b.ty);
  r.ty = a.ty + (a.s * b.tx + a.c * b.ty);
  return r;
}

struct SbiScratch {
  float *t, *hrow, *cur, *prev, *jx, *jy, *warped;
  uint8_t* small;
  double (*red)[16];
  double* fin;
  unsigned* usum;
  Se2* c2c;
  double* mean_off;
};

PTAM_DEV SbiScratch sbi_scratch(unsigned char* raw, int n, double (*red)[16], double* fin, unsigned* usum, Se2* c2c, double* mean_off) {
  SbiScratch m;
  m.t = reinterpret_cast<float*>(raw);
  m.hrow = m.t + n; m.cur = m.hrow + n; m.prev = m.cur + n;
  m.jx = m.prev + n; m.jy = m.jx + n; m.warped = m.jy + n;
  m.small = reinterpret_cast<uint8_t*>(m.warped + n);
  m.red = red; m.fin = fin; m.usum = usum; m.c2c = c2c; m.mean_off = mean_off;
  return m;
}

// SmallBlurryImage::MakeFromKF (ImageProcess.cc:279-304): halfSample of level 3, zero mean, Gaussian blur
// with the given taps.  Whole CTA; the result is in m.cur (synchronised).
PTAM_DEV void sbi_make_small(const SbiDev& sb, const float* taps, int ks, const uint8_t* l3, int pitch, const SbiScratch& m) {
  const int n = sb.n, w = sb.w, h = sb.h;
  float* t = m.t; float* hrow = m.hrow; float* cur = m.cur; uint8_t* small = m.small; unsigned* usum = m.usum;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // ---- halfSample(level 3) and its sum
  unsigned part = 0;
  for (int i = tid; i < n; i += blockDim.x) {
    const int y = i / w, x = i - y * w;
    const uint8_t* r0 = l3 + (size_t)(2 * y) * pitch + 2 * x;
    const uint8_t* r1 = r0 + pitch;
    const int v = ((int)r0[0] + r0[1] + r1[0] + r1[1]) / 4;
    small[i] = (uint8_t)v;
    part += (unsigned)v;
  }
  part = __reduce_add_sync(kFull, part);
  if (lane == 0) usum[warp] = part;
  __syncthreads();
  unsigned total = 0;
  for (int q = 0; q < 8; q++) total += usum[q];
  const float mean = ((float)total) / (float)n;
  for (int i = tid; i < n; i += blockDim.x) t[i] = (float)small[i] - mean;
  __syncthreads();
  // ---- Gaussian blur: rows, then columns; replicated borders
  for (int i = tid; i < n; i += blockDim.x) {
    const int y = i / w, x = i - y * w;
    float a = t[i] * taps[0];
    for (int k = 1; k <= ks; k++) a += (t[y * w + max(x - k, 0)] + t[y * w + min(x + k, w - 1)]) * taps[k];
    hrow[i] = a;
  }
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) {
    const int y = i / w, x = i - y * w;
    float a = hrow[i] * taps[0];
    for (int k = 1; k <= ks; k++) a += (hrow[max(y - k, 0) * w + x] + hrow[min(y + k, h - 1) * w + x]) * taps[k];
    cur[i] = a;
  }
  __syncthreads();
}

// MakeJacs of m.prev (ImageProcess.cc:170-191), six ESM iterations of IteratePosRelToTarget of m.cur against
// it (:313-412) and SE3fromSE2 (:421-473).  Whole CTA; thread 0 returns the rotation matrix and the final score.
PTAM_DEV double sbi_esm_rotation(const SbiDev& sb, const SbiScratch& m, double* R) {
  const int n = sb.n, w = sb.w, h = sb.h;
  float* cur = m.cur; float* prev = m.prev; float* jx = m.jx; float* jy = m.jy; float* warped = m.warped;
  double (*red)[16] = m.red; double* fin = m.fin;
  Se2& c2c_s = *m.c2c; double& mean_off_s = *m.mean_off;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // ---- MakeJacs of the previous small image
  for (int i = tid; i < n; i += blockDim.x) {
    const int y = i / w, x = i - y * w;
    float gx = 0.f, gy = 0.f;
    if (x >= 1 && y >= 1 && x < w - 1 && y < h - 1) { gx = prev[i + 1] - prev[i - 1]; gy = prev[i + w] - prev[i - w]; }
    jx[i] = gx; jy[i] = gy;
  }
  if (tid == 0) { c2c_s.c = 1.0; c2c_s.s = 0.0; c2c_s.tx = 0.0; c2c_s.ty = 0.0; mean_off_s = 0.0; }
  __syncthreads();
  // ---- IteratePosRelToTarget, nIterations = 6
  const int cx = w / 2, cy = h / 2;
  double final_score = 0.0;
  for (int it = 0; it < 6; it++) {
    const Se2 c2c = c2c_s;
    const double mean_off = mean_off_s;
    Se2 wfc{1.0, 0.0, (double)cx, (double)cy}, wfc_inv{1.0, 0.0, -(double)cx, -(double)cy};
    const Se2 xf = se2_mul(se2_mul(wfc, c2c), wfc_inv);
    const double xb = w - 1, yb = h - 1;
    for (int i = tid; i < n; i += blockDim.x) {
      const int y = i / w, x = i - y * w;
      const double px = xf.tx + ((double)x * xf.c + (double)y * (-xf.s));
      const double py = xf.ty + ((double)x * xf.s + (double)y * xf.c);
      float v = -9e20f;
      if (0 <= px && 0 <= py && px < xb && py < yb) {
        const int lx = (int)px, ly = (int)py;
        const double fx = px - lx, fy = py - ly;
        const float* r0 = cur + ly * w + lx;
        const float* r1 = r0 + w;
        v = (float)((1 - fy) * ((1 - fx) * (double)r0[0] + fx * (double)r0[1]) + fy * ((1 - fx) * (double)r1[0] + fx * (double)r1[1]));
      }
      warped[i] = v;
    }
    __syncthreads();
    double a[15];
#pragma unroll
    for (int q = 0; q < 15; q++) a[q] = 0.0;
    for (int i = tid; i < n; i += blockDim.x) {
      const int y = i / w, x = i - y * w;
      if (!(x >= 1 && y >= 1 && x < w - 1 && y < h - 1)) continue;
      const float l = warped[i - 1], r = warped[i + 1], u = warped[i - w], dn = warped[i + w], here = warped[i];
      if (l + r + u + dn + here < -9999.9) continue;
      const double g0 = r - l, g1 = dn - u;
      const double sg0 = 0.25 * (g0 + (double)jx[i]), sg1 = 0.25 * (g1 + (double)jy[i]);
      const double J0 = sg0, J1 = sg1, J2 = -(y - cy) * sg0 + (x - cx) * sg1;
      const double diff = here - prev[i] + mean_off;
      a[14] += diff * diff;
      a[0] += diff * J0; a[1] += diff * J1; a[2] += diff * J2; a[3] += diff;
      a[4] += J0 * J0; a[5] += J1 * J0; a[6] += J1 * J1; a[7] += J2 * J0; a[8] += J2 * J1; a[9] += J2 * J2;
      a[10] += J0; a[11] += J1; a[12] += J2; a[13] += 1.0;
    }
#pragma unroll
    for (int q = 0; q < 15; q++) a[q] = warp_sum(a[q]);
    if (lane == 0)
#pragma unroll
      for (int q = 0; q < 15; q++) red[warp][q] = a[q];
    __syncthreads();
    if (tid < 15) {
      double v = 0;
      for (int q = 0; q < 8; q++) v += red[q][tid];
      fin[tid] = v;
    }
    __syncthreads();
    if (tid == 0) {
      double M[16], upd[4];
      int v = 0;
      for (int j = 0; j < 4; j++)
        for (int i = 0; i <= j; i++) { M[4 * j + i] = fin[4 + v]; M[4 * i + j] = fin[4 + v]; v++; }
      ldlt_factor<4>(M);
      ldlt_backsub<4>(M, fin, upd);
      Se2 u;
      u.tx = -upd[0]; u.ty = -upd[1];
      u.c = cos(-upd[2]); u.s = sin(-upd[2]);
      c2c_s = se2_mul(c2c, u);
      mean_off_s = mean_off - upd[3];
      final_score = fin[14];
    }
    __syncthreads();
  }
  if (tid != 0) return 0.0;
  // ---- SE3fromSE2 (ImageProcess.cc:421-473) and its logarithm
  const Se2 se2 = c2c_s;
  const CamModel& cam = sb.cam_small;
  double turned[2][2], orig[2][3];
  const double off[2] = {5.0, -5.0};
  for (int i = 0; i < 2; i++) {
    turned[i][0] = (double)cx + (se2.c * off[i] + se2.tx);
    turned[i][1] = (double)cy + (se2.s * off[i] + se2.ty);
    const double d0 = (((double)cx + off[i]) - cam.center[0]) * cam.inv_focal[0];
    const double d1 = ((double)cy - cam.center[1]) * cam.inv_focal[1];
    const double dr = sqrt(d0 * d0 + d1 * d1);
    const double rr = cam.w == 0.0 ? dr : tan(dr * cam.w) * cam.one_over_tan2;  // invrtrans (ATANCamera.h:151-157)
    const double f = dr > 0.01 ? rr / dr : 1.0;
    orig[i][0] = f * d0; orig[i][1] = f * d1; orig[i][2] = 1.0;
  }
  R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
  for (int it = 0; it < 3; it++) {
    double C[9] = {10, 0, 0, 0, 10, 0, 0, 0, 10}, b[3] = {0, 0, 0};
    for (int i = 0; i < 2; i++) {
      double vc[3];
      so3_rotate(R, orig[i], vc);
      const CamProj q = cam_project(cam, vc[0] / vc[2], vc[1] / vc[2]);
      const double err[2] = {turned[i][0] - q.im[0], turned[i][1] - q.im[1]};
      double dv[4];
      cam_derivs(cam, q, dv);
      const double ooz = 1.0 / vc[2];
      const double gen[3][3] = {{0, -vc[2], vc[1]}, {vc[2], 0, -vc[0]}, {-vc[1], vc[0], 0}};
      double Jr[2][3];
      for (int m = 0; m < 3; m++) {
        const double a0 = (gen[m][0] - vc[0] * gen[m][2] * ooz) * ooz, a1 = (gen[m][1] - vc[1] * gen[m][2] * ooz) * ooz;
        Jr[0][m] = dv[0] * a0 + dv[1] * a1;
        Jr[1][m] = dv[2] * a0 + dv[3] * a1;
      }
      for (int r = 0; r < 2; r++)
        for (int aa = 0; aa < 3; aa++) {
          for (int c = 0; c < 3; c++) C[3 * aa + c] += Jr[r][aa] * Jr[r][c];
          b[aa] += err[r] * Jr[r][aa];
        }
    }
    double x[3], E[9], N[9];
    ldlt_factor<3>(C);
    ldlt_backsub<3>(C, b, x);
    so3_exp(x, E);
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) N[3 * r + c] = E[3 * r] * R[c] + E[3 * r + 1] * R[3 + c] + E[3 * r + 2] * R[6 + c];
    for (int i = 0; i < 9; i++) R[i] = N[i];
  }
  return final_score;
}

__global__ void __launch_bounds__(256) k_sbi(TrackerDev d) {
  extern __shared__ __align__(16) unsigned char sbi_raw[];
  __shared__ double red[8][16];
  __shared__ double fin[16];
  __shared__ unsigned usum[8];
  __shared__ Se2 c2c_s;
  __shared__ double mean_off_s;
  const SbiDev& sb = d.sbi;
  const int n = sb.n;
  const SbiScratch m = sbi_scratch(sbi_raw, n, red, fin, usum, &c2c_s, &mean_off_s);
  const int s = blockIdx.x + d.s_off, tid = threadIdx.x;
  StreamCtl& ctl = d.ctl[s];
  int pitch;
  const uint8_t* l3 = level_image(d, s, 3, pitch);
  sbi_make_small(sb, sb.taps, sb.ks, l3, pitch, m);
  const int idx_prev = ctl.sbi_idx, has_prev = ctl.sbi_valid;
  float* g_cur = sb.tmpl + ((size_t)(1 - idx_prev) * d.S + s) * n;
  const float* g_prev = sb.tmpl + ((size_t)idx_prev * d.S + s) * n;
  for (int i = tid; i < n; i += blockDim.x) {
    const float a = m.cur[i];
    g_cur[i] = a;
    m.prev[i] = has_prev ? g_prev[i] : a;  // first frame: both small images are made from it (Tracker.cc:99-100)
  }
  __syncthreads();
  double R[9];
  const double final_score = sbi_esm_rotation(sb, m, R);
  if (tid != 0) return;
  double rot[3];
  so3_ln(R, rot);
  ctl.sbi_rot[0] = rot[0]; ctl.sbi_rot[1] = rot[1]; ctl.sbi_rot[2] = rot[2];
  ctl.sbi_score = final_score;
  ctl.sbi_idx = 1 - idx_prev;
  ctl.sbi_valid = 1;
}

// =============================================================================================
// Relocaliser.
//   k_kf_sbi   one CTA: the SmallBlurryImage (default blur 2.5) KeyFrame::MakeKeyFrame_Rest gives a keyframe
//              (KeyFrame.cc:80-81), stored behind the keyframe's pyramid.
//   k_reloc    one CTA per stream, a no-op unless the stream arrives with mnLostFrames >= 3 (Tracker.cc:133):
//              Relocaliser::AttemptRecovery (Relocaliser.cc:12-38) — small blurry image of the current frame,
//              SSD against every stored keyframe's (block-reduced in f64), first minimum, ESM rotation
//              against it (the same code as k_sbi), pose = rotation * keyframe pose — and
//              Tracker::AttemptRecovery (Tracker.cc:196-207): pose, zero velocity, mbJustRecoveredSoUseCoarse.
//              k_pvs_select / k_pose then skip the motion model (frame_mode 1) or the whole frame (2).
// =============================================================================================
__global__ void __launch_bounds__(256) k_kf_sbi(TrackerDev d, int kf) {
  extern __shared__ __align__(16) unsigned char sbi_raw[];
  __shared__ double red[8][16];
  __shared__ double fin[16];
  __shared__ unsigned usum[8];
  __shared__ Se2 c2c_s;
  __shared__ double mean_off_s;
  const SbiDev& sb = d.sbi;
  const SbiScratch m = sbi_scratch(sbi_raw, sb.n, red, fin, usum, &c2c_s, &mean_off_s);
  const uint8_t* base = d.kf_ptrs[kf];
  sbi_make_small(sb, d.taps25, d.ks25, base + d.g.lev[3].img_off, d.g.lev[3].pitch, m);
  float* out = reinterpret_cast<float*>(const_cast<uint8_t*>(base) + d.kf_sbi_off);
  for (int i = threadIdx.x; i < sb.n; i += blockDim.x) out[i] = m.cur[i];
}

__global__ void __launch_bounds__(256) k_reloc(TrackerDev d) {
  extern __shared__ __align__(16) unsigned char sbi_raw[];
  __shared__ double red[8][16];
  __shared__ double fin[16];
  __shared__ unsigned usum[8];
  __shared__ Se2 c2c_s;
  __shared__ double mean_off_s;
  __shared__ int best_s;
  const SbiDev& sb = d.sbi;
  const int n = sb.n;
  const int s = blockIdx.x + d.s_off, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  StreamCtl& ctl = d.ctl[s];
  if (ctl.st.lost_frames < 3) {  // uniform: nothing below writes lost_frames
    if (tid == 0) { ctl.frame_mode = 0; ctl.reloc_kf = -1; ctl.reloc_score = 0.0; }
    return;
  }
  const SbiScratch m = sbi_scratch(sbi_raw, n, red, fin, usum, &c2c_s, &mean_off_s);
  int pitch;
  const uint8_t* l3 = level_image(d, s, 3, pitch);
  sbi_make_small(sb, d.taps25, d.ks25, l3, pitch, m);  // kCurrent.pSBI = new SmallBlurryImage(kCurrent)
  // SSDofImgs against every keyframe (ImageProcess.cc:88-105), strict '<': the first minimum wins
  double best_score = 99999999999999.9;
  int best = -1;
  for (int kf = 0; kf < d.n_kf; kf++) {
    const float* kt = reinterpret_cast<const float*>(d.kf_ptrs[kf] + d.kf_sbi_off);
    double part = 0.0;
    for (int i = tid; i < n; i += blockDim.x) { const double dd = m.cur[i] - kt[i]; part += dd * dd; }
    part = warp_sum(part);
    if (lane == 0) red[warp][0] = part;
    __syncthreads();
    if (tid == 0) {
      double ssd = 0.0;
      for (int q = 0; q < 8; q++) ssd += red[q][0];
      if (ssd < best_score) { best_score = ssd; best = kf; }
    }
    __syncthreads();
  }
  if (tid == 0) best_s = best;
  __syncthreads();
  best = best_s;
  {
    const float* kt = reinterpret_cast<const float*>(d.kf_ptrs[best] + d.kf_sbi_off);
    for (int i = tid; i < n; i += blockDim.x) m.prev[i] = kt[i];
  }
  __syncthreads();
  double R[9];
  const double score = sbi_esm_rotation(sb, m, R);  // CalcSBIRotation(best keyframe's SBI, camera), 6 iterations
  if (tid != 0) return;
  ctl.reloc_kf = best; ctl.reloc_score = score;
  if (!(score < 9e6)) { ctl.frame_mode = 2; return; }  // Reloc2.MaxScore (Relocaliser.cc:37)
  double rot[12], np[12];
  for (int i = 0; i < 9; i++) rot[i] = R[i];
  rot[9] = 0.0; rot[10] = 0.0; rot[11] = 0.0;
  se3_mul(rot, d.kf_pose + 12 * best, np);  // mse3Best = rotation * keyframe pose
  ptam_tracker_state& st = ctl.st;
  for (int i = 0; i < 12; i++) st.se3_cam_from_world[i] = np[i];
  for (int i = 0; i < 6; i++) st.velocity[i] = 0.0;
  st.just_recovered_so_use_coarse = 1;
  ctl.frame_mode = 1;
}

// =============================================================================================
// KeyFrame::MakeKeyFrame_Rest (KeyFrame.cc:61-82) for one stream's current frame:
//   k_rest_score   thread per FAST corner: libCVD fast_corner_score_9 in closed form — the largest
//                  threshold b in [10, 255) at which the pixel is still a FAST-9 corner is
//                  max over the 16 arcs of 9 of min(ring - p) (or of min(p - ring)), minus one —
//                  written to a per-pixel score map (score + 1; 0 = no corner);
//   k_rest_select  one CTA per level, corners in raster order: nonmax_suppression (dropped when an
//                  8-neighbour corner has a strictly greater score), then ShiTomasiScoreAtPoint
//                  (ImageProcess.cc:20-47) for the survivors 10 px inside the image; both lists are
//                  compacted in order.
// =============================================================================================
struct RestDev {
  uint8_t* smap;        // score map, pyramid geometry (LevelDesc::img_off / pitch)
  int2* max_corners;    // per level at LevelDesc::corner_off
  int2* cand_pos;
  double* cand_score;
  int* counts;          // [0..3] n_max per level, [4..7] n_candidates per level
  double min_st_score;
};

__global__ void __launch_bounds__(256) k_rest_score(TrackerDev d, RestDev r, int s, int l) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.ctl[s].n_corners[l]) return;
  const LevelDesc& L = d.g.lev[l];
  int pitch;
  const uint8_t* im = level_image(d, s, l, pitch);
  const int2 c = d.corners[(size_t)s * d.g.corner_stride + L.corner_off + i];
  const uint8_t* pc = im + (size_t)c.y * pitch + c.x;
  const int p = *pc;
  const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
  const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
  int db[16];
#pragma unroll
  for (int j = 0; j < 16; j++) db[j] = (int)pc[dy[j] * pitch + dx[j]] - p;
  int mb = -256, md = -256;
#pragma unroll
  for (int st = 0; st < 16; st++) {
    int lo = 256, hi = -256;
#pragma unroll
    for (int k = 0; k < 9; k++) { const int v = db[(st + k) & 15]; lo = min(lo, v); hi = max(hi, v); }
    mb = max(mb, lo);    // brightest arc: all ring - p >= lo
    md = max(md, -hi);   // darkest arc: all p - ring >= -hi
  }
  int score = max(mb, md) - 1;
  score = max(10, min(score, 254));  // the bisection of fast_corner_score_9 never leaves [barrier, 254]
  r.smap[L.img_off + (size_t)c.y * L.pitch + c.x] = (uint8_t)(score + 1);
}

__global__ void __launch_bounds__(1024) k_rest_select(TrackerDev d, RestDev r, int s) {
  __shared__ int wk[32], wc[32];
  __shared__ int base_k, base_c;
  const int l = blockIdx.x;
  const LevelDesc& L = d.g.lev[l];
  int pitch;
  const uint8_t* im = level_image(d, s, l, pitch);
  const int n = d.ctl[s].n_corners[l];
  const int2* corners = d.corners + (size_t)s * d.g.corner_stride + L.corner_off;
  const uint8_t* smap = r.smap + L.img_off;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { base_k = 0; base_c = 0; }
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += blockDim.x) {
    const int i = i0 + threadIdx.x;
    bool keep = false, cand = false;
    int2 c = make_int2(0, 0);
    double st = 0.0;
    if (i < n) {
      c = corners[i];
      const int own = smap[(size_t)c.y * L.pitch + c.x];
      keep = true;
#pragma unroll
      for (int dy = -1; dy <= 1; dy++)
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) {
          if (!dx && !dy) continue;
          const int x = c.x + dx, y = c.y + dy;
          if (x < 0 || y < 0 || x >= L.w || y >= L.h) continue;
          if ((int)smap[(size_t)y * L.pitch + x] > own) keep = false;
        }
      if (keep && c.x >= 10 && c.y >= 10 && c.x < L.w - 10 && c.y < L.h - 10) {
        double dXX = 0, dYY = 0, dXY = 0;
        for (int y = c.y - 3; y <= c.y + 3; y++)
          for (int x = c.x - 3; x <= c.x + 3; x++) {
            const uint8_t* q = im + (size_t)y * pitch + x;
            const double gx = (double)q[1] - (double)q[-1];
            const double gy = (double)q[pitch] - (double)q[-pitch];
            dXX += gx * gx; dYY += gy * gy; dXY += gx * gy;
          }
        dXX = dXX / (2.0 * 49); dYY = dYY / (2.0 * 49); dXY = dXY / (2.0 * 49);
        st = 0.5 * (dXX + dYY - sqrt((dXX + dYY) * (dXX + dYY) - 4 * (dXX * dYY - dXY * dXY)));
        cand = st > r.min_st_score;
      }
    }
    const unsigned bk = __ballot_sync(kFull, keep), bc = __ballot_sync(kFull, cand);
    if (lane == 0) { wk[warp] = __popc(bk); wc[warp] = __popc(bc); }
    __syncthreads();
    int ok = base_k, oc = base_c;
    for (int q = 0; q < warp; q++) { ok += wk[q]; oc += wc[q]; }
    const unsigned lt = (1u << lane) - 1u;
    if (keep) r.max_corners[L.corner_off + ok + __popc(bk & lt)] = c;
    if (cand) { const int o = oc + __popc(bc & lt); r.cand_pos[L.corner_off + o] = c; r.cand_score[L.corner_off + o] = st; }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tk = 0, tc = 0;
      for (int q = 0; q < (int)(blockDim.x >> 5); q++) { tk += wk[q]; tc += wc[q]; }
      base_k += tk; base_c += tc;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { r.counts[l] = base_k; r.counts[4 + l] = base_c; }
}

// =============================================================================================
// Projection helpers (TrackerData::Project, Tracker.h:70-86)
// =============================================================================================
struct ProjOut { double v3cam[3]; double v2image[2]; bool in_image; bool reached_cam; CamProj q; };

PTAM_DEV void project_point(const TrackerDev& d, const double* pose, const double* world, ProjOut& o) {
  o.in_image = false; o.reached_cam = false;
  se3_apply(pose, world, o.v3cam);
  if (o.v3cam[2] < 0.001) return;
  const double ix = o.v3cam[0] / o.v3cam[2], iy = o.v3cam[1] / o.v3cam[2];
  if (ix * ix + iy * iy > d.cam.largest_radius * d.cam.largest_radius) return;
  o.q = cam_project(d.cam, ix, iy);
  o.reached_cam = true;
  o.v2image[0] = o.q.im[0]; o.v2image[1] = o.q.im[1];
  if (o.q.invalid) return;
  if (o.v2image[0] < 0 || o.v2image[1] < 0 || o.v2image[0] > d.cam.img_w || o.v2image[1] > d.cam.img_h) return;
  o.in_image = true;
}

// =============================================================================================
// k_pvs_select — one CTA (1024 threads) per stream.
// =============================================================================================
__global__ void __launch_bounds__(1024) k_pvs_select(TrackerDev d) {
  __shared__ double pose[12];
  __shared__ int wcnt[kLevels][32];
  __shared__ int running[kLevels];
  __shared__ int seg_src[6], seg_off[6], seg_n[6], seg_dst[6];
  __shared__ int nseg;
  const int s = blockIdx.x + d.s_off;
  StreamCtl& ctl = d.ctl[s];
  // a frame whose relocalisation failed is not tracked at all (frame_mode is written by k_reloc, before this kernel)
  const int fmode = (d.mode == 0 && d.reloc_on) ? ctl.frame_mode : 0;
  const bool skip_frame = fmode == 2;
  const int n = skip_frame ? 0 : d.pt_count[s];
  const int cap = d.p.cap;
  const size_t gb = (size_t)s * cap;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0 && d.mode >= 1) {
    // ReFind / PatchFinder unit entry: the pose is given (the keyframe's se3CfromW); no motion model, tracker state untouched
    for (int i = 0; i < 12; i++) { pose[i] = d.refind_pose[12 * s + i]; ctl.pose[i] = pose[i]; }
    for (int l = 0; l < kLevels; l++) { running[l] = 0; ctl.attempted[l] = 0; ctl.found[l] = 0; }
    ctl.n_cand = 0;
  } else if (threadIdx.x == 0 && fmode != 0) {
    // recovery frame (Tracker.cc:170-178): mnFrame++; the pose is the relocaliser's (mode 1) or stays (mode 2,
    // nothing is tracked); no motion model
    ctl.st.frame++;
    for (int i = 0; i < 12; i++) { ctl.start_pose[i] = ctl.st.se3_cam_from_world[i]; pose[i] = ctl.st.se3_cam_from_world[i]; ctl.pose[i] = pose[i]; }
    for (int l = 0; l < kLevels; l++) { running[l] = 0; ctl.attempted[l] = 0; ctl.found[l] = 0; }
    ctl.n_cand = 0;
  } else if (threadIdx.x == 0) {
    // mnFrame++, PredictPoseWithMotionModel (Tracker.cc:1012-1029)
    ctl.st.frame++;
    double ex[12], np[12];
    for (int i = 0; i < 12; i++) ctl.start_pose[i] = ctl.st.se3_cam_from_world[i];
    double v6[6];
    for (int i = 0; i < 6; i++) v6[i] = ctl.st.velocity[i];
    if (d.prm.use_rotation_estimator) {  // mbUseSBIInit (Tracker.cc:1016-1027): rotation from k_sbi
      v6[3] = ctl.sbi_rot[0]; v6[4] = ctl.sbi_rot[1]; v6[5] = ctl.sbi_rot[2];
      v6[0] = 0.0; v6[1] = 0.0;
    }
    se3_exp(v6, ex);
    se3_mul(ex, ctl.start_pose, np);
    for (int i = 0; i < 12; i++) { pose[i] = np[i]; ctl.pose[i] = np[i]; }
    for (int l = 0; l < kLevels; l++) { running[l] = 0; ctl.attempted[l] = 0; ctl.found[l] = 0; }
    ctl.n_cand = 0;
  }
  __syncthreads();
  int* pvs = d.p.pvs + (size_t)s * kLevels * cap;
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    int lvl = -1;
    if (i < n) {
      const size_t g = gb + i;
      int fl = d.p.flags[g] & (F_TEMPLATE_BAD | F_HAS_TEMPLATE);
      ProjOut o;
      project_point(d, pose, d.p.world + 3 * g, o);
      d.p.v3cam[3 * g] = o.v3cam[0]; d.p.v3cam[3 * g + 1] = o.v3cam[1]; d.p.v3cam[3 * g + 2] = o.v3cam[2];
      if (o.reached_cam) { d.p.v2image[2 * g] = o.v2image[0]; d.p.v2image[2 * g + 1] = o.v2image[1]; }
      if (o.in_image) {
        fl |= F_IN_IMAGE;
        double dv[4];
        cam_derivs(d.cam, o.q, dv);
        for (int k = 0; k < 4; k++) d.p.derivs[4 * g + k] = dv[k];
        // PatchFinder::CalcSearchLevelAndWarpMatrix (PatchFinder.cc:52-84)
        const double ooz = 1.0 / o.v3cam[2];
        double mr[3], md[3], wi[4];
        so3_rotate(pose, d.p.right + 3 * g, mr);
        so3_rotate(pose, d.p.down + 3 * g, md);
        double a0 = (mr[0] - o.v3cam[0] * mr[2] * ooz), a1 = (mr[1] - o.v3cam[1] * mr[2] * ooz);
        wi[0] = (dv[0] * a0 + dv[1] * a1) * ooz; wi[2] = (dv[2] * a0 + dv[3] * a1) * ooz;
        a0 = (md[0] - o.v3cam[0] * md[2] * ooz); a1 = (md[1] - o.v3cam[1] * md[2] * ooz);
        wi[1] = (dv[0] * a0 + dv[1] * a1) * ooz; wi[3] = (dv[2] * a0 + dv[3] * a1) * ooz;
        for (int k = 0; k < 4; k++) d.p.warp_inv[4 * g + k] = wi[k];
        double det = wi[0] * wi[3] - wi[1] * wi[2];
        int sl = 0;
        while (det > 3 && sl < kLevels - 1) { sl++; det *= 0.25; }
        d.p.search_level[g] = sl;
        if (d.mode == 1) {  // ReFind_Common ignores the level verdict; the template is always re-made
          lvl = sl; fl = F_IN_IMAGE | F_IN_PVS;
        } else if (det > 3 || det < 0.25) fl |= F_TEMPLATE_BAD;
        else { lvl = sl; fl |= F_IN_PVS; }
      }
      d.p.flags[g] = fl;
      d.p.level[g] = lvl;
    }
    // ordered compaction into the four per-level PVS lists (map order == identity shuffle)
    unsigned same = 0;
#pragma unroll
    for (int l = 0; l < kLevels; l++) {
      const unsigned b = __ballot_sync(kFull, lvl == l);
      if (lane == 0) wcnt[l][warp] = __popc(b);
      if (lvl == l) same = b;
    }
    __syncthreads();
    if (lvl >= 0) {
      int before = 0;
      for (int wq = 0; wq < warp; wq++) before += wcnt[lvl][wq];
      const int pos = running[lvl] + before + __popc(same & ((1u << lane) - 1));
      pvs[(size_t)lvl * cap + pos] = i;
    }
    __syncthreads();
    if (threadIdx.x < kLevels) {
      int tot = 0;
      for (int wq = 0; wq < (int)(blockDim.x >> 5); wq++) tot += wcnt[threadIdx.x][wq];
      running[threadIdx.x] += tot;
    }
    __syncthreads();
  }
  // ---- selection (Tracker.cc:485-611), identity shuffle --------------------------------------
  if (threadIdx.x == 0 && d.mode >= 1) {  // ReFind / unit entry: every point of the PVS is searched, fine stage only
    int dst = 0, ns = 0;
    for (int l = kLevels - 1; l >= 0; l--) {
      ctl.n_pvs[l] = running[l];
      if (running[l]) { seg_src[ns] = l; seg_off[ns] = 0; seg_n[ns] = running[l]; seg_dst[ns] = dst; dst += running[l]; ns++; }
    }
    ctl.n_coarse = 0; ctl.n_l3 = 0; ctl.n_fine = dst;
    ctl.try_coarse = 0; ctl.coarse_range = 0; ctl.did_coarse = 0;
    nseg = ns;
  } else if (threadIdx.x == 0) {
    int n3 = running[3], n2 = running[2], n1 = running[1], n0 = running[0];
    for (int l = 0; l < kLevels; l++) ctl.n_pvs[l] = running[l];
    unsigned coarse_max = d.prm.coarse_max, coarse_range = d.prm.coarse_range;
    bool try_coarse = true;
    if (d.prm.disable_coarse || ctl.st.msd_scaled_velocity_magnitude < d.prm.coarse_min_velocity || coarse_max == 0) try_coarse = false;
    if (ctl.st.just_recovered_so_use_coarse && !skip_frame) {
      try_coarse = true; coarse_max *= 2; coarse_range *= 2; ctl.st.just_recovered_so_use_coarse = 0;
    }
    int ns = 0, dst = 0;
    int o3 = 0, o2 = 0;  // consumed from the front of the level-3 / level-2 lists
    int ncoarse = 0;
    if (try_coarse && (unsigned)(n3 + n2) > (unsigned)d.prm.coarse_min) {
      unsigned c3 = (unsigned)n3 <= coarse_max ? n3 : coarse_max;
      o3 = c3;
      unsigned have = c3;
      bool keep3 = true;
      unsigned c2 = 0;
      if (have < coarse_max) {
        const unsigned more = coarse_max - have;
        if ((unsigned)n2 <= more) { c2 = n2; keep3 = false; }  // sic: vNextToSearch = avPVS[2] (Tracker.cc:533)
        else c2 = more;
        o2 = c2;
      }
      if (keep3 && c3) { seg_src[ns] = 3; seg_off[ns] = 0; seg_n[ns] = c3; seg_dst[ns] = dst; dst += c3; ns++; }
      if (c2) { seg_src[ns] = 2; seg_off[ns] = 0; seg_n[ns] = c2; seg_dst[ns] = dst; dst += c2; ns++; }
      ncoarse = dst;
      ctl.try_coarse = 1;
    } else ctl.try_coarse = 0;
    ctl.n_coarse = ncoarse;
    ctl.coarse_range = coarse_range;
    ctl.did_coarse = 0;
    // remaining level-3 points
    const int nl3 = n3 - o3;
    if (nl3) { seg_src[ns] = 3; seg_off[ns] = o3; seg_n[ns] = nl3; seg_dst[ns] = dst; dst += nl3; ns++; }
    ctl.n_l3 = nl3;
    int use = d.prm.max_patches_per_frame - dst;
    if (use < 0) use = 0;
    int nfine = 0;
    const int cnt[3] = {n2 - o2, n1, n0};
    const int off[3] = {o2, 0, 0};
    for (int k = 0; k < 3; k++) {
      int take = cnt[k];
      if (take > use - nfine) take = use - nfine;
      if (take > 0) { seg_src[ns] = 2 - k; seg_off[ns] = off[k]; seg_n[ns] = take; seg_dst[ns] = dst; dst += take; ns++; nfine += take; }
    }
    ctl.n_fine = nfine;
    nseg = ns;
  }
  __syncthreads();
  int* iter = d.p.iter_idx + (size_t)s * cap;
  for (int k = 0; k < nseg; k++)
    for (int j = threadIdx.x; j < seg_n[k]; j += blockDim.x)
      iter[seg_dst[k] + j] = pvs[(size_t)seg_src[k] * cap + seg_off[k] + j];
}

// =============================================================================================
// k_search — one warp per iteration-set entry.  stage 0: coarse set; stage 1: level-3 + fine sets.
// =============================================================================================
constexpr int kMaxSSD = 8 * 8 * 500;  // PatchFinder.cc:18-19

PTAM_DEV double level_zero_pos(double p, int l) { return (p + 0.5) * (double)(1 << l) - 0.5; }
PTAM_DEV double level_n_pos(double p, int l) { return (p + 0.5) / (double)(1 << l) - 0.5; }

// k_search_prep — one THREAD per iteration-set entry: everything of SearchForPoints that is scalar per
// point (re-projection with the current pose, Tracker.h:89-94; inverse warp matrix and the template
// cache test, PatchFinder.cc:100-110), so that the warp-per-point k_search does not repeat it 32 times.
__global__ void __launch_bounds__(128) k_search_prep(TrackerDev d, int stage) {
  const int s = blockIdx.y + d.s_off;
  StreamCtl& ctl = d.ctl[s];
  int begin, end;
  if (stage == 0) { begin = 0; end = ctl.n_coarse; }
  else { begin = ctl.n_coarse; end = ctl.n_coarse + ctl.n_l3 + ctl.n_fine; }
  const int k = begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= end) return;
  const int cap = d.p.cap;
  const size_t g = (size_t)s * cap + d.p.iter_idx[(size_t)s * cap + k];
  int fl = d.p.flags[g] & ~F_REFRESH;
  const bool reproject = stage == 1 && (k < ctl.n_coarse + ctl.n_l3 || ctl.did_coarse);
  if (reproject) {
    // ProjectAndDerivs with bFound == false: Project only (Tracker.h:89-94)
    ProjOut o;
    project_point(d, ctl.pose, d.p.world + 3 * g, o);
    d.p.v3cam[3 * g] = o.v3cam[0]; d.p.v3cam[3 * g + 1] = o.v3cam[1]; d.p.v3cam[3 * g + 2] = o.v3cam[2];
    if (o.reached_cam) { d.p.v2image[2 * g] = o.v2image[0]; d.p.v2image[2 * g + 1] = o.v2image[1]; }
    fl = o.in_image ? (fl | F_IN_IMAGE) : (fl & ~F_IN_IMAGE);
  }
  const int sl = d.p.search_level[g];
  // ---- MakeTemplateCoarseCont, the decision part (PatchFinder.cc:98-110)
  const double wi0 = d.p.warp_inv[4 * g], wi1 = d.p.warp_inv[4 * g + 1], wi2 = d.p.warp_inv[4 * g + 2], wi3 = d.p.warp_inv[4 * g + 3];
  const double det = wi0 * wi3 - wi2 * wi1;
  const double idet = 1.0 / det;
  const double sc = (double)(1 << sl);
  double m2[4];
  m2[0] = (wi3 * idet) * sc; m2[3] = (wi0 * idet) * sc;
  m2[2] = (-wi2 * idet) * sc; m2[1] = (-wi1 * idet) * sc;
  bool refresh = !(fl & F_HAS_TEMPLATE);
  if (!refresh) {
    for (int i = 0; !refresh && i < 2; i++) {
      const double d0 = m2[i] - d.p.last_warp[4 * g + i], d1 = m2[2 + i] - d.p.last_warp[4 * g + 2 + i];
      if (d0 * d0 + d1 * d1 > 0.07 * 0.07) refresh = true;
    }
  }
  if (refresh) {
    fl |= F_REFRESH;
    for (int q = 0; q < 4; q++) d.p.m2[4 * g + q] = m2[q];
  }
  d.p.flags[g] = fl;
  // ---- FindPatchCoarse, the scalar part (PatchFinder.cc:160-191): search centre and range in level
  // coordinates (ir() truncation, C++ integer division), the candidates themselves come from the corner mask in k_search
  {
    const unsigned range = d.mode == 1 ? 4u : d.mode == 2 ? d.unit_range : (stage == 0 ? (unsigned)ctl.coarse_range : (ctl.did_coarse ? 5u : 10u));
    const int scale = 1 << sl;
    const int posx = (int)d.p.v2image[2 * g] / scale, posy = (int)d.p.v2image[2 * g + 1] / scale;
    const unsigned r = (range + scale - 1) / scale;
    d.p.geo[2 * g] = make_int4(posx, posy, (int)r, 0);
  }
}

__global__ void __launch_bounds__(128, 12) k_search(TrackerDev d, int stage) {
  __shared__ __align__(8) uint8_t stmpl[4][64];
  __shared__ int2 squeue[4][64];
  const int s = blockIdx.y + d.s_off;
  StreamCtl& ctl = d.ctl[s];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int item = blockIdx.x * 4 + warp;
  int begin, end;
  if (stage == 0) { begin = 0; end = ctl.n_coarse; }
  else { begin = ctl.n_coarse; end = ctl.n_coarse + ctl.n_l3 + ctl.n_fine; }
  const int k = begin + item;
  if (k >= end) return;
  const int cap = d.p.cap;
  const int idx = d.p.iter_idx[(size_t)s * cap + k];
  const size_t g = (size_t)s * cap + idx;
  int fl = d.p.flags[g];
  unsigned range;
  int subpix_its;
  if (stage == 0) { range = ctl.coarse_range; subpix_its = d.prm.coarse_subpix_its; }
  else {
    range = ctl.did_coarse ? 5 : 10;
    subpix_its = k < ctl.n_coarse + ctl.n_l3 ? 8 : 0;
  }
  const bool refind = d.mode == 1;
  const double v2image[2] = {d.p.v2image[2 * g], d.p.v2image[2 * g + 1]};  // re-projected by k_search_prep
  const int sl = d.p.search_level[g];
  if (refind) subpix_its = sl > 0 ? 8 : 0;  // MapMaker.cc:1003-1014
  if (d.mode == 2) subpix_its = d.unit_subpix_its;
  const bool refresh = (fl & F_REFRESH) != 0;
  fl &= ~F_REFRESH;
  // lane owns template pixels (row = lane/4, cols 2*(lane%4), +1)
  const int trow = lane >> 2, tcol = (lane & 3) * 2;
  int t0, t1, tsum, tsumsq;
  if (refresh) {
    const int kf = d.p.src_kf[g], slv = d.p.src_level[g];
    const LevelDesc& SL = d.g.lev[slv];
    const uint8_t* src = d.kf_ptrs[kf] + SL.img_off;
    const int2 c = d.p.center[g];
    const double m2[4] = {d.p.m2[4 * g], d.p.m2[4 * g + 1], d.p.m2[4 * g + 2], d.p.m2[4 * g + 3]};
    // CVD::transform: p = inOrig + M (out - outOrig), accumulated across/down like libCVD
    const double ax = m2[0], ay = m2[2];
    const double crx = m2[1] - 8 * ax, cry = m2[3] - 8 * ay;
    double px = (double)c.x - (m2[0] * 4.0 + m2[1] * 4.0), py = (double)c.y - (m2[2] * 4.0 + m2[3] * 4.0);
    for (int i = 0; i < trow; i++) {
      for (int j = 0; j < 8; j++) { px += ax; py += ay; }
      px += crx; py += cry;
    }
    for (int j = 0; j < tcol; j++) { px += ax; py += ay; }
    const double xb = (double)(SL.w - 1), yb = (double)(SL.h - 1);
    int outside = 0;
    int tv[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      if (0 <= px && 0 <= py && px < xb && py < yb) {
        double x = px, y = py;
        const int lx = (int)x, ly = (int)y;
        x -= lx; y -= ly;
        const uint8_t* r0 = src + (size_t)ly * SL.pitch + lx;
        const uint8_t* r1 = r0 + SL.pitch;
        const double v = (1 - y) * ((1 - x) * (double)r0[0] + x * (double)r0[1]) + y * ((1 - x) * (double)r1[0] + x * (double)r1[1]);
        tv[q] = (int)(uint8_t)(int)v;
      } else { tv[q] = 0; outside++; }
      px += ax; py += ay;
    }
    t0 = tv[0]; t1 = tv[1];
    outside = warp_sum_int(outside);
    tsum = warp_sum_int(t0 + t1);
    tsumsq = warp_sum_int(t0 * t0 + t1 * t1);
    fl = outside ? (fl | F_TEMPLATE_BAD) : (fl & ~F_TEMPLATE_BAD);
    fl |= F_HAS_TEMPLATE;
    *reinterpret_cast<uchar2*>(d.p.tmpl + 64 * g + 2 * lane) = make_uchar2((uint8_t)t0, (uint8_t)t1);
    if (lane == 0) {
      d.p.tsum[g] = tsum; d.p.tsumsq[g] = tsumsq;
      for (int q = 0; q < 4; q++) d.p.last_warp[4 * g + q] = m2[q];
    }
  } else {
    const uchar2 t = *reinterpret_cast<const uchar2*>(d.p.tmpl + 64 * g + 2 * lane);
    t0 = t.x; t1 = t.y;
    tsum = d.p.tsum[g]; tsumsq = d.p.tsumsq[g];
  }
  if (fl & F_TEMPLATE_BAD) {  // Tracker.cc:873-876
    fl &= ~(F_IN_IMAGE | F_FOUND);
    if (lane == 0) d.p.flags[g] = fl;
    return;
  }
  if (lane == 0) atomicAdd(&ctl.attempted[sl], 1);
  stmpl[warp][2 * lane] = (uint8_t)t0; stmpl[warp][2 * lane + 1] = (uint8_t)t1;
  __syncwarp();
  // ---- FindPatchCoarse (PatchFinder.cc:160-211) ------------------------------------------------
  bool found = false;
  double coarse[2] = {0, 0};
  {
    const int4 g0 = d.p.geo[2 * g];
    const int posx = g0.x, posy = g0.y;
    const unsigned r = (unsigned)g0.z;
    const LevelDesc& L = d.g.lev[sl];
    int pitch;
    const uint8_t* im = level_image(d, s, sl, pitch);
    // candidate corners: inside the disc of radius r around (posx, posy) and 4 px off the border (PatchFinder.cc:193-196,
    // ImageProcess.cc:134).  They are read straight from the corner MASK of the level: lane -> one row of the
    // window, the one to three mask words that cover [posx - r, posx + r]; no corner list, no row LUT on this path.
    const int x_lo = max(posx - (int)r, 4), x_hi = min(posx + (int)r, L.w - 5);
    const int y_lo = max(posy - (int)r, 4), y_hi = min(posy + (int)r, L.h - 5);
    if (x_lo <= x_hi && y_lo <= y_hi) {
      const uint32_t* mask = d.mask + (size_t)s * d.g.mask_stride + L.mask_off;
      // Candidates are queued per warp; ZMSSD then runs four candidates at a time, eight lanes per candidate, one
      // 8-pixel window row per lane: three aligned 32-bit loads + byte_perm, six dp4a, group reduction.  The winner
      // is the minimum of (ssd, raster position), i.e. the first minimum in raster order (PatchFinder.cc:198).
      const int sub = lane & 7, grp = lane >> 3;
      const unsigned tw0 = *reinterpret_cast<const unsigned*>(&stmpl[warp][8 * sub]);
      const unsigned tw1 = *reinterpret_cast<const unsigned*>(&stmpl[warp][8 * sub + 4]);
      int best_ssd = kMaxSSD + 1, best_idx = 0x7fffffff;
      auto process = [&](int n) {
        for (int r0 = 0; r0 < n; r0 += 4) {
          const int kq = r0 + grp;
          const bool valid = kq < n;
          const int ek = squeue[warp][valid ? kq : 0].x;  // y << 16 | x: the raster position
          const int cx = ek & 0xffff, cy = ek >> 16;
          const uint8_t* ip = im + (size_t)(cy - 4 + sub) * pitch + (cx - 4);
          const unsigned a = (unsigned)(reinterpret_cast<uintptr_t>(ip) & 3);
          const unsigned* wp = reinterpret_cast<const unsigned*>(ip - a);
          const unsigned w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = a ? __ldg(wp + 2) : 0u;
          const unsigned sel = 0x3210u + 0x1111u * a;
          const unsigned v0 = __byte_perm(w0, w1, sel), v1 = __byte_perm(w1, w2, sel);
          int isum = (int)__dp4a(v0, 0x01010101u, __dp4a(v1, 0x01010101u, 0u));
          int isq = (int)__dp4a(v0, v0, __dp4a(v1, v1, 0u));
          int cross = (int)__dp4a(v0, tw0, __dp4a(v1, tw1, 0u));
#pragma unroll
          for (int o = 4; o > 0; o >>= 1) {  // sum over the 8 lanes (window rows) of the group
            isum += __shfl_xor_sync(kFull, isum, o);
            isq += __shfl_xor_sync(kFull, isq, o);
            cross += __shfl_xor_sync(kFull, cross, o);
          }
          const int SA = tsum, SB = isum;
          const int ssd = ((2 * SA * SB - SA * SA - SB * SB) / 64 + isq + tsumsq - 2 * cross);  // C++ truncating division
          if (valid && (ssd < best_ssd || (ssd == best_ssd && ek < best_idx))) { best_ssd = ssd; best_idx = ek; }
        }
        __syncwarp();
      };
      int qn = 0, n_eval = 0;
      const unsigned lt = (1u << lane) - 1u;
      const int w_lo = x_lo >> 5, w_hi = x_hi >> 5;
      for (int yb = y_lo; yb <= y_hi; yb += 32) {
        const int y = yb + lane;
        const int ddy = posy - y;
        for (int wi = w_lo; wi <= w_hi; wi++) {
          unsigned m = 0;
          if (y <= y_hi) {
            m = __ldg(mask + (size_t)y * L.nwords + wi);
            const int b_lo = max(x_lo - 32 * wi, 0), b_hi = min(x_hi - 32 * wi, 31);  // bits of this word inside [x_lo, x_hi]
            m &= (0xffffffffu << b_lo) & (0xffffffffu >> (31 - b_hi));
          }
          while (__any_sync(kFull, m != 0)) {
            bool pass = false;
            int cx = 0;
            if (m) {
              const int b = __ffs(m) - 1;
              m &= m - 1;
              cx = 32 * wi + b;
              const int ddx = posx - cx;
              pass = !((unsigned)(ddx * ddx + ddy * ddy) > r * r);
            }
            const unsigned bal = __ballot_sync(kFull, pass);
            if (pass) squeue[warp][qn + __popc(bal & lt)] = make_int2(cx | (y << 16), 0);
            qn += __popc(bal);
            n_eval += __popc(bal);
            if (qn > 32) { __syncwarp(); process(qn); qn = 0; }
          }
        }
      }
      __syncwarp();
      process(qn);
      if (lane == 0 && n_eval) atomicAdd(&ctl.n_cand, n_eval);
      {
        const int bs = __reduce_min_sync(kFull, best_ssd);
        best_idx = __reduce_min_sync(kFull, best_ssd == bs ? best_idx : 0x7fffffff);
        best_ssd = bs;
      }
      int bx = 0, by = 0;
      if (best_ssd < kMaxSSD) { bx = best_idx & 0xffff; by = best_idx >> 16; }
      if (best_ssd < kMaxSSD) {
        coarse[0] = level_zero_pos((double)bx, sl);
        coarse[1] = level_zero_pos((double)by, sl);
        found = true;
      }
    }
  }
  fl |= F_SEARCHED;
  if (!found) {
    fl &= ~F_FOUND;
    if (lane == 0) d.p.flags[g] = fl;
    return;
  }
  fl |= F_FOUND;
  double v2found[2] = {coarse[0], coarse[1]};
  if (subpix_its > 0) {
    fl |= F_SUBPIX;
    // ---- MakeSubPixTemplate (PatchFinder.cc:219-240); the template is already in stmpl[warp] ----
    float jx[2] = {0.f, 0.f}, jy[2] = {0.f, 0.f};
    int tq[2] = {0, 0};
    double h00 = 0, h01 = 0, h02 = 0, h11 = 0, h12 = 0, h22 = 0;
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int pi = lane + 32 * q;
      if (pi < 36) {
        const int y = pi / 6 + 1, x = pi % 6 + 1;
        const uint8_t* T = stmpl[warp];
        const double gx = 0.5 * (double)((int)T[8 * y + x + 1] - (int)T[8 * y + x - 1]);
        const double gy = 0.5 * (double)((int)T[8 * (y + 1) + x] - (int)T[8 * (y - 1) + x]);
        jx[q] = (float)gx; jy[q] = (float)gy; tq[q] = T[8 * y + x];
        h00 += gx * gx; h01 += gx * gy; h02 += gx; h11 += gy * gy; h12 += gy; h22 += 1.0;
      }
    }
    double H[9], Hi[9];
    H[0] = warp_sum(h00); H[1] = H[3] = warp_sum(h01); H[2] = H[6] = warp_sum(h02);
    H[4] = warp_sum(h11); H[5] = H[7] = warp_sum(h12); H[8] = warp_sum(h22);
    ldlt_inverse<3>(H, Hi);
    double sp[2] = {coarse[0], coarse[1]};
    double mean_diff = 0.0;
    // ---- IterateSubPixToConvergence (PatchFinder.cc:250-318) -----------------------------------
    const LevelDesc& L = d.g.lev[sl];
    int pitch;
    const uint8_t* im = level_image(d, s, sl, pitch);
    bool converged = false;
    for (int it = 0; it < subpix_its; it++) {
      const double cx = level_n_pos(sp[0], sl), cy = level_n_pos(sp[1], sl);
      const int rx = (int)(cx > 0.0 ? cx + 0.5 : cx - 0.5), ry = (int)(cy > 0.0 ? cy + 0.5 : cy - 0.5);
      if (!(rx >= 5 && ry >= 5 && rx < L.w - 5 && ry < L.h - 5)) break;  // went off the image
      const double bx = cx - 4, by = cy - 4;
      const double dX = bx - floor(bx), dY = by - floor(by);
      const float fTL = (float)((1.0 - dX) * (1.0 - dY));
      const float fTR = (float)((dX) * (1.0 - dY));
      const float fBL = (float)((1.0 - dX) * (dY));
      const float fBR = (float)((dX) * (dY));
      const int ibx = (int)bx, iby = (int)by;
      double a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const int pi = lane + 32 * q;
        if (pi < 36) {
          const int y = pi / 6 + 1, x = pi % 6 + 1;
          const uint8_t* tl = im + (size_t)(iby + y) * pitch + ibx + x;
          const float fp = fTL * (float)tl[0] + fTR * (float)tl[1] + fBL * (float)tl[pitch] + fBR * (float)tl[pitch + 1];
          const double diff = (double)(fp - (float)tq[q]) + mean_diff;
          a0 += diff * (double)jx[q]; a1 += diff * (double)jy[q]; a2 += diff;
        }
      }
      a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
      const double u0 = Hi[0] * a0 + Hi[1] * a1 + Hi[2] * a2;
      const double u1 = Hi[3] * a0 + Hi[4] * a1 + Hi[5] * a2;
      const double u2 = Hi[6] * a0 + Hi[7] * a1 + Hi[8] * a2;
      sp[0] -= u0 * (double)(1 << sl);
      sp[1] -= u1 * (double)(1 << sl);
      mean_diff -= u2;
      const double uu = u0 * u0 + u1 * u1;
      if (uu < 0.03 * 0.03) { converged = true; break; }
    }
    if (!converged && !refind) {  // Tracker.cc:898-903 (ReFind_Common does not look at the result)
      fl &= ~F_FOUND;
      if (lane == 0) d.p.flags[g] = fl;
      return;
    }
    v2found[0] = sp[0]; v2found[1] = sp[1];
  } else {
    fl &= ~F_SUBPIX;
  }
  if (lane == 0) {
    atomicAdd(&ctl.found[sl], 1);
    d.p.flags[g] = fl;
    d.p.v2found[2 * g] = v2found[0]; d.p.v2found[2 * g + 1] = v2found[1];
    d.p.sqrt_inv_noise[g] = 1.0 / (double)(1 << sl);
  }
}

// =============================================================================================
// k_pose — one CTA per stream.  stage 0: coarse-stage Gauss-Newton (Tracker.cc:552-568) when enough
// coarse points were found; stage 1: the ten fine iterations (Tracker.cc:614-643), measurement
// export statistics, UpdateMotionModel, AssessTrackingQuality.
// =============================================================================================
#ifndef PTAM_POSE_THREADS
#define PTAM_POSE_THREADS 256
#endif
constexpr int kPoseThreads = PTAM_POSE_THREADS;  // two CTAs (streams) resident per SM

// exact k-th smallest (0-based) of n non-negative doubles: MSB-first radix select on the IEEE bit
// patterns (order-isomorphic to the values), 8 bits per pass; the 256-bin histogram is scanned by
// warp 0 (8 bins per lane + shuffle scan), so a pass costs two barriers and no serial loop.
// Squared errors of one frame share their top bits, so the histogram updates of the FIRST pass are aggregated per
// warp (match.any on the digit: one shared-memory atomic per distinct digit and warp instead of one per element on
// the same bin; the later digits are spread, where match.any costs more than it saves), and the passes stop as soon as
// the bin of the k-th element holds ONE element: a last sweep fetches that element (usually after 3-4 of the 8 passes; ties run all 8).
// keys are read through `at(i)`.  All threads of the block must call it.
template <class At>
PTAM_DEV double block_select_kth(At at, int n, int kth, int* hist /*2 x 256*/, unsigned long long* sh_prefix, int* sh_k) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { *sh_prefix = 0ull; sh_k[0] = kth; sh_k[1] = 0; }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (int pass = 0; pass < 8; pass++) {
    const int shift = 56 - 8 * pass;
    int* h = hist + 256 * (pass & 1);
    int* hnext = hist + 256 * ((pass + 1) & 1);
    const unsigned long long prefix = *sh_prefix;
    for (int base = 0; base < n; base += blockDim.x) {
      const int i = base + threadIdx.x;
      unsigned digit = 256u + lane;  // not a bin: lanes without an element match nobody
      if (i < n) {
        const unsigned long long key = (unsigned long long)__double_as_longlong(at(i));
        if (pass == 0 || (key >> (shift + 8)) == (prefix >> (shift + 8))) digit = (unsigned)(key >> shift) & 255u;
      }
      if (pass == 0) {  // sign + top exponent bits: one or two bins for the whole frame -> one atomic per warp and bin
        const unsigned peers = __match_any_sync(kFull, digit);
        if (digit < 256u && lane == __ffs(peers) - 1) atomicAdd(&h[digit], __popc(peers));
      } else if (digit < 256u) {  // later digits are spread: match.any walks the distinct values, plain atomics are cheaper
        atomicAdd(&h[digit], 1);
      }
    }
    __syncthreads();
    if (warp == 0) {
      int c[8], tot = 0;
#pragma unroll
      for (int q = 0; q < 8; q++) { c[q] = h[8 * lane + q]; tot += c[q]; }
      int inc = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += v;
      }
      const int kk = sh_k[0];
      __syncwarp();  // every lane has read sh_k[0] before the one owner rewrites it
      const int excl = inc - tot;
      if (kk >= excl && kk < inc) {  // exactly one lane
        int r = kk - excl, q = 0;
        while (q < 7 && r >= c[q]) { r -= c[q]; q++; }
        sh_k[0] = r;
        sh_k[1] = c[q] == 1 ? shift : 0;  // the k-th element is alone in its bin: bits below `shift` come from the element itself
        *sh_prefix = prefix | ((unsigned long long)(8 * lane + q) << shift);
      }
    } else {
      for (int i = threadIdx.x - 32; i < 256; i += blockDim.x - 32) hnext[i] = 0;  // for the next pass
    }
    __syncthreads();
    const int uniq_shift = sh_k[1];
    if (uniq_shift) {  // block-uniform
      const unsigned long long pfx = *sh_prefix;
      __syncthreads();  // everybody has read the prefix before its owner overwrites it
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long key = (unsigned long long)__double_as_longlong(at(i));
        if ((key >> uniq_shift) == (pfx >> uniq_shift)) *sh_prefix = key;  // one element
      }
      __syncthreads();
      break;
    }
  }
  return __longlong_as_double((long long)*sh_prefix);
}

// Sum v[i] over the 32 lanes for 32 values at once: after the call lane i holds the total of v[i]
// in v[0].  Halving exchange: 16 + 8 + 4 + 2 + 1 = 31 shuffles instead of 32 x 5.
PTAM_DEV void warp_transpose_sum32(double (&v)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < off; k++) {
      const double send = up ? v[k] : v[k + off];
      const double keep = up ? v[k + off] : v[k];
      v[k] = keep + __shfl_xor_sync(kFull, send, off);
    }
  }
}

// Per-point working set of the Gauss-Newton loop.  Found sets of up to kPoseSmemPts points (the
// tracker's MaxPatchesPerFrame default is 1000) keep it in shared memory for all ten iterations
// (96 B per point, SoA by found index): v2Found, sqrt-inverse-noise, v2Image, v3Cam, the 2x2 camera
// derivatives.  The 2x6 Jacobian is not stored: CalcJacobian (Tracker.h:125-136) is re-evaluated from
// v3Cam / derivatives, which only change on the non-linear iterations, so LinearUpdate sees exactly
// the Jacobian of the last non-linear iteration as in the reference.  Larger sets fall back to the
// same code over the global arrays.
constexpr int kPoseSmemPts = 1024;
constexpr int kPoseSmemBytes = 12 * kPoseSmemPts * (int)sizeof(double);

struct PoseStoreSmem {
  double* base;  // [12][kPoseSmemPts]
  PTAM_DEV double& found(int i, int k) const { return base[(0 + k) * kPoseSmemPts + i]; }
  PTAM_DEV double& sn(int i) const { return base[2 * kPoseSmemPts + i]; }
  PTAM_DEV double& image(int i, int k) const { return base[(3 + k) * kPoseSmemPts + i]; }
  PTAM_DEV double& v3(int i, int k) const { return base[(5 + k) * kPoseSmemPts + i]; }
  PTAM_DEV double& dv(int i, int k) const { return base[(8 + k) * kPoseSmemPts + i]; }
};
struct PoseStoreGlobal {
  const PointArrays* p; size_t gb; const int* fidx;
  PTAM_DEV double& found(int i, int k) const { return p->v2found[2 * (gb + fidx[i]) + k]; }
  PTAM_DEV double& sn(int i) const { return p->sqrt_inv_noise[gb + fidx[i]]; }
  PTAM_DEV double& image(int i, int k) const { return p->v2image[2 * (gb + fidx[i]) + k]; }
  PTAM_DEV double& v3(int i, int k) const { return p->v3cam[3 * (gb + fidx[i]) + k]; }
  PTAM_DEV double& dv(int i, int k) const { return p->derivs[4 * (gb + fidx[i]) + k]; }
};

// TrackerData::CalcJacobian (Tracker.h:125-136): rows of the 2x6 Jacobian w.r.t. the SE3 generators
PTAM_DEV void calc_jacobian(double X, double Y, double Z, double dv0, double dv1, double dv2, double dv3, double* J0, double* J1) {
  const double ooz = 1.0 / Z;
  const double gx[6] = {1, 0, 0, 0, Z, -Y}, gy[6] = {0, 1, 0, -Z, 0, X}, gz[6] = {0, 0, 1, Y, -X, 0};
#pragma unroll
  for (int m = 0; m < 6; m++) {
    const double a0 = (gx[m] - X * gz[m] * ooz) * ooz;
    const double a1 = (gy[m] - Y * gz[m] * ooz) * ooz;
    J0[m] = dv0 * a0 + dv1 * a1;
    J1[m] = dv2 * a0 + dv3 * a1;
  }
}

#ifdef PTAM_POSE_CLOCKS
__device__ long long g_pose_clk[32];
#define PCLK(k) do { if (threadIdx.x == 0 && blockIdx.x == 0) { const long long t_ = clock64(); g_pose_clk[(k) + 16 * (stage == 0)] += t_ - pclk_last; pclk_last = t_; } } while (0)
#else
#define PCLK(k) do { } while (0)
#endif

struct PoseShared {
  double s_e2[kPoseSmemPts];
  double pose[12];
  double red[kPoseThreads / 32][27];
  double mu_s[6];
  int hist[512];
  unsigned long long sh_prefix;
  int sh_k[2];
};

// The ten iterations of Tracker.cc:552-568 (stage 0) / 614-643 (stage 1) over the found set.
template <class Store>
PTAM_DEV void pose_iterations(const TrackerDev& d, const Store& st, PoseShared& sh, int stage, int nf, size_t gb,
                              const int* fidx, double* e2_global) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int est = d.prm.mestimator;
  double* pose = sh.pose;
  double last_mu[6] = {0, 0, 0, 0, 0, 0};
  const bool e2_smem = nf <= kPoseSmemPts;
  const bool unit = d.mode == 2;  // ptam_pose_update: ONE CalcPoseUpdate at the projections of the patch search, not applied
#ifdef PTAM_POSE_CLOCKS
  long long pclk_last = clock64();
#endif
  for (int it = 0; it < (unit ? 1 : 10); it++) {
    const bool nonlin = stage == 0 || it == 0 || it == 4 || it == 9;
    const double override_sigma = unit ? d.unit_override_sigma : (it > 5 ? (stage == 0 ? 1.0 : 16.0) : 0.0);
    const bool mark = unit ? d.unit_mark != 0 : (stage == 1 && it == 9);
    // per-point update: reprojection / linear update, scaled error
    for (int i = threadIdx.x; i < nf; i += blockDim.x) {
      double v2i[2] = {st.image(i, 0), st.image(i, 1)};
      if (it != 0) {
        if (nonlin) {  // ProjectAndDerivs (Tracker.h:89-94)
          ProjOut o;
          const size_t g = gb + fidx[i];
          project_point(d, pose, d.p.world + 3 * g, o);
          {  // Project() leaves bInImage behind (Tracker.h:72,85)
            const int fl = d.p.flags[g];
            const int nfl = o.in_image ? (fl | F_IN_IMAGE) : (fl & ~F_IN_IMAGE);
            if (nfl != fl) d.p.flags[g] = nfl;
          }
          st.v3(i, 0) = o.v3cam[0]; st.v3(i, 1) = o.v3cam[1]; st.v3(i, 2) = o.v3cam[2];
          if (o.reached_cam) {
            v2i[0] = o.v2image[0]; v2i[1] = o.v2image[1];
            double dv[4];
            cam_derivs(d.cam, o.q, dv);
            for (int q = 0; q < 4; q++) st.dv(i, q) = dv[q];
          }
        } else {  // LinearUpdate (Tracker.h:139-142) with the Jacobian of the last non-linear iteration
          double J0[6], J1[6];
          calc_jacobian(st.v3(i, 0), st.v3(i, 1), st.v3(i, 2), st.dv(i, 0), st.dv(i, 1), st.dv(i, 2), st.dv(i, 3), J0, J1);
          double a0 = 0, a1 = 0;
          for (int q = 0; q < 6; q++) { a0 += J0[q] * last_mu[q]; a1 += J1[q] * last_mu[q]; }
          v2i[0] += a0; v2i[1] += a1;
        }
        st.image(i, 0) = v2i[0]; st.image(i, 1) = v2i[1];
      }
      const double sn = st.sn(i);
      const double e0 = sn * (st.found(i, 0) - v2i[0]), e1 = sn * (st.found(i, 1) - v2i[1]);
      const double ee = e0 * e0 + e1 * e1;
      if (e2_smem) sh.s_e2[i] = ee; else e2_global[fidx[i]] = ee;
    }
    __syncthreads();
    PCLK(nonlin ? 0 : 1);
    double mu[6] = {0, 0, 0, 0, 0, 0};
    if (nf > 0) {
      double sigma2;
      if (override_sigma > 0) sigma2 = override_sigma;
      else {
        double med;
        if (e2_smem) med = block_select_kth([&](int i) { return sh.s_e2[i]; }, nf, nf / 2, sh.hist, &sh.sh_prefix, sh.sh_k);
        else med = block_select_kth([&](int i) { return e2_global[fidx[i]]; }, nf, nf / 2, sh.hist, &sh.sh_prefix, sh.sh_k);
        sigma2 = mest_sigma_from_median(med, nf, est);
      }
      PCLK(2);
      // weighted normal equations (TooN WLS<6>::add_mJ twice per point)
      double acc[32];
#pragma unroll
      for (int q = 0; q < 32; q++) acc[q] = 0;
      for (int i = threadIdx.x; i < nf; i += blockDim.x) {
        const double sn = st.sn(i);
        const double e0 = sn * (st.found(i, 0) - st.image(i, 0)), e1 = sn * (st.found(i, 1) - st.image(i, 1));
        const double es = e0 * e0 + e1 * e1;
        const double wgt = mest_weight(es, sigma2, est);
        if (wgt == 0.0) { if (mark) d.p.outliers[gb + fidx[i]]++; continue; }
        if (mark) d.p.inliers[gb + fidx[i]]++;
        double J0[6], J1[6];
        calc_jacobian(st.v3(i, 0), st.v3(i, 1), st.v3(i, 2), st.dv(i, 0), st.dv(i, 1), st.dv(i, 2), st.dv(i, 3), J0, J1);
#pragma unroll
        for (int r = 0; r < 2; r++) {
          double Jr[6], Jw[6];
#pragma unroll
          for (int q = 0; q < 6; q++) { Jr[q] = sn * (r ? J1[q] : J0[q]); Jw[q] = Jr[q] * wgt; }
          const double er = r ? e1 : e0;
          int c = 0;
#pragma unroll
          for (int a = 0; a < 6; a++) {
#pragma unroll
            for (int b = 0; b <= a; b++) { acc[c] = fma(Jw[a], Jr[b], acc[c]); c++; }
          }
#pragma unroll
          for (int a = 0; a < 6; a++) acc[21 + a] = fma(er, Jw[a], acc[21 + a]);
        }
      }
      PCLK(3);
      warp_transpose_sum32(acc);
      if (lane < 27) sh.red[warp][lane] = acc[0];
      __syncthreads();
      PCLK(4);
      if (warp == 0) {  // cross-warp sums, the 6x6 solve and the pose update without further block barriers
        if (lane < 27) {
          double t = 0;
          for (int wq = 0; wq < kPoseThreads / 32; wq++) t += sh.red[wq][lane];
          sh.red[0][lane] = t;
        }
        __syncwarp();
        if (lane == 0) {
          double tot[27];
          for (int q = 0; q < 27; q++) tot[q] = sh.red[0][q];
          double C[36], b[6], x[6];
          int c = 0;
          for (int a = 0; a < 6; a++)
            for (int bb = 0; bb <= a; bb++) { C[6 * a + bb] = tot[c]; C[6 * bb + a] = tot[c]; c++; }
          for (int a = 0; a < 6; a++) { C[7 * a] += 100.0; b[a] = tot[21 + a]; }  // add_prior(100)
          ldlt_factor<6>(C);
          ldlt_backsub<6>(C, b, x);
          for (int a = 0; a < 6; a++) sh.mu_s[a] = x[a];
          if (!unit) {  // mse3CamFromWorld = SE3<>::exp(v6Update) * mse3CamFromWorld
            double ex[12], np[12];
            se3_exp(x, ex);
            se3_mul(ex, pose, np);
            for (int i = 0; i < 12; i++) pose[i] = np[i];
          }
        }
      }
      __syncthreads();
      PCLK(5);
      for (int a = 0; a < 6; a++) mu[a] = sh.mu_s[a];
    }
    if (unit) {
      if (threadIdx.x < 6) d.unit_mu[6 * blockIdx.x + threadIdx.x] = mu[threadIdx.x];
      break;
    }
    for (int a = 0; a < 6; a++) last_mu[a] = mu[a];
    PCLK(6);
  }
}

__global__ void __launch_bounds__(kPoseThreads, 2) k_pose(TrackerDev d, int stage) {
  extern __shared__ __align__(16) unsigned char pose_dyn[];  // kPoseSmemBytes: the per-point working set
  __shared__ PoseShared sh;
  __shared__ double red2[kPoseThreads / 32][2];
  __shared__ int sh_cnt[kPoseThreads / 32];
  __shared__ int nfound_s;
  double* pose = sh.pose;
  const int s = blockIdx.x + d.s_off;
  StreamCtl& ctl = d.ctl[s];
  const int cap = d.p.cap;
  const size_t gb = (size_t)s * cap;
  const int* iter = d.p.iter_idx + gb;
  double* e2 = d.p.e2 + gb;
  int* fidx = d.p.pvs + (size_t)s * kLevels * cap;  // PVS lists are dead after selection: reuse as found-index scratch
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_set = stage == 0 ? ctl.n_coarse : ctl.n_coarse + ctl.n_l3 + ctl.n_fine;
  const int fmode = (d.mode == 0 && d.reloc_on) ? ctl.frame_mode : 0;
  if (stage == 0 && threadIdx.x < kLevels) ctl.res_n_corners[threadIdx.x] = d.ncorn[s * kLevels + threadIdx.x];
  if (fmode == 2) return;  // relocalisation failed: the reference does nothing else this frame
#ifdef PTAM_POSE_CLOCKS
  long long pclk_last = clock64();
#endif

  // compact the found entries (order preserved) once: the found set does not change during GN
  if (threadIdx.x == 0) nfound_s = 0;
  for (int i = threadIdx.x; i < 12; i += blockDim.x) pose[i] = ctl.pose[i];
  __syncthreads();
  for (int base = 0; base < n_set; base += blockDim.x) {
    const int k = base + threadIdx.x;
    bool f = false;
    if (k < n_set) f = (d.p.flags[gb + iter[k]] & F_FOUND) != 0;
    const unsigned b = __ballot_sync(kFull, f);
    if (lane == 0) sh_cnt[warp] = __popc(b);
    __syncthreads();
    int before = nfound_s;
    for (int wq = 0; wq < warp; wq++) before += sh_cnt[wq];
    if (f) fidx[before + __popc(b & ((1u << lane) - 1))] = iter[k];
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int wq = 0; wq < (int)(blockDim.x >> 5); wq++) t += sh_cnt[wq]; nfound_s += t; }
    __syncthreads();
  }
  const int nf = nfound_s;
  bool run = true;
  if (stage == 0) {
    run = ctl.try_coarse && nf >= d.prm.coarse_min;
    if (threadIdx.x == 0) ctl.did_coarse = run ? 1 : 0;
  }
  const bool in_smem = d.pose_ws_smem && nf <= kPoseSmemPts;
  const PoseStoreSmem ss{reinterpret_cast<double*>(pose_dyn)};
  const PoseStoreGlobal sg{&d.p, gb, fidx};
  if (run) {
    if (in_smem) {
      for (int i = threadIdx.x; i < nf; i += blockDim.x) {  // gather the working set once
        ss.found(i, 0) = sg.found(i, 0); ss.found(i, 1) = sg.found(i, 1); ss.sn(i) = sg.sn(i);
        ss.image(i, 0) = sg.image(i, 0); ss.image(i, 1) = sg.image(i, 1);
        for (int q = 0; q < 3; q++) ss.v3(i, q) = sg.v3(i, q);
        for (int q = 0; q < 4; q++) ss.dv(i, q) = sg.dv(i, q);
      }
      __syncthreads();
      PCLK(8);
      pose_iterations(d, ss, sh, stage, nf, gb, fidx, e2);
      PCLK(9);
      for (int i = threadIdx.x; i < nf; i += blockDim.x) {  // what later stages and the getters read
        sg.image(i, 0) = ss.image(i, 0); sg.image(i, 1) = ss.image(i, 1);
        for (int q = 0; q < 3; q++) sg.v3(i, q) = ss.v3(i, q);
        for (int q = 0; q < 4; q++) sg.dv(i, q) = ss.dv(i, q);
      }
    } else {
      pose_iterations(d, sg, sh, stage, nf, gb, fidx, e2);
    }
  }
  if (d.mode == 2) {  // unit entry: nothing of the tracker's state moves
    if (threadIdx.x == 0) d.unit_nfound[s] = nf;
    if (nf == 0 && threadIdx.x < 6) d.unit_mu[6 * s + threadIdx.x] = 0.0;
    return;
  }
  if (threadIdx.x < 12) ctl.pose[threadIdx.x] = pose[threadIdx.x];
  if (stage == 0) return;
  // ---- scene depth (Tracker.cc:680-697) ---------------------------------------------------------
  {
    double sum = 0, sumsq = 0;
    for (int i = threadIdx.x; i < nf; i += blockDim.x) {
      const double z = d.p.v3cam[3 * (gb + fidx[i]) + 2];
      sum += z; sumsq += z * z;
    }
    sum = warp_sum(sum); sumsq = warp_sum(sumsq);
    __syncthreads();
    if (lane == 0) { red2[warp][0] = sum; red2[warp][1] = sumsq; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double sum = 0, sumsq = 0;
    for (int wq = 0; wq < kPoseThreads / 32; wq++) { sum += red2[wq][0]; sumsq += red2[wq][1]; }
    ptam_tracker_state& st = ctl.st;
    if (nf > 20) {
      st.scene_depth_mean = sum / nf;
      st.scene_depth_sigma = sqrt((sumsq / nf) - st.scene_depth_mean * st.scene_depth_mean);
    }
    if (fmode == 0) {  // a recovery frame runs TrackMap and AssessTrackingQuality only (Tracker.cc:174-177)
      // UpdateMotionModel (Tracker.cc:1035-1056)
      double inv[12], rel[12], motion[6];
      se3_inverse(ctl.start_pose, inv);
      se3_mul(pose, inv, rel);
      se3_ln(rel, motion);
      if (d.prm.use_constant_velocity) for (int q = 0; q < 6; q++) st.velocity[q] = motion[q];
      else for (int q = 0; q < 6; q++) st.velocity[q] = 0.9 * (0.5 * motion[q] + 0.5 * st.velocity[q]);
      double m = 0;
      for (int q = 0; q < 6; q++) {
        double v = st.velocity[q];
        if (q < 3) v *= 1.0 / st.scene_depth_mean;
        m += v * v;
      }
      st.msd_scaled_velocity_magnitude = sqrt(m);
    }
    // AssessTrackingQuality (Tracker.cc:1062-1107)
    int ta = 0, tf = 0, la = 0, lf = 0;
    for (int l = 0; l < kLevels; l++) {
      ta += ctl.attempted[l]; tf += ctl.found[l];
      if (l >= 2) { la += ctl.attempted[l]; lf += ctl.found[l]; }
    }
    ctl.needs_kf_distance = 0;
    if (tf == 0 || ta == 0) st.tracking_quality = 0;
    else {
      const double tfrac = (double)tf / ta;
      const double lfrac = la > 10 ? (double)lf / la : tfrac;
      if (tfrac > d.prm.quality_good) st.tracking_quality = 2;
      else if (lfrac < d.prm.quality_lost) st.tracking_quality = 0;
      else ctl.needs_kf_distance = 1;
    }
    if (st.tracking_quality == 0) st.lost_frames++; else st.lost_frames = 0;
    for (int i = 0; i < 12; i++) st.se3_cam_from_world[i] = pose[i];
  }
  PCLK(10);
}

// =============================================================================================
// MapMaker::AddPointEpipolar (MapMaker.cc:529-688; SURVEY 8f rank 3, second half) up to the sub-pixel
// position in the target keyframe.  Source = a stored keyframe, target = the stream's current frame.
//   k_epi_implane  thread per target corner: vImplaneCorners (MapMaker.cc:608-614) = UnProject of the
//                  truncated level-zero position.
//   k_epi_search   warp per candidate: epipolar segment of the candidate's view ray in the target's
//                  z = 1 plane (f64, every lane redundantly), un-warped 8x8 template
//                  (MakeTemplateCoarseNoWarp), scan of ALL target corners of the level against the
//                  segment (one corner per lane per pass), survivors queued per warp and scored four at
//                  a time with the same 8-lane dp4a ZMSSD as k_search, first minimum <= mnMaxSSD, then
//                  MakeSubPixTemplate + IterateSubPixToConvergence(10) in the same warp.
// The triangulation that follows (a 4x4 SVD per accepted point, MapMaker.cc:176-187) stays with the caller.
// =============================================================================================
struct EpiDev {
  int stream, level, src_kf, n_cand;
  double src[12], tgt[12];   // se3CfromW of the source / target keyframe
  double d_start, d_end;     // depth range along the ray (MapMaker.cc:552-556)
  double max_dist_sq;        // (OnePixelDist (4 + LevelScale))^2
  const int2* cand;          // irLevelPos of the candidates in the source level
  double2* implane;          // [n_corners of the level]
  int* found;                // 1: found and converged
  int* best;                 // index of the best corner in the target level's list, -1 if none
  double* sub;               // [n][2] sub-pixel level-zero position in the target
};

PTAM_DEV void cam_unproject(const CamModel& cam, double ix, double iy, double& ox, double& oy) {  // ATANCamera.cc:125-140
  const double d0 = (ix - cam.center[0]) * cam.inv_focal[0];
  const double d1 = (iy - cam.center[1]) * cam.inv_focal[1];
  const double dr = sqrt(d0 * d0 + d1 * d1);
  const double rr = cam.w == 0.0 ? dr : tan(dr * cam.w) * cam.one_over_tan2;
  const double f = dr > 0.01 ? rr / dr : 1.0;
  ox = f * d0; oy = f * d1;
}

__global__ void __launch_bounds__(256) k_epi_implane(TrackerDev d, EpiDev e) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nc = d.ctl[e.stream].n_corners[e.level];
  if (i >= nc) return;
  const int2 c = d.corners[(size_t)e.stream * d.g.corner_stride + d.g.lev[e.level].corner_off + i];
  const double px = (double)(int)level_zero_pos((double)c.x, e.level), py = (double)(int)level_zero_pos((double)c.y, e.level);
  double ox, oy;
  cam_unproject(d.cam, px, py, ox, oy);
  e.implane[i] = make_double2(ox, oy);
}

__global__ void __launch_bounds__(128) k_epi_search(TrackerDev d, EpiDev e) {
  __shared__ __align__(8) uint8_t stmpl[4][64];
  __shared__ int2 squeue[4][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ci = blockIdx.x * 4 + warp;
  if (ci >= e.n_cand) return;
  const int lv = e.level, s = e.stream;
  const int2 cp = e.cand[ci];
  auto fail = [&](int best_corner) {
    if (lane == 0) { e.found[ci] = 0; e.best[ci] = best_corner; e.sub[2 * ci] = 0.0; e.sub[2 * ci + 1] = 0.0; }
  };
  // ---- epipolar segment (MapMaker.cc:541-596) ---------------------------------------------------
  double normal[2], along[2], norm_dist, min_len, max_len;
  {
    double ux, uy;
    cam_unproject(d.cam, level_zero_pos((double)cp.x, lv), level_zero_pos((double)cp.y, lv), ux, uy);
    double ray[3] = {ux, uy, 1.0};
    const double nr = sqrt(ray[0] * ray[0] + ray[1] * ray[1] + ray[2] * ray[2]);
    for (int i = 0; i < 3; i++) ray[i] = ray[i] / nr;
    double ray_w[3], dirn[3], c_w[3], c_t[3];
    for (int i = 0; i < 3; i++) ray_w[i] = e.src[i] * ray[0] + e.src[3 + i] * ray[1] + e.src[6 + i] * ray[2];
    for (int i = 0; i < 3; i++) dirn[i] = e.tgt[3 * i] * ray_w[0] + e.tgt[3 * i + 1] * ray_w[1] + e.tgt[3 * i + 2] * ray_w[2];
    for (int i = 0; i < 3; i++) c_w[i] = -(e.src[i] * e.src[9] + e.src[3 + i] * e.src[10] + e.src[6 + i] * e.src[11]);
    for (int i = 0; i < 3; i++) c_t[i] = (e.tgt[3 * i] * c_w[0] + e.tgt[3 * i + 1] * c_w[1] + e.tgt[3 * i + 2] * c_w[2]) + e.tgt[9 + i];
    double rs[3], re[3];
    for (int i = 0; i < 3; i++) { rs[i] = c_t[i] + e.d_start * dirn[i]; re[i] = c_t[i] + e.d_end * dirn[i]; }
    if (re[2] <= rs[2] || re[2] <= 0.0) { fail(-1); return; }
    if (rs[2] <= 0.0) {
      const double k = 0.001 - rs[2] / dirn[2];
      for (int i = 0; i < 3; i++) rs[i] += dirn[i] * k;
    }
    const double A[2] = {rs[0] / rs[2], rs[1] / rs[2]}, B[2] = {re[0] / re[2], re[1] / re[2]};
    double al[2] = {A[0] - B[0], A[1] - B[1]};
    if (al[0] * al[0] + al[1] * al[1] < 1e-8) { fail(-1); return; }
    const double na = sqrt(al[0] * al[0] + al[1] * al[1]);
    al[0] = al[0] / na; al[1] = al[1] / na;
    along[0] = al[0]; along[1] = al[1];
    normal[0] = al[1]; normal[1] = -al[0];
    norm_dist = A[0] * normal[0] + A[1] * normal[1];
    if (fabs(norm_dist) > d.cam.largest_radius) { fail(-1); return; }
    const double la = al[0] * A[0] + al[1] * A[1], lb = al[0] * B[0] + al[1] * B[1];
    min_len = fmin(la, lb) - 0.05;
    max_len = fmax(la, lb) + 0.05;
    if (min_len < -2.0) min_len = -2.0;
    if (max_len < -2.0) max_len = -2.0;
    if (min_len > 2.0) min_len = 2.0;
    if (max_len > 2.0) max_len = 2.0;
  }
  // ---- MakeTemplateCoarseNoWarp (PatchFinder.cc:135-150): the 8x8 block around the candidate ------
  const LevelDesc& L = d.g.lev[lv];
  if (!(cp.x >= 5 && cp.y >= 5 && cp.x < L.w - 5 && cp.y < L.h - 5)) { fail(-1); return; }
  int tsum, tsumsq;
  {
    const uint8_t* src = d.kf_ptrs[e.src_kf] + L.img_off;
    const int trow = lane >> 2, tcol = (lane & 3) * 2;
    const uint8_t* sp = src + (size_t)(cp.y - 4 + trow) * L.pitch + (cp.x - 4 + tcol);
    const int t0 = sp[0], t1 = sp[1];
    tsum = warp_sum_int(t0 + t1);
    tsumsq = warp_sum_int(t0 * t0 + t1 * t1);
    stmpl[warp][2 * lane] = (uint8_t)t0; stmpl[warp][2 * lane + 1] = (uint8_t)t1;
  }
  __syncwarp();
  // ---- scan of all target corners (MapMaker.cc:616-636) ------------------------------------------
  int pitch;
  const uint8_t* im = level_image(d, s, lv, pitch);
  const int nc = d.ctl[s].n_corners[lv];
  const int2* corners = d.corners + (size_t)s * d.g.corner_stride + L.corner_off;
  const int sub = lane & 7, grp = lane >> 3;
  const unsigned tw0 = *reinterpret_cast<const unsigned*>(&stmpl[warp][8 * sub]);
  const unsigned tw1 = *reinterpret_cast<const unsigned*>(&stmpl[warp][8 * sub + 4]);
  int best_ssd = kMaxSSD + 1, best_idx = 0x7fffffff;
  auto process = [&](int n) {
    for (int r0 = 0; r0 < n; r0 += 4) {
      const int kq = r0 + grp;
      const bool valid = kq < n;
      const int2 q = squeue[warp][valid ? kq : 0];
      const int cx = q.x & 0xffff, cy = q.x >> 16;
      const uint8_t* ip = im + (size_t)(cy - 4 + sub) * pitch + (cx - 4);
      const unsigned a = (unsigned)(reinterpret_cast<uintptr_t>(ip) & 3);
      const unsigned* wp = reinterpret_cast<const unsigned*>(ip - a);
      const unsigned w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = a ? __ldg(wp + 2) : 0u;
      const unsigned sel = 0x3210u + 0x1111u * a;
      const unsigned v0 = __byte_perm(w0, w1, sel), v1 = __byte_perm(w1, w2, sel);
      int isum = (int)__dp4a(v0, 0x01010101u, __dp4a(v1, 0x01010101u, 0u));
      int isq = (int)__dp4a(v0, v0, __dp4a(v1, v1, 0u));
      int cross = (int)__dp4a(v0, tw0, __dp4a(v1, tw1, 0u));
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        isum += __shfl_xor_sync(kFull, isum, o);
        isq += __shfl_xor_sync(kFull, isq, o);
        cross += __shfl_xor_sync(kFull, cross, o);
      }
      const int SA = tsum, SB = isum;
      const int ssd = ((2 * SA * SB - SA * SA - SB * SB) / 64 + isq + tsumsq - 2 * cross);
      if (valid && (ssd < best_ssd || (ssd == best_ssd && q.y < best_idx))) { best_ssd = ssd; best_idx = q.y; }
    }
    __syncwarp();
  };
  int qn = 0;
  const unsigned lt = (1u << lane) - 1u;
  for (int b0 = 0; b0 < nc; b0 += 32) {
    const int i = b0 + lane;
    int2 c = make_int2(0, 0);
    bool pass = false;
    if (i < nc) {
      const double2 ip = e.implane[i];
      const double dd = norm_dist - (ip.x * normal[0] + ip.y * normal[1]);
      if (!(dd * dd > e.max_dist_sq)) {
        const double al = ip.x * along[0] + ip.y * along[1];
        if (!(al < min_len) && !(al > max_len)) {
          c = corners[i];
          pass = c.x >= 4 && c.y >= 4 && c.x < L.w - 4 && c.y < L.h - 4;  // else ZMSSDAtPoint = mnMaxSSD + 1: never the best
        }
      }
    }
    const unsigned m = __ballot_sync(kFull, pass);
    if (pass) squeue[warp][qn + __popc(m & lt)] = make_int2(c.x | (c.y << 16), i);
    qn += __popc(m);
    if (qn > 32) { __syncwarp(); process(qn); qn = 0; }
  }
  __syncwarp();
  process(qn);
  {
    const int bs = __reduce_min_sync(kFull, best_ssd);
    best_idx = __reduce_min_sync(kFull, best_ssd == bs ? best_idx : 0x7fffffff);
    best_ssd = bs;
  }
  if (!(best_ssd < kMaxSSD + 1)) { fail(-1); return; }  // nBest == -1
  const int2 bc = corners[best_idx];
  // ---- MakeSubPixTemplate + IterateSubPixToConvergence(kTarget, 10) (MapMaker.cc:640-646) -------
  float jx[2] = {0.f, 0.f}, jy[2] = {0.f, 0.f};
  int tq[2] = {0, 0};
  double h00 = 0, h01 = 0, h02 = 0, h11 = 0, h12 = 0, h22 = 0;
#pragma unroll
  for (int q = 0; q < 2; q++) {
    const int pi = lane + 32 * q;
    if (pi < 36) {
      const int y = pi / 6 + 1, x = pi % 6 + 1;
      const uint8_t* T = stmpl[warp];
      const double gx = 0.5 * (double)((int)T[8 * y + x + 1] - (int)T[8 * y + x - 1]);
      const double gy = 0.5 * (double)((int)T[8 * (y + 1) + x] - (int)T[8 * (y - 1) + x]);
      jx[q] = (float)gx; jy[q] = (float)gy; tq[q] = T[8 * y + x];
      h00 += gx * gx; h01 += gx * gy; h02 += gx; h11 += gy * gy; h12 += gy; h22 += 1.0;
    }
  }
  double H[9], Hi[9];
  H[0] = warp_sum(h00); H[1] = H[3] = warp_sum(h01); H[2] = H[6] = warp_sum(h02);
  H[4] = warp_sum(h11); H[5] = H[7] = warp_sum(h12); H[8] = warp_sum(h22);
  ldlt_inverse<3>(H, Hi);
  double sp[2] = {level_zero_pos((double)bc.x, lv), level_zero_pos((double)bc.y, lv)};
  double mean_diff = 0.0;
  bool converged = false;
  for (int it = 0; it < 10; it++) {
    const double cx = level_n_pos(sp[0], lv), cy = level_n_pos(sp[1], lv);
    const int rx = (int)(cx > 0.0 ? cx + 0.5 : cx - 0.5), ry = (int)(cy > 0.0 ? cy + 0.5 : cy - 0.5);
    if (!(rx >= 5 && ry >= 5 && rx < L.w - 5 && ry < L.h - 5)) break;
    const double bx = cx - 4, by = cy - 4;
    const double dX = bx - floor(bx), dY = by - floor(by);
    const float fTL = (float)((1.0 - dX) * (1.0 - dY));
    const float fTR = (float)((dX) * (1.0 - dY));
    const float fBL = (float)((1.0 - dX) * (dY));
    const float fBR = (float)((dX) * (dY));
    const int ibx = (int)bx, iby = (int)by;
    double a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int pi = lane + 32 * q;
      if (pi < 36) {
        const int y = pi / 6 + 1, x = pi % 6 + 1;
        const uint8_t* tl = im + (size_t)(iby + y) * pitch + ibx + x;
        const float fp = fTL * (float)tl[0] + fTR * (float)tl[1] + fBL * (float)tl[pitch] + fBR * (float)tl[pitch + 1];
        const double diff = (double)(fp - (float)tq[q]) + mean_diff;
        a0 += diff * (double)jx[q]; a1 += diff * (double)jy[q]; a2 += diff;
      }
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    const double u0 = Hi[0] * a0 + Hi[1] * a1 + Hi[2] * a2;
    const double u1 = Hi[3] * a0 + Hi[4] * a1 + Hi[5] * a2;
    const double u2 = Hi[6] * a0 + Hi[7] * a1 + Hi[8] * a2;
    sp[0] -= u0 * (double)(1 << lv);
    sp[1] -= u1 * (double)(1 << lv);
    mean_diff -= u2;
    if (u0 * u0 + u1 * u1 < 0.03 * 0.03) { converged = true; break; }
  }
  if (!converged) { fail(best_idx); return; }
  if (lane == 0) { e.found[ci] = 1; e.best[ci] = best_idx; e.sub[2 * ci] = sp[0]; e.sub[2 * ci + 1] = sp[1]; }
}

}  // namespace ptam
