// Host mirror of the reference's shared-state structs on path T: Level / Measurement / KeyFrame
// (include/KeyFrame.h:40-152) and MapPoint / Map (include/Map.h:28-101), with
// KeyFrame::MakeKeyFrame_Lite (src/KeyFrame.cc:18-54) forwarding to the CUDA library: pyramid,
// FAST-10 corners in raster order and the row LUT are computed on the device and copied back into
// the same members the reference fills (Level::im, vCorners, vCornerRowLUT).
#pragma once
#include <cmath>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/ptam_b200.h"
#include "ATANCamera.h"

#ifndef LEVELS
#define LEVELS 4
#endif

namespace ptam_b200 {

struct MapPoint;

struct Measurement {  // KeyFrame.h:44-50
  int nLevel;
  bool bSubPix;
  TooN::Vector<2> v2RootPos;
  enum { SRC_TRACKER, SRC_REFIND, SRC_ROOT, SRC_TRAIL, SRC_EPIPOLAR } Source;
};

struct Candidate {  // KeyFrame.h:36-41
  CVD::ImageRef irLevelPos;
  TooN::Vector<2> v2RootPos;
  double dSTScore;
};

struct Level {  // KeyFrame.h:55-125
  CVD::Image<CVD::byte> im;
  std::vector<CVD::ImageRef> vCorners;
  std::vector<int> vCornerRowLUT;
  std::vector<CVD::ImageRef> vMaxCorners;
  std::vector<Candidate> vCandidates;
  static int LevelScale(int nLevel) { return 1 << nLevel; }
  static double LevelZeroPos(double dLevelPos, int nLevel) { return (dLevelPos + 0.5) * LevelScale(nLevel) - 0.5; }
  static double LevelNPos(double dRootPos, int nLevel) { return (dRootPos + 0.5) / LevelScale(nLevel) - 0.5; }
  static TooN::Vector<2> LevelZeroPos(CVD::ImageRef ir, int nLevel) { return TooN::makeVector(LevelZeroPos(ir.x, nLevel), LevelZeroPos(ir.y, nLevel)); }
  static TooN::Vector<2> LevelZeroPos(TooN::Vector<2> v, int nLevel) { return TooN::makeVector(LevelZeroPos(v[0], nLevel), LevelZeroPos(v[1], nLevel)); }
  static TooN::Vector<2> LevelNPos(TooN::Vector<2> v, int nLevel) { return TooN::makeVector(LevelNPos(v[0], nLevel), LevelNPos(v[1], nLevel)); }
};

// One S=1 tracker handle per (device, image size) and thread, used by KeyFrame::MakeKeyFrame_Lite
// when it is called outside a Tracker (MapMaker builds keyframes too, MapMaker.cc:282-285).
inline ptam_tracker* keyframe_context(int w, int h, int device = 0) {
  struct Holder {
    ptam_tracker* t = nullptr; int w = 0, h = 0, dev = 0;
    ~Holder() { if (t) ptam_tracker_destroy(t); }
  };
  static thread_local std::vector<std::unique_ptr<Holder> > pool;
  for (auto& c : pool) if (c->w == w && c->h == h && c->dev == device) return c->t;
  const double cam[5] = {1.0803, 1.43987, 0.519983, 0.548655, 0.244943};  // unused by keyframe making
  ptam_tracker* t = ptam_tracker_create(device, cam, w, h, 1, nullptr);
  if (!t) throw std::runtime_error(std::string("ptam_tracker_create: ") + ptam_global_last_error());
  pool.emplace_back(new Holder);
  pool.back()->t = t; pool.back()->w = w; pool.back()->h = h; pool.back()->dev = device;
  return t;
}

struct KeyFrame {  // KeyFrame.h:130-150
  TooN::SE3<> se3CfromW;
  bool bFixed = false;
  Level aLevels[LEVELS];
  std::map<MapPoint*, Measurement> mMeasurements;
  double dSceneDepthMean = 1.0, dSceneDepthSigma = 1.0;

  // Copies level `l` of `stream` of a tracker handle into aLevels[l].
  void FetchLevel(ptam_tracker* t, int stream, int l) {
    int w = 0, h = 0;
    ptam_tracker_level_size(t, l, &w, &h);
    Level& L = aLevels[l];
    L.im.resize(CVD::ImageRef(w, h));
    L.vCornerRowLUT.assign(h, 0);
    const int n = ptam_tracker_get_level(t, stream, l, L.im.data(), nullptr, 0, L.vCornerRowLUT.data());
    if (n < 0) throw std::runtime_error(ptam_tracker_last_error(t));
    static_assert(sizeof(CVD::ImageRef) == 2 * sizeof(int32_t), "ImageRef is two ints");
    L.vCorners.resize(n);
    if (n) ptam_tracker_get_level(t, stream, l, nullptr, reinterpret_cast<int32_t*>(L.vCorners.data()), n, nullptr);
  }

  void MakeKeyFrame_Lite(CVD::BasicImage<CVD::byte>& im) {  // KeyFrame.cc:18-54
    ptam_tracker* t = keyframe_context(im.size().x, im.size().y);
    const uint8_t* ptrs[1] = {im.data()};
    if (ptam_tracker_make_keyframes(t, ptrs, im.row_stride()) != PTAM_OK) throw std::runtime_error(ptam_tracker_last_error(t));
    for (int l = 0; l < LEVELS; l++) FetchLevel(t, 0, l);
  }

  // KeyFrame.cc:61-82: FAST non-max suppression and Shi-Tomasi candidates on the device
  // (MapMaker.CandidateMinShiTomasiScore: 70 in code, 400 in the shipped settings.cfg:27).
  // The level-0 image is sent again: the keyframe context may have processed other frames since.
  void MakeKeyFrame_Rest(double dCandidateMinSTScore = 70.0) {
    CVD::Image<CVD::byte>& im0 = aLevels[0].im;
    ptam_tracker* t = keyframe_context(im0.size().x, im0.size().y);
    const uint8_t* ptrs[1] = {im0.data()};
    if (ptam_tracker_make_keyframes(t, ptrs, im0.row_stride()) != PTAM_OK) throw std::runtime_error(ptam_tracker_last_error(t));
    if (ptam_tracker_keyframe_rest(t, 0, dCandidateMinSTScore) != PTAM_OK) throw std::runtime_error(ptam_tracker_last_error(t));
    for (int l = 0; l < LEVELS; l++) {
      Level& L = aLevels[l];
      int nc = 0;
      const int nm = ptam_tracker_get_level_rest(t, 0, l, nullptr, 0, nullptr, nullptr, 0, &nc);
      if (nm < 0) throw std::runtime_error(ptam_tracker_last_error(t));
      L.vMaxCorners.resize(nm);
      std::vector<CVD::ImageRef> cpos(nc);
      std::vector<double> cscore(nc);
      ptam_tracker_get_level_rest(t, 0, l, reinterpret_cast<int32_t*>(L.vMaxCorners.data()), nm,
                                  reinterpret_cast<int32_t*>(cpos.data()), cscore.data(), nc, &nc);
      L.vCandidates.resize(nc);
      for (int i = 0; i < nc; i++) { L.vCandidates[i].irLevelPos = cpos[i]; L.vCandidates[i].dSTScore = cscore[i]; }
    }
  }
};

struct MapPoint {  // Map.h:46-98 (the fields the tracker and the map maker's hot paths read and write)
  TooN::Vector<3> v3WorldPos;
  bool bBad = false;
  KeyFrame* pPatchSourceKF = nullptr;
  int nSourceLevel = 0;
  CVD::ImageRef irCenter;
  // the source patch's plane: unit view rays of its centre and of one level-pixel right / down of it, and
  // the plane normal, all in the source keyframe's camera frame (Map.h:68-75)
  TooN::Vector<3> v3Center_NC, v3OneDownFromCenter_NC, v3OneRightFromCenter_NC, v3Normal_NC;
  TooN::Vector<3> v3PixelDown_W, v3PixelRight_W;
  int nMEstimatorOutlierCount = 0, nMEstimatorInlierCount = 0;

  // Map.cc:40-65: world-frame steps of one source-level pixel right / down on the patch's plane
  void RefreshPixelVectors() {
    const KeyFrame& k = *pPatchSourceKF;
    const TooN::Vector<3> on_plane_c = k.se3CfromW * v3WorldPos;
    const double height = std::fabs(on_plane_c * v3Normal_NC);
    auto hit = [&](const TooN::Vector<3>& ray) {  // (ray * height) / rate, in the reference's order of operations
      const double rate = std::fabs(ray * v3Normal_NC);
      return TooN::makeVector(ray[0] * height / rate, ray[1] * height / rate, ray[2] * height / rate);
    };
    const TooN::Vector<3> c = hit(v3Center_NC), r = hit(v3OneRightFromCenter_NC), d = hit(v3OneDownFromCenter_NC);
    const TooN::SO3<> Rwc = k.se3CfromW.get_rotation().inverse();
    v3PixelRight_W = Rwc * (r - c);
    v3PixelDown_W = Rwc * (d - c);
  }
};

struct Map {  // Map.h:28-44
  std::vector<MapPoint*> vpPoints;
  std::vector<MapPoint*> vpPointsTrash;
  // Map.cc:20-30 (whoever owns the points empties the trash: the tracker may still hold a pointer for a frame)
  void MoveBadPointsToTrash() {
    for (int i = (int)vpPoints.size() - 1; i >= 0; i--)
      if (vpPoints[i]->bBad) { vpPointsTrash.push_back(vpPoints[i]); vpPoints.erase(vpPoints.begin() + i); nRevision++; }
  }
  std::vector<KeyFrame*> vpKeyFrames;
  bool bGood = false;
  // bumped by whoever edits vpPoints / point geometry, so that the Tracker re-uploads the map
  // (the reference shares the containers between threads with no signalling at all, Map.h:8-13)
  unsigned nRevision = 0;
  bool IsGood() const { return bGood; }
};

}  // namespace ptam_b200
