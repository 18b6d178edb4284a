// Host mirror of class MapMaker (reference include/MapMaker.h:38-160) on the hot paths: the bundle-adjustment
// side — MapMaker is the only caller of Bundle — and the epipolar search for new map points
// (AddPointsEpipolar).  BundleAdjustAll / BundleAdjustRecent choose the keyframe and point sets
// (src/MapMaker.cc:767-836), BundleAdjust marshals them into a Bundle, runs it and writes the adjusted state
// and the outlier bookkeeping back into the map (src/MapMaker.cc:838-933).  Same member names, argument
// meaning and side effects as the reference; the arithmetic runs behind ptam_bundle_* (host/Bundle.h).
//
// Identity of the bundle ids matters for a drop-in: the reference walks std::set<KeyFrame*> /
// std::set<MapPoint*> (pointer order) for cameras and points, and mMap.vpKeyFrames (insertion order) with each
// keyframe's std::map<MapPoint*, Measurement> (pointer order) for the measurements; so does this class.
#pragma once
#include <algorithm>
#include <functional>
#include <cmath>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "Bundle.h"
#include "KeyFrame.h"

namespace ptam_b200 {

// MapMaker.h:28-35: the map maker's per-point bookkeeping (MapPoint::pMMData in the reference; kept beside the
// point here because the mirror's MapPoint only carries what the tracker touches)
struct MapMakerData {
  std::set<KeyFrame*> sMeasurementKFs;  // keyframes holding a measurement of the point
  std::set<KeyFrame*> sNeverRetryKFs;   // keyframes in which re-finding it is pointless
  int GoodMeasCount() const { return (int)sMeasurementKFs.size(); }
};

class MapMaker {
 public:
  MapMaker(Map& m, const ATANCamera& cam, int device = 0, const ptam_bundle_params* params = nullptr)
      : mMap(m), mCamera(cam), mnDevice(device) {
    if (params) { mParams = *params; mbHaveParams = true; }
  }
  ~MapMaker() { if (mpAssoc) ptam_tracker_destroy(mpAssoc); }
  MapMaker(const MapMaker&) = delete;
  MapMaker& operator=(const MapMaker&) = delete;

  // MapMaker.cc:172-187: the point seen at v2A / v2B (z = 1 plane coordinates in frames A / B), in frame B: the
  // right singular vector of the smallest singular value of the 4x4 DLT matrix (the reference asks TooN's
  // LAPACK-backed SVD<4>; here a one-sided Jacobi SVD, which needs no library)
  static TooN::Vector<3> Triangulate(const TooN::SE3<>& se3AfromB, const TooN::Vector<2>& v2A, const TooN::Vector<2>& v2B) {
    double P[3][4];
    for (int r = 0; r < 3; r++) {
      for (int c = 0; c < 3; c++) P[r][c] = se3AfromB.get_rotation().get_matrix()(r, c);
      P[r][3] = se3AfromB.get_translation()[r];
    }
    double A[4][4] = {{-1.0, 0.0, v2B[0], 0.0}, {0.0, -1.0, v2B[1], 0.0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
    for (int c = 0; c < 4; c++) { A[2][c] = v2A[0] * P[2][c] - P[0][c]; A[3][c] = v2A[1] * P[2][c] - P[1][c]; }
    // Hestenes: rotate pairs of columns of A until they are orthogonal, accumulating the rotations in V
    double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    for (int sweep = 0; sweep < 60; sweep++) {
      double off = 0.0;
      for (int p = 0; p < 3; p++)
        for (int q = p + 1; q < 4; q++) {
          double a = 0, b = 0, g = 0;
          for (int k = 0; k < 4; k++) { a += A[k][p] * A[k][p]; b += A[k][q] * A[k][q]; g += A[k][p] * A[k][q]; }
          if (g == 0.0) continue;
          off = std::max(off, std::fabs(g) / std::sqrt(a * b));
          const double zeta = (b - a) / (2.0 * g);
          const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
          const double cs = 1.0 / std::sqrt(1.0 + t * t), sn = cs * t;
          for (int k = 0; k < 4; k++) {
            const double ap = A[k][p], aq = A[k][q];
            A[k][p] = cs * ap - sn * aq; A[k][q] = sn * ap + cs * aq;
            const double vp = V[k][p], vq = V[k][q];
            V[k][p] = cs * vp - sn * vq; V[k][q] = sn * vp + cs * vq;
          }
        }
      if (off < 1e-15) break;
    }
    int smallest = 0;
    double best = -1.0;
    for (int c = 0; c < 4; c++) {
      double n2 = 0;
      for (int k = 0; k < 4; k++) n2 += A[k][c] * A[k][c];
      if (best < 0 || n2 < best) { best = n2; smallest = c; }
    }
    double w = V[3][smallest];
    if (w == 0.0) w = 0.00001;
    return TooN::makeVector(V[0][smallest] / w, V[1][smallest] / w, V[2][smallest] / w);
  }

  // The candidate loop of MapMaker::AddSomeMapPoints (MapMaker.cc:496-527) with MapMaker::AddPointEpipolar
  // (MapMaker.cc:529-688) for every candidate of kSrc's level: the epipolar search up to the sub-pixel position
  // in kTarget runs on the device for all candidates at once (ptam_tracker_epipolar_search); triangulation,
  // MapPoint construction and the two measurements follow here as in the reference (MapMaker.cc:648-687).
  // Both keyframes must have been made (MakeKeyFrame_Lite; kSrc also MakeKeyFrame_Rest for its candidates).
  // Returns the number of points added to the map; they are also appended to mvpNewQueue (mqNewQueue).
  int AddPointsEpipolar(KeyFrame& kSrc, KeyFrame& kTarget, int nLevel) {
    const std::vector<Candidate>& cands = kSrc.aLevels[nLevel].vCandidates;
    const int n = (int)cands.size();
    if (n == 0) return 0;
    ptam_tracker* t = Assoc(kSrc.aLevels[0].im.size());
    auto stored = mStoreId.find(&kSrc);
    if (stored == mStoreId.end()) {
      const int id = ptam_tracker_add_keyframe(t, kSrc.aLevels[0].im.data(), kSrc.aLevels[0].im.row_stride());
      if (id < 0) throw std::runtime_error(ptam_tracker_last_error(t));
      stored = mStoreId.emplace(&kSrc, id).first;
    }
    const uint8_t* target_image[1] = {kTarget.aLevels[0].im.data()};
    if (ptam_tracker_make_keyframes(t, target_image, kTarget.aLevels[0].im.row_stride()) != PTAM_OK)
      throw std::runtime_error(ptam_tracker_last_error(t));
    std::vector<int32_t> xy(2 * (size_t)n), found(n), best(n);
    std::vector<double> sub(2 * (size_t)n);
    for (int i = 0; i < n; i++) { xy[2 * i] = cands[i].irLevelPos.x; xy[2 * i + 1] = cands[i].irLevelPos.y; }
    double src12[12], tgt12[12];
    se3_to_array(kSrc.se3CfromW, src12);
    se3_to_array(kTarget.se3CfromW, tgt12);
    if (ptam_tracker_epipolar_search(t, 0, nLevel, stored->second, src12, kSrc.dSceneDepthMean, kSrc.dSceneDepthSigma, tgt12,
                                     mdWiggleScale, n, xy.data(), found.data(), best.data(), sub.data()) != PTAM_OK)
      throw std::runtime_error(ptam_tracker_last_error(t));
    const int nLevelScale = Level::LevelScale(nLevel);
    const TooN::SE3<> se3SrcfromTarget = kSrc.se3CfromW * kTarget.se3CfromW.inverse();
    const TooN::SE3<> se3WfromTarget = kTarget.se3CfromW.inverse();
    auto unit_ray = [&](const TooN::Vector<2>& im) {
      const TooN::Vector<2> p = mCamera.UnProject(im);
      const double nrm = std::sqrt(p[0] * p[0] + p[1] * p[1] + 1.0);
      return TooN::makeVector(p[0] / nrm, p[1] / nrm, 1.0 / nrm);
    };
    int added = 0;
    for (int i = 0; i < n; i++) {
      if (!found[i]) continue;
      const TooN::Vector<2> v2RootPos = Level::LevelZeroPos(cands[i].irLevelPos, nLevel);
      const TooN::Vector<2> v2SubPosTarget = TooN::makeVector(sub[2 * i], sub[2 * i + 1]);
      mOwnedPoints.emplace_back(new MapPoint);  // the reference's Map owns its points; here their maker does
      MapPoint* pNew = mOwnedPoints.back().get();
      pNew->v3WorldPos = se3WfromTarget * Triangulate(se3SrcfromTarget, mCamera.UnProject(v2RootPos), mCamera.UnProject(v2SubPosTarget));
      pNew->pPatchSourceKF = &kSrc;
      pNew->nSourceLevel = nLevel;
      pNew->v3Normal_NC = TooN::makeVector(0.0, 0.0, -1.0);
      pNew->irCenter = cands[i].irLevelPos;
      pNew->v3Center_NC = unit_ray(v2RootPos);
      pNew->v3OneRightFromCenter_NC = unit_ray(v2RootPos + TooN::makeVector((double)nLevelScale, 0.0));
      pNew->v3OneDownFromCenter_NC = unit_ray(v2RootPos + TooN::makeVector(0.0, (double)nLevelScale));
      pNew->RefreshPixelVectors();
      mMap.vpPoints.push_back(pNew);
      mMap.nRevision++;
      mvpNewQueue.push_back(pNew);
      Measurement m;
      m.Source = Measurement::SRC_ROOT;
      m.v2RootPos = v2RootPos;
      m.nLevel = nLevel;
      m.bSubPix = true;
      kSrc.mMeasurements[pNew] = m;
      m.Source = Measurement::SRC_EPIPOLAR;
      m.v2RootPos = v2SubPosTarget;
      kTarget.mMeasurements[pNew] = m;
      MapMakerData& md = MMData(pNew);
      md.sMeasurementKFs.insert(&kSrc);
      md.sMeasurementKFs.insert(&kTarget);
      added++;
    }
    return added;
  }

  // MapMaker.cc:1026-1040 with ReFind_Common (MapMaker.cc:943-1018) for every map point: the data association of
  // a keyframe the tracker handed over.  Projection, warp / template, the radius-4 coarse search and the
  // sub-pixel refinement of all points run on the device in one call (ptam_tracker_refind_in_keyframes); the
  // reference's bookkeeping follows here: points already measured in k, or given up on, are left alone; a point
  // that was not found (for whatever reason: behind the camera, off the image, bad template, no match) is never
  // retried in k; a found one gets its Measurement (SRC_REFIND) in k.  Returns the number of new measurements.
  int ReFindInSingleKeyFrame(KeyFrame& k) { return ReFindIn(k, nullptr); }

  // MapMaker.cc:1046-1065: the points made since the last call get their chance in every keyframe of the map
  // (ReFind_Common per keyframe and point; the pairs are independent, so the device call per keyframe covers all
  // queued points at once).  The reference also stops when a keyframe arrives from the tracker; the mirror has no
  // keyframe queue, the caller simply calls again.
  int ReFindNewlyMade() {
    std::set<MapPoint*> fresh;
    for (MapPoint* p : mvpNewQueue)
      if (!p->bBad) fresh.insert(p);
    mvpNewQueue.clear();
    int nFound = 0;
    if (!fresh.empty())
      for (KeyFrame* kf : mMap.vpKeyFrames) nFound += ReFindIn(*kf, &fresh);
    return nFound;
  }

  // MapMaker.cc:1070-1082: measurements bundle adjustment threw out get a second chance
  int ReFindFromFailureQueue() {
    std::map<KeyFrame*, std::set<MapPoint*> > by_keyframe;
    for (auto& kp : mvFailureQueue) by_keyframe[kp.first].insert(kp.second);
    mvFailureQueue.clear();
    int nFound = 0;
    for (auto& ks : by_keyframe) nFound += ReFindIn(*ks.first, &ks.second);
    return nFound;
  }

 private:
  // ReFind_Common for the points of `only` (all map points when null) in keyframe k
  int ReFindIn(KeyFrame& k, const std::set<MapPoint*>* only) {
    const size_t n = mMap.vpPoints.size();
    if (n == 0) return 0;
    ptam_tracker* t = Assoc(k.aLevels[0].im.size());
    SyncAssocMap(t);
    const uint8_t* image[1] = {k.aLevels[0].im.data()};
    double se3[12];
    se3_to_array(k.se3CfromW, se3);
    if (ptam_tracker_refind_in_keyframes(t, image, k.aLevels[0].im.row_stride(), se3) != PTAM_OK)
      throw std::runtime_error(ptam_tracker_last_error(t));
    std::vector<int32_t> flags(n), level(n);
    std::vector<double> found(2 * n);
    if (ptam_tracker_get_points(t, 0, flags.data(), level.data(), found.data(), nullptr, nullptr, nullptr) < 0)
      throw std::runtime_error(ptam_tracker_last_error(t));
    int nFoundNow = 0;
    for (size_t i = 0; i < n; i++) {
      MapPoint* p = mMap.vpPoints[i];
      if (only && !only->count(p)) continue;
      MapMakerData& md = MMData(p);
      if (md.sMeasurementKFs.count(&k) || md.sNeverRetryKFs.count(&k)) continue;
      if (!(flags[i] & PTAM_PT_FOUND)) { md.sNeverRetryKFs.insert(&k); continue; }
      Measurement m;
      m.nLevel = level[i];
      m.Source = Measurement::SRC_REFIND;
      m.bSubPix = level[i] > 0;
      m.v2RootPos = TooN::makeVector(found[2 * i], found[2 * i + 1]);
      k.mMeasurements[p] = m;
      md.sMeasurementKFs.insert(&k);
      nFoundNow++;
    }
    return nFoundNow;
  }

 public:

  // bookkeeping of a point, created on first use (the reference allocates pMMData when the point is made)
  MapMakerData& MMData(MapPoint* p) { return mMMData[p]; }

  // MapMaker.cc:696-704
  static double KeyFrameLinearDist(const KeyFrame& k1, const KeyFrame& k2) {
    const TooN::Vector<3> c1 = k1.se3CfromW.inverse().get_translation(), c2 = k2.se3CfromW.inverse().get_translation();
    double d2 = 0.0;
    for (int i = 0; i < 3; i++) d2 += (c2[i] - c1[i]) * (c2[i] - c1[i]);
    return std::sqrt(d2);
  }

  // MapMaker.cc:711-731: the N keyframes nearest to k, nearest first
  std::vector<KeyFrame*> NClosestKeyFrames(KeyFrame& k, unsigned N) {
    std::vector<std::pair<double, KeyFrame*> > scored;
    for (KeyFrame* other : mMap.vpKeyFrames)
      if (other != &k) scored.emplace_back(KeyFrameLinearDist(k, *other), other);
    N = std::min<unsigned>(N, (unsigned)scored.size());
    std::partial_sort(scored.begin(), scored.begin() + N, scored.end());
    std::vector<KeyFrame*> out;
    for (unsigned i = 0; i < N; i++) out.push_back(scored[i].second);
    return out;
  }

  int QueueSize() const { return (int)mvpKeyFrameQueue.size(); }   // MapMaker.h:52
  double GetWiggleScale() const { return mdWiggleScale; }           // MapMaker.h:57

  // MapMaker.cc:754-764: is the camera far enough from every keyframe, relative to the scene depth, for a new one?
  bool IsNeedNewKeyFrame(KeyFrame& kCurrent) {
    const double dist = KeyFrameLinearDist(kCurrent, *ClosestKeyFrame(kCurrent)) * (1.0 / kCurrent.dSceneDepthMean);
    return dist > mdMaxKFDistWiggleMult * mdWiggleScaleDepthNormalized;
  }

  // MapMaker.cc:738-752: the keyframe of the map nearest to k (not k itself)
  KeyFrame* ClosestKeyFrame(KeyFrame& k) {
    KeyFrame* best = nullptr;
    double best_dist = 9999999999.9;
    for (KeyFrame* other : mMap.vpKeyFrames) {
      if (other == &k) continue;
      const double d = KeyFrameLinearDist(k, *other);
      if (d < best_dist) { best_dist = d; best = other; }
    }
    if (!best) throw std::logic_error("ClosestKeyFrame: the map has no other keyframe");
    return best;
  }

  // MapMaker.cc:413-441: drop the candidates of one level that sit within 10 level-pixels of a point the keyframe
  // already measures on that level or the next coarser one
  static void ThinCandidates(KeyFrame& k, int nLevel) {
    std::vector<CVD::ImageRef> busy;
    const int scale = Level::LevelScale(nLevel);
    auto rounded = [](double v) { return (int)(v > 0.0 ? v + 0.5 : v - 0.5); };  // CVD::ir_rounded
    for (auto& pm : k.mMeasurements)
      if (pm.second.nLevel == nLevel || pm.second.nLevel == nLevel + 1)
        busy.emplace_back(rounded(pm.second.v2RootPos[0] / scale), rounded(pm.second.v2RootPos[1] / scale));
    std::vector<Candidate>& cands = k.aLevels[nLevel].vCandidates;
    std::vector<Candidate> kept;
    for (const Candidate& c : cands) {
      bool clear = true;
      for (const CVD::ImageRef& b : busy) {
        const int dx = b.x - c.irLevelPos.x, dy = b.y - c.irLevelPos.y;
        if ((unsigned)(dx * dx + dy * dy) < 100u) { clear = false; break; }
      }
      if (clear) kept.push_back(c);
    }
    cands.swap(kept);
  }

  // MapMaker.cc:449-458: new points for the newest keyframe on one level, by epipolar search in its nearest neighbour
  int AddSomeMapPoints(int nLevel) {
    KeyFrame& kSrc = *mMap.vpKeyFrames.back();
    KeyFrame& kTarget = *ClosestKeyFrame(kSrc);
    ThinCandidates(kSrc, nLevel);
    return AddPointsEpipolar(kSrc, kTarget, nLevel);
  }

  // MapMaker.cc:480-488: the tracker hands over a keyframe (it is copied; an ongoing adjustment is asked to stop)
  void AddKeyFrame(KeyFrame& k) {
    mOwnedKeyFrames.emplace_back(new KeyFrame(k));
    mvpKeyFrameQueue.push_back(mOwnedKeyFrames.back().get());
    if (mbBundleRunning) mbBundleAbortRequested = true;
  }

  // MapMaker.cc:493-519: candidates of the oldest queued keyframe, its measurements entered into the points'
  // bookkeeping, the points the tracker missed re-found, new points on levels 3, 0, 1, 2
  void AddKeyFrameFromTopOfQueue() {
    if (mvpKeyFrameQueue.empty()) return;
    KeyFrame* pK = mvpKeyFrameQueue.front();
    mvpKeyFrameQueue.erase(mvpKeyFrameQueue.begin());
    pK->MakeKeyFrame_Rest(mdCandidateMinSTScore);
    mMap.vpKeyFrames.push_back(pK);
    mMap.nRevision++;  // a new keyframe (and its pose) for the tracker's relocaliser
    for (auto& pm : pK->mMeasurements) {
      MMData(pm.first).sMeasurementKFs.insert(pK);
      pm.second.Source = Measurement::SRC_TRACKER;
    }
    ReFindInSingleKeyFrame(*pK);
    AddSomeMapPoints(3);
    AddSomeMapPoints(0);
    AddSomeMapPoints(1);
    AddSomeMapPoints(2);
    mbBundleConverged_Full = false;
    mbBundleConverged_Recent = false;
  }

  // MapMaker.cc:131-153: points the tracker's M-estimator rejected more often than not become bad; every bad
  // point loses its measurements and leaves the map
  void HandleBadPoints() {
    if (mfnRefreshCounters) mfnRefreshCounters();  // the tracker's M-estimator counts live on the device between keyframes
    for (MapPoint* p : mMap.vpPoints)
      if (p->nMEstimatorOutlierCount > 20 && p->nMEstimatorOutlierCount > p->nMEstimatorInlierCount) p->bBad = true;
    for (MapPoint* p : mMap.vpPoints)
      if (p->bBad)
        for (KeyFrame* kf : mMap.vpKeyFrames) kf->mMeasurements.erase(p);
    mMap.MoveBadPointsToTrash();
  }

  // One pass of the map-maker thread's priority list (MapMaker::run, MapMaker.cc:83-116), for callers without the
  // reference's thread.  The reference gives the failure queue its second chance on a 1-in-20 draw of rand();
  // here the caller decides.
  void RunOnce(bool bRefindFailures = false) {
    if (mbResetRequested || !mMap.IsGood()) return;  // CHECK_RESET (MapMaker.cc:85): a requested reset is the caller's to carry out
    if (!mbBundleConverged_Recent && QueueSize() == 0) BundleAdjustRecent();
    if (mbBundleConverged_Recent && QueueSize() == 0) ReFindNewlyMade();
    if (mbBundleConverged_Recent && !mbBundleConverged_Full && QueueSize() == 0) BundleAdjustAll();
    if (mbBundleConverged_Recent && mbBundleConverged_Full && bRefindFailures && QueueSize() == 0) ReFindFromFailureQueue();
    HandleBadPoints();
    if (QueueSize() > 0) AddKeyFrameFromTopOfQueue();
  }

  // MapMaker.cc:767-782: every keyframe, every point
  void BundleAdjustAll() {
    std::set<KeyFrame*> adjust, fixed;
    for (KeyFrame* kf : mMap.vpKeyFrames) (kf->bFixed ? fixed : adjust).insert(kf);
    std::set<MapPoint*> points(mMap.vpPoints.begin(), mMap.vpPoints.end());
    BundleAdjust(adjust, fixed, points, false);
  }

  // MapMaker.cc:787-829: the newest keyframe and its four nearest neighbours, the points they measure, and
  // every other keyframe measuring one of those points as a fixed camera
  void BundleAdjustRecent() {
    if (mMap.vpKeyFrames.size() < 8) { mbBundleConverged_Recent = true; return; }
    std::set<KeyFrame*> adjust;
    KeyFrame* newest = mMap.vpKeyFrames.back();
    adjust.insert(newest);
    for (KeyFrame* kf : NClosestKeyFrames(*newest, 4))
      if (!kf->bFixed) adjust.insert(kf);
    std::set<MapPoint*> points;
    for (KeyFrame* kf : adjust)
      for (auto& pm : kf->mMeasurements) points.insert(pm.first);
    std::set<KeyFrame*> fixed;
    for (KeyFrame* kf : mMap.vpKeyFrames) {
      if (adjust.count(kf)) continue;
      for (auto& pm : kf->mMeasurements)
        if (points.count(pm.first)) { fixed.insert(kf); break; }
    }
    BundleAdjust(adjust, fixed, points, true);
  }

  // MapMaker.cc:838-933
  void BundleAdjust(std::set<KeyFrame*> sAdjustSet, std::set<KeyFrame*> sFixedSet, std::set<MapPoint*> sMapPoints, bool bRecent) {
    Bundle b(mCamera, mnDevice, mbHaveParams ? &mParams : nullptr);
    mbBundleRunning = true;
    mbBundleRunningIsRecent = bRecent;
    // bundle id <-> map object; ids are handed out in insertion order, so two vectors and two maps suffice
    std::vector<KeyFrame*> view_of;
    std::vector<MapPoint*> point_of;
    std::map<KeyFrame*, int> id_of_view;
    std::map<MapPoint*, int> id_of_point;
    auto add_view = [&](KeyFrame* kf, bool fixed) {
      const int id = b.AddCamera(kf->se3CfromW, fixed);
      id_of_view[kf] = id;
      if ((int)view_of.size() <= id) view_of.resize(id + 1, nullptr);
      view_of[id] = kf;
    };
    for (KeyFrame* kf : sAdjustSet) add_view(kf, kf->bFixed);  // adjustable ones first, then the fixed ones
    for (KeyFrame* kf : sFixedSet) add_view(kf, true);
    for (MapPoint* p : sMapPoints) {
      const int id = b.AddPoint(p->v3WorldPos);
      id_of_point[p] = id;
      if ((int)point_of.size() <= id) point_of.resize(id + 1, nullptr);
      point_of[id] = p;
    }
    // measurements of the chosen points in the chosen keyframes, sigma^2 = (2^level)^2 (MapMaker.cc:879-880)
    for (KeyFrame* kf : mMap.vpKeyFrames) {
      auto v = id_of_view.find(kf);
      if (v == id_of_view.end()) continue;
      for (auto& pm : kf->mMeasurements) {
        auto p = id_of_point.find(pm.first);
        if (p == id_of_point.end()) continue;
        const double s = (double)Level::LevelScale(pm.second.nLevel);
        b.AddMeas(v->second, p->second, pm.second.v2RootPos, s * s);
      }
    }
    if (b.NumMeasurements() == 0) {  // nothing to adjust (a sparse map): the reference would assert in FindSigmaSquared (Tools.h:155)
      mbBundleRunning = false;
      mbBundleAbortRequested = false;
      if (bRecent) mbBundleConverged_Recent = true; else mbBundleConverged_Full = true;
      return;
    }
    const int nAccepted = b.Compute(&mbBundleAbortRequested);
    if (nAccepted < 0) {
      // A negative count from the reference's Bundle means "the adjustment blew up": ditch the map
      // (MapMaker.cc:887-892).  The library's negative codes are API / CUDA / NCCL errors instead: those are the
      // caller's problem, not the map's.
      mbBundleRunning = false;
      mbBundleAbortRequested = false;
      throw std::runtime_error(std::string("Bundle::Compute failed: ") + b.LastError());
    }
    if (nAccepted > 0) {
      for (auto& pi : id_of_point) pi.first->v3WorldPos = b.GetPoint(pi.second);
      for (auto& vi : id_of_view) vi.first->se3CfromW = b.GetCamera(vi.second);
      if (bRecent) mbBundleConverged_Recent = false;
      mbBundleConverged_Full = false;
      // the reference's tracker reads the live map; here it holds a device copy, refreshed on a revision change
      mMap.nRevision++;
    }
    if (b.Converged()) {
      mbBundleConverged_Recent = true;
      if (!bRecent) mbBundleConverged_Full = true;
    }
    mbBundleRunning = false;
    mbBundleAbortRequested = false;
    // outlier measurements (MapMaker.cc:914-932): a point whose root measurement is an outlier, or that would be
    // left with too few measurements, is bad; otherwise the measurement goes, and is retried or blacklisted
    // depending on where it came from
    for (auto& pc : b.GetOutlierMeasurements()) {
      MapPoint* pp = point_of[pc.first];
      KeyFrame* pk = view_of[pc.second];
      Measurement& m = pk->mMeasurements[pp];
      MapMakerData& md = MMData(pp);
      if (md.GoodMeasCount() <= 2 || m.Source == Measurement::SRC_ROOT) {
        pp->bBad = true;
      } else {
        if (m.Source == Measurement::SRC_TRACKER || m.Source == Measurement::SRC_EPIPOLAR) mvFailureQueue.emplace_back(pk, pp);
        else md.sNeverRetryKFs.insert(pk);
        pk->mMeasurements.erase(pp);
        md.sMeasurementKFs.erase(pk);
      }
    }
  }

  // state the reference keeps as (mostly private) members, MapMaker.h:118-158
  bool mbBundleConverged_Full = true, mbBundleConverged_Recent = true;
  bool mbBundleRunning = false, mbBundleRunningIsRecent = false;
  bool mbBundleAbortRequested = false, mbResetRequested = false;
  // set by Tracker::SetMapMaker: brings MapPoint::nMEstimatorOutlierCount / InlierCount up to date from the device
  std::function<void()> mfnRefreshCounters;
  std::vector<std::pair<KeyFrame*, MapPoint*> > mvFailureQueue;
  std::vector<MapPoint*> mvpNewQueue;  // mqNewQueue (MapMaker.h:130): points waiting to be re-found in older keyframes
  std::vector<KeyFrame*> mvpKeyFrameQueue;   // keyframes from the tracker waiting to be processed (MapMaker.h:128)
  double mdCandidateMinSTScore = 70.0;       // MapMaker.CandidateMinShiTomasiScore (KeyFrame.cc:63)
  double mdWiggleScaleDepthNormalized = 0.1; // mdWiggleScale / the first keyframe's scene depth (MapMaker.cc:392)
  double mdMaxKFDistWiggleMult = 0.05;       // MapMaker.MaxKFDistWiggleMult (MapMaker.cc:760)
  double mdWiggleScale = 0.1;          // MapMaker.WiggleScale (MapMaker.cc:225), the stereo baseline in map units

 private:
  // the map maker's own data-association handle (the reference gives it its own PatchFinder, MapMaker.cc:977),
  // created with this camera on first use
  ptam_tracker* Assoc(CVD::ImageRef size) {
    if (!mpAssoc) {
      double p[5];
      for (int i = 0; i < 5; i++) p[i] = mCamera.GetParams()[i];
      mpAssoc = ptam_tracker_create(mnDevice, p, size.x, size.y, 1, nullptr);
      if (!mpAssoc) throw std::runtime_error(std::string("ptam_tracker_create: ") + ptam_global_last_error());
    }
    return mpAssoc;
  }
  // the whole map as the handle's point set (source keyframes go to its keyframe store on first sight)
  void SyncAssocMap(ptam_tracker* t) {
    if (mbAssocMapValid && mnAssocRevision == mMap.nRevision && mnAssocPoints == mMap.vpPoints.size()) return;
    const size_t n = mMap.vpPoints.size();
    std::vector<double> world(3 * n), right(3 * n), down(3 * n);
    std::vector<int32_t> kf(n), lvl(n), ctr(2 * n);
    for (size_t i = 0; i < n; i++) {
      const MapPoint& p = *mMap.vpPoints[i];
      auto it = mStoreId.find(p.pPatchSourceKF);
      if (it == mStoreId.end()) {
        Level& L0 = p.pPatchSourceKF->aLevels[0];
        const int id = ptam_tracker_add_keyframe(t, L0.im.data(), L0.im.row_stride());
        if (id < 0) throw std::runtime_error(ptam_tracker_last_error(t));
        it = mStoreId.emplace(p.pPatchSourceKF, id).first;
      }
      for (int c = 0; c < 3; c++) { world[3 * i + c] = p.v3WorldPos[c]; right[3 * i + c] = p.v3PixelRight_W[c]; down[3 * i + c] = p.v3PixelDown_W[c]; }
      kf[i] = it->second; lvl[i] = p.nSourceLevel; ctr[2 * i] = p.irCenter.x; ctr[2 * i + 1] = p.irCenter.y;
    }
    if (ptam_tracker_set_map(t, 0, (int)n, world.data(), right.data(), down.data(), kf.data(), lvl.data(), ctr.data()) != PTAM_OK)
      throw std::runtime_error(ptam_tracker_last_error(t));
    mnAssocRevision = mMap.nRevision; mnAssocPoints = n; mbAssocMapValid = true;
  }
  ptam_tracker* mpAssoc = nullptr;
  unsigned mnAssocRevision = 0;
  size_t mnAssocPoints = 0;
  bool mbAssocMapValid = false;
  std::map<KeyFrame*, int> mStoreId;  // keyframes already resident in the handle's keyframe store
  std::vector<std::unique_ptr<MapPoint> > mOwnedPoints;
  std::vector<std::unique_ptr<KeyFrame> > mOwnedKeyFrames;
  Map& mMap;
  ATANCamera mCamera;
  int mnDevice;
  ptam_bundle_params mParams{};
  bool mbHaveParams = false;
  std::map<MapPoint*, MapMakerData> mMMData;
};

}  // namespace ptam_b200
