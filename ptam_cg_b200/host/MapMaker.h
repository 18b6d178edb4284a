// Host mirror of the bundle-adjustment side of class MapMaker (reference include/MapMaker.h:38-160): the only
// caller of Bundle.  BundleAdjustAll / BundleAdjustRecent choose the keyframe and point sets
// (src/MapMaker.cc:767-836), BundleAdjust marshals them into a Bundle, runs it and writes the adjusted state
// and the outlier bookkeeping back into the map (src/MapMaker.cc:838-933).  Same member names, argument
// meaning and side effects as the reference; the arithmetic runs behind ptam_bundle_* (host/Bundle.h).
//
// Identity of the bundle ids matters for a drop-in: the reference walks std::set<KeyFrame*> /
// std::set<MapPoint*> (pointer order) for cameras and points, and mMap.vpKeyFrames (insertion order) with each
// keyframe's std::map<MapPoint*, Measurement> (pointer order) for the measurements; so does this class.
#pragma once
#include <algorithm>
#include <cmath>
#include <map>
#include <set>
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
    const int nAccepted = b.Compute(&mbBundleAbortRequested);
    if (nAccepted < 0) {  // the reference ditches the map (MapMaker.cc:887-892)
      mbResetRequested = true;
      return;
    }
    if (nAccepted > 0) {
      for (auto& pi : id_of_point) pi.first->v3WorldPos = b.GetPoint(pi.second);
      for (auto& vi : id_of_view) vi.first->se3CfromW = b.GetCamera(vi.second);
      if (bRecent) mbBundleConverged_Recent = false;
      mbBundleConverged_Full = false;
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
  std::vector<std::pair<KeyFrame*, MapPoint*> > mvFailureQueue;

 private:
  Map& mMap;
  ATANCamera mCamera;
  int mnDevice;
  ptam_bundle_params mParams{};
  bool mbHaveParams = false;
  std::map<MapPoint*, MapMakerData> mMMData;
};

}  // namespace ptam_b200
