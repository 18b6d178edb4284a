// Host mirror of class Tracker (reference include/Tracker.h:155-241) for the measured path:
// TrackFrame(CVD::Image<byte>&, bool) = MakeKeyFrame_Lite + PredictPoseWithMotionModel + TrackMap +
// UpdateMotionModel + AssessTrackingQuality (src/Tracker.cc:86-188,442-698,1012-1107), all on the
// device through ptam_tracker_track_frames; GetCurrentPose() as in Tracker.h:163.
// With SetMapMaker: the keyframe hand-over heuristic and the distance branch of AssessTrackingQuality; the
// relocaliser runs on the device once every stored keyframe has a pose (LastResult().recovery reports it).
// Not mirrored (out of scope): trail tracking / stereo initialisation, GUI commands, GL drawing.
#pragma once
#include <map>
#include "KeyFrame.h"
#include "MapMaker.h"

namespace ptam_b200 {

class Tracker {
 public:
  // reference: Tracker(ImageRef irVideoSize, const ATANCamera &c, Map &m, MapMaker &mm)
  Tracker(CVD::ImageRef irVideoSize, const ATANCamera& c, Map& m, int device = 0, const ptam_tracker_params* params = nullptr)
      : mMap(m), mCamera(c), mirSize(irVideoSize) {
    double p[5];
    for (int i = 0; i < 5; i++) p[i] = c.GetParams()[i];
    h = ptam_tracker_create(device, p, irVideoSize.x, irVideoSize.y, 1, params);
    if (!h) throw std::runtime_error(std::string("ptam_tracker_create: ") + ptam_global_last_error());
  }
  ~Tracker() { if (h) ptam_tracker_destroy(h); }
  Tracker(const Tracker&) = delete;
  Tracker& operator=(const Tracker&) = delete;

  // bDraw is accepted for signature compatibility; there is no GL on this path.
  void TrackFrame(CVD::Image<CVD::byte>& imFrame, bool bDraw) {
    (void)bDraw;
    if (imFrame.size() != mirSize) throw std::invalid_argument("frame size differs from irVideoSize");
    SyncMap();
    const uint8_t* ptrs[1] = {imFrame.data()};
    if (ptam_tracker_track_frames(h, ptrs, imFrame.row_stride(), &mLast) != PTAM_OK) throw std::runtime_error(ptam_tracker_last_error(h));
    mse3CamFromWorld = se3_from_array(mLast.se3_cam_from_world);
    mCurrentKF.se3CfromW = mse3CamFromWorld;  // Tracker.cc:662
    mCurrentKF.dSceneDepthMean = mLast.scene_depth_mean;
    mCurrentKF.dSceneDepthSigma = mLast.scene_depth_sigma;
    mbKFCurrent = false;
    mnFrame++;
    if (mpMapMaker && mMap.IsGood() && mLast.recovery == 0) ConsultMapMaker();
  }

  // reference: the fourth constructor argument (MapMaker &mm).  With a map maker attached TrackFrame also does
  // the two things the reference's tracker asks it for: the last branch of AssessTrackingQuality
  // (Tracker.cc:1094-1099) and the keyframe hand-over heuristic (Tracker.cc:146-166).
  void SetMapMaker(MapMaker* mm) {
    mpMapMaker = mm;
    if (mm) mm->mfnRefreshCounters = [this]() { RefreshPointCounters(); };
  }
  // MapPoint::nMEstimatorOutlierCount / InlierCount are lifetime counters in the reference (Tracker.cc:990-997
  // increments the map's own objects).  The device counts from zero after every map upload, so the host keeps
  // the counts at upload time and adds the device's since then.
  void RefreshPointCounters() {
    const size_t n = mvUploaded.size();
    if (!n || !mbMapUploaded) return;
    std::vector<int32_t> outl(n), inl(n);
    if (ptam_tracker_get_points(h, 0, nullptr, nullptr, nullptr, nullptr, outl.data(), inl.data()) < 0) throw std::runtime_error(ptam_tracker_last_error(h));
    for (size_t i = 0; i < n; i++) {
      mvUploaded[i]->nMEstimatorOutlierCount = mvBaseOutl[i] + outl[i];
      mvUploaded[i]->nMEstimatorInlierCount = mvBaseInl[i] + inl[i];
    }
  }
  TooN::SE3<> GetCurrentPose() { return mse3CamFromWorld; }

  // Sets mse3CamFromWorld (and zero velocity) — what the reference does after stereo initialisation
  // (Tracker.cc:332-335) or recovery.
  void SetCurrentPose(const TooN::SE3<>& se3) {
    ptam_tracker_state st;
    ptam_tracker_get_state(h, 0, &st);
    se3_to_array(se3, st.se3_cam_from_world);
    for (double& v : st.velocity) v = 0;
    st.msd_scaled_velocity_magnitude = 0;
    ptam_tracker_set_state(h, 0, &st);
    mse3CamFromWorld = se3;
  }

  // mCurrentKF as the reference leaves it after TrackFrame: levels with corners (MakeKeyFrame_Lite)
  // and the measurements of the found points (Tracker.cc:665-677) + per-point M-estimator counters
  // (Tracker.cc:990-997).  Fetched lazily: it is only needed when the frame becomes a keyframe.
  KeyFrame& CurrentKeyFrame() {
    if (mbKFCurrent) return mCurrentKF;
    for (int l = 0; l < LEVELS; l++) mCurrentKF.FetchLevel(h, 0, l);
    const size_t n = mvUploaded.size();
    std::vector<int32_t> flags(n), level(n);
    std::vector<double> found(2 * n);
    if (n) ptam_tracker_get_points(h, 0, flags.data(), level.data(), found.data(), nullptr, nullptr, nullptr);
    RefreshPointCounters();
    mCurrentKF.mMeasurements.clear();
    for (size_t i = 0; i < n; i++) {
      if (!(flags[i] & PTAM_PT_FOUND)) continue;
      Measurement m;
      m.nLevel = level[i];
      m.bSubPix = (flags[i] & PTAM_PT_SUBPIX) != 0;
      m.v2RootPos = TooN::makeVector(found[2 * i], found[2 * i + 1]);
      m.Source = Measurement::SRC_TRACKER;
      mCurrentKF.mMeasurements[mvUploaded[i]] = m;
    }
    mbKFCurrent = true;
    return mCurrentKF;
  }
  const ptam_track_result& LastResult() const { return mLast; }
  ptam_tracker* handle() { return h; }

 private:
  // Upload the map when its revision changed: new source keyframes go to the device keyframe store
  // (their level-0 pixels; the library rebuilds the pyramid), points as SoA.
  void SyncMap() {
    if (mbMapUploaded && mnRevision == mMap.nRevision && mvUploaded.size() == mMap.vpPoints.size()) return;
    // the upload below zeroes the device's per-point counters: bank what they hold first.  Points that left the
    // map are in its trash (Map::MoveBadPointsToTrash), still valid objects until their owner empties it.
    RefreshPointCounters();
    const size_t n = mMap.vpPoints.size();
    std::vector<double> world(3 * n), right(3 * n), down(3 * n);
    std::vector<int32_t> kf(n), lvl(n), ctr(2 * n);
    for (size_t i = 0; i < n; i++) {
      MapPoint& p = *mMap.vpPoints[i];
      auto it = mKFIds.find(p.pPatchSourceKF);
      if (it == mKFIds.end()) {
        Level& L0 = p.pPatchSourceKF->aLevels[0];
        const int id = ptam_tracker_add_keyframe(h, L0.im.data(), L0.im.row_stride());
        if (id < 0) throw std::runtime_error(ptam_tracker_last_error(h));
        it = mKFIds.emplace(p.pPatchSourceKF, id).first;
      }
      for (int k = 0; k < 3; k++) { world[3 * i + k] = p.v3WorldPos[k]; right[3 * i + k] = p.v3PixelRight_W[k]; down[3 * i + k] = p.v3PixelDown_W[k]; }
      kf[i] = it->second; lvl[i] = p.nSourceLevel; ctr[2 * i] = p.irCenter.x; ctr[2 * i + 1] = p.irCenter.y;
    }
    if (ptam_tracker_set_map(h, 0, (int)n, world.data(), right.data(), down.data(), kf.data(), lvl.data(), ctr.data()) != PTAM_OK)
      throw std::runtime_error(ptam_tracker_last_error(h));
    // every keyframe of the map with its current pose: what the relocaliser compares a lost frame against
    // (Relocaliser.cc:12-38 walks mMap.vpKeyFrames); bundle adjustment moves the poses, a revision bump brings them here
    for (KeyFrame* k : mMap.vpKeyFrames) {
      auto it = mKFIds.find(k);
      if (it == mKFIds.end()) {
        if (k->aLevels[0].im.size() != mirSize) continue;   // not made yet: the relocaliser stays off until it is (LastResult().recovery stays 0)
        const int id = ptam_tracker_add_keyframe(h, k->aLevels[0].im.data(), k->aLevels[0].im.row_stride());
        if (id < 0) throw std::runtime_error(ptam_tracker_last_error(h));
        it = mKFIds.emplace(k, id).first;
      }
      double pose[12];
      se3_to_array(k->se3CfromW, pose);
      if (ptam_tracker_set_keyframe_pose(h, it->second, pose) != PTAM_OK) throw std::runtime_error(ptam_tracker_last_error(h));
    }
    mvUploaded = mMap.vpPoints;
    mvBaseOutl.resize(n); mvBaseInl.resize(n);
    for (size_t i = 0; i < n; i++) { mvBaseOutl[i] = mvUploaded[i]->nMEstimatorOutlierCount; mvBaseInl[i] = mvUploaded[i]->nMEstimatorInlierCount; }
    mnRevision = mMap.nRevision;
    mbMapUploaded = true;
  }

  void ConsultMapMaker() {
    // AssessTrackingQuality left the quality as it was because the found fractions were inconclusive: far from
    // every keyframe means lost (the device cannot know: the keyframe poses live with the map maker)
    if (mLast.quality_needs_kf_distance && !mMap.vpKeyFrames.empty()) {
      KeyFrame* closest = mpMapMaker->ClosestKeyFrame(mCurrentKF);
      if (MapMaker::KeyFrameLinearDist(mCurrentKF, *closest) > mpMapMaker->GetWiggleScale() * 10.0) {
        ptam_tracker_state st;
        if (ptam_tracker_get_state(h, 0, &st) != PTAM_OK) throw std::runtime_error(ptam_tracker_last_error(h));
        if (st.tracking_quality != 0) {   // BAD now: the first lost frame (mnLostFrames was reset while the quality was not BAD)
          st.tracking_quality = 0;
          st.lost_frames = 1;
          if (ptam_tracker_set_state(h, 0, &st) != PTAM_OK) throw std::runtime_error(ptam_tracker_last_error(h));
        }
        mLast.tracking_quality = 0;
      }
    }
    // a good, well-separated frame becomes a keyframe unless the map maker is already behind (isNeedFrame is
    // computed but not used by the reference, Tracker.cc:160)
    if (mLast.tracking_quality == 2 && mnFrame - mnLastKeyFrameDropped > 20 && mpMapMaker->QueueSize() < 3) {
      mpMapMaker->AddKeyFrame(CurrentKeyFrame());
      mnLastKeyFrameDropped = mnFrame;
    }
  }

  Map& mMap;
  ATANCamera mCamera;
  CVD::ImageRef mirSize;
  ptam_tracker* h = nullptr;
  MapMaker* mpMapMaker = nullptr;
  int mnFrame = 0, mnLastKeyFrameDropped = -20;   // Tracker.cc:62-63
  KeyFrame mCurrentKF;
  TooN::SE3<> mse3CamFromWorld;
  ptam_track_result mLast{};
  std::map<KeyFrame*, int> mKFIds;
  std::vector<MapPoint*> mvUploaded;
  std::vector<int> mvBaseOutl, mvBaseInl;  // the points' lifetime counts when they were uploaded
  unsigned mnRevision = 0;
  bool mbMapUploaded = false, mbKFCurrent = false;
};

}  // namespace ptam_b200
