// Host mirror of class PatchFinder (reference include/PatchFinder.h:54-98, src/PatchFinder.cc) over the unit
// entry points of the CUDA library (ptam_patch_search_batch / ptam_patch_get_results, include/ptam_b200.h).
//
// The reference's object works on ONE map point at a time, five steps in a row:
//   CalcSearchLevelAndWarpMatrix -> MakeTemplateCoarseCont -> FindPatchCoarse -> MakeSubPixTemplate ->
//   IterateSubPixToConvergence.
// The device runs those steps for whole batches (that is what the tracker and the map maker use: Tracker.h,
// MapMaker.h here).  This class keeps the reference's method-by-method surface for callers written against it
// and for unit parity: every method forwards to the device with a one-point map and returns what the reference's
// getter would; nothing is computed on the host.  Each step is one small device call, so use it where the
// reference's call pattern must be kept, not for throughput — SearchBatch below is the throughput form.
//
// Differences forced by the boundary, all visible in the signatures:
//  * the constructor also takes the camera and the image size (the reference's PatchFinder receives the camera
//    derivatives from its caller; the device evaluates ATANCamera::GetProjectionDerivs itself at the same pose,
//    so the m2CamDerivs argument of CalcSearchLevelAndWarpMatrix is accepted and not read);
//  * FindPatchCoarse searches around ir(v2Image) of the pose given to step 1 — what every caller in the
//    reference passes (Tracker.cc:881, MapMaker.cc:993); another centre is rejected;
//  * MakeTemplateCoarseNoWarp / SetSubPixPos (epipolar search of the map maker) live in
//    ptam_tracker_epipolar_search (MapMaker.h: AddPointsEpipolar), not here.
#pragma once
#include <map>
#include "KeyFrame.h"

namespace ptam_b200 {

class PatchFinder {
 public:
  int mnMaxSSD;  // PatchFinder.cc:18-19: the device's kMaxSSD is the same 8 * 8 * 500

  PatchFinder(const ATANCamera& cam, CVD::ImageRef irSize, int nPatchSize = 8, int device = 0) : mirSize(irSize) {
    if (nPatchSize != 8) throw std::invalid_argument("the device PatchFinder is the reference's default 8x8");
    mnMaxSSD = 8 * 8 * 500;
    double p[5];
    for (int i = 0; i < 5; i++) p[i] = cam.GetParams()[i];
    h = ptam_tracker_create(device, p, irSize.x, irSize.y, 1, nullptr);
    if (!h) throw std::runtime_error(std::string("ptam_tracker_create: ") + ptam_global_last_error());
  }
  ~PatchFinder() { if (h) ptam_tracker_destroy(h); }
  PatchFinder(const PatchFinder&) = delete;
  PatchFinder& operator=(const PatchFinder&) = delete;

  // Step 1 (PatchFinder.cc:52-84).  Returns mnSearchLevel, or -1 for an inappropriate warp / a point that does not
  // project into the image (the reference's callers test the projection before calling).
  int CalcSearchLevelAndWarpMatrix(MapPoint& p, TooN::SE3<> se3CFromW, TooN::Matrix<2>& m2CamDerivs) {
    (void)m2CamDerivs;
    SetPoint(p);
    se3_to_array(se3CFromW, mPose);
    mbHavePose = true;
    mnRange = 0; mnSubPix = 0;
    Run(nullptr);
    return mnSearchLevel;
  }
  int GetLevel() { return mnSearchLevel; }
  const TooN::Matrix<2>& GetWarpInverse() const { return mm2WarpInverse; }  // mm2WarpInverse (protected in the reference)

  // Step 2 (PatchFinder.cc:98-127): the template was made by the device call of step 1 (same warp: the device's
  // per-point cache holds it); this fetches it.
  void MakeTemplateCoarseCont(MapPoint& p) {
    if (&p != mpPoint || !mbHavePose) throw std::logic_error("MakeTemplateCoarseCont: call CalcSearchLevelAndWarpMatrix for this point first");
    mimTemplate.resize(CVD::ImageRef(8, 8));
    int32_t sums[2] = {0, 0};
    if (ptam_tracker_get_templates(h, 0, mimTemplate.data(), sums) < 0) throw std::runtime_error(ptam_tracker_last_error(h));
    mnTemplateSum = sums[0]; mnTemplateSumSq = sums[1];
  }
  bool TemplateBad() { return mbTemplateBad; }
  const CVD::Image<CVD::byte>& GetTemplate() const { return mimTemplate; }
  int GetTemplateSum() const { return mnTemplateSum; }
  int GetTemplateSumSq() const { return mnTemplateSumSq; }

  // Step 3 (PatchFinder.cc:160-211)
  bool FindPatchCoarse(CVD::ImageRef ir, KeyFrame& kf, unsigned int nRange) {
    if (!mbHavePose) throw std::logic_error("FindPatchCoarse: call CalcSearchLevelAndWarpMatrix first");
    mnRange = nRange; mnSubPix = 0;
    Run(&kf);
    if (mnSearchLevel >= 0 && (ir.x != mirPredicted.x || ir.y != mirPredicted.y))
      throw std::invalid_argument("FindPatchCoarse: the device searches around ir(v2Image) of the pose given to step 1");
    mv2CoarsePos = mv2Pos;
    return mbFound;
  }
  CVD::ImageRef GetCoarsePos() { return CVD::ImageRef((int)mv2CoarsePos[0], (int)mv2CoarsePos[1]); }
  TooN::Vector<2> GetCoarsePosAsVector() { return mv2CoarsePos; }

  // Steps 4 + 5 (PatchFinder.cc:219-318): the sub-pixel template is built on the device in the call that iterates
  void MakeSubPixTemplate() {}
  bool IterateSubPixToConvergence(KeyFrame& kf, int nMaxIts) {
    if (!mbHavePose) throw std::logic_error("IterateSubPixToConvergence: call steps 1-3 first");
    if (nMaxIts <= 0) return false;
    mnSubPix = nMaxIts;
    Run(&kf);
    if (mbFound) mv2SubPixPos = mv2Pos;
    return mbFound && mbSubPix;
  }
  TooN::Vector<2> GetSubPixPos() { return mv2SubPixPos; }
  TooN::Matrix<2> GetCov() {  // PatchFinder.h:92: an appropriately scaled identity
    TooN::Matrix<2> m;
    m(0, 0) = m(1, 1) = (double)Level::LevelScale(mnSearchLevel); m(0, 1) = m(1, 0) = 0.0;
    return m;
  }

  // Throughput form: all points against one keyframe in one device call (what Tracker::SearchForPoints does for
  // its lists, Tracker.cc:867-912).  Outputs may be null; found positions are sub-pixel where nSubPixIts > 0.
  void SearchBatch(const std::vector<MapPoint*>& vpPoints, KeyFrame& kf, const TooN::SE3<>& se3CFromW, unsigned nRange, int nSubPixIts,
                   std::vector<int>* pvLevel, std::vector<char>* pvFound, std::vector<TooN::Vector<2> >* pvPos) {
    UploadPoints(vpPoints);
    mpPoint = nullptr; mbHavePose = false;
    SetFrame(kf);
    double pose[12];
    se3_to_array(se3CFromW, pose);
    if (ptam_patch_search_batch(h, pose, nRange, nSubPixIts) != PTAM_OK) throw std::runtime_error(ptam_tracker_last_error(h));
    const size_t n = vpPoints.size();
    std::vector<int32_t> level(n), found(n);
    std::vector<double> pos(2 * n);
    if (n && ptam_patch_get_results(h, 0, level.data(), nullptr, nullptr, found.data(), pos.data(), nullptr) < 0)
      throw std::runtime_error(ptam_tracker_last_error(h));
    if (pvLevel) pvLevel->assign(level.begin(), level.end());
    if (pvFound) pvFound->assign(found.begin(), found.end());
    if (pvPos) { pvPos->resize(n); for (size_t i = 0; i < n; i++) (*pvPos)[i] = TooN::makeVector(pos[2 * i], pos[2 * i + 1]); }
  }

 private:
  void UploadPoints(const std::vector<MapPoint*>& pts) {
    const size_t n = pts.size();
    std::vector<double> world(3 * n), right(3 * n), down(3 * n);
    std::vector<int32_t> kf(n), lvl(n), ctr(2 * n);
    for (size_t i = 0; i < n; i++) {
      MapPoint& p = *pts[i];
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
  }
  void SetPoint(MapPoint& p) {
    if (&p == mpPoint) return;
    UploadPoints(std::vector<MapPoint*>(1, &p));  // a new point: the device's template cache starts empty, as a new PatchFinder's
    mpPoint = &p;
  }
  void SetFrame(KeyFrame& kf) {
    if (&kf == mpFrame) return;
    CVD::Image<CVD::byte>& im = kf.aLevels[0].im;
    if (im.size() != mirSize) throw std::invalid_argument("keyframe size differs from the PatchFinder's image size");
    const uint8_t* ptrs[1] = {im.data()};
    if (ptam_tracker_make_keyframes(h, ptrs, im.row_stride()) != PTAM_OK) throw std::runtime_error(ptam_tracker_last_error(h));
    mpFrame = &kf;
  }
  void Run(KeyFrame* kf) {
    if (kf) SetFrame(*kf);
    else if (!mpFrame) {  // step 1 needs no image, the device call does: any frame of the right size will do
      mBlank.aLevels[0].im.resize(mirSize);
      std::memset(mBlank.aLevels[0].im.data(), 0, (size_t)mirSize.x * mirSize.y);
      SetFrame(mBlank);
    }
    if (ptam_patch_search_batch(h, mPose, mnRange, mnSubPix) != PTAM_OK) throw std::runtime_error(ptam_tracker_last_error(h));
    int32_t level = -1, bad = 0, found = 0, sub = 0;
    double wi[4] = {0, 0, 0, 0}, pos[2] = {0, 0}, v2im[2] = {0, 0};
    if (ptam_patch_get_results(h, 0, &level, wi, &bad, &found, pos, &sub) < 0) throw std::runtime_error(ptam_tracker_last_error(h));
    ptam_tracker_get_points(h, 0, nullptr, nullptr, nullptr, v2im, nullptr, nullptr);
    mnSearchLevel = level; mbTemplateBad = bad != 0; mbFound = found != 0; mbSubPix = sub != 0;
    mm2WarpInverse(0, 0) = wi[0]; mm2WarpInverse(0, 1) = wi[1]; mm2WarpInverse(1, 0) = wi[2]; mm2WarpInverse(1, 1) = wi[3];
    mv2Pos = TooN::makeVector(pos[0], pos[1]);
    mirPredicted = CVD::ImageRef((int)v2im[0], (int)v2im[1]);
  }

  ptam_tracker* h = nullptr;
  CVD::ImageRef mirSize;
  std::map<KeyFrame*, int> mKFIds;
  MapPoint* mpPoint = nullptr;
  KeyFrame* mpFrame = nullptr;
  KeyFrame mBlank;
  double mPose[12];
  bool mbHavePose = false;
  unsigned mnRange = 0;
  int mnSubPix = 0;
  int mnSearchLevel = -1;
  bool mbTemplateBad = false, mbFound = false, mbSubPix = false;
  TooN::Matrix<2> mm2WarpInverse;
  TooN::Vector<2> mv2Pos, mv2CoarsePos, mv2SubPixPos;
  CVD::ImageRef mirPredicted;
  CVD::Image<CVD::byte> mimTemplate;
  int mnTemplateSum = 0, mnTemplateSumSq = 0;
};

}  // namespace ptam_b200
