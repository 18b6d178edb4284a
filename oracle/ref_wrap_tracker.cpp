// TEST INFRASTRUCTURE — C ABI around the reference's OWN Tracker / MapMaker / KeyFrame / PatchFinder /
// SmallBlurryImage classes (src/Tracker.cc, MapMaker.cc, KeyFrame.cc, PatchFinder.cc, ImageProcess.cc,
// Map.cc, Relocaliser.cc, ATANCamera.cc ... compiled in place from /root/reference against the header
// stand-ins in oracle/shim/), with the signatures of the oracle (orc_tracker_*) and the product
// (ptam_tracker_*), prefix ref_.  tests/test_ref_pin_tracker.py drives the oracle and this library with
// the same frames, maps and states and compares.  Nothing here ships.
#include "Tracker.h"
#include "MapMaker.h"
#include "../include/ptam_b200.h"
#include <gvars3/instances.h>
#include <cstring>
#include <functional>
#include <memory>
#include <string>

namespace {
std::function<void()> g_wait_hook;

struct RefMapMaker : public MapMaker {
  RefMapMaker(Map& m, const ATANCamera& c) : MapMaker(m, c) { mdWiggleScale = 1e30; mdWiggleScaleDepthNormalized = 1e30; }
  using MapMaker::mbResetRequested;
  using MapMaker::mdWiggleScale;
  using MapMaker::mvpKeyFrameQueue;
  void ResetNow() { if (mbResetRequested) Reset(); }
  int ReFindAllIn(KeyFrame& k) { return ReFindInSingleKeyFrame(k); }
  bool Epipolar(KeyFrame& src, KeyFrame& tgt, int level, int cand) { return AddPointEpipolar(src, tgt, level, cand); }
  void TopOfQueue() { AddKeyFrameFromTopOfQueue(); }
  using MapMaker::mbBundleConverged_Full;
  using MapMaker::mbBundleConverged_Recent;
  using MapMaker::mbBundleRunning;
  using MapMaker::mvFailureQueue;
  void AdjustAll() { BundleAdjustAll(); }
  void AdjustRecent() { BundleAdjustRecent(); }
};
struct RefTracker : public Tracker {
  RefTracker(CVD::ImageRef sz, const ATANCamera& c, Map& m, MapMaker& mm) : Tracker(sz, c, m, mm) {}
  using Tracker::manMeasAttempted;
  using Tracker::manMeasFound;
  using Tracker::mbDidCoarse;
  using Tracker::mbJustRecoveredSoUseCoarse;
  using Tracker::mCurrentKF;
  using Tracker::mdMSDScaledVelocityMagnitude;
  using Tracker::mnFrame;
  using Tracker::mnLostFrames;
  using Tracker::mpSBIThisFrame;
  using Tracker::mse3CamFromWorld;
  using Tracker::mv6CameraVelocity;
  using Tracker::mCamera;
  using Tracker::SearchForPoints;
  using Tracker::CalcPoseUpdate;
  struct RelocPeek : public Relocaliser { using Relocaliser::mnBest; };
  int reloc_best() { return static_cast<RelocPeek&>(mRelocaliser).mnBest; }
  void set_quality(int q) { mTrackingQuality = q == 0 ? BAD : GOOD; }
  int get_quality() const { return mTrackingQuality == BAD ? 0 : 2; }
};
// SmallBlurryImage::mirSize is a process-wide static fixed by the first keyframe (ImageProcess.cc:281-282):
// a handle of another resolution has to re-arm it
struct SbiSize : public SmallBlurryImage { static void rearm() { mirSize = CVD::ImageRef(-1, -1); } };
// the reference's PatchFinder keeps its working state protected: read it through a derived class
struct PeekFinder : public PatchFinder {
  using PatchFinder::mimTemplate;
  using PatchFinder::mm2WarpInverse;
  using PatchFinder::mnTemplateSum;
  using PatchFinder::mnTemplateSumSq;
  using PatchFinder::mpLastTemplateMapPoint;
};
struct Stream {
  std::vector<TrackerData*> unit_set;      // what the last ref_patch_search_batch searched
  std::vector<char> unit_projected;        // per map point: projected into the frame in that call
  bool unit_mode = false;
  Map map;
  std::unique_ptr<RefMapMaker> mm;
  std::unique_ptr<RefTracker> trk;
  std::unique_ptr<KeyFrame> refind_kf;
  ptam_track_result res;
  bool refind_mode = false;
  bool keep_queue = false;   // test hook: let the tracker's keyframes pile up in MapMaker::mvpKeyFrameQueue
};
struct Handle {
  int W = 0, H = 0, S = 0;
  std::unique_ptr<ATANCamera> cam;
  std::vector<KeyFrame*> store;
  std::vector<std::unique_ptr<Stream>> streams;
  std::string err;
};

TooN::SE3<> se3_from12(const double* p) {
  TooN::Matrix<3> R;
  for (int i = 0; i < 9; i++) R(i / 3, i % 3) = p[i];
  return TooN::SE3<>(TooN::SE3<>::raw(R), TooN::makeVector(p[9], p[10], p[11]));
}
void se3_to12(const TooN::SE3<>& s, double* p) {
  for (int i = 0; i < 9; i++) p[i] = s.get_rotation().get_matrix()(i / 3, i % 3);
  for (int i = 0; i < 3; i++) p[9 + i] = s.get_translation()[i];
}
CVD::Image<CVD::byte> wrap_image(const uint8_t* im, int w, int h, int stride) {
  CVD::Image<CVD::byte> out(CVD::ImageRef(w, h));
  for (int y = 0; y < h; y++) std::memcpy(out[y], im + (size_t)y * stride, w);
  return out;
}
const char* kEstimators[3] = {"Tukey", "Cauchy", "Huber"};
}  // namespace

#include <execinfo.h>
#include <csignal>
static void segv_handler(int) { void* bt[64]; int n = backtrace(bt, 64); backtrace_symbols_fd(bt, n, 2); _exit(139); }
extern "C" void ref_debug_install_segv_handler() { signal(SIGSEGV, segv_handler); }
extern "C" int ptam_ref_usleep(unsigned int) { if (g_wait_hook) g_wait_hook(); return 0; }

extern "C" {

void ref_tracker_default_params(ptam_tracker_params* p) {
  p->coarse_min = 20; p->coarse_max = 60; p->coarse_range = 30; p->coarse_subpix_its = 8;
  p->disable_coarse = 0; p->max_patches_per_frame = 1000; p->mestimator = 0; p->use_constant_velocity = 1;
  p->coarse_min_velocity = 0.006; p->quality_good = 0.3; p->quality_lost = 0.13;
  p->use_rotation_estimator = 1; p->reserved0 = 0; p->rotation_estimator_blur = 0.75;
}

void* ref_tracker_create(int, const double* cam_params, int width, int height, int n_streams, const ptam_tracker_params* params) {
  using GVars3::GV3;
  ptam_tracker_params p;
  if (params) p = *params; else ref_tracker_default_params(&p);
  TooN::Vector<5> cp;
  for (int i = 0; i < 5; i++) cp[i] = cam_params[i];
  GV3::set<TooN::Vector<5>>("Camera.Parameters", cp);
  // the GVars3 keys Tracker.cc reads (Tracker.cc:95-96,491-496,596,931,1040,1088-1089)
  GV3::set<unsigned int>("Tracker.CoarseMin", (unsigned)p.coarse_min);
  GV3::set<unsigned int>("Tracker.CoarseMax", (unsigned)p.coarse_max);
  GV3::set<unsigned int>("Tracker.CoarseRange", (unsigned)p.coarse_range);
  GV3::set<int>("Tracker.CoarseSubPixIts", p.coarse_subpix_its);
  GV3::set<int>("Tracker.DisableCoarse", p.disable_coarse);
  GV3::set<double>("Tracker.CoarseMinVelocity", p.coarse_min_velocity);
  GV3::set<int>("Tracker.MaxPatchesPerFrame", p.max_patches_per_frame);
  GV3::set<std::string>("Tracker.MEstimator", kEstimators[p.mestimator >= 0 && p.mestimator < 3 ? p.mestimator : 0]);
  GV3::set<int>("Tracker.UseConstantVelocity", p.use_constant_velocity);
  GV3::set<double>("Tracker.TrackingQualityGood", p.quality_good);
  GV3::set<double>("Tracker.TrackingQualityLost", p.quality_lost);
  GV3::set<int>("Tracker.UseRotationEstimator", p.use_rotation_estimator);
  GV3::set<double>("Tracker.RotationEstimatorBlur", p.rotation_estimator_blur);
  GV3::set<int>("Tracker.DrawFASTCorners", 0);
  SbiSize::rearm();
  Handle* h = new Handle;
  h->W = width; h->H = height; h->S = n_streams;
  h->cam.reset(new ATANCamera("Camera"));
  h->cam->SetImageSize(TooN::makeVector((double)width, (double)height));
  for (int s = 0; s < n_streams; s++) {
    std::unique_ptr<Stream> st(new Stream);
    std::memset(&st->res, 0, sizeof st->res);
    st->mm.reset(new RefMapMaker(st->map, *h->cam));
    RefMapMaker* mm = st->mm.get();
    g_wait_hook = [mm]() { mm->ResetNow(); };  // Tracker::Reset waits for the map-maker thread (Tracker.cc:71-74)
    st->trk.reset(new RefTracker(CVD::ImageRef(width, height), *h->cam, st->map, *st->mm));
    g_wait_hook = nullptr;
    st->trk->set_quality(2);
    h->streams.push_back(std::move(st));
  }
  return h;
}
void ref_tracker_destroy(void* hp) {
  Handle* h = (Handle*)hp;
  h->streams.clear();
  for (KeyFrame* k : h->store) delete k;
  delete h;
}
const char* ref_tracker_last_error(const void* h) { return ((const Handle*)h)->err.c_str(); }

int ref_tracker_add_keyframe(void* hp, const uint8_t* image, int stride) {
  Handle* h = (Handle*)hp;
  CVD::Image<CVD::byte> im = wrap_image(image, h->W, h->H, stride);
  KeyFrame* k = new KeyFrame;
  k->bFixed = false;
  k->dSceneDepthMean = 1.0; k->dSceneDepthSigma = 1.0;
  k->MakeKeyFrame_Lite(im);
  k->MakeKeyFrame_Rest();  // the relocaliser compares against KeyFrame::pSBI (KeyFrame.cc:80-81)
  h->store.push_back(k);
  for (auto& s : h->streams) s->map.vpKeyFrames = h->store;
  return (int)h->store.size() - 1;
}

int ref_tracker_set_map(void* hp, int stream, int n, const double* world, const double* right, const double* down,
                        const int32_t* src_kf, const int32_t* src_level, const int32_t* center) {
  Handle* h = (Handle*)hp;
  if (stream < 0 || stream >= h->S) return PTAM_ERR_INVALID;
  Stream& s = *h->streams[stream];
  for (MapPoint* p : s.map.vpPoints) { delete p->pTData; delete p->pMMData; delete p; }
  s.map.vpPoints.clear();
  for (int i = 0; i < n; i++) {
    if (src_kf[i] < 0 || src_kf[i] >= (int)h->store.size() || src_level[i] < 0 || src_level[i] >= LEVELS) return PTAM_ERR_INVALID;
    MapPoint* p = new MapPoint;
    p->v3WorldPos = TooN::makeVector(world[3 * i], world[3 * i + 1], world[3 * i + 2]);
    p->v3PixelRight_W = TooN::makeVector(right[3 * i], right[3 * i + 1], right[3 * i + 2]);
    p->v3PixelDown_W = TooN::makeVector(down[3 * i], down[3 * i + 1], down[3 * i + 2]);
    p->pPatchSourceKF = h->store[src_kf[i]];
    p->nSourceLevel = src_level[i];
    p->irCenter = CVD::ImageRef(center[2 * i], center[2 * i + 1]);
    p->pMMData = new MapMakerData;
    // the tracker allocates TrackerData lazily and leaves its flags uninitialised until the point first
    // enters the image (Tracker.cc:457-458, Tracker.h:41-61): allocate it here with defined values
    p->pTData = new TrackerData(p);
    p->pTData->bInImage = p->pTData->bPotentiallyVisible = p->pTData->bSearched = p->pTData->bFound = p->pTData->bDidSubPix = false;
    p->pTData->nSearchLevel = -1;
    p->pTData->v2Found = TooN::Zeros; p->pTData->v2Image = TooN::Zeros;
    s.map.vpPoints.push_back(p);
  }
  s.map.vpKeyFrames = h->store;
  s.map.bGood = true;
  return 0;
}

int ref_tracker_set_keyframe_pose(void* hp, int kf, const double* se3) {
  Handle* h = (Handle*)hp;
  if (kf < 0 || kf >= (int)h->store.size()) return PTAM_ERR_INVALID;
  h->store[kf]->se3CfromW = se3_from12(se3);
  return PTAM_OK;
}
int ref_tracker_set_state(void* hp, int stream, const ptam_tracker_state* st) {
  Handle* h = (Handle*)hp;
  if (stream < 0 || stream >= h->S) return PTAM_ERR_INVALID;
  RefTracker& t = *h->streams[stream]->trk;
  t.mse3CamFromWorld = se3_from12(st->se3_cam_from_world);
  for (int k = 0; k < 6; k++) t.mv6CameraVelocity[k] = st->velocity[k];
  t.mdMSDScaledVelocityMagnitude = st->msd_scaled_velocity_magnitude;
  t.mCurrentKF.dSceneDepthMean = st->scene_depth_mean;
  t.mCurrentKF.dSceneDepthSigma = st->scene_depth_sigma;
  t.mbJustRecoveredSoUseCoarse = st->just_recovered_so_use_coarse != 0;
  t.set_quality(st->tracking_quality);
  t.mnLostFrames = st->lost_frames;
  t.mnFrame = st->frame;
  return 0;
}
int ref_tracker_get_state(void* hp, int stream, ptam_tracker_state* st) {
  Handle* h = (Handle*)hp;
  if (stream < 0 || stream >= h->S) return PTAM_ERR_INVALID;
  RefTracker& t = *h->streams[stream]->trk;
  std::memset(st, 0, sizeof *st);
  se3_to12(t.mse3CamFromWorld, st->se3_cam_from_world);
  for (int k = 0; k < 6; k++) st->velocity[k] = t.mv6CameraVelocity[k];
  st->msd_scaled_velocity_magnitude = t.mdMSDScaledVelocityMagnitude;
  st->scene_depth_mean = t.mCurrentKF.dSceneDepthMean;
  st->scene_depth_sigma = t.mCurrentKF.dSceneDepthSigma;
  st->just_recovered_so_use_coarse = t.mbJustRecoveredSoUseCoarse;
  st->tracking_quality = t.get_quality();
  st->lost_frames = t.mnLostFrames;
  st->frame = t.mnFrame;
  return 0;
}
int ref_tracker_make_keyframes(void* hp, const uint8_t* const* images, int stride) {
  Handle* h = (Handle*)hp;
  for (int s = 0; s < h->S; s++) {
    CVD::Image<CVD::byte> im = wrap_image(images[s], h->W, h->H, stride);
    h->streams[s]->trk->mCurrentKF.MakeKeyFrame_Lite(im);
    h->streams[s]->refind_mode = false;
  }
  return 0;
}
int ref_tracker_track_frames(void* hp, const uint8_t* const* images, int stride, ptam_track_result* results) {
  Handle* h = (Handle*)hp;
  for (int s = 0; s < h->S; s++) {
    Stream& st = *h->streams[s];
    RefTracker& t = *st.trk;
    st.refind_mode = false;
    if (!st.keep_queue) st.mm->mvpKeyFrameQueue.clear();  // keyframes the tracker hands to the (absent) map-maker thread are dropped
    CVD::Image<CVD::byte> im = wrap_image(images[s], h->W, h->H, stride);
    const int lost_before = t.mnLostFrames;
    t.TrackFrame(im, false);
    ptam_track_result& r = st.res;
    std::memset(&r, 0, sizeof r);
    r.reloc_keyframe = -1;
    if (lost_before >= 3) {  // the recovery branch ran (Tracker.cc:170-178); AssessTrackingQuality moves mnLostFrames only on success
      r.recovery = t.mnLostFrames != lost_before ? 1 : 2;
      r.reloc_keyframe = t.reloc_best();
    }
    se3_to12(t.mse3CamFromWorld, r.se3_cam_from_world);
    r.scene_depth_mean = t.mCurrentKF.dSceneDepthMean; r.scene_depth_sigma = t.mCurrentKF.dSceneDepthSigma;
    for (int l = 0; l < LEVELS; l++) {
      r.meas_attempted[l] = t.manMeasAttempted[l]; r.meas_found[l] = t.manMeasFound[l];
      r.n_corners[l] = (int)t.mCurrentKF.aLevels[l].vCorners.size();
    }
    r.did_coarse = t.mbDidCoarse;
    r.tracking_quality = t.get_quality();
    if (results) results[s] = r;
  }
  return 0;
}
int ref_tracker_synchronize(void*) { return 0; }
int ref_tracker_level_size(const void* hp, int level, int* w, int* h) {
  const Handle* t = (const Handle*)hp;
  int ww = t->W, hh = t->H;
  for (int l = 0; l < level; l++) { ww /= 2; hh /= 2; }
  *w = ww; *h = hh; return 0;
}
static KeyFrame& current_kf(Stream& s) { return s.refind_mode ? *s.refind_kf : s.trk->mCurrentKF; }
int ref_tracker_get_level(void* hp, int stream, int level, uint8_t* pixels, int32_t* corners_xy, int cap, int32_t* row_lut) {
  Handle* h = (Handle*)hp;
  Level& L = current_kf(*h->streams[stream]).aLevels[level];
  const CVD::ImageRef sz = L.im.size();
  if (pixels) for (int y = 0; y < sz.y; y++) std::memcpy(pixels + (size_t)y * sz.x, L.im[y], sz.x);
  if (corners_xy)
    for (size_t i = 0; i < L.vCorners.size() && (int)i < cap; i++) { corners_xy[2 * i] = L.vCorners[i].x; corners_xy[2 * i + 1] = L.vCorners[i].y; }
  if (row_lut) for (size_t i = 0; i < L.vCornerRowLUT.size(); i++) row_lut[i] = L.vCornerRowLUT[i];
  return (int)L.vCorners.size();
}
// MapMaker::ReFindInSingleKeyFrame (MapMaker.cc:1028-1040) on a keyframe made from images[s] with pose se3 + 12 s
int ref_tracker_refind_in_keyframes(void* hp, const uint8_t* const* images, int stride, const double* se3) {
  Handle* h = (Handle*)hp;
  for (int s = 0; s < h->S; s++) {
    Stream& st = *h->streams[s];
    st.refind_kf.reset(new KeyFrame);
    st.refind_kf->bFixed = false;
    CVD::Image<CVD::byte> im = wrap_image(images[s], h->W, h->H, stride);
    st.refind_kf->MakeKeyFrame_Lite(im);
    st.refind_kf->se3CfromW = se3_from12(se3 + 12 * s);
    for (MapPoint* p : st.map.vpPoints) { p->pMMData->sMeasurementKFs.clear(); p->pMMData->sNeverRetryKFs.clear(); }
    st.mm->ReFindAllIn(*st.refind_kf);
    st.refind_mode = true;
  }
  return PTAM_OK;
}
// MapMaker::AddPointEpipolar (MapMaker.cc:529-688) for every candidate; source = stored keyframe, target =
// the stream's current frame.  The new MapPoint the reference creates is read (sub-pixel target
// position = kTarget.mMeasurements[pNew].v2RootPos, triangulated v3WorldPos) and removed again.
// NB the reference caches UnProject of every pixel in a function-local static sized by the first call.
// Test hook (not part of the shared ABI): the stream's map maker as the reference's tracker sees it — sets
// MapMaker::mdWiggleScale (wiggle_scale >= 0; the wrapper otherwise keeps it at 1e30 so that the distance branch of
// Tracker::AssessTrackingQuality never fires), optionally drops the keyframes the tracker queued (clear_queue: 1 once,
// 2 once and keep the queue across frames from now on, 3 back to dropping it every frame), returns MapMaker::QueueSize().
int ref_tracker_mapmaker_ctl(void* hp, int stream, double wiggle_scale, int clear_queue) {
  Handle* h = (Handle*)hp;
  if (stream < 0 || stream >= h->S) return PTAM_ERR_INVALID;
  Stream& st = *h->streams[stream];
  if (wiggle_scale >= 0) st.mm->mdWiggleScale = wiggle_scale;
  if (clear_queue) { for (KeyFrame* k : st.mm->mvpKeyFrameQueue) delete k; st.mm->mvpKeyFrameQueue.clear(); }
  if (clear_queue == 2) st.keep_queue = true;    // ... and from now on keep what the tracker queues
  if (clear_queue == 3) st.keep_queue = false;
  return st.mm->QueueSize();
}
// Test hook (not part of the shared ABI): the reference's own MapMaker::BundleAdjustAll / BundleAdjustRecent
// (MapMaker.cc:767-933) on a map given as arrays: keyframes (pose, fixed), points, measurements (keyframe, point, root
// position, level, Source).  Keyframes and points live in arrays, so pointer order = index order, as in
// ptam_cg_b200/host/mapmaker_check.cc.  Outputs in that program's layout.  Returns the failure-queue length, or < 0.
int ref_mapmaker_bundle_adjust(void* hp, int stream, int mode, int max_iterations, int C, const double* cams, const int32_t* fixed, int P,
                               const double* pts, int M, const int32_t* mcam, const int32_t* mpt, const double* uv, const int32_t* level,
                               const int32_t* src, double* out_pts, double* out_cams, int32_t* out_bad, int32_t* out_nmeas,
                               int32_t* out_queue, int32_t* out_never, int never_cap, int32_t* out_n_never, int32_t* out_flags) {
  Handle* h = (Handle*)hp;
  if (stream < 0 || stream >= h->S) return PTAM_ERR_INVALID;
  Stream& st = *h->streams[stream];
  GVars3::GV3::set<int>("Bundle.MaxIterations", max_iterations);   // the Bundle keys are process-wide: pin them all
  GVars3::GV3::set<double>("Bundle.UpdateSquaredConvergenceLimit", 1e-6);
  GVars3::GV3::set<double>("Bundle.MinTukeySigma", 0.4);
  GVars3::GV3::set<std::string>("Bundle.MEstimator", "Tukey");
  GVars3::GV3::set<int>("Bundle.Cout", 0);
  std::unique_ptr<KeyFrame[]> kfs(new KeyFrame[C]);
  std::unique_ptr<MapPoint[]> points(new MapPoint[P]);
  std::vector<KeyFrame*> saved_kfs = st.map.vpKeyFrames;
  std::vector<MapPoint*> saved_pts = st.map.vpPoints;
  st.map.vpKeyFrames.clear(); st.map.vpPoints.clear();
  for (int c = 0; c < C; c++) { kfs[c].se3CfromW = se3_from12(cams + 12 * c); kfs[c].bFixed = fixed[c] != 0; st.map.vpKeyFrames.push_back(&kfs[c]); }
  for (int p = 0; p < P; p++) {
    points[p].v3WorldPos = TooN::makeVector(pts[3 * p], pts[3 * p + 1], pts[3 * p + 2]);
    points[p].pMMData = new MapMakerData;
    st.map.vpPoints.push_back(&points[p]);
  }
  for (int m = 0; m < M; m++) {
    Measurement me;
    me.nLevel = level[m]; me.bSubPix = false; me.v2RootPos = TooN::makeVector(uv[2 * m], uv[2 * m + 1]);
    me.Source = static_cast<decltype(me.Source)>(src[m]);
    kfs[mcam[m]].mMeasurements[&points[mpt[m]]] = me;
    points[mpt[m]].pMMData->sMeasurementKFs.insert(&kfs[mcam[m]]);
  }
  st.mm->mbBundleConverged_Full = true; st.mm->mbBundleConverged_Recent = true;
  st.mm->mvFailureQueue.clear();
  if (mode == 0) st.mm->AdjustAll(); else st.mm->AdjustRecent();
  int n_never = 0;
  for (int p = 0; p < P; p++) {
    for (int k = 0; k < 3; k++) out_pts[3 * p + k] = points[p].v3WorldPos[k];
    out_bad[p] = points[p].bBad ? 1 : 0;
    for (KeyFrame* kf : points[p].pMMData->sNeverRetryKFs)
      if (n_never < never_cap) { out_never[2 * n_never] = (int)(kf - kfs.get()); out_never[2 * n_never + 1] = p; n_never++; }
  }
  *out_n_never = n_never;
  for (int c = 0; c < C; c++) { se3_to12(kfs[c].se3CfromW, out_cams + 12 * c); out_nmeas[c] = (int)kfs[c].mMeasurements.size(); }
  const int nq = (int)st.mm->mvFailureQueue.size();
  for (int q = 0; q < nq; q++) { out_queue[2 * q] = (int)(st.mm->mvFailureQueue[q].first - kfs.get()); out_queue[2 * q + 1] = (int)(st.mm->mvFailureQueue[q].second - points.get()); }
  out_flags[0] = st.mm->mbBundleConverged_Full; out_flags[1] = st.mm->mbBundleConverged_Recent;
  out_flags[2] = st.mm->mbResetRequested; out_flags[3] = st.mm->mbBundleRunning;
  st.mm->mvFailureQueue.clear();
  for (int p = 0; p < P; p++) delete points[p].pMMData;
  st.map.vpKeyFrames = saved_kfs; st.map.vpPoints = saved_pts;
  GVars3::GV3::set<int>("Bundle.MaxIterations", 20);
  return nq;
}
// Test hook (not part of the shared ABI): the reference's own MapMaker::AddKeyFrame + AddKeyFrameFromTopOfQueue
// (MapMaker.cc:480-519: MakeKeyFrame_Rest, ReFindInSingleKeyFrame, ThinCandidates / ClosestKeyFrame / AddPointEpipolar on
// levels 3, 0, 1, 2) on the stream's map.  The stored keyframes must have poses (ref_tracker_set_keyframe_pose) and the
// map must be set; every existing point gets its root measurement in its source keyframe first.  The keyframe from the
// "tracker" = image, pose, scene depth and n_meas measurements (point, level | position).  Outputs, in the layout of
// ptam_cg_b200/host/mapmaker_check.cc (addkf): per old point (has measurement, Source, level, in sMeasurementKFs,
// in sNeverRetryKFs) + position; the thinned candidate lists (count, then x y pairs, per level); new points per level
// and the index of the closest stored keyframe.
int ref_mapmaker_add_keyframe(void* hp, int stream, const uint8_t* image, int stride, const double* se3, double depth_mean,
                              double depth_sigma, double wiggle, int n_meas, const int32_t* meas_pt_level, const double* meas_pos,
                              int32_t* out_meas, double* out_pos, int32_t* out_cand, int cand_cap, int32_t* out_new) {
  Handle* h = (Handle*)hp;
  if (stream < 0 || stream >= h->S) return PTAM_ERR_INVALID;
  Stream& st = *h->streams[stream];
  const size_t n_old = st.map.vpPoints.size();
  for (KeyFrame* k : h->store) k->mMeasurements.clear();
  for (MapPoint* p : st.map.vpPoints) {
    p->pMMData->sMeasurementKFs.clear(); p->pMMData->sNeverRetryKFs.clear();
    Measurement root;
    root.nLevel = p->nSourceLevel; root.bSubPix = true; root.Source = Measurement::SRC_ROOT;
    root.v2RootPos = Level::LevelZeroPos(p->irCenter, p->nSourceLevel);
    p->pPatchSourceKF->mMeasurements[p] = root;
    p->pMMData->sMeasurementKFs.insert(p->pPatchSourceKF);
  }
  st.map.vpKeyFrames = h->store;
  GVars3::GV3::set<double>("MapMaker.CandidateMinShiTomasiScore", 70.0);
  KeyFrame from_tracker;
  from_tracker.bFixed = false;
  CVD::Image<CVD::byte> im = wrap_image(image, h->W, h->H, stride);
  from_tracker.MakeKeyFrame_Lite(im);
  from_tracker.se3CfromW = se3_from12(se3);
  from_tracker.dSceneDepthMean = depth_mean; from_tracker.dSceneDepthSigma = depth_sigma;
  for (int j = 0; j < n_meas; j++) {
    Measurement m;
    m.nLevel = meas_pt_level[2 * j + 1]; m.bSubPix = m.nLevel > 0; m.Source = Measurement::SRC_REFIND;
    m.v2RootPos = TooN::makeVector(meas_pos[2 * j], meas_pos[2 * j + 1]);
    from_tracker.mMeasurements[st.map.vpPoints[meas_pt_level[2 * j]]] = m;
  }
  st.mm->mdWiggleScale = wiggle;
  st.mm->AddKeyFrame(from_tracker);
  st.mm->TopOfQueue();
  st.mm->mdWiggleScale = 1e30;
  KeyFrame& k = *st.map.vpKeyFrames.back();
  for (size_t i = 0; i < n_old; i++) {
    MapPoint* p = st.map.vpPoints[i];
    auto it = k.mMeasurements.find(p);
    const bool has = it != k.mMeasurements.end();
    out_meas[5 * i] = has; out_meas[5 * i + 1] = has ? (int)it->second.Source : -1; out_meas[5 * i + 2] = has ? it->second.nLevel : -1;
    out_meas[5 * i + 3] = (int)p->pMMData->sMeasurementKFs.count(&k); out_meas[5 * i + 4] = (int)p->pMMData->sNeverRetryKFs.count(&k);
    out_pos[2 * i] = has ? it->second.v2RootPos[0] : 0.0; out_pos[2 * i + 1] = has ? it->second.v2RootPos[1] : 0.0;
  }
  int o = 0;
  for (int l = 0; l < LEVELS; l++) {
    const auto& v = k.aLevels[l].vCandidates;
    if (o + 1 + 2 * (int)v.size() > cand_cap) return PTAM_ERR_CAPACITY;
    out_cand[o++] = (int)v.size();
    for (const auto& c : v) { out_cand[o++] = c.irLevelPos.x; out_cand[o++] = c.irLevelPos.y; }
  }
  for (int l = 0; l < LEVELS; l++) out_new[l] = 0;
  for (size_t i = n_old; i < st.map.vpPoints.size(); i++) out_new[st.map.vpPoints[i]->nSourceLevel]++;
  double best = 1e300; int closest = -1;
  for (size_t c = 0; c < h->store.size(); c++) {
    const TooN::Vector<3> d = h->store[c]->se3CfromW.inverse().get_translation() - k.se3CfromW.inverse().get_translation();
    if (std::sqrt(d * d) < best) { best = std::sqrt(d * d); closest = (int)c; }
  }
  out_new[4] = closest;
  // take the new keyframe and its points out again: the handle's other entry points expect the map as it was set
  for (size_t i = n_old; i < st.map.vpPoints.size(); i++) { MapPoint* p = st.map.vpPoints[i]; delete p->pMMData; delete p; }
  st.map.vpPoints.resize(n_old);
  for (MapPoint* p : st.map.vpPoints) { p->pMMData->sMeasurementKFs.clear(); p->pMMData->sNeverRetryKFs.clear(); }
  for (KeyFrame* kk : h->store) kk->mMeasurements.clear();
  st.map.vpKeyFrames = h->store;
  delete &k;
  return o;
}
// world position + pixel-right / pixel-down vectors (9 doubles) of every point the last ref_tracker_epipolar_search
// added, in candidate order: what MapMaker.cc:648-672 (Triangulate, RefreshPixelVectors) computed.  Test hook for
// the host mirror's MapMaker::AddPointsEpipolar; not part of the shared ABI.
static std::vector<double> g_last_epi_points;
int ref_tracker_epipolar_last_points(double* out, int cap_points) {
  const int n = (int)g_last_epi_points.size() / 9;
  if (out) std::memcpy(out, g_last_epi_points.data(), sizeof(double) * 9 * (size_t)std::min(n, cap_points));
  return n;
}
int ref_tracker_epipolar_search(void* hp, int stream, int level, int src_kf, const double* src_se3, double src_depth_mean,
                                double src_depth_sigma, const double* target_se3, double wiggle_scale, int n_cand,
                                const int32_t* cand_xy, int32_t* found, int32_t* best_corner, double* sub_pos) {
  Handle* h = (Handle*)hp;
  if (stream < 0 || stream >= h->S || level < 0 || level >= LEVELS || src_kf < 0 || src_kf >= (int)h->store.size()) return PTAM_ERR_INVALID;
  Stream& st = *h->streams[stream];
  KeyFrame& src = *h->store[src_kf];
  KeyFrame& tgt = current_kf(st);
  src.se3CfromW = se3_from12(src_se3);
  src.dSceneDepthMean = src_depth_mean; src.dSceneDepthSigma = src_depth_sigma;
  tgt.se3CfromW = se3_from12(target_se3);
  tgt.aLevels[level].bImplaneCornersCached = false;
  tgt.aLevels[level].vImplaneCorners.clear();
  st.mm->mdWiggleScale = wiggle_scale;
  std::vector<Candidate> saved = src.aLevels[level].vCandidates;
  src.aLevels[level].vCandidates.clear();
  g_last_epi_points.clear();
  for (int c = 0; c < n_cand; c++) {
    Candidate cd;
    cd.irLevelPos = CVD::ImageRef(cand_xy[2 * c], cand_xy[2 * c + 1]);
    cd.dSTScore = 0;
    src.aLevels[level].vCandidates.push_back(cd);
  }
  for (int c = 0; c < n_cand; c++) {
    found[c] = 0; best_corner[c] = -1; sub_pos[2 * c] = sub_pos[2 * c + 1] = 0;
    const size_t before = st.map.vpPoints.size();
    if (!st.mm->Epipolar(src, tgt, level, c)) continue;
    MapPoint* p = st.map.vpPoints.back();
    found[c] = 1;
    for (int k = 0; k < 3; k++) g_last_epi_points.push_back(p->v3WorldPos[k]);   // the point the reference's own
    for (int k = 0; k < 3; k++) g_last_epi_points.push_back(p->v3PixelRight_W[k]);  // Triangulate / RefreshPixelVectors made
    for (int k = 0; k < 3; k++) g_last_epi_points.push_back(p->v3PixelDown_W[k]);
    const Measurement& m = tgt.mMeasurements[p];
    sub_pos[2 * c] = m.v2RootPos[0]; sub_pos[2 * c + 1] = m.v2RootPos[1];
    tgt.mMeasurements.erase(p); src.mMeasurements.erase(p);
    st.map.vpPoints.resize(before);
    delete p->pMMData; delete p;
  }
  src.aLevels[level].vCandidates = saved;
  st.mm->mdWiggleScale = 1e30;
  return PTAM_OK;
}
int ref_tracker_keyframe_rest(void* hp, int stream, double min_shi_tomasi_score) {
  Handle* h = (Handle*)hp;
  if (stream < 0 || stream >= h->S) return PTAM_ERR_INVALID;
  GVars3::GV3::set<double>("MapMaker.CandidateMinShiTomasiScore", min_shi_tomasi_score);
  current_kf(*h->streams[stream]).MakeKeyFrame_Rest();
  return PTAM_OK;
}
int ref_tracker_get_level_rest(void* hp, int stream, int level, int32_t* max_xy, int max_cap, int32_t* cand_xy, double* cand_score,
                               int cand_cap, int* n_cand) {
  Handle* h = (Handle*)hp;
  if (stream < 0 || stream >= h->S || level < 0 || level >= LEVELS) return PTAM_ERR_INVALID;
  Level& L = current_kf(*h->streams[stream]).aLevels[level];
  if (max_xy)
    for (size_t i = 0; i < L.vMaxCorners.size() && (int)i < max_cap; i++) { max_xy[2 * i] = L.vMaxCorners[i].x; max_xy[2 * i + 1] = L.vMaxCorners[i].y; }
  for (size_t i = 0; i < L.vCandidates.size() && (int)i < cand_cap; i++) {
    if (cand_xy) { cand_xy[2 * i] = L.vCandidates[i].irLevelPos.x; cand_xy[2 * i + 1] = L.vCandidates[i].irLevelPos.y; }
    if (cand_score) cand_score[i] = L.vCandidates[i].dSTScore;
  }
  if (n_cand) *n_cand = (int)L.vCandidates.size();
  return (int)L.vMaxCorners.size();
}
int ref_tracker_get_points(void* hp, int stream, int32_t* flags, int32_t* level, double* v2_found, double* v2_image,
                           int32_t* outl, int32_t* inl) {
  Handle* h = (Handle*)hp;
  Stream& s = *h->streams[stream];
  for (size_t i = 0; i < s.map.vpPoints.size(); i++) {
    MapPoint* p = s.map.vpPoints[i];
    int f = 0, lv = -1;
    double fx = 0, fy = 0, ix = 0, iy = 0;
    if (s.refind_mode) {
      auto it = s.refind_kf->mMeasurements.find(p);
      if (it != s.refind_kf->mMeasurements.end()) {
        f |= PTAM_PT_FOUND | (it->second.bSubPix ? PTAM_PT_SUBPIX : 0);
        lv = it->second.nLevel; fx = it->second.v2RootPos[0]; fy = it->second.v2RootPos[1];
      }
    } else if (p->pTData) {
      // raw TrackerData fields.  The reference keeps no "entered the PVS this frame" flag
      // (bPotentiallyVisible is never set), and bSearched / bFound / v2Found are stale for points that
      // did not enter it: PTAM_PT_IN_PVS is never reported here, the test masks with the oracle's.
      const TrackerData& d = *p->pTData;
      if (d.bInImage) f |= PTAM_PT_IN_IMAGE;
      if (d.bSearched) f |= PTAM_PT_SEARCHED;
      if (d.bFound) f |= PTAM_PT_FOUND;
      if (d.bFound && d.bDidSubPix) f |= PTAM_PT_SUBPIX;
      lv = d.nSearchLevel;
      fx = d.v2Found[0]; fy = d.v2Found[1];
      ix = d.v2Image[0]; iy = d.v2Image[1];
    }
    if (flags) flags[i] = f;
    if (level) level[i] = lv;
    if (v2_found) { v2_found[2 * i] = fx; v2_found[2 * i + 1] = fy; }
    if (v2_image) { v2_image[2 * i] = ix; v2_image[2 * i + 1] = iy; }
    if (outl) outl[i] = p->nMEstimatorOutlierCount;
    if (inl) inl[i] = p->nMEstimatorInlierCount;
  }
  return (int)s.map.vpPoints.size();
}
// The unit entry points on the reference's OWN classes: TrackerData::Project, ATANCamera::GetProjectionDerivs,
// PatchFinder::CalcSearchLevelAndWarpMatrix as Tracker::TrackMap calls them (Tracker.cc:452-476), then the
// reference's own Tracker::SearchForPoints (Tracker.cc:867-912) on that list against mCurrentKF.
int ref_patch_search_batch(void* hp, const double* se3, unsigned range, int subpix_its) {
  Handle* h = (Handle*)hp;
  if (!se3 || subpix_its < 0) return PTAM_ERR_INVALID;
  for (int s = 0; s < h->S; s++) {
    Stream& st = *h->streams[s];
    RefTracker& t = *st.trk;
    st.refind_mode = false; st.unit_mode = true;
    TooN::SE3<> pose = se3_from12(se3 + 12 * s);
    st.unit_set.clear();
    st.unit_projected.assign(st.map.vpPoints.size(), 0);
    for (size_t i = 0; i < st.map.vpPoints.size(); i++) {
      MapPoint& p = *st.map.vpPoints[i];
      if (!p.pTData) p.pTData = new TrackerData(&p);
      TrackerData& TD = *p.pTData;
      TD.bSearched = TD.bFound = TD.bDidSubPix = false;
      TD.nSearchLevel = -1;
      TD.Project(pose, t.mCamera);
      if (!TD.bInImage) continue;
      st.unit_projected[i] = 1;
      TD.m2CamDerivs = t.mCamera.GetProjectionDerivs();
      TD.nSearchLevel = TD.Finder.CalcSearchLevelAndWarpMatrix(TD.Point, pose, TD.m2CamDerivs);
      if (TD.nSearchLevel == -1) continue;
      st.unit_set.push_back(&TD);
    }
    t.SearchForPoints(st.unit_set, range, subpix_its);
  }
  return PTAM_OK;
}
int ref_patch_get_results(void* hp, int stream, int32_t* level, double* warp_inverse, int32_t* template_bad, int32_t* found,
                          double* pos, int32_t* subpix_converged) {
  Handle* h = (Handle*)hp;
  if (stream < 0 || stream >= h->S) return PTAM_ERR_INVALID;
  Stream& st = *h->streams[stream];
  for (size_t i = 0; i < st.map.vpPoints.size(); i++) {
    MapPoint* p = st.map.vpPoints[i];
    const bool have = p->pTData != nullptr;
    const bool proj = have && i < st.unit_projected.size() && st.unit_projected[i];
    const bool fnd = have && p->pTData->nSearchLevel >= 0 && p->pTData->bFound;
    if (level) level[i] = have ? p->pTData->nSearchLevel : -1;
    if (warp_inverse) {
      for (int q = 0; q < 4; q++) warp_inverse[4 * i + q] = 0.0;
      if (proj) { const TooN::Matrix<2>& m = static_cast<PeekFinder&>(p->pTData->Finder).mm2WarpInverse; for (int q = 0; q < 4; q++) warp_inverse[4 * i + q] = m(q / 2, q % 2); }
    }
    if (template_bad) template_bad[i] = proj && p->pTData->Finder.TemplateBad() ? 1 : 0;  // mbTemplateBad is not initialised before the first use
    if (found) found[i] = fnd ? 1 : 0;
    if (pos) { pos[2 * i] = fnd ? p->pTData->v2Found[0] : 0.0; pos[2 * i + 1] = fnd ? p->pTData->v2Found[1] : 0.0; }
    if (subpix_converged) subpix_converged[i] = (fnd && p->pTData->bDidSubPix) ? 1 : 0;
  }
  return (int)st.map.vpPoints.size();
}
// CalcJacobian (Tracker.h:125-136) + the reference's own Tracker::CalcPoseUpdate (Tracker.cc:928-1005)
int ref_pose_update(void* hp, double override_sigma_squared, int mark_outliers, double* mu6, int32_t* n_found) {
  Handle* h = (Handle*)hp;
  for (int s = 0; s < h->S; s++) {
    Stream& st = *h->streams[s];
    int nf = 0;
    for (TrackerData* td : st.unit_set)
      if (td->bFound) { td->CalcJacobian(); nf++; }
    TooN::Vector<6> mu = st.trk->CalcPoseUpdate(st.unit_set, override_sigma_squared, mark_outliers != 0);
    if (mu6) for (int k = 0; k < 6; k++) mu6[6 * s + k] = mu[k];
    if (n_found) n_found[s] = nf;
  }
  return PTAM_OK;
}
// the coarse templates of the points' PatchFinders (protected in the reference: read through PeekFinder)
int ref_tracker_get_templates(void* hp, int stream, uint8_t* tmpl, int32_t* sums) {
  Handle* h = (Handle*)hp;
  if (stream < 0 || stream >= h->S) return PTAM_ERR_INVALID;
  Stream& st = *h->streams[stream];
  for (size_t i = 0; i < st.map.vpPoints.size(); i++) {
    MapPoint* p = st.map.vpPoints[i];
    bool has = false;
    if (p->pTData) {
      PeekFinder& f = static_cast<PeekFinder&>(p->pTData->Finder);
      has = f.mpLastTemplateMapPoint == p;  // a template has been made for this point (PatchFinder.cc:124)
      if (has) {
        if (tmpl) for (int y = 0; y < 8; y++) std::memcpy(tmpl + 64 * i + 8 * y, f.mimTemplate[y], 8);
        if (sums) { sums[2 * i] = f.mnTemplateSum; sums[2 * i + 1] = f.mnTemplateSumSq; }
      }
    }
    if (!has) { if (tmpl) std::memset(tmpl + 64 * i, 0, 64); if (sums) { sums[2 * i] = 0; sums[2 * i + 1] = 0; } }
  }
  return (int)st.map.vpPoints.size();
}
int ref_tracker_get_iteration_set(void*, int, int32_t*, int) { return PTAM_ERR_INVALID; }  // a local of TrackMap
int ref_tracker_get_sbi(void* hp, int stream, float* tmpl, int cap, double* rot3, double* score) {
  Handle* h = (Handle*)hp;
  if (stream < 0 || stream >= h->S) return PTAM_ERR_INVALID;
  SmallBlurryImage* sbi = h->streams[stream]->trk->mpSBIThisFrame;
  if (!sbi) return 0;
  const CVD::ImageRef sz = sbi->mimTemplate.size();
  if (tmpl) for (int y = 0, i = 0; y < sz.y; y++) for (int x = 0; x < sz.x && i < cap; x++, i++) tmpl[i] = sbi->mimTemplate[y][x];
  if (rot3) rot3[0] = rot3[1] = rot3[2] = 0;  // not kept by the reference (a temporary of PredictPoseWithMotionModel)
  if (score) *score = 0;
  return sz.x * sz.y;
}
}  // extern "C"
