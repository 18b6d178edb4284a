// Exercises the MapMaker host mirror (host/MapMaker.h: BundleAdjustAll / BundleAdjustRecent / BundleAdjust over
// the Bundle mirror) on a map read from raw arrays, and writes the map back: adjusted points and keyframe
// poses, bad flags, surviving measurement counts, the failure queue and the never-retry sets.
// Driven by tests/test_zz_host_mapmaker_gpu.py (CUDA library) and tests/test_host_mapmaker_cpu.py (the same
// source compiled against the CPU oracle's identical ABI, to check the marshalling without a GPU).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include "MapMaker.h"
#include "Tracker.h"

using namespace ptam_b200;
using namespace TooN;

template <class T>
static std::vector<T> rd(const std::string& dir, const char* name) {
  std::ifstream f(dir + "/" + name, std::ios::binary | std::ios::ate);
  if (!f) { std::cerr << "missing " << name << "\n"; std::exit(2); }
  const size_t bytes = (size_t)f.tellg();
  f.seekg(0);
  std::vector<T> v(bytes / sizeof(T));
  f.read(reinterpret_cast<char*>(v.data()), bytes);
  return v;
}
template <class T>
static void wr(const std::string& dir, const char* name, const std::vector<T>& v) {
  std::ofstream f(dir + "/" + name, std::ios::binary);
  f.write(reinterpret_cast<const char*>(v.data()), v.size() * sizeof(T));
}

// MapMaker::AddPointsEpipolar between two keyframes made from raw images: the new map points (world position,
// pixel-right / pixel-down vectors) and their two measurements, level by level.
static int run_epipolar(const std::string& dir) {
  auto dims = rd<int32_t>(dir, "epi_dims.i32");  // W, H
  const int W = dims[0], H = dims[1];
  auto src_im = rd<uint8_t>(dir, "epi_src.u8");
  auto tgt_im = rd<uint8_t>(dir, "epi_tgt.u8");
  auto src_pose = rd<double>(dir, "epi_src_pose.f64");
  auto tgt_pose = rd<double>(dir, "epi_tgt_pose.f64");
  auto depth = rd<double>(dir, "epi_depth.f64");  // source scene depth mean, sigma, wiggle scale
  ATANCamera cam("Camera", makeVector(1.0803, 1.43987, 0.519983, 0.548655, 0.244943), CVD::ImageRef(W, H));
  KeyFrame kSrc, kTgt;
  CVD::BasicImage<CVD::byte> is(src_im.data(), CVD::ImageRef(W, H)), it(tgt_im.data(), CVD::ImageRef(W, H));
  kSrc.MakeKeyFrame_Lite(is);
  kSrc.MakeKeyFrame_Rest();     // candidates (KeyFrame.cc:61-82)
  kTgt.MakeKeyFrame_Lite(it);
  kSrc.se3CfromW = se3_from_array(src_pose.data());
  kTgt.se3CfromW = se3_from_array(tgt_pose.data());
  kSrc.dSceneDepthMean = depth[0]; kSrc.dSceneDepthSigma = depth[1];
  Map map;
  map.vpKeyFrames = {&kSrc, &kTgt};
  MapMaker mm(map, cam);
  mm.mdWiggleScale = depth[2];
  std::vector<int32_t> counts, ncand;
  for (int l = 0; l < LEVELS; l++) {
    ncand.push_back((int32_t)kSrc.aLevels[l].vCandidates.size());
    counts.push_back(mm.AddPointsEpipolar(kSrc, kTgt, l));
  }
  std::vector<double> pts, meas;
  std::vector<int32_t> levels;
  for (MapPoint* p : map.vpPoints) {
    for (int k = 0; k < 3; k++) pts.push_back(p->v3WorldPos[k]);
    for (int k = 0; k < 3; k++) pts.push_back(p->v3PixelRight_W[k]);
    for (int k = 0; k < 3; k++) pts.push_back(p->v3PixelDown_W[k]);
    const Measurement& ms = kSrc.mMeasurements[p];
    const Measurement& mt = kTgt.mMeasurements[p];
    meas.insert(meas.end(), {ms.v2RootPos[0], ms.v2RootPos[1], mt.v2RootPos[0], mt.v2RootPos[1]});
    levels.push_back(p->nSourceLevel);
    if (ms.Source != Measurement::SRC_ROOT || mt.Source != Measurement::SRC_EPIPOLAR || p->pPatchSourceKF != &kSrc ||
        mm.MMData(p).GoodMeasCount() != 2) { std::cerr << "bad bookkeeping of a new point\n"; return 1; }
  }
  if (mm.mvpNewQueue.size() != map.vpPoints.size()) { std::cerr << "new-point queue out of step\n"; return 1; }
  // a third keyframe joins the map: MapMaker::ReFindNewlyMade looks for the new points in every keyframe (the two
  // they were made from already measure them)
  {
    std::ifstream third(dir + "/epi_third.u8", std::ios::binary);
    if (third) {
      auto im3 = rd<uint8_t>(dir, "epi_third.u8");
      auto pose3 = rd<double>(dir, "epi_third_pose.f64");
      static KeyFrame k3;
      CVD::BasicImage<CVD::byte> i3(im3.data(), CVD::ImageRef(W, H));
      k3.MakeKeyFrame_Lite(i3);
      k3.se3CfromW = se3_from_array(pose3.data());
      map.vpKeyFrames.push_back(&k3);
      const int nRefound = mm.ReFindNewlyMade();
      std::vector<int32_t> r3 = {nRefound, (int32_t)mm.mvpNewQueue.size()};
      std::vector<double> p3;
      for (MapPoint* p : map.vpPoints) {
        auto m3 = k3.mMeasurements.find(p);
        const bool has = m3 != k3.mMeasurements.end();
        r3.push_back(has ? 1 : 0); r3.push_back(has ? m3->second.nLevel : -1);
        r3.push_back((int32_t)mm.MMData(p).sNeverRetryKFs.count(&k3)); r3.push_back(mm.MMData(p).GoodMeasCount());
        p3.push_back(has ? m3->second.v2RootPos[0] : 0.0); p3.push_back(has ? m3->second.v2RootPos[1] : 0.0);
      }
      wr(dir, "epi_out_third.i32", r3); wr(dir, "epi_out_third_pos.f64", p3);
      std::printf("epipolar: %d of the new points re-found in the third keyframe\n", nRefound);
    }
  }
  wr(dir, "epi_out_counts.i32", counts); wr(dir, "epi_out_ncand.i32", ncand); wr(dir, "epi_out_points.f64", pts);
  wr(dir, "epi_out_meas.f64", meas); wr(dir, "epi_out_levels.i32", levels);
  std::printf("epipolar: %zu new points from %d + %d + %d + %d candidates\n", map.vpPoints.size(), ncand[0], ncand[1], ncand[2], ncand[3]);
  return 0;
}

// MapMaker::ReFindInSingleKeyFrame: a map (source keyframes, points) and a new keyframe; some points already have
// a measurement in it or are on its never-retry list.  Output per point: measurement (level, sub-pixel flag,
// root position) and never-retry flag.
static int run_refind(const std::string& dir) {
  auto dims = rd<int32_t>(dir, "trk_dims.i32");  // W, H, n_kf, n_pts
  const int W = dims[0], H = dims[1], nkf = dims[2], npts = dims[3];
  auto kfim = rd<uint8_t>(dir, "trk_kf.u8");
  auto world = rd<double>(dir, "trk_world.f64");
  auto right = rd<double>(dir, "trk_right.f64");
  auto down = rd<double>(dir, "trk_down.f64");
  auto skf = rd<int32_t>(dir, "trk_srckf.i32");
  auto slv = rd<int32_t>(dir, "trk_srclevel.i32");
  auto ctr = rd<int32_t>(dir, "trk_center.i32");
  auto newim = rd<uint8_t>(dir, "rf_image.u8");
  auto newpose = rd<double>(dir, "rf_pose.f64");
  auto pre = rd<int32_t>(dir, "rf_pre.i32");  // per point: 0 nothing, 1 already measured in k, 2 never retry in k
  ATANCamera cam("Camera", makeVector(1.0803, 1.43987, 0.519983, 0.548655, 0.244943), CVD::ImageRef(W, H));
  Map map;
  std::vector<KeyFrame> kfs(nkf);
  for (int k = 0; k < nkf; k++) {
    CVD::BasicImage<CVD::byte> im(kfim.data() + (size_t)k * W * H, CVD::ImageRef(W, H));
    kfs[k].MakeKeyFrame_Lite(im);
    map.vpKeyFrames.push_back(&kfs[k]);
  }
  std::vector<MapPoint> points(npts);
  for (int i = 0; i < npts; i++) {
    MapPoint& p = points[i];
    p.v3WorldPos = makeVector(world[3 * i], world[3 * i + 1], world[3 * i + 2]);
    p.v3PixelRight_W = makeVector(right[3 * i], right[3 * i + 1], right[3 * i + 2]);
    p.v3PixelDown_W = makeVector(down[3 * i], down[3 * i + 1], down[3 * i + 2]);
    p.pPatchSourceKF = &kfs[skf[i]];
    p.nSourceLevel = slv[i];
    p.irCenter = CVD::ImageRef(ctr[2 * i], ctr[2 * i + 1]);
    map.vpPoints.push_back(&p);
  }
  map.bGood = true; map.nRevision++;
  KeyFrame k;
  CVD::BasicImage<CVD::byte> im(newim.data(), CVD::ImageRef(W, H));
  k.MakeKeyFrame_Lite(im);
  k.se3CfromW = se3_from_array(newpose.data());
  MapMaker mm(map, cam);
  for (int i = 0; i < npts; i++) {
    if (pre[i] == 1) {
      Measurement m; m.nLevel = 0; m.bSubPix = false; m.v2RootPos = makeVector(-1.0, -1.0); m.Source = Measurement::SRC_TRACKER;
      k.mMeasurements[&points[i]] = m;
      mm.MMData(&points[i]).sMeasurementKFs.insert(&k);
    } else if (pre[i] == 2) {
      mm.MMData(&points[i]).sNeverRetryKFs.insert(&k);
    }
  }
  const int nFound = mm.ReFindInSingleKeyFrame(k);
  const int nAgain = mm.ReFindInSingleKeyFrame(k);  // everything is now measured or given up on: nothing changes
  // what MapMaker::BundleAdjust does to an outlier measurement that came from the tracker (MapMaker.cc:923-929),
  // for every fifth re-found point; MapMaker::ReFindFromFailureQueue must bring exactly those back, unchanged
  std::map<MapPoint*, Measurement> before;
  for (int i = 0; i < npts; i += 5) {
    auto it = k.mMeasurements.find(&points[i]);
    if (it == k.mMeasurements.end() || it->second.Source != Measurement::SRC_REFIND) continue;
    before[&points[i]] = it->second;
    mm.mvFailureQueue.emplace_back(&k, &points[i]);
    k.mMeasurements.erase(it);
    mm.MMData(&points[i]).sMeasurementKFs.erase(&k);
  }
  const int nQueued = (int)mm.mvFailureQueue.size();
  const int nSecondChance = mm.ReFindFromFailureQueue();
  int nSame = 0;
  for (auto& pm : before) {
    auto it = k.mMeasurements.find(pm.first);
    if (it != k.mMeasurements.end() && it->second.nLevel == pm.second.nLevel && it->second.v2RootPos[0] == pm.second.v2RootPos[0] &&
        it->second.v2RootPos[1] == pm.second.v2RootPos[1]) nSame++;
  }
  std::vector<int32_t> out;   // per point: has measurement, source, level, subpix, never retry
  std::vector<double> pos;
  for (int i = 0; i < npts; i++) {
    auto it = k.mMeasurements.find(&points[i]);
    const bool has = it != k.mMeasurements.end();
    out.insert(out.end(), {has ? 1 : 0, has ? (int32_t)it->second.Source : -1, has ? it->second.nLevel : -1, has && it->second.bSubPix ? 1 : 0,
                           (int32_t)mm.MMData(&points[i]).sNeverRetryKFs.count(&k)});
    pos.push_back(has ? it->second.v2RootPos[0] : 0.0); pos.push_back(has ? it->second.v2RootPos[1] : 0.0);
  }
  wr(dir, "rf_out_points.i32", out); wr(dir, "rf_out_pos.f64", pos); wr(dir, "rf_out_counts.i32", std::vector<int32_t>{nFound, nAgain, nQueued, nSecondChance, nSame, (int32_t)mm.mvFailureQueue.size()});
  std::printf("refind: %d new measurements among %d points (second pass %d)\n", nFound, npts, nAgain);
  return 0;
}

static int ClosestIndex(MapMaker& mm, KeyFrame& k, std::vector<KeyFrame>& kfs) { return (int)(mm.ClosestKeyFrame(k) - kfs.data()); }

// MapMaker::AddKeyFrame + AddKeyFrameFromTopOfQueue: a map of two keyframes and their points, and a keyframe
// from the tracker with some measurements.  Output: the thinned candidate lists the epipolar searches ran on,
// the measurements the keyframe ended up with, the new points.
static int run_add_keyframe(const std::string& dir) {
  auto dims = rd<int32_t>(dir, "trk_dims.i32");  // W, H, n_kf, n_pts
  const int W = dims[0], H = dims[1], nkf = dims[2], npts = dims[3];
  auto kfim = rd<uint8_t>(dir, "trk_kf.u8");
  auto kfpose = rd<double>(dir, "ak_kf_poses.f64");
  auto world = rd<double>(dir, "trk_world.f64");
  auto right = rd<double>(dir, "trk_right.f64");
  auto down = rd<double>(dir, "trk_down.f64");
  auto skf = rd<int32_t>(dir, "trk_srckf.i32");
  auto slv = rd<int32_t>(dir, "trk_srclevel.i32");
  auto ctr = rd<int32_t>(dir, "trk_center.i32");
  auto newim = rd<uint8_t>(dir, "rf_image.u8");
  auto newpose = rd<double>(dir, "rf_pose.f64");
  auto depth = rd<double>(dir, "ak_depth.f64");     // scene depth mean, sigma of the new keyframe; wiggle scale
  auto tm_idx = rd<int32_t>(dir, "ak_meas_idx.i32");  // the tracker's measurements in the new keyframe: point, level
  auto tm_pos = rd<double>(dir, "ak_meas_pos.f64");
  ATANCamera cam("Camera", makeVector(1.0803, 1.43987, 0.519983, 0.548655, 0.244943), CVD::ImageRef(W, H));
  Map map;
  std::vector<KeyFrame> kfs(nkf);
  std::vector<MapPoint> points(npts);
  MapMaker mm(map, cam);
  mm.mdWiggleScale = depth[2];
  for (int k = 0; k < nkf; k++) {
    CVD::BasicImage<CVD::byte> im(kfim.data() + (size_t)k * W * H, CVD::ImageRef(W, H));
    kfs[k].MakeKeyFrame_Lite(im);
    kfs[k].se3CfromW = se3_from_array(&kfpose[12 * k]);
    map.vpKeyFrames.push_back(&kfs[k]);
  }
  for (int i = 0; i < npts; i++) {
    MapPoint& p = points[i];
    p.v3WorldPos = makeVector(world[3 * i], world[3 * i + 1], world[3 * i + 2]);
    p.v3PixelRight_W = makeVector(right[3 * i], right[3 * i + 1], right[3 * i + 2]);
    p.v3PixelDown_W = makeVector(down[3 * i], down[3 * i + 1], down[3 * i + 2]);
    p.pPatchSourceKF = &kfs[skf[i]];
    p.nSourceLevel = slv[i];
    p.irCenter = CVD::ImageRef(ctr[2 * i], ctr[2 * i + 1]);
    map.vpPoints.push_back(&p);
    Measurement root;
    root.nLevel = slv[i]; root.bSubPix = true; root.Source = Measurement::SRC_ROOT;
    root.v2RootPos = Level::LevelZeroPos(p.irCenter, slv[i]);
    kfs[skf[i]].mMeasurements[&p] = root;
    mm.MMData(&p).sMeasurementKFs.insert(&kfs[skf[i]]);
  }
  map.bGood = true; map.nRevision++;
  KeyFrame fromTracker;
  CVD::BasicImage<CVD::byte> im(newim.data(), CVD::ImageRef(W, H));
  fromTracker.MakeKeyFrame_Lite(im);
  fromTracker.se3CfromW = se3_from_array(newpose.data());
  fromTracker.dSceneDepthMean = depth[0]; fromTracker.dSceneDepthSigma = depth[1];
  for (size_t j = 0; j < tm_idx.size() / 2; j++) {
    Measurement m;
    m.nLevel = tm_idx[2 * j + 1]; m.bSubPix = m.nLevel > 0; m.Source = Measurement::SRC_REFIND;  // the hand-over resets it
    m.v2RootPos = makeVector(tm_pos[2 * j], tm_pos[2 * j + 1]);
    fromTracker.mMeasurements[&points[tm_idx[2 * j]]] = m;
  }
  mm.mbBundleConverged_Full = mm.mbBundleConverged_Recent = true;
  mm.AddKeyFrame(fromTracker);
  if (mm.mvpKeyFrameQueue.size() != 1 || map.vpKeyFrames.size() != (size_t)nkf) { std::cerr << "AddKeyFrame must only queue\n"; return 1; }
  mm.AddKeyFrameFromTopOfQueue();
  if (!mm.mvpKeyFrameQueue.empty() || map.vpKeyFrames.size() != (size_t)nkf + 1 || mm.mbBundleConverged_Full || mm.mbBundleConverged_Recent) {
    std::cerr << "queue / map / convergence flags after AddKeyFrameFromTopOfQueue\n"; return 1;
  }
  KeyFrame& k = *map.vpKeyFrames.back();
  if (&k == &fromTracker) { std::cerr << "the keyframe must be copied\n"; return 1; }
  std::vector<int32_t> cand, meas, newpts;
  std::vector<double> meas_pos;
  for (int l = 0; l < LEVELS; l++) {
    cand.push_back((int32_t)k.aLevels[l].vCandidates.size());
    for (auto& c : k.aLevels[l].vCandidates) { cand.push_back(c.irLevelPos.x); cand.push_back(c.irLevelPos.y); }
  }
  for (int i = 0; i < npts; i++) {
    auto it = k.mMeasurements.find(&points[i]);
    const bool has = it != k.mMeasurements.end();
    meas.insert(meas.end(), {has ? 1 : 0, has ? (int32_t)it->second.Source : -1, has ? it->second.nLevel : -1,
                             (int32_t)mm.MMData(&points[i]).sMeasurementKFs.count(&k), (int32_t)mm.MMData(&points[i]).sNeverRetryKFs.count(&k)});
    meas_pos.push_back(has ? it->second.v2RootPos[0] : 0.0); meas_pos.push_back(has ? it->second.v2RootPos[1] : 0.0);
  }
  int per_level[LEVELS] = {0, 0, 0, 0};
  for (size_t i = npts; i < map.vpPoints.size(); i++) {
    MapPoint* p = map.vpPoints[i];
    per_level[p->nSourceLevel]++;
    if (p->pPatchSourceKF != &k || !k.mMeasurements.count(p)) { std::cerr << "new point not rooted in the new keyframe\n"; return 1; }
  }
  for (int l = 0; l < LEVELS; l++) newpts.push_back(per_level[l]);
  newpts.push_back((int32_t)(ClosestIndex(mm, k, kfs)));
  wr(dir, "ak_out_cand.i32", cand); wr(dir, "ak_out_meas.i32", meas); wr(dir, "ak_out_meas_pos.f64", meas_pos); wr(dir, "ak_out_new.i32", newpts);
  std::printf("add keyframe: %zu new points (levels %d %d %d %d)\n", map.vpPoints.size() - npts, per_level[0], per_level[1], per_level[2], per_level[3]);
  return 0;
}

// The two threads of the reference in one loop (System.cc:94 + MapMaker::run, MapMaker.cc:87-160): the Tracker mirror
// follows a sequence; one of its frames becomes a keyframe and goes through MapMaker::AddKeyFrame /
// AddKeyFrameFromTopOfQueue (new points by epipolar search), the tracker carries on with the enlarged map; at the end
// the new points are re-found in the older keyframes and the whole map is bundle-adjusted.
static int run_loop(const std::string& dir) {
  auto dims = rd<int32_t>(dir, "trk_dims.i32");  // W, H, n_kf, n_pts, n_frames, frame to hand over
  const int W = dims[0], H = dims[1], nkf = dims[2], npts = dims[3], nfr = dims[4], hand_over = dims[5];
  auto kfim = rd<uint8_t>(dir, "trk_kf.u8");
  auto kfpose = rd<double>(dir, "ak_kf_poses.f64");
  auto frames = rd<uint8_t>(dir, "trk_frames.u8");
  auto world = rd<double>(dir, "trk_world.f64");
  auto right = rd<double>(dir, "trk_right.f64");
  auto down = rd<double>(dir, "trk_down.f64");
  auto skf = rd<int32_t>(dir, "trk_srckf.i32");
  auto slv = rd<int32_t>(dir, "trk_srclevel.i32");
  auto ctr = rd<int32_t>(dir, "trk_center.i32");
  auto pose0 = rd<double>(dir, "trk_pose0.f64");
  ATANCamera cam("Camera", makeVector(1.0803, 1.43987, 0.519983, 0.548655, 0.244943), CVD::ImageRef(W, H));
  Map map;
  std::vector<KeyFrame> kfs(nkf);
  std::vector<MapPoint> points(npts);
  MapMaker mm(map, cam);
  for (int k = 0; k < nkf; k++) {
    CVD::BasicImage<CVD::byte> im(kfim.data() + (size_t)k * W * H, CVD::ImageRef(W, H));
    kfs[k].MakeKeyFrame_Lite(im);
    kfs[k].se3CfromW = se3_from_array(&kfpose[12 * k]);
    kfs[k].bFixed = k == 0;  // the first keyframe holds the gauge (MapMaker.cc:304)
    map.vpKeyFrames.push_back(&kfs[k]);
  }
  for (int i = 0; i < npts; i++) {
    MapPoint& p = points[i];
    p.v3WorldPos = makeVector(world[3 * i], world[3 * i + 1], world[3 * i + 2]);
    p.v3PixelRight_W = makeVector(right[3 * i], right[3 * i + 1], right[3 * i + 2]);
    p.v3PixelDown_W = makeVector(down[3 * i], down[3 * i + 1], down[3 * i + 2]);
    p.pPatchSourceKF = &kfs[skf[i]];
    p.nSourceLevel = slv[i];
    p.irCenter = CVD::ImageRef(ctr[2 * i], ctr[2 * i + 1]);
    map.vpPoints.push_back(&p);
    Measurement root;
    root.nLevel = slv[i]; root.bSubPix = true; root.Source = Measurement::SRC_ROOT;
    root.v2RootPos = Level::LevelZeroPos(p.irCenter, slv[i]);
    kfs[skf[i]].mMeasurements[&p] = root;
    mm.MMData(&p).sMeasurementKFs.insert(&kfs[skf[i]]);
  }
  map.bGood = true; map.nRevision++;
  Tracker trk(CVD::ImageRef(W, H), cam, map);
  trk.SetCurrentPose(se3_from_array(pose0.data()));
  std::vector<double> poses;
  std::vector<int32_t> found, info;
  CVD::Image<CVD::byte> frame(CVD::ImageRef(W, H));
  for (int f = 0; f < nfr; f++) {
    std::memcpy(frame.data(), frames.data() + (size_t)f * W * H, (size_t)W * H);
    trk.TrackFrame(frame, false);
    double a[12];
    se3_to_array(trk.GetCurrentPose(), a);
    poses.insert(poses.end(), a, a + 12);
    const ptam_track_result& r = trk.LastResult();
    found.push_back(r.meas_found[0] + r.meas_found[1] + r.meas_found[2] + r.meas_found[3]);
    if (f == hand_over) {
      const size_t before = map.vpPoints.size();
      mm.AddKeyFrame(trk.CurrentKeyFrame());
      mm.AddKeyFrameFromTopOfQueue();
      info.push_back((int32_t)(map.vpPoints.size() - before));
      info.push_back((int32_t)map.vpKeyFrames.back()->mMeasurements.size());
    }
  }
  // the map-maker thread's priority list until it has nothing left to do (MapMaker::run): with fewer than eight
  // keyframes the local adjustment is skipped, the new points are re-found in the older keyframes, the whole map is
  // adjusted (again and again until it converges), outlier measurements get their second chance, bad points leave
  const std::vector<MapPoint*> all_points = map.vpPoints;   // before any of them is moved to the trash
  std::vector<double> before_ba;
  for (MapPoint* p : all_points) for (int k = 0; k < 3; k++) before_ba.push_back(p->v3WorldPos[k]);
  const size_t new_queue = mm.mvpNewQueue.size();
  {
    // a bundle adjustment that moves the map must reach the tracker's device copy even when no point is trashed:
    // the revision has to change (the reference's tracker reads the live map)
    const unsigned rev0 = map.nRevision;
    const size_t n0 = map.vpPoints.size();
    mm.BundleAdjustAll();
    bool moved = false;
    for (size_t i = 0; i < all_points.size() && !moved; i++)
      for (int k = 0; k < 3; k++) moved = moved || all_points[i]->v3WorldPos[k] != before_ba[3 * i + k];
    if (moved && map.vpPoints.size() == n0 && map.nRevision == rev0) { std::fprintf(stderr, "BundleAdjust moved the map without a revision bump\n"); return 1; }
  }
  int passes = 0;
  do { mm.RunOnce(true); passes++; } while (passes < 12 && !(mm.mbBundleConverged_Full && mm.mbBundleConverged_Recent && mm.QueueSize() == 0));
  std::vector<double> after_ba;
  int nBad = 0, nDangling = 0;
  for (MapPoint* p : all_points) { for (int k = 0; k < 3; k++) after_ba.push_back(p->v3WorldPos[k]); nBad += p->bBad; }
  for (MapPoint* p : map.vpPoints) nDangling += p->bBad;                       // a bad point still in the map
  for (KeyFrame* kf : map.vpKeyFrames)
    for (auto& pm : kf->mMeasurements) nDangling += pm.first->bBad;            // a measurement of a point that left
  const int nRefound = (int)new_queue - (int)mm.mvpNewQueue.size();
  info.insert(info.end(), {nRefound, (int32_t)all_points.size(), (int32_t)map.vpKeyFrames.size(), mm.mbResetRequested ? 1 : 0,
                           mm.mbBundleConverged_Full ? 1 : 0, nBad, (int32_t)mm.mvFailureQueue.size(), passes, nDangling,
                           (int32_t)map.vpPoints.size(), (int32_t)map.vpPointsTrash.size()});
  // and the tracker carries on with the cleaned-up, adjusted map
  std::memcpy(frame.data(), frames.data() + (size_t)(nfr - 1) * W * H, (size_t)W * H);
  trk.TrackFrame(frame, false);
  {
    double a[12];
    se3_to_array(trk.GetCurrentPose(), a);
    poses.insert(poses.end(), a, a + 12);
    const ptam_track_result& r = trk.LastResult();
    found.push_back(r.meas_found[0] + r.meas_found[1] + r.meas_found[2] + r.meas_found[3]);
  }
  std::vector<double> kf0(12);
  se3_to_array(kfs[0].se3CfromW, kf0.data());
  wr(dir, "loop_out_poses.f64", poses); wr(dir, "loop_out_found.i32", found); wr(dir, "loop_out_info.i32", info);
  wr(dir, "loop_out_before_ba.f64", before_ba); wr(dir, "loop_out_after_ba.f64", after_ba); wr(dir, "loop_out_kf0.f64", kf0);
  std::printf("loop: %d frames, %d new points at the hand-over, %d re-found later, map of %zu points in %zu keyframes\n", nfr, info[0], nRefound,
              map.vpPoints.size(), map.vpKeyFrames.size());
  return 0;
}

// The tracker with a map maker attached: per frame the tracking quality and lost-frame counter after the last branch
// of Tracker::AssessTrackingQuality (Tracker.cc:1094-1099), and the map maker's queue after the keyframe hand-over
// heuristic (Tracker.cc:146-166).
static int run_heuristics(const std::string& dir) {
  auto dims = rd<int32_t>(dir, "trk_dims.i32");  // W, H, n_kf, n_pts, n_frames
  const int W = dims[0], H = dims[1], nkf = dims[2], npts = dims[3], nfr = dims[4];
  auto kfim = rd<uint8_t>(dir, "trk_kf.u8");
  auto kfpose = rd<double>(dir, "ak_kf_poses.f64");
  auto frames = rd<uint8_t>(dir, "trk_frames.u8");
  auto world = rd<double>(dir, "trk_world.f64");
  auto right = rd<double>(dir, "trk_right.f64");
  auto down = rd<double>(dir, "trk_down.f64");
  auto skf = rd<int32_t>(dir, "trk_srckf.i32");
  auto slv = rd<int32_t>(dir, "trk_srclevel.i32");
  auto ctr = rd<int32_t>(dir, "trk_center.i32");
  auto pose0 = rd<double>(dir, "trk_pose0.f64");
  auto hp = rd<double>(dir, "hq_params.f64");  // TrackingQualityGood, TrackingQualityLost, wiggle scale
  ATANCamera cam("Camera", makeVector(1.0803, 1.43987, 0.519983, 0.548655, 0.244943), CVD::ImageRef(W, H));
  Map map;
  std::vector<KeyFrame> kfs(nkf);
  std::vector<MapPoint> points(npts);
  for (int k = 0; k < nkf; k++) {
    CVD::BasicImage<CVD::byte> im(kfim.data() + (size_t)k * W * H, CVD::ImageRef(W, H));
    kfs[k].MakeKeyFrame_Lite(im);
    kfs[k].se3CfromW = se3_from_array(&kfpose[12 * k]);
    map.vpKeyFrames.push_back(&kfs[k]);
  }
  for (int i = 0; i < npts; i++) {
    MapPoint& p = points[i];
    p.v3WorldPos = makeVector(world[3 * i], world[3 * i + 1], world[3 * i + 2]);
    p.v3PixelRight_W = makeVector(right[3 * i], right[3 * i + 1], right[3 * i + 2]);
    p.v3PixelDown_W = makeVector(down[3 * i], down[3 * i + 1], down[3 * i + 2]);
    p.pPatchSourceKF = &kfs[skf[i]];
    p.nSourceLevel = slv[i];
    p.irCenter = CVD::ImageRef(ctr[2 * i], ctr[2 * i + 1]);
    map.vpPoints.push_back(&p);
  }
  map.bGood = true; map.nRevision++;
  MapMaker mm(map, cam);
  mm.mdWiggleScale = hp[2];
  ptam_tracker_params prm;
  ptam_tracker_default_params(&prm);
  prm.quality_good = hp[0]; prm.quality_lost = hp[1];
  Tracker trk(CVD::ImageRef(W, H), cam, map, 0, &prm);
  trk.SetMapMaker(&mm);
  trk.SetCurrentPose(se3_from_array(pose0.data()));
  std::vector<int32_t> out;
  CVD::Image<CVD::byte> frame(CVD::ImageRef(W, H));
  for (int f = 0; f < nfr; f++) {
    std::memcpy(frame.data(), frames.data() + (size_t)f * W * H, (size_t)W * H);
    trk.TrackFrame(frame, false);
    ptam_tracker_state st;
    ptam_tracker_get_state(trk.handle(), 0, &st);
    out.insert(out.end(), {st.tracking_quality, st.lost_frames, mm.QueueSize(), trk.LastResult().quality_needs_kf_distance});
  }
  wr(dir, "hq_out.i32", out);
  std::printf("heuristics: %d frames, %d keyframes queued\n", nfr, mm.QueueSize());
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) { std::cerr << "usage: mapmaker_check <dir> [ba|epi|refind|addkf|loop|heur]\n"; return 2; }
  const std::string dir = argv[1];
  try {
    if (argc > 2 && std::string(argv[2]) == "epi") return run_epipolar(dir);
    if (argc > 2 && std::string(argv[2]) == "refind") return run_refind(dir);
    if (argc > 2 && std::string(argv[2]) == "addkf") return run_add_keyframe(dir);
    if (argc > 2 && std::string(argv[2]) == "loop") return run_loop(dir);
    if (argc > 2 && std::string(argv[2]) == "heur") return run_heuristics(dir);
    auto cams = rd<double>(dir, "mm_cams.f64");
    auto fixed = rd<int32_t>(dir, "mm_fixed.i32");
    auto pts = rd<double>(dir, "mm_pts.f64");
    auto mcam = rd<int32_t>(dir, "mm_mcam.i32");
    auto mpt = rd<int32_t>(dir, "mm_mpt.i32");
    auto uv = rd<double>(dir, "mm_uv.f64");
    auto lvl = rd<int32_t>(dir, "mm_level.i32");
    auto src = rd<int32_t>(dir, "mm_src.i32");
    auto mode = rd<int32_t>(dir, "mm_mode.i32");  // [0]: 0 = BundleAdjustAll, 1 = BundleAdjustRecent; [1]: max iterations
    const size_t C = fixed.size(), P = pts.size() / 3;
    // contiguous storage: pointer order = index order, which is the order std::set / std::map walk them in
    std::vector<KeyFrame> kfs(C);
    std::vector<MapPoint> points(P);
    Map map;
    for (size_t c = 0; c < C; c++) {
      kfs[c].se3CfromW = se3_from_array(&cams[12 * c]);
      kfs[c].bFixed = fixed[c] != 0;
      map.vpKeyFrames.push_back(&kfs[c]);
    }
    for (size_t p = 0; p < P; p++) {
      points[p].v3WorldPos = makeVector(pts[3 * p], pts[3 * p + 1], pts[3 * p + 2]);
      map.vpPoints.push_back(&points[p]);
    }
    map.bGood = true;
    ATANCamera cam("Camera");
    ptam_bundle_params prm;
    ptam_bundle_default_params(&prm);
    prm.max_iterations = mode[1];
    MapMaker mm(map, cam, 0, &prm);
    for (size_t m = 0; m < mcam.size(); m++) {
      Measurement me;
      me.nLevel = lvl[m];
      me.bSubPix = false;
      me.v2RootPos = makeVector(uv[2 * m], uv[2 * m + 1]);
      me.Source = static_cast<decltype(me.Source)>(src[m]);
      kfs[mcam[m]].mMeasurements[&points[mpt[m]]] = me;
      mm.MMData(&points[mpt[m]]).sMeasurementKFs.insert(&kfs[mcam[m]]);
    }
    if (mode[0] == 0) mm.BundleAdjustAll(); else mm.BundleAdjustRecent();
    std::vector<double> opts, ocams(12 * C);
    std::vector<int32_t> bad, nmeas, queue, never;
    for (size_t p = 0; p < P; p++) {
      for (int k = 0; k < 3; k++) opts.push_back(points[p].v3WorldPos[k]);
      bad.push_back(points[p].bBad ? 1 : 0);
      for (KeyFrame* kf : mm.MMData(&points[p]).sNeverRetryKFs) { never.push_back((int32_t)(kf - kfs.data())); never.push_back((int32_t)p); }
    }
    for (size_t c = 0; c < C; c++) { se3_to_array(kfs[c].se3CfromW, &ocams[12 * c]); nmeas.push_back((int32_t)kfs[c].mMeasurements.size()); }
    for (auto& q : mm.mvFailureQueue) { queue.push_back((int32_t)(q.first - kfs.data())); queue.push_back((int32_t)(q.second - points.data())); }
    std::vector<int32_t> flags = {mm.mbBundleConverged_Full ? 1 : 0, mm.mbBundleConverged_Recent ? 1 : 0, mm.mbResetRequested ? 1 : 0,
                                  mm.mbBundleRunning ? 1 : 0};
    wr(dir, "mm_out_pts.f64", opts); wr(dir, "mm_out_cams.f64", ocams); wr(dir, "mm_out_bad.i32", bad); wr(dir, "mm_out_nmeas.i32", nmeas);
    wr(dir, "mm_out_queue.i32", queue); wr(dir, "mm_out_never.i32", never); wr(dir, "mm_out_flags.i32", flags);
    std::printf("mapmaker: mode %d, %zu keyframes, %zu points, failure queue %zu\n", mode[0], C, P, mm.mvFailureQueue.size());
  } catch (const std::exception& e) {
    std::cerr << "mapmaker_check failed: " << e.what() << "\n";
    return 1;
  }
  return 0;
}
