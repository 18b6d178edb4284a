// Exercises the C++ host mirror (Bundle / KeyFrame / Tracker with TooN/CVD-style types) end to end
// against libptam_b200.so.  Driven by tests/test_host_cpp_gpu.py: reads raw arrays from a directory,
// runs them through the classes exactly as MapMaker::BundleAdjust (MapMaker.cc:852-904) and
// System::UpdateFrame (System.cc:94) would, writes the results back as raw arrays.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>
#include "Bundle.h"
#include "Tracker.h"
#include "PatchFinder.h"

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

static void run_bundle(const std::string& dir) {
  auto cams = rd<double>(dir, "ba_cams.f64");
  auto fixed = rd<int32_t>(dir, "ba_fixed.i32");
  auto pts = rd<double>(dir, "ba_pts.f64");
  auto mcam = rd<int32_t>(dir, "ba_mcam.i32");
  auto mpt = rd<int32_t>(dir, "ba_mpt.i32");
  auto uv = rd<double>(dir, "ba_uv.f64");
  auto s2 = rd<double>(dir, "ba_s2.f64");
  ATANCamera cam("Camera");
  Bundle b(cam);
  for (size_t c = 0; c < fixed.size(); c++) b.AddCamera(se3_from_array(&cams[12 * c]), fixed[c] != 0);
  for (size_t p = 0; p < pts.size() / 3; p++) b.AddPoint(makeVector(pts[3 * p], pts[3 * p + 1], pts[3 * p + 2]));
  for (size_t m = 0; m < mcam.size(); m++) b.AddMeas(mcam[m], mpt[m], makeVector(uv[2 * m], uv[2 * m + 1]), s2[m]);
  bool abort = false;
  const int accepted = b.Compute(&abort);
  std::vector<double> opts, ocams(12 * fixed.size());
  for (size_t p = 0; p < pts.size() / 3; p++) { Vector<3> v = b.GetPoint((int)p); opts.insert(opts.end(), {v[0], v[1], v[2]}); }
  for (size_t c = 0; c < fixed.size(); c++) se3_to_array(b.GetCamera((int)c), &ocams[12 * c]);
  auto outl = b.GetOutlierMeasurements();
  std::vector<int32_t> oo, meta = {accepted, b.Converged() ? 1 : 0, (int32_t)outl.size()};
  for (auto& pc : outl) { oo.push_back(pc.first); oo.push_back(pc.second); }
  wr(dir, "ba_out_pts.f64", opts); wr(dir, "ba_out_cams.f64", ocams); wr(dir, "ba_out_meta.i32", meta); wr(dir, "ba_out_outliers.i32", oo);
  std::printf("bundle: accepted %d converged %d outliers %zu\n", accepted, (int)b.Converged(), outl.size());
}

static void run_tracker(const std::string& dir) {
  auto dims = rd<int32_t>(dir, "trk_dims.i32");  // W, H, n_kf, n_pts, n_frames
  const int W = dims[0], H = dims[1], nkf = dims[2], npts = dims[3], nfr = dims[4];
  auto kfim = rd<uint8_t>(dir, "trk_kf.u8");
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
  for (int k = 0; k < nkf; k++) {
    CVD::BasicImage<CVD::byte> im(kfim.data() + (size_t)k * W * H, CVD::ImageRef(W, H));
    kfs[k].MakeKeyFrame_Lite(im);  // KeyFrame.cc:18-54 through the device
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
  Tracker trk(CVD::ImageRef(W, H), cam, map);
  trk.SetCurrentPose(se3_from_array(pose0.data()));
  std::vector<double> poses;
  std::vector<int32_t> counts;
  CVD::Image<CVD::byte> frame(CVD::ImageRef(W, H));
  for (int f = 0; f < nfr; f++) {
    std::memcpy(frame.data(), frames.data() + (size_t)f * W * H, (size_t)W * H);
    trk.TrackFrame(frame, false);
    double a[12];
    se3_to_array(trk.GetCurrentPose(), a);
    poses.insert(poses.end(), a, a + 12);
    const ptam_track_result& r = trk.LastResult();
    for (int l = 0; l < LEVELS; l++) counts.push_back(r.meas_found[l]);
  }
  KeyFrame& kf = trk.CurrentKeyFrame();
  std::vector<int32_t> last = {(int32_t)kf.mMeasurements.size()};
  for (int l = 0; l < LEVELS; l++) last.push_back((int32_t)kf.aLevels[l].vCorners.size());
  // corners of source keyframe 0, level 0 (raster order) for a bit-exact check
  std::vector<int32_t> c0;
  for (auto& c : kfs[0].aLevels[0].vCorners) { c0.push_back(c.x); c0.push_back(c.y); }
  // ---- class PatchFinder, the reference's five steps method by method on the first points of the map, against
  // the last frame at the pose the test gives (tests/host_util.py compares with ptam_patch_search_batch)
  {
    auto pf_pose = rd<double>(dir, "pf_pose.f64");
    auto pf_cfg = rd<int32_t>(dir, "pf_cfg.i32");  // n points, range, sub-pixel iterations
    KeyFrame target;
    target.MakeKeyFrame_Lite(frame);
    const SE3<> se3 = se3_from_array(pf_pose.data());
    std::vector<int32_t> oi;   // per point: level, template bad, found coarse, coarse x, coarse y, template sum, sub-pixel converged
    std::vector<double> od;    // per point: warp inverse (4), sub-pixel position (2)
    std::vector<uint8_t> ot;   // per point: 64 template bytes
    PatchFinder finder(cam, CVD::ImageRef(W, H));
    for (int i = 0; i < pf_cfg[0] && i < npts; i++) {
      MapPoint& p = points[i];
      Matrix<2> derivs;
      const int lvl = finder.CalcSearchLevelAndWarpMatrix(p, se3, derivs);
      int bad = 0, found = 0, cx = 0, cy = 0, tsum = 0, conv = 0;
      double sx = 0, sy = 0;
      uint8_t tmpl[64] = {0};
      if (lvl >= 0) {
        finder.MakeTemplateCoarseCont(p);
        bad = finder.TemplateBad();
        if (!bad) {
          tsum = finder.GetTemplateSum();
          std::memcpy(tmpl, finder.GetTemplate().data(), 64);
          Vector<3> v3 = se3 * p.v3WorldPos;
          Vector<2> v2 = cam.Project(makeVector(v3[0] / v3[2], v3[1] / v3[2]));
          found = finder.FindPatchCoarse(CVD::ImageRef((int)v2[0], (int)v2[1]), target, (unsigned)pf_cfg[1]);
          if (found) {
            cx = finder.GetCoarsePos().x; cy = finder.GetCoarsePos().y;
            finder.MakeSubPixTemplate();
            conv = finder.IterateSubPixToConvergence(target, pf_cfg[2]);
            if (conv) { sx = finder.GetSubPixPos()[0]; sy = finder.GetSubPixPos()[1]; }
          }
        }
      }
      oi.insert(oi.end(), {lvl, bad, found, cx, cy, tsum, conv});
      const Matrix<2>& wi = finder.GetWarpInverse();
      od.insert(od.end(), {wi(0, 0), wi(0, 1), wi(1, 0), wi(1, 1), sx, sy});
      ot.insert(ot.end(), tmpl, tmpl + 64);
    }
    wr(dir, "pf_out.i32", oi); wr(dir, "pf_out.f64", od); wr(dir, "pf_out_tmpl.u8", ot);
    // and the batched form on all points
    std::vector<int> lv; std::vector<char> fd; std::vector<Vector<2> > ps;
    finder.SearchBatch(map.vpPoints, target, se3, (unsigned)pf_cfg[1], pf_cfg[2], &lv, &fd, &ps);
    std::vector<int32_t> bi; std::vector<double> bd;
    for (size_t i = 0; i < lv.size(); i++) { bi.push_back(lv[i]); bi.push_back(fd[i]); bd.push_back(ps[i][0]); bd.push_back(ps[i][1]); }
    wr(dir, "pf_batch.i32", bi); wr(dir, "pf_batch.f64", bd);
  }
  kfs[0].MakeKeyFrame_Rest();  // KeyFrame.cc:61-82 through the device
  std::vector<int32_t> rest = {(int32_t)kfs[0].aLevels[0].vMaxCorners.size(), (int32_t)kfs[0].aLevels[0].vCandidates.size()};
  for (auto& c : kfs[0].aLevels[0].vCandidates) { rest.push_back(c.irLevelPos.x); rest.push_back(c.irLevelPos.y); }
  wr(dir, "trk_out_kf0_rest.i32", rest);
  wr(dir, "trk_out_poses.f64", poses); wr(dir, "trk_out_found.i32", counts); wr(dir, "trk_out_last.i32", last); wr(dir, "trk_out_kf0_corners.i32", c0);
  std::printf("tracker: %d frames, last frame %zu measurements\n", nfr, kf.mMeasurements.size());
}

int main(int argc, char** argv) {
  if (argc < 2) { std::cerr << "usage: host_check <dir>\n"; return 2; }
  try {
    run_bundle(argv[1]);
    run_tracker(argv[1]);
  } catch (const std::exception& e) {
    std::cerr << "host_check failed: " << e.what() << "\n";
    return 1;
  }
  return 0;
}
