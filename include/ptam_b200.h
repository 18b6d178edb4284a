/*
 * ptam_b200.h — C-ABI of the B200-native PTAM hot paths (libptam_b200.so).
 *
 * The reference (cggos/ptam_cg) has no FFI/plugin layer: its boundary is ordinary C++ class methods
 * on TooN/CVD value types.  This header is the thin C-ABI those classes are re-hosted on
 * (the headers in ptam_cg_b200/host/ mirror Tracker / KeyFrame / PatchFinder / Bundle on top of it).
 * Plain pointers and sizes only; all host pointers are caller-owned and only touched during the
 * call; all device memory is owned by the handle.  Every function returns 0 (PTAM_OK) or a
 * negative error code unless stated; ptam_*_last_error() gives the text.  Algorithmic outcomes
 * (patch not found, bad template, BA not converged) are data, not errors.  A handle may be used by
 * one thread at a time; tracker and bundle handles own separate CUDA streams and may run
 * concurrently (the reference runs Tracker and MapMaker on two threads, MapMaker.h:37-38).
 *
 * There is NO CPU fallback: creating a handle without a usable sm_100 device fails.
 *
 * SE3 layout everywhere: 12 doubles = rotation matrix row-major (9) then translation (3),
 * "camera from world" as in KeyFrame::se3CfromW (KeyFrame.h:136).
 */
#ifndef PTAM_B200_H
#define PTAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTAM_LEVELS 4 /* KeyFrame.h:34 */

enum {
  PTAM_OK = 0,
  PTAM_ERR_INVALID = -1,  /* bad argument */
  PTAM_ERR_CUDA = -2,     /* CUDA runtime error (sticky per handle) */
  PTAM_ERR_NO_DEVICE = -3,
  PTAM_ERR_CAPACITY = -4,
  PTAM_ERR_NCCL = -5
};

/* per-point result flags (TrackerData booleans, Tracker.h:54-61) */
enum {
  PTAM_PT_IN_IMAGE = 1,   /* bInImage after the last Project */
  PTAM_PT_IN_PVS = 2,     /* entered the potentially-visible set this frame */
  PTAM_PT_SEARCHED = 4,   /* bSearched */
  PTAM_PT_FOUND = 8,      /* bFound */
  PTAM_PT_SUBPIX = 16,    /* bDidSubPix */
  PTAM_PT_TEMPLATE_BAD = 32
};

/* ------------------------------------------------------------------------------------------
 * Path T — per-frame tracker.
 * Replaces: Tracker::TrackFrame (Tracker.cc:86-188, the MakeKeyFrame_Lite / PredictPose /
 * TrackMap / UpdateMotionModel / AssessTrackingQuality part), KeyFrame::MakeKeyFrame_Lite
 * (KeyFrame.cc:18-54), PatchFinder steps 1-5 as driven by Tracker::SearchForPoints
 * (Tracker.cc:867-912, PatchFinder.cc:52-318), Tracker::CalcPoseUpdate (Tracker.cc:928-1005).
 * ------------------------------------------------------------------------------------------ */

/* GVars3 keys the tracker reads, with the reference defaults (Tracker.cc:491-496,596,931,1088-1089;
 * Tracker.cc:95-96). */
typedef struct ptam_tracker_params {
  int32_t coarse_min;            /* Tracker.CoarseMin            = 20  */
  int32_t coarse_max;            /* Tracker.CoarseMax            = 60  */
  int32_t coarse_range;          /* Tracker.CoarseRange          = 30  */
  int32_t coarse_subpix_its;     /* Tracker.CoarseSubPixIts      = 8   */
  int32_t disable_coarse;        /* Tracker.DisableCoarse        = 0   */
  int32_t max_patches_per_frame; /* Tracker.MaxPatchesPerFrame   = 1000 */
  int32_t mestimator;            /* Tracker.MEstimator: 0 Tukey (default), 1 Cauchy, 2 Huber */
  int32_t use_constant_velocity; /* Tracker.UseConstantVelocity  = 1   */
  double coarse_min_velocity;    /* Tracker.CoarseMinVelocity    = 0.006 */
  double quality_good;           /* Tracker.TrackingQualityGood  = 0.3  */
  double quality_lost;           /* Tracker.TrackingQualityLost  = 0.13 */
  int32_t use_rotation_estimator; /* Tracker.UseRotationEstimator = 1 (Tracker.cc:96): SmallBlurryImage ESM rotation
                                     replaces the rotational velocity in PredictPoseWithMotionModel */
  int32_t reserved0;
  double rotation_estimator_blur; /* Tracker.RotationEstimatorBlur = 0.75 (Tracker.cc:95) */
} ptam_tracker_params;

/* Tracker member state carried from frame to frame (Tracker.h:176-215). */
typedef struct ptam_tracker_state {
  double se3_cam_from_world[12];        /* mse3CamFromWorld */
  double velocity[6];                   /* mv6CameraVelocity */
  double msd_scaled_velocity_magnitude; /* mdMSDScaledVelocityMagnitude */
  double scene_depth_mean;              /* mCurrentKF.dSceneDepthMean */
  double scene_depth_sigma;             /* mCurrentKF.dSceneDepthSigma */
  int32_t just_recovered_so_use_coarse; /* mbJustRecoveredSoUseCoarse */
  int32_t tracking_quality;             /* 0 BAD, 1 DODGY, 2 GOOD (Tracker.h:203) */
  int32_t lost_frames;                  /* mnLostFrames */
  int32_t frame;                        /* mnFrame */
} ptam_tracker_state;

/* What one TrackFrame leaves behind, per stream. */
typedef struct ptam_track_result {
  double se3_cam_from_world[12]; /* pose after TrackMap (mCurrentKF.se3CfromW, Tracker.cc:662) */
  double scene_depth_mean;       /* Tracker.cc:680-697 (unchanged when <= 20 points found) */
  double scene_depth_sigma;
  int32_t meas_attempted[PTAM_LEVELS]; /* manMeasAttempted */
  int32_t meas_found[PTAM_LEVELS];     /* manMeasFound */
  int32_t n_corners[PTAM_LEVELS];      /* FAST corners per pyramid level of the current frame */
  int32_t did_coarse;                  /* mbDidCoarse */
  int32_t n_coarse;                    /* size of the coarse-stage search set */
  int32_t n_level3;                    /* level-3 points searched in the fine stage */
  int32_t n_fine;                      /* level 2..0 points searched in the fine stage */
  int32_t tracking_quality;            /* after AssessTrackingQuality */
  int32_t quality_needs_kf_distance;   /* 1: reference would consult MapMaker::ClosestKeyFrame
                                          (Tracker.cc:1095-1099); left to the caller */
  int32_t n_pvs[PTAM_LEVELS];          /* PVS size per level before selection */
  int32_t n_candidates;                /* ZMSSD windows evaluated (FindPatchCoarse candidates, PatchFinder.cc:193-200) */
  int32_t recovery;                    /* 0: tracked normally; 1: the frame arrived with mnLostFrames >= 3, the relocaliser
                                          succeeded and TrackMap + AssessTrackingQuality ran from its pose (Tracker.cc:170-178);
                                          2: the relocaliser's score was too high, nothing else was done this frame */
  int32_t reloc_keyframe;              /* Relocaliser::mnBest (-1 when recovery == 0) */
  int32_t reserved1;
  double reloc_score;                  /* final ESM score of CalcSBIRotation against that keyframe (Relocaliser.cc:34-37) */
} ptam_track_result;

typedef struct ptam_tracker ptam_tracker;

void ptam_tracker_default_params(ptam_tracker_params* p);

/* n_streams independent trackers (own pose, map, per-point template cache) that share one camera
 * model, image size and keyframe store and are processed as one batch per call.  n_streams = 1 is
 * the reference's single Tracker.  Returns NULL on failure (see ptam_global_last_error). */
ptam_tracker* ptam_tracker_create(int device, const double cam_params[5], int width, int height,
                                  int n_streams, const ptam_tracker_params* params);
void ptam_tracker_destroy(ptam_tracker* t);
const char* ptam_tracker_last_error(const ptam_tracker* t);
const char* ptam_global_last_error(void);

/* Store a source keyframe (pyramid only) so map points can take their templates from it
 * (MapPoint::pPatchSourceKF, Map.h:67).  Returns the keyframe id (>= 0) or an error. */
int ptam_tracker_add_keyframe(ptam_tracker* t, const uint8_t* image, int stride);

/* Replace stream `stream`'s map.  SoA of MapPoint fields the tracker reads (Map.h:64-84):
 * world_pos / pixel_right_w / pixel_down_w: n*3 doubles; src_kf: id from add_keyframe;
 * src_level: nSourceLevel; ir_center: n*2 ints (x,y) in source-level pixels.  Clears the per-point
 * template cache and the M-estimator inlier/outlier counters. */
int ptam_tracker_set_map(ptam_tracker* t, int stream, int n_points, const double* world_pos,
                         const double* pixel_right_w, const double* pixel_down_w,
                         const int32_t* src_kf, const int32_t* src_level, const int32_t* ir_center);

/* Relocaliser (Relocaliser.cc:12-38, Tracker.cc:170-178,196-207).  Giving every stored
 * keyframe its pose (KeyFrame::se3CfromW) switches the recovery branch of Tracker::TrackFrame on: a frame that
 * arrives with mnLostFrames >= 3 is not tracked from the motion model; instead the SmallBlurryImage (blur 2.5)
 * of the frame is compared (SSD) with that of every stored keyframe, the ESM rotation against the best one
 * gives pose = rotation * keyframe pose, and if the final ESM score is below Reloc2.MaxScore (9e6) the pose,
 * a zero velocity and mbJustRecoveredSoUseCoarse are installed and TrackMap + AssessTrackingQuality run.
 * Without keyframe poses a lost stream keeps tracking from its motion model (the recovery branch is off). */
int ptam_tracker_set_keyframe_pose(ptam_tracker* t, int keyframe, const double se3_cam_from_world[12]);
int ptam_tracker_set_state(ptam_tracker* t, int stream, const ptam_tracker_state* s);
int ptam_tracker_get_state(ptam_tracker* t, int stream, ptam_tracker_state* s);

/* KeyFrame::MakeKeyFrame_Lite only, for every stream: images[s] is stream s's W x H u8 frame. */
int ptam_tracker_make_keyframes(ptam_tracker* t, const uint8_t* const* images, int stride);

/* One TrackFrame per stream: MakeKeyFrame_Lite + SmallBlurryImage update + PredictPoseWithMotionModel
 * (with the SmallBlurryImage rotation estimator when use_rotation_estimator) + TrackMap +
 * UpdateMotionModel + AssessTrackingQuality.  std::random_shuffle (Tracker.cc:483,601) is the identity permutation.
 * images: n_streams host pointers; results: n_streams structs (may be NULL). */
int ptam_tracker_track_frames(ptam_tracker* t, const uint8_t* const* images, int stride,
                              ptam_track_result* results);

/* Same, with the frames already resident in device memory: frame s starts at
 * d_images + s * frame_pitch_bytes, rows `stride` bytes apart.  results may be NULL (no D2H, no
 * host synchronisation: the call only enqueues work on the handle's stream). */
int ptam_tracker_track_frames_device(ptam_tracker* t, const uint8_t* d_images,
                                     size_t frame_pitch_bytes, int stride,
                                     ptam_track_result* results);
/* Pipelined form of ptam_tracker_track_frames for sustained streaming: submit enqueues the H2D copy
 * of this batch on a copy stream (into one of two landing buffers) and the tracking kernels behind
 * it, then returns; collect blocks until the OLDEST submitted batch is done and hands out its
 * results.  At most two batches may be in flight, so the copy of batch i+1 overlaps the kernels of
 * batch i.  Host frame buffers must stay valid until the matching collect (page-locked buffers are
 * read by DMA; pageable ones are staged inside submit). */
int ptam_tracker_submit_frames(ptam_tracker* t, const uint8_t* const* images, int stride);
/* The same for frames that already live in device memory (layout as ptam_tracker_track_frames_device).  The
 * frames must be complete when `ready_event` (a cudaEvent_t, or NULL = complete now) has fired, and stay
 * untouched until the matching collect.  Unlike ptam_tracker_track_frames_device the batch is NOT ordered
 * behind the work already queued on the handle's stream: its pyramid / FAST / corner-list kernels run on a
 * second stream beside the last Gauss-Newton iterations of the previous batch. */
int ptam_tracker_submit_frames_device(ptam_tracker* t, const uint8_t* d_images, size_t frame_pitch_bytes, int stride,
                                      void* ready_event);
int ptam_tracker_collect(ptam_tracker* t, ptam_track_result* results);
int ptam_tracker_synchronize(ptam_tracker* t);
/* cudaStream_t of the handle, as an opaque pointer (for CUDA-event timing by the caller). */
void* ptam_tracker_cuda_stream(ptam_tracker* t);
/* Number of kernel launches issued by the handle so far. */
int64_t ptam_tracker_launch_count(const ptam_tracker* t);
/* Per-kernel device timing: when on, every launch is bracketed by CUDA events on the handle's
 * stream and each call synchronises to accumulate them.  Kernel ids: 0 k_pyramid, 1 k_fast,
 * 2 k_compact, 3 k_sbi + k_pvs_select, 4 k_search(coarse), 5 k_pose(coarse), 6 k_search(fine), 7 k_pose(fine).
 * Turning it on or off resets the accumulators. */
int ptam_tracker_set_profiling(ptam_tracker* t, int on);
int ptam_tracker_get_kernel_times(ptam_tracker* t, double ms_total[8], int64_t launches[8]);

/* Read back the current frame's keyframe level of one stream (Level, KeyFrame.h:55-125):
 * pixels (w*h, may be NULL), corners as interleaved (x,y) int32 in raster order (cap pairs, may be
 * NULL), row_lut (h ints, may be NULL).  Returns the number of corners on that level. */
int ptam_tracker_get_level(ptam_tracker* t, int stream, int level, uint8_t* pixels,
                           int32_t* corners_xy, int corners_cap, int32_t* row_lut);
int ptam_tracker_level_size(const ptam_tracker* t, int level, int* w, int* h);

/* Per-point results of the last frame of one stream (any pointer may be NULL):
 * flags: PTAM_PT_* bits; level: nSearchLevel (-1 if not in the PVS); v2_found / v2_image: n*2. */
int ptam_tracker_get_points(ptam_tracker* t, int stream, int32_t* flags, int32_t* level,
                            double* v2_found, double* v2_image, int32_t* outlier_count,
                            int32_t* inlier_count);
/* MapMaker::ReFindInSingleKeyFrame / ReFind_Common (MapMaker.cc:943-1040), one keyframe per stream:
 * images[s] becomes stream s's current frame (MakeKeyFrame_Lite), se3 + 12 s is that keyframe's
 * se3CfromW, and every point of stream s's map is looked for in it: projection, warp matrix and
 * level, template (always re-made), FindPatchCoarse with search radius 4, sub-pixel refinement on
 * levels > 0 (its convergence is not checked, as in the reference).  Results per point through
 * ptam_tracker_get_points: PTAM_PT_FOUND / PTAM_PT_SUBPIX, level, v2_found (= Measurement::v2RootPos,
 * Source SRC_REFIND); a projected point without PTAM_PT_FOUND is one the reference adds to
 * sNeverRetryKFs.  The call overwrites the handle's per-point search state and template cache and
 * leaves the tracking state (pose, velocity) alone: give the map maker a handle of its own, as the
 * reference gives it its own PatchFinder (MapMaker.cc:977). */
int ptam_tracker_refind_in_keyframes(ptam_tracker* t, const uint8_t* const* images, int stride,
                                     const double* se3_cam_from_world /* n_streams * 12 */);
/* MapMaker::AddPointEpipolar (MapMaker.cc:529-688) for a batch of candidates, up to the sub-pixel position in
 * the target keyframe.  Source = stored keyframe src_kf (ptam_tracker_add_keyframe) with pose src_se3 and
 * scene depth mean / sigma (KeyFrame::dSceneDepthMean/Sigma); target = stream's current frame (its FAST
 * corners of `level` are the search set) with pose target_se3; cand_xy = Candidate::irLevelPos pairs in the
 * source level (e.g. from ptam_tracker_get_level_rest); wiggle_scale = MapMaker::mdWiggleScale.
 * Per candidate: the epipolar segment of its view ray over the depth range [max(wiggle, mean - sigma),
 * min(40 wiggle, mean + sigma)] in the target's z = 1 plane, all target corners within OnePixelDist (4 +
 * LevelScale) of it, un-warped 8x8 template (MakeTemplateCoarseNoWarp), first minimum ZMSSD <= mnMaxSSD,
 * MakeSubPixTemplate + IterateSubPixToConvergence(10).  found[i] = 1 when all of that succeeded (the
 * reference's `return true` apart from the triangulation), best_corner[i] = index of the winning corner in
 * the target level's list (-1: none), sub_pos = Finder.GetSubPixPos() (level-zero pixels).  The
 * triangulation and MapPoint construction that follow (a 4x4 SVD per accepted point, MapMaker.cc:648-687)
 * stay with the caller. */
int ptam_tracker_epipolar_search(ptam_tracker* t, int stream, int level, int src_kf, const double src_se3[12],
                                 double src_depth_mean, double src_depth_sigma, const double target_se3[12],
                                 double wiggle_scale, int n_cand, const int32_t* cand_xy, int32_t* found,
                                 int32_t* best_corner, double* sub_pos);
/* KeyFrame::MakeKeyFrame_Rest (KeyFrame.cc:61-82) for the current frame of one stream — what the map
 * maker needs when the frame becomes a keyframe: fast_nonmax(im, vCorners, 10, vMaxCorners) on every
 * level, then the Shi-Tomasi candidates (ImageProcess.cc:20-47; in_image_with_border 10, score >
 * min_shi_tomasi_score = MapMaker.CandidateMinShiTomasiScore, 70 in code, 400 in settings.cfg:27).
 * (The relocaliser's SmallBlurryImage of that function is the one ptam_tracker_get_sbi returns.)
 * get_level_rest: vMaxCorners as (x,y) pairs in raster order (returns their number), vCandidates as
 * irLevelPos pairs + dSTScore (their number in *n_cand). */
int ptam_tracker_keyframe_rest(ptam_tracker* t, int stream, double min_shi_tomasi_score);
int ptam_tracker_get_level_rest(ptam_tracker* t, int stream, int level, int32_t* max_corners_xy, int max_cap,
                                int32_t* cand_xy, double* cand_score, int cand_cap, int* n_cand);
/* SmallBlurryImage of the last frame of one stream (ImageProcess.cc:279-304): mimTemplate (w*h floats,
 * at most cap), the rotation estimate CalcSBIRotation gave against the previous frame (so3 log, 3
 * doubles) and its final ESM score.  Returns w*h of the small image ((W/8)/2 x (H/8)/2). */
int ptam_tracker_get_sbi(ptam_tracker* t, int stream, float* tmpl, int cap, double rot3[3], double* score);
/* Cached coarse templates of one stream: tmpl n*64 bytes, sums n*2 ints (sum, sum of squares). */
int ptam_tracker_get_templates(ptam_tracker* t, int stream, uint8_t* tmpl, int32_t* sums);
/* vIterationSet of the last frame in order (coarse set, level-3 set, fine set); returns its size. */
int ptam_tracker_get_iteration_set(ptam_tracker* t, int stream, int32_t* idx, int cap);

/* ---- PatchFinder / CalcPoseUpdate unit entry points ---------------------------------------------
 * The per-frame path runs PatchFinder and CalcPoseUpdate inside ptam_tracker_track_frames; these two calls expose
 * the same device code one step at a time, for unit parity and for callers that hold their own pose (the host
 * mirror ptam_cg_b200/host/PatchFinder.h is built on them).
 *
 * ptam_patch_search_batch = class PatchFinder, steps 1-5 (reference include/PatchFinder.h:54-98), for EVERY map
 * point of every stream against the stream's current frame (pyramid + FAST corners of the last
 * ptam_tracker_make_keyframes / _track_frames call), at the caller's poses se3[S][12], driven as
 * Tracker::SearchForPoints drives them (src/Tracker.cc:867-912):
 *   TrackerData::Project + GetProjectionDerivs (Tracker.h:70-94), CalcSearchLevelAndWarpMatrix
 *   (PatchFinder.cc:52-84; level -1 rejects), MakeTemplateCoarseCont (:98-127, per-point template cache),
 *   FindPatchCoarse(ir(v2Image), kf, range) (:160-211), and for subpix_its > 0 MakeSubPixTemplate +
 *   IterateSubPixToConvergence(kf, subpix_its) (:219-318; a point that does not converge is not found).
 * Tracker state (pose, velocity, counters of lost frames) is not touched.
 * ptam_patch_get_results: level[n] (mnSearchLevel or -1), warp_inverse[n][4] (mm2WarpInverse row-major, zero for a
 * point that does not project into the frame), template_bad[n] (TemplateBad() of the points that do), found[n], pos[n][2] (GetSubPixPos()
 * if the sub-pixel step ran, else GetCoarsePosAsVector(); level-zero pixels), subpix_converged[n]; templates and
 * their sums through ptam_tracker_get_templates.  Any output may be NULL; returns n.
 *
 * ptam_pose_update = Tracker::CalcPoseUpdate(vTD, dOverrideSigma, bMarkOutliers) (src/Tracker.cc:928-1005) once per
 * stream over the points FOUND by the last ptam_patch_search_batch: CalcJacobian (Tracker.h:125-136) at that call's
 * poses, sigma^2 from the M-estimator (Tools.h:128-254) unless override_sigma_squared > 0, WLS<6> with the 100 I
 * prior.  mu6[S][6] = v6Update (the caller applies SE3::exp(mu) * pose, Tracker.cc:567,641), n_found[S]. */
int ptam_patch_search_batch(ptam_tracker* t, const double* se3_cam_from_world, unsigned range, int subpix_its);
int ptam_patch_get_results(ptam_tracker* t, int stream, int32_t* level, double* warp_inverse, int32_t* template_bad,
                           int32_t* found, double* pos, int32_t* subpix_converged);
int ptam_pose_update(ptam_tracker* t, double override_sigma_squared, int mark_outliers, double* mu6, int32_t* n_found);

/* ------------------------------------------------------------------------------------------
 * Path B — bundle adjuster.  Replaces class Bundle (Bundle.h:105-156), whose only caller is
 * MapMaker::BundleAdjust (MapMaker.cc:838-933).
 * ------------------------------------------------------------------------------------------ */
typedef struct ptam_bundle_params {
  int32_t max_iterations;              /* Bundle.MaxIterations = 20 (Bundle.cc:40) */
  int32_t mestimator;                  /* Bundle.MEstimator: 0 Tukey (default), 1 Cauchy, 2 Huber */
  double update_squared_convergence;   /* Bundle.UpdateSquaredConvergenceLimit = 1e-6 (Bundle.cc:41) */
  double min_tukey_sigma;              /* Bundle.MinTukeySigma = 0.4 (Bundle.cc:234) */
} ptam_bundle_params;

typedef struct ptam_bundle_stats {
  int32_t accepted;        /* mnAccepted */
  int32_t lambda_trials;   /* mnCounter */
  int32_t lm_steps;        /* calls of Do_LM_Step */
  int32_t converged;       /* mbConverged */
  int32_t hit_max_iterations;
  int32_t n_outliers;      /* measurements erased so far */
  double sigma_squared;    /* mdSigmaSquared of the last step */
  double lambda;           /* mdLambda */
  double last_error;       /* dCurrentError of the last step */
  double last_new_error;   /* dNewError of the last lambda trial */
} ptam_bundle_stats;

typedef struct ptam_bundle ptam_bundle;

void ptam_bundle_default_params(ptam_bundle_params* p);
/* Bundle::Bundle(const ATANCamera&) — Bundle.cc:35-43.  width/height = ATANCamera image size. */
ptam_bundle* ptam_bundle_create(int device, const double cam_params[5], int width, int height,
                                const ptam_bundle_params* params);
void ptam_bundle_destroy(ptam_bundle* b);
const char* ptam_bundle_last_error(const ptam_bundle* b);

/* Bundle::AddCamera (Bundle.cc:46-63) → camera id. */
int ptam_bundle_add_camera(ptam_bundle* b, const double se3_cam_from_world[12], int fixed);
/* Bundle::AddPoint (Bundle.cc:66-78) → point id; NaN positions are zeroed as in the reference. */
int ptam_bundle_add_point(ptam_bundle* b, const double xyz[3]);
/* Bundle::AddMeas (Bundle.cc:81-93). */
int ptam_bundle_add_meas(ptam_bundle* b, int cam, int point, const double uv[2], double sigma_squared);
/* Bulk ingest in the same insertion order as repeated Add* calls. */
int ptam_bundle_add_cameras(ptam_bundle* b, int n, const double* se3, const int32_t* fixed);
int ptam_bundle_add_points(ptam_bundle* b, int n, const double* xyz);
int ptam_bundle_add_measurements(ptam_bundle* b, int n, const int32_t* cam, const int32_t* point,
                                 const double* uv, const double* sigma_squared);

/* Multi-GPU (no reference counterpart): one process per GPU, every rank feeds the SAME
 * graph through the Add* calls, and the handle of rank r keeps the points of its contiguous range
 * (ptam_bundle_shard_plan, balanced by measurement count) with all their measurements; cameras are
 * replicated.  Inside ptam_bundle_compute / _lm_step the partial reduced camera system (S, vE) of
 * every rank is summed (the cross-camera J^T J reduction: the packed lower triangle through ncclAllReduce,
 * or, for two ranks of one node, in place through an NVLink peer window the ranks map with CUDA IPC) and
 * solved identically everywhere; the sigma-squared order statistic is found exactly by a radix select on the
 * all-gathered squared errors; the LM step's error sums and abort votes ride with the first exchange, the
 * trial's in one small all-reduce.  After compute every rank holds all points, cameras and the merged
 * outlier list.  The peer window belongs to the communicator: ONE sharded handle per communicator at a time
 * (as MapMaker builds one Bundle at a time), and ptam_nccl_comm_destroy releases it.
 * In a sharded run compute / lm_step / get_point(s) / get_outliers / get_stats are COLLECTIVE.
 *   ptam_nccl_unique_id   rank 0 creates the id and ships it to the other ranks by any host channel;
 *   ptam_bundle_init_shard  creates the handle's own communicator from it (ncclCommInitRank);
 *   ptam_bundle_set_shard   alternatively adopts an ncclComm_t the caller already owns. */
#define PTAM_NCCL_UNIQUE_ID_BYTES 128
int ptam_nccl_unique_id(unsigned char id[PTAM_NCCL_UNIQUE_ID_BYTES]);
int ptam_bundle_init_shard(ptam_bundle* b, int rank, int world, const unsigned char id[PTAM_NCCL_UNIQUE_ID_BYTES]);
int ptam_bundle_set_shard(ptam_bundle* b, int rank, int world, void* nccl_comm);
/* A communicator that outlives individual handles (MapMaker builds a new Bundle per adjustment,
 * MapMaker.cc:840; a unique id can initialise only one communicator): create it once per process,
 * hand it to every handle with ptam_bundle_set_shard, destroy it at shutdown.  NULL on failure. */
void* ptam_nccl_comm_create(int device, int rank, int world, const unsigned char id[PTAM_NCCL_UNIQUE_ID_BYTES]);
void ptam_nccl_comm_destroy(void* comm);
/* Host-only: point_begin[world + 1], shard r owns points [point_begin[r], point_begin[r + 1]). */
int ptam_bundle_shard_plan(int n_points, int n_meas, const int32_t* meas_point, int world, int32_t* point_begin);

/* Bundle::Compute (Bundle.cc:116-158): returns the number of accepted LM steps (>= 0) or a negative
 * error.  abort_flag is polled between device phases like *pbAbortSignal (Bundle.cc:134,338);
 * may be NULL. */
int ptam_bundle_compute(ptam_bundle* b, const volatile unsigned char* abort_flag);
/* One Do_LM_Step only (Bundle.cc:209-551), for step-wise parity tests.  Call
 * ptam_bundle_begin() once before the first step (does what Compute does before its loop). */
int ptam_bundle_begin(ptam_bundle* b);
int ptam_bundle_lm_step(ptam_bundle* b, const volatile unsigned char* abort_flag);

/* Persistent graph.  The reference rebuilds a Bundle from the map for every
 * MapMaker::BundleAdjust call (MapMaker.cc:852-882), although consecutive calls on the same keyframe set
 * (BundleAdjustAll until converged, MapMaker.cc:67-77) feed back exactly what the previous Compute left
 * behind: adjusted poses and points, measurement list minus the erased outliers.  That state stays on the
 * device: ptam_bundle_recompute runs Bundle::Compute again on it (LM control reset as in Bundle.cc:121-126,
 * outlier list restarted) without the host-side graph rebuild and upload; same return value as compute.
 * ptam_bundle_update_camera / _point overwrite one pose / position of the resident graph in between. */
int ptam_bundle_recompute(ptam_bundle* b, const volatile unsigned char* abort_flag);
int ptam_bundle_update_camera(ptam_bundle* b, int n, const double se3_cam_from_world[12]);
int ptam_bundle_update_point(ptam_bundle* b, int n, const double xyz[3]);
int ptam_bundle_converged(const ptam_bundle* b);
int ptam_bundle_get_point(ptam_bundle* b, int n, double xyz[3]);
int ptam_bundle_get_camera(ptam_bundle* b, int n, double se3[12]);
int ptam_bundle_get_points(ptam_bundle* b, double* xyz /* n_points*3 */);
int ptam_bundle_get_cameras(ptam_bundle* b, double* se3 /* n_cameras*12 */);
/* Bundle::GetOutlierMeasurements: (point, camera) pairs in erase order; returns the total count
 * (may exceed cap; only cap pairs are written). */
int ptam_bundle_get_outliers(ptam_bundle* b, int32_t* point_cam_pairs, int cap);
int ptam_bundle_get_stats(ptam_bundle* b, ptam_bundle_stats* s);
/* Reduced camera system of the last lambda trial, for parity tests: S n*n row-major (both
 * triangles), vE n; returns n = 6 * non-fixed cameras. */
int ptam_bundle_get_reduced_system(ptam_bundle* b, double* S, double* vE, int cap_n);
int ptam_bundle_synchronize(ptam_bundle* b);
void* ptam_bundle_cuda_stream(ptam_bundle* b);
int64_t ptam_bundle_launch_count(const ptam_bundle* b);
/* Task table of the persistent dense solve (reference: Cholesky<>(mS).backsub(vE), Bundle.cc:457-458) for an n x n
 * reduced system: *k_start = first panel of the persistent kernel (the panels before it run one launch each),
 * task_off[k], k = *k_start .. n_panels: first ticket of round k (see csrc/ldlt_dag.cuh).  Host arithmetic only, no
 * device is touched; for tests of the task order.  Returns the number of 64-wide panels, or a negative PTAM_ERR_*
 * when `cap` is smaller than panels + 1. */
int ptam_bundle_solve_schedule(int n, int tail_tiles, int* k_start, int32_t* task_off, int cap);
/* Per-phase device timing (CUDA events on the handle's stream).  Phase ids: 0 project (Bundle.cc:219-225),
 * 1 sigma-squared select (:230-237), 2 Jacobian/accumulate (:251-332), 3 V*^-1 + S/vE init (:341-392),
 * 4 Schur build (:396-446), 5 cross-shard all-reduce of S/vE, 6 dense LDL^T solve (:457-458),
 * 7 updates + FindNewError (:461-506), 8 reserved, 9 commit + outlier erase (:512-547) — device time per call;
 * 10 set-up of Compute (host CSR build, upload, pair list: GenerateMeasLUTs / OffDiagScripts, :558-599) and
 * 11 the rest of Compute's wall clock (host LM control, blocking scalar read-backs, launch gaps) — host wall
 * clock, one entry per Compute.  Turning it on or off resets the accumulators. */
#define PTAM_BA_PHASES 12
int ptam_bundle_set_profiling(ptam_bundle* b, int on);
int ptam_bundle_get_phase_times(ptam_bundle* b, double ms_total[PTAM_BA_PHASES], int64_t count[PTAM_BA_PHASES]);

#ifdef __cplusplus
}
#endif
#endif /* PTAM_B200_H */
