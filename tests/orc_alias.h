// Test-only: lets the host mirror's C++ sources (ptam_cg_b200/host/*.h, written against include/ptam_b200.h)
// link against the CPU oracle, which exports the same ABI under the prefix orc_.  Used by
// tests/test_host_mapmaker_cpu.py to run mapmaker_check.cc without a GPU; the product never sees this file.
#pragma once
#define ptam_bundle_default_params orc_bundle_default_params
#define ptam_bundle_create orc_bundle_create
#define ptam_bundle_destroy orc_bundle_destroy
#define ptam_bundle_last_error orc_bundle_last_error
#define ptam_bundle_add_camera orc_bundle_add_camera
#define ptam_bundle_add_point orc_bundle_add_point
#define ptam_bundle_add_meas orc_bundle_add_meas
#define ptam_bundle_compute orc_bundle_compute
#define ptam_bundle_recompute orc_bundle_recompute
#define ptam_bundle_update_camera orc_bundle_update_camera
#define ptam_bundle_update_point orc_bundle_update_point
#define ptam_bundle_converged orc_bundle_converged
#define ptam_bundle_get_point orc_bundle_get_point
#define ptam_bundle_get_camera orc_bundle_get_camera
#define ptam_bundle_get_outliers orc_bundle_get_outliers
#define ptam_tracker_create orc_tracker_create
#define ptam_tracker_destroy orc_tracker_destroy
#define ptam_tracker_last_error orc_tracker_last_error
#define ptam_tracker_level_size orc_tracker_level_size
#define ptam_tracker_get_level orc_tracker_get_level
#define ptam_tracker_make_keyframes orc_tracker_make_keyframes
#define ptam_tracker_keyframe_rest orc_tracker_keyframe_rest
#define ptam_tracker_get_level_rest orc_tracker_get_level_rest
#define ptam_tracker_default_params orc_tracker_default_params
#define ptam_tracker_add_keyframe orc_tracker_add_keyframe
#define ptam_tracker_set_map orc_tracker_set_map
#define ptam_tracker_set_keyframe_pose orc_tracker_set_keyframe_pose
#define ptam_tracker_set_state orc_tracker_set_state
#define ptam_tracker_get_state orc_tracker_get_state
#define ptam_tracker_get_points orc_tracker_get_points
#define ptam_tracker_track_frames orc_tracker_track_frames
#define ptam_tracker_epipolar_search orc_tracker_epipolar_search
#define ptam_tracker_refind_in_keyframes orc_tracker_refind_in_keyframes
#define ptam_tracker_get_templates orc_tracker_get_templates
#define ptam_patch_search_batch orc_patch_search_batch
#define ptam_patch_get_results orc_patch_get_results
#define ptam_pose_update orc_pose_update
#define ptam_global_last_error orc_test_global_last_error
#ifdef __cplusplus
extern "C"
#endif
inline const char* orc_test_global_last_error(void) { return "(oracle build: no global error string)"; }
