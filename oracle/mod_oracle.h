/*
 * mod_oracle.h -- C interface of the CPU ORACLE for the scene-flow / clusterer hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under moving_object_detector_b200/ (the product) may include,
 * link or call this.  Allowed users: tests/, __graft_entry__.smoke(), and bench.py's cpu_baseline /
 * --impl reference legs (as the checker / the timed CPU baseline, never as the shipped path).
 *
 * PARITY STATUS: *pinned to the reference's own code, with its third-party dependencies shimmed*.  The reference
 * ships no tests, golden vectors or fixtures and its dependencies (ROS Melodic, PCL, Eigen, OpenCV) are not in this
 * image, but its hot-path translation units compile UNMODIFIED, from where they lie under /root/reference, against
 * the stand-in headers of oracle/ref_shim/ (oracle/Makefile target ref_full -> oracle/_ref/libmod_ref.so, driven by
 * oracle/ref_full_wrap.cpp).  tests/test_reference_build.py requires this oracle to equal what
 * SceneFlowConstructor::construct() and ClustererNodelet::dataCB() publish, bit for bit, on the golden fixtures
 * (which are generated from that build), randomised and adversarial inputs; tests/test_gpu_reference.py compares
 * the CUDA path with it directly.  What remains restated rather than executed is the arithmetic inside the
 * un-vendored dependencies (image_geometry projection, tf2_eigen/Eigen products and norms, pcl::getMinMax3D, cv::Vec
 * dot, the COLORMAP_HSV table taken from cv2): see the header of oracle/ref_shim/ref_shim.h and SURVEY.md Appendix B.
 *
 * The struct layouts deliberately equal those in include/mod_b200.h so that a test can hand the same
 * ctypes buffers to both sides, but the two headers share no code.
 */
#ifndef MOD_ORACLE_H_
#define MOD_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* left CameraInfo P[0],P[5],P[2],P[6],P[3],P[7]  (image_geometry::PinholeCameraModel fx() fy() cx() cy() Tx() Ty()) */
typedef struct { double fx, fy, cx, cy, Tx, Ty; } orc_camera;

/* stereo_msgs/DisparityImage: scalars are float32 on the wire; image is 32FC1 with a byte row stride */
typedef struct {
  float f, T, min_disparity, max_disparity;
  const float* data;
  int32_t step_bytes;
} orc_disparity;

/* geometry_msgs/Transform, exactly as stored at scene_flow_constructor.cpp:248-249 */
typedef struct { double tx, ty, tz, qx, qy, qz, qw; } orc_transform;

/* moving_object_msgs/MovingObject content (the reference computes all of it in float, then widens) */
typedef struct {
  int32_t id;
  float center[3];
  float velocity[3];
  float bbox[3];
  int32_t n_points;
} orc_object;

/* per-pixel class of the constructVelocityPC state machine */
enum { ORC_CLASS_NAN = 0,      /* I   : all six fields NaN                          */
       ORC_CLASS_XYZ_ONLY = 1, /* II  : xyz valid, velocity NaN                     */
       ORC_CLASS_STATIC = 2,   /* III : xyz valid, velocity = 0                     */
       ORC_CLASS_MOVING = 3 }; /* IV  : xyz valid, velocity = (now - prev) / dt     */

/* bits of the optional near-threshold side channel */
enum { ORC_NEAR_FLOW_DIFF = 1,  /* |flow - static_flow| within 1 float ulp of dynamic_flow_diff */
       ORC_NEAR_SPEED = 2 };    /* |v| within 1 float ulp of dynamic_speed                       */

/*
 * SceneFlowConstructor::construct() minus ROS publishing (scene_flow_constructor.cpp:91-147).
 * out_points32: w*h records of 8 floats {x,y,z,0,vx,vy,vz,0} (PointXYZVelocity, PointCloud2.data compatible).
 * opt_depth (w*h), opt_static_flow (w*h*2), opt_class (w*h) may be NULL.
 * Returns 0, or <0 on bad arguments.
 */
int orc_scene_flow(int w, int h, const orc_camera* cam, const orc_disparity* now, const orc_disparity* prev,
                   const float* flow_xy, int32_t flow_step_bytes, const orc_transform* prev2now, double dt_sec,
                   int32_t dynamic_flow_diff, float* out_points32, float* opt_depth, float* opt_static_flow,
                   uint8_t* opt_class);

/*
 * ClustererNodelet::clustering() + publishMovingObjects() (clusterer_nodelet.cpp:85-95,324-343).
 * opt_dynamic_map (w*h, 0/1), opt_labels (w*h, -1 = none), opt_near (w*h) may be NULL.
 * Returns 0; -2 if more than max_objs objects were produced (n_objs then holds the true count).
 */
int orc_cluster(int w, int h, const float* points32, int32_t cluster_size, double depth_diff, double dynamic_speed,
                int32_t neighbor_distance, uint8_t* opt_dynamic_map, int32_t* opt_labels, uint8_t* opt_near,
                orc_object* objs, int32_t max_objs, int32_t* n_objs);

/* near-threshold flags for the constructor's 2-D residual gate; same inputs as orc_scene_flow */
int orc_scene_flow_near(int w, int h, const orc_camera* cam, const orc_disparity* now, const orc_disparity* prev,
                        const float* flow_xy, int32_t flow_step_bytes, const orc_transform* prev2now,
                        int32_t dynamic_flow_diff, uint8_t* out_near);

/* ~clusters_image: publishClustersImage (clusterer_nodelet.cpp:293-322) + ColorSet (color_set.cpp:11-26) on a final label
 * map; colormap_bgr768 is OpenCV's COLORMAP_HSV table (256 x BGR).  Returns 0, -3 where cv::Mat::at would throw. */
int orc_clusters_image(int w, int h, const int32_t* labels, int32_t number_of_clusters, const uint8_t* colormap_bgr768,
                       uint8_t* out_bgr8);

/* ~clusters marker point order: clusterMap2IndicesCluster (clusterer_nodelet.cpp:97-117), u outer / v inner.
 * out_offsets has number_of_clusters + 1 entries, out_indices one entry per labelled pixel. */
int orc_cluster_members(int w, int h, const int32_t* labels, int32_t number_of_clusters, int32_t* out_offsets,
                        int32_t* out_indices);

/* image_crop::ImageCrop::cameraCallback (image_crop/src/image_crop.cpp:20-43): centre crop to target_w x target_h of an
 * 8-bit image converted to bgr8 (encoding 0 bgr8, 1 rgb8, 2 bgra8, 3 rgba8, 4 mono8) and the CameraInfo fix-up
 * (K[2], K[5], P[2], P[6] shifted by the crop origin; width / height replaced).  Returns 0, or -12 where the
 * reference's cv::Mat ROI assertion throws (crop window not inside the image). */
int orc_image_crop(int w, int h, int encoding, const uint8_t* data, int32_t step_bytes, int target_w, int target_h,
                   uint8_t* out_bgr8);
void orc_crop_camera_info(uint32_t w, uint32_t h, int target_w, int target_h, const double* K9, const double* P12,
                          double* out_K9, double* out_P12);

/* quaternion -> 3x3 rotation (row-major, 9 doubles), Eigen::Quaterniond::toRotationMatrix order of operations */
void orc_quat_to_matrix(const orc_transform* t, double* R9);

/* label-equivalence table, restated (lookup_table.h:10-33); used by the _ref cross-check test.
 * ops: array of (op, a, b) int triples, op 0 = addLabel, 1 = link(a,b), 2 = lookup(a) -> appended to out. */
int orc_lookup_table_run(int size, const int32_t* ops, int n_ops, int32_t* out, int max_out);

/*
 * CPU baseline harness: runs the full reference-structured path (constructor + clusterer) over n_pairs
 * frame pairs with n_threads std::threads (one pair per thread at a time), repeating `reps` times.
 * Pair i uses disparity frames prev_index[i] / now_index[i] out of `disp_frames` (n_frames * w * h floats, dense)
 * and flow pair i (w*h*2 floats, dense), transform i, dt[i].
 * Returns wall seconds of the best repetition in *best_sec and the total number of objects in *n_objects.
 */
int orc_bench_pipeline(int w, int h, const orc_camera* cam, float f, float T, float min_disp, float max_disp,
                       const float* disp_frames, const int32_t* prev_index, const int32_t* now_index,
                       const float* flows, const orc_transform* transforms, const double* dt, int n_pairs,
                       int32_t dynamic_flow_diff, int32_t cluster_size, double depth_diff, double dynamic_speed,
                       int32_t neighbor_distance, int n_threads, int reps, int stage_mask /*1=constructor,2=clusterer*/,
                       double* best_sec, int64_t* n_objects);

#ifdef __cplusplus
}
#endif
#endif /* MOD_ORACLE_H_ */
