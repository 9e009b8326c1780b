/*
 * mod_b200.h -- C ABI of the B200-native scene-flow / motion-segmentation hot path.
 *
 * The reference (ActiveIntelligentSystemsLab/moving_object_detector) has no FFI: its seams are C++ method
 * boundaries inside two ROS processes.  This header introduces a thin C ABI exactly at those two seams so the
 * unchanged ROS host side (node / nodelet, dynamic_reconfigure, topics, messages) can call the CUDA path:
 *
 *   seam 1  SceneFlowConstructor::construct()                scene_flow_constructor/src/scene_flow_constructor.cpp:91-147
 *           (DisparityImageProcessor::toPointCloud x2         disparity_image_proc/src/disparity_image_processor.cpp:86-103,
 *            transformPCPreviousToNow :409-429, calculateStaticOpticalFlow :65-89, constructVelocityPC :149-212)
 *   seam 2  ClustererNodelet::clustering() + publishMovingObjects()
 *                                                             scene_flow_clusterer/src/clusterer_nodelet.cpp:85-95,324-343
 *
 * All entry points are extern "C", take plain pointers and sizes, never throw, and return a mod_status.
 * Parameters that dynamic_reconfigure can change (cfg/SceneFlowConstructor.cfg:8-9, cfg/Clusterer.cfg:8-11) are
 * passed BY VALUE PER CALL, never latched in the context.
 *
 * Threading: one call at a time per context (the reference joins the previous construct() worker before starting
 * the next, scene_flow_constructor.cpp:389-392, and the clusterer callback queue is single-threaded,
 * clusterer_nodelet.cpp:25,35).  Contexts are independent (one per GPU / per stream).
 *
 * There is NO CPU fallback: every compute entry point returns MOD_ERR_CUDA / MOD_ERR_NO_DEVICE when no sm_100a
 * device is usable.
 */
#ifndef MOD_B200_H_
#define MOD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOD_B200_ABI_VERSION 1

typedef struct mod_ctx mod_ctx;

typedef enum {
  MOD_OK = 0,
  MOD_ERR_INVALID_ARG = -1,   /* null pointer, non-positive size, size larger than the context was created for   */
  MOD_ERR_CAPACITY = -2,      /* more moving objects than max_objs; *n_objs holds the true count, objs is filled   */
                              /* up to max_objs (the reference's message array is unbounded)                       */
  MOD_ERR_INVALID_PARAM = -3, /* parameter value at which the reference itself throws / is undefined               */
                              /* (cluster_size == 0: cluster.at(0) on an empty cluster, clusterer_nodelet.cpp:174; */
                              /*  neighbor_distance < 0), or outside the dynamic_reconfigure range                 */
                              /* (neighbor_distance > 10, cfg/Clusterer.cfg:11)                                    */
  MOD_ERR_CUDA = -4,          /* a CUDA runtime call failed; see mod_last_error()                                  */
  MOD_ERR_NO_DEVICE = -5,     /* no CUDA device / not compute capability 10.x                                      */
  MOD_ERR_CLUSTER_SLOTS = -6  /* more surviving clusters in one frame than the context's max_clusters              */
} mod_status;

/* left CameraInfo: P[0], P[5], P[2], P[6], P[3], P[7] -- what image_geometry::PinholeCameraModel::fx() fy() cx()
 * cy() Tx() Ty() return after fromCameraInfo (disparity_image_processor.cpp:7,14; scene_flow_constructor.cpp:370-371) */
typedef struct { double fx, fy, cx, cy, Tx, Ty; } mod_camera;

/* stereo_msgs/DisparityImage as consumed at disparity_image_processor.cpp:3-15: float32 scalars, 32FC1 image with a
 * byte row stride (image.step).  `data` is a HOST pointer in the host entry points. */
typedef struct {
  float f, T, min_disparity, max_disparity;
  const float* data;
  int32_t step_bytes;
} mod_disparity;

/* geometry_msgs/Transform exactly as produced at scene_flow_constructor.cpp:248-249 (translation + quaternion) */
typedef struct { double tx, ty, tz, qx, qy, qz, qw; } mod_transform;

/* dynamic_reconfigure SceneFlowConstructor (cfg/SceneFlowConstructor.cfg:8-9); max_color_velocity is carried for
 * interface completeness -- no computation in the reference reads it (scene_flow_constructor.cpp:406). */
typedef struct { int32_t dynamic_flow_diff; double max_color_velocity; } mod_sf_params;

/* dynamic_reconfigure Clusterer (cfg/Clusterer.cfg:8-11; members at clusterer_nodelet.h:54-57) */
typedef struct {
  int32_t cluster_size;
  double depth_diff;
  double dynamic_speed;
  int32_t neighbor_distance;
} mod_cl_params;

/* moving_object_msgs/MovingObject (msg/MovingObject.msg:3-7).  The reference computes centre, velocity and
 * bounding box in float and widens to the message's float64 fields (clusterer_nodelet.cpp:151-181), so float is
 * lossless here; orientation is always (0,0,0,1) (clusterer_nodelet.cpp:163-166). */
typedef struct {
  int32_t id;
  float center[3];
  float velocity[3];
  float bbox[3];
  int32_t n_points;
} mod_object;

/* One frame pair of the device-resident batched path.  All pointers are DEVICE pointers to dense images
 * (row stride = w elements): disparity 32FC1, flow 32FC2.  In a sequence disp_prev of pair i is simply
 * disp_now of pair i-1 (scene_flow_constructor.cpp:398). */
typedef struct {
  const float* disp_now;
  const float* disp_prev;
  const float* flow_xy;
  float f_now, T_now, min_disparity_now, max_disparity_now;
  float f_prev, T_prev, min_disparity_prev, max_disparity_prev;
  mod_transform prev2now;
  double dt_sec; /* (stamp_now - stamp_prev).toSec(), scene_flow_constructor.cpp:162-164 */
} mod_pair_dev;

/* Per-frame counters of the last call (diagnostics; the reference only logs wall time). */
typedef struct {
  int32_t n_dynamic;   /* pixels in dynamic_map_ (clusterer_nodelet.cpp:40-54)                 */
  int32_t n_clusters;  /* clusters that survived removeSmallClusters (clusterer_nodelet.cpp:354-393) */
  int32_t n_objects;   /* accepted MovingObjects                                                */
  int32_t status;      /* MOD_OK, MOD_ERR_CAPACITY or MOD_ERR_CLUSTER_SLOTS for this frame      */
} mod_frame_stats;

/* ---- context --------------------------------------------------------------------------------------------- */

/* Creates a context on CUDA device `device` able to process frames up to max_w x max_h, max_batch frame pairs per
 * batched call and max_clusters surviving clusters per frame (0 = default 4096, at most 32768; a frame with more
 * fails with MOD_ERR_CLUSTER_SLOTS and the caller re-creates the context with a larger value -- the host classes do).
 * Owns device workspaces, one stream and pinned staging.  Returns MOD_ERR_NO_DEVICE when there is no usable sm_100a GPU. */
int mod_create(int device, int max_w, int max_h, int max_batch, int max_clusters, mod_ctx** out_ctx);
void mod_destroy(mod_ctx* ctx);
/* Message of the last error on this context ("" if none); valid until the next call on the context. */
const char* mod_last_error(const mod_ctx* ctx);
int mod_abi_version(void);
/* Number of CUDA kernels this context has launched so far (evidence that the GPU path ran; used by bench.py). */
unsigned long long mod_launch_count(const mod_ctx* ctx);
/* The context's own non-blocking stream (a cudaStream_t), usable as the `stream` argument of the *_dev calls. */
void* mod_stream(const mod_ctx* ctx);

/* Numerics self-check used by the tests: counts, into *d_n_bad, the operand pairs (d_a[i], d_b[i]) for which the
 * kernel's shared-reciprocal double division (one reciprocal per divisor Z / dt, scene_flow_constructor.cpp:84,200-202)
 * differs bitwise from IEEE double division (__ddiv_rn, i.e. the reference's divsd). */
int mod_selfcheck_ddiv(mod_ctx* ctx, void* stream, const double* d_a, const double* d_b, size_t n,
                       unsigned long long* d_n_bad);

/* ---- O(1) host-side parameter preparation, exposed so the CPU test-suite can pin it without a GPU ---------- */
/* tf2::transformToEigen + Eigen::Quaterniond::toRotationMatrix (scene_flow_constructor.cpp:411; no normalisation):
 * row-major R9 and translation t3, in the order of operations of Eigen 3.3. */
void mod_host_rotation(const mod_transform* t, double* R9, double* t3);
/* float equivalents of the reference's float-vs-double comparisons (clusterer_nodelet.cpp:51,176,198):
 * (double)a >= th  <=>  a >= mod_host_threshold_ge(th);   (double)a > th  <=>  a > mod_host_threshold_gt(th). */
float mod_host_threshold_ge(double th);
float mod_host_threshold_gt(double th);
/* sqrt-free form of "std::sqrt(s) >= thr" for a float sum of squares s >= +0 (scene_flow_constructor.cpp:198,
 * clusterer_nodelet.cpp:51): the smallest float s2 with sqrtf(s2) >= thr, so that sqrtf(s) >= thr <=> s >= s2
 * (sqrtf is correctly rounded and monotonic; NaN compares false on both sides). */
float mod_host_sq_threshold(float thr);

/* Pinned host memory for message buffers (lets H2D/D2H run at full PCIe rate); optional. */
void* mod_host_alloc(size_t bytes);
void mod_host_free(void* p);
/* Pin / unpin a buffer the caller owns and keeps across frames (e.g. a message object that is reused): copies from /
 * to it then are single DMAs.  Unregister before the memory is released.  MOD_OK or MOD_ERR_CUDA. */
int mod_host_register(void* p, size_t bytes);
int mod_host_unregister(void* p);
/* Ordinary (pageable) buffers - the std::vector storage of ROS messages (scene_flow_constructor.cpp:351-362,
 * clusterer_nodelet.cpp:225-226) - are staged through pinned memory by n host threads (the caller's included) while
 * the copy engine moves the previous chunk; process-wide.  Default: $MOD_B200_COPY_THREADS, else min(4, hardware
 * threads / 2).  n <= 0: plain cudaMemcpyAsync (the driver stages on one thread). */
void mod_set_host_copy_threads(int n);
int mod_host_copy_threads(void);
/* Frame identities for the NEXT mod_scene_flow / mod_pipeline call on this context (then forgotten): caller-chosen
 * non-zero ids of the two disparity images' contents, e.g. the messages' stamps in nanoseconds.  The image uploaded
 * as `now` stays on the device; when the following call names the same id as `prev_id` (same size) it is not uploaded
 * again - the reference keeps `disparity_previous_ = disparity_now_` the same way (scene_flow_constructor.cpp:398).
 * The ids are a promise about contents: equal ids, equal images.  Without ids (or with 0) both images are uploaded. */
int mod_set_frame_ids(mod_ctx* ctx, unsigned long long now_id, unsigned long long prev_id);
unsigned long long mod_resident_hits(const mod_ctx* ctx); /* uploads saved so far (diagnostic) */

/* ---- seam 1: SceneFlowConstructor::construct() minus ROS publishing (scene_flow_constructor.cpp:91-147) ------ */
/* Host buffers in, host buffers out, synchronous.
 *  out_points32 : w*h PointXYZVelocity records of 32 bytes {x,y,z,0,vx,vy,vz,0} = PointCloud2.data with
 *                 point_step 32, row_step 32*w (pcl_point_xyz_velocity.h:8-34).  May be NULL.
 *  opt_depth    : w*h float, ~depth topic (disparity_image_processor.cpp:105-121).  May be NULL.
 *  opt_static_flow : w*h*2 float, ~synthetic_optical_flow (scene_flow_constructor.cpp:65-89).  May be NULL.
 * Degrade rules of the reference (scene_flow_constructor.cpp:104-133): the caller must not call this without flow,
 * previous disparity or transform -- in those cases the reference publishes no scene_flow at all. */
int mod_scene_flow(mod_ctx* ctx, int w, int h, const mod_camera* cam, const mod_disparity* now,
                   const mod_disparity* prev, const float* flow_xy, int32_t flow_step_bytes,
                   const mod_transform* prev2now, double dt_sec, const mod_sf_params* params, void* out_points32,
                   float* opt_depth, float* opt_static_flow);

/* ---- seam 2: ClustererNodelet::clustering() + publishMovingObjects() (clusterer_nodelet.cpp:85-95,324-343) --- */
/*  points32   : w*h 32-byte records (PointCloud2.data of ~scene_flow), host.
 *  opt_labels : w*h int32, the final cluster_map_ (-1 = NOT_BELONGED_), ids in the reference's order.  May be NULL.
 *  objs/n_objs: MovingObjectArray.moving_object_array in id order. */
int mod_cluster(mod_ctx* ctx, int w, int h, const void* points32, const mod_cl_params* params, int32_t* opt_labels,
                mod_object* objs, int32_t max_objs, int32_t* n_objs);
/* The same on the cloud the last mod_scene_flow call on this context left on the device (whether or not it was
 * downloaded): constructor and clusterer in one process, the reference's two nodes without the PointCloud2 round
 * trip between them (scene_flow_constructor.cpp:141-142 -> clusterer_nodelet.cpp:221-242).  MOD_ERR_INVALID_ARG when
 * there is no such cloud of this size. */
int mod_cluster_resident(mod_ctx* ctx, int w, int h, const mod_cl_params* params, int32_t* opt_labels,
                         mod_object* objs, int32_t max_objs, int32_t* n_objs);

/* Counters of the frame the last mod_cluster / mod_cluster_resident / mod_pipeline call on this context processed:
 * n_clusters is number_of_clusters_ after removeSmallClusters (clusterer_nodelet.cpp:354-393), which sizes the colour
 * set of ~clusters and ~clusters_image (:269-322, color_set.cpp:16-26); n_objects counts the accepted MovingObjects
 * (:176).  MOD_ERR_INVALID_ARG before the first such call. */
int mod_last_frame_stats(const mod_ctx* ctx, mod_frame_stats* out);
/* Surviving clusters per frame this context was created for (mod_create's max_clusters after defaulting / clamping). */
int mod_max_clusters(const mod_ctx* ctx);

/* ---- fused host entry: seam 1 + seam 2 with the cloud staying on the GPU ------------------------------------ */
/* opt_out_points32 is only copied back when non-NULL (the reference publishes ~scene_flow only when it has
 * subscribers, scene_flow_constructor.cpp:141-142). */
int mod_pipeline(mod_ctx* ctx, int w, int h, const mod_camera* cam, const mod_disparity* now,
                 const mod_disparity* prev, const float* flow_xy, int32_t flow_step_bytes,
                 const mod_transform* prev2now, double dt_sec, const mod_sf_params* sf_params,
                 const mod_cl_params* cl_params, void* opt_out_points32, int32_t* opt_labels, mod_object* objs,
                 int32_t max_objs, int32_t* n_objs);

/* ---- device-resident batched path (benchmark / multi-frame drivers) ----------------------------------------- */
/* All d_* pointers are device memory; `stream` is a cudaStream_t used as given (NULL = CUDA's legacy default stream,
 * as in every CUDA API; pass mod_stream(ctx) for the context's own stream).  The calls are ASYNCHRONOUS on that
 * stream and do not synchronise the host.
 *  d_points32 : batch * frame_stride_points records of 32 bytes (frame_stride_points >= w*h).
 *  d_mask     : batch * w*h bytes, the clusterer's dynamic_map_ (|v| >= dynamic_speed), fused into the same kernel.
 *               May be NULL (then no mask is produced and dynamic_speed is ignored).
 *  d_opt_depth / d_opt_static_flow : batch * w*h (*2) floats or NULL. */
int mod_scene_flow_batched_dev(mod_ctx* ctx, void* stream, int w, int h, int batch, const mod_camera* cam,
                               const mod_pair_dev* pairs /* host array[batch] */, const mod_sf_params* params,
                               double dynamic_speed, void* d_points32, size_t frame_stride_points, uint8_t* d_mask,
                               float* d_opt_depth, float* d_opt_static_flow);

/*  d_points32   : 32-byte aligned (records are moved as 256-bit / 128-bit vectors), in both calls.
 *  d_mask       : batch * w*h bytes as written by mod_scene_flow_batched_dev (16-byte aligned), or NULL to derive it
 *                 from the cloud (calculateDynamicMap, clusterer_nodelet.cpp:40-54).
 *  d_opt_labels : batch * w*h int32 or NULL.
 *  d_objs       : batch * max_objs records; d_stats: batch records (n_objects, status per frame). */
int mod_cluster_batched_dev(mod_ctx* ctx, void* stream, int w, int h, int batch, const void* d_points32,
                            size_t frame_stride_points, const uint8_t* d_mask, const mod_cl_params* params,
                            int32_t* d_opt_labels, mod_object* d_objs, int32_t max_objs, mod_frame_stats* d_stats);

/* ---- ~clusters_image ------------------------------------------------------------------------------------------ */
/* ClustererNodelet::publishClustersImage (scene_flow_clusterer/src/clusterer_nodelet.cpp:293-322) with ColorSet
 * (src/color_set.cpp:16-26): a BGR8 image, black where the label is -1 (NOT_BELONGED_), otherwise entry
 * label * 255 / n_clusters of OpenCV's COLORMAP_HSV table (cv::applyColorMap, an un-vendored dependency; the table is
 * generated from cv2 by tests/golden/make_colormap.py).  n_clusters is the number of surviving clusters of the frame
 * (mod_frame_stats.n_clusters / *n_objs of mod_cluster), as in color_set_.resize(number_of_clusters_).
 * Host seam: labels and out_bgr8 (w*h*3 bytes) are host pointers. */
int mod_clusters_image(mod_ctx* ctx, int w, int h, const int32_t* labels, int32_t n_clusters, uint8_t* out_bgr8);
/* Device-resident batch: d_labels batch*w*h, d_stats batch records (as written by mod_cluster_batched_dev),
 * d_out_bgr8 batch*w*h*3 bytes.  Asynchronous on `stream`. */
int mod_clusters_image_batched_dev(mod_ctx* ctx, void* stream, int w, int h, int batch, const int32_t* d_labels,
                                   const mod_frame_stats* d_stats, uint8_t* d_out_bgr8);
/* The 256-entry COLORMAP_HSV table as 768 bytes (b, g, r). */
void mod_host_hsv_colormap(uint8_t* bgr768);

/* ---- image_crop (upstream glue) --------------------------------------------------------- */
/* image_crop::ImageCrop::cameraCallback (image_crop/src/image_crop.cpp:20-43): the camera image, converted to bgr8 as
 * cv_bridge::toCvShare(msg, "bgr8") does (:23; channel shuffles only), is cropped to target_w x target_h around the
 * image centre: origin ((width - target_w) / 2, (height - target_h) / 2) in the reference's UNSIGNED 32-bit arithmetic
 * (:24).  A window that does not lie inside the image makes the reference's cv::Mat ROI assertion throw (or, where
 * cv::Rect's int sum overflows, read out of bounds); here it returns MOD_ERR_INVALID_ARG. */
enum { MOD_ENC_BGR8 = 0, MOD_ENC_RGB8 = 1, MOD_ENC_BGRA8 = 2, MOD_ENC_RGBA8 = 3, MOD_ENC_MONO8 = 4 };
/* Host seam: data = sensor_msgs/Image.data (h rows of step_bytes), out_bgr8 = target_w*target_h*3 bytes, dense. */
int mod_image_crop(mod_ctx* ctx, int w, int h, int encoding, const uint8_t* data, int32_t step_bytes, int target_w,
                   int target_h, uint8_t* out_bgr8);
/* Device-resident batch: frame i starts at d_data + i*frame_stride_bytes; d_out_bgr8 = batch dense crops.
 * Asynchronous on `stream`.  The source needs w, h <= the context's max_w, max_h. */
int mod_image_crop_batched_dev(mod_ctx* ctx, void* stream, int w, int h, int batch, int encoding, const uint8_t* d_data,
                               size_t frame_stride_bytes, int32_t step_bytes, int target_w, int target_h,
                               uint8_t* d_out_bgr8);
/* The CameraInfo fix-up of the same callback (:32-41), host only: K[2], P[2] = K[2] - (width - target_w) / 2 and
 * K[5], P[6] = K[5] - (height - target_h) / 2 (both taken from K, unsigned quotients), everything else copied; the
 * caller sets width / height to the target size.  Returns MOD_ERR_INVALID_ARG for a window outside the image. */
int mod_host_crop_camera_info(uint32_t w, uint32_t h, int target_w, int target_h, const double* K9, const double* P12,
                              double* out_K9, double* out_P12);

/* ---- MONO8 conversion (upstream glue) ------------------------------------------------------ */
/* SceneFlowConstructor::estimateCameraMotion (scene_flow_constructor/src/scene_flow_constructor.cpp:220-221):
 * cv_bridge::toCvCopy(image, MONO8) of the left and right camera image in front of the viso2 odometer, i.e.
 * cv::cvtColor(COLOR_{BGR,RGB,BGRA,RGBA}2GRAY) on 8-bit pixels (a mono8 source is copied):
 *     grey = (cB*B + cG*G + cR*R + (1 << (shift-1))) >> shift.
 * The constants belong to the OpenCV the node is linked against, so the caller names them:
 *   MOD_GRAY_OPENCV3  (cR, cG, cB, shift) = (4899, 9617, 1868, 14)  -- OpenCV 2.4 / 3.x, the reference's ROS Melodic build
 *   MOD_GRAY_OPENCV4  (cR, cG, cB, shift) = (9798, 19235, 3735, 15) -- OpenCV 4.x
 * Output: dense mono8, h rows of w bytes (what viso2's process() receives as dims = {w, h, w}). */
enum { MOD_GRAY_OPENCV3 = 0, MOD_GRAY_OPENCV4 = 1 };
/* sensor_msgs/Image.encoding -> MOD_ENC_* ("bgr8", "rgb8", "bgra8", "rgba8", "mono8"); MOD_ERR_INVALID_ARG for the
 * encodings neither image_crop nor the MONO8 conversion take.  Host only. */
int mod_host_encoding(const char* ros_encoding);
/* Host seam: data = sensor_msgs/Image.data (h rows of step_bytes) in `encoding` (MOD_ENC_*). */
int mod_image_to_mono8(mod_ctx* ctx, int w, int h, int encoding, const uint8_t* data, int32_t step_bytes, int rule,
                       uint8_t* out_mono8);
/* Device-resident batch: frame i starts at d_data + i*frame_stride_bytes; d_out_mono8 = batch dense images.
 * Asynchronous on `stream`. */
int mod_image_to_mono8_batched_dev(mod_ctx* ctx, void* stream, int w, int h, int batch, int encoding,
                                   const uint8_t* d_data, size_t frame_stride_bytes, int32_t step_bytes, int rule,
                                   uint8_t* d_out_mono8);

/* ---- static-flow filter (diagnostics) ------------------------------------------------------------------------ */
/* The scene-flow kernel decides the residual gate (scene_flow_constructor.cpp:191-198) from a float evaluation of
 * the static flow with a rigorous error bound and re-evaluates undecided pixels with the reference's exact
 * arithmetic, so outputs do not depend on this switch; it exists so tests can compare both kernels.  Default: on.
 * A request for the static-flow image (opt_static_flow) always takes the exact kernel. */
int mod_set_static_flow_filter(mod_ctx* ctx, int enabled);
/* Number of pixels the filter could not decide since the last reset (synchronises the device). */
int mod_filter_fallbacks(mod_ctx* ctx, int reset, unsigned long long* count);

/* ---- host batched streaming path (end-to-end: H2D + kernels + D2H inside the call) --------------------------- */
/* A sequence of n_frames disparity images (host, dense, preferably pinned) forms n_frames-1 frame pairs:
 * pair i = (prev = frame i, now = frame i+1), flows[i] and transforms[i]/dt[i] belong to pair i.
 * Each disparity frame is uploaded once; uploads, kernels and downloads of successive pairs overlap on separate
 * streams.  opt_out_points32 (n_pairs * w*h*32 bytes) is downloaded only when non-NULL. */
int mod_pipeline_sequence_host(mod_ctx* ctx, int w, int h, int n_frames, const mod_camera* cam, float f, float T,
                               float min_disparity, float max_disparity, const float* disp_frames,
                               const float* flows, const mod_transform* transforms, const double* dt_sec,
                               const mod_sf_params* sf_params, const mod_cl_params* cl_params,
                               void* opt_out_points32, mod_object* objs, int32_t max_objs_per_pair,
                               mod_frame_stats* stats /* host array[n_frames-1] */);

#ifdef __cplusplus
}
#endif
#endif /* MOD_B200_H_ */
