// scene_flow_host.h -- C++ host side of the B200 path: the two classes a ROS maintainer swaps in for the reference's
// compute methods.  They keep the reference's names, parameters, degrade rules and message layouts and call the GPU
// only through the C ABI of include/mod_b200.h (no CUDA types here; link with -lmod_b200).
//
//   scene_flow_constructor::SceneFlowConstructorB200::construct()   <- SceneFlowConstructor::construct
//        scene_flow_constructor/src/scene_flow_constructor.cpp:91-147 (+ :149-212, :65-89, :409-429 and
//        disparity_image_proc/src/disparity_image_processor.cpp:86-121, all inside ONE mod_scene_flow / mod_pipeline call)
//   scene_flow_clusterer::ClustererB200::dataCB()                   <- ClustererNodelet::dataCB
//        scene_flow_clusterer/src/clusterer_nodelet.cpp:221-242 (clustering :85-95 + publishMovingObjects :324-343)
//
// ROS plumbing (NodeHandle, message_filters, image_transport, dynamic_reconfigure servers, nodelet/pluginlib export)
// stays exactly as in the reference and is not duplicated here: INTEGRATION.md shows the few lines that change in
// scene_flow_constructor.cpp / clusterer_nodelet.cpp.  "Publishers" are modelled as sinks with a subscriber count so
// that the reference's getNumSubscribers() gating is preserved.
#pragma once
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "../../include/mod_b200.h"
#include "ros_compat.h"

#ifdef MOD_HAVE_ROS
// the dynamic_reconfigure-generated parameter structs (cfg/SceneFlowConstructor.cfg, cfg/Clusterer.cfg)
#include <scene_flow_clusterer/ClustererConfig.h>
#include <scene_flow_constructor/SceneFlowConstructorConfig.h>
#endif

namespace scene_flow_constructor {

#ifndef MOD_HAVE_ROS
// cfg/SceneFlowConstructor.cfg:8-9 (generated SceneFlowConstructorConfig has exactly these two members)
struct SceneFlowConstructorConfig {
  int dynamic_flow_diff = 5;
  double max_color_velocity = 1.0;
};
#endif

template <typename Msg>
struct Sink {  // ros::Publisher stand-in: publish() + getNumSubscribers()
  std::function<void(const Msg&)> publish;
  int num_subscribers = 0;
  int getNumSubscribers() const { return publish ? num_subscribers : 0; }
};

class SceneFlowConstructorB200 {
 public:
  // max_w / max_h size the device workspaces; device = CUDA ordinal.  Throws std::runtime_error when no sm_100a GPU
  // is usable (there is no CPU fallback).
  SceneFlowConstructorB200(int max_w, int max_h, int device = 0);
  ~SceneFlowConstructorB200();
  SceneFlowConstructorB200(const SceneFlowConstructorB200&) = delete;
  SceneFlowConstructorB200& operator=(const SceneFlowConstructorB200&) = delete;

  // dynamic_reconfigure callback (scene_flow_constructor.cpp:401-407): latches the two parameters
  void reconfigureCB(const SceneFlowConstructorConfig& config, uint32_t level);

  // left CameraInfo as latched at scene_flow_constructor.cpp:368-374 (PinholeCameraModel::fromCameraInfo: binning
  // and ROI offsets are applied to P exactly as image_geometry does)
  void setCameraInfo(const sensor_msgs::CameraInfo& left_camera_info);
  // frame_id the ~depth image is stamped with.  setCameraInfo takes it from the CameraInfo header; the reference node
  // takes it from the first left image (:373) - a node that wants exactly that calls this afterwards.
  void setCameraFrame(const std::string& frame_id) { camera_frame_id_ = frame_id; }

  // Same contract as SceneFlowConstructor::construct (:91-147): any argument may be null.
  //   ~optical_flow is published whenever left_flow is set and subscribed           (:99-100)
  //   ~depth needs disparity_now and a subscriber                                     (:111-119)
  //   nothing else happens without left_flow                                          (:122-123)
  //   ~scene_flow / ~synthetic_optical_flow need disparity_now, disparity_previous and transform_prev2now (:127-146)
  // Returns true when the scene-flow cloud was computed (whether or not anybody subscribed to it).
  bool construct(const std::shared_ptr<const stereo_msgs::DisparityImage>& disparity_now,
                 const std::shared_ptr<const stereo_msgs::DisparityImage>& disparity_previous,
                 const std::shared_ptr<const sensor_msgs::Image>& left_flow /* 32FC2 */,
                 const std::shared_ptr<const geometry_msgs::Transform>& transform_prev2now);

  Sink<sensor_msgs::Image> optflow_pub_;          // ~optical_flow            32FC2
  Sink<sensor_msgs::Image> depth_pub_;            // ~depth                   32FC1
  Sink<sensor_msgs::PointCloud2> pc_with_velocity_pub_;  // ~scene_flow       PointXYZVelocity, point_step 32
  Sink<sensor_msgs::Image> static_flow_pub_;      // ~synthetic_optical_flow  32FC2
  // Optional: the clusterer in the same process (mod_cluster_resident on the cloud construct() just left on the GPU),
  // so that ~scene_flow's 32 B/px never cross the PCIe link or a ROS socket unless somebody subscribes to it.  When
  // this sink is set and subscribed, construct() publishes what ClustererNodelet::dataCB would have published on
  // ~moving_objects for that cloud, after its own topics.  clusterer_* are the Clusterer.cfg parameters.
  Sink<moving_object_msgs::MovingObjectArray> moving_objects_pub_;
  int clusterer_cluster_size = 2500;
  double clusterer_depth_diff = 0.15, clusterer_dynamic_speed = 0.3;
  int clusterer_neighbor_distance = 4;

  // cv_bridge::toCvCopy(image, MONO8) of estimateCameraMotion (:220-221): dense width x height bytes for viso2's
  // process().  rule = MOD_GRAY_OPENCV3 (the reference's OpenCV 3.2 constants) or MOD_GRAY_OPENCV4.  false + lastError()
  // for an encoding cv_bridge would have to convert differently (16-bit, Bayer, ...).
  bool toMono8(const sensor_msgs::Image& image, std::vector<uint8_t>& mono8, int rule = MOD_GRAY_OPENCV3);

  const std::string& lastError() const { return last_error_; }
  mod_ctx* context() const { return ctx_; }

 private:
  mod_ctx* ctx_ = nullptr;
  int dynamic_flow_diff_ = 5;
  double max_color_velocity_ = 1.0;
  bool have_camera_ = false;
  mod_camera camera_{};
  std::string camera_frame_id_;
  std::string last_error_;
  // output messages, kept across frames (construct() is never re-entered: one call at a time per object)
  sensor_msgs::PointCloud2 cloud_msg_;
  sensor_msgs::Image depth_msg_, static_flow_msg_;
  std::vector<float> zero_flow_;
  std::vector<mod_object> objs_;
  int max_w_ = 0, max_h_ = 0, device_ = 0;
  // the disparity message whose image the last call left on the device as `now`, and the id it was uploaded under
  std::shared_ptr<const stereo_msgs::DisparityImage> resident_now_;
  unsigned long long resident_now_id_ = 0, frame_counter_ = 0;
};

// pcl::toROSMsg(pcl::PointCloud<pcl::PointXYZVelocity>) layout (pcl_point_xyz_velocity.h:8-34): fills everything
// except `data` (which the GPU path writes directly) for an organized w x h cloud.
void describeVelocityCloud(sensor_msgs::PointCloud2& msg, uint32_t width, uint32_t height);

}  // namespace scene_flow_constructor

namespace scene_flow_clusterer {

#ifndef MOD_HAVE_ROS
// cfg/Clusterer.cfg:8-11
struct ClustererConfig {
  int cluster_size = 2500;
  double depth_diff = 0.15;
  double dynamic_speed = 0.3;
  int neighbor_distance = 4;
};
#endif

// pcl::fromROSMsg<pcl::PointXYZVelocity> equivalents (clusterer_nodelet.cpp:226).  velocityCloudLayout: 1 = the
// constructor's own layout (32-byte records, dense rows: usable in place), 0 = another layout that carries the fields
// by name (repack), -1 = malformed.  repackVelocityCloud writes width*height records {x,y,z,0,vx,vy,vz,0}; a field the
// message lacks (or carries with another datatype) stays 0, as in PCL.
int velocityCloudLayout(const sensor_msgs::PointCloud2& msg);
void repackVelocityCloud(const sensor_msgs::PointCloud2& msg, std::vector<uint8_t>& out);

class ClustererB200 {
 public:
  ClustererB200(int max_w, int max_h, int device = 0);
  ~ClustererB200();
  ClustererB200(const ClustererB200&) = delete;
  ClustererB200& operator=(const ClustererB200&) = delete;

  void reconfigureCB(const ClustererConfig& config, uint32_t level);  // clusterer_nodelet.cpp:345-352

  // ClustererNodelet::dataCB (:221-242): each output is computed and published only when subscribed (:233-238), in
  // the reference's order ~clusters, ~clusters_image, ~moving_objects; one mod_cluster call serves all three (the
  // label image is requested only when a visualisation is subscribed).
  // Returns false (and publishes nothing) on a malformed cloud or a GPU error.
  bool dataCB(const sensor_msgs::PointCloud2& input_pc_msg);

  scene_flow_constructor::Sink<visualization_msgs::MarkerArray> clusters_pub_;       // ~clusters        (:269-291)
  scene_flow_constructor::Sink<sensor_msgs::Image> clusters_image_pub_;              // ~clusters_image  BGR8 (:293-322)
  scene_flow_constructor::Sink<moving_object_msgs::MovingObjectArray> dynamic_objects_pub_;  // ~moving_objects

  const std::string& lastError() const { return last_error_; }

 private:
  mod_ctx* ctx_ = nullptr;
  int cluster_size_th_ = 2500;
  double depth_diff_th_ = 0.15;
  double dynamic_speed_th_ = 0.3;
  int neighbor_distance_th_ = 4;
  std::string last_error_;
  std::vector<mod_object> objs_;
  int max_w_ = 0, max_h_ = 0, device_ = 0;
};

}  // namespace scene_flow_clusterer

// ---------------------------------------------------------------------------------------------------------------
// image_crop::ImageCrop (image_crop/include/image_crop.h, src/image_crop.cpp:20-43): the node detect_with_zed.launch
// puts in front of both cameras.  cameraCallback crops on the GPU through mod_image_crop and shifts the principal
// point with mod_host_crop_camera_info; ~cropped_image carries bgr8 pixels and the adjusted CameraInfo.
namespace image_crop {

class ImageCropB200 {
 public:
  ImageCropB200(int max_w, int max_h, int target_width, int target_height, int device = 0);  // ~target_width / ~target_height
  ~ImageCropB200();
  ImageCropB200(const ImageCropB200&) = delete;
  ImageCropB200& operator=(const ImageCropB200&) = delete;

  // false (nothing published, lastError() says why) where the reference's cv::Mat ROI assertion would throw
  bool cameraCallback(const sensor_msgs::Image& image_msg, const sensor_msgs::CameraInfo& camera_info);
  const std::string& lastError() const { return last_error_; }

  std::function<void(const sensor_msgs::Image&, const sensor_msgs::CameraInfo&)> camera_pub_;  // ~cropped_image

 private:
  mod_ctx* ctx_ = nullptr;
  int target_width_, target_height_;
  std::string last_error_;
};

}  // namespace image_crop
