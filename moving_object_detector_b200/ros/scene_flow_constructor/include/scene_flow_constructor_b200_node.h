// The scene-flow constructor node, B200 edition.  External surface identical to the reference's
// scene_flow_constructor::SceneFlowConstructor (scene_flow_constructor/include/scene_flow_constructor.h:32-35,
// src/scene_flow_constructor.cpp:22-62): subscribes `left_image` / `right_image` (image_transport) and their
// `camera_info` topics through an exact-time synchroniser of queue 1; advertises `~depth` (32FC1), `~optical_flow`
// (32FC2), `~scene_flow` (sensor_msgs/PointCloud2 of pcl::PointXYZVelocity) and `~synthetic_optical_flow` (32FC2),
// each computed only while it has subscribers; reads `~visual_odometry/*` and serves the dynamic_reconfigure
// parameters dynamic_flow_diff / max_color_velocity; broadcasts the integrated odometry on tf.
//
// Per stereo frame (stereoCallback, reference :364-399) the three upstream estimators - sgm_gpu (disparity), pwc_net
// (optical flow), libviso2 (camera motion) - run concurrently, exactly as before: they are inputs of this path.
// What used to be five CPU passes over the image (construct(), reference :91-147) is one call into libmod_b200.so
// through SceneFlowConstructorB200::construct, on a worker so that the next frame's estimators overlap it; at most
// one construct() is in flight, as in the reference (:389-392).
#pragma once
#include <dynamic_reconfigure/server.h>
#include <geometry_msgs/Transform.h>
#include <image_transport/image_transport.h>
#include <image_transport/subscriber_filter.h>
#include <message_filters/subscriber.h>
#include <message_filters/time_synchronizer.h>
#include <pwc_net/pwc_net.h>
#include <ros/ros.h>
#include <sensor_msgs/CameraInfo.h>
#include <sensor_msgs/Image.h>
#include <sgm_gpu/sgm_gpu.h>
#include <stereo_msgs/DisparityImage.h>
#include <tf2/LinearMath/Transform.h>
#include <tf2_ros/transform_broadcaster.h>
#include <tf2_ros/transform_listener.h>
#include <viso_stereo.h>

#include <future>
#include <memory>
#include <mutex>
#include <string>

#include "scene_flow_host.h"

namespace scene_flow_constructor {

class SceneFlowConstructor {
 public:
  SceneFlowConstructor();
  ~SceneFlowConstructor();  // waits for the construct() in flight

 private:
  // what the upstream estimators hand to construct(); null = that estimator failed on this frame
  typedef std::shared_ptr<const stereo_msgs::DisparityImage> DisparityPtr;
  typedef std::shared_ptr<const sensor_msgs::Image> FlowPtr;               // 32FC2, header of the left image
  typedef std::shared_ptr<const geometry_msgs::Transform> MotionPtr;       // previous camera frame -> current one

  void stereoCallback(const sensor_msgs::ImageConstPtr& left_image, const sensor_msgs::ImageConstPtr& right_image,
                      const sensor_msgs::CameraInfoConstPtr& left_camera_info,
                      const sensor_msgs::CameraInfoConstPtr& right_camera_info);
  void reconfigureCB(SceneFlowConstructorConfig& config, uint32_t level);

  DisparityPtr estimateDisparity(const sensor_msgs::Image& left_image, const sensor_msgs::Image& right_image,
                                 const sensor_msgs::CameraInfo& left_camera_info,
                                 const sensor_msgs::CameraInfo& right_camera_info);
  FlowPtr estimateOpticalFlow(const sensor_msgs::Image& previous_left_image, const sensor_msgs::Image& left_image);
  MotionPtr estimateCameraMotion(const sensor_msgs::ImageConstPtr& left_image, const sensor_msgs::ImageConstPtr& right_image,
                                 const sensor_msgs::CameraInfo& left_camera_info,
                                 const sensor_msgs::CameraInfo& right_camera_info);
  void integrateAndBroadcastTF(const tf2::Transform& delta_transform, const ros::Time& timestamp);

  bool ensureEngine(uint32_t width, uint32_t height);
  void construct(DisparityPtr disparity_now, DisparityPtr disparity_previous, FlowPtr left_flow, MotionPtr transform_prev2now);

  // outputs
  ros::Publisher optflow_pub_, depth_pub_, pc_with_velocity_pub_, static_flow_pub_;
  // ~fused_clusterer (default false): also run the clusterer in this process, on the cloud still on the GPU, and
  // publish ~moving_objects (remap it to where scene_flow_clusterer's ~moving_objects went); its four parameters are
  // served by a second dynamic_reconfigure server under ~clusterer.  The separate clusterer node is then not needed
  // and ~scene_flow's 32 B/px only leave the GPU while somebody subscribes to it.
  bool fused_clusterer_ = false;
  ros::Publisher moving_objects_pub_;
  std::shared_ptr<dynamic_reconfigure::Server<scene_flow_clusterer::ClustererConfig>> clusterer_reconfigure_server_;
  scene_flow_clusterer::ClustererConfig clusterer_config_;  // guarded by config_mutex_

  // inputs: stereo images + camera infos, exact-time synchronised
  std::shared_ptr<image_transport::ImageTransport> image_transport_;
  image_transport::SubscriberFilter left_image_sub_, right_image_sub_;
  message_filters::Subscriber<sensor_msgs::CameraInfo> left_caminfo_sub_, right_caminfo_sub_;
  typedef message_filters::TimeSynchronizer<sensor_msgs::Image, sensor_msgs::Image, sensor_msgs::CameraInfo, sensor_msgs::CameraInfo>
      StereoSynchronizer;
  std::shared_ptr<StereoSynchronizer> stereo_synchronizer_;

  // parameters
  std::shared_ptr<dynamic_reconfigure::Server<SceneFlowConstructorConfig>> reconfigure_server_;
  std::mutex config_mutex_;             // reconfigureCB runs on the spinner thread, construct() on the worker
  SceneFlowConstructorConfig config_;
  int device_ = 0;                      // ~device: CUDA ordinal

  // upstream estimators
  std::shared_ptr<sgm_gpu::SgmGpu> sgm_gpu_;
  pwc_net::PwcNet pwc_net_;
  std::shared_ptr<VisualOdometryStereo> visual_odometer_;  // created from the first pair of CameraInfos
  VisualOdometryStereo::parameters visual_odometer_params_;

  // odometry on tf
  std::string base_link_frame_id_, odom_frame_id_, camera_frame_id_;
  tf2_ros::Buffer tf_buffer_;
  std::shared_ptr<tf2_ros::TransformListener> tf_listener_;
  tf2_ros::TransformBroadcaster tf_broadcaster_;
  tf2::Transform integrated_pose_;

  // frame-to-frame state (reference :397-398)
  sensor_msgs::ImageConstPtr previous_left_image_;
  DisparityPtr disparity_previous_;

  // GPU side
  std::unique_ptr<SceneFlowConstructorB200> engine_;  // sized by the first frame; camera model latched there (:368-375)
  uint32_t engine_w_ = 0, engine_h_ = 0;
  bool gpu_unusable_ = false;
  sensor_msgs::CameraInfo latched_camera_info_;
  bool camera_latched_ = false;
  std::future<void> construct_job_;
};

}  // namespace scene_flow_constructor
