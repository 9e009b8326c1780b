// The clusterer nodelet, B200 edition.  External surface identical to the reference's
// scene_flow_clusterer::ClustererNodelet (scene_flow_clusterer/include/clusterer_nodelet.h:37-39,
// src/clusterer_nodelet.cpp:20-37): pluginlib class scene_flow_clusterer/scene_flow_clusterer, subscribes
// `scene_flow` (sensor_msgs/PointCloud2, queue 10, public namespace), advertises `~moving_objects`
// (moving_object_msgs/MovingObjectArray), `~clusters` (visualization_msgs/MarkerArray) and `~clusters_image`
// (image_transport), each computed only while it has subscribers, and serves the dynamic_reconfigure parameters
// cluster_size / depth_diff / dynamic_speed / neighbor_distance.  All compute is ClustererB200::dataCB -> mod_cluster.
#pragma once
#include <dynamic_reconfigure/server.h>
#include <image_transport/image_transport.h>
#include <nodelet/nodelet.h>
#include <ros/ros.h>
#include <sensor_msgs/PointCloud2.h>

#include <memory>

#include "scene_flow_host.h"

namespace scene_flow_clusterer {

class ClustererNodelet : public nodelet::Nodelet {
 public:
  void onInit() override;

 private:
  void dataCB(const sensor_msgs::PointCloud2ConstPtr& cloud);
  void reconfigureCB(ClustererConfig& config, uint32_t level);
  bool ensureEngine(uint32_t width, uint32_t height);

  ros::Subscriber cloud_sub_;
  ros::Publisher moving_objects_pub_;
  ros::Publisher clusters_pub_;
  std::shared_ptr<image_transport::ImageTransport> image_transport_;
  image_transport::Publisher clusters_image_pub_;
  std::shared_ptr<dynamic_reconfigure::Server<ClustererConfig>> reconfigure_server_;

  ClustererConfig config_;                 // last values from dynamic_reconfigure
  std::unique_ptr<ClustererB200> engine_;  // GPU context, sized by the first cloud (re-created if a larger one arrives)
  uint32_t engine_w_ = 0, engine_h_ = 0;
  int device_ = 0;                         // ~device: CUDA ordinal
};

}  // namespace scene_flow_clusterer
