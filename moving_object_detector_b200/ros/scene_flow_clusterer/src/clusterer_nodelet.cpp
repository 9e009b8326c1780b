#include "clusterer_nodelet_b200.h"

#include <pluginlib/class_list_macros.h>

#include <stdexcept>

PLUGINLIB_EXPORT_CLASS(scene_flow_clusterer::ClustererNodelet, nodelet::Nodelet)

namespace scene_flow_clusterer {

void ClustererNodelet::onInit() {
  ros::NodeHandle& nh = getNodeHandle();
  ros::NodeHandle& pnh = getPrivateNodeHandle();
  pnh.param("device", device_, 0);

  // the server calls back at once with the parameter-server / default values, so config_ is valid from here on
  reconfigure_server_ = std::make_shared<dynamic_reconfigure::Server<ClustererConfig>>(pnh);
  reconfigure_server_->setCallback(
      [this](ClustererConfig& config, uint32_t level) { reconfigureCB(config, level); });

  image_transport_ = std::make_shared<image_transport::ImageTransport>(pnh);
  clusters_image_pub_ = image_transport_->advertise("clusters_image", 1);
  moving_objects_pub_ = pnh.advertise<moving_object_msgs::MovingObjectArray>("moving_objects", 1);
  clusters_pub_ = pnh.advertise<visualization_msgs::MarkerArray>("clusters", 1);
  cloud_sub_ = nh.subscribe<sensor_msgs::PointCloud2>("scene_flow", 10, &ClustererNodelet::dataCB, this);
}

void ClustererNodelet::reconfigureCB(ClustererConfig& config, uint32_t level) {
  NODELET_INFO("Reconfigure request: cluster_size %d, depth_diff %f, dynamic_speed %f, neighbor_distance %d",
               config.cluster_size, config.depth_diff, config.dynamic_speed, config.neighbor_distance);
  config_ = config;
  if (engine_) engine_->reconfigureCB(config_, level);
}

bool ClustererNodelet::ensureEngine(uint32_t width, uint32_t height) {
  if (engine_ && width <= engine_w_ && height <= engine_h_) return true;
  try {
    engine_.reset();  // release the old context's device memory first
    engine_.reset(new ClustererB200(static_cast<int>(width), static_cast<int>(height), device_));
  } catch (const std::runtime_error& e) {
    NODELET_ERROR("scene_flow_clusterer: %s", e.what());  // no CPU fallback: the cloud is dropped, loudly
    return false;
  }
  engine_w_ = width;
  engine_h_ = height;
  engine_->reconfigureCB(config_, 0);
  return true;
}

void ClustererNodelet::dataCB(const sensor_msgs::PointCloud2ConstPtr& cloud) {
  if (cloud->width == 0 || cloud->height == 0 || !ensureEngine(cloud->width, cloud->height)) return;
  // every output is produced only while somebody listens; the engine publishes ~clusters, ~clusters_image,
  // ~moving_objects in that order
  engine_->clusters_pub_ = {[this](const visualization_msgs::MarkerArray& m) { clusters_pub_.publish(m); },
                            static_cast<int>(clusters_pub_.getNumSubscribers())};
  engine_->clusters_image_pub_ = {[this](const sensor_msgs::Image& m) { clusters_image_pub_.publish(m); },
                                  static_cast<int>(clusters_image_pub_.getNumSubscribers())};
  engine_->dynamic_objects_pub_ = {[this](const moving_object_msgs::MovingObjectArray& m) { moving_objects_pub_.publish(m); },
                                   static_cast<int>(moving_objects_pub_.getNumSubscribers())};
  if (!engine_->dataCB(*cloud)) NODELET_ERROR("scene_flow_clusterer: %s", engine_->lastError().c_str());
}

}  // namespace scene_flow_clusterer
