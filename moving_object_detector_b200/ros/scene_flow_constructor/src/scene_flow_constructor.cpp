#include "scene_flow_constructor_b200_node.h"

#include <cv_bridge/cv_bridge.h>
#include <image_geometry/stereo_camera_model.h>
#include <image_transport/camera_common.h>
#include <sensor_msgs/image_encodings.h>
#include <tf2_geometry_msgs/tf2_geometry_msgs.h>

#include <opencv2/core/core.hpp>

#include <cstring>
#include <stdexcept>

namespace scene_flow_constructor {

namespace {

// ~visual_odometry/*: libviso2's matcher, bucketing and stereo parameters under the names the reference reads
// (scene_flow_constructor/include/odometry_params.h:15-52).  A parameter that is not set keeps libviso2's default.
void loadOdometerParams(const ros::NodeHandle& nh, VisualOdometryStereo::parameters& p) {
  nh.getParam("nms_n", p.match.nms_n);
  nh.getParam("nms_tau", p.match.nms_tau);
  nh.getParam("match_binsize", p.match.match_binsize);
  nh.getParam("match_radius", p.match.match_radius);
  nh.getParam("match_disp_tolerance", p.match.match_disp_tolerance);
  nh.getParam("outlier_disp_tolerance", p.match.outlier_disp_tolerance);
  nh.getParam("outlier_flow_tolerance", p.match.outlier_flow_tolerance);
  nh.getParam("multi_stage", p.match.multi_stage);
  nh.getParam("half_resolution", p.match.half_resolution);
  nh.getParam("refinement", p.match.refinement);
  nh.getParam("max_features", p.bucket.max_features);
  nh.getParam("bucket_width", p.bucket.bucket_width);
  nh.getParam("bucket_height", p.bucket.bucket_height);
  nh.getParam("ransac_iters", p.ransac_iters);
  nh.getParam("inlier_threshold", p.inlier_threshold);
  nh.getParam("reweighting", p.reweighting);
}

}  // namespace

SceneFlowConstructor::SceneFlowConstructor() {
  ros::NodeHandle nh;
  ros::NodeHandle pnh("~");
  pnh.param("device", device_, 0);

  ros::NodeHandle odometry_nh(pnh, "visual_odometry");
  loadOdometerParams(odometry_nh, visual_odometer_params_);
  odometry_nh.param("base_link_frame_id", base_link_frame_id_, std::string("base_link"));
  odometry_nh.param("odom_frame_id", odom_frame_id_, std::string("odom"));
  integrated_pose_.setIdentity();
  tf_listener_ = std::make_shared<tf2_ros::TransformListener>(tf_buffer_);

  sgm_gpu_ = std::make_shared<sgm_gpu::SgmGpu>(pnh);
  image_transport_ = std::make_shared<image_transport::ImageTransport>(pnh);

  // the server calls back at once with the parameter-server / default values, so config_ is valid from here on
  reconfigure_server_ = std::make_shared<dynamic_reconfigure::Server<SceneFlowConstructorConfig>>(pnh);
  reconfigure_server_->setCallback(
      [this](SceneFlowConstructorConfig& config, uint32_t level) { reconfigureCB(config, level); });

  depth_pub_ = pnh.advertise<sensor_msgs::Image>("depth", 1);
  optflow_pub_ = pnh.advertise<sensor_msgs::Image>("optical_flow", 1);
  pc_with_velocity_pub_ = pnh.advertise<sensor_msgs::PointCloud2>("scene_flow", 1);
  static_flow_pub_ = pnh.advertise<sensor_msgs::Image>("synthetic_optical_flow", 1);
  pnh.param("fused_clusterer", fused_clusterer_, false);
  if (fused_clusterer_) {
    moving_objects_pub_ = pnh.advertise<moving_object_msgs::MovingObjectArray>("moving_objects", 1);
    clusterer_reconfigure_server_ =
        std::make_shared<dynamic_reconfigure::Server<scene_flow_clusterer::ClustererConfig>>(ros::NodeHandle(pnh, "clusterer"));
    clusterer_reconfigure_server_->setCallback([this](scene_flow_clusterer::ClustererConfig& config, uint32_t) {
      std::lock_guard<std::mutex> lock(config_mutex_);
      clusterer_config_ = config;
    });
  }

  const std::string left_topic = nh.resolveName("left_image"), right_topic = nh.resolveName("right_image");
  left_image_sub_.subscribe(*image_transport_, left_topic, 1);
  right_image_sub_.subscribe(*image_transport_, right_topic, 1);
  left_caminfo_sub_.subscribe(nh, image_transport::getCameraInfoTopic(left_topic), 1);
  right_caminfo_sub_.subscribe(nh, image_transport::getCameraInfoTopic(right_topic), 1);
  stereo_synchronizer_ = std::make_shared<StereoSynchronizer>(left_image_sub_, right_image_sub_, left_caminfo_sub_,
                                                              right_caminfo_sub_, 1);
  stereo_synchronizer_->registerCallback(&SceneFlowConstructor::stereoCallback, this);
}

SceneFlowConstructor::~SceneFlowConstructor() {
  if (construct_job_.valid()) construct_job_.wait();
}

void SceneFlowConstructor::reconfigureCB(SceneFlowConstructorConfig& config, uint32_t /*level*/) {
  ROS_INFO("Reconfigure request: dynamic_flow_diff %d, max_color_velocity %f", config.dynamic_flow_diff,
           config.max_color_velocity);
  std::lock_guard<std::mutex> lock(config_mutex_);
  config_ = config;  // handed to the engine by value at the start of every construct()
}

bool SceneFlowConstructor::ensureEngine(uint32_t width, uint32_t height) {
  if (gpu_unusable_) return false;
  if (engine_ && width <= engine_w_ && height <= engine_h_) return true;
  try {
    engine_.reset();  // release the old context's device memory first
    engine_.reset(new SceneFlowConstructorB200(static_cast<int>(width), static_cast<int>(height), device_));
  } catch (const std::runtime_error& e) {
    // no CPU fallback: the node stops, loudly, instead of publishing nothing for ever
    ROS_FATAL("scene_flow_constructor: %s", e.what());
    gpu_unusable_ = true;
    ros::shutdown();
    return false;
  }
  engine_w_ = width;
  engine_h_ = height;
  engine_->setCameraInfo(latched_camera_info_);
  engine_->setCameraFrame(camera_frame_id_);
  return true;
}

void SceneFlowConstructor::stereoCallback(const sensor_msgs::ImageConstPtr& left_image,
                                          const sensor_msgs::ImageConstPtr& right_image,
                                          const sensor_msgs::CameraInfoConstPtr& left_camera_info,
                                          const sensor_msgs::CameraInfoConstPtr& right_camera_info) {
  const ros::WallTime start = ros::WallTime::now();

  if (!camera_latched_) {
    // the first frame fixes the camera model and the frame the ~depth image is stamped with (reference :368-375)
    latched_camera_info_ = *left_camera_info;
    camera_frame_id_ = left_image->header.frame_id;
    camera_latched_ = true;
  }

  // disparity and optical flow on their own threads, the odometer on this one; all three finish before the frame
  // moves on (reference :377-387).  Optical flow needs a previous left image, i.e. starts with the second frame.
  std::future<DisparityPtr> disparity_job = std::async(std::launch::async, [&] {
    return estimateDisparity(*left_image, *right_image, *left_camera_info, *right_camera_info);
  });
  std::future<FlowPtr> flow_job;
  if (previous_left_image_)
    flow_job = std::async(std::launch::async, [&] { return estimateOpticalFlow(*previous_left_image_, *left_image); });
  const MotionPtr transform_prev2now = estimateCameraMotion(left_image, right_image, *left_camera_info, *right_camera_info);
  const FlowPtr left_flow = flow_job.valid() ? flow_job.get() : FlowPtr();
  const DisparityPtr disparity_now = disparity_job.get();

  if (construct_job_.valid()) construct_job_.get();  // at most one construct() in flight (reference :389-390)
  if (ensureEngine(left_image->width, left_image->height))
    construct_job_ = std::async(std::launch::async, &SceneFlowConstructor::construct, this, disparity_now,
                                disparity_previous_, left_flow, transform_prev2now);

  ROS_INFO("process time: %f", (ros::WallTime::now() - start).toSec());
  previous_left_image_ = left_image;
  disparity_previous_ = disparity_now;
}

void SceneFlowConstructor::construct(DisparityPtr disparity_now, DisparityPtr disparity_previous, FlowPtr left_flow,
                                     MotionPtr transform_prev2now) {
  {
    std::lock_guard<std::mutex> lock(config_mutex_);
    engine_->reconfigureCB(config_, 0);
    engine_->clusterer_cluster_size = clusterer_config_.cluster_size;
    engine_->clusterer_depth_diff = clusterer_config_.depth_diff;
    engine_->clusterer_dynamic_speed = clusterer_config_.dynamic_speed;
    engine_->clusterer_neighbor_distance = clusterer_config_.neighbor_distance;
  }
  // every output is produced only while somebody listens (reference :99, :114, :141, :144)
  engine_->optflow_pub_ = {[this](const sensor_msgs::Image& m) { optflow_pub_.publish(m); },
                           static_cast<int>(optflow_pub_.getNumSubscribers())};
  engine_->depth_pub_ = {[this](const sensor_msgs::Image& m) { depth_pub_.publish(m); },
                         static_cast<int>(depth_pub_.getNumSubscribers())};
  engine_->pc_with_velocity_pub_ = {[this](const sensor_msgs::PointCloud2& m) { pc_with_velocity_pub_.publish(m); },
                                    static_cast<int>(pc_with_velocity_pub_.getNumSubscribers())};
  engine_->static_flow_pub_ = {[this](const sensor_msgs::Image& m) { static_flow_pub_.publish(m); },
                               static_cast<int>(static_flow_pub_.getNumSubscribers())};
  if (fused_clusterer_)
    engine_->moving_objects_pub_ = {[this](const moving_object_msgs::MovingObjectArray& m) { moving_objects_pub_.publish(m); },
                                    static_cast<int>(moving_objects_pub_.getNumSubscribers())};
  engine_->construct(disparity_now, disparity_previous, left_flow, transform_prev2now);
  if (!engine_->lastError().empty()) ROS_ERROR("scene_flow_constructor: %s", engine_->lastError().c_str());
}

SceneFlowConstructor::DisparityPtr SceneFlowConstructor::estimateDisparity(const sensor_msgs::Image& left_image,
                                                                           const sensor_msgs::Image& right_image,
                                                                           const sensor_msgs::CameraInfo& left_camera_info,
                                                                           const sensor_msgs::CameraInfo& right_camera_info) {
  std::shared_ptr<stereo_msgs::DisparityImage> disparity = std::make_shared<stereo_msgs::DisparityImage>();
  if (sgm_gpu_->computeDisparity(left_image, right_image, left_camera_info, right_camera_info, *disparity)) return disparity;
  ROS_ERROR_STREAM("Disparity estimation failed for the frame stamped " << left_image.header.stamp);
  return DisparityPtr();
}

SceneFlowConstructor::FlowPtr SceneFlowConstructor::estimateOpticalFlow(const sensor_msgs::Image& previous_left_image,
                                                                        const sensor_msgs::Image& left_image) {
  // The flow is wanted as a 32FC2 sensor_msgs/Image (that is what ~optical_flow carries and what the GPU path
  // uploads), so PWC-Net is given a cv::Mat header over the message's own buffer; should it hand back other storage,
  // the rows are copied.  Header = the left image's (reference :281).
  std::shared_ptr<sensor_msgs::Image> msg = std::make_shared<sensor_msgs::Image>();
  msg->header = left_image.header;
  msg->width = left_image.width;
  msg->height = left_image.height;
  msg->encoding = sensor_msgs::image_encodings::TYPE_32FC2;
  msg->is_bigendian = 0;
  msg->step = left_image.width * 8;
  msg->data.resize(static_cast<size_t>(msg->step) * msg->height);
  cv::Mat flow(static_cast<int>(msg->height), static_cast<int>(msg->width), CV_32FC2, msg->data.data(), msg->step);
  if (!pwc_net_.estimateOpticalFlow(previous_left_image, left_image, flow)) {
    ROS_ERROR_STREAM("Optical flow estimation failed between the frames stamped " << previous_left_image.header.stamp
                                                                                   << " and " << left_image.header.stamp);
    return FlowPtr();
  }
  if (flow.data != msg->data.data()) {
    msg->width = static_cast<uint32_t>(flow.cols);
    msg->height = static_cast<uint32_t>(flow.rows);
    msg->step = msg->width * 8;
    msg->data.resize(static_cast<size_t>(msg->step) * msg->height);
    for (int r = 0; r < flow.rows; r++)
      std::memcpy(&msg->data[static_cast<size_t>(r) * msg->step], flow.data + static_cast<size_t>(r) * flow.step, msg->step);
  }
  return msg;
}

SceneFlowConstructor::MotionPtr SceneFlowConstructor::estimateCameraMotion(const sensor_msgs::ImageConstPtr& left_image,
                                                                           const sensor_msgs::ImageConstPtr& right_image,
                                                                           const sensor_msgs::CameraInfo& left_camera_info,
                                                                           const sensor_msgs::CameraInfo& right_camera_info) {
  if (!visual_odometer_) {
    // calibration from the first CameraInfo pair (reference :306-320)
    image_geometry::StereoCameraModel model;
    model.fromCameraInfo(left_camera_info, right_camera_info);
    visual_odometer_params_.base = model.baseline();
    visual_odometer_params_.calib.f = model.left().fx();
    visual_odometer_params_.calib.cu = model.left().cx();
    visual_odometer_params_.calib.cv = model.left().cy();
    visual_odometer_ = std::make_shared<VisualOdometryStereo>(visual_odometer_params_);
  }

  const cv_bridge::CvImageConstPtr left = cv_bridge::toCvCopy(left_image, sensor_msgs::image_encodings::MONO8);
  const cv_bridge::CvImageConstPtr right = cv_bridge::toCvCopy(right_image, sensor_msgs::image_encodings::MONO8);
  ROS_ASSERT(left->image.step[0] == right->image.step[0]);
  ROS_ASSERT(left->image.rows == right->image.rows && left->image.cols == right->image.cols);
  int32_t dims[3] = {left->image.cols, left->image.rows, static_cast<int32_t>(left->image.step[0])};
  if (!visual_odometer_->process(left->image.data, right->image.data, dims)) {
    ROS_ERROR_STREAM("Visual odometry failed for the frame stamped " << left_image->header.stamp);
    return MotionPtr();
  }

  // libviso2 reports the motion of the left camera from the previous frame to this one as a 4x4 matrix
  const Matrix m = visual_odometer_->getMotion();
  const tf2::Transform camera_motion(
      tf2::Matrix3x3(m.val[0][0], m.val[0][1], m.val[0][2], m.val[1][0], m.val[1][1], m.val[1][2], m.val[2][0],
                     m.val[2][1], m.val[2][2]),
      tf2::Vector3(m.val[0][3], m.val[1][3], m.val[2][3]));
  integrateAndBroadcastTF(camera_motion.inverse(), left_image->header.stamp);
  return std::make_shared<geometry_msgs::Transform>(tf2::toMsg(camera_motion));
}

void SceneFlowConstructor::integrateAndBroadcastTF(const tf2::Transform& delta_transform, const ros::Time& timestamp) {
  // odom -> base_link = (base -> camera) * integrated camera pose * (base -> camera)^-1 (reference :322-349)
  integrated_pose_ *= delta_transform;

  tf2::Stamped<tf2::Transform> base_to_sensor;
  std::string why;
  if (tf_buffer_.canTransform(base_link_frame_id_, camera_frame_id_, timestamp, &why)) {
    tf2::fromMsg(tf_buffer_.lookupTransform(base_link_frame_id_, camera_frame_id_, timestamp), base_to_sensor);
  } else {
    ROS_ERROR_THROTTLE(10.0, "No tf from '%s' to '%s' (%s); taking it as identity", base_link_frame_id_.c_str(),
                       camera_frame_id_.c_str(), why.c_str());
    base_to_sensor.setIdentity();
  }
  const tf2::Transform base_pose = base_to_sensor * integrated_pose_ * base_to_sensor.inverse();
  geometry_msgs::TransformStamped msg = tf2::toMsg(tf2::Stamped<tf2::Transform>(base_pose, timestamp, odom_frame_id_));
  msg.child_frame_id = base_link_frame_id_;
  tf_broadcaster_.sendTransform(msg);
}

}  // namespace scene_flow_constructor
