// scene_flow_host.cpp -- see scene_flow_host.h.  Host logic only; every computation is one C-ABI call.
#include "scene_flow_host.h"

#include <cstring>
#include <stdexcept>
#include <vector>

namespace {

std::string ctxError(mod_ctx* ctx, const char* what, int rc) {
  return std::string(what) + " failed (" + std::to_string(rc) + "): " + (ctx ? mod_last_error(ctx) : "no context");
}

mod_ctx* createContext(int max_w, int max_h, int device, int max_clusters = 0) {
  mod_ctx* ctx = nullptr;
  const int rc = mod_create(device, max_w, max_h, 1, max_clusters, &ctx);
  if (rc != MOD_OK)
    throw std::runtime_error("mod_create failed (" + std::to_string(rc) +
                             "): no usable sm_100a GPU; this path has no CPU fallback");
  return ctx;
}

// The reference's MovingObjectArray is unbounded (clusterer_nodelet.cpp:324-343); a context holds max_clusters
// survivors per frame.  When a frame has more (MOD_ERR_CLUSTER_SLOTS) the context is re-created with room for twice as
// many and the caller repeats the frame.  Returns false when the context is already at the library's limit.
bool growContext(mod_ctx*& ctx, int max_w, int max_h, int device) {
  const int have = mod_max_clusters(ctx);
  mod_ctx* bigger = nullptr;
  if (mod_create(device, max_w, max_h, 1, have * 2, &bigger) != MOD_OK) return false;
  if (mod_max_clusters(bigger) <= have) {
    mod_destroy(bigger);
    return false;
  }
  mod_destroy(ctx);
  ctx = bigger;
  return true;
}

// mod_object records -> MovingObjectArray: float -> float64 widening as in the reference's assignments
// (clusterer_nodelet.cpp:151-181), identity orientation, header = the cloud's (:327)
void fillMovingObjects(const mod_object* objs, int32_t n, const std_msgs::Header& header,
                       moving_object_msgs::MovingObjectArray& out) {
  out.header = header;
  out.moving_object_array.resize(static_cast<size_t>(n));
  for (int32_t i = 0; i < n; i++) {
    const mod_object& o = objs[i];
    moving_object_msgs::MovingObject& m = out.moving_object_array[static_cast<size_t>(i)];
    m.id = o.id;
    m.center.position.x = o.center[0];
    m.center.position.y = o.center[1];
    m.center.position.z = o.center[2];
    m.center.orientation.x = 0;
    m.center.orientation.y = 0;
    m.center.orientation.z = 0;
    m.center.orientation.w = 1;
    m.velocity.x = o.velocity[0];
    m.velocity.y = o.velocity[1];
    m.velocity.z = o.velocity[2];
    m.bounding_box.x = o.bbox[0];
    m.bounding_box.y = o.bbox[1];
    m.bounding_box.z = o.bbox[2];
  }
}

mod_disparity toC(const stereo_msgs::DisparityImage& d) {
  mod_disparity c;
  c.f = d.f;
  c.T = d.T;
  c.min_disparity = d.min_disparity;
  c.max_disparity = d.max_disparity;
  c.data = reinterpret_cast<const float*>(d.image.data.data());  // cv::Mat_<float>(h, w, data, step), dip.cpp:5
  c.step_bytes = static_cast<int32_t>(d.image.step);
  return c;
}

void fillImage(sensor_msgs::Image& img, const std_msgs::Header& header, uint32_t w, uint32_t h, const char* encoding,
               uint32_t bytes_per_pixel) {
  img.header = header;
  img.width = w;
  img.height = h;
  img.encoding = encoding;
  img.is_bigendian = 0;
  img.step = w * bytes_per_pixel;
  img.data.resize(static_cast<size_t>(img.step) * h);
}

}  // namespace

namespace scene_flow_constructor {

SceneFlowConstructorB200::SceneFlowConstructorB200(int max_w, int max_h, int device)
    : ctx_(createContext(max_w, max_h, device)), max_w_(max_w), max_h_(max_h), device_(device) {}

SceneFlowConstructorB200::~SceneFlowConstructorB200() { mod_destroy(ctx_); }

void SceneFlowConstructorB200::reconfigureCB(const SceneFlowConstructorConfig& config, uint32_t /*level*/) {
  dynamic_flow_diff_ = config.dynamic_flow_diff;
  max_color_velocity_ = config.max_color_velocity;
}

void SceneFlowConstructorB200::setCameraInfo(const sensor_msgs::CameraInfo& info) {
  // image_geometry::PinholeCameraModel::fromCameraInfo: P is copied, then adjusted for ROI offset and binning
  // (no-ops for binning 0/1 and a zero ROI, which is what image_crop emits: image_crop.cpp:32-38)
  const double bx = info.binning_x == 0 ? 1.0 : static_cast<double>(info.binning_x);
  const double by = info.binning_y == 0 ? 1.0 : static_cast<double>(info.binning_y);
  double fx = info.P[0], fy = info.P[5], cx = info.P[2], cy = info.P[6], Tx = info.P[3], Ty = info.P[7];
  cx -= info.roi.x_offset;
  cy -= info.roi.y_offset;
  if (bx > 1.0) { fx /= bx; cx /= bx; Tx /= bx; }
  if (by > 1.0) { fy /= by; cy /= by; Ty /= by; }
  camera_ = mod_camera{fx, fy, cx, cy, Tx, Ty};
  camera_frame_id_ = info.header.frame_id;
  have_camera_ = true;
}

void describeVelocityCloud(sensor_msgs::PointCloud2& msg, uint32_t width, uint32_t height) {
  static const char* names[6] = {"x", "y", "z", "vx", "vy", "vz"};
  static const uint32_t offsets[6] = {0, 4, 8, 16, 20, 24};
  msg.fields.resize(6);
  for (int i = 0; i < 6; i++) {
    msg.fields[i].name = names[i];
    msg.fields[i].offset = offsets[i];
    msg.fields[i].datatype = sensor_msgs::PointField::FLOAT32;
    msg.fields[i].count = 1;
  }
  msg.width = width;
  msg.height = height;
  msg.is_bigendian = false;
  msg.point_step = 32;
  msg.row_step = 32 * width;
  msg.is_dense = true;  // pcl::PointCloud(w, h, value) leaves is_dense = true although the cloud holds NaN
}

bool SceneFlowConstructorB200::toMono8(const sensor_msgs::Image& image, std::vector<uint8_t>& mono8, int rule) {
  last_error_.clear();
  const int encoding = mod_host_encoding(image.encoding.c_str());
  if (encoding < 0) {
    last_error_ = "toMono8: unsupported encoding " + image.encoding;
    return false;
  }
  if (image.data.size() < static_cast<size_t>(image.step) * image.height) {
    last_error_ = "toMono8: image data shorter than step * height";
    return false;
  }
  mono8.resize(static_cast<size_t>(image.width) * image.height);
  if (mono8.empty()) return true;
  const int rc = mod_image_to_mono8(ctx_, static_cast<int>(image.width), static_cast<int>(image.height), encoding,
                                    image.data.data(), static_cast<int32_t>(image.step), rule, mono8.data());
  if (rc != MOD_OK) {
    last_error_ = std::string("toMono8: ") + mod_last_error(ctx_);
    return false;
  }
  return true;
}

bool SceneFlowConstructorB200::construct(const std::shared_ptr<const stereo_msgs::DisparityImage>& disparity_now,
                                         const std::shared_ptr<const stereo_msgs::DisparityImage>& disparity_previous,
                                         const std::shared_ptr<const sensor_msgs::Image>& left_flow,
                                         const std::shared_ptr<const geometry_msgs::Transform>& transform_prev2now) {
  last_error_.clear();
  if (left_flow && optflow_pub_.getNumSubscribers() > 0) optflow_pub_.publish(*left_flow);  // :99-100

  const bool full = disparity_now && disparity_previous && left_flow && transform_prev2now && have_camera_;
  const bool want_depth = disparity_now && depth_pub_.getNumSubscribers() > 0 && have_camera_;

  if (!full) {
    // ~depth alone (:111-119): the reference publishes it before the left_flow gate (:122-123).  The kernel needs a
    // full pair, so the depth image is produced by pairing the frame with itself under an identity transform; the
    // depth output only depends on disparity_now (disparity_image_processor.cpp:105-121).
    if (!want_depth) return false;
    const uint32_t w = disparity_now->image.width, h = disparity_now->image.height;
    zero_flow_.assign(static_cast<size_t>(w) * h * 2, 0.0f);
    sensor_msgs::Image& depth = depth_msg_;
    std_msgs::Header hd;
    hd.frame_id = camera_frame_id_;
    hd.stamp = disparity_now->header.stamp;  // publishDepthImage(:431-438)
    fillImage(depth, hd, w, h, "32FC1", 4);
    const mod_disparity now = toC(*disparity_now);
    const mod_transform ident{0, 0, 0, 0, 0, 0, 1};
    const mod_sf_params p{dynamic_flow_diff_, max_color_velocity_};
    resident_now_.reset();  // both slots are rewritten with this frame; nothing is promised for the next call
    const int rc = mod_scene_flow(ctx_, static_cast<int>(w), static_cast<int>(h), &camera_, &now, &now,
                                  zero_flow_.data(), static_cast<int32_t>(w * 8), &ident, 1.0, &p, nullptr,
                                  reinterpret_cast<float*>(depth.data.data()), nullptr);
    if (rc != MOD_OK) {
      last_error_ = ctxError(ctx_, "mod_scene_flow(depth)", rc);
      return false;
    }
    depth_pub_.publish(depth);
    return false;
  }

  const uint32_t w = disparity_now->image.width, h = disparity_now->image.height;
  if (disparity_previous->image.width != w || disparity_previous->image.height != h || left_flow->width != w ||
      left_flow->height != h) {
    last_error_ = "frame size mismatch between disparity_now, disparity_previous and left_flow";
    return false;
  }
  const bool want_cloud = pc_with_velocity_pub_.getNumSubscribers() > 0;    // :141
  const bool want_static = static_flow_pub_.getNumSubscribers() > 0;        // :144

  // The output messages live in the object and are refilled every frame: a fresh 32 B/px vector per frame costs more
  // in page faults and zero-filling than the whole GPU path (profiles/r01_single_frame.jsonl).  publish(const M&)
  // serialises before it returns, so the buffers are free again for the next frame.
  sensor_msgs::PointCloud2& cloud = cloud_msg_;
  sensor_msgs::Image& depth = depth_msg_;
  sensor_msgs::Image& static_flow = static_flow_msg_;
  if (want_cloud) {
    describeVelocityCloud(cloud, w, h);
    cloud.header = left_flow->header;  // publishPointcloud(..., left_flow->header), :142
    cloud.data.resize(static_cast<size_t>(w) * h * 32);
  }
  if (want_depth) {
    std_msgs::Header hd;
    hd.frame_id = camera_frame_id_;
    hd.stamp = disparity_now->header.stamp;
    fillImage(depth, hd, w, h, "32FC1", 4);
  }
  if (want_static) fillImage(static_flow, left_flow->header, w, h, "32FC2", 8);

  const mod_disparity now = toC(*disparity_now), prev = toC(*disparity_previous);
  const geometry_msgs::Transform& t = *transform_prev2now;
  const mod_transform tr{t.translation.x, t.translation.y, t.translation.z, t.rotation.x, t.rotation.y, t.rotation.z,
                         t.rotation.w};
  // (stamp_now - stamp_previous).toSec(), :162-164
  const double dt = (disparity_now->header.stamp - disparity_previous->header.stamp).toSec();
  const mod_sf_params p{dynamic_flow_diff_, max_color_velocity_};
  // In a stream the previous disparity is the message the last call uploaded as `now` (the reference's
  // disparity_previous_ = disparity_now_, :398): it is still on the device.  Identity is the message OBJECT, not its
  // stamp (stamps repeat when a bag loops or /clock resets): the last `now` message is kept alive here, so its address
  // cannot be recycled, and every upload gets a fresh id.
  const unsigned long long prev_id =
      (resident_now_ && resident_now_.get() == disparity_previous.get()) ? resident_now_id_ : 0ull;
  const unsigned long long now_id = ++frame_counter_;
  mod_set_frame_ids(ctx_, now_id, prev_id);
  resident_now_.reset();
  const int rc = mod_scene_flow(ctx_, static_cast<int>(w), static_cast<int>(h), &camera_, &now, &prev,
                                reinterpret_cast<const float*>(left_flow->data.data()),
                                static_cast<int32_t>(left_flow->step), &tr, dt, &p,
                                want_cloud ? cloud.data.data() : nullptr,
                                want_depth ? reinterpret_cast<float*>(depth.data.data()) : nullptr,
                                want_static ? reinterpret_cast<float*>(static_flow.data.data()) : nullptr);
  if (rc != MOD_OK) {
    last_error_ = ctxError(ctx_, "mod_scene_flow", rc);
    return false;
  }
  resident_now_ = disparity_now;
  resident_now_id_ = now_id;
  if (want_depth) depth_pub_.publish(depth);
  if (want_cloud) pc_with_velocity_pub_.publish(cloud);
  if (want_static) static_flow_pub_.publish(static_flow);
  if (moving_objects_pub_.getNumSubscribers() > 0) {  // the clusterer, in process, on the cloud still on the GPU
    const mod_cl_params cp{clusterer_cluster_size, clusterer_depth_diff, clusterer_dynamic_speed,
                           clusterer_neighbor_distance};
    objs_.resize(static_cast<size_t>(mod_max_clusters(ctx_)));
    int32_t n = 0;
    int rc2 = mod_cluster_resident(ctx_, static_cast<int>(w), static_cast<int>(h), &cp, nullptr, objs_.data(),
                                   static_cast<int32_t>(objs_.size()), &n);
    while (rc2 == MOD_ERR_CLUSTER_SLOTS && growContext(ctx_, max_w_, max_h_, device_)) {
      // more surviving clusters than the context had room for: the new context recomputes this frame's cloud
      resident_now_.reset();
      objs_.resize(static_cast<size_t>(mod_max_clusters(ctx_)));
      rc2 = mod_scene_flow(ctx_, static_cast<int>(w), static_cast<int>(h), &camera_, &now, &prev,
                           reinterpret_cast<const float*>(left_flow->data.data()), static_cast<int32_t>(left_flow->step),
                           &tr, dt, &p, nullptr, nullptr, nullptr);
      if (rc2 == MOD_OK)
        rc2 = mod_cluster_resident(ctx_, static_cast<int>(w), static_cast<int>(h), &cp, nullptr, objs_.data(),
                                   static_cast<int32_t>(objs_.size()), &n);
    }
    if (rc2 != MOD_OK) {
      last_error_ = ctxError(ctx_, "mod_cluster_resident", rc2);
      return true;  // the scene flow itself was computed and published
    }
    moving_object_msgs::MovingObjectArray out;
    fillMovingObjects(objs_.data(), n, left_flow->header, out);  // the header ~scene_flow carries
    moving_objects_pub_.publish(out);
  }
  return true;
}

}  // namespace scene_flow_constructor

namespace scene_flow_clusterer {

// The six FLOAT32 fields pcl::PointXYZVelocity registers (pcl_point_xyz_velocity.h:27-34) and their offsets in the record
static const char* const kVelocityFieldNames[6] = {"x", "y", "z", "vx", "vy", "vz"};
static const uint32_t kVelocityFieldOffsets[6] = {0, 4, 8, 16, 20, 24};

// pcl's FieldMatches: same name, same datatype, same count (a count of 0 means 1)
static const sensor_msgs::PointField* findVelocityField(const sensor_msgs::PointCloud2& msg, int i) {
  const sensor_msgs::PointField* hit = nullptr;
  for (const auto& f : msg.fields)
    if (f.name == kVelocityFieldNames[i] && f.datatype == sensor_msgs::PointField::FLOAT32 && f.count <= 1) hit = &f;
  return hit;
}

int velocityCloudLayout(const sensor_msgs::PointCloud2& msg) {
  if (msg.data.size() < static_cast<size_t>(msg.row_step) * msg.height ||
      static_cast<size_t>(msg.point_step) * msg.width > msg.row_step)
    return -1;
  bool native = msg.point_step == 32 && msg.row_step == 32u * msg.width;
  for (int i = 0; i < 6; i++) {
    const sensor_msgs::PointField* f = findVelocityField(msg, i);
    if (f != nullptr && static_cast<size_t>(f->offset) + 4 > msg.point_step) return -1;
    native = native && f != nullptr && f->offset == kVelocityFieldOffsets[i];
  }
  return native ? 1 : 0;
}

void repackVelocityCloud(const sensor_msgs::PointCloud2& msg, std::vector<uint8_t>& out) {
  // fromPCLPointCloud2: points are value-initialised (all zero), then every registered field that has a match in the
  // message is copied; a field without a match stays 0 ("Failed to find match for field")
  const size_t n = static_cast<size_t>(msg.width) * msg.height;
  out.assign(n * 32, 0);
  for (int i = 0; i < 6; i++) {
    const sensor_msgs::PointField* f = findVelocityField(msg, i);
    if (f == nullptr) continue;
    for (uint32_t r = 0; r < msg.height; r++) {
      const uint8_t* src = msg.data.data() + static_cast<size_t>(r) * msg.row_step + f->offset;
      uint8_t* dst = out.data() + static_cast<size_t>(r) * msg.width * 32 + kVelocityFieldOffsets[i];
      for (uint32_t c = 0; c < msg.width; c++) std::memcpy(dst + static_cast<size_t>(c) * 32, src + static_cast<size_t>(c) * msg.point_step, 4);
    }
  }
}

ClustererB200::ClustererB200(int max_w, int max_h, int device)
    : ctx_(createContext(max_w, max_h, device)), max_w_(max_w), max_h_(max_h), device_(device) {}
ClustererB200::~ClustererB200() { mod_destroy(ctx_); }

void ClustererB200::reconfigureCB(const ClustererConfig& config, uint32_t /*level*/) {
  cluster_size_th_ = config.cluster_size;
  depth_diff_th_ = config.depth_diff;
  dynamic_speed_th_ = config.dynamic_speed;
  neighbor_distance_th_ = config.neighbor_distance;
}

bool ClustererB200::dataCB(const sensor_msgs::PointCloud2& msg) {
  last_error_.clear();
  const bool want_markers = clusters_pub_.getNumSubscribers() > 0;
  const bool want_image = clusters_image_pub_.getNumSubscribers() > 0;
  const bool want_objects = dynamic_objects_pub_.getNumSubscribers() > 0;
  if (!want_markers && !want_image && !want_objects) return true;  // nothing is subscribed
  // pcl::fromROSMsg (clusterer_nodelet.cpp:226) maps fields by name: the constructor's own cloud (32-byte records,
  // x@0 y@4 z@8 vx@16 vy@20 vz@24, dense rows) is handed to the GPU as it is, any other layout is repacked first
  std::vector<uint8_t> repacked;
  const int layout = velocityCloudLayout(msg);
  if (layout < 0) {
    last_error_ = "scene_flow cloud: data is shorter than row_step * height, or a field lies outside point_step";
    return false;
  }
  if (layout == 0) repackVelocityCloud(msg, repacked);
  const uint8_t* cloud = layout == 1 ? msg.data.data() : repacked.data();
  const mod_cl_params p{cluster_size_th_, depth_diff_th_, dynamic_speed_th_, neighbor_distance_th_};
  std::vector<mod_object>& objs = objs_;
  int32_t n = 0;
  const size_t n_points = static_cast<size_t>(msg.width) * msg.height;
  std::vector<int32_t> labels;  // cluster_map_ after removeSmallClusters; only the visualisations need it on the host
  if (want_markers || want_image) labels.resize(n_points);
  int rc;
  do {  // the reference's object array is unbounded: a frame with more survivors than the context holds is repeated
        // on a context with room for twice as many
    objs.resize(static_cast<size_t>(mod_max_clusters(ctx_)));
    rc = mod_cluster(ctx_, static_cast<int>(msg.width), static_cast<int>(msg.height), cloud, &p,
                     labels.empty() ? nullptr : labels.data(), objs.data(), static_cast<int32_t>(objs.size()), &n);
  } while (rc == MOD_ERR_CLUSTER_SLOTS && growContext(ctx_, max_w_, max_h_, device_));
  if (rc != MOD_OK) {
    last_error_ = ctxError(ctx_, "mod_cluster", rc);
    return false;
  }
  // number_of_clusters_ after removeSmallClusters (clusterer_nodelet.cpp:354-393) sizes the colour set of both
  // visualisations (:269-322); n counts the MovingObjects accepted at :176
  mod_frame_stats fs{};
  const int n_clusters = mod_last_frame_stats(ctx_, &fs) == MOD_OK ? fs.n_clusters : n;
  if (want_markers) {  // publishClusters :269-291, cluster2Marker :119-145
    uint8_t table[768];
    mod_host_hsv_colormap(table);
    visualization_msgs::MarkerArray clusters_msg;
    visualization_msgs::Marker delete_marker;
    delete_marker.action = visualization_msgs::Marker::DELETEALL;
    clusters_msg.markers.push_back(delete_marker);
    clusters_msg.markers.resize(static_cast<size_t>(n_clusters) + 1);
    for (int32_t id = 0; id < n_clusters; id++) {
      visualization_msgs::Marker& m = clusters_msg.markers[static_cast<size_t>(id) + 1];
      const uint8_t* bgr = table + 3 * (id * 255 / n_clusters);  // ColorSet::resize(clusters.size()), color_set.cpp:21-25
      m.header = msg.header;
      m.id = id;
      m.type = visualization_msgs::Marker::POINTS;
      m.action = visualization_msgs::Marker::ADD;
      m.scale.x = 0.05;
      m.scale.y = 0.05;
      m.scale.z = 0;
      m.color.a = 1.0;
      m.color.r = bgr[2] / 255.0;
      m.color.g = bgr[1] / 255.0;
      m.color.b = bgr[0] / 255.0;
    }
    // clusterMap2IndicesCluster :97-117: u outer / v inner, so that is the order of every marker's points
    const float* pts = reinterpret_cast<const float*>(cloud);
    for (uint32_t u = 0; u < msg.width; u++) {
      for (uint32_t v = 0; v < msg.height; v++) {
        const size_t i = static_cast<size_t>(msg.width) * v + u;
        const int32_t c = labels[i];
        if (c < 0 || c >= n_clusters) continue;
        geometry_msgs::Point pt;
        pt.x = pts[8 * i];
        pt.y = pts[8 * i + 1];
        pt.z = pts[8 * i + 2];
        clusters_msg.markers[static_cast<size_t>(c) + 1].points.push_back(pt);
      }
    }
    clusters_pub_.publish(clusters_msg);
  }
  if (want_image) {  // publishClustersImage :293-322
    sensor_msgs::Image image;
    image.header.frame_id = msg.header.frame_id;
    image.header.stamp = msg.header.stamp;
    image.height = msg.height;
    image.width = msg.width;
    image.encoding = "bgr8";
    image.is_bigendian = 0;
    image.step = msg.width * 3;
    image.data.resize(n_points * 3);
    const int rc2 = mod_clusters_image(ctx_, static_cast<int>(msg.width), static_cast<int>(msg.height), labels.data(),
                                       n_clusters, image.data.data());
    if (rc2 != MOD_OK) {
      last_error_ = ctxError(ctx_, "mod_clusters_image", rc2);
      return false;
    }
    clusters_image_pub_.publish(image);
  }
  if (!want_objects) return true;
  moving_object_msgs::MovingObjectArray out;
  fillMovingObjects(objs.data(), n, msg.header, out);  // moving_objects_msg.header = input_header_, :327
  dynamic_objects_pub_.publish(out);
  return true;
}

}  // namespace scene_flow_clusterer

namespace image_crop {

ImageCropB200::ImageCropB200(int max_w, int max_h, int target_width, int target_height, int device)
    : ctx_(createContext(max_w, max_h, device)), target_width_(target_width), target_height_(target_height) {}

ImageCropB200::~ImageCropB200() { mod_destroy(ctx_); }

bool ImageCropB200::cameraCallback(const sensor_msgs::Image& image_msg, const sensor_msgs::CameraInfo& camera_info) {
  last_error_.clear();
  // cv_bridge::toCvShare(image_msg, "bgr8") (:23): the 8-bit colour encodings convert by channel shuffles
  const int encoding = mod_host_encoding(image_msg.encoding.c_str());
  if (encoding < 0) {
    last_error_ = "image_crop: unsupported encoding " + image_msg.encoding;
    return false;
  }
  if (image_msg.data.size() < static_cast<size_t>(image_msg.step) * image_msg.height) {
    last_error_ = "image_crop: image data shorter than step * height";
    return false;
  }
  sensor_msgs::CameraInfo cropped_caminfo = camera_info;  // :34
  if (mod_host_crop_camera_info(image_msg.width, image_msg.height, target_width_, target_height_, camera_info.K.data(),
                                camera_info.P.data(), cropped_caminfo.K.data(), cropped_caminfo.P.data()) != MOD_OK) {
    last_error_ = "image_crop: the target window does not lie inside the image";
    return false;
  }
  cropped_caminfo.width = static_cast<uint32_t>(target_width_);   // :39-40
  cropped_caminfo.height = static_cast<uint32_t>(target_height_);
  sensor_msgs::Image cropped;  // CvImage::toImageMsg of the bgr8 copy: dense rows
  cropped.header = image_msg.header;
  cropped.width = static_cast<uint32_t>(target_width_);
  cropped.height = static_cast<uint32_t>(target_height_);
  cropped.encoding = "bgr8";
  cropped.is_bigendian = 0;
  cropped.step = 3u * cropped.width;
  cropped.data.resize(static_cast<size_t>(cropped.step) * cropped.height);
  const int rc = mod_image_crop(ctx_, static_cast<int>(image_msg.width), static_cast<int>(image_msg.height), encoding,
                                image_msg.data.data(), static_cast<int32_t>(image_msg.step), target_width_, target_height_,
                                cropped.data.data());
  if (rc != MOD_OK) {
    last_error_ = ctxError(ctx_, "mod_image_crop", rc);
    return false;
  }
  if (camera_pub_) camera_pub_(cropped, cropped_caminfo);  // :42
  return true;
}

}  // namespace image_crop
