// ref_full_wrap.cpp -- C entry points that drive the REFERENCE's own hot-path code.
//
// TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs).
//
// oracle/Makefile compiles this file together with the reference's unmodified translation units
//   disparity_image_proc/src/disparity_image_processor.cpp, scene_flow_constructor/src/scene_flow_constructor.cpp,
//   scene_flow_clusterer/src/{clusterer_nodelet,color_set,lookup_table}.cpp
// from where they lie under /root/reference, against the stand-in dependency headers of oracle/ref_shim/ref_shim.h,
// into oracle/_ref/libmod_ref.so.  No reference source is copied into this repository.  The wrapper only does what a
// ROS graph would do around the two nodes: build the input messages, call SceneFlowConstructor::construct() /
// ClustererNodelet::dataCB() (private members, reached with -fno-access-control), and read back what the nodes
// published.  Struct layouts equal oracle/mod_oracle.h so tests hand the same ctypes buffers to both.
#include <cstdint>
#include <cstring>
#include <thread>
#include <chrono>
#include <atomic>

#include "scene_flow_constructor.h"   // reference: scene_flow_constructor/include/scene_flow_constructor.h
#include "clusterer_nodelet.h"        // reference: scene_flow_clusterer/include/clusterer_nodelet.h
#include "image_crop.h"               // reference: image_crop/include/image_crop.h
#include <sensor_msgs/PointCloud2.h>
#include <visualization_msgs/MarkerArray.h>

#include "msg_describe.h"
#include "node_sequence.h"

namespace cv {
const uchar*& mod_ref_colormap_table() {
  static const uchar* table = nullptr;
  return table;
}
}  // namespace cv

namespace {

struct ref_camera { double fx, fy, cx, cy, Tx, Ty; };
struct ref_disparity { float f, T, min_disparity, max_disparity; const float* data; int32_t step_bytes; };
struct ref_transform { double tx, ty, tz, qx, qy, qz, qw; };

uint8_t g_colormap[768];

sensor_msgs::CameraInfo make_camera_info(int w, int h, const ref_camera& c) {
  sensor_msgs::CameraInfo info;
  info.width = w; info.height = h;
  // P = [fx 0 cx Tx; 0 fy cy Ty; 0 0 1 0]
  info.P[0] = c.fx; info.P[2] = c.cx; info.P[3] = c.Tx;
  info.P[5] = c.fy; info.P[6] = c.cy; info.P[7] = c.Ty;
  info.P[10] = 1.0;
  return info;
}

stereo_msgs::DisparityImage make_disparity_msg(int w, int h, const ref_disparity& d, uint32_t sec, uint32_t nsec) {
  stereo_msgs::DisparityImage m;
  m.header.stamp = ros::Time(sec, nsec);
  m.header.frame_id = "disparity_frame";
  m.image.width = w; m.image.height = h; m.image.encoding = "32FC1";
  m.image.step = (uint32_t)d.step_bytes;
  m.image.data.resize((size_t)d.step_bytes * h);
  for (int r = 0; r < h; r++)
    std::memcpy(&m.image.data[(size_t)r * d.step_bytes], (const char*)d.data + (size_t)r * d.step_bytes, (size_t)w * 4);
  m.f = d.f; m.T = d.T; m.min_disparity = d.min_disparity; m.max_disparity = d.max_disparity;
  return m;
}

// What the constructor node needs before construct() can run: the state stereoCallback sets on the first frame
// (scene_flow_constructor.cpp:368-375) and the dynamic_reconfigure value (reconfigureCB, :401-407).
void prime_constructor(scene_flow_constructor::SceneFlowConstructor& node, int w, int h, const sensor_msgs::CameraInfo& info,
                       int32_t dynamic_flow_diff, int subscribers) {
  node.left_cam_model_.reset(new image_geometry::PinholeCameraModel());
  node.left_cam_model_->fromCameraInfo(info);
  node.image_width_ = w;
  node.image_height_ = h;
  node.camera_frame_id_ = "left_camera_optical";  // left_image->header.frame_id (:373)
  scene_flow_constructor::SceneFlowConstructorConfig cfg;
  cfg.dynamic_flow_diff = dynamic_flow_diff;
  node.reconfigureCB(cfg, 0);
  node.pc_with_velocity_pub_.state->num_subscribers = (subscribers & 1) ? 1 : 0;
  node.depth_pub_.state->num_subscribers = (subscribers & 2) ? 1 : 0;
  node.static_flow_pub_.state->num_subscribers = (subscribers & 4) ? 1 : 0;
  node.optflow_pub_.state->num_subscribers = (subscribers & 8) ? 1 : 0;
}

std::shared_ptr<cv_bridge::CvImage> make_flow(int w, int h, const float* flow_xy, int32_t flow_step_bytes,
                                              uint32_t sec = 0, uint32_t nsec = 0) {
  std_msgs::Header hd;  // estimateOpticalFlow stamps the flow with the left image's header (:281)
  hd.seq = 42; hd.stamp = ros::Time(sec + 1, nsec); hd.frame_id = "flow_frame";  // distinct from the disparity header on purpose
  std::shared_ptr<cv_bridge::CvImage> flow(new cv_bridge::CvImage(hd, sensor_msgs::image_encodings::TYPE_32FC2));
  flow->image = cv::Mat(h, w, CV_32FC2, (void*)flow_xy, (size_t)flow_step_bytes);  // pwc_net fills a cv::Mat of this type
  return flow;
}

geometry_msgs::TransformPtr make_transform(const ref_transform& t) {
  geometry_msgs::TransformPtr m(new geometry_msgs::Transform());
  m->translation.x = t.tx; m->translation.y = t.ty; m->translation.z = t.tz;
  m->rotation.x = t.qx; m->rotation.y = t.qy; m->rotation.z = t.qz; m->rotation.w = t.qw;
  return m;
}

sensor_msgs::PointCloud2Ptr make_cloud_msg(int w, int h, const void* points32) {
  sensor_msgs::PointCloud2Ptr m(new sensor_msgs::PointCloud2());
  m->width = w; m->height = h; m->point_step = 32; m->row_step = 32u * (uint32_t)w; m->is_dense = 1;
  const char* names[6] = {"x", "y", "z", "vx", "vy", "vz"};
  const uint32_t offs[6] = {0, 4, 8, 16, 20, 24};
  for (int i = 0; i < 6; i++) {
    sensor_msgs::PointField f; f.name = names[i]; f.offset = offs[i]; f.datatype = sensor_msgs::PointField::FLOAT32; f.count = 1;
    m->fields.push_back(f);
  }
  m->data.resize((size_t)w * h * 32);
  std::memcpy(m->data.data(), points32, m->data.size());
  return m;
}

void prime_clusterer(scene_flow_clusterer::ClustererNodelet& node, int32_t cluster_size, double depth_diff, double dynamic_speed,
                     int32_t neighbor_distance, int subscribers) {
  node.onInit();
  scene_flow_clusterer::ClustererConfig cfg;
  cfg.cluster_size = cluster_size; cfg.depth_diff = depth_diff; cfg.dynamic_speed = dynamic_speed;
  cfg.neighbor_distance = neighbor_distance;
  node.reconfigureCB(cfg, 0);
  node.dynamic_objects_pub_.state->num_subscribers = (subscribers & 1) ? 1 : 0;
  node.clusters_pub_.state->num_subscribers = (subscribers & 2) ? 1 : 0;
  node.clusters_image_pub_.state->num_subscribers = (subscribers & 4) ? 1 : 0;
}

// the last ~scene_flow message, so that a test can feed the clusterer the very message the constructor published
sensor_msgs::PointCloud2Ptr g_last_cloud;
// metadata of everything the last ref_scene_flow call published (oracle/msg_describe.h)
std::string g_last_description;

}  // namespace

extern "C" {

int ref_set_colormap(const uint8_t* bgr768) {
  std::memcpy(g_colormap, bgr768, 768);
  cv::mod_ref_colormap_table() = g_colormap;
  return 0;
}

/* defaults the nodes latch from dynamic_reconfigure at start-up (Server::setCallback -> reconfigureCB):
 * out5 = dynamic_flow_diff, cluster_size, neighbor_distance ; out2d = depth_diff, dynamic_speed */
int ref_defaults(int32_t* out3, double* out2d) {
  if (!cv::mod_ref_colormap_table()) return -20;
  scene_flow_constructor::SceneFlowConstructor c;
  scene_flow_clusterer::ClustererNodelet n;
  n.onInit();
  out3[0] = c.dynamic_flow_diff_; out3[1] = n.cluster_size_th_; out3[2] = n.neighbor_distance_th_;
  out2d[0] = n.depth_diff_th_; out2d[1] = n.dynamic_speed_th_;
  return 0;
}

/*
 * SceneFlowConstructor::construct (scene_flow_constructor.cpp:91-147) on one frame pair.
 * now / prev / flow_xy / prev2now may be NULL (the reference's degrade rules).  subscribers: bit0 ~scene_flow,
 * bit1 ~depth, bit2 ~synthetic_optical_flow, bit3 ~optical_flow.  Outputs may be NULL.
 * *published gets the same bits for what was actually published.  cloud_meta (20 ints): width, height, point_step,
 * row_step, is_dense, is_bigendian, n_fields, names_ok, then (offset, datatype) of the six fields.
 * Returns 0, -10 if the reference threw std::out_of_range, -11 on another exception.
 */
int ref_scene_flow(int w, int h, const ref_camera* cam, const ref_disparity* now, const ref_disparity* prev,
                   const float* flow_xy, int32_t flow_step_bytes, const ref_transform* prev2now,
                   uint32_t prev_sec, uint32_t prev_nsec, uint32_t now_sec, uint32_t now_nsec, int32_t dynamic_flow_diff,
                   int subscribers, void* out_points32, float* out_depth, float* out_static_flow, float* out_flow_echo,
                   int32_t* published, int32_t* cloud_meta) {
  try {
    const sensor_msgs::CameraInfo info = make_camera_info(w, h, *cam);
    scene_flow_constructor::SceneFlowConstructor node;
    prime_constructor(node, w, h, info, dynamic_flow_diff, subscribers);
    std::shared_ptr<DisparityImageProcessor> dnow, dprev;
    if (now) dnow.reset(new DisparityImageProcessor(make_disparity_msg(w, h, *now, now_sec, now_nsec), info));
    if (prev) dprev.reset(new DisparityImageProcessor(make_disparity_msg(w, h, *prev, prev_sec, prev_nsec), info));
    std::shared_ptr<cv_bridge::CvImage> flow;
    if (flow_xy) flow = make_flow(w, h, flow_xy, flow_step_bytes, now_sec, now_nsec);
    geometry_msgs::TransformPtr tf;
    if (prev2now) tf = make_transform(*prev2now);

    node.construct(dnow, dprev, flow, tf);

    int pub = 0;
    std::vector<std::pair<long, std::string>> described;  // (publication order, description)
    if (node.pc_with_velocity_pub_.state->n_published) {
      pub |= 1;
      const sensor_msgs::PointCloud2& m = *std::any_cast<sensor_msgs::PointCloud2>(&node.pc_with_velocity_pub_.state->last);
      described.emplace_back(node.pc_with_velocity_pub_.state->order, mod_describe::cloud("scene_flow", m));
      g_last_cloud.reset(new sensor_msgs::PointCloud2(m));
      if (out_points32) std::memcpy(out_points32, m.data.data(), std::min(m.data.size(), (size_t)w * h * 32));
      if (cloud_meta) {
        cloud_meta[0] = m.width; cloud_meta[1] = m.height; cloud_meta[2] = m.point_step; cloud_meta[3] = m.row_step;
        cloud_meta[4] = m.is_dense; cloud_meta[5] = m.is_bigendian; cloud_meta[6] = (int32_t)m.fields.size();
        const char* names[6] = {"x", "y", "z", "vx", "vy", "vz"};
        int ok = m.fields.size() == 6;
        for (size_t i = 0; i < m.fields.size() && i < 6; i++) {
          ok = ok && m.fields[i].name == names[i] && m.fields[i].count == 1;
          cloud_meta[8 + 2 * i] = m.fields[i].offset; cloud_meta[9 + 2 * i] = m.fields[i].datatype;
        }
        cloud_meta[7] = ok;
      }
    }
    if (node.depth_pub_.state->n_published) {
      pub |= 2;
      const sensor_msgs::Image& m = *std::any_cast<sensor_msgs::Image>(&node.depth_pub_.state->last);
      described.emplace_back(node.depth_pub_.state->order, mod_describe::image("depth", m));
      if (out_depth) std::memcpy(out_depth, m.data.data(), std::min(m.data.size(), (size_t)w * h * 4));
    }
    if (node.static_flow_pub_.state->n_published) {
      pub |= 4;
      const sensor_msgs::Image& m = *std::any_cast<sensor_msgs::Image>(&node.static_flow_pub_.state->last);
      described.emplace_back(node.static_flow_pub_.state->order, mod_describe::image("synthetic_optical_flow", m));
      if (out_static_flow) std::memcpy(out_static_flow, m.data.data(), std::min(m.data.size(), (size_t)w * h * 8));
    }
    if (node.optflow_pub_.state->n_published) {
      pub |= 8;
      const sensor_msgs::Image& m = *std::any_cast<sensor_msgs::Image>(&node.optflow_pub_.state->last);
      described.emplace_back(node.optflow_pub_.state->order, mod_describe::image("optical_flow", m));
      if (out_flow_echo) std::memcpy(out_flow_echo, m.data.data(), std::min(m.data.size(), (size_t)w * h * 8));
    }
    std::sort(described.begin(), described.end());
    g_last_description.clear();
    for (const auto& d : described) g_last_description += d.second;
    if (published) *published = pub;
    return 0;
  } catch (const std::out_of_range&) {
    return -10;
  } catch (...) {
    return -11;
  }
}

/* metadata (header, layout, encoding; not the payload) of the messages the last ref_scene_flow call published */
int ref_last_description(char* buf, int cap) {
  if (!buf || cap <= 0) return -1;
  std::snprintf(buf, (size_t)cap, "%s", g_last_description.c_str());
  return (int)g_last_description.size();
}

/*
 * ClustererNodelet::dataCB (clusterer_nodelet.cpp:221-242) on one PointCloud2.  points32 = w*h records of 32 bytes
 * {x,y,z,_,vx,vy,vz,_}, or NULL to use the message the last ref_scene_flow call published.  subscribers: bit0
 * ~moving_objects, bit1 ~clusters, bit2 ~clusters_image.
 * out_objs: max_objs x 14 doubles {id, center.position xyz, center.orientation xyzw, velocity xyz, bounding_box xyz}.
 * out_labels = cluster_map_ after clustering; *number_of_clusters = number_of_clusters_; out_dynamic = dynamic_map_.
 * Markers (bit1): marker_offsets[n_markers+1], marker_points (xyz doubles per point), marker_rgba (4 floats per
 * marker); *n_markers excludes the leading DELETEALL marker, whose presence is reported in *published bit 3.
 */
int ref_cluster(int w, int h, const void* points32, int32_t cluster_size, double depth_diff, double dynamic_speed,
                int32_t neighbor_distance, int subscribers, uint8_t* out_dynamic, int32_t* out_labels,
                int32_t* number_of_clusters, double* out_objs, int32_t max_objs, int32_t* n_objs, uint8_t* out_bgr8,
                int32_t* marker_offsets, double* marker_points, float* marker_rgba, int32_t max_markers,
                int32_t* n_markers, int32_t* published) {
  if (!cv::mod_ref_colormap_table()) return -20;
  try {
    sensor_msgs::PointCloud2ConstPtr msg;
    if (points32) msg = make_cloud_msg(w, h, points32);
    else msg = g_last_cloud;
    if (!msg) return -21;
    scene_flow_clusterer::ClustererNodelet node;
    prime_clusterer(node, cluster_size, depth_diff, dynamic_speed, neighbor_distance, subscribers);

    node.dataCB(msg);

    const size_t n = (size_t)msg->width * msg->height;
    if (out_dynamic) for (size_t i = 0; i < n; i++) out_dynamic[i] = node.dynamic_map_.at(i) ? 1 : 0;
    if (out_labels) std::memcpy(out_labels, node.cluster_map_.data(), n * 4);
    if (number_of_clusters) *number_of_clusters = node.number_of_clusters_;
    int pub = 0, rc = 0;
    if (node.dynamic_objects_pub_.state->n_published) {
      pub |= 1;
      const moving_object_msgs::MovingObjectArray& m =
          *std::any_cast<moving_object_msgs::MovingObjectArray>(&node.dynamic_objects_pub_.state->last);
      const int cnt = (int)m.moving_object_array.size();
      if (n_objs) *n_objs = cnt;
      if (cnt > max_objs) rc = -2;
      for (int i = 0; i < cnt && i < max_objs && out_objs; i++) {
        const moving_object_msgs::MovingObject& o = m.moving_object_array[i];
        double* d = out_objs + 14 * (size_t)i;
        d[0] = o.id;
        d[1] = o.center.position.x; d[2] = o.center.position.y; d[3] = o.center.position.z;
        d[4] = o.center.orientation.x; d[5] = o.center.orientation.y; d[6] = o.center.orientation.z; d[7] = o.center.orientation.w;
        d[8] = o.velocity.x; d[9] = o.velocity.y; d[10] = o.velocity.z;
        d[11] = o.bounding_box.x; d[12] = o.bounding_box.y; d[13] = o.bounding_box.z;
      }
    } else if (n_objs) *n_objs = 0;
    if (node.clusters_pub_.state->n_published) {
      pub |= 2;
      const visualization_msgs::MarkerArray& m = *std::any_cast<visualization_msgs::MarkerArray>(&node.clusters_pub_.state->last);
      if (!m.markers.empty() && m.markers[0].action == visualization_msgs::Marker::DELETEALL) pub |= 8;
      const int cnt = (int)m.markers.size() - 1;
      if (n_markers) *n_markers = cnt;
      if (cnt > max_markers) rc = -3;
      size_t off = 0;
      for (int i = 0; i < cnt && i < max_markers && marker_offsets; i++) {
        const visualization_msgs::Marker& k = m.markers[i + 1];
        if (k.id != i || k.type != visualization_msgs::Marker::POINTS || k.action != visualization_msgs::Marker::ADD) rc = -4;
        marker_offsets[i] = (int32_t)off;
        for (const geometry_msgs::Point& p : k.points) {
          if (marker_points) { marker_points[3 * off] = p.x; marker_points[3 * off + 1] = p.y; marker_points[3 * off + 2] = p.z; }
          off++;
        }
        marker_offsets[i + 1] = (int32_t)off;
        if (marker_rgba) { marker_rgba[4 * i] = k.color.r; marker_rgba[4 * i + 1] = k.color.g; marker_rgba[4 * i + 2] = k.color.b; marker_rgba[4 * i + 3] = k.color.a; }
      }
    } else if (n_markers) *n_markers = 0;
    if (node.clusters_image_pub_.state->n_published) {
      pub |= 4;
      const sensor_msgs::Image& m = *std::any_cast<sensor_msgs::Image>(&node.clusters_image_pub_.state->last);
      if (out_bgr8) std::memcpy(out_bgr8, m.data.data(), std::min(m.data.size(), n * 3));
    }
    if (published) *published = pub;
    return rc;
  } catch (const std::out_of_range&) {
    return -10;
  } catch (...) {
    return -11;
  }
}

/*
 * image_crop::ImageCrop::cameraCallback (image_crop/src/image_crop.cpp:20-43) on one image + CameraInfo.
 * encoding: 0 bgr8, 1 rgb8, 2 bgra8, 3 rgba8, 4 mono8.  K9 / P12: the CameraInfo matrices; out_K9 / out_P12 / out_wh
 * receive the published CameraInfo's K, P, width, height; out_bgr8 the published image (dense target_w*target_h*3).
 * Returns 0, -12 if the reference threw (ROI outside the image), -11 on another failure.
 */
int ref_image_crop(int w, int h, int encoding, const uint8_t* data, int32_t step_bytes, const double* K9, const double* P12,
                   int target_w, int target_h, uint8_t* out_bgr8, double* out_K9, double* out_P12, int32_t* out_wh) {
  static const char* enc_names[5] = {"bgr8", "rgb8", "bgra8", "rgba8", "mono8"};
  if (encoding < 0 || encoding > 4) return -11;
  try {
    ros::mod_ref_int_params() = {{"target_width", target_w}, {"target_height", target_h}};
    image_crop::ImageCrop node;
    sensor_msgs::ImagePtr img(new sensor_msgs::Image());
    img->width = w; img->height = h; img->encoding = enc_names[encoding]; img->step = (uint32_t)step_bytes;
    img->data.assign(data, data + (size_t)step_bytes * h);
    sensor_msgs::CameraInfoPtr info(new sensor_msgs::CameraInfo());
    info->width = w; info->height = h;
    for (int i = 0; i < 9; i++) info->K[i] = K9[i];
    for (int i = 0; i < 12; i++) info->P[i] = P12[i];
    node.cameraCallback(img, info);
    const image_transport::CameraPubState& st = *node.camera_pub_.state;
    if (st.n_published != 1 || st.image.encoding != "bgr8") return -11;
    if (out_bgr8) std::memcpy(out_bgr8, st.image.data.data(), std::min(st.image.data.size(), (size_t)target_w * target_h * 3));
    for (int i = 0; i < 9 && out_K9; i++) out_K9[i] = st.info.K[i];
    for (int i = 0; i < 12 && out_P12; i++) out_P12[i] = st.info.P[i];
    if (out_wh) { out_wh[0] = st.info.width; out_wh[1] = st.info.height; out_wh[2] = st.image.width; out_wh[3] = st.image.height; out_wh[4] = st.image.step; }
    return 0;
  } catch (const std::runtime_error&) {
    return -12;
  } catch (...) {
    return -11;
  }
}

/*
 * CPU baseline harness over the reference's own code: n_pairs frame pairs, n_threads std::threads (one node pair per
 * thread, pairs handed out by an atomic counter), best of `reps`.  Per pair: construct() with a ~scene_flow subscriber,
 * then dataCB() on the published PointCloud2 with a ~moving_objects subscriber -- the work the two ROS nodes do per
 * frame, without TCPROS.  Same argument meaning as orc_bench_pipeline (oracle/mod_oracle.h).
 */
int ref_bench_pipeline(int w, int h, const ref_camera* cam, float f, float T, float min_disp, float max_disp,
                       const float* disp_frames, const int32_t* prev_index, const int32_t* now_index,
                       const float* flows, const ref_transform* transforms, const double* dt, int n_pairs,
                       int32_t dynamic_flow_diff, int32_t cluster_size, double depth_diff, double dynamic_speed,
                       int32_t neighbor_distance, int n_threads, int reps, int stage_mask, double* best_sec,
                       int64_t* n_objects) {
  if (!cv::mod_ref_colormap_table()) return -20;
  const size_t px = (size_t)w * h;
  const sensor_msgs::CameraInfo info = make_camera_info(w, h, *cam);
  double best = 1e300;
  std::atomic<int64_t> objs_total{0};
  std::atomic<int> failed{0};
  for (int rep = 0; rep < reps; rep++) {
    std::atomic<int> next{0};
    objs_total = 0;
    const auto t0 = std::chrono::steady_clock::now();
    auto worker = [&]() {
      try {
        scene_flow_constructor::SceneFlowConstructor ctor;
        prime_constructor(ctor, w, h, info, dynamic_flow_diff, 1);
        scene_flow_clusterer::ClustererNodelet cl;
        prime_clusterer(cl, cluster_size, depth_diff, dynamic_speed, neighbor_distance, 1);
        for (;;) {
          const int i = next.fetch_add(1);
          if (i >= n_pairs) break;
          // stamps chosen so that (now - prev).toSec() reproduces dt[i] to the nanosecond
          const uint64_t ns = (uint64_t)std::llround(dt[i] * 1e9);
          ref_disparity dn{f, T, min_disp, max_disp, disp_frames + (size_t)now_index[i] * px, w * 4};
          ref_disparity dp{f, T, min_disp, max_disp, disp_frames + (size_t)prev_index[i] * px, w * 4};
          std::shared_ptr<DisparityImageProcessor> dnow(new DisparityImageProcessor(
              make_disparity_msg(w, h, dn, (uint32_t)(1000 + ns / 1000000000ull), (uint32_t)(ns % 1000000000ull)), info));
          std::shared_ptr<DisparityImageProcessor> dprev(new DisparityImageProcessor(make_disparity_msg(w, h, dp, 1000, 0), info));
          std::shared_ptr<cv_bridge::CvImage> flow = make_flow(w, h, flows + (size_t)i * px * 2, w * 8);
          geometry_msgs::TransformPtr tf = make_transform(transforms[i]);
          ctor.pc_with_velocity_pub_.state->n_published = 0;
          ctor.construct(dnow, dprev, flow, tf);
          if (!(stage_mask & 2)) continue;
          if (!ctor.pc_with_velocity_pub_.state->n_published) continue;
          sensor_msgs::PointCloud2ConstPtr msg(
              new sensor_msgs::PointCloud2(*std::any_cast<sensor_msgs::PointCloud2>(&ctor.pc_with_velocity_pub_.state->last)));
          cl.dataCB(msg);
          const moving_object_msgs::MovingObjectArray& m =
              *std::any_cast<moving_object_msgs::MovingObjectArray>(&cl.dynamic_objects_pub_.state->last);
          objs_total += (int64_t)m.moving_object_array.size();
        }
      } catch (...) {
        failed = 1;
      }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; t++) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (sec < best) best = sec;
  }
  if (best_sec) *best_sec = best;
  if (n_objects) *n_objects = objs_total.load();
  return failed ? -11 : 0;
}

/*
 * The reference NODE driven through a sequence of stereo frames: SceneFlowConstructor::stereoCallback
 * (scene_flow_constructor.cpp:364-399) once per frame, with the stand-in sgm_gpu / pwc_net / libviso2 replaying the
 * given script, so that the frame-to-frame state (previous image / disparity, :397-398), the estimator-failure rules
 * (:250-254, :270-276, :284-289), the odometry broadcast (:322-349) and construct() run exactly as written.  `log`
 * receives every publication and tf broadcast in order (oracle/node_sequence.h).
 */
int ref_node_sequence(int w, int h, int n_frames, const mod_node_sequence::camera* cam, float f, float T, float min_d,
                      float max_d, const float* disparities, const float* flows, const double* motions,
                      const int32_t* ok_flags, const uint32_t* stamps, int subscribers, int32_t dynamic_flow_diff,
                      int fused, int32_t cluster_size, char* log, int log_cap, int32_t* shutdown_requested) {
  // fused: the reference's TWO nodes chained the way the launch files wire them - every ~scene_flow cloud goes through
  // ClustererNodelet::dataCB (only ~moving_objects subscribed), whose publication lands in the same log
  std::unique_ptr<scene_flow_clusterer::ClustererNodelet> clusterer;
  size_t seen = 0;
  auto setup = [](scene_flow_constructor::SceneFlowConstructor&) {};
  auto finish = [&](scene_flow_constructor::SceneFlowConstructor& node) {
    if (node.construct_thread_.joinable()) node.construct_thread_.join();
    ros::PubLog* events = ros::mod_ref_pub_log();
    if (!fused || !events) return;
    std::vector<sensor_msgs::PointCloud2ConstPtr> clouds;
    {
      std::lock_guard<std::mutex> l(events->mu);
      for (; seen < events->events.size(); seen++)
        if (events->events[seen].topic == "scene_flow")
          clouds.push_back(std::make_shared<sensor_msgs::PointCloud2>(
              *std::any_cast<sensor_msgs::PointCloud2>(&events->events[seen].msg)));
    }
    for (const auto& cloud : clouds) {
      if (!clusterer) {
        clusterer.reset(new scene_flow_clusterer::ClustererNodelet());
        prime_clusterer(*clusterer, cluster_size, 0.15, 0.3, 4, 1);
      }
      clusterer->dataCB(cloud);
    }
    std::lock_guard<std::mutex> l(events->mu);
    seen = events->events.size();
  };
  return mod_node_sequence::run<scene_flow_constructor::SceneFlowConstructor>(
      w, h, n_frames, cam, f, T, min_d, max_d, disparities, flows, motions, ok_flags, stamps, subscribers,
      dynamic_flow_diff, fused, setup, finish, log, log_cap, shutdown_requested);
}

}  // extern "C"
