// ros_pkg_wrap.cpp -- drives the PRODUCT's drop-in clusterer nodelet
// (moving_object_detector_b200/ros/scene_flow_clusterer/src/clusterer_nodelet.cpp, written against the real ROS API)
// the way a ROS graph would, with the stand-in ROS headers of oracle/ref_shim/ref_shim.h underneath.
//
// TEST INFRASTRUCTURE ONLY.  There is no ROS in this image, so the catkin package cannot be built as it will be on a
// robot; compiling its sources unmodified against the same stand-in headers that carry the REFERENCE's nodelet
// (oracle/_ref/libmod_ref.so) lets a test hand both nodelets the same sensor_msgs/PointCloud2 and compare what they
// publish on ~moving_objects, ~clusters and ~clusters_image message for message (tests/test_gpu_ros_plugin.py).
// Built by `make -C oracle ros_pkg` into oracle/_shim/libclusterer_nodelet_b200.so; needs no reference source.
#include <cstdint>
#include <cstring>

#include "clusterer_nodelet_b200.h"
#include "scene_flow_constructor_b200_node.h"
#include <visualization_msgs/MarkerArray.h>

#include "msg_describe.h"
#include "node_sequence.h"

namespace {
sensor_msgs::PointCloud2Ptr make_cloud_msg(int w, int h, const void* points32) {
  sensor_msgs::PointCloud2Ptr m(new sensor_msgs::PointCloud2());
  m->header.frame_id = "left_camera";
  m->header.stamp = ros::Time(1234, 5678);
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
}  // namespace

struct rb_camera { double fx, fy, cx, cy, Tx, Ty; };
struct rb_disparity { float f, T, min_disparity, max_disparity; const float* data; int32_t step_bytes; };
struct rb_transform { double tx, ty, tz, qx, qy, qz, qw; };

static std::shared_ptr<stereo_msgs::DisparityImage> rb_disparity_msg(int w, int h, const rb_disparity& d, uint32_t sec,
                                                                       uint32_t nsec) {
  auto m = std::make_shared<stereo_msgs::DisparityImage>();
  m->header.stamp = ros::Time(sec, nsec);
  m->header.frame_id = "disparity_frame";
  m->image.width = w; m->image.height = h; m->image.encoding = "32FC1"; m->image.step = (uint32_t)d.step_bytes;
  m->image.data.resize((size_t)d.step_bytes * h);
  for (int r = 0; r < h; r++)
    std::memcpy(&m->image.data[(size_t)r * d.step_bytes], (const char*)d.data + (size_t)r * d.step_bytes, (size_t)w * 4);
  m->f = d.f; m->T = d.T; m->min_disparity = d.min_disparity; m->max_disparity = d.max_disparity;
  return m;
}

extern "C" {

/*
 * SceneFlowConstructorB200::construct (the product's stand-in for SceneFlowConstructor::construct) on one frame pair,
 * with the argument meaning of ref_scene_flow (oracle/ref_full_wrap.cpp).  description receives the metadata of the
 * published messages in the format of oracle/msg_describe.h.
 */
int rosb200_scene_flow(int w, int h, const rb_camera* cam, const rb_disparity* now, const rb_disparity* prev,
                       const float* flow_xy, int32_t flow_step_bytes, const rb_transform* prev2now, uint32_t prev_sec,
                       uint32_t prev_nsec, uint32_t now_sec, uint32_t now_nsec, int32_t dynamic_flow_diff,
                       int subscribers, void* out_points32, float* out_depth, float* out_static_flow,
                       float* out_flow_echo, int32_t* published, char* description, int description_cap) {
  try {
    namespace sfc = scene_flow_constructor;
    sfc::SceneFlowConstructorB200 node(w, h);
    sfc::SceneFlowConstructorConfig cfg;
    cfg.dynamic_flow_diff = dynamic_flow_diff;
    node.reconfigureCB(cfg, 0);
    sensor_msgs::CameraInfo info;
    info.header.frame_id = "left_camera_optical";
    info.width = w; info.height = h;
    info.P[0] = cam->fx; info.P[2] = cam->cx; info.P[3] = cam->Tx;
    info.P[5] = cam->fy; info.P[6] = cam->cy; info.P[7] = cam->Ty; info.P[10] = 1.0;
    node.setCameraInfo(info);
    std::string desc;
    int pub = 0;
    const size_t P = (size_t)w * h;
    node.pc_with_velocity_pub_ = {[&](const sensor_msgs::PointCloud2& m) {
      pub |= 1; desc += mod_describe::cloud("scene_flow", m);
      if (out_points32) std::memcpy(out_points32, m.data.data(), std::min(m.data.size(), P * 32));
    }, (subscribers & 1) ? 1 : 0};
    node.depth_pub_ = {[&](const sensor_msgs::Image& m) {
      pub |= 2; desc += mod_describe::image("depth", m);
      if (out_depth) std::memcpy(out_depth, m.data.data(), std::min(m.data.size(), P * 4));
    }, (subscribers & 2) ? 1 : 0};
    node.static_flow_pub_ = {[&](const sensor_msgs::Image& m) {
      pub |= 4; desc += mod_describe::image("synthetic_optical_flow", m);
      if (out_static_flow) std::memcpy(out_static_flow, m.data.data(), std::min(m.data.size(), P * 8));
    }, (subscribers & 4) ? 1 : 0};
    node.optflow_pub_ = {[&](const sensor_msgs::Image& m) {
      pub |= 8; desc += mod_describe::image("optical_flow", m);
      if (out_flow_echo) {
        for (int r = 0; r < h; r++) std::memcpy((char*)out_flow_echo + (size_t)r * w * 8, &m.data[(size_t)r * m.step], (size_t)w * 8);
      }
    }, (subscribers & 8) ? 1 : 0};

    std::shared_ptr<stereo_msgs::DisparityImage> dnow, dprev;
    if (now) dnow = rb_disparity_msg(w, h, *now, now_sec, now_nsec);
    if (prev) dprev = rb_disparity_msg(w, h, *prev, prev_sec, prev_nsec);
    std::shared_ptr<sensor_msgs::Image> flow;
    if (flow_xy) {
      flow = std::make_shared<sensor_msgs::Image>();
      flow->header.seq = 42; flow->header.stamp = ros::Time(now_sec + 1, now_nsec); flow->header.frame_id = "flow_frame";
      flow->width = w; flow->height = h; flow->encoding = "32FC2"; flow->step = (uint32_t)w * 8;  // cv_bridge emits dense rows
      flow->data.resize(P * 8);
      for (int r = 0; r < h; r++) std::memcpy(&flow->data[(size_t)r * w * 8], (const char*)flow_xy + (size_t)r * flow_step_bytes, (size_t)w * 8);
    }
    std::shared_ptr<geometry_msgs::Transform> tf;
    if (prev2now) {
      tf = std::make_shared<geometry_msgs::Transform>();
      tf->translation.x = prev2now->tx; tf->translation.y = prev2now->ty; tf->translation.z = prev2now->tz;
      tf->rotation.x = prev2now->qx; tf->rotation.y = prev2now->qy; tf->rotation.z = prev2now->qz; tf->rotation.w = prev2now->qw;
    }
    node.construct(dnow, dprev, flow, tf);
    if (published) *published = pub;
    if (description && description_cap > 0) std::snprintf(description, (size_t)description_cap, "%s", desc.c_str());
    return node.lastError().empty() ? 0 : -13;
  } catch (...) {
    return -11;
  }
}

/* topics the nodelet advertises and the defaults it latches at start-up: out3 = cluster_size, neighbor_distance,
 * n_topics_ok (3 when ~moving_objects, ~clusters, ~clusters_image are advertised under those names); out2d = depth_diff,
 * dynamic_speed */
int rosb200_defaults(int32_t* out3, double* out2d) {
  scene_flow_clusterer::ClustererNodelet n;
  n.onInit();
  out3[0] = n.config_.cluster_size; out3[1] = n.config_.neighbor_distance;
  out3[2] = (n.moving_objects_pub_.state && n.moving_objects_pub_.state->topic == "moving_objects") +
            (n.clusters_pub_.state && n.clusters_pub_.state->topic == "clusters") +
            (n.clusters_image_pub_.state && n.clusters_image_pub_.state->topic == "clusters_image");
  out2d[0] = n.config_.depth_diff; out2d[1] = n.config_.dynamic_speed;
  return 0;
}

/* One PointCloud2 through ClustererNodelet::dataCB.  Same argument meaning as ref_cluster (oracle/ref_full_wrap.cpp)
 * minus the internal maps: subscribers bit0 ~moving_objects, bit1 ~clusters, bit2 ~clusters_image; out_objs 14 doubles
 * per object; markers as offsets / xyz doubles / rgba floats, *n_markers excluding the leading DELETEALL marker whose
 * presence sets bit 3 of *published; header_ok = every published header equals the cloud's. */
int rosb200_cluster(int w, int h, const void* points32, int32_t cluster_size, double depth_diff, double dynamic_speed,
                    int32_t neighbor_distance, int subscribers, double* out_objs, int32_t max_objs, int32_t* n_objs,
                    uint8_t* out_bgr8, int32_t* marker_offsets, double* marker_points, float* marker_rgba,
                    int32_t max_markers, int32_t* n_markers, int32_t* published, int32_t* header_ok) {
  try {
    scene_flow_clusterer::ClustererNodelet node;
    node.onInit();
    scene_flow_clusterer::ClustererConfig cfg;
    cfg.cluster_size = cluster_size; cfg.depth_diff = depth_diff; cfg.dynamic_speed = dynamic_speed;
    cfg.neighbor_distance = neighbor_distance;
    node.reconfigureCB(cfg, 0);
    node.moving_objects_pub_.state->num_subscribers = (subscribers & 1) ? 1 : 0;
    node.clusters_pub_.state->num_subscribers = (subscribers & 2) ? 1 : 0;
    node.clusters_image_pub_.state->num_subscribers = (subscribers & 4) ? 1 : 0;
    sensor_msgs::PointCloud2ConstPtr msg = make_cloud_msg(w, h, points32);

    node.dataCB(msg);

    const size_t n = (size_t)w * h;
    int pub = 0, rc = 0, hdr_ok = 1;
    auto same_header = [&](const std_msgs::Header& hd) {
      return hd.frame_id == msg->header.frame_id && hd.stamp.sec == msg->header.stamp.sec && hd.stamp.nsec == msg->header.stamp.nsec;
    };
    if (n_objs) *n_objs = 0;
    if (n_markers) *n_markers = 0;
    if (node.moving_objects_pub_.state->n_published) {
      pub |= 1;
      const moving_object_msgs::MovingObjectArray& m =
          *std::any_cast<moving_object_msgs::MovingObjectArray>(&node.moving_objects_pub_.state->last);
      hdr_ok = hdr_ok && same_header(m.header);
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
    }
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
        hdr_ok = hdr_ok && same_header(k.header);
        if (k.id != i || k.type != visualization_msgs::Marker::POINTS || k.action != visualization_msgs::Marker::ADD ||
            k.scale.x != 0.05 || k.scale.y != 0.05 || k.scale.z != 0)
          rc = -4;
        marker_offsets[i] = (int32_t)off;
        for (const geometry_msgs::Point& p : k.points) {
          if (marker_points) { marker_points[3 * off] = p.x; marker_points[3 * off + 1] = p.y; marker_points[3 * off + 2] = p.z; }
          off++;
        }
        marker_offsets[i + 1] = (int32_t)off;
        if (marker_rgba) { marker_rgba[4 * i] = k.color.r; marker_rgba[4 * i + 1] = k.color.g; marker_rgba[4 * i + 2] = k.color.b; marker_rgba[4 * i + 3] = k.color.a; }
      }
    }
    if (node.clusters_image_pub_.state->n_published) {
      pub |= 4;
      const sensor_msgs::Image& m = *std::any_cast<sensor_msgs::Image>(&node.clusters_image_pub_.state->last);
      hdr_ok = hdr_ok && same_header(m.header) && m.encoding == "bgr8" && m.width == (uint32_t)w && m.height == (uint32_t)h && m.step == 3u * w;
      if (out_bgr8) std::memcpy(out_bgr8, m.data.data(), std::min(m.data.size(), n * 3));
    }
    if (published) *published = pub;
    if (header_ok) *header_ok = hdr_ok;
    return rc;
  } catch (...) {
    return -11;
  }
}

/*
 * The product's drop-in constructor NODE (moving_object_detector_b200/ros/scene_flow_constructor) driven through a
 * sequence of stereo frames; same arguments and log format as ref_node_sequence (oracle/ref_full_wrap.cpp).
 */
int rosb200_node_sequence(int w, int h, int n_frames, const mod_node_sequence::camera* cam, float f, float T, float min_d,
                          float max_d, const float* disparities, const float* flows, const double* motions,
                          const int32_t* ok_flags, const uint32_t* stamps, int subscribers, int32_t dynamic_flow_diff,
                          int fused, int32_t cluster_size, char* log, int log_cap, int32_t* shutdown_requested) {
  auto setup = [&](scene_flow_constructor::SceneFlowConstructor& node) {
    if (!fused) return;
    node.clusterer_config_.cluster_size = cluster_size;  // what a dynamic_reconfigure request on ~clusterer would set
    node.moving_objects_pub_.state->num_subscribers = 1;
  };
  auto finish = [](scene_flow_constructor::SceneFlowConstructor& node) {
    if (node.construct_job_.valid()) node.construct_job_.wait();
  };
  return mod_node_sequence::run<scene_flow_constructor::SceneFlowConstructor>(
      w, h, n_frames, cam, f, T, min_d, max_d, disparities, flows, motions, ok_flags, stamps, subscribers,
      dynamic_flow_diff, fused, setup, finish, log, log_cap, shutdown_requested);
}

/* what the constructor node advertises and latches at start-up: out[0] = dynamic_flow_diff default, out[1] = number of
 * the four topics advertised under the reference's names */
int rosb200_constructor_defaults(int32_t* out2) {
  scene_flow_constructor::SceneFlowConstructor n;
  out2[0] = n.config_.dynamic_flow_diff;
  out2[1] = (n.depth_pub_.state && n.depth_pub_.state->topic == "depth") +
            (n.optflow_pub_.state && n.optflow_pub_.state->topic == "optical_flow") +
            (n.pc_with_velocity_pub_.state && n.pc_with_velocity_pub_.state->topic == "scene_flow") +
            (n.static_flow_pub_.state && n.static_flow_pub_.state->topic == "synthetic_optical_flow");
  return 0;
}

}  // extern "C"
