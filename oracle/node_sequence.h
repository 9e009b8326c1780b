// node_sequence.h -- drives a scene_flow_constructor NODE (the reference's, or the product's drop-in) through a
// sequence of stereo frames the way a ROS graph would: one stereoCallback per frame, with the stand-in upstream
// estimators of oracle/ref_shim/ref_shim.h replaying a script (disparity, optical flow, camera motion, failures), and
// renders everything the node published or broadcast on tf, in order, as text.  Shared by oracle/ref_full_wrap.cpp
// (reference node) and oracle/ros_pkg_wrap.cpp (product node) so that a test can compare the two logs as strings.
// TEST INFRASTRUCTURE ONLY.
#pragma once
#include <geometry_msgs/TransformStamped.h>
#include <moving_object_msgs/MovingObjectArray.h>
#include <sensor_msgs/CameraInfo.h>
#include <sensor_msgs/Image.h>
#include <sensor_msgs/PointCloud2.h>

#include <any>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>

#include "msg_describe.h"

namespace mod_node_sequence {

struct camera { double fx, fy, cx, cy, Tx, Ty; };

inline uint64_t fnv1a(uint64_t h, const void* p, size_t n) {
  const uint8_t* b = (const uint8_t*)p;
  for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}
// hash of float payloads with every NaN taken as the same value (the reference writes std::nanf(""), a GPU may
// produce another quiet NaN; tests/helpers.assert_cloud_bit_exact uses the same rule)
inline uint64_t hash_floats(uint64_t h, const uint8_t* p, size_t n_floats, size_t* n_finite = nullptr) {
  for (size_t i = 0; i < n_floats; i++) {
    uint32_t u;
    std::memcpy(&u, p + 4 * i, 4);
    if ((u & 0x7f800000u) == 0x7f800000u && (u & 0x007fffffu)) u = 0x7fc00000u;
    if (n_finite && (u & 0x7f800000u) != 0x7f800000u) ++*n_finite;
    h = fnv1a(h, &u, 4);
  }
  return h;
}
const uint64_t kSeed = 1469598103934665603ull;

inline std::string render(const std::string& topic, const std::any& msg) {
  char b[256];
  if (const sensor_msgs::Image* m = std::any_cast<sensor_msgs::Image>(&msg)) {
    // ~optical_flow echoes the estimator's bytes: hashed raw.  ~depth / ~synthetic_optical_flow: NaN-normalised floats.
    size_t finite = 0;
    uint64_t h = hash_floats(kSeed, m->data.data(), m->data.size() / 4, &finite);
    if (topic == "optical_flow") h = fnv1a(kSeed, m->data.data(), m->data.size());
    std::snprintf(b, sizeof b, "  payload=%016llx finite=%zu\n", (unsigned long long)h, finite);
    return mod_describe::image(topic.c_str(), *m) + b;
  }
  if (const sensor_msgs::PointCloud2* m = std::any_cast<sensor_msgs::PointCloud2>(&msg)) {
    // the six registered fields of every point; the record's pad bytes are uninitialised in the reference
    uint64_t h = kSeed;
    size_t finite = 0;
    const size_t n = (size_t)m->width * m->height;
    for (size_t i = 0; i < n && (i + 1) * m->point_step <= m->data.size(); i++) {
      h = hash_floats(h, &m->data[i * m->point_step], 3, &finite);
      h = hash_floats(h, &m->data[i * m->point_step + 16], 3, &finite);
    }
    std::snprintf(b, sizeof b, "  payload=%016llx finite=%zu\n", (unsigned long long)h, finite);
    return mod_describe::cloud(topic.c_str(), *m) + b;
  }
  if (const geometry_msgs::TransformStamped* m = std::any_cast<geometry_msgs::TransformStamped>(&msg)) {
    std::snprintf(b, sizeof b, " child=%s t=(%a %a %a) q=(%a %a %a %a)\n", m->child_frame_id.c_str(),
                  m->transform.translation.x, m->transform.translation.y, m->transform.translation.z,
                  m->transform.rotation.x, m->transform.rotation.y, m->transform.rotation.z, m->transform.rotation.w);
    return topic + ": " + mod_describe::header(m->header) + b;
  }
  if (const moving_object_msgs::MovingObjectArray* m = std::any_cast<moving_object_msgs::MovingObjectArray>(&msg)) {
    std::string out = topic + ": " + mod_describe::header(m->header);
    std::snprintf(b, sizeof b, " objects=%zu\n", m->moving_object_array.size());
    out += b;
    for (const auto& o : m->moving_object_array) {
      std::snprintf(b, sizeof b, "  id=%d center=(%a %a %a) q=(%g %g %g %g) bbox=(%a %a %a)\n", o.id, o.center.position.x,
                    o.center.position.y, o.center.position.z, o.center.orientation.x, o.center.orientation.y,
                    o.center.orientation.z, o.center.orientation.w, o.bounding_box.x, o.bounding_box.y, o.bounding_box.z);
      out += b;
      // the velocity is a member at the median SPEED (std::sort leaves ties unspecified): the speed is what is pinned
      const float vx = (float)o.velocity.x, vy = (float)o.velocity.y, vz = (float)o.velocity.z;
      std::snprintf(b, sizeof b, "  speed=%a\n", (double)std::sqrt((vx * vx + vy * vy) + vz * vz));
      out += b;
    }
    return out;
  }
  return topic + ": <unknown message type>\n";
}

// ok_flags[k]: bit0 disparity, bit1 camera motion, bit2 optical flow (prev -> k) succeed on frame k.
// subscribers: bit0 ~scene_flow, bit1 ~depth, bit2 ~synthetic_optical_flow, bit3 ~optical_flow.
// finish(node) must wait for the node's construct() worker.  Returns 0, -11 on an exception, -12 when the log does
// not fit; *shutdown_requested tells whether the node asked ros::shutdown().
// fused != 0: the private parameter ~fused_clusterer is set before the node is built (the product's node then runs the
// clusterer in process; the reference's node does not know the parameter - its wrapper chains the reference's
// clusterer nodelet behind ~scene_flow in `finish`).  setup(node) runs once after construction.
template <class Node, class Setup, class Finish>
int run(int w, int h, int n_frames, const camera* cam, float f, float T, float min_d, float max_d,
        const float* disparities, const float* flows, const double* motions, const int32_t* ok_flags,
        const uint32_t* stamps, int subscribers, int32_t dynamic_flow_diff, int fused, Setup setup, Finish finish,
        char* log, int log_cap, int32_t* shutdown_requested) {
  mod_ref_upstream::Script script;
  script.width = w; script.height = h;
  const size_t P = (size_t)w * h;
  for (int k = 0; k < n_frames; k++) {
    mod_ref_upstream::Frame fr;
    fr.disparity_ok = ok_flags[k] & 1; fr.motion_ok = ok_flags[k] & 2; fr.flow_ok = ok_flags[k] & 4;
    fr.disparity.assign(disparities + k * P, disparities + (k + 1) * P);
    fr.f = f; fr.T = T; fr.min_disparity = min_d; fr.max_disparity = max_d;
    fr.flow.assign(flows + k * P * 2, flows + (k + 1) * P * 2);
    std::memcpy(fr.motion, motions + 16 * k, sizeof fr.motion);
    script.frames.push_back(std::move(fr));
  }
  ros::PubLog events;
  std::string text;
  int rc = 0;
  mod_ref_upstream::script() = &script;
  ros::mod_ref_pub_log() = &events;
  ros::mod_ref_shutdown_requested() = false;
  ros::mod_ref_int_params().clear();
  if (fused) ros::mod_ref_int_params().push_back({"fused_clusterer", 1});
  try {
    Node node;
    setup(node);
    if (dynamic_flow_diff >= 0) {
      scene_flow_constructor::SceneFlowConstructorConfig cfg;
      cfg.dynamic_flow_diff = dynamic_flow_diff;
      node.reconfigureCB(cfg, 0);
    }
    node.pc_with_velocity_pub_.state->num_subscribers = (subscribers & 1) ? 1 : 0;
    node.depth_pub_.state->num_subscribers = (subscribers & 2) ? 1 : 0;
    node.static_flow_pub_.state->num_subscribers = (subscribers & 4) ? 1 : 0;
    node.optflow_pub_.state->num_subscribers = (subscribers & 8) ? 1 : 0;

    sensor_msgs::CameraInfoPtr left_info(new sensor_msgs::CameraInfo()), right_info(new sensor_msgs::CameraInfo());
    left_info->width = right_info->width = w; left_info->height = right_info->height = h;
    left_info->P[0] = cam->fx; left_info->P[2] = cam->cx; left_info->P[3] = cam->Tx;
    left_info->P[5] = cam->fy; left_info->P[6] = cam->cy; left_info->P[7] = cam->Ty; left_info->P[10] = 1.0;
    *right_info = *left_info;
    right_info->P[3] = -cam->fx * (double)T;  // right camera of a rectified pair: Tx = -fx * baseline
    left_info->header.frame_id = "left_camera_info_frame";  // differs from the image's on purpose (reference uses the image's, :373)
    right_info->header.frame_id = "right_camera_info_frame";

    for (int k = 0; k < n_frames; k++) {
      sensor_msgs::ImagePtr left(new sensor_msgs::Image()), right(new sensor_msgs::Image());
      left->header.seq = 100 + k; left->header.stamp = ros::Time(stamps[2 * k], stamps[2 * k + 1]);
      left->header.frame_id = "left_camera_optical";
      left->width = w; left->height = h; left->encoding = "mono8"; left->step = w;
      left->data.assign(P, (uint8_t)(17 * k + 3));
      left->data[0] = (uint8_t)k;  // the scripted odometer checks that it is handed this frame's pixels
      *right = *left;
      right->header.frame_id = "right_camera_optical";
      left_info->header.stamp = right_info->header.stamp = left->header.stamp;

      const size_t before = events.events.size();
      node.stereoCallback(left, right, left_info, right_info);
      finish(node);
      char b[64];
      std::snprintf(b, sizeof b, "frame %d\n", k);
      text += b;
      std::lock_guard<std::mutex> l(events.mu);
      for (size_t i = before; i < events.events.size(); i++) text += render(events.events[i].topic, events.events[i].msg);
    }
    finish(node);
  } catch (...) {
    rc = -11;
  }
  ros::mod_ref_int_params().clear();
  mod_ref_upstream::script() = nullptr;
  ros::mod_ref_pub_log() = nullptr;
  if (shutdown_requested) *shutdown_requested = ros::mod_ref_shutdown_requested() ? 1 : 0;
  if (rc == 0 && (int)text.size() + 1 > log_cap) rc = -12;
  if (log && log_cap > 0) std::snprintf(log, (size_t)log_cap, "%s", text.c_str());
  return rc;
}

}  // namespace mod_node_sequence
