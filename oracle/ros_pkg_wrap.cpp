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
#include <visualization_msgs/MarkerArray.h>

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

extern "C" {

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

}  // extern "C"
