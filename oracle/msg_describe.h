// msg_describe.h -- canonical one-line descriptions of the ROS messages the constructor node publishes (everything
// but the payload bytes), shared by oracle/ref_full_wrap.cpp (reference node) and oracle/ros_pkg_wrap.cpp (product
// host classes) so that a test can compare the two sides' message metadata as strings.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <sensor_msgs/Image.h>
#include <sensor_msgs/PointCloud2.h>

#include <cstdio>
#include <string>

namespace mod_describe {

inline std::string header(const std_msgs::Header& h) {
  char b[160];
  std::snprintf(b, sizeof b, "stamp=%u.%09u frame=%s", h.stamp.sec, h.stamp.nsec, h.frame_id.c_str());
  return b;
}

inline std::string image(const char* topic, const sensor_msgs::Image& m) {
  char b[200];
  std::snprintf(b, sizeof b, " w=%u h=%u enc=%s bigendian=%d step=%u bytes=%zu", m.width, m.height, m.encoding.c_str(),
                (int)m.is_bigendian, m.step, m.data.size());
  return std::string(topic) + ": " + header(m.header) + b + "\n";
}

inline std::string cloud(const char* topic, const sensor_msgs::PointCloud2& m) {
  char b[200];
  std::snprintf(b, sizeof b, " w=%u h=%u point_step=%u row_step=%u bigendian=%d dense=%d bytes=%zu fields=", m.width,
                m.height, m.point_step, m.row_step, (int)m.is_bigendian, (int)m.is_dense, m.data.size());
  std::string s = std::string(topic) + ": " + header(m.header) + b;
  for (const auto& f : m.fields) {
    std::snprintf(b, sizeof b, "%s:%u:%d:%u,", f.name.c_str(), f.offset, (int)f.datatype, f.count);
    s += b;
  }
  return s + "\n";
}

}  // namespace mod_describe
