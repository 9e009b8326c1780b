// ros_compat.h -- minimal stand-ins for the ROS message types the hot path touches, used ONLY when the real ROS
// headers are not available (this build image has no ROS).  Field names, types and order follow the .msg definitions
// so that the host classes compile unchanged against the real headers on a ROS box (define MOD_HAVE_ROS there):
//   std_msgs/Header, sensor_msgs/{Image,CameraInfo,PointField,PointCloud2,RegionOfInterest}, stereo_msgs/DisparityImage,
//   geometry_msgs/{Vector3,Quaternion,Transform,Point,Pose}, std_msgs/ColorRGBA, visualization_msgs/{Marker,MarkerArray}
//   (the fields cluster2Marker sets), moving_object_msgs/{MovingObject,MovingObjectArray}
//   (moving_object_msgs/msg/MovingObject.msg:1-7, MovingObjectArray.msg:1-2 in the reference tree).
#pragma once
#ifdef MOD_HAVE_ROS
#include <geometry_msgs/Transform.h>
#include <moving_object_msgs/MovingObjectArray.h>
#include <sensor_msgs/CameraInfo.h>
#include <sensor_msgs/Image.h>
#include <sensor_msgs/PointCloud2.h>
#include <stereo_msgs/DisparityImage.h>
#include <visualization_msgs/MarkerArray.h>
#else
#include <array>
#include <cstdint>
#include <string>
#include <vector>

namespace ros {
struct Time {
  uint32_t sec = 0, nsec = 0;
  double toSec() const { return static_cast<double>(sec) + 1e-9 * static_cast<double>(nsec); }
};
struct Duration {  // ros::Duration: int32 sec + int32 nsec, normalised; toSec() = sec + 1e-9 * nsec
  int32_t sec = 0, nsec = 0;
  double toSec() const { return static_cast<double>(sec) + 1e-9 * static_cast<double>(nsec); }
};
inline Duration operator-(const Time& a, const Time& b) {
  int64_t s = static_cast<int64_t>(a.sec) - b.sec, n = static_cast<int64_t>(a.nsec) - b.nsec;
  while (n < 0) { n += 1000000000LL; s--; }
  while (n >= 1000000000LL) { n -= 1000000000LL; s++; }
  Duration d;
  d.sec = static_cast<int32_t>(s);
  d.nsec = static_cast<int32_t>(n);
  return d;
}
}  // namespace ros

namespace std_msgs {
struct Header { uint32_t seq = 0; ros::Time stamp; std::string frame_id; };
}
namespace sensor_msgs {
struct RegionOfInterest { uint32_t x_offset = 0, y_offset = 0, height = 0, width = 0; bool do_rectify = false; };
struct Image {
  std_msgs::Header header;
  uint32_t height = 0, width = 0;
  std::string encoding;
  uint8_t is_bigendian = 0;
  uint32_t step = 0;
  std::vector<uint8_t> data;
};
struct CameraInfo {
  std_msgs::Header header;
  uint32_t height = 0, width = 0;
  std::string distortion_model;
  std::vector<double> D;
  std::array<double, 9> K{}, R{};
  std::array<double, 12> P{};
  uint32_t binning_x = 0, binning_y = 0;
  RegionOfInterest roi;
};
struct PointField {
  enum { INT8 = 1, UINT8, INT16, UINT16, INT32, UINT32, FLOAT32, FLOAT64 };
  std::string name;
  uint32_t offset = 0;
  uint8_t datatype = 0;
  uint32_t count = 0;
};
struct PointCloud2 {
  std_msgs::Header header;
  uint32_t height = 0, width = 0;
  std::vector<PointField> fields;
  bool is_bigendian = false;
  uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
  bool is_dense = false;
};
}  // namespace sensor_msgs
namespace stereo_msgs {
struct DisparityImage {
  std_msgs::Header header;
  sensor_msgs::Image image;  // 32FC1
  float f = 0, T = 0;
  sensor_msgs::RegionOfInterest valid_window;
  float min_disparity = 0, max_disparity = 0, delta_d = 0;
};
}  // namespace stereo_msgs
namespace geometry_msgs {
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Transform { Vector3 translation; Quaternion rotation; };
struct Pose { Point position; Quaternion orientation; };
}  // namespace geometry_msgs
namespace std_msgs {
struct ColorRGBA { float r = 0, g = 0, b = 0, a = 0; };
}
namespace visualization_msgs {
struct Marker {  // visualization_msgs/Marker.msg (the members the clusterer touches plus the ones between them)
  enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4, LINE_LIST = 5, CUBE_LIST = 6, SPHERE_LIST = 7,
         POINTS = 8 };
  enum { ADD = 0, MODIFY = 0, DELETE = 2, DELETEALL = 3 };
  std_msgs::Header header;
  std::string ns;
  int32_t id = 0;
  int32_t type = 0;
  int32_t action = 0;
  geometry_msgs::Pose pose;
  geometry_msgs::Vector3 scale;
  std_msgs::ColorRGBA color;
  std::vector<geometry_msgs::Point> points;
};
struct MarkerArray { std::vector<Marker> markers; };
}  // namespace visualization_msgs
namespace moving_object_msgs {
struct MovingObject {
  int32_t id = 0;
  geometry_msgs::Pose center;
  geometry_msgs::Vector3 velocity;
  geometry_msgs::Vector3 bounding_box;
};
struct MovingObjectArray {
  std_msgs::Header header;
  std::vector<MovingObject> moving_object_array;
};
}  // namespace moving_object_msgs
#endif  // MOD_HAVE_ROS
