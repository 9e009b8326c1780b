// ref_shim.h -- stand-ins for the third-party headers the REFERENCE's hot-path translation units include
// (roscpp, message types, image_geometry, tf2, tf2_eigen, Eigen, PCL, OpenCV core, cv_bridge, image_transport,
// message_filters, dynamic_reconfigure, nodelet, pluginlib, libviso2, sgm_gpu, pwc_net).
//
// TEST INFRASTRUCTURE ONLY.  None of ROS Melodic / PCL 1.8 / Eigen 3.3 / OpenCV 3.2 is installed in this image, so the
// reference cannot be built as shipped.  With this header on the include path (oracle/Makefile writes one forwarding
// header per <path/name.h> the reference includes into oracle/_shim/include/), the reference's OWN sources
//   disparity_image_proc/src/disparity_image_processor.cpp
//   scene_flow_constructor/src/scene_flow_constructor.cpp
//   scene_flow_clusterer/src/{clusterer_nodelet,color_set,lookup_table}.cpp
//   image_crop/src/image_crop.cpp
// compile UNMODIFIED from where they lie under /root/reference into oracle/_ref/libmod_ref.so.  Every loop, branch,
// cast and container access of the hot path is then the reference's; what is restated HERE is only the arithmetic that
// lives in the un-vendored dependencies, each piece citing the upstream function it follows:
//   image_geometry::PinholeCameraModel::{fromCameraInfo, projectPixelTo3dRay, project3dToPixel}
//   tf2::transformToEigen + Eigen::Quaterniond::toRotationMatrix + Isometry3d * Vector3d
//   Eigen::Vector3f::norm, Vector4f +,-,/ ; pcl::getMinMax3D ; pcl::PointCloud::at ; pcl::to/fromROSMsg
//   cv::Mat::at, cv::Vec2f::dot, cv::applyColorMap(COLORMAP_HSV) (table supplied by the caller, generated from cv2)
//   ros::Time - ros::Time, ros::Duration::toSec ; dynamic_reconfigure::Server::setCallback (calls back with defaults)
// Everything else (node handles, publishers, subscribers, tf buffers, the upstream estimators) is inert plumbing:
// publishers record the last message so the wrapper (ref_full_wrap.cpp) can read what the reference would have sent.
#ifndef MOD_REF_SHIM_H_
#define MOD_REF_SHIM_H_

#include <algorithm>
#include <any>
#include <array>
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <iostream>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

// ---------------------------------------------------------------------------------------------------------- boost
namespace boost {
using std::bind;
template <class S> using function = std::function<S>;
template <class T> using shared_ptr = std::shared_ptr<T>;
}  // namespace boost
using namespace std::placeholders;  // boost/bind.hpp puts _1, _2 ... in the global namespace

// ----------------------------------------------------------------------------------------------------- rosconsole
#define MOD_REF_NOLOG(...) do { } while (0)
#define ROS_DEBUG(...) MOD_REF_NOLOG()
#define ROS_INFO(...) MOD_REF_NOLOG()
#define ROS_WARN(...) MOD_REF_NOLOG()
#define ROS_ERROR(...) MOD_REF_NOLOG()
#define ROS_FATAL(...) MOD_REF_NOLOG()
#define ROS_DEBUG_STREAM(x) MOD_REF_NOLOG()
#define ROS_INFO_STREAM(x) MOD_REF_NOLOG()
#define ROS_WARN_STREAM(x) MOD_REF_NOLOG()
#define ROS_ERROR_STREAM(x) MOD_REF_NOLOG()
#define ROS_ERROR_THROTTLE(...) MOD_REF_NOLOG()
#define ROS_ASSERT(x) MOD_REF_NOLOG()
#define NODELET_INFO(...) MOD_REF_NOLOG()
#define NODELET_INFO_STREAM(x) MOD_REF_NOLOG()
#define NODELET_DEBUG(...) MOD_REF_NOLOG()
#define NODELET_WARN(...) MOD_REF_NOLOG()
#define NODELET_ERROR(...) MOD_REF_NOLOG()
#define NODELET_ERROR_STREAM(x) MOD_REF_NOLOG()
#define NODELET_WARN_STREAM(x) MOD_REF_NOLOG()
#define PLUGINLIB_EXPORT_CLASS(cls, base)

// ------------------------------------------------------------------------------------------------------- rostime
namespace ros {
// rostime Duration: int32 sec / nsec, normalised so that 0 <= nsec < 1e9; toSec() = (double)sec + 1e-9 * (double)nsec
struct Duration {
  int32_t sec = 0, nsec = 0;
  Duration() {}
  Duration& fromNSec(int64_t t) {
    int64_t s = t / 1000000000LL, n = t % 1000000000LL;
    if (n < 0) { n += 1000000000LL; s -= 1; }
    sec = (int32_t)s; nsec = (int32_t)n;
    return *this;
  }
  double toSec() const { return (double)sec + 1e-9 * (double)nsec; }
};
// rostime TimeBase<T,D>::operator-: D().fromNSec(toNSec() - rhs.toNSec())
struct Time {
  uint32_t sec = 0, nsec = 0;
  Time() {}
  Time(uint32_t s, uint32_t n) : sec(s), nsec(n) {}
  static Time now() { return Time(); }
  uint64_t toNSec() const { return (uint64_t)sec * 1000000000ull + (uint64_t)nsec; }
  Duration operator-(const Time& r) const { return Duration().fromNSec((int64_t)toNSec() - (int64_t)r.toNSec()); }
};
struct WallDuration { double toSec() const { return 0.0; } };
struct WallTime {
  static WallTime now() { return WallTime(); }
  WallDuration operator-(const WallTime&) const { return WallDuration(); }
};
}  // namespace ros

// ------------------------------------------------------------------------------------------------------ messages
namespace std_msgs {
struct Header { uint32_t seq = 0; ros::Time stamp; std::string frame_id; };
}  // namespace std_msgs

namespace geometry_msgs {
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 0; };
struct Pose { Point position; Quaternion orientation; };
struct Transform { Vector3 translation; Quaternion rotation; };
struct TransformStamped { std_msgs::Header header; std::string child_frame_id; Transform transform; };
typedef std::shared_ptr<Transform> TransformPtr;
typedef std::shared_ptr<const Transform> TransformConstPtr;
}  // namespace geometry_msgs

namespace sensor_msgs {
struct RegionOfInterest { uint32_t x_offset = 0, y_offset = 0, height = 0, width = 0; bool do_rectify = false; };
struct CameraInfo {
  std_msgs::Header header;
  uint32_t height = 0, width = 0;
  std::string distortion_model;
  std::vector<double> D;
  std::array<double, 9> K{}, R{};   // boost::array in ROS: operator[], data(), size()
  std::array<double, 12> P{};
  uint32_t binning_x = 0, binning_y = 0;
  RegionOfInterest roi;
};
typedef std::shared_ptr<CameraInfo> CameraInfoPtr;
typedef std::shared_ptr<const CameraInfo> CameraInfoConstPtr;
struct Image {
  std_msgs::Header header;
  uint32_t height = 0, width = 0;
  std::string encoding;
  uint8_t is_bigendian = 0;
  uint32_t step = 0;
  std::vector<uint8_t> data;
};
typedef std::shared_ptr<Image> ImagePtr;
typedef std::shared_ptr<const Image> ImageConstPtr;
struct PointField {
  enum { INT8 = 1, UINT8 = 2, INT16 = 3, UINT16 = 4, INT32 = 5, UINT32 = 6, FLOAT32 = 7, FLOAT64 = 8 };
  std::string name;
  uint32_t offset = 0;
  uint8_t datatype = 0;
  uint32_t count = 0;
};
struct PointCloud2 {
  std_msgs::Header header;
  uint32_t height = 0, width = 0;
  std::vector<PointField> fields;
  uint8_t is_bigendian = 0;
  uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
  uint8_t is_dense = 0;
};
typedef std::shared_ptr<PointCloud2> PointCloud2Ptr;
typedef std::shared_ptr<const PointCloud2> PointCloud2ConstPtr;
namespace image_encodings {
const std::string MONO8 = "mono8";
const std::string BGR8 = "bgr8";
const std::string TYPE_32FC1 = "32FC1";
const std::string TYPE_32FC2 = "32FC2";
}  // namespace image_encodings
}  // namespace sensor_msgs

namespace stereo_msgs {
struct DisparityImage {
  std_msgs::Header header;
  sensor_msgs::Image image;
  float f = 0, T = 0;
  sensor_msgs::RegionOfInterest valid_window;
  float min_disparity = 0, max_disparity = 0, delta_d = 0;
};
typedef std::shared_ptr<DisparityImage> DisparityImagePtr;
typedef std::shared_ptr<const DisparityImage> DisparityImageConstPtr;
}  // namespace stereo_msgs

namespace std_msgs { struct ColorRGBA { float r = 0, g = 0, b = 0, a = 0; }; }

namespace visualization_msgs {
struct Marker {
  enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4, LINE_LIST = 5, CUBE_LIST = 6, SPHERE_LIST = 7,
         POINTS = 8 };
  enum { ADD = 0, MODIFY = 0, DELETE = 2, DELETEALL = 3 };
  std_msgs::Header header;
  std::string ns;
  int32_t id = 0, type = 0, action = 0;
  geometry_msgs::Pose pose;
  geometry_msgs::Vector3 scale;
  std_msgs::ColorRGBA color;
  std::vector<geometry_msgs::Point> points;
};
struct MarkerArray { std::vector<Marker> markers; };
}  // namespace visualization_msgs

// moving_object_msgs/msg/MovingObject.msg, MovingObjectArray.msg (generated code in the reference build)
namespace moving_object_msgs {
struct MovingObject {
  int32_t id = 0;
  geometry_msgs::Pose center;
  geometry_msgs::Vector3 velocity;
  geometry_msgs::Vector3 bounding_box;
};
struct MovingObjectArray { std_msgs::Header header; std::vector<MovingObject> moving_object_array; };
}  // namespace moving_object_msgs

// ------------------------------------------------------------------------------------------------------- roscpp
namespace ros {
inline long& mod_ref_publish_counter() { static thread_local long n = 0; return n; }  // orders publications across topics
struct PubState { std::string topic; int num_subscribers = 0; int n_published = 0; long order = 0; std::any last; };
// Optional process-wide record of EVERYTHING that is published or broadcast on tf, in order and from whichever thread
// (the constructor node publishes from a worker): a wrapper that drives a node through several frames installs one.
struct PubEvent { std::string topic; std::any msg; };
struct PubLog {
  std::mutex mu;
  std::vector<PubEvent> events;
  void add(const std::string& topic, std::any msg) { std::lock_guard<std::mutex> l(mu); events.push_back(PubEvent{topic, std::move(msg)}); }
};
inline PubLog*& mod_ref_pub_log() { static PubLog* log = nullptr; return log; }
inline bool& mod_ref_shutdown_requested() { static bool flag = false; return flag; }
inline void shutdown() { mod_ref_shutdown_requested() = true; }
class Publisher {
 public:
  std::shared_ptr<PubState> state;
  uint32_t getNumSubscribers() const { return state ? (uint32_t)state->num_subscribers : 0u; }
  template <class M> void publish(const M& msg) const {
    if (!state) return;
    state->last = msg; state->n_published++; state->order = ++mod_ref_publish_counter();
    if (PubLog* log = mod_ref_pub_log()) log->add(state->topic, state->last);
  }
  template <class M> void publish(const std::shared_ptr<M>& msg) const { publish(typename std::remove_const<M>::type(*msg)); }
};
class Subscriber {};
// integer parameters a test wants the nodes to find on the parameter server (image_crop's target_width / target_height)
inline std::vector<std::pair<std::string, int>>& mod_ref_int_params() {
  static thread_local std::vector<std::pair<std::string, int>> params;
  return params;
}
class NodeHandle {
 public:
  NodeHandle() {}
  explicit NodeHandle(const std::string&) {}
  NodeHandle(const NodeHandle&, const std::string&) {}
  template <class T> bool getParam(const std::string&, T&) const { return false; }
  bool getParam(const std::string& name, int& v) const {
    for (const auto& kv : mod_ref_int_params())
      if (kv.first == name) { v = kv.second; return true; }
    return false;
  }
  template <class T, class U> bool param(const std::string& name, T& v, const U& d) const {
    v = d;
    if constexpr (std::is_same<T, bool>::value || std::is_same<T, int>::value) {  // a test may have set it (ints and flags)
      for (const auto& kv : mod_ref_int_params())
        if (kv.first == name) { v = static_cast<T>(kv.second); return true; }
    }
    return false;
  }
  std::string resolveName(const std::string& n) const { return n; }
  template <class M> Publisher advertise(const std::string& topic, uint32_t) {
    Publisher p; p.state = std::make_shared<PubState>(); p.state->topic = topic; return p;
  }
  template <class M, class T>
  Subscriber subscribe(const std::string&, uint32_t, void (T::*)(const std::shared_ptr<const M>&), T*) { return Subscriber(); }
};
}  // namespace ros

namespace nodelet {
class Nodelet {
 public:
  virtual ~Nodelet() {}
  virtual void onInit() = 0;
 protected:
  ros::NodeHandle& getNodeHandle() { return nh_; }
  ros::NodeHandle& getPrivateNodeHandle() { return pnh_; }
 private:
  ros::NodeHandle nh_, pnh_;
};
}  // namespace nodelet

// dynamic_reconfigure: Server::setCallback stores the callback and immediately calls it with the current
// configuration (dynamic_reconfigure/server.h: setCallback -> callCallback(config_, ~0)); with an empty parameter
// server that is the .cfg defaults.  Config structs restate cfg/Clusterer.cfg:8-13 and cfg/SceneFlowConstructor.cfg:8-11.
namespace dynamic_reconfigure {
template <class ConfigT> class Server {
 public:
  typedef boost::function<void(ConfigT&, uint32_t)> CallbackType;
  explicit Server(const ros::NodeHandle&) {}
  void setCallback(const CallbackType& cb) { callback_ = cb; if (callback_) callback_(config_, ~0u); }
 private:
  CallbackType callback_;
  ConfigT config_;
};
}  // namespace dynamic_reconfigure
namespace scene_flow_clusterer {
struct ClustererConfig { int cluster_size = 2500; double depth_diff = 0.15; double dynamic_speed = 0.3; int neighbor_distance = 4; };
}
namespace scene_flow_constructor {
struct SceneFlowConstructorConfig { int dynamic_flow_diff = 5; double max_color_velocity = 1.0; };
}

// ------------------------------------------------------------------------------------------------- OpenCV (core)
typedef unsigned char uchar;
#define CV_8U 0
#define CV_32F 5
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_8UC4 CV_MAKETYPE(CV_8U, 4)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
namespace cv {
template <class T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<int> Point2i;
typedef Point_<double> Point2d;
typedef Point2i Point;
template <class T> struct Point3_ {
  T x, y, z;
  Point3_() : x(0), y(0), z(0) {}
  Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
typedef Point3_<double> Point3d;
// cv::Vec<_Tp,cn> (core/matx.hpp): dot() accumulates in _Tp starting from 0: s = 0; s += a[i]*b[i]
template <class T, int N> struct Vec {
  T val[N];
  Vec() { for (int i = 0; i < N; i++) val[i] = T(0); }
  Vec(T a, T b) { static_assert(N == 2, "Vec2"); val[0] = a; val[1] = b; }
  Vec(T a, T b, T c) { static_assert(N == 3, "Vec3"); val[0] = a; val[1] = b; val[2] = c; }
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
  Vec operator-(const Vec& o) const { Vec r; for (int i = 0; i < N; i++) r.val[i] = val[i] - o.val[i]; return r; }
  T dot(const Vec& o) const { T s = 0; for (int i = 0; i < N; i++) s += val[i] * o.val[i]; return s; }
};
typedef Vec<float, 2> Vec2f;
typedef Vec<uchar, 3> Vec3b;

// cv::Mat restricted to 2-D; at<T>() indexes data + row*step + col*sizeof(T) exactly like core/mat.inl.hpp
// (release builds do not bounds-check).  at<T>(i) on a continuous matrix is the linear element index.
struct Rect {
  int x = 0, y = 0, width = 0, height = 0;
  Rect() {}
  Rect(int x_, int y_, int w_, int h_) : x(x_), y(y_), width(w_), height(h_) {}
};
struct MatStep {  // cv::MatStep: converts to size_t (bytes per row), step[0] is the same number
  size_t p[2] = {0, 0};
  MatStep() {}
  MatStep(size_t s) { p[0] = s; }
  operator size_t() const { return p[0]; }
  MatStep& operator=(size_t s) { p[0] = s; return *this; }
  size_t operator[](int i) const { return p[i]; }
};
class Mat {
 public:
  int rows = 0, cols = 0;
  uchar* data = nullptr;
  MatStep step;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void* ext, size_t step_) : rows(r), cols(c), data((uchar*)ext), step(step_), type_(type) {
    if ((size_t)step == 0) step = (size_t)c * elemSize();
  }
  void create(int r, int c, int type) {
    if (own_ && rows == r && cols == c && type_ == type) return;
    rows = r; cols = c; type_ = type; step = (size_t)c * elemSize();
    own_ = std::shared_ptr<uchar>(new uchar[std::max<size_t>(1, step * (size_t)r)], std::default_delete<uchar[]>());
    data = own_.get();
  }
  int type() const { return type_; }
  size_t elemSize() const { static const size_t dsz[8] = {1, 1, 2, 2, 4, 4, 8, 2}; return dsz[type_ & 7] * (size_t)((type_ >> CV_CN_SHIFT) + 1); }
  bool isContinuous() const { return step == (size_t)cols * elemSize() || rows <= 1; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  template <class T> T& at(int r, int c) { return *(T*)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
  template <class T> const T& at(int r, int c) const { return *(const T*)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
  template <class T> T& at(Point2i p) { return at<T>(p.y, p.x); }
  template <class T> const T& at(Point2i p) const { return at<T>(p.y, p.x); }
  // Mat::operator()(const Rect&): a view into the same pixels; core/src/matrix.cpp asserts the ROI lies inside
  Mat operator()(const Rect& r) const {
    // (OpenCV evaluates roi.x + roi.width in int; where that overflows its behaviour is undefined -- evaluated in 64 bits here)
    if (!(0 <= r.x && 0 <= r.width && (long long)r.x + r.width <= cols && 0 <= r.y && 0 <= r.height &&
          (long long)r.y + r.height <= rows))
      throw std::runtime_error("cv::Mat::Mat(roi): assertion failed, roi outside the matrix");
    Mat m(*this);
    m.rows = r.height; m.cols = r.width;
    m.data = data + (size_t)r.y * step + (size_t)r.x * elemSize();
    return m;
  }
  void copyTo(Mat& dst) const {
    dst.create(rows, cols, type_);
    for (int r = 0; r < rows; r++) std::memcpy(dst.data + (size_t)r * dst.step, data + (size_t)r * step, (size_t)cols * elemSize());
  }
  template <class T> T& at(int i0) {
    if (isContinuous() || rows == 1) return ((T*)data)[i0];
    if (cols == 1) return *(T*)(data + (size_t)i0 * step);
    int r = i0 / cols;
    return at<T>(r, i0 - r * cols);
  }
 private:
  int type_ = 0;
  std::shared_ptr<uchar> own_;
};
template <class T> struct MatDepth;
template <> struct MatDepth<float> { enum { type = CV_32FC1 }; };
template <class T> class Mat_ : public Mat {
 public:
  Mat_(int r, int c, T* ext, size_t step_) : Mat(r, c, MatDepth<T>::type, ext, step_) {}
};

// cv::applyColorMap(COLORMAP_HSV): per-pixel lookup of an 8UC1 image in a 256-entry BGR table.  The table is not
// restated: the wrapper installs the one generated from the cv2 package (tests/golden/hsv_colormap_bgr.npy).
enum ColormapTypes { COLORMAP_AUTUMN = 0, COLORMAP_JET = 2, COLORMAP_HSV = 9 };
const uchar*& mod_ref_colormap_table();
inline void applyColorMap(const Mat& src, Mat& dst, int) {
  const uchar* lut = mod_ref_colormap_table();
  if (!lut) throw std::runtime_error("ref_shim: colour map table not installed");
  dst.create(src.rows, src.cols, CV_8UC3);
  for (int r = 0; r < src.rows; r++)
    for (int c = 0; c < src.cols; c++) {
      const uchar v = src.at<uchar>(r, c);
      dst.at<Vec3b>(r, c) = Vec3b(lut[3 * v], lut[3 * v + 1], lut[3 * v + 2]);
    }
}
}  // namespace cv

// ------------------------------------------------------------------------------------------------------ cv_bridge
namespace cv_bridge {
class CvImage {
 public:
  std_msgs::Header header;
  std::string encoding;
  cv::Mat image;
  CvImage() {}
  CvImage(const std_msgs::Header& h, const std::string& e, const cv::Mat& m = cv::Mat()) : header(h), encoding(e), image(m) {}
  // cv_bridge::CvImage::toImageMsg: rows copied densely, step = cols * elemSize
  sensor_msgs::ImagePtr toImageMsg() const {
    sensor_msgs::ImagePtr m = std::make_shared<sensor_msgs::Image>();
    m->header = header; m->encoding = encoding; m->height = image.rows; m->width = image.cols;
    m->step = (uint32_t)(image.cols * image.elemSize());
    m->data.resize((size_t)m->step * m->height);
    for (int r = 0; r < image.rows; r++) std::memcpy(&m->data[(size_t)r * m->step], image.data + (size_t)r * image.step, m->step);
    return m;
  }
};
typedef std::shared_ptr<CvImage> CvImagePtr;
typedef std::shared_ptr<const CvImage> CvImageConstPtr;
// cv_bridge::toCvCopy(msg, "mono8"): a deep copy when the source already is mono8.  Colour sources would go through
// cv::cvtColor's fixed-point BGR2GRAY, which is not restated (it feeds libviso2, an upstream estimator): tests that drive
// the constructor node through stereoCallback hand it mono8 images.
inline CvImagePtr toCvCopy(const sensor_msgs::ImageConstPtr& src, const std::string& encoding) {
  if (encoding != "mono8" || src->encoding != "mono8") throw std::runtime_error("ref_shim: toCvCopy only mono8 -> mono8");
  CvImagePtr out = std::make_shared<CvImage>();
  out->header = src->header;
  out->encoding = encoding;
  cv::Mat(src->height, src->width, CV_8UC1, const_cast<uint8_t*>(src->data.data()), src->step).copyTo(out->image);
  return out;
}
// cv_bridge::toCvShare(msg, "bgr8") (cv_bridge/src/cv_bridge.cpp): shares the message's pixels when the encodings are
// equal, otherwise converts with cv::cvtColor.  The colour conversions among 8-bit bgr8 / rgb8 / bgra8 / rgba8 / mono8
// are pure channel shuffles (RGB2BGR swaps, BGRA2BGR drops alpha, GRAY2BGR replicates), restated here.
inline CvImageConstPtr toCvShare(const sensor_msgs::ImageConstPtr& src, const std::string& encoding) {
  if (encoding != "bgr8") throw std::runtime_error("ref_shim: toCvShare only to bgr8");
  CvImagePtr out = std::make_shared<CvImage>();
  out->header = src->header;
  out->encoding = encoding;
  const int h = (int)src->height, w = (int)src->width;
  uint8_t* base = const_cast<uint8_t*>(src->data.data());
  if (src->encoding == "bgr8") {
    out->image = cv::Mat(h, w, CV_8UC3, base, src->step);
    return out;
  }
  int bpp, ib, ig, ir;
  if (src->encoding == "rgb8") { bpp = 3; ib = 2; ig = 1; ir = 0; }
  else if (src->encoding == "bgra8") { bpp = 4; ib = 0; ig = 1; ir = 2; }
  else if (src->encoding == "rgba8") { bpp = 4; ib = 2; ig = 1; ir = 0; }
  else if (src->encoding == "mono8") { bpp = 1; ib = 0; ig = 0; ir = 0; }
  else throw std::runtime_error("ref_shim: unsupported source encoding " + src->encoding);
  out->image.create(h, w, CV_8UC3);
  for (int r = 0; r < h; r++)
    for (int c = 0; c < w; c++) {
      const uint8_t* p = base + (size_t)r * src->step + (size_t)c * bpp;
      out->image.at<cv::Vec3b>(r, c) = cv::Vec3b(p[ib], p[ig], p[ir]);
    }
  return out;
}
}  // namespace cv_bridge

// ------------------------------------------------------------------------------- image_transport, message_filters
namespace image_transport {
class Publisher : public ros::Publisher {};
struct CameraPubState { int n_published = 0; sensor_msgs::Image image; sensor_msgs::CameraInfo info; };
class CameraPublisher {
 public:
  std::shared_ptr<CameraPubState> state;
  void publish(const sensor_msgs::Image& image, const sensor_msgs::CameraInfo& info) const {
    if (state) { state->image = image; state->info = info; state->n_published++; }
  }
};
class CameraSubscriber {};
class ImageTransport {
 public:
  explicit ImageTransport(const ros::NodeHandle&) {}
  Publisher advertise(const std::string& topic, uint32_t) {
    Publisher p; p.state = std::make_shared<ros::PubState>(); p.state->topic = topic; return p;
  }
  CameraPublisher advertiseCamera(const std::string&, uint32_t);
  template <class T>
  CameraSubscriber subscribeCamera(const std::string&, uint32_t,
                                   void (T::*)(const sensor_msgs::ImageConstPtr&, const sensor_msgs::CameraInfoConstPtr&), T*);
};
class SubscriberFilter {
 public:
  void subscribe(ImageTransport&, const std::string&, uint32_t) {}
};

inline CameraPublisher ImageTransport::advertiseCamera(const std::string&, uint32_t) {
  CameraPublisher p; p.state = std::make_shared<CameraPubState>(); return p;
}
template <class T>
CameraSubscriber ImageTransport::subscribeCamera(const std::string&, uint32_t,
                                                 void (T::*)(const sensor_msgs::ImageConstPtr&,
                                                             const sensor_msgs::CameraInfoConstPtr&), T*) {
  return CameraSubscriber();
}
inline std::string getCameraInfoTopic(const std::string& base) { return base + "/camera_info"; }
}  // namespace image_transport
namespace message_filters {
template <class M> class Subscriber {
 public:
  void subscribe(ros::NodeHandle&, const std::string&, uint32_t) {}
};
template <class A, class B, class C, class D> class TimeSynchronizer {
 public:
  template <class F0, class F1, class F2, class F3> TimeSynchronizer(F0&, F1&, F2&, F3&, uint32_t) {}
  template <class MemFn, class T> void registerCallback(MemFn, T*) {}
};
}  // namespace message_filters

// ------------------------------------------------------------------------------------------------- image_geometry
namespace image_geometry {
// vision_opencv image_geometry/src/pinhole_camera_model.cpp.  fromCameraInfo keeps P as sent when binning is 0/1 and
// the ROI is empty (the only case on this path); fx = P(0,0), fy = P(1,1), cx = P(0,2), cy = P(1,2), Tx = P(0,3),
// Ty = P(1,3).  projectPixelTo3dRay: ray = ((u - cx - Tx)/fx, (v - cy - Ty)/fy, 1).  project3dToPixel:
// uv = ((fx*X + Tx)/Z + cx, (fy*Y + Ty)/Z + cy).  All double.
class PinholeCameraModel {
 public:
  bool fromCameraInfo(const sensor_msgs::CameraInfo& msg) { for (int i = 0; i < 12; i++) P_[i] = msg.P[i]; init_ = true; return true; }
  bool fromCameraInfo(const sensor_msgs::CameraInfoConstPtr& msg) { return fromCameraInfo(*msg); }
  double fx() const { return P_[0]; }
  double fy() const { return P_[5]; }
  double cx() const { return P_[2]; }
  double cy() const { return P_[6]; }
  double Tx() const { return P_[3]; }
  double Ty() const { return P_[7]; }
  cv::Point3d projectPixelTo3dRay(const cv::Point2d& uv) const {
    cv::Point3d ray;
    ray.x = (uv.x - cx() - Tx()) / fx();
    ray.y = (uv.y - cy() - Ty()) / fy();
    ray.z = 1.0;
    return ray;
  }
  cv::Point2d project3dToPixel(const cv::Point3d& xyz) const {
    cv::Point2d uv;
    uv.x = (fx() * xyz.x + Tx()) / xyz.z + cx();
    uv.y = (fy() * xyz.y + Ty()) / xyz.z + cy();
    return uv;
  }
 private:
  double P_[12] = {0};
  bool init_ = false;
};
class StereoCameraModel {  // only used by initializeOdometer (upstream glue, never run here)
 public:
  bool fromCameraInfo(const sensor_msgs::CameraInfo& l, const sensor_msgs::CameraInfo& r) { left_.fromCameraInfo(l); right_.fromCameraInfo(r); return true; }
  const PinholeCameraModel& left() const { return left_; }
  double baseline() const { return -right_.Tx() / right_.fx(); }
 private:
  PinholeCameraModel left_, right_;
};
}  // namespace image_geometry

// ---------------------------------------------------------------------------------------------------------- Eigen
#define EIGEN_ALIGN16 __attribute__((aligned(16)))
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
namespace Eigen {
// Fixed-size float vectors.  norm() = sqrt(squaredNorm()); for a 3-vector (not vectorisable) Eigen's default
// redux sums left to right: (x*x + y*y) + z*z.
struct Vector3f {
  float v[3];
  Vector3f() {}
  Vector3f(float x, float y, float z) { v[0] = x; v[1] = y; v[2] = z; }
  explicit Vector3f(const float* p) { v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; }
  float norm() const { return std::sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]); }
  float operator()(int i) const { return v[i]; }
};
template <class T> struct Map;
template <> struct Map<Vector3f> {
  const float* p;
  explicit Map(const float* q) : p(q) {}
  float norm() const { return std::sqrt((p[0] * p[0] + p[1] * p[1]) + p[2] * p[2]); }
};
struct Vector4f {
  float v[4];
  float operator()(int i) const { return v[i]; }
  float& operator()(int i) { return v[i]; }
  Vector4f operator-(const Vector4f& o) const { Vector4f r; for (int i = 0; i < 4; i++) r.v[i] = v[i] - o.v[i]; return r; }
  Vector4f operator+(const Vector4f& o) const { Vector4f r; for (int i = 0; i < 4; i++) r.v[i] = v[i] + o.v[i]; return r; }
  Vector4f operator/(float s) const { Vector4f r; for (int i = 0; i < 4; i++) r.v[i] = v[i] / s; return r; }
};
struct Vector3d {
  double v[3];
  Vector3d(double x, double y, double z) { v[0] = x; v[1] = y; v[2] = z; }
  double x() const { return v[0]; }
  double y() const { return v[1]; }
  double z() const { return v[2]; }
};
// Transform<double,3,Isometry> * Vector3d = linear() * v + translation(); a 3x3 * 3x1 product in Eigen 3.3 evaluates
// every row as ((r0*x + r1*y) + r2*z), with or without SSE2 packets (column-axpy gives the same association).
struct Isometry3d {
  double R[3][3], t[3];
  Vector3d operator*(const Vector3d& p) const {
    return Vector3d(((R[0][0] * p.v[0] + R[0][1] * p.v[1]) + R[0][2] * p.v[2]) + t[0],
                    ((R[1][0] * p.v[0] + R[1][1] * p.v[1]) + R[1][2] * p.v[2]) + t[1],
                    ((R[2][0] * p.v[0] + R[2][1] * p.v[1]) + R[2][2] * p.v[2]) + t[2]);
  }
};
}  // namespace Eigen

// ------------------------------------------------------------------------------------------------------------ tf2
namespace tf2 {
class Vector3 {
 public:
  Vector3() {}
  Vector3(double x, double y, double z) { m[0] = x; m[1] = y; m[2] = z; }
  void setX(double x) { m[0] = x; }
  void setY(double y) { m[1] = y; }
  void setZ(double z) { m[2] = z; }
  double x() const { return m[0]; }
  double y() const { return m[1]; }
  double z() const { return m[2]; }
  double m[3] = {0, 0, 0};
};
// tf2 LinearMath (tf2/LinearMath/{Matrix3x3,Transform,Quaternion}.h), restated as far as estimateCameraMotion /
// integrateAndBroadcastTF use it.  This is odometry glue AROUND the hot path: the reference node and the product node
// both call these very functions, so a test that drives both through stereoCallback compares how they USE tf2, not tf2.
//   Transform::inverse()      = (basis^T, basis^T * -origin)
//   Transform * Transform     = (b1 * b2, b1 * o2 + o1);  operator*= likewise
//   Matrix3x3::getRotation(q) = the trace / largest-diagonal branch of Matrix3x3.h
//   toMsg(Transform)          = translation <- origin, rotation <- getRotation()
class Matrix3x3 {
 public:
  Matrix3x3() {}
  Matrix3x3(double a, double b, double c, double d, double e, double f, double g, double h, double i) {
    m[0] = a; m[1] = b; m[2] = c; m[3] = d; m[4] = e; m[5] = f; m[6] = g; m[7] = h; m[8] = i;
  }
  double at(int r, int c) const { return m[3 * r + c]; }
  Matrix3x3 transpose() const { return Matrix3x3(m[0], m[3], m[6], m[1], m[4], m[7], m[2], m[5], m[8]); }
  Matrix3x3 operator*(const Matrix3x3& o) const {
    Matrix3x3 r;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) r.m[3 * i + j] = at(i, 0) * o.at(0, j) + at(i, 1) * o.at(1, j) + at(i, 2) * o.at(2, j);
    return r;
  }
  Vector3 operator*(const Vector3& v) const {
    return Vector3(m[0] * v.x() + m[1] * v.y() + m[2] * v.z(), m[3] * v.x() + m[4] * v.y() + m[5] * v.z(),
                   m[6] * v.x() + m[7] * v.y() + m[8] * v.z());
  }
  void getRotation(double q[4]) const {  // q = x, y, z, w
    const double trace = m[0] + m[4] + m[8];
    if (trace > 0.0) {
      double s = std::sqrt(trace + 1.0);
      q[3] = s * 0.5;
      s = 0.5 / s;
      q[0] = (at(2, 1) - at(1, 2)) * s;
      q[1] = (at(0, 2) - at(2, 0)) * s;
      q[2] = (at(1, 0) - at(0, 1)) * s;
    } else {
      const int i = at(0, 0) < at(1, 1) ? (at(1, 1) < at(2, 2) ? 2 : 1) : (at(0, 0) < at(2, 2) ? 2 : 0);
      const int j = (i + 1) % 3, k = (i + 2) % 3;
      double s = std::sqrt(at(i, i) - at(j, j) - at(k, k) + 1.0);
      q[i] = s * 0.5;
      s = 0.5 / s;
      q[3] = (at(k, j) - at(j, k)) * s;
      q[j] = (at(j, i) + at(i, j)) * s;
      q[k] = (at(k, i) + at(i, k)) * s;
    }
  }
  double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
};
class Transform {
 public:
  Transform() {}
  Transform(const Matrix3x3& b, const Vector3& o) : basis(b), origin(o) {}
  void setIdentity() { basis = Matrix3x3(); origin = Vector3(); }
  Transform inverse() const {
    const Matrix3x3 inv = basis.transpose();
    return Transform(inv, inv * Vector3(-origin.x(), -origin.y(), -origin.z()));
  }
  Transform operator*(const Transform& t) const {
    const Vector3 o = basis * t.origin;
    return Transform(basis * t.basis, Vector3(o.x() + origin.x(), o.y() + origin.y(), o.z() + origin.z()));
  }
  Transform& operator*=(const Transform& t) { *this = *this * t; return *this; }
  Matrix3x3 basis;
  Vector3 origin;
};
template <class T> class Stamped : public T {
 public:
  Stamped() {}
  Stamped(const T& v, const ros::Time& t, const std::string& f) : T(v), stamp_(t), frame_id_(f) {}
  ros::Time stamp_;
  std::string frame_id_;
};
inline geometry_msgs::Transform toMsg(const Transform& t) {
  geometry_msgs::Transform m;
  m.translation.x = t.origin.x(); m.translation.y = t.origin.y(); m.translation.z = t.origin.z();
  double q[4];
  t.basis.getRotation(q);
  m.rotation.x = q[0]; m.rotation.y = q[1]; m.rotation.z = q[2]; m.rotation.w = q[3];
  return m;
}
inline geometry_msgs::TransformStamped toMsg(const Stamped<Transform>& t) {
  geometry_msgs::TransformStamped m;
  m.header.stamp = t.stamp_;
  m.header.frame_id = t.frame_id_;
  m.transform = toMsg(static_cast<const Transform&>(t));
  return m;
}
inline void fromMsg(const geometry_msgs::TransformStamped&, Stamped<Transform>& out) { out.setIdentity(); }  // never reached: the stand-in tf buffer knows no frames

// tf2_eigen: transformToEigen(geometry_msgs::Transform) =
//   Isometry3d(Translation3d(t.x, t.y, t.z) * Quaterniond(q.w, q.x, q.y, q.z))
// i.e. linear = Quaterniond::toRotationMatrix() (Eigen/src/Geometry/Quaternion.h; the quaternion is NOT normalised),
// translation = t.
inline Eigen::Isometry3d transformToEigen(const geometry_msgs::Transform& m) {
  const double x = m.rotation.x, y = m.rotation.y, z = m.rotation.z, w = m.rotation.w;
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  Eigen::Isometry3d T;
  T.R[0][0] = 1.0 - (tyy + tzz); T.R[0][1] = txy - twz;         T.R[0][2] = txz + twy;
  T.R[1][0] = txy + twz;         T.R[1][1] = 1.0 - (txx + tzz); T.R[1][2] = tyz - twx;
  T.R[2][0] = txz - twy;         T.R[2][1] = tyz + twx;         T.R[2][2] = 1.0 - (txx + tyy);
  T.t[0] = m.translation.x; T.t[1] = m.translation.y; T.t[2] = m.translation.z;
  return T;
}
}  // namespace tf2

namespace tf2_ros {
class Buffer {
 public:
  bool canTransform(const std::string&, const std::string&, const ros::Time&, std::string* = nullptr) const { return false; }
  geometry_msgs::TransformStamped lookupTransform(const std::string&, const std::string&, const ros::Time&) const { return geometry_msgs::TransformStamped(); }
};
class TransformListener { public: explicit TransformListener(Buffer&) {} };
class TransformBroadcaster {
 public:
  void sendTransform(const geometry_msgs::TransformStamped& m) { if (ros::PubLog* log = ros::mod_ref_pub_log()) log->add("/tf", m); }
};
}  // namespace tf2_ros

// ------------------------------------------------------------------------- upstream estimators (scripted stand-ins)
// libviso2, sgm_gpu and pwc_net are inputs of this path.  Their stand-ins replay a script a test installs: frame k of
// the script holds what each estimator returns on the k-th stereo frame (and whether it fails).  Every estimator
// instance counts its own calls, so two nodes driven one after the other read the same script from the start.
// Without a script they fail, as before.
namespace mod_ref_upstream {
struct Frame {
  bool disparity_ok = false;
  std::vector<float> disparity;            // width x height, dense
  float f = 0, T = 0, min_disparity = 0, max_disparity = 0;
  bool motion_ok = false;
  double motion[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};  // libviso2 getMotion(): previous -> current
  bool flow_ok = false;
  std::vector<float> flow;                 // width x height x 2: optical flow of the previous frame -> this frame
};
struct Script { int width = 0, height = 0; std::vector<Frame> frames; };
inline const Script*& script() { static const Script* s = nullptr; return s; }
}  // namespace mod_ref_upstream

struct Matrix { double val[4][4]; };  // libviso2 Matrix
class Matcher {
 public:
  struct parameters { int nms_n = 0, nms_tau = 0, match_binsize = 0, match_radius = 0, match_disp_tolerance = 0, outlier_disp_tolerance = 0, outlier_flow_tolerance = 0, multi_stage = 0, half_resolution = 0, refinement = 0; };
};
class VisualOdometry {
 public:
  struct calibration { double f = 1, cu = 0, cv = 0; };
  struct bucketing { int max_features = 0; double bucket_width = 0, bucket_height = 0; };
  struct parameters { Matcher::parameters match; bucketing bucket; calibration calib; };
};
class VisualOdometryStereo : public VisualOdometry {
 public:
  struct parameters : public VisualOdometry::parameters { double base = 1; int ransac_iters = 0; double inlier_threshold = 0; bool reweighting = true; };
  explicit VisualOdometryStereo(parameters p) : param(p) {}
  bool process(uint8_t* left, uint8_t* right, int32_t* dims, bool = false) {
    const mod_ref_upstream::Script* s = mod_ref_upstream::script();
    if (!s || calls_ >= s->frames.size()) return false;
    last_ = &s->frames[calls_++];
    // the images must be what the node was given: the script's frames are mono8 images whose first byte is the frame index
    if (!left || !right || dims[0] != s->width || dims[1] != s->height || left[0] != (uint8_t)(calls_ - 1)) return false;
    return last_->motion_ok;
  }
  Matrix getMotion() {
    Matrix m;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) m.val[i][j] = last_ ? last_->motion[i][j] : (i == j);
    return m;
  }
  parameters param;
 private:
  size_t calls_ = 0;
  const mod_ref_upstream::Frame* last_ = nullptr;
};
namespace sgm_gpu {
class SgmGpu {
 public:
  explicit SgmGpu(const ros::NodeHandle&) {}
  // sgm_gpu stamps the DisparityImage (and its image) with the left image's header
  bool computeDisparity(const sensor_msgs::Image& left, const sensor_msgs::Image&, const sensor_msgs::CameraInfo&, const sensor_msgs::CameraInfo&, stereo_msgs::DisparityImage& out) {
    const mod_ref_upstream::Script* s = mod_ref_upstream::script();
    if (!s || calls_ >= s->frames.size()) return false;
    const mod_ref_upstream::Frame& fr = s->frames[calls_++];
    if (!fr.disparity_ok) return false;
    out.header = left.header;
    out.image.header = left.header;
    out.image.width = (uint32_t)s->width; out.image.height = (uint32_t)s->height;
    out.image.encoding = "32FC1"; out.image.is_bigendian = 0; out.image.step = (uint32_t)s->width * 4;
    out.image.data.resize((size_t)out.image.step * s->height);
    std::memcpy(out.image.data.data(), fr.disparity.data(), out.image.data.size());
    out.f = fr.f; out.T = fr.T; out.min_disparity = fr.min_disparity; out.max_disparity = fr.max_disparity; out.delta_d = 1.0f;
    return true;
  }
 private:
  size_t calls_ = 0;
};
}  // namespace sgm_gpu
namespace pwc_net {
class PwcNet {
 public:
  // first call = flow from frame 0 to frame 1 (the node has no previous image on frame 0)
  bool estimateOpticalFlow(const sensor_msgs::Image&, const sensor_msgs::Image&, cv::Mat& flow) {
    const mod_ref_upstream::Script* s = mod_ref_upstream::script();
    if (!s || calls_ + 1 >= s->frames.size()) return false;
    const mod_ref_upstream::Frame& fr = s->frames[++calls_];
    if (!fr.flow_ok) return false;
    flow.create(s->height, s->width, CV_32FC2);
    for (int r = 0; r < s->height; r++)
      std::memcpy(flow.data + (size_t)r * flow.step, fr.flow.data() + (size_t)r * s->width * 2, (size_t)s->width * 8);
    return true;
  }
 private:
  size_t calls_ = 0;
};
}  // namespace pwc_net

// ------------------------------------------------------------------------------------------------------------ PCL
#define PCL_ADD_POINT4D \
  union EIGEN_ALIGN16 { float data[4]; struct { float x; float y; float z; }; };
namespace pcl {
struct PCLHeader { uint32_t seq = 0; uint64_t stamp = 0; std::string frame_id; };
// pcl::PointXYZ (point_types.hpp): 16 bytes, default (0,0,0) with data[3] = 1
struct EIGEN_ALIGN16 PointXYZ {
  PCL_ADD_POINT4D
  PointXYZ() { x = y = z = 0.0f; data[3] = 1.0f; }
  PointXYZ(float x_, float y_, float z_) { x = x_; y = y_; z = z_; data[3] = 1.0f; }
};
struct PointIndices { PCLHeader header; std::vector<int> indices; };
typedef std::vector<PointIndices> IndicesClusters;

struct FieldDesc { const char* name; size_t offset; uint8_t datatype; };
namespace traits { template <class PointT> struct FieldList; }

// pcl::PointCloud<PointT> (point_cloud.h, 1.8): at(col,row) = points.at(row*width + col) for organised clouds (the
// only check is std::vector::at on the LINEAR index), throws for height <= 1; (w,h,value) ctor leaves is_dense = true;
// (cloud, indices) ctor copies the listed points into an unorganised cloud and inherits is_dense.
template <class PointT> class PointCloud {
 public:
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
  typedef typename std::vector<PointT>::iterator iterator;
  typedef typename std::vector<PointT>::const_iterator const_iterator;
  PCLHeader header;
  std::vector<PointT> points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;
  PointCloud() {}
  PointCloud(uint32_t w, uint32_t h, const PointT& value = PointT()) : points((size_t)w * h, value), width(w), height(h), is_dense(true) {}
  PointCloud(const PointCloud<PointT>& pc, const std::vector<int>& indices)
      : header(pc.header), points(indices.size()), width((uint32_t)indices.size()), height(1), is_dense(pc.is_dense) {
    for (size_t i = 0; i < indices.size(); i++) points[i] = pc.points[indices[i]];
  }
  const PointT& at(int column, int row) const {
    if (height > 1) return points.at((size_t)row * width + column);
    throw std::runtime_error("Can't use 2D indexing with a unorganized point cloud");
  }
  PointT& at(int column, int row) {
    if (height > 1) return points.at((size_t)row * width + column);
    throw std::runtime_error("Can't use 2D indexing with a unorganized point cloud");
  }
  const PointT& at(size_t n) const { return points.at(n); }
  PointT& at(size_t n) { return points.at(n); }
  size_t size() const { return points.size(); }
  iterator begin() { return points.begin(); }
  iterator end() { return points.end(); }
  const_iterator begin() const { return points.begin(); }
  const_iterator end() const { return points.end(); }
};

// pcl::getMinMax3D(cloud, min, max) (common/impl/common.hpp, 1.8): Array4f min/max over data[0..3] starting from
// +-FLT_MAX; dense clouds take every point, non-dense clouds skip points with a non-finite x, y or z.  Eigen's
// packet min/max are SSE minps/maxps: min(a,b) = a < b ? a : b with a = the running extreme.
template <class PointT> void getMinMax3D(const PointCloud<PointT>& cloud, Eigen::Vector4f& min_pt, Eigen::Vector4f& max_pt) {
  float mn[4], mx[4];
  for (int k = 0; k < 4; k++) { mn[k] = FLT_MAX; mx[k] = -FLT_MAX; }
  for (size_t i = 0; i < cloud.points.size(); i++) {
    const float* d = cloud.points[i].data;
    if (!cloud.is_dense && (!std::isfinite(d[0]) || !std::isfinite(d[1]) || !std::isfinite(d[2]))) continue;
    for (int k = 0; k < 4; k++) {
      mn[k] = mn[k] < d[k] ? mn[k] : d[k];
      mx[k] = mx[k] > d[k] ? mx[k] : d[k];
    }
  }
  for (int k = 0; k < 4; k++) { min_pt(k) = mn[k]; max_pt(k) = mx[k]; }
}

// pcl::toROSMsg / fromROSMsg (pcl_conversions + conversions.h): organised width/height, one PointField per registered
// field (count 1), point_step = sizeof(PointT), row_step = point_step*width, data = the points array byte for byte;
// the way back copies each registered field it finds by name into value-initialised points.
template <class PointT> void toROSMsg(const PointCloud<PointT>& cloud, sensor_msgs::PointCloud2& msg) {
  msg.width = cloud.width; msg.height = cloud.height;
  if (cloud.width == 0 && cloud.height == 0) { msg.width = (uint32_t)cloud.points.size(); msg.height = 1; }
  msg.fields.clear();
  for (const FieldDesc& f : traits::FieldList<PointT>::get()) {
    sensor_msgs::PointField pf; pf.name = f.name; pf.offset = (uint32_t)f.offset; pf.datatype = f.datatype; pf.count = 1;
    msg.fields.push_back(pf);
  }
  msg.is_bigendian = 0;
  msg.point_step = sizeof(PointT);
  msg.row_step = (uint32_t)(sizeof(PointT) * msg.width);
  msg.is_dense = cloud.is_dense;
  msg.data.resize(sizeof(PointT) * cloud.points.size());
  if (!cloud.points.empty()) std::memcpy(msg.data.data(), cloud.points.data(), msg.data.size());
}
template <class PointT> void fromROSMsg(const sensor_msgs::PointCloud2& msg, PointCloud<PointT>& cloud) {
  cloud.width = msg.width; cloud.height = msg.height; cloud.is_dense = msg.is_dense != 0;
  const size_t n = (size_t)msg.width * msg.height;
  cloud.points.assign(n, PointT());
  for (const FieldDesc& f : traits::FieldList<PointT>::get()) {
    const sensor_msgs::PointField* src = nullptr;
    for (const sensor_msgs::PointField& pf : msg.fields)
      if (pf.name == f.name && pf.datatype == f.datatype) src = &pf;
    if (!src) continue;  // PCL warns "Failed to find match for field" and leaves the member untouched
    for (uint32_t r = 0; r < msg.height; r++)
      for (uint32_t c = 0; c < msg.width; c++)
        std::memcpy((char*)&cloud.points[(size_t)r * msg.width + c] + f.offset,
                    &msg.data[(size_t)r * msg.row_step + (size_t)c * msg.point_step + src->offset], 4);
  }
}
}  // namespace pcl

// POINT_CLOUD_REGISTER_POINT_STRUCT(name, (type, member, tag)(type, member, tag)...) without Boost.PP: the sequence is
// walked by two macros that call each other.  Expands inside namespace pcl like the original.
#define MOD_REF_PCL_TYPE_float ::sensor_msgs::PointField::FLOAT32
#define MOD_REF_FIELD(pt, type, member) {#member, offsetof(pt, member), MOD_REF_PCL_TYPE_##type},
#define MOD_REF_SEQ_A(type, member, tag) MOD_REF_FIELD(MOD_REF_POINT_T, type, member) MOD_REF_SEQ_B
#define MOD_REF_SEQ_B(type, member, tag) MOD_REF_FIELD(MOD_REF_POINT_T, type, member) MOD_REF_SEQ_A
#define MOD_REF_SEQ_A_END
#define MOD_REF_SEQ_B_END
#define MOD_REF_SEQ_END_(...) __VA_ARGS__##_END
#define MOD_REF_SEQ_END(...) MOD_REF_SEQ_END_(__VA_ARGS__)
#define POINT_CLOUD_REGISTER_POINT_STRUCT(name, seq)                                             \
  namespace pcl { namespace traits {                                                             \
  template <> struct FieldList<name> {                                                           \
    typedef name MOD_REF_POINT_T;                                                                \
    static std::vector<FieldDesc> get() { return {MOD_REF_SEQ_END(MOD_REF_SEQ_A seq)}; }             \
  };                                                                                             \
  } }

#endif  // MOD_REF_SHIM_H_
