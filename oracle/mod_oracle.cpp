// mod_oracle.cpp -- CPU ORACLE: a dependency-free restatement of the reference's hot path.
//
// TEST INFRASTRUCTURE ONLY (see mod_oracle.h).  PARITY STATUS: pinned to the reference's own code with shimmed
// dependencies -- tests/test_reference_build.py requires every output below to equal, bit for bit, what the
// reference's unmodified translation units publish when compiled against oracle/ref_shim/ (oracle/_ref/libmod_ref.so).
// Every function cites the reference lines it restates (paths relative to /root/reference) so it can also be
// reviewed line-against-line.
//
// Build: g++ -std=c++17 -O3 -DNDEBUG -msse3 -ffp-contract=off  (catkin Release + add_definitions(-msse3),
// scene_flow_constructor/CMakeLists.txt:4).  x86-64 without -mfma never contracts a*b+c, and
// -ffp-contract=off makes that explicit, so every float / double operation below rounds exactly once,
// which is what the CUDA path reproduces with __fmul_rn/__dadd_rn/... intrinsics.
//
// The structure (five full-image passes with AoS 16-B / 32-B points, column-major walks where the
// reference walks column-major, fresh allocations per frame, bounds-checked .at(), std::vector<bool>
// dynamic map, per-cluster copy + std::sort) is kept on purpose: this file is also the timed CPU baseline.

#include "mod_oracle.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------------
// Point records
// ---------------------------------------------------------------------------------------------------
struct alignas(16) PointXYZ {  // pcl::PointXYZ: float data[4], 16 B
  float x, y, z, w;
};
struct alignas(16) PointXYZVelocity {  // scene_flow_constructor/pcl_point_xyz_velocity.h:8-24, 32 B
  float x, y, z, pad0;
  float vx, vy, vz, pad1;
};
static_assert(sizeof(PointXYZ) == 16, "PointXYZ must be 16 bytes");
static_assert(sizeof(PointXYZVelocity) == 32, "PointXYZVelocity must be 32 bytes");

// pcl::PointCloud<T>(w, h, value): organized, row-major storage; at(col,row) = points.at(row*width+col)
// (PCL 1.8, bounds check on the linear index only).
template <typename T>
struct OrganizedCloud {
  int width = 0, height = 0;
  std::vector<T> points;
  OrganizedCloud() = default;
  OrganizedCloud(int w, int h, const T& v) : width(w), height(h), points(static_cast<size_t>(w) * h, v) {}
  OrganizedCloud(int w, int h) : width(w), height(h), points(static_cast<size_t>(w) * h) {}
  T& at(int col, int row) { return points.at(static_cast<size_t>(row) * width + col); }
  const T& at(int col, int row) const { return points.at(static_cast<size_t>(row) * width + col); }
  size_t size() const { return points.size(); }
};

// x86 cvttss2si: NaN / out-of-range -> 0x80000000.  The reference relies on this ("previous.x = std::round(..)"
// with an int destination, scene_flow_constructor.h:210-211); stated explicitly here instead of being UB.
inline int cvtFloatToIntX86(float r) {
  if (!(r >= -2147483648.0f && r < 2147483648.0f)) return INT_MIN;
  return static_cast<int>(r);
}

// ---------------------------------------------------------------------------------------------------
// disparity_image_proc/src/disparity_image_processor.cpp
// ---------------------------------------------------------------------------------------------------
class DisparityImage {
 public:
  DisparityImage(const orc_disparity& d, int w, int h, const orc_camera& cam)
      : msg_(d), width_(w), height_(h), cam_(cam) {}

  int width() const { return width_; }
  int height() const { return height_; }

  // cv::Mat_<float>(rows, cols, ptr, step).at<float>(v,u)   (.cpp:5,12)
  float texel(int u, int v) const {
    const char* row = reinterpret_cast<const char*>(msg_.data) + static_cast<ptrdiff_t>(v) * msg_.step_bytes;
    return reinterpret_cast<const float*>(row)[u];
  }

  // getDisparity, .cpp:17-31 : bounds, then the two range tests (a NaN passes both)
  bool disparityAt(int u, int v, float& disparity) const {
    if (u < 0 || u >= width_) return false;
    if (v < 0 || v >= height_) return false;
    disparity = texel(u, v);
    if (msg_.max_disparity < disparity)
      return false;
    else if (msg_.min_disparity > disparity)
      return false;
    return true;
  }

  // getPoint3D, .cpp:33-50.  z is float (f*T in float, then a float divide); the ray is
  // image_geometry::PinholeCameraModel::projectPixelTo3dRay in double;
  // x,y = double * float -> double, narrowed on store.
  bool pointAt(int u, int v, PointXYZ& p) const {
    float disparity;
    if (!disparityAt(u, v, disparity)) return false;
    if (disparity == 0.0) return false;
    const float focal_length = msg_.f;
    const float baseline = msg_.T;
    p.z = focal_length * baseline / disparity;
    const double ray_x = (static_cast<double>(u) - cam_.cx - cam_.Tx) / cam_.fx;
    const double ray_y = (static_cast<double>(v) - cam_.cy - cam_.Ty) / cam_.fy;
    p.x = static_cast<float>(ray_x * p.z);
    p.y = static_cast<float>(ray_y * p.z);
    return true;
  }

  // toPointCloud, .cpp:86-103: NaN default, u outer / v inner
  void toPointCloud(OrganizedCloud<PointXYZ>& cloud) const {
    const float nan = std::nanf("");
    PointXYZ default_value{nan, nan, nan, 1.0f};
    cloud = OrganizedCloud<PointXYZ>(width_, height_, default_value);
    for (int u = 0; u < width_; u++) {
      for (int v = 0; v < height_; v++) {
        PointXYZ p{0, 0, 0, 1.0f};
        if (pointAt(u, v, p)) cloud.at(u, v) = p;
      }
    }
  }

  // toDepthImage, .cpp:105-121
  void toDepthImage(float* depth) const {
    const float nan = std::nanf("");
    for (int u = 0; u < width_; u++) {
      for (int v = 0; v < height_; v++) {
        PointXYZ p{0, 0, 0, 1.0f};
        depth[static_cast<size_t>(v) * width_ + u] = pointAt(u, v, p) ? p.z : nan;
      }
    }
  }

 private:
  orc_disparity msg_;
  int width_, height_;
  orc_camera cam_;
};

// ---------------------------------------------------------------------------------------------------
// tf2::transformToEigen + Eigen::Quaterniond::toRotationMatrix (Eigen 3.3)
// ---------------------------------------------------------------------------------------------------
struct Isometry {
  double R[9];  // row-major
  double t[3];
};

Isometry isometryFromTransform(const orc_transform& m) {
  Isometry iso;
  const double x = m.qx, y = m.qy, z = m.qz, w = m.qw;
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  iso.R[0] = 1.0 - (tyy + tzz);
  iso.R[1] = txy - twz;
  iso.R[2] = txz + twy;
  iso.R[3] = txy + twz;
  iso.R[4] = 1.0 - (txx + tzz);
  iso.R[5] = tyz - twx;
  iso.R[6] = txz - twy;
  iso.R[7] = tyz + twx;
  iso.R[8] = 1.0 - (txx + tyy);
  iso.t[0] = m.tx;
  iso.t[1] = m.ty;
  iso.t[2] = m.tz;
  return iso;
}

// Isometry3d * Vector3d: res = t; res += R*p  (Eigen/src/Geometry/Transform.h, affine x vector case).
// The order of the three-term sum inside Eigen is unpinned (packet vs scalar redux); the oracle fixes it
// as ((R0*px + R1*py) + R2*pz) + t and the CUDA kernel uses the same order.
inline void applyIsometry(const Isometry& T, double px, double py, double pz, double out[3]) {
  for (int r = 0; r < 3; r++) {
    const double s = (T.R[3 * r + 0] * px + T.R[3 * r + 1] * py) + T.R[3 * r + 2] * pz;
    out[r] = s + T.t[r];
  }
}

// ---------------------------------------------------------------------------------------------------
// scene_flow_constructor/src/scene_flow_constructor.cpp + include/scene_flow_constructor.h
// ---------------------------------------------------------------------------------------------------
class SceneFlowConstructor {
 public:
  SceneFlowConstructor(int w, int h, const orc_camera& cam, int dynamic_flow_diff)
      : image_width_(w), image_height_(h), cam_(cam), dynamic_flow_diff_(dynamic_flow_diff) {}

  // transformPCPreviousToNow, .cpp:409-429 (u outer / v inner; NaN x passes through untouched)
  void transformPCPreviousToNow(const OrganizedCloud<PointXYZ>& pc_previous, OrganizedCloud<PointXYZ>& out,
                                const orc_transform& previous_to_now) const {
    const Isometry T = isometryFromTransform(previous_to_now);
    out = OrganizedCloud<PointXYZ>(image_width_, image_height_);
    for (int u = 0; u < image_width_; u++) {
      for (int v = 0; v < image_height_; v++) {
        const PointXYZ& p = pc_previous.at(u, v);
        if (std::isnan(p.x)) {
          out.at(u, v) = pc_previous.at(u, v);
          continue;
        }
        double q[3];
        applyIsometry(T, p.x, p.y, p.z, q);
        out.at(u, v) = PointXYZ{static_cast<float>(q[0]), static_cast<float>(q[1]), static_cast<float>(q[2]), 1.0f};
      }
    }
  }

  // calculateStaticOpticalFlow, .cpp:65-89; project3dToPixel is image_geometry's
  // ((fx*X + Tx)/Z + cx, (fy*Y + Ty)/Z + cy) in double, no Z<=0 guard.
  void calculateStaticOpticalFlow(const OrganizedCloud<PointXYZ>& pc_previous_transformed,
                                  std::vector<float>& static_flow) const {
    static_flow.assign(static_cast<size_t>(image_width_) * image_height_ * 2, 0.0f);
    for (int y = 0; y < image_height_; y++) {
      for (int x = 0; x < image_width_; x++) {
        const PointXYZ p = pc_previous_transformed.at(x, y);
        float* dst = &static_flow[(static_cast<size_t>(y) * image_width_ + x) * 2];
        if (std::isnan(p.x)) {
          dst[0] = std::nanf("");
          dst[1] = std::nanf("");
          continue;
        }
        const double X = p.x, Y = p.y, Z = p.z;
        const double px = (cam_.fx * X + cam_.Tx) / Z + cam_.cx;
        const double py = (cam_.fy * Y + cam_.Ty) / Z + cam_.cy;
        dst[0] = static_cast<float>(px - x);
        dst[1] = static_cast<float>(py - y);
      }
    }
  }

  // isValid, scene_flow_constructor.h:240-249 (only x is inspected)
  static bool isValid(const PointXYZ& p) {
    if (std::isnan(p.x)) return false;
    if (std::isinf(p.x)) return false;
    return true;
  }

  const float* flowAt(const float* flow, int flow_step_bytes, int x, int y) const {
    const char* row = reinterpret_cast<const char*>(flow) + static_cast<ptrdiff_t>(y) * flow_step_bytes;
    return reinterpret_cast<const float*>(row) + 2 * x;
  }

  // getPreviousPoint, scene_flow_constructor.h:195-214
  bool getPreviousPoint(int now_x, int now_y, int& prev_x, int& prev_y, const float* flow, int flow_step) const {
    if (now_x < 0 || now_x >= image_width_ || now_y < 0 || now_y >= image_height_) return false;
    const float* f = flowAt(flow, flow_step, now_x, now_y);
    if (std::isnan(f[0]) || std::isnan(f[1])) return false;
    prev_x = cvtFloatToIntX86(std::round(now_x - f[0]));  // int -> float, float subtract, roundf
    prev_y = cvtFloatToIntX86(std::round(now_y - f[1]));
    return true;
  }

  // getRightPoint, scene_flow_constructor.h:215-227 (the computed right pixel is never used downstream)
  static bool getRightPoint(int left_x, int left_y, const DisparityImage& disparity_image) {
    float disparity;
    if (!disparity_image.disparityAt(left_x, left_y, disparity)) return false;
    if (std::isnan(disparity) || std::isinf(disparity) || disparity < 0) return false;
    return true;
  }

  // getMatchPoints, scene_flow_constructor.h:173-194
  bool getMatchPoints(int now_x, int now_y, int& prev_x, int& prev_y, const float* flow, int flow_step,
                      const DisparityImage& disparity_now, const DisparityImage& disparity_previous) const {
    if (!getPreviousPoint(now_x, now_y, prev_x, prev_y, flow, flow_step)) return false;
    if (!getRightPoint(now_x, now_y, disparity_now)) return false;
    if (!getRightPoint(prev_x, prev_y, disparity_previous)) return false;
    return true;
  }

  // constructVelocityPC, .cpp:149-212 (+ initializeVelocityPC, .cpp:293-303)
  void constructVelocityPC(const OrganizedCloud<PointXYZ>& pc_now,
                           const OrganizedCloud<PointXYZ>& pc_previous_transformed, const float* flow,
                           int flow_step, const std::vector<float>& static_flow, const DisparityImage& disparity_now,
                           const DisparityImage& disparity_previous, double dt_sec,
                           OrganizedCloud<PointXYZVelocity>& velocity_pc, uint8_t* opt_class,
                           uint8_t* opt_near) const {
    const float nan = std::nanf("");
    PointXYZVelocity default_value{nan, nan, nan, 0.0f, nan, nan, nan, 0.0f};
    velocity_pc = OrganizedCloud<PointXYZVelocity>(image_width_, image_height_, default_value);
    const float threshold = static_cast<float>(dynamic_flow_diff_);  // float >= int promotes the int to float

    for (int y = 0; y < image_height_; y++) {
      for (int x = 0; x < image_width_; x++) {
        const size_t idx = static_cast<size_t>(y) * image_width_ + x;
        if (opt_class) opt_class[idx] = ORC_CLASS_NAN;
        if (opt_near) opt_near[idx] = 0;
        PointXYZVelocity& out = velocity_pc.at(x, y);
        const PointXYZ now = pc_now.at(x, y);
        if (!isValid(now)) continue;
        out.x = now.x;
        out.y = now.y;
        out.z = now.z;
        if (opt_class) opt_class[idx] = ORC_CLASS_XYZ_ONLY;

        int prev_x, prev_y;
        if (!getMatchPoints(x, y, prev_x, prev_y, flow, flow_step, disparity_now, disparity_previous)) continue;

        const PointXYZ previous = pc_previous_transformed.at(prev_x, prev_y);
        if (!isValid(previous)) continue;

        const float* f = flowAt(flow, flow_step, x, y);
        const float sx = static_flow[idx * 2 + 0];
        const float sy = static_flow[idx * 2 + 1];
        if (std::isnan(sx)) continue;

        // cv::Vec2f - cv::Vec2f, then Matx::dot: s = 0; s += a0*a0; s += a1*a1 (float), then std::sqrt(float)
        const float ex = f[0] - sx;
        const float ey = f[1] - sy;
        float dot = 0.0f;
        dot += ex * ex;
        dot += ey * ey;
        const float residual = std::sqrt(dot);
        if (opt_near &&
            residual >= std::nextafterf(threshold, -INFINITY) && residual <= std::nextafterf(threshold, INFINITY))
          opt_near[idx] |= ORC_NEAR_FLOW_DIFF;

        if (residual >= threshold) {
          out.vx = static_cast<float>((now.x - previous.x) / dt_sec);  // float diff, double divide, narrow
          out.vy = static_cast<float>((now.y - previous.y) / dt_sec);
          out.vz = static_cast<float>((now.z - previous.z) / dt_sec);
          if (opt_class) opt_class[idx] = ORC_CLASS_MOVING;
        } else {
          out.vx = 0.0f;
          out.vy = 0.0f;
          out.vz = 0.0f;
          if (opt_class) opt_class[idx] = ORC_CLASS_STATIC;
        }
      }
    }
  }

  // construct, .cpp:91-147 with every publisher "subscribed"
  int construct(const DisparityImage& disparity_now, const DisparityImage& disparity_previous, const float* flow,
                int flow_step, const orc_transform& prev2now, double dt_sec, float* out_points32, float* opt_depth,
                float* opt_static_flow, uint8_t* opt_class, uint8_t* opt_near) const {
    OrganizedCloud<PointXYZ> pc_previous, pc_now;
    disparity_previous.toPointCloud(pc_previous);
    disparity_now.toPointCloud(pc_now);
    if (opt_depth) disparity_now.toDepthImage(opt_depth);

    OrganizedCloud<PointXYZ> pc_previous_transformed;
    transformPCPreviousToNow(pc_previous, pc_previous_transformed, prev2now);

    std::vector<float> static_flow;
    calculateStaticOpticalFlow(pc_previous_transformed, static_flow);

    OrganizedCloud<PointXYZVelocity> pc_with_velocity;
    constructVelocityPC(pc_now, pc_previous_transformed, flow, flow_step, static_flow, disparity_now,
                        disparity_previous, dt_sec, pc_with_velocity, opt_class, opt_near);

    // publishPointcloud -> pcl::toROSMsg = memcpy of the 32-byte records (.cpp:351-362)
    if (out_points32) std::memcpy(out_points32, pc_with_velocity.points.data(), pc_with_velocity.size() * 32);
    if (opt_static_flow) std::memcpy(opt_static_flow, static_flow.data(), static_flow.size() * sizeof(float));
    return 0;
  }

 private:
  int image_width_, image_height_;
  orc_camera cam_;
  int dynamic_flow_diff_;
};

// ---------------------------------------------------------------------------------------------------
// scene_flow_clusterer/include/lookup_table.h + src/lookup_table.cpp
// ---------------------------------------------------------------------------------------------------
class EquivalenceTable {
 public:
  void resize(size_t n) { table_.resize(n); }
  size_t size() const { return table_.size(); }
  void reset() { max_label_ = -1; }  // only rewinds the counter (lookup_table.cpp:11-14)
  int addLabel() {
    max_label_++;
    table_.at(max_label_) = max_label_;
    return max_label_;
  }
  int lookup(int source) {  // path halving
    while (source != table_.at(source)) {
      table_.at(source) = table_.at(table_.at(source));
      source = table_.at(source);
    }
    return source;
  }
  void link(int a, int b) {  // larger root points at smaller root
    const int ra = lookup(a);
    const int rb = lookup(b);
    if (ra > rb)
      table_.at(ra) = rb;
    else
      table_.at(rb) = ra;
  }

 private:
  std::vector<int> table_;
  int max_label_ = -1;
};

// Eigen::Vector3f::norm() = sqrtf(squaredNorm()); summation order unpinned in Eigen, fixed here and in the
// kernels as (x*x + y*y) + z*z without contraction.
inline float norm3(float x, float y, float z) { return std::sqrt((x * x + y * y) + z * z); }

// ---------------------------------------------------------------------------------------------------
// scene_flow_clusterer/src/clusterer_nodelet.cpp
// ---------------------------------------------------------------------------------------------------
class Clusterer {
 public:
  Clusterer(int cluster_size, double depth_diff, double dynamic_speed, int neighbor_distance)
      : cluster_size_th_(cluster_size),
        depth_diff_th_(depth_diff),
        dynamic_speed_th_(dynamic_speed),
        neighbor_distance_th_(neighbor_distance) {}

  static constexpr int NOT_BELONGED = -1;

  // dataCB, .cpp:221-242: fromROSMsg = memcpy into the AoS cloud
  void setInput(int w, int h, const float* points32) {
    input_ = OrganizedCloud<PointXYZVelocity>(w, h);
    std::memcpy(input_.points.data(), points32, input_.size() * 32);
  }

  // calculateDynamicMap, .cpp:40-54 (float norm compared with a double threshold; NaN -> false)
  void calculateDynamicMap(uint8_t* opt_near) {
    if (dynamic_map_.size() != input_.size()) dynamic_map_.resize(input_.size());
    std::fill(dynamic_map_.begin(), dynamic_map_.end(), false);
    for (size_t i = 0; i < input_.size(); i++) {
      const PointXYZVelocity& p = input_.points.at(i);
      const float speed = norm3(p.vx, p.vy, p.vz);
      if (speed >= dynamic_speed_th_) dynamic_map_.at(i) = true;
      if (opt_near) {
        const float thr = static_cast<float>(dynamic_speed_th_);
        opt_near[i] = (speed >= std::nextafterf(thr, -INFINITY) && speed <= std::nextafterf(thr, INFINITY))
                          ? ORC_NEAR_SPEED
                          : 0;
      }
    }
  }

  void initClusterMap() {  // .cpp:244-251
    number_of_clusters_ = 0;
    if (cluster_map_.size() != input_.size()) cluster_map_.resize(input_.size());
    std::fill(cluster_map_.begin(), cluster_map_.end(), NOT_BELONGED);
  }

  bool isInRange(int u, int v) const { return !(u < 0 || u >= input_.width || v < 0 || v >= input_.height); }
  bool isDynamic(int u, int v) const { return dynamic_map_.at(static_cast<size_t>(input_.width) * v + u); }
  int& clusterAt(int u, int v) { return cluster_map_.at(static_cast<size_t>(v) * input_.width + u); }
  float depthDiff(int u1, int v1, int u2, int v2) const {  // clusterer_nodelet.h:83-86 (float)
    return std::abs(input_.at(u1, v1).z - input_.at(u2, v2).z);
  }

  // comparePoints, .cpp:186-219
  void comparePoints(int iu, int iv, int cu, int cv) {
    if (!isInRange(cu, cv)) return;
    if (!isDynamic(cu, cv)) return;
    if (depthDiff(iu, iv, cu, cv) > depth_diff_th_) return;  // float promoted to double; NaN passes
    int& c1 = clusterAt(iu, iv);
    int& c2 = clusterAt(cu, cv);
    if (c1 == NOT_BELONGED && c2 == NOT_BELONGED) {
      const int fresh = lookup_table_.addLabel();
      c1 = fresh;
      c2 = fresh;
    } else if (c1 != NOT_BELONGED && c2 == NOT_BELONGED) {
      c2 = c1;
    } else if (c1 == NOT_BELONGED && c2 != NOT_BELONGED) {
      c1 = c2;
    } else {
      if (c1 != c2) lookup_table_.link(c1, c2);
    }
  }

  // calculateInitialClusterMap, .cpp:56-83: raster scan, upper-left (N+1)^2-1 window
  void calculateInitialClusterMap() {
    if (lookup_table_.size() != input_.size()) lookup_table_.resize(input_.size());
    lookup_table_.reset();
    for (int v = 0; v < input_.height; v++) {
      for (int u = 0; u < input_.width; u++) {
        if (!isDynamic(u, v)) continue;
        for (int dv = -neighbor_distance_th_; dv <= 0; dv++) {
          for (int du = -neighbor_distance_th_; du <= 0; du++) {
            if (dv == 0 && du == 0) continue;
            comparePoints(u, v, u + du, v + dv);
          }
        }
      }
    }
  }

  void integrateConnectedClusters() {  // .cpp:253-267
    number_of_clusters_ = 0;
    for (size_t i = 0; i < cluster_map_.size(); i++) {
      const int initial = cluster_map_.at(i);
      if (initial == NOT_BELONGED) continue;
      const int final_label = lookup_table_.lookup(initial);
      cluster_map_.at(i) = final_label;
      if (final_label > number_of_clusters_ - 1) number_of_clusters_ = final_label + 1;
    }
  }

  void removeSmallClusters() {  // .cpp:354-393 (size_t vs int comparison kept)
    if (number_of_clusters_ <= 0) return;
    std::vector<size_t> cluster_size(number_of_clusters_, 0);
    for (size_t i = 0; i < cluster_map_.size(); i++) {
      const int c = cluster_map_.at(i);
      if (c == NOT_BELONGED) continue;
      cluster_size.at(c) += 1;
    }
    std::vector<int> old2new(number_of_clusters_);
    for (size_t i = 0; i < cluster_size.size(); i++) {
      if (cluster_size.at(i) < static_cast<size_t>(cluster_size_th_)) {
        old2new.at(i) = NOT_BELONGED;
        number_of_clusters_--;
      } else {
        old2new.at(i) = static_cast<int>(i - (cluster_size.size() - number_of_clusters_));
      }
    }
    for (size_t i = 0; i < cluster_map_.size(); i++) {
      const int old_label = cluster_map_.at(i);
      if (old_label == NOT_BELONGED) continue;
      cluster_map_.at(i) = old2new.at(old_label);
    }
  }

  // clusterMap2IndicesCluster, .cpp:97-117 (u outer / v inner)
  void clusterMap2IndicesCluster(std::vector<std::vector<int>>& clusters) {
    if (number_of_clusters_ <= 0) return;
    clusters.resize(number_of_clusters_);
    for (int u = 0; u < input_.width; u++) {
      for (int v = 0; v < input_.height; v++) {
        const int c = clusterAt(u, v);
        if (c == NOT_BELONGED) continue;
        clusters.at(c).push_back(input_.width * v + u);
      }
    }
  }

  void clustering(std::vector<std::vector<int>>& clusters, uint8_t* opt_near) {  // .cpp:85-95
    calculateDynamicMap(opt_near);
    initClusterMap();
    calculateInitialClusterMap();
    integrateConnectedClusters();
    removeSmallClusters();
    clusterMap2IndicesCluster(clusters);
  }

  // cluster2MovingObject, .cpp:147-184.  Returns 1 accepted, 0 rejected, -3 if the reference would throw
  // (empty cluster -> cluster.at(0) out of range; only reachable with cluster_size == 0).
  int cluster2MovingObject(const std::vector<int>& indices, orc_object& obj) const {
    std::vector<PointXYZVelocity> cluster;  // pcl::PointCloud(cloud, indices) copies the indexed points
    cluster.reserve(indices.size());
    for (int i : indices) cluster.push_back(input_.points.at(i));

    // pcl::getMinMax3D on a dense cloud: running min/max from +-FLT_MAX
    float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(),
                   std::numeric_limits<float>::max()};
    float mx[3] = {-std::numeric_limits<float>::max(), -std::numeric_limits<float>::max(),
                   -std::numeric_limits<float>::max()};
    for (const PointXYZVelocity& p : cluster) {
      const float c[3] = {p.x, p.y, p.z};
      for (int k = 0; k < 3; k++) {
        mn[k] = (c[k] < mn[k]) ? c[k] : mn[k];
        mx[k] = (c[k] > mx[k]) ? c[k] : mx[k];
      }
    }
    for (int k = 0; k < 3; k++) {
      obj.bbox[k] = mx[k] - mn[k];             // Vector4f max - min
      obj.center[k] = (mn[k] + mx[k]) / 2;     // (min + max) / 2 in float
    }

    std::sort(cluster.begin(), cluster.end(), [](const PointXYZVelocity& a, const PointXYZVelocity& b) {
      return norm3(a.vx, a.vy, a.vz) > norm3(b.vx, b.vy, b.vz);
    });
    if (cluster.empty()) return -3;
    const PointXYZVelocity& median = cluster.at(cluster.size() / 2);
    if (norm3(median.vx, median.vy, median.vz) < dynamic_speed_th_) return 0;
    obj.velocity[0] = median.vx;
    obj.velocity[1] = median.vy;
    obj.velocity[2] = median.vz;
    obj.n_points = static_cast<int32_t>(cluster.size());
    return 1;
  }

  // publishMovingObjects, .cpp:324-343: ids are the running count of accepted clusters
  int collectMovingObjects(const std::vector<std::vector<int>>& clusters, orc_object* objs, int max_objs,
                           int* n_objs) const {
    int id = 0;
    int status = 0;
    for (const auto& indices : clusters) {
      orc_object obj{};
      const int r = cluster2MovingObject(indices, obj);
      if (r < 0) return r;
      if (r == 1) {
        obj.id = id;
        if (id < max_objs && objs)
          objs[id] = obj;
        else
          status = -2;
        id++;
      }
    }
    if (n_objs) *n_objs = id;
    return status;
  }

  const std::vector<bool>& dynamicMap() const { return dynamic_map_; }
  const std::vector<int>& clusterMap() const { return cluster_map_; }

 private:
  int cluster_size_th_;
  double depth_diff_th_;
  double dynamic_speed_th_;
  int neighbor_distance_th_;
  OrganizedCloud<PointXYZVelocity> input_;
  std::vector<int> cluster_map_;
  int number_of_clusters_ = 0;
  EquivalenceTable lookup_table_;
  std::vector<bool> dynamic_map_;
};

}  // namespace

// ===================================================================================================
// C interface
// ===================================================================================================
extern "C" {

int orc_scene_flow(int w, int h, const orc_camera* cam, const orc_disparity* now, const orc_disparity* prev,
                   const float* flow_xy, int32_t flow_step_bytes, const orc_transform* prev2now, double dt_sec,
                   int32_t dynamic_flow_diff, float* out_points32, float* opt_depth, float* opt_static_flow,
                   uint8_t* opt_class) {
  if (w <= 0 || h <= 0 || !cam || !now || !prev || !flow_xy || !prev2now || !out_points32) return -1;
  DisparityImage dnow(*now, w, h, *cam), dprev(*prev, w, h, *cam);
  SceneFlowConstructor sfc(w, h, *cam, dynamic_flow_diff);
  return sfc.construct(dnow, dprev, flow_xy, flow_step_bytes, *prev2now, dt_sec, out_points32, opt_depth,
                       opt_static_flow, opt_class, nullptr);
}

int orc_scene_flow_near(int w, int h, const orc_camera* cam, const orc_disparity* now, const orc_disparity* prev,
                        const float* flow_xy, int32_t flow_step_bytes, const orc_transform* prev2now,
                        int32_t dynamic_flow_diff, uint8_t* out_near) {
  if (w <= 0 || h <= 0 || !cam || !now || !prev || !flow_xy || !prev2now || !out_near) return -1;
  DisparityImage dnow(*now, w, h, *cam), dprev(*prev, w, h, *cam);
  SceneFlowConstructor sfc(w, h, *cam, dynamic_flow_diff);
  return sfc.construct(dnow, dprev, flow_xy, flow_step_bytes, *prev2now, 1.0, nullptr, nullptr, nullptr, nullptr,
                       out_near);
}

int orc_cluster(int w, int h, const float* points32, int32_t cluster_size, double depth_diff, double dynamic_speed,
                int32_t neighbor_distance, uint8_t* opt_dynamic_map, int32_t* opt_labels, uint8_t* opt_near,
                orc_object* objs, int32_t max_objs, int32_t* n_objs) {
  if (w <= 0 || h <= 0 || !points32) return -1;
  Clusterer cl(cluster_size, depth_diff, dynamic_speed, neighbor_distance);
  cl.setInput(w, h, points32);
  std::vector<std::vector<int>> clusters;
  cl.clustering(clusters, opt_near);
  const size_t n = static_cast<size_t>(w) * h;
  if (opt_dynamic_map)
    for (size_t i = 0; i < n; i++) opt_dynamic_map[i] = cl.dynamicMap()[i] ? 1 : 0;
  if (opt_labels) std::memcpy(opt_labels, cl.clusterMap().data(), n * sizeof(int32_t));
  return cl.collectMovingObjects(clusters, objs, max_objs, n_objs);
}

// ---- debug / visual side outputs of the clusterer ----------------------------------------------
// ClustererNodelet::publishClustersImage (clusterer_nodelet.cpp:293-322) + ColorSet (color_set.cpp:11-26).  The
// 256-entry table cv::applyColorMap(COLORMAP_HSV) maps through is OpenCV's (un-vendored); the caller passes it in
// (tests use the fixture generated from cv2, tests/golden/hsv_colormap_bgr.npy).
int orc_clusters_image(int w, int h, const int32_t* labels, int32_t number_of_clusters, const uint8_t* colormap_bgr768,
                       uint8_t* out_bgr8) {
  if (w <= 0 || h <= 0 || !labels || !colormap_bgr768 || !out_bgr8) return -1;
  // ColorSet::resize(set_size): returns early for set_size < 1; input_map[i] = i * 255 / set_size; applyColorMap
  std::vector<uint8_t> color_set;  // BGR per entry
  if (number_of_clusters >= 1) {
    color_set.resize(static_cast<size_t>(number_of_clusters) * 3);
    for (int i = 0; i < number_of_clusters; i++) {
      const uint8_t v = static_cast<uint8_t>(i * 255 / number_of_clusters);
      for (int k = 0; k < 3; k++) color_set[3 * static_cast<size_t>(i) + k] = colormap_bgr768[3 * v + k];
    }
  }
  const size_t n = static_cast<size_t>(w) * h;
  for (size_t i = 0; i < n; i++) {
    int b = 0, g = 0, r = 0;
    const int cluster = labels[i];
    if (cluster != -1) {  // NOT_BELONGED_
      if (cluster < 0 || static_cast<size_t>(cluster) * 3 >= color_set.size()) return -3;  // cv::Mat::at out of range
      b = color_set[3 * static_cast<size_t>(cluster)];
      g = color_set[3 * static_cast<size_t>(cluster) + 1];
      r = color_set[3 * static_cast<size_t>(cluster) + 2];
    }
    out_bgr8[3 * i] = static_cast<uint8_t>(b);  // cv::Vec3b(b, g, r)
    out_bgr8[3 * i + 1] = static_cast<uint8_t>(g);
    out_bgr8[3 * i + 2] = static_cast<uint8_t>(r);
  }
  return 0;
}

// clusterMap2IndicesCluster (clusterer_nodelet.cpp:97-117) on a final label map: per cluster the linear indices
// W*v+u in u-outer / v-inner order -- the point order of the ~clusters markers (cluster2Marker :119-145).
// out_offsets: number_of_clusters + 1 entries; out_indices: one entry per labelled pixel.
int orc_cluster_members(int w, int h, const int32_t* labels, int32_t number_of_clusters, int32_t* out_offsets,
                        int32_t* out_indices) {
  if (w <= 0 || h <= 0 || !labels || number_of_clusters < 0 || !out_offsets || !out_indices) return -1;
  std::vector<std::vector<int>> clusters(static_cast<size_t>(number_of_clusters));
  for (int u = 0; u < w; u++) {
    for (int v = 0; v < h; v++) {
      const int c = labels[static_cast<size_t>(w) * v + u];
      if (c == -1) continue;
      if (c < 0 || c >= number_of_clusters) return -3;  // indices_clusters.at(cluster_number) would throw
      clusters[static_cast<size_t>(c)].push_back(w * v + u);
    }
  }
  int32_t pos = 0;
  for (int c = 0; c < number_of_clusters; c++) {
    out_offsets[c] = pos;
    for (int i : clusters[static_cast<size_t>(c)]) out_indices[pos++] = i;
  }
  out_offsets[number_of_clusters] = pos;
  return 0;
}

// ---- image_crop ------------------------------------------------------------------------------------
// ImageCrop::cameraCallback, image_crop/src/image_crop.cpp:20-43.  roi origin = (msg.width - target_width_) / 2 in
// UNSIGNED 32-bit arithmetic (uint32_t - int), converted to int by cv::Rect; Mat::operator()(Rect) asserts the window
// lies inside the image (:25-26).  cv_bridge::toCvShare(msg, "bgr8") (:23) converts other 8-bit encodings by channel
// shuffles.  cropped_cx = K[2] - (double)((width - target_width_) / 2) with the same unsigned quotient (:32-33).
static inline void crop_origin(uint32_t w, uint32_t h, int tw, int th, uint32_t* ox, uint32_t* oy) {
  *ox = (w - static_cast<uint32_t>(tw)) / 2u;
  *oy = (h - static_cast<uint32_t>(th)) / 2u;
}

int orc_image_crop(int w, int h, int encoding, const uint8_t* data, int32_t step_bytes, int target_w, int target_h,
                   uint8_t* out_bgr8) {
  static const int bpp_of[5] = {3, 3, 4, 4, 1};
  static const int idx_of[5][3] = {{0, 1, 2}, {2, 1, 0}, {0, 1, 2}, {2, 1, 0}, {0, 0, 0}};  // source byte of B, G, R
  if (encoding < 0 || encoding > 4 || !data || !out_bgr8) return -1;
  uint32_t ox, oy;
  crop_origin(static_cast<uint32_t>(w), static_cast<uint32_t>(h), target_w, target_h, &ox, &oy);
  const int x0 = static_cast<int>(ox), y0 = static_cast<int>(oy);
  if (!(0 <= x0 && 0 <= target_w && static_cast<long long>(x0) + target_w <= w && 0 <= y0 && 0 <= target_h &&
        static_cast<long long>(y0) + target_h <= h))
    return -12;
  const int bpp = bpp_of[encoding];
  for (int r = 0; r < target_h; r++)
    for (int c = 0; c < target_w; c++) {
      const uint8_t* p = data + static_cast<size_t>(y0 + r) * step_bytes + static_cast<size_t>(x0 + c) * bpp;
      uint8_t* q = out_bgr8 + (static_cast<size_t>(r) * target_w + c) * 3;
      q[0] = p[idx_of[encoding][0]];
      q[1] = p[idx_of[encoding][1]];
      q[2] = p[idx_of[encoding][2]];
    }
  return 0;
}

// cv_bridge::toCvCopy(image, MONO8) (scene_flow_constructor.cpp:220-221) = cv::cvtColor(..., COLOR_*2GRAY) on 8-bit
// pixels: OpenCV's RGB2Gray<uchar> (imgproc/src/color.cpp; color_rgb.simd.hpp from 4.x on) is the fixed-point sum
// (cb*B + cg*G + cr*R + (1 << (shift-1))) >> shift.  rule 0: OpenCV 2.4 / 3.x (R2Y 4899, G2Y 9617, B2Y 1868, yuv_shift
// 14) -- the reference's OpenCV 3.2; rule 1: OpenCV 4.x (RY15 9798, GY15 19235, BY15 3735, gray_shift 15), which is the
// cv2 of this image and what tests/test_mono8.py pins the arithmetic against.  A mono8 source is copied.
int orc_to_mono8(int w, int h, int encoding, const uint8_t* data, int32_t step_bytes, int rule, uint8_t* out) {
  static const int bpp_of[5] = {3, 3, 4, 4, 1};
  static const int rules[2][4] = {{4899, 9617, 1868, 14}, {9798, 19235, 3735, 15}};  // cR, cG, cB, shift
  if (encoding < 0 || encoding > 4 || rule < 0 || rule > 1 || !data || !out) return -1;
  const int bpp = bpp_of[encoding];
  const bool blue_first = encoding == 0 || encoding == 2;
  const int cr = rules[rule][0], cg = rules[rule][1], cb = rules[rule][2], shift = rules[rule][3];
  for (int r = 0; r < h; r++)
    for (int c = 0; c < w; c++) {
      const uint8_t* p = data + static_cast<size_t>(r) * step_bytes + static_cast<size_t>(c) * bpp;
      uint8_t* q = out + static_cast<size_t>(r) * w + c;
      if (bpp == 1) {
        *q = p[0];
      } else {
        const int B = blue_first ? p[0] : p[2], G = p[1], R = blue_first ? p[2] : p[0];
        *q = static_cast<uint8_t>((cb * B + cg * G + cr * R + (1 << (shift - 1))) >> shift);
      }
    }
  return 0;
}

void orc_crop_camera_info(uint32_t w, uint32_t h, int target_w, int target_h, const double* K9, const double* P12,
                          double* out_K9, double* out_P12) {
  uint32_t ox, oy;
  crop_origin(w, h, target_w, target_h, &ox, &oy);
  for (int i = 0; i < 9; i++) out_K9[i] = K9[i];
  for (int i = 0; i < 12; i++) out_P12[i] = P12[i];
  const double cx = K9[2] - static_cast<double>(ox), cy = K9[5] - static_cast<double>(oy);  // :32-33 (from K, also for P)
  out_K9[2] = cx;
  out_K9[5] = cy;
  out_P12[2] = cx;
  out_P12[6] = cy;
}

void orc_quat_to_matrix(const orc_transform* t, double* R9) {
  const Isometry iso = isometryFromTransform(*t);
  std::memcpy(R9, iso.R, sizeof(iso.R));
}

int orc_lookup_table_run(int size, const int32_t* ops, int n_ops, int32_t* out, int max_out) {
  EquivalenceTable table;
  table.resize(size);
  table.reset();
  int n_out = 0;
  try {
    for (int i = 0; i < n_ops; i++) {
      const int op = ops[3 * i], a = ops[3 * i + 1], b = ops[3 * i + 2];
      int r;
      if (op == 0)
        r = table.addLabel();
      else if (op == 1) {
        table.link(a, b);
        continue;
      } else
        r = table.lookup(a);
      if (n_out < max_out) out[n_out] = r;
      n_out++;
    }
  } catch (const std::out_of_range&) {
    return -1;
  }
  return n_out;
}

int orc_bench_pipeline(int w, int h, const orc_camera* cam, float f, float T, float min_disp, float max_disp,
                       const float* disp_frames, const int32_t* prev_index, const int32_t* now_index,
                       const float* flows, const orc_transform* transforms, const double* dt, int n_pairs,
                       int32_t dynamic_flow_diff, int32_t cluster_size, double depth_diff, double dynamic_speed,
                       int32_t neighbor_distance, int n_threads, int reps, int stage_mask, double* best_sec,
                       int64_t* n_objects) {
  if (w <= 0 || h <= 0 || n_pairs <= 0 || n_threads <= 0 || reps <= 0) return -1;
  const size_t P = static_cast<size_t>(w) * h;
  double best = 1e300;
  std::atomic<int64_t> total_objects{0};
  // When only the clusterer is timed, the clouds are produced once up front (untimed).
  std::vector<std::vector<float>> clouds;
  if (!(stage_mask & 1)) {
    clouds.resize(n_pairs);
    for (int i = 0; i < n_pairs; i++) {
      clouds[i].resize(P * 8);
      orc_disparity now{f, T, min_disp, max_disp, disp_frames + P * now_index[i], w * 4};
      orc_disparity prev{f, T, min_disp, max_disp, disp_frames + P * prev_index[i], w * 4};
      orc_scene_flow(w, h, cam, &now, &prev, flows + P * 2 * i, w * 8, &transforms[i], dt[i], dynamic_flow_diff,
                     clouds[i].data(), nullptr, nullptr, nullptr);
    }
  }
  for (int rep = 0; rep < reps; rep++) {
    total_objects = 0;
    std::atomic<int> next{0};
    auto worker = [&]() {
      std::vector<float> cloud;
      std::vector<orc_object> objs(4096);
      for (;;) {
        const int i = next.fetch_add(1);
        if (i >= n_pairs) break;
        const float* cloud_ptr;
        if (stage_mask & 1) {
          cloud.resize(P * 8);
          orc_disparity now{f, T, min_disp, max_disp, disp_frames + P * now_index[i], w * 4};
          orc_disparity prev{f, T, min_disp, max_disp, disp_frames + P * prev_index[i], w * 4};
          orc_scene_flow(w, h, cam, &now, &prev, flows + P * 2 * i, w * 8, &transforms[i], dt[i],
                         dynamic_flow_diff, cloud.data(), nullptr, nullptr, nullptr);
          cloud_ptr = cloud.data();
        } else {
          cloud_ptr = clouds[i].data();
        }
        if (stage_mask & 2) {
          int32_t n = 0;
          orc_cluster(w, h, cloud_ptr, cluster_size, depth_diff, dynamic_speed, neighbor_distance, nullptr, nullptr,
                      nullptr, objs.data(), static_cast<int32_t>(objs.size()), &n);
          total_objects += n;
        }
      }
    };
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; t++) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    best = std::min(best, sec);
  }
  if (best_sec) *best_sec = best;
  if (n_objects) *n_objects = total_objects.load();
  return 0;
}

}  // extern "C"
