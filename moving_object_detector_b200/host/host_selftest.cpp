// host_selftest.cpp -- drives the C++ host classes the way the ROS callbacks would.
//   host_selftest --no-gpu                     checks that construction fails loudly without a GPU, the PointCloud2
//                                              description and ros::Time arithmetic (runs in the CPU test-suite)
//   host_selftest <case.bin> <out.bin>         runs construct() -> ~scene_flow -> dataCB() -> ~moving_objects,
//                                              ~clusters_image, ~clusters on one frame pair read from case.bin and
//                                              writes cloud + objects + image + markers (GPU tests compare them with
//                                              the oracle)
// case.bin: int32 w,h ; double cam[6] (fx,fy,cx,cy,Tx,Ty) ; float f,T,min,max ; double tr[7] ; uint32 sec0,nsec0,sec1,nsec1 ;
//           int32 dynamic_flow_diff, cluster_size ; double depth_diff, dynamic_speed ; int32 neighbor_distance, pad ;
//           float disp_prev[w*h], disp_now[w*h], flow[w*h*2]
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "scene_flow_host.h"

namespace sfc = scene_flow_constructor;
namespace cl = scene_flow_clusterer;

static int fail(const char* what) {
  std::fprintf(stderr, "host_selftest: FAILED: %s\n", what);
  return 1;
}

static int runNoGpu() {
  sensor_msgs::PointCloud2 pc;
  sfc::describeVelocityCloud(pc, 7, 3);
  if (pc.point_step != 32 || pc.row_step != 224 || pc.fields.size() != 6 || pc.fields[3].name != "vx" ||
      pc.fields[3].offset != 16 || pc.fields[5].offset != 24 || !pc.is_dense)
    return fail("PointCloud2 description");
  {
    // a foreign publisher's layout: 28-byte points {vz, x, intensity(u8 x4), y, vx, z, vy}, padded rows, no "vy" match
    // (FLOAT64 datatype) -> fromROSMsg semantics: fields by name, the unmatched one stays 0
        pc.data.assign(static_cast<size_t>(pc.row_step) * pc.height, 0);
    if (cl::velocityCloudLayout(pc) != 1) return fail("native layout not recognised");
    sensor_msgs::PointCloud2 other;
    other.width = 5; other.height = 2; other.point_step = 28; other.row_step = 28 * 5 + 12;
    const char* names[7] = {"vz", "x", "intensity", "y", "vx", "z", "vy"};
    for (int i = 0; i < 7; i++) {
      sensor_msgs::PointField f;
      f.name = names[i]; f.offset = 4 * i; f.count = 1;
      f.datatype = (i == 6) ? sensor_msgs::PointField::FLOAT64 : sensor_msgs::PointField::FLOAT32;
      other.fields.push_back(f);
    }
    other.data.assign(static_cast<size_t>(other.row_step) * other.height, 0xAB);
    for (uint32_t r = 0; r < 2; r++)
      for (uint32_t c = 0; c < 5; c++)
        for (int i = 0; i < 7; i++) {
          const float v = 100.0f * r + 10.0f * c + i;
          std::memcpy(&other.data[r * other.row_step + c * 28 + 4 * i], &v, 4);
        }
    if (cl::velocityCloudLayout(other) != 0) return fail("foreign layout not classified as repackable");
    std::vector<uint8_t> rec;
    cl::repackVelocityCloud(other, rec);
    if (rec.size() != 10 * 32) return fail("repack size");
    const float* q = reinterpret_cast<const float*>(rec.data());
    for (uint32_t r = 0; r < 2; r++)
      for (uint32_t c = 0; c < 5; c++) {
        const float b = 100.0f * r + 10.0f * c;
        const float* p = q + 8 * (r * 5 + c);
        if (p[0] != b + 1 || p[1] != b + 3 || p[2] != b + 5 || p[3] != 0 || p[4] != b + 4 || p[5] != 0 || p[6] != b + 0 ||
            p[7] != 0)
          return fail("repack values");
      }
    other.data.resize(10);
    if (cl::velocityCloudLayout(other) != -1) return fail("truncated cloud accepted");
  }
  ros::Time a, b;
  a.sec = 10; a.nsec = 100000000;
  b.sec = 9;  b.nsec = 900000000;
  if (std::fabs((a - b).toSec() - 0.2) > 1e-12) return fail("ros::Time difference");
  try {
    sfc::SceneFlowConstructorB200 c(64, 64, 0);
    std::printf("host_selftest: a GPU is present; --no-gpu checks skipped construction failure\n");
  } catch (const std::runtime_error& e) {
    std::printf("host_selftest: construction refused as expected: %s\n", e.what());
  }
  std::printf("host_selftest: no-gpu checks ok\n");
  return 0;
}

template <typename T>
static bool readv(FILE* f, T* p, size_t n) { return std::fread(p, sizeof(T), n, f) == n; }

int main(int argc, char** argv) {
  if (argc == 2 && std::strcmp(argv[1], "--no-gpu") == 0) return runNoGpu();
  if (argc != 3) return fail("usage: host_selftest --no-gpu | <case.bin> <out.bin>");
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return fail("cannot open case file");
  int32_t wh[2], ip[2], nd[2];
  double cam[6], tr[7], th[2];
  float meta[4];
  uint32_t stamps[4];
  if (!readv(f, wh, 2) || !readv(f, cam, 6) || !readv(f, meta, 4) || !readv(f, tr, 7) || !readv(f, stamps, 4) ||
      !readv(f, ip, 2) || !readv(f, th, 2) || !readv(f, nd, 2))
    return fail("short case header");
  const uint32_t w = wh[0], h = wh[1];
  const size_t P = static_cast<size_t>(w) * h;
  auto mkdisp = [&](uint32_t sec, uint32_t nsec) {
    auto d = std::make_shared<stereo_msgs::DisparityImage>();
    d->header.stamp.sec = sec;
    d->header.stamp.nsec = nsec;
    d->image.width = w;
    d->image.height = h;
    d->image.encoding = "32FC1";
    d->image.step = w * 4;
    d->image.data.resize(P * 4);
    d->f = meta[0]; d->T = meta[1]; d->min_disparity = meta[2]; d->max_disparity = meta[3];
    return d;
  };
  auto prev = mkdisp(stamps[0], stamps[1]), now = mkdisp(stamps[2], stamps[3]);
  auto flow = std::make_shared<sensor_msgs::Image>();
  flow->header = now->header;
  flow->width = w; flow->height = h; flow->encoding = "32FC2"; flow->step = w * 8;
  flow->data.resize(P * 8);
  if (!readv(f, prev->image.data.data(), P * 4) || !readv(f, now->image.data.data(), P * 4) ||
      !readv(f, flow->data.data(), P * 8))
    return fail("short case payload");
  std::fclose(f);
  auto transform = std::make_shared<geometry_msgs::Transform>();
  transform->translation.x = tr[0]; transform->translation.y = tr[1]; transform->translation.z = tr[2];
  transform->rotation.x = tr[3]; transform->rotation.y = tr[4]; transform->rotation.z = tr[5]; transform->rotation.w = tr[6];
  sensor_msgs::CameraInfo info;
  info.width = w; info.height = h;
  info.P = {cam[0], 0, cam[2], cam[4], 0, cam[1], cam[3], cam[5], 0, 0, 1, 0};
  info.header.frame_id = "left_camera";

  try {
    sfc::SceneFlowConstructorB200 constructor(static_cast<int>(w), static_cast<int>(h));
    cl::ClustererB200 clusterer(static_cast<int>(w), static_cast<int>(h));
    constructor.reconfigureCB(sfc::SceneFlowConstructorConfig{ip[0], 1.0}, 0);
    clusterer.reconfigureCB(cl::ClustererConfig{ip[1], th[0], th[1], nd[0]}, 0);
    constructor.setCameraInfo(info);

    // degrade rules first: nothing may be published without flow / previous frame / transform
    int published = 0;
    constructor.pc_with_velocity_pub_.publish = [&](const sensor_msgs::PointCloud2&) { published++; };
    constructor.pc_with_velocity_pub_.num_subscribers = 1;
    if (constructor.construct(now, prev, nullptr, transform) || published) return fail("published without flow");
    if (constructor.construct(now, nullptr, flow, transform) || published) return fail("published without previous frame");
    if (constructor.construct(now, prev, flow, nullptr) || published) return fail("published without transform");

    sensor_msgs::PointCloud2 cloud;
    moving_object_msgs::MovingObjectArray objects;
    bool got_objects = false;
    constructor.pc_with_velocity_pub_.publish = [&](const sensor_msgs::PointCloud2& m) { cloud = m; published++; };
    clusterer.dynamic_objects_pub_.publish = [&](const moving_object_msgs::MovingObjectArray& m) { objects = m; got_objects = true; };
    clusterer.dynamic_objects_pub_.num_subscribers = 1;
    visualization_msgs::MarkerArray markers;
    sensor_msgs::Image clusters_image;
    int order = 0, markers_at = -1, image_at = -1;
    clusterer.clusters_pub_.publish = [&](const visualization_msgs::MarkerArray& m) { markers = m; markers_at = order++; };
    clusterer.clusters_image_pub_.publish = [&](const sensor_msgs::Image& m) { clusters_image = m; image_at = order++; };
    if (!constructor.construct(now, prev, flow, transform)) return fail(constructor.lastError().c_str());
    if (published != 1 || cloud.data.size() != P * 32) return fail("~scene_flow not published");
    // unsubscribed visualisations must not be computed or published
    if (!clusterer.dataCB(cloud)) return fail(clusterer.lastError().c_str());
    if (!got_objects) return fail("~moving_objects not published");
    if (markers_at >= 0 || image_at >= 0) return fail("visualisation published without a subscriber");
    clusterer.clusters_pub_.num_subscribers = 1;
    clusterer.clusters_image_pub_.num_subscribers = 1;
    if (!clusterer.dataCB(cloud)) return fail(clusterer.lastError().c_str());
    if (markers_at != 0 || image_at != 1) return fail("~clusters / ~clusters_image not published in the reference's order");
    if (clusters_image.encoding != "bgr8" || clusters_image.step != w * 3 || clusters_image.data.size() != P * 3)
      return fail("~clusters_image layout");

    {
      // the same cloud from a foreign publisher: packed 24-byte points in another field order, padded rows.  The
      // nodelet maps fields by name (pcl::fromROSMsg), so the objects must be identical.
      sensor_msgs::PointCloud2 other;
      other.header = cloud.header;
      other.width = w; other.height = h; other.point_step = 24; other.row_step = 24 * w + 8; other.is_dense = cloud.is_dense;
      const char* names[6] = {"vx", "vy", "vz", "z", "y", "x"};
      const int from[6] = {4, 5, 6, 2, 1, 0};
      for (int i = 0; i < 6; i++) {
        sensor_msgs::PointField fd;
        fd.name = names[i]; fd.offset = 4 * i; fd.datatype = sensor_msgs::PointField::FLOAT32; fd.count = 1;
        other.fields.push_back(fd);
      }
      other.data.assign(static_cast<size_t>(other.row_step) * h, 0xCD);
      const float* src = reinterpret_cast<const float*>(cloud.data.data());
      for (uint32_t r = 0; r < h; r++)
        for (uint32_t c = 0; c < w; c++)
          for (int i = 0; i < 6; i++)
            std::memcpy(&other.data[static_cast<size_t>(r) * other.row_step + c * 24 + 4 * i],
                        &src[8 * (static_cast<size_t>(r) * w + c) + from[i]], 4);
      const moving_object_msgs::MovingObjectArray first = objects;
      got_objects = false;
      if (!clusterer.dataCB(other) || !got_objects) return fail("foreign PointCloud2 layout rejected");
      if (objects.moving_object_array.size() != first.moving_object_array.size()) return fail("foreign layout: object count");
      for (size_t i = 0; i < first.moving_object_array.size(); i++) {
        const auto &a = first.moving_object_array[i], &b = objects.moving_object_array[i];
        if (a.id != b.id || a.center.position.x != b.center.position.x || a.center.position.z != b.center.position.z ||
            a.bounding_box.y != b.bounding_box.y || a.velocity.x != b.velocity.x || a.velocity.z != b.velocity.z)
          return fail("foreign layout: objects differ");
      }
    }

    {
      // image_crop node: rgba8 camera image -> centred bgr8 crop + shifted principal point
      const uint32_t tw = w / 2 + 1, th = h / 2;
      image_crop::ImageCropB200 crop(static_cast<int>(w), static_cast<int>(h), static_cast<int>(tw), static_cast<int>(th));
      sensor_msgs::Image src;
      src.width = w; src.height = h; src.encoding = "rgba8"; src.step = w * 4 + 4;
      src.data.resize(static_cast<size_t>(src.step) * h);
      for (size_t i = 0; i < src.data.size(); i++) src.data[i] = static_cast<uint8_t>((i * 2654435761u) >> 13);
      sensor_msgs::CameraInfo ci = info;
      ci.K = {cam[0], 0, cam[2], 0, cam[1], cam[3], 0, 0, 1};
      sensor_msgs::Image got;
      sensor_msgs::CameraInfo got_info;
      int n_pub = 0;
      crop.camera_pub_ = [&](const sensor_msgs::Image& m, const sensor_msgs::CameraInfo& c) { got = m; got_info = c; n_pub++; };
      if (!crop.cameraCallback(src, ci) || n_pub != 1) return fail(crop.lastError().c_str());
      const uint32_t x0 = (w - tw) / 2, y0 = (h - th) / 2;
      if (got.width != tw || got.height != th || got.step != tw * 3 || got.encoding != "bgr8") return fail("cropped image layout");
      for (uint32_t r = 0; r < th; r++)
        for (uint32_t c = 0; c < tw; c++) {
          const uint8_t* p = &src.data[static_cast<size_t>(y0 + r) * src.step + (x0 + c) * 4];
          const uint8_t* q = &got.data[static_cast<size_t>(r) * got.step + c * 3];
          if (q[0] != p[2] || q[1] != p[1] || q[2] != p[0]) return fail("cropped pixels");
        }
      if (got_info.width != tw || got_info.height != th || got_info.K[2] != ci.K[2] - x0 || got_info.K[5] != ci.K[5] - y0 ||
          got_info.P[2] != ci.K[2] - x0 || got_info.P[6] != ci.K[5] - y0 || got_info.P[0] != ci.P[0] || got_info.P[3] != ci.P[3])
        return fail("cropped CameraInfo");
      image_crop::ImageCropB200 too_big(static_cast<int>(w), static_cast<int>(h), static_cast<int>(w + 1), static_cast<int>(h));
      too_big.camera_pub_ = crop.camera_pub_;
      if (too_big.cameraCallback(src, ci) || n_pub != 1) return fail("a window larger than the image was published");

      // estimateCameraMotion's MONO8 conversion of the same (strided, rgba8) camera image, both grey rules
      std::vector<uint8_t> mono;
      for (int rule = MOD_GRAY_OPENCV3; rule <= MOD_GRAY_OPENCV4; rule++) {
        if (!constructor.toMono8(src, mono, rule) || mono.size() != static_cast<size_t>(w) * h)
          return fail(constructor.lastError().c_str());
        const int cr = rule == MOD_GRAY_OPENCV3 ? 4899 : 9798, cg = rule == MOD_GRAY_OPENCV3 ? 9617 : 19235,
                  cb = rule == MOD_GRAY_OPENCV3 ? 1868 : 3735, sh = rule == MOD_GRAY_OPENCV3 ? 14 : 15;
        for (uint32_t r = 0; r < h; r++)
          for (uint32_t c = 0; c < w; c++) {
            const uint8_t* p = &src.data[static_cast<size_t>(r) * src.step + c * 4];  // r, g, b, a
            const int want = (cr * p[0] + cg * p[1] + cb * p[2] + (1 << (sh - 1))) >> sh;
            if (mono[static_cast<size_t>(r) * w + c] != want) return fail("toMono8 pixels");
          }
      }
      sensor_msgs::Image bayer = src;
      bayer.encoding = "bayer_rggb8";
      if (constructor.toMono8(bayer, mono)) return fail("an encoding cv_bridge converts differently was accepted");
    }

    FILE* o = std::fopen(argv[2], "wb");
    if (!o) return fail("cannot open output file");
    const int32_t n = static_cast<int32_t>(objects.moving_object_array.size());
    std::fwrite(cloud.data.data(), 1, cloud.data.size(), o);
    std::fwrite(&n, sizeof(n), 1, o);
    for (const auto& m : objects.moving_object_array) {
      const double rec[10] = {static_cast<double>(m.id), m.center.position.x, m.center.position.y, m.center.position.z,
                              m.velocity.x, m.velocity.y, m.velocity.z, m.bounding_box.x, m.bounding_box.y, m.bounding_box.z};
      std::fwrite(rec, sizeof(double), 10, o);
    }
    // ~clusters_image, then ~clusters: int32 n_markers; per marker int32 id,type,action,n_points; float r,g,b,a;
    // double scale x,y,z; n_points x 3 doubles
    std::fwrite(clusters_image.data.data(), 1, clusters_image.data.size(), o);
    const int32_t nm = static_cast<int32_t>(markers.markers.size());
    std::fwrite(&nm, sizeof(nm), 1, o);
    for (const auto& m : markers.markers) {
      const int32_t head[4] = {m.id, m.type, m.action, static_cast<int32_t>(m.points.size())};
      const float color[4] = {m.color.r, m.color.g, m.color.b, m.color.a};
      const double scale[3] = {m.scale.x, m.scale.y, m.scale.z};
      std::fwrite(head, sizeof(int32_t), 4, o);
      std::fwrite(color, sizeof(float), 4, o);
      std::fwrite(scale, sizeof(double), 3, o);
      for (const auto& pt : m.points) {
        const double xyz[3] = {pt.x, pt.y, pt.z};
        std::fwrite(xyz, sizeof(double), 3, o);
      }
    }
    std::fclose(o);
    std::printf("host_selftest: ok, %d moving objects\n", n);
  } catch (const std::exception& e) {
    return fail(e.what());
  }
  return 0;
}
