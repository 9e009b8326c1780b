// abi.cu -- the C ABI of include/mod_b200.h: context, workspaces, host<->device staging and kernel sequencing.
// No computation happens on the host except O(1) parameter preparation (quaternion -> rotation matrix in the
// reference's order of operations, float equivalents of the double thresholds).

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"
#include "hostcopy.cuh"

namespace mod {

float float_ceil_of_double(double th) {
  float f = static_cast<float>(th);
  if (static_cast<double>(f) < th) f = nextafterf(f, INFINITY);
  return f;
}
float float_floor_of_double(double th) {
  float f = static_cast<float>(th);
  if (static_cast<double>(f) > th) f = nextafterf(f, -INFINITY);
  return f;
}

// smallest float s2 >= +0 with sqrtf(s2) >= thr (sqrtf: IEEE correctly rounded on the host and as __fsqrt_rn)
float sq_threshold(float thr) {
  if (std::isnan(thr)) return thr;          // comparisons with NaN are false on both sides
  if (!(thr > 0.0f)) return 0.0f;           // every s >= +0 passes
  if (std::isinf(thr)) return thr;          // only s = +inf passes
  float s = thr * thr;                      // within a few ulps of the boundary
  if (std::isinf(s)) s = 3.402823466e38f;
  while (s > 0.0f && sqrtf(nextafterf(s, 0.0f)) >= thr) s = nextafterf(s, 0.0f);
  while (sqrtf(s) < thr) s = nextafterf(s, INFINITY);
  return s;
}

}  // namespace mod

using namespace mod;

struct mod_ctx {
  int device = 0;
  int max_w = 0, max_h = 0, max_batch = 0, max_clusters = 0;
  size_t maxP = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t s_in = nullptr, s_out = nullptr;
  char err[512] = {0};
  unsigned long long launches = 0;  // kernels launched by this context (mod_launch_count)

  // ray look-up tables (projectPixelTo3dRay), cached per camera / size
  double* d_ray = nullptr;
  float* d_rayf = nullptr;   // float roundings of the tables (static-flow filter), same layout
  unsigned long long* d_filter_fallbacks = nullptr;
  bool static_flow_filter = true;  // mod_set_static_flow_filter
  mod_camera lut_cam{};
  int lut_w = 0, lut_h = 0;
  int lut_stride = 0;        // offset of the row table inside d_ray / d_rayf

  // clusterer workspaces (max_batch frames)
  int* d_parent = nullptr;
  int* d_count = nullptr;
  TileStat* d_tstat = nullptr;
  uint32_t* d_list = nullptr;
  uint32_t* d_list_speed = nullptr;
  uint32_t* d_border_mask = nullptr;
  uint32_t* d_member_speed = nullptr;
  int* d_member_pixel = nullptr;
  int* d_frame_ints = nullptr;  // n_slots[max_batch], n_dynamic[max_batch]
  int* d_work = nullptr;        // work_count, work_cursor, list_count, border_count, root_count
  ClusterStat* d_stats = nullptr;
  int2* d_work_items = nullptr;
  int* d_member_offset = nullptr;
  int* d_cursor = nullptr;
  uint8_t* d_mask_own = nullptr;  // mask derived from a cloud when the caller has none (max_batch frames)

  // single-frame staging for the host seams (lazily allocated)
  bool host_ready = false;
  float* d_disp_now = nullptr;
  float* d_disp_prev = nullptr;
  float* d_flow = nullptr;
  float* d_points = nullptr;
  uint8_t* d_mask = nullptr;
  int32_t* d_labels = nullptr;
  float* d_depth = nullptr;
  float* d_static = nullptr;
  mod_object* d_objs = nullptr;
  mod_frame_stats* d_fstats = nullptr;
  mod_frame_stats* h_fstats = nullptr;  // pinned
  bool have_fstats = false;             // h_fstats holds the counters of a finished host-seam clusterer call
  HostStager stager;                    // pageable message buffers <-> device (hostcopy.cuh)
  // Frame identities for the host seams (mod_set_frame_ids): the disparity image uploaded as `now` stays on the device
  // and is reused when the next call names it as `prev` - in a stream every frame is uploaded once, as in the batched
  // sequence path, instead of twice.
  unsigned long long next_now_id = 0, next_prev_id = 0;  // consumed by the next mod_scene_flow / mod_pipeline call
  unsigned long long resident_id = 0;   // id of the image held in d_disp slot `resident_slot` (0: none)
  int resident_slot = 0, resident_w = 0, resident_h = 0;
  unsigned long long resident_hits = 0;
  int cloud_w = 0, cloud_h = 0;         // size of the cloud the last mod_scene_flow left in d_points (0: none)
  // the clusterer's kernel sequences of recent small-batch calls as instantiated CUDA graphs (run_cluster_dev), keyed
  // on the launch descriptors; a few entries, because the streaming path alternates between two sets of buffers
  struct ClGraph {
    cudaGraphExec_t exec = nullptr;
    ClLaunch L;
    ClExtra X;
    int launches = 0;
    unsigned long long last_use = 0;
  };
  static constexpr int kClGraphs = 4;
  ClGraph cl_graphs[kClGraphs];
  unsigned long long cl_graph_clock = 0;

  // streaming buffers of mod_pipeline_sequence_host (lazily allocated)
  struct SeqBuf {
    float* d_disp = nullptr;   // (chunk + 1) frames
    float* d_flow = nullptr;   // chunk pairs
    float* d_points = nullptr; // chunk pairs
    uint8_t* d_mask = nullptr;
    mod_object* d_objs = nullptr;
    mod_frame_stats* d_stats = nullptr;
    cudaEvent_t ev_in = nullptr, ev_comp = nullptr, ev_out = nullptr;
    int max_objs = 0;
  };
  int seq_chunk = 0;
  SeqBuf seq[2];
};

namespace {

int fail(mod_ctx* ctx, int code, const char* fmt, ...) {
  if (ctx) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

#define MOD_CK(ctx, call)                                                                                   \
  do {                                                                                                      \
    cudaError_t e__ = (call);                                                                               \
    if (e__ != cudaSuccess)                                                                                 \
      return fail((ctx), MOD_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e__)); \
  } while (0)

template <typename T>
cudaError_t dev_alloc(T** p, size_t count) {
  return cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T));
}

// tf2::transformToEigen + Quaterniond::toRotationMatrix (Eigen 3.3 order of operations).
// O(1) host-side parameter preparation; IEEE double on the host equals what the reference's host computes.
void fill_transform(const mod_transform& m, SfPair& p) {
  const double x = m.qx, y = m.qy, z = m.qz, w = m.qw;
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  p.R[0] = 1.0 - (tyy + tzz);
  p.R[1] = txy - twz;
  p.R[2] = txz + twy;
  p.R[3] = txy + twz;
  p.R[4] = 1.0 - (txx + tzz);
  p.R[5] = tyz - twx;
  p.R[6] = txz - twy;
  p.R[7] = tyz + twx;
  p.R[8] = 1.0 - (txx + tyy);
  p.t[0] = m.tx;
  p.t[1] = m.ty;
  p.t[2] = m.tz;
}

// Smallest previous-frame disparity for which isValid(pc_previous_transformed.at(previous)) (scene_flow_constructor.cpp
// :187-189) is certain without evaluating the point: for 0 < d_safe <= d the depth f*T/d is at most z_max, and z_max is
// chosen so that |R0 . (rx z, ry z, z) + t0| stays far below FLT_MAX with every intermediate finite.  Anything the
// bound cannot cover (non-finite or negative f*T, non-finite transform or camera) yields +inf: the kernel then
// evaluates the point for every pixel that reaches the gate.
float safe_disparity(const SfPair& p, float fT_prev, double ray_x_max, double ray_y_max) {
  const double fT = static_cast<double>(fT_prev);
  const double s = std::fabs(p.R[0]) * ray_x_max + std::fabs(p.R[1]) * ray_y_max + std::fabs(p.R[2]);
  const double t0 = std::fabs(p.t[0]);
  if (!(fT >= 0.0) || !std::isfinite(fT) || !std::isfinite(s) || !std::isfinite(t0) || !std::isfinite(ray_x_max) ||
      !std::isfinite(ray_y_max) || t0 > 1e30 || s > 1e30 || ray_x_max > 1e30 || ray_y_max > 1e30)
    return INFINITY;
  const double z_max = 1e30 / (s + ray_x_max + ray_y_max + 1.0);  // |x|,|y|,|z|,|m.x| <= ~1e30 + 1e30 << FLT_MAX
  double d = 4.0 * fT / z_max;                                    // float rounding of f*T/d is covered by the factor 4
  if (!(d > 1e-30)) d = 1e-30;                                    // stay clear of subnormal float disparities
  return mod::float_ceil_of_double(d);
}

// ---- static-flow filter constants (sceneflow.cu, "filtered residual gate") ----------------------------------------
// A constant is usable by the float evaluation if it is finite and either 0 or of moderate magnitude, so that its
// float rounding has relative error <= 2^-24 and no product of the bounded operands can overflow or go subnormal.
bool filter_const_ok(double c, double lo = 9.094947017729282e-13 /* 2^-40 */, double hi = 1099511627776.0 /* 2^40 */) {
  if (!std::isfinite(c)) return false;
  const double a = std::fabs(c);
  return a == 0.0 || (a >= lo && a <= hi);
}

float float_up(double v) { return mod::float_ceil_of_double(v * (1.0 + 1e-6)); }

// Camera-level constants; returns false when the filter must stay off for this camera / size.
bool fill_filter_camera(SfLaunch& L, const mod_camera& cam, int w, int h) {
  const double u = 5.9604644775390625e-08;  // 2^-24
  const double lo_f = 9.765625e-4 /* 2^-10 */, hi_f = 1073741824.0 /* 2^30 */;
  const bool ok = filter_const_ok(cam.fx, lo_f, hi_f) && cam.fx != 0.0 && filter_const_ok(cam.fy, lo_f, hi_f) &&
                  cam.fy != 0.0 && filter_const_ok(cam.cx) && filter_const_ok(cam.cy) && filter_const_ok(cam.Tx) &&
                  filter_const_ok(cam.Ty) && std::fabs(cam.cx) <= 1073741824.0 && std::fabs(cam.cy) <= 1073741824.0 &&
                  w <= (1 << 20) && h <= (1 << 20);
  L.fxf = static_cast<float>(cam.fx);
  L.fyf = static_cast<float>(cam.fy);
  L.cxf = static_cast<float>(cam.cx);
  L.cyf = static_cast<float>(cam.cy);
  L.Txf = static_cast<float>(cam.Tx);
  L.Tyf = static_cast<float>(cam.Ty);
  const double phi = std::fmax(std::fabs(cam.fx), std::fabs(cam.fy));
  const double Tm = std::fmax(std::fabs(cam.Tx), std::fabs(cam.Ty));
  L.aN = float_up(41.0 * u * phi);
  L.bN = float_up(2.1 * (phi * 1.01 * 8.673617379884035e-19 /* 2^-60 */ + 2.0 * u * Tm) + 8.673617379884035e-19);
  L.Cw = float_up(std::fabs(cam.cx) + std::fabs(cam.cy) + w + h);
  return ok;
}

// Pair-level constants; rho = +inf switches the filter off for the pair (every pixel is then evaluated exactly).
void fill_filter_pair(SfPair& p, bool camera_ok, float fT_prev, float min_prev, float max_prev) {
  bool ok = camera_ok && filter_const_ok(static_cast<double>(fT_prev)) && fT_prev != 0.0f && std::isfinite(min_prev) &&
            std::isfinite(max_prev) && std::fabs(min_prev) <= 1099511627776.0f &&
            std::fabs(max_prev) <= 1099511627776.0f;
  double rho = 0.0, tau = 0.0;
  float Rf[9], tf[3];
  for (int k = 0; k < 9; k++) {
    ok = ok && filter_const_ok(p.R[k]);
    Rf[k] = static_cast<float>(p.R[k]);
    rho = std::fmax(rho, std::fabs(p.R[k]));
  }
  for (int k = 0; k < 3; k++) {
    ok = ok && filter_const_ok(p.t[k]);
    tf[k] = static_cast<float>(p.t[k]);
    tau = std::fmax(tau, std::fabs(p.t[k]));
  }
  p.r0 = make_float4(Rf[0], Rf[1], Rf[2], tf[0]);
  p.r1 = make_float4(Rf[3], Rf[4], Rf[5], tf[1]);
  p.r2 = make_float4(Rf[6], Rf[7], Rf[8], tf[2]);
  p.g2.w = ok ? float_up(rho) : INFINITY;
  p.tau = ok ? float_up(tau) : INFINITY;
}

// The reference's disparity tests as closed intervals (common.cuh, SfPair).  getDisparity rejects d > max and
// d < min (disparity_image_processor.cpp:25-28), so a NaN bound rejects nothing; getRightPoint additionally wants
// 0 <= d < inf (scene_flow_constructor.h:220; -0.0f passes, and -0.0f >= 0.0f holds).
struct DisparityBounds {
  float lo_pt, hi_pt, lo_ch, hi_ch;
};
DisparityBounds fill_disparity_bounds(float dmin, float dmax) {
  DisparityBounds b;
  b.lo_pt = std::isnan(dmin) ? -INFINITY : dmin;
  b.hi_pt = std::isnan(dmax) ? INFINITY : dmax;
  b.lo_ch = b.lo_pt > 0.0f ? b.lo_pt : 0.0f;
  b.hi_ch = b.hi_pt < 3.402823466e38f ? b.hi_pt : 3.402823466e38f;
  return b;
}

bool same_camera(const mod_camera& a, const mod_camera& b) { return std::memcmp(&a, &b, sizeof(a)) == 0; }

int ensure_lut(mod_ctx* ctx, cudaStream_t s, int w, int h, const mod_camera& cam) {
  if (ctx->lut_w == w && ctx->lut_h == h && same_camera(ctx->lut_cam, cam)) return MOD_OK;
  launch_ray_lut(s, ctx->d_ray, ctx->d_ray + ctx->lut_stride, ctx->d_rayf, ctx->d_rayf + ctx->lut_stride, w, h, cam);
  MOD_CK(ctx, cudaGetLastError());
  ctx->launches++;
  ctx->lut_cam = cam;
  ctx->lut_w = w;
  ctx->lut_h = h;
  return MOD_OK;
}

int check_size(mod_ctx* ctx, int w, int h, int batch) {
  if (!ctx) return MOD_ERR_INVALID_ARG;
  if (w <= 0 || h <= 0 || batch <= 0) return fail(ctx, MOD_ERR_INVALID_ARG, "non-positive size %dx%d batch %d", w, h, batch);
  if (static_cast<long long>(w) * h >= (1ll << 30))  // the kernels index a frame (and its 2-float flow) with 32-bit ints
    return fail(ctx, MOD_ERR_INVALID_ARG, "frame %dx%d has 2^30 pixels or more", w, h);
  if (w > ctx->max_w || h > ctx->max_h)
    return fail(ctx, MOD_ERR_INVALID_ARG, "frame %dx%d exceeds the context's %dx%d", w, h, ctx->max_w, ctx->max_h);
  if (batch > ctx->max_batch) return fail(ctx, MOD_ERR_INVALID_ARG, "batch %d exceeds the context's %d", batch, ctx->max_batch);
  return MOD_OK;
}

// Scene-flow over `batch` device-resident pairs, chunked so the per-pair constants fit the parameter bank.
int run_scene_flow_dev(mod_ctx* ctx, cudaStream_t s, int w, int h, int batch, const mod_camera& cam,
                       const mod_pair_dev* pairs, const mod_sf_params& params, double dynamic_speed, float* d_points,
                       size_t frame_stride_points, uint8_t* d_mask, float* d_depth, float* d_static) {
  int rc = ensure_lut(ctx, s, w, h, cam);
  if (rc != MOD_OK) return rc;
  SfLaunch L;
  std::memset(&L, 0, sizeof(L));
  L.ray_x = ctx->d_ray;
  L.ray_y = ctx->d_ray + ctx->lut_stride;
  L.fx = cam.fx;
  L.fy = cam.fy;
  L.cx = cam.cx;
  L.cy = cam.cy;
  L.Tx = cam.Tx;
  L.Ty = cam.Ty;
  L.w = w;
  L.h = h;
  // std::sqrt(s) >= dynamic_flow_diff_: the int is promoted to float (cpp:198); sqrt-free form, see sq_threshold
  L.flow_thr2 = sq_threshold(static_cast<float>(params.dynamic_flow_diff));
  // (double)norm >= dynamic_speed_th_  <=>  norm >= ceil_float(th)  <=>  norm^2 >= sq_threshold(ceil_float(th))
  L.speed_thr2 = sq_threshold(float_ceil_of_double(dynamic_speed));
  // |ray| is largest at an image border (the ray is affine in the pixel coordinate)
  const double rx0 = std::fabs((0.0 - cam.cx - cam.Tx) / cam.fx), rx1 = std::fabs((w - 1.0 - cam.cx - cam.Tx) / cam.fx);
  const double ry0 = std::fabs((0.0 - cam.cy - cam.Ty) / cam.fy), ry1 = std::fabs((h - 1.0 - cam.cy - cam.Ty) / cam.fy);
  const double ray_x_max = 1.001 * (rx0 > rx1 ? rx0 : rx1), ray_y_max = 1.001 * (ry0 > ry1 ? ry0 : ry1);
  L.points = d_points;
  L.points_frame_stride = frame_stride_points;
  L.mask = d_mask;
  L.depth = d_depth;
  L.static_flow = d_static;
  L.ray_xf = ctx->d_rayf;
  L.ray_yf = ctx->d_rayf + ctx->lut_stride;
  const bool camera_ok = fill_filter_camera(L, cam, w, h);
  L.use_filter = ctx->static_flow_filter ? 1 : 0;
  L.filter_fallbacks = ctx->d_filter_fallbacks;
  for (int first = 0; first < batch; first += kSfMaxPairsPerLaunch) {
    const int n = (batch - first < kSfMaxPairsPerLaunch) ? batch - first : kSfMaxPairsPerLaunch;
    for (int k = 0; k < n; k++) {
      const mod_pair_dev& src = pairs[first + k];
      SfPair& p = L.pair[k];
      p.disp_now = src.disp_now;
      p.disp_prev = src.disp_prev;
      p.flow = src.flow_xy;
      fill_transform(src.prev2now, p);
      p.dt = src.dt_sec;
      const float fT_now = src.f_now * src.T_now;  // float product, disparity_image_processor.cpp:43
      const float fT_prev = src.f_prev * src.T_prev;
      const DisparityBounds bn = fill_disparity_bounds(src.min_disparity_now, src.max_disparity_now);
      const DisparityBounds bp = fill_disparity_bounds(src.min_disparity_prev, src.max_disparity_prev);
      p.g0 = make_float4(fT_now, bn.lo_pt, bn.hi_pt, bn.lo_ch);
      p.g1 = make_float4(bn.hi_ch, fT_prev, bp.lo_pt, bp.hi_pt);
      p.g2 = make_float4(bp.lo_ch, bp.hi_ch, safe_disparity(p, fT_prev, ray_x_max, ray_y_max), 0.0f);
      fill_filter_pair(p, camera_ok, fT_prev, src.min_disparity_prev, src.max_disparity_prev);
    }
    L.first_frame = first;
    MOD_CK(ctx, launch_scene_flow(s, L, n, first > 0));
    ctx->launches++;
  }
  return MOD_OK;
}

#ifndef MOD_CLUSTER_GRAPH_MAX_BATCH
#define MOD_CLUSTER_GRAPH_MAX_BATCH 8
#endif
constexpr int kClusterGraphMaxBatch = MOD_CLUSTER_GRAPH_MAX_BATCH;  // 0: never replay a graph

int run_cluster_dev(mod_ctx* ctx, cudaStream_t s, int w, int h, int batch, const float* d_points,
                    size_t frame_stride_points, const uint8_t* d_mask, const mod_cl_params& params,
                    int32_t* d_labels, mod_object* d_objs, int32_t max_objs, mod_frame_stats* d_stats) {
  if (params.cluster_size == 0)
    return fail(ctx, MOD_ERR_INVALID_PARAM, "cluster_size == 0: the reference throws std::out_of_range "
                "(cluster.at(0) on an empty cluster, clusterer_nodelet.cpp:174)");
  if (params.neighbor_distance < 0) return fail(ctx, MOD_ERR_INVALID_PARAM, "neighbor_distance < 0");
  if (params.neighbor_distance > kClMaxNeighborDistance)  // the tile kernels' halo and window arrays are sized for this
    return fail(ctx, MOD_ERR_INVALID_PARAM, "neighbor_distance %d > %d (cfg/Clusterer.cfg:11 limits it to [1, 10])",
                params.neighbor_distance, kClMaxNeighborDistance);
  if (max_objs < 0 || !d_objs || !d_stats) return fail(ctx, MOD_ERR_INVALID_ARG, "objs / stats missing");
  const size_t P = static_cast<size_t>(w) * h;
  if (P * static_cast<size_t>(batch) >= (1ull << 32))
    return fail(ctx, MOD_ERR_INVALID_ARG, "batch * w * h must stay below 2^32 pixels per call; split the batch");
  if (d_mask && (reinterpret_cast<uintptr_t>(d_mask) & 15u))
    return fail(ctx, MOD_ERR_INVALID_ARG, "d_mask must be 16-byte aligned");
  if (reinterpret_cast<uintptr_t>(d_points) & 31u)  // records are read as two float4 halves / written as one 256-bit store
    return fail(ctx, MOD_ERR_INVALID_ARG, "d_points32 must be 32-byte aligned");
  const float speed_thr = float_ceil_of_double(params.dynamic_speed);
  if (!d_mask) {
    MOD_CK(ctx, launch_dynamic_map(s, d_points, frame_stride_points, ctx->d_mask_own, w, h, batch, speed_thr));
    ctx->launches++;
    d_mask = ctx->d_mask_own;
  }
  ClLaunch L;
  std::memset(&L, 0, sizeof(L));
  L.points = d_points;
  L.points_frame_stride = frame_stride_points;
  L.mask = d_mask;
  L.w = w;
  L.h = h;
  L.batch = batch;
  L.n = params.neighbor_distance;
  L.depth_thr = float_floor_of_double(params.depth_diff);
  L.speed_thr = speed_thr;
  // removeSmallClusters compares size_t < (size_t)cluster_size (cpp:374): a negative threshold removes everything;
  // components need at least one edge, i.e. two pixels, to be labelled at all.
  L.min_size = params.cluster_size < 0 ? 0x7fffffff : (params.cluster_size < 2 ? 2 : params.cluster_size);
  L.max_clusters = ctx->max_clusters;
  L.max_objs = max_objs;
  L.parent = ctx->d_parent;
  L.count = ctx->d_count;
  L.tstat = ctx->d_tstat;
  L.n_slots = ctx->d_frame_ints;
  L.n_dynamic = ctx->d_frame_ints + batch;  // adjacent: one reset kernel clears both
  L.stats = ctx->d_stats;
  L.labels = d_labels;
  L.objs = d_objs;
  L.frame_stats = d_stats;
  ClExtra X;
  X.work_count = ctx->d_work;
  X.work_cursor = ctx->d_work + 1;
  X.list = ctx->d_list;
  X.list_speed = ctx->d_list_speed;
  X.border_mask = ctx->d_border_mask;
  X.work_items = ctx->d_work_items;
  X.member_offset = ctx->d_member_offset;
  X.cursor = ctx->d_cursor;
  X.member_speed = ctx->d_member_speed;
  X.member_pixel = ctx->d_member_pixel;
  (void)P;
  int n_launches = 0;
  // Small batches (a live node: one frame per call) are bound by launch latency - eight to ten kernels of a few
  // microseconds each.  Their launch parameters are a pure function of (L, X), which only change with the frame size,
  // the reconfigure values or the caller's buffers, so the sequence is captured once into a CUDA graph and replayed.
  cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
  const bool graph_ok = batch <= kClusterGraphMaxBatch && s != nullptr &&
                        cudaStreamIsCapturing(s, &capturing) == cudaSuccess && capturing == cudaStreamCaptureStatusNone;
  if (!graph_ok) {
    MOD_CK(ctx, launch_cluster(s, L, X, &n_launches));
    ctx->launches += n_launches;
    return MOD_OK;
  }
  mod_ctx::ClGraph* g = nullptr;
  mod_ctx::ClGraph* victim = &ctx->cl_graphs[0];
  for (auto& c : ctx->cl_graphs) {
    if (c.exec && std::memcmp(&c.L, &L, sizeof(L)) == 0 && std::memcmp(&c.X, &X, sizeof(X)) == 0) g = &c;
    if (c.last_use < victim->last_use) victim = &c;  // least recently used (unused entries carry 0)
  }
  if (!g) {
    g = victim;
    if (g->exec) {
      cudaGraphExecDestroy(g->exec);
      g->exec = nullptr;
    }
    MOD_CK(ctx, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    const cudaError_t launched = launch_cluster(s, L, X, &n_launches);
    cudaGraph_t graph = nullptr;
    const cudaError_t ended = cudaStreamEndCapture(s, &graph);
    MOD_CK(ctx, launched);
    MOD_CK(ctx, ended);
    const cudaError_t inst = cudaGraphInstantiate(&g->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (inst != cudaSuccess) g->exec = nullptr;
    MOD_CK(ctx, inst);
    g->L = L;
    g->X = X;
    g->launches = n_launches;
  }
  g->last_use = ++ctx->cl_graph_clock;
  MOD_CK(ctx, cudaGraphLaunch(g->exec, s));
  ctx->launches += g->launches;
  return MOD_OK;
}

int ensure_host_staging(mod_ctx* ctx) {
  if (ctx->host_ready) return MOD_OK;
  const size_t P = ctx->maxP;
  // a call that failed half-way leaves its buffers in place: only the missing ones are allocated on the next attempt
#define MOD_ENSURE(ptr, count) \
  if (!(ptr)) MOD_CK(ctx, dev_alloc(&(ptr), (count)))
  MOD_ENSURE(ctx->d_disp_now, P);
  MOD_ENSURE(ctx->d_disp_prev, P);
  MOD_ENSURE(ctx->d_flow, P * 2);
  MOD_ENSURE(ctx->d_points, P * 8);
  MOD_ENSURE(ctx->d_mask, P);
  MOD_ENSURE(ctx->d_labels, P);
  MOD_ENSURE(ctx->d_depth, P);
  MOD_ENSURE(ctx->d_static, P * 2);
  MOD_ENSURE(ctx->d_objs, static_cast<size_t>(ctx->max_clusters));
  MOD_ENSURE(ctx->d_fstats, 1);
#undef MOD_ENSURE
  if (!ctx->h_fstats) MOD_CK(ctx, cudaMallocHost(reinterpret_cast<void**>(&ctx->h_fstats), sizeof(mod_frame_stats)));
  ctx->host_ready = true;
  return MOD_OK;
}

int upload_image(mod_ctx* ctx, cudaStream_t s, float* dst, const float* src, int step_bytes, int w, int h,
                 int channels) {
  const size_t row = static_cast<size_t>(w) * channels * sizeof(float);
  if (static_cast<size_t>(step_bytes) < row) return fail(ctx, MOD_ERR_INVALID_ARG, "row step %d smaller than a row", step_bytes);
  if (static_cast<size_t>(step_bytes) == row) {  // dense rows (what cv_bridge and sgm_gpu emit): one staged copy
    MOD_CK(ctx, ctx->stager.h2d(dst, src, row * h, s));
  } else {
    MOD_CK(ctx, cudaMemcpy2DAsync(dst, row, src, static_cast<size_t>(step_bytes), row, h, cudaMemcpyHostToDevice, s));
  }
  return MOD_OK;
}

// Uploads the two disparity images of a host-seam call into the context's two staging slots and says which slot holds
// which.  With frame identities (mod_set_frame_ids) a `prev` that is the previous call's `now` is already there.
int upload_disparities(mod_ctx* ctx, cudaStream_t s, int w, int h, const mod_disparity& now, const mod_disparity& prev,
                       const float** d_now, const float** d_prev) {
  const unsigned long long now_id = ctx->next_now_id, prev_id = ctx->next_prev_id;
  ctx->next_now_id = ctx->next_prev_id = 0;
  float* slots[2] = {ctx->d_disp_now, ctx->d_disp_prev};
  int sp = 0;
  const bool reuse = prev_id != 0 && prev_id == ctx->resident_id && w == ctx->resident_w && h == ctx->resident_h;
  ctx->resident_id = 0;  // nothing is promised while the slots are being rewritten (or if an upload fails)
  int rc;
  if (reuse) {
    sp = ctx->resident_slot;
    ctx->resident_hits++;
  } else if ((rc = upload_image(ctx, s, slots[sp], prev.data, prev.step_bytes, w, h, 1)) != MOD_OK) {
    return rc;
  }
  const int sn = sp ^ 1;
  if ((rc = upload_image(ctx, s, slots[sn], now.data, now.step_bytes, w, h, 1)) != MOD_OK) return rc;
  ctx->resident_id = now_id;
  ctx->resident_slot = sn;
  ctx->resident_w = w;
  ctx->resident_h = h;
  *d_now = slots[sn];
  *d_prev = slots[sp];
  return MOD_OK;
}

mod_pair_dev make_pair(const float* d_now, const float* d_prev, const float* d_flow, const mod_disparity& now,
                       const mod_disparity& prev, const mod_transform& T, double dt) {
  mod_pair_dev p;
  std::memset(&p, 0, sizeof(p));
  p.disp_now = d_now;
  p.disp_prev = d_prev;
  p.flow_xy = d_flow;
  p.f_now = now.f;
  p.T_now = now.T;
  p.min_disparity_now = now.min_disparity;
  p.max_disparity_now = now.max_disparity;
  p.f_prev = prev.f;
  p.T_prev = prev.T;
  p.min_disparity_prev = prev.min_disparity;
  p.max_disparity_prev = prev.max_disparity;
  p.prev2now = T;
  p.dt_sec = dt;
  return p;
}

// Downloads objects + stats of one frame and converts the per-frame status into the return code.
int fetch_objects(mod_ctx* ctx, cudaStream_t s, mod_object* objs, int32_t max_objs, int32_t* n_objs) {
  ctx->have_fstats = false;
  MOD_CK(ctx, cudaMemcpyAsync(ctx->h_fstats, ctx->d_fstats, sizeof(mod_frame_stats), cudaMemcpyDeviceToHost, s));
  MOD_CK(ctx, cudaStreamSynchronize(s));
  ctx->have_fstats = true;
  const mod_frame_stats fs = *ctx->h_fstats;
  if (n_objs) *n_objs = fs.n_objects;
  const int n_copy = fs.n_objects < max_objs ? fs.n_objects : max_objs;
  if (n_copy > 0 && objs)
    MOD_CK(ctx, cudaMemcpy(objs, ctx->d_objs, static_cast<size_t>(n_copy) * sizeof(mod_object), cudaMemcpyDeviceToHost));
  if (fs.status == MOD_ERR_CLUSTER_SLOTS)
    return fail(ctx, MOD_ERR_CLUSTER_SLOTS, "more than %d surviving clusters in one frame", ctx->max_clusters);
  if (fs.status == MOD_ERR_CAPACITY) return fail(ctx, MOD_ERR_CAPACITY, "%d objects > max_objs %d", fs.n_objects, max_objs);
  return MOD_OK;
}

}  // namespace

extern "C" {

int mod_abi_version(void) { return MOD_B200_ABI_VERSION; }

void mod_host_rotation(const mod_transform* t, double* R9, double* t3) {
  if (!t) return;
  SfPair p;
  fill_transform(*t, p);
  if (R9) std::memcpy(R9, p.R, sizeof(p.R));
  if (t3) std::memcpy(t3, p.t, sizeof(p.t));
}
float mod_host_threshold_ge(double th) { return float_ceil_of_double(th); }
float mod_host_threshold_gt(double th) { return float_floor_of_double(th); }
float mod_host_sq_threshold(float thr) { return sq_threshold(thr); }

void* mod_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
  return p;
}
void mod_host_free(void* p) {
  if (p) cudaFreeHost(p);
}
int mod_host_register(void* p, size_t bytes) {
  if (!p || bytes == 0) return MOD_ERR_INVALID_ARG;
  if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) != cudaSuccess) {
    cudaGetLastError();
    return MOD_ERR_CUDA;
  }
  return MOD_OK;
}
int mod_host_unregister(void* p) {
  if (!p) return MOD_ERR_INVALID_ARG;
  if (cudaHostUnregister(p) != cudaSuccess) {
    cudaGetLastError();
    return MOD_ERR_CUDA;
  }
  return MOD_OK;
}
int mod_set_frame_ids(mod_ctx* ctx, unsigned long long now_id, unsigned long long prev_id) {
  if (!ctx) return MOD_ERR_INVALID_ARG;
  ctx->next_now_id = now_id;
  ctx->next_prev_id = prev_id;
  return MOD_OK;
}
unsigned long long mod_resident_hits(const mod_ctx* ctx) { return ctx ? ctx->resident_hits : 0; }
void mod_set_host_copy_threads(int n) { set_host_copy_threads(n); }
int mod_host_copy_threads(void) { return host_copy_threads(); }

const char* mod_last_error(const mod_ctx* ctx) { return ctx ? ctx->err : "null context"; }

unsigned long long mod_launch_count(const mod_ctx* ctx) { return ctx ? ctx->launches : 0; }

int mod_selfcheck_ddiv(mod_ctx* ctx, void* stream, const double* d_a, const double* d_b, size_t n,
                       unsigned long long* d_n_bad) {
  if (!ctx || !d_a || !d_b || !d_n_bad) return MOD_ERR_INVALID_ARG;
  MOD_CK(ctx, cudaSetDevice(ctx->device));
  MOD_CK(ctx, launch_ddiv_check(static_cast<cudaStream_t>(stream), d_a, d_b, n, d_n_bad));
  ctx->launches++;
  return MOD_OK;
}

int mod_set_static_flow_filter(mod_ctx* ctx, int enabled) {
  if (!ctx) return MOD_ERR_INVALID_ARG;
  ctx->static_flow_filter = enabled != 0;
  return MOD_OK;
}

int mod_filter_fallbacks(mod_ctx* ctx, int reset, unsigned long long* count) {
  if (!ctx || !count) return MOD_ERR_INVALID_ARG;
  MOD_CK(ctx, cudaSetDevice(ctx->device));
  MOD_CK(ctx, cudaDeviceSynchronize());
  MOD_CK(ctx, cudaMemcpy(count, ctx->d_filter_fallbacks, sizeof(*count), cudaMemcpyDeviceToHost));
  if (reset) MOD_CK(ctx, cudaMemset(ctx->d_filter_fallbacks, 0, sizeof(*count)));
  return MOD_OK;
}

int mod_last_frame_stats(const mod_ctx* ctx, mod_frame_stats* out) {
  if (!ctx || !out) return MOD_ERR_INVALID_ARG;
  if (!ctx->h_fstats || !ctx->have_fstats) return MOD_ERR_INVALID_ARG;
  *out = *ctx->h_fstats;
  return MOD_OK;
}

int mod_max_clusters(const mod_ctx* ctx) { return ctx ? ctx->max_clusters : 0; }

void* mod_stream(const mod_ctx* ctx) { return ctx ? static_cast<void*>(ctx->stream) : nullptr; }

int mod_create(int device, int max_w, int max_h, int max_batch, int max_clusters, mod_ctx** out_ctx) {
  if (!out_ctx) return MOD_ERR_INVALID_ARG;
  *out_ctx = nullptr;
  if (max_w <= 0 || max_h <= 0 || max_batch <= 0 || max_clusters < 0) return MOD_ERR_INVALID_ARG;
  if (static_cast<long long>(max_w) * max_h >= (1ll << 30)) return MOD_ERR_INVALID_ARG;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0 || device < 0 || device >= n_dev) return MOD_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MOD_ERR_NO_DEVICE;
  if (prop.major != 10) return MOD_ERR_NO_DEVICE;  // kernels are built for sm_100a only
  if (cudaSetDevice(device) != cudaSuccess) return MOD_ERR_NO_DEVICE;

  mod_ctx* ctx = new (std::nothrow) mod_ctx();
  if (!ctx) return MOD_ERR_INVALID_ARG;
  ctx->device = device;
  ctx->max_w = max_w;
  ctx->max_h = max_h;
  ctx->max_batch = max_batch;
  ctx->max_clusters = max_clusters == 0 ? 4096 : (max_clusters > kClMaxClusters ? kClMaxClusters : max_clusters);
  ctx->maxP = static_cast<size_t>(max_w) * max_h;
  const size_t N = ctx->maxP * max_batch;
  const size_t C = static_cast<size_t>(ctx->max_clusters) * max_batch;
  cudaError_t e = cudaSuccess;
  auto ok = [&](cudaError_t r) {
    if (e == cudaSuccess) e = r;
  };
  ok(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  ok(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
  ok(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
  // the row table starts at max_w rounded up to even so that the second table of each pair stays 16 / 8-byte aligned
  ctx->lut_stride = max_w + (max_w & 1);
  ok(dev_alloc(&ctx->d_ray, static_cast<size_t>(ctx->lut_stride) + max_h));
  ok(dev_alloc(&ctx->d_rayf, static_cast<size_t>(ctx->lut_stride) + max_h));
  ok(dev_alloc(&ctx->d_filter_fallbacks, 1));
  if (e == cudaSuccess) ok(cudaMemset(ctx->d_filter_fallbacks, 0, sizeof(unsigned long long)));
  ok(dev_alloc(&ctx->d_parent, N));
  ok(dev_alloc(&ctx->d_count, N));
  ok(dev_alloc(&ctx->d_tstat, N));
  ok(dev_alloc(&ctx->d_list, N));
  ok(dev_alloc(&ctx->d_list_speed, N));
  ok(dev_alloc(&ctx->d_border_mask, N));
  ok(dev_alloc(&ctx->d_member_speed, N));
  ok(dev_alloc(&ctx->d_member_pixel, N));
  ok(dev_alloc(&ctx->d_mask_own, N));
  ok(dev_alloc(&ctx->d_frame_ints, static_cast<size_t>(2) * max_batch));
  ok(dev_alloc(&ctx->d_work, 8));
  ok(dev_alloc(&ctx->d_stats, C));
  ok(dev_alloc(&ctx->d_work_items, C));
  ok(dev_alloc(&ctx->d_member_offset, C));
  ok(dev_alloc(&ctx->d_cursor, C));
  if (e != cudaSuccess) {
    mod_destroy(ctx);
    return MOD_ERR_CUDA;
  }
  *out_ctx = ctx;
  return MOD_OK;
}

void mod_destroy(mod_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  cudaFree(ctx->d_ray);
  cudaFree(ctx->d_rayf);
  cudaFree(ctx->d_filter_fallbacks);
  cudaFree(ctx->d_parent);
  cudaFree(ctx->d_count);
  cudaFree(ctx->d_tstat);
  cudaFree(ctx->d_list);
  cudaFree(ctx->d_list_speed);
  cudaFree(ctx->d_border_mask);
  cudaFree(ctx->d_member_speed);
  cudaFree(ctx->d_member_pixel);
  cudaFree(ctx->d_mask_own);
  cudaFree(ctx->d_frame_ints);
  cudaFree(ctx->d_work);
  cudaFree(ctx->d_stats);
  cudaFree(ctx->d_work_items);
  cudaFree(ctx->d_member_offset);
  cudaFree(ctx->d_cursor);
  cudaFree(ctx->d_disp_now);
  cudaFree(ctx->d_disp_prev);
  cudaFree(ctx->d_flow);
  cudaFree(ctx->d_points);
  cudaFree(ctx->d_mask);
  cudaFree(ctx->d_labels);
  cudaFree(ctx->d_depth);
  cudaFree(ctx->d_static);
  cudaFree(ctx->d_objs);
  cudaFree(ctx->d_fstats);
  if (ctx->h_fstats) cudaFreeHost(ctx->h_fstats);
  for (auto& b : ctx->seq) {
    cudaFree(b.d_disp);
    cudaFree(b.d_flow);
    cudaFree(b.d_points);
    cudaFree(b.d_mask);
    cudaFree(b.d_objs);
    cudaFree(b.d_stats);
    if (b.ev_in) cudaEventDestroy(b.ev_in);
    if (b.ev_comp) cudaEventDestroy(b.ev_comp);
    if (b.ev_out) cudaEventDestroy(b.ev_out);
  }
  for (auto& c : ctx->cl_graphs)
    if (c.exec) cudaGraphExecDestroy(c.exec);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
  if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
  delete ctx;
}

int mod_scene_flow(mod_ctx* ctx, int w, int h, const mod_camera* cam, const mod_disparity* now,
                   const mod_disparity* prev, const float* flow_xy, int32_t flow_step_bytes,
                   const mod_transform* prev2now, double dt_sec, const mod_sf_params* params, void* out_points32,
                   float* opt_depth, float* opt_static_flow) {
  int rc = check_size(ctx, w, h, 1);
  if (rc != MOD_OK) return rc;
  if (!cam || !now || !prev || !now->data || !prev->data || !flow_xy || !prev2now || !params)
    return fail(ctx, MOD_ERR_INVALID_ARG, "null argument");
  MOD_CK(ctx, cudaSetDevice(ctx->device));
  ctx->cloud_w = ctx->cloud_h = 0;  // no resident cloud until this call has enqueued its kernel (mod_cluster_resident)
  if ((rc = ensure_host_staging(ctx)) != MOD_OK) return rc;
  cudaStream_t s = ctx->stream;
  const size_t P = static_cast<size_t>(w) * h;
  const float *d_now = nullptr, *d_prev = nullptr;
  if ((rc = upload_disparities(ctx, s, w, h, *now, *prev, &d_now, &d_prev)) != MOD_OK) return rc;
  if ((rc = upload_image(ctx, s, ctx->d_flow, flow_xy, flow_step_bytes, w, h, 2)) != MOD_OK) return rc;
  const mod_pair_dev pair = make_pair(d_now, d_prev, ctx->d_flow, *now, *prev, *prev2now, dt_sec);
  rc = run_scene_flow_dev(ctx, s, w, h, 1, *cam, &pair, *params, 0.0, ctx->d_points, P, nullptr,
                          opt_depth ? ctx->d_depth : nullptr, opt_static_flow ? ctx->d_static : nullptr);
  if (rc != MOD_OK) return rc;
  ctx->cloud_w = w;
  ctx->cloud_h = h;
  if (out_points32) MOD_CK(ctx, ctx->stager.d2h(out_points32, ctx->d_points, P * 32, s));
  if (opt_depth) MOD_CK(ctx, ctx->stager.d2h(opt_depth, ctx->d_depth, P * 4, s));
  if (opt_static_flow) MOD_CK(ctx, ctx->stager.d2h(opt_static_flow, ctx->d_static, P * 8, s));
  MOD_CK(ctx, cudaStreamSynchronize(s));
  return MOD_OK;
}

int mod_cluster_resident(mod_ctx* ctx, int w, int h, const mod_cl_params* params, int32_t* opt_labels,
                         mod_object* objs, int32_t max_objs, int32_t* n_objs) {
  int rc = check_size(ctx, w, h, 1);
  if (rc != MOD_OK) return rc;
  if (!params || max_objs < 0 || (max_objs > 0 && !objs)) return fail(ctx, MOD_ERR_INVALID_ARG, "null argument");
  if (ctx->cloud_w != w || ctx->cloud_h != h)
    return fail(ctx, MOD_ERR_INVALID_ARG, "no %dx%d cloud from a previous mod_scene_flow call on this context", w, h);
  MOD_CK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const size_t P = static_cast<size_t>(w) * h;
  const int32_t cap = max_objs < ctx->max_clusters ? max_objs : ctx->max_clusters;
  rc = run_cluster_dev(ctx, s, w, h, 1, ctx->d_points, P, nullptr, *params, opt_labels ? ctx->d_labels : nullptr,
                       ctx->d_objs, cap, ctx->d_fstats);
  if (rc != MOD_OK) return rc;
  if (opt_labels) MOD_CK(ctx, ctx->stager.d2h(opt_labels, ctx->d_labels, P * 4, s));
  return fetch_objects(ctx, s, objs, cap, n_objs);
}

int mod_cluster(mod_ctx* ctx, int w, int h, const void* points32, const mod_cl_params* params, int32_t* opt_labels,
                mod_object* objs, int32_t max_objs, int32_t* n_objs) {
  int rc = check_size(ctx, w, h, 1);
  if (rc != MOD_OK) return rc;
  if (!points32 || !params || max_objs < 0 || (max_objs > 0 && !objs)) return fail(ctx, MOD_ERR_INVALID_ARG, "null argument");
  MOD_CK(ctx, cudaSetDevice(ctx->device));
  if ((rc = ensure_host_staging(ctx)) != MOD_OK) return rc;
  cudaStream_t s = ctx->stream;
  const size_t P = static_cast<size_t>(w) * h;
  ctx->cloud_w = ctx->cloud_h = 0;  // replaced by the caller's cloud
  MOD_CK(ctx, ctx->stager.h2d(ctx->d_points, points32, P * 32, s));
  const int32_t cap = max_objs < ctx->max_clusters ? max_objs : ctx->max_clusters;
  rc = run_cluster_dev(ctx, s, w, h, 1, ctx->d_points, P, nullptr, *params, opt_labels ? ctx->d_labels : nullptr,
                       ctx->d_objs, cap, ctx->d_fstats);
  if (rc != MOD_OK) return rc;
  if (opt_labels) MOD_CK(ctx, ctx->stager.d2h(opt_labels, ctx->d_labels, P * 4, s));
  return fetch_objects(ctx, s, objs, cap, n_objs);
}

int mod_pipeline(mod_ctx* ctx, int w, int h, const mod_camera* cam, const mod_disparity* now,
                 const mod_disparity* prev, const float* flow_xy, int32_t flow_step_bytes,
                 const mod_transform* prev2now, double dt_sec, const mod_sf_params* sf_params,
                 const mod_cl_params* cl_params, void* opt_out_points32, int32_t* opt_labels, mod_object* objs,
                 int32_t max_objs, int32_t* n_objs) {
  int rc = check_size(ctx, w, h, 1);
  if (rc != MOD_OK) return rc;
  if (!cam || !now || !prev || !now->data || !prev->data || !flow_xy || !prev2now || !sf_params || !cl_params ||
      max_objs < 0 || (max_objs > 0 && !objs))
    return fail(ctx, MOD_ERR_INVALID_ARG, "null argument");
  MOD_CK(ctx, cudaSetDevice(ctx->device));
  ctx->cloud_w = ctx->cloud_h = 0;  // no resident cloud until this call has enqueued its kernel (mod_cluster_resident)
  if ((rc = ensure_host_staging(ctx)) != MOD_OK) return rc;
  cudaStream_t s = ctx->stream;
  const size_t P = static_cast<size_t>(w) * h;
  const float *d_now = nullptr, *d_prev = nullptr;
  if ((rc = upload_disparities(ctx, s, w, h, *now, *prev, &d_now, &d_prev)) != MOD_OK) return rc;
  if ((rc = upload_image(ctx, s, ctx->d_flow, flow_xy, flow_step_bytes, w, h, 2)) != MOD_OK) return rc;
  const mod_pair_dev pair = make_pair(d_now, d_prev, ctx->d_flow, *now, *prev, *prev2now, dt_sec);
  rc = run_scene_flow_dev(ctx, s, w, h, 1, *cam, &pair, *sf_params, cl_params->dynamic_speed, ctx->d_points, P,
                          ctx->d_mask, nullptr, nullptr);
  if (rc != MOD_OK) return rc;
  const int32_t cap = max_objs < ctx->max_clusters ? max_objs : ctx->max_clusters;
  rc = run_cluster_dev(ctx, s, w, h, 1, ctx->d_points, P, ctx->d_mask, *cl_params,
                       opt_labels ? ctx->d_labels : nullptr, ctx->d_objs, cap, ctx->d_fstats);
  if (rc != MOD_OK) return rc;
  ctx->cloud_w = w;
  ctx->cloud_h = h;
  if (opt_out_points32) MOD_CK(ctx, ctx->stager.d2h(opt_out_points32, ctx->d_points, P * 32, s));
  if (opt_labels) MOD_CK(ctx, ctx->stager.d2h(opt_labels, ctx->d_labels, P * 4, s));
  return fetch_objects(ctx, s, objs, cap, n_objs);
}

int mod_scene_flow_batched_dev(mod_ctx* ctx, void* stream, int w, int h, int batch, const mod_camera* cam,
                               const mod_pair_dev* pairs, const mod_sf_params* params, double dynamic_speed,
                               void* d_points32, size_t frame_stride_points, uint8_t* d_mask, float* d_opt_depth,
                               float* d_opt_static_flow) {
  if (!ctx) return MOD_ERR_INVALID_ARG;
  if (w <= 0 || h <= 0 || batch <= 0 || w > ctx->max_w || h > ctx->max_h)
    return fail(ctx, MOD_ERR_INVALID_ARG, "bad size %dx%d batch %d", w, h, batch);
  if (!cam || !pairs || !params || !d_points32 || frame_stride_points < static_cast<size_t>(w) * h)
    return fail(ctx, MOD_ERR_INVALID_ARG, "null argument or frame stride smaller than a frame");
  if (reinterpret_cast<uintptr_t>(d_points32) & 31u)  // one 256-bit store per record
    return fail(ctx, MOD_ERR_INVALID_ARG, "d_points32 must be 32-byte aligned");
  MOD_CK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return run_scene_flow_dev(ctx, s, w, h, batch, *cam, pairs, *params, dynamic_speed, static_cast<float*>(d_points32),
                            frame_stride_points, d_mask, d_opt_depth, d_opt_static_flow);
}

int mod_cluster_batched_dev(mod_ctx* ctx, void* stream, int w, int h, int batch, const void* d_points32,
                            size_t frame_stride_points, const uint8_t* d_mask, const mod_cl_params* params,
                            int32_t* d_opt_labels, mod_object* d_objs, int32_t max_objs, mod_frame_stats* d_stats) {
  int rc = check_size(ctx, w, h, batch);
  if (rc != MOD_OK) return rc;
  if (!d_points32 || !params || frame_stride_points < static_cast<size_t>(w) * h)
    return fail(ctx, MOD_ERR_INVALID_ARG, "null argument or frame stride smaller than a frame");
  MOD_CK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return run_cluster_dev(ctx, s, w, h, batch, static_cast<const float*>(d_points32), frame_stride_points, d_mask,
                         *params, d_opt_labels, d_objs, max_objs, d_stats);
}

int mod_clusters_image_batched_dev(mod_ctx* ctx, void* stream, int w, int h, int batch, const int32_t* d_labels,
                                   const mod_frame_stats* d_stats, uint8_t* d_out_bgr8) {
  int rc = check_size(ctx, w, h, batch);
  if (rc != MOD_OK) return rc;
  if (!d_labels || !d_stats || !d_out_bgr8) return fail(ctx, MOD_ERR_INVALID_ARG, "null argument");
  MOD_CK(ctx, cudaSetDevice(ctx->device));
  MOD_CK(ctx, launch_clusters_image(static_cast<cudaStream_t>(stream), d_labels, d_stats, 0, w, h, batch, d_out_bgr8));
  ctx->launches++;
  return MOD_OK;
}

int mod_clusters_image(mod_ctx* ctx, int w, int h, const int32_t* labels, int32_t n_clusters, uint8_t* out_bgr8) {
  int rc = check_size(ctx, w, h, 1);
  if (rc != MOD_OK) return rc;
  if (!labels || !out_bgr8) return fail(ctx, MOD_ERR_INVALID_ARG, "null argument");
  MOD_CK(ctx, cudaSetDevice(ctx->device));
  rc = ensure_host_staging(ctx);
  if (rc != MOD_OK) return rc;
  const size_t P = static_cast<size_t>(w) * h;
  cudaStream_t s = ctx->stream;
  ctx->cloud_w = ctx->cloud_h = 0;  // the cloud buffer doubles as staging here
  uint8_t* d_bgr = reinterpret_cast<uint8_t*>(ctx->d_points);  // 32 bytes per pixel of staging, 3 needed
  MOD_CK(ctx, ctx->stager.h2d(ctx->d_labels, labels, P * sizeof(int32_t), s));
  MOD_CK(ctx, launch_clusters_image(s, ctx->d_labels, nullptr, n_clusters, w, h, 1, d_bgr));
  ctx->launches++;
  MOD_CK(ctx, ctx->stager.d2h(out_bgr8, d_bgr, P * 3, s));
  MOD_CK(ctx, cudaStreamSynchronize(s));
  return MOD_OK;
}

// crop origin of image_crop.cpp:24 (uint32_t - int, unsigned division) and the ROI test of cv::Mat::operator()(Rect)
static bool crop_window(uint32_t w, uint32_t h, int tw, int th, int* x0, int* y0) {
  const uint32_t ox = (w - static_cast<uint32_t>(tw)) / 2u, oy = (h - static_cast<uint32_t>(th)) / 2u;
  *x0 = static_cast<int>(ox);
  *y0 = static_cast<int>(oy);
  return *x0 >= 0 && tw >= 0 && static_cast<long long>(*x0) + tw <= static_cast<long long>(w) && *y0 >= 0 && th >= 0 &&
         static_cast<long long>(*y0) + th <= static_cast<long long>(h);
}

int mod_image_crop_batched_dev(mod_ctx* ctx, void* stream, int w, int h, int batch, int encoding, const uint8_t* d_data,
                               size_t frame_stride_bytes, int32_t step_bytes, int target_w, int target_h,
                               uint8_t* d_out_bgr8) {
  if (!ctx) return MOD_ERR_INVALID_ARG;
  if (w <= 0 || h <= 0 || batch <= 0) return fail(ctx, MOD_ERR_INVALID_ARG, "bad image size or batch");
  if (encoding < MOD_ENC_BGR8 || encoding > MOD_ENC_MONO8) return fail(ctx, MOD_ERR_INVALID_ARG, "unknown encoding %d", encoding);
  static const int bpp_of[5] = {3, 3, 4, 4, 1};
  if (step_bytes < w * bpp_of[encoding]) return fail(ctx, MOD_ERR_INVALID_ARG, "step %d shorter than a row", step_bytes);
  int x0, y0;
  if (!crop_window(static_cast<uint32_t>(w), static_cast<uint32_t>(h), target_w, target_h, &x0, &y0))
    return fail(ctx, MOD_ERR_INVALID_ARG, "crop window %dx%d does not lie inside the %dx%d image", target_w, target_h, w, h);
  if (!d_data || (!d_out_bgr8 && target_w > 0 && target_h > 0)) return fail(ctx, MOD_ERR_INVALID_ARG, "null argument");
  MOD_CK(ctx, cudaSetDevice(ctx->device));
  MOD_CK(ctx, launch_image_crop(static_cast<cudaStream_t>(stream), d_data, frame_stride_bytes, step_bytes, encoding, x0,
                                y0, target_w, target_h, batch, d_out_bgr8));
  if (target_w > 0 && target_h > 0) ctx->launches++;
  return MOD_OK;
}

int mod_image_crop(mod_ctx* ctx, int w, int h, int encoding, const uint8_t* data, int32_t step_bytes, int target_w,
                   int target_h, uint8_t* out_bgr8) {
  int rc = check_size(ctx, w, h, 1);
  if (rc != MOD_OK) return rc;
  if (!data) return fail(ctx, MOD_ERR_INVALID_ARG, "null argument");
  if (step_bytes <= 0 || static_cast<size_t>(step_bytes) > 16 * static_cast<size_t>(ctx->max_w))
    return fail(ctx, MOD_ERR_INVALID_ARG, "step %d outside (0, 16*max_w]", step_bytes);
  MOD_CK(ctx, cudaSetDevice(ctx->device));
  rc = ensure_host_staging(ctx);
  if (rc != MOD_OK) return rc;
  // staging: the 32 bytes-per-pixel cloud buffer holds the source (at most 16 bytes per pixel of row stride) in its
  // first half and the crop (3 bytes per pixel) in its second half
  cudaStream_t s = ctx->stream;
  ctx->cloud_w = ctx->cloud_h = 0;  // the cloud buffer doubles as staging here
  uint8_t* d_src = reinterpret_cast<uint8_t*>(ctx->d_points);
  uint8_t* d_out = d_src + 16 * static_cast<size_t>(ctx->max_w) * ctx->max_h;
  MOD_CK(ctx, ctx->stager.h2d(d_src, data, static_cast<size_t>(step_bytes) * h, s));
  rc = mod_image_crop_batched_dev(ctx, s, w, h, 1, encoding, d_src, 0, step_bytes, target_w, target_h, d_out);
  if (rc != MOD_OK) return rc;
  const size_t out_bytes = static_cast<size_t>(target_w) * target_h * 3;
  if (out_bytes > 0) {
    if (!out_bgr8) return fail(ctx, MOD_ERR_INVALID_ARG, "null argument");
    MOD_CK(ctx, ctx->stager.d2h(out_bgr8, d_out, out_bytes, s));
  }
  MOD_CK(ctx, cudaStreamSynchronize(s));
  return MOD_OK;
}

int mod_host_encoding(const char* ros_encoding) {
  static const char* const names[5] = {"bgr8", "rgb8", "bgra8", "rgba8", "mono8"};
  if (!ros_encoding) return MOD_ERR_INVALID_ARG;
  for (int i = 0; i < 5; i++)
    if (strcmp(ros_encoding, names[i]) == 0) return i;
  return MOD_ERR_INVALID_ARG;
}

int mod_image_to_mono8_batched_dev(mod_ctx* ctx, void* stream, int w, int h, int batch, int encoding,
                                   const uint8_t* d_data, size_t frame_stride_bytes, int32_t step_bytes, int rule,
                                   uint8_t* d_out_mono8) {
  if (!ctx) return MOD_ERR_INVALID_ARG;
  if (w <= 0 || h <= 0 || batch <= 0) return fail(ctx, MOD_ERR_INVALID_ARG, "bad image size or batch");
  if (encoding < MOD_ENC_BGR8 || encoding > MOD_ENC_MONO8) return fail(ctx, MOD_ERR_INVALID_ARG, "unknown encoding %d", encoding);
  if (rule != MOD_GRAY_OPENCV3 && rule != MOD_GRAY_OPENCV4) return fail(ctx, MOD_ERR_INVALID_ARG, "unknown grey rule %d", rule);
  static const int bpp_of[5] = {3, 3, 4, 4, 1};
  if (step_bytes < w * bpp_of[encoding]) return fail(ctx, MOD_ERR_INVALID_ARG, "step %d shorter than a row", step_bytes);
  if (!d_data || !d_out_mono8) return fail(ctx, MOD_ERR_INVALID_ARG, "null argument");
  MOD_CK(ctx, cudaSetDevice(ctx->device));
  MOD_CK(ctx, launch_to_mono8(static_cast<cudaStream_t>(stream), d_data, frame_stride_bytes, step_bytes, encoding, w, h,
                              batch, rule, d_out_mono8));
  ctx->launches++;
  return MOD_OK;
}

int mod_image_to_mono8(mod_ctx* ctx, int w, int h, int encoding, const uint8_t* data, int32_t step_bytes, int rule,
                       uint8_t* out_mono8) {
  int rc = check_size(ctx, w, h, 1);
  if (rc != MOD_OK) return rc;
  if (!data || !out_mono8) return fail(ctx, MOD_ERR_INVALID_ARG, "null argument");
  if (step_bytes <= 0 || static_cast<size_t>(step_bytes) > 16 * static_cast<size_t>(ctx->max_w))
    return fail(ctx, MOD_ERR_INVALID_ARG, "step %d outside (0, 16*max_w]", step_bytes);
  MOD_CK(ctx, cudaSetDevice(ctx->device));
  rc = ensure_host_staging(ctx);
  if (rc != MOD_OK) return rc;
  // staging as in mod_image_crop: source in the first half of the cloud buffer, result in the second half
  cudaStream_t s = ctx->stream;
  ctx->cloud_w = ctx->cloud_h = 0;
  uint8_t* d_src = reinterpret_cast<uint8_t*>(ctx->d_points);
  uint8_t* d_out = d_src + 16 * static_cast<size_t>(ctx->max_w) * ctx->max_h;
  MOD_CK(ctx, ctx->stager.h2d(d_src, data, static_cast<size_t>(step_bytes) * h, s));
  rc = mod_image_to_mono8_batched_dev(ctx, s, w, h, 1, encoding, d_src, 0, step_bytes, rule, d_out);
  if (rc != MOD_OK) return rc;
  MOD_CK(ctx, ctx->stager.d2h(out_mono8, d_out, static_cast<size_t>(w) * h, s));
  MOD_CK(ctx, cudaStreamSynchronize(s));
  return MOD_OK;
}

int mod_host_crop_camera_info(uint32_t w, uint32_t h, int target_w, int target_h, const double* K9, const double* P12,
                              double* out_K9, double* out_P12) {
  int x0, y0;
  if (!K9 || !P12 || !out_K9 || !out_P12 || !crop_window(w, h, target_w, target_h, &x0, &y0)) return MOD_ERR_INVALID_ARG;
  for (int i = 0; i < 9; i++) out_K9[i] = K9[i];
  for (int i = 0; i < 12; i++) out_P12[i] = P12[i];
  // image_crop.cpp:32-33: double - (uint32 quotient), both principal-point entries taken from K
  const double cx = K9[2] - static_cast<double>(static_cast<uint32_t>(x0));
  const double cy = K9[5] - static_cast<double>(static_cast<uint32_t>(y0));
  out_K9[2] = cx;
  out_K9[5] = cy;
  out_P12[2] = cx;
  out_P12[6] = cy;
  return MOD_OK;
}

void mod_host_hsv_colormap(uint8_t* bgr768) {
  if (bgr768) host_hsv_colormap(bgr768);
}

int mod_pipeline_sequence_host(mod_ctx* ctx, int w, int h, int n_frames, const mod_camera* cam, float f, float T,
                               float min_disparity, float max_disparity, const float* disp_frames,
                               const float* flows, const mod_transform* transforms, const double* dt_sec,
                               const mod_sf_params* sf_params, const mod_cl_params* cl_params,
                               void* opt_out_points32, mod_object* objs, int32_t max_objs_per_pair,
                               mod_frame_stats* stats) {
  const int n_pairs = n_frames - 1;
  if (!ctx) return MOD_ERR_INVALID_ARG;
  if (n_pairs <= 0) return fail(ctx, MOD_ERR_INVALID_ARG, "need at least two frames");
  int rc = check_size(ctx, w, h, 1);
  if (rc != MOD_OK) return rc;
  if (!cam || !disp_frames || !flows || !transforms || !dt_sec || !sf_params || !cl_params || !objs || !stats ||
      max_objs_per_pair <= 0)
    return fail(ctx, MOD_ERR_INVALID_ARG, "null argument");
  MOD_CK(ctx, cudaSetDevice(ctx->device));
  const size_t P = static_cast<size_t>(w) * h;

  // chunk = as many pairs as the clusterer workspace holds, at most 8: large enough to amortise launches,
  // small enough that upload / compute / download of neighbouring chunks overlap.
  int chunk = ctx->max_batch < 8 ? ctx->max_batch : 8;
  if (ctx->seq_chunk == 0) {
    for (auto& b : ctx->seq) {  // only what an earlier, failed attempt did not already allocate
      if (!b.d_disp) MOD_CK(ctx, dev_alloc(&b.d_disp, ctx->maxP * (chunk + 1)));
      if (!b.d_flow) MOD_CK(ctx, dev_alloc(&b.d_flow, ctx->maxP * 2 * chunk));
      if (!b.d_points) MOD_CK(ctx, dev_alloc(&b.d_points, ctx->maxP * 8 * chunk));
      if (!b.d_mask) MOD_CK(ctx, dev_alloc(&b.d_mask, ctx->maxP * chunk));
      if (!b.d_stats) MOD_CK(ctx, dev_alloc(&b.d_stats, static_cast<size_t>(chunk)));
      if (!b.ev_in) MOD_CK(ctx, cudaEventCreateWithFlags(&b.ev_in, cudaEventDisableTiming));
      if (!b.ev_comp) MOD_CK(ctx, cudaEventCreateWithFlags(&b.ev_comp, cudaEventDisableTiming));
      if (!b.ev_out) MOD_CK(ctx, cudaEventCreateWithFlags(&b.ev_out, cudaEventDisableTiming));
    }
    ctx->seq_chunk = chunk;
  }
  chunk = ctx->seq_chunk;
  for (auto& b : ctx->seq) {
    if (b.max_objs < max_objs_per_pair) {
      cudaFree(b.d_objs);
      b.d_objs = nullptr;
      MOD_CK(ctx, dev_alloc(&b.d_objs, static_cast<size_t>(chunk) * max_objs_per_pair));
      b.max_objs = max_objs_per_pair;
    }
  }

  mod_disparity meta{f, T, min_disparity, max_disparity, nullptr, 0};
  std::vector<mod_pair_dev> pairs(chunk);
  int n_chunks = (n_pairs + chunk - 1) / chunk;
  for (int c = 0; c < n_chunks; c++) {
    mod_ctx::SeqBuf& b = ctx->seq[c & 1];
    const int first = c * chunk;
    const int n = (n_pairs - first < chunk) ? n_pairs - first : chunk;
    // the buffer is free once its previous download (chunk c-2) finished
    if (c >= 2) MOD_CK(ctx, cudaStreamWaitEvent(ctx->s_in, b.ev_out, 0));
    // slot 0 = frame `first`: uploaded for chunk 0, otherwise carried device-to-device from the previous chunk's
    // last slot.  The carry runs on the upload stream, i.e. after the previous chunk's upload and before the
    // next upload into that buffer, so no extra event is needed.
    if (c == 0) {
      MOD_CK(ctx, ctx->stager.h2d(b.d_disp, disp_frames, P * 4, ctx->s_in));
    } else {
      mod_ctx::SeqBuf& pb = ctx->seq[(c - 1) & 1];
      MOD_CK(ctx, cudaMemcpyAsync(b.d_disp, pb.d_disp + P * chunk, P * 4, cudaMemcpyDeviceToDevice, ctx->s_in));
    }
    // frames first+1 .. first+n go to slots 1..n (pinned sources: one DMA each; pageable ones: staged, hostcopy.cuh)
    MOD_CK(ctx, ctx->stager.h2d(b.d_disp + P, disp_frames + P * (first + 1), P * 4 * n, ctx->s_in));
    MOD_CK(ctx, ctx->stager.h2d(b.d_flow, flows + P * 2 * first, P * 8 * n, ctx->s_in));
    MOD_CK(ctx, cudaEventRecord(b.ev_in, ctx->s_in));

    MOD_CK(ctx, cudaStreamWaitEvent(ctx->stream, b.ev_in, 0));
    for (int k = 0; k < n; k++)
      pairs[k] = make_pair(b.d_disp + P * (k + 1), b.d_disp + P * k, b.d_flow + P * 2 * k, meta, meta,
                           transforms[first + k], dt_sec[first + k]);
    rc = run_scene_flow_dev(ctx, ctx->stream, w, h, n, *cam, pairs.data(), *sf_params, cl_params->dynamic_speed,
                            b.d_points, P, b.d_mask, nullptr, nullptr);
    if (rc != MOD_OK) return rc;
    rc = run_cluster_dev(ctx, ctx->stream, w, h, n, b.d_points, P, b.d_mask, *cl_params, nullptr, b.d_objs,
                         max_objs_per_pair, b.d_stats);
    if (rc != MOD_OK) return rc;
    MOD_CK(ctx, cudaEventRecord(b.ev_comp, ctx->stream));

    MOD_CK(ctx, cudaStreamWaitEvent(ctx->s_out, b.ev_comp, 0));
    if (opt_out_points32)
      MOD_CK(ctx, cudaMemcpyAsync(static_cast<char*>(opt_out_points32) + P * 32 * first, b.d_points, P * 32 * n,
                                  cudaMemcpyDeviceToHost, ctx->s_out));
    MOD_CK(ctx, cudaMemcpyAsync(objs + static_cast<size_t>(first) * max_objs_per_pair, b.d_objs,
                                sizeof(mod_object) * static_cast<size_t>(n) * max_objs_per_pair, cudaMemcpyDeviceToHost,
                                ctx->s_out));
    MOD_CK(ctx, cudaMemcpyAsync(stats + first, b.d_stats, sizeof(mod_frame_stats) * n, cudaMemcpyDeviceToHost, ctx->s_out));
    MOD_CK(ctx, cudaEventRecord(b.ev_out, ctx->s_out));
  }
  MOD_CK(ctx, cudaStreamSynchronize(ctx->s_out));
  MOD_CK(ctx, cudaStreamSynchronize(ctx->stream));
  int worst = MOD_OK;
  for (int i = 0; i < n_pairs; i++)
    if (stats[i].status != MOD_OK && worst == MOD_OK) worst = stats[i].status;
  if (worst != MOD_OK) return fail(ctx, worst, "a frame reported status %d", worst);
  return MOD_OK;
}

}  // extern "C"
