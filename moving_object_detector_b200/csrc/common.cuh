// common.cuh -- shared declarations of the sm_100a kernels behind include/mod_b200.h
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mod_b200.h"

namespace mod {

// ---------------------------------------------------------------------------------------------------------
// Scene-flow launch descriptor.  It is passed BY VALUE as a __grid_constant__ kernel parameter, so the per-pair
// ego-motion transforms, timestamps and disparity scalars live in the constant bank (no global loads, no
// cudaMemcpyToSymbol ordering hazards, CUDA-graph friendly).
// ---------------------------------------------------------------------------------------------------------
struct SfPair {
  const float* disp_now;   // dense 32FC1, device
  const float* disp_prev;  // dense 32FC1, device
  const float* flow;       // dense 32FC2, device
  double dt;               // (stamp_now - stamp_prev).toSec()
  // Hot per-pair scalars in 16-byte groups (one constant-bank load each).  The disparity tests of the reference
  // (getDisparity's range test, getPoint3D's d != 0, getRightPoint's NaN / inf / negative test) are folded into
  // closed intervals prepared on the host (abi.cu, fill_disparity_bounds):
  //   "gives a point"  <=>  lo_pt <= d <= hi_pt  and  d != 0        (NaN fails both comparisons)
  //   "usable in the chain" (range test and 0 <= d < inf)  <=>  lo_ch <= d <= hi_ch
  float4 g0;               // fT_now (f*T in float, disparity_image_processor.cpp:43), lo_pt_now, hi_pt_now, lo_ch_now
  float4 g1;               // hi_ch_now, fT_prev, lo_pt_prev, hi_pt_prev
  float4 g2;               // lo_ch_prev, hi_ch_prev, d_safe, rho
  float4 r0, r1, r2;       // float copies of (R[3i], R[3i+1], R[3i+2], t[i]) for the static-flow filter
  float tau;               // >= max |t_i|; rho (g2.w) >= max |R_ij|; rho = +inf disables the filter for the pair
  float pad[3];
  double R[9];             // row-major rotation, Eigen::Quaterniond::toRotationMatrix order of operations
  double t[3];
};
// d_safe (g2.z): previous disparities >= d_safe are known to ego-transform to a finite, non-NaN x

constexpr int kSfMaxPairsPerLaunch = 64;  // 64 * 240 B = 15 KB of the 32 KB parameter space

struct SfLaunch {
  SfPair pair[kSfMaxPairsPerLaunch];
  const double* ray_x;  // [w]  ((double)u - cx - Tx) / fx   (PinholeCameraModel::projectPixelTo3dRay)
  const double* ray_y;  // [h]  ((double)v - cy - Ty) / fy
  const float* ray_xf;  // float roundings of the two tables (static-flow filter)
  const float* ray_yf;
  double fx, fy, cx, cy, Tx, Ty;
  float fxf, fyf, cxf, cyf, Txf, Tyf;  // float roundings of the camera (static-flow filter)
  float aN, bN;          // error of the projected numerator: dN' = aN * S + bN
  float Cw;              // >= |cx| + |cy| + w + h
  int prefetch_rows;     // rows ahead whose inputs each thread pulls into L2 (set by launch_scene_flow; 0 = off)
  int use_filter;        // 0: evaluate the static flow exactly for every pixel
  unsigned long long* filter_fallbacks;  // pixels the filter could not decide (re-evaluated exactly); may be null
  int w, h;
  float flow_thr2;       // sqrtf(s) >= (float)dynamic_flow_diff  <=>  s >= flow_thr2   (sq_threshold)
  float speed_thr2;      // (double)sqrtf(s) >= dynamic_speed     <=>  s >= speed_thr2
  float* points;         // 8 floats per point
  size_t points_frame_stride;  // in points
  uint8_t* mask;         // may be null
  float* depth;          // may be null
  float* static_flow;    // may be null
  int first_frame;       // output frame index of pair[0]
};

constexpr int kClMaxNeighborDistance = 10;  // cfg/Clusterer.cfg:11 (neighbor_distance in [1, 10]); sizes the halos
constexpr int kClMaxClusters = 32768;       // surviving clusters per frame a context can be created for (cl_rank's
                                            // shared-memory key table: 128 KB)

// Clusterer launch descriptor (one per batched call; every kernel grid-strides over the batch's flat pixel range).
struct ClLaunch {
  const float* points;   // 8 floats per point
  size_t points_frame_stride;  // in points
  const uint8_t* mask;   // dynamic map, 1 byte per pixel, dense per frame
  int w, h, batch;
  int n;                 // neighbor_distance
  float depth_thr;       // largest float <= depth_diff   ((double)a > th  <=>  a > th_down)
  float speed_thr;       // smallest float >= dynamic_speed
  int min_size;          // max(cluster_size, 2): isolated dynamic pixels never get a label in the reference
  int max_clusters;      // survivor slots per frame
  int max_objs;
  // per-pixel workspaces, batch * w*h
  int* parent;
  int* count;            // at tile roots: component size (cl_tile, merged by cl_roots); at final roots after
                         // cl_slots: survivor slot or -1
  struct TileStat* tstat;  // at tile roots: partial statistics (cl_tile), merged into the final root by cl_roots
  // per-frame workspaces
  int* n_slots;          // [batch]
  int* n_dynamic;        // [batch]
  struct ClusterStat* stats;   // [batch * max_clusters]
  // outputs
  int32_t* labels;       // may be null
  mod_object* objs;
  mod_frame_stats* frame_stats;
};

struct __align__(32) TileStat {
  int key;           // min raster index of a member with an accepted up-left neighbour
  uint32_t mn[3];    // order-preserving encodings of float min x,y,z
  uint32_t mx[3];
  int pad;
};

struct ClusterStat {
  int root;          // pixel index of the union-find root
  int count;         // number of member pixels
  int key;           // min raster index of a member that has an accepted up-left neighbour (reference label order)
  int rank;          // final cluster id (rank of key among the frame's survivors)
  uint32_t mn[3];    // order-preserving encodings of float min x,y,z
  uint32_t mx[3];
  int median_idx;       // pixel index whose speed equals the selected value (smallest index among ties)
  int pad;
};

void launch_ray_lut(cudaStream_t s, double* d_ray_x, double* d_ray_y, float* d_ray_xf, float* d_ray_yf, int w, int h,
                    const mod_camera& cam);
// follows_own_launch: the previous operation on the stream is another launch of the same batched call (disjoint frames)
cudaError_t launch_scene_flow(cudaStream_t s, const SfLaunch& L, int n_pairs, bool follows_own_launch = false);
cudaError_t launch_ddiv_check(cudaStream_t s, const double* a, const double* b, size_t n, unsigned long long* n_bad);
cudaError_t launch_dynamic_map(cudaStream_t s, const float* points, size_t frame_stride_points, uint8_t* mask, int w,
                               int h, int batch, float speed_thr);
// Workspaces of the clusterer that are not indexed per pixel in ClLaunch.
struct ClExtra {
  int* work_count;         // [5]   work_count, work_cursor, list_count, border_count, root_count (one reset)
  int* work_cursor;        // = work_count + 1
  uint32_t* list;          // [batch * w*h]  compacted global indices (frame*P + i) of the dynamic pixels
  uint32_t* list_speed;    // [batch * w*h]  |v| bits of the same entries
  uint32_t* border_mask;   // [batch * w*h]  per border-list entry: window cells of its cross-tile edges (N <= 4)
  int2* work_items;        // [batch * max_clusters]  (frame, slot) of every surviving cluster
  int* member_offset;      // [batch * max_clusters]  start of the cluster's member list inside its frame
  int* cursor;             // [batch * max_clusters]  fill level of the member list
  uint32_t* member_speed;  // [batch * w*h]
  int* member_pixel;       // [batch * w*h]
};
// Returns the number of kernels it launched through *n_launches.
cudaError_t launch_cluster(cudaStream_t s, const ClLaunch& L, const ClExtra& X, int* n_launches);

// ~clusters_image: labels (+ per-frame n_clusters from stats, or fixed_n when stats is null) -> BGR8
cudaError_t launch_clusters_image(cudaStream_t s, const int32_t* labels, const mod_frame_stats* stats, int fixed_n,
                                  int w, int h, int batch, uint8_t* out_bgr);

// glue.cu: image_crop (image_crop/src/image_crop.cpp:20-43).  encoding: MOD_ENC_* of include/mod_b200.h
cudaError_t launch_image_crop(cudaStream_t s, const uint8_t* src, size_t src_frame_stride, int step, int encoding,
                              int x0, int y0, int tw, int th, int batch, uint8_t* out);
// glue.cu: cv_bridge::toCvCopy(image, MONO8) (scene_flow_constructor.cpp:220-221).  rule: MOD_GRAY_* of include/mod_b200.h
cudaError_t launch_to_mono8(cudaStream_t s, const uint8_t* src, size_t src_frame_stride, int step, int encoding, int w,
                            int h, int batch, int rule, uint8_t* out);
void host_hsv_colormap(uint8_t* bgr768);

// float threshold equivalents of the reference's float-vs-double comparisons
float float_ceil_of_double(double th);   // smallest float >= th
float float_floor_of_double(double th);  // largest float <= th
float sq_threshold(float thr);           // smallest float s2 with sqrtf(s2) >= thr

}  // namespace mod
