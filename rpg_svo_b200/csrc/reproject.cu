// rpg_svo_b200/csrc/reproject.cu -- C ABI entry point
//   svo_b200_reproject_map  <- Reprojector::reprojectMap (svo/src/reprojector.cpp:64-217)
//
// The reference walks the grid cell by cell and calls Matcher::findMatchDirect once per candidate until a cell has its
// match -- a data-dependent sequential loop.  Here the device projects every map point / candidate into the frame and
// aligns ALL in-frame points speculatively (one warp each: Point::getCloseViewObs over the observation list, then
// findMatchDirect), and the host replays the reference's policy -- cell order, quality sort, one match per cell, maxFts
// stop, point counters and deletions -- over those results, touching only what the sequential code would have reached
// (SURVEY.md 8f row 2).  Everything the replay needs is per-point and independent of the replay order, so the results
// equal the sequential ones.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "ctx.h"
#include "warp_align.cuh"

namespace svo {

constexpr int kRpWarps = 4;

struct ReprojIn {
  const int* e_pt;           // E: enumerated point indices (keyframe points in reference order, then candidates)
  const uint8_t* e_skip;     // E: input type == TYPE_DELETED -> projected but never matched
  const double* pt_pos;      // n_points*3
  const int* pt_obs_offset;  // n_points+1
  const int* pt_obs;
  const int* ftr_kf;
  const double* ftr_px;
  const double* ftr_f;
  const int* ftr_level;
  const int* ftr_type;
  const double* ftr_grad;
  const double* kf_T;        // n_kfs*12
  const FrameDesc* kf_frames;
};
struct ReprojOut {
  double* px;         // E*2  frame->w2c(pos)
  uint8_t* in_frame;  // E
  int* cell;          // E
  uint8_t* success;   // E    findMatchDirect
  double* px_match;   // E*2
  int* search_level;  // E
  double* A;          // E*4
  int* ref_ftr;       // E    feature picked by getCloseViewObs
};

__global__ void __launch_bounds__(kRpWarps * 32) reproject_match_kernel(
    FrameDesc cur, Cam cam, int E, int n_kfs, ReprojIn in, ReprojOut out, const double* __restrict__ cur_T_f_w, int cell_size,
    int grid_n_cols, int find_match, int max_search_level, int align_max_iter) {
  extern __shared__ double kf_pos[];  // n_kfs*3: Frame::pos() = T_f_w_.inverse().translation()
  __shared__ WarpAlignScratch scratch[kRpWarps];
  for (int k = threadIdx.x; k < n_kfs; k += blockDim.x) {
    const Pose Ti = pose_inv(pose_from_rt12(in.kf_T + 12 * (size_t)k));
    kf_pos[3 * k] = Ti.t[0]; kf_pos[3 * k + 1] = Ti.t[1]; kf_pos[3 * k + 2] = Ti.t[2];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * kRpWarps + warp;
  if (e >= E) return;
  const int p = in.e_pt[e];
  const double pos[3] = {in.pt_pos[3 * p], in.pt_pos[3 * p + 1], in.pt_pos[3 * p + 2]};
  const Pose T_cur_w = pose_from_rt12(cur_T_f_w);
  // Reprojector::reprojectPoint (:206-217)
  double pc[3], u, v;
  pose_apply(T_cur_w, pos, pc);
  world2cam(cam, pc, u, v);
  const int ui = (int)u, vi = (int)v;
  const bool inside = ui >= 8 && ui < cam.width - 8 && vi >= 8 && vi < cam.height - 8;  // isInFrame(px.cast<int>(), 8)
  const int cell = inside ? (int)(v / cell_size) * grid_n_cols + (int)(u / cell_size) : -1;
  int success = 0, search_level = 0, ref = -1;
  double A[4] = {0, 0, 0, 0}, h_inv = 0.0, pu = u, pv = v;
  if (inside && find_match && !in.e_skip[e]) {
    // Point::getCloseViewObs (point.cpp:97-117): first observation with the largest cos(angle), must be > 60 deg
    const Pose T_cur_w_inv = pose_inv(T_cur_w);
    double ox = T_cur_w_inv.t[0] - pos[0], oy = T_cur_w_inv.t[1] - pos[1], oz = T_cur_w_inv.t[2] - pos[2];
    const double on = sqrt(ox * ox + oy * oy + oz * oz);
    ox /= on; oy /= on; oz /= on;
    const int b = in.pt_obs_offset[p], en = in.pt_obs_offset[p + 1];
    double best_c = 0.0;
    int best_j = b;
    for (int j = b + lane; j < en; j += 32) {
      const int k = in.ftr_kf[in.pt_obs[j]];
      double dx = kf_pos[3 * k] - pos[0], dy = kf_pos[3 * k + 1] - pos[1], dz = kf_pos[3 * k + 2] - pos[2];
      const double dn = sqrt(dx * dx + dy * dy + dz * dz);
      dx /= dn; dy /= dn; dz /= dn;
      const double c = ox * dx + oy * dy + oz * dz;
      if (c > best_c) { best_c = c; best_j = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double oc = __shfl_xor_sync(0xffffffffu, best_c, o);
      const int oj = __shfl_xor_sync(0xffffffffu, best_j, o);
      if (oc > best_c || (oc == best_c && oj < best_j)) { best_c = oc; best_j = oj; }
    }
    if (en > b) {
      ref = in.pt_obs[best_j];
      if (!(best_c < 0.5)) {
        const int k = in.ftr_kf[ref];
        const double ref_px[2] = {in.ftr_px[2 * ref], in.ftr_px[2 * ref + 1]};
        const double f_ref[3] = {in.ftr_f[3 * ref], in.ftr_f[3 * ref + 1], in.ftr_f[3 * ref + 2]};
        const double grad[2] = {in.ftr_grad[2 * ref], in.ftr_grad[2 * ref + 1]};
        success = warp_find_match_direct(cur, cam, in.kf_frames[k], pose_from_rt12(in.kf_T + 12 * (size_t)k), T_cur_w, ref_px,
                                         f_ref, in.ftr_level[ref], in.ftr_type[ref], grad, pos, max_search_level,
                                         align_max_iter, scratch[warp], pu, pv, search_level, A, h_inv)
                      ? 1 : 0;
      }
    }
  }
  if (lane == 0) {
    out.px[2 * e] = u; out.px[2 * e + 1] = v;
    out.in_frame[e] = inside ? 1 : 0;
    out.cell[e] = cell;
    out.success[e] = (uint8_t)success;
    out.px_match[2 * e] = pu; out.px_match[2 * e + 1] = pv;
    out.search_level[e] = search_level;
    for (int k = 0; k < 4; ++k) out.A[4 * e + k] = A[k];
    out.ref_ftr[e] = ref;
  }
}

}  // namespace svo

using namespace svo;

extern "C" int svo_b200_reproject_map(svo_b200_ctx* ctx, const svo_b200_map_view* m, const svo_b200_frame* const* kf_frames,
                                      const svo_b200_frame* cur, const double* cur_T_f_w, const svo_b200_camera* cam,
                                      const svo_b200_reproject_options* opt, const int* cell_order, int* pt_type_io,
                                      int* pt_n_failed_io, int* pt_n_succeeded_io, uint8_t* pt_action_out,
                                      int* overlap_kf_out, int64_t* overlap_count_out, int* new_point_out,
                                      double* new_px_out, int* new_level_out, int* new_type_out, double* new_grad_out,
                                      svo_b200_reproject_stats* stats) {
  if (!ctx || !m || !cur || !cur_T_f_w || !cam || !opt || !cell_order || !pt_type_io || !pt_n_failed_io ||
      !pt_n_succeeded_io || !pt_action_out || !overlap_kf_out || !overlap_count_out || !new_point_out || !new_px_out ||
      !new_level_out || !new_type_out || !new_grad_out || !stats)
    return set_err(ctx, SVO_B200_EINVAL, "reproject_map: NULL argument");
  if (m->n_kfs < 0 || m->n_ftrs < 0 || m->n_points < 0 || m->n_candidates < 0 || (m->n_kfs > 0 && !kf_frames))
    return set_err(ctx, SVO_B200_EINVAL, "reproject_map: negative sizes / missing keyframe handles");
  if ((m->n_kfs > 0 && (!m->kf_T_f_w || !m->kf_keypt_pos || !m->kf_keypt_valid || !m->kf_fts_offset)) ||
      (m->n_ftrs > 0 && (!m->ftr_kf || !m->ftr_px || !m->ftr_f || !m->ftr_level || !m->ftr_type || !m->ftr_grad || !m->ftr_point)) ||
      (m->n_points > 0 && (!m->pt_pos || !m->pt_obs_offset)) || (m->n_candidates > 0 && !m->cand_point))
    return set_err(ctx, SVO_B200_EINVAL, "reproject_map: NULL array in the map view");
  for (int k = 0; k < m->n_kfs; ++k)
    if (!kf_frames[k]) return set_err(ctx, SVO_B200_EINVAL, "reproject_map: kf_frames[%d] is NULL", k);
  const int n_kf_fts = m->n_kfs ? m->kf_fts_offset[m->n_kfs] : 0, n_obs_total = m->n_points ? m->pt_obs_offset[m->n_points] : 0;
  if (n_kf_fts < 0 || n_obs_total < 0 || (n_kf_fts > 0 && !m->kf_fts) || (n_obs_total > 0 && !m->pt_obs))
    return set_err(ctx, SVO_B200_EINVAL, "reproject_map: inconsistent offsets in the map view");
  if (opt->grid_size <= 0 || opt->max_fts < 0 || opt->max_n_kfs < 0 || opt->max_search_level < 0 ||
      opt->max_search_level >= cur->n_levels)
    return set_err(ctx, SVO_B200_EINVAL, "reproject_map: bad options (grid_size %d, max_search_level %d of %d levels)",
                   opt->grid_size, opt->max_search_level, cur->n_levels);
  for (int i = 0; i < m->n_ftrs; ++i) {
    if (m->ftr_kf[i] < 0 || m->ftr_kf[i] >= m->n_kfs || m->ftr_point[i] >= m->n_points || m->ftr_point[i] < -1)
      return set_err(ctx, SVO_B200_EINVAL, "reproject_map: feature %d refers outside the map view", i);
    if (m->ftr_level[i] < 0 || m->ftr_level[i] >= kf_frames[m->ftr_kf[i]]->n_levels)
      return set_err(ctx, SVO_B200_EINVAL, "reproject_map: ftr_level[%d] outside the pyramid", i);
  }
  for (int j = 0; j < n_kf_fts; ++j)
    if (m->kf_fts[j] < 0 || m->kf_fts[j] >= m->n_ftrs) return set_err(ctx, SVO_B200_EINVAL, "reproject_map: kf_fts[%d] out of range", j);
  for (int j = 0; j < n_obs_total; ++j)
    if (m->pt_obs[j] < 0 || m->pt_obs[j] >= m->n_ftrs) return set_err(ctx, SVO_B200_EINVAL, "reproject_map: pt_obs[%d] out of range", j);
  for (int c = 0; c < m->n_candidates; ++c)
    if (m->cand_point[c] < 0 || m->cand_point[c] >= m->n_points) return set_err(ctx, SVO_B200_EINVAL, "reproject_map: cand_point[%d] out of range", c);

  std::memset(stats, 0, sizeof(*stats));
  for (int p = 0; p < m->n_points; ++p) pt_action_out[p] = SVO_B200_PT_NONE;
  // initializeGrid (:47-58)
  const int cell_size = opt->grid_size;
  Cam cm;
  { const int rc_cam = cam_to_dev(ctx, cam, cm); if (rc_cam) return rc_cam; }
  const int grid_n_cols = (int)std::ceil((double)cam->width / cell_size);
  const int grid_n_rows = (int)std::ceil((double)cam->height / cell_size);
  const size_t n_cells = (size_t)grid_n_cols * grid_n_rows;
  for (size_t i = 0; i < n_cells; ++i)
    if (cell_order[i] < 0 || (size_t)cell_order[i] >= n_cells) return set_err(ctx, SVO_B200_EINVAL, "reproject_map: cell_order[%zu] out of range", i);

  // Map::getCloseKeyframes (map.cpp:106-127) + sort by distance (:76-77); list::sort is stable
  const double* Tc = cur_T_f_w;
  std::vector<std::pair<int, double>> close_kfs;
  for (int k = 0; k < m->n_kfs; ++k)
    for (int i = 0; i < 5; ++i) {
      if (!m->kf_keypt_valid[5 * k + i]) continue;
      const double* kp = m->kf_keypt_pos + 3 * (5 * (size_t)k + i);
      const double x = Tc[0] * kp[0] + Tc[1] * kp[1] + Tc[2] * kp[2] + Tc[3];  // Frame::isVisible (frame.cpp:141-150)
      const double y = Tc[4] * kp[0] + Tc[5] * kp[1] + Tc[6] * kp[2] + Tc[7];
      const double z = Tc[8] * kp[0] + Tc[9] * kp[1] + Tc[10] * kp[2] + Tc[11];
      if (z < 0.0) continue;
      double u, v;
      cam_world2cam(cm, x / z, y / z, u, v);  // Frame::w2c
      if (u >= 0.0 && v >= 0.0 && u < cam->width && v < cam->height) {
        const double* Tk = m->kf_T_f_w + 12 * (size_t)k;
        const double dx = Tc[3] - Tk[3], dy = Tc[7] - Tk[7], dz = Tc[11] - Tk[11];
        close_kfs.emplace_back(k, std::sqrt(dx * dx + dy * dy + dz * dz));
        break;
      }
    }
  std::stable_sort(close_kfs.begin(), close_kfs.end(),
                   [](const std::pair<int, double>& a, const std::pair<int, double>& b) { return a.second < b.second; });

  // enumeration in the reference's order (:81-104, :108-127); src = overlap slot or -1 for a candidate
  std::vector<int> e_pt, e_src;
  std::vector<uint8_t> projected((size_t)m->n_points, 0);  // point->last_projected_kf_id_ == frame->id_
  size_t n_ov = 0;
  for (auto it = close_kfs.begin(); it != close_kfs.end() && n_ov < (size_t)opt->max_n_kfs; ++it, ++n_ov) {
    const int k = it->first;
    overlap_kf_out[n_ov] = k;
    overlap_count_out[n_ov] = 0;
    for (int j = m->kf_fts_offset[k]; j < m->kf_fts_offset[k + 1]; ++j) {
      const int p = m->ftr_point[m->kf_fts[j]];
      if (p < 0 || projected[p]) continue;
      projected[p] = 1;
      e_pt.push_back(p);
      e_src.push_back((int)n_ov);
    }
  }
  stats->n_overlap = (int)n_ov;
  for (int c = 0; c < m->n_candidates; ++c) { e_pt.push_back(m->cand_point[c]); e_src.push_back(-1); }
  const int E = (int)e_pt.size();
  if (E == 0) return 0;

  // ---- device: project + speculative getCloseViewObs / findMatchDirect for every enumerated point ----
  cudaSetDevice(ctx->device);
  const int n_obs = n_obs_total;
  Carver c;
  const size_t o_ept = c.take(sizeof(int) * E), o_skip = c.take(E), o_pos = c.take(sizeof(double) * 3 * m->n_points),
               o_ooff = c.take(sizeof(int) * (m->n_points + 1)), o_obs = c.take(sizeof(int) * n_obs),
               o_fkf = c.take(sizeof(int) * m->n_ftrs), o_fpx = c.take(sizeof(double) * 2 * m->n_ftrs),
               o_ff = c.take(sizeof(double) * 3 * m->n_ftrs), o_flv = c.take(sizeof(int) * m->n_ftrs),
               o_fty = c.take(sizeof(int) * m->n_ftrs), o_fgr = c.take(sizeof(double) * 2 * m->n_ftrs),
               o_kT = c.take(sizeof(double) * 12 * m->n_kfs), o_kfr = c.take(sizeof(FrameDesc) * m->n_kfs),
               o_cT = c.take(sizeof(double) * 12);
  const size_t in_bytes = c.off;
  const size_t o_px = c.take(sizeof(double) * 2 * E), o_in = c.take(E), o_cell = c.take(sizeof(int) * E), o_su = c.take(E),
               o_pm = c.take(sizeof(double) * 2 * E), o_sl = c.take(sizeof(int) * E), o_A = c.take(sizeof(double) * 4 * E),
               o_rf = c.take(sizeof(int) * E);
  int rc;
  if ((rc = ensure_host(ctx, ctx->h_in, c.off))) return rc;
  if ((rc = ensure_dev(ctx, ctx->d_in, c.off))) return rc;
  SVO_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  uint8_t* h = static_cast<uint8_t*>(ctx->h_in.p);
  uint8_t* d = static_cast<uint8_t*>(ctx->d_in.p);
  auto cp = [&](size_t off, const void* src, size_t bytes) { if (bytes) memcpy(h + off, src, bytes); };  // empty tables may be NULL
  cp(o_ept, e_pt.data(), sizeof(int) * E);
  for (int e = 0; e < E; ++e) (h + o_skip)[e] = pt_type_io[e_pt[e]] == 0;
  cp(o_pos, m->pt_pos, sizeof(double) * 3 * m->n_points);
  cp(o_ooff, m->pt_obs_offset, sizeof(int) * (m->n_points + 1));
  cp(o_obs, m->pt_obs, sizeof(int) * n_obs);
  cp(o_fkf, m->ftr_kf, sizeof(int) * m->n_ftrs);
  cp(o_fpx, m->ftr_px, sizeof(double) * 2 * m->n_ftrs);
  cp(o_ff, m->ftr_f, sizeof(double) * 3 * m->n_ftrs);
  cp(o_flv, m->ftr_level, sizeof(int) * m->n_ftrs);
  cp(o_fty, m->ftr_type, sizeof(int) * m->n_ftrs);
  cp(o_fgr, m->ftr_grad, sizeof(double) * 2 * m->n_ftrs);
  cp(o_kT, m->kf_T_f_w, sizeof(double) * 12 * m->n_kfs);
  for (int k = 0; k < m->n_kfs; ++k) reinterpret_cast<FrameDesc*>(h + o_kfr)[k] = make_desc(kf_frames[k]);
  cp(o_cT, cur_T_f_w, sizeof(double) * 12);
  SVO_CUDA_CHECK(ctx, cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
  ReprojIn in = {reinterpret_cast<const int*>(d + o_ept), d + o_skip, reinterpret_cast<const double*>(d + o_pos),
                 reinterpret_cast<const int*>(d + o_ooff), reinterpret_cast<const int*>(d + o_obs),
                 reinterpret_cast<const int*>(d + o_fkf), reinterpret_cast<const double*>(d + o_fpx),
                 reinterpret_cast<const double*>(d + o_ff), reinterpret_cast<const int*>(d + o_flv),
                 reinterpret_cast<const int*>(d + o_fty), reinterpret_cast<const double*>(d + o_fgr),
                 reinterpret_cast<const double*>(d + o_kT), reinterpret_cast<const FrameDesc*>(d + o_kfr)};
  ReprojOut out = {reinterpret_cast<double*>(d + o_px), d + o_in, reinterpret_cast<int*>(d + o_cell), d + o_su,
                   reinterpret_cast<double*>(d + o_pm), reinterpret_cast<int*>(d + o_sl), reinterpret_cast<double*>(d + o_A),
                   reinterpret_cast<int*>(d + o_rf)};
  const int blocks = (E + kRpWarps - 1) / kRpWarps;
  kt_begin(ctx);
  reproject_match_kernel<<<blocks, kRpWarps * 32, sizeof(double) * 3 * m->n_kfs, ctx->stream>>>(
      make_desc(cur), cm, E, m->n_kfs, in, out, reinterpret_cast<const double*>(d + o_cT), cell_size, grid_n_cols,
      opt->find_match_direct, opt->max_search_level, opt->align_max_iter);
  ctx->launches++;
  kt_end(ctx);
  SVO_CUDA_CHECK(ctx, cudaGetLastError());
  SVO_CUDA_CHECK(ctx, cudaMemcpyAsync(h + o_px, d + o_px, c.off - o_px, cudaMemcpyDeviceToHost, ctx->stream));
  SVO_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  const double* r_px = reinterpret_cast<const double*>(h + o_px);
  const uint8_t* r_in = h + o_in;
  const int* r_cell = reinterpret_cast<const int*>(h + o_cell);
  const uint8_t* r_su = h + o_su;
  const double* r_pm = reinterpret_cast<const double*>(h + o_pm);
  const int* r_sl = reinterpret_cast<const int*>(h + o_sl);
  const double* r_A = reinterpret_cast<const double*>(h + o_A);
  const int* r_rf = reinterpret_cast<const int*>(h + o_rf);
  (void)r_px;

  // ---- host replay of the sequential policy ----
  std::vector<std::vector<int>> cells(n_cells);  // enumeration indices, push_back order
  for (int e = 0; e < E; ++e) {
    const int p = e_pt[e];
    if (r_in[e]) {
      cells[(size_t)r_cell[e]].push_back(e);
      ++stats->n_projected;
      if (e_src[e] >= 0) overlap_count_out[e_src[e]]++;
      if (opt->find_match_direct && pt_type_io[p] != 0) ++stats->n_speculative;
    } else if (e_src[e] < 0) {  // candidate that does not reproject (:113-122)
      pt_n_failed_io[p] += 3;
      if (pt_n_failed_io[p] > 30) {
        pt_type_io[p] = 0;
        pt_action_out[p] = SVO_B200_PT_CANDIDATE_ERASED;
      }
    }
  }
  for (size_t i = 0; i < n_cells; ++i) {
    std::vector<int>& cell = cells[(size_t)cell_order[i]];
    // cell.sort(pointQualityComparator): stable, better type first (:144-149,153)
    std::stable_sort(cell.begin(), cell.end(), [&](int l, int r) { return pt_type_io[e_pt[l]] > pt_type_io[e_pt[r]]; });
    bool matched = false;
    for (size_t ci = 0; ci < cell.size(); ++ci) {
      const int e = cell[ci], p = e_pt[e];
      ++stats->n_trials;
      if (pt_type_io[p] == 0) continue;  // TYPE_DELETED: erased from the cell
      const bool found_match = opt->find_match_direct ? r_su[e] != 0 : true;
      if (!found_match) {
        pt_n_failed_io[p]++;
        if (pt_type_io[p] == 2 && pt_n_failed_io[p] > 15) { pt_type_io[p] = 0; pt_action_out[p] = SVO_B200_PT_SAFE_DELETE; }
        if (pt_type_io[p] == 1 && pt_n_failed_io[p] > 30) { pt_type_io[p] = 0; pt_action_out[p] = SVO_B200_PT_DELETE_CANDIDATE; }
        continue;
      }
      pt_n_succeeded_io[p]++;
      if (pt_type_io[p] == 2 && pt_n_succeeded_io[p] > 10) pt_type_io[p] = 3;
      const int q = stats->n_new++;
      new_point_out[q] = p;
      new_px_out[2 * q] = opt->find_match_direct ? r_pm[2 * e] : r_px[2 * e];
      new_px_out[2 * q + 1] = opt->find_match_direct ? r_pm[2 * e + 1] : r_px[2 * e + 1];
      new_level_out[q] = r_sl[e];
      new_type_out[q] = 0;
      new_grad_out[2 * q] = 1.0;
      new_grad_out[2 * q + 1] = 0.0;
      const int ref = r_rf[e];
      if (ref >= 0 && m->ftr_type[ref] == 1) {  // EDGELET: grad = normalize(A_cur_ref * ref grad) (:190-195)
        const double gx = m->ftr_grad[2 * ref], gy = m->ftr_grad[2 * ref + 1];
        const double ax = r_A[4 * e] * gx + r_A[4 * e + 1] * gy, ay = r_A[4 * e + 2] * gx + r_A[4 * e + 3] * gy;
        const double nn = std::sqrt(ax * ax + ay * ay);
        new_type_out[q] = 1;
        new_grad_out[2 * q] = ax / nn;
        new_grad_out[2 * q + 1] = ay / nn;
      }
      matched = true;
      break;
    }
    if (matched) ++stats->n_matches;
    if (stats->n_matches > (int64_t)opt->max_fts) break;
  }
  return 0;
}
