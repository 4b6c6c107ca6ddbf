// oracle/ref_wrap_reproject.cpp -- TEST INFRASTRUCTURE ONLY.
// C entry point around the reference's OWN svo::Reprojector + svo::Map (svo/src/reprojector.cpp, map.cpp compiled where
// they lie, see oracle/Makefile target `ref`): rebuilds the pointer graph (keyframes, features, points, candidates) from
// the flat map view of oracle/svo_oracle.h, runs Reprojector::reprojectMap and flattens the result again.
#include <svo/config.h>
#include <svo/feature.h>
#include <svo/frame.h>
#include <svo/map.h>
#include <svo/point.h>
#define private public  // grid_.cell_order is private; the reference shuffles it with rand(), the tests must fix it
#include <svo/reprojector.h>
#undef private
#include <vikit/abstract_camera.h>

#include <chrono>
#include <map>
#include <memory>
#include <set>
#include <vector>

#include "svo_oracle.h"

vk::AbstractCamera* ref_make_camera(int w, int h, const double* c);  // oracle/ref_wrap.cpp
extern double g_ref_last_seconds;                                       // oracle/ref_wrap.cpp

using namespace svo;

namespace {
SE3 se3_from12(const double* T) {
  Matrix3d R;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R(i, j) = T[i * 4 + j];
  return SE3(R, Vector3d(T[3], T[7], T[11]));
}
FramePtr make_frame(vk::AbstractCamera* cam, const uint8_t* img, int w, int h, int n_levels, const double* T_f_w) {
  Config::nPyrLevels() = 1;
  Config::kltMaxLevel() = n_levels - 1;
  cv::Mat m(h, w, const_cast<uint8_t*>(img), (size_t)w);
  FramePtr f(new Frame(cam, m, 0.0));
  f->T_f_w_ = se3_from12(T_f_w);
  return f;
}
}  // namespace

extern "C" void ref_reproject_map(const orc_map_view* m, const uint8_t* kf_l0s /*n_kfs images*/, const uint8_t* cur_l0, int w, int h,
                                  int n_levels, const double* cam4, const double* cur_T_f_w, const orc_reproject_options* opt,
                                  const int* cell_order, int* pt_type_io, int* pt_n_failed_io, int* pt_n_succeeded_io,
                                  uint8_t* pt_action_out, int* overlap_kf_out, int64_t* overlap_count_out, int* new_point,
                                  double* new_px, int* new_level, int* new_type, double* new_grad, orc_reproject_stats* st) {
  std::unique_ptr<vk::AbstractCamera> cam_owner(ref_make_camera(w, h, cam4));
  vk::AbstractCamera& cam = *cam_owner;
  std::vector<FramePtr> kfs;
  for (int k = 0; k < m->n_kfs; ++k) kfs.push_back(make_frame(&cam, kf_l0s + (size_t)k * w * h, w, h, n_levels, m->kf_T_f_w + 12 * k));
  FramePtr cur = make_frame(&cam, cur_l0, w, h, n_levels, cur_T_f_w);
  Config::nPyrLevels() = opt->max_search_level + 1;
  Config::gridSize() = opt->grid_size;
  Config::maxFts() = opt->max_fts;

  std::vector<Point*> pts(m->n_points);
  std::map<Point*, int> pt_index;
  for (int p = 0; p < m->n_points; ++p) {
    pts[p] = new Point(Vector3d(m->pt_pos[3 * p], m->pt_pos[3 * p + 1], m->pt_pos[3 * p + 2]));
    pts[p]->type_ = (Point::PointType)pt_type_io[p];
    pts[p]->n_failed_reproj_ = pt_n_failed_io[p];
    pts[p]->n_succeeded_reproj_ = pt_n_succeeded_io[p];
    pt_index[pts[p]] = p;
  }
  std::vector<Feature*> fts(m->n_ftrs);
  std::vector<bool> owned_by_frame(m->n_ftrs, false);
  for (int i = 0; i < m->n_ftrs; ++i) {
    Frame* fr = kfs[m->ftr_kf[i]].get();
    Point* p = m->ftr_point[i] >= 0 ? pts[m->ftr_point[i]] : NULL;
    fts[i] = new Feature(fr, p, Vector2d(m->ftr_px[2 * i], m->ftr_px[2 * i + 1]),
                         Vector3d(m->ftr_f[3 * i], m->ftr_f[3 * i + 1], m->ftr_f[3 * i + 2]), m->ftr_level[i]);
    fts[i]->type = m->ftr_type[i] ? Feature::EDGELET : Feature::CORNER;
    fts[i]->grad = Vector2d(m->ftr_grad[2 * i], m->ftr_grad[2 * i + 1]);
  }
  for (int k = 0; k < m->n_kfs; ++k)
    for (int j = m->kf_fts_offset[k]; j < m->kf_fts_offset[k + 1]; ++j) {
      kfs[k]->addFeature(fts[m->kf_fts[j]]);
      owned_by_frame[m->kf_fts[j]] = true;
    }
  for (int p = 0; p < m->n_points; ++p)
    for (int j = m->pt_obs_offset[p]; j < m->pt_obs_offset[p + 1]; ++j) pts[p]->obs_.push_back(fts[m->pt_obs[j]]);
  // key points: stand-in features carrying only point->pos_ (all Map::getCloseKeyframes reads)
  std::vector<std::unique_ptr<Point>> key_points;
  std::vector<std::unique_ptr<Feature>> key_features;
  for (int k = 0; k < m->n_kfs; ++k)
    for (int i = 0; i < 5; ++i) {
      kfs[k]->key_pts_[i] = NULL;
      if (!m->kf_keypt_valid[5 * k + i]) continue;
      const double* kp = m->kf_keypt_pos + 3 * (5 * (size_t)k + i);
      key_points.emplace_back(new Point(Vector3d(kp[0], kp[1], kp[2])));
      key_features.emplace_back(new Feature(kfs[k].get(), key_points.back().get(), Vector2d(0, 0), Vector3d(0, 0, 1), 0));
      kfs[k]->key_pts_[i] = key_features.back().get();
    }

  {
    Map map;
    for (auto& kf : kfs) map.addKeyframe(kf);
    for (int c = 0; c < m->n_candidates; ++c) {
      Point* p = pts[m->cand_point[c]];
      map.point_candidates_.candidates_.push_back(MapPointCandidates::PointCandidate(p, p->obs_.front()));
    }
    Reprojector rp(&cam, map);
    rp.options_.max_n_kfs = (size_t)opt->max_n_kfs;
    rp.options_.find_match_direct = opt->find_match_direct != 0;
    for (size_t i = 0; i < rp.grid_.cell_order.size(); ++i) rp.grid_.cell_order[i] = cell_order[i];
    std::vector<std::pair<FramePtr, size_t>> overlap;
    {
      const auto t0 = std::chrono::steady_clock::now();
      rp.reprojectMap(cur, overlap);
      g_ref_last_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }

    memset(st, 0, sizeof(*st));
    st->n_matches = (int64_t)rp.n_matches_;
    st->n_trials = (int64_t)rp.n_trials_;
    st->n_overlap = (int)overlap.size();
    for (size_t i = 0; i < overlap.size(); ++i) {
      for (int k = 0; k < m->n_kfs; ++k) if (kfs[k] == overlap[i].first) overlap_kf_out[i] = k;
      overlap_count_out[i] = (int64_t)overlap[i].second;
    }
    st->n_projected = -1;  // not observable after the cells have been consumed
    for (Feature* f : cur->fts_) {
      const int q = st->n_new++;
      new_point[q] = pt_index.at(f->point);
      new_px[2 * q] = f->px[0]; new_px[2 * q + 1] = f->px[1];
      new_level[q] = f->level;
      new_type[q] = f->type == Feature::EDGELET ? 1 : 0;
      new_grad[2 * q] = f->grad[0]; new_grad[2 * q + 1] = f->grad[1];
    }
    for (int p = 0; p < m->n_points; ++p) {
      pt_type_io[p] = (int)pts[p]->type_;
      pt_n_failed_io[p] = pts[p]->n_failed_reproj_;
      pt_n_succeeded_io[p] = pts[p]->n_succeeded_reproj_;
      pt_action_out[p] = ORC_PT_NONE;
    }
    for (Point* p : map.trash_points_) pt_action_out[pt_index.at(p)] = ORC_PT_SAFE_DELETE;
    for (Point* p : map.point_candidates_.trash_points_) pt_action_out[pt_index.at(p)] = ORC_PT_DELETE_CANDIDATE;  // or _ERASED
    // ownership: ~Map deletes its trash and the remaining candidates (points and their features); the rest is ours
    std::set<Point*> map_owned(map.trash_points_.begin(), map.trash_points_.end());
    map_owned.insert(map.point_candidates_.trash_points_.begin(), map.point_candidates_.trash_points_.end());
    for (auto& c : map.point_candidates_.candidates_) map_owned.insert(c.first);
    for (int p = 0; p < m->n_points; ++p)
      if (map_owned.count(pts[p])) pts[p] = NULL;
  }
  // features: frames delete their fts_; candidate features are deleted by the map; unreferenced ones are ours
  std::set<int> cand_pts(m->cand_point, m->cand_point + m->n_candidates);
  for (int i = 0; i < m->n_ftrs; ++i)
    if (!owned_by_frame[i] && !(m->ftr_point[i] >= 0 && cand_pts.count(m->ftr_point[i]))) delete fts[i];
  for (Point* p : pts) delete p;
}
