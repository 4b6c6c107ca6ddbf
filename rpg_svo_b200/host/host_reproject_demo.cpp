// rpg_svo_b200/host/host_reproject_demo.cpp -- drives svo::Reprojector of svo_host.h the way FrameHandlerMono does
// (frame_handler_mono.cpp:147-150: reprojector_.reprojectMap(new_frame_, overlap_kfs_)): rebuilds a map of keyframes,
// features, points and candidates from a binary dump written by tests/test_host_cpp_gpu.py, calls reprojectMap and writes
// the features added to the frame and the point bookkeeping for comparison with the CPU oracle.
//   usage: host_reproject_demo in.bin out.bin
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "svo_host.h"

template <class T>
static std::vector<T> rd(FILE* f, size_t n) {
  std::vector<T> v(n);
  if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "host_reproject_demo: short read\n"); exit(2); }
  return v;
}

int main(int argc, char** argv) {
  if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
  FILE* fi = fopen(argv[1], "rb");
  if (!fi) { perror("open input"); return 2; }
  const auto hdr = rd<int>(fi, 14);
  const int w = hdr[0], h = hdr[1], n_levels = hdr[2], n_kfs = hdr[3], F = hdr[4], P = hdr[5], C = hdr[6], n_kf_fts = hdr[7],
            n_obs = hdr[8], n_cells = hdr[9], grid = hdr[10], max_fts = hdr[11], max_n_kfs = hdr[12], n_pyr = hdr[13];
  const auto camv = rd<double>(fi, 4);
  const auto imgs = rd<uint8_t>(fi, (size_t)(n_kfs + 1) * w * h);  // keyframes, then the current frame
  const auto kf_T = rd<double>(fi, 12 * (size_t)n_kfs);
  const auto cur_T = rd<double>(fi, 12);
  const auto keypt_ftr = rd<int>(fi, 5 * (size_t)n_kfs);  // feature index or -1
  const auto kf_fts_offset = rd<int>(fi, n_kfs + 1);
  const auto kf_fts = rd<int>(fi, n_kf_fts);
  const auto ftr_kf = rd<int>(fi, F);
  const auto ftr_px = rd<double>(fi, 2 * (size_t)F);
  const auto ftr_f = rd<double>(fi, 3 * (size_t)F);
  const auto ftr_level = rd<int>(fi, F);
  const auto ftr_type = rd<int>(fi, F);
  const auto ftr_grad = rd<double>(fi, 2 * (size_t)F);
  const auto ftr_point = rd<int>(fi, F);
  const auto pt_pos = rd<double>(fi, 3 * (size_t)P);
  const auto pt_obs_offset = rd<int>(fi, P + 1);
  const auto pt_obs = rd<int>(fi, n_obs);
  const auto cand_point = rd<int>(fi, C);
  const auto pt_type = rd<int>(fi, P);
  const auto pt_failed = rd<int>(fi, P);
  const auto pt_succ = rd<int>(fi, P);
  const auto cell_order = rd<int>(fi, n_cells);
  fclose(fi);

  try {
    svo::Context ctx(0);
    svo::PinholeCamera cam(w, h, camv[0], camv[1], camv[2], camv[3]);
    svo::Map map;
    std::vector<svo::FramePtr> kfs;
    for (int k = 0; k < n_kfs; ++k) {
      kfs.emplace_back(new svo::Frame(ctx, &cam, imgs.data() + (size_t)k * w * h, n_levels, k));
      std::memcpy(kfs[k]->T_f_w_.m, &kf_T[12 * k], sizeof(double) * 12);
      kfs[k]->setKeyframe();
      map.addKeyframe(kfs[k]);
    }
    svo::FramePtr cur(new svo::Frame(ctx, &cam, imgs.data() + (size_t)n_kfs * w * h, n_levels, n_kfs));
    std::memcpy(cur->T_f_w_.m, cur_T.data(), sizeof(double) * 12);
    std::vector<svo::Point*> pts(P);
    for (int p = 0; p < P; ++p) {
      pts[p] = new svo::Point({pt_pos[3 * p], pt_pos[3 * p + 1], pt_pos[3 * p + 2]});
      pts[p]->type_ = (svo::Point::PointType)pt_type[p];
      pts[p]->n_failed_reproj_ = pt_failed[p];
      pts[p]->n_succeeded_reproj_ = pt_succ[p];
    }
    std::vector<svo::Feature*> fts(F);
    for (int i = 0; i < F; ++i) {
      fts[i] = new svo::Feature(kfs[ftr_kf[i]].get(), ftr_point[i] >= 0 ? pts[ftr_point[i]] : nullptr, {ftr_px[2 * i], ftr_px[2 * i + 1]},
                                {ftr_f[3 * i], ftr_f[3 * i + 1], ftr_f[3 * i + 2]}, ftr_level[i]);
      fts[i]->type = ftr_type[i] ? svo::Feature::EDGELET : svo::Feature::CORNER;
      fts[i]->grad = {ftr_grad[2 * i], ftr_grad[2 * i + 1]};
    }
    for (int k = 0; k < n_kfs; ++k) {
      for (int j = kf_fts_offset[k]; j < kf_fts_offset[k + 1]; ++j) kfs[k]->addFeature(fts[kf_fts[j]]);
      for (int i = 0; i < 5; ++i) kfs[k]->key_pts_[i] = keypt_ftr[5 * k + i] >= 0 ? fts[keypt_ftr[5 * k + i]] : nullptr;
    }
    for (int p = 0; p < P; ++p)
      for (int j = pt_obs_offset[p]; j < pt_obs_offset[p + 1]; ++j) pts[p]->obs_.push_back(fts[pt_obs[j]]);
    for (int c = 0; c < C; ++c)
      map.point_candidates_.candidates_.push_back(svo::MapPointCandidates::PointCandidate(pts[cand_point[c]], pts[cand_point[c]]->obs_.front()));

    svo::Reprojector::Options opt;
    opt.grid_size = grid; opt.max_fts = max_fts; opt.max_n_kfs = (size_t)max_n_kfs; opt.n_pyr_levels = n_pyr;
    svo::Reprojector reprojector(&cam, map, opt);
    reprojector.cellOrder() = cell_order;
    std::vector<std::pair<svo::FramePtr, size_t>> overlap_kfs;
    reprojector.reprojectMap(cur, overlap_kfs);

    FILE* fo = fopen(argv[2], "wb");
    if (!fo) { perror("open output"); return 2; }
    const long long head[4] = {(long long)reprojector.n_matches_, (long long)reprojector.n_trials_, (long long)cur->fts_.size(),
                               (long long)overlap_kfs.size()};
    fwrite(head, sizeof(long long), 4, fo);
    for (auto& o : overlap_kfs) {
      long long rec[2] = {-1, (long long)o.second};
      for (int k = 0; k < n_kfs; ++k) if (kfs[k] == o.first) rec[0] = k;
      fwrite(rec, sizeof(long long), 2, fo);
    }
    for (svo::Feature* f : cur->fts_) {
      int pi = -1;
      for (int p = 0; p < P; ++p) if (pts[p] == f->point) pi = p;
      const int rec[3] = {pi, f->level, (int)f->type};
      fwrite(rec, sizeof(int), 3, fo);
      const double d[4] = {f->px[0], f->px[1], f->grad[0], f->grad[1]};
      fwrite(d, sizeof(double), 4, fo);
    }
    for (int p = 0; p < P; ++p) {
      const int rec[3] = {(int)pts[p]->type_, pts[p]->n_failed_reproj_, pts[p]->n_succeeded_reproj_};
      fwrite(rec, sizeof(int), 3, fo);
    }
    fclose(fo);
    printf("host_reproject_demo: %zu matches in %zu trials, %zu overlap keyframes, %zu points in the map trash, %zu candidates in trash\n",
           reprojector.n_matches_, reprojector.n_trials_, overlap_kfs.size(), map.trash_points_.size(),
           map.point_candidates_.trash_points_.size());
    // ownership as in the reference: the map deletes its trash and remaining candidates; frames delete their features
    std::vector<bool> map_owned(P, false);
    for (svo::Point* t : map.trash_points_) for (int p = 0; p < P; ++p) if (pts[p] == t) map_owned[p] = true;
    for (svo::Point* t : map.point_candidates_.trash_points_) for (int p = 0; p < P; ++p) if (pts[p] == t) map_owned[p] = true;
    for (auto& c : map.point_candidates_.candidates_) for (int p = 0; p < P; ++p) if (pts[p] == c.first) map_owned[p] = true;
    map.emptyTrash();
    for (int p = 0; p < P; ++p) if (!map_owned[p]) delete pts[p];
  } catch (const std::exception& e) {
    fprintf(stderr, "host_reproject_demo: %s\n", e.what());
    return 1;
  }
  return 0;
}
