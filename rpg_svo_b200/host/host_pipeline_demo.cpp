// rpg_svo_b200/host/host_pipeline_demo.cpp -- executes the rest of the C++ host surface of svo_host.h the way the
// reference's own code calls it:
//   units     feature_alignment::align2D / align1D with the reference's argument lists (feature_alignment.h:29-44),
//             svo::Matcher::findMatchDirect / findEpipolarMatchDirect + the public scratch members their callers read
//             (matcher.h:92-121), DepthFilter::addKeyframe / addFrame -> updateSeeds over several frames until seeds
//             converge into MapPointCandidates::newCandidatePoint (depth_filter.cpp:197-291, map.cpp:213-218),
//             getSeedsCopy, the halt flag;
//   pipeline  svo::FrameHandlerMono::addImage (frame_handler_mono.cpp:129-245): SparseImgAlign -> Reprojector ->
//             pose_optimizer -> Point::optimize on the tracking thread's context while the depth filter runs on the
//             mapper thread with a context of its own (two host threads, two CUDA streams, concurrently).
// Inputs come from a binary dump written by tests/test_host_cpp_gpu.py, results go to a second file that the test compares
// with the CPU oracle.   usage: host_pipeline_demo units|pipeline in.bin out.bin
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "svo_host.h"

template <class T>
static void rd(FILE* f, T* p, size_t n) {
  if (n && fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "host_pipeline_demo: short read\n"); exit(2); }
}
template <class T>
static void wr(FILE* f, const T* p, size_t n) { fwrite(p, sizeof(T), n, f); }

static svo::AbstractCamera* read_camera(FILE* fi) {
  int wh[3];
  rd(fi, wh, 3);  // width, height, model
  double c[9];
  rd(fi, c, 9);   // fx fy cx cy d0..d4
  if (wh[2] == SVO_B200_CAM_ATAN)  // pixel parameters back to the normalised ones the vikit constructor takes
    return new svo::ATANCamera(wh[0], wh[1], c[0] / wh[0], c[1] / wh[1], (c[2] + 0.5) / wh[0], (c[3] + 0.5) / wh[1], c[4]);
  return new svo::PinholeCamera(wh[0], wh[1], c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8]);
}

static int run_units(FILE* fi, FILE* fo) {
  std::unique_ptr<svo::AbstractCamera> cam(read_camera(fi));
  const int w = cam->width(), h = cam->height();
  int hdr[6];
  rd(fi, hdr, 6);
  const int n_levels = hdr[0], M = hdr[1], E = hdr[2], A = hdr[3], S = hdr[4], F = hdr[5];
  std::vector<uint8_t> ref_img((size_t)w * h), cur_img((size_t)w * h);
  rd(fi, ref_img.data(), ref_img.size());
  rd(fi, cur_img.data(), cur_img.size());
  svo::Context ctx(0);
  svo::FramePtr ref(new svo::Frame(ctx, cam.get(), ref_img.data(), n_levels, 0.0));
  svo::FramePtr cur(new svo::Frame(ctx, cam.get(), cur_img.data(), n_levels, 1.0));
  rd(fi, ref->T_f_w_.m, 12);
  rd(fi, cur->T_f_w_.m, 12);

  // ---- Matcher::findMatchDirect: one Point with one observation (the reference feature) per candidate
  std::vector<svo::Point*> points;
  {
    std::vector<double> px(2 * M), f(3 * M), grad(2 * M), pos(3 * M), px_cur(2 * M);
    std::vector<int> level(M), type(M);
    rd(fi, px.data(), px.size()); rd(fi, f.data(), f.size()); rd(fi, level.data(), level.size()); rd(fi, type.data(), type.size());
    rd(fi, grad.data(), grad.size()); rd(fi, pos.data(), pos.size()); rd(fi, px_cur.data(), px_cur.size());
    svo::Matcher matcher;
    for (int i = 0; i < M; ++i) {
      svo::Feature* ftr = new svo::Feature(ref.get(), nullptr, {px[2 * i], px[2 * i + 1]}, {f[3 * i], f[3 * i + 1], f[3 * i + 2]}, level[i]);
      ftr->type = type[i] ? svo::Feature::EDGELET : svo::Feature::CORNER;
      ftr->grad = {grad[2 * i], grad[2 * i + 1]};
      ref->addFeature(ftr);
      svo::Point* pt = new svo::Point({pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]}, ftr);
      ftr->point = pt;
      points.push_back(pt);
      svo::Vector2d pc{px_cur[2 * i], px_cur[2 * i + 1]};
      const int ok = matcher.findMatchDirect(*pt, *cur, pc) ? 1 : 0;
      const double rec[8] = {(double)ok, pc[0], pc[1], (double)matcher.search_level_, matcher.A_cur_ref_[0], matcher.A_cur_ref_[1],
                             matcher.A_cur_ref_[2], matcher.A_cur_ref_[3]};
      wr(fo, rec, 8);
      if (matcher.ref_ftr_ != ftr) { fprintf(stderr, "findMatchDirect: ref_ftr_ is not the point's observation\n"); return 1; }
    }
  }
  // ---- Matcher::findEpipolarMatchDirect
  {
    std::vector<double> px(2 * E), f(3 * E), grad(2 * E), d(3 * E);
    std::vector<int> level(E), type(E);
    rd(fi, px.data(), px.size()); rd(fi, f.data(), f.size()); rd(fi, level.data(), level.size()); rd(fi, type.data(), type.size());
    rd(fi, grad.data(), grad.size()); rd(fi, d.data(), d.size());
    svo::Matcher matcher;
    for (int i = 0; i < E; ++i) {
      svo::Feature ftr(ref.get(), nullptr, {px[2 * i], px[2 * i + 1]}, {f[3 * i], f[3 * i + 1], f[3 * i + 2]}, level[i]);
      ftr.type = type[i] ? svo::Feature::EDGELET : svo::Feature::CORNER;
      ftr.grad = {grad[2 * i], grad[2 * i + 1]};
      double depth = 0.0;
      const int ok = matcher.findEpipolarMatchDirect(*ref, *cur, ftr, d[3 * i], d[3 * i + 1], d[3 * i + 2], depth) ? 1 : 0;
      const double rec[7] = {(double)ok, depth, matcher.px_cur_[0], matcher.px_cur_[1], (double)matcher.search_level_,
                             matcher.epi_length_, matcher.reject_ ? 1.0 : 0.0};
      wr(fo, rec, 7);
    }
  }
  // ---- feature_alignment::align2D / align1D on cur_frame.img_pyr_[level]  (matcher.cpp:160-171)
  {
    std::vector<int> level(A);
    std::vector<uint8_t> pwb(100 * (size_t)A), patch(64 * (size_t)A);
    std::vector<double> px0(2 * A);
    std::vector<float> dir(2 * A);
    rd(fi, level.data(), level.size()); rd(fi, pwb.data(), pwb.size()); rd(fi, patch.data(), patch.size());
    rd(fi, px0.data(), px0.size()); rd(fi, dir.data(), dir.size());
    for (int i = 0; i < A; ++i) {
      svo::Vector2d p2{px0[2 * i], px0[2 * i + 1]}, p1 = p2;
      double h_inv = 0.0;
      const int ok2 = svo::feature_alignment::align2D(cur->img_pyr_[level[i]], &pwb[100 * (size_t)i], &patch[64 * (size_t)i], 10, p2) ? 1 : 0;
      const int ok1 = svo::feature_alignment::align1D(cur->img_pyr_[level[i]], {dir[2 * i], dir[2 * i + 1]}, &pwb[100 * (size_t)i],
                                                      &patch[64 * (size_t)i], 10, p1, h_inv) ? 1 : 0;
      const double rec[7] = {(double)ok2, p2[0], p2[1], (double)ok1, p1[0], p1[1], h_inv};
      wr(fo, rec, 7);
    }
  }
  // ---- DepthFilter over F further frames: seeds -> candidates through the newCandidatePoint callback
  {
    std::vector<double> spx(2 * S);
    std::vector<int> slevel(S);
    double depth[2];
    rd(fi, spx.data(), spx.size()); rd(fi, slevel.data(), slevel.size()); rd(fi, depth, 2);
    svo::MapPointCandidates candidates;
    svo::DepthFilter filter(std::bind(&svo::MapPointCandidates::newCandidatePoint, &candidates, std::placeholders::_1, std::placeholders::_2));
    svo::FramePtr kf(new svo::Frame(ctx, cam.get(), ref_img.data(), n_levels, 2.0));
    kf->T_f_w_ = ref->T_f_w_;
    kf->setKeyframe();
    std::vector<svo::Feature*> seeds_ftrs;
    for (int i = 0; i < S; ++i) seeds_ftrs.push_back(new svo::Feature(kf.get(), {spx[2 * i], spx[2 * i + 1]}, slevel[i]));
    filter.addKeyframe(kf, seeds_ftrs, depth[0], depth[1]);
    std::list<svo::Seed> copy;
    filter.getSeedsCopy(kf, copy);
    if ((int)copy.size() != S) { fprintf(stderr, "getSeedsCopy: %zu seeds, expected %d\n", copy.size(), S); return 1; }
    std::vector<svo::FramePtr> frames;
    std::vector<uint8_t> img((size_t)w * h);
    for (int k = 0; k < F; ++k) {
      rd(fi, img.data(), img.size());
      svo::FramePtr fr(new svo::Frame(ctx, cam.get(), img.data(), n_levels, 3.0 + k));
      rd(fi, fr->T_f_w_.m, 12);
      frames.push_back(fr);
      if (k == 0) {  // a halted filter must leave the seeds untouched (depth_filter.cpp:211-212)
        filter.seeds_updating_halt_ = true;
        filter.addFrame(fr);
        filter.seeds_updating_halt_ = false;
        std::list<svo::Seed> again;
        filter.getSeedsCopy(kf, again);
        auto a = copy.begin();
        for (auto& b : again) { if (a->mu != b.mu || a->sigma2 != b.sigma2 || a->a != b.a || a->b != b.b) { fprintf(stderr, "halted updateSeeds changed a seed\n"); return 1; } ++a; }
      }
      filter.addFrame(fr);  // synchronous: updateSeeds(frame)
      const long long cnt[3] = {(long long)filter.getSeeds().size(), (long long)candidates.candidates_.size(), (long long)filter.n_updates_};
      wr(fo, cnt, 3);
    }
    const long long ns = (long long)filter.getSeeds().size(), nc = (long long)candidates.candidates_.size();
    wr(fo, &ns, 1);
    for (const svo::Seed& sd : filter.getSeeds()) {
      const double rec[7] = {sd.ftr->px[0], sd.ftr->px[1], sd.a, sd.b, sd.mu, sd.z_range, sd.sigma2};
      wr(fo, rec, 7);
    }
    wr(fo, &nc, 1);
    for (auto& c : candidates.candidates_) {
      // the candidate's first observation must be the seed feature (Point(xyz_world, it->ftr), depth_filter.cpp:265)
      if (c.first->obs_.empty() || c.first->obs_.front() != c.second || c.second->point != c.first) { fprintf(stderr, "candidate without its seed observation\n"); return 1; }
      const double rec[5] = {c.second->px[0], c.second->px[1], c.first->pos_[0], c.first->pos_[1], c.first->pos_[2]};
      wr(fo, rec, 5);
    }
    // candidates own their point and feature (MapPointCandidates::reset); unconverged seed features are ours
    for (const svo::Seed& sd : filter.getSeeds()) delete sd.ftr;
  }
  for (svo::Point* p : points) delete p;
  return 0;
}

static int run_pipeline(FILE* fi, FILE* fo) {
  std::unique_ptr<svo::AbstractCamera> cam(read_camera(fi));
  const int w = cam->width(), h = cam->height();
  int hdr[3];
  rd(fi, hdr, 3);
  const int K = hdr[0], N = hdr[1], use_thread = hdr[2];
  svo::FrameHandlerMono::Options opt;
  opt.mapper_thread = use_thread != 0;
  svo::FrameHandlerMono vo(cam.get(), opt, 0);
  std::vector<uint8_t> img((size_t)w * h);
  rd(fi, img.data(), img.size());
  const int n_levels = std::max(opt.n_pyr_levels, opt.klt_max_level + 1);
  svo::FramePtr first(new svo::Frame(vo.trackingContext(), cam.get(), img.data(), n_levels, 0.0));
  rd(fi, first->T_f_w_.m, 12);
  std::vector<double> px(2 * N), pos(3 * N);
  rd(fi, px.data(), px.size());
  rd(fi, pos.data(), pos.size());
  std::vector<svo::Point*> points;
  for (int i = 0; i < N; ++i) {
    svo::Feature* ftr = new svo::Feature(first.get(), {px[2 * i], px[2 * i + 1]}, 0);
    svo::Point* pt = new svo::Point({pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]});
    ftr->point = pt;
    points.push_back(pt);
    first->addFeature(ftr);
  }
  vo.setFirstFrame(first);
  for (int k = 1; k < K; ++k) {
    rd(fi, img.data(), img.size());
    const int res = (int)vo.addImage(img.data(), (double)k);
    const svo::FrameHandlerMono::FrameLog& lg = vo.log();
    const double rec[6] = {(double)res, (double)lg.img_align_n_tracked, (double)lg.repr_n_matches, (double)lg.sfba_n_edges_final,
                           lg.sfba_error_init, lg.sfba_error_final};
    wr(fo, rec, 6);
    wr(fo, vo.lastFrame()->T_f_w_.m, 12);
  }
  // let the mapper thread drain its queue, then report what it built
  for (int spin = 0; spin < 2000 && !vo.depthFilter()->idle(); ++spin) std::this_thread::sleep_for(std::chrono::milliseconds(1));
  vo.depthFilter()->stopThread();
  const long long out[4] = {(long long)vo.depthFilter()->getSeeds().size(), (long long)vo.map().point_candidates_.candidates_.size(),
                            (long long)vo.depthFilter()->n_updates_, (long long)vo.map().keyframes_.size()};
  wr(fo, out, 4);
  printf("host_pipeline_demo: %d frames, %lld seeds left, %lld candidates, %lld seed updates, %lld keyframes\n", K - 1, out[0], out[1], out[2], out[3]);
  return 0;
}

int main(int argc, char** argv) {
  if (argc != 4) { fprintf(stderr, "usage: %s units|pipeline in.bin out.bin\n", argv[0]); return 2; }
  FILE* fi = fopen(argv[2], "rb");
  if (!fi) { perror("open input"); return 2; }
  FILE* fo = fopen(argv[3], "wb");
  if (!fo) { perror("open output"); return 2; }
  int rc = 1;
  try {
    rc = !strcmp(argv[1], "units") ? run_units(fi, fo) : run_pipeline(fi, fo);
  } catch (const std::exception& e) {
    fprintf(stderr, "host_pipeline_demo: %s\n", e.what());
    rc = 1;
  }
  fclose(fi);
  fclose(fo);
  return rc;
}
