// rpg_svo_b200/host/host_demo.cpp -- drives the C++ host classes (svo_host.h) the way the reference's
// own programs drive SVO: build two Frames with Features/Points, then
//   svo::SparseImgAlign(max, min, 30, GaussNewton, false, false).run(ref, cur)   (svo/test/test_sparse_img_align.cpp:121-123)
//   svo::pose_optimizer::optimizeGaussNewton(2.0, 10, false, frame, ...)          (svo/test/test_pose_optimizer.cpp:99-104)
//   svo::feature_detection::FastDetector(w, h, 30, 3).detect(frame, 20.0, fts)    (svo/test/test_feature_detection.cpp, depth_filter.cpp:114-119)
// Inputs come from a binary dump written by tests/test_host_cpp_gpu.py, results go to a second file that
// the test compares with the CPU oracle.   usage: host_demo in.bin out.bin
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "svo_host.h"

template <class T>
static void rd(FILE* f, T* p, size_t n) {
  if (fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "host_demo: short read\n"); exit(2); }
}

int main(int argc, char** argv) {
  if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
  FILE* fi = fopen(argv[1], "rb");
  if (!fi) { perror("open input"); return 2; }
  int hdr[6];
  rd(fi, hdr, 6);
  const int w = hdr[0], h = hdr[1], n_levels = hdr[2], N = hdr[3], max_level = hdr[4], min_level = hdr[5];
  double camv[4];
  rd(fi, camv, 4);
  std::vector<uint8_t> ref_img((size_t)w * h), cur_img((size_t)w * h), has_point(N);
  rd(fi, ref_img.data(), ref_img.size());
  rd(fi, cur_img.data(), cur_img.size());
  svo::SE3 T_ref_w, T_cur_w_init;
  rd(fi, T_ref_w.m, 12);
  rd(fi, T_cur_w_init.m, 12);
  std::vector<double> px(2 * N), f(3 * N), pos(3 * N), f_cur(3 * N);
  std::vector<int> level(N);
  rd(fi, px.data(), px.size());
  rd(fi, f.data(), f.size());
  rd(fi, pos.data(), pos.size());
  rd(fi, has_point.data(), has_point.size());
  rd(fi, f_cur.data(), f_cur.size());
  rd(fi, level.data(), level.size());
  fclose(fi);

  try {
    svo::Context ctx(0);
    svo::PinholeCamera cam(w, h, camv[0], camv[1], camv[2], camv[3]);
    svo::FramePtr frame_ref(new svo::Frame(ctx, &cam, ref_img.data(), n_levels, 0.0));
    svo::FramePtr frame_cur(new svo::Frame(ctx, &cam, cur_img.data(), n_levels, 1.0));
    frame_ref->T_f_w_ = T_ref_w;
    frame_cur->T_f_w_ = T_cur_w_init;  // processFrame: new_frame.T_f_w_ = last_frame.T_f_w_ (frame_handler_mono.cpp:132)
    std::vector<svo::Point*> points;
    for (int i = 0; i < N; ++i) {
      svo::Point* pt = has_point[i] ? new svo::Point({pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]}) : nullptr;
      if (pt) points.push_back(pt);
      frame_ref->addFeature(new svo::Feature(frame_ref.get(), pt, {px[2 * i], px[2 * i + 1]}, {f[3 * i], f[3 * i + 1], f[3 * i + 2]}, 0));
      frame_cur->addFeature(new svo::Feature(frame_cur.get(), pt, {0.0, 0.0}, {f_cur[3 * i], f_cur[3 * i + 1], f_cur[3 * i + 2]}, level[i]));
    }
    svo::SparseImgAlign img_align(max_level, min_level, 30, svo::SparseImgAlign::GaussNewton, false, false);
    const size_t n_tracked = img_align.run(frame_ref, frame_cur);
    const svo::Matrix6d fisher = img_align.getFisherInformation();
    const svo::SE3 T_after_align = frame_cur->T_f_w_;

    double estimated_scale = 0, error_init = 0, error_final = 0;
    size_t num_obs = 0;
    svo::pose_optimizer::optimizeGaussNewton(2.0, 10, false, frame_cur, estimated_scale, error_init, error_final, num_obs);
    std::vector<uint8_t> hp_after;
    for (svo::Feature* ft : frame_cur->fts_) hp_after.push_back(ft->point != nullptr);

    // seed initialisation on a keyframe (depth_filter.cpp:114-119): new corners away from the existing features
    svo::feature_detection::FastDetector detector(w, h, 30, 3);
    svo::Features new_features;
    detector.setExistingFeatures(frame_ref->fts_);
    detector.detect(frame_ref.get(), 20.0, new_features);

    FILE* fo = fopen(argv[2], "wb");
    if (!fo) { perror("open output"); return 2; }
    long long nt = (long long)n_tracked, no = (long long)num_obs;
    fwrite(T_after_align.m, sizeof(double), 12, fo);
    fwrite(&nt, sizeof(nt), 1, fo);
    fwrite(fisher.data(), sizeof(double), 36, fo);
    fwrite(frame_cur->T_f_w_.m, sizeof(double), 12, fo);
    const double sc[3] = {estimated_scale, error_init, error_final};
    fwrite(sc, sizeof(double), 3, fo);
    fwrite(&no, sizeof(no), 1, fo);
    fwrite(hp_after.data(), 1, hp_after.size(), fo);
    const int n_new = (int)new_features.size();
    fwrite(&n_new, sizeof(int), 1, fo);
    for (svo::Feature* nf : new_features) {
      const int rec[3] = {(int)nf->px[0], (int)nf->px[1], nf->level};
      fwrite(rec, sizeof(int), 3, fo);
      delete nf;
    }
    fclose(fo);
    printf("host_demo: tracked %zu patches, pose-opt kept %zu observations (err %.3f -> %.3f px)\n", n_tracked, num_obs,
           error_init, error_final);
    for (svo::Point* p : points) delete p;
  } catch (const std::exception& e) {
    fprintf(stderr, "host_demo: %s\n", e.what());
    return 1;
  }
  return 0;
}
