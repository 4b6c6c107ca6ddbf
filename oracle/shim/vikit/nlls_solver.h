// oracle/shim/vikit/nlls_solver.h -- TEST INFRASTRUCTURE ONLY: [EXT] vk::NLLSSolver<D,T> restated from the published
// rpg_vikit sources (nlls_solver.h / nlls_solver_impl.hpp): Gauss-Newton driver only (the reference uses GaussNewton).
#pragma once
#include <Eigen/Core>
#include <vikit/math_utils.h>
#include <vikit/robust_cost.h>
#include <iostream>
namespace vk {
using namespace Eigen;
template <int D, typename T>
class NLLSSolver {
 public:
  typedef T ModelType;
  enum Method { GaussNewton, LevenbergMarquardt };
  enum ScaleEstimatorType { UnitScale, TDistScale, MADScale, NormalScale };
  enum WeightFunctionType { UnitWeight, TDistWeight, TukeyWeight, HuberWeight };
 protected:
  Matrix<double, D, D> H_;
  Matrix<double, D, 1> Jres_;
  Matrix<double, D, 1> x_;
  bool have_prior_;
  ModelType prior_;
  Matrix<double, D, D> I_prior_;
  double chi2_, rho_;
  Method method_;
  virtual double computeResiduals(const ModelType& model, bool linearize_system, bool compute_weight_scale) = 0;
  virtual int solve() = 0;
  virtual void update(const ModelType& old_model, ModelType& new_model) = 0;
  virtual void applyPrior(const ModelType&) {}
  virtual void startIteration() {}
  virtual void finishIteration() {}
  void optimizeGaussNewton(ModelType& model) {
    if (use_weights_) computeResiduals(model, false, true);
    ModelType old_model(model);
    for (iter_ = 0; iter_ < n_iter_; ++iter_) {
      rho_ = 0;
      startIteration();
      H_.setZero();
      Jres_.setZero();
      n_meas_ = 0;
      double new_chi2 = computeResiduals(model, true, false);
      if (have_prior_) applyPrior(model);
      if (!solve()) stop_ = true;
      if ((iter_ > 0 && new_chi2 > chi2_) || stop_) {
        model = old_model;
        break;
      }
      ModelType new_model;
      update(model, new_model);
      old_model = model;
      model = new_model;
      chi2_ = new_chi2;
      finishIteration();
      if (vk::norm_max(x_) <= eps_) break;
    }
  }
 public:
  double mu_init_, mu_, nu_init_, nu_;
  size_t n_iter_init_, n_iter_, n_trials_, n_trials_max_, n_meas_;
  bool stop_, verbose_;
  double eps_;
  size_t iter_;
  bool use_weights_;
  float scale_;
  robust_cost::ScaleEstimatorPtr scale_estimator_;
  robust_cost::WeightFunctionPtr weight_function_;
  NLLSSolver()
      : have_prior_(false), method_(LevenbergMarquardt), mu_init_(0.01f), mu_(mu_init_), nu_init_(2.0), nu_(nu_init_),
        n_iter_init_(15), n_iter_(n_iter_init_), n_trials_(0), n_trials_max_(5), n_meas_(0), stop_(false), verbose_(true),
        eps_(0.0000000001), iter_(0), use_weights_(false), scale_(0.0) {}
  virtual ~NLLSSolver() {}
  void optimize(ModelType& model) { if (method_ == GaussNewton) optimizeGaussNewton(model); }
  void reset() {
    have_prior_ = false; chi2_ = 1e10; mu_ = mu_init_; nu_ = nu_init_; n_meas_ = 0; n_iter_ = n_iter_init_; iter_ = 0; stop_ = false;
  }
  const double& getChi2() const { return chi2_; }
  const Matrix<double, D, D>& getInformationMatrix() const { return H_; }
};
}  // namespace vk
