// oracle/shim/vikit/robust_cost.h -- TEST INFRASTRUCTURE ONLY: [EXT] vk::robust_cost (Tukey weights, MAD scale).
#pragma once
#include <vikit/math_utils.h>
#include <memory>
#include <vector>
namespace vk { namespace robust_cost {
class ScaleEstimator { public: virtual ~ScaleEstimator() {} virtual float compute(std::vector<float>& errors) const = 0; };
typedef std::shared_ptr<ScaleEstimator> ScaleEstimatorPtr;
class MADScaleEstimator : public ScaleEstimator {
 public:
  float compute(std::vector<float>& errors) const override { return 1.48f * vk::getMedian(errors); }
};
class WeightFunction { public: virtual ~WeightFunction() {} virtual float value(const float& x) const = 0; };
typedef std::shared_ptr<WeightFunction> WeightFunctionPtr;
class TukeyWeightFunction : public WeightFunction {
  float b_square;
 public:
  TukeyWeightFunction(const float b = 4.6851f) : b_square(b * b) {}
  float value(const float& x) const override {
    const float x_square = x * x;
    if (x_square <= b_square) { const float tmp = 1.0f - x_square / b_square; return tmp * tmp; }
    return 0.0f;
  }
};
} }
