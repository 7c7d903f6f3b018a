/* MultivariateNormal.hpp — unnormalised Gaussian density exp(-1/2 (x-mu)^T Sigma^-1 (x-mu))
 * (reference Model/MultivariateNormal.hpp:39-64); parameters {mean, covariance} (:49-50). */
#ifndef SVGDCPP_B200_MULTIVARIATE_NORMAL_HPP
#define SVGDCPP_B200_MULTIVARIATE_NORMAL_HPP

#include "Model.hpp"

class MultivariateNormal : public Model {
public:
    MultivariateNormal() {}
    MultivariateNormal(const Eigen::VectorXd &mean, const Eigen::MatrixXd &covariance) : Model(static_cast<size_t>(mean.rows()))
    {
        if (covariance.rows() != mean.rows() || covariance.cols() != mean.rows())
            throw DimensionMismatchException("Dimensions of parameter vectors/matrices do not match.");
        model_parameters_ = {mean, covariance};
        ComputeNormalizationConstant();
    }

    void UpdateParameters(const std::vector<Eigen::MatrixXd> &params) override
    {
        const Eigen::MatrixXd &mean = params.at(0), &covariance = params.at(1);
        if (covariance.rows() != mean.rows() || covariance.cols() != mean.rows())
            throw DimensionMismatchException("Dimensions of parameter vectors/matrices do not match each other (# of rows must be equal).");
        if (mean.rows() != dimension_)
            throw DimensionMismatchException("Dimensions of parameter vectors/matrices do not match original dimension.");
        model_parameters_ = {mean, covariance};
        ComputeNormalizationConstant();
    }

    std::unique_ptr<Model> CloneUniquePointer() const override { return std::make_unique<MultivariateNormal>(*this); }
    std::shared_ptr<Model> CloneSharedPointer() const override { return std::make_shared<MultivariateNormal>(*this); }

    /* Normalised variants (reference :143-175). */
    double EvaluateModelNormalized(const Eigen::VectorXd &x) const { return norm_const_ * EvaluateModel(x); }
    double EvaluateLogModelNormalized(const Eigen::VectorXd &x) const { return std::log(norm_const_) + EvaluateLogModel(x); }
    Eigen::VectorXd EvaluateModelGradNormalized(const Eigen::VectorXd &x) const
    {
        Eigen::VectorXd g = EvaluateModelGrad(x);
        g *= norm_const_;
        return g;
    }
    double GetNormalizationConstant() const { return norm_const_; }

protected:
    /* 1 / ((2 pi)^(d/2) sqrt(det Sigma)), reference :182-186; determinant by Gaussian elimination. */
    void ComputeNormalizationConstant()
    {
        const int d = dimension_;
        std::vector<double> a(model_parameters_[1].data(), model_parameters_[1].data() + static_cast<size_t>(d) * d);
        double det = 1.0;
        for (int c = 0; c < d; ++c) {
            int p = c;
            for (int r = c + 1; r < d; ++r)
                if (std::fabs(a[static_cast<size_t>(r) * d + c]) > std::fabs(a[static_cast<size_t>(p) * d + c])) p = r;
            if (a[static_cast<size_t>(p) * d + c] == 0.0) { det = 0.0; break; }
            if (p != c) {
                for (int k = 0; k < d; ++k) std::swap(a[static_cast<size_t>(p) * d + k], a[static_cast<size_t>(c) * d + k]);
                det = -det;
            }
            det *= a[static_cast<size_t>(c) * d + c];
            for (int r = c + 1; r < d; ++r) {
                double f = a[static_cast<size_t>(r) * d + c] / a[static_cast<size_t>(c) * d + c];
                for (int k = c; k < d; ++k) a[static_cast<size_t>(r) * d + k] -= f * a[static_cast<size_t>(c) * d + k];
            }
        }
        norm_const_ = 1.0 / (std::pow(2.0 * M_PI, d / 2.0) * std::sqrt(det));
    }

    double norm_const_ = 0.0;
};
#endif
