/* GaussianRBFKernel.hpp — k(x, x') = exp(-(x-x')^T A (x-x')), A = a I
 * (reference Kernel/GaussianRBFKernel.hpp:47-88).  ScaleMethod::Median recomputes
 * a = log(n) / median(|x_i - x_j|)^2 from the current particles at every Step (:141-188); on the
 * device that is an exact radix select over all n^2 distances (svgdcpp_b200/csrc/select.cuh).
 * ScaleMethod::Hessian (:189-210) makes A the mean negative Hessian of log p over the particles / (2 d); the device
 * path serves it for the built-in Gaussian models (svgdcpp_b200/csrc/kernels_hessian.cuh). */
#ifndef SVGDCPP_B200_GAUSSIAN_RBF_KERNEL_HPP
#define SVGDCPP_B200_GAUSSIAN_RBF_KERNEL_HPP

#include "../Model/Model.hpp"
#include "Kernel.hpp"

class GaussianRBFKernel : public Kernel {
public:
    enum class ScaleMethod {
        Median = 0,
        Hessian = 1,
        Constant = 2 /* the reference's "TODO: constant scale"; value a given by UpdateParameters({a I}) */
    };

    GaussianRBFKernel() {}
    GaussianRBFKernel(const std::shared_ptr<Eigen::MatrixXd> &coord_mat_ptr, const ScaleMethod &method = ScaleMethod::Median,
                      const std::shared_ptr<Model> &model_ptr = nullptr)
        : Kernel(static_cast<size_t>(coord_mat_ptr->rows())), scale_method_(method), coord_matrix_ptr_(coord_mat_ptr), target_model_ptr_(model_ptr)
    {
        if (scale_method_ == ScaleMethod::Hessian && !model_ptr) throw UnsetException("Hessian-based scale requires a model.");
    }

    /* With ScaleMethod::Median / Hessian the scale is recomputed from the particles at every Step and overwrites the parameters
     * (reference GaussianRBFKernel.hpp:141-156), so UpdateParameters only takes effect for ScaleMethod::Constant, where
     * params[0] must be a I (the device kernel's scale is a scalar). */
    void UpdateParameters(const std::vector<Eigen::MatrixXd> &params) override
    {
        Kernel::UpdateParameters(params);
        if (scale_method_ != ScaleMethod::Constant || params.empty() || params[0].size() == 0) return;
        const Eigen::MatrixXd &A = params[0];
        if (A.rows() != dimension_ || A.cols() != dimension_) throw DimensionMismatchException("Kernel parameter matrix must be dimension x dimension.");
        for (int r = 0; r < dimension_; ++r)
            for (int c = 0; c < dimension_; ++c)
                if (A(r, c) != (r == c ? A(0, 0) : 0.0))
                    throw std::invalid_argument(SVGDCPP_LOG_PREFIX + "[Argument Error] the device RBF kernel takes a scalar scale: A must be a * I.");
        fixed_scale_ = A(0, 0);
    }

    std::unique_ptr<Kernel> CloneUniquePointer() const override { return std::make_unique<GaussianRBFKernel>(*this); }
    std::shared_ptr<Kernel> CloneSharedPointer() const override { return std::make_shared<GaussianRBFKernel>(*this); }

    void Upload(svgdb_ctx *ctx) const override
    {
        svgdcpp_b200::ThrowOnError(svgdb_set_kernel_rbf(ctx, static_cast<int>(scale_method_), fixed_scale_), svgdb_last_error(ctx));
    }

protected:
    ScaleMethod scale_method_ = ScaleMethod::Median;
    double fixed_scale_ = 0.0;
    std::shared_ptr<Eigen::MatrixXd> coord_matrix_ptr_;
    std::shared_ptr<Model> target_model_ptr_;
};
#endif
