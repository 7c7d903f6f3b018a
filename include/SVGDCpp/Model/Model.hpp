/* Model.hpp — Model base of the facade (reference Model/Model.hpp).
 *
 * What has a device implementation: MultivariateNormal, sums of them built with operator+
 * (reference Model.hpp:55-92 — how the reference forms its "Gaussian mixture",
 * examples/gaussian_mixture_model/gmm_example.cpp:24), and user models that provide a device
 * gradient through SetDeviceGradient() (replaces the CppAD Jacobian of Model.hpp:335-338).
 * Arbitrary host lambdas (UpdateModel) and - * / composition need the CppAD tape and are out of
 * scope: they raise UnsetException when handed to SVGD; there is no CPU fallback. */
#ifndef SVGDCPP_B200_MODEL_HPP
#define SVGDCPP_B200_MODEL_HPP

#include "../Core.hpp"

class Model {
public:
    Model() {}
    explicit Model(const size_t &dim) : dimension_(static_cast<int>(dim)) {}
    virtual ~Model() {}

    /* Sum of two models (reference Model.hpp:55-92): unweighted, unnormalised. */
    Model operator+(const Model &obj) const
    {
        if (dimension_ != obj.dimension_)
            throw DimensionMismatchException("Only models with the same variable dimensions can be added.");
        if (!IsSet() || !obj.IsSet())
            throw UnsetException("One of the model functions is unset; functional composition requires both model functions to be set.");
        if (hook_ || obj.hook_)
            throw UnsetException("Models with a device-gradient hook cannot be composed on the device path.");
        Model sum(static_cast<size_t>(dimension_));
        sum.model_parameters_ = model_parameters_;
        sum.model_parameters_.insert(sum.model_parameters_.end(), obj.model_parameters_.begin(), obj.model_parameters_.end());
        return sum;
    }

    virtual std::unique_ptr<Model> CloneUniquePointer() const { return std::make_unique<Model>(*this); }
    virtual std::shared_ptr<Model> CloneSharedPointer() const { return std::make_shared<Model>(*this); }

    virtual void Initialize() {}
    virtual void Step() {}

    /* Parameters are {mean_0, cov_0, mean_1, cov_1, ...} like the reference's concatenated
     * model_parameters_ (Model.hpp:68-72). */
    virtual void UpdateParameters(const std::vector<Eigen::MatrixXd> &params)
    {
        if (params.size() % 2 != 0) throw DimensionMismatchException("Expected {mean, covariance} pairs.");
        model_parameters_ = params;
    }
    std::vector<Eigen::MatrixXd> GetParameters() const { return model_parameters_; }

    /* Device hook: see svgdb_grad_fn in svgd_b200.h. */
    void SetDeviceGradient(svgdb_grad_fn fn, void *user = nullptr)
    {
        hook_ = fn;
        hook_user_ = user;
        model_parameters_.clear();
    }

    int Dimension() const { return dimension_; }
    bool IsSet() const { return hook_ != nullptr || !model_parameters_.empty(); }

    /* Point evaluations of the reference's Model (Model.hpp:290-338).  They run on the device like everything else: a
     * one-particle context is created for the call (an inspection aid, not a fast path; the values of all particles of a run come
     * from svgdb_compute_log_model / svgdb_compute_log_model_grad).  log p goes through log-sum-exp, so it stays finite where the
     * reference's log(sum exp) underflows.  Models behind a device-gradient hook only have EvaluateLogModelGrad. */
    double EvaluateLogModel(const Eigen::VectorXd &x) const
    {
        double v = 0.0;
        EvaluateAt(x, &v, nullptr);
        return v;
    }
    double EvaluateModel(const Eigen::VectorXd &x) const { return std::exp(EvaluateLogModel(x)); }
    Eigen::VectorXd EvaluateLogModelGrad(const Eigen::VectorXd &x) const
    {
        Eigen::VectorXd g(dimension_);
        EvaluateAt(x, nullptr, g.data());
        return g;
    }
    Eigen::VectorXd EvaluateModelGrad(const Eigen::VectorXd &x) const
    {
        double v = 0.0;
        Eigen::VectorXd g(dimension_);
        EvaluateAt(x, &v, g.data());
        g *= std::exp(v); /* grad p = p grad log p */
        return g;
    }

    /* Pushes this model into a device context (called by SVGD). */
    void Upload(svgdb_ctx *ctx) const
    {
        if (!IsSet()) throw UnsetException("Model function is unset.");
        if (hook_) {
            svgdcpp_b200::ThrowOnError(svgdb_set_model_device_hook(ctx, hook_, hook_user_), svgdb_last_error(ctx));
            return;
        }
        const size_t n_comp = model_parameters_.size() / 2;
        const size_t d = static_cast<size_t>(dimension_);
        std::vector<double> means(n_comp * d), covs(n_comp * d * d);
        for (size_t c = 0; c < n_comp; ++c) {
            const Eigen::MatrixXd &mu = model_parameters_[2 * c];
            const Eigen::MatrixXd &cov = model_parameters_[2 * c + 1];
            if (static_cast<size_t>(mu.size()) != d || static_cast<size_t>(cov.rows()) != d || static_cast<size_t>(cov.cols()) != d)
                throw DimensionMismatchException("Dimensions of parameter vectors/matrices do not match.");
            for (size_t k = 0; k < d; ++k) means[c * d + k] = mu.data()[k];
            for (size_t k = 0; k < d * d; ++k) covs[c * d * d + k] = cov.data()[k];
        }
        svgdcpp_b200::ThrowOnError(svgdb_set_model_mvn_sum(ctx, static_cast<int32_t>(n_comp), means.data(), covs.data()), svgdb_last_error(ctx));
    }

protected:
    void EvaluateAt(const Eigen::VectorXd &x, double *logp, double *grad) const
    {
        if (x.rows() != dimension_) throw DimensionMismatchException("Argument dimension does not match the model dimension.");
        svgdb_ctx *ctx = nullptr;
        int rc = svgdb_create(&ctx, 0, 1, dimension_, SVGDB_PRECISION_F64);
        std::string msg = ctx ? svgdb_last_error(ctx) : "svgdb_create failed";
        try {
            svgdcpp_b200::ThrowOnError(rc, msg.c_str());
            Upload(ctx);
            svgdcpp_b200::ThrowOnError(svgdb_set_particles(ctx, x.data()), svgdb_last_error(ctx));
            if (logp) svgdcpp_b200::ThrowOnError(svgdb_compute_log_model(ctx, logp), svgdb_last_error(ctx));
            if (grad) svgdcpp_b200::ThrowOnError(svgdb_compute_log_model_grad(ctx, grad), svgdb_last_error(ctx));
        } catch (...) {
            svgdb_destroy(ctx);
            throw;
        }
        svgdb_destroy(ctx);
    }

    int dimension_ = -1;
    std::vector<Eigen::MatrixXd> model_parameters_;
    svgdb_grad_fn hook_ = nullptr;
    void *hook_user_ = nullptr;
};
#endif
