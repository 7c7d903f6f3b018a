/* Kernel.hpp — Kernel base of the facade (reference Kernel/Kernel.hpp).  The driver's per-pair
 * virtual calls (EvaluateKernel / EvaluateKernelGrad / UpdateLocation, Kernel.hpp:279-330) do not
 * exist on the device path: the whole N x N interaction is one fused kernel, so a Kernel here only
 * describes WHICH interaction to run.  Generic taped lambdas and + - * / composition
 * (Kernel.hpp:55-223) are out of scope. */
#ifndef SVGDCPP_B200_KERNEL_HPP
#define SVGDCPP_B200_KERNEL_HPP

#include "../Core.hpp"

class Kernel {
public:
    Kernel() {}
    explicit Kernel(const size_t &dim) : dimension_(static_cast<int>(dim)) {}
    virtual ~Kernel() {}

    virtual void Initialize() {}
    virtual void Step() {}
    virtual void UpdateParameters(const std::vector<Eigen::MatrixXd> &params) { kernel_parameters_ = params; }
    std::vector<Eigen::MatrixXd> GetParameters() const { return kernel_parameters_; }
    virtual std::unique_ptr<Kernel> CloneUniquePointer() const { return std::make_unique<Kernel>(*this); }
    virtual std::shared_ptr<Kernel> CloneSharedPointer() const { return std::make_shared<Kernel>(*this); }

    int Dimension() const { return dimension_; }

    /* Pushes this kernel into a device context (called by SVGD). */
    virtual void Upload(svgdb_ctx *) const { throw UnsetException("Kernel function is unset."); }

protected:
    int dimension_ = -1;
    std::vector<Eigen::MatrixXd> kernel_parameters_;
};
#endif
