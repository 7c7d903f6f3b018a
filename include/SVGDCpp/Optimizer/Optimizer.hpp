/* Optimizer.hpp — Optimizer base (reference Optimizer/Optimizer.hpp:19-48).  The update itself is the
 * epilogue of the fused pair-interaction kernel; these classes carry the hyper-parameters. */
#ifndef SVGDCPP_B200_OPTIMIZER_HPP
#define SVGDCPP_B200_OPTIMIZER_HPP

#include "../Core.hpp"

class Optimizer {
public:
    Optimizer(const double &lr, const double &epsilon = 1.0e-8) : learning_rate_(lr), stabilizer_(epsilon) {}
    virtual ~Optimizer() {}
    virtual void Initialize() {}
    /* Pushes kind + hyper-parameters into a device context (called by SVGD). */
    virtual void Upload(svgdb_ctx *ctx) const = 0;

protected:
    double learning_rate_;
    double stabilizer_;
};
#endif
