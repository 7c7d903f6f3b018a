/* AdaGrad.hpp — reference Optimizer/AdaGrad.hpp:31-65: s += g^2, step = lr g / (eps + sqrt(s)). */
#ifndef SVGDCPP_B200_ADAGRAD_HPP
#define SVGDCPP_B200_ADAGRAD_HPP

#include "Optimizer.hpp"

class AdaGrad : public Optimizer {
public:
    AdaGrad(const size_t &dimension, const size_t &num_particles, const double &lr, const double &epsilon = 1.0e-8)
        : Optimizer(lr, epsilon), dimension_(dimension), num_particles_(num_particles) {}
    void Upload(svgdb_ctx *ctx) const override
    {
        svgdcpp_b200::ThrowOnError(svgdb_set_optimizer(ctx, SVGDB_OPT_ADAGRAD, learning_rate_, 0.0, 0.0, stabilizer_), svgdb_last_error(ctx));
    }

protected:
    size_t dimension_, num_particles_;
};
#endif
