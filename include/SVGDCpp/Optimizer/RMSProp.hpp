/* RMSProp.hpp — reference Optimizer/RMSProp.hpp:33-74: s = b s + (1-b) g^2, step = lr g / (eps + sqrt(s)). */
#ifndef SVGDCPP_B200_RMSPROP_HPP
#define SVGDCPP_B200_RMSPROP_HPP

#include "Optimizer.hpp"

class RMSProp : public Optimizer {
public:
    RMSProp(const size_t &dimension, const size_t &num_particles, const double &lr, const double &beta, const double &epsilon = 1.0e-8)
        : Optimizer(lr, epsilon), dimension_(dimension), num_particles_(num_particles), decay_rate_(beta)
    {
        if (beta > 1.0 || beta < 0.0)
            throw std::invalid_argument(SVGDCPP_LOG_PREFIX + "[Argument Error] Invalid value for decay parameter beta.");
    }
    void Upload(svgdb_ctx *ctx) const override
    {
        svgdcpp_b200::ThrowOnError(svgdb_set_optimizer(ctx, SVGDB_OPT_RMSPROP, learning_rate_, decay_rate_, 0.0, stabilizer_), svgdb_last_error(ctx));
    }

protected:
    size_t dimension_, num_particles_;
    double decay_rate_;
};
#endif
