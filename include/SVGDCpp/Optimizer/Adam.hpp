/* Adam.hpp — reference Optimizer/Adam.hpp:33-96: m = b1 m + (1-b1) g, v = b2 v + (1-b2) g^2,
 * step = lr (m / (1-b1^t)) / (eps + sqrt(v / (1-b2^t))), eps outside the square root. */
#ifndef SVGDCPP_B200_ADAM_HPP
#define SVGDCPP_B200_ADAM_HPP

#include "Optimizer.hpp"

class Adam : public Optimizer {
public:
    Adam(const size_t &dimension, const size_t &num_particles, const double &lr, const double &beta1, const double &beta2,
         const double &epsilon = 1.0e-8)
        : Optimizer(lr, epsilon), dimension_(dimension), num_particles_(num_particles), decay_rate_1_(beta1), decay_rate_2_(beta2)
    {
        if (beta1 >= 1.0 || beta1 < 0.0 || beta2 >= 1.0 || beta2 < 0.0)
            throw std::invalid_argument(SVGDCPP_LOG_PREFIX + "[Argument Error] Invalid value for decay parameter beta.");
    }
    void Upload(svgdb_ctx *ctx) const override
    {
        svgdcpp_b200::ThrowOnError(svgdb_set_optimizer(ctx, SVGDB_OPT_ADAM, learning_rate_, decay_rate_1_, decay_rate_2_, stabilizer_), svgdb_last_error(ctx));
    }

protected:
    size_t dimension_, num_particles_;
    double decay_rate_1_, decay_rate_2_;
};
#endif
