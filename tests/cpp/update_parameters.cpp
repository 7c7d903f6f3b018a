// SVGD::UpdateModelParameters / SVGD::UpdateKernelParameters between two Run() calls through the facade (reference SVGD.hpp:304-332,
// MultivariateNormal::UpdateParameters Model/MultivariateNormal.hpp:94-115): 5 iterations on one Gaussian target, the target is
// re-parametrised, 5 more iterations with the optimizer state kept; then the same with a constant-scale kernel whose scale is
// changed in between.  Final coordinates of both runs go to stdout with 17 digits; tests/test_gpu_updates.py checks them
// against the oracle.
#include <iomanip>
#include <iostream>

#include "Core"
#include "Kernel"
#include "Model"
#include "Optimizer"

static void print_matrix(const Eigen::MatrixXd &m)
{
    std::cout << std::setprecision(17);
    for (Eigen::Index r = 0; r < m.rows(); ++r) {
        for (Eigen::Index c = 0; c < m.cols(); ++c) std::cout << m(r, c) << (c + 1 < m.cols() ? " " : "\n");
    }
}

int main()
{
    const size_t dim = 2, num_particles = 8, num_iterations = 5;
    Eigen::Vector2d mean1(0.5, -0.25), mean2(-1.0, 0.75);
    Eigen::Matrix2d cov1, cov2;
    cov1 << 0.5, 0.2, 0.2, 0.8;
    cov2 << 1.5, -0.3, -0.3, 0.6;
    Eigen::MatrixXd start(dim, num_particles);
    start << 1.5, -0.75, 0.25, 2.0, -1.25, 0.5, -2.0, 1.0,
        -0.5, 1.0, 0.75, -1.5, 0.125, 2.25, 0.5, -1.0;
    for (int scenario = 0; scenario < 2; ++scenario) {
        auto x0 = std::make_shared<Eigen::MatrixXd>(start);
        std::shared_ptr<Model> model_ptr = std::make_shared<MultivariateNormal>(mean1, cov1);
        const auto method = scenario == 0 ? GaussianRBFKernel::ScaleMethod::Median : GaussianRBFKernel::ScaleMethod::Constant;
        std::shared_ptr<Kernel> kernel_ptr = std::make_shared<GaussianRBFKernel>(x0, method, model_ptr);
        if (scenario == 1) kernel_ptr->UpdateParameters({0.8 * Eigen::MatrixXd::Identity(dim, dim)});
        std::shared_ptr<Optimizer> opt_ptr = std::make_shared<Adam>(dim, num_particles, 1.0e-1, 0.9, 0.999);
        SVGD svgd(dim, num_iterations, x0, kernel_ptr, model_ptr, opt_ptr);
        svgd.Initialize();
        svgd.Run();
        Eigen::MatrixXd m2 = mean2, c2 = cov2;
        svgd.UpdateModelParameters({m2, c2});
        if (scenario == 1) svgd.UpdateKernelParameters({0.3 * Eigen::MatrixXd::Identity(dim, dim)});
        svgd.Run();
        print_matrix(*x0);
    }
    return 0;
}
