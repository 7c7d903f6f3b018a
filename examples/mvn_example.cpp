// The reference's examples/multivariate_normal/mvn_example.cpp scenario on the B200 facade:
// 2-D MVN target, 10 particles, median-heuristic RBF kernel, AdaGrad(0.1), 1000 iterations.
// Prints the same two blocks as the reference program; tests compare them with the published output
// (reference examples/README.md:7-12).
#include <iostream>

#include "Core"
#include "Kernel"
#include "Model"
#include "Optimizer"

int main()
{
    Eigen::Vector2d mean(-0.6871, 0.8010);
    Eigen::Matrix2d covariance;
    covariance << 0.2260, 0.1652, 0.1652, 0.6779;
    covariance *= 5;
    std::shared_ptr<Model> mvn_ptr = std::make_shared<MultivariateNormal>(mean, covariance);

    size_t dim = 2, num_particles = 10, num_iterations = 1000;
    auto x0 = std::make_shared<Eigen::MatrixXd>(3 * Eigen::MatrixXd::Random(dim, num_particles));
    std::cout << "Initial particle coordinates" << std::endl << *x0 << std::endl;

    std::shared_ptr<Kernel> rbf_ptr = std::make_shared<GaussianRBFKernel>(x0, GaussianRBFKernel::ScaleMethod::Median, mvn_ptr);
    std::shared_ptr<Optimizer> opt_ptr = std::make_shared<AdaGrad>(dim, num_particles, 1.0e-1);

    SVGD svgd(dim, num_iterations, x0, rbf_ptr, mvn_ptr, opt_ptr);
    svgd.Initialize();
    svgd.Run();

    std::cout << "Final particle coordinates" << std::endl << *x0 << std::endl;
    return 0;
}
