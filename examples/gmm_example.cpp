// The reference's examples/gaussian_mixture_model/gmm_example.cpp scenario on the B200 facade:
// sum of two 2-D Gaussians (Model::operator+), 20 particles, median RBF kernel, Adam, 1000 iterations.
#include <iostream>

#include "Core"
#include "Kernel"
#include "Model"
#include "Optimizer"

int main()
{
    Eigen::Vector2d mean1(3.6871, -2.801), mean2(-2.9802, 4.3387);
    Eigen::Matrix2d cov1, cov2;
    cov1 << 0.5001, 0.2426, 0.2426, 0.8420;
    cov2 << 0.6779, -0.1652, -0.1652, 0.2260;
    cov1 *= 5;
    cov2 *= 5;

    MultivariateNormal mvn1(mean1, cov1);
    MultivariateNormal mvn2(mean2, cov2);
    Model gmm = mvn1 + mvn2;
    std::shared_ptr<Model> gmm_ptr = std::make_shared<Model>(gmm);

    size_t dim = 2, num_particles = 20, num_iterations = 1000;
    auto x0 = std::make_shared<Eigen::MatrixXd>(8 * Eigen::MatrixXd::Random(dim, num_particles));
    std::cout << "Initial particle coordinates" << std::endl << *x0 << std::endl;

    std::shared_ptr<Kernel> rbf_ptr = std::make_shared<GaussianRBFKernel>(x0, GaussianRBFKernel::ScaleMethod::Median, gmm_ptr);
    std::shared_ptr<Optimizer> opt_ptr = std::make_shared<Adam>(dim, num_particles, 1.0e-1, 0.9, 0.999);

    SVGD svgd(dim, num_iterations, x0, rbf_ptr, gmm_ptr, opt_ptr);
    svgd.Initialize();
    svgd.Run();

    std::cout << "Final particle coordinates" << std::endl << *x0 << std::endl;
    return 0;
}
