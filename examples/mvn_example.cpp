// Ten particles, a correlated 2-D Gaussian target, the median-heuristic RBF kernel and AdaGrad for 1000 iterations -- the
// scenario whose output the reference publishes (reference examples/README.md:7-12), here driven through SVGDOptions on the
// B200 facade.  tests/test_facade_gpu.py compares the two printed blocks with that published output character by character.
#include <iostream>

#include "Core"
#include "Kernel"
#include "Model"
#include "Optimizer"

namespace {

constexpr size_t kDim = 2, kParticles = 10, kIterations = 1000;

void Report(const char *title, const Eigen::MatrixXd &particles)
{
    std::cout << title << std::endl << particles << std::endl;
}

} // namespace

int main()
{
    // target density: N((-0.6871, 0.8010), 5 S)
    Eigen::Matrix2d sigma;
    sigma << 0.2260, 0.1652, 0.1652, 0.6779;
    sigma *= 5;
    std::shared_ptr<Model> target = std::make_shared<MultivariateNormal>(Eigen::Vector2d(-0.6871, 0.8010), sigma);

    // particles: 3 * uniform(-1, 1), drawn like Eigen::MatrixXd::Random (unseeded)
    auto particles = std::make_shared<Eigen::MatrixXd>(3 * Eigen::MatrixXd::Random(kDim, kParticles));
    Report("Initial particle coordinates", *particles);

    SVGDOptions options;
    options.Dimension = kDim;
    options.NumIterations = kIterations;
    options.CoordinateMatrixPtr = particles;
    options.ModelPtr = target;
    options.KernelPtr = std::make_shared<GaussianRBFKernel>(particles, GaussianRBFKernel::ScaleMethod::Median, target);
    options.OptimizerPtr = std::make_shared<AdaGrad>(kDim, kParticles, 0.1);

    SVGD driver(options);
    driver.Initialize();
    driver.Run(); // the particle matrix is updated in place

    Report("Final particle coordinates", *particles);
    return 0;
}
