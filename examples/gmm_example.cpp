// Twenty particles and a two-mode target built as the sum of two Gaussians (Model::operator+), median-heuristic RBF kernel,
// Adam for 1000 iterations -- the scenario of the reference's Gaussian-mixture notebook (cell 4 of gmm_example.ipynb), on the
// B200 facade.  tests/test_facade_gpu.py compares the printed particles with the notebook's captured output.
#include <iostream>

#include "Core"
#include "Kernel"
#include "Model"
#include "Optimizer"

namespace {

constexpr size_t kDim = 2, kParticles = 20, kIterations = 1000;

struct Component {
    double mean[2];
    double cov[4]; // row-major, before the common factor 5
};

const Component kModes[2] = {
    {{3.6871, -2.801}, {0.5001, 0.2426, 0.2426, 0.8420}},
    {{-2.9802, 4.3387}, {0.6779, -0.1652, -0.1652, 0.2260}},
};

MultivariateNormal MakeMode(const Component &c)
{
    Eigen::Matrix2d cov;
    cov << c.cov[0], c.cov[1], c.cov[2], c.cov[3];
    cov *= 5;
    return MultivariateNormal(Eigen::Vector2d(c.mean[0], c.mean[1]), cov);
}

void Report(const char *title, const Eigen::MatrixXd &particles)
{
    std::cout << title << std::endl << particles << std::endl;
}

} // namespace

int main()
{
    // unweighted, unnormalised sum of the two modes: what the reference calls its Gaussian mixture
    std::shared_ptr<Model> target = std::make_shared<Model>(MakeMode(kModes[0]) + MakeMode(kModes[1]));

    auto particles = std::make_shared<Eigen::MatrixXd>(8 * Eigen::MatrixXd::Random(kDim, kParticles));
    Report("Initial particle coordinates", *particles);

    SVGDOptions options;
    options.Dimension = kDim;
    options.NumIterations = kIterations;
    options.CoordinateMatrixPtr = particles;
    options.ModelPtr = target;
    options.KernelPtr = std::make_shared<GaussianRBFKernel>(particles, GaussianRBFKernel::ScaleMethod::Median, target);
    options.OptimizerPtr = std::make_shared<Adam>(kDim, kParticles, 0.1, 0.9, 0.999);

    SVGD driver(options);
    driver.Initialize();
    driver.Run();

    Report("Final particle coordinates", *particles);
    return 0;
}
