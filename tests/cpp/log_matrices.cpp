// SVGDOptions::LogIntermediateMatrices through the facade (reference SVGD.hpp:346-365): a 2-D Gaussian target, 6 particles,
// RBF kernel with the median scale, Adam, 3 iterations; the log goes to argv[1].  tests/test_facade_gpu.py parses it and
// checks every printed number against the oracle.  Then the model's point evaluations, printed to stdout.
#include <iomanip>
#include <iostream>

#include "Core"
#include "Kernel"
#include "Model"
#include "Optimizer"

int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    const size_t dim = 2, num_particles = 6, num_iterations = 3;
    Eigen::Vector2d mean(0.5, -0.25);
    Eigen::Matrix2d covariance;
    covariance << 0.5, 0.2, 0.2, 0.8;
    auto x0 = std::make_shared<Eigen::MatrixXd>(dim, num_particles);
    *x0 << 1.5, -0.75, 0.25, 2.0, -1.25, 0.5,
        -0.5, 1.0, 0.75, -1.5, 0.125, 2.25;
    std::shared_ptr<Model> model_ptr = std::make_shared<MultivariateNormal>(mean, covariance);
    std::shared_ptr<Kernel> kernel_ptr = std::make_shared<GaussianRBFKernel>(x0, GaussianRBFKernel::ScaleMethod::Median, model_ptr);
    std::shared_ptr<Optimizer> opt_ptr = std::make_shared<Adam>(dim, num_particles, 1.0e-1, 0.9, 0.999);
    SVGDOptions options;
    options.Dimension = dim;
    options.NumIterations = num_iterations;
    options.CoordinateMatrixPtr = x0;
    options.KernelPtr = kernel_ptr;
    options.ModelPtr = model_ptr;
    options.OptimizerPtr = opt_ptr;
    options.LogIntermediateMatrices = true;
    options.IntermediateMatricesOutputPath = argv[1];
    SVGD svgd(options);
    svgd.Initialize();
    svgd.Run();

    // point evaluations of the model (Model.hpp:290-338, MultivariateNormal.hpp:143-175) at a fixed argument
    Eigen::Vector2d x(0.75, -1.5);
    MultivariateNormal mvn(mean, covariance);
    Eigen::VectorXd g = mvn.EvaluateLogModelGrad(x), gp = mvn.EvaluateModelGradNormalized(x);
    std::cout << std::setprecision(17) << mvn.EvaluateLogModel(x) << " " << mvn.EvaluateModel(x) << " " << g(0) << " " << g(1) << " "
              << mvn.EvaluateLogModelNormalized(x) << " " << mvn.EvaluateModelNormalized(x) << " " << gp(0) << " " << gp(1) << std::endl;
    return 0;
}
