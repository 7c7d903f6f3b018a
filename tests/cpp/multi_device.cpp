// SVGDOptions::Devices through the facade: the same run on one GPU and sharded over the GPUs given on the command line (one host
// thread per GPU inside Run(), rows of the shared coordinate matrix sharded, NCCL over NVLink).  Prints both finals with 17 digits;
// tests/test_gpu_multi.py checks them against each other and against the oracle.
#include <cstdlib>
#include <iomanip>
#include <iostream>

#include "Core"
#include "Kernel"
#include "Model"
#include "Optimizer"

int main(int argc, char **argv)
{
    const size_t dim = 3, num_particles = 700, num_iterations = 6;
    Eigen::VectorXd mean(dim);
    mean << 0.5, -0.25, 1.0;
    Eigen::MatrixXd covariance(dim, dim);
    covariance << 0.9, 0.2, -0.1, 0.2, 0.8, 0.3, -0.1, 0.3, 1.2;
    Eigen::MatrixXd start(dim, num_particles);
    unsigned long long state = 12345;
    for (size_t j = 0; j < num_particles; ++j)
        for (size_t r = 0; r < dim; ++r) { // a fixed LCG: the test regenerates the same numbers
            state = state * 6364136223846793005ULL + 1442695040888963407ULL;
            start(r, j) = 4.0 * (static_cast<double>(state >> 11) / 9007199254740992.0) - 2.0;
        }
    std::cout << std::setprecision(17);
    for (int pass = 0; pass < 2; ++pass) {
        auto x0 = std::make_shared<Eigen::MatrixXd>(start);
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
        if (pass == 1)
            for (int a = 1; a < argc; ++a) options.Devices.push_back(std::atoi(argv[a]));
        SVGD svgd(options);
        svgd.Initialize();
        svgd.Run();
        std::cout << "devices " << svgd.NumDevices() << "\n";
        for (size_t r = 0; r < dim; ++r) {
            for (size_t j = 0; j < num_particles; ++j) std::cout << (*x0)(r, j) << (j + 1 < num_particles ? " " : "\n");
        }
    }
    return 0;
}
