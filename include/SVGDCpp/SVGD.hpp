/* SVGD.hpp — the SVGD driver of the facade (reference SVGD.hpp:27-52 SVGDOptions, :84-511 SVGD).
 *
 * Same constructors, Initialize(), Run(), UpdateKernelParameters(), UpdateModelParameters(); Step()
 * is public here (protected in the reference, :368-373).  The body of Step() is one call into the
 * C ABI: median bandwidth, grad log p, the N x N interaction, the optimizer update and the bound
 * clamp all run on the GPU; the shared coordinate matrix is refreshed in place when Run()/Step()
 * return, keeping the reference's in-place contract (:393). */
#ifndef SVGDCPP_B200_SVGD_HPP
#define SVGDCPP_B200_SVGD_HPP

#include <cmath>
#include <exception>
#include <fstream>
#include <sstream>
#include <thread>

#include "Core.hpp"
#include "Kernel/Kernel.hpp"
#include "Model/Model.hpp"
#include "Optimizer/Optimizer.hpp"

struct SVGDOptions {
    size_t Dimension = 0;
    size_t NumIterations = 0;
    std::shared_ptr<Eigen::MatrixXd> CoordinateMatrixPtr = nullptr;
    std::shared_ptr<Kernel> KernelPtr = nullptr;
    std::shared_ptr<Model> ModelPtr = nullptr;
    std::shared_ptr<Optimizer> OptimizerPtr = nullptr;
    Eigen::VectorXd LowerBound = Eigen::VectorXd::Constant(1, -INFINITY);
    Eigen::VectorXd UpperBound = Eigen::VectorXd::Constant(1, INFINITY);
    std::string IntermediateMatricesOutputPath = "log.txt";
    bool Parallel = false;               /* accepted; the device path is always parallel */
    bool LogIntermediateMatrices = false; /* inspection path: K and grad K are formed on demand, one step at a time (small n) */
    int Device = 0;                                  /* CUDA device ordinal */
    int PrecisionMode = SVGDB_PRECISION_F64;         /* svgdb_precision */
    int Tc32Variant = SVGDB_TC32_AUTO;               /* svgdb_tc32_variant: arithmetic of the tensor-core pair kernel (TC32 mode only) */
    /* GPUs to shard the particle rows over (row blocks, NCCL all-gathers over NVLink; DESIGN.md "Multi-GPU").  Empty: `Device`
     * alone -- or, with `Parallel` set (the reference's "use what the machine has", SVGD.hpp:49, 239-249), every visible GPU.
     * More than one GPU: every call runs one host thread per GPU for its duration; the coordinate matrix stays one host matrix. */
    std::vector<int> Devices;
    SVGDOptions() {}
};

class SVGD {
public:
    SVGD(const SVGDOptions &o)
        : SVGD(o.Dimension, o.NumIterations, o.CoordinateMatrixPtr, o.KernelPtr, o.ModelPtr, o.OptimizerPtr, o.LowerBound, o.UpperBound,
               o.Parallel, o.LogIntermediateMatrices, o.IntermediateMatricesOutputPath, o.Device, o.PrecisionMode, o.Devices, o.Tc32Variant) {}

    SVGD(const size_t &dim, const size_t &iter, const std::shared_ptr<Eigen::MatrixXd> &coord_mat_ptr, const std::shared_ptr<Kernel> &kernel_ptr,
         const std::shared_ptr<Model> &model_ptr, const std::shared_ptr<Optimizer> &optimizer_ptr, const bool &parallel = false)
        : SVGD(dim, iter, coord_mat_ptr, kernel_ptr, model_ptr, optimizer_ptr, Eigen::VectorXd::Constant(1, -INFINITY),
               Eigen::VectorXd::Constant(1, INFINITY), parallel) {}

    SVGD(const size_t &dim, const size_t &iter, const std::shared_ptr<Eigen::MatrixXd> &coord_mat_ptr, const std::shared_ptr<Kernel> &kernel_ptr,
         const std::shared_ptr<Model> &model_ptr, const std::shared_ptr<Optimizer> &optimizer_ptr, const Eigen::VectorXd &bound_lower,
         const Eigen::VectorXd &bound_upper, const bool &parallel = false, const bool &log_intermediate_matrices = false,
         const std::string &intermediate_matrices_output_path = "log.txt", int device = 0, int precision_mode = SVGDB_PRECISION_F64,
         const std::vector<int> &devices = {}, int tc32_variant = SVGDB_TC32_AUTO)
        : num_iterations_(iter), parallel_(parallel), log_intermediate_matrices_(log_intermediate_matrices),
          intermediate_matrices_output_path_(intermediate_matrices_output_path)
    {
        if (!coord_mat_ptr) throw std::invalid_argument(SVGDCPP_LOG_PREFIX + "[Argument Error] Invalid coordinate matrix pointer.");
        dimension_ = static_cast<int>(coord_mat_ptr->rows());
        if (dimension_ != static_cast<int>(dim))
            throw DimensionMismatchException("Specified dimension does not match the particle coordinate matrix.");
        coord_matrix_ptr_ = coord_mat_ptr;

        const bool default_bounds = bound_lower.rows() == 1 && bound_upper.rows() == 1 && bound_lower(0) == -INFINITY && bound_upper(0) == INFINITY;
        check_bounds_ = !default_bounds;
        if (check_bounds_) {
            if (bound_lower.rows() != dimension_ && bound_lower.rows() != 1)
                throw DimensionMismatchException("The provided lower bounds have incorrect dimensions.");
            std::cout << SVGDCPP_LOG_PREFIX + "Bound checking enabled, lower bound set to " << bound_lower.transpose() << "." << std::endl;
            if (bound_upper.rows() != dimension_ && bound_upper.rows() != 1)
                throw DimensionMismatchException("The provided upper bounds have incorrect dimensions.");
            std::cout << SVGDCPP_LOG_PREFIX + "Bound checking enabled, upper bound set to " << bound_upper.transpose() << "." << std::endl;
        }
        kernel_ptr_ = kernel_ptr;
        model_ptr_ = model_ptr;
        optimizer_ptr_ = optimizer_ptr;
        if (kernel_ptr_ == nullptr) throw std::invalid_argument(SVGDCPP_LOG_PREFIX + "[Argument Error] Invalid Kernel object pointer.");
        if (model_ptr_ == nullptr) throw std::invalid_argument(SVGDCPP_LOG_PREFIX + "[Argument Error] Invalid Model object pointer.");
        if (optimizer_ptr_ == nullptr) throw std::invalid_argument(SVGDCPP_LOG_PREFIX + "[Argument Error] Invalid Optimizer object pointer.");

        /* which GPUs */
        devices_ = devices;
        if (devices_.empty()) {
            int visible = 0;
            if (parallel_ && svgdb_device_count(&visible) == SVGDB_OK && visible > 1)
                for (int k = 0; k < visible; ++k) devices_.push_back(k);
            else
                devices_.push_back(device);
        }
        if (static_cast<int64_t>(devices_.size()) > coord_matrix_ptr_->cols()) devices_.resize(static_cast<size_t>(coord_matrix_ptr_->cols()));
        if (model_ptr_->Dimension() != dimension_) throw DimensionMismatchException("Model dimension does not match the particle coordinate matrix.");
        const int world = static_cast<int>(devices_.size());
        ctxs_.assign(static_cast<size_t>(world), nullptr);
        unsigned char nccl_id[128] = {0};
        if (world > 1) svgdcpp_b200::ThrowOnError(svgdb_nccl_unique_id(nccl_id, sizeof(nccl_id)), "NCCL is not available for a multi-GPU run");
        std::vector<double> lb(dimension_), ub(dimension_);
        if (check_bounds_) {
            // a 1-row bound is replicated over every coordinate (reference replicate(1, n), :203,215)
            const int nb_l = static_cast<int>(bound_lower.rows()), nb_u = static_cast<int>(bound_upper.rows());
            for (int k = 0; k < dimension_; ++k) { lb[k] = bound_lower(nb_l == 1 ? 0 : k); ub[k] = bound_upper(nb_u == 1 ? 0 : k); }
        }
        try {
            OnEveryRank([&](int r) {
                svgdb_ctx *&c = ctxs_[static_cast<size_t>(r)];
                int rc = svgdb_create(&c, devices_[static_cast<size_t>(r)], static_cast<int64_t>(coord_matrix_ptr_->cols()), dimension_, precision_mode);
                if (rc != SVGDB_OK) svgdcpp_b200::ThrowOnError(rc, c ? svgdb_last_error(c) : "svgdb_create failed");
                if (world > 1) CheckOn(c, svgdb_comm_init(c, world, r, nccl_id, sizeof(nccl_id)));
                CheckOn(c, svgdb_set_tc32_variant(c, tc32_variant));
                if (check_bounds_) CheckOn(c, svgdb_set_bounds(c, lb.data(), ub.data(), dimension_));
                model_ptr_->Upload(c);
                kernel_ptr_->Upload(c);
                optimizer_ptr_->Upload(c);
            });
        } catch (...) {
            for (svgdb_ctx *c : ctxs_) svgdb_destroy(c);
            ctxs_.clear();
            throw;
        }
        ctx_ = ctxs_[0];
        if (parallel_)
            std::cout << SVGDCPP_LOG_PREFIX << "device path: all particle pairs run in parallel on " << world << " GPU" << (world > 1 ? "s." : ".") << std::endl;
    }

    SVGD(const SVGD &) = delete;
    SVGD &operator=(const SVGD &) = delete;
    ~SVGD()
    {
        for (svgdb_ctx *c : ctxs_) svgdb_destroy(c);
    }

    void Initialize()
    {
        model_ptr_->Initialize();
        kernel_ptr_->Initialize();
        optimizer_ptr_->Initialize();
        for (svgdb_ctx *c : ctxs_) CheckOn(c, svgdb_initialize(c));
    }

    void UpdateKernelParameters(const std::vector<Eigen::MatrixXd> &params)
    {
        kernel_ptr_->UpdateParameters(params);
        kernel_ptr_->Initialize();
        for (svgdb_ctx *c : ctxs_) kernel_ptr_->Upload(c);
    }

    void UpdateModelParameters(const std::vector<Eigen::MatrixXd> &params)
    {
        model_ptr_->UpdateParameters(params);
        model_ptr_->Initialize();
        for (svgdb_ctx *c : ctxs_) model_ptr_->Upload(c);
    }

    void Run()
    {
        if (!log_intermediate_matrices_) {
            Step(num_iterations_);
            return;
        }
        if (ctxs_.size() > 1)
            throw std::invalid_argument(SVGDCPP_LOG_PREFIX + "[Argument Error] LogIntermediateMatrices is an inspection path for one GPU.");
        /* Reference Run() with LogIntermediateMatrices (:338-366): after every step, the gradient / kernel / kernel-gradient matrices
         * that step used and the updated coordinates, in Eigen's default text format.  The device step never forms the n x n
         * matrices, so they are computed on demand from the pre-step particles (svgdb_compute_kernel_matrices; small n only). */
        const Eigen::Index n = coord_matrix_ptr_->cols();
        std::ostringstream log;
        for (size_t iter = 0; iter < num_iterations_; ++iter) {
            Eigen::MatrixXd grad(dimension_, n), kernel(n, n), kernel_grad(n * dimension_, n);
            Check(svgdb_set_particles(ctx_, coord_matrix_ptr_->data()));
            Check(svgdb_compute_log_model_grad(ctx_, grad.data()));
            Check(svgdb_compute_kernel_matrices(ctx_, kernel.data(), kernel_grad.data(), nullptr));
            Step(1);
            log << "========== Step " << iter + 1 << " =========="
                << "\nLogModelGrad=\n" << grad << "\n\nKernel=\n" << kernel << "\n\nKernelGrad=\n" << kernel_grad << "\n\nCoordMat=\n"
                << *coord_matrix_ptr_ << "\n\n";
        }
        std::ofstream output_file(intermediate_matrices_output_path_);
        if (!output_file)
            throw std::runtime_error(SVGDCPP_LOG_PREFIX + "[Runtime Error] Cannot open " + intermediate_matrices_output_path_ + " for writing.");
        output_file << log.str();
    }

    /* `iters` SVGD steps on the device; the coordinate matrix is read before and written after. */
    void Step(size_t iters = 1)
    {
        model_ptr_->Step();
        if (iters == 0) return;
        /* upload, iterate, download; the download of the last iteration overlaps its pair kernel (svgdb_step_host).  Several GPUs:
         * every rank moves its own block of rows of the (particle-contiguous) coordinate matrix, the rest travels over NVLink. */
        double *x = coord_matrix_ptr_->data();
        OnEveryRank([&](int r) {
            svgdb_ctx *c = ctxs_[static_cast<size_t>(r)];
            int64_t row0 = 0, n_rows = 0;
            CheckOn(c, svgdb_local_rows(c, &row0, &n_rows));
            double *mine = x + row0 * dimension_;
            CheckOn(c, svgdb_step_host(c, mine, mine, static_cast<int64_t>(iters)));
        });
    }

    /* ComputePhi of the reference (:407-454) for inspection: phi is dim x n. */
    Eigen::MatrixXd ComputePhi(double *scale_out = nullptr)
    {
        Eigen::MatrixXd phi(dimension_, coord_matrix_ptr_->cols());
        std::vector<Eigen::MatrixXd> others(ctxs_.size() > 1 ? ctxs_.size() - 1 : 0, Eigen::MatrixXd(dimension_, coord_matrix_ptr_->cols()));
        OnEveryRank([&](int r) { // every rank takes part in the collectives and receives the whole phi
            svgdb_ctx *c = ctxs_[static_cast<size_t>(r)];
            CheckOn(c, svgdb_set_particles(c, coord_matrix_ptr_->data()));
            CheckOn(c, svgdb_compute_phi(c, r == 0 ? phi.data() : others[static_cast<size_t>(r - 1)].data(), r == 0 ? scale_out : nullptr));
        });
        return phi;
    }

    int NumDevices() const { return static_cast<int>(ctxs_.size()); }

    svgdb_ctx *Context() { return ctx_; }

protected:
    void Check(int rc) { svgdcpp_b200::ThrowOnError(rc, svgdb_last_error(ctx_)); }
    static void CheckOn(svgdb_ctx *c, int rc) { svgdcpp_b200::ThrowOnError(rc, svgdb_last_error(c)); }

    /* Runs f(rank) for every GPU of this object: in place for one GPU, on one host thread per GPU otherwise (the library's
     * collectives need all ranks inside the call at the same time).  The first exception is rethrown on the caller's thread. */
    template <class F>
    void OnEveryRank(F f)
    {
        const int world = static_cast<int>(ctxs_.size());
        if (world <= 1) {
            f(0);
            return;
        }
        std::vector<std::exception_ptr> errors(static_cast<size_t>(world));
        std::vector<std::thread> threads;
        for (int r = 0; r < world; ++r)
            threads.emplace_back([&, r]() {
                try {
                    f(r);
                } catch (...) {
                    errors[static_cast<size_t>(r)] = std::current_exception();
                }
            });
        for (std::thread &t : threads) t.join();
        for (const std::exception_ptr &e : errors)
            if (e) std::rethrow_exception(e);
    }

    int dimension_ = -1;
    size_t num_iterations_;
    const bool parallel_ = false;
    bool log_intermediate_matrices_ = false;
    std::string intermediate_matrices_output_path_ = "log.txt";
    bool check_bounds_ = false;
    std::shared_ptr<Kernel> kernel_ptr_;
    std::shared_ptr<Model> model_ptr_;
    std::shared_ptr<Optimizer> optimizer_ptr_;
    std::shared_ptr<Eigen::MatrixXd> coord_matrix_ptr_;
    svgdb_ctx *ctx_ = nullptr;          /* rank 0 (the only one with a single GPU) */
    std::vector<svgdb_ctx *> ctxs_;     /* one context per GPU */
    std::vector<int> devices_;
};
#endif
