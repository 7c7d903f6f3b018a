"""Host-side mirror of the SVGDCpp interface for the accelerated path, over the C ABI.

Same names, argument meaning and error behaviour as the reference classes so that parity tests read
like the reference's own programs (examples/multivariate_normal/mvn_example.cpp):

    x0 = 3 * eigen_random(dim, n)                      # dim x n, like Eigen::MatrixXd
    model = MultivariateNormal(mean, cov)              # Model/MultivariateNormal.hpp:39
    kernel = GaussianRBFKernel(x0, ScaleMethod.Median, model)   # Kernel/GaussianRBFKernel.hpp:47
    opt = AdaGrad(dim, n, 1e-1)                        # Optimizer/AdaGrad.hpp:31
    svgd = SVGD(dim, iters, x0, kernel, model, opt)    # SVGD.hpp:118
    svgd.Initialize(); svgd.Run()                      # x0 is updated in place (SVGD.hpp:393)

All numerics run in libsvgd_b200.so on the GPU; nothing here computes (the point evaluations of a
model exponentiate / scale the device's log p).  The C++ facade with the
identical API lives in include/SVGDCpp/.
"""
from __future__ import annotations

import ctypes as C
import enum
import math
import threading

import numpy as np

from . import _capi

_dp = C.POINTER(C.c_double)


class DimensionMismatchException(Exception):  # Exceptions.hpp:23-35
    def __init__(self, message):
        super().__init__("SVGDCpp: [Dimension Error] " + message)


class UnsetException(Exception):  # Exceptions.hpp:38-50
    def __init__(self, message):
        super().__init__("SVGDCpp: [Unset Error] " + message)


def _raise_for(code, ctx_handle):
    if code == _capi.OK:
        return
    msg = _capi.load().svgdb_last_error(ctx_handle)
    msg = msg.decode() if msg else "error %d" % code
    if code == _capi.ERR_DIMENSION:
        raise DimensionMismatchException(msg)
    if code == _capi.ERR_UNSET:
        raise UnsetException(msg)
    if code == _capi.ERR_INVALID:
        raise ValueError("SVGDCpp: " + msg)  # std::invalid_argument
    raise RuntimeError("SVGDCpp: [Runtime Error] " + msg)


def _ptr(a):
    return a.ctypes.data_as(_dp) if a is not None else None


# ---- models -------------------------------------------------------------------------------------
class Model:
    """Model base (Model/Model.hpp).  Only sums of MultivariateNormal objects (operator+,
    Model.hpp:55-92) and device-gradient hooks have a device implementation; an arbitrary taped
    lambda has none and raises UnsetException at SVGD construction (no CPU fallback)."""

    def __init__(self, dim):
        self.dimension_ = int(dim)
        self._components = []   # list of (mean, cov)
        self._hook = None       # (ctypes callback, user pointer)

    def __add__(self, other):
        if self.dimension_ != other.dimension_:
            raise DimensionMismatchException("Only models with the same variable dimensions can be added.")
        if (not self._components and self._hook is None) or (not other._components and other._hook is None):
            raise UnsetException("One of the model functions is unset; functional composition requires both model functions to be set.")
        if self._hook is not None or other._hook is not None:
            raise UnsetException("device-gradient hooks cannot be composed; only sums of MultivariateNormal are supported on the device")
        out = Model(self.dimension_)
        out._components = list(self._components) + list(other._components)
        return out

    def SetDeviceGradient(self, fn, user=None):
        """Device hook replacing Model::EvaluateLogModelGrad (Model.hpp:335-338); `fn` is a
        _capi.GRAD_FN (or a raw function pointer) enqueuing work on the given stream."""
        self._hook = (fn, user)
        self._components = []

    def GetParameters(self):
        out = []
        for mean, cov in self._components:
            out += [mean.copy().reshape(-1, 1), cov.copy()]
        return out

    def Initialize(self):
        pass

    def Step(self):
        pass

    # -- point evaluations (Model.hpp:290-338), on the device through a one-particle context: an inspection aid ------------
    def _evaluate_at(self, x, want_logp, want_grad):
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))
        if x.shape[0] != self.dimension_:
            raise DimensionMismatchException("Argument dimension does not match the model dimension.")
        lib = _capi.load()
        ctx = C.c_void_p()
        rc = lib.svgdb_create(C.byref(ctx), 0, 1, self.dimension_, _capi.PRECISION_F64)
        try:
            if rc != _capi.OK:
                _raise_for(rc, ctx)
            if self._hook is not None:
                fn, user = self._hook
                rc = lib.svgdb_set_model_device_hook(ctx, C.cast(fn, C.c_void_p), user)
            elif self._components:
                means = np.ascontiguousarray(np.stack([c[0] for c in self._components]))
                covs = np.ascontiguousarray(np.stack([c[1] for c in self._components]))
                rc = lib.svgdb_set_model_mvn_sum(ctx, means.shape[0], _ptr(means), _ptr(covs))
            else:
                raise UnsetException("Model function is unset.")
            if rc != _capi.OK:
                _raise_for(rc, ctx)
            rc = lib.svgdb_set_particles(ctx, _ptr(x))
            if rc != _capi.OK:
                _raise_for(rc, ctx)
            logp, grad = np.zeros(1), np.zeros(self.dimension_)
            if want_logp:
                rc = lib.svgdb_compute_log_model(ctx, _ptr(logp))
                if rc != _capi.OK:
                    _raise_for(rc, ctx)
            if want_grad:
                rc = lib.svgdb_compute_log_model_grad(ctx, _ptr(grad))
                if rc != _capi.OK:
                    _raise_for(rc, ctx)
            return float(logp[0]), grad
        finally:
            if ctx:
                lib.svgdb_destroy(ctx)

    def EvaluateLogModel(self, x):
        return self._evaluate_at(x, True, False)[0]

    def EvaluateModel(self, x):
        return math.exp(self.EvaluateLogModel(x))

    def EvaluateLogModelGrad(self, x):
        return self._evaluate_at(x, False, True)[1]

    def EvaluateModelGrad(self, x):
        logp, grad = self._evaluate_at(x, True, True)
        return math.exp(logp) * grad   # grad p = p grad log p


class MultivariateNormal(Model):
    def __init__(self, mean, covariance):
        mean = np.asarray(mean, dtype=np.float64).reshape(-1)
        covariance = np.asarray(covariance, dtype=np.float64)
        super().__init__(mean.shape[0])
        if covariance.ndim != 2 or covariance.shape[0] != mean.shape[0] or covariance.shape[1] != mean.shape[0]:
            raise DimensionMismatchException("Dimensions of parameter vectors/matrices do not match.")
        self._components = [(mean.copy(), covariance.copy())]
        self._compute_norm_const()

    def UpdateParameters(self, params):  # MultivariateNormal.hpp:94-115
        mean = np.asarray(params[0], dtype=np.float64).reshape(-1)
        covariance = np.asarray(params[1], dtype=np.float64)
        if covariance.ndim != 2 or covariance.shape[0] != mean.shape[0] or covariance.shape[1] != mean.shape[0]:
            raise DimensionMismatchException("Dimensions of parameter vectors/matrices do not match each other (# of rows must be equal).")
        if mean.shape[0] != self.dimension_:
            raise DimensionMismatchException("Dimensions of parameter vectors/matrices do not match original dimension.")
        self._components = [(mean.copy(), covariance.copy())]
        self._compute_norm_const()

    def _compute_norm_const(self):  # MultivariateNormal.hpp:182-186
        cov = self._components[0][1]
        with np.errstate(all="ignore"):
            det = np.float64(np.linalg.det(cov))
            self.norm_const_ = float(np.float64(1.0) / (np.float64(math.pow(2.0 * math.pi, self.dimension_ / 2.0)) * np.sqrt(det)))

    def GetNormalizationConstant(self):
        return self.norm_const_

    def EvaluateModelNormalized(self, x):  # MultivariateNormal.hpp:143-175
        return self.norm_const_ * self.EvaluateModel(x)

    def EvaluateLogModelNormalized(self, x):
        return math.log(self.norm_const_) + self.EvaluateLogModel(x)

    def EvaluateModelGradNormalized(self, x):
        return self.norm_const_ * self.EvaluateModelGrad(x)


# ---- kernel -------------------------------------------------------------------------------------
class ScaleMethod(enum.IntEnum):  # GaussianRBFKernel.hpp:25-30 (+ a constant scale)
    Median = 0
    Hessian = 1
    Fixed = 2


class GaussianRBFKernel:
    ScaleMethod = ScaleMethod

    def __init__(self, coord_mat, method=ScaleMethod.Median, model=None, fixed_scale=0.0):
        if method == ScaleMethod.Hessian and model is None:
            raise UnsetException("Hessian-based scale requires a model.")
        self.coord_matrix_ = coord_mat
        self.dimension_ = int(np.asarray(coord_mat).shape[0])
        self.scale_method_ = ScaleMethod(method)
        self.fixed_scale_ = float(fixed_scale)
        self.target_model_ = model
        self.kernel_parameters_ = []

    def UpdateParameters(self, params):
        """Kernel::UpdateParameters.  With ScaleMethod.Median / Hessian the reference recomputes the scale from the particles at every
        Step and overwrites whatever was set (GaussianRBFKernel.hpp:141-156), so the parameters only take effect for the constant
        scale (ScaleMethod.Fixed), where params[0] must be a I."""
        A = np.atleast_2d(np.asarray(params[0], dtype=np.float64))
        d = self.dimension_
        if A.shape != (d, d):
            raise DimensionMismatchException("Kernel parameter matrix must be %d x %d." % (d, d))
        self.kernel_parameters_ = [A.copy()]
        if self.scale_method_ == ScaleMethod.Fixed:
            a = float(A[0, 0])
            if not np.array_equal(A, a * np.eye(d)):
                raise ValueError("SVGDCpp: [Argument Error] the device RBF kernel takes a scalar scale: A must be a * I.")
            self.fixed_scale_ = a


# ---- optimizers ---------------------------------------------------------------------------------
class Optimizer:
    def __init__(self, lr, epsilon=1.0e-8):
        self.learning_rate_ = float(lr)
        self.stabilizer_ = float(epsilon)


class Adam(Optimizer):  # Optimizer/Adam.hpp:33-49
    kind = _capi.OPT_ADAM

    def __init__(self, dimension, num_particles, lr, beta1, beta2, epsilon=1.0e-8):
        super().__init__(lr, epsilon)
        if beta1 >= 1.0 or beta1 < 0.0 or beta2 >= 1.0 or beta2 < 0.0:
            raise ValueError("SVGDCpp: [Argument Error] Invalid value for decay parameter beta.")
        self.dimension_, self.num_particles_ = int(dimension), int(num_particles)
        self.decay_rate_1_, self.decay_rate_2_ = float(beta1), float(beta2)


class AdaGrad(Optimizer):  # Optimizer/AdaGrad.hpp:31-37
    kind = _capi.OPT_ADAGRAD

    def __init__(self, dimension, num_particles, lr, epsilon=1.0e-8):
        super().__init__(lr, epsilon)
        self.dimension_, self.num_particles_ = int(dimension), int(num_particles)
        self.decay_rate_1_ = self.decay_rate_2_ = 0.0


class RMSProp(Optimizer):  # Optimizer/RMSProp.hpp:33-46
    kind = _capi.OPT_RMSPROP

    def __init__(self, dimension, num_particles, lr, beta, epsilon=1.0e-8):
        super().__init__(lr, epsilon)
        if beta > 1.0 or beta < 0.0:
            raise ValueError("SVGDCpp: [Argument Error] Invalid value for decay parameter beta.")
        self.dimension_, self.num_particles_ = int(dimension), int(num_particles)
        self.decay_rate_1_, self.decay_rate_2_ = float(beta), 0.0


# ---- driver -------------------------------------------------------------------------------------
def _eigen_str(M):
    """Eigen's default operator<< for a matrix: 6 significant digits, columns right-aligned to the widest entry."""
    M = np.atleast_2d(np.asarray(M, dtype=np.float64))
    cells = [["%.6g" % v for v in row] for row in M]
    width = max(len(c) for row in cells for c in row)
    return "\n".join(" ".join(c.rjust(width) for c in row) for row in cells)


class SVGDOptions:  # SVGD.hpp:27-52 (+ device / precision / sharding fields with defaults)
    def __init__(self):
        self.Dimension = 0
        self.NumIterations = 0
        self.CoordinateMatrixPtr = None
        self.KernelPtr = None
        self.ModelPtr = None
        self.OptimizerPtr = None
        self.LowerBound = np.array([-np.inf])
        self.UpperBound = np.array([np.inf])
        self.IntermediateMatricesOutputPath = "log.txt"
        self.Parallel = False
        self.LogIntermediateMatrices = False
        self.Device = 0
        self.Precision = _capi.PRECISION_F64
        self.Tc32Variant = _capi.TC32_AUTO
        self.Devices = []   # GPUs to shard the particle rows over; empty: Device alone, or every visible GPU with Parallel set


class SVGD:
    """SVGD driver (SVGD.hpp:84-511).  `coord_mat` is the dim x n particle matrix, updated in place
    by Run()/Step() exactly like the reference's shared Eigen::MatrixXd."""

    def __init__(self, dim, iter=None, coord_mat=None, kernel=None, model=None, optimizer=None,
                 bound_lower=None, bound_upper=None, parallel=False, log_intermediate_matrices=False,
                 intermediate_matrices_output_path="log.txt", device=0, precision=_capi.PRECISION_F64, tc32_variant=_capi.TC32_AUTO,
                 devices=None):
        if isinstance(dim, SVGDOptions):
            o = dim
            dim, iter, coord_mat, kernel, model, optimizer = (o.Dimension, o.NumIterations, o.CoordinateMatrixPtr,
                                                              o.KernelPtr, o.ModelPtr, o.OptimizerPtr)
            bound_lower, bound_upper, parallel = o.LowerBound, o.UpperBound, o.Parallel
            log_intermediate_matrices, intermediate_matrices_output_path = o.LogIntermediateMatrices, o.IntermediateMatricesOutputPath
            device, precision, tc32_variant, devices = o.Device, o.Precision, o.Tc32Variant, list(o.Devices)
        self._lib = _capi.load()
        self._ctx = C.c_void_p()
        if coord_mat is None:
            raise ValueError("SVGDCpp: [Argument Error] Invalid coordinate matrix pointer.")
        self.coord_matrix_ = coord_mat
        self.dimension_ = int(coord_mat.shape[0])
        self.num_particles_ = int(coord_mat.shape[1])
        self.num_iterations_ = int(iter)
        self.parallel_ = bool(parallel)  # accepted for source compatibility; the GPU path is always parallel
        if self.dimension_ != int(dim):
            raise DimensionMismatchException("Specified dimension does not match the particle coordinate matrix.")
        lb = np.asarray(bound_lower if bound_lower is not None else [-np.inf], dtype=np.float64).reshape(-1)
        ub = np.asarray(bound_upper if bound_upper is not None else [np.inf], dtype=np.float64).reshape(-1)
        self.check_bounds_ = not (lb.size == 1 and ub.size == 1 and lb[0] == -np.inf and ub[0] == np.inf)
        if self.check_bounds_:
            if lb.size not in (1, self.dimension_):
                raise DimensionMismatchException("The provided lower bounds have incorrect dimensions.")
            if ub.size not in (1, self.dimension_):
                raise DimensionMismatchException("The provided upper bounds have incorrect dimensions.")
        if kernel is None:
            raise ValueError("SVGDCpp: [Argument Error] Invalid Kernel object pointer.")
        if model is None:
            raise ValueError("SVGDCpp: [Argument Error] Invalid Model object pointer.")
        if optimizer is None:
            raise ValueError("SVGDCpp: [Argument Error] Invalid Optimizer object pointer.")
        self.log_intermediate_matrices_ = bool(log_intermediate_matrices)
        self.intermediate_matrices_output_path_ = intermediate_matrices_output_path
        self.kernel_, self.model_, self.optimizer_ = kernel, model, optimizer

        # which GPUs: `devices`, else `device` alone -- or every visible GPU when the reference's `parallel` flag is set (SVGD.hpp:49, 239-249)
        devs = [int(x) for x in devices] if devices else []
        if not devs:
            cnt = C.c_int(0)
            if self.parallel_ and self._lib.svgdb_device_count(C.byref(cnt)) == _capi.OK and cnt.value > 1:
                devs = list(range(cnt.value))
            else:
                devs = [int(device)]
        devs = devs[:max(1, self.num_particles_)]
        world = len(devs)
        self._ctxs = [C.c_void_p() for _ in devs]
        self._ctx = self._ctxs[0]
        uid = (C.c_ubyte * 128)()
        if world > 1 and self._lib.svgdb_nccl_unique_id(C.cast(uid, C.c_void_p), 128) != _capi.OK:
            raise RuntimeError("SVGDCpp: [Runtime Error] NCCL is not available for a multi-GPU run")
        lbf = ubf = None
        if self.check_bounds_:
            lbf = np.ascontiguousarray(np.broadcast_to(lb, (self.dimension_,)) if lb.size == 1 else lb)
            ubf = np.ascontiguousarray(np.broadcast_to(ub, (self.dimension_,)) if ub.size == 1 else ub)
        o = optimizer

        def setup(r):
            ctx = self._ctxs[r]
            rc = self._lib.svgdb_create(C.byref(ctx), devs[r], self.num_particles_, self.dimension_, int(precision))
            self._check_on(ctx, rc)
            if world > 1:
                self._check_on(ctx, self._lib.svgdb_comm_init(ctx, world, r, C.cast(uid, C.c_void_p), 128))
            self._check_on(ctx, self._lib.svgdb_set_tc32_variant(ctx, int(tc32_variant)))
            if self.check_bounds_:
                self._check_on(ctx, self._lib.svgdb_set_bounds(ctx, _ptr(lbf), _ptr(ubf), self.dimension_))
            self._push_model(ctx)
            self._push_kernel(ctx)
            self._check_on(ctx, self._lib.svgdb_set_optimizer(ctx, o.kind, o.learning_rate_, o.decay_rate_1_, o.decay_rate_2_, o.stabilizer_))

        try:
            self._on_every_rank(setup)
        except Exception:
            for ctx in self._ctxs:
                if ctx:
                    self._lib.svgdb_destroy(ctx)
            self._ctxs = []
            self._ctx = C.c_void_p()
            raise
        # pinned staging buffer (particle-major): the device-to-host copy of a step overlaps its pair kernel only from pinned memory
        self._hp = C.c_void_p()
        count = self.num_particles_ * self.dimension_
        if self._lib.svgdb_host_alloc(C.byref(self._hp), count * 8) != _capi.OK:
            raise MemoryError("pinned host allocation of %d bytes failed" % (count * 8))
        self._host = np.ctypeslib.as_array((C.c_double * count).from_address(self._hp.value)).reshape(self.num_particles_, self.dimension_)
        self._dirty_host = True

    # -- plumbing ------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != _capi.OK:
            _raise_for(rc, self._ctx)

    @staticmethod
    def _check_on(ctx, rc):
        if rc != _capi.OK:
            _raise_for(rc, ctx)

    def _on_every_rank(self, fn):
        """fn(rank) for every GPU of this object: in place for one, on one host thread per GPU otherwise (the library's collectives
        need all ranks inside the call at once; ctypes releases the GIL during the calls).  The first exception is re-raised here."""
        world = len(self._ctxs)
        if world <= 1:
            fn(0)
            return
        errors = [None] * world

        def run(r):
            try:
                fn(r)
            except BaseException as e:  # noqa: BLE001 - re-raised below
                errors[r] = e

        threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for e in errors:
            if e is not None:
                raise e

    def _push_model(self, ctx=None):
        if ctx is None:
            for c in self._ctxs:
                self._push_model(c)
            return
        m = self.model_
        if m._hook is not None:
            fn, user = m._hook
            self._hook_keepalive = fn
            self._check_on(ctx, self._lib.svgdb_set_model_device_hook(ctx, C.cast(fn, C.c_void_p), user))
            return
        if not m._components:
            raise UnsetException("Model function is unset.")
        means = np.ascontiguousarray(np.stack([c[0] for c in m._components]))          # C x d
        covs = np.ascontiguousarray(np.stack([c[1] for c in m._components]))           # C x d x d
        if means.shape[1] != self.dimension_:
            raise DimensionMismatchException("Model dimension does not match the particle coordinate matrix.")
        self._check_on(ctx, self._lib.svgdb_set_model_mvn_sum(ctx, means.shape[0], _ptr(means), _ptr(covs)))

    def _push_kernel(self, ctx=None):
        if ctx is None:
            for c in self._ctxs:
                self._push_kernel(c)
            return
        k = self.kernel_
        if k.dimension_ != self.dimension_:
            raise DimensionMismatchException("Kernel dimension does not match the particle coordinate matrix.")
        self._check_on(ctx, self._lib.svgdb_set_kernel_rbf(ctx, int(k.scale_method_), k.fixed_scale_))

    def _upload(self):
        self._host[...] = np.asarray(self.coord_matrix_, dtype=np.float64).T
        for ctx in self._ctxs:
            self._check_on(ctx, self._lib.svgdb_set_particles(ctx, _ptr(self._host)))

    def _download(self):
        self._check(self._lib.svgdb_get_particles(self._ctx, _ptr(self._host)))
        self.coord_matrix_[...] = self._host.T

    def NumDevices(self):
        return len(self._ctxs)

    # -- reference API ---------------------------------------------------------------------------
    def Initialize(self):  # SVGD.hpp:268-296
        self.model_.Initialize()
        for ctx in self._ctxs:
            self._check_on(ctx, self._lib.svgdb_initialize(ctx))

    def UpdateKernelParameters(self, params):  # SVGD.hpp:304-321: params[0] = A (= a I for the constant scale)
        self.kernel_.UpdateParameters(params)
        self._push_kernel()

    def UpdateModelParameters(self, params):  # SVGD.hpp:328-332
        self.model_.UpdateParameters(params)
        self._push_model()

    def Step(self, iters=1):  # SVGD.hpp:373-400 (protected there; public here)
        if int(iters) < 1:
            return
        self._host[...] = np.asarray(self.coord_matrix_, dtype=np.float64).T

        def step(r):  # every rank moves its own block of rows of the staging buffer; the rest travels over NVLink
            ctx = self._ctxs[r]
            r0, nr = C.c_int64(0), C.c_int64(0)
            self._check_on(ctx, self._lib.svgdb_local_rows(ctx, C.byref(r0), C.byref(nr)))
            mine = self._host[r0.value:r0.value + nr.value]
            self._check_on(ctx, self._lib.svgdb_step_host(ctx, _ptr(mine), _ptr(mine), int(iters)))

        self._on_every_rank(step)
        self.coord_matrix_[...] = self._host.T

    def Run(self):  # SVGD.hpp:338-366
        if not self.log_intermediate_matrices_:
            self.Step(self.num_iterations_)
            return
        if len(self._ctxs) > 1:
            raise ValueError("SVGDCpp: [Argument Error] LogIntermediateMatrices is an inspection path for one GPU.")
        # LogIntermediateMatrices: one step at a time, the matrices of that step formed on demand (inspection path, small n)
        chunks = []
        for it in range(self.num_iterations_):
            G = self.EvaluateLogModelGrad()
            K, dK = self.ComputeKernelMatrices()
            self.Step(1)
            chunks.append("========== Step %d ==========\nLogModelGrad=\n%s\n\nKernel=\n%s\n\nKernelGrad=\n%s\n\nCoordMat=\n%s\n\n"
                          % (it + 1, _eigen_str(G), _eigen_str(K), _eigen_str(dK), _eigen_str(self.coord_matrix_)))
        try:
            with open(self.intermediate_matrices_output_path_, "w") as f:
                f.write("".join(chunks))
        except OSError:
            raise RuntimeError("SVGDCpp: [Runtime Error] Cannot open %s for writing." % self.intermediate_matrices_output_path_)

    def ComputeKernelMatrices(self):
        """kernel_matrix_ (n x n, [j, i] = k(x_j, x_i)) and kernel_grad_matrix_ ((n dim) x n, rows j dim .. j dim + dim of column i =
        grad k(x_j, x_i)) of SVGD::ComputePhi (SVGD.hpp:434-448) for the current particles; small n only."""
        self._upload()
        n, d = self.num_particles_, self.dimension_
        K = np.empty((n, n))          # C order [i, j] == Eigen column-major (j, i)
        dK = np.empty((n, n * d))     # C order [i, j d + c] == Eigen column-major (j d + c, i)
        self._check(self._lib.svgdb_compute_kernel_matrices(self._ctx, _ptr(K), _ptr(dK), None))
        return K.T.copy(), dK.T.copy()

    # -- extras used by tests / bench ---------------------------------------------------------------
    def ComputePhi(self):
        """SVGD::ComputePhi (SVGD.hpp:407-454) on the current coordinates: (phi as dim x n, scale a)."""
        self._upload()
        world = len(self._ctxs)
        phis = [np.empty((self.num_particles_, self.dimension_), dtype=np.float64) for _ in range(world)]
        scales = [C.c_double(0.0) for _ in range(world)]
        self._on_every_rank(lambda r: self._check_on(self._ctxs[r], self._lib.svgdb_compute_phi(self._ctxs[r], _ptr(phis[r]), C.byref(scales[r]))))
        return phis[0].T.copy(), scales[0].value

    def GetScaleMatrix(self):
        """The kernel's inverse scale matrix (GaussianRBFKernel::GetParameters()[0]) of the last Step / Compute call."""
        A = np.zeros((self.dimension_, self.dimension_))
        self._check(self._lib.svgdb_get_scale_matrix(self._ctx, _ptr(A)))
        return A

    def ComputeScale(self):
        self._upload()
        scales = [C.c_double(0.0) for _ in self._ctxs]
        self._on_every_rank(lambda r: self._check_on(self._ctxs[r], self._lib.svgdb_compute_scale(self._ctxs[r], C.byref(scales[r]))))
        return scales[0].value

    def EvaluateLogModelGrad(self):
        self._upload()
        Gs = [np.empty((self.num_particles_, self.dimension_), dtype=np.float64) for _ in self._ctxs]
        self._on_every_rank(lambda r: self._check_on(self._ctxs[r], self._lib.svgdb_compute_log_model_grad(self._ctxs[r], _ptr(Gs[r]))))
        return Gs[0].T.copy()

    def Stats(self):
        st = _capi.Stats()
        self._check(self._lib.svgdb_get_stats(self._ctx, C.byref(st)))
        return {name: getattr(st, name) for name, _ in st._fields_}

    def context(self):
        return self._ctx

    def close(self):
        for ctx in getattr(self, "_ctxs", []):
            if ctx:
                self._lib.svgdb_destroy(ctx)
        self._ctxs = []
        if getattr(self, "_ctx", None) is not None:
            self._ctx = C.c_void_p()
        if getattr(self, "_hp", None) is not None and self._hp:
            self._host = None
            self._lib.svgdb_host_free(self._hp)
            self._hp = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
