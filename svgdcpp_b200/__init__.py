"""svgdcpp_b200 — B200-native (sm_100a) SVGD inner loop behind the SVGDCpp API.

The package holds only what the hot path needs: `csrc/` (hand-written CUDA kernels and the C ABI of
include/svgd_b200.h, built into lib/libsvgd_b200.so) and a thin host-side mirror of the reference's
SVGD / Kernel / Model / Optimizer interface.  There is no CPU implementation: importing works
anywhere, but constructing an SVGD object requires the built library and a B200.
"""
from . import _capi
from .svgd import (SVGD, AdaGrad, Adam, DimensionMismatchException, GaussianRBFKernel, Model,
                   MultivariateNormal, Optimizer, RMSProp, ScaleMethod, SVGDOptions, UnsetException)

__all__ = ["SVGD", "SVGDOptions", "GaussianRBFKernel", "ScaleMethod", "Model", "MultivariateNormal",
           "Optimizer", "Adam", "AdaGrad", "RMSProp", "DimensionMismatchException", "UnsetException", "_capi"]
