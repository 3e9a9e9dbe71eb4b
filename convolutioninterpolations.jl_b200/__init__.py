"""convolutioninterpolations.jl_b200 - B200 (sm_100a) implementation of the uniform-grid evaluation hot path
of ConvolutionInterpolations.jl behind the reference's own API names.

The directory name contains a dot, so import it through the `convitp_b200` shim at the repo root:

    import convitp_b200 as ci
    itp = ci.convolution_interpolation((x, y, z), vals, kernel="b5")
    ys = itp.eval(xq, yq, zq)
"""
from ._lib import (CIB200Error, DomainError, InexactError, UnsupportedModeError, SO_PATH, SYMBOLS, check, launch_count,
                   lib)
from . import sharding
from .api import (ConvolutionExtrapolation, FastConvolutionInterpolation, Flat, Line, LinearSubgridTableWarning, Natural, Throw,
                  convolution_interpolation, convolution_resample, convolution_smooth)

__all__ = ["convolution_interpolation", "convolution_resample", "convolution_smooth", "ConvolutionExtrapolation",
           "FastConvolutionInterpolation", "Throw", "Flat", "Line", "Natural", "CIB200Error", "DomainError",
           "InexactError", "UnsupportedModeError", "LinearSubgridTableWarning", "lib", "launch_count", "SO_PATH", "SYMBOLS", "sharding"]
