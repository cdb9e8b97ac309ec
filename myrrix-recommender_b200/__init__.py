"""B200-native ALS factorization core: drop-in for Myrrix's MatrixFactorizer path.

The product is libmyrrix_als.so (include/myrrix_als.h, csrc/*.cu*).  This package is
the host-side mirror of the reference interface
(online/src/net/myrrix/online/factorizer/MatrixFactorizer.java:31-77,
 online/src/net/myrrix/online/factorizer/als/AlternatingLeastSquares.java:66) that the
parity tests and bench.py drive.  No CPU compute path exists here.
"""
from . import _native, foldin, ingest, initial_y, model_io, recommender
from .factorizer import (AlternatingLeastSquares, MatrixFactorizer, NativeALS,
                         SingularMatrixSolverException, SolverException, properties)

__all__ = ["AlternatingLeastSquares", "MatrixFactorizer", "NativeALS",
           "SingularMatrixSolverException", "SolverException", "properties", "_native", "ingest", "foldin", "model_io",
           "initial_y", "recommender"]
