"""Writes tests/golden/reference_goldens.json: the golden vectors the reference's OWN unit
tests hold for the ALS path, transcribed from (paths relative to /root/reference):

  online/test/net/myrrix/online/factorizer/als/AlternatingLeastSquaresTest.java:39-57   (testALS)
  online/test/net/myrrix/online/factorizer/als/AlternatingLeastSquaresTest.java:60-78   (testALSPredictingR)
      inputs :92-113, run config :115 (k=2, threshold 1e-4, max 40 iterations)
  online/test/net/myrrix/online/factorizer/als/NegativeInputTest.java:37-80
  common/test/net/myrrix/common/math/MatrixUtilsTest.java:63-73
  common/test/net/myrrix/common/math/SimpleVectorMathTest.java:26-37
  common/test/net/myrrix/common/MyrrixTest.java:34-35  (FLOAT_EPSILON 1e-6, DOUBLE_EPSILON 1e-12)

The reference is Java and cannot run in this image (no JVM), so these are the only
reference-produced numbers available; this script only re-emits them as JSON.
Run: python tests/golden/make_golden.py
"""
import json
import os

G = {
    "float_epsilon": 1e-6,
    "double_epsilon": 1e-12,
    "als": {
        "R": [[0, 2, 3, 1, 0], [0, 0, 4, 5, 0], [1, 0, 0, 0, 2], [3, 0, 1, 0, 5], [0, 2, 2, 2, 0]],
        "Y0": [[0.1, 0.2], [0.2, 0.5], [0.3, 0.1], [0.2, 0.2], [0.5, 0.4]],
        "features": 2, "threshold": 0.0001, "max_iterations": 40,
        "product": [[-0.030258, 0.852781, 1.004839, 1.024087, -0.036206],
                    [0.077046, 0.751232, 0.949796, 0.910322, 0.073047],
                    [0.916777, -0.196005, 0.335926, -0.163591, 0.929028],
                    [0.987400, 0.130943, 0.772403, 0.235522, 0.998354],
                    [-0.028683, 0.850540, 1.003130, 1.021514, -0.034598]],
        "product_reconstruct_r": [[0.0678369, 0.6574759, 2.1020291, 2.0976211, 0.1115919],
                                  [-0.0176293, 1.3062225, 4.1365933, 4.1739127, -0.0380586],
                                  [1.0854513, -0.0344434, 0.1725342, -0.1564803, 1.8502977],
                                  [2.8377915, 0.0528524, 0.9041158, 0.0474437, 4.8365208],
                                  [-0.0057799, 0.6608552, 2.0936351, 2.1115670, -0.0139042]],
    },
    "negative_input": {
        "R": [[1, 1, 1, 0], [0, -1, 1, 1], [-1, 0, 0, 1]],
        "Y0": [[0.1, 0.2], [0.2, 0.5], [0.3, 0.1], [0.2, 0.2]],
        "features": 2, "threshold": 0.0001, "max_iterations": 40,
        "product": [[0.899032, 0.900162, 0.990150, -0.026642],
                    [0.181214, 0.089988, 0.787198, 1.012226],
                    [-0.104165, -0.178240, 0.360391, 0.825856]],
    },
    "transpose_times_self": {
        "M": [[4.0, -1.0, -5.0], [2.0, 0.0, 3.0]],
        "MTM": [[20.0, -4.0, -14.0], [-4.0, 1.0, 5.0], [-14.0, 5.0, 34.0]],
    },
    "simple_vector_math": {
        "vec1": [-1.0, 2.5, 3.0], "vec2": [1.5, -1.5, 0.0],
        "dot": -5.25, "norm1": 4.03112887414928, "norm2": 2.12132034355964,
    },
}

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.json")
    with open(out, "w") as f:
        json.dump(G, f, indent=1)
    print(out)
