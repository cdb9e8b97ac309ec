/*
 * Reference-side binding a Myrrix maintainer would add (NOT compiled in this repo's image:
 * there is no JDK here; see INTEGRATION.md).  Implements the reference's plug-in interface
 *   net.myrrix.online.factorizer.MatrixFactorizer  (online/.../factorizer/MatrixFactorizer.java:31-77)
 * over the C ABI in include/myrrix_als.h via the JNI stub in bindings/jni/myrrix_als_jni.c.
 * The only production change is DelegateGenerationManager.java:406-410, which constructs this
 * class instead of AlternatingLeastSquares (same constructor arguments).
 */
package net.myrrix.online.factorizer.als;

import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.nio.FloatBuffer;
import java.nio.IntBuffer;
import java.nio.LongBuffer;
import java.util.concurrent.ExecutionException;

import org.apache.commons.math3.util.FastMath;

import net.myrrix.common.LangUtils;
import net.myrrix.common.collection.FastByIDFloatMap;
import net.myrrix.common.collection.FastByIDMap;
import net.myrrix.common.math.SingularMatrixSolverException;
import net.myrrix.common.math.SimpleVectorMath;
import net.myrrix.common.random.RandomManager;
import net.myrrix.common.random.RandomUtils;
import net.myrrix.common.stats.DoubleWeightedMean;
import net.myrrix.online.factorizer.MatrixFactorizer;

public final class CudaAlternatingLeastSquares implements MatrixFactorizer {

  static {
    System.loadLibrary("myrrix_als_jni"); // links libmyrrix_als.so
  }

  // status codes of include/myrrix_als.h
  private static final int ALS_OK = 0;
  private static final int ALS_E_SINGULAR = 1;

  private final FastByIDMap<FastByIDFloatMap> RbyRow;
  private final FastByIDMap<FastByIDFloatMap> RbyColumn;
  private final int features;
  private final double estimateErrorConvergenceThreshold;
  private final int maxIterations;
  private FastByIDMap<float[]> X;
  private FastByIDMap<float[]> Y;
  private FastByIDMap<float[]> previousY;

  public CudaAlternatingLeastSquares(FastByIDMap<FastByIDFloatMap> RbyRow,
                                     FastByIDMap<FastByIDFloatMap> RbyColumn,
                                     int features,
                                     double estimateErrorConvergenceThreshold,
                                     int maxIterations) {
    // same preconditions as AlternatingLeastSquares.java:137-141
    if (RbyRow == null || RbyColumn == null) {
      throw new NullPointerException();
    }
    if (features <= 0) {
      throw new IllegalArgumentException("features must be positive: " + features);
    }
    if (!(estimateErrorConvergenceThreshold > 0.0 && estimateErrorConvergenceThreshold < 1.0)) {
      throw new IllegalArgumentException("threshold must be in (0,1): " + estimateErrorConvergenceThreshold);
    }
    this.RbyRow = RbyRow;
    this.RbyColumn = RbyColumn;
    this.features = features;
    this.estimateErrorConvergenceThreshold = estimateErrorConvergenceThreshold;
    this.maxIterations = maxIterations;
  }

  @Override public FastByIDMap<float[]> getX() { return X; }
  @Override public FastByIDMap<float[]> getY() { return Y; }
  @Override public void setPreviousX(FastByIDMap<float[]> previousX) { /* ignored, as ALS.java:162-165 */ }
  @Override public void setPreviousY(FastByIDMap<float[]> previousY) { this.previousY = previousY; }

  // ---- native methods (bindings/jni/myrrix_als_jni.c), one per C-ABI entry point used ----
  private static native long nCreate(int features, double alpha, double lambda, boolean reconstructR,
                                     boolean lossIgnoresUnspecified, double singularityThreshold, int device);
  private static native void nDestroy(long handle);
  private static native int nSetInteractions(long handle, long nUsers, long nItems,
                                             ByteBuffer rowPtr, ByteBuffer colIdx, ByteBuffer val);
  private static native int nSetInteractionsByColumn(long handle, ByteBuffer colPtr, ByteBuffer rowIdx, ByteBuffer val);
  private static native int nSetY(long handle, ByteBuffer y);
  private static native int nHalfX(long handle);
  private static native int nHalfY(long handle);
  private static native int nSync(long handle);
  private static native int nProbe(long handle, int[] users, int[] items, double[] out);
  private static native int nGetX(long handle, ByteBuffer out);
  private static native int nGetY(long handle, ByteBuffer out);
  private static native String nLastError(long handle);
  private static native int nSingularRank(long handle);

  @Override
  public Void call() throws ExecutionException, InterruptedException {
    boolean randomY = previousY == null || previousY.isEmpty();
    // constructInitialY (ALS.java:264-335) stays in Java: reuse the reference's own code path
    // for the random / feature-count-change cases, then flatten.
    FastByIDMap<float[]> initialY = ReferenceInitialY.construct(previousY, RbyColumn, features);

    // long ID -> dense index, slot-order walk of the keys (FastByIDMap.java:499-533)
    long[] userIDs = keys(RbyRow);
    long[] itemIDs = keysOfFactors(initialY);
    FastByIDMap<Integer> itemIndex = index(itemIDs);
    FastByIDMap<Integer> userIndex = index(userIDs);

    ByteBuffer[] byRow = flatten(RbyRow, userIDs, itemIndex);       // row_ptr, col_idx, val
    ByteBuffer[] byCol = flatten(RbyColumn, itemIDs, userIndex);    // items without entries: empty rows
    ByteBuffer y0 = directFloats((long) itemIDs.length * features);
    FloatBuffer y0f = y0.asFloatBuffer();
    for (long itemID : itemIDs) {
      y0f.put(initialY.get(itemID));
    }

    long h = nCreate(features, alpha(), lambda(),
                     Boolean.parseBoolean(System.getProperty("model.reconstructRMatrix", "false")),
                     Boolean.parseBoolean(System.getProperty("model.lossIgnoresUnspecified", "false")),
                     Double.parseDouble(System.getProperty("common.matrix.singularityThreshold", "1.0e-5")),
                     Integer.getInteger("model.cuda.device", 0));
    try {
      check(h, nSetInteractions(h, userIDs.length, itemIDs.length, byRow[0], byRow[1], byRow[2]));
      check(h, nSetInteractionsByColumn(h, byCol[0], byCol[1], byCol[2]));
      check(h, nSetY(h, y0));

      if (!Boolean.parseBoolean(System.getProperty("model.als.iterate", "true"))) {  // ALS.java:196-204
        check(h, nHalfX(h));
        check(h, nSync(h));
        copyOut(h, userIDs, itemIDs);
        return null;
      }

      int[] testUsers = sample(userIDs, userIndex);   // RandomUtils.chooseAboutNFromStream, :207-214
      int[] testItems = sample(keys(RbyColumn), itemIndex);
      double[] estimates = new double[testUsers.length * testItems.length];  // X empty: zeros (:215-223)
      double[] fresh = new double[estimates.length];
      int iterationNumber = 0;
      while (true) {
        check(h, nHalfX(h));   // iterateXFromY
        check(h, nHalfY(h));   // iterateYFromX
        check(h, nSync(h));    // surfaces ALS_E_SINGULAR exactly where Future.get() would (:349)
        check(h, nProbe(h, testUsers, testItems, fresh));
        DoubleWeightedMean averageAbsoluteEstimateDiff = new DoubleWeightedMean();
        for (int i = 0; i < fresh.length; i++) {
          averageAbsoluteEstimateDiff.increment(FastMath.abs(fresh[i] - estimates[i]), FastMath.max(0.0, fresh[i]));
          estimates[i] = fresh[i];
        }
        iterationNumber++;
        if (maxIterations > 0 && iterationNumber >= maxIterations) {
          break;
        }
        double convergenceValue = averageAbsoluteEstimateDiff.getResult();
        if (!LangUtils.isFinite(convergenceValue)) {
          break;
        }
        if (!(randomY && iterationNumber == 1) && convergenceValue < estimateErrorConvergenceThreshold) {
          break;
        }
      }
      copyOut(h, userIDs, itemIDs);
    } finally {
      nDestroy(h);
    }
    return null;
  }

  private void check(long h, int status) throws ExecutionException {
    if (status == ALS_OK) {
      return;
    }
    if (status == ALS_E_SINGULAR) {
      // unchecked, unwrapped: DelegateGenerationManager.java:345-354 lowers model.features and retries
      throw new SingularMatrixSolverException(nSingularRank(h), nLastError(h));
    }
    throw new ExecutionException(new IllegalStateException(nLastError(h)));
  }

  private void copyOut(long h, long[] userIDs, long[] itemIDs) throws ExecutionException {
    ByteBuffer xb = directFloats((long) userIDs.length * features);
    ByteBuffer yb = directFloats((long) itemIDs.length * features);
    check(h, nGetX(h, xb));
    check(h, nGetY(h, yb));
    X = unflatten(xb.asFloatBuffer(), userIDs);   // ordinary float[] rows: Generation mutates them in place
    Y = unflatten(yb.asFloatBuffer(), itemIDs);
  }

  private FastByIDMap<float[]> unflatten(FloatBuffer buf, long[] ids) {
    FastByIDMap<float[]> result = new FastByIDMap<float[]>(ids.length);
    for (long id : ids) {
      float[] row = new float[features];
      buf.get(row);
      result.put(id, row);
    }
    return result;
  }

  private static ByteBuffer directFloats(long n) {
    return ByteBuffer.allocateDirect((int) (n * 4)).order(ByteOrder.nativeOrder());
  }

  private static double alpha() {
    String p = System.getProperty("model.als.alpha");
    return p == null ? AlternatingLeastSquares.DEFAULT_ALPHA : LangUtils.parseDouble(p);
  }

  private static double lambda() {
    String p = System.getProperty("model.als.lambda");
    return p == null ? AlternatingLeastSquares.DEFAULT_LAMBDA : LangUtils.parseDouble(p);
  }

  // keys(), keysOfFactors(), index(), flatten(), sample(): straightforward walks of the
  // FastByIDMap / FastByIDFloatMap entry sets into direct LongBuffer/IntBuffer/FloatBuffer
  // (row_ptr int64, col_idx int32, val fp32), elided here for brevity -- the Python mirror
  // myrrix-recommender_b200/factorizer.py::_flatten is the executable specification.
  private static long[] keys(FastByIDMap<?> m) { throw new UnsupportedOperationException("see INTEGRATION.md"); }
  private static long[] keysOfFactors(FastByIDMap<float[]> m) { throw new UnsupportedOperationException(); }
  private static FastByIDMap<Integer> index(long[] ids) { throw new UnsupportedOperationException(); }
  private static ByteBuffer[] flatten(FastByIDMap<FastByIDFloatMap> R, long[] rowIDs, FastByIDMap<Integer> colIndex) {
    throw new UnsupportedOperationException();
  }
  private static int[] sample(long[] ids, FastByIDMap<Integer> index) { throw new UnsupportedOperationException(); }
}
