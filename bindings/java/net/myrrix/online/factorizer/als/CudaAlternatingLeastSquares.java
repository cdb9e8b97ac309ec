/*
 * Reference-side binding a Myrrix maintainer would add (NOT compiled in this repo's image:
 * there is no JDK here; see INTEGRATION.md).  Implements the reference's plug-in interface
 *   net.myrrix.online.factorizer.MatrixFactorizer  (online/.../factorizer/MatrixFactorizer.java:31-77)
 * over the C ABI in include/myrrix_als.h via the JNI stub in bindings/jni/myrrix_als_jni.c.
 * The only production change is DelegateGenerationManager.java:406-410, which constructs this
 * class instead of AlternatingLeastSquares (same constructor arguments).
 *
 * Everything the C ABI needs is built here, in the JVM: dense indices for the long IDs, both CSR
 * orientations, the initial Y (same rules as AlternatingLeastSquares.constructInitialY, :264-335),
 * the convergence sample and the stop rule (:206-257).  Large arrays live in native memory
 * allocated by the stub and are filled / read through direct ByteBuffer windows of at most
 * 1 GiB, so nothing is limited by the 2 GiB size of a single ByteBuffer (the headline X is 2.56 GB).
 */
package net.myrrix.online.factorizer.als;

import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.util.ArrayList;
import java.util.List;
import java.util.concurrent.ExecutionException;

import org.apache.commons.math3.random.RandomGenerator;
import org.apache.commons.math3.util.FastMath;
import org.apache.mahout.cf.taste.impl.common.LongPrimitiveIterator;

import net.myrrix.common.LangUtils;
import net.myrrix.common.collection.FastByIDFloatMap;
import net.myrrix.common.collection.FastByIDMap;
import net.myrrix.common.math.SimpleVectorMath;
import net.myrrix.common.math.SingularMatrixSolverException;
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
  private static final int ALS_E_NONFINITE = 2;

  private static final int NUM_USER_ITEMS_TO_TEST_CONVERGENCE = 100;  // ALS.java:78
  private static final int MAX_FAR_FROM_VECTORS = 100000;             // ALS.java:80
  private static final long WINDOW_BYTES = 1L << 30;

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
                                             long rowPtrAddr, long colIdxAddr, long valAddr);
  private static native int nSetInteractionsByColumn(long handle, long colPtrAddr, long rowIdxAddr, long valAddr);
  private static native int nSetPresentEmptyRows(long handle, int which, int[] rows);
  private static native int nSetY(long handle, long yAddr);
  private static native int nHalfX(long handle);
  private static native int nHalfY(long handle);
  private static native int nSync(long handle);
  private static native int nProbe(long handle, int[] users, int[] items, double[] out);
  private static native int nCall(long handle, int[] testUsers, int[] testItems, int maxIterations,
                                  double convergenceThreshold, boolean randomY, boolean xIsEmpty,
                                  int[] iterationsRun, double[] lastConvergenceValue);
  private static native int nGetX(long handle, long outAddr);
  private static native int nGetY(long handle, long outAddr);
  private static native String nLastError(long handle);
  private static native int nSingularRank(long handle);
  // native memory: malloc / free, and a direct ByteBuffer window [offset, offset + length) onto it
  private static native long nAlloc(long bytes);
  private static native void nFree(long address);
  private static native ByteBuffer nWindow(long address, long offset, int length);

  /** A native array addressed in windows of at most 1 GiB (little-endian like the device). */
  private static final class NativeArray {
    final long address;
    final long bytes;
    private ByteBuffer window;
    private long windowStart = -1;
    NativeArray(long bytes) {
      this.bytes = bytes;
      this.address = nAlloc(Math.max(bytes, 8L));
      if (address == 0L) {
        throw new OutOfMemoryError("native allocation of " + bytes + " bytes failed");
      }
    }
    private ByteBuffer at(long offset) {  // window containing `offset`; elements never straddle (8 | 1 GiB)
      long start = offset - (offset % WINDOW_BYTES);
      if (start != windowStart) {
        int length = (int) Math.min(WINDOW_BYTES, bytes - start);
        window = nWindow(address, start, length).order(ByteOrder.nativeOrder());
        windowStart = start;
      }
      return window;
    }
    void putLong(long index, long v) { at(index * 8).putLong((int) ((index * 8) % WINDOW_BYTES), v); }
    void putInt(long index, int v) { at(index * 4).putInt((int) ((index * 4) % WINDOW_BYTES), v); }
    void putFloat(long index, float v) { at(index * 4).putFloat((int) ((index * 4) % WINDOW_BYTES), v); }
    float getFloat(long index) { return at(index * 4).getFloat((int) ((index * 4) % WINDOW_BYTES)); }
    void free() { nFree(address); }
  }

  /** CSR of one orientation in native memory + the rows that are keys without entries. */
  private static final class Csr {
    NativeArray ptr, idx, val;
    int[] presentEmptyRows;
    void free() {
      if (ptr != null) { ptr.free(); }
      if (idx != null) { idx.free(); }
      if (val != null) { val.free(); }
    }
  }

  @Override
  public Void call() throws ExecutionException, InterruptedException {
    RandomGenerator random = RandomManager.getRandom();
    boolean randomY = previousY == null || previousY.isEmpty();                 // ALS.java:181
    FastByIDMap<float[]> initialY = constructInitialY(previousY, random);        // ALS.java:182

    // long ID -> dense index in the maps' own iteration order (FastByIDMap.java:499-533).  Every
    // row of Y counts in Y^T Y, including stale rows that are not keys of RbyColumn.
    long[] userIDs = keys(RbyRow);
    long[] itemIDs = keys(initialY);
    FastByIDMap<Integer> userIndex = index(userIDs);
    FastByIDMap<Integer> itemIndex = index(itemIDs);

    Csr byRow = null;
    Csr byCol = null;
    NativeArray y0 = null;
    long h = 0L;
    try {
      byRow = flatten(RbyRow, userIDs, itemIndex);
      byCol = flatten(RbyColumn, itemIDs, userIndex);  // items that are not keys of RbyColumn: empty, absent
      y0 = new NativeArray((long) itemIDs.length * features * 4L);
      long o = 0;
      for (long itemID : itemIDs) {
        for (float v : initialY.get(itemID)) {
          y0.putFloat(o++, v);
        }
      }
      h = nCreate(features, alpha(), lambda(),
                  Boolean.parseBoolean(System.getProperty("model.reconstructRMatrix", "false")),
                  Boolean.parseBoolean(System.getProperty("model.lossIgnoresUnspecified", "false")),
                  Double.parseDouble(System.getProperty("common.matrix.singularityThreshold", "1.0e-5")),
                  Integer.getInteger("model.cuda.device", 0));
      check(h, nSetInteractions(h, userIDs.length, itemIDs.length,
                                byRow.ptr.address, byRow.idx.address, byRow.val.address));
      check(h, nSetInteractionsByColumn(h, byCol.ptr.address, byCol.idx.address, byCol.val.address));
      // keys whose maps InputFilesReader.removeSmall emptied (:202-211) are still walked by
      // addWorkers (ALS.java:391-410): solved as W = G, b = 0
      if (byRow.presentEmptyRows.length > 0) {
        check(h, nSetPresentEmptyRows(h, 0, byRow.presentEmptyRows));
      }
      if (byCol.presentEmptyRows.length > 0) {
        check(h, nSetPresentEmptyRows(h, 1, byCol.presentEmptyRows));
      }
      check(h, nSetY(h, y0.address));
      byRow.free(); byRow = null;   // the library has its own copies now
      byCol.free(); byCol = null;
      y0.free(); y0 = null;

      if (!Boolean.parseBoolean(System.getProperty("model.als.iterate", "true"))) {  // ALS.java:196-204
        check(h, nHalfX(h));
        check(h, nSync(h));
        copyOut(h, userIDs, itemIDs, initialY);
        return null;
      }

      // RandomUtils.chooseAboutNFromStream over the key sets (ALS.java:206-214)
      int[] testUsers = sample(RbyRow.keySetIterator(), RbyRow.size(), userIndex, random);
      int[] testItems = sample(RbyColumn.keySetIterator(), RbyColumn.size(), itemIndex, random);
      if (!Boolean.parseBoolean(System.getProperty("model.als.hostStopRule", "false"))) {
        // the loop below, run by the library in one call with the statistic evaluated on the device
        // (als_call); row-update errors (ALS_E_SINGULAR, ...) come back from it
        check(h, nCall(h, testUsers, testItems, maxIterations, estimateErrorConvergenceThreshold, randomY, true,
                       new int[1], new double[1]));
        copyOut(h, userIDs, itemIDs, initialY);
        return null;
      }
      double[] estimates = new double[testUsers.length * testItems.length];  // X empty: zeros (:215-223)
      double[] fresh = new double[estimates.length];
      int iterationNumber = 0;
      while (true) {
        check(h, nHalfX(h));   // iterateXFromY (:228)
        check(h, nHalfY(h));   // iterateYFromX (:229)
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
      copyOut(h, userIDs, itemIDs, initialY);
    } finally {
      if (h != 0L) { nDestroy(h); }
      if (byRow != null) { byRow.free(); }
      if (byCol != null) { byCol.free(); }
      if (y0 != null) { y0.free(); }
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
    if (status == ALS_E_NONFINITE) {
      // a SolverException that is not the singular one: "waiting for more data", generation dropped
      // (DelegateGenerationManager.java:375-378); SolverException's own constructors are protected
      throw new net.myrrix.common.math.IllConditionedSolverException(nLastError(h));
    }
    throw new ExecutionException(new IllegalStateException(nLastError(h)));
  }

  /** getX()/getY(): ordinary float[] rows, because Generation keeps and mutates them (fold-in). */
  private void copyOut(long h, long[] userIDs, long[] itemIDs, FastByIDMap<float[]> initialY)
      throws ExecutionException {
    NativeArray xb = new NativeArray((long) userIDs.length * features * 4L);
    NativeArray yb = new NativeArray((long) itemIDs.length * features * 4L);
    try {
      check(h, nGetX(h, xb.address));
      check(h, nGetY(h, yb.address));
      X = new FastByIDMap<float[]>(userIDs.length);
      long o = 0;
      for (long id : userIDs) {
        float[] row = new float[features];
        for (int f = 0; f < features; f++) {
          row[f] = xb.getFloat(o++);
        }
        X.put(id, row);
      }
      // Y keeps the identity of the map it started from: the reference adopts previousY in place
      // when the feature count is unchanged (ALS.java:304-308) and updates its rows
      o = 0;
      for (long id : itemIDs) {
        float[] row = initialY.get(id);
        for (int f = 0; f < features; f++) {
          row[f] = yb.getFloat(o++);
        }
      }
      Y = initialY;
    } finally {
      xb.free();
      yb.free();
    }
  }

  // ---- constructInitialY: the rules of AlternatingLeastSquares.java:264-335 -------------------
  private FastByIDMap<float[]> constructInitialY(FastByIDMap<float[]> previous, RandomGenerator random) {
    FastByIDMap<float[]> y;
    if (previous == null || previous.isEmpty()) {
      y = new FastByIDMap<float[]>(RbyColumn.size());
    } else {
      int oldFeatures = previous.entrySet().iterator().next().getValue().length;
      if (oldFeatures == features) {
        y = previous;  // adopted in place; the caller passes a clone (DelegateGenerationManager.java:419-426)
      } else {
        y = new FastByIDMap<float[]>(previous.size());
        for (FastByIDMap.MapEntry<float[]> entry : previous.entrySet()) {
          float[] old = entry.getValue();
          float[] resized = new float[features];
          System.arraycopy(old, 0, resized, 0, Math.min(old.length, features));
          for (int i = old.length; i < features; i++) {
            resized[i] = (float) random.nextGaussian();  // new dimensions start random
          }
          SimpleVectorMath.normalize(resized);
          y.put(entry.getKey(), resized);
        }
      }
    }
    List<float[]> recent = new ArrayList<float[]>();
    for (FastByIDMap.MapEntry<float[]> entry : y.entrySet()) {
      if (recent.size() >= MAX_FAR_FROM_VECTORS) {
        break;
      }
      recent.add(entry.getValue());
    }
    LongPrimitiveIterator it = RbyColumn.keySetIterator();
    while (it.hasNext()) {
      long id = it.nextLong();
      if (!y.containsKey(id)) {
        float[] fresh = RandomUtils.randomUnitVectorFarFrom(features, recent, random);
        y.put(id, fresh);
        if (recent.size() < MAX_FAR_FROM_VECTORS) {
          recent.add(fresh);
        }
      }
    }
    return y;
  }

  // ---- flattening ------------------------------------------------------------------------------
  private static long[] keys(FastByIDMap<?> map) {
    long[] ids = new long[map.size()];
    LongPrimitiveIterator it = map.keySetIterator();
    int i = 0;
    while (it.hasNext()) {
      ids[i++] = it.nextLong();
    }
    return ids;
  }

  private static FastByIDMap<Integer> index(long[] ids) {
    FastByIDMap<Integer> index = new FastByIDMap<Integer>(ids.length);
    for (int i = 0; i < ids.length; i++) {
      index.put(ids[i], i);
    }
    return index;
  }

  /**
   * {rowID: {colID: value}} -> CSR over rowIDs (dense order), columns through colIndex.  A rowID
   * that is not a key of R (a stale Y row) becomes an empty, absent row; a key with an empty map
   * becomes an empty row that is listed in presentEmptyRows.
   */
  private static Csr flatten(FastByIDMap<FastByIDFloatMap> R, long[] rowIDs, FastByIDMap<Integer> colIndex) {
    long nnz = 0;
    for (long rowID : rowIDs) {
      FastByIDFloatMap row = R.get(rowID);
      if (row != null) {
        nnz += row.size();
      }
    }
    Csr csr = new Csr();
    List<Integer> presentEmpty = new ArrayList<Integer>();
    try {
      csr.ptr = new NativeArray((rowIDs.length + 1L) * 8L);
      csr.idx = new NativeArray(nnz * 4L);
      csr.val = new NativeArray(nnz * 4L);
      long e = 0;
      for (int r = 0; r < rowIDs.length; r++) {
        csr.ptr.putLong(r, e);
        FastByIDFloatMap row = R.get(rowIDs[r]);
        if (row == null) {
          continue;
        }
        if (row.isEmpty()) {
          presentEmpty.add(r);
          continue;
        }
        for (FastByIDFloatMap.MapEntry entry : row.entrySet()) {
          Integer col = colIndex.get(entry.getKey());
          if (col == null) {
            throw new IllegalStateException("RbyRow and RbyColumn disagree on ID " + entry.getKey());
          }
          csr.idx.putInt(e, col);
          csr.val.putFloat(e, entry.getValue());
          e++;
        }
      }
      csr.ptr.putLong(rowIDs.length, e);
    } catch (RuntimeException ex) {
      csr.free();
      throw ex;
    }
    csr.presentEmptyRows = new int[presentEmpty.size()];
    for (int i = 0; i < csr.presentEmptyRows.length; i++) {
      csr.presentEmptyRows[i] = presentEmpty.get(i);
    }
    return csr;
  }

  private static int[] sample(LongPrimitiveIterator ids, int size, FastByIDMap<Integer> index,
                              RandomGenerator random) {
    long[] chosen = RandomUtils.chooseAboutNFromStream(NUM_USER_ITEMS_TO_TEST_CONVERGENCE, ids, size, random);
    int[] dense = new int[chosen.length];
    for (int i = 0; i < chosen.length; i++) {
      dense[i] = index.get(chosen[i]);
    }
    return dense;
  }

  private static double alpha() {
    String p = System.getProperty("model.als.alpha");
    return p == null ? AlternatingLeastSquares.DEFAULT_ALPHA : LangUtils.parseDouble(p);
  }

  private static double lambda() {
    String p = System.getProperty("model.als.lambda");
    return p == null ? AlternatingLeastSquares.DEFAULT_LAMBDA : LangUtils.parseDouble(p);
  }
}
