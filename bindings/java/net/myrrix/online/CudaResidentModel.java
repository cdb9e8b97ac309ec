/*
 * The reference-side binding of the resident-model entry points of include/myrrix_als.h
 * (als_recommend, als_recommend_batch, als_top_n, als_set_fold_in_state, als_fold_in) -- what a
 * Myrrix maintainer would add next to ServerRecommender to keep a built model on the GPU between
 * builds.  NOT compiled here (no JDK in this image); natives in bindings/jni/myrrix_model_jni.c.
 *
 * It replaces, for dense row indices, the work ServerRecommender does AFTER its ID lookups and locks:
 *   multithreadedTopN            online/src/net/myrrix/online/ServerRecommender.java:443-509
 *     -> RecommendIterator.next  online/src/net/myrrix/online/RecommendIterator.java:68-110
 *     -> TopN                    common/src/net/myrrix/common/TopN.java:55-131
 *   updateFeatures               online/src/net/myrrix/online/ServerRecommender.java:865-907
 * The ID <-> dense index maps are the ones CudaAlternatingLeastSquares built for the factorization
 * (slot-order walk of the FastByIDMap keys); IDRescorer callbacks stay in Java: isFiltered() is
 * evaluated on the caller's side into the `exclude` list, rescore() is not supported on the device
 * (such calls keep the reference path).
 */
package net.myrrix.online;

import java.util.ArrayList;
import java.util.List;

import org.apache.mahout.cf.taste.recommender.RecommendedItem;

import net.myrrix.common.MutableRecommendedItem;
import net.myrrix.common.collection.FastByIDMap;

public final class CudaResidentModel {

  static {
    System.loadLibrary("myrrix_model_jni");
  }

  private static final int ALS_OK = 0;
  private static final int ALS_E_NONFINITE = 2;

  private final long handle;                    // the als_handle the factorization left alive
  private final long[] itemIDs;                 // dense item row -> ID
  private final FastByIDMap<Integer> userIndex; // ID -> dense user row
  private final FastByIDMap<Integer> itemIndex;

  private static native int nRecommend(long handle, int[] users, int howMany, boolean considerKnownItems,
                                       int[] exclude, int[] outItems, float[] outValues, int[] outCount);
  private static native int nRecommendBatch(long handle, int[] users, int howMany, boolean considerKnownItems,
                                            int[] outItems, float[] outValues, int[] outCounts);
  private static native int nTopN(long handle, int which, float[] features, int nVectors, int[] exclude,
                                  int howMany, int[] outIDs, float[] outValues, int[] outCount);
  private static native int nSetFoldInState(long handle, int which, double[] qrt, double[] rdiag, int[] perm,
                                            double learnRate);
  private static native int nFoldIn(long handle, int[] users, int[] items, float[] values);
  private static native String nLastError(long handle);

  public CudaResidentModel(long handle, long[] itemIDs, FastByIDMap<Integer> userIndex,
                           FastByIDMap<Integer> itemIndex) {
    this.handle = handle;
    this.itemIDs = itemIDs;
    this.userIndex = userIndex;
    this.itemIndex = itemIndex;
  }

  /** recommendToMany after the ID lookups (ServerRecommender.java:366-441). */
  public List<RecommendedItem> recommend(long[] userIDs, int howMany, boolean considerKnownItems,
                                         long[] filteredItemIDs) {
    int[] users = dense(userIDs, userIndex);
    if (users.length == 0) {
      throw new IllegalArgumentException("no such user");   // NoSuchUserException in the caller (:391-393)
    }
    int[] outItems = new int[howMany];
    float[] outValues = new float[howMany];
    int[] outCount = new int[1];
    check(nRecommend(handle, users, howMany, considerKnownItems, dense(filteredItemIDs, itemIndex),
                     outItems, outValues, outCount));
    List<RecommendedItem> result = new ArrayList<RecommendedItem>(outCount[0]);
    for (int i = 0; i < outCount[0]; i++) {
      result.add(new MutableRecommendedItem(itemIDs[outItems[i]], outValues[i]));
    }
    return result;   // value descending, equal values by item ascending: TopN.selectTopNFromQueue's order
  }

  /** updateFeatures for one write (setPreference on a known user and item); returns false for new IDs. */
  public boolean foldIn(long userID, long itemID, float value) {
    Integer u = userIndex.get(userID);
    Integer i = itemIndex.get(itemID);
    if (u == null || i == null) {
      return false;   // new IDs wait for the next build, as their vectors do in the reference
    }
    check(nFoldIn(handle, new int[] {u}, new int[] {i}, new float[] {value}));
    return true;
  }

  /** Generation.recomputeState: the factors of the generation's solvers (myrrix_foldin.h), per side. */
  public void setFoldInState(int which, double[] qrt, double[] rdiag, int[] perm, double learnRate) {
    check(nSetFoldInState(handle, which, qrt, rdiag, perm, learnRate));
  }

  private static int[] dense(long[] ids, FastByIDMap<Integer> index) {
    if (ids == null) {
      return new int[0];
    }
    int[] tmp = new int[ids.length];
    int n = 0;
    for (long id : ids) {
      Integer row = index.get(id);
      if (row != null) {
        tmp[n++] = row;
      }
    }
    return java.util.Arrays.copyOf(tmp, n);
  }

  private void check(int status) {
    if (status == ALS_OK) {
      return;
    }
    if (status == ALS_E_NONFINITE) {
      throw new IllegalStateException("Bad recommendation value");   // RecommendIterator.java:99
    }
    throw new IllegalStateException(nLastError(handle));
  }

}
