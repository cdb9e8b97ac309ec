/*
 * The Java side of include/myrrix_ingest.h: what a Myrrix maintainer adds next to
 * InputFilesReader (online-local/src/net/myrrix/online/generation/InputFilesReader.java) so that
 * DelegateGenerationManager.java:333 can hand the parsed matrix straight to
 * CudaAlternatingLeastSquares without building FastByIDMap<FastByIDFloatMap> first.
 * Source only: this image has no JDK.
 */
package net.myrrix.online.generation;

import java.io.File;
import java.io.FilenameFilter;
import java.io.IOException;
import java.io.InputStream;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.util.Arrays;

import com.google.common.io.ByteStreams;
import com.google.common.io.PatternFilenameFilter;

import net.myrrix.common.io.ByLastModifiedComparator;

final class NativeInputFilesReader {

  static { System.loadLibrary("myrrix_ingest_jni"); }

  /** Dense CSR by user plus the long IDs behind the dense indices. */
  static final class Interactions {
    ByteBuffer userIDs;  // int64 [nUsers]
    ByteBuffer itemIDs;  // int64 [nItems]
    ByteBuffer rowPtr;   // int64 [nUsers + 1]
    ByteBuffer colIdx;   // int32 [nnz]
    ByteBuffer val;      // fp32  [nnz]
    long nUsers, nItems, nnz;
  }

  private NativeInputFilesReader() {}

  /** Same files, same order and same line semantics as InputFilesReader.readInputFiles (:64-196). */
  static Interactions readInputFiles(File inputDir) throws IOException {
    FilenameFilter csvFilter = new PatternFilenameFilter(".+\\.csv(\\.(zip|gz))?");
    File[] inputFiles = inputDir.listFiles(csvFilter);
    long h = nCreate(Float.parseFloat(System.getProperty("model.decay.zeroThreshold", "0.0001")));
    try {
      if (inputFiles != null) {
        Arrays.sort(inputFiles, ByLastModifiedComparator.INSTANCE);
        for (File f : inputFiles) {
          // decompression stays in Java (FileLineIterator.getFileInputStream)
          InputStream in = net.myrrix.common.iterator.FileLineIterator.getFileInputStream(f);
          byte[] bytes;
          try { bytes = ByteStreams.toByteArray(in); } finally { in.close(); }
          ByteBuffer direct = ByteBuffer.allocateDirect(bytes.length);
          direct.put(bytes);
          int rc = nAddFile(h, direct, bytes.length);
          if (rc == 2) throw new IOException("Too many bad lines; aborting");
          if (rc != 0) throw new IOException(nLastError(h));
        }
      }
      if (nFinish(h) != 0) throw new IOException(nLastError(h));
      Interactions r = new Interactions();
      r.nUsers = nCount(h, 0); r.nItems = nCount(h, 1); r.nnz = nCount(h, 2);
      r.userIDs = direct(8 * r.nUsers); r.itemIDs = direct(8 * r.nItems);
      r.rowPtr = direct(8 * (r.nUsers + 1)); r.colIdx = direct(4 * r.nnz); r.val = direct(4 * r.nnz);
      nGetIds(h, 0, r.userIDs); nGetIds(h, 1, r.itemIDs);
      nGetCsr(h, r.rowPtr, r.colIdx, r.val);
      return r;
    } finally {
      nDestroy(h);
    }
  }

  private static ByteBuffer direct(long bytes) {
    return ByteBuffer.allocateDirect((int) Math.max(bytes, 1)).order(ByteOrder.nativeOrder());
  }

  private static native long nCreate(float zeroThreshold);
  private static native void nDestroy(long handle);
  private static native int nAddFile(long handle, ByteBuffer bytes, long len);
  private static native int nFinish(long handle);
  private static native long nCount(long handle, int kind);
  private static native int nGetIds(long handle, int which, ByteBuffer out);
  private static native int nGetCsr(long handle, ByteBuffer rowPtr, ByteBuffer colIdx, ByteBuffer val);
  private static native int nGetKnown(long handle, ByteBuffer rowPtr, ByteBuffer colIdx);
  private static native int nGetTags(long handle, int which, ByteBuffer out);
  private static native String nLastError(long handle);
}
