/* myrrix_model_io.h -- C ABI of the model file codec (libmyrrix_model_io.so; plain C++).
 *
 * Replaces the byte layout of the reference's model.bin.gz, minus the gzip layer:
 *   GenerationSerializer.writeObject / readObject   online-local/src/net/myrrix/online/generation/GenerationSerializer.java:96-131
 *       writeKnownIDs :152-169, writeMatrix :195-211 (count, then per row: long id, int length, floats),
 *       writeIDSet :223-234, writeClusters :257-276
 *   IOUtils.writeObjectToFile / readObjectFromFile   common/src/net/myrrix/common/io/IOUtils.java:259-283
 *       (ObjectOutputStream over a GZIP stream: the caller adds / removes the gzip layer)
 * The bytes are a Java Object Serialization stream (protocol version 2): stream header, one
 * object of class net.myrrix.online.generation.GenerationSerializer (serialVersionUID 1, flags
 * SC_SERIALIZABLE | SC_WRITE_METHOD, one declared field `generation` whose value is not
 * written because writeObject never calls defaultWriteObject), then the writeObject payload
 * as block-data records (written here in the 1024-byte blocks ObjectOutputStream produces; any
 * block sizes are accepted when reading), then TC_ENDBLOCKDATA.
 * PARITY UNPINNED: no JVM in the build image, no model file in the reference tree.
 * All integers inside the payload are big-endian (DataOutput).
 */
#ifndef MYRRIX_MODEL_IO_H
#define MYRRIX_MODEL_IO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum model_io_status {
  MODEL_IO_OK = 0,
  MODEL_IO_E_ARG = 1,
  MODEL_IO_E_FORMAT = 2,     /* not a GenerationSerializer stream / truncated */
  MODEL_IO_E_NONFINITE = 3,  /* Preconditions.checkState(LangUtils.isFinite(f)) (:187, :205) */
  MODEL_IO_E_RAGGED = 4,     /* rows of one matrix have different lengths */
  MODEL_IO_E_OOM = 5
};

typedef struct {
  int32_t features;
  int64_t n_users; const int64_t* user_ids; const float* x; /* [n_users][features] row-major */
  int64_t n_items; const int64_t* item_ids; const float* y; /* [n_items][features] */
  /* knownItemIDs: has_known = 0 writes the reference's "null" marker (count -1) */
  int32_t has_known;
  int64_t n_known_users; const int64_t* known_user_ids;
  const int64_t* known_ptr;      /* [n_known_users + 1] */
  const int64_t* known_item_ids; /* long item IDs, [known_ptr[n_known_users]] */
  int64_t n_item_tags; const int64_t* item_tags;
  int64_t n_user_tags; const int64_t* user_tags;
  /* clusters are written as two empty lists (the ALS path produces none) */
} model_io_desc;

/* Serialise; *out is malloc'ed (release with model_io_free). */
int model_io_write(const model_io_desc* d, uint8_t** out, size_t* out_len);
void model_io_free(void* p);

typedef struct model_io_reader model_io_reader;
/* Parse the (already gunzipped) bytes. */
int model_io_read(const uint8_t* bytes, size_t len, model_io_reader** out);
void model_io_reader_destroy(model_io_reader* r);
/* what: 0 features, 1 n_users, 2 n_items, 3 has_known, 4 n_known_users, 5 known entries,
 * 6 n_item_tags, 7 n_user_tags, 8 user clusters, 9 item clusters (counted, contents skipped) */
int64_t model_io_count(const model_io_reader* r, int what);
/* which: 0 = X, 1 = Y; ids [n], m [n][features] */
int model_io_get_matrix(const model_io_reader* r, int which, int64_t* ids, float* m);
int model_io_get_known(const model_io_reader* r, int64_t* user_ids, int64_t* ptr, int64_t* item_ids);
int model_io_get_tags(const model_io_reader* r, int which, int64_t* out);

#ifdef __cplusplus
}
#endif
#endif
