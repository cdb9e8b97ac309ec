/* myrrix_ingest.h -- C ABI of the input canonicalisation that feeds the ALS core
 * (libmyrrix_ingest.so; plain C++, no CUDA, loads on any host).
 *
 * Replaces, for the path CSV files -> interaction matrix, the reference's
 *   InputFilesReader.readInputFiles   online-local/src/net/myrrix/online/generation/InputFilesReader.java:64-196
 *   InputFilesReader.removeSmall      ...InputFilesReader.java:198-211
 *   MatrixUtils.addTo / remove        common/src/net/myrrix/common/math/MatrixUtils.java:64-121
 *   FastByIDFloatMap.increment        common/src/net/myrrix/common/collection/FastByIDFloatMap.java:129-138
 *   OneWayMigrator.toLongID           common/src/net/myrrix/common/OneWayMigrator.java:27 (Mahout AbstractIDMigrator:
 *                                     first 8 bytes of MD5(UTF-8), big-endian)
 * and hands back what als_set_interactions (myrrix_als.h) takes: a CSR with dense 0-based
 * indices, duplicates summed in fp32 in input order, deletions applied, |v| < threshold pruned,
 * plus the long IDs behind the dense indices (the FastByIDMap keys of RbyRow / RbyColumn).
 *
 * Line semantics kept from the reference (InputFilesReader.java:93-150):
 *   - empty lines and lines starting with '#' are skipped;
 *   - fields are split on ',' and trimmed; "user,item" means strength 1.0; "user,item," (empty
 *     third field) deletes the entry; further fields are ignored;
 *   - a quoted first/second field is a tag, hashed to a long; two tags on a line are rejected;
 *   - a line that does not parse is "bad" (the very first line of the input is forgiven as a
 *     header); the 102nd bad line aborts (INGEST_E_BAD_LINES);
 *   - non-finite strengths are bad lines.
 * A user (item) whose entries were all deleted disappears; one whose entries were all pruned
 * as near-zero stays, with an empty row (the reference keeps the map key: removeSmall only
 * empties the inner maps).
 *
 * Dense indices are assigned in order of first appearance among the IDs that survive.
 * Not thread-safe per handle.
 */
#ifndef MYRRIX_INGEST_H
#define MYRRIX_INGEST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ingest_handle ingest_handle;

enum ingest_status {
  INGEST_OK = 0,
  INGEST_E_ARG = 1,
  INGEST_E_BAD_LINES = 2, /* "Too many bad lines; aborting" (InputFilesReader.java:95-97) */
  INGEST_E_STATE = 3,     /* wrong call order */
  INGEST_E_OOM = 4,
  INGEST_E_RANGE = 5      /* more than 2^31-1 distinct users or items */
};

enum ingest_count_kind {
  INGEST_N_USERS = 0,      /* keys of RbyRow */
  INGEST_N_ITEMS = 1,      /* keys of RbyColumn */
  INGEST_NNZ = 2,          /* entries after pruning */
  INGEST_KNOWN_NNZ = 3,    /* entries before pruning (knownItemIDs) */
  INGEST_LINES = 4,
  INGEST_BAD_LINES = 5,
  INGEST_N_ITEM_TAGS = 6,  /* itemTagIDs: users given as tags (InputFilesReader.java:152-154) */
  INGEST_N_USER_TAGS = 7   /* userTagIDs: items given as tags (:156-158) */
};

/* zero_threshold: model.decay.zeroThreshold, 0.0001 in the reference (InputFilesReader.java:58-59). */
int ingest_create(float zero_threshold, ingest_handle** out);
void ingest_destroy(ingest_handle* h);
/* Files are cut at line boundaries and parsed on up to max_threads threads (0 = all host
 * threads) in chunks of at least min_chunk_bytes (default 1 MiB). The result does not depend
 * on either value. */
int ingest_set_parallelism(ingest_handle* h, int max_threads, size_t min_chunk_bytes);

/* The decompressed bytes of one input file; call once per file, files in last-modified order
 * (InputFilesReader.java:86). Lines end with \n, \r or \r\n. */
int ingest_add_file(ingest_handle* h, const char* data, size_t len);

/* Apply additions / deletions in input order, prune, build the dense maps and the CSR. */
int ingest_finish(ingest_handle* h);

int64_t ingest_count(const ingest_handle* h, int kind);

/* which: 0 = user IDs [n_users], 1 = item IDs [n_items] (dense index -> long ID). */
int ingest_get_ids(const ingest_handle* h, int which, int64_t* out);
/* CSR by user of the pruned matrix: row_ptr [n_users+1], col_idx / val [nnz]; columns ascending. */
int ingest_get_csr(const ingest_handle* h, int64_t* row_ptr, int32_t* col_idx, float* val);
/* knownItemIDs as a CSR pattern (entries before pruning): row_ptr [n_users+1], col_idx [known_nnz]. */
int ingest_get_known(const ingest_handle* h, int64_t* row_ptr, int32_t* col_idx);
/* which: 0 = itemTagIDs, 1 = userTagIDs (long IDs, order of first appearance). */
int ingest_get_tags(const ingest_handle* h, int which, int64_t* out);
/* Hash of a tag string (the long ID the reference derives for quoted fields). */
int64_t ingest_tag_id(const char* utf8, size_t len);

const char* ingest_last_error(const ingest_handle* h);

#ifdef __cplusplus
}
#endif
#endif
