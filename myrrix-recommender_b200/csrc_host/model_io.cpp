// model_io.cpp -- model.bin codec (include/myrrix_model_io.h): the Java Object Serialization
// framing around GenerationSerializer's writeObject payload, written and parsed by hand.
#include "../../include/myrrix_model_io.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <string>
#include <vector>

namespace {

const char kClassName[] = "net.myrrix.online.generation.GenerationSerializer";
const char kFieldName[] = "generation";
const char kFieldType[] = "Lnet/myrrix/online/generation/Generation;";
enum : uint8_t {
  TC_NULL = 0x70, TC_REFERENCE = 0x71, TC_CLASSDESC = 0x72, TC_OBJECT = 0x73, TC_STRING = 0x74,
  TC_BLOCKDATA = 0x77, TC_ENDBLOCKDATA = 0x78, TC_BLOCKDATALONG = 0x7A
};
constexpr size_t kMaxBlock = 1024;  // ObjectOutputStream.BlockDataOutputStream.MAX_BLOCK_SIZE

// ---- writer: raw bytes + block-data mode -------------------------------------------------------
struct Writer {
  std::vector<uint8_t> out;
  uint8_t blk[kMaxBlock];
  size_t pos = 0;
  void raw(const void* p, size_t n) { const uint8_t* b = (const uint8_t*)p; out.insert(out.end(), b, b + n); }
  void raw8(uint8_t v) { out.push_back(v); }
  void raw16(uint16_t v) { raw8((uint8_t)(v >> 8)); raw8((uint8_t)v); }
  void utf(const char* s) { raw16((uint16_t)strlen(s)); raw(s, strlen(s)); }
  void drain() {
    if (pos == 0) return;
    if (pos <= 0xff) { raw8(TC_BLOCKDATA); raw8((uint8_t)pos); }
    else { raw8(TC_BLOCKDATALONG); raw8(0); raw8(0); raw8((uint8_t)(pos >> 8)); raw8((uint8_t)pos); }
    raw(blk, pos);
    pos = 0;
  }
  void byte(uint8_t v) { if (pos >= kMaxBlock) drain(); blk[pos++] = v; }
  void be(uint64_t v, int n) { for (int i = n - 1; i >= 0; i--) byte((uint8_t)(v >> (8 * i))); }
  void i32(int32_t v) { be((uint32_t)v, 4); }
  void i64(int64_t v) { be((uint64_t)v, 8); }
  void f32(float v) { uint32_t u; memcpy(&u, &v, 4); be(u, 4); }
};

// ---- reader: raw cursor, then a cursor over the concatenated block data -------------------------
struct Cursor {
  const uint8_t* p;
  size_t n, at = 0;
  bool ok = true;
  uint8_t u8() { if (at + 1 > n) { ok = false; return 0; } return p[at++]; }
  uint64_t be(int k) { uint64_t v = 0; for (int i = 0; i < k; i++) v = (v << 8) | u8(); return v; }
  bool skip(size_t k) { if (at + k > n) { ok = false; return false; } at += k; return true; }
  bool utf(std::string* s) {
    const size_t len = (size_t)be(2);
    if (!ok || at + len > n) { ok = false; return false; }
    if (s) s->assign((const char*)p + at, len);
    at += len;
    return true;
  }
};

}  // namespace

struct model_io_reader {
  int32_t features = 0;
  std::vector<int64_t> ids[2];
  std::vector<float> m[2];
  bool has_known = false;
  std::vector<int64_t> known_users, known_ptr, known_items, tags[2];
  int64_t clusters[2] = {0, 0};
};

extern "C" {

int model_io_write(const model_io_desc* d, uint8_t** out, size_t* out_len) {
  if (!d || !out || !out_len || d->features < 0 || d->n_users < 0 || d->n_items < 0) return MODEL_IO_E_ARG;
  try {
    Writer w;
    w.raw16(0xACED); w.raw16(5);        // STREAM_MAGIC, STREAM_VERSION
    w.raw8(TC_OBJECT);
    w.raw8(TC_CLASSDESC);
    w.utf(kClassName);
    for (int i = 0; i < 7; i++) w.raw8(0);
    w.raw8(1);                          // serialVersionUID = 1L (GenerationSerializer.java:51)
    w.raw8(0x03);                       // SC_WRITE_METHOD | SC_SERIALIZABLE
    w.raw16(1);                         // one declared field: `private Generation generation`
    w.raw8('L'); w.utf(kFieldName); w.raw8(TC_STRING); w.utf(kFieldType);
    w.raw8(TC_ENDBLOCKDATA);            // no class annotation
    w.raw8(TC_NULL);                    // no serialisable superclass
    // ---- writeObject payload (GenerationSerializer.java:96-106) in block-data mode ----
    if (!d->has_known) {
      w.i32(-1);                        // NULL_COUNT (:155-156)
    } else {
      w.i32((int32_t)d->n_known_users);
      for (int64_t u = 0; u < d->n_known_users; u++) {
        w.i64(d->known_user_ids[u]);
        const int64_t a = d->known_ptr[u], b = d->known_ptr[u + 1];
        w.i32((int32_t)(b - a));
        for (int64_t e = a; e < b; e++) w.i64(d->known_item_ids[e]);
      }
    }
    const int64_t n[2] = {d->n_users, d->n_items};
    const int64_t* ids[2] = {d->user_ids, d->item_ids};
    const float* mat[2] = {d->x, d->y};
    for (int which = 0; which < 2; which++) {  // writeMatrix (:195-211)
      w.i32((int32_t)n[which]);
      for (int64_t r = 0; r < n[which]; r++) {
        w.i64(ids[which][r]);
        w.i32(d->features);
        for (int f = 0; f < d->features; f++) {
          const float v = mat[which][(size_t)r * d->features + f];
          if (!isfinite(v)) return MODEL_IO_E_NONFINITE;
          w.f32(v);
        }
      }
    }
    w.i32((int32_t)d->n_item_tags);
    for (int64_t i = 0; i < d->n_item_tags; i++) w.i64(d->item_tags[i]);
    w.i32((int32_t)d->n_user_tags);
    for (int64_t i = 0; i < d->n_user_tags; i++) w.i64(d->user_tags[i]);
    w.i32(0);                           // user clusters
    w.i32(0);                           // item clusters
    w.drain();
    w.raw8(TC_ENDBLOCKDATA);
    uint8_t* buf = (uint8_t*)malloc(w.out.size() ? w.out.size() : 1);
    if (!buf) return MODEL_IO_E_OOM;
    memcpy(buf, w.out.data(), w.out.size());
    *out = buf;
    *out_len = w.out.size();
  } catch (const std::bad_alloc&) {
    return MODEL_IO_E_OOM;
  }
  return MODEL_IO_OK;
}

void model_io_free(void* p) { free(p); }

int model_io_read(const uint8_t* bytes, size_t len, model_io_reader** out) {
  if (!bytes || !out) return MODEL_IO_E_ARG;
  try {
    Cursor c{bytes, len};
    if (c.be(2) != 0xACED || c.be(2) != 5 || c.u8() != TC_OBJECT || c.u8() != TC_CLASSDESC) return MODEL_IO_E_FORMAT;
    std::string name;
    if (!c.utf(&name) || name != kClassName) return MODEL_IO_E_FORMAT;
    c.skip(8);  // serialVersionUID
    const uint8_t flags = c.u8();
    if (!c.ok || !(flags & 0x01)) return MODEL_IO_E_FORMAT;  // the payload exists only with SC_WRITE_METHOD
    const int n_fields = (int)c.be(2);
    for (int f = 0; f < n_fields && c.ok; f++) {
      const uint8_t type = c.u8();
      c.utf(nullptr);
      if (type == 'L' || type == '[') {  // className1: a string object or a back reference
        const uint8_t tc = c.u8();
        if (tc == TC_STRING) c.utf(nullptr);
        else if (tc == TC_REFERENCE) c.skip(4);
        else return MODEL_IO_E_FORMAT;
      }
    }
    if (!c.ok || c.u8() != TC_ENDBLOCKDATA || c.u8() != TC_NULL) return MODEL_IO_E_FORMAT;
    if ((flags & 0x02) == 0) return MODEL_IO_E_FORMAT;
    // default field values would come here only if writeObject had called defaultWriteObject
    // gather the block-data records up to TC_ENDBLOCKDATA
    std::vector<uint8_t> data;
    for (;;) {
      const uint8_t tc = c.u8();
      if (!c.ok) return MODEL_IO_E_FORMAT;
      if (tc == TC_ENDBLOCKDATA) break;
      size_t n;
      if (tc == TC_BLOCKDATA) n = c.u8();
      else if (tc == TC_BLOCKDATALONG) n = (size_t)c.be(4);
      else return MODEL_IO_E_FORMAT;
      if (!c.ok || c.at + n > c.n) return MODEL_IO_E_FORMAT;
      data.insert(data.end(), c.p + c.at, c.p + c.at + n);
      c.at += n;
    }
    Cursor p{data.data(), data.size()};
    model_io_reader* r = new model_io_reader();
    auto fail = [&](int rc) { delete r; return rc; };
    const int32_t known = (int32_t)p.be(4);  // readKnownIDs (:133-150)
    r->has_known = known != -1;
    r->known_ptr.push_back(0);
    for (int32_t u = 0; r->has_known && u < known && p.ok; u++) {
      r->known_users.push_back((int64_t)p.be(8));
      const int32_t cnt = (int32_t)p.be(4);
      for (int32_t e = 0; e < cnt && p.ok; e++) r->known_items.push_back((int64_t)p.be(8));
      r->known_ptr.push_back((int64_t)r->known_items.size());
    }
    r->features = -1;
    for (int which = 0; which < 2; which++) {  // readMatrix (:174-190)
      const int32_t rows = (int32_t)p.be(4);
      for (int32_t i = 0; i < rows && p.ok; i++) {
        r->ids[which].push_back((int64_t)p.be(8));
        const int32_t k = (int32_t)p.be(4);
        if (r->features < 0) r->features = k;
        if (k != r->features) return fail(MODEL_IO_E_RAGGED);
        for (int32_t f = 0; f < k && p.ok; f++) {
          const uint32_t bits = (uint32_t)p.be(4);
          float v;
          memcpy(&v, &bits, 4);
          if (!isfinite(v)) return fail(MODEL_IO_E_NONFINITE);
          r->m[which].push_back(v);
        }
      }
    }
    if (r->features < 0) r->features = 0;
    for (int which = 0; which < 2; which++) {  // readIDSet (:213-221)
      const int32_t cnt = (int32_t)p.be(4);
      for (int32_t i = 0; i < cnt && p.ok; i++) r->tags[which].push_back((int64_t)p.be(8));
    }
    for (int which = 0; which < 2; which++) {  // readClusters (:236-255): counted, contents skipped
      const int32_t cnt = (int32_t)p.be(4);
      r->clusters[which] = cnt;
      for (int32_t i = 0; i < cnt && p.ok; i++) {
        const int32_t members = (int32_t)p.be(4);
        p.skip((size_t)members * 8);
        const int32_t centroid = (int32_t)p.be(4);
        p.skip((size_t)centroid * 4);
      }
    }
    if (!p.ok || p.at != p.n) return fail(MODEL_IO_E_FORMAT);
    *out = r;
  } catch (const std::bad_alloc&) {
    return MODEL_IO_E_OOM;
  }
  return MODEL_IO_OK;
}

void model_io_reader_destroy(model_io_reader* r) { delete r; }

int64_t model_io_count(const model_io_reader* r, int what) {
  if (!r) return -1;
  switch (what) {
    case 0: return r->features;
    case 1: return (int64_t)r->ids[0].size();
    case 2: return (int64_t)r->ids[1].size();
    case 3: return r->has_known ? 1 : 0;
    case 4: return (int64_t)r->known_users.size();
    case 5: return (int64_t)r->known_items.size();
    case 6: return (int64_t)r->tags[0].size();
    case 7: return (int64_t)r->tags[1].size();
    case 8: return r->clusters[0];
    case 9: return r->clusters[1];
    default: return -1;
  }
}

int model_io_get_matrix(const model_io_reader* r, int which, int64_t* ids, float* m) {
  if (!r || (which != 0 && which != 1)) return MODEL_IO_E_ARG;
  if (!r->ids[which].empty()) {
    if (!ids || (!m && !r->m[which].empty())) return MODEL_IO_E_ARG;
    memcpy(ids, r->ids[which].data(), r->ids[which].size() * sizeof(int64_t));
    if (!r->m[which].empty()) memcpy(m, r->m[which].data(), r->m[which].size() * sizeof(float));
  }
  return MODEL_IO_OK;
}

int model_io_get_known(const model_io_reader* r, int64_t* user_ids, int64_t* ptr, int64_t* item_ids) {
  if (!r || !ptr) return MODEL_IO_E_ARG;
  memcpy(ptr, r->known_ptr.data(), r->known_ptr.size() * sizeof(int64_t));
  if (!r->known_users.empty()) {
    if (!user_ids) return MODEL_IO_E_ARG;
    memcpy(user_ids, r->known_users.data(), r->known_users.size() * sizeof(int64_t));
  }
  if (!r->known_items.empty()) {
    if (!item_ids) return MODEL_IO_E_ARG;
    memcpy(item_ids, r->known_items.data(), r->known_items.size() * sizeof(int64_t));
  }
  return MODEL_IO_OK;
}

int model_io_get_tags(const model_io_reader* r, int which, int64_t* out) {
  if (!r || (which != 0 && which != 1)) return MODEL_IO_E_ARG;
  if (!r->tags[which].empty()) {
    if (!out) return MODEL_IO_E_ARG;
    memcpy(out, r->tags[which].data(), r->tags[which].size() * sizeof(int64_t));
  }
  return MODEL_IO_OK;
}

}  // extern "C"
