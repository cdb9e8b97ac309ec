// ingest.cpp -- input canonicalisation in front of the ALS core (include/myrrix_ingest.h).
//
// Plain C++ (no CUDA): CSV bytes -> (user, item, strength | delete) events -> interaction
// matrix as CSR with dense indices.  Follows InputFilesReader.readInputFiles
// (online-local/src/net/myrrix/online/generation/InputFilesReader.java:64-211) line for line;
// what differs is the machine mapping: the reference mutates two hash-of-hash maps per line
// (MatrixUtils.addTo / remove), here the events are appended to one array, stably sorted by
// (user, item) and folded per cell in input order -- the same fp32 sums in the same order, one
// sequential pass over 16-byte records instead of two dependent hash probes per line.
#include "../../include/myrrix_ingest.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <new>
#include <string>
#include <thread>
#include <vector>

namespace {

// ---- MD5 (RFC 1321), for OneWayMigrator.toLongID: first 8 digest bytes, big-endian ----------
struct Md5 {
  uint32_t a = 0x67452301u, b = 0xefcdab89u, c = 0x98badcfeu, d = 0x10325476u;
  static uint32_t rol(uint32_t x, int s) { return (x << s) | (x >> (32 - s)); }
  void block(const unsigned char* p) {
    static const uint32_t K[64] = {
        0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
        0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
        0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
        0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
        0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
        0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
        0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
        0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
    static const int S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22,
                              5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20,
                              4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                              6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
    uint32_t m[16];
    for (int i = 0; i < 16; i++)
      m[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) |
             ((uint32_t)p[4 * i + 3] << 24);
    uint32_t A = a, B = b, C = c, D = d;
    for (int i = 0; i < 64; i++) {
      uint32_t f;
      int g;
      if (i < 16) { f = (B & C) | (~B & D); g = i; }
      else if (i < 32) { f = (D & B) | (~D & C); g = (5 * i + 1) & 15; }
      else if (i < 48) { f = B ^ C ^ D; g = (3 * i + 5) & 15; }
      else { f = C ^ (B | ~D); g = (7 * i) & 15; }
      const uint32_t t = D;
      D = C; C = B;
      B = B + rol(A + f + K[i] + m[g], S[i]);
      A = t;
    }
    a += A; b += B; c += C; d += D;
  }
};

int64_t md5_first8_be(const char* s, size_t len) {
  Md5 h;
  size_t off = 0;
  for (; off + 64 <= len; off += 64) h.block(reinterpret_cast<const unsigned char*>(s) + off);
  unsigned char tail[128];
  const size_t rem = len - off;
  memset(tail, 0, sizeof(tail));
  memcpy(tail, s + off, rem);
  tail[rem] = 0x80;
  const size_t padded = (rem < 56) ? 64 : 128;
  const uint64_t bits = (uint64_t)len * 8;
  for (int i = 0; i < 8; i++) tail[padded - 8 + i] = (unsigned char)(bits >> (8 * i));
  h.block(tail);
  if (padded == 128) h.block(tail + 64);
  const uint32_t w0 = h.a, w1 = h.b;  // digest bytes 0..7 = a, b little-endian
  uint64_t v = 0;
  for (int i = 0; i < 4; i++) v = (v << 8) | ((w0 >> (8 * i)) & 0xffu);
  for (int i = 0; i < 4; i++) v = (v << 8) | ((w1 >> (8 * i)) & 0xffu);
  return (int64_t)v;
}

// ---- token parsing with the reference's (Java) acceptance rules --------------------------------
inline bool is_space(char c) {  // Splitter.trimResults(): whitespace (ASCII subset)
  return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r';
}
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }
inline bool is_hex(char c) { return is_digit(c) || (c >= 'a' && c <= 'f') || (c >= 'A' && c <= 'F'); }

// Long.parseLong: [+-]digits, no overflow.
bool parse_long(const char* s, size_t n, int64_t* out) {
  if (n == 0) return false;
  size_t i = 0;
  bool neg = false;
  if (s[0] == '-' || s[0] == '+') { neg = s[0] == '-'; i = 1; }
  if (i == n) return false;
  const uint64_t limit = neg ? (uint64_t)1 << 63 : ((uint64_t)1 << 63) - 1;
  uint64_t v = 0;
  for (; i < n; i++) {
    if (!is_digit(s[i])) return false;
    const uint64_t dgt = (uint64_t)(s[i] - '0');
    if (v > (limit - dgt) / 10) return false;
    v = v * 10 + dgt;
  }
  *out = neg ? (int64_t)(0 - v) : (int64_t)v;
  return true;
}

// LangUtils.parseFloat = Float.parseFloat + finiteness: [+-] (decimal | hex) literal with optional
// exponent and optional f/F/d/D suffix.  "NaN" / "Infinity" parse in Java but are rejected as
// non-finite, so they are simply refused here.
bool parse_float(const char* s, size_t n, float* out) {
  if (n == 0 || n > 200) return false;
  {
    // Fast path (Clinger): [+-]digits[.digits] with a mantissa below 2^24 and at most 10
    // fraction digits -- mantissa and power of ten are both exact in fp32, so the single
    // fp32 division is the correctly rounded result.
    static const float kPow10[11] = {1e0f, 1e1f, 1e2f, 1e3f, 1e4f, 1e5f, 1e6f, 1e7f, 1e8f, 1e9f, 1e10f};
    size_t k = 0;
    const bool neg = s[0] == '-';
    if (s[0] == '+' || s[0] == '-') k = 1;
    uint32_t m = 0;
    int frac = -1, digits = 0;
    bool simple = k < n;
    for (; k < n && simple; k++) {
      const char ch = s[k];
      if (is_digit(ch)) {
        m = m * 10 + (uint32_t)(ch - '0');
        digits++;
        if (frac >= 0) frac++;
        if (m >= (1u << 24) || frac > 10) simple = false;
      } else if (ch == '.' && frac < 0) {
        frac = 0;
      } else {
        simple = false;
      }
    }
    if (simple && digits > 0) {
      const float v = frac > 0 ? (float)m / kPow10[frac] : (float)m;
      *out = neg ? -v : v;
      return true;
    }
  }
  size_t i = 0;
  if (s[i] == '+' || s[i] == '-') i++;
  size_t end = n;
  if (end > i && (s[end - 1] == 'f' || s[end - 1] == 'F' || s[end - 1] == 'd' || s[end - 1] == 'D')) end--;
  if (end <= i) return false;
  if (end - i > 2 && s[i] == '0' && (s[i + 1] == 'x' || s[i + 1] == 'X')) {
    size_t k = i + 2, digits = 0;
    while (k < end && is_hex(s[k])) { k++; digits++; }
    if (k < end && s[k] == '.') { k++; while (k < end && is_hex(s[k])) { k++; digits++; } }
    if (digits == 0 || k >= end || (s[k] != 'p' && s[k] != 'P')) return false;  // binary exponent is mandatory
    k++;
    if (k < end && (s[k] == '+' || s[k] == '-')) k++;
    size_t ed = 0;
    while (k < end && is_digit(s[k])) { k++; ed++; }
    if (ed == 0 || k != end) return false;
  } else {
    // a hex literal's 'd'/'f' are digits, a decimal literal's are suffixes: handled by `end`
    size_t k = i, digits = 0;
    while (k < end && is_digit(s[k])) { k++; digits++; }
    if (k < end && s[k] == '.') { k++; while (k < end && is_digit(s[k])) { k++; digits++; } }
    if (digits == 0) return false;
    if (k < end && (s[k] == 'e' || s[k] == 'E')) {
      k++;
      if (k < end && (s[k] == '+' || s[k] == '-')) k++;
      size_t ed = 0;
      while (k < end && is_digit(s[k])) { k++; ed++; }
      if (ed == 0) return false;
    }
    if (k != end) return false;
  }
  char buf[208];
  memcpy(buf, s, end);
  buf[end] = 0;
  char* stop = nullptr;
  const float v = strtof(buf, &stop);  // correctly rounded, like Float.parseFloat
  if (stop != buf + end || !isfinite(v)) return false;
  *out = v;
  return true;
}

// ---- long ID -> provisional dense index (order of first appearance) ----------------------------
// Open addressing, linear probing, key and index in one 16-byte slot (one cache line per probe).
struct IdMap {
  struct Slot {
    int64_t key;
    uint32_t idx1;  // index + 1 (0 = empty)
    uint32_t pad;
  };
  std::vector<Slot> slots;
  std::vector<int64_t> ids;  // index -> key
  size_t mask = 0;
  IdMap() { rehash(1 << 12); }
  static uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
  }
  void rehash(size_t cap) {
    std::vector<Slot> t(cap, Slot{0, 0, 0});
    mask = cap - 1;
    for (size_t idx = 0; idx < ids.size(); idx++) {
      size_t s = mix((uint64_t)ids[idx]) & mask;
      while (t[s].idx1) s = (s + 1) & mask;
      t[s].key = ids[idx];
      t[s].idx1 = (uint32_t)idx + 1;
    }
    slots.swap(t);
  }
  void prefetch(int64_t key) const { __builtin_prefetch(&slots[mix((uint64_t)key) & mask]); }
  // returns the index, or UINT32_MAX when the 2^31-1 limit is hit
  uint32_t get_or_add(int64_t key) {
    size_t s = mix((uint64_t)key) & mask;
    while (slots[s].idx1) {
      if (slots[s].key == key) return slots[s].idx1 - 1;
      s = (s + 1) & mask;
    }
    if (ids.size() >= 0x7fffffffu) return 0xffffffffu;
    slots[s].key = key;
    slots[s].idx1 = (uint32_t)ids.size() + 1;
    ids.push_back(key);
    if (ids.size() * 10 > (mask + 1) * 7) rehash((mask + 1) * 2);
    return (uint32_t)ids.size() - 1;
  }
};

struct Event {
  uint64_t key;  // provisional user index << 32 | provisional item index
  float v;       // NaN = delete (InputFilesReader.java:160-165)
};

}  // namespace

struct ingest_handle {
  float zero_threshold = 1e-4f;
  IdMap users, items, item_tags, user_tags;
  std::vector<Event> events;
  long long lines = 0, bad_lines = 0;
  size_t min_chunk = 1u << 20;
  int max_threads = 0;  // 0 = all host threads
  bool finished = false;
  // results
  std::vector<int64_t> user_ids, item_ids;
  std::vector<int64_t> row_ptr, known_ptr;
  std::vector<int32_t> col_idx, known_idx;
  std::vector<float> val;
  std::string err;
};

namespace {

// Stable LSD radix sort of one bucket by (user, item).  Only the bits that vary are sorted on:
// the item index (< n_items) and the user index relative to a lower bound of the bucket's users.
void radix_sort_events(Event* base, size_t m, uint64_t user_lo, size_t n_items) {
  if (m < 2) return;
  if (m < 4096) {
    std::stable_sort(base, base + m, [](const Event& a, const Event& b) { return a.key < b.key; });
    return;
  }
  int ibits = 1;
  while (((uint64_t)1 << ibits) < (uint64_t)n_items) ibits++;
  uint64_t umax = 0;
  for (size_t e = 0; e < m; e++) umax = std::max(umax, (base[e].key >> 32) - user_lo);
  int ubits = 1;
  while (((uint64_t)1 << ubits) <= umax) ubits++;
  const int total = ibits + ubits;
  const int passes = (total + 10) / 11;
  const int digit = (total + passes - 1) / passes;  // <= 11 bits per pass
  const uint64_t imask = ((uint64_t)1 << ibits) - 1;
  auto shrink = [&](uint64_t key) { return (((key >> 32) - user_lo) << ibits) | (key & imask); };
  std::vector<Event> tmp(m);
  Event* src = base;
  Event* dst = tmp.data();
  std::vector<size_t> count((size_t)1 << digit);
  for (int p = 0; p < passes; p++) {
    const int shift = p * digit;
    const uint64_t dmask = ((uint64_t)1 << digit) - 1;
    std::fill(count.begin(), count.end(), 0);
    for (size_t e = 0; e < m; e++) count[(shrink(src[e].key) >> shift) & dmask]++;
    size_t acc = 0;
    for (size_t d = 0; d < count.size(); d++) { const size_t c = count[d]; count[d] = acc; acc += c; }
    for (size_t e = 0; e < m; e++) dst[count[(shrink(src[e].key) >> shift) & dmask]++] = src[e];
    std::swap(src, dst);
  }
  if (src != base) memcpy(base, src, m * sizeof(Event));
}

// ---- parsing: a file is cut into chunks at line boundaries and the chunks are parsed on all
// host threads; each chunk keeps its own ID maps (order of first appearance inside the chunk)
// and events with chunk-local indices.  Chunks are then stitched in file order, which
// reproduces exactly what a sequential reader would have assigned.
struct LocalEvent {
  uint32_t u, i;  // chunk-local indices
  float v;        // NaN = delete (InputFilesReader.java:160-165)
};

struct Chunk {
  const char* begin = nullptr;
  const char* end = nullptr;
  bool starts_input = false;  // its first line is line 1 of the whole input (header rule)
  IdMap users, items;
  std::vector<LocalEvent> events;
  std::vector<int64_t> item_tags, user_tags;  // in order of appearance (duplicates allowed)
  long long lines = 0, bad = 0;
  // ordinal (1-based, within the chunk) of the first 102 bad lines: enough to replay the
  // reference's "more than 100 bad lines and another line arrives" rule across chunks
  std::vector<long long> bad_at;
  bool overflow = false;  // more than 2^31-1 distinct IDs
  struct Pending { int64_t u, i; float v; };
  static constexpr int kBatch = 32;
  Pending pending[kBatch];
  int n_pending = 0;
};

void flush_pending(Chunk& c) {
  for (int k = 0; k < c.n_pending; k++) {
    const uint32_t u = c.users.get_or_add(c.pending[k].u);
    const uint32_t i = c.items.get_or_add(c.pending[k].i);
    if (u == 0xffffffffu || i == 0xffffffffu) { c.overflow = true; continue; }
    c.events.push_back(LocalEvent{u, i, c.pending[k].v});
  }
  c.n_pending = 0;
}

// One line (no terminator) of chunk c.
void take_line(Chunk& c, const char* s, size_t n) {
  c.lines++;
  if (n == 0 || s[0] == '#') return;
  // Splitter.on(',').trimResults(): only the first three fields matter
  const char* tok[3];
  size_t len[3];
  int nt = 0;
  size_t start = 0;
  for (size_t i = 0; i <= n && nt < 3; i++) {
    if (i == n || s[i] == ',') {
      size_t a = start, b = i;
      while (a < b && is_space(s[a])) a++;
      while (b > a && is_space(s[b - 1])) b--;
      tok[nt] = s + a;
      len[nt] = b - a;
      nt++;
      start = i + 1;
    }
  }
  auto bad = [&](bool forgivable) {
    // IllegalArgumentException on the first line = header (:135-141); too few columns is
    // always bad (:131-134)
    if (forgivable && c.starts_input && c.lines == 1) return;
    c.bad++;
    if (c.bad_at.size() < 102) c.bad_at.push_back(c.lines);
  };
  if (nt < 2) {
    // the user field is parsed before the item field is asked for: an unparseable user on a
    // one-field line is a NumberFormatException (forgivable), a parseable one NoSuchElement
    int64_t tmp;
    const bool tag = len[0] > 0 && tok[0][0] == '"';
    if (!tag && !parse_long(tok[0], len[0], &tmp)) return bad(true);
    return bad(false);
  }
  int64_t ids[2];
  bool is_tag[2];
  for (int f = 0; f < 2; f++) {
    is_tag[f] = len[f] > 0 && tok[f][0] == '"';
    if (is_tag[f]) {
      if (len[f] < 2) return bad(false);  // lone quote: the reference dies on substring(1, 0)
      ids[f] = md5_first8_be(tok[f] + 1, len[f] - 2);  // substring(1, length - 1) (:111-113)
    } else if (!parse_long(tok[f], len[f], &ids[f])) {
      return bad(true);
    }
  }
  float value = 1.0f;  // no third field (:127-129)
  if (nt == 3) {
    if (len[2] == 0) value = NAN;  // empty value = delete (:125)
    else if (!parse_float(tok[2], len[2], &value)) return bad(true);
  }
  if (is_tag[0] && is_tag[1]) return bad(false);       // two tags (:144-148)
  if (is_tag[0]) c.item_tags.push_back(ids[0]);        // itemTagIDs.add(userID)  (:150-152)
  if (is_tag[1]) c.user_tags.push_back(ids[1]);        // userTagIDs.add(itemID)  (:154-156)
  // the two map probes are cache misses on large inputs: queue the line behind a prefetch of
  // its slots and resolve a batch at a time (order is preserved)
  c.users.prefetch(ids[0]);
  c.items.prefetch(ids[1]);
  c.pending[c.n_pending++] = Chunk::Pending{ids[0], ids[1], value};
  if (c.n_pending == Chunk::kBatch) flush_pending(c);
}

void parse_chunk(Chunk& c) {
  const char* p = c.begin;
  while (p < c.end) {
    const char* e = p;
    while (e < c.end && *e != '\n' && *e != '\r') e++;
    take_line(c, p, (size_t)(e - p));
    if (e < c.end && *e == '\r' && e + 1 < c.end && e[1] == '\n') e++;  // \r\n
    p = e + 1;
  }
  flush_pending(c);
}

}  // namespace

extern "C" {

int ingest_create(float zero_threshold, ingest_handle** out) {
  if (!out || !(zero_threshold >= 0.f)) return INGEST_E_ARG;
  ingest_handle* h = new (std::nothrow) ingest_handle();
  if (!h) return INGEST_E_OOM;
  h->zero_threshold = zero_threshold;
  *out = h;
  return INGEST_OK;
}

void ingest_destroy(ingest_handle* h) { delete h; }

int ingest_set_parallelism(ingest_handle* h, int max_threads, size_t min_chunk_bytes) {
  if (!h || max_threads < 0) return INGEST_E_ARG;
  h->max_threads = max_threads;
  h->min_chunk = min_chunk_bytes ? min_chunk_bytes : 1;
  return INGEST_OK;
}

int ingest_add_file(ingest_handle* h, const char* data, size_t len) {
  if (!h || (!data && len)) return INGEST_E_ARG;
  if (h->finished) return INGEST_E_STATE;
  try {
    // cut at line boundaries (never between \r and \n)
    size_t parts = h->max_threads > 0 ? (size_t)h->max_threads : std::thread::hardware_concurrency();
    if (parts < 1) parts = 1;
    while (parts > 1 && len / parts < h->min_chunk) parts--;
    std::vector<Chunk> chunks(parts);
    const char* cut = data;
    for (size_t c = 0; c < parts; c++) {
      chunks[c].begin = cut;
      const char* e = (c + 1 == parts) ? data + len : data + len / parts * (c + 1);
      if (e < cut) e = cut;
      while (e < data + len && !(e > data && (e[-1] == '\n' || (e[-1] == '\r' && *e != '\n')))) e++;
      chunks[c].end = cut = e;
    }
    chunks[0].starts_input = (h->lines == 0);
    const bool trace = getenv("MYRRIX_INGEST_TRACE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
      return std::chrono::duration<double>(b - a).count();
    };
    const auto t0 = now();
    {
      std::vector<std::thread> th;
      for (size_t c = 1; c < parts; c++) th.emplace_back(parse_chunk, std::ref(chunks[c]));
      parse_chunk(chunks[0]);
      for (auto& t : th) t.join();
    }
    const auto t1 = now();
    // replay "if (badLines > 100) throw" (checked when a line arrives, :95-97) in file order
    for (Chunk& c : chunks) {
      if (c.lines > 0 && h->bad_lines > 100) { h->err = "Too many bad lines; aborting"; return INGEST_E_BAD_LINES; }
      const long long need = 101 - h->bad_lines;  // this chunk's bad line that makes the count 101
      if (need >= 1 && (long long)c.bad_at.size() >= need && c.bad_at[need - 1] < c.lines) {
        h->err = "Too many bad lines; aborting";
        return INGEST_E_BAD_LINES;
      }
      h->lines += c.lines;
      h->bad_lines += c.bad;
      if (c.overflow) { h->err = "more than 2^31-1 distinct users or items"; return INGEST_E_RANGE; }
    }
    // stitch: chunk-local indices -> global provisional indices, in file order (= the order a
    // sequential reader would have met the IDs in)
    std::vector<std::vector<uint32_t>> umap(parts), imap(parts);
    std::vector<size_t> offset(parts + 1, h->events.size());
    // (the user and the item map are independent: one thread each; probes run behind a prefetch)
    auto stitch = [&](bool users) {
      IdMap& g = users ? h->users : h->items;
      for (size_t c = 0; c < parts; c++) {
        const std::vector<int64_t>& ids = users ? chunks[c].users.ids : chunks[c].items.ids;
        std::vector<uint32_t>& m = users ? umap[c] : imap[c];
        m.resize(ids.size());
        constexpr size_t kAhead = 16;
        for (size_t k = 0; k < ids.size(); k++) {
          if (k + kAhead < ids.size()) g.prefetch(ids[k + kAhead]);
          m[k] = g.get_or_add(ids[k]);
        }
      }
    };
    {
      std::thread tu(stitch, true);
      stitch(false);
      tu.join();
    }
    for (size_t c = 0; c < parts; c++) {
      Chunk& ch = chunks[c];
      for (int64_t t : ch.item_tags) h->item_tags.get_or_add(t);
      for (int64_t t : ch.user_tags) h->user_tags.get_or_add(t);
      offset[c + 1] = offset[c] + ch.events.size();
    }
    if (h->users.ids.size() >= 0x7fffffffu || h->items.ids.size() >= 0x7fffffffu) {
      h->err = "more than 2^31-1 distinct users or items";
      return INGEST_E_RANGE;
    }
    const auto t2 = now();
    h->events.resize(offset[parts]);
    {
      auto emit = [&](size_t c) {
        Event* out = h->events.data() + offset[c];
        const std::vector<uint32_t>&um = umap[c], &im = imap[c];
        for (const LocalEvent& e : chunks[c].events)
          *out++ = Event{((uint64_t)um[e.u] << 32) | im[e.i], e.v};
      };
      std::vector<std::thread> th;
      for (size_t c = 1; c < parts; c++) th.emplace_back(emit, c);
      emit(0);
      for (auto& t : th) t.join();
    }
    if (trace)
      fprintf(stderr, "ingest_add_file: %zu chunks, parse %.3f s, stitch %.3f s, emit %.3f s\n", parts,
              secs(t0, t1), secs(t1, t2), secs(t2, now()));
  } catch (const std::bad_alloc&) {
    h->err = "out of memory";
    return INGEST_E_OOM;
  }
  return INGEST_OK;
}

int ingest_finish(ingest_handle* h) {
  if (!h) return INGEST_E_ARG;
  if (h->finished) return INGEST_E_STATE;
  try {
    // Events are partitioned by user range into one bucket per thread (stable: input order
    // inside a bucket is kept), and every bucket is sorted by (user, item), folded and turned
    // into its slice of the CSR independently -- users never straddle buckets.
    std::vector<Event>& ev = h->events;
    const size_t n = ev.size();
    const size_t nu0 = h->users.ids.size(), ni0 = h->items.ids.size();
    size_t T = h->max_threads > 0 ? (size_t)h->max_threads : std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    while (T > 1 && n < 1024 * T) T--;
    auto run = [&](auto&& fn) {  // fn(t) for t in [0, T) on T threads
      std::vector<std::thread> th;
      for (size_t t = 1; t < T; t++) th.emplace_back(fn, t);
      fn((size_t)0);
      for (auto& x : th) x.join();
    };
    auto bucket_of = [&](uint64_t key) { return (size_t)((key >> 32) * T / (nu0 ? nu0 : 1)); };
    auto slice = [&](size_t t) { return n / T * t + std::min(t, n % T); };

    // 1. stable partition by bucket
    std::vector<size_t> bstart(T + 1, 0);
    if (T > 1) {
      std::vector<std::vector<size_t>> cnt(T, std::vector<size_t>(T, 0));
      run([&](size_t t) {
        for (size_t e = slice(t); e < slice(t + 1); e++) cnt[t][bucket_of(ev[e].key)]++;
      });
      std::vector<std::vector<size_t>> pos(T, std::vector<size_t>(T, 0));
      size_t acc = 0;
      for (size_t b = 0; b < T; b++) {
        bstart[b] = acc;
        for (size_t t = 0; t < T; t++) { pos[t][b] = acc; acc += cnt[t][b]; }
      }
      bstart[T] = acc;
      std::vector<Event> tmp(n);
      run([&](size_t t) {
        std::vector<size_t> p = pos[t];
        for (size_t e = slice(t); e < slice(t + 1); e++) tmp[p[bucket_of(ev[e].key)]++] = ev[e];
      });
      ev.swap(tmp);
    } else {
      bstart[1] = n;
    }

    // 2. per bucket: sort, fold every cell in input order -- increment (fp32 sum) or remove
    //    (FastByIDFloatMap.increment :129-138; MatrixUtils.removeByRow :112-121); surviving
    //    cells are compacted to the front of the bucket
    std::vector<uint8_t> user_alive(nu0, 0);
    std::vector<std::vector<uint8_t>> item_alive_t(T, std::vector<uint8_t>(ni0, 0));
    std::vector<size_t> kept(T, 0), kept_unpruned(T, 0);
    const float zt = h->zero_threshold;
    run([&](size_t t) {
      Event* base = ev.data() + bstart[t];
      const size_t m = bstart[t + 1] - bstart[t];
      radix_sort_events(base, m, nu0 ? (uint64_t)t * nu0 / T : 0, ni0);
      std::vector<uint8_t>& item_alive = item_alive_t[t];
      size_t w = 0, unpruned = 0;
      for (size_t a = 0; a < m;) {
        size_t b = a;
        bool present = false;
        float sum = 0.f;
        for (; b < m && base[b].key == base[a].key; b++) {
          if (isnan(base[b].v)) present = false;
          else if (!present) { present = true; sum = base[b].v; }
          else sum = sum + base[b].v;
        }
        if (present) {
          user_alive[base[a].key >> 32] = 1;  // users are private to the bucket
          item_alive[base[a].key & 0xffffffffu] = 1;
          base[w++] = Event{base[a].key, sum};
          // removeSmall (:198-211): |v| < threshold leaves the matrix, the (possibly empty) row stays
          if (!(fabsf(sum) < zt)) unpruned++;
        }
        a = b;
      }
      kept[t] = w;
      kept_unpruned[t] = unpruned;
    });

    // 3. final dense indices: order of first appearance among the survivors (rows whose
    //    entries were all deleted left the maps, MatrixUtils.java:116-119)
    std::vector<uint32_t> umap(nu0, 0), imap(ni0, 0);
    for (size_t u = 0; u < nu0; u++)
      if (user_alive[u]) { umap[u] = (uint32_t)h->user_ids.size(); h->user_ids.push_back(h->users.ids[u]); }
    for (size_t i = 0; i < ni0; i++) {
      bool alive = false;
      for (size_t t = 0; t < T && !alive; t++) alive = item_alive_t[t][i] != 0;
      if (alive) { imap[i] = (uint32_t)h->item_ids.size(); h->item_ids.push_back(h->items.ids[i]); }
    }
    const size_t nu = h->user_ids.size();

    // 4. CSR slices per bucket
    std::vector<size_t> koff(T + 1, 0), coff(T + 1, 0);
    for (size_t t = 0; t < T; t++) { koff[t + 1] = koff[t] + kept[t]; coff[t + 1] = coff[t] + kept_unpruned[t]; }
    h->row_ptr.assign(nu + 1, 0);
    h->known_ptr.assign(nu + 1, 0);
    h->known_idx.resize(koff[T]);
    h->col_idx.resize(coff[T]);
    h->val.resize(coff[T]);
    run([&](size_t t) {
      const Event* base = ev.data() + bstart[t];
      size_t ko = koff[t], co = coff[t];
      for (size_t e = 0; e < kept[t]; e++) {
        const uint32_t u = umap[base[e].key >> 32], i = imap[base[e].key & 0xffffffffu];
        h->known_ptr[u + 1]++;
        h->known_idx[ko++] = (int32_t)i;
        if (!(fabsf(base[e].v) < zt)) {
          h->row_ptr[u + 1]++;
          h->col_idx[co] = (int32_t)i;
          h->val[co++] = base[e].v;
        }
      }
    });
    for (size_t u = 0; u < nu; u++) {
      h->row_ptr[u + 1] += h->row_ptr[u];
      h->known_ptr[u + 1] += h->known_ptr[u];
    }
    std::vector<Event>().swap(ev);
  } catch (const std::bad_alloc&) {
    h->err = "out of memory";
    return INGEST_E_OOM;
  }
  h->finished = true;
  return INGEST_OK;
}

int64_t ingest_count(const ingest_handle* h, int kind) {
  if (!h) return -1;
  switch (kind) {
    case INGEST_LINES: return h->lines;
    case INGEST_BAD_LINES: return h->bad_lines;
    case INGEST_N_ITEM_TAGS: return (int64_t)h->item_tags.ids.size();
    case INGEST_N_USER_TAGS: return (int64_t)h->user_tags.ids.size();
    default: break;
  }
  if (!h->finished) return -1;
  switch (kind) {
    case INGEST_N_USERS: return (int64_t)h->user_ids.size();
    case INGEST_N_ITEMS: return (int64_t)h->item_ids.size();
    case INGEST_NNZ: return (int64_t)h->col_idx.size();
    case INGEST_KNOWN_NNZ: return (int64_t)h->known_idx.size();
    default: return -1;
  }
}

int ingest_get_ids(const ingest_handle* h, int which, int64_t* out) {
  if (!h || !out || (which != 0 && which != 1)) return INGEST_E_ARG;
  if (!h->finished) return INGEST_E_STATE;
  const std::vector<int64_t>& v = which == 0 ? h->user_ids : h->item_ids;
  if (!v.empty()) memcpy(out, v.data(), v.size() * sizeof(int64_t));
  return INGEST_OK;
}

int ingest_get_csr(const ingest_handle* h, int64_t* row_ptr, int32_t* col_idx, float* val) {
  if (!h || !row_ptr || (!col_idx && !h->col_idx.empty()) || (!val && !h->val.empty())) return INGEST_E_ARG;
  if (!h->finished) return INGEST_E_STATE;
  memcpy(row_ptr, h->row_ptr.data(), h->row_ptr.size() * sizeof(int64_t));
  if (!h->col_idx.empty()) {
    memcpy(col_idx, h->col_idx.data(), h->col_idx.size() * sizeof(int32_t));
    memcpy(val, h->val.data(), h->val.size() * sizeof(float));
  }
  return INGEST_OK;
}

int ingest_get_known(const ingest_handle* h, int64_t* row_ptr, int32_t* col_idx) {
  if (!h || !row_ptr || (!col_idx && !h->known_idx.empty())) return INGEST_E_ARG;
  if (!h->finished) return INGEST_E_STATE;
  memcpy(row_ptr, h->known_ptr.data(), h->known_ptr.size() * sizeof(int64_t));
  if (!h->known_idx.empty()) memcpy(col_idx, h->known_idx.data(), h->known_idx.size() * sizeof(int32_t));
  return INGEST_OK;
}

int ingest_get_tags(const ingest_handle* h, int which, int64_t* out) {
  if (!h || !out || (which != 0 && which != 1)) return INGEST_E_ARG;
  const std::vector<int64_t>& v = which == 0 ? h->item_tags.ids : h->user_tags.ids;
  if (!v.empty()) memcpy(out, v.data(), v.size() * sizeof(int64_t));
  return INGEST_OK;
}

int64_t ingest_tag_id(const char* utf8, size_t len) { return md5_first8_be(utf8 ? utf8 : "", len); }

const char* ingest_last_error(const ingest_handle* h) { return h ? h->err.c_str() : "null handle"; }

}  // extern "C"
