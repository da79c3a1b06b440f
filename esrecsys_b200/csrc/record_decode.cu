// Native decoders of the reference's on-disk record formats (HOST code).
// The reference decodes these in pure Python (wikipedia/cooccurrence_matrix.py:62-83: bz2 -> line -> base64 ->
// protobuf ParseFromString -> per-element yield) or through tf.data (spotify/input_pipeline.py:23-49); once the
// training step runs at > 1 G pairs/s that decoder is the end-to-end bottleneck by orders of magnitude.
//
//  esr_decode_cooccur_b64      text of a *.cooccur.pb.b64 part (after bz2): one base64 line per
//                              topicspace.corpus.nlp.CooccurrenceRow {1: index varint, 2: packed varint other_index,
//                              3: packed float count}  ->  (i, j, count) triples, exactly the order get_item() yields
//  esr_decode_tfrecord_int64   TFRecord framing + tf.train.Example{features{feature: map<string, Feature{int64_list}>}}
//                              -> per-key concatenated int64 values + per-record offsets (spotify/input_pipeline.py:23-37)
// Both are re-entrant and allocation-free (the caller owns every buffer), so Python can run one call per file on a
// thread pool: ctypes releases the GIL.
#include <stdint.h>
#include <string.h>

#include <new>
#include <vector>

#include "esr.h"

namespace {

struct B64 {
  int8_t t[256];
  B64() {
    memset(t, -1, sizeof(t));
    const char* a = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    for (int i = 0; i < 64; ++i) t[(unsigned char)a[i]] = (int8_t)i;
  }
};
const B64 g_b64;

// decodes [s, e) (standard alphabet, '=' padding) into out; returns bytes written or -1
int64_t b64_decode(const unsigned char* s, const unsigned char* e, unsigned char* out) {
  uint32_t acc = 0;
  int bits = 0;
  int64_t n = 0;
  for (; s < e; ++s) {
    if (*s == '=') break;
    const int8_t v = g_b64.t[*s];
    if (v < 0) {
      if (*s == '\r') continue;
      return -1;
    }
    acc = (acc << 6) | (uint32_t)v;
    bits += 6;
    if (bits >= 8) {
      bits -= 8;
      out[n++] = (unsigned char)((acc >> bits) & 0xFF);
    }
  }
  return n;
}

inline bool varint(const unsigned char*& p, const unsigned char* e, uint64_t* v) {
  uint64_t r = 0;
  for (int shift = 0; p < e && shift < 64; shift += 7) {
    const unsigned char b = *p++;
    r |= (uint64_t)(b & 0x7F) << shift;
    if (!(b & 0x80)) {
      *v = r;
      return true;
    }
  }
  return false;
}

inline bool skip_field(const unsigned char*& p, const unsigned char* e, uint32_t wire) {
  uint64_t v;
  switch (wire) {
    case 0: return varint(p, e, &v);
    case 1: if (e - p < 8) return false; p += 8; return true;
    case 2: if (!varint(p, e, &v) || (uint64_t)(e - p) < v) return false; p += v; return true;
    case 5: if (e - p < 4) return false; p += 4; return true;
    default: return false;
  }
}

}  // namespace

extern "C" int64_t esr_decode_cooccur_b64(const char* text, size_t n_bytes, int32_t* out_i, int32_t* out_j, float* out_count,
                                          int64_t cap, int64_t* n_rows, size_t* consumed) {
  if (!text || !out_i || !out_j || !out_count || cap < 0) return ESR_EINVAL;
  const unsigned char* p = reinterpret_cast<const unsigned char*>(text);
  const unsigned char* end = p + n_bytes;
  // <= 1001 entries per row message at the reference's default --max_row_size (wikipedia/make_cooccurrence.py:87): ~9 KB
  // worst case, decoded on the stack; longer lines (a larger --max_row_size, many 5-byte varints) take a heap buffer.
  unsigned char stack_buf[16384];
  std::vector<unsigned char> heap_buf;
  unsigned char* buf = stack_buf;
  size_t buf_cap = sizeof(stack_buf);
  int64_t n = 0, rows = 0;
  while (p < end) {
    const unsigned char* nl = static_cast<const unsigned char*>(memchr(p, '\n', (size_t)(end - p)));
    if (!nl) break;  // incomplete last line: the caller resumes from `consumed`
    if (nl == p) {
      p = nl + 1;
      continue;
    }
    const size_t need = (size_t)(nl - p) / 4 * 3 + 3;
    if (need > buf_cap) {
      try {
        heap_buf.resize(need + need / 2);
      } catch (const std::bad_alloc&) {
        return ESR_ENOMEM;  // not "malformed": the row is valid, the host is out of memory
      }
      buf = heap_buf.data();
      buf_cap = heap_buf.size();
    }
    const int64_t len = b64_decode(p, nl, buf);
    if (len < 0) return ESR_EINVAL;
    // ---- CooccurrenceRow ----
    const unsigned char* q = buf;
    const unsigned char* qe = buf + len;
    uint64_t index = 0;
    const int64_t first = n;
    int64_t n_other = 0, n_count = 0;
    bool full = false;
    while (q < qe && !full) {
      uint64_t tag;
      if (!varint(q, qe, &tag)) return ESR_EINVAL;
      const uint32_t field = (uint32_t)(tag >> 3), wire = (uint32_t)(tag & 7);
      if (field == 1 && wire == 0) {
        if (!varint(q, qe, &index)) return ESR_EINVAL;
      } else if (field == 2 && (wire == 2 || wire == 0)) {
        uint64_t l = 0;
        const unsigned char* fe = qe;
        if (wire == 2) {
          if (!varint(q, qe, &l) || (uint64_t)(qe - q) < l) return ESR_EINVAL;
          fe = q + l;
        }
        do {
          uint64_t v;
          if (!varint(q, fe, &v)) return ESR_EINVAL;
          if (first + n_other >= cap) {
            full = true;
            break;
          }
          out_j[first + n_other++] = (int32_t)v;
        } while (wire == 2 && q < fe);
      } else if (field == 3 && (wire == 2 || wire == 5)) {
        uint64_t l = 4;
        if (wire == 2 && (!varint(q, qe, &l) || (l & 3))) return ESR_EINVAL;
        if ((uint64_t)(qe - q) < l) return ESR_EINVAL;
        for (uint64_t k = 0; k < l; k += 4) {
          if (first + n_count >= cap) {
            full = true;
            break;
          }
          float f;
          memcpy(&f, q + k, 4);  // little-endian host
          out_count[first + n_count++] = f;
        }
        q += l;
      } else if (!skip_field(q, qe, wire)) {
        return ESR_EINVAL;
      }
    }
    if (full) break;  // this row does not fit: stop BEFORE it
    if (n_other != n_count) return ESR_EINVAL;  // the reader indexes count[i] for i < len(other_index)
    for (int64_t k = 0; k < n_other; ++k) out_i[first + k] = (int32_t)index;
    n += n_other;
    ++rows;
    p = nl + 1;
  }
  if (n_rows) *n_rows = rows;
  if (consumed) *consumed = (size_t)(p - reinterpret_cast<const unsigned char*>(text));
  return n;
}

// TFRecord: u64 length | u32 masked crc32c(length) | data | u32 masked crc32c(data).  CRCs are not verified
// (tf.data verifies them; corrupt files are out of scope for the hot path).
// For each of the n_keys feature names: values of record r are vals[k][off[k][r] .. off[k][r+1]).
extern "C" int64_t esr_decode_tfrecord_int64(const uint8_t* data, size_t n_bytes, int32_t n_keys, const char* const* keys,
                                             int64_t* const* vals, const int64_t* val_cap, int64_t* const* offs,
                                             int64_t max_records, size_t* consumed) {
  if (!data || n_keys <= 0 || n_keys > 16 || !keys || !vals || !val_cap || !offs || max_records < 0) return ESR_EINVAL;
  size_t klen[16];
  int64_t fill[16];
  for (int k = 0; k < n_keys; ++k) {
    if (!keys[k] || !vals[k] || !offs[k]) return ESR_EINVAL;
    klen[k] = strlen(keys[k]);
    fill[k] = 0;
    offs[k][0] = 0;
  }
  const unsigned char* p = data;
  const unsigned char* end = data + n_bytes;
  int64_t rec = 0;
  while (rec < max_records && (size_t)(end - p) >= 12) {
    uint64_t len;
    memcpy(&len, p, 8);
    const uint64_t avail = (uint64_t)(end - p);
    if (avail < 16 || len > avail - 16) break;  // incomplete record (also a corrupt length: 16 + len must not wrap)
    const unsigned char* q = p + 12;
    const unsigned char* qe = q + len;
    int64_t start[16];
    for (int k = 0; k < n_keys; ++k) start[k] = fill[k];
    bool full = false;
    // Example { 1: Features { repeated 1: map entry { 1: string key, 2: Feature { 3: Int64List { 1: packed/unpacked } } } } }
    while (q < qe && !full) {
      uint64_t tag, l;
      if (!varint(q, qe, &tag)) return ESR_EINVAL;
      if (tag != ((1u << 3) | 2)) {
        if (!skip_field(q, qe, (uint32_t)(tag & 7))) return ESR_EINVAL;
        continue;
      }
      if (!varint(q, qe, &l) || (uint64_t)(qe - q) < l) return ESR_EINVAL;
      const unsigned char* f = q;
      const unsigned char* fe = q + l;
      q = fe;
      while (f < fe && !full) {  // Features
        if (!varint(f, fe, &tag)) return ESR_EINVAL;
        if (tag != ((1u << 3) | 2)) {
          if (!skip_field(f, fe, (uint32_t)(tag & 7))) return ESR_EINVAL;
          continue;
        }
        if (!varint(f, fe, &l) || (uint64_t)(fe - f) < l) return ESR_EINVAL;
        const unsigned char* m = f;
        const unsigned char* me = f + l;
        f = me;
        const unsigned char* name = nullptr;
        uint64_t name_len = 0;
        const unsigned char* feat = nullptr;
        const unsigned char* feat_e = nullptr;
        while (m < me) {  // map entry
          if (!varint(m, me, &tag)) return ESR_EINVAL;
          if ((tag & 7) != 2) {
            if (!skip_field(m, me, (uint32_t)(tag & 7))) return ESR_EINVAL;
            continue;
          }
          if (!varint(m, me, &l) || (uint64_t)(me - m) < l) return ESR_EINVAL;
          if ((tag >> 3) == 1) {
            name = m;
            name_len = l;
          } else if ((tag >> 3) == 2) {
            feat = m;
            feat_e = m + l;
          }
          m += l;
        }
        int k = -1;
        for (int c = 0; c < n_keys && name; ++c)
          if (klen[c] == name_len && memcmp(keys[c], name, name_len) == 0) k = c;
        if (k < 0 || !feat) continue;
        while (feat < feat_e && !full) {  // Feature: field 3 = Int64List
          if (!varint(feat, feat_e, &tag)) return ESR_EINVAL;
          if (tag != ((3u << 3) | 2)) {
            if (!skip_field(feat, feat_e, (uint32_t)(tag & 7))) return ESR_EINVAL;
            continue;
          }
          if (!varint(feat, feat_e, &l) || (uint64_t)(feat_e - feat) < l) return ESR_EINVAL;
          const unsigned char* v = feat;
          const unsigned char* ve = feat + l;
          feat = ve;
          while (v < ve && !full) {  // Int64List: field 1, packed (wire 2) or repeated varint (wire 0)
            if (!varint(v, ve, &tag)) return ESR_EINVAL;
            if (tag == ((1u << 3) | 2)) {
              if (!varint(v, ve, &l) || (uint64_t)(ve - v) < l) return ESR_EINVAL;
              const unsigned char* pe = v + l;
              while (v < pe) {
                uint64_t x;
                if (!varint(v, pe, &x)) return ESR_EINVAL;
                if (fill[k] >= val_cap[k]) {
                  full = true;
                  break;
                }
                vals[k][fill[k]++] = (int64_t)x;
              }
            } else if (tag == ((1u << 3) | 0)) {
              uint64_t x;
              if (!varint(v, ve, &x)) return ESR_EINVAL;
              if (fill[k] >= val_cap[k]) full = true;
              else vals[k][fill[k]++] = (int64_t)x;
            } else if (!skip_field(v, ve, (uint32_t)(tag & 7))) {
              return ESR_EINVAL;
            }
          }
        }
      }
    }
    if (full) {  // roll this record back and stop before it
      for (int k = 0; k < n_keys; ++k) fill[k] = start[k];
      break;
    }
    ++rec;
    for (int k = 0; k < n_keys; ++k) offs[k][rec] = fill[k];
    p += 16 + len;
  }
  if (consumed) *consumed = (size_t)(p - data);
  return rec;
}
