// TEST INFRASTRUCTURE: sanitizer fuzz driver for the host record decoders of libesr (csrc/record_decode.cu is plain
// host C++; it is compiled here with -fsanitize=address,undefined next to this driver, no CUDA involved).
// Inputs: valid streams built by a small encoder (CooccurrenceRow base64 lines / TFRecord tf.train.Example), then random
// mutations (bit flips, truncation, spliced garbage, hostile lengths).  Every input and output buffer is an exact-size
// heap allocation, so any out-of-bounds access aborts the run.  Valid streams must decode to what was encoded; mutated
// ones must return a count or a negative ESR_E* code.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "esr.h"

static uint64_t g_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() {
  g_state ^= g_state << 13;
  g_state ^= g_state >> 7;
  g_state ^= g_state << 17;
  return g_state;
}
static uint32_t rnd_below(uint32_t n) { return n ? (uint32_t)(rnd() % n) : 0; }

static void put_varint(std::string& s, uint64_t v) {
  while (v >= 0x80) {
    s.push_back((char)((v & 0x7F) | 0x80));
    v >>= 7;
  }
  s.push_back((char)v);
}
static void put_bytes(std::string& s, uint32_t field, const std::string& payload) {
  put_varint(s, (field << 3) | 2);
  put_varint(s, payload.size());
  s += payload;
}
static std::string b64(const std::string& in) {
  static const char* a = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
  std::string out;
  size_t i = 0;
  for (; i + 2 < in.size(); i += 3) {
    uint32_t v = ((unsigned char)in[i] << 16) | ((unsigned char)in[i + 1] << 8) | (unsigned char)in[i + 2];
    out += a[v >> 18]; out += a[(v >> 12) & 63]; out += a[(v >> 6) & 63]; out += a[v & 63];
  }
  if (i + 1 == in.size()) {
    uint32_t v = (unsigned char)in[i] << 16;
    out += a[v >> 18]; out += a[(v >> 12) & 63]; out += "==";
  } else if (i + 2 == in.size()) {
    uint32_t v = ((unsigned char)in[i] << 16) | ((unsigned char)in[i + 1] << 8);
    out += a[v >> 18]; out += a[(v >> 12) & 63]; out += a[(v >> 6) & 63]; out += '=';
  }
  return out;
}

struct Triples {
  std::vector<int32_t> i, j;
  std::vector<float> c;
};

static std::string make_cooccur(Triples* t, int rows) {
  std::string text;
  for (int r = 0; r < rows; ++r) {
    const uint32_t index = rnd_below(1u << 20);
    const int m = (int)rnd_below(40);
    std::string other, count, msg;
    for (int k = 0; k < m; ++k) {
      const uint32_t o = rnd_below(1u << 20);
      const float f = (float)(rnd_below(100000)) / 7.0f;
      put_varint(other, o);
      count.append(reinterpret_cast<const char*>(&f), 4);
      t->i.push_back((int32_t)index); t->j.push_back((int32_t)o); t->c.push_back(f);
    }
    put_varint(msg, (1 << 3) | 0);
    put_varint(msg, index);
    if (m) {
      put_bytes(msg, 2, other);
      put_bytes(msg, 3, count);
    }
    text += b64(msg);
    text += '\n';
  }
  return text;
}

static std::string make_tfrecords(std::vector<std::vector<int64_t>>* per_key, int n_keys, const char* const* keys, int records) {
  std::string out;
  for (int r = 0; r < records; ++r) {
    std::string features;
    for (int k = 0; k < n_keys; ++k) {
      const int m = (int)rnd_below(9);
      std::string packed;
      for (int q = 0; q < m; ++q) {
        const int64_t v = (int64_t)(rnd() >> (rnd_below(2) ? 40 : 1));
        put_varint(packed, (uint64_t)v);
        (*per_key)[k].push_back(v);
      }
      (*per_key)[k].push_back(INT64_MIN);  // record separator in the expectation
      std::string list, feature, entry;
      put_bytes(list, 1, packed);
      put_bytes(feature, 3, list);
      put_bytes(entry, 1, keys[k]);
      put_bytes(entry, 2, feature);
      put_bytes(features, 1, entry);
    }
    std::string ex;
    put_bytes(ex, 1, features);
    const uint64_t len = ex.size();
    out.append(reinterpret_cast<const char*>(&len), 8);
    out.append("\0\0\0\0", 4);
    out += ex;
    out.append("\0\0\0\0", 4);
  }
  return out;
}

static void mutate(std::string& s) {
  if (s.empty()) return;
  switch (rnd_below(6)) {
    case 0: for (int k = 0, n = 1 + (int)rnd_below(8); k < n; ++k) s[rnd_below((uint32_t)s.size())] ^= (char)(1u << rnd_below(8)); break;
    case 1: s.resize(rnd_below((uint32_t)s.size())); break;
    case 2: { const uint32_t at = rnd_below((uint32_t)s.size()); for (int k = 0, n = 1 + (int)rnd_below(16); k < n; ++k) s.insert(s.begin() + at, (char)rnd()); } break;
    case 3: { const uint32_t at = rnd_below((uint32_t)s.size()); for (uint32_t k = at; k < s.size() && k < at + 10; ++k) s[k] = (char)0xFF; } break;  // hostile varints / lengths
    case 4: for (auto& ch : s) if (rnd_below(50) == 0) ch = (char)rnd(); break;
    default: { const uint32_t at = rnd_below((uint32_t)s.size()); s[at] = '\n'; } break;
  }
}

template <class T>
struct Exact {  // exact-size heap block: one byte past the end is poisoned by ASAN
  T* p;
  size_t n;
  explicit Exact(size_t n_) : p((T*)malloc(n_ ? n_ * sizeof(T) : 1)), n(n_) {}
  ~Exact() { free(p); }
};

static int fail(const char* what, int iter) {
  fprintf(stderr, "decode_fuzz: %s (iteration %d)\n", what, iter);
  return 1;
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  if (argc > 2) g_state ^= strtoull(argv[2], nullptr, 10) * 0x2545F4914F6CDD1Dull;
  const char* keys[3] = {"track_context", "next_track", "neg_track"};
  for (int it = 0; it < iters; ++it) {
    // ---------------- CooccurrenceRow text ----------------
    {
      Triples t;
      std::string text = make_cooccur(&t, 1 + (int)rnd_below(12));
      const bool valid = rnd_below(3) == 0;
      if (!valid) mutate(text);
      const int64_t cap = valid && rnd_below(2) ? (int64_t)t.i.size() : (int64_t)rnd_below((uint32_t)t.i.size() + 8);
      Exact<char> in(text.size());
      memcpy(in.p, text.data(), text.size());
      Exact<int32_t> oi((size_t)cap), oj((size_t)cap);
      Exact<float> oc((size_t)cap);
      int64_t rows = -7;
      size_t used = (size_t)-1;
      const int64_t n = esr_decode_cooccur_b64(in.p, text.size(), oi.p, oj.p, oc.p, cap, &rows, &used);
      if (n < 0) {
        if (valid) return fail("valid cooccur text rejected", it);
        if (n != ESR_EINVAL) return fail("unexpected error code (cooccur)", it);
      } else {
        if (n > cap || used > text.size() || rows < 0) return fail("cooccur result out of range", it);
        if (valid) {
          for (int64_t k = 0; k < n; ++k)
            if (oi.p[k] != t.i[k] || oj.p[k] != t.j[k] || memcmp(&oc.p[k], &t.c[k], 4) != 0) return fail("cooccur triple mismatch", it);
          if (cap >= (int64_t)t.i.size() && (n != (int64_t)t.i.size() || used != text.size())) return fail("cooccur stream not fully decoded", it);
        }
      }
    }
    // ---------------- TFRecord / tf.train.Example ----------------
    {
      std::vector<std::vector<int64_t>> exp(3);
      const int records = 1 + (int)rnd_below(6);
      std::string data = make_tfrecords(&exp, 3, keys, records);
      const bool valid = rnd_below(3) == 0;
      if (!valid) mutate(data);
      const int64_t max_records = valid ? records : (int64_t)rnd_below(10);
      int64_t caps[3];
      for (int k = 0; k < 3; ++k) caps[k] = valid ? (int64_t)exp[k].size() : (int64_t)rnd_below(40);
      Exact<uint8_t> in(data.size());
      memcpy(in.p, data.data(), data.size());
      Exact<int64_t> v0((size_t)caps[0]), v1((size_t)caps[1]), v2((size_t)caps[2]);
      Exact<int64_t> o0((size_t)max_records + 1), o1((size_t)max_records + 1), o2((size_t)max_records + 1);
      int64_t* vals[3] = {v0.p, v1.p, v2.p};
      int64_t* offs[3] = {o0.p, o1.p, o2.p};
      size_t used = (size_t)-1;
      const int64_t n = esr_decode_tfrecord_int64(in.p, data.size(), 3, keys, vals, caps, offs, max_records, &used);
      if (n < 0) {
        if (valid) return fail("valid TFRecord stream rejected", it);
        if (n != ESR_EINVAL) return fail("unexpected error code (tfrecord)", it);
      } else {
        if (n > max_records || used > data.size()) return fail("tfrecord result out of range", it);
        for (int k = 0; k < 3; ++k)
          for (int64_t r = 0; r < n; ++r)
            if (offs[k][r] > offs[k][r + 1] || offs[k][r + 1] > caps[k]) return fail("tfrecord offsets not monotone / out of range", it);
        if (valid) {
          if (n != records || used != data.size()) return fail("tfrecord stream not fully decoded", it);
          for (int k = 0; k < 3; ++k) {
            size_t e = 0;
            for (int64_t r = 0; r < n; ++r) {
              for (int64_t q = offs[k][r]; q < offs[k][r + 1]; ++q)
                if (exp[k][e++] != vals[k][q]) return fail("tfrecord value mismatch", it);
              if (exp[k][e++] != INT64_MIN) return fail("tfrecord record boundary mismatch", it);
            }
          }
        }
      }
    }
  }
  printf("decode_fuzz ok: %d iterations\n", iters);
  return 0;
}
