// Host-side Fiat-Shamir transcript: Merlin v1.0 over STROBE-128 / Keccak-f[1600], and
// SHAKE256 for generator derivation. In production the Rust host keeps using the `merlin`
// and `sha3` crates (BASELINE.json north_star); without a Rust toolchain this C++ mirror
// stands in for them so that whole proofs can be produced and compared byte for byte.
//   reference call sites: src/transcript.rs:13-37, src/random.rs:9-28, src/commitments.rs:15-33
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace sph {

inline uint64_t rol64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

inline void keccak_f1600(uint64_t A[25]) {
  static const uint64_t RC[24] = {
      0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL,
      0x000000000000808BULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
      0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,
      0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
      0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
      0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  static const int ROT[5][5] = {{0, 36, 3, 41, 18}, {1, 44, 10, 45, 2}, {62, 6, 43, 15, 61},
                                {28, 55, 25, 21, 56}, {27, 20, 39, 8, 14}};
  for (int rnd = 0; rnd < 24; rnd++) {
    uint64_t C[5], D[5], B[25];
    for (int x = 0; x < 5; x++) C[x] = A[x] ^ A[x + 5] ^ A[x + 10] ^ A[x + 15] ^ A[x + 20];
    for (int x = 0; x < 5; x++) D[x] = C[(x + 4) % 5] ^ rol64(C[(x + 1) % 5], 1);
    for (int i = 0; i < 25; i++) A[i] ^= D[i % 5];
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) B[y + 5 * ((2 * x + 3 * y) % 5)] = rol64(A[x + 5 * y], ROT[x][y]);
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) A[x + 5 * y] = B[x + 5 * y] ^ ((~B[(x + 1) % 5 + 5 * y]) & B[(x + 2) % 5 + 5 * y]);
    A[0] ^= RC[rnd];
  }
}

class Strobe128 {
 public:
  explicit Strobe128(const std::string &label) {
    memset(st_, 0, sizeof st_);
    const uint8_t init[6] = {1, R + 2, 1, 0, 1, 96};
    memcpy(st_, init, 6);
    memcpy(st_ + 6, "STROBEv1.0.2", 12);
    permute();
    meta_ad((const uint8_t *)label.data(), label.size(), false);
  }
  void meta_ad(const uint8_t *d, size_t n, bool more) {
    begin_op(FLAG_M | FLAG_A, more);
    absorb(d, n);
  }
  void ad(const uint8_t *d, size_t n, bool more) {
    begin_op(FLAG_A, more);
    absorb(d, n);
  }
  void prf(uint8_t *out, size_t n, bool more) {
    begin_op(FLAG_I | FLAG_A | FLAG_C, more);
    for (size_t i = 0; i < n; i++) {
      out[i] = st_[pos_];
      st_[pos_] = 0;
      if (++pos_ == R) run_f();
    }
  }

 private:
  static constexpr int R = 166;
  enum { FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32 };
  uint8_t st_[200];
  uint8_t pos_ = 0, pos_begin_ = 0, cur_flags_ = 0;

  void permute() {
    uint64_t A[25];
    for (int i = 0; i < 25; i++) {
      A[i] = 0;
      for (int k = 7; k >= 0; k--) A[i] = (A[i] << 8) | st_[8 * i + k];
    }
    keccak_f1600(A);
    for (int i = 0; i < 25; i++)
      for (int k = 0; k < 8; k++) st_[8 * i + k] = (uint8_t)(A[i] >> (8 * k));
  }
  void run_f() {
    st_[pos_] ^= pos_begin_;
    st_[pos_ + 1] ^= 0x04;
    st_[R + 1] ^= 0x80;
    permute();
    pos_ = 0;
    pos_begin_ = 0;
  }
  void absorb(const uint8_t *d, size_t n) {
    for (size_t i = 0; i < n; i++) {
      st_[pos_] ^= d[i];
      if (++pos_ == R) run_f();
    }
  }
  void begin_op(uint8_t flags, bool more) {
    if (more) return;
    uint8_t old_begin = pos_begin_;
    pos_begin_ = pos_ + 1;
    cur_flags_ = flags;
    uint8_t hdr[2] = {old_begin, flags};
    absorb(hdr, 2);
    if ((flags & (FLAG_C | FLAG_K)) && pos_ != 0) run_f();
  }
};

// merlin::Transcript
class Transcript {
 public:
  explicit Transcript(const std::string &label) : strobe_("Merlin v1.0") { append_message("dom-sep", label); }
  void append_message(const std::string &label, const uint8_t *m, size_t n) {
    strobe_.meta_ad((const uint8_t *)label.data(), label.size(), false);
    uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
    strobe_.meta_ad(len, 4, true);
    strobe_.ad(m, n, false);
  }
  void append_message(const std::string &label, const std::string &m) {
    append_message(label, (const uint8_t *)m.data(), m.size());
  }
  void append_u64(const std::string &label, uint64_t x) {
    uint8_t b[8];
    for (int i = 0; i < 8; i++) b[i] = (uint8_t)(x >> (8 * i));
    append_message(label, b, 8);
  }
  void challenge_bytes(const std::string &label, uint8_t *out, size_t n) {
    strobe_.meta_ad((const uint8_t *)label.data(), label.size(), false);
    uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
    strobe_.meta_ad(len, 4, true);
    strobe_.prf(out, n, false);
  }

 private:
  Strobe128 strobe_;
};

// SHAKE256 XOF (FIPS 202): absorb everything, then squeeze
inline std::vector<uint8_t> shake256(const std::vector<uint8_t> &in, size_t out_len) {
  const size_t rate = 136;
  uint64_t A[25] = {0};
  std::vector<uint8_t> m(in);
  m.push_back(0x1F);
  while (m.size() % rate) m.push_back(0);
  m.back() |= 0x80;
  for (size_t off = 0; off < m.size(); off += rate) {
    for (size_t i = 0; i < rate; i++) A[i / 8] ^= (uint64_t)m[off + i] << (8 * (i % 8));
    keccak_f1600(A);
  }
  std::vector<uint8_t> out(out_len);
  size_t pos = 0;
  while (pos < out_len) {
    for (size_t i = 0; i < rate && pos < out_len; i++) out[pos++] = (uint8_t)(A[i / 8] >> (8 * (i % 8)));
    if (pos < out_len) keccak_f1600(A);
  }
  return out;
}

}  // namespace sph
