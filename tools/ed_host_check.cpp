// CPU harness for spartan_parallel_b200/csrc/ed25519.cuh (the same source the CUDA
// kernels compile): reads commands on stdin, prints hex results. Driven by
// tests/test_ed_host.py against the oracle.
//   mul  <a hex32> <b hex32>        -> canonical bytes of a*b mod p
//   sub  <a> <b>                    -> a-b
//   smul <scalar hex32 LE> <point compressed hex32>  -> compressed k*P   (or "invalid")
//   add  <p> <q>                    -> compressed p+q
//   dec  <p>                        -> "ok" / "invalid"
//   uni  <lo hex32> <hi hex32>      -> compressed from_uniform_bytes(lo || hi)
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>

#include "../spartan_parallel_b200/csrc/ed25519.cuh"
using namespace spg;

static void unhex(const std::string &s, uint8_t out[32]) {
  for (int i = 0; i < 32; i++) out[i] = (uint8_t)std::stoi(s.substr(2 * i, 2), nullptr, 16);
}
static std::string hex(const uint8_t b[32]) {
  char buf[65];
  for (int i = 0; i < 32; i++) snprintf(buf + 2 * i, 3, "%02x", b[i]);
  return std::string(buf, 64);
}

int main() {
  std::string op, a, b;
  while (std::cin >> op >> a) {
    uint8_t x[32], y[32], o[32];
    unhex(a, x);
    if (op == "uni") {  // 64 uniform bytes given as two hex32 words
      std::cin >> b;
      uint8_t u[64];
      unhex(a, u);
      unhex(b, u + 32);
      ge p = ristretto_from_uniform_bytes(u);
      ristretto_compress(p, o);
      std::cout << hex(o) << "\n";
      continue;
    }
    if (op == "dec") {
      ge p;
      std::cout << (ristretto_decompress(x, &p) ? "ok" : "invalid") << "\n";
      continue;
    }
    std::cin >> b;
    unhex(b, y);
    if (op == "mul" || op == "sub") {
      fe f = fe_frombytes(x), g = fe_frombytes(y);
      fe r = op == "mul" ? fe_mul(f, g) : fe_sub(f, g);
      fe_tobytes(r, o);
      std::cout << hex(o) << "\n";
    } else if (op == "smul") {
      ge p;
      if (!ristretto_decompress(y, &p)) {
        std::cout << "invalid\n";
        continue;
      }
      ge_cached pc = ge_to_cached(p);
      ge acc = ge_identity();
      for (int bit = 255; bit >= 0; bit--) {
        acc = ge_double(acc);
        if ((x[bit >> 3] >> (bit & 7)) & 1) acc = ge_add(acc, pc);
      }
      ristretto_compress(acc, o);
      std::cout << hex(o) << "\n";
    } else if (op == "add") {
      ge p, q;
      if (!ristretto_decompress(x, &p) || !ristretto_decompress(y, &q)) {
        std::cout << "invalid\n";
        continue;
      }
      ge r = ge_add(p, ge_to_cached(q));
      ristretto_compress(r, o);
      std::cout << hex(o) << "\n";
    }
  }
  return 0;
}
