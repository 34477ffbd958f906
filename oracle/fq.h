/*
 * ORACLE (test infrastructure, NOT product code).
 *
 * CPU restatement of the reference's Curve25519 scalar field
 * (/root/reference/src/scalar/ristretto255.rs). Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * Representation is the reference's: four 64-bit little-endian limbs holding
 * a*R mod q (R = 2^256), always fully reduced (ristretto255.rs:193-199).
 *
 * Parity pin: every known-answer test of ristretto255.rs:776-1201 is replayed in
 * tests/test_oracle_field.py, plus libsodium (PyNaCl) as an independent F_q.
 */
#ifndef SPG_ORACLE_FQ_H
#define SPG_ORACLE_FQ_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t l[4]; } ofq;

/* constants, ristretto255.rs:248-253, 304-328 */
extern const ofq OFQ_MODULUS, OFQ_R, OFQ_R2, OFQ_R3;
#define OFQ_INV 0xd2b51da312547e1bULL

ofq ofq_zero(void);
ofq ofq_one(void);                       /* = R, ristretto255.rs:370-372 */
ofq ofq_add(const ofq *a, const ofq *b); /* :736-745 */
ofq ofq_sub(const ofq *a, const ofq *b); /* :718-732 */
ofq ofq_neg(const ofq *a);               /* :749-763 */
ofq ofq_mul(const ofq *a, const ofq *b); /* :690-714 */
ofq ofq_square(const ofq *a);            /* :476-504 */
ofq ofq_montgomery_reduce(const uint64_t r[8]); /* :642-686 */
ofq ofq_from_u64(uint64_t v);            /* :212-216 */
ofq ofq_from_raw(const uint64_t v[4]);   /* :470-472 */
ofq ofq_from_u512(const uint64_t v[8]);  /* :448-466 */
int ofq_from_bytes(const uint8_t b[32], ofq *out); /* :391-415, returns 1 if canonical */
void ofq_to_bytes(const ofq *a, uint8_t out[32]);  /* :419-431 */
ofq ofq_from_bytes_wide(const uint8_t b[64]);      /* :435-446 */
ofq ofq_pow(const ofq *a, const uint64_t by[4]);   /* :508-519 */
ofq ofq_invert(const ofq *a);            /* addition chain, :541-595 (0 -> 0) */
ofq ofq_batch_invert(ofq *inputs, size_t n); /* :597-639 */
int ofq_eq(const ofq *a, const ofq *b);

/* vector helpers used by the python side */
void ofq_vec_mul(const ofq *a, const ofq *b, ofq *out, size_t n);
void ofq_vec_add(const ofq *a, const ofq *b, ofq *out, size_t n);
void ofq_vec_sub(const ofq *a, const ofq *b, ofq *out, size_t n);
void ofq_vec_from_u512(const uint64_t *wide, ofq *out, size_t n);

#ifdef __cplusplus
}
#endif
#endif
