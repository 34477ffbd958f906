/*
 * ORACLE (test infrastructure, NOT product code).
 *
 * CPU restatement of the (pi, D) recurrence of the w3 witness sections SNARK::prove builds
 * (/root/reference/src/lib.rs:1378-1400 perm_exec_w3, :862-880 mem_gen, :1533-1570 block_w3):
 * the same sequential loop, last proof first, one segment per proving instance.
 * Parity pin: the reference records no witness values; pinned by the closed form checked in
 * tests/test_oracle_witness.py (pi[q] = product of x over the valid suffix).
 */
#include "witness.h"

/* w3: rows of `width` scalars; columns v_col, x_col given, pi_col and d_col written */
void owit_perm_fill(ofq *w3, size_t width, const size_t *seg_len, size_t n_seg, size_t v_col, size_t x_col,
                    size_t pi_col, size_t d_col) {
  const ofq one = ofq_one();
  size_t start = 0;
  for (size_t s = 0; s < n_seg; s++) {
    size_t n = seg_len[s];
    for (size_t k = n; k-- > 0;) { /* for q in (0..n).rev(), lib.rs:1382 */
      ofq *row = w3 + (start + k) * width;
      if (k != n - 1) {
        const ofq *nxt = row + width;
        /* w3[q][3] = w3[q][1] * (w3[q+1][2] + ONE - w3[q+1][0]), lib.rs:1391-1393 */
        ofq t = ofq_add(&nxt[pi_col], &one);
        t = ofq_sub(&t, &nxt[v_col]);
        row[d_col] = ofq_mul(&row[x_col], &t);
      } else {
        row[d_col] = row[x_col]; /* lib.rs:1395 */
      }
      row[pi_col] = ofq_mul(&row[v_col], &row[d_col]); /* lib.rs:1398 */
    }
    start += n;
  }
}
