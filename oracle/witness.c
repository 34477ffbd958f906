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

/* ------------------------------------------------------------------------------------------------
 * The witness sections SNARK::prove derives from the primary ones before committing them
 * (SURVEY 8(f)2). Restated loop for loop; the tables are row-major like the flattened
 * `w2_list_p` / `w3_list_p` the reference commits (lib.rs:1424, :1443, :1630, :1655).
 * Parity pin: none in the reference (it records no witness values); checked by the relations the
 * derived columns must satisfy (tests/test_oracle_witness.py).
 */

/* perm_w0 = (tau, r, r^2, ..., r^(used-1), 0 ...), lib.rs:1328-1338 (used = 2 * num_inputs_unpadded) */
void owit_perm_w0(const ofq *tau, const ofq *r, size_t used, size_t total, ofq *out) {
  ofq r_tmp = *r;
  for (size_t i = 0; i < total; i++) out[i] = ofq_zero();
  if (used > 0) out[0] = *tau;
  for (size_t i = 1; i < used; i++) {
    out[i] = r_tmp;
    r_tmp = ofq_mul(&r_tmp, r);
  }
}

/* INPUT part shared by perm_exec_w2 (lib.rs:1346-1375) and block_w2 (lib.rs:1515-1531):
 * in = the row of exec_inputs / block_vars, n = num_inputs_unpadded; w2 row zeroed by the caller */
static void owit_input_part(const ofq *in, const ofq *w0, size_t n, ofq *w2) {
  const ofq one = ofq_one();
  w2[0] = in[0];
  w2[1] = in[0];
  for (size_t i = 1; i < 2 * (n - 1); i++) {
    ofq t = ofq_mul(&w0[i], &in[i + 2]);
    w2[2 + i] = ofq_add(&w2[2 + i], &t);
  }
  for (size_t i = 0; i + 1 < n; i++) {
    ofq perm = i == 0 ? one : w0[i];
    ofq t = ofq_mul(&perm, &in[2 + i]);
    w2[0] = ofq_add(&w2[0], &t);
    t = ofq_mul(&perm, &in[2 + (n - 1) + i]);
    w2[2] = ofq_add(&w2[2], &t);
  }
  w2[0] = ofq_mul(&w2[0], &in[0]);
  ofq ZO = w2[2];
  w2[1] = ofq_add(&w2[1], &ZO);
  w2[1] = ofq_mul(&w2[1], &in[0]);
}

/* x = v * (tau - sum(w2[3..sum_end]) - in[2])  (lib.rs:1384-1388, :1533-1537) */
static ofq owit_input_x(const ofq *in, const ofq *w2, size_t sum_end, const ofq *tau) {
  ofq s = ofq_zero();
  for (size_t i = 3; i < sum_end; i++) s = ofq_add(&s, &w2[i]);
  ofq t = ofq_sub(tau, &s);
  t = ofq_sub(&t, &in[2]);
  return ofq_mul(&in[0], &t);
}

/* (pi, D) step of one pair of columns, last row of the segment first (lib.rs:1389-1398) */
static void owit_pair(ofq *row, const ofq *nxt /* NULL for the last row */, const ofq *x, size_t pi_col, size_t d_col) {
  const ofq one = ofq_one();
  if (nxt) {
    ofq t = ofq_add(&nxt[pi_col], &one);
    t = ofq_sub(&t, &nxt[0]);
    row[d_col] = ofq_mul(x, &t);
  } else {
    row[d_col] = *x;
  }
  row[pi_col] = ofq_mul(&row[0], &row[d_col]);
}

/* perm_exec_w2 / perm_exec_w3, lib.rs:1346-1400. inputs: rows x in_width, w2: rows x num_ios, w3: rows x 8 */
void owit_exec(const ofq *inputs, size_t rows, size_t in_width, const ofq *w0, const ofq *tau, size_t n, size_t num_ios,
               ofq *w2, ofq *w3) {
  for (size_t q = rows; q-- > 0;) {
    const ofq *in = inputs + q * in_width;
    ofq *r2 = w2 + q * num_ios, *r3 = w3 + q * 8;
    for (size_t i = 0; i < num_ios; i++) r2[i] = ofq_zero();
    for (size_t i = 0; i < 8; i++) r3[i] = ofq_zero();
    owit_input_part(in, w0, n, r2);
    r3[0] = in[0];
    r3[1] = owit_input_x(in, r2, num_ios, tau);
    r3[4] = r2[0];
    r3[5] = r2[1];
    owit_pair(r3, q + 1 < rows ? r3 + 8 : NULL, &r3[1], 2, 3);
  }
}

/* block_w2 / block_w3 of ONE instance, lib.rs:1511-1613. vars: rows x vars_width (inputs first, then the
 * memory operations from io_width on), w2: rows x w2_width, w3: rows x 8 */
void owit_block(const ofq *vars, size_t rows, size_t vars_width, const ofq *w0, const ofq *tau, const ofq *r, size_t n,
                size_t io_width, size_t phy_ops, size_t vir_ops, size_t w2_width, ofq *w2, ofq *w3) {
  ofq r2s = ofq_mul(r, r), r3s = ofq_mul(&r2s, r);
  for (size_t q = rows; q-- > 0;) {
    const ofq *in = vars + q * vars_width;
    ofq *a = w2 + q * w2_width, *b = w3 + q * 8;
    const ofq *nxt = q + 1 < rows ? b + 8 : NULL;
    ofq cnst = in[0];
    for (size_t i = 0; i < w2_width; i++) a[i] = ofq_zero();
    for (size_t i = 0; i < 8; i++) b[i] = ofq_zero();
    owit_input_part(in, w0, n, a);
    b[0] = in[0];
    b[1] = owit_input_x(in, a, w2_width, tau); /* the memory entries of the row are still zero here */
    owit_pair(b, nxt, &b[1], 2, 3);
    /* PHY: PMR = r * PD, PMC = (cnst or PMC[i-1]) * (tau - PA - PMR), lib.rs:1541-1553 */
    for (size_t i = 0; i < phy_ops; i++) {
      size_t pmr = 2 * n + 2 * i, pmc = pmr + 1;
      a[pmr] = ofq_mul(r, &in[io_width + 2 * i + 1]);
      ofq t = i == 0 ? cnst : a[pmc - 2];
      ofq u = ofq_sub(tau, &in[io_width + 2 * i]);
      u = ofq_sub(&u, &a[pmr]);
      a[pmc] = ofq_mul(&t, &u);
    }
    ofq px = phy_ops == 0 ? cnst : a[2 * n + 2 * (phy_ops - 1) + 1];
    owit_pair(b, nxt, &px, 4, 5);
    /* VIR: VMR1..3 = r, r^2, r^3 times (VD, VL, VT), VMC chain, lib.rs:1570-1597 */
    for (size_t i = 0; i < vir_ops; i++) {
      size_t src = io_width + 2 * phy_ops + 4 * i, dst = 2 * n + 2 * phy_ops + 4 * i;
      a[dst] = ofq_mul(r, &in[src + 1]);
      a[dst + 1] = ofq_mul(&r2s, &in[src + 2]);
      a[dst + 2] = ofq_mul(&r3s, &in[src + 3]);
      ofq t = i == 0 ? cnst : a[dst - 1];
      ofq u = ofq_sub(tau, &in[src]);
      u = ofq_sub(&u, &a[dst]);
      u = ofq_sub(&u, &a[dst + 1]);
      u = ofq_sub(&u, &a[dst + 2]);
      a[dst + 3] = ofq_mul(&t, &u);
    }
    ofq vx = vir_ops == 0 ? cnst : a[2 * n + 2 * phy_ops + 4 * (vir_ops - 1) + 3];
    owit_pair(b, nxt, &vx, 6, 7);
  }
}

/* mem_gen, lib.rs:832-880. mems: rows x in_width rows (v, _, addr, data, ...), w2: rows x mem_width, w3: rows x 8 */
void owit_mem(const ofq *mems, size_t rows, size_t in_width, const ofq *tau, const ofq *r, size_t mem_width, ofq *w2,
              ofq *w3) {
  for (size_t q = rows; q-- > 0;) {
    const ofq *m = mems + q * in_width;
    ofq *a = w2 + q * mem_width, *b = w3 + q * 8;
    for (size_t i = 0; i < mem_width; i++) a[i] = ofq_zero();
    for (size_t i = 0; i < 8; i++) b[i] = ofq_zero();
    a[3] = ofq_mul(r, &m[3]);
    b[0] = m[0];
    ofq t = ofq_sub(tau, &m[2]);
    t = ofq_sub(&t, &a[3]);
    b[1] = ofq_mul(&m[0], &t);
    owit_pair(b, q + 1 < rows ? b + 8 : NULL, &b[1], 2, 3);
    t = ofq_add(&m[0], &m[2]);
    t = ofq_add(&t, &a[3]);
    b[4] = ofq_mul(&m[0], &t);
    b[5] = m[0];
  }
}

/* w3_shifted: rows 1.. of the instance followed by a zero row (lib.rs:1667-1676, :925-929) */
void owit_shift(const ofq *w3, size_t rows, size_t width, ofq *out) {
  for (size_t q = 0; q + 1 < rows; q++)
    for (size_t i = 0; i < width; i++) out[q * width + i] = w3[(q + 1) * width + i];
  for (size_t i = 0; i < width; i++) out[(rows - 1) * width + i] = ofq_zero();
}
