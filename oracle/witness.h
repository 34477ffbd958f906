/* ORACLE (test infrastructure, NOT product code): see witness.c */
#ifndef SPG_ORACLE_WITNESS_H
#define SPG_ORACLE_WITNESS_H
#include "fq.h"
#ifdef __cplusplus
extern "C" {
#endif
void owit_perm_fill(ofq *w3, size_t width, const size_t *seg_len, size_t n_seg, size_t v_col, size_t x_col,
                    size_t pi_col, size_t d_col);
void owit_perm_w0(const ofq *tau, const ofq *r, size_t used, size_t total, ofq *out);
void owit_exec(const ofq *inputs, size_t rows, size_t in_width, const ofq *w0, const ofq *tau, size_t n, size_t num_ios,
               ofq *w2, ofq *w3);
void owit_block(const ofq *vars, size_t rows, size_t vars_width, const ofq *w0, const ofq *tau, const ofq *r, size_t n,
                size_t io_width, size_t phy_ops, size_t vir_ops, size_t w2_width, ofq *w2, ofq *w3);
void owit_mem(const ofq *mems, size_t rows, size_t in_width, const ofq *tau, const ofq *r, size_t mem_width, ofq *w2,
              ofq *w3);
void owit_shift(const ofq *w3, size_t rows, size_t width, ofq *out);
#ifdef __cplusplus
}
#endif
#endif
