/* ORACLE (test infrastructure, NOT product code): see witness.c */
#ifndef SPG_ORACLE_WITNESS_H
#define SPG_ORACLE_WITNESS_H
#include "fq.h"
#ifdef __cplusplus
extern "C" {
#endif
void owit_perm_fill(ofq *w3, size_t width, const size_t *seg_len, size_t n_seg, size_t v_col, size_t x_col,
                    size_t pi_col, size_t d_col);
#ifdef __cplusplus
}
#endif
#endif
