/* zpic-b200 :: access to the state of the host random stream (internal; random.h is the reference's API) */
#ifndef ZB_RAND_H
#define ZB_RAND_H
#include <stdint.h>
/* (z, w) = the reference's m_z, m_w; (have, value) = iset, gset of rand_norm (em2d/random.c:16-17, 69-70) */
void zb_rand_get_state( uint32_t* z, uint32_t* w, int* have, double* value );
void zb_rand_set_state( uint32_t z, uint32_t w, int have, double value );
#endif
