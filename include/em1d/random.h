/* zpic-b200 :: host pseudo-random numbers (reference em2d/random.h).
 * Kept on the host and bit-exact: initial momenta and moving-window injection draw
 * from this single global stream. */
#ifndef ZPIC_B200_RANDOM_H
#define ZPIC_B200_RANDOM_H
#include <stdint.h>
/* NB the reference definition assigns the FIRST argument to m_w (random.c:25-29) */
void set_rand_seed( uint32_t m_z_, uint32_t m_w_ );
/* replaces em1d/random.c:48-53 */
uint32_t rand_uint32( void );
/* replaces em1d/random.c:67-101 */
double rand_norm( void );
#endif
