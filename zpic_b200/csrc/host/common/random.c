/* zpic-b200 :: host random numbers.
 * Same generator and same stream as the reference (em2d/random.c:16-101) because the
 * initial momenta and every moving-window injection are drawn from it and must match
 * the reference bit for bit: two 16-bit multiply-with-carry generators glued into a
 * 32-bit word, and polar Box-Muller in double precision with one cached deviate. */
#include <math.h>
#include "random.h"

static uint32_t mwc_w = 12345;   /* default seeds of the reference (random.c:16-17) */
static uint32_t mwc_z = 67890;
static int    have_spare = 0;
static double spare = 0.0;

void set_rand_seed( uint32_t first, uint32_t second )
{
	/* the reference assigns its first argument to m_w (random.c:25-29) */
	mwc_w = first;
	mwc_z = second;
}

/* the stream's state, for the device-side continuation of the stream (zdev_ref_normals, zpic_dev.h) */
void zb_rand_get_state( uint32_t* z, uint32_t* w, int* have, double* value )
{
	*z = mwc_z; *w = mwc_w; *have = have_spare; *value = spare;
}

void zb_rand_set_state( uint32_t z, uint32_t w, int have, double value )
{
	mwc_z = z; mwc_w = w; have_spare = have; spare = value;
}

uint32_t rand_uint32( void )
{
	mwc_z = 36969u * (mwc_z & 0xffffu) + (mwc_z >> 16);
	mwc_w = 18000u * (mwc_w & 0xffffu) + (mwc_w >> 16);
	return (mwc_z << 16) + mwc_w;
}

double rand_norm( void )
{
	if (have_spare) { have_spare = 0; return spare; }

	double a, b, r2;
	do {
		a = ( rand_uint32() + 0.5 ) / 2147483649.0 - 1.0;
		b = ( rand_uint32() + 0.5 ) / 2147483649.0 - 1.0;
		r2 = a*a + b*b;
	} while ( r2 == 0.0 || r2 >= 1.0 );

	double f = sqrt( -2.0 * log(r2) / r2 );
	spare = a * f;
	have_spare = 1;
	return b * f;
}
