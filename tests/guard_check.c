/* CPU check of the guarded host mirrors (zpic_b200/csrc/host/common/zb_guard.c), built and run by
 * tests/test_guard_cpu.py: a "device" array stands in for device memory; reads of a stale mirror must be served by
 * exactly one fill, the first write to a clean mirror must be reported, foreign faults must reach the old handler. */
#include <setjmp.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "zb_guard.h"

#define N 100000
static float device[N];
static float* mirror;
static int fills = 0, dirties = 0;
static sigjmp_buf jb;

static void fill( void* owner, int kind ) {
	(void) kind;
	zb_guard_set(mirror, ZB_G_RW);
	memcpy(mirror, owner, N * sizeof(float));
	zb_guard_set(mirror, ZB_G_READ);
	fills++;
}
static void dirty( void* owner, int kind ) { (void) owner; (void) kind; dirties++; }
static void foreign( int sig ) { (void) sig; siglongjmp(jb, 1); }

#define CHECK(c) do { if (!(c)) { printf("FAILED line %d: %s\n", __LINE__, #c); return 1; } } while (0)

int main( void ) {
	signal(SIGSEGV, foreign);                      /* somebody else's handler, installed first */
	CHECK(zb_guard_enabled());
	mirror = zb_guard_alloc(N * sizeof(float));
	CHECK(mirror && zb_guard_state(mirror) == ZB_G_RW);
	for (int i = 0; i < N; i++) CHECK(mirror[i] == 0.0f);
	zb_guard_bind(mirror, device, 7, fill, dirty);

	/* the device advances: the mirror is stale */
	for (int i = 0; i < N; i++) device[i] = 0.5f * i;
	zb_guard_set(mirror, ZB_G_NONE);
	volatile float* m = mirror;
	float s = 0;
	for (int i = 0; i < N; i += 1000) s += m[i];   /* first read faults, the rest are free */
	CHECK(fills == 1 && dirties == 0);
	CHECK(m[N - 1] == 0.5f * (N - 1));
	CHECK(zb_guard_state(mirror) == ZB_G_READ);

	/* a write to the clean mirror is noticed, once */
	m[17] = -3.0f; m[18] = -4.0f;
	CHECK(dirties == 1 && fills == 1 && zb_guard_state(mirror) == ZB_G_RW);
	CHECK(m[17] == -3.0f && m[16] == 8.0f);

	/* a WRITE to a stale mirror: fill, then dirty */
	for (int i = 0; i < N; i++) device[i] = 2.0f * i;
	zb_guard_set(mirror, ZB_G_NONE);
	m[5] += 1.0f;
	CHECK(fills == 2 && dirties == 2 && m[5] == 11.0f && m[6] == 12.0f);

	/* growing keeps the owner and the requested part of the contents */
	float* g = zb_guard_realloc(mirror, 4 * N * sizeof(float), 10 * sizeof(float));
	CHECK(g != mirror && g[5] == 11.0f && g[9] == 18.0f && g[10] == 0.0f);
	mirror = g; m = g;
	zb_guard_set(mirror, ZB_G_NONE);
	CHECK(m[3] == 6.0f && fills == 3);

	/* a fault that is not ours goes to the handler that was there before */
	if (sigsetjmp(jb, 1) == 0) {
		volatile int* bad = (volatile int*) 8;
		*bad = 1;
		CHECK(0);
	}
	zb_guard_free(mirror);
	CHECK(zb_guard_state(mirror) == -1);
	printf("guard ok: %lu fills, %lu dirties\n", zb_guard_fills(), zb_guard_dirties());
	return 0;
}
