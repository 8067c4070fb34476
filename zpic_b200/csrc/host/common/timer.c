/* zpic-b200 :: wall-clock timers, microsecond ticks from gettimeofday like the
 * reference (em2d/timer.c:24-84).  Device work is timed with CUDA events
 * (zdev_event_*); these only serve the reference API (sim_timings). */
#include <stddef.h>
#include <sys/time.h>
#include "timer.h"

uint64_t timer_ticks( void )
{
	struct timeval now;
	gettimeofday(&now, NULL);
	return (uint64_t) now.tv_sec * 1000000u + (uint64_t) now.tv_usec;
}

double timer_interval_seconds( uint64_t start, uint64_t end ) { return (end - start) * 1.0e-6; }

double timer_cpu_seconds( void )
{
	struct timeval now;
	gettimeofday(&now, NULL);
	return (double) now.tv_sec + 1.0e-6 * (double) now.tv_usec;
}

double timer_resolution( void )
{
	uint64_t a = timer_ticks(), b;
	do { b = timer_ticks(); } while (b == a);
	return (b - a) * 1.0e-6;
}
