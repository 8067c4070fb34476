/* zpic-b200 :: wall-clock timers (reference em2d/timer.h) */
#ifndef ZPIC_B200_TIMER_H
#define ZPIC_B200_TIMER_H
#include <stdint.h>
uint64_t timer_ticks( void );
double timer_interval_seconds( uint64_t start, uint64_t end );
double timer_cpu_seconds( void );
double timer_resolution( void );
#endif
