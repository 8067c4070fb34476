/* zpic-b200 :: wall-clock timers (reference em2d/timer.h) */
#ifndef ZPIC_B200_TIMER_H
#define ZPIC_B200_TIMER_H
#include <stdint.h>
/* replaces em2d/timer.c:24-29 */
uint64_t timer_ticks( void );
/* replaces em2d/timer.c:38-41 */
double timer_interval_seconds( uint64_t start, uint64_t end );
/* replaces em2d/timer.c:51-59 */
double timer_cpu_seconds( void );
/* replaces em2d/timer.c:72-84 */
double timer_resolution( void );
#endif
