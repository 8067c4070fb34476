/* zpic-b200 :: em2d electric current density (reference em2d/current.h) */
#ifndef ZPIC_B200_EM2D_CURRENT_H
#define ZPIC_B200_EM2D_CURRENT_H

#include "zpic.h"

/* digital filter kinds (reference current.h:18-22) */
enum smooth_type { NONE, BINOMIAL, COMPENSATED };

/* filter configuration (reference current.h:29-34) */
typedef struct Smooth {
	enum smooth_type xtype, ytype;
	int xlevel, ylevel;
} t_smooth;

/* J grid: (nx0+3) x (nx1+3) float3, guards {1 lower, 2 upper} per axis,
 * J points at cell (0,0) of J_buf (reference current.h:41-70, current.c:30-53).
 * On the device the same layout is used so host mirrors are plain copies. */
typedef struct Current {
	float3 *J;
	float3 *J_buf;
	int nx[2];
	int nrow;
	int gc[2][2];
	float box[2];
	float dx[2];
	t_smooth smooth;
	float dt;
	int iter;
	int moving_window;
} t_current;

/* replaces em2d/current.c:30-79 */
void current_new( t_current *current, int nx[], float box[], float dt );
/* replaces em2d/current.c:86-91 */
void current_delete( t_current *current );
/* device: cudaMemsetAsync of the whole J buffer (reference current.c:98-107) */
void current_zero( t_current *current );
/* device: guard fold + smoothing kernels (reference current.c:118-183, 297-459) */
void current_update( t_current *current );
/* host, after mirror sync (reference current.c:194-283) */
void current_report( const t_current *current, const int jc );

#endif
