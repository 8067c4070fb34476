/* zpic-b200 :: em1d electric current density (reference em1d/current.h) */
#ifndef ZPIC_B200_EM1D_CURRENT_H
#define ZPIC_B200_EM1D_CURRENT_H

#include "zpic.h"

enum smooth_type { NONE, BINOMIAL, COMPENSATED };
/* reference em1d/current.h:29-32 */
enum current_boundary { CURRENT_BC_NONE, CURRENT_BC_PERIODIC };

typedef struct Smooth {
	enum smooth_type xtype;
	int xlevel;
} t_smooth;

/* J grid: nx+3 float3, guards {1 lower, 2 upper} (reference em1d/current.h:47-75) */
typedef struct Current {
	float3 *J;
	float3 *J_buf;
	int nx;
	int gc[2];
	float box;
	float dx;
	t_smooth smooth;
	float dt;
	int iter;
	enum current_boundary bc_type;
} t_current;

/* replaces em1d/current.c:31-73 */
void current_new( t_current *current, int nx, float box, float dt );
/* replaces em1d/current.c:80-86 */
void current_delete( t_current *current );
/* replaces em1d/current.c:93-101 */
void current_zero( t_current *current );
/* device: periodic guard fold + binomial / compensated filter (reference em1d/current.c:112-155, 265-333) */
void current_update( t_current *current );
/* replaces em1d/current.c:167-230 */
void current_report( const t_current *current, const int jc );

#endif
