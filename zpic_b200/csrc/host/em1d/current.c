/* zpic-b200 :: em1d current density, host side of the API (reference em1d/current.c) */
#include <stdlib.h>
#include <stdio.h>
#include "zb_state.h"
#include "zdf.h"

void current_new( t_current *current, int nx, float box, float dt )
{
	zb_grid_drop_cur(current);
	current->nx = nx;
	current->gc[0] = 1; current->gc[1] = 2;           /* reference current.c:33 */
	current->J_buf = zb_guard_alloc(((size_t) nx + 3) * sizeof(float3));
	if (!current->J_buf) { fprintf(stderr, "(*error*) current_new: out of memory\n"); exit(-1); }
	current->J = current->J_buf + 1;
	current->box = box;
	current->dx = box / nx;
	current->smooth = (t_smooth) { .xtype = NONE, .xlevel = 0 };
	current->iter = 0;
	current->dt = dt;
	current->bc_type = CURRENT_BC_PERIODIC;
	zb_grid_of_cur(current, 1);
	zb_guard_bind_cur(current);
}

void current_delete( t_current *current )
{
	zb_grid_drop_cur(current);
	zb_guard_free(current->J_buf);
	current->J_buf = NULL;
}

void current_zero( t_current *current )
{
	zb_grid* e = zb_grid_of_cur(current, 1);
	zdev_current1d_zero(zb_dev(e));
	e->j_host_stale = 1;
	zb_guard_refresh();
}

void current_update( t_current *current )
{
	zb_grid* e = zb_grid_of_cur(current, 1);
	zdev_current1d_update(zb_dev(e), current->bc_type == CURRENT_BC_PERIODIC,
	                      (int) current->smooth.xtype, current->smooth.xlevel);
	e->j_host_stale = 1;
	current->iter++;
	zb_guard_refresh();
}

void current_report( const t_current *current, const int jc )
{
	if (jc < 0 || jc > 2) {
		fprintf(stderr, "(*error*) Invalid current component (jc) selected, returning\n");
		return;
	}
	zb_cur_to_host(current);
	float* buf = malloc((size_t) current->nx * sizeof(float));
	const float* f = (const float*) current->J;
	for (int i = 0; i < current->nx; i++) buf[i] = f[3*i + jc];

	char name[8], label[8];
	snprintf(name, sizeof name, "J%1d", jc);
	snprintf(label, sizeof label, "J_%c", "xyz"[jc]);
	t_zdf_grid_axis axis[1] = { { .min = 0.0, .max = current->box, .name = "x", .label = "x", .units = "c/\\omega_p" } };
	t_zdf_grid_info info = { .ndims = 1, .name = name, .label = label, .units = "e \\omega_p^2 / c", .axis = axis };
	info.count[0] = current->nx;
	t_zdf_iteration iter = { .name = "ITERATION", .n = current->iter, .t = current->iter * current->dt, .time_units = "1/\\omega_p" };
	zdf_save_grid(buf, zdf_float32, &info, &iter, "CURRENT");
	free(buf);
	zb_guard_refresh();
}
