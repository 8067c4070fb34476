/* zpic-b200 :: em2d current density, host side of the API (reference em2d/current.c).
 * The host J_buf is only a mirror for diagnostics; zero / fold / smooth run on the
 * device (csrc/dev/zdev_grid2d.cu). */
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include "zb_state.h"
#include "zdf.h"

void current_new( t_current *current, int nx[], float box[], float dt )
{
	zb_grid_drop_cur(current);

	/* guard cells for linear interpolation: 1 below, 2 above (reference current.c:33-34) */
	for (int i = 0; i < 2; i++) {
		current->nx[i] = nx[i];
		current->gc[i][0] = 1;
		current->gc[i][1] = 2;
		current->box[i] = box[i];
		current->dx[i] = box[i] / nx[i];
	}
	current->nrow = nx[0] + 3;
	size_t ncell = (size_t) (nx[0] + 3) * (nx[1] + 3);
	current->J_buf = zb_guard_alloc(ncell * sizeof(float3));
	if (!current->J_buf) { fprintf(stderr, "(*error*) current_new: out of memory\n"); exit(-1); }
	zb_guard_bind_cur(current);
	current->J = current->J_buf + 1 + current->nrow;

	current->smooth = (t_smooth) { .xtype = NONE, .ytype = NONE, .xlevel = 0, .ylevel = 0 };
	current->iter = 0;
	current->dt = dt;
	current->moving_window = 0;

	zb_grid_of_cur(current, 1);
}

void current_delete( t_current *current )
{
	zb_grid_drop_cur(current);
	if (zdev_ready()) zdev_host_unpin(current->J_buf);
	zb_guard_free(current->J_buf);
	current->J_buf = NULL;
}

void current_zero( t_current *current )
{
	zb_grid* e = zb_grid_of_cur(current, 1);
	zdev_current_zero(zb_dev(e));
	e->j_host_stale = 1;
	zb_guard_refresh();
}

void current_update( t_current *current )
{
	zb_grid* e = zb_grid_of_cur(current, 1);
	/* smoothing parameters and the window flag are re-read every step: decks and
	   Python may change them after sim_new ("Config / flags") */
	zdev_current_update(zb_dev(e), current->moving_window,
	                    (int) current->smooth.xtype, (int) current->smooth.ytype,
	                    current->smooth.xlevel, current->smooth.ylevel);
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
	if (zb_par_rank() != 0) { zb_guard_refresh(); return; }          /* one file per box: rank 0 writes it */

	const int nx = current->nx[0], ny = current->nx[1];
	float* buf = malloc((size_t) nx * ny * sizeof(float));
	for (int j = 0; j < ny; j++) {
		const float* row = (const float*) (current->J + (size_t) j * current->nrow);
		for (int i = 0; i < nx; i++) buf[(size_t) j * nx + i] = row[3*i + jc];
	}

	/* names as produced by the reference's truncating snprintf (current.c:239-241) */
	char name[8], label[8];
	snprintf(name, sizeof name, "J%1d", jc);
	snprintf(label, sizeof label, "J_%c", "xyz"[jc]);

	t_zdf_grid_axis axis[2] = {
		{ .min = 0.0, .max = current->box[0], .name = "x", .label = "x", .units = "c/\\omega_p" },
		{ .min = 0.0, .max = current->box[1], .name = "y", .label = "y", .units = "c/\\omega_p" }
	};
	t_zdf_grid_info info = { .ndims = 2, .name = name, .label = label,
	                         .units = "e \\omega_p^2 / c", .axis = axis };
	info.count[0] = nx; info.count[1] = ny;
	t_zdf_iteration iter = { .name = "ITERATION", .n = current->iter,
	                         .t = current->iter * current->dt, .time_units = "1/\\omega_p" };
	zdf_save_grid(buf, zdf_float32, &info, &iter, "CURRENT");
	free(buf);
	zb_guard_refresh();
}
