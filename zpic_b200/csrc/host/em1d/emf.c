/* zpic-b200 :: em1d electromagnetic fields, host side of the API (reference em1d/emf.c).
 * Host: construction, plane-wave laser launch (libm double precision, emf.c:160-262), initial / external
 * field set-up, reports.  Device: field advance incl. the Mur open boundary, guards, window shift,
 * external-field superposition, energy (csrc/dev/zdev_grid1d.cu). */
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include "zb_state.h"
#include "zpic_b200.h"
#include "timer.h"
#include "zdf.h"

static double emf_seconds = 0.0;
double emf_time( void ) { return emf_seconds; }

void emf_new( t_emf *emf, int nx, float box, const float dt )
{
	zb_grid_drop_emf(emf);
	emf->nx = nx;
	emf->gc[0] = 1; emf->gc[1] = 2;                     /* reference emf.c:41 */
	emf->E_buf = zb_guard_alloc(((size_t) nx + 3) * sizeof(float3));       /* guarded mirrors: ../common/zb_guard.h */
	emf->B_buf = zb_guard_alloc(((size_t) nx + 3) * sizeof(float3));
	if (!emf->E_buf || !emf->B_buf) { fprintf(stderr, "(*error*) emf_new: out of memory\n"); exit(-1); }
	emf->E = emf->E_buf + 1;
	emf->B = emf->B_buf + 1;
	emf->box = box;
	emf->dx = box / nx;
	emf->dt = dt;
	emf->iter = 0;
	emf->moving_window = 0;
	emf->n_move = 0;
	emf->bc_type = EMF_BC_PERIODIC;
	memset(emf->mur_fld, 0, sizeof emf->mur_fld);
	memset(emf->mur_tmp, 0, sizeof emf->mur_tmp);
	memset(&emf->ext_fld, 0, sizeof emf->ext_fld);
	emf->E_part = emf->E;
	emf->B_part = emf->B;
	zb_grid* e = zb_grid_of_emf(emf, 1);
	e->eb_dev_stale = 0;
	zb_guard_bind_emf(emf);
}

void emf_delete( t_emf *emf )
{
	zb_grid_drop_emf(emf);
	zb_guard_free(emf->E_buf); zb_guard_free(emf->B_buf);
	emf->E_buf = emf->B_buf = NULL;
	if (emf->ext_fld.E_type > EMF_FLD_TYPE_NONE) zb_guard_free(emf->ext_fld.E_part_buf);
	if (emf->ext_fld.B_type > EMF_FLD_TYPE_NONE) zb_guard_free(emf->ext_fld.B_part_buf);
	emf->E_part = emf->B_part = NULL;
}

/* longitudinal sin^2 envelope (reference em1d/emf.c:160-183) */
static float envelope( const t_emf_laser* l, const float z )
{
	if (z > l->start) return 0.0;
	if (z > l->start - l->rise) {
		float csi = z - l->start;
		float e = sin( M_PI_2 * csi / l->rise );
		return e*e;
	}
	if (z > l->start - (l->rise + l->flat)) return 1.0;
	if (z > l->start - (l->rise + l->flat + l->fall)) {
		float csi = z - (l->start - l->rise - l->flat - l->fall);
		float e = sin( M_PI_2 * csi / l->fall );
		return e*e;
	}
	return 0.0;
}

void emf_add_laser( t_emf* const emf, t_emf_laser* laser )
{
	if (laser->fwhm != 0) {
		if (laser->fwhm <= 0) { fprintf(stderr, "Invalid laser FWHM, must be > 0, aborting.\n"); exit(-1); }
		laser->rise = laser->fwhm; laser->fall = laser->fwhm; laser->flat = 0.;
	}
	if (laser->rise <= 0) { fprintf(stderr, "Invalid laser RISE, must be > 0, aborting.\n"); exit(-1); }
	if (laser->flat < 0)  { fprintf(stderr, "Invalid laser FLAT, must be >= 0, aborting.\n"); exit(-1); }
	if (laser->fall <= 0) { fprintf(stderr, "Invalid laser FALL, must be > 0, aborting.\n"); exit(-1); }

	zb_emf_to_host(emf);
	float3* E = emf->E; float3* B = emf->B;
	const float dx = emf->dx;
	const float amp = laser->omega0 * laser->a0;
	const float cos_pol = cos( laser->polarization );
	const float sin_pol = sin( laser->polarization );
	const float k = laser->omega0;
	for (int i = 0; i < emf->nx; i++) {
		float z = i * dx, z_2 = z + dx/2;
		float lenv = amp * envelope(laser, z), lenv_2 = amp * envelope(laser, z_2);
		E[i].y += +lenv * cos( k * z ) * cos_pol;
		E[i].z += +lenv * cos( k * z ) * sin_pol;
		B[i].y += -lenv_2 * cos( k * z_2 ) * sin_pol;
		B[i].z += +lenv_2 * cos( k * z_2 ) * cos_pol;
	}
	/* guard refresh as the reference's emf_update_gc does it: lower guard, and only ONE upper guard
	   cell (its upper loop runs to gc[0]; em1d/emf.c:476-500) */
	if (emf->bc_type == EMF_BC_PERIODIC) {
		E[-1] = E[emf->nx - 1]; B[-1] = B[emf->nx - 1];
		E[emf->nx] = E[0];      B[emf->nx] = B[0];
	}
	zb_grid_of_emf(emf, 1)->eb_dev_stale = 1;
}

void emf_init_fld( t_emf* const emf, t_emf_init_fld* init_fld )
{
	if (emf->iter != 0) {
		fprintf(stderr, "emf_init_fld should only be called at initialization, aborting...\n");
		exit(-1);
	}
	zb_emf_to_host(emf);
	for (int f = 0; f < 2; f++) {
		float3* A = f ? emf->B : emf->E;
		enum emf_fld_type type = f ? init_fld->B_type : init_fld->E_type;
		float3 v0 = f ? init_fld->B_0 : init_fld->E_0;
		float3 (*fn)(int, float, void*) = f ? init_fld->B_custom : init_fld->E_custom;
		void* data = f ? init_fld->B_custom_data : init_fld->E_custom_data;
		if (type == EMF_FLD_TYPE_NONE) continue;
		for (int i = -emf->gc[0]; i < emf->nx + emf->gc[1]; i++)
			A[i] = (type == EMF_FLD_TYPE_UNIFORM) ? v0 : fn(i, emf->dx, data);
	}
	zb_grid_of_emf(emf, 1)->eb_dev_stale = 1;
}

static float3* eval_ext( const t_emf* emf, float3 (*fn)(int, float, void*), void* data )
{
	float3* buf = malloc(((size_t) emf->nx + 3) * sizeof(float3));
	for (int i = -emf->gc[0]; i < emf->nx + emf->gc[1]; i++) buf[i + 1] = fn(i, emf->dx, data);
	return buf;
}

void emf_set_ext_fld( t_emf* const emf, t_emf_ext_fld* ext_fld )
{
	zb_emf_to_device(emf);
	zb_grid* e = zb_grid_of_emf(emf, 1);
	const size_t bytes = ((size_t) emf->nx + 3) * sizeof(float3);
	if (emf->ext_fld.E_type > EMF_FLD_TYPE_NONE) zb_guard_free(emf->ext_fld.E_part_buf);
	if (emf->ext_fld.B_type > EMF_FLD_TYPE_NONE) zb_guard_free(emf->ext_fld.B_part_buf);
	if ((unsigned) ext_fld->E_type > EMF_FLD_TYPE_CUSTOM || (unsigned) ext_fld->B_type > EMF_FLD_TYPE_CUSTOM) {
		fprintf(stderr, "Invalid external field type, aborting.\n");
		exit(-1);
	}
	emf->ext_fld.E_type = ext_fld->E_type;
	emf->ext_fld.B_type = ext_fld->B_type;
	if (ext_fld->E_type == EMF_FLD_TYPE_NONE) { emf->E_part = emf->E; emf->ext_fld.E_part_buf = NULL; }
	else {
		emf->ext_fld.E_0 = ext_fld->E_0;
		emf->ext_fld.E_custom = ext_fld->E_custom; emf->ext_fld.E_custom_data = ext_fld->E_custom_data;
		emf->ext_fld.E_part_buf = zb_guard_alloc(bytes);
		emf->E_part = emf->ext_fld.E_part_buf + 1;
	}
	if (ext_fld->B_type == EMF_FLD_TYPE_NONE) { emf->B_part = emf->B; emf->ext_fld.B_part_buf = NULL; }
	else {
		emf->ext_fld.B_0 = ext_fld->B_0;
		emf->ext_fld.B_custom = ext_fld->B_custom; emf->ext_fld.B_custom_data = ext_fld->B_custom_data;
		emf->ext_fld.B_part_buf = zb_guard_alloc(bytes);
		emf->B_part = emf->ext_fld.B_part_buf + 1;
	}
	float e0[3] = { ext_fld->E_0.x, ext_fld->E_0.y, ext_fld->E_0.z };
	float b0[3] = { ext_fld->B_0.x, ext_fld->B_0.y, ext_fld->B_0.z };
	zdev_emf1d_set_ext_uniform(zb_dev(e), ext_fld->E_type == EMF_FLD_TYPE_UNIFORM, e0, ext_fld->B_type == EMF_FLD_TYPE_UNIFORM, b0);
	float3 *ge = NULL, *gb = NULL;
	if (ext_fld->E_type == EMF_FLD_TYPE_CUSTOM) ge = eval_ext(emf, ext_fld->E_custom, ext_fld->E_custom_data);
	if (ext_fld->B_type == EMF_FLD_TYPE_CUSTOM) gb = eval_ext(emf, ext_fld->B_custom, ext_fld->B_custom_data);
	if (ge || gb) zdev_emf1d_set_ext_grid(zb_dev(e), (const float*) ge, (const float*) gb);
	free(ge); free(gb);
	e->part_host_stale = 1;
	zb_guard_bind_emf(emf);
	zb_guard_refresh();
}

/* custom-field callback that reads a table the caller filled (zpic_b200.h; the em1d form: value of cell ix at
   table[3 * (ix + 1)]) */
float3 zpic_b200_table_field( int ix, float dx, void* data )
{
	const zpic_b200_field_table* t = (const zpic_b200_field_table*) data;
	const float* v = t->table + 3 * (size_t) (ix + 1);
	(void) dx;
	return (float3) { v[0], v[1], v[2] };
}

void emf_advance( t_emf *emf, const t_current *current )
{
	uint64_t t0 = timer_ticks();
	zb_emf_to_device(emf);
	zb_grid* e = zb_grid_of_emf(emf, 1);
	zb_grid* c = zb_grid_of_cur(current, 1);
	/* bc_type is re-read every step: decks poke it after sim_new (em1d/input/absorbing.c:42) */
	int shift = 0;
	if (emf->moving_window) shift = ( (emf->iter + 1) * emf->dt ) > emf->dx * ( emf->n_move + 1 );
	zdev_emf1d_advance(zb_dev(e), zb_dev(c), emf->dt, emf->dx, (int) emf->bc_type, shift);
	e->eb_host_stale = 1;
	e->part_host_stale = 1;
	emf->iter += 1;
	if (shift) emf->n_move++;
	if (!zb_opt_lazy()) zdev_sync();
	zb_guard_refresh();
	emf_seconds += timer_interval_seconds(t0, timer_ticks());
}

void emf_get_energy( const t_emf *emf, double energy[] )
{
	zb_emf_to_device((t_emf*) emf);
	zdev_emf1d_energy(zb_dev(zb_grid_of_emf(emf, 1)), energy);
	for (int i = 0; i < 6; i++) energy[i] *= 0.5 * emf->dx;
}

void emf_report( const t_emf *emf, const char field, const int fc )
{
	if (fc < 0 || fc > 2) {
		fprintf(stderr, "(*error*) Invalid field component (fc) selected, returning\n");
		return;
	}
	zb_emf_to_host(emf);
	char name[16], label[16];
	const float3* f;
	const char comp = "xyz"[fc];
	switch (field) {
	case EFLD:  f = emf->E;      snprintf(name, 16, "E%1d", fc);      snprintf(label, 16, "E_%c", comp); break;
	case BFLD:  f = emf->B;      snprintf(name, 16, "B%1d", fc);      snprintf(label, 16, "B_%c", comp); break;
	case EPART: f = emf->E_part; snprintf(name, 16, "E%1d-part", fc); snprintf(label, 16, "E_{%cp}", comp); break;
	case BPART: f = emf->B_part; snprintf(name, 16, "B%1d-part", fc); snprintf(label, 16, "B_{%cp}", comp); break;
	default:
		fprintf(stderr, "Invalid field type selected, returning\n");
		return;
	}
	float* buf = malloc((size_t) emf->nx * sizeof(float));
	for (int i = 0; i < emf->nx; i++) buf[i] = ((const float*) f)[3*i + fc];
	t_zdf_grid_axis axis[1] = { { .min = 0.0 + emf->n_move * emf->dx, .max = emf->box + emf->n_move * emf->dx,
	                              .name = "x", .label = "x", .units = "c/\\omega_p" } };
	t_zdf_grid_info info = { .ndims = 1, .name = name, .label = label, .units = "m_e c \\omega_p e^{-1}", .axis = axis };
	info.count[0] = emf->nx;
	t_zdf_iteration iter = { .name = "ITERATION", .n = emf->iter, .t = emf->iter * emf->dt, .time_units = "1/\\omega_p" };
	zdf_save_grid(buf, zdf_float32, &info, &iter, "EMF");
	free(buf);
	zb_guard_refresh();
}
