/* zpic-b200 :: em2d electromagnetic fields, host side of the API (reference em2d/emf.c).
 *
 * What stays on the host: construction, the one-off laser launch (libm double
 * precision, reference emf.c:158-350), initial / external field set-up and the
 * ZDF reports.  What runs on the device: the field advance, guard cells, window
 * shift, external-field superposition and the energy reduction
 * (csrc/dev/zdev_grid2d.cu).  E_buf/B_buf are mirrors, refreshed on demand.
 */
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

static size_t emf_ncell( const t_emf* emf ) {
	return (size_t) (emf->gc[0][0] + emf->nx[0] + emf->gc[0][1]) *
	       (size_t) (emf->gc[1][0] + emf->nx[1] + emf->gc[1][1]);
}

void emf_new( t_emf *emf, int nx[], float box[], const float dt )
{
	zb_grid_drop_emf(emf);

	for (int i = 0; i < 2; i++) {
		emf->nx[i] = nx[i];
		emf->gc[i][0] = 1;       /* reference emf.c:59-60 */
		emf->gc[i][1] = 2;
		emf->box[i] = box[i];
		emf->dx[i] = box[i] / nx[i];
	}
	emf->nrow = nx[0] + 3;
	size_t n = emf_ncell(emf);
	emf->E_buf = zb_guard_alloc(n * sizeof(float3));       /* guarded mirrors: ../common/zb_guard.h */
	emf->B_buf = zb_guard_alloc(n * sizeof(float3));
	if (!emf->E_buf || !emf->B_buf) { fprintf(stderr, "(*error*) emf_new: out of memory\n"); exit(-1); }
	zb_guard_bind_emf(emf);
	emf->E = emf->E_buf + 1 + emf->nrow;
	emf->B = emf->B_buf + 1 + emf->nrow;

	emf->dt = dt;
	emf->iter = 0;
	emf->moving_window = 0;
	emf->n_move = 0;

	memset(&emf->ext_fld, 0, sizeof emf->ext_fld);
	emf->ext_fld.E_type = EMF_FLD_TYPE_NONE;
	emf->ext_fld.B_type = EMF_FLD_TYPE_NONE;
	emf->E_part = emf->E;
	emf->B_part = emf->B;

	zb_grid* e = zb_grid_of_emf(emf, 1);
	e->eb_dev_stale = 0;      /* both sides start from zero */
}

void emf_delete( t_emf *emf )
{
	zb_grid_drop_emf(emf);
	if (zdev_ready()) { zdev_host_unpin(emf->E_buf); zdev_host_unpin(emf->B_buf); }
	zb_guard_free(emf->E_buf); zb_guard_free(emf->B_buf);
	emf->E_buf = emf->B_buf = NULL;
	if (emf->ext_fld.E_type > EMF_FLD_TYPE_NONE) zb_guard_free(emf->ext_fld.E_part_buf);
	if (emf->ext_fld.B_type > EMF_FLD_TYPE_NONE) zb_guard_free(emf->ext_fld.B_part_buf);
	emf->E_part = emf->B_part = NULL;
}

/* ------------------------------------------------------------------ laser launch (host) */

/* transverse profile and phase of a Gaussian beam (reference emf.c:158-170).
 * float intermediates, double libm calls: kept identical for bit-equal fields. */
static float gaussian_beam( const t_emf_laser* l, const float z_, const float r )
{
	float z = z_ - l->focus;
	float z0 = l->omega0 * ( l->W0 * l->W0 ) / 2;
	float rho2 = r * r;
	float curv = rho2 * z / ( z0*z0 + z*z );
	float rWl2 = ( z0*z0 ) / ( z0*z0 + z*z );
	float gouy = atan2( z, z0 );
	return sqrt( sqrt(rWl2) ) * exp( - rho2 * rWl2 / ( l->W0 * l->W0 ) ) *
	       cos( l->omega0 * ( z + curv ) - gouy );
}

/* longitudinal sin^2 envelope (reference emf.c:179-202) */
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

/* longitudinal components from div E = div B = 0, integrating from the right edge
 * (reference div_corr_x, emf.c:213-231; double accumulators) */
static void fix_divergence_x( t_emf* emf )
{
	float3* E = emf->E; float3* B = emf->B;
	const int nrow = emf->nrow;
	const double dx_dy = emf->dx[0] / emf->dx[1];
	for (int j = 0; j < emf->nx[1]; j++) {
		double ex = 0.0, bx = 0.0;
		for (int i = emf->nx[0] - 1; i >= 0; i--) {
			ex += dx_dy * ( E[i+1 + j*nrow].y - E[i+1 + (j-1)*nrow].y );
			E[i + j*nrow].x = ex;
			bx += dx_dy * ( B[i + (j+1)*nrow].y - B[i + j*nrow].y );
			B[i + j*nrow].x = bx;
		}
	}
}

/* periodic guard refresh on the HOST mirror, only used right after the launch
 * (reference emf_update_gc as called from emf_add_laser, emf.c:348) */
static void host_update_gc( t_emf* emf )
{
	const int nrow = emf->nrow, nx = emf->nx[0], ny = emf->nx[1];
	float3* F[2] = { emf->E, emf->B };
	for (int f = 0; f < 2; f++) {
		float3* A = F[f];
		if (!emf->moving_window)
			for (int j = -1; j < ny + 2; j++) {
				A[-1 + j*nrow] = A[nx - 1 + j*nrow];
				A[nx + j*nrow] = A[0 + j*nrow];
				A[nx + 1 + j*nrow] = A[1 + j*nrow];
			}
		for (int i = -1; i < nx + 2; i++) {
			A[i - nrow] = A[i + (ny - 1)*nrow];
			A[i + ny*nrow] = A[i];
			A[i + (ny + 1)*nrow] = A[i + nrow];
		}
	}
}

void emf_add_laser( t_emf* const emf, t_emf_laser* laser )
{
	/* parameter checks and fwhm override, as the reference (emf.c:246-270);
	   note the caller's struct is modified */
	if (laser->fwhm != 0) {
		if (laser->fwhm <= 0) { fprintf(stderr, "Invalid laser FWHM, must be > 0, aborting.\n"); exit(-1); }
		laser->rise = laser->fwhm; laser->fall = laser->fwhm; laser->flat = 0.;
	}
	if (laser->rise <= 0) { fprintf(stderr, "Invalid laser RISE, must be > 0, aborting.\n"); exit(-1); }
	if (laser->flat < 0)  { fprintf(stderr, "Invalid laser FLAT, must be >= 0, aborting.\n"); exit(-1); }
	if (laser->fall <= 0) { fprintf(stderr, "Invalid laser FALL, must be > 0, aborting.\n"); exit(-1); }

	zb_emf_to_host(emf);      /* superimpose on the current fields */

	float3* E = emf->E; float3* B = emf->B;
	const int nrow = emf->nrow;
	const float dx = emf->dx[0], dy = emf->dx[1];
	const float r_center = laser->axis;
	const float amp = laser->omega0 * laser->a0;
	const float cos_pol = cos( laser->polarization );
	const float sin_pol = sin( laser->polarization );

	if (laser->type == PLANE) {
		const float k = laser->omega0;
		for (int i = 0; i < emf->nx[0]; i++) {
			float z = i * dx, z_2 = z + dx/2;
			float lenv = amp * envelope(laser, z), lenv_2 = amp * envelope(laser, z_2);
			for (int j = 0; j < emf->nx[1]; j++) {
				E[i + j*nrow].y += +lenv * cos( k * z ) * cos_pol;
				E[i + j*nrow].z += +lenv * cos( k * z ) * sin_pol;
				B[i + j*nrow].y += -lenv_2 * cos( k * z_2 ) * sin_pol;
				B[i + j*nrow].z += +lenv_2 * cos( k * z_2 ) * cos_pol;
			}
		}
	} else if (laser->type == GAUSSIAN) {
		for (int i = 0; i < emf->nx[0]; i++) {
			float z = i * dx, z_2 = z + dx/2;
			float lenv = amp * envelope(laser, z), lenv_2 = amp * envelope(laser, z_2);
			for (int j = 0; j < emf->nx[1]; j++) {
				float r = j * dy - r_center, r_2 = r + dy/2;
				E[i + j*nrow].y += +lenv * gaussian_beam(laser, z, r_2) * cos_pol;
				E[i + j*nrow].z += +lenv * gaussian_beam(laser, z, r) * sin_pol;
				B[i + j*nrow].y += -lenv_2 * gaussian_beam(laser, z_2, r) * sin_pol;
				B[i + j*nrow].z += +lenv_2 * gaussian_beam(laser, z_2, r_2) * cos_pol;
			}
		}
		fix_divergence_x(emf);
	}

	host_update_gc(emf);
	zb_grid_of_emf(emf, 1)->eb_dev_stale = 1;
}

/* ------------------------------------------------------------------ initial / external fields */

void emf_init_fld( t_emf* const emf, t_emf_init_fld* init_fld )
{
	if (emf->iter != 0) {
		fprintf(stderr, "emf_init_fld should only be called at initialization, aborting...\n");
		exit(-1);
	}
	zb_emf_to_host(emf);
	const int nrow = emf->nrow;
	for (int f = 0; f < 2; f++) {
		float3* A = f ? emf->B : emf->E;
		enum emf_fld_type type = f ? init_fld->B_type : init_fld->E_type;
		float3 v0 = f ? init_fld->B_0 : init_fld->E_0;
		float3 (*fn)(int, float, int, float, void*) = f ? init_fld->B_custom : init_fld->E_custom;
		void* data = f ? init_fld->B_custom_data : init_fld->E_custom_data;
		if (type == EMF_FLD_TYPE_NONE) continue;
		for (int j = -emf->gc[1][0]; j < emf->nx[1] + emf->gc[1][1]; j++)
			for (int i = -emf->gc[0][0]; i < emf->nx[0] + emf->gc[0][1]; i++)
				A[j*nrow + i] = (type == EMF_FLD_TYPE_UNIFORM) ? v0 : fn(i, emf->dx[0], j, emf->dx[1], data);
	}
	zb_grid_of_emf(emf, 1)->eb_dev_stale = 1;
}

/* evaluate a custom external field once over the whole buffer (the callbacks take no
 * time argument, so the reference's per-step re-evaluation yields the same values;
 * SURVEY.md App. D) */
static float3* eval_ext_grid( const t_emf* emf, float3 (*fn)(int, float, int, float, void*), void* data )
{
	float3* buf = malloc(emf_ncell(emf) * sizeof(float3));
	float3* org = buf + 1 + emf->nrow;
	for (int j = -emf->gc[1][0]; j < emf->nx[1] + emf->gc[1][1]; j++)
		for (int i = -emf->gc[0][0]; i < emf->nx[0] + emf->gc[0][1]; i++)
			org[j*emf->nrow + i] = fn(i, emf->dx[0], j, emf->dx[1], data);
	return buf;
}

void emf_set_ext_fld( t_emf* const emf, t_emf_ext_fld* ext_fld )
{
	zb_emf_to_device(emf);
	zb_grid* e = zb_grid_of_emf(emf, 1);
	const size_t bytes = emf_ncell(emf) * sizeof(float3);

	if (emf->ext_fld.E_type > EMF_FLD_TYPE_NONE) zb_guard_free(emf->ext_fld.E_part_buf);
	if (emf->ext_fld.B_type > EMF_FLD_TYPE_NONE) zb_guard_free(emf->ext_fld.B_part_buf);

	emf->ext_fld.E_type = ext_fld->E_type;
	emf->ext_fld.B_type = ext_fld->B_type;
	if ((unsigned) ext_fld->E_type > EMF_FLD_TYPE_CUSTOM || (unsigned) ext_fld->B_type > EMF_FLD_TYPE_CUSTOM) {
		fprintf(stderr, "Invalid external field type, aborting.\n");
		exit(-1);
	}

	/* host-visible buffers of the fields seen by particles (exposed by the Python API) */
	if (ext_fld->E_type == EMF_FLD_TYPE_NONE) { emf->E_part = emf->E; emf->ext_fld.E_part_buf = NULL; }
	else {
		emf->ext_fld.E_0 = ext_fld->E_0;
		emf->ext_fld.E_custom = ext_fld->E_custom; emf->ext_fld.E_custom_data = ext_fld->E_custom_data;
		emf->ext_fld.E_part_buf = zb_guard_alloc(bytes);
		emf->E_part = emf->ext_fld.E_part_buf + 1 + emf->nrow;
	}
	if (ext_fld->B_type == EMF_FLD_TYPE_NONE) { emf->B_part = emf->B; emf->ext_fld.B_part_buf = NULL; }
	else {
		emf->ext_fld.B_0 = ext_fld->B_0;
		emf->ext_fld.B_custom = ext_fld->B_custom; emf->ext_fld.B_custom_data = ext_fld->B_custom_data;
		emf->ext_fld.B_part_buf = zb_guard_alloc(bytes);
		emf->B_part = emf->ext_fld.B_part_buf + 1 + emf->nrow;
	}

	/* device side */
	float e0[3] = { ext_fld->E_0.x, ext_fld->E_0.y, ext_fld->E_0.z };
	float b0[3] = { ext_fld->B_0.x, ext_fld->B_0.y, ext_fld->B_0.z };
	zdev_emf_set_ext_uniform(zb_dev(e), ext_fld->E_type == EMF_FLD_TYPE_UNIFORM, e0,
	                               ext_fld->B_type == EMF_FLD_TYPE_UNIFORM, b0);
	float3 *ge = NULL, *gb = NULL;
	if (ext_fld->E_type == EMF_FLD_TYPE_CUSTOM) ge = eval_ext_grid(emf, ext_fld->E_custom, ext_fld->E_custom_data);
	if (ext_fld->B_type == EMF_FLD_TYPE_CUSTOM) gb = eval_ext_grid(emf, ext_fld->B_custom, ext_fld->B_custom_data);
	if (e->slab.on) {
		/* the device grid is this rank's window of the box-wide buffers */
		const int nrl = e->slab.nxl + 3, nrows = emf->nx[1] + 3;
		for (int f = 0; f < 2; f++) {
			float3** gp = f ? &gb : &ge;
			if (!*gp) continue;
			float3* w = malloc((size_t) nrl * nrows * sizeof(float3));
			for (int r = 0; r < nrows; r++)
				memcpy(w + (size_t) r * nrl, *gp + (size_t) r * emf->nrow + e->slab.x0, (size_t) nrl * sizeof(float3));
			free(*gp); *gp = w;
		}
	}
	if (ge || gb) zdev_emf_set_ext_grid(zb_dev(e), (const float*) ge, (const float*) gb);
	free(ge); free(gb);
	e->part_host_stale = 1;
	zb_guard_bind_emf(emf);
	zb_guard_refresh();
}

/* custom-field callback that reads a table the caller filled (zpic_b200.h) */
float3 zpic_b200_table_field( int ix, float dx, int iy, float dy, void* data )
{
	const zpic_b200_field_table* t = (const zpic_b200_field_table*) data;
	const float* v = t->table + 3 * ((size_t) (ix + 1) + (size_t) (iy + 1) * t->nrow);
	(void) dx; (void) dy;
	return (float3) { v[0], v[1], v[2] };
}

/* ------------------------------------------------------------------ advance */

void emf_advance( t_emf *emf, const t_current *current )
{
	uint64_t t0 = timer_ticks();

	zb_emf_to_device(emf);
	zb_grid* e = zb_grid_of_emf(emf, 1);
	zb_grid* c = zb_grid_of_cur(current, 1);

	/* window test with the iteration number already advanced, in float, exactly as
	   emf_move_window does (reference emf.c:707-712, :650) */
	int shift = 0;
	if (emf->moving_window)
		shift = ( (emf->iter + 1) * emf->dt ) > emf->dx[0] * ( emf->n_move + 1 );

	zdev_emf_advance(zb_dev(e), zb_dev(c), emf->dt, emf->dx[0], emf->dx[1], emf->moving_window, shift);
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
	zb_grid* e = zb_grid_of_emf(emf, 1);
	zdev_emf_energy(zb_dev(e), energy);
	if (e->slab.on) zb_par_allreduce_sum_d(energy, 6);      /* the slabs' interiors tile the box */
	for (int i = 0; i < 6; i++) energy[i] *= 0.5 * emf->dx[0] * emf->dx[1];
}

/* ------------------------------------------------------------------ report */

void emf_report( const t_emf *emf, const char field, const int fc )
{
	if (fc < 0 || fc > 2) {
		fprintf(stderr, "(*error*) Invalid field component (fc) selected, returning\n");
		return;
	}
	zb_emf_to_host(emf);
	if (zb_par_rank() != 0) { zb_guard_refresh(); return; }          /* one file per box: rank 0 writes it */

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

	const int nx = emf->nx[0], ny = emf->nx[1];
	float* buf = malloc((size_t) nx * ny * sizeof(float));
	for (int j = 0; j < ny; j++) {
		const float* row = (const float*) (f + (size_t) j * emf->nrow);
		for (int i = 0; i < nx; i++) buf[(size_t) j * nx + i] = row[3*i + fc];
	}

	t_zdf_grid_axis axis[2] = {
		{ .min = 0.0, .max = emf->box[0], .name = "x", .label = "x", .units = "c/\\omega_p" },
		{ .min = 0.0, .max = emf->box[1], .name = "y", .label = "y", .units = "c/\\omega_p" }
	};
	t_zdf_grid_info info = { .ndims = 2, .name = name, .label = label,
	                         .units = "m_e c \\omega_p e^{-1}", .axis = axis };
	info.count[0] = nx; info.count[1] = ny;
	t_zdf_iteration iter = { .name = "ITERATION", .n = emf->iter,
	                         .t = emf->iter * emf->dt, .time_units = "1/\\omega_p" };
	zdf_save_grid(buf, zdf_float32, &info, &iter, "EMF");
	free(buf);
	zb_guard_refresh();
}
