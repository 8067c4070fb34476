/* zpic-b200 :: em1d particle species, host side of the API (reference em1d/particles.c).
 * Host: species construction and injector (sequential, global random stream: bit-identical initial
 * conditions and window columns), ZDF diagnostics.  Device: spec_advance and the charge deposit
 * (csrc/dev/zdev_spec1d.cu).  spec->part is a mirror refreshed on demand. */
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include "zb_state.h"
#include "random.h"
#include "../common/zb_rand.h"
#include "timer.h"
#include "zdf.h"

static double   push_seconds = 0.0;
static uint64_t push_count = 0;
double   spec_time( void )  { return push_seconds; }
uint64_t spec_npush( void ) { return push_count; }
double   spec_perf( void )  { return (push_count > 0) ? push_seconds / push_count : -1.0; }

/* ------------------------------------------------------------------ injection (host) */

/* `mirror`: the buffer is a species' host mirror (a guarded mapping, ../common/zb_guard.h; `keep` particles of it
   are worth copying), otherwise a scratch buffer from the C allocator */
static void grow( t_part** buf, int* np_max, int size, const t_species* mirror, int keep )
{
	if (size > *np_max) {
		*np_max = ( size/1024 + 1 ) * 1024;
		if (mirror) {
			*buf = zb_guard_realloc(*buf, (size_t) *np_max * sizeof(t_part), (size_t) keep * sizeof(t_part));
			zb_guard_bind_spec(mirror, *buf);
		} else *buf = realloc(*buf, (size_t) *np_max * sizeof(t_part));
		if (!*buf) { fprintf(stderr, "(*error*) species buffer: out of memory\n"); exit(-1); }
	}
}
/* room for `size` particles in the mirror; what it holds is NOT preserved (it is about to be overwritten) */
void zb_spec_reserve( t_species* spec, const int size )
{
	if (size <= spec->np_max) return;
	const long long want = (long long) spec->np_max + spec->np_max / 4;
	grow(&spec->part, &spec->np_max, (want > size && want < 0x7ffffc00LL) ? (int) want : size, spec, 0);
}
void spec_grow_buffer( t_species* spec, const int size )
{
	if (size <= spec->np_max) return;
	zb_spec_to_host(spec);        /* callers append to what the buffer holds (Species.add, em1d.pyx): make it current */
	grow(&spec->part, &spec->np_max, size, spec, spec->np);
}

/* upper bound of the particles the profile puts in cells range[0]..range[1]
 * (reference spec_np_inj, em1d/particles.c:147-232) */
static int count_upper_bound( t_species* spec, const int range[] )
{
	const t_density* d = &spec->density;
	switch (d->type) {
	case STEP: {
		int i0 = d->start / spec->dx - spec->n_move;
		if (i0 > range[1]) return 0;
		if (i0 < range[0]) i0 = range[0];
		return ( range[1] - i0 + 1 ) * spec->ppc;
	}
	case SLAB: {
		int i0 = d->start / spec->dx - spec->n_move;
		int i1 = d->end / spec->dx - spec->n_move;
		if (i0 > range[1] || i1 < range[0]) return 0;
		if (i0 < range[0]) i0 = range[0];
		if (i1 > range[1]) i1 = range[1];
		return ( i1 - i0 + 1 ) * spec->ppc;
	}
	case RAMP: {
		float x0 = d->start, x1 = d->end;
		float a = (range[0] + spec->n_move) * spec->dx;
		float b = (range[1] + 1 + spec->n_move) * spec->dx;
		if ( (x1 <= x0) || (a > x1) || (b < x0) ) return 0;
		if (a < x0) a = x0;
		if (b > x1) b = x1;
		float n0 = d->ramp[0], n1 = d->ramp[1];
		float q = (b-a)*( n0 + 0.5 * (a+b-2*x0)*(n1-n0)/(x1-x0));
		return q * spec->ppc / spec->dx;
	}
	case CUSTOM: {
		double q = 0.5 * ( d->custom((range[0] + spec->n_move) * spec->dx, d->custom_data) +
		                   d->custom((range[1] + 1 + spec->n_move) * spec->dx, d->custom_data) );
		for (int i = range[0] + 1; i <= range[1]; i++) q += d->custom((i + spec->n_move) * spec->dx, d->custom_data);
		return ceil(q * spec->ppc);
	}
	case EMPTY:
		return 0;
	default:
		return ( range[1] - range[0] + 1 ) * spec->ppc;
	}
}

/* positions (reference spec_set_x, em1d/particles.c:240-445); returns the new particle count */
static int place_particles( t_species* spec, const int range[], t_part* part, const int ip0 )
{
	const int npc = spec->ppc;
	int ip = ip0;
	float* pos = malloc((size_t) npc * sizeof(float));
	for (int i = 0; i < npc; i++) pos[i] = ( i + 0.5 ) / npc;

	switch (spec->density.type) {
	case RAMP: {
		/* inverse CDF of a linear ramp, continued across calls through total_np_inj */
		double r0 = spec->density.start / spec->dx, r1 = spec->density.end / spec->dx;
		if ( ((range[0] + spec->n_move) > r1) || ((range[1] + spec->n_move) < r0) ) break;
		double n0 = spec->density.ramp[0], n1 = spec->density.ramp[1];
		if (r0 < 0) { n0 += - r0 * (n1-n0) / (r1-r0); r0 = 0; }
		const double cpp = 1.0 / spec->ppc;
		for (int k = spec->density.total_np_inj; ; k++) {
			double Rs = (k + 0.5) * cpp / (r1 - r0);
			double p = 2 * Rs / ( sqrt( n0*n0 + 2 * (n1-n0) * Rs ) + n0 );
			if (p > 1) break;
			p = r0 + (r1-r0) * p;
			int ix = p;
			if (ix - spec->n_move < range[0]) {
				fprintf(stderr, "(*error*) attempting to inject outside of valid range.\n");
				break;
			}
			if (ix - spec->n_move > range[1]) break;
			part[ip].ix = ix - spec->n_move;
			part[ip].x = p - ix;
			ip++;
		}
		break;
	}
	case CUSTOM: {
		const double dx = spec->dx, cpp = 1.0 / spec->ppc;
		int k = spec->density.total_np_inj;
		int ix = range[0];
		double n0, n1 = spec->density.custom((ix + spec->n_move) * dx, spec->density.custom_data);
		double d0, d1 = spec->density.custom_q_inj, Rs;
		while (ix <= range[1]) {
			n0 = n1;
			n1 = spec->density.custom((ix + 1 + spec->n_move) * dx, spec->density.custom_data);
			d0 = d1;
			d1 += 0.5 * (n0 + n1);
			while ( ( Rs = (k + 0.5) * cpp ) < d1 ) {
				double p = 2 * (Rs-d0) / ( sqrt( n0*n0 + 2 * (n1-n0) * (Rs-d0) ) + n0 );
				part[ip].ix = ix;
				part[ip].x = p;
				ip++; k++;
			}
			ix++;
		}
		spec->density.custom_q_inj = d1;
		break;
	}
	case EMPTY:
		break;
	default: {
		/* UNIFORM / STEP / SLAB lattice */
		const enum density_type type = spec->density.type;
		float lo = 0, hi = 0;
		if (type == STEP || type == SLAB) lo = spec->density.start / spec->dx - spec->n_move;
		if (type == SLAB) hi = spec->density.end / spec->dx - spec->n_move;
		for (int i = range[0]; i <= range[1]; i++)
			for (int k = 0; k < npc; k++) {
				if (type == STEP && !( i + pos[k] > lo )) continue;
				if (type == SLAB && !( i + pos[k] > lo && i + pos[k] < hi )) continue;
				part[ip].ix = i;
				part[ip].x = pos[k];
				ip++;
			}
	}
	}
	free(pos);
	spec->density.total_np_inj += ip - ip0;
	return ip;
}

/* thermal momenta minus the cell mean plus fluid momentum (reference spec_set_u, em1d/particles.c:88-130) */
static void draw_momenta( t_species* spec, t_part* part, int first, int last )
{
	for (int i = first; i <= last; i++) {
		part[i].ux = spec->uth[0] * rand_norm();
		part[i].uy = spec->uth[1] * rand_norm();
		part[i].uz = spec->uth[2] * rand_norm();
	}
	float3* mean = calloc(spec->nx, sizeof(float3));
	int* count = calloc(spec->nx, sizeof(int));
	for (int i = first; i <= last; i++) {
		const int c = part[i].ix;
		mean[c].x += part[i].ux; mean[c].y += part[i].uy; mean[c].z += part[i].uz;
		count[c] += 1;
	}
	for (int c = 0; c < spec->nx; c++) {
		const float norm = (count[c] > 0) ? 1.0f / count[c] : 0;
		mean[c].x *= norm; mean[c].y *= norm; mean[c].z *= norm;
	}
	for (int i = first; i <= last; i++) {
		const int c = part[i].ix;
		part[i].ux += spec->ufl[0] - mean[c].x;
		part[i].uy += spec->ufl[1] - mean[c].y;
		part[i].uz += spec->ufl[2] - mean[c].z;
	}
	free(count); free(mean);
}

void spec_inject_into( t_species* spec, const int range[], t_part** buf, int* np, int* np_max )
{
	const int first = *np;
	grow(buf, np_max, *np + count_upper_bound(spec, range), (buf == &spec->part) ? spec : NULL, *np);
	*np = place_particles(spec, range, *buf, *np);
	draw_momenta(spec, *buf, first, *np - 1);
}

/* Lattice profiles (UNIFORM / STEP / SLAB): the in-cell positions k with lo[i] <= k < hi[i] of cell i carry plasma -
   the clip of place_particles() evaluated with the same float expressions.  Returns the particle count, or -1 when
   the reference-stream device injector does not cover the profile (RAMP, CUSTOM). */
static long long lattice_cells( const t_species* spec, int** lo_out, int** hi_out )
{
	const enum density_type type = spec->density.type;
	if (type != UNIFORM && type != STEP && type != SLAB) return -1;
	const int nx = spec->nx, npc = spec->ppc;
	float* pos = malloc((size_t) npc * sizeof(float));
	for (int i = 0; i < npc; i++) pos[i] = ( i + 0.5 ) / npc;
	float lo = 0, hi = 0;
	if (type == STEP || type == SLAB) lo = spec->density.start / spec->dx - spec->n_move;
	if (type == SLAB) hi = spec->density.end / spec->dx - spec->n_move;
	int* klo = malloc((size_t) nx * sizeof(int)); int* khi = malloc((size_t) nx * sizeof(int));
	long long total = 0;
	for (int i = 0; i < nx; i++) {
		int a = 0, b = npc;
		if (type != UNIFORM) {
			a = npc; b = 0;
			for (int k = 0; k < npc; k++) {
				int in = 1;
				if (type == STEP) in = ( i + pos[k] > lo );
				if (type == SLAB) in = ( i + pos[k] > lo && i + pos[k] < hi );
				if (in) { if (k < a) a = k; b = k + 1; }
			}
			if (b <= a) { a = 0; b = 0; }
		}
		klo[i] = a; khi[i] = b;
		total += b - a;
	}
	free(pos);
	*lo_out = klo; *hi_out = khi;
	return total;
}

void spec_new( t_species* spec, char name[], const float m_q, const int ppc,
               const float *ufl, const float *uth,
               const int nx, float box, const float dt, t_density* density )
{
	zb_spec_drop(spec);
	strncpy(spec->name, name, MAX_SPNAME_LEN);
	spec->name[MAX_SPNAME_LEN] = 0;
	spec->nx = nx;
	spec->ppc = ppc;
	spec->box = box;
	spec->dx = box / nx;
	spec->m_q = m_q;
	spec->q = copysign( 1.0f, m_q ) / ppc;
	spec->dt = dt;
	spec->energy = 0;
	spec->np_max = 0;
	spec->part = NULL;
	if (density) {
		spec->density = *density;
		if (spec->density.n == 0.) spec->density.n = 1.0;
	} else {
		spec->density = (t_density) { .type = UNIFORM, .n = 1.0 };
	}
	spec->density.total_np_inj = 0;
	spec->density.custom_q_inj = 0.;
	spec->q *= fabsf( spec->density.n );
	for (int i = 0; i < 3; i++) {
		spec->ufl[i] = ufl ? ufl[i] : 0;
		spec->uth[i] = uth ? uth[i] : 0;
	}
	spec->iter = 0;
	spec->moving_window = 0;
	spec->n_move = 0;
	spec->np = 0;
	const int range[2] = { 0, nx - 1 };
	int *lat_lo = NULL, *lat_hi = NULL;
	long long total2 = -1;
	if (zb_opt_device_init() == 2) total2 = lattice_cells(spec, &lat_lo, &lat_hi);
	if (total2 >= 0) {
		/* device_init = 2: the reference's own initial population on the reference's random stream, generated on the
		   device at the first step; the host stream is moved past its 3 deviates per particle now (see the em2d twin) */
		zb_spec* e = zb_spec_of(spec, 1);
		zb_rand_get_state(&e->rs_z, &e->rs_w, &e->rs_have, &e->rs_spare);
		uint32_t z = e->rs_z, w = e->rs_w; int have = e->rs_have; double spare = e->rs_spare;
		if (zdev_ref_normals(&z, &w, &have, &spare, 3 * total2, spec->uth, NULL) == 0) {
			zb_rand_set_state(z, w, have, spare);
			e->device_init = 2;
			e->lat_lo = lat_lo; e->lat_hi = lat_hi;
			spec->np = (total2 > 0x7fffffffLL) ? 0x7fffffff : (int) total2;
			spec->density.total_np_inj += total2;
		} else {
			free(lat_lo); free(lat_hi);
			total2 = -1;
		}
	}
	if (total2 >= 0) {
		/* done above */
	} else if (zb_opt_device_init() == 1 && spec->density.type == UNIFORM) {
		zb_spec* e = zb_spec_of(spec, 1);
		e->device_init = 1;
		e->device_seed = ((uint64_t) rand_uint32() << 32) | rand_uint32();
		long long total = (long long) nx * ppc;
		spec->np = (total > 0x7fffffffLL) ? 0x7fffffff : (int) total;
	} else {
		spec_inject_into(spec, range, &spec->part, &spec->np, &spec->np_max);
	}
	spec->n_sort = 16;
	spec->bc_type = PART_BC_PERIODIC;
}

void spec_delete( t_species* spec )
{
	zb_spec_drop(spec);
	zb_guard_free(spec->part);
	spec->part = NULL;
	spec->np = -1;
}

/* stand-alone window move on the host mirror (reference em1d/particles.c:663-684) */
void spec_move_window( t_species *spec )
{
	if ( (spec->iter * spec->dt) > (spec->dx * (spec->n_move + 1)) ) {
		zb_spec_to_host(spec);
		for (int i = 0; i < spec->np; i++) spec->part[i].ix--;
		spec->n_move++;
		const int range[2] = { spec->nx - 1, spec->nx - 1 };
		spec_inject_into(spec, range, &spec->part, &spec->np, &spec->np_max);
		zb_spec_of(spec, 1)->dev_stale = 1;
		zb_guard_refresh();
	}
}

/* ------------------------------------------------------------------ advance (device) */

void spec_advance( t_species* spec, t_emf* emf, t_current* current )
{
	uint64_t t0 = timer_ticks();
	zb_spec_to_device(spec);
	zb_emf_to_device(emf);
	zb_spec* s = zb_spec_of(spec, 1);
	zb_grid* gf = zb_grid_of_emf(emf, 1);
	zb_grid* gc = zb_grid_of_cur(current, 1);

	zdev_push1d_params prm;
	prm.tem   = 0.5 * spec->dt / spec->m_q;           /* double arithmetic, as the reference (particles.c:925) */
	prm.dt_dx = spec->dt / spec->dx;
	prm.qnx   = spec->q * spec->dx / spec->dt;
	prm.q     = spec->q;
	prm.absorbing = spec->moving_window || spec->bc_type == PART_BC_OPEN;
	prm.shift_window = spec->moving_window && ( ((spec->iter + 1) * spec->dt) > (spec->dx * (spec->n_move + 1)) );

	zdev_spec1d_advance(zb_spec_dev(s), zb_dev(gf), zb_dev(gc), &prm);
	s->host_stale = 1;
	gc->j_host_stale = 1;
	spec->iter += 1;

	if (prm.shift_window) {
		spec->n_move++;
		const int range[2] = { spec->nx - 1, spec->nx - 1 };
		t_part* col = NULL; int ncol = 0, ncol_max = 0;
		spec_inject_into(spec, range, &col, &ncol, &ncol_max);
		zdev_spec1d_append(zb_spec_dev(s), col, ncol);
		free(col);
	}
	if (!zb_opt_lazy()) {
		double esum; int64_t np;
		zdev_spec1d_fetch(zb_spec_dev(s), &esum, &np);
		spec->energy = spec->q * spec->m_q * esum * spec->dx;
		/* the count is taken after the append: the injected column is already in it */
		spec->np = (np > 0x7fffffffLL) ? 0x7fffffff : (int) np;
		s->np_seen = spec->np;
		/* a guarded mirror is filled from inside a fault handler, where the buffer cannot move: make room now */
		if (zb_guard_enabled() && spec->np > spec->np_max && spec->np < 0x7ff00000) zb_spec_reserve(spec, spec->np);
	}
	zb_guard_refresh();
	push_count += spec->np;
	push_seconds += timer_interval_seconds(t0, timer_ticks());
}

void spec_deposit_charge( const t_species* spec, float* charge )
{
	zb_spec_to_device((t_species*) spec);
	zdev_spec1d_deposit_charge(zb_spec_dev(zb_spec_of(spec, 1)), spec->q, spec->moving_window, charge);
}

/* ------------------------------------------------------------------ reports (host, ZDF) */

static void report_particles( const t_species *spec )
{
	static const char* quants[]  = { "x", "ux", "uy", "uz" };
	static const char* qlabels[] = { "x", "u_x", "u_y", "u_z" };
	static const char* qunits[]  = { "c/\\omega_p", "c", "c", "c" };
	t_zdf_iteration iter = { .name = "ITERATION", .n = spec->iter, .t = spec->iter * spec->dt, .time_units = "1/\\omega_p" };
	t_zdf_part_info info = { .name = (char*) spec->name, .label = (char*) spec->name, .nquants = 4,
	                         .quants = (char**) quants, .qlabels = (char**) qlabels, .qunits = (char**) qunits, .np = spec->np };
	char path[1024];
	snprintf(path, 1024, "PARTICLES/%s", spec->name);
	t_zdf_file file;
	zdf_open_part_file(&file, &info, &iter, path);
	const int np = spec->np;
	const t_part* p = spec->part;
	float* data = malloc((size_t) (np > 0 ? np : 1) * sizeof(float));
	for (int i = 0; i < np; i++) data[i] = ( spec->n_move + p[i].ix + p[i].x ) * spec->dx;
	zdf_add_quant_part_file(&file, quants[0], data, np);
	for (int i = 0; i < np; i++) data[i] = p[i].ux;
	zdf_add_quant_part_file(&file, quants[1], data, np);
	for (int i = 0; i < np; i++) data[i] = p[i].uy;
	zdf_add_quant_part_file(&file, quants[2], data, np);
	for (int i = 0; i < np; i++) data[i] = p[i].uz;
	zdf_add_quant_part_file(&file, quants[3], data, np);
	free(data);
	zdf_close_file(&file);
}

static void report_charge( const t_species *spec )
{
	float* rho = calloc((size_t) spec->nx + 1, sizeof(float));
	spec_deposit_charge(spec, rho);
	t_zdf_grid_axis axis = { .min = spec->n_move * spec->dx, .max = spec->box + spec->n_move * spec->dx,
	                         .name = "x", .label = "x", .units = "c/\\omega_p" };
	char name[128], label[128];
	snprintf(name, 128, "%s-charge", spec->name);
	snprintf(label, 128, "%s \\rho", spec->name);
	t_zdf_grid_info info = { .ndims = 1, .name = name, .label = label, .units = "n_e", .axis = &axis };
	info.count[0] = spec->nx;
	t_zdf_iteration iter = { .name = "ITERATION", .n = spec->iter, .t = spec->iter * spec->dt, .time_units = "1/\\omega_p" };
	char path[1024];
	snprintf(path, 1024, "CHARGE/%s", spec->name);
	zdf_save_grid(rho, zdf_float32, &info, &iter, path);
	free(rho);
}

static inline float pha_value( const t_species* spec, const t_part* p, int quant )
{
	switch (quant) {
	case X1: return ( p->x + p->ix ) * spec->dx;
	case U1: return p->ux;
	case U2: return p->uy;
	case U3: return p->uz;
	}
	return 0;
}

void spec_deposit_pha( const t_species *spec, const int rep_type,
                       const int pha_nx[], const float pha_range[][2], float* buf )
{
	zb_spec_to_host(spec);
	const int nrow = pha_nx[0];
	const int quant1 = rep_type & 0x000F, quant2 = (rep_type & 0x00F0) >> 4;
	const float x1min = pha_range[0][0], x2min = pha_range[1][0];
	const float rdx1 = pha_nx[0] / ( pha_range[0][1] - pha_range[0][0] );
	const float rdx2 = pha_nx[1] / ( pha_range[1][1] - pha_range[1][0] );
	for (int k = 0; k < spec->np; k++) {
		const t_part* p = &spec->part[k];
		float nx1 = ( pha_value(spec, p, quant1) - x1min ) * rdx1;
		float nx2 = ( pha_value(spec, p, quant2) - x2min ) * rdx2;
		int i1 = (int) (nx1 + 0.5f), i2 = (int) (nx2 + 0.5f);
		float w1 = nx1 - i1 + 0.5f, w2 = nx2 - i2 + 0.5f;
		int idx = i1 + nrow * i2;
		const int in1a = (i1 >= 0 && i1 < pha_nx[0]), in1b = (i1 + 1 >= 0 && i1 + 1 < pha_nx[0]);
		if (i2 >= 0 && i2 < pha_nx[1]) {
			if (in1a) buf[idx]     += (1.0f - w1) * (1.0f - w2) * spec->q;
			if (in1b) buf[idx + 1] += w1 * (1.0f - w2) * spec->q;
		}
		idx += nrow;
		if (i2 + 1 >= 0 && i2 + 1 < pha_nx[1]) {
			if (in1a) buf[idx]     += (1.0f - w1) * w2 * spec->q;
			if (in1b) buf[idx + 1] += w1 * w2 * spec->q;
		}
	}
}

static void report_pha( const t_species *spec, const int rep_type, const int pha_nx[], const float pha_range[][2] )
{
	float* buf = calloc((size_t) pha_nx[0] * pha_nx[1], sizeof(float));
	spec_deposit_pha(spec, rep_type, pha_nx, pha_range, buf);
	const int q1 = rep_type & 0x000F, q2 = (rep_type & 0x00F0) >> 4;
	static const char* ax_name[]  = { "x1", "x2", "x3", "u1", "u2", "u3" };
	static const char* ax_label[] = { "x", "y", "z", "u_x", "u_y", "u_z" };
	const char* u1 = (q1 == X1) ? "c/\\omega_p" : "m_e c";
	const char* u2 = (q2 == X1) ? "c/\\omega_p" : "m_e c";
	t_zdf_grid_axis axis[2] = {
		{ .min = pha_range[0][0], .max = pha_range[0][1], .name = (char*) ax_name[q1-1], .label = (char*) ax_label[q1-1], .units = (char*) u1 },
		{ .min = pha_range[1][0], .max = pha_range[1][1], .name = (char*) ax_name[q2-1], .label = (char*) ax_label[q2-1], .units = (char*) u2 }
	};
	char name[64], label[64];
	snprintf(name, 64, "%s-%s%s", spec->name, ax_name[q1-1], ax_name[q2-1]);
	snprintf(label, 64, "%s %s-%s", spec->name, ax_label[q1-1], ax_label[q2-1]);
	t_zdf_grid_info info = { .ndims = 2, .name = name, .label = label, .units = "a.u.", .axis = axis };
	info.count[0] = pha_nx[0]; info.count[1] = pha_nx[1];
	t_zdf_iteration iter = { .name = "ITERATION", .n = spec->iter, .t = spec->iter * spec->dt, .time_units = "1/\\omega_p" };
	char path[1024];
	snprintf(path, 1024, "PHASESPACE/%s", spec->name);
	zdf_save_grid(buf, zdf_float32, &info, &iter, path);
	free(buf);
}

void spec_report( const t_species *spec, const int rep_type, const int pha_nx[], const float pha_range[][2] )
{
	switch (rep_type & 0xF000) {
	case CHARGE: report_charge(spec); break;
	case PHA: report_pha(spec, rep_type, pha_nx, pha_range); break;
	case PARTICLES: zb_spec_to_host(spec); report_particles(spec); break;
	}
	zb_guard_refresh();
}
