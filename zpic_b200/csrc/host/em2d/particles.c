/* zpic-b200 :: em2d particle species, host side of the API (reference em2d/particles.c).
 *
 * Host: species construction and the particle injector (sequential, drawing from the
 * single global random stream so that initial conditions and every moving-window
 * column are bit-identical to the reference), diagnostics output.
 * Device: spec_advance - field interpolation, Boris push, current deposit,
 * boundaries, window shift, tile binning (csrc/dev/zdev_spec2d.cu) - and the charge
 * deposit.  spec->part is a mirror that is refreshed only when somebody reads it.
 */
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

static float density_one( float x, void* data ) { (void) x; (void) data; return 1.0; }

/* `mirror`: the buffer is a species' host mirror (a guarded mapping, ../common/zb_guard.h; `keep` particles of it
   are worth copying), otherwise a scratch buffer from the C allocator */
static void grow( t_part** buf, int* np_max, int size, const t_species* mirror, int keep )
{
	if (size > *np_max) {
		*np_max = ( size/1024 + 1 ) * 1024;       /* 1024-particle chunks (reference particles.c:463) */
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
	if (spec->part && zdev_ready()) zdev_host_forget(spec->part);
	/* a growing population (moving window) asks again and again: a quarter of headroom */
	const long long want = (long long) spec->np_max + spec->np_max / 4;
	grow(&spec->part, &spec->np_max, (want > size && want < 0x7ffffc00LL) ? (int) want : size, spec, 0);
}

void spec_grow_buffer( t_species* spec, const int size )
{
	if (size <= spec->np_max) return;
	/* callers (Species.add of the Python module, em2d.pyx:253) append to what the buffer holds: make it current */
	zb_spec_to_host(spec);
	if (spec->part && zdev_ready()) zdev_host_forget(spec->part);   /* the buffer may move */
	grow(&spec->part, &spec->np_max, size, spec, spec->np);
}

/* upper bound of the number of particles the profile puts in `range`
 * (reference spec_np_inj, particles.c:367-449, incl. the dx[1] slip in SLAB) */
static int count_upper_bound( t_species* spec, const int range[][2] )
{
	const int ncx = range[0][1] - range[0][0] + 1, ncy = range[1][1] - range[1][0] + 1;
	switch (spec->density.type) {
	case STEP: {
		int i0 = spec->density.start / spec->dx[0] - spec->n_move;
		int n;
		if (i0 > range[0][1]) n = 0;
		else { if (i0 < range[0][0]) i0 = range[0][0]; n = ( range[0][1] - i0 + 1 ) * spec->ppc[0]; }
		return n * ncy * spec->ppc[1];
	}
	case SLAB: {
		int i0 = spec->density.start / spec->dx[0] - spec->n_move;
		int i1 = spec->density.end / spec->dx[1] - spec->n_move;
		int n;
		if (i0 > range[0][1] || i1 < range[0][0]) n = 0;
		else {
			if (i0 < range[0][0]) i0 = range[0][0];
			if (i1 > range[0][1]) i1 = range[0][1];
			n = ( i1 - i0 + 1 ) * spec->ppc[0];
		}
		return n * ncy * spec->ppc[1];
	}
	case CUSTOM: {
		/* trapezoidal integrals of the two profile factors */
		double x = (range[0][0] + spec->n_move) * spec->dx[0];
		double qx = spec->density.custom_x(x, spec->density.custom_data_x);
		x = (range[0][1] + 1 + spec->n_move) * spec->dx[0];
		qx += spec->density.custom_x(x, spec->density.custom_data_x);
		qx *= 0.5;
		for (int i = range[0][0] + 1; i <= range[0][1]; i++) {
			x = (i + spec->n_move) * spec->dx[0];
			qx += spec->density.custom_x(x, spec->density.custom_data_x);
		}
		double y = range[1][0] * spec->dx[1];
		double qy = spec->density.custom_y(y, spec->density.custom_data_y);
		y = (range[1][1] + 1) * spec->dx[1];
		qy += spec->density.custom_y(y, spec->density.custom_data_y);
		qy *= 0.5;
		for (int j = range[1][0] + 1; j <= range[1][1]; j++) {
			y = j * spec->dx[1];
			qy += spec->density.custom_y(y, spec->density.custom_data_y);
		}
		return ceil(qx * spec->ppc[0]) * ceil(qy * spec->ppc[1]);
	}
	case EMPTY:
		return 0;
	default:
		return ncx * spec->ppc[0] * ncy * spec->ppc[1];
	}
}

/* place particles in `range` following the density profile; returns the new count
 * (reference spec_set_x, particles.c:159-354) */
static int place_particles( t_species* spec, const int range[][2], t_part* part, int ip )
{
	const int npc = spec->ppc[0] * spec->ppc[1];
	const float dpcx = 1.0f / spec->ppc[0], dpcy = 1.0f / spec->ppc[1];

	/* positions inside a cell, x fastest */
	float* cx = malloc(npc * sizeof(float)); float* cy = malloc(npc * sizeof(float));
	for (int j = 0, k = 0; j < spec->ppc[1]; j++)
		for (int i = 0; i < spec->ppc[0]; i++, k++) {
			cx[k] = dpcx * ( i + 0.5 );
			cy[k] = dpcy * ( j + 0.5 );
		}

	const enum density_type type = spec->density.type;
	if (type == CUSTOM) {
		/* inverse-CDF placement on a piecewise linear density, x then y, with the running
		   x integral carried across calls so window columns continue the sequence */
		const double dx = spec->dx[0], dy = spec->dx[1];
		const double cppx = 1.0 / spec->ppc[0], cppy = 1.0 / spec->ppc[1];
		const double thresh = 4 * cppx * cppy;
		const int n_movex = spec->n_move;
		const int ix0 = range[0][0], iy0 = range[1][0];

		unsigned long kx = spec->density.custom_x_total_part;
		double d0x, d1x = spec->density.custom_x_total_q;
		double n0x, n1x = spec->density.custom_x((ix0 + n_movex) * dx, spec->density.custom_data_x);
		const double n0y0 = spec->density.custom_y(iy0 * dy, spec->density.custom_data_y);

		for (int ix = ix0; ix <= range[0][1]; ix++) {
			n0x = n1x;
			n1x = spec->density.custom_x((ix + 1 + n_movex) * dx, spec->density.custom_data_x);
			d0x = d1x;
			d1x += 0.5 * (n0x + n1x);
			double Rsx;
			while ( (Rsx = (kx + 0.5) * cppx) < d1x ) {
				double x = 2 * (Rsx - d0x) / ( sqrt( n0x*n0x + 2 * (n1x - n0x) * (Rsx - d0x) ) + n0x );
				double nx = (0.5 - x) * n0x + (0.5 + x) * n1x;
				int ky = 0;
				double n0y, n1y = n0y0, d0y, d1y = 0;
				for (int iy = iy0; iy <= range[1][1]; iy++) {
					n0y = n1y;
					n1y = spec->density.custom_y((iy + 1) * dy, spec->density.custom_data_y);
					d0y = d1y;
					d1y += 0.5 * (n0y + n1y);
					double Rsy;
					while ( (Rsy = (ky + 0.5) * cppy) < d1y ) {
						double y = 2 * (Rsy - d0y) / ( sqrt( n0y*n0y + 2 * (n1y - n0y) * (Rsy - d0y) ) + n0y );
						double ny = (0.5 - y) * n0y + (0.5 + y) * n1y;
						if (nx * ny > thresh) {
							part[ip].ix = ix; part[ip].iy = iy;
							part[ip].x = x;   part[ip].y = y;
							ip++;
						}
						ky++;
					}
				}
				kx++;
			}
		}
		spec->density.custom_x_total_q = d1x;
		spec->density.custom_x_total_part = kx;
	} else if (type != EMPTY) {
		/* UNIFORM / STEP / SLAB: regular lattice, optionally clipped in x */
		float lo = 0, hi = 0;
		if (type == STEP || type == SLAB) lo = spec->density.start / spec->dx[0] - spec->n_move;
		if (type == SLAB) hi = spec->density.end / spec->dx[0] - spec->n_move;
		for (int j = range[1][0]; j <= range[1][1]; j++)
			for (int i = range[0][0]; i <= range[0][1]; i++)
				for (int k = 0; k < npc; k++) {
					if (type == STEP && !( i + cx[k] > lo )) continue;
					if (type == SLAB && !( i + cx[k] > lo && i + cx[k] < hi )) continue;
					part[ip].ix = i; part[ip].iy = j;
					part[ip].x = cx[k]; part[ip].y = cy[k];
					ip++;
				}
	}
	free(cx); free(cy);
	return ip;
}

/* thermal momenta for part[first..last], cell-mean removed, fluid momentum added
 * (reference spec_set_u, particles.c:96-146; the cell index uses nx[1] as stride,
 * App. B item 5 - kept, it decides which particles share a mean).
 * The reference allocates and sweeps nx0*nx1 accumulators on every call, which is invisible next to its
 * own push but would dominate a GPU step when a moving window injects one column per iteration.  The
 * sums are order dependent only within a cell, so accumulating in particle order into a hash table of
 * the cells actually touched gives bit-identical means at O(particles) cost. */
static void draw_momenta( t_species* spec, t_part* part, int first, int last )
{
	for (int i = first; i <= last; i++) {
		part[i].ux = spec->uth[0] * rand_norm();
		part[i].uy = spec->uth[1] * rand_norm();
		part[i].uz = spec->uth[2] * rand_norm();
	}
	const int n = last - first + 1;
	if (n <= 0) return;

	const int stride = spec->nx[1];
	size_t cap = 16;
	while (cap < (size_t) 2 * n) cap <<= 1;
	int* key = malloc(cap * sizeof(int));
	int* count = calloc(cap, sizeof(int));
	float3* mean = calloc(cap, sizeof(float3));
	int* slot_of = malloc((size_t) n * sizeof(int));
	memset(key, 0xff, cap * sizeof(int));              /* -1 = free */

	for (int i = first; i <= last; i++) {
		const int c = part[i].ix + stride * part[i].iy;
		size_t h = ((size_t) (unsigned) c * 2654435761u) & (cap - 1);
		while (key[h] != -1 && key[h] != c) h = (h + 1) & (cap - 1);
		key[h] = c;
		mean[h].x += part[i].ux; mean[h].y += part[i].uy; mean[h].z += part[i].uz;
		count[h] += 1;
		slot_of[i - first] = (int) h;
	}
	for (size_t h = 0; h < cap; h++) {
		if (key[h] == -1) continue;
		const float norm = 1.0f / count[h];
		mean[h].x *= norm; mean[h].y *= norm; mean[h].z *= norm;
	}
	for (int i = first; i <= last; i++) {
		const int h = slot_of[i - first];
		part[i].ux += spec->ufl[0] - mean[h].x;
		part[i].uy += spec->ufl[1] - mean[h].y;
		part[i].uz += spec->ufl[2] - mean[h].z;
	}
	free(slot_of); free(mean); free(count); free(key);
}

/* inject into an arbitrary AoS buffer (the species mirror at start-up, a scratch
 * column buffer when the window moves) - reference spec_inject_particles :476-492 */
void spec_inject_into( t_species* spec, const int range[][2], t_part** buf, int* np, int* np_max )
{
	const int first = *np;
	const int is_mirror = (buf == &spec->part);
	grow(buf, np_max, *np + count_upper_bound(spec, range), is_mirror ? spec : NULL, *np);
	*np = place_particles(spec, range, *buf, *np);
	draw_momenta(spec, *buf, first, *np - 1);
}

/* the cells [ix0, ix1) x [iy0, iy1) a density profile fills completely, if it is of that kind (device-side
   initialisation only handles those): whole cells, decided at the cell centre */
static int device_init_rect( const t_species* spec, int rect[4] )
{
	const t_density* d = &spec->density;
	const int nx = spec->nx[0], ny = spec->nx[1];
	rect[0] = 0; rect[1] = nx; rect[2] = 0; rect[3] = ny;
	switch (d->type) {
	case UNIFORM: return 1;
	case STEP:
	case SLAB: {
		int i0 = 0, i1 = nx;
		while (i0 < nx && (i0 + 0.5f) * spec->dx[0] < d->start) i0++;
		if (d->type == SLAB) { i1 = i0; while (i1 < nx && (i1 + 0.5f) * spec->dx[0] < d->end) i1++; }
		rect[0] = i0; rect[1] = i1;
		return 1;
	}
	case CUSTOM: {
		if (d->custom_x && d->custom_x != &density_one) return 0;
		if (!d->custom_y) return 1;
		int j0 = -1, j1 = -1;
		for (int j = 0; j < ny; j++) {
			const float v = d->custom_y((j + 0.5f) * spec->dx[1], d->custom_data_y);
			if (v != 0.0f && v != 1.0f) return 0;              /* a real profile: the host injector */
			if (v == 1.0f) { if (j0 < 0) j0 = j; else if (j1 >= 0) return 0; }     /* a second band */
			else if (j0 >= 0 && j1 < 0) j1 = j;
		}
		if (j0 < 0) { rect[2] = rect[3] = 0; return 1; }
		rect[2] = j0; rect[3] = (j1 < 0) ? ny : j1;
		return 1;
	}
	default: return 0;
	}
}

/* Lattice profiles (UNIFORM / STEP / SLAB): the in-cell x positions kx with lo[i] <= kx < hi[i] of column i carry
   plasma - the clip of place_particles() evaluated with the same float expressions; every row is alike.  Returns
   the particles per row, or -1 when the reference-stream device injector does not cover the species (CUSTOM
   profiles; a warm plasma in a box with nx[0] != nx[1], whose cell means mix cells - SURVEY.md App. B 5). */
static long long lattice_columns( const t_species* spec, int** lo_out, int** hi_out )
{
	const enum density_type type = spec->density.type;
	if (type != UNIFORM && type != STEP && type != SLAB) return -1;
	const int warm = spec->uth[0] != 0 || spec->uth[1] != 0 || spec->uth[2] != 0;
	if (warm && spec->nx[0] != spec->nx[1]) return -1;
	const int nx = spec->nx[0], ppcx = spec->ppc[0];
	const float dpcx = 1.0f / ppcx;
	float lo = 0, hi = 0;
	if (type == STEP || type == SLAB) lo = spec->density.start / spec->dx[0] - spec->n_move;
	if (type == SLAB) hi = spec->density.end / spec->dx[0] - spec->n_move;
	int* klo = malloc((size_t) nx * sizeof(int)); int* khi = malloc((size_t) nx * sizeof(int));
	long long per_row = 0;
	for (int i = 0; i < nx; i++) {
		int a = ppcx, b = 0;                     /* first and one-past-last passing position (they form a range) */
		for (int k = 0; k < ppcx; k++) {
			const float cx = dpcx * ( k + 0.5 );
			int in = 1;
			if (type == STEP) in = ( i + cx > lo );
			if (type == SLAB) in = ( i + cx > lo && i + cx < hi );
			if (in) { if (k < a) a = k; b = k + 1; }
		}
		if (b <= a) { a = 0; b = 0; }
		klo[i] = a; khi[i] = b;
		per_row += (long long) (b - a) * spec->ppc[1];
	}
	*lo_out = klo; *hi_out = khi;
	return per_row;
}

void spec_new( t_species* spec, char name[], const float m_q, const int ppc[],
               const float *ufl, const float *uth,
               const int nx[], float box[], const float dt, t_density* density )
{
	zb_spec_drop(spec);

	strncpy(spec->name, name, MAX_SPNAME_LEN);
	spec->name[MAX_SPNAME_LEN] = 0;

	int npc = 1;
	for (int i = 0; i < 2; i++) {
		spec->nx[i] = nx[i];
		spec->ppc[i] = ppc[i];
		npc *= ppc[i];
		spec->box[i] = box[i];
		spec->dx[i] = box[i] / nx[i];
	}
	spec->m_q = m_q;
	spec->q = copysign( 1.0f, m_q ) / npc;
	spec->dt = dt;
	spec->energy = 0;

	spec->np_max = 0;
	spec->part = NULL;

	if (density) spec->density = *density;
	else spec->density = (t_density) { .type = UNIFORM, .n = 1.0 };
	if (spec->density.n == 0.) spec->density.n = 1.0;
	if (spec->density.type == CUSTOM) {
		if (!spec->density.custom_x) spec->density.custom_x = &density_one;
		if (!spec->density.custom_y) spec->density.custom_y = &density_one;
		spec->density.custom_x_total_part = 0;
		spec->density.custom_x_total_q = 0;
	}
	spec->q *= fabsf( spec->density.n );

	for (int i = 0; i < 3; i++) {
		spec->ufl[i] = ufl ? ufl[i] : 0;
		spec->uth[i] = uth ? uth[i] : 0;
	}

	spec->iter = 0;
	spec->moving_window = 0;
	spec->n_move = 0;

	spec->np = 0;
	const int range[][2] = { {0, nx[0]-1}, {0, nx[1]-1} };
	int rect[4];
	int *lat_lo = NULL, *lat_hi = NULL;
	long long per_row = -1;
	if (zb_opt_device_init() == 2) per_row = lattice_columns(spec, &lat_lo, &lat_hi);
	if (per_row >= 0) {
		/* device_init = 2: the reference's own initial population - same positions, same injection order, momenta
		   from the same global random stream - generated on the device at the first step (zdev_refrng.cu).  The
		   host stream is moved past this species' 3 deviates per particle NOW, so whatever is created next (the
		   second species, a window column) draws what it would have drawn in the reference. */
		zb_spec* e = zb_spec_of(spec, 1);
		zb_rand_get_state(&e->rs_z, &e->rs_w, &e->rs_have, &e->rs_spare);
		uint32_t z = e->rs_z, w = e->rs_w; int have = e->rs_have; double spare = e->rs_spare;
		const long long total = per_row * nx[1];
		if (zdev_ref_normals(&z, &w, &have, &spare, 3 * total, spec->uth, NULL) == 0) {
			zb_rand_set_state(z, w, have, spare);
			e->device_init = 2;
			e->lat_lo = lat_lo; e->lat_hi = lat_hi;
			spec->np = (total > 0x7fffffffLL) ? 0x7fffffff : (int) total;
		} else {
			/* seeds outside the generators' linear range: the host injector */
			free(lat_lo); free(lat_hi);
			per_row = -1;
		}
	}
	if (per_row >= 0) {
		/* done above */
	} else if (zb_opt_device_init() == 1 && device_init_rect(spec, rect)) {
		/* opt-in for populations too large for a host mirror: the plasma fills a rectangle of cells at the nominal
		   particles per cell (whole box; from a STEP / inside a SLAB along x; the band a 0/1 CUSTOM profile along
		   y selects), generated by a counter-based generator on the device at the first step (one draw of the
		   host stream seeds it, so runs stay reproducible and species differ) - same distribution as
		   spec_set_x / spec_set_u, not the reference random stream */
		zb_spec* e = zb_spec_of(spec, 1);
		e->device_init = 1;
		e->device_seed = ((uint64_t) rand_uint32() << 32) | rand_uint32();
		memcpy(e->dev_rect, rect, sizeof rect);
		long long total = (long long) (rect[1] - rect[0]) * (rect[3] - rect[2]) * npc;
		spec->np = (total > 0x7fffffffLL) ? 0x7fffffff : (int) total;
	} else {
		spec_inject_into(spec, range, &spec->part, &spec->np, &spec->np_max);
	}

	spec->n_sort = 16;    /* kept for API compatibility; device tiles are re-binned every step */
}

void spec_delete( t_species* spec )
{
	zb_spec_drop(spec);
	if (spec->part && zdev_ready()) zdev_host_forget(spec->part);
	zb_guard_free(spec->part);
	spec->part = NULL;
	spec->np = -1;
}

/* Stand-alone window move on the host mirror (reference particles.c:619-640).  The
 * device path does not come through here: spec_advance folds the shift into the push
 * kernel and uploads the injected column. */
void spec_move_window( t_species *spec )
{
	if ( (spec->iter * spec->dt) > (spec->dx[0] * (spec->n_move + 1)) ) {
		zb_spec_to_host(spec);
		for (int i = 0; i < spec->np; i++) spec->part[i].ix--;
		spec->n_move++;
		const int range[][2] = { {spec->nx[0]-1, spec->nx[0]-1}, {0, spec->nx[1]-1} };
		spec_inject_into(spec, range, &spec->part, &spec->np, &spec->np_max);
		zb_spec* e = zb_spec_of(spec, 1);
		e->dev_stale = 1;
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

	/* per-step scalars exactly as the reference forms them (particles.c:1111-1117) */
	zdev_push2d_params prm;
	prm.tem   = 0.5 * spec->dt / spec->m_q;
	prm.dt_dx = spec->dt / spec->dx[0];
	prm.dt_dy = spec->dt / spec->dx[1];
	prm.qnx   = spec->q * spec->dx[0] / spec->dt;
	prm.qny   = spec->q * spec->dx[1] / spec->dt;
	prm.q     = spec->q;
	prm.moving_window = spec->moving_window;
	prm.slab_left = prm.slab_right = 0;      /* single-slab API path; slabs are driven through zpic_dev.h */
	/* the window test uses the already incremented iteration (particles.c:1234-1240, :621) */
	prm.shift_window = spec->moving_window &&
		( ((spec->iter + 1) * spec->dt) > (spec->dx[0] * (spec->n_move + 1)) );

	zdev_spec2d_advance(zb_spec_dev(s), zb_dev(gf), zb_dev(gc), &prm);
	s->host_stale = 1;
	gc->j_host_stale = 1;
	spec->iter += 1;

	if (prm.shift_window) {
		/* new plasma enters through the right edge: host injector (global random stream),
		   then the column is appended to the device tiles */
		spec->n_move++;
		if (s->device_made) {
			/* a species that was generated on the device: so is its new column, wherever the profile (whole cells,
			   decided at the cell centre like the initial fill) has plasma at the window's right edge */
			const float xc = (spec->n_move + spec->nx[0] - 1 + 0.5f) * spec->dx[0];
			const t_density* d = &spec->density;
			const int in = (d->type == UNIFORM || d->type == CUSTOM) || (d->type == STEP && xc >= d->start) ||
			               (d->type == SLAB && xc >= d->start && xc < d->end);
			if (in && (!s->slab.on || s->slab.is_last))
				zdev_spec2d_inject_column(zb_spec_dev(s), spec->ppc[0], spec->ppc[1], spec->ufl, spec->uth, s->device_seed,
				                          s->slab.nxl - 1, s->dev_rect[2], s->dev_rect[3],
				                          (uint64_t) (spec->n_move + spec->nx[0] - 1));
		} else {
			const int range[][2] = { {spec->nx[0]-1, spec->nx[0]-1}, {0, spec->nx[1]-1} };
			t_part* col = NULL; int ncol = 0, ncol_max = 0;
			spec_inject_into(spec, range, &col, &ncol, &ncol_max);
			if (!s->slab.on) zdev_spec2d_append(zb_spec_dev(s), col, ncol);
			else if (s->slab.is_last) {
				/* every rank runs the injector (the global random stream stays in step); the column belongs to the last slab */
				for (int i = 0; i < ncol; i++) col[i].ix -= s->slab.x0;
				zdev_spec2d_append(zb_spec_dev(s), col, ncol);
			}
			free(col);
		}
	}

	if (!zb_opt_lazy()) {
		double esum; int64_t np;
		zdev_spec2d_fetch(zb_spec_dev(s), &esum, &np);
		if (s->slab.on) {          /* the diagnostics of the API are those of the whole box */
			long long n = np;
			zb_par_allreduce_sum_d(&esum, 1);
			zb_par_allreduce_sum_ll(&n, 1);
			np = n;
		}
		spec->energy = spec->q * spec->m_q * esum * spec->dx[0] * spec->dx[1];
		/* the count is taken after the append: the injected column is already in it */
		spec->np = (np > 0x7fffffffLL) ? 0x7fffffff : (int) np;
		s->np_seen = spec->np;
		push_count += spec->np;
		/* a guarded mirror is filled from inside a fault handler, where the buffer cannot move: make room now */
		if (zb_guard_enabled() && spec->np > spec->np_max && spec->np < 0x7ff00000 && zb_par_init() <= 1)
			zb_spec_reserve(spec, spec->np);
	} else {
		push_count += spec->np;      /* population estimate; exact after the next sync */
	}
	zb_guard_refresh();
	push_seconds += timer_interval_seconds(t0, timer_ticks());
}

/* ------------------------------------------------------------------ charge density */

void spec_deposit_charge( const t_species* spec, float* charge )
{
	zb_spec_to_device((t_species*) spec);
	zb_spec* s = zb_spec_of(spec, 1);
	if (!s->slab.on) {
		zdev_spec2d_deposit_charge(zb_spec_dev(s), spec->q, spec->moving_window, charge);
		return;
	}
	/* slabs: every rank deposits its own particles on its (nxl+1) x (ny+1) nodes; the slabs are joined in box
	   coordinates (a slab's last node column is its right neighbour's first: summed), then the box's periodic
	   folds of the reference (particles.c:1310-1322) are applied and the result ADDED to the caller's array */
	const int nx = spec->nx[0], ny = spec->nx[1], nxl = s->slab.nxl, x0 = s->slab.x0;
	float* loc = malloc((size_t) (nxl + 1) * (ny + 1) * sizeof(float));
	zdev_spec2d_deposit_charge_raw(zb_spec_dev(s), spec->q, loc);
	float* box = calloc((size_t) (nx + 1) * (ny + 1), sizeof(float));
	for (int j = 0; j <= ny; j++)
		for (int i = 0; i <= nxl; i++) box[(size_t) j * (nx + 1) + x0 + i] = loc[(size_t) j * (nxl + 1) + i];
	zb_par_allreduce_sum_f(box, (size_t) (nx + 1) * (ny + 1));
	if (!spec->moving_window)
		for (int j = 0; j <= ny; j++) box[(size_t) j * (nx + 1)] += box[(size_t) j * (nx + 1) + nx];
	for (int i = 0; i <= nx; i++) box[i] += box[(size_t) ny * (nx + 1) + i];
	for (size_t k = 0; k < (size_t) (nx + 1) * (ny + 1); k++) charge[k] += box[k];
	free(loc); free(box);
}

/* ------------------------------------------------------------------ reports (host, ZDF) */

static void report_particles( const t_species *spec )
{
	static const char* quants[]  = { "x", "y", "ux", "uy", "uz" };
	static const char* qlabels[] = { "x", "y", "u_x", "u_y", "u_z" };
	static const char* qunits[]  = { "c/\\omega_p", "c/\\omega_p", "c", "c", "c" };

	t_zdf_iteration iter = { .name = "ITERATION", .n = spec->iter,
	                         .t = spec->iter * spec->dt, .time_units = "1/\\omega_p" };
	t_zdf_part_info info = { .name = (char*) spec->name, .label = (char*) spec->name, .nquants = 5,
	                         .quants = (char**) quants, .qlabels = (char**) qlabels,
	                         .qunits = (char**) qunits, .np = spec->np };
	char path[1024];
	snprintf(path, 1024, "PARTICLES/%s", spec->name);
	t_zdf_file file;
	zdf_open_part_file(&file, &info, &iter, path);

	const int np = spec->np;
	const t_part* p = spec->part;
	float* data = malloc((size_t) (np > 0 ? np : 1) * sizeof(float));
	for (int i = 0; i < np; i++) data[i] = ( spec->n_move + p[i].ix + p[i].x ) * spec->dx[0];
	zdf_add_quant_part_file(&file, quants[0], data, np);
	for (int i = 0; i < np; i++) data[i] = ( p[i].iy + p[i].y ) * spec->dx[1];
	zdf_add_quant_part_file(&file, quants[1], data, np);
	for (int i = 0; i < np; i++) data[i] = p[i].ux;
	zdf_add_quant_part_file(&file, quants[2], data, np);
	for (int i = 0; i < np; i++) data[i] = p[i].uy;
	zdf_add_quant_part_file(&file, quants[3], data, np);
	for (int i = 0; i < np; i++) data[i] = p[i].uz;
	zdf_add_quant_part_file(&file, quants[4], data, np);
	free(data);
	zdf_close_file(&file);
}

static void report_charge( const t_species *spec )
{
	const int nx = spec->nx[0], ny = spec->nx[1];
	float* rho = calloc((size_t) (nx + 1) * (ny + 1), sizeof(float));
	spec_deposit_charge(spec, rho);
	if (zb_par_rank() != 0) { free(rho); return; }       /* one file per box: rank 0 writes it */
	float* buf = malloc((size_t) nx * ny * sizeof(float));
	for (int j = 0; j < ny; j++)
		memcpy(buf + (size_t) j * nx, rho + (size_t) j * (nx + 1), nx * sizeof(float));
	free(rho);

	t_zdf_grid_axis axis[2] = {
		{ .min = spec->n_move * spec->dx[0], .max = spec->box[0] + spec->n_move * spec->dx[0],
		  .name = "x", .label = "x", .units = "c/\\omega_p" },
		{ .min = 0.0, .max = spec->box[1], .name = "y", .label = "y", .units = "c/\\omega_p" }
	};
	char name[128], label[128];
	snprintf(name, 128, "%s-charge", spec->name);
	snprintf(label, 128, "%s \\rho", spec->name);
	t_zdf_grid_info info = { .ndims = 2, .name = name, .label = label, .units = "n_e", .axis = axis };
	info.count[0] = nx; info.count[1] = ny;
	t_zdf_iteration iter = { .name = "ITERATION", .n = spec->iter,
	                         .t = spec->iter * spec->dt, .time_units = "1/\\omega_p" };
	char path[1024];
	snprintf(path, 1024, "CHARGE/%s", spec->name);
	zdf_save_grid(buf, zdf_float32, &info, &iter, path);
	free(buf);
}

/* value of one phasespace axis quantity for particle p (reference spec_pha_axis :1512-1538) */
static inline float pha_value( const t_species* spec, const t_part* p, int quant )
{
	switch (quant) {
	case X1: return ( p->x + p->ix ) * spec->dx[0];
	case X2: return ( p->y + p->iy ) * spec->dx[1];
	case U1: return p->ux;
	case U2: return p->uy;
	case U3: return p->uz;
	}
	return 0;
}

/* linear deposit on a 2D phasespace grid (reference :1569-1632).  When the device holds the current
   population (any time after the first step) the deposit runs there - no 28 B/particle download; while the
   host buffer is still the authoritative copy (iteration 0, Species.add) it is done here like the reference. */
void spec_deposit_pha( const t_species *spec, const int rep_type,
                       const int pha_nx[], const float pha_range[][2], float* buf )
{
	{
		zb_spec* e = zb_spec_of(spec, 0);
		if (e && e->host_stale) {
			if (!e->slab.on) {
				zdev_spec2d_deposit_pha(zb_spec_dev(e), rep_type & 0x000F, (rep_type & 0x00F0) >> 4, pha_nx, pha_range,
				                        spec->q, spec->dx[0], spec->dx[1], buf);
				return;
			}
			/* slabs: every rank deposits its own particles (box coordinates), the grids are summed */
			const size_t n = (size_t) pha_nx[0] * pha_nx[1];
			float* part = calloc(n, sizeof(float));
			zdev_spec2d_deposit_pha(zb_spec_dev(e), rep_type & 0x000F, (rep_type & 0x00F0) >> 4, pha_nx, pha_range,
			                        spec->q, spec->dx[0], spec->dx[1], part);
			zb_par_allreduce_sum_f(part, n);
			for (size_t k = 0; k < n; k++) buf[k] += part[k];
			free(part);
			return;
		}
	}

	const int nrow = pha_nx[0];
	const int quant1 = rep_type & 0x000F;
	const int quant2 = (rep_type & 0x00F0) >> 4;
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

static void report_pha( const t_species *spec, const int rep_type,
                        const int pha_nx[], const float pha_range[][2] )
{
	float* buf = calloc((size_t) pha_nx[0] * pha_nx[1], sizeof(float));
	spec_deposit_pha(spec, rep_type, pha_nx, pha_range, buf);
	if (zb_par_rank() != 0) { free(buf); return; }

	const int q1 = rep_type & 0x000F, q2 = (rep_type & 0x00F0) >> 4;
	static const char* ax_name[]  = { "x1", "x2", "x3", "u1", "u2", "u3" };
	static const char* ax_label[] = { "x", "y", "z", "u_x", "u_y", "u_z" };
	const char* u1 = (q1 <= X2) ? "c/\\omega_p" : "m_e c";
	const char* u2 = (q2 <= X2) ? "c/\\omega_p" : "m_e c";

	t_zdf_grid_axis axis[2] = {
		{ .min = pha_range[0][0], .max = pha_range[0][1], .name = (char*) ax_name[q1-1],
		  .label = (char*) ax_label[q1-1], .units = (char*) u1 },
		{ .min = pha_range[1][0], .max = pha_range[1][1], .name = (char*) ax_name[q2-1],
		  .label = (char*) ax_label[q2-1], .units = (char*) u2 }
	};
	char name[64], label[64];
	snprintf(name, 64, "%s-%s%s", spec->name, ax_name[q1-1], ax_name[q2-1]);
	snprintf(label, 64, "%s %s-%s", spec->name, ax_label[q1-1], ax_label[q2-1]);
	t_zdf_grid_info info = { .ndims = 2, .name = name, .label = label, .units = "a.u.", .axis = axis };
	info.count[0] = pha_nx[0]; info.count[1] = pha_nx[1];
	t_zdf_iteration iter = { .name = "ITERATION", .n = spec->iter,
	                         .t = spec->iter * spec->dt, .time_units = "1/\\omega_p" };
	char path[1024];
	snprintf(path, 1024, "PHASESPACE/%s", spec->name);
	zdf_save_grid(buf, zdf_float32, &info, &iter, path);
	free(buf);
}

void spec_report( const t_species *spec, const int rep_type,
                  const int pha_nx[], const float pha_range[][2] )
{
	switch (rep_type & 0xF000) {
	case CHARGE:
		report_charge(spec);
		break;
	case PHA:
		report_pha(spec, rep_type, pha_nx, pha_range);
		break;
	case PARTICLES:
		zb_spec_to_host(spec);
		if (zb_par_rank() == 0) report_particles(spec);
		break;
	}
	zb_guard_refresh();
}
