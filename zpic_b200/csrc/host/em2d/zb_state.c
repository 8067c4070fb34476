/* zpic-b200 :: registry of device twins + mirror coherence (see zb_state.h) */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "zb_state.h"
#include "zpic_b200.h"

#define ZB_MAX 256
#define ZB_MAX_RANKS_LOCAL 64
static zb_grid grids[ZB_MAX];
static int n_grids = 0;
static zb_spec specs[ZB_MAX];
static int n_specs = 0;

static int opt_lazy = -1, opt_ids = -1, opt_coherent = -1, opt_devinit = -1;
static int env_flag( const char* name ) { const char* e = getenv(name); return e && atoi(e) != 0; }
int zb_opt_lazy( void ) { if (opt_lazy < 0) opt_lazy = env_flag("ZPIC_LAZY"); return opt_lazy; }
int zb_opt_track_ids( void ) { if (opt_ids < 0) opt_ids = env_flag("ZPIC_TRACK_IDS"); return opt_ids; }
int zb_opt_coherent( void ) { if (opt_coherent < 0) opt_coherent = env_flag("ZPIC_COHERENT"); return opt_coherent; }
int zb_opt_device_init( void ) {
	if (opt_devinit < 0) { const char* e = getenv("ZPIC_DEVICE_INIT"); opt_devinit = e ? atoi(e) : 0; }
	return opt_devinit;
}

void zpic_b200_set_option( const char* name, int value ) {
	if (!strcmp(name, "lazy")) opt_lazy = value;
	else if (!strcmp(name, "track_ids")) opt_ids = value;
	else if (!strcmp(name, "coherent")) opt_coherent = value;
	else if (!strcmp(name, "device_init")) opt_devinit = value;
	else fprintf(stderr, "(*warning*) zpic_b200_set_option: unknown option %s\n", name);
}

zb_slab zb_slab_make( int nx, int window )
{
	zb_slab g;
	memset(&g, 0, sizeof g);
	const int n = zb_par_init();
	g.rank = zb_par_rank(); g.nranks = n;
	g.nxl = nx; g.x0 = 0; g.left = g.right = -1;
	if (n <= 1) return g;
	if (nx % n) {
		fprintf(stderr, "(*error*) zpic-b200: %d cells along x cannot be cut into %d equal slabs\n", nx, n);
		exit(-1);
	}
	g.on = 1;
	g.nxl = nx / n; g.x0 = g.rank * g.nxl;
	const int ring = !window;
	g.left  = (ring || g.rank > 0)     ? (g.rank + n - 1) % n : -1;
	g.right = (ring || g.rank < n - 1) ? (g.rank + 1) % n : -1;
	g.is_last = g.rank == n - 1;
	g.wrap_left = ring && g.rank == 0;
	g.wrap_right = ring && g.rank == n - 1;
	return g;
}

static zb_grid* grid_new( int nx, int ny ) {
	if (n_grids == ZB_MAX) { fprintf(stderr, "(*error*) zpic-b200: too many live field objects\n"); exit(-1); }
	zb_grid* e = &grids[n_grids++];
	memset(e, 0, sizeof(*e));
	e->nx = nx; e->ny = ny;      /* the device object is created on first device use */
	return e;
}

zdev_grid2d* zb_dev( zb_grid* e ) {
	if (!e->g) {
		const int window = (e->emf && e->emf->moving_window) || (e->cur && e->cur->moving_window);
		e->slab = zb_slab_make(e->nx, window);
		e->g = zdev_grid2d_create(e->slab.nxl, e->ny);
		if (e->slab.on) zdev_grid2d_set_slab(e->g, e->slab.left, e->slab.right, e->slab.wrap_left, e->slab.wrap_right);
	}
	return e->g;
}

zb_grid* zb_grid_of_emf( const t_emf* emf, int create ) {
	for (int i = 0; i < n_grids; i++) if (grids[i].emf == emf) return &grids[i];
	if (!create) return NULL;
	zb_grid* e = grid_new(emf->nx[0], emf->nx[1]);
	e->emf = emf;
	e->eb_dev_stale = 1;   /* whatever the host buffers hold now goes up before first use */
	return e;
}

zb_grid* zb_grid_of_cur( const t_current* cur, int create ) {
	for (int i = 0; i < n_grids; i++) if (grids[i].cur == cur) return &grids[i];
	if (!create) return NULL;
	zb_grid* e = grid_new(cur->nx[0], cur->nx[1]);
	e->cur = cur;
	return e;
}

static void grid_remove( zb_grid* e ) {
	if (e->g) zdev_grid2d_destroy(e->g);
	*e = grids[--n_grids];
}

void zb_grid_pair( const t_emf* emf, const t_current* cur ) {
	zb_grid* ge = zb_grid_of_emf(emf, 1);
	zb_grid* gc = zb_grid_of_cur(cur, 0);
	if (gc == ge) return;
	if (gc) { if (gc->emf) gc->cur = NULL; else grid_remove(gc); }
	ge = zb_grid_of_emf(emf, 1);      /* grid_remove may have moved it */
	ge->cur = cur;
}

void zb_grid_drop_emf( const t_emf* emf ) {
	zb_grid* e = zb_grid_of_emf(emf, 0);
	if (!e) return;
	if (e->cur) e->emf = NULL; else grid_remove(e);
}
void zb_grid_drop_cur( const t_current* cur ) {
	zb_grid* e = zb_grid_of_cur(cur, 0);
	if (!e) return;
	if (e->emf) e->cur = NULL; else grid_remove(e);
}

zb_spec* zb_spec_of( const t_species* spec, int create ) {
	for (int i = 0; i < n_specs; i++) if (specs[i].spec == spec) return &specs[i];
	if (!create) return NULL;
	if (n_specs == ZB_MAX) { fprintf(stderr, "(*error*) zpic-b200: too many live species\n"); exit(-1); }
	zb_spec* e = &specs[n_specs++];
	memset(e, 0, sizeof(*e));
	e->spec = spec;
	e->dev_stale = 1;
	return e;
}

zdev_spec2d* zb_spec_dev( zb_spec* e ) {
	if (!e->d) {
		const t_species* spec = e->spec;
		e->slab = zb_slab_make(spec->nx[0], spec->moving_window);
		e->d = zdev_spec2d_create(e->slab.nxl, spec->nx[1], spec->ppc[0] * spec->ppc[1], e->slab.on ? 0 : zb_opt_track_ids());
		if (e->slab.on) zdev_spec2d_set_slab(e->d, e->slab.left, e->slab.right, e->slab.x0, spec->nx[0]);
	}
	return e->d;
}

void zb_spec_drop( const t_species* spec ) {
	zb_spec* e = zb_spec_of(spec, 0);
	if (!e) return;
	if (e->d) zdev_spec2d_destroy(e->d);
	free(e->lat_lo); free(e->lat_hi);
	*e = specs[--n_specs];
}

/* ---------------------------------------------------------------- coherence */

/* the E, B mirrors are transferred again and again: page-lock them once (unpinned by emf_delete) */
static void zb_pin_emf( const t_emf* emf ) {
	const size_t bytes = (size_t) (emf->nx[0] + 3) * (emf->nx[1] + 3) * sizeof(float3);
	zdev_host_pin(emf->E_buf, bytes);
	zdev_host_pin(emf->B_buf, bytes);
}

/* ---- slabs: the mirrors are global, the device grids are this rank's window of them ---- */

static void grid_up( zb_grid* e, int which, const float3* host_buf, int nrow ) {
	if (e->slab.on) zdev_grid2d_upload_window(zb_dev(e), which, (const float*) host_buf, nrow, e->slab.x0);
	else zdev_grid2d_upload(zb_dev(e), which, (const float*) host_buf);
}

/* every rank's mirror ends up holding the whole box: own window from the device, the rest through the job's
   shared scratch area (diagnostics path: reports, host-side field edits) */
static void grid_down( zb_grid* e, int which, float3* host_buf, int nrow, int nrows ) {
	if (!e->slab.on) { zdev_grid2d_download(zb_dev(e), which, (float*) host_buf); return; }
	const zb_slab* g = &e->slab;
	zdev_grid2d_download_window(zb_dev(e), which, (float*) host_buf, nrow, g->x0);
	float3* all = zb_par_scratch((size_t) nrow * nrows * sizeof(float3));
	/* buffer column c = cell c-1: the slab owns buffer columns [x0+1, x0+1+nxl); the box's own guard columns
	   come from the first (column 0) and the last rank (the two upper ones) */
	const int c0 = (g->rank == 0) ? 0 : g->x0 + 1;
	const int c1 = g->is_last ? nrow : g->x0 + 1 + g->nxl;
	for (int r = 0; r < nrows; r++)
		memcpy(all + (size_t) r * nrow + c0, host_buf + (size_t) r * nrow + c0, (size_t) (c1 - c0) * sizeof(float3));
	zb_par_barrier();
	memcpy(host_buf, all, (size_t) nrow * nrows * sizeof(float3));
	zb_par_barrier();
}

void zb_emf_to_device( t_emf* emf ) {
	zb_grid* e = zb_grid_of_emf(emf, 1);
	if (e->eb_dev_stale) {
		zb_pin_emf(emf);
		grid_up(e, ZDEV_E, emf->E_buf, emf->nrow);
		grid_up(e, ZDEV_B, emf->B_buf, emf->nrow);
		e->eb_dev_stale = 0;
		e->eb_host_stale = 0;
	}
}

void zb_emf_to_host( const t_emf* emf ) {
	zb_grid* e = zb_grid_of_emf(emf, 0);
	if (!e) return;
	const int nrows = emf->nx[1] + 3;
	/* the library is about to write (and its callers may go on writing): open the mirrors */
	zb_guard_set(emf->E_buf, ZB_G_RW); zb_guard_set(emf->B_buf, ZB_G_RW);
	zb_guard_set(emf->ext_fld.E_part_buf, ZB_G_RW); zb_guard_set(emf->ext_fld.B_part_buf, ZB_G_RW);
	if (e->eb_host_stale) {
		zb_pin_emf(emf);
		grid_down(e, ZDEV_E, emf->E_buf, emf->nrow, nrows);
		grid_down(e, ZDEV_B, emf->B_buf, emf->nrow, nrows);
		e->eb_host_stale = 0;
	}
	if (e->part_host_stale) {
		if (emf->ext_fld.E_type != EMF_FLD_TYPE_NONE && emf->ext_fld.E_part_buf)
			grid_down(e, ZDEV_EPART, emf->ext_fld.E_part_buf, emf->nrow, nrows);
		if (emf->ext_fld.B_type != EMF_FLD_TYPE_NONE && emf->ext_fld.B_part_buf)
			grid_down(e, ZDEV_BPART, emf->ext_fld.B_part_buf, emf->nrow, nrows);
		e->part_host_stale = 0;
	}
}

void zb_cur_to_host( const t_current* cur ) {
	zb_grid* e = zb_grid_of_cur(cur, 0);
	zb_guard_set(cur->J_buf, ZB_G_RW);
	if (!e || !e->j_host_stale) return;
	zdev_host_pin(cur->J_buf, (size_t) (cur->nx[0] + 3) * (cur->nx[1] + 3) * sizeof(float3));   /* unpinned by current_delete */
	grid_down(e, ZDEV_J, cur->J_buf, cur->nrow, cur->nx[1] + 3);
	e->j_host_stale = 0;
}

void zb_spec_to_device( t_species* spec ) {
	zb_spec* e = zb_spec_of(spec, 1);
	if (e->device_init == 2) {
		/* the reference's population, generated on the device from the state the stream had at spec_new */
		zdev_spec2d* d = zb_spec_dev(e);
		uint32_t z = e->rs_z, w = e->rs_w; int have = e->rs_have; double spare = e->rs_spare;
		if (zdev_spec2d_inject_lattice(d, spec->ppc[0], spec->ppc[1], spec->ufl, spec->uth, e->lat_lo, e->lat_hi,
		                               &z, &w, &have, &spare)) {
			fprintf(stderr, "(*error*) zpic-b200: device-side injection lost the random stream\n");
			exit(-1);
		}
		free(e->lat_lo); free(e->lat_hi); e->lat_lo = e->lat_hi = NULL;
		e->device_init = 0; e->dev_stale = 0; e->host_stale = 1;
		return;
	}
	if (e->device_init) {
		/* throughput configurations: the population never existed on the host */
		zdev_spec2d* d = zb_spec_dev(e);
		zdev_spec2d_inject_rect(d, spec->ppc[0], spec->ppc[1], spec->ufl, spec->uth, e->device_seed,
		                        e->dev_rect[0] - e->slab.x0, e->dev_rect[1] - e->slab.x0, e->dev_rect[2], e->dev_rect[3]);
		e->device_init = 0; e->device_made = 1; e->dev_stale = 0; e->host_stale = 1;
		return;
	}
	/* host code that appended particles or reallocated the buffer did so on a current
	   mirror (the Python layer syncs first); take the host copy as the truth then */
	if (!e->host_stale && (e->part_seen != spec->part || e->np_seen != spec->np)) e->dev_stale = 1;
	if (e->dev_stale) {
		zdev_spec2d* d = zb_spec_dev(e);
		if (e->slab.on) {
			/* the mirror holds the whole box on every rank: this rank takes the particles of its slab */
			const int x0 = e->slab.x0, x1 = e->slab.x0 + e->slab.nxl;
			int n = 0;
			for (int i = 0; i < spec->np; i++) n += (spec->part[i].ix >= x0 && spec->part[i].ix < x1);
			t_part* mine = malloc((size_t) (n > 0 ? n : 1) * sizeof(t_part));
			n = 0;
			for (int i = 0; i < spec->np; i++)
				if (spec->part[i].ix >= x0 && spec->part[i].ix < x1) { mine[n] = spec->part[i]; mine[n].ix -= x0; n++; }
			zdev_spec2d_upload(d, mine, n);
			free(mine);
		} else {
			zdev_spec2d_upload(d, spec->part, spec->np);
		}
		e->dev_stale = 0; e->host_stale = 0;
		e->part_seen = spec->part; e->np_seen = spec->np;
	}
}

void zb_spec_to_host( const t_species* cspec ) {
	t_species* spec = (t_species*) cspec;    /* the mirror is a cache of device state */
	zb_spec* e = zb_spec_of(spec, 0);
	if (e && e->device_init) zb_spec_to_device(spec);      /* never materialised yet: generate it, then mirror it */
	zb_guard_set(spec->part, ZB_G_RW);
	if (!e || !e->host_stale) return;
	if (e->slab.on) {
		/* every rank's mirror gets the whole population: own slab from the device (box coordinates), the
		   others' through the shared scratch area, in rank order */
		long long cnt[ZB_MAX_RANKS_LOCAL], mine = zdev_spec2d_np(zb_spec_dev(e));
		zb_par_allgather(&mine, sizeof mine, cnt);
		long long total = 0, off = 0;
		for (int r = 0; r < e->slab.nranks; r++) { if (r < e->slab.rank) off += cnt[r]; total += cnt[r]; }
		if (total > 0x7fffffffLL) {
			fprintf(stderr, "(*error*) zpic-b200: %lld particles do not fit the int-sized host mirror of the reference API\n", total);
			exit(-1);
		}
		t_part* all = zb_par_scratch((size_t) (total > 0 ? total : 1) * sizeof(t_part));
		t_part* dst = all + off;
		zdev_spec2d_download(zb_spec_dev(e), dst, mine);
		for (long long i = 0; i < mine; i++) dst[i].ix += e->slab.x0;
		zb_par_barrier();
		zb_spec_reserve(spec, (int) total);
		memcpy(spec->part, all, (size_t) total * sizeof(t_part));
		spec->np = (int) total;
		zb_par_barrier();
	} else {
		int64_t np = zdev_spec2d_np(zb_spec_dev(e));
		zb_spec_reserve(spec, (int) np);
		spec->np = (int) zdev_spec2d_download(zb_spec_dev(e), spec->part, spec->np_max);
	}
	e->host_stale = 0;
	e->part_seen = spec->part; e->np_seen = spec->np;
}

/* ---------------------------------------------------------------- guarded mirrors */

enum { ZB_K_E, ZB_K_B, ZB_K_EPART, ZB_K_BPART, ZB_K_J, ZB_K_PART };

/* a stale mirror was touched by host code (we are inside the fault handler): bring it over.  E and B travel
   together, like in zb_emf_to_host. */
static void guard_fill( void* owner, int kind ) {
	switch (kind) {
	case ZB_K_E: case ZB_K_B: case ZB_K_EPART: case ZB_K_BPART: zb_emf_to_host((const t_emf*) owner); break;
	case ZB_K_J: zb_cur_to_host((const t_current*) owner); break;
	case ZB_K_PART: zb_spec_to_host((const t_species*) owner); break;
	}
	zb_guard_refresh();
}

/* first host write to a mirror that was in sync: it goes up before the next device step */
static void guard_dirty( void* owner, int kind ) {
	if (kind == ZB_K_E || kind == ZB_K_B) { zb_grid* e = zb_grid_of_emf((const t_emf*) owner, 0); if (e) e->eb_dev_stale = 1; }
	else if (kind == ZB_K_PART) { zb_spec* e = zb_spec_of((const t_species*) owner, 0); if (e) e->dev_stale = 1; }
}

void zb_guard_bind_emf( const t_emf* emf ) {
	zb_guard_bind(emf->E_buf, (void*) emf, ZB_K_E, guard_fill, guard_dirty);
	zb_guard_bind(emf->B_buf, (void*) emf, ZB_K_B, guard_fill, guard_dirty);
	zb_guard_bind(emf->ext_fld.E_part_buf, (void*) emf, ZB_K_EPART, guard_fill, guard_dirty);
	zb_guard_bind(emf->ext_fld.B_part_buf, (void*) emf, ZB_K_BPART, guard_fill, guard_dirty);
}
void zb_guard_bind_cur( const t_current* cur ) { zb_guard_bind(cur->J_buf, (void*) cur, ZB_K_J, guard_fill, guard_dirty); }
void zb_guard_bind_spec( const t_species* spec, void* buf ) { zb_guard_bind(buf, (void*) spec, ZB_K_PART, guard_fill, guard_dirty); }

/* protection of every mirror from the coherence flags: device newer -> no access; in sync -> read only (a write
   is noticed); host newer -> open.  J is never uploaded, so it is simply open once it is current.  Particle
   mirrors are guarded only when the particle count is known after every step (not in lazy mode: the handler
   cannot move a buffer that turns out too small). */
void zb_guard_refresh( void ) {
	if (!zb_guard_enabled()) return;
	/* slabs: refreshing a mirror is a collective of the job (every rank's mirror holds the whole box), which a
	   page fault on one rank cannot start - the mirrors stay open and callers use the explicit sync calls */
	if (zb_par_init() > 1) return;
	for (int i = 0; i < n_grids; i++) {
		zb_grid* e = &grids[i];
		if (e->emf) {
			const int st = e->eb_host_stale ? ZB_G_NONE : (e->eb_dev_stale ? ZB_G_RW : ZB_G_READ);
			zb_guard_set(e->emf->E_buf, st); zb_guard_set(e->emf->B_buf, st);
			const int sp = e->part_host_stale ? ZB_G_NONE : ZB_G_READ;
			if (e->emf->ext_fld.E_type != EMF_FLD_TYPE_NONE) zb_guard_set(e->emf->ext_fld.E_part_buf, sp);
			if (e->emf->ext_fld.B_type != EMF_FLD_TYPE_NONE) zb_guard_set(e->emf->ext_fld.B_part_buf, sp);
		}
		if (e->cur) zb_guard_set(e->cur->J_buf, e->j_host_stale ? ZB_G_NONE : ZB_G_RW);
	}
	const int lazy = zb_opt_lazy();
	for (int i = 0; i < n_specs; i++) {
		zb_spec* e = &specs[i];
		if (!e->spec->part) continue;
		const int st = lazy ? ZB_G_RW : (e->host_stale ? ZB_G_NONE : (e->dev_stale ? ZB_G_RW : ZB_G_READ));
		zb_guard_set(e->spec->part, st);
	}
}

/* ---------------------------------------------------------------- public extras */

void zpic_b200_sync_host( t_simulation* sim ) {
	zb_emf_to_host(&sim->emf);
	zb_cur_to_host(&sim->current);
	for (int i = 0; i < sim->n_species; i++) zb_spec_to_host(&sim->species[i]);
	zb_guard_refresh();
}

void zpic_b200_touch_host( t_simulation* sim ) {
	zpic_b200_touch_emf(&sim->emf);
	for (int i = 0; i < sim->n_species; i++) zpic_b200_touch_species(&sim->species[i]);
}

void zpic_b200_touch_emf( t_emf* emf ) {
	zb_emf_to_host(emf);
	zb_grid* e = zb_grid_of_emf(emf, 1);
	e->eb_dev_stale = 1;
	zb_guard_refresh();
}

void zpic_b200_touch_species( t_species* spec ) {
	zb_spec_to_host(spec);
	zb_spec* e = zb_spec_of(spec, 1);
	e->dev_stale = 1;
	zb_guard_refresh();
}

void zpic_b200_sync_species( t_species* spec ) { zb_spec_to_host(spec); zb_guard_refresh(); }
void zpic_b200_sync_emf( t_emf* emf ) { zb_emf_to_host(emf); zb_guard_refresh(); }
void zpic_b200_sync_current( t_current* cur ) { zb_cur_to_host(cur); zb_guard_refresh(); }

/* device handles of the twins (zpic_dev.h objects), for tools that drive or time the
   device seam directly (bench.py) */
void* zpic_b200_species_handle( t_species* spec ) {
	zb_spec_to_device(spec);
	return zb_spec_dev(zb_spec_of(spec, 1));
}
void* zpic_b200_grid_handle( t_emf* emf ) { return zb_dev(zb_grid_of_emf(emf, 1)); }
