/* zpic-b200 :: registry of device twins + mirror coherence (see zb_state.h) */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "zb_state.h"
#include "zpic_b200.h"

#define ZB_MAX 256
static zb_grid grids[ZB_MAX];
static int n_grids = 0;
static zb_spec specs[ZB_MAX];
static int n_specs = 0;

static int opt_lazy = -1, opt_ids = -1, opt_coherent = -1, opt_devinit = -1;
static int env_flag( const char* name ) { const char* e = getenv(name); return e && atoi(e) != 0; }
int zb_opt_lazy( void ) { if (opt_lazy < 0) opt_lazy = env_flag("ZPIC_LAZY"); return opt_lazy; }
int zb_opt_track_ids( void ) { if (opt_ids < 0) opt_ids = env_flag("ZPIC_TRACK_IDS"); return opt_ids; }
int zb_opt_coherent( void ) { if (opt_coherent < 0) opt_coherent = env_flag("ZPIC_COHERENT"); return opt_coherent; }
int zb_opt_device_init( void ) { if (opt_devinit < 0) opt_devinit = env_flag("ZPIC_DEVICE_INIT"); return opt_devinit; }

void zpic_b200_set_option( const char* name, int value ) {
	if (!strcmp(name, "lazy")) opt_lazy = value;
	else if (!strcmp(name, "track_ids")) opt_ids = value;
	else if (!strcmp(name, "coherent")) opt_coherent = value;
	else if (!strcmp(name, "device_init")) opt_devinit = value;
	else fprintf(stderr, "(*warning*) zpic_b200_set_option: unknown option %s\n", name);
}

static zb_grid* grid_new( int nx, int ny ) {
	if (n_grids == ZB_MAX) { fprintf(stderr, "(*error*) zpic-b200: too many live field objects\n"); exit(-1); }
	zb_grid* e = &grids[n_grids++];
	memset(e, 0, sizeof(*e));
	e->nx = nx; e->ny = ny;      /* the device object is created on first device use */
	return e;
}

zdev_grid2d* zb_dev( zb_grid* e ) {
	if (!e->g) e->g = zdev_grid2d_create(e->nx, e->ny);
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
		e->d = zdev_spec2d_create(spec->nx[0], spec->nx[1], spec->ppc[0] * spec->ppc[1], zb_opt_track_ids());
	}
	return e->d;
}

void zb_spec_drop( const t_species* spec ) {
	zb_spec* e = zb_spec_of(spec, 0);
	if (!e) return;
	if (e->d) zdev_spec2d_destroy(e->d);
	*e = specs[--n_specs];
}

/* ---------------------------------------------------------------- coherence */

/* the E, B mirrors are transferred again and again: page-lock them once (unpinned by emf_delete) */
static void zb_pin_emf( const t_emf* emf ) {
	const size_t bytes = (size_t) (emf->nx[0] + 3) * (emf->nx[1] + 3) * sizeof(float3);
	zdev_host_pin(emf->E_buf, bytes);
	zdev_host_pin(emf->B_buf, bytes);
}

void zb_emf_to_device( t_emf* emf ) {
	zb_grid* e = zb_grid_of_emf(emf, 1);
	if (e->eb_dev_stale) {
		zb_pin_emf(emf);
		zdev_grid2d_upload(zb_dev(e), ZDEV_E, (const float*) emf->E_buf);
		zdev_grid2d_upload(zb_dev(e), ZDEV_B, (const float*) emf->B_buf);
		e->eb_dev_stale = 0;
		e->eb_host_stale = 0;
	}
}

void zb_emf_to_host( const t_emf* emf ) {
	zb_grid* e = zb_grid_of_emf(emf, 0);
	if (!e) return;
	if (e->eb_host_stale) {
		zb_pin_emf(emf);
		zdev_grid2d_download(zb_dev(e), ZDEV_E, (float*) emf->E_buf);
		zdev_grid2d_download(zb_dev(e), ZDEV_B, (float*) emf->B_buf);
		e->eb_host_stale = 0;
	}
	if (e->part_host_stale) {
		if (emf->ext_fld.E_type != EMF_FLD_TYPE_NONE && emf->ext_fld.E_part_buf)
			zdev_grid2d_download(zb_dev(e), ZDEV_EPART, (float*) emf->ext_fld.E_part_buf);
		if (emf->ext_fld.B_type != EMF_FLD_TYPE_NONE && emf->ext_fld.B_part_buf)
			zdev_grid2d_download(zb_dev(e), ZDEV_BPART, (float*) emf->ext_fld.B_part_buf);
		e->part_host_stale = 0;
	}
}

void zb_cur_to_host( const t_current* cur ) {
	zb_grid* e = zb_grid_of_cur(cur, 0);
	if (!e || !e->j_host_stale) return;
	zdev_host_pin(cur->J_buf, (size_t) (cur->nx[0] + 3) * (cur->nx[1] + 3) * sizeof(float3));   /* unpinned by current_delete */
	zdev_grid2d_download(zb_dev(e), ZDEV_J, (float*) cur->J_buf);
	e->j_host_stale = 0;
}

void zb_spec_to_device( t_species* spec ) {
	zb_spec* e = zb_spec_of(spec, 1);
	if (e->device_init) {
		/* throughput configurations: the population never existed on the host */
		zdev_spec2d_inject_uniform(zb_spec_dev(e), spec->ppc[0], spec->ppc[1], spec->ufl, spec->uth, e->device_seed);
		e->device_init = 0; e->dev_stale = 0; e->host_stale = 1;
		return;
	}
	/* host code that appended particles or reallocated the buffer did so on a current
	   mirror (the Python layer syncs first); take the host copy as the truth then */
	if (!e->host_stale && (e->part_seen != spec->part || e->np_seen != spec->np)) e->dev_stale = 1;
	if (e->dev_stale) {
		zdev_spec2d_upload(zb_spec_dev(e), spec->part, spec->np);
		e->dev_stale = 0; e->host_stale = 0;
		e->part_seen = spec->part; e->np_seen = spec->np;
	}
}

void zb_spec_to_host( const t_species* cspec ) {
	t_species* spec = (t_species*) cspec;    /* the mirror is a cache of device state */
	zb_spec* e = zb_spec_of(spec, 0);
	if (!e || !e->host_stale) return;
	int64_t np = zdev_spec2d_np(zb_spec_dev(e));
	spec_grow_buffer(spec, (int) np);
	spec->np = (int) zdev_spec2d_download(zb_spec_dev(e), spec->part, spec->np_max);
	e->host_stale = 0;
	e->part_seen = spec->part; e->np_seen = spec->np;
}

/* ---------------------------------------------------------------- public extras */

void zpic_b200_sync_host( t_simulation* sim ) {
	zb_emf_to_host(&sim->emf);
	zb_cur_to_host(&sim->current);
	for (int i = 0; i < sim->n_species; i++) zb_spec_to_host(&sim->species[i]);
}

void zpic_b200_touch_host( t_simulation* sim ) {
	zpic_b200_touch_emf(&sim->emf);
	for (int i = 0; i < sim->n_species; i++) zpic_b200_touch_species(&sim->species[i]);
}

void zpic_b200_touch_emf( t_emf* emf ) {
	zb_emf_to_host(emf);
	zb_grid* e = zb_grid_of_emf(emf, 1);
	e->eb_dev_stale = 1;
}

void zpic_b200_touch_species( t_species* spec ) {
	zb_spec_to_host(spec);
	zb_spec* e = zb_spec_of(spec, 1);
	e->dev_stale = 1;
}

void zpic_b200_sync_species( t_species* spec ) { zb_spec_to_host(spec); }
void zpic_b200_sync_emf( t_emf* emf ) { zb_emf_to_host(emf); }
void zpic_b200_sync_current( t_current* cur ) { zb_cur_to_host(cur); }

/* device handles of the twins (zpic_dev.h objects), for tools that drive or time the
   device seam directly (bench.py) */
void* zpic_b200_species_handle( t_species* spec ) {
	zb_spec_to_device(spec);
	return zb_spec_dev(zb_spec_of(spec, 1));
}
void* zpic_b200_grid_handle( t_emf* emf ) { return zb_dev(zb_grid_of_emf(emf, 1)); }
