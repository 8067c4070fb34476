/* zpic-b200 :: em1d registry of device twins + mirror coherence (see zb_state.h) */
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

static zb_grid* grid_new( int nx ) {
	if (n_grids == ZB_MAX) { fprintf(stderr, "(*error*) zpic-b200: too many live field objects\n"); exit(-1); }
	zb_grid* e = &grids[n_grids++];
	memset(e, 0, sizeof(*e));
	e->nx = nx;
	return e;
}
zdev_grid1d* zb_dev( zb_grid* e ) { if (!e->g) e->g = zdev_grid1d_create(e->nx); return e->g; }

zb_grid* zb_grid_of_emf( const t_emf* emf, int create ) {
	for (int i = 0; i < n_grids; i++) if (grids[i].emf == emf) return &grids[i];
	if (!create) return NULL;
	zb_grid* e = grid_new(emf->nx);
	e->emf = emf; e->eb_dev_stale = 1; e->mur_dev_stale = 1;
	return e;
}
zb_grid* zb_grid_of_cur( const t_current* cur, int create ) {
	for (int i = 0; i < n_grids; i++) if (grids[i].cur == cur) return &grids[i];
	if (!create) return NULL;
	zb_grid* e = grid_new(cur->nx);
	e->cur = cur;
	return e;
}
static void grid_remove( zb_grid* e ) { if (e->g) zdev_grid1d_destroy(e->g); *e = grids[--n_grids]; }

void zb_grid_pair( const t_emf* emf, const t_current* cur ) {
	zb_grid* ge = zb_grid_of_emf(emf, 1);
	zb_grid* gc = zb_grid_of_cur(cur, 0);
	if (gc == ge) return;
	if (gc) { if (gc->emf) gc->cur = NULL; else grid_remove(gc); }
	ge = zb_grid_of_emf(emf, 1);
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
	e->spec = spec; e->dev_stale = 1;
	return e;
}
zdev_spec1d* zb_spec_dev( zb_spec* e ) {
	if (!e->d) e->d = zdev_spec1d_create(e->spec->nx, e->spec->ppc, zb_opt_track_ids());
	return e->d;
}
void zb_spec_drop( const t_species* spec ) {
	zb_spec* e = zb_spec_of(spec, 0);
	if (!e) return;
	if (e->d) zdev_spec1d_destroy(e->d);
	free(e->lat_lo); free(e->lat_hi);
	*e = specs[--n_specs];
}

void zb_emf_to_device( t_emf* emf ) {
	zb_grid* e = zb_grid_of_emf(emf, 1);
	if (e->eb_dev_stale) {
		zdev_grid1d_upload(zb_dev(e), ZDEV_E, (const float*) emf->E_buf);
		zdev_grid1d_upload(zb_dev(e), ZDEV_B, (const float*) emf->B_buf);
		e->eb_dev_stale = 0; e->eb_host_stale = 0;
	}
	if (e->mur_dev_stale) {
		float st[12];
		memcpy(st, emf->mur_fld, 6 * sizeof(float));
		memcpy(st + 6, emf->mur_tmp, 6 * sizeof(float));
		zdev_emf1d_set_mur(zb_dev(e), st);
		e->mur_dev_stale = 0;
	}
}
void zb_emf_to_host( const t_emf* cemf ) {
	t_emf* emf = (t_emf*) cemf;
	zb_grid* e = zb_grid_of_emf(emf, 0);
	if (!e) return;
	zb_guard_set(emf->E_buf, ZB_G_RW); zb_guard_set(emf->B_buf, ZB_G_RW);
	zb_guard_set(emf->ext_fld.E_part_buf, ZB_G_RW); zb_guard_set(emf->ext_fld.B_part_buf, ZB_G_RW);
	if (e->eb_host_stale) {
		zdev_grid1d_download(zb_dev(e), ZDEV_E, (float*) emf->E_buf);
		zdev_grid1d_download(zb_dev(e), ZDEV_B, (float*) emf->B_buf);
		float st[12];
		zdev_emf1d_get_mur(zb_dev(e), st);
		memcpy(emf->mur_fld, st, 6 * sizeof(float));
		memcpy(emf->mur_tmp, st + 6, 6 * sizeof(float));
		e->eb_host_stale = 0;
	}
	if (e->part_host_stale) {
		if (emf->ext_fld.E_type != EMF_FLD_TYPE_NONE && emf->ext_fld.E_part_buf)
			zdev_grid1d_download(zb_dev(e), ZDEV_EPART, (float*) emf->ext_fld.E_part_buf);
		if (emf->ext_fld.B_type != EMF_FLD_TYPE_NONE && emf->ext_fld.B_part_buf)
			zdev_grid1d_download(zb_dev(e), ZDEV_BPART, (float*) emf->ext_fld.B_part_buf);
		e->part_host_stale = 0;
	}
}
void zb_cur_to_host( const t_current* cur ) {
	zb_grid* e = zb_grid_of_cur(cur, 0);
	zb_guard_set(cur->J_buf, ZB_G_RW);
	if (!e || !e->j_host_stale) return;
	zdev_grid1d_download(zb_dev(e), ZDEV_J, (float*) cur->J_buf);
	e->j_host_stale = 0;
}
void zb_spec_to_device( t_species* spec ) {
	zb_spec* e = zb_spec_of(spec, 1);
	if (e->device_init == 2) {
		/* the reference's population, generated on the device from the state the stream had at spec_new */
		uint32_t z = e->rs_z, w = e->rs_w; int have = e->rs_have; double spare = e->rs_spare;
		if (zdev_spec1d_inject_lattice(zb_spec_dev(e), spec->ppc, spec->ufl, spec->uth, e->lat_lo, e->lat_hi, &z, &w, &have, &spare)) {
			fprintf(stderr, "(*error*) zpic-b200: device-side injection lost the random stream\n");
			exit(-1);
		}
		free(e->lat_lo); free(e->lat_hi); e->lat_lo = e->lat_hi = NULL;
		e->device_init = 0; e->dev_stale = 0; e->host_stale = 1;
		return;
	}
	if (e->device_init) {
		zdev_spec1d_inject_uniform(zb_spec_dev(e), spec->ppc, spec->ufl, spec->uth, e->device_seed);
		e->device_init = 0; e->dev_stale = 0; e->host_stale = 1;
		return;
	}
	if (!e->host_stale && (e->part_seen != spec->part || e->np_seen != spec->np)) e->dev_stale = 1;
	if (e->dev_stale) {
		zdev_spec1d_upload(zb_spec_dev(e), spec->part, spec->np);
		e->dev_stale = 0; e->host_stale = 0;
		e->part_seen = spec->part; e->np_seen = spec->np;
	}
}
void zb_spec_to_host( const t_species* cspec ) {
	t_species* spec = (t_species*) cspec;
	zb_spec* e = zb_spec_of(spec, 0);
	if (e && e->device_init) zb_spec_to_device(spec);      /* never materialised yet: generate it, then mirror it */
	zb_guard_set(spec->part, ZB_G_RW);
	if (!e || !e->host_stale) return;
	int64_t np = zdev_spec1d_np(zb_spec_dev(e));
	zb_spec_reserve(spec, (int) np);
	spec->np = (int) zdev_spec1d_download(zb_spec_dev(e), spec->part, spec->np_max);
	e->host_stale = 0;
	e->part_seen = spec->part; e->np_seen = spec->np;
}

/* ---- guarded mirrors: see csrc/host/em2d/zb_state.c */
enum { ZB_K_E, ZB_K_B, ZB_K_EPART, ZB_K_BPART, ZB_K_J, ZB_K_PART };

static void guard_fill( void* owner, int kind ) {
	switch (kind) {
	case ZB_K_E: case ZB_K_B: case ZB_K_EPART: case ZB_K_BPART: zb_emf_to_host((const t_emf*) owner); break;
	case ZB_K_J: zb_cur_to_host((const t_current*) owner); break;
	case ZB_K_PART: zb_spec_to_host((const t_species*) owner); break;
	}
	zb_guard_refresh();
}
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

void zb_guard_refresh( void ) {
	if (!zb_guard_enabled()) return;
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
		zb_guard_set(e->spec->part, lazy ? ZB_G_RW : (e->host_stale ? ZB_G_NONE : (e->dev_stale ? ZB_G_RW : ZB_G_READ)));
	}
}

void zpic_b200_sync_host( t_simulation* sim ) {
	zb_emf_to_host(&sim->emf);
	zb_cur_to_host(&sim->current);
	for (int i = 0; i < sim->n_species; i++) zb_spec_to_host(&sim->species[i]);
	zb_guard_refresh();
}
void zpic_b200_touch_emf( t_emf* emf ) {
	zb_emf_to_host(emf);
	zb_grid* e = zb_grid_of_emf(emf, 1);
	e->eb_dev_stale = 1; e->mur_dev_stale = 1;
	zb_guard_refresh();
}
void zpic_b200_touch_species( t_species* spec ) {
	zb_spec_to_host(spec);
	zb_spec_of(spec, 1)->dev_stale = 1;
	zb_guard_refresh();
}
void zpic_b200_touch_host( t_simulation* sim ) {
	zpic_b200_touch_emf(&sim->emf);
	for (int i = 0; i < sim->n_species; i++) zpic_b200_touch_species(&sim->species[i]);
}
void zpic_b200_sync_species( t_species* spec ) { zb_spec_to_host(spec); zb_guard_refresh(); }
void zpic_b200_sync_emf( t_emf* emf ) { zb_emf_to_host(emf); zb_guard_refresh(); }
void zpic_b200_sync_current( t_current* cur ) { zb_cur_to_host(cur); zb_guard_refresh(); }
void* zpic_b200_species_handle( t_species* spec ) { zb_spec_to_device(spec); return zb_spec_dev(zb_spec_of(spec, 1)); }
void* zpic_b200_grid_handle( t_emf* emf ) { return zb_dev(zb_grid_of_emf(emf, 1)); }
