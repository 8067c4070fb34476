/* zpic-b200 :: host <-> device bookkeeping for the em2d API layer (internal).
 *
 * The public structs (include/em2d) carry no device members; the device twin of
 * every t_emf / t_current / t_species is found through a small registry keyed by
 * the host object's address.  Each entry also tracks which side holds the current
 * data so mirrors are only copied when somebody needs them (SURVEY.md 8b,
 * "coherence contract").
 */
#ifndef ZB_STATE_H
#define ZB_STATE_H

#include "zpic_dev.h"
#include "simulation.h"
#include "../common/zb_par.h"
#include "../common/zb_guard.h"

/* Slab decomposition along x (one process per GPU, zb_par.h).  The host objects of the API stay GLOBAL on every
 * rank - same sizes, same mirrors, same random stream as a single-process run; the device twins are the rank's
 * slab: columns [x0, x0 + nxl) with the reference's guard layout as halo.  Periodic boxes form a ring; under a
 * moving window the slabs form an open chain whose two physical edges do nothing (em2d/emf.c:581, current.c:124). */
typedef struct zb_slab {
	int on;                    /* 0: single domain */
	int rank, nranks;
	int nxl, x0;               /* width and first column of this rank's slab */
	int left, right;           /* neighbour ranks, -1: none */
	int wrap_left, wrap_right; /* that edge is the periodic box boundary */
	int is_last;
} zb_slab;
zb_slab zb_slab_make( int nx_global, int moving_window );

/* E/B/J device grids shared by one t_emf and (after sim_new) one t_current */
typedef struct zb_grid {
	const t_emf* emf;          /* owner keys, either may be NULL */
	const t_current* cur;
	zdev_grid2d* g;            /* NULL until first device use: see zb_dev() */
	int nx, ny;
	int eb_dev_stale;          /* host E_buf/B_buf were modified after the last upload */
	int eb_host_stale;         /* device advanced E/B after the last download */
	int j_host_stale;          /* device J newer than host J_buf */
	int part_host_stale;       /* device E_part/B_part newer than the host *_part_buf */
	zb_slab slab;              /* fixed when the device object is created */
} zb_grid;

typedef struct zb_spec {
	const t_species* spec;
	zdev_spec2d* d;            /* NULL until first device use: see zb_spec_dev() */
	int device_init;           /* population is generated on the device at first use (no host mirror yet): 1 with a
	                              counter-based generator, 2 on the reference random stream (lattice profiles) */
	int* lat_lo; int* lat_hi;  /*   2: in-cell x positions [lo, hi) of every box column that carry plasma */
	uint32_t rs_z, rs_w; int rs_have; double rs_spare;   /* 2: the stream's state where this species' draws begin */
	int device_made;           /* ... and was: the moving window's new columns are generated on the device too */
	uint64_t device_seed;
	int dev_rect[4];           /*   ... in the cells [0,1) x [2,3) of the box */
	int dev_stale;             /* host part[] newer than the device copy (or never uploaded) */
	int host_stale;            /* device newer than host part[] */
	const t_part* part_seen;   /* host buffer address / count at the last transfer: a change */
	int np_seen;               /*   means host code touched the buffer (Species.add, realloc) */
	zb_slab slab;              /* fixed when the device object is created */
} zb_spec;

zb_grid* zb_grid_of_emf( const t_emf* emf, int create );
zb_grid* zb_grid_of_cur( const t_current* cur, int create );
/* device grid object of an entry, created on demand */
zdev_grid2d* zb_dev( zb_grid* e );
/* make emf and current share one device grid object (done by sim_new) */
void zb_grid_pair( const t_emf* emf, const t_current* cur );
void zb_grid_drop_emf( const t_emf* emf );
void zb_grid_drop_cur( const t_current* cur );

zb_spec* zb_spec_of( const t_species* spec, int create );
zdev_spec2d* zb_spec_dev( zb_spec* e );
void zb_spec_drop( const t_species* spec );

/* bring one side up to date */
void zb_emf_to_device( t_emf* emf );       /* upload E/B if the host copy is newer */
void zb_emf_to_host( const t_emf* emf );   /* download E/B (and *_part_buf) if the device copy is newer */
void zb_cur_to_host( const t_current* cur );
void zb_spec_to_device( t_species* spec );
void zb_spec_to_host( const t_species* spec );

/* options (environment: ZPIC_LAZY, ZPIC_TRACK_IDS, ZPIC_COHERENT) */
int zb_opt_lazy( void );        /* 1: do not fetch energy / np after every spec_advance */
int zb_opt_track_ids( void );   /* 1: keep injection order recoverable in the host mirror */
int zb_opt_coherent( void );    /* 1: host mirrors are refreshed before and after every sim_iter */
int zb_opt_device_init( void ); /* species are initialised on the device: 1 counter-based generator (not the reference
                                   random stream), 2 the reference random stream (lattice profiles; others on the host) */

void spec_inject_into( t_species* spec, const int range[][2], t_part** buf, int* np, int* np_max );
void zb_spec_reserve( t_species* spec, const int size );

/* guarded mirrors (../common/zb_guard.h): tell the guard who owns a mirror, and bring every mirror's protection
   in line with the coherence flags (called where a public entry point returns to the caller) */
void zb_guard_bind_emf( const t_emf* emf );
void zb_guard_bind_cur( const t_current* cur );
void zb_guard_bind_spec( const t_species* spec, void* buf );
void zb_guard_refresh( void );

#endif
