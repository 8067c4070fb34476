/* zpic-b200 :: host <-> device bookkeeping for the em1d API layer (internal; 1-D twin of
 * csrc/host/em2d/zb_state.h - registry of device twins keyed by host object address + coherence flags) */
#ifndef ZB_STATE_1D_H
#define ZB_STATE_1D_H

#include "zpic_dev.h"
#include "simulation.h"
#include "../common/zb_guard.h"

typedef struct zb_grid {
	const t_emf* emf;
	const t_current* cur;
	zdev_grid1d* g;            /* NULL until first device use */
	int nx;
	int eb_dev_stale, eb_host_stale, j_host_stale, part_host_stale;
	int mur_dev_stale;         /* host mur_fld / mur_tmp newer than the device copy */
} zb_grid;

typedef struct zb_spec {
	const t_species* spec;
	zdev_spec1d* d;
	int device_init;           /* generated on the device at first use: 1 counter-based, 2 the reference random stream */
	uint64_t device_seed;
	int* lat_lo; int* lat_hi;  /* 2: in-cell positions [lo, hi) of every cell that carry plasma */
	uint32_t rs_z, rs_w; int rs_have; double rs_spare;   /* 2: the stream's state where this species' draws begin */
	int dev_stale, host_stale;
	const t_part* part_seen;
	int np_seen;
} zb_spec;

zb_grid* zb_grid_of_emf( const t_emf* emf, int create );
zb_grid* zb_grid_of_cur( const t_current* cur, int create );
zdev_grid1d* zb_dev( zb_grid* e );
void zb_grid_pair( const t_emf* emf, const t_current* cur );
void zb_grid_drop_emf( const t_emf* emf );
void zb_grid_drop_cur( const t_current* cur );
zb_spec* zb_spec_of( const t_species* spec, int create );
zdev_spec1d* zb_spec_dev( zb_spec* e );
void zb_spec_drop( const t_species* spec );

void zb_emf_to_device( t_emf* emf );
void zb_emf_to_host( const t_emf* emf );
void zb_cur_to_host( const t_current* cur );
void zb_spec_to_device( t_species* spec );
void zb_spec_to_host( const t_species* spec );

int zb_opt_lazy( void );
int zb_opt_track_ids( void );
int zb_opt_coherent( void );
int zb_opt_device_init( void );

void spec_inject_into( t_species* spec, const int range[], t_part** buf, int* np, int* np_max );
void zb_spec_reserve( t_species* spec, const int size );

/* guarded mirrors (../common/zb_guard.h), as in the em2d layer */
void zb_guard_bind_emf( const t_emf* emf );
void zb_guard_bind_cur( const t_current* cur );
void zb_guard_bind_spec( const t_species* spec, void* buf );
void zb_guard_refresh( void );

#endif
