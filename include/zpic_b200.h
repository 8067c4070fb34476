/* zpic-b200 :: extensions to the reference API (host side).
 *
 * The reference keeps all state in host memory, so callers read and write
 * spec->part, emf->E_buf, current->J_buf ... freely between sim_iter() calls
 * (SURVEY.md 3.3).  Here the device copy is authoritative while stepping; every
 * reference entry point that consumes host data (*_report, emf_get_energy,
 * spec_deposit_*, sim_report_energy) refreshes the mirror it needs by itself.
 * Code that touches the raw buffers directly brackets the access with the calls
 * below (the Python module does so inside its property getters).
 */
#ifndef ZPIC_B200_H
#define ZPIC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

struct Simulation; struct EMF; struct Current; struct Species;

/* make every host mirror of the simulation current (device -> host where needed) */
void zpic_b200_sync_host( struct Simulation* sim );
void zpic_b200_sync_species( struct Species* spec );
void zpic_b200_sync_emf( struct EMF* emf );
void zpic_b200_sync_current( struct Current* cur );
/* declare that host code is about to modify / has modified the raw buffers: the
 * mirror is refreshed first and re-uploaded before the next device step */
void zpic_b200_touch_host( struct Simulation* sim );
void zpic_b200_touch_species( struct Species* spec );
void zpic_b200_touch_emf( struct EMF* emf );
/* "lazy" (skip the per-step fetch of energy / particle count: fully asynchronous
 * stepping), "track_ids" (host mirror keeps injection order), "coherent" (mirrors
 * refreshed around every sim_iter: strict drop-in semantics, host<->device copies
 * every step).  Also settable through ZPIC_LAZY / ZPIC_TRACK_IDS / ZPIC_COHERENT. */
void zpic_b200_set_option( const char* name, int value );

/* opaque device handles (zdev_spec2d* / zdev_grid2d* of include/zpic_dev.h) behind a host object */
void* zpic_b200_species_handle( struct Species* spec );
void* zpic_b200_grid_handle( struct EMF* emf );

#ifdef __cplusplus
}
#endif
#endif
