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

/* A custom-field callback (t_emf_ext_fld / t_emf_init_fld: E_custom, B_custom with *_custom_data) for callers that
 * cannot hand over a C function - Python through ctypes cannot return a struct from a callback: `data` points at a
 * zpic_b200_field_table whose values the caller computed for every cell of the buffer, guards included
 * (value of cell (ix, iy) at table[3 * ((ix + 1) + (iy + 1) * nrow)], nrow = nx + 3; the em1d library exports the same
 * name with the em1d callback signature (int ix, float dx, void* data), value of cell ix at table[3 * (ix + 1)]). */
typedef struct zpic_b200_field_table { int nrow; const float* table; } zpic_b200_field_table;
#ifdef ZPIC_B200_WITH_FLOAT3          /* (declared where the reference's float3 is known: include/em2d/zpic.h) */
float3 zpic_b200_table_field( int ix, float dx, int iy, float dy, void* data );
#endif

/* opaque device handles (zdev_spec2d* / zdev_grid2d* of include/zpic_dev.h) behind a host object */
void* zpic_b200_species_handle( struct Species* spec );
void* zpic_b200_grid_handle( struct EMF* emf );

#ifdef __cplusplus
}
#endif
#endif
