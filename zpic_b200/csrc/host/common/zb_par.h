/* zpic-b200 :: the ranks of one slab-decomposed run (one process per GPU, one node) - host side.
 *
 * A run is SPMD: every rank executes the same deck / the same API calls (so the reference's global random
 * stream stays in step everywhere); rank r owns the cell columns [r*nx/N, (r+1)*nx/N) on its own GPU.
 * This module is the small host-side substrate the API layer needs for that: who am I, a barrier, sums and
 * gathers for DIAGNOSTICS (energies, particle counts, reports), and the exchange of the CUDA IPC handles of
 * the device mailboxes.  It never carries time-step data: guard cells and migrating particles travel from GPU
 * to GPU over NVLink, written by the sending kernels straight into the neighbour's memory (zdev_slab.cuh).
 *
 * Transport: one POSIX shared-memory segment per job (all ranks are on one node by construction).
 * Ranks come from ZPIC_RANK / ZPIC_NRANKS or, under torchrun, RANK / WORLD_SIZE; the job key from ZPIC_JOB or
 * MASTER_PORT.  ZPIC_SLABS=0 turns the decomposition off (every rank then runs the whole box).
 */
#ifndef ZB_PAR_H
#define ZB_PAR_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* attach to the job (idempotent); returns the number of ranks (1: no decomposition) */
int zb_par_init( void );
int zb_par_rank( void );
int zb_par_nranks( void );
void zb_par_barrier( void );

/* in-place sums over all ranks (every rank gets the result); n <= 1024 */
void zb_par_allreduce_sum_d( double* v, int n );
void zb_par_allreduce_sum_ll( long long* v, int n );
/* every rank contributes `bytes` (<= 1024) bytes; all[r*bytes ..] = contribution of rank r */
void zb_par_allgather( const void* mine, size_t bytes, void* all );

/* A shared scratch area of at least `bytes`, the same memory on every rank (collective: all ranks call it with
   the same size; the pointer stays valid until the next call with a larger size).  Callers bracket their use
   with zb_par_barrier(). */
void* zb_par_scratch( size_t bytes );

/* sum of float arrays over all ranks through the scratch area, result in every rank's `v` */
void zb_par_allreduce_sum_f( float* v, size_t n );

void zb_par_finalize( void );

#ifdef __cplusplus
}
#endif
#endif
