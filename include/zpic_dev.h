/* zpic-b200 :: the device seam (C ABI).
 *
 * Everything the host C layer (include/em2d, include/em1d) asks of the GPU goes
 * through the plain-C entry points below: opaque handles, plain pointers and
 * sizes, no C++ or torch types.  Each entry point names the reference routine it
 * replaces (paths relative to the reference tree).  The implementations are
 * hand-written sm_100a CUDA kernels (zpic_b200/csrc/dev); there is no CPU
 * fallback: if no CUDA device can be initialised zdev_init() returns non-zero
 * and the host layer aborts.
 *
 * Conventions
 *  - all grids use the reference guard-cell geometry: (nx+3) x (ny+3) float3
 *    cells, 1 lower / 2 upper guards, row stride nrow = nx+3 (em2d/emf.c:59-89,
 *    em2d/current.c:33-53).  Host buffers passed in/out are the *_buf arrays
 *    (guards included).
 *  - particle records exchanged with the host are the 28-byte AoS t_part
 *    (em2d/particles.h:29-37); on the device they live as tile-binned SoA.
 *  - all work is enqueued on one per-process stream; calls returning data to the
 *    host synchronise that stream, the others do not.
 *  - errors: CUDA failures print a message to stderr and exit(-1), the
 *    reference's own convention for fatal errors (em2d/simulation.c:106-110).
 */
#ifndef ZPIC_DEV_H
#define ZPIC_DEV_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ runtime */

/* Select the CUDA device (ordinal, or -1 = $ZPIC_DEVICE / $LOCAL_RANK / 0) and
 * create the stream.  Idempotent.  Returns 0 on success. */
int  zdev_init( int device );
/* 1 once zdev_init succeeded */
int  zdev_ready( void );
/* block until all enqueued device work finished */
void zdev_sync( void );
/* the cudaStream_t used for every launch (so callers can record events on it) */
void* zdev_stream( void );
/* number of kernels launched by this library since load (bench "gpu_launches") */
uint64_t zdev_launch_count( void );
/* device-side timers: cudaEvent pairs on the library stream */
void* zdev_event_create( void );
void  zdev_event_record( void* ev );
float zdev_event_elapsed_ms( void* start, void* stop );   /* synchronises on stop */
void  zdev_event_destroy( void* ev );
/* free / total device memory in bytes */
void zdev_mem_info( size_t* free_b, size_t* total_b );
/* enqueue all further work on a caller-owned cudaStream_t (NULL: back to the library stream) */
void zdev_set_stream( void* stream );
/* write `bytes` of scratch (>= L2 size) to evict the L2 between timed steps */
void zdev_flush_l2( void );
/* page-lock a host buffer that is transferred repeatedly (>= 1 MB; done once per buffer, silently skipped if it
 * cannot be locked); the owner calls zdev_host_unpin before freeing it */
void zdev_host_pin( const void* host_ptr, size_t bytes );
void zdev_host_unpin( const void* host_ptr );

/* ---------------------------------------------------------- em2d grids (E,B,J) */

typedef struct zdev_grid2d zdev_grid2d;

enum zdev_fld { ZDEV_E = 0, ZDEV_B = 1, ZDEV_J = 2, ZDEV_EPART = 3, ZDEV_BPART = 4 };

/* emf_new / current_new: a grid object holds zeroed E,B and/or J device buffers, each
 * allocated on first use, so one object can back a t_emf, a t_current or (as sim_new
 * arranges) both (em2d/emf.c:56-113, current.c:30-79).  Entry points that combine
 * fields and currents take the field grid `g` and the current grid `g_cur`, which
 * may be the same object. */
zdev_grid2d* zdev_grid2d_create( int nx, int ny );
/* emf_delete + current_delete (em2d/emf.c:123-142, current.c:86-91) */
void zdev_grid2d_destroy( zdev_grid2d* g );
/* host mirror -> device / device -> host mirror, whole buffer incl. guards */
void zdev_grid2d_upload( zdev_grid2d* g, int which, const float* host_buf );
void zdev_grid2d_download( zdev_grid2d* g, int which, float* host_buf );
/* the same for a grid that is one slab of a wider box: the slab's window (guards included) of a host buffer
 * whose rows are host_nrow cells long; local buffer column c = host buffer column x0 + c */
void zdev_grid2d_upload_window( zdev_grid2d* g, int which, const float* host_buf, int host_nrow, int x0 );
void zdev_grid2d_download_window( zdev_grid2d* g, int which, float* host_buf, int host_nrow, int x0 );
/* Slab decomposition along x, one process per GPU (csrc/dev/zdev_slab.cuh): this grid is the slab
 * between the ranks `left` and `right` (-1: none - the open end of a moving-window chain).  Collective over the
 * ranks of the job (zb_par.h).  From then on zdev_current_update and zdev_emf_advance exchange the guard columns
 * with the neighbour slabs - kernels that store straight into the neighbour's memory over NVLink and flag the
 * message, no host synchronisation, no library collective.  wrap_*: that edge is the periodic box boundary
 * (the reference leaves guard columns un-filtered by kernel_y there, em2d/current.c:382-411). */
void zdev_grid2d_set_slab( zdev_grid2d* g, int left, int right, int wrap_left, int wrap_right );
/* raw device pointer of a grid buffer (cell [-1,-1]); for tests / multi-GPU glue */
float* zdev_grid2d_ptr( zdev_grid2d* g, int which );

/* current_zero (em2d/current.c:98-107) */
void zdev_current_zero( zdev_grid2d* g );
/* current_update = current_update_gc + current_smooth, iter++ left to the host
 * (em2d/current.c:118-183, 297-459).  xtype/ytype: 0 none, 1 binomial, 2 compensated.
 * Reproduces the reference quirk that the y passes are counted with xlevel
 * (current.c:449). */
void zdev_current_update( zdev_grid2d* g, int moving_window,
                          int xtype, int ytype, int xlevel, int ylevel );
/* the two halves separately (multi-GPU inserts halo exchanges between them) */
void zdev_current_update_gc( zdev_grid2d* g, int moving_window );
void zdev_current_smooth( zdev_grid2d* g, int moving_window,
                          int xtype, int ytype, int xlevel, int ylevel );

/* uniform external fields: E_part = E + E0 on every cell incl. guards
 * (em2d/emf.c:838-914, UNIFORM branch).  type 0 disables (E_part aliases E). */
void zdev_emf_set_ext_uniform( zdev_grid2d* g, int e_on, const float e0[3],
                               int b_on, const float b0[3] );
/* custom external fields evaluated once on the host (callbacks are time independent,
 * SURVEY.md App. D): host_ext is a full (nx+3)x(ny+3) float3 buffer or NULL */
void zdev_emf_set_ext_grid( zdev_grid2d* g, const float* host_ext_e, const float* host_ext_b );

/* emf_advance minus iter++ (em2d/emf.c:688-716): yee_b(dt/2), yee_e(dt), yee_b(dt/2),
 * emf_update_gc, emf_update_part_fld, and - when shift_window != 0 -
 * emf_move_window's left shift + zeroing (emf.c:648-675).  The host decides
 * shift_window with the reference's float test (emf.c:650). */
void zdev_emf_advance( zdev_grid2d* g, zdev_grid2d* g_cur, float dt, float dx, float dy,
                       int moving_window, int shift_window );
/* The three stencils of emf_advance run as ONE kernel by default (60 B instead of 132 B of traffic per cell,
 * bit-identical results): 2 (default) marches the rows through registers, 1 works on shared-memory tiles, 0 selects
 * the three separate kernels below ($ZPIC_FUSED_YEE=0 / 1 / 2 does the same). */
void zdev_yee_set_fused( int on );
/* pieces, for kernel-level parity tests */
void zdev_yee_b( zdev_grid2d* g, float dt_dx, float dt_dy );            /* emf.c:500-522 */
void zdev_yee_e( zdev_grid2d* g, zdev_grid2d* g_cur, float dt_dx, float dt_dy, float dt );  /* emf.c:531-562 */
void zdev_emf_update_gc( zdev_grid2d* g, int moving_window );           /* emf.c:573-637 */
void zdev_emf_move_window( zdev_grid2d* g );                            /* emf.c:659-670 */
/* ---- pieces for the slab decomposition (halo exchange between the steps) ----
 * pack / unpack: cell columns [i0, i0+ncols) x rows [j0, j0+nrows) of one grid <-> a dense device
 * buffer of float3 [row][col]; add != 0 accumulates (the J guard fold across slabs,
 * em2d/current.c:128-131), else overwrites (guard refresh, em2d/emf.c:583-607). */
void zdev_grid2d_pack_cols( zdev_grid2d* g, int which, int i0, int ncols, int j0, int nrows, float* dev_out );
void zdev_grid2d_unpack_cols( zdev_grid2d* g, int which, int i0, int ncols, int j0, int nrows, const float* dev_in, int add );
/* y half of current_update_gc (em2d/current.c:142-157) */
void zdev_current_fold_y( zdev_grid2d* g );
/* pass list of current_smooth (em2d/current.c:427-459) and one [sa,sb,sa] pass; keep_x_guards = the
 * moving-window rule (x guards not refreshed, current.c:346) - slabs refresh them by exchange */
int  zdev_smooth_plan( int xtype, int ytype, int xlevel, int ylevel, int* dirs, float* sa, float* sb );
void zdev_smooth_pass( zdev_grid2d* g, int dir, float sa, float sb, int keep_x_guards );
/* emf_move_window's shift; zero_right = 0 for slabs whose right edge is interior (em2d/emf.c:659-670) */
void zdev_emf_shift( zdev_grid2d* g, int zero_right );
void zdev_emf_update_part_fld( zdev_grid2d* g );                        /* emf.c:838-914 */
/* emf_get_energy: 6 interior sums of squares in double, scaled by 0.5*dx*dy by the
 * caller (em2d/emf.c:729-750).  Synchronises. */
void zdev_emf_energy( zdev_grid2d* g, double sums[6] );

/* ---------------------------------------------------------- em2d particles */

typedef struct zdev_spec2d zdev_spec2d;

/* per-step scalars, computed on the host exactly as the reference does
 * (em2d/particles.c:1111-1117) */
typedef struct zdev_push2d_params {
	float tem;      /* (float)(0.5*dt/m_q), double arithmetic on the host */
	float dt_dx;    /* dt/dx[0] */
	float dt_dy;    /* dt/dx[1] */
	float qnx;      /* q*dx[0]/dt */
	float qny;      /* q*dx[1]/dt */
	float q;        /* charge per particle */
	int   moving_window;  /* 1: absorbing x, periodic y (particles.c:1237-1251) */
	int   shift_window;   /* 1: all ix-- this step (particles.c:619-632) */
	int   slab_left;      /* 1: the lower x edge is a slab boundary: leavers are exported, not wrapped/absorbed */
	int   slab_right;     /* 1: same for the upper x edge */
} zdev_push2d_params;

/* Device species for an nx x ny grid.  ppc_hint = expected particles per cell
 * (sizes the tiles and their capacity), track_ids != 0 carries the injection index
 * of every particle so the host mirror can be restored in reference order. */
zdev_spec2d* zdev_spec2d_create( int nx, int ny, int ppc_hint, int track_ids );
void zdev_spec2d_destroy( zdev_spec2d* s );
/* host AoS -> device tiles (replaces the whole population) */
void zdev_spec2d_upload( zdev_spec2d* s, const void* part_aos, int64_t np );
/* With ZPIC_ZERO_COPY_MIN=<bytes> set, host buffers of at least that size passed to upload / download are
 * pinned + mapped on first use and then accessed by the kernels directly (zero copy, off by default); call
 * this before such a buffer is freed or reallocated. */
void zdev_host_forget( const void* host_ptr );
/* append host AoS particles (moving-window injection, Species.add) */
void zdev_spec2d_append( zdev_spec2d* s, const void* part_aos, int64_t np );
/* device tiles -> host AoS; returns the number written (<= max_np).  With
 * track_ids and a never-absorbed population the order is the injection order. */
int64_t zdev_spec2d_download( zdev_spec2d* s, void* part_aos, int64_t max_np );
/* current particle count (synchronises) */
int64_t zdev_spec2d_np( zdev_spec2d* s );
/* uniform-density, thermal+fluid initialisation ON the device with a counter-based
 * generator: same distribution and per-cell mean removal as spec_set_x/spec_set_u
 * (em2d/particles.c:85-149, 159-354) but NOT the same random stream.  Used for the
 * throughput configurations whose host mirrors would not fit (hard part 5). */
void zdev_spec2d_inject_uniform( zdev_spec2d* s, int ppcx, int ppcy,
                                 const float ufl[3], const float uth[3], uint64_t seed );
/* the same in the cells [ix0, ix1) x [iy0, iy1) only (a plasma that starts at some x, a species that fills a band) */
void zdev_spec2d_inject_rect( zdev_spec2d* s, int ppcx, int ppcy,
                              const float ufl[3], const float uth[3], uint64_t seed, int ix0, int ix1, int iy0, int iy1 );
/* the moving window's new column (cells ix, iy0 <= iy < iy1) generated the same way and appended on the device
 * (species that were initialised on the device; the others take the host injector, em2d/particles.c:619-640) */
void zdev_spec2d_inject_column( zdev_spec2d* s, int ppcx, int ppcy, const float ufl[3], const float uth[3], uint64_t seed,
                                int ix, int iy0, int iy1, uint64_t column_number );
/* the same in the rows iy0 <= iy < iy1 only (half-box species of a shear-flow deck) */
void zdev_spec2d_inject_band( zdev_spec2d* s, int ppcx, int ppcy,
                              const float ufl[3], const float uth[3], uint64_t seed, int iy0, int iy1 );

/* The reference's OWN random stream continued on the device (em2d/random.c:48-101: the two multiply-with-carry
 * generators by modular jump-ahead, the polar Box-Muller rejections by a prefix sum over the acceptance flags).
 * (z, w, have_spare, spare) are the variables m_z, m_w, iset, gset of random.c:16-17, 69-70; on return they are
 * what `count` calls of rand_norm() would have left.  d_out (device memory, `count` floats; NULL: advance the
 * state only) receives (float) (scale[m % 3] * deviate m), the narrowing of spec_set_u (em2d/particles.c:97-101).
 * Returns 1 without doing anything when the state is outside the generators' linear range (a seed word of
 * a 2^16 - 1 or more): the caller then draws on the host. */
int zdev_ref_normals( uint32_t* z, uint32_t* w, int* have_spare, double* spare, long long count,
                      const float scale[3], float* d_out );
/* host half of the above alone: the generators' state after `draws` calls of rand_uint32() (random.c:48-53) */
int zdev_ref_jump( uint32_t* z, uint32_t* w, unsigned long long draws );
/* Initial population of a lattice profile (UNIFORM / STEP / SLAB, em2d/particles.c:167-180, 335-347) generated
 * on the device on the reference random stream: the in-cell x positions kx_lo[i] <= kx < kx_hi[i] of box column
 * i carry plasma (host arrays, one entry per column of the WHOLE box), every row alike; injection order,
 * positions, thermal momenta, per-cell mean removal and fluid momentum as spec_set_x / spec_set_u produce them
 * (:96-142; the cell means are those of a square box or a cold plasma - the reference's accumulator index uses
 * nx[1] as its stride).  The stream state is advanced like zdev_ref_normals does. */
int zdev_spec2d_inject_lattice( zdev_spec2d* s, int ppcx, int ppcy, const float ufl[3], const float uth[3],
                                const int* kx_lo, const int* kx_hi,
                                uint32_t* z, uint32_t* w, int* have_spare, double* spare );

/* spec_advance minus the host bookkeeping (em2d/particles.c:1125-1259):
 * interpolate_fld (:1029-1071) + Boris push (:1146-1207) + dep_current_zamb
 * (:773-924) into g_cur's J, then boundaries / window shift and tile re-binning.
 * Kinetic energy sum (double, unscaled) and the new particle count are left on
 * the device until zdev_spec2d_fetch(). */
void zdev_spec2d_advance( zdev_spec2d* s, zdev_grid2d* g, zdev_grid2d* g_cur, const zdev_push2d_params* p );
/* energy sum of the last advance and current particle count (synchronises).
 * Also raises the fatal tile-capacity error if a tile overflowed. */
void zdev_spec2d_fetch( zdev_spec2d* s, double* energy_sum, int64_t* np );
/* spec_deposit_charge on the device (em2d/particles.c:1289-1324): charge is a host
 * (nx+1)*(ny+1) float array that is ADDED to, like the reference does */
void zdev_spec2d_deposit_charge( zdev_spec2d* s, float q, int moving_window, float* charge );
/* the multi-threaded host copy the charge deposit uses between the caller's pageable array and its pinned staging
   buffer (host only, no device call: exercised on the CPU by tests/test_abi_symbols.py) */
void zdev_spec2d_par_memcpy( void* dst, const void* src, size_t bytes );
/* the tile layout after a regrow event (a density spike filled a tile): host only, no device call - see
   zdev_spec2d.cu; returns the largest tile capacity or -1 when tile *bad_tile needs more than a push CTA can take */
int64_t zdev_spec2d_plan_regrow( int ntx, int nty, int TX, int TY, const int64_t* off, const int* np, const int* ovf,
                                 int64_t* off_new, int* bad_tile );
/* spec_deposit_pha on the device (em2d/particles.c:1569-1632): quant1/2 = the reference's
 * X1 (1), X2 (2), U1 (4), U2 (5), U3 (6), pha_buf is a host pha_nx[0]*pha_nx[1] float array that is ADDED to */
void zdev_spec2d_deposit_pha( zdev_spec2d* s, int quant1, int quant2, const int pha_nx[2], const float pha_range[2][2],
                              float q, float dx, float dy, float* pha_buf );
/* Slab decomposition along x (one process per GPU).  After zdev_spec2d_advance
 * with slab_left/right set, the particles that crossed a slab edge sit in two device lists of
 * 28-byte t_part records whose ix is already expressed in the neighbour's frame (all slabs
 * have the same width).  The caller moves them (NCCL / peer copy) and hands what it received to
 * zdev_spec2d_append_device.  export_counts synchronises the stream. */
void  zdev_spec2d_export_counts( zdev_spec2d* s, int64_t counts[2] );
void* zdev_spec2d_export_ptr( zdev_spec2d* s, int side );
void  zdev_spec2d_append_device( zdev_spec2d* s, const void* dev_part_aos, int64_t np );
/* The linked form of the same (one process per GPU on one node, csrc/dev/zdev_slab.cuh): collective over the ranks
 * of the job.  From then on zdev_spec2d_advance writes the particles that leave through a slab edge straight into
 * the neighbour's mailbox over NVLink and appends what the neighbours sent, all on the device; slab_left / slab_right
 * of the push parameters are taken from the link.  gx0, gnx: the slab's first column in the whole box and the box
 * width (the device-side injector and the phasespace axis work in box coordinates). */
void  zdev_spec2d_set_slab( zdev_spec2d* s, int left, int right, int gx0, int gnx );
/* linked slabs: enqueue the (deferred) wait for the neighbours' particles + their append now; called at the end of a
 * time step so that the transfer overlaps the current / field phase and the other species (anything that touches
 * the species does it implicitly) */
void  zdev_spec2d_flush_import( zdev_spec2d* s );
/* the slab's own charge deposit: rho = (nx+1)*(ny+1) floats, overwritten, not folded */
void  zdev_spec2d_deposit_charge_raw( zdev_spec2d* s, float q, float* rho );
/* Device timing of the push kernel alone (k_push2d, not the migration pass): when enabled
 * every launch is bracketed by CUDA events on the library stream; the accumulated time and
 * launch count are read back (and optionally reset) per species.  Used by bench.py for
 * the roofline figure. */
void zdev_set_push_timing( int on );
void zdev_spec2d_push_timing( zdev_spec2d* s, double* total_ms, int64_t* launches, int reset );
/* tile geometry chosen for this species (cells per tile in x,y; number of tiles) */
void zdev_spec2d_tile_info( zdev_spec2d* s, int* tx, int* ty, int* ntiles, int64_t* capacity );

/* ============================================================ em1d ============================ */
/* Grids: nx+3 float3 cells, guards {1 lower, 2 upper} (em1d/emf.c:41-65, em1d/current.c:33-50).
 * Boundary types are the reference enums: fields 0 none / 1 periodic / 2 open (Mur),
 * current 0 none / 1 periodic (em1d/emf.h:83-87, em1d/current.h:29-32). */

typedef struct zdev_grid1d zdev_grid1d;

zdev_grid1d* zdev_grid1d_create( int nx );
void zdev_grid1d_destroy( zdev_grid1d* g );
void zdev_grid1d_upload( zdev_grid1d* g, int which, const float* host_buf );
void zdev_grid1d_download( zdev_grid1d* g, int which, float* host_buf );
/* current_zero (em1d/current.c:97-104) */
void zdev_current1d_zero( zdev_grid1d* g );
/* current_update minus iter++ (em1d/current.c:112-155, 265-333): guard fold when periodic, then
 * xlevel binomial passes (+ compensator for xtype 2) */
void zdev_current1d_update( zdev_grid1d* g, int bc_periodic, int xtype, int xlevel );
/* emf_advance minus iter++ (em1d/emf.c:548-590): yee_b(dt/2), yee_e(dt), mur_abc when bc_type is open
 * (em1d/emf.c:379-408), yee_b(dt/2), emf_update_gc when periodic (em1d/emf.c:470-502, incl. its
 * one-cell upper refresh), external fields, window shift (em1d/emf.c:513-537) */
void zdev_emf1d_advance( zdev_grid1d* g, zdev_grid1d* g_cur, float dt, float dx, int bc_type, int shift_window );
/* Mur boundary state: mur_fld[2], mur_tmp[2] as 12 floats (host struct order) */
void zdev_emf1d_set_mur( zdev_grid1d* g, const float state[12] );
void zdev_emf1d_get_mur( zdev_grid1d* g, float state[12] );
void zdev_emf1d_set_ext_uniform( zdev_grid1d* g, int e_on, const float e0[3], int b_on, const float b0[3] );
void zdev_emf1d_set_ext_grid( zdev_grid1d* g, const float* host_ext_e, const float* host_ext_b );
/* emf_get_energy sums (em1d/emf.c:600-620) */
void zdev_emf1d_energy( zdev_grid1d* g, double sums[6] );

typedef struct zdev_spec1d zdev_spec1d;

/* per-step scalars (em1d/particles.c:925-931) */
typedef struct zdev_push1d_params {
	float tem;      /* (float)(0.5*dt/m_q) */
	float dt_dx;    /* dt/dx */
	float qnx;      /* q*dx/dt */
	float q;
	int   absorbing;      /* 1: moving window or PART_BC_OPEN (em1d/particles.c:1044-1057) */
	int   shift_window;   /* 1: all ix-- this step (em1d/particles.c:663-684) */
} zdev_push1d_params;

zdev_spec1d* zdev_spec1d_create( int nx, int ppc_hint, int track_ids );
void zdev_spec1d_destroy( zdev_spec1d* s );
/* host records are the 20-byte t_part of em1d/particles.h:29-35 */
void zdev_spec1d_upload( zdev_spec1d* s, const void* part_aos, int64_t np );
void zdev_spec1d_append( zdev_spec1d* s, const void* part_aos, int64_t np );
int64_t zdev_spec1d_download( zdev_spec1d* s, void* part_aos, int64_t max_np );
int64_t zdev_spec1d_np( zdev_spec1d* s );
/* slots allocated per buffer (grows when tiles fill up) */
int64_t zdev_spec1d_capacity( zdev_spec1d* s );
void zdev_spec1d_inject_uniform( zdev_spec1d* s, int ppc, const float ufl[3], const float uth[3], uint64_t seed );
/* the 1-D twin of zdev_spec2d_inject_lattice: UNIFORM / STEP / SLAB plasma (em1d/particles.c:245-262) generated on
 * the device on the reference random stream; the in-cell positions klo[i] <= k < khi[i] of cell i carry plasma */
int zdev_spec1d_inject_lattice( zdev_spec1d* s, int ppc, const float ufl[3], const float uth[3],
                                const int* klo, const int* khi,
                                uint32_t* z, uint32_t* w, int* have_spare, double* spare );
/* spec_advance minus host bookkeeping (em1d/particles.c:936-1062): interpolate_fld (:864-886), Boris,
 * dep_current_zamb (:707-779), boundaries / window shift, per-step re-binning */
void zdev_spec1d_advance( zdev_spec1d* s, zdev_grid1d* g, zdev_grid1d* g_cur, const zdev_push1d_params* p );
void zdev_spec1d_fetch( zdev_spec1d* s, double* energy_sum, int64_t* np );
/* spec_deposit_charge (em1d/particles.c:1085-1106): charge has nx+1 entries, added to */
void zdev_spec1d_deposit_charge( zdev_spec1d* s, float q, int moving_window, float* charge );
void zdev_spec1d_push_timing( zdev_spec1d* s, double* total_ms, int64_t* launches, int reset );

#ifdef __cplusplus
}
#endif
#endif
