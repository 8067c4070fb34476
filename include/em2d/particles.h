/* zpic-b200 :: em2d particle species (reference em2d/particles.h) */
#ifndef ZPIC_B200_EM2D_PARTICLES_H
#define ZPIC_B200_EM2D_PARTICLES_H

#include "zpic.h"
#include "emf.h"
#include "current.h"
#include <stdint.h>

#define MAX_SPNAME_LEN 32

/* host AoS record, 28 bytes, observable from Python as a structured dtype
 * (reference particles.h:29-37). The device keeps tile-binned SoA arrays; this
 * record only exists in the host mirror. */
typedef struct Particle {
	int ix, iy;
	float x, y;
	float ux, uy, uz;
} t_part;

enum density_type { UNIFORM, EMPTY, STEP, SLAB, CUSTOM };

/* injection profile (reference particles.h:55-78) */
typedef struct Density {
	float n;                                  /* reference density (0 is read as 1) */
	enum density_type type;
	float start, end;                         /* STEP / SLAB limits along x */
	float (*custom_x)(float, void*);  void *custom_data_x;     /* CUSTOM: n(x,y) = n * custom_x(x) * custom_y(y) */
	float (*custom_y)(float, void*);  void *custom_data_y;
	unsigned long custom_x_total_part;        /* injector bookkeeping for the moving window */
	double custom_x_total_q;
} t_density;

/* species container (reference particles.h:85-132) */
typedef struct Species {
	char name[MAX_SPNAME_LEN+1];
	t_part *part;                 /* host mirror of the population (the device copy is authoritative while stepping) */
	int np, np_max;               /* particles in use / allocated in the mirror */
	float m_q;                    /* mass over charge */
	double energy;                /* kinetic energy of the last advance */
	float q;                      /* charge of one simulation particle */
	int ppc[2];
	t_density density;
	float ufl[3], uth[3];         /* fluid / thermal momenta */
	int nx[2];
	float dx[2], box[2], dt;
	int iter, moving_window, n_move, n_sort;
} t_species;

/* replaces em2d/particles.c:535-609 */
void spec_new( t_species* spec, char name[], const float m_q, const int ppc[],
               const float ufl[], const float uth[],
               const int nx[], float box[], const float dt, t_density* density );
/* replaces em2d/particles.c:647-651 */
void spec_delete( t_species* spec );
/* replaces em2d/particles.c:460-466 */
void spec_grow_buffer( t_species* spec, const int size );
/* device: fused interpolate + Boris + split-segment deposit, boundaries, window,
 * tile re-binning (reference particles.c:1104-1269) */
void spec_advance( t_species* spec, t_emf* emf, t_current* current );
/* replaces em2d/particles.c:619-640 */
void spec_move_window( t_species *spec );
/* replaces em2d/particles.c:46-49 */
uint64_t spec_npush( void );
/* replaces em2d/particles.c:36-39 */
double spec_time( void );
/* replaces em2d/particles.c:56-59 */
double spec_perf( void );

/* diagnostics selectors (reference particles.h:230-246) */
#define CHARGE      0x1000
#define PHA         0x2000
#define PARTICLES   0x3000
#define X1          0x0001
#define X2          0x0002
#define U1          0x0004
#define U2          0x0005
#define U3          0x0006
#define PHASESPACE(a,b) ((a) + (b)*16 + PHA)

/* replaces em2d/particles.c:1569-1632 */
void spec_deposit_pha( const t_species *spec, const int rep_type,
                       const int pha_nx[], const float pha_range[][2], float* buf );
/* replaces em2d/particles.c:1725-1741 */
void spec_report( const t_species *spec, const int rep_type,
                  const int pha_nx[], const float pha_range[][2] );
/* replaces em2d/particles.c:1289-1324 */
void spec_deposit_charge( const t_species* spec, float* charge );

#endif
