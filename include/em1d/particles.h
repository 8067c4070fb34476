/* zpic-b200 :: em1d particle species (reference em1d/particles.h) */
#ifndef ZPIC_B200_EM1D_PARTICLES_H
#define ZPIC_B200_EM1D_PARTICLES_H

#include "zpic.h"
#include "emf.h"
#include "current.h"
#include <stdint.h>

#define MAX_SPNAME_LEN 32

/* host AoS record, 20 bytes (reference em1d/particles.h:29-35) */
typedef struct Particle {
	int ix;
	float x;
	float ux, uy, uz;
} t_part;

enum density_type { UNIFORM, EMPTY, STEP, SLAB, RAMP, CUSTOM };

/* reference em1d/particles.h:52-70 */
typedef struct Density {
	float n;                                  /* reference density (0 is read as 1) */
	enum density_type type;
	float start, end, ramp[2];                /* STEP / SLAB / RAMP limits and the two ramp densities */
	float (*custom)(float, void*);  void *custom_data;        /* CUSTOM: n(x) = n * custom(x) */
	unsigned long total_np_inj;               /* injector bookkeeping for the moving window */
	double custom_q_inj;
} t_density;

/* reference em1d/particles.h:76-80 */
enum part_boundary { PART_BC_NONE, PART_BC_PERIODIC, PART_BC_OPEN };

/* reference em1d/particles.h:86-135 */
typedef struct Species {
	char name[MAX_SPNAME_LEN+1];
	t_part *part;                 /* host mirror of the population (the device copy is authoritative while stepping) */
	int np, np_max;               /* particles in use / allocated in the mirror */
	float m_q;                    /* mass over charge */
	double energy;                /* kinetic energy of the last advance */
	float q;                      /* charge of one simulation particle */
	int ppc;
	t_density density;
	float ufl[3], uth[3];         /* fluid / thermal momenta */
	int nx;
	float dx, box, dt;
	int iter, moving_window, n_move;
	enum part_boundary bc_type;   /* re-read every step (decks poke it) */
	int n_sort;
} t_species;

/* replaces em1d/particles.c:511-584 */
void spec_new( t_species* spec, char name[], const float m_q, const int ppc,
               const float ufl[], const float uth[],
               const int nx, float box, const float dt, t_density* density );
/* replaces em1d/particles.c:621-625 */
void spec_delete( t_species* spec );
/* replaces em1d/particles.c:460-466 */
void spec_grow_buffer( t_species* spec, const int size );
/* device: interpolate + Boris + split-segment deposit + boundaries (reference em1d/particles.c:919-1074) */
void spec_advance( t_species* spec, t_emf* emf, t_current* current );
/* replaces em1d/particles.c:594-614 */
void spec_move_window( t_species *spec );
/* replaces em1d/particles.c:48-51 */
uint64_t spec_npush( void );
/* replaces em1d/particles.c:38-41 */
double spec_time( void );
/* replaces em1d/particles.c:58-61 */
double spec_perf( void );

#define CHARGE      0x1000
#define PHA         0x2000
#define PARTICLES   0x3000
#define X1          0x0001
#define U1          0x0004
#define U2          0x0005
#define U3          0x0006
#define PHASESPACE(a,b) ((a) + (b)*16 + PHA)

/* replaces em1d/particles.c:1323-1386 */
void spec_deposit_pha( const t_species *spec, const int rep_type,
                       const int pha_nx[], const float pha_range[][2], float* buf );
/* replaces em1d/particles.c:1477-1493 */
void spec_report( const t_species *spec, const int rep_type,
                  const int pha_nx[], const float pha_range[][2] );
/* replaces em1d/particles.c:1093-1112 */
void spec_deposit_charge( const t_species* spec, float* charge );

#endif
