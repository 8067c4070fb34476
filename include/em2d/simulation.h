/* zpic-b200 :: em2d simulation object (reference em2d/simulation.h) */
#ifndef ZPIC_B200_EM2D_SIMULATION_H
#define ZPIC_B200_EM2D_SIMULATION_H

#include <stdint.h>
#include "particles.h"
#include "emf.h"
#include "current.h"

/* reference simulation.h:13-29 */
typedef struct Simulation {
	float dt, tmax;              /* time step, end of the run */
	int ndump, n_species;        /* report period (0: never); length of species[] */
	t_species* species;          /* malloc'ed by the caller, freed by sim_delete */
	t_emf emf;
	t_current current;
	int moving_window;
} t_simulation;

/* supplied by the input deck */
void sim_init( t_simulation* sim );
void sim_report( t_simulation* sim );

/* one time step on the device: current_zero, spec_advance x n_species,
 * current_update, emf_advance (reference simulation.c:45-56) */
void sim_iter( t_simulation* sim );
/* replaces em2d/simulation.c:183-205 */
void sim_report_energy( t_simulation* sim );
/* replaces em2d/simulation.c:93-112 */
void sim_new( t_simulation* sim, int nx[], float box[], float dt, float tmax, int ndump,
              t_species* species, int n_species );
/* replaces em2d/simulation.c:25-32 */
int  report( int n, int ndump );
/* replaces em2d/simulation.c:66-78 */
void sim_timings( t_simulation* sim, uint64_t t0, uint64_t t1 );
/* replaces em2d/simulation.c:123-126 */
void sim_add_laser( t_simulation* sim, t_emf_laser* laser );
/* replaces em2d/simulation.c:212-222 */
void sim_delete( t_simulation* sim );
/* replaces em2d/simulation.c:154-163 */
void sim_set_moving_window( t_simulation* sim );
/* replaces em2d/simulation.c:134-147 */
void sim_set_smooth( t_simulation* sim, t_smooth* smooth );
/* replaces em2d/simulation.c:174-176 */
void sim_set_ext_fld( t_simulation* sim, t_emf_ext_fld* ext_fld );

#endif
