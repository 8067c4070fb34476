/* zpic-b200 :: em1d simulation object (reference em1d/simulation.h) */
#ifndef ZPIC_B200_EM1D_SIMULATION_H
#define ZPIC_B200_EM1D_SIMULATION_H

#include <stdint.h>
#include "particles.h"
#include "emf.h"
#include "current.h"

typedef struct Simulation {
	float dt, tmax;              /* time step, end of the run */
	int ndump, n_species;        /* report period (0: never); length of species[] */
	t_species* species;          /* malloc'ed by the caller, freed by sim_delete */
	t_emf emf;
	t_current current;
	int moving_window;
} t_simulation;

void sim_init( t_simulation* sim );
void sim_report( t_simulation* sim );

/* replaces em1d/simulation.c:45-56 */
void sim_iter( t_simulation* sim );
/* replaces em1d/simulation.c:179-201 */
void sim_report_energy( t_simulation* sim );
/* replaces em1d/simulation.c:92-111 */
void sim_new( t_simulation* sim, int nx, float box, float dt, float tmax, int ndump, t_species* species, int n_species );
/* replaces em1d/simulation.c:65-77 */
void sim_timings( t_simulation* sim, uint64_t t0, uint64_t t1 );
/* replaces em1d/simulation.c:122-124 */
void sim_add_laser( t_simulation* sim, t_emf_laser* laser );
/* replaces em1d/simulation.c:208-217 */
void sim_delete( t_simulation* sim );
/* replaces em1d/simulation.c:159-172 */
void sim_set_moving_window( t_simulation* sim );
/* replaces em1d/simulation.c:145-152 */
void sim_set_smooth( t_simulation* sim, t_smooth* smooth );
/* replaces em1d/simulation.c:135-137 */
void sim_set_ext_fld( t_simulation* sim, t_emf_ext_fld* ext_fld );
/* replaces em1d/simulation.c:25-32 */
int report( int n, int ndump );

#endif
