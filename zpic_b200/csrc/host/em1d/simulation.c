/* zpic-b200 :: em1d simulation object (reference em1d/simulation.c) */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "zb_state.h"
#include "zpic_b200.h"
#include "timer.h"

int report( int n, int ndump ) { return (ndump > 0) ? !(n % ndump) : 0; }

void sim_iter( t_simulation* sim )
{
	if (zb_opt_coherent()) zpic_b200_touch_host(sim);
	current_zero( &sim->current );
	for (int i = 0; i < sim->n_species; i++)
		spec_advance( &sim->species[i], &sim->emf, &sim->current );
	current_update( &sim->current );
	emf_advance( &sim->emf, &sim->current );
	if (zb_opt_coherent()) zpic_b200_sync_host(sim);
}

void sim_timings( t_simulation* sim, uint64_t t0, uint64_t t1 )
{
	(void) sim;
	fprintf(stderr, "Time for spec. advance = %f s\n", spec_time());
	fprintf(stderr, "Time for emf   advance = %f s\n", emf_time());
	fprintf(stderr, "Total simulation time  = %f s\n", timer_interval_seconds(t0, t1));
	fprintf(stderr, "\n");
	double perf = spec_perf();
	if (perf > 0) {
		fprintf(stderr, "Particle advance [nsec/part] = %f \n", 1.e9 * perf);
		fprintf(stderr, "Particle advance [Mpart/sec] = %f \n", 1.e-6 / perf);
	}
}

void sim_new( t_simulation* sim, int nx, float box, float dt, float tmax, int ndump, t_species* species, int n_species )
{
	sim->dt = dt;
	sim->tmax = tmax;
	sim->ndump = ndump;
	sim->moving_window = 0;
	emf_new( &sim->emf, nx, box, dt );
	current_new( &sim->current, nx, box, dt );
	zb_grid_pair( &sim->emf, &sim->current );
	sim->n_species = n_species;
	sim->species = species;

	/* Courant condition (reference em1d/simulation.c:104-109) */
	float cour = sim->emf.dx;
	if (dt >= cour) {
		fprintf(stderr, "Invalid timestep, courant condition violation, dtmax = %f \n", cour);
		exit(-1);
	}
}

void sim_add_laser( t_simulation* sim, t_emf_laser* laser ) { emf_add_laser( &sim->emf, laser ); }

void sim_set_smooth( t_simulation* sim, t_smooth* smooth )
{
	if ( (smooth->xtype != NONE) && (smooth->xlevel <= 0) ) {
		fprintf(stderr, "Invalid smooth level along x direction\n");
		exit(-1);
	}
	sim->current.smooth = *smooth;
}

/* the window also switches the field and current boundaries off (reference em1d/simulation.c:159-173) */
void sim_set_moving_window( t_simulation* sim )
{
	sim->emf.moving_window = 1;
	sim->emf.bc_type = EMF_BC_NONE;
	sim->current.bc_type = CURRENT_BC_NONE;
	for (int i = 0; i < sim->n_species; i++) sim->species[i].moving_window = 1;
}

void sim_set_ext_fld( t_simulation* sim, t_emf_ext_fld* ext_fld ) { emf_set_ext_fld( &sim->emf, ext_fld ); }

void sim_report_energy( t_simulation* sim )
{
	double emf_energy[6];
	emf_get_energy( &sim->emf, emf_energy );
	double tot_emf = emf_energy[0];            /* counted twice, as the reference does (simulation.c:187-190) */
	for (int i = 0; i < 6; i++) tot_emf += emf_energy[i];
	double tot_part = 0;
	for (int i = 0; i < sim->n_species; i++) tot_part += sim->species[i].energy;
	printf("Energy (fields | particles | total) = %e %e %e\n", tot_emf, tot_part, tot_emf + tot_part);
}

void sim_delete( t_simulation* sim )
{
	for (int i = 0; i < sim->n_species; i++) spec_delete( &sim->species[i] );
	free( sim->species );
	current_delete( &sim->current );
	emf_delete( &sim->emf );
}
