/* zpic-b200 :: em2d simulation object (reference em2d/simulation.c).
 * sim_iter keeps the reference's sequencing; each stage enqueues device kernels. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "zb_state.h"
#include "zpic_b200.h"
#include "timer.h"

int report( int n, int ndump )
{
	return (ndump > 0) ? !(n % ndump) : 0;
}

void sim_iter( t_simulation* sim )
{
	if (zb_opt_coherent()) zpic_b200_touch_host(sim);   /* host buffers are the truth: re-upload */

	current_zero( &sim->current );
	for (int i = 0; i < sim->n_species; i++)
		spec_advance( &sim->species[i], &sim->emf, &sim->current );
	current_update( &sim->current );
	emf_advance( &sim->emf, &sim->current );
	/* slabs: the particles the neighbour slabs sent during spec_advance are waited for and appended only now -
	   their transfer ran behind the other species' push and the current / field phase */
	for (int i = 0; i < sim->n_species; i++) {
		zb_spec* e = zb_spec_of(&sim->species[i], 0);
		if (e && e->d && e->slab.on) zdev_spec2d_flush_import(e->d);
	}

	if (zb_opt_coherent()) zpic_b200_sync_host(sim);     /* and refresh every mirror */
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

void sim_new( t_simulation* sim, int nx[], float box[], float dt, float tmax, int ndump,
              t_species* species, int n_species )
{
	zb_par_init();       /* one process per GPU: join the job (no-op for a single process) */
	sim->dt = dt;
	sim->tmax = tmax;
	sim->ndump = ndump;
	sim->moving_window = 0;

	emf_new( &sim->emf, nx, box, dt );
	current_new( &sim->current, nx, box, dt );
	zb_grid_pair( &sim->emf, &sim->current );     /* one device grid object for E, B and J */

	sim->n_species = n_species;
	sim->species = species;

	/* Courant condition (reference simulation.c:106-110) */
	float cour = sqrtf( 1.0f / ( 1.0f/(sim->emf.dx[0]*sim->emf.dx[0]) + 1.0f/(sim->emf.dx[1]*sim->emf.dx[1]) ) );
	if (dt >= cour) {
		fprintf(stderr, "Invalid timestep, courant condition violation, dtmax = %f \n", cour);
		exit(-1);
	}
}

void sim_add_laser( t_simulation* sim, t_emf_laser* laser )
{
	emf_add_laser( &sim->emf, laser );
}

void sim_set_smooth( t_simulation* sim, t_smooth* smooth )
{
	if ( (smooth->xtype != NONE) && (smooth->xlevel <= 0) ) {
		fprintf(stderr, "Invalid smooth level along x direction\n");
		exit(-1);
	}
	if ( (smooth->ytype != NONE) && (smooth->ylevel <= 0) ) {
		fprintf(stderr, "Invalid smooth level along y direction\n");
		exit(-1);
	}
	sim->current.smooth = *smooth;
}

void sim_set_moving_window( t_simulation* sim )
{
	sim->emf.moving_window = 1;
	sim->current.moving_window = 1;
	for (int i = 0; i < sim->n_species; i++) sim->species[i].moving_window = 1;
}

void sim_set_ext_fld( t_simulation* sim, t_emf_ext_fld* ext_fld )
{
	emf_set_ext_fld( &sim->emf, ext_fld );
}

void sim_report_energy( t_simulation* sim )
{
	double emf_energy[6];
	emf_get_energy( &sim->emf, emf_energy );
	/* the reference starts the field total from component 0 and then adds all six
	   (simulation.c:191-194): reproduced so printed numbers compare */
	double tot_emf = emf_energy[0];
	for (int i = 0; i < 6; i++) tot_emf += emf_energy[i];

	double tot_part = 0;
	for (int i = 0; i < sim->n_species; i++) tot_part += sim->species[i].energy;

	if (zb_par_rank() == 0)
		printf("Energy (fields | particles | total) = %e %e %e\n", tot_emf, tot_part, tot_emf + tot_part);
}

void sim_delete( t_simulation* sim )
{
	for (int i = 0; i < sim->n_species; i++) spec_delete( &sim->species[i] );
	free( sim->species );
	current_delete( &sim->current );
	emf_delete( &sim->emf );
}
