/* oracle/orc_em1d.h - interface of the em1d CPU restatement (TEST INFRASTRUCTURE ONLY, see orc_em1d.c) */
#ifndef ORC_EM1D_H
#define ORC_EM1D_H

typedef struct { int ix; float x, ux, uy, uz; } orc1_part;     /* em1d/particles.h:29-35 */

typedef struct orc1_species {
	orc1_part* part;
	int np;
	float m_q, q;
	double energy;
	int iter, n_move, n_sort;
	int open_bc;             /* PART_BC_OPEN */
} orc1_species;

typedef struct {
	int nx;
	float dx, dt;
	float *E, *B, *J;        /* (nx+3)*3 floats each */
	int iter, n_move, moving_window;
	int emf_bc;              /* 0 none, 1 periodic, 2 open (Mur) */
	int cur_bc;              /* 0 none, 1 periodic */
	float mur_fld[6], mur_tmp[6];
	int xtype, xlevel;
	int n_species;
	orc1_species* species;
} orc1_sim;

double orc1d_spec_push(orc1_part* part, int np, const float* E, const float* B, float* J, const float prm[4]);
int  orc1d_spec_boundary(orc1_part* part, int np, int nx, int absorbing);
void orc1d_spec_sort(orc1_part* part, int np, int nx);
void orc1d_current_update(float* J, int nx, int periodic, int xtype, int xlevel);
void orc1d_emf_advance(orc1_sim* s);
void orc1d_emf_energy(const float* E, const float* B, int nx, double out[6]);
void orc1d_deposit_charge(const orc1_part* part, int np, float q, int nx, int moving_window, float* charge);
void orc1d_sim_iter(orc1_sim* s);

#endif
