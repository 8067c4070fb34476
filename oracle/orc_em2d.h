/* oracle/orc_em2d.h - interface of the CPU restatement (TEST INFRASTRUCTURE ONLY, see orc_em2d.c) */
#ifndef ORC_EM2D_H
#define ORC_EM2D_H

/* same 28-byte record as the reference API (em2d/particles.h:29-37) */
typedef struct { int ix, iy; float x, y, ux, uy, uz; } orc_part;

typedef struct orc_species {
	orc_part* part;          /* caller-owned buffer with room for injected particles */
	int np;
	float m_q, q;
	double energy;
	int iter, n_move, n_sort;
	void (*inject)(struct orc_species*, void*);
	void* inject_ctx;
} orc_species;

typedef struct {
	int nx, ny;
	float dx, dy, dt;
	float *E, *B, *J;        /* (nx+3)*(ny+3)*3 floats each */
	int iter, n_move, moving_window;
	int xtype, ytype, xlevel, ylevel;
	int n_species;
	orc_species* species;
} orc_sim;

void orc2d_yee_b(float* B, const float* E, int nx, int ny, float dt_dx, float dt_dy);
void orc2d_yee_e(float* E, const float* B, const float* J, int nx, int ny, float dt_dx, float dt_dy, float dt);
void orc2d_guard_copy(float* F, int nx, int ny, int moving_window);
void orc2d_shift_left(float* F, int nx, int ny);
void orc2d_emf_advance(float* E, float* B, const float* J, int nx, int ny, float dt, float dx, float dy,
                       int moving_window, int shift);
void orc2d_emf_energy(const float* E, const float* B, int nx, int ny, double out[6]);
void orc2d_current_gc(float* J, int nx, int ny, int moving_window);
void orc2d_smooth_pass(float* J, int nx, int ny, int dir, float sa, float sb, int keep_x_guards);
int orc2d_smooth_plan(int xtype, int ytype, int xlevel, int ylevel, int* dirs, float* sa, float* sb);
void orc2d_current_smooth(float* J, int nx, int ny, int moving_window, int xtype, int ytype, int xlevel, int ylevel);
double orc2d_spec_push(orc_part* part, int np, const float* E, const float* B, float* J, int nx, int ny,
                       const float prm[6]);
int orc2d_spec_boundary(orc_part* part, int np, int nx, int ny, int moving_window);
void orc2d_spec_sort(orc_part* part, int np, int nx, int ny);
void orc2d_deposit_charge(const orc_part* part, int np, float q, int nx, int ny, int moving_window, float* charge);
void orc2d_sim_iter(orc_sim* s);

#endif
