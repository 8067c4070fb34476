/* oracle/orc_em2d.c - CPU restatement of the em2d time step.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker, never the product: only tests/, __graft_entry__.smoke() and
 * the cpu_baseline leg of bench.py may load it.  It restates, on plain arrays and in our
 * own structure, the algorithm of the reference hot path
 *
 *     sim_iter            em2d/simulation.c:45-56
 *     spec_advance        em2d/particles.c:1104-1269  (interpolate_fld :1029-1071,
 *                         dep_current_zamb :773-924, boundaries :1237-1259)
 *     current_update      em2d/current.c:118-183, 297-459
 *     emf_advance         em2d/emf.c:500-716
 *     emf_get_energy      em2d/emf.c:729-750
 *     spec_deposit_charge em2d/particles.c:1289-1324
 *
 * Parity is PINNED: tests/test_oracle.py runs this restatement and the unmodified reference
 * (oracle/_ref, built from /root/reference by oracle/Makefile) on the same decks and demands
 * bit-identical particles, fields and currents; tests/golden/ holds vectors generated from
 * the reference for boxes that do not have /root/reference.
 *
 * Build: gcc -O2 -std=c99 -ffp-contract=off (strict IEEE, no contraction) - see Makefile.
 *
 * Conventions: grids are (nx+3)*(ny+3) cells of 3 floats, guards {1 lower, 2 upper};
 * G(i,j) addresses cell (i,j), i in [-1,nx+1].  Particles are the 28-byte records of the
 * reference API.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "orc_em2d.h"

#define NROW(nx) ((nx) + 3)
/* float index of component c of cell (i,j) */
#define G(i, j, c) (3 * (((i) + 1) + ((j) + 1) * nrow) + (c))
enum { X = 0, Y = 1, Z = 2 };

/* ------------------------------------------------------------------ fields */

void orc2d_yee_b(float* B, const float* E, int nx, int ny, float dt_dx, float dt_dy)
{
	const int nrow = NROW(nx);
	for (int j = -1; j <= ny; j++)
		for (int i = -1; i <= nx; i++) {
			B[G(i,j,X)] += ( - dt_dy * ( E[G(i,j+1,Z)] - E[G(i,j,Z)] ) );
			B[G(i,j,Y)] += (   dt_dx * ( E[G(i+1,j,Z)] - E[G(i,j,Z)] ) );
			B[G(i,j,Z)] += ( - dt_dx * ( E[G(i+1,j,Y)] - E[G(i,j,Y)] ) +
			                   dt_dy * ( E[G(i,j+1,X)] - E[G(i,j,X)] ) );
		}
}

void orc2d_yee_e(float* E, const float* B, const float* J, int nx, int ny, float dt_dx, float dt_dy, float dt)
{
	const int nrow = NROW(nx);
	for (int j = 0; j <= ny + 1; j++)
		for (int i = 0; i <= nx + 1; i++) {
			E[G(i,j,X)] += ( + dt_dy * ( B[G(i,j,Z)] - B[G(i,j-1,Z)] ) ) - dt * J[G(i,j,X)];
			E[G(i,j,Y)] += ( - dt_dx * ( B[G(i,j,Z)] - B[G(i-1,j,Z)] ) ) - dt * J[G(i,j,Y)];
			E[G(i,j,Z)] += ( + dt_dx * ( B[G(i,j,Y)] - B[G(i-1,j,Y)] ) -
			                   dt_dy * ( B[G(i,j,X)] - B[G(i,j-1,X)] ) ) - dt * J[G(i,j,Z)];
		}
}

static void copy_cell(float* F, int dst, int src) { F[dst] = F[src]; F[dst+1] = F[src+1]; F[dst+2] = F[src+2]; }

/* periodic guard refresh; x copies are skipped when the window moves */
void orc2d_guard_copy(float* F, int nx, int ny, int moving_window)
{
	const int nrow = NROW(nx);
	if (!moving_window)
		for (int j = -1; j <= ny + 1; j++) {
			copy_cell(F, G(-1,j,0), G(nx-1,j,0));
			copy_cell(F, G(nx,j,0), G(0,j,0));
			copy_cell(F, G(nx+1,j,0), G(1,j,0));
		}
	for (int i = -1; i <= nx + 1; i++) {
		copy_cell(F, G(i,-1,0), G(i,ny-1,0));
		copy_cell(F, G(i,ny,0), G(i,0,0));
		copy_cell(F, G(i,ny+1,0), G(i,1,0));
	}
}

/* shift one cell to the left, zero the three rightmost columns */
void orc2d_shift_left(float* F, int nx, int ny)
{
	const int nrow = NROW(nx);
	for (int j = -1; j <= ny + 1; j++) {
		memmove(&F[G(-1,j,0)], &F[G(0,j,0)], (size_t) 3 * (nx + 2) * sizeof(float));
		memset(&F[G(nx-1,j,0)], 0, (size_t) 9 * sizeof(float));
	}
}

/* one field step; `shift` = the window moves after this step (decided by the caller with the
 * reference's float test) */
void orc2d_emf_advance(float* E, float* B, const float* J, int nx, int ny, float dt, float dx, float dy,
                       int moving_window, int shift)
{
	const float dth = dt / 2.0f;
	orc2d_yee_b(B, E, nx, ny, dth / dx, dth / dy);
	orc2d_yee_e(E, B, J, nx, ny, dt / dx, dt / dy, dt);
	orc2d_yee_b(B, E, nx, ny, dth / dx, dth / dy);
	orc2d_guard_copy(E, nx, ny, moving_window);
	orc2d_guard_copy(B, nx, ny, moving_window);
	if (shift) { orc2d_shift_left(E, nx, ny); orc2d_shift_left(B, nx, ny); }
}

void orc2d_emf_energy(const float* E, const float* B, int nx, int ny, double out[6])
{
	const int nrow = NROW(nx);
	for (int k = 0; k < 6; k++) out[k] = 0;
	for (int j = 0; j < ny; j++)
		for (int i = 0; i < nx; i++)
			for (int c = 0; c < 3; c++) {
				out[c]     += E[G(i,j,c)] * E[G(i,j,c)];
				out[3 + c] += B[G(i,j,c)] * B[G(i,j,c)];
			}
}

/* ------------------------------------------------------------------ current */

static void fold_cell(float* J, int lo, int up)
{
	for (int c = 0; c < 3; c++) { J[lo + c] += J[up + c]; J[up + c] = J[lo + c]; }
}

void orc2d_current_gc(float* J, int nx, int ny, int moving_window)
{
	const int nrow = NROW(nx);
	if (!moving_window)
		for (int j = -1; j <= ny + 1; j++)
			for (int i = -1; i <= 1; i++) fold_cell(J, G(i,j,0), G(nx+i,j,0));
	for (int i = -1; i <= nx + 1; i++)
		for (int j = -1; j <= 1; j++) fold_cell(J, G(i,j,0), G(i,ny+j,0));
}

/* one [sa,sb,sa] pass along x: rows 0..ny-1 only, x guards of those rows refreshed unless the
 * window moves; inputs are the values before the pass (out of place via a row copy) */
static void pass_x(float* J, int nx, int ny, float sa, float sb, int moving_window)
{
	const int nrow = NROW(nx);
	float* old = malloc((size_t) 3 * nrow * sizeof(float));
	for (int j = 0; j < ny; j++) {
		memcpy(old, &J[G(-1,j,0)], (size_t) 3 * nrow * sizeof(float));
		const float* o = old + 3;                 /* o[3*i + c] = old value of cell (i,j) */
		for (int i = 0; i < nx; i++)
			for (int c = 0; c < 3; c++)
				J[G(i,j,c)] = sa * o[3*(i-1)+c] + sb * o[3*i+c] + sa * o[3*(i+1)+c];
		if (!moving_window) {
			copy_cell(J, G(-1,j,0), G(nx-1,j,0));
			copy_cell(J, G(nx,j,0), G(0,j,0));
			copy_cell(J, G(nx+1,j,0), G(1,j,0));
		}
	}
	free(old);
}

/* one pass along y: columns 0..nx-1, then y guards of every column */
static void pass_y(float* J, int nx, int ny, float sa, float sb)
{
	const int nrow = NROW(nx);
	float* old = malloc((size_t) 3 * nrow * (ny + 3) * sizeof(float));
	memcpy(old, J, (size_t) 3 * nrow * (ny + 3) * sizeof(float));
	for (int j = 0; j < ny; j++)
		for (int i = 0; i < nx; i++)
			for (int c = 0; c < 3; c++)
				J[G(i,j,c)] = sa * old[G(i,j-1,c)] + sb * old[G(i,j,c)] + sa * old[G(i,j+1,c)];
	free(old);
	for (int i = -1; i <= nx + 1; i++) {
		copy_cell(J, G(i,-1,0), G(i,ny-1,0));
		copy_cell(J, G(i,ny,0), G(i,0,0));
		copy_cell(J, G(i,ny+1,0), G(i,1,0));
	}
}

static void compensator(int n, float* sa, float* sb)
{
	float a = -1;
	float b = (4.0 + 2.0*n) / n;
	float total = 2*a + b;
	*sa = a / total; *sb = b / total;
}

/* single pass / pass list, for the slab-decomposition tests that exchange halos between passes */
void orc2d_smooth_pass(float* J, int nx, int ny, int dir, float sa, float sb, int keep_x_guards)
{
	if (dir == 0) pass_x(J, nx, ny, sa, sb, keep_x_guards); else pass_y(J, nx, ny, sa, sb);
}
int orc2d_smooth_plan(int xtype, int ytype, int xlevel, int ylevel, int* dirs, float* sa, float* sb)
{
	int n = 0;
	if (xtype) {
		for (int k = 0; k < xlevel; k++) { dirs[n] = 0; sa[n] = 0.25f; sb[n] = 0.5f; n++; }
		if (xtype == 2) { dirs[n] = 0; compensator(xlevel, &sa[n], &sb[n]); n++; }
	}
	if (ytype) {
		for (int k = 0; k < xlevel; k++) { dirs[n] = 1; sa[n] = 0.25f; sb[n] = 0.5f; n++; }
		if (ytype == 2) { dirs[n] = 1; compensator(ylevel, &sa[n], &sb[n]); n++; }
	}
	return n;
}

/* types: 0 none, 1 binomial, 2 compensated.  NB the y passes are counted with xlevel, as the
 * reference does (em2d/current.c:449) */
void orc2d_current_smooth(float* J, int nx, int ny, int moving_window, int xtype, int ytype, int xlevel, int ylevel)
{
	float sa, sb;
	if (xtype) {
		for (int k = 0; k < xlevel; k++) pass_x(J, nx, ny, 0.25f, 0.5f, moving_window);
		if (xtype == 2) { compensator(xlevel, &sa, &sb); pass_x(J, nx, ny, sa, sb, moving_window); }
	}
	if (ytype) {
		for (int k = 0; k < xlevel; k++) pass_y(J, nx, ny, 0.25f, 0.5f);
		if (ytype == 2) { compensator(ylevel, &sa, &sb); pass_y(J, nx, ny, sa, sb); }
	}
}

/* ------------------------------------------------------------------ particles */

typedef struct { float x0, x1, y0, y1, dx, dy, qvz; int ix, iy; } piece;

static void cut_y(piece* s, piece* n, int dj)
{
	const int jb = (dj == 1);
	const float delta = (s->y1 - jb) / s->dy;
	n->y0 = 1 - jb;
	n->y1 = s->y1 - dj;
	n->dy = s->dy * delta;
	n->iy = s->iy + dj;
	const float xc = s->x0 + s->dx * (1.0f - delta);
	n->x0 = xc; n->x1 = s->x1; n->dx = s->dx * delta; n->ix = s->ix;
	n->qvz = s->qvz * delta;
	s->y1 = jb;
	s->dy *= (1.0f - delta);
	s->dx *= (1.0f - delta);
	s->x1 = xc;
	s->qvz *= (1.0f - delta);
}

static void deposit_piece(float* J, int nrow, const piece* s, float qnx, float qny)
{
	const float S0x[2] = { 1.0f - s->x0, s->x0 }, S1x[2] = { 1.0f - s->x1, s->x1 };
	const float S0y[2] = { 1.0f - s->y0, s->y0 }, S1y[2] = { 1.0f - s->y1, s->y1 };
	const float wl1 = qnx * s->dx, wl2 = qny * s->dy;
	const float wp1[2] = { 0.5f*(S0y[0] + S1y[0]), 0.5f*(S0y[1] + S1y[1]) };
	const float wp2[2] = { 0.5f*(S0x[0] + S1x[0]), 0.5f*(S0x[1] + S1x[1]) };
	const int i = s->ix, j = s->iy;
	J[G(i,j,X)]   += wl1 * wp1[0];
	J[G(i,j+1,X)] += wl1 * wp1[1];
	J[G(i,j,Y)]   += wl2 * wp2[0];
	J[G(i+1,j,Y)] += wl2 * wp2[1];
	for (int b = 0; b < 2; b++)
		for (int a = 0; a < 2; a++)
			J[G(i+a,j+b,Z)] += s->qvz * ( S0x[a]*S0y[b] + S1x[a]*S1y[b] + ( S0x[a]*S1y[b] - S1x[a]*S0y[b] ) / 2.0f );
}

/* charge conserving deposit of one move (linear shapes, trajectory cut at cell faces) */
static void deposit_move(float* J, int nrow, int ix, int iy, int di, int dj,
                         float x0, float y0, float dx, float dy, float qnx, float qny, float qvz)
{
	piece vp[3];
	int n = 1;
	vp[0] = (piece) { .x0 = x0, .y0 = y0, .dx = dx, .dy = dy, .x1 = x0 + dx, .y1 = y0 + dy,
	                  .qvz = qvz / 2.0, .ix = ix, .iy = iy };
	if (di != 0) {
		const int ib = (di == 1);
		const float delta = (x0 + dx - ib) / dx;
		const float yc = y0 + dy * (1.0f - delta);
		vp[1] = (piece) { .x0 = 1 - ib, .x1 = (x0 + dx) - di, .dx = dx * delta, .ix = ix + di,
		                  .y0 = yc, .y1 = vp[0].y1, .dy = dy * delta, .iy = iy, .qvz = vp[0].qvz * delta };
		vp[0].x1 = ib;
		vp[0].dx *= (1.0f - delta);
		vp[0].dy *= (1.0f - delta);
		vp[0].y1 = yc;
		vp[0].qvz *= (1.0f - delta);
		n = 2;
	}
	if (dj != 0) {
		const int first_crosses = (vp[0].y1 < 0.0f || vp[0].y1 >= 1.0f);
		piece* s = first_crosses ? &vp[0] : &vp[1];
		cut_y(s, &vp[n], dj);
		if (first_crosses && n == 2) { vp[1].y0 -= dj; vp[1].y1 -= dj; vp[1].iy += dj; }
		n++;
	}
	for (int k = 0; k < n; k++) deposit_piece(J, nrow, &vp[k], qnx, qny);
}

static void gather(const float* E, const float* B, int nrow, int i, int j, float w1, float w2, float Ep[3], float Bp[3])
{
	const int ih = i + ((w1 < 0.5f) ? -1 : 0), jh = j + ((w2 < 0.5f) ? -1 : 0);
	const float w1h = w1 + ((w1 < 0.5f) ? 0.5f : -0.5f), w2h = w2 + ((w2 < 0.5f) ? 0.5f : -0.5f);
#define LERP2(F, c, a, b, wa, wb) \
	( ( F[G(a,b,c)] * (1.0f - (wa)) + F[G((a)+1,b,c)] * (wa) ) * (1.0f - (wb)) + \
	  ( F[G(a,(b)+1,c)] * (1.0f - (wa)) + F[G((a)+1,(b)+1,c)] * (wa) ) * (wb) )
	Ep[X] = LERP2(E, X, ih, j,  w1h, w2);
	Ep[Y] = LERP2(E, Y, i,  jh, w1,  w2h);
	Ep[Z] = LERP2(E, Z, i,  j,  w1,  w2);
	Bp[X] = LERP2(B, X, i,  jh, w1,  w2h);
	Bp[Y] = LERP2(B, Y, ih, j,  w1h, w2);
	Bp[Z] = LERP2(B, Z, ih, jh, w1h, w2h);
#undef LERP2
}

/* Push + deposit of one species (the particle loop of spec_advance).  Returns the (unscaled)
 * kinetic energy sum.  prm: tem, dt_dx, dt_dy, qnx, qny, q. */
double orc2d_spec_push(orc_part* part, int np, const float* E, const float* B, float* J, int nx, int ny,
                       const float prm[6])
{
	const int nrow = NROW(nx);
	const float tem = prm[0], dt_dx = prm[1], dt_dy = prm[2], qnx = prm[3], qny = prm[4], q = prm[5];
	double energy = 0;

	for (int k = 0; k < np; k++) {
		orc_part* p = &part[k];
		float Ep[3], Bp[3];
		gather(E, B, nrow, p->ix, p->iy, p->x, p->y, Ep, Bp);

		Ep[X] *= tem; Ep[Y] *= tem; Ep[Z] *= tem;
		float utx = p->ux + Ep[X], uty = p->uy + Ep[Y], utz = p->uz + Ep[Z];
		const float utsq = utx*utx + uty*uty + utz*utz;
		const float gamma = sqrtf( 1.0f + utsq );
		energy += utsq / (gamma + 1);
		const float tg = tem / gamma;
		Bp[X] *= tg; Bp[Y] *= tg; Bp[Z] *= tg;
		const float otsq = 2.0f / ( 1.0f + Bp[X]*Bp[X] + Bp[Y]*Bp[Y] + Bp[Z]*Bp[Z] );
		float ux = utx + uty*Bp[Z] - utz*Bp[Y];
		float uy = uty + utz*Bp[X] - utx*Bp[Z];
		float uz = utz + utx*Bp[Y] - uty*Bp[X];
		Bp[X] *= otsq; Bp[Y] *= otsq; Bp[Z] *= otsq;
		utx += uy*Bp[Z] - uz*Bp[Y];
		uty += uz*Bp[X] - ux*Bp[Z];
		utz += ux*Bp[Y] - uy*Bp[X];
		ux = utx + Ep[X]; uy = uty + Ep[Y]; uz = utz + Ep[Z];
		p->ux = ux; p->uy = uy; p->uz = uz;

		const float rg = 1.0f / sqrtf( 1.0f + ux*ux + uy*uy + uz*uz );
		const float dx = dt_dx * rg * ux, dy = dt_dy * rg * uy;
		float x1 = p->x + dx, y1 = p->y + dy;
		const int di = (x1 >= 1.0f) - (x1 < 0.0f), dj = (y1 >= 1.0f) - (y1 < 0.0f);
		x1 -= di; y1 -= dj;
		const float qvz = q * uz * rg;

		deposit_move(J, nrow, p->ix, p->iy, di, dj, p->x, p->y, dx, dy, qnx, qny, qvz);

		p->x = x1; p->y = y1;
		p->ix += di; p->iy += dj;
	}
	return energy;
}

/* boundaries after the push (and after the window shift / injection):
 * moving window: absorbing in x (swap with last; the swapped-in particle is re-tested), periodic
 * in y; otherwise periodic in both directions.  Returns the new particle count. */
int orc2d_spec_boundary(orc_part* part, int np, int nx, int ny, int moving_window)
{
	if (moving_window) {
		int k = 0;
		while (k < np) {
			if (part[k].ix < 0 || part[k].ix >= nx) { part[k] = part[--np]; continue; }
			part[k].iy += ((part[k].iy < 0) ? ny : 0) - ((part[k].iy >= ny) ? ny : 0);
			k++;
		}
	} else {
		for (int k = 0; k < np; k++) {
			part[k].ix += ((part[k].ix < 0) ? nx : 0) - ((part[k].ix >= nx) ? nx : 0);
			part[k].iy += ((part[k].iy < 0) ? ny : 0) - ((part[k].iy >= ny) ? ny : 0);
		}
	}
	return np;
}

/* stable counting sort by cell (the permutation spec_sort applies, em2d/particles.c:942-1007) */
void orc2d_spec_sort(orc_part* part, int np, int nx, int ny)
{
	const int ncell = nx * ny;
	int* start = calloc((size_t) ncell + 1, sizeof(int));
	orc_part* tmp = malloc((size_t) (np > 0 ? np : 1) * sizeof(orc_part));
	for (int k = 0; k < np; k++) start[part[k].ix + part[k].iy * nx + 1]++;
	for (int c = 0; c < ncell; c++) start[c + 1] += start[c];
	for (int k = 0; k < np; k++) tmp[start[part[k].ix + part[k].iy * nx]++] = part[k];
	memcpy(part, tmp, (size_t) np * sizeof(orc_part));
	free(tmp); free(start);
}

/* node centred linear charge deposit; charge has (nx+1)*(ny+1) entries and is added to */
void orc2d_deposit_charge(const orc_part* part, int np, float q, int nx, int ny, int moving_window, float* charge)
{
	const int nr = nx + 1;
	for (int k = 0; k < np; k++) {
		const int idx = part[k].ix + nr * part[k].iy;
		const float w1 = part[k].x, w2 = part[k].y;
		charge[idx]          += ( 1.0f - w1 ) * ( 1.0f - w2 ) * q;
		charge[idx + 1]      += (        w1 ) * ( 1.0f - w2 ) * q;
		charge[idx + nr]     += ( 1.0f - w1 ) * (        w2 ) * q;
		charge[idx + 1 + nr] += (        w1 ) * (        w2 ) * q;
	}
	if (!moving_window)
		for (int j = 0; j < ny + 1; j++) charge[j * nr] += charge[nx + j * nr];
	for (int i = 0; i < nx + 1; i++) charge[i] += charge[i + ny * nr];
}

/* ------------------------------------------------------------------ one full iteration */

/* sim_iter for a set of species sharing one grid.  When the window moves, sp->inject (if set)
 * is called between the index shift and the boundary pass, where the reference injects the new
 * right-hand column (it needs the host random stream, which the caller owns). */
void orc2d_sim_iter(orc_sim* s)
{
	const size_t n = (size_t) 3 * NROW(s->nx) * (s->ny + 3);
	memset(s->J, 0, n * sizeof(float));
	for (int k = 0; k < s->n_species; k++) {
		orc_species* sp = &s->species[k];
		const float prm[6] = { (float) (0.5 * s->dt / sp->m_q), s->dt / s->dx, s->dt / s->dy,
		                       sp->q * s->dx / s->dt, sp->q * s->dy / s->dt, sp->q };
		const double e = orc2d_spec_push(sp->part, sp->np, s->E, s->B, s->J, s->nx, s->ny, prm);
		sp->energy = sp->q * sp->m_q * e * s->dx * s->dy;
		sp->iter++;
		if (s->moving_window && ( (sp->iter * s->dt) > (s->dx * (sp->n_move + 1)) )) {
			for (int i = 0; i < sp->np; i++) sp->part[i].ix--;
			sp->n_move++;
			if (sp->inject) sp->inject(sp, sp->inject_ctx);
		}
		sp->np = orc2d_spec_boundary(sp->part, sp->np, s->nx, s->ny, s->moving_window);
		if (sp->n_sort > 0 && !(sp->iter % sp->n_sort)) orc2d_spec_sort(sp->part, sp->np, s->nx, s->ny);
	}
	orc2d_current_gc(s->J, s->nx, s->ny, s->moving_window);
	orc2d_current_smooth(s->J, s->nx, s->ny, s->moving_window, s->xtype, s->ytype, s->xlevel, s->ylevel);
	const int shift = s->moving_window && ( ((s->iter + 1) * s->dt) > s->dx * (s->n_move + 1) );
	orc2d_emf_advance(s->E, s->B, s->J, s->nx, s->ny, s->dt, s->dx, s->dy, s->moving_window, shift);
	s->iter++;
	if (shift) s->n_move++;
}
