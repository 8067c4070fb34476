/* oracle/orc_em1d.c - CPU restatement of the em1d time step.  TEST INFRASTRUCTURE ONLY (see
 * oracle/README.md): only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load it.
 *
 * Restates, on plain arrays, the reference hot path
 *     sim_iter          em1d/simulation.c:45-56
 *     spec_advance      em1d/particles.c:919-1074 (interpolate_fld :864-886, dep_current_zamb :707-779)
 *     current_update    em1d/current.c:112-155, 265-333
 *     emf_advance       em1d/emf.c:379-590 (yee_b/yee_e, mur_abc, emf_update_gc, emf_move_window)
 * Parity is PINNED against the unmodified reference build (tests/test_oracle_em1d.py: bit-identical
 * on the two-stream deck incl. sorts, an open-boundary laser deck and a moving-window deck).
 * Build: gcc -O2 -std=c99 -ffp-contract=off.  Grids: nx+3 cells of 3 floats, F(i,c) = cell i in [-1,nx+1].
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "orc_em1d.h"

#define F(i, c) (3 * ((i) + 1) + (c))
enum { X = 0, Y = 1, Z = 2 };

/* ------------------------------------------------------------------ particles */

typedef struct { float x0, x1, dx, qvy, qvz; int ix; } piece1;

static void deposit_piece(float* J, const piece1* s, float qnx)
{
	const float S0x[2] = { 1.0f - s->x0, s->x0 }, S1x[2] = { 1.0f - s->x1, s->x1 };
	J[F(s->ix, X)]     += qnx * s->dx;
	J[F(s->ix, Y)]     += s->qvy * ( S0x[0] + S1x[0] + ( S0x[0] - S1x[0] ) / 2.0f );
	J[F(s->ix + 1, Y)] += s->qvy * ( S0x[1] + S1x[1] + ( S0x[1] - S1x[1] ) / 2.0f );
	J[F(s->ix, Z)]     += s->qvz * ( S0x[0] + S1x[0] + ( S0x[0] - S1x[0] ) / 2.0f );
	J[F(s->ix + 1, Z)] += s->qvz * ( S0x[1] + S1x[1] + ( S0x[1] - S1x[1] ) / 2.0f );
}

/* prm: tem, dt_dx, qnx, q.  Returns the unscaled kinetic energy sum. */
double orc1d_spec_push(orc1_part* part, int np, const float* E, const float* B, float* J, const float prm[4])
{
	const float tem = prm[0], dt_dx = prm[1], qnx = prm[2], q = prm[3];
	double energy = 0;
	for (int k = 0; k < np; k++) {
		orc1_part* p = &part[k];
		const int i = p->ix;
		const float w1 = p->x;
		const int ih = i + ((w1 < 0.5f) ? -1 : 0);
		const float w1h = w1 + ((w1 < 0.5f) ? 0.5f : -0.5f);
		float Ep[3], Bp[3];
		Ep[X] = E[F(ih,X)] * (1.0f - w1h) + E[F(ih+1,X)] * w1h;
		Ep[Y] = E[F(i,Y)]  * (1.0f - w1)  + E[F(i+1,Y)]  * w1;
		Ep[Z] = E[F(i,Z)]  * (1.0f - w1)  + E[F(i+1,Z)]  * w1;
		Bp[X] = B[F(i,X)]  * (1.0f - w1)  + B[F(i+1,X)]  * w1;
		Bp[Y] = B[F(ih,Y)] * (1.0f - w1h) + B[F(ih+1,Y)] * w1h;
		Bp[Z] = B[F(ih,Z)] * (1.0f - w1h) + B[F(ih+1,Z)] * w1h;

		Ep[X] *= tem; Ep[Y] *= tem; Ep[Z] *= tem;
		float utx = p->ux + Ep[X], uty = p->uy + Ep[Y], utz = p->uz + Ep[Z];
		const float u2 = utx*utx + uty*uty + utz*utz;
		const float gamma = sqrtf( 1 + u2 );
		energy += u2 / ( 1 + gamma );
		const float gtem = tem / gamma;
		Bp[X] *= gtem; Bp[Y] *= gtem; Bp[Z] *= gtem;
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
		const float dx = dt_dx * rg * ux;
		float x1 = p->x + dx;
		const int di = (x1 >= 1.0f) - (x1 < 0.0f);
		x1 -= di;
		const float qvy = q * uy * rg, qvz = q * uz * rg;

		/* split at the cell face (dep_current_zamb) */
		piece1 a = { .x0 = p->x, .dx = dx, .x1 = p->x + dx, .qvy = qvy / 2.0, .qvz = qvz / 2.0, .ix = p->ix };
		if (di != 0) {
			const int ib = (di == 1);
			const float delta = (p->x + dx - ib) / dx;
			piece1 b = { .x0 = 1 - ib, .x1 = (p->x + dx) - di, .dx = dx * delta, .ix = p->ix + di,
			             .qvy = a.qvy * delta, .qvz = a.qvz * delta };
			a.x1 = ib;
			a.dx *= (1.0f - delta);
			a.qvy *= (1.0f - delta);
			a.qvz *= (1.0f - delta);
			deposit_piece(J, &a, qnx);
			deposit_piece(J, &b, qnx);
		} else {
			deposit_piece(J, &a, qnx);
		}
		p->x = x1;
		p->ix += di;
	}
	return energy;
}

int orc1d_spec_boundary(orc1_part* part, int np, int nx, int absorbing)
{
	if (absorbing) {
		int k = 0;
		while (k < np) {
			if (part[k].ix < 0 || part[k].ix >= nx) { part[k] = part[--np]; continue; }
			k++;
		}
	} else {
		for (int k = 0; k < np; k++) part[k].ix += ((part[k].ix < 0) ? nx : 0) - ((part[k].ix >= nx) ? nx : 0);
	}
	return np;
}

void orc1d_spec_sort(orc1_part* part, int np, int nx)
{
	int* start = calloc((size_t) nx + 1, sizeof(int));
	orc1_part* tmp = malloc((size_t) (np > 0 ? np : 1) * sizeof(orc1_part));
	for (int k = 0; k < np; k++) start[part[k].ix + 1]++;
	for (int c = 0; c < nx; c++) start[c + 1] += start[c];
	for (int k = 0; k < np; k++) tmp[start[part[k].ix]++] = part[k];
	memcpy(part, tmp, (size_t) np * sizeof(orc1_part));
	free(tmp); free(start);
}

void orc1d_deposit_charge(const orc1_part* part, int np, float q, int nx, int moving_window, float* charge)
{
	for (int k = 0; k < np; k++) {
		charge[part[k].ix]     += ( 1.0f - part[k].x ) * q;
		charge[part[k].ix + 1] += (        part[k].x ) * q;
	}
	if (!moving_window) charge[0] += charge[nx];
}

/* ------------------------------------------------------------------ current */

void orc1d_current_update(float* J, int nx, int periodic, int xtype, int xlevel)
{
	if (periodic)
		for (int i = -1; i < 2; i++)
			for (int c = 0; c < 3; c++) { J[F(i,c)] += J[F(nx+i,c)]; J[F(nx+i,c)] = J[F(i,c)]; }
	if (!xtype) return;
	float* old = malloc((size_t) 3 * (nx + 3) * sizeof(float));
	const int npass = xlevel + (xtype == 2);
	for (int pass = 0; pass < npass; pass++) {
		float sa = 0.25f, sb = 0.5f;
		if (pass == xlevel) { float a = -1, b = (4.0 + 2.0*xlevel) / xlevel, total = 2*a + b; sa = a / total; sb = b / total; }
		memcpy(old, J, (size_t) 3 * (nx + 3) * sizeof(float));
		for (int i = 0; i < nx; i++)
			for (int c = 0; c < 3; c++) J[F(i,c)] = sa * old[F(i-1,c)] + sb * old[F(i,c)] + sa * old[F(i+1,c)];
		if (periodic)
			for (int c = 0; c < 3; c++) { J[F(-1,c)] = J[F(nx-1,c)]; J[F(nx,c)] = J[F(0,c)]; J[F(nx+1,c)] = J[F(1,c)]; }
	}
	free(old);
}

/* ------------------------------------------------------------------ fields */

static void yee_b(float* B, const float* E, int nx, float dt_dx)
{
	for (int i = -1; i <= nx; i++) {
		B[F(i,Y)] += (   dt_dx * ( E[F(i+1,Z)] - E[F(i,Z)] ) );
		B[F(i,Z)] += ( - dt_dx * ( E[F(i+1,Y)] - E[F(i,Y)] ) );
	}
}

void orc1d_emf_advance(orc1_sim* s)
{
	const int nx = s->nx;
	float *E = s->E, *B = s->B; const float* J = s->J;
	const float dt = s->dt, dth = dt / 2.0f, dt_dx = dt / s->dx;
	yee_b(B, E, nx, dth / s->dx);
	for (int i = 0; i <= nx + 1; i++) {
		E[F(i,X)] += (                                      - dt * J[F(i,X)] );
		E[F(i,Y)] += ( - dt_dx * ( B[F(i,Z)] - B[F(i-1,Z)] ) - dt * J[F(i,Y)] );
		E[F(i,Z)] += ( + dt_dx * ( B[F(i,Y)] - B[F(i-1,Y)] ) - dt * J[F(i,Z)] );
	}
	if (s->emf_bc == 2) {
		/* first order Mur boundary; mur_fld / mur_tmp hold [lower xyz, upper xyz] */
		const float S = (s->dt - s->dx) / (s->dt + s->dx);
		for (int side = 0; side < 2; side++) {
			const int in = side ? nx - 1 : 0, out = side ? nx : -1;
			float* fld = s->mur_fld + 3 * side; float* tmp = s->mur_tmp + 3 * side;
			fld[Y] = tmp[Y] + S * ( E[F(in,Y)] - fld[Y] );
			fld[Z] = tmp[Z] + S * ( E[F(in,Z)] - fld[Z] );
			E[F(out,Y)] = fld[Y]; E[F(out,Z)] = fld[Z];
			tmp[Y] = E[F(in,Y)]; tmp[Z] = E[F(in,Z)];
		}
	}
	yee_b(B, E, nx, dth / s->dx);
	if (s->emf_bc == 1)
		for (int c = 0; c < 3; c++) {
			E[F(-1,c)] = E[F(nx-1,c)]; B[F(-1,c)] = B[F(nx-1,c)];
			E[F(nx,c)] = E[F(0,c)];    B[F(nx,c)] = B[F(0,c)];      /* only one upper cell, as the reference */
		}
	s->iter++;
	if (s->moving_window && ( (s->iter * s->dt) > s->dx * (s->n_move + 1) )) {
		memmove(&E[F(-1,0)], &E[F(0,0)], (size_t) 3 * (nx + 2) * sizeof(float));
		memmove(&B[F(-1,0)], &B[F(0,0)], (size_t) 3 * (nx + 2) * sizeof(float));
		memset(&E[F(nx-1,0)], 0, 9 * sizeof(float));
		memset(&B[F(nx-1,0)], 0, 9 * sizeof(float));
		s->n_move++;
	}
}

void orc1d_emf_energy(const float* E, const float* B, int nx, double out[6])
{
	for (int k = 0; k < 6; k++) out[k] = 0;
	for (int i = 0; i < nx; i++)
		for (int c = 0; c < 3; c++) { out[c] += E[F(i,c)] * E[F(i,c)]; out[3+c] += B[F(i,c)] * B[F(i,c)]; }
}

/* ------------------------------------------------------------------ one iteration (no window injection: caller's) */

void orc1d_sim_iter(orc1_sim* s)
{
	memset(s->J, 0, (size_t) 3 * (s->nx + 3) * sizeof(float));
	for (int k = 0; k < s->n_species; k++) {
		orc1_species* sp = &s->species[k];
		const float prm[4] = { (float) (0.5 * s->dt / sp->m_q), s->dt / s->dx, sp->q * s->dx / s->dt, sp->q };
		const double e = orc1d_spec_push(sp->part, sp->np, s->E, s->B, s->J, prm);
		sp->energy = sp->q * sp->m_q * e * s->dx;
		sp->iter++;
		const int absorbing = s->moving_window || sp->open_bc;
		if (s->moving_window && ( (sp->iter * s->dt) > (s->dx * (sp->n_move + 1)) )) {
			for (int i = 0; i < sp->np; i++) sp->part[i].ix--;
			sp->n_move++;
		}
		sp->np = orc1d_spec_boundary(sp->part, sp->np, s->nx, absorbing);
		if (sp->n_sort > 0 && !(sp->iter % sp->n_sort)) orc1d_spec_sort(sp->part, sp->np, s->nx);
	}
	orc1d_current_update(s->J, s->nx, s->cur_bc == 1, s->xtype, s->xlevel);
	orc1d_emf_advance(s);
}
