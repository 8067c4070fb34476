// zpic-b200 :: em1d field and current grids on the device.
//
// Same philosophy as zdev_grid2d.cu: the device buffers have exactly the reference layout
// (nx+3 AoS float3, guards {1 lower, 2 upper}; em1d/emf.c:41-65, em1d/current.c:33-50), one thread per
// cell, --fmad=false, expression order of the reference => bit-identical results.  The grids are tiny
// next to the particles (50 MB at 2^22 cells), so these kernels are launch-latency bound.
#include "zdev_common.cuh"

struct zdev_grid1d {
	int nx, n;               // n = nx + 3
	f3 *E, *B, *J, *tmp;     // point at cell -1
	f3 *Epart, *Bpart, *Eext, *Bext;
	int e_ext, b_ext;        // 0 none, 1 uniform, 2 grid
	f3 e0, b0;
	f3* mur;                 // 4 x f3: mur_fld[0], mur_fld[1], mur_tmp[0], mur_tmp[1]
	double* d_sums;
};

static f3* g1_alloc(zdev_grid1d* g) {
	f3* p; ZDEV_CHECK(cudaMalloc(&p, (size_t) g->n * sizeof(f3)));
	ZDEV_CHECK(cudaMemsetAsync(p, 0, (size_t) g->n * sizeof(f3), zdev_strm));
	return p;
}
static void need_EB(zdev_grid1d* g) { if (!g->E) { g->E = g1_alloc(g); g->B = g1_alloc(g); g->Epart = g->E; g->Bpart = g->B; } }
static void need_J(zdev_grid1d* g) { if (!g->J) g->J = g1_alloc(g); }
static void need_tmp(zdev_grid1d* g) { if (!g->tmp) g->tmp = g1_alloc(g); }

extern "C" zdev_grid1d* zdev_grid1d_create(int nx) {
	zdev_require_init();
	zdev_grid1d* g = (zdev_grid1d*) calloc(1, sizeof(zdev_grid1d));
	g->nx = nx; g->n = nx + 3;
	ZDEV_CHECK(cudaMalloc(&g->d_sums, 6 * sizeof(double)));
	ZDEV_CHECK(cudaMalloc(&g->mur, 4 * sizeof(f3)));
	ZDEV_CHECK(cudaMemsetAsync(g->mur, 0, 4 * sizeof(f3), zdev_strm));
	return g;
}

extern "C" void zdev_grid1d_destroy(zdev_grid1d* g) {
	if (!g) return;
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	if (g->e_ext) cudaFree(g->Epart);
	if (g->b_ext) cudaFree(g->Bpart);
	cudaFree(g->Eext); cudaFree(g->Bext);
	cudaFree(g->E); cudaFree(g->B); cudaFree(g->J); cudaFree(g->tmp); cudaFree(g->mur); cudaFree(g->d_sums);
	free(g);
}

static f3* g1_sel(zdev_grid1d* g, int which) {
	if (which == ZDEV_J) need_J(g); else need_EB(g);
	switch (which) {
	case ZDEV_E: return g->E;
	case ZDEV_B: return g->B;
	case ZDEV_J: return g->J;
	case ZDEV_EPART: return g->Epart;
	case ZDEV_BPART: return g->Bpart;
	}
	fprintf(stderr, "(*error*) zdev_grid1d: invalid grid selector %d\n", which); exit(-1);
}

extern "C" void zdev_grid1d_upload(zdev_grid1d* g, int which, const float* h) {
	ZDEV_CHECK(cudaMemcpyAsync(g1_sel(g, which), h, (size_t) g->n * sizeof(f3), cudaMemcpyHostToDevice, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
}
extern "C" void zdev_grid1d_download(zdev_grid1d* g, int which, float* h) {
	ZDEV_CHECK(cudaMemcpyAsync(h, g1_sel(g, which), (size_t) g->n * sizeof(f3), cudaMemcpyDeviceToHost, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
}

extern "C" void zdev_current1d_zero(zdev_grid1d* g) {
	need_J(g);
	ZDEV_CHECK(cudaMemsetAsync(g->J, 0, (size_t) g->n * sizeof(f3), zdev_strm));
}

// ------------------------------------------------------------------ Yee solver (em1d/emf.c:422-460)

// i in [-1, nx]; Bx is static in 1D
__global__ void k1_yee_b(f3* __restrict__ B, const f3* __restrict__ E, int nx, float dt_dx) {
	int i = blockIdx.x * blockDim.x + threadIdx.x - 1;
	if (i > nx) return;
	f3 e = E[i + 1], e1 = E[i + 2], b = B[i + 1];
	b.y += (   dt_dx * ( e1.z - e.z ) );
	b.z += ( - dt_dx * ( e1.y - e.y ) );
	B[i + 1] = b;
}
// i in [0, nx+1]
__global__ void k1_yee_e(f3* __restrict__ E, const f3* __restrict__ B, const f3* __restrict__ J, int nx, float dt_dx, float dt) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i > nx + 1) return;
	f3 b = B[i + 1], bm = B[i], j = J[i + 1], e = E[i + 1];
	e.x += (                               - dt * j.x );
	e.y += ( - dt_dx * ( b.z - bm.z ) - dt * j.y );
	e.z += ( + dt_dx * ( b.y - bm.y ) - dt * j.z );
	E[i + 1] = e;
}

// first-order Mur absorbing boundary (em1d/emf.c:379-408); single thread, four cells
__global__ void k1_mur(f3* __restrict__ E, f3* __restrict__ mur, int nx, float S) {
	f3* Ec = E + 1;                         // cell 0
	f3* fld = mur; f3* tmp = mur + 2;
	fld[0].y = tmp[0].y + S * (Ec[0].y - fld[0].y);
	fld[0].z = tmp[0].z + S * (Ec[0].z - fld[0].z);
	Ec[-1].y = fld[0].y; Ec[-1].z = fld[0].z;
	tmp[0].y = Ec[0].y; tmp[0].z = Ec[0].z;
	fld[1].y = tmp[1].y + S * (Ec[nx - 1].y - fld[1].y);
	fld[1].z = tmp[1].z + S * (Ec[nx - 1].z - fld[1].z);
	Ec[nx].y = fld[1].y; Ec[nx].z = fld[1].z;
	tmp[1].y = Ec[nx - 1].y; tmp[1].z = Ec[nx - 1].z;
}

// periodic guards: E[-1] = E[nx-1]; upper loop runs to gc[0] = 1, so only E[nx] = E[0] (em1d/emf.c:476-500)
__global__ void k1_emf_gc(f3* __restrict__ E, f3* __restrict__ B, int nx) {
	f3* e = E + 1; f3* b = B + 1;
	e[-1] = e[nx - 1]; b[-1] = b[nx - 1];
	e[nx] = e[0];      b[nx] = b[0];
}

// new[i] = old[i+1] for i in [-1, nx]; cells nx-1..nx+1 zeroed (em1d/emf.c:519-531)
__global__ void k1_shift(f3* __restrict__ dst, const f3* __restrict__ src, int nx) {
	int c = blockIdx.x * blockDim.x + threadIdx.x;      // buffer index = cell + 1
	if (c > nx + 2) return;
	f3 z = {0.f, 0.f, 0.f};
	dst[c] = (c < nx) ? src[c + 1] : z;
}

__global__ void k1_add_uniform(f3* __restrict__ dst, const f3* __restrict__ src, int n, f3 v) {
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	f3 e = src[k]; e.x += v.x; e.y += v.y; e.z += v.z; dst[k] = e;
}
__global__ void k1_add_grid(f3* __restrict__ dst, const f3* __restrict__ src, const f3* __restrict__ ext, int n) {
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	f3 e = src[k], v = ext[k]; e.x += v.x; e.y += v.y; e.z += v.z; dst[k] = e;
}

static void update_part_fld(zdev_grid1d* g) {
	int grd = zdev_div_up(g->n, 256);
	if (g->e_ext == 1) ZDEV_LAUNCH(k1_add_uniform, grd, 256, 0, g->Epart, g->E, g->n, g->e0);
	else if (g->e_ext == 2) ZDEV_LAUNCH(k1_add_grid, grd, 256, 0, g->Epart, g->E, g->Eext, g->n);
	if (g->b_ext == 1) ZDEV_LAUNCH(k1_add_uniform, grd, 256, 0, g->Bpart, g->B, g->n, g->b0);
	else if (g->b_ext == 2) ZDEV_LAUNCH(k1_add_grid, grd, 256, 0, g->Bpart, g->B, g->Bext, g->n);
}

static void set_ext(zdev_grid1d* g, int is_b, int mode, const float v[3], const float* host_grid) {
	need_EB(g);
	f3** part = is_b ? &g->Bpart : &g->Epart;
	f3*  self = is_b ? g->B : g->E;
	int* flag = is_b ? &g->b_ext : &g->e_ext;
	f3** ext  = is_b ? &g->Bext : &g->Eext;
	if (*flag && !mode) { cudaFree(*part); *part = self; }
	if (!*flag && mode) ZDEV_CHECK(cudaMalloc(part, (size_t) g->n * sizeof(f3)));
	if (!mode) *part = self;
	*flag = mode;
	if (mode == 1) { f3 t = {v[0], v[1], v[2]}; if (is_b) g->b0 = t; else g->e0 = t; }
	if (mode == 2) {
		if (!*ext) ZDEV_CHECK(cudaMalloc(ext, (size_t) g->n * sizeof(f3)));
		ZDEV_CHECK(cudaMemcpyAsync(*ext, host_grid, (size_t) g->n * sizeof(f3), cudaMemcpyHostToDevice, zdev_strm));
		ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	}
}
extern "C" void zdev_emf1d_set_ext_uniform(zdev_grid1d* g, int e_on, const float e0[3], int b_on, const float b0[3]) {
	set_ext(g, 0, e_on ? 1 : 0, e0, nullptr);
	set_ext(g, 1, b_on ? 1 : 0, b0, nullptr);
	update_part_fld(g);
}
extern "C" void zdev_emf1d_set_ext_grid(zdev_grid1d* g, const float* he, const float* hb) {
	if (he) set_ext(g, 0, 2, nullptr, he);
	if (hb) set_ext(g, 1, 2, nullptr, hb);
	update_part_fld(g);
}

extern "C" void zdev_emf1d_set_mur(zdev_grid1d* g, const float st[12]) {
	ZDEV_CHECK(cudaMemcpyAsync(g->mur, st, 12 * sizeof(float), cudaMemcpyHostToDevice, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
}
extern "C" void zdev_emf1d_get_mur(zdev_grid1d* g, float st[12]) {
	ZDEV_CHECK(cudaMemcpyAsync(st, g->mur, 12 * sizeof(float), cudaMemcpyDeviceToHost, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
}

extern "C" void zdev_emf1d_advance(zdev_grid1d* g, zdev_grid1d* gj, float dt, float dx, int bc_type, int shift_window) {
	need_EB(g); need_J(gj);
	if (g->nx != gj->nx) { fprintf(stderr, "(*error*) zpic-b200: field / current grid size mismatch\n"); exit(-1); }
	const float dth = dt / 2.0f;
	const int grd = zdev_div_up(g->nx + 2, 256);
	ZDEV_LAUNCH(k1_yee_b, grd, 256, 0, g->B, g->E, g->nx, dth / dx);
	ZDEV_LAUNCH(k1_yee_e, grd, 256, 0, g->E, g->B, gj->J, g->nx, dt / dx, dt);
	if (bc_type == 2) ZDEV_LAUNCH(k1_mur, 1, 1, 0, g->E, g->mur, g->nx, (dt - dx) / (dt + dx));
	ZDEV_LAUNCH(k1_yee_b, grd, 256, 0, g->B, g->E, g->nx, dth / dx);
	if (bc_type == 1) ZDEV_LAUNCH(k1_emf_gc, 1, 1, 0, g->E, g->B, g->nx);
	update_part_fld(g);
	if (shift_window) {
		need_tmp(g);
		const int alias_e = (g->Epart == g->E), alias_b = (g->Bpart == g->B);
		const int gs = zdev_div_up(g->n, 256);
		ZDEV_LAUNCH(k1_shift, gs, 256, 0, g->tmp, g->E, g->nx);
		{ f3* t = g->E; g->E = g->tmp; g->tmp = t; }
		ZDEV_LAUNCH(k1_shift, gs, 256, 0, g->tmp, g->B, g->nx);
		{ f3* t = g->B; g->B = g->tmp; g->tmp = t; }
		if (alias_e) g->Epart = g->E;
		if (alias_b) g->Bpart = g->B;
	}
}

// ------------------------------------------------------------------ current (em1d/current.c:112-132, 265-333)

__global__ void k1_fold(f3* __restrict__ J, int nx) {
	f3* a = J + 1;
	for (int i = -1; i < 2; i++) {
		f3 lo = a[i], up = a[nx + i];
		lo.x += up.x; lo.y += up.y; lo.z += up.z;
		a[i] = lo; a[nx + i] = lo;
	}
}

// one [sa,sb,sa] pass, out of place; guards refreshed from the filtered interior when periodic
__global__ void k1_smooth(f3* __restrict__ dst, const f3* __restrict__ src, int nx, float sa, float sb, int periodic) {
	int i = blockIdx.x * blockDim.x + threadIdx.x - 1;
	if (i > nx + 1) return;
	int iw = i;
	if (i < 0 || i >= nx) {
		if (!periodic) { dst[i + 1] = src[i + 1]; return; }
		iw = (i < 0) ? i + nx : i - nx;
	}
	f3 fl = src[iw], f0 = src[iw + 1], fu = src[iw + 2], fs;
	fs.x = sa * fl.x + sb * f0.x + sa * fu.x;
	fs.y = sa * fl.y + sb * f0.y + sa * fu.y;
	fs.z = sa * fl.z + sb * f0.z + sa * fu.z;
	dst[i + 1] = fs;
}

extern "C" void zdev_current1d_update(zdev_grid1d* g, int bc_periodic, int xtype, int xlevel) {
	need_J(g);
	if (bc_periodic) ZDEV_LAUNCH(k1_fold, 1, 1, 0, g->J, g->nx);
	if (xtype != 0) {
		need_tmp(g);
		const int grd = zdev_div_up(g->nx + 3, 256);
		for (int k = 0; k < xlevel; k++) {
			ZDEV_LAUNCH(k1_smooth, grd, 256, 0, g->tmp, g->J, g->nx, 0.25f, 0.5f, bc_periodic);
			f3* t = g->J; g->J = g->tmp; g->tmp = t;
		}
		if (xtype == 2) {
			// reference get_smooth_comp, em1d/current.c:245-255
			float a = -1, b = (float) ((4.0 + 2.0 * xlevel) / xlevel), total = 2 * a + b;
			ZDEV_LAUNCH(k1_smooth, grd, 256, 0, g->tmp, g->J, g->nx, a / total, b / total, bc_periodic);
			f3* t = g->J; g->J = g->tmp; g->tmp = t;
		}
	}
}

// ------------------------------------------------------------------ energy (em1d/emf.c:600-620)

__global__ void k1_energy(const f3* __restrict__ E, const f3* __restrict__ B, int nx, double* out) {
	double s[6] = {0, 0, 0, 0, 0, 0};
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nx; i += gridDim.x * blockDim.x) {
		f3 e = E[i + 1], b = B[i + 1];
		s[0] += e.x * e.x; s[1] += e.y * e.y; s[2] += e.z * e.z;
		s[3] += b.x * b.x; s[4] += b.y * b.y; s[5] += b.z * b.z;
	}
	#pragma unroll
	for (int q = 0; q < 6; q++) {
		double v = s[q];
		for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
		if ((threadIdx.x & 31) == 0) atomicAdd(&out[q], v);
	}
}

extern "C" void zdev_emf1d_energy(zdev_grid1d* g, double sums[6]) {
	need_EB(g);
	ZDEV_CHECK(cudaMemsetAsync(g->d_sums, 0, 6 * sizeof(double), zdev_strm));
	int grd = zdev_div_up(g->nx, 256);
	if (grd > 2 * zdev_num_sm) grd = 2 * zdev_num_sm;
	ZDEV_LAUNCH(k1_energy, grd, 256, 0, g->E, g->B, g->nx, g->d_sums);
	ZDEV_CHECK(cudaMemcpyAsync(sums, g->d_sums, 6 * sizeof(double), cudaMemcpyDeviceToHost, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
}

// accessors for zdev_spec1d.cu
f3* zdev_grid1d_Epart(zdev_grid1d* g) { need_EB(g); return g->Epart; }
f3* zdev_grid1d_Bpart(zdev_grid1d* g) { need_EB(g); return g->Bpart; }
f3* zdev_grid1d_J(zdev_grid1d* g) { need_J(g); return g->J; }
int zdev_grid1d_nx(zdev_grid1d* g) { return g->nx; }
