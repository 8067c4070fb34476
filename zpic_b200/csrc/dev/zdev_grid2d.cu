// zpic-b200 :: em2d field and current grids on the device.
//
// Layout: identical to the reference host buffers - (nx+3) x (ny+3) AoS float3,
// guards {1 lower, 2 upper}, row stride nrow = nx+3 (reference em2d/emf.c:59-89,
// em2d/current.c:33-53) - so host mirrors are straight cudaMemcpy's and the
// reference loop bounds carry over literally.  All kernels are pure HBM-streaming
// stencils: one thread per cell, x fastest so a warp touches one contiguous
// 384-byte span per float3 array.  Compiled with --fmad=false: every expression
// below keeps the reference's operation order so results are bit-identical.
#include "zdev_common.cuh"
#include "zdev_slab.cuh"
#include <algorithm>

struct zdev_grid2d {
	int nx, ny, nrow, nrows;
	size_t ncell;            // (nx+3)*(ny+3)
	f3 *E, *B, *J;           // point at buffer start, i.e. cell (-1,-1)
	f3 *tmp;                 // scratch of the same size (smoothing ping-pong, window shift, fused field advance)
	f3 *tmp2;                // second scratch, only for the fused field advance (E and B both move out of place)
	f3 *Epart, *Bpart;       // fields seen by particles; alias E/B unless external fields are on
	f3 *Eext, *Bext;         // cached custom external fields (or null)
	int e_ext, b_ext;        // 0 none, 1 uniform, 2 grid
	f3 e0, b0;
	double* d_sums;          // 6 doubles
	// slab decomposition along x (zdev_slab.cuh): this grid is one slab of a wider box
	int slab;                // 1: guard columns towards a neighbour slab are exchanged, not wrapped
	int wrap_left, wrap_right;   // that edge of the slab is the (periodic) box boundary
	zdev_link link;
};

// cell (i,j), i in [-1,nx+1], j in [-1,ny+1] -> linear index from buffer start
__device__ __forceinline__ int cidx(int i, int j, int nrow) { return (i + 1) + (j + 1) * nrow; }

extern "C" zdev_grid2d* zdev_grid2d_create(int nx, int ny) {
	zdev_require_init();
	zdev_grid2d* g = (zdev_grid2d*) calloc(1, sizeof(zdev_grid2d));
	g->nx = nx; g->ny = ny; g->nrow = nx + 3; g->nrows = ny + 3;
	g->ncell = (size_t) g->nrow * g->nrows;
	ZDEV_CHECK(cudaMalloc(&g->d_sums, 6 * sizeof(double)));
	return g;
}

// Buffers are allocated (zeroed) on first use: a grid object that only backs a
// t_current never pays for E/B and vice versa.
static f3* grid_alloc_zero(zdev_grid2d* g) {
	f3* p; size_t bytes = g->ncell * sizeof(f3);
	ZDEV_CHECK(cudaMalloc(&p, bytes));
	ZDEV_CHECK(cudaMemsetAsync(p, 0, bytes, zdev_strm));
	return p;
}
static void need_EB(zdev_grid2d* g) {
	if (g->E) return;
	g->E = grid_alloc_zero(g); g->B = grid_alloc_zero(g);
	g->Epart = g->E; g->Bpart = g->B;
}
static void need_J(zdev_grid2d* g) { if (!g->J) g->J = grid_alloc_zero(g); }
static void need_tmp(zdev_grid2d* g) { if (!g->tmp) g->tmp = grid_alloc_zero(g); }
static void need_tmp2(zdev_grid2d* g) { if (!g->tmp2) g->tmp2 = grid_alloc_zero(g); }

extern "C" void zdev_grid2d_destroy(zdev_grid2d* g) {
	if (!g) return;
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	if (g->e_ext) cudaFree(g->Epart);
	if (g->b_ext) cudaFree(g->Bpart);
	cudaFree(g->Eext); cudaFree(g->Bext);
	cudaFree(g->E); cudaFree(g->B); cudaFree(g->J); cudaFree(g->tmp); cudaFree(g->tmp2); cudaFree(g->d_sums);
	if (g->slab) zdev_link_close(&g->link);
	free(g);
}

static f3* grid_sel(zdev_grid2d* g, int which) {
	if (which == ZDEV_J) need_J(g); else need_EB(g);
	switch (which) {
	case ZDEV_E: return g->E;
	case ZDEV_B: return g->B;
	case ZDEV_J: return g->J;
	case ZDEV_EPART: return g->Epart;
	case ZDEV_BPART: return g->Bpart;
	}
	fprintf(stderr, "(*error*) zdev_grid2d: invalid grid selector %d\n", which); exit(-1);
}

extern "C" void zdev_grid2d_upload(zdev_grid2d* g, int which, const float* host_buf) {
	ZDEV_CHECK(cudaMemcpyAsync(grid_sel(g, which), host_buf, g->ncell * sizeof(f3), cudaMemcpyHostToDevice, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
}
extern "C" void zdev_grid2d_download(zdev_grid2d* g, int which, float* host_buf) {
	ZDEV_CHECK(cudaMemcpyAsync(host_buf, grid_sel(g, which), g->ncell * sizeof(f3), cudaMemcpyDeviceToHost, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
}
extern "C" float* zdev_grid2d_ptr(zdev_grid2d* g, int which) { return (float*) grid_sel(g, which); }

// the slab's window of a wider host buffer (row stride host_nrow cells): local buffer column c = host buffer
// column x0 + c, guards included (the host mirrors of a decomposed run are global, the device grids local)
extern "C" void zdev_grid2d_upload_window(zdev_grid2d* g, int which, const float* host_buf, int host_nrow, int x0) {
	ZDEV_CHECK(cudaMemcpy2DAsync(grid_sel(g, which), (size_t) g->nrow * sizeof(f3), (const f3*) host_buf + x0,
	                             (size_t) host_nrow * sizeof(f3), (size_t) g->nrow * sizeof(f3), g->nrows,
	                             cudaMemcpyHostToDevice, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
}
extern "C" void zdev_grid2d_download_window(zdev_grid2d* g, int which, float* host_buf, int host_nrow, int x0) {
	ZDEV_CHECK(cudaMemcpy2DAsync((f3*) host_buf + x0, (size_t) host_nrow * sizeof(f3), grid_sel(g, which),
	                             (size_t) g->nrow * sizeof(f3), (size_t) g->nrow * sizeof(f3), g->nrows,
	                             cudaMemcpyDeviceToHost, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
}

// ------------------------------------------------------------------ slab links (zdev_slab.cuh)

static const int SMOOTH_HALO_MAX = 16;        // the wide halo of the fused smoothing passes: passes + 2 columns at most

extern "C" void zdev_grid2d_set_slab(zdev_grid2d* g, int left, int right, int wrap_left, int wrap_right) {
	if (g->slab) return;
	// the largest messages: three columns of two grids, or the wide halo of the fused smoothing passes
	// (SMOOTH_HALO_MAX columns of one grid), every row
	zdev_link_open(&g->link, (size_t) SMOOTH_HALO_MAX * g->nrows * sizeof(f3), left, right);
	g->slab = 1; g->wrap_left = wrap_left; g->wrap_right = wrap_right;
}

struct slab_msg { f3* buf; unsigned* flag; int i0, ncols; };      // one side of an exchange; buf == nullptr: nothing

// blockIdx.y = side.  Columns [i0, i0+ncols) x rows [j0, j0+nrows) of up to two grids -> the neighbour's payload
__global__ void k_slab_send(const f3* __restrict__ G0, const f3* __restrict__ G1, int ngrids, int nrow, int j0, int nrows,
                            slab_msg L, slab_msg R, unsigned seq_l, unsigned seq_r, unsigned* ticket) {
	const slab_msg m = blockIdx.y ? R : L;
	const unsigned seq = blockIdx.y ? seq_r : seq_l;
	if (!m.buf) return;
	const int per = m.ncols * nrows, n = ngrids * per;
	for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
		const int gsel = k / per, q = k - gsel * per, r = q / m.ncols, c = q - r * m.ncols;
		m.buf[k] = (gsel ? G1 : G0)[cidx(m.i0 + c, j0 + r, nrow)];
	}
	slab_publish(ticket + blockIdx.y, gridDim.x, m.flag, seq);
}
// ... and my payload -> my guard columns (add = 1: the J fold)
__global__ void k_slab_recv(f3* __restrict__ G0, f3* __restrict__ G1, int ngrids, int nrow, int j0, int nrows,
                            slab_msg L, slab_msg R, unsigned seq_l, unsigned seq_r, int add) {
	const slab_msg m = blockIdx.y ? R : L;
	const unsigned seq = blockIdx.y ? seq_r : seq_l;
	if (!m.buf) return;
	slab_wait(m.flag, seq);
	const int per = m.ncols * nrows, n = ngrids * per;
	for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
		const int gsel = k / per, q = k - gsel * per, r = q / m.ncols, c = q - r * m.ncols;
		const float* src = reinterpret_cast<const float*>(m.buf + k);
		f3 v; v.x = __ldcg(src); v.y = __ldcg(src + 1); v.z = __ldcg(src + 2);       // written by another GPU: past L1
		f3* d = &(gsel ? G1 : G0)[cidx(m.i0 + c, j0 + r, nrow)];
		if (add) { f3 o = *d; o.x += v.x; o.y += v.y; o.z += v.z; *d = o; } else *d = v;
	}
}

// One exchange with both neighbours: columns [sl, sl+nsl) go to the left neighbour and [sr, sr+nsr) to the right
// one; what the left neighbour sent lands in [rl, rl+nrl), the right one's in [rr, rr+nrr).
static void slab_exchange(zdev_grid2d* g, f3* G0, f3* G1, int sl, int nsl, int sr, int nsr, int rl, int nrl, int rr, int nrr,
                          int j0, int nrows, int add, bool do_l, bool do_r) {
	zdev_link& K = g->link;
	do_l = do_l && K.left >= 0; do_r = do_r && K.right >= 0;
	if (!do_l && !do_r) return;
	const int ngrids = G1 ? 2 : 1;
	slab_msg SL = { nullptr, nullptr, 0, 0 }, SR = SL, RL = SL, RR = SL;
	unsigned seq_l = 0, seq_r = 0;
	if (do_l) {
		seq_l = ++K.seq[0];
		SL = { (f3*) zdev_link_out(K, 0, seq_l), &zdev_link_out_hdr(K, 0)->flag[1], sl, nsl };
		RL = { (f3*) zdev_link_in(K, 0, seq_l), &zdev_link_in_hdr(K)->flag[0], rl, nrl };
	}
	if (do_r) {
		seq_r = ++K.seq[1];
		SR = { (f3*) zdev_link_out(K, 1, seq_r), &zdev_link_out_hdr(K, 1)->flag[0], sr, nsr };
		RR = { (f3*) zdev_link_in(K, 1, seq_r), &zdev_link_in_hdr(K)->flag[1], rr, nrr };
	}
	const int nmax = ngrids * 3 * nrows;
	dim3 grd(std::max(1, std::min(zdev_div_up(nmax, 256), 32)), 2);
	ZDEV_LAUNCH(k_slab_send, grd, 256, 0, G0, G1, ngrids, g->nrow, j0, nrows, SL, SR, seq_l, seq_r, K.ticket);
	ZDEV_LAUNCH(k_slab_recv, grd, 256, 0, G0, G1, ngrids, g->nrow, j0, nrows, RL, RR, seq_l, seq_r, add);
}
// guards <- neighbour interior: (-1) <- left's (nx-1); (nx, nx+1) <- right's (0, 1)
static void slab_halo_refresh(zdev_grid2d* g, f3* G0, f3* G1, int j0, int nrows, bool skip_wrap) {
	slab_exchange(g, G0, G1, 0, 2, g->nx - 1, 1, -1, 1, g->nx, 2, j0, nrows, 0,
	              !(skip_wrap && g->wrap_left), !(skip_wrap && g->wrap_right));
}

extern "C" void zdev_current_zero(zdev_grid2d* g) {
	need_J(g);
	ZDEV_CHECK(cudaMemsetAsync(g->J, 0, g->ncell * sizeof(f3), zdev_strm));
}

// ------------------------------------------------------------------ Yee solver

// reference em2d/emf.c:500-522 : i in [-1,nx], j in [-1,ny]
__global__ void k_yee_b(f3* __restrict__ B, const f3* __restrict__ E, int nx, int ny, int nrow,
                        float dt_dx, float dt_dy) {
	int i = blockIdx.x * blockDim.x + threadIdx.x - 1;
	int j = blockIdx.y * blockDim.y + threadIdx.y - 1;
	if (i > nx || j > ny) return;
	int c = cidx(i, j, nrow);
	f3 e = E[c], ex = E[c + 1], ey = E[c + nrow], b = B[c];
	b.x += ( - dt_dy * ( ey.z - e.z ) );
	b.y += (   dt_dx * ( ex.z - e.z ) );
	b.z += ( - dt_dx * ( ex.y - e.y ) + dt_dy * ( ey.x - e.x ) );
	B[c] = b;
}

// reference em2d/emf.c:531-562 : i in [0,nx+1], j in [0,ny+1]
__global__ void k_yee_e(f3* __restrict__ E, const f3* __restrict__ B, const f3* __restrict__ J,
                        int nx, int ny, int nrow, float dt_dx, float dt_dy, float dt) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	int j = blockIdx.y * blockDim.y + threadIdx.y;
	if (i > nx + 1 || j > ny + 1) return;
	int c = cidx(i, j, nrow);
	f3 b = B[c], bx = B[c - 1], by = B[c - nrow], jc = J[c], e = E[c];
	e.x += ( + dt_dy * ( b.z - by.z ) ) - dt * jc.x;
	e.y += ( - dt_dx * ( b.z - bx.z ) ) - dt * jc.y;
	e.z += ( + dt_dx * ( b.y - bx.y ) - dt_dy * ( b.x - by.x ) ) - dt * jc.z;
	E[c] = e;
}

extern "C" void zdev_yee_b(zdev_grid2d* g, float dt_dx, float dt_dy) {
	need_EB(g);
	dim3 blk(64, 4), grd(zdev_div_up(g->nx + 2, 64), zdev_div_up(g->ny + 2, 4));
	ZDEV_LAUNCH(k_yee_b, grd, blk, 0, g->B, g->E, g->nx, g->ny, g->nrow, dt_dx, dt_dy);
}
static void check_same_shape(zdev_grid2d* a, zdev_grid2d* b) {
	if (a->nx != b->nx || a->ny != b->ny) { fprintf(stderr, "(*error*) zpic-b200: field / current grid size mismatch\n"); exit(-1); }
}
extern "C" void zdev_yee_e(zdev_grid2d* g, zdev_grid2d* gj, float dt_dx, float dt_dy, float dt) {
	need_EB(g); need_J(gj); check_same_shape(g, gj);
	dim3 blk(64, 4), grd(zdev_div_up(g->nx + 2, 64), zdev_div_up(g->ny + 2, 4));
	ZDEV_LAUNCH(k_yee_e, grd, blk, 0, g->E, g->B, gj->J, g->nx, g->ny, g->nrow, dt_dx, dt_dy, dt);
}

// ---- yee_b(dt/2), yee_e(dt), yee_b(dt/2) in ONE pass over the grids (reference em2d/emf.c:694-698).
// The three stencils separately move 132 B per cell (E is read twice and rewritten once, B is rewritten
// twice and read a third time); fused they move 60: every CTA stages the E and B neighbourhood of its tile
// in shared memory, applies the three updates there with the reference's own loop bounds as masks, and
// writes the tile of the new E and B OUT OF PLACE (a neighbouring CTA may still be reading the old halo).
//   new B on the tile T needs new E on T + its upper neighbours, new E there needs the half-step B on one
//   more lower layer, and that needs old E on one more upper layer:
//   sE = old E on [A-1, A+W+1] x [R-1, R+H+1],  sB = old B on [A-1, A+W] x [R-1, R+H]  (buffer indices).
// The halo results are computed redundantly with exactly the arithmetic of their owner CTA, so every cell
// (guards included) is bit-identical to the three separate kernels.
#ifndef YF_H_N
#define YF_H_N 16
#endif
#ifndef YF_TY_N
#define YF_TY_N 4
#endif
constexpr int YF_W = 64, YF_H = YF_H_N, YF_TY = YF_TY_N, YF_THREADS = YF_W * YF_TY;
constexpr int YF_EW = YF_W + 3, YF_EH = YF_H + 3, YF_BW = YF_W + 2, YF_BH = YF_H + 2, YF_JW = YF_W + 1, YF_JH = YF_H + 1;

// 4-byte asynchronous copy global -> shared (LDGSTS): no register, no wait; ok = false writes a zero
__device__ __forceinline__ void yf_cp4(float* sdst, const float* gsrc, bool ok) {
	const unsigned sa = (unsigned) __cvta_generic_to_shared(sdst);
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(sa), "l"(gsrc), "r"(ok ? 4 : 0) : "memory");
}
// yee_b on one cell: e = E here, ex = E at i+1, ey = E at j+1 (reference em2d/emf.c:514-520)
__device__ __forceinline__ void yf_b(f3& b, const f3 e, const f3 ex, const f3 ey, float dt_dx, float dt_dy) {
	b.x += ( - dt_dy * ( ey.z - e.z ) );
	b.y += (   dt_dx * ( ex.z - e.z ) );
	b.z += ( - dt_dx * ( ex.y - e.y ) + dt_dy * ( ey.x - e.x ) );
}

// The neighbourhoods arrive as flat runs of floats per buffer row through asynchronous 4-byte copies (rows
// of the reference layout are only 4-byte aligned, so no bulk copies): every thread issues all its copies
// back to back and waits once, instead of one DRAM round trip per row.  In the stencil phases thread (tx, ty)
// of the 64 x YF_TY CTA walks rows ty, ty+YF_TY, ... at column tx; the one or two extra halo columns of a
// region are taken by the first threads of each row.  12-byte cells in shared memory are conflict free
// (stride 3 words).
__global__ void __launch_bounds__(YF_THREADS)
k_yee_fused(f3* __restrict__ Eout, f3* __restrict__ Bout, const f3* __restrict__ E, const f3* __restrict__ B,
            const f3* __restrict__ J, int nrow, int nrows, float hdt_dx, float hdt_dy, float dt_dx, float dt_dy, float dt) {
	__shared__ f3 sE[YF_EH * YF_EW];
	__shared__ f3 sB[YF_BH * YF_BW];
	__shared__ f3 sJ[YF_JH * YF_JW];
	const int A = blockIdx.x * YF_W, R = blockIdx.y * YF_H;       // buffer column / row of the tile origin
	const int tx = threadIdx.x, ty = threadIdx.y;
	{	// ---- old E on [A-1, A+W+1] x [R-1, R+H+1], old B on [A-1, A+W] x [R-1, R+H], J on [A, A+W] x [R, R+H]
		const int tid = ty * YF_W + tx;
		const long rowf = (long) nrow * 3;                          // floats per buffer row
		const float* gE = reinterpret_cast<const float*>(E);
		const float* gB = reinterpret_cast<const float*>(B);
		const float* gJ = reinterpret_cast<const float*>(J);
		float* fE = reinterpret_cast<float*>(sE);
		float* fB = reinterpret_cast<float*>(sB);
		float* fJ = reinterpret_cast<float*>(sJ);
		#pragma unroll 1
		for (int r = 0; r < YF_EH; r++) {
			const int bj = R - 1 + r;
			const bool rok = bj >= 0 && bj < nrows;
			for (int f = tid; f < YF_EW * 3; f += YF_THREADS) {
				const long fo = (long) (A - 1) * 3 + f;             // float offset inside the buffer row
				const bool ok = rok && fo >= 0 && fo < rowf;
				const long go = ok ? (long) bj * rowf + fo : 0;
				yf_cp4(fE + r * (YF_EW * 3) + f, gE + go, ok);
				if (r < YF_BH && f < YF_BW * 3) yf_cp4(fB + r * (YF_BW * 3) + f, gB + go, ok);
			}
		}
		#pragma unroll 1
		for (int r = 0; r < YF_JH; r++) {
			const int bj = R + r;
			for (int f = tid; f < YF_JW * 3; f += YF_THREADS) {
				const long fo = (long) A * 3 + f;
				const bool ok = bj < nrows && fo < rowf;
				yf_cp4(fJ + r * (YF_JW * 3) + f, gJ + (ok ? (long) bj * rowf + fo : 0), ok);
			}
		}
		asm volatile("cp.async.wait_all;" ::: "memory");
	}
	__syncthreads();
	// ---- yee_b(dt/2) on the whole sB region: cells i in [-1,nx], j in [-1,ny] = buffer [0,nrow-2] x [0,nrows-2]
	#pragma unroll
	for (int r = ty; r < YF_BH; r += YF_TY) {
		const int bj = R - 1 + r;
		const bool rok = bj >= 0 && bj <= nrows - 2;
		#pragma unroll
		for (int h = 0; h < 2; h++) {
			const int c = h ? YF_W + tx : tx;
			if (h && tx >= 2) break;
			const int bi = A - 1 + c;
			if (rok && bi >= 0 && bi <= nrow - 2) {
				f3 b = sB[r * YF_BW + c];
				yf_b(b, sE[r * YF_EW + c], sE[r * YF_EW + c + 1], sE[(r + 1) * YF_EW + c], hdt_dx, hdt_dy);
				sB[r * YF_BW + c] = b;
			}
		}
	}
	__syncthreads();
	// ---- yee_e(dt) on the tile and its upper neighbours: cells i in [0,nx+1], j in [0,ny+1] = buffer [1,nrow-1] x [1,nrows-1]
	#pragma unroll
	for (int rr = ty; rr < YF_JH; rr += YF_TY) {
		const int bj = R + rr;
		const bool rok = bj >= 1 && bj <= nrows - 1;
		#pragma unroll
		for (int h = 0; h < 2; h++) {
			const int cc = h ? YF_W : tx;
			if (h && tx >= 1) break;
			const int bi = A + cc;
			if (rok && bi >= 1 && bi <= nrow - 1) {
				const int r = rr + 1, c = cc + 1;                   // position in sE / sB
				const f3 b = sB[r * YF_BW + c], bx = sB[r * YF_BW + c - 1], by = sB[(r - 1) * YF_BW + c];
				const f3 jc = sJ[rr * YF_JW + cc];
				f3 e = sE[r * YF_EW + c];
				e.x += ( + dt_dy * ( b.z - by.z ) ) - dt * jc.x;
				e.y += ( - dt_dx * ( b.z - bx.z ) ) - dt * jc.y;
				e.z += ( + dt_dx * ( b.y - bx.y ) - dt_dy * ( b.x - by.x ) ) - dt * jc.z;
				sE[r * YF_EW + c] = e;
			}
		}
	}
	__syncthreads();
	// ---- yee_b(dt/2) on the tile from the new E; both grids leave from registers
	#pragma unroll
	for (int rr = ty; rr < YF_H; rr += YF_TY) {
		const int bi = A + tx, bj = R + rr;
		if (bi < nrow && bj < nrows) {
			const int r = rr + 1, c = tx + 1;
			const f3 e = sE[r * YF_EW + c];
			f3 b = sB[r * YF_BW + c];
			if (bi <= nrow - 2 && bj <= nrows - 2)
				yf_b(b, e, sE[r * YF_EW + c + 1], sE[(r + 1) * YF_EW + c], hdt_dx, hdt_dy);
			const long o = (long) bj * nrow + bi;
			Eout[o] = e;
			Bout[o] = b;
		}
	}
}

// The same one-pass field advance with the rows in REGISTERS (mode 2).  A warp owns 32 consecutive buffer
// columns, one per lane, and marches up a stripe of rows: per step it loads one row of E (the row above), B and J,
// and has everything else in registers - the half-step B of this row and the one below, the new E of this row and
// the one below.  Neighbours along x come from the lanes next door (six shuffles per cell and step), neighbours
// along y from the previous / next step:
//     Bh[j] = B[j]  + yee_b(E[j],  E[j](i+1),  E[j+1])           valid on lanes 0..30
//     En[j] = E[j]  + yee_e(Bh[j], Bh[j](i-1), Bh[j-1], J[j])    valid on lanes 1..30
//     Bn[j-1] = Bh[j-1] + yee_b(En[j-1], En[j-1](i+1), En[j])    valid on lanes 1..29  -> rows j-1 of E and B leave
// so 29 columns per warp come out (consecutive warps overlap by three lanes) and a stripe of H rows costs three
// extra row loads.  The tile kernel above needs ~16 warp-instructions per cell (asynchronous 4-byte copies into
// shared memory, three passes over the tile with index arithmetic and masks); this one ~3.5 and no shared memory
// or barrier at all, which leaves the DRAM traffic (60 B per cell) as the bound.  Every cell, guards included,
// goes through the expressions of yee_b / yee_e (em2d/emf.c:509-520, 537-559) on the same operands: bit-identical.
#define YM_COLS 29
__device__ __forceinline__ f3 ym_load(const f3* __restrict__ G, int bi, int bj, int nrow, int nrows) {
	f3 v = {0.0f, 0.0f, 0.0f};
	if (bi >= 0 && bi < nrow && bj >= 0 && bj < nrows) v = G[(long) bj * nrow + bi];
	return v;
}
__device__ __forceinline__ f3 ym_shfl_down(const f3 v) {
	f3 r; r.x = __shfl_down_sync(0xffffffffu, v.x, 1); r.y = __shfl_down_sync(0xffffffffu, v.y, 1); r.z = __shfl_down_sync(0xffffffffu, v.z, 1);
	return r;
}
__global__ void __launch_bounds__(256)
k_yee_march(f3* __restrict__ Eout, f3* __restrict__ Bout, const f3* __restrict__ E, const f3* __restrict__ B,
            const f3* __restrict__ J, int nrow, int nrows, int H, float hdt_dx, float hdt_dy, float dt_dx, float dt_dy, float dt) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int cg = blockIdx.x * 8 + warp;                          // column group: buffer columns cg*29 .. cg*29+28
	if (cg * YM_COLS >= nrow) return;
	const int bi = cg * YM_COLS - 1 + lane;                        // lane 1 holds the group's first column
	const int R = blockIdx.y * H;
	const bool col_b = bi >= 0 && bi <= nrow - 2;                  // yee_b: i in [-1,nx] = buffer [0, nrow-2]
	const bool col_e = bi >= 1 && bi <= nrow - 1;                  // yee_e: i in [0,nx+1] = buffer [1, nrow-1]
	const bool mine = lane >= 1 && lane <= YM_COLS && bi < nrow;
	f3 Ec = ym_load(E, bi, R - 1, nrow, nrows);                    // E of the current row
	f3 Bh_prev = {0, 0, 0}, En_prev = {0, 0, 0};
	const int r_end = min(R + H, nrows);                           // rows [R, r_end) leave this stripe
	// (issuing the three row loads of step r+1 before the arithmetic of step r was measured: 0.2467 against 0.2455 ms
	//  per grid half at 4096^2, 40 instead of 32 registers - 64 resident warps per SM already cover the latency)
	for (int r = R - 1; r <= r_end; r++) {
		const f3 Eup = ym_load(E, bi, r + 1, nrow, nrows);
		f3 Bh = ym_load(B, bi, r, nrow, nrows);
		const f3 Jc = ym_load(J, bi, r, nrow, nrows);
		// half-step B of row r (yee_b: j in [-1,ny] = buffer rows [0, nrows-2])
		{
			const f3 Ex = ym_shfl_down(Ec);
			if (col_b && r >= 0 && r <= nrows - 2) yf_b(Bh, Ec, Ex, Eup, hdt_dx, hdt_dy);
		}
		// new E of row r (yee_e: j in [0,ny+1] = buffer rows [1, nrows-1]); rows below the stripe are not needed
		f3 En = Ec;
		{
			f3 bx;
			bx.y = __shfl_up_sync(0xffffffffu, Bh.y, 1); bx.z = __shfl_up_sync(0xffffffffu, Bh.z, 1);
			if (col_e && r >= max(R, 1) && r <= nrows - 1) {
				const f3 b = Bh, by = Bh_prev;
				En.x += ( + dt_dy * ( b.z - by.z ) ) - dt * Jc.x;
				En.y += ( - dt_dx * ( b.z - bx.z ) ) - dt * Jc.y;
				En.z += ( + dt_dx * ( b.y - bx.y ) - dt_dy * ( b.x - by.x ) ) - dt * Jc.z;
			}
		}
		// row r-1 is complete: second half step of B from the new E, and out
		{
			const f3 Ex = ym_shfl_down(En_prev);
			if (r - 1 >= R) {
				f3 Bn = Bh_prev;
				if (col_b && r - 1 <= nrows - 2) yf_b(Bn, En_prev, Ex, En, hdt_dx, hdt_dy);
				if (mine) {
					const long o = (long) (r - 1) * nrow + bi;
					Eout[o] = En_prev;
					Bout[o] = Bn;
				}
			}
		}
		Ec = Eup; Bh_prev = Bh; En_prev = En;
	}
}

static int fused_yee = -1;
static int fused_yee_on() {
	if (fused_yee < 0) { const char* e = getenv("ZPIC_FUSED_YEE"); fused_yee = !e ? 2 : (e[0] == '0' ? 0 : (e[0] == '1' ? 1 : 2)); }
	return fused_yee;
}
extern "C" void zdev_yee_set_fused(int on) { fused_yee = (on == 2) ? 2 : (on ? 1 : 0); }
static void yee_fused(zdev_grid2d* g, zdev_grid2d* gj, float dt, float dx, float dy) {
	need_EB(g); need_J(gj); check_same_shape(g, gj); need_tmp(g); need_tmp2(g);
	const float dtb = dt / 2.0f;
	const int alias_e = (g->Epart == g->E), alias_b = (g->Bpart == g->B);
	if (fused_yee_on() == 2) {
		// stripes of H rows: tall ones amortise the three extra row loads, short ones keep a small grid's SMs busy
		const int ncg = zdev_div_up(g->nrow, YM_COLS);
		int H = 64;
		while (H > 16 && (long) ncg * zdev_div_up(g->nrows, H) < (long) 32 * zdev_num_sm) H >>= 1;
		dim3 grd(zdev_div_up(ncg, 8), zdev_div_up(g->nrows, H));
		ZDEV_LAUNCH(k_yee_march, grd, 256, 0, g->tmp, g->tmp2, g->E, g->B, gj->J, g->nrow, g->nrows, H,
		            dtb / dx, dtb / dy, dt / dx, dt / dy, dt);
	} else {
		dim3 grd(zdev_div_up(g->nrow, YF_W), zdev_div_up(g->nrows, YF_H));
		ZDEV_LAUNCH(k_yee_fused, grd, dim3(YF_W, YF_TY), 0, g->tmp, g->tmp2, g->E, g->B, gj->J, g->nrow, g->nrows,
		            dtb / dx, dtb / dy, dt / dx, dt / dy, dt);
	}
	{ f3* t = g->E; g->E = g->tmp; g->tmp = t; }
	{ f3* t = g->B; g->B = g->tmp2; g->tmp2 = t; }
	if (alias_e) g->Epart = g->E;
	if (alias_b) g->Bpart = g->B;
}

// ------------------------------------------------------------------ guard cells

// periodic x copies for every row (reference em2d/emf.c:583-607): one thread per row
__global__ void k_gc_x_copy(f3* __restrict__ A, f3* __restrict__ Bf, int nx, int nrows, int nrow) {
	int r = blockIdx.x * blockDim.x + threadIdx.x;   // buffer row 0..nrows-1
	if (r >= nrows) return;
	f3* a = A + (size_t) r * nrow + 1;                // cell (0, r-1)
	a[-1] = a[nx - 1]; a[nx] = a[0]; a[nx + 1] = a[1];
	if (Bf) { f3* b = Bf + (size_t) r * nrow + 1; b[-1] = b[nx - 1]; b[nx] = b[0]; b[nx + 1] = b[1]; }
}
// periodic y copies for every column incl. x guards (reference em2d/emf.c:611-635)
__global__ void k_gc_y_copy(f3* __restrict__ A, f3* __restrict__ Bf, int ny, int nrow) {
	int c = blockIdx.x * blockDim.x + threadIdx.x;   // buffer column 0..nrow-1
	if (c >= nrow) return;
	size_t s = nrow;
	f3* a = A + c + s;                                // cell (c-1, 0)
	a[-(long) s] = a[(size_t)(ny - 1) * s]; a[(size_t) ny * s] = a[0]; a[(size_t)(ny + 1) * s] = a[s];
	if (Bf) { f3* b = Bf + c + s;
		b[-(long) s] = b[(size_t)(ny - 1) * s]; b[(size_t) ny * s] = b[0]; b[(size_t)(ny + 1) * s] = b[s]; }
}

extern "C" void zdev_emf_update_gc(zdev_grid2d* g, int moving_window) {
	need_EB(g);
	if (!moving_window)
		ZDEV_LAUNCH(k_gc_x_copy, zdev_div_up(g->nrows, 128), 128, 0, g->E, g->B, g->nx, g->nrows, g->nrow);
	ZDEV_LAUNCH(k_gc_y_copy, zdev_div_up(g->nrow, 128), 128, 0, g->E, g->B, g->ny, g->nrow);
}

// J fold: lower += upper, then upper = lower (reference em2d/current.c:124-157)
__global__ void k_fold_x(f3* __restrict__ J, int nx, int nrows, int nrow) {
	int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= nrows) return;
	f3* a = J + (size_t) r * nrow + 1;
	#pragma unroll
	for (int i = -1; i < 2; i++) {
		f3 lo = a[i], up = a[nx + i];
		lo.x += up.x; lo.y += up.y; lo.z += up.z;
		a[i] = lo; a[nx + i] = lo;
	}
}
__global__ void k_fold_y(f3* __restrict__ J, int ny, int nrow) {
	int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= nrow) return;
	size_t s = nrow;
	f3* a = J + c + s;
	#pragma unroll
	for (int j = -1; j < 2; j++) {
		f3 lo = a[(long) j * (long) s], up = a[(size_t)(ny + j) * s];
		lo.x += up.x; lo.y += up.y; lo.z += up.z;
		a[(long) j * (long) s] = lo; a[(size_t)(ny + j) * s] = lo;
	}
}

extern "C" void zdev_current_update_gc(zdev_grid2d* g, int moving_window) {
	need_J(g);
	if (!moving_window)
		ZDEV_LAUNCH(k_fold_x, zdev_div_up(g->nrows, 128), 128, 0, g->J, g->nx, g->nrows, g->nrow);
	ZDEV_LAUNCH(k_fold_y, zdev_div_up(g->nrow, 128), 128, 0, g->J, g->ny, g->nrow);
}

// ------------------------------------------------------------------ smoothing

__device__ __forceinline__ f3 stencil3(f3 fl, f3 f0, f3 fu, float sa, float sb) {
	f3 fs;   // reference em2d/current.c:334-336, left-to-right
	fs.x = sa * fl.x + sb * f0.x + sa * fu.x;
	fs.y = sa * fl.y + sb * f0.y + sa * fu.y;
	fs.z = sa * fl.z + sb * f0.z + sa * fu.z;
	return fs;
}

// One [sa,sb,sa] pass along x, out of place (reference kernel_x, current.c:316-354).
// Only rows 0..ny-1 are filtered; x guards of those rows are refreshed from the
// filtered interior unless the window moves; other rows pass through unchanged.
__global__ void k_smooth_x(f3* __restrict__ dst, const f3* __restrict__ src, int nx, int ny, int nrow,
                           float sa, float sb, int moving_window) {
	int i = blockIdx.x * blockDim.x + threadIdx.x - 1;
	int j = blockIdx.y * blockDim.y + threadIdx.y - 1;
	if (i > nx + 1 || j > ny + 1) return;
	int c = cidx(i, j, nrow);
	if (j < 0 || j >= ny) { dst[c] = src[c]; return; }
	int iw = i;
	if (i < 0 || i >= nx) {
		if (moving_window) { dst[c] = src[c]; return; }
		iw = (i < 0) ? i + nx : i - nx;
	}
	int cw = cidx(iw, j, nrow);
	dst[c] = stencil3(src[cw - 1], src[cw], src[cw + 1], sa, sb);
}

// One pass along y (reference kernel_y, current.c:366-413): columns 0..nx-1 filtered,
// then y guards of ALL columns copied from the (new) interior rows.
__global__ void k_smooth_y(f3* __restrict__ dst, const f3* __restrict__ src, int nx, int ny, int nrow,
                           float sa, float sb) {
	int i = blockIdx.x * blockDim.x + threadIdx.x - 1;
	int j = blockIdx.y * blockDim.y + threadIdx.y - 1;
	if (i > nx + 1 || j > ny + 1) return;
	int jw = (j < 0) ? j + ny : ((j >= ny) ? j - ny : j);
	int cw = cidx(i, jw, nrow);
	int c = cidx(i, j, nrow);
	if (i < 0 || i >= nx) { dst[c] = src[cw]; return; }
	dst[c] = stencil3(src[cw - nrow], src[cw], src[cw + nrow], sa, sb);
}

// ALL the [sa,sb,sa] passes along x in one kernel (reference current_smooth's x loop, current.c:427-447, each pass =
// kernel_x, :316-354).  A WARP owns 32 consecutive cells of one row, one per lane, and runs the passes in registers:
// the neighbours come from the lanes next door (two shuffles per component and pass), so every pass the valid
// stretch shrinks by one lane per side and V = 32 - 2 passes cells in the middle come out; consecutive warps
// overlap by 2 passes cells.  Every cell goes through exactly the stencils, on exactly the values, of the
// pass-by-pass form (the overlap recomputes what the neighbouring warp computes), so the result is bit-identical;
// the grid is read and written once instead of once per pass and nothing waits at a barrier.  The first warp of a
// row starts so that guard -1 is its first valid lane; the guards nx, nx+1 fall to the last one.  What lies beyond
// the row's ends:
//   periodic box        the cells of the other end (the reference refreshes its x guards from there after every
//                       pass, so guards and their images evolve alike);
//   moving window       nothing - the guards keep their raw values through all passes (current.c:346: no refresh)
//                       and the cells next to them see those (`frozen` cells);
//   neighbour slab      the neighbour's cells, H = passes + 2 columns sent ONCE before the kernel (zdev_slab.cuh)
//                       instead of a guard refresh after every pass; the warps at the row's ends wait for the
//                       message and read the payload.
// Rows outside 0..ny-1 are not filtered (kernel_x only walks the interior rows): copied through.
#define SMX_MAXP 12
#define SMX_WARPS 8
struct smx_coef { float sa[SMX_MAXP], sb[SMX_MAXP]; };
__global__ void __launch_bounds__(SMX_WARPS * 32)
k_smooth_x_multi(f3* __restrict__ dst, const f3* __restrict__ src, int nx, int ny, int nrow, int npass, smx_coef cf,
                 int frozen_lo, int frozen_hi, slab_msg L, slab_msg R, unsigned seq_l, unsigned seq_r) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int j = (int) blockIdx.y - 1;
	const int V = 32 - 2 * npass;                                    // valid cells per warp
	const int kw = blockIdx.x * SMX_WARPS + warp;                    // which stretch of the row
	const int i = -1 - npass + kw * V + lane;                        // this lane's cell (extended index)
	if (-1 + kw * V > nx + 1) return;                                // past the row's end
	const bool mine = lane >= npass && lane < 32 - npass && i <= nx + 1;
	if (j < 0 || j >= ny) {                                          // pass-through rows
		if (mine) { const int c = cidx(i, j, nrow); dst[c] = src[c]; }
		return;
	}
	const int H = npass + 2;
	f3 v;
	if (L.buf && -1 - npass + kw * V < 0) {                          // (warp-uniform) some lanes read the left halo
		if (lane == 0) slab_wait_lane(L.flag, seq_l);
		__syncwarp();
	}
	if (R.buf && -1 - npass + kw * V + 31 >= nx) {
		if (lane == 0) slab_wait_lane(R.flag, seq_r);
		__syncwarp();
	}
	if (i < 0 && L.buf) {
		const float* q = reinterpret_cast<const float*>(L.buf + (size_t) j * H + (i + H));      // column i of the left halo
		v.x = __ldcg(q); v.y = __ldcg(q + 1); v.z = __ldcg(q + 2);
	} else if (i >= nx && R.buf) {
		const float* q = reinterpret_cast<const float*>(R.buf + (size_t) j * H + min(i - nx, H - 1));
		v.x = __ldcg(q); v.y = __ldcg(q + 1); v.z = __ldcg(q + 2);
	} else {
		int iw = i;
		if (i < 0) iw = frozen_lo ? max(i, -1) : ((i % nx) + nx) % nx;
		else if (i >= nx) iw = frozen_hi ? min(i, nx + 1) : i % nx;
		v = src[cidx(iw, j, nrow)];
	}
	const bool keep = (frozen_lo && i < 0) || (frozen_hi && i >= nx);
	#pragma unroll
	for (int p = 0; p < SMX_MAXP; p++) {                             // (unrolled: the coefficients stay kernel parameters)
		if (p >= npass) break;
		const float sa = cf.sa[p], sb = cf.sb[p];
		f3 lo, hi;
		lo.x = __shfl_up_sync(0xffffffffu, v.x, 1); hi.x = __shfl_down_sync(0xffffffffu, v.x, 1);
		lo.y = __shfl_up_sync(0xffffffffu, v.y, 1); hi.y = __shfl_down_sync(0xffffffffu, v.y, 1);
		lo.z = __shfl_up_sync(0xffffffffu, v.z, 1); hi.z = __shfl_down_sync(0xffffffffu, v.z, 1);
		const f3 fs = stencil3(lo, v, hi, sa, sb);                   // current.c:334-336, left to right
		if (!keep) v = fs;                                           // (lanes 0 and 31 see themselves: outside the valid stretch)
	}
	if (mine) dst[cidx(i, j, nrow)] = v;
}

// reference get_smooth_comp, em2d/current.c:297-304 (double -> float on b)
static void smooth_comp(int n, float* sa, float* sb) {
	float a = -1;
	float b = (float) ((4.0 + 2.0 * n) / n);
	float total = 2 * a + b;
	*sa = a / total; *sb = b / total;
}

static void smooth_pass(zdev_grid2d* g, int dir, float sa, float sb, int moving_window) {
	need_J(g); need_tmp(g);
	dim3 blk(64, 4), grd(zdev_div_up(g->nx + 3, 64), zdev_div_up(g->ny + 3, 4));
	if (dir == 0) ZDEV_LAUNCH(k_smooth_x, grd, blk, 0, g->tmp, g->J, g->nx, g->ny, g->nrow, sa, sb, moving_window);
	else          ZDEV_LAUNCH(k_smooth_y, grd, blk, 0, g->tmp, g->J, g->nx, g->ny, g->nrow, sa, sb);
	f3* t = g->J; g->J = g->tmp; g->tmp = t;
}

static int smooth_fused = -1;             // ZPIC_FUSED_SMOOTH=0: one kernel per pass (same results bit for bit)
static bool use_fused_smooth(int npass) {
	if (smooth_fused < 0) { const char* e = getenv("ZPIC_FUSED_SMOOTH"); smooth_fused = e ? atoi(e) != 0 : 1; }
	return smooth_fused && npass >= 2 && npass <= SMX_MAXP;
}
// all x passes of the plan [first, first + npass) in one kernel; L / R: wide halos of neighbour slabs (or none)
static void smooth_x_fused(zdev_grid2d* g, const float* sa, const float* sb, int npass, int frozen_lo, int frozen_hi,
                           slab_msg L, slab_msg R, unsigned seq_l, unsigned seq_r) {
	need_J(g); need_tmp(g);
	smx_coef cf;
	for (int p = 0; p < npass; p++) { cf.sa[p] = sa[p]; cf.sb[p] = sb[p]; }
	const int V = 32 - 2 * npass;                                   // cells a warp delivers
	dim3 grd(zdev_div_up(zdev_div_up(g->nx + 3, V), SMX_WARPS), g->ny + 3);
	ZDEV_LAUNCH(k_smooth_x_multi, grd, SMX_WARPS * 32, 0, g->tmp, g->J, g->nx, g->ny, g->nrow, npass, cf, frozen_lo, frozen_hi, L, R, seq_l, seq_r);
	f3* t = g->J; g->J = g->tmp; g->tmp = t;
}

extern "C" void zdev_current_smooth(zdev_grid2d* g, int moving_window, int xtype, int ytype, int xlevel, int ylevel) {
	float sa, sb;
	if (xtype != 0) {
		const int npx = xlevel + (xtype == 2 ? 1 : 0);
		if (use_fused_smooth(npx)) {
			float ca[SMX_MAXP], cb[SMX_MAXP];
			for (int i = 0; i < xlevel; i++) { ca[i] = 0.25f; cb[i] = 0.5f; }
			if (xtype == 2) smooth_comp(xlevel, &ca[xlevel], &cb[xlevel]);
			const slab_msg none = { nullptr, nullptr, 0, 0 };
			smooth_x_fused(g, ca, cb, npx, moving_window, moving_window, none, none, 0, 0);
		} else {
			for (int i = 0; i < xlevel; i++) smooth_pass(g, 0, 0.25f, 0.5f, moving_window);
			if (xtype == 2) { smooth_comp(xlevel, &sa, &sb); smooth_pass(g, 0, sa, sb, moving_window); }
		}
	}
	if (ytype != 0) {
		// sic: the reference counts the y passes with xlevel (em2d/current.c:449)
		for (int i = 0; i < xlevel; i++) smooth_pass(g, 1, 0.25f, 0.5f, moving_window);
		if (ytype == 2) { smooth_comp(ylevel, &sa, &sb); smooth_pass(g, 1, sa, sb, moving_window); }
	}
}

// the pass list of current_smooth (reference em2d/current.c:427-459) for callers that have to put a
// halo exchange between passes: dirs[k] = 0 (x) / 1 (y), coefficients [sa, sb, sa]; returns the count
extern "C" int zdev_smooth_plan(int xtype, int ytype, int xlevel, int ylevel, int* dirs, float* sa, float* sb) {
	int n = 0;
	if (xtype != 0) {
		for (int i = 0; i < xlevel; i++) { dirs[n] = 0; sa[n] = 0.25f; sb[n] = 0.5f; n++; }
		if (xtype == 2) { dirs[n] = 0; smooth_comp(xlevel, &sa[n], &sb[n]); n++; }
	}
	if (ytype != 0) {
		for (int i = 0; i < xlevel; i++) { dirs[n] = 1; sa[n] = 0.25f; sb[n] = 0.5f; n++; }
		if (ytype == 2) { dirs[n] = 1; smooth_comp(ylevel, &sa[n], &sb[n]); n++; }
	}
	return n;
}
extern "C" void zdev_smooth_pass(zdev_grid2d* g, int dir, float sa, float sb, int keep_x_guards) {
	smooth_pass(g, dir, sa, sb, keep_x_guards);
}
extern "C" void zdev_current_fold_y(zdev_grid2d* g) {
	need_J(g);
	ZDEV_LAUNCH(k_fold_y, zdev_div_up(g->nrow, 128), 128, 0, g->J, g->ny, g->nrow);
}

// columns [i0, i0+ncols) x rows [j0, j0+nrows) of a grid <-> dense device buffer [row][col] of float3
__global__ void k_pack_cols(const f3* __restrict__ G, int nrow, int i0, int ncols, int j0, int nrows, f3* __restrict__ out) {
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= ncols * nrows) return;
	int r = k / ncols, c = k - r * ncols;
	out[k] = G[cidx(i0 + c, j0 + r, nrow)];
}
__global__ void k_unpack_cols(f3* __restrict__ G, int nrow, int i0, int ncols, int j0, int nrows, const f3* __restrict__ in, int add) {
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= ncols * nrows) return;
	int r = k / ncols, c = k - r * ncols;
	f3 v = in[k];
	f3* d = &G[cidx(i0 + c, j0 + r, nrow)];
	if (add) { f3 o = *d; o.x += v.x; o.y += v.y; o.z += v.z; *d = o; } else *d = v;
}
extern "C" void zdev_grid2d_pack_cols(zdev_grid2d* g, int which, int i0, int ncols, int j0, int nrows, float* dev_out) {
	int n = ncols * nrows;
	ZDEV_LAUNCH(k_pack_cols, zdev_div_up(n, 256), 256, 0, grid_sel(g, which), g->nrow, i0, ncols, j0, nrows, (f3*) dev_out);
}
extern "C" void zdev_grid2d_unpack_cols(zdev_grid2d* g, int which, int i0, int ncols, int j0, int nrows, const float* dev_in, int add) {
	int n = ncols * nrows;
	ZDEV_LAUNCH(k_unpack_cols, zdev_div_up(n, 256), 256, 0, grid_sel(g, which), g->nrow, i0, ncols, j0, nrows, (const f3*) dev_in, add);
}

extern "C" void zdev_current_update(zdev_grid2d* g, int moving_window, int xtype, int ytype, int xlevel, int ylevel) {
	if (!g->slab) {
		zdev_current_update_gc(g, moving_window);
		zdev_current_smooth(g, moving_window, xtype, ytype, xlevel, ylevel);
		return;
	}
	// One slab of a wider box.  Guard fold: the three aliased columns (nx-1, nx, nx+1) <-> (-1, 0, 1)
	// of two neighbours are swapped and ADDED on both sides (addition commutes: both end up with the sums the
	// reference's fold-and-copy-back leaves, em2d/current.c:128-137); then the local y fold.
	need_J(g);
	slab_exchange(g, g->J, nullptr, -1, 3, g->nx - 1, 3, -1, 3, g->nx - 1, 3, -1, g->nrows, 1, true, true);
	ZDEV_LAUNCH(k_fold_y, zdev_div_up(g->nrow, 128), 128, 0, g->J, g->ny, g->nrow);
	// Smoothing: a pass leaves the x guards alone (as under a moving window); guards that mirror a neighbour's
	// cells are refreshed from it after every x pass (em2d/current.c:346-352).
	int dirs[64]; float sa[64], sb[64];
	const int np = zdev_smooth_plan(xtype, ytype, xlevel, ylevel, dirs, sa, sb);
	bool y_passes = false;
	int k0 = 0;
	int npx = 0;
	while (npx < np && dirs[npx] == 0) npx++;                // the x passes come first in the plan
	if (use_fused_smooth(npx) && npx + 2 <= SMOOTH_HALO_MAX && g->nx >= npx + 2) {
		// ONE exchange of H = passes + 2 columns per side (my lowest H columns to the left neighbour, my highest H
		// to the right one), then all x passes in one kernel whose edge blocks read the neighbours' columns from
		// the mailbox; a side without a neighbour (the open ends of a moving-window chain) keeps its guards frozen
		zdev_link& K = g->link;
		const int H = npx + 2;
		const bool do_l = K.left >= 0, do_r = K.right >= 0;
		slab_msg SL = { nullptr, nullptr, 0, 0 }, SR = SL, RL = SL, RR = SL;
		unsigned seq_l = 0, seq_r = 0;
		if (do_l) {
			seq_l = ++K.seq[0];
			SL = { (f3*) zdev_link_out(K, 0, seq_l), &zdev_link_out_hdr(K, 0)->flag[1], 0, H };
			RL = { (f3*) zdev_link_in(K, 0, seq_l), &zdev_link_in_hdr(K)->flag[0], -H, H };
		}
		if (do_r) {
			seq_r = ++K.seq[1];
			SR = { (f3*) zdev_link_out(K, 1, seq_r), &zdev_link_out_hdr(K, 1)->flag[0], g->nx - H, H };
			RR = { (f3*) zdev_link_in(K, 1, seq_r), &zdev_link_in_hdr(K)->flag[1], g->nx, H };
		}
		if (do_l || do_r) {
			dim3 grd(std::max(1, std::min(zdev_div_up(H * g->ny, 256), 32)), 2);
			ZDEV_LAUNCH(k_slab_send, grd, 256, 0, g->J, (const f3*) nullptr, 1, g->nrow, 0, g->ny, SL, SR, seq_l, seq_r, K.ticket);
		}
		smooth_x_fused(g, sa, sb, npx, !do_l, !do_r, RL, RR, seq_l, seq_r);
		k0 = npx;
	}
	for (int k = k0; k < np; k++) {
		smooth_pass(g, dirs[k], sa[k], sb[k], 1);
		if (dirs[k] == 0) slab_halo_refresh(g, g->J, nullptr, 0, g->ny, false);
		else y_passes = true;
	}
	// kernel_y leaves the x guard columns un-filtered (em2d/current.c:382-411).  Across the box boundary that is
	// what the reference feeds to yee_e, so the wrap-around edge keeps it; guards that mirror interior cells of a
	// neighbour slab must see the filtered values.
	if (y_passes) slab_halo_refresh(g, g->J, nullptr, -1, g->nrows, true);
}

// ------------------------------------------------------------------ moving window

// new[i] = old[i+1] for i in [-1,nx-2]; columns nx-1..nx+1 zeroed (reference emf.c:659-670)
// zero_right = 0 (slab whose right edge is not the box edge): columns nx-1, nx come from the halo,
// column nx+1 keeps its old value until the caller refreshes the halo
__global__ void k_shift_left(f3* __restrict__ dst, const f3* __restrict__ src, int nx, int nrows, int nrow, int zero_right) {
	int c = blockIdx.x * blockDim.x + threadIdx.x;   // buffer column
	int r = blockIdx.y * blockDim.y + threadIdx.y;   // buffer row
	if (c >= nrow || r >= nrows) return;
	size_t k = (size_t) r * nrow + c;
	f3 z = {0.f, 0.f, 0.f};
	if (zero_right) dst[k] = (c < nx) ? src[k + 1] : z;            // buffer column c = cell c-1
	else            dst[k] = (c < nrow - 1) ? src[k + 1] : src[k];
}

extern "C" void zdev_emf_shift(zdev_grid2d* g, int zero_right) {
	need_EB(g); need_tmp(g);
	dim3 blk(64, 4), grd(zdev_div_up(g->nrow, 64), zdev_div_up(g->nrows, 4));
	int alias_e = (g->Epart == g->E), alias_b = (g->Bpart == g->B);
	ZDEV_LAUNCH(k_shift_left, grd, blk, 0, g->tmp, g->E, g->nx, g->nrows, g->nrow, zero_right);
	{ f3* t = g->E; g->E = g->tmp; g->tmp = t; }
	ZDEV_LAUNCH(k_shift_left, grd, blk, 0, g->tmp, g->B, g->nx, g->nrows, g->nrow, zero_right);
	{ f3* t = g->B; g->B = g->tmp; g->tmp = t; }
	if (alias_e) g->Epart = g->E;
	if (alias_b) g->Bpart = g->B;
}

extern "C" void zdev_emf_move_window(zdev_grid2d* g) { zdev_emf_shift(g, 1); }

// ------------------------------------------------------------------ external fields

__global__ void k_add_uniform(f3* __restrict__ dst, const f3* __restrict__ src, size_t n, f3 v) {
	size_t k = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	f3 e = src[k]; e.x += v.x; e.y += v.y; e.z += v.z; dst[k] = e;
}
__global__ void k_add_grid(f3* __restrict__ dst, const f3* __restrict__ src, const f3* __restrict__ ext, size_t n) {
	size_t k = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	f3 e = src[k], v = ext[k]; e.x += v.x; e.y += v.y; e.z += v.z; dst[k] = e;
}

// reference emf_update_part_fld, em2d/emf.c:838-914
static void update_part_fld(zdev_grid2d* g) {
	int grd = zdev_div_up(g->ncell, 256);
	if (g->e_ext == 1) ZDEV_LAUNCH(k_add_uniform, grd, 256, 0, g->Epart, g->E, g->ncell, g->e0);
	else if (g->e_ext == 2) ZDEV_LAUNCH(k_add_grid, grd, 256, 0, g->Epart, g->E, g->Eext, g->ncell);
	if (g->b_ext == 1) ZDEV_LAUNCH(k_add_uniform, grd, 256, 0, g->Bpart, g->B, g->ncell, g->b0);
	else if (g->b_ext == 2) ZDEV_LAUNCH(k_add_grid, grd, 256, 0, g->Bpart, g->B, g->Bext, g->ncell);
}

static void set_ext(zdev_grid2d* g, int is_b, int mode, const float v[3], const float* host_grid) {
	need_EB(g);
	f3** part = is_b ? &g->Bpart : &g->Epart;
	f3*  self = is_b ? g->B : g->E;
	int* flag = is_b ? &g->b_ext : &g->e_ext;
	f3** ext  = is_b ? &g->Bext : &g->Eext;
	if (*flag && !mode) { cudaFree(*part); *part = self; }
	if (!*flag && mode) ZDEV_CHECK(cudaMalloc(part, g->ncell * sizeof(f3)));
	if (!mode) *part = self;
	*flag = mode;
	if (mode == 1) { f3 t = {v[0], v[1], v[2]}; if (is_b) g->b0 = t; else g->e0 = t; }
	if (mode == 2) {
		if (!*ext) ZDEV_CHECK(cudaMalloc(ext, g->ncell * sizeof(f3)));
		ZDEV_CHECK(cudaMemcpyAsync(*ext, host_grid, g->ncell * sizeof(f3), cudaMemcpyHostToDevice, zdev_strm));
		ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	}
}

extern "C" void zdev_emf_update_part_fld(zdev_grid2d* g) { need_EB(g); update_part_fld(g); }

extern "C" void zdev_emf_set_ext_uniform(zdev_grid2d* g, int e_on, const float e0[3], int b_on, const float b0[3]) {
	set_ext(g, 0, e_on ? 1 : 0, e0, nullptr);
	set_ext(g, 1, b_on ? 1 : 0, b0, nullptr);
	update_part_fld(g);
}
extern "C" void zdev_emf_set_ext_grid(zdev_grid2d* g, const float* he, const float* hb) {
	if (he) set_ext(g, 0, 2, nullptr, he);
	if (hb) set_ext(g, 1, 2, nullptr, hb);
	update_part_fld(g);
}

// ------------------------------------------------------------------ field advance

extern "C" void zdev_emf_advance(zdev_grid2d* g, zdev_grid2d* gj, float dt, float dx, float dy, int moving_window, int shift_window) {
	// reference em2d/emf.c:694-698, scalars formed exactly as in yee_b/yee_e (:509-510, :537-538)
	if (fused_yee_on()) {
		yee_fused(g, gj, dt, dx, dy);
	} else {
		float dtb = dt / 2.0f;
		zdev_yee_b(g, dtb / dx, dtb / dy);
		zdev_yee_e(g, gj, dt / dx, dt / dy, dt);
		zdev_yee_b(g, dtb / dx, dtb / dy);
	}
	if (g->slab) {
		// guards towards a neighbour slab come from its interior (every row), then the local periodic y copies
		slab_halo_refresh(g, g->E, g->B, -1, g->nrows, false);
		zdev_emf_update_gc(g, 1);
		update_part_fld(g);
		if (shift_window) {
			zdev_emf_shift(g, g->link.right < 0);     // interior right edges shift without zeroing ...
			slab_halo_refresh(g, g->E, g->B, -1, g->nrows, false);       // ... and take the new columns from the neighbour
		}
		return;
	}
	zdev_emf_update_gc(g, moving_window);
	update_part_fld(g);
	if (shift_window) {
		zdev_emf_move_window(g);
		// E_part/B_part buffers are refreshed by the next emf_advance, as in the reference
	}
}

// ------------------------------------------------------------------ energy

__global__ void k_energy(const f3* __restrict__ E, const f3* __restrict__ B, int nx, int ny, int nrow, double* out) {
	double s[6] = {0, 0, 0, 0, 0, 0};
	long n = (long) nx * ny;
	for (long k = (long) blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long) gridDim.x * blockDim.x) {
		int j = (int) (k / nx), i = (int) (k - (long) j * nx);
		f3 e = E[cidx(i, j, nrow)], b = B[cidx(i, j, nrow)];
		// float products widened to double, as in reference em2d/emf.c:739-744
		s[0] += e.x * e.x; s[1] += e.y * e.y; s[2] += e.z * e.z;
		s[3] += b.x * b.x; s[4] += b.y * b.y; s[5] += b.z * b.z;
	}
	__shared__ double red[6][8];
	int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	#pragma unroll
	for (int q = 0; q < 6; q++) {
		double v = s[q];
		for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
		if (lane == 0) red[q][w] = v;
	}
	__syncthreads();
	if (threadIdx.x < 6) {
		double v = 0;
		for (int k = 0; k < (int) (blockDim.x >> 5); k++) v += red[threadIdx.x][k];
		atomicAdd(&out[threadIdx.x], v);
	}
}

extern "C" void zdev_emf_energy(zdev_grid2d* g, double sums[6]) {
	need_EB(g);
	ZDEV_CHECK(cudaMemsetAsync(g->d_sums, 0, 6 * sizeof(double), zdev_strm));
	int grd = zdev_div_up((long) g->nx * g->ny, 256);
	if (grd > 4 * zdev_num_sm) grd = 4 * zdev_num_sm;
	ZDEV_LAUNCH(k_energy, grd, 256, 0, g->E, g->B, g->nx, g->ny, g->nrow, g->d_sums);
	ZDEV_CHECK(cudaMemcpyAsync(sums, g->d_sums, 6 * sizeof(double), cudaMemcpyDeviceToHost, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
}

// accessors used by the particle kernels (zdev_spec2d.cu)
f3* zdev_grid2d_Epart(zdev_grid2d* g) { need_EB(g); return g->Epart; }
f3* zdev_grid2d_Bpart(zdev_grid2d* g) { need_EB(g); return g->Bpart; }
f3* zdev_grid2d_J(zdev_grid2d* g) { need_J(g); return g->J; }
int zdev_grid2d_nx(zdev_grid2d* g) { return g->nx; }
int zdev_grid2d_ny(zdev_grid2d* g) { return g->ny; }
