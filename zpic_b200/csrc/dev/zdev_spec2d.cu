// zpic-b200 :: em2d particle species on the device.
//
// Data layout in HBM.  The grid is cut into tiles of TX x TY cells; every tile owns
// a fixed-capacity segment [tile_off[t], tile_off[t+1]) of seven SoA arrays
// (ix, iy, x, y, ux, uy, uz; optional injection tag) of which the first tile_np[t]
// slots are live.  One CTA advances one tile: it stages the (TX+2)x(TY+2) E/B
// neighbourhood in shared memory once, streams the tile's particles through
// registers with fully coalesced SoA loads, and writes survivors back compacted in
// place.  Particles that leave the tile go through a small global "migrants" list
// that a second kernel appends to the destination tiles; per step a particle is
// therefore read once and written once (56 B), plus the few percent that migrate.
//
// Replaces reference em2d/particles.c:1104-1269 (spec_advance incl. boundaries and
// spec_move_window's index shift) and :942-1007 (spec_sort: binning by tile is kept
// current every step instead of a counting sort every n_sort steps).
#include "zdev_common.cuh"
#include "pic2d_core.cuh"
#include <vector>
#include <cstring>

// accessors implemented in zdev_grid2d.cu
f3* zdev_grid2d_Epart(zdev_grid2d* g);
f3* zdev_grid2d_Bpart(zdev_grid2d* g);
f3* zdev_grid2d_J(zdev_grid2d* g);
int zdev_grid2d_nx(zdev_grid2d* g);
int zdev_grid2d_ny(zdev_grid2d* g);

// host AoS record (include/em2d/particles.h t_part)
struct part_aos { int ix, iy; float x, y, ux, uy, uz; };

// SoA view handed to kernels by value
struct soa2d {
	int *ix, *iy;
	float *x, *y, *ux, *uy, *uz;
	int *tag;            // null unless ids are tracked
};

// control block in device memory, zeroed at the start of every advance
struct ctl2d {
	double energy;                   // sum of utsq/(gamma+1)
	unsigned long long np;           // live particles after the step
	unsigned int n_mig;              // entries in the migrants list
	unsigned int flags;              // 1: tile capacity overflow, 2: migrants list overflow
};

struct zdev_spec2d {
	int nx, ny;
	int TX, TY, ntx, nty, ntiles;
	int ppc_hint, track_ids;
	double slack;
	int64_t cap_total;               // total SoA slots
	soa2d p;                         // tile-binned particles
	int64_t* tile_off;               // device, ntiles+1
	int* tile_np;                    // device, ntiles
	soa2d mig;                       // migrants list (global cell indices)
	unsigned int mig_cap;
	ctl2d* ctl;                      // device
	int64_t np_host;                 // last known particle count
	int ids_valid;                   // tags are a permutation of [0,np)
	std::vector<int64_t>* h_off;     // host copy of tile_off
};

static const int PUSH_THREADS = 256;
static const int MAX_TILE = 16;      // cells per tile edge (shared-memory E/B tile is (MAX_TILE+2)^2)

static void soa_alloc(soa2d& a, int64_t n, int with_tag) {
	size_t nn = (size_t) (n > 0 ? n : 1);
	ZDEV_CHECK(cudaMalloc(&a.ix, nn * 4)); ZDEV_CHECK(cudaMalloc(&a.iy, nn * 4));
	ZDEV_CHECK(cudaMalloc(&a.x, nn * 4));  ZDEV_CHECK(cudaMalloc(&a.y, nn * 4));
	ZDEV_CHECK(cudaMalloc(&a.ux, nn * 4)); ZDEV_CHECK(cudaMalloc(&a.uy, nn * 4));
	ZDEV_CHECK(cudaMalloc(&a.uz, nn * 4));
	a.tag = nullptr;
	if (with_tag) ZDEV_CHECK(cudaMalloc(&a.tag, nn * 4));
}
static void soa_free(soa2d& a) {
	cudaFree(a.ix); cudaFree(a.iy); cudaFree(a.x); cudaFree(a.y);
	cudaFree(a.ux); cudaFree(a.uy); cudaFree(a.uz); cudaFree(a.tag);
	memset(&a, 0, sizeof(a));
}

extern "C" zdev_spec2d* zdev_spec2d_create(int nx, int ny, int ppc_hint, int track_ids) {
	zdev_require_init();
	zdev_spec2d* s = new zdev_spec2d();
	memset(s, 0, sizeof(*s));
	s->nx = nx; s->ny = ny;
	s->ppc_hint = ppc_hint > 0 ? ppc_hint : 1;
	s->track_ids = track_ids;
	// tile size: aim at ~4096 particles per tile (amortises the E/B staging and the J
	// traffic, keeps migration to a few percent); overridable for experiments
	int cells = 4096 / s->ppc_hint;
	int tx = MAX_TILE, ty = MAX_TILE;
	while (tx * ty > cells && tx * ty > 16) { if (ty >= tx) ty >>= 1; else tx >>= 1; }
	if (const char* e = getenv("ZPIC_TILE_X")) tx = atoi(e);
	if (const char* e = getenv("ZPIC_TILE_Y")) ty = atoi(e);
	if (tx < 1) tx = 1; if (ty < 1) ty = 1;
	if (tx > MAX_TILE) tx = MAX_TILE; if (ty > MAX_TILE) ty = MAX_TILE;
	if (tx > nx) tx = nx; if (ty > ny) ty = ny;
	s->TX = tx; s->TY = ty;
	s->ntx = (nx + tx - 1) / tx; s->nty = (ny + ty - 1) / ty;
	s->ntiles = s->ntx * s->nty;
	s->slack = 0.0;
	if (const char* e = getenv("ZPIC_TILE_SLACK")) s->slack = atof(e);
	ZDEV_CHECK(cudaMalloc(&s->tile_off, (size_t) (s->ntiles + 1) * sizeof(int64_t)));
	ZDEV_CHECK(cudaMalloc(&s->tile_np, (size_t) s->ntiles * sizeof(int)));
	ZDEV_CHECK(cudaMemsetAsync(s->tile_np, 0, (size_t) s->ntiles * sizeof(int), zdev_strm));
	ZDEV_CHECK(cudaMalloc(&s->ctl, sizeof(ctl2d)));
	ZDEV_CHECK(cudaMemsetAsync(s->ctl, 0, sizeof(ctl2d), zdev_strm));
	s->h_off = new std::vector<int64_t>();
	return s;
}

static void spec_free_particles(zdev_spec2d* s) {
	if (s->cap_total) { soa_free(s->p); soa_free(s->mig); }
	s->cap_total = 0; s->mig_cap = 0;
}

extern "C" void zdev_spec2d_destroy(zdev_spec2d* s) {
	if (!s) return;
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	spec_free_particles(s);
	cudaFree(s->tile_off); cudaFree(s->tile_np); cudaFree(s->ctl);
	delete s->h_off;
	delete s;
}

extern "C" void zdev_spec2d_tile_info(zdev_spec2d* s, int* tx, int* ty, int* ntiles, int64_t* capacity) {
	if (tx) *tx = s->TX; if (ty) *ty = s->TY; if (ntiles) *ntiles = s->ntiles; if (capacity) *capacity = s->cap_total;
}

// Lay out tile segments for the given per-tile populations and (re)allocate the SoA
// arrays.  Capacity per tile = slack * max(population, nominal fill) rounded to 32.
static void spec_layout(zdev_spec2d* s, const std::vector<int>& cnt, int64_t np) {
	double slack = s->slack;
	if (slack <= 0.0) slack = (np > (int64_t) 200000000) ? 1.25 : 2.0;
	std::vector<int64_t>& off = *s->h_off;
	off.assign(s->ntiles + 1, 0);
	for (int ty = 0; ty < s->nty; ty++) for (int tx = 0; tx < s->ntx; tx++) {
		int t = tx + ty * s->ntx;
		int cx = (tx + 1) * s->TX <= s->nx ? s->TX : s->nx - tx * s->TX;
		int cy = (ty + 1) * s->TY <= s->ny ? s->TY : s->ny - ty * s->TY;
		int64_t nominal = (int64_t) cx * cy * s->ppc_hint;
		int64_t want = cnt[t] > nominal ? cnt[t] : nominal;
		int64_t cap = (int64_t) (want * slack) + 64;
		cap = (cap + 31) & ~(int64_t) 31;
		off[t + 1] = off[t] + cap;
	}
	int64_t total = off[s->ntiles];
	spec_free_particles(s);
	soa_alloc(s->p, total, s->track_ids);
	s->cap_total = total;
	int64_t mc = total / 4 + 65536;
	if (mc > 0x7fffffff) mc = 0x7fffffff;
	s->mig_cap = (unsigned int) mc;
	soa_alloc(s->mig, s->mig_cap, s->track_ids);
	ZDEV_CHECK(cudaMemcpyAsync(s->tile_off, off.data(), (size_t) (s->ntiles + 1) * sizeof(int64_t),
	                           cudaMemcpyHostToDevice, zdev_strm));
	ZDEV_CHECK(cudaMemsetAsync(s->tile_np, 0, (size_t) s->ntiles * sizeof(int), zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
}

// ------------------------------------------------------------------ host <-> device

__global__ void k_count_tiles(const part_aos* __restrict__ a, int64_t np, int TX, int TY, int ntx, int* cnt) {
	int64_t k = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= np) return;
	atomicAdd(&cnt[a[k].ix / TX + (a[k].iy / TY) * ntx], 1);
}

// append AoS records to their tiles; tag = tag0 + index when ids are tracked
__global__ void k_scatter_tiles(const part_aos* __restrict__ a, int64_t np, int TX, int TY, int ntx,
                                soa2d p, const int64_t* __restrict__ off, int* tile_np, ctl2d* ctl, int tag0) {
	int64_t k = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= np) return;
	part_aos r = a[k];
	int t = r.ix / TX + (r.iy / TY) * ntx;
	int slot = atomicAdd(&tile_np[t], 1);
	int64_t d = off[t] + slot;
	if (d >= off[t + 1]) { atomicOr(&ctl->flags, 1u); return; }
	p.ix[d] = r.ix; p.iy[d] = r.iy; p.x[d] = r.x; p.y[d] = r.y; p.ux[d] = r.ux; p.uy[d] = r.uy; p.uz[d] = r.uz;
	if (p.tag) p.tag[d] = tag0 + (int) k;
}

static void check_flags(zdev_spec2d* s, unsigned int flags) {
	if (flags & 1u) {
		fprintf(stderr, "(*error*) zpic-b200: particle tile capacity exceeded (tile %dx%d cells); "
		        "raise ZPIC_TILE_SLACK (current %.2f) and rerun, aborting.\n", s->TX, s->TY, s->slack);
		exit(-1);
	}
	if (flags & 2u) {
		fprintf(stderr, "(*error*) zpic-b200: particle migration list overflow (capacity %u), aborting.\n", s->mig_cap);
		exit(-1);
	}
}

static void spec_append_dev(zdev_spec2d* s, const part_aos* d_aos, int64_t np, int tag0) {
	if (np <= 0) return;
	ZDEV_LAUNCH(k_scatter_tiles, zdev_div_up(np, 256), 256, 0, d_aos, np, s->TX, s->TY, s->ntx,
	            s->p, s->tile_off, s->tile_np, s->ctl, tag0);
}

extern "C" void zdev_spec2d_upload(zdev_spec2d* s, const void* part, int64_t np) {
	part_aos* d_aos = nullptr;
	std::vector<int> cnt(s->ntiles, 0);
	if (np > 0) {
		ZDEV_CHECK(cudaMalloc(&d_aos, (size_t) np * sizeof(part_aos)));
		ZDEV_CHECK(cudaMemcpyAsync(d_aos, part, (size_t) np * sizeof(part_aos), cudaMemcpyHostToDevice, zdev_strm));
		int* d_cnt; ZDEV_CHECK(cudaMalloc(&d_cnt, (size_t) s->ntiles * sizeof(int)));
		ZDEV_CHECK(cudaMemsetAsync(d_cnt, 0, (size_t) s->ntiles * sizeof(int), zdev_strm));
		ZDEV_LAUNCH(k_count_tiles, zdev_div_up(np, 256), 256, 0, d_aos, np, s->TX, s->TY, s->ntx, d_cnt);
		ZDEV_CHECK(cudaMemcpyAsync(cnt.data(), d_cnt, (size_t) s->ntiles * sizeof(int), cudaMemcpyDeviceToHost, zdev_strm));
		ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
		cudaFree(d_cnt);
	}
	spec_layout(s, cnt, np);
	ZDEV_CHECK(cudaMemsetAsync(s->ctl, 0, sizeof(ctl2d), zdev_strm));
	spec_append_dev(s, d_aos, np, 0);
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	if (d_aos) cudaFree(d_aos);
	s->np_host = np;
	s->ids_valid = s->track_ids;
}

extern "C" void zdev_spec2d_append(zdev_spec2d* s, const void* part, int64_t np) {
	if (np <= 0) return;
	if (!s->cap_total) { zdev_spec2d_upload(s, part, np); return; }
	part_aos* d_aos;
	ZDEV_CHECK(cudaMalloc(&d_aos, (size_t) np * sizeof(part_aos)));
	ZDEV_CHECK(cudaMemcpyAsync(d_aos, part, (size_t) np * sizeof(part_aos), cudaMemcpyHostToDevice, zdev_strm));
	spec_append_dev(s, d_aos, np, (int) s->np_host);
	ctl2d h;
	ZDEV_CHECK(cudaMemcpyAsync(&h, s->ctl, sizeof(h), cudaMemcpyDeviceToHost, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	cudaFree(d_aos);
	check_flags(s, h.flags);
	s->np_host += np;
}

// tiles -> AoS.  prefix[t] = first output slot of tile t (ignored when scattering by tag)
__global__ void k_gather_aos(soa2d p, const int64_t* __restrict__ off, const int* __restrict__ tile_np,
                             const int64_t* __restrict__ prefix, part_aos* __restrict__ out, int by_tag) {
	int t = blockIdx.x;
	int n = tile_np[t];
	int64_t b = off[t], o = prefix[t];
	for (int k = threadIdx.x; k < n; k += blockDim.x) {
		part_aos r;
		r.ix = p.ix[b + k]; r.iy = p.iy[b + k]; r.x = p.x[b + k]; r.y = p.y[b + k];
		r.ux = p.ux[b + k]; r.uy = p.uy[b + k]; r.uz = p.uz[b + k];
		int64_t d = by_tag ? (int64_t) p.tag[b + k] : o + k;
		out[d] = r;
	}
}

extern "C" int64_t zdev_spec2d_download(zdev_spec2d* s, void* part, int64_t max_np) {
	if (!s->cap_total) return 0;
	std::vector<int> cnt(s->ntiles);
	ZDEV_CHECK(cudaMemcpyAsync(cnt.data(), s->tile_np, (size_t) s->ntiles * sizeof(int), cudaMemcpyDeviceToHost, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	std::vector<int64_t> prefix(s->ntiles);
	int64_t np = 0;
	for (int t = 0; t < s->ntiles; t++) { prefix[t] = np; np += cnt[t]; }
	s->np_host = np;
	if (np == 0) return 0;
	if (np > max_np) {
		fprintf(stderr, "(*error*) zpic-b200: host particle buffer too small (%lld > %lld)\n", (long long) np, (long long) max_np);
		exit(-1);
	}
	int64_t* d_prefix; part_aos* d_aos;
	ZDEV_CHECK(cudaMalloc(&d_prefix, (size_t) s->ntiles * sizeof(int64_t)));
	ZDEV_CHECK(cudaMalloc(&d_aos, (size_t) np * sizeof(part_aos)));
	ZDEV_CHECK(cudaMemcpyAsync(d_prefix, prefix.data(), (size_t) s->ntiles * sizeof(int64_t), cudaMemcpyHostToDevice, zdev_strm));
	ZDEV_LAUNCH(k_gather_aos, s->ntiles, 256, 0, s->p, s->tile_off, s->tile_np, d_prefix, d_aos,
	            (s->track_ids && s->ids_valid) ? 1 : 0);
	ZDEV_CHECK(cudaMemcpyAsync(part, d_aos, (size_t) np * sizeof(part_aos), cudaMemcpyDeviceToHost, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	cudaFree(d_prefix); cudaFree(d_aos);
	return np;
}

extern "C" int64_t zdev_spec2d_np(zdev_spec2d* s) {
	if (!s->cap_total) return 0;
	std::vector<int> cnt(s->ntiles);
	ZDEV_CHECK(cudaMemcpyAsync(cnt.data(), s->tile_np, (size_t) s->ntiles * sizeof(int), cudaMemcpyDeviceToHost, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	int64_t np = 0;
	for (int t = 0; t < s->ntiles; t++) np += cnt[t];
	s->np_host = np;
	return np;
}

// ------------------------------------------------------------------ device-side uniform injection

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
	z += 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}
// standard normal triple for particle `gid` (counter based: no state, reproducible)
__device__ __forceinline__ void normal3(uint64_t seed, uint64_t gid, float& a, float& b, float& c) {
	uint64_t r0 = mix64(seed ^ (gid * 2 + 0) * 0xD1342543DE82EF95ull);
	uint64_t r1 = mix64(seed ^ (gid * 2 + 1) * 0xD1342543DE82EF95ull);
	float u0 = ((uint32_t) (r0 >> 40) + 0.5f) * (1.0f / 16777216.0f);
	float u1 = ((uint32_t) (r0 & 0xffffff) + 0.5f) * (1.0f / 16777216.0f);
	float u2 = ((uint32_t) (r1 >> 40) + 0.5f) * (1.0f / 16777216.0f);
	float u3 = ((uint32_t) (r1 & 0xffffff) + 0.5f) * (1.0f / 16777216.0f);
	float m0 = sqrtf(-2.0f * logf(u0)), m1 = sqrtf(-2.0f * logf(u2));
	float s0, c0, s1, c1;
	sincospif(2.0f * u1, &s0, &c0);
	sincospif(2.0f * u3, &s1, &c1);
	a = m0 * c0; b = m0 * s0; c = m1 * c1;
	(void) s1;
}

// one thread per cell: positions as spec_set_x's UNIFORM branch (particles.c:167-180,
// 335-347), momenta as spec_set_u (thermal, minus the cell mean, plus fluid; :96-142)
__global__ void k_inject_uniform(soa2d p, const int64_t* __restrict__ off, int* tile_np,
                                 int nx, int ny, int TX, int TY, int ntx, int ppcx, int ppcy,
                                 f3 ufl, f3 uth, uint64_t seed) {
	int64_t cell = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (cell >= (int64_t) nx * ny) return;
	int iy = (int) (cell / nx), ix = (int) (cell - (int64_t) iy * nx);
	int tx = ix / TX, ty = iy / TY, t = tx + ty * ntx;
	int cx = (tx + 1) * TX <= nx ? TX : nx - tx * TX;
	int lx = ix - tx * TX, ly = iy - ty * TY;
	int npc = ppcx * ppcy;
	int64_t base = off[t] + (int64_t) (lx + ly * cx) * npc;
	uint64_t gid0 = (uint64_t) cell * npc;
	float sx = 0, sy = 0, sz = 0;
	for (int k = 0; k < npc; k++) {
		float a, b, c; normal3(seed, gid0 + k, a, b, c);
		sx += uth.x * a; sy += uth.y * b; sz += uth.z * c;
	}
	float norm = 1.0f / npc;
	sx *= norm; sy *= norm; sz *= norm;
	float dpcx = 1.0f / ppcx, dpcy = 1.0f / ppcy;
	for (int k = 0; k < npc; k++) {
		float a, b, c; normal3(seed, gid0 + k, a, b, c);
		int kx = k % ppcx, ky = k / ppcx;
		int64_t d = base + k;
		p.ix[d] = ix; p.iy[d] = iy;
		p.x[d] = (float) (dpcx * (kx + 0.5)); p.y[d] = (float) (dpcy * (ky + 0.5));
		p.ux[d] = uth.x * a + (ufl.x - sx);
		p.uy[d] = uth.y * b + (ufl.y - sy);
		p.uz[d] = uth.z * c + (ufl.z - sz);
		if (p.tag) p.tag[d] = (int) (gid0 + k);
	}
	if (lx == 0 && ly == 0) {
		int cy = (ty + 1) * TY <= ny ? TY : ny - ty * TY;
		tile_np[t] = cx * cy * npc;
	}
}

extern "C" void zdev_spec2d_inject_uniform(zdev_spec2d* s, int ppcx, int ppcy, const float ufl[3], const float uth[3], uint64_t seed) {
	int npc = ppcx * ppcy;
	std::vector<int> cnt(s->ntiles);
	int64_t np = 0;
	for (int ty = 0; ty < s->nty; ty++) for (int tx = 0; tx < s->ntx; tx++) {
		int cx = (tx + 1) * s->TX <= s->nx ? s->TX : s->nx - tx * s->TX;
		int cy = (ty + 1) * s->TY <= s->ny ? s->TY : s->ny - ty * s->TY;
		cnt[tx + ty * s->ntx] = cx * cy * npc; np += (int64_t) cx * cy * npc;
	}
	spec_layout(s, cnt, np);
	f3 fl = {ufl[0], ufl[1], ufl[2]}, th = {uth[0], uth[1], uth[2]};
	int64_t ncell = (int64_t) s->nx * s->ny;
	ZDEV_LAUNCH(k_inject_uniform, zdev_div_up(ncell, 128), 128, 0, s->p, s->tile_off, s->tile_np,
	            s->nx, s->ny, s->TX, s->TY, s->ntx, ppcx, ppcy, fl, th, seed);
	s->np_host = np;
	s->ids_valid = s->track_ids && np < 0x7fffffff;
}

// ------------------------------------------------------------------ the push

struct push_geom {
	int nx, ny, nrow;        // grid
	int TX, TY, ntx;         // tiling
};

// scatter one segment's 8 contributions into the global J grid (L2 reductions)
__device__ __forceinline__ void deposit_seg_global(f3* __restrict__ J, int nrow, const seg2d& s, float qnx, float qny) {
	float w[8];
	seg_weights(s, qnx, qny, w);
	f3* c = J + (s.ix + 1) + (s.iy + 1) * nrow;
	atomicAdd(&c[0].x, w[0]);
	atomicAdd(&c[nrow].x, w[1]);
	atomicAdd(&c[0].y, w[2]);
	atomicAdd(&c[1].y, w[3]);
	atomicAdd(&c[0].z, w[4]);
	atomicAdd(&c[1].z, w[5]);
	atomicAdd(&c[nrow].z, w[6]);
	atomicAdd(&c[nrow + 1].z, w[7]);
}

// One CTA per tile.  Dynamic shared memory: E and B neighbourhoods, (TX+2)*(TY+2) f3 each.
__global__ void __launch_bounds__(PUSH_THREADS)
k_push2d(soa2d p, const int64_t* __restrict__ tile_off, int* __restrict__ tile_np, soa2d mig, unsigned int mig_cap,
         ctl2d* __restrict__ ctl, const f3* __restrict__ E, const f3* __restrict__ B, f3* __restrict__ J,
         push_geom g, zdev_push2d_params prm) {
	extern __shared__ f3 s_fld[];
	__shared__ int s_wcnt[2][PUSH_THREADS / 32];
	__shared__ double s_en[PUSH_THREADS / 32];

	const int t = blockIdx.x;
	const int tx = t % g.ntx, ty = t / g.ntx;
	const int x0 = tx * g.TX, y0 = ty * g.TY;
	const int cx = min(g.TX, g.nx - x0), cy = min(g.TY, g.ny - y0);
	const int srow = g.TX + 2;
	f3* sE = s_fld;
	f3* sB = s_fld + srow * (g.TY + 2);

	// stage the field neighbourhood: cells [x0-1, x0+cx] x [y0-1, y0+cy]
	for (int k = threadIdx.x; k < (cx + 2) * (cy + 2); k += blockDim.x) {
		int r = k / (cx + 2), c = k - r * (cx + 2);
		int gi = (x0 + c) + (y0 + r) * g.nrow;        // buffer index of cell (x0-1+c, y0-1+r)
		sE[c + r * srow] = E[gi];
		sB[c + r * srow] = B[gi];
	}
	__syncthreads();
	// cell (i,j) of the grid -> sE[(i-x0+1) + (j-y0+1)*srow]
	const f3* sE0 = sE + 1 + srow;
	const f3* sB0 = sB + 1 + srow;

	const int n = tile_np[t];
	const int64_t base = tile_off[t];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int run_out = 0;
	double energy = 0.0;

	for (int c0 = 0, it = 0; c0 < n; c0 += PUSH_THREADS, it ^= 1) {
		const int i = c0 + threadIdx.x;
		const bool active = i < n;
		int ix = 0, iy = 0, tag = 0;
		float x = 0, y = 0, ux = 0, uy = 0, uz = 0;
		int fate = 0;                                 // 0 drop / inactive, 1 stay, 2 migrate
		if (active) {
			const int64_t k = base + i;
			ix = p.ix[k]; iy = p.iy[k]; x = p.x[k]; y = p.y[k];
			ux = p.ux[k]; uy = p.uy[k]; uz = p.uz[k];
			if (p.tag) tag = p.tag[k];

			f3 Ep, Bp;
			interp_EB(sE0, sB0, srow, ix - x0, iy - y0, x, y, Ep, Bp);
			energy += boris(Ep, Bp, prm.tem, ux, uy, uz);

			float rg = 1.0f / sqrtf(1.0f + ux * ux + uy * uy + uz * uz);
			float dx = prm.dt_dx * rg * ux;
			float dy = prm.dt_dy * rg * uy;
			float x1 = x + dx, y1 = y + dy;
			int di = ltrim(x1), dj = ltrim(y1);
			x1 -= di; y1 -= dj;
			float qvz = prm.q * uz * rg;

			seg2d vp[3];
			int vnp = split_trajectory(ix, iy, di, dj, x, y, dx, dy, qvz, vp);
			deposit_seg_global(J, g.nrow, vp[0], prm.qnx, prm.qny);
			if (vnp > 1) deposit_seg_global(J, g.nrow, vp[1], prm.qnx, prm.qny);
			if (vnp > 2) deposit_seg_global(J, g.nrow, vp[2], prm.qnx, prm.qny);

			x = x1; y = y1;
			ix += di - prm.shift_window; iy += dj;
			// boundaries (reference particles.c:1237-1259)
			fate = 1;
			if (prm.moving_window) {
				if (ix < 0 || ix >= g.nx) fate = 0;
			} else {
				ix += ((ix < 0) ? g.nx : 0) - ((ix >= g.nx) ? g.nx : 0);
			}
			iy += ((iy < 0) ? g.ny : 0) - ((iy >= g.ny) ? g.ny : 0);
			if (fate && (ix < x0 || ix >= x0 + cx || iy < y0 || iy >= y0 + cy)) fate = 2;
		}

		// in-place compaction of the survivors of this chunk
		const unsigned stay_m = __ballot_sync(0xffffffffu, fate == 1);
		const unsigned mig_m = __ballot_sync(0xffffffffu, fate == 2);
		if (lane == 0) s_wcnt[it][warp] = __popc(stay_m);
		__syncthreads();      // every load of this chunk precedes every store
		int woff = 0, total = 0;
		#pragma unroll
		for (int w = 0; w < PUSH_THREADS / 32; w++) {
			int c = s_wcnt[it][w];
			woff += (w < warp) ? c : 0;
			total += c;
		}
		if (fate == 1) {
			const int64_t d = base + run_out + woff + __popc(stay_m & ((1u << lane) - 1));
			p.ix[d] = ix; p.iy[d] = iy; p.x[d] = x; p.y[d] = y; p.ux[d] = ux; p.uy[d] = uy; p.uz[d] = uz;
			if (p.tag) p.tag[d] = tag;
		}
		run_out += total;
		if (mig_m) {
			unsigned int mbase = 0;
			if (lane == 0) mbase = atomicAdd(&ctl->n_mig, (unsigned int) __popc(mig_m));
			mbase = __shfl_sync(0xffffffffu, mbase, 0);
			if (fate == 2) {
				unsigned int d = mbase + __popc(mig_m & ((1u << lane) - 1));
				if (d < mig_cap) {
					mig.ix[d] = ix; mig.iy[d] = iy; mig.x[d] = x; mig.y[d] = y;
					mig.ux[d] = ux; mig.uy[d] = uy; mig.uz[d] = uz;
					if (mig.tag) mig.tag[d] = tag;
				} else atomicOr(&ctl->flags, 2u);
			}
		}
	}

	// tile epilogue: population, energy
	for (int o = 16; o > 0; o >>= 1) energy += __shfl_down_sync(0xffffffffu, energy, o);
	if (lane == 0) s_en[warp] = energy;
	__syncthreads();
	if (threadIdx.x == 0) {
		double e = 0;
		for (int w = 0; w < PUSH_THREADS / 32; w++) e += s_en[w];
		if (n > 0) atomicAdd(&ctl->energy, e);
		tile_np[t] = run_out;
		if (run_out) atomicAdd(&ctl->np, (unsigned long long) run_out);
	}
}

// append the migrants to their destination tiles
__global__ void k_migrate2d(soa2d p, const int64_t* __restrict__ tile_off, int* __restrict__ tile_np, soa2d mig,
                            unsigned int mig_cap, ctl2d* __restrict__ ctl, push_geom g) {
	unsigned int n = ctl->n_mig;
	if (n > mig_cap) n = mig_cap;
	unsigned int accepted = 0;
	for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
		int ix = mig.ix[k], iy = mig.iy[k];
		int t = ix / g.TX + (iy / g.TY) * g.ntx;
		int slot = atomicAdd(&tile_np[t], 1);
		int64_t d = tile_off[t] + slot;
		if (d >= tile_off[t + 1]) { atomicOr(&ctl->flags, 1u); continue; }
		p.ix[d] = ix; p.iy[d] = iy; p.x[d] = mig.x[k]; p.y[d] = mig.y[k];
		p.ux[d] = mig.ux[k]; p.uy[d] = mig.uy[k]; p.uz[d] = mig.uz[k];
		if (p.tag) p.tag[d] = mig.tag[k];
		accepted++;
	}
	for (int o = 16; o > 0; o >>= 1) accepted += __shfl_down_sync(0xffffffffu, accepted, o);
	if ((threadIdx.x & 31) == 0 && accepted) atomicAdd(&ctl->np, (unsigned long long) accepted);
}

extern "C" void zdev_spec2d_advance(zdev_spec2d* s, zdev_grid2d* grid, zdev_grid2d* gcur, const zdev_push2d_params* prm) {
	if (zdev_grid2d_nx(grid) != s->nx || zdev_grid2d_ny(grid) != s->ny ||
	    zdev_grid2d_nx(gcur) != s->nx || zdev_grid2d_ny(gcur) != s->ny) {
		fprintf(stderr, "(*error*) zdev_spec2d_advance: species / grid size mismatch\n"); exit(-1);
	}
	ZDEV_CHECK(cudaMemsetAsync(s->ctl, 0, sizeof(ctl2d), zdev_strm));
	if (!s->cap_total) return;
	push_geom g = { s->nx, s->ny, s->nx + 3, s->TX, s->TY, s->ntx };
	size_t smem = (size_t) 2 * (s->TX + 2) * (s->TY + 2) * sizeof(f3);
	ZDEV_LAUNCH(k_push2d, s->ntiles, PUSH_THREADS, smem, s->p, s->tile_off, s->tile_np, s->mig, s->mig_cap, s->ctl,
	            zdev_grid2d_Epart(grid), zdev_grid2d_Bpart(grid), zdev_grid2d_J(gcur), g, *prm);
	int mg = 2 * zdev_num_sm;
	ZDEV_LAUNCH(k_migrate2d, mg, 256, 0, s->p, s->tile_off, s->tile_np, s->mig, s->mig_cap, s->ctl, g);
	if (prm->moving_window) s->ids_valid = 0;
}

extern "C" void zdev_spec2d_fetch(zdev_spec2d* s, double* energy_sum, int64_t* np) {
	ctl2d h;
	ZDEV_CHECK(cudaMemcpyAsync(&h, s->ctl, sizeof(h), cudaMemcpyDeviceToHost, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	check_flags(s, h.flags);
	s->np_host = (int64_t) h.np;
	if (energy_sum) *energy_sum = h.energy;
	if (np) *np = (int64_t) h.np;
}

// ------------------------------------------------------------------ charge deposit

// reference spec_deposit_charge, em2d/particles.c:1289-1324 (node centred, linear)
__global__ void k_deposit_charge(soa2d p, const int64_t* __restrict__ off, const int* __restrict__ tile_np,
                                 float* __restrict__ rho, int nrow, float q) {
	int t = blockIdx.x;
	int n = tile_np[t];
	int64_t b = off[t];
	for (int k = threadIdx.x; k < n; k += blockDim.x) {
		int idx = p.ix[b + k] + nrow * p.iy[b + k];
		float w1 = p.x[b + k], w2 = p.y[b + k];
		atomicAdd(&rho[idx], (1.0f - w1) * (1.0f - w2) * q);
		atomicAdd(&rho[idx + 1], (w1) * (1.0f - w2) * q);
		atomicAdd(&rho[idx + nrow], (1.0f - w1) * (w2) * q);
		atomicAdd(&rho[idx + 1 + nrow], (w1) * (w2) * q);
	}
}
__global__ void k_charge_fold(float* __restrict__ rho, int nx, int ny, int moving_window) {
	int nrow = nx + 1;
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	// x fold first (all rows), then y fold: done by two launches of this kernel
	if (moving_window >= 0) { if (!moving_window && k <= ny) rho[(size_t) k * nrow] += rho[nx + (size_t) k * nrow]; }
	else { if (k <= nx) rho[k] += rho[k + (size_t) ny * nrow]; }
}

extern "C" void zdev_spec2d_deposit_charge(zdev_spec2d* s, float q, int moving_window, float* charge) {
	size_t n = (size_t) (s->nx + 1) * (s->ny + 1);
	float* d_rho; ZDEV_CHECK(cudaMalloc(&d_rho, n * sizeof(float)));
	ZDEV_CHECK(cudaMemcpyAsync(d_rho, charge, n * sizeof(float), cudaMemcpyHostToDevice, zdev_strm));
	if (s->cap_total) ZDEV_LAUNCH(k_deposit_charge, s->ntiles, 256, 0, s->p, s->tile_off, s->tile_np, d_rho, s->nx + 1, q);
	ZDEV_LAUNCH(k_charge_fold, zdev_div_up(s->ny + 1, 128), 128, 0, d_rho, s->nx, s->ny, moving_window ? 1 : 0);
	ZDEV_LAUNCH(k_charge_fold, zdev_div_up(s->nx + 1, 128), 128, 0, d_rho, s->nx, s->ny, -1);
	ZDEV_CHECK(cudaMemcpyAsync(charge, d_rho, n * sizeof(float), cudaMemcpyDeviceToHost, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	cudaFree(d_rho);
}
