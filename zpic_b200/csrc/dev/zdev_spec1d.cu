// zpic-b200 :: em1d particle species on the device.
//
// One-dimensional twin of zdev_spec2d.cu (see there for the design): the grid is cut into tiles of TX
// cells, each tile owns a fixed-capacity segment of 20-byte records {x, ux, uy, uz, cell} plus a 16-bit
// sort key per slot, double buffered (A -> B every step).  One CTA per tile: stage the E/B neighbourhood
// in shared memory, counting-sort the slot indices by cell, then stream the particles in cell order with
// a software-pipelined loop; the five current contributions of non-crossing particles are combined per
// run of equal cell with a segmented warp scan, cell-crossing particles go through a warp-private queue
// and are split into their two segments 32 at a time.
//
// Replaces reference em1d/particles.c:919-1074 (spec_advance: interpolate_fld :864-886, Boris push,
// dep_current_zamb :707-779, periodic / open / moving-window boundaries) and :791-850 (spec_sort).
// Arithmetic order follows the reference exactly (--fmad=false): one step is bit-identical.
#include "zdev_common.cuh"
#include "pic2d_core.cuh"      // div_exact / sqrt_exact / ltrim
#include <vector>
#include <cstring>

f3* zdev_grid1d_Epart(zdev_grid1d* g);
f3* zdev_grid1d_Bpart(zdev_grid1d* g);
f3* zdev_grid1d_J(zdev_grid1d* g);
int zdev_grid1d_nx(zdev_grid1d* g);

struct part1_aos { int ix; float x, ux, uy, uz; };           // host record (em1d/particles.h:29-35)
struct rec20 { float x, ux, uy, uz; int cell; };             // device record; cell = tile-local index
#define KEY1_EMPTY 0xffffu

struct buf1d {
	rec20* rec;
	unsigned short* key;     // tile buffers only
	int* tag;                // optional
};

struct ctl1d {
	double energy;
	unsigned long long np;
	unsigned int n_mig;
	unsigned int flags;      // 1 tile overflow, 2 migrants overflow
};

struct zdev_spec1d {
	int nx, TX, ntiles, ppc_hint, track_ids;
	double slack;
	int64_t cap_total;
	int max_cap;
	buf1d p, q;
	int64_t* tile_off;
	int *tile_np, *tile_np_q;
	rec20* mig; int* mig_tag;        // migrants: cell = GLOBAL cell index
	unsigned int mig_cap;
	ctl1d* ctl;
	int64_t np_host;
	int ids_valid;
	std::vector<int64_t>* h_off;
	std::vector<cudaEvent_t>* ev; int ev_next, ev_pending; double push_ms; int64_t push_launches;
};

static const int P1_THREADS = 256;
static const int P1_WARPS = P1_THREADS / 32;
static const int XQ1_CAP = 64;
static const int EV1_RING = 64;

static void buf_alloc(buf1d& b, int64_t n, int with_tag) {
	size_t nn = (size_t) (n > 0 ? n : 1);
	memset(&b, 0, sizeof b);
	ZDEV_CHECK(cudaMalloc(&b.rec, nn * sizeof(rec20)));
	ZDEV_CHECK(cudaMalloc(&b.key, nn * 2));
	if (with_tag) ZDEV_CHECK(cudaMalloc(&b.tag, nn * 4));
}
static void buf_free(buf1d& b) { cudaFree(b.rec); cudaFree(b.key); cudaFree(b.tag); memset(&b, 0, sizeof b); }

extern "C" zdev_spec1d* zdev_spec1d_create(int nx, int ppc_hint, int track_ids) {
	zdev_require_init();
	zdev_spec1d* s = new zdev_spec1d();
	memset(s, 0, sizeof(*s));
	s->nx = nx; s->ppc_hint = ppc_hint > 0 ? ppc_hint : 1; s->track_ids = track_ids;
	int tx = 512;
	while (tx > 4 && (int64_t) tx * s->ppc_hint > 8192) tx >>= 1;
	if (const char* e = getenv("ZPIC_TILE_X1D")) tx = atoi(e);
	if (tx < 1) tx = 1; if (tx > 512) tx = 512;
	s->TX = tx;
	s->ntiles = (nx + tx - 1) / tx;
	if (const char* e = getenv("ZPIC_TILE_SLACK")) s->slack = atof(e);
	ZDEV_CHECK(cudaMalloc(&s->tile_off, (size_t) (s->ntiles + 1) * sizeof(int64_t)));
	ZDEV_CHECK(cudaMalloc(&s->tile_np, (size_t) s->ntiles * sizeof(int)));
	ZDEV_CHECK(cudaMalloc(&s->tile_np_q, (size_t) s->ntiles * sizeof(int)));
	ZDEV_CHECK(cudaMemsetAsync(s->tile_np, 0, (size_t) s->ntiles * sizeof(int), zdev_strm));
	ZDEV_CHECK(cudaMemsetAsync(s->tile_np_q, 0, (size_t) s->ntiles * sizeof(int), zdev_strm));
	ZDEV_CHECK(cudaMalloc(&s->ctl, sizeof(ctl1d)));
	ZDEV_CHECK(cudaMemsetAsync(s->ctl, 0, sizeof(ctl1d), zdev_strm));
	s->h_off = new std::vector<int64_t>();
	return s;
}

static void free_particles(zdev_spec1d* s) {
	if (s->cap_total) { buf_free(s->p); buf_free(s->q); cudaFree(s->mig); cudaFree(s->mig_tag); s->mig = nullptr; s->mig_tag = nullptr; }
	s->cap_total = 0; s->mig_cap = 0;
}

extern "C" void zdev_spec1d_destroy(zdev_spec1d* s) {
	if (!s) return;
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	free_particles(s);
	cudaFree(s->tile_off); cudaFree(s->tile_np); cudaFree(s->tile_np_q); cudaFree(s->ctl);
	if (s->ev) { for (auto& e : *s->ev) cudaEventDestroy(e); delete s->ev; }
	delete s->h_off;
	delete s;
}

static void layout(zdev_spec1d* s, const std::vector<int>& cnt, int64_t np) {
	double slack = s->slack;
	if (slack <= 0.0) slack = (np > (int64_t) 200000000) ? 1.25 : 2.0;
	std::vector<int64_t>& off = *s->h_off;
	off.assign(s->ntiles + 1, 0);
	int64_t max_cap = 0;
	for (int t = 0; t < s->ntiles; t++) {
		int cx = (t + 1) * s->TX <= s->nx ? s->TX : s->nx - t * s->TX;
		int64_t nominal = (int64_t) cx * s->ppc_hint;
		int64_t want = cnt[t] > nominal ? cnt[t] : nominal;
		int64_t cap = (int64_t) (want * slack) + 64;
		cap = (cap + 31) & ~(int64_t) 31;
		off[t + 1] = off[t] + cap;
		if (cap > max_cap) max_cap = cap;
	}
	if (max_cap * 4 > 160 * 1024) {
		fprintf(stderr, "(*error*) zpic-b200: %lld particles in one %d-cell tile exceed the shared-memory index "
		        "buffer; use smaller tiles (ZPIC_TILE_X1D)\n", (long long) max_cap, s->TX);
		exit(-1);
	}
	int64_t total = off[s->ntiles];
	free_particles(s);
	buf_alloc(s->p, total, s->track_ids);
	buf_alloc(s->q, total, s->track_ids);
	s->cap_total = total; s->max_cap = (int) max_cap;
	int64_t mc = total / 8 + 65536;
	if (mc > 0x7fffffff) mc = 0x7fffffff;
	s->mig_cap = (unsigned int) mc;
	ZDEV_CHECK(cudaMalloc(&s->mig, (size_t) mc * sizeof(rec20)));
	if (s->track_ids) ZDEV_CHECK(cudaMalloc(&s->mig_tag, (size_t) mc * 4));
	ZDEV_CHECK(cudaMemcpyAsync(s->tile_off, off.data(), (size_t) (s->ntiles + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, zdev_strm));
	ZDEV_CHECK(cudaMemsetAsync(s->tile_np, 0, (size_t) s->ntiles * sizeof(int), zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
}

// ------------------------------------------------------------------ host <-> device

__global__ void k1_count_tiles(const part1_aos* __restrict__ a, int64_t np, int TX, int* cnt) {
	int64_t k = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (k < np) atomicAdd(&cnt[a[k].ix / TX], 1);
}
__global__ void k1_scatter(const part1_aos* __restrict__ a, int64_t np, int TX, buf1d p, const int64_t* __restrict__ off,
                           int* tile_np, ctl1d* ctl, int tag0) {
	int64_t k = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= np) return;
	part1_aos r = a[k];
	int t = r.ix / TX;
	int slot = atomicAdd(&tile_np[t], 1);
	int64_t d = off[t] + slot;
	if (d >= off[t + 1]) { atomicOr(&ctl->flags, 1u); return; }
	rec20 v = { r.x, r.ux, r.uy, r.uz, r.ix - t * TX };
	p.rec[d] = v;
	p.key[d] = (unsigned short) v.cell;
	if (p.tag) p.tag[d] = tag0 + (int) k;
}

static void check_flags(zdev_spec1d* s, unsigned int flags) {
	if (flags & 1u) {
		fprintf(stderr, "(*error*) zpic-b200: particle tile capacity exceeded (%d-cell tiles); raise ZPIC_TILE_SLACK "
		        "(current %.2f) and rerun, aborting.\n", s->TX, s->slack);
		exit(-1);
	}
	if (flags & 2u) { fprintf(stderr, "(*error*) zpic-b200: particle migration list overflow (capacity %u), aborting.\n", s->mig_cap); exit(-1); }
}

extern "C" void zdev_spec1d_upload(zdev_spec1d* s, const void* part, int64_t np) {
	part1_aos* d_aos = nullptr;
	std::vector<int> cnt(s->ntiles, 0);
	if (np > 0) {
		ZDEV_CHECK(cudaMalloc(&d_aos, (size_t) np * sizeof(part1_aos)));
		ZDEV_CHECK(cudaMemcpyAsync(d_aos, part, (size_t) np * sizeof(part1_aos), cudaMemcpyHostToDevice, zdev_strm));
		int* d_cnt; ZDEV_CHECK(cudaMalloc(&d_cnt, (size_t) s->ntiles * sizeof(int)));
		ZDEV_CHECK(cudaMemsetAsync(d_cnt, 0, (size_t) s->ntiles * sizeof(int), zdev_strm));
		ZDEV_LAUNCH(k1_count_tiles, zdev_div_up(np, 256), 256, 0, d_aos, np, s->TX, d_cnt);
		ZDEV_CHECK(cudaMemcpyAsync(cnt.data(), d_cnt, (size_t) s->ntiles * sizeof(int), cudaMemcpyDeviceToHost, zdev_strm));
		ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
		cudaFree(d_cnt);
	}
	bool fits = s->cap_total > 0;
	if (fits) { const std::vector<int64_t>& off = *s->h_off; for (int t = 0; t < s->ntiles && fits; t++) fits = cnt[t] <= off[t + 1] - off[t]; }
	if (fits) ZDEV_CHECK(cudaMemsetAsync(s->tile_np, 0, (size_t) s->ntiles * sizeof(int), zdev_strm));
	else layout(s, cnt, np);
	ZDEV_CHECK(cudaMemsetAsync(s->ctl, 0, sizeof(ctl1d), zdev_strm));
	if (np > 0) ZDEV_LAUNCH(k1_scatter, zdev_div_up(np, 256), 256, 0, d_aos, np, s->TX, s->p, s->tile_off, s->tile_np, s->ctl, 0);
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	if (d_aos) cudaFree(d_aos);
	s->np_host = np;
	s->ids_valid = s->track_ids;
}

extern "C" void zdev_spec1d_append(zdev_spec1d* s, const void* part, int64_t np) {
	if (np <= 0) return;
	if (!s->cap_total) { zdev_spec1d_upload(s, part, np); return; }
	part1_aos* d_aos;
	ZDEV_CHECK(cudaMalloc(&d_aos, (size_t) np * sizeof(part1_aos)));
	ZDEV_CHECK(cudaMemcpyAsync(d_aos, part, (size_t) np * sizeof(part1_aos), cudaMemcpyHostToDevice, zdev_strm));
	ZDEV_LAUNCH(k1_scatter, zdev_div_up(np, 256), 256, 0, d_aos, np, s->TX, s->p, s->tile_off, s->tile_np, s->ctl, (int) s->np_host);
	ctl1d h;
	ZDEV_CHECK(cudaMemcpyAsync(&h, s->ctl, sizeof h, cudaMemcpyDeviceToHost, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	cudaFree(d_aos);
	check_flags(s, h.flags);
	s->np_host += np;
}

__global__ void k1_count_live(buf1d p, const int64_t* __restrict__ off, const int* __restrict__ tile_np, int* live) {
	int t = blockIdx.x, n = tile_np[t], c = 0;
	int64_t b = off[t];
	for (int k = threadIdx.x; k < n; k += blockDim.x) c += (p.key[b + k] != KEY1_EMPTY);
	for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
	if ((threadIdx.x & 31) == 0 && c) atomicAdd(&live[t], c);
}

__global__ void k1_gather(buf1d p, const int64_t* __restrict__ off, const int* __restrict__ tile_np,
                          const int64_t* __restrict__ prefix, part1_aos* __restrict__ out, int by_tag, int TX) {
	int t = blockIdx.x, n = tile_np[t];
	int64_t b = off[t], o = prefix[t];
	__shared__ int s_run;
	if (threadIdx.x == 0) s_run = 0;
	__syncthreads();
	for (int k0 = 0; k0 < n; k0 += blockDim.x) {
		int k = k0 + threadIdx.x;
		bool live = (k < n) && p.key[b + k] != KEY1_EMPTY;
		unsigned m = __ballot_sync(0xffffffffu, live);
		int wbase = 0;
		if ((threadIdx.x & 31) == 0 && m) wbase = atomicAdd(&s_run, __popc(m));
		wbase = __shfl_sync(0xffffffffu, wbase, 0);
		if (live) {
			rec20 v = p.rec[b + k];
			part1_aos r = { t * TX + v.cell, v.x, v.ux, v.uy, v.uz };
			int64_t d = by_tag ? (int64_t) p.tag[b + k] : o + wbase + __popc(m & ((1u << (threadIdx.x & 31)) - 1));
			out[d] = r;
		}
	}
}

static int64_t live_counts(zdev_spec1d* s, std::vector<int>& cnt) {
	int* d_live; ZDEV_CHECK(cudaMalloc(&d_live, (size_t) s->ntiles * sizeof(int)));
	ZDEV_CHECK(cudaMemsetAsync(d_live, 0, (size_t) s->ntiles * sizeof(int), zdev_strm));
	ZDEV_LAUNCH(k1_count_live, s->ntiles, 256, 0, s->p, s->tile_off, s->tile_np, d_live);
	cnt.resize(s->ntiles);
	ZDEV_CHECK(cudaMemcpyAsync(cnt.data(), d_live, (size_t) s->ntiles * sizeof(int), cudaMemcpyDeviceToHost, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	cudaFree(d_live);
	int64_t np = 0;
	for (int t = 0; t < s->ntiles; t++) np += cnt[t];
	return np;
}

extern "C" int64_t zdev_spec1d_np(zdev_spec1d* s) {
	if (!s->cap_total) return 0;
	std::vector<int> cnt;
	return s->np_host = live_counts(s, cnt);
}

extern "C" int64_t zdev_spec1d_download(zdev_spec1d* s, void* part, int64_t max_np) {
	if (!s->cap_total) return 0;
	std::vector<int> cnt;
	int64_t np = live_counts(s, cnt);
	s->np_host = np;
	if (np == 0) return 0;
	if (np > max_np) { fprintf(stderr, "(*error*) zpic-b200: host particle buffer too small (%lld > %lld)\n", (long long) np, (long long) max_np); exit(-1); }
	std::vector<int64_t> prefix(s->ntiles);
	int64_t acc = 0;
	for (int t = 0; t < s->ntiles; t++) { prefix[t] = acc; acc += cnt[t]; }
	int64_t* d_prefix; part1_aos* d_aos;
	ZDEV_CHECK(cudaMalloc(&d_prefix, (size_t) s->ntiles * sizeof(int64_t)));
	ZDEV_CHECK(cudaMalloc(&d_aos, (size_t) np * sizeof(part1_aos)));
	ZDEV_CHECK(cudaMemcpyAsync(d_prefix, prefix.data(), (size_t) s->ntiles * sizeof(int64_t), cudaMemcpyHostToDevice, zdev_strm));
	ZDEV_LAUNCH(k1_gather, s->ntiles, 256, 0, s->p, s->tile_off, s->tile_np, d_prefix, d_aos, (s->track_ids && s->ids_valid) ? 1 : 0, s->TX);
	ZDEV_CHECK(cudaMemcpyAsync(part, d_aos, (size_t) np * sizeof(part1_aos), cudaMemcpyDeviceToHost, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	cudaFree(d_prefix); cudaFree(d_aos);
	return np;
}

// ------------------------------------------------------------------ device-side uniform injection

__device__ __forceinline__ uint64_t mix64_1d(uint64_t z) {
	z += 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}
__device__ __forceinline__ void normal3_1d(uint64_t seed, uint64_t gid, float& a, float& b, float& c) {
	uint64_t r0 = mix64_1d(seed ^ (gid * 2 + 0) * 0xD1342543DE82EF95ull);
	uint64_t r1 = mix64_1d(seed ^ (gid * 2 + 1) * 0xD1342543DE82EF95ull);
	float u0 = ((uint32_t) (r0 >> 40) + 0.5f) * (1.0f / 16777216.0f), u1 = ((uint32_t) (r0 & 0xffffff) + 0.5f) * (1.0f / 16777216.0f);
	float u2 = ((uint32_t) (r1 >> 40) + 0.5f) * (1.0f / 16777216.0f), u3 = ((uint32_t) (r1 & 0xffffff) + 0.5f) * (1.0f / 16777216.0f);
	float m0 = sqrtf(-2.0f * logf(u0)), m1 = sqrtf(-2.0f * logf(u2)), s0, c0, s1, c1;
	sincospif(2.0f * u1, &s0, &c0); sincospif(2.0f * u3, &s1, &c1);
	a = m0 * c0; b = m0 * s0; c = m1 * c1; (void) s1;
}

// one warp per cell (ppc is large in 1D): positions (k+0.5)/ppc (em1d/particles.c:245-250), thermal momenta
// minus the cell mean plus fluid (em1d/particles.c:88-130)
__global__ void k1_inject_uniform(buf1d p, const int64_t* __restrict__ off, int* tile_np, int nx, int TX, int ppc,
                                  f3 ufl, f3 uth, uint64_t seed) {
	int cell = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (cell >= nx) return;
	int t = cell / TX, lc = cell - t * TX;
	int64_t base = off[t] + (int64_t) lc * ppc;
	uint64_t gid0 = (uint64_t) cell * ppc;
	float sx = 0, sy = 0, sz = 0;
	for (int k = lane; k < ppc; k += 32) { float a, b, c; normal3_1d(seed, gid0 + k, a, b, c); sx += uth.x * a; sy += uth.y * b; sz += uth.z * c; }
	for (int o = 16; o > 0; o >>= 1) { sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); sz += __shfl_xor_sync(0xffffffffu, sz, o); }
	float norm = 1.0f / ppc; sx *= norm; sy *= norm; sz *= norm;
	for (int k = lane; k < ppc; k += 32) {
		float a, b, c; normal3_1d(seed, gid0 + k, a, b, c);
		rec20 v = { (float) ((k + 0.5) / ppc), uth.x * a + (ufl.x - sx), uth.y * b + (ufl.y - sy), uth.z * c + (ufl.z - sz), lc };
		p.rec[base + k] = v; p.key[base + k] = (unsigned short) lc;
		if (p.tag) p.tag[base + k] = (int) (gid0 + k);
	}
	if (lc == 0 && lane == 0) { int cx = (t + 1) * TX <= nx ? TX : nx - t * TX; tile_np[t] = cx * ppc; }
}

extern "C" void zdev_spec1d_inject_uniform(zdev_spec1d* s, int ppc, const float ufl[3], const float uth[3], uint64_t seed) {
	std::vector<int> cnt(s->ntiles);
	int64_t np = 0;
	for (int t = 0; t < s->ntiles; t++) { int cx = (t + 1) * s->TX <= s->nx ? s->TX : s->nx - t * s->TX; cnt[t] = cx * ppc; np += cnt[t]; }
	layout(s, cnt, np);
	f3 fl = {ufl[0], ufl[1], ufl[2]}, th = {uth[0], uth[1], uth[2]};
	ZDEV_LAUNCH(k1_inject_uniform, zdev_div_up((int64_t) s->nx * 32, 256), 256, 0, s->p, s->tile_off, s->tile_np, s->nx, s->TX, ppc, fl, th, seed);
	s->np_host = np;
	s->ids_valid = s->track_ids && np < 0x7fffffff;
}

// ------------------------------------------------------------------ the push

struct seg1d { float x0, x1, dx, qvy, qvz; int ix; };

// the 5 contributions of one in-cell segment (em1d/particles.c:763-777): Jx[ix]; Jy[ix], Jy[ix+1]; Jz[ix], Jz[ix+1]
__device__ __forceinline__ void seg1_weights(const seg1d& s, float qnx, float w[5]) {
	float S0x0 = 1.0f - s.x0, S0x1 = s.x0, S1x0 = 1.0f - s.x1, S1x1 = s.x1;
	w[0] = qnx * s.dx;
	w[1] = s.qvy * (S0x0 + S1x0 + (S0x0 - S1x0) / 2.0f);
	w[2] = s.qvy * (S0x1 + S1x1 + (S0x1 - S1x1) / 2.0f);
	w[3] = s.qvz * (S0x0 + S1x0 + (S0x0 - S1x0) / 2.0f);
	w[4] = s.qvz * (S0x1 + S1x1 + (S0x1 - S1x1) / 2.0f);
}
__device__ __forceinline__ void red1(f3* __restrict__ c, const float w[5]) {
	atomicAdd(&c[0].x, w[0]);
	atomicAdd(&c[0].y, w[1]); atomicAdd(&c[1].y, w[2]);
	atomicAdd(&c[0].z, w[3]); atomicAdd(&c[1].z, w[4]);
}

struct xq1_entry { int ix, di; float x0, dx, qvy, qvz; };

// split a cell-crossing move into its two segments (em1d/particles.c:723-757) and deposit both
__device__ __forceinline__ void drain1(const xq1_entry* q, int n, int lane, f3* __restrict__ J, float qnx) {
	if (lane >= n) return;
	xq1_entry e = q[lane];
	seg1d a, b;
	a.x0 = e.x0; a.dx = e.dx; a.x1 = e.x0 + e.dx; a.qvy = e.qvy * 0.5f; a.qvz = e.qvz * 0.5f; a.ix = e.ix;
	const int ib = (e.di == 1);
	const float delta = (e.x0 + e.dx - ib) / e.dx;
	b.x0 = 1 - ib; b.x1 = (e.x0 + e.dx) - e.di; b.dx = e.dx * delta; b.ix = e.ix + e.di;
	b.qvy = a.qvy * delta; b.qvz = a.qvz * delta;
	a.x1 = ib; a.dx *= (1.0f - delta); a.qvy *= (1.0f - delta); a.qvz *= (1.0f - delta);
	float w[5];
	seg1_weights(a, qnx, w); red1(J + a.ix + 1, w);
	seg1_weights(b, qnx, w); red1(J + b.ix + 1, w);
}

// dynamic shared memory: perm[max_cap] ints, then 6 field planes of (TX+2) floats, then cnt[TX]
__global__ void __launch_bounds__(P1_THREADS, 2)
k_push1d(buf1d A, buf1d Bo, const int64_t* __restrict__ tile_off, const int* __restrict__ tile_np, int* __restrict__ tile_np_out,
         rec20* __restrict__ mig, int* __restrict__ mig_tag, unsigned int mig_cap, ctl1d* __restrict__ ctl,
         const f3* __restrict__ E, const f3* __restrict__ B, f3* __restrict__ J, int nx, int TX, int max_cap,
         zdev_push1d_params prm) {
	extern __shared__ int smem[];
	int* s_perm = smem;
	float* s_fld = reinterpret_cast<float*>(smem + max_cap);
	int* s_cnt = reinterpret_cast<int*>(s_fld + 6 * (TX + 2));
	__shared__ int s_wsum[P1_WARPS];
	__shared__ xq1_entry s_xq[P1_WARPS][XQ1_CAP];

	const int t = blockIdx.x, x0 = t * TX, cx = min(TX, nx - x0);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int n = tile_np[t];
	const int64_t base = tile_off[t];
	const int PL = TX + 2;

	for (int k = threadIdx.x; k < cx + 2; k += P1_THREADS) {        // cells x0-1 .. x0+cx
		f3 e = E[x0 + k], b = B[x0 + k];
		s_fld[k] = e.x; s_fld[k + PL] = e.y; s_fld[k + 2 * PL] = e.z;
		s_fld[k + 3 * PL] = b.x; s_fld[k + 4 * PL] = b.y; s_fld[k + 5 * PL] = b.z;
	}
	for (int k = threadIdx.x; k < TX; k += P1_THREADS) s_cnt[k] = 0;
	__syncthreads();

	// ---- phase A: index sort by cell
	for (int i = threadIdx.x; i < n; i += P1_THREADS) {
		unsigned c = A.key[base + i];
		if (c != KEY1_EMPTY) atomicAdd(&s_cnt[c], 1);
	}
	__syncthreads();
	int nlive;
	{	// exclusive scan over TX <= 512 counters, two per thread
		const int i0 = 2 * threadIdx.x;
		int a = (i0 < TX) ? s_cnt[i0] : 0, b = (i0 + 1 < TX) ? s_cnt[i0 + 1] : 0;
		int v = a + b, incl = v;
		for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += u; }
		if (lane == 31) s_wsum[warp] = incl;
		__syncthreads();
		int woff = 0, tot = 0;
		#pragma unroll
		for (int w = 0; w < P1_WARPS; w++) { int c = s_wsum[w]; woff += (w < warp) ? c : 0; tot += c; }
		const int ex = woff + incl - v;
		if (i0 < TX) s_cnt[i0] = ex;
		if (i0 + 1 < TX) s_cnt[i0 + 1] = ex + a;
		nlive = tot;
		__syncthreads();
	}
	for (int i = threadIdx.x; i < n; i += P1_THREADS) {
		unsigned c = A.key[base + i];
		if (c != KEY1_EMPTY) s_perm[atomicAdd(&s_cnt[c], 1)] = i;
	}
	__syncthreads();

	// ---- phase B
	xq1_entry* xq = s_xq[warp];
	int nxq = 0;
	float energy = 0.0f;
	f3* const J0 = J + x0 + 1;
	const float* Ex = s_fld + 1;          // plane index 0 is cell -1 of the tile
	const int first = (nlive > 0) ? s_perm[0] : 0;
	rec20 nv; int ntag = 0;
	{
		const int pn = warp * 32 + lane;
		const int64_t k = base + ((pn < nlive) ? s_perm[pn] : first);
		nv = A.rec[k];
		if (A.tag) ntag = A.tag[k];
	}
	for (int p0 = warp * 32; p0 < nlive; p0 += P1_THREADS) {
		const int p = p0 + lane;
		const bool active = p < nlive;
		const rec20 v = nv;
		const int tag = ntag;
		{
			const int pn = p + P1_THREADS;
			const int64_t k = base + ((pn < nlive) ? s_perm[pn] : first);
			nv = A.rec[k];
			if (A.tag) ntag = A.tag[k];
		}
		const int lx = v.cell;
		const int key = active ? lx : 0x7fffffff;
		float x = v.x, ux = v.ux, uy = v.uy, uz = v.uz;
		float w[5];
		int fate, ncell = -1, gix = 0;
		bool crosses;
		xq1_entry xe;
		{
			// interpolate_fld (em1d/particles.c:864-886)
			const int h = (x < 0.5f) ? 1 : 0;
			const float w1 = x, w1h = x + (h ? 0.5f : -0.5f);
			const int i = lx, ih = lx - h;
			f3 Ep, Bp;
			Ep.x = Ex[ih] * (1.0f - w1h) + Ex[ih + 1] * w1h;
			Ep.y = Ex[i + PL] * (1.0f - w1) + Ex[i + 1 + PL] * w1;
			Ep.z = Ex[i + 2 * PL] * (1.0f - w1) + Ex[i + 1 + 2 * PL] * w1;
			Bp.x = Ex[i + 3 * PL] * (1.0f - w1) + Ex[i + 1 + 3 * PL] * w1;
			Bp.y = Ex[ih + 4 * PL] * (1.0f - w1h) + Ex[ih + 1 + 4 * PL] * w1h;
			Bp.z = Ex[ih + 5 * PL] * (1.0f - w1h) + Ex[ih + 1 + 5 * PL] * w1h;
			// Boris (em1d/particles.c:953-1000): same rotation as 2D; energy term u2/(1+gamma)
			const float en = boris(Ep, Bp, prm.tem, ux, uy, uz);
			energy += active ? en : 0.0f;

			const float rg = div_exact(1.0f, sqrt_exact(1.0f + ux * ux + uy * uy + uz * uz));
			const float dx = prm.dt_dx * rg * ux;
			float x1 = x + dx;
			const int di = ltrim(x1);
			x1 -= di;
			const float qvy = prm.q * uy * rg, qvz = prm.q * uz * rg;

			crosses = active && (di != 0);
			xe.ix = x0 + lx; xe.di = di; xe.x0 = x; xe.dx = dx; xe.qvy = qvy; xe.qvz = qvz;
			seg1d s0; s0.x0 = x; s0.dx = dx; s0.x1 = x + dx; s0.qvy = qvy * 0.5f; s0.qvz = qvz * 0.5f; s0.ix = 0;
			seg1_weights(s0, prm.qnx, w);
			const bool zero = crosses || !active;
			#pragma unroll
			for (int q = 0; q < 5; q++) w[q] = zero ? 0.0f : w[q];

			x = x1;
			int ix = x0 + lx + di - prm.shift_window;
			fate = active ? 1 : 0;
			if (prm.absorbing) { if (ix < 0 || ix >= nx) fate = 0; }
			else ix += ((ix < 0) ? nx : 0) - ((ix >= nx) ? nx : 0);
			const int nlx = ix - x0;
			if (fate) { if (nlx < 0 || nlx >= cx) { fate = 2; gix = ix; } else ncell = nlx; }
		}

		// current of the non-crossing particles: segmented inclusive scan over runs of equal cell
		{
			const int prev = __shfl_up_sync(0xffffffffu, key, 1);
			const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || key != prev);
			const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
			#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const bool take = (lane - d) >= start;
				#pragma unroll
				for (int q = 0; q < 5; q++) { float u = __shfl_up_sync(0xffffffffu, w[q], d); if (take) w[q] += u; }
			}
			const bool tail = (lane == 31) || ((heads >> (lane + 1)) & 1u);
			if (tail && active) red1(J0 + lx, w);
		}
		{
			const unsigned xm = __ballot_sync(0xffffffffu, crosses);
			if (xm) {
				if (crosses) xq[nxq + __popc(xm & ((1u << lane) - 1))] = xe;
				nxq += __popc(xm);
				__syncwarp();
				if (nxq >= 32) { drain1(xq + nxq - 32, 32, lane, J, prm.qnx); nxq -= 32; __syncwarp(); }
			}
		}
		if (active) {
			const int64_t d = base + p;
			Bo.key[d] = (fate == 1) ? (unsigned short) ncell : (unsigned short) KEY1_EMPTY;
			if (fate == 1) {
				rec20 o = { x, ux, uy, uz, ncell };
				Bo.rec[d] = o;
				if (Bo.tag) Bo.tag[d] = tag;
			}
		}
		const unsigned mig_m = __ballot_sync(0xffffffffu, fate == 2);
		if (mig_m) {
			unsigned int mbase = 0;
			if (lane == 0) mbase = atomicAdd(&ctl->n_mig, (unsigned int) __popc(mig_m));
			mbase = __shfl_sync(0xffffffffu, mbase, 0);
			if (fate == 2) {
				unsigned int d = mbase + __popc(mig_m & ((1u << lane) - 1));
				if (d < mig_cap) { rec20 o = { x, ux, uy, uz, gix }; mig[d] = o; if (mig_tag) mig_tag[d] = tag; }
				else atomicOr(&ctl->flags, 2u);
			}
		}
	}
	if (nxq) drain1(xq, nxq, lane, J, prm.qnx);

	double e = (double) energy;
	for (int o = 16; o > 0; o >>= 1) e += __shfl_down_sync(0xffffffffu, e, o);
	if (lane == 0 && nlive > 0) atomicAdd(&ctl->energy, e);
	if (threadIdx.x == 0) tile_np_out[t] = nlive;
}

__global__ void k1_migrate(buf1d p, const int64_t* __restrict__ tile_off, int* __restrict__ tile_np, const rec20* __restrict__ mig,
                           const int* __restrict__ mig_tag, unsigned int mig_cap, ctl1d* __restrict__ ctl, int TX) {
	unsigned int n = ctl->n_mig;
	if (n > mig_cap) n = mig_cap;
	for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
		rec20 v = mig[k];
		int t = v.cell / TX;
		int slot = atomicAdd(&tile_np[t], 1);
		int64_t d = tile_off[t] + slot;
		if (d >= tile_off[t + 1]) { atomicOr(&ctl->flags, 1u); continue; }
		v.cell -= t * TX;
		p.rec[d] = v; p.key[d] = (unsigned short) v.cell;
		if (p.tag) p.tag[d] = mig_tag[k];
	}
}

__global__ void k1_count_total(buf1d p, const int64_t* __restrict__ off, const int* __restrict__ tile_np, int ntiles, ctl1d* ctl) {
	unsigned long long c = 0;
	for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
		int n = tile_np[t]; int64_t b = off[t];
		for (int k = threadIdx.x; k < n; k += blockDim.x) c += (p.key[b + k] != KEY1_EMPTY);
	}
	for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
	if ((threadIdx.x & 31) == 0 && c) atomicAdd(&ctl->np, c);
}

static void collect_timing(zdev_spec1d* s) {
	if (!s->ev) return;
	for (int k = 0; k < s->ev_pending; k++) {
		int slot = (s->ev_next - s->ev_pending + k + EV1_RING) % EV1_RING;
		float ms = 0;
		ZDEV_CHECK(cudaEventSynchronize((*s->ev)[2 * slot + 1]));
		ZDEV_CHECK(cudaEventElapsedTime(&ms, (*s->ev)[2 * slot], (*s->ev)[2 * slot + 1]));
		s->push_ms += ms; s->push_launches++;
	}
	s->ev_pending = 0;
}
extern "C" void zdev_spec1d_push_timing(zdev_spec1d* s, double* total_ms, int64_t* launches, int reset) {
	collect_timing(s);
	if (total_ms) *total_ms = s->push_ms;
	if (launches) *launches = s->push_launches;
	if (reset) { s->push_ms = 0; s->push_launches = 0; }
}

extern "C" void zdev_spec1d_advance(zdev_spec1d* s, zdev_grid1d* grid, zdev_grid1d* gcur, const zdev_push1d_params* prm) {
	if (zdev_grid1d_nx(grid) != s->nx || zdev_grid1d_nx(gcur) != s->nx) {
		fprintf(stderr, "(*error*) zdev_spec1d_advance: species / grid size mismatch\n"); exit(-1);
	}
	ZDEV_CHECK(cudaMemsetAsync(s->ctl, 0, sizeof(ctl1d), zdev_strm));
	if (!s->cap_total) return;
	size_t smem = (size_t) s->max_cap * sizeof(int) + (size_t) 6 * (s->TX + 2) * sizeof(float) + (size_t) s->TX * sizeof(int);
	static size_t configured = 0;
	if (smem > configured) {
		ZDEV_CHECK(cudaFuncSetAttribute(k_push1d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		configured = smem;
	}
	int slot = -1;
	if (zdev_time_push) {
		if (!s->ev) { s->ev = new std::vector<cudaEvent_t>(2 * EV1_RING); for (auto& e : *s->ev) ZDEV_CHECK(cudaEventCreate(&e)); }
		if (s->ev_pending == EV1_RING) collect_timing(s);
		slot = s->ev_next; s->ev_next = (s->ev_next + 1) % EV1_RING; s->ev_pending++;
		ZDEV_CHECK(cudaEventRecord((*s->ev)[2 * slot], zdev_strm));
	}
	ZDEV_LAUNCH(k_push1d, s->ntiles, P1_THREADS, smem, s->p, s->q, s->tile_off, s->tile_np, s->tile_np_q, s->mig, s->mig_tag,
	            s->mig_cap, s->ctl, zdev_grid1d_Epart(grid), zdev_grid1d_Bpart(grid), zdev_grid1d_J(gcur), s->nx, s->TX, s->max_cap, *prm);
	if (slot >= 0) ZDEV_CHECK(cudaEventRecord((*s->ev)[2 * slot + 1], zdev_strm));
	{ buf1d t = s->p; s->p = s->q; s->q = t; }
	{ int* t = s->tile_np; s->tile_np = s->tile_np_q; s->tile_np_q = t; }
	ZDEV_LAUNCH(k1_migrate, 2 * zdev_num_sm, 256, 0, s->p, s->tile_off, s->tile_np, s->mig, s->mig_tag, s->mig_cap, s->ctl, s->TX);
	if (prm->absorbing) s->ids_valid = 0;
}

extern "C" void zdev_spec1d_fetch(zdev_spec1d* s, double* energy_sum, int64_t* np) {
	ctl1d h;
	memset(&h, 0, sizeof h);
	if (s->cap_total && np) ZDEV_LAUNCH(k1_count_total, 4 * zdev_num_sm, 256, 0, s->p, s->tile_off, s->tile_np, s->ntiles, s->ctl);
	ZDEV_CHECK(cudaMemcpyAsync(&h, s->ctl, sizeof h, cudaMemcpyDeviceToHost, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	check_flags(s, h.flags);
	if (np) {
		ZDEV_CHECK(cudaMemsetAsync(&s->ctl->np, 0, sizeof(unsigned long long), zdev_strm));
		s->np_host = (int64_t) h.np; *np = (int64_t) h.np;
	}
	if (energy_sum) *energy_sum = h.energy;
}

// ------------------------------------------------------------------ charge deposit (em1d/particles.c:1085-1106)

__global__ void k1_deposit_charge(buf1d p, const int64_t* __restrict__ off, const int* __restrict__ tile_np, float* __restrict__ rho, float q, int TX) {
	int t = blockIdx.x, n = tile_np[t];
	int64_t b = off[t];
	for (int k = threadIdx.x; k < n; k += blockDim.x) {
		if (p.key[b + k] == KEY1_EMPTY) continue;
		rec20 v = p.rec[b + k];
		int idx = t * TX + v.cell;
		atomicAdd(&rho[idx], (1.0f - v.x) * q);
		atomicAdd(&rho[idx + 1], (v.x) * q);
	}
}
__global__ void k1_charge_fold(float* rho, int nx) { rho[0] += rho[nx]; }

extern "C" void zdev_spec1d_deposit_charge(zdev_spec1d* s, float q, int moving_window, float* charge) {
	size_t n = (size_t) s->nx + 1;
	float* d_rho; ZDEV_CHECK(cudaMalloc(&d_rho, n * sizeof(float)));
	ZDEV_CHECK(cudaMemcpyAsync(d_rho, charge, n * sizeof(float), cudaMemcpyHostToDevice, zdev_strm));
	if (s->cap_total) ZDEV_LAUNCH(k1_deposit_charge, s->ntiles, 256, 0, s->p, s->tile_off, s->tile_np, d_rho, q, s->TX);
	if (!moving_window) ZDEV_LAUNCH(k1_charge_fold, 1, 1, 0, d_rho, s->nx);
	ZDEV_CHECK(cudaMemcpyAsync(charge, d_rho, n * sizeof(float), cudaMemcpyDeviceToHost, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	cudaFree(d_rho);
}
