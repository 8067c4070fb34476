// zpic-b200 :: em1d particle species on the device.
//
// One-dimensional twin of zdev_spec2d.cu (see there for the design): the grid is cut into tiles of TX
// cells, each tile owns a fixed-capacity segment of slots in chunks of 32, a chunk being five 128-byte
// rows  x[32] ux[32] uy[32] uz[32] cell[32]  (cell = tile-local index), plus a 16-bit sort key per slot,
// double buffered (A -> B every step).  One CTA per tile: TMA bulk load of the tile's keys while the E/B
// neighbourhood is staged, counting sort of the slot indices by cell in shared memory, then every warp
// streams a contiguous range of the sorted particles, 64 per iteration, two particles per thread as the
// halves of packed fp32 registers (FFMA2 / FADD2 / FMUL2, each half rounded like the scalar reference
// operation).  Every move deposits its first in-cell piece through a per-lane accumulator that is reduced
// across the warp only when the warp moves on to the next cell; the remainder of a cell-crossing move
// (19 % of the particles of the two-stream deck) goes through a warp-private queue, one segment each.
//
// Replaces reference em1d/particles.c:919-1074 (spec_advance: interpolate_fld :864-886, Boris push,
// dep_current_zamb :707-779, periodic / open / moving-window boundaries) and :791-850 (spec_sort).
// Arithmetic order follows the reference exactly (no contraction): one step is bit-identical.
#include "zdev_common.cuh"
#include <type_traits>
#include "pic2d_core.cuh"      // ltrim
#include "pic2d_packed.cuh"    // packed fp32 arithmetic, boris2
#include "zdev_tma.cuh"
#include <vector>
#include <cstring>

f3* zdev_grid1d_Epart(zdev_grid1d* g);
f3* zdev_grid1d_Bpart(zdev_grid1d* g);
f3* zdev_grid1d_J(zdev_grid1d* g);
int zdev_grid1d_nx(zdev_grid1d* g);

struct part1_aos { int ix; float x, ux, uy, uz; };           // host record (em1d/particles.h:29-35)
struct rec20 { float x, ux, uy, uz; int cell; };             // one particle as a value; cell = tile-local index
#define KEY1_EMPTY 0xffffu
// The sort key is the cell AND the quarter of the cell the particle sits in (SUB1 bins): inside a cell the
// particles are then ordered by position, so the ones about to cross a face (a contiguous slice at one end
// when the species drifts, 19 % per step in the two-stream deck) leave and arrive as a block and the gathers
// of the next step read runs of consecutive slots instead of isolated ones.
#ifndef SUB1
#define SUB1 4
#endif
__host__ __device__ __forceinline__ unsigned short key1_of(int cell, float x) {
	int b = (int) (x * SUB1);
	b = b < 0 ? 0 : (b > SUB1 - 1 ? SUB1 - 1 : b);
	return (unsigned short) (cell * SUB1 + b);
}
#define REC1_CHUNK_WORDS 128          // 4 rows x 32 slots: x, ux, uy, uz (the cell is in the key)

struct buf1d {
	float* rec;              // chunked records
	unsigned short* key;
	int* tag;                // optional
};
// migrants: one fixed segment per tile, [tile_off[t]/div, tile_off[t+1]/div), of reference-format records
// carrying UNWRAPPED global cell indices (boundary conditions are applied by k1_migrate)
struct mig1d { part1_aos* rec; int* tag; int* np; int div; };

__device__ __forceinline__ size_t rec1_word(int64_t slot) { return (size_t) (slot >> 5) * REC1_CHUNK_WORDS + (size_t) (slot & 31); }
// the record of a LIVE slot; its cell is carried by its sort key (key1_of)
__device__ __forceinline__ rec20 rec1_load(const float* __restrict__ rec, int64_t slot, unsigned short key) {
	const float* q = rec + rec1_word(slot);
	rec20 r; r.x = q[0]; r.ux = q[32]; r.uy = q[64]; r.uz = q[96]; r.cell = (int) key / SUB1;
	return r;
}
// (the caller stores the key: key1_of(r.cell, r.x))
__device__ __forceinline__ void rec1_store(float* __restrict__ rec, int64_t slot, const rec20& r) {
	float* q = rec + rec1_word(slot);
	q[0] = r.x; q[32] = r.ux; q[64] = r.uy; q[96] = r.uz;
}

struct ctl1d {
	double energy;
	unsigned long long np;
	unsigned int n_ovf;      // entries in the overflow list
	unsigned int flags;      // 1: some tile was full (particles parked in the overflow list), 2: a tile's migrants
	                         // segment overflowed, 8: the overflow list overflowed
};

struct zdev_spec1d {
	int nx, TX, ntiles, ppc_hint, track_ids;
	double slack;
	int64_t cap_total;
	int max_cap;
	buf1d p, q;
	int64_t* tile_off;
	int *tile_np, *tile_np_q;
	mig1d mig;                       // per-tile migrants segments
	part1_aos* ovf; int* ovf_tag;    // particles that found their destination tile full (global cell index):
	unsigned int ovf_cap;            //   the host grows the tiles and re-appends them before the next push
	ctl1d* ctl;
	int64_t np_host;
	int ids_valid;
	std::vector<int64_t>* h_off;
	std::vector<cudaEvent_t>* ev; int ev_next, ev_pending; double push_ms; int64_t push_launches;
};

static const int P1_THREADS = 256;
static const int P1_WARPS = P1_THREADS / 32;
static const int XQ1_CAP = 96;       // 31 left over + 64 new entries at most
static const int EV1_RING = 64;

static void buf_alloc(buf1d& b, int64_t n, int with_tag) {
	size_t nn = (size_t) (n > 0 ? n : 1);
	memset(&b, 0, sizeof b);
	if (nn & 31) nn = (nn + 31) & ~(size_t) 31;      // whole chunks
	ZDEV_CHECK(cudaMalloc(&b.rec, nn * 16));
	ZDEV_CHECK(cudaMalloc(&b.key, nn * 2));
	if (with_tag) ZDEV_CHECK(cudaMalloc(&b.tag, nn * 4));
}
static void buf_free(buf1d& b) { cudaFree(b.rec); cudaFree(b.key); cudaFree(b.tag); memset(&b, 0, sizeof b); }

extern "C" zdev_spec1d* zdev_spec1d_create(int nx, int ppc_hint, int track_ids) {
	zdev_require_init();
	{	// the packed-multiply addend (pic2d_packed.cuh)
		static bool negzero_set = false;
		if (!negzero_set) {
			const float2 nz = make_float2(-0.0f, -0.0f);
			ZDEV_CHECK(cudaMemcpyToSymbol(c_negzero2, &nz, sizeof nz));
			negzero_set = true;
		}
	}
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
	if (s->cap_total) { buf_free(s->p); buf_free(s->q); }
	cudaFree(s->mig.rec); cudaFree(s->mig.tag); cudaFree(s->mig.np);
	memset(&s->mig, 0, sizeof s->mig);
	cudaFree(s->ovf); cudaFree(s->ovf_tag); s->ovf = nullptr; s->ovf_tag = nullptr; s->ovf_cap = 0;
	s->cap_total = 0;
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

// migrants segments: 1/4 of every tile (two-stream decks move ~10 % of a 32-cell tile per step, a window
// shift a whole cell's worth on top)
static void mig1_alloc(zdev_spec1d* s) {
	cudaFree(s->mig.rec); cudaFree(s->mig.tag); cudaFree(s->mig.np);
	memset(&s->mig, 0, sizeof s->mig);
	s->mig.div = 4;
	ZDEV_CHECK(cudaMalloc(&s->mig.rec, (size_t) (s->cap_total / s->mig.div + 32) * sizeof(part1_aos)));
	if (s->track_ids) ZDEV_CHECK(cudaMalloc(&s->mig.tag, (size_t) (s->cap_total / s->mig.div + 32) * 4));
	ZDEV_CHECK(cudaMalloc(&s->mig.np, (size_t) s->ntiles * sizeof(int)));
	ZDEV_CHECK(cudaMemsetAsync(s->mig.np, 0, (size_t) s->ntiles * sizeof(int), zdev_strm));
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
	if (max_cap > 0xfff0) {
		fprintf(stderr, "(*error*) zpic-b200: %lld particles in one %d-cell tile exceed the shared-memory index "
		        "buffer; use smaller tiles (ZPIC_TILE_X1D)\n", (long long) max_cap, s->TX);
		exit(-1);
	}
	int64_t total = off[s->ntiles];
	free_particles(s);
	buf_alloc(s->p, total, s->track_ids);
	buf_alloc(s->q, total, s->track_ids);
	s->cap_total = total; s->max_cap = (int) max_cap;
	mig1_alloc(s);
	s->ovf_cap = (unsigned int) (total / 32 > (1 << 20) ? total / 32 : (1 << 20));
	ZDEV_CHECK(cudaMalloc(&s->ovf, (size_t) s->ovf_cap * sizeof(part1_aos)));
	if (s->track_ids) ZDEV_CHECK(cudaMalloc(&s->ovf_tag, (size_t) s->ovf_cap * 4));
	ZDEV_CHECK(cudaMemcpyAsync(s->tile_off, off.data(), (size_t) (s->ntiles + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, zdev_strm));
	ZDEV_CHECK(cudaMemsetAsync(s->tile_np, 0, (size_t) s->ntiles * sizeof(int), zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
}

// ------------------------------------------------------------------ host <-> device

__global__ void k1_count_tiles(const part1_aos* __restrict__ a, int64_t np, int TX, int* cnt) {
	int64_t k = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (k < np) atomicAdd(&cnt[a[k].ix / TX], 1);
}
// a particle whose destination tile is full: park it (global cell index) for the host to deal with
__device__ __forceinline__ void ovf1_push(ctl1d* ctl, part1_aos* ovf, int* ovf_tag, unsigned int cap, const part1_aos& r, int tag) {
	atomicOr(&ctl->flags, 1u);
	const unsigned int k = atomicAdd(&ctl->n_ovf, 1u);
	if (k < cap) { ovf[k] = r; if (ovf_tag) ovf_tag[k] = tag; }
	else atomicOr(&ctl->flags, 8u);
}

__global__ void k1_scatter(const part1_aos* __restrict__ a, int64_t np, int TX, buf1d p, const int64_t* __restrict__ off,
                           int* tile_np, ctl1d* ctl, int tag0, const int* __restrict__ tags,
                           part1_aos* ovf, int* ovf_tag, unsigned int ovf_cap) {
	int64_t k = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= np) return;
	part1_aos r = a[k];
	const int tag = tags ? tags[k] : tag0 + (int) k;
	int t = r.ix / TX;
	int slot = atomicAdd(&tile_np[t], 1);
	int64_t d = off[t] + slot;
	if (d >= off[t + 1]) { atomicSub(&tile_np[t], 1); ovf1_push(ctl, ovf, ovf_tag, ovf_cap, r, tag); return; }
	rec20 v = { r.x, r.ux, r.uy, r.uz, r.ix - t * TX };
	rec1_store(p.rec, d, v);
	p.key[d] = key1_of(v.cell, v.x);
	if (p.tag) p.tag[d] = tag;
}

static void spec1_resolve_overflow(zdev_spec1d* s);
static void check_flags(zdev_spec1d* s, unsigned int flags) {
	if (flags & 8u) {
		fprintf(stderr, "(*error*) zpic-b200: more than %u particles found their tile full in one step (%d-cell tiles); raise "
		        "ZPIC_TILE_SLACK (current %.2f) and rerun, aborting.\n", s->ovf_cap, s->TX, s->slack);
		exit(-1);
	}
	if (flags & 2u) {
		fprintf(stderr, "(*error*) zpic-b200: a tile's migrants segment overflowed (1/%d of the tile capacity); raise "
		        "ZPIC_TILE_SLACK (current %.2f) and rerun, aborting.\n", s->mig.div, s->slack);
		exit(-1);
	}
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
	if (np > 0) ZDEV_LAUNCH(k1_scatter, zdev_div_up(np, 256), 256, 0, d_aos, np, s->TX, s->p, s->tile_off, s->tile_np, s->ctl, 0,
	                        (const int*) nullptr, s->ovf, s->ovf_tag, s->ovf_cap);
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
	ZDEV_LAUNCH(k1_scatter, zdev_div_up(np, 256), 256, 0, d_aos, np, s->TX, s->p, s->tile_off, s->tile_np, s->ctl, (int) s->np_host,
	            (const int*) nullptr, s->ovf, s->ovf_tag, s->ovf_cap);
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	cudaFree(d_aos);
	s->np_host += np;
	spec1_resolve_overflow(s);
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
			rec20 v = rec1_load(p.rec, b + k, p.key[b + k]);
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

extern "C" int64_t zdev_spec1d_capacity(zdev_spec1d* s) { return s->cap_total; }

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

// ------------------------------------------------------------------ growing tiles (see zdev_spec2d.cu)

__global__ void k1_relayout(buf1d src, const int64_t* __restrict__ off_src, buf1d dst, const int64_t* __restrict__ off_dst,
                            const int* __restrict__ tile_np) {
	const int t = blockIdx.x, n = tile_np[t];
	const int64_t a = off_src[t], b = off_dst[t];
	for (int k = threadIdx.x; k < n; k += blockDim.x) {
		const unsigned short key = src.key[a + k];
		dst.key[b + k] = key;
		if (key != KEY1_EMPTY) {
			rec1_store(dst.rec, b + k, rec1_load(src.rec, a + k, key));
			if (src.tag) dst.tag[b + k] = src.tag[a + k];
		}
	}
}

// Particles that found their tile full were parked in the overflow list by the append / migrate kernels: give
// the tiles that need it more room (full ones 2x, those above 80 % 1.5x what they hold), copy the population
// to the new layout and append the parked particles.  Runs after every advance / append, so nothing ever
// misses a push.
static void spec1_resolve_overflow(zdev_spec1d* s) {
	ctl1d h;
	ZDEV_CHECK(cudaMemcpyAsync(&h, s->ctl, sizeof h, cudaMemcpyDeviceToHost, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	while (h.flags & 1u) {
		check_flags(s, h.flags & (2u | 8u));
		const int64_t n_ovf = h.n_ovf;
		std::vector<int> np_t(s->ntiles), ovf_t(s->ntiles, 0);
		int* d_cnt; ZDEV_CHECK(cudaMalloc(&d_cnt, (size_t) s->ntiles * sizeof(int)));
		ZDEV_CHECK(cudaMemsetAsync(d_cnt, 0, (size_t) s->ntiles * sizeof(int), zdev_strm));
		ZDEV_LAUNCH(k1_count_tiles, zdev_div_up(n_ovf, 256), 256, 0, s->ovf, n_ovf, s->TX, d_cnt);
		ZDEV_CHECK(cudaMemcpyAsync(ovf_t.data(), d_cnt, (size_t) s->ntiles * sizeof(int), cudaMemcpyDeviceToHost, zdev_strm));
		ZDEV_CHECK(cudaMemcpyAsync(np_t.data(), s->tile_np, (size_t) s->ntiles * sizeof(int), cudaMemcpyDeviceToHost, zdev_strm));
		ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
		cudaFree(d_cnt);
		std::vector<int64_t> off_new(s->ntiles + 1, 0);
		const std::vector<int64_t>& off = *s->h_off;
		int64_t max_cap = 0;
		for (int t = 0; t < s->ntiles; t++) {
			int64_t cap = off[t + 1] - off[t];
			const int64_t need = (int64_t) np_t[t] + ovf_t[t];
			if (ovf_t[t] > 0 || need > cap - cap / 10) {          // (within 10 %: at the 1.25 slack of large runs every tile sits at 80 %)
				int64_t grown = (ovf_t[t] > 0 ? 2 * need : need + need / 2) + 64;
				grown = (grown + 31) & ~(int64_t) 31;
				if (grown > cap) cap = grown;
			}
			off_new[t + 1] = off_new[t] + cap;
			if (cap > max_cap) max_cap = cap;
		}
		if (max_cap > 0xfff0) {
			fprintf(stderr, "(*error*) zpic-b200: %lld particles in one %d-cell tile exceed the shared-memory index "
			        "buffer; use smaller tiles (ZPIC_TILE_X1D)\n", (long long) max_cap, s->TX);
			exit(-1);
		}
		const int64_t total = off_new[s->ntiles];
		part1_aos* d_wait; int* d_wait_tag = nullptr;
		ZDEV_CHECK(cudaMalloc(&d_wait, (size_t) n_ovf * sizeof(part1_aos)));
		ZDEV_CHECK(cudaMemcpyAsync(d_wait, s->ovf, (size_t) n_ovf * sizeof(part1_aos), cudaMemcpyDeviceToDevice, zdev_strm));
		if (s->ovf_tag) {
			ZDEV_CHECK(cudaMalloc(&d_wait_tag, (size_t) n_ovf * 4));
			ZDEV_CHECK(cudaMemcpyAsync(d_wait_tag, s->ovf_tag, (size_t) n_ovf * 4, cudaMemcpyDeviceToDevice, zdev_strm));
		}
		ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
		buf_free(s->q);                                   // scratch between steps
		buf1d pn;
		buf_alloc(pn, total, s->track_ids);
		int64_t* d_off_new; ZDEV_CHECK(cudaMalloc(&d_off_new, (size_t) (s->ntiles + 1) * sizeof(int64_t)));
		ZDEV_CHECK(cudaMemcpyAsync(d_off_new, off_new.data(), (size_t) (s->ntiles + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, zdev_strm));
		ZDEV_LAUNCH(k1_relayout, s->ntiles, 256, 0, s->p, s->tile_off, pn, d_off_new, s->tile_np);
		ZDEV_CHECK(cudaMemcpyAsync(s->tile_off, d_off_new, (size_t) (s->ntiles + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, zdev_strm));
		ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
		cudaFree(d_off_new);
		buf_free(s->p);
		s->p = pn;
		buf_alloc(s->q, total, s->track_ids);
		*s->h_off = off_new;
		s->cap_total = total;
		s->max_cap = (int) max_cap;
		mig1_alloc(s);
		ZDEV_CHECK(cudaMemsetAsync(&s->ctl->n_ovf, 0, 2 * sizeof(unsigned int), zdev_strm));    // n_ovf, flags
		ZDEV_LAUNCH(k1_scatter, zdev_div_up(n_ovf, 256), 256, 0, d_wait, n_ovf, s->TX, s->p, s->tile_off, s->tile_np, s->ctl, 0,
		            (const int*) d_wait_tag, s->ovf, s->ovf_tag, s->ovf_cap);
		ZDEV_CHECK(cudaMemcpyAsync(&h, s->ctl, sizeof h, cudaMemcpyDeviceToHost, zdev_strm));
		ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
		cudaFree(d_wait); cudaFree(d_wait_tag);
	}
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
		rec1_store(p.rec, base + k, v); p.key[base + k] = key1_of(lc, v.x);
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

// ---- the same population as the reference's host injector, on the reference random stream (zdev_refrng.cu; the
// 1-D twin of k_inject_lattice in zdev_spec2d.cu).  One thread per cell of [i0, i1): the in-cell positions
// klo <= k < khi of cell i carry plasma (STEP / SLAB clip them, em1d/particles.c:245-262), pre[] = particles in
// the cells before i; th[] holds the three thermal components of the band's particles in injection order (null:
// a cold plasma).  Cell means are summed in particle order like spec_set_u does (em1d/particles.c:88-130).
__global__ void k1_inject_lattice(buf1d p, const int64_t* __restrict__ off, int* tile_np, int nx, int TX, int ppc, f3 ufl,
                                  const int* __restrict__ klo, const int* __restrict__ khi, const int64_t* __restrict__ pre,
                                  int i0, int i1, const float* __restrict__ th) {
	const int cell = i0 + blockIdx.x * blockDim.x + threadIdx.x;
	if (cell >= i1) return;
	const int t = cell / TX, lc = cell - t * TX, gt0 = t * TX;
	const int cxw = (t + 1) * TX <= nx ? TX : nx - t * TX;
	if (lc == 0) tile_np[t] = (int) (pre[gt0 + cxw] - pre[gt0]);
	const int lo = klo[cell], cnt = khi[cell] - lo;
	if (cnt <= 0) return;
	const int64_t base = off[t] + (pre[cell] - pre[gt0]);
	const float* q = th ? th + 3 * (pre[cell] - pre[i0]) : nullptr;
	float sx = 0, sy = 0, sz = 0;
	if (q) {
		for (int k = 0; k < cnt; k++) { sx += q[3 * k]; sy += q[3 * k + 1]; sz += q[3 * k + 2]; }
		const float norm = 1.0f / cnt;
		sx *= norm; sy *= norm; sz *= norm;
	}
	for (int k = 0; k < cnt; k++) {
		const float a = q ? q[3 * k] : 0.0f, b = q ? q[3 * k + 1] : 0.0f, c = q ? q[3 * k + 2] : 0.0f;
		rec20 v = { (float) ((lo + k + 0.5) / ppc), a + (ufl.x - sx), b + (ufl.y - sy), c + (ufl.z - sz), lc };
		rec1_store(p.rec, base + k, v); p.key[base + k] = key1_of(lc, v.x);
		if (p.tag) p.tag[base + k] = (int) (pre[cell] + k);
	}
}

extern "C" int zdev_spec1d_inject_lattice(zdev_spec1d* s, int ppc, const float ufl[3], const float uth[3],
                                          const int* klo, const int* khi,
                                          uint32_t* z, uint32_t* w, int* have_spare, double* spare) {
	{ uint32_t zz = *z, ww = *w; if (zdev_ref_jump(&zz, &ww, 0)) return 1; }
	const int nx = s->nx;
	std::vector<int64_t> pre(nx + 1, 0);
	for (int i = 0; i < nx; i++) pre[i + 1] = pre[i] + std::max(khi[i] - klo[i], 0);
	const int64_t total = pre[nx];
	std::vector<int> cnt(s->ntiles);
	for (int t = 0; t < s->ntiles; t++) {
		const int cx = (t + 1) * s->TX <= nx ? s->TX : nx - t * s->TX;
		cnt[t] = (int) (pre[t * s->TX + cx] - pre[t * s->TX]);
	}
	layout(s, cnt, total);
	ZDEV_CHECK(cudaMemsetAsync(s->tile_np, 0, (size_t) s->ntiles * sizeof(int), zdev_strm));
	int *d_lo, *d_hi; int64_t* d_pre;
	ZDEV_CHECK(cudaMalloc(&d_lo, (size_t) nx * sizeof(int)));
	ZDEV_CHECK(cudaMalloc(&d_hi, (size_t) nx * sizeof(int)));
	ZDEV_CHECK(cudaMalloc(&d_pre, (size_t) (nx + 1) * sizeof(int64_t)));
	ZDEV_CHECK(cudaMemcpyAsync(d_lo, klo, (size_t) nx * sizeof(int), cudaMemcpyHostToDevice, zdev_strm));
	ZDEV_CHECK(cudaMemcpyAsync(d_hi, khi, (size_t) nx * sizeof(int), cudaMemcpyHostToDevice, zdev_strm));
	ZDEV_CHECK(cudaMemcpyAsync(d_pre, pre.data(), (size_t) (nx + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, zdev_strm));
	const f3 fl = {ufl[0], ufl[1], ufl[2]};
	const bool cold = uth[0] == 0.0f && uth[1] == 0.0f && uth[2] == 0.0f;
	if (cold || total == 0) {
		zdev_ref_normals(z, w, have_spare, spare, 3 * total, uth, nullptr);
		if (total > 0)
			ZDEV_LAUNCH(k1_inject_lattice, zdev_div_up(nx, 128), 128, 0, s->p, s->tile_off, s->tile_np, nx, s->TX, ppc, fl,
			            d_lo, d_hi, d_pre, 0, nx, (const float*) nullptr);
	} else {
		// bands of cells holding at most 2^26 particles: their thermal components are generated, then placed
		const int64_t band = (int64_t) 1 << 26;
		float* d_th;
		ZDEV_CHECK(cudaMalloc(&d_th, (size_t) std::min<int64_t>(total, band + ppc) * 3 * sizeof(float)));
		for (int i0 = 0; i0 < nx; ) {
			int i1 = i0 + 1;
			while (i1 < nx && pre[i1 + 1] - pre[i0] <= band) i1++;
			zdev_ref_normals(z, w, have_spare, spare, 3 * (pre[i1] - pre[i0]), uth, d_th);
			ZDEV_LAUNCH(k1_inject_lattice, zdev_div_up(i1 - i0, 128), 128, 0, s->p, s->tile_off, s->tile_np, nx, s->TX, ppc, fl,
			            d_lo, d_hi, d_pre, i0, i1, (const float*) d_th);
			i0 = i1;
		}
		ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
		ZDEV_CHECK(cudaFree(d_th));
	}
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	ZDEV_CHECK(cudaFree(d_lo)); ZDEV_CHECK(cudaFree(d_hi)); ZDEV_CHECK(cudaFree(d_pre));
	s->np_host = total;
	s->ids_valid = s->track_ids && total < 0x7fffffff;
	return 0;
}

// ------------------------------------------------------------------ the push

// one queued remainder of a cell-crossing move (a single in-cell segment, already in the frame of the cell
// behind the face): 5 words, odd stride = conflict-free shared-memory stores
struct xq1_entry { int lx; float x0, dx, qvy, qvz; };

// The five current contributions of an in-cell segment x0 -> x1 (em1d/particles.c:763-777):
// Jx[ix]; Jy[ix], Jy[ix+1]; Jz[ix], Jz[ix+1], with qvy, qvz already halved and scaled to the segment:
//   S0x0 + S1x0 + (S0x0 - S1x0)/2 = 1.5 (1-x0) + 0.5 (1-x1) = 2 - (1.5 x0 + 0.5 x1)
// J is compared by tolerance (summation order), so this uses fused multiply-adds.
__device__ __forceinline__ void seg1_weights(float x0, float x1, float dx, float qvy, float qvz, float qnx, float w[5]) {
	const float a1 = __fmaf_rn(1.5f, x0, 0.5f * x1), a0 = 2.0f - a1;
	w[0] = qnx * dx;
	w[1] = qvy * a0; w[2] = qvy * a1;
	w[3] = qvz * a0; w[4] = qvz * a1;
}
// scatter into the global J grid (L2 reductions); c = cell of the segment
__device__ __forceinline__ void red1(f3* __restrict__ c, const float w[5]) {
	atomicAdd(&c[0].x, w[0]);
	atomicAdd(&c[0].y, w[1]); atomicAdd(&c[1].y, w[2]);
	atomicAdd(&c[0].z, w[3]); atomicAdd(&c[1].z, w[4]);
}
// Deposit up to 32 queued remainders, one per lane.  The queue is filled in sorted cell order, so entries with
// the same destination cell sit next to each other (all of them, when the species drifts one way): combine
// each run with a segmented warp scan and issue the five L2 reductions once per run - 32 lanes hammering the
// same five addresses serialise in the L2 atomic unit.
__device__ __forceinline__ void drain1(const xq1_entry* q, int n, int lane, f3* __restrict__ J0, float qnx) {
	const bool act = lane < n;
	float w[5];
	int key = 0x7fffffff;
	if (act) {
		const xq1_entry e = q[lane];
		seg1_weights(e.x0, e.x0 + e.dx, e.dx, e.qvy, e.qvz, qnx, w);
		key = e.lx;
	} else {
		#pragma unroll
		for (int k = 0; k < 5; k++) w[k] = 0.0f;
	}
	const int prev = __shfl_up_sync(0xffffffffu, key, 1);
	const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || key != prev);
	const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const bool take = (lane - d) >= start;
		#pragma unroll
		for (int k = 0; k < 5; k++) { float u = __shfl_up_sync(0xffffffffu, w[k], d); if (take) w[k] += u; }
	}
	const bool tail = (lane == 31) || ((heads >> (lane + 1)) & 1u);
	if (tail && act) red1(J0 + key, w);
}

// Sum acc[0..4] over the 32 lanes with the transposed butterfly of the 2-D kernel (slots 5..7 are zero) and
// add the totals to cell `cell` of the tile: lane 4*k ends up with contribution k.
__device__ __forceinline__ void flush_cell1(const float acc[5], int cell, int lane, f3* __restrict__ J0) {
	const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4;
	// stage 1 (xor 16): pairs (0,4) (1,5) (2,6) (3,7) with 5..7 = 0
	float v4[4];
	{
		float send = b16 ? acc[0] : acc[4], keep = b16 ? acc[4] : acc[0];
		v4[0] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
		#pragma unroll
		for (int q = 1; q < 4; q++) {
			float s2 = b16 ? acc[q] : 0.0f, k2 = b16 ? 0.0f : acc[q];
			v4[q] = k2 + __shfl_xor_sync(0xffffffffu, s2, 16);
		}
	}
	float v2[2], v1;
	#pragma unroll
	for (int q = 0; q < 2; q++) {
		float send = b8 ? v4[q] : v4[q + 2], keep = b8 ? v4[q + 2] : v4[q];
		v2[q] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
	}
	{
		float send = b4 ? v2[0] : v2[1], keep = b4 ? v2[1] : v2[0];
		v1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
	}
	v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
	v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
	const int k = lane >> 2;
	if ((lane & 3) == 0 && k < 5) {
		const int comp = (k == 0) ? 0 : ((k < 3) ? 1 : 2);
		const int right = (k == 2) | (k == 4);
		atomicAdd(reinterpret_cast<float*>(J0 + cell + right) + comp, v1);
	}
}

struct pair1_rec { f2 x, ux, uy, uz; int ca, cb, ta, tb; };


// dynamic shared memory of k_push1d: [keys during the sort | field pairs + queues afterwards][perm][raw planes]
static size_t push1_smem_front(int TX, int max_cap) {
	// + 128 keys: the sort walks 64 * S >= n keys and the ones past n are padded with KEY1_EMPTY in place
	size_t late = (size_t) 6 * (TX + 2) * 8 + (size_t) P1_WARPS * XQ1_CAP * sizeof(xq1_entry), keys = ((size_t) (max_cap + 128) * 2 + 15) & ~(size_t) 15;
	late = (late + 15) & ~(size_t) 15;
	return late > keys ? late : keys;
}
// perm[] = 32-bit (key << 16 | slot) entries; the holes are sorted too (behind the live entries)
static size_t push1_smem_perm(int max_cap) { return (((size_t) (max_cap + 128) * 4 + 15) & ~(size_t) 15); }
static size_t push1_smem_bytes(int TX, int max_cap) {
	return push1_smem_front(TX, max_cap) + push1_smem_perm(max_cap) + (size_t) 6 * (TX + 2) * 4;
}

template <bool TAGS>
__global__ void __launch_bounds__(P1_THREADS, 2)
k_push1d(buf1d A, buf1d Bo, const int64_t* __restrict__ tile_off, const int* __restrict__ tile_np, int* __restrict__ tile_np_out,
         mig1d mig, ctl1d* __restrict__ ctl,
         const f3* __restrict__ E, const f3* __restrict__ B, f3* __restrict__ J, int nx, int TX,
         zdev_push1d_params prm, unsigned smem_front, unsigned smem_perm
         ) {
	extern __shared__ __align__(16) unsigned char s_dyn[];
	const int PL = TX + 2;
	f2* const s_f2 = reinterpret_cast<f2*>(s_dyn);                     // 6 planes of (F[k], F[k+1])
	xq1_entry* const s_xq = reinterpret_cast<xq1_entry*>(s_dyn + 6 * PL * 8);
	const unsigned short* const s_key = reinterpret_cast<const unsigned short*>(s_dyn);
	unsigned* const s_perm = reinterpret_cast<unsigned*>(s_dyn + smem_front);
	float* const s_raw = reinterpret_cast<float*>(s_dyn + smem_front + smem_perm);
	__shared__ int s_cnt[512 * SUB1 + 1], s_cur[512 * SUB1 + 1];       // bin NK: the holes
	const int NK = TX * SUB1;                            // sort keys of the tile
	__shared__ int s_wsum[P1_WARPS];
	__shared__ int s_nmig, s_done;
	__shared__ __align__(8) unsigned long long s_bar;

	const int t = blockIdx.x, x0 = t * TX, cx = min(TX, nx - x0);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int n = tile_np[t];
	const int64_t base = tile_off[t];

	if (threadIdx.x == 0) {
		s_nmig = 0; s_done = 0;
		mbar_init(&s_bar, 1);
		if (n > 0) bulk_load(s_dyn, A.key + base, (unsigned) ((n * 2 + 15) & ~15), &s_bar);
	}
	for (int k = threadIdx.x; k < cx + 2; k += P1_THREADS) {        // cells x0-1 .. x0+cx
		f3 e = E[x0 + k], b = B[x0 + k];
		s_raw[k] = e.x; s_raw[k + PL] = e.y; s_raw[k + 2 * PL] = e.z;
		s_raw[k + 3 * PL] = b.x; s_raw[k + 4 * PL] = b.y; s_raw[k + 5 * PL] = b.z;
	}
	for (int k = threadIdx.x; k <= NK; k += P1_THREADS) s_cnt[k] = 0;
	__syncthreads();
	if (n > 0) mbar_wait(&s_bar, 0);

	// ---- phase A: counting sort of slot indices by cell (see zdev_spec2d.cu: every thread walks its own
	//      stretch of consecutive keys, an odd number of words long, so that the lanes of a warp are in
	//      different cells and banks at any moment)
	// Lane l owns the words [l*S, (l+1)*S) of the key array (S odd: the lanes of a warp read distinct banks
	// and sit S*2 keys apart, i.e. in different cells as long as a cell holds fewer particles than that);
	// the warps split every lane's stretch into WARPS consecutive pieces.
	// Branch-free (as in zdev_spec2d.cu): the keys past n (up to the 64 * S the lanes cover) are padded with
	// KEY1_EMPTY, and a hole counts as key NK - it gets a counter and a stretch of perm[] behind the live entries.
	const int S = ((((n + 1) >> 1) + 31) >> 5) | 1;
	{
		unsigned short* const kw = const_cast<unsigned short*>(s_key);
		for (int k = n + threadIdx.x; k < 64 * S; k += P1_THREADS) kw[k] = (unsigned short) KEY1_EMPTY;
		__syncthreads();
	}
	const int ws = (S + P1_WARPS - 1) / P1_WARPS;
	const int w0 = lane * S + warp * ws, wn = min(ws, S - warp * ws);
	const unsigned* const s_key2 = reinterpret_cast<const unsigned*>(s_key) + w0;
	#pragma unroll 4
	for (int j = 0; j < wn; j++) {
		const unsigned two = s_key2[j];
		atomicAdd(&s_cnt[min(two & 0xffffu, (unsigned) NK)], 1);
		atomicAdd(&s_cnt[min(two >> 16, (unsigned) NK)], 1);
	}
	__syncthreads();
	int nlive;
	{	// exclusive scan over the NK <= 2048 counters, 2 * SUB1 consecutive ones per thread
		constexpr int PER = 2 * SUB1;
		const int i0 = PER * threadIdx.x;
		int c[PER], v = 0;
		#pragma unroll
		for (int k = 0; k < PER; k++) { c[k] = (i0 + k < NK) ? s_cnt[i0 + k] : 0; v += c[k]; }
		int incl = v;
		for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += u; }
		if (lane == 31) s_wsum[warp] = incl;
		__syncthreads();
		int woff = 0, tot = 0;
		#pragma unroll
		for (int w = 0; w < P1_WARPS; w++) { int cw = s_wsum[w]; woff += (w < warp) ? cw : 0; tot += cw; }
		int run = woff + incl - v;
		#pragma unroll
		for (int k = 0; k < PER; k++) { if (i0 + k < NK) s_cur[i0 + k] = run; run += c[k]; }
		if (threadIdx.x == 0) s_cur[NK] = tot;               // the holes start behind the live entries
		nlive = tot;
		__syncthreads();
	}
	#pragma unroll 4
	for (int j = 0; j < wn; j++) {
		const unsigned two = s_key2[j];
		const unsigned c0 = min(two & 0xffffu, (unsigned) NK), c1 = min(two >> 16, (unsigned) NK);
		const unsigned i = 2u * (unsigned) (w0 + j);
		s_perm[atomicAdd(&s_cur[c0], 1)] = (c0 << 16) | i;
		s_perm[atomicAdd(&s_cur[c1], 1)] = (c1 << 16) | (i + 1u);
	}
	__syncthreads();                                    // the keys are dead: their bytes become field pairs + queues
	// ---- the fields as (F[k], F[k+1]) pairs: one LDS.64 per component and particle
	for (int k = threadIdx.x; k < 6 * PL; k += P1_THREADS) {
		const int pl = k / PL, o = k - pl * PL;
		const int o1 = (o + 1 < PL) ? k + 1 : k;
		s_f2[k] = mk2(s_raw[k], s_raw[o1]);
	}
	__syncthreads();

	// ---- phase B: every warp streams a contiguous range of the sorted particles, 64 per iteration (lane l
	//      owns the particles l and l+32 of the block); no block barriers from here on
	xq1_entry* const xq = s_xq + warp * XQ1_CAP;
	const unsigned lt = (1u << lane) - 1u;
	int nxq = 0;
	float energy = 0.0f;
	f3* const J0 = J + x0 + 1;                                   // cell x0
	const f2* const F2 = s_f2 + 1;                               // plane index 0 is cell -1 of the tile
	const float* const Arec = A.rec + (size_t) (base >> 5) * REC1_CHUNK_WORDS;
	float* const Brec = Bo.rec + (size_t) (base >> 5) * REC1_CHUNK_WORDS;
	const int chunk = ((nlive + P1_WARPS * 64 - 1) / (P1_WARPS * 64)) * 64;
	const int pbeg = warp * chunk, pend = min(pbeg + chunk, nlive);
	const int mig_cap = (int) ((tile_off[t + 1] - base) / mig.div);
	const int64_t mig_base = base / mig.div;

	float acc[5];
	#pragma unroll
	for (int q = 0; q < 5; q++) acc[q] = 0.0f;
	int cur = -1;

	// current of 32 consecutive sorted particles (one per lane): accumulate per lane while the warp stays in
	// one cell, reduce when it moves on; several cells per 32 lanes: segmented scan
	auto deposit32 = [&](int key, float (&w)[5], bool act) {
		const int prev = __shfl_up_sync(0xffffffffu, key, 1);
		const unsigned heads = __ballot_sync(0xffffffffu, (lane == 0) ? (key != cur) : (key != prev));
		if (heads == 0u) {
			#pragma unroll
			for (int q = 0; q < 5; q++) acc[q] += w[q];
			return;
		}
		const int b = __ffs(heads) - 1;
		const bool lo = lane < b;
		const float mlo = lo ? 1.0f : 0.0f, mhi = lo ? 0.0f : 1.0f;
		if (cur >= 0) {
			#pragma unroll
			for (int q = 0; q < 5; q++) acc[q] = __fmaf_rn(w[q], mlo, acc[q]);
			flush_cell1(acc, cur, lane, J0);
		}
		if ((heads & (heads - 1u)) == 0u) {
			#pragma unroll
			for (int q = 0; q < 5; q++) acc[q] = w[q] * mhi;
			cur = __shfl_sync(0xffffffffu, key, 31);
			if (cur >= TX) cur = -1;
		} else {
			#pragma unroll
			for (int q = 0; q < 5; q++) w[q] *= mhi;
			const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
			#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const bool take = (lane - d) >= start;
				#pragma unroll
				for (int q = 0; q < 5; q++) { float u = __shfl_up_sync(0xffffffffu, w[q], d); if (take) w[q] += u; }
			}
			const bool tail = (lane == 31) || ((heads >> (lane + 1)) & 1u);
			if (tail && act && !lo) red1(J0 + key, w);
			#pragma unroll
			for (int q = 0; q < 5; q++) acc[q] = 0.0f;
			cur = -1;
		}
	};

	auto load_pair = [&](int pa, pair1_rec& r) {
		const unsigned va = s_perm[pa < pend ? pa : pbeg], vb = s_perm[pa + 32 < pend ? pa + 32 : pbeg];
		const int ia = va & 0xffffu, ib = vb & 0xffffu;
		const float* qa = Arec + (ia >> 5) * REC1_CHUNK_WORDS + (ia & 31);
		const float* qb = Arec + (ib >> 5) * REC1_CHUNK_WORDS + (ib & 31);
		r.x = mk2(qa[0], qb[0]); r.ux = mk2(qa[32], qb[32]); r.uy = mk2(qa[64], qb[64]); r.uz = mk2(qa[96], qb[96]);
		r.ca = (int) (va >> 16) / SUB1; r.cb = (int) (vb >> 16) / SUB1;             // the sort key carries the cell
		r.ta = r.tb = 0;
		if (TAGS) { r.ta = A.tag[base + ia]; r.tb = A.tag[base + ib]; }
	};

	pair1_rec nv;
	if (pbeg < pend) load_pair(pbeg + lane, nv);

	// one iteration = 64 particles; FULL: all 64 exist (every iteration but the last of a warp's range), so the
	// activity masks are compile-time true and the predicates, selects and store guards they feed disappear
	auto advance64 = [&](const int p0, auto full_tag) {
		constexpr bool FULL = decltype(full_tag)::value;
		const int pa = p0 + lane, pb = pa + 32;
		const bool actA = FULL || pa < pend, actB = FULL || pb < pend;
		const pair1_rec v = nv;
		if (p0 + 64 < pend) load_pair(pa + 64, nv);        // software pipeline: next iteration's records

		const int lxa = v.ca, lxb = v.cb;
		f2 x = v.x, ux = v.ux, uy = v.uy, uz = v.uz;
		f2 dx, qvy, qvz;
		{
			// interpolate_fld (em1d/particles.c:864-886): F[i]*(1-w) + F[i+1]*w, the two products as one FFMA2
			f2 Ex, Ey, Ez, Bx, By, Bz;
			#pragma unroll
			for (int h2 = 0; h2 < 2; h2++) {
				const float xx = h2 ? x.y : x.x;
				const int lx = h2 ? lxb : lxa;
				const int h = (xx < 0.5f) ? 1 : 0;
				const float w1h = xx + (h ? 0.5f : -0.5f);
				const f2 wn = mk2(1.0f - xx, xx), wh = mk2(1.0f - w1h, w1h);
				const int i = lx, ih = lx - h;
				f2 m;
				float ex, ey, ez, bx, by, bz;
				m = mul2(F2[ih], wh);          ex = m.x + m.y;
				m = mul2(F2[i + PL], wn);      ey = m.x + m.y;
				m = mul2(F2[i + 2 * PL], wn);  ez = m.x + m.y;
				m = mul2(F2[i + 3 * PL], wn);  bx = m.x + m.y;
				m = mul2(F2[ih + 4 * PL], wh); by = m.x + m.y;
				m = mul2(F2[ih + 5 * PL], wh); bz = m.x + m.y;
				if (h2) { Ex.y = ex; Ey.y = ey; Ez.y = ez; Bx.y = bx; By.y = by; Bz.y = bz; }
				else    { Ex.x = ex; Ey.x = ey; Ez.x = ez; Bx.x = bx; By.x = by; Bz.x = bz; }
			}
			// Boris (em1d/particles.c:953-1000): same rotation as 2D; energy term u2/(1+gamma)
			const f2 en = boris2(Ex, Ey, Ez, Bx, By, Bz, prm.tem, ux, uy, uz);
			energy += (actA ? en.x : 0.0f) + (actB ? en.y : 0.0f);
		}
		{
			const f2 usq = add2(add2(add2(bc2(1.0f), mul2(ux, ux)), mul2(uy, uy)), mul2(uz, uz));
			const f2 rg = div_exact2(bc2(1.0f), sqrt_exact2(usq));
			dx = mul2(mul2(rg, prm.dt_dx), ux);
			qvy = mul2(mul2(uy, prm.q), rg);
			qvz = mul2(mul2(uz, prm.q), rg);
		}
		const f2 x1 = add2(x, dx);
		const int dia = ltrim(x1.x), dib = ltrim(x1.y);
		const bool xa = actA && dia != 0, xb = actB && dib != 0;

		// first in-cell piece of every move through the lanes (the whole move when no face is crossed)
		const f2 fx = mk2(dia > 0 ? 1.0f : 0.0f, dib > 0 ? 1.0f : 0.0f);
		f2 t1 = __fmul2_rn(sub2(fx, x), rcp_approx2(dx));
		t1.x = dia ? fminf(fmaxf(t1.x, 0.0f), 1.0f) : 1.0f;
		t1.y = dib ? fminf(fmaxf(t1.y, 0.0f), 1.0f) : 1.0f;
		const f2 dx0 = __fmul2_rn(dx, t1);
		f2 xe = add2(x, dx0);
		xe.x = dia ? fx.x : xe.x; xe.y = dib ? fx.y : xe.y;
		{
			const f2 kh = mk2(actA ? 0.5f : 0.0f, actB ? 0.5f : 0.0f);
			const f2 qy = __fmul2_rn(__fmul2_rn(qvy, kh), t1), qz = __fmul2_rn(__fmul2_rn(qvz, kh), t1);
			const f2 a1 = fma2(bc2(1.5f), x, __fmul2_rn(bc2(0.5f), xe)), a0 = sub2(bc2(2.0f), a1);
			const f2 w0 = __fmul2_rn(__fmul2_rn(dx0, mk2(actA ? 1.0f : 0.0f, actB ? 1.0f : 0.0f)), bc2(prm.qnx));
			const f2 w1 = __fmul2_rn(qy, a0), w2 = __fmul2_rn(qy, a1), w3 = __fmul2_rn(qz, a0), w4 = __fmul2_rn(qz, a1);
			float w[5];
			w[0] = w0.x; w[1] = w1.x; w[2] = w2.x; w[3] = w3.x; w[4] = w4.x;
			deposit32(actA ? lxa : 0x7fffffff, w, actA);
			w[0] = w0.y; w[1] = w1.y; w[2] = w2.y; w[3] = w3.y; w[4] = w4.y;
			deposit32(actB ? lxb : 0x7fffffff, w, actB);
		}

		// --- queue of the remainders (always a single segment in 1D); drain 32 at a time
		{
			const unsigned ma = __ballot_sync(0xffffffffu, xa), mb = __ballot_sync(0xffffffffu, xb);
			if (ma | mb) {
				if (xa) {
					const float rem = 1.0f - t1.x;
					xq1_entry e; e.lx = lxa + dia; e.x0 = 1.0f - fx.x; e.dx = dx.x * rem;
					e.qvy = qvy.x * 0.5f * rem; e.qvz = qvz.x * 0.5f * rem;
					xq[nxq + __popc(ma & lt)] = e;
				}
				nxq += __popc(ma);
				if (xb) {
					const float rem = 1.0f - t1.y;
					xq1_entry e; e.lx = lxb + dib; e.x0 = 1.0f - fx.y; e.dx = dx.y * rem;
					e.qvy = qvy.y * 0.5f * rem; e.qvz = qvz.y * 0.5f * rem;
					xq[nxq + __popc(mb & lt)] = e;
				}
				nxq += __popc(mb);
				__syncwarp();
				while (nxq >= 32) { drain1(xq + nxq - 32, 32, lane, J0, prm.qnx); nxq -= 32; }
				__syncwarp();
			}
		}

		// --- new positions; survivors go to their sorted slot in B, leavers to the tile's migrants segment
		const f2 xn = sub2(x1, mk2((float) dia, (float) dib));
		const int nlxa = lxa + dia - prm.shift_window, nlxb = lxb + dib - prm.shift_window;
		const bool sta = actA && (unsigned) nlxa < (unsigned) cx, stb = actB && (unsigned) nlxb < (unsigned) cx;
		if (actA) {
			float* qd = Brec + (pa >> 5) * REC1_CHUNK_WORDS + (pa & 31);
			qd[0] = xn.x; qd[32] = ux.x; qd[64] = uy.x; qd[96] = uz.x;
			Bo.key[base + pa] = sta ? key1_of(nlxa, xn.x) : (unsigned short) KEY1_EMPTY;
			if (TAGS) Bo.tag[base + pa] = v.ta;
		}
		if (actB) {
			float* qd = Brec + (pb >> 5) * REC1_CHUNK_WORDS + (pb & 31);
			qd[0] = xn.y; qd[32] = ux.y; qd[64] = uy.y; qd[96] = uz.y;
			Bo.key[base + pb] = stb ? key1_of(nlxb, xn.y) : (unsigned short) KEY1_EMPTY;
			if (TAGS) Bo.tag[base + pb] = v.tb;
		}
		{
			const bool la = actA && !sta, lb = actB && !stb;
			const unsigned ma = __ballot_sync(0xffffffffu, la), mb = __ballot_sync(0xffffffffu, lb);
			if (ma | mb) {
				int slot = 0;
				if (lane == 0) slot = atomicAdd(&s_nmig, __popc(ma) + __popc(mb));
				slot = __shfl_sync(0xffffffffu, slot, 0);
				if (la) {
					const int d = slot + __popc(ma & lt);
					if (d < mig_cap) {
						part1_aos r = { x0 + nlxa, xn.x, ux.x, uy.x, uz.x };
						mig.rec[mig_base + d] = r;
						if (TAGS) mig.tag[mig_base + d] = v.ta;
					}
				}
				if (lb) {
					const int d = slot + __popc(ma) + __popc(mb & lt);
					if (d < mig_cap) {
						part1_aos r = { x0 + nlxb, xn.y, ux.y, uy.y, uz.y };
						mig.rec[mig_base + d] = r;
						if (TAGS) mig.tag[mig_base + d] = v.tb;
					}
				}
			}
		}
	};
	{
		int p0 = pbeg;
		for (; p0 + 64 <= pend; p0 += 64) advance64(p0, std::true_type());
		if (p0 < pend) advance64(p0, std::false_type());
	}
	if (cur >= 0) flush_cell1(acc, cur, lane, J0);
	if (nxq) drain1(xq, nxq, lane, J0, prm.qnx);

	double e = (double) energy;
	for (int o = 16; o > 0; o >>= 1) e += __shfl_down_sync(0xffffffffu, e, o);
	if (lane == 0 && nlive > 0) atomicAdd(&ctl->energy, e);
	if (lane == 0) {
		__threadfence_block();
		if (atomicAdd(&s_done, 1) == P1_WARPS - 1) {
			const int nm = atomicAdd(&s_nmig, 0);
			if (nm > mig_cap) atomicOr(&ctl->flags, 2u);
			mig.np[t] = min(nm, mig_cap);
			tile_np_out[t] = nlive;
		}
	}
}

// Boundary conditions for the particles that left their tile (em1d/particles.c:1044-1060: absorbing under a
// moving window or open boundaries, else periodic), then append them to their destination tiles.
__global__ void k1_migrate(buf1d p, const int64_t* __restrict__ tile_off, int* __restrict__ tile_np, mig1d mig,
                           ctl1d* __restrict__ ctl, int TX, int ntiles, int nx, int absorbing,
                           part1_aos* __restrict__ ovf, int* __restrict__ ovf_tag, unsigned int ovf_cap) {
	const int lane = threadIdx.x & 31;
	const int nwarp = (gridDim.x * blockDim.x) >> 5;
	for (int ts = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ts < ntiles; ts += nwarp) {
		const int n = mig.np[ts];
		const int64_t mb = tile_off[ts] / mig.div;
		for (int k = lane; k < n; k += 32) {
			part1_aos r = mig.rec[mb + k];
			int ix = r.ix;
			if (ix < 0 || ix >= nx) {
				if (absorbing) continue;
				ix += (ix < 0) ? nx : -nx;
			}
			int t = ix / TX;
			int slot = atomicAdd(&tile_np[t], 1);
			int64_t d = tile_off[t] + slot;
			if (d >= tile_off[t + 1]) {
				atomicSub(&tile_np[t], 1);
				r.ix = ix;
				ovf1_push(ctl, ovf, ovf_tag, ovf_cap, r, p.tag ? mig.tag[mb + k] : 0);
				continue;
			}
			rec20 v = { r.x, r.ux, r.uy, r.uz, ix - t * TX };
			rec1_store(p.rec, d, v);
			p.key[d] = key1_of(v.cell, v.x);
			if (p.tag) p.tag[d] = mig.tag[mb + k];
		}
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
	size_t smem = push1_smem_bytes(s->TX, s->max_cap);
	static size_t configured = 0;
	if (smem > configured) {
		ZDEV_CHECK(cudaFuncSetAttribute(k_push1d<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		ZDEV_CHECK(cudaFuncSetAttribute(k_push1d<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		configured = smem;
	}
	int slot = -1;
	if (zdev_time_push) {
		if (!s->ev) { s->ev = new std::vector<cudaEvent_t>(2 * EV1_RING); for (auto& e : *s->ev) ZDEV_CHECK(cudaEventCreate(&e)); }
		if (s->ev_pending == EV1_RING) collect_timing(s);
		slot = s->ev_next; s->ev_next = (s->ev_next + 1) % EV1_RING; s->ev_pending++;
		ZDEV_CHECK(cudaEventRecord((*s->ev)[2 * slot], zdev_strm));
	}
	const unsigned front = (unsigned) push1_smem_front(s->TX, s->max_cap), permb = (unsigned) push1_smem_perm(s->max_cap);
	if (s->track_ids)
		ZDEV_LAUNCH(k_push1d<true>, s->ntiles, P1_THREADS, smem, s->p, s->q, s->tile_off, s->tile_np, s->tile_np_q, s->mig, s->ctl,
		            zdev_grid1d_Epart(grid), zdev_grid1d_Bpart(grid), zdev_grid1d_J(gcur), s->nx, s->TX, *prm, front, permb);
	else
		ZDEV_LAUNCH(k_push1d<false>, s->ntiles, P1_THREADS, smem, s->p, s->q, s->tile_off, s->tile_np, s->tile_np_q, s->mig, s->ctl,
		            zdev_grid1d_Epart(grid), zdev_grid1d_Bpart(grid), zdev_grid1d_J(gcur), s->nx, s->TX, *prm, front, permb);
	if (slot >= 0) ZDEV_CHECK(cudaEventRecord((*s->ev)[2 * slot + 1], zdev_strm));
	{ buf1d t = s->p; s->p = s->q; s->q = t; }
	{ int* t = s->tile_np; s->tile_np = s->tile_np_q; s->tile_np_q = t; }
	ZDEV_LAUNCH(k1_migrate, 4 * zdev_num_sm, 256, 0, s->p, s->tile_off, s->tile_np, s->mig, s->ctl, s->TX, s->ntiles, s->nx, prm->absorbing,
	            s->ovf, s->ovf_tag, s->ovf_cap);
	spec1_resolve_overflow(s);
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
		rec20 v = rec1_load(p.rec, b + k, p.key[b + k]);
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
