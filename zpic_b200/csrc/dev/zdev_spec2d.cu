// zpic-b200 :: em2d particle species on the device.
//
// Data layout in HBM.  The grid is cut into tiles of TX x TY cells; every tile owns a
// fixed-capacity segment [tile_off[t], tile_off[t+1]) of
//     rec[]  24-byte records {x, y, ux, uy, uz, cell}, cell = lx | ly<<16 (tile local),
//            moved with three 64-bit accesses off one address register;
//     key[]  16-bit cell number lx + ly*TX (0xffff = empty slot), the only thing the
//            index sort has to read;
//     tag[]  optional injection index (parity runs),
// of which the first tile_np[t] slots are in use.  There are two such buffers, A and B,
// and every step streams A -> B.
//
// One CTA advances one tile (k_push2d):
//   phase A  an index-only counting sort of the tile's particles by cell, done in shared
//            memory with native integer atomics: perm[] lists the live slots in cell
//            order (empty slots drop out here, so compaction is free);
//   phase B  warps walk perm[] in order with no block barriers: gather the particle from
//            A, interpolate E/B from the shared-memory field tile, Boris push, move.
//            Because the lanes of a warp now sit in the same one or two cells, the eight
//            current contributions of the (non-crossing) particles are combined with a
//            segmented warp scan and only the last lane of each run issues the 8 L2
//            reductions.  The few particles that cross a cell face go to a warp-private
//            queue and are split/deposited 32 at a time, so the rare path costs no
//            divergence.  Survivors are written to B at their sorted position
//            (coalesced); particles that left the tile go to a global migrants list and
//            leave an empty slot behind.
// k_migrate2d then appends the migrants to their destination tiles in B.
// Per step a particle is read once and written once (2 x 26 B; 56 B with the reference's
// 28-byte record), plus the few percent that migrate.
//
// Replaces reference em2d/particles.c:1104-1269 (spec_advance incl. boundaries and
// spec_move_window's index shift) and :942-1007 (spec_sort - here the cell order is
// rebuilt every step instead of every n_sort steps).
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

// device particle record, 24 bytes, 8-byte aligned
struct rec24 { float x, y, ux, uy, uz; int cell; };
static_assert(sizeof(rec24) == 24, "rec24 must be 24 bytes");
#define KEY_EMPTY 0xffffu

// buffer view handed to kernels by value
struct soa2d {
	rec24* rec;          // tile buffers: cell = lx | ly<<16; migrants list: cell = global ix
	unsigned short* key; // tile buffers only
	int *iy;             // migrants list only: global iy
	int *tag;            // null unless ids are tracked
};

__device__ __forceinline__ rec24 rec_load(const rec24* p) {
	const float2* q = reinterpret_cast<const float2*>(p);
	float2 a = q[0], b = q[1], c = q[2];
	rec24 r; r.x = a.x; r.y = a.y; r.ux = b.x; r.uy = b.y; r.uz = c.x; r.cell = __float_as_int(c.y);
	return r;
}
__device__ __forceinline__ void rec_store(rec24* p, float x, float y, float ux, float uy, float uz, int cell) {
	float2* q = reinterpret_cast<float2*>(p);
	q[0] = make_float2(x, y); q[1] = make_float2(ux, uy); q[2] = make_float2(uz, __int_as_float(cell));
}

// control block in device memory, zeroed at the start of every advance
struct ctl2d {
	double energy;                   // sum of utsq/(gamma+1)
	unsigned long long np;           // live particles after the step
	unsigned int n_mig;              // entries in the migrants list
	unsigned int flags;              // 1: tile capacity overflow, 2: migrants list overflow, 4: export list overflow
	unsigned int n_exp[2];           // slab mode: particles handed to the left / right neighbour
	unsigned int pad[2];
};

struct zdev_spec2d {
	int nx, ny;
	int TX, TY, ntx, nty, ntiles;
	int ppc_hint, track_ids;
	double slack;
	int64_t cap_total;               // total SoA slots per buffer
	int max_cap;                     // largest tile capacity (sizes perm[] in shared memory)
	soa2d p, q;                      // current (A) and next (B) tile-binned buffers
	int64_t* tile_off;               // device, ntiles+1
	int* tile_np;                    // device, ntiles: slots in use in p
	int* tile_np_q;                  // device, ntiles: slots in use in q
	soa2d mig;                       // migrants list (global cell indices)
	unsigned int mig_cap;
	part_aos* stage; int64_t stage_cap;   // persistent staging for appended host particles
	part_aos* exp_buf[2];            // slab mode: export lists (AoS, ix already in the neighbour's frame)
	unsigned int exp_cap;
	ctl2d* ctl;                      // device
	int64_t np_host;                 // last known particle count
	int ids_valid;                   // tags are a permutation of [0,np)
	std::vector<int64_t>* h_off;     // host copy of tile_off
	// optional device timing of the push kernel alone (bench roofline): ring of event pairs
	std::vector<cudaEvent_t>* ev;    // 2*EV_RING events, created on first use
	int ev_next, ev_pending;
	double push_ms; int64_t push_launches, push_particles;
};
static const int EV_RING = 64;

static void spec_collect_timing(zdev_spec2d* s) {
	if (!s->ev) return;
	for (int k = 0; k < s->ev_pending; k++) {
		int slot = (s->ev_next - s->ev_pending + k + EV_RING) % EV_RING;
		float ms = 0;
		ZDEV_CHECK(cudaEventSynchronize((*s->ev)[2 * slot + 1]));
		ZDEV_CHECK(cudaEventElapsedTime(&ms, (*s->ev)[2 * slot], (*s->ev)[2 * slot + 1]));
		s->push_ms += ms; s->push_launches++;
	}
	s->ev_pending = 0;
}

extern "C" void zdev_spec2d_push_timing(zdev_spec2d* s, double* total_ms, int64_t* launches, int reset) {
	spec_collect_timing(s);
	if (total_ms) *total_ms = s->push_ms;
	if (launches) *launches = s->push_launches;
	if (reset) { s->push_ms = 0; s->push_launches = 0; }
}

#ifndef PUSH_THREADS_N
#define PUSH_THREADS_N 256
#endif
static const int PUSH_THREADS = PUSH_THREADS_N;
static const int PUSH_WARPS = PUSH_THREADS / 32;
#ifndef PUSH_MIN_BLOCKS
#define PUSH_MIN_BLOCKS 2           // CTAs per SM the register allocation is sized for
#endif
static const int XQ_CAP = 64;        // warp-private queue of cell-crossing particles

static void soa_alloc(soa2d& a, int64_t n, int with_tag, int is_mig) {
	size_t nn = (size_t) (n > 0 ? n : 1);
	memset(&a, 0, sizeof(a));
	ZDEV_CHECK(cudaMalloc(&a.rec, nn * sizeof(rec24)));
	if (is_mig) ZDEV_CHECK(cudaMalloc(&a.iy, nn * 4));
	else ZDEV_CHECK(cudaMalloc(&a.key, nn * 2));
	if (with_tag) ZDEV_CHECK(cudaMalloc(&a.tag, nn * 4));
}
static void soa_free(soa2d& a) {
	cudaFree(a.rec); cudaFree(a.key); cudaFree(a.iy); cudaFree(a.tag);
	memset(&a, 0, sizeof(a));
}

// supported tile shapes (kernel template instantiations)
static bool tile_supported(int tx, int ty) {
	return (tx == 16 && ty == 16) || (tx == 16 && ty == 8) || (tx == 8 && ty == 8) ||
	       (tx == 8 && ty == 4) || (tx == 4 && ty == 4);
}

extern "C" zdev_spec2d* zdev_spec2d_create(int nx, int ny, int ppc_hint, int track_ids) {
	zdev_require_init();
	zdev_spec2d* s = new zdev_spec2d();
	memset(s, 0, sizeof(*s));
	s->nx = nx; s->ny = ny;
	s->ppc_hint = ppc_hint > 0 ? ppc_hint : 1;
	s->track_ids = track_ids;
	// tile shape: aim at ~8192 particles per tile (measured best on B200: amortises the field
	// staging and the per-tile index sort, keeps the per-step migration to a few percent and
	// the shared-memory index buffer small enough for two CTAs per SM)
	int cells = 8192 / s->ppc_hint;
	int tx = 16, ty = 16;
	if (cells < 256) { tx = 16; ty = 8; }
	if (cells < 128) { tx = 8; ty = 8; }
	if (cells < 64)  { tx = 8; ty = 4; }
	if (cells < 32)  { tx = 4; ty = 4; }
	if (const char* e = getenv("ZPIC_TILE_X")) tx = atoi(e);
	if (const char* e = getenv("ZPIC_TILE_Y")) ty = atoi(e);
	if (!tile_supported(tx, ty)) {
		fprintf(stderr, "(*error*) zpic-b200: unsupported tile shape %dx%d (use 16x16, 16x8, 8x8, 8x4 or 4x4)\n", tx, ty);
		exit(-1);
	}
	s->TX = tx; s->TY = ty;
	s->ntx = (nx + tx - 1) / tx; s->nty = (ny + ty - 1) / ty;
	s->ntiles = s->ntx * s->nty;
	s->slack = 0.0;
	if (const char* e = getenv("ZPIC_TILE_SLACK")) s->slack = atof(e);
	ZDEV_CHECK(cudaMalloc(&s->tile_off, (size_t) (s->ntiles + 1) * sizeof(int64_t)));
	ZDEV_CHECK(cudaMalloc(&s->tile_np, (size_t) s->ntiles * sizeof(int)));
	ZDEV_CHECK(cudaMalloc(&s->tile_np_q, (size_t) s->ntiles * sizeof(int)));
	ZDEV_CHECK(cudaMemsetAsync(s->tile_np, 0, (size_t) s->ntiles * sizeof(int), zdev_strm));
	ZDEV_CHECK(cudaMemsetAsync(s->tile_np_q, 0, (size_t) s->ntiles * sizeof(int), zdev_strm));
	ZDEV_CHECK(cudaMalloc(&s->ctl, sizeof(ctl2d)));
	ZDEV_CHECK(cudaMemsetAsync(s->ctl, 0, sizeof(ctl2d), zdev_strm));
	s->h_off = new std::vector<int64_t>();
	return s;
}

static void spec_free_particles(zdev_spec2d* s) {
	for (int k = 0; k < 2; k++) { cudaFree(s->exp_buf[k]); s->exp_buf[k] = nullptr; }
	cudaFree(s->stage); s->stage = nullptr; s->stage_cap = 0;
	s->exp_cap = 0;
	if (s->cap_total) { soa_free(s->p); soa_free(s->q); soa_free(s->mig); }
	s->cap_total = 0; s->mig_cap = 0;
}

extern "C" void zdev_spec2d_destroy(zdev_spec2d* s) {
	if (!s) return;
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	spec_free_particles(s);
	cudaFree(s->tile_off); cudaFree(s->tile_np); cudaFree(s->tile_np_q); cudaFree(s->ctl);
	if (s->ev) { for (auto& e : *s->ev) cudaEventDestroy(e); delete s->ev; }
	delete s->h_off;
	delete s;
}

extern "C" void zdev_spec2d_tile_info(zdev_spec2d* s, int* tx, int* ty, int* ntiles, int64_t* capacity) {
	if (tx) *tx = s->TX; if (ty) *ty = s->TY; if (ntiles) *ntiles = s->ntiles; if (capacity) *capacity = s->cap_total;
}

// Lay out tile segments for the given per-tile populations and (re)allocate the SoA
// buffers.  Capacity per tile = slack * max(population, nominal fill), rounded to 32.
static void spec_layout(zdev_spec2d* s, const std::vector<int>& cnt, int64_t np) {
	double slack = s->slack;
	if (slack <= 0.0) slack = (np > (int64_t) 200000000) ? 1.25 : 2.0;
	std::vector<int64_t>& off = *s->h_off;
	off.assign(s->ntiles + 1, 0);
	int64_t max_cap = 0;
	for (int ty = 0; ty < s->nty; ty++) for (int tx = 0; tx < s->ntx; tx++) {
		int t = tx + ty * s->ntx;
		int cx = (tx + 1) * s->TX <= s->nx ? s->TX : s->nx - tx * s->TX;
		int cy = (ty + 1) * s->TY <= s->ny ? s->TY : s->ny - ty * s->TY;
		int64_t nominal = (int64_t) cx * cy * s->ppc_hint;
		int64_t want = cnt[t] > nominal ? cnt[t] : nominal;
		int64_t cap = (int64_t) (want * slack) + 64;
		cap = (cap + 31) & ~(int64_t) 31;
		off[t + 1] = off[t] + cap;
		if (cap > max_cap) max_cap = cap;
	}
	if (max_cap * 4 > 160 * 1024) {
		fprintf(stderr, "(*error*) zpic-b200: %lld particles in one %dx%d tile exceed the shared-memory index "
		        "buffer; use smaller tiles (ZPIC_TILE_X/Y)\n", (long long) max_cap, s->TX, s->TY);
		exit(-1);
	}
	int64_t total = off[s->ntiles];
	spec_free_particles(s);
	soa_alloc(s->p, total, s->track_ids, 0);
	soa_alloc(s->q, total, s->track_ids, 0);
	s->cap_total = total;
	s->max_cap = (int) max_cap;
	int64_t mc = total / 8 + 65536;
	if (mc > 0x7fffffff) mc = 0x7fffffff;
	s->mig_cap = (unsigned int) mc;
	soa_alloc(s->mig, s->mig_cap, s->track_ids, 1);
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
	int tx = r.ix / TX, ty = r.iy / TY, t = tx + ty * ntx;
	int slot = atomicAdd(&tile_np[t], 1);
	int64_t d = off[t] + slot;
	if (d >= off[t + 1]) { atomicOr(&ctl->flags, 1u); return; }
	const int lx = r.ix - tx * TX, ly = r.iy - ty * TY;
	rec_store(p.rec + d, r.x, r.y, r.ux, r.uy, r.uz, lx | (ly << 16));
	p.key[d] = (unsigned short) (lx + ly * TX);
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
	if (flags & 4u) {
		fprintf(stderr, "(*error*) zpic-b200: slab export list overflow (capacity %u), aborting.\n", s->exp_cap);
		exit(-1);
	}
}

static void spec_append_dev(zdev_spec2d* s, const part_aos* d_aos, int64_t np, int tag0) {
	if (np <= 0) return;
	ZDEV_LAUNCH(k_scatter_tiles, zdev_div_up(np, 256), 256, 0, d_aos, np, s->TX, s->TY, s->ntx,
	            s->p, s->tile_off, s->tile_np, s->ctl, tag0);
}

// Host buffers of at least this size are pinned + mapped once and then read / written by the binning
// and gather kernels directly over PCIe (zero copy): no device staging copy of the population, and
// repeated transfers of the same mirror (ZPIC_COHERENT) run at link speed instead of pageable speed.
static const size_t MAP_MIN_BYTES = (size_t) 32 << 20;
struct host_map { const void* ptr; size_t bytes; void* dev; };
static host_map g_maps[8];

static void* map_host(const void* ptr, size_t bytes) {
	size_t min_bytes = MAP_MIN_BYTES;
	if (const char* e = getenv("ZPIC_ZERO_COPY_MIN")) min_bytes = (size_t) atoll(e);     // tests: force / forbid
	if (bytes < min_bytes || bytes == 0) return nullptr;
	for (auto& m : g_maps) if (m.ptr == ptr && m.bytes >= bytes) return m.dev;
	host_map* slot = &g_maps[0];
	for (auto& m : g_maps) { if (m.ptr == ptr) { slot = &m; break; } if (!m.ptr) slot = &m; }
	if (slot->ptr) { cudaHostUnregister((void*) slot->ptr); slot->ptr = nullptr; }
	if (cudaHostRegister((void*) ptr, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable) != cudaSuccess) {
		cudaGetLastError();
		return nullptr;                    // not pinnable (e.g. over the locked-memory limit): staged path
	}
	void* dev = nullptr;
	if (cudaHostGetDevicePointer(&dev, (void*) ptr, 0) != cudaSuccess) { cudaGetLastError(); cudaHostUnregister((void*) ptr); return nullptr; }
	slot->ptr = ptr; slot->bytes = bytes; slot->dev = dev;
	return dev;
}
// a host buffer is about to be freed / reallocated by its owner
extern "C" void zdev_host_forget(const void* ptr) {
	for (auto& m : g_maps) if (m.ptr == ptr) { cudaHostUnregister((void*) m.ptr); m.ptr = nullptr; }
}

extern "C" void zdev_spec2d_upload(zdev_spec2d* s, const void* part, int64_t np) {
	const size_t bytes = (size_t) np * sizeof(part_aos);
	part_aos* d_stage = nullptr;
	const part_aos* src = (np > 0) ? (const part_aos*) map_host(part, bytes) : nullptr;
	if (np > 0 && !src) {
		ZDEV_CHECK(cudaMalloc(&d_stage, bytes));
		ZDEV_CHECK(cudaMemcpyAsync(d_stage, part, bytes, cudaMemcpyHostToDevice, zdev_strm));
		src = d_stage;
	}
	ctl2d h;
	bool done = false;
	if (np > 0 && s->cap_total > 0) {
		// optimistic: bin straight into the existing tile layout, fall back if a tile overflows
		ZDEV_CHECK(cudaMemsetAsync(s->tile_np, 0, (size_t) s->ntiles * sizeof(int), zdev_strm));
		ZDEV_CHECK(cudaMemsetAsync(s->ctl, 0, sizeof(ctl2d), zdev_strm));
		spec_append_dev(s, src, np, 0);
		ZDEV_CHECK(cudaMemcpyAsync(&h, s->ctl, sizeof h, cudaMemcpyDeviceToHost, zdev_strm));
		ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
		done = !(h.flags & 1u);
	}
	if (!done) {
		std::vector<int> cnt(s->ntiles, 0);
		if (np > 0) {
			int* d_cnt; ZDEV_CHECK(cudaMalloc(&d_cnt, (size_t) s->ntiles * sizeof(int)));
			ZDEV_CHECK(cudaMemsetAsync(d_cnt, 0, (size_t) s->ntiles * sizeof(int), zdev_strm));
			ZDEV_LAUNCH(k_count_tiles, zdev_div_up(np, 256), 256, 0, src, np, s->TX, s->TY, s->ntx, d_cnt);
			ZDEV_CHECK(cudaMemcpyAsync(cnt.data(), d_cnt, (size_t) s->ntiles * sizeof(int), cudaMemcpyDeviceToHost, zdev_strm));
			ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
			cudaFree(d_cnt);
		}
		spec_layout(s, cnt, np);
		ZDEV_CHECK(cudaMemsetAsync(s->ctl, 0, sizeof(ctl2d), zdev_strm));
		spec_append_dev(s, src, np, 0);
		ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	}
	if (d_stage) cudaFree(d_stage);
	s->np_host = np;
	s->ids_valid = s->track_ids;
}

// Appends are on the per-step path of a moving window (one injected column per move): the staging
// buffer is kept, nothing is synchronised, and a capacity overflow is reported by the next fetch.
extern "C" void zdev_spec2d_append(zdev_spec2d* s, const void* part, int64_t np) {
	if (np <= 0) return;
	if (!s->cap_total) { zdev_spec2d_upload(s, part, np); return; }
	if (np > s->stage_cap) {
		if (s->stage) { ZDEV_CHECK(cudaStreamSynchronize(zdev_strm)); cudaFree(s->stage); }
		s->stage_cap = np + np / 2 + 1024;
		ZDEV_CHECK(cudaMalloc(&s->stage, (size_t) s->stage_cap * sizeof(part_aos)));
	}
	// pageable source: the copy is staged by the driver before the call returns, so `part` may be freed
	ZDEV_CHECK(cudaMemcpyAsync(s->stage, part, (size_t) np * sizeof(part_aos), cudaMemcpyHostToDevice, zdev_strm));
	spec_append_dev(s, s->stage, np, (int) s->np_host);
	s->np_host += np;
}

// per tile: number of live slots (slots whose cell is not -1)
__global__ void k_count_live(soa2d p, const int64_t* __restrict__ off, const int* __restrict__ tile_np, int* live) {
	int t = blockIdx.x;
	int n = tile_np[t], c = 0;
	int64_t b = off[t];
	for (int k = threadIdx.x; k < n; k += blockDim.x) c += (p.key[b + k] != KEY_EMPTY);
	for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
	__shared__ int w[8];
	if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = c;
	__syncthreads();
	if (threadIdx.x == 0) { int s = 0; for (int k = 0; k < (int) (blockDim.x >> 5); k++) s += w[k]; live[t] = s; }
}

// tiles -> AoS.  prefix[t] = first output slot of tile t (ignored when scattering by tag)
__global__ void k_gather_aos(soa2d p, const int64_t* __restrict__ off, const int* __restrict__ tile_np,
                             const int64_t* __restrict__ prefix, part_aos* __restrict__ out, int by_tag,
                             int TX, int TY, int ntx) {
	int t = blockIdx.x;
	int n = tile_np[t];
	int x0 = (t % ntx) * TX, y0 = (t / ntx) * TY;
	int64_t b = off[t], o = prefix[t];
	__shared__ int s_run;
	if (threadIdx.x == 0) s_run = 0;
	__syncthreads();
	for (int k0 = 0; k0 < n; k0 += blockDim.x) {
		int k = k0 + threadIdx.x;
		bool live = (k < n) && p.key[b + k] != KEY_EMPTY;
		unsigned m = __ballot_sync(0xffffffffu, live);
		int wbase = 0;
		if ((threadIdx.x & 31) == 0 && m) wbase = atomicAdd(&s_run, __popc(m));
		wbase = __shfl_sync(0xffffffffu, wbase, 0);
		if (live) {
			rec24 v = rec_load(p.rec + b + k);
			part_aos r;
			r.ix = x0 + (v.cell & 0xffff); r.iy = y0 + (v.cell >> 16);
			r.x = v.x; r.y = v.y; r.ux = v.ux; r.uy = v.uy; r.uz = v.uz;
			int64_t d = by_tag ? (int64_t) p.tag[b + k] : o + wbase + __popc(m & ((1u << (threadIdx.x & 31)) - 1));
			out[d] = r;
		}
	}
}

// live particles per tile -> host vector; returns the total
static int64_t spec_live_counts(zdev_spec2d* s, std::vector<int>& cnt) {
	int* d_live; ZDEV_CHECK(cudaMalloc(&d_live, (size_t) s->ntiles * sizeof(int)));
	ZDEV_LAUNCH(k_count_live, s->ntiles, 256, 0, s->p, s->tile_off, s->tile_np, d_live);
	cnt.resize(s->ntiles);
	ZDEV_CHECK(cudaMemcpyAsync(cnt.data(), d_live, (size_t) s->ntiles * sizeof(int), cudaMemcpyDeviceToHost, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	cudaFree(d_live);
	int64_t np = 0;
	for (int t = 0; t < s->ntiles; t++) np += cnt[t];
	return np;
}

extern "C" int64_t zdev_spec2d_download(zdev_spec2d* s, void* part, int64_t max_np) {
	if (!s->cap_total) return 0;
	std::vector<int> cnt;
	int64_t np = spec_live_counts(s, cnt);
	std::vector<int64_t> prefix(s->ntiles);
	int64_t acc = 0;
	for (int t = 0; t < s->ntiles; t++) { prefix[t] = acc; acc += cnt[t]; }
	s->np_host = np;
	if (np == 0) return 0;
	if (np > max_np) {
		fprintf(stderr, "(*error*) zpic-b200: host particle buffer too small (%lld > %lld)\n", (long long) np, (long long) max_np);
		exit(-1);
	}
	int64_t* d_prefix; part_aos* d_aos = nullptr;
	ZDEV_CHECK(cudaMalloc(&d_prefix, (size_t) s->ntiles * sizeof(int64_t)));
	ZDEV_CHECK(cudaMemcpyAsync(d_prefix, prefix.data(), (size_t) s->ntiles * sizeof(int64_t), cudaMemcpyHostToDevice, zdev_strm));
	// large mirrors are written by the gather kernel directly (mapped pinned host memory)
	part_aos* dst = (part_aos*) map_host(part, (size_t) max_np * sizeof(part_aos));
	if (!dst) { ZDEV_CHECK(cudaMalloc(&d_aos, (size_t) np * sizeof(part_aos))); dst = d_aos; }
	ZDEV_LAUNCH(k_gather_aos, s->ntiles, 256, 0, s->p, s->tile_off, s->tile_np, d_prefix, dst,
	            (s->track_ids && s->ids_valid) ? 1 : 0, s->TX, s->TY, s->ntx);
	if (d_aos) ZDEV_CHECK(cudaMemcpyAsync(part, d_aos, (size_t) np * sizeof(part_aos), cudaMemcpyDeviceToHost, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	cudaFree(d_prefix); cudaFree(d_aos);
	return np;
}

extern "C" int64_t zdev_spec2d_np(zdev_spec2d* s) {
	if (!s->cap_total) return 0;
	std::vector<int> cnt;
	s->np_host = spec_live_counts(s, cnt);
	return s->np_host;
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
		rec_store(p.rec + d, (float) (dpcx * (kx + 0.5)), (float) (dpcy * (ky + 0.5)),
		          uth.x * a + (ufl.x - sx), uth.y * b + (ufl.y - sy), uth.z * c + (ufl.z - sz), lx | (ly << 16));
		p.key[d] = (unsigned short) (lx + ly * TX);
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
	int ntx;                 // tiles per row
};

// scatter one segment's 8 contributions into the global J grid (L2 reductions)
__device__ __forceinline__ void red_weights(f3* __restrict__ c, int nrow, const float w[8]) {
	atomicAdd(&c[0].x, w[0]);
	atomicAdd(&c[nrow].x, w[1]);
	atomicAdd(&c[0].y, w[2]);
	atomicAdd(&c[1].y, w[3]);
	atomicAdd(&c[0].z, w[4]);
	atomicAdd(&c[1].z, w[5]);
	atomicAdd(&c[nrow].z, w[6]);
	atomicAdd(&c[nrow + 1].z, w[7]);
}
__device__ __forceinline__ void deposit_seg_global(f3* __restrict__ J, int nrow, const seg2d& s, float qnx, float qny) {
	float w[8];
	seg_weights(s, qnx, qny, w);
	red_weights(J + (s.ix + 1) + (s.iy + 1) * nrow, nrow, w);
}

// one queued cell-crossing move (warp-private shared-memory queue)
struct xq_entry { int ix, iy, dij; float x0, y0, dx, dy, qvz; };

// split + deposit up to 32 queued moves, one per lane
__device__ __forceinline__ void drain_crossers(const xq_entry* q, int n, int lane, f3* __restrict__ J, int nrow,
                                               float qnx, float qny) {
	if (lane < n) {
		xq_entry e = q[lane];
		seg2d vp[3];
		int vnp = split_trajectory(e.ix, e.iy, (e.dij & 3) - 1, ((e.dij >> 2) & 3) - 1, e.x0, e.y0, e.dx, e.dy, e.qvz, vp);
		deposit_seg_global(J, nrow, vp[0], qnx, qny);
		deposit_seg_global(J, nrow, vp[1], qnx, qny);
		if (vnp > 2) deposit_seg_global(J, nrow, vp[2], qnx, qny);
	}
}

// One CTA per tile.  Dynamic shared memory: perm[max_cap] ints.
// (A variant specialised for the plain periodic case - no window, slab or tag logic - was measured
// 6 % SLOWER than this generic kernel on B200: 3.87 vs 3.66 ms at 67 M particles; not kept.)
template <int TX, int TY>
__global__ void __launch_bounds__(PUSH_THREADS, PUSH_MIN_BLOCKS)
k_push2d(soa2d A, soa2d Bo, const int64_t* __restrict__ tile_off, const int* __restrict__ tile_np,
         int* __restrict__ tile_np_out, soa2d mig, unsigned int mig_cap,
         ctl2d* __restrict__ ctl, const f3* __restrict__ E, const f3* __restrict__ B, f3* __restrict__ J,
         push_geom g, zdev_push2d_params prm) {
	constexpr int SROW = TX + 2;
	constexpr int PLANE = SROW * (TY + 2);
	constexpr int NC = TX * TY;
	extern __shared__ int s_perm[];
	__shared__ float s_fld[6 * PLANE];
	__shared__ int s_cnt[NC];
	__shared__ int s_wsum[PUSH_WARPS];
	__shared__ xq_entry s_xq[PUSH_WARPS][XQ_CAP];

	const int t = blockIdx.x;
	const int tx = t % g.ntx, ty = t / g.ntx;
	const int x0 = tx * TX, y0 = ty * TY;
	const int cx = min(TX, g.nx - x0), cy = min(TY, g.ny - y0);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int n = tile_np[t];
	const int64_t base = tile_off[t];

	// ---- stage the field neighbourhood as planes: cells [x0-1, x0+cx] x [y0-1, y0+cy]
	for (int k = threadIdx.x; k < (cx + 2) * (cy + 2); k += PUSH_THREADS) {
		int r = k / (cx + 2), c = k - r * (cx + 2);
		int gi = (x0 + c) + (y0 + r) * g.nrow;        // buffer index of cell (x0-1+c, y0-1+r)
		f3 e = E[gi], b = B[gi];
		int o = c + r * SROW;
		s_fld[o] = e.x; s_fld[o + PLANE] = e.y; s_fld[o + 2 * PLANE] = e.z;
		s_fld[o + 3 * PLANE] = b.x; s_fld[o + 4 * PLANE] = b.y; s_fld[o + 5 * PLANE] = b.z;
	}
	for (int k = threadIdx.x; k < NC; k += PUSH_THREADS) s_cnt[k] = 0;
	__syncthreads();

	// ---- phase A: counting sort of slot indices by cell
	for (int i = threadIdx.x; i < n; i += PUSH_THREADS) {
		unsigned c = A.key[base + i];
		if (c != KEY_EMPTY) atomicAdd(&s_cnt[c], 1);
	}
	__syncthreads();
	int nlive;
	{	// exclusive scan of s_cnt (NC <= 256 == PUSH_THREADS): s_cnt becomes the write cursor
		int v = (threadIdx.x < NC) ? s_cnt[threadIdx.x] : 0;
		int incl = v;
		for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += u; }
		if (lane == 31) s_wsum[warp] = incl;
		__syncthreads();
		int woff = 0, tot = 0;
		#pragma unroll
		for (int w = 0; w < PUSH_WARPS; w++) { int c = s_wsum[w]; woff += (w < warp) ? c : 0; tot += c; }
		if (threadIdx.x < NC) s_cnt[threadIdx.x] = woff + incl - v;
		nlive = tot;
		__syncthreads();
	}
	for (int i = threadIdx.x; i < n; i += PUSH_THREADS) {
		unsigned c = A.key[base + i];
		if (c != KEY_EMPTY) s_perm[atomicAdd(&s_cnt[c], 1)] = i;
	}
	__syncthreads();

	// ---- phase B: warps stream the sorted particles, no block barriers from here on
	xq_entry* xq = s_xq[warp];
	int nxq = 0;
	float energy = 0.0f;      // per-thread partial in float (<= 64 terms), widened once per tile
	f3* const J0 = J + (x0 + 1) + (y0 + 1) * g.nrow;          // cell (x0,y0)

	// software pipeline: the record of the next iteration is requested before the current one
	// is processed, so its HBM/L2 latency hides behind ~600 instructions of arithmetic
	rec24 nv; int ntag = 0;
	const int first = (nlive > 0) ? s_perm[0] : 0;
	{
		const int pn = warp * 32 + lane;
		const int64_t k = base + ((pn < nlive) ? s_perm[pn] : first);
		nv = rec_load(A.rec + k);
		if (A.tag) ntag = A.tag[k];
	}

	for (int p0 = warp * 32; p0 < nlive; p0 += PUSH_THREADS) {
		const int p = p0 + lane;
		const bool active = p < nlive;
		const rec24 v = nv;
		const int tag = ntag;
		{
			const int pn = p + PUSH_THREADS;
			const int64_t k = base + ((pn < nlive) ? s_perm[pn] : first);
			nv = rec_load(A.rec + k);
			if (A.tag) ntag = A.tag[k];
		}

		// Lanes past the end of the tile (last iteration only) run the arithmetic on a copy of
		// the tile's first particle and are masked out of every side effect below.
		float w[8];
		const int lx = v.cell & 0xffff, ly = v.cell >> 16;
		const int key = active ? lx + ly * TX : 0x7fffffff;
		float x = v.x, y = v.y, ux = v.ux, uy = v.uy, uz = v.uz;
		int fate, ncell = -1, gix = 0, giy = 0;
		bool crosses;
		xq_entry xe;
		{
			f3 Ep, Bp;
			interp_EB_planes<SROW, PLANE>(s_fld, lx, ly, x, y, Ep, Bp);
			const float en = boris(Ep, Bp, prm.tem, ux, uy, uz);
			energy += active ? en : 0.0f;

			float rg = div_exact(1.0f, sqrt_exact(1.0f + ux * ux + uy * uy + uz * uz));
			float dx = prm.dt_dx * rg * ux;
			float dy = prm.dt_dy * rg * uy;
			float x1 = x + dx, y1 = y + dy;
			int di = ltrim(x1), dj = ltrim(y1);
			x1 -= di; y1 -= dj;
			float qvz = prm.q * uz * rg;

			crosses = active && ((di | dj) != 0);
			xe.ix = x0 + lx; xe.iy = y0 + ly; xe.dij = (di + 1) | ((dj + 1) << 2);
			xe.x0 = x; xe.y0 = y; xe.dx = dx; xe.dy = dy; xe.qvz = qvz;
			{
				seg2d s0;
				s0.x0 = x; s0.y0 = y; s0.dx = dx; s0.dy = dy; s0.x1 = x + dx; s0.y1 = y + dy;
				s0.qvz = qvz * 0.5f; s0.ix = 0; s0.iy = 0;
				seg_weights(s0, prm.qnx, prm.qny, w);
				const bool zero = crosses || !active;      // crossers deposit through the queue
				#pragma unroll
				for (int q = 0; q < 8; q++) w[q] = zero ? 0.0f : w[q];
			}

			x = x1; y = y1;
			int ix = x0 + lx + di - prm.shift_window, iy = y0 + ly + dj;
			// boundaries (reference particles.c:1237-1259)
			// an x edge is either a slab boundary (particle handed to the neighbour rank through the
			// migrants list, keeping its out-of-range index), absorbing (moving window) or periodic
			fate = active ? 1 : 0;
			if (ix < 0) {
				if (!prm.slab_left) { if (prm.moving_window) fate = 0; else ix += g.nx; }
			} else if (ix >= g.nx) {
				if (!prm.slab_right) { if (prm.moving_window) fate = 0; else ix -= g.nx; }
			}
			iy += ((iy < 0) ? g.ny : 0) - ((iy >= g.ny) ? g.ny : 0);
			const int nlx = ix - x0, nly = iy - y0;
			if (fate) {
				if (nlx < 0 || nlx >= cx || nly < 0 || nly >= cy) { fate = 2; gix = ix; giy = iy; }
				else ncell = nlx | (nly << 16);
			}
		}

		// --- current of the non-crossing particles, combined per run of equal cell
		{
			const int prev = __shfl_up_sync(0xffffffffu, key, 1);
			const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || key != prev);
			if (heads == 1u) {
				// all 32 lanes in one cell: transposed butterfly, 9 shuffles for the 8 sums, after
				// which lane 4*k holds the total of contribution k and issues its single reduction
				float v4[4], v2[2], v1;
				const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4;
				#pragma unroll
				for (int q = 0; q < 4; q++) {
					float send = b16 ? w[q] : w[q + 4], keep = b16 ? w[q + 4] : w[q];
					v4[q] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
				}
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
				if ((lane & 3) == 0) {
					const int k = lane >> 2;                       // contribution index, see seg_weights()
					const int comp = (k < 2) ? 0 : ((k < 4) ? 1 : 2);
					const int right = (k == 3) | (k == 5) | (k == 7);
					const int up = (k == 1) | (k == 6) | (k == 7);
					float* a = reinterpret_cast<float*>(J0 + lx + right + (ly + up) * g.nrow) + comp;
					atomicAdd(a, v1);
				}
			} else {
				// general case: segmented inclusive scan, the last lane of each run holds its totals
				const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
				#pragma unroll
				for (int d = 1; d < 32; d <<= 1) {
					const bool take = (lane - d) >= start;
					#pragma unroll
					for (int q = 0; q < 8; q++) {
						float u = __shfl_up_sync(0xffffffffu, w[q], d);
						if (take) w[q] += u;
					}
				}
				const bool tail = (lane == 31) || ((heads >> (lane + 1)) & 1u);
				if (tail && active) red_weights(J0 + lx + ly * g.nrow, g.nrow, w);
			}
		}

		// --- cell crossers: queue, drain 32 at a time
		{
			const unsigned xm = __ballot_sync(0xffffffffu, crosses);
			if (xm) {
				if (crosses) xq[nxq + __popc(xm & ((1u << lane) - 1))] = xe;
				nxq += __popc(xm);
				__syncwarp();
				if (nxq >= 32) {
					drain_crossers(xq + nxq - 32, 32, lane, J, g.nrow, prm.qnx, prm.qny);
					nxq -= 32;
					__syncwarp();
				}
			}
		}

		// --- write the survivors to their sorted slot in B, route the leavers
		if (active) {
			const int64_t d = base + p;
			Bo.key[d] = (fate == 1) ? (unsigned short) ((ncell & 0xffff) + (ncell >> 16) * TX) : (unsigned short) KEY_EMPTY;
			if (fate == 1) {
				rec_store(Bo.rec + d, x, y, ux, uy, uz, ncell);
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
				if (d < mig_cap) {
					rec_store(mig.rec + d, x, y, ux, uy, uz, gix);
					mig.iy[d] = giy;
					if (mig.tag) mig.tag[d] = tag;
				} else atomicOr(&ctl->flags, 2u);
			}
		}
	}
	if (nxq) drain_crossers(xq, nxq, lane, J, g.nrow, prm.qnx, prm.qny);

	// ---- tile epilogue (no block barrier: warps retire independently): slots in use, energy
	double e = (double) energy;
	for (int o = 16; o > 0; o >>= 1) e += __shfl_down_sync(0xffffffffu, e, o);
	if (lane == 0 && nlive > 0) atomicAdd(&ctl->energy, e);
	if (threadIdx.x == 0) tile_np_out[t] = nlive;
}

// append the migrants to their destination tiles and count the population
__global__ void k_migrate2d(soa2d p, const int64_t* __restrict__ tile_off, int* __restrict__ tile_np, soa2d mig,
                            unsigned int mig_cap, ctl2d* __restrict__ ctl, int TX, int TY, int ntx, int nx,
                            part_aos* __restrict__ exp_l, part_aos* __restrict__ exp_r, unsigned int exp_cap) {
	unsigned int n = ctl->n_mig;
	if (n > mig_cap) n = mig_cap;
	for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
		rec24 v = rec_load(mig.rec + k);
		int ix = v.cell, iy = mig.iy[k];
		if (ix < 0 || ix >= nx) {
			// leaves the slab: export in the neighbour's frame (all slabs have the same width)
			const int side = ix >= nx;
			unsigned int slot = atomicAdd(&ctl->n_exp[side], 1u);
			if (slot >= exp_cap) { atomicOr(&ctl->flags, 4u); continue; }
			part_aos r; r.ix = side ? ix - nx : ix + nx; r.iy = iy;
			r.x = v.x; r.y = v.y; r.ux = v.ux; r.uy = v.uy; r.uz = v.uz;
			(side ? exp_r : exp_l)[slot] = r;
			continue;
		}
		int tx = ix / TX, ty = iy / TY, t = tx + ty * ntx;
		int slot = atomicAdd(&tile_np[t], 1);
		int64_t d = tile_off[t] + slot;
		if (d >= tile_off[t + 1]) { atomicOr(&ctl->flags, 1u); continue; }
		const int lx = ix - tx * TX, ly = iy - ty * TY;
		rec_store(p.rec + d, v.x, v.y, v.ux, v.uy, v.uz, lx | (ly << 16));
		p.key[d] = (unsigned short) (lx + ly * TX);
		if (p.tag) p.tag[d] = mig.tag[k];
	}
}

// total live particles (slots with cell >= 0) -> ctl->np
__global__ void k_count_total(soa2d p, const int64_t* __restrict__ off, const int* __restrict__ tile_np, int ntiles, ctl2d* ctl) {
	unsigned long long c = 0;
	for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
		int n = tile_np[t];
		int64_t b = off[t];
		for (int k = threadIdx.x; k < n; k += blockDim.x) c += (p.key[b + k] != KEY_EMPTY);
	}
	for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
	if ((threadIdx.x & 31) == 0 && c) atomicAdd(&ctl->np, c);
}

template <int TX, int TY>
static void launch_push(zdev_spec2d* s, const f3* E, const f3* B, f3* J, const push_geom& g, const zdev_push2d_params& prm) {
	size_t smem = (size_t) s->max_cap * sizeof(int);
	static size_t configured = 0;
	if (smem > configured) {
		ZDEV_CHECK(cudaFuncSetAttribute(k_push2d<TX, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		configured = smem;
	}
	int slot = -1;
	if (zdev_time_push) {
		if (!s->ev) {
			s->ev = new std::vector<cudaEvent_t>(2 * EV_RING);
			for (auto& e : *s->ev) ZDEV_CHECK(cudaEventCreate(&e));
		}
		if (s->ev_pending == EV_RING) spec_collect_timing(s);
		slot = s->ev_next; s->ev_next = (s->ev_next + 1) % EV_RING; s->ev_pending++;
		ZDEV_CHECK(cudaEventRecord((*s->ev)[2 * slot], zdev_strm));
	}
	ZDEV_LAUNCH((k_push2d<TX, TY>), s->ntiles, PUSH_THREADS, smem, s->p, s->q, s->tile_off, s->tile_np, s->tile_np_q,
	            s->mig, s->mig_cap, s->ctl, E, B, J, g, prm);
	if (slot >= 0) ZDEV_CHECK(cudaEventRecord((*s->ev)[2 * slot + 1], zdev_strm));
}

extern "C" void zdev_spec2d_advance(zdev_spec2d* s, zdev_grid2d* grid, zdev_grid2d* gcur, const zdev_push2d_params* prm) {
	if (zdev_grid2d_nx(grid) != s->nx || zdev_grid2d_ny(grid) != s->ny ||
	    zdev_grid2d_nx(gcur) != s->nx || zdev_grid2d_ny(gcur) != s->ny) {
		fprintf(stderr, "(*error*) zdev_spec2d_advance: species / grid size mismatch\n"); exit(-1);
	}
	ZDEV_CHECK(cudaMemsetAsync(s->ctl, 0, sizeof(ctl2d), zdev_strm));
	if (!s->cap_total) return;
	push_geom g = { s->nx, s->ny, s->nx + 3, s->ntx };
	const f3* E = zdev_grid2d_Epart(grid); const f3* B = zdev_grid2d_Bpart(grid); f3* J = zdev_grid2d_J(gcur);
	if      (s->TX == 16 && s->TY == 16) launch_push<16, 16>(s, E, B, J, g, *prm);
	else if (s->TX == 16 && s->TY == 8)  launch_push<16, 8>(s, E, B, J, g, *prm);
	else if (s->TX == 8  && s->TY == 8)  launch_push<8, 8>(s, E, B, J, g, *prm);
	else if (s->TX == 8  && s->TY == 4)  launch_push<8, 4>(s, E, B, J, g, *prm);
	else                                 launch_push<4, 4>(s, E, B, J, g, *prm);
	// B becomes the current buffer
	{ soa2d t = s->p; s->p = s->q; s->q = t; }
	{ int* t = s->tile_np; s->tile_np = s->tile_np_q; s->tile_np_q = t; }
	if ((prm->slab_left || prm->slab_right) && !s->exp_cap) {
		// a window shift sends a whole column at once: size the export lists for two columns
		int64_t cap = (int64_t) 2 * s->ppc_hint * s->ny + 65536;
		s->exp_cap = (unsigned int) cap;
		for (int k = 0; k < 2; k++) ZDEV_CHECK(cudaMalloc(&s->exp_buf[k], (size_t) cap * sizeof(part_aos)));
	}
	ZDEV_LAUNCH(k_migrate2d, 2 * zdev_num_sm, 256, 0, s->p, s->tile_off, s->tile_np, s->mig, s->mig_cap, s->ctl,
	            s->TX, s->TY, s->ntx, s->nx, s->exp_buf[0], s->exp_buf[1], s->exp_cap);
	if (prm->moving_window || prm->slab_left || prm->slab_right) s->ids_valid = 0;
}

extern "C" void zdev_spec2d_fetch(zdev_spec2d* s, double* energy_sum, int64_t* np) {
	ctl2d h;
	memset(&h, 0, sizeof h);
	if (s->cap_total && np)
		ZDEV_LAUNCH(k_count_total, 4 * zdev_num_sm, 256, 0, s->p, s->tile_off, s->tile_np, s->ntiles, s->ctl);
	ZDEV_CHECK(cudaMemcpyAsync(&h, s->ctl, sizeof(h), cudaMemcpyDeviceToHost, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	check_flags(s, h.flags);
	if (np) {
		// the count is accumulated into ctl->np by k_count_total: reset so a second fetch does not double it
		ZDEV_CHECK(cudaMemsetAsync(&s->ctl->np, 0, sizeof(unsigned long long), zdev_strm));
		s->np_host = (int64_t) h.np;
		*np = (int64_t) h.np;
	}
	if (energy_sum) *energy_sum = h.energy;
}

// ------------------------------------------------------------------ charge deposit

// reference spec_deposit_charge, em2d/particles.c:1289-1324 (node centred, linear)
__global__ void k_deposit_charge(soa2d p, const int64_t* __restrict__ off, const int* __restrict__ tile_np,
                                 float* __restrict__ rho, int nrow, float q, int TX, int TY, int ntx) {
	int t = blockIdx.x;
	int n = tile_np[t];
	int x0 = (t % ntx) * TX, y0 = (t / ntx) * TY;
	int64_t b = off[t];
	for (int k = threadIdx.x; k < n; k += blockDim.x) {
		if (p.key[b + k] == KEY_EMPTY) continue;
		rec24 v = rec_load(p.rec + b + k);
		int idx = (x0 + (v.cell & 0xffff)) + nrow * (y0 + (v.cell >> 16));
		float w1 = v.x, w2 = v.y;
		atomicAdd(&rho[idx], (1.0f - w1) * (1.0f - w2) * q);
		atomicAdd(&rho[idx + 1], (w1) * (1.0f - w2) * q);
		atomicAdd(&rho[idx + nrow], (1.0f - w1) * (w2) * q);
		atomicAdd(&rho[idx + 1 + nrow], (w1) * (w2) * q);
	}
}
__global__ void k_charge_fold(float* __restrict__ rho, int nx, int ny, int mode) {
	int nrow = nx + 1;
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (mode == 0) { if (k <= ny) rho[(size_t) k * nrow] += rho[nx + (size_t) k * nrow]; }   // x fold
	else           { if (k <= nx) rho[k] += rho[k + (size_t) ny * nrow]; }                   // y fold
}

extern "C" void zdev_spec2d_deposit_charge(zdev_spec2d* s, float q, int moving_window, float* charge) {
	size_t n = (size_t) (s->nx + 1) * (s->ny + 1);
	float* d_rho; ZDEV_CHECK(cudaMalloc(&d_rho, n * sizeof(float)));
	ZDEV_CHECK(cudaMemcpyAsync(d_rho, charge, n * sizeof(float), cudaMemcpyHostToDevice, zdev_strm));
	if (s->cap_total)
		ZDEV_LAUNCH(k_deposit_charge, s->ntiles, 256, 0, s->p, s->tile_off, s->tile_np, d_rho, s->nx + 1, q,
		            s->TX, s->TY, s->ntx);
	if (!moving_window) ZDEV_LAUNCH(k_charge_fold, zdev_div_up(s->ny + 1, 128), 128, 0, d_rho, s->nx, s->ny, 0);
	ZDEV_LAUNCH(k_charge_fold, zdev_div_up(s->nx + 1, 128), 128, 0, d_rho, s->nx, s->ny, 1);
	ZDEV_CHECK(cudaMemcpyAsync(charge, d_rho, n * sizeof(float), cudaMemcpyDeviceToHost, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	cudaFree(d_rho);
}

// ------------------------------------------------------------------ slab decomposition support

extern "C" void zdev_spec2d_export_counts(zdev_spec2d* s, int64_t counts[2]) {
	ctl2d h;
	ZDEV_CHECK(cudaMemcpyAsync(&h, s->ctl, sizeof(h), cudaMemcpyDeviceToHost, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	check_flags(s, h.flags);
	counts[0] = h.n_exp[0]; counts[1] = h.n_exp[1];
}
extern "C" void* zdev_spec2d_export_ptr(zdev_spec2d* s, int side) { return s->exp_buf[side ? 1 : 0]; }

extern "C" void zdev_spec2d_append_device(zdev_spec2d* s, const void* dev_aos, int64_t np) {
	if (np <= 0) return;
	if (!s->cap_total) {
		fprintf(stderr, "(*error*) zdev_spec2d_append_device: species has no tile layout yet (upload or inject first)\n");
		exit(-1);
	}
	spec_append_dev(s, (const part_aos*) dev_aos, np, 0);
	s->np_host += np;
	s->ids_valid = 0;
}
