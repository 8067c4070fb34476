// zpic-b200 :: em2d particle species on the device.
//
// Data layout in HBM.  The grid is cut into tiles of TX x TY cells; every tile owns a
// fixed-capacity segment [tile_off[t], tile_off[t+1]) of slots (a multiple of 32) in
//     rec[]  chunks of 32 slots, each chunk five 128-byte rows  x[32] y[32] ux[32] uy[32] uz[32]:
//            a warp reading 32 neighbouring slots touches one line per field;
//     key[]  16-bit tile-local cell number lx + ly*TX (0xffff = empty slot): the only thing the
//            index sort has to read, and the only place the cell is stored (TX is a power of two:
//            lx = key & (TX-1), ly = key / TX) - 22 bytes per particle in all;
//     tag[]  optional injection index (parity runs),
// of which the first tile_np[t] slots are in use.  There are two such buffers, A and B,
// and every step streams A -> B.
//
// One CTA advances one tile (k_push2d), every thread two particles at a time:
//   phase A  an index-only counting sort of the tile's particles by cell, done in shared
//            memory with native integer atomics: perm[] lists (cell << 16 | slot) of the live
//            slots in cell order (empty slots drop out here, so compaction is free);
//   phase B  each warp walks a contiguous range of perm[], 64 particles per iteration, lane l
//            owning the sorted neighbours 2l and 2l+1.  The two particles travel through the
//            Boris push as the two halves of packed fp32 registers: one FMUL2/FADD2/FFMA2 per
//            operation of the reference's expression tree, each half rounded exactly like the
//            scalar operation (no contraction).  E/B come from a shared-memory tile that stores
//            the four corners of every cell as one float4 per component (6 LDS.128 per particle).
//            Current: the lanes of a warp sit in one or two cells, so the eight contributions
//            are accumulated per lane across iterations and reduced (transposed butterfly, 9
//            shuffles) only when the warp moves on to the next cell.  Particles that cross a cell
//            face (~9 %) go to a warp-private queue and are split/deposited 32 at a time.
//            Survivors are written to B at their sorted position (coalesced 8-byte pairs);
//            particles that left the tile go to the tile's migrants segment (unwrapped global
//            cell indices) and leave an empty slot behind.
// k_migrate2d then applies the boundary conditions to the migrants and appends them to their
// destination tiles in B (or to the slab export lists).
// Per step a particle is read once and written once (2 x 22 B; 56 B with the reference's
// 28-byte record), plus the few percent that migrate.
//
// Replaces reference em2d/particles.c:1104-1269 (spec_advance incl. boundaries and
// spec_move_window's index shift) and :942-1007 (spec_sort - here the cell order is
// rebuilt every step instead of every n_sort steps).
#include "zdev_common.cuh"
#include "pic2d_core.cuh"
#include "pic2d_packed.cuh"
#include "zdev_tma.cuh"
#include "zdev_slab.cuh"
#include <vector>
#include <thread>
#include <algorithm>
#include <cstring>
#include <chrono>
#include <type_traits>

// accessors implemented in zdev_grid2d.cu
f3* zdev_grid2d_Epart(zdev_grid2d* g);
f3* zdev_grid2d_Bpart(zdev_grid2d* g);
f3* zdev_grid2d_J(zdev_grid2d* g);
int zdev_grid2d_nx(zdev_grid2d* g);
int zdev_grid2d_ny(zdev_grid2d* g);

// host AoS record (include/em2d/particles.h t_part)
struct part_aos { int ix, iy; float x, y, ux, uy, uz; };

// one particle as a value (registers); in memory its six words live in a 32-slot chunk, see above
struct rec20 { float x, y, ux, uy, uz; };
#define KEY_EMPTY 0xffffu
#define REC_CHUNK_WORDS 160          // 5 rows x 32 slots

// buffer view handed to kernels by value
struct soa2d {
	float* rec;          // chunked records
	unsigned short* key; // lx + ly*TX, KEY_EMPTY for a hole
	int *tag;            // null unless ids are tracked
};
// migrants: one fixed segment per tile, [tile_off[t]/div, tile_off[t+1]/div), of reference-format records
// carrying UNWRAPPED global cell indices (boundary conditions are applied by k_migrate2d)
struct mig2d {
	part_aos* rec;
	int* tag;            // null unless ids are tracked
	int* np;             // entries written per tile this step
	int div;
};

// word index of slot `slot` (field f is 32*f words further)
__device__ __forceinline__ size_t rec_word(int64_t slot) {
	return (size_t) (slot >> 5) * REC_CHUNK_WORDS + (size_t) (slot & 31);
}
__device__ __forceinline__ rec20 rec_load(const float* __restrict__ rec, int64_t slot) {
	const float* q = rec + rec_word(slot);
	rec20 r; r.x = q[0]; r.y = q[32]; r.ux = q[64]; r.uy = q[96]; r.uz = q[128];
	return r;
}
__device__ __forceinline__ void rec_store(float* __restrict__ rec, int64_t slot, float x, float y, float ux, float uy, float uz) {
	float* q = rec + rec_word(slot);
	q[0] = x; q[32] = y; q[64] = ux; q[96] = uy; q[128] = uz;
}

// control block in device memory, zeroed at the start of every advance
struct ctl2d {
	double energy;                   // sum of utsq/(gamma+1)
	unsigned long long np;           // live particles after the step
	unsigned int n_ovf;              // entries in the overflow list
	unsigned int flags;              // 1: some tile was full (particles parked in the overflow list), 2: a tile's migrants
	                                 // segment overflowed, 4: export list overflow, 8: the overflow list overflowed
	unsigned int n_exp[2];           // slab mode: particles handed to the left / right neighbour
	unsigned int pad[2];
};

struct part_aos;
// linked slabs (zdev_slab.cuh): where k_migrate2d announces its exports, and where k_slab_import finds the imports
struct slab_pub { unsigned* ticket; unsigned* flag[2]; unsigned* count[2]; unsigned seq[2]; };
struct slab_in { const unsigned* flag[2]; const unsigned* count[2]; const part_aos* rec[2]; unsigned seq[2]; };

struct zdev_spec2d {
	int nx, ny;
	int TX, TY, ntx, nty, ntiles;
	int ppc_hint, track_ids;
	double slack;
	int64_t cap_total;               // total SoA slots per buffer in use by the tile layout
	int64_t alloc_total;             // slots allocated per buffer (headroom: tiles grow without reallocation)
	int max_cap;                     // largest tile capacity (sizes perm[] in shared memory)
	soa2d p, q;                      // current (A) and next (B) tile-binned buffers
	int64_t* tile_off;               // device, ntiles+1
	int* tile_np;                    // device, ntiles: slots in use in p
	int* tile_np_q;                  // device, ntiles: slots in use in q
	mig2d mig;                       // per-tile migrants segments
	int64_t mig_cap;                 // total entries allocated in mig.rec
	part_aos* ovf; int* ovf_tag;     // particles that found their destination tile full (global cell indices):
	unsigned int ovf_cap;            //   the host grows the tiles and re-appends them before the next push
	part_aos* ovf_alt; int* ovf_tag_alt;   // the second list: a regrow swaps the two (the parked particles wait in one while
	                                 //   the re-append may park into the other) - no copy, no allocation per event
	int* ev_cnt; int64_t* ev_off;    // device scratch of a regrow event (per-tile counts, the new offsets), kept
	int* ev_host;                    // pinned: 2 x ntiles counts come back here
	int appended;                    // appends since the last overflow check
	int* tile_list;                  // device: ids of the ordinary tiles, then of the few that outgrew them
	int n_small, cap_small;          // how many ordinary tiles, and their largest capacity
	part_aos* stage; int64_t stage_cap;   // persistent staging for appended host particles
	part_aos* exp_buf[2];            // slab mode: export lists (AoS, ix already in the neighbour's frame)
	unsigned int exp_cap;
	ctl2d* ctl;                      // device
	int64_t np_host;                 // last known particle count
	int np_known;                    // np_host is exact (no absorbing boundary / slab exchange since it was counted)
	ctl2d last; int last_valid;      // control block of the last advance as read by the overflow check
	// The step's control block travels to a pinned host copy behind the step's kernels; the host looks at it
	// (full tiles? overflowed lists?) only when it next needs to - normally at the start of the next advance,
	// when the copy has long arrived - so a time step never waits for the stream to drain.
	ctl2d* h_ctl; cudaEvent_t ev_ctl; int ctl_pending;
	// slab decomposition (zdev_slab.cuh): particles that leave through a slab edge are written straight into the
	// neighbour's mailbox by k_migrate2d and appended there by k_slab_import
	int slab; zdev_link link;
	int gx0, gnx;                    // this slab's first column in the whole box, and the box width
	// The import is DEFERRED: the message was sent by the step's k_migrate2d, but nothing in the rest of the step
	// (current, fields, the other species) needs the arrivals, so the wait + append is enqueued at the end of the
	// step (zdev_spec2d_flush_import, or whoever touches the species first): the transfer overlaps that work.
	int import_pending; slab_in pending_in;
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

static void soa_alloc(soa2d& a, int64_t n, int with_tag) {
	size_t nn = (size_t) (n > 0 ? n : 32);           // n is a multiple of 32 (whole chunks)
	memset(&a, 0, sizeof(a));
	ZDEV_CHECK(cudaMalloc(&a.rec, nn * 20));
	ZDEV_CHECK(cudaMalloc(&a.key, nn * 2));
	if (with_tag) ZDEV_CHECK(cudaMalloc(&a.tag, nn * 4));
}
static void soa_free(soa2d& a) {
	cudaFree(a.rec); cudaFree(a.key); cudaFree(a.tag);
	memset(&a, 0, sizeof(a));
}
// migrants segments: 1/div of every tile's capacity (div from the tile shape, mig_div_for below; ZPIC_MIG_DIV
// overrides).  Under a moving window a
// shift empties a whole column of every tile on top of the ordinary leavers, and laser-driven plasma leaves a
// tile at close to one cell per step along a whole edge: div = 2 there.
static void mig_alloc(zdev_spec2d* s, int div);
static void mig_free(zdev_spec2d* s);
static int64_t tile_cap_limit(int TX, int TY);

// supported tile shapes (kernel template instantiations)
static bool tile_supported(int tx, int ty) {
	return (tx == 16 && ty == 16) || (tx == 16 && ty == 8) || (tx == 8 && ty == 8) ||
	       (tx == 8 && ty == 4) || (tx == 4 && ty == 4);
}

extern "C" zdev_spec2d* zdev_spec2d_create(int nx, int ny, int ppc_hint, int track_ids) {
	zdev_require_init();
	{	// the packed-multiply addend (pic2d_packed.cuh)
		static bool negzero_set = false;
		if (!negzero_set) {
			const float2 nz = make_float2(-0.0f, -0.0f);
			ZDEV_CHECK(cudaMemcpyToSymbol(c_negzero2, &nz, sizeof nz));
			negzero_set = true;
		}
	}
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
	ZDEV_CHECK(cudaHostAlloc((void**) &s->h_ctl, sizeof(ctl2d), cudaHostAllocPortable));
	ZDEV_CHECK(cudaEventCreateWithFlags(&s->ev_ctl, cudaEventDisableTiming));
	s->gnx = nx;
	s->h_off = new std::vector<int64_t>();
	return s;
}

static void mig_free(zdev_spec2d* s) {
	cudaFree(s->mig.rec); cudaFree(s->mig.tag); cudaFree(s->mig.np);
	memset(&s->mig, 0, sizeof(s->mig)); s->mig_cap = 0;
}
// A tile's migrants segment is 1/div of the tile's capacity.  No particle moves more than one cell per axis and step
// (v < c and the Courant limit dt < dx, dt < dy), so at most 1 - (1 - 1/TX)(1 - 1/TY) of a tile's particles can
// leave it in one step, whatever the plasma does: the largest div that covers this bound makes the segment
// overflow (flag 2, particles lost) impossible for a stable deck - 8 for 16x16 tiles, 5 for 16x8, 4 for 8x8,
// 2 for 8x4 and 4x4.  A window shift moves one more column of every tile: div 2 (zdev_spec2d_advance).
// $ZPIC_MIG_DIV overrides the choice either way (8 saves 28 B x 7.5 % of the slots on a big 16x8 run).
static int mig_div_for(int TX, int TY) {
	const double leave = 1.0 - (1.0 - 1.0 / TX) * (1.0 - 1.0 / TY);
	const int d = (int) (1.0 / leave);
	return d < 2 ? 2 : (d > 8 ? 8 : d);
}
static void mig_alloc(zdev_spec2d* s, int div) {
	mig_free(s);
	if (const char* e = getenv("ZPIC_MIG_DIV")) { const int v = atoi(e); if (v >= 1 && (v < div || div > 2)) div = v; }
	s->mig.div = div;
	s->mig_cap = s->alloc_total / div + 32;
	ZDEV_CHECK(cudaMalloc(&s->mig.rec, (size_t) s->mig_cap * sizeof(part_aos)));
	if (s->track_ids) ZDEV_CHECK(cudaMalloc(&s->mig.tag, (size_t) s->mig_cap * 4));
	ZDEV_CHECK(cudaMalloc(&s->mig.np, (size_t) s->ntiles * sizeof(int)));
	ZDEV_CHECK(cudaMemsetAsync(s->mig.np, 0, (size_t) s->ntiles * sizeof(int), zdev_strm));
}

static void spec_free_particles(zdev_spec2d* s) {
	for (int k = 0; k < 2; k++) { cudaFree(s->exp_buf[k]); s->exp_buf[k] = nullptr; }
	cudaFree(s->stage); s->stage = nullptr; s->stage_cap = 0;
	if (!s->slab) s->exp_cap = 0;        // (linked slabs export into the neighbours' mailboxes, sized once)
	if (s->cap_total) { soa_free(s->p); soa_free(s->q); }
	mig_free(s);
	cudaFree(s->ovf); cudaFree(s->ovf_tag); s->ovf = nullptr; s->ovf_tag = nullptr; s->ovf_cap = 0;
	cudaFree(s->ovf_alt); cudaFree(s->ovf_tag_alt); s->ovf_alt = nullptr; s->ovf_tag_alt = nullptr;
	cudaFree(s->ev_cnt); cudaFree(s->ev_off); cudaFreeHost(s->ev_host); s->ev_cnt = nullptr; s->ev_off = nullptr; s->ev_host = nullptr;
	s->cap_total = 0; s->alloc_total = 0;
}

extern "C" void zdev_spec2d_destroy(zdev_spec2d* s) {
	if (!s) return;
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	spec_free_particles(s);
	cudaFree(s->tile_off); cudaFree(s->tile_np); cudaFree(s->tile_np_q); cudaFree(s->ctl); cudaFree(s->tile_list);
	cudaFreeHost(s->h_ctl); cudaEventDestroy(s->ev_ctl);
	if (s->slab) zdev_link_close(&s->link);
	if (s->ev) { for (auto& e : *s->ev) cudaEventDestroy(e); delete s->ev; }
	delete s->h_off;
	delete s;
}

extern "C" void zdev_spec2d_tile_info(zdev_spec2d* s, int* tx, int* ty, int* ntiles, int64_t* capacity) {
	if (tx) *tx = s->TX; if (ty) *ty = s->TY; if (ntiles) *ntiles = s->ntiles; if (capacity) *capacity = s->cap_total;
}

// The shared memory of a push CTA is sized by the capacity of the tile it sorts.  A few tiles in a density
// spike can grow to many times the ordinary capacity; sizing every CTA for them would halve the occupancy of
// the whole launch.  The tiles are therefore split into the ordinary ones (capacity up to ~2x the nominal
// fill) and the outgrown ones, pushed by a second small launch with its own shared-memory size.
static void spec_build_tile_lists(zdev_spec2d* s) {
	const std::vector<int64_t>& off = *s->h_off;
	const int64_t limit = (int64_t) (2.2 * s->TX * s->TY * s->ppc_hint) + 128;
	std::vector<int> small, big;
	int cap_small = 32;
	for (int t = 0; t < s->ntiles; t++) {
		const int64_t cap = off[t + 1] - off[t];
		if (cap <= limit) { small.push_back(t); if (cap > cap_small) cap_small = (int) cap; }
		else big.push_back(t);
	}
	s->n_small = (int) small.size();
	s->cap_small = cap_small;
	small.insert(small.end(), big.begin(), big.end());
	if (!s->tile_list) ZDEV_CHECK(cudaMalloc(&s->tile_list, (size_t) s->ntiles * sizeof(int)));
	ZDEV_CHECK(cudaMemcpyAsync(s->tile_list, small.data(), (size_t) s->ntiles * sizeof(int), cudaMemcpyHostToDevice, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
}

// Lay out tile segments for the given per-tile populations and (re)allocate the SoA
// buffers.  Capacity per tile = slack * max(population, nominal fill), rounded to 32.
static void spec_layout(zdev_spec2d* s, const std::vector<int>& cnt, int64_t np) {
	double slack = s->slack;
	// perm[] costs 4 bytes of shared memory per slot of capacity: tiles whose nominal fill is large keep 25 % of
	// room (two push CTAs + 92 KB of L1 per SM), small ones 100 %; full tiles grow on demand either way
	if (slack <= 0.0) {
		int64_t fill = (int64_t) s->TX * s->TY * s->ppc_hint;
		for (int t = 0; t < s->ntiles; t++) if (cnt[t] > fill) fill = cnt[t];
		slack = (fill > 4500 || np > (int64_t) 200000000) ? 1.25 : 2.0;
	}
	std::vector<int64_t>& off = *s->h_off;
	off.assign(s->ntiles + 1, 0);
	int64_t max_cap = 0;
	const int64_t cap_limit = tile_cap_limit(s->TX, s->TY);
	for (int ty = 0; ty < s->nty; ty++) for (int tx = 0; tx < s->ntx; tx++) {
		int t = tx + ty * s->ntx;
		int cx = (tx + 1) * s->TX <= s->nx ? s->TX : s->nx - tx * s->TX;
		int cy = (ty + 1) * s->TY <= s->ny ? s->TY : s->ny - ty * s->TY;
		int64_t nominal = (int64_t) cx * cy * s->ppc_hint;
		int64_t want = cnt[t] > nominal ? cnt[t] : nominal;
		int64_t cap = (int64_t) (want * slack) + 64;
		cap = std::min((cap + 31) & ~(int64_t) 31, cap_limit);
		if (want + 32 > cap) {
			fprintf(stderr, "(*error*) zpic-b200: %lld particles in one %dx%d tile exceed what a push CTA can index and hold in "
			        "shared memory (%lld); use smaller tiles (ZPIC_TILE_X/Y)\n", (long long) want, s->TX, s->TY, (long long) cap_limit);
			exit(-1);
		}
		off[t + 1] = off[t] + cap;
		if (cap > max_cap) max_cap = cap;
	}
	int64_t total = off[s->ntiles];
	spec_free_particles(s);
	// headroom for tiles that grow (a density spike): 1/8 of the layout, at most 64 M slots
	s->alloc_total = (total + std::min<int64_t>(total / 8, (int64_t) 64 << 20) + 65536 + 31) & ~(int64_t) 31;
	soa_alloc(s->p, s->alloc_total, s->track_ids);
	soa_alloc(s->q, s->alloc_total, s->track_ids);
	s->cap_total = total;
	s->max_cap = (int) max_cap;
	mig_alloc(s, mig_div_for(s->TX, s->TY));
	s->ovf_cap = (unsigned int) (total / 32 > (1 << 20) ? total / 32 : (1 << 20));
	ZDEV_CHECK(cudaMalloc(&s->ovf, (size_t) s->ovf_cap * sizeof(part_aos)));
	if (s->track_ids) ZDEV_CHECK(cudaMalloc(&s->ovf_tag, (size_t) s->ovf_cap * 4));
	// the second overflow list and the scratch of the regrow events, here and not at the first event: allocating
	// behind queued work stalled the first event of a run for up to 36 ms (spec_resolve_overflow)
	ZDEV_CHECK(cudaMalloc(&s->ovf_alt, (size_t) s->ovf_cap * sizeof(part_aos)));
	if (s->track_ids) ZDEV_CHECK(cudaMalloc(&s->ovf_tag_alt, (size_t) s->ovf_cap * 4));
	ZDEV_CHECK(cudaMalloc(&s->ev_cnt, (size_t) s->ntiles * sizeof(int)));
	ZDEV_CHECK(cudaMalloc(&s->ev_off, (size_t) (s->ntiles + 1) * sizeof(int64_t)));
	ZDEV_CHECK(cudaHostAlloc((void**) &s->ev_host, (size_t) 2 * s->ntiles * sizeof(int), cudaHostAllocPortable));
	ZDEV_CHECK(cudaMemcpyAsync(s->tile_off, off.data(), (size_t) (s->ntiles + 1) * sizeof(int64_t),
	                           cudaMemcpyHostToDevice, zdev_strm));
	ZDEV_CHECK(cudaMemsetAsync(s->tile_np, 0, (size_t) s->ntiles * sizeof(int), zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	spec_build_tile_lists(s);
}

// ------------------------------------------------------------------ host <-> device

__global__ void k_count_tiles(const part_aos* __restrict__ a, int64_t np, int TX, int TY, int ntx, int* cnt) {
	int64_t k = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= np) return;
	atomicAdd(&cnt[a[k].ix / TX + (a[k].iy / TY) * ntx], 1);
}

// a particle whose destination tile is full: park it (global cell indices) for the host to deal with
__device__ __forceinline__ void ovf_push(ctl2d* ctl, part_aos* ovf, int* ovf_tag, unsigned int cap, const part_aos& r, int tag) {
	atomicOr(&ctl->flags, 1u);
	const unsigned int k = atomicAdd(&ctl->n_ovf, 1u);
	if (k < cap) { ovf[k] = r; if (ovf_tag) ovf_tag[k] = tag; }
	else atomicOr(&ctl->flags, 8u);
}

// append AoS records to their tiles; tag = tags[k] if given, else tag0 + index, when ids are tracked
__global__ void k_scatter_tiles(const part_aos* __restrict__ a, int64_t np, int TX, int TY, int ntx,
                                soa2d p, const int64_t* __restrict__ off, int* tile_np, ctl2d* ctl, int tag0,
                                const int* __restrict__ tags, part_aos* ovf, int* ovf_tag, unsigned int ovf_cap) {
	int64_t k = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= np) return;
	part_aos r = a[k];
	const int tag = tags ? tags[k] : tag0 + (int) k;
	int tx = r.ix / TX, ty = r.iy / TY, t = tx + ty * ntx;
	int slot = atomicAdd(&tile_np[t], 1);
	int64_t d = off[t] + slot;
	if (d >= off[t + 1]) { atomicSub(&tile_np[t], 1); ovf_push(ctl, ovf, ovf_tag, ovf_cap, r, tag); return; }
	const int lx = r.ix - tx * TX, ly = r.iy - ty * TY;
	rec_store(p.rec, d, r.x, r.y, r.ux, r.uy, r.uz);
	p.key[d] = (unsigned short) (lx + ly * TX);
	if (p.tag) p.tag[d] = tag;
}

static void spec_resolve_overflow(zdev_spec2d* s);
// send the control block to the host behind everything enqueued so far (no wait)
static void spec_snapshot_ctl(zdev_spec2d* s) {
	ZDEV_CHECK(cudaMemcpyAsync(s->h_ctl, s->ctl, sizeof(ctl2d), cudaMemcpyDeviceToHost, zdev_strm));
	ZDEV_CHECK(cudaEventRecord(s->ev_ctl, zdev_strm));
	s->ctl_pending = 1;
}
// look at the last snapshot (waits for it if it is still in flight) and deal with full tiles
static void spec_flush_import(zdev_spec2d* s);
static void spec_settle(zdev_spec2d* s) { spec_flush_import(s); if (s->ctl_pending) spec_resolve_overflow(s); }
static void check_flags(zdev_spec2d* s, unsigned int flags) {
	if (flags & 8u) {
		fprintf(stderr, "(*error*) zpic-b200: more than %u particles found their tile full in one step (tile %dx%d cells); "
		        "raise ZPIC_TILE_SLACK (current %.2f) and rerun, aborting.\n", s->ovf_cap, s->TX, s->TY, s->slack);
		exit(-1);
	}
	if (flags & 2u) {
		fprintf(stderr, "(*error*) zpic-b200: a tile's migrants segment overflowed (1/%d of the tile capacity): more particles "
		        "left a %dx%d tile in one step than a time step below the Courant limit allows (or ZPIC_MIG_DIV is set "
		        "too high); set ZPIC_MIG_DIV=1 and rerun, aborting.\n", s->mig.div, s->TX, s->TY);
		exit(-1);
	}
	if (flags & 4u) {
		fprintf(stderr, "(*error*) zpic-b200: slab export list overflow (capacity %u), aborting.\n", s->exp_cap);
		exit(-1);
	}
}

static void spec_append_dev(zdev_spec2d* s, const part_aos* d_aos, int64_t np, int tag0, const int* d_tags = nullptr) {
	if (np <= 0) return;
	ZDEV_LAUNCH(k_scatter_tiles, zdev_div_up(np, 256), 256, 0, d_aos, np, s->TX, s->TY, s->ntx,
	            s->p, s->tile_off, s->tile_np, s->ctl, tag0, d_tags, s->ovf, s->ovf_tag, s->ovf_cap);
	s->appended = 1;
}

// Zero-copy option (ZPIC_ZERO_COPY_MIN=<bytes>): host buffers of at least that size are pinned + mapped once
// and then read / written by the binning and gather kernels directly over PCIe, without a device staging
// copy of the population.  Off by default: measured on B200 (PCIe Gen5) the gather kernel's 28-byte record
// writes into mapped memory run at ~9 GB/s, half the speed of gather-to-HBM + one bulk copy.
static const size_t MAP_MIN_BYTES = ~(size_t) 0;
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
	s->np_host = np; s->np_known = 1;
	s->ids_valid = s->track_ids;
}

// Appends are on the per-step path of a moving window (one injected column per move): the staging
// buffer is kept; the only synchronisation is the overflow check (a 48-byte copy).
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
	spec_snapshot_ctl(s);              // a full tile is dealt with before anybody looks at the population again
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
		const unsigned key = (k < n) ? p.key[b + k] : KEY_EMPTY;
		bool live = key != KEY_EMPTY;
		unsigned m = __ballot_sync(0xffffffffu, live);
		int wbase = 0;
		if ((threadIdx.x & 31) == 0 && m) wbase = atomicAdd(&s_run, __popc(m));
		wbase = __shfl_sync(0xffffffffu, wbase, 0);
		if (live) {
			rec20 v = rec_load(p.rec, b + k);
			part_aos r;
			r.ix = x0 + (int) (key % TX); r.iy = y0 + (int) (key / TX);
			r.x = v.x; r.y = v.y; r.ux = v.ux; r.uy = v.uy; r.uz = v.uz;
			int64_t d = by_tag ? (int64_t) p.tag[b + k] : o + wbase + __popc(m & ((1u << (threadIdx.x & 31)) - 1));
			out[d] = r;
		}
	}
}

// live particles per tile -> host vector; returns the total
static int64_t spec_live_counts(zdev_spec2d* s, std::vector<int>& cnt) {
	spec_settle(s);
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
	s->np_host = np; s->np_known = 1;
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
	s->np_host = spec_live_counts(s, cnt); s->np_known = 1;
	return s->np_host;
}

// ------------------------------------------------------------------ growing tiles

// copy every tile's slots (holes included) from one layout to another
__global__ void k_relayout(soa2d src, const int64_t* __restrict__ off_src, soa2d dst, const int64_t* __restrict__ off_dst,
                           const int* __restrict__ tile_np) {
	const int t = blockIdx.x, n = tile_np[t];
	const int64_t a = off_src[t], b = off_dst[t];
	for (int k = threadIdx.x; k < n; k += blockDim.x) {
		const unsigned short key = src.key[a + k];
		dst.key[b + k] = key;
		if (key != KEY_EMPTY) {
			const rec20 v = rec_load(src.rec, a + k);
			rec_store(dst.rec, b + k, v.x, v.y, v.ux, v.uy, v.uz);
			if (src.tag) dst.tag[b + k] = src.tag[a + k];
		}
	}
}

// Tiles are fixed-capacity segments; a plasma that piles up (the density spike behind a laser pulse is several
// times the initial density) outgrows them.  Particles that find their tile full are parked in the overflow
// list by the append / migrate kernels; here the tiles that need it get 1.5x the room they need now, the
// population is copied to the new layout and the parked particles are appended.  Called (with one small
// device -> host copy) after every advance, so nothing ever misses a push.
// The layout after a regrow event (host only, no device call: exercised on the CPU by tests/test_abi_symbols.py).
// off[] = the ntx*nty+1 offsets of the current layout, np[] / ovf[] = what every tile holds / what is parked for it,
// off_new[] = the new offsets.  A tile that was full gets twice what it needs now and so do the tiles up to two away
// from it (the spike that filled it is moving or widening: they are next - on the LWFA deck 37 events in 200 steps
// without the neighbours, 27 with one ring, 17 with two); every tile within 10 % of its capacity gets 1.5 x its need
// (the 10 % used to be 25 %: with the 1.25 slack of large runs EVERY tile sits at 80 %, so the first event grew them
// all - 340 M -> 444 M slots and a reallocation on the LWFA deck); nothing shrinks, and nothing grows past what a push
// CTA can take (tile_cap_limit).  Returns the largest capacity, or -1 (*bad_tile = which) when a tile NEEDS more than
// that.  x does not wrap (slab edges, moving windows), y is periodic like the box.
extern "C" int64_t zdev_spec2d_plan_regrow(int ntx, int nty, int TX, int TY, const int64_t* off, const int* np, const int* ovf,
                                           int64_t* off_new, int* bad_tile) {
	const int ntiles = ntx * nty;
	const int64_t cap_limit = tile_cap_limit(TX, TY);
	std::vector<int64_t> want(ntiles);
	for (int t = 0; t < ntiles; t++) {
		const int64_t cap = off[t + 1] - off[t], need = (int64_t) np[t] + ovf[t];
		want[t] = (need > cap - cap / 10) ? std::max(cap, need + need / 2 + 256) : cap;
	}
	for (int t = 0; t < ntiles; t++) {
		if (ovf[t] <= 0) continue;
		const int64_t grown = 2 * ((int64_t) np[t] + ovf[t]) + 256;
		const int tx = t % ntx, ty = t / ntx;
		for (int dy = -2; dy <= 2; dy++) for (int dx = -2; dx <= 2; dx++) {
			const int ux = tx + dx, uy = (ty + dy + 2 * nty) % nty;
			if (ux < 0 || ux >= ntx) continue;
			int64_t& w = want[ux + uy * ntx];
			if (grown > w) w = grown;
		}
	}
	int64_t max_cap = 0;
	off_new[0] = 0;
	for (int t = 0; t < ntiles; t++) {
		const int64_t need = (int64_t) np[t] + ovf[t];
		const int64_t cap = std::min((want[t] + 31) & ~(int64_t) 31, cap_limit);
		if (need + 32 > cap) { if (bad_tile) *bad_tile = t; return -1; }
		off_new[t + 1] = off_new[t] + cap;
		if (cap > max_cap) max_cap = cap;
	}
	return max_cap;
}

static void spec_resolve_overflow(zdev_spec2d* s) {
	s->appended = 0;
	ctl2d h;
	if (s->ctl_pending) {
		ZDEV_CHECK(cudaEventSynchronize(s->ev_ctl));
		h = *s->h_ctl;
		s->ctl_pending = 0;
	} else {
		ZDEV_CHECK(cudaMemcpyAsync(&h, s->ctl, sizeof h, cudaMemcpyDeviceToHost, zdev_strm));
		ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	}
	while (h.flags & 1u) {
		const auto t_start = std::chrono::steady_clock::now();
		double t_ph[5] = {0, 0, 0, 0, 0};
		auto mark = [&](int k) { t_ph[k] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(); };
		check_flags(s, h.flags & (2u | 4u | 8u));
		const int64_t n_ovf = h.n_ovf;
		// the scratch of the events is kept (a cudaMalloc / cudaFree pair behind queued work costs 1 - 4 ms each: measured
		// on the LWFA deck, where a tile outgrows its segment every ~10 steps); normally allocated with the layout
		if (!s->ev_cnt) {
			ZDEV_CHECK(cudaMalloc(&s->ev_cnt, (size_t) s->ntiles * sizeof(int)));
			ZDEV_CHECK(cudaMalloc(&s->ev_off, (size_t) (s->ntiles + 1) * sizeof(int64_t)));
			ZDEV_CHECK(cudaHostAlloc((void**) &s->ev_host, (size_t) 2 * s->ntiles * sizeof(int), cudaHostAllocPortable));
		}
		if (!s->ovf_alt) {
			ZDEV_CHECK(cudaMalloc(&s->ovf_alt, (size_t) s->ovf_cap * sizeof(part_aos)));
			if (s->track_ids) ZDEV_CHECK(cudaMalloc(&s->ovf_tag_alt, (size_t) s->ovf_cap * 4));
		}
		// what every tile holds and what is waiting for it
		int* d_cnt = s->ev_cnt;
		const int* ovf_t = s->ev_host; const int* np_t = s->ev_host + s->ntiles;
		ZDEV_CHECK(cudaMemsetAsync(d_cnt, 0, (size_t) s->ntiles * sizeof(int), zdev_strm));
		ZDEV_LAUNCH(k_count_tiles, zdev_div_up(n_ovf, 256), 256, 0, s->ovf, n_ovf, s->TX, s->TY, s->ntx, d_cnt);
		ZDEV_CHECK(cudaMemcpyAsync(s->ev_host, d_cnt, (size_t) s->ntiles * sizeof(int), cudaMemcpyDeviceToHost, zdev_strm));
		ZDEV_CHECK(cudaMemcpyAsync(s->ev_host + s->ntiles, s->tile_np, (size_t) s->ntiles * sizeof(int), cudaMemcpyDeviceToHost, zdev_strm));
		ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
		mark(0);
		std::vector<int64_t> off_new(s->ntiles + 1, 0);
		int bad_tile = -1;
		const int64_t max_cap = zdev_spec2d_plan_regrow(s->ntx, s->nty, s->TX, s->TY, s->h_off->data(), np_t, ovf_t, off_new.data(), &bad_tile);
		if (max_cap < 0) {
			fprintf(stderr, "(*error*) zpic-b200: %lld particles in one %dx%d tile exceed what a push CTA can index and hold in "
			        "shared memory (%lld); use smaller tiles (ZPIC_TILE_X/Y)\n", (long long) np_t[bad_tile] + ovf_t[bad_tile],
			        s->TX, s->TY, (long long) tile_cap_limit(s->TX, s->TY));
			exit(-1);
		}
		const int64_t total = off_new[s->ntiles];
		const int64_t slots_before = s->cap_total;
		// the parked particles wait in their list while the other list takes over (appending them may park others again)
		part_aos* d_wait = s->ovf; int* d_wait_tag = s->ovf_tag;
		s->ovf = s->ovf_alt; s->ovf_tag = s->ovf_tag_alt;
		s->ovf_alt = d_wait; s->ovf_tag_alt = d_wait_tag;
		int64_t* d_off_new = s->ev_off;
		ZDEV_CHECK(cudaMemcpyAsync(d_off_new, off_new.data(), (size_t) (s->ntiles + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, zdev_strm));
		mark(1);
		if (total <= s->alloc_total) {
			// the new layout fits the allocation: population -> the other buffer (scratch between steps), swap
			ZDEV_LAUNCH(k_relayout, s->ntiles, 256, 0, s->p, s->tile_off, s->q, d_off_new, s->tile_np);
			ZDEV_CHECK(cudaMemcpyAsync(s->tile_off, d_off_new, (size_t) (s->ntiles + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, zdev_strm));
			{ soa2d t = s->p; s->p = s->q; s->q = t; }
			ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));        // (off_new, a host vector, was the source of an async copy)
		} else {
			// population -> new, larger buffers (q is scratch between steps: release it first)
			ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
			soa_free(s->q);
			s->alloc_total = (total + std::min<int64_t>(total / 8, (int64_t) 64 << 20) + 65536 + 31) & ~(int64_t) 31;
			soa2d pn;
			soa_alloc(pn, s->alloc_total, s->track_ids);
			ZDEV_LAUNCH(k_relayout, s->ntiles, 256, 0, s->p, s->tile_off, pn, d_off_new, s->tile_np);
			ZDEV_CHECK(cudaMemcpyAsync(s->tile_off, d_off_new, (size_t) (s->ntiles + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, zdev_strm));
			ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
			soa_free(s->p);
			s->p = pn;
			soa_alloc(s->q, s->alloc_total, s->track_ids);
			mig_alloc(s, s->mig.div);
		}
		mark(2);
		*s->h_off = off_new;
		s->cap_total = total;
		s->max_cap = (int) max_cap;
		spec_build_tile_lists(s);
		mark(3);
		// clear the overflow state (energy, counts and export counters of the step stay) and append
		ZDEV_CHECK(cudaMemsetAsync(&s->ctl->n_ovf, 0, 2 * sizeof(unsigned int), zdev_strm));    // n_ovf, flags
		spec_append_dev(s, d_wait, n_ovf, 0, d_wait_tag);
		s->appended = 0;
		ZDEV_CHECK(cudaMemcpyAsync(&h, s->ctl, sizeof h, cudaMemcpyDeviceToHost, zdev_strm));
		ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
		if (getenv("ZPIC_VERBOSE"))
			fprintf(stderr, "zpic-b200: %lld particles found their %dx%d tile full: slots %lld -> %lld, largest tile %d, "
			        "%d outgrown tiles, %.1f ms (counts %.1f, scratch %.1f, relayout %.1f, lists %.1f)\n", (long long) n_ovf, s->TX, s->TY,
			        (long long) slots_before, (long long) total, s->max_cap, s->ntiles - s->n_small,
			        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(),
			        t_ph[0], t_ph[1] - t_ph[0], t_ph[2] - t_ph[1], t_ph[3] - t_ph[2]);
	}
	s->last = h; s->last_valid = 1;
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
// 335-347), momenta as spec_set_u (thermal, minus the cell mean, plus fluid; :96-142).
// Only the cells of the rectangle [ix0, ix1) x [iy0, iy1) get particles; a tile packs its part of it row by row
// from its first slot.
__global__ void k_inject_uniform(soa2d p, const int64_t* __restrict__ off, int* tile_np,
                                 int nx, int ny, int TX, int TY, int ntx, int ppcx, int ppcy,
                                 f3 ufl, f3 uth, uint64_t seed, int ix0, int ix1, int iy0, int iy1, int gx0, int gnx) {
	int64_t cell = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (cell >= (int64_t) nx * ny) return;
	int iy = (int) (cell / nx), ix = (int) (cell - (int64_t) iy * nx);
	if (iy < iy0 || iy >= iy1 || ix < ix0 || ix >= ix1) return;       // outside the rectangle: no particles
	int tx = ix / TX, ty = iy / TY, t = tx + ty * ntx;
	int cx = (tx + 1) * TX <= nx ? TX : nx - tx * TX;
	int lx = ix - tx * TX, ly = iy - ty * TY;
	int npc = ppcx * ppcy;
	const int lx0 = max(ix0 - tx * TX, 0), ly0 = max(iy0 - ty * TY, 0);
	const int w = min(tx * TX + cx, ix1) - max(tx * TX, ix0);          // cells of a tile row inside the rectangle
	int64_t base = off[t] + (int64_t) ((lx - lx0) + (ly - ly0) * w) * npc;
	uint64_t gid0 = ((uint64_t) (gx0 + ix) + (uint64_t) gnx * iy) * npc;     // numbered by cell of the WHOLE box
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
		rec_store(p.rec, d, (float) (dpcx * (kx + 0.5)), (float) (dpcy * (ky + 0.5)),
		          uth.x * a + (ufl.x - sx), uth.y * b + (ufl.y - sy), uth.z * c + (ufl.z - sz));
		p.key[d] = (unsigned short) (lx + ly * TX);
		if (p.tag) p.tag[d] = (int) (gid0 + k);
	}
	if (lx == lx0 && ly == ly0) {
		int cy = (ty + 1) * TY <= ny ? TY : ny - ty * TY;
		const int rows = min(ty * TY + cy, iy1) - max(ty * TY, iy0);
		tile_np[t] = w * rows * npc;
	}
}

// uniform plasma in the cells [ix0, ix1) x [iy0, iy1) only (the half-box species of a shear-flow deck, a plasma
// that starts at some x); tiles outside get the minimum capacity (they grow on demand when particles arrive)
extern "C" void zdev_spec2d_inject_rect(zdev_spec2d* s, int ppcx, int ppcy, const float ufl[3], const float uth[3], uint64_t seed,
                                        int ix0, int ix1, int iy0, int iy1) {
	int npc = ppcx * ppcy;
	ix0 = std::max(ix0, 0); iy0 = std::max(iy0, 0);
	ix1 = std::min(ix1, s->nx); iy1 = std::min(iy1, s->ny);
	std::vector<int> cnt(s->ntiles);
	int64_t np = 0;
	for (int ty = 0; ty < s->nty; ty++) for (int tx = 0; tx < s->ntx; tx++) {
		int cx = (tx + 1) * s->TX <= s->nx ? s->TX : s->nx - tx * s->TX;
		int cy = (ty + 1) * s->TY <= s->ny ? s->TY : s->ny - ty * s->TY;
		int rows = std::min(ty * s->TY + cy, iy1) - std::max(ty * s->TY, iy0);
		int cols = std::min(tx * s->TX + cx, ix1) - std::max(tx * s->TX, ix0);
		if (rows < 0) rows = 0;
		if (cols < 0) cols = 0;
		cnt[tx + ty * s->ntx] = cols * rows * npc; np += (int64_t) cols * rows * npc;
	}
	// a species that fills only part of the box (a half-box beam) gets its capacity where it is, not the nominal
	// fill everywhere; one that merely starts a little inside the box keeps the nominal layout (under a moving
	// window the empty tiles fill up within a few steps, and growing tiles step after step is expensive)
	const bool partial = (double) np < 0.9 * (double) s->nx * s->ny * npc;
	const int hint = s->ppc_hint;
	if (partial) s->ppc_hint = 0;               // capacities from the actual populations, not the nominal fill
	spec_layout(s, cnt, np);
	s->ppc_hint = hint;
	if (partial) spec_build_tile_lists(s);
	f3 fl = {ufl[0], ufl[1], ufl[2]}, th = {uth[0], uth[1], uth[2]};
	int64_t ncell = (int64_t) s->nx * s->ny;
	ZDEV_CHECK(cudaMemsetAsync(s->tile_np, 0, (size_t) s->ntiles * sizeof(int), zdev_strm));
	if (np > 0)
		ZDEV_LAUNCH(k_inject_uniform, zdev_div_up(ncell, 128), 128, 0, s->p, s->tile_off, s->tile_np,
		            s->nx, s->ny, s->TX, s->TY, s->ntx, ppcx, ppcy, fl, th, seed, ix0, ix1, iy0, iy1, s->gx0, s->gnx);
	s->np_host = np; s->np_known = 1;
	s->ids_valid = s->track_ids && np < 0x7fffffff;
}

extern "C" void zdev_spec2d_inject_band(zdev_spec2d* s, int ppcx, int ppcy, const float ufl[3], const float uth[3], uint64_t seed,
                                        int iy0, int iy1) {
	zdev_spec2d_inject_rect(s, ppcx, ppcy, ufl, uth, seed, 0, s->nx, iy0, iy1);
}

extern "C" void zdev_spec2d_inject_uniform(zdev_spec2d* s, int ppcx, int ppcy, const float ufl[3], const float uth[3], uint64_t seed) {
	zdev_spec2d_inject_rect(s, ppcx, ppcy, ufl, uth, seed, 0, s->nx, 0, s->ny);
}

// ---- the same population as the reference's host injector, on the reference random stream (zdev_refrng.cu)
// One thread per cell of the rows [j0, j1).  Column i of the box holds the in-cell x positions kx_lo <= kx < kx_hi
// (STEP / SLAB clip them, particles.c:335-347), pre[] = particles per row in the columns before i, R per row; the
// cell's particles are numbered (iy R + pre[i] + ky w + kx - kx_lo) like spec_set_x's loops order them, and
// th[] holds the three thermal components of the rows' particles in that order (null: a cold plasma).
__global__ void k_inject_lattice(soa2d p, const int64_t* __restrict__ off, int* tile_np, int nx, int ny, int TX, int TY, int ntx,
                                 int ppcx, int ppcy, f3 ufl, const int* __restrict__ kx_lo, const int* __restrict__ kx_hi,
                                 const int* __restrict__ pre, int gx0, int64_t R, int j0, int j1, const float* __restrict__ th) {
	const int64_t c = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= (int64_t) nx * (j1 - j0)) return;
	const int iy = j0 + (int) (c / nx), ix = (int) (c % nx), g = gx0 + ix;
	const int tx = ix / TX, ty = iy / TY, t = tx + ty * ntx;
	const int cxw = (tx + 1) * TX <= nx ? TX : nx - tx * TX, cyw = (ty + 1) * TY <= ny ? TY : ny - ty * TY;
	const int lx = ix - tx * TX, ly = iy - ty * TY;
	const int gt0 = gx0 + tx * TX;
	const int rowcnt = pre[gt0 + cxw] - pre[gt0];
	if (lx == 0 && ly == 0) tile_np[t] = rowcnt * cyw;
	const int lo = kx_lo[g], w = kx_hi[g] - lo;
	if (w <= 0) return;
	const int cnt = w * ppcy;
	const int64_t base = off[t] + (int64_t) ly * rowcnt + (pre[g] - pre[gt0]);
	const int64_t n0 = (int64_t) iy * R + pre[g];                    // injection index of the cell's first particle
	const float* q = th ? th + 3 * ((int64_t) (iy - j0) * R + pre[g]) : nullptr;
	float sx = 0, sy = 0, sz = 0;
	if (q) {
		for (int k = 0; k < cnt; k++) { sx += q[3 * k]; sy += q[3 * k + 1]; sz += q[3 * k + 2]; }
		const float norm = 1.0f / cnt;
		sx *= norm; sy *= norm; sz *= norm;
	}
	const float dpcx = 1.0f / ppcx, dpcy = 1.0f / ppcy;
	for (int k = 0; k < cnt; k++) {
		const int ky = k / w, kx = lo + k - ky * w;
		const float tx_ = q ? q[3 * k] : 0.0f, ty_ = q ? q[3 * k + 1] : 0.0f, tz_ = q ? q[3 * k + 2] : 0.0f;
		const int64_t d = base + k;
		rec_store(p.rec, d, (float) (dpcx * (kx + 0.5)), (float) (dpcy * (ky + 0.5)),
		          tx_ + (ufl.x - sx), ty_ + (ufl.y - sy), tz_ + (ufl.z - sz));
		p.key[d] = (unsigned short) (lx + ly * TX);
		if (p.tag) p.tag[d] = (int) (n0 + k);
	}
}

extern "C" int zdev_spec2d_inject_lattice(zdev_spec2d* s, int ppcx, int ppcy, const float ufl[3], const float uth[3],
                                          const int* kx_lo, const int* kx_hi,
                                          uint32_t* z, uint32_t* w, int* have_spare, double* spare) {
	const int gnx = s->gnx > 0 ? s->gnx : s->nx, gx0 = s->gnx > 0 ? s->gx0 : 0;
	std::vector<int> pre(gnx + 1, 0);
	for (int i = 0; i < gnx; i++) pre[i + 1] = pre[i] + std::max(kx_hi[i] - kx_lo[i], 0) * ppcy;
	const int64_t R = pre[gnx];
	const int64_t total = R * s->ny;                               // the whole box: the stream is the whole box's
	const bool cold = uth[0] == 0.0f && uth[1] == 0.0f && uth[2] == 0.0f;
	{	// the state must be usable before anything is laid out
		uint32_t zz = *z, ww = *w;
		if (zdev_ref_jump(&zz, &ww, 0)) return 1;
	}
	std::vector<int> cnt(s->ntiles);
	int64_t np = 0;
	for (int ty = 0; ty < s->nty; ty++) for (int tx = 0; tx < s->ntx; tx++) {
		const int cx = (tx + 1) * s->TX <= s->nx ? s->TX : s->nx - tx * s->TX;
		const int cy = (ty + 1) * s->TY <= s->ny ? s->TY : s->ny - ty * s->TY;
		const int64_t c = (int64_t) (pre[gx0 + tx * s->TX + cx] - pre[gx0 + tx * s->TX]) * cy;
		cnt[tx + ty * s->ntx] = (int) c; np += c;
	}
	const bool partial = (double) np < 0.9 * (double) s->nx * s->ny * ppcx * ppcy;
	const int hint = s->ppc_hint;
	if (partial) s->ppc_hint = 0;
	spec_layout(s, cnt, np);
	s->ppc_hint = hint;
	if (partial) spec_build_tile_lists(s);
	ZDEV_CHECK(cudaMemsetAsync(s->tile_np, 0, (size_t) s->ntiles * sizeof(int), zdev_strm));
	int *d_lo, *d_hi, *d_pre;
	ZDEV_CHECK(cudaMalloc(&d_lo, (size_t) gnx * sizeof(int)));
	ZDEV_CHECK(cudaMalloc(&d_hi, (size_t) gnx * sizeof(int)));
	ZDEV_CHECK(cudaMalloc(&d_pre, (size_t) (gnx + 1) * sizeof(int)));
	ZDEV_CHECK(cudaMemcpyAsync(d_lo, kx_lo, (size_t) gnx * sizeof(int), cudaMemcpyHostToDevice, zdev_strm));
	ZDEV_CHECK(cudaMemcpyAsync(d_hi, kx_hi, (size_t) gnx * sizeof(int), cudaMemcpyHostToDevice, zdev_strm));
	ZDEV_CHECK(cudaMemcpyAsync(d_pre, pre.data(), (size_t) (gnx + 1) * sizeof(int), cudaMemcpyHostToDevice, zdev_strm));
	const f3 fl = {ufl[0], ufl[1], ufl[2]};
	if (cold || R == 0) {
		// 0 * deviate: only the stream position matters (3 deviates per particle all the same, particles.c:97-101)
		zdev_ref_normals(z, w, have_spare, spare, 3 * total, uth, nullptr);
		if (np > 0)
			ZDEV_LAUNCH(k_inject_lattice, zdev_div_up((int64_t) s->nx * s->ny, 128), 128, 0, s->p, s->tile_off, s->tile_np,
			            s->nx, s->ny, s->TX, s->TY, s->ntx, ppcx, ppcy, fl, d_lo, d_hi, d_pre, gx0, R, 0, s->ny, (const float*) nullptr);
	} else {
		// bands of rows: the thermal components of a band are generated into a staging array, then placed
		const int rows = (int) std::max<int64_t>(1, std::min<int64_t>(s->ny, ((int64_t) 1 << 26) / R));
		float* d_th;
		ZDEV_CHECK(cudaMalloc(&d_th, (size_t) rows * R * 3 * sizeof(float)));
		for (int j0 = 0; j0 < s->ny; j0 += rows) {
			const int j1 = std::min(j0 + rows, s->ny);
			zdev_ref_normals(z, w, have_spare, spare, 3 * R * (j1 - j0), uth, d_th);
			if (np > 0)
				ZDEV_LAUNCH(k_inject_lattice, zdev_div_up((int64_t) s->nx * (j1 - j0), 128), 128, 0, s->p, s->tile_off, s->tile_np,
				            s->nx, s->ny, s->TX, s->TY, s->ntx, ppcx, ppcy, fl, d_lo, d_hi, d_pre, gx0, R, j0, j1, (const float*) d_th);
		}
		ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
		ZDEV_CHECK(cudaFree(d_th));
	}
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	ZDEV_CHECK(cudaFree(d_lo)); ZDEV_CHECK(cudaFree(d_hi)); ZDEV_CHECK(cudaFree(d_pre));
	s->np_host = np; s->np_known = 1;
	s->ids_valid = s->track_ids && total < 0x7fffffff && np == total;
	return 0;
}

// The moving window's new column (cells ix, iy0 <= iy < iy1) generated on the device like k_inject_uniform does and
// appended to its tiles: one thread per cell reserves the cell's slots with one atomic.  `col` numbers the column
// (cells that ever entered the box get distinct particle numbers).
__global__ void k_inject_column(soa2d p, const int64_t* __restrict__ off, int* tile_np, ctl2d* ctl, int TX, int TY, int ntx,
                                int ppcx, int ppcy, f3 ufl, f3 uth, uint64_t seed, int ix, int iy0, int iy1, uint64_t col,
                                part_aos* ovf, int* ovf_tag, unsigned int ovf_cap) {
	const int iy = iy0 + blockIdx.x * blockDim.x + threadIdx.x;
	if (iy >= iy1) return;
	const int npc = ppcx * ppcy;
	const uint64_t gid0 = (col * 0x100000ull + (uint64_t) iy) * npc + 0x4000000000000000ull;
	float sx = 0, sy = 0, sz = 0;
	for (int k = 0; k < npc; k++) {
		float a, b, c; normal3(seed, gid0 + k, a, b, c);
		sx += uth.x * a; sy += uth.y * b; sz += uth.z * c;
	}
	const float norm = 1.0f / npc;
	sx *= norm; sy *= norm; sz *= norm;
	const int tx = ix / TX, ty = iy / TY, t = tx + ty * ntx;
	const int lx = ix - tx * TX, ly = iy - ty * TY;
	const int slot0 = atomicAdd(&tile_np[t], npc);
	const float dpcx = 1.0f / ppcx, dpcy = 1.0f / ppcy;
	const int64_t room = off[t + 1] - off[t];
	int parked = 0;
	for (int k = 0; k < npc; k++) {
		float a, b, c; normal3(seed, gid0 + k, a, b, c);
		const int kx = k % ppcx, ky = k / ppcx;
		const float x = (float) (dpcx * (kx + 0.5)), y = (float) (dpcy * (ky + 0.5));
		const float ux = uth.x * a + (ufl.x - sx), uy = uth.y * b + (ufl.y - sy), uz = uth.z * c + (ufl.z - sz);
		if (slot0 + k < room) {
			const int64_t d = off[t] + slot0 + k;
			rec_store(p.rec, d, x, y, ux, uy, uz);
			p.key[d] = (unsigned short) (lx + ly * TX);
			if (p.tag) p.tag[d] = 0;
		} else {
			part_aos r; r.ix = ix; r.iy = iy; r.x = x; r.y = y; r.ux = ux; r.uy = uy; r.uz = uz;
			ovf_push(ctl, ovf, ovf_tag, ovf_cap, r, 0);
			parked++;
		}
	}
	if (parked) atomicSub(&tile_np[t], parked);      // (the tile is full: nobody else got a slot behind ours)
}

extern "C" void zdev_spec2d_inject_column(zdev_spec2d* s, int ppcx, int ppcy, const float ufl[3], const float uth[3], uint64_t seed,
                                          int ix, int iy0, int iy1, uint64_t column_number) {
	if (!s->cap_total || iy1 <= iy0) return;
	f3 fl = {ufl[0], ufl[1], ufl[2]}, th = {uth[0], uth[1], uth[2]};
	ZDEV_LAUNCH(k_inject_column, zdev_div_up(iy1 - iy0, 128), 128, 0, s->p, s->tile_off, s->tile_np, s->ctl, s->TX, s->TY, s->ntx,
	            ppcx, ppcy, fl, th, seed, ix, iy0, iy1, column_number, s->ovf, s->ovf_tag, s->ovf_cap);
	s->appended = 1;
	s->np_host += (int64_t) (iy1 - iy0) * ppcx * ppcy;
	s->ids_valid = 0;
	spec_snapshot_ctl(s);
}

// ------------------------------------------------------------------ the push

struct push_geom {
	int nx, ny, nrow;        // grid
	int ntx;                 // tiles per row
};

// Where the current goes: the global J grid itself, through L2 float reductions (RED.ADD.F32).  v points
// at the tile's cell (-1,-1), W3 = 3*nrow; component k of tile cell (lx,ly) = (lx+1)*3 + (ly+1)*W3 + k.
// (Measured alternatives on B200: a shared-memory tile of floats or of 64-bit fixed point compiles to
// compare-and-swap loops, 34 vs 50 Gpush/s; a tile of 32-bit coarse+fine fixed-point pairs with native
// integer atomics 53 vs 56 Gpush/s.)
struct jtile { float* v; };
#ifdef ABL_NO_RED      // ablation builds (timing experiments only, results are wrong): ABL_NO_RED, ABL_NO_QUEUE, ABL_NO_DEPOSIT
__device__ __forceinline__ void jt_add(const jtile& t, int idx, float w) { if (w == 12345.678f) atomicAdd(t.v + idx, w); }
#else
__device__ __forceinline__ void jt_add(const jtile& t, int idx, float w) { atomicAdd(t.v + idx, w); }
#endif
// the 8 contributions of one segment in cell c (index of its first component), W3 = 3 * row length
__device__ __forceinline__ void jt_weights(const jtile& t, int c, int W3, const float w[8]) {
	jt_add(t, c, w[0]);
	jt_add(t, c + W3, w[1]);
	jt_add(t, c + 1, w[2]);
	jt_add(t, c + 3 + 1, w[3]);
	jt_add(t, c + 2, w[4]);
	jt_add(t, c + 3 + 2, w[5]);
	jt_add(t, c + W3 + 2, w[6]);
	jt_add(t, c + W3 + 3 + 2, w[7]);
}
__device__ __forceinline__ void deposit_seg(const jtile& t, int W3, const seg2d& s, float qnx, float qny) {
	float w[8];
	seg_weights(s, qnx, qny, w);
	jt_weights(t, (s.ix + 1) * 3 + (s.iy + 1) * W3, W3, w);
}

// one queued move (warp-private shared-memory queue): a move that leaves its cell, as the push computed it.
// cij = cell key | (di+1) << 16 | (dj+1) << 18.  The lanes have deposited its first in-cell piece; the drain
// recomputes where that piece ended (the same operations on the same operands, so the pieces join exactly) and
// deposits the rest.  Queueing the raw move instead of the remainder keeps the enqueue - which runs for every
// 64 particles although only ~6 of them cross - down to a few instructions; the arithmetic runs in the drain,
// 32 entries at a time.
struct __align__(8) xq_entry { int cij; float x, y, dx, dy, qvz; };

// split + deposit up to 32 queued moves, one per lane
template <int TX>
__device__ __forceinline__ void drain_queue(const xq_entry* q, int n, int lane, const jtile& t, int W3,
                                            float qnx, float qny) {
	if (lane < n) {
		const xq_entry e = q[lane];
		const int key = e.cij & 0xffff, di = ((e.cij >> 16) & 3) - 1, dj = ((e.cij >> 18) & 3) - 1;
		// end of the first piece: the scalar twins of the packed operations of the push (same roundings)
		const float fx = di > 0 ? 1.0f : 0.0f, fy = dj > 0 ? 1.0f : 0.0f;
		float tx = (fx - e.x) * rcp_approx1(e.dx), ty = (fy - e.y) * rcp_approx1(e.dy);
		tx = di ? tx : 2.0f; ty = dj ? ty : 2.0f;
		const float t1 = fmaxf(fminf(fminf(tx, ty), 1.0f), 0.0f);
		const bool xf = di != 0 && tx <= ty, yf = dj != 0 && !xf;
		const float xe = e.x + e.dx * t1, ye = e.y + e.dy * t1, r1 = 1.0f - t1;
		seg2d vp[2];
		int vnp = split_once((key & (TX - 1)) + (xf ? di : 0), key / TX + (yf ? dj : 0), xf ? 0 : di, yf ? 0 : dj,
		                     xf ? 1.0f - fx : xe, yf ? 1.0f - fy : ye, e.dx * r1, e.dy * r1, e.qvz * r1, vp);
		deposit_seg(t, W3, vp[0], qnx, qny);
		if (vnp > 1) deposit_seg(t, W3, vp[1], qnx, qny);
	}
}

// Sum acc[0..7] over the 32 lanes (transposed butterfly: 9 shuffles for the 8 sums), after which lane 4*k
// holds the total of contribution k, and add the totals to cell `cell` (= lx + ly*TX) of the tile.
template <int TX>
__device__ __forceinline__ void flush_cell(const float acc[8], int cell, int lane, const jtile& t, int W3) {
	float v4[4], v2[2], v1;
	const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4;
	#pragma unroll
	for (int q = 0; q < 4; q++) {
		float send = b16 ? acc[q] : acc[q + 4], keep = b16 ? acc[q + 4] : acc[q];
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
		const int k = lane >> 2;
		const int comp = (k < 2) ? 0 : ((k < 4) ? 1 : 2);
		const int right = (k == 3) | (k == 5) | (k == 7);
		const int up = (k == 1) | (k == 6) | (k == 7);
		jt_add(t, ((cell & (TX - 1)) + 1) * 3 + (cell / TX + 1) * W3 + comp + right * 3 + up * W3, v1);
	}

}

// the six words of the sorted neighbours (pa, pa+1), as loaded from their source slots
struct pair_rec { f2 x, y, ux, uy, uz; int ca, cb, ta, tb; };      // ca, cb: cell numbers (keys)

#ifndef XQ_CAP_N
#define XQ_CAP_N 96                  // 31 left over + 64 new entries at most
#endif
static const int XQ_CAP = XQ_CAP_N;

// dynamic shared memory of k_push2d: [front][perm], front = the tile's keys during the sort, afterwards the
// corner tile + the warps' queues in the same bytes; the raw field planes (staged before the sort, dead once the
// corner tile is built, i.e. before the first queue entry is written) sit in the tail of the queue area, behind
// the keys; perm[] = 32-bit (cell << 16 | slot) entries.
struct push_smem { size_t front, raw, total; };
static push_smem push_smem_layout(int TX, int TY, int max_cap) {
	const size_t plane = (size_t) (TX + 2) * (TY + 2);
	const size_t corner = 6 * plane * 16, queues = (size_t) PUSH_WARPS * XQ_CAP * sizeof(xq_entry);
	// + 128 keys: the sort walks 64 * S >= n keys and the ones past n are padded with KEY_EMPTY in place
	const size_t keys = ((size_t) (max_cap + 128) * 2 + 15) & ~(size_t) 15;
	push_smem L;
	L.raw = keys > corner ? keys : corner;
	L.front = corner + queues > L.raw + 6 * plane * 4 ? corner + queues : L.raw + 6 * plane * 4;
	L.front = (L.front + 15) & ~(size_t) 15;
	L.total = L.front + (size_t) (max_cap + 128) * 4;     // the holes are sorted too (behind the live entries)
	return L;
}
// the largest tile capacity a push CTA can take: 16-bit slot numbers in perm[], and its shared memory (227 KB per CTA,
// of which up to 3.1 KB are the kernel's static arrays)
static int64_t tile_cap_limit(int TX, int TY) {
	int64_t lo = 32, hi = 0xfff0 & ~31;
	while (lo < hi) {
		const int64_t mid = ((lo + hi) / 2 + 31) & ~(int64_t) 31;
		if (push_smem_layout(TX, TY, (int) mid).total <= (size_t) 223 * 1024) lo = mid; else hi = mid - 32;
	}
	return lo;
}

// One CTA per tile.
template <int TX, int TY, bool TAGS>
__global__ void __launch_bounds__(PUSH_THREADS, PUSH_MIN_BLOCKS)
k_push2d(soa2d A, soa2d Bo, const int64_t* __restrict__ tile_off, const int* __restrict__ tile_np,
         int* __restrict__ tile_np_out, mig2d mig,
         ctl2d* __restrict__ ctl, const f3* __restrict__ E, const f3* __restrict__ B, f3* __restrict__ J,
         push_geom g, zdev_push2d_params prm, unsigned smem_front, unsigned smem_raw, const int* __restrict__ tile_list) {
	constexpr int SROW = TX + 2;
	constexpr int PLANE = SROW * (TY + 2);
	constexpr int NC = TX * TY;
	static_assert((TX & (TX - 1)) == 0, "TX must be a power of two");
	extern __shared__ __align__(16) unsigned char s_dyn[];
	float4* const s_f4 = reinterpret_cast<float4*>(s_dyn);
	xq_entry* const s_xq = reinterpret_cast<xq_entry*>(s_dyn + 6 * PLANE * 16);
	const unsigned short* const s_key = reinterpret_cast<const unsigned short*>(s_dyn);
	unsigned* const s_perm = reinterpret_cast<unsigned*>(s_dyn + smem_front);
	float* const s_raw = reinterpret_cast<float*>(s_dyn + smem_raw);
	const int JW3 = 3 * g.nrow;
	__shared__ int s_cnt[NC + 1], s_cur[NC + 1];          // bin NC: the holes
	__shared__ int s_nmig, s_done;
	__shared__ __align__(8) unsigned long long s_bar;

	const int t = tile_list[blockIdx.x];
	const int tx = t % g.ntx, ty = t / g.ntx;
	const int x0 = tx * TX, y0 = ty * TY;
	const int cx = min(TX, g.nx - x0), cy = min(TY, g.ny - y0);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int n = tile_np[t];
	const int64_t base = tile_off[t];
	// an empty tile (the other half of the box of a half-box species, vacuum ahead of a plasma edge) has nothing to
	// stage, sort or push: it only reports that it stays empty
	if (n == 0) {
		if (threadIdx.x == 0) { mig.np[t] = 0; tile_np_out[t] = 0; }
		return;
	}

	// ---- the tile's keys: one bulk copy, in flight while the fields are staged
	if (threadIdx.x == 0) {
		s_nmig = 0; s_done = 0;
		mbar_init(&s_bar, 1);
		if (n > 0) bulk_load(s_dyn, A.key + base, (unsigned) ((n * 2 + 15) & ~15), &s_bar);
	}
	// ---- stage the field neighbourhood: cells [x0-1, x0+cx] x [y0-1, y0+cy] as raw planes ...
	for (int k = threadIdx.x; k < (cx + 2) * (cy + 2); k += PUSH_THREADS) {
		int r = k / (cx + 2), c = k - r * (cx + 2);
		int gi = (x0 + c) + (y0 + r) * g.nrow;        // buffer index of cell (x0-1+c, y0-1+r)
		f3 e = E[gi], b = B[gi];
		int o = c + r * SROW;
		s_raw[o] = e.x; s_raw[o + PLANE] = e.y; s_raw[o + 2 * PLANE] = e.z;
		s_raw[o + 3 * PLANE] = b.x; s_raw[o + 4 * PLANE] = b.y; s_raw[o + 5 * PLANE] = b.z;
	}
	for (int k = threadIdx.x; k <= NC; k += PUSH_THREADS) s_cnt[k] = 0;
	__syncthreads();
	if (n > 0) mbar_wait(&s_bar, 0);

	// ---- phase A: counting sort of slot indices by cell.  The keys arrive almost sorted (B was written in
	//      cell order by the previous step), so 32 CONSECUTIVE keys would hammer one counter (shared-memory
	//      atomics on one address serialise).  Instead every thread walks its own stretch of consecutive
	//      keys: at any moment the lanes of a warp are a stretch apart, i.e. in different cells, and the
	//      atomics spread over the counters.  Stretches are an odd number of 32-bit words long, so the
	//      lanes also read distinct banks.  Plain uniform loops: no votes, nothing divergent.
	// Lane l owns the words [l*S, (l+1)*S) of the key array (S odd: the lanes of a warp read distinct banks
	// and sit S*2 keys apart, i.e. in different cells as long as a cell holds fewer particles than that);
	// the warps split every lane's stretch into WARPS consecutive pieces.
	// Branch-free: the keys past n (up to the 64 * S the lanes cover) are padded with KEY_EMPTY, and a hole counts
	// as cell NC - it gets a counter and a stretch of perm[] behind the live entries like any other cell.
	const int S = ((((n + 1) >> 1) + 31) >> 5) | 1;
	{
		unsigned short* const kw = const_cast<unsigned short*>(s_key);
		for (int k = n + threadIdx.x; k < 64 * S; k += PUSH_THREADS) kw[k] = (unsigned short) KEY_EMPTY;
		__syncthreads();
	}
	const int ws = (S + PUSH_WARPS - 1) / PUSH_WARPS;
	const int w0 = lane * S + warp * ws, wn = min(ws, S - warp * ws);
	const unsigned* const s_key2 = reinterpret_cast<const unsigned*>(s_key) + w0;
	#pragma unroll 4
	for (int j = 0; j < wn; j++) {
		const unsigned two = s_key2[j];
		atomicAdd(&s_cnt[min(two & 0xffffu, (unsigned) NC)], 1);
		atomicAdd(&s_cnt[min(two >> 16, (unsigned) NC)], 1);
	}
	__syncthreads();
	int nlive;
	{	// exclusive scan of s_cnt (NC <= 256 == PUSH_THREADS); the holes start at nlive
		__shared__ int s_wsum[PUSH_WARPS];
		int v = (threadIdx.x < NC) ? s_cnt[threadIdx.x] : 0;
		int incl = v;
		for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += u; }
		if (lane == 31) s_wsum[warp] = incl;
		__syncthreads();
		int woff = 0, tot = 0;
		#pragma unroll
		for (int w = 0; w < PUSH_WARPS; w++) { int c = s_wsum[w]; woff += (w < warp) ? c : 0; tot += c; }
		if (threadIdx.x < NC) s_cur[threadIdx.x] = woff + incl - v;
		if (threadIdx.x == 0) s_cur[NC] = tot;
		nlive = tot;
		__syncthreads();
	}
	#pragma unroll 4
	for (int j = 0; j < wn; j++) {
		const unsigned two = s_key2[j];
		const unsigned c0 = min(two & 0xffffu, (unsigned) NC), c1 = min(two >> 16, (unsigned) NC);
		const unsigned i = 2u * (unsigned) (w0 + j);
		s_perm[atomicAdd(&s_cur[c0], 1)] = (c0 << 16) | i;
		s_perm[atomicAdd(&s_cur[c1], 1)] = (c1 << 16) | (i + 1u);
	}
	__syncthreads();                                    // the keys are dead: their bytes become corner tile + queues
	// ---- the fields as the four corners of every cell (entries of the last row / column are never read)
	for (int k = threadIdx.x; k < 6 * PLANE; k += PUSH_THREADS) {
		const int pl = k / PLANE, o = k - pl * PLANE;
		const int r = o / SROW, c = o - r * SROW;
		const int o1 = (c + 1 < SROW) ? o + 1 : o, o2 = (r + 1 < TY + 2) ? o + SROW : o;
		const int o3 = o2 + (o1 - o);
		const float* P = s_raw + pl * PLANE;
		s_f4[k] = make_float4(P[o], P[o2], P[o1], P[o3]);
	}
	__syncthreads();

	// ---- phase B: every warp streams a contiguous range of the sorted particles, 64 per iteration (lane l
	//      owns the particles l and l+32 of the iteration's block); no block barriers from here on
	xq_entry* const xq = s_xq + warp * XQ_CAP;
	const unsigned lt = (1u << lane) - 1u;
	int nxq = 0;
	float energy = 0.0f;      // per-thread partial in float (a few dozen terms), widened once per tile
	jtile jt;
	jt.v = reinterpret_cast<float*>(J + x0 + y0 * g.nrow);       // tile cell (-1,-1) = buffer cell (x0, y0)
	const float* const Arec = A.rec + (size_t) (base >> 5) * REC_CHUNK_WORDS;
	float* const Brec = Bo.rec + (size_t) (base >> 5) * REC_CHUNK_WORDS;
	const int chunk = ((nlive + PUSH_WARPS * 64 - 1) / (PUSH_WARPS * 64)) * 64;
	const int pbeg = warp * chunk, pend = min(pbeg + chunk, nlive);
	const int mig_cap = (int) ((tile_off[t + 1] - base) / mig.div);
	const int64_t mig_base = base / mig.div;

#ifndef PUSH_CARRY_PPC
#define PUSH_CARRY_PPC 24
#endif
	const int carry_max = (nlive >= PUSH_CARRY_PPC * NC) ? 1 : -1;      // see deposit32 (the scan's last run)
	// per-lane partial sums of the 8 contributions to cell `cur` (warp-uniform; -1: none)
	float acc[8];
	#pragma unroll
	for (int q = 0; q < 8; q++) acc[q] = 0.0f;
	int cur = -1;

	// Current of 32 consecutive sorted particles (one per lane, zero for the lanes that deposit through the
	// queue): accumulate per lane while the warp stays in one cell, reduce when it moves on.
	auto deposit32 = [&](int key, float (&w)[8], bool act, int lx, int ly) {
		const int prev = __shfl_up_sync(0xffffffffu, key, 1);
		const unsigned heads = __ballot_sync(0xffffffffu, (lane == 0) ? (key != cur) : (key != prev));
		if (heads == 0u) {
			#pragma unroll
			for (int q = 0; q < 8; q++) acc[q] += w[q];
			return;
		}
		const int b = __ffs(heads) - 1;                // lanes below b continue cell `cur`
		const bool lo = lane < b;
		// masks as multipliers: the selects would all land on the (half-rate) ALU pipe
		const float mlo = lo ? 1.0f : 0.0f, mhi = lo ? 0.0f : 1.0f;
		if (cur >= 0) {
			#pragma unroll
			for (int q = 0; q < 8; q++) acc[q] = __fmaf_rn(w[q], mlo, acc[q]);
			flush_cell<TX>(acc, cur, lane, jt, JW3);
		}
		if ((heads & (heads - 1u)) == 0u) {
			// one new cell starts at lane b and runs to the end of the warp: it becomes `cur`.  (Two or three new cells
			// through a loop of such butterflies instead of the scan below: no faster at 32 particles per cell, 3 - 8 %
			// slower at 16, and the loop form cost the one-cell case 11 instructions per 32 particles -
			// profiles/r02_push_lowppc_ab.txt.)
			#pragma unroll
			for (int q = 0; q < 8; q++) acc[q] = w[q] * mhi;
			cur = __shfl_sync(0xffffffffu, key, 31);
			if (cur >= NC) cur = -1;                   // the range ended inside these 32
		} else {
			// many cells start here (a few particles per cell): segmented inclusive scan, the last lane of each run
			// holds its totals
			#pragma unroll
			for (int q = 0; q < 8; q++) w[q] *= mhi;
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
			// Every finished run is added to J by its last lane (8 reductions from one lane: ~10 instructions, against
			// ~50 for the butterfly).  The last run is not finished.  In a tile with >= PUSH_CARRY_PPC particles per cell,
			// when at most one cell starts after lane 0, its totals stay in acc and its cell becomes `cur`: the next 32
			// lanes then continue it and are likely to see one boundary only (the cheaper path above).  With more
			// cells per 32 lanes the next group takes this path whatever happens, and a carried run would only add a
			// butterfly to it - the run is flushed here.  Measured, G push/s at 64 / 32 / 16 / 8 per cell and ms per
			// LWFA step (16 per cell): never carry 56.4 / 46.1 / 39.9 / 32.2 / 1.78, carry in every tile 56.8 / 48.3 /
			// 38.9 / 31.7 / 1.94 (profiles/r02_push_lowppc_ab.txt).
			const bool carry = __popc(heads & ~1u) <= carry_max;
			const bool tail = (lane == 31 ? !carry : (bool) ((heads >> (lane + 1)) & 1u));
			if (tail && act && !lo) jt_weights(jt, (lx + 1) * 3 + (ly + 1) * JW3, JW3, w);
			#pragma unroll
			for (int q = 0; q < 8; q++) acc[q] = (carry && lane == 31) ? w[q] : 0.0f;
			cur = carry ? __shfl_sync(0xffffffffu, key, 31) : -1;
			if (cur >= NC) cur = -1;
		}
	};

	// source slots of the particles (pa, pa+32); lanes past the end of the range read the range's first particle
	auto load_pair = [&](int pa, pair_rec& r) {
		const unsigned va = s_perm[pa < pend ? pa : pbeg], vb = s_perm[pa + 32 < pend ? pa + 32 : pbeg];
		const int ia = va & 0xffffu, ib = vb & 0xffffu;
		const float* qa = Arec + (ia >> 5) * REC_CHUNK_WORDS + (ia & 31);
		const float* qb = Arec + (ib >> 5) * REC_CHUNK_WORDS + (ib & 31);
		r.x = mk2(qa[0], qb[0]); r.y = mk2(qa[32], qb[32]);
		r.ux = mk2(qa[64], qb[64]); r.uy = mk2(qa[96], qb[96]); r.uz = mk2(qa[128], qb[128]);
		r.ca = (int) (va >> 16); r.cb = (int) (vb >> 16);
		r.ta = r.tb = 0;
		if (TAGS) { r.ta = A.tag[base + ia]; r.tb = A.tag[base + ib]; }
	};

#ifndef PUSH_PREFETCH
#define PUSH_PREFETCH 1              // software pipeline: request the next iteration's records before the math
#endif
	pair_rec nv;
	if (PUSH_PREFETCH && pbeg < pend) load_pair(pbeg + lane, nv);

	// one iteration = 64 particles; FULL: all 64 exist (every iteration but the last of a warp's range), so the
	// activity masks are compile-time true and the predicates, selects and store guards they feed disappear
	auto advance64 = [&](const int p0, auto full_tag) {
		constexpr bool FULL = decltype(full_tag)::value;
		const int pa = p0 + lane, pb = pa + 32;
		const bool actA = FULL || pa < pend, actB = FULL || pb < pend;
		if (!PUSH_PREFETCH) load_pair(pa, nv);
		const pair_rec v = nv;
		if (PUSH_PREFETCH && p0 + 64 < pend) load_pair(pa + 64, nv);

		constexpr int LOG_TX = (TX == 16) ? 4 : ((TX == 8) ? 3 : 2);
		const int lxa = v.ca & (TX - 1), lya = v.ca >> LOG_TX, lxb = v.cb & (TX - 1), lyb = v.cb >> LOG_TX;
		f2 x = v.x, y = v.y, ux = v.ux, uy = v.uy, uz = v.uz;
		f2 dx, dy, qvz;
		{
			f2 Ex, Ey, Ez, Bx, By, Bz;
			interp_EB_f4<SROW, PLANE>(s_f4, lxa, lya, x.x, y.x, Ex.x, Ey.x, Ez.x, Bx.x, By.x, Bz.x);
			interp_EB_f4<SROW, PLANE>(s_f4, lxb, lyb, x.y, y.y, Ex.y, Ey.y, Ez.y, Bx.y, By.y, Bz.y);
			const f2 en = boris2(Ex, Ey, Ez, Bx, By, Bz, prm.tem, ux, uy, uz);
			energy += (actA ? en.x : 0.0f) + (actB ? en.y : 0.0f);
		}
		{
			const f2 usq = add2(add2(add2(bc2(1.0f), mul2(ux, ux)), mul2(uy, uy)), mul2(uz, uz));
			const f2 rg = div_exact2(bc2(1.0f), sqrt_exact2(usq));
			dx = mul2(mul2(rg, prm.dt_dx), ux);
			dy = mul2(mul2(rg, prm.dt_dy), uy);
			qvz = mul2(mul2(uz, prm.q), rg);
		}
		const f2 x1 = add2(x, dx), y1 = add2(y, dy);
		const int dia = ltrim(x1.x), dib = ltrim(x1.y), dja = ltrim(y1.x), djb = ltrim(y1.y);

		// Every move deposits its FIRST in-cell piece - up to the first cell face it crosses, the whole move
		// for the 91 % that cross none - through the lanes (its cell is the cell of its neighbours in the
		// sorted order).  Only the remainder of a crossing move goes to the queue: 1 piece (2 for a corner
		// cut) in the neighbouring cell.  The pieces join at (xe, ye), so charge is conserved exactly as in
		// the reference's split (particles.c:785-879), which cuts the same straight line at the same faces.
		const bool xa = actA && ((dia | dja) != 0);
		const bool xb = actB && ((dib | djb) != 0);
		const f2 fx = mk2(dia > 0 ? 1.0f : 0.0f, dib > 0 ? 1.0f : 0.0f), fy = mk2(dja > 0 ? 1.0f : 0.0f, djb > 0 ? 1.0f : 0.0f);
		f2 tx = __fmul2_rn(sub2(fx, x), rcp_approx2(dx)), ty = __fmul2_rn(sub2(fy, y), rcp_approx2(dy));
		tx.x = dia ? tx.x : 2.0f; tx.y = dib ? tx.y : 2.0f;
		ty.x = dja ? ty.x : 2.0f; ty.y = djb ? ty.y : 2.0f;
		const f2 t1 = mk2(fmaxf(fminf(fminf(tx.x, ty.x), 1.0f), 0.0f), fmaxf(fminf(fminf(tx.y, ty.y), 1.0f), 0.0f));
		const bool xfa = dia != 0 && tx.x <= ty.x, xfb = dib != 0 && tx.y <= ty.y;       // the x face comes first
		const bool yfa = dja != 0 && !xfa, yfb = djb != 0 && !xfb;
		const f2 dx0 = __fmul2_rn(dx, t1), dy0 = __fmul2_rn(dy, t1);
		f2 xe = add2(x, dx0), ye = add2(y, dy0);
		xe.x = xfa ? fx.x : xe.x; xe.y = xfb ? fx.y : xe.y;
		ye.x = yfa ? fy.x : ye.x; ye.y = yfb ? fy.y : ye.y;
		{
			f2 w2[8];
			const f2 kz = mk2(actA ? 0.5f : 0.0f, actB ? 0.5f : 0.0f);
			seg_weights2(x, y, xe, ye, dx0, dy0, __fmul2_rn(qvz, t1), mul2(kz, prm.qnx), mul2(kz, prm.qny), kz, w2);
			float w[8];
			#pragma unroll
			for (int q = 0; q < 8; q++) w[q] = w2[q].x;
#ifndef ABL_NO_DEPOSIT
			deposit32(actA ? v.ca : 0x7fffffff, w, actA, lxa, lya);
			#pragma unroll
			for (int q = 0; q < 8; q++) w[q] = w2[q].y;
			deposit32(actB ? v.cb : 0x7fffffff, w, actB, lxb, lyb);
#else
			acc[0] += w[0] + w[1] + w[2] + w[3] + w[4] + w[5] + w[6] + w[7];
			#pragma unroll
			for (int q = 0; q < 8; q++) w[q] = w2[q].y;
			acc[1] += w[0] + w[1] + w[2] + w[3] + w[4] + w[5] + w[6] + w[7];
#endif
		}

		// --- queue of the remainders; drain 32 at a time
		{
#ifdef ABL_NO_QUEUE
			const unsigned ma = 0, mb = 0;
#else
			const unsigned ma = __ballot_sync(0xffffffffu, xa), mb = __ballot_sync(0xffffffffu, xb);
#endif
			if (ma | mb) {
				if (xa) {
					xq_entry e;
					e.cij = v.ca + ((dia + 1) << 16) + ((dja + 1) << 18);
					e.x = x.x; e.y = y.x; e.dx = dx.x; e.dy = dy.x; e.qvz = qvz.x;
					xq[nxq + __popc(ma & lt)] = e;
				}
				nxq += __popc(ma);
				if (xb) {
					xq_entry e;
					e.cij = v.cb + ((dib + 1) << 16) + ((djb + 1) << 18);
					e.x = x.y; e.y = y.y; e.dx = dx.y; e.dy = dy.y; e.qvz = qvz.y;
					xq[nxq + __popc(mb & lt)] = e;
				}
				nxq += __popc(mb);
				__syncwarp();
				while (nxq >= 32) {
					drain_queue<TX>(xq + nxq - 32, 32, lane, jt, JW3, prm.qnx, prm.qny);
					nxq -= 32;
				}
				__syncwarp();
			}
		}

		// --- new positions; survivors go to their sorted slot in B, leavers to the tile's migrants segment
		const f2 xn = sub2(x1, mk2((float) dia, (float) dib)), yn = sub2(y1, mk2((float) dja, (float) djb));
		const int nlxa = lxa + dia - prm.shift_window, nlya = lya + dja;
		const int nlxb = lxb + dib - prm.shift_window, nlyb = lyb + djb;
		const bool sta = actA && (unsigned) nlxa < (unsigned) cx && (unsigned) nlya < (unsigned) cy;
		const bool stb = actB && (unsigned) nlxb < (unsigned) cx && (unsigned) nlyb < (unsigned) cy;
		// p0 is a multiple of 64: pa sits in chunk p0/32 at lane, pb in the next chunk at lane
		if (actA) {
			float* qd = Brec + (pa >> 5) * REC_CHUNK_WORDS + (pa & 31);
			qd[0] = xn.x; qd[32] = yn.x; qd[64] = ux.x; qd[96] = uy.x; qd[128] = uz.x;
			Bo.key[base + pa] = sta ? (unsigned short) (nlxa + nlya * TX) : (unsigned short) KEY_EMPTY;
			if (TAGS) Bo.tag[base + pa] = v.ta;
		}
		if (actB) {
			float* qd = Brec + (pb >> 5) * REC_CHUNK_WORDS + (pb & 31);
			qd[0] = xn.y; qd[32] = yn.y; qd[64] = ux.y; qd[96] = uy.y; qd[128] = uz.y;
			Bo.key[base + pb] = stb ? (unsigned short) (nlxb + nlyb * TX) : (unsigned short) KEY_EMPTY;
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
						part_aos r; r.ix = x0 + nlxa; r.iy = y0 + nlya;
						r.x = xn.x; r.y = yn.x; r.ux = ux.x; r.uy = uy.x; r.uz = uz.x;
						mig.rec[mig_base + d] = r;
						if (TAGS) mig.tag[mig_base + d] = v.ta;
					}
				}
				if (lb) {
					const int d = slot + __popc(ma) + __popc(mb & lt);
					if (d < mig_cap) {
						part_aos r; r.ix = x0 + nlxb; r.iy = y0 + nlyb;
						r.x = xn.y; r.y = yn.y; r.ux = ux.y; r.uy = uy.y; r.uz = uz.y;
						mig.rec[mig_base + d] = r;
						if (TAGS) mig.tag[mig_base + d] = v.tb;
					}
				}
			}
		}
	};
	{
		int p0 = pbeg;
		// (unrolling this loop by two to drop the pipeline's register moves doubles the 21 KB body past the 32 KB
		//  instruction cache: measured -24 %)
		for (; p0 + 64 <= pend; p0 += 64) advance64(p0, std::true_type());
		if (p0 < pend) advance64(p0, std::false_type());
	}
	if (cur >= 0) flush_cell<TX>(acc, cur, lane, jt, JW3);
	if (nxq) drain_queue<TX>(xq, nxq, lane, jt, JW3, prm.qnx, prm.qny);

	// ---- tile epilogue (no block barrier: warps retire independently): energy, slots in use, migrants
	double e = (double) energy;
	for (int o = 16; o > 0; o >>= 1) e += __shfl_down_sync(0xffffffffu, e, o);
	if (lane == 0 && nlive > 0) atomicAdd(&ctl->energy, e);
	if (lane == 0) {
		__threadfence_block();
		if (atomicAdd(&s_done, 1) == PUSH_WARPS - 1) {
			// last warp out: every reservation in s_nmig has been made
			const int nm = atomicAdd(&s_nmig, 0);
			if (nm > mig_cap) atomicOr(&ctl->flags, 2u);
			mig.np[t] = min(nm, mig_cap);
			tile_np_out[t] = nlive;
		}
	}
}

// Boundary conditions for the particles that left their tile (reference particles.c:1237-1259: periodic y
// always; x periodic, absorbing under a moving window, or handed to the neighbour slab), then append them
// to their destination tiles.  One warp per tile segment.
__global__ void __launch_bounds__(256, 3) k_migrate2d(soa2d p, const int64_t* __restrict__ tile_off, int* __restrict__ tile_np, mig2d mig,
                            ctl2d* __restrict__ ctl, int TX, int TY, int ntx, int ntiles, int nx, int ny,
                            int moving_window, int slab_left, int slab_right,
                            part_aos* __restrict__ exp_l, part_aos* __restrict__ exp_r, unsigned int exp_cap,
                            part_aos* __restrict__ ovf, int* __restrict__ ovf_tag, unsigned int ovf_cap, slab_pub pub) {
	// The work per record is a chain of dependent memory operations (record -> slot reservation -> tile bounds ->
	// stores), so a warp keeps MIG_U records per lane in flight and runs the chain phase by phase; reservations of
	// lanes that target the same tile (most leavers of a tile go to the same neighbour) or the same export list are
	// combined into one atomic per warp instruction.
	constexpr int MIG_U = 4;
	const int lane = threadIdx.x & 31;
	const unsigned lt = (1u << lane) - 1u;
	const int nwarp = (gridDim.x * blockDim.x) >> 5;
	const int lgx = 31 - __clz(TX), lgy = 31 - __clz(TY);             // tile shapes are powers of two
	for (int ts = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ts < ntiles; ts += nwarp) {
		const int n = mig.np[ts];
		const int64_t mb = tile_off[ts] / mig.div;
		for (int k0 = 0; k0 < n; k0 += 32 * MIG_U) {                  // (warp-uniform trip count: the lanes vote below)
			part_aos r[MIG_U];
			int t[MIG_U], ex[MIG_U], tag[MIG_U];
			#pragma unroll
			for (int u = 0; u < MIG_U; u++) {
				const int k = k0 + 32 * u + lane;
				t[u] = -1; ex[u] = -1; tag[u] = 0;
				if (k < n) { r[u] = mig.rec[mb + k]; if (p.tag) tag[u] = mig.tag[mb + k]; t[u] = 0; }
			}
			// boundary conditions: periodic y always; x periodic, absorbing under a moving window, or export
			#pragma unroll
			for (int u = 0; u < MIG_U; u++) {
				if (t[u] < 0) continue;
				int ix = r[u].ix, iy = r[u].iy;
				iy += ((iy < 0) ? ny : 0) - ((iy >= ny) ? ny : 0);
				if (ix < 0 || ix >= nx) {
					const int side = ix >= nx;
					if (side ? slab_right : slab_left) { ex[u] = side; ix += side ? -nx : nx; t[u] = -1; }   // the neighbour's frame
					else if (moving_window) t[u] = -1;                                                  // absorbed
					else ix += side ? -nx : nx;
				}
				r[u].ix = ix; r[u].iy = iy;
				if (t[u] >= 0) t[u] = (ix >> lgx) + (iy >> lgy) * ntx;
			}
			// leaves the slab: export (all slabs have the same width); one reservation per side and instruction
			if (slab_left | slab_right) {
				#pragma unroll
				for (int u = 0; u < MIG_U; u++) {
					#pragma unroll
					for (int side = 0; side < 2; side++) {
						const unsigned m = __ballot_sync(0xffffffffu, ex[u] == side);
						if (m == 0u) continue;
						const int leader = __ffs(m) - 1;
						unsigned int base = 0;
						if (lane == leader) base = atomicAdd(&ctl->n_exp[side], (unsigned) __popc(m));
						base = __shfl_sync(0xffffffffu, base, leader);
						if (ex[u] == side) {
							const unsigned int slot = base + __popc(m & lt);
							if (slot >= exp_cap) atomicOr(&ctl->flags, 4u);
							else (side ? exp_r : exp_l)[slot] = r[u];
						}
					}
				}
			}
			// reserve a slot in the destination tile
			int slot[MIG_U];
			#pragma unroll
			for (int u = 0; u < MIG_U; u++) {
				const unsigned grp = __match_any_sync(0xffffffffu, t[u]);
				const int leader = __ffs(grp) - 1;
				int base = 0;
				if (lane == leader && t[u] >= 0) base = atomicAdd(&tile_np[t[u]], __popc(grp));
				slot[u] = __shfl_sync(0xffffffffu, base, leader) + __popc(grp & lt);
			}
			int64_t lo[MIG_U], hi[MIG_U];
			#pragma unroll
			for (int u = 0; u < MIG_U; u++) if (t[u] >= 0) { lo[u] = tile_off[t[u]]; hi[u] = tile_off[t[u] + 1]; }
			#pragma unroll
			for (int u = 0; u < MIG_U; u++) {
				if (t[u] < 0) continue;
				const int64_t d = lo[u] + slot[u];
				if (d >= hi[u]) {
					atomicSub(&tile_np[t[u]], 1);
					ovf_push(ctl, ovf, ovf_tag, ovf_cap, r[u], tag[u]);
					continue;
				}
				const int lx = r[u].ix & (TX - 1), ly = r[u].iy & (TY - 1);
				rec_store(p.rec, d, r[u].x, r[u].y, r[u].ux, r[u].uy, r[u].uz);
				p.key[d] = (unsigned short) (lx + ly * TX);
				if (p.tag) p.tag[d] = tag[u];
			}
		}
	}
	// linked slabs: the export lists ARE the neighbours' mailboxes; tell them how many records arrived
	if (pub.flag[0]) slab_publish(pub.ticket, gridDim.x, pub.flag[0], pub.seq[0], pub.count[0], &ctl->n_exp[0]);
	if (pub.flag[1]) slab_publish(pub.ticket + 1, gridDim.x, pub.flag[1], pub.seq[1], pub.count[1], &ctl->n_exp[1]);
}

// What the neighbour slabs sent (blockIdx.y = side): wait for the message, then append its records to their tiles
__global__ void k_slab_import(soa2d p, const int64_t* __restrict__ tile_off, int* __restrict__ tile_np, ctl2d* __restrict__ ctl,
                              int TX, int TY, int ntx, slab_in in, unsigned int cap, part_aos* __restrict__ ovf,
                              int* __restrict__ ovf_tag, unsigned int ovf_cap) {
	const int side = blockIdx.y;
	if (!in.flag[side]) return;
	slab_wait(in.flag[side], in.seq[side]);
	const unsigned n = min(__ldcg(in.count[side]), cap);       // (a sender whose list overflowed aborts at its next step)
	const int* src = reinterpret_cast<const int*>(in.rec[side]);
	for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
		const int* w = src + (size_t) k * 7;
		part_aos r;
		r.ix = __ldcg(w); r.iy = __ldcg(w + 1);
		r.x = __int_as_float(__ldcg(w + 2)); r.y = __int_as_float(__ldcg(w + 3));
		r.ux = __int_as_float(__ldcg(w + 4)); r.uy = __int_as_float(__ldcg(w + 5)); r.uz = __int_as_float(__ldcg(w + 6));
		const int tx = r.ix / TX, ty = r.iy / TY, t = tx + ty * ntx;
		const int slot = atomicAdd(&tile_np[t], 1);
		const int64_t d = tile_off[t] + slot;
		if (d >= tile_off[t + 1]) { atomicSub(&tile_np[t], 1); ovf_push(ctl, ovf, ovf_tag, ovf_cap, r, 0); continue; }
		const int lx = r.ix - tx * TX, ly = r.iy - ty * TY;
		rec_store(p.rec, d, r.x, r.y, r.ux, r.uy, r.uz);
		p.key[d] = (unsigned short) (lx + ly * TX);
		if (p.tag) p.tag[d] = 0;
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
	size_t smem = push_smem_layout(TX, TY, s->max_cap).total;
	static size_t configured = 0;
	if (smem > configured) {
		ZDEV_CHECK(cudaFuncSetAttribute(k_push2d<TX, TY, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		ZDEV_CHECK(cudaFuncSetAttribute(k_push2d<TX, TY, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
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
	// (the few outgrown tiles of a density spike - one long-running CTA each - started first on a side stream so that the
	//  ordinary tiles fill the machine around them: measured on the LWFA probe, 1.874 against 1.860 ms/step; not kept)
	for (int grp = 0; grp < 2; grp++) {
		const int ntl = grp ? s->ntiles - s->n_small : s->n_small;
		if (ntl <= 0) continue;
		const int cap = grp ? s->max_cap : s->cap_small;
		const int* list = s->tile_list + (grp ? s->n_small : 0);
		const push_smem L = push_smem_layout(TX, TY, cap);
		const size_t sm = L.total;
		const unsigned front = (unsigned) L.front, permb = (unsigned) L.raw;
		if (s->track_ids)
			ZDEV_LAUNCH((k_push2d<TX, TY, true>), ntl, PUSH_THREADS, sm, s->p, s->q, s->tile_off, s->tile_np, s->tile_np_q,
			            s->mig, s->ctl, E, B, J, g, prm, front, permb, list);
		else
			ZDEV_LAUNCH((k_push2d<TX, TY, false>), ntl, PUSH_THREADS, sm, s->p, s->q, s->tile_off, s->tile_np, s->tile_np_q,
			            s->mig, s->ctl, E, B, J, g, prm, front, permb, list);
	}
	if (slot >= 0) ZDEV_CHECK(cudaEventRecord((*s->ev)[2 * slot + 1], zdev_strm));
}

extern "C" void zdev_spec2d_advance(zdev_spec2d* s, zdev_grid2d* grid, zdev_grid2d* gcur, const zdev_push2d_params* prm) {
	if (zdev_grid2d_nx(grid) != s->nx || zdev_grid2d_ny(grid) != s->ny ||
	    zdev_grid2d_nx(gcur) != s->nx || zdev_grid2d_ny(gcur) != s->ny) {
		fprintf(stderr, "(*error*) zdev_spec2d_advance: species / grid size mismatch\n"); exit(-1);
	}
	spec_settle(s);                  // the last step (or an append since) may have hit a full tile: grow before pushing
	if (s->cap_total && s->appended) spec_resolve_overflow(s);
	s->last_valid = 0;
	const int slab_left = s->slab ? (s->link.left >= 0) : prm->slab_left;
	const int slab_right = s->slab ? (s->link.right >= 0) : prm->slab_right;
	zdev_push2d_params prm_local = *prm;
	prm_local.slab_left = slab_left; prm_local.slab_right = slab_right;
	prm = &prm_local;
	if (prm->moving_window || prm->slab_left || prm->slab_right) s->np_known = 0;
	ZDEV_CHECK(cudaMemsetAsync(s->ctl, 0, sizeof(ctl2d), zdev_strm));
	if (!s->cap_total) return;
	// a window shift sends a whole column of every tile through the migrants segments
	if (prm->moving_window && s->mig.div > 2) mig_alloc(s, 2);
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
	if ((prm->slab_left || prm->slab_right) && !s->exp_cap && !s->slab) {
		// a window shift sends a whole column at once: size the export lists for two columns
		int64_t cap = (int64_t) 2 * s->ppc_hint * s->ny + 65536;
		s->exp_cap = (unsigned int) cap;
		for (int k = 0; k < 2; k++) ZDEV_CHECK(cudaMalloc(&s->exp_buf[k], (size_t) cap * sizeof(part_aos)));
	}
	slab_pub pub; memset(&pub, 0, sizeof pub);
	slab_in in; memset(&in, 0, sizeof in);
	part_aos* exp_l = s->exp_buf[0]; part_aos* exp_r = s->exp_buf[1];
	if (s->slab) {
		zdev_link& K = s->link;
		pub.ticket = K.ticket;
		for (int side = 0; side < 2; side++) {
			if ((side ? K.right : K.left) < 0) continue;
			const unsigned seq = ++K.seq[side];
			zdev_mbox_hdr* out = zdev_link_out_hdr(K, side);
			pub.flag[side] = &out->flag[1 - side]; pub.count[side] = &out->count[1 - side][seq & 1u]; pub.seq[side] = seq;
			(side ? exp_r : exp_l) = (part_aos*) zdev_link_out(K, side, seq);
			zdev_mbox_hdr* me = zdev_link_in_hdr(K);
			in.flag[side] = &me->flag[side]; in.count[side] = &me->count[side][seq & 1u]; in.seq[side] = seq;
			in.rec[side] = (const part_aos*) zdev_link_in(K, side, seq);
		}
	}
	ZDEV_LAUNCH(k_migrate2d, 8 * zdev_num_sm, 256, 0, s->p, s->tile_off, s->tile_np, s->mig, s->ctl,
	            s->TX, s->TY, s->ntx, s->ntiles, s->nx, s->ny, prm->moving_window, prm->slab_left, prm->slab_right,
	            exp_l, exp_r, s->exp_cap, s->ovf, s->ovf_tag, s->ovf_cap, pub);
	if (s->slab) { s->pending_in = in; s->import_pending = 1; }      // the snapshot follows the import
	else spec_snapshot_ctl(s);
	if (prm->moving_window || prm->slab_left || prm->slab_right) s->ids_valid = 0;
}

static void spec_flush_import(zdev_spec2d* s) {
	if (!s->import_pending) return;
	s->import_pending = 0;
	ZDEV_LAUNCH(k_slab_import, dim3(2 * zdev_num_sm, 2), 256, 0, s->p, s->tile_off, s->tile_np, s->ctl, s->TX, s->TY, s->ntx,
	            s->pending_in, s->exp_cap, s->ovf, s->ovf_tag, s->ovf_cap);
	spec_snapshot_ctl(s);
}
extern "C" void zdev_spec2d_flush_import(zdev_spec2d* s) { spec_flush_import(s); }

extern "C" void zdev_spec2d_fetch(zdev_spec2d* s, double* energy_sum, int64_t* np) {
	// the overflow check at the end of the advance has already brought the step's control block to the host;
	// the particle count only changes through absorbing boundaries, slab exchange and appends
	spec_settle(s);
	if (s->last_valid && (!np || s->np_known)) {
		check_flags(s, s->last.flags);
		if (np) *np = s->np_host;
		if (energy_sum) *energy_sum = s->last.energy;
		return;
	}
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
		s->np_host = (int64_t) h.np; s->np_known = 1;
		*np = (int64_t) h.np;
	}
	if (energy_sum) *energy_sum = h.energy;
}

// ------------------------------------------------------------------ charge deposit

// reference spec_deposit_charge, em2d/particles.c:1289-1324 (node centred, linear).  The slots of a tile are
// almost in cell order: the four weights of neighbouring slots in the same cell are combined with a segmented
// warp scan and only the last lane of a run issues the four L2 reductions (one per particle and node is
// ~30x slower: 64 particles of a cell serialise on the same four addresses).
__global__ void k_deposit_charge(soa2d p, const int64_t* __restrict__ off, const int* __restrict__ tile_np,
                                 float* __restrict__ rho, int nrow, float q, int TX, int TY, int ntx) {
	const int t = blockIdx.x, lane = threadIdx.x & 31;
	const int n = tile_np[t];
	const int x0 = (t % ntx) * TX, y0 = (t / ntx) * TY;
	const int64_t b = off[t];
	for (int k0 = (threadIdx.x >> 5) * 32; k0 < n; k0 += blockDim.x) {
		const int k = k0 + lane;
		const unsigned key = (k < n) ? p.key[b + k] : KEY_EMPTY;
		float w[4] = {0.0f, 0.0f, 0.0f, 0.0f};
		int idx = 0;
		if (key != KEY_EMPTY) {
			const float* r = p.rec + rec_word(b + k);
			const float w1 = r[0], w2 = r[32];
			idx = (x0 + (int) (key % TX)) + nrow * (y0 + (int) (key / TX));
			w[0] = (1.0f - w1) * (1.0f - w2) * q; w[1] = (w1) * (1.0f - w2) * q;
			w[2] = (1.0f - w1) * (w2) * q;        w[3] = (w1) * (w2) * q;
		}
		const unsigned prev = __shfl_up_sync(0xffffffffu, key, 1);
		const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || key != prev);
		const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const bool take = (lane - d) >= start;
			#pragma unroll
			for (int c = 0; c < 4; c++) { float u = __shfl_up_sync(0xffffffffu, w[c], d); if (take) w[c] += u; }
		}
		const bool tail = (lane == 31) || ((heads >> (lane + 1)) & 1u);
		if (tail && key != KEY_EMPTY) {
			atomicAdd(&rho[idx], w[0]); atomicAdd(&rho[idx + 1], w[1]);
			atomicAdd(&rho[idx + nrow], w[2]); atomicAdd(&rho[idx + 1 + nrow], w[3]);
		}
	}
}
__global__ void k_charge_fold(float* __restrict__ rho, int nx, int ny, int mode) {
	int nrow = nx + 1;
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (mode == 0) { if (k <= ny) rho[(size_t) k * nrow] += rho[nx + (size_t) k * nrow]; }   // x fold
	else           { if (k <= nx) rho[k] += rho[k + (size_t) ny * nrow]; }                   // y fold
}

// scratch of the charge diagnostic, kept between calls: device grid + pinned host staging (a pageable 4 MB
// copy each way and a cudaMalloc/cudaFree pair per call cost 10x the kernel)
static float* g_rho_dev = nullptr;
static float* g_dep_dev = nullptr;
static float* g_rho_pin = nullptr;
static size_t g_rho_cap = 0;
static cudaEvent_t g_rho_ev[4] = {nullptr, nullptr, nullptr, nullptr};

// the caller's array is pageable memory: large ones are copied to / from the pinned scratch by a few threads (one
// thread moves ~10 GB/s; the 67 MB charge grid of a 4096^2 box cost 5 ms each way)
static void par_memcpy(void* dst, const void* src, size_t bytes) {
	const size_t min_chunk = (size_t) 4 << 20;
	unsigned nt = std::thread::hardware_concurrency();
	nt = std::max(1u, std::min(std::min(nt, 8u), (unsigned) (bytes / min_chunk)));
	{ static int knob = -1; if (knob < 0) { const char* e = getenv("ZPIC_PAR_MEMCPY"); knob = !e || e[0] != '0'; } if (!knob) nt = 1; }
	if (nt <= 1) { memcpy(dst, src, bytes); return; }
	std::vector<std::thread> th;
	const size_t chunk = ((bytes + nt - 1) / nt + 4095) & ~(size_t) 4095;     // nt chunks cover every byte (rounded UP)
	for (unsigned k = 0; k < nt; k++) {
		const size_t o = (size_t) k * chunk;
		if (o >= bytes) break;
		const size_t len = std::min(chunk, bytes - o);
		th.emplace_back([=] { memcpy((char*) dst + o, (const char*) src + o, len); });
	}
	for (auto& t : th) t.join();
}

extern "C" void zdev_spec2d_par_memcpy(void* dst, const void* src, size_t bytes) { par_memcpy(dst, src, bytes); }   // host only: tests/test_abi_symbols.py

__global__ void k_charge_add(float* __restrict__ rho, const float* __restrict__ dep, size_t n) {
	for (size_t k = (size_t) blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t) gridDim.x * blockDim.x) rho[k] += dep[k];
}

// The deposit ADDS to the caller's array and then folds the periodic edges of the SUM (the reference folds whatever the
// array held, particles.c:1310-1322), so the caller's values have to come up.  They come up BEHIND the kernel: the
// particles are deposited on a zeroed grid while the host copies the (pageable) array into the pinned buffer and the
// upload runs; the two are added on the device, folded, and come back in four pieces - the host copies one out while
// the next one arrives.  (At 4096^2 the host copies are 2/3 of the call: 13.1 ms per species before.)
extern "C" void zdev_spec2d_deposit_charge(zdev_spec2d* s, float q, int moving_window, float* charge) {
	spec_settle(s);
	size_t n = (size_t) (s->nx + 1) * (s->ny + 1);
	if (n > g_rho_cap) {
		ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
		cudaFree(g_rho_dev); cudaFree(g_dep_dev); cudaFreeHost(g_rho_pin);
		ZDEV_CHECK(cudaMalloc(&g_rho_dev, n * sizeof(float)));
		ZDEV_CHECK(cudaMalloc(&g_dep_dev, n * sizeof(float)));
		ZDEV_CHECK(cudaHostAlloc((void**) &g_rho_pin, n * sizeof(float), cudaHostAllocPortable));
		g_rho_cap = n;
		if (!g_rho_ev[0]) for (int k = 0; k < 4; k++) ZDEV_CHECK(cudaEventCreateWithFlags(&g_rho_ev[k], cudaEventDisableTiming));
	}
	float* d_rho = g_rho_dev;
	ZDEV_CHECK(cudaMemsetAsync(g_dep_dev, 0, n * sizeof(float), zdev_strm));
	if (s->cap_total)
		ZDEV_LAUNCH(k_deposit_charge, s->ntiles, 256, 0, s->p, s->tile_off, s->tile_np, g_dep_dev, s->nx + 1, q,
		            s->TX, s->TY, s->ntx);
	par_memcpy(g_rho_pin, charge, n * sizeof(float));                        // (the kernel is running)
	ZDEV_CHECK(cudaMemcpyAsync(d_rho, g_rho_pin, n * sizeof(float), cudaMemcpyHostToDevice, zdev_strm));
	ZDEV_LAUNCH(k_charge_add, 4 * zdev_num_sm, 256, 0, d_rho, g_dep_dev, n);
	if (!moving_window) ZDEV_LAUNCH(k_charge_fold, zdev_div_up(s->ny + 1, 128), 128, 0, d_rho, s->nx, s->ny, 0);
	ZDEV_LAUNCH(k_charge_fold, zdev_div_up(s->nx + 1, 128), 128, 0, d_rho, s->nx, s->ny, 1);
	const int pieces = n >= ((size_t) 1 << 20) ? 4 : 1;
	const size_t per = (n / pieces + 1023) & ~(size_t) 1023;
	for (int k = 0; k < pieces; k++) {
		const size_t o = (size_t) k * per, len = std::min(per, n - std::min(o, n));
		if (len) ZDEV_CHECK(cudaMemcpyAsync(g_rho_pin + o, d_rho + o, len * sizeof(float), cudaMemcpyDeviceToHost, zdev_strm));
		ZDEV_CHECK(cudaEventRecord(g_rho_ev[k], zdev_strm));
	}
	for (int k = 0; k < pieces; k++) {
		const size_t o = (size_t) k * per, len = std::min(per, n - std::min(o, n));
		ZDEV_CHECK(cudaEventSynchronize(g_rho_ev[k]));
		if (len) par_memcpy(charge + o, g_rho_pin + o, len * sizeof(float));
	}
}

// the slab's own deposit alone: rho = (nx+1)*(ny+1) floats, overwritten, NOT folded (the caller joins the slabs,
// adds the shared edge columns and applies the box's periodic folds)
extern "C" void zdev_spec2d_deposit_charge_raw(zdev_spec2d* s, float q, float* rho) {
	spec_settle(s);
	const size_t n = (size_t) (s->nx + 1) * (s->ny + 1);
	float* d_rho; ZDEV_CHECK(cudaMalloc(&d_rho, n * sizeof(float)));
	ZDEV_CHECK(cudaMemsetAsync(d_rho, 0, n * sizeof(float), zdev_strm));
	if (s->cap_total)
		ZDEV_LAUNCH(k_deposit_charge, s->ntiles, 256, 0, s->p, s->tile_off, s->tile_np, d_rho, s->nx + 1, q, s->TX, s->TY, s->ntx);
	ZDEV_CHECK(cudaMemcpyAsync(rho, d_rho, n * sizeof(float), cudaMemcpyDeviceToHost, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	cudaFree(d_rho);
}

// ------------------------------------------------------------------ phasespace density

// reference spec_deposit_pha, em2d/particles.c:1569-1632: linear deposit of the charge on a 2-D grid over two of
// {x1, x2, u1, u2, u3} (axis values as spec_pha_axis :1512-1538).  Grids of up to PHA_SMEM_BINS bins are
// accumulated per CTA in shared memory and merged with one L2 reduction per touched bin; larger ones go
// straight to L2.
#define PHA_SMEM_BINS 8192
struct pha_params { int q1, q2, n1, n2; float min1, min2, rd1, rd2, q, dx, dy; int gx0; };

__device__ __forceinline__ float pha_axis(int quant, float x, float y, float ux, float uy, float uz, int ix, int iy, float dx, float dy) {
	switch (quant) {                      // the reference's X1, X2, U1, U2, U3 (em2d/particles.h:236-240)
	case 1: return (x + ix) * dx;
	case 2: return (y + iy) * dy;
	case 4: return ux;
	case 5: return uy;
	default: return uz;
	}
}

__global__ void k_deposit_pha(soa2d p, const int64_t* __restrict__ off, const int* __restrict__ tile_np, int ntiles,
                              float* __restrict__ buf, pha_params a, int TX, int TY, int ntx, int use_smem) {
	extern __shared__ float s_pha[];
	const int nbins = a.n1 * a.n2;
	if (use_smem) {
		for (int k = threadIdx.x; k < nbins; k += blockDim.x) s_pha[k] = 0.0f;
		__syncthreads();
	}
	float* const dst = use_smem ? s_pha : buf;
	for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
		const int n = tile_np[t];
		const int x0 = (t % ntx) * TX, y0 = (t / ntx) * TY;
		const int64_t b = off[t];
		for (int k = threadIdx.x; k < n; k += blockDim.x) {
			const unsigned key = p.key[b + k];
			if (key == KEY_EMPTY) continue;
			const rec20 v = rec_load(p.rec, b + k);
			const int ix = a.gx0 + x0 + (int) (key % TX), iy = y0 + (int) (key / TX);
			const float nx1 = (pha_axis(a.q1, v.x, v.y, v.ux, v.uy, v.uz, ix, iy, a.dx, a.dy) - a.min1) * a.rd1;
			const float nx2 = (pha_axis(a.q2, v.x, v.y, v.ux, v.uy, v.uz, ix, iy, a.dx, a.dy) - a.min2) * a.rd2;
			const int i1 = (int) (nx1 + 0.5f), i2 = (int) (nx2 + 0.5f);
			const float w1 = nx1 - i1 + 0.5f, w2 = nx2 - i2 + 0.5f;
			const int idx = i1 + a.n1 * i2;
			const bool in1a = (i1 >= 0 && i1 < a.n1), in1b = (i1 + 1 >= 0 && i1 + 1 < a.n1);
			if (i2 >= 0 && i2 < a.n2) {
				if (in1a) atomicAdd(&dst[idx], (1.0f - w1) * (1.0f - w2) * a.q);
				if (in1b) atomicAdd(&dst[idx + 1], w1 * (1.0f - w2) * a.q);
			}
			if (i2 + 1 >= 0 && i2 + 1 < a.n2) {
				if (in1a) atomicAdd(&dst[idx + a.n1], (1.0f - w1) * w2 * a.q);
				if (in1b) atomicAdd(&dst[idx + a.n1 + 1], w1 * w2 * a.q);
			}
		}
	}
	if (use_smem) {
		__syncthreads();
		for (int k = threadIdx.x; k < nbins; k += blockDim.x) { const float v = s_pha[k]; if (v != 0.0f) atomicAdd(&buf[k], v); }
	}
}

extern "C" void zdev_spec2d_deposit_pha(zdev_spec2d* s, int quant1, int quant2, const int pha_nx[2], const float pha_range[2][2],
                                        float q, float dx, float dy, float* host_buf) {
	spec_settle(s);
	const size_t n = (size_t) pha_nx[0] * pha_nx[1];
	float* d_buf; ZDEV_CHECK(cudaMalloc(&d_buf, n * sizeof(float)));
	ZDEV_CHECK(cudaMemcpyAsync(d_buf, host_buf, n * sizeof(float), cudaMemcpyHostToDevice, zdev_strm));
	if (s->cap_total) {
		pha_params a;
		a.q1 = quant1; a.q2 = quant2; a.n1 = pha_nx[0]; a.n2 = pha_nx[1];
		a.min1 = pha_range[0][0]; a.min2 = pha_range[1][0];
		a.rd1 = pha_nx[0] / (pha_range[0][1] - pha_range[0][0]);       // float arithmetic as the reference (:1583-1584)
		a.rd2 = pha_nx[1] / (pha_range[1][1] - pha_range[1][0]);
		a.q = q; a.dx = dx; a.dy = dy; a.gx0 = s->gx0;
		const int use_smem = n <= PHA_SMEM_BINS;
		const int grid = s->ntiles < 4 * zdev_num_sm ? s->ntiles : 4 * zdev_num_sm;
		ZDEV_LAUNCH(k_deposit_pha, grid, 256, use_smem ? n * sizeof(float) : 0, s->p, s->tile_off, s->tile_np, s->ntiles,
		            d_buf, a, s->TX, s->TY, s->ntx, use_smem);
	}
	ZDEV_CHECK(cudaMemcpyAsync(host_buf, d_buf, n * sizeof(float), cudaMemcpyDeviceToHost, zdev_strm));
	ZDEV_CHECK(cudaStreamSynchronize(zdev_strm));
	cudaFree(d_buf);
}

// ------------------------------------------------------------------ slab decomposition support

extern "C" void zdev_spec2d_export_counts(zdev_spec2d* s, int64_t counts[2]) {
	spec_settle(s);
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
	spec_settle(s);
	spec_append_dev(s, (const part_aos*) dev_aos, np, 0);
	s->np_host += np;
	s->ids_valid = 0;
	spec_snapshot_ctl(s);
}

// This species is one slab of a wider box: open the links to the neighbour slabs (collective over the ranks).
// gx0 / gnx: the slab's first column in the whole box and the box width (the device-side injector numbers its
// particles by GLOBAL cell, so the decomposed population is the single-domain one).
extern "C" void zdev_spec2d_set_slab(zdev_spec2d* s, int left, int right, int gx0, int gnx) {
	if (s->slab) return;
	// a window shift sends a whole column at once: room for two columns of the nominal fill, and then some
	const int64_t cap = (int64_t) 2 * s->ppc_hint * s->ny + 65536;
	s->exp_cap = (unsigned int) cap;
	zdev_link_open(&s->link, (size_t) cap * sizeof(part_aos), left, right);
	s->slab = 1; s->gx0 = gx0; s->gnx = gnx;
}
