// zpic-b200 :: GPU-to-GPU links of a slab-decomposed run (one process per GPU, all on one NVLink / NVSwitch node).
//
// Every object that exchanges data with the two neighbour slabs (a grid: guard columns of E, B, J; a species:
// particles that crossed a slab edge) owns a MAILBOX in its own device memory:
//
//     [ header: flag[2], count[2][2] | payload[from-left][parity 0,1] | payload[from-right][parity 0,1] ]
//
// and maps the mailboxes of its neighbours through CUDA IPC (handles travel once, at set-up, through the
// host-side job segment, zb_par.h).  An exchange has no host involvement and no collective library call:
//   send   the packing kernel stores straight into the neighbour's payload over NVLink; its last block fences
//          (system scope) and stores the exchange number into the neighbour's flag;
//   recv   the unpacking kernel of the neighbour spins on its own flag until it carries the exchange number,
//          then reads the payload past L1.
// The messages over one edge are numbered alike on both of its ends (the run is SPMD, every exchange over an edge
// is two-way) and payloads alternate between two parity buffers: message n+2 is sent after message n+1 of the
// neighbour has been received, which the neighbour sent after consuming message n - so a payload is never
// overwritten before it has been read and no acknowledgement is needed.  Every send of a rank is enqueued
// before the matching receive, so two neighbours cannot wait for each other.
#pragma once
#include "zdev_common.cuh"
#include "../host/common/zb_par.h"

struct zdev_mbox_hdr {
	unsigned flag[2];            // [side the data comes from]: number of the last exchange that has fully arrived
	unsigned count[2][2];        // [side][parity]: records in the payload (particle links)
	unsigned pad[10];
};

struct zdev_link {
	int left, right;             // neighbour ranks, -1: none
	char* mine;                  // my mailbox
	char* peer[2];               // the mailboxes of my left / right neighbour (IPC mappings)
	size_t payload;              // bytes of one payload buffer
	unsigned seq[2];             // messages exchanged so far with the left / right neighbour (both ends of an
	                             // edge count alike; an exchange that skips an edge does not count there)
	unsigned* ticket;            // 2 device counters: last-block detection of the sending kernels
};

__host__ __device__ __forceinline__ size_t zdev_mbox_payload_off(size_t payload, int side, int parity) {
	return sizeof(zdev_mbox_hdr) + (size_t) (2 * side + parity) * payload;
}
static inline size_t zdev_mbox_bytes(size_t payload) { return sizeof(zdev_mbox_hdr) + 4 * payload; }

// my payload holding what arrived from `side` (0: left neighbour, 1: right) in exchange `seq`
static inline char* zdev_link_in(const zdev_link& L, int side, unsigned seq) {
	return L.mine + zdev_mbox_payload_off(L.payload, side, (int) (seq & 1u));
}
// where my message to the neighbour on `side` goes: the neighbour sees me on ITS other side
static inline char* zdev_link_out(const zdev_link& L, int side, unsigned seq) {
	return L.peer[side] + zdev_mbox_payload_off(L.payload, 1 - side, (int) (seq & 1u));
}
static inline zdev_mbox_hdr* zdev_link_out_hdr(const zdev_link& L, int side) { return (zdev_mbox_hdr*) L.peer[side]; }
static inline zdev_mbox_hdr* zdev_link_in_hdr(const zdev_link& L) { return (zdev_mbox_hdr*) L.mine; }

static inline void zdev_link_open(zdev_link* L, size_t payload, int left, int right) {
	memset(L, 0, sizeof *L);
	L->left = left; L->right = right;
	L->payload = (payload + 255) & ~(size_t) 255;
	const size_t bytes = zdev_mbox_bytes(L->payload);
	ZDEV_CHECK(cudaMalloc(&L->mine, bytes));
	ZDEV_CHECK(cudaMemset(L->mine, 0, bytes));
	ZDEV_CHECK(cudaMalloc(&L->ticket, 2 * sizeof(unsigned)));
	ZDEV_CHECK(cudaMemset(L->ticket, 0, 2 * sizeof(unsigned)));
	ZDEV_CHECK(cudaDeviceSynchronize());
	cudaIpcMemHandle_t h;
	ZDEV_CHECK(cudaIpcGetMemHandle(&h, L->mine));
	const int n = zb_par_nranks();
	cudaIpcMemHandle_t* all = (cudaIpcMemHandle_t*) malloc((size_t) n * sizeof h);
	zb_par_allgather(&h, sizeof h, all);
	if (left >= 0) ZDEV_CHECK(cudaIpcOpenMemHandle((void**) &L->peer[0], all[left], cudaIpcMemLazyEnablePeerAccess));
	if (right >= 0) {
		if (right == left) L->peer[1] = L->peer[0];
		else ZDEV_CHECK(cudaIpcOpenMemHandle((void**) &L->peer[1], all[right], cudaIpcMemLazyEnablePeerAccess));
	}
	free(all);
	zb_par_barrier();
}

static inline void zdev_link_close(zdev_link* L) {
	if (!L->mine) return;
	zb_par_barrier();                        // nobody is still writing into my mailbox
	if (L->peer[0]) cudaIpcCloseMemHandle(L->peer[0]);
	if (L->peer[1] && L->peer[1] != L->peer[0]) cudaIpcCloseMemHandle(L->peer[1]);
	zb_par_barrier();                        // ... and nobody still has it mapped
	cudaFree(L->mine); cudaFree(L->ticket);
	memset(L, 0, sizeof *L);
}

// --- device side of the protocol

__device__ __forceinline__ unsigned ld_flag(const unsigned* p) {
	unsigned v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_flag(unsigned* p, unsigned v) {
	asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// Called by every thread of a sending kernel after its stores into the neighbour's payload: the last block to
// arrive publishes exchange number `seq` in the neighbour's flag (the fence / ticket pattern of a grid-wide
// reduction, at system scope because the reader is another GPU).  `extra`, if given, is a second word stored
// before the flag (the record count of a particle link).
__device__ __forceinline__ void slab_publish(unsigned* ticket, unsigned nblocks, unsigned* peer_flag, unsigned seq,
                                             unsigned* peer_extra = nullptr, const unsigned* extra_src = nullptr) {
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0) {
		const unsigned t = atomicAdd(ticket, 1u);
		if (t == nblocks - 1u) {
			*ticket = 0u;
			__threadfence_system();
			if (peer_extra) { *(volatile unsigned*) peer_extra = *(volatile const unsigned*) extra_src; __threadfence_system(); }
			st_flag(peer_flag, seq);
		}
	}
}

// the same for ONE lane of a warp that is about to read the payload (the caller follows with __syncwarp())
__device__ __forceinline__ void slab_wait_lane(const unsigned* my_flag, unsigned seq);

// Called by every thread of a receiving kernel before it reads the payload
__device__ __forceinline__ void slab_wait_lane(const unsigned* my_flag, unsigned seq) {
	while ((int) (ld_flag(my_flag) - seq) < 0) __nanosleep(200);
}
__device__ __forceinline__ void slab_wait(const unsigned* my_flag, unsigned seq) {
	if (threadIdx.x == 0) {
		while ((int) (ld_flag(my_flag) - seq) < 0) __nanosleep(200);
	}
	__syncthreads();
}
