/* zpic-b200 :: host-side substrate of a slab-decomposed run (see zb_par.h) */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <signal.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>
#include "zb_par.h"

#define ZB_PAR_MAGIC 0x5a504943u           /* "ZPIC" */
#define ZB_PAR_MAX_RANKS 64
#define ZB_PAR_SLOT 1024                   /* bytes per rank in the small exchange area */
#define ZB_PAR_RED 1024                    /* doubles per rank in the reduction area */

typedef struct {
	_Atomic uint32_t magic;
	int nranks;
	int creator_pid;
	_Atomic int attached;
	_Atomic int bar_count;
	_Atomic int bar_sense;
	_Atomic int scratch_gen;               /* generation of the big scratch segment */
	size_t scratch_bytes;
	unsigned char slot[ZB_PAR_MAX_RANKS][ZB_PAR_SLOT];
	double red[ZB_PAR_MAX_RANKS][ZB_PAR_RED];
} zb_par_hdr;

static int par_rank = 0, par_nranks = 1, par_ready = 0;
static zb_par_hdr* hdr = NULL;
static char job_name[96];
static int local_sense = 0;
static void* scratch_ptr = NULL;
static size_t scratch_len = 0;
static int scratch_gen_seen = 0;

static int env_int( const char* a, const char* b, int dflt ) {
	const char* e = getenv(a);
	if (!e && b) e = getenv(b);
	return e ? atoi(e) : dflt;
}

static void nap( int* spins ) {
	if (++*spins < 2000) return;
	if (*spins < 20000) { sched_yield(); return; }
	struct timespec ts = { 0, 50000 };
	nanosleep(&ts, NULL);
}

static void die( const char* what ) {
	fprintf(stderr, "(*error*) zpic-b200 rank %d: %s (%s)\n", par_rank, what, strerror(errno));
	exit(-1);
}

int zb_par_rank( void ) { return par_rank; }
int zb_par_nranks( void ) { return par_nranks; }

static void at_exit_unlink( void ) { zb_par_finalize(); }

int zb_par_init( void )
{
	if (par_ready) return par_nranks;
	par_ready = 1;
	const char* off = getenv("ZPIC_SLABS");
	int n = env_int("ZPIC_NRANKS", "WORLD_SIZE", 1);
	if ((off && atoi(off) == 0) || n <= 1) { par_nranks = 1; par_rank = 0; return 1; }
	if (n > ZB_PAR_MAX_RANKS) { fprintf(stderr, "(*error*) zpic-b200: at most %d ranks\n", ZB_PAR_MAX_RANKS); exit(-1); }
	par_nranks = n;
	par_rank = env_int("ZPIC_RANK", "RANK", 0);
	const char* job = getenv("ZPIC_JOB");
	if (!job) job = getenv("MASTER_PORT");
	if (!job) job = "default";
	snprintf(job_name, sizeof job_name, "/zpic_b200_%s_%d", job, (int) getuid());

	const size_t bytes = sizeof(zb_par_hdr);
	if (par_rank == 0) {
		shm_unlink(job_name);                                  /* a stale segment of a dead job */
		int fd = shm_open(job_name, O_CREAT | O_EXCL | O_RDWR, 0600);
		if (fd < 0) die("cannot create the job's shared-memory segment");
		if (ftruncate(fd, (off_t) bytes) != 0) die("ftruncate");
		hdr = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
		close(fd);
		if (hdr == MAP_FAILED) die("mmap");
		memset(hdr, 0, bytes);
		hdr->nranks = n;
		hdr->creator_pid = (int) getpid();
		atomic_store(&hdr->magic, ZB_PAR_MAGIC);
		int spins = 0;
		while (atomic_load(&hdr->attached) < n - 1) nap(&spins);
		atexit(at_exit_unlink);
	} else {
		/* wait for a segment whose creator is alive (a leftover of a crashed job is replaced by rank 0) */
		int spins = 0;
		for (;;) {
			int fd = shm_open(job_name, O_RDWR, 0600);
			if (fd >= 0) {
				struct stat st;
				if (fstat(fd, &st) == 0 && (size_t) st.st_size >= bytes) {
					zb_par_hdr* h = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
					close(fd);
					if (h != MAP_FAILED) {
						int ok = 0;
						for (int k = 0; k < 200 && !ok; k++) {
							if (atomic_load(&h->magic) == ZB_PAR_MAGIC && h->nranks == n &&
							    h->creator_pid > 0 && kill(h->creator_pid, 0) == 0) ok = 1;
							else { struct timespec ts = { 0, 1000000 }; nanosleep(&ts, NULL); }
						}
						if (ok) { hdr = h; break; }
						munmap(h, bytes);
					}
				} else close(fd);
			}
			nap(&spins);
			if (spins > 4000000) { errno = ETIMEDOUT; die("rank 0 never created the job segment"); }
		}
		atomic_fetch_add(&hdr->attached, 1);
	}
	zb_par_barrier();
	return par_nranks;
}

void zb_par_barrier( void )
{
	if (par_nranks <= 1) return;
	local_sense = !local_sense;
	if (atomic_fetch_add(&hdr->bar_count, 1) == par_nranks - 1) {
		atomic_store(&hdr->bar_count, 0);
		atomic_store(&hdr->bar_sense, local_sense);
	} else {
		int spins = 0;
		while (atomic_load(&hdr->bar_sense) != local_sense) nap(&spins);
	}
}

void zb_par_allreduce_sum_d( double* v, int n )
{
	if (par_nranks <= 1) return;
	if (n > ZB_PAR_RED) { fprintf(stderr, "(*error*) zb_par_allreduce_sum_d: too many values\n"); exit(-1); }
	memcpy(hdr->red[par_rank], v, (size_t) n * sizeof(double));
	zb_par_barrier();
	for (int k = 0; k < n; k++) {
		double s = 0;
		for (int r = 0; r < par_nranks; r++) s += hdr->red[r][k];     /* same order on every rank */
		v[k] = s;
	}
	zb_par_barrier();
}

void zb_par_allreduce_sum_ll( long long* v, int n )
{
	if (par_nranks <= 1) return;
	if (n > ZB_PAR_RED) { fprintf(stderr, "(*error*) zb_par_allreduce_sum_ll: too many values\n"); exit(-1); }
	memcpy(hdr->red[par_rank], v, (size_t) n * sizeof(long long));
	zb_par_barrier();
	for (int k = 0; k < n; k++) {
		long long s = 0;
		for (int r = 0; r < par_nranks; r++) s += ((const long long*) hdr->red[r])[k];
		v[k] = s;
	}
	zb_par_barrier();
}

void zb_par_allgather( const void* mine, size_t bytes, void* all )
{
	if (par_nranks <= 1) { memcpy(all, mine, bytes); return; }
	if (bytes > ZB_PAR_SLOT) { fprintf(stderr, "(*error*) zb_par_allgather: contribution too large\n"); exit(-1); }
	memcpy(hdr->slot[par_rank], mine, bytes);
	zb_par_barrier();
	for (int r = 0; r < par_nranks; r++) memcpy((char*) all + (size_t) r * bytes, hdr->slot[r], bytes);
	zb_par_barrier();
}

void* zb_par_scratch( size_t bytes )
{
	static void* solo = NULL; static size_t solo_len = 0;
	if (par_nranks <= 1) {
		if (bytes > solo_len) { free(solo); solo = malloc(bytes); solo_len = bytes; }
		return solo;
	}
	char name[128];
	zb_par_barrier();                                      /* nobody is still using the old area */
	if (par_rank == 0 && bytes > hdr->scratch_bytes) {
		if (hdr->scratch_bytes) { snprintf(name, sizeof name, "%s_s%d", job_name, atomic_load(&hdr->scratch_gen)); shm_unlink(name); }
		const int gen = atomic_load(&hdr->scratch_gen) + 1;
		snprintf(name, sizeof name, "%s_s%d", job_name, gen);
		shm_unlink(name);
		int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
		if (fd < 0) die("cannot create the shared scratch segment");
		if (ftruncate(fd, (off_t) bytes) != 0) die("shared scratch: ftruncate (is /dev/shm large enough?)");
		close(fd);
		hdr->scratch_bytes = bytes;
		atomic_store(&hdr->scratch_gen, gen);
	}
	zb_par_barrier();
	const int gen = atomic_load(&hdr->scratch_gen);
	if (gen != scratch_gen_seen) {
		if (scratch_ptr) munmap(scratch_ptr, scratch_len);
		snprintf(name, sizeof name, "%s_s%d", job_name, gen);
		int fd = shm_open(name, O_RDWR, 0600);
		if (fd < 0) die("cannot open the shared scratch segment");
		scratch_len = hdr->scratch_bytes;
		scratch_ptr = mmap(NULL, scratch_len, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
		close(fd);
		if (scratch_ptr == MAP_FAILED) die("shared scratch: mmap");
		scratch_gen_seen = gen;
	}
	return scratch_ptr;
}

void zb_par_allreduce_sum_f( float* v, size_t n )
{
	if (par_nranks <= 1 || n == 0) return;
	float* s = zb_par_scratch((size_t) par_nranks * n * sizeof(float));
	memcpy(s + (size_t) par_rank * n, v, n * sizeof(float));
	zb_par_barrier();
	for (size_t k = 0; k < n; k++) {
		float acc = 0;
		for (int r = 0; r < par_nranks; r++) acc += s[(size_t) r * n + k];
		v[k] = acc;
	}
	zb_par_barrier();
}

void zb_par_finalize( void )
{
	if (par_nranks <= 1 || !hdr) return;
	if (par_rank == 0) {
		char name[128];
		if (hdr->scratch_bytes) { snprintf(name, sizeof name, "%s_s%d", job_name, atomic_load(&hdr->scratch_gen)); shm_unlink(name); }
		shm_unlink(job_name);
	}
}
