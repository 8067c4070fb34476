/* zpic-b200 :: guarded host mirrors - see zb_guard.h */
#define _GNU_SOURCE
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <unistd.h>
#include <sys/mman.h>
#include "zb_guard.h"

#define ZB_G_MAX 1024

typedef struct {
	char* base; size_t len;          /* whole pages */
	int state;
	void* owner; int kind;
	zb_guard_fn fill, dirty;
} region;

static region regions[ZB_G_MAX];
static int n_regions = 0;
static int enabled = -1;
static int handler_on = 0;
static struct sigaction old_segv, old_bus;
static unsigned long n_fills = 0, n_dirties = 0;
static volatile int in_fill = 0;

int zb_guard_enabled( void )
{
	if (enabled < 0) { const char* e = getenv("ZPIC_GUARD"); enabled = e ? (atoi(e) != 0) : 1; }
	return enabled;
}

unsigned long zb_guard_fills( void ) { return n_fills; }
unsigned long zb_guard_dirties( void ) { return n_dirties; }

static region* find( const void* p )
{
	const char* a = p;
	for (int i = 0; i < n_regions; i++)
		if (a >= regions[i].base && a < regions[i].base + regions[i].len) return &regions[i];
	return NULL;
}

static const int prot_of[3] = { PROT_READ | PROT_WRITE, PROT_READ, PROT_NONE };

static void apply( region* r, int state )
{
	if (r->state == state) return;
	if (mprotect(r->base, r->len, prot_of[state]) != 0) {
		perror("(*error*) zpic-b200: mprotect of a host mirror");
		exit(-1);
	}
	r->state = state;
}

static void chain( int sig, siginfo_t* si, void* uc, const struct sigaction* old )
{
	if (old->sa_flags & SA_SIGINFO) {
		if (old->sa_sigaction) { old->sa_sigaction(sig, si, uc); return; }
	} else if (old->sa_handler != SIG_DFL && old->sa_handler != SIG_IGN) {
		old->sa_handler(sig);
		return;
	}
	/* default action: restore it and let the access fault again */
	signal(sig, SIG_DFL);
}

static void on_fault( int sig, siginfo_t* si, void* uc )
{
	region* r = find(si->si_addr);
	if (r && r->state == ZB_G_NONE && r->fill && !in_fill) {
		/* the device is newer: bring the mirror over, then retry the access */
		in_fill = 1;
		n_fills++;
		r->fill(r->owner, r->kind);
		in_fill = 0;
		r = find(si->si_addr);                     /* (the table may have been compacted) */
		if (r && r->state != ZB_G_NONE) return;
	} else if (r && r->state == ZB_G_READ) {
		/* first host write to a mirror that was in sync */
		apply(r, ZB_G_RW);
		n_dirties++;
		if (r->dirty) r->dirty(r->owner, r->kind);
		return;
	}
	chain(sig, si, uc, sig == SIGBUS ? &old_bus : &old_segv);
}

static void install( void )
{
	if (handler_on) return;
	struct sigaction sa;
	memset(&sa, 0, sizeof sa);
	sa.sa_sigaction = on_fault;
	sa.sa_flags = SA_SIGINFO | SA_NODEFER;
	sigemptyset(&sa.sa_mask);
	sigaction(SIGSEGV, &sa, &old_segv);
	sigaction(SIGBUS, &sa, &old_bus);
	handler_on = 1;
}

static size_t round_pages( size_t bytes )
{
	const size_t pg = (size_t) sysconf(_SC_PAGESIZE);
	if (bytes == 0) bytes = 1;
	return (bytes + pg - 1) / pg * pg;
}

void* zb_guard_alloc( size_t bytes )
{
	if (!zb_guard_enabled()) return calloc(bytes ? bytes : 1, 1);
	if (n_regions == ZB_G_MAX) { fprintf(stderr, "(*error*) zpic-b200: too many host mirrors\n"); exit(-1); }
	size_t len = round_pages(bytes);
	void* p;
	const size_t huge = (size_t) 2 << 20;
	if (len >= 2 * huge) {
		/* a large mirror (the E, B, J grids of a big box, a particle buffer): 2 MB-aligned and advised as huge pages -
		   changing the protection of 200 MB walks 50 000 page-table entries per call with 4 KB pages, 100 with 2 MB */
		len = (len + huge - 1) / huge * huge;
		char* raw = mmap(NULL, len + huge, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
		if (raw == MAP_FAILED) { fprintf(stderr, "(*error*) zpic-b200: host mirror of %zu bytes: out of memory\n", bytes); exit(-1); }
		char* al = (char*) (((uintptr_t) raw + huge - 1) / huge * huge);
		if (al > raw) munmap(raw, (size_t) (al - raw));
		if (al + len < raw + len + huge) munmap(al + len, (size_t) (raw + len + huge - (al + len)));
		p = al;
#ifdef MADV_HUGEPAGE
		madvise(p, len, MADV_HUGEPAGE);
#endif
	} else {
		p = mmap(NULL, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
		if (p == MAP_FAILED) { fprintf(stderr, "(*error*) zpic-b200: host mirror of %zu bytes: out of memory\n", bytes); exit(-1); }
	}
	install();
	region* r = &regions[n_regions++];
	memset(r, 0, sizeof *r);
	r->base = p; r->len = len; r->state = ZB_G_RW;
	return p;
}

void zb_guard_free( void* p )
{
	if (!p) return;
	region* r = find(p);
	if (!r) { free(p); return; }
	munmap(r->base, r->len);
	*r = regions[--n_regions];
}

void* zb_guard_realloc( void* p, size_t bytes, size_t keep )
{
	region* r = p ? find(p) : NULL;
	if (p && !r) return realloc(p, bytes);
	if (!zb_guard_enabled()) return realloc(p, bytes);
	if (r && round_pages(bytes) <= r->len) return p;
	void* q = zb_guard_alloc(bytes);
	if (r) {
		r = find(p);
		region* n = find(q);
		n->owner = r->owner; n->kind = r->kind; n->fill = r->fill; n->dirty = r->dirty;
		if (keep) {
			if (r->state == ZB_G_NONE) apply(r, ZB_G_READ);   /* (callers keep nothing of a stale mirror) */
			memcpy(q, p, keep);
		}
		zb_guard_free(p);
	}
	return q;
}

void zb_guard_bind( void* p, void* owner, int kind, zb_guard_fn fill, zb_guard_fn dirty )
{
	region* r = p ? find(p) : NULL;
	if (!r) return;
	r->owner = owner; r->kind = kind; r->fill = fill; r->dirty = dirty;
}

void zb_guard_set( void* p, int state )
{
	region* r = p ? find(p) : NULL;
	if (r) apply(r, state);
}

int zb_guard_state( const void* p )
{
	region* r = p ? find(p) : NULL;
	return r ? r->state : -1;
}
