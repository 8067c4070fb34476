/* zpic-b200 :: guarded host mirrors (internal).
 *
 * The reference keeps all state in host buffers that callers read and write between iterations
 * (sim->emf.E_buf, species[i].part, ... - the Cython module hands out numpy views of them,
 * python/source/em2d.pyx:305-312, 1044-1296).  Here those buffers are MIRRORS of device state.  So that
 * unmodified callers still see reference semantics, every mirror is a page-aligned mapping whose protection
 * says which side is newer:
 *     ZB_G_NONE   the device is newer: any host access faults, the fault handler downloads the mirror
 *                 (fill callback) and retries the access;
 *     ZB_G_READ   both sides agree: reads are free, the first write faults, marks the mirror dirty
 *                 (dirty callback: it is uploaded before the next device step) and proceeds;
 *     ZB_G_RW     the host copy is the truth (or the library itself is working on it).
 * Off with ZPIC_GUARD=0 (plain allocations; callers then bracket raw accesses with zpic_b200_sync_* /
 * zpic_b200_touch_*, include/zpic_b200.h).  Not covered: passing a stale mirror straight to a system call
 * (write(2) returns EFAULT instead of faulting) - copy it first.
 */
#ifndef ZB_GUARD_H
#define ZB_GUARD_H
#include <stddef.h>

enum { ZB_G_RW = 0, ZB_G_READ = 1, ZB_G_NONE = 2 };

typedef void (*zb_guard_fn)( void* owner, int kind );

int   zb_guard_enabled( void );
/* zeroed, page-aligned; plain calloc when guards are off */
void* zb_guard_alloc( size_t bytes );
void  zb_guard_free( void* p );
/* a mapping of at least `bytes` holding the first `keep` bytes of p (p is released; p may be NULL) */
void* zb_guard_realloc( void* p, size_t bytes, size_t keep );
/* who to ask when the mirror is touched: fill() must leave it readable, dirty() records a host write */
void  zb_guard_bind( void* p, void* owner, int kind, zb_guard_fn fill, zb_guard_fn dirty );
/* change the protection (no-op for unguarded pointers and when the state is already `state`) */
void  zb_guard_set( void* p, int state );
int   zb_guard_state( const void* p );      /* -1: not a guarded mapping */
/* counters for tests / diagnostics */
unsigned long zb_guard_fills( void );
unsigned long zb_guard_dirties( void );

#endif
