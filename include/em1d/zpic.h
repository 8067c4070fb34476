/* zpic-b200 :: em1d public API (host side).
 *
 * These headers restate the C interface of the reference em1d code so that input
 * decks, the Cython module and any other caller link against libzpic_b200_em2d
 * instead of the reference objects ("link-time symbol replacement").
 * Struct layouts are field-for-field those of the reference (em1d/zpic.h:18-22,
 * current.h:41-70, emf.h:83-120, particles.h:29-132, simulation.h:13-29): device
 * state is kept in a side registry keyed by the host object address, never in
 * extra struct members, so objects compiled against either header set are
 * interchangeable.
 */
#ifndef ZPIC_B200_EM2D_ZPIC_H
#define ZPIC_B200_EM2D_ZPIC_H

/* AoS grid element, 12 bytes (reference em1d/zpic.h:18-22) */
typedef struct Float3 { float x, y, z; } float3;

#ifndef M_PI
#define M_PI   3.14159265358979323846264338327950288
#endif
#ifndef M_PI_2
#define M_PI_2 1.57079632679489661923132169163975144
#endif
#ifndef M_PI_4
#define M_PI_4 0.785398163397448309615660845819875721
#endif

#endif
