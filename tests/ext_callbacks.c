/* Test-only C callbacks for CUSTOM external fields (ctypes cannot return a struct from a Python callback).
   ext_wire_B restates the deck function of the reference's em2d/input/extfld.c:17-48 (field of a wire through
   (x0, y0), evaluated at the staggered position of every component); ext_ripple_E is an arbitrary smooth E. */
#include <math.h>

typedef struct { float x, y, z; } f3;

f3 ext_wire_B(int ix, float dx, int iy, float dy, void* data)
{
	const float x0 = 6.4f, y0 = 6.4f;
	f3 b;
	float x = ix * dx - x0, y = (iy + 0.5) * dy - y0;
	b.x = -y / (x * x + y * y);
	x = (ix + 0.5) * dx - x0; y = iy * dy - y0;
	b.y = x / (x * x + y * y);
	b.z = 0;
	(void) data;
	return b;
}

f3 ext_ripple_E(int ix, float dx, int iy, float dy, void* data)
{
	f3 e = { 0.0f, 1e-3f * sinf(0.3f * ix * dx), 2e-3f * cosf(0.2f * iy * dy) };
	(void) data;
	return e;
}
