/* zpic-b200 :: ZDF writer (format: SURVEY.md App. C; reference em2d/zdf.c:78-90,
 * 769-1268, 1500-1624).  Little-endian hosts only, like the shipped reference build.
 *
 * File    = "ZDF1" + records
 * Record  = u32 id|version, string name, u64 payload length, payload
 * String  = u32 length + bytes zero-padded to a multiple of 4
 */
#include <stdlib.h>
#include <string.h>
#include <errno.h>
#include <sys/stat.h>
#include "zdf.h"

#if defined(__BYTE_ORDER__) && __BYTE_ORDER__ != __ORDER_LITTLE_ENDIAN__
#error zpic-b200 ZDF writer supports little-endian hosts only
#endif

enum {
	REC_INT32 = 0x00010000, REC_DOUBLE = 0x00020000, REC_STRING = 0x00030000,
	REC_DATASET = 0x00100002, REC_ITERATION = 0x00200001, REC_GRID_INFO = 0x00210001,
	REC_PART_INFO = 0x00220002
};

static size_t pad4( size_t n ) { return (n + 3) & ~(size_t) 3; }

static int put( t_zdf_file* f, const void* p, size_t n ) { return fwrite(p, 1, n, f->fp) == n; }
static int put_u32( t_zdf_file* f, uint32_t v ) { return put(f, &v, 4); }
static int put_i32( t_zdf_file* f, int32_t v )  { return put(f, &v, 4); }
static int put_u64( t_zdf_file* f, uint64_t v ) { return put(f, &v, 8); }
static int put_f64( t_zdf_file* f, double v )   { return put(f, &v, 8); }

static size_t str_size( const char* s ) { size_t n = s ? strlen(s) : 0; return 4 + pad4(n); }

static int put_str( t_zdf_file* f, const char* s )
{
	static const char zeros[4] = {0, 0, 0, 0};
	uint32_t n = s ? (uint32_t) strlen(s) : 0;
	if (!put_u32(f, n)) return 0;
	if (n && !put(f, s, n)) return 0;
	size_t extra = pad4(n) - n;
	return extra ? put(f, zeros, extra) : 1;
}

/* record header; returns its size or 0 */
static size_t put_header( t_zdf_file* f, uint32_t id, const char* name, uint64_t length )
{
	if (!put_u32(f, id) || !put_str(f, name) || !put_u64(f, length)) return 0;
	return 4 + str_size(name) + 8;
}

size_t zdf_sizeof( enum zdf_data_type t )
{
	switch (t) {
	case zdf_int8: case zdf_uint8: return 1;
	case zdf_int16: case zdf_uint16: return 2;
	case zdf_int32: case zdf_uint32: case zdf_float32: return 4;
	case zdf_int64: case zdf_uint64: case zdf_float64: return 8;
	default: return 0;
	}
}

/* mkdir -p */
static int make_path( const char* path )
{
	char tmp[1024];
	size_t n = strlen(path);
	if (n == 0 || n >= sizeof tmp) return -1;
	memcpy(tmp, path, n + 1);
	for (char* p = tmp + 1; ; p++) {
		if (*p == '/' || *p == 0) {
			char c = *p; *p = 0;
			if (mkdir(tmp, 0755) && errno != EEXIST) return errno;
			*p = c;
			if (!c) break;
		}
	}
	return 0;
}

int zdf_open_file( t_zdf_file* zdf, const char* filename, enum zdf_file_access_mode mode )
{
	zdf->mode = mode;
	zdf->ndatasets = 0;
	if (mode != ZDF_CREATE) {
		fprintf(stderr, "(*error*) zdf_open_file: this build only writes ZDF files.\n");
		return 0;
	}
	if (!(zdf->fp = fopen(filename, "w+b"))) {
		perror("(*error*) Unable to open ZDF file for writing");
		return 0;
	}
	if (!put(zdf, "ZDF1", 4)) {
		fprintf(stderr, "(*error*) Unable to write magic number to ZDF file.\n");
		zdf_close_file(zdf);
		return 0;
	}
	return 1;
}

int zdf_close_file( t_zdf_file* zdf )
{
	if (fclose(zdf->fp)) { perror("(*error*) Unable to close ZDF file"); return 0; }
	zdf->fp = NULL;
	return 1;
}

size_t zdf_add_string( t_zdf_file* zdf, const char* name, const char* str )
{
	size_t h = put_header(zdf, REC_STRING, name, str_size(str));
	if (!h || !put_str(zdf, str)) return 0;
	return h + str_size(str);
}

size_t zdf_add_int32( t_zdf_file* zdf, const char* name, const int32_t value )
{
	size_t h = put_header(zdf, REC_INT32, name, 4);
	if (!h || !put_i32(zdf, value)) return 0;
	return h + 4;
}

size_t zdf_add_double( t_zdf_file* zdf, const char* name, const double value )
{
	size_t h = put_header(zdf, REC_DOUBLE, name, 8);
	if (!h || !put_f64(zdf, value)) return 0;
	return h + 8;
}

size_t zdf_add_iteration( t_zdf_file* zdf, const t_zdf_iteration* it )
{
	size_t len = 4 + 8 + str_size(it->time_units);
	size_t h = put_header(zdf, REC_ITERATION, it->name, len);
	if (!h || !put_i32(zdf, it->n) || !put_f64(zdf, it->t) || !put_str(zdf, it->time_units)) return 0;
	return h + len;
}

static size_t grid_info_size( const t_zdf_grid_info* g )
{
	size_t n = 4 + 8 * (size_t) g->ndims + str_size(g->label) + str_size(g->units) + 4;
	if (g->axis)
		for (unsigned i = 0; i < g->ndims; i++)
			n += str_size(g->axis[i].name) + 4 + 16 + str_size(g->axis[i].label) + str_size(g->axis[i].units);
	return n;
}

size_t zdf_add_grid_info( t_zdf_file* zdf, const t_zdf_grid_info* g )
{
	size_t len = grid_info_size(g);
	size_t h = put_header(zdf, REC_GRID_INFO, g->name, len);
	if (!h || !put_u32(zdf, g->ndims)) return 0;
	for (unsigned i = 0; i < g->ndims; i++) if (!put_u64(zdf, g->count[i])) return 0;
	if (!put_str(zdf, g->label) || !put_str(zdf, g->units) || !put_i32(zdf, g->axis != NULL)) return 0;
	if (g->axis)
		for (unsigned i = 0; i < g->ndims; i++) {
			const t_zdf_grid_axis* a = &g->axis[i];
			if (!put_str(zdf, a->name) || !put_i32(zdf, a->type) || !put_f64(zdf, a->min) ||
			    !put_f64(zdf, a->max) || !put_str(zdf, a->label) || !put_str(zdf, a->units)) return 0;
		}
	return h + len;
}

size_t zdf_add_part_info( t_zdf_file* zdf, const t_zdf_part_info* p )
{
	size_t len = str_size(p->label) + 8 + 4;
	for (unsigned i = 0; i < p->nquants; i++)
		len += str_size(p->quants[i]) + str_size(p->qlabels[i]) + str_size(p->qunits[i]);
	size_t h = put_header(zdf, REC_PART_INFO, p->name, len);
	if (!h || !put_str(zdf, p->label) || !put_u64(zdf, p->np) || !put_u32(zdf, p->nquants)) return 0;
	for (unsigned i = 0; i < p->nquants; i++) if (!put_str(zdf, p->quants[i])) return 0;
	for (unsigned i = 0; i < p->nquants; i++) if (!put_str(zdf, p->qlabels[i])) return 0;
	for (unsigned i = 0; i < p->nquants; i++) if (!put_str(zdf, p->qunits[i])) return 0;
	return h + len;
}

size_t zdf_add_dataset( t_zdf_file* zdf, t_zdf_dataset* ds )
{
	size_t count = 1;
	for (unsigned i = 0; i < ds->ndims; i++) count *= ds->count[i];
	size_t bytes = count * zdf_sizeof(ds->data_type);
	size_t len = 4 + 4 + 4 + 8 * (size_t) ds->ndims + bytes;

	size_t h = put_header(zdf, REC_DATASET, ds->name, len);
	if (!h) return 0;
	ds->offset = (uint64_t) ftello(zdf->fp);
	ds->id = ++zdf->ndatasets;
	if (!put_u32(zdf, (uint32_t) ds->id) || !put_i32(zdf, ds->data_type) || !put_u32(zdf, ds->ndims)) return 0;
	for (unsigned i = 0; i < ds->ndims; i++) if (!put_u64(zdf, ds->count[i])) return 0;
	if (bytes && !put(zdf, ds->data, bytes)) return 0;
	if (pad4(bytes) != bytes) { const char z[4] = {0}; if (!put(zdf, z, pad4(bytes) - bytes)) return 0; }
	return h + len;
}

int zdf_open_grid_file( t_zdf_file *zdf, const t_zdf_grid_info *info,
                        const t_zdf_iteration *iteration, char const path[] )
{
	char filename[1200];
	make_path(path);
	snprintf(filename, sizeof filename, "%s/%s-%06u.zdf", path, info->name, (unsigned) iteration->n);
	if (!zdf_open_file(zdf, filename, ZDF_CREATE)) {
		fprintf(stderr, "(*error*) Unable to open ZDF file, aborting.\n");
		return -1;
	}
	if (!zdf_add_string(zdf, "TYPE", "grid")) return 0;
	if (!zdf_add_grid_info(zdf, info)) return 0;
	if (!zdf_add_iteration(zdf, iteration)) return 0;
	return 1;
}

int zdf_save_grid( const void* data, enum zdf_data_type data_type, const t_zdf_grid_info *info,
                   const t_zdf_iteration *iteration, char const path[] )
{
	t_zdf_file zdf;
	if (zdf_open_grid_file(&zdf, info, iteration, path) != 1) return 0;
	t_zdf_dataset ds = { .name = info->name, .data_type = data_type, .ndims = info->ndims, .data = (void*) data };
	for (unsigned i = 0; i < info->ndims; i++) ds.count[i] = info->count[i];
	if (!zdf_add_dataset(&zdf, &ds)) return 0;
	return zdf_close_file(&zdf);
}

int zdf_open_part_file( t_zdf_file *zdf, t_zdf_part_info *info,
                        const t_zdf_iteration *iteration, char const path[] )
{
	char filename[1200];
	make_path(path);
	snprintf(filename, sizeof filename, "%s/%s-%s-%06u.zdf", path, "particles", info->name, (unsigned) iteration->n);
	if (!zdf_open_file(zdf, filename, ZDF_CREATE)) {
		fprintf(stderr, "(*error*) Unable to open ZDF file, aborting.\n");
		return -1;
	}
	if (!zdf_add_string(zdf, "TYPE", "particles")) return 0;
	if (!zdf_add_part_info(zdf, info)) return 0;
	if (!zdf_add_iteration(zdf, iteration)) return 0;
	return 1;
}

int zdf_add_quant_part_file( t_zdf_file *zdf, const char *name, const float* data, const uint64_t np )
{
	t_zdf_dataset ds = { .name = (char*) name, .data_type = zdf_float32, .ndims = 1, .data = (void*) data };
	ds.count[0] = np;
	return (int) zdf_add_dataset(zdf, &ds);
}
