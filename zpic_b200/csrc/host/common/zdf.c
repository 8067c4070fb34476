/* zpic-b200 :: ZDF writer (format: SURVEY.md App. C; reference em2d/zdf.c:78-90,
 * 735-1624).  Little-endian hosts only, like the shipped reference build.
 *
 * File    = "ZDF1" + records
 * Record  = u32 id|version, string name, u64 payload length, payload
 * String  = u32 length + bytes zero-padded to a multiple of 4
 * Vector  = raw elements; 8-bit vectors are zero-padded to a multiple of 4 bytes, wider ones are not
 *           (and the padding is not counted in the record length) - reference zdf.c:709-756
 * Chunked dataset = start record (dataset header) + "<id>-chunk" records + "<id>-end" record
 */
#include <stdlib.h>
#include <string.h>
#include <errno.h>
#include <inttypes.h>
#include <sys/stat.h>
#include "zdf.h"

#if defined(__BYTE_ORDER__) && __BYTE_ORDER__ != __ORDER_LITTLE_ENDIAN__
#error zpic-b200 ZDF writer supports little-endian hosts only
#endif

enum {
	REC_INT32 = 0x00010000, REC_DOUBLE = 0x00020000, REC_STRING = 0x00030000,
	REC_DATASET = 0x00100002, REC_CDSET_START = 0x00110000, REC_CDSET_CHUNK = 0x00120000,
	REC_CDSET_END = 0x00130000, REC_ITERATION = 0x00200001, REC_GRID_INFO = 0x00210001,
	REC_PART_INFO = 0x00220002, REC_TRACK_INFO = 0x00230001
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
	if (mode == ZDF_CREATE) {
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
	if (mode != ZDF_READ && mode != ZDF_UPDATE) {
		fprintf(stderr, "(*error*) zdf_open_file: unsupported mode.\n");
		return 0;
	}
	/* existing file: check the magic number; updates continue at the end of the file */
	if (!(zdf->fp = fopen(filename, mode == ZDF_READ ? "r" : "r+"))) {
		perror(mode == ZDF_READ ? "(*error*) Unable to open ZDF file for reading"
		                        : "(*error*) Unable to open ZDF file for reading / writing.\n");
		return 0;
	}
	char magic[4];
	if (fread(magic, 1, 4, zdf->fp) != 4) {
		fprintf(stderr, "(*error*) Unable to read magic number from ZDF file.\n");
		zdf_close_file(zdf);
		return 0;
	}
	if (memcmp(magic, "ZDF1", 4)) {
		fprintf(stderr, "(*error*) Invalid magic number, file is not a proper ZDF file.\n");
		zdf_close_file(zdf);
		return 0;
	}
	if (mode == ZDF_UPDATE) fseeko(zdf->fp, 0, SEEK_END);
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

size_t zdf_add_track_info( t_zdf_file* zdf, const t_zdf_track_info* tr )
{
	size_t len = str_size(tr->label) + 4 * 4;
	for (unsigned i = 0; i < tr->nquants; i++)
		len += str_size(tr->quants[i]) + str_size(tr->qlabels[i]) + str_size(tr->qunits[i]);
	size_t h = put_header(zdf, REC_TRACK_INFO, tr->name, len);
	if (!h || !put_str(zdf, tr->label) || !put_u32(zdf, tr->ntracks) || !put_u32(zdf, tr->ndump) ||
	    !put_u32(zdf, tr->niter) || !put_u32(zdf, tr->nquants)) return 0;
	for (unsigned i = 0; i < tr->nquants; i++) if (!put_str(zdf, tr->quants[i])) return 0;
	for (unsigned i = 0; i < tr->nquants; i++) if (!put_str(zdf, tr->qlabels[i])) return 0;
	for (unsigned i = 0; i < tr->nquants; i++) if (!put_str(zdf, tr->qunits[i])) return 0;
	return h + len;
}

/* Raw elements.  The reference pads 8-bit vectors (only) to a multiple of 4 bytes and reports a vector of
 * zero elements like a failed write (zdf.c:709-756); callers that compare return values see the same here. */
size_t zdf_vector_write( t_zdf_file* zdf, const void* data, enum zdf_data_type data_type, size_t len )
{
	const size_t esize = zdf_sizeof(data_type);
	if (!esize) {
		fprintf(stderr, "(*error*) zdf_vector_write: Unsupported datatype.\n");
		return 0;
	}
	size_t bytes = len * esize;
	if (bytes && !put(zdf, data, bytes)) return 0;
	if (esize == 1 && pad4(bytes) != bytes) {
		const char z[4] = {0, 0, 0, 0};
		if (!put(zdf, z, pad4(bytes) - bytes)) return 0;
		bytes = pad4(bytes);
	}
	return bytes;
}

static size_t ds_elements( const uint64_t count[], unsigned ndims )
{
	size_t n = 1;
	for (unsigned i = 0; i < ndims; i++) n *= count[i];
	return n;
}

/* dataset header: u32 id, i32 type, u32 ndims, u64 count[ndims]; remembers where it sits so that
 * zdf_extend_dataset can rewrite it (reference zdf.c:1168-1199) */
static size_t ds_header_size( const t_zdf_dataset* ds ) { return 4 + 4 + 4 + 8 * (size_t) ds->ndims; }
static size_t put_ds_header( t_zdf_file* zdf, t_zdf_dataset* ds )
{
	off_t at = ftello(zdf->fp);
	if (at < 0) return 0;
	ds->offset = (uint64_t) at;
	if (!put_u32(zdf, (uint32_t) ds->id) || !put_i32(zdf, ds->data_type) || !put_u32(zdf, ds->ndims)) return 0;
	for (unsigned i = 0; i < ds->ndims; i++) if (!put_u64(zdf, ds->count[i])) return 0;
	return ds_header_size(ds);
}

size_t zdf_add_dataset( t_zdf_file* zdf, t_zdf_dataset* ds )
{
	const size_t count = ds_elements(ds->count, ds->ndims);
	const size_t len = ds_header_size(ds) + count * zdf_sizeof(ds->data_type);
	size_t h = put_header(zdf, REC_DATASET, ds->name, len);
	if (!h) return 0;
	ds->id = ++zdf->ndatasets;
	if (!put_ds_header(zdf, ds)) return 0;
	if (!zdf_vector_write(zdf, ds->data, ds->data_type, count)) return 0;
	return h + len;
}

/* ------------------------------------------------------------------ chunked datasets */

size_t zdf_start_cdset( t_zdf_file* zdf, t_zdf_dataset* ds )
{
	size_t h = put_header(zdf, REC_CDSET_START, ds->name, ds_header_size(ds));
	if (!h) return 0;
	ds->id = ++zdf->ndatasets;
	size_t d = put_ds_header(zdf, ds);
	return d ? h + d : 0;
}

/* what a chunk record occupies before its data: record id, the 16-character name "<id>-chunk" as a string,
 * record length, dataset id, count / start / stride */
size_t size_zdf_chunk_header( const t_zdf_dataset* ds )
{
	return 4 + (4 + 16) + 8 + 4 + 3 * 8 * (size_t) ds->ndims;
}

size_t zdf_write_chunk_header( t_zdf_file* zdf, t_zdf_dataset* ds, t_zdf_chunk* chunk )
{
	char name[16];
	snprintf(name, sizeof name, "%08" PRIx64 "-chunk", ds->id);
	const size_t meta = 4 + 3 * 8 * (size_t) ds->ndims;
	size_t h = put_header(zdf, REC_CDSET_CHUNK, name, ds_elements(chunk->count, ds->ndims) * zdf_sizeof(ds->data_type) + meta);
	if (!h || !put_u32(zdf, (uint32_t) ds->id)) return 0;
	for (unsigned i = 0; i < ds->ndims; i++) if (!put_u64(zdf, chunk->count[i])) return 0;
	for (unsigned i = 0; i < ds->ndims; i++) if (!put_u64(zdf, chunk->start[i])) return 0;
	for (unsigned i = 0; i < ds->ndims; i++) if (!put_u64(zdf, chunk->stride[i])) return 0;
	return h + meta;
}

size_t zdf_write_cdset( t_zdf_file* zdf, t_zdf_dataset* ds, t_zdf_chunk* chunk )
{
	size_t h = zdf_write_chunk_header(zdf, ds, chunk);
	if (!h) return 0;
	size_t v = zdf_vector_write(zdf, chunk->data, ds->data_type, ds_elements(chunk->count, ds->ndims));
	return v ? h + v : 0;
}

size_t zdf_end_cdset( t_zdf_file* zdf, t_zdf_dataset* ds )
{
	char name[16];
	snprintf(name, sizeof name, "%08" PRIx64 "-end", ds->id);
	return put_header(zdf, REC_CDSET_END, name, 0);
}

/* ------------------------------------------------------------------ updating a file */

static int get( t_zdf_file* f, void* p, size_t n ) { return fread(p, 1, n, f->fp) == n; }

/* record header of the record at the file position: id, name (malloc'ed), payload length */
static int get_header( t_zdf_file* f, uint32_t* id, char** name, uint64_t* length )
{
	uint32_t n;
	*name = NULL;
	if (!get(f, id, 4) || !get(f, &n, 4)) return 0;
	char* s = malloc(pad4(n) + 1);
	if (!s) return 0;
	if (pad4(n) && !get(f, s, pad4(n))) { free(s); return 0; }
	s[n] = 0;
	if (!get(f, length, 8)) { free(s); return 0; }
	*name = s;
	return 1;
}

size_t zdf_open_dataset( t_zdf_file* zdf, t_zdf_dataset* ds )
{
	if (fseek(zdf->fp, 4, SEEK_SET)) return (size_t) -1;
	for (;;) {
		uint32_t id; char* name; uint64_t length;
		if (!get_header(zdf, &id, &name, &length)) return 0;          /* ran off the end: not there */
		const int hit = (id == REC_CDSET_START || id == REC_DATASET) && !strcmp(ds->name, name);
		free(name);
		if (hit) break;
		if (fseeko(zdf->fp, (off_t) length, SEEK_CUR)) return 0;
	}
	/* the dataset header follows */
	off_t at = ftello(zdf->fp);
	uint32_t id32; int32_t type;
	if (at < 0 || !get(zdf, &id32, 4) || !get(zdf, &type, 4) || !get(zdf, &ds->ndims, 4)) return 0;
	ds->offset = (uint64_t) at;
	ds->id = id32;
	ds->data_type = (enum zdf_data_type) type;
	for (unsigned i = 0; i < ds->ndims; i++) if (!get(zdf, &ds->count[i], 8)) return 0;
	if (fseeko(zdf->fp, 0, SEEK_END)) return 0;
	return 1;
}

int zdf_extend_dataset( t_zdf_file* zdf, t_zdf_dataset* ds, uint64_t* new_count )
{
	for (unsigned i = 0; i < ds->ndims; i++) {
		if (new_count[i] < ds->count[i]) {
			fprintf(stderr, "(*error*) Invalid value for zdf_extend_dataset.\n");
			fprintf(stderr, "(*error*) New size is smaller than original size.\n");
			return -1;
		}
		ds->count[i] = new_count[i];         /* like the reference: dimensions before the bad one stay changed */
	}
	if (fseeko(zdf->fp, (off_t) ds->offset, SEEK_SET)) return 0;
	if (!put_ds_header(zdf, ds)) return 0;
	if (fseeko(zdf->fp, 0, SEEK_END)) return 0;
	return 1;
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
