/* zpic-b200 :: ZDF output (reference em2d/zdf.h).
 * Writer for the self-describing little-endian "ZDF1" container the reference uses
 * for every diagnostic.  Files written here are byte-compatible
 * with the reference writer, so python/lib/zdf.py and zpic_b200/zdf.py read both.
 * Public types keep the reference's field order (they are part of its C API). */
#ifndef ZPIC_B200_ZDF_H
#define ZPIC_B200_ZDF_H

#include <stdint.h>
#include <stdio.h>

#define zdf_max_dims 3
/* strings and 8-bit vectors are padded to this many bytes; length of the "ZDF1" magic (reference zdf.h) */
#define BYTES_PER_ZDF_UNIT 4
#define ZDF_MAGIC_LENGTH 4

enum zdf_data_type {
	zdf_null, zdf_int8, zdf_uint8, zdf_int16, zdf_uint16, zdf_int32, zdf_uint32,
	zdf_int64, zdf_uint64, zdf_float32, zdf_float64
};
enum zdf_file_access_mode { ZDF_CREATE, ZDF_READ, ZDF_UPDATE };

typedef struct ZDF_File {
	FILE *fp;
	enum zdf_file_access_mode mode;
	uint32_t ndatasets;
} t_zdf_file;

typedef struct ZDF_Dataset {
	char* name;
	enum zdf_data_type data_type;
	uint32_t ndims;
	uint64_t count[zdf_max_dims];
	void* data;
	uint64_t id;
	uint64_t offset;
} t_zdf_dataset;

/* one piece of a chunked dataset (reference zdf.h:117-122): `count` elements per direction placed at `start`
 * with `stride` inside the dataset; data is contiguous */
typedef struct ZDF_Chunk {
	uint64_t count[zdf_max_dims];
	uint64_t start[zdf_max_dims];
	uint64_t stride[zdf_max_dims];
	void* data;
} t_zdf_chunk;

enum zdf_axis_type { zdf_linear, zdf_log10, zdf_log2 };

typedef struct ZDF_GridAxis {
	char* name;
	enum zdf_axis_type type;
	double min, max;
	char* label;
	char* units;
} t_zdf_grid_axis;

typedef struct ZDF_GridInfo {
	char* name;
	uint32_t ndims;
	uint64_t count[zdf_max_dims];
	char* label;
	char* units;
	t_zdf_grid_axis *axis;
} t_zdf_grid_info;

typedef struct ZDF_Iteration {
	char* name;
	int32_t n;
	double t;
	char* time_units;
} t_zdf_iteration;

typedef struct ZDF_PartInfo {
	char* name;
	char* label;
	uint64_t np;
	uint32_t nquants;
	char** quants;
	char** qlabels;
	char** qunits;
} t_zdf_part_info;

/* particle tracks metadata (reference zdf.h:192-203) */
typedef struct ZDF_TrackInfo {
	char* name;
	char* label;
	uint32_t ntracks;
	uint32_t ndump;
	uint32_t niter;
	uint32_t nquants;
	char** quants;
	char** qlabels;
	char** qunits;
} t_zdf_track_info;

/* replaces em1d/zdf.c:135-155 */
size_t zdf_sizeof( enum zdf_data_type data_type );
/* replaces em1d/zdf.c:183-270 */
int zdf_open_file( t_zdf_file* zdf, const char* filename, enum zdf_file_access_mode mode );
/* replaces em1d/zdf.c:166-174 */
int zdf_close_file( t_zdf_file* zdf );
/* replaces em1d/zdf.c:893-909 */
size_t zdf_add_string( t_zdf_file* zdf, const char* name, const char* str );
/* replaces em1d/zdf.c:918-934 */
size_t zdf_add_int32( t_zdf_file* zdf, const char* name, const int32_t value );
/* replaces em1d/zdf.c:943-956 */
size_t zdf_add_double( t_zdf_file* zdf, const char* name, const double value );
/* replaces em1d/zdf.c:969-987 */
size_t zdf_add_iteration( t_zdf_file* zdf, const t_zdf_iteration* iter );
/* replaces em1d/zdf.c:1021-1056 */
size_t zdf_add_grid_info( t_zdf_file* zdf, const t_zdf_grid_info* grid );
/* replaces em1d/zdf.c:1083-1109 */
size_t zdf_add_part_info( t_zdf_file* zdf, const t_zdf_part_info* part );
/* replaces em1d/zdf.c:1241-1268 */
size_t zdf_add_dataset( t_zdf_file* zdf, t_zdf_dataset* dataset );
/* raw vector of `len` elements: 8-bit data is zero-padded to a multiple of 4 bytes, wider types are not;
 * returns the bytes written, 0 on error or when len is 0 (replaces em1d/zdf.c:735-756) */
size_t zdf_vector_write( t_zdf_file* zdf, const void* data, enum zdf_data_type data_type, size_t len );
/* replaces em1d/zdf.c:1135-1161 */
size_t zdf_add_track_info( t_zdf_file* zdf, const t_zdf_track_info* tracks );
/* Chunked datasets: a start record with the dataset header, any number of chunk records
 * ("<id>-chunk": dataset id, count, start, stride, data), an end record ("<id>-end").
 * replaces em1d/zdf.c:1277-1296 */
size_t zdf_start_cdset( t_zdf_file* zdf, t_zdf_dataset* dataset );
/* replaces em1d/zdf.c:1303-1307 */
size_t size_zdf_chunk_header( const t_zdf_dataset* dataset );
/* replaces em1d/zdf.c:1316-1359 */
size_t zdf_write_chunk_header( t_zdf_file* zdf, t_zdf_dataset* dataset, t_zdf_chunk* chunk );
/* replaces em1d/zdf.c:1369-1385 */
size_t zdf_write_cdset( t_zdf_file* zdf, t_zdf_dataset* dataset, t_zdf_chunk* chunk );
/* replaces em1d/zdf.c:1393-1409 */
size_t zdf_end_cdset( t_zdf_file* zdf, t_zdf_dataset* dataset );
/* find a (chunked) dataset by name in a file opened with ZDF_READ / ZDF_UPDATE and load its header; the file
 * position ends up at the end of the file. 1 on success; 0 on read errors, which includes running off the
 * end of the file when there is no such dataset (replaces em1d/zdf.c:1412-1457) */
size_t zdf_open_dataset( t_zdf_file* zdf, t_zdf_dataset* dataset );
/* grow the dimensions recorded in the dataset header in place; 1 on success, -1 if a dimension would
 * shrink (replaces em1d/zdf.c:1470-1488) */
int zdf_extend_dataset( t_zdf_file* zdf, t_zdf_dataset* dataset, uint64_t* new_count );
/* replaces em1d/zdf.c:1500-1528 */
int zdf_open_grid_file( t_zdf_file *file, const t_zdf_grid_info *info,
                        const t_zdf_iteration *iteration, char const path[] );
/* replaces em1d/zdf.c:1540-1562 */
int zdf_save_grid( const void* data, enum zdf_data_type data_type, const t_zdf_grid_info *info,
                   const t_zdf_iteration *iteration, char const path[] );
/* replaces em1d/zdf.c:1572-1600 */
int zdf_open_part_file( t_zdf_file *file, t_zdf_part_info *info,
                        const t_zdf_iteration *iteration, char const path[] );
/* replaces em1d/zdf.c:1610-1624 */
int zdf_add_quant_part_file( t_zdf_file *zdf, const char *name, const float* data, const uint64_t np );

#endif
