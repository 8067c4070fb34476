/* zpic-b200 :: em2d electromagnetic fields (reference em2d/emf.h) */
#ifndef ZPIC_B200_EM2D_EMF_H
#define ZPIC_B200_EM2D_EMF_H

#include "zpic.h"
#include "current.h"

enum emf_fld_type { EMF_FLD_TYPE_NONE, EMF_FLD_TYPE_UNIFORM, EMF_FLD_TYPE_CUSTOM };

/* externally imposed fields (reference emf.h:28-44) */
typedef struct EMF_ExternalField {
	enum emf_fld_type E_type, B_type;
	float3 E_0, B_0;
	float3 (*E_custom)(int, float, int, float, void*);
	float3 (*B_custom)(int, float, int, float, void*);
	void *E_custom_data, *B_custom_data;
	float3 *E_part_buf, *B_part_buf;
} t_emf_ext_fld;

/* initial field values (reference emf.h:50-65) */
typedef struct EMF_InitialField {
	enum emf_fld_type E_type, B_type;
	float3 E_0, B_0;
	float3 (*E_custom)(int, float, int, float, void*);
	float3 (*B_custom)(int, float, int, float, void*);
	void *E_custom_data, *B_custom_data;
} t_emf_init_fld;

enum emf_diag { EFLD, BFLD, EPART, BPART };

/* E/B grids, same geometry as t_current (reference emf.h:83-120) */
typedef struct EMF {
	float3 *E, *B;
	float3 *E_buf, *B_buf;
	float3 *E_part, *B_part;
	int nx[2];
	int nrow;
	int gc[2][2];
	float box[2];
	float dx[2];
	float dt;
	int iter;
	int moving_window;
	int n_move;
	t_emf_ext_fld ext_fld;
} t_emf;

enum emf_laser_type { PLANE, GAUSSIAN };

/* laser pulse description (reference emf.h:135-157) */
typedef struct EMF_Laser {
	enum emf_laser_type type;
	float start, fwhm, rise, flat, fall;
	float a0, omega0, polarization;
	float W0, focus, axis;
} t_emf_laser;

/* device reduction, 6 doubles (reference emf.c:729-750) */
void emf_get_energy( const t_emf *emf, double energy[] );
/* replaces em2d/emf.c:56-113 */
void emf_new( t_emf *emf, int nx[], float box[], const float dt );
/* replaces em2d/emf.c:123-142 */
void emf_delete( t_emf *emf );
/* replaces em2d/emf.c:368-485 */
void emf_report( const t_emf *emf, const char field, const int fc );
/* host double-precision launch, then device refresh (reference emf.c:242-350) */
void emf_add_laser( t_emf* const emf, t_emf_laser* laser );
/* replaces em2d/emf.c:922-980 */
void emf_init_fld( t_emf* const emf, t_emf_init_fld* init_fld );
/* replaces em2d/emf.c:764-831 */
void emf_set_ext_fld( t_emf* const emf, t_emf_ext_fld* ext_fld );
/* device: yee_b, yee_e, yee_b, guard refresh, ext. fields, window shift (reference emf.c:688-716) */
void emf_advance( t_emf *emf, const t_current *current );
/* replaces em2d/emf.c:34-37 */
double emf_time( void );

#endif
