/* zpic-b200 :: em1d electromagnetic fields (reference em1d/emf.h) */
#ifndef ZPIC_B200_EM1D_EMF_H
#define ZPIC_B200_EM1D_EMF_H

#include "zpic.h"
#include "current.h"

enum emf_fld_type { EMF_FLD_TYPE_NONE, EMF_FLD_TYPE_UNIFORM, EMF_FLD_TYPE_CUSTOM };

typedef struct EMF_ExternalField {
	enum emf_fld_type E_type, B_type;
	float3 E_0, B_0;
	float3 (*E_custom)(int, float, void*);
	float3 (*B_custom)(int, float, void*);
	void *E_custom_data, *B_custom_data;
	float3 *E_part_buf, *B_part_buf;
} t_emf_ext_fld;

typedef struct EMF_InitialField {
	enum emf_fld_type E_type, B_type;
	float3 E_0, B_0;
	float3 (*E_custom)(int, float, void*);
	float3 (*B_custom)(int, float, void*);
	void *E_custom_data, *B_custom_data;
} t_emf_init_fld;

enum emf_diag { EFLD, BFLD, EPART, BPART };
/* reference em1d/emf.h:83-87 */
enum emf_boundary { EMF_BC_NONE, EMF_BC_PERIODIC, EMF_BC_OPEN };

/* reference em1d/emf.h:94-132 */
typedef struct EMF {
	float3 *E, *B;
	float3 *E_buf, *B_buf;
	float3 *E_part, *B_part;
	int nx;
	int gc[2];
	float box;
	float dx;
	float dt;
	int iter;
	int moving_window;
	int n_move;
	enum emf_boundary bc_type;
	float3 mur_fld[2];      /* state of the first-order Mur boundary (lower / upper) */
	float3 mur_tmp[2];
	t_emf_ext_fld ext_fld;
} t_emf;

/* plane-wave laser pulse (reference em1d/emf.h:139-154) */
typedef struct EMF_Laser {
	float start, fwhm, rise, flat, fall;
	float a0, omega0, polarization;
} t_emf_laser;

/* replaces em1d/emf.c:590-609 */
void emf_get_energy( const t_emf *emf, double energy[] );
/* replaces em1d/emf.c:56-114 */
void emf_new( t_emf *emf, int nx, float box, const float dt );
/* replaces em1d/emf.c:124-142 */
void emf_delete( t_emf *emf );
/* replaces em1d/emf.c:278-365 */
void emf_report( const t_emf *emf, const char field, const int fc );
/* replaces em1d/emf.c:191-260 */
void emf_add_laser( t_emf* const emf, t_emf_laser* laser );
/* replaces em1d/emf.c:764-811 */
void emf_init_fld( t_emf* const emf, t_emf_init_fld* init_fld );
/* replaces em1d/emf.c:623-686 */
void emf_set_ext_fld( t_emf* const emf, t_emf_ext_fld* ext_fld );
/* device: yee_b, yee_e, Mur boundary, yee_b, guards, ext. fields, window shift (reference em1d/emf.c:548-590) */
void emf_advance( t_emf *emf, const t_current *current );
/* replaces em1d/emf.c:34-37 */
double emf_time( void );

#endif
