/* vicres.h -- C ABI of the B200-native VIC-Res land-surface step (and routing pass).
 *
 * Drop-in boundary for the reference's per-cell call seam
 *
 *     ErrorFlag = full_energy(cellnum, rec, &atmos[rec], &all_vars, dmy, &global_param,
 *                             &lake_con, &soil_con, veg_con, veg_hist);     runoff/model/vicNl.c:301
 *     ErrorFlag = put_data(&all_vars, &atmos[rec], &soil_con, veg_con, &lake_con,
 *                          out_data_files, out_data, &save_data, &dmy[rec], rec);  vicNl.c:306
 *
 * which the reference runs inside `while(cell){ for(rec) }` (vicNl.c:197-335).  Here the two
 * loops are batched: all cells of a handle advance together through a block of records on the
 * GPU; state stays resident in HBM between blocks.  Plain C types only -- no torch / C++ types.
 *
 * Layout conventions (struct-of-arrays, cell index fastest):
 *   per-cell scalar  a[c]                      c in [0,C)
 *   per-layer        a[l*C + c]                l in [0,nlayer)
 *   per-band         a[b*C + c]                b in [0,nbands)
 *   per-tile         CSR: tiles of cell c are t in [tile_ptr[c], tile_ptr[c+1]); the LAST tile of
 *                    a cell may be the bare-soil tile (veg_class == veglib.nclass), exactly as
 *                    veg_con[Nveg] in the reference (read_vegparam.c:431-434)
 *   forcing          f[field][rec*C + c]       (what atmos_data_struct holds, vicNl_def.h:1090-1113)
 *   output           out[(v*nrec + rec)*C + c] v = position in vr_options.out_vars
 *
 * Every function returns VR_OK (0) or a negative vr_status; vr_last_error() gives the text.
 * The library never calls exit() (the reference's nrerror/vicerror do).
 */
#ifndef VICRES_H
#define VICRES_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VR_MAX_LAYERS 3   /* MAX_LAYERS, vicNl_def.h:171 */
#define VR_MAX_BANDS 10   /* MAX_BANDS,  vicNl_def.h:173 */
#define VR_ZWT_NPOINTS 11      /* MAX_ZWTVMOIST, vicNl_def.h */
#define VR_MAX_OUT 64

typedef enum vr_status {
  VR_OK = 0,
  VR_ERR_ARG = -1,          /* bad argument / inconsistent sizes */
  VR_ERR_UNSUPPORTED = -2,  /* option outside the water-balance hot path */
  VR_ERR_CUDA = -3,         /* CUDA runtime failure; no CPU fallback exists */
  VR_ERR_STATE = -4,        /* call sequence error (e.g. step before set_cells) */
  VR_ERR_CELL = -999        /* == the reference's ERROR (vicNl_def.h:167): a cell failed */
} vr_status;

/* Forcing fields consumed by the step, in the order of vr_forcing.field[] */
enum {
  VR_F_AIR_TEMP = 0, /* C      atmos.air_temp  */
  VR_F_PREC,         /* mm     atmos.prec      */
  VR_F_WIND,         /* m/s    atmos.wind      */
  VR_F_VP,           /* Pa     atmos.vp        */
  VR_F_VPD,          /* Pa     atmos.vpd       */
  VR_F_PRESSURE,     /* Pa     atmos.pressure  */
  VR_F_DENSITY,      /* kg/m3  atmos.density   */
  VR_F_SHORTWAVE,    /* W/m2   atmos.shortwave */
  VR_F_LONGWAVE,     /* W/m2   atmos.longwave  */
  VR_NFORCING
};

/* Output variables (subset of the reference's OUT_* list, vicNl_def.h:486-659, aggregated over
 * tiles x bands exactly as put_data.c:298-632).  VR_OUT_SOIL_LIQ0+l is layer l. */
enum {
  VR_OUT_PREC = 0, VR_OUT_EVAP, VR_OUT_RUNOFF, VR_OUT_BASEFLOW, VR_OUT_WDEW,
  VR_OUT_SOIL_LIQ0, VR_OUT_SOIL_LIQ1, VR_OUT_SOIL_LIQ2,
  VR_OUT_NET_SHORT, VR_OUT_R_NET, VR_OUT_EVAP_CANOP, VR_OUT_TRANSP_VEG, VR_OUT_EVAP_BARE,
  VR_OUT_SUB_CANOP, VR_OUT_SUB_SNOW, VR_OUT_AERO_RESIST, VR_OUT_SURF_TEMP, VR_OUT_ALBEDO,
  VR_OUT_REL_HUMID, VR_OUT_IN_LONG, VR_OUT_AIR_TEMP, VR_OUT_WIND,
  VR_OUT_SWE, VR_OUT_SNOW_DEPTH, VR_OUT_SNOW_CANOPY, VR_OUT_SNOW_COVER,
  /* extras beyond the reference's default fluxes+snow files */
  VR_OUT_WATER_ERROR, VR_OUT_INFLOW, VR_OUT_SNOW_MELT, VR_OUT_RAINF, VR_OUT_SNOWF,
  VR_OUT_NET_LONG, VR_OUT_ASAT, VR_OUT_SOIL_WET, VR_OUT_ROOTMOIST, VR_OUT_GRND_FLUX,
  VR_OUT_DELTAH, VR_OUT_AERO_COND, VR_OUT_SHORTWAVE, VR_OUT_LONGWAVE,
  VR_OUT_DELSOILMOIST, VR_OUT_DELSWE, VR_OUT_DELINTERCEPT, VR_OUT_SNOW_SURF_TEMP,
  VR_OUT_SNOW_PACK_TEMP, VR_OUT_SALBEDO,
  /* potential evaporation of the six reference covers (compute_pot_evap.c; needs vr_veglib.LAI/albedo) */
  VR_OUT_PET_SATSOIL, VR_OUT_PET_H2OSURF, VR_OUT_PET_SHORT, VR_OUT_PET_TALL, VR_OUT_PET_NATVEG, VR_OUT_PET_VEGNOCR,
  /* water table position, cm (compute_zwt.c; needs vr_cells.zwt_zwt/zwt_moist) */
  VR_OUT_ZWT, VR_OUT_ZWT_LUMPED,
  /* the record's own forcing values as put_data.c:271-292 passes them on: kg/m3, kPa, kPa, kPa, kg/kg */
  VR_OUT_DENSITY, VR_OUT_PRESSURE, VR_OUT_VP, VR_OUT_VPD, VR_OUT_QAIR,
  VR_NOUTVAR
};

/* Run-wide options: the subset of option_struct / global_param_struct (vicNl_def.h:722-932)
 * that the water-balance path reads.  Defaults as initialize_global.c:146-211. */
typedef struct vr_options {
  int32_t nlayer;           /* options.Nlayer, 2 or 3                                   */
  int32_t nbands;           /* options.SNOW_BAND, 1..10                                 */
  int32_t dt_hours;         /* global_param.dt (TIME_STEP), 1..24                       */
  int32_t snow_step_hours;  /* options.SNOW_STEP, a divisor of dt_hours: NF = dt_hours / snow_step_hours passes of the
                               snow model per record (surface_fluxes.c:385-396, get_global_param.c:761-770)         */
  double  max_snow_temp;    /* global_param.MAX_SNOW_TEMP                               */
  double  min_rain_temp;    /* global_param.MIN_RAIN_TEMP                               */
  double  wind_h;           /* global_param.wind_h (WIND_H)                             */
  int32_t n_out;            /* number of output variables requested                     */
  int32_t out_vars[VR_MAX_OUT]; /* VR_OUT_* ids, in output order                        */
  int32_t cells_per_block_hint; /* records per pipelined H2D/compute/D2H chunk of vr_step_block; 0 = default (16) */
  int32_t moistfract;       /* options.MOISTFRACT: OUT_SOIL_LIQ / OUT_SOIL_MOIST as volume fractions (each tile's layer moisture
                               divided by depth * 1000 before the area weighting, put_data.c:906-911)                  */
  int32_t reserved[6];
} vr_options;

/* Vegetation library incl. nothing but real classes; the bare-soil pseudo class (index nclass) is
 * derived per cell from soil roughness as read_soilparam.c:551-554 does. Monthly arrays [cls*12+m]. */
typedef struct vr_veglib {
  int32_t nclass;
  const int32_t *overstory;      /* [nclass] veg_lib.overstory  */
  const double  *rarc;           /* [nclass] */
  const double  *rmin;           /* [nclass] */
  const double  *wind_h;         /* [nclass] */
  const double  *RGL;            /* [nclass] (float in the reference; pass the float value) */
  const double  *rad_atten;      /* [nclass] */
  const double  *wind_atten;     /* [nclass] */
  const double  *trunk_ratio;    /* [nclass] */
  const double  *roughness;      /* [nclass*12] */
  const double  *displacement;   /* [nclass*12] */
  const double  *LAI;            /* [nclass*12] veg_lib.LAI    -- read only for VR_OUT_PET_NATVEG / _VEGNOCR; else may be NULL */
  const double  *albedo;         /* [nclass*12] veg_lib.albedo -- the same                                                     */
} vr_veglib;

/* Cells: soil_con_struct + veg_con_struct[] restated as SoA (vicNl_def.h:937-1031). All values
 * are the DERIVED ones the reference holds after read_soilparam/read_vegparam/read_snowband
 * (float-rounded depths, max_moist, Wcr, Wpwp ...); vicres_host.h parsers produce them. */
typedef struct vr_cells {
  int64_t ncell;
  /* per cell */
  const double *lat, *elevation, *b_infilt, *Ds, *Dsmax, *Ws, *c_expt, *avg_temp, *dp,
               *rough, *snow_rough;
  /* per layer [l*C+c] */
  const double *expt, *Ksat, *depth, *max_moist, *Wcr, *Wpwp, *resid_moist, *init_moist,
               *bulk_density, *soil_density, *bulk_dens_min, *soil_dens_min, *quartz, *organic,
               *porosity;
  /* per band [b*C+c] */
  const double *AreaFract, *Tfactor, *Pfactor;
  /* tiles (CSR) */
  const int64_t *tile_ptr;       /* [C+1] */
  const int32_t *tile_class;     /* [T] index into veglib, or veglib.nclass for bare soil */
  const double  *tile_Cv;        /* [T] */
  const double  *tile_root;      /* [T*nlayer] root[l] of tile t at t*nlayer+l (float values) */
  const double  *tile_LAI;       /* [T*12] veg_con.LAI   (monthly climatology)     */
  const double  *tile_albedo;    /* [T*12] veg_con.albedo                           */
  const double  *tile_vegcover;  /* [T*12] veg_con.vegcover                         */
  /* soil moisture v water table curves, soil_con.zwtvmoist_zwt / _moist (read_soilparam.c:822-921): curve k =
   * layer k for k < nlayer, nlayer = top layers lumped, nlayer+1 = whole column; VR_ZWT_NPOINTS points each;
   * value of point i of curve k of cell c at ((k*VR_ZWT_NPOINTS + i)*C + c).  Read only for VR_OUT_ZWT*; else NULL. */
  const double  *zwt_zwt, *zwt_moist;
} vr_cells;

/* One block of records for all cells. Pointers are HOST memory for vr_step_block and DEVICE
 * memory for vr_step_block_device. */
typedef struct vr_forcing {
  int32_t nrec;
  const int32_t *month;        /* [nrec] dmy.month 1..12            (host memory in both calls) */
  const int32_t *day_in_year;  /* [nrec] dmy.day_in_year 1..366     (host memory in both calls) */
  const double  *field[VR_NFORCING];   /* each [nrec*C]; with NF > 1 each [nrec*(NF+1)*C]: per record the NF
                                          sub-step values, then the record's own (atmos-><field>[0..NF-1], [NR]) */
  const int32_t *snowflag;     /* [nrec*C] atmos->snowflag[NR], or NULL; read only when NF > 1 (same memory space
                                  as field[])                                                                  */
} vr_forcing;

typedef struct vr_handle vr_handle;

/* lifecycle */
int  vr_create(const vr_options *opt, int device, vr_handle **out);
void vr_destroy(vr_handle *h);
const char *vr_last_error(const vr_handle *h);   /* h may be NULL: error of a failed vr_create */
void vr_default_options(vr_options *opt, int nlayer);  /* reference defaults + default 20+L-1+4 outputs */

/* parameters: replaces read_veglib / read_soilparam / read_vegparam / read_snowband hand-off */
int  vr_set_veglib(vr_handle *h, const vr_veglib *lib);
int  vr_set_cells(vr_handle *h, const vr_cells *cells);

/* Optional ensemble axis (toolbox calibration / eFAST / EET): cell c reads forcing
 * column col[c] of forcing arrays that have ncol columns, so P parameter sets x N basin cells are
 * P*N "cells" here that share N forcing columns (the reference re-reads the same forcing file in
 * every vicNl run, toolbox/calibration/basin_calibration_eNSGAII.py:608-641). col == NULL restores
 * the identity map. tair0 of vr_init_state is then indexed by forcing column as well. */
int  vr_set_forcing_map(vr_handle *h, const int64_t *col /* [C] host */, int64_t ncol);

/* initialize_model_state(): surf_temp = atmos[0].air_temp[NR] per cell (vicNl.c:260-264) and the
 * put_data(rec = -nrecs) storage initialisation (vicNl.c:287) */
int  vr_init_state(vr_handle *h, const double *tair0 /* [C] host */);

/* the step: full_energy + put_data for nrec consecutive records of every cell.
 * out: [n_out][nrec][C]; cell_status: [C] 0 or the first failing record + 1 (may be NULL). */
int  vr_step_block(vr_handle *h, const vr_forcing *f, double *out_host, int32_t *cell_status_host);
int  vr_step_block_device(vr_handle *h, const vr_forcing *f_dev, double *out_dev, void *cuda_stream);
int  vr_synchronize(vr_handle *h);
/* Compatibility shim for record-by-record callers and single-cell tests: ONE record, the pair of calls of the reference's
 * record loop, `full_energy(cellnum, rec, &atmos[rec], &all_vars, dmy, &global_param, &lake_con, &soil_con, veg_con, veg_hist);
 * put_data(&all_vars, &atmos[rec], &soil_con, veg_con, &lake_con, out_data_files, out_data, &save_data, &dmy[rec], rec);`
 * (vicNl.c:301-306), for every cell of the handle.  rec must be the next record of the run (records done so far:
 * a skipped or repeated record is VR_ERR_STATE, where the reference would silently use stale storage terms).
 * atmos: [VR_NFORCING][NF + 1 or 1][C] host values of this record in VR_F_* order (atmos-><field>[0..NF-1], [NR]);
 * snowflag [C] or NULL; month 1..12 and day_in_year of dmy[rec]; out: [n_out][C]; cell_status [C] or NULL.
 * Returns VR_OK, or VR_ERR_CELL as vr_step_block does (the reference's ERROR return of full_energy). */
int  vr_full_energy_compat(vr_handle *h, int64_t rec, int32_t month, int32_t day_in_year, const double *atmos,
                           const int32_t *snowflag, double *out, int32_t *cell_status);

/* state hand-off (write_model_state / read_initial_model_state analogue): opaque flat blob */
int64_t vr_state_size(const vr_handle *h);                 /* bytes */
int  vr_get_state(vr_handle *h, void *blob_host);
int  vr_set_state(vr_handle *h, const void *blob_host);

/* The reference's own state file (write_model_state.c:100-300 with the header of open_state_file.c; ASCII "%f"
 * columns or BINARY_STATE_FILE): per cell `cellnum Nveg Nbands dz_node[3] Zsum_node[3]`, then per vegetation tile
 * (the bare-soil tile last) x snow band `veg band moist[L] ice[L] [Wdew] last_snow MELTING coverage swq surf_temp
 * surf_water pack_temp pack_water density coldcontent snow_canopy T[3]`.  gridcel: [C] the cells' numbers.
 * vr_read_state_file is read_initial_model_state + the state branch of initialize_model_state (:233-330, 590-633):
 * moisture capped at max_moist, snow pack terms made consistent, depth from swq / density, LongUnderOut from T[0],
 * Tfoliage from tair0 (atmos[0].air_temp of the restarted run), then the put_data(rec = -nrecs) storages. */
int  vr_write_state_file(vr_handle *h, const char *path, int32_t year, int32_t month, int32_t day, int32_t binary,
                         const int32_t *gridcel);
int  vr_read_state_file(vr_handle *h, const char *path, int32_t binary, const double *tair0, const int32_t *gridcel);

/* introspection used by bench/tests */
int64_t vr_num_units(const vr_handle *h);      /* active tile x band pairs */
int64_t vr_num_launches(const vr_handle *h);   /* kernels launched by this handle so far */
int  vr_last_kernel_ms(vr_handle *h, float *ms); /* CUDA-event time of all kernels of the last block of records */
/* CUDA-event times, summed over every chunk of records launched since the previous call of this function (at most
 * 8192 chunks are kept): units_ms = the step kernels of the chunk (k_forcing_slots, k_classify, then per record
 * k_front, k_back and -- on the snow stream -- k_snow, k_back_list), cells_ms = k_cells (put_data of all cells);
 * launch_pairs = chunks, records = the records they covered.  Synchronises with the launches; resets the sums. */
int  vr_kernel_times(vr_handle *h, double *units_ms, double *cells_ms, int64_t *launch_pairs, int64_t *records);
/* diagnostic: measured FP64 fused-multiply-add thread-instructions per second of the device (independent
 * chains on every SM), the denominator of bench.py's FP64 roofline */
int  vr_diag_fp64_peak(int device, double *dfma_per_s);
/* diagnostic, since the last call: out[0..2] = summed durations (ms) of the step kernels k_front, k_snow, k_back
   (timed only when VICRES_TIME_KERNELS=1 was in the environment at vr_create), out[3] = records timed, out[4] =
   unit-records that ran the snow instance, out[5] = unit-records in total, out[6] = summed durations of k_back_list */
int  vr_step_diag(vr_handle *h, double *out7);
/* diagnostic: z[i] = the device's power function (vic_physics.cuh m_pow) of x[i], y[i]; host pointers */
int  vr_diag_pow(int device, int64_t n, const double *x, const double *y, double *z);
/* TFALLBACK counters summed over all units, as put_data.c:658-664 prints them:
 * counts[0] = Tsnowsurf (snow_melt.c:361-366), counts[1] = Tfoliage (snow_intercept.c:418-423) */
int  vr_fallback_counts(vr_handle *h, int64_t counts[2]);

#ifdef __cplusplus
}
#endif
#endif /* VICRES_H */
