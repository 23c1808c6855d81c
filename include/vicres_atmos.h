/* vicres_atmos.h -- C ABI of the forcing preparation: the raw forcing-file columns of every cell
 * (PREC TMAX TMIN WIND daily, or PREC AIR_TEMP WIND sub-daily) -> the nine atmos fields the cell step
 * consumes (include/vicres.h, VR_F_*), on the GPU.
 *
 * Replaces, for every cell at once, the reference's per-cell call (runoff/model/vicNl.c:239-241)
 *
 *     initialize_atmos(atmos, dmy, filep.forcing, veg_con, veg_hist, &soil_con, out_data_files, out_data);
 *
 * i.e. initialize_atmos.c:137-1818 with the MTCLIM 4.3 preprocessor behind it (mtclim_wrapper.c:69-278,
 * mtclim_vic.c:404-1395: calc_tair, calc_prcp, snowpack, calc_srad_humidity_iterative,
 * compute_srad_humidity_onetime, calc_pet, atm_pres, pulled_boxcar), calc_air_temperature.c:15-201
 * (HourlyT, hermite, hermint, set_max_min_hour), calc_longwave.c:7-76 and svp.c:7-22.
 *
 * The output arrays have the layout of vr_forcing.field[] (field f, record r, cell c at out[f][r*C + c]), so the
 * device variant's result goes straight into vr_step_block_device without touching the host.
 *
 * Supported forcing files (everything else -> VR_ERR_UNSUPPORTED, nothing is silently ignored):
 *   VR_ATMOS_DAILY_TMAX_TMIN    FORCE_DT 24, columns PREC TMAX TMIN [WIND]   (the bundled basin's format)
 *   VR_ATMOS_SUBDAILY_AIR_TEMP  FORCE_DT < 24, columns PREC AIR_TEMP [WIND]
 * each optionally with observed SHORTWAVE, LONGWAVE, VP, PRESSURE and DENSITY columns at the same FORCE_DT; SNOW_STEP == TIME_STEP
 * (NF == 1) or, for daily steps, below it (vr_atmos_prepare_sub); QAIR / REL_HUMID in place of VP and the split columns
 * RAINF + SNOWF, WIND_E + WIND_N (vr_atmos_raw).  There is no CPU path: without a CUDA device the calls return VR_ERR_CUDA.
 */
#ifndef VICRES_ATMOS_H
#define VICRES_ATMOS_H
#include <stdint.h>
#include "vicres.h"
#ifdef __cplusplus
extern "C" {
#endif

enum { VR_ATMOS_DAILY_TMAX_TMIN = 0, VR_ATMOS_SUBDAILY_AIR_TEMP = 1 };
/* option values as vicNl_def.h:208-223 */
enum { VR_VP_ITER_NONE = 0, VR_VP_ITER_ALWAYS = 1, VR_VP_ITER_ANNUAL = 2, VR_VP_ITER_CONVERGE = 3 };
enum { VR_LW_TVA = 0, VR_LW_ANDERSON, VR_LW_BRUTSAERT, VR_LW_SATTERLUND, VR_LW_IDSO, VR_LW_PRATA };
enum { VR_LW_CLOUD_BRAS = 0, VR_LW_CLOUD_DEARDORFF = 1 };

/* The part of global_param_struct / option_struct / param_set_struct that initialize_atmos reads.
 * Defaults (vr_atmos_default_options) as initialize_global.c:146-211. */
typedef struct vr_atmos_options {
  int32_t time_step;        /* TIME_STEP, hours                                                  */
  int32_t nrecs;            /* model records                                                     */
  int32_t startyear, startmonth, startday, starthour;
  int32_t layout;           /* VR_ATMOS_*                                                        */
  int32_t force_dt;         /* FORCE_DT of the file: 24 for the daily layout, a divisor of 24 below 24 otherwise */
  int32_t has_wind;         /* 0: no WIND column -> DEFAULT_WIND_SPEED 3.0 (initialize_atmos.c:681-689) */
  int32_t vp_interp;        /* VP_INTERP        default TRUE                                     */
  int32_t vp_iter;          /* VP_ITER          default VR_VP_ITER_ALWAYS                        */
  int32_t mtclim_swe_corr;  /* MTCLIM_SWE_CORR  default FALSE                                    */
  int32_t lw_type;          /* LW_TYPE          default VR_LW_PRATA                              */
  int32_t lw_cloud;         /* LW_CLOUD         default VR_LW_CLOUD_DEARDORFF                    */
  int32_t plapse;           /* PLAPSE           default TRUE                                     */
  int32_t alma_input;       /* ALMA_INPUT       default FALSE; TRUE: PREC is a rate (mm/s), temperatures are K, VP and
                             * PRESSURE are Pa (initialize_atmos.c:374-402)                      */
  double  min_wind_speed;   /* MIN_WIND_SPEED   default 0.1                                      */
  double  sw_prec_thresh;   /* SW_PREC_THRESH   default 0 (mm; compared with MTCLIM's cm value as the reference does) */
  int32_t reserved[8];      /* [0] cells per batch (0 = 32768); [1] = 1: untiled daily-radiation kernel,
                             * [2] = 1: solar kernel in the caller's cell order instead of latitude order (tests) */
  int32_t snow_step;        /* SNOW_STEP, hours: 0 = TIME_STEP.  Below TIME_STEP (daily runs only, get_global_param.c:761-765)
                             * every record carries NF = TIME_STEP / SNOW_STEP windows plus its own values         */
  int32_t pad_;
  double  max_snow_temp;    /* MAX_SNOW_TEMP    default 0.5: atmos->snowflag (initialize_atmos.c:1736-1753)        */
} vr_atmos_options;

void vr_atmos_default_options(vr_atmos_options *o);

/* Per-cell site values, as soil_con_struct holds them after read_soilparam (so lat/lng/elevation/time_zone_lng
 * are the FLOAT values widened to double; time_zone_lng = off_gmt*360/24, read_soilparam.c:934).  slope, aspect,
 * ehoriz, whoriz may be NULL = 0 (the reference hard-codes them, read_soilparam.c:936-939). */
typedef struct vr_atmos_cells {
  int64_t ncell;
  const double *lat, *lng, *time_zone_lng, *elevation, *annual_prec;
  const double *slope, *aspect, *ehoriz, *whoriz;
  const double *min_Tfactor;   /* the lowest Tfactor of the cell's snow bands (initialize_atmos.c:1737-1741), NULL = 0:
                                * read only for the snow flags */
} vr_atmos_cells;

/* Raw forcing columns, record r of cell c at col[r*C + c], nrec_forcing = nrecs*time_step/force_dt records
 * (what read_atmos_data returns per cell, vr_read_forcing_file). t_a = TMAX or AIR_TEMP, t_b = TMIN or NULL. */
typedef struct vr_atmos_raw {
  int64_t nrec_forcing;
  const double *prec, *t_a, *t_b, *wind;
  /* optional observed columns at the file's FORCE_DT, NULL when absent: SHORTWAVE (W/m2), LONGWAVE (W/m2), VP (kPa as in
   * the file; converted to Pa like initialize_atmos.c:404-414).  With them MTCLIM runs in its "observed" modes
   * (ctrl.insw / ctrl.invp, mtclim_vic.c:982-995, 1173-1185). */
  const double *shortwave, *longwave, *vp;
  /* optional observed PRESSURE (kPa as in the file) and DENSITY (kg/m3) columns (initialize_atmos.c:1032-1170): the one
   * that is missing is derived from the other and the air temperature, both missing = the PLAPSE estimate */
  const double *pressure, *density;
  /* optional humidity columns, used when VP is absent (initialize_atmos.c:773-866, 1176-1215): QAIR (kg/kg) and REL_HUMID (%).
   * With PRESSURE (QAIR) or in the AIR_TEMP layout (REL_HUMID) they become the observed VP before MTCLIM; in the daily
   * TMAX/TMIN layout without those they replace MTCLIM's daily vapour pressure afterwards, from the record's own pressure
   * or air temperature.  Sub-daily QAIR without PRESSURE is refused (the reference indexes atmos->pressure with a loop
   * variable left over from an earlier loop, :1302). */
  const double *qair, *rel_humid;
  /* optional split columns that REPLACE prec / wind when both of a pair are given (:424-460): RAINF + SNOWF, and
   * sqrt(WIND_E^2 + WIND_N^2) */
  const double *rainf, *snowf, *wind_e, *wind_n;
} vr_atmos_raw;

/* Number of local days MTCLIM runs over for a cell with this integer hour offset (Ndays or Ndays+1). */
int32_t vr_atmos_ndays(const vr_atmos_options *o);

/* host buffers in, host buffers out: out[f] is [nrecs*C] for f < VR_NFORCING */
int  vr_atmos_prepare(int device, const vr_atmos_options *o, const vr_atmos_cells *cells, const vr_atmos_raw *raw,
                      double *const out[VR_NFORCING]);
/* device buffers in (raw columns and cell arrays in DEVICE memory), device buffers out; runs on `cuda_stream`
 * (NULL = default stream) and returns after the launches are queued */
int  vr_atmos_prepare_device(int device, const vr_atmos_options *o, const vr_atmos_cells *cells_dev,
                             const vr_atmos_raw *raw_dev, double *const out_dev[VR_NFORCING], void *cuda_stream);
/* The same with SNOW_STEP < TIME_STEP (o->snow_step; NF = TIME_STEP / SNOW_STEP): out[f] is [nrecs][NF + 1][C] -- per record
 * the NF windows' values, then the record's own (atmos-><field>[0..NF-1], [NR]) -- which is the layout vr_forcing takes for
 * a sub-stepped snow model, and snowflag (may be NULL) [nrecs][C] = atmos->snowflag[NR].  With NF == 1 they are the calls
 * above plus the flags. */
int  vr_atmos_prepare_sub(int device, const vr_atmos_options *o, const vr_atmos_cells *cells, const vr_atmos_raw *raw,
                          double *const out[VR_NFORCING], int32_t *snowflag);
int  vr_atmos_prepare_sub_device(int device, const vr_atmos_options *o, const vr_atmos_cells *cells_dev,
                                 const vr_atmos_raw *raw_dev, double *const out_dev[VR_NFORCING], int32_t *snowflag_dev,
                                 void *cuda_stream);
/* CUDA-event time of the kernels of the last vr_atmos_prepare* call on this thread: [0] solar geometry tables,
 * [1] MTCLIM daily pass, [2] sub-daily disaggregation; launches = kernels launched */
int  vr_atmos_last_times(double ms[3], int64_t *launches);
const char *vr_atmos_last_error(void);
/* diagnostic: the device's correctly rounded acos(x[i]) and cos(y[i]), y in [0, pi] (vic_cr_trig.cuh); host pointers */
int  vr_atmos_diag_cr(int device, int64_t n, const double *x, double *acos_out, const double *y, double *cos_out);

/* diagnostic: the solar-geometry tables of n cells (lng = time_zone_lng, i.e. no sub-hour offset, unless given):
 * tab[(row*n + c)], rows 0..364 ttmax0, 365..729 flat potential radiation, 730..1094 slope potential radiation,
 * 1095..1459 daylength, then 365*24 rows of hourly radiation fractions (yearday-major); host pointers.
 * exact_walk != 0 runs the hour-angle walk with the reference's libm call per step instead of the production
 * kernel's rotation / table form (vic_atmos.cuh, solar_day_t) */
int  vr_atmos_diag_solar(int device, int64_t n, const double *lat, const double *elevation, const double *lng,
                         const double *time_zone_lng, double *tab, int32_t exact_walk);

#ifdef __cplusplus
}
#endif
#endif
