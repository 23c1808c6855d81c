/* vicres_files.h -- host-side readers/writers for the reference's own file formats, so that the B200
 * path is a drop-in for vicNl's parameter hand-off and flux output:
 *   soil parameter file     runoff/model/read_soilparam.c:8-640   (one row per cell; derived max_moist, Wcr,
 *                           Wpwp, porosity ..., depths rounded to mm through a float, :342)
 *   vegetation library      runoff/model/read_veglib.c:8-190
 *   vegetation parameters   runoff/model/read_vegparam.c:8-340 + calc_root_fraction.c:8-150
 *   snow band file          runoff/model/read_snowband.c:8-125
 *   flux output             runoff/model/write_data.c ASCII branch ("%04i\t%02i\t%02i\t" + "%.4f" columns)
 * The readers produce exactly the numbers the reference holds after ITS readers (checked bit for
 * bit against a dump of the reference's structs, tests/test_files.py).  Plain C ABI; host only.
 */
#ifndef VICRES_FILES_H
#define VICRES_FILES_H
#include "vicres.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct vr_param_files {
  int32_t nlayer;            /* NLAYER */
  int32_t nbands;            /* SNOW_BAND */
  int32_t root_zones;        /* ROOT_ZONES */
  int32_t vegparam_lai;      /* VEGPARAM_LAI TRUE: every tile line is followed by 12 monthly LAI values */
  int32_t lai_from_vegparam; /* LAI_SRC FROM_VEGPARAM */
  const char *soil, *veglib, *vegparam;
  const char *snowband;      /* may be NULL when nbands == 1 */
  int32_t baseflow_nijssen;  /* BASEFLOW NIJSSEN2001: soil columns 5-8 are d1..d4 of Nijssen et al. (2001), converted to the ARNO
                              * Ds, Dsmax, Ws as read_soilparam.c:719-726 does (c stays) */
  int32_t organic_fract;     /* ORGANIC_FRACT TRUE: after the mineral soil densities the soil file holds, per layer, the organic
                              * fraction and the organic bulk and soil densities (read_soilparam.c:444-506, 647-655) */
} vr_param_files;

typedef struct vr_params vr_params;          /* owns every array the structs below point to */

int  vr_params_read(const vr_param_files *f, vr_params **out);
void vr_params_free(vr_params *p);
const char *vr_params_last_error(void);
const vr_cells  *vr_params_cells(const vr_params *p);     /* cells flagged "run" (first column) in file order */
const vr_veglib *vr_params_veglib(const vr_params *p);
const int32_t *vr_params_gridcel(const vr_params *p);     /* [C] cell numbers            */
const float   *vr_params_lat(const vr_params *p);         /* [C] as stored (float)       */
const float   *vr_params_lng(const vr_params *p);
/* [C] site values initialize_atmos reads, as soil_con holds them (floats widened): which = 0 lat, 1 lng,
 * 2 time_zone_lng (off_gmt*360/24, read_soilparam.c:934), 3 elevation, 4 annual_prec, 5 the lowest Tfactor of the cell's
 * snow bands -> vr_atmos_cells */
const double  *vr_params_site(const vr_params *p, int32_t which);

/* The reference's ASCII result files (write_data.c:196-226, names from make_in_and_outfiles.c:46-86):
 * one file "<result_dir>/<prefix>_<lat>_<lng>" per cell (lat/lng printed with grid_decimal places),
 * one row per record: "%04i\t%02i\t%02i\t" (plus "%02i\t" hour when hour != NULL, i.e. dt < 24) and
 * the selected columns in "%.4f" separated by "\t ".  The default lists (set_output_defaults.c:66-93)
 * are prefix "fluxes" = the first 20+nlayer-1 default outputs and prefix "snow" = the last 4.
 * out: [n_total][nrec][C] as vr_step_block returns it; cols: [n_cols] rows of `out` to print. */
int  vr_write_results(const char *result_dir, const char *prefix, int32_t grid_decimal, int64_t ncell,
                      const float *lat, const float *lng, int32_t nrec, const int32_t *year, const int32_t *month,
                      const int32_t *day, const int32_t *hour, int32_t n_cols, const int32_t *cols,
                      const double *out);

/* ---- output selection, aggregation and the result writer in full (parse_output_info.c:55-125, set_output_defaults.c,
 * put_data.c:665-790, write_data.c) ------------------------------------------------------------------------------ */
#define VR_MAX_OUTFILES 8
#define VR_MAX_OUTCOLS 64
enum { VR_OUT_TYPE_DEFAULT = 0, VR_OUT_TYPE_CHAR, VR_OUT_TYPE_SINT, VR_OUT_TYPE_USINT, VR_OUT_TYPE_INT, VR_OUT_TYPE_FLOAT,
       VR_OUT_TYPE_DOUBLE };                       /* vicNl_def.h: OUT_TYPE_* */
/* The files a run writes and their columns.  A column is one ELEMENT of a reference variable (OUT_SOIL_LIQ is nlayer
 * columns); format / type / mult belong to the variable as in the reference (out_data[varid]). */
typedef struct vr_outsel {
  int32_t nfiles;
  char    prefix[VR_MAX_OUTFILES][64];
  int32_t ncols[VR_MAX_OUTFILES];
  int32_t var[VR_MAX_OUTFILES][VR_MAX_OUTCOLS];          /* VR_OUT_* id of the column                          */
  char    format[VR_MAX_OUTFILES][VR_MAX_OUTCOLS][12];   /* printf format of the ASCII column ("%.4f")          */
  int32_t type[VR_MAX_OUTFILES][VR_MAX_OUTCOLS];         /* binary type (OUT_TYPE_FLOAT unless the OUTVAR line says otherwise) */
  double  mult[VR_MAX_OUTFILES][VR_MAX_OUTCOLS];         /* binary multiplier                                   */
  /* PRT_HEADER (write_header.c): the column's name as the header prints it (the reference's variable name, "_<element>"
   * appended when the variable has several), the number of VARIABLES of the file, and what the caller fills in before
   * vr_write_results_sel when the files are to carry the header: header != 0, the run's record count and OUT_STEP in hours,
   * the date of the FIRST MODEL RECORD (year, month, day, hour), ALMA_OUTPUT */
  char    name[VR_MAX_OUTFILES][VR_MAX_OUTCOLS][32];
  int32_t nvars[VR_MAX_OUTFILES];
  int32_t header, hdr_nrecs, hdr_out_dt, hdr_start[4], hdr_alma;
  int32_t no_date;          /* OUTPUT_FORCE TRUE: the files hold the forcing, one row per snow step, without date columns
                             * (write_data.c:118-131, 206-217); set by vr_outsel_read */
} vr_outsel;
/* N_OUTFILES / OUTFILE / OUTVAR lines of a global parameter file; without them the default pair "fluxes" (20 variables,
 * 20 + nlayer - 1 columns) and "snow" (4) of set_output_defaults.c:66-162.  An unknown variable name is an error, as in
 * the reference (set_output_var). */
int  vr_outsel_read(const char *global_path, int32_t nlayer, vr_outsel *out);
/* the VR_OUT_* ids the step must produce for a selection, each once, in order of first appearance: vars[VR_MAX_OUT].
 * mode bit 0 (OUT_STEP > TIME_STEP): OUT_AERO_RESIST is produced as OUT_AERO_COND, because the reference aggregates the
 * conductance and inverts the aggregate (put_data.c:686-688); mode bit 1 (ALMA_OUTPUT): OUT_SUB_CANOP is produced whenever
 * OUT_SUB_SNOW is asked for, because the ALMA value of the latter includes the former (put_data.c:742). */
int  vr_outsel_step_vars(const vr_outsel *s, int32_t mode, int32_t *vars, int32_t *n);
/* Temporal aggregation of put_data.c:665-685 on the device: in_dev [n][nrec][ncell] (device) -> out_dev [n][nrec/ratio][ncell]
 * (device; may alias nothing), ratio = OUT_STEP / TIME_STEP records per output interval, vars[n] = the VR_OUT_* ids of the
 * rows (a row that holds AERO_COND for a requested AERO_RESIST is passed as VR_OUT_AERO_RESIST with cond_for_resist != 0):
 * fluxes are summed, storages take the interval's last record, the rest is averaged as sum of x / ratio, in record order. */
int  vr_aggregate_device(int device, int32_t n, const int32_t *vars, int32_t cond_for_resist, int32_t nrec, int32_t ratio,
                         int64_t ncell, const double *in_dev, double *out_dev, void *cuda_stream);
/* ALMA_OUTPUT (put_data.c:697-760), in place on [n][nrec][ncell] device values that are already aggregated to OUT_STEP =
 * out_dt_hours: moisture fluxes become rates (/ seconds of the interval), temperatures Kelvin, OUT_DELTAH * seconds, and
 * OUT_SUB_SNOW gets OUT_SUB_CANOP added (which therefore has to be one of the rows). */
int  vr_alma_output_device(int device, int32_t n, const int32_t *vars, int32_t nrec, int64_t ncell, int32_t out_dt_hours,
                           double *data_dev, void *cuda_stream);
/* OUTPUT_FORCE (write_forcing_file.c:46-85): the rows of the forcing files from the prepared forcing.  fields_dev: the nine
 * VR_F_* arrays as vr_atmos_prepare_sub_device leaves them ([nrecs][NF + 1][C], or [nrecs][C] when nf == 1); out_dev:
 * [n][nrecs * nf][C], row rec * nf + j = window j of record rec.  vars: VR_OUT_AIR_TEMP, DENSITY, LONGWAVE, PREC, PRESSURE,
 * QAIR, REL_HUMID, SHORTWAVE, VP, VPD, WIND, RAINF, SNOWF (the rain / snow split by the two temperature thresholds);
 * anything else is VR_ERR_UNSUPPORTED.  vr_alma_output_device with out_dt_hours = TIME_STEP does its ALMA units. */
int  vr_forcing_outputs_device(int device, int32_t n, const int32_t *vars, int32_t nrecs, int32_t nf, int64_t ncell,
                               double max_snow_temp, double min_rain_temp, const double *const fields_dev[9], double *out_dev,
                               void *cuda_stream);
/* the same with host buffers: fields [9] each [nrecs * (nf > 1 ? nf + 1 : 1) * ncell], out [n][nrecs * nf][ncell] */
int  vr_forcing_outputs(int device, int32_t n, const int32_t *vars, int32_t nrecs, int32_t nf, int64_t ncell, double max_snow_temp,
                        double min_rain_temp, const double *const fields[9], double *out);
/* the same on a host buffer (copy in, kernel, copy out) */
int  vr_alma_output(int device, int32_t n, const int32_t *vars, int32_t nrec, int64_t ncell, int32_t out_dt_hours, double *data);
/* the same with host buffers (copies in, kernel, copies out) */
int  vr_aggregate(int device, int32_t n, const int32_t *vars, int32_t cond_for_resist, int32_t nrec, int32_t ratio, int64_t ncell,
                  const double *in, double *out);
/* File `file` of a selection for every cell: ASCII (write_data.c:196-226: date, then the columns in their formats separated by
 * "\t ") or binary (write_data.c:103-193: 3 or 4 ints of date, then every column value * mult cast to its type).  hour == NULL
 * when the output interval is a day or longer.  rows[c] = row of `out` ([.][nrec][ncell]) that holds column c of the file. */
int  vr_write_results_sel(const char *result_dir, const vr_outsel *s, int32_t file, int32_t grid_decimal, int64_t ncell,
                          const float *lat, const float *lng, int32_t nrec, const int32_t *year, const int32_t *month,
                          const int32_t *day, const int32_t *hour, int32_t binary, const int32_t *rows, const double *out);

/* ---- global parameter file (get_global_param.c:210-727) and calendar (make_dmy.c) ---------------------- */
#define VR_MAX_FORCE_TYPES 16
typedef struct vr_global {
  /* what vr_options needs */
  int32_t nlayer, time_step, snow_step;            /* NLAYER, TIME_STEP, SNOW_STEP                              */
  double  max_snow_temp, min_rain_temp, wind_h;    /* MAX_SNOW_TEMP, MIN_RAIN_TEMP, WIND_H                      */
  double  min_wind_speed, measure_h;               /* MIN_WIND_SPEED, MEASURE_H (used by the forcing preparation) */
  /* simulated period: nrecs is computed from the END* date when NRECS is absent (make_dmy.c:52-90) */
  int32_t startyear, startmonth, startday, starthour, nrecs, skipyear;
  /* parameter files, as vr_param_files wants them */
  int32_t snow_band, root_zones, vegparam_lai, lai_from_vegparam, grid_decimal;
  char soil[512], veglib[512], vegparam[512], snowband[512], result_dir[512];
  /* forcing description of FORCING1 (first file type only) */
  char forcing1[512];
  int32_t force_ascii, force_little_endian, force_dt, n_force_types;
  int32_t forceyear, forcemonth, forceday, forcehour;
  char force_type[VR_MAX_FORCE_TYPES][16];         /* FORCE_TYPE names in column order (SKIP included)         */
  int32_t force_signed[VR_MAX_FORCE_TYPES];        /* binary files: SIGNED / UNSIGNED                           */
  double  force_multiplier[VR_MAX_FORCE_TYPES];    /* binary files: value = short / multiplier                  */
  int32_t binary_output, out_step, compress;
  /* options this library does not implement are not an error to PARSE; vr_global_unsupported() names them */
  int32_t full_energy, frozen_soil, lakes, carbon, blowing, corrprec, compute_treeline, dist_prcp, organic_fract,
          july_tavg_supplied, alma_input, init_state;
  /* options of the forcing preparation (include/vicres_atmos.h); values as vicNl_def.h:208-223 */
  int32_t vp_interp, vp_iter, mtclim_swe_corr, lw_type, lw_cloud, plapse;
  double  sw_prec_thresh;
  /* state files (INIT_STATE <file>; STATENAME <prefix> + STATEYEAR/STATEMONTH/STATEDAY: the file written at the end of
   * that day is "<prefix>_YYYYMMDD", get_global_param.c:961-964; BINARY_STATE_FILE) -> vr_read/write_state_file */
  char    init_state_file[512], statename[512];
  int32_t stateyear, statemonth, stateday, binary_state_file;
  /* a second forcing file (FORCING2 <prefix>: != 0, read through vr_global_read_forcing2) and a custom output selection (N_OUTFILES / OUTFILE / OUTVAR lines,
   * parse_output_info.c:62-118; counted here, parsed by vr_outsel_read) */
  int32_t forcing2, n_outfile_lines, n_outvar_lines;
  /* the first option of the file that changes the results and has a value this library does not implement ("KEY VALUE":
   * PRT_SNOW_BAND, SNOW_ALBEDO, SNOW_DENSITY,
   * AERO_RESIST_CANSNOW, GRND_FLUX_TYPE, SPATIAL_SNOW ...), or empty: named by vr_global_unsupported, never ignored */
  char    other_unsupported[96];
  int32_t prt_header, alma_output, output_force;   /* PRT_HEADER, ALMA_OUTPUT, OUTPUT_FORCE                     */
  int32_t moistfract;                              /* MOISTFRACT -> vr_options.moistfract                       */
  int32_t baseflow_nijssen;                        /* BASEFLOW NIJSSEN2001 -> vr_param_files.baseflow_nijssen   */
} vr_global;

int  vr_global_read(const char *path, vr_global *out);
/* FORCING2: the same global file read with the description of the SECOND forcing file in the forcing fields (forcing1 = its
 * path prefix, force_type[], force_dt, format, FORCEYEAR ...), so that vr_force_skip / vr_read_forcing_file read that file;
 * everything else equals what vr_global_read returns.  A type both files hold is taken from the second, as the reference's
 * param_set.TYPE[].SUPPLIED ends up pointing at it.  Both files have to share FORCE_DT here (the hosts check). */
int  vr_global_read_forcing2(const char *path, vr_global *out2);
/* NULL when the water-balance path of this library covers the file's options, else a text naming the first
 * option it does not (FULL_ENERGY TRUE, FROZEN_SOIL TRUE, SNOW_STEP != TIME_STEP, ...) */
const char *vr_global_unsupported(const vr_global *g);
/* calendar of the nrecs model records (make_dmy.c:92-120); each array [g->nrecs] */
int  vr_make_dmy(const vr_global *g, int32_t *year, int32_t *month, int32_t *day, int32_t *hour, int32_t *day_in_year);
/* records to skip at the head of the forcing file (make_dmy.c:122-150) */
int32_t vr_force_skip(const vr_global *g);
/* SKIPYEAR: the model records before the first one written (make_dmy.c:162-169: per skipped year the 365 or 366 days
 * of the year the running offset has reached; year[] from vr_make_dmy).  put_data.c:765 writes an output interval only
 * when its last record is >= this.  An offset beyond the run (which the reference reads past its array) counts 365 days. */
int32_t vr_output_skip(const vr_global *g, const int32_t *year);
/* one cell's forcing file "<forcing1><lat>_<lng>" (read_atmos_data.c:120-235): columns in FORCE_TYPE order, ASCII or
 * 2-byte scaled binary (with or without the 0xFFFF header); data: [n_force_types][nrec_forcing] (SKIP columns are
 * read and kept), nrec_forcing = nrecs * time_step / force_dt.  Returns the records read or a negative status. */
int64_t vr_read_forcing_file(const vr_global *g, const char *path, int64_t nrec_forcing, double *data);

/* out[i] = the value in[i] has after a round trip through a "%.<decimals>f" text column (the reference's flux
 * files are the hand-off between vicNl and rout): lets the step's outputs go to the routing pass in memory and
 * still equal what rout would have read. */
int  vr_round_decimals(int64_t n, const double *in, double *out, int32_t decimals);

#ifdef __cplusplus
}
#endif
#endif
