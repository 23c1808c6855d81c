/* vicres_route.h -- C ABI of the B200-native VIC-Res routing pass (unit hydrographs + convolution).
 *
 * Replaces, for one station, the reference's
 *   MAKE_UHM            routing/model/unit_hyd_routines.f:7-50     per-cell 48-tap Green's function
 *   SEARCH_WHOLECATCHMENT / SEARCH_CATCHMENTRS   init_routines.f:103-168, reservoir.f:578-763
 *                       catchment cell lists of the station and of every reservoir above it
 *   MAKE_GRID_UH        unit_hyd_routines.f:55-178                per-cell UH_S(107) to its node
 *   MAKE_CONVOLUTION    reservoir.f:216-526                       node inflow = sum over cells of UH_S * (runoff+baseflow)
 * (the reference drives them from rout.f:318-478 and reservoir.f:769-1140).  All arithmetic is fp32,
 * as in the Fortran (REAL), with the reference's accumulation order: cells in catchment-list order,
 * taps ascending.
 *   MAKE_RESERVOIR_UH   unit_hyd_routines.f:218-346               UH_R(107) reservoir -> next reservoir / station
 *   MAKE_CONVOLUTIONRS  reservoir.f:1142-1684                     the reservoir day loop (rules 0-6, spill,
 *                       environmental flow, irrigation, seepage, head, energy, release routing through UH_R)
 * Open-water (Penman 1948, table driven) evaporation on reservoir-surface cells (RESER == 9999,
 * reservoir.f:410-475,496-503) is part of vr_route_convolve_in; the two shorter convolve calls reject grids
 * with 9999 cells inside a catchment with VR_ERR_UNSUPPORTED because they lack the meteorology it needs.
 *
 * Grid arrays are in FILE order: index r*ncol + c where r = 0 is the FIRST data line of the ESRI
 * ASCII file (northern-most row) and c = 0 its first value.  Station col/row are the numbers of
 * stations.txt (rout.f:333-341).
 */
#ifndef VICRES_ROUTE_H
#define VICRES_ROUTE_H
#include <stdint.h>
#include "vicres.h"

#ifdef __cplusplus
extern "C" {
#endif

#define VR_UH_LE 48        /* LE, hourly taps of the cell impulse response (rout.f:31-40) */
#define VR_UH_DAY 96       /* UH_DAY */
#define VR_UH_KE 12        /* KE, taps of UH_BOX (UH.all) */
#define VR_UH_NS (VR_UH_KE + VR_UH_DAY - 1)   /* 107 daily taps of UH_S */

typedef struct vr_route_grid {
  int32_t ncol, nrow;
  const int32_t *dir;     /* D8 codes 0..8 (READ_DIREC, reservoir.f:149-211) */
  const int32_t *reser;   /* reservoirlocation.txt: 0, reservoir id, 9999 = reservoir surface; may be NULL */
  const float *velo, *diff, *xmask;   /* per cell (a constant in configuration.txt is a filled grid) */
} vr_route_grid;

typedef struct vr_route vr_route;

int  vr_route_create(const vr_route_grid *g, int32_t station_col, int32_t station_row,
                     const float *uh_box /* [VR_UH_KE] */, int device, vr_route **out);
void vr_route_destroy(vr_route *h);
const char *vr_route_last_error(const vr_route *h);

/* nodes: the reservoirs found above the station, in the reference's discovery order, then the station */
int32_t vr_route_num_nodes(const vr_route *h);
int  vr_route_node_info(const vr_route *h, int32_t node, int32_t *res_id, int64_t *ncells);
/* catchment cells of a node in the reference's list order, as file-order grid indices r*ncol+c */
int  vr_route_node_cells(const vr_route *h, int32_t node, int32_t *grid_index);
/* UH_S of a node: [ncells][VR_UH_NS] (what the reference writes to <name>.uh_s) */
int  vr_route_get_uh(vr_route *h, int32_t node, float *uh_s);

/* flows[node*ndays + i] (m3/s): runoff, baseflow are [nrow*ncol][ndays] mm/day by file-order grid
 * cell (cells without data hold zeros, as the reference inserts for missing flux files); `scale` is
 * the mm/day -> m3/s factor (the reference hard-codes 0.18308, reservoir.f:479-480). Host pointers. */
int  vr_route_convolve(vr_route *h, int32_t ndays, const float *runoff, const float *baseflow, float scale,
                       float *flows);
/* Full form of MAKE_CONVOLUTION's per-cell handling:
 *   cell_scale[ncell]   per-cell mm/day -> m3/s factor instead of `scale` (the shipped prebuilt binary used
 *                       FACTOR*0.028 with FACTOR = FRACTION*35.315*AREA/86400000, reservoir.f:357-364)
 *   ev_vege, ev_soil, tr_vege [ncell][ndays] + cell_factor[ncell]: the reference's correction on days whose VIC
 *                       evaporation components sum below zero (reservoir.f:496-499)
 *   present[ncell]      0 = flux file missing: zeros are routed and the correction sees the PREVIOUS cell's
 *                       evaporation series (the reference resets only runoff/baseflow, reservoir.f:398-409)
 * Any of them may be NULL. */
int  vr_route_convolve_ex(vr_route *h, int32_t ndays, const float *runoff, const float *baseflow, float scale,
                          const float *cell_scale, const float *ev_vege, const float *ev_soil, const float *tr_vege,
                          const float *cell_factor, const unsigned char *present, float *flows);
/* Everything MAKE_CONVOLUTION reads, including the columns of the flux files that drive the open-water
 * evaporation of reservoir-surface cells.  Unused pointers may be NULL (then as vr_route_convolve_ex). */
typedef struct vr_route_inputs {
  const float *runoff, *baseflow;          /* [ncell][ndays] mm/day                                        */
  float scale;                             /* mm/day -> m3/s when cell_scale == NULL                       */
  const float *cell_scale;                 /* [ncell]                                                      */
  const float *ev_vege, *ev_soil, *tr_vege;/* [ncell][ndays]                                               */
  const float *cell_factor;                /* [ncell] FACTOR                                               */
  const unsigned char *present;            /* [ncell] flux file found                                      */
  const float *rh, *tair, *wind;           /* [ncell][ndays] RHD (%), AIRTEMP (C), WINDS (m/s)             */
  const float *cell_lat, *cell_lon;        /* [ncell] JLOC, ILOC (reservoir.f:353-356)                     */
  const int32_t *node_ope_year;            /* [nnodes] OPEYEAR of the node's reservoir (reservoir.f:318-325); station 0 */
  int32_t start_year, start_month;         /* START_YEAR, START_MO                                         */
  int32_t node_begin, node_end;            /* only nodes [node_begin, node_end) are computed (the other rows of `flows` are left
                                              untouched); 0, 0 = all.  Nodes are independent in MAKE_CONVOLUTION, so ranks of a
                                              multi-GPU run take node ranges and all-gather the rows: no change of summation order */
  int32_t reserved[2];
} vr_route_inputs;
/* The flux hand-off between the two passes on the device: rows `rows[v]` of the cell step's output block
 * out_dev [.][ndays][ncell] (device, as vr_step_block_device leaves it) -> the routing pass's columns dst_dev[v]
 * [ngrid][ndays] float (device), cell c going to grid cell grid_of_cell[c] (host array; -1 = not on the routing grid), every
 * value through the "%.4f" text round trip of the flux files (write_data.c -> READ(20,*), reservoir.f:373-389) exactly as
 * vr_round_decimals + the conversion to REAL do on the host.  present_dev [ngrid] (device, may be NULL) is set for the grid
 * cells that received a VIC cell.  The input arrays of vr_route_inputs / vr_route_convolve_ex may be host OR device memory. */
int  vr_route_fluxes_from_step(int device, const double *out_dev, int64_t ncell, int32_t ndays, int32_t nvar, const int32_t *rows,
                               const int64_t *grid_of_cell, int64_t ngrid, float *const dst_dev[], unsigned char *present_dev);
/* res_evaporation: [nnodes][ndays] RES_EVAPORATION (sum of 0.9*E0 over the node's surface cells) or NULL */
int  vr_route_convolve_in(vr_route *h, int32_t ndays, const vr_route_inputs *in, float *flows, float *res_evaporation);
int  vr_route_last_kernel_ms(vr_route *h, float *ms);

/* ---- reservoirs (MAKE_CONVOLUTIONRS) -------------------------------------------------------------- */

/* One resN.txt (reservoir.f:1021-1131), units as in the file (volumes 1000 m3, flows m3/s, levels m). */
typedef struct vr_reservoir {
  float hresermax, hresermin;          /* HRESERMAX, HRESERMIN                                     */
  float v_live, v_dead;                /* VRESERTHAT, VDEAD                                        */
  float head, q_design;                /* HYDRAUHEAD, QRESERTHAT                                   */
  int32_t ope_year;                    /* OPEYEAR                                                  */
  float v_initial;                     /* VINITIAL                                                 */
  float seepage, infiltration;
  int32_t hydropower_type;             /* 0 run-of-river, 1 storage                                */
  int32_t hydropower_generator;        /* 1 generators on                                          */
  float env_flow;                      /* fraction of q_design                                     */
  int32_t rule;                        /* 0..6                                                     */
  float hmax, hmin, op1[2];            /* rule 1: levels and the day numbers T1, T2                */
  float reslv[12];                     /* rule 2: monthly target level                             */
  float demand, x1, x2, x3, x4;        /* rule 3 (x1, x4 in radians)                               */
  float demand5[12], op5x1[12], op5x2[12], op5x3[12], op5x4[12];   /* rule 5                       */
  int32_t bathymetry;                  /* BTHMET_ACTIVE                                            */
  float bth[4];                        /* BTHMET_A..D                                              */
  const float *series;                 /* rule 4 release / rule 6 storage series [ndays], already aligned to the
                                          first simulated day (the reference's STARTDAY offset); else NULL */
  const float *irrigation;             /* [ndays] withdrawals (WITHIRRIGATION = 1) or NULL         */
} vr_reservoir;

typedef struct vr_res_options {
  int32_t start_year, start_month;     /* START_YEAR, START_MO (CURRENTYEAR / CRTDATE, reservoir.f:1154,1172) */
  int32_t legacy_rule_curve;           /* 0: today's reservoir.f (80 % lower/upper bands, :1222-1265);
                                          1: the zone logic the shipped prebuilt `rout` still runs (kept as
                                          comments in reservoir.f:1268-1302) -- the variant parity is pinned on */
  int32_t reserved[5];
} vr_res_options;

/* Per-day series of every reservoir; each pointer may be NULL. Layout [set][reservoir][ndays (+1)]. */
typedef struct vr_res_series {
  float *vol;        /* [nsets][nres][ndays+1]  VOL(J,1..NDAY+1)           */
  float *hho;        /* [nsets][nres][ndays+1]  HHO                        */
  float *flowin;     /* [nsets][nres][ndays]    FLOWIN                     */
  float *flowout;    /* [nsets][nres][ndays]    FLOWOUT                    */
  float *turbine;    /* [nsets][nres][ndays]    FLOWOUT_TURB               */
  float *energy;     /* [nsets][nres][ndays]    ENERGYPRO                  */
  float *htk;        /* [nsets][nres][ndays]    HTK (target level)         */
} vr_res_series;

/* node -> RES_DIRECT(:,2) (id of the next reservoir downstream, 0 = none) and the node its release is routed
 * to (reservoir.f:1529-1567: the next reservoir when it lies above this station, else the station) */
int  vr_route_res_info(const vr_route *h, int32_t node, int32_t *downstream_id, int32_t *target_node);
/* UH_R of a reservoir node: [VR_UH_NS] (what the reference writes to RES<i>_<to>.uh_r) */
int  vr_route_get_uh_r(vr_route *h, int32_t node, float *uh_r);
/* replace it with a user-supplied one, as the reference does when it reads an existing .uh_r file
 * (reservoir.f:1544-1554) */
int  vr_route_set_uh_r(vr_route *h, int32_t node, const float *uh_r);
/* The day loop for nsets independent parameter sets (toolbox/optimization evaluates many rule sets on the
 * same naturalised inflows).  natural: [nnodes][ndays] from vr_route_convolve; res: [nsets][nnodes-1] in node
 * order; flow: [nsets][ndays] regulated station discharge.  Host pointers. */
int  vr_route_reservoirs(vr_route *h, int32_t ndays, int32_t nsets, const float *natural, const vr_reservoir *res,
                         const vr_res_options *opt, float *flow, const vr_res_series *out);

/* Days [day_begin, day_end) of the same day loop (0-based): the reference's STEPBYSTEP mode (reservoir.f:1142-1150,
 * 1596-1675: one day per run of `rout`, VOL / FLOWOUT / RESFLOWS carried from run to run in files).  With day_begin > 0
 * `out` carries the state IN as well: out->vol[.][day_begin] and out->flowout[.][0 .. day_begin) must hold what the
 * previous call left (the other series of `out`, when given, are kept for the earlier days).  `flow` is valid for the
 * days < day_end.  Calling it for [0, d), [d, ndays) gives bit for bit what vr_route_reservoirs gives for [0, ndays). */
int  vr_route_reservoirs_range(vr_route *h, int32_t ndays, int32_t nsets, int32_t day_begin, int32_t day_end, const float *natural,
                               const vr_reservoir *res, const vr_res_options *opt, float *flow, const vr_res_series *out);

/* ---- calibration objectives (toolbox/calibration/basin_calibration_eNSGAII.py:86-175) ---------------------
 * out[set][4] = NSE, TRMSE (Box-Cox 0.3), MSDE, ROCE of sim[set][ndays] (station discharge of every parameter
 * set, e.g. the `flow` of vr_route_reservoirs) against obs[ndays]; handle_negatives = the toolbox's abs(). */
int  vr_objectives(int device, int32_t nsets, int32_t ndays, const float *sim, const double *obs,
                   int32_t handle_negatives, double *out);

#ifdef __cplusplus
}
#endif
#endif
